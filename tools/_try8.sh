# Development sweep (8 GPUs under gpurun): decompositions 2x2x2 vs 4x2x1 of the 256^3 x 4096 block;
# summarised in profiles/r02_decomposition.md.
run() { echo "== decomp $1 cfg $2"; MGB_BENCH_QUICK=1 MGB_HPSI_CFG=$2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --decomp $1 $3 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith(chr(123)):
        d=json.loads(l); print(d['ms_per_step'], d['roofline']['frac'], d['path']['kernel']); p=d.get('pieces') or {}; print({k:v.get('ms') for k,v in p.items() if isinstance(v,dict)})"; }
run 2x2x2 "" --no-pieces
run 4x2x1 "" --no-pieces
run 2x2x2 8,2,2,3,0 --no-pieces
