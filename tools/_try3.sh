# Development sweep (2 GPUs under gpurun): z-split H psi, in-place z columns vs pushed column buffers x tile configurations;
# produced gpurun_out/r02_try3.log, summarised in profiles/r02_decomposition.md.
run() { echo "== decomp $1 zmode $2 cfg $3"; MGB_HPSI_CFG=$3 MGB_HPSI_TIMING=1 MGB_ZHALO=$2 MGB_BENCH_QUICK=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --decomp $1 --no-pieces 2>&1 | grep -E "mgb timing|ms_per_step" | grep -v "rank 1" | sed -E 's/.*"ms_per_step": ([0-9.]+).*"kernel": "([^"]*)".*/\1 \2/' | cut -c1-200; }
run 1x1x2 inplace 8,1,2,6,128
run 1x1x2 inplace 8,1,2,6,0
run 1x1x2 inplace 8,1,1,8,128
run 1x1x2 inplace 8,1,3,4,128
run 1x1x2 inplace 4,2,2,6,128
run 1x1x2 inplace 4,1,4,8,128
run 1x1x2 push 8,1,2,6,128
