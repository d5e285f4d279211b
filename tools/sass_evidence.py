"""Static evidence of what the kernels are made of: per kernel family of
mgmol_b200/libmgmol_b200.so, how many template instantiations exist and the
range of SASS instruction counts for TMA loads/stores (UTMALDG / UTMASTG /
UBLKCP), mbarrier operations (SYNCS), FP64 tensor MMA (DMMA), TF32 tensor MMA
(HMMA...TF32), cp.async (LDGSTS) and FP64/FP32 FMAs.  Runs without a GPU:

    python tools/sass_evidence.py > profiles/r01_sass_evidence.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mgmol_b200", "libmgmol_b200.so")
COLS = [("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("UBLKCP", r"\bUBLKCP"),
        ("SYNCS", r"\bSYNCS"), ("UTCHMMA", r"\bUTCHMMA"), ("UTCBAR", r"\bUTCBAR"),
        ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("DMMA", r"\bDMMA"),
        ("HMMA.TF32", r"\bHMMA[.\w]*TF32"),
        ("LDGSTS", r"\bLDGSTS"), ("DFMA", r"\bDFMA"), ("FFMA", r"\bFFMA"),
        ("instr", r"/\*[0-9a-f]{4}\*/")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    parts = re.split(r"\n\s*Function : ", sass)[1:]
    names = [p.split("\n", 1)[0].strip() for p in parts]
    dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True,
                         text=True).stdout.splitlines()
    # registers / static shared memory / local-memory stack per kernel
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True,
                         text=True).stdout
    usage = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", res):
        usage[m.group(1)] = tuple(int(m.group(i)) for i in (2, 3, 4, 5))
    fam = collections.OrderedDict()
    famres = collections.OrderedDict()
    for raw, name, body in zip(names, dem, parts):
        base = re.sub(r"^void\s+", "", name)
        base = re.sub(r"[<(].*", "", base).replace("mgb::", "")
        if base.startswith("(anonymous namespace)::"):
            base = base[len("(anonymous namespace)::"):]
        counts = [len(re.findall(pat, body)) for _, pat in COLS]
        fam.setdefault(base, []).append(counts)
        famres.setdefault(base, []).append(usage.get(raw, (0, 0, 0, 0)))
    print("# SASS make-up of the kernels in `mgmol_b200/libmgmol_b200.so` (sm_100a)\n")
    print("Produced by `python tools/sass_evidence.py` from `cuobjdump -sass` (no GPU needed). One row per "
          "kernel family; `n` = template instantiations; every other cell = min–max count of that "
          "instruction over the instantiations. UTMALDG/UTMASTG = TMA tensor loads/stores "
          "(`cp.async.bulk.tensor`), SYNCS = mbarrier operations, UTCHMMA = `tcgen05.mma` (5th-generation "
          "tensor cores; `kind::tf32` here), UTCBAR = `tcgen05.commit`, LDTM / STTM = `tcgen05.ld` / "
          "`tcgen05.st` (TMEM), DMMA = FP64 tensor MMA, "
          "HMMA…TF32 = TF32 tensor MMA (`mma.sync.m16n8k8.tf32`, the 3×TF32 float contractions), "
          "LDGSTS = `cp.async`; regs / stack (local memory, i.e. spills when non-zero) / static shared "
          "memory from `cuobjdump --dump-resource-usage` (the TMA rings are dynamic shared memory and do "
          "not show here).\n")
    print("| kernel | n | " + " | ".join(c for c, _ in COLS) + " | regs | stack B | static smem B |")
    print("|---|---|" + "---|" * (len(COLS) + 3))

    def rng(v):
        return str(min(v)) if min(v) == max(v) else "%d–%d" % (min(v), max(v))
    for base, rows in sorted(fam.items(), key=lambda kv: kv[0]):
        cells = [rng([r[i] for r in rows]) for i in range(len(COLS))]
        rr = famres[base]
        cells += [rng([r[0] for r in rr]), rng([r[1] + r[3] for r in rr]), rng([r[2] for r in rr])]
        print("| `%s` | %d | %s |" % (base, len(rows), " | ".join(cells)))
    tot = len(parts)
    print("\n%d kernels in all.\n" % tot)
    print("Reading the table: the two streaming stencil families (`k_hpsi_tma`, `k_mg_jacobi`) are TMA + "
          "mbarrier pipelines; the `true` PEER instantiations of `k_hpsi_tma` carry the extra tensor maps "
          "of the neighbours' blocks (93 vs 48 UTMALDG). The contractions are tensor-pipe kernels fed by "
          "`cp.async` rings: DMMA for `ORBDTYPE double` (FP64 has no tcgen05 form; DMMA is the FP64 tensor "
          "instruction of sm_100a). For `ORBDTYPE float` the error-compensated 3xTF32 contractions run "
          "on tcgen05 (`k_gemm_tn_umma`, `k_gemm_nn_umma`: UTCHMMA with the A operand in TMEM, UTCBAR "
          "commits, LDTM drains, STTM operand stores, TMA-fed); the `mma.sync` TF32 kernels "
          "(`k_gemm_*_tf32`) remain as `mgb_set_f32_contraction(2)`. The "
          "\"literal\" kernels (`k_del2_*`, `k_rhs_*`, `k_axpy`, `k_hpsi_generic` ...) are compiled with "
          "`-fmad=false` on purpose: they reproduce the reference's separately rounded multiply and add, "
          "hence no FMA in them.")


if __name__ == "__main__":
    sys.exit(main())
