"""Build libmgmol_b200.so in-tree with nvcc for sm_100a (no JIT cache: the built
.so travels to the GPU box with the repo snapshot).

    python -m mgmol_b200.build [--force] [--verbose]
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libmgmol_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
          "-Xcompiler", "-fvisibility=default", "-I", os.path.join(ROOT, "include"),
          "-I", CSRC]

# translation unit -> extra flags.  The "literal" units reproduce the
# reference's rounding bit for bit and must not contract a*b+c into an FMA.
SOURCES = {
    "api.cu": [],
    "fd_ghosted.cu": ["-fmad=false"],
    "hpsi_generic.cu": ["-fmad=false"],
    "hpsi_fused.cu": [],
    "mg_precond.cu": [],
    "mg_fused.cu": [],
    "masks.cu": [],
    "blas1_cols.cu": [],
    "contractions.cu": [],
    "comm.cu": [],
    "poisson.cu": [],
    # the scatter reproduces the reference's (T) roundings: no FMA contraction
    "kb_nonlocal.cu": ["-fmad=false"],
}


def _stamp(src, flags):
    h = hashlib.sha256()
    h.update(" ".join(flags).encode())
    with open(src, "rb") as f:
        h.update(f.read())
    for name in sorted(os.listdir(CSRC)):
        if name.endswith((".h", ".cuh")):
            with open(os.path.join(CSRC, name), "rb") as f:
                h.update(f.read())
    # the C ABI header for every unit; the C++ headers for the unit that
    # instantiates their templates
    incs = ["mgmol_b200.h"]
    if os.path.basename(src) == "poisson.cu":
        incs += ["mgmol_b200.hpp", "mgmol_b200_poisson.hpp"]
    for name in incs:
        with open(os.path.join(ROOT, "include", name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _compile(name, flags, verbose, force):
    src = os.path.join(CSRC, name)
    obj = os.path.join(OBJ, name.replace(".cu", ".o"))
    stamp_file = obj + ".stamp"
    allflags = ARCH + COMMON + flags
    stamp = _stamp(src, allflags)
    if (not force and os.path.exists(obj) and os.path.exists(stamp_file)
            and open(stamp_file).read() == stamp):
        return obj, False
    cmd = [NVCC] + allflags + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed on " + name)
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return obj, True


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    present = {n: f for n, f in SOURCES.items() if os.path.exists(os.path.join(CSRC, n))}
    with ThreadPoolExecutor(max_workers=8) as ex:
        results = list(ex.map(lambda kv: _compile(kv[0], kv[1], verbose, force),
                              present.items()))
    objs = [o for o, _ in results]
    if any(changed for _, changed in results) or not os.path.exists(LIB) or force:
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
