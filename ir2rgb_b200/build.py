"""Build libflowops.so (the C-ABI library of hand-written sm_100a kernels) in-tree with nvcc.

    python -m ir2rgb_b200.build [--force] [-v]

The library links only against the CUDA runtime (shared, so that it uses the runtime instance -- and
therefore the current device and streams -- of the hosting process, e.g. PyTorch's); there are no
torch headers and no torch dispatch anywhere in it.  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
LIB = os.path.join(HERE, "libflowops.so")
SOURCES = ["api.cu", "cnorm.cu", "warp.cu", "warp16.cu", "corr.cu", "corr_generic.cu", "corr_fast.cu", "corr_bwd.cu", "fused.cu", "epilogue.cu"]
HEADERS = ["common.cuh", "io16.cuh", "corr.cuh", "warp.cuh", "warp_rows.cuh", "warp_rows_bwd.cuh", "tma.cuh", os.path.join("..", "..", "include", "flowops.h")]

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    os.makedirs(OBJ, exist_ok=True)
    jobs = []
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s.replace(".cu", ".o"))
        if force or _stale(obj, [src] + hdrs):
            jobs.append([NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n%s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return r

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            list(ex.map(run, jobs))
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in srcs]
    if force or jobs or _stale(LIB, objs):
        run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "shared", "-o", LIB] + objs +
            ["-Xlinker", "-rpath", "-Xlinker", "/usr/local/cuda/lib64"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
