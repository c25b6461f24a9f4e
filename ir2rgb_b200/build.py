"""Build libflowops.so (the C-ABI library of hand-written sm_100a kernels) in-tree with nvcc.

    python -m ir2rgb_b200.build [--force] [-v] [--out PATH -DNAME=VALUE ...]

`--out` with `-D` flags builds a VARIANT of the library (own object directory next to PATH) for A/B timing with
tools/ab_ops.py; the kernels expose a few tuning macros for that (grep FLOWOPS_TUNE in csrc/).  The default build
takes no -D flags.

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
SOURCES = ["api.cu", "cnorm.cu", "warp.cu", "warp16.cu", "corr.cu", "corr_generic.cu", "corr_fast.cu", "corr_tc.cu", "corr_tc_bwd.cu", "corr_bwd.cu", "fused.cu", "epilogue.cu", "flowhead.cu"]
HEADERS = ["common.cuh", "io16.cuh", "corr.cuh", "warp.cuh", "warp_rows.cuh", "warp_rows_bwd.cuh", "warp_win_bwd.cuh", "warp_fx_bwd.cuh", "tma.cuh", "tc.cuh", os.path.join("..", "..", "include", "flowops.h")]

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, out=None, defines=()):
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    lib_path, obj_dir = LIB, OBJ
    if out is not None:                       # a variant build: never touches the in-tree library or its objects
        lib_path = os.path.abspath(out)
        obj_dir = lib_path + ".obj"
        force = True
    os.makedirs(obj_dir, exist_ok=True)
    jobs = []
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(obj_dir, s.replace(".cu", ".o"))
        if force or _stale(obj, [src] + hdrs):
            jobs.append([NVCC] + NVCC_FLAGS + list(defines) + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj])

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
    objs = [os.path.join(obj_dir, s.replace(".cu", ".o")) for s in srcs]
    if force or jobs or _stale(lib_path, objs):
        run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "shared", "-o", lib_path] + objs +
            ["-Xlinker", "-rpath", "-Xlinker", "/usr/local/cuda/lib64"])
    return lib_path


if __name__ == "__main__":
    argv = sys.argv[1:]
    out = argv[argv.index("--out") + 1] if "--out" in argv else None
    print(build(force="--force" in argv, verbose="-v" in argv, out=out, defines=[a for a in argv if a.startswith("-D")]))
