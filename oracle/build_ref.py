"""Build the reference's own CUDA extensions for sm_100 into oracle/_ref/ (git-ignored).

Test infrastructure only: the resulting modules are the GPU-side oracle (and the reference arm of
bench.py).  Sources are compiled from where they lie in /root/reference -- they are staged into a
throw-away directory under /tmp because two mechanical patches are needed:

  1. the gencode list (sm_37 ... sm_70 in */setup.py is rejected by CUDA 12) is replaced by
     -gencode arch=compute_100,code=sm_100 (passed here on the command line; setup.py is not used);
  2. `.type()` -> `.scalar_type()` inside the AT_DISPATCH_* calls (removed API in torch >= 2.x).

Only built artefacts are written into the repo tree (oracle/_ref/, git-ignored), never the sources: the three .so
files, and -- so that bench.py's reference arm and the GPU tests can run the reference's OWN `FlowNet2` class and its
OWN operator wrappers on the GPU box, where /root/reference does not exist -- the byte-compiled (.pyc, sourceless)
form of the unmodified Python modules of models/flownet2_pytorch that class needs (oracle/_ref/refpy/).
Run:  python oracle/build_ref.py          (needs /root/reference; ~2-5 min on 8 cores, no GPU needed)
"""
import glob
import os
import re
import shutil
import sys
import tempfile

REF_ROOT = os.environ.get("IR2RGB_REFERENCE", "/root/reference")
NETS = os.path.join(REF_ROOT, "models", "flownet2_pytorch", "networks")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")

EXTS = {
    "correlation_cuda": ("correlation_package", ["correlation_cuda.cc", "correlation_cuda_kernel.cu"]),
    "resample2d_cuda": ("resample2d_package", ["resample2d_cuda.cc", "resample2d_kernel.cu"]),
    "channelnorm_cuda": ("channelnorm_package", ["channelnorm_cuda.cc", "channelnorm_kernel.cu"]),
}


# the reference's unmodified Python modules that define FlowNet2 and the three operator wrappers (paths below
# models/flownet2_pytorch); byte-compiled into PYC_OUT as sourceless modules
PYC_OUT = os.path.join(OUT, "refpy")
PY_MODULES = ["__init__.py", "models.py", "networks/__init__.py", "networks/submodules.py", "networks/FlowNetC.py",
              "networks/FlowNetS.py", "networks/FlowNetSD.py", "networks/FlowNetFusion.py",
              "networks/correlation_package/__init__.py", "networks/correlation_package/correlation.py",
              "networks/resample2d_package/__init__.py", "networks/resample2d_package/resample2d.py",
              "networks/channelnorm_package/__init__.py", "networks/channelnorm_package/channelnorm.py"]


def available():
    return os.path.isdir(NETS)


PYC_EXT = ".rpyc"          # not ".pyc": snapshot tools commonly drop *.pyc / __pycache__; loaded by RefpyFinder below


def _pyc_path(m):
    return os.path.join(PYC_OUT, "flownet2_pytorch", m[:-3] + PYC_EXT)


def pyc_built():
    return all(os.path.exists(_pyc_path(m)) for m in PY_MODULES)


class RefpyFinder:
    """sys.meta_path finder for the byte-compiled reference package: flownet2_pytorch[.sub.module] ->
    oracle/_ref/refpy/flownet2_pytorch/sub/module.rpyc (packages: .../__init__.rpyc)."""

    @staticmethod
    def find_spec(fullname, path=None, target=None):
        import importlib.machinery
        import importlib.util
        if fullname != "flownet2_pytorch" and not fullname.startswith("flownet2_pytorch."):
            return None
        rel = fullname.split(".")
        base = os.path.join(PYC_OUT, *rel)
        pkg_init = os.path.join(base, "__init__" + PYC_EXT)
        if os.path.exists(pkg_init):
            loader = importlib.machinery.SourcelessFileLoader(fullname, pkg_init)
            return importlib.util.spec_from_file_location(fullname, pkg_init, loader=loader, submodule_search_locations=[base])
        if os.path.exists(base + PYC_EXT):
            loader = importlib.machinery.SourcelessFileLoader(fullname, base + PYC_EXT)
            return importlib.util.spec_from_file_location(fullname, base + PYC_EXT, loader=loader)
        return None


def install_finder():
    import sys
    if not any(f is RefpyFinder for f in sys.meta_path):
        sys.meta_path.insert(0, RefpyFinder)


def build_pyc(force=False):
    """Byte-compile the reference's FlowNet2 package (unmodified) into oracle/_ref/refpy/flownet2_pytorch/**.pyc."""
    import py_compile
    if pyc_built() and not force:
        return PYC_OUT
    src_root = os.path.join(REF_ROOT, "models", "flownet2_pytorch")
    for m in PY_MODULES:
        dst = _pyc_path(m)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        py_compile.compile(os.path.join(src_root, m), cfile=dst, dfile="reference:models/flownet2_pytorch/" + m,
                           doraise=True, invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    return PYC_OUT


def built():
    return all(glob.glob(os.path.join(OUT, name + "*.so")) for name in EXTS)


def build(force=False, verbose=False):
    if not available():
        raise RuntimeError("reference sources not found under %s" % NETS)
    build_pyc(force)
    if built() and not force:
        return OUT
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ.setdefault("MAX_JOBS", str(os.cpu_count() or 4))
    from torch.utils.cpp_extension import load
    os.makedirs(OUT, exist_ok=True)
    for name, (pkg, sources) in EXTS.items():
        stage = tempfile.mkdtemp(prefix="ir2rgb_ref_%s_" % name)
        for fn in os.listdir(os.path.join(NETS, pkg)):
            if fn.endswith((".cc", ".cu", ".cuh", ".h")):
                with open(os.path.join(NETS, pkg, fn)) as f:
                    text = f.read()
                text = re.sub(r"\.type\(\),", ".scalar_type(),", text)       # patch 2
                with open(os.path.join(stage, fn), "w") as f:
                    f.write(text)
        build_dir = os.path.join(stage, "build")
        os.makedirs(build_dir)
        load(name=name, sources=[os.path.join(stage, s) for s in sources],
             extra_cflags=["-O2", "-std=c++17"],
             extra_cuda_cflags=["-gencode", "arch=compute_100,code=sm_100", "-std=c++17"],   # patch 1
             build_directory=build_dir, verbose=verbose, is_python_module=False)
        so = glob.glob(os.path.join(build_dir, name + "*.so"))[0]
        shutil.copy(so, os.path.join(OUT, name + ".so"))
        shutil.rmtree(stage, ignore_errors=True)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built:", sorted(os.listdir(OUT)))
