"""CPU tests of the C-ABI boundary: the library builds for sm_100a, loads through ctypes, exports every
symbol include/flowops.h declares, and rejects bad arguments before touching a device."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "flowops.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(flowops_\w+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for s in ["flowops_cnorm_fwd", "flowops_cnorm_bwd", "flowops_warp_fwd", "flowops_warp_bwd",
              "flowops_corr_fwd", "flowops_corr_bwd", "flowops_corr_out_shape",
              "flowops_corr_fwd_workspace_bytes", "flowops_corr_bwd_workspace_bytes",
              "flowops_version", "flowops_last_error"]:
        assert s in syms


def test_library_exports_every_declared_symbol(flowops_lib):
    from ir2rgb_b200 import _lib
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(raw, s), "libflowops.so does not export %s" % s
    # and the ctypes signature table covers exactly the header
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_library_is_sm100a_and_has_no_torch_dependency(flowops_lib):
    from ir2rgb_b200 import _lib
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libcudart" in out
    for forbidden in ("libtorch", "libc10", "libpython"):
        assert forbidden not in out
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if os.path.exists(cuobjdump):
        elf = subprocess.run([cuobjdump, "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
        assert "sm_100a" in elf
        assert not re.search(r"sm_(?!100a)\d+", elf), "only sm_100a code is shipped"


def test_fast_correlation_kernel_uses_tma(flowops_lib):
    from ir2rgb_b200 import _lib
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTMALDG" in sass, "the Correlation fast path stages its tiles with TMA"
    assert "RED.E.ADD.F32" in sass or "REDG" in sass or "RED." in sass, "warp backward uses fire-and-forget reductions"


def test_tensor_core_correlation_is_tcgen05(flowops_lib):
    """The Correlation forward runs on the 5th-generation tensor cores: tcgen05.mma (UTC*MMA), accumulators read back from
    TMEM with tcgen05.ld (LDTM), 5-D TMA loads, TMEM allocation -- in SASS, where the PTX names never appear."""
    from ir2rgb_b200 import _lib
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", "-fun", "corr_fwd_tc", _lib.LIB_PATH], capture_output=True, text=True).stdout
    if "corr_fwd_tc" not in sass:      # older cuobjdump: no -fun filter on shared objects
        sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert re.search(r"\bUTC\w*MMA\b", sass), "tcgen05.mma"
    assert "LDTM" in sass, "tcgen05.ld"
    assert "UTMALDG.5D" in sass, "5-D TMA loads of the K-major planes"
    assert not re.search(r"\bHMMA\b|\bHGMMA\b", sass), "no legacy mma.sync / wgmma path"


def test_tensor_core_correlation_backward_is_tcgen05_with_the_a_operand_in_tmem(flowops_lib):
    """The Correlation backward: tcgen05.mma, the skewed-gradient operand written into TMEM with tcgen05.st (STTM),
    accumulators read back with tcgen05.ld (LDTM), 5-D TMA loads of the feature planes."""
    from ir2rgb_b200 import _lib
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    funcs = [f for f in sass.split("Function : ")[1:] if f.startswith("_ZN7flowops3tcb11corr_bwd_tcILb1E")]
    assert funcs, "corr_bwd_tc<true, ...> is in the library"
    for f in funcs:
        assert re.search(r"\bUTC\w*MMA\b", f) and "STTM" in f and "LDTM" in f and "UTMALDG.5D" in f


def test_bad_arguments_are_rejected_without_a_device(flowops_lib):
    lib = flowops_lib
    assert lib.flowops_version() == 1
    assert lib.flowops_cnorm_fwd(None, None, 1, 3, 4, 4, None) == -1
    assert b"null pointer" in lib.flowops_last_error()
    one = ctypes.c_void_p(16)
    assert lib.flowops_cnorm_fwd(one, one, 0, 3, 4, 4, None) == -1
    assert lib.flowops_warp_fwd(one, one, one, 1, 3, 4, 4, 7, None, None, None) == -1         # unknown mode
    assert lib.flowops_warp_fwd(one, one, one, 1, 3, 4, 4, 1, None, None, None) == -1         # tables missing
    assert lib.flowops_warp_bwd(one, one, one, None, None, 1, 3, 4, 4, 0, None, None, None) == -1
    assert lib.flowops_corr_fwd(one, one, one, 1, 4, 8, 8, 20, 2, 20, 1, 2, 0, None, 0, None) == -1  # even kernel
    assert lib.flowops_corr_bwd(one, one, one, one, one, 1, 4, 8, 8, 4, 1, 4, 2, 2, None, 0, None) == -2
    # fast path without its workspace
    assert lib.flowops_corr_fwd(one, one, one, 1, 4, 8, 8, 20, 1, 20, 1, 2, 0, None, 0, None) == -3


def test_shapes_and_workspace(flowops_lib):
    lib = flowops_lib
    oc, oh, ow = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert lib.flowops_corr_out_shape(48, 64, 20, 1, 20, 1, 2, ctypes.byref(oc), ctypes.byref(oh), ctypes.byref(ow)) == 0
    assert (oc.value, oh.value, ow.value) == (441, 48, 64)
    assert lib.flowops_corr_out_shape(10, 10, 4, 1, 4, 2, 2, ctypes.byref(oc), ctypes.byref(oh), ctypes.byref(ow)) == 0
    assert (oc.value, oh.value, ow.value) == (25, 5, 5)
    # FlowNetC configuration, FP32-FMA kernel: two parity-plane copies (the role of the reference's rbot1/rbot2, without
    # the 20-pixel padding); the f2 planes carry 4 extra columns per row for TMA start alignment
    prev = lib.flowops_corr_get_impl()
    try:
        lib.flowops_corr_set_impl(0)
        assert lib.flowops_corr_fwd_workspace_bytes(8, 256, 48, 64, 20, 1, 20, 1, 2) == 8 * 4 * 256 * 24 * (32 + 36) * 4
        # tensor-core kernel: two K-major plane copies without any padding (the kernel stores either output layout itself),
        # i.e. less than the FP32-FMA kernel's planes; the entry point reports the larger of the two so that the caller's
        # workspace serves whichever kernel the call selects
        lib.flowops_corr_set_impl(1)
        assert lib.flowops_corr_fwd_workspace_bytes(8, 256, 48, 64, 20, 1, 20, 1, 2) == max(2 * 8 * 256 * 48 * 64, 8 * 4 * 256 * 24 * (32 + 36)) * 4
    finally:
        lib.flowops_corr_set_impl(prev)
    assert lib.flowops_corr_fwd_workspace_bytes(1, 8, 9, 11, 4, 1, 4, 1, 1) == 0               # generic kernel


def test_oracle_shape_agrees_with_library(flowops_lib, c_oracle):
    oc, oh, ow = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    for args in [(48, 64, 20, 1, 20, 1, 2), (9, 11, 4, 1, 4, 1, 1), (10, 10, 4, 1, 4, 2, 2), (7, 9, 3, 3, 2, 1, 1),
                 (11, 13, 6, 1, 4, 3, 2)]:
        assert flowops_lib.flowops_corr_out_shape(*args, ctypes.byref(oc), ctypes.byref(oh), ctypes.byref(ow)) == 0
        assert (oc.value, oh.value, ow.value) == c_oracle.corr_shape(*args)
