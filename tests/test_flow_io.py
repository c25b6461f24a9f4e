"""`.flo` IO against the reference's own reader / writer (imported from /root/reference in the build container) and a
committed golden file everywhere."""
import importlib.util
import os

import numpy as np
import pytest

from ir2rgb_b200.utils import flow_io

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/models/flownet2_pytorch/utils/flow_utils.py"
GOLD = os.path.join(HERE, "golden", "flow_5x7.flo")


def _sample():
    rng = np.random.default_rng(42)
    return (10 * rng.standard_normal((5, 7, 2))).astype(np.float32)


def test_golden_file_written_by_the_reference_reads_back():
    flow = flow_io.read_flow(GOLD)
    assert flow.shape == (5, 7, 2) and flow.dtype == np.float32
    assert np.array_equal(flow, _sample())


def test_write_is_byte_identical_to_the_golden_file(tmp_path):
    out = tmp_path / "a.flo"
    flow_io.write_flow(str(out), _sample())
    assert out.read_bytes() == open(GOLD, "rb").read()
    u, v = _sample()[:, :, 0], _sample()[:, :, 1]
    flow_io.write_flow(str(out), u, v)
    assert out.read_bytes() == open(GOLD, "rb").read()


def test_layout_helpers_round_trip():
    f = _sample()
    t = flow_io.flow_to_tensor_layout(f)
    assert t.shape == (2, 5, 7) and np.array_equal(t[0], f[:, :, 0])
    assert np.array_equal(flow_io.flow_from_tensor_layout(t), f)


def test_bad_files_raise(tmp_path):
    p = tmp_path / "bad.flo"
    p.write_bytes(b"\x00" * 12)
    with pytest.raises(ValueError):
        flow_io.read_flow(str(p))
    good = open(GOLD, "rb").read()
    p.write_bytes(good[:-4])
    with pytest.raises(ValueError):
        flow_io.read_flow(str(p))


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present (GPU box): covered by the golden file")
def test_against_reference_reader_and_writer(tmp_path):
    spec = importlib.util.spec_from_file_location("ref_flow_utils", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    f = (3 * np.random.default_rng(1).standard_normal((9, 4, 2))).astype(np.float32)
    a, b = tmp_path / "ref.flo", tmp_path / "new.flo"
    ref.writeFlow(str(a), f)
    flow_io.write_flow(str(b), f)
    assert a.read_bytes() == b.read_bytes()
    assert np.array_equal(ref.readFlow(str(b)), flow_io.read_flow(str(a)))
