"""Middlebury `.flo` files: the on-disk format of optical flow next to the hot path (counterpart of the reference's
models/flownet2_pytorch/utils/flow_utils.py:5-55 -- the format FlowNet2's tooling reads and writes).

Layout (little endian): float32 tag 202021.25 ("PIEH"), int32 width, int32 height, then height * width interleaved
(u, v) float32 pairs in row-major order.  The flow tensors of this package are [2, H, W] (channel 0 = u = dx); the
helpers accept / return the reference's [H, W, 2] arrays and convert from / to that layout.
"""
import numpy as np

TAG = np.float32(202021.25)


def read_flow(path):
    """Returns an [H, W, 2] float32 array (flow_utils.py:5-25); raises on a bad tag or a truncated file instead of
    returning None."""
    with open(path, "rb") as f:
        head = np.fromfile(f, dtype="<f4", count=1)
        if head.size != 1 or head[0] != TAG:
            raise ValueError("%s: not a .flo file (bad magic number)" % path)
        wh = np.fromfile(f, dtype="<i4", count=2)
        if wh.size != 2 or wh[0] <= 0 or wh[1] <= 0:
            raise ValueError("%s: bad .flo header" % path)
        w, h = int(wh[0]), int(wh[1])
        data = np.fromfile(f, dtype="<f4", count=2 * w * h)
    if data.size != 2 * w * h:
        raise ValueError("%s: truncated .flo file (%d of %d values)" % (path, data.size, 2 * w * h))
    return data.reshape(h, w, 2).astype(np.float32, copy=False)


def write_flow(path, uv, v=None):
    """uv: [H, W, 2], or u and v as two [H, W] arrays (flow_utils.py:27-55)."""
    if v is None:
        uv = np.asarray(uv)
        if uv.ndim != 3 or uv.shape[2] != 2:
            raise ValueError("expected an [H, W, 2] array, got %s" % (uv.shape,))
        u, v = uv[:, :, 0], uv[:, :, 1]
    else:
        u, v = np.asarray(uv), np.asarray(v)
    if u.shape != v.shape or u.ndim != 2:
        raise ValueError("u and v must be [H, W] arrays of the same shape")
    h, w = u.shape
    inter = np.empty((h, w, 2), dtype="<f4")
    inter[:, :, 0], inter[:, :, 1] = u, v
    with open(path, "wb") as f:
        np.array([TAG], dtype="<f4").tofile(f)
        np.array([w, h], dtype="<i4").tofile(f)
        inter.tofile(f)


def flow_to_tensor_layout(hw2):
    """[H, W, 2] -> [2, H, W] (channel 0 = dx, channel 1 = dy: what Resample2d / networks.resample take)."""
    return np.ascontiguousarray(np.transpose(hw2, (2, 0, 1)))


def flow_from_tensor_layout(chw):
    """[2, H, W] (numpy or torch) -> [H, W, 2] float32."""
    a = chw.detach().cpu().numpy() if hasattr(chw, "detach") else np.asarray(chw)
    return np.ascontiguousarray(np.transpose(a, (1, 2, 0))).astype(np.float32, copy=False)
