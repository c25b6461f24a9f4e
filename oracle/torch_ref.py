"""Pure-PyTorch restatement of the flow hot path (the second oracle BASELINE.json names).

Test infrastructure only.  Everything here is differentiable and dtype-generic (fp32 / fp64), so the
backwards come from autograd and an fp64 run gives the "truth" the fp32 kernels are scored against.

Closed forms: SURVEY.md Appendix A (checked there against line-by-line transliterations of the
reference kernels).
"""
import torch
import torch.nn.functional as F


def correlation(in1, in2, pad_size=20, kernel_size=1, max_displacement=20, stride1=1, stride2=2):
    """correlation_cuda_kernel.cu:74-147 for kernel_size=1, stride1=1, pad_size == max_displacement.

    out[n, (tj+r)*D + (ti+r), y, x] = mean_c in1[n,c,y,x] * in2pad[n,c,y + tj*s2, x + ti*s2]
    (tj = vertical = slow index; in2pad zero outside the image).
    """
    assert kernel_size == 1 and stride1 == 1 and pad_size == max_displacement
    B, C, H, W = in1.shape
    r = max_displacement // stride2
    md = r * stride2
    in2p = F.pad(in2, (md, md, md, md))
    outs = []
    for tj in range(-r, r + 1):
        for ti in range(-r, r + 1):
            ys, xs = md + tj * stride2, md + ti * stride2
            outs.append((in1 * in2p[:, :, ys:ys + H, xs:xs + W]).sum(1) / C)
    return torch.stack(outs, 1)


def vid2vid_grid(B, H, W, device, dtype):
    """models/networks.py:15-28 (get_grid): linspace(-1, 1) in fp32 on the host, then cast/moved."""
    hor = torch.linspace(-1.0, 1.0, W).view(1, 1, 1, W).expand(B, 1, H, W)
    ver = torch.linspace(-1.0, 1.0, H).view(1, 1, H, 1).expand(B, 1, H, W)
    return torch.cat([hor, ver], 1).to(dtype).to(device)


def resample2d(img, flow):
    """resample2d_kernel.cu:16-64 == grid_sample(bilinear, border, align_corners=True) on vid2vid's
    grid.  Built from explicit pixel coordinates to avoid the normalise /
    un-normalise round trip."""
    B, C, H, W = img.shape
    dev, dt = img.device, img.dtype
    xs = torch.arange(W, device=dev, dtype=dt).view(1, 1, W) + flow[:, 0]
    ys = torch.arange(H, device=dev, dtype=dt).view(1, H, 1) + flow[:, 1]
    x0, y0 = torch.floor(xs), torch.floor(ys)
    a, b = (xs - x0).unsqueeze(1), (ys - y0).unsqueeze(1)
    xL = x0.long().clamp(0, W - 1)
    xR = (x0.long() + 1).clamp(0, W - 1)
    yT = y0.long().clamp(0, H - 1)
    yB = (y0.long() + 1).clamp(0, H - 1)
    flat = img.reshape(B, C, H * W)

    def take(yy, xx):
        idx = (yy * W + xx).view(B, 1, H * W).expand(B, C, H * W)
        return flat.gather(2, idx).view(B, C, H, W)

    return ((1 - a) * (1 - b) * take(yT, xL) + a * (1 - b) * take(yT, xR)
            + (1 - a) * b * take(yB, xL) + a * b * take(yB, xR))


def networks_resample(img, flow, grid=None):
    """models/networks.py:93-100 / base_model.py:129-136, as run: grid_sample with the default
    align_corners=False on a grid built for the align_corners=True convention."""
    B, C, H, W = img.shape
    if grid is None:
        grid = vid2vid_grid(B, H, W, flow.device, flow.dtype)
    nflow = torch.cat([flow[:, 0:1] / ((W - 1.0) / 2.0), flow[:, 1:2] / ((H - 1.0) / 2.0)], dim=1)
    final_grid = (grid + nflow).permute(0, 2, 3, 1)
    return F.grid_sample(img, final_grid, mode='bilinear', padding_mode='border', align_corners=False)


def channelnorm(x):
    """channelnorm_kernel.cu:51-59."""
    return torch.sqrt((x * x).sum(1, keepdim=True))


def channelnorm_bwd(x, y, gy):
    """channelnorm_kernel.cu:92-95 (note the 1e-9, which autograd of sqrt does not have)."""
    return gy * x / (y + 1e-9)
