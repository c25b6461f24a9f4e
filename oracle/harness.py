"""FlowNet2 harnesses built from the ORACLE operators -- test infrastructure only.

``make_flownet(kind)`` returns the ``FlowNet`` wrapper (models/flownet.py semantics) whose three custom
operators are replaced by
    kind="torch"  the pure-PyTorch oracle (runs on CPU: bench.py's cpu_baseline leg, CPU integration tests)
    kind="ref"    the reference's own CUDA extensions from oracle/_ref (GPU integration tests)
    kind="refclass"  the reference's OWN `FlowNet2` class (models/flownet2_pytorch/models.py:30-161, unmodified,
                  byte-compiled by oracle/build_ref.py into oracle/_ref/refpy because /root/reference does not
                  exist on the GPU box) with its own Python operator wrappers importing its own rebuilt CUDA
                  extensions -- nothing of this repo's architecture restatement is on that path
                  (GPU: bench.py --impl reference, GPU integration tests)
and whose glue is the reference's unfused chain (models.py:109-150, flownet.py:50).  The conv body is
the same stock-PyTorch architecture in all arms; weights are shared by passing ``state_dict``.
"""
import torch
import torch.nn as nn

from . import torch_ref as tr


class _TorchCorrelation(nn.Module):
    def forward(self, a, b):
        return tr.correlation(a, b, 20, 1, 20, 1, 2)


class _TorchResample2d(nn.Module):
    def forward(self, img, flow):
        return tr.resample2d(img.contiguous(), flow)


class _TorchChannelNorm(nn.Module):
    def forward(self, x):
        return tr.channelnorm(x)


class _RefCorrelation(nn.Module):
    def forward(self, a, b):
        from . import ref_ext
        return ref_ext.correlation_forward(a.contiguous(), b.contiguous(), 20, 1, 20, 1, 2)


class _RefResample2d(nn.Module):
    def forward(self, img, flow):
        from . import ref_ext
        return ref_ext.resample2d_forward(img.contiguous(), flow.contiguous())      # resample2d.py:45


class _RefChannelNorm(nn.Module):
    def forward(self, x):
        from . import ref_ext
        return ref_ext.channelnorm_forward(x.contiguous())


def swap_ops(flownet2, kind):
    """Replace the custom-operator modules of a FlowNet2 instance (attribute names as in the reference:
    flownetc.corr, resample, channelnorm) and switch the fused glue off."""
    mods = {"torch": (_TorchCorrelation, _TorchResample2d, _TorchChannelNorm),
            "ref": (_RefCorrelation, _RefResample2d, _RefChannelNorm)}[kind]
    flownet2.flownetc.corr = mods[0]()
    flownet2.resample = mods[1]()
    flownet2.channelnorm = mods[2]()
    flownet2.fuse_glue = False
    # ... and every libflowops epilogue / concat fusion of the restated conv body: the oracle network must run
    # stock torch layers only (conv -> bias -> LeakyReLU, torch.cat), like the reference's submodules.py
    for m in flownet2.modules():
        if type(m).__name__ == "ConvAct":
            m.fusable = lambda x: False
        if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
            m.__dict__["_flowops_stock"] = True          # submodules.apply_conv: bias stays inside the torch convolution
    return flownet2


def reference_flownet2_available():
    from . import build_ref, ref_ext
    return ref_ext.available() and build_ref.pyc_built()


def load_reference_flownet2_module():
    """Import the reference's unmodified `flownet2_pytorch.models` from oracle/_ref/refpy with the three compiled
    extension modules it imports (`correlation_cuda`, `resample2d_cuda`, `channelnorm_cuda`) resolved to the
    reference's own rebuilt .so files in oracle/_ref."""
    import importlib
    import sys
    from . import build_ref, ref_ext
    if not reference_flownet2_available():
        raise RuntimeError("oracle/_ref is incomplete: run `python oracle/build_ref.py` where /root/reference exists")
    for name in ("correlation_cuda", "resample2d_cuda", "channelnorm_cuda"):
        sys.modules[name] = ref_ext._load(name)
    build_ref.install_finder()
    return importlib.import_module("flownet2_pytorch.models")


class OracleFlowNet(nn.Module):
    """models/flownet.py:20-57 with oracle operators (no libflowops anywhere on this path)."""

    def __init__(self, kind, device, state_dict=None):
        super().__init__()
        if kind == "refclass":
            net = load_reference_flownet2_module().FlowNet2()             # the reference's class, ops and glue as they are
            if state_dict is not None:
                net.load_state_dict(state_dict)
            self.flowNet = net.to(device).eval()
        else:
            from ir2rgb_b200.models.flownet2_pytorch import models as m   # architecture definition only
            net = m.FlowNet2()
            if state_dict is not None:
                net.load_state_dict(state_dict)
            self.flowNet = swap_ops(net, kind).to(device).eval()
        # flownet.py:50 calls `self.resample`, which in the reference resolves to the METHOD Model.resample
        # (base_model.py:129: get_grid + F.grid_sample), not to the Resample2d submodule of the same name
        self.resample = tr.networks_resample

    @torch.no_grad()
    def forward(self, im1, im2):
        data1 = torch.cat([im1.unsqueeze(2), im2.unsqueeze(2)], dim=2)
        flow1 = self.flowNet(data1)
        t = im1 - self.resample(im2, flow1)
        conf = (torch.sum(t * t, dim=1, keepdim=True) < 0.02).float()               # flownet.py:50,56-57
        return flow1, conf
