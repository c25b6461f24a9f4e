"""Check the restated FlowNet2 architecture against the reference's own class definitions and write
tests/golden/flownet2_arch.json (state_dict keys -> shapes, + output digests of the sub-networks that
contain no custom operator).

Build container only: imports /root/reference/models/flownet2_pytorch (its `*_cuda` imports are
satisfied by the rebuilt extensions in oracle/_ref, which import fine without a GPU).
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("IR2RGB_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
sys.path.insert(0, os.path.join(REF, "models"))


def main():
    import flownet2_pytorch.models as ref_models            # the reference's FlowNet2
    from ir2rgb_b200.models.flownet2_pytorch import models as new_models
    torch.manual_seed(0)
    ref = ref_models.FlowNet2().eval()
    new = new_models.FlowNet2().eval()
    sd = ref.state_dict()
    assert list(sd.keys()) == list(new.state_dict().keys()) or set(sd.keys()) == set(new.state_dict().keys())
    new.load_state_dict(sd)                                   # strict: same keys, same shapes
    arch = {k: list(v.shape) for k, v in sd.items()}
    digests = {}
    with torch.no_grad():
        torch.manual_seed(1)
        for name, cin in [("flownets_1", 12), ("flownets_2", 12), ("flownets_d", 6), ("flownetfusion", 11)]:
            x = torch.randn(1, cin, 64, 128)
            a, b = getattr(ref, name)(x), getattr(new, name)(x)
            a = a[0] if isinstance(a, tuple) else a
            b = b[0] if isinstance(b, tuple) else b
            assert torch.equal(a, b), name
            digests[name] = {"shape": list(a.shape), "sum": float(a.double().sum())}
    out = {"n_params": sum(v.numel() for v in sd.values()), "state_dict": arch, "subnet_equal_to_reference": digests}
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "flownet2_arch.json"), "w"), indent=0)
    print("params", out["n_params"], "sub-networks bit-identical to the reference:", sorted(digests))


if __name__ == "__main__":
    main()
