"""Generates tests/golden/vid2vid_arch.json from the reference's own classes (run in the build container, where
/root/reference exists): state-dict keys and shapes of the composite generator and the multi-scale discriminators
at BASELINE configs[4] (ngf 128, 3 down-sampling layers, 9 blocks, batch norm; ndf 64, 3 layers, num_D 2)."""
import json
import os
import sys

sys.path.insert(0, "/root/reference")
from models import networks as ref          # noqa: E402

kw = dict(gen_blocks=9, n_local_enhancers=1, feat_num=3, n_blocks_local=3, fg=False, no_flow=False)
G = ref.build_generator_module(9, 3, 6, 128, "composite", 3, "batch", 0, **kw)
D = ref.build_discriminator_module(6, 64, 3, "batch", 2, True)
DT = ref.build_discriminator_module(3 * 3 + 2 * 2, 64, 3, "batch", 2, True)
out = {name: {k: list(v.shape) for k, v in m.state_dict().items()} for name, m in (("G", G), ("D", D), ("D_T", DT))}
out["n_params"] = {name: sum(p.numel() for p in m.parameters()) for name, m in (("G", G), ("D", D), ("D_T", DT))}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "vid2vid_arch.json"), "w"))
print(out["n_params"])
