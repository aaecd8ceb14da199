"""Where does ConvNeXt-Tiny lose precision against the fp32 oracle?  Variants of the weights / input isolate the stages."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from effocr_b200 import synth
from effocr_b200.engine import ConvNextEngine
from oracle import convnext as OC, transform as OT
torch.manual_seed(0)
base = OC.init_convnext_tiny_state_dict(seed=0)
xr = torch.randn(3, 3, 224, 224, generator=torch.Generator().manual_seed(1))
crops, _ = synth.synthetic_crops(6, seed=0)
xc = torch.from_numpy(np.stack([OT.paired_transform(c) for c in crops]))
def run(name, sd, x):
    with torch.no_grad():
        ref = OC.convnext_forward(sd, x)
    out = ConvNextEngine(sd, max_batch=4).forward(x.cuda()).cpu()
    rel = ((out - ref).norm(dim=1) / ref.norm(dim=1))
    print(f"{name:50s} max rel {rel.max():.2e} mean {rel.mean():.2e}", flush=True)
for nm, x in (("randn", xr), ("crops", xc)):
    run(f"{nm}: default", base, x)
    run(f"{nm}: fp16-representable weights", {k: (v.half().float() if v.dim() > 1 else v) for k, v in base.items()}, x)
    run(f"{nm}: gamma = 0 (stem + downsample + head only)", {k: (v * 0 if k.endswith("gamma") else v) for k, v in base.items()}, x)
    for st in range(4):
        sd = {k: (v * 0 if (k.endswith("gamma") and f"stages.{st}." not in k) else v) for k, v in base.items()}
        run(f"{nm}: only stage {st}'s blocks active", sd, x)
