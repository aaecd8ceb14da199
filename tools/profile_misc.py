"""ncu target: crop_resize (1024 synthetic crops, patch-major) and tcgen05 attention (batch 1024, 6 heads)."""
import sys
import torch
sys.path.insert(0, ".")
from effocr_b200 import ops, synth
from effocr_b200.pipeline import PackedCrops

crops, _ = synth.synthetic_crops(1024, seed=0)
px, im, bx, n = PackedCrops(crops).to_device()
out = torch.empty(n * 196, 768, device="cuda", dtype=torch.float16)
qkv = (torch.randn(1024 * 197, 1152, device="cuda") * 0.5).half()
for _ in range(3):
    ops.crop_resize(px, im, bx, n, ops.CROP_PATCH_F16, out=out)
    ops.attention(qkv, 1024, 6)
torch.cuda.synchronize()
print("done")
