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
# kNN: 1024 unit-norm queries against a 10k x 384 index, k = 10
from effocr_b200.engine import FlatIPIndex
g = torch.Generator().manual_seed(1)
ix = FlatIPIndex(384)
ix.add(torch.nn.functional.normalize(torch.randn(10000, 384, generator=g), dim=1))
q = torch.nn.functional.normalize(torch.randn(1024, 384, generator=g), dim=1).cuda()
for _ in range(2):
    ix.search_device(q, 10)
torch.cuda.synchronize()
print("knn done")
