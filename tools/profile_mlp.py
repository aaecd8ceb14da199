"""ncu target: the fused MLP block kernel at ViT-S batch 1024 (M = 201728)."""
import sys
import torch
sys.path.insert(0, ".")
from effocr_b200 import ops
M, D, HID = 201728, 384, 1536
torch.manual_seed(0)
h = (torch.randn(M, D, device="cuda") * 0.7).half()
w1 = (torch.randn(HID, D, device="cuda") * 0.05).half()
w2 = (torch.randn(D, HID, device="cuda") * 0.05).half()
b1 = torch.randn(HID, device="cuda"); b2 = torch.randn(D, device="cuda")
x = torch.randn(M, D, device="cuda")
for _ in range(3):
    ops.mlp_fused(x, h, w1, b1, w2, b2)
torch.cuda.synchronize()
print("done")
