"""ncu target: one ViT-S layer at batch 1024 (M = 201728): QKV GEMM, attention, proj GEMM, LayerNorm, fused MLP block;
one launch each after warm-up."""
import sys
import torch
sys.path.insert(0, ".")
from effocr_b200 import ops

M = 201728
dev = "cuda"
torch.manual_seed(0)
h = (torch.randn(M, 384, device=dev) * 0.5).half()
mid = (torch.randn(M, 1536, device=dev) * 0.5).half()
x = torch.randn(M, 384, device=dev)
wqkv = (torch.randn(1152, 384, device=dev) * 0.05).half()
wproj = (torch.randn(384, 384, device=dev) * 0.05).half()
wfc1 = (torch.randn(1536, 384, device=dev) * 0.05).half()
wfc2 = (torch.randn(384, 1536, device=dev) * 0.05).half()
b1152 = torch.randn(1152, device=dev); b384 = torch.randn(384, device=dev); b1536 = torch.randn(1536, device=dev)
qkv = torch.empty(M, 1152, device=dev, dtype=torch.float16)
o1536 = torch.empty(M, 1536, device=dev, dtype=torch.float16)
g = torch.ones(384, device=dev)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for _ in range(reps):
    ops.gemm(h, wqkv, bias=b1152, out=qkv)
    ops.attention(qkv, 1024, 6)
    ops.gemm(h, wproj, bias=b384, out_dtype=torch.float32, resid=x, out=x)
    ops.layernorm(x, g, b384)
    ops.mlp_fused(x, h, wfc1, b1536, wfc2, b384)  # fc1 + GELU + fc2 + residual in one kernel
torch.cuda.synchronize()
print("done")
