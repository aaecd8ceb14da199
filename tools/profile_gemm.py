"""ncu target: one ViT-S layer at batch 1024 (M = 201728).  The product path is three kernels -- norm1 + QKV (ln_gemm), attention,
block tail (projection + residual + norm2 + MLP + residual) -- followed here by the kernels they replaced (QKV GEMM, proj_ln, LayerNorm,
mlp_fused) for comparison; one launch each per repetition after warm-up."""
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
h2 = torch.empty(M, 384, device=dev, dtype=torch.float16)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for _ in range(reps):
    ops.ln_gemm(x, g, b384, wqkv, bias=b1152, out=qkv)      # norm1 + QKV projection in one kernel
    ops.attention(qkv, 1024, 6)
    ops.block_tail(x, h, wproj, b384, g, b384, wfc1, b1536, wfc2, b384)  # projection + residual + norm2 + MLP + residual
    ops.gemm(h, wqkv, bias=b1152, out=qkv)
    ops.attention(qkv, 1024, 6)
    ops.proj_ln(x, h, wproj, b384, g, b384, out=h2)  # projection + residual + norm2 in one kernel
    ops.layernorm(x, g, b384)                        # norm1 of the next block
    ops.mlp_fused(x, h, wfc1, b1536, wfc2, b384)  # fc1 + GELU + fc2 + residual in one kernel
torch.cuda.synchronize()
print("done")
