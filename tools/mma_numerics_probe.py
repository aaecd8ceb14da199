"""What does tcgen05.mma kind::f16 do with (a) fp16 subnormal inputs and (b) long fp32 accumulation chains?
Runs the product's own GEMM (effocr_gemm_f16, fp32 output) on inputs whose exact result is known in fp64."""
import sys
sys.path.insert(0, ".")
import torch
from effocr_b200 import ops
torch.manual_seed(0)
M, N = 256, 128
for K in (64, 512, 4608):
    a = torch.rand(M, K).half()          # positive: a truncating adder shows up as a negative bias growing with K
    w = torch.rand(N, K).half()
    ref = a.double() @ w.double().t()
    out = ops.gemm(a.cuda(), w.cuda(), out_dtype=torch.float32).cpu().double()
    rel = (out - ref) / ref
    print(f"K={K}: positive inputs: mean rel err {rel.mean().item():+.3e}, max |rel| {rel.abs().max().item():.3e}  (2^-24 = 5.96e-08)")
    a2 = (torch.rand(M, K) - 0.5).half(); w2 = (torch.rand(N, K) - 0.5).half()
    ref2 = a2.double() @ w2.double().t()
    out2 = ops.gemm(a2.cuda(), w2.cuda(), out_dtype=torch.float32).cpu().double()
    print(f"K={K}: signed inputs: max abs err {(out2 - ref2).abs().max().item():.3e} (|ref| max {ref2.abs().max().item():.2f})")
# subnormal operands
K = 64
a = torch.full((M, K), 3.0e-5).half()     # subnormal in fp16 (< 6.1e-5)
w = torch.ones(N, K).half()
out = ops.gemm(a.cuda(), w.cuda(), out_dtype=torch.float32).cpu()
print(f"subnormal A (3e-5) x ones, K=64: got {out[0, 0].item():.6e}, exact {(a[0].double().sum()).item():.6e}")
w = torch.full((N, K), 3.0e-5).half()
a = torch.ones(M, K).half()
out = ops.gemm(a.cuda(), w.cuda(), out_dtype=torch.float32).cpu()
print(f"ones x subnormal W (3e-5), K=64: got {out[0, 0].item():.6e}, exact {(w[0].double().sum()).item():.6e}")
