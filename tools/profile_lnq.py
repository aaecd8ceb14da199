"""ncu target: the fused norm1 + QKV kernel (ops.ln_gemm) against its unfused pair at M = 201728; CUDA-event timings printed."""
import sys
import torch
sys.path.insert(0, ".")
from effocr_b200 import ops

M = 201728
dev = "cuda"
torch.manual_seed(0)
x = torch.randn(M, 384, device=dev)
wqkv = (torch.randn(1152, 384, device=dev) * 0.05).half()
b1152 = torch.randn(1152, device=dev)
g = torch.ones(384, device=dev); b = torch.zeros(384, device=dev)
qkv = torch.empty(M, 1152, device=dev, dtype=torch.float16)
big = torch.empty(256 << 20, device=dev, dtype=torch.uint8)


def timed(fn, n=10):
    ts = []
    for _ in range(n):
        big.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def unfused():
    h = ops.layernorm(x, g, b)
    ops.gemm(h, wqkv, bias=b1152, out=qkv)


def fused():
    ops.ln_gemm(x, g, b, wqkv, bias=b1152, out=qkv)


for _ in range(2):
    unfused(); fused()
torch.cuda.synchronize()
if len(sys.argv) > 1 and sys.argv[1] == "time":
    print("unfused us", timed(unfused), "fused us", timed(fused))
print("done")
