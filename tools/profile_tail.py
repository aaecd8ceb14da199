"""CUDA-event timing of the one-kernel block tail (ops.block_tail) against proj_ln + mlp_fused at M = 201728."""
import sys
import torch
sys.path.insert(0, ".")
from effocr_b200 import ops

M, D, HID = 201728, 384, 1536
torch.manual_seed(0)
x = torch.randn(M, D, device="cuda")
att = (torch.randn(M, D, device="cuda") * 0.7).half()
wp = (torch.randn(D, D, device="cuda") * 0.05).half()
w1 = (torch.randn(HID, D, device="cuda") * 0.05).half()
w2 = (torch.randn(D, HID, device="cuda") * 0.05).half()
bp, g, be, b2 = (torch.randn(D, device="cuda") for _ in range(4))
b1 = torch.randn(HID, device="cuda")
h = torch.empty(M, D, device="cuda", dtype=torch.float16)
big = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)


def timed(fn, n=8):
    ts = []
    for _ in range(n):
        big.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def two():
    ops.proj_ln(x, att, wp, bp, g, be, out=h)
    ops.mlp_fused(x, h, w1, b1, w2, b2)


def one():
    ops.block_tail(x, att, wp, bp, g, be, w1, b1, w2, b2)


for _ in range(2):
    two(); one()
torch.cuda.synchronize()
if len(sys.argv) > 1 and sys.argv[1] == "time":
    print("proj_ln + mlp_fused us", timed(two), "block_tail us", timed(one))
print("done")
