"""Stand-alone timing of the kernels of one ViT-S layer at batch 1024 (M = 201 728 rows), CUDA events, 20 launches each.
A/B switches are environment variables read once per process, so run once per variant:
    EFFOCR_PLN_PREFETCH=0 python tools/ab_kernels.py proj_ln"""
import sys
sys.path.insert(0, ".")
import torch
from effocr_b200 import ops
which = set(sys.argv[1:]) or {"ln", "qkv", "ln_qkv", "attention", "proj_ln", "mlp"}
B, T, D, HID = 1024, 197, 384, 1536
M = B * T
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(M, D, device="cuda", generator=g)
h = torch.randn(M, D, device="cuda", generator=g).half()
wqkv = (torch.randn(3 * D, D, device="cuda", generator=g) * 0.02).half()
wproj = (torch.randn(D, D, device="cuda", generator=g) * 0.02).half()
w1 = (torch.randn(HID, D, device="cuda", generator=g) * 0.02).half()
w2 = (torch.randn(D, HID, device="cuda", generator=g) * 0.02).half()
bq, bp, b1, b2 = (torch.zeros(n, device="cuda") for n in (3 * D, D, HID, D))
gam, bet = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
qkv = torch.randn(M, 3 * D, device="cuda", generator=g).half()
att = torch.randn(M, D, device="cuda", generator=g).half()
hout = torch.empty(M, D, device="cuda", dtype=torch.float16)
qout = torch.empty(M, 3 * D, device="cuda", dtype=torch.float16)
fns = {
    "ln": lambda: ops.layernorm(x, gam, bet),
    "qkv": lambda: ops.gemm(h, wqkv, bias=bq, out=qout),
    "ln_qkv": lambda: ops.ln_gemm(x, gam, bet, wqkv, bq, out=qout),
    "attention": lambda: ops.attention(qkv, B, 6),
    "proj_ln": lambda: ops.proj_ln(x, att, wproj, bp, gam, bet, out=hout),
    "mlp": lambda: ops.mlp_fused(x, h, w1, b1, w2, b2),
}
for name, fn in fns.items():
    if name not in which:
        continue
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        fn()
    b.record()
    torch.cuda.synchronize()
    print(f"{name:10s} {a.elapsed_time(b) / 20 * 1e3:8.1f} us", flush=True)
