"""GPU self-test + timing of the fused MLP block kernel against torch fp32 and the unfused fc1/fc2 GEMM pair."""
import sys
import torch
sys.path.insert(0, ".")
from effocr_b200 import ops


def ref_update(h, w1, b1, w2, b2):
    p = torch.nn.functional.gelu(h.float() @ w1.float().t() + b1).half().float()
    return p @ w2.float().t() + b2


def check(M, D, HID, seed=0):
    torch.manual_seed(seed)
    h = (torch.randn(M, D, device="cuda") * 0.7).half()
    w1 = (torch.randn(HID, D, device="cuda") * 0.05).half()
    w2 = (torch.randn(D, HID, device="cuda") * 0.05).half()
    b1 = torch.randn(HID, device="cuda") * 0.3
    b2 = torch.randn(D, device="cuda") * 0.3
    x0 = torch.randn(M, D, device="cuda")
    x = x0.clone()
    ops.mlp_fused(x, h, w1, b1, w2, b2)
    torch.cuda.synchronize()
    upd = ref_update(h, w1, b1, w2, b2)
    rel = ((x - x0) - upd).norm() / upd.norm()
    mx = ((x - x0) - upd).abs().max()
    print(f"M={M} D={D} HID={HID}: rel {rel:.3e} max-abs {mx:.3e}", flush=True)
    return rel.item()


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


if __name__ == "__main__":
    bad = 0
    shapes = [] if "--time-only" in sys.argv else  [(256, 384, 1536), (100, 384, 1536), (300, 384, 256), (5000, 384, 1536), (40000, 384, 1536),
                        (70001, 384, 1536), (777, 192, 768), (50000, 192, 768), (201728, 384, 1536)]
    for (M, D, HID) in shapes:
        bad += check(M, D, HID) > 2e-4
    print("FAILED" if bad else "all ok", flush=True)
    M, D, HID = 201728, 384, 1536
    h = (torch.randn(M, D, device="cuda") * 0.7).half()
    w1 = (torch.randn(HID, D, device="cuda") * 0.05).half()
    w2 = (torch.randn(D, HID, device="cuda") * 0.05).half()
    b1 = torch.randn(HID, device="cuda"); b2 = torch.randn(D, device="cuda")
    x = torch.randn(M, D, device="cuda")
    mid = torch.empty(M, HID, device="cuda", dtype=torch.float16)
    t_f = timeit(lambda: ops.mlp_fused(x, h, w1, b1, w2, b2))

    def unfused():
        ops.gemm(h, w1, bias=b1, act=1, out=mid)
        ops.gemm(mid, w2, bias=b2, out_dtype=torch.float32, resid=x, out=x)
    t_u = timeit(unfused)
    fl = 4.0 * M * D * HID
    print(f"fused {t_f*1e3:.1f} us ({fl/t_f/1e9:.0f} TFLOP/s)   unfused fc1+fc2 {t_u*1e3:.1f} us ({fl/t_u/1e9:.0f} TFLOP/s)")
