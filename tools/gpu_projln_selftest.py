"""GPU self-test + timing of the fused projection + residual + LayerNorm kernel vs torch fp32 and the unfused pair."""
import sys
import torch
sys.path.insert(0, ".")
from effocr_b200 import ops


def check(M, seed=0):
    D = 384
    torch.manual_seed(seed)
    att = (torch.randn(M, D, device="cuda") * 0.7).half()
    w = (torch.randn(D, D, device="cuda") * 0.05).half()
    b = torch.randn(D, device="cuda") * 0.3
    g = 1.0 + 0.2 * torch.randn(D, device="cuda")
    be = 0.2 * torch.randn(D, device="cuda")
    x0 = torch.randn(M, D, device="cuda") * 2.0 + 0.5
    x = x0.clone()
    h = ops.proj_ln(x, att, w, b, g, be)
    torch.cuda.synchronize()
    xr = x0 + (att.float() @ w.float().t() + b)
    hr = torch.nn.functional.layer_norm(xr, (D,), g, be, 1e-6)
    ex = ((x - xr).norm() / xr.norm()).item()
    eh = ((h.float() - hr).norm() / hr.norm()).item()
    print(f"M={M}: x rel {ex:.3e}  h rel {eh:.3e}  finite {bool(torch.isfinite(h.float()).all())}", flush=True)
    return ex < 2e-6 and eh < 4e-4


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
    ok = all(check(M) for M in (256, 100, 5000, 40000, 70001, 201728))
    print("all ok" if ok else "FAILED", flush=True)
    M, D = 201728, 384
    att = (torch.randn(M, D, device="cuda") * 0.7).half()
    w = (torch.randn(D, D, device="cuda") * 0.05).half()
    b = torch.randn(D, device="cuda"); g = torch.ones(D, device="cuda"); be = torch.zeros(D, device="cuda")
    x = torch.randn(M, D, device="cuda")
    h = torch.empty(M, D, device="cuda", dtype=torch.float16)
    t_f = timeit(lambda: ops.proj_ln(x, att, w, b, g, be, out=h))

    def unfused():
        ops.gemm(att, w, bias=b, out_dtype=torch.float32, resid=x, out=x)
        ops.layernorm(x, g, be)
    t_u = timeit(unfused)
    print(f"fused {t_f*1e3:.1f} us ({930e6/t_f/1e6:.0f} GB/s algorithmic)   unfused proj + LN {t_u*1e3:.1f} us")
