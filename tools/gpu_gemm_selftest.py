"""First-light test of the tcgen05 GEMM on a real B200 (run through gpurun)."""
import sys, time, json
import torch
sys.path.insert(0, ".")
from effocr_b200 import _lib

lib = _lib.load()
_lib.require_device()
torch.manual_seed(0)
dev = "cuda"
results = []

def run(M, N, K, act=0, out_f32=0, bias=True, gamma=False, resid=False, bn=0, lda_pad=0):
    A = (torch.randn(M, K + lda_pad, device=dev) * 0.5).half()
    W = (torch.randn(N, K, device=dev) * 0.05).half()
    b = torch.randn(N, device=dev) if bias else None
    g = torch.randn(N, device=dev) if gamma else None
    odt = torch.float32 if out_f32 else torch.float16
    R = torch.randn(M, N, device=dev).to(odt) if resid else None
    ldo = (N + 7) // 8 * 8
    out_full = torch.full((M, ldo), float("nan"), device=dev, dtype=odt)
    out = out_full[:, :N]
    if resid:
        R_full = torch.randn(M, ldo, device=dev).to(odt); R = R_full[:, :N]
    Av = A[:, :K]
    st = lib.effocr_gemm_f16(A.data_ptr(), A.stride(0), W.data_ptr(), W.stride(0), M, N, K,
                             _lib.ptr(b), _lib.ptr(g), _lib.ptr(R), ldo, out.data_ptr(), ldo,
                             act, out_f32, bn, _lib.stream_ptr())
    _lib.check(st, "gemm")
    torch.cuda.synchronize()
    ref = Av.float() @ W.float().t()
    if bias: ref = ref + b
    if act == 1: ref = torch.nn.functional.gelu(ref)
    if act == 2: ref = torch.nn.functional.silu(ref)
    if gamma: ref = ref * g
    if resid: ref = ref + R.float()
    err = (out.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    nan = int(torch.isnan(out.float()).sum().item())
    ok = nan == 0 and err <= (2e-3 * scale + 2e-3 if not out_f32 else 1e-3 * scale + 1e-4)
    rec = dict(M=M, N=N, K=K, act=act, f32=out_f32, gamma=gamma, resid=resid, bn=bn, err=err, scale=scale, nan=nan, ok=ok)
    print(json.dumps(rec), flush=True)
    results.append(rec)
    return ok

cases = [
    dict(M=128, N=64, K=64, bn=64),
    dict(M=128, N=128, K=64, bn=128),
    dict(M=128, N=256, K=128, bn=256),
    dict(M=128, N=192, K=384, bn=192),
    dict(M=256, N=384, K=384),
    dict(M=1000, N=384, K=384),
    dict(M=5000, N=1152, K=384),
    dict(M=197 * 64, N=1536, K=384, act=1),
    dict(M=197 * 64, N=384, K=1536, out_f32=1, resid=True),
    dict(M=3000, N=100, K=72, act=2),
    dict(M=3000, N=21, K=512, out_f32=1, bn=64),
    dict(M=4000, N=256, K=288, act=2, resid=True),
    dict(M=777, N=384, K=384, out_f32=1, gamma=True, resid=True),
    dict(M=333, N=200, K=40, lda_pad=8),
    dict(M=201728, N=1152, K=384),
]
allok = True
for c in cases:
    try:
        allok &= run(**c)
    except Exception as e:
        print("EXC", c, repr(e), flush=True)
        allok = False
        break

# quick perf probe: TMA-store epilogue vs direct-store epilogue
if allok:
    for (M, N, K, act, f32, res) in [(201728, 1152, 384, 0, 0, False), (201728, 1536, 384, 1, 0, False),
                                      (201728, 1536, 384, 0, 0, False),
                                      (201728, 384, 1536, 0, 1, True), (201728, 384, 384, 0, 1, True)]:
        A = (torch.randn(M, K, device=dev) * 0.5).half(); W = (torch.randn(N, K, device=dev) * 0.05).half()
        b = torch.randn(N, device=dev); odt = torch.float32 if f32 else torch.float16
        out = torch.zeros(M, N, device=dev, dtype=odt)
        R = out if res else None
        for direct in (0, 0x10000):
            def go():
                _lib.check(lib.effocr_gemm_f16(A.data_ptr(), K, W.data_ptr(), K, M, N, K, b.data_ptr(), 0, _lib.ptr(R), N,
                                               out.data_ptr(), N, act, f32, direct, _lib.stream_ptr()))
            for _ in range(3): go()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): go()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(json.dumps(dict(perf=True, M=M, N=N, K=K, act=act, f32=f32, direct=bool(direct), ms=ms, tflops=2.0 * M * N * K / ms / 1e9)), flush=True)
print("ALL_OK" if allok else "FAILED", flush=True)
# A-stationary vs plain TMA-epilogue schedule
for (M, N, K, act, f32, res) in [(201728, 1152, 384, 0, 0, False), (201728, 1536, 384, 1, 0, False), (201728, 384, 384, 0, 1, True)]:
    A = (torch.randn(M, K, device=dev) * 0.5).half(); W = (torch.randn(N, K, device=dev) * 0.05).half()
    b = torch.randn(N, device=dev); odt = torch.float32 if f32 else torch.float16
    out = torch.zeros(M, N, device=dev, dtype=odt); R = out if res else None
    for flag in (0, 0x20000):
        def go():
            _lib.check(lib.effocr_gemm_f16(A.data_ptr(), K, W.data_ptr(), K, M, N, K, b.data_ptr(), 0, _lib.ptr(R), N,
                                           out.data_ptr(), N, act, f32, flag, _lib.stream_ptr()))
        for _ in range(3): go()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): go()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(json.dumps(dict(perf2=True, M=M, N=N, K=K, act=act, f32=f32, astat=(flag == 0), ms=ms, tflops=2.0 * M * N * K / ms / 1e9)), flush=True)
