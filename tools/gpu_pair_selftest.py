"""CTA-pair (cta_group::2) GEMM bring-up: correctness vs torch + timing vs the single-CTA schedules."""
import json, sys
import torch
sys.path.insert(0, ".")
from effocr_b200 import _lib
lib = _lib.load(); _lib.require_device()
dev = "cuda"; torch.manual_seed(0)
PAIR, PLAIN = 0x30000, 0x20000

def run(M, N, K, act, f32, res, flag, check=True, reps=0):
    A = (torch.randn(M, K, device=dev) * 0.5).half(); W = (torch.randn(N, K, device=dev) * 0.05).half()
    b = torch.randn(N, device=dev); odt = torch.float32 if f32 else torch.float16
    out = torch.randn(M, N, device=dev).to(odt) if res else torch.full((M, N), float("nan"), device=dev, dtype=odt)
    base = out.clone() if res else None
    R = out if res else None
    def go():
        _lib.check(lib.effocr_gemm_f16(A.data_ptr(), K, W.data_ptr(), K, M, N, K, b.data_ptr(), 0, _lib.ptr(R), N,
                                       out.data_ptr(), N, act, f32, flag, _lib.stream_ptr()))
    go(); torch.cuda.synchronize()
    rec = dict(M=M, N=N, K=K, act=act, f32=f32, res=res, pair=(flag == PAIR))
    if check:
        ref = A.float() @ W.float().t() + b
        if act == 1: ref = torch.nn.functional.gelu(ref)
        if res: ref = ref + base.float()
        rel = ((out.float() - ref).norm() / ref.norm()).item()
        rec.update(rel=rel, nan=int(torch.isnan(out.float()).sum()), ok=bool(rel < (5e-6 if f32 else 6e-4)))
    if reps:
        for _ in range(3): go()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): go()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        rec.update(ms=ms, tflops=2.0 * M * N * K / ms / 1e9)
    print(json.dumps(rec), flush=True)
    return rec.get("ok", True)

ok = True
for (M, N, K, act, f32, res) in [(256 * 80, 256, 64, 0, 0, False), (256 * 80, 192, 128, 0, 0, False), (30000, 1152, 384, 0, 0, False),
                                  (25000, 1536, 384, 1, 0, False), (20000, 384, 1536, 0, 1, True), (19999, 600, 320, 0, 0, False)]:
    ok &= run(M, N, K, act, f32, res, PAIR)
if ok:
    for (M, N, K, act, f32, res) in [(201728, 1152, 384, 0, 0, False), (201728, 1536, 384, 1, 0, False), (201728, 1536, 384, 0, 0, False),
                                      (201728, 384, 1536, 0, 1, True), (201728, 384, 384, 0, 1, True)]:
        for flag in (PAIR, 0):
            run(M, N, K, act, f32, res, flag, check=False, reps=10)
print("PAIR_OK" if ok else "PAIR_FAILED")
