import sys, torch, json
sys.path.insert(0, ".")
from effocr_b200 import ops
qkv = (torch.randn(1024 * 197, 1152, device="cuda") * 0.3).half()
for impl in (0, 1):
    for _ in range(3): ops.attention(qkv, 1024, 6, impl=impl)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.attention(qkv, 1024, 6, impl=impl)
    e1.record(); torch.cuda.synchronize()
    print(json.dumps(dict(impl=impl, ms=e0.elapsed_time(e1) / 20)))
