"""A/B timing of the attention kernels on one box: impl 2 = tcgen05 two-pass softmax, 3 = tcgen05 single TMEM pass,
1 = first-generation mma.sync kernel.  Prints ms per launch (batch 1024 x 6 heads x 197 tokens) and max |diff|."""
import json
import sys

import torch

sys.path.insert(0, ".")
from effocr_b200 import ops

qkv = (torch.randn(1024 * 197, 1152, device="cuda") * 0.3).half()
ref = ops.attention(qkv, 1024, 6, impl=1).float()
for impl in (0, 2, 4, 0, 2):
    for _ in range(3):
        out = ops.attention(qkv, 1024, 6, impl=impl)
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        ops.attention(qkv, 1024, 6, impl=impl)
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps(dict(impl=impl, ms=e0.elapsed_time(e1) / 50, maxdiff=float((out.float() - ref).abs().max()))))
