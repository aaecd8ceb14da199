"""Timing ablations of the default attention kernel (results are wrong on purpose): which part of the softmax warps' work sets
the period.  EFFOCR_ATT_ABLATE bits: 1 no global stores, 2 no P stores, 4 no MUFU (FMA instead of ex2), 8 no maximum search.
The ablated kernels are not part of the product library: build it first with
  EFFOCR_NVCC_EXTRA=-DEFFOCR_ATT_ABLATION python -m effocr_b200.build --force   (and rebuild without it afterwards)
usage (GPU box): PYTHONPATH=. python tools/att_ablate.py"""
import os
import sys

import torch

sys.path.insert(0, ".")
from effocr_b200 import ops

qkv = (torch.randn(1024 * 197, 1152, device="cuda") * 0.3).half()
for mode in (0, 1, 2, 4, 8, 3, 7, 15, 0):
    os.environ["EFFOCR_ATT_ABLATE"] = str(mode)
    for _ in range(3):
        ops.attention(qkv, 1024, 6, impl=0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.attention(qkv, 1024, 6, impl=0)
    e1.record()
    torch.cuda.synchronize()
    print("ablate %2d: %.1f us per launch" % (mode, e0.elapsed_time(e1) / 20 * 1e3), flush=True)
