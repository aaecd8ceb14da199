"""Where the default attention kernel (attention_tc2b_kernel) waits: cycle totals of every mbarrier wait of the MMA-issuing
warp and of one softmax warp per query-tile group of CTA 0, over one launch (batch 1024 x 6 heads x 197 tokens).
usage (GPU box): PYTHONPATH=. python tools/att_timeline.py [impl: 0 two key blocks / 8 softmax warps, 5 three blocks / 16 warps]"""
import os
import sys

import torch

sys.path.insert(0, ".")
from effocr_b200 import ops

IMPL = int(sys.argv[1]) if len(sys.argv) > 1 else 0
qkv = (torch.randn(1024 * 197, 1152, device="cuda") * 0.3).half()
for _ in range(3):
    ops.attention(qkv, 1024, 6, impl=IMPL)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.attention(qkv, 1024, 6, impl=IMPL)
e1.record()
torch.cuda.synchronize()
print("impl %d: %.1f us per launch" % (IMPL,  e0.elapsed_time(e1) / 20 * 1e3))
dbg = torch.zeros(32, dtype=torch.int64, device="cuda")
os.environ["EFFOCR_ATT_DBG_PTR"] = str(dbg.data_ptr())
ops.attention(qkv, 1024, 6, impl=IMPL)
torch.cuda.synchronize()
d = dbg.tolist()
n = max(d[11], 1)
names = ["k_full", "q_full[0]", "t_free[0]", "p_full[0]", "pb_full[0]", "q_full[1]", "t_free[1]", "p_full[1]", "pb_full[1]", "v_full"]
print("MMA warp, CTA 0: %d units, %.0f cycles per unit" % (d[11], d[10] / n))
for i, nm in enumerate(names):
    print("  wait %-11s %7.0f cycles per unit" % (nm, d[i] / n))
for g in range(4 if IMPL == 5 else 2):
    b = 12 + 4 * g
    print("softmax warp %s group %d: s_full wait %.0f, softmax %.0f, o_full wait %.0f, epilogue %.0f cycles per unit"
          % ("XY"[g >> 1] if IMPL == 5 else "", g & 1 if IMPL == 5 else g, d[b] / n, d[b + 1] / n, d[b + 2] / n, d[b + 3] / n))
