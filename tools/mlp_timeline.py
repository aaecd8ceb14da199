"""Pipeline hand-off timeline of the fused MLP kernel (CTA 0, third tile): clock64 stamps logged by the kernel itself."""
import os
import sys
import torch
sys.path.insert(0, ".")
dbg = torch.zeros(512, dtype=torch.int64, device="cuda")
os.environ["EFFOCR_MLP_DBG_PTR"] = str(dbg.data_ptr())
from effocr_b200 import ops
M, D, HID = 201728, 384, 1536
torch.manual_seed(0)
h = (torch.randn(M, D, device="cuda") * 0.7).half()
w1 = (torch.randn(HID, D, device="cuda") * 0.05).half()
w2 = (torch.randn(D, HID, device="cuda") * 0.05).half()
b1 = torch.randn(HID, device="cuda"); b2 = torch.randn(D, device="cuda")
x = torch.randn(M, D, device="cuda")
for _ in range(3):
    ops.mlp_fused(x, h, w1, b1, w2, b2)
torch.cuda.synchronize()
t = dbg.cpu().view(8, 64)
t0 = int(t[0, 0])
names = ["mma: G1 waits done", "mma: G2 waits done", "epi: sfull seen", "epi: GELU done", "epi: pempty seen", "epi: pfull arrived"]
print("chunk " + " ".join(f"{n[:18]:>19s}" for n in names))
for j in range(24):
    print(f"{j:5d} " + " ".join(f"{int(t[k, j]) - t0:19d}" for k in range(6)))
print("tile end: before ofull wait", int(t[6, 0]) - t0, " ofull seen", int(t[6, 1]) - t0, " drain done", int(t[6, 2]) - t0)
