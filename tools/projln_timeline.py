"""clock64 timeline of the fused projection + LayerNorm kernel's epilogue (CTA 0, warp 4, third tile)."""
import os, sys, torch
sys.path.insert(0, ".")
dbg = torch.zeros(64, dtype=torch.int64, device="cuda")
os.environ["EFFOCR_PLN_DBG_PTR"] = str(dbg.data_ptr())
from effocr_b200 import ops
M, D = 201728, 384
att = (torch.randn(M, D, device="cuda") * 0.7).half()
w = (torch.randn(D, D, device="cuda") * 0.05).half()
b = torch.randn(D, device="cuda"); g = torch.ones(D, device="cuda"); be = torch.zeros(D, device="cuda")
x = torch.randn(M, D, device="cuda"); h = torch.empty(M, D, device="cuda", dtype=torch.float16)
for _ in range(3):
    ops.proj_ln(x, att, w, b, g, be, out=h)
torch.cuda.synchronize()
t = dbg.cpu().tolist(); t0 = t[0]
names = {0: "wait tfull", 1: "tfull seen", 14: "pass1 done", 15: "stats merged", 22: "tile done"}
for k in range(3):
    names.update({2 + 4 * k: f"p1 c{k} ld issued", 3 + 4 * k: f"p1 c{k} x+acc ready", 4 + 4 * k: f"p1 c{k} computed", 5 + 4 * k: f"p1 c{k} quad synced"})
    names.update({16 + 2 * k: f"p2 c{k} loaded", 17 + 2 * k: f"p2 c{k} staging free"})
for i in sorted(names):
    print(f"{names[i]:24s} {t[i] - t0:8d}")
