"""clock64 timeline of the one-kernel block tail (CTA 0): per tile, when the MMA issuer and epilogue warp 4 reached each phase."""
import os, sys, torch
sys.path.insert(0, ".")
dbg = torch.zeros(2 * 512, dtype=torch.int64, device="cuda")  # [role][tile][8] stamps; [1000..1006] MMA-issuer wait totals
from effocr_b200 import ops
M, D, HID = 201728, 384, 1536
x = torch.randn(M, D, device="cuda")
att = (torch.randn(M, D, device="cuda") * 0.7).half()
wp = (torch.randn(D, D, device="cuda") * 0.05).half()
w1 = (torch.randn(HID, D, device="cuda") * 0.05).half()
w2 = (torch.randn(D, HID, device="cuda") * 0.05).half()
bp, g, be, b2 = (torch.randn(D, device="cuda") for _ in range(4))
b1 = torch.randn(HID, device="cuda")
for _ in range(2):
    ops.block_tail(x, att, wp, bp, g, be, w1, b1, w2, b2)
torch.cuda.synchronize()
os.environ["EFFOCR_TAIL_DBG_PTR"] = str(dbg.data_ptr())
ops.block_tail(x, att, wp, bp, g, be, w1, b1, w2, b2)
torch.cuda.synchronize()
acc = dbg.cpu()[1000:1007].tolist()
dbg[1000:1007] = 0
t = dbg.cpu().view(2, 64, 8)
t0 = int(t[t > 0].min())
names_m = ["tile start", "O drained", "proj issued", "h ready", "last fc2 issued"]
names_e = ["tile start", "proj done", "pass1 done", "pass2 done", "first S", "last GELU", "fc2 done", "drained"]
for tile in range(1, 6):
    m = [int(v) - t0 for v in t[0, tile, :5]]
    e = [int(v) - t0 for v in t[1, tile]]
    print(f"tile {tile}  MMA: " + "  ".join(f"{n} {v}" for n, v in zip(names_m, m)))
    print(f"        EPI: " + "  ".join(f"{n} {v}" for n, v in zip(names_e, e)))
    print(f"        EPI phase lengths: wait-proj {e[1]-e[0]}  pass1 {e[2]-e[1]}  pass2 {e[3]-e[2]}  to-first-S {e[4]-e[3]}  gelu-loop {e[5]-e[4]}  to-fc2-done {e[6]-e[5]}  drain {e[7]-e[6]}  | tile {e[7]-e[0]}")
if acc[6]:
    n = acc[6]
    print("MMA issuer, cycles waited per tile on:  O drained %d  att landed %d  weights %d  h %d  S buffer free %d  P written %d"
          % tuple(a // n for a in acc[:6]))
