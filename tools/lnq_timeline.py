"""clock64 timeline of the fused norm1 + QKV kernel (CTA 0): MMA issuer per tile, epilogue warp 4 per tile, LayerNorm
warp 0 per row block.  Prints, per row block, when each role waited and for how long (cycles)."""
import os, sys, torch
sys.path.insert(0, ".")
dbg = torch.zeros(3 * 1024, dtype=torch.int64, device="cuda")
from effocr_b200 import ops
M = 201728
x = torch.randn(M, 384, device="cuda")
w = (torch.randn(1152, 384, device="cuda") * 0.05).half()
b = torch.randn(1152, device="cuda"); g = torch.ones(384, device="cuda"); be = torch.zeros(384, device="cuda")
out = torch.empty(M, 1152, device="cuda", dtype=torch.float16)
for _ in range(2):
    ops.ln_gemm(x, g, be, w, bias=b, out=out)
torch.cuda.synchronize()
os.environ["EFFOCR_LNQ_DBG_PTR"] = str(dbg.data_ptr())
ops.ln_gemm(x, g, be, w, bias=b, out=out)
torch.cuda.synchronize()
t = dbg.cpu().view(3, 1024)
mma, epi, ln = t[0].tolist(), t[1].tolist(), t[2].tolist()
t0 = min(v for v in mma + epi + ln if v > 0)
NT = int(sys.argv[1]) if len(sys.argv) > 1 else 6
NT = int(sys.argv[1]) if len(sys.argv) > 1 else 6   # N tiles per row block (1152 / BLOCK_N)
nblk = 0
while mma[4 * NT * nblk] > 0:
    nblk += 1
print("blocks", nblk, "tiles per block", NT)
for bk in range(min(nblk, 6)):
    L = [v - t0 for v in ln[8 * bk: 8 * bk + 7]]
    if ln[8 * bk] > 0: print(f"block {bk}: LN ready {L[0]:8d} | pair0 released {L[1]:8d} handed {L[2]:8d} | pair1 rel {L[3]:8d} handed {L[4]:8d} | pair2 rel {L[5]:8d} handed {L[6]:8d}")
    for nb in range(NT):
        i = bk * NT + nb
        m = [v - t0 for v in mma[4 * i: 4 * i + 4]]
        e = [v - t0 for v in epi[3 * i: 3 * i + 3]]
        print(f"   tile {nb}: MMA wait-tmem {m[0]:8d} tmem-free {m[1]:8d} operands {m[2]:8d} issued {m[3]:8d} (tile {m[3]-m[0]:6d}) | EPI wait {e[0]:8d} full {e[1]:8d} drained {e[2]:8d} (drain {e[2]-e[1]:6d})")
