"""Experiment: ViT-S forward of 1024 crops as ONE launch sequence on all SMs vs TWO half-batches on two streams, every
persistent kernel capped at half of the SMs (EFFOCR_SM_LIMIT=74), so that HBM-bound kernels of one half can overlap
tensor-bound kernels of the other.  Usage: python tools/overlap_probe.py  (runs itself twice through the env var)."""
import os, subprocess, sys, time
sys.path.insert(0, ".")
if len(sys.argv) == 1:
    for lim, mode in (("0", "single"), ("74", "dual"), ("0", "dual"), ("74", "single")):
        env = dict(os.environ, EFFOCR_SM_LIMIT=lim)
        subprocess.run([sys.executable, __file__, mode], env=env)
    sys.exit(0)
import torch
from effocr_b200.engine import VitEngine
from effocr_b200.encoders import TimmViTParams
mode = sys.argv[1]
torch.manual_seed(0)
sd = {"net." + k: v for k, v in TimmViTParams("vit_small_patch16_224").state_dict().items()}
B = 1024
x = torch.randn(B * 196, 768, device="cuda", dtype=torch.float16)
if mode == "single":
    eng = VitEngine(sd, max_batch=B)
    def step():
        eng.forward(x)
else:
    e1, e2 = VitEngine(sd, max_batch=B // 2), VitEngine(sd, max_batch=B // 2)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    xa, xb = x[:B // 2 * 196], x[B // 2 * 196:]
    def step():
        main = torch.cuda.current_stream()
        s1.wait_stream(main); s2.wait_stream(main)
        with torch.cuda.stream(s1):
            e1.forward(xa)
        with torch.cuda.stream(s2):
            e2.forward(xb)
        main.wait_stream(s1); main.wait_stream(s2)
for _ in range(3):
    step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    step()
b.record()
torch.cuda.synchronize()
print(f"mode={mode} sm_limit={os.environ.get('EFFOCR_SM_LIMIT')}: {a.elapsed_time(b) / 10:.3f} ms per 1024 crops", flush=True)
