"""One ConvNeXt-Tiny forward at batch 1024 for an ncu launch list (per-stage kernel times)."""
import sys
sys.path.insert(0, ".")
import torch
from effocr_b200.engine import ConvNextEngine
from oracle import convnext as OC
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
eng = ConvNextEngine(OC.init_convnext_tiny_state_dict(seed=0), max_batch=B)
x = torch.randn(B, 3, 224, 224, device="cuda")
for _ in range(2):
    eng.forward(x)
torch.cuda.synchronize()
