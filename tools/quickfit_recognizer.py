"""TEST INFRASTRUCTURE -- quick-fit a small recogniser so that transcription parity is meaningful.

With random-init encoders every glyph embedding has cosine >= 0.97 with every other and the top-1/top-2 margin is
~1e-5, so "identical transcriptions" cannot be asserted.  This script trains a 2-block
ViT (D = 192, timm key names, the architecture of oracle/vit.py) for a few hundred steps on rendered ASCII glyphs
with a cosine-softmax head (the reference trains with a supervised-contrastive loss,
/root/reference/train_effocr_recognizer.py:126-157; any loss that separates glyph classes serves the purpose) and
saves the weights as tests/golden/quickfit_vit_d2.npz (fp16).  Run on the GPU box:

    gpurun -- 'python tools/quickfit_recognizer.py gpurun_out/quickfit_vit_d2.npz'

Training uses torch autograd over the ORACLE forward (oracle/vit.py) -- it is not part of the product.
"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from effocr_b200 import synth
from oracle import transform as OT, vit as OV

out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/quickfit_vit_d2.npz"
dev = "cuda" if torch.cuda.is_available() else "cpu"
torch.manual_seed(0)
rng = np.random.default_rng(0)
glyphs = synth.ASCII_GLYPHS

# ---- data: character crops cut from rendered lines exactly like the pipeline cuts them (same distribution as the
# held-out lines: font sizes 34..43, arbitrary sub-pixel positions, neighbouring-glyph bleed), plus isolated renders
crops, labels = synth.synthetic_crops(9000, seed=1)
labels = [glyphs.index(c) for c in labels]
for gi, ch in enumerate(glyphs):
    for size in (34, 37, 40, 43):
        img, cb, _wb, chars = synth.render_line(ch, font_size=size, x0=6, width=128)
        if len(cb) == 1:
            a, b = int(round(float(cb[0][0]))), int(round(float(cb[0][2])))
            if b > a:
                crops.append(np.ascontiguousarray(img[:, a:b, :]))
                labels.append(gi)
X = torch.from_numpy(np.stack([OT.paired_transform(c) for c in crops])).half()
Y = torch.tensor(labels)
print(f"{len(crops)} training crops of {len(glyphs)} glyphs", flush=True)

# ---- model: timm-keyed ViT, D = 192, 3 heads, 2 blocks, mlp 768
full = OV.init_vit_state_dict("vit_tiny_patch16_224", seed=0)
sd = {k: v.clone().to(dev).requires_grad_(True) for k, v in full.items()
      if not k.startswith("net.blocks.") or int(k.split(".")[2]) < 2}
head = (torch.randn(len(glyphs), 192, device=dev) * 0.02).requires_grad_(True)
opt = torch.optim.AdamW(list(sd.values()) + [head], lr=1e-3, weight_decay=0.01)
X, Y = X.to(dev), Y.to(dev)
steps, bs = 900, 192
t0 = time.time()
for step in range(steps):
    idx = torch.randint(0, len(X), (bs,), device=dev)
    xb = X[idx].float() + 0.05 * torch.randn(bs, 3, 224, 224, device=dev)
    emb = torch.nn.functional.normalize(OV.vit_forward(sd, xb), dim=1)
    logits = emb @ torch.nn.functional.normalize(head, dim=1).t() / 0.07
    loss = torch.nn.functional.cross_entropy(logits, Y[idx])
    opt.zero_grad(set_to_none=True)
    loss.backward()
    for g in opt.param_groups:
        g["lr"] = 1e-3 * min(1.0, (step + 1) / 20) * (0.5 * (1 + np.cos(np.pi * step / steps)))
    opt.step()
    if step % 100 == 0 or step == steps - 1:
        acc = (logits.argmax(1) == Y[idx]).float().mean().item()
        print(f"step {step} loss {loss.item():.4f} acc {acc:.3f} ({time.time() - t0:.0f}s)", flush=True)

# ---- evaluate retrieval on held-out line crops: index = one canonical render per glyph
final = {k: v.detach().cpu() for k, v in sd.items()}
with torch.no_grad():
    proto = [synth.render_line(ch, font_size=40, x0=6, width=128) for ch in glyphs]
    pc = [np.ascontiguousarray(im[:, int(round(float(cb[0][0]))):int(round(float(cb[0][2]))), :]) for im, cb, _, _ in proto]
    xb = OV.l2_normalize(OV.vit_forward(final, torch.from_numpy(np.stack([OT.paired_transform(c) for c in pc]))))
    test_crops, test_labels = synth.synthetic_crops(600, seed=123)
    q = OV.l2_normalize(OV.vit_forward(final, torch.from_numpy(np.stack([OT.paired_transform(c) for c in test_crops]))))
    s = q @ xb.t()
    top2 = torch.topk(s, 2, dim=1)
    pred = [glyphs[i] for i in top2.indices[:, 0].tolist()]
    acc = np.mean([p == t for p, t in zip(pred, test_labels)])
    margin = (top2.values[:, 0] - top2.values[:, 1])
    print(f"held-out top-1 accuracy {acc:.3f}; margin min {margin.min().item():.4f} median {margin.median().item():.4f}", flush=True)
np.savez_compressed(out_path, **{k: v.numpy().astype(np.float16) for k, v in final.items()})
print("saved", out_path, flush=True)
