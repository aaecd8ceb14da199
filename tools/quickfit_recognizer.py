"""TEST INFRASTRUCTURE -- quick-fit a recogniser so that transcription parity is meaningful.

With random-init encoders every glyph embedding has cosine >= 0.97 with every other and the top-1/top-2 margin is
~1e-5, so "identical transcriptions" cannot be asserted.  This script trains a timm-keyed ViT
(the architecture of oracle/vit.py) for a few hundred steps on rendered ASCII glyphs with a cosine-softmax head (the
reference trains with a supervised-contrastive loss, /root/reference/train_effocr_recognizer.py:126-157; any loss that
separates glyph classes serves the purpose) and saves the weights as fp16 .npz.  Run on the GPU box:

    gpurun -- 'python tools/quickfit_recognizer.py gpurun_out/quickfit_vit_d2.npz'                       (2 blocks, D = 192)
    gpurun -- 'python tools/quickfit_recognizer.py gpurun_out/quickfit_vit_small.npz --arch vit_small_patch16_224
               --depth 12 --steps 1500 --composites 2000 --index 10000'                                  (the headline model)

`--composites N`: the first N composite glyphs of synth.glyph_image (the stand-in for a large CJK charset that
bench.py's 10 000-glyph index is made of) join the label set as one class each, so that the index prototypes spread out
and character queries keep wide top-1 / top-2 margins against the full index.

Training uses torch autograd over the ORACLE forward (oracle/vit.py) -- it is not part of the product; the training
inputs come from the product's crop kernel only because the oracle's per-crop transform is a slow Python loop.
"""
import argparse
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from effocr_b200 import synth
from oracle import transform as OT, vit as OV

ap = argparse.ArgumentParser()
ap.add_argument("out", nargs="?", default="gpurun_out/quickfit_vit_d2.npz")
ap.add_argument("--arch", default="vit_tiny_patch16_224")
ap.add_argument("--depth", type=int, default=2)
ap.add_argument("--steps", type=int, default=900)
ap.add_argument("--batch", type=int, default=192)
ap.add_argument("--lr", type=float, default=1e-3)
ap.add_argument("--composites", type=int, default=0)
ap.add_argument("--index", type=int, default=0, help="evaluate margins against this many synth.glyph_images prototypes")
ap.add_argument("--crops", type=int, default=9000)
args = ap.parse_args()

dev = "cuda" if torch.cuda.is_available() else "cpu"
torch.manual_seed(0)
glyphs = synth.ASCII_GLYPHS


def transform_all(crops):
    """u8 crops -> f16 [n, 3, 224, 224] through the product crop kernel on the GPU (oracle transform on the CPU)."""
    if dev == "cuda":
        from effocr_b200 import ops
        from effocr_b200.pipeline import PackedCrops

        outs = []
        for i0 in range(0, len(crops), 2048):
            px, im, bx, n = PackedCrops(crops[i0:i0 + 2048]).to_device()
            outs.append(ops.crop_resize(px, im, bx, n, ops.CROP_NCHW_F16))
        return torch.cat(outs)
    return torch.from_numpy(np.stack([OT.paired_transform(c) for c in crops])).half()


# ---- data: character crops cut from rendered lines exactly like the pipeline cuts them (same distribution as the
# held-out lines: font sizes 34..43, arbitrary sub-pixel positions, neighbouring-glyph bleed), plus isolated renders
crops, labels = synth.synthetic_crops(args.crops, seed=1)
labels = [glyphs.index(c) for c in labels]
for gi, ch in enumerate(glyphs):
    for size in (34, 37, 40, 43):
        img, cb, _wb, chars = synth.render_line(ch, font_size=size, x0=6, width=128)
        if len(cb) == 1:
            a, b = int(round(float(cb[0][0]))), int(round(float(cb[0][2])))
            if b > a:
                crops.append(np.ascontiguousarray(img[:, a:b, :]))
                labels.append(gi)
for gi in range(len(glyphs)):  # the index prototypes of bench.py (synth.glyph_image) are 64 x 64 renders at size 44
    crops.append(synth.glyph_image(gi))
    labels.append(gi)
n_classes = len(glyphs) + args.composites
for ci in range(args.composites):
    crops.append(synth.glyph_image(len(glyphs) + ci))
    labels.append(len(glyphs) + ci)
X = transform_all(crops).to(dev)
Y = torch.tensor(labels, device=dev)
n_line = len(crops) - args.composites
print(f"{len(crops)} training crops, {n_classes} classes ({args.composites} composite glyphs)", flush=True)

# ---- model: timm-keyed ViT, first `depth` blocks of the architecture
full = OV.init_vit_state_dict(args.arch, seed=0)
sd = {k: v.clone().to(dev) for k, v in full.items() if not k.startswith("net.blocks.") or int(k.split(".")[2]) < args.depth}
for k in sd:  # residual-branch outputs scaled down with depth (GPT-2 style): a 12-block stack trains without warm-up tricks
    if k.endswith(("attn.proj.weight", "mlp.fc2.weight")):
        sd[k] = sd[k] / (2.0 * args.depth) ** 0.5
sd = {k: v.requires_grad_(True) for k, v in sd.items()}
D = OV.VIT_CONFIGS[args.arch][0]
head = (torch.randn(n_classes, D, device=dev) * 0.02).requires_grad_(True)
decay = [v for k, v in sd.items() if v.dim() > 1] + [head]
no_decay = [v for k, v in sd.items() if v.dim() <= 1]
opt = torch.optim.AdamW([{"params": decay, "weight_decay": 0.01}, {"params": no_decay, "weight_decay": 0.0}], lr=args.lr)
steps, bs = args.steps, args.batch
warm = max(20, steps // 15)
t0 = time.time()
amp = torch.autocast("cuda", dtype=torch.bfloat16) if dev == "cuda" else torch.autocast("cpu", enabled=False)
for step in range(steps):
    n_comp = bs // 4 if args.composites else 0
    idx = torch.randint(0, n_line, (bs - n_comp,), device=dev)
    if n_comp:
        idx = torch.cat([idx, n_line + torch.randint(0, args.composites, (n_comp,), device=dev)])
    xb = X[idx].float() + 0.05 * torch.randn(bs, 3, 224, 224, device=dev)
    sh = torch.randint(-6, 7, (2,))  # small translations: crops are cut at arbitrary sub-pixel offsets
    xb = torch.roll(xb, shifts=(int(sh[0]), int(sh[1])), dims=(2, 3))
    with amp:
        feat = OV.vit_forward(sd, xb)
    emb = torch.nn.functional.normalize(feat.float(), dim=1)
    logits = emb @ torch.nn.functional.normalize(head, dim=1).t() / 0.07
    loss = torch.nn.functional.cross_entropy(logits, Y[idx])
    opt.zero_grad(set_to_none=True)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(list(sd.values()) + [head], 1.0)
    for g in opt.param_groups:
        g["lr"] = args.lr * min(1.0, (step + 1) / warm) * (0.5 * (1 + np.cos(np.pi * step / steps)))
    opt.step()
    if step % 100 == 0 or step == steps - 1:
        acc = (logits.argmax(1) == Y[idx]).float().mean().item()
        print(f"step {step} loss {loss.item():.4f} acc {acc:.3f} ({time.time() - t0:.0f}s)", flush=True)

# ---- evaluate retrieval on held-out line crops (fp32 oracle forward over the fp16-rounded weights that are saved)
final = {k: v.detach().half().float() for k, v in sd.items()}


def embed(crop_list):
    out = []
    with torch.no_grad():
        x = transform_all(crop_list).to(dev)
        for i0 in range(0, len(x), 512):
            out.append(OV.l2_normalize(OV.vit_forward(final, x[i0:i0 + 512].float())))
    return torch.cat(out)


proto = [synth.render_line(ch, font_size=40, x0=6, width=128) for ch in glyphs]
pc = [np.ascontiguousarray(im[:, int(round(float(cb[0][0]))):int(round(float(cb[0][2]))), :]) for im, cb, _, _ in proto]
test_crops, test_labels = synth.synthetic_crops(1024, seed=123)
q = embed(test_crops)
for name, xb_crops in (("94 canonical renders", pc),) + ((("%d synth.glyph_images" % args.index, synth.glyph_images(args.index)),) if args.index else ()):
    xb = embed(xb_crops)
    s = q.double() @ xb.double().t()
    top2 = torch.topk(s, 2, dim=1)
    pred = top2.indices[:, 0].tolist()
    acc = np.mean([p < len(glyphs) and glyphs[p] == t for p, t in zip(pred, test_labels)])
    margin = (top2.values[:, 0] - top2.values[:, 1])
    print(f"index = {name}: held-out top-1 accuracy {acc:.3f}; margin min {margin.min().item():.4f} "
          f"p05 {margin.quantile(0.05).item():.4f} median {margin.median().item():.4f}; "
          f"frac(margin > 4e-3) {(margin > 4e-3).double().mean().item():.4f}", flush=True)
np.savez_compressed(args.out, **{k: v.cpu().numpy().astype(np.float16) for k, v in final.items()})
print("saved", args.out, flush=True)
