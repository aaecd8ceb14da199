"""TEST INFRASTRUCTURE -- quick-fit a YOLOv5s localizer (nc = 2: character, word) on synthetic text lines.

A random-init detector gives every location nearly the same confidence, so which boxes survive `conf_thres` is decided
by rounding noise and "identical boxes / identical strings through the localizer" cannot be asserted.  This script
trains the oracle's YOLOv5s (oracle/yolo.py, ultralytics-keyed state dict) with the published YOLOv5 loss (anchor-ratio
target assignment with the two nearest neighbour cells, CIoU box loss, IoU-weighted objectness BCE, class BCE) on
rendered 64 x 1024 lines -- in the reference's 640 x 640 letterbox geometry AND at the native 64 x 1024 shape -- and
saves fp16 weights (BatchNorm statistics included) as .npz.  Run on the GPU box:

    gpurun -- 'python tools/quickfit_localizer.py gpurun_out/quickfit_yolov5s.npz --steps 2500'

Training runs torch autograd over the ORACLE forward; it is not part of the product.
"""
import argparse
import math
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from effocr_b200 import synth
from oracle import yolo as OY

ap = argparse.ArgumentParser()
ap.add_argument("out", nargs="?", default="gpurun_out/quickfit_yolov5s.npz")
ap.add_argument("--steps", type=int, default=2500)
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--lines", type=int, default=3000)
ap.add_argument("--lr", type=float, default=2e-3)
ap.add_argument("--native-frac", type=float, default=0.3)
ap.add_argument("--tracking", type=float, default=4.0, help="letter-spacing of the rendered lines (NMS runs at IoU 0.01)")
args = ap.parse_args()
dev = "cuda" if torch.cuda.is_available() else "cpu"
torch.manual_seed(0)
torch.backends.cudnn.allow_tf32 = True
torch.backends.cuda.matmul.allow_tf32 = True
rng = np.random.default_rng(0)
NC, NO = 2, 7

# ---- data: rendered lines; the letterboxed content strip (40 x 640) comes from the OpenCV-exact resize restatement
lines = synth.synthetic_lines(args.lines // 2, seed=4242, tracking=args.tracking) + synth.synthetic_lines(args.lines - args.lines // 2, seed=4243, tracking=args.tracking + 2)
strips = np.stack([OY.cv2_resize_linear_u8(np.ascontiguousarray(l[0][:, :, ::-1]), 640, 40)[:, :, ::-1] for l in lines])  # RGB
natives = np.stack([l[0] for l in lines])
strips_t = torch.from_numpy(np.ascontiguousarray(strips)).to(dev)  # [N, 40, 640, 3] u8
natives_t = torch.from_numpy(natives).to(dev)  # [N, 64, 1024, 3] u8
boxes = []  # per line: [n, 5] (cls, x0, y0, x1, y1) in ORIGINAL pixels
for _img, cb, wb, _chars in lines:
    b = [np.concatenate([np.zeros((len(cb), 1), np.float32), cb], 1), np.concatenate([np.ones((len(wb), 1), np.float32), wb], 1)]
    boxes.append(np.concatenate(b, 0).astype(np.float32))
print(f"{len(lines)} lines, {sum(len(b) for b in boxes)} boxes", flush=True)


def make_batch(ids, native: bool):
    """-> x f32 [B,3,H,W] (RGB / 255), targets [nt, 6] = (image, cls, xc, yc, w, h) normalised to the model input."""
    if native:
        x = natives_t[ids].permute(0, 3, 1, 2).float() / 255.0
        H, W, sx, sy, oy = 64, 1024, 1.0, 1.0, 0.0
    else:
        x = torch.full((len(ids), 3, 640, 640), 114.0 / 255.0, device=dev)
        x[:, :, 300:340, :] = strips_t[ids].permute(0, 3, 1, 2).float() / 255.0
        H, W, sx, sy, oy = 640, 640, 0.625, 0.625, 300.0
    t = []
    for bi, i in enumerate(ids.tolist()):
        b = boxes[i]
        x0, y0, x1, y1 = b[:, 1] * sx, b[:, 2] * sy + oy, b[:, 3] * sx, b[:, 4] * sy + oy
        t.append(np.stack([np.full(len(b), bi, np.float32), b[:, 0], (x0 + x1) / 2 / W, (y0 + y1) / 2 / H, (x1 - x0) / W, (y1 - y0) / H], 1))
    return x, torch.from_numpy(np.concatenate(t, 0)).to(dev)


# ---- model: oracle forward with BatchNorm in training mode (running statistics updated in the state dict)
sd = {k: v.clone().to(dev) for k, v in OY.init_yolov5s_state_dict(nc=NC, seed=0).items()}
for k in list(sd):
    if k.endswith("bn.weight"):
        sd[k] = torch.ones_like(sd[k])
    elif k.endswith(("bn.bias", "bn.running_mean")):
        sd[k] = torch.zeros_like(sd[k])
    elif k.endswith("bn.running_var"):
        sd[k] = torch.ones_like(sd[k])
trainable = {k: v.requires_grad_(True) for k, v in sd.items() if k.endswith(("conv.weight", "bn.weight", "bn.bias")) or k.startswith("model.24.m.")}
TRAINING = [True]


def conv_bn_act(sd_, prefix, x, k, s, dtype):
    w = sd_[prefix + "conv.weight"]
    y = F.conv2d(x, w, None, stride=s, padding=2 if k == 6 else k // 2)
    y = F.batch_norm(y, sd_[prefix + "bn.running_mean"], sd_[prefix + "bn.running_var"], sd_[prefix + "bn.weight"],
                     sd_[prefix + "bn.bias"], TRAINING[0], 0.03, OY.BN_EPS)
    return F.silu(y)


OY._conv = conv_bn_act  # the oracle's layer walker (C3 / SPPF / Detect) stays as is
ANCH = sd["model.24.anchors"]  # [3, 3, 2] in grid units


def bbox_ciou(p, t, eps=1e-7):
    """p, t: [n, 4] xywh (same units) -> CIoU [n]."""
    px1, px2, py1, py2 = p[:, 0] - p[:, 2] / 2, p[:, 0] + p[:, 2] / 2, p[:, 1] - p[:, 3] / 2, p[:, 1] + p[:, 3] / 2
    tx1, tx2, ty1, ty2 = t[:, 0] - t[:, 2] / 2, t[:, 0] + t[:, 2] / 2, t[:, 1] - t[:, 3] / 2, t[:, 1] + t[:, 3] / 2
    inter = (torch.min(px2, tx2) - torch.max(px1, tx1)).clamp(0) * (torch.min(py2, ty2) - torch.max(py1, ty1)).clamp(0)
    union = p[:, 2] * p[:, 3] + t[:, 2] * t[:, 3] - inter + eps
    iou = inter / union
    cw = torch.max(px2, tx2) - torch.min(px1, tx1)
    ch = torch.max(py2, ty2) - torch.min(py1, ty1)
    c2 = cw ** 2 + ch ** 2 + eps
    rho2 = (t[:, 0] - p[:, 0]) ** 2 + (t[:, 1] - p[:, 1]) ** 2
    v = (4 / math.pi ** 2) * (torch.atan(t[:, 2] / (t[:, 3] + eps)) - torch.atan(p[:, 2] / (p[:, 3] + eps))) ** 2
    with torch.no_grad():
        alpha = v / (v - iou + (1 + eps))
    return iou - (rho2 / c2 + v * alpha)


def compute_loss(raw, targets, anchor_t=4.0):
    lbox = lobj = lcls = torch.zeros((), device=dev)
    nt = targets.shape[0]
    ai = torch.arange(3, device=dev).float().view(3, 1).repeat(1, nt)
    tg = torch.cat((targets.repeat(3, 1, 1), ai[..., None]), 2)  # [3, nt, 7]
    off = torch.tensor([[0, 0], [1, 0], [0, 1], [-1, 0], [0, -1]], device=dev).float() * 0.5
    balance = (4.0, 1.0, 0.4)
    for i, pi in enumerate(raw):
        bs, _, ny, nx, _ = pi.shape
        gain = torch.tensor([1, 1, nx, ny, nx, ny, 1], device=dev).float()
        t = tg * gain
        anchors = ANCH[i]
        r = t[..., 4:6] / anchors[:, None]
        t = t[torch.max(r, 1 / r).max(2)[0] < anchor_t]
        gxy = t[:, 2:4]
        gxi = gain[[2, 3]] - gxy
        j, k = ((gxy % 1 < 0.5) & (gxy > 1)).T
        l, m = ((gxi % 1 < 0.5) & (gxi > 1)).T
        sel = torch.stack((torch.ones_like(j), j, k, l, m))
        t = t.repeat((5, 1, 1))[sel]
        offsets = (torch.zeros_like(gxy)[None] + off[:, None])[sel]
        b, c = t[:, 0].long(), t[:, 1].long()
        gxy, gwh, a = t[:, 2:4], t[:, 4:6], t[:, 6].long()
        gij = (gxy - offsets).long()
        gi, gj = gij[:, 0].clamp(0, nx - 1), gij[:, 1].clamp(0, ny - 1)
        tobj = torch.zeros(pi.shape[:4], device=dev)
        if len(b):
            ps = pi[b, a, gj, gi]
            pxy = ps[:, 0:2].sigmoid() * 2 - 0.5
            pwh = (ps[:, 2:4].sigmoid() * 2) ** 2 * anchors[a]
            iou = bbox_ciou(torch.cat((pxy, pwh), 1), torch.cat((gxy - gij, gwh), 1))
            lbox = lbox + (1.0 - iou).mean()
            tobj[b, a, gj, gi] = iou.detach().clamp(0)
            tc = torch.zeros_like(ps[:, 5:])
            tc[torch.arange(len(b)), c] = 1.0
            lcls = lcls + F.binary_cross_entropy_with_logits(ps[:, 5:], tc)
        lobj = lobj + F.binary_cross_entropy_with_logits(pi[..., 4], tobj) * balance[i]
    return 0.05 * lbox + 1.0 * lobj + 0.3 * lcls, (lbox.item(), lobj.item(), lcls.item())


params = list(trainable.values())
decay = [v for k, v in trainable.items() if v.dim() > 1]
no_decay = [v for k, v in trainable.items() if v.dim() <= 1]
opt = torch.optim.AdamW([{"params": decay, "weight_decay": 5e-4}, {"params": no_decay, "weight_decay": 0.0}], lr=args.lr)
n_train = args.lines - 64
t0 = time.time()
for step in range(args.steps):
    native = rng.random() < args.native_frac
    ids = torch.from_numpy(rng.integers(0, n_train, args.batch * (2 if native else 1)))
    x, tg = make_batch(ids, native)
    x = x + 0.02 * torch.randn_like(x)
    _out, raw = OY.yolov5s_forward(sd, x, return_raw=True)
    loss, parts = compute_loss(raw, tg)
    opt.zero_grad(set_to_none=True)
    (loss * args.batch).backward()
    torch.nn.utils.clip_grad_norm_(params, 10.0)
    for g in opt.param_groups:
        g["lr"] = args.lr * min(1.0, (step + 1) / 100) * (0.05 + 0.95 * 0.5 * (1 + math.cos(math.pi * step / args.steps)))
    opt.step()
    if step % 100 == 0 or step == args.steps - 1:
        print(f"step {step} loss {loss.item():.4f} box {parts[0]:.3f} obj {parts[1]:.4f} cls {parts[2]:.4f} ({time.time() - t0:.0f}s)", flush=True)

# ---- evaluate on held-out lines with the fp16-rounded weights that are saved (fp32 oracle forward, reference NMS)
TRAINING[0] = False
final = {k: v.detach().half().float() for k, v in sd.items()}
ids = torch.arange(n_train, args.lines)
for native in (False, True):
    with torch.no_grad():
        x, _ = make_batch(ids, native)
        pred = OY.yolov5s_forward(final, x).cpu()
    det = OY.non_max_suppression(pred, conf_thres=0.35, iou_thres=0.01, max_det=1000)
    sx, sy, oy = (1.0, 1.0, 0.0) if native else (0.625, 0.625, 300.0)
    n_gt = n_hit = n_det = 0
    for i, d in zip(ids.tolist(), det):
        g = boxes[i][boxes[i][:, 0] == 0][:, 1:] * np.array([sx, sy, sx, sy], np.float32) + np.array([0, oy, 0, oy], np.float32)
        p = d[d[:, 5] == 0][:, :4].numpy()
        n_gt += len(g)
        n_det += len(p)
        if len(p) and len(g):
            ix = (np.minimum(g[:, None, 2], p[None, :, 2]) - np.maximum(g[:, None, 0], p[None, :, 0])).clip(0)
            iy = (np.minimum(g[:, None, 3], p[None, :, 3]) - np.maximum(g[:, None, 1], p[None, :, 1])).clip(0)
            inter = ix * iy
            union = ((g[:, 2] - g[:, 0]) * (g[:, 3] - g[:, 1]))[:, None] + ((p[:, 2] - p[:, 0]) * (p[:, 3] - p[:, 1]))[None, :] - inter
            n_hit += int(((inter / union).max(1) > 0.5).sum())
    conf = torch.cat([d[:, 4] for d in det] + [torch.zeros(1)])
    print(f"native={native}: held-out char boxes gt {n_gt} detected {n_det} matched(IoU>0.5) {n_hit}; words {sum(int((d[:, 5] == 1).sum()) for d in det)}; "
          f"kept conf median {conf.median().item():.3f} p10 {conf.quantile(0.1).item():.3f}; kept boxes with conf in (0.35, 0.45): {int(((conf > 0.35) & (conf < 0.45)).sum())} of {len(conf) - 1}",
          flush=True)
np.savez_compressed(args.out, **{k: v.cpu().numpy().astype(np.float16) for k, v in final.items()})
print("saved", args.out, flush=True)
