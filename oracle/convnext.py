"""TEST INFRASTRUCTURE (oracle) -- CPU restatement of the ConvNeXt recognizer encoder.

`timm.create_model("convnext_tiny", num_classes=0)` (reached from /root/reference/models/encoders.py:58; BASELINE
config 4) is restated on a timm-keyed state dict (`net.` prefix): stem Conv2d(3,96,k4,s4) + LayerNorm2d; stages of
depths (3,3,9,3) / dims (96,192,384,768), stages 2-4 preceded by LayerNorm2d + Conv2d(k2,s2); block =
dwconv7x7 -> LayerNorm -> Linear(C,4C) -> GELU(erf) -> Linear(4C,C) -> * gamma -> + x; head = global average pool ->
LayerNorm2d -> flatten -> [B, 768].  LayerNorm eps 1e-6.

Pinned (tests/test_oracle.py) against torchvision.models.convnext_tiny with `classifier[2] = Identity` through the
key map below (same math; timm itself is not installable here).
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn.functional as F

DEPTHS = (3, 3, 9, 3)
DIMS = (96, 192, 384, 768)
LN_EPS = 1e-6


def init_convnext_tiny_state_dict(seed: int = 0, prefix: str = "net.", gamma_scale: float = 0.5):
    """trunc_normal(.02) conv / linear weights, small random biases, LN (1 + noise, noise); gamma ~ U(0, gamma_scale)
    (timm initialises gamma to 1e-6, which would hide every block: a trained checkpoint has O(0.1..1) values)."""
    g = torch.Generator().manual_seed(seed)

    def tn(*shape):
        t = torch.empty(*shape)
        torch.nn.init.trunc_normal_(t, std=0.02, a=-0.04, b=0.04, generator=g)
        return t

    def ln(c, key, sd):
        sd[key + ".weight"] = 1.0 + 0.05 * torch.randn(c, generator=g)
        sd[key + ".bias"] = 0.02 * torch.randn(c, generator=g)

    sd = OrderedDict()
    sd["stem.0.weight"] = tn(96, 3, 4, 4) * 5
    sd["stem.0.bias"] = 0.02 * torch.randn(96, generator=g)
    ln(96, "stem.1", sd)
    for i, (depth, dim) in enumerate(zip(DEPTHS, DIMS)):
        if i > 0:
            ln(DIMS[i - 1], f"stages.{i}.downsample.0", sd)
            sd[f"stages.{i}.downsample.1.weight"] = tn(dim, DIMS[i - 1], 2, 2) * 3
            sd[f"stages.{i}.downsample.1.bias"] = 0.02 * torch.randn(dim, generator=g)
        for j in range(depth):
            p = f"stages.{i}.blocks.{j}."
            sd[p + "conv_dw.weight"] = tn(dim, 1, 7, 7) * 5
            sd[p + "conv_dw.bias"] = 0.02 * torch.randn(dim, generator=g)
            ln(dim, p + "norm", sd)
            sd[p + "mlp.fc1.weight"] = tn(4 * dim, dim) * 3
            sd[p + "mlp.fc1.bias"] = 0.02 * torch.randn(4 * dim, generator=g)
            sd[p + "mlp.fc2.weight"] = tn(dim, 4 * dim) * 3
            sd[p + "mlp.fc2.bias"] = 0.02 * torch.randn(dim, generator=g)
            sd[p + "gamma"] = gamma_scale * torch.rand(dim, generator=g)
    ln(768, "head.norm", sd)
    return OrderedDict((prefix + k, v) for k, v in sd.items())


def _ln2d(x, w, b):
    """LayerNorm over channels of an NCHW tensor (timm LayerNorm2d)."""
    return F.layer_norm(x.permute(0, 2, 3, 1), (x.shape[1],), w, b, LN_EPS).permute(0, 3, 1, 2)


def convnext_forward(sd, x: torch.Tensor, prefix: str = "net.", dtype=torch.float32):
    """x f32 [B,3,224,224] -> [B, 768] pooled pre-logits (timm num_classes=0)."""
    w = {k[len(prefix):]: v.to(dtype) for k, v in sd.items() if k.startswith(prefix)}
    x = x.to(dtype)
    x = F.conv2d(x, w["stem.0.weight"], w["stem.0.bias"], stride=4)
    x = _ln2d(x, w["stem.1.weight"], w["stem.1.bias"])
    for i, (depth, dim) in enumerate(zip(DEPTHS, DIMS)):
        if i > 0:
            x = _ln2d(x, w[f"stages.{i}.downsample.0.weight"], w[f"stages.{i}.downsample.0.bias"])
            x = F.conv2d(x, w[f"stages.{i}.downsample.1.weight"], w[f"stages.{i}.downsample.1.bias"], stride=2)
        for j in range(depth):
            p = f"stages.{i}.blocks.{j}."
            h = F.conv2d(x, w[p + "conv_dw.weight"], w[p + "conv_dw.bias"], padding=3, groups=dim)
            h = h.permute(0, 2, 3, 1)
            h = F.layer_norm(h, (dim,), w[p + "norm.weight"], w[p + "norm.bias"], LN_EPS)
            h = F.linear(h, w[p + "mlp.fc1.weight"], w[p + "mlp.fc1.bias"])
            h = F.gelu(h)
            h = F.linear(h, w[p + "mlp.fc2.weight"], w[p + "mlp.fc2.bias"])
            h = (h * w[p + "gamma"]).permute(0, 3, 1, 2)
            x = x + h
    x = x.mean(dim=(2, 3), keepdim=True)
    x = _ln2d(x, w["head.norm.weight"], w["head.norm.bias"])
    return x.flatten(1)


def timm_to_torchvision(sd, prefix: str = "net."):
    w = {k[len(prefix):]: v for k, v in sd.items()}
    out = OrderedDict()
    out["features.0.0.weight"], out["features.0.0.bias"] = w["stem.0.weight"], w["stem.0.bias"]
    out["features.0.1.weight"], out["features.0.1.bias"] = w["stem.1.weight"], w["stem.1.bias"]
    for i, depth in enumerate(DEPTHS):
        if i > 0:
            for n, m in ((0, 0), (1, 1)):
                out[f"features.{2 * i}.{m}.weight"] = w[f"stages.{i}.downsample.{n}.weight"]
                out[f"features.{2 * i}.{m}.bias"] = w[f"stages.{i}.downsample.{n}.bias"]
        for j in range(depth):
            s, t = f"stages.{i}.blocks.{j}.", f"features.{2 * i + 1}.{j}."
            for a, b in (("conv_dw", "block.0"), ("norm", "block.2"), ("mlp.fc1", "block.3"), ("mlp.fc2", "block.5")):
                out[t + b + ".weight"] = w[s + a + ".weight"]
                out[t + b + ".bias"] = w[s + a + ".bias"]
            out[t + "layer_scale"] = w[s + "gamma"].reshape(-1, 1, 1)
    out["classifier.0.weight"], out["classifier.0.bias"] = w["head.norm.weight"], w["head.norm.bias"]
    return out
