"""TEST INFRASTRUCTURE (oracle) -- CPU restatement of the recognizer encoder forward pass.

The reference builds its encoder with `timm.create_model(name, num_classes=0)`
(/root/reference/models/encoders.py:58) and calls it at infer_effocr.py:314 /
onnx_engines/recognizer_engine.py:27.  timm is an un-vendored, un-pinned dependency
(requirements.txt:27), so its published ViT algorithm (timm.models.vision_transformer
.VisionTransformer, `vit_{tiny,small,base}_patch16_224`) is restated here in
plain torch ops on timm-keyed state dicts (`net.` prefix as saved by the reference,
train_effocr_recognizer.py:65-72).

Pinned (tests/test_oracle.py: test_vit_matches_golden_reference_hf_backend, test_vit_matches_torchvision) against
  * the reference's own "hf" back-end, run live: AutoEncoderFactory("hf", dir) (encoders.py:72-91)
    over a random-init transformers.ViTModel, weights mapped key-by-key, and
  * torchvision.models.vision_transformer.VisionTransformer with the same weights,
and by the committed fixture tests/golden/vit_tiny_golden.npz (oracle/make_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

VIT_CONFIGS = {
    # name: (embed dim, heads, depth, mlp dim)
    "vit_tiny_patch16_224": (192, 3, 12, 768),
    "vit_small_patch16_224": (384, 6, 12, 1536),
    "vit_base_patch16_224": (768, 12, 12, 3072),
}
PATCH = 16
IMG = 224
TOKENS = (IMG // PATCH) ** 2 + 1
LN_EPS = 1e-6


def init_vit_state_dict(name: str, seed: int = 0, prefix: str = "net.", scale: float = 1.0) -> "OrderedDict[str, torch.Tensor]":
    """timm-style random init (trunc_normal std .02 weights, zero bias, LN (1, 0), cls ~ N(0, 1e-6)).

    `scale` > 1 widens the weight distribution ("trained-like" magnitudes) so that embeddings of
    different glyphs separate by more than fp16 noise; scale = 1 is timm's init.
    """
    d, _h, depth, mlp = VIT_CONFIGS[name]
    g = torch.Generator().manual_seed(seed)

    def tn(*shape, std=0.02):
        t = torch.empty(*shape)
        torch.nn.init.trunc_normal_(t, std=std * scale, a=-2 * std * scale, b=2 * std * scale, generator=g)
        return t

    sd = OrderedDict()
    sd["cls_token"] = torch.randn(1, 1, d, generator=g) * 1e-6
    sd["pos_embed"] = tn(1, TOKENS, d)
    # timm leaves the patch conv at nn.Conv2d's default (kaiming-uniform, fan_in 768)
    bound = 1.0 / math.sqrt(3 * PATCH * PATCH)
    sd["patch_embed.proj.weight"] = (torch.rand(d, 3, PATCH, PATCH, generator=g) * 2 - 1) * bound
    sd["patch_embed.proj.bias"] = (torch.rand(d, generator=g) * 2 - 1) * bound
    for i in range(depth):
        p = f"blocks.{i}."
        sd[p + "norm1.weight"] = torch.ones(d)
        sd[p + "norm1.bias"] = torch.zeros(d)
        sd[p + "attn.qkv.weight"] = tn(3 * d, d)
        sd[p + "attn.qkv.bias"] = torch.zeros(3 * d)
        sd[p + "attn.proj.weight"] = tn(d, d)
        sd[p + "attn.proj.bias"] = torch.zeros(d)
        sd[p + "norm2.weight"] = torch.ones(d)
        sd[p + "norm2.bias"] = torch.zeros(d)
        sd[p + "mlp.fc1.weight"] = tn(mlp, d)
        sd[p + "mlp.fc1.bias"] = torch.zeros(mlp)
        sd[p + "mlp.fc2.weight"] = tn(d, mlp)
        sd[p + "mlp.fc2.bias"] = torch.zeros(d)
    sd["norm.weight"] = torch.ones(d)
    sd["norm.bias"] = torch.zeros(d)
    return OrderedDict((prefix + k, v) for k, v in sd.items())


def randomize_affine(sd, seed: int = 1, prefix: str = "net."):
    """Give biases / LayerNorm affine parameters non-trivial values (a trained checkpoint has them);
    keeps tests from passing with a kernel that ignores a bias."""
    g = torch.Generator().manual_seed(seed)
    out = OrderedDict()
    for k, v in sd.items():
        if k.endswith("bias") and "patch_embed" not in k:
            out[k] = torch.randn(v.shape, generator=g) * 0.02
        elif "norm" in k and k.endswith("weight"):
            out[k] = 1.0 + torch.randn(v.shape, generator=g) * 0.05
        else:
            out[k] = v
    return out


def infer_config(sd, prefix: str = "net."):
    d = sd[prefix + "cls_token"].shape[-1]
    depth = 1 + max(int(k[len(prefix):].split(".")[1]) for k in sd if k.startswith(prefix + "blocks."))
    mlp = sd[prefix + "blocks.0.mlp.fc1.weight"].shape[0]
    heads = d // 64
    return d, heads, depth, mlp


def vit_forward(sd, x: torch.Tensor, prefix: str = "net.", dtype=torch.float32, return_tokens: bool = False):
    """x f32 [B, 3, 224, 224] -> pooled pre-logits [B, D] (final-LN'd CLS token), timm num_classes=0.

    Per block (timm Block.forward): x += proj(softmax(q k^T / sqrt(hd)) v), q,k,v = split(qkv(LN1 x))
    with fused rows ordered [q | k | v], each [heads, hd];  x += fc2(GELU_erf(fc1(LN2 x))).
    """
    d, heads, depth, _mlp = infer_config(sd, prefix)
    hd = d // heads
    w = {k[len(prefix):]: v.to(dtype) for k, v in sd.items() if k.startswith(prefix)}
    x = x.to(dtype)
    b = x.shape[0]
    t = F.conv2d(x, w["patch_embed.proj.weight"], w["patch_embed.proj.bias"], stride=PATCH)
    t = t.flatten(2).transpose(1, 2)  # [B, 196, D], patches row-major
    t = torch.cat([w["cls_token"].expand(b, -1, -1), t], dim=1) + w["pos_embed"]
    n = t.shape[1]
    for i in range(depth):
        p = f"blocks.{i}."
        h = F.layer_norm(t, (d,), w[p + "norm1.weight"], w[p + "norm1.bias"], LN_EPS)
        qkv = F.linear(h, w[p + "attn.qkv.weight"], w[p + "attn.qkv.bias"])
        qkv = qkv.reshape(b, n, 3, heads, hd).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        att = (q @ k.transpose(-2, -1)) * (hd ** -0.5)
        att = att.softmax(dim=-1)
        o = (att @ v).transpose(1, 2).reshape(b, n, d)
        t = t + F.linear(o, w[p + "attn.proj.weight"], w[p + "attn.proj.bias"])
        h = F.layer_norm(t, (d,), w[p + "norm2.weight"], w[p + "norm2.bias"], LN_EPS)
        h = F.gelu(F.linear(h, w[p + "mlp.fc1.weight"], w[p + "mlp.fc1.bias"]))
        t = t + F.linear(h, w[p + "mlp.fc2.weight"], w[p + "mlp.fc2.bias"])
    t = F.layer_norm(t, (d,), w["norm.weight"], w["norm.bias"], LN_EPS)
    return t if return_tokens else t[:, 0]


def l2_normalize(e: torch.Tensor) -> torch.Tensor:
    """torch.nn.functional.normalize(p=2, dim=1), eps 1e-12 (infer_effocr.py:316,
    infer_effocr_onnx_multi.py:371)."""
    return e / e.norm(dim=1, keepdim=True).clamp_min(1e-12)


# ---------------------------------------------------------------------------- key maps
def timm_to_hf(sd, prefix: str = "net."):
    """timm-keyed -> transformers.ViTModel keys (inverse of the mapping documented in
    /root/reference/scripts/trocr_fairseq_to_pytorch_chkpt.py:30-88)."""
    d, _heads, depth, _ = infer_config(sd, prefix)
    w = {k[len(prefix):]: v for k, v in sd.items()}
    out = OrderedDict()
    out["embeddings.cls_token"] = w["cls_token"]
    out["embeddings.position_embeddings"] = w["pos_embed"]
    out["embeddings.patch_embeddings.projection.weight"] = w["patch_embed.proj.weight"]
    out["embeddings.patch_embeddings.projection.bias"] = w["patch_embed.proj.bias"]
    for i in range(depth):
        s, t = f"blocks.{i}.", f"encoder.layer.{i}."
        qw, kw, vw = w[s + "attn.qkv.weight"].split(d, 0)
        qb, kb, vb = w[s + "attn.qkv.bias"].split(d, 0)
        for nm, ww, bb in (("query", qw, qb), ("key", kw, kb), ("value", vw, vb)):
            out[t + f"attention.attention.{nm}.weight"] = ww
            out[t + f"attention.attention.{nm}.bias"] = bb
        out[t + "attention.output.dense.weight"] = w[s + "attn.proj.weight"]
        out[t + "attention.output.dense.bias"] = w[s + "attn.proj.bias"]
        out[t + "layernorm_before.weight"] = w[s + "norm1.weight"]
        out[t + "layernorm_before.bias"] = w[s + "norm1.bias"]
        out[t + "layernorm_after.weight"] = w[s + "norm2.weight"]
        out[t + "layernorm_after.bias"] = w[s + "norm2.bias"]
        out[t + "intermediate.dense.weight"] = w[s + "mlp.fc1.weight"]
        out[t + "intermediate.dense.bias"] = w[s + "mlp.fc1.bias"]
        out[t + "output.dense.weight"] = w[s + "mlp.fc2.weight"]
        out[t + "output.dense.bias"] = w[s + "mlp.fc2.bias"]
    out["layernorm.weight"] = w["norm.weight"]
    out["layernorm.bias"] = w["norm.bias"]
    return out


def hf_to_timm(hf_sd, prefix: str = "net.", hf_prefix: str = ""):
    """transformers.ViTModel keys -> timm keys (used by the product's "hf" encoder back-end too)."""
    w = {k[len(hf_prefix):]: v for k, v in hf_sd.items() if k.startswith(hf_prefix)}
    depth = 1 + max(int(k.split(".")[2]) for k in w if k.startswith("encoder.layer."))
    out = OrderedDict()
    out["cls_token"] = w["embeddings.cls_token"]
    out["pos_embed"] = w["embeddings.position_embeddings"]
    out["patch_embed.proj.weight"] = w["embeddings.patch_embeddings.projection.weight"]
    out["patch_embed.proj.bias"] = w["embeddings.patch_embeddings.projection.bias"]
    for i in range(depth):
        t, s = f"blocks.{i}.", f"encoder.layer.{i}."
        out[t + "norm1.weight"] = w[s + "layernorm_before.weight"]
        out[t + "norm1.bias"] = w[s + "layernorm_before.bias"]
        out[t + "attn.qkv.weight"] = torch.cat([w[s + f"attention.attention.{n}.weight"] for n in ("query", "key", "value")], 0)
        out[t + "attn.qkv.bias"] = torch.cat([w[s + f"attention.attention.{n}.bias"] for n in ("query", "key", "value")], 0)
        out[t + "attn.proj.weight"] = w[s + "attention.output.dense.weight"]
        out[t + "attn.proj.bias"] = w[s + "attention.output.dense.bias"]
        out[t + "norm2.weight"] = w[s + "layernorm_after.weight"]
        out[t + "norm2.bias"] = w[s + "layernorm_after.bias"]
        out[t + "mlp.fc1.weight"] = w[s + "intermediate.dense.weight"]
        out[t + "mlp.fc1.bias"] = w[s + "intermediate.dense.bias"]
        out[t + "mlp.fc2.weight"] = w[s + "output.dense.weight"]
        out[t + "mlp.fc2.bias"] = w[s + "output.dense.bias"]
    out["norm.weight"] = w["layernorm.weight"]
    out["norm.bias"] = w["layernorm.bias"]
    return OrderedDict((prefix + k, v) for k, v in out.items())


def timm_to_torchvision(sd, prefix: str = "net."):
    d, _heads, depth, _ = infer_config(sd, prefix)
    w = {k[len(prefix):]: v for k, v in sd.items()}
    out = OrderedDict()
    out["class_token"] = w["cls_token"]
    out["encoder.pos_embedding"] = w["pos_embed"]
    out["conv_proj.weight"] = w["patch_embed.proj.weight"]
    out["conv_proj.bias"] = w["patch_embed.proj.bias"]
    for i in range(depth):
        s, t = f"blocks.{i}.", f"encoder.layers.encoder_layer_{i}."
        out[t + "ln_1.weight"] = w[s + "norm1.weight"]
        out[t + "ln_1.bias"] = w[s + "norm1.bias"]
        out[t + "self_attention.in_proj_weight"] = w[s + "attn.qkv.weight"]
        out[t + "self_attention.in_proj_bias"] = w[s + "attn.qkv.bias"]
        out[t + "self_attention.out_proj.weight"] = w[s + "attn.proj.weight"]
        out[t + "self_attention.out_proj.bias"] = w[s + "attn.proj.bias"]
        out[t + "ln_2.weight"] = w[s + "norm2.weight"]
        out[t + "ln_2.bias"] = w[s + "norm2.bias"]
        out[t + "mlp.0.weight"] = w[s + "mlp.fc1.weight"]
        out[t + "mlp.0.bias"] = w[s + "mlp.fc1.bias"]
        out[t + "mlp.3.weight"] = w[s + "mlp.fc2.weight"]
        out[t + "mlp.3.bias"] = w[s + "mlp.fc2.bias"]
    out["encoder.ln.weight"] = w["norm.weight"]
    out["encoder.ln.bias"] = w["norm.bias"]
    return out
