"""Drop-in for /root/reference/models/encoders.py: `AutoEncoderFactory(backend, modelpath)`.

Same call surface (returned class: `__init__(model=modelpath, device='cuda')`, `.net`, `forward`,
`@classmethod load(checkpoint)`, state-dict keys `net.*`) so infer_effocr.py:487-496,177-179,314
run unchanged; the forward pass runs in effocr_b200's sm_100a kernels (csrc/vit.cu), not in
timm / transformers.  There is no eager fallback: forward() on a machine without a B200 raises.

Differences forced by the environment, all documented in INTEGRATION.md:
  * the reference's `timm.create_model(..., pretrained=True)` (encoders.py:58) downloads weights;
    there is no network here, so a fresh AutoEncoder() is timm-initialised (trunc_normal .02) and
    real weights arrive through `load(checkpoint)` / `load_state_dict` exactly as in the reference.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

from . import _lib
from .engine import CONVNEXT_DEPTHS, CONVNEXT_DIMS, VIT_CONFIGS, ConvNextEngine, VitEngine

_TIMM_ALIASES = {
    "vit_tiny_patch16_224.augreg_in21k_ft_in1k": "vit_tiny_patch16_224",
    "vit_small_patch16_224.augreg_in21k_ft_in1k": "vit_small_patch16_224",
    "vit_base_patch16_224.augreg2_in21k_ft_in1k": "vit_base_patch16_224",
}


class TimmViTParams(torch.nn.Module):
    """Parameter container with timm VisionTransformer's parameter names and shapes."""

    def __init__(self, name: str):
        super().__init__()
        name = _TIMM_ALIASES.get(name, name)
        if name not in VIT_CONFIGS:
            raise NotImplementedError(f"encoder '{name}' is not implemented in effocr_b200 (have: {sorted(VIT_CONFIGS)})")
        d, _heads, depth, mlp = VIT_CONFIGS[name]
        self.arch = name
        P = torch.nn.Parameter

        def tn(*shape):
            t = torch.empty(*shape)
            torch.nn.init.trunc_normal_(t, std=0.02)
            return P(t)

        self.cls_token = P(torch.randn(1, 1, d) * 1e-6)
        self.pos_embed = tn(1, 197, d)
        self.patch_embed = torch.nn.Module()
        self.patch_embed.proj = torch.nn.Conv2d(3, d, 16, 16)
        self.blocks = torch.nn.ModuleList()
        for _ in range(depth):
            blk = torch.nn.Module()
            blk.norm1 = torch.nn.LayerNorm(d, eps=1e-6)
            blk.attn = torch.nn.Module()
            blk.attn.qkv = torch.nn.Linear(d, 3 * d)
            blk.attn.proj = torch.nn.Linear(d, d)
            blk.norm2 = torch.nn.LayerNorm(d, eps=1e-6)
            blk.mlp = torch.nn.Module()
            blk.mlp.fc1 = torch.nn.Linear(d, mlp)
            blk.mlp.fc2 = torch.nn.Linear(mlp, d)
            for lin in (blk.attn.qkv, blk.attn.proj, blk.mlp.fc1, blk.mlp.fc2):
                torch.nn.init.trunc_normal_(lin.weight, std=0.02)
                torch.nn.init.zeros_(lin.bias)
            self.blocks.append(blk)
        self.norm = torch.nn.LayerNorm(d, eps=1e-6)
        self.num_features = self.embed_dim = d


class TimmConvNeXtParams(torch.nn.Module):
    """Parameter container with timm convnext_tiny's parameter names and shapes (timm init: trunc_normal .02,
    zero bias, layer-scale gamma 1e-6)."""

    def __init__(self, name: str = "convnext_tiny"):
        super().__init__()
        if name.split(".")[0] != "convnext_tiny":
            raise NotImplementedError(f"encoder '{name}' is not implemented in effocr_b200")
        self.arch = "convnext_tiny"
        nn = torch.nn

        def init(m):
            torch.nn.init.trunc_normal_(m.weight, std=0.02)
            torch.nn.init.zeros_(m.bias)
            return m

        self.stem = nn.Sequential(init(nn.Conv2d(3, 96, 4, 4)), nn.LayerNorm(96, eps=1e-6))
        self.stages = nn.ModuleList()
        for i, (depth, dim) in enumerate(zip(CONVNEXT_DEPTHS, CONVNEXT_DIMS)):
            st = nn.Module()
            st.downsample = nn.Identity() if i == 0 else nn.Sequential(nn.LayerNorm(CONVNEXT_DIMS[i - 1], eps=1e-6),
                                                                         init(nn.Conv2d(CONVNEXT_DIMS[i - 1], dim, 2, 2)))
            st.blocks = nn.ModuleList()
            for _ in range(depth):
                b = nn.Module()
                b.conv_dw = init(nn.Conv2d(dim, dim, 7, padding=3, groups=dim))
                b.norm = nn.LayerNorm(dim, eps=1e-6)
                b.mlp = nn.Module()
                b.mlp.fc1 = init(nn.Linear(dim, 4 * dim))
                b.mlp.fc2 = init(nn.Linear(4 * dim, dim))
                b.gamma = nn.Parameter(torch.full((dim,), 1e-6))
                st.blocks.append(b)
            self.stages.append(st)
        self.head = nn.Module()
        self.head.norm = nn.LayerNorm(768, eps=1e-6)
        self.num_features = 768


def hf_vit_to_timm(hf_sd, prefix: str = "") -> "OrderedDict[str, torch.Tensor]":
    """transformers.ViTModel state dict -> timm names (the mapping the reference documents in
    scripts/trocr_fairseq_to_pytorch_chkpt.py:30-88, inverted); fused qkv rows are [q | k | v]."""
    w = {k[len(prefix):]: v for k, v in hf_sd.items() if k.startswith(prefix)}
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
    return out


class _EngineBackedEncoder(torch.nn.Module):
    """Common machinery: an nn.Module whose forward() runs the C-ABI engine built from its own
    current parameters (rebuilt when the parameters change)."""

    max_batch = 1024
    ln_eps = 1e-6  # timm ViT LayerNorm epsilon

    def _timm_state(self):  # -> dict of timm-keyed tensors (no prefix)
        raise NotImplementedError

    def _params_version(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def engine(self):
        ver = self._params_version()
        if getattr(self, "_engine", None) is None or self._engine_ver != ver:
            sd = {"net." + k: v for k, v in self._timm_state().items()}
            if "net.stem.0.weight" in sd:
                eng = ConvNextEngine(sd, prefix="net.", max_batch=min(self.max_batch, 4096))
            else:
                eng = VitEngine(sd, prefix="net.", max_batch=self.max_batch, ln_eps=self.ln_eps)
            object.__setattr__(self, "_engine", eng)
            object.__setattr__(self, "_engine_ver", ver)
        return self._engine

    def forward(self, x):
        if not isinstance(x, torch.Tensor):
            x = torch.as_tensor(x)
        if not torch.cuda.is_available():
            raise _lib.EffocrError("effocr_b200 encoders run only on a B200 (no CPU / eager fallback)")
        dev_in = x.device
        y = self.engine().forward(x.to("cuda", torch.float32))
        return y if dev_in.type == "cuda" else y.to(dev_in)


def AutoEncoderFactory(backend, modelpath):
    """models/encoders.py:50-97."""

    if backend == "timm":

        class AutoEncoder(_EngineBackedEncoder):

            def __init__(self, model=modelpath, device="cuda"):
                super().__init__()
                net = TimmConvNeXtParams(model) if str(model).startswith("convnext") else TimmViTParams(model)
                if device != "cpu" and torch.cuda.is_available():
                    net.to(device)
                self.net = net

            def _timm_state(self):
                return self.net.state_dict()

            @classmethod
            def load(cls, checkpoint):
                ptnet = cls()
                ptnet.load_state_dict(torch.load(checkpoint, map_location="cpu"))
                return ptnet

    elif backend == "hf":

        class AutoEncoder(_EngineBackedEncoder):

            def __init__(self, model=modelpath, device="cuda"):
                super().__init__()
                from transformers import AutoModel

                net = AutoModel.from_pretrained(model)
                if device != "cpu" and torch.cuda.is_available():
                    net.to(device)
                self.net = net
                self.ln_eps = float(getattr(net.config, "layer_norm_eps", 1e-12))

            def _timm_state(self):
                return hf_vit_to_timm(self.net.state_dict())

            @classmethod
            def load(cls, checkpoint):
                ptnet = cls()
                ptnet.load_state_dict(torch.load(checkpoint, map_location="cpu"))
                return ptnet

    else:
        raise NotImplementedError

    return AutoEncoder
