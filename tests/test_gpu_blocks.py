"""GPU parity of the building-block kernels (GEMM epilogues, LayerNorm, attention) through the C ABI,
against plain fp32 torch on the same inputs.  Tolerances: fp16 operand rounding (rel 2^-11)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30)).item()


def _needs_ab():
    """A/B kernel variants (direct-store GEMM epilogues, earlier attention kernels) exist only in a library built with
    EFFOCR_AB=1; the product build refuses them loudly."""
    from effocr_b200 import _lib
    if not (_lib.load().effocr_build_flags() & 1):
        pytest.skip("A/B kernel variant: compiled out of the product build (EFFOCR_AB=1 python -m effocr_b200.build --force)")


@pytest.mark.parametrize("M,N,K,act,f32,resid,gamma,bn", [
    (128, 64, 64, 0, 0, False, False, 64),
    (300, 384, 384, 0, 0, False, False, 0),
    (1000, 1152, 384, 0, 0, False, False, 0),
    (1000, 1536, 384, 1, 0, False, False, 0),
    (1000, 384, 1536, 0, 1, True, False, 0),
    (777, 384, 384, 0, 1, True, True, 0),
    (3000, 100, 72, 2, 0, False, False, 0),
    (4000, 256, 288, 2, 0, True, False, 0),
    (3000, 21, 512, 0, 1, False, False, 64),
    (20000, 1152, 384, 0, 0, False, False, 0),   # many tiles per CTA: phase wrap of both pipelines
    (50000, 1152, 384, 0, 0, False, False, 0),   # A-stationary schedule (K <= 384, >= 2 row blocks per SM)
    (45000, 1536, 384, 1, 0, False, False, 0),
    (40000, 600, 320, 2, 0, False, False, 0),    # ragged N and K tail on the CTA-pair / A-stationary schedules
    (30011, 1152, 384, 0, 0, False, False, 0x20000),   # single-CTA TMA epilogue (A-stationary off, pair off)
    (30011, 384, 1536, 0, 1, False, False, 0x20000),
])
@pytest.mark.parametrize("direct", [False, True])
def test_gemm_epilogues(M, N, K, act, f32, resid, gamma, bn, direct):
    """Default dispatch = CTA-pair (cta_group::2) kernel for wide N and many tiles, single-CTA TMA-epilogue kernel
    otherwise; direct=True forces the first-generation direct-store epilogue."""
    from effocr_b200 import ops
    if direct or resid:  # a separate residual buffer / the forced direct-store epilogue: A/B-only paths
        _needs_ab()
    torch.manual_seed(0)
    a = (torch.randn(M, K, device="cuda") * 0.5).half()
    w = (torch.randn(N, K, device="cuda") * 0.05).half()
    b = torch.randn(N, device="cuda")
    g = torch.randn(N, device="cuda") if gamma else None
    odt = torch.float32 if f32 else torch.float16
    ld = (N + 7) // 8 * 8
    r = torch.randn(M, ld, device="cuda").to(odt)[:, :N] if resid else None
    out = ops.gemm(a, w, bias=b, act=act, out_dtype=odt, resid=r, gamma=g, block_n=bn, direct_epilogue=direct)
    ref = a.float() @ w.float().t() + b
    if act == 1:
        ref = torch.nn.functional.gelu(ref)
    if act == 2:
        ref = torch.nn.functional.silu(ref)
    if gamma:
        ref = ref * g
    if resid:
        ref = ref + r.float()
    tol = 2e-6 if f32 else 6e-4  # fp32 out: accumulation-order noise only; fp16 out: 2^-11 rounding
    assert _rel(out, ref) < tol
    assert torch.isfinite(out.float()).all()


@pytest.mark.parametrize("direct", [False, True])
@pytest.mark.parametrize("M,N,K,gamma", [(2000, 384, 384, False), (5000, 384, 1536, False), (1333, 96, 384, True),
                                          (4100, 768, 3072, True), (60000, 384, 384, False)])
def test_gemm_inplace_residual(M, N, K, gamma, direct):
    """x <- x + (a @ w.T + bias) * gamma: TMA reduce-add epilogue vs the direct read-modify-write one."""
    from effocr_b200 import ops
    if direct:
        _needs_ab()
    torch.manual_seed(1)
    a = (torch.randn(M, K, device="cuda") * 0.5).half()
    w = (torch.randn(N, K, device="cuda") * 0.05).half()
    b = torch.randn(N, device="cuda")
    g = torch.randn(N, device="cuda") if gamma else None
    x = torch.randn(M, N, device="cuda")
    upd = a.float() @ w.float().t() + b
    ref = x + (upd * g if gamma else upd)
    ops.gemm(a, w, bias=b, gamma=g, out_dtype=torch.float32, resid=x, out=x, direct_epilogue=direct)
    assert _rel(x, ref) < (2e-6 if K <= 1536 else 5e-6)  # fp32 accumulation-order noise grows with K


@pytest.mark.parametrize("M,N", [(100, 1152), (128, 1152), (1000, 1152), (40000, 1152), (201728 // 4, 1152), (3000, 768), (777, 200)])
def test_ln_gemm_equals_layernorm_then_gemm(M, N):
    """norm1 + QKV in one kernel: the A operand is normalised on the SM with the arithmetic of the stand-alone LayerNorm
    kernel, so the result equals layernorm -> gemm bit for bit (same fp16 operand, same MMA order per tile); row tails,
    one .. many row blocks per CTA (phase wrap of the A barriers), ragged N."""
    from effocr_b200 import ops
    torch.manual_seed(3)
    D = 384
    x = torch.randn(M, D, device="cuda") * 2 + 0.5
    g = torch.randn(D, device="cuda")
    b = torch.randn(D, device="cuda")
    w = (torch.randn(N, D, device="cuda") * 0.05).half()
    bias = torch.randn(N, device="cuda")
    out = ops.ln_gemm(x, g, b, w, bias)
    h = ops.layernorm(x, g, b, 1e-6, torch.float16)
    ref = ops.gemm(h, w, bias=bias)
    assert _rel(out, ref) < 1e-6
    ref32 = torch.nn.functional.layer_norm(x, (D,), g, b, 1e-6) @ w.float().t() + bias
    assert _rel(out, ref32) < 6e-4


@pytest.mark.parametrize("M", [100, 256, 1576, 40000, 201728 // 4 + 77])
def test_block_tail_equals_proj_ln_then_mlp(M):
    """Projection + residual + norm2 + MLP + residual in one kernel against the two kernels it replaces (proj_ln, mlp_fused)
    and against plain fp32 torch: same operands, the only arithmetic difference is that fc2 accumulates on top of the
    residual inside the fp32 accumulator.  Row tails, one .. several tiles per CTA pair (every barrier's phase wrap)."""
    from effocr_b200 import ops
    torch.manual_seed(5)
    D, HID = 384, 1536
    x0 = torch.randn(M, D, device="cuda") * 2 + 0.5
    att = (torch.randn(M, D, device="cuda") * 0.7).half()
    wp = (torch.randn(D, D, device="cuda") * 0.05).half()
    w1 = (torch.randn(HID, D, device="cuda") * 0.05).half()
    w2 = (torch.randn(D, HID, device="cuda") * 0.05).half()
    bp, g, be, b2 = (torch.randn(D, device="cuda") for _ in range(4))
    b1 = torch.randn(HID, device="cuda")
    x = x0.clone()
    ops.block_tail(x, att, wp, bp, g, be, w1, b1, w2, b2)
    ref = x0.clone()
    h = ops.proj_ln(ref, att, wp, bp, g, be)
    ops.mlp_fused(ref, h, w1, b1, w2, b2)
    # not bit-equal: the fp32 accumulator also carries the residual, and the row statistics are summed two lanes at a time
    # (packed fp32x2), which flips the fp16 rounding of an occasional h element
    assert _rel(x, ref) < 5e-5, _rel(x, ref)
    t = x0 + att.float() @ wp.float().t() + bp
    hh = torch.nn.functional.layer_norm(t, (D,), g, be, 1e-6).half().float()
    t = t + torch.nn.functional.gelu(hh @ w1.float().t() + b1).half().float() @ w2.float().t() + b2
    assert _rel(x, t) < 2e-4, _rel(x, t)


@pytest.mark.parametrize("dim", [96, 192, 384, 768])
@pytest.mark.parametrize("f32", [False, True])
def test_layernorm(dim, f32):
    from effocr_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(1003, dim, device="cuda") * 3 + 1
    g = torch.randn(dim, device="cuda")
    b = torch.randn(dim, device="cuda")
    out = ops.layernorm(x, g, b, 1e-6, torch.float32 if f32 else torch.float16)
    ref = torch.nn.functional.layer_norm(x, (dim,), g, b, 1e-6)
    assert _rel(out, ref) < (2e-6 if f32 else 4e-4)


@pytest.mark.parametrize("impl", [0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("batch,heads", [(1, 3), (5, 6), (64, 6), (200, 3)])
def test_attention(batch, heads, impl):
    from effocr_b200 import ops
    if impl != 0:
        _needs_ab()
    torch.manual_seed(0)
    T, D = 197, heads * 64
    qkv = (torch.randn(batch * T, 3 * D, device="cuda") * 1.5).half()
    out = ops.attention(qkv, batch, heads, impl=impl)
    q, k, v = qkv.float().reshape(batch, T, 3, heads, 64).permute(2, 0, 3, 1, 4)
    att = (q @ k.transpose(-2, -1) * 0.125).softmax(-1)
    ref = (att @ v).transpose(1, 2).reshape(batch * T, D)
    assert _rel(out, ref) < 1.5e-3  # P rounded to fp16 before P.V


def test_l2_normalize():
    from effocr_b200 import ops
    x = torch.randn(300, 384, device="cuda")
    x[7] = 0
    out = ops.l2_normalize(x)
    ref = torch.nn.functional.normalize(x, p=2, dim=1)
    assert torch.allclose(out, ref, atol=1e-6, rtol=1e-5)


@pytest.mark.parametrize("M,D,HID", [(100, 384, 1536), (256, 384, 1536), (300, 384, 256), (1182, 192, 768), (40000, 384, 1536),
                                     (70001, 384, 1536), (50000, 192, 768)])
def test_mlp_fused(M, D, HID):
    """x += GELU(h W1^T + b1) W2^T + b2 in ONE kernel (hidden activations stay on chip) vs fp32 torch with the hidden
    activations rounded to fp16 like the unfused path; row tails, one tile .. many tiles per CTA pair (phase wrap)."""
    from effocr_b200 import ops
    torch.manual_seed(0)
    h = (torch.randn(M, D, device="cuda") * 0.7).half()
    w1 = (torch.randn(HID, D, device="cuda") * 0.05).half()
    w2 = (torch.randn(D, HID, device="cuda") * 0.05).half()
    b1 = torch.randn(HID, device="cuda") * 0.3
    b2 = torch.randn(D, device="cuda") * 0.3
    x0 = torch.randn(M, D, device="cuda")
    x = x0.clone()
    ops.mlp_fused(x, h, w1, b1, w2, b2)
    p = torch.nn.functional.gelu(h.float() @ w1.float().t() + b1).half().float()
    upd = p @ w2.float().t() + b2
    assert _rel(x - x0, upd) < 1e-4  # fp16 rounding flips of the hidden activations (GELU approximation 8.6e-7 abs)
    # and against the unfused kernels of this library
    mid = ops.gemm(h, w1, bias=b1, act=1)
    y = x0.clone()
    ops.gemm(mid.contiguous(), w2, bias=b2, out_dtype=torch.float32, resid=y, out=y)
    assert _rel(x - x0, y - x0) < 1e-4
    assert torch.isfinite(x).all()


@pytest.mark.parametrize("M", [100, 256, 5000, 70001])
def test_proj_ln_fused(M):
    """x += att Wp^T + bp and h = LayerNorm(x) in ONE full-row kernel vs fp32 torch and vs the unfused kernels."""
    from effocr_b200 import ops
    D = 384
    torch.manual_seed(0)
    att = (torch.randn(M, D, device="cuda") * 0.7).half()
    w = (torch.randn(D, D, device="cuda") * 0.05).half()
    b = torch.randn(D, device="cuda") * 0.3
    g = 1.0 + 0.2 * torch.randn(D, device="cuda")
    be = 0.2 * torch.randn(D, device="cuda")
    x0 = torch.randn(M, D, device="cuda") * 2.0 + 0.5
    x = x0.clone()
    h = ops.proj_ln(x, att, w, b, g, be)
    xr = x0 + (att.float() @ w.float().t() + b)
    hr = torch.nn.functional.layer_norm(xr, (D,), g, be, 1e-6)
    assert _rel(x, xr) < 2e-6          # fp32 accumulation-order noise only
    assert _rel(h, hr) < 4e-4          # fp16 output rounding
    y = x0.clone()
    ops.gemm(att, w, bias=b, out_dtype=torch.float32, resid=y, out=y)
    hu = ops.layernorm(y, g, be)
    assert _rel(x, y) < 1e-6 and _rel(h, hu) < 3e-4
    assert torch.isfinite(h.float()).all()
