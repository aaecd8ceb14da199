"""GPU parity of the recognizer path (crop -> ViT -> L2 norm -> kNN) through the C ABI against the
CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star / SURVEY.md section 8c):
  * crop kernel: max-abs <= 1e-3 on the fp16 output, <= 2e-5 on the fp32 output;
  * embeddings: ||e_gpu - e_ref|| / ||e_ref|| <= 1e-3 (fp16 operands, fp32 accumulate);
  * kNN given identical embeddings: identical ids wherever the fp64 margin exceeds 1e-6
    (fp32 summation-order noise), distances within 2e-6.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _random_crops(rng, n, hmax=64, wlo=10, whi=48):
    return [rng.integers(0, 256, (hmax, int(rng.integers(wlo, whi + 1)), 3), dtype=np.uint8) for _ in range(n)]


@pytest.mark.parametrize("layout", ["nchw_f32", "nchw_f16", "patch_f16"])
def test_crop_resize_matches_oracle(layout):
    from effocr_b200 import ops
    from oracle import transform as T
    rng = np.random.default_rng(0)
    shapes = [(64, 23), (64, 64), (64, 100), (30, 64), (224, 50), (300, 40), (1, 1), (2, 7), (64, 10), (48, 225), (64, 400)]
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for (h, w) in shapes]
    pixels, images, _ = ops.pack_images(imgs)
    boxes = [(i, 0, 0, imgs[i].shape[1], imgs[i].shape[0]) for i in range(len(imgs))]
    # sub-rectangles, numpy-slice semantics (negative start wraps, stop clamps) and an empty slice
    boxes += [(2, 10, 0, 60, 64), (2, -30, 0, 1000, 64), (1, 5, 3, 5, 64), (10, 100, 0, 164, 64)]
    bt, n = ops.pack_boxes(boxes)
    code = {"nchw_f32": ops.CROP_NCHW_F32, "nchw_f16": ops.CROP_NCHW_F16, "patch_f16": ops.CROP_PATCH_F16}[layout]
    out = ops.crop_resize(pixels, images, bt, n, code).float().cpu().numpy()
    if layout == "patch_f16":  # [n*196, 768] -> [n,3,224,224]; the patch-major layout is white-centred (see crop.cu)
        out = out.reshape(n, 14, 14, 3, 16, 16).transpose(0, 3, 1, 4, 2, 5).reshape(n, 3, 224, 224)
        white = (1.0 - np.asarray(T.IMAGENET_MEAN, np.float32)) / np.asarray(T.IMAGENET_STD, np.float32)
        out += white[None, :, None, None]  # an empty rectangle is the reference's all-zero crop: stored as -white
    # half an fp16 ulp: 1e-3 at |x| ~ 2.6 (NCHW), 2e-3 at the centred layout's darkest ink (|x - white| up to 4.8; white is 0)
    tol = {"nchw_f32": 2e-5, "nchw_f16": 1.1e-3, "patch_f16": 2.1e-3}[layout]
    for j, (i, x0, y0, x1, y1) in enumerate(boxes):
        crop = imgs[i][y0:y1, x0:x1, :]
        if crop.size == 0:
            assert np.abs(out[j]).max() <= (0 if layout != "patch_f16" else 1e-3)  # fp16(-white) + white
            continue
        ref = T.paired_transform(crop)
        assert np.abs(out[j] - ref).max() <= tol, (j, boxes[j])


@pytest.mark.parametrize("name,batch", [("vit_tiny_patch16_224", 5), ("vit_small_patch16_224", 3)])
def test_vit_embeddings_match_oracle(name, batch):
    from effocr_b200.engine import VitEngine
    from oracle import vit as V
    sd = V.randomize_affine(V.init_vit_state_dict(name, seed=0))
    x = torch.randn(batch, 3, 224, 224, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        ref = V.vit_forward(sd, x)
    eng = VitEngine(sd, max_batch=2)  # forces chunking (2 + 2 + 1)
    out = eng.forward(x.cuda()).cpu()
    rel = ((out - ref).norm(dim=1) / ref.norm(dim=1)).max().item()
    assert rel <= 1e-3, rel


V_MEAN, V_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def test_vit_patch_input_equals_nchw_input():
    from effocr_b200.engine import VitEngine
    from oracle import vit as V
    sd = V.init_vit_state_dict("vit_tiny_patch16_224", seed=3)
    eng = VitEngine(sd, max_batch=4)
    x = torch.randn(4, 3, 224, 224, generator=torch.Generator().manual_seed(2)).half().float()
    e0 = eng.forward(x.cuda())
    white = ((1.0 - torch.tensor(V_MEAN)) / torch.tensor(V_STD)).view(1, 3, 1, 1)  # patch-major inputs are white-centred
    patches = (x - white).reshape(4, 3, 14, 16, 14, 16).permute(0, 2, 4, 1, 3, 5).reshape(4 * 196, 768).half().cuda()
    e1 = eng.forward(patches)
    assert torch.equal(e0, e1)  # same kernels, same operands: bit-identical, batch-composition independent
    e2 = eng.forward(x[1:2].cuda())
    assert torch.equal(e0[1:2], e2)


@pytest.mark.parametrize("n,d,nq,k", [(94, 192, 37, 1), (94, 192, 37, 10), (10000, 384, 1024, 10), (1000, 768, 5, 32),
                                      (5, 384, 3, 10), (300, 100, 200, 4)])
def test_knn_matches_oracle(n, d, nq, k):
    from effocr_b200.engine import FlatIPIndex
    from oracle import knn as K
    g = torch.Generator().manual_seed(n + d)
    xb = torch.nn.functional.normalize(torch.randn(n, d, generator=g) + 2.0, dim=1)  # clustered: cos ~ 0.8
    q = torch.nn.functional.normalize(torch.randn(nq, d, generator=g) + 2.0, dim=1)
    index = FlatIPIndex(d)
    index.add(xb)
    dist, idx = index.search_device(q.cuda(), k)
    dist, idx = dist.cpu(), idx.cpu()
    rd, ri = K.flat_ip_search(xb, q, k)
    s64, margin = K.margins(xb, q, k)
    kk = min(k, n)
    assert torch.all(idx[:, kk:] == -1)
    decidable = margin > 1e-6
    if nq >= 30:
        assert decidable.float().mean() > 0.8
    assert torch.equal(idx[decidable][:, :kk], ri[decidable][:, :kk])
    assert torch.allclose(dist[:, :kk], rd[:, :kk], atol=2e-6, rtol=0)
    # undecidable rows: returned ids must still be genuine top-k up to the noise floor
    got = torch.gather(s64, 1, idx[:, :kk])
    best = torch.sort(s64, dim=1, descending=True).values[:, :kk]
    assert (best - got).abs().max() < 2e-6


def test_knn_ties_lowest_id_and_duplicates():
    from effocr_b200.engine import FlatIPIndex
    d = 192
    base = torch.nn.functional.normalize(torch.randn(40, d, generator=torch.Generator().manual_seed(0)), dim=1)
    xb = torch.cat([base, base, base[:5]], 0)  # exact duplicates: ids i, i+40 (and i+80 for i < 5)
    index = FlatIPIndex(d)
    index.add(xb)
    _, idx = index.search_device(base.cuda(), 3)
    idx = idx.cpu()
    for i in range(40):
        exp = [i, i + 40, i + 80] if i < 5 else [i, i + 40]
        assert idx[i, :len(exp)].tolist() == exp


def test_knn_remove_ids_compacts():
    from effocr_b200.engine import FlatIPIndex
    d = 192
    xb = torch.nn.functional.normalize(torch.randn(50, d, generator=torch.Generator().manual_seed(1)), dim=1)
    index = FlatIPIndex(d)
    index.add(xb)
    index.remove_ids(np.array([3, 10]))
    assert index.ntotal == 48
    _, idx = index.search_device(xb[[2, 4, 11, 49]].cuda(), 1)
    assert idx.cpu().flatten().tolist() == [2, 3, 9, 47]


def test_convnext_embeddings_match_oracle():
    """ConvNeXt-Tiny (BASELINE config 4 encoder): fp16 operands vs the fp32 oracle, chunked batches."""
    from effocr_b200.engine import ConvNextEngine
    from oracle import convnext as OC
    sd = OC.init_convnext_tiny_state_dict(seed=0)
    x = torch.randn(3, 3, 224, 224, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        ref = OC.convnext_forward(sd, x)
    eng = ConvNextEngine(sd, max_batch=2)
    out = eng.forward(x.cuda()).cpu()
    rel = ((out - ref).norm(dim=1) / ref.norm(dim=1)).max().item()
    print(f"convnext_tiny: max relative embedding error {rel:.2e}")
    assert rel <= 1e-3, rel  # north_star tolerance (measured 4.1e-4 with the white-centred 4x4-patch rows)


def test_convnext_crop_patch4_path_equals_nchw_path():
    from effocr_b200 import ops
    from effocr_b200.engine import ConvNextEngine
    from oracle import convnext as OC
    sd = OC.init_convnext_tiny_state_dict(seed=1)
    eng = ConvNextEngine(sd, max_batch=4)
    rng = np.random.default_rng(0)
    crops = _random_crops(rng, 4)
    pixels, images, _ = ops.pack_images(crops)
    bt, n = ops.pack_boxes([(i, 0, 0, c.shape[1], c.shape[0]) for i, c in enumerate(crops)])
    nchw = ops.crop_resize(pixels, images, bt, n, ops.CROP_NCHW_F32)  # fp32: both paths round (value - white) once
    e0 = eng.forward(nchw)
    ops.crop_resize(pixels, images, bt, n, ops.CROP_PATCH4_F16, out=eng.patch_buffer(n))
    e1 = eng.forward(None, batch=n)
    assert torch.equal(e0, e1)


def test_autoencoder_factory_convnext_and_vit_surface():
    from effocr_b200.encoders import AutoEncoderFactory
    for name, dim in (("vit_tiny_patch16_224", 192), ("convnext_tiny", 768)):
        enc = AutoEncoderFactory("timm", name)(device="cuda").eval()
        assert any(k.startswith("net.") for k in enc.state_dict())
        with torch.no_grad():
            y = enc(torch.randn(2, 3, 224, 224, device="cuda"))
        assert y.shape == (2, dim) and y.is_cuda and torch.isfinite(y).all()


def test_knn_stress_config4_shape():
    """BASELINE config 4: 50k-glyph index, D = 768, 4096 queries, k = 10 -- score matrix (819 MB) never materialised."""
    from effocr_b200.engine import FlatIPIndex
    from oracle import knn as K
    g = torch.Generator().manual_seed(4)
    xb = torch.nn.functional.normalize(torch.randn(50000, 768, generator=g) + 1.0, dim=1)
    q = torch.nn.functional.normalize(torch.randn(4096, 768, generator=g) + 1.0, dim=1)
    index = FlatIPIndex(768)
    index.add(xb)
    dist, idx = index.search_device(q.cuda(), 10)
    sub = slice(0, 256)  # oracle on a slice (fp64 margins on 256 x 50k)
    rd, ri = K.flat_ip_search(xb, q[sub], 10)
    _, margin = K.margins(xb, q[sub], 10)
    dec = margin > 1e-6
    assert torch.equal(idx.cpu()[sub][dec], ri[dec])
    assert torch.allclose(dist.cpu()[sub], rd, atol=3e-6, rtol=0)


def test_recognize_stream_equals_recognize_packed():
    """The one-batch-in-flight streaming API returns, batch by batch, exactly what the synchronous call returns."""
    import numpy as np
    import torch
    from effocr_b200 import synth
    from effocr_b200.pipeline import PackedCrops, RecognizerPipeline
    from oracle import vit as OV
    sd = OV.randomize_affine(OV.init_vit_state_dict("vit_tiny_patch16_224", seed=0))
    xb = torch.nn.functional.normalize(torch.randn(300, 192, generator=torch.Generator().manual_seed(3)), dim=1)
    pipe = RecognizerPipeline(sd, xb, max_batch=64)
    batches = [PackedCrops(synth.synthetic_crops(n, seed=s)[0]) for n, s in ((40, 1), (7, 2), (64, 3), (90, 4))]
    ref = [pipe.recognize_packed(p, 5) for p in batches]
    got = list(pipe.recognize_stream(iter(batches), 5))
    assert len(got) == len(ref)
    for (d0, i0), (d1, i1) in zip(ref, got):
        assert np.array_equal(i0, i1) and np.array_equal(d0, d1)


def test_last_block_class_token_only_equals_full_last_block(tmp_path):
    """The engine runs the last encoder block for the class token only (the pooled token).  Against the same engine
    with the full last block (EFFOCR_VIT_LAST_BLOCK_FULL=1, read at first use, hence two subprocesses) the embeddings
    agree to fp16-operand noise, and both stay within the 1e-3 bound of the CPU oracle."""
    import os
    import subprocess
    import sys
    import numpy as np
    script = (
        "import sys, numpy as np, torch; sys.path.insert(0, '.');"
        "from effocr_b200 import synth; from effocr_b200.pipeline import PackedCrops, RecognizerPipeline;"
        "from oracle import vit as OV;"
        "sd = OV.randomize_affine(OV.init_vit_state_dict('vit_small_patch16_224', seed=0));"
        "pipe = RecognizerPipeline(sd, torch.zeros(1, 384), max_batch=64);"
        "crops, _ = synth.synthetic_crops(40, seed=3);"
        "px, im, bx, n = PackedCrops(crops).to_device();"
        "np.save(sys.argv[1], pipe.embed_boxes(px, im, bx, n).cpu().numpy())")
    outs = []
    for full in ("0", "1"):
        path = str(tmp_path / f"emb_{full}.npy")
        env = dict(os.environ, EFFOCR_VIT_LAST_BLOCK_FULL=full)
        subprocess.run([sys.executable, "-c", script, path], check=True, env=env, cwd=str(__import__("pathlib").Path(__file__).resolve().parent.parent))
        outs.append(np.load(path))
    a, b = outs
    rel = np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)
    assert rel.max() < 1e-3, rel.max()
    # and the oracle
    import torch
    from effocr_b200 import synth
    from oracle import transform as OT, vit as OV
    sd = OV.randomize_affine(OV.init_vit_state_dict("vit_small_patch16_224", seed=0))
    crops, _ = synth.synthetic_crops(40, seed=3)
    with torch.no_grad():
        ref = OV.l2_normalize(OV.vit_forward(sd, torch.from_numpy(np.stack([OT.paired_transform(c) for c in crops])))).numpy()
    for e in (a, b):
        assert (np.linalg.norm(e - ref, axis=1) / np.linalg.norm(ref, axis=1)).max() <= 1e-3


def test_three_kernel_layer_equals_separate_kernels(tmp_path):
    """An encoder layer runs as three kernels (norm1 + QKV, attention, block tail).  Against the same engine with the
    separate kernels they replaced (EFFOCR_LN_QKV=0 EFFOCR_BLOCK_TAIL=0: LayerNorm, QKV GEMM, proj_ln, mlp_fused; the
    switches are read at first use, hence two subprocesses) the embeddings agree inside the oracle bound: norm1 + QKV is
    bit-identical, the block tail differs by the order of a few fp32 additions, which flips the fp16 rounding of an
    occasional operand element; over twelve layers that is ~1e-4 of the embedding norm (measured max 2.2e-4)."""
    import os
    import subprocess
    import sys
    script = (
        "import sys, numpy as np, torch; sys.path.insert(0, '.'); sys.path.insert(0, 'tests');"
        "import driver_fixture as DF;"
        "from effocr_b200 import synth; from effocr_b200.pipeline import PackedCrops, RecognizerPipeline;"
        "sd = DF.load_npz_state(DF.VIT_WEIGHTS);"
        "pipe = RecognizerPipeline(sd, torch.zeros(1, 384), max_batch=512);"
        "crops, _ = synth.synthetic_crops(300, seed=11);"  # 300 x 197 rows: several 256-row tiles per CTA pair
        "px, im, bx, n = PackedCrops(crops).to_device();"
        "np.save(sys.argv[1], pipe.embed_boxes(px, im, bx, n).cpu().numpy())")
    root = str(__import__("pathlib").Path(__file__).resolve().parent.parent)
    outs = []
    for fused in ("1", "0"):
        path = str(tmp_path / f"emb_{fused}.npy")
        env = dict(os.environ, EFFOCR_LN_QKV=fused, EFFOCR_BLOCK_TAIL=fused)
        subprocess.run([sys.executable, "-c", script, path], check=True, env=env, cwd=root)
        outs.append(np.load(path))
    a, b = outs
    rel = np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)
    assert rel.max() < 5e-4, rel.max()
