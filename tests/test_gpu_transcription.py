"""GPU: transcription parity with a TRAINED recogniser (tests/golden/quickfit_vit_d2.npz, made by
tools/quickfit_recognizer.py): with real margins between glyph classes the B200 path must produce exactly the
oracle's strings (CER vs ref = 0, BASELINE.json metric), and both must actually read the rendered text."""
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden" / "quickfit_vit_d2.npz"


def _setup():
    from effocr_b200 import synth
    from oracle import transform as OT, vit as OV
    sd = {k: torch.from_numpy(v.astype(np.float32)) for k, v in np.load(GOLDEN).items()}
    glyphs = synth.ASCII_GLYPHS
    protos = [synth.render_line(ch, font_size=40, x0=6, width=128) for ch in glyphs]
    pc = [np.ascontiguousarray(im[:, int(round(float(cb[0][0]))):int(round(float(cb[0][2]))), :]) for im, cb, _, _ in protos]
    with torch.no_grad():
        xb = OV.l2_normalize(OV.vit_forward(sd, torch.from_numpy(np.stack([OT.paired_transform(c) for c in pc]))))
    return sd, xb, glyphs


@pytest.mark.skipif(not GOLDEN.exists(), reason="quick-fit weights not generated")
def test_transcriptions_identical_to_oracle_and_readable():
    from effocr_b200 import synth, textproc
    from effocr_b200.infer import crop_rect_torch_path
    from effocr_b200.pipeline import PackedCrops, RecognizerPipeline
    from oracle import knn as OK, transform as OT, vit as OV
    sd, xb, glyphs = _setup()
    pipe = RecognizerPipeline(sd, xb, candidate_chars=glyphs, max_batch=512)
    lines = synth.synthetic_lines(16, seed=77)
    images = [l[0] for l in lines]
    rects, per_line = [], []
    for li, (img, cb, wb, chars) in enumerate(lines):
        # renderer boxes stand in for the localizer (mmdet-style [x0,y0,x1,y1,score]); torch-path semantics
        cbs = [list(map(float, b)) + [1.0] for b in cb]
        wbs = [list(map(float, b)) + [1.0] for b in wb]
        sc, wei = textproc.en_preprocess(cbs, wbs, score_thresh=0.5, score_thresh_word=0.5)
        r = [crop_rect_torch_path(b, img.shape[0], img.shape[1], vertical=False) for b in sc]
        per_line.append((sc, wei, len(r), "".join(chars)))
        rects += [(li,) + x for x in r]
    # ---- B200 path: one crop launch over all lines, ViT, L2 norm, exact kNN (k = 10 as infer_effocr.py:112)
    dist, idx = pipe.recognize_packed(PackedCrops(images, rects), k=10)
    # ---- oracle path on the same rectangles
    crops = [images[li][y0:y1, x0:x1, :] for (li, x0, y0, x1, y1) in rects]
    with torch.no_grad():
        emb = OV.l2_normalize(OV.vit_forward(sd, torch.from_numpy(np.stack([OT.paired_transform(c) for c in crops]))))
    rd, ri = OK.flat_ip_search(xb, emb, 10)
    _, margin = OK.margins(xb, emb, 1)
    assert margin.median() > 1e-2  # trained weights: margins three orders of magnitude above the random-init case
    pos, pairs_ref, pairs_gt = 0, [], []
    for (sc, wei, n, gt) in per_line:
        heights = [b[3] - b[1] for b in sc]
        bottoms = [b[3] for b in sc]
        got = [glyphs[j] for j in idx[pos:pos + n, 0]]
        exp = [glyphs[int(j)] for j in ri[pos:pos + n, 0]]
        t_got = textproc.en_postprocess(got, wei, heights, bottoms)
        t_exp = textproc.en_postprocess(exp, wei, heights, bottoms)
        pairs_ref.append((t_exp, t_got))
        pairs_gt.append((gt, (t_got or "").replace(" ", "")))
        pos += n
    acc, cer_vs_ref = textproc.textline_evaluation(pairs_ref)
    assert cer_vs_ref == 0.0 and acc == 100.0, [p for p in pairs_ref if p[0] != p[1]][:3]
    # top-10 neighbour lists agree wherever the oracle's ordering is decidable: here the two arms embed with
    # different arithmetic (fp16 operands vs fp32), so consecutive scores must differ by 4 x the 1e-3 tolerance
    _, m10 = OK.margins(xb, emb, 10)
    dec = (m10 > 4e-3).numpy()
    assert dec.any() and np.array_equal(idx[dec], ri.numpy()[dec])
    assert np.array_equal(idx[:, 0], ri.numpy()[:, 0])  # top-1 ids identical for every character
    _, cer_gt = textproc.textline_evaluation(pairs_gt, no_spaces_in_eval=True)
    assert cer_gt < 0.15, cer_gt  # the quick-fit recogniser reads the synthetic lines


def test_ad_hoc_index_on_device_matches_oracle_index(tmp_path):
    """SURVEY 8f N2 (`--ad_hoc_index_root_dir`, infer_effocr.py:190-201): a folder of rendered glyphs -> prototypes embedded
    on the device (fused crop kernel + encoder + L2 norm) -> new index.  Against the oracle's index of the same renders
    (per-glyph transform -> ViT -> normalise, what InferenceModel.train_knn computes): same characters in the same order,
    prototypes within the embedding tolerance, and recognition through the new index reads held-out crops like the
    oracle's index does."""
    from PIL import Image
    from effocr_b200 import lineio, synth
    from effocr_b200.pipeline import RecognizerPipeline
    from oracle import knn as OK, transform as OT, vit as OV
    vit_s = GOLDEN.parent / "quickfit_vit_small.npz"  # the full-depth quick-fit ViT-S (tools/quickfit_recognizer.py)
    if not vit_s.exists():
        pytest.skip("quick-fit ViT-S weights not generated")
    sd = {k: torch.from_numpy(v.astype(np.float32)) for k, v in np.load(vit_s).items()}
    glyphs = synth.ASCII_GLYPHS[:40]
    renders = {}
    for ch in glyphs:
        im, cb, _wb, _c = synth.render_line(ch, font_size=40, x0=6, width=128)
        crop = np.ascontiguousarray(im[:, int(round(float(cb[0][0]))):int(round(float(cb[0][2]))), :])
        cls = f"0x{ord(ch):x}"
        (tmp_path / cls).mkdir()
        Image.fromarray(crop).save(tmp_path / cls / f"{cls}_NotoSerif-Regular.png")
        Image.fromarray(crop[:, ::-1]).save(tmp_path / cls / f"{cls}_OtherFont.png")           # another font: ignored
        Image.fromarray(crop).save(tmp_path / cls / f"PAIRED_{cls}_NotoSerif-Regular_0.png")   # a scanned crop: ignored
        renders[ch] = crop
    pipe = RecognizerPipeline(sd, torch.zeros(1, 384), max_batch=16)  # max_batch < 40: chunked index construction
    chars = lineio.build_ad_hoc_index(str(tmp_path), pipe, lang="en")
    order = sorted(glyphs, key=lambda c: f"0x{ord(c):x}")  # ImageFolder walks the class directories in sorted order
    assert chars == order == pipe.candidate_chars and pipe.index.ntotal == 40
    with torch.no_grad():
        ref = OV.l2_normalize(OV.vit_forward(sd, torch.from_numpy(np.stack([OT.paired_transform(renders[c]) for c in order]))))
    got = torch.from_numpy(pipe.index.reconstruct_n())
    rel = ((got - ref).norm(dim=1) / ref.norm(dim=1)).max().item()
    print(f"ad-hoc index: max relative prototype error {rel:.2e}")
    assert rel <= 1e-3, rel
    crops, labels = synth.synthetic_crops(400, seed=9)
    keep = [i for i, l in enumerate(labels) if l in order][:128]
    _d, idx = pipe.recognize_crops([crops[i] for i in keep], k=1)
    with torch.no_grad():
        q = OV.l2_normalize(OV.vit_forward(sd, torch.from_numpy(np.stack([OT.paired_transform(crops[i]) for i in keep]))))
    _rd, ri = OK.flat_ip_search(ref, q, 1)
    assert np.array_equal(idx[:, 0], ri.numpy()[:, 0])
    assert np.mean([order[j] == labels[i] for i, j in zip(keep, idx[:, 0])]) > 0.8
