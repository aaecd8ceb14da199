"""Deterministic synthetic inputs for tests and bench.py: rendered text-line
images with ground-truth char / word boxes, character crops cut from them, and glyph renders for a
prototype index.  The reference's TTF files live under /root/reference (absent on the GPU box), so
lines are drawn with Pillow's bundled scalable default font (FreeType); everything is seeded.
"""
from __future__ import annotations

import numpy as np

ASCII_GLYPHS = [chr(c) for c in range(33, 127)]  # the "94-glyph" printable-ASCII charset of BASELINE.json


_FONT_FILES: list = []   # set_font_dir(): TTF/OTF files used round-robin instead of Pillow's default font
_font_cache: dict = {}


def set_font_dir(path) -> int:
    """Render with the TTF / OTF files below `path` (e.g. the reference's english_font_files/, round-robin per line, as
    SURVEY.md section 8d describes) instead of Pillow's bundled default font.  `None` restores the default.  Returns the
    number of font files found.  The committed golden fixtures and quick-fit weights were made with the DEFAULT font:
    tests never call this; bench.py does when --font-dir is given and says so in its `data` field."""
    import glob
    import os

    _FONT_FILES.clear()
    _font_cache.clear()
    if path:
        for ext in ("*.ttf", "*.otf", "*.TTF", "*.OTF"):
            _FONT_FILES.extend(sorted(glob.glob(os.path.join(str(path), "**", ext), recursive=True)))
    return len(_FONT_FILES)


def _font(size: int, which: int = 0):
    from PIL import ImageFont

    if _FONT_FILES:
        key = (_FONT_FILES[which % len(_FONT_FILES)], size)
        if key not in _font_cache:
            _font_cache[key] = ImageFont.truetype(key[0], size=size)
        return _font_cache[key]
    return ImageFont.load_default(size=size)


def render_line(text: str, height: int = 64, width: int = 1024, font_size: int = 40, x0: int = 6, tracking: float = 0.0,
                font_index: int = 0):
    """-> (u8 [height, width, 3] RGB, char_boxes [n,4] float32 xyxy, word_boxes [m,4], chars list).
    `tracking`: extra pixels between consecutive glyphs (letter-spacing).  The reference's localizer runs its NMS at
    IoU 0.01 (infer_effocr_onnx_multi.py:441), which only keeps neighbouring characters whose boxes do not touch."""
    from PIL import Image, ImageDraw

    font = _font(font_size, font_index)
    img = Image.new("RGB", (width, height), (255, 255, 255))
    draw = ImageDraw.Draw(img)
    x = float(x0)
    char_boxes, chars, word_boxes = [], [], []
    word_start = None
    for ch in text:
        adv = font.getlength(ch)
        if ch == " ":
            if word_start is not None:
                word_boxes.append([word_start, 0.0, x, float(height)])
                word_start = None
            x += adv
            continue
        if x + adv >= width - 2:
            break
        draw.text((x, 4), ch, font=font, fill=(0, 0, 0))
        l, t, r, b = draw.textbbox((x, 4), ch, font=font)
        char_boxes.append([float(l), float(max(t, 0)), float(max(r, l + 1)), float(min(b, height))])
        chars.append(ch)
        if word_start is None:
            word_start = float(l)
        x += adv + tracking
    if word_start is not None:
        word_boxes.append([word_start, 0.0, x, float(height)])
    return (np.asarray(img, dtype=np.uint8).copy(), np.asarray(char_boxes, dtype=np.float32).reshape(-1, 4),
            np.asarray(word_boxes, dtype=np.float32).reshape(-1, 4), chars)


def random_text(rng: np.random.Generator, n_glyphs: int) -> str:
    words, left = [], n_glyphs
    while left > 0:
        k = int(min(left, rng.integers(1, 9)))
        words.append("".join(ASCII_GLYPHS[int(i)] for i in rng.integers(0, len(ASCII_GLYPHS), k)))
        left -= k
    return " ".join(words)


def synthetic_lines(n: int, seed: int = 0, height: int = 64, width: int = 1024, tracking: float = 0.0):
    """n rendered lines, 20-40 glyphs each (fewer if the line fills up)."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        text = random_text(rng, int(rng.integers(20, 41)))
        out.append(render_line(text, height, width, font_size=int(rng.integers(34, 44)), tracking=tracking, font_index=i))
    return out


def synthetic_crops(n: int, seed: int = 0):
    """n u8 crops [64, w, 3] (full line height, w = the glyph's box width) cut from rendered lines
    with the renderer's boxes -- the shape of crop the reference feeds its transform (double clipping)."""
    crops, labels = [], []
    s = seed
    while len(crops) < n:
        for img, cb, _wb, chars in synthetic_lines(8, seed=s):
            for box, ch in zip(cb, chars):
                x0, x1 = int(round(float(box[0]))), int(round(float(box[2])))
                if x1 > x0:
                    crops.append(np.ascontiguousarray(img[:, x0:x1, :]))
                    labels.append(ch)
                if len(crops) == n:
                    return crops, labels
        s += 1000003
    return crops, labels


def glyph_image(index: int, size: int = 64) -> np.ndarray:
    """Deterministic synthetic 'glyph' number `index`: the first 94 are the printable ASCII glyphs,
    later ones overlay two or three ASCII glyphs at index-derived offsets (distinct shapes, same ink
    statistics) -- stands in for a large CJK charset when no CJK font is available."""
    from PIL import Image, ImageDraw

    font = _font(44)
    img = Image.new("RGB", (size, size), (255, 255, 255))
    draw = ImageDraw.Draw(img)
    n = len(ASCII_GLYPHS)
    parts = [index % n]
    rest = index // n
    while rest > 0:
        parts.append(rest % n)
        rest //= n
    for j, p in enumerate(parts):
        dx = 6 + (j * 11 + (index >> (3 * j)) % 7) % 20
        dy = 2 + (j * 7 + (index >> (2 * j)) % 5) % 10
        draw.text((dx, dy), ASCII_GLYPHS[p], font=font, fill=(0, 0, 0))
    return np.asarray(img, dtype=np.uint8).copy()


def glyph_images(n: int, size: int = 64):
    return [glyph_image(i, size) for i in range(n)]


def random_yolov5s_state_dict(nc: int = 2, seed: int = 0, obj_bias: float | None = None, char_bias: float = 2.0,
                              obj_gain: float = 1.0):
    """Random-init YOLOv5s (ultralytics `model.{i}...` keys; width 0.50 / depth 0.33) for benchmarks and smoke
    runs -- there is no network for checkpoints.  Kaiming-uniform convolutions, near-identity BatchNorm statistics,
    Detect biases a la ultralytics (`obj <- log(8 / (640 / s)^2)` unless `obj_bias` is given); `char_bias` is added to
    the class-0 (character) logit so that, as with a trained EffOCR localizer, most surviving boxes are characters;
    `obj_gain` scales the objectness rows of the Detect convolutions (a random-init head gives every location nearly
    the same confidence; a gain of ~25 spreads the scores like a trained detector's, so box counts are not
    hyper-sensitive to the threshold)."""
    import math

    import torch

    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(prefix, c1, c2, k):
        bound = math.sqrt(3.0 / (c1 * k * k))
        sd[prefix + "conv.weight"] = (torch.rand(c2, c1, k, k, generator=g) * 2 - 1) * bound
        sd[prefix + "bn.weight"] = 1.0 + 0.1 * torch.randn(c2, generator=g)
        sd[prefix + "bn.bias"] = 0.1 * torch.randn(c2, generator=g)
        sd[prefix + "bn.running_mean"] = 0.1 * torch.randn(c2, generator=g)
        sd[prefix + "bn.running_var"] = 1.0 + 0.2 * torch.rand(c2, generator=g)

    def c3(i, c1, c2, n):
        p, c_ = f"model.{i}.", c2 // 2
        conv(p + "cv1.", c1, c_, 1)
        conv(p + "cv2.", c1, c_, 1)
        conv(p + "cv3.", 2 * c_, c2, 1)
        for j in range(n):
            conv(p + f"m.{j}.cv1.", c_, c_, 1)
            conv(p + f"m.{j}.cv2.", c_, c_, 3)

    conv("model.0.", 3, 32, 6)
    conv("model.1.", 32, 64, 3)
    c3(2, 64, 64, 1)
    conv("model.3.", 64, 128, 3)
    c3(4, 128, 128, 2)
    conv("model.5.", 128, 256, 3)
    c3(6, 256, 256, 3)
    conv("model.7.", 256, 512, 3)
    c3(8, 512, 512, 1)
    conv("model.9.cv1.", 512, 256, 1)
    conv("model.9.cv2.", 1024, 512, 1)
    conv("model.10.", 512, 256, 1)
    c3(13, 512, 256, 1)
    conv("model.14.", 256, 128, 1)
    c3(17, 256, 128, 1)
    conv("model.18.", 128, 128, 3)
    c3(20, 256, 256, 1)
    conv("model.21.", 256, 256, 3)
    c3(23, 512, 512, 1)
    no = 5 + nc
    strides = (8, 16, 32)
    for l, (c, s) in enumerate(zip((128, 256, 512), strides)):
        bound = 1.0 / math.sqrt(c)
        wdet = (torch.rand(3 * no, c, 1, 1, generator=g) * 2 - 1) * bound
        wdet.view(3, no, c)[:, 4] *= obj_gain
        sd[f"model.24.m.{l}.weight"] = wdet
        b = ((torch.rand(3 * no, generator=g) * 2 - 1) * bound).view(3, no)
        b[:, 4] += math.log(8 / (640 / s) ** 2) if obj_bias is None else obj_bias
        b[:, 5:] += math.log(0.6 / (nc - 0.99999))
        b[:, 5] += char_bias
        sd[f"model.24.m.{l}.bias"] = b.reshape(-1)
    anchors = torch.tensor((((10, 13), (16, 30), (33, 23)), ((30, 61), (62, 45), (59, 119)), ((116, 90), (156, 198), (373, 326))),
                           dtype=torch.float32)
    sd["model.24.anchors"] = anchors / torch.tensor(strides, dtype=torch.float32).view(3, 1, 1)
    return sd


def background_suppressed_yolo_state(nc: int = 2, seed: int = 0, obj_gain: float = 25.0, background_logit: float = -5.0,
                                     input_shape=(640, 640)):
    """random_yolov5s_state_dict whose objectness biases are shifted so that the uniform grey letterbox padding scores
    `background_logit` at every scale / anchor (a random-init head otherwise gives the padding -- 94 % of a
    letterboxed 64 x 1024 line -- one constant confidence that thousands of boxes tie on).  Needs the GPU engine:
    the padding response is measured by one forward pass over an all-grey image."""
    import torch

    from .localizer_engine import YoloEngine

    sd = random_yolov5s_state_dict(nc=nc, seed=seed, obj_bias=0.0, obj_gain=obj_gain)
    eng = YoloEngine(sd, max_batch=1, max_shape=tuple(input_shape))
    grey = torch.full((1, 3, input_shape[0], input_shape[1]), 114.0 / 255.0, device="cuda", dtype=torch.float32)
    pred = eng.forward(grey)[0].float().cpu()
    del eng
    no, pos = 5 + nc, 0
    for l, s in enumerate((8, 16, 32)):
        ny, nx = input_shape[0] // s, input_shape[1] // s
        p = pred[pos:pos + 3 * ny * nx, 4].view(3, ny, nx)
        pos += 3 * ny * nx
        inner = p[:, ny // 4:ny - ny // 4, nx // 4:nx - nx // 4].reshape(3, -1).median(dim=1).values.clamp(1e-6, 1 - 1e-6)
        logit = torch.log(inner / (1 - inner))
        sd[f"model.24.m.{l}.bias"].view(3, no)[:, 4] += background_logit - logit
    return sd
