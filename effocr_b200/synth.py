"""Deterministic synthetic inputs for tests and bench.py: rendered text-line
images with ground-truth char / word boxes, character crops cut from them, and glyph renders for a
prototype index.  The reference's TTF files live under /root/reference (absent on the GPU box), so
lines are drawn with Pillow's bundled scalable default font (FreeType); everything is seeded.
"""
from __future__ import annotations

import numpy as np

ASCII_GLYPHS = [chr(c) for c in range(33, 127)]  # the "94-glyph" printable-ASCII charset of BASELINE.json


def _font(size: int):
    from PIL import ImageFont

    return ImageFont.load_default(size=size)


def render_line(text: str, height: int = 64, width: int = 1024, font_size: int = 40, x0: int = 6):
    """-> (u8 [height, width, 3] RGB, char_boxes [n,4] float32 xyxy, word_boxes [m,4], chars list)."""
    from PIL import Image, ImageDraw

    font = _font(font_size)
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
        x += adv
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


def synthetic_lines(n: int, seed: int = 0, height: int = 64, width: int = 1024):
    """n rendered lines, 20-40 glyphs each (fewer if the line fills up)."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        text = random_text(rng, int(rng.integers(20, 41)))
        out.append(render_line(text, height, width, font_size=int(rng.integers(34, 44))))
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
