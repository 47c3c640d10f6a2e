"""JPEG files for the loader tests, made with PIL's own encoder (the writer of the reference's patch files:
deepzoom_tiler.py saves tiles with PIL, quality 70)."""
import io

import numpy as np
from PIL import Image


def histology_like(h, w, seed):
    """Smooth pink/purple blobs + noise: coefficient statistics closer to a stained-tissue patch than white noise."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.zeros((h, w, 3))
    base = np.array([225.0, 190.0, 215.0])
    img += base
    for _ in range(12):
        cy, cx, r = rng.uniform(0, h), rng.uniform(0, w), rng.uniform(3, max(4, min(h, w) / 4))
        blob = np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * r * r))
        img -= blob[..., None] * rng.uniform(30, 140, 3)
    img += rng.normal(0, 6, (h, w, 3))
    return np.clip(img, 0, 255).astype(np.uint8)


def noise(h, w, seed):
    return (np.random.default_rng(seed).random((h, w, 3)) * 255).astype(np.uint8)


def encode(img, **kw):
    b = io.BytesIO()
    Image.fromarray(img).save(b, format="JPEG", **kw)
    return b.getvalue()


def pil_rgb(data):
    with Image.open(io.BytesIO(data)) as im:
        return np.asarray(im.convert("RGB"))


def patch_files(n, h=224, w=224, quality=70, seed=0, **kw):
    """n patch files as deepzoom_tiler.py writes them (PIL defaults: 4:2:0, standard Huffman tables)."""
    return [encode(histology_like(h, w, seed + i), quality=quality, **kw) for i in range(n)]
