"""Synthetic mold-frame generator.

Host-side numpy only.  Used by bench.py, the tests and the golden-vector
script so that every arm (CUDA path, CPU oracle, reference arm) sees the same
bytes for the same seed.  Not part of the reference: the reference ships no
data; the spec of these frames is SURVEY.md Appendix D / section 8(d).
"""
from __future__ import annotations

import numpy as np

DEFAULT_W = 4096
DEFAULT_H = 3000


def make_frame(seed, boxes, H=DEFAULT_H, W=DEFAULT_W, inset=24, jitter=3,
               max_discs=3, salt_p=0.0):
    """One mono uint8 frame: bright field N(200,6), one dark plate N(70,5) per
    unit rect (inset `inset` px, jittered by U{-jitter..jitter}), 0..max_discs
    bright discs per plate (foreign material), optional salt noise inside
    plates (config 5).  `boxes` is an iterable of (x, y, w, h)."""
    rng = np.random.default_rng(int(seed))
    img = np.clip(200.0 + rng.normal(0.0, 6.0, size=(H, W)), 0, 255).astype(np.uint8)
    for (x, y, w, h) in boxes:
        jx = int(rng.integers(-jitter, jitter + 1))
        jy = int(rng.integers(-jitter, jitter + 1))
        y0 = y + inset + jy
        y1 = y + h - inset + jy
        x0 = x + inset + jx
        x1 = x + w - inset + jx
        y0c, y1c = max(0, y0), min(H, y1)
        x0c, x1c = max(0, x0), min(W, x1)
        if y1c <= y0c or x1c <= x0c:
            continue
        ph, pw = y1c - y0c, x1c - x0c
        plate = np.clip(70.0 + rng.normal(0.0, 5.0, size=(ph, pw)), 0, 255).astype(np.uint8)
        if salt_p > 0.0:
            salt = rng.random(size=(ph, pw)) < salt_p
            plate[salt] = 255
        img[y0c:y1c, x0c:x1c] = plate
        n = int(rng.integers(0, max_discs + 1))
        for _ in range(n):
            margin = 5
            if pw <= 2 * margin or ph <= 2 * margin:
                break
            cx = int(rng.integers(x0c + margin, x1c - margin))
            cy = int(rng.integers(y0c + margin, y1c - margin))
            rad = int(rng.integers(2, 8))
            val = int(rng.integers(130, 255))
            ya, yb = max(0, cy - rad), min(H, cy + rad + 1)
            xa, xb = max(0, cx - rad), min(W, cx + rad + 1)
            yy, xx = np.ogrid[ya:yb, xa:xb]
            disc = (xx - cx) ** 2 + (yy - cy) ** 2 <= rad * rad
            img[ya:yb, xa:xb][disc] = val
    return img


def dense_grid_boxes(W=16384, H=12000, unit=96, origin=32, pitch=128):
    """Config-5 grid: unit x unit rects on a `pitch` lattice."""
    boxes = []
    y = origin
    while y + unit <= H:
        x = origin
        while x + unit <= W:
            boxes.append((x, y, unit, unit))
            x += pitch
        y += pitch
    return boxes
