"""The consumer contract of the mask/CSV export (indexing_ui.py:2703-2730) on batch results: the rows of
`masks_summary.csv` without re-reading the PNGs -- `mask_stats` (segmentation.py:103-111) of every unit's final
segmentation mask, from the (area, sum x, sum y) the kernel can emit next to the masks."""
from __future__ import annotations

import csv
from typing import Iterable, List

import numpy as np

FIELDS = ['index', 'mask', 'area', 'centroid_x', 'centroid_y']


def masks_summary_rows(seg_stats) -> List[dict]:
    """seg_stats: int64 [n_units, 3] of one image.  Same rows as export_masks_and_csv builds (indexing_ui.py:2716-2722):
    empty masks give area 0 and centroid (0, 0) (segmentation.py:106-107); else float means of x and y."""
    rows = []
    for i, (area, sx, sy) in enumerate(np.asarray(seg_stats, dtype=np.int64).reshape(-1, 3).tolist()):
        cx, cy = (0, 0) if area == 0 else (float(sx / area), float(sy / area))
        rows.append({'index': i, 'mask': f'mask_{i:04d}.png', 'area': int(area), 'centroid_x': cx, 'centroid_y': cy})
    return rows


def write_masks_summary_csv(path: str, rows: Iterable[dict]) -> None:
    """indexing_ui.py:2723-2729."""
    with open(path, 'w', newline='') as cf:
        writer = csv.DictWriter(cf, fieldnames=FIELDS)
        writer.writeheader()
        for row in rows:
            writer.writerow(row)


def write_mask_png(path: str, packed_rows, w: int, h: int) -> None:
    """A unit's mask as a 1-bit grayscale PNG straight from the packed-bit output of the kernel (VI_MASKS_PACKED:
    rows of 32-bit words, first pixel = most significant bit of its byte -- a PNG scanline of bit depth 1 is exactly
    the first ceil(w / 8) bytes of such a row).  Decoders expand it to 0 / 255, so `mask_stats` of the re-read file
    (export_masks_and_csv, indexing_ui.py:2703-2722) sees the same mask as from the reference's 8-bit `mask_%04d.png`."""
    import struct
    import zlib
    rows = np.asarray(packed_rows, dtype=np.uint8).reshape(h, -1)
    nb = (w + 7) // 8
    if rows.shape[1] < nb:
        raise ValueError(f"packed rows hold {rows.shape[1]} bytes, a {w}-pixel scanline needs {nb}")
    raw = np.zeros((h, nb + 1), np.uint8)                   # filter type 0 in front of every scanline
    raw[:, 1:] = rows[:, :nb]
    if w % 8:
        raw[:, nb] &= np.uint8((0xFF << (8 - w % 8)) & 0xFF)   # padding bits of the last byte are zero

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)
    with open(path, 'wb') as f:
        f.write(b"\x89PNG\r\n\x1a\n")
        f.write(chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 1, 0, 0, 0, 0)))
        f.write(chunk(b"IDAT", zlib.compress(raw.tobytes(), 6)))
        f.write(chunk(b"IEND", b""))
