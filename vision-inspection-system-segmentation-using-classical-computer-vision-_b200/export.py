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
