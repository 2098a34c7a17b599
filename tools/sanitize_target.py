"""Small workload for compute-sanitizer: 1 frame x 48 default units (full path), a ragged grid, and the compat calls."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import vi_b200
from vi_b200 import synth
from vi_b200.grid import Grid, generate_grid
boxes = generate_grid((251, 232, 316, 315), 4, 2, 1, 1, 133, 136, 252, 0)          # 8 units
frame = synth.make_frame(3, [b for b, _ in boxes], H=1100, W=2064)
insp = vi_b200.Inspector(0)
excl = [{'shape': 'rect', 'x': 50, 'y': 60, 'w': 70, 'h': 30}, {'shape': 'circle', 'cx': 200, 'cy': 180, 'r': 25}]
insp.configure(Grid(boxes=boxes, exclusions=excl), is_reference=True)
for p in (vi_b200.default_params(), vi_b200.default_params(erode_px=1, threshold=6), vi_b200.default_params(defect_method=1),
          vi_b200.default_params(seg_method=1), vi_b200.default_params(gaussian_blur=5, morph_kernel=5, erode_px=17)):
    rec, _, _ = insp.inspect_batch(torch.from_numpy(frame[None]).cuda(), p)
    torch.cuda.synchronize()
rng = np.random.default_rng(1)
for shape in ((64, 97), (12, 15), (1, 1), (3, 200)):
    im = rng.integers(0, 256, size=shape, dtype=np.uint8)
    insp.segment_cell(im, vi_b200.default_params())
    m = ((rng.random(shape) < 0.5) * 255).astype(np.uint8)
    insp.detect_defects(im, m, vi_b200.default_params(erode_px=1, threshold=5, min_area=0))
print("done", int(rec.sum()))
