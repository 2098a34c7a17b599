"""Per-phase SM-cycle breakdown of the fused kernel (diagnostics).
    python tools/phase_profile.py [n_images]
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import vi_b200
from vi_b200 import synth, _lib
from vi_b200.grid import Grid, generate_grid

NAMES = ["gather", "blur+hist", "wait for cell min/max", "cells rest+otsu wait", "threshold", "close/open", "hole fill",
         "centroid ccl", "excl+seg out", "erosion", "roi ccl", "dirty+exact", "open3", "defect hole fill", "area filter",
         "defect out"]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    boxes = generate_grid((251, 232, 316, 315), 4, 6, 2, 1, 133, 136, 252, 0)
    frames = np.stack([synth.make_frame(s, [b for b, _ in boxes]) for s in range(n)])
    insp = vi_b200.Inspector(0)
    insp.configure(Grid(boxes=boxes), is_reference=True)
    d = torch.from_numpy(frames).cuda()
    prof = torch.zeros((n * 48, 48), dtype=torch.int64, device="cuda")      # [units][kProfSlots]
    insp.inspect_batch(d)
    torch.cuda.synchronize()
    _lib.check(insp._lib.vi_debug_set_profile(insp._ctx, prof.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rec, _, _ = insp.inspect_batch(d)
    e1.record()
    torch.cuda.synchronize()
    _lib.check(insp._lib.vi_debug_set_profile(insp._ctx, None))
    p = prof.cpu().numpy().astype(np.float64)
    rec = rec.cpu().numpy().view(vi_b200.RECORD_DTYPE).reshape(-1)
    mean = p.mean(axis=0)
    tot = mean[:len(NAMES)].sum()
    print(f"{n} images, {n*48} units, kernel {e0.elapsed_time(e1):.3f} ms; mean cycles/unit {mean[:37].sum() + mean[44] + mean[45] + mean[47]:.0f} (phase slots {tot:.0f} + sub-phase slots {mean[len(NAMES):37].sum() + mean[44] + mean[45] + mean[47]:.0f}; slots 37-43 run beside them on the Otsu warp, 46 is a count)")
    for i, nm in enumerate(NAMES):
        print(f"  {i:2d} {nm:18s} {mean[i]:10.0f}  {100*mean[i]/tot:5.1f}%   max {p[:, i].max():10.0f}")
    for i, nm in ((20, "rank: V"), (21, "rank: C"), (22, "rank: classify"), (23, "ccl: count+scan"), (24, "ccl: extract"), (25, "ccl: link"), (26, "ccl: jump B"), (27, "ccl: unions"), (28, "ccl: jump D"), (29, "thr: gray mask"), (30, "hist: zero"), (31, "hist: blur3 loop"), (32, "gather: wait for rows"), (33, "warp0: approx thr"), (34, "warp0: levels+tables"), (35, "finish: zero cand"), (36, "finish: plane scan"), (37, "otsu warp: scan"), (38, "otsu: p, ip"), (39, "otsu: q1 sums"), (40, "otsu: y"), (41, "otsu: chain"), (42, "otsu: sigma, arg max"), (43, "otsu: bins (count)"), (44, "finish: plane loads (thread 0)"), (45, "finish: thread 0 roi tests"), (46, "thr: listed band pixels (count, x16 warps)"), (47, "thr: band pass A")):
        print(f"  {i:2d} {nm:18s} {mean[i]:10.0f}")
    print("  gather cycles: first unit of each CTA %.0f, later units %.0f" % (p[:148, 0].mean(), p[148:, 0].mean()))
    print("  n_ambiguous mean %.1f max %d; n_runs max %d" % (rec['n_ambiguous'].mean(), rec['n_ambiguous'].max(), rec['n_runs'].max()))


if __name__ == "__main__":
    main()
