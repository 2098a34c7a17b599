"""Throughput of the non-headline BASELINE configs on one GPU (diagnostics; parity for them is in tests/).
    python tools/config_bench.py [5] [4]
config 5: one 16384x12000 frame, 127x93 = 11,811 units of 96x96 (pitch 128), salt noise, thr 8 / min-area 0 / erode 1.
config 4: 8 frames of the grid.json grid, rect + circle exclusions, centroid shift, erosion radius sweep."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import vi_b200
from vi_b200 import synth
from vi_b200.grid import Grid, generate_grid


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def config5():
    boxes = synth.dense_grid_boxes()
    t0 = time.time()
    frame = synth.make_frame(7, boxes, H=12000, W=16384, inset=8, salt_p=0.02)
    insp = vi_b200.Inspector(0)
    insp.configure(Grid(boxes=[(b, i) for i, b in enumerate(boxes)]), is_reference=True)
    p = vi_b200.default_params(threshold=8, min_area=0, erode_px=1)
    d = torch.from_numpy(frame[None]).cuda()
    out = {}
    def run():
        out['r'] = insp.inspect_batch(d, p)
    ms = timed(run)
    rec = out['r'][0].cpu().numpy().view(vi_b200.RECORD_DTYPE).reshape(-1)
    n = len(boxes)
    from vi_b200 import _lib
    prof = torch.zeros((n, 48), dtype=torch.int64, device="cuda")      # [units][kProfSlots]
    _lib.check(insp._lib.vi_debug_set_profile(insp._ctx, prof.data_ptr()))
    insp.inspect_batch(d, p); torch.cuda.synchronize()
    _lib.check(insp._lib.vi_debug_set_profile(insp._ctx, None))
    m = prof.cpu().numpy().astype(np.float64).mean(axis=0)
    names = ["gather", "blur+hist", "wait for cell min/max", "cells rest+otsu", "threshold", "close/open", "hole fill", "centroid ccl",
             "excl+seg out", "erosion", "roi ccl", "dirty+exact", "open3", "defect hole fill", "area filter", "defect out"]
    sub = {20: "rank V", 21: "rank C", 22: "rank classify", 30: "hist zero", 31: "hist blur3 loop", 33: "warp0 approx thr", 34: "warp0 levels", 36: "finish plane scan", 23: "ccl count+scan", 24: "ccl extract", 25: "ccl link", 26: "ccl jump B",
           27: "ccl unions", 28: "ccl jump D", 29: "thr gray mask"}
    tot = m[:16].sum() + sum(m[k] for k in sub)
    print("  cycles per unit %.0f:" % tot, ", ".join(f"{nm} {m[i]:.0f}" for i, nm in enumerate(names)))
    print("  sub-slots:", ", ".join(f"{nm} {m[k]:.0f}" for k, nm in sub.items()))
    print(f"config 5: {n} units of 96x96 in {ms:.3f} ms -> {n / ms * 1e3:,.0f} units/s, {n * 96 * 96 / ms / 1e6:.2f} Gpix/s (unit px); "
          f"NG {int((rec['status'] == 1).sum())}, kept components mean {rec['n_kept'].mean():.1f}, max runs {rec['n_runs'].max()}, "
          f"ambiguous px mean {rec['n_ambiguous'].mean():.1f}  (frame synthesis {time.time() - t0:.0f} s)")


def config4():
    boxes = generate_grid((251, 232, 316, 315), 4, 6, 2, 1, 133, 136, 252, 0)
    frames = np.stack([synth.make_frame(s, [b for b, _ in boxes]) for s in range(8)])
    excl = [{'shape': 'rect', 'x': 50, 'y': 60, 'w': 70, 'h': 30}, {'shape': 'circle', 'cx': 200, 'cy': 180, 'r': 25}]
    insp = vi_b200.Inspector(0)
    insp.configure(Grid(boxes=boxes, exclusions=excl), is_reference=True)
    d = torch.from_numpy(frames).cuda()
    rec0 = insp.inspect_batch(d[:1])[0].cpu().numpy().view(vi_b200.RECORD_DTYPE).reshape(-1)
    refc = {int(r['unit']): (float(r['cx']), float(r['cy'])) for r in rec0}
    insp.configure(Grid(boxes=boxes, exclusions=excl, ref_centroids=refc), is_reference=False)
    for r in (1, 6, 10, 11, 17, 32, 63):
        p = vi_b200.default_params(erode_px=r)
        ms = timed(lambda: insp.inspect_batch(d, p))
        print(f"config 4: erode r={r:2d}: {8 * 48 / ms * 1e3:,.0f} units/s ({ms:.3f} ms for 384 units = 2.6 units per SM)")


if __name__ == "__main__":
    which = sys.argv[1:] or ["5", "4"]
    if "5" in which: config5()
    if "4" in which: config4()
