"""Generate tests/golden/*.npz by running the REFERENCE'S OWN CODE.

Run in the build container only (reads /root/reference):

    python tests/golden/make_golden.py

* ``segmentation.py`` is imported unmodified (its Qt import is guarded).
* ``indexing_ui.py`` is imported unmodified on top of ``fakeqt`` (a numpy-backed
  stand-in for PyQt6) and the real ``MainWindow`` methods are executed:
  ``import_grid`` / ``update_grid_preview`` / ``populate_thumbnails`` /
  ``run_segmentation_all`` / ``test_defect_detection_all`` / ``run_inspection``.
  Nothing of the reference is copied: it is executed where it lies.

Inputs are regenerated from seeds by ``vi_b200.synth`` (sha256 of every frame
is stored so generator drift is caught).  Masks are stored bit-packed.
Recorded environment: cv2 / numpy versions (the reference pins neither).
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import fakeqt  # noqa: E402

fakeqt.install()
sys.path.insert(0, REF)
import cv2  # noqa: E402
import indexing_ui  # noqa: E402  (the unmodified reference UI module)
import segmentation as ref_seg  # noqa: E402  (the unmodified reference module)

from vi_b200 import synth  # noqa: E402

ROLE = indexing_ui.ROLE_BASE
ref_seg.qimage_to_gray_array = lambda q: q.arr.copy()   # mono identity, SURVEY A.1


class _Dialog:
    path = None

    @staticmethod
    def getOpenFileName(*a, **k):
        return _Dialog.path, ''

    @staticmethod
    def getSaveFileName(*a, **k):
        return _Dialog.path, ''


indexing_ui.QtWidgets.QFileDialog = _Dialog


def make_window(params):
    """A MainWindow without __init__ (which builds the widget tree): only the
    attributes the hot-path methods read, with the reference's defaults."""
    mw = object.__new__(indexing_ui.MainWindow)
    S = fakeqt.Spin
    mw.img_widget = fakeqt.Stub()
    iw = type('IW', (), {})()
    iw.grid_rects = []
    iw.image = None
    iw.fixed_img_rect = None
    iw.selected_cell_index = None
    iw.inspection_mode = False
    iw.inspection_results = {}
    iw.update = lambda *a, **k: None
    mw.img_widget = iw
    mw.thumb_list = fakeqt.QListWidget()
    mw.seg_method = S(params.get('seg_method', 'otsu'))
    mw.gauss_spin = S(params.get('gaussian_blur', 3))
    mw.morph_spin = S(params.get('morph_kernel', 3))
    mw.adapt_block = S(params.get('adapt_block', 51))
    mw.adapt_C = S(params.get('adapt_C', 10))
    mw.defect_method = S(params.get('defect_method', 'threshold'))
    mw.defect_threshold = S(params.get('threshold', 24))
    mw.defect_min_area = S(params.get('min_area', 20))
    mw.defect_mask_erode = S(params.get('erode_px', 6))
    mw.overlay_mode = S('Both')
    for n in ('units_x', 'units_y', 'blocks_x', 'blocks_y', 'unit_space_x', 'unit_space_y',
              'block_space_x', 'block_space_y', 'excl_index', 'defect_unit_spin'):
        setattr(mw, n, S(0))
    mw.exclusions = []
    mw._exclusion_ref_centroids = {}
    mw._image_states = {}
    mw._reference_image_path = None
    mw._current_image_path = None
    mw.logs = []
    mw.log = lambda msg: mw.logs.append(str(msg))
    sb = fakeqt.Stub()
    mw.statusBar = lambda: sb
    for n in ('refresh_thumbnail_icons', 'refresh_canvas_overlays', '_snapshot_current_results',
              'update_selected_overlay', 'on_exclusion_index_changed', 'exit_inspection_mode'):
        setattr(mw, n, lambda *a, **k: None)
    return mw


def set_image(mw, frame, path):
    mw.img_widget.image = fakeqt.QImage(frame)
    mw._current_image_path = path
    if mw._reference_image_path is None:
        mw._reference_image_path = path
    mw.populate_thumbnails()            # real reference method (:3096-3125)


def collect(mw, role):
    out = []
    for i in range(mw.thumb_list.count()):
        pm = mw.thumb_list.item(i).data(role)
        out.append(pm.arr.copy() if isinstance(pm, fakeqt.QPixmap) else None)
    return out


def pack_masks(masks, shape_list):
    """-> (present u8[n], packed object array of packbits rows)"""
    present = np.array([m is not None for m in masks], np.uint8)
    blobs = []
    for m, (h, w) in zip(masks, shape_list):
        if m is None:
            m = np.zeros((h, w), np.uint8)
        assert m.shape == (h, w) and set(np.unique(m)) <= {0, 255}
        blobs.append(np.packbits(m > 0))
    return present, np.concatenate(blobs) if blobs else np.zeros(0, np.uint8)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def parse_areas(logs):
    """'Unit {row}: defect area={area} px -> {verdict}' (indexing_ui.py:1619)."""
    areas, verdicts = {}, {}
    for ln in logs:
        if ln.startswith('Unit ') and 'defect area=' in ln:
            row = int(ln.split(':')[0][5:])
            areas[row] = int(ln.split('defect area=')[1].split(' px')[0])
            verdicts[row] = ln.rsplit('-> ', 1)[1]
    return areas, verdicts


def run_case(name, frames, boxes_src, params, exclusions=(), erode_list=None, H=3000, W=4096,
             synth_kw=None, max_units=None):
    """frames: list of seeds; frames[0] is the reference image.  boxes_src:
    ('json', path) -> real import_grid; ('list', [(x,y,w,h),...]) -> written to a
    legacy bare-list JSON and imported through the real import_grid."""
    synth_kw = synth_kw or {}
    mw = make_window(params)
    kind, src = boxes_src
    if kind == 'json':
        with open(src) as f:
            gj = json.load(f)
        if max_units:
            gj['boxes'] = gj['boxes'][:max_units]
        tmp = os.path.join('/tmp', f'golden_{name}_grid.json')
        with open(tmp, 'w') as f:
            json.dump(gj, f)
    else:
        tmp = os.path.join('/tmp', f'golden_{name}_grid.json')
        with open(tmp, 'w') as f:
            json.dump([{'x': b[0], 'y': b[1], 'w': b[2], 'h': b[3]} for b in src], f)
    out = {}
    frame_hashes = []
    for fi, seed in enumerate(frames):
        # boxes are needed by the generator before import_grid has run once
        with open(tmp) as f:
            raw = json.load(f)
        raw_boxes = raw['boxes'] if isinstance(raw, dict) else raw
        gen_boxes = [(b['x'], b['y'], b['w'], b['h']) for b in raw_boxes]
        frame = synth.make_frame(seed, gen_boxes, H=H, W=W, **synth_kw)
        frame_hashes.append(sha(frame))
        set_image(mw, frame, f'img{seed}.png')
        if fi == 0:
            _Dialog.path = tmp
            mw.import_grid()                        # real (:2831-2934)
            mw.exclusions = [dict(e) for e in exclusions]
        boxes = [(r, idx) for r, idx in mw.img_widget.grid_rects]
        shapes = [(r[3], r[2]) for r, _ in boxes]
        mw.run_segmentation_all()                   # real (:2203-2368)
        seg = collect(mw, ROLE + 1)
        p, blob = pack_masks(seg, shapes)
        assert p.all()
        out[f'f{fi}_seg'] = blob
        if fi == 0:
            rc = mw._exclusion_ref_centroids
            out['ref_centroids'] = np.array([[rc[i][0], rc[i][1]] if i in rc else [np.nan, np.nan]
                                             for i in range(len(boxes))], np.float64)
        for r in (erode_list or [params.get('erode_px', 6)]):
            mw.defect_mask_erode.setValue(int(r))
            mw.logs.clear()
            mw.test_defect_detection_all()          # real (:1574-1632): logs areas
            areas, verdicts = parse_areas(mw.logs)
            mw.img_widget.inspection_mode = False
            ok = mw.run_inspection()                # real (:1634-1709)
            assert ok
            dm = collect(mw, ROLE + 2)
            p, blob = pack_masks(dm, shapes)
            out[f'f{fi}_r{r}_def_present'] = p
            out[f'f{fi}_r{r}_def'] = blob
            res = mw.img_widget.inspection_results
            out[f'f{fi}_r{r}_ng'] = np.array([int(bool(res[idx])) for _, idx in boxes], np.uint8)
            out[f'f{fi}_r{r}_area'] = np.array([areas.get(i, 0) for i in range(len(boxes))], np.int64)
            for i in range(len(boxes)):
                assert (verdicts.get(i) == 'NG') == bool(res[boxes[i][1]]) or i not in verdicts
    out['boxes'] = np.array([[r[0], r[1], r[2], r[3], idx] for r, idx in boxes], np.int32)
    out['meta'] = np.array(json.dumps(dict(
        name=name, seeds=list(frames), H=H, W=W, params=params, exclusions=list(exclusions),
        erode_list=list(erode_list or [params.get('erode_px', 6)]), synth_kw=synth_kw,
        frame_sha256=frame_hashes, cv2=cv2.__version__, numpy=np.__version__)))
    path = os.path.join(HERE, f'{name}.npz')
    np.savez_compressed(path, **out)
    print(name, 'units', len(boxes), 'frames', len(frames), os.path.getsize(path), 'bytes')


def stage_goldens():
    """Direct calls into the unmodified reference segmentation.py on seeded crops
    and hand-built masks (stage-level known answers)."""
    rng = np.random.default_rng(1234)
    out = {}
    boxes = [(8, 8, 316, 315)]
    crops = []
    for seed in range(6):
        fr = synth.make_frame(100 + seed, boxes, H=331, W=332)
        crops.append(fr[8:8 + 315, 8:8 + 316].copy())
    crops.append(rng.integers(0, 256, size=(64, 97), dtype=np.uint8))            # uniform noise
    crops.append(np.full((40, 50), 123, np.uint8))                               # constant
    g = np.tile(np.linspace(0, 255, 120).astype(np.uint8), (90, 1))             # ramp
    crops.append(g)
    small = rng.integers(0, 256, size=(12, 15), dtype=np.uint8)                  # smaller than median window
    crops.append(small)
    cfgs = [dict(), dict(gaussian_blur=5, morph_kernel=5), dict(gaussian_blur=0, morph_kernel=0),
            dict(gaussian_blur=4, morph_kernel=2), dict(gaussian_blur=7, morph_kernel=7),
            dict(gaussian_blur=31, morph_kernel=31), dict(method='adaptive'),
            dict(method='adaptive', adapt_block=11, adapt_C=-3), dict(method='bogus')]
    n = 0
    for ci, c in enumerate(crops):
        out[f'crop{ci}'] = c if c.size < 20000 else np.zeros(0, np.uint8)  # big ones regenerate from seed
        for ki, kw in enumerate(cfgs):
            m = ref_seg.segment_cell(c, **kw)
            out[f'seg_c{ci}_k{ki}'] = np.packbits(m > 0)
            st = ref_seg.mask_stats(m)
            out[f'stats_c{ci}_k{ki}'] = np.array([st['area'], st['centroid'][0], st['centroid'][1]], np.float64)
            n += 1
    # hole-fill known answers on adversarial masks
    masks = []
    m = np.zeros((40, 40), np.uint8); m[5:35, 5:35] = 255; m[10:20, 10:20] = 0; m[12:15, 12:15] = 255
    masks.append(m)                                              # ring with island
    m = np.zeros((30, 30), np.uint8); m[::2, ::2] = 255; m[1::2, 1::2] = 255
    masks.append(m)                                              # checkerboard (8-conn fg, 4-conn bg)
    m = np.full((25, 33), 255, np.uint8); m[8:12, 0:10] = 0; m[15:18, 15:20] = 0
    masks.append(m)                                              # full frame, notch to border + hole
    m = np.zeros((20, 20), np.uint8)
    for i in range(4, 16):
        m[i, 4] = m[i, 15] = m[4, i] = m[15, i] = 255
    m[4, 4] = 0; m[5, 5] = 255
    masks.append(m)                                              # ring closed only diagonally
    masks.append((rng.random((120, 130)) < 0.55).astype(np.uint8) * 255)
    masks.append((rng.random((64, 64)) < 0.5).astype(np.uint8) * 7)   # non-0/255 foreground values
    masks.append(np.zeros((9, 9), np.uint8))
    masks.append(np.full((9, 9), 255, np.uint8))
    masks.append(np.zeros((1, 17), np.uint8))
    for mi, m in enumerate(masks):
        out[f'hole_in{mi}'] = m
        out[f'hole_out{mi}'] = ref_seg.fill_internal_holes(m)
    out['meta'] = np.array(json.dumps(dict(cfgs=cfgs, n_crops=len(crops), n_masks=len(masks),
                                           crop_seeds=[100 + s for s in range(6)],
                                           cv2=cv2.__version__, numpy=np.__version__)))
    path = os.path.join(HERE, 'stages.npz')
    np.savez_compressed(path, **out)
    print('stages', n, 'segment_cell cases', len(masks), 'hole masks', os.path.getsize(path), 'bytes')


def main():
    grid_json = os.path.join(REF, 'grid.json')
    stage_goldens()
    # config 1: one 4096x3000 frame, the repo's grid.json, defaults
    run_case('config1', [0], ('json', grid_json), {})
    # config 4 (subset): exclusions + centroid shift + erosion sweep; frame 0 = reference image
    excl = [{'shape': 'rect', 'x': 50, 'y': 60, 'w': 70, 'h': 30},
            {'shape': 'circle', 'cx': 200, 'cy': 180, 'r': 25}]
    run_case('config4', [0, 1, 2], ('json', grid_json), {}, exclusions=excl,
             erode_list=[1, 2, 6, 17, 40, 63], max_units=8)
    # config 5 (subset): dense small units, salt noise, low threshold, min_area 0
    boxes5 = synth.dense_grid_boxes(W=1088, H=832, unit=96, origin=32, pitch=128)
    run_case('config5', [7], ('list', boxes5), dict(threshold=8, min_area=0, erode_px=1),
             H=832, W=1088, synth_kw=dict(inset=8, salt_p=0.02))
    # non-default branches
    run_case('canny', [3], ('json', grid_json), dict(defect_method='canny'), max_units=8)
    run_case('adaptive', [4], ('json', grid_json), dict(seg_method='adaptive'), max_units=8)
    run_case('blur5_morph5', [5], ('json', grid_json), dict(gaussian_blur=5, morph_kernel=5), max_units=8)


if __name__ == '__main__':
    main()
