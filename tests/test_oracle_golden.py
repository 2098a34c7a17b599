"""The cv2 oracle (oracle/ref_cv2.py) against outputs of the reference's own
code (tests/golden/*.npz, see make_golden.py).  CPU only."""
import json

import numpy as np
import pytest

from oracle import ref_cv2 as R

CASES = ['config1', 'config4', 'config5', 'canny', 'adaptive', 'blur5_morph5']


def _params(meta):
    p = meta['params']
    return R.Params(**{k: v for k, v in p.items()})


@pytest.mark.parametrize('name', CASES)
def test_pipeline_matches_reference(golden, name):
    g = golden(name)
    meta = g.meta
    refc = {i: (float(c[0]), float(c[1])) for i, c in enumerate(g.z['ref_centroids']) if not np.isnan(c[0])}
    for fi in range(len(meta['seeds'])):
        frame = g.frame(fi)
        seg_gold = g.masks(f'f{fi}_seg')
        for r in meta['erode_list']:
            p = _params(meta)
            p.erode_px = r
            recs, segs, defs = R.inspect_frame(frame, g.boxes, p, meta['exclusions'],
                                               ref_centroids=(None if fi == 0 else refc),
                                               is_reference=(fi == 0))
            dgold = g.defects(fi, r)
            ng = g.z[f'f{fi}_r{r}_ng']
            area = g.z[f'f{fi}_r{r}_area']
            for i in range(len(g.boxes)):
                assert np.array_equal(segs[i], seg_gold[i]), (name, fi, r, i, 'seg')
                assert (defs[i] is None) == (dgold[i] is None), (name, fi, r, i)
                if defs[i] is not None:
                    assert np.array_equal(defs[i], dgold[i]), (name, fi, r, i, 'defect')
                assert recs[i]['defect_area'] == area[i]
                assert (recs[i]['status'] == R.STATUS_NG) == bool(ng[i])
            if fi == 0:
                for i, rec in enumerate(recs):
                    assert rec['cx'] == g.z['ref_centroids'][i][0] and rec['cy'] == g.z['ref_centroids'][i][1]
                    assert rec['dx'] == 0 and rec['dy'] == 0


def test_stage_goldens():
    from vi_b200 import synth
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'stages.npz'))
    meta = json.loads(str(z['meta']))
    for ci in range(meta['n_crops']):
        c = z[f'crop{ci}']
        if c.size == 0:
            fr = synth.make_frame(meta['crop_seeds'][ci], [(8, 8, 316, 315)], H=331, W=332)
            c = fr[8:8 + 315, 8:8 + 316].copy()
        for ki, kw in enumerate(meta['cfgs']):
            m = R.segment_cell(c, **kw)
            assert np.array_equal(np.packbits(m > 0), z[f'seg_c{ci}_k{ki}']), (ci, kw)
            st = R.mask_stats(m)
            assert [st['area'], st['centroid'][0], st['centroid'][1]] == list(z[f'stats_c{ci}_k{ki}'])
    for mi in range(meta['n_masks']):
        assert np.array_equal(R.fill_internal_holes(z[f'hole_in{mi}']), z[f'hole_out{mi}']), mi


def test_grid_generator_reproduces_grid_json(golden):
    g = golden('config1')
    grid = R.generate_grid((251, 232, 316, 315), 4, 6, 2, 1, 133, 136, 252, 0)
    assert grid == g.boxes


def test_fill_holes_error_behaviour():
    assert R.fill_internal_holes(None) is None
    with pytest.raises(ValueError):
        R.fill_internal_holes(np.zeros((2, 2, 2), np.uint8))
    assert R.mask_stats(np.zeros((4, 4), np.uint8)) == {'area': 0, 'centroid': (0, 0)}
