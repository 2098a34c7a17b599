"""Property tests (hypothesis) of the integer restatements against the reference's own cv2 calls: random small masks
and images find the corner cases hand-built ones miss.  CPU only."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
hyp = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st
from hypothesis.extra import numpy as hnp

from oracle import ref_cv2 as R
from oracle import restate as S

SETTINGS = dict(max_examples=200, deadline=None)


def masks(max_side=24):
    shapes = st.tuples(st.integers(1, max_side), st.integers(1, max_side))
    return shapes.flatmap(lambda s: hnp.arrays(np.bool_, s)).map(lambda a: a.astype(np.uint8) * 255)


def images(max_side=40):
    shapes = st.tuples(st.integers(1, max_side), st.integers(1, max_side))
    return shapes.flatmap(lambda s: hnp.arrays(np.uint8, s))


@settings(**SETTINGS)
@given(masks())
def test_fill_holes_is_the_reference_flood(m):
    assert np.array_equal(S.fill_holes_4bg(m), R.fill_internal_holes(m))


@settings(**SETTINGS)
@given(masks(), st.integers(0, 12), st.integers(0, 400))
def test_contour_filter_is_findcontours(m, min_area, seg_area):
    cnts, _ = cv2.findContours(m.copy(), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    ref = np.zeros_like(m)
    max_area = max(min_area, int(seg_area * 0.98))
    found = 0
    for c in cnts:
        if min_area <= cv2.contourArea(c) <= max_area:
            cv2.drawContours(ref, [c], -1, 255, -1)
            found += 1
    out, n = S.contour_free_filter(m, min_area, seg_area)
    assert n == found
    assert (out is None) if found == 0 else np.array_equal(out, ref)


@settings(**SETTINGS)
@given(masks())
def test_largest_component_follows_cv2_label_order(m):
    nlab, labels, stats, _ = cv2.connectedComponentsWithStats((m > 0).astype(np.uint8), connectivity=8)
    lc = S.largest_component(m)
    if nlab <= 1:
        assert lc is None
        return
    best = 1 + int(np.argmax(stats[1:, cv2.CC_STAT_AREA]))
    assert np.array_equal(lc[0], labels == best)
    assert S.largest_component_centroid(m) == R.largest_component_centroid(m)


@settings(**SETTINGS)
@given(masks(), st.integers(0, 9))
def test_square_erosion_is_cv2_iterations(m, r):
    ref = m if r == 0 else cv2.erode(m, None, iterations=r)
    assert np.array_equal(S.erode_square(m, r), ref)


@settings(max_examples=60, deadline=None)
@given(images(30), st.integers(0, 255), st.lists(st.integers(0, 254), min_size=6, max_size=6))
def test_lattice_rank_decision_is_the_median_residual(im, thr, levels):
    direct = cv2.absdiff(im, cv2.medianBlur(im, 21)) > thr
    assert np.array_equal(S.residual_mask_lattice(im, thr, levels), direct)


@settings(**SETTINGS)
@given(images(48))
def test_otsu_scan_is_opencvs_on_non_degenerate_histograms(im):
    ipp = cv2.ipp.useIPP()
    cv2.ipp.setUseIPP(False)                       # OpenCV's own C++ scan (exact ties: see test_restate_vs_cv2.py)
    try:
        t = int(cv2.threshold(im, 0, 255, cv2.THRESH_BINARY_INV + cv2.THRESH_OTSU)[0])
    finally:
        cv2.ipp.setUseIPP(ipp)
    assert S.otsu_from_hist(np.bincount(im.ravel(), minlength=256)) == t


@settings(**SETTINGS)
@given(images(40), st.sampled_from([3, 5, 7, 9, 15]))
def test_gaussian_fixed_point_is_cv2(im, k):
    assert np.array_equal(S.gaussian_blur_u8(im, k), cv2.GaussianBlur(im, (k, k), 0))
