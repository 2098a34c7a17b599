"""Every SURVEY Appendix-A equivalence as a permanent test: the cv2-free
integer restatement (oracle/restate.py, the form the kernels implement) equals
the cv2 arm (oracle/ref_cv2.py, pinned to the reference by the goldens)."""
import cv2
import numpy as np
import pytest

from oracle import ref_cv2 as R
from oracle import restate as S
from vi_b200 import synth


def _crops(n=4, seed0=200):
    out = []
    for s in range(n):
        fr = synth.make_frame(seed0 + s, [(8, 8, 316, 315)], H=331, W=332)
        out.append(fr[8:323, 8:324].copy())
    return out


def _adversarial_masks():
    rng = np.random.default_rng(5)
    ms = []
    ms.append(np.zeros((17, 23), np.uint8))
    ms.append(np.full((17, 23), 255, np.uint8))
    m = np.zeros((40, 40), np.uint8); m[5:35, 5:35] = 255; m[10:30, 10:30] = 0; m[15:25, 15:25] = 255; m[18:22, 18:22] = 0
    ms.append(m)                                    # nested rings with island
    m = np.zeros((30, 30), np.uint8); np.fill_diagonal(m, 255); ms.append(m)   # diagonal line
    m = np.zeros((30, 31), np.uint8); m[::2, ::2] = 255; m[1::2, 1::2] = 255; ms.append(m)  # checkerboard
    m = np.zeros((30, 30), np.uint8); m[0, :] = m[-1, :] = 255; m[:, 0] = m[:, -1] = 255; m[10:14, 10:14] = 255; ms.append(m)
    m = np.zeros((33, 35), np.uint8); m[16, :] = 255; m[:, 17] = 255; ms.append(m)    # crossing lines
    m = np.zeros((9, 9), np.uint8); m[4, 4] = 255; ms.append(m)                      # single pixel
    m = np.zeros((12, 12), np.uint8); m[2:5, 2:5] = 255; m[5:8, 5:8] = 255; ms.append(m)  # diagonal touch
    m = np.zeros((20, 20), np.uint8); m[4:16, 4:16] = 255; m[6:14, 6:14] = 0; m[9, 4:6] = 0; ms.append(m)  # ring with 4-conn gap
    m = np.zeros((20, 20), np.uint8)
    for i in range(4, 16):
        m[i, 4] = m[i, 15] = m[4, i] = m[15, i] = 255
    m[4, 4] = 0; m[5, 5] = 255; ms.append(m)        # thin ring closed only diagonally
    ms.append((rng.random((200, 210)) < 0.02).astype(np.uint8) * 255)
    ms.append((rng.random((150, 160)) < 0.55).astype(np.uint8) * 255)
    ms.append((rng.random((64, 70)) < 0.3).astype(np.uint8) * 255)
    # two equal-area components: tie-break by cv2 label order
    m = np.zeros((20, 40), np.uint8); m[10:14, 2:6] = 255; m[3:7, 30:34] = 255; ms.append(m)
    m = np.zeros((21, 40), np.uint8); m[5:9, 30:34] = 255; m[4:8, 2:6] = 255; m[4, 2] = 0; m[8, 5] = 255; ms.append(m)
    return ms


@pytest.mark.parametrize('k', list(range(0, 33)))
def test_gaussian_bit_exact(k):
    rng = np.random.default_rng(k)
    for img in (rng.integers(0, 256, size=(61, 47), dtype=np.uint8), _crops(1)[0],
                rng.integers(0, 256, size=(5, 4), dtype=np.uint8)):
        kk = k if k % 2 == 1 else k + 1
        if k == 0:
            assert np.array_equal(S.gaussian_blur_u8(img, 0), img)
            continue
        assert np.array_equal(S.gaussian_blur_u8(img, k), cv2.GaussianBlur(img, (kk, kk), 0)), k


def test_adaptive_taps_and_mask():
    """float32 Gaussian mean, the +-1 LSB class.  Taps are bit-equal to cv2's for every block size of the
    widget range; the mask restatement may differ from cv2 only at .5 ties of the float mean (cv2's scalar tail columns,
    numpy's emulated FMA): stated bound 1e-4 of the pixels, measured here 0."""
    for bs in range(3, 202, 2):
        assert np.array_equal(S.gaussian_kernel_f32(bs), cv2.getGaussianKernel(bs, 0, cv2.CV_32F).ravel()), bs
    rng = np.random.default_rng(2)
    imgs = [_crops(1)[0], rng.integers(0, 256, size=(61, 47), dtype=np.uint8),
            cv2.GaussianBlur(rng.integers(0, 256, size=(90, 100), dtype=np.uint8), (9, 9), 0)]
    bad = tot = 0
    for bs, C in ((51, 10), (11, -3), (3, 0), (201, 10), (21, -50), (9, 2)):
        for im in imgs:
            ref = cv2.adaptiveThreshold(im, 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, cv2.THRESH_BINARY_INV, bs, C)
            got = S.adaptive_inv_mask(im, bs, C)
            bad += int((got != ref).sum()); tot += ref.size
    assert bad <= 1e-4 * tot, (bad, tot)


def test_canny_restatement_is_cv2():
    """indexing_ui.py:1537: integer, so bit-exact; cv2's result does not depend on its thread count."""
    rng = np.random.default_rng(4)
    imgs = [rng.integers(0, 256, size=(120, 130), dtype=np.uint8), _crops(1)[0],
            cv2.GaussianBlur(rng.integers(0, 256, size=(200, 210), dtype=np.uint8), (7, 7), 0),
            rng.integers(0, 256, size=(5, 7), dtype=np.uint8), np.full((9, 9), 77, np.uint8),
            rng.integers(0, 256, size=(1, 1), dtype=np.uint8), rng.integers(0, 256, size=(1, 40), dtype=np.uint8)]
    for thr in (24, 3, 1, 0, 100, 255, 8):
        lo, hi = max(1, thr // 2), max(2, thr)
        for im in imgs:
            ref = cv2.Canny(im, lo, hi)
            assert np.array_equal(S.canny_edges(im, lo, hi), ref), (thr, im.shape)
    n0 = cv2.getNumThreads()
    try:
        cv2.setNumThreads(1)
        one = cv2.Canny(imgs[1], 12, 24)
    finally:
        cv2.setNumThreads(n0)
    assert np.array_equal(one, cv2.Canny(imgs[1], 12, 24))


def test_otsu_matches_cv2():
    rng = np.random.default_rng(0)
    imgs = _crops(4)
    for i in range(120):
        kind = i % 4
        if kind == 0:
            im = rng.integers(0, 256, size=(40, 50), dtype=np.uint8)
        elif kind == 1:
            a, b = rng.integers(0, 256, size=2)
            im = np.where(rng.random((60, 70)) < rng.random(), a, b).astype(np.uint8)
            im = np.clip(im + rng.normal(0, rng.integers(1, 20), im.shape), 0, 255).astype(np.uint8)
        elif kind == 2:
            im = cv2.GaussianBlur(rng.integers(0, 256, size=(80, 80), dtype=np.uint8), (9, 9), 0)
        else:
            im = np.full((10, 10), rng.integers(0, 256), np.uint8)
        imgs.append(im)
    for im in imgs:
        t = R.otsu_threshold(im)
        ts, mask = S.otsu_inv_mask(im)
        assert ts == t
        _, mref = cv2.threshold(im, 0, 255, cv2.THRESH_BINARY_INV + cv2.THRESH_OTSU)
        assert np.array_equal(mask, mref)


@pytest.mark.parametrize('k', list(range(1, 32)))
def test_ellipse_se(k):
    assert np.array_equal(S.ellipse_se(k), cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (k, k)))


@pytest.mark.parametrize('k', [1, 2, 3, 4, 5, 6, 7, 9, 12, 15, 21, 31])
def test_close_open(k):
    se = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (k, k))
    for m in _adversarial_masks():
        ref = cv2.morphologyEx(cv2.morphologyEx(m, cv2.MORPH_CLOSE, se), cv2.MORPH_OPEN, se)
        assert np.array_equal(S.morph_close_open(m, k), ref), (k, m.shape)


@pytest.mark.parametrize('r', [1, 2, 3, 6, 10, 17, 40, 63, 200])
def test_erode_square(r):
    for m in _adversarial_masks() + [R.segment_cell(c) for c in _crops(2)]:
        assert np.array_equal(S.erode_square(m, r), cv2.erode(m, None, iterations=r)), (r, m.shape)


def test_fill_holes():
    for m in _adversarial_masks():
        assert np.array_equal(S.fill_holes_4bg(m), R.fill_internal_holes(m))


def test_largest_component_tiebreak_and_centroid():
    rng = np.random.default_rng(3)
    masks = _adversarial_masks()
    for _ in range(60):
        h, w = rng.integers(2, 40, size=2)
        masks.append((rng.random((h, w)) < rng.uniform(0.2, 0.7)).astype(np.uint8) * 255)
    for m in masks:
        src = (m > 0).astype(np.uint8)
        nlab, labels, stats, _ = cv2.connectedComponentsWithStats(src, connectivity=8)
        lc = S.largest_component(m)
        if nlab <= 1:
            assert lc is None
            continue
        best = 1 + int(np.argmax(stats[1:, cv2.CC_STAT_AREA]))
        assert np.array_equal(lc[0], labels == best)
        assert S.largest_component_centroid(m) == R.largest_component_centroid(m)


def test_median21_and_rank_formulation():
    rng = np.random.default_rng(11)
    imgs = [_crops(1)[0][:120, :130].copy(), rng.integers(0, 256, size=(50, 60), dtype=np.uint8),
            rng.integers(0, 256, size=(12, 15), dtype=np.uint8),
            cv2.GaussianBlur(rng.integers(0, 256, size=(70, 70), dtype=np.uint8), (15, 15), 0)]
    for im in imgs:
        assert np.array_equal(S.median21(im), cv2.medianBlur(im, 21))
        for thr in (0, 3, 8, 24, 100, 255):
            direct = cv2.absdiff(im, cv2.medianBlur(im, 21)) > thr
            for levels in ([], [64, 70, 76, 190, 200, 210], [0, 254], list(range(7, 255, 8))):
                assert np.array_equal(S.residual_mask_rank(im, thr, levels), direct), (thr, levels)


def test_lattice_rank_formulation():
    """The kernels' 3x3-lattice bound (csrc/vi_rank.cuh) is exact for any level set."""
    rng = np.random.default_rng(12)
    imgs = [_crops(1)[0][:100, :115].copy(), rng.integers(0, 256, size=(40, 47), dtype=np.uint8),
            rng.integers(0, 256, size=(12, 15), dtype=np.uint8), rng.integers(60, 90, size=(31, 29), dtype=np.uint8),
            cv2.GaussianBlur(rng.integers(0, 256, size=(64, 70), dtype=np.uint8), (15, 15), 0)]
    for im in imgs:
        for thr in (0, 8, 24, 100):
            direct = cv2.absdiff(im, cv2.medianBlur(im, 21)) > thr
            for levels in ([62, 70, 78, 192, 200, 208], [0, 1, 2, 3, 4, 5], [254] * 6, [10, 60, 110, 160, 210, 250]):
                st = {}
                assert np.array_equal(S.residual_mask_lattice(im, thr, levels, stats=st), direct), (thr, levels)


def test_contour_free_filter():
    rng = np.random.default_rng(2)
    masks = _adversarial_masks()
    for m in list(masks):
        masks.append(S.open_cross3(m))
    masks.append((rng.random((600, 600)) < 0.02).astype(np.uint8) * 255)
    for m in masks:
        for min_area, seg_area in ((0, m.size), (20, m.size), (3, 50), (20, 0)):
            cnts, _ = cv2.findContours(m.copy(), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
            ref = np.zeros_like(m)
            max_area = max(min_area, int(seg_area * 0.98))
            found = 0
            for c in cnts:
                a = cv2.contourArea(c)
                if min_area <= a <= max_area:
                    cv2.drawContours(ref, [c], -1, 255, -1)
                    found += 1
            out, n = S.contour_free_filter(m, min_area, seg_area)
            assert n == found
            if found == 0:
                assert out is None
            else:
                assert np.array_equal(out, ref)


def test_open_cross3():
    se = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (3, 3))
    for m in _adversarial_masks():
        assert np.array_equal(S.open_cross3(m), cv2.morphologyEx(m, cv2.MORPH_OPEN, se))


def test_full_unit_matches_cv2_arm():
    excl = [{'shape': 'rect', 'x': 50, 'y': 60, 'w': 70, 'h': 30}, {'shape': 'circle', 'cx': 200, 'cy': 180, 'r': 25}]
    for ci, gray in enumerate(_crops(3, seed0=300)):
        info = {}
        seg = S.segment_cell(gray, info=info)
        assert np.array_equal(seg, R.segment_cell(gray))
        assert info['otsu_t'] == R.otsu_threshold(cv2.GaussianBlur(gray, (3, 3), 0))
        a = seg.copy(); b = seg.copy()
        S.apply_exclusions(a, excl, 3, -2); R.apply_exclusions(b, excl, 3, -2)
        assert np.array_equal(a, b)
        for r, thr, mn in ((6, 24, 20), (1, 8, 0), (40, 3, 5)):
            i1, i2 = {}, {}
            d_ref = R.detect_defects(gray, a, 'threshold', thr, mn, r, i1)
            d_new = S.detect_defects(gray, a, thr, mn, r, levels=[62, 70, 78, 192, 200, 208], info=i2)
            assert (d_ref is None) == (d_new is None)
            if d_ref is not None:
                assert np.array_equal(d_ref, d_new)
                assert np.array_equal(i1['roi'], i2['roi'])


def test_otsu_exact_ties_follow_opencv_source_not_ipp():
    """On mirror-symmetric histograms two thresholds tie exactly and rounding decides.  The restatement equals
    OpenCV's C++ scan (IPP dispatch off) on every one of them; the closed-source IPP routine that the cv2 wheel
    dispatches to by default (when present) may break such ties differently -- reported, not asserted."""
    import cv2
    from oracle import restate as S
    rng = np.random.default_rng(11)
    cases = []
    for _ in range(150):
        c = int(rng.integers(40, 216)); d = int(rng.integers(5, 40)); s = float(rng.uniform(0.5, 6))
        half = np.clip(rng.normal(c - d, s, 1600), 0, 255).astype(np.uint8)
        cases.append(np.concatenate([half, (2 * c - half.astype(int)).clip(0, 255).astype(np.uint8)]).reshape(40, 80))
    ipp = cv2.ipp.useIPP()
    try:
        cv2.ipp.setUseIPP(False)
        for im in cases:
            t = int(cv2.threshold(im, 0, 255, cv2.THRESH_BINARY_INV + cv2.THRESH_OTSU)[0])
            assert S.otsu_from_hist(np.bincount(im.ravel(), minlength=256)) == t
        cv2.ipp.setUseIPP(True)
        diff = sum(S.otsu_from_hist(np.bincount(im.ravel(), minlength=256)) !=
                   int(cv2.threshold(im, 0, 255, cv2.THRESH_BINARY_INV + cv2.THRESH_OTSU)[0]) for im in cases)
        print(f"exactly tied histograms on which the IPP dispatch differs from OpenCV's own scan: {diff} of {len(cases)}")
    finally:
        cv2.ipp.setUseIPP(ipp)


def test_ingest_restatements_equal_the_reference_calls():
    """Frame ingest: the ARGB32 -> gray path of qimage_to_gray_array (channels reversed, then cv2's
    BGR2GRAY) and the loader's 16-bit rule, restated in integers."""
    rng = np.random.default_rng(5)
    a = rng.integers(0, 256, size=(97, 131, 4), dtype=np.uint8)
    assert np.array_equal(S.gray_from_argb32(a), R.gray_from_argb32(a))
    # every gray level is a fixed point (mono frames pass through unchanged)
    m = np.repeat(np.arange(256, dtype=np.uint8)[None, :, None], 4, axis=2)
    assert np.array_equal(R.gray_from_argb32(m)[0], np.arange(256, dtype=np.uint8))
    assert np.array_equal(S.gray_from_argb32(m)[0], np.arange(256, dtype=np.uint8))
    # channel extremes, where the R/B swap of the reference shows
    ext = np.array([[[255, 0, 0, 255], [0, 255, 0, 255], [0, 0, 255, 255], [255, 255, 255, 0]]], np.uint8)
    assert np.array_equal(S.gray_from_argb32(ext), R.gray_from_argb32(ext))
    assert S.gray_from_argb32(ext)[0].tolist() == [76, 150, 29, 255]          # blue byte takes the red weight
    u = rng.integers(0, 65536, size=(50, 70), dtype=np.uint16)
    u[0, :4] = [0, 255, 256, 65535]
    assert np.array_equal(S.gray8_from_gray16(u), R.gray8_from_gray16(u))
