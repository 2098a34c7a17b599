import sys, collections, hashlib
import numpy as np
sys.path.insert(0, '.')
import cv2
from oracle import ref_cv2 as R, restate as S
rng = np.random.default_rng(3)
rng.integers(0, 256, size=(64, 97), dtype=np.uint8)
im = rng.integers(0, 256, size=(12, 15), dtype=np.uint8)
hh = lambda a: hashlib.md5(a.tobytes()).hexdigest()[:6]
k = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (2, 2))
b = cv2.GaussianBlur(im.copy(), (5, 5), 0)
t, th = cv2.threshold(b, 0, 255, cv2.THRESH_BINARY_INV + cv2.THRESH_OTSU)
c = cv2.morphologyEx(th, cv2.MORPH_CLOSE, k)
o = cv2.morphologyEx(c, cv2.MORPH_OPEN, k)
cnt = collections.Counter(); cnt2 = collections.Counter(); cnt3 = collections.Counter()
for rep in range(3000):
    cnt[hh(R.fill_internal_holes(o))] += 1                  # flood fill alone on a fixed input
    cnt2[hh(R.segment_cell(im, gaussian_blur=4, morph_kernel=2))] += 1
    m2 = cv2.morphologyEx(cv2.morphologyEx(th.copy(), cv2.MORPH_CLOSE, k), cv2.MORPH_OPEN, k)
    cnt3[hh(m2)] += 1
print("fill alone", dict(cnt)); print("segment_cell", dict(cnt2)); print("close+open alone", dict(cnt3))
print("restatement", hh(S.segment_cell(im, gaussian_blur=4, morph_kernel=2)))
