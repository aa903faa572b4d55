"""CPU: each OpenCV primitive restated in the oracle against a live cv2 (skipped where cv2 is absent)."""
import numpy as np
import pytest

import synth

cv2 = pytest.importorskip("cv2")
cv2.setNumThreads(1)


def _pyr_sizes(p, w, h, n):
    return [(int(np.rint(np.float32(w) * np.float32(p.inv_sf[l]))), int(np.rint(np.float32(h) * np.float32(p.inv_sf[l])))) for l in range(n)]


@pytest.mark.parametrize("w,h,n", [(1241, 376, 8), (640, 480, 8), (517, 291, 4)])
def test_resize_chain(oracle, w, h, n):
    p = oracle.Extractor(1000, 1.2, n).params
    for gen in (synth.scene_s1, synth.scene_s2):
        prev = gen(w, h, 5)
        for (lw, lh) in _pyr_sizes(p, w, h, n)[1:]:
            a = oracle.resize(prev, lw, lh)
            b = cv2.resize(prev, (lw, lh), interpolation=cv2.INTER_LINEAR)
            assert np.array_equal(a, b)
            prev = b


def test_border_and_blur(oracle):
    for im in (synth.scene_s1(333, 211, 1), synth.scene_s2(200, 64, 2), synth.scene_s2(9, 8, 3)):
        assert np.array_equal(oracle.border101(im, 19, 19, 19, 19) if min(im.shape) > 19 else oracle.border101(im, 3, 3, 3, 3),
                              cv2.copyMakeBorder(im, *([19] * 4 if min(im.shape) > 19 else [3] * 4), cv2.BORDER_REFLECT_101))
        assert np.array_equal(oracle.gaussian7(im), cv2.GaussianBlur(im, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101))


def _cv_fast(im, th):
    k = cv2.FastFeatureDetector_create(th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16).detect(im)
    return [(int(p.pt[0]), int(p.pt[1]), int(p.response)) for p in k]


def test_fast_whole_images_and_cells(oracle):
    rng = np.random.default_rng(8)
    s1 = synth.scene_s1(640, 360, 4); s2 = synth.scene_s2(320, 200, 4)
    for im in (s1, s2):
        for th in (20, 7):
            xs, ys, sc = oracle.fast9(im, th)
            assert list(zip(xs.tolist(), ys.tolist(), sc.tolist())) == _cv_fast(im, th)
    for i in range(60):   # cell-sized sub-images incl. ones too small to hold any tested pixel
        ww, hh = int(rng.integers(5, 46)), int(rng.integers(5, 46))
        src = s1 if i % 2 else s2
        x0, y0 = int(rng.integers(0, src.shape[1] - ww)), int(rng.integers(0, src.shape[0] - hh))
        sub = np.ascontiguousarray(src[y0:y0 + hh, x0:x0 + ww])
        for th in (20, 7):
            xs, ys, sc = oracle.fast9(sub, th)
            assert list(zip(xs.tolist(), ys.tolist(), sc.tolist())) == _cv_fast(sub, th)


def test_threshold_fallback_equivalence(oracle):
    """NMS(20) == {k in NMS(7): score >= 20} -- what lets the GPU kernel score each cell once."""
    s1 = synth.scene_s1(400, 300, 12)
    rng = np.random.default_rng(1)
    for _ in range(80):
        x0, y0 = int(rng.integers(0, 360)), int(rng.integers(0, 260))
        sub = np.ascontiguousarray(s1[y0:y0 + 38, x0:x0 + 37])
        a = list(zip(*[v.tolist() for v in oracle.fast9(sub, 20)]))
        b = [k for k in zip(*[v.tolist() for v in oracle.fast9(sub, 7)]) if k[2] >= 20]
        assert a == b


def test_fast_atan2(oracle):
    rng = np.random.default_rng(2)
    for y, x in rng.integers(-250000, 250000, (5000, 2)):
        assert np.float32(oracle.fast_atan2(float(y), float(x))) == np.float32(cv2.fastAtan2(float(y), float(x)))


def test_knn2_matches_bfmatcher(oracle):
    q, t = synth.matching_set(100, 3000, seed=4)
    t[7] = q[0]; t[2000] = q[0]
    bf = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2)
    idx, d1, d2 = oracle.knn2(q, t)
    assert idx.tolist() == [m[0].trainIdx for m in bf]
    assert d1.tolist() == [int(m[0].distance) for m in bf]
    assert d2.tolist() == [int(m[1].distance) for m in bf]
