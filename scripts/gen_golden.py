#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ (run in the BUILD container only).

Sources of truth, in order of authority:
  * the reference's own src/orbextractor.cpp, compiled unmodified against oracle/cvshim
    (oracle/_ref/liborbref.so, `make -C oracle ref`; needs /root/reference) and run under the
    monotone bump allocator ("canonical" tie order, see oracle/cvshim/ref_glue.cpp);
  * cv2 4.13 for the five OpenCV primitives the reference calls (resize, copyMakeBorder, FAST,
    GaussianBlur, fastAtan2) and cv2.BFMatcher(NORM_HAMMING) for kNN-2.
Neither travels to the GPU box; the fixtures do.
"""
import ctypes as C
import hashlib
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "opendlv-perception-vision-orbslam2_b200")]
import orb_oracle_py as O  # noqa: E402
import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
cv2.setNumThreads(1)


class Cfg(C.Structure):
    _fields_ = [("nfeatures", C.c_int), ("scale", C.c_float), ("nlevels", C.c_int), ("ini", C.c_int), ("min", C.c_int)]


R = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "liborbref.so"))
R.orbref_extract.restype = C.c_int
R.orbref_extract.argtypes = [C.POINTER(Cfg), C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p,
                             C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]


def ref_extract(img, nf, nl, canonical=1, levels=False):
    cfg = Cfg(nf, 1.2, nl, 20, 7)
    cap = nf + 256
    kps = np.zeros(cap, O.KP_DTYPE); desc = np.zeros((cap, 32), np.uint8)
    lv = None
    if levels:
        bufs = [np.zeros(img.size, np.uint8) for _ in range(nl)]
        ptrs = (C.c_void_p * nl)(*[b.ctypes.data for b in bufs])
        lw = (C.c_int * nl)(); lh = (C.c_int * nl)()
        n = R.orbref_extract(C.byref(cfg), canonical, img.ctypes.data, img.shape[1], img.shape[0], img.strides[0],
                             kps.ctypes.data, desc.ctypes.data, cap, ptrs, lw, lh)
        lv = [bufs[l][:lw[l] * lh[l]].reshape(lh[l], lw[l]).copy() for l in range(nl)]
    else:
        n = R.orbref_extract(C.byref(cfg), canonical, img.ctypes.data, img.shape[1], img.shape[0], img.strides[0],
                             kps.ctypes.data, desc.ctypes.data, cap, None, None, None)
    assert n >= 0
    return kps[:n].copy(), desc[:n].copy(), lv


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ---- 1. OpenCV primitives from cv2 ---------------------------------------------------------
rng = np.random.default_rng(20261018)
img = synth.scene_s1(97, 61, 42)
noise = rng.integers(0, 256, (61, 97), dtype=np.uint8)
prim = {"img": img, "noise": noise}
prim["resize_img"] = cv2.resize(img, (81, 51), interpolation=cv2.INTER_LINEAR)
prim["resize_noise"] = cv2.resize(noise, (81, 51), interpolation=cv2.INTER_LINEAR)
prim["border_noise"] = cv2.copyMakeBorder(noise, 19, 19, 19, 19, cv2.BORDER_REFLECT_101)
prim["blur_img"] = cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
prim["blur_noise"] = cv2.GaussianBlur(noise, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
for name, im in (("img", img), ("noise", noise)):
    for th in (20, 7):
        k = cv2.FastFeatureDetector_create(th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16).detect(im)
        prim[f"fast{th}_{name}"] = np.array([[kp.pt[0], kp.pt[1], kp.response] for kp in k], np.int32).reshape(-1, 3)
yx = rng.integers(-300000, 300000, (256, 2)).astype(np.float32)
yx[:4] = [[0, 0], [0, 5], [5, 0], [-3, -3]]
prim["atan2_yx"] = yx
prim["atan2_out"] = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in yx], np.float32)
np.savez_compressed(os.path.join(OUT, "cv2_primitives.npz"), **prim)

# ---- 2. whole extractor from the reference's own translation unit -----------------------------
ext = {}
tiny = synth.scene_s1(320, 200, 31337)
k, d, lv = ref_extract(tiny, 300, 4, 1, levels=True)
ext["tiny_img"] = tiny; ext["tiny_kps"] = k; ext["tiny_desc"] = d
for l, a in enumerate(lv):
    ext[f"tiny_level{l}"] = a
ks, ds, _ = ref_extract(tiny, 300, 4, 0)
ext["tiny_kps_stockmalloc"] = ks
kitti = synth.scene_s1(1241, 376, 1000)
k, d, lv = ref_extract(kitti, 2000, 8, 1, levels=True)
ext["kitti_img_sha256"] = np.array(sha(kitti))
ext["kitti_kps"] = k; ext["kitti_desc"] = d
ext["kitti_level_sha256"] = np.array([sha(a) for a in lv])
ext["kitti_level_shapes"] = np.array([a.shape for a in lv], np.int32)
noisef = synth.scene_s2(640, 360, 11)
k, d, _ = ref_extract(noisef, 1000, 6, 1)
ext["noise_img_sha256"] = np.array(sha(noisef))
ext["noise_kps"] = k; ext["noise_desc"] = d
np.savez_compressed(os.path.join(OUT, "ref_extractor.npz"), **ext)

# ---- 3. matcher: reference loop semantics cross-checked with cv2.BFMatcher ---------------------
q, t = synth.matching_set(64, 500, seed=123)
t[10] = q[3]; t[400] = q[3]
bf = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2)
idx = np.array([m[0].trainIdx for m in bf], np.int32)
d1 = np.array([m[0].distance for m in bf], np.int32)
d2 = np.array([m[1].distance for m in bf], np.int32)
oi, o1, o2 = O.knn2(q, t)
assert np.array_equal(idx, oi) and np.array_equal(d1, o1) and np.array_equal(d2, o2), "oracle loop != cv2.BFMatcher"
np.savez_compressed(os.path.join(OUT, "matcher.npz"), q=q, t=t, idx=idx, d1=d1, d2=d2)
for f in sorted(os.listdir(OUT)):
    print(f, os.path.getsize(os.path.join(OUT, f)))
