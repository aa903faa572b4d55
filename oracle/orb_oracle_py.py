"""ctypes binding of the CPU oracle (oracle/orb_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product library
(liborbx.so) never touches anything in this directory.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liborboracle.so")
MAX_LEVELS = 16


class Keypoint(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("size", C.c_float), ("angle", C.c_float),
                ("response", C.c_float), ("octave", C.c_int32), ("class_id", C.c_int32)]


KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28


class Params(C.Structure):
    _fields_ = [("nfeatures", C.c_int), ("scale_factor", C.c_float), ("nlevels", C.c_int),
                ("ini_th", C.c_int), ("min_th", C.c_int), ("taps", C.c_int * 7),
                ("sf", C.c_float * MAX_LEVELS), ("inv_sf", C.c_float * MAX_LEVELS),
                ("sigma2", C.c_float * MAX_LEVELS), ("inv_sigma2", C.c_float * MAX_LEVELS),
                ("quota", C.c_int * MAX_LEVELS), ("umax", C.c_int * 16)]


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("orb_oracle.c", "orb_oracle.h", "orb_pattern.inc")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return _LIB_PATH
    subprocess.check_call(["make", "-s", "-C", _HERE, "_build/liborboracle.so"])
    return _LIB_PATH


_lib = None
_u8p = C.POINTER(C.c_uint8)
_ip = C.POINTER(C.c_int)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    L.orbo_create.restype = C.c_void_p
    L.orbo_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, _ip]
    L.orbo_destroy.argtypes = [C.c_void_p]
    L.orbo_get_params.restype = C.POINTER(Params)
    L.orbo_get_params.argtypes = [C.c_void_p]
    L.orbo_set_tie_rule.argtypes = [C.c_void_p, C.c_int]
    L.orbo_extract.restype = C.c_int
    L.orbo_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int]
    L.orbo_extract_batch_mt.restype = C.c_int
    L.orbo_extract_batch_mt.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t,
                                        C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    L.orbo_level.restype = C.c_void_p
    L.orbo_level.argtypes = [C.c_void_p, C.c_int, _ip, _ip, C.POINTER(C.c_size_t)]
    L.orbo_blurred.restype = C.c_void_p
    L.orbo_blurred.argtypes = [C.c_void_p, C.c_int, _ip, _ip, C.POINTER(C.c_size_t)]
    L.orbo_candidates.restype = C.c_int
    L.orbo_candidates.argtypes = [C.c_void_p, C.c_int, C.POINTER(_ip), C.POINTER(_ip), C.POINTER(_ip)]
    L.orbo_resize_linear_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_size_t]
    L.orbo_border_reflect101_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_size_t] + [C.c_int] * 4
    L.orbo_fast9_nms.restype = C.c_int
    L.orbo_fast9_nms.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.orbo_grid_fast.restype = C.c_int
    L.orbo_grid_fast.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.orbo_gaussian7_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_size_t, _ip]
    L.orbo_fast_atan2.restype = C.c_float
    L.orbo_fast_atan2.argtypes = [C.c_float, C.c_float]
    L.orbo_distribute.restype = C.c_int
    L.orbo_distribute.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int] + [C.c_int] * 6 + [C.c_void_p, C.c_int]
    L.orbo_ic_angle.restype = C.c_float
    L.orbo_ic_angle.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, _ip]
    L.orbo_rbrief.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_float, C.c_void_p]
    L.orbo_pattern.restype = _ip
    L.orbo_descriptor_distance.restype = C.c_int
    L.orbo_descriptor_distance.argtypes = [C.c_void_p, C.c_void_p]
    L.orbo_knn2.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.orbo_stereo_matches.restype = C.c_int
    L.orbo_stereo_matches.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 7 + \
                                     [C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    L.orbo_knn2_csr.argtypes = [C.c_void_p, C.c_int, C.c_void_p] + [C.c_void_p] * 6
    L.orbo_distinctive.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.orbo_voc_transform.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.orbo_knn2_mt.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _c_u8(img):
    img = np.asarray(img)
    assert img.dtype == np.uint8 and img.ndim == 2
    if img.strides[1] != 1:
        img = np.ascontiguousarray(img)
    return img


def resize(img, dw, dh):
    img = _c_u8(img)
    out = np.empty((dh, dw), np.uint8)
    lib().orbo_resize_linear_u8(_ptr(img), img.shape[1], img.shape[0], img.strides[0], _ptr(out), dw, dh, dw)
    return out


def border101(img, t, b, l, r):
    img = _c_u8(img)
    out = np.empty((img.shape[0] + t + b, img.shape[1] + l + r), np.uint8)
    lib().orbo_border_reflect101_u8(_ptr(img), img.shape[1], img.shape[0], img.strides[0], _ptr(out), out.shape[1], t, b, l, r)
    return out


def fast9(img, th):
    img = _c_u8(img)
    cap = img.size
    xs, ys, sc = (np.empty(cap, np.int32) for _ in range(3))
    n = lib().orbo_fast9_nms(_ptr(img), img.shape[1], img.shape[0], img.strides[0], th, _ptr(xs), _ptr(ys), _ptr(sc), cap)
    assert n >= 0
    return xs[:n].copy(), ys[:n].copy(), sc[:n].copy()


def grid_fast(img, ini_th=20, min_th=7):
    img = _c_u8(img)
    cap = img.size
    xs, ys, sc = (np.empty(cap, np.int32) for _ in range(3))
    n = lib().orbo_grid_fast(_ptr(img), img.shape[1], img.shape[0], img.strides[0], ini_th, min_th, _ptr(xs), _ptr(ys), _ptr(sc), cap)
    if n < 0:
        raise RuntimeError(f"orbo_grid_fast failed: {n}")
    return xs[:n].copy(), ys[:n].copy(), sc[:n].copy()


def gaussian7(img, taps=(18, 34, 48, 56, 48, 34, 18)):
    img = _c_u8(img)
    out = np.empty(img.shape, np.uint8)
    t = (C.c_int * 7)(*taps)
    lib().orbo_gaussian7_u8(_ptr(img), img.shape[1], img.shape[0], img.strides[0], _ptr(out), out.shape[1], t)
    return out


def fast_atan2(y, x):
    return lib().orbo_fast_atan2(float(y), float(x))


def distribute(xs, ys, sc, min_x, max_x, min_y, max_y, n_quota, tie_rule=0):
    xs = np.ascontiguousarray(xs, np.int32); ys = np.ascontiguousarray(ys, np.int32); sc = np.ascontiguousarray(sc, np.int32)
    cap = max(len(xs), 1)
    out = np.empty(cap, np.int32)
    n = lib().orbo_distribute(_ptr(xs), _ptr(ys), _ptr(sc), len(xs), min_x, max_x, min_y, max_y, n_quota, tie_rule, _ptr(out), cap)
    if n < 0:
        raise RuntimeError(f"orbo_distribute failed: {n}")
    return out[:n].copy()


def ic_angle(img, cx, cy, umax):
    img = _c_u8(img)
    um = (C.c_int * 16)(*umax)
    return lib().orbo_ic_angle(_ptr(img), img.strides[0], cx, cy, um)


def rbrief(blurred, cx, cy, angle):
    blurred = _c_u8(blurred)
    d = np.empty(32, np.uint8)
    lib().orbo_rbrief(_ptr(blurred), blurred.strides[0], cx, cy, float(angle), _ptr(d))
    return d


def pattern():
    p = lib().orbo_pattern()
    return np.ctypeslib.as_array(p, shape=(1024,)).copy()


def descriptor_distance(a, b):
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    return lib().orbo_descriptor_distance(_ptr(a), _ptr(b))


def knn2(q, t, nthreads=1):
    q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
    nq, nt = len(q), len(t)
    idx, d1, d2 = (np.empty(nq, np.int32) for _ in range(3))
    if nthreads <= 1:
        lib().orbo_knn2(_ptr(q), nq, _ptr(t), nt, _ptr(idx), _ptr(d1), _ptr(d2))
    else:
        lib().orbo_knn2_mt(_ptr(q), nq, _ptr(t), nt, _ptr(idx), _ptr(d1), _ptr(d2), nthreads)
    return idx, d1, d2


def knn2_csr(q, t, offsets, indices):
    q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
    offsets = np.ascontiguousarray(offsets, np.int32); indices = np.ascontiguousarray(indices, np.int32)
    nq = len(q)
    i1, d1, i2, d2 = (np.empty(nq, np.int32) for _ in range(4))
    lib().orbo_knn2_csr(_ptr(q), nq, _ptr(t), _ptr(offsets), _ptr(indices), _ptr(i1), _ptr(d1), _ptr(i2), _ptr(d2))
    return i1, d1, i2, d2


def distinctive(desc, offsets, indices):
    """OrbMapPoint::ComputeDistinctiveDescriptors for every CSR list -> (best position in the list, its median)."""
    desc = np.ascontiguousarray(desc, np.uint8)
    offsets = np.ascontiguousarray(offsets, np.int32); indices = np.ascontiguousarray(indices, np.int32)
    n = len(offsets) - 1
    best, med = np.empty(n, np.int32), np.empty(n, np.int32)
    lib().orbo_distinctive(_ptr(desc), _ptr(offsets), _ptr(indices), n, _ptr(best), _ptr(med))
    return best, med


def voc_transform(child_off, child_ids, node_desc, word_id, L, levels_up, feat):
    """OrbVocabulary::transform5 for every row of feat -> (word id, node id at level L - levels_up)."""
    child_off = np.ascontiguousarray(child_off, np.int32); child_ids = np.ascontiguousarray(child_ids, np.int32)
    node_desc = np.ascontiguousarray(node_desc, np.uint8); word_id = np.ascontiguousarray(word_id, np.int32)
    feat = np.ascontiguousarray(feat, np.uint8)
    n = len(feat)
    word, node = np.empty(n, np.int32), np.empty(n, np.int32)
    lib().orbo_voc_transform(_ptr(child_off), _ptr(child_ids), _ptr(node_desc), _ptr(word_id), L, levels_up, _ptr(feat), n, _ptr(word), _ptr(node))
    return word, node


def write_vocabulary_text(path, child_off, child_ids, node_desc, weight, k, L):
    """The DBoW2 text format OrbVocabulary::loadFromTextFile reads (orbvocabulary.cpp:39-118): a header "k L s w", then one
    line per non-root node in id order: "parent isLeaf d0 .. d31 weight".  No trailing newline: the loader's
    `while(!f.eof())` would otherwise parse an empty line into a node with an uninitialised parent."""
    child_off = np.asarray(child_off); child_ids = np.asarray(child_ids)
    n = len(child_off) - 1
    parent = np.zeros(n, np.int64)
    for v in range(n):
        ch = child_ids[child_off[v]:child_off[v + 1]]
        parent[ch] = v
        assert (np.diff(ch) > 0).all(), "children must be listed in id order"
    lines = [f"{k} {L} 0 0"]
    for v in range(1, n):
        leaf = int(child_off[v + 1] == child_off[v])
        lines.append(f"{parent[v]} {leaf} " + " ".join(str(int(b)) for b in node_desc[v]) + f" {float(weight[v])!r}")
    with open(path, "w") as f:
        f.write("\n".join(lines))


class RefVocabulary:
    """The reference's own OrbVocabulary (oracle/_ref/libvocref.so, built from /root/reference by `make -C oracle ref`)."""
    PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libvocref.so")

    def __init__(self, text_path, lib_name=None):
        # lib_name="libvocdropin.so": the same class with the body of transform4 replaced by the liborbx-backed one (GPU box)
        R = C.CDLL(self.PATH if lib_name is None else os.path.join(os.path.dirname(self.PATH), lib_name))
        R.vocref_load.restype = C.c_void_p
        R.vocref_load.argtypes = [C.c_char_p]
        R.vocref_free.argtypes = [C.c_void_p]
        R.vocref_size.argtypes = [C.c_void_p]
        R.vocref_transform_each.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        R.vocref_transform4.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        self.R = R
        self.h = R.vocref_load(text_path.encode())

    def size(self):
        return self.R.vocref_size(self.h)

    def transform_each(self, feat, levels_up):
        feat = np.ascontiguousarray(feat, np.uint8)
        n = len(feat)
        word, node = np.empty(n, np.int32), np.empty(n, np.int32)
        self.R.vocref_transform_each(self.h, _ptr(feat), n, levels_up, _ptr(word), _ptr(node))
        return word, node

    def transform4(self, feat, levels_up):
        feat = np.ascontiguousarray(feat, np.uint8)
        n = len(feat)
        ids, vals = np.zeros(n, np.uint32), np.zeros(n, np.float64)
        nodes, feats = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
        nb, nf = C.c_int(), C.c_int()
        rc = self.R.vocref_transform4(self.h, _ptr(feat), n, levels_up, _ptr(ids), _ptr(vals), n, C.byref(nb),
                                      _ptr(nodes), _ptr(feats), n, C.byref(nf))
        assert rc == 0
        return ids[:nb.value], vals[:nb.value], nodes[:nf.value], feats[:nf.value]

    def close(self):
        if self.h:
            self.R.vocref_free(self.h)
            self.h = None


def ref_distinctive(desc, offsets, indices, bad=None):
    """The reference's own OrbMapPoint::ComputeDistinctiveDescriptors (oracle/_ref/libmpref.so) -> (descriptor[n][32], has[n])."""
    R = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libmpref.so"))
    R.mpref_distinctive.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    desc = np.ascontiguousarray(desc, np.uint8)
    offsets = np.ascontiguousarray(offsets, np.int32); indices = np.ascontiguousarray(indices, np.int32)
    n = len(offsets) - 1
    out, has = np.zeros((n, 32), np.uint8), np.zeros(n, np.int32)
    b = None if bad is None else np.ascontiguousarray(bad, np.uint8)
    R.mpref_distinctive(_ptr(desc), len(desc), _ptr(offsets), _ptr(indices), None if b is None else _ptr(b), n, _ptr(out), _ptr(has))
    return out, has


def search_by_projection(keys, uright, occupied, desc, bounds, mp_desc, mp_x, mp_y, mp_level, mp_radius, nnratio=0.8, th_high=100,
                         mp_observed=None):
    """ORBmatcher::SearchByProjection(frame, map points, th) with the frame grid -> (mp_match, assigned, nmatches).
    mp_observed[i] != 0: map point i has observations, so once accepted it hides its key point from later map points."""
    keys = np.ascontiguousarray(keys); uright = np.ascontiguousarray(uright, np.float32)
    occ = None if occupied is None else np.ascontiguousarray(occupied, np.uint8)
    desc = np.ascontiguousarray(desc, np.uint8); mp_desc = np.ascontiguousarray(mp_desc, np.uint8)
    mp_x = np.ascontiguousarray(mp_x, np.float32); mp_y = np.ascontiguousarray(mp_y, np.float32)
    mp_level = np.ascontiguousarray(mp_level, np.int32); mp_radius = np.ascontiguousarray(mp_radius, np.float32)
    n, nmp = len(keys), len(mp_desc)
    match = np.empty(nmp, np.int32); assigned = np.empty(max(n, 1), np.int32)
    f = lib().orbo_search_by_projection
    f.restype = C.c_int
    f.argtypes = [C.c_void_p] * 4 + [C.c_int] + [C.c_float] * 4 + [C.c_void_p] * 6 + [C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p]
    obs = None if mp_observed is None else np.ascontiguousarray(mp_observed, np.uint8)
    nm = f(_ptr(keys), _ptr(uright), None if occ is None else _ptr(occ), _ptr(desc), n, *[float(v) for v in bounds],
           _ptr(mp_desc), _ptr(mp_x), _ptr(mp_y), _ptr(mp_level), _ptr(mp_radius), None if obs is None else _ptr(obs), nmp,
           float(nnratio), int(th_high), _ptr(match), _ptr(assigned))
    return match, assigned[:n], nm


def filter_keypoints(keys, desc, box):
    """OrbFrame::FilterKeyPoints -> (keys, descriptors) without the key points strictly inside box = (x0, x1, y0, y1)."""
    k = np.ascontiguousarray(keys).copy(); d = np.ascontiguousarray(desc, np.uint8).copy()
    b = np.ascontiguousarray(box, np.float32)
    f = lib().orbo_filter_keypoints
    f.restype = C.c_int; f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    n = f(_ptr(k), _ptr(d), len(k), _ptr(b))
    return k[:n], d[:n]


def assign_grid(keys, bounds):
    """OrbFrame::AssignFeaturesToGrid as CSR -> (cell_start[3073], cell_items)."""
    k = np.ascontiguousarray(keys)
    start = np.zeros(64 * 48 + 1, np.int32); items = np.zeros(max(len(k), 1), np.int32)
    f = lib().orbo_assign_grid
    f.restype = None; f.argtypes = [C.c_void_p, C.c_int] + [C.c_float] * 4 + [C.c_void_p, C.c_void_p]
    f(_ptr(k), len(k), *[float(v) for v in bounds], _ptr(start), _ptr(items))
    return start, items[:start[-1]].copy()


def area_distances(keys, desc, bounds, q_desc, q_x, q_y, q_r, q_min_level, q_max_level, cap=1 << 20):
    """OrbFrame::GetFeaturesInArea for every window + DescriptorDistance -> (offsets, indices, dist or None)."""
    keys = np.ascontiguousarray(keys); desc = np.ascontiguousarray(desc, np.uint8)
    qd = None if q_desc is None else np.ascontiguousarray(q_desc, np.uint8)
    q_x = np.ascontiguousarray(q_x, np.float32); q_y = np.ascontiguousarray(q_y, np.float32); q_r = np.ascontiguousarray(q_r, np.float32)
    l0 = np.ascontiguousarray(q_min_level, np.int32); l1 = np.ascontiguousarray(q_max_level, np.int32)
    nq = len(q_x)
    offsets = np.zeros(nq + 1, np.int32); indices = np.zeros(cap, np.int32); dist = np.zeros(cap, np.int32)
    f = lib().orbo_area_distances
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_float] * 4 + [C.c_void_p] * 6 + [C.c_int] + [C.c_void_p] * 3 + [C.c_int]
    total = f(_ptr(keys), _ptr(desc), len(keys), *[float(v) for v in bounds], None if qd is None else _ptr(qd), _ptr(q_x), _ptr(q_y),
              _ptr(q_r), _ptr(l0), _ptr(l1), nq, _ptr(offsets), _ptr(indices), _ptr(dist), cap)
    assert total <= cap
    return offsets, indices[:total].copy(), (None if qd is None else dist[:total].copy())


def ref_descriptor_distance(a, b):
    """The reference's own ORBmatcher::DescriptorDistance (orbmatcher.cpp:1662-1677, oracle/_ref/libframeref.so), row by row."""
    R = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libframeref.so"))
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    out = np.zeros(len(a), np.int32)
    R.frameref_descriptor_distance.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    R.frameref_descriptor_distance(_ptr(a), _ptr(b), len(a), _ptr(out))
    return out


def ref_search_by_projection(pairA, pairB, mbf, mb, th=3.0, nnratio=0.8, mp_step=1, dx=0.0, dy=0.0,
                             nfeatures=2000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, canonical=1, mp_dup=1, obs_mod=0):
    """The reference's own ORBmatcher::SearchByProjection(frame, map points, th) (src/orbmatcher.cpp:42-124 compiled
    unmodified into oracle/_ref/libframeref.so) on two reference OrbFrames; see oracle/cvshim/frame_glue.cpp.
    mp_dup map points per key point of A (copies of one another: they collide on the same key point of B); with obs_mod > 0
    map point k carries an observation (AddObservingKeyframe) unless k % obs_mod == 0, which makes the rule of :87-89 live
    inside the call.
    -> dict(mp_desc, mp_x, mp_y, mp_level, mp_radius, mp_observed, b_keys, b_desc, b_octave, b_uright, b_occupied, bounds,
    offsets, indices, assigned, nmatches)."""
    class Cfg(C.Structure):
        _fields_ = [("nfeatures", C.c_int), ("scale", C.c_float), ("nlevels", C.c_int), ("ini", C.c_int), ("min", C.c_int)]
    R = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libframeref.so"))
    R.frameref_search_by_projection.restype = C.c_int
    R.frameref_search_by_projection.argtypes = [C.POINTER(Cfg), C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.c_int] + [C.c_float] * 4 + \
        [C.c_int, C.c_float, C.c_float, C.c_int, C.c_int] + [C.POINTER(C.c_int)] + [C.c_void_p] * 3 + [C.POINTER(C.c_int)] + [C.c_void_p] * 11 + \
        [C.c_int, C.c_int, C.c_void_p]
    imgs = [np.ascontiguousarray(a, np.uint8) for a in (*pairA, *pairB)]
    h, w = imgs[0].shape
    cap, list_cap = (nfeatures + 512) * max(1, int(mp_dup)), 1 << 22
    mp_obs = np.zeros(cap, np.uint8)
    mp_desc = np.zeros((cap, 32), np.uint8); mp_x = np.zeros(cap, np.float32); mp_r = np.zeros(cap, np.float32)
    b_desc = np.zeros((cap, 32), np.uint8); b_oct = np.zeros(cap, np.int32); b_ur = np.zeros(cap, np.float32)
    b_occ = np.zeros(cap, np.int32); offsets = np.zeros(cap + 1, np.int32); indices = np.zeros(list_cap, np.int32)
    assigned = np.zeros(cap, np.int32)
    b_keys = np.zeros(cap, KP_DTYPE); mp_y = np.zeros(cap, np.float32); mp_level = np.zeros(cap, np.int32); bounds = np.zeros(4, np.float32)
    nmp, nb = C.c_int(), C.c_int()
    nm = R.frameref_search_by_projection(C.byref(Cfg(nfeatures, scale_factor, nlevels, ini_th, min_th)), int(canonical), *[_ptr(a) for a in imgs], w, h,
                                         float(mbf), float(mb), float(th), float(nnratio), int(mp_step), float(dx), float(dy), cap, list_cap,
                                         C.byref(nmp), _ptr(mp_desc), _ptr(mp_x), _ptr(mp_r), C.byref(nb), _ptr(b_desc), _ptr(b_oct),
                                         _ptr(b_ur), _ptr(b_occ), _ptr(offsets), _ptr(indices), _ptr(assigned), _ptr(b_keys), _ptr(mp_y), _ptr(mp_level), _ptr(bounds),
                                         int(mp_dup), int(obs_mod), _ptr(mp_obs))
    if nm < 0:
        raise RuntimeError("frameref_search_by_projection: output buffers too small")
    nmp, nb = nmp.value, nb.value
    return dict(mp_desc=mp_desc[:nmp].copy(), mp_x=mp_x[:nmp].copy(), mp_radius=mp_r[:nmp].copy(), b_desc=b_desc[:nb].copy(),
                b_octave=b_oct[:nb].copy(), b_uright=b_ur[:nb].copy(), b_occupied=b_occ[:nb].copy(), offsets=offsets[:nmp + 1].copy(),
                indices=indices[:offsets[nmp]].copy(), assigned=assigned[:nb].copy(), nmatches=nm,
                b_keys=b_keys[:nb].copy(), mp_y=mp_y[:nmp].copy(), mp_level=mp_level[:nmp].copy(), bounds=bounds,
                mp_observed=mp_obs[:nmp].copy())


def transform4(child_off, child_ids, node_desc, word_id, weight, L, levels_up, feat):
    """OrbVocabulary::transform4 (orbvocabulary.cpp:168-201) on top of the restated transform5: features with a positive
    word weight enter the bag of words (weights summed per word, then L1-normalised, orbbowvector.cpp:29-69) and the
    feature vector (node -> feature indices in feature order, orbfeaturevector.cpp:27-40), both ordered by key."""
    word, node = voc_transform(child_off, child_ids, node_desc, word_id, L, levels_up, feat)
    wt_of_word = np.asarray(weight, np.float64)[np.flatnonzero(np.asarray(word_id) >= 0)]
    bow, fv = {}, {}
    for i, (w, nd) in enumerate(zip(word.tolist(), node.tolist())):
        if wt_of_word[w] > 0:
            bow[w] = bow.get(w, 0.0) + wt_of_word[w]
            fv.setdefault(nd, []).append(i)
    norm = 0.0
    for w in sorted(bow):
        norm += abs(bow[w])
    ids = np.array(sorted(bow), np.uint32)
    vals = np.array([bow[w] / norm if norm > 0 else bow[w] for w in sorted(bow)], np.float64)
    nodes = np.array([nd for nd in sorted(fv) for _ in fv[nd]], np.uint32)
    feats = np.array([i for nd in sorted(fv) for i in fv[nd]], np.uint32)
    return ids, vals, nodes, feats


def stereo_matches_levels(lvL, lvR, sf, isf, kl, dl, kr, dr, mbf, mb):
    """OrbFrame::ComputeStereoMatches (orbframe.cpp:511-705) over explicit pyramid levels (lists of 2-D uint8 arrays)."""
    nlv = len(lvL)
    lvL = [np.ascontiguousarray(a) for a in lvL]; lvR = [np.ascontiguousarray(a) for a in lvR]
    pl = (C.c_void_p * nlv)(*[a.ctypes.data for a in lvL]); pr = (C.c_void_p * nlv)(*[a.ctypes.data for a in lvR])
    lw = (C.c_int * nlv)(*[a.shape[1] for a in lvL]); lh = (C.c_int * nlv)(*[a.shape[0] for a in lvL])
    ls = (C.c_size_t * nlv)(*[a.strides[0] for a in lvL])
    sf = (C.c_float * nlv)(*sf[:nlv]); isf = (C.c_float * nlv)(*isf[:nlv])
    kl = np.ascontiguousarray(kl); kr = np.ascontiguousarray(kr)
    dl = np.ascontiguousarray(dl, np.uint8); dr = np.ascontiguousarray(dr, np.uint8)
    u = np.empty(len(kl), np.float32); d = np.empty(len(kl), np.float32)
    n = lib().orbo_stereo_matches(_ptr(kl), _ptr(dl), len(kl), _ptr(kr), _ptr(dr), len(kr), pl, pr, lw, lh, ls, sf, isf,
                                  float(mbf), float(mb), _ptr(u), _ptr(d))
    return u, d, n


def stereo_matches(exL, exR, kl, dl, kr, dr, mbf, mb):
    """OrbFrame::ComputeStereoMatches over two oracle extractors that just processed the left / right image."""
    nlv = exL.nlevels
    return stereo_matches_levels([exL.level(l) for l in range(nlv)], [exR.level(l) for l in range(nlv)],
                                 exL.params.sf, exL.params.inv_sf, kl, dl, kr, dr, mbf, mb)


def ref_stereo_frame(left, right, mbf, mb, nfeatures=2000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, canonical=1, bbox=None, lib_name="libframeref.so"):
    """The reference's own OrbFrame stereo constructor + ComputeStereoMatches (oracle/_ref/libframeref.so: src/orbframe.cpp
    and src/orbextractor.cpp compiled unmodified) -> dict(kl, dl, kr, dr, levelsL, levelsR, uRight, depth, grid_*).
    lib_name="libdropinref.so": the same reference OrbFrame compiled against the drop-in OrbExtractor (GPU box only)."""
    class Cfg(C.Structure):
        _fields_ = [("nfeatures", C.c_int), ("scale", C.c_float), ("nlevels", C.c_int), ("ini", C.c_int), ("min", C.c_int)]
    R = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", lib_name))
    R.frameref_stereo.restype = C.c_int
    R.frameref_stereo.argtypes = [C.POINTER(Cfg), C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float] + \
        [C.c_void_p] * 4 + [C.c_int, C.POINTER(C.c_int)] + [C.c_void_p] * 9
    left = np.ascontiguousarray(left, np.uint8); right = np.ascontiguousarray(right, np.uint8)
    h, w = left.shape
    cap = nfeatures + 512
    kl = np.zeros(cap, KP_DTYPE); kr = np.zeros(cap, KP_DTYPE)
    dl = np.zeros((cap, 32), np.uint8); dr = np.zeros((cap, 32), np.uint8)
    lvL = [np.zeros(w * h, np.uint8) for _ in range(nlevels)]; lvR = [np.zeros(w * h, np.uint8) for _ in range(nlevels)]
    pl = (C.c_void_p * nlevels)(*[a.ctypes.data for a in lvL]); pr = (C.c_void_p * nlevels)(*[a.ctypes.data for a in lvR])
    lw = (C.c_int * nlevels)(); lh = (C.c_int * nlevels)()
    u = np.zeros(cap, np.float32); d = np.zeros(cap, np.float32)
    box = None if bbox is None else np.ascontiguousarray(bbox, np.float32)
    gstart = np.zeros(64 * 48 + 1, np.int32); gitems = np.zeros(cap, np.int32)
    nr = C.c_int()
    n = R.frameref_stereo(C.byref(Cfg(nfeatures, scale_factor, nlevels, ini_th, min_th)), int(canonical), _ptr(left), _ptr(right), w, h,
                          float(mbf), float(mb), _ptr(kl), _ptr(dl), _ptr(kr), _ptr(dr), cap, C.byref(nr), pl, pr, lw, lh, _ptr(u), _ptr(d),
                          None if box is None else _ptr(box), _ptr(gstart), _ptr(gitems))
    if n < 0:
        raise RuntimeError("frameref_stereo: more keypoints than the output buffers hold")
    return dict(kl=kl[:n].copy(), dl=dl[:n].copy(), kr=kr[:nr.value].copy(), dr=dr[:nr.value].copy(),
                levelsL=[lvL[l][:lw[l] * lh[l]].reshape(lh[l], lw[l]).copy() for l in range(nlevels)],
                levelsR=[lvR[l][:lw[l] * lh[l]].reshape(lh[l], lw[l]).copy() for l in range(nlevels)],
                uRight=u[:n].copy(), depth=d[:n].copy(), grid_start=gstart, grid_items=gitems[:gstart[-1]].copy())


class Extractor:
    """Mirror of OrbExtractor (orbextractor.hpp:90-109) over the C oracle."""

    def __init__(self, nfeatures=2000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, taps=None, tie_rule=0):
        t = (C.c_int * 7)(*taps) if taps is not None else None
        self._h = lib().orbo_create(nfeatures, scale_factor, nlevels, ini_th, min_th, t)
        if not self._h:
            raise ValueError("bad extractor parameters")
        lib().orbo_set_tie_rule(self._h, tie_rule)
        self.params = Params.from_buffer_copy(lib().orbo_get_params(self._h).contents)  # own copy: outlives the handle
        self.nfeatures, self.nlevels = nfeatures, nlevels

    def __del__(self):
        if getattr(self, "_h", None) and lib is not None:      # module globals are already gone at interpreter shutdown
            try:
                lib().orbo_destroy(self._h)
            except Exception:
                pass
            self._h = None

    def cap(self, img_shape):
        return int(sum(self.params.quota[:self.nlevels])) + 8 * self.nlevels + 64

    def extract(self, img):
        img = _c_u8(img)
        cap = self.cap(img.shape)
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = lib().orbo_extract(self._h, _ptr(img), img.shape[1], img.shape[0], img.strides[0], _ptr(kps), _ptr(desc), cap)
        if n < 0:
            raise RuntimeError(f"orbo_extract failed: {n}")
        return kps[:n].copy(), desc[:n].copy()

    def extract_batch_mt(self, imgs, nthreads):
        imgs = [np.ascontiguousarray(i) for i in imgs]
        h, w = imgs[0].shape
        cap = self.cap(imgs[0].shape)
        n = len(imgs)
        kps = np.zeros((n, cap), KP_DTYPE)
        desc = np.zeros((n, cap, 32), np.uint8)
        counts = np.zeros(n, np.int32)
        ptrs = (C.c_void_p * n)(*[i.ctypes.data for i in imgs])
        lib().orbo_extract_batch_mt(self._h, ptrs, n, w, h, w, _ptr(kps), _ptr(desc), cap, _ptr(counts), nthreads)
        return kps, desc, counts

    def level(self, l):
        w, h, s = C.c_int(), C.c_int(), C.c_size_t()
        p = lib().orbo_level(self._h, l, C.byref(w), C.byref(h), C.byref(s))
        if not p:
            return None
        buf = (C.c_uint8 * (s.value * h.value)).from_address(p)
        a = np.frombuffer(buf, np.uint8).reshape(h.value, s.value)[:, :w.value]
        return a.copy()

    def blurred(self, l):
        w, h, s = C.c_int(), C.c_int(), C.c_size_t()
        p = lib().orbo_blurred(self._h, l, C.byref(w), C.byref(h), C.byref(s))
        if not p:
            return None
        buf = (C.c_uint8 * (s.value * h.value)).from_address(p)
        return np.frombuffer(buf, np.uint8).reshape(h.value, s.value)[:, :w.value].copy()

    def candidates(self, l):
        xs, ys, sc = _ip(), _ip(), _ip()
        n = lib().orbo_candidates(self._h, l, C.byref(xs), C.byref(ys), C.byref(sc))
        if n <= 0:
            z = np.zeros(0, np.int32)
            return z, z, z
        f = lambda p: np.ctypeslib.as_array(p, shape=(n,)).copy()
        return f(xs), f(ys), f(sc)
