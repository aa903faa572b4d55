"""ctypes binding of liborbx.so (include/orbx.h) for tests, bench.py and Python users.

The classes mirror the reference's interface for this path:
  Extractor  <->  OrbExtractor   (include/orbextractor.hpp:90-109 of the reference)
  Matcher    <->  ORBmatcher::DescriptorDistance + best/second-best loop (orbmatcher.hpp:48)
There is no fallback: if liborbx.so is missing or no sm_100 GPU is usable, construction raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ORBX_LIB") or os.path.join(_HERE, "liborbx.so")     # ORBX_LIB: another build of the same library (debugging)
MAX_LEVELS = 16

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])

ERR_NAMES = {0: "OK", -1: "ERR_ARG", -2: "ERR_SHAPE", -3: "ERR_CAPACITY", -4: "ERR_CUDA", -5: "ERR_NOMEM"}


class Config(C.Structure):
    _fields_ = [("nfeatures", C.c_int), ("scale_factor", C.c_float), ("nlevels", C.c_int),
                ("ini_th_fast", C.c_int), ("min_th_fast", C.c_int), ("max_width", C.c_int),
                ("max_height", C.c_int), ("max_batch", C.c_int), ("device", C.c_int),
                ("blur_taps", C.c_int * 7), ("tie_rule", C.c_int)]


class OrbxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


_lib = None
EXPORTS = ["orbx_version", "orbx_create", "orbx_destroy", "orbx_last_error", "orbx_extract", "orbx_extract_batch",
           "orbx_extract_batch_async", "orbx_wait", "orbx_multi_create", "orbx_multi_destroy", "orbx_multi_last_error", "orbx_multi_devices",
           "orbx_multi_max_keypoints", "orbx_multi_extract_batch", "orbx_multi_extract_batch_async", "orbx_multi_wait", "orbx_multi_handle",
           "orbx_multi_frame_range", "orbm_multi_create", "orbm_multi_destroy", "orbm_multi_last_error", "orbm_multi_devices", "orbm_multi_set_train",
           "orbm_multi_knn2", "orbm_multi_matcher",
           "orbx_extract_batch_device", "orbx_set_device_split", "orbx_device_results", "orbx_pipe_create", "orbx_pipe_destroy", "orbx_pipe_last_error", "orbx_pipe_depth",
           "orbx_pipe_submit", "orbx_pipe_join", "orbx_pipe_handle", "orbx_fetch_results", "orbx_filter_keypoints", "orbx_stereo_match", "orbx_stereo_match_batch", "orbx_max_keypoints", "orbx_host_alloc", "orbx_host_free", "orbx_last_launches", "orbx_get_level",
           "orbx_scale_tables", "orbx_profile_stages", "orbx_debug_blurred", "orbx_debug_enable_candidates", "orbx_debug_candidates",
           "orbm_create", "orbm_destroy", "orbm_last_error", "orbm_knn2", "orbm_set_train", "orbm_knn2_resident",
           "orbm_knn2_device", "orbm_window_create", "orbm_window_attach_ipc", "orbm_window_attach_peer", "orbm_knn2_sharded", "orbm_window_status", "orbm_window_fetch", "orbm_window_records", "orbm_knn2_csr", "orbm_knn2_csr_device", "orbm_distance_csr", "orbm_search_by_projection", "orbm_area_distances", "orbm_assign_grid", "orbm_distinctive", "orbm_distance_pairs", "orbm_measure_popc",
           "orbv_create", "orbv_destroy", "orbv_last_error", "orbv_transform", "orbv_transform_device"]


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, ip, fp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float)
    L.orbx_version.restype = C.c_char_p
    L.orbx_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.orbx_destroy.argtypes = [vp]
    L.orbx_last_error.restype = C.c_char_p
    L.orbx_last_error.argtypes = [vp]
    L.orbx_extract.argtypes = [vp, vp, C.c_int, C.c_int, C.c_size_t, vp, C.c_int, vp, ip]
    L.orbx_extract_batch.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_size_t, vp, C.c_int, vp, vp]
    L.orbx_extract_batch_async.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_size_t, vp, C.c_int, vp, vp, ip]
    L.orbx_wait.argtypes = [vp, C.c_int]
    L.orbx_multi_create.argtypes = [C.POINTER(Config), ip, C.c_int, C.POINTER(vp)]
    L.orbx_multi_destroy.argtypes = [vp]
    L.orbx_multi_last_error.restype = C.c_char_p
    L.orbx_multi_last_error.argtypes = [vp]
    L.orbx_multi_devices.argtypes = [vp]
    L.orbx_multi_max_keypoints.argtypes = [vp]
    L.orbx_multi_extract_batch.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_size_t, vp, C.c_int, vp, vp]
    L.orbx_multi_extract_batch_async.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_size_t, vp, C.c_int, vp, vp, ip]
    L.orbx_multi_wait.argtypes = [vp, C.c_int]
    L.orbx_multi_handle.restype = vp
    L.orbx_multi_handle.argtypes = [vp, C.c_int]
    L.orbx_multi_frame_range.argtypes = [vp, C.c_int, C.c_int, ip, ip]
    L.orbm_multi_create.argtypes = [ip, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.orbm_multi_destroy.argtypes = [vp]
    L.orbm_multi_last_error.restype = C.c_char_p
    L.orbm_multi_last_error.argtypes = [vp]
    L.orbm_multi_devices.argtypes = [vp]
    L.orbm_multi_set_train.argtypes = [vp, vp, C.c_int]
    L.orbm_multi_knn2.argtypes = [vp, vp, C.c_int, vp, vp, vp]
    L.orbm_multi_matcher.restype = vp
    L.orbm_multi_matcher.argtypes = [vp, C.c_int]
    L.orbx_last_launches.argtypes = [vp]
    L.orbx_extract_batch_device.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_int, vp]
    L.orbx_device_results.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), ip]
    L.orbx_fetch_results.argtypes = [vp, vp, vp, C.c_int, vp, vp]
    L.orbx_filter_keypoints.argtypes = [vp, C.c_int, C.c_int, fp]
    L.orbx_stereo_match.argtypes = [vp, C.c_int, vp, C.c_int, C.c_float, C.c_float, vp, vp, C.c_int, ip, ip]
    L.orbx_stereo_match_batch.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, vp, vp, C.c_int, ip, ip]
    L.orbx_max_keypoints.argtypes = [vp]
    L.orbx_set_device_split.argtypes = [vp, C.c_int]
    L.orbx_pipe_create.argtypes = [C.POINTER(Config), C.c_int, C.POINTER(vp)]
    L.orbx_pipe_destroy.argtypes = [vp]; L.orbx_pipe_destroy.restype = None
    L.orbx_pipe_last_error.argtypes = [vp]; L.orbx_pipe_last_error.restype = C.c_char_p
    L.orbx_pipe_depth.argtypes = [vp]
    L.orbx_pipe_submit.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_int, vp, ip]
    L.orbx_pipe_join.argtypes = [vp, C.c_int, vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), ip]
    L.orbx_pipe_handle.argtypes = [vp, C.c_int]; L.orbx_pipe_handle.restype = vp
    L.orbx_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.orbx_host_free.argtypes = [vp]; L.orbx_host_free.restype = None
    L.orbx_get_level.argtypes = [vp, C.c_int, C.c_int, C.POINTER(vp), ip, ip, C.POINTER(C.c_size_t)]
    L.orbx_scale_tables.argtypes = [vp, fp, fp, fp, fp, ip]
    L.orbx_profile_stages.argtypes = [vp, C.c_int, fp, C.c_int]
    L.orbx_debug_blurred.argtypes = [vp, C.c_int, C.c_int, vp, C.c_size_t, ip, ip]
    L.orbx_debug_enable_candidates.argtypes = [vp, C.c_int]
    L.orbx_debug_candidates.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, C.c_int]
    L.orbm_create.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.orbm_destroy.argtypes = [vp]
    L.orbm_last_error.restype = C.c_char_p
    L.orbm_last_error.argtypes = [vp]
    L.orbm_knn2.argtypes = [vp, vp, C.c_int, vp, C.c_int, vp, vp, vp]
    L.orbm_set_train.argtypes = [vp, vp, C.c_int]
    L.orbm_knn2_resident.argtypes = [vp, vp, C.c_int, vp, vp, vp]
    L.orbm_knn2_device.argtypes = [vp, vp, C.c_int, vp, C.c_int, vp, vp]
    L.orbm_window_create.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp]
    L.orbm_window_attach_ipc.argtypes = [vp, C.c_int, vp]
    L.orbm_window_attach_peer.argtypes = [vp, C.c_int, vp]
    L.orbm_knn2_sharded.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_int, vp]
    L.orbm_window_status.argtypes = [vp, vp]
    L.orbm_window_records.argtypes = [vp, C.POINTER(vp)]
    L.orbm_window_fetch.argtypes = [vp, vp, vp, C.c_int]
    L.orbm_knn2_csr.argtypes = [vp, vp, C.c_int, vp, C.c_int, vp, vp, vp, vp, vp, vp]
    L.orbm_knn2_csr_device.argtypes = [vp, vp, C.c_int, vp, vp, vp, vp, vp]
    L.orbm_measure_popc.argtypes = [vp, C.POINTER(C.c_double)]
    L.orbm_distance_pairs.argtypes = [vp, vp, vp, C.c_int, vp]
    L.orbm_distance_csr.argtypes = [vp, vp, C.c_int, vp, C.c_int, vp, vp, vp]
    L.orbm_assign_grid.argtypes = [vp, vp, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, vp, vp]
    L.orbm_area_distances.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, C.c_int, vp, vp, vp, C.c_int, C.POINTER(C.c_int32)]
    L.orbm_search_by_projection.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, C.c_int, C.c_float, C.c_int, vp, vp, C.POINTER(C.c_int32)]
    L.orbm_distinctive.argtypes = [vp, vp, C.c_int, vp, vp, C.c_int, vp, vp]
    L.orbv_create.argtypes = [C.c_int, C.c_int, vp, vp, vp, vp, vp, C.c_int, C.POINTER(vp)]
    L.orbv_destroy.argtypes = [vp]
    L.orbv_last_error.restype = C.c_char_p
    L.orbv_last_error.argtypes = [vp]
    L.orbv_transform.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp, vp]
    L.orbv_transform_device.argtypes = [vp, vp, C.c_size_t, C.c_int, C.c_int, vp, vp]
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Extractor:
    """OrbExtractor(nFeatures, scaleFactor, nLevels, iniThFAST, minThFAST) on a B200."""

    def __init__(self, nfeatures=2000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7,
                 max_width=1241, max_height=376, max_batch=1, device=0, taps=None, tie_rule=0):
        cfg = Config(nfeatures, scale_factor, nlevels, ini_th, min_th, max_width, max_height, max_batch, device,
                     (C.c_int * 7)(*(taps if taps is not None else [0] * 7)), tie_rule)
        self._h = C.c_void_p()
        rc = lib().orbx_create(C.byref(cfg), C.byref(self._h))
        if rc != 0:
            msg = lib().orbx_last_error(self._h).decode() if self._h else "invalid configuration"
            if self._h:
                lib().orbx_destroy(self._h)
                self._h = None
            raise OrbxError(rc, msg)
        self.nlevels, self.nfeatures, self.max_batch = nlevels, nfeatures, max_batch

    def close(self):
        if getattr(self, "_h", None):
            if getattr(self, "_owned", True):
                lib().orbx_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, rc):
        if rc < 0:
            raise OrbxError(rc, lib().orbx_last_error(self._h).decode())
        return rc

    @property
    def max_keypoints(self):
        return lib().orbx_max_keypoints(self._h)

    # getters of orbextractor.cpp:557-579
    def tables(self):
        n = self.nlevels
        arrs = [np.zeros(n, np.float32) for _ in range(4)] + [np.zeros(n, np.int32)]
        fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
        self._check(lib().orbx_scale_tables(self._h, *[a.ctypes.data_as(fp) for a in arrs[:4]], arrs[4].ctypes.data_as(ip)))
        return dict(zip(["scale", "inv_scale", "sigma2", "inv_sigma2", "quota"], arrs))

    def extract(self, img):
        """ExtractFeatures: 2-D uint8 array -> (keypoints[KP_DTYPE], descriptors[n,32])."""
        kps, desc, counts = self.extract_batch([img])
        return kps[0][:counts[0]].copy(), desc[0][:counts[0]].copy()

    def extract_batch(self, imgs, out=None):
        imgs = [i if (i.dtype == np.uint8 and i.ndim == 2 and i.strides[1] == 1) else np.ascontiguousarray(i, np.uint8) for i in imgs]
        h, w = imgs[0].shape
        pitch = imgs[0].strides[0]
        for i in imgs:
            if i.shape != (h, w) or i.strides[0] != pitch:
                raise ValueError("all frames of a batch must share shape and pitch")
        n = len(imgs)
        cap = self.max_keypoints
        if out is None:
            out = (np.zeros((n, cap), KP_DTYPE), np.zeros((n, cap, 32), np.uint8), np.zeros(n, np.int32))
        kps, desc, counts = out
        ptrs = (C.c_void_p * n)(*[i.ctypes.data for i in imgs])
        self._check(lib().orbx_extract_batch(self._h, ptrs, n, w, h, pitch, _ptr(kps), kps.shape[1], _ptr(desc), _ptr(counts)))
        return kps, desc, counts

    @staticmethod
    def frame_pointers(imgs):
        """The `const uint8_t *const *` argument of orbx_extract_batch for a list of frames, built once and reusable
        (what a C++ caller holds anyway); keeps the frames alive."""
        ptrs = (C.c_void_p * len(imgs))(*[i.ctypes.data for i in imgs])
        ptrs._frames = imgs
        return ptrs

    def extract_batch_ptrs(self, ptrs, n, width, height, pitch, out):
        """orbx_extract_batch with a prepared pointer array and prepared output arrays: nothing but the C call."""
        kps, desc, counts = out
        self._check(lib().orbx_extract_batch(self._h, ptrs, n, width, height, pitch, kps.ctypes.data, kps.shape[1],
                                             desc.ctypes.data, counts.ctypes.data))
        return out

    def extract_batch_async(self, ptrs, n, width, height, pitch, out):
        """orbx_extract_batch_async: enqueue the call and return its ticket; `out` and the frames stay untouched until wait()."""
        kps, desc, counts = out
        t = C.c_int()
        self._check(lib().orbx_extract_batch_async(self._h, ptrs, n, width, height, pitch, kps.ctypes.data, kps.shape[1],
                                                   desc.ctypes.data, counts.ctypes.data, C.byref(t)))
        return t.value

    def wait(self, ticket):
        """orbx_wait: block until the call behind `ticket` has delivered its results."""
        self._check(lib().orbx_wait(self._h, int(ticket)))

    def extract_batch_device(self, dptr, frame_stride, pitch, batch, width, height, stream=None):
        """Frames already in HBM (raw device pointer); results stay in HBM (see device_results)."""
        self._check(lib().orbx_extract_batch_device(self._h, C.c_void_p(dptr), frame_stride, pitch, batch, width, height,
                                                    C.c_void_p(stream) if stream else None))

    def device_results(self):
        k, d, c, s = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int()
        self._check(lib().orbx_device_results(self._h, C.byref(k), C.byref(d), C.byref(c), C.byref(s)))
        return k.value, d.value, c.value, s.value

    def fetch_results(self, batch, stream=None):
        cap = self.max_keypoints
        kps = np.zeros((batch, cap), KP_DTYPE); desc = np.zeros((batch, cap, 32), np.uint8); counts = np.zeros(batch, np.int32)
        self._check(lib().orbx_fetch_results(self._h, C.c_void_p(stream) if stream else None, _ptr(kps), cap, _ptr(desc), _ptr(counts)))
        return kps, desc, counts

    def filter_keypoints(self, box, frame0=0, n_frames=1):
        """OrbFrame::FilterKeyPoints on the device-resident results of the last extraction: key points strictly inside
        box = (x0, x1, y0, y1) are removed in HBM (no-op unless box[1] > 2, as in the reference)."""
        b = (C.c_float * 4)(*[float(v) for v in box])
        self._check(lib().orbx_filter_keypoints(self._h, frame0, n_frames, b))

    def last_launches(self):
        """Kernel launches the last extract call enqueued (all chunks / halves), counted inside the library."""
        return int(lib().orbx_last_launches(self._h))

    STAGES = ("resize", "fast", "octree", "blur", "describe")

    def profile_stages(self, reps=5):
        ms = np.zeros(5, np.float32)
        self._check(lib().orbx_profile_stages(self._h, reps, ms.ctypes.data_as(C.POINTER(C.c_float)), 5))
        return dict(zip(self.STAGES, ms.tolist()))

    def level(self, frame, level):
        """m_vImagePyramid[level] of `frame` of the last call, as a host array (copy)."""
        p, w, h, pitch = C.c_void_p(), C.c_int(), C.c_int(), C.c_size_t()
        self._check(lib().orbx_get_level(self._h, frame, level, C.byref(p), C.byref(w), C.byref(h), C.byref(pitch)))
        buf = (C.c_uint8 * (pitch.value * h.value)).from_address(p.value)
        return np.frombuffer(buf, np.uint8).reshape(h.value, pitch.value)[:, :w.value].copy()

    def blurred(self, frame, level, shape):
        out = np.zeros(shape, np.uint8)
        w, h = C.c_int(), C.c_int()
        self._check(lib().orbx_debug_blurred(self._h, frame, level, _ptr(out), out.size, C.byref(w), C.byref(h)))
        assert (h.value, w.value) == tuple(shape)
        return out

    def enable_candidates(self, on=True):
        self._check(lib().orbx_debug_enable_candidates(self._h, int(on)))

    def candidates(self, frame, level, cap=1 << 22):
        xs, ys, sc = (np.zeros(cap, np.int32) for _ in range(3))
        n = self._check(lib().orbx_debug_candidates(self._h, frame, level, _ptr(xs), _ptr(ys), _ptr(sc), cap))
        return xs[:n].copy(), ys[:n].copy(), sc[:n].copy()


class Pipe:
    """orbx_pipe: `depth` device-resident extractions in flight on one GPU (submit / join by ticket)."""

    def __init__(self, depth=3, nfeatures=2000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7,
                 max_width=1241, max_height=376, max_batch=1, device=0, taps=None, tie_rule=0):
        cfg = Config(nfeatures, scale_factor, nlevels, ini_th, min_th, max_width, max_height, max_batch, device,
                     (C.c_int * 7)(*(taps if taps is not None else [0] * 7)), tie_rule)
        self._h = C.c_void_p()
        rc = lib().orbx_pipe_create(C.byref(cfg), depth, C.byref(self._h))
        if rc != 0:
            msg = lib().orbx_pipe_last_error(self._h).decode() if self._h else "invalid configuration"
            if self._h:
                lib().orbx_pipe_destroy(self._h)
                self._h = None
            raise OrbxError(rc, msg)
        self.depth = depth
        self._cfg = (nlevels, nfeatures, max_batch)

    def close(self):
        if getattr(self, "_h", None):
            lib().orbx_pipe_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, rc):
        if rc < 0:
            raise OrbxError(rc, lib().orbx_pipe_last_error(self._h).decode())
        return rc

    def submit(self, dptr, frame_stride, pitch, batch, width, height, stream=None):
        t = C.c_int()
        self._check(lib().orbx_pipe_submit(self._h, C.c_void_p(dptr), frame_stride, pitch, batch, width, height,
                                           C.c_void_p(stream) if stream else None, C.byref(t)))
        return t.value

    def join(self, ticket, stream=None):
        """-> (d_kps, d_desc, d_counts, kp_stride) of that submission; `stream` waits for it."""
        k, d, c, s = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int()
        self._check(lib().orbx_pipe_join(self._h, ticket, C.c_void_p(stream) if stream else None, C.byref(k), C.byref(d), C.byref(c), C.byref(s)))
        return k.value, d.value, c.value, s.value

    def extractor(self, ticket):
        """A non-owning Extractor view of the slot that holds `ticket` (fetch_results, levels, stereo matching of that batch)."""
        h = lib().orbx_pipe_handle(self._h, ticket)
        if not h:
            raise OrbxError(-1, "the ticket's slot has been reused")
        ex = Extractor.__new__(Extractor)
        ex._h, ex._owned = C.c_void_p(h), False
        ex.nlevels, ex.nfeatures, ex.max_batch = self._cfg
        return ex


def stereo_match(left, frame_left, right, frame_right, mbf, mb=0.0):
    """OrbFrame::ComputeStereoMatches on the device-resident results of two extractions -> (uRight, depth, n_matches)."""
    cap = left.max_keypoints
    u = np.zeros(cap, np.float32); d = np.zeros(cap, np.float32)
    nl, nm = C.c_int(), C.c_int()
    left._check(lib().orbx_stereo_match(left._h, frame_left, right._h, frame_right, mbf, mb, _ptr(u), _ptr(d), cap,
                                        C.byref(nl), C.byref(nm)))
    return u[:nl.value].copy(), d[:nl.value].copy(), nm.value


def stereo_match_batch(ex, n_pairs, frame_left0=0, frame_right0=1, frame_step=2, mbf=386.1448, mb=0.0, out=None):
    """ComputeStereoMatches for n_pairs pairs of the last batch of `ex` in one launch -> (uRight[n][cap], depth[n][cap], n_left, n_matches)."""
    cap = ex.max_keypoints
    if out is None:
        out = (np.zeros((n_pairs, cap), np.float32), np.zeros((n_pairs, cap), np.float32))
    u, d = out
    nl, nm = np.zeros(n_pairs, np.int32), np.zeros(n_pairs, np.int32)
    ex._check(lib().orbx_stereo_match_batch(ex._h, n_pairs, frame_left0, frame_right0, frame_step, mbf, mb, _ptr(u), _ptr(d), cap,
                                            nl.ctypes.data_as(C.POINTER(C.c_int)), nm.ctypes.data_as(C.POINTER(C.c_int))))
    return u, d, nl, nm


class FrameView(C.Structure):
    """orbm_frame_view (include/orbx.h)."""
    _fields_ = [("keys", C.c_void_p), ("u_right", C.c_void_p), ("occupied", C.c_void_p), ("desc", C.c_void_p), ("n", C.c_int32),
                ("min_x", C.c_float), ("min_y", C.c_float), ("max_x", C.c_float), ("max_y", C.c_float)]


class Matcher:
    """Brute-force Hamming kNN-2 with the reference's best / second-best semantics."""

    def __init__(self, max_queries=2000, max_train=100000, device=0):
        self._h = C.c_void_p()
        rc = lib().orbm_create(device, max_queries, max_train, C.byref(self._h))
        if rc != 0:
            msg = lib().orbm_last_error(self._h).decode() if self._h else "invalid configuration"
            if self._h:
                lib().orbm_destroy(self._h)
                self._h = None
            raise OrbxError(rc, msg)

    def close(self):
        if getattr(self, "_h", None):
            lib().orbm_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, rc):
        if rc < 0:
            raise OrbxError(rc, lib().orbm_last_error(self._h).decode())
        return rc

    def knn2(self, q, t):
        q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
        nq = len(q)
        idx, d1, d2 = (np.zeros(nq, np.int32) for _ in range(3))
        self._check(lib().orbm_knn2(self._h, _ptr(q), nq, _ptr(t), len(t), _ptr(idx), _ptr(d1), _ptr(d2)))
        return idx, d1, d2

    def set_train(self, t):
        t = np.ascontiguousarray(t, np.uint8)
        self._check(lib().orbm_set_train(self._h, _ptr(t), len(t)))

    def knn2_resident(self, q):
        q = np.ascontiguousarray(q, np.uint8)
        nq = len(q)
        idx, d1, d2 = (np.zeros(nq, np.int32) for _ in range(3))
        self._check(lib().orbm_knn2_resident(self._h, _ptr(q), nq, _ptr(idx), _ptr(d1), _ptr(d2)))
        return idx, d1, d2

    def knn2_device(self, dq, nq, dt, nt, dout, stream=None):
        self._check(lib().orbm_knn2_device(self._h, C.c_void_p(dq), nq, C.c_void_p(dt), nt, C.c_void_p(dout),
                                           C.c_void_p(stream) if stream else None))

    # ---- query-sharded kNN-2 with the gather fused into the kernel
    def window_create(self, nq_total, n_ranks, rank):
        """-> the 64-byte CUDA IPC handle of this rank's result window (bytes)"""
        buf = C.create_string_buffer(64)
        self._check(lib().orbm_window_create(self._h, nq_total, n_ranks, rank, buf))
        self._win = (nq_total, n_ranks, rank)
        return buf.raw

    def window_attach_ipc(self, peer_rank, handle):
        self._check(lib().orbm_window_attach_ipc(self._h, peer_rank, C.create_string_buffer(handle, 64)))

    def window_attach_peer(self, peer_rank, peer):
        self._check(lib().orbm_window_attach_peer(self._h, peer_rank, peer._h))

    def knn2_sharded(self, dq, nq_local, q_offset, dt, nt, stream=None):
        self._check(lib().orbm_knn2_sharded(self._h, C.c_void_p(dq), nq_local, q_offset, C.c_void_p(dt), nt, C.c_void_p(stream or 0)))

    def window_status(self, stream=None):
        self._check(lib().orbm_window_status(self._h, C.c_void_p(stream or 0)))

    def window_fetch(self, nq, stream=None):
        """the records of all ranks after the last knn2_sharded call -> int32 [nq, 4] = {idx, d1, d2, pad}"""
        rec = np.zeros((nq, 4), np.int32)
        self._check(lib().orbm_window_fetch(self._h, C.c_void_p(stream or 0), _ptr(rec), nq))
        return rec

    def window_records_ptr(self):
        p = C.c_void_p()
        self._check(lib().orbm_window_records(self._h, C.byref(p)))
        return p.value

    def knn2_csr(self, q, t, offsets, indices):
        """Per-query candidate lists (CSR): (idx1, d1, idx2, d2) as SearchByProjection's inner loop leaves them."""
        q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.int32); indices = np.ascontiguousarray(indices, np.int32)
        nq = len(q)
        i1, d1, i2, d2 = (np.zeros(nq, np.int32) for _ in range(4))
        self._check(lib().orbm_knn2_csr(self._h, _ptr(q), nq, _ptr(t), len(t), _ptr(offsets), _ptr(indices),
                                        _ptr(i1), _ptr(d1), _ptr(i2), _ptr(d2)))
        return i1, d1, i2, d2

    def distance_csr(self, q, t, offsets, indices):
        """DescriptorDistance of every (query, candidate) entry of the CSR lists, in list order."""
        q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.int32); indices = np.ascontiguousarray(indices, np.int32)
        dist = np.zeros(len(indices), np.int32)
        self._check(lib().orbm_distance_csr(self._h, _ptr(q), len(q), _ptr(t), len(t), _ptr(offsets), _ptr(indices), _ptr(dist)))
        return dist

    def search_by_projection(self, keys, uright, occupied, desc, bounds, mp_desc, mp_x, mp_y, mp_level, mp_radius,
                             nnratio=0.8, th_high=100, mp_observed=None):
        """ORBmatcher::SearchByProjection(frame, map points, th) with the frame grid (AssignFeaturesToGrid,
        GetFeaturesInArea) on the device -> (mp_match[n_mp], assigned[n_keypoints], nmatches).  keys: KP_DTYPE records
        (m_undistortedKeys), bounds = (m_minX, m_minY, m_maxX, m_maxY), mp_radius = r * scaleFactor[level], mp_observed[i] != 0
        where map point i has observations (it then hides the key point it is stored on from the map points after it)."""
        keys = np.ascontiguousarray(keys, KP_DTYPE); uright = np.ascontiguousarray(uright, np.float32)
        occ = None if occupied is None else np.ascontiguousarray(occupied, np.uint8)
        desc = np.ascontiguousarray(desc, np.uint8); mp_desc = np.ascontiguousarray(mp_desc, np.uint8)
        mp_x = np.ascontiguousarray(mp_x, np.float32); mp_y = np.ascontiguousarray(mp_y, np.float32)
        mp_level = np.ascontiguousarray(mp_level, np.int32); mp_radius = np.ascontiguousarray(mp_radius, np.float32)
        n, nmp = len(keys), len(mp_desc)
        view = FrameView(keys.ctypes.data if n else None, uright.ctypes.data if n else None, None if occ is None or not n else occ.ctypes.data,
                         desc.ctypes.data if n else None, n, *[float(v) for v in bounds])
        match = np.zeros(nmp, np.int32); assigned = np.zeros(max(n, 1), np.int32); nm = C.c_int32()
        obs = None if mp_observed is None else np.ascontiguousarray(mp_observed, np.uint8)
        self._check(lib().orbm_search_by_projection(self._h, C.byref(view), _ptr(mp_desc), _ptr(mp_x), _ptr(mp_y), _ptr(mp_level),
                                                    _ptr(mp_radius), None if obs is None else _ptr(obs), nmp, C.c_float(nnratio),
                                                    int(th_high), _ptr(match), _ptr(assigned), C.byref(nm)))
        return match, assigned[:n], nm.value

    def assign_grid(self, keys, bounds):
        """OrbFrame::AssignFeaturesToGrid as CSR -> (cell_start[64 * 48 + 1], cell_items); cell = ix * 48 + iy."""
        keys = np.ascontiguousarray(keys, KP_DTYPE)
        start = np.zeros(64 * 48 + 1, np.int32); items = np.zeros(max(len(keys), 1), np.int32)
        self._check(lib().orbm_assign_grid(self._h, _ptr(keys) if len(keys) else None, len(keys), *[C.c_float(float(v)) for v in bounds],
                                           _ptr(start), _ptr(items)))
        return start, items[:start[-1]].copy()

    def area_distances(self, keys, desc, bounds, q_desc, q_x, q_y, q_r, q_min_level, q_max_level, cap=None):
        """OrbFrame::GetFeaturesInArea for every window of one frame + DescriptorDistance of each feature found ->
        (offsets[nq + 1], indices, dist or None), the lists in the reference's order."""
        keys = np.ascontiguousarray(keys, KP_DTYPE); desc = np.ascontiguousarray(desc, np.uint8)
        qd = None if q_desc is None else np.ascontiguousarray(q_desc, np.uint8)
        q_x = np.ascontiguousarray(q_x, np.float32); q_y = np.ascontiguousarray(q_y, np.float32); q_r = np.ascontiguousarray(q_r, np.float32)
        l0 = np.ascontiguousarray(q_min_level, np.int32); l1 = np.ascontiguousarray(q_max_level, np.int32)
        n, nq = len(keys), len(q_x)
        view = FrameView(keys.ctypes.data if n else None, None, None, desc.ctypes.data if n else None, n, *[float(v) for v in bounds])
        cap = int(cap) if cap is not None else max(1024, 16 * nq)
        while True:
            offsets = np.zeros(nq + 1, np.int32); indices = np.zeros(max(cap, 1), np.int32); dist = np.zeros(max(cap, 1), np.int32)
            total = C.c_int32()
            rc = lib().orbm_area_distances(self._h, C.byref(view), None if qd is None else _ptr(qd), _ptr(q_x), _ptr(q_y), _ptr(q_r),
                                           _ptr(l0), _ptr(l1), nq, _ptr(offsets), _ptr(indices), _ptr(dist), cap, C.byref(total))
            if rc == -3 and total.value > cap:               # ORBX_ERR_CAPACITY: the library says how many entries there are
                cap = total.value
                continue
            self._check(rc)
            return offsets, indices[:total.value].copy(), (None if qd is None else dist[:total.value].copy())

    def measure_popc(self):
        """POPC lane-operations per clock per SM measured on this GPU."""
        r = C.c_double()
        self._check(lib().orbm_measure_popc(self._h, C.byref(r)))
        return r.value

    def distinctive(self, desc, offsets, indices):
        """OrbMapPoint::ComputeDistinctiveDescriptors for every CSR list of rows of `desc` -> (best position, median)."""
        desc = np.ascontiguousarray(desc, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.int32); indices = np.ascontiguousarray(indices, np.int32)
        n = len(offsets) - 1
        best, med = np.zeros(n, np.int32), np.zeros(n, np.int32)
        self._check(lib().orbm_distinctive(self._h, _ptr(desc), len(desc), _ptr(offsets), _ptr(indices), n, _ptr(best), _ptr(med)))
        return best, med

    def distance_pairs(self, a, b):
        a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
        out = np.zeros(len(a), np.int32)
        self._check(lib().orbm_distance_pairs(self._h, _ptr(a), _ptr(b), len(a), _ptr(out)))
        return out


class Vocabulary:
    """OrbVocabulary's tree on the device: transform5 for whole descriptor sets (orbvocabulary.cpp:203-242)."""

    def __init__(self, child_off, child_ids, node_desc, word_id, weight, L, device=0):
        child_off = np.ascontiguousarray(child_off, np.int32); child_ids = np.ascontiguousarray(child_ids, np.int32)
        node_desc = np.ascontiguousarray(node_desc, np.uint8); word_id = np.ascontiguousarray(word_id, np.int32)
        weight = np.ascontiguousarray(weight, np.float64)
        self._h = C.c_void_p()
        rc = lib().orbv_create(device, len(word_id), _ptr(child_off), _ptr(child_ids), _ptr(node_desc), _ptr(word_id),
                               _ptr(weight), L, C.byref(self._h))
        if rc != 0:
            msg = lib().orbv_last_error(self._h).decode() if self._h else "invalid vocabulary"
            if self._h:
                lib().orbv_destroy(self._h)
                self._h = None
            raise OrbxError(rc, msg)

    def close(self):
        if getattr(self, "_h", None):
            lib().orbv_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, rc):
        if rc < 0:
            raise OrbxError(rc, lib().orbv_last_error(self._h).decode())
        return rc

    def transform(self, desc, levels_up=4):
        """-> (word id, weight, node id at level L - levels_up) per descriptor row."""
        desc = np.ascontiguousarray(desc, np.uint8)
        n = len(desc)
        word, node, w = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.float64)
        self._check(lib().orbv_transform(self._h, _ptr(desc), n, levels_up, _ptr(word), _ptr(w), _ptr(node)))
        return word, w, node

    def transform_device(self, d_desc, stride, n, levels_up, d_out, stream=None):
        self._check(lib().orbv_transform_device(self._h, C.c_void_p(d_desc), stride, n, levels_up, C.c_void_p(d_out),
                                                C.c_void_p(stream) if stream else None))


class MultiExtractor:
    """orbx_multi_*: one handle over several GPUs of the box (frames of a batch sharded over the device slots)."""

    def __init__(self, devices, nfeatures=2000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, max_width=1241, max_height=376,
                 max_batch=64, taps=None, tie_rule=0):
        cfg = Config(nfeatures, scale_factor, nlevels, ini_th, min_th, max_width, max_height, max_batch, 0,
                     (C.c_int * 7)(*(taps or [0] * 7)), tie_rule)
        self._h = C.c_void_p()
        dev = (C.c_int * len(devices))(*devices)
        rc = lib().orbx_multi_create(C.byref(cfg), dev, len(devices), C.byref(self._h))
        if rc != 0:
            msg = lib().orbx_multi_last_error(self._h).decode() if self._h else "orbx_multi_create failed"
            if self._h:
                lib().orbx_multi_destroy(self._h)
            self._h = None
            raise OrbxError(rc, msg)
        self.n_devices = len(devices)

    def close(self):
        if getattr(self, "_h", None):
            lib().orbx_multi_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, rc):
        if rc != 0:
            raise OrbxError(rc, lib().orbx_multi_last_error(self._h).decode())

    @property
    def max_keypoints(self):
        return lib().orbx_multi_max_keypoints(self._h)

    def extract_batch_ptrs(self, ptrs, n, width, height, pitch, out):
        kps, desc, counts = out
        self._check(lib().orbx_multi_extract_batch(self._h, ptrs, n, width, height, pitch, kps.ctypes.data, kps.shape[1],
                                                   desc.ctypes.data, counts.ctypes.data))
        return out

    def extract_batch(self, imgs, out=None):
        imgs = [np.ascontiguousarray(i, np.uint8) for i in imgs]
        h, w = imgs[0].shape
        n, cap = len(imgs), self.max_keypoints
        if out is None:
            out = (np.zeros((n, cap), KP_DTYPE), np.zeros((n, cap, 32), np.uint8), np.zeros(n, np.int32))
        ptrs = Extractor.frame_pointers(imgs)
        return self.extract_batch_ptrs(ptrs, n, w, h, w, out)

    def extract_batch_async(self, ptrs, n, width, height, pitch, out):
        kps, desc, counts = out
        t = C.c_int()
        self._check(lib().orbx_multi_extract_batch_async(self._h, ptrs, n, width, height, pitch, kps.ctypes.data, kps.shape[1],
                                                         desc.ctypes.data, counts.ctypes.data, C.byref(t)))
        return t.value

    def wait(self, ticket):
        self._check(lib().orbx_multi_wait(self._h, int(ticket)))

    def frame_range(self, batch, slot):
        a, b = C.c_int(), C.c_int()
        self._check(lib().orbx_multi_frame_range(self._h, batch, slot, C.byref(a), C.byref(b)))
        return a.value, b.value


class MultiMatcher:
    """orbm_multi_*: query-sharded kNN-2 over several GPUs of one process, records exchanged by the kernels through peer stores."""

    def __init__(self, devices, max_queries=2000, max_train=100000):
        self._h = C.c_void_p()
        dev = (C.c_int * len(devices))(*devices)
        rc = lib().orbm_multi_create(dev, len(devices), max_queries, max_train, C.byref(self._h))
        if rc != 0:
            msg = lib().orbm_multi_last_error(self._h).decode() if self._h else "orbm_multi_create failed"
            if self._h:
                lib().orbm_multi_destroy(self._h)
            self._h = None
            raise OrbxError(rc, msg)

    def close(self):
        if getattr(self, "_h", None):
            lib().orbm_multi_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, rc):
        if rc != 0:
            raise OrbxError(rc, lib().orbm_multi_last_error(self._h).decode())

    def set_train(self, t):
        t = np.ascontiguousarray(t, np.uint8)
        self._check(lib().orbm_multi_set_train(self._h, _ptr(t), len(t)))

    def knn2(self, q):
        q = np.ascontiguousarray(q, np.uint8)
        n = len(q)
        idx, d1, d2 = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int32)
        self._check(lib().orbm_multi_knn2(self._h, _ptr(q), n, _ptr(idx), _ptr(d1), _ptr(d2)))
        return idx, d1, d2


def random_vocabulary(k=10, L=3, seed=0):
    """A synthetic k-ary tree of depth L in the array form of orbv_create (the reference's ORBvoc.txt is not in the tree):
    node ids in creation order (children of a node are consecutive), random descriptors, leaves numbered as words."""
    rng = np.random.default_rng(seed)
    child_off, child_ids, level_of = [0], [], [0]
    frontier, n = [0], 1
    kids = {}
    for lev in range(L):
        nxt = []
        for v in frontier:
            kids[v] = list(range(n, n + k)); n += k
            nxt += kids[v]
        frontier = nxt
    for v in range(n):
        child_ids += kids.get(v, [])
        child_off.append(len(child_ids))
    node_desc = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    word_id = np.full(n, -1, np.int32)
    leaves = [v for v in range(n) if v not in kids]
    word_id[leaves] = np.arange(len(leaves), dtype=np.int32)
    weight = np.where(word_id >= 0, rng.uniform(0.1, 5.0, n), 0.0)
    return np.array(child_off, np.int32), np.array(child_ids, np.int32), node_desc, word_id, weight, L


def filter_projection_candidates(r):
    """Host side of SearchByProjection: drop candidates whose key point already carries an observed map
    point (orbmatcher.cpp:87-89) or whose right coordinate is too far from the projection (:91-96).  `r`: mapping with
    offsets / indices (CSR lists per map point), mp_x, mp_radius (per map point), b_uright, b_occupied (per key point)."""
    off, ind = r["offsets"], r["indices"]
    owner = np.repeat(np.arange(len(off) - 1), np.diff(off))
    ur = r["b_uright"][ind]
    er = np.abs(r["mp_x"][owner] - ur).astype(np.float32)                   # float subtraction, fabs, back to float (:93)
    keep = (r["b_occupied"][ind] == 0) & ~((ur > 0) & (er > r["mp_radius"][owner]))
    lens = np.bincount(owner[keep], minlength=len(off) - 1)
    off2 = np.zeros(len(off), np.int32); off2[1:] = np.cumsum(lens)
    return off2, ind[keep]


def accept_projection_matches(idx1, d1, idx2, d2, octave, nnratio, th_high=100):
    """The acceptance of orbmatcher.cpp:116-123 applied map point by map point, in order (a later map point overwrites
    an earlier one on the same key point) -> assigned[n_keypoints], nmatches."""
    assigned = np.full(len(octave), -1, np.int32)
    n = 0
    for i in range(len(idx1)):
        if d1[i] > th_high:
            continue
        lv = octave[idx1[i]]; lv2 = octave[idx2[i]] if idx2[i] >= 0 else -1
        if lv == lv2 and np.float32(d1[i]) > np.float32(nnratio) * np.float32(d2[i]):
            continue
        assigned[idx1[i]] = i
        n += 1
    return assigned, n


def ratio_test(d1, d2, ratio=0.7, th=100):
    """Host-side acceptance exactly as the reference applies it (orbmatcher.cpp:234-236):
    d1 <= TH and (float)d1 < ratio * (float)d2."""
    d1f = d1.astype(np.float32); d2f = d2.astype(np.float32)
    return (d1 <= th) & (d1f < np.float32(ratio) * d2f)
