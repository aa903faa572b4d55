#!/usr/bin/env python3
"""Wall time of the reference's own stereo OrbFrame constructor in four builds of the same unmodified src/orbframe.cpp:
all-reference (CPU), with the drop-in extractor (pyramid levels downloaded when the reference's stereo matcher reads them), with the
drop-in extractor and stereo matcher (no pyramid leaves HBM), and the latter with ORBX_ADAPTER_EAGER_PYRAMID (all levels downloaded
in every call, the adapter's earlier behaviour) (GPU box only)."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "opendlv-perception-vision-orbslam2_b200"))
import synth
class Cfg(C.Structure):
    _fields_ = [("nfeatures", C.c_int), ("scale", C.c_float), ("nlevels", C.c_int), ("ini", C.c_int), ("min", C.c_int)]
w, h = 1241, 376
l, r = synth.stereo_pair(w, h, 11)
l = np.ascontiguousarray(l); r = np.ascontiguousarray(r)
for name, reps in (("libframeref.so", 5), ("libdropinref.so", 50), ("libdropin2ref.so", 50), ("libdropin3ref.so", 50)):
    p = os.path.join(ROOT, "oracle", "_ref", name)
    if not os.path.exists(p):
        print(name, "not built"); continue
    R = C.CDLL(p)
    R.frameref_time_constructor.restype = C.c_double
    R.frameref_time_constructor.argtypes = [C.POINTER(Cfg), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int]
    us = R.frameref_time_constructor(C.byref(Cfg(2000, 1.2, 8, 20, 7)), l.ctypes.data, r.ctypes.data, w, h, 386.1, reps)
    print(f"{name:20s} OrbFrame stereo constructor: {us / 1e3:8.3f} ms", flush=True)
