import ctypes as C, os, sys, numpy as np
sys.path.insert(0, "opendlv-perception-vision-orbslam2_b200")
import synth
class Cfg(C.Structure):
    _fields_ = [("nfeatures", C.c_int), ("scale", C.c_float), ("nlevels", C.c_int), ("ini", C.c_int), ("min", C.c_int)]
R = C.CDLL("oracle/_ref/libdriverref.so")
R.driverref_check.restype = C.c_int
R.driverref_check.argtypes = [C.POINTER(Cfg)] + [C.c_void_p] * 4 + [C.c_int, C.c_int] + [C.c_float] * 7 + [C.c_void_p]
for (w,h,sa,sb,t1,t2,ratio,dx,dy) in [(1241, 376, 11, 11, 3.0, 15.0, 0.8, 0.7, -0.4), (640, 360, 5, 5, 1.0, 7.0, 0.6, 0.0, 0.0), (752, 480, 9, 10, 3.0, 15.0, 0.9, 1.5, 1.0)]:
    (la, ra), (lb, rb) = synth.stereo_pair(w, h, sa), synth.stereo_pair(w, h, sb)
    imgs = [np.ascontiguousarray(a, np.uint8) for a in (la, ra, lb, rb)]
    out = np.full(32, -99, np.int32)
    rc = R.driverref_check(C.byref(Cfg(2000, 1.2, 8, 20, 7)), *[a.ctypes.data for a in imgs], w, h, 386.1, 0.537, t1, t2, ratio, dx, dy, out.ctypes.data)
    print(rc, out[:28].reshape(7,4).tolist(), 'us: SearchByProjection ref/gpu', out[28], out[29], 'SearchByBoW ref/gpu', out[30], out[31])
