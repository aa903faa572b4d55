"""GPU: the reference's own ORBmatcher drivers against their liborbx-backed replacements (cpp/orbmatcher_drivers_b200.hpp,
class ORBmatcherB200) inside the reference's own data model.  oracle/_ref/libdriverref.so holds the reference's
orbframe.cpp / orbmatcher.cpp / orbmappoint.cpp / orbextractor.cpp compiled unmodified plus the drop-in header; it is built
in the build container (`make -C oracle ref`) and travels to the GPU box with the other built libraries."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libdriverref.so")


SCRIPT = """
import ctypes as C, sys, numpy as np
sys.path[:0] = {path!r}
import synth
class Cfg(C.Structure):
    _fields_ = [("nfeatures", C.c_int), ("scale", C.c_float), ("nlevels", C.c_int), ("ini", C.c_int), ("min", C.c_int)]
R = C.CDLL({lib!r})
R.driverref_check.restype = C.c_int
R.driverref_check.argtypes = [C.POINTER(Cfg)] + [C.c_void_p] * 4 + [C.c_int, C.c_int] + [C.c_float] * 7 + [C.c_void_p]
w, h, sa, sb, th_points, th_frames, ratio, dx, dy = {args!r}
(la, ra), (lb, rb) = synth.stereo_pair(w, h, sa), synth.stereo_pair(w, h, sb)
imgs = [np.ascontiguousarray(a, np.uint8) for a in (la, ra, lb, rb)]
out = np.full(64, -99, np.int32)
rc = R.driverref_check(C.byref(Cfg(2000, 1.2, 8, 20, 7)), *[a.ctypes.data for a in imgs], w, h, 386.1, 0.537,
                       th_points, th_frames, ratio, dx, dy, out.ctypes.data)
print("RESULT", rc, *out.tolist())
"""


@pytest.mark.skipif(not os.path.exists(LIB), reason="libdriverref.so is built where the reference tree is mounted")
@pytest.mark.parametrize("w,h,sa,sb,th_points,th_frames,ratio,dx,dy", [
    (1241, 376, 11, 11, 3.0, 15.0, 0.8, 0.7, -0.4), (640, 360, 5, 5, 1.0, 7.0, 0.6, 0.0, 0.0), (752, 480, 9, 10, 3.0, 15.0, 0.9, 1.5, 1.0)])
def test_drop_in_drivers_equal_reference_drivers(w, h, sa, sb, th_points, th_frames, ratio, dx, dy):
    # the reference code + liborbx run in a process of their own: a fault there must not take the test session down
    code = SCRIPT.format(path=sys.path[:4], lib=LIB, args=(w, h, sa, sb, th_points, th_frames, ratio, dx, dy))
    p = subprocess.run([sys.executable, "-c", code], check=True, timeout=300, capture_output=True, text=True)
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT")][-1].split()
    rc, out = int(line[1]), np.array([int(v) for v in line[2:]], np.int32)
    assert rc == 0
    n_ref, n_gpu, bad, assigned = out[:4]
    assert n_ref == n_gpu and bad == 0 and assigned > 0, f"SearchByProjection(frame, map points): {out[:4]}"
    for m, name in enumerate(("forward", "backward", "neither")):
        n_ref, n_gpu, bad, assigned = out[4 + 4 * m: 8 + 4 * m]
        assert n_ref == n_gpu and bad == 0, f"SearchByProjection(current, last) {name}: {out[4 + 4 * m: 8 + 4 * m]}"
    assert out[4:16:4].max() > 20          # the frame-to-frame search did find matches in some mode
    n_ref, n_gpu, bad, assigned = out[16:20]
    assert n_ref == n_gpu and bad == 0 and assigned > 0, f"SearchByBoW(key frame, frame): {out[16:20]}"
    n_ref, n_gpu, bad, assigned = out[20:24]
    assert n_ref == n_gpu and bad == 0 and assigned > 0, f"SearchForInitialization: {out[20:24]}"
    n_ref, n_gpu, bad, assigned = out[24:28]
    assert n_ref == n_gpu and bad == 0 and assigned > 0, f"SearchByProjection(current, key frame): {out[24:28]}"
    n_ref, n_gpu, bad, assigned = out[32:36]
    assert n_ref == n_gpu and bad == 0 and assigned > 0, f"SearchByBoW(key frame, key frame): {out[32:36]}"
    for only, name in enumerate(("all features", "stereo only")):
        n_ref, n_gpu, bad, pairs = out[36 + 4 * only: 40 + 4 * only]
        assert n_ref == n_gpu and bad == 0 and (pairs > 0 or only == 1), f"SearchForTriangulation ({name}): {out[36 + 4 * only: 40 + 4 * only]}"
    n_ref, n_gpu, bad, held = out[44:48]
    assert n_ref == n_gpu and bad == 0 and held > 0, f"SearchByProjection(key frame, Scw): {out[44:48]}"
    n_ref, n_gpu, bad, held, repl = out[48:53]
    assert n_ref == n_gpu and bad == 0 and held > 0, f"Fuse(key frame, Scw): {out[48:53]}"
    n_ref, n_gpu, bad, held, corrupt = out[53:58]
    assert n_ref == n_gpu and bad == 0 and held > 0, f"Fuse(key frame, points): {out[53:58]}"
    n_ref, n_gpu, bad, held = out[58:62]
    assert n_ref == n_gpu and bad == 0 and held > 0, f"SearchBySim3: {out[58:62]}"
    print("key-frame drivers:", out[32:62].tolist())
