#!/usr/bin/env python3
"""Time orbm_search_by_projection (2000 map points against 2000 key points, host arrays in and out)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "opendlv-perception-vision-orbslam2_b200"))
import orbx, synth
W, H = 1241, 376
ex = orbx.Extractor(2000, 1.2, 8, max_width=W, max_height=H)
k, d = ex.extract(synth.scene_s1(W, H, 7))
sf = (np.float32(1.2) ** np.arange(8, dtype=np.float32)).astype(np.float32)
args = (k, np.full(len(k), -1, np.float32), None, d, (0.0, 0.0, float(W), float(H)), d, k["x"] + np.float32(1.25), k["y"].copy(),
        k["octave"].copy(), (np.float32(4.0) * sf[k["octave"]]).astype(np.float32))
m = orbx.Matcher(16, 16)
for _ in range(5):
    m.search_by_projection(*args)
t0 = time.perf_counter()
for _ in range(200):
    r = m.search_by_projection(*args)
dt = (time.perf_counter() - t0) / 200
print(f"search_by_projection: {dt * 1e3:.3f} ms per call, {len(k) / dt / 1e6:.1f} M map points/s, matches {r[2]}")
