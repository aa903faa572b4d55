#!/usr/bin/env python3
"""Golden fixture for the candidate-list row (tests/golden/ref_projection.npz); run in the BUILD container only.

Source of truth: the reference's own ORBmatcher::SearchByProjection(frame, map points, th) -- src/orbmatcher.cpp
(with its own DescriptorDistance), src/orbframe.cpp, src/orbmappoint.cpp, src/orbextractor.cpp compiled UNMODIFIED
against oracle/cvshim (oracle/_ref/libframeref.so, `make -C oracle ref`, driver in oracle/cvshim/frame_glue.cpp).
Stored per case: the map points' descriptors and tracking fields, the searched frame's descriptors / octaves / mvuRight /
occupied flags, the candidate lists the reference's own GetFeaturesInArea produced, and the assignment it made.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "opendlv-perception-vision-orbslam2_b200")]
import orb_oracle_py as O  # noqa: E402
import synth  # noqa: E402

# w, h, seed A, seed B, dx, dy, th, nnratio, nfeatures
CASES = [(640, 360, 5, 5, 0.7, 0.4, 3.0, 0.8, 1000), (800, 240, 21, 21, -1.5, 1.0, 1.0, 0.6, 1200), (640, 360, 5, 6, 0.0, 0.0, 3.0, 0.9, 1000)]
out = {"cases": np.array(CASES, np.float64)}
for c, (w, h, sa, sb, dx, dy, th, ratio, nf) in enumerate(CASES):
    r = O.ref_search_by_projection(synth.stereo_pair(w, h, sa), synth.stereo_pair(w, h, sb), 386.1, 0.537, th=th, nnratio=ratio,
                                   dx=dx, dy=dy, nfeatures=nf)
    for k, v in r.items():
        out[f"{k}_{c}"] = np.asarray(v)
    print(w, h, "map points", len(r["mp_desc"]), "candidates", len(r["indices"]), "matches", r["nmatches"])
path = os.path.join(ROOT, "tests", "golden", "ref_projection.npz")
np.savez_compressed(path, **out)
print("ref_projection.npz", os.path.getsize(path))

# ---- the sequential rule inside the call (orbmatcher.cpp:87-89 after :121): map points that carry observations, as every
# local map point of Tracking::SearchLocalPoints does, and twins that collide on one key point.
# w, h, seed A, seed B, dx, dy, th, nnratio, nfeatures, map points per key point, obs_mod (map point k is observed unless k % obs_mod == 0)
CASES_OBS = [(640, 360, 5, 5, 0.7, 0.4, 3.0, 0.8, 1000, 2, 4), (800, 240, 21, 21, -1.5, 1.0, 1.0, 0.6, 1200, 2, 3),
             (640, 360, 5, 6, 0.0, 0.0, 3.0, 0.9, 1000, 3, 5), (1241, 376, 11, 11, 0.5, 0.0, 3.0, 0.8, 1000, 1, 7)]
out = {"cases": np.array(CASES_OBS, np.float64)}
for c, (w, h, sa, sb, dx, dy, th, ratio, nf, dup, obs) in enumerate(CASES_OBS):
    r = O.ref_search_by_projection(synth.stereo_pair(w, h, sa), synth.stereo_pair(w, h, sb), 386.1, 0.537, th=th, nnratio=ratio,
                                   dx=dx, dy=dy, nfeatures=nf, mp_dup=dup, obs_mod=obs)
    for k in ("mp_desc", "mp_x", "mp_y", "mp_level", "mp_radius", "mp_observed", "b_keys", "b_desc", "b_uright", "b_occupied", "bounds", "assigned", "nmatches"):
        out[f"{k}_{c}"] = np.asarray(r[k])
    print(w, h, "map points", len(r["mp_desc"]), "observed", int(r["mp_observed"].sum()), "matches", r["nmatches"])
path = os.path.join(ROOT, "tests", "golden", "ref_projection_observed.npz")
np.savez_compressed(path, **out)
print("ref_projection_observed.npz", os.path.getsize(path))
