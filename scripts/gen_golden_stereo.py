#!/usr/bin/env python3
"""Golden fixture for the stereo row (tests/golden/ref_stereo.npz); run in the BUILD container only.

Source of truth: the reference's own OrbFrame -- src/orbframe.cpp (stereo constructor + ComputeStereoMatches) on top of
src/orbextractor.cpp, both compiled UNMODIFIED against oracle/cvshim (oracle/_ref/libframeref.so, `make -C oracle ref`,
glue in oracle/cvshim/frame_glue.cpp).  Stored per case: the key points and descriptors the reference extracted (its
monotone heap order), a SHA-256 of every pyramid level of both images, the mvuRight / m_depths it computed and its
m_grid (AssignFeaturesToGrid); the third case hands the constructor a bounding box (FilterKeyPoints).  The images
are regenerated from the seed (synth.stereo_pair) and the pyramids from the oracle (pinned bit for bit elsewhere).
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "opendlv-perception-vision-orbslam2_b200")]
import orb_oracle_py as O  # noqa: E402
import synth  # noqa: E402

CASES = [(640, 360, 5, 1000, 6, 200.0, 0.4), (800, 240, 21, 1200, 8, 386.1, 0.537),   # w, h, seed, nfeatures, nlevels, mbf, mb
         (640, 360, 5, 1000, 6, 200.0, 0.4)]
# bounding box handed to the OrbFrame constructor (x0, x1, y0, y1): FilterKeyPoints drops what is strictly inside; zeros = none
BOXES = [(0, 0, 0, 0), (0, 0, 0, 0), (200.0, 420.0, 100.0, 260.0)]
out = {"cases": np.array(CASES, np.float64), "boxes": np.array(BOXES, np.float32)}
for c, (w, h, seed, nf, nl, mbf, mb) in enumerate(CASES):
    left, right = synth.stereo_pair(w, h, seed)
    r = O.ref_stereo_frame(left, right, mbf, mb, nfeatures=nf, nlevels=nl, bbox=BOXES[c] if BOXES[c][1] > 0 else None)
    for k in ("kl", "dl", "kr", "dr", "uRight", "depth", "grid_start", "grid_items"):
        out[f"{k}_{c}"] = r[k]
    out[f"sha_{c}"] = np.array([hashlib.sha256(a.tobytes()).hexdigest() for a in r["levelsL"] + r["levelsR"]])
    print(w, h, len(r["kl"]), len(r["kr"]), "matches", int((r["uRight"] >= 0).sum()))
path = os.path.join(ROOT, "tests", "golden", "ref_stereo.npz")
np.savez_compressed(path, **out)
print("ref_stereo.npz", os.path.getsize(path))
