#!/usr/bin/env python3
"""Golden fixture for the distinctive-descriptor row (tests/golden/ref_mappoint.npz); run in the BUILD container only.

Source of truth: the reference's own OrbMapPoint::ComputeDistinctiveDescriptors -- src/orbmappoint.cpp compiled
UNMODIFIED against oracle/cvshim (oracle/_ref/libmpref.so, `make -C oracle ref`, glue in oracle/cvshim/mappoint_glue.cpp).
Map points observe 0..40 key frames each (clusters of near-duplicates, unrelated descriptors, lists of length 1 and 2
whose medians tie, some key frames flagged bad) and the descriptor the reference keeps is recorded per point.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "opendlv-perception-vision-orbslam2_b200")]
import orb_oracle_py as O  # noqa: E402
import synth  # noqa: E402

desc, offsets, indices, bad = synth.observation_lists(n_points=400, seed=2027)
out, has = O.ref_distinctive(desc, offsets, indices, bad)
path = os.path.join(ROOT, "tests", "golden", "ref_mappoint.npz")
np.savez_compressed(path, desc=desc, offsets=offsets, indices=indices, bad=bad, out=out, has=has)
print("ref_mappoint.npz", os.path.getsize(path), "points with a descriptor:", int(has.sum()))
