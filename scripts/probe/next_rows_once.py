#!/usr/bin/env python3
"""One call of every 'next'-row entry point on a KITTI-sized stereo batch (for an ncu capture of their kernels)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "opendlv-perception-vision-orbslam2_b200"))
import orbx, synth
W, H, pairs = 1241, 376, 8
frames = synth.stereo_batch(2, W, H, pairs)
ex = orbx.Extractor(2000, 1.2, 8, max_width=W, max_height=H, max_batch=2 * pairs)
kps, desc, cnt = ex.extract_batch(frames)
kps, desc, cnt = kps.copy(), desc.copy(), cnt.copy()
orbx.stereo_match_batch(ex, pairs, 0, 1, 2, 386.1448, 0.5372)
ex.filter_keypoints((300.0, 900.0, 80.0, 300.0), 0, 2 * pairs)
m = orbx.Matcher(4096, 8192)
k0, d0, k1, d1 = kps[0, :cnt[0]], desc[0, :cnt[0]], kps[2, :cnt[2]], desc[2, :cnt[2]]
sf = (np.float32(1.2) ** np.arange(8, dtype=np.float32)).astype(np.float32)
bounds = (0.0, 0.0, float(W), float(H))
args = (k1, np.full(len(k1), -1, np.float32), None, d1, bounds, d0, k0["x"] + np.float32(1.25), k0["y"].copy(), k0["octave"].copy(),
        (np.float32(4.0) * sf[k0["octave"]]).astype(np.float32))
m.search_by_projection(*args)
m.area_distances(k1, d1, bounds, d0, args[6], args[7], args[9], k0["octave"] - 1, k0["octave"] + 1)
rng = np.random.default_rng(0)
offs = np.concatenate([[0], np.cumsum(rng.integers(2, 25, 2000))]).astype(np.int32)
inds = rng.integers(0, len(d0), int(offs[-1])).astype(np.int32)
m.distinctive(d0, offs, inds)
m.knn2_csr(d0[:2000], d1, offs[:2001] if len(d0) >= 2000 else offs[:len(d0) + 1], rng.integers(0, len(d1), int(offs[min(2000, len(d0))])).astype(np.int32))
voc = orbx.random_vocabulary(10, 5, seed=1)
V = orbx.Vocabulary(*voc[:5], voc[5])
V.transform(np.concatenate([d0, d1]), 4)
print("next rows once: ok")
