#!/usr/bin/env python3
"""Corner configurations of the extractor against the oracle: 1-10 features, frames barely larger than one FAST cell, one
level, scale factors at both ends of the supported range, equal thresholds, thresholds at their limits."""
import os, sys, itertools
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "opendlv-perception-vision-orbslam2_b200"), os.path.join(ROOT, "oracle")]
import orbx, synth
import orb_oracle_py as O
bad = ran = 0
rng = np.random.default_rng(5)
for (w, h), nf, nl, sf, (ini, mn) in itertools.product([(64, 64), (70, 63), (97, 71), (128, 96), (200, 150), (333, 90)], [1, 2, 5, 40],
                                                        [1, 2, 4], [1.01, 1.2, 1.35], [(20, 7), (7, 7), (254, 127), (2, 1)]):
    d = f"{w}x{h} nf={nf} nl={nl} sf={sf} th={ini}/{mn}"
    try:
        oex = O.Extractor(nf, sf, nl, ini, mn)
    except ValueError:
        continue
    try:
        ex = orbx.Extractor(nf, sf, nl, ini, mn, max_width=w, max_height=h, max_batch=2)
    except orbx.OrbxError:
        continue
    imgs = [synth.scene_s1(w, h, int(rng.integers(0, 1 << 20))), rng.integers(0, 256, (h, w), dtype=np.uint8)]
    try:
        kps, desc, cnt = ex.extract_batch(imgs)
        for f in range(2):
            try:
                okp, od = oex.extract(imgs[f])
            except Exception as e:
                print(d, "oracle failed:", e); continue
            if not (cnt[f] == len(okp) and kps[f, :cnt[f]].tobytes() == okp.tobytes() and np.array_equal(desc[f, :cnt[f]], od)):
                bad += 1; print(d, f"-> MISMATCH frame {f}: {cnt[f]} vs {len(okp)}")
        ran += 1
    except orbx.OrbxError as e:
        msg = str(e)
        if "ERR_SHAPE" in msg: continue
        bad += 1; print(d, "-> error:", msg[:120])
    ex.close()
print(f"soak_edges: {ran} configurations ran, {bad} bad")
sys.exit(1 if bad else 0)
