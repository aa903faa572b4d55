#!/usr/bin/env python3
"""Randomised soak of the extractor against the oracle (not part of the test suite): random frame sizes, feature counts,
level counts, scale factors, thresholds, batch sizes and scene kinds; every frame must be bit-identical.
usage: soak.py [n_cases] [seed] [big]"""
import os, sys, time, traceback
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "opendlv-perception-vision-orbslam2_b200"), os.path.join(ROOT, "oracle")]
import orbx, synth
import orb_oracle_py as O

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
bad = 0
t00 = time.time()
for case in range(n_cases):
    big = len(sys.argv) > 3 and sys.argv[3] == "big"
    w = int(rng.integers(160, 4100 if big else 2100)); h = int(rng.integers(120, min(w, 2200 if big else 1300)))
    if w >= 5 * h:
        h = w // 4
    nl = int(rng.integers(1, 13 if big else 9)); sf = float(rng.choice([1.1, 1.15, 1.2, 1.25, 1.3, 1.4]))
    nf = int(rng.integers(50, 8000 if big else 4000)); ini = int(rng.integers(8, 60)); mn = int(rng.integers(1, ini))
    batch = int(rng.integers(1, 3 if big else 5))
    kinds = [("s1", "s2", "noise", "flat", "steps")[int(rng.integers(0, 5))] for _ in range(batch)]
    desc_s = f"case {case}: {w}x{h} nf={nf} nl={nl} sf={sf} th={ini}/{mn} batch={batch} {kinds}"
    try:
        oex = O.Extractor(nf, sf, nl, ini, mn)
    except ValueError:
        print(desc_s, "-> oracle rejects the configuration"); continue
    imgs = []
    for k in kinds:
        s = int(rng.integers(0, 1 << 30))
        if k == "s1": imgs.append(synth.scene_s1(w, h, s))
        elif k == "s2": imgs.append(synth.scene_s2(w, h, s))
        elif k == "noise": imgs.append(np.random.default_rng(s).integers(0, 256, (h, w), dtype=np.uint8))
        elif k == "flat": imgs.append(np.full((h, w), s % 256, np.uint8))
        else: imgs.append(((np.add.outer(np.arange(h) // 13, np.arange(w) // 17) % 2) * (40 + s % 200)).astype(np.uint8))
    try:
        ex = orbx.Extractor(nf, sf, nl, ini, mn, max_width=w, max_height=h, max_batch=batch)
    except orbx.OrbxError as e:
        print(desc_s, "-> rejected:", str(e)[:80]); continue
    try:
        kps, desc, cnt = ex.extract_batch(imgs)
        ok = True
        for f, img in enumerate(imgs):
            okp, od = oex.extract(img)
            n = int(cnt[f])
            if n != len(okp) or kps[f, :n].tobytes() != okp.tobytes() or not np.array_equal(desc[f, :n], od):
                ok = False
                print(desc_s, f"-> MISMATCH frame {f}: n {n} vs {len(okp)}")
        if not ok: bad += 1
        else: print(desc_s, "-> ok", int(cnt.sum()))
    except orbx.OrbxError as e:
        bad += 1                                  # a configuration orbx_create accepted must run
        print(desc_s, "-> error:", str(e)[:100])
    except Exception:
        bad += 1; traceback.print_exc()
    ex.close()
print(f"soak: {n_cases} cases, {bad} bad, {time.time() - t00:.0f} s")
sys.exit(1 if bad else 0)
