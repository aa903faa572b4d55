#!/usr/bin/env python3
"""Randomised soak of the Hamming matcher and the vocabulary descent against the oracle (not part of the test suite).
usage: soak_match.py [n_cases] [seed]"""
import os, sys, time, traceback
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "opendlv-perception-vision-orbslam2_b200"), os.path.join(ROOT, "oracle")]
import orbx
import orb_oracle_py as O

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
bad = 0
t00 = time.time()
for case in range(n_cases):
    nq = int(np.exp(rng.uniform(0, np.log(4000)))); nt = int(np.exp(rng.uniform(0, np.log(200000))))
    d = f"case {case}: nq={nq} nt={nt}"
    try:
        t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
        q = t[rng.integers(0, nt, nq)].copy()
        flips = rng.integers(0, 60, nq)
        for i in range(nq):
            if flips[i]:
                b = rng.integers(0, 256, flips[i]); np.bitwise_xor.at(q[i], b // 8, (1 << (b % 8)).astype(np.uint8))
        if nt > 3: t[rng.integers(0, nt, max(1, nt // 50))] = t[rng.integers(0, nt, max(1, nt // 50))]      # duplicate rows
        m = orbx.Matcher(max_queries=max(nq, 1), max_train=max(nt, 1))
        gi, g1, g2 = m.knn2(q, t); oi, o1, o2 = O.knn2(q, t, nthreads=8)
        ok = np.array_equal(gi, oi) and np.array_equal(g1, o1) and np.array_equal(g2, o2)
        m.set_train(t); ri, r1, r2 = m.knn2_resident(q)
        ok &= np.array_equal(ri, oi) and np.array_equal(r1, o1) and np.array_equal(r2, o2)
        n = min(nq, nt)
        ok &= np.array_equal(m.distance_pairs(q[:n], t[:n]), np.unpackbits(q[:n] ^ t[:n], axis=1).sum(1))
        m.close()
        # vocabulary: random branching / depth
        k = int(rng.integers(2, 21)); L = int(rng.integers(1, 6 if k <= 10 else 4)); lu = int(rng.integers(0, L + 2))
        voc = orbx.random_vocabulary(k, L, seed=int(rng.integers(0, 1 << 30)))
        V = orbx.Vocabulary(*voc[:5], voc[5])
        f = q[: max(1, min(nq, 3000))]
        gw, gwt, gn = V.transform(f, lu); ow, on = O.voc_transform(voc[0], voc[1], voc[2], voc[3], voc[5], lu, f)
        ok &= np.array_equal(gw, ow) and np.array_equal(gn, on)
        V.close()
        print(d, f"k={k} L={L} lu={lu}", "-> ok" if ok else "-> MISMATCH", flush=True)
        bad += (not ok)
    except Exception:
        bad += 1; print(d, "-> EXCEPTION"); traceback.print_exc()
print(f"soak_match: {n_cases} cases, {bad} bad, {time.time() - t00:.0f} s")
sys.exit(1 if bad else 0)
