#!/usr/bin/env python3
"""Four host threads, one extractor + one matcher handle each, different frame sizes, calls interleaving freely (graph
capture, shared-memory opt-ins and stream work of different handles overlap) -- every result against precomputed oracle
results (not part of the test suite).  usage: soak_threads.py [calls_per_thread]"""
import os, sys, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "opendlv-perception-vision-orbslam2_b200"), os.path.join(ROOT, "oracle")]
import orbx, synth
import orb_oracle_py as O

calls = int(sys.argv[1]) if len(sys.argv) > 1 else 40
cfgs = [(1241, 376, 2000, 8), (752, 480, 1200, 8), (556, 514, 3662, 7), (640, 360, 800, 5)]
pools, wants = [], []
for (w, h, nf, nl) in cfgs:
    imgs = [synth.scene_s1(w, h, 7 + i) for i in range(6)]
    oex = O.Extractor(nf, 1.2, nl)
    pools.append(imgs); wants.append([oex.extract(im) for im in imgs])
errors = []

def worker(t):
    w, h, nf, nl = cfgs[t]
    rng = np.random.default_rng(t)
    try:
        ex = orbx.Extractor(nf, 1.2, nl, max_width=w, max_height=h, max_batch=6)
        m = orbx.Matcher(4096, 4096)
        for c in range(calls):
            b = int(rng.integers(1, 7)); idx = rng.integers(0, 6, b)
            kps, desc, cnt = ex.extract_batch([pools[t][i] for i in idx])
            for f in range(b):
                okp, od = wants[t][idx[f]]
                if not (cnt[f] == len(okp) and kps[f, :cnt[f]].tobytes() == okp.tobytes() and np.array_equal(desc[f, :cnt[f]], od)):
                    errors.append(f"thread {t} call {c} frame {f}: extraction mismatch")
            if b >= 2:
                q, tr = desc[0, :cnt[0]], desc[1, :cnt[1]]
                gi, g1, g2 = m.knn2(q, tr)
                oi, o1, o2 = O.knn2(q, tr)
                if not (np.array_equal(gi, oi) and np.array_equal(g1, o1) and np.array_equal(g2, o2)):
                    errors.append(f"thread {t} call {c}: knn2 mismatch")
        ex.close(); m.close()
    except Exception as e:
        errors.append(f"thread {t}: {type(e).__name__}: {e}")

t0 = time.time()
ths = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
for t in ths: t.start()
for t in ths: t.join()
print(f"soak_threads: 4 threads x {calls} calls, {len(errors)} errors, {time.time() - t0:.0f} s")
for e in errors[:10]: print(" ", e)
sys.exit(1 if errors else 0)
