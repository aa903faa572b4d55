#!/usr/bin/env python3
"""Randomised soak of the 'next' rows against the oracle (not part of the test suite): stereo matching (single + batch),
FilterKeyPoints, grid assignment, SearchByProjection / GetFeaturesInArea on real extraction output, distinctive descriptors,
candidate lists.  usage: soak_next.py [n_cases] [seed]"""
import os, sys, time, traceback
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "opendlv-perception-vision-orbslam2_b200"), os.path.join(ROOT, "oracle")]
import orbx, synth
import orb_oracle_py as O

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
bad = 0
t00 = time.time()
m = orbx.Matcher(4096, 8192)
for case in range(n_cases):
    w = int(rng.integers(320, 1700)); h = int(rng.integers(200, min(w, 900)))
    if w >= 4 * h: h = w // 3
    nl = int(rng.integers(3, 9)); nf = int(rng.integers(300, 3000)); pairs = int(rng.integers(1, 4))
    mbf = float(rng.uniform(100, 500)); mb = float(rng.choice([0.0, 0.1, 0.54, 3.0]))
    d = f"case {case}: {w}x{h} nf={nf} nl={nl} pairs={pairs} mbf={mbf:.1f} mb={mb}"
    try:
        frames = synth.stereo_batch(int(rng.integers(1, 10000)), w, h, pairs)
        try:
            ex = orbx.Extractor(nf, 1.2, nl, max_width=w, max_height=h, max_batch=2 * pairs)
        except orbx.OrbxError as e:
            print(d, "-> rejected", str(e)[:60]); continue
        kps, desc, cnt = ex.extract_batch(frames)
        kps, desc, cnt = kps.copy(), desc.copy(), cnt.copy()
        ok = True
        # stereo, batch and single
        u, dep, nl_, nm = orbx.stereo_match_batch(ex, pairs, 0, 1, 2, mbf, mb)
        for p in range(pairs):
            eL, eR = O.Extractor(nf, 1.2, nl), O.Extractor(nf, 1.2, nl)
            kl, dl = eL.extract(frames[2 * p]); kr, dr = eR.extract(frames[2 * p + 1])
            ou, od, on = O.stereo_matches(eL, eR, kl, dl, kr, dr, mbf, mb)
            if not (nl_[p] == len(kl) and nm[p] == on and np.array_equal(u[p, :len(kl)].view(np.uint32), ou.view(np.uint32))
                    and np.array_equal(dep[p, :len(kl)].view(np.uint32), od.view(np.uint32))):
                ok = False; print(d, f"-> STEREO MISMATCH pair {p}")
            if p == 0:
                su, sd, snm = orbx.stereo_match(ex, 0, ex, 1, mbf, mb)
                if not (np.array_equal(su.view(np.uint32), ou.view(np.uint32)) and snm == on):
                    ok = False; print(d, "-> STEREO(single) MISMATCH")
        # grid + projection + area on frame 0's key points against frame 1's
        k0, d0 = kps[0, :cnt[0]], desc[0, :cnt[0]]; k1, d1 = kps[1, :cnt[1]], desc[1, :cnt[1]]
        bounds = (0.0, 0.0, float(w), float(h))
        if len(k0) and len(k1):
            gs, gi = m.assign_grid(k1, bounds); os_, oi_ = O.assign_grid(k1, bounds)
            if not (np.array_equal(gs, os_) and np.array_equal(gi, oi_)): ok = False; print(d, "-> GRID MISMATCH")
            sf = (np.float32(1.2) ** np.arange(8, dtype=np.float32)).astype(np.float32)
            args = (k1, rng.uniform(-1, w, len(k1)).astype(np.float32), (rng.random(len(k1)) < 0.1).astype(np.uint8), d1, bounds, d0,
                    k0["x"] + np.float32(rng.uniform(-20, 20)), k0["y"] + np.float32(rng.uniform(-3, 3)), k0["octave"].copy(),
                    (np.float32(rng.choice([2.5, 4.0, 15.0])) * sf[k0["octave"]]).astype(np.float32))
            g = m.search_by_projection(*args, nnratio=float(rng.uniform(0.5, 1.0)) if False else 0.8, th_high=100); o = O.search_by_projection(*args, nnratio=0.8, th_high=100)
            if not (np.array_equal(g[0], o[0]) and np.array_equal(g[1], o[1]) and g[2] == o[2]): ok = False; print(d, "-> PROJECTION MISMATCH")
            l0 = k0["octave"] - 1; l1 = k0["octave"] + 1
            ga = m.area_distances(k1, d1, bounds, d0, args[6], args[7], args[9], l0, l1); oa = O.area_distances(k1, d1, bounds, d0, args[6], args[7], args[9], l0, l1, cap=1 << 22)
            if not (np.array_equal(ga[0], oa[0]) and np.array_equal(ga[1], oa[1]) and np.array_equal(ga[2], oa[2])): ok = False; print(d, "-> AREA MISMATCH")
        # filter on every frame
        box = (float(rng.uniform(0, w / 2)), float(rng.uniform(w / 2, w)), float(rng.uniform(0, h / 2)), float(rng.uniform(h / 2, h)))
        ex.filter_keypoints(box, 0, 2 * pairs)
        fk, fd, fc = ex.fetch_results(2 * pairs)
        for f in range(2 * pairs):
            okp, odd = O.filter_keypoints(kps[f, :cnt[f]], desc[f, :cnt[f]], box)
            if not (fc[f] == len(okp) and fk[f, :fc[f]].tobytes() == okp.tobytes() and np.array_equal(fd[f, :fc[f]], odd)):
                ok = False; print(d, f"-> FILTER MISMATCH frame {f}")
        # distinctive descriptors + candidate lists over this frame's descriptors
        if len(k0) > 10:
            sizes = rng.integers(0, 40, 200); offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
            inds = rng.integers(0, len(d0), int(offs[-1])).astype(np.int32)
            gb = m.distinctive(d0, offs, inds); ob = O.distinctive(d0, offs, inds)
            if not (np.array_equal(gb[0], ob[0]) and np.array_equal(gb[1], ob[1])): ok = False; print(d, "-> DISTINCTIVE MISMATCH")
            if len(k1):
                inds2 = rng.integers(0, len(d1), int(offs[-1])).astype(np.int32)
                q = d0[rng.integers(0, len(d0), 200)]
                gc = m.knn2_csr(q, d1, offs, inds2); oc = O.knn2_csr(q, d1, offs, inds2)
                if not all(np.array_equal(a, b) for a, b in zip(gc, oc)): ok = False; print(d, "-> CSR MISMATCH")
        if ok: print(d, "-> ok", int(cnt.sum()), "kp,", int(nm.sum()), "stereo matches")
        else: bad += 1
        ex.close()
    except Exception:
        bad += 1; print(d, "-> EXCEPTION"); traceback.print_exc()
print(f"soak_next: {n_cases} cases, {bad} bad, {time.time() - t00:.0f} s")
sys.exit(1 if bad else 0)
