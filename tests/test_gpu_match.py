"""GPU parity of the Hamming matcher: liborbx orbm_* vs the CPU oracle loop (bit-exact)."""
import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu


def test_knn2_small_vs_oracle(oracle):
    import orbx
    q, t = synth.matching_set(300, 5000, seed=77)
    t[100] = q[5]; t[4000] = q[5]           # exact duplicates: lowest index wins, d2 == d1 == 0
    t[50] = t[51]                            # duplicate train rows
    m = orbx.Matcher(max_queries=300, max_train=5000)
    idx, d1, d2 = m.knn2(q, t)
    oi, o1, o2 = oracle.knn2(q, t)
    assert np.array_equal(idx, oi) and np.array_equal(d1, o1) and np.array_equal(d2, o2)
    assert idx[5] == 100 and d1[5] == 0 and d2[5] == 0
    m.close()


@pytest.mark.parametrize("nq,nt", [(1, 1), (1, 2), (7, 255), (129, 257), (128, 256), (33, 1000)])
def test_knn2_ragged_sizes(oracle, nq, nt):
    import orbx
    rng = np.random.default_rng(nq * 1000 + nt)
    q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    m = orbx.Matcher(max_queries=256, max_train=2048)
    idx, d1, d2 = m.knn2(q, t)
    oi, o1, o2 = oracle.knn2(q, t)
    assert np.array_equal(idx, oi) and np.array_equal(d1, o1) and np.array_equal(d2, o2)
    m.close()


def test_knn2_distance_256_is_never_a_match(oracle):
    import orbx
    q = np.zeros((2, 32), np.uint8)
    t = np.full((3, 32), 255, np.uint8)      # all 256 bits differ: 'dist < 256' never fires (orbmatcher.cpp:208-232)
    m = orbx.Matcher(max_queries=8, max_train=8)
    idx, d1, d2 = m.knn2(q, t)
    oi, o1, o2 = oracle.knn2(q, t)
    assert idx.tolist() == [-1, -1] and d1.tolist() == [256, 256] and d2.tolist() == [256, 256]
    assert np.array_equal(idx, oi) and np.array_equal(d1, o1) and np.array_equal(d2, o2)
    m.close()


def test_knn2_full_size_properties(oracle):
    """C5 at full size (2000 x 100000): every query against an exact matrix-product restatement, a sample against the
    oracle's sequential loop, and size-independent properties."""
    import orbx
    q, t = synth.matching_set(2000, 100000, seed=5000)
    m = orbx.Matcher(max_queries=2000, max_train=100000)
    idx, d1, d2 = m.knn2(q, t)
    # (1) reported distance is the true distance to the reported index; (2) d1 <= d2
    x = np.unpackbits(q ^ t[idx], axis=1).sum(1)
    assert np.array_equal(x, d1)
    assert (d1 <= d2).all()
    # (3) a sample of queries against the oracle's sequential loop
    sel = np.arange(0, 2000, 40)
    oi, o1, o2 = oracle.knn2(q[sel], t, nthreads=8)
    assert np.array_equal(idx[sel], oi) and np.array_equal(d1[sel], o1) and np.array_equal(d2[sel], o2)
    # (3b) ALL 2000 queries against an independent exact restatement: Hamming distance = |a| + |b| - 2 a.b on the unpacked bits
    # (float32 matrix products of 0/1 values up to 256 are exact), best = first minimum in index order (orbmatcher.cpp:208-232:
    # strict '<' keeps the earliest), second = minimum over the other rows
    tb = np.unpackbits(t, axis=1).astype(np.float32)
    tn = tb.sum(1)
    for a in range(0, 2000, 250):
        qb = np.unpackbits(q[a:a + 250], axis=1).astype(np.float32)
        D = (qb.sum(1)[:, None] + tn[None, :] - 2.0 * (qb @ tb.T)).astype(np.int32)
        bi = D.argmin(1)
        b1 = D[np.arange(len(bi)), bi]
        D[np.arange(len(bi)), bi] = 1 << 20
        b2 = D.min(1)
        assert np.array_equal(idx[a:a + 250], bi) and np.array_equal(d1[a:a + 250], b1) and np.array_equal(d2[a:a + 250], b2), a
    # (4) resident path and permutation of the query order give the same per-query answers
    m.set_train(t)
    perm = np.random.default_rng(1).permutation(2000)
    i2, a2, b2 = m.knn2_resident(q[perm])
    assert np.array_equal(i2, idx[perm]) and np.array_equal(a2, d1[perm]) and np.array_equal(b2, d2[perm])
    # ratio test as the reference applies it on the host
    acc = orbx.ratio_test(d1, d2, 0.7, 100)
    assert acc.sum() > 0
    m.close()


def test_distance_pairs(oracle):
    import orbx
    rng = np.random.default_rng(3)
    a = rng.integers(0, 256, (500, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (500, 32), dtype=np.uint8)
    m = orbx.Matcher(max_queries=512, max_train=512)
    d = m.distance_pairs(a, b)
    ref = np.array([oracle.descriptor_distance(a[i], b[i]) for i in range(500)], np.int32)
    assert np.array_equal(d, ref)
    m.close()


def _random_csr(rng, nq, nt, max_len):
    lens = rng.integers(0, max_len + 1, nq)
    lens[rng.integers(0, nq, max(1, nq // 10))] = 0            # empty lists (vIndices.empty(), orbmatcher.cpp:71)
    offsets = np.zeros(nq + 1, np.int32); offsets[1:] = np.cumsum(lens)
    indices = rng.integers(0, nt, int(offsets[-1])).astype(np.int32)   # unsorted, with repeats
    return offsets, indices


def test_knn2_csr_vs_oracle(oracle):
    """Candidate-list matching = inner loop of SearchByProjection (orbmatcher.cpp:76-114)."""
    import orbx
    rng = np.random.default_rng(12)
    q, t = synth.matching_set(500, 4000, seed=8)
    t[17] = t[18] = q[3]                                       # equal distances inside one list: earlier position wins
    offsets, indices = _random_csr(rng, 500, 4000, 90)
    indices[offsets[3]:offsets[3] + 2] = [18, 17] if offsets[4] - offsets[3] >= 2 else indices[offsets[3]:offsets[3] + 2]
    m = orbx.Matcher(max_queries=500, max_train=4000)
    g = m.knn2_csr(q, t, offsets, indices)
    o = oracle.knn2_csr(q, t, offsets, indices)
    for a, b, name in zip(g, o, ("idx1", "d1", "idx2", "d2")):
        assert np.array_equal(a, b), name
    empty = np.flatnonzero(np.diff(offsets) == 0)
    assert (g[0][empty] == -1).all() and (g[1][empty] == 256).all()
    with pytest.raises(orbx.OrbxError):
        m.knn2_csr(q, t, offsets, np.full_like(indices, 4000))  # out-of-range candidate index
    m.close()


def test_knn2_csr_full_lists_equal_bruteforce(oracle):
    import orbx
    q, t = synth.matching_set(64, 1500, seed=3)
    offsets = (np.arange(65) * 1500).astype(np.int32)
    indices = np.tile(np.arange(1500, dtype=np.int32), 64)
    m = orbx.Matcher(max_queries=64, max_train=1500)
    i1, d1, i2, d2 = m.knn2_csr(q, t, offsets, indices)
    bi, b1, b2 = m.knn2(q, t)
    assert np.array_equal(i1, bi) and np.array_equal(d1, b1) and np.array_equal(d2, b2)
    m.close()


def test_search_by_projection_vs_reference_fixture():
    """tests/golden/ref_projection.npz holds the inputs and the result of the reference's own ORBmatcher::SearchByProjection
    (src/orbmatcher.cpp:42-124, compiled unmodified).  Both GPU routes reproduce its B.m_mapPoints and nmatches: filtered
    lists -> orbm_knn2_csr -> acceptance; all candidates -> orbm_distance_csr -> the loop replayed with its exclusions; and
    orbm_search_by_projection, which also builds the frame grid and the candidate lists on the device."""
    import os
    import orbx
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_projection.npz"))
    for c, case in enumerate(g["cases"]):
        ratio = float(case[7])
        r = {k: g[f"{k}_{c}"] for k in ("mp_desc", "mp_x", "mp_y", "mp_level", "mp_radius", "b_keys", "b_desc", "b_octave", "b_uright",
                                        "b_occupied", "bounds", "offsets", "indices", "assigned")}
        ref, nref = r["assigned"], int(g[f"nmatches_{c}"])
        m = orbx.Matcher(max_queries=len(r["mp_desc"]), max_train=len(r["b_desc"]))
        # third route: the whole function on the device, grid assignment and GetFeaturesInArea included
        keys = r["b_keys"].view(orbx.KP_DTYPE).reshape(-1)
        match, asg, nm = m.search_by_projection(keys, r["b_uright"], r["b_occupied"], r["b_desc"], r["bounds"], r["mp_desc"], r["mp_x"],
                                                r["mp_y"], r["mp_level"], r["mp_radius"], ratio, 100)
        asg[(asg == -1) & (ref == -2)] = -2
        assert nm == nref and np.array_equal(asg, ref) and int((match >= 0).sum()) == nref
        # GetFeaturesInArea on the device: the reference's own lists, in its order, with every DescriptorDistance
        off, ind, dist = m.area_distances(keys, r["b_desc"], r["bounds"], r["mp_desc"], r["mp_x"], r["mp_y"], r["mp_radius"],
                                          r["mp_level"] - 1, r["mp_level"])
        assert np.array_equal(off, r["offsets"]) and np.array_equal(ind, r["indices"])
        owner = np.repeat(np.arange(len(off) - 1), np.diff(off))
        assert np.array_equal(dist, np.unpackbits(r["mp_desc"][owner] ^ r["b_desc"][ind], axis=1).sum(1))
        off2, ind2 = orbx.filter_projection_candidates(r)
        i1, d1, i2, d2 = m.knn2_csr(r["mp_desc"], r["b_desc"], off2, ind2)
        got, n = orbx.accept_projection_matches(i1, d1, i2, d2, r["b_octave"], ratio)
        got[(got == -1) & (ref == -2)] = -2
        assert n == nref and np.array_equal(got, ref)
        # second route: every candidate's distance in one launch, then the reference's loop verbatim on the host
        dist = m.distance_csr(r["mp_desc"], r["b_desc"], r["offsets"], r["indices"])
        got2, n2 = np.full(len(ref), -1, np.int32), 0
        off, ind, ur = r["offsets"], r["indices"], r["b_uright"]
        for i in range(len(off) - 1):
            best, best2, lv, lv2, bi = 256, 256, -1, -1, -1
            for k in range(off[i], off[i + 1]):
                idx = int(ind[k])
                if r["b_occupied"][idx]:
                    continue
                if ur[idx] > 0 and np.float32(abs(r["mp_x"][i] - ur[idx])) > r["mp_radius"][i]:
                    continue
                d = int(dist[k])
                if d < best:
                    best2, best, lv2, lv, bi = best, d, lv, int(r["b_octave"][idx]), idx
                elif d < best2:
                    lv2, best2 = int(r["b_octave"][idx]), d
            if best <= 100 and not (lv == lv2 and np.float32(best) > np.float32(ratio) * np.float32(best2)):
                got2[bi] = i; n2 += 1
        got2[(got2 == -1) & (ref == -2)] = -2
        assert n2 == nref and np.array_equal(got2, ref)
        m.close()


def test_search_by_projection_sequential_rule_vs_reference_fixture(oracle):
    """tests/golden/ref_projection_observed.npz: the reference's own function with observed map points and colliding twins,
    so that a map point stored at orbmatcher.cpp:121 hides its key point from later ones (:87-89).  orbm_search_by_projection
    iterates that rule to its fixpoint on the device and reproduces m_mapPoints / nmatches; the oracle agrees entry by entry."""
    import os
    import orbx
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_projection_observed.npz"))
    for c, case in enumerate(g["cases"]):
        r = {k: g[f"{k}_{c}"] for k in ("mp_desc", "mp_x", "mp_y", "mp_level", "mp_radius", "mp_observed", "b_keys", "b_desc", "b_uright",
                                        "b_occupied", "bounds", "assigned")}
        ref, nref = r["assigned"], int(g[f"nmatches_{c}"])
        keys = r["b_keys"].view(orbx.KP_DTYPE).reshape(-1)
        m = orbx.Matcher(max_queries=len(r["mp_desc"]), max_train=len(r["b_desc"]))
        args = (keys, r["b_uright"], r["b_occupied"], r["b_desc"], r["bounds"], r["mp_desc"], r["mp_x"], r["mp_y"], r["mp_level"], r["mp_radius"],
                float(case[7]), 100)
        match, asg, nm = m.search_by_projection(*args, mp_observed=r["mp_observed"])
        om, oa, on = oracle.search_by_projection(*args, mp_observed=r["mp_observed"])
        assert nm == on and np.array_equal(match, om) and np.array_equal(asg, oa)
        asg[(asg == -1) & (ref == -2)] = -2
        assert nm == nref and np.array_equal(asg, ref) and int((match >= 0).sum()) == nref
        m.close()


@pytest.mark.parametrize("n,nmp,w,h,frac", [(2000, 4000, 1241, 376, 0.8), (500, 3000, 320, 240, 1.0), (3000, 1500, 1920, 1080, 0.3), (60, 2000, 200, 120, 1.0)])
def test_search_by_projection_sequential_rule_vs_oracle(oracle, n, nmp, w, h, frac):
    """Dense collisions: many more map points than key points, most of them observed, long dependency chains (a map point
    pushed off its key point takes the next one, which pushes the next map point ...).  The device fixpoint equals the
    sequential restatement entry by entry."""
    import orbx
    rng = np.random.default_rng(n + nmp)
    keys = np.zeros(n, orbx.KP_DTYPE)
    keys["x"] = rng.uniform(0, w, n).astype(np.float32); keys["y"] = rng.uniform(0, h, n).astype(np.float32)
    keys["octave"] = rng.integers(0, 4, n)
    desc = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    ur = np.where(rng.random(n) < 0.5, keys["x"] - rng.uniform(0, 30, n).astype(np.float32), -1).astype(np.float32)
    occ = (rng.random(n) < 0.1).astype(np.uint8)
    src = rng.integers(0, n, nmp)
    mp_desc = desc[src].copy()
    flips = rng.integers(0, 256, (nmp, 32), dtype=np.uint8) & rng.integers(0, 256, (nmp, 32), dtype=np.uint8) & rng.integers(0, 256, (nmp, 32), dtype=np.uint8)
    mp_desc ^= flips                                                   # ~32 bit flips: under TH_HIGH, ties and near-ties are common
    mp_x = (keys["x"][src] + rng.uniform(-3, 3, nmp)).astype(np.float32); mp_y = (keys["y"][src] + rng.uniform(-3, 3, nmp)).astype(np.float32)
    mp_level = np.minimum(keys["octave"][src] + rng.integers(0, 2, nmp), 7).astype(np.int32)
    mp_radius = (np.float32(12.0) * np.float32(1.2) ** mp_level).astype(np.float32)
    obs = (rng.random(nmp) < frac).astype(np.uint8)
    m = orbx.Matcher(max_queries=nmp, max_train=n)
    args = (keys, ur, occ, desc, (0.0, 0.0, float(w), float(h)), mp_desc, mp_x, mp_y, mp_level, mp_radius, 0.9, 100)
    gm, ga, gn = m.search_by_projection(*args, mp_observed=obs)
    om, oa, on = oracle.search_by_projection(*args, mp_observed=obs)
    assert gn == on and np.array_equal(gm, om) and np.array_equal(ga, oa)
    assert on > 0 and on != oracle.search_by_projection(*args)[2]      # the rule changed the outcome
    m.close()


def test_distance_csr_and_host_replay_of_search_by_projection(oracle):
    """orbm_distance_csr gives every candidate's distance; replaying the reference's SearchByProjection loop
    (orbmatcher.cpp:76-124, with its 'keypoint already carries a map point' exclusion) on those distances equals the
    same loop computing DescriptorDistance itself."""
    import orbx
    rng = np.random.default_rng(5)
    nq, nt = 300, 1500
    q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    for i in range(nq):                                   # train row i = query i with a few flipped bits
        bits = np.unpackbits(q[i]); bits[rng.choice(256, int(rng.integers(0, 30)), replace=False)] ^= 1
        t[i] = np.packbits(bits)
    lists = []
    for i in range(nq):                                   # own match, the neighbours' matches (conflicts), random rows; some empty
        if i % 17 == 0:
            lists.append(np.zeros(0, np.int64)); continue
        c = np.concatenate([[i, (i + 1) % nq, (i - 1) % nq], rng.integers(0, nt, int(rng.integers(0, 30)))])
        lists.append(rng.permutation(c))
    offsets = np.concatenate([[0], np.cumsum([len(c) for c in lists])]).astype(np.int32)
    indices = np.concatenate(lists).astype(np.int32)
    octave = rng.integers(0, 8, nt)
    m = orbx.Matcher(max_queries=nq, max_train=nt)
    dist = m.distance_csr(q, t, offsets, indices)
    ref = np.array([oracle.descriptor_distance(q[i], t[indices[k]]) for i in range(nq) for k in range(offsets[i], offsets[i + 1])], np.int32)
    assert np.array_equal(dist, ref)

    def run(get_dist):
        taken, matches = np.zeros(nt, bool), []
        for i in range(nq):
            best, best2, lv, lv2, bi = 256, 256, -1, -1, -1
            for k in range(offsets[i], offsets[i + 1]):
                idx = int(indices[k])
                if taken[idx]:
                    continue
                d = get_dist(i, k)
                if d < best:
                    best2, best, lv2, lv, bi = best, d, lv, int(octave[idx]), idx
                elif d < best2:
                    lv2, best2 = int(octave[idx]), d
            if best <= 100 and not (lv == lv2 and best > 0.8 * best2):
                taken[bi] = True
                matches.append((i, bi, best))
        return matches
    a = run(lambda i, k: int(dist[k]))
    b = run(lambda i, k: oracle.descriptor_distance(q[i], t[indices[k]]))
    assert a == b and len(a) > 50
    m.close()


def _projection_case(rng, n, nmp, w, h, clustered):
    import orbx
    keys = np.zeros(n, orbx.KP_DTYPE)
    if clustered:                                             # many key points in few cells: long cell lists
        cx, cy = rng.uniform(0, w, 12), rng.uniform(0, h, 12)
        k = rng.integers(0, 12, n)
        keys["x"] = (cx[k] + rng.normal(0, 6, n)).astype(np.float32); keys["y"] = (cy[k] + rng.normal(0, 6, n)).astype(np.float32)
    else:
        keys["x"] = rng.uniform(-8, w + 8, n).astype(np.float32); keys["y"] = rng.uniform(-8, h + 8, n).astype(np.float32)   # some outside the grid
    keys["x"][::5] = np.round(keys["x"][::5])                 # integer coordinates: cell boundaries and |dx| == r ties
    keys["octave"] = rng.integers(0, 8, n)
    desc = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    uright = np.where(rng.random(n) < 0.6, keys["x"] - rng.uniform(0, 40, n), -1).astype(np.float32)
    occupied = (rng.random(n) < 0.1).astype(np.uint8)
    src = rng.integers(0, n, nmp)
    mp_desc = desc[src].copy()
    flip = rng.integers(0, 256, (nmp, 6)); on = rng.random((nmp, 6)) < 0.7
    for j in range(6):
        mp_desc[np.arange(nmp), flip[:, j] // 8] ^= (on[:, j] * (1 << (flip[:, j] % 8))).astype(np.uint8)
    mp_x = (keys["x"][src] + rng.uniform(-3, 3, nmp)).astype(np.float32); mp_y = (keys["y"][src] + rng.uniform(-3, 3, nmp)).astype(np.float32)
    mp_x[::17] = rng.uniform(-30, w + 30, len(mp_x[::17])); mp_y[::19] = rng.uniform(-30, h + 30, len(mp_y[::19]))   # off-image projections
    mp_level = np.clip(keys["octave"][src] + rng.integers(-1, 2, nmp), 0, 7).astype(np.int32)
    sf = np.float32(1.2) ** np.arange(8, dtype=np.float32)
    mp_radius = (np.where(rng.random(nmp) < 0.5, np.float32(2.5), np.float32(4.0)).astype(np.float32) * sf[mp_level]).astype(np.float32)
    mp_radius[::23] = np.float32(300.0)                       # huge windows: clamped cell ranges, many candidates
    return keys, uright, occupied, desc, (0.0, 0.0, float(w), float(h)), mp_desc, mp_x, mp_y, mp_level, mp_radius


@pytest.mark.parametrize("n,nmp,w,h,clustered", [(2000, 1500, 1241, 376, False), (3000, 2500, 640, 480, True), (1, 5, 100, 100, False),
                                                  (5000, 33, 1920, 1080, False), (12000, 700, 1920, 1080, True)])   # > 8192: cells beyond the on-chip table
def test_search_by_projection_vs_oracle(oracle, n, nmp, w, h, clustered):
    """orbm_search_by_projection against the oracle restatement (itself pinned to the reference's orbmatcher.cpp /
    orbframe.cpp) on inputs the fixture does not reach: key points outside the grid, cell-boundary coordinates, long cell
    lists, clamped windows, off-image projections."""
    import orbx
    rng = np.random.default_rng(n + nmp)
    case = _projection_case(rng, n, nmp, w, h, clustered)
    m = orbx.Matcher(max_queries=16, max_train=16)
    for ratio, th, occ in ((0.8, 100, True), (0.6, 50, False)):
        args = list(case)
        if not occ:
            args[2] = None
        gm, ga, gn = m.search_by_projection(*args, nnratio=ratio, th_high=th)
        om, oa, on = oracle.search_by_projection(*args, nnratio=ratio, th_high=th)
        assert gn == on and np.array_equal(gm, om) and np.array_equal(ga, oa)
    if n > 1:
        assert on > 0
    # an empty frame is not an error
    e = m.search_by_projection(case[0][:0], case[1][:0], None, case[3][:0], case[4], *case[5:])
    assert e[2] == 0 and (e[0] == -1).all()
    with pytest.raises(orbx.OrbxError):
        m.search_by_projection(*case[:4], (0.0, 0.0, 0.0, 10.0), *case[5:])
    m.close()


@pytest.mark.parametrize("n,nq,w,h,clustered", [(2000, 1500, 1241, 376, False), (3000, 2500, 640, 480, True), (7, 3000, 320, 240, False)])
def test_area_distances_vs_oracle(oracle, n, nq, w, h, clustered):
    """orbm_area_distances (GetFeaturesInArea + DescriptorDistance for many windows) against the oracle restatement:
    level windows with the reference's -1 conventions, huge and off-image windows, too-small output arrays."""
    import orbx
    rng = np.random.default_rng(7 * n + nq)
    keys, _, _, desc, bounds, q_desc, q_x, q_y, q_level, q_r = _projection_case(rng, n, nq, w, h, clustered)
    kind = rng.integers(0, 4, nq)
    l0 = np.where(kind == 0, -1, np.where(kind == 1, q_level, q_level - 1)).astype(np.int32)   # no check / [l, l] / [l-1, l]
    l1 = np.where(kind == 0, -1, q_level).astype(np.int32)
    l0[kind == 3] = 2; l1[kind == 3] = -1                                                       # minLevel only (:337, :352-364)
    m = orbx.Matcher(max_queries=16, max_train=16)
    go, gi, gd = m.area_distances(keys, desc, bounds, q_desc, q_x, q_y, q_r, l0, l1, cap=64)      # grows through ORBX_ERR_CAPACITY
    oo, oi, od = oracle.area_distances(keys, desc, bounds, q_desc, q_x, q_y, q_r, l0, l1, cap=1 << 22)
    assert np.array_equal(go, oo) and np.array_equal(gi, oi) and np.array_equal(gd, od) and len(oi) > 0
    go2, gi2, gd2 = m.area_distances(keys, desc, bounds, None, q_x, q_y, q_r, l0, l1)            # lists only
    assert gd2 is None and np.array_equal(go2, oo) and np.array_equal(gi2, oi)
    m.close()
