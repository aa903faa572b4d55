"""CPU checks of the oracle restatements for the 'next' rows N2 / N4 against independent numpy formulations."""
import numpy as np


def _ham(a, b):
    return int(np.unpackbits(a ^ b).sum())


def test_distinctive_oracle_matches_numpy(oracle):
    rng = np.random.default_rng(0)
    desc = rng.integers(0, 256, (200, 32), dtype=np.uint8)
    offsets, indices = [0], []
    for n in (1, 2, 3, 4, 9, 16, 0, 31):
        indices += rng.integers(0, 200, n).tolist()
        offsets.append(len(indices))
    best, med = oracle.distinctive(desc, offsets, indices)
    for p in range(len(offsets) - 1):
        ix = indices[offsets[p]:offsets[p + 1]]
        if not ix:
            assert best[p] == -1
            continue
        n = len(ix)
        D = np.array([[_ham(desc[i], desc[j]) for j in ix] for i in ix])
        meds = np.sort(D, axis=1)[:, int(0.5 * (n - 1))]
        assert best[p] == int(np.argmin(meds)) and med[p] == int(meds.min())


def test_voc_transform_oracle_matches_numpy(oracle):
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "opendlv-perception-vision-orbslam2_b200"))
    rng = np.random.default_rng(1)
    # k = 4, L = 3 tree in array form
    k, L = 4, 3
    kids, n, frontier = {}, 1, [0]
    for _ in range(L):
        nxt = []
        for v in frontier:
            kids[v] = list(range(n, n + k)); n += k; nxt += kids[v]
        frontier = nxt
    child_off, child_ids = [0], []
    for v in range(n):
        child_ids += kids.get(v, []); child_off.append(len(child_ids))
    node_desc = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    word_id = np.full(n, -1, np.int32)
    leaves = [v for v in range(n) if v not in kids]
    word_id[leaves] = np.arange(len(leaves))
    feat = rng.integers(0, 256, (64, 32), dtype=np.uint8)
    for levels_up in (0, 1, 2, 5):
        word, node = oracle.voc_transform(child_off, child_ids, node_desc, word_id, L, levels_up, feat)
        for i, f in enumerate(feat):
            v, lev, nid = 0, 0, 0
            while v in kids:
                lev += 1
                ds = [_ham(f, node_desc[c]) for c in kids[v]]
                v = kids[v][int(np.argmin(ds))]
                if lev == L - levels_up:
                    nid = v
            assert word[i] == word_id[v] and node[i] == nid
