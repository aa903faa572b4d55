"""GPU parity of the 'next' rows N2 (bag-of-words descent) and N4 (distinctive descriptor) against the oracle
restatements of orbvocabulary.cpp:203-242 and orbmappoint.cpp:314-383.  Integer work: results must be identical."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _lists(rng, n_points, n_desc, sizes):
    offsets, indices = [0], []
    for p in range(n_points):
        n = int(sizes[p % len(sizes)])
        indices += rng.integers(0, n_desc, n).tolist()
        offsets.append(len(indices))
    return np.array(offsets, np.int32), np.array(indices, np.int32)


def test_distinctive_descriptor_vs_oracle(oracle):
    import orbx
    rng = np.random.default_rng(3)
    n_desc = 5000
    desc = rng.integers(0, 256, (n_desc, 32), dtype=np.uint8)
    # clusters of near-duplicates so that medians tie and the first-row rule matters
    for c in range(0, n_desc, 50):
        base = desc[c].copy()
        for j in range(1, 20):
            d = np.unpackbits(base); d[rng.choice(256, int(rng.integers(0, 12)), replace=False)] ^= 1
            desc[c + j] = np.packbits(d)
    sizes = [1, 2, 3, 0, 5, 8, 13, 21, 32, 33, 47, 64, 100, 7, 4]
    offsets, indices = _lists(rng, 600, n_desc, sizes)
    # lists drawn from one cluster (many equal distances)
    o2, i2 = [0], []
    for p in range(100):
        c = 50 * int(rng.integers(0, n_desc // 50))
        i2 += (c + rng.integers(0, 20, int(rng.integers(2, 25)))).tolist()
        o2.append(len(i2))
    m = orbx.Matcher(max_queries=16, max_train=16)
    for off, ind in ((offsets, indices), (np.array(o2, np.int32), np.array(i2, np.int32))):
        gb, gm = m.distinctive(desc, off, ind)
        ob, om = oracle.distinctive(desc, off, ind)
        assert np.array_equal(gb, ob), f"first differing point {np.flatnonzero(gb != ob)[:5]}"
        assert np.array_equal(gm, om)
    # one long list
    off, ind = np.array([0, 300], np.int32), rng.integers(0, n_desc, 300).astype(np.int32)
    assert m.distinctive(desc, off, ind)[0].tolist() == oracle.distinctive(desc, off, ind)[0].tolist()
    with pytest.raises(orbx.OrbxError):
        m.distinctive(desc, np.array([0, 800], np.int32), rng.integers(0, n_desc, 800).astype(np.int32))
    m.close()


def test_distinctive_descriptor_vs_reference_fixture():
    """tests/golden/ref_mappoint.npz: descriptors kept by the reference's own OrbMapPoint::ComputeDistinctiveDescriptors."""
    import os
    import orbx
    import synth
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_mappoint.npz"))
    off2, ind2 = synth.drop_bad_observations(g["offsets"], g["indices"], g["bad"])
    m = orbx.Matcher(max_queries=16, max_train=16)
    best, _ = m.distinctive(g["desc"], off2, ind2)
    n = np.diff(off2)
    assert ((best == -1) == (n == 0)).all() and ((g["has"] == 0) == (n == 0)).all()
    ok = n > 0
    assert np.array_equal(g["desc"][ind2[off2[:-1][ok] + best[ok]]], g["out"][ok])
    m.close()


@pytest.mark.parametrize("k,L,levels_up", [(10, 3, 1), (10, 4, 2), (3, 5, 4), (10, 2, 4), (32, 2, 1)])
def test_vocabulary_descent_vs_oracle(oracle, k, L, levels_up):
    import orbx
    voc = orbx.random_vocabulary(k, L, seed=k * 10 + L)
    child_off, child_ids, node_desc, word_id, weight, _ = voc
    # duplicate child descriptors: the first child with the least distance must win
    node_desc[child_ids[child_off[0] + 1]] = node_desc[child_ids[child_off[0]]]
    rng = np.random.default_rng(11)
    feat = rng.integers(0, 256, (3001, 32), dtype=np.uint8)
    feat[:50] = node_desc[rng.integers(1, len(node_desc), 50)]           # exact hits
    v = orbx.Vocabulary(child_off, child_ids, node_desc, word_id, weight, L)
    gw, gwt, gn = v.transform(feat, levels_up)
    ow, on = oracle.voc_transform(child_off, child_ids, node_desc, word_id, L, levels_up, feat)
    assert np.array_equal(gw, ow) and np.array_equal(gn, on)
    assert np.array_equal(gwt, weight[np.flatnonzero(word_id >= 0)][ow])
    v.close()


def test_vocabulary_on_device_resident_descriptors(oracle):
    """transform5 straight on the extractor's descriptors in HBM (OrbFrame::ComputeBoW without a host round trip)."""
    import torch
    import orbx
    import synth
    w, h = 640, 360
    ex = orbx.Extractor(nfeatures=1000, nlevels=6, max_width=w, max_height=h, max_batch=1)
    kps, desc, counts = ex.extract_batch([synth.scene_s1(w, h, 77)])
    n = int(counts[0])
    child_off, child_ids, node_desc, word_id, weight, L = orbx.random_vocabulary(10, 3, seed=5)
    v = orbx.Vocabulary(child_off, child_ids, node_desc, word_id, weight, L)
    _, d_desc, _, stride = ex.device_results()
    out = torch.zeros((n, 2), dtype=torch.int32, device="cuda")
    v.transform_device(d_desc, 32, n, 1, out.data_ptr())
    torch.cuda.synchronize()
    ow, on = oracle.voc_transform(child_off, child_ids, node_desc, word_id, L, 1, desc[0][:n])
    got = out.cpu().numpy()
    assert np.array_equal(got[:, 0], ow) and np.array_equal(got[:, 1], on)
    v.close(); ex.close()


def test_vocabulary_descent_vs_reference_fixture():
    """GPU against the outputs of the reference's own OrbVocabulary frozen in tests/golden/ref_vocabulary.npz."""
    import os
    import orbx
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_vocabulary.npz"))
    L = int(g["L"])
    v = orbx.Vocabulary(g["child_off"], g["child_ids"], g["node_desc"], g["word_id"], g["weight"], L)
    for lu in (0, 1, 2, 4):
        w, wt, n = v.transform(g["feat"], lu)
        kept = wt > 0
        assert np.array_equal(w[kept], g[f"word_lu{lu}"][kept]) and np.array_equal(n[kept], g[f"node_lu{lu}"][kept])
        assert (g[f"word_lu{lu}"][~kept] == -1).all()
    v.close()
