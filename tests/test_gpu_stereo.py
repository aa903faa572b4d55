"""GPU parity of orbx_stereo_match (OrbFrame::ComputeStereoMatches, orbframe.cpp:511-705) against the
oracle restatement and against tests/golden/ref_stereo.npz, the output of the reference's own src/orbframe.cpp
(the restatement itself is pinned to that translation unit in tests/test_oracle_vs_ref.py)."""
import os

import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu


def _oracle_pair(oracle, l, r, nf, nl, mbf, mb):
    eL = oracle.Extractor(nf, 1.2, nl); eR = oracle.Extractor(nf, 1.2, nl)
    kl, dl = eL.extract(l); kr, dr = eR.extract(r)
    return oracle.stereo_matches(eL, eR, kl, dl, kr, dr, mbf, mb), len(kl)


@pytest.mark.parametrize("mb", [0.0, 0.5372])          # 0: the reference's first-frame value (maxD = +inf)
def test_two_handles_like_orbframe(oracle, mb):
    import orbx
    w, h, nf, nl, mbf = 1241, 376, 2000, 8, 386.1448
    for seed in (2000, 2001):
        l, r = synth.stereo_pair(w, h, seed)
        exL = orbx.Extractor(nf, 1.2, nl, max_width=w, max_height=h)
        exR = orbx.Extractor(nf, 1.2, nl, max_width=w, max_height=h)
        exL.extract(l); exR.extract(r)
        u, d, nm = orbx.stereo_match(exL, 0, exR, 0, mbf, mb)
        (ou, od, on), n_left = _oracle_pair(oracle, l, r, nf, nl, mbf, mb)
        assert len(u) == n_left and nm == on and on > 50
        assert np.array_equal(u.view(np.uint32), ou.view(np.uint32)), f"uRight differs at {np.flatnonzero(u != ou)[:5]}"
        assert np.array_equal(d.view(np.uint32), od.view(np.uint32))
        exL.close(); exR.close()


def test_stereo_vs_reference_fixture():
    """Images from the seed -> GPU extraction (left / right) -> FilterKeyPoints in HBM -> GPU stereo matching -> grid, against
    the key points, descriptors, mvuRight / m_depths and m_grid the reference's own OrbFrame produced for the same pair
    (scripts/gen_golden_stereo.py; the third case carries a bounding box)."""
    import orbx
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_stereo.npz"))
    for c, (w, h, seed, nf, nl, mbf, mb) in enumerate(g["cases"]):
        w, h, seed, nf, nl = int(w), int(h), int(seed), int(nf), int(nl)
        left, right = synth.stereo_pair(w, h, seed)
        ex = orbx.Extractor(nf, 1.2, nl, max_width=w, max_height=h, max_batch=2)
        kps, desc, cnt = ex.extract_batch([left, right])
        n_before = cnt.copy()
        ex.filter_keypoints(g["boxes"][c], 0, 2)
        kps, desc, cnt = ex.fetch_results(2)
        kl, dl, kr, dr = kps[0, :cnt[0]], desc[0, :cnt[0]], kps[1, :cnt[1]], desc[1, :cnt[1]]
        assert kl.tobytes() == g[f"kl_{c}"].tobytes() and kr.tobytes() == g[f"kr_{c}"].tobytes()
        assert np.array_equal(dl, g[f"dl_{c}"]) and np.array_equal(dr, g[f"dr_{c}"])
        assert (cnt < n_before).all() if g["boxes"][c][1] > 2 else (cnt == n_before).all()
        u, d, nm = orbx.stereo_match(ex, 0, ex, 1, float(mbf), float(mb))
        assert int((u >= 0).sum()) == int((g[f"uRight_{c}"] >= 0).sum()) > 50 and nm >= int((u >= 0).sum())   # nm counts before the median filter
        assert np.array_equal(u.view(np.uint32), g[f"uRight_{c}"].view(np.uint32))
        assert np.array_equal(d.view(np.uint32), g[f"depth_{c}"].view(np.uint32))
        m = orbx.Matcher(16, 16)
        start, items = m.assign_grid(kl, (0.0, 0.0, float(w), float(h)))
        assert np.array_equal(start, g[f"grid_start_{c}"]) and np.array_equal(items, g[f"grid_items_{c}"])
        m.close(); ex.close()


def test_filter_keypoints_vs_oracle(oracle):
    """orbx_filter_keypoints on frames of a larger batch: more than one 1024-entry chunk per frame, a box that removes
    everything, a box that removes nothing, and the reference's 'box[1] <= 2 means no box' rule."""
    import orbx
    w, h, nf = 1241, 376, 3000
    frames = synth.stereo_batch(4, w, h, 2)
    ex = orbx.Extractor(nf, 1.2, 8, max_width=w, max_height=h, max_batch=4)
    k0, d0, c0 = ex.extract_batch(frames)
    k0, d0, c0 = k0.copy(), d0.copy(), c0.copy()
    assert c0.min() > 2048
    boxes = [(300.0, 900.0, 80.0, 300.0), (-10.0, 5000.0, -10.0, 5000.0), (600.0, 601.0, 100.0, 101.0), (0.0, 2.0, 0.0, 1000.0)]
    for f, box in enumerate(boxes):
        ex.filter_keypoints(box, f, 1)
    k1, d1, c1 = ex.fetch_results(4)
    for f, box in enumerate(boxes):
        ok, od = oracle.filter_keypoints(k0[f, :c0[f]], d0[f, :c0[f]], box)
        assert c1[f] == len(ok) and k1[f, :c1[f]].tobytes() == ok.tobytes() and np.array_equal(d1[f, :c1[f]], od)
    assert c1[1] == 0 and c1[3] == c0[3] and 0 < c1[0] < c0[0]
    with pytest.raises(orbx.OrbxError):
        ex.filter_keypoints(boxes[0], 3, 2)
    ex.close()


def test_one_handle_batch_of_pairs(oracle):
    import orbx
    w, h, nf, nl, mbf = 640, 360, 1000, 6, 300.0
    frames = synth.stereo_batch(9, w, h, 2)              # L0 R0 L1 R1
    ex = orbx.Extractor(nf, 1.2, nl, max_width=w, max_height=h, max_batch=4)
    ex.extract_batch(frames)
    for p in range(2):
        u, d, nm = orbx.stereo_match(ex, 2 * p, ex, 2 * p + 1, mbf, 0.0)
        (ou, od, on), _ = _oracle_pair(oracle, frames[2 * p], frames[2 * p + 1], nf, nl, mbf, 0.0)
        assert nm == on and np.array_equal(u.view(np.uint32), ou.view(np.uint32)) and np.array_equal(d.view(np.uint32), od.view(np.uint32))
    ex.close()


def test_no_matches_is_not_an_error(oracle):
    import orbx
    w, h = 640, 360
    ex = orbx.Extractor(500, 1.2, 4, max_width=w, max_height=h, max_batch=2)
    ex.extract_batch([synth.scene_s1(w, h, 1), synth.scene_s3(w, h, "const")])   # right image has no keypoints
    u, d, nm = orbx.stereo_match(ex, 0, ex, 1, 300.0, 0.0)
    assert nm == 0 and (u == -1).all() and (d == -1).all()
    ex.close()


def test_batch_of_pairs_in_one_launch(oracle):
    """orbx_stereo_match_batch: all pairs of a batch at once, equal to the per-pair call and to the oracle."""
    import orbx
    w, h, nf, nl, mbf = 640, 360, 1000, 6, 300.0
    frames = synth.stereo_batch(12, w, h, 5)             # L0 R0 ... L4 R4
    ex = orbx.Extractor(nf, 1.2, nl, max_width=w, max_height=h, max_batch=10)
    ex.extract_batch(frames)
    u, d, n_left, n_match = orbx.stereo_match_batch(ex, 5, 0, 1, 2, mbf, 0.0)
    for p in range(5):
        u1, d1, nm1 = orbx.stereo_match(ex, 2 * p, ex, 2 * p + 1, mbf, 0.0)
        n = int(n_left[p])
        assert n == len(u1) and int(n_match[p]) == nm1
        assert np.array_equal(u[p, :n].view(np.uint32), u1.view(np.uint32)) and np.array_equal(d[p, :n].view(np.uint32), d1.view(np.uint32))
    for p in (0, 4):
        (ou, od, on), _ = _oracle_pair(oracle, frames[2 * p], frames[2 * p + 1], nf, nl, mbf, 0.0)
        n = int(n_left[p])
        assert int(n_match[p]) == on and np.array_equal(u[p, :n].view(np.uint32), ou.view(np.uint32)) and np.array_equal(d[p, :n].view(np.uint32), od.view(np.uint32))
    with pytest.raises(orbx.OrbxError):
        orbx.stereo_match_batch(ex, 6, 0, 1, 2, mbf, 0.0)      # sixth pair is outside the batch
    ex.close()


def test_stereo_with_more_than_8192_keypoints_per_frame(oracle):
    """nfeatures = 9000: the median filter's sort buffer (capacity rounded up to 16384 keys) needs 64 KB of dynamic shared
    memory, i.e. the opt-in."""
    import orbx
    w, h, nf, nl = 1920, 1080, 9000, 8
    left, right = synth.stereo_pair(w, h, 4)
    ex = orbx.Extractor(nf, 1.2, nl, max_width=w, max_height=h, max_batch=2)
    kps, desc, cnt = ex.extract_batch([left, right])
    assert ex.max_keypoints > 8192 and cnt[0] > 2000      # the sort buffer is sized by the capacity, not by the count
    u, d, nm = orbx.stereo_match(ex, 0, ex, 1, 500.0, 0.5)
    (ou, od, on), n_left = _oracle_pair(oracle, left, right, nf, nl, 500.0, 0.5)
    assert len(u) == n_left and nm == on and on > 100
    assert np.array_equal(u.view(np.uint32), ou.view(np.uint32)) and np.array_equal(d.view(np.uint32), od.view(np.uint32))
    ex.close()


def test_stereo_on_a_pipe_slot(oracle):
    """Stereo batches submitted through an orbx_pipe: the extractor that holds a ticket's batch serves ComputeStereoMatches of
    exactly that batch while the next submission is already running on another slot."""
    import torch
    import orbx
    w, h, nf, nl, mbf, mb = 752, 480, 1200, 6, 386.1448, 0.5372
    pairs = [synth.stereo_pair(w, h, 2100 + k) for k in range(3)]
    pipe = orbx.Pipe(depth=2, nfeatures=nf, nlevels=nl, max_width=w, max_height=h, max_batch=2)
    dev = [torch.from_numpy(np.stack([l, r])).cuda() for (l, r) in pairs]
    torch.cuda.synchronize()
    tk = [pipe.submit(dev[0].data_ptr(), h * w, w, 2, w, h)]
    for k in range(3):
        if k + 1 < 3:
            tk.append(pipe.submit(dev[k + 1].data_ptr(), h * w, w, 2, w, h))     # the next pair is in flight during the match
        pipe.join(tk[k])
        u, d, n_left, n_match = orbx.stereo_match_batch(pipe.extractor(tk[k]), 1, 0, 1, 2, mbf, mb)
        (ou, od, on), n_l = _oracle_pair(oracle, pairs[k][0], pairs[k][1], nf, nl, mbf, mb)
        assert int(n_left[0]) == n_l and int(n_match[0]) == on and on > 20
        assert np.array_equal(u[0][:n_l].view(np.uint32), ou.view(np.uint32)) and np.array_equal(d[0][:n_l].view(np.uint32), od.view(np.uint32))
    pipe.close()
