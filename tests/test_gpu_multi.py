"""GPU: the multi-GPU entry points of the library itself (include/orbx.h: orbx_multi_*, orbm_multi_*, orbm_window_* /
orbm_knn2_sharded).  A device ordinal may be listed more than once, so the sharding, the host threads, the peer windows and
the flag protocol are exercised on a one-GPU box too; with two or more GPUs visible the same tests also run across them."""
import multiprocessing as mp
import os
import sys

import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu


def _device_sets():
    import torch
    sets = [[0], [0, 0], [0, 0, 0]]
    n = torch.cuda.device_count()
    if n >= 2:
        sets.append(list(range(min(n, 8))))
    return sets


def test_multi_extractor_equals_single_gpu_calls(oracle):
    """orbx_multi_extract_batch(_async): frames sharded over the device slots land in the caller's arrays exactly where the
    single-GPU call puts them -- every batch size from fewer frames than slots upwards, two asynchronous calls in flight."""
    import orbx
    w, h = 640, 360
    ref = orbx.Extractor(nfeatures=1000, nlevels=6, max_width=w, max_height=h, max_batch=21)
    oex = oracle.Extractor(nfeatures=1000, nlevels=6)
    for devices in _device_sets():
        mx = orbx.MultiExtractor(devices, nfeatures=1000, nlevels=6, max_width=w, max_height=h, max_batch=21)
        assert mx.n_devices == len(devices) and mx.max_keypoints == ref.max_keypoints
        total = 0
        for s in range(len(devices)):
            a, n = mx.frame_range(21, s)
            assert a == total
            total += n
        assert total == 21
        for k, b in enumerate((1, 2, 7, 21)):
            imgs = synth.frames(70 + k, w, h, b)
            kr, dr, cr = ref.extract_batch(imgs)
            km, dm, cm = mx.extract_batch(imgs)
            assert np.array_equal(cm, cr), (devices, b)
            for f in range(b):
                n = int(cr[f])
                assert km[f][:n].tobytes() == kr[f][:n].tobytes() and np.array_equal(dm[f][:n], dr[f][:n]), (devices, b, f)
        # two calls in flight, results of both intact
        cap = mx.max_keypoints
        sets = [synth.frames(80 + k, w, h, 21) for k in range(3)]
        outs = [(np.zeros((21, cap), orbx.KP_DTYPE), np.zeros((21, cap, 32), np.uint8), np.zeros(21, np.int32)) for _ in sets]
        ptrs = [orbx.Extractor.frame_pointers([np.ascontiguousarray(i) for i in fr]) for fr in sets]
        t = [mx.extract_batch_async(ptrs[0], 21, w, h, w, outs[0])]
        t.append(mx.extract_batch_async(ptrs[1], 21, w, h, w, outs[1]))
        mx.wait(t[0])
        t.append(mx.extract_batch_async(ptrs[2], 21, w, h, w, outs[2]))
        mx.wait(t[1]); mx.wait(t[2])
        with pytest.raises(orbx.OrbxError):
            mx.wait(t[2])
        for k, fr in enumerate(sets):
            kr, dr, cr = ref.extract_batch(fr)
            assert np.array_equal(outs[k][2], cr)
            for f in (0, 10, 20):
                n = int(cr[f])
                assert outs[k][0][f][:n].tobytes() == kr[f][:n].tobytes() and np.array_equal(outs[k][1][f][:n], dr[f][:n])
        okps, odesc = oex.extract(sets[2][20])
        n = int(outs[2][2][20])
        assert n == len(okps) and outs[2][0][20][:n].tobytes() == okps.tobytes() and np.array_equal(outs[2][1][20][:n], odesc)
        mx.close()
    ref.close()


@pytest.mark.parametrize("nq,nt", [(2000, 100000), (250, 30000), (5, 999), (2, 300), (129, 257), (5, 100000)])   # the last: 37 fold groups
def test_sharded_knn2_in_one_process(oracle, nq, nt):
    """orbm_multi_knn2: every device slot matches its block of queries in one launch and stores its records into every
    slot's window (peer stores + flags, no collective call); the host reads the whole result from slot 0.  Equal to the
    single-GPU kernel pair on every query and to the oracle on a sample; query counts below the number of slots leave a slot
    without queries (it still publishes its flag)."""
    import orbx
    q, t = synth.matching_set(nq, nt, seed=77 + nq)
    m1 = orbx.Matcher(max_queries=nq, max_train=nt)
    i1, a1, b1 = m1.knn2(q, t)
    m1.close()
    pick = np.linspace(0, nq - 1, min(nq, 40)).astype(int)
    oi, o1, o2 = oracle.knn2(q[pick], t)
    assert np.array_equal(i1[pick], oi) and np.array_equal(a1[pick], o1) and np.array_equal(b1[pick], o2)
    for devices in _device_sets():
        mm = orbx.MultiMatcher(devices, max_queries=nq, max_train=nt)
        mm.set_train(t)
        for rep in range(3):                      # consecutive calls: epochs, double-buffered windows, counters reset
            qi = q if rep != 1 else q[::-1].copy()
            idx, d1, d2 = mm.knn2(qi)
            ri, ra, rb = (i1, a1, b1) if rep != 1 else (i1[::-1], a1[::-1], b1[::-1])
            assert np.array_equal(idx, ri) and np.array_equal(d1, ra) and np.array_equal(d2, rb), (devices, rep)
        if nq >= 8:
            idx, d1, d2 = mm.knn2(q[:nq // 2 + 1])   # a smaller query block on the same windows
            assert np.array_equal(idx, i1[:nq // 2 + 1]) and np.array_equal(d2, b1[:nq // 2 + 1])
        mm.close()


def _ipc_rank(rank, world, conn, nq, nt, repo_paths):
    """one process = one rank: own matcher + window, peers' windows mapped through CUDA IPC handles"""
    sys.path[:0] = repo_paths
    import torch
    import orbx
    import synth as S
    dev = rank % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    q, t = S.matching_set(nq, nt, seed=5)
    m = orbx.Matcher(max_queries=nq, max_train=nt, device=dev)
    handle = m.window_create(nq, world, rank)
    conn.send(handle)
    handles = conn.recv()
    for r in range(world):
        if r != rank:
            m.window_attach_ipc(r, handles[r])
    block = (nq + world - 1) // world
    lo, hi = min(rank * block, nq), min(rank * block + block, nq)
    dq = torch.from_numpy(np.ascontiguousarray(q[lo:hi]) if hi > lo else np.zeros((1, 32), np.uint8)).cuda()
    dt = torch.from_numpy(t).cuda()
    ok = True
    for rep in range(4):
        m.knn2_sharded(dq.data_ptr(), hi - lo, lo, dt.data_ptr(), nt)
        rec = m.window_fetch(nq)
        m1 = orbx.Matcher(max_queries=nq, max_train=nt, device=dev)
        i1, a1, b1 = m1.knn2(q, t)
        m1.close()
        ok = ok and np.array_equal(rec[:, 0], i1) and np.array_equal(rec[:, 1], a1) and np.array_equal(rec[:, 2], b1)
    conn.send(bool(ok))
    conn.recv()                                    # keep the window alive until every rank has finished reading
    m.close()


@pytest.mark.parametrize("world,nq,nt", [(2, 2000, 50000), (3, 50, 4000)])
def test_sharded_knn2_across_processes_ipc(world, nq, nt):
    """One process per rank (the torchrun shape): windows exchanged as CUDA IPC handles, every rank's kernel stores its records
    into the other processes' windows; each rank ends up with the complete, correct result.  Ranks share GPU 0 on a one-GPU box."""
    ctx = mp.get_context("spawn")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    paths = [os.path.join(root, "opendlv-perception-vision-orbslam2_b200"), root]
    pipes, procs = [], []
    for r in range(world):
        a, b = ctx.Pipe()
        p = ctx.Process(target=_ipc_rank, args=(r, world, b, nq, nt, paths))
        p.start()
        pipes.append(a); procs.append(p)
    try:
        handles = []
        for a in pipes:
            assert a.poll(120), "a rank did not come up"
            handles.append(a.recv())
        for a in pipes:
            a.send(handles)
        results = []
        for a in pipes:
            assert a.poll(120), "a rank did not finish"
            results.append(a.recv())
        for a in pipes:
            a.send(True)
        assert all(results), results
    finally:
        for p in procs:
            p.join(timeout=60)
            if p.is_alive():
                p.kill()
