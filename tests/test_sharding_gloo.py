"""CPU, world_size 2 over gloo: the N>1 partitioning used by bench.py (frames for extraction, query
blocks + one all_gather for matching).  The oracle stands in for the per-rank compute here; the
sharded answers must equal the unsharded ones."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q_out):
    for p in (os.path.join(ROOT, "opendlv-perception-vision-orbslam2_b200"), os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    import orb_oracle_py as oracle
    import shard
    import synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ---- matching: query blocks, train replicated, one all_gather of 16-byte records
        nq, nt = 101, 700                         # ragged: 101 is not a multiple of 2
        q, t = synth.matching_set(nq, nt, seed=31)
        b = shard.query_block(nq, world)
        lo, hi = shard.query_range(nq, rank, world)
        local = torch.zeros((b, 4), dtype=torch.int32)
        idx, d1, d2 = oracle.knn2(q[lo:hi], t)
        local[:hi - lo, 0] = torch.from_numpy(idx); local[:hi - lo, 1] = torch.from_numpy(d1); local[:hi - lo, 2] = torch.from_numpy(d2)
        allr = shard.gather_match_records(local, nq).numpy()
        gi, g1, g2 = oracle.knn2(q, t)
        ok_match = np.array_equal(allr[:, 0], gi) and np.array_equal(allr[:, 1], g1) and np.array_equal(allr[:, 2], g2)
        # ---- extraction: contiguous frame blocks, nothing exchanged but the counts
        frames = synth.frames(7, 320, 200, 5)     # 5 frames over 2 ranks: 2 + 3
        flo, fhi = shard.frame_range(len(frames), rank, world)
        ex = oracle.Extractor(300, 1.2, 4)
        mine = [len(ex.extract(f)[0]) for f in frames[flo:fhi]]
        counts = shard.gather_counts(torch.tensor(mine, dtype=torch.int32)).tolist()
        full = [len(ex.extract(f)[0]) for f in frames]
        q_out.put((rank, ok_match, counts == full, (flo, fhi)))
    finally:
        dist.destroy_process_group()


def test_world_size_2_partitioning_matches_unsharded():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q_out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q_out)) for r in range(world)]
    [p.start() for p in procs]
    res = [q_out.get(timeout=240) for _ in range(world)]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    for rank, ok_match, ok_counts, rng in res:
        assert ok_match, f"rank {rank}: gathered matches differ from the unsharded answer"
        assert ok_counts, f"rank {rank}: gathered counts differ"
    assert sorted(r[3] for r in res) == [(0, 2), (2, 5)]


def test_ranges_cover_everything_once():
    sys.path.insert(0, os.path.join(ROOT, "opendlv-perception-vision-orbslam2_b200"))
    import shard
    for n in (1, 5, 64, 2000, 2001):
        for world in (1, 2, 4, 8):
            fr = [shard.frame_range(n, r, world) for r in range(world)]
            assert fr[0][0] == 0 and fr[-1][1] == n and all(fr[i][1] == fr[i + 1][0] for i in range(world - 1))
            qr = [shard.query_range(n, r, world) for r in range(world)]
            assert sum(h - l for l, h in qr) == n and all(h - l <= shard.query_block(n, world) for l, h in qr)
