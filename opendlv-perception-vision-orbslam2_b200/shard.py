"""Partitioning of the path over the GPUs of one box (one process per GPU, torch.distributed).

Extraction: frames are independent units (the reference already runs left/right images in separate
threads with separate instances, orbframe.cpp:73-76) -> rank r takes a contiguous block of frames,
nothing is exchanged on the data path.
Matching: query-block sharded -- every rank holds the whole train set (3.2 MB for 100k map points),
matches its block of queries, and ONE all_gather of the 16-byte result records {idx, d1, d2, pad}
delivers all results everywhere (32 KB for 2000 queries; latency-bound, so the records are written
by the kernel straight into the gather's send buffer).
"""
import torch
import torch.distributed as dist


def frame_range(n_frames, rank, world):
    """Contiguous block of frames for `rank`: sizes differ by at most one."""
    lo = n_frames * rank // world
    hi = n_frames * (rank + 1) // world
    return lo, hi


def query_block(n_queries, world):
    """Queries per rank (equal blocks, the last one padded) -- all_gather needs equal sizes."""
    return (n_queries + world - 1) // world


def query_range(n_queries, rank, world):
    b = query_block(n_queries, world)
    lo = min(rank * b, n_queries)
    return lo, min(lo + b, n_queries)


def gather_match_records(local, n_queries, out=None, group=None):
    """local: int32 [query_block, 4] records of this rank (rows past its range are padding).
    Returns int32 [n_queries, 4] with every rank's results, in query order."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local[:n_queries]
    if out is None:
        out = torch.empty((world * local.shape[0], local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local, group=group)
    return out[:n_queries]


def gather_counts(local_counts, group=None):
    """Per-frame keypoint counts of every rank (frames in global order); the only thing an
    extraction job may want to exchange -- 4 bytes per frame."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local_counts
    sizes = [torch.zeros(1, dtype=torch.int64, device=local_counts.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([local_counts.numel()], dtype=torch.int64, device=local_counts.device), group=group)
    m = int(max(s.item() for s in sizes))
    pad = torch.zeros(m, dtype=local_counts.dtype, device=local_counts.device)
    pad[:local_counts.numel()] = local_counts
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:int(s.item())] for p, s in zip(parts, sizes)])
