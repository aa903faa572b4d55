"""Deterministic synthetic inputs for the ORB front-end benchmarks and parity tests.

Generators follow SURVEY.md §8(d): seed = 1000*config_id + frame_index, PCG64.
  S1  corner-rich scene with flat areas (rectangles + mild blur + noise)
  S2  i.i.d. uniform u8 noise (stress: ~1 % of pixels are FAST corners)
  S3  degenerate: constant image / a two-level step edge
  C5  Hamming matching set with planted near-duplicates
"""
import numpy as np


def _blur3(a, sigma=0.7):
    k = np.exp(-np.arange(-1, 2, dtype=np.float64) ** 2 / (2 * sigma * sigma))
    k = (k / k.sum()).astype(np.float32)
    p = np.pad(a, 1, mode="reflect")
    h = k[0] * p[:, :-2] + k[1] * p[:, 1:-1] + k[2] * p[:, 2:]
    return k[0] * h[:-2] + k[1] * h[1:-1] + k[2] * h[2:]


def scene_s1(w, h, seed):
    rng = np.random.default_rng(seed)
    img = np.full((h, w), 110.0, np.float32)
    n = (w * h) // 1500
    xs = rng.integers(0, w, n); ys = rng.integers(0, h, n)
    ws = rng.integers(4, 60, n); hs = rng.integers(4, 40, n)
    vs = rng.integers(20, 235, n)
    for x, y, rw, rh, v in zip(xs, ys, ws, hs, vs):
        img[y:y + rh, x:x + rw] = v
    img = _blur3(img)
    img = img + rng.normal(0.0, 2.5, img.shape).astype(np.float32)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def scene_s2(w, h, seed):
    return np.random.default_rng(seed).integers(0, 256, (h, w), dtype=np.uint8)


def scene_s3(w, h, kind="const"):
    img = np.full((h, w), 128, np.uint8)
    if kind == "step":
        img[:, w // 2:] = 30
    return img


def stereo_pair(w, h, seed):
    """left = S1; right = left shifted by a per-row-constant disparity in [5,60] + independent noise."""
    rng = np.random.default_rng(seed + 500)
    left = scene_s1(w, h, seed)
    disp = rng.integers(5, 61, h)
    right = np.empty_like(left)
    for y in range(h):
        right[y] = np.roll(left[y], -int(disp[y]))
    noise = rng.normal(0.0, 1.5, right.shape)
    right = np.clip(np.rint(right.astype(np.float32) + noise), 0, 255).astype(np.uint8)
    return left, right


def frames(config_id, w, h, count, kind="s1"):
    gen = {"s1": scene_s1, "s2": scene_s2}[kind]
    return [gen(w, h, 1000 * config_id + i) for i in range(count)]


def stereo_batch(config_id, w, h, pairs):
    out = []
    for i in range(pairs):
        l, r = stereo_pair(w, h, 1000 * config_id + i)
        out += [l, r]
    return out


def matching_set(nq=2000, nt=100000, seed=5000):
    """C5: uniform random descriptors; for each query plant one train row with k1~U{0..40}
    flipped bits and a second with k1+U{1..30} flipped bits."""
    rng = np.random.default_rng(seed)
    q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    slots = rng.permutation(nt)[:2 * nq] if nt >= 2 * nq else rng.integers(0, nt, 2 * nq)
    for i in range(nq):
        k1 = int(rng.integers(0, 41))
        k2 = k1 + int(rng.integers(1, 31))
        for k, s in ((k1, slots[2 * i]), (k2, slots[2 * i + 1])):
            bits = np.unpackbits(q[i])
            flip = rng.choice(256, size=k, replace=False)
            bits[flip] ^= 1
            t[s] = np.packbits(bits)
    return q, t


def observation_lists(n_points=400, seed=2027, pool=4000, max_obs=40):
    """Descriptor pool + per-map-point observation lists (CSR) for ComputeDistinctiveDescriptors (orbmappoint.cpp:314-383):
    every third point observes near-duplicates of one descriptor, the others unrelated rows; lists of length 0, 1 and 2
    (medians tie: the first observation wins) appear at fixed places; bad[k] flags observations whose key frame is bad."""
    rng = np.random.default_rng(seed)
    desc = rng.integers(0, 256, (pool, 32), dtype=np.uint8)
    for c in range(0, pool, 50):
        for j in range(1, 25):
            bits = np.unpackbits(desc[c]); bits[rng.choice(256, int(rng.integers(0, 14)), replace=False)] ^= 1
            desc[c + j] = np.packbits(bits)
    offsets, indices = [0], []
    for p in range(n_points):
        n = (0, 1, 2, 2)[p % 4] if p % 10 < 4 and p % 20 < 10 else int(rng.integers(3, max_obs + 1))
        if p % 3 == 0:
            c = 50 * int(rng.integers(0, pool // 50)); ix = (c + rng.choice(25, min(n, 25), replace=False)).tolist()
        else:
            ix = rng.choice(pool, n, replace=False).tolist()
        indices += ix; offsets.append(len(indices))
    indices = np.asarray(indices, np.int32)
    bad = (rng.random(len(indices)) < 0.08).astype(np.uint8)
    return desc, np.asarray(offsets, np.int32), indices, bad


def drop_bad_observations(offsets, indices, bad):
    """What the reference's loop does with bad key frames (orbmappoint.cpp:331-333): they never enter vDescriptors."""
    keep = np.asarray(bad) == 0
    lens = np.add.reduceat(np.r_[keep.astype(np.int64), 0], np.asarray(offsets[:-1], np.int64)) if len(indices) else np.zeros(len(offsets) - 1, np.int64)
    lens[np.diff(offsets) == 0] = 0
    off2 = np.zeros(len(offsets), np.int32); off2[1:] = np.cumsum(lens)
    return off2, np.asarray(indices, np.int32)[keep]
