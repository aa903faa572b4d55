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
