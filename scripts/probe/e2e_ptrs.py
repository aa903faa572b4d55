#!/usr/bin/env python3
"""e2e frames/s of orbx_extract_batch (64-frame batch, pinned host buffers, prepared pointer array) for the
chunk plan / lane count / input mode in the environment."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "opendlv-perception-vision-orbslam2_b200"))
import torch, orbx, synth
W, H, b = 1241, 376, 64
frames = synth.stereo_batch(2, W, H, 32)
ex = orbx.Extractor(2000, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=b)
pools, ptrs = [], []
for p in range(6):
    hb = torch.empty((b, H, W), dtype=torch.uint8).pin_memory()
    for f in range(b):
        hb[f] = torch.from_numpy(np.roll(frames[(f + 5 * p) % b], 3 * p, axis=1))
    pools.append(hb); ptrs.append(orbx.Extractor.frame_pointers([hb[f].numpy() for f in range(b)]))
cap = ex.max_keypoints
out = (torch.zeros(b * cap * 28, dtype=torch.uint8).pin_memory().numpy().view(orbx.KP_DTYPE).reshape(b, cap),
       torch.zeros((b, cap, 32), dtype=torch.uint8).pin_memory().numpy(), np.zeros(b, np.int32))
for i in range(6):
    ex.extract_batch_ptrs(ptrs[i % 6], b, W, H, W, out)
ref = out[0].copy(), out[1].copy(), out[2].copy()
t0 = time.perf_counter()
n = 60
for i in range(n):
    ex.extract_batch_ptrs(ptrs[i % 6], b, W, H, W, out)
dt = (time.perf_counter() - t0) / n
env = {k: os.environ.get(k) for k in ("ORBX_ZEROCOPY", "ORBX_LANES", "ORBX_CHUNKS", "ORBX_CHUNK_PLAN") if os.environ.get(k)}
print(f"{env}: {dt * 1e3:.3f} ms/call  {b / dt:.0f} frames/s  kp {int(out[2].sum())}", flush=True)
ex.close()
