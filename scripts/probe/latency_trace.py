#!/usr/bin/env python3
"""Latency of orbx_extract_batch for 1 and 2 frames (pinned host buffers); with ORBX_TRACE=1 the library prints the
per-chunk timeline (H2D done / kernels done / D2H done)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "opendlv-perception-vision-orbslam2_b200"))
import torch, orbx, synth
W, H = 1241, 376
frames = synth.stereo_batch(2, W, H, 1)
for b in (1, 2):
    ex = orbx.Extractor(2000, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=b)
    hb = torch.empty((b, H, W), dtype=torch.uint8).pin_memory()
    for f in range(b):
        hb[f] = torch.from_numpy(frames[f])
    ptrs = orbx.Extractor.frame_pointers([hb[f].numpy() for f in range(b)])
    cap = ex.max_keypoints
    out = (torch.zeros(b * cap * 28, dtype=torch.uint8).pin_memory().numpy().view(orbx.KP_DTYPE).reshape(b, cap),
           torch.zeros((b, cap, 32), dtype=torch.uint8).pin_memory().numpy(), np.zeros(b, np.int32))
    for i in range(20):
        ex.extract_batch_ptrs(ptrs, b, W, H, W, out)
    t0 = time.perf_counter()
    n = 200
    for i in range(n):
        ex.extract_batch_ptrs(ptrs, b, W, H, W, out)
    dt = (time.perf_counter() - t0) / n
    print(f"batch {b}: {dt * 1e3:.4f} ms per call, launches {ex.last_launches()}", flush=True)
    ex.close()
