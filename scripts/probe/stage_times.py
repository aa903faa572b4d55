#!/usr/bin/env python3
"""Per-stage device times (orbx_profile_stages) and e2e latency of orbx_extract_batch at several batch sizes."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "opendlv-perception-vision-orbslam2_b200"))
import torch, orbx, synth
W, H = 1241, 376
frames = synth.stereo_batch(2, W, H, 32)
for b in (1, 2, 8, 16, 32, 64):
    ex = orbx.Extractor(2000, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=b)
    hb = torch.empty((b, H, W), dtype=torch.uint8).pin_memory()
    for f in range(b):
        hb[f] = torch.from_numpy(frames[f])
    imgs = [hb[f].numpy() for f in range(b)]
    cap = ex.max_keypoints
    out = (torch.zeros(b * cap * 28, dtype=torch.uint8).pin_memory().numpy().view(orbx.KP_DTYPE).reshape(b, cap),
           torch.zeros((b, cap, 32), dtype=torch.uint8).pin_memory().numpy(), np.zeros(b, np.int32))
    for _ in range(5):
        ex.extract_batch(imgs, out=out)
    t0 = time.perf_counter()
    n = 30
    for _ in range(n):
        ex.extract_batch(imgs, out=out)
    dt = (time.perf_counter() - t0) / n
    st = ex.profile_stages(reps=5)
    print(f"batch {b:3d}: e2e {dt * 1e3:7.3f} ms/call ({b / dt:9.0f} frames/s)  stages(us): " +
          " ".join(f"{k}={v * 1e3:.0f}" for k, v in st.items()), flush=True)
    ex.close()
