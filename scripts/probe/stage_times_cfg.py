#!/usr/bin/env python3
"""Per-stage device times and e2e rate for the HD / UHD configurations of BASELINE.json (configs[2], configs[3])."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "opendlv-perception-vision-orbslam2_b200"))
import torch, orbx, synth
for name, W, H, nf, nl, b in (("kitti", 1241, 376, 2000, 8, 64), ("hd", 1920, 1080, 4000, 8, 16), ("uhd", 3840, 2160, 8000, 12, 4)):
    frames = [synth.scene_s1(W, H, 9000 + i) for i in range(min(b, 4))]
    ex = orbx.Extractor(nf, 1.2, nl, 20, 7, max_width=W, max_height=H, max_batch=b)
    hb = torch.empty((b, H, W), dtype=torch.uint8).pin_memory()
    for f in range(b):
        hb[f] = torch.from_numpy(np.roll(frames[f % len(frames)], 5 * f, axis=1))
    ptrs = orbx.Extractor.frame_pointers([hb[f].numpy() for f in range(b)])
    cap = ex.max_keypoints
    out = (torch.zeros(b * cap * 28, dtype=torch.uint8).pin_memory().numpy().view(orbx.KP_DTYPE).reshape(b, cap),
           torch.zeros((b, cap, 32), dtype=torch.uint8).pin_memory().numpy(), np.zeros(b, np.int32))
    for _ in range(3):
        ex.extract_batch_ptrs(ptrs, b, W, H, W, out)
    t0 = time.perf_counter()
    n = 10
    for _ in range(n):
        ex.extract_batch_ptrs(ptrs, b, W, H, W, out)
    dt = (time.perf_counter() - t0) / n
    d = hb.cuda()
    st = torch.cuda.Stream()
    for _ in range(3):
        ex.extract_batch_device(d.data_ptr(), H * W, W, b, W, H, st.cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(n):
        ex.extract_batch_device(d.data_ptr(), H * W, W, b, W, H, st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    stg = ex.profile_stages(reps=3)
    print(f"{name:5s} {W}x{H} nf={nf} L={nl} batch {b}: device {b / (ms * 1e-3):9.0f} frames/s ({ms:.3f} ms)  e2e {b / dt:9.0f} frames/s ({dt * 1e3:.3f} ms)  "
          f"kps/frame {int(out[2].mean())}  stages(us): " + " ".join(f"{k}={v * 1e3:.0f}" for k, v in stg.items()), flush=True)
    ex.close()
