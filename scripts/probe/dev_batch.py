#!/usr/bin/env python3
"""Run a few device-resident batches of one BASELINE configuration (for ncu captures and quick timings).
usage: dev_batch.py [kitti|hd|uhd] [steps]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "opendlv-perception-vision-orbslam2_b200"))
import torch, orbx, synth
CFG = {"kitti": (1241, 376, 2000, 8, 64), "hd": (1920, 1080, 4000, 8, 16), "uhd": (3840, 2160, 8000, 12, 4)}
name = sys.argv[1] if len(sys.argv) > 1 else "kitti"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
W, H, nf, nl, b = CFG[name]
if name == "kitti":
    frames = synth.stereo_batch(2, W, H, b // 2)
else:
    base = [synth.scene_s1(W, H, 9000 + i) for i in range(min(b, 4))]
    frames = [np.roll(base[f % len(base)], 5 * f, axis=1) for f in range(b)]
ex = orbx.Extractor(nf, 1.2, nl, 20, 7, max_width=W, max_height=H, max_batch=b)
d = torch.from_numpy(np.stack(frames)).cuda()
st = torch.cuda.Stream()
for _ in range(2):
    ex.extract_batch_device(d.data_ptr(), H * W, W, b, W, H, st.cuda_stream)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(steps):
    ex.extract_batch_device(d.data_ptr(), H * W, W, b, W, H, st.cuda_stream)
e1.record(st)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
stg = ex.profile_stages(reps=3)
print(f"{name} batch {b}: {ms:.4f} ms/step {b / (ms * 1e-3):.0f} frames/s  stages(us): " + " ".join(f"{k}={v * 1e3:.1f}" for k, v in stg.items()), flush=True)
ex.close()
