#!/usr/bin/env python3
"""Stage times (each kernel alone, CUDA events) of small batches: where the latency of a single frame goes."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "opendlv-perception-vision-orbslam2_b200"))
import torch, orbx, synth
W, H = 1241, 376
frames = synth.stereo_batch(2, W, H, 4)
for b in (1, 2, 8):
    ex = orbx.Extractor(2000, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=b)
    d = torch.from_numpy(np.stack(frames[:b])).cuda()
    st = torch.cuda.Stream()
    for _ in range(5):
        ex.extract_batch_device(d.data_ptr(), H * W, W, b, W, H, st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(50):
        ex.extract_batch_device(d.data_ptr(), H * W, W, b, W, H, st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    stg = ex.profile_stages(reps=10)
    print(f"batch {b}: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us/step  stages(us): " + " ".join(f"{k}={v * 1e3:.1f}" for k, v in stg.items()), flush=True)
    ex.close()
