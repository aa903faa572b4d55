#!/usr/bin/env python3
"""Upper bound of cross-call overlap: two extractor handles on two streams take alternate 64-frame batches.
usage: two_handles.py [kitti|hd|uhd] [steps]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "opendlv-perception-vision-orbslam2_b200"))
import torch, orbx, synth
CFG = {"kitti": (1241, 376, 2000, 8, 64), "hd": (1920, 1080, 4000, 8, 16), "uhd": (3840, 2160, 8000, 12, 4)}
name = sys.argv[1] if len(sys.argv) > 1 else "kitti"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
W, H, nf, nl, b = CFG[name]
if name == "kitti":
    frames = synth.stereo_batch(2, W, H, b // 2)
else:
    base = [synth.scene_s1(W, H, 9000 + i) for i in range(min(b, 4))]
    frames = [np.roll(base[f % len(base)], 5 * f, axis=1) for f in range(b)]
d = torch.from_numpy(np.stack(frames)).cuda()
for nh in tuple(int(v) for v in os.environ.get("NH", "1,2,3").split(",")):
    exs = [orbx.Extractor(nf, 1.2, nl, 20, 7, max_width=W, max_height=H, max_batch=b) for _ in range(nh)]
    sts = [torch.cuda.Stream() for _ in range(nh)]
    for k in range(2 * nh):
        exs[k % nh].extract_batch_device(d.data_ptr(), H * W, W, b, W, H, sts[k % nh].cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    e0.record(main)
    for s in sts: s.wait_stream(main)
    for k in range(steps * nh):
        exs[k % nh].extract_batch_device(d.data_ptr(), H * W, W, b, W, H, sts[k % nh].cuda_stream)
    for s in sts: main.wait_stream(s)
    e1.record(main)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (steps * nh)
    print(f"{name} batch {b}, {nh} handle(s) alternating: {ms:.4f} ms/step {b / (ms * 1e-3):.0f} frames/s", flush=True)
    [e.close() for e in exs]
