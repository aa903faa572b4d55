#!/usr/bin/env python3
"""Exactly three device-resident 64-frame KITTI steps on one handle (13 kernel launches each with ORBX_SPLIT=1); ncu captures the last:
ORBX_SPLIT=1 ncu --set full --clock-control none --import-source on --launch-skip 26 --launch-count 13 -o prof python one_step.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "opendlv-perception-vision-orbslam2_b200"))
import torch, orbx, synth
W, H, nf, nl, b = 1241, 376, 2000, 8, 64
frames = synth.stereo_batch(2, W, H, b // 2)
ex = orbx.Extractor(nf, 1.2, nl, 20, 7, max_width=W, max_height=H, max_batch=b)
d = torch.from_numpy(np.stack(frames)).cuda()
st = torch.cuda.Stream()
for _ in range(3):
    ex.extract_batch_device(d.data_ptr(), H * W, W, b, W, H, st.cuda_stream)
    torch.cuda.synchronize()
print("launches per step", ex.last_launches())
ex.close()
