#!/usr/bin/env python3
"""kNN-2 of a rank's share at N = 8 (250 queries x 100000) on one GPU: a few calls for an ncu launch list."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "opendlv-perception-vision-orbslam2_b200"))
import torch, orbx, synth
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 250
q, t = synth.matching_set(2000, 100000)
dq = torch.from_numpy(q).cuda(); dt = torch.from_numpy(t).cuda()
st = torch.cuda.Stream()
m = orbx.Matcher(max_queries=2000, max_train=100000)
out = torch.zeros((nq, 4), dtype=torch.int32, device="cuda")
for _ in range(6):
    m.knn2_device(dq.data_ptr(), nq, dt.data_ptr(), 100000, out.data_ptr(), st.cuda_stream)
torch.cuda.synchronize()
m.close()
