#!/usr/bin/env python3
"""k_knn2_sharded against k_knn2_partial + k_knn2_merge on ONE GPU: (a) one rank, (b) two / eight ranks as slots of the same device
(every slot's kernel waits for the others' flags), for the query counts a rank sees at N = 1, 2, 8."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "opendlv-perception-vision-orbslam2_b200"))
import torch, orbx, synth
NQ, NT = 2000, 100000
q, t = synth.matching_set(NQ, NT)
dq = torch.from_numpy(q).cuda(); dt = torch.from_numpy(t).cuda()
st = torch.cuda.Stream()
def timeit(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(n): fn()
    e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for nq in (2000, 1000, 250):
    m = orbx.Matcher(max_queries=NQ, max_train=NT)
    out = torch.zeros((nq, 4), dtype=torch.int32, device="cuda")
    a = timeit(lambda: m.knn2_device(dq.data_ptr(), nq, dt.data_ptr(), NT, out.data_ptr(), st.cuda_stream))
    m.window_create(nq, 1, 0)
    b = timeit(lambda: m.knn2_sharded(dq.data_ptr(), nq, 0, dt.data_ptr(), NT, st.cuda_stream))
    print(f"nq {nq}: partial+merge {a * 1e3:.1f} us, sharded kernel with one rank {b * 1e3:.1f} us", flush=True)
    m.close()
