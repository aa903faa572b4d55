#!/usr/bin/env python3
"""soak_handle.py with a report of WHAT differs in a mismatching frame (debugging aid)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "opendlv-perception-vision-orbslam2_b200"), os.path.join(ROOT, "oracle")]
import torch, orbx, synth
import orb_oracle_py as O
n_calls = int(sys.argv[1]) if len(sys.argv) > 1 else 80
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
sizes = [(1241, 376), (752, 480)]
pool = {s: [synth.scene_s1(s[0], s[1], 100 + i) if i % 3 else synth.scene_s2(s[0], s[1], 100 + i) for i in range(12)] for s in sizes}
want = {}
for s in sizes:
    oex = O.Extractor(1500, 1.2, 8)
    want[s] = [oex.extract(img) for img in pool[s]]
ex = orbx.Extractor(1500, 1.2, 8, max_width=1241, max_height=480, max_batch=24)
bad = 0
for call in range(n_calls):
    s = sizes[int(rng.integers(0, 2))]; w, h = s
    b = int(rng.integers(1, 25)); idx = rng.integers(0, 12, b)
    mode = ("pinned", "pageable", "pitched", "device")[int(rng.integers(0, 4))]
    if mode == "pinned":
        hb = torch.empty((b, h, w), dtype=torch.uint8).pin_memory()
        for f in range(b): hb[f] = torch.from_numpy(pool[s][idx[f]])
        imgs = [hb[f].numpy() for f in range(b)]
        kps, desc, cnt = ex.extract_batch(imgs)
    elif mode == "pageable":
        kps, desc, cnt = ex.extract_batch([pool[s][i].copy() for i in idx])
    elif mode == "pitched":
        big = np.zeros((b, h, w + 37), np.uint8)
        for f in range(b): big[f, :, :w] = pool[s][idx[f]]
        kps, desc, cnt = ex.extract_batch([big[f, :, :w] for f in range(b)])
    else:
        dev = torch.from_numpy(np.stack([pool[s][i] for i in idx])).cuda()
        torch.cuda.synchronize()
        ex.extract_batch_device(dev.data_ptr(), w * h, w, b, w, h)
        kps, desc, cnt = ex.fetch_results(b)
    for f in range(b):
        okp, od = want[s][idx[f]]
        n = int(cnt[f])
        if not (n == len(okp) and kps[f, :n].tobytes() == okp.tobytes() and np.array_equal(desc[f, :n], od)):
            bad += 1
            g = kps[f, :n]
            msg = f"call {call}: {w}x{h} batch {b} {mode} frame {f} (pool {idx[f]}): count {n} vs {len(okp)}"
            if n == len(okp):
                diff = [k for k in range(n) if g[k].tobytes() != okp[k].tobytes()]
                dd = [k for k in range(n) if not np.array_equal(desc[f, k], od[k])]
                msg += f"; {len(diff)} keypoint records differ (first {diff[:5]}), {len(dd)} descriptors differ (first {dd[:5]})"
                for k in diff[:3]:
                    msg += f"\n    gpu {g[k]}  oracle {okp[k]}"
            else:
                go = np.bincount(g["octave"], minlength=8); oo = np.bincount(okp["octave"], minlength=8)
                msg += f"; per level gpu {go.tolist()} oracle {oo.tolist()}"
            print(msg, flush=True)
print(f"soak_handle_dbg: {n_calls} calls, {bad} bad frames")
