import os, sys
import numpy as np
sys.path[:0] = ["opendlv-perception-vision-orbslam2_b200", "oracle"]
import orbx, synth
import orb_oracle_py as O
cases = [(556, 514, 3662, 7, 1.25, 27, 2, 4), (2088, 522, 898, 5, 1.25, 18, 10, 2)]
for (w, h, nf, nl, sf, ini, mn, batch) in cases:
    imgs = [synth.scene_s1(w, h, 5 + i) for i in range(batch)]
    ex = orbx.Extractor(nf, sf, nl, ini, mn, max_width=w, max_height=h, max_batch=batch)
    try:
        kps, desc, cnt = ex.extract_batch(imgs)
        oex = O.Extractor(nf, sf, nl, ini, mn)
        ok = all(kps[f, :cnt[f]].tobytes() == oex.extract(imgs[f])[0].tobytes() for f in range(batch))
        print(w, h, "ok", cnt.tolist(), "parity", ok)
    except orbx.OrbxError as e:
        print(w, h, "ERROR", e)
    try:
        k1, d1 = ex.extract(imgs[0]); print("  single ok", len(k1))
    except orbx.OrbxError as e:
        print("  single ERROR", e)
    ex.close()
