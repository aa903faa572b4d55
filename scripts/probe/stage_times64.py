import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "opendlv-perception-vision-orbslam2_b200"))
import numpy as np, orbx, synth
W, H, b = 1241, 376, 64
frames = synth.stereo_batch(2, W, H, 32)
ex = orbx.Extractor(2000, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=b)
ex.extract_batch(frames)
for r in range(3):
    st = ex.profile_stages(reps=10)
    print(os.environ.get("ORBX_EXP_LISTDIV"), " ".join(f"{k}={v * 1e3:.1f}" for k, v in st.items()), flush=True)
