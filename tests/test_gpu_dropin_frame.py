"""GPU: the reference's own OrbFrame with the drop-in extractor inside.

oracle/_ref/libdropinref.so is the reference's src/orbframe.cpp (stereo constructor: two extractor threads, FilterKeyPoints,
ComputeStereoMatches, AssignFeaturesToGrid) compiled UNMODIFIED, with cvshim/dropin/orbextractor.hpp shadowing the
reference's include/orbextractor.hpp by the liborbx-backed class of cpp/orbextractor_b200.hpp -- the substitution
INTEGRATION.md section 1 describes.  oracle/_ref/libframeref.so is the same frame code with the reference's own
src/orbextractor.cpp.  Both frames must be identical: key points, descriptors, the pyramids the frame reads through
m_vImagePyramid, mvuRight, m_depths, m_grid.

oracle/_ref/libdropin2ref.so goes one step further: the body of OrbFrame::ComputeStereoMatches is replaced as well (the
reference's definition is weakened in the object file, cvshim/frame_glue.cpp supplies the liborbx-backed one of
INTEGRATION.md section 2b, FilterKeyPoints on the device-resident results included), so the reference's unmodified stereo
constructor runs extraction AND stereo matching on the GPU."""
import os
import subprocess
import sys

import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu
REFDIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")


@pytest.mark.skipif(not (os.path.exists(os.path.join(REFDIR, "libdropinref.so")) and os.path.exists(os.path.join(REFDIR, "libframeref.so"))),
                    reason="the reference translation units are built where the reference tree is mounted")
@pytest.mark.parametrize("lib_name", ["libdropinref.so", "libdropin2ref.so"])
@pytest.mark.parametrize("w,h,seed,nf,nl,mbf,mb,bbox", [
    (1241, 376, 11, 2000, 8, 386.1, 0.537, None), (640, 360, 5, 1000, 6, 200.0, 0.4, (200.0, 420.0, 100.0, 260.0)),
    (752, 480, 9, 1200, 8, 435.2, 0.11, None)])
def test_reference_frame_with_drop_in_extractor(oracle, tmp_path, lib_name, w, h, seed, nf, nl, mbf, mb, bbox):
    if not os.path.exists(os.path.join(REFDIR, lib_name)):
        pytest.skip(f"{lib_name} not built")
    left, right = synth.stereo_pair(w, h, seed)
    ref = oracle.ref_stereo_frame(left, right, mbf, mb, nfeatures=nf, nlevels=nl, canonical=1, bbox=bbox)
    # the reference frame code + liborbx in one process of their own: a fault there must not take the test session down
    out = str(tmp_path / "frame.npz")
    code = (f"import sys, numpy as np; sys.path[:0] = {sys.path[:4]!r}; import orb_oracle_py as O, synth\n"
            f"l, r = synth.stereo_pair({w}, {h}, {seed})\n"
            f"g = O.ref_stereo_frame(l, r, {mbf}, {mb}, nfeatures={nf}, nlevels={nl}, canonical=0, bbox={bbox!r}, lib_name={lib_name!r})\n"
            f"lv = {{f'L{{i}}': a for i, a in enumerate(g.pop('levelsL'))}}; lv.update({{f'R{{i}}': a for i, a in enumerate(g.pop('levelsR'))}})\n"
            f"np.savez({out!r}, **g, **lv)\n")
    subprocess.run([sys.executable, "-c", code], check=True, timeout=300)
    z = np.load(out)
    got = {k: z[k] for k in z.files}
    got["levelsL"] = [z[f"L{i}"] for i in range(nl)]; got["levelsR"] = [z[f"R{i}"] for i in range(nl)]
    assert len(got["kl"]) == len(ref["kl"]) > 100 and got["kl"].tobytes() == ref["kl"].tobytes() and got["kr"].tobytes() == ref["kr"].tobytes()
    assert np.array_equal(got["dl"], ref["dl"]) and np.array_equal(got["dr"], ref["dr"])
    for l in range(nl):
        assert np.array_equal(got["levelsL"][l], ref["levelsL"][l]) and np.array_equal(got["levelsR"][l], ref["levelsR"][l]), l
    assert (ref["uRight"] >= 0).sum() > 50
    assert np.array_equal(got["uRight"].view(np.uint32), ref["uRight"].view(np.uint32))
    assert np.array_equal(got["depth"].view(np.uint32), ref["depth"].view(np.uint32))
    assert np.array_equal(got["grid_start"], ref["grid_start"]) and np.array_equal(got["grid_items"], ref["grid_items"])
