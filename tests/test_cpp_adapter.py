"""The drop-in C++ adapter (cpp/orbextractor_b200.hpp, cpp/orbmatcher_b200.hpp): compiles against the
cv shim on CPU; on the GPU it is driven like OrbFrame drives the reference class and compared with the oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "opendlv-perception-vision-orbslam2_b200")


def _build(tmp_path):
    exe = str(tmp_path / "adapter_main")
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-Wall", "-I", os.path.join(ROOT, "cpp"), "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "oracle", "cvshim"), os.path.join(ROOT, "tests", "cpp", "adapter_main.cpp"),
                           "-o", exe, "-L", PKG, "-lorbx", f"-Wl,-rpath,{PKG}"])
    return exe


def test_adapter_compiles_with_reference_signatures(tmp_path):
    exe = _build(tmp_path)
    assert os.path.exists(exe)
    hdr = open(os.path.join(ROOT, "cpp", "orbextractor_b200.hpp")).read()
    # the public surface of the reference's include/orbextractor.hpp:92-109
    for sig in ("OrbExtractor(int nFeatures, float scaleFactor, int nLevels, int initialFastTh, int minFastTh)",
                "void ExtractFeatures(cv::InputArray image, std::vector<cv::KeyPoint> &keypoints, cv::OutputArray descriptors)",
                "int getLevels()", "double getScaleFactor()", "std::vector<float> getScaleFactors()",
                "std::vector<float> getInverseScaleFactors()", "std::vector<float> getScaleSigmaSquares()",
                "std::vector<float> getInverseScaleSigmaSquares()",
                "Pyramid m_vImagePyramid;   // std::vector<cv::Mat> m_vImagePyramid; in the reference"):
        assert sig in hdr, sig


@pytest.mark.gpu
def test_adapter_end_to_end(tmp_path, oracle):
    exe = _build(tmp_path)
    w, h, nf, nl = 752, 480, 1200, 8       # EuRoC-sized frame
    img = synth.scene_s1(w, h, 2024)
    raw = tmp_path / "in.raw"; out = tmp_path / "out.bin"
    img.tofile(raw)
    subprocess.check_call([exe, str(raw), str(w), str(h), str(nf), str(nl), str(out)])
    b = open(out, "rb").read()
    n = struct.unpack_from("<i", b, 0)[0]; o = 4
    kps = np.frombuffer(b, oracle.KP_DTYPE, n, o); o += 28 * n
    desc = np.frombuffer(b, np.uint8, 32 * n, o).reshape(n, 32); o += 32 * n
    levels = struct.unpack_from("<i", b, o)[0]; o += 4
    sf = np.frombuffer(b, np.float32, levels, o); o += 4 * levels
    is2 = np.frombuffer(b, np.float32, levels, o); o += 4 * levels
    oex = oracle.Extractor(nf, 1.2, nl)
    okps, odesc = oex.extract(img)
    assert n == len(okps) and kps.tobytes() == okps.tobytes() and np.array_equal(desc, odesc)
    assert levels == nl and np.array_equal(sf, np.array(oex.params.sf[:nl], np.float32))
    assert np.array_equal(is2, np.array(oex.params.inv_sigma2[:nl], np.float32))
    for l in range(nl):
        lw, lh = struct.unpack_from("<ii", b, o); o += 8
        a = np.frombuffer(b, np.uint8, lw * lh, o).reshape(lh, lw); o += lw * lh
        assert np.array_equal(a, oex.level(l)), f"m_vImagePyramid[{l}]"
    idx = np.frombuffer(b, np.int32, n, o); o += 4 * n
    d1 = np.frombuffer(b, np.int32, n, o); o += 4 * n
    d2 = np.frombuffer(b, np.int32, n, o); o += 4 * n
    oi, o1, o2 = oracle.knn2(desc, desc)
    assert np.array_equal(idx, oi) and np.array_equal(d1, o1) and np.array_equal(d2, o2)
    assert (d1 == 0).all()
    nm, nu = struct.unpack_from("<ii", b, o); o += 8
    uR = np.frombuffer(b, np.float32, nu, o); o += 4 * nu
    dep = np.frombuffer(b, np.float32, nu, o); o += 4 * nu
    oex2 = oracle.Extractor(nf, 1.2, nl); oex2.extract(img)
    ou, od, on = oracle.stereo_matches(oex, oex2, okps, odesc, okps, odesc, 400.0, 0.0)
    assert nu == n and nm == on and nm > 0        # (identical images: all rejected by the median filter, as in the reference)
    assert np.array_equal(uR.view(np.uint32), ou.view(np.uint32)) and np.array_equal(dep.view(np.uint32), od.view(np.uint32))
    # ---- vocabulary adapter (transform4 bookkeeping) and distinctive descriptors
    nb, nfv = struct.unpack_from("<ii", b, o); o += 8
    bsum = struct.unpack_from("<d", b, o)[0]; o += 8
    words = np.frombuffer(b, np.int32, n, o); o += 4 * n
    nodes = np.frombuffer(b, np.int32, n, o); o += 4 * n
    best = np.frombuffer(b, np.int32, 3, o); o += 12
    k, L, nn = 3, 2, 13
    off, ids = [0], []
    for v in range(nn):
        if v < 1 + k:
            ids += [1 + v * k + c for c in range(k)]
        off.append(len(ids))
    wid = np.array([-1] * (1 + k) + list(range(k * k)), np.int32)
    wt = np.array([0.0 if (v < 1 + k or v % 4 == 0) else 0.5 + v for v in range(nn)])
    ow, on = oracle.voc_transform(off, ids, desc[:nn], wid, L, 1, desc)
    assert np.array_equal(words, ow) and np.array_equal(nodes, on)
    leaf_of_word = np.flatnonzero(wid >= 0)
    kept = wt[leaf_of_word][ow] > 0
    assert nfv == int(kept.sum()) and nb == len(set(ow[kept].tolist())) and abs(bsum - 1.0) < 1e-12
    ob, _ = oracle.distinctive(desc, [0, 5, 5, 12], [0, 1, 2, 3, 4, 7, 7, 8, 9, 10, 11, 12])
    assert best.tolist() == ob.tolist() and best[1] == -1
    # ---- SearchByProjection adapter
    nm = struct.unpack_from("<i", b, o)[0]; o += 4
    asg = np.frombuffer(b, np.int32, n, o); o += 4 * n
    src = np.arange(0, n - 1, 2)
    occ = np.zeros(n, np.uint8); occ[src[src % 10 == 0]] = 1
    om, oa, onm = oracle.search_by_projection(kps, np.full(n, -1, np.float32), occ, desc, (0.0, 0.0, float(w), float(h)), desc[src],
                                              kps["x"][src] + np.float32(1.25), kps["y"][src], kps["octave"][src],
                                              np.float32(4.0) * sf[kps["octave"][src]], 0.8, 100,
                                              mp_observed=(np.arange(len(src)) % 3 != 0).astype(np.uint8))
    assert nm == onm and nm > 100 and np.array_equal(asg, oa)
    # ---- GetFeaturesInArea adapter
    nq, total = struct.unpack_from("<ii", b, o); o += 8
    aoff = np.frombuffer(b, np.int32, nq + 1, o); o += 4 * (nq + 1)
    aind = np.frombuffer(b, np.int32, total, o); o += 4 * total
    adist = np.frombuffer(b, np.int32, total, o); o += 4 * total
    oo, oi, od = oracle.area_distances(kps, desc, (0.0, 0.0, float(w), float(h)), desc[src], kps["x"][src] + np.float32(1.25), kps["y"][src],
                                       np.float32(4.0) * sf[kps["octave"][src]], kps["octave"][src] - 1, kps["octave"][src])
    assert nq == len(src) and np.array_equal(aoff, oo) and np.array_equal(aind, oi) and np.array_equal(adist, od)
    # ---- AssignFeaturesToGrid adapter (the reference's m_grid, cell by cell)
    gs, gi = oracle.assign_grid(kps, (0.0, 0.0, float(w), float(h)))
    for cell in range(64 * 48):
        cnt = struct.unpack_from("<i", b, o)[0]; o += 4
        got = np.frombuffer(b, np.int32, cnt, o); o += 4 * cnt
        assert np.array_equal(got, gi[gs[cell]:gs[cell + 1]]), cell
    # ---- FilterKeyPoints adapter: what the stereo matcher saw on the left after filtering
    nu = struct.unpack_from("<i", b, o)[0]; o += 4
    fk, _ = oracle.filter_keypoints(kps, desc, (0.25 * w, 0.75 * w, 0.25 * h, 0.75 * h))
    assert nu == len(fk) and 0 < nu < n
    assert o == len(b)
