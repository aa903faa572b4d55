"""CPU: liborbx.so loads and exports every symbol include/orbx.h declares; without a GPU the entry
points fail loudly (no CPU fallback exists)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "orbx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(orb[xmv]_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    import orbx
    L = orbx.lib()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"liborbx.so does not export {n}"
    assert sorted(orbx.EXPORTS) == names
    assert b"sm_100a" in L.orbx_version()


def test_keypoint_record_layout():
    import orbx
    assert orbx.KP_DTYPE.itemsize == 28      # cv::KeyPoint: 5 floats + 2 ints
    assert C.sizeof(orbx.Config) == 4 * 9 + 4 * 7 + 4


def test_no_gpu_means_loud_failure():
    import orbx
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(orbx.OrbxError) as e:
        orbx.Extractor()
    assert e.value.code == -4
    with pytest.raises(orbx.OrbxError) as e:
        orbx.Matcher()
    assert e.value.code == -4


def test_bad_config_is_rejected_before_touching_cuda():
    import orbx
    with pytest.raises(orbx.OrbxError) as e:
        orbx.Extractor(nlevels=0)
    assert e.value.code == -1
    with pytest.raises(orbx.OrbxError) as e:
        orbx.Extractor(ini_th=5, min_th=7)    # the single-pass fallback needs minTh <= iniTh
    assert e.value.code == -1


def test_product_never_links_the_oracle():
    import subprocess
    out = subprocess.run(["nm", "-D", os.path.join(ROOT, "opendlv-perception-vision-orbslam2_b200", "liborbx.so")],
                         capture_output=True, text=True).stdout
    assert "orbo_" not in out and "orbref_" not in out
