import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "opendlv-perception-vision-orbslam2_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def pytest_sessionstart(session):
    """A fresh checkout has no built artefacts (they are git-ignored): build liborbx.so once where nvcc exists, exactly as
    __graft_entry__.build() does.  Nothing is built implicitly on a box without the toolchain -- the tests then fail loudly."""
    import shutil
    import subprocess
    lib = os.path.join(ROOT, "opendlv-perception-vision-orbslam2_b200", "liborbx.so")
    if not os.path.exists(lib) and (shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")):
        env = dict(os.environ, PATH=os.environ.get("PATH", "") + ":/usr/local/cuda/bin")
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "opendlv-perception-vision-orbslam2_b200", "csrc")], env=env)
    # the reference's own translation units against the header shim, only where the reference tree is mounted
    if os.path.isdir("/root/reference/src"):
        subprocess.call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])      # incremental: a no-op when up to date


@pytest.fixture(scope="session")
def oracle():
    import orb_oracle_py
    orb_oracle_py.lib()
    return orb_oracle_py
