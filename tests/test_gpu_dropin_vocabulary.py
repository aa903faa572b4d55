"""GPU: the reference's own OrbVocabulary with the drop-in transform4 inside.

oracle/_ref/libvocdropin.so is the reference's src/orbvocabulary.cpp (text loader, tree, scoring) compiled unmodified, with
the body of OrbVocabulary::transform4 replaced by the liborbx-backed one of INTEGRATION.md section 2c
(cpp/orbvocabulary_b200.hpp; the reference's definition is weakened in the object file).  oracle/_ref/libvocref.so is the
all-reference build.  The same vocabulary file, loaded by the reference's loader in both, must give the same bag of words
and feature vector."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
REFDIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")


@pytest.mark.skipif(not (os.path.exists(os.path.join(REFDIR, "libvocdropin.so")) and os.path.exists(os.path.join(REFDIR, "libvocref.so"))),
                    reason="the reference translation units are built where the reference tree is mounted")
@pytest.mark.parametrize("k,L", [(10, 3), (4, 5), (20, 2)])
def test_reference_vocabulary_with_drop_in_transform4(oracle, tmp_path, k, L):
    import orbx
    child_off, child_ids, node_desc, word_id, weight, _ = orbx.random_vocabulary(k, L, seed=300 + k)
    weight = np.round(weight, 6)
    weight[np.flatnonzero(word_id >= 0)[::5]] = 0.0                      # stopped words
    node_desc[child_ids[child_off[0] + 1]] = node_desc[child_ids[child_off[0]]]   # a tie between the first two children of the root
    feat = np.random.default_rng(k).integers(0, 256, (2000, 32), dtype=np.uint8)
    path = str(tmp_path / "voc.txt")
    oracle.write_vocabulary_text(path, child_off, child_ids, node_desc, weight, k, L)
    np.save(str(tmp_path / "feat.npy"), feat)
    ref = oracle.RefVocabulary(path)
    want = {lu: ref.transform4(feat, lu) for lu in (0, 1, L - 1, L + 2)}
    ref.close()
    # the drop-in build + liborbx in a process of their own
    out = str(tmp_path / "got.npz")
    code = (f"import sys, numpy as np; sys.path[:0] = {sys.path[:4]!r}; import orb_oracle_py as O\n"
            f"feat = np.load({str(tmp_path / 'feat.npy')!r}); v = O.RefVocabulary({path!r}, lib_name='libvocdropin.so'); res = {{}}\n"
            f"for lu in {(0, 1, L - 1, L + 2)!r}:\n"
            f"    r = v.transform4(feat, lu)\n"
            f"    for j, a in enumerate(r): res[f'{{lu}}_{{j}}'] = a\n"
            f"v.close(); np.savez({out!r}, **res)\n")
    subprocess.run([sys.executable, "-c", code], check=True, timeout=300)
    got = np.load(out)
    for lu, w in want.items():
        assert len(w[0]) > 0
        for j, a in enumerate(w):
            assert np.array_equal(got[f"{lu}_{j}"], a), (lu, j)
