#!/usr/bin/env python3
"""Golden fixture for the bag-of-words row (tests/golden/ref_vocabulary.npz); run in the BUILD container only.

Source of truth: the reference's own OrbVocabulary -- src/orbvocabulary.cpp, orbdescriptor.cpp, orbbowvector.cpp,
orbfeaturevector.cpp compiled UNMODIFIED against oracle/cvshim (oracle/_ref/libvocref.so, `make -C oracle ref`).
A synthetic 5-ary tree of depth 3 (the reference's ORBvoc.txt is not in its tree) is written in the DBoW2 text format,
loaded by the reference's own loader and run through transform4 / transform5.
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "opendlv-perception-vision-orbslam2_b200")]
import orb_oracle_py as O  # noqa: E402
import orbx  # noqa: E402  (random_vocabulary only: pure numpy)

k, L = 5, 3
child_off, child_ids, node_desc, word_id, weight, _ = orbx.random_vocabulary(k, L, seed=2026)
weight = np.round(weight, 6)                      # survives the text round trip exactly
weight[np.flatnonzero(word_id >= 0)[::7]] = 0.0   # some stopped words
node_desc[child_ids[1]] = node_desc[child_ids[0]]  # a tie between the first two children of the root
rng = np.random.default_rng(7)
feat = rng.integers(0, 256, (400, 32), dtype=np.uint8)
feat[:40] = node_desc[rng.integers(1, len(node_desc), 40)]
with tempfile.TemporaryDirectory() as d:
    path = os.path.join(d, "voc.txt")
    O.write_vocabulary_text(path, child_off, child_ids, node_desc, weight, k, L)
    ref = O.RefVocabulary(path)
    assert ref.size() == int((word_id >= 0).sum())
    out = {"k": k, "L": L, "child_off": child_off, "child_ids": child_ids, "node_desc": node_desc, "word_id": word_id,
           "weight": weight, "feat": feat}
    for lu in (0, 1, 2, 4):
        w, n = ref.transform_each(feat, lu)
        ids, vals, nodes, feats = ref.transform4(feat, lu)
        out[f"word_lu{lu}"] = w; out[f"node_lu{lu}"] = n
        out[f"bow_ids_lu{lu}"] = ids; out[f"bow_vals_lu{lu}"] = vals; out[f"fv_nodes_lu{lu}"] = nodes; out[f"fv_feats_lu{lu}"] = feats
    ref.close()
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_vocabulary.npz"), **out)
print("ref_vocabulary.npz", os.path.getsize(os.path.join(ROOT, "tests", "golden", "ref_vocabulary.npz")))
