"""CPU, only where the reference checkout is mounted (/root/reference; skipped on the GPU box): the real
datasets through OUR loaders + host graph preparation against the dataset-level integer goldens recorded by
tests/golden/make_golden.py."""
import os

import numpy as np
import pytest

from relationprediction_b200.common import io
from relationprediction_b200.ops import Graph
from relationprediction_b200 import _lib

REF = "/root/reference/data"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference datasets not mounted")


@pytest.mark.parametrize("name", ["FB-Toutanova", "wn18", "FB15k"])
def test_dataset_goldens_and_graph_invariants(toy, name):
    gold = toy["dataset_stats"][name]
    d = os.path.join(REF, name)
    tr = io.read_triplets_as_array(os.path.join(d, "train.txt"), os.path.join(d, "entities.dict"),
                                   os.path.join(d, "relations.dict"))
    V, R = gold["V"], gold["R"]
    assert tr.shape == (gold["E_train"], 3) and tr[:5].tolist() == gold["first_triples"]
    assert [int(tr[:, k].astype(np.int64).sum()) for k in range(3)] == gold["checksum_s_r_o"]
    deg = np.bincount(np.concatenate([tr[:, 0], tr[:, 2]]), minlength=V)
    assert int((deg == 0).sum()) == gold["isolated"] and int(deg.max()) == gold["max_degree"]
    g = Graph(tr, V, R)   # host-side build
    info = g.info()
    assert info[0] == 2 * len(tr) and info[3] == 2 * R
    rowptr = g.export(_lib.X_DST_ROWPTR)
    indeg = np.bincount(tr[:, 2], minlength=V) + np.bincount(tr[:, 0], minlength=V)   # fwd into o, bwd into s
    np.testing.assert_array_equal(np.diff(rowptr), indeg)
    norm = g.export(_lib.X_DST_NORM)
    relw = g.export(_lib.X_DST_RELW)
    # per direction the norms of every destination row sum to 1 (or the row has no message of that direction)
    rows = np.repeat(np.arange(V), np.diff(rowptr))
    for lo, hi in ((0, R), (R, 2 * R)):
        sel = (relw >= lo) & (relw < hi)
        sums = np.bincount(rows[sel], weights=norm[sel].astype(np.float64), minlength=V)
        assert np.all((np.abs(sums - 1) < 1e-4) | (sums == 0))
    # (dst, weight id) run count reported by the library == independent count
    key = rows.astype(np.int64) * (2 * R) + relw
    assert info[9] == 1 + int((key[1:] != key[:-1]).sum())
    if name == "FB-Toutanova":
        assert info[9] == 149689   # 544 230 messages collapse to 149 689 block mat-vecs
