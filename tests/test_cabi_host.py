"""CPU: the C-ABI library builds, loads, exports every symbol include/rgcn_b200.h declares, and its
host-side graph preparation is bit-exact against an independent numpy restatement (oracle)."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import rgcn_oracle as oracle
from relationprediction_b200 import _lib
from relationprediction_b200.ops import Graph
from conftest import synthetic_kg, ROOT


def test_library_loads_and_exports_header_symbols():
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "rgcn_b200.h")).read()
    declared = set(re.findall(r"\b((?:rgcn|distmult)_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.rgcn_version() >= 100


def _check_views(g, dst, src, relw, norm, V_dst, V_src, n_relw):
    ref = oracle.sorted_views(dst, src, relw, norm, V_dst, V_src, n_relw, g.info()[13],
                              adaptive="RGCN_SUPERTILE_ROWS" not in os.environ)
    pairs = [(_lib.X_DST_ROWPTR, "dst_rowptr"), (_lib.X_DST_SRC, "dst_src"), (_lib.X_DST_RELW, "dst_relw"),
             (_lib.X_DST_NORM, "dst_norm"), (_lib.X_DST_MID, "dst_mid"), (_lib.X_SRC_ROWPTR, "src_rowptr"),
             (_lib.X_SRC_DST, "src_dst"), (_lib.X_SRC_RELW, "src_relw"), (_lib.X_SRC_NORM, "src_norm"),
             (_lib.X_SRC_MID, "src_mid"), (_lib.X_REL_PTR, "rel_ptr"), (_lib.X_REL_DST, "rel_row"),
             (_lib.X_REL_SRC, "rel_nbr"), (_lib.X_REL_NORM, "rel_norm"), (_lib.X_REL_MID, "rel_mid"),
             (_lib.X_REL2_PTR, "rel2_ptr"), (_lib.X_REL2_SRC, "rel2_row"), (_lib.X_REL2_DST, "rel2_nbr"),
             (_lib.X_REL2_NORM, "rel2_norm"), (_lib.X_REL2_MID, "rel2_mid")]
    for which, key in pairs:
        got = g.export(which)
        np.testing.assert_array_equal(got, ref[key], err_msg=key)  # bit-exact, floats included
    np.testing.assert_array_equal(g.export(_lib.X_MSG_NORM), norm)


def test_toy_graph_prep_bit_exact(toy):
    tr = np.array(toy["train"], dtype=np.int32)
    V, R = toy["V"], toy["R"]
    g = Graph(tr, V, R)  # host only
    dst, src, relw, norm = oracle.messages_from_triples(tr, R, V)
    assert g.M == 86 and g.V_dst == g.V_src == 16 and g.n_relw == 18
    _check_views(g, dst, src, relw, norm, V, V, 2 * R)
    np.testing.assert_array_equal(g.export(_lib.X_MSG_NORM)[:43], np.float32(toy["norm_f_canonical"]))
    np.testing.assert_array_equal(g.export(_lib.X_MSG_NORM)[43:], np.float32(toy["norm_b_canonical"]))


def test_explicit_norm_mode_carries_tf_compat_values(toy):
    tr = np.array(toy["train"], dtype=np.int32)
    nf, nb = oracle.graph_norms(tr, 16, "tf_unsorted_compat")
    g = Graph(tr, 16, 9, norm_mode="explicit", norm_f=nf, norm_b=nb)
    np.testing.assert_array_equal(g.export(_lib.X_MSG_NORM), np.concatenate([nf, nb]))
    g2 = Graph(tr, 16, 9, norm_mode="none")
    assert (g2.export(_lib.X_MSG_NORM) == 1).all()


@pytest.mark.parametrize("skewed", [False, True])
def test_synthetic_graph_prep_bit_exact_and_work_items(skewed):
    V, R, E = 3000, 37, 40000
    tr = synthetic_kg(V, R, E, seed=7, skewed=skewed)
    g = Graph(tr, V, R)
    dst, src, relw, norm = oracle.messages_from_triples(tr, R, V)
    _check_views(g, dst, src, relw, norm, V, V, 2 * R)
    info = g.info()
    # (dst, relw) run count
    p = np.lexsort((relw, dst))
    runs = 1 + int(((dst[p][1:] != dst[p][:-1]) | (relw[p][1:] != relw[p][:-1])).sum())
    assert info[9] == runs
    item_max = info[12]
    deg = np.bincount(dst, minlength=V)
    n_items = int(sum(1 if x <= item_max else -(-x // (-(-x // -(-x // item_max)))) for x in deg))
    assert info[7] == int((deg > item_max).sum())
    assert info[4] >= V and abs(info[4] - n_items) <= info[7]


def test_supertiled_weight_major_views(monkeypatch):
    monkeypatch.setenv("RGCN_SUPERTILE_ROWS", "257")
    V, R, E = 3000, 11, 30000
    tr = synthetic_kg(V, R, E, seed=9, skewed=True)
    g = Graph(tr, V, R)
    assert g.info()[13] == 257 and g.info()[14] == 12
    dst, src, relw, norm = oracle.messages_from_triples(tr, R, V)
    _check_views(g, dst, src, relw, norm, V, V, 2 * R)


def test_sparse_views_get_larger_supertiles():
    """view_supertile_rows: a weight-id-major view with fewer than 48 messages per (supertile, weight id) item grows
    its supertiles x2 up to 32768 rows; dense views keep the default 8192; the views stay exact either way."""
    V, R = 40000, 20
    tr = synthetic_kg(V, R, 1000, seed=3, skewed=False)            # 2000 messages over 5 x 40 default items: sparse
    g = Graph(tr, V, R)
    assert g.info()[13] == 8192 and g.info()[14] == 2               # capped at 32768 rows: ceil(40000 / 32768)
    dst, src, relw, norm = oracle.messages_from_triples(tr, R, V)
    _check_views(g, dst, src, relw, norm, V, V, 2 * R)
    tr = synthetic_kg(V, 2, 60000, seed=4, skewed=False)           # 120000 messages over 5 x 4 items: dense
    g = Graph(tr, V, 2)
    assert g.info()[14] == 5
    # in between: one doubling is enough (16384 rows: 3 supertiles x 40 weight ids x 48 <= 6000 messages)
    tr = synthetic_kg(V, R, 3000, seed=5, skewed=False)
    assert 3 * 40 * 48 <= 6000 < 5 * 40 * 48
    g = Graph(tr, V, R)
    assert g.info()[14] == 3
    dst, src, relw, norm = oracle.messages_from_triples(tr, R, V)
    _check_views(g, dst, src, relw, norm, V, V, 2 * R)


def test_empty_and_ragged_graphs():
    g = Graph(np.zeros((0, 3), np.int32), 5, 2)
    assert g.M == 0 and g.info()[4] == 5  # one (empty) work item per destination row
    assert g.export(_lib.X_DST_ROWPTR).tolist() == [0] * 6
    # a node that only sends, a self loop, duplicate triples
    tr = np.array([[0, 1, 0], [0, 1, 0], [3, 0, 0], [4, 1, 2]], np.int32)
    g = Graph(tr, 5, 2)
    dst, src, relw, norm = oracle.messages_from_triples(tr, 2, 5)
    _check_views(g, dst, src, relw, norm, 5, 5, 4)


def test_message_constructor_with_halo_rows():
    rng = np.random.RandomState(0)
    V_dst, V_src, n_relw, M = 50, 80, 6, 700
    dst = rng.randint(0, V_dst, M).astype(np.int32)
    src = rng.randint(0, V_src, M).astype(np.int32)
    relw = rng.randint(0, n_relw, M).astype(np.int32)
    norm = rng.rand(M).astype(np.float32)
    g = Graph.from_messages(dst, src, relw, norm, V_dst, V_src, n_relw)
    assert (g.V_dst, g.V_src, g.n_relw, g.M) == (V_dst, V_src, n_relw, M)
    _check_views(g, dst, src, relw, norm, V_dst, V_src, n_relw)


def test_error_codes_instead_of_exceptions():
    lib = _lib.load()
    with pytest.raises(_lib.RgcnError, match="out of range"):
        Graph(np.array([[0, 0, 9]], np.int32), 5, 2)
    with pytest.raises(_lib.RgcnError, match="out of range"):
        Graph(np.array([[0, 7, 1]], np.int32), 5, 2)
    g = Graph(np.array([[0, 0, 1]], np.int32), 5, 2)  # host-only graph
    buf = ctypes.create_string_buffer(1024)
    rc = lib.rgcn_block_forward(g.handle, 8, 2, buf, buf, buf, buf, None, 1.0, 1, buf, buf, 1024, None)
    assert rc == -5  # RGCN_ERR_NODEVICE: no silent CPU path
    assert b"host-only" in lib.rgcn_last_error()
    assert lib.rgcn_graph_export(g.handle, 99, buf, 1024) == -1


def test_next_row_entry_points_reject_bad_arguments_without_a_gpu():
    """Argument validation of the optimizer and sampler entry points happens before any device work."""
    lib = _lib.load()
    buf = ctypes.create_string_buffer(64)
    assert lib.rgcn_adam_update(None, buf, buf, buf, 4, 0.01, 0.9, 0.999, 1e-8, 1, None, 0.0, None) == -1
    assert lib.rgcn_adam_update(buf, buf, buf, buf, 4, 0.01, 0.9, 0.999, 1e-8, 0, None, 0.0, None) == -1  # 1-based step
    assert b"step" in lib.rgcn_last_error()
    assert lib.rgcn_adam_update(buf, buf, buf, buf, 0, 0.01, 0.9, 0.999, 1e-8, 1, None, 0.0, None) == 0     # empty tensor
    assert lib.rgcn_sumsq_accumulate(None, 4, buf, None) == -1
    assert lib.rgcn_sumsq_accumulate(buf, 0, buf, None) == 0
    h = ctypes.c_void_p()
    tri = np.array([[0, 0, 1], [1, 0, 7]], np.int32)
    assert lib.rgcn_sampler_create(ctypes.c_void_p(tri.ctypes.data), 2, 5, ctypes.byref(h)) == -1          # id 7 >= V
    assert not h.value
    assert lib.rgcn_sampler_create(ctypes.c_void_p(tri.ctypes.data), 2, 8, None) == -1
    assert lib.rgcn_sampler_create(ctypes.c_void_p(tri.ctypes.data), 2, 8, ctypes.byref(h)) == 0 and h.value
    out = np.empty(4, np.int32)
    assert lib.rgcn_sampler_draw(h, 3, 1, ctypes.c_void_p(out.ctypes.data)) == -1                           # > E
    assert lib.rgcn_sampler_draw(None, 1, 1, ctypes.c_void_p(out.ctypes.data)) == -1
    assert lib.rgcn_sampler_draw(h, 2, 1, ctypes.c_void_p(out.ctypes.data)) == 0 and sorted(out[:2].tolist()) == [0, 1]
    lib.rgcn_sampler_destroy(h)
    lib.rgcn_sampler_destroy(None)


def test_built_library_has_no_async_copy_on_undefined_uniform_registers():
    """ptxas 12.9 once emitted LDGSTS (cp.async + L2 cache hint) reading uniform registers no instruction writes; the
    kernel faulted with 'illegal instruction' on the GPU.  scripts/check_sass_ur.py scans the built .so for it."""
    import shutil
    import sys
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import check_sass_ur
    _lib.load()
    n, bad = check_sass_ur.check(_lib.LIB_PATH)
    assert n > 50 and not bad, bad
