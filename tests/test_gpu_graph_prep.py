"""GPU: graph preparation on the device (graph_device.cu, CUB radix sorts) is bit-exact against the
host builder (graph.cu), which is itself pinned against the numpy restatement in test_cabi_host.py."""
import numpy as np
import pytest

from relationprediction_b200 import _lib
from relationprediction_b200.ops import Graph
from conftest import synthetic_kg

pytestmark = pytest.mark.gpu
ALL = list(range(21))


def same(gd, gh):
    idev, ihost = gd.info(), gh.info()
    for k in (0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 12, 13, 14, 15):
        assert idev[k] == ihost[k], (k, idev, ihost)
    for which in ALL:
        np.testing.assert_array_equal(gd.export(which), gh.export(which), err_msg=str(which))


@pytest.mark.parametrize("V,R,E,skewed", [(16, 9, 43, False), (3000, 37, 40000, True), (20000, 237, 300000, True),
                                          (50000, 1000, 200000, False), (5, 2, 0, False)])
def test_device_prep_equals_host_prep(V, R, E, skewed, toy):
    tr = np.array(toy["train"], np.int32) if (V, E) == (16, 43) else synthetic_kg(V, R, E, seed=3, skewed=skewed)
    same(Graph(tr, V, R, device=0), Graph(tr, V, R))


def test_device_prep_small_items_and_supertiles(monkeypatch):
    monkeypatch.setenv("RGCN_ITEM_MAX", "8")
    monkeypatch.setenv("RGCN_SUPERTILE_ROWS", "100")
    tr = synthetic_kg(1500, 11, 20000, seed=4, skewed=True)
    same(Graph(tr, 1500, 11, device=0), Graph(tr, 1500, 11))


def test_device_prep_explicit_and_none_norms(toy):
    tr = np.array(toy["train"], np.int32)
    rng = np.random.RandomState(0)
    nf, nb = rng.rand(43).astype(np.float32), rng.rand(43).astype(np.float32)
    same(Graph(tr, 16, 9, norm_mode="explicit", norm_f=nf, norm_b=nb, device=0),
         Graph(tr, 16, 9, norm_mode="explicit", norm_f=nf, norm_b=nb))
    same(Graph(tr, 16, 9, norm_mode="none", device=0), Graph(tr, 16, 9, norm_mode="none"))


def test_device_prep_message_constructor_and_host_fallback(monkeypatch):
    rng = np.random.RandomState(1)
    V_dst, V_src, n_relw, M = 500, 800, 6, 9000
    dst = rng.randint(0, V_dst, M).astype(np.int32)
    src = rng.randint(0, V_src, M).astype(np.int32)
    relw = rng.randint(0, n_relw, M).astype(np.int32)
    norm = rng.rand(M).astype(np.float32)
    gh = Graph.from_messages(dst, src, relw, norm, V_dst, V_src, n_relw)
    same(Graph.from_messages(dst, src, relw, norm, V_dst, V_src, n_relw, device=0), gh)
    monkeypatch.setenv("RGCN_PREP", "host")   # host sort + upload path stays available
    same(Graph.from_messages(dst, src, relw, norm, V_dst, V_src, n_relw, device=0), gh)


def test_device_prep_rejects_bad_indices():
    with pytest.raises(_lib.RgcnError, match="out of range"):
        Graph(np.array([[0, 0, 9]], np.int32), 5, 2, device=0)
    with pytest.raises(_lib.RgcnError, match="out of range"):
        Graph.from_messages(np.array([7], np.int32), np.array([0], np.int32), np.array([0], np.int32),
                            np.array([1.0], np.float32), 5, 5, 2, device=0)
