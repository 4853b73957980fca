"""CPU: the next-row host pieces against goldens produced by the reference's own importable numpy
modules (tests/golden/make_golden.py): evaluation ranking (common/evaluation.py) and the negative
sampler (common/auxilliaries.py)."""
import os

import numpy as np
import pytest

from relationprediction_b200.common import auxilliaries, evaluation
from conftest import GOLDEN


def test_ranking_matches_reference_evaluation():
    g = np.load(os.path.join(GOLDEN, "eval_golden.npz"))
    T, train = g["T"], g["train"]

    class FakeModel:
        def score_all_subjects(self, tr):
            return np.stack([T[r, :, o] for s, r, o in tr])

        def score_all_objects(self, tr):
            return np.stack([T[r, s, :] for s, r, o in tr])
    sc = evaluation.Scorer()
    sc.register_data(train)
    sc.register_model(FakeModel())
    res = sc.compute_scores(train[:40]).get_summary().results
    np.testing.assert_allclose([res['Raw'][m] for m in ('MRR', 'H@1', 'H@3', 'H@10')], g["raw"], rtol=1e-12)
    np.testing.assert_allclose([res['Filtered'][m] for m in ('MRR', 'H@1', 'H@3', 'H@10')], g["filtered"], rtol=1e-12)


def test_negative_sampler_matches_reference_stream():
    g = np.load(os.path.join(GOLDEN, "negsample_golden.npz"))
    np.random.seed(123)   # same numpy global stream, same draw order (binomial, then randint)
    idx, lab = auxilliaries.NegativeSampler(3, 50).transform(g["batch"])
    np.testing.assert_array_equal(idx, g["idx"])
    np.testing.assert_array_equal(lab, g["labels"])
    assert idx.dtype == np.int32 and lab.dtype == np.float32


def test_train_driver_host_helpers(toy, tmp_path):
    from relationprediction_b200 import train as driver
    from relationprediction_b200.common import settings_reader
    ent = {int(k): v for k, v in toy["entities"].items()}
    rel = {int(k): v for k, v in toy["relations"].items()}
    (tmp_path / "entities.dict").write_text("".join("%d\t%s\n" % kv for kv in sorted(ent.items())))
    (tmp_path / "relations.dict").write_text("".join("%d\t%s\n" % kv for kv in sorted(rel.items())))
    for split in ("train", "valid", "test"):
        (tmp_path / (split + ".txt")).write_text(
            "".join("%s\t%s\t%s\n" % (ent[s], rel[r], ent[o]) for s, r, o in toy[split]))
    splits, entities, relations = driver.load_dataset(str(tmp_path))
    assert splits["train"].tolist() == toy["train"] and len(entities) == 16 and len(relations) == 9
    # section merge exactly like train.py:69-86
    s = driver.merge_settings(settings_reader.read_string(toy["settings_text"]["gcn_block.exp"]), 16, 9, 43)
    assert s["Encoder"]["EntityCount"] == 16 and s["Encoder"]["CodeDimension"] == "500"
    assert s["Decoder"]["NegativeSampleRate"] == "10" and s["Optimizer"]["GraphBatchSize"] == "30000"
    assert s["Evaluation"]["EdgeCount"] == 43
    # neighbourhood-expansion sampler: distinct edge ids, grows from seen vertices (train.py:161-198)
    train = splits["train"]
    adj = [[] for _ in range(16)]
    for i, (a, _, b) in enumerate(train.tolist()):
        adj[a].append((i, b))
        adj[b].append((i, a))
    deg = np.array([len(a) for a in adj])
    np.random.seed(1)
    ids = driver.sample_edge_neighborhood(adj, deg, len(train), 20)
    assert len(set(ids.tolist())) == 20 and ids.min() >= 0 and ids.max() < 43
    touched = set()
    for k, e in enumerate(ids.tolist()):
        a, _, b = train[e]
        if k > 0:
            assert a in touched or b in touched   # every later edge touches an already seen vertex
        touched.update((int(a), int(b)))


def test_library_edge_sampler_properties_and_distribution(toy):
    """csrc/sampler.cu against the reference-restated Python sampler: hard invariants on every draw and the
    first-edge distribution (fallback: uniform vertex with edges, then uniform incident edge) in closed form."""
    from relationprediction_b200 import train as driver
    train = np.array(toy["train"], dtype=np.int32)
    E, V = len(train), 16
    np.random.seed(5)
    for size in (1, 10, 43):
        ids = driver.sample_edge_neighborhood_fast(train, V, size)
        assert len(ids) == size and len(set(ids.tolist())) == size and ids.min() >= 0 and ids.max() < E
    # every edge after the first touches an already seen vertex (Toy's graph is connected)
    for _ in range(50):
        ids = driver.sample_edge_neighborhood_fast(train, V, 30)
        touched = set()
        for k, e in enumerate(ids.tolist()):
            a, _, b = train[e]
            if k:
                assert int(a) in touched or int(b) in touched
            touched.update((int(a), int(b)))
    # first-edge law: P(e) = (1/#nonisolated) * (1/deg(s) + 1/deg(o))
    deg = np.bincount(np.concatenate([train[:, 0], train[:, 2]]), minlength=V)
    p = (1.0 / deg[train[:, 0]] + 1.0 / deg[train[:, 2]]) / (deg > 0).sum()
    assert abs(p.sum() - 1) < 1e-12
    n = 40000
    firsts = np.array([driver.sample_edge_neighborhood_fast(train, V, 1)[0] for _ in range(n)])
    freq = np.bincount(firsts, minlength=E) / n
    assert np.abs(freq - p).max() < 4 * np.sqrt(p.max() / n) + 1e-3
    # same summary statistic as the Python restatement: mean number of distinct vertices in a 15-edge sample
    adj = [[] for _ in range(V)]
    for i, (a, _, b) in enumerate(train.tolist()):
        adj[a].append((i, b))
        adj[b].append((i, a))
    dg = np.array([len(a) for a in adj])

    def n_vertices(ids):
        return len(set(train[ids][:, 0].tolist()) | set(train[ids][:, 2].tolist()))
    ref = np.mean([n_vertices(driver.sample_edge_neighborhood(adj, dg, E, 15)) for _ in range(400)])
    fast = np.mean([n_vertices(driver.sample_edge_neighborhood_fast(train, V, 15)) for _ in range(4000)])
    assert abs(ref - fast) < 0.35, (ref, fast)


def test_library_edge_sampler_scale_and_self_loops():
    import time
    from relationprediction_b200 import train as driver
    from conftest import synthetic_kg
    tr = synthetic_kg(14541, 237, 272115, seed=3, skewed=True)
    tr[:50, 2] = tr[:50, 0]   # self loops
    t0 = time.perf_counter()
    ids = driver.sample_edge_neighborhood_fast(tr, 14541, 30000)
    dt = time.perf_counter() - t0
    assert len(set(ids.tolist())) == 30000 and dt < 2.0   # the reference's numpy loop takes ~5 s
    ids_all = driver.sample_edge_neighborhood_fast(tr, 14541, len(tr))   # exhausts every edge exactly once
    assert sorted(ids_all.tolist()) == list(range(len(tr)))


def test_tf_adam_and_clip_restatement_agrees_with_torch():
    """N1 oracle pin: the TF-1.x Adam / clip_by_global_norm restatement against torch's own implementations
    (they differ only in where epsilon enters, which vanishes for eps -> 0 / large norms)."""
    import torch
    from oracle import rgcn_oracle as oracle
    rng = np.random.RandomState(3)
    shapes = [(7, 5), (11,), (3, 4, 2)]
    ps = [rng.normal(size=s) for s in shapes]
    tp = [torch.nn.Parameter(torch.tensor(p, dtype=torch.float64)) for p in ps]
    ms, vs = [np.zeros(s) for s in shapes], [np.zeros(s) for s in shapes]
    opt = torch.optim.Adam(tp, lr=0.01, betas=(0.9, 0.999), eps=1e-30)
    for t in range(1, 6):
        gs = [rng.normal(size=s) * 3 for s in shapes]
        for p, g in zip(tp, gs):
            p.grad = torch.tensor(g, dtype=torch.float64)
        torch.nn.utils.clip_grad_norm_(tp, 1.0)
        cl, gn = oracle.tf_clip_by_global_norm(gs, 1.0)
        assert abs(gn - np.sqrt(sum((g ** 2).sum() for g in gs))) < 1e-12
        for p, c in zip(tp, cl):
            assert np.abs(p.grad.numpy() - c).max() < 1e-5      # torch divides by (norm + 1e-6)
            p.grad = torch.tensor(c, dtype=torch.float64)
        opt.step()
        oracle.tf_adam_step(ps, cl, ms, vs, t, eps=1e-30)
        for p, ref in zip(tp, ps):
            assert np.abs(p.detach().numpy() - ref).max() < 1e-12
    # below the threshold nothing is scaled; zero gradients stay zero
    small = [np.full((3,), 1e-3)]
    assert np.allclose(oracle.tf_clip_by_global_norm(small, 1.0)[0][0], small[0])
    assert np.all(oracle.tf_clip_by_global_norm([np.zeros(4)], 1.0)[0][0] == 0)


def test_packed_dataset_and_sample_stream(tmp_path):
    """train.py host plumbing that needs no GPU: the packed dataset form and the prefetching sample stream."""
    from relationprediction_b200 import train as T
    rng = np.random.RandomState(0)
    arrs = {k: rng.randint(0, 9, size=(n, 3)).astype(np.int32) for k, n in (("train", 40), ("valid", 5), ("test", 6))}
    p = str(tmp_path / "d.npz")
    np.savez_compressed(p, V=16, R=9, **arrs)
    splits, ents, rels = T.load_dataset_npz(p)
    assert len(ents) == 16 and len(rels) == 9
    for k in arrs:
        assert splits[k].dtype == np.int32 and np.array_equal(splits[k], arrs[k])
    counter = {"n": 0}
    lock = __import__("threading").Lock()

    def sample():
        with lock:
            counter["n"] += 1
            return counter["n"]
    for threads in (0, 3):
        counter["n"] = 0
        s = T.sample_stream(sample, threads)
        got = [next(s) for _ in range(20)]
        s.close()
        assert len(set(got)) == 20 and min(got) >= 1      # every sample delivered once, none duplicated


def test_early_stopper_follows_reference_rule(capsys):
    """optimization/shared/algorithms.py:139-161: compare with the PREVIOUS check (not the best), strict
    improvement required, drops inside the burn-in are ignored."""
    from relationprediction_b200.train import EarlyStopper
    es = EarlyStopper("2000", "6000")
    assert [es.due(i) for i in (1999, 2000, 4000, 4001)] == [False, True, True, False]
    assert es.update(2000, 0.10) is False          # first check: nothing to compare with
    assert es.update(4000, 0.09) is False          # drop inside the burn-in: ignored ...
    assert "Ignoring criterion" in capsys.readouterr().out
    assert es.update(6000, 0.095) is False         # ... and the comparison base moved to 0.09; 6000 is not > burn-in
    assert es.update(8000, 0.20) is False
    assert es.update(10000, 0.20) is True          # equal is not an improvement
    assert "Stopping criterion reached" in capsys.readouterr().out
    es2 = EarlyStopper(10)
    assert es2.update(10, 0.5) is False and es2.update(20, 0.4) is True


def test_sampler_handle_matches_one_shot_and_is_thread_compatible():
    """rgcn_sampler_create / _draw: same seed -> same sample as rgcn_sample_edge_neighborhood; concurrent draws on
    one handle do not disturb each other (the handle is read-only)."""
    import ctypes
    import threading
    from relationprediction_b200 import _lib
    from relationprediction_b200.train import EdgeNeighborhoodSampler
    rng = np.random.RandomState(4)
    V, E = 300, 4000
    tr = np.stack([rng.randint(0, V, E), rng.randint(0, 5, E), (rng.zipf(1.5, E) - 1) % V], 1).astype(np.int32)
    tr[:7, 2] = tr[:7, 0]                                    # a few self loops
    s = EdgeNeighborhoodSampler(tr, V)
    for seed, n in ((1, 500), (2, 4000), (3, 0), (4, 1)):
        one = np.empty(max(n, 1), dtype=np.int32)
        rc = _lib.load().rgcn_sample_edge_neighborhood(ctypes.c_void_p(tr.ctypes.data), E, V, n, seed,
                                                       ctypes.c_void_p(one.ctypes.data))
        assert rc == 0
        got = s.draw(n, seed=seed)
        assert np.array_equal(got, one[:n]) and len(set(got.tolist())) == n
    expect = {seed: s.draw(1500, seed=seed) for seed in range(8)}
    results = {}

    def work(seed):
        for _ in range(5):
            results[seed] = s.draw(1500, seed=seed)
    threads = [threading.Thread(target=work, args=(seed,)) for seed in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for seed in range(8):
        assert np.array_equal(results[seed], expect[seed])
    with pytest.raises(_lib.RgcnError):
        s.draw(E + 1)
    s.close()


def test_one_call_training_sample_structure_and_distribution():
    """rgcn_sampler_draw_batch: the batch is the edge-neighbourhood sample of the same seed, the graph split a
    duplicate-free subset of it, the negatives follow NegativeSampler.transform's layout (copy k of triple i at row
    (k+1)*n + i, relation kept, exactly one of subject / object replaced by a uniform entity, fair coin), labels 1/0;
    deterministic per seed; bad sizes are refused."""
    from relationprediction_b200 import _lib
    from relationprediction_b200.train import EdgeNeighborhoodSampler
    rng = np.random.RandomState(9)
    V, E, R = 400, 6000, 7
    tr = np.unique(np.stack([rng.randint(0, V, E), rng.randint(0, R, E), rng.randint(0, V, E)], 1).astype(np.int32), axis=0)
    E = len(tr)
    s = EdgeNeighborhoodSampler(tr, V)
    n, split, k = 2000, 900, 10
    gs, X, Y = s.draw_batch(n, split, k, seed=11)
    assert gs.shape == (split, 3) and X.shape == ((k + 1) * n, 3) and Y.shape == ((k + 1) * n,)
    assert np.array_equal(X[:n], tr[s.draw(n, seed=11)])            # same batch as the plain draw of that seed
    assert (Y[:n] == 1).all() and (Y[n:] == 0).all()
    batch = set(map(tuple, X[:n].tolist()))
    assert len(batch) == n
    split_rows = list(map(tuple, gs.tolist()))
    assert len(set(split_rows)) == split and all(r in batch for r in split_rows)
    pos, neg = np.tile(X[:n], (k, 1)), X[n:]
    assert np.array_equal(neg[:, 1], pos[:, 1])
    ds, do = neg[:, 0] != pos[:, 0], neg[:, 2] != pos[:, 2]
    assert not (ds & do).any()                                       # never both ends
    assert neg[:, [0, 2]].min() >= 0 and neg[:, [0, 2]].max() < V
    # fair coin (a replacement equal to the original hides 1/V of the corruptions on either side)
    assert abs(ds.mean() - 0.5) < 0.02 and abs(do.mean() - 0.5) < 0.02
    # replacement entities uniform over V: chi-square over the object-corrupted rows
    counts = np.bincount(neg[do, 2], minlength=V).astype(np.float64)
    chi2 = ((counts - counts.mean()) ** 2 / counts.mean()).sum()
    assert chi2 < V + 6 * np.sqrt(2 * V)
    g2, X2, _ = s.draw_batch(n, split, 0, seed=11)
    assert np.array_equal(g2, gs) and np.array_equal(X2, X[:n])          # deterministic per seed, independent of k
    g3, X3, Y3 = s.draw_batch(n, 0, 0, seed=5)
    assert g3.shape == (0, 3) and X3.shape == (n, 3) and (Y3 == 1).all()
    for bad in ((0, 0, 1), (E + 1, 0, 1), (10, 11, 1), (10, 5, -1)):
        with pytest.raises(_lib.RgcnError):
            s.draw_batch(*bad)
    s.close()
