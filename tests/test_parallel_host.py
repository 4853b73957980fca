"""CPU, world_size 2 (gloo): the 1-D node-shard plan and the halo exchange (forward rows, backward
gradient return) against a dense single-process reference.  Kernels are not involved here; the GPU
equivalence (sharded == single GPU) is in the -m gpu suite / bench --gpus N."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from relationprediction_b200 import parallel
from conftest import synthetic_kg


def test_shard_plans_partition_the_global_message_list():
    V, R, E = 1000, 7, 9000
    tr = synthetic_kg(V, R, E, seed=5, skewed=True)
    dst, src, relw, norm = parallel.global_messages(tr, V, R)
    for world in (1, 2, 3, 8):
        seen = np.zeros(2 * E, dtype=np.int64)
        for rank in range(world):
            p = parallel.ShardPlan(tr, V, R, rank, world)
            assert p.n_local == p.hi - p.lo and p.bounds[0] == 0 and p.bounds[-1] == V
            seen[p.msg_global_id] += 1
            # every local message maps back to its global (dst, src, relw, norm)
            ext = np.concatenate([np.arange(p.lo, p.hi), p.halo_nodes])
            np.testing.assert_array_equal(p.msg_dst + p.lo, dst[p.msg_global_id])
            np.testing.assert_array_equal(ext[p.msg_src], src[p.msg_global_id])
            np.testing.assert_array_equal(p.msg_relw, relw[p.msg_global_id])
            np.testing.assert_array_equal(p.msg_norm, norm[p.msg_global_id])
            assert ((p.halo_nodes < p.lo) | (p.halo_nodes >= p.hi)).all()
            assert p.recv_counts.sum() == p.n_halo and p.recv_counts[rank] == 0
        assert (seen == 1).all()          # each message lives on exactly one rank (its destination's)
        # what rank a sends to rank b is exactly what b expects from a
        plans = [parallel.ShardPlan(tr, V, R, r, world) for r in range(world)]
        for a in range(world):
            off = 0
            for b in range(world):
                n = int(plans[a].send_counts[b])
                rows = plans[a].send_rows[off:off + n] + plans[a].lo
                off += n
                assert n == int(plans[b].recv_counts[a])
                start = int(plans[b].recv_counts[:a].sum())
                np.testing.assert_array_equal(rows, plans[b].halo_nodes[start:start + n])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, V, R, E, d, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        tr = synthetic_kg(V, R, E, seed=5, skewed=True)
        sg = parallel.ShardedGraph(tr, V, R, rank, world, "cpu")   # host-only graph handle
        p = sg.plan
        g = torch.Generator().manual_seed(0)
        H = torch.randn(V, d, generator=g)
        H_local = H[p.lo:p.hi].clone().requires_grad_(True)
        H_ext = sg.halo_exchange(H_local)
        ext = np.concatenate([np.arange(p.lo, p.hi), p.halo_nodes])
        assert torch.equal(H_ext.detach(), H[torch.as_tensor(ext)])
        # backward: weight each extended row by a global per-node coefficient on each rank
        coef = torch.arange(1, V + 1, dtype=torch.float32)[torch.as_tensor(ext)] * (rank + 1)
        (H_ext * coef[:, None]).sum().backward()
        np.save(os.path.join(out_dir, "grad_%d.npy" % rank), H_local.grad.numpy())
        np.save(os.path.join(out_dir, "ext_%d.npy" % rank), ext)
        # weight-grad all-reduce bucket
        w1 = torch.ones(3, 2, requires_grad=True)
        w2 = torch.ones(5, requires_grad=True)
        w1.grad = torch.full((3, 2), float(rank + 1))
        w2.grad = torch.full((5,), 10.0 * (rank + 1))
        sg.allreduce_weight_grads([w1, w2])
        tot = sum(range(1, world + 1))
        assert torch.equal(w1.grad, torch.full((3, 2), float(tot))) and torch.equal(w2.grad, torch.full((5,), 10.0 * tot))
    finally:
        dist.destroy_process_group()


def test_halo_exchange_world2_gloo(tmp_path):
    V, R, E, d, world = 400, 5, 3000, 12, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, V, R, E, d, str(tmp_path)), nprocs=world, join=True)
    # dense reference: d/dH[v] of sum_r sum_{rows of rank r's extended space} coef
    expect = np.zeros((V, d), np.float32)
    for rank in range(world):
        ext = np.load(tmp_path / ("ext_%d.npy" % rank))
        np.add.at(expect, ext, (np.arange(1, V + 1, dtype=np.float32)[ext] * (rank + 1))[:, None] * np.ones(d, np.float32))
    bounds = parallel.node_bounds(V, world)
    for rank in range(world):
        got = np.load(tmp_path / ("grad_%d.npy" % rank))
        np.testing.assert_allclose(got, expect[bounds[rank]:bounds[rank + 1]], rtol=1e-6)


def _ring_worker(rank, world, port, V, R, E, d, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        tr = synthetic_kg(V, R, E, seed=8, skewed=True)
        sg = parallel.ShardedGraph(tr, V, R, rank, world, "cpu", pipelined=True)
        p = sg.plan
        assert sg.pipelined and set(sg.graph_halo_peer) == {q for q in range(world) if q != rank and p.recv_counts[q]}
        # per-peer halo sub-graphs partition the remote-source messages
        assert sum(g.M for g in sg.graph_halo_peer.values()) == sg.graph_halo.M
        g = torch.Generator().manual_seed(0)
        H = torch.randn(V, d, generator=g)
        H_local = H[p.lo:p.hi].contiguous()
        send_all = H_local.index_select(0, sg.send_rows)
        H_halo = torch.empty(p.n_halo, d)
        works = parallel.ring_post_forward(sg, send_all, H_halo)
        for k in range(1, world):
            for w in works[k - 1]:
                w.wait()
            rf = (rank - k) % world
            lo, hi = sg.halo_off[rf], sg.halo_off[rf + 1]
            # after step k exactly the rows of rank me-k are guaranteed to be there
            assert torch.equal(H_halo[lo:hi], H[torch.as_tensor(p.halo_nodes[lo:hi].astype(np.int64))])
        # backward ring: gradient of halo row (global id v) coming from rank r is (r+1) * (v+1)
        back = torch.zeros(int(p.send_counts.sum()), d)
        wl, alive = [], []
        for k in range(1, world):
            rf = (rank - k) % world
            lo, hi = sg.halo_off[rf], sg.halo_off[rf + 1]
            dX = None
            if hi > lo:
                dX = ((rank + 1) * (torch.as_tensor(p.halo_nodes[lo:hi].astype(np.float32)) + 1))[:, None] * torch.ones(d)
                alive.append(dX)
            wl.append(parallel.ring_post_backward_step(sg, k, dX, back))
        for wk in wl:
            for w in wk:
                w.wait()
        dH = torch.zeros(p.n_local, d)
        dH.index_add_(0, sg.send_rows, back)
        np.save(os.path.join(out_dir, "ring_%d.npy" % rank), dH.numpy())
        np.save(os.path.join(out_dir, "halo_%d.npy" % rank), p.halo_nodes)
    finally:
        dist.destroy_process_group()


def test_pipelined_ring_exchange_world3_gloo(tmp_path):
    V, R, E, d, world = 300, 5, 2500, 6, 3
    mp.spawn(_ring_worker, args=(world, _free_port(), V, R, E, d, str(tmp_path)), nprocs=world, join=True)
    expect = np.zeros((V, d), np.float32)
    for r in range(world):
        halo = np.load(tmp_path / ("halo_%d.npy" % r))
        expect[halo] += ((r + 1) * (halo.astype(np.float32) + 1))[:, None]
    bounds = parallel.node_bounds(V, world)
    for r in range(world):
        got = np.load(tmp_path / ("ring_%d.npy" % r))
        np.testing.assert_allclose(got, expect[bounds[r]:bounds[r + 1]], rtol=1e-6)


# ---- feature-sharded block layer (experimental): host logic on gloo, kernels stubbed by the oracle -------------
class _OracleGraph(object):
    def __init__(self, triples, n_entities, n_relations, norm_mode="canonical", norm_f=None, norm_b=None,
                 device=None):
        from oracle import rgcn_oracle as oracle
        self.triples = np.asarray(triples, dtype=np.int32).reshape(-1, 3)
        self.V_dst = self.V_src = int(n_entities)
        self.M = 2 * len(self.triples)
        self.nf, self.nb = oracle.graph_norms(self.triples, n_entities, norm_mode, np.float64)


def _messages_only(X, Wf, Wb, graph):
    from oracle import rgcn_oracle as oracle
    zero_self = torch.zeros(X.shape[1], X.shape[1], dtype=X.dtype)
    return oracle.concat_gcn_forward(X, graph.triples, Wf, Wb, zero_self, graph.nf, graph.nb, None, 1.0, False,
                                     X.dtype)


def _stub_aggregate_(out, X, Wf, Wb, graph, n_blocks):
    assert Wf.shape[1] == n_blocks and X.shape == (graph.V_src, n_blocks * Wf.shape[2])
    out += _messages_only(X, Wf, Wb, graph)
    return out


def _stub_aggregate_backward(X, Wf, Wb, G, graph, n_blocks, dWf=None, dWb=None):
    with torch.enable_grad():      # called from inside an autograd.Function.backward
        Xl, Wfl, Wbl = (t.detach().clone().requires_grad_(True) for t in (X, Wf, Wb))
        gx, gf, gb = torch.autograd.grad(_messages_only(Xl, Wfl, Wbl, graph), [Xl, Wfl, Wbl], grad_outputs=G)
    return gx, gf, gb


def _feature_worker(rank, world, port, V, R, E, B, s, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from relationprediction_b200 import ops
        ops.Graph, ops.block_aggregate_, ops.block_aggregate_backward = (_OracleGraph, _stub_aggregate_,
                                                                          _stub_aggregate_backward)
        d = B * s
        tr = synthetic_kg(V, R, E, seed=3, skewed=True)
        fs = parallel.FeatureShardedGraph(tr, V, R, rank, world, "cpu", B, s)
        assert fs.d == d and fs.d_local % 4 == 0 and fs.n_halo == 0
        g = torch.Generator().manual_seed(1)
        dt = torch.float64
        H = torch.randn(V, d, generator=g, dtype=dt)
        Wf, Wb = (torch.randn(R, B, s, s, generator=g, dtype=dt).requires_grad_(True) for _ in range(2))
        Ws = torch.randn(d, d, generator=g, dtype=dt).requires_grad_(True)
        mask = (torch.rand(V, d, generator=g) < 0.8).to(torch.uint8)
        dOut = torch.randn(V, d, generator=g, dtype=dt)
        # the transposes are each other's adjoint and round-trip to the identity
        Xl = H[fs.lo:fs.hi].contiguous()
        work, Xf = fs.to_feature_async(Xl)
        work.wait()
        c0, c1 = fs.col_bounds[rank], fs.col_bounds[rank + 1]
        assert torch.equal(Xf, H[:, c0:c1])
        assert torch.equal(fs.to_node_add(Xf, torch.zeros_like(Xl)), Xl)
        H_local = Xl.clone().requires_grad_(True)
        out = fs.block_layer(H_local, Wf, Wb, Ws, B, mask[fs.lo:fs.hi], 0.8, True)
        out.backward(dOut[fs.lo:fs.hi])
        fs.allreduce_weight_grads([Wf, Wb, Ws])
        np.savez(os.path.join(out_dir, "fs_%d.npz" % rank), out=out.detach().numpy(), dH=H_local.grad.numpy(),
                 dWf=Wf.grad.numpy(), dWb=Wb.grad.numpy(), dWs=Ws.grad.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,B,s", [(2, 8, 4), (3, 6, 8), (2, 8, 5)])
def test_feature_sharded_block_layer_gloo(tmp_path, world, B, s):
    from oracle import rgcn_oracle as oracle
    V, R, E = 61, 4, 500
    d = B * s
    port = _free_port()
    mp.spawn(_feature_worker, args=(world, port, V, R, E, B, s, str(tmp_path)), nprocs=world, join=True)
    tr = synthetic_kg(V, R, E, seed=3, skewed=True)
    g = torch.Generator().manual_seed(1)
    dt = torch.float64
    H = torch.randn(V, d, generator=g, dtype=dt).requires_grad_(True)
    Wf, Wb = (torch.randn(R, B, s, s, generator=g, dtype=dt).requires_grad_(True) for _ in range(2))
    Ws = torch.randn(d, d, generator=g, dtype=dt).requires_grad_(True)
    mask = (torch.rand(V, d, generator=g) < 0.8).to(torch.uint8)
    dOut = torch.randn(V, d, generator=g, dtype=dt)
    nf, nb = oracle.graph_norms(tr, V, "canonical", np.float64)
    ref = oracle.concat_gcn_forward(H, tr, Wf, Wb, Ws, nf, nb, mask, 0.8, True, dt)
    ref.backward(dOut)
    bounds = parallel.node_bounds(V, world)
    for rank in range(world):
        z = np.load(tmp_path / ("fs_%d.npz" % rank))
        lo, hi = bounds[rank], bounds[rank + 1]
        np.testing.assert_allclose(z["out"], ref.detach().numpy()[lo:hi], rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose(z["dH"], H.grad.numpy()[lo:hi], rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose(z["dWf"], Wf.grad.numpy(), rtol=1e-10, atol=1e-10)   # summed over ranks
        np.testing.assert_allclose(z["dWb"], Wb.grad.numpy(), rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose(z["dWs"], Ws.grad.numpy(), rtol=1e-10, atol=1e-10)


def test_block_bounds_keep_rows_float4_aligned():
    assert parallel.block_bounds(100, 5, 8) == [0, 12, 24, 36, 48, 60, 72, 84, 100]
    assert parallel.block_bounds(64, 8, 8) == [0, 8, 16, 24, 32, 40, 48, 56, 64]
    for B, s, world in [(100, 5, 8), (100, 5, 3), (64, 8, 4), (12, 6, 2)]:
        b = parallel.block_bounds(B, s, world)
        assert b[0] == 0 and b[-1] == B and all(x < y for x, y in zip(b, b[1:]))
        assert all(((y - x) * s) % 4 == 0 for x, y in zip(b, b[1:]))
    with pytest.raises(ValueError):
        parallel.block_bounds(10, 5, 2)      # 10 blocks of 5 cannot be dealt in groups of 4 blocks
    with pytest.raises(ValueError):
        parallel.block_bounds(4, 8, 8)       # fewer column groups than ranks


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_device_plan_equals_host_plan(world):
    """ShardPlanDevice (torch ops; runs on whatever device holds the edge list -- here CPU tensors) reproduces the
    numpy ShardPlan element for element, explicit norms included."""
    V, R, E = 900, 6, 7000
    tr = synthetic_kg(V, R, E, seed=9, skewed=True)
    rng = np.random.RandomState(1)
    nf, nb = rng.uniform(0.1, 1, E).astype(np.float32), rng.uniform(0.1, 1, E).astype(np.float32)
    t = torch.from_numpy(tr)
    for mode, kw_h, kw_d in (("canonical", {}, {}),
                             ("explicit", dict(norm_f=nf, norm_b=nb), dict(norm_f=torch.from_numpy(nf), norm_b=torch.from_numpy(nb))),
                             ("none", {}, {})):
        for rank in range(world):
            h = parallel.ShardPlan(tr, V, R, rank, world, norm_mode=mode, **kw_h)
            dv = parallel.ShardPlanDevice(t, V, R, rank, world, norm_mode=mode, keep_global_ids=True, **kw_d)
            assert (dv.lo, dv.hi, dv.n_local, dv.n_halo) == (h.lo, h.hi, h.n_local, h.n_halo)
            for name in ("msg_dst", "msg_src", "msg_relw", "msg_norm", "msg_global_id", "halo_nodes", "send_rows"):
                np.testing.assert_array_equal(getattr(dv, name).numpy(), getattr(h, name), err_msg=name)
            np.testing.assert_array_equal(dv.recv_counts, h.recv_counts)
            np.testing.assert_array_equal(dv.send_counts, h.send_counts)


def _agree_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = [parallel.all_ranks_agree(True, None, "cpu"),                 # everyone fine
               parallel.all_ranks_agree(rank != 1, None, "cpu"),            # one rank short of memory
               parallel.all_ranks_agree(False, None, "cpu")]
        np.save(os.path.join(out_dir, "agree_%d.npy" % rank), np.array(res))
    finally:
        dist.destroy_process_group()


def test_ranks_agree_before_collective_setup(tmp_path):
    """all_ranks_agree: the peer-mapped halo buffers are only set up when EVERY rank can (a rank that cannot must take
    all the others to the NCCL transport with it, not leave them waiting in the rendezvous)."""
    world = 3
    mp.spawn(_agree_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for rank in range(world):
        assert np.load(tmp_path / ("agree_%d.npy" % rank)).tolist() == [True, False, False]
