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
