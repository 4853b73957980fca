"""GPU (>= 2 devices): the 1-D node-sharded layer (halo all-to-all over NCCL) -- and the experimental
feature-sharded variant -- reproduce the single-GPU forward output, input gradient and (after the all-reduce)
weight gradients."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import synthetic_kg

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, V, R, E, d, B, out_dir, mode):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from relationprediction_b200 import parallel
        tr = synthetic_kg(V, R, E, seed=5, skewed=True)
        if mode == "feature":
            sg = p = parallel.FeatureShardedGraph(tr, V, R, rank, world, dev, B, d // B)
        elif mode == "device-plan":   # edge list on the GPU: ShardPlanDevice + rgcn_graph_create_messages_device
            sg = parallel.ShardedGraph(torch.from_numpy(tr).to(dev), V, R, rank, world, dev)
            assert isinstance(sg.plan, parallel.ShardPlanDevice) and sg.graph is None
            p = sg.plan
        else:   # "peer": rows pushed into peer-mapped halo buffers (rgcn_rows_gather); "overlapped": NCCL all-to-all
            sg = parallel.ShardedGraph(tr, V, R, rank, world, dev, overlap=(mode != "plain"),
                                       pipelined=(mode == "pipelined"),
                                       transport={"peer": "peer", "overlapped": "nccl"}.get(mode))
            assert sg.pipelined == (mode == "pipelined")
            p = sg.plan
        g = torch.Generator().manual_seed(0)
        s = d // B
        H = torch.randn(V, d, generator=g)
        dOut = torch.randn(V, d, generator=g)
        Wf = (torch.randn(R, B, s, s, generator=g) * 0.3).to(dev).requires_grad_(True)
        Wb = (torch.randn(R, B, s, s, generator=g) * 0.3).to(dev).requires_grad_(True)
        Ws = (torch.randn(d, d, generator=g) * 0.05).to(dev).requires_grad_(True)
        Hl = H[p.lo:p.hi].to(dev).requires_grad_(True)
        for step in range(2 if mode == "peer" else 1):   # peer: the second step reuses the buffers behind the barriers
            for t in (Hl, Wf, Wb, Ws):
                t.grad = None
            out = sg.block_layer(Hl, Wf, Wb, Ws, B, None, 1.0, True)
            out.backward(dOut[p.lo:p.hi].to(dev))
        if mode == "peer":
            assert sg.halo_transport() == "peer" and len(sg._peer_slots) == 1
        elif mode == "overlapped":
            assert sg.halo_transport() == "nccl"
        sg.allreduce_weight_grads([Wf, Wb, Ws])
        torch.cuda.synchronize()
        np.savez(os.path.join(out_dir, "r%d.npz" % rank), out=out.detach().cpu().numpy(), dH=Hl.grad.cpu().numpy(),
                 dWf=Wf.grad.cpu().numpy(), dWb=Wb.grad.cpu().numpy(), dWs=Ws.grad.cpu().numpy(), lo=p.lo, hi=p.hi)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["peer", "device-plan", "pipelined", "overlapped", "plain", "feature"])
@pytest.mark.parametrize("world", [2, 4])
def test_sharded_block_layer_equals_single_gpu(tmp_path, world, mode):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    from relationprediction_b200 import ops
    V, R, E, d, B = 3000, 11, 40000, 500, 100
    mp.spawn(_worker, args=(world, _free_port(), V, R, E, d, B, str(tmp_path), mode), nprocs=world, join=True)
    tr = synthetic_kg(V, R, E, seed=5, skewed=True)
    g = torch.Generator().manual_seed(0)
    s = d // B
    H = torch.randn(V, d, generator=g)
    dOut = torch.randn(V, d, generator=g)
    Wf = (torch.randn(R, B, s, s, generator=g) * 0.3).cuda().requires_grad_(True)
    Wb = (torch.randn(R, B, s, s, generator=g) * 0.3).cuda().requires_grad_(True)
    Ws = (torch.randn(d, d, generator=g) * 0.05).cuda().requires_grad_(True)
    Hc = H.cuda().requires_grad_(True)
    graph = ops.Graph(tr, V, R, device=0)
    out = ops.block_layer(Hc, Wf, Wb, Ws, graph, B, None, 1.0, True)
    out.backward(dOut.cuda())

    def rel(a, b):
        return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))
    for rank in range(world):
        z = np.load(tmp_path / ("r%d.npz" % rank))
        lo, hi = int(z["lo"]), int(z["hi"])
        assert rel(z["out"], out.detach().cpu().numpy()[lo:hi]) < 1e-5
        assert rel(z["dH"], Hc.grad.cpu().numpy()[lo:hi]) < 1e-5
        assert rel(z["dWf"], Wf.grad.cpu().numpy()) < 1e-5
        assert rel(z["dWb"], Wb.grad.cpu().numpy()) < 1e-5
        assert rel(z["dWs"], Ws.grad.cpu().numpy()) < 1e-5
