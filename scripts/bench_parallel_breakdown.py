"""torchrun --nproc-per-node N scripts/bench_parallel_breakdown.py : where does a sharded step go?"""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, ".")
from bench import synthetic_kg
from relationprediction_b200 import ops, parallel

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
V, R, E, d, B = 14541 * world, 237, 272115 * world, 500, 100
tr = synthetic_kg(V, R, E, seed=1234, skewed=True)
sg = parallel.ShardedGraph(tr, V, R, rank, world, dev)
p = sg.plan
s = d // B
g = torch.Generator(device=dev).manual_seed(1)
H = torch.randn(p.n_local, d, device=dev, generator=g)
dOut = torch.randn(p.n_local, d, device=dev, generator=g)
Wf = (torch.randn(R, B, s, s, device=dev, generator=g) * 0.2).requires_grad_(True)
Wb = (torch.randn(R, B, s, s, device=dev, generator=g) * 0.2).requires_grad_(True)
Ws = (torch.randn(d, d, device=dev, generator=g) * 0.05).requires_grad_(True)

def T(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

H_ext = torch.empty(p.n_local + p.n_halo, d, device=dev)
send = H.index_select(0, sg.send_rows)
res = {"n_local": p.n_local, "n_halo": p.n_halo, "send_rows": int(sg.send_rows.numel()), "msgs": int(p.msg_dst.shape[0])}
res["index_select"] = T(lambda: H.index_select(0, sg.send_rows))
res["copy_local"] = T(lambda: H_ext[:p.n_local].copy_(H))
res["all_to_all_fwd"] = T(lambda: dist.all_to_all_single(H_ext[p.n_local:], send, p.recv_counts.tolist(), p.send_counts.tolist()))
Hx = H_ext.clone().requires_grad_(True)
res["layer_fwd"] = T(lambda: ops.block_layer(Hx.detach(), Wf.detach(), Wb.detach(), Ws.detach(), sg.graph, B, None, 1.0, True))
def fb():
    Hx.grad = None; Wf.grad = None; Wb.grad = None; Ws.grad = None
    ops.block_layer(Hx, Wf, Wb, Ws, sg.graph, B, None, 1.0, True).backward(dOut)
res["layer_fwd_bwd"] = T(fb)
back = torch.empty(int(p.send_counts.sum()), d, device=dev)
dH_ext = torch.randn(p.n_local + p.n_halo, d, device=dev)
res["all_to_all_bwd"] = T(lambda: dist.all_to_all_single(back, dH_ext[p.n_local:].contiguous(), p.send_counts.tolist(), p.recv_counts.tolist()))
dHl = dH_ext[:p.n_local].clone()
res["index_add"] = T(lambda: dHl.index_add_(0, sg.send_rows, back))
def ar():
    sg.allreduce_weight_grads([Wf, Wb, Ws])
fb()
res["allreduce"] = T(ar)
Hl = H.clone().requires_grad_(True)
def full():
    Hl.grad = None; Wf.grad = None; Wb.grad = None; Ws.grad = None
    sg.block_layer(Hl, Wf, Wb, Ws, B, None, 1.0, True).backward(dOut)
    sg.allreduce_weight_grads([Wf, Wb, Ws])
res["full_step_overlapped"] = T(full)
sg.overlap = False
res["full_step_plain"] = T(full)
if rank == 0:
    print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in res.items()})
dist.destroy_process_group()
