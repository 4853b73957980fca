import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from bench import synthetic_kg
from relationprediction_b200 import ops
dev = torch.device("cuda", 0)
V, R, E, d, B = 14541, 237, 272115, 500, 100
s = d // B
tr = synthetic_kg(V, R, E, seed=1234, skewed=True)
tri_pin = torch.from_numpy(tr).pin_memory()
H_pin = torch.randn(V, d).pin_memory(); dOut_pin = torch.randn(V, d).pin_memory()
out_host = torch.empty(V, d).pin_memory(); dH_host = torch.empty(V, d).pin_memory()
Wf = (torch.randn(R, B, s, s, device=dev) * 0.2).requires_grad_(True)
Wb = (torch.randn(R, B, s, s, device=dev) * 0.2).requires_grad_(True)
Ws = (torch.randn(d, d, device=dev) * 0.05).requires_grad_(True)
dW_host = [torch.empty_like(t, device="cpu").pin_memory() for t in (Wf, Wb, Ws)]
cs = torch.cuda.Stream(device=dev)

def seq():
    t0 = time.perf_counter()
    g2 = ops.Graph(tri_pin.numpy(), V, R, device=0)
    t1 = time.perf_counter()
    h = H_pin.to(dev, non_blocking=True).requires_grad_(True); do = dOut_pin.to(dev, non_blocking=True)
    for t in (Wf, Wb, Ws): t.grad = None
    o = ops.block_layer(h, Wf, Wb, Ws, g2, B, None, 1.0, True); o.backward(do)
    out_host.copy_(o.detach(), non_blocking=True); dH_host.copy_(h.grad, non_blocking=True)
    for hh, t in zip(dW_host, (Wf, Wb, Ws)): hh.copy_(t.grad, non_blocking=True)
    torch.cuda.synchronize()
    return t1 - t0

def side_h2d_first():
    with torch.cuda.stream(cs):
        h = H_pin.to(dev, non_blocking=True); do = dOut_pin.to(dev, non_blocking=True)
    ev = cs.record_event()
    t0 = time.perf_counter()
    g2 = ops.Graph(tri_pin.numpy(), V, R, device=0)
    t1 = time.perf_counter()
    torch.cuda.current_stream().wait_event(ev)
    h.requires_grad_(True)
    for t in (Wf, Wb, Ws): t.grad = None
    o = ops.block_layer(h, Wf, Wb, Ws, g2, B, None, 1.0, True); o.backward(do)
    out_host.copy_(o.detach(), non_blocking=True); dH_host.copy_(h.grad, non_blocking=True)
    for hh, t in zip(dW_host, (Wf, Wb, Ws)): hh.copy_(t.grad, non_blocking=True)
    torch.cuda.synchronize()
    return t1 - t0

def prep_then_side():
    t0 = time.perf_counter()
    tri_dev = tri_pin.to(dev, non_blocking=True)   # tiny copy first, then the big ones on the side stream
    with torch.cuda.stream(cs):
        h = H_pin.to(dev, non_blocking=True); do = dOut_pin.to(dev, non_blocking=True)
    ev = cs.record_event()
    g2 = ops.Graph(tri_pin.numpy(), V, R, device=0)
    t1 = time.perf_counter()
    torch.cuda.current_stream().wait_event(ev)
    h.requires_grad_(True)
    for t in (Wf, Wb, Ws): t.grad = None
    o = ops.block_layer(h, Wf, Wb, Ws, g2, B, None, 1.0, True)
    fe = torch.cuda.current_stream().record_event()
    with torch.cuda.stream(cs):
        cs.wait_event(fe); out_host.copy_(o.detach(), non_blocking=True)
    o.backward(do)
    dH_host.copy_(h.grad, non_blocking=True)
    for hh, t in zip(dW_host, (Wf, Wb, Ws)): hh.copy_(t.grad, non_blocking=True)
    torch.cuda.synchronize()
    return t1 - t0

for name, fn in (("sequential", seq), ("side-stream H2D before prep", side_h2d_first), ("prep first, side H2D, out D2H overlap", prep_then_side), ("sequential again", seq)):
    fn(); fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter(); preps = [fn() for _ in range(10)]; dt = (time.perf_counter() - t0) / 10
    print("%-45s %.3f ms/step  (graph prep call %.3f ms)  -> %.1f M-edges/s" % (name, dt * 1e3, np.mean(preps) * 1e3, E / dt / 1e6))
# raw copy bandwidth
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): a = H_pin.to(dev, non_blocking=True)
torch.cuda.synchronize(); print("H2D 29MB pinned: %.3f ms" % ((time.perf_counter() - t0) / 10 * 1e3))
t0 = time.perf_counter()
for _ in range(10): out_host.copy_(a, non_blocking=True)
torch.cuda.synchronize(); print("D2H 29MB pinned: %.3f ms" % ((time.perf_counter() - t0) / 10 * 1e3))
