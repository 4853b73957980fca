"""Host-side operators over the C-ABI: graph handle + autograd Functions.

PyTorch is used for storage (CUDA tensors), streams and autograd bookkeeping only: every
forward/backward below is ONE call into librgcn_b200.so (include/rgcn_b200.h).  There is no CPU or
eager-torch fallback: tensors that are not CUDA fp32 raise.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib

_NORM = {"canonical": _lib.RGCN_NORM_CANONICAL, "explicit": _lib.RGCN_NORM_EXPLICIT,
         "none": _lib.RGCN_NORM_NONE}


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _np_ptr(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else ctypes.c_void_p(0)


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _check_cuda_f32(name, t, shape=None):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32):
        raise _lib.RgcnError("%s must be a CUDA float32 tensor (no CPU fallback exists)" % name)
    if not t.is_contiguous():
        raise _lib.RgcnError("%s must be contiguous" % name)
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise _lib.RgcnError("%s has shape %s, expected %s" % (name, tuple(t.shape), tuple(shape)))


class Graph:
    """Prepared message-passing graph (opaque C handle).

    Replaces Representation/MessageGraph (extras/graph_representations.py): triples [E,3] (s,r,o)
    -> 2E messages, per-direction 1/in-degree normalisation, three sorted views + warp work lists.
    `device=None` builds the host structure only (used by the CPU tests of the index work).
    """

    def __init__(self, triples, n_entities, n_relations, norm_mode="canonical", norm_f=None,
                 norm_b=None, device=None, _handle=None):
        self._lib = _lib.load()
        self._h = ctypes.c_void_p(0)
        self.device = device
        if _handle is not None:
            self._h = _handle
        else:
            tri = np.ascontiguousarray(np.asarray(triples, dtype=np.int32).reshape(-1, 3))
            nf = None if norm_f is None else np.ascontiguousarray(norm_f, dtype=np.float32)
            nb = None if norm_b is None else np.ascontiguousarray(norm_b, dtype=np.float32)
            dev = -1 if device is None else int(device)
            st = _stream(device) if device is not None else ctypes.c_void_p(0)
            rc = self._lib.rgcn_graph_create(_np_ptr(tri), tri.shape[0], int(n_entities),
                                             int(n_relations), _NORM[norm_mode], _np_ptr(nf),
                                             _np_ptr(nb), dev, st, ctypes.byref(self._h))
            _lib.check(rc, "rgcn_graph_create")
        info = self.info()
        self.M, self.V_dst, self.V_src, self.n_relw = info[0], info[1], info[2], info[3]

    @classmethod
    def from_messages(cls, dst, src, relw, norm, V_dst, V_src, n_relw, device=None):
        lib = _lib.load()
        dst = np.ascontiguousarray(dst, dtype=np.int32)
        src = np.ascontiguousarray(src, dtype=np.int32)
        relw = np.ascontiguousarray(relw, dtype=np.int32)
        norm = np.ascontiguousarray(norm, dtype=np.float32)
        h = ctypes.c_void_p(0)
        dev = -1 if device is None else int(device)
        st = _stream(device) if device is not None else ctypes.c_void_p(0)
        rc = lib.rgcn_graph_create_messages(_np_ptr(dst), _np_ptr(src), _np_ptr(relw), _np_ptr(norm),
                                            dst.shape[0], int(V_dst), int(V_src), int(n_relw), dev,
                                            st, ctypes.byref(h))
        _lib.check(rc, "rgcn_graph_create_messages")
        return cls(None, 0, 0, device=device, _handle=h)

    @staticmethod
    def _dev_i32(name, t, n=None):
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.int32 and t.is_contiguous()):
            raise _lib.RgcnError("%s must be a contiguous CUDA int32 tensor" % name)
        if n is not None and t.numel() != n:
            raise _lib.RgcnError("%s has %d elements, expected %d" % (name, t.numel(), n))
        return t

    @classmethod
    def from_device_triples(cls, triples, n_entities, n_relations, norm_mode="canonical", norm_f=None, norm_b=None):
        """Graph from an int32 [E,3] (s,r,o) tensor that already lives on the GPU (rgcn_graph_create_device):
        the edge list never visits the host."""
        lib = _lib.load()
        tri = cls._dev_i32("triples", triples)
        if tri.dim() != 2 or tri.shape[1] != 3:
            raise _lib.RgcnError("triples must be [E,3]")
        for nm, t in (("norm_f", norm_f), ("norm_b", norm_b)):
            if t is not None:
                _check_cuda_f32(nm, t, (tri.shape[0],))
        dev = tri.device.index if tri.device.index is not None else torch.cuda.current_device()
        h = ctypes.c_void_p(0)
        rc = lib.rgcn_graph_create_device(_ptr(tri), tri.shape[0], int(n_entities), int(n_relations),
                                          _NORM[norm_mode], _ptr(norm_f), _ptr(norm_b), dev, _stream(tri.device),
                                          ctypes.byref(h))
        _lib.check(rc, "rgcn_graph_create_device")
        return cls(None, 0, 0, device=dev, _handle=h)

    @classmethod
    def from_device_messages(cls, dst, src, relw, norm, V_dst, V_src, n_relw):
        """Graph from message arrays resident on the GPU (rgcn_graph_create_messages_device)."""
        lib = _lib.load()
        M = dst.numel()
        cls._dev_i32("dst", dst)
        cls._dev_i32("src", src, M)
        cls._dev_i32("relw", relw, M)
        _check_cuda_f32("norm", norm, (M,))
        dev = dst.device.index if dst.device.index is not None else torch.cuda.current_device()
        h = ctypes.c_void_p(0)
        rc = lib.rgcn_graph_create_messages_device(_ptr(dst), _ptr(src), _ptr(relw), _ptr(norm), M, int(V_dst),
                                                   int(V_src), int(n_relw), dev, _stream(dst.device),
                                                   ctypes.byref(h))
        _lib.check(rc, "rgcn_graph_create_messages_device")
        return cls(None, 0, 0, device=dev, _handle=h)

    def info(self):
        arr = (ctypes.c_int64 * 16)()
        _lib.check(self._lib.rgcn_graph_info(self._h, arr), "rgcn_graph_info")
        return [int(x) for x in arr]

    def export(self, which):
        n = self._lib.rgcn_graph_export_bytes(self._h, which)
        if n < 0:
            _lib.check(int(n), "rgcn_graph_export_bytes")
        dtype = np.float32 if which in (_lib.X_DST_NORM, _lib.X_SRC_NORM, _lib.X_REL_NORM,
                                        _lib.X_MSG_NORM, _lib.X_REL2_NORM) else np.int32
        out = np.empty(n // 4, dtype=dtype)
        _lib.check(self._lib.rgcn_graph_export(self._h, which, _np_ptr(out), n), "rgcn_graph_export")
        return out

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            if self.device is not None and os.environ.get("RGCN_ASYNC_FREE") == "1":
                # opt-in (not yet validated on a GPU): stream-ordered frees on the current stream, no device sync
                self._lib.rgcn_graph_destroy_async(self._h, _stream(self.device))
            else:
                self._lib.rgcn_graph_destroy(self._h)
            self._h = ctypes.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h


# ---- tf.clip_by_global_norm sees IndexedSlices -----------------------------------------------------------------
# Variables the reference reads through tf.nn.embedding_lookup (block tables, relation table) get un-aggregated
# per-edge / per-triple gradient slices, and the global clipping norm is taken over those slice values.  When enabled,
# the backward passes below also produce sum |slice|^2 on the device and park it on the parameter tensor as
# `_slice_sumsq`; optim.ClippedAdam uses it in place of the dense gradient's sum of squares.
_SLICE_NORMS = False


def set_slice_norms(enabled):
    global _SLICE_NORMS
    _SLICE_NORMS = bool(enabled)


def _add_slice_sumsq(param, value):
    prev = getattr(param, "_slice_sumsq", None)
    param._slice_sumsq = value if prev is None else prev + value


def _workspace(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def _mask_arg(mask, V, d):
    if mask is None:
        return None
    if not (mask.is_cuda and mask.dtype == torch.uint8 and mask.is_contiguous()
            and tuple(mask.shape) == (V, d)):
        raise _lib.RgcnError("drop_mask must be a contiguous CUDA uint8 [V_dst, d] keep-mask")
    return mask


class _BlockLayerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, H, Wf, Wb, Wself, graph, n_blocks, drop_mask, keep, relu):
        lib = _lib.load()
        d = H.shape[1]
        R = graph.n_relw // 2
        s = d // n_blocks
        _check_cuda_f32("H", H, (graph.V_src, d))
        _check_cuda_f32("W_forward", Wf, (R, n_blocks, s, s))
        _check_cuda_f32("W_backward", Wb, (R, n_blocks, s, s))
        _check_cuda_f32("W_self", Wself, (d, d))
        mask = _mask_arg(drop_mask, graph.V_dst, d)
        dev = H.device
        out = torch.empty(graph.V_dst, d, dtype=torch.float32, device=dev)
        nb = lib.rgcn_block_workspace_bytes(graph.handle, d, n_blocks, 0)
        if nb < 0:
            _lib.check(int(nb), "rgcn_block_workspace_bytes")
        ws = _workspace(nb, dev)
        rc = lib.rgcn_block_forward(graph.handle, d, n_blocks, _ptr(H), _ptr(Wf), _ptr(Wb),
                                    _ptr(Wself), _ptr(mask), float(keep), int(bool(relu)), _ptr(out),
                                    _ptr(ws), ws.numel(), _stream(dev))
        _lib.check(rc, "rgcn_block_forward")
        ctx.graph, ctx.n_blocks, ctx.keep, ctx.relu = graph, n_blocks, float(keep), bool(relu)
        ctx.mask = mask
        ctx.params = (Wf, Wb)   # the caller's tensor objects (slice norms are parked on them)
        ctx.save_for_backward(H, Wf, Wb, Wself, out)
        return out

    @staticmethod
    def backward(ctx, dOut):
        lib = _lib.load()
        H, Wf, Wb, Wself, out = ctx.saved_tensors
        graph, B = ctx.graph, ctx.n_blocks
        d = H.shape[1]
        dOut = dOut.contiguous()
        _check_cuda_f32("dOut", dOut, (graph.V_dst, d))
        dev = H.device
        dH = torch.empty_like(H)
        dWf, dWb, dWself = torch.empty_like(Wf), torch.empty_like(Wb), torch.empty_like(Wself)
        nb = lib.rgcn_block_workspace_bytes(graph.handle, d, B, 1)
        ws = _workspace(nb, dev)
        rc = lib.rgcn_block_backward(graph.handle, d, B, _ptr(H), _ptr(Wf), _ptr(Wb), _ptr(Wself),
                                     _ptr(ctx.mask), ctx.keep, int(ctx.relu), _ptr(out), _ptr(dOut),
                                     _ptr(dH), _ptr(dWf), _ptr(dWb), _ptr(dWself), _ptr(ws),
                                     ws.numel(), _stream(dev))
        _lib.check(rc, "rgcn_block_backward")
        if _SLICE_NORMS:
            G = (dOut * (out > 0)) if ctx.relu else dOut
            nb2 = lib.rgcn_block_slice_sumsq_workspace_bytes(graph.handle, d, B)
            ws2 = _workspace(nb2, dev)
            ss = torch.empty(2, dtype=torch.float32, device=dev)
            rc = lib.rgcn_block_slice_sumsq(graph.handle, d, B, _ptr(H), _ptr(G.contiguous()), _ptr(ss), _ptr(ws2),
                                            ws2.numel(), _stream(dev))
            _lib.check(rc, "rgcn_block_slice_sumsq")
            _add_slice_sumsq(ctx.params[0], ss[0])
            _add_slice_sumsq(ctx.params[1], ss[1])
        return dH, dWf, dWb, dWself, None, None, None, None, None


def block_layer(H, W_forward, W_backward, W_self, graph, n_blocks, drop_mask=None, keep=1.0,
                relu=True):
    """Block-diagonal R-GCN layer (ConcatGcn, gcn_basis_concat.py:35-83), differentiable."""
    return _BlockLayerFn.apply(H, W_forward, W_backward, W_self, graph, int(n_blocks), drop_mask,
                               keep, relu)


class _BasisLayerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, H, Vf, Vb, Cf, Cb, Wself, graph, drop_mask, keep, relu):
        lib = _lib.load()
        d = H.shape[1]
        R = graph.n_relw // 2
        B = Cf.shape[1]
        _check_cuda_f32("H", H, (graph.V_src, d))
        _check_cuda_f32("W_forward", Vf, (d, B, d))
        _check_cuda_f32("W_backward", Vb, (d, B, d))
        _check_cuda_f32("C_forward", Cf, (R, B))
        _check_cuda_f32("C_backward", Cb, (R, B))
        _check_cuda_f32("W_self", Wself, (d, d))
        mask = _mask_arg(drop_mask, graph.V_dst, d)
        dev = H.device
        out = torch.empty(graph.V_dst, d, dtype=torch.float32, device=dev)
        saved = torch.empty(graph.V_dst, 2 * d * B, dtype=torch.float32, device=dev)
        nb = lib.rgcn_basis_workspace_bytes(graph.handle, d, B, 0)
        if nb < 0:
            _lib.check(int(nb), "rgcn_basis_workspace_bytes")
        ws = _workspace(nb, dev)
        rc = lib.rgcn_basis_forward(graph.handle, d, B, _ptr(H), _ptr(Vf), _ptr(Vb), _ptr(Cf),
                                    _ptr(Cb), _ptr(Wself), _ptr(mask), float(keep), int(bool(relu)),
                                    _ptr(out), _ptr(saved), _ptr(ws), ws.numel(), _stream(dev))
        _lib.check(rc, "rgcn_basis_forward")
        ctx.graph, ctx.keep, ctx.relu, ctx.mask = graph, float(keep), bool(relu), mask
        ctx.save_for_backward(H, Vf, Vb, Cf, Cb, Wself, out, saved)
        return out

    @staticmethod
    def backward(ctx, dOut):
        lib = _lib.load()
        H, Vf, Vb, Cf, Cb, Wself, out, saved = ctx.saved_tensors
        graph = ctx.graph
        d, B = H.shape[1], Cf.shape[1]
        dOut = dOut.contiguous()
        dev = H.device
        dH = torch.empty_like(H)
        dVf, dVb = torch.empty_like(Vf), torch.empty_like(Vb)
        dCf, dCb, dWself = torch.empty_like(Cf), torch.empty_like(Cb), torch.empty_like(Wself)
        nb = lib.rgcn_basis_workspace_bytes(graph.handle, d, B, 1)
        ws = _workspace(nb, dev)
        rc = lib.rgcn_basis_backward(graph.handle, d, B, _ptr(H), _ptr(Vf), _ptr(Vb), _ptr(Cf),
                                     _ptr(Cb), _ptr(Wself), _ptr(ctx.mask), ctx.keep, int(ctx.relu),
                                     _ptr(out), _ptr(saved), _ptr(dOut), _ptr(dH), _ptr(dVf),
                                     _ptr(dVb), _ptr(dCf), _ptr(dCb), _ptr(dWself), _ptr(ws),
                                     ws.numel(), _stream(dev))
        _lib.check(rc, "rgcn_basis_backward")
        return dH, dVf, dVb, dCf, dCb, dWself, None, None, None, None


def basis_layer(H, W_forward, W_backward, C_forward, C_backward, W_self, graph, drop_mask=None,
                keep=1.0, relu=True):
    """Basis-decomposition R-GCN layer (BasisGcn, gcn_basis.py:39-88), differentiable."""
    return _BasisLayerFn.apply(H, W_forward, W_backward, C_forward, C_backward, W_self, graph,
                               drop_mask, keep, relu)


class _DistMultFn(torch.autograd.Function):
    """Returns (energies[N], loss, reg): loss = mean sigmoid-CE (0 if Y is None), reg = un-scaled L2."""

    @staticmethod
    def forward(ctx, codes, rel, X, Y):
        lib = _lib.load()
        _check_cuda_f32("codes", codes)
        _check_cuda_f32("relation table", rel)
        if not (X.is_cuda and X.dtype == torch.int32 and X.is_contiguous() and X.dim() == 2
                and X.shape[1] == 3):
            raise _lib.RgcnError("X must be a contiguous CUDA int32 [N,3] tensor")
        if Y is not None:
            _check_cuda_f32("Y", Y, (X.shape[0],))
        V, d = codes.shape
        dev = codes.device
        N = X.shape[0]
        energies = torch.empty(N, dtype=torch.float32, device=dev)
        loss2 = torch.empty(2, dtype=torch.float32, device=dev)
        rc = lib.distmult_forward(_ptr(codes), _ptr(rel), V, rel.shape[0], d, _ptr(X), N, _ptr(Y),
                                  _ptr(energies), _ptr(loss2), _stream(dev))
        _lib.check(rc, "distmult_forward")
        ctx.has_y = Y is not None
        ctx.rel_param = rel
        ctx.save_for_backward(codes, rel, X, Y if Y is not None else torch.empty(0, device=dev),
                              energies)
        return energies, loss2[0], loss2[1]

    @staticmethod
    def backward(ctx, g_energy, g_loss, g_reg):
        lib = _lib.load()
        codes, rel, X, Y, energies = ctx.saved_tensors
        Y = Y if ctx.has_y else None
        V, d = codes.shape
        dev = codes.device
        dcodes = torch.zeros_like(codes)
        drel = torch.zeros_like(rel)
        ge = None
        if g_energy is not None:
            ge = g_energy.contiguous()
        # upstream scalar gradients stay on the device (no host sync): passed as g_scale_dev[2]
        gs = torch.zeros(2, dtype=torch.float32, device=dev)
        if g_loss is not None:
            gs[0] = g_loss
        if g_reg is not None:
            gs[1] = g_reg
        ss = torch.zeros(1, dtype=torch.float32, device=dev) if _SLICE_NORMS else None
        rc = lib.distmult_backward_slices(_ptr(codes), _ptr(rel), V, rel.shape[0], d, _ptr(X), X.shape[0],
                                          _ptr(Y), _ptr(energies), 1.0, 1.0, _ptr(gs), _ptr(ge),
                                          _ptr(dcodes), _ptr(drel), _ptr(ss), _stream(dev))
        _lib.check(rc, "distmult_backward_slices")
        if ss is not None:
            _add_slice_sumsq(ctx.rel_param, ss[0])
        return dcodes, drel, None, None


def distmult(codes, rel, X, Y=None):
    """DistMult energies + sigmoid cross-entropy + L2 term (bilinear_diag.py:14-34,63-69)."""
    return _DistMultFn.apply(codes, rel, X, Y)


def gemm_tf32x3(A, B, b_is_nk=False, out=None, accumulate=False):
    """C = A @ B (B [K,N]) or A @ B.T (B [N,K], b_is_nk=True) on the tcgen05 tensor cores with the
    3xTF32 split (fp32-level accuracy).  Thin wrapper over rgcn_gemm_tf32x3 (include/rgcn_b200.h)."""
    lib = _lib.load()
    _check_cuda_f32("A", A)
    _check_cuda_f32("B", B)
    M, K = A.shape
    N = B.shape[0] if b_is_nk else B.shape[1]
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32, device=A.device)
    ws = torch.empty(2 * N * K, dtype=torch.float32, device=A.device)
    rc = lib.rgcn_gemm_tf32x3(_ptr(A), A.stride(0), _ptr(B), B.stride(0), int(b_is_nk), _ptr(out),
                              out.stride(0), M, N, K, int(accumulate), _ptr(ws), ws.numel() * 4,
                              _stream(A.device))
    _lib.check(rc, "rgcn_gemm_tf32x3")
    return out


def gemm_tn_tf32x3(A, B, out=None, accumulate=False):
    """C = A.T @ B with A [K,M], B [K,N] (contraction over the slow dimension), tcgen05 3xTF32, split-K."""
    lib = _lib.load()
    _check_cuda_f32("A", A)
    _check_cuda_f32("B", B)
    K, M = A.shape
    N = B.shape[1]
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32, device=A.device)
    rc = lib.rgcn_gemm_tn_tf32x3(_ptr(A), A.stride(0), _ptr(B), B.stride(0), _ptr(out), out.stride(0), M, N, K,
                                 int(accumulate), _stream(A.device))
    _lib.check(rc, "rgcn_gemm_tn_tf32x3")
    return out


def rows_add_(dst, rows, src):
    """dst[rows[i]] += src[i] for UNIQUE rows (rgcn_rows_add): the non-atomic unpack of one peer's halo gradients."""
    lib = _lib.load()
    _check_cuda_f32("dst", dst)
    _check_cuda_f32("src", src)
    if not (rows.is_cuda and rows.dtype == torch.int64 and rows.is_contiguous() and rows.numel() == src.shape[0]):
        raise _lib.RgcnError("rows must be a contiguous CUDA int64 tensor with one entry per row of src")
    _lib.check(lib.rgcn_rows_add(_ptr(dst), _ptr(rows), _ptr(src), src.shape[0], dst.shape[1], _stream(dst.device)),
               "rgcn_rows_add")
    return dst


def relu_backward(dOut, out):
    """G = dOut * (out > 0) in one pass (rgcn_relu_backward)."""
    lib = _lib.load()
    dOut = dOut.contiguous()
    _check_cuda_f32("dOut", dOut)
    _check_cuda_f32("out", out, dOut.shape)
    G = torch.empty_like(dOut)
    _lib.check(lib.rgcn_relu_backward(_ptr(dOut), _ptr(out), _ptr(G), dOut.numel(), _stream(dOut.device)),
               "rgcn_relu_backward")
    return G


def rows_gather_to(dst_ptr, src, rows, max_ctas=0):
    """memory at dst_ptr [len(rows), d] = src[rows] (rgcn_rows_gather).  dst_ptr is a raw device address, normally a
    PEER GPU's halo buffer mapped into this process (parallel.PeerHalo): the halo push of the node-sharded path."""
    lib = _lib.load()
    _check_cuda_f32("src", src)
    if not (rows.is_cuda and rows.dtype == torch.int64 and rows.is_contiguous()):
        raise _lib.RgcnError("rows must be a contiguous CUDA int64 tensor")
    _lib.check(lib.rgcn_rows_gather(int(dst_ptr), _ptr(src), _ptr(rows), rows.numel(), src.shape[1], int(max_ctas),
                                    _stream(src.device)), "rgcn_rows_gather")


def block_aggregate_(out, X, W_forward, W_backward, graph, n_blocks):
    """out[dst] += sum_m norm_m W[relw_m] . X[src_m] (messages only, in place; rgcn_block_aggregate)."""
    lib = _lib.load()
    d = X.shape[1]
    _check_cuda_f32("X", X, (graph.V_src, d))
    _check_cuda_f32("out", out, (graph.V_dst, d))
    nb = lib.rgcn_block_aggregate_workspace_bytes(graph.handle, d, n_blocks, 0)
    ws = _workspace(nb, X.device)
    rc = lib.rgcn_block_aggregate(graph.handle, d, n_blocks, _ptr(X), _ptr(W_forward), _ptr(W_backward),
                                  _ptr(out), _ptr(ws), ws.numel(), _stream(X.device))
    _lib.check(rc, "rgcn_block_aggregate")
    return out


def block_aggregate_backward(X, W_forward, W_backward, G, graph, n_blocks, dWf=None, dWb=None):
    """Returns (dX, dWf, dWb) of block_aggregate_; when dWf/dWb are given they are accumulated into."""
    lib = _lib.load()
    d = X.shape[1]
    _check_cuda_f32("X", X, (graph.V_src, d))
    _check_cuda_f32("G", G, (graph.V_dst, d))
    acc = dWf is not None
    if not acc:
        dWf, dWb = torch.empty_like(W_forward), torch.empty_like(W_backward)
    dX = torch.empty_like(X)
    nb = lib.rgcn_block_aggregate_workspace_bytes(graph.handle, d, n_blocks, 1)
    ws = _workspace(nb, X.device)
    rc = lib.rgcn_block_aggregate_backward(graph.handle, d, n_blocks, _ptr(X), _ptr(W_forward), _ptr(W_backward),
                                           _ptr(G), _ptr(dX), _ptr(dWf), _ptr(dWb), int(acc), _ptr(ws),
                                           ws.numel(), _stream(X.device))
    _lib.check(rc, "rgcn_block_aggregate_backward")
    return dX, dWf, dWb


class DistMultRanker(object):
    """Fused all-entity scoring + ranking over one entity code matrix (distmult_rank, include/rgcn_b200.h): the
    hi/lo split of `codes` is made once and reused by every chunk / corruption side."""

    def __init__(self, codes, rel):
        _check_cuda_f32("codes", codes)
        _check_cuda_f32("relation table", rel)
        self.codes, self.rel = codes, rel
        self._ws, self._ws_n, self._split_ready = None, -1, False

    def rank(self, X, side, known_mask=None):
        """X int32 [n,3] CUDA; side 0 = subjects corrupted, 1 = objects; known_mask uint32 [n, ceil(V/32)] CUDA or None.
        Returns (raw_rank, filtered_rank or None) int32 CUDA tensors."""
        lib = _lib.load()
        V, d = self.codes.shape
        if not (X.is_cuda and X.dtype == torch.int32 and X.is_contiguous() and X.dim() == 2 and X.shape[1] == 3):
            raise _lib.RgcnError("X must be a contiguous CUDA int32 [n,3] tensor")
        n = X.shape[0]
        words = (V + 31) // 32
        if known_mask is not None and not (known_mask.is_cuda and known_mask.dtype == torch.int32
                                           and known_mask.is_contiguous() and tuple(known_mask.shape) == (n, words)):
            raise _lib.RgcnError("known_mask must be a contiguous CUDA int32 [n, ceil(V/32)] tensor (bit masks)")
        dev = self.codes.device
        if self._ws is None or n > self._ws_n:
            nb = lib.distmult_rank_workspace_bytes(V, d, n)
            if nb < 0:
                _lib.check(int(nb), "distmult_rank_workspace_bytes")
            self._ws, self._ws_n, self._split_ready = _workspace(nb, dev), n, False
        raw = torch.empty(n, dtype=torch.int32, device=dev)
        filt = torch.empty(n, dtype=torch.int32, device=dev) if known_mask is not None else None
        rc = lib.distmult_rank(_ptr(self.codes), _ptr(self.rel), V, self.rel.shape[0], d, _ptr(X), n, int(side),
                               _ptr(known_mask), int(self._split_ready), _ptr(raw), _ptr(filt), _ptr(self._ws),
                               self._ws.numel(), _stream(dev))
        _lib.check(rc, "distmult_rank")
        self._split_ready = True
        return raw, filt
