"""Torch-tensor wrappers over the C ABI.  PyTorch is plumbing here: device memory, streams.

Every function requires CUDA tensors on an sm_100 device and raises otherwise.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import torch

from . import _lib
from ._lib import FesrError, ModelDims, ParamGrads, Params, check

_device_ok = {}


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise FesrError("fesr_b200 runs on a B200 (sm_100a) only: got a CPU tensor and there is no CPU fallback")
    dev = next(t.device for t in tensors if t is not None)
    if dev.index not in _device_ok:
        with torch.cuda.device(dev):
            check(_lib.load().fesr_device_check(), "fesr_device_check")
        _device_ok[dev.index] = True
    return dev


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _f32c(t):
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.to(torch.float32).contiguous()
    return t


class _Workspace:
    """Grow-only byte buffers per (device, tag): the caller-owned workspace the C ABI asks for."""

    def __init__(self):
        self.bufs = {}

    def get(self, dev, tag, nbytes):
        key = (dev.index, tag)
        buf = self.bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            self.bufs.pop(key, None)
            _prepared.pop(key, None)          # a new buffer holds no prepared weights
            buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=dev)
            self.bufs[key] = buf
        return buf

    def clear(self):
        self.bufs.clear()
        _prepared.clear()


# (device, workspace tag) -> (state, tensors): `state` = (workspace ptr, dims, weights generation, parameter (ptr, version)
# pairs) last prepared there; `tensors` keeps those parameter tensors alive so that the caching allocator cannot hand
# their addresses to another model while the entry says "prepared".
_prepared = {}
workspace = _Workspace()
_weights_gen = 0        # bumped by every out-of-band parameter write (fesr_adam_step works through raw pointers)


def weights_generation() -> int:
    return _weights_gen


def invalidate_prepared_weights():
    """A parameter was written without PyTorch noticing (no `_version` bump): every cached prepared copy is stale."""
    global _weights_gen
    _weights_gen += 1
    _prepared.clear()


def _prepared_state(key):
    ent = _prepared.get(key)
    return None if ent is None else ent[0]


# ---------------------------------------------------------------------------------- graph
@dataclass
class Csr:
    """Destination-sorted CSR of a (block-diagonal) graph."""
    rowptr: torch.Tensor       # [n+1] int32
    src: torch.Tensor          # [E] int32, CSR order
    perm: torch.Tensor | None  # [E] int32 original edge id per CSR slot (None: identity)
    n: int
    E: int


def csr_build(edge_index: torch.Tensor, n: int) -> Csr:
    """edge_index [2,E] int64 (any order) -> CSR by destination (fesr_csr_build)."""
    dev = _require_cuda(edge_index)
    if edge_index.dtype != torch.int64 or edge_index.dim() != 2 or edge_index.shape[0] != 2:
        raise FesrError("edge_index must be an int64 tensor of shape [2, E]")
    edge_index = edge_index.contiguous()
    E = int(edge_index.shape[1])
    lib = _lib.load()
    rowptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
    src = torch.empty(E, dtype=torch.int32, device=dev)
    perm = torch.empty(E, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        nbytes = lib.fesr_csr_workspace_bytes(n, E)
        ws = workspace.get(dev, "sort", nbytes)
        check(lib.fesr_csr_build(_ptr(edge_index), E, n, _ptr(rowptr), _ptr(src), _ptr(perm), _ptr(ws), ws.numel(),
                                 _stream(dev)), "fesr_csr_build")
    return Csr(rowptr, src, perm, n, E)


# ---------------------------------------------------------------------------------- model
def make_params(struct_cls, tensors: dict):
    """tensors: fc1_w, fc1_b, mlp_w (list), mlp_b (list), lin_w, lin_b, root, bias, fc2_w, fc2_b."""
    p = struct_cls()
    keep = []
    for name in ("fc1_w", "fc1_b", "lin_w", "lin_b", "root", "bias", "fc2_w", "fc2_b"):
        t = tensors.get(name)
        if t is not None:
            if t.dtype != torch.float32 or not t.is_contiguous() or not t.is_cuda:
                raise FesrError(f"parameter {name} must be a contiguous fp32 CUDA tensor")
            keep.append(t)
            setattr(p, name, t.data_ptr())
    for name in ("mlp_w", "mlp_b"):
        arr = getattr(p, name)
        for i, t in enumerate(tensors[name]):
            if t.dtype != torch.float32 or not t.is_contiguous() or not t.is_cuda:
                raise FesrError(f"parameter {name}[{i}] must be a contiguous fp32 CUDA tensor")
            keep.append(t)
            arr[i] = t.data_ptr()
    return p, keep


def nnconv_forward(dims: ModelDims, tensors: dict, x: torch.Tensor, csr: Csr, edge_attr: torch.Tensor,
                   precision: int = _lib.PREC_FP32, keep_for_backward: bool = False, ws_tag: str = "fwd",
                   x_ready=None):
    """Runs fesr_nnconv_forward.  Returns y [n, out_ch] (and the workspace tensor when kept).
    x_ready: a CUDA event after which `x` is valid (its host -> device copy is still running on another stream).
    The pass is then issued in two phases -- the edge MLP, which does not read x, first; the stream waits for the
    event; the rest -- so that the copy hides under the edge MLP (predict only)."""
    dev = _require_cuda(x, edge_attr, csr.rowptr)
    x = _f32c(x)
    edge_attr = _f32c(edge_attr.reshape(-1))
    n, E = csr.n, csr.E
    if x.shape[0] != n or x.shape[1] != dims.in_ch:
        raise FesrError(f"x must be [{n}, {dims.in_ch}], got {tuple(x.shape)}")
    if edge_attr.numel() != E:
        raise FesrError(f"edge_attr must have {E} entries, got {edge_attr.numel()}")
    lib = _lib.load()
    p, keep = make_params(Params, tensors)
    y = torch.empty(n, dims.out_ch, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        qflags = int(keep_for_backward)
        if keep_for_backward and precision == _lib.PREC_TF32 and os.environ.get("FESR_Z16", "1") != "0":
            qflags |= _lib.FWD_KEEP_Z16          # the tf32 arm keeps its Z stash as fp16: half the bytes per layer
        nbytes = lib.fesr_forward_workspace_bytes(C.byref(dims), n, E, qflags)
        flags = int(keep_for_backward)
        if keep_for_backward:
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        else:
            ws = workspace.get(dev, ws_tag, nbytes)
            # the prepared weight copies at the head of the cached workspace stay valid as long as the same
            # parameter tensors hold the same values (predict loops): skip the preparation kernels then
            state = (ws.data_ptr(), bytes(dims), _weights_gen, tuple((t.data_ptr(), t._version) for t in keep))
            if _prepared_state((dev.index, ws_tag)) == state:
                flags |= _lib.FWD_WEIGHTS_PREPARED
            _prepared[(dev.index, ws_tag)] = None
        if x_ready is not None and not keep_for_backward and n > 0:
            check(lib.fesr_nnconv_forward(C.byref(dims), C.byref(p), None, _ptr(csr.rowptr), _ptr(csr.src),
                                          _ptr(csr.perm), _ptr(edge_attr), n, E, precision, flags | _lib.FWD_EDGE_ONLY,
                                          None, _ptr(ws), ws.numel(), _stream(dev)), "fesr_nnconv_forward (edge phase)")
            torch.cuda.current_stream(dev).wait_event(x_ready)
            flags |= _lib.FWD_EDGE_DONE
        elif x_ready is not None:
            torch.cuda.current_stream(dev).wait_event(x_ready)
        check(lib.fesr_nnconv_forward(C.byref(dims), C.byref(p), _ptr(x), _ptr(csr.rowptr), _ptr(csr.src),
                                      _ptr(csr.perm), _ptr(edge_attr), n, E, precision, flags,
                                      _ptr(y), _ptr(ws), ws.numel(), _stream(dev)), "fesr_nnconv_forward")
        if not keep_for_backward:
            _prepared[(dev.index, ws_tag)] = (state, keep)
    del keep
    return (y, ws) if keep_for_backward else y


def overflow_flag(dims: ModelDims, dev, ws_tag: str = "fwd"):
    """int32 [1] view of the fp16 range flag of the last forward that ran in the cached workspace `ws_tag` (None if
    no forward has run there yet).  Reading its value synchronises: predict() copies it to the host together with
    the results instead and raises when they are first touched."""
    buf = workspace.bufs.get((torch.device(dev).index, ws_tag))
    if buf is None:
        return None
    off = int(_lib.load().fesr_forward_overflow_offset(C.byref(dims)))
    return buf[off:off + 4].view(torch.int32)


class ForwardPlan:
    """Everything of a predict-time forward that does not change from call to call (validated once): the ctypes
    parameter struct, the graph, the workspace size.  `run_forward_plan` then costs one tensor allocation and one
    or two C calls -- the host time before the first kernel is enqueued is GPU idle time in an end-to-end step."""
    __slots__ = ("dims", "params", "keep", "csr", "edge_attr", "n", "E", "precision", "ws_tag", "nbytes", "state_tail",
                 "dev", "edge_in")


def make_forward_plan(dims: ModelDims, tensors: dict, csr: Csr, edge_attr: torch.Tensor, precision: int,
                      ws_tag: str = "fwd") -> ForwardPlan:
    dev = _require_cuda(edge_attr, csr.rowptr)
    edge_attr = _f32c(edge_attr.reshape(-1))
    if edge_attr.numel() != csr.E:
        raise FesrError(f"edge_attr must have {csr.E} entries, got {edge_attr.numel()}")
    pl = ForwardPlan()
    pl.dims, pl.csr, pl.edge_attr, pl.n, pl.E, pl.precision, pl.ws_tag, pl.dev = dims, csr, edge_attr, csr.n, csr.E, precision, ws_tag, dev
    pl.params, pl.keep = make_params(Params, tensors)
    with torch.cuda.device(dev):
        pl.nbytes = _lib.load().fesr_forward_workspace_bytes(C.byref(dims), csr.n, csr.E, 0)
    pl.state_tail = (bytes(dims), tuple((t.data_ptr(), t._version) for t in pl.keep))
    pl.edge_in = None
    return pl


def run_forward_plan(pl: ForwardPlan, x: torch.Tensor | None, x_ready=None, edge_only: bool = False, out=None):
    """edge_only: issue just the edge phase (FESR_FWD_EDGE_ONLY: weight preparation + edge MLP, nothing that reads x)
    and return None -- the caller does so BEFORE it starts the host -> device copies of a step, so that the GPU is
    busy while the host is still issuing them; the next call on this plan then runs the rest (FESR_FWD_EDGE_DONE)."""
    dev, lib, csr = pl.dev, _lib.load(), pl.csr
    if not edge_only:
        if not x.is_cuda or x.dtype != torch.float32 or not x.is_contiguous() or x.shape != (pl.n, pl.dims.in_ch):
            raise FesrError(f"x must be a contiguous fp32 CUDA tensor of shape [{pl.n}, {pl.dims.in_ch}], got "
                            f"{tuple(x.shape)} {x.dtype} on {x.device}")
        if out is not None:
            # `out`: the caller's [n, out_ch] block (the rank's slot of the all-gather buffer) -- no copy afterwards
            if not out.is_cuda or out.dtype != torch.float32 or not out.is_contiguous() or out.shape != (pl.n, pl.dims.out_ch):
                raise FesrError(f"out must be a contiguous fp32 CUDA tensor of shape [{pl.n}, {pl.dims.out_ch}]")
            y = out
        else:
            y = torch.empty(pl.n, pl.dims.out_ch, dtype=torch.float32, device=dev)
    if pl.n == 0:
        return None if edge_only else y
    ws = workspace.get(dev, pl.ws_tag, pl.nbytes)
    key = (dev.index, pl.ws_tag)
    state = (ws.data_ptr(), _weights_gen) + pl.state_tail
    flags = _lib.FWD_WEIGHTS_PREPARED if _prepared_state(key) == state else 0
    stream = torch.cuda.current_stream(dev)
    sp = C.c_void_p(stream.cuda_stream)
    args = (C.byref(pl.dims), C.byref(pl.params))
    graph = (_ptr(csr.rowptr), _ptr(csr.src), _ptr(csr.perm), _ptr(pl.edge_attr), pl.n, pl.E, pl.precision)
    edge_in = pl.edge_in == ws.data_ptr() and flags != 0      # an edge phase of THIS plan is the last thing in the workspace
    pl.edge_in = None
    _prepared[key] = None
    with torch.cuda.device(dev):
        if edge_only or (x_ready is not None and not edge_in):
            check(lib.fesr_nnconv_forward(*args, None, *graph, flags | _lib.FWD_EDGE_ONLY, None, _ptr(ws), ws.numel(), sp),
                  "fesr_nnconv_forward (edge phase)")
            edge_in = True
            if edge_only:
                pl.edge_in = ws.data_ptr()
                _prepared[key] = (state, pl.keep)
                return None
        if x_ready is not None:
            stream.wait_event(x_ready)
        if edge_in:
            flags |= _lib.FWD_EDGE_DONE
        check(lib.fesr_nnconv_forward(*args, _ptr(x), *graph, flags, _ptr(y), _ptr(ws), ws.numel(), sp), "fesr_nnconv_forward")
    _prepared[key] = (state, pl.keep)
    return y


# ---------------------------------------------------------------------------------- node weight
def node_weight(pred, target, csr: Csr, edge_attr, node_ptr=None, clamp_max=float('inf'), out=None):
    """Per-subdomain scalar of GradientbasedLoss.compute_node_weight -> [S] fp32 (written into `out` when given)."""
    dev = _require_cuda(pred, target, edge_attr)
    pred, target = _f32c(pred), _f32c(target)
    edge_attr = _f32c(edge_attr.reshape(-1))
    n_sub = 1 if node_ptr is None else int(node_ptr.numel() - 1)
    if out is None:
        out = torch.empty(n_sub, dtype=torch.float32, device=dev)
    elif not out.is_cuda or out.dtype != torch.float32 or not out.is_contiguous() or out.numel() != n_sub:
        raise FesrError(f"out must be a contiguous fp32 CUDA tensor with {n_sub} entries")
    scratch = torch.empty(max(csr.n, 1), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.load().fesr_node_weight(_ptr(pred), _ptr(target), pred.shape[1], _ptr(csr.rowptr), _ptr(csr.src),
                                           _ptr(csr.perm), _ptr(edge_attr), _ptr(node_ptr), n_sub, csr.n, csr.E,
                                           clamp_max, _ptr(out), _ptr(scratch), _stream(dev)), "fesr_node_weight")
    return out


# ---------------------------------------------------------------------------------- stitch
@dataclass
class Occurrence:
    occ_ptr: torch.Tensor   # [N+1] int32
    occ_idx: torch.Tensor   # [n_tot] int32
    N: int
    n_tot: int


def occurrence_build(global_ids: torch.Tensor, N: int) -> Occurrence:
    dev = _require_cuda(global_ids)
    if global_ids.dtype != torch.int64:
        raise FesrError("global_ids must be int64")
    global_ids = global_ids.contiguous()
    n_tot = int(global_ids.numel())
    lib = _lib.load()
    occ_ptr = torch.empty(N + 1, dtype=torch.int32, device=dev)
    occ_idx = torch.empty(n_tot, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        nbytes = lib.fesr_occurrence_workspace_bytes(n_tot, N)
        ws = workspace.get(dev, "sort", nbytes)
        check(lib.fesr_occurrence_build(_ptr(global_ids), n_tot, N, _ptr(occ_ptr), _ptr(occ_idx), _ptr(ws),
                                        ws.numel(), _stream(dev)), "fesr_occurrence_build")
    return Occurrence(occ_ptr, occ_idx, N, n_tot)


def stitch_mean(values: torch.Tensor, occ: Occurrence, global_ids: torch.Tensor | None = None,
                want_merged: bool = False, want_count: bool = True, node_range=None):
    """values [n_tot, c] -> (field [N, c], count [N] | None, merged [n_tot, c] | None).
    node_range = (lo, hi): only the global nodes lo..hi-1 are stitched -> field [hi-lo, c], count [hi-lo] (the
    occurrence table is indexed from its row `lo`; a rank of a sharded run stitches its own slice of the mesh)."""
    dev = _require_cuda(values, occ.occ_ptr)
    values = _f32c(values)
    c = int(values.shape[1])
    lo, hi = (0, occ.N) if node_range is None else (int(node_range[0]), int(node_range[1]))
    if not (0 <= lo <= hi <= occ.N):
        raise FesrError(f"node_range {node_range} outside [0, {occ.N}]")
    if want_merged and (global_ids is None or (lo, hi) != (0, occ.N)):
        raise FesrError("merged output needs global_ids and the full node range")
    m = hi - lo
    field = torch.empty(m, c, dtype=torch.float32, device=dev)
    count = torch.empty(m, dtype=torch.int32, device=dev) if want_count else None
    merged = torch.empty(occ.n_tot, c, dtype=torch.float32, device=dev) if want_merged else None
    if m == 0:
        return field, count, merged
    occ_ptr = occ.occ_ptr if lo == 0 else occ.occ_ptr[lo:]
    with torch.cuda.device(dev):
        check(_lib.load().fesr_stitch_mean(_ptr(values), c, _ptr(occ_ptr), _ptr(occ.occ_idx), _ptr(global_ids),
                                           occ.n_tot, m, _ptr(field), _ptr(count), _ptr(merged), _stream(dev)),
              "fesr_stitch_mean")
    return field, count, merged


# ---------------------------------------------------------------------------------- train helpers
def mse_loss(pred, target, want_grad=True):
    dev = _require_cuda(pred, target)
    pred, target = _f32c(pred), _f32c(target)
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    grad = torch.empty_like(pred) if want_grad else None
    ws = workspace.get(dev, "reduce", _lib.REDUCE_WS_BYTES)
    with torch.cuda.device(dev):
        check(_lib.load().fesr_mse_loss(_ptr(pred), _ptr(target), pred.numel(), _ptr(loss), _ptr(grad), _ptr(ws),
                                        _stream(dev)), "fesr_mse_loss")
    return loss, grad


def adam_step(param, grad, exp_avg, exp_avg_sq, lr, step, beta1=0.9, beta2=0.999, eps=1e-8):
    dev = _require_cuda(param, grad, exp_avg, exp_avg_sq)
    with torch.cuda.device(dev):
        check(_lib.load().fesr_adam_step(_ptr(param), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq), param.numel(),
                                         lr, beta1, beta2, eps, step, _stream(dev)), "fesr_adam_step")
    invalidate_prepared_weights()        # written through a raw pointer: no tensor `_version` changed


# ---------------------------------------------------------------------------------- assembly
@dataclass
class Partition:
    home_leaf: torch.Tensor   # [C] int32
    tree_axis: torch.Tensor   # [2^k - 1] int32
    tree_split: torch.Tensor  # [2^k - 1] fp32
    leaf_ptr: torch.Tensor    # [S+1] int32
    leaf_cells: torch.Tensor  # [pairs] int32
    levels: int
    mode: int


@dataclass
class SubdomainBatch:
    """All subdomains of one mesh as one block-diagonal graph (device tensors)."""
    node_ptr: torch.Tensor    # [S+1] int32
    edge_ptr: torch.Tensor    # [S+1] int32
    global_ids: torch.Tensor  # [n_tot] int64, ascending inside each subdomain
    edge_src: torch.Tensor    # [e_tot] int32 batch-level
    edge_dst: torch.Tensor    # [e_tot] int32 batch-level, (subdomain, dst, src) order
    edge_attr: torch.Tensor   # [e_tot] fp32
    rowptr: torch.Tensor      # [n_tot + 1] int32
    n_sub: int
    n_tot: int
    e_tot: int

    @property
    def csr(self) -> Csr:
        return Csr(self.rowptr, self.edge_src, None, self.n_tot, self.e_tot)


def partition_cells(pos: torch.Tensor, cells: torch.Tensor, levels: int, mode: int = _lib.ALL_INTERSECTING) -> Partition:
    """kd decomposition of the cells into 2**levels leaves (+ halo assignment)."""
    dev = _require_cuda(pos, cells)
    if pos.dtype != torch.float32 or cells.dtype != torch.int32:
        raise FesrError("pos must be fp32 [N,3] and cells int32 [C,4]")
    pos, cells = pos.contiguous(), cells.contiguous()
    N, Cn = int(pos.shape[0]), int(cells.shape[0])
    S = 1 << levels
    lib = _lib.load()
    home = torch.empty(max(Cn, 1), dtype=torch.int32, device=dev)[:Cn]
    tree_axis = torch.zeros(max(S - 1, 1), dtype=torch.int32, device=dev)[:S - 1]
    tree_split = torch.zeros(max(S - 1, 1), dtype=torch.float32, device=dev)[:S - 1]
    leaf_ptr = torch.empty(S + 1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        st = _stream(dev)
        ws = workspace.get(dev, "asm", lib.fesr_partition_workspace_bytes(Cn, levels))
        check(lib.fesr_partition_cells(_ptr(pos), _ptr(cells), N, Cn, levels, _ptr(home), _ptr(tree_axis),
                                       _ptr(tree_split), _ptr(ws), ws.numel(), st), "fesr_partition_cells")
        total = C.c_int64(0)
        ws = workspace.get(dev, "asm", lib.fesr_assign_workspace_bytes(Cn, levels, 0))
        check(lib.fesr_assign_count(_ptr(pos), _ptr(cells), Cn, levels, mode, _ptr(home), _ptr(tree_axis),
                                    _ptr(tree_split), _ptr(leaf_ptr), C.byref(total), _ptr(ws), ws.numel(), st),
              "fesr_assign_count")
        P = int(total.value)
        leaf_cells = torch.empty(max(P, 1), dtype=torch.int32, device=dev)[:P]
        ws = workspace.get(dev, "asm", lib.fesr_assign_workspace_bytes(Cn, levels, P))
        check(lib.fesr_assign_fill(_ptr(pos), _ptr(cells), Cn, levels, mode, _ptr(home), _ptr(tree_axis),
                                   _ptr(tree_split), _ptr(leaf_ptr), P, _ptr(leaf_cells), _ptr(ws), ws.numel(), st),
              "fesr_assign_fill")
    return Partition(home, tree_axis, tree_split, leaf_ptr, leaf_cells, levels, mode)


def build_subdomains(pos: torch.Tensor, cells: torch.Tensor, leaf_ptr: torch.Tensor, leaf_cells: torch.Tensor) -> SubdomainBatch:
    """Node compaction + edge build + CSR for every subdomain (one block-diagonal batch)."""
    dev = _require_cuda(pos, cells, leaf_ptr)
    pos, cells = pos.contiguous(), cells.contiguous()
    N = int(pos.shape[0])
    S = int(leaf_ptr.numel() - 1)
    P = int(leaf_cells.numel())
    lib = _lib.load()
    node_ptr = torch.empty(S + 1, dtype=torch.int32, device=dev)
    edge_ptr = torch.empty(S + 1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        st = _stream(dev)
        ws = workspace.get(dev, "asm", lib.fesr_subdomain_workspace_bytes(P, N, S))
        totals = (C.c_int64 * 2)(0, 0)
        check(lib.fesr_subdomain_count(_ptr(cells), _ptr(leaf_ptr), _ptr(leaf_cells), S, P, N, _ptr(node_ptr),
                                       _ptr(edge_ptr), totals, _ptr(ws), ws.numel(), st), "fesr_subdomain_count")
        n_tot, e_tot = int(totals[0]), int(totals[1])
        gids = torch.empty(max(n_tot, 1), dtype=torch.int64, device=dev)[:n_tot]
        esrc = torch.empty(max(e_tot, 1), dtype=torch.int32, device=dev)[:e_tot]
        edst = torch.empty(max(e_tot, 1), dtype=torch.int32, device=dev)[:e_tot]
        eattr = torch.empty(max(e_tot, 1), dtype=torch.float32, device=dev)[:e_tot]
        rowptr = torch.empty(n_tot + 1, dtype=torch.int32, device=dev)
        check(lib.fesr_subdomain_fill(_ptr(pos), _ptr(node_ptr), _ptr(edge_ptr), S, P, n_tot, e_tot, _ptr(gids),
                                      _ptr(esrc), _ptr(edst), _ptr(eattr), _ptr(rowptr), _ptr(ws), ws.numel(), st),
              "fesr_subdomain_fill")
    return SubdomainBatch(node_ptr, edge_ptr, gids, esrc, edst, eattr, rowptr, S, n_tot, e_tot)


def assemble(pos, cells, levels: int, mode: int = _lib.ALL_INTERSECTING):
    """partition_cells + build_subdomains: the GPU replacement of get_partition_domain."""
    part = partition_cells(pos, cells, levels, mode)
    return part, build_subdomains(pos, cells, part.leaf_ptr, part.leaf_cells)


# ---------------------------------------------------------------------------------- ALDS routing
_F64_CACHE = {}


def _f64(a, dev):
    """fp64 device copy of a (small, read-only) model array.  numpy arrays -- sklearn's fitted attributes -- are uploaded
    once per (array object, device) and verified by content on every reuse: the routed predict called this five times
    per step, each a pageable host -> device copy with the GPU idle behind it."""
    if torch.is_tensor(a):
        return a.to(dev, dtype=torch.float64).contiguous()
    import numpy as np
    arr = np.ascontiguousarray(a, dtype=np.float64)
    key = (id(a), str(dev))
    ent = _F64_CACHE.get(key)
    if ent is not None and ent[0].shape == arr.shape and np.array_equal(ent[0], arr):
        return ent[1]
    if len(_F64_CACHE) > 64:
        _F64_CACHE.clear()
    t = torch.from_numpy(arr.copy()).to(dev)
    _F64_CACHE[key] = (arr.copy(), t)
    return t


def route(x, node_ptr, pca_mean, pca_components, scaler_mean=None, scaler_scale=None, centroids=None, rows=280):
    """PCA latent of the first `rows` nodes of every subdomain (+ k-means label when a classifier
    is given).  Returns (labels int32 [S] | None, latent fp64 [S, n_comp])."""
    dev = _require_cuda(x, node_ptr)
    x = _f32c(x)
    S = int(node_ptr.numel() - 1)
    comps = _f64(pca_components, dev)
    n_comp = int(comps.shape[0])
    mean = _f64(pca_mean, dev)
    if mean.numel() != rows * x.shape[1] or comps.shape[1] != rows * x.shape[1]:
        raise FesrError(f"PCA was fitted on {mean.numel()} features, expected rows*channels = {rows * x.shape[1]}")
    latent = torch.empty(S, n_comp, dtype=torch.float64, device=dev)
    labels = None
    sm = ss = cc = None
    k = 0
    if centroids is not None:
        sm, ss, cc = _f64(scaler_mean, dev), _f64(scaler_scale, dev), _f64(centroids, dev)
        k = int(cc.shape[0])
        labels = torch.empty(S, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.load().fesr_route(_ptr(x), int(x.shape[1]), _ptr(node_ptr), S, rows, _ptr(mean), _ptr(comps), n_comp,
                                     _ptr(sm), _ptr(ss), _ptr(cc), k, _ptr(labels), _ptr(latent), _stream(dev)),
              "fesr_route")
    return labels, latent


def cluster(latent, scaler_mean, scaler_scale, centroids):
    dev = _require_cuda(latent)
    latent = latent.to(torch.float64).contiguous()
    S, n_comp = int(latent.shape[0]), int(latent.shape[1])
    sm, ss, cc = _f64(scaler_mean, dev), _f64(scaler_scale, dev), _f64(centroids, dev)
    labels = torch.empty(S, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.load().fesr_cluster(_ptr(latent), S, n_comp, _ptr(sm), _ptr(ss), _ptr(cc), int(cc.shape[0]),
                                       _ptr(labels), _stream(dev)), "fesr_cluster")
    return labels


def gather_rows(src, index, out=None):
    """out[r] = src[index[r]] for fp32 rows of a multiple of 4 floats (fesr_gather_rows; models/scheduler_gnn.py:240-251)."""
    dev = _require_cuda(src)
    src = src.contiguous()
    index = index.to(torch.int64).contiguous()
    rf = 1
    for dim in src.shape[1:]:
        rf *= int(dim)
    if out is None:
        out = torch.empty((int(index.numel()),) + tuple(src.shape[1:]), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.load().fesr_gather_rows(_ptr(src), _ptr(index), int(index.numel()), rf, _ptr(out), _stream(dev)),
              "fesr_gather_rows")
    return out


def scatter_rows(src, index, out):
    """out[index[r]] = src[r] (distinct indices; fesr_scatter_rows; reorder_predictions, models/scheduler_gnn.py:302-309)."""
    dev = _require_cuda(src)
    src = src.contiguous()
    index = index.to(torch.int64).contiguous()
    rf = 1
    for dim in src.shape[1:]:
        rf *= int(dim)
    if not out.is_contiguous():
        raise FesrError("scatter_rows needs a contiguous destination")
    with torch.cuda.device(dev):
        check(_lib.load().fesr_scatter_rows(_ptr(src), _ptr(index), int(index.numel()), rf, _ptr(out), _stream(dev)),
              "fesr_scatter_rows")
    return out


# ---------------------------------------------------------------------------------- backward
def reverse_csr(csr: Csr) -> Csr:
    """CSR of the reversed graph over the forward CSR slots: rowptr groups edges by SOURCE,
    .src is the original destination, .perm maps reversed slot -> forward slot."""
    cached = getattr(csr, "_rev", None)
    if cached is not None:
        return cached
    dev = csr.rowptr.device
    deg = (csr.rowptr[1:] - csr.rowptr[:-1]).long()
    dst = torch.repeat_interleave(torch.arange(csr.n, device=dev, dtype=torch.int64), deg)
    rev = csr_build(torch.stack([dst, csr.src.long()]).contiguous(), csr.n)
    csr._rev = rev
    return rev


def nnconv_backward(dims: ModelDims, tensors: dict, x, csr: Csr, edge_attr, precision, grad_y, fwd_ws,
                    need_grad_x: bool = False, grads: dict | None = None):
    """Runs fesr_nnconv_backward. Returns (grads dict shaped like `tensors`, grad_x | None).
    grads: accumulators shaped like `tensors` that the call ADDS into (e.g. views of a flat gradient buffer);
    default: fresh zero tensors."""
    dev = _require_cuda(x, grad_y)
    x = _f32c(x)
    grad_y = _f32c(grad_y)
    edge_attr = _f32c(edge_attr.reshape(-1))
    rev = reverse_csr(csr)
    lib = _lib.load()
    p, keep = make_params(Params, tensors)
    gt = grads if grads is not None else \
        {k: (None if v is None else torch.zeros_like(v) if torch.is_tensor(v) else [torch.zeros_like(t) for t in v])
         for k, v in tensors.items()}
    g, keep_g = make_params(ParamGrads, gt)
    grad_x = torch.empty_like(x) if need_grad_x else None
    n, E = csr.n, csr.E
    with torch.cuda.device(dev):
        nbytes = lib.fesr_backward_workspace_bytes(C.byref(dims), n, E)
        ws = workspace.get(dev, "bwd", nbytes)
        check(lib.fesr_nnconv_backward(C.byref(dims), C.byref(p), _ptr(x), _ptr(csr.rowptr), _ptr(csr.src),
                                       _ptr(csr.perm), _ptr(rev.rowptr), _ptr(rev.src), _ptr(rev.perm),
                                       _ptr(edge_attr), n, E, precision, _ptr(grad_y), _ptr(fwd_ws), C.byref(g),
                                       _ptr(grad_x), _ptr(ws), ws.numel(), _stream(dev)), "fesr_nnconv_backward")
    del keep, keep_g
    return gt, grad_x


# ---------------------------------------------------------------------------------- low-res -> high-res transfer
def interp_gaussian(src_pos: torch.Tensor, src_val: torch.Tensor, dst_pos: torch.Tensor, radius: float,
                    sharpness: float = 2.0, null_value: float = 0.0, want_count: bool = False):
    """Gaussian-kernel point interpolation (fesr_interp_gaussian): src_val [n_src, c] at src_pos [n_src, 3] ->
    [n_dst, c] at dst_pos.  c in {1, 3, 4}; a 1-D src_val is treated as one channel and comes back 1-D."""
    dev = _require_cuda(src_pos, src_val, dst_pos)
    src_pos, dst_pos = _f32c(src_pos), _f32c(dst_pos)
    flat = src_val.dim() == 1
    src_val = _f32c(src_val.unsqueeze(1) if flat else src_val.flatten(1))
    n_src, c, n_dst = int(src_pos.shape[0]), int(src_val.shape[1]), int(dst_pos.shape[0])
    if src_pos.shape[1:] != (3,) or dst_pos.shape[1:] != (3,) or src_val.shape[0] != n_src:
        raise FesrError("interp_gaussian: positions must be [n, 3] and src_val [n_src, c]")
    lib = _lib.load()
    out = torch.empty(n_dst, c, dtype=torch.float32, device=dev)
    count = torch.empty(n_dst, dtype=torch.int32, device=dev) if want_count else None
    with torch.cuda.device(dev):
        ws = workspace.get(dev, "sort", lib.fesr_interp_workspace_bytes(n_src, c))
        check(lib.fesr_interp_gaussian(_ptr(src_pos), _ptr(src_val), c, n_src, _ptr(dst_pos), n_dst, float(radius),
                                       float(sharpness), float(null_value), _ptr(out), _ptr(count), _ptr(ws),
                                       ws.numel(), _stream(dev)), "fesr_interp_gaussian")
    out = out.reshape(-1) if flat else out
    return (out, count) if want_count else out


# ---------------------------------------------------------------------------------- wall shear stress (post-processing)
def wall_shear_stress(pos: torch.Tensor, cells: torch.Tensor, velocity: torch.Tensor, dynamic_viscosity: float = 1.0):
    """compute_wss.compute_wall_shear_stress on a tetrahedral mesh: point gradients of `velocity` [N, 3] (mean of the
    incident tets' gradients), boundary faces (outward), point normals, tau_wall = tau - (tau . n) n.
    -> dict(surface_nodes [M] int64, faces [F, 3] int32, normals [M, 3], gradient [N, 9], wss [M, 3], wss_magnitude [M])."""
    dev = _require_cuda(pos, cells, velocity)
    if pos.dtype != torch.float32 or cells.dtype != torch.int32 or cells.dim() != 2 or cells.shape[1] != 4:
        raise FesrError("pos must be fp32 [N, 3] and cells int32 [C, 4]")
    pos, cells, velocity = pos.contiguous(), cells.contiguous(), _f32c(velocity)
    N, Cn = int(pos.shape[0]), int(cells.shape[0])
    if velocity.shape != (N, 3):
        raise FesrError(f"velocity must be [{N}, 3], got {tuple(velocity.shape)}")
    lib = _lib.load()
    st = _stream(dev)
    with torch.cuda.device(dev):
        grad_c = torch.empty(Cn, 9, dtype=torch.float32, device=dev)
        check(lib.fesr_tet_gradient(_ptr(pos), _ptr(cells), _ptr(velocity), Cn, _ptr(grad_c), st), "fesr_tet_gradient")
        occ = occurrence_build(cells.reshape(-1).long(), N)                # node -> (cell, corner) incidences
        grad_p = torch.empty(N, 9, dtype=torch.float32, device=dev)
        check(lib.fesr_incident_mean(_ptr(grad_c), 9, _ptr(occ.occ_ptr), _ptr(occ.occ_idx), 4, N, 0, _ptr(grad_p), st),
              "fesr_incident_mean")
        faces = torch.empty(max(4 * Cn, 1), 3, dtype=torch.int32, device=dev)
        face_cell = torch.empty(max(4 * Cn, 1), dtype=torch.int32, device=dev)
        face_n = torch.empty(max(4 * Cn, 1), 3, dtype=torch.float32, device=dev)
        ws = workspace.get(dev, "sort", lib.fesr_boundary_faces_workspace_bytes(Cn))
        cnt = C.c_int64(0)
        check(lib.fesr_boundary_faces(_ptr(pos), _ptr(cells), N, Cn, _ptr(faces), _ptr(face_cell), _ptr(face_n),
                                      C.byref(cnt), _ptr(ws), ws.numel(), st), "fesr_boundary_faces")
        F = int(cnt.value)
        faces, face_n = faces[:F].contiguous(), face_n[:F].contiguous()
        focc = occurrence_build(faces.reshape(-1).long(), N)               # node -> boundary-face incidences
        normal_p = torch.empty(N, 3, dtype=torch.float32, device=dev)
        check(lib.fesr_incident_mean(_ptr(face_n), 3, _ptr(focc.occ_ptr), _ptr(focc.occ_idx), 3, N, 1, _ptr(normal_p), st),
              "fesr_incident_mean")
        surf = torch.nonzero(focc.occ_ptr[1:] > focc.occ_ptr[:-1]).reshape(-1)
        M = int(surf.numel())
        surf32 = surf.to(torch.int32)
        tau = torch.empty(M, 3, dtype=torch.float32, device=dev)
        mag = torch.empty(M, dtype=torch.float32, device=dev)
        check(lib.fesr_wall_shear_stress(_ptr(grad_p), _ptr(normal_p), _ptr(surf32), M, float(dynamic_viscosity),
                                         _ptr(tau), _ptr(mag), st), "fesr_wall_shear_stress")
    return {"surface_nodes": surf, "faces": faces, "face_cell": face_cell[:F], "normals": normal_p[surf], "gradient": grad_p,
            "wss": tau, "wss_magnitude": mag}
