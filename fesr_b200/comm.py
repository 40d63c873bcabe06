"""libfesr's own NCCL communicator (include/fesr.h: fesr_comm_*): the collectives of the sharded path are issued
by the library on the caller's CUDA stream, in place on caller-owned buffers -- no torch.distributed call, no
pad / cat / clone around them.  `torch.distributed` stays the bootstrap side channel that ships the NCCL
unique id (and what the CPU `gloo` tests of the host logic run on).

Replaces the reference's mp.Process + Manager().dict() fan-out / fan-in (models/scheduler_gnn.py:254-291)
and DistributedDataParallel's gradient all-reduce (:386).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import FesrError, check

_state = {"rank": 0, "world": 1, "ready": False}


def ready() -> bool:
    return _state["ready"]


def rank() -> int:
    return _state["rank"]


def world() -> int:
    return _state["world"]


def init_from_torch_distributed(group=None) -> bool:
    """Creates the communicator for the ranks of the initialised torch.distributed world (one process per GPU,
    torch.cuda.current_device() = this rank's GPU).  Idempotent.  Returns False on a single-process run."""
    import torch.distributed as dist
    if _state["ready"]:
        return True
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return False
    if not torch.cuda.is_available():
        raise FesrError("fesr_b200.comm needs a B200 per rank; there is no CPU collective path")
    lib = _lib.load()
    r, w = dist.get_rank(group), dist.get_world_size(group)
    uid = (C.c_char * 128)()
    if r == 0:
        check(lib.fesr_comm_unique_id(uid), "fesr_comm_unique_id")
    box = [bytes(uid.raw) if r == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    uid.raw = box[0]
    with torch.cuda.device(torch.cuda.current_device()):
        check(lib.fesr_comm_init(uid, r, w), "fesr_comm_init")
    _state.update(rank=r, world=w, ready=True)
    return True


def destroy():
    if _state["ready"]:
        check(_lib.load().fesr_comm_destroy(), "fesr_comm_destroy")
        _state.update(rank=0, world=1, ready=False)


def allgatherv_pred(slots: torch.Tensor, stream=None):
    """slots [world, slot_elems] fp32 (contiguous, CUDA): this rank has filled slot `rank`; afterwards every slot
    is filled.  In place, asynchronous on `stream` (default: the current stream)."""
    if not _state["ready"]:
        raise FesrError("fesr_b200.comm is not initialised (comm.init_from_torch_distributed())")
    if not slots.is_cuda or slots.dtype != torch.float32 or not slots.is_contiguous() or slots.dim() != 2 \
            or slots.shape[0] != _state["world"]:
        raise FesrError(f"slots must be a contiguous fp32 CUDA tensor [world={_state['world']}, slot_elems]")
    st = stream if stream is not None else torch.cuda.current_stream(slots.device)
    with torch.cuda.device(slots.device):
        check(_lib.load().fesr_allgatherv_pred(C.c_void_p(slots.data_ptr()), int(slots.shape[1]),
                                               C.c_void_p(st.cuda_stream)), "fesr_allgatherv_pred")


def allreduce_grads(flat: torch.Tensor, stream=None):
    """flat fp32 CUDA buffer <- mean over the ranks, in place, on `stream` (default: the current stream)."""
    if not _state["ready"]:
        raise FesrError("fesr_b200.comm is not initialised (comm.init_from_torch_distributed())")
    if not flat.is_cuda or flat.dtype != torch.float32 or not flat.is_contiguous():
        raise FesrError("flat must be a contiguous fp32 CUDA tensor")
    st = stream if stream is not None else torch.cuda.current_stream(flat.device)
    with torch.cuda.device(flat.device):
        check(_lib.load().fesr_allreduce_grads(C.c_void_p(flat.data_ptr()), int(flat.numel()),
                                               C.c_void_p(st.cuda_stream)), "fesr_allreduce_grads")
