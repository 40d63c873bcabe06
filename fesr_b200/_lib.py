"""ctypes binding of libfesr.so (the C ABI declared in include/fesr.h).

There is no fallback: if the shared library is missing or the device is not sm_100 every
entry point raises.  ``python -m fesr_b200.build`` (or ``__graft_entry__.build()``) builds it.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FESR_LIB_PATH") or os.path.join(HERE, "lib", "libfesr.so")      # (env: tools/dev A/B builds)

KERNELNN, TEECNET = 0, 1
PREC_FP32, PREC_TF32, PREC_TF32X3, PREC_F16 = 0, 1, 2, 3
ONE_REGION, ALL_INTERSECTING = 0, 1
FWD_KEEP, FWD_WEIGHTS_PREPARED, FWD_EDGE_ONLY, FWD_EDGE_DONE, FWD_KEEP_Z16 = 1, 2, 4, 8, 16
REDUCE_WS_BYTES = 8192
PRECISIONS = {"fp32": PREC_FP32, "tf32": PREC_TF32, "f16": PREC_F16, "fp16": PREC_F16}


class FesrError(RuntimeError):
    pass


class ModelDims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("kind", "w", "wp", "in_ch", "out_ch", "layers", "n_hidden")] + \
               [("hidden", C.c_int32 * 4)] + \
               [(n, C.c_int32) for n in ("k1", "kt", "ktp", "passes", "kp", "k1p", "zk_main", "zk", "leaky")]


_FP = C.c_void_p


class Params(C.Structure):
    _fields_ = [("fc1_w", _FP), ("fc1_b", _FP), ("mlp_w", _FP * 4), ("mlp_b", _FP * 4),
                ("lin_w", _FP), ("lin_b", _FP), ("root", _FP), ("bias", _FP), ("fc2_w", _FP), ("fc2_b", _FP)]


class ParamGrads(Params):
    pass


_i64, _i32, _sz, _vp, _f = C.c_int64, C.c_int32, C.c_size_t, C.c_void_p, C.c_float

# name -> (restype, argtypes); the CPU test checks this table against include/fesr.h
SIGNATURES = {
    "fesr_version": (C.c_int, []),
    "fesr_last_error": (C.c_char_p, []),
    "fesr_device_check": (C.c_int, []),
    "fesr_launch_count": (C.c_longlong, []),
    "fesr_profile_enable": (C.c_int, [C.c_int]),
    "fesr_profile_collect": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.c_int]),
    "fesr_model_dims_init": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(ModelDims)]),
    "fesr_csr_workspace_bytes": (_sz, [_i64, _i64]),
    "fesr_csr_build": (C.c_int, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "fesr_forward_workspace_bytes": (_sz, [C.POINTER(ModelDims), _i64, _i64, C.c_int]),
    "fesr_forward_overflow_offset": (_sz, [C.POINTER(ModelDims)]),
    "fesr_nnconv_forward": (C.c_int, [C.POINTER(ModelDims), C.POINTER(Params), _vp, _vp, _vp, _vp, _vp, _i64, _i64,
                                      C.c_int, C.c_int, _vp, _vp, _sz, _vp]),
    "fesr_backward_workspace_bytes": (_sz, [C.POINTER(ModelDims), _i64, _i64]),
    "fesr_nnconv_backward": (C.c_int, [C.POINTER(ModelDims), C.POINTER(Params), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                       _i64, _i64, C.c_int, _vp, _vp, C.POINTER(ParamGrads), _vp, _vp, _sz, _vp]),
    "fesr_mse_loss": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "fesr_adam_step": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _f, _f, _f, _f, _i64, _vp]),
    "fesr_node_weight": (C.c_int, [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _i64, _f, _vp, _vp, _vp]),
    "fesr_occurrence_workspace_bytes": (_sz, [_i64, _i64]),
    "fesr_occurrence_build": (C.c_int, [_vp, _i64, _i64, _vp, _vp, _vp, _sz, _vp]),
    "fesr_stitch_mean": (C.c_int, [_vp, _i32, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp]),
    "fesr_partition_workspace_bytes": (_sz, [_i64, _i32]),
    "fesr_partition_cells": (C.c_int, [_vp, _vp, _i64, _i64, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "fesr_assign_count": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, C.POINTER(_i64), _vp, _sz, _vp]),
    "fesr_assign_workspace_bytes": (_sz, [_i64, _i32, _i64]),
    "fesr_assign_fill": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _sz, _vp]),
    "fesr_subdomain_workspace_bytes": (_sz, [_i64, _i64, _i32]),
    "fesr_subdomain_count": (C.c_int, [_vp, _vp, _vp, _i32, _i64, _i64, _vp, _vp, C.POINTER(_i64), _vp, _sz, _vp]),
    "fesr_subdomain_fill": (C.c_int, [_vp, _vp, _vp, _i32, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "fesr_cluster": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _vp, _i32, _vp, _vp]),
    "fesr_gather_rows": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp]),
    "fesr_scatter_rows": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp]),
    "fesr_interp_workspace_bytes": (_sz, [_i64, _i32]),
    "fesr_interp_gaussian": (C.c_int, [_vp, _vp, _i32, _i64, _vp, _i64, _f, _f, _f, _vp, _vp, _vp, _sz, _vp]),
    "fesr_tet_gradient": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "fesr_incident_mean": (C.c_int, [_vp, _i32, _vp, _vp, _i32, _i64, _i32, _vp, _vp]),
    "fesr_boundary_faces_workspace_bytes": (_sz, [_i64]),
    "fesr_boundary_faces": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, C.POINTER(_i64), _vp, _sz, _vp]),
    "fesr_wall_shear_stress": (C.c_int, [_vp, _vp, _vp, _i64, _f, _vp, _vp, _vp]),
    "fesr_comm_unique_id": (C.c_int, [_vp]),
    "fesr_comm_init": (C.c_int, [_vp, C.c_int, C.c_int]),
    "fesr_comm_destroy": (C.c_int, []),
    "fesr_comm_rank": (C.c_int, []),
    "fesr_comm_world": (C.c_int, []),
    "fesr_allgatherv_pred": (C.c_int, [_vp, _i64, _vp]),
    "fesr_allreduce_grads": (C.c_int, [_vp, _i64, _vp]),
    "fesr_route": (C.c_int, [_vp, _i32, _vp, _i32, _i32, _vp, _vp, _i32, _vp, _vp, _vp, _i32, _vp, _vp, _vp]),
}

_lib = None


def load() -> C.CDLL:
    """Loads libfesr.so and types every entry point.  Raises FesrError if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FesrError(
            f"{LIB_PATH} not found: the CUDA library is not built. Run `python -m fesr_b200.build` "
            "(nvcc, sm_100a). fesr_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().fesr_last_error().decode(errors="replace")
        raise FesrError(f"{what or 'libfesr call'} failed ({rc}): {msg}")


PROF_KINDS = ("prepare", "edge_hidden", "fc_in", "zbuild", "node_gemm", "fc_out", "node_weight", "stitch", "graph",
              "backward", "layer_fused")


def profile_enable(on: bool):
    check(load().fesr_profile_enable(int(on)), "fesr_profile_enable")


def profile_collect():
    n = len(PROF_KINDS)
    ms = (C.c_double * n)()
    cnt = (C.c_longlong * n)()
    check(load().fesr_profile_collect(ms, cnt, n), "fesr_profile_collect")
    return {k: (ms[i], cnt[i]) for i, k in enumerate(PROF_KINDS)}


def launch_count() -> int:
    return int(load().fesr_launch_count())


def model_dims(kind: int, w: int, in_ch: int, out_ch: int, layers: int) -> ModelDims:
    d = ModelDims()
    check(load().fesr_model_dims_init(kind, w, in_ch, out_ch, layers, C.byref(d)), "fesr_model_dims_init")
    return d
