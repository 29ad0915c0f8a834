"""ctypes binding of the C-ABI CUDA library (include/p2t_b200.h).

There is no CPU fallback: if the shared library is missing the import of any compute entry fails
loudly (build it with `python prot2text-v2-esm3_b200/build.py` or `__graft_entry__.build()`).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("P2T_LIB_PATH") or os.path.join(_HERE, "libp2t_b200.so")  # the override serves build-variant A/B runs

_vp, _i, _ll, _f, _ull = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_ulonglong


class OverlapReduce(C.Structure):
    """p2t_overlap_reduce_t of include/p2t_b200.h"""
    _fields_ = [("peers", C.POINTER(C.c_void_p)), ("world", C.c_int), ("rank", C.c_int), ("n_bytes", C.c_longlong),
                ("f32_from_byte", C.c_longlong)]


# name -> argument ctypes, in header order (include/p2t_b200.h)
SIGNATURES = {
    "p2t_gemm_bf16": [_vp, _ll, _i, _vp, _ll, _i, _vp, _ll, _i, _i, _i, _i, _f, _vp, _vp, _vp, _i, _vp],
    "p2t_rows_plan": [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "p2t_rows_plan_counts": [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "p2t_stage_rows_h2d": [_vp, _ll, _ll, _vp, _vp, _i, _vp, C.POINTER(_vp), _i],
    "p2t_stage_rows_pull": [_vp, _vp, _vp, _i, _vp, _i, _vp],
    "p2t_row_inv_norm": [_vp, _i, _vp, _i, _vp, _vp],
    "p2t_gather_rows": [_vp, _ll, _vp, _vp, _i, _i, _vp, _vp],
    "p2t_adapter_fwd": [_vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _f, _ull, _vp, _vp, _i, _vp],
    "p2t_adapter_scale_rows": [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp],
    "p2t_pool_fwd": [_vp, _i, _ll, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _ll, _vp, _vp, _vp, _vp],
    "p2t_loss_bwd_coef": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp],
    "p2t_readout_last": [_vp, _vp, _i, _i, _i, _vp, _vp],
    "p2t_readout_last_bwd": [_vp, _vp, _i, _i, _i, _vp, _vp],
    "p2t_l2norm_fwd": [_vp, _i, _i, _vp, _vp, _vp, _vp],
    "p2t_l2norm_bwd": [_vp, _vp, _vp, _i, _i, _vp, _vp],
    "p2t_pool_bwd_coef": [_vp, _ll, _vp, _ll, _vp, _i, _i, _i, _vp, _vp, _vp],
    "p2t_readout_bwd": [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "p2t_adapter_tail_bwd": [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _i, _vp, _vp, _vp],
    "p2t_adapter_tail_bwd_dy": [_vp, _vp, _vp, _vp, _i, _vp, _i, _i, _vp, _vp],
    "p2t_adapter_bwd": [_vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i,
                        C.POINTER(OverlapReduce), _i, _vp],
    "p2t_bias_grads": [_vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _i, _vp],
    "p2t_loss_fused": [_vp, _vp, C.POINTER(_vp), _i, _i, _ll, _vp, _i, _i, _i, _i, _f, _f, _f, _f, _i, _i, _i,
                       _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "p2t_similarity": [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp, _i, _vp],
    "p2t_infonce_col_stats": [_vp, _i, _i, _vp, _vp, _vp, _i, _vp],
    "p2t_infonce_ce": [_vp, _vp, _i, _i, _f, _f, _f, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _vp],
    "p2t_infonce_stats": [_vp, _vp, _vp, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp],
    "p2t_infonce_finish": [_vp, _vp, _vp, _i, _i, _i, _f, _f, _f, _f, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp],
    "p2t_infonce_grad": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp, _vp, _i, _vp],
    "p2t_loss_mean": [_vp, _i, _f, _vp, _i, _vp],
    "p2t_f32_to_bf16": [_vp, _ll, _vp, _vp],
    "p2t_bf16_to_f32": [_vp, _ll, _vp, _vp],
    "p2t_colsum": [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp],
    "p2t_dropout_mask": [_i, _i, _f, _ull, _i, _vp, _vp],
    "p2t_adapter_scatter_rows": [_vp, _vp, _i, _i, _i, _i, _vp, _ll, _vp, _vp, _vp, _vp, _vp],
    "p2t_peer_alloc": [_ull, C.POINTER(_vp), C.c_char_p],
    "p2t_peer_open": [C.c_char_p, C.POINTER(_vp)],
    "p2t_peer_close": [_vp],
    "p2t_peer_free": [_vp],
    "p2t_peer_allgather": [C.POINTER(_vp), _i, _i, _vp, _ll, _vp, _i, _vp],
    "p2t_peer_allreduce_mean": [C.POINTER(_vp), _i, _i, _ll, _ll, _vp, _i, _vp],
    "p2t_peer_reset": [_vp, _vp],
    "p2t_copy_d2d": [_vp, _vp, _ull, _vp],
    "p2t_peer_status": [_vp, C.POINTER(C.c_uint)],
    "p2t_adamw_step": [_i, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp),
                       C.POINTER(_ll), _vp, _vp, _vp, _vp, _f, _f, _f, _f, _f, _i, _vp],
    "p2t_launch_timing_mark": [_vp],
    "p2t_launch_timing_collect": [C.POINTER(C.c_double), _i, C.POINTER(C.c_int), C.c_char_p, _i],
    "p2t_gemm_timing_collect": [C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_double), _i],
}
NON_STATUS = {"p2t_abi_version": (_i, []), "p2t_last_error": (C.c_char_p, []),
              "p2t_launch_count": (_ull, []), "p2t_reset_launch_count": (None, []),
              "p2t_gemm_timing_enable": (None, [_i]), "p2t_launch_timing_enable": (None, [_i]), "p2t_gemm_workspace_bytes": (_ull, []),
              "p2t_peer_ctrl_bytes": (_ull, []), "p2t_loss_fused_eligible": (_i, [_i, _i, _i, _i]), "p2t_adamw_workspace_floats": (_i, [_i, C.POINTER(_ll)])}

_lib = None


class P2TError(RuntimeError):
    pass


def load():
    """dlopen the library once; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise P2TError(
            f"{LIB_PATH} not found: the CUDA extension is not built and there is no CPU fallback. "
            "Run `python prot2text-v2-esm3_b200/build.py`.")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _i
    for name, (res, argtypes) in NON_STATUS.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = res
    _lib = lib
    return lib


def call(name: str, *args) -> None:
    """Invoke a status-returning entry point; raise P2TError with the library's message on failure."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.p2t_last_error()
        raise P2TError(f"{name} failed (code {rc}): {msg.decode() if msg else ''}")


def launch_count() -> int:
    return int(load().p2t_launch_count())


def reset_launch_count() -> None:
    load().p2t_reset_launch_count()


def gemm_timing_enable(on: bool) -> None:
    load().p2t_gemm_timing_enable(int(on))


def gemm_timing_collect(cap: int = 4096):
    """(total GEMM-kernel milliseconds, number of GEMM launches, per-launch ms list) since enable; synchronise first."""
    ms, n = C.c_double(0.0), C.c_int(0)
    each = (C.c_double * cap)()
    call("p2t_gemm_timing_collect", C.byref(ms), C.byref(n), each, cap)
    return ms.value, n.value, list(each[:min(n.value, cap)])


def launch_timing_enable(on: bool) -> None:
    load().p2t_launch_timing_enable(int(on))


def launch_timing_mark(stream: int) -> None:
    call("p2t_launch_timing_mark", stream)


def launch_timing_collect(cap: int = 1 << 16):
    """[(kernel name, ms since the previous stamp)] in launch order since enable; "mark" entries open a step.
    Synchronise first."""
    n = C.c_int(0)
    ms = (C.c_double * cap)()
    names = C.create_string_buffer(cap * 40)
    call("p2t_launch_timing_collect", ms, cap, C.byref(n), names, len(names))
    labels = names.value.decode().split("\n")[:n.value]
    return list(zip(labels, list(ms[:len(labels)])))
