"""ctypes binding of libpfo_b200.so (the C ABI declared in include/pfo_b200.h).

There is no CPU fallback: if the library is missing or a launch fails this module raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_int, c_int32, c_int64, c_uint32, c_uint64, c_float, c_double, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpfo_b200.so")

P = c_void_p
_SIGS = {
    "pfo_abi_version": (c_int, []),
    "pfo_neighbor_sample": (c_int, [P, P, P, P, P, P, c_int64, c_int, c_int, c_uint64, c_uint32, P, c_int, P, c_int64,
                                    P, P, P, P, P]),
    "pfo_mark_nodes": (c_int, [P, c_int64, c_int, P, P]),
    "pfo_compact_workspace_ints": (c_int64, [c_int64]),
    "pfo_compact_nodes": (c_int, [P, c_int64, P, P, P, P, P]),
    "pfo_map_slots": (c_int, [P, c_int64, c_int, P, P, P]),
    "pfo_linear_f32": (c_int, [P, c_int64, P, P, c_int64, c_int, P, P, c_int64, P, c_int64, c_int64, P, c_int, c_int,
                               c_float, c_int, P, P, c_int64, c_int, P]),
    "pfo_linear_bf16": (c_int, [P, c_int64, P, P, c_int64, c_int, P, P, c_int64, P, c_int64, c_int64, P, c_int, c_int,
                                c_float, c_int, P, P, c_int64, c_int, P]),
    "pfo_linear_tf32": (c_int, [P, c_int64, P, P, c_int64, c_int, P, P, c_int64, P, c_int64, c_int64, P, c_int, c_int,
                                c_float, c_int, P, P, c_int64, c_int, c_int, P]),
    "pfo_wgrad_tf32_workspace_floats": (c_int64, [c_int64, c_int, c_int, c_int]),
    "pfo_wgrad_tf32": (c_int, [P, c_int64, P, c_int64, P, c_int64, P, c_int, c_int, P, c_int64, P, c_int, P, c_int, P]),
    "pfo_wgrad_workspace_floats": (c_int64, [c_int64, c_int, c_int, c_int]),
    "pfo_wgrad_f32": (c_int, [P, c_int64, P, c_int64, P, c_int64, P, c_int, c_int, P, c_int64, P, c_int, P, P]),
    "pfo_gather_state": (c_int, [P, P, c_int64, c_int, c_int, P, P, c_int64, P, P, P, P, P, c_int64, P, c_int64, P, P, P]),
    "pfo_cell_forward": (c_int, [P, P, c_int64, c_int, c_int, c_int, P, P, c_int64, P, P, P, P, P, P]),
    "pfo_cell_backward": (c_int, [P, P, c_int64, c_int, c_int, c_int, P, P, c_int64, P, P, P, P, P, P]),
    "pfo_pack_cell": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, P, P, P]),
    "pfo_unpack_cell_grads": (c_int, [P, P, c_int, c_int, c_int, c_int, P, P, P, P, P]),
    "pfo_persist_rank": (c_int, [P, P, c_int, c_int, P, P, P, P, P, P, P, P]),
    "pfo_store_messages": (c_int, [P, P, P, P, c_int, c_int, c_int, P, P, P, P, P, P, P, P, P, P, c_int64, P, P, P, P]),
    "pfo_store_messages_mean": (c_int, [P, P, P, P, c_int, c_int, c_int, P, P, P, P, P, P, P, P, P, P, P, P, c_int64,
                                        P, P, P, P]),
    "pfo_build_messages": (c_int, [P, P, P, P, c_int, c_int, c_int, P, P, P, P, P, P, P, P, c_int64, P, P]),
    "pfo_apply_messages": (c_int, [P, P, c_int64, c_int, c_int, P, P, P, c_int64, P, P, P, P, c_int64, P, P, P, P]),
    "pfo_build_routed_messages": (c_int, [P, P, P, P, P, P, c_int, c_int, c_int, P, P, P, P, P, P, P, c_int, c_int,
                                          c_int, P, c_int64, P]),
    "pfo_apply_routed_messages": (c_int, [P, c_int64, c_int64, c_int, c_int, P, P, P, P, P, c_int64, P, P, P, P]),
    "pfo_route_plan": (c_int, [P, c_int64, P, c_int, c_int, P, P, P, P, P]),
    "pfo_scatter_rows": (c_int, [P, c_int64, P, c_int64, P, c_int, P, c_int64, P]),
    "pfo_gather_words": (c_int, [P, c_int64, P, c_int64, c_int, P, c_int64, c_uint32, P]),
    "pfo_pack_queries": (c_int, [P, P, P, P, c_int64, P, P]),
    "pfo_unpack_queries": (c_int, [P, c_int64, P, P, P, P]),
    "pfo_unroute_neighbors": (c_int, [P, P, c_int64, c_int, P, P, P, P]),
    "pfo_route_reply_rows": (c_int, [P, P, P, c_int64, c_int, P, P]),
    "pfo_unroute_rows": (c_int, [P, P, P, P, P, c_int64, c_int, P, P, P, P]),
    "pfo_peer_alloc": (c_int, [c_int64, P, P]),
    "pfo_peer_open": (c_int, [P, P]),
    "pfo_peer_close": (c_int, [P]),
    "pfo_peer_free": (c_int, [P]),
    "pfo_peer_max_ranks": (c_int, []),
    "pfo_peer_max_barriers": (c_int, []),
    "pfo_peer_header_bytes": (c_int64, []),
    "pfo_peer_push": (c_int, [P, P, c_int64, c_int, c_int, c_int64, P]),
    "pfo_peer_barrier": (c_int, [P, c_int, c_int, c_int, P, c_double, P]),
    "pfo_time_embedding_fwd": (c_int, [P, P, c_int64, c_int64, c_int, P, P, P, c_float, c_float, c_float,
                                       c_float, P, P, P, P, P]),
    "pfo_time_embedding_bwd": (c_int, [P, c_int64, c_int, P, P, P, P, P, P, P, P, P, c_int64, P]),
    "pfo_time_encode": (c_int, [P, P, P, c_int64, c_int, c_int, P, P, P]),
    "pfo_adam_flat": (c_int, [P, P, P, P, c_int64, c_float, c_float, c_float, c_float, P, P]),
    "pfo_reduce_partials": (c_int, [P, c_int, c_int, P, c_int, P]),
    "pfo_scatter_add_rows": (c_int, [P, c_int64, P, c_int64, c_int, P, c_int64, P]),
    "pfo_gather_rows": (c_int, [P, c_int64, P, c_int64, c_int, P, c_int64, P]),
    "pfo_attn_nbr_fwd": (c_int, [P, P, c_int64, P, P, P, P, P, P, c_int64, c_int, c_int, c_int, c_int, c_int,
                                 c_float, c_uint64, c_uint32, P, P, c_int64, P, P, P]),
    "pfo_attn_nbr_fwd_rows": (c_int, [P, P, P, c_int64, P, P, P, P, P, P, c_int64, c_int, c_int, c_int, c_int, c_int,
                                      c_float, c_uint64, c_uint32, P, P, c_int64, P, P, P]),
    "pfo_attn_nbr_bwd_workspace_floats": (c_int64, [c_int]),
    "pfo_attn_nbr_bwd": (c_int, [P, P, c_int64, P, P, P, c_int64, P, P, P, P, P, P, c_int64, c_int, c_int, c_int, c_int, c_int,
                                 c_float, c_uint64, c_uint32, P, P, P, c_int64, P, c_int, P, P]),
    "pfo_fold_attention_workspace_doubles": (c_int64, [c_int, c_int, c_int]),
    "pfo_fold_attention_fwd": (c_int, [P] * 9 + [c_int, c_int, c_int, c_int, P, P, P, P, P]),
    "pfo_fold_attention_bwd": (c_int, [P] * 9 + [c_int, c_int, c_int, c_int, P] + [P] * 3 + [P] * 9 + [P]),
    "pfo_bpr": (c_int, [P, P, P, c_int, c_int, c_int, P, P, P, P, c_float, P, P]),
    "pfo_eval_score": (c_int, [P, P, P, c_int, c_int, c_int, c_int, P, P, P, P]),
    "pfo_eval_metrics": (c_int, [P, P, c_int, P, P, c_int, c_int, P, P, P, P, P, c_int, c_int, c_int, P, P, P]),
    "pfo_mv_select": (c_int, [P, P, P, P, P, P, c_int, P, c_int, c_int, c_int, c_int, c_double, c_double, c_int,
                              c_int, c_uint64, c_int, P, P, P, P, c_int, P]),
    "pfo_sample_candidates": (c_int, [P, P, P, P, c_int, c_int, c_int, c_uint64, P, P]),
}

EXPORTS = tuple(_SIGS)
ABI_VERSION = 3         # PFO_ABI_VERSION of include/pfo_b200.h this binding was written against
_lib = None
LAUNCHES = 0            # kernels launched through this binding (bench.py reports it)
_LAUNCHES_PER_CALL = {"pfo_compact_nodes": 1, "pfo_fold_attention_fwd": 2, "pfo_fold_attention_bwd": 2,
                      "pfo_fold_attention_workspace_doubles": 0, "pfo_wgrad_f32": 2, "pfo_wgrad_tf32": 2,
                      "pfo_wgrad_tf32_workspace_floats": 0, "pfo_time_embedding_bwd": 2,
                      "pfo_attn_nbr_bwd": 2, "pfo_bpr": 2, "pfo_eval_metrics": 2, "pfo_apply_messages": 2, "pfo_apply_routed_messages": 2, "pfo_abi_version": 0, "pfo_peer_alloc": 0, "pfo_peer_open": 0, "pfo_peer_close": 0, "pfo_peer_free": 0,
                      "pfo_peer_max_ranks": 0, "pfo_peer_max_barriers": 0, "pfo_peer_header_bytes": 0,
                      "pfo_compact_workspace_ints": 0, "pfo_wgrad_workspace_floats": 0,
                      "pfo_attn_nbr_bwd_workspace_floats": 0}


_WGRAD_FUSED = os.environ.get("PFO_WGRAD_FUSED", "0")[:1] == "1"


class PfoError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built: no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PfoError(f"{LIB_PATH} is missing: run `python -m pfotgnrec_b200.build` "
                       "(or __graft_entry__.build()); there is no CPU fallback")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)      # AttributeError if the header and the library disagree
        fn.restype, fn.argtypes = res, args
    if lib.pfo_abi_version() != ABI_VERSION:
        raise PfoError("libpfo_b200.so ABI version mismatch")
    _lib = lib
    return lib


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def use_device(device):
    """Make `device` the process's current CUDA device.  Every entry point launches on the current stream of the
    CURRENT device (and cudaFuncSetAttribute / the SM count are per current device), so whoever binds this package to
    cuda:N -- the trainer, the engine, the drop-in TGN, the command-line driver -- selects it first; kernels launched
    for tensors of another device would otherwise run on device 0 against device-N pointers."""
    dev = torch.device(device)
    if dev.type == "cuda" and torch.cuda.is_available():
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        if torch.cuda.current_device() != idx:
            torch.cuda.set_device(idx)
    return dev


def call(name, *args):
    """Invoke an entry point on the current torch stream and raise on a non-zero return."""
    global LAUNCHES
    lib = load()
    rc = getattr(lib, name)(*args, stream())
    if name == "pfo_compact_nodes":          # one CTA for small id spaces (<= 1024 bitmap words), three kernels beyond
        LAUNCHES += 1 if (int(args[1]) + 31) // 32 <= 1024 else 3
    elif name == "pfo_wgrad_tf32":           # the slab reduction is a second launch unless PFO_WGRAD_FUSED=1
        LAUNCHES += 1 if _WGRAD_FUSED else 2
    else:
        LAUNCHES += _LAUNCHES_PER_CALL.get(name, 1)
    if rc != 0:
        raise PfoError(f"{name} failed with cudaError {rc}")


class _NoRange:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NO_RANGE = _NoRange()
NVTX = os.environ.get("PFO_NVTX", "0") not in ("", "0")


class nvtx_range:
    """`with nvtx_range("stage"):` -- an NVTX range around a stage of the step (sampling, node table, attention, BPR,
    backward, optimiser, ...) when PFO_NVTX=1, so that ncu / nsys group the launches by stage; a no-op object otherwise
    (the ranges are host-side markers: they are not captured into CUDA graphs, profile with --no-graph)."""

    def __new__(cls, name):
        if not NVTX:
            return _NO_RANGE
        return object.__new__(cls)

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        torch.cuda.nvtx.range_push(self.name)
        return self

    def __exit__(self, *exc):
        torch.cuda.nvtx.range_pop()
        return False


def query(name, *args):
    """Host-only helper entry points (workspace sizes)."""
    return getattr(load(), name)(*args)
