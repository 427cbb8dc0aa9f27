"""ctypes binding of libgsp.so — the only way the Python layer reaches the CUDA kernels.

The prototypes mirror include/gsp.h one to one. There is no CPU fallback: if the library is missing or
no CUDA device is present, every compute entry point raises (loudly) instead of degrading.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libgsp.so")

SELECT_BINS = 2048
SELECT_PASSES = 6
SELECT_STATE_BYTES = 16384
SELECT_SLOT_WORDS = 2050
SELECT_SCRATCH_BYTES = 18944


class GspError(RuntimeError):
    pass


class GraphInfo(C.Structure):
    _fields_ = [
        ("num_nodes", C.c_int64), ("num_input_edges", C.c_int64), ("nnz", C.c_int64),
        ("num_undirected", C.c_int64), ("max_degree", C.c_int64), ("sum_degree_sq", C.c_double),
        ("symmetric", C.c_int32), ("input_canonical", C.c_int32), ("unit_weights", C.c_int32),
        ("device", C.c_int32),
    ]


_P = C.c_void_p
_I64 = C.c_int64
_I32 = C.c_int32
_INT = C.c_int
_F64 = C.c_double

# name -> (restype, argtypes); every symbol include/gsp.h declares
PROTOTYPES = {
    "gsp_version": (_INT, []),
    "gsp_last_error": (C.c_char_p, []),
    "gsp_launch_count": (C.c_uint64, []),
    "gsp_trim_scratch": (_INT, []),
    "gsp_set_allocator": (_INT, [_P, _P]),
    "gsp_graph_create": (_INT, [_I64, _I64, _P, _P, _P, _P, C.POINTER(_P)]),
    "gsp_graph_destroy": (None, [_P]),
    "gsp_graph_get_info": (_INT, [_P, C.POINTER(GraphInfo)]),
    "gsp_graph_export": (_INT, [_P, _P, _P, _P, _P, _P]),
    "gsp_graph_degrees": (_INT, [_P, _P, _P]),
    "gsp_graph_undirected_ids": (_INT, [_P, _P, _P]),
    "gsp_jaccard": (_INT, [_P, _I64, _I64, _P, _P, _P]),
    "gsp_adamic_adar": (_INT, [_P, _P, _I64, _I64, _P, _P]),
    "gsp_jaccard_adamic_adar": (_INT, [_P, _P, _I64, _I64, _P, _P, _P, _P]),
    "gsp_aa_node_weights": (_INT, [_P, _P, _P]),
    "gsp_aa_node_weights_from_table": (_INT, [_P, _P, _I64, _P, _P]),
    "gsp_copy_f64": (_INT, [_P, _P, _I64, _P]),
    "gsp_jaccard_owned": (_INT, [_P, _I64, _I64, _P, _P, _P]),
    "gsp_adamic_adar_owned": (_INT, [_P, _P, _I64, _I64, _P, _P]),
    "gsp_jaccard_adamic_adar_owned": (_INT, [_P, _P, _I64, _I64, _P, _P, _P]),
    "gsp_owner_costs": (_INT, [_P, _P, _P]),
    "gsp_graph_set_owner_deal": (_INT, [_P, _P, _I32, _P]),
    "gsp_jaccard_owned_scatter": (_INT, [_P, _I64, _I64, _P, _I32, _I64, _P]),
    "gsp_adamic_adar_owned_scatter": (_INT, [_P, _P, _I64, _I64, _P, _I32, _I64, _P]),
    "gsp_jaccard_adamic_adar_owned_scatter": (_INT, [_P, _P, _I64, _I64, _P, _P, _I32, _I64, _P]),
    "gsp_gcn_norm": (_INT, [_I64, _I64, _P, _P, _P, _P, _P, _P, _P, _P]),
    "gsp_target_order": (_INT, [_I64, _I64, _P, _P, _P, _P]),
    "gsp_gcn_propagate": (_INT, [_I64, _P, _P, _P, _P, _P, _I32, _I64, _P, _I64, _P]),
    "gsp_node_triangles": (_INT, [_P, _P, _P, _P, _P]),
    "gsp_connected_components": (_INT, [_P, _P, C.POINTER(C.c_int32), _P]),
    "gsp_degree_product": (_INT, [_P, _I64, _I64, _P, _P]),
    "gsp_featcos_normalize_f32": (_INT, [_I64, _I32, _P, _I64, _P, _I64, _P]),
    "gsp_featcos_f32": (_INT, [_P, _P, _I32, _I64, _I64, _I64, _P, _P]),
    "gsp_featcos_normalize_f32_packed": (_INT, [_I64, _I32, _P, _I64, _P, _I64, _P]),
    "gsp_featcos_f32_packed": (_INT, [_P, _P, _I32, _I64, _I64, _I64, _P, _P]),
    "gsp_featcos_normalize_f64": (_INT, [_I64, _I32, _P, _I64, _P, _I64, _P]),
    "gsp_featcos_f64": (_INT, [_P, _P, _I32, _I64, _I64, _I64, _P, _P]),
    "gsp_approx_er_partial": (_INT, [_P, _P, _I64, _I32, _I32, _F64, _F64, _I64, _I64, _P, _P, _P]),
    "gsp_approx_er_partial_philox": (_INT, [_P, C.c_uint64, _I32, _I32, _I32, _I32, _F64, _F64, _I64, _I64, _P, _P, _P]),
    "gsp_philox_projection": (_INT, [C.c_uint64, _I64, _I32, _I32, _I32, _P, _P]),
    "gsp_er_finalize": (_INT, [_P, _I64, _P]),
    "gsp_laplacian_solve": (_INT, [_P, _P, _I32, _I32, _F64, _F64, _P, _P, _P]),
    "gsp_sssp_batch": (_INT, [_P, _P, _I64, _I32, _P, _I32, C.POINTER(C.c_int32), _P]),
    "gsp_sssp_sources": (_INT, [_P, _P, _P, _I32, _P, _I32, C.POINTER(C.c_int32), _P]),
    "gsp_select_begin": (_INT, [_P, _I64, _INT, _P]),
    "gsp_select_histogram": (_INT, [_P, _I64, _P, _P, _INT, _P, _P]),
    "gsp_select_pick": (_INT, [_P, _P, _INT, _P]),
    "gsp_select_count_ties": (_INT, [_P, _I64, _P, _P, _P, _P]),
    "gsp_select_write_mask": (_INT, [_P, _I64, _P, _P, _P, _P, _INT, _P, _P]),
    "gsp_select_mask": (_INT, [_P, _I64, _I64, _INT, _P, _INT, _P, _P]),
    "gsp_select_compact": (_INT, [_P, _I64, _I64, _INT, _P, _I64, _P, _P, _I64, _P, _INT, _P, _P]),
    "gsp_select_histogram_slot": (_INT, [_P, _I64, _P, _INT, _P, _P]),
    "gsp_select_pick_slots": (_INT, [_P, _P, _I32, _INT, _P]),
    "gsp_select_tally": (_INT, [_P, _I64, _P, _P, _P, _P]),
    "gsp_select_emit": (_INT, [_P, _I64, _P, _P, _P, _I32, _I32, _P, _I64, _P, _P, _I64, _P, _INT, _P, _P]),
    "gsp_degree_aware_guarantee": (_INT, [_P, _P, _I64, _I64, _I32, _P, _P, _P]),
    "gsp_compact_edges": (_INT, [_P, _I64, _I64, _P, _P, _INT, _P, _I64, _P, _P, _P]),
}

_lib = None


def load():
    """Load libgsp.so (no device needed for loading; compute calls need CUDA)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GspError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback for the sparsification engine."
            )
        lib = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in PROTOTYPES.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
        _install_torch_allocator(lib)
    return _lib


_ALLOC_CB = C.CFUNCTYPE(C.c_void_p, C.c_size_t, C.c_int, C.c_void_p)
_FREE_CB = C.CFUNCTYPE(None, C.c_void_p)
_allocator_callbacks = None      # keeps the ctypes thunks alive for the life of the process


def _install_torch_allocator(lib) -> None:
    """Hand torch's caching allocator to libgsp (include/gsp.h, gsp_set_allocator): graph arrays and scratch then come out
    of — and go back to — the pool torch manages, so nothing the library frees stays invisible to the training that
    follows, and building a graph per call recycles blocks instead of paying cudaMalloc / cudaFree of gigabytes.
    GSP_ALLOCATOR=cuda keeps the library's own cudaMalloc + private scratch pool."""
    global _allocator_callbacks
    if os.environ.get("GSP_ALLOCATOR", "torch") != "torch" or not torch.cuda.is_available():
        return

    def alloc(nbytes, device, stream):
        try:
            return torch.cuda.caching_allocator_alloc(int(nbytes), int(device), int(stream or 0))
        except Exception:
            return None

    def free(ptr):
        try:
            torch.cuda.caching_allocator_delete(ptr)
        except Exception:
            pass

    _allocator_callbacks = (_ALLOC_CB(alloc), _FREE_CB(free))
    lib.gsp_set_allocator(C.cast(_allocator_callbacks[0], C.c_void_p), C.cast(_allocator_callbacks[1], C.c_void_p))


def require_cuda() -> None:
    if not torch.cuda.is_available():
        raise GspError("the sparsification engine needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def check(rc: int) -> None:
    if rc != 0:
        msg = load().gsp_last_error()
        raise GspError(f"libgsp error {rc}: {msg.decode() if msg else '?'}")


def ptr(t):
    """Device (or host) address of a tensor, or NULL."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
