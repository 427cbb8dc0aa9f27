"""TEST INFRASTRUCTURE ONLY — ctypes front-end of the plain-C oracle (`gsp_oracle.c`).

Exposes the oracle with the reference's own function names and argument meaning
(reference `src/sparsification/__init__.py:16-28`) so parity tests read like the
reference's tests. Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
cpu-baseline legs may import this module; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libgsp_oracle.so")
_lib = None

_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "gsp_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp = C.c_void_p
        L.gspo_csr_build.restype = C.c_int64
        L.gspo_csr_build.argtypes = [C.c_int64, C.c_int64, _i64p, _i64p, vp, _i64p, _i32p, _f64p]
        L.gspo_transpose.restype = None
        L.gspo_transpose.argtypes = [C.c_int64, _i64p, _i32p, _i64p, _i32p]
        L.gspo_jaccard.restype = None
        L.gspo_jaccard.argtypes = [C.c_int64, _i64p, _i32p, vp, vp, vp, _f64p]
        L.gspo_adamic_adar.restype = None
        L.gspo_adamic_adar.argtypes = [C.c_int64, _i64p, _i32p, _f64p, _f64p]
        L.gspo_degree_product.restype = None
        L.gspo_degree_product.argtypes = [C.c_int64, _i64p, _i32p, vp, _f64p]
        L.gspo_pairwise_sum_f32.restype = C.c_float
        L.gspo_pairwise_sum_f32.argtypes = [_f32p, C.c_int64]
        L.gspo_pairwise_sum_f64.restype = C.c_double
        L.gspo_pairwise_sum_f64.argtypes = [_f64p, C.c_int64]
        L.gspo_featcos_f32.restype = None
        L.gspo_featcos_f32.argtypes = [C.c_int64, C.c_int64, _f32p, _i64p, _i32p, _f64p, vp]
        L.gspo_featcos_f64.restype = None
        L.gspo_featcos_f64.argtypes = [C.c_int64, C.c_int64, _f64p, _i64p, _i32p, _f64p, vp]
        L.gspo_select.restype = None
        L.gspo_select.argtypes = [C.c_int64, _f64p, C.c_int64, C.c_int64, C.c_int, _u8p]
        L.gspo_degree_aware.restype = None
        L.gspo_degree_aware.argtypes = [C.c_int64, _i64p, _f64p, C.c_int64, C.c_int64, C.c_int64, _u8p]
        L.gspo_count_upper.restype = C.c_int64
        L.gspo_count_upper.argtypes = [C.c_int64, _i64p, _i32p]
        L.gspo_approx_er.restype = C.c_int
        L.gspo_approx_er.argtypes = [C.c_int64, _i64p, _i32p, vp, C.c_int64, _f64p, C.c_int64, C.c_double,
                                     C.c_double, _f64p, vp, vp]
        L.gspo_metric_backbone.restype = C.c_int
        L.gspo_metric_backbone.argtypes = [C.c_int64, C.c_int64, _i64p, _i64p, _f64p, C.c_double, _u8p]
        _lib = L
    return _lib


def _vp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class CSR:
    """Canonical CSR (int64 indptr, int32 indices, float64 data=multiplicity)."""

    def __init__(self, n, indptr, indices, data):
        self.n, self.indptr, self.indices, self.data = int(n), indptr, indices, data
        self.nnz = int(indptr[-1])

    @property
    def rows(self):
        return np.repeat(np.arange(self.n, dtype=np.int64), np.diff(self.indptr))

    def is_symmetric(self):
        tptr, tidx = self.transpose()
        return np.array_equal(tptr, self.indptr) and np.array_equal(tidx, self.indices)

    def transpose(self):
        tptr = np.zeros(self.n + 1, np.int64)
        tidx = np.zeros(max(self.nnz, 1), np.int32)
        lib().gspo_transpose(self.n, self.indptr, self.indices, tptr, tidx)
        return tptr, tidx[: self.nnz]


def csr_from_edge_index(edge_index: np.ndarray, num_nodes: int, values=None) -> CSR:
    """reference core.py:70-74."""
    row = np.ascontiguousarray(edge_index[0], dtype=np.int64)
    col = np.ascontiguousarray(edge_index[1], dtype=np.int64)
    e = len(row)
    indptr = np.zeros(num_nodes + 1, np.int64)
    indices = np.zeros(max(e, 1), np.int32)
    data = np.zeros(max(e, 1), np.float64)
    val = None if values is None else np.ascontiguousarray(values, dtype=np.float64)
    nnz = lib().gspo_csr_build(num_nodes, e, row, col, _vp(val), indptr, indices, data)
    if nnz < 0:
        raise ValueError("edge index out of range")
    return CSR(num_nodes, indptr, indices[:nnz].copy(), data[:nnz].copy())


def csr_from_scipy(adj) -> CSR:
    """Canonicalise a SciPy sparse matrix exactly like `adj.nonzero()` order (explicit zeros dropped)."""
    coo = adj.tocoo()
    return csr_from_edge_index(np.vstack([coo.row, coo.col]), adj.shape[0], values=coo.data.astype(np.float64))


def _as_csr(adj) -> CSR:
    return adj if isinstance(adj, CSR) else csr_from_scipy(adj)


# ---- reference-named scoring functions (metrics.py:17,67,178,301) -----------------------
def calculate_jaccard_scores(adj, return_counts=False):
    g = _as_csr(adj)
    inter = np.zeros(max(g.nnz, 1), np.int32)
    out = np.zeros(max(g.nnz, 1), np.float64)
    if g.is_symmetric():
        lib().gspo_jaccard(g.n, g.indptr, g.indices, None, None, _vp(inter), out)
    else:
        tptr, tidx = g.transpose()
        tidx = np.ascontiguousarray(tidx)
        lib().gspo_jaccard(g.n, g.indptr, g.indices, _vp(tptr), _vp(tidx), _vp(inter), out)
    return (out[: g.nnz], inter[: g.nnz]) if return_counts else out[: g.nnz]


def adamic_adar_node_weights(deg: np.ndarray) -> np.ndarray:
    """reference metrics.py:104-108 (NumPy expression; bits are libm-defined)."""
    return 1.0 / np.sqrt(np.maximum(np.log(deg.astype(np.float64) + 1), 1e-10))


def calculate_adamic_adar_scores(adj, node_weights=None):
    g = _as_csr(adj)
    w = adamic_adar_node_weights(np.diff(g.indptr)) if node_weights is None else np.ascontiguousarray(node_weights)
    out = np.zeros(max(g.nnz, 1), np.float64)
    lib().gspo_adamic_adar(g.n, g.indptr, g.indices, w, out)
    return out[: g.nnz]


def calculate_feature_cosine_scores(adj, features: np.ndarray):
    g = _as_csr(adj)
    x = np.ascontiguousarray(features)
    out = np.zeros(max(g.nnz, 1), np.float64)
    if x.dtype == np.float32:
        lib().gspo_featcos_f32(g.n, x.shape[1], x, g.indptr, g.indices, out, None)
    else:
        lib().gspo_featcos_f64(g.n, x.shape[1], x.astype(np.float64), g.indptr, g.indices, out, None)
    return out[: g.nnz]


def degree_product_scores(adj):
    g = _as_csr(adj)
    out = np.zeros(max(g.nnz, 1), np.float64)
    lib().gspo_degree_product(g.n, g.indptr, g.indices, _vp(g.data), out)
    return out[: g.nnz]


def jl_dimension(n: int, epsilon: float) -> int:
    return max(int(24 * np.log(max(n, 2)) / (epsilon ** 2)), 1)


def calculate_approx_effective_resistance_scores(adj, epsilon=0.3, seed=42, max_cg_iters=500, cg_tol=1e-6,
                                                 k=None, projection=None, return_iters=False):
    g = _as_csr(adj)
    m = lib().gspo_count_upper(g.n, g.indptr, g.indices)
    out = np.zeros(max(g.nnz, 1), np.float64)
    if m == 0:
        return (out[: g.nnz], np.zeros(0, np.int32)) if return_iters else out[: g.nnz]
    if projection is not None:
        R = np.ascontiguousarray(projection, dtype=np.float64)
        k = R.shape[1]
    else:
        if k is None:
            k = jl_dimension(g.n, epsilon)
        R = np.random.default_rng(seed).standard_normal((m, k)) / np.sqrt(k)
    iters = np.zeros(k, np.int32)
    lib().gspo_approx_er(g.n, g.indptr, g.indices, _vp(g.data), k, R, max_cg_iters, cg_tol, 1e-6, out,
                         _vp(iters), None)
    return (out[: g.nnz], iters) if return_iters else out[: g.nnz]


# ---- selection (core.py:232-240, 415-451) ---------------------------------------------
def threshold_mask(scores, num_edges, retention_ratio, keep_lowest=False):
    s = np.ascontiguousarray(scores, dtype=np.float64)
    mask = np.zeros(max(num_edges, 1), np.uint8)
    lib().gspo_select(len(s), s if len(s) else np.zeros(1), num_edges, int(num_edges * retention_ratio),
                      int(keep_lowest), mask)
    return mask[:num_edges].astype(bool)


def degree_aware_mask(scores, src, num_nodes, num_edges, retention_ratio, min_edges_per_node=1):
    s = np.ascontiguousarray(scores, dtype=np.float64)
    if len(s) < num_edges:
        raise IndexError("scores shorter than edge list (duplicate edges): reference raises IndexError")
    mask = np.zeros(max(num_edges, 1), np.uint8)
    lib().gspo_degree_aware(num_edges, np.ascontiguousarray(src, dtype=np.int64), s, num_nodes,
                            int(num_edges * retention_ratio), min_edges_per_node, mask)
    return mask[:num_edges].astype(bool)


def pairwise_sum(a: np.ndarray):
    a = np.ascontiguousarray(a)
    if a.dtype == np.float32:
        return np.float32(lib().gspo_pairwise_sum_f32(a, len(a)))
    return np.float64(lib().gspo_pairwise_sum_f64(a.astype(np.float64), len(a)))


def scores_to_cost(scores: np.ndarray, distance_metric: bool = False) -> np.ndarray:
    """reference core.py:82-116 (similarity -> distance, d = 1/p - 1)."""
    similarity = 1.0 / np.maximum(scores, 1e-10) if distance_metric else scores.copy()
    top = similarity.max()
    if top <= 0:
        return np.ones_like(scores)
    proximity = similarity / top
    positive = proximity[proximity > 0]
    floor = (positive.min() * 0.01) if len(positive) > 0 else 1e-6
    proximity[proximity <= 0] = floor
    return 1.0 / proximity - 1.0


def metric_backbone_mask(edge_index: np.ndarray, num_nodes: int, edge_weights: np.ndarray, epsilon: float = 1e-9):
    """reference metric_backbone.py:58-112: keep-mask over the edge_index columns."""
    e = edge_index.shape[1]
    mask = np.zeros(max(e, 1), np.uint8)
    lib().gspo_metric_backbone(num_nodes, e, np.ascontiguousarray(edge_index[0], dtype=np.int64),
                               np.ascontiguousarray(edge_index[1], dtype=np.int64),
                               np.ascontiguousarray(edge_weights[:e], dtype=np.float64), float(epsilon), mask)
    return mask[:e].astype(bool)
