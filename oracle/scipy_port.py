"""TEST INFRASTRUCTURE ONLY — NumPy/SciPy port of the reference's CPU path.

A compact restatement of *how the reference computes* (same SciPy/NumPy calls
in the same order, so the same third-party arithmetic: `scipy.sparse` SpGEMM,
`scipy.sparse.linalg.cg`, NumPy pairwise sums, `np.argsort`, PCG64). It exists

  * as the timed CPU baseline of `bench.py` (`cpu_baseline.kind == "port"`,
    and the `--impl reference` arm): it has the reference's cost model
    (SpGEMM fill-in, k sequential CG solves, O(N*E) degree-aware scans are
    replaced by their closed form only where stated), and
  * as a second checker beside the plain-C restatement `gsp_oracle.c`.

Pinned against the live reference by `tests/test_oracle_pin.py` (build
container) and against `tests/golden/*.npz` (everywhere).
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu-baseline legs may
import this module; the product package never does.

Third-party arithmetic this relies on (un-vendored by the reference,
`pyproject.toml:28-32`): numpy 2.3.5, scipy 1.18.1 in this image.
"""
from __future__ import annotations

import warnings

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


# --------------------------------------------------------------------------- CSR
def build_adjacency(edge_index: np.ndarray, num_nodes: int) -> sp.csr_matrix:
    """reference core.py:70-74 — canonical CSR, duplicates summed into `data`."""
    e = edge_index.shape[1]
    return sp.csr_matrix((np.ones(e), (edge_index[0], edge_index[1])), shape=(num_nodes, num_nodes))


def _binary(adj):
    return (adj > 0).astype(np.float64)


# ------------------------------------------------------------------------ scoring
def jaccard(adj: sp.csr_matrix) -> np.ndarray:
    """reference metrics.py:43-64."""
    ab = _binary(adj)
    deg = np.asarray(ab.sum(axis=1)).ravel()
    two_hop = ab @ ab
    r, c = ab.nonzero()
    inter = np.asarray(two_hop[r, c]).ravel()
    union = deg[r] + deg[c] - inter
    out = np.zeros_like(inter, dtype=np.float64)
    np.divide(inter, union, out=out, where=union > 0)
    return np.nan_to_num(out, nan=0.0, posinf=0.0, neginf=0.0)


def adamic_adar_node_weights(deg: np.ndarray) -> np.ndarray:
    """reference metrics.py:104-108 — w = 1/sqrt(max(log(deg+1), 1e-10)) (NumPy libm)."""
    return 1.0 / np.sqrt(np.maximum(np.log(deg + 1), 1e-10))


def adamic_adar(adj: sp.csr_matrix) -> np.ndarray:
    """reference metrics.py:99-121."""
    ab = _binary(adj)
    deg = np.asarray(ab.sum(axis=1)).ravel()
    w = adamic_adar_node_weights(deg)
    scaled = ab @ sp.diags(w, format="csr")
    prod = scaled @ scaled.T
    r, c = ab.nonzero()
    return np.nan_to_num(np.asarray(prod[r, c]).ravel(), nan=0.0, posinf=0.0, neginf=0.0)


def feature_cosine(adj: sp.csr_matrix, features: np.ndarray) -> np.ndarray:
    """reference metrics.py:344-358 — runs in the dtype of `features`."""
    nrm = np.maximum(np.linalg.norm(features, axis=1, keepdims=True), 1e-10)
    unit = features / nrm
    r, c = adj.nonzero()
    dots = np.sum(unit[r] * unit[c], axis=1)
    return np.maximum(dots, 0.0).astype(np.float64)


def degree_product(adj: sp.csr_matrix) -> np.ndarray:
    """reference core.py:167-172 (uses raw adjacency values, i.e. multiplicities)."""
    deg = np.asarray(adj.sum(axis=1)).ravel()
    r, c = adj.nonzero()
    return deg[r] * deg[c]


def jl_dimension(n: int, epsilon: float) -> int:
    """reference metrics.py:248."""
    return max(int(24 * np.log(max(n, 2)) / (epsilon ** 2)), 1)


def projection_matrix(m: int, k: int, seed: int) -> np.ndarray:
    """reference metrics.py:232,272 — PCG64 standard normals, row-major fill, / sqrt(k)."""
    return np.random.default_rng(seed).standard_normal((m, k)) / np.sqrt(k)


def approx_effective_resistance(adj, epsilon=0.3, seed=42, max_cg_iters=500, cg_tol=1e-6,
                                k=None, projection=None, return_iters=False):
    """reference metrics.py:232-298. `k` / `projection` are extensions (BASELINE config 4 fixes k=64)."""
    n = adj.shape[0]
    rows, cols = adj.nonzero()
    upper = rows < cols
    ue, ve = rows[upper], cols[upper]
    m = len(ue)
    if m == 0:
        z = np.zeros(len(rows), dtype=np.float64)
        return (z, np.zeros(0, dtype=np.int32)) if return_iters else z
    if projection is not None:
        R = np.asarray(projection, dtype=np.float64)
        k = R.shape[1]
    else:
        if k is None:
            k = jl_dimension(n, epsilon)
        R = projection_matrix(m, k, seed)
    deg = np.array(adj.sum(axis=1)).ravel()
    lap = sp.diags(deg, format="csr") - adj + 1e-6 * sp.eye(n, format="csr")
    ar = np.arange(m)
    inc = sp.csr_matrix((np.concatenate([np.ones(m), -np.ones(m)]),
                         (np.concatenate([ue, ve]), np.concatenate([ar, ar]))), shape=(n, m))
    Y = inc @ R
    Z = np.zeros((n, k), dtype=np.float64)
    iters = np.zeros(k, dtype=np.int32)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        for j in range(k):
            count = [0]
            z, info = spla.cg(lap, Y[:, j], maxiter=max_cg_iters, rtol=cg_tol,
                              callback=(lambda _x, c=count: c.__setitem__(0, c[0] + 1)) if return_iters else None)
            iters[j] = count[0]
            if info != 0 or np.any(np.isnan(z)):
                z = np.nan_to_num(z, nan=0.0, posinf=0.0, neginf=0.0)
            Z[:, j] = z
    diff = Z[rows] - Z[cols]
    r_eff = np.sum(diff ** 2, axis=1)
    r_eff = np.maximum(np.nan_to_num(r_eff, nan=1e-10, posinf=1e-10, neginf=1e-10), 1e-10).astype(np.float64)
    return (r_eff, iters) if return_iters else r_eff


# ---------------------------------------------------------------------- selection
def threshold_mask(scores: np.ndarray, num_edges: int, retention_ratio: float, keep_lowest=False,
                   stable=True) -> np.ndarray:
    """reference core.py:232-240 (with the stable-sort tie contract when `stable`)."""
    num_keep = int(num_edges * retention_ratio)
    order = np.argsort(scores, kind="stable" if stable else None)
    chosen = order[:num_keep] if keep_lowest else order[-num_keep:]
    mask = np.zeros(num_edges, dtype=bool)
    mask[chosen] = True
    return mask


def sampled_mask(scores: np.ndarray, num_edges: int, retention_ratio: float, seed=42) -> np.ndarray:
    """reference core.py:333-349."""
    rng = np.random.default_rng(seed)
    floor = 1e-8
    s = np.nan_to_num(scores, nan=floor, posinf=floor, neginf=floor)
    p = np.maximum(s, floor)
    p = p / p.sum()
    chosen = rng.choice(num_edges, size=int(num_edges * retention_ratio), replace=False, p=p)
    mask = np.zeros(num_edges, dtype=bool)
    mask[chosen] = True
    return mask


def degree_aware_mask_reference_loops(scores, src, num_nodes, num_edges, retention_ratio,
                                      min_edges_per_node=1, stable=True) -> np.ndarray:
    """reference core.py:415-451 with the reference's own O(N*E) + O(E^2) loops (timing baseline)."""
    kind = "stable" if stable else None
    num_keep = int(num_edges * retention_ratio)
    mask = np.zeros(num_edges, dtype=bool)
    for node in range(num_nodes):
        inc = np.where(src == node)[0]
        if len(inc) == 0:
            continue
        k = min(min_edges_per_node, len(inc))
        mask[inc[np.argsort(scores[inc], kind=kind)[-k:]]] = True
    if mask.sum() < num_keep:
        for idx in np.argsort(scores, kind=kind)[::-1]:
            if mask.sum() >= num_keep:
                break
            if not mask[idx]:
                mask[idx] = True
    return mask


def degree_aware_mask(scores, src, num_nodes, num_edges, retention_ratio, min_edges_per_node=1) -> np.ndarray:
    """Closed form of reference core.py:415-451 under the stable-sort contract."""
    num_keep = int(num_edges * retention_ratio)
    pos = np.arange(num_edges)
    # per source node: last min(m, deg) entries of the stable ascending sort of its out-edge scores
    order = np.lexsort((pos, scores[:num_edges], src))  # by src, then score, then position
    s_sorted = src[order]
    seg_end = np.r_[s_sorted[1:] != s_sorted[:-1], True] if num_edges else np.zeros(0, bool)
    # rank from the end of each segment
    idx_in = np.arange(num_edges)
    ends = np.flatnonzero(seg_end)
    seg_id = np.searchsorted(ends, idx_in, side="left")
    from_end = ends[seg_id] - idx_in
    mask = np.zeros(num_edges, dtype=bool)
    mask[order[from_end < min_edges_per_node]] = True
    have = int(mask.sum())
    if have < num_keep:
        desc = np.argsort(scores, kind="stable")[::-1]
        fill = desc[~mask[desc]][: num_keep - have]
        mask[fill] = True
    return mask


# ----------------------------------------------------------------- random baseline
def precompute_random_scores(edge_index: np.ndarray, num_nodes: int, seed=42):
    """reference random.py:23-33."""
    lo = np.minimum(edge_index[0], edge_index[1]).astype(np.int64)
    hi = np.maximum(edge_index[0], edge_index[1]).astype(np.int64)
    _, inverse = np.unique(lo * (num_nodes + 1) + hi, return_inverse=True)
    n_und = int(inverse.max()) + 1
    return np.random.default_rng(seed).random(n_und), inverse


def random_mask(undirected_scores, inverse_idx, retention_ratio, stable=True) -> np.ndarray:
    """reference random.py:45-50."""
    n_und = len(undirected_scores)
    n_keep = max(1, int(n_und * retention_ratio))
    keep = np.zeros(n_und, dtype=bool)
    keep[np.argsort(undirected_scores, kind="stable" if stable else None)[-n_keep:]] = True
    return keep[inverse_idx]


def minmax_edge_weight(all_scores, mask, keep_lowest=False) -> np.ndarray:
    """reference scripts/nb05_roman_empire/roman_empire_gpu.py:248-254 (float32 result)."""
    s = all_scores[mask]
    mn, mx = s.min(), s.max()
    w = (s - mn) / (mx - mn + 1e-8)
    if keep_lowest:
        w = 1.0 - w
    return w.astype(np.float32)
