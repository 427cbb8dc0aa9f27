"""`calculate_*_scores(adj, ...)` — drop-ins for reference `src/sparsification/metrics.py:17-358`.

Same names, argument meaning, return type (float64 ndarray in `adj.nonzero()` order) and error behaviour as
the reference functions, but the SciPy matrix is only a container: it is uploaded once and every score is
computed by libgsp.so on the GPU. No CPU fallback.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import _lib
from .engine import DeviceGraph


def _device(device=None) -> torch.device:
    _lib.require_cuda()
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def graph_from_scipy(adj, device=None) -> DeviceGraph:
    """Canonical device graph of a SciPy sparse matrix (entries == `adj.nonzero()`, values kept)."""
    dev = _device(device)
    coo = adj.tocoo()
    val = np.asarray(coo.data, dtype=np.float64)
    keep = val != 0                                   # adj.nonzero() drops explicit zeros
    row = torch.from_numpy(np.ascontiguousarray(coo.row[keep], dtype=np.int64))
    col = torch.from_numpy(np.ascontiguousarray(coo.col[keep], dtype=np.int64))
    ei = torch.stack([row, col]).to(dev)
    values = None if np.all(val[keep] == 1.0) else torch.from_numpy(np.ascontiguousarray(val[keep])).to(dev)
    if values is not None and bool((values < 0).any()):
        raise ValueError("negative adjacency entries are not supported")
    return DeviceGraph(ei, adj.shape[0], values)


def _numpy_aa_weights(g: DeviceGraph) -> torch.Tensor:
    # reference metrics.py:104-108 evaluated once per distinct degree (libm-defined constants), gathered on device
    return g.aa_node_weights_numpy()


def calculate_jaccard_scores(adj) -> np.ndarray:
    """reference metrics.py:17-64."""
    return graph_from_scipy(adj).jaccard().cpu().numpy()


def calculate_adamic_adar_scores(adj) -> np.ndarray:
    """reference metrics.py:67-121."""
    g = graph_from_scipy(adj)
    return g.adamic_adar(_numpy_aa_weights(g)).cpu().numpy()


def calculate_feature_cosine_scores(adj, features: np.ndarray) -> np.ndarray:
    """reference metrics.py:301-358 (runs in the dtype of `features`)."""
    g = graph_from_scipy(adj)
    x = torch.from_numpy(np.ascontiguousarray(features))
    return g.feature_cosine(g.normalize_features(x)).cpu().numpy()


def jl_dimension(n: int, epsilon: float) -> int:
    """k = 24 ln(n) / eps^2 (reference metrics.py:248)."""
    return max(int(24 * np.log(max(n, 2)) / (epsilon ** 2)), 1)


def _column_block(n: int, k: int, budget_bytes: float = 24e9) -> int:
    """Columns solved together: four fp64 [n, kb] CG vectors must fit the scratch budget."""
    kb = int(budget_bytes // (32 * max(n, 1)))
    return max(1, min(k, max(kb, 32)))


def _approx_er_on_graph(g: DeviceGraph, epsilon=0.3, seed=42, max_cg_iters=500, cg_tol=1e-6, k: Optional[int] = None,
                        projection=None, group=None, return_iters=False):
    """ApproxER on a device graph. `projection`: None / "numpy" (the reference's PCG64 matrix, generated on the
    host exactly like metrics.py:232,272), "device" (torch Philox normals on the GPU, same distribution,
    no host work: Philox entries generated inside the projection kernel, the [m, k] matrix is never materialised), or an
    explicit [m, k] array/tensor. With a process `group` the projection columns are split over its ranks and the per-edge
    partial sums are all-reduced (NCCL)."""
    m = g.num_undirected
    nnz = g.nnz
    dev = g.device
    if m == 0:
        out = torch.zeros(nnz, dtype=torch.float64, device=dev)
        return (out, torch.zeros(0, dtype=torch.int32, device=dev)) if return_iters else out
    if projection is not None and not isinstance(projection, str):
        R = torch.as_tensor(projection, dtype=torch.float64)
        k = R.size(1)
    else:
        if k is None:
            k = jl_dimension(g.num_nodes, epsilon)
        if projection == "device":
            R = None
        else:
            R = torch.from_numpy(np.random.default_rng(seed).standard_normal((m, k)) / np.sqrt(k))
    rank, world = 0, 1
    if group is not None:
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    c_lo, c_hi = (k * rank) // world, (k * (rank + 1)) // world
    total = torch.zeros(nnz, dtype=torch.float64, device=dev)
    iters = torch.zeros(k, dtype=torch.int32, device=dev)
    kb = _column_block(g.num_nodes, max(c_hi - c_lo, 1))
    for c0 in range(c_lo, c_hi, kb):
        c1 = min(c0 + kb, c_hi)
        if R is None:
            part, it = g.approx_er_partial_philox(seed, c0, c1 - c0, k, max_cg_iters, cg_tol, 1e-6, return_iters=True)
        else:
            block = R[:, c0:c1]
            if not block.is_cuda:
                block = block.contiguous().to(dev, non_blocking=True)
            part, it = g.approx_er_partial(block, max_cg_iters, cg_tol, 1e-6, return_iters=True)
        total += part
        iters[c0:c1] = it
    if group is not None and world > 1:
        dist.all_reduce(total, group=group)
        dist.all_reduce(iters, group=group)
    g.er_finalize(total)
    return (total, iters) if return_iters else total


def calculate_approx_effective_resistance_scores(adj, epsilon: float = 0.3, seed: int = 42, max_cg_iters: int = 500,
                                                 cg_tol: float = 1e-6, k: Optional[int] = None, projection=None):
    """reference metrics.py:178-298. `k` / `projection` are extensions (BASELINE.json config 4 fixes k = 64)."""
    g = graph_from_scipy(adj)
    return _approx_er_on_graph(g, epsilon, seed, max_cg_iters, cg_tol, k, projection).cpu().numpy()


def _exact_er_on_graph(g: DeviceGraph, rtol: float = 1e-11, max_iters: int = 0) -> torch.Tensor:
    """Exact effective resistance R(u,v) = L+[u,u] + L+[v,v] - 2 L+[u,v] (reference metrics.py:158-175, which forms the
    dense pseudo-inverse of L + 1e-10 I: O(n^3)). Here only the columns of L+ that the edges need are computed, by the
    batched Laplacian CG of the ApproxER path (`gsp_laplacian_solve`) with one right-hand side per node,
    b_j = e_j - 1_C(j)/|C(j)| (the component's constant vector projected out, so the 1e10-sized null-space term of the
    reference's regularised inverse — which cancels in R anyway — never enters the arithmetic). Small graphs only, like
    the reference: n^2 fp64 per column block. Asymmetric patterns (outside the reference's contract) keep the dense path."""
    n = g.num_nodes
    if n > 20000:
        raise ValueError("exact effective resistance is O(n^2) memory and n solves; use approx_er for large graphs")
    indptr, indices, data, rows = g.export(with_data=True, with_rows=True)
    rows, cols = rows.long(), indices.long()
    if not g.symmetric:
        lap = torch.zeros((n, n), dtype=torch.float64, device=g.device)
        lap.index_put_((rows, cols), -data, accumulate=True)
        deg = torch.zeros(n, dtype=torch.float64, device=g.device).index_add_(0, rows, data)
        lap += torch.diag(deg + 1e-10)
        pinv = torch.linalg.pinv(lap)
        return torch.clamp_min(pinv[rows, rows] + pinv[cols, cols] - 2.0 * pinv[rows, cols], 1e-10)
    from .topology import connected_components

    label, _ = connected_components(g)
    label = label.long()
    size = torch.bincount(label, minlength=n).to(torch.float64)
    inv_size = 1.0 / size[label]                                  # per node: 1 / |its component|
    z_diag = torch.zeros(n, dtype=torch.float64, device=g.device)
    z_edge = torch.zeros(g.nnz, dtype=torch.float64, device=g.device)
    block = max(1, min(n, int(2e9 // (8 * max(n, 1)))))            # <= 2 GB per [n, block] fp64 vector (the solver keeps five)
    iters_cap = max_iters if max_iters > 0 else max(2000, 8 * n)
    for c0 in range(0, n, block):
        c1 = min(n, c0 + block)
        j = torch.arange(c0, c1, device=g.device)
        rhs = -(label.unsqueeze(1) == label[j].unsqueeze(0)).to(torch.float64) * inv_size[j].unsqueeze(0)
        rhs[j, j - c0] += 1.0
        x = g.laplacian_solve(rhs, max_iters=iters_cap, rtol=rtol, reg=1e-10)
        z_diag[j] = x[j, j - c0]
        sel = (cols >= c0) & (cols < c1)
        z_edge[sel] = x[rows[sel], cols[sel] - c0]
        del rhs, x
    return torch.clamp_min(z_diag[rows] + z_diag[cols] - 2.0 * z_edge, 1e-10)


def calculate_effective_resistance_scores(adj) -> np.ndarray:
    """reference metrics.py:124-175."""
    return _exact_er_on_graph(graph_from_scipy(adj)).cpu().numpy()
