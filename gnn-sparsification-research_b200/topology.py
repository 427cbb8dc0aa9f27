"""Topology metrics of a (sparsified) graph on the GPU — drop-ins for reference `src/sparsification/metrics.py:361-578`
(`compute_geodesic_preservation`, `compute_topology_metrics`, `compute_topology_preservation`; NetworkX in the reference,
SURVEY §8f-4).

Same names, argument (a symmetric SciPy CSR adjacency), dictionary keys and conventions: undirected edge count with a self
loop counted once, NetworkX degrees (a loop counts twice), average local clustering coefficient, connected components and
the share of the largest one, algebraic connectivity of the graph (of its largest component when disconnected). Triangles
come from the intersection counts of the Jaccard pass (`gsp_node_triangles`), components from `gsp_connected_components`.
The algebraic connectivity of a component of up to `dense_limit` nodes is a dense symmetric eigenproblem on the device
(`torch.linalg.eigvalsh`); larger components go through shift-and-invert Lanczos on the batched Laplacian CG of the
ApproxER path (`_fiedler_shift_invert`), so there is no size beyond which the value is missing.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict

import numpy as np
import torch

from ._lib import check, ptr
from .engine import DeviceGraph
from .metrics import graph_from_scipy


def node_triangles(g: DeviceGraph):
    """(pairs int64[n], degree int32[n]): NetworkX's `t` (= 2 x triangles) and `d` of every node (self loops discounted)."""
    _, inter = g.jaccard(return_counts=True)
    pairs = torch.empty(g.num_nodes, dtype=torch.int64, device=g.device)
    degree = torch.empty(g.num_nodes, dtype=torch.int32, device=g.device)
    with torch.cuda.device(g.device):
        check(g._lib.gsp_node_triangles(g._handle, ptr(inter), ptr(pairs), ptr(degree), g._stream()))
    return pairs, degree


def clustering_coefficients(g: DeviceGraph) -> torch.Tensor:
    """fp64[n] local clustering coefficients `t / (d (d - 1))`, 0 where undefined (nx.clustering)."""
    pairs, degree = node_triangles(g)
    d = degree.to(torch.float64)
    denom = d * (d - 1.0)
    return torch.where((pairs > 0) & (denom > 0), pairs.to(torch.float64) / denom.clamp_min(1.0), torch.zeros_like(d))


def connected_components(g: DeviceGraph):
    """(label int32[n] = smallest node id of each node's component, sweeps used)."""
    label = torch.empty(g.num_nodes, dtype=torch.int32, device=g.device)
    rounds = C.c_int32(0)
    with torch.cuda.device(g.device):
        check(g._lib.gsp_connected_components(g._handle, ptr(label), C.byref(rounds), g._stream()))
    return label, int(rounds.value)


def _fiedler_shift_invert(g: DeviceGraph, member: torch.Tensor, size: int, tol: float = 1e-9, max_steps: int = 60) -> float:
    """Second smallest Laplacian eigenvalue of the component `member` (bool[n]) by shift-and-invert Lanczos: the operator
    is (L + sigma I)^-1 restricted to the component and to the complement of its constant vector, applied with the batched
    CG of the ApproxER path (`gsp_laplacian_solve`, one column); the wanted eigenvalue is the LARGEST of that operator, so a
    few dozen Lanczos steps with full reorthogonalisation settle it where plain Lanczos on L would need thousands on
    chain-like graphs. Weights = the matrix values, like nx.algebraic_connectivity (reference metrics.py:480-511)."""
    dev = g.device
    n = g.num_nodes
    ones = member.to(torch.float64) / np.sqrt(size)
    sigma = 1e-9 * max(float(g.max_degree), 1.0)
    gen = torch.Generator(device=dev)
    gen.manual_seed(0)
    v = torch.randn(n, dtype=torch.float64, device=dev, generator=gen) * member
    v -= torch.dot(ones, v) * ones
    v /= torch.linalg.norm(v)
    basis, alphas, betas = [v], [], []
    lam = float("nan")
    for step in range(max_steps):
        w = g.laplacian_solve(basis[-1].unsqueeze(1), max_iters=200000, rtol=1e-12, reg=sigma).squeeze(1) * member
        w -= torch.dot(ones, w) * ones
        alpha = float(torch.dot(basis[-1], w))
        alphas.append(alpha)
        for _ in range(2):                                        # full reorthogonalisation, twice
            q = torch.stack(basis, dim=1)
            w -= q @ (q.T @ w)
        beta = float(torch.linalg.norm(w))
        t = np.diag(alphas) + np.diag(betas, 1) + np.diag(betas, -1) if betas else np.array([[alpha]])
        theta, vecs = np.linalg.eigh(t)
        new = 1.0 / theta[-1] - sigma
        residual = abs(beta * vecs[-1, -1]) / abs(theta[-1])       # Ritz residual of the largest pair, relative
        if residual < tol or beta < 1e-14 or (step > 0 and abs(new - lam) <= tol * abs(new) and residual < 1e-6):
            return float(new)
        lam = new
        betas.append(beta)
        basis.append(w / beta)
    return float(lam)


def _algebraic_connectivity(g: DeviceGraph, label: torch.Tensor, root: int, size: int, dense_limit: int) -> float:
    if size <= 1:
        return 0.0
    if size > dense_limit:
        return _fiedler_shift_invert(g, label == root, size)
    indptr, indices, data, rows = g.export(with_data=True, with_rows=True)
    nodes = torch.nonzero(label == root).flatten()
    local = torch.full((g.num_nodes,), -1, dtype=torch.int64, device=g.device)
    local[nodes] = torch.arange(size, device=g.device)
    r, c = local[rows.long()], local[indices.long()]
    keep = (r >= 0) & (c >= 0) & (r != c)
    a = torch.zeros((size, size), dtype=torch.float64, device=g.device)
    a[r[keep], c[keep]] = data[keep] if data is not None else 1.0
    a = torch.maximum(a, a.T)
    lap = torch.diag(a.sum(dim=1)) - a
    return float(torch.linalg.eigvalsh(lap)[1])


def topology_metrics_on_graph(g: DeviceGraph, dense_limit: int = 8192, with_connectivity: bool = True) -> Dict:
    if not g.symmetric:
        raise ValueError("topology metrics need a symmetric (undirected) adjacency matrix")
    n = g.num_nodes
    if n == 0:
        return {"num_nodes": 0, "num_edges": 0, "avg_degree": 0.0, "clustering_coefficient": 0.0, "algebraic_connectivity": 0.0,
                "num_connected_components": 0, "largest_component_ratio": 0.0}
    pairs, degree = node_triangles(g)
    loops = g.degrees().to(torch.int64) - degree.to(torch.int64)              # 1 where the row holds the node itself
    num_loops = int(loops.sum())
    num_edges = int(degree.sum(dtype=torch.int64)) // 2 + num_loops
    avg_degree = float((degree.to(torch.int64) + 2 * loops).sum()) / n
    d = degree.to(torch.float64)
    denom = d * (d - 1.0)
    c = torch.where((pairs > 0) & (denom > 0), pairs.to(torch.float64) / denom.clamp_min(1.0), torch.zeros_like(d))
    label, _ = connected_components(g)
    sizes = torch.bincount(label.long(), minlength=n)
    num_components = int((sizes > 0).sum())
    largest = int(sizes.max())
    out = {
        "num_nodes": n, "num_edges": num_edges, "avg_degree": avg_degree,
        "clustering_coefficient": float(c.sum()) / n, "algebraic_connectivity": 0.0,
        "num_connected_components": num_components, "largest_component_ratio": largest / n,
    }
    if with_connectivity and n > 1:
        out["algebraic_connectivity"] = _algebraic_connectivity(g, label, int(sizes.argmax()), largest, dense_limit)
    return out


def compute_topology_metrics(adj, dense_limit: int = 8192) -> Dict:
    """reference metrics.py:445-520. An asymmetric matrix (what sparsify_sampled / sparsify_degree_aware / a directed
    top-k boundary leave behind) is read like the reference's `nx.from_scipy_sparse_array(adj)` reads it: an undirected
    edge wherever either direction is stored (weight: the larger of the two entries)."""
    import scipy.sparse as sp

    adj = sp.csr_matrix(adj)
    if (adj != adj.T).nnz:
        adj = sp.csr_matrix(adj.maximum(adj.T))
    return topology_metrics_on_graph(graph_from_scipy(adj), dense_limit=dense_limit)


def compute_topology_preservation(original_adj, sparse_adj) -> Dict:
    """reference metrics.py:523-578."""
    orig, sparse = compute_topology_metrics(original_adj), compute_topology_metrics(sparse_adj)
    return {
        "edge_retention": sparse["num_edges"] / orig["num_edges"] if orig["num_edges"] > 0 else 0.0,
        "clustering_preservation": (sparse["clustering_coefficient"] / orig["clustering_coefficient"]
                                    if orig["clustering_coefficient"] > 0 else 1.0),
        # (NaN — a component beyond the dense eigen-solver's limit — propagates instead of reading as "0 preserved")
        "connectivity_preservation": (sparse["algebraic_connectivity"] / orig["algebraic_connectivity"]
                                      if not orig["algebraic_connectivity"] <= 0 else 0.0),
        "component_change": sparse["num_connected_components"] - orig["num_connected_components"],
        "original_metrics": orig,
        "sparse_metrics": sparse,
    }


def hop_distances(g: DeviceGraph, sources, scratch_bytes: float = 16e9) -> torch.Tensor:
    """fp64 [len(sources), n] hop counts (inf = unreachable) from the given source nodes: batched (min,+) relaxation with unit
    edge lengths (`gsp_sssp_sources`), the device form of the reference's repeated `nx.shortest_path_length` calls."""
    n = g.num_nodes
    src = torch.as_tensor(list(sources), dtype=torch.int32, device=g.device)
    out = torch.empty((src.numel(), n), dtype=torch.float64, device=g.device)
    if src.numel() == 0 or n == 0:
        return out
    if int(src.min()) < 0 or int(src.max()) >= n:
        raise ValueError("source ids outside [0, num_nodes)")
    ones = torch.ones(max(g.nnz, 1), dtype=torch.float64, device=g.device)
    batch = int(max(1, min(src.numel(), scratch_bytes // (8 * max(n, 1)))))
    rounds = C.c_int32(0)
    for s0 in range(0, src.numel(), batch):
        s1 = min(src.numel(), s0 + batch)
        dist = torch.empty((n, s1 - s0), dtype=torch.float64, device=g.device)       # [n, S]: sources contiguous
        with torch.cuda.device(g.device):
            check(g._lib.gsp_sssp_sources(g._handle, ptr(ones), ptr(src[s0:s1].contiguous()), s1 - s0, ptr(dist), max(n, 1),
                                          C.byref(rounds), g._stream()))
        out[s0:s1] = dist.T
        del dist
    return out


def _undirected_pattern_graph(adj) -> DeviceGraph:
    import scipy.sparse as sp

    pat = sp.csr_matrix(adj) != 0
    pat = sp.csr_matrix((pat + pat.T).astype(np.float64))       # NetworkX builds an undirected graph from either direction
    pat.data[:] = 1.0
    return graph_from_scipy(pat)


def compute_geodesic_preservation(original_adj, sparse_adj, n_samples: int = 500, seed: int = 42) -> Dict:
    """reference metrics.py:361-442: share of sampled node pairs whose hop distance survives sparsification. The pair sample is
    drawn on the host exactly like the reference does (NumPy PCG64: the stream is observable behaviour); the distances
    come from the device."""
    n = original_adj.shape[0]
    rng = np.random.default_rng(seed)
    pairs = set()
    attempts = 0
    while len(pairs) < n_samples and attempts < n_samples * 10:
        u, v = rng.integers(0, n, size=2)
        if u != v:
            pairs.add((min(u, v), max(u, v)))
        attempts += 1
    pairs = list(pairs)
    preserved = increased = disconnected = 0
    increases = []
    if pairs:
        sources = sorted({int(u) for u, _ in pairs})
        column = {u: i for i, u in enumerate(sources)}
        rows = torch.tensor([column[int(u)] for u, _ in pairs], dtype=torch.int64)
        cols = torch.tensor([int(v) for _, v in pairs], dtype=torch.int64)
        def sampled(adj):
            d = hop_distances(_undirected_pattern_graph(adj), sources)
            return d[rows.to(d.device), cols.to(d.device)].cpu().numpy()

        d_orig, d_sparse = sampled(original_adj), sampled(sparse_adj)
        for a, b in zip(d_orig, d_sparse):
            if np.isinf(a):
                continue                         # already disconnected in the original graph
            if np.isinf(b):
                disconnected += 1
            elif b == a:
                preserved += 1
            else:
                increased += 1
                increases.append(int(b - a))
    total_valid = preserved + increased + disconnected
    return {
        "preservation_ratio": preserved / total_valid if total_valid > 0 else 0.0,
        "pairs_tested": len(pairs),
        "preserved_count": preserved,
        "increased_count": increased,
        "disconnected_count": disconnected,
        "avg_distance_increase": np.mean(increases) if increases else 0.0,
        "max_distance_increase": max(increases) if increases else 0,
    }
