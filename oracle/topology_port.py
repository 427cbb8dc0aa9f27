"""TEST INFRASTRUCTURE — NumPy/SciPy restatement of the reference's output-graph topology metrics
(`src/sparsification/metrics.py:445-520` `compute_topology_metrics`, `:523-578` `compute_topology_preservation`), which
the reference evaluates through NetworkX. Only tests/ may import this module. Pinned against the live reference (NetworkX
3.6.1 in the build container) by tests/test_oracle_pin.py and against tests/golden/topology_metrics.json.

Semantics restated (an asymmetric matrix — a sampled / degree-aware sparsified graph — is read as the undirected graph
with an edge wherever either direction is stored):
* `nx.from_scipy_sparse_array(adj)`: one undirected edge per non-zero pair, a diagonal entry is a self loop;
  `number_of_edges` counts a loop once, `degree` counts it twice (metrics.py:461-465).
* `nx.average_clustering`: per node `2 T(v) / (d(d-1))` over the neighbour set WITHOUT the node itself, 0 when d < 2,
  averaged over all nodes (metrics.py:468).
* connected components and the share of the largest one (metrics.py:471-474).
* algebraic connectivity: second smallest Laplacian eigenvalue of the graph, or of its largest component when it is
  disconnected (metrics.py:478-509); weights = the matrix values; self loops cancel in `D - A`.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
from scipy.sparse.csgraph import connected_components


def compute_topology_metrics(adj: sp.csr_matrix, with_connectivity: bool = True) -> dict:
    adj = sp.csr_matrix(adj)
    if (adj != adj.T).nnz:      # nx.from_scipy_sparse_array builds an UNDIRECTED graph: either direction is the edge
        adj = sp.csr_matrix(adj.maximum(adj.T))
    n = adj.shape[0]
    pattern = sp.csr_matrix((np.ones(adj.nnz), adj.indices, adj.indptr), shape=adj.shape)
    diag = pattern.diagonal()
    off = pattern - sp.diags(diag)
    off.eliminate_zeros()
    num_loops = int(diag.sum())
    num_edges = int(off.nnz // 2 + num_loops)
    d = np.asarray(off.sum(axis=1)).ravel()                       # neighbours other than the node itself
    nx_degree = d + 2 * diag
    avg_degree = float(nx_degree.mean()) if n else 0.0
    tri2 = np.asarray((off @ off).multiply(off).sum(axis=1)).ravel()   # 2 * triangles through each node
    denom = d * (d - 1)
    with np.errstate(divide="ignore", invalid="ignore"):
        c = np.where((tri2 > 0) & (denom > 0), tri2 / np.where(denom > 0, denom, 1), 0.0)
    clustering = float(c.sum() / n) if n else 0.0
    num_comp, labels = connected_components(pattern, directed=False)
    sizes = np.bincount(labels, minlength=num_comp) if n else np.zeros(0, dtype=np.int64)
    largest = int(sizes.max()) if n else 0
    out = {
        "num_nodes": n, "num_edges": num_edges, "avg_degree": avg_degree, "clustering_coefficient": clustering,
        "num_connected_components": int(num_comp), "largest_component_ratio": largest / n if n else 0.0,
    }
    if with_connectivity:
        ac = 0.0
        if n > 1:
            keep = np.flatnonzero(labels == int(np.argmax(sizes)))
            if len(keep) > 1:
                a = adj[keep][:, keep].toarray().astype(np.float64)
                a = np.maximum(a, a.T)
                np.fill_diagonal(a, 0.0)
                lap = np.diag(a.sum(axis=1)) - a
                ac = float(np.sort(np.linalg.eigvalsh(lap))[1])
        out["algebraic_connectivity"] = ac
    return out


def compute_topology_preservation(original_adj, sparse_adj) -> dict:
    o, s = compute_topology_metrics(original_adj), compute_topology_metrics(sparse_adj)
    return {
        "edge_retention": s["num_edges"] / o["num_edges"] if o["num_edges"] > 0 else 0.0,
        "clustering_preservation": (s["clustering_coefficient"] / o["clustering_coefficient"]
                                    if o["clustering_coefficient"] > 0 else 1.0),
        "connectivity_preservation": (s["algebraic_connectivity"] / o["algebraic_connectivity"]
                                      if o["algebraic_connectivity"] > 0 else 0.0),
        "component_change": s["num_connected_components"] - o["num_connected_components"],
        "original_metrics": o, "sparse_metrics": s,
    }


def sample_pairs(n: int, n_samples: int, seed: int):
    """The reference's pair sampling (metrics.py:388-402): PCG64 draws until `n_samples` distinct unordered pairs."""
    rng = np.random.default_rng(seed)
    pairs = set()
    attempts = 0
    while len(pairs) < n_samples and attempts < n_samples * 10:
        u, v = rng.integers(0, n, size=2)
        if u != v:
            pairs.add((min(u, v), max(u, v)))
        attempts += 1
    return list(pairs)


def compute_geodesic_preservation(original_adj, sparse_adj, n_samples: int = 500, seed: int = 42) -> dict:
    """metrics.py:361-442 with SciPy breadth-first distances in place of nx.shortest_path_length (hop counts)."""
    from scipy.sparse.csgraph import shortest_path

    n = original_adj.shape[0]
    pairs = sample_pairs(n, n_samples, seed)
    sources = sorted({int(u) for u, _ in pairs})
    col = {u: i for i, u in enumerate(sources)}

    def hops(adj):
        pat = sp.csr_matrix(adj) != 0
        pat = (pat + pat.T).astype(np.float64)
        return shortest_path(pat, method="D", directed=False, unweighted=True, indices=sources) if sources else np.zeros((0, n))

    d_o, d_s = hops(original_adj), hops(sparse_adj)
    preserved = increased = disconnected = 0
    inc = []
    for u, v in pairs:
        a = d_o[col[int(u)], int(v)]
        if np.isinf(a):
            continue
        b = d_s[col[int(u)], int(v)]
        if np.isinf(b):
            disconnected += 1
        elif b == a:
            preserved += 1
        else:
            increased += 1
            inc.append(int(b - a))
    total = preserved + increased + disconnected
    return {
        "preservation_ratio": preserved / total if total > 0 else 0.0, "pairs_tested": len(pairs), "preserved_count": preserved,
        "increased_count": increased, "disconnected_count": disconnected,
        "avg_distance_increase": float(np.mean(inc)) if inc else 0.0, "max_distance_increase": max(inc) if inc else 0,
    }
