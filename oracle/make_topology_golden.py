"""TEST INFRASTRUCTURE — tests/golden/topology_metrics.json and geodesic_preservation.json from the LIVE reference
(`compute_topology_metrics`, `compute_geodesic_preservation`, reference src/sparsification/metrics.py:361-520, NetworkX) for the symmetric golden fixture graphs and three generated
ones. Run in the build container:   python oracle/make_topology_golden.py"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from gsr_b200.synthetic import chain_with_shortcuts, rmat_graph  # noqa: E402

FIXTURES = ("karate_unsorted", "triangle", "star_isolated", "two_triangles", "rmat_300", "chain_400")


def graphs():
    for name in FIXTURES:
        g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        yield name, g["edge_index"], int(g["num_nodes"])
    yield "rmat_2000", rmat_graph(2000, 16000, 11, seed=7), 2000
    yield "chain_shortcuts_1500", chain_with_shortcuts(1500, 40, seed=3), 1500
    rng = np.random.default_rng(2)                         # several components, self loops, isolated nodes
    a = rng.integers(0, 60, 90)
    b = (a // 10) * 10 + rng.integers(0, 10, 90)
    loops = np.array([3, 3, 17, 41])
    row = np.concatenate([a, b, loops])
    col = np.concatenate([b, a, loops])
    yield "blocks_with_loops", np.vstack([row, col]), 70
    # a sparsified graph as sparsify_sampled / sparsify_degree_aware leave it: (u, v) kept without (v, u); the reference's
    # nx.from_scipy_sparse_array takes either direction as the undirected edge (metrics.py:461-462)
    g = np.load(os.path.join(ROOT, "tests", "golden", "rmat_300.npz"))
    ei = g["edge_index"]
    keep = (ei[0] * 31 + ei[1] * 17) % 5 != 0
    yield "asymmetric_kept", ei[:, keep], int(g["num_nodes"])


def adjacency(ei, n):
    return sp.csr_matrix((np.ones(ei.shape[1]), (ei[0], ei[1])), shape=(n, n))


GEODESIC_CASES = (("karate_unsorted", 200), ("rmat_300", 200), ("chain_400", 150), ("rmat_2000", 300), ("blocks_with_loops", 120))


def thinned(ei):
    """Deterministic sub-graph: drops every undirected edge whose endpoints hash to 0 mod 3 (keeps symmetry)."""
    lo, hi = np.minimum(ei[0], ei[1]), np.maximum(ei[0], ei[1])
    return ei[:, (lo * 7 + hi * 13) % 3 != 0]


def main():
    ref = ref_loader.load()
    metrics = __import__(ref.__name__ + ".metrics", fromlist=["compute_topology_metrics"])
    geo = {}
    by_name = {name: (ei, n) for name, ei, n in graphs()}
    for name, samples in GEODESIC_CASES:
        ei, n = by_name[name]
        m = metrics.compute_geodesic_preservation(adjacency(ei, n), adjacency(thinned(ei), n), n_samples=samples, seed=42)
        geo[name] = {k: (float(v) if isinstance(v, (float, np.floating)) else int(v)) for k, v in m.items()}
        geo[name]["n_samples"] = samples
        print("geodesic", name, geo[name])
    with open(os.path.join(ROOT, "tests", "golden", "geodesic_preservation.json"), "w") as f:
        json.dump(geo, f, indent=1, sort_keys=True)
    out = {}
    for name, ei, n in graphs():
        adj = adjacency(ei, n)
        adj.data[:] = 1.0
        m = metrics.compute_topology_metrics(adj)
        out[name] = {k: (float(v) if isinstance(v, (float, np.floating)) else int(v)) for k, v in m.items()}
        print(name, out[name])
    with open(os.path.join(ROOT, "tests", "golden", "topology_metrics.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
