"""TEST INFRASTRUCTURE — tests/golden/exact_er.npz from the LIVE reference (`calculate_effective_resistance_scores`,
reference src/sparsification/metrics.py:124-175: dense pinv of L + 1e-10 I) for the golden fixture graphs, built the way
`GraphSparsifier.__init__` builds its adjacency (core.py:70-74: duplicates summed into weights). Also stores, per graph,
how far the reference is from the regularisation-free pseudo-inverse (its own cancellation noise: the 1e10-sized null-space
term of the regularised inverse cancels in R only up to rounding), which is the tolerance the GPU test uses.
Run in the build container:   python oracle/make_exact_er_golden.py"""
from __future__ import annotations

import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402

FIXTURES = ("karate_unsorted", "triangle", "star_isolated", "two_triangles", "rmat_300", "chain_400")


def adjacency(ei, n):
    return sp.csr_matrix((np.ones(ei.shape[1]), (ei[0], ei[1])), shape=(n, n))


def clean_pinv_resistance(adj):
    """R from the pseudo-inverse of the UNregularised Laplacian (null space removed by the SVD cut-off): the value the
    reference approximates."""
    lap = (sp.diags(np.asarray(adj.sum(axis=1)).ravel()) - adj).toarray()
    pinv = np.linalg.pinv(lap, rcond=1e-12, hermitian=True)
    r, c = adj.nonzero()
    return np.maximum(pinv[r, r] + pinv[c, c] - 2.0 * pinv[r, c], 1e-10)


def main():
    ref = ref_loader.load()
    metrics = __import__(ref.__name__ + ".metrics", fromlist=["calculate_effective_resistance_scores"])
    out = {}
    for name in FIXTURES:
        g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        adj = adjacency(g["edge_index"], int(g["num_nodes"]))
        want = metrics.calculate_effective_resistance_scores(adj)
        clean = clean_pinv_resistance(adj)
        noise = float(np.max(np.abs(want - clean) / np.abs(clean)))
        out[name] = want
        out[name + "__clean"] = clean
        out[name + "__reference_noise"] = np.float64(noise)
        print(f"{name}: {len(want)} edges, R in [{want.min():.4g}, {want.max():.4g}], reference vs clean pinv: {noise:.2e}")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "exact_er.npz"), **out)


if __name__ == "__main__":
    main()
