"""TEST INFRASTRUCTURE — generate `tests/golden/*.npz` from the LIVE reference.

Run in the build container (needs /root/reference):   python oracle/make_golden.py

The reference's own tests hold no golden score vectors (property tests
plus one `_scores_to_cost` KAT), so parity is pinned on outputs of the unmodified
reference code, imported through `oracle/ref_loader.py` with the stable-argsort
contract. Each fixture stores the inputs (edge_index, features, the NumPy-computed
Adamic-Adar node weights) and what the reference returned for them: score vectors
of every metric, threshold / inverse / degree-aware / sampled keep-masks, the
`-W` edge weights and the symmetric random baseline. Files are kept small enough
to commit (a few hundred KB in total).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from gsr_b200.data import Data  # noqa: E402
from gsr_b200.synthetic import chain_with_shortcuts, features, rmat_graph  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
RETENTIONS = (0.9, 0.5, 0.2)


def karate_unsorted():
    import networkx as nx

    g = nx.karate_club_graph()
    el = list(g.edges())
    # same construction as reference tests/test_sparsification.py:36-41 (NOT (row,col)-sorted)
    ei = np.array([[u, v] for u, v in el] + [[v, u] for u, v in el], dtype=np.int64).T.copy()
    return ei, g.number_of_nodes()


def cases():
    ei, n = karate_unsorted()
    yield "karate_unsorted", ei, n, features(n, 16, 21), 1.2
    yield "triangle", np.array([[0, 0, 1, 1, 2, 2], [1, 2, 0, 2, 0, 1]], dtype=np.int64), 3, features(3, 5, 22), 1.0
    star = np.array([[0, 0, 0, 1, 2, 3], [1, 2, 3, 0, 0, 0]], dtype=np.int64)
    yield "star_isolated", star, 5, features(5, 9, 23), 1.0              # node 4 isolated
    two_tri = np.array([[0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5], [1, 2, 0, 2, 0, 1, 4, 5, 3, 5, 3, 4]], dtype=np.int64)
    yield "two_triangles", two_tri, 6, features(6, 130, 24), 1.0
    yield "rmat_300", rmat_graph(300, 2400, 9, seed=31), 300, features(300, 37, 25), 1.4
    yield "rmat_1500_bow", rmat_graph(1500, 9000, 11, seed=32), 1500, features(1500, 300, 26, "bow"), 2.0
    ei = chain_with_shortcuts(400, 30, seed=33)
    yield "chain_400", ei, 400, features(400, 128, 27), 1.6
    dup = np.concatenate([karate_unsorted()[0], karate_unsorted()[0][:, :9]], axis=1)
    # duplicating one direction makes the value matrix asymmetric: CG diverges chaotically there, no ER fixture
    yield "karate_duplicates", dup, 34, None, None
    asym = rmat_graph(120, 900, 7, seed=34)[:, ::2].copy()
    yield "asymmetric_120", asym, 120, features(120, 8, 28), None


def main() -> None:
    ref = ref_loader.load(stable=True)
    metrics = sys.modules[ref.__name__ + ".metrics"]
    os.makedirs(OUT, exist_ok=True)
    for name, ei, n, x, er_eps in cases():
        data = Data(edge_index=torch.from_numpy(ei), x=None if x is None else torch.from_numpy(x), num_nodes=n)
        sp = ref.GraphSparsifier(data, "cpu")
        e = ei.shape[1]
        out = {"edge_index": ei, "num_nodes": np.int64(n), "nnz": np.int64(sp.adj.nnz),
               "csr_indptr": sp.adj.indptr.astype(np.int64), "csr_indices": sp.adj.indices.astype(np.int32),
               "csr_data": sp.adj.data.astype(np.float64)}
        if x is not None:
            out["x"] = x
        deg = np.asarray((sp.adj > 0).sum(axis=1)).ravel().astype(np.float64)
        out["aa_node_w"] = 1.0 / np.sqrt(np.maximum(np.log(deg + 1), 1e-10))     # metrics.py:104-108
        names = ["jaccard", "adamic_adar", "degree"] + (["feature_cosine"] if x is not None else [])
        for m in names:
            out[f"score_{m}"] = sp.compute_scores(m)
        if er_eps is not None:
            out["er_epsilon"] = np.float64(er_eps)
            out["score_approx_er"] = metrics.calculate_approx_effective_resistance_scores(sp.adj, epsilon=er_eps)
            sp._score_cache["approx_effective_resistance"] = out["score_approx_er"]
            names.append("approx_er")
        for m in names:
            if m == "degree":
                continue
            for r in RETENTIONS:
                tag = f"{m}_{int(r * 100)}"
                for kl in (False, True):
                    _, mask = sp.sparsify(m, r, return_mask=True, keep_lowest=kl)
                    out[f"mask_{'low' if kl else 'top'}_{tag}"] = mask.numpy()
                    s = out[f"score_{m}"][mask.numpy()[: len(out[f'score_{m}'])]] if sp.adj.nnz == e else None
                    if s is not None and len(s):
                        w = (s - s.min()) / (s.max() - s.min() + 1e-8)       # roman_empire_gpu.py:248-254
                        out[f"weight_{'low' if kl else 'top'}_{tag}"] = (1.0 - w if kl else w).astype(np.float32)
                if sp.adj.nnz == e and m in ("jaccard", "adamic_adar"):
                    _, mask = sp.sparsify_degree_aware(m, r, return_mask=True)
                    out[f"mask_dega_{tag}"] = mask.numpy()
                    _, mask = sp.sparsify_degree_aware(m, r, min_edges_per_node=2, return_mask=True)
                    out[f"mask_dega2_{tag}"] = mask.numpy()
                    _, mask = sp.sparsify_sampled(m, r, seed=42, return_mask=True)
                    out[f"mask_samp_{tag}"] = mask.numpy()
        if sp.adj.nnz == e and n <= 2000:        # global metric backbone (NetworkX APSP), metric_backbone.py:58-112
            for m in ("jaccard", "adamic_adar"):
                _, stats = sp.sparsify_metric_backbone(m)
                out[f"mask_backbone_{m}"] = stats["keep_mask"]
        us, inv = ref.precompute_random_scores(data, seed=42)
        out["random_undirected_scores"], out["random_inverse_idx"] = us, inv.astype(np.int64)
        for r in RETENTIONS:
            kept = ref.random_sparsify(data, us, inv, r, "cpu").edge_index.numpy()
            out[f"random_edge_index_{int(r * 100)}"] = kept
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **out)
        print(f"{name}: n={n} E={e} nnz={sp.adj.nnz} keys={len(out)}")


if __name__ == "__main__":
    main()
