"""Pins the oracles (plain-C restatement + SciPy port) before anything trusts them.

* against the committed golden fixtures (outputs of the live reference, made by
  oracle/make_golden.py) — runs everywhere, including the GPU box;
* against the live reference imported from /root/reference — build container only.
"""
import os
import sys

import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import ref_loader
from oracle import scipy_port as port
from tests.helpers import RETENTIONS, bits_equal, golden_names, load_golden


@pytest.mark.parametrize("name", golden_names())
def test_c_oracle_matches_golden(name):
    g = load_golden(name)
    ei, n = g["edge_index"], int(g["num_nodes"])
    e = ei.shape[1]
    csr = co.csr_from_edge_index(ei, n)
    assert csr.nnz == int(g["nnz"])
    assert np.array_equal(csr.indptr, g["csr_indptr"]) and np.array_equal(csr.indices, g["csr_indices"])
    assert bits_equal(csr.data, g["csr_data"])
    scores = {
        "jaccard": co.calculate_jaccard_scores(csr),
        "adamic_adar": co.calculate_adamic_adar_scores(csr, node_weights=g["aa_node_w"]),
        "degree": co.degree_product_scores(csr),
    }
    if "x" in g:
        scores["feature_cosine"] = co.calculate_feature_cosine_scores(csr, g["x"])
    for m, s in scores.items():
        assert bits_equal(s, g[f"score_{m}"]), m            # integer/fp64/fp32-tree work: bit-exact
    if "score_approx_er" in g:
        er = co.calculate_approx_effective_resistance_scores(csr, epsilon=float(g["er_epsilon"]))
        np.testing.assert_allclose(er, g["score_approx_er"], rtol=1e-4)      # north_star tolerance
        scores["approx_er"] = g["score_approx_er"]
    for m in ("jaccard", "adamic_adar"):
        if f"mask_backbone_{m}" in g:
            assert np.array_equal(co.metric_backbone_mask(ei, n, co.scores_to_cost(g[f"score_{m}"])), g[f"mask_backbone_{m}"]), m
    for m, s in scores.items():
        if m == "degree":
            continue
        for r in RETENTIONS:
            tag = f"{m}_{int(r * 100)}"
            for kl in (False, True):
                assert np.array_equal(co.threshold_mask(s, e, r, kl), g[f"mask_{'low' if kl else 'top'}_{tag}"]), tag
            if f"mask_dega_{tag}" in g:
                assert np.array_equal(co.degree_aware_mask(s, ei[0], n, e, r, 1), g[f"mask_dega_{tag}"])
                assert np.array_equal(co.degree_aware_mask(s, ei[0], n, e, r, 2), g[f"mask_dega2_{tag}"])


@pytest.mark.parametrize("name", golden_names())
def test_scipy_port_matches_golden(name):
    g = load_golden(name)
    ei, n = g["edge_index"], int(g["num_nodes"])
    e = ei.shape[1]
    adj = port.build_adjacency(ei, n)
    # Jaccard is integer counts + one IEEE divide: identical on any host. AA goes through
    # libm log (SIMD-dispatch dependent), so it is pinned with the fixture's own weights in
    # the C-oracle test and to 1e-12 here.
    assert bits_equal(port.jaccard(adj), g["score_jaccard"])
    np.testing.assert_allclose(port.adamic_adar(adj), g["score_adamic_adar"], rtol=1e-12)
    assert bits_equal(port.degree_product(adj), g["score_degree"])
    if "x" in g:
        assert bits_equal(port.feature_cosine(adj, g["x"]), g["score_feature_cosine"])
    if "score_approx_er" in g:
        er = port.approx_effective_resistance(adj, epsilon=float(g["er_epsilon"]))
        np.testing.assert_allclose(er, g["score_approx_er"], rtol=1e-6)
    s = g["score_jaccard"]
    for r in RETENTIONS:
        tag = f"jaccard_{int(r * 100)}"
        assert np.array_equal(port.threshold_mask(s, e, r), g[f"mask_top_{tag}"])
        assert np.array_equal(port.threshold_mask(s, e, r, keep_lowest=True), g[f"mask_low_{tag}"])
        if f"mask_dega_{tag}" in g:
            assert np.array_equal(port.degree_aware_mask(s, ei[0], n, e, r, 1), g[f"mask_dega_{tag}"])
            assert np.array_equal(port.degree_aware_mask(s, ei[0], n, e, r, 2), g[f"mask_dega2_{tag}"])
            assert np.array_equal(port.sampled_mask(s, e, r, 42), g[f"mask_samp_{tag}"])
            mask = g[f"mask_top_{tag}"]
            assert bits_equal(port.minmax_edge_weight(s, mask), g[f"weight_top_{tag}"])
    us, inv = port.precompute_random_scores(ei, n, 42)
    assert bits_equal(us, g["random_undirected_scores"]) and np.array_equal(inv, g["random_inverse_idx"])
    for r in RETENTIONS:
        assert np.array_equal(ei[:, port.random_mask(us, inv, r)], g[f"random_edge_index_{int(r * 100)}"])


def test_pairwise_sum_tree_matches_numpy():
    """the restated tree must equal np.sum / np.add.reduce bit-for-bit."""
    rng = np.random.default_rng(0)
    for d in list(range(1, 40)) + [64, 100, 127, 128, 129, 200, 255, 256, 257, 300, 512, 1000, 1433]:
        a32 = rng.standard_normal(d, dtype=np.float32)
        assert co.pairwise_sum(a32).tobytes() == np.add.reduce(a32).tobytes(), d
        a64 = rng.standard_normal(d)
        assert co.pairwise_sum(a64).tobytes() == np.add.reduce(a64).tobytes(), d


def test_scores_to_cost_kat():
    """The one known-answer test the reference holds for this path (tests/test_sparsification.py:143-150)."""
    from gsr_b200.core import GraphSparsifier  # host logic only: no device call
    costs = GraphSparsifier._scores_to_cost(None, np.array([0.5, 1.0, 0.25]), "jaccard")
    np.testing.assert_allclose(costs, [1.0, 0.0, 3.0])


needs_ref = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present (GPU box)")


@needs_ref
@pytest.mark.parametrize("shape", [(800, 6400, 10, 41), (2708, 10556, 12, 1)])
def test_oracles_match_live_reference(shape):
    import torch
    from gsr_b200.data import Data
    from gsr_b200.synthetic import features, rmat_graph

    n, e, scale, seed = shape
    ref = ref_loader.load(stable=True)
    ei = rmat_graph(n, e, scale, seed)
    x = features(n, 65, seed)
    sp = ref.GraphSparsifier(Data(edge_index=torch.from_numpy(ei), x=torch.from_numpy(x), num_nodes=n), "cpu")
    csr = co.csr_from_edge_index(ei, n)
    adj = port.build_adjacency(ei, n)
    for m, fn, pfn in (("jaccard", co.calculate_jaccard_scores, port.jaccard),
                       ("adamic_adar", co.calculate_adamic_adar_scores, port.adamic_adar),
                       ("degree", co.degree_product_scores, port.degree_product)):
        want = sp.compute_scores(m)
        assert bits_equal(fn(csr), want), m
        assert bits_equal(pfn(adj), want), m
    want = sp.compute_scores("feature_cosine")
    assert bits_equal(co.calculate_feature_cosine_scores(csr, x), want)
    assert bits_equal(port.feature_cosine(adj, x), want)
    s = sp.compute_scores("jaccard")
    for r in (0.8, 0.5, 0.13):
        for kl in (False, True):
            _, mask = sp.sparsify("jaccard", r, return_mask=True, keep_lowest=kl)
            assert np.array_equal(co.threshold_mask(s, e, r, kl), mask.numpy())
    if n <= 1000:
        _, mask = sp.sparsify_degree_aware("jaccard", 0.5, return_mask=True)
        assert np.array_equal(co.degree_aware_mask(s, ei[0], n, e, 0.5, 1), mask.numpy())
        metrics = sys.modules[ref.__name__ + ".metrics"]
        want = metrics.calculate_approx_effective_resistance_scores(sp.adj, epsilon=1.5)
        np.testing.assert_allclose(co.calculate_approx_effective_resistance_scores(csr, epsilon=1.5), want, rtol=1e-4)
        assert bits_equal(port.approx_effective_resistance(adj, epsilon=1.5), want)


@needs_ref
def test_reference_property_tests_hold_for_oracle():
    """The reference's own property tests (tests/test_sparsification.py:168-220) re-run on the oracle."""
    import scipy.sparse as sp

    tri = sp.csr_matrix(np.array([[0, 1, 1], [1, 0, 1], [1, 1, 0]]))
    s = co.calculate_jaccard_scores(tri)
    assert np.all(s >= 0) and np.all(s <= 1)
    iso = sp.csr_matrix(np.array([[0, 1, 0], [1, 0, 0], [0, 0, 0]]))
    assert np.all(np.isfinite(co.calculate_jaccard_scores(iso)))
    star = sp.csr_matrix(np.array([[0, 1, 1, 1], [1, 0, 0, 0], [1, 0, 0, 0], [1, 0, 0, 0]]))
    assert np.all(np.isfinite(co.calculate_adamic_adar_scores(star)))
    assert np.all(co.calculate_approx_effective_resistance_scores(tri) > 0)


@needs_ref
@pytest.mark.parametrize("shape", [(34, None), (400, 3000), (1500, 9000)])
def test_metric_backbone_oracle_matches_live_reference(shape):
    """the oracle's Dijkstra backbone against the reference's NetworkX APSP (metric_backbone.py:58-112)."""
    import torch
    from gsr_b200.data import Data
    from gsr_b200.synthetic import rmat_graph
    from tests.helpers import load_golden

    n, e = shape
    ei = load_golden("karate_unsorted")["edge_index"] if e is None else rmat_graph(n, e, 11, seed=n)
    ref = ref_loader.load(stable=True)
    sp = ref.GraphSparsifier(Data(edge_index=torch.from_numpy(ei), num_nodes=n), "cpu")
    for metric in ("jaccard", "adamic_adar"):
        s = sp.compute_scores(metric)
        cost = sp._scores_to_cost(s, metric)
        assert bits_equal(co.scores_to_cost(s), cost)
        _, stats = sp.sparsify_metric_backbone(metric)
        assert np.array_equal(co.metric_backbone_mask(ei, n, cost), stats["keep_mask"]), metric


def test_topology_port_matches_goldens_and_live_reference():
    """oracle/topology_port.py against tests/golden/topology_metrics.json (live reference, NetworkX) and, in the build
    container, against the reference function itself."""
    import json

    from oracle import topology_port
    from oracle.make_topology_golden import adjacency, graphs

    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "topology_metrics.json")))
    live = None
    if ref_loader.available():
        ref = ref_loader.load()
        live = __import__(ref.__name__ + ".metrics", fromlist=["compute_topology_metrics"]).compute_topology_metrics
    seen = 0
    for name, ei, n in graphs():
        adj = adjacency(ei, n)
        adj.data[:] = 1.0
        got = topology_port.compute_topology_metrics(adj)
        for want in ([gold[name]] + ([live(adj)] if live is not None and n <= 400 else [])):
            for k in ("num_nodes", "num_edges", "num_connected_components"):
                assert got[k] == want[k], (name, k)
            for k in ("avg_degree", "clustering_coefficient", "largest_component_ratio"):
                assert abs(got[k] - want[k]) <= 1e-12 * max(1.0, abs(want[k])), (name, k)
            assert abs(got["algebraic_connectivity"] - want["algebraic_connectivity"]) <= 1e-6 * max(1.0, want["algebraic_connectivity"]), name
        seen += 1
    assert seen == len(gold) == 10


def test_geodesic_port_matches_goldens():
    import json

    from oracle import topology_port
    from oracle.make_topology_golden import GEODESIC_CASES, adjacency, graphs, thinned

    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "geodesic_preservation.json")))
    by_name = {name: (ei, n) for name, ei, n in graphs()}
    for name, samples in GEODESIC_CASES:
        ei, n = by_name[name]
        got = topology_port.compute_geodesic_preservation(adjacency(ei, n), adjacency(thinned(ei), n), n_samples=samples, seed=42)
        for k, v in got.items():
            assert abs(v - gold[name][k]) <= 1e-12, (name, k)


def test_pyg_port_matches_dense_normalised_adjacency():
    """oracle/pyg_port.py (restatement of torch_geometric's gcn_norm + propagate; the library itself is not installed, so it
    stays "parity unpinned" against PyG) against a dense D^-1/2 (A + I') D^-1/2 built independently, with self loops that keep
    their weight, duplicate edges and isolated nodes — and against NetworkX's normalised Laplacian on the karate club."""
    from oracle import pyg_port

    rng = np.random.default_rng(0)
    n = 25
    ei = np.vstack([rng.integers(0, 20, 120), rng.integers(0, 20, 120)])          # nodes 20..24 isolated, some loops/duplicates
    w = (rng.random(120) + 0.1).astype(np.float32)
    out_ei, out_w = pyg_port.gcn_norm(ei, w, n)
    mask = ei[0] != ei[1]
    assert np.array_equal(out_ei[:, : mask.sum()], ei[:, mask])                   # non-loop edges first, in order
    assert np.array_equal(out_ei[:, mask.sum():], np.vstack([np.arange(n)] * 2))  # then one loop per node
    dense = np.zeros((n, n))
    for (r, c), v in zip(ei[:, mask].T, w[mask]):
        dense[c, r] += v                                                          # aggregation at the target
    loop = np.ones(n)
    for (r, _), v in zip(ei[:, ~mask].T, w[~mask]):
        loop[r] = v                                                               # the last loop edge of a node wins
    dense += np.diag(loop)
    deg = dense.sum(axis=1)
    want = dense / np.sqrt(deg)[:, None] / np.sqrt(deg)[None, :]
    got = np.zeros((n, n))
    for (r, c), v in zip(out_ei.T, out_w):
        got[c, r] += v
    assert np.abs(got - want).max() < 5e-7
    x = rng.standard_normal((n, 6)).astype(np.float32)
    assert np.abs(pyg_port.propagate(out_ei, out_w, x, n) - want @ x).max() < 5e-6
    # unweighted graph without loops: weights are exactly 1 / sqrt((d_u + 1)(d_v + 1)) up to fp32 rounding
    ring = np.vstack([np.arange(8), (np.arange(8) + 1) % 8])
    ring = np.hstack([ring, ring[::-1]])
    _, rw = pyg_port.gcn_norm(ring, None, 8)
    assert np.allclose(rw, 1.0 / 3.0, rtol=2e-7)
    # a second, library-made pin: NetworkX's normalised Laplacian of the graph with unit self loops added is
    # I - D^-1/2 (A + I) D^-1/2 (row sums of the adjacency matrix, so a loop counts once — the GCN convention)
    import networkx as nx

    kc = nx.karate_club_graph()
    m = kc.number_of_nodes()
    for weight in (None, "weight"):
        g = nx.Graph()
        g.add_nodes_from(range(m))
        g.add_weighted_edges_from((u, v, float(d["weight"]) if weight else 1.0) for u, v, d in kc.edges(data=True))
        pairs = np.array([(u, v) for u, v in g.edges()] + [(v, u) for u, v in g.edges()]).T
        vals = np.array([g[u][v]["weight"] for u, v in pairs.T], dtype=np.float32)
        g.add_weighted_edges_from((i, i, 1.0) for i in range(m))
        want = np.eye(m) - nx.normalized_laplacian_matrix(g, nodelist=range(m), weight="weight").toarray()
        k_ei, k_w = pyg_port.gcn_norm(pairs, vals if weight else None, m)
        got = np.zeros((m, m))
        for (r, c), v in zip(k_ei.T, k_w):
            got[c, r] += v
        assert np.abs(got - want).max() < 5e-7, weight
