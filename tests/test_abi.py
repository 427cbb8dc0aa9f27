"""CPU-side checks: the C-ABI library loads and exports every symbol include/gsp.h declares; host logic
(argument validation, name handling, label table) behaves like the reference without touching a device."""
import os
import re

import numpy as np
import pytest
import torch

import gsr_b200
from gsr_b200 import _lib, labels

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "gsp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gsp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import gsr_b200.build as b

    b.build_library()
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"libgsp.so does not export {name}"
    assert set(declared) == set(_lib.PROTOTYPES), "ctypes prototypes out of sync with include/gsp.h"
    assert lib.gsp_version() == 100


def test_header_constants_match_python():
    text = open(os.path.join(ROOT, "include", "gsp.h")).read()
    assert int(re.search(r"#define GSP_SELECT_BINS (\d+)", text).group(1)) == _lib.SELECT_BINS
    assert int(re.search(r"#define GSP_SELECT_PASSES (\d+)", text).group(1)) == _lib.SELECT_PASSES
    assert int(re.search(r"#define GSP_SELECT_STATE_BYTES (\d+)", text).group(1)) == _lib.SELECT_STATE_BYTES


def _tiny():
    ei = torch.tensor([[0, 0, 1, 1, 2, 2], [1, 2, 0, 2, 0, 1]], dtype=torch.long)
    return gsr_b200.Data(edge_index=ei, num_nodes=3)


def test_reference_attributes_and_validation_without_device():
    sp = gsr_b200.GraphSparsifier(_tiny(), "cpu")
    assert (sp.num_nodes, sp.num_edges, sp.verbose, sp._score_cache) == (3, 6, False, {})
    assert sp.stats == {"num_nodes": 3, "num_edges": 6, "density": 1.0, "avg_degree": 2.0}
    for bad in (0, -0.1, 1.01):
        for fn in (sp.sparsify, sp.sparsify_sampled, sp.sparsify_degree_aware):
            with pytest.raises(ValueError, match="retention_ratio must be in"):
                fn("jaccard", bad)
    with pytest.raises(ValueError, match="not supported"):
        sp.compute_scores("pagerank")
    # r == 1.0 short-circuits before any scoring (reference core.py:224-227)
    out, mask = sp.sparsify("jaccard", 1.0, return_mask=True)
    assert out.edge_index.equal(sp.data.edge_index) and out is not sp.data and mask.all() and mask.dtype == torch.bool
    for alias, want in (("AA", "adamic_adar"), ("adamic-adar", "adamic_adar"), ("approx_er", "approx_effective_resistance"),
                        ("ER", "effective_resistance"), ("rand", "random"), ("Feature Cosine", "feature_cosine")):
        assert sp._normalize_metric_name(alias) == want


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the loud failure on a GPU-less host")
def test_no_cpu_fallback():
    sp = gsr_b200.GraphSparsifier(_tiny(), "cpu")
    with pytest.raises(_lib.GspError, match="no CPU fallback"):
        sp.compute_scores("jaccard")
    import scipy.sparse as sparse
    with pytest.raises(_lib.GspError, match="no CPU fallback"):
        gsr_b200.calculate_jaccard_scores(sparse.csr_matrix(np.array([[0, 1], [1, 0]])))
    with pytest.raises(_lib.GspError, match="no CPU fallback"):
        sp.prefetch_scores(["jaccard", "adamic_adar"])
    from gsr_b200 import gcn, topology
    with pytest.raises(_lib.GspError, match="no CPU fallback"):
        gcn.gcn_norm(torch.tensor([[0, 1], [1, 0]]), None, 2)
    with pytest.raises(_lib.GspError, match="no CPU fallback"):
        gcn.GcnPropagation(torch.tensor([[0, 1], [1, 0]]), None, 2)
    with pytest.raises(_lib.GspError, match="no CPU fallback"):
        topology.compute_topology_metrics(sparse.csr_matrix(np.array([[0, 1], [1, 0]])))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gnn-sparsification-research_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle", ""), f"{f} mentions the oracle"
                assert "/root/reference" not in src


def test_label_table_and_composite_names():
    cfg = labels.SPARSIFICATION_CONFIGS
    assert cfg["Jaccard-T"] == ("jaccard", False, False, "threshold")
    assert cfg["AA-IT"] == ("adamic_adar", True, False, "threshold")
    assert cfg["ApproxER-IT-W"] == ("approx_er", True, True, "threshold")
    assert cfg["FeatCos-T-W"] == ("feature_cosine", False, True, "threshold")
    assert cfg["Jaccard-Samp"][3] == "sampled" and cfg["Jaccard-DegA"][3] == "degree_aware"
    assert cfg["Random"] == ("random", False, False, "threshold")
    assert labels.parse_composite_metric("degree_aware_adamic_adar") == ("degree_aware", "adamic_adar")
    assert labels.parse_composite_metric("sampled_jaccard") == ("sampled", "jaccard")
    assert labels.parse_composite_metric("approx_er_inv") == ("inverse", "approx_er")
    assert labels.parse_composite_metric("metric_backbone_jaccard") == ("metric_backbone", "jaccard")
    assert labels.parse_composite_metric("feature_cosine") == ("threshold", "feature_cosine")


def test_data_container_clone_is_deep():
    d = _tiny()
    c = d.clone()
    c.edge_index[0, 0] = 7
    assert d.edge_index[0, 0] == 0 and c.num_nodes == 3


def test_bench_reference_arm_runs_on_cpu():
    """`bench.py --impl reference` (the CPU arm the driver times) prints one well-formed JSON line without a GPU."""
    import json
    import subprocess
    import sys

    env = dict(os.environ, GSP_BENCH_CPU_SCALE="9")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "edges/s" and line["value"] > 0
    # the unmodified reference where /root/reference exists (build container), the oracle's port elsewhere (GPU box)
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["cpu_baseline"]["cores"] == 1 and line["cpu_baseline"]["host_cores_available"] >= 1
    assert line["higher_is_better"] is True and line["metric"].startswith("edges scored/sec")


def test_profiles_traffic_record_is_well_formed():
    import json

    t = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
    assert t["workload"] == {"scale": 24, "edge_factor": 16, "dim": 128, "n_gpus": 1}
    for k in ("jaccard", "adamic_adar"):
        assert 1e10 < t[k]["dram_bytes_per_launch"] < 1e12


def test_data_container_semantics_on_the_host():
    """`Data.to` / `clone` keep the attribute order and never leak the upload bookkeeping (host-only check; the asynchronous
    CUDA path is covered by the GPU tests)."""
    d = gsr_b200.Data(x=torch.randn(4, 3), edge_index=torch.tensor([[0, 1], [1, 0]]), num_nodes=4, note="kept")
    moved = d.to("cpu", non_blocking=True)
    assert list(moved.__dict__) == list(d.__dict__) and moved.note == "kept" and moved.num_nodes == 4
    assert not any(k.startswith("_gsp_") for k in moved.__dict__)                 # events only for uploads to a CUDA device
    moved._gsp_edge_index_ready = object()
    assert not hasattr(moved.clone(), "_gsp_edge_index_ready") and "gsp" not in repr(moved)
    assert torch.equal(moved.clone().edge_index, d.edge_index) and moved.clone().edge_index is not moved.edge_index
