"""B200-native edge-scoring sparsification engine — drop-in for the reference's `src.sparsification`
(reference src/sparsification/__init__.py:16-28): same class, function and method names; the work runs in
hand-written sm_100a CUDA (libgsp.so) behind the C ABI of include/gsp.h."""
from .core import GraphSparsifier
from .data import Data
from .metrics import (
    calculate_adamic_adar_scores,
    calculate_approx_effective_resistance_scores,
    calculate_effective_resistance_scores,
    calculate_feature_cosine_scores,
    calculate_jaccard_scores,
)
from .metric_backbone import compute_metric_backbone
from .topology import compute_geodesic_preservation, compute_topology_metrics, compute_topology_preservation
from .random import precompute_random_scores, random_sparsify

__all__ = [
    "GraphSparsifier",
    "Data",
    "calculate_jaccard_scores",
    "calculate_adamic_adar_scores",
    "calculate_effective_resistance_scores",
    "calculate_approx_effective_resistance_scores",
    "calculate_feature_cosine_scores",
    "compute_geodesic_preservation",
    "compute_topology_metrics",
    "compute_topology_preservation",
    "compute_metric_backbone",
    "precompute_random_scores",
    "random_sparsify",
]
