"""Method labels and composite metric names used by the reference's experiment drivers.

`SPARSIFICATION_CONFIGS` restates the label table of reference
`scripts/nb05_roman_empire/roman_empire_gpu.py:81-102` (label -> metric, keep_lowest, weighted, variant) and
`parse_composite_metric` the name grammar of `src/hpo/config.py:22-35` /
`scripts/nb07_synthetic/run_synthetic_hpo.py:239-248`, so a caller can ask for "Jaccard-IT-W" or
"degree_aware_adamic_adar" and receive exactly the `edge_index` / `edge_weight` the GCN* trainer is fed.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

_METRICS = {"Jaccard": "jaccard", "AA": "adamic_adar", "ApproxER": "approx_er", "FeatCos": "feature_cosine"}

# label -> (metric, keep_lowest, weighted, variant)
SPARSIFICATION_CONFIGS = {"Random": ("random", False, False, "threshold")}
for _name, _metric in _METRICS.items():
    SPARSIFICATION_CONFIGS[f"{_name}-T"] = (_metric, False, False, "threshold")
    SPARSIFICATION_CONFIGS[f"{_name}-IT"] = (_metric, True, False, "threshold")
    SPARSIFICATION_CONFIGS[f"{_name}-T-W"] = (_metric, False, True, "threshold")
    SPARSIFICATION_CONFIGS[f"{_name}-IT-W"] = (_metric, True, True, "threshold")
    SPARSIFICATION_CONFIGS[f"{_name}-Samp"] = (_metric, False, False, "sampled")
    SPARSIFICATION_CONFIGS[f"{_name}-DegA"] = (_metric, False, False, "degree_aware")

SPARSIFIER_METRICS = ["jaccard", "adamic_adar", "approx_er", "feature_cosine"]
INVERSE_METRICS = [m + "_inv" for m in SPARSIFIER_METRICS]
DEGREE_AWARE_METRICS = ["degree_aware_" + m for m in ("jaccard", "adamic_adar", "approx_er")]
SAMPLED_METRICS = ["sampled_" + m for m in ("jaccard", "adamic_adar", "approx_er")]
BACKBONE_METRICS = ["metric_backbone_" + m for m in ("jaccard", "adamic_adar", "approx_er")]


def parse_composite_metric(metric: str) -> Tuple[str, str]:
    """(strategy, base_metric) for names like "degree_aware_jaccard", "sampled_aa", "jaccard_inv"."""
    for prefix in ("metric_backbone_", "degree_aware_", "sampled_"):
        if metric.startswith(prefix):
            return prefix.rstrip("_"), metric[len(prefix):]
    if metric.endswith("_inv"):
        return "inverse", metric[:-4]
    if metric == "random":
        return "random", "random"
    return "threshold", metric


def sparsify_by_label(sparsifier, label: str, retention_ratio: float, seed: int = 42):
    """(Data, edge_weight or None, mask) for one of the labelled methods, as roman_empire_gpu.py:228-256 builds them."""
    metric, keep_lowest, weighted, variant = SPARSIFICATION_CONFIGS[label]
    weight: Optional[torch.Tensor] = None
    if variant == "sampled":
        data, mask = sparsifier.sparsify_sampled(metric, retention_ratio, seed=seed, return_mask=True)
    elif variant == "degree_aware":
        data, mask = sparsifier.sparsify_degree_aware(metric, retention_ratio, return_mask=True)
    elif weighted and metric != "random":
        data, weight, mask = sparsifier.sparsify_with_weights(metric, retention_ratio, keep_lowest=keep_lowest)
    else:
        data, mask = sparsifier.sparsify(metric, retention_ratio, return_mask=True, keep_lowest=keep_lowest)
    return data, weight, mask


def sparsify_by_composite(sparsifier, metric: str, retention_ratio: float):
    """Dispatch of run_synthetic_hpo.py:281-299 for composite metric names (random excluded: see random.py)."""
    strategy, base = parse_composite_metric(metric)
    if strategy == "threshold":
        return sparsifier.sparsify(base, retention_ratio)
    if strategy == "inverse":
        return sparsifier.sparsify(base, retention_ratio, keep_lowest=True)
    if strategy == "degree_aware":
        return sparsifier.sparsify_degree_aware(base, retention_ratio)
    if strategy == "sampled":
        return sparsifier.sparsify_sampled(base, retention_ratio)
    if strategy == "metric_backbone":
        return sparsifier.sparsify_metric_backbone(base)[0]
    raise ValueError(f"Unknown strategy: {strategy}")


def compute_edge_weights(data, metric: str = "jaccard", device: str = "cuda"):
    """Second "-W" convention of the reference (`src/experiments/ablation.py:119-145`): re-score the (already
    sparsified) graph `data`, min-max normalise over ALL its edges, clip to [0.1, 1], float32 on `device`.
    Scores come from the GPU engine; the normalisation runs on the device in fp64 like the NumPy original."""
    from .core import GraphSparsifier

    sp = GraphSparsifier(data, device)
    scores = sp._device_scores(metric)
    if scores.numel() == 0:
        return torch.empty(0, dtype=torch.float32, device=device)
    lo, hi = scores.min(), scores.max()
    if bool(hi > lo):
        normalized = (scores - lo) / (hi - lo)
    else:
        normalized = torch.ones_like(scores)
    return normalized.clamp(0.1, 1.0).to(torch.float32).to(device)
