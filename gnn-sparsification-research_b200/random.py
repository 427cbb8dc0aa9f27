"""Symmetric random baseline — drop-in for reference `src/sparsification/random.py:14-52`.

One PCG64 score per undirected edge, nested across retention rates. The scores themselves are NumPy's
random stream (that *is* the reference behaviour); grouping the two directions of an edge, the top-k
over undirected edges and the edge_index compaction run on the GPU.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .engine import compact_edges, select_mask


def _cuda_of(data, device=None) -> torch.device:
    _lib.require_cuda()
    if data.edge_index.is_cuda:
        return data.edge_index.device
    if device is not None and torch.device(device).type == "cuda":
        d = torch.device(device)
        return d if d.index is not None else torch.device("cuda", torch.cuda.current_device())
    return torch.device("cuda", torch.cuda.current_device())


def precompute_random_scores(data, seed: int = 42):
    """(undirected_scores float64[n_undirected], inverse_idx int64[E]) — reference random.py:14-33."""
    dev = _cuda_of(data)
    ei = data.edge_index.to(dev)
    n = data.num_nodes
    lo, hi = torch.minimum(ei[0], ei[1]), torch.maximum(ei[0], ei[1])
    keys = lo.long() * (n + 1) + hi.long()
    _, inverse = torch.unique(keys, return_inverse=True)       # sorted unique, like np.unique
    inverse_idx = inverse.cpu().numpy().astype(np.int64)
    n_undirected = int(inverse_idx.max()) + 1
    rng = np.random.default_rng(seed)
    return rng.random(n_undirected), inverse_idx


def random_sparsify(data, undirected_scores, inverse_idx, retention_ratio: float, device: str):
    """Keep the top `max(1, int(n_undirected * r))` undirected edges, both directions (reference random.py:36-52)."""
    if retention_ratio == 1.0:
        return data.clone()
    dev = _cuda_of(data, device)
    n_undirected = len(undirected_scores)
    n_keep = max(1, int(n_undirected * retention_ratio))
    scores = torch.from_numpy(np.ascontiguousarray(undirected_scores, dtype=np.float64)).to(dev)
    keep_undir = select_mask(scores, min(n_keep, n_undirected), keep_lowest=False)
    inv = torch.from_numpy(np.ascontiguousarray(inverse_idx)).to(dev)
    mask = keep_undir[inv].contiguous()
    kept = int(mask.sum().item())
    ei = data.edge_index.to(dev)
    out, _, _ = compact_edges(ei if ei.dtype == torch.int64 else ei.long(), mask, kept)
    sparse = data.clone()
    sparse.edge_index = out.to(device)
    return sparse
