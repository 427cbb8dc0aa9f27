"""Global metric backbone on the GPU — drop-in for reference `src/sparsification/metric_backbone.py:28-141`.

`compute_metric_backbone(data, edge_weights, epsilon, verbose)` keeps edge (u, v) iff `w_uv <= d_G(u, v) + epsilon`
where `d_G` is the shortest-path length in the undirected graph G built from the `u < v` columns of `edge_index`
(duplicates merged with `min`, exactly like the reference's NetworkX construction). The all-pairs Dijkstra of the
reference becomes batched in-place (min,+) relaxation on the device (`gsp_sssp_batch`); only distances between
adjacent nodes are read back out of each source batch. Same return value: `(Data, stats)` with the reference's keys.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Tuple

import numpy as np
import torch

from . import _lib
from .engine import DeviceGraph, compact_edges


def _device_of(data) -> torch.device:
    _lib.require_cuda()
    if data.edge_index.is_cuda:
        return data.edge_index.device
    return torch.device("cuda", torch.cuda.current_device())


def backbone_mask(edge_index: torch.Tensor, num_nodes: int, weights: torch.Tensor, epsilon: float = 1e-9,
                  scratch_bytes: float = 16e9):
    """uint8 keep-mask over the columns of `edge_index` (device tensors in, device tensor out) + sweep count."""
    lib = _lib.load()
    dev = edge_index.device
    n, e = int(num_nodes), edge_index.size(1)
    row, col = edge_index[0], edge_index[1]
    w = weights.to(device=dev, dtype=torch.float64)
    if w.numel() < e:
        raise IndexError(f"index {e - 1} is out of bounds for axis 0 with size {w.numel()}")
    w = w[:e]
    mask = torch.zeros(e, dtype=torch.uint8, device=dev)
    if e == 0:
        return mask, 0
    # G: one undirected edge per distinct (u < v) column, weight = min over its duplicates (metric_backbone.py:70-77)
    up = row < col
    keys = row[up] * n + col[up]
    uniq, inv = torch.unique(keys, return_inverse=True)
    wmin = torch.full((uniq.numel(),), float("inf"), dtype=torch.float64, device=dev).scatter_reduce(0, inv, w[up], "amin")
    lo, hi = uniq // n, uniq % n
    dkeys = torch.cat([lo * n + hi, hi * n + lo])
    order = torch.argsort(dkeys)
    dkeys = dkeys[order]
    gw = torch.cat([wmin, wmin])[order].contiguous()               # lengths in canonical (row, col) order
    gei = torch.stack([dkeys // n, dkeys % n])
    graph = DeviceGraph(gei, n)
    # distance u -> v for every column, read out of the batch that holds source u
    batch = int(max(1, min(n, scratch_bytes // (8 * max(n, 1)))))
    inf = float("inf")
    dist_uv = torch.full((e,), inf, dtype=torch.float64, device=dev)
    same = row == col
    dist_uv[same] = 0.0                                            # dist_matrix[u][u] = 0
    rounds_total = 0
    if uniq.numel() > 0:
        rounds = C.c_int32(0)
        src_sorted, pos_sorted = torch.sort(row)
        for s0 in range(0, n, batch):
            s1 = min(n, s0 + batch)
            a, b = torch.searchsorted(src_sorted, torch.tensor([s0, s1], device=dev)).tolist()
            if a == b:
                continue
            dist = torch.empty((n, s1 - s0), dtype=torch.float64, device=dev)
            with torch.cuda.device(dev):
                _lib.check(lib.gsp_sssp_batch(graph._handle, _lib.ptr(gw), s0, s1 - s0, _lib.ptr(dist), max(n, 1),
                                              C.byref(rounds), _lib.stream_ptr(dev)))
            rounds_total = max(rounds_total, rounds.value)
            idx = pos_sorted[a:b]
            dist_uv[idx] = dist[col[idx], row[idx] - s0]
            del dist
    keep = torch.isinf(dist_uv) | (w <= dist_uv + epsilon)         # metric_backbone.py:100-108
    return keep.to(torch.uint8), rounds_total


def compute_metric_backbone(data, edge_weights, epsilon: float = 1e-9, verbose: bool = True) -> Tuple[object, Dict]:
    """reference metric_backbone.py:28-141 (same arguments, same `(sparse_data, stats)` result)."""
    dev = _device_of(data)
    ei = data.edge_index.to(dev)
    if ei.dtype != torch.int64:
        ei = ei.long()
    n, e = data.num_nodes, ei.size(1)
    w_host = np.asarray(edge_weights, dtype=np.float64)
    if verbose:
        print("Computing Global Metric Backbone")
        print(f"  Nodes: {n:,}, Edges: {e:,}")
        print(f"  Epsilon: {epsilon}")
    mask, rounds = backbone_mask(ei, n, torch.from_numpy(np.ascontiguousarray(w_host)), epsilon)
    kept = int(mask.sum().item())
    out, _, _ = compact_edges(ei.contiguous(), mask, kept)
    keep_mask = mask.cpu().numpy().astype(bool)
    sparse_data = data.clone()
    sparse_data.edge_index = out.cpu()                             # the reference returns a CPU edge_index here
    stats = {
        "original_edges": e,
        "retained_edges": kept,
        "removed_edges": e - kept,
        "retention_ratio": float(kept / e) if e else float("nan"),
        "edges_metric": kept,
        "edges_semi_metric": e - kept,
        "epsilon": epsilon,
        "sparse_weights": w_host[:e][keep_mask],
        "keep_mask": keep_mask,
        "relaxation_sweeps": rounds,
    }
    if verbose:
        print(f"  Metric (retained):     {kept:,} ({stats['retention_ratio']:.1%})")
        print(f"  Semi-metric (removed): {e - kept:,}")
    return sparse_data, stats
