"""Seeded synthetic graphs of the shapes named in BASELINE.json.

The reference ships no data (its `src/data` loader is absent and there is no
network), so every test and benchmark runs on R-MAT power-law graphs folded to
the named node count: undirected, symmetrised, self-loop-free, de-duplicated,
`edge_index` int64 sorted by (row, col) — the form PyG's coalesced datasets
have, for which CSR position == `edge_index` column.

`rmat_graph` is the host/NumPy generator (tests, fixtures, CPU-baseline
samples); `rmat_graph_device` builds the same family with torch ops on the GPU
for the 10^8-edge benchmark shapes (different random stream, same distribution).
"""
from __future__ import annotations

import numpy as np
import torch

RMAT_ABCD = (0.57, 0.19, 0.19, 0.05)

# name -> (num_nodes, directed_edges, feature_dim, rmat_scale, seed)
SHAPES = {
    "cora": (2708, 10556, 1433, 12, 1),
    "roman_empire": (22662, 65854, 300, 15, 2),
    "arxiv": (169343, 2315598, 128, 18, 3),
    "products": (2449029, 123718280, 100, 22, 4),
    "rmat24": (1 << 24, 268435456, 128, 24, 5),
}


def _rmat_pairs_numpy(rng, scale, count, abcd):
    a, b, c, _ = abcd
    src = np.zeros(count, dtype=np.int64)
    dst = np.zeros(count, dtype=np.int64)
    for _ in range(scale):
        r = rng.random(count)
        right = ((r >= a) & (r < a + b)) | (r >= a + b + c)      # quadrants b, d: column bit set
        down = r >= a + b                                         # quadrants c, d: row bit set
        src = (src << 1) | down
        dst = (dst << 1) | right
    return src, dst


def rmat_graph(num_nodes: int, num_directed_edges: int, scale: int | None = None, seed: int = 0,
               abcd=RMAT_ABCD) -> np.ndarray:
    """Symmetric R-MAT graph with exactly `num_directed_edges` (even) directed edges; int64 [2, E] sorted."""
    if num_directed_edges % 2:
        raise ValueError("directed edge count must be even (every undirected edge is stored twice)")
    want = num_directed_edges // 2
    if want > num_nodes * (num_nodes - 1) // 2:
        raise ValueError("too many edges for the node count")
    if scale is None:
        scale = max(1, int(np.ceil(np.log2(max(num_nodes, 2)))))
    rng = np.random.default_rng(seed)
    keys = np.zeros(0, dtype=np.int64)
    while len(keys) < want:
        batch = int((want - len(keys)) * 1.3) + 1024
        s, d = _rmat_pairs_numpy(rng, scale, batch, abcd)
        s %= num_nodes
        d %= num_nodes
        lo, hi = np.minimum(s, d), np.maximum(s, d)
        k = (lo * num_nodes + hi)[lo != hi]
        # keep first occurrences in generation order so the result does not depend on batch size
        merged = np.concatenate([keys, k])
        _, first = np.unique(merged, return_index=True)
        keys = merged[np.sort(first)]
    keys = keys[:want]
    lo, hi = keys // num_nodes, keys % num_nodes
    row = np.concatenate([lo, hi])
    col = np.concatenate([hi, lo])
    order = np.lexsort((col, row))
    return np.vstack([row[order], col[order]])


def rmat_graph_device(num_nodes: int, num_directed_edges: int, scale: int, seed: int, device="cuda",
                      abcd=RMAT_ABCD) -> torch.Tensor:
    """GPU generator for the large benchmark shapes; returns int64 [2, E] sorted by (row, col) on `device`."""
    want = num_directed_edges // 2
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    a, b, c, _ = abcd
    keys = torch.zeros(0, dtype=torch.int64, device=device)
    while keys.numel() < want:
        batch = int((want - keys.numel()) * 1.25) + 4096
        src = torch.zeros(batch, dtype=torch.int64, device=device)
        dst = torch.zeros(batch, dtype=torch.int64, device=device)
        for _ in range(scale):
            r = torch.rand(batch, device=device, generator=gen)
            right = ((r >= a) & (r < a + b)) | (r >= a + b + c)
            down = r >= a + b
            src = (src << 1) | down
            dst = (dst << 1) | right
            del r, right, down
        src %= num_nodes
        dst %= num_nodes
        lo, hi = torch.minimum(src, dst), torch.maximum(src, dst)
        k = (lo * num_nodes + hi)[lo != hi]
        del src, dst, lo, hi
        keys = torch.unique(torch.cat([keys, k]))          # sorted unique (order-free: any subset is fine)
        del k
    if keys.numel() > want:                                 # drop a seeded random subset of the surplus
        perm = torch.randperm(keys.numel(), device=device, generator=gen)[:want]
        keys = keys[perm]
        del perm
    lo, hi = keys // num_nodes, keys % num_nodes
    del keys
    dkeys = torch.cat([lo * num_nodes + hi, hi * num_nodes + lo])
    del lo, hi
    dkeys, _ = torch.sort(dkeys)
    return torch.stack([dkeys // num_nodes, dkeys % num_nodes])


def features(num_nodes: int, dim: int, seed: int, kind: str = "normal") -> np.ndarray:
    """fp32 node features: N(0,1) or a Cora-like row-normalised Bernoulli bag of words."""
    rng = np.random.default_rng(seed + 1000)
    if kind == "bow":
        x = (rng.random((num_nodes, dim)) < 0.0127).astype(np.float32)
        s = x.sum(axis=1, keepdims=True)
        return (x / np.maximum(s, 1.0)).astype(np.float32)
    return rng.standard_normal((num_nodes, dim), dtype=np.float32)


def named_graph(name: str, with_features: bool = True):
    """(edge_index int64 [2,E], x fp32 [N,d] or None, num_nodes) for a BASELINE.json shape (host generator)."""
    n, e, d, scale, seed = SHAPES[name]
    ei = rmat_graph(n, e, scale, seed)
    x = features(n, d, seed, "bow" if name == "cora" else "normal") if with_features else None
    return ei, x, n


def chain_with_shortcuts(num_nodes: int, num_shortcuts: int, seed: int = 0) -> np.ndarray:
    """Path graph + random chords: Roman-empire-like, drives CG to its iteration cap."""
    rng = np.random.default_rng(seed)
    a = np.arange(num_nodes - 1, dtype=np.int64)
    lo, hi = [a], [a + 1]
    if num_shortcuts:
        s = rng.integers(0, num_nodes, size=(2, num_shortcuts))
        l2, h2 = np.minimum(s[0], s[1]), np.maximum(s[0], s[1])
        keep = h2 - l2 > 1
        lo.append(l2[keep]); hi.append(h2[keep])
    lo, hi = np.concatenate(lo), np.concatenate(hi)
    keys = np.unique(lo * num_nodes + hi)
    lo, hi = keys // num_nodes, keys % num_nodes
    row, col = np.concatenate([lo, hi]), np.concatenate([hi, lo])
    order = np.lexsort((col, row))
    return np.vstack([row[order], col[order]])
