"""TEST INFRASTRUCTURE — CPU restatement of the torch_geometric pieces the reference's GCN consumer runs on the kept
sub-graph. Only tests/ may import this module; the product path is gnn-sparsification-research_b200/gcn.py
+ csrc/gcn.cu.

Third-party dependency, not vendored under /root/reference and not installed in this image: torch-geometric >= 2.3.0
(reference pyproject.toml:31; no lock file). The functions restate its published algorithm:

* `gcn_norm`  — torch_geometric/nn/conv/gcn_conv.py `gcn_norm` (edge_index branch) with
  torch_geometric/utils/loop.py `add_remaining_self_loops`, called by `GCNConv(normalize=True, add_self_loops=True,
  improved=False)`, flow "source_to_target" — reference call sites src/models/gnn.py:222-223 (construction) and :244
  (`conv(x, edge_index, edge_weight=edge_weight)`).
* `propagate` — `MessagePassing.propagate` of GCNConv: message `edge_weight.view(-1, 1) * x_j`, sum aggregation at
  edge_index[1] (a scatter-add, sequential in edge order on the CPU).

PARITY UNPINNED against torch_geometric itself (it cannot be imported here); pinned only to this restatement. fp32
throughout, one rounding per operation (NumPy float32 scalars / arrays), sums in edge-list order.
"""
from __future__ import annotations

import numpy as np


def add_remaining_self_loops(edge_index: np.ndarray, edge_weight, num_nodes: int, fill_value: float = 1.0):
    row, col = edge_index[0], edge_index[1]
    mask = row != col
    loop_index = np.arange(num_nodes, dtype=np.int64)
    if edge_weight is not None:
        loop_attr = np.full(num_nodes, fill_value, dtype=np.float32)
        inv = ~mask
        loop_attr[row[inv]] = edge_weight[inv]            # duplicates: the last assignment wins
        edge_weight = np.concatenate([edge_weight[mask], loop_attr]).astype(np.float32)
    out = np.concatenate([edge_index[:, mask], np.vstack([loop_index, loop_index])], axis=1)
    return out, edge_weight


def gcn_norm(edge_index: np.ndarray, edge_weight=None, num_nodes: int | None = None):
    """(edge_index int64 [2, K], edge_weight float32 [K]) of `D^-1/2 (A + I) D^-1/2`, torch_geometric layout."""
    edge_index = np.asarray(edge_index, dtype=np.int64)
    if num_nodes is None:
        num_nodes = int(edge_index.max()) + 1 if edge_index.size else 0
    if edge_weight is None:
        edge_weight = np.ones(edge_index.shape[1], dtype=np.float32)
    edge_weight = np.asarray(edge_weight, dtype=np.float32)
    edge_index, edge_weight = add_remaining_self_loops(edge_index, edge_weight, num_nodes, 1.0)
    row, col = edge_index[0], edge_index[1]
    deg = np.zeros(num_nodes, dtype=np.float32)
    np.add.at(deg, col, edge_weight)                      # unbuffered, sequential in edge order (fp32)
    with np.errstate(divide="ignore", invalid="ignore"):
        dinv = (np.float32(1.0) / np.sqrt(deg)).astype(np.float32)
    dinv[np.isinf(dinv)] = 0.0
    return edge_index, (dinv[row] * edge_weight * dinv[col]).astype(np.float32)


def propagate(edge_index: np.ndarray, edge_weight: np.ndarray, x: np.ndarray, num_nodes: int) -> np.ndarray:
    """out[t] = sum over edges into t (edge order) of w * x[source]; fp32, multiply then add."""
    x = np.asarray(x, dtype=np.float32)
    out = np.zeros((num_nodes, x.shape[1]), dtype=np.float32)
    msg = (edge_weight.astype(np.float32)[:, None] * x[edge_index[0]]).astype(np.float32)
    np.add.at(out, edge_index[1], msg)
    return out
