"""GCN normalisation and propagation of the kept sub-graph on the GPU (the consumer of the sparsifier).

The reference's GCN / GCN* (`src/models/gnn.py:222-223,244`) call `GCNConv(..., cached=False, normalize=True)` with the
optional "-W" edge weights, so torch_geometric's `gcn_norm` and the scatter-add propagate run again in every layer of
every forward pass over the same kept edges. `gcn_norm` below has torch_geometric's signature and return value
(`torch_geometric.nn.conv.gcn_conv.gcn_norm`, edge_index form) so its output can be fed to `GCNConv(normalize=False)`;
`GcnPropagation` is the normalised adjacency as an operator (`x -> Â x`, what `GCNConv.propagate` computes).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from ._lib import check, load, ptr, require_cuda, stream_ptr


def _device_edges(edge_index: torch.Tensor, edge_weight: Optional[torch.Tensor], device):
    if edge_index.dim() != 2 or edge_index.size(0) != 2:
        raise ValueError("edge_index must have shape [2, E]")
    ei = edge_index.to(device=device, dtype=torch.int64).contiguous()
    w = None
    if edge_weight is not None:
        if edge_weight.numel() != ei.size(1):
            raise ValueError("edge_weight must have one entry per edge")
        w = edge_weight.to(device=device, dtype=torch.float32).contiguous().view(-1)
    return ei, w


def gcn_norm(edge_index: torch.Tensor, edge_weight: Optional[torch.Tensor] = None, num_nodes: Optional[int] = None,
             device=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """`(edge_index, edge_weight)` with self loops added and `D^-1/2 A D^-1/2` weights (fp32), like torch_geometric's
    `gcn_norm(edge_index, edge_weight, num_nodes, improved=False, add_self_loops=True, flow="source_to_target")`:
    existing self loops are removed from the list and one loop per node is appended (keeping an existing loop's weight)."""
    require_cuda()
    if device is None:
        device = edge_index.device if edge_index.is_cuda else torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    ei, w = _device_edges(edge_index, edge_weight, device)
    e = ei.size(1)
    if num_nodes is None:
        num_nodes = int(ei.max()) + 1 if e else 0
    if e and (int(ei.min()) < 0 or int(ei.max()) >= num_nodes):
        raise ValueError("edge_index holds node ids outside [0, num_nodes)")
    out = torch.empty((2, e + num_nodes), dtype=torch.int64, device=device)
    out_w = torch.empty(e + num_nodes, dtype=torch.float32, device=device)
    count = torch.zeros(1, dtype=torch.int64, device=device)
    lib = load()
    with torch.cuda.device(device):
        check(lib.gsp_gcn_norm(num_nodes, e, ptr(ei[0]), ptr(ei[1]), ptr(w), ptr(out[0]), ptr(out[1]), ptr(out_w), ptr(count),
                               stream_ptr(device)))
    k = int(count)
    if k == e + num_nodes:
        return out, out_w
    return torch.stack((out[0, :k], out[1, :k])), out_w[:k].clone()


class GcnPropagation:
    """`x -> Â x` for a fixed (kept) edge list: `out[t] = sum_{e: target(e) = t} w_e * x[source(e)]` (GCNConv.propagate,
    flow source_to_target, sum aggregation), fp32, sums in edge-list order. With `normalize=True` (GCNConv's default) the
    list first goes through `gcn_norm`. The target grouping is built once and reused by every call (every layer, every epoch)."""

    def __init__(self, edge_index: torch.Tensor, edge_weight: Optional[torch.Tensor] = None, num_nodes: Optional[int] = None,
                 normalize: bool = True, device=None):
        require_cuda()
        if device is None:
            device = edge_index.device if edge_index.is_cuda else torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        if num_nodes is None:
            num_nodes = int(edge_index.max()) + 1 if edge_index.numel() else 0
        self.num_nodes = int(num_nodes)
        if normalize:
            self.edge_index, self.edge_weight = gcn_norm(edge_index, edge_weight, self.num_nodes, device=self.device)
        else:
            self.edge_index, w = _device_edges(edge_index, edge_weight, self.device)
            self.edge_weight = w if w is not None else torch.ones(self.edge_index.size(1), dtype=torch.float32, device=self.device)
        e = self.edge_index.size(1)
        self._row = self.edge_index[0].contiguous()
        self.indptr = torch.empty(self.num_nodes + 1, dtype=torch.int64, device=self.device)
        self.perm = torch.empty(max(e, 1), dtype=torch.int64, device=self.device)
        self._lib = load()
        with torch.cuda.device(self.device):
            check(self._lib.gsp_target_order(self.num_nodes, e, ptr(self.edge_index[1].contiguous()), ptr(self.indptr), ptr(self.perm),
                                             stream_ptr(self.device)))

    def __call__(self, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Forward only: the result is written by the kernel into a plain tensor and carries no autograd graph."""
        if x.dim() != 2 or x.size(0) != self.num_nodes:
            raise ValueError("x must have shape [num_nodes, dim]")
        if x.requires_grad and torch.is_grad_enabled():
            raise RuntimeError("GcnPropagation is inference-only (no backward pass): call it under torch.no_grad() or on a "
                               "detached tensor; inside a trained model it would silently cut the gradient to x")
        x = x.to(device=self.device, dtype=torch.float32)
        if x.stride(1) != 1:
            x = x.contiguous()
        if out is None:
            out = torch.empty((self.num_nodes, x.size(1)), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(self._lib.gsp_gcn_propagate(self.num_nodes, ptr(self.indptr), ptr(self.perm), ptr(self._row), ptr(self.edge_weight),
                                              ptr(x), x.size(1), x.stride(0), ptr(out), out.stride(0), stream_ptr(self.device)))
        return out
