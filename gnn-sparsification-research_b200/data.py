"""Minimal graph container used when PyTorch Geometric is not installed.

The reference hands `torch_geometric.data.Data` objects in and out of the
sparsification engine (reference `src/sparsification/core.py:12,63-76,244-245`).
The engine only relies on a handful of attributes and two methods:

    data.edge_index : LongTensor [2, E]
    data.x          : optional FloatTensor [N, d]
    data.num_nodes  : int
    data.clone()    : deep copy (every tensor attribute is cloned)
    data.to(device) : move every tensor attribute

`Data` below provides exactly that surface so the engine works without PyG.
When PyG *is* installed, real `torch_geometric.data.Data` objects are accepted
unchanged: the engine never constructs a `Data` itself, it only calls
`data.clone()` and re-assigns `edge_index` on the clone.
"""
from __future__ import annotations

import copy
from typing import Any, Iterator, Tuple

import torch


class Data:
    """Attribute bag with PyG-like `clone()` / `to()` / `num_nodes` semantics."""

    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, num_nodes=None, **kwargs: Any) -> None:
        self.x = x
        self.edge_index = edge_index
        self.edge_attr = edge_attr
        self.y = y
        if num_nodes is not None:
            self._num_nodes = int(num_nodes)
        for key, value in kwargs.items():
            setattr(self, key, value)

    # PyG infers num_nodes from x / edge_index when it was not given explicitly.
    @property
    def num_nodes(self) -> int:
        explicit = self.__dict__.get("_num_nodes")
        if explicit is not None:
            return explicit
        if self.x is not None:
            return int(self.x.size(0))
        if self.edge_index is not None and self.edge_index.numel() > 0:
            return int(self.edge_index.max()) + 1
        return 0

    @num_nodes.setter
    def num_nodes(self, value: int) -> None:
        self._num_nodes = int(value)

    @property
    def num_edges(self) -> int:
        return 0 if self.edge_index is None else int(self.edge_index.size(1))

    def _items(self) -> Iterator[Tuple[str, Any]]:
        return iter(self.__dict__.items())

    def clone(self) -> "Data":
        out = self.__class__.__new__(self.__class__)
        for key, value in self._items():
            if key.startswith("_gsp_"):          # stream bookkeeping of `to()`, not data
                continue
            out.__dict__[key] = value.clone() if torch.is_tensor(value) else copy.deepcopy(value)
        return out

    def to(self, device, non_blocking: bool = False) -> "Data":
        """Move every tensor attribute. `edge_index` goes first and, for an asynchronous upload to a CUDA device, an event
        marks its arrival: `GraphSparsifier` builds the CSR as soon as the edge list is there, while the (much larger)
        feature matrix is still crossing PCIe."""
        out = self.__class__.__new__(self.__class__)
        dev = torch.device(device) if not isinstance(device, torch.device) else device
        ordered = sorted(self._items(), key=lambda kv: kv[0] != "edge_index")        # stable: edge_index first
        for key, value in ordered:
            if key.startswith("_gsp_"):
                continue
            out.__dict__[key] = value.to(device, non_blocking=non_blocking) if torch.is_tensor(value) else value
            if (key == "edge_index" and non_blocking and dev.type == "cuda" and torch.is_tensor(value) and not value.is_cuda
                    and torch.cuda.is_available()):
                event = torch.cuda.Event()
                event.record(torch.cuda.current_stream(out.__dict__[key].device))
                out.__dict__["_gsp_edge_index_ready"] = event
        # keep the attribute order of the source object
        out.__dict__ = {k: out.__dict__[k] for k in list(self.__dict__.keys()) + ["_gsp_edge_index_ready"] if k in out.__dict__}
        return out

    def cpu(self) -> "Data":
        return self.to("cpu")

    def keys(self):
        return [k for k, v in self._items() if v is not None and not k.startswith("_")]

    def __repr__(self) -> str:
        parts = []
        for key, value in self._items():
            if value is None or key.startswith("_gsp_"):
                continue
            if torch.is_tensor(value):
                parts.append(f"{key.lstrip('_')}={list(value.shape)}")
            else:
                parts.append(f"{key.lstrip('_')}={value!r}")
        return f"Data({', '.join(parts)})"
