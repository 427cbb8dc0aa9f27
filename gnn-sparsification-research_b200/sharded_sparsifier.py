"""`GraphSparsifier` over the GPUs of one node: one process per GPU, the reference's method API unchanged.

The reference's drivers only know `GraphSparsifier(data, device)` + `compute_scores` + `sparsify*`
(reference scripts/nb05_roman_empire/roman_empire_gpu.py:213-259); its multi-GPU mode is one process per method.
`GraphSparsifier(data, device, group=...)` (which returns this class) keeps that surface and lets the ranks of a
`torch.distributed` group share ONE (graph, metric) job the way SURVEY §8e lays out:

  * inputs      `sharded_to_device(data, device, group)`: every rank uploads 1/N of `edge_index` and `x` from its own
                (pinned) host memory and the slices are all-gathered over NVLink — the replicas cost one PCIe crossing of
                1/N of the bytes per rank instead of N full uploads through the shared host;
  * scoring     Jaccard / Adamic-Adar owner-sharded with the exchange fused into the scoring kernel (peer stores into
                NVLink-mapped symmetric memory; NCCL reduce-scatter when symmetric memory is unavailable), FeatCos /
                degree on this rank's contiguous slice of canonical positions (the feature normalisation, 2.7 ms, is
                replicated: gathering row-sharded results costs four times as much), ApproxER column-sharded with an
                all-reduce of the partial sums;
  * selection   distributed radix select (16 KB histogram all-reduces), local compaction, kept `edge_index` slices
                all-gathered in rank order (== position order);
  * outputs     `compute_scores` returns THIS RANK'S slice of the fp64 score vector and `return_mask=True` this rank's
                slice of the mask (`local_range` = [lo, hi) of canonical positions): the host copies scale with 1/N too.
                `gather_outputs=True` returns the full vectors on every rank, like the single-GPU class.

Degree-aware and "-W" variants need global per-node maxima / extrema: the score slices are all-gathered over NVLink
(fp64 [E], once per metric) and the single-GPU kernels run replicated on every rank (the scoring, which is the cost,
was sharded) — identical results on every rank, bit-equal to the single-GPU class. Sampled / random use the host RNG
streams of the reference and are replicated the same way.

Graphs the sharded kernels do not cover (asymmetric pattern, duplicate edges: nnz != num_edges) are scored replicated on
every rank; results are still those of the single-GPU class.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import sharding
from .core import GraphSparsifier, _copy_stream, _to_host
from .engine import DeviceGraph, ShardedSelect

_PEER_CACHE: Dict[tuple, Optional[sharding.PeerScoreSlices]] = {}


def _peer_slices(nnz: int, group, device, tag: str) -> Optional[sharding.PeerScoreSlices]:
    """Symmetric-memory score slices, allocated once per (group, length, device, role): the rendezvous is a collective."""
    key = (id(group), int(nnz), str(device), tag)
    if key not in _PEER_CACHE:
        try:
            _PEER_CACHE[key] = sharding.PeerScoreSlices(nnz, group, device)
        except Exception:      # symmetric memory unavailable (old driver, gloo group ...): reduce-scatter path
            _PEER_CACHE[key] = None
        # every rank must take the same path
        flag = torch.tensor([0 if _PEER_CACHE[key] is None else 1], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag) == 0:
            _PEER_CACHE[key] = None
    return _PEER_CACHE[key]


def sharded_upload(t: torch.Tensor, dim: int, device, group) -> torch.Tensor:
    """Host tensor -> full replica on `device` on every rank: this rank copies its 1/N block along `dim` over PCIe
    (asynchronously when `t` is page-locked) and the blocks are all-gathered in place over NVLink."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if t.is_cuda or world == 1 or t.dim() == 0 or t.numel() < (1 << 16):
        return t.to(device, non_blocking=True)
    dim = dim % t.dim()
    if dim != 0 and not (t.dim() == 2 and dim == 1):
        return t.to(device, non_blocking=True)
    n = t.size(dim)
    length = (n + world - 1) // world
    lo, hi = min(rank * length, n), min((rank + 1) * length, n)
    if dim == 0:
        full = torch.empty((length * world,) + tuple(t.shape[1:]), dtype=t.dtype, device=device)
        full[lo:hi].copy_(t[lo:hi], non_blocking=True)
        dist.all_gather_into_tensor(full, full[rank * length:(rank + 1) * length], group=group)
        return full[:n]
    # [r, n] sharded along the columns (edge_index): one in-place gather per row keeps every row contiguous
    full = torch.empty((t.size(0), length * world), dtype=t.dtype, device=device)
    for r in range(t.size(0)):
        # row by row: each piece is contiguous on both sides (a strided host slice takes torch's slow staged copy: 6 GB/s)
        full[r, lo:hi].copy_(t[r, lo:hi], non_blocking=True)
        dist.all_gather_into_tensor(full[r], full[r, rank * length:(rank + 1) * length], group=group)
    return full[:, :n] if length * world == n else full[:, :n].contiguous()


def sharded_to_device(data, device, group):
    """`data.to(device)` for the ranks of `group`: same result (every tensor attribute a full replica on `device`), but
    each rank's PCIe link carries 1/N of `edge_index` and of every per-node tensor."""
    out = data.__class__.__new__(data.__class__)
    items = list(data._items()) if hasattr(data, "_items") else list(data.__dict__.items())
    moved = {}
    for key, value in sorted(items, key=lambda kv: kv[0] != "edge_index"):     # stable: edge_index first
        if key.startswith("_gsp_"):
            continue
        if torch.is_tensor(value):
            was_host = not value.is_cuda
            value = sharded_upload(value, 1 if key == "edge_index" else 0, device, group)
            if key == "edge_index" and was_host and value.is_cuda:
                # like Data.to(non_blocking=True): the sparsifier builds the CSR behind this event, on a side stream, while
                # the (larger) feature slices are still crossing PCIe and NVLink
                event = torch.cuda.Event()
                event.record(torch.cuda.current_stream(value.device))
                moved["_gsp_edge_index_ready"] = event
        moved[key] = value
    for key, _ in items:                                                        # keep the attribute order of the source
        if key in moved:
            out.__dict__[key] = moved[key]
    if "_gsp_edge_index_ready" in moved:
        out.__dict__["_gsp_edge_index_ready"] = moved["_gsp_edge_index_ready"]
    return out


class ShardedGraphSparsifier(GraphSparsifier):
    """See the module docstring. Construct it on every rank of `group` with the same `data`; every method is a
    collective call (same metric / ratio / order on every rank)."""

    def __init__(self, data, device: str, compute_device: Optional[str] = None, group=None, gather_outputs: bool = False):
        super().__init__(data, device, compute_device)
        self._group = group if group is not None else dist.group.WORLD
        self._world = dist.get_world_size(self._group)
        self._rank = dist.get_rank(self._group)
        self.gather_outputs = gather_outputs
        self._slices: Dict[str, torch.Tensor] = {}        # this rank's fp64 slice per metric
        self._node_range: Optional[Tuple[int, int]] = None
        self._sharded_ok: Optional[bool] = None

    # ------------------------------------------------------------------------------ layout
    def _build_graph(self) -> DeviceGraph:
        if self._graph is None and self.data.edge_index.is_cuda:
            return super()._build_graph()      # (side-stream build behind the arrival event of `sharded_to_device`)
        if self._graph is None:
            dev = self._cuda_device()
            ei = self.data.edge_index
            if ei.dtype != torch.int64:
                ei = ei.long()
            if not ei.is_cuda:      # host input: 1/N per PCIe link, replicas assembled over NVLink
                ei = sharded_upload(ei, 1, dev, self._group)
            self._ei_dev = ei.to(dev).contiguous()
            self._graph = DeviceGraph(self._ei_dev, self.num_nodes)
        return self._graph

    @property
    def sharded(self) -> bool:
        """True when the sharded kernels apply (symmetric pattern, one canonical position per edge_index column)."""
        if self._sharded_ok is None:
            g = self.graph
            self._sharded_ok = bool(self._world > 1 and g.symmetric and g.nnz == self.num_edges and g.nnz > 0)
        return self._sharded_ok

    @property
    def local_range(self) -> Tuple[int, int]:
        """[lo, hi) of canonical positions whose scores / mask bytes this rank returns."""
        if not self.sharded:
            return 0, self.graph.nnz
        _, slices = sharding.equal_slices(self.graph.nnz, self._world)
        return slices[self._rank]

    def _owner_range(self) -> Tuple[int, int]:
        if self._node_range is None:
            self._node_range = sharding.install_owner_deal(self.graph, self._world, self._rank)
        return self._node_range

    def _local(self, t: torch.Tensor) -> torch.Tensor:
        lo, hi = self.local_range
        return t[: hi - lo]

    # ------------------------------------------------------------------------------ scoring
    def _neighbourhood_slices(self, want_jaccard: bool, want_aa: bool) -> None:
        g = self.graph
        w = self._aa_node_weights() if want_aa else None
        nr = self._owner_range()
        if want_jaccard and want_aa:
            pj, pa = (_peer_slices(g.nnz, self._group, g.device, "jaccard"), _peer_slices(g.nnz, self._group, g.device, "aa"))
            if pj is not None and pa is not None:
                j, a = sharding.owner_sharded_jaccard_adamic_adar_p2p(g, pj, pa, nr, w)
                j, a = j.clone(), a.clone()          # the symmetric buffers are reused by the next call
            else:
                j, a = sharding.owner_sharded_jaccard_adamic_adar(g, self._group, nr, w)
            self._slices["jaccard"], self._slices["adamic_adar"] = self._local(j), self._local(a)
            return
        metric = "jaccard" if want_jaccard else "adamic_adar"
        peer = _peer_slices(g.nnz, self._group, g.device, "jaccard" if want_jaccard else "aa")
        if peer is not None:
            s = sharding.owner_sharded_scores_p2p(g, metric, peer, nr, w).clone()
        else:
            s = sharding.owner_sharded_scores(g, metric, self._group, nr, w)
        self._slices[metric] = self._local(s)

    def _normalized_features(self) -> torch.Tensor:
        """Normalised feature replica (reference metrics.py:344-346). Computed on every rank from its replica of `x`:
        2.7 ms at 16.7 M x 128 fp32, against ~11 ms for all-gathering row-sharded results over NVLink (measured on 8 GPUs:
        the gather moves 7/8 of 8.6 GB to every rank)."""
        if self._xhat is None:
            self._xhat = self.graph.normalize_features(self.data.x)
        return self._xhat

    def _slice_scores(self, key: str) -> torch.Tensor:
        """This rank's slice of the fp64 scores of `key` (device)."""
        if key in self._slices:
            return self._slices[key]
        g = self.graph
        lo, hi = self.local_range
        if key in self._score_cache:                     # injected host scores
            full = np.ascontiguousarray(self._score_cache[key], dtype=np.float64)
            t = torch.from_numpy(full[lo:hi] if full.shape[0] == g.nnz else full).to(g.device)
        elif key in ("jaccard", "adamic_adar"):
            self._neighbourhood_slices(key == "jaccard", key == "adamic_adar")
            return self._slices[key]
        elif key == "feature_cosine":
            if getattr(self.data, "x", None) is None:
                raise ValueError("feature_cosine requires node features (data.x)")
            t = g.feature_cosine(self._normalized_features(), lo, hi)
        elif key == "degree":
            t = g.degree_product(lo, hi)
        elif key == "approx_effective_resistance":
            from .metrics import _approx_er_on_graph
            t = _approx_er_on_graph(g, group=self._group, **self.approx_er_options)[lo:hi].contiguous()
        elif key == "effective_resistance":
            t = self._exact_effective_resistance()[lo:hi].contiguous()      # dense O(n^3) utility: replicated
        elif key == "random":
            # reference core.py:165-166: NumPy's global legacy RNG; rank 0 draws, every rank receives the same vector
            box = [np.random.rand(g.nnz) if self._rank == 0 else None]
            dist.broadcast_object_list(box, src=dist.get_global_rank(self._group, 0), group=self._group)
            t = torch.from_numpy(box[0][lo:hi]).to(g.device)
        else:
            raise ValueError(f"Internal error: Unhandled metric '{key}'")
        self._slices[key] = t
        return t

    def _gathered(self, key: str) -> torch.Tensor:
        """Full fp64 score vector on this rank's device (all-gather of the slices over NVLink, cached)."""
        if key not in self._dev_scores:
            s = self._slice_scores(key)
            length, _ = sharding.equal_slices(self.graph.nnz, self._world)
            full = torch.empty(length * self._world, dtype=torch.float64, device=s.device)
            mine = full[self._rank * length:(self._rank + 1) * length]
            mine[: s.numel()].copy_(s)
            dist.all_gather_into_tensor(full, mine, group=self._group)
            self._dev_scores[key] = full[: self.graph.nnz]
        return self._dev_scores[key]

    def _device_scores(self, metric: str) -> torch.Tensor:
        """Full-length device scores (what the replicated variants — degree-aware, "-W", sampled — consume)."""
        if not self.sharded:
            return super()._device_scores(metric)
        return self._gathered(self._normalize_metric_name(metric))

    def prefetch_scores(self, metrics, to_host: bool = False) -> None:
        """Announce the metric list: Jaccard and Adamic-Adar requested together come from ONE streaming pass per rank;
        `to_host=True` queues the read-back of every slice behind its kernel (copy stream, page-locked buffers)."""
        if not self.sharded:
            return super().prefetch_scores(metrics, to_host)
        keys = [self._normalize_metric_name(m) for m in metrics]
        pending = [k for k in keys if k not in self._slices and k not in self._score_cache]
        if "jaccard" in pending and "adamic_adar" in pending:
            self._neighbourhood_slices(True, True)
        for k in keys:
            t = self._slice_scores(k)
            if to_host and not self.gather_outputs and ("local", k) not in self._host_cache() and k not in self._host_pending:
                cs = _copy_stream(t.device)
                cs.wait_stream(torch.cuda.current_stream(t.device))
                host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                with torch.cuda.stream(cs):
                    host.copy_(t, non_blocking=True)
                    done = torch.cuda.Event()
                    done.record(cs)
                t.record_stream(cs)
                self._host_pending[k] = (host, done)

    def compute_scores(self, metric: str) -> np.ndarray:
        """float64 ndarray: this rank's slice (`local_range`), or the full vector with `gather_outputs=True`."""
        if not self.sharded:
            return super().compute_scores(metric)
        key = self._normalize_metric_name(metric)
        cache_key = key if self.gather_outputs else ("local", key)
        if cache_key not in self._host_cache():
            if not self.gather_outputs and key in self._host_pending:
                host, done = self._host_pending.pop(key)
                done.synchronize()
                self._host_cache()[cache_key] = host.numpy()
            else:
                src = self._gathered(key) if self.gather_outputs else self._slice_scores(key)
                self._host_cache()[cache_key] = _to_host(src).numpy()
        return self._host_cache()[cache_key]

    def _host_cache(self) -> dict:
        # local slices are kept apart from `_score_cache` (whose entries are full-length vectors by contract)
        if self.gather_outputs:
            return self._score_cache
        if not hasattr(self, "_local_host"):
            self._local_host = {}
        return self._local_host

    # ------------------------------------------------------------------------------ selection
    def sparsify(self, metric: str, retention_ratio: float, return_mask: bool = False, keep_lowest: bool = False):
        """Global top (bottom) `int(num_edges * r)` by score across all ranks (reference core.py:193-249): every rank
        returns the same `Data` (full kept `edge_index` on `device`); the mask is this rank's slice unless
        `gather_outputs`."""
        if not self.sharded:
            return super().sparsify(metric, retention_ratio, return_mask, keep_lowest)
        self._check_ratio(retention_ratio)
        if retention_ratio == 1.0:
            return self._full(return_mask)
        key = self._normalize_metric_name(metric)
        scores = self._slice_scores(key)
        g = self.graph
        num_keep = int(self.num_edges * retention_ratio)
        if keep_lowest:
            take = min(num_keep, g.nnz)
        elif num_keep == 0:
            take = g.nnz                                   # order[-0:] is the whole array (reference slicing quirk)
        else:
            take = min(num_keep, g.nnz)
        sparse_data, _, mask = self._select_sharded(scores, take, keep_lowest, return_mask)
        return (sparse_data, mask) if return_mask else sparse_data

    def _select_sharded(self, scores: torch.Tensor, take: int, keep_lowest: bool, want_mask: bool, with_weights: bool = False):
        """Distributed boundary search + this rank's mask slice, kept columns (and "-W" weights) from one fused emit
        (`engine.ShardedSelect`), then the ranks' kept lists concatenated in rank (= position) order."""
        lo, hi = self.local_range
        if getattr(self, "_select", None) is None:
            self._select = ShardedSelect(scores.device, self._group)
        mask_local, kept_local, w_local, count = self._select(scores, take, keep_lowest, self._ei_dev[:, lo:hi],
                                                              with_weights=with_weights, invert_weights=keep_lowest)
        k = int(count.item())
        kept = sharding.all_gather_variable(kept_local[:, :k], self._group, dim=1)
        w = sharding.all_gather_variable(w_local[:k], self._group, dim=0).to(self.device) if with_weights else None
        sparse_data = self.data.clone()
        sparse_data.edge_index = kept.to(self.device)
        if not want_mask:
            return sparse_data, w, None
        if self.gather_outputs:
            mask_local = sharding.all_gather_variable(mask_local, self._group, dim=0)
        return sparse_data, w, _to_host(mask_local.view(torch.bool))

    def sparsify_with_weights(self, metric: str, retention_ratio: float, keep_lowest: bool = False):
        """Threshold sparsification + min-max "-W" weights (reference roman_empire_gpu.py:248-256): the extrema of the
        kept scores are the boundary score and the globally best score, both known to every rank after the distributed
        boundary search, so each rank's emit writes the weights of its own kept edges; returns (Data, full float32
        weights on `device`, mask)."""
        if not self.sharded:
            return super().sparsify_with_weights(metric, retention_ratio, keep_lowest)
        self._check_ratio(retention_ratio)
        key = self._normalize_metric_name(metric)
        scores = self._slice_scores(key)
        num_keep = int(self.num_edges * retention_ratio)
        if retention_ratio == 1.0 or (num_keep == 0 and not keep_lowest):
            take = self.graph.nnz                           # everything (ratio 1, or the reference's order[-0:] quirk)
        else:
            take = min(num_keep, self.graph.nnz)
        return self._select_sharded(scores, take, keep_lowest, True, with_weights=True)

    # The replicated variants run the single-GPU code on the gathered score vector (`_device_scores`);
    # sparsify_degree_aware is inherited unchanged, the two that read host scores need the full vector.
    def _with_full_host_scores(self, fn, *args, **kwargs):
        saved, self.gather_outputs = self.gather_outputs, True
        try:
            out = fn(*args, **kwargs)
        finally:
            self.gather_outputs = saved
        return out

    def sparsify_sampled(self, metric: str, retention_ratio: float, seed: int = 42, return_mask: bool = False,
                         method: str = "numpy"):
        if not self.sharded:
            return super().sparsify_sampled(metric, retention_ratio, seed, return_mask, method)
        out = self._with_full_host_scores(super().sparsify_sampled, metric, retention_ratio, seed, return_mask, method)
        if return_mask and not self.gather_outputs:
            lo, hi = self.local_range
            return out[0], out[1][lo:hi]
        return out

    def sparsify_metric_backbone(self, metric: str, epsilon: float = 1e-9):
        if not self.sharded:
            return super().sparsify_metric_backbone(metric, epsilon)
        return self._with_full_host_scores(super().sparsify_metric_backbone, metric, epsilon)

    def _finish(self, mask_dev: torch.Tensor, num_kept: int, return_mask: bool):
        out = super()._finish(mask_dev, num_kept, return_mask)
        if return_mask and self.sharded and not self.gather_outputs:
            data, mask = out
            lo, hi = self.local_range
            return data, mask[lo:hi]
        return out
