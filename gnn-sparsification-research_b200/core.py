"""`GraphSparsifier` — drop-in for reference `src/sparsification/core.py:24-490`, executed on a B200.

Same constructor, attributes, method names, argument order, return types and error behaviour as the
reference class; every score vector and keep-mask is produced by the CUDA kernels of libgsp.so (through
`engine.py`). Scores are cached on the device; the NumPy arrays the reference API promises are
materialised lazily (`compute_scores` returns `np.ndarray[float64]` in canonical CSR order, exactly
like the reference). There is no CPU fallback: without CUDA every scoring/selection call raises.

Reference behaviours reproduced on purpose:
  * positional aliasing — scores are in canonical-CSR order but masks index `edge_index` columns by
    position (`core.py:239-242`), also for unsorted or duplicated `edge_index`;
  * `num_keep = int(num_edges * r)` and Python's slicing quirks (`order[-0:]` is everything);
  * ties resolved as a stable argsort would (top-k keeps the highest positions of the boundary tie
    class, bottom-k the lowest) — the reference's default argsort is an unstable SIMD sort, so this is
    the reproducible form of its contract;
  * `random` scores come from NumPy's global legacy RNG, `sparsify_sampled` draws with NumPy's PCG64
    `Generator.choice` on the host (its sequential fp64 cumsum cannot be matched bit-for-bit by a
    parallel scan) — RNG streams are part of the reference's observable behaviour.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib
from .engine import DeviceGraph, compact_edges, degree_aware_guarantee, select_compact, select_mask


def _to_host(t: torch.Tensor) -> torch.Tensor:
    """Device -> host through page-locked memory (torch's caching host allocator recycles the blocks), so the
    fp64 score vectors and masks the reference API returns on the host move at PCIe speed."""
    host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    host.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return host


_SIDE_STREAMS: Dict[int, "torch.cuda.Stream"] = {}
_COPY_STREAMS: Dict[int, "torch.cuda.Stream"] = {}


def _copy_stream(dev: torch.device) -> "torch.cuda.Stream":
    """One device-to-host copy stream per device for the whole process (read-backs overlap the kernels of the compute
    stream and, the link being full duplex, the uploads still in flight)."""
    index = dev.index if dev.index is not None else torch.cuda.current_device()
    if index not in _COPY_STREAMS:
        _COPY_STREAMS[index] = torch.cuda.Stream(torch.device("cuda", index))
    return _COPY_STREAMS[index]


def _side_stream(dev: torch.device) -> "torch.cuda.Stream":
    """One side stream per device for the whole process: torch's caching allocator keeps a pool per stream, so a fresh stream
    per sparsifier would strand the multi-GB score buffers of the previous one and pay cudaMalloc again."""
    index = dev.index if dev.index is not None else torch.cuda.current_device()
    if index not in _SIDE_STREAMS:
        _SIDE_STREAMS[index] = torch.cuda.Stream(torch.device("cuda", index))
    return _SIDE_STREAMS[index]


class GraphSparsifier:
    """Engine for graph sparsification via edge metric thresholding (reference core.py:24)."""

    SUPPORTED_METRICS = {
        "jaccard", "adamic-adar", "adamic_adar", "aa", "effective_resistance", "effective-resistance", "er",
        "approx_effective_resistance", "approx_er", "random", "rand", "degree", "feature_cosine", "feature-cosine",
    }
    _DISTANCE_METRICS = {"effective_resistance", "approx_effective_resistance"}

    def __new__(cls, data=None, device=None, compute_device=None, group=None, **kwargs):
        # extension: `group=` (a torch.distributed process group, one process per GPU) returns the sharded engine with
        # the same method API (sharded_sparsifier.py)
        if cls is GraphSparsifier and group is not None:
            from .sharded_sparsifier import ShardedGraphSparsifier
            return super().__new__(ShardedGraphSparsifier)
        return super().__new__(cls)

    def __init__(self, data, device: str, compute_device: Optional[str] = None, group=None) -> None:
        self.data = data
        self.device = device
        self.num_nodes = data.num_nodes
        self.num_edges = data.edge_index.size(1)
        self.verbose: bool = False
        self._score_cache: Dict[str, np.ndarray] = {}
        # --- B200 engine state (built lazily so argument validation needs no device) ---
        self._compute_device = compute_device
        self._graph: Optional[DeviceGraph] = None
        self._side: Optional[torch.cuda.Stream] = None      # set when the graph is built beside an asynchronous upload
        self._main: Optional[torch.cuda.Stream] = None
        self._side_pending = False
        self._ei_dev: Optional[torch.Tensor] = None
        self._dev_scores: Dict[str, torch.Tensor] = {}
        self._host_pending: Dict[str, Tuple[torch.Tensor, "torch.cuda.Event"]] = {}   # read-backs in flight (prefetch_scores)
        self._xhat: Optional[torch.Tensor] = None
        # extensions (not in the reference signature): ApproxER knobs, AA weight source
        self.approx_er_options = dict(epsilon=0.3, seed=42, max_cg_iters=500, cg_tol=1e-6, k=None, projection=None)
        self.aa_weights = "numpy"   # "numpy": libm-defined constant table from NumPy (bit parity); "device": CUDA log

    # ------------------------------------------------------------------------------ engine plumbing
    @property
    def scores(self) -> Dict[str, np.ndarray]:
        """Alias some reference callers read (scripts/nb07_synthetic/run_synthetic_hpo.py:275)."""
        return self._score_cache

    def _cuda_device(self) -> torch.device:
        _lib.require_cuda()
        if self._compute_device is not None:
            return torch.device(self._compute_device)
        ei = self.data.edge_index
        if ei.is_cuda:
            return ei.device
        dev = torch.device(self.device) if isinstance(self.device, str) else self.device
        if isinstance(dev, torch.device) and dev.type == "cuda":
            return torch.device("cuda", torch.cuda.current_device()) if dev.index is None else dev
        return torch.device("cuda", torch.cuda.current_device())

    def _build_graph(self) -> DeviceGraph:
        if self._graph is None:
            dev = self._cuda_device()
            ei = self.data.edge_index
            if ei.dtype != torch.int64:
                ei = ei.long()
            self._ei_dev = ei.to(dev, non_blocking=True).contiguous()
            ready = getattr(self.data, "_gsp_edge_index_ready", None)
            if ready is not None and ei.is_cuda and self._ei_dev.data_ptr() == ei.data_ptr():
                # `Data.to(cuda, non_blocking=True)` marked the arrival of the edge list: build the CSR on a side stream that
                # waits only for it, so the build overlaps the upload of the feature matrix queued behind it
                self._main = torch.cuda.current_stream(dev)
                self._side = _side_stream(dev)
                self._side.wait_event(ready)
                with torch.cuda.stream(self._side):
                    self._graph = DeviceGraph(self._ei_dev, self.num_nodes)
                    self._graph_ready = torch.cuda.Event()
                    self._graph_ready.record(self._side)
                self._ei_dev.record_stream(self._side)
                self._side_pending = True
            else:
                self._graph = DeviceGraph(self._ei_dev, self.num_nodes)
        return self._graph

    def _join_side(self) -> None:
        """ALL work queued on the side stream (graph build, neighbourhood scoring started by `prefetch_scores`) becomes a
        dependency of the caller's stream. Every access to cached device scores goes through here."""
        if self._side_pending:
            torch.cuda.current_stream(self._side.device).wait_stream(self._side)
            self._side_pending = False
            self._graph_ready = None

    def _join_graph(self) -> None:
        """Only the graph build becomes a dependency of the caller's stream (work that needs the graph but not the
        neighbourhood scores of a pass still running on the side stream)."""
        ready = getattr(self, "_graph_ready", None)
        if ready is not None:
            torch.cuda.current_stream(self._side.device).wait_event(ready)
            self._graph_ready = None

    @property
    def graph(self) -> DeviceGraph:
        self._build_graph()
        self._join_graph()
        return self._graph

    @property
    def adj(self):
        """SciPy CSR view of the canonical adjacency (reference attribute `adj`, core.py:71-74); host copy."""
        import scipy.sparse as sp

        g = self.graph
        indptr, indices, data, _ = g.export(with_data=True)
        return sp.csr_matrix((data.cpu().numpy(), indices.cpu().numpy(), indptr.cpu().numpy()),
                             shape=(self.num_nodes, self.num_nodes))

    # ------------------------------------------------------------------------------ name handling
    def _normalize_metric_name(self, metric: str) -> str:
        """reference core.py:118-138."""
        key = metric.lower().replace("-", "_").replace(" ", "_")
        table = {
            "jaccard": "jaccard", "adamic_adar": "adamic_adar", "aa": "adamic_adar",
            "effective_resistance": "effective_resistance", "er": "effective_resistance",
            "approx_effective_resistance": "approx_effective_resistance", "approx_er": "approx_effective_resistance",
            "random": "random", "rand": "random", "degree": "degree", "feature_cosine": "feature_cosine",
        }
        if key in table:
            return table[key]
        raise ValueError(f"Metric '{metric}' not supported. " f"Choose from: {self.SUPPORTED_METRICS}")

    def _scores_to_cost(self, scores: np.ndarray, metric: str) -> np.ndarray:
        """Similarity -> distance for the metric backbone, d = 1/p - 1 (reference core.py:82-116)."""
        key = GraphSparsifier._normalize_metric_name(self, metric)
        if key in GraphSparsifier._DISTANCE_METRICS:
            similarity = 1.0 / np.maximum(scores, 1e-10)
        else:
            similarity = scores.copy()
        top = similarity.max()
        if top <= 0:
            return np.ones_like(scores)
        proximity = similarity / top
        positive = proximity[proximity > 0]
        floor = (positive.min() * 0.01) if len(positive) > 0 else 1e-6
        proximity[proximity <= 0] = floor
        return 1.0 / proximity - 1.0

    # ------------------------------------------------------------------------------ scoring
    def _device_scores(self, metric: str) -> torch.Tensor:
        """fp64 score tensor on the GPU, canonical CSR order, length nnz (cached per metric)."""
        key = self._normalize_metric_name(metric)
        if key in self._dev_scores:
            self._join_side()
            return self._dev_scores[key]
        g = self.graph
        if key in self._score_cache:                       # injected / host-side scores: upload once
            t = torch.from_numpy(np.ascontiguousarray(self._score_cache[key], dtype=np.float64)).to(g.device)
        elif key == "jaccard":
            t = g.jaccard()
        elif key == "adamic_adar":
            t = g.adamic_adar(self._aa_node_weights())
        elif key == "effective_resistance":
            t = self._exact_effective_resistance()
        elif key == "approx_effective_resistance":
            t = self._approx_effective_resistance()
        elif key == "random":
            # reference core.py:165-166: NumPy's global legacy RNG (host state is the behaviour)
            t = torch.from_numpy(np.random.rand(g.nnz)).to(g.device)
        elif key == "degree":
            t = g.degree_product()
        elif key == "feature_cosine":
            if getattr(self.data, "x", None) is None:
                raise ValueError("feature_cosine requires node features (data.x)")
            if self._xhat is None:
                self._xhat = g.normalize_features(self.data.x)
            t = g.feature_cosine(self._xhat)
        else:  # unreachable
            raise ValueError(f"Internal error: Unhandled metric '{key}'")
        self._dev_scores[key] = t
        return t

    def prefetch_scores(self, metrics, to_host: bool = False) -> None:
        """Extension: announce the metrics a caller is about to use (the reference's drivers loop over a fixed method list,
        scripts/nb05_roman_empire/roman_empire_gpu.py:81-102, src/experiments/ablation.py:220-270). Jaccard and Adamic-Adar
        requested together come from ONE streaming pass over the neighbour lists (`gsp_jaccard_adamic_adar`): the hit
        ballots of the ordered Adamic-Adar sum also give the intersection count. Results are bit-identical to separate
        `compute_scores` calls; everything else is computed as usual and cached on the device. `to_host=True` also starts the
        read-back of every announced vector as soon as its kernel is queued (see `_start_host_copies`)."""
        keys = [self._normalize_metric_name(m) for m in metrics]
        pending = [k for k in keys if k not in self._dev_scores and k not in self._score_cache]
        if "jaccard" in pending and "adamic_adar" in pending:
            g = self._build_graph()
            if self._side_pending:
                # the graph was built beside the feature upload: the neighbourhood pass only needs the graph, so it starts on
                # the same side stream while the features are still arriving (everything else joins before it touches either)
                with torch.cuda.stream(self._side):
                    jac, aa = g.jaccard_adamic_adar(g.aa_node_weights_numpy() if self.aa_weights == "numpy" else None)
                for t in (jac, aa):
                    t.record_stream(self._main)
            else:
                jac, aa = g.jaccard_adamic_adar(self._aa_node_weights())
            self._dev_scores["jaccard"], self._dev_scores["adamic_adar"] = jac, aa
        if to_host:
            self._start_host_copies([k for k in keys if k in self._dev_scores])    # the neighbourhood scores first
        for k in keys:
            if k not in self._dev_scores and k not in self._score_cache:
                # everything else runs BEHIND the neighbourhood pass: beside it, feature-cosine stretched the pass from 184 to
                # ~240 ms (two 768-thread hub CTAs fill an SM; the kernels time-slice at CTA granularity)
                self._join_side()
            self._device_scores(k)
            if to_host:
                self._start_host_copies([k])

    def _start_host_copies(self, keys) -> None:
        """Queue the device-to-host copy of the fp64 vectors the reference API returns on the host (page-locked buffers,
        a copy stream that waits only for the producing stream): `compute_scores` then waits for ITS vector instead of
        running each 2 GB read-back after the previous kernel, one blocking copy at a time."""
        for key in keys:
            if key in self._score_cache or key in self._host_pending:
                continue
            t = self._dev_scores[key]
            cs = _copy_stream(t.device)
            producer = self._side if (self._side_pending and self._side is not None) else torch.cuda.current_stream(t.device)
            ready = torch.cuda.Event()
            ready.record(producer)
            host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            cs.wait_event(ready)
            with torch.cuda.stream(cs):
                host.copy_(t, non_blocking=True)
                done = torch.cuda.Event()
                done.record(cs)
            t.record_stream(cs)
            self._host_pending[key] = (host, done)

    def _aa_node_weights(self) -> Optional[torch.Tensor]:
        """Adamic-Adar node weights 1/sqrt(max(log(deg+1),1e-10)) (reference metrics.py:104-108).

        The weight depends only on the integer degree, and its last bit is defined by NumPy's libm/SIMD
        `log`. "numpy" evaluates the reference expression once per distinct degree value (a constant table
        of at most max_degree+1 entries) and gathers it on the device, so scores are bit-equal to the
        reference; "device" uses CUDA's log (<= 1 ulp apart, scores within 1e-6)."""
        if self.aa_weights != "numpy":
            return None
        return self.graph.aa_node_weights_numpy()

    def _approx_effective_resistance(self) -> torch.Tensor:
        from .metrics import _approx_er_on_graph
        return _approx_er_on_graph(self.graph, **self.approx_er_options)

    def _exact_effective_resistance(self) -> torch.Tensor:
        from .metrics import _exact_er_on_graph
        return _exact_er_on_graph(self.graph)

    def compute_scores(self, metric: str) -> np.ndarray:
        """Edge scores as float64 ndarray in canonical CSR order (reference core.py:140-191)."""
        key = self._normalize_metric_name(metric)
        if key not in self._score_cache:
            if key in self._host_pending:               # read-back started by prefetch_scores(to_host=True)
                host, done = self._host_pending.pop(key)
                done.synchronize()
                self._score_cache[key] = host.numpy()
            else:
                self._score_cache[key] = _to_host(self._device_scores(key)).numpy()
        return self._score_cache[key]

    # ------------------------------------------------------------------------------ selection
    @staticmethod
    def _check_ratio(retention_ratio: float) -> None:
        if not 0 < retention_ratio <= 1:
            raise ValueError(f"retention_ratio must be in (0, 1], got {retention_ratio}")

    def _full(self, return_mask: bool):
        if return_mask:
            return self.data.clone(), torch.ones(self.num_edges, dtype=torch.bool)
        return self.data.clone()

    def _finish(self, mask_dev: torch.Tensor, num_kept: int, return_mask: bool):
        """mask (uint8, device, length num_edges) -> reference return value (core.py:242-249)."""
        kept, _, _ = compact_edges(self._ei_dev, mask_dev, num_kept)
        sparse_data = self.data.clone()
        sparse_data.edge_index = kept.to(self.device)
        if return_mask:
            return sparse_data, _to_host(mask_dev.view(torch.bool))   # 0/1 bytes reinterpreted, no conversion pass
        return sparse_data

    def _threshold_mask(self, scores: torch.Tensor, num_keep: int, keep_lowest: bool) -> Tuple[torch.Tensor, int]:
        """Device keep-mask over `num_edges` positions with the reference's slicing semantics (core.py:232-240)."""
        nnz = scores.numel()
        if keep_lowest:
            take = min(num_keep, nnz)                 # order[:num_keep]
        elif num_keep == 0:
            take = nnz                                # order[-0:] is the whole array
        else:
            take = min(num_keep, nnz)                 # order[-num_keep:]
        mask = torch.zeros(self.num_edges, dtype=torch.uint8, device=scores.device)
        if take > 0:
            select_mask(scores, take, keep_lowest, out=mask[:nnz])
        return mask, take

    def _take(self, num_keep: int, keep_lowest: bool, nnz: int) -> int:
        """How many of the `nnz` scored positions the reference's slices keep (core.py:232-237): `order[:k]`,
        `order[-k:]` — and `order[-0:]` is the whole array."""
        if not keep_lowest and num_keep == 0:
            return nnz
        return min(num_keep, nnz)

    def _select_finish(self, scores: torch.Tensor, take: int, keep_lowest: bool, return_mask: bool, with_weights: bool = False):
        """Boundary search, keep-mask, `edge_index[:, mask]` and (optionally) the "-W" weights in one library call
        (`gsp_select_compact`); positions >= nnz (duplicate edges) are never kept (reference core.py:239-245)."""
        nnz = scores.numel()
        mask = None
        if return_mask:
            mask = (torch.zeros if nnz < self.num_edges or take == 0 else torch.empty)(self.num_edges, dtype=torch.uint8,
                                                                                       device=scores.device)
        weights = None
        if take > 0:
            kept, weights, _ = select_compact(scores, take, keep_lowest, self._ei_dev, mask, with_weights, invert_weights=keep_lowest)
        else:
            kept = torch.empty((2, 0), dtype=torch.int64, device=scores.device)
            weights = torch.empty(0, dtype=torch.float32, device=scores.device) if with_weights else None
        sparse_data = self.data.clone()
        sparse_data.edge_index = kept.to(self.device)
        host_mask = _to_host(mask.view(torch.bool)) if return_mask else None   # 0/1 bytes reinterpreted, no conversion pass
        return sparse_data, weights, host_mask

    def sparsify(self, metric: str, retention_ratio: float, return_mask: bool = False, keep_lowest: bool = False):
        """Keep the top (or bottom) `int(num_edges * retention_ratio)` edges by score (reference core.py:193-249)."""
        self._check_ratio(retention_ratio)
        if retention_ratio == 1.0:
            return self._full(return_mask)
        scores = self._device_scores(metric)
        take = self._take(int(self.num_edges * retention_ratio), keep_lowest, scores.numel())
        sparse_data, _, mask = self._select_finish(scores, take, keep_lowest, return_mask)
        return (sparse_data, mask) if return_mask else sparse_data

    def sparsify_with_weights(self, metric: str, retention_ratio: float, keep_lowest: bool = False):
        """Extension: threshold sparsification plus the min-max "-W" edge weights in one device pass.

        Returns (Data, edge_weight float32 on `self.device`, mask). Weight formula: reference
        scripts/nb05_roman_empire/roman_empire_gpu.py:248-256."""
        self._check_ratio(retention_ratio)
        scores = self._device_scores(metric)
        if retention_ratio == 1.0:
            mask = torch.ones(self.num_edges, dtype=torch.uint8, device=scores.device)
            if scores.numel() < self.num_edges:
                raise IndexError("boolean index did not match indexed array (duplicate edges)")
            ei, w, _ = compact_edges(self._ei_dev, mask, self.num_edges, scores=scores, with_weights=True, invert_weights=keep_lowest)
            sparse_data = self.data.clone()
            sparse_data.edge_index = ei.to(self.device)
            return sparse_data, w.to(self.device), _to_host(mask.view(torch.bool))
        take = self._take(int(self.num_edges * retention_ratio), keep_lowest, scores.numel())
        sparse_data, w, mask = self._select_finish(scores, take, keep_lowest, True, with_weights=True)
        return sparse_data, w.to(self.device), mask

    def sparsify_metric_backbone(self, metric: str, epsilon: float = 1e-9):
        """Global metric backbone: keep edge (u,v) iff its cost <= shortest-path cost + epsilon (reference core.py:251-279).

        Returns `(sparse_data_on_device, stats)`; the retention ratio follows from the graph, not from a parameter."""
        from .metric_backbone import compute_metric_backbone

        distances = self._scores_to_cost(self.compute_scores(metric), metric)
        sparse_data, stats = compute_metric_backbone(self.data, distances, epsilon=epsilon, verbose=self.verbose)
        return sparse_data.to(self.device), stats

    def sparsify_sampled(self, metric: str, retention_ratio: float, seed: int = 42, return_mask: bool = False,
                         method: str = "numpy"):
        """Sample edges without replacement with probability proportional to score (reference core.py:281-357).

        `method="numpy"` (default) reproduces the reference's draw bit for bit: NumPy's PCG64 `Generator.choice`
        with its sequential fp64 cumsum runs on the host (an RNG stream is observable behaviour). `method="device"`
        is an opt-in extension that never leaves the GPU: Efraimidis-Spirakis keys u^(1/p) (u ~ torch Philox) and the
        radix select — the same sampling distribution (successive weighted draws without replacement), a different
        random stream, so not mask-identical to the reference."""
        self._check_ratio(retention_ratio)
        if retention_ratio == 1.0:
            return self._full(return_mask)
        if method == "device":
            scores = self._device_scores(metric)
            if scores.numel() != self.num_edges:
                raise ValueError("'a' and 'p' must have same size")
            floor = 1e-8
            p = torch.nan_to_num(scores, nan=floor, posinf=floor, neginf=floor).clamp_min(floor)
            gen = torch.Generator(device=p.device)
            gen.manual_seed(int(seed))
            u = torch.rand(p.shape, dtype=torch.float64, device=p.device, generator=gen).clamp_min(1e-300)
            keys = torch.log(u) / p                                   # log of u^(1/p): larger = drawn earlier
            num_keep = int(self.num_edges * retention_ratio)
            mask_dev, kept = self._threshold_mask(keys, num_keep, keep_lowest=False) if num_keep > 0 else (
                torch.zeros(self.num_edges, dtype=torch.uint8, device=p.device), 0)
            return self._finish(mask_dev, kept, return_mask)
        if method != "numpy":
            raise ValueError("method must be 'numpy' or 'device'")
        rng = np.random.default_rng(seed)
        scores = self.compute_scores(metric)
        floor = 1e-8
        scores = np.nan_to_num(scores, nan=floor, posinf=floor, neginf=floor)
        probs = np.maximum(scores, floor)
        probs = probs / probs.sum()
        num_keep = int(self.num_edges * retention_ratio)
        # NumPy's Generator.choice: PCG64 stream + sequential fp64 cumsum are the reference's observable RNG behaviour
        selected = rng.choice(self.num_edges, size=num_keep, replace=False, p=probs)
        mask = np.zeros(self.num_edges, dtype=np.uint8)
        mask[selected] = 1
        self.graph  # make sure the device edge list exists
        mask_dev = torch.from_numpy(mask).to(self._ei_dev.device)
        return self._finish(mask_dev, num_keep, return_mask)

    def sparsify_degree_aware(self, metric: str, retention_ratio: float, min_edges_per_node: int = 1,
                              return_mask: bool = False):
        """Per-node minimum edge budget, then global fill (reference core.py:359-461)."""
        self._check_ratio(retention_ratio)
        if retention_ratio == 1.0:
            return self._full(return_mask)
        scores = self._device_scores(metric)
        if scores.numel() < self.num_edges and self.num_edges > 0:
            # reference: scores[incident_indices] with positions >= nnz (duplicate edges) -> IndexError
            raise IndexError(f"index {self.num_edges - 1} is out of bounds for axis 0 with size {scores.numel()}")
        num_keep = int(self.num_edges * retention_ratio)
        src = self._ei_dev[0]
        m = int(min_edges_per_node)
        if m < 0:
            raise ValueError("min_edges_per_node must be >= 0")   # (the reference's `[-k:]` with k < 0 is an accident, not an API)
        if m == 0:
            # reference core.py:432-434: `argsort(...)[-0:]` is the whole array, so EVERY incident edge is guaranteed and
            # the budget is exceeded: the full graph comes back
            mask = torch.ones(self.num_edges, dtype=torch.uint8, device=scores.device)
            return self._finish(mask, self.num_edges, return_mask)
        mask, marked = degree_aware_guarantee(src, scores, self.num_nodes, m)
        guaranteed = int(marked.item())            # one 8-byte host read: the output size depends on it
        kept = guaranteed
        if guaranteed < num_keep:
            select_mask(scores, num_keep - guaranteed, keep_lowest=False, exclude=mask, out=mask, or_into=True)
            kept = num_keep
        return self._finish(mask, kept, return_mask)

    def get_retention_curve_data(self, metric: str, retention_rates):
        """reference core.py:463-480."""
        return [self.sparsify(metric, rate) for rate in retention_rates]

    @property
    def stats(self) -> dict:
        """reference core.py:482-490."""
        return {
            "num_nodes": self.num_nodes,
            "num_edges": self.num_edges,
            "density": self.num_edges / (self.num_nodes * (self.num_nodes - 1)),
            "avg_degree": self.num_edges / self.num_nodes,
        }
