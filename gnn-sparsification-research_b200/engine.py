"""Device-tensor layer over the C ABI: torch owns the memory and the streams, libgsp.so does the work.

Everything here takes and returns CUDA tensors on the graph's device and enqueues on torch's current
stream; nothing synchronises except graph construction (one host read of the merged nnz) and the
explicit `.item()` calls documented below.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr


class DeviceGraph:
    """Canonical CSR of a directed edge list, resident on one GPU (reference core.py:70-74)."""

    def __init__(self, edge_index: torch.Tensor, num_nodes: int, values: Optional[torch.Tensor] = None):
        _lib.require_cuda()
        if not edge_index.is_cuda:
            raise _lib.GspError("DeviceGraph needs a CUDA edge_index")
        if edge_index.dtype != torch.int64 or edge_index.dim() != 2 or edge_index.size(0) != 2:
            raise ValueError("edge_index must be an int64 tensor of shape [2, E]")
        self.device = edge_index.device
        self._lib = _lib.load()
        self._handle = C.c_void_p()
        ei = edge_index if edge_index.is_contiguous() else edge_index.contiguous()
        e = ei.size(1)
        val = None
        if values is not None:
            val = values.to(device=self.device, dtype=torch.float64).contiguous()
        with torch.cuda.device(self.device):
            row_ptr = C.c_void_p(ei.data_ptr()) if e else None
            col_ptr = C.c_void_p(ei.data_ptr() + 8 * e) if e else None
            check(self._lib.gsp_graph_create(int(num_nodes), e, row_ptr, col_ptr, ptr(val), stream_ptr(self.device),
                                             C.byref(self._handle)))
        info = _lib.GraphInfo()
        check(self._lib.gsp_graph_get_info(self._handle, C.byref(info)))
        self.num_nodes = info.num_nodes
        self.num_input_edges = info.num_input_edges
        self.nnz = info.nnz
        self.num_undirected = info.num_undirected
        self.max_degree = info.max_degree
        self.sum_degree_sq = info.sum_degree_sq
        self.symmetric = bool(info.symmetric)
        self.input_canonical = bool(info.input_canonical)
        self.unit_weights = bool(info.unit_weights)

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            try:
                self._lib.gsp_graph_destroy(h)
            except Exception:
                pass
            self._handle = None

    # -- helpers ---------------------------------------------------------------------------------
    def _stream(self):
        return stream_ptr(self.device)

    def _range(self, e_begin, e_end) -> Tuple[int, int]:
        e_begin = 0 if e_begin is None else int(e_begin)
        e_end = self.nnz if e_end is None else int(e_end)
        if not 0 <= e_begin <= e_end <= self.nnz:
            raise ValueError(f"edge range [{e_begin}, {e_end}) outside [0, {self.nnz}]")
        return e_begin, e_end

    def _empty(self, n, dtype):
        return torch.empty(int(n), dtype=dtype, device=self.device)

    # -- CSR export ------------------------------------------------------------------------------
    def export(self, with_data=True, with_rows=False):
        indptr = self._empty(self.num_nodes + 1, torch.int64)
        indices = self._empty(self.nnz, torch.int32)
        data = self._empty(self.nnz, torch.float64) if with_data else None
        rows = self._empty(self.nnz, torch.int32) if with_rows else None
        with torch.cuda.device(self.device):
            check(self._lib.gsp_graph_export(self._handle, ptr(indptr), ptr(indices), ptr(data), ptr(rows), self._stream()))
        return indptr, indices, data, rows

    def degrees(self) -> torch.Tensor:
        deg = self._empty(self.num_nodes, torch.int32)
        with torch.cuda.device(self.device):
            check(self._lib.gsp_graph_degrees(self._handle, ptr(deg), self._stream()))
        return deg

    def undirected_ids(self) -> torch.Tensor:
        uid = self._empty(self.nnz, torch.int32)
        with torch.cuda.device(self.device):
            check(self._lib.gsp_graph_undirected_ids(self._handle, ptr(uid), self._stream()))
        return uid

    # -- scoring ---------------------------------------------------------------------------------
    def jaccard(self, e_begin=None, e_end=None, return_counts=False, out=None):
        b, e = self._range(e_begin, e_end)
        score = self._empty(e - b, torch.float64) if out is None else out
        inter = self._empty(e - b, torch.int32) if return_counts else None
        with torch.cuda.device(self.device):
            check(self._lib.gsp_jaccard(self._handle, b, e, ptr(inter), ptr(score), self._stream()))
        return (score, inter) if return_counts else score

    def aa_node_weights(self) -> torch.Tensor:
        w = self._empty(self.num_nodes, torch.float64)
        with torch.cuda.device(self.device):
            check(self._lib.gsp_aa_node_weights(self._handle, ptr(w), self._stream()))
        return w

    def aa_node_weights_numpy(self) -> torch.Tensor:
        """Node weights with NumPy's bits: the reference expression (metrics.py:104-108) evaluated once per distinct
        degree value on the host (a table of max_degree + 1 constants), gathered per node on the device."""
        import numpy as np

        if getattr(self, "_aa_table", None) is None:
            t = np.arange(self.max_degree + 1, dtype=np.float64)
            host = torch.from_numpy(1.0 / np.sqrt(np.maximum(np.log(t + 1), 1e-10))).pin_memory()
            # fetched by a kernel from page-locked memory: the copy engine may be busy with a feature upload queued earlier
            self._aa_table = self._empty(host.numel(), torch.float64)
            with torch.cuda.device(self.device):
                check(self._lib.gsp_copy_f64(ptr(host), ptr(self._aa_table), host.numel(), self._stream()))
            self._aa_table_host = host      # stays alive until the kernel has read it (and for reuse)
        w = self._empty(self.num_nodes, torch.float64)
        with torch.cuda.device(self.device):
            check(self._lib.gsp_aa_node_weights_from_table(self._handle, ptr(self._aa_table), self._aa_table.numel(), ptr(w),
                                                           self._stream()))
        return w

    def adamic_adar(self, node_weights: Optional[torch.Tensor] = None, e_begin=None, e_end=None, out=None):
        b, e = self._range(e_begin, e_end)
        if node_weights is not None:
            node_weights = node_weights.to(device=self.device, dtype=torch.float64).contiguous()
            if node_weights.numel() != self.num_nodes:
                raise ValueError("node_weights must have num_nodes entries")
        score = self._empty(e - b, torch.float64) if out is None else out
        with torch.cuda.device(self.device):
            check(self._lib.gsp_adamic_adar(self._handle, ptr(node_weights), b, e, ptr(score), self._stream()))
        return score

    def jaccard_adamic_adar(self, node_weights: Optional[torch.Tensor] = None, e_begin=None, e_end=None, return_counts=False,
                            out_jaccard=None, out_adamic_adar=None):
        """Both neighbourhood scores from ONE streaming pass (bit-identical to `jaccard()` and `adamic_adar()`)."""
        b, e = self._range(e_begin, e_end)
        if node_weights is not None:
            node_weights = node_weights.to(device=self.device, dtype=torch.float64).contiguous()
            if node_weights.numel() != self.num_nodes:
                raise ValueError("node_weights must have num_nodes entries")
        jac = self._empty(e - b, torch.float64) if out_jaccard is None else out_jaccard
        aa = self._empty(e - b, torch.float64) if out_adamic_adar is None else out_adamic_adar
        inter = self._empty(e - b, torch.int32) if return_counts else None
        with torch.cuda.device(self.device):
            check(self._lib.gsp_jaccard_adamic_adar(self._handle, ptr(node_weights), b, e, ptr(inter), ptr(jac), ptr(aa),
                                                    self._stream()))
        return (jac, aa, inter) if return_counts else (jac, aa)

    # -- owner-sharded scoring (multi-GPU; see sharding.py) -------------------------------------------
    def owner_costs(self) -> torch.Tensor:
        cost = self._empty(self.num_nodes, torch.float64)
        with torch.cuda.device(self.device):
            check(self._lib.gsp_owner_costs(self._handle, ptr(cost), self._stream()))
        return cost

    def set_owner_deal(self, owner_rank: Optional[torch.Tensor], rank: int = 0) -> None:
        """Dealt ownership (`gsp_graph_set_owner_deal`): from now on the `*_owned*` calls of this handle evaluate only
        owners o with `owner_rank[o] == rank` (uint8[num_nodes], copied); None clears the deal."""
        if owner_rank is not None:
            owner_rank = owner_rank.to(device=self.device, dtype=torch.uint8).contiguous()
            if owner_rank.numel() != self.num_nodes:
                raise ValueError("owner_rank must have num_nodes entries")
        with torch.cuda.device(self.device):
            check(self._lib.gsp_graph_set_owner_deal(self._handle, ptr(owner_rank), int(rank), self._stream()))

    def jaccard_owned(self, node_begin: int, node_end: int, out: torch.Tensor, counts: Optional[torch.Tensor] = None):
        """Scores of the pairs owned by nodes [node_begin, node_end) into the full-length (>= nnz) buffer `out`."""
        with torch.cuda.device(self.device):
            check(self._lib.gsp_jaccard_owned(self._handle, int(node_begin), int(node_end), ptr(counts), ptr(out), self._stream()))
        return out

    def adamic_adar_owned(self, node_weights: Optional[torch.Tensor], node_begin: int, node_end: int, out: torch.Tensor):
        with torch.cuda.device(self.device):
            check(self._lib.gsp_adamic_adar_owned(self._handle, ptr(node_weights), int(node_begin), int(node_end), ptr(out),
                                                  self._stream()))
        return out

    def jaccard_adamic_adar_owned(self, node_weights: Optional[torch.Tensor], node_begin: int, node_end: int,
                                  out_jaccard: torch.Tensor, out_adamic_adar: torch.Tensor):
        with torch.cuda.device(self.device):
            check(self._lib.gsp_jaccard_adamic_adar_owned(self._handle, ptr(node_weights), int(node_begin), int(node_end),
                                                          ptr(out_jaccard), ptr(out_adamic_adar), self._stream()))
        return out_jaccard, out_adamic_adar

    def owned_scatter(self, metric: str, node_begin: int, node_end: int, slices_dev_ptr: int, world: int, slice_len: int,
                      node_weights: Optional[torch.Tensor] = None, jaccard_slices_dev_ptr: Optional[int] = None) -> None:
        """Score the pairs owned by [node_begin, node_end) and store every score straight into the owning rank's slice;
        `slices_dev_ptr` is the device address of an array of `world` slice base pointers (peer memory allowed).
        metric "jaccard+adamic_adar": Adamic-Adar goes to `slices_dev_ptr`, Jaccard to `jaccard_slices_dev_ptr`."""
        sl = C.c_void_p(int(slices_dev_ptr))
        with torch.cuda.device(self.device):
            if metric == "jaccard":
                check(self._lib.gsp_jaccard_owned_scatter(self._handle, int(node_begin), int(node_end), sl, int(world),
                                                          int(slice_len), self._stream()))
            elif metric == "adamic_adar":
                check(self._lib.gsp_adamic_adar_owned_scatter(self._handle, ptr(node_weights), int(node_begin), int(node_end), sl,
                                                              int(world), int(slice_len), self._stream()))
            elif metric == "jaccard+adamic_adar":
                check(self._lib.gsp_jaccard_adamic_adar_owned_scatter(
                    self._handle, ptr(node_weights), int(node_begin), int(node_end), C.c_void_p(int(jaccard_slices_dev_ptr)), sl,
                    int(world), int(slice_len), self._stream()))
            else:
                raise ValueError(metric)

    def degree_product(self, e_begin=None, e_end=None, out=None):
        b, e = self._range(e_begin, e_end)
        score = self._empty(e - b, torch.float64) if out is None else out
        with torch.cuda.device(self.device):
            check(self._lib.gsp_degree_product(self._handle, b, e, ptr(score), self._stream()))
        return score

    def normalize_features(self, x: torch.Tensor, packed: Optional[bool] = None) -> torch.Tensor:
        """Row-normalised features in the dtype of `x` (fp32 or fp64), reference metrics.py:344-346.

        fp32 features with dim in {32, 64, 96, 128} are stored in the accumulator-major packed layout of
        `gsp_featcos_f32_packed` (the returned tensor carries `_gsp_packed = True`; it is only meaningful as input of
        `feature_cosine`). `packed=False` forces the plain layout."""
        if x.dim() != 2 or x.size(0) != self.num_nodes:
            raise ValueError("features must have shape [num_nodes, d]")
        if x.dtype not in (torch.float32, torch.float64):
            x = x.to(torch.float64)  # NumPy promotes integer features to float64 in linalg.norm / divide
        x = x.to(self.device).contiguous()
        xhat = torch.empty_like(x)
        dim = x.size(1)
        use_packed = x.dtype == torch.float32 and dim in (32, 64, 96, 128) and packed is not False
        if use_packed:
            fn = self._lib.gsp_featcos_normalize_f32_packed
        else:
            fn = self._lib.gsp_featcos_normalize_f32 if x.dtype == torch.float32 else self._lib.gsp_featcos_normalize_f64
        with torch.cuda.device(self.device):
            check(fn(x.size(0), dim, ptr(x), dim, ptr(xhat), dim, self._stream()))
        xhat._gsp_packed = use_packed
        return xhat

    def feature_cosine(self, xhat: torch.Tensor, e_begin=None, e_end=None, out=None):
        b, e = self._range(e_begin, e_end)
        score = self._empty(e - b, torch.float64) if out is None else out
        if getattr(xhat, "_gsp_packed", False):
            fn = self._lib.gsp_featcos_f32_packed
        else:
            fn = self._lib.gsp_featcos_f32 if xhat.dtype == torch.float32 else self._lib.gsp_featcos_f64
        with torch.cuda.device(self.device):
            check(fn(self._handle, ptr(xhat), xhat.size(1), xhat.stride(0), b, e, ptr(score), self._stream()))
        return score

    def approx_er_partial(self, projection: torch.Tensor, max_iters=500, rtol=1e-6, reg=1e-6, e_begin=None, e_end=None,
                          return_iters=False):
        """Partial resistance sums over the columns of `projection` (fp64 [m, k], any row stride)."""
        b, e = self._range(e_begin, e_end)
        if projection.dim() != 2 or projection.size(0) != self.num_undirected or projection.dtype != torch.float64:
            raise ValueError("projection must be fp64 [num_undirected, k]")
        if projection.stride(1) != 1:
            projection = projection.contiguous()
        k = projection.size(1)
        out = self._empty(e - b, torch.float64)
        iters = self._empty(k, torch.int32) if return_iters else None
        with torch.cuda.device(self.device):
            check(self._lib.gsp_approx_er_partial(self._handle, ptr(projection), projection.stride(0), k, int(max_iters),
                                                  float(rtol), float(reg), b, e, ptr(out), ptr(iters), self._stream()))
        return (out, iters) if return_iters else out

    def approx_er_partial_philox(self, seed: int, col_begin: int, k: int, k_total: int, max_iters=500, rtol=1e-6, reg=1e-6,
                                 return_iters=False):
        """Partial resistance sums over the projection columns [col_begin, col_begin + k) of the `k_total`-column Philox
        matrix of `seed`, generated inside the projection kernel (no [m, k] matrix in memory)."""
        out = self._empty(self.nnz, torch.float64)
        iters = self._empty(k, torch.int32) if return_iters else None
        with torch.cuda.device(self.device):
            check(self._lib.gsp_approx_er_partial_philox(self._handle, int(seed) & (2 ** 64 - 1), int(col_begin), int(k), int(k_total),
                                                         int(max_iters), float(rtol), float(reg), 0, self.nnz, ptr(out), ptr(iters),
                                                         self._stream()))
        return (out, iters) if return_iters else out

    def philox_projection(self, seed: int, col_begin: int, k: int, k_total: int) -> torch.Tensor:
        """fp64 [num_undirected, k]: the entries `approx_er_partial_philox` draws, written out (tests, inspection)."""
        r = torch.empty((self.num_undirected, k), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            check(self._lib.gsp_philox_projection(int(seed) & (2 ** 64 - 1), self.num_undirected, int(col_begin), int(k), int(k_total),
                                                  ptr(r), self._stream()))
        return r

    def laplacian_solve(self, rhs: torch.Tensor, max_iters=500, rtol=1e-6, reg=1e-6, return_iters=False):
        """X = CG(D - A + reg*I, rhs) column by column (fp64 [n, k]; SciPy cg semantics per column; `gsp_laplacian_solve`)."""
        if rhs.dim() != 2 or rhs.size(0) != self.num_nodes or rhs.dtype != torch.float64:
            raise ValueError("rhs must be fp64 [num_nodes, k]")
        rhs = rhs.to(self.device).contiguous()
        k = rhs.size(1)
        x = torch.empty_like(rhs)
        iters = self._empty(k, torch.int32) if return_iters else None
        with torch.cuda.device(self.device):
            check(self._lib.gsp_laplacian_solve(self._handle, ptr(rhs), k, int(max_iters), float(rtol), float(reg), ptr(x), ptr(iters),
                                                self._stream()))
        return (x, iters) if return_iters else x

    def er_finalize(self, partial: torch.Tensor) -> torch.Tensor:
        with torch.cuda.device(self.device):
            check(self._lib.gsp_er_finalize(ptr(partial), partial.numel(), self._stream()))
        return partial


# ---- selection / compaction on raw tensors -----------------------------------------------------------
def select_mask(scores: torch.Tensor, num_keep: int, keep_lowest: bool = False, exclude: Optional[torch.Tensor] = None,
                out: Optional[torch.Tensor] = None, or_into: bool = False) -> torch.Tensor:
    """uint8 keep-mask of the `num_keep` highest (lowest) scores, ties broken by position (stable-argsort contract)."""
    lib = _lib.load()
    n = scores.numel()
    mask = torch.empty(n, dtype=torch.uint8, device=scores.device) if out is None else out
    with torch.cuda.device(scores.device):
        check(lib.gsp_select_mask(ptr(scores), n, int(num_keep), int(bool(keep_lowest)), ptr(exclude), int(bool(or_into)),
                                  ptr(mask), stream_ptr(scores.device)))
    return mask


def select_mask_sharded(scores: torch.Tensor, num_keep: int, keep_lowest: bool, group, exclude=None, out=None,
                        or_into: bool = False) -> torch.Tensor:
    """Distributed radix select: `scores` is this rank's contiguous slice (rank order == position order).

    Only the 16 KB histograms (all-reduce) and one tie count per rank (all-gather) cross NVLink; the protocol is
    `sharding.distributed_select`."""
    from .sharding import LibgspSelectOps, distributed_select

    return distributed_select(LibgspSelectOps(scores, exclude, out, or_into), num_keep, keep_lowest, group)


def select_compact(scores: torch.Tensor, num_keep: int, keep_lowest: bool, edge_index: torch.Tensor,
                   mask: Optional[torch.Tensor] = None, with_weights: bool = False, invert_weights: bool = False,
                   out: Optional[torch.Tensor] = None):
    """Threshold selection + `edge_index[:, mask]` (+ min-max "-W" weights) in one library call (`gsp_select_compact`).

    `scores` covers the first `scores.numel()` columns of `edge_index` (positional aliasing: canonical position p <->
    column p); `mask` (uint8, >= scores.numel() entries, optional) receives the keep flags of those columns. Returns
    (kept edge_index [2, capacity], weights or None, device int64[1] count); capacity = min(num_keep, n) columns."""
    lib = _lib.load()
    dev = scores.device
    n = scores.numel()
    ei = edge_index if edge_index.is_contiguous() else edge_index.contiguous()
    cap = min(int(num_keep), n)
    if out is None:
        out = torch.empty((2, cap), dtype=torch.int64, device=dev)
    w = torch.empty(out.size(1), dtype=torch.float32, device=dev) if with_weights else None
    count = torch.empty(1, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        check(lib.gsp_select_compact(ptr(scores), n, int(num_keep), int(bool(keep_lowest)), ptr(ei), ei.size(1), ptr(mask),
                                     ptr(out), out.size(1), ptr(w), int(bool(invert_weights)), ptr(count), stream_ptr(dev)))
    return out, w, count


class ShardedSelect:
    """Reusable buffers + the call sequence of the sharded select + compaction (`gsp_select_histogram_slot` ...
    `gsp_select_emit`, include/gsp.h): per method two kinds of collectives — an all-gather of one 16 KB slot per radix pass
    and one of (below, ties) per rank — and no host synchronisation."""

    def __init__(self, device, group):
        import torch.distributed as dist

        self.dev, self.group = device, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.state = torch.empty(_lib.SELECT_STATE_BYTES, dtype=torch.uint8, device=device)
        self.scratch = torch.empty(_lib.SELECT_SCRATCH_BYTES, dtype=torch.uint8, device=device)
        self.slot = torch.empty(_lib.SELECT_SLOT_WORDS, dtype=torch.int64, device=device)
        self.slots = torch.empty(self.world * _lib.SELECT_SLOT_WORDS, dtype=torch.int64, device=device)
        self.totals = torch.empty(2, dtype=torch.int64, device=device)
        self.rank_totals = torch.empty(2 * self.world, dtype=torch.int64, device=device)

    def __call__(self, scores: torch.Tensor, num_keep: int, keep_lowest: bool, edge_index_local: torch.Tensor,
                 mask: Optional[torch.Tensor] = None, with_weights: bool = False, invert_weights: bool = False,
                 out: Optional[torch.Tensor] = None):
        """`scores` / `edge_index_local` [2, n]: this rank's slice. Returns (mask slice, kept columns [2, capacity],
        weights or None, device int64[1] count of kept columns on this rank)."""
        import torch.distributed as dist

        lib = _lib.load()
        n = scores.numel()
        ei = edge_index_local if edge_index_local.is_contiguous() else edge_index_local.contiguous()
        if mask is None:
            mask = torch.empty(n, dtype=torch.uint8, device=self.dev)
        if out is None:
            out = torch.empty((2, n), dtype=torch.int64, device=self.dev)
        w = torch.empty(out.size(1), dtype=torch.float32, device=self.dev) if with_weights else None
        count = torch.empty(1, dtype=torch.int64, device=self.dev)
        with torch.cuda.device(self.dev):
            s = stream_ptr(self.dev)
            check(lib.gsp_select_begin(ptr(self.state), int(num_keep), int(bool(keep_lowest)), s))
            for p in range(_lib.SELECT_PASSES):
                check(lib.gsp_select_histogram_slot(ptr(scores), n, ptr(self.state), p, ptr(self.slot), s))
                dist.all_gather_into_tensor(self.slots, self.slot, group=self.group)
                check(lib.gsp_select_pick_slots(ptr(self.state), ptr(self.slots), self.world, p, s))
            check(lib.gsp_select_tally(ptr(scores), n, ptr(self.state), ptr(self.scratch), ptr(self.totals), s))
            dist.all_gather_into_tensor(self.rank_totals, self.totals, group=self.group)
            check(lib.gsp_select_emit(ptr(scores), n, ptr(self.state), ptr(self.scratch), ptr(self.rank_totals), self.rank,
                                      self.world, ptr(ei), ei.size(1), ptr(mask), ptr(out), out.size(1), ptr(w),
                                      int(bool(invert_weights)), ptr(count), s))
        return mask, out, w, count


def degree_aware_guarantee(src: torch.Tensor, scores: torch.Tensor, num_nodes: int, min_per_node: int):
    """(uint8 mask of guaranteed edges, int64[1] count) — reference core.py:421-435."""
    lib = _lib.load()
    n = src.numel()
    mask = torch.empty(n, dtype=torch.uint8, device=src.device)
    marked = torch.empty(1, dtype=torch.int64, device=src.device)
    with torch.cuda.device(src.device):
        check(lib.gsp_degree_aware_guarantee(ptr(src), ptr(scores), n, int(num_nodes), int(min_per_node), ptr(mask),
                                             ptr(marked), stream_ptr(src.device)))
    return mask, marked


def compact_edges(edge_index: torch.Tensor, mask: torch.Tensor, num_kept: int, scores: Optional[torch.Tensor] = None,
                  with_weights: bool = False, invert_weights: bool = False, out: Optional[torch.Tensor] = None):
    """`edge_index[:, mask]` (+ optional min-max "-W" weights) without leaving the device.

    `num_kept` is the capacity of the output (columns); the true count comes back in the third return value (device
    int64[1]), so a caller that does not know the count can pass an upper bound and avoid a host synchronisation."""
    lib = _lib.load()
    dev = edge_index.device
    e = edge_index.size(1)
    ei = edge_index if edge_index.is_contiguous() else edge_index.contiguous()
    if out is None:
        out = torch.empty((2, int(num_kept)), dtype=torch.int64, device=dev)
    else:
        num_kept = out.size(1)
    w = torch.empty(int(num_kept), dtype=torch.float32, device=dev) if with_weights else None
    count = torch.empty(1, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        check(lib.gsp_compact_edges(ptr(ei), e, e, ptr(mask), ptr(scores) if with_weights else None,
                                    int(bool(invert_weights)), ptr(out), int(num_kept), ptr(w), ptr(count),
                                    stream_ptr(dev)))
    return out, w, count
