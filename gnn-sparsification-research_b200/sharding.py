"""Multi-GPU plumbing: one process per GPU, `torch.distributed` (NCCL over NVLink on the box, gloo in CPU tests).

The path shards naturally:

  * scoring      every rank holds the replicated CSR (+ features). FeatCos / degree score one contiguous range of
                 canonical edges per rank with no communication. Jaccard / Adamic-Adar on a symmetric graph are
                 owner-sharded so no pair is evaluated twice: rank r evaluates the pairs owned by its node range
                 (balanced by `gsp_owner_costs`) and the exchange is fused into the scoring kernel: slices live in
                 NVLink-mapped symmetric memory and every score is stored straight at the rank that owns its
                 position (`PeerScoreSlices`, `owner_sharded_scores_p2p`); fallback: full-length zero-filled buffer +
                 one reduce-scatter (`owner_sharded_scores`); asymmetric graphs: edge ranges;
  * selection    distributed radix select: per pass each rank histograms its slice, the 2048-bin histogram is
                 all-reduced (16 KB), every rank picks the same digit; one all-gather of per-rank tie counts
                 resolves the (score, position) boundary; each rank writes its mask slice;
  * ApproxER     projection columns are split over ranks, the per-edge partial sums are all-reduced;
  * outputs      kept edge_index / mask slices are all-gathered (rank order == position order).

The collective protocol lives here, independent of where the local work runs: `distributed_select` drives any
object with the five local operations of `SelectOps` — `LibgspSelectOps` (CUDA, libgsp.so) on a GPU box, a NumPy
stand-in with the same contract in the world_size-2 gloo tests.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import _lib


# ------------------------------------------------------------------------------------------ partitioning
def balanced_cuts(cost_prefix: torch.Tensor, world: int) -> List[int]:
    """Cut points of `world` contiguous ranges with ~equal cost; `cost_prefix` is the inclusive prefix sum."""
    n = cost_prefix.numel()
    if n == 0:
        return [0] * (world + 1)
    total = cost_prefix[-1].item()
    targets = torch.tensor([total * r / world for r in range(1, world)], dtype=cost_prefix.dtype, device=cost_prefix.device)
    inner = torch.searchsorted(cost_prefix, targets).tolist() if world > 1 else []
    cuts = [0] + [int(c) for c in inner] + [n]
    for i in range(1, len(cuts)):           # monotone even for degenerate costs
        cuts[i] = max(cuts[i], cuts[i - 1])
    return cuts


def balanced_edge_ranges(graph, world: int) -> List[Tuple[int, int]]:
    """Contiguous canonical-edge ranges with ~equal estimated work (min endpoint degree + a constant per edge)."""
    if world == 1:
        return [(0, graph.nnz)]
    indptr, indices, _, rows = graph.export(with_data=False, with_rows=True)
    deg = indptr[1:] - indptr[:-1]
    cost = torch.minimum(deg[rows.long()], deg[indices.long()]).double() + 8.0
    cuts = balanced_cuts(torch.cumsum(cost, 0), world)
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


def equal_slices(count: int, world: int) -> Tuple[int, List[Tuple[int, int]]]:
    """(slice length L, [(lo, hi)] per rank) with L = ceil(count / world): the layout `reduce_scatter_tensor` produces."""
    length = (count + world - 1) // world if world > 0 else count
    return length, [(min(r * length, count), min((r + 1) * length, count)) for r in range(world)]


def owner_node_ranges(graph, world: int) -> List[Tuple[int, int]]:
    """Contiguous node ranges with ~equal intersection work (`gsp_owner_costs`): rank r evaluates the pairs they own."""
    if world == 1:
        return [(0, graph.num_nodes)]
    cuts = balanced_cuts(torch.cumsum(graph.owner_costs(), 0), world)
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


def owner_deal(costs: torch.Tensor, world: int, head: int = 4096) -> torch.Tensor:
    """uint8[n] rank of every owner. Owners are sorted by estimated work (descending, stable); the `head` heaviest are
    placed one by one on the least-loaded rank (longest-processing-time greedy, on the host: an outlier such as node 0 of
    R-MAT scale 24, 3x the work of the runner-up, is compensated by the owners that follow it), the rest are dealt in snake
    order (0 .. N-1, N-1 .. 0, ...). Every rank receives the same mix of hub, medium and small owners, so whatever the cost
    estimate gets wrong about a class is spread evenly instead of landing on the rank whose node range holds the hubs
    (contiguous ranges: 28.3 ms on rank 0 against 23.9-25.6 ms on the other seven, R-MAT scale 24)."""
    import heapq

    n = costs.numel()
    order = torch.argsort(costs, descending=True, stable=True)
    i = torch.arange(n, device=costs.device) % (2 * world)
    dealt = torch.where(i < world, i, 2 * world - 1 - i).to(torch.uint8)
    head = min(head - head % (2 * world), n - n % (2 * world))      # whole snake cycles, so the tail starts at rank 0
    if head > 0:
        top = costs[order[:head]].cpu().tolist()
        loads = [(0.0, r) for r in range(world)]
        placed = []
        for c in top:
            load, r = heapq.heappop(loads)
            placed.append(r)
            heapq.heappush(loads, (load + c, r))
        dealt[:head] = torch.tensor(placed, dtype=torch.uint8).to(costs.device)
    out = torch.empty(n, dtype=torch.uint8, device=costs.device)
    out[order] = dealt
    return out


def install_owner_deal(graph, world: int, rank: int) -> Tuple[int, int]:
    """Deal the owners of `graph` over `world` ranks (identical on every rank: the costs and the stable sort are
    deterministic) and install this rank's share; returns the node range to pass to the `*_owned*` calls (all nodes)."""
    if world > 1:
        if world > 256:
            raise ValueError("owner deal supports at most 256 ranks")
        graph.set_owner_deal(owner_deal(graph.owner_costs(), world), rank)
    return 0, graph.num_nodes


def owner_sharded_scores(graph, metric: str, group, node_range: Tuple[int, int], node_weights=None,
                         scratch: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Jaccard / Adamic-Adar on `world` GPUs without duplicated work: this rank evaluates the undirected pairs owned by
    its node range, writes each score at both directed positions of a zero-filled full-length buffer, and a
    reduce-scatter (sum; every position has exactly one non-zero contributor, so the sum is exact) hands every
    rank its contiguous slice of `equal_slices(nnz, world)`. NCCL over NVLink; fp64 [nnz] crosses the fabric once."""
    world = dist.get_world_size(group)
    length, _ = equal_slices(graph.nnz, world)
    full = (scratch[: length * world] if scratch is not None and scratch.numel() >= length * world
            else torch.empty(length * world, dtype=torch.float64, device=graph.device))
    full.zero_()
    if metric == "jaccard":
        graph.jaccard_owned(node_range[0], node_range[1], full)
    elif metric == "adamic_adar":
        graph.adamic_adar_owned(node_weights, node_range[0], node_range[1], full)
    else:
        raise ValueError(metric)
    out = torch.empty(length, dtype=torch.float64, device=graph.device)
    dist.reduce_scatter_tensor(out, full, group=group)
    return out


def owner_sharded_jaccard_adamic_adar(graph, group, node_range: Tuple[int, int], node_weights=None,
                                      scratch: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Both neighbourhood scores from one streaming pass per rank (`gsp_jaccard_adamic_adar_owned`), then one
    reduce-scatter of the two zero-filled full-length vectors laid out as [world, 2, length] — returns this rank's
    (jaccard, adamic_adar) slices."""
    world = dist.get_world_size(group)
    length, _ = equal_slices(graph.nnz, world)
    need = 2 * length * world
    both = scratch if scratch is not None and scratch.numel() >= need else torch.empty(need, dtype=torch.float64, device=graph.device)
    both = both[:need]
    both.zero_()
    jac_full, aa_full = both[:length * world], both[length * world:]
    graph.jaccard_adamic_adar_owned(node_weights, node_range[0], node_range[1], jac_full, aa_full)
    if world == 1:
        return jac_full[:length], aa_full[:length]
    # rank k's output chunk = [jaccard slice k | adamic-adar slice k]
    send = torch.stack((jac_full.view(world, length), aa_full.view(world, length)), dim=1).contiguous()
    out = torch.empty(2 * length, dtype=torch.float64, device=graph.device)
    dist.reduce_scatter_tensor(out, send.view(-1), group=group)
    return out[:length], out[length:]


class PeerScoreSlices:
    """Per-rank fp64 score slices in NVLink-mapped symmetric memory (torch.distributed._symmetric_memory).

    Every rank allocates its slice of `equal_slices(nnz, world)` positions, the rendezvous maps all slices into every
    process, and the scoring kernel (`gsp_*_owned_scatter`) stores each score directly at its owner — the exchange is
    fused into the compute kernel as plain peer stores, one crossing of the fabric per score (the reduce-scatter path moves
    the full zero-padded vector from every rank). Raises if symmetric memory is unavailable; callers fall back to
    `owner_sharded_scores`."""

    def __init__(self, nnz: int, group, device):
        import torch.distributed._symmetric_memory as symm_mem

        self.group = group
        self.world = dist.get_world_size(group)
        self.length, _ = equal_slices(nnz, self.world)
        self.tensor = symm_mem.empty(self.length, dtype=torch.float64, device=device)
        self.handle = symm_mem.rendezvous(self.tensor, group)
        self.slices_dev_ptr = int(self.handle.buffer_ptrs_dev)

    def barrier(self):
        self.handle.barrier()


def owner_sharded_scores_p2p(graph, metric: str, peer: PeerScoreSlices, node_range: Tuple[int, int], node_weights=None):
    """Owner-sharded Jaccard / Adamic-Adar with the exchange fused into the scoring kernel (peer stores over NVLink)."""
    peer.barrier()                      # every rank is done reading the previous contents of its slice
    graph.owned_scatter(metric, node_range[0], node_range[1], peer.slices_dev_ptr, peer.world, peer.length, node_weights)
    peer.barrier()                      # all peers' stores into this rank's slice have landed
    return peer.tensor


def owner_sharded_jaccard_adamic_adar_p2p(graph, peer_jaccard: PeerScoreSlices, peer_adamic_adar: PeerScoreSlices,
                                          node_range: Tuple[int, int], node_weights=None, kernel_events=None):
    """Fused pass + fused exchange: one streaming kernel per rank stores both scores of every pair it owns straight into
    the owning ranks' slices (two symmetric-memory buffers). `kernel_events`: an optional pair of CUDA events recorded
    around this rank's kernels alone (between the barriers), for the per-rank load-balance figure of the benchmark."""
    peer_jaccard.barrier()
    peer_adamic_adar.barrier()
    if kernel_events is not None:
        kernel_events[0].record()
    graph.owned_scatter("jaccard+adamic_adar", node_range[0], node_range[1], peer_adamic_adar.slices_dev_ptr, peer_adamic_adar.world,
                        peer_adamic_adar.length, node_weights, jaccard_slices_dev_ptr=peer_jaccard.slices_dev_ptr)
    if kernel_events is not None:
        kernel_events[1].record()
    peer_jaccard.barrier()
    peer_adamic_adar.barrier()
    return peer_jaccard.tensor, peer_adamic_adar.tensor


def column_slice(k: int, rank: int, world: int) -> Tuple[int, int]:
    """Projection columns [lo, hi) of rank `rank` (ApproxER column sharding)."""
    return (k * rank) // world, (k * (rank + 1)) // world


# ------------------------------------------------------------------------------------------ distributed select
class SelectOps:
    """Local operations of the radix select on this rank's slice (see include/gsp.h, selection section)."""

    passes = _lib.SELECT_PASSES
    bins = _lib.SELECT_BINS

    def begin(self, num_keep: int, keep_lowest: bool) -> None: ...
    def histogram(self, pass_index: int) -> torch.Tensor: ...          # int64[bins] on the collective's device
    def pick(self, pass_index: int, hist: torch.Tensor) -> None: ...
    def count_ties(self) -> torch.Tensor: ...                           # int64[1]
    def write_mask(self, ties_before: torch.Tensor, ties_total: torch.Tensor): ...


def distributed_select(ops: SelectOps, num_keep: int, keep_lowest: bool, group=None):
    """Run the select protocol; returns whatever `ops.write_mask` returns (this rank's mask slice)."""
    world = dist.get_world_size(group) if group is not None or dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    ops.begin(num_keep, keep_lowest)
    for p in range(ops.passes):
        hist = ops.histogram(p)
        if world > 1:
            dist.all_reduce(hist, group=group)
        ops.pick(p, hist)
    ties = ops.count_ties()
    if world > 1:
        all_ties = [torch.empty_like(ties) for _ in range(world)]
        dist.all_gather(all_ties, ties, group=group)
        all_ties = torch.cat(all_ties)
        before = all_ties[:rank].sum().reshape(1)
        total = all_ties.sum().reshape(1)
    else:
        before, total = torch.zeros_like(ties), ties
    return ops.write_mask(before, total)


class LibgspSelectOps(SelectOps):
    """The CUDA implementation: every step is one libgsp.so call on torch's current stream, no host sync."""

    def __init__(self, scores: torch.Tensor, exclude: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                 or_into: bool = False):
        self.lib = _lib.load()
        self.scores, self.exclude, self.or_into = scores, exclude, or_into
        self.dev = scores.device
        self.n = scores.numel()
        self.mask = torch.empty(self.n, dtype=torch.uint8, device=self.dev) if out is None else out
        self.state = torch.empty(_lib.SELECT_STATE_BYTES, dtype=torch.uint8, device=self.dev)
        self.hist = torch.empty(_lib.SELECT_BINS, dtype=torch.int64, device=self.dev)
        self.ties = torch.zeros(1, dtype=torch.int64, device=self.dev)

    def _s(self):
        return _lib.stream_ptr(self.dev)

    def begin(self, num_keep, keep_lowest):
        with torch.cuda.device(self.dev):
            _lib.check(self.lib.gsp_select_begin(_lib.ptr(self.state), int(num_keep), int(bool(keep_lowest)), self._s()))

    def histogram(self, p):
        with torch.cuda.device(self.dev):
            _lib.check(self.lib.gsp_select_histogram(_lib.ptr(self.scores), self.n, _lib.ptr(self.exclude),
                                                     _lib.ptr(self.state), p, _lib.ptr(self.hist), self._s()))
        return self.hist

    def pick(self, p, hist):
        with torch.cuda.device(self.dev):
            _lib.check(self.lib.gsp_select_pick(_lib.ptr(self.state), _lib.ptr(hist), p, self._s()))

    def count_ties(self):
        with torch.cuda.device(self.dev):
            _lib.check(self.lib.gsp_select_count_ties(_lib.ptr(self.scores), self.n, _lib.ptr(self.exclude),
                                                      _lib.ptr(self.state), _lib.ptr(self.ties), self._s()))
        return self.ties

    def write_mask(self, ties_before, ties_total):
        with torch.cuda.device(self.dev):
            _lib.check(self.lib.gsp_select_write_mask(_lib.ptr(self.scores), self.n, _lib.ptr(self.exclude),
                                                      _lib.ptr(self.state), _lib.ptr(ties_before.contiguous()),
                                                      _lib.ptr(ties_total.contiguous()), int(self.or_into),
                                                      _lib.ptr(self.mask), self._s()))
        return self.mask


# ------------------------------------------------------------------------------------------ gathers
def all_gather_variable(local: torch.Tensor, group=None, dim: int = -1) -> torch.Tensor:
    """Concatenate per-rank tensors of different length along `dim` in rank order (== canonical position order)."""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    dim = dim % local.dim()
    count = torch.tensor([local.size(dim)], dtype=torch.int64, device=local.device)
    counts = [torch.empty_like(count) for _ in range(world)]
    dist.all_gather(counts, count, group=group)
    sizes = [int(c.item()) for c in counts]
    width = max(sizes)
    moved = local.movedim(dim, 0).contiguous()
    padded = torch.zeros((width,) + tuple(moved.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: moved.size(0)] = moved
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    out = torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)
    return out.movedim(0, dim)


def sharded_threshold_sparsify(graph, edge_index: torch.Tensor, local_scores: torch.Tensor, e_range: Tuple[int, int],
                               num_keep: int, keep_lowest: bool, group=None, with_weights: bool = False):
    """Global top-/bottom-`num_keep` over score slices held by the ranks of `group`.

    Returns (kept edge_index [2, K] gathered on every rank, local uint8 mask slice). `with_weights` is not
    offered here: the "-W" min-max needs global extrema (two more scalar all-reduces) — see DESIGN.md."""
    from .engine import compact_edges

    ops = LibgspSelectOps(local_scores)
    mask = distributed_select(ops, num_keep, keep_lowest, group)
    kept_local = int(mask.sum().item())
    lo, hi = e_range
    local_ei = edge_index[:, lo:hi].contiguous()
    kept, _, _ = compact_edges(local_ei, mask, kept_local)
    return all_gather_variable(kept, group, dim=1), mask
