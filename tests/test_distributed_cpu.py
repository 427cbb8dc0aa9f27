"""world_size-2 gloo tests (CPU) of the N > 1 host logic: the distributed radix-select protocol, the variable-length
gathers and the partitioning helpers. The local select operations are NumPy stand-ins with the contract of
include/gsp.h (test infrastructure); on a GPU box the same protocol drives libgsp.so (tests/test_gpu_parity.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gsr_b200 import sharding

SHIFTS = (53, 42, 31, 20, 9, 0)
BITS = (11, 11, 11, 11, 11, 9)


def ordered_keys(scores: np.ndarray) -> np.ndarray:
    s = np.where(scores == 0.0, 0.0, scores)                     # -0.0 ties with +0.0
    b = s.view(np.uint64)
    neg = (b >> np.uint64(63)).astype(bool)
    return np.where(neg, ~b, b | np.uint64(1 << 63))


class NumpySelectOps(sharding.SelectOps):
    def __init__(self, scores: np.ndarray):
        self.scores = scores

    def begin(self, num_keep, keep_lowest):
        k = ordered_keys(self.scores)
        self.keys = k if keep_lowest else ~k
        self.keep_lowest, self.remaining, self.prefix, self.empty = keep_lowest, int(num_keep), np.uint64(0), num_keep <= 0

    def _live(self, p):
        hi = SHIFTS[p] + BITS[p]
        if hi >= 64:
            return np.ones(len(self.keys), bool)
        return (self.keys >> np.uint64(hi)) == (self.prefix >> np.uint64(hi))

    def histogram(self, p):
        digits = ((self.keys[self._live(p)] >> np.uint64(SHIFTS[p])) & np.uint64((1 << BITS[p]) - 1)).astype(np.int64)
        return torch.from_numpy(np.bincount(digits, minlength=self.bins).astype(np.int64))

    def pick(self, p, hist):
        if self.empty:
            return
        h = hist.numpy()
        want = min(self.remaining, int(h.sum()))
        csum = np.cumsum(h)
        d = int(np.searchsorted(csum, want, side="left"))
        self.prefix |= np.uint64(d) << np.uint64(SHIFTS[p])
        self.remaining = want - int(csum[d] - h[d])

    def count_ties(self):
        return torch.tensor([0 if self.empty else int((self.keys == self.prefix).sum())], dtype=torch.int64)

    def write_mask(self, before, total):
        if self.empty:
            return np.zeros(len(self.keys), bool)
        tie = self.keys == self.prefix
        rank = int(before) + np.cumsum(tie) - 1
        lo, hi = (0, self.remaining) if self.keep_lowest else (int(total) - self.remaining, int(total))
        return (self.keys < self.prefix) | (tie & (rank >= lo) & (rank < hi))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, scores, cut, cases, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = (0, cut) if rank == 0 else (cut, len(scores))
        results = []
        for num_keep, keep_lowest in cases:
            mask = sharding.distributed_select(NumpySelectOps(scores[lo:hi].copy()), num_keep, keep_lowest)
            gathered = sharding.all_gather_variable(torch.from_numpy(mask.astype(np.uint8)))
            results.append(gathered.numpy().astype(bool))
        # variable-length gather of [2, K] edge lists keeps rank order
        local = torch.arange(2 * (3 + 2 * rank)).reshape(2, -1) + 100 * rank
        ei = sharding.all_gather_variable(local, dim=1)
        if rank == 0:
            out.put((results, ei.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["ties", "normal", "signed_zero"])
def test_distributed_select_protocol_world2(kind):
    rng = np.random.default_rng(5)
    n = 5000
    scores = {"ties": rng.integers(0, 5, n).astype(np.float64) / 4.0,
              "normal": rng.standard_normal(n),
              "signed_zero": np.where(rng.random(n) < 0.5, -0.0, 0.0) + (rng.random(n) < 0.1)}[kind]
    cases = [(0, False), (1, True), (n // 3, False), (n // 3, True), (n - 1, False), (n, True), (1234, False)]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, scores, 1777, cases, out)) for r in range(2)]
    for p in procs:
        p.start()
    results, ei = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    order = np.argsort(scores, kind="stable")
    for (num_keep, keep_lowest), got in zip(cases, results):
        want = np.zeros(n, bool)
        want[order[:num_keep] if keep_lowest else order[n - num_keep:]] = True
        assert np.array_equal(got, want), (kind, num_keep, keep_lowest)
    assert ei.shape == (2, 8) and ei[0].tolist() == [0, 1, 2, 100, 101, 102, 103, 104]


def test_partition_helpers():
    cost = torch.tensor([1.0, 1, 1, 1, 10, 1, 1, 1, 1, 1]).cumsum(0)
    cuts = sharding.balanced_cuts(cost, 2)
    assert cuts[0] == 0 and cuts[-1] == 10 and 3 <= cuts[1] <= 5
    assert sharding.balanced_cuts(torch.zeros(0), 4) == [0, 0, 0, 0, 0]
    cuts = sharding.balanced_cuts(torch.ones(7).cumsum(0), 8)
    assert cuts == sorted(cuts) and cuts[-1] == 7
    slices = [sharding.column_slice(64, r, 8) for r in range(8)]
    assert slices[0] == (0, 8) and slices[-1] == (56, 64)
    slices = [sharding.column_slice(10, r, 4) for r in range(4)]
    assert sum(hi - lo for lo, hi in slices) == 10 and all(a[1] == b[0] for a, b in zip(slices, slices[1:]))


def test_owner_deal_balances_an_outlier_and_is_deterministic():
    """`sharding.owner_deal`: the heaviest owners placed greedily on the least-loaded rank, the rest dealt in snake order."""
    g = torch.Generator().manual_seed(4)
    costs = torch.rand(50000, generator=g, dtype=torch.float64) ** 6 * 1e6
    costs[17] = 3.0 * float(costs.max())                  # one owner with three times the work of the runner-up
    costs[100:200] = 0.0                                   # owners with nothing to do
    for world in (2, 3, 8):
        deal = sharding.owner_deal(costs, world)
        assert deal.dtype == torch.uint8 and deal.numel() == costs.numel() and int(deal.max()) == world - 1
        assert sharding.owner_deal(costs.clone(), world).equal(deal)
        loads = torch.stack([costs[deal == r].sum() for r in range(world)])
        assert float(loads.max() - loads.min()) <= 1e-3 * float(loads.mean()), (world, loads.tolist())
        counts = torch.bincount(deal.long(), minlength=world)
        assert int(counts.max() - counts.min()) <= 0.02 * costs.numel()       # and a similar number of owners each
    few = sharding.owner_deal(torch.tensor([5.0, 1.0, 1.0]), 2)               # fewer owners than one snake cycle
    assert few.tolist() == [0, 1, 1]
    assert sharding.owner_deal(torch.zeros(0, dtype=torch.float64), 4).numel() == 0


class _OwnedScoresStandIn:
    """Stand-in for engine.DeviceGraph in the owner-sharded exchange: position p is "owned" by node p % num_nodes, its
    Jaccard score is p + 1 and its Adamic-Adar score 1000 + p (the CUDA calls write exactly the owned positions of
    zero-filled full-length buffers; here NumPy-style indexing does)."""

    def __init__(self, nnz, num_nodes):
        self.nnz, self.num_nodes, self.device = nnz, num_nodes, torch.device("cpu")
        self.pos = torch.arange(nnz)

    def _owned(self, node_begin, node_end):
        owner = self.pos % self.num_nodes
        return (owner >= node_begin) & (owner < node_end)

    def jaccard_owned(self, node_begin, node_end, out, counts=None):
        m = self._owned(node_begin, node_end)
        out[: self.nnz][m] = (self.pos[m] + 1).double()
        return out

    def adamic_adar_owned(self, node_weights, node_begin, node_end, out):
        m = self._owned(node_begin, node_end)
        out[: self.nnz][m] = (self.pos[m] + 1000).double()
        return out

    def jaccard_adamic_adar_owned(self, node_weights, node_begin, node_end, out_jaccard, out_adamic_adar):
        self.jaccard_owned(node_begin, node_end, out_jaccard)
        self.adamic_adar_owned(node_weights, node_begin, node_end, out_adamic_adar)
        return out_jaccard, out_adamic_adar


def _owner_exchange_worker(rank, world, port, nnz, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = _OwnedScoresStandIn(nnz, 10)
        node_range = (0, 4) if rank == 0 else (4, 10)
        length, slices = sharding.equal_slices(nnz, world)
        lo, hi = slices[rank]
        res = {}
        for scratch in (None, torch.empty(2 * length * world, dtype=torch.float64)):    # the bench shares one 2x scratch
            tag = "none" if scratch is None else "shared"
            res[f"j_{tag}"] = sharding.owner_sharded_scores(g, "jaccard", None, node_range, scratch=scratch)[: hi - lo].clone()
            res[f"a_{tag}"] = sharding.owner_sharded_scores(g, "adamic_adar", None, node_range, scratch=scratch)[: hi - lo].clone()
            j, a = sharding.owner_sharded_jaccard_adamic_adar(g, None, node_range, scratch=scratch)
            res[f"fj_{tag}"], res[f"fa_{tag}"] = j[: hi - lo].clone(), a[: hi - lo].clone()
        out.put((rank, lo, hi, {k: v.numpy() for k, v in res.items()}))
    finally:
        dist.destroy_process_group()


def test_owner_sharded_exchange_world2():
    """Reduce-scatter form of the owner-sharded Jaccard / Adamic-Adar exchange (separate and fused passes), with and without
    a caller-provided scratch buffer that is larger than one full vector."""
    nnz = 1001                                   # not divisible by the world size: the last slice is padded
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_owner_exchange_worker, args=(r, 2, port, nnz, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = [out.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pos = np.arange(nnz, dtype=np.float64)
    for rank, lo, hi, res in got:
        for tag in ("none", "shared"):
            assert np.array_equal(res[f"j_{tag}"], pos[lo:hi] + 1) and np.array_equal(res[f"fj_{tag}"], pos[lo:hi] + 1), (rank, tag)
            assert np.array_equal(res[f"a_{tag}"], pos[lo:hi] + 1000) and np.array_equal(res[f"fa_{tag}"], pos[lo:hi] + 1000), (rank, tag)
