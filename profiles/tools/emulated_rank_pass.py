"""One GPU standing in for rank r of an N-rank owner deal (R-MAT scale 24): the fused Jaccard + Adamic-Adar pass of that
rank's owners with (a) plain local stores (`*_owned`), (b) the peer-scatter code path with every "peer" slice on this GPU.
Separates what an owner-sharded rank loses to its shorter work lists (tails of the three launches) from what the
NVLink stores cost: compare with rank_spread of the multi-GPU bench lines. Usage: python emulated_rank_pass.py [N] [ranks...]"""
import sys

import torch

import gsr_b200  # noqa: F401
from gsr_b200 import engine, sharding
from gsr_b200.synthetic import rmat_graph_device

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ranks = [int(a) for a in sys.argv[2:]] or [0, world // 2, world - 1]
dev = torch.device("cuda:0")
scale = 24
n, e = 1 << scale, (1 << scale) * 16
ei = rmat_graph_device(n, e, scale, seed=5, device=dev)
g = engine.DeviceGraph(ei, n)
w = g.aa_node_weights_numpy()
deal = sharding.owner_deal(g.owner_costs(), world)
length, _ = sharding.equal_slices(g.nnz, world)
full_j = torch.zeros(length * world, dtype=torch.float64, device=dev)
full_a = torch.zeros_like(full_j)
ptr_j = torch.tensor([full_j.data_ptr() + 8 * length * r for r in range(world)], dtype=torch.int64, device=dev)
ptr_a = torch.tensor([full_a.data_ptr() + 8 * length * r for r in range(world)], dtype=torch.int64, device=dev)


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


t_full = timed(lambda: g.jaccard_adamic_adar(w, out_jaccard=full_j[:g.nnz], out_adamic_adar=full_a[:g.nnz]))
print(f"single GPU, all owners: {t_full:.2f} ms  (1/{world} = {t_full / world:.2f} ms)")
for r in ranks:
    g.set_owner_deal(deal, r)
    t_local = timed(lambda: g.jaccard_adamic_adar_owned(w, 0, n, full_j, full_a))
    t_scatter = timed(lambda: g.owned_scatter("jaccard+adamic_adar", 0, n, ptr_a.data_ptr(), world, length, w, ptr_j.data_ptr()))
    print(f"rank {r}/{world}: owned (local stores) {t_local:.2f} ms, scatter path (local 'peers') {t_scatter:.2f} ms")
