"""GCN normalisation + propagation of the kept sub-graph at the bench workload (R-MAT scale S, top-50 % Jaccard edges,
128-d fp32 features): milliseconds and algorithmic GB/s. Usage: python profiles/tools/gcn_probe.py [scale]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import gsr_b200
from gsr_b200 import engine
from gsr_b200.gcn import GcnPropagation, gcn_norm
from gsr_b200.synthetic import rmat_graph_device

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 24
dev = torch.device("cuda", 0)
n, d = 1 << scale, 128
ei = rmat_graph_device(n, 16 * n, scale, seed=5, device=dev)
g = engine.DeviceGraph(ei, n)
scores = g.jaccard()
mask = engine.select_mask(scores, ei.size(1) // 2, False)
kept, w, _ = engine.compact_edges(ei, mask, ei.size(1) // 2, scores=scores, with_weights=True)
del scores, mask, g
x = torch.randn((n, d), dtype=torch.float32, device=dev)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, out


k = kept.size(1)
t_norm, (nei, nw) = timed(lambda: gcn_norm(kept, w, n))
t_plan, prop = timed(lambda: GcnPropagation(nei, nw, n, normalize=False))
out = torch.empty((n, d), dtype=torch.float32, device=dev)
t_prop, _ = timed(lambda: prop(x, out=out), reps=5)
e2 = nei.size(1)
alg = e2 * (4.0 * d + 8 + 8 + 4) + n * (4.0 * d + 16)      # gathered rows + perm/source/weight + output rows + indptr
print(json.dumps({"scale": scale, "kept_edges": k, "edges_with_loops": e2, "dim": d, "gcn_norm_ms": t_norm, "target_order_ms": t_plan,
                  "propagate_ms": t_prop, "propagate_alg_gb": alg / 1e9, "propagate_gbs": alg / t_prop / 1e6,
                  "launches": gsr_b200._lib.load().gsp_launch_count()}))
