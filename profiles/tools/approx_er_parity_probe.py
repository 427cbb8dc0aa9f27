"""ApproxER on a products-SHAPED, down-scaled R-MAT graph (average degree ~50) against the SciPy port of the reference
(oracle/scipy_port.py: the reference's own `scipy.sparse.linalg.cg` calls), 4 of the 64 projection columns of BASELINE
config 4, in the two regimes the benchmark graph family produces:
  capped     500 iterations are not enough (every column stops at the cap, the partial iterate is kept, metrics.py:287-288)
  converged  the same solve with the cap lifted
Prints relative errors and kept-set agreement; tests/test_gpu_parity.py::test_approx_er_products_shaped asserts the bars.
usage: python profiles/tools/approx_er_parity_probe.py [scale] [cap]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import gsr_b200
from gsr_b200.metrics import _approx_er_on_graph
from gsr_b200.synthetic import rmat_graph
from oracle import scipy_port as port

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n = 1 << scale
e = n * 50 // 2 * 2
ei = rmat_graph(n, e, scale, seed=4)
adj = port.build_adjacency(ei, n)
m = e // 2
R = port.projection_matrix(m, 64, 42)[:, :4].copy()      # 4 of the reference's 64 PCG64 columns (scaled by 1/sqrt(64))
sp = gsr_b200.GraphSparsifier(gsr_b200.Data(edge_index=torch.from_numpy(ei), num_nodes=n), "cuda:0")
for cap in ([int(sys.argv[2])] if len(sys.argv) > 2 else [500, 20000]):
    t0 = time.time()
    want, wit = port.approx_effective_resistance(adj, projection=R, max_cg_iters=cap, return_iters=True)
    t1 = time.time()
    got, git = _approx_er_on_graph(sp.graph, projection=R, max_cg_iters=cap, return_iters=True)
    got = got.cpu().numpy(); git = git.cpu().numpy()
    rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-300)
    line = {"scale": scale, "edges": e, "cap": cap, "cpu_s": round(t1 - t0, 1), "iters_cpu": wit.tolist(), "iters_gpu": git.tolist(),
            "rel_max": float(rel.max()), "rel_p999": float(np.quantile(rel, 0.999)), "rel_median": float(np.median(rel))}
    for r in (0.2, 0.5, 0.8):
        k = int(e * r)
        a = np.zeros(e, bool); a[np.argsort(want, kind="stable")[e - k:]] = True
        b = np.zeros(e, bool); b[np.argsort(got, kind="stable")[e - k:]] = True
        line[f"kept_agree_{r}"] = float((a & b).sum() / k)
    print(line, flush=True)
