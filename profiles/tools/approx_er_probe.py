"""Probe: ApproxER (k=64, rtol 1e-6, 500 iters) on the ogbn-products-shaped R-MAT graph (BASELINE config 4)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from gsr_b200 import engine
from gsr_b200.metrics import _approx_er_on_graph
from gsr_b200.synthetic import SHAPES, rmat_graph_device

name = sys.argv[1] if len(sys.argv) > 1 else "products"
k = int(sys.argv[2]) if len(sys.argv) > 2 else 64
n, e, d, scale, seed = SHAPES[name]
dev = torch.device("cuda:0")
t0 = time.time(); ei = rmat_graph_device(n, e, scale, seed, dev); torch.cuda.synchronize(); print("gen", time.time() - t0)
g = engine.DeviceGraph(ei, n)
print("nnz", g.nnz, "m", g.num_undirected, "maxdeg", g.max_degree)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    s, it = _approx_er_on_graph(g, k=k, projection="device", return_iters=True)
    torch.cuda.synchronize(); dt = time.time() - t0
    itc = it.cpu()
    print(f"approx_er k={k}: {dt*1e3:.1f} ms  iters min/mean/max {itc.min().item()}/{itc.float().mean().item():.1f}/{itc.max().item()}  score mean {s.mean().item():.4f}")
