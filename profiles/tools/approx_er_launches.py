"""Short ApproxER run (20 CG iterations) for ncu launch lists: products-shaped graph, k columns."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from gsr_b200 import engine
from gsr_b200.metrics import _approx_er_on_graph
from gsr_b200.synthetic import SHAPES, rmat_graph_device

k = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n, e, d, scale, seed = SHAPES["products"]
dev = torch.device("cuda:0")
g = engine.DeviceGraph(rmat_graph_device(n, e, scale, seed, dev), n)
_approx_er_on_graph(g, k=k, max_cg_iters=20, projection="device")
torch.cuda.synchronize()
print("done")
