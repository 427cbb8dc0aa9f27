"""ms per CG iteration of ApproxER on the products-shaped graph (BASELINE config 4, k = 64, generated projection): two
iteration caps, the difference divided by the extra iterations. (Round 2 used it with a GSP_SPMM_VARIANT knob to compare
gather depths / occupancies of the SpMM: 4-deep x 4 CTAs 9.8-10.2 ms, 8-deep x 3 CTAs 11.2 ms, 4-deep x 5 CTAs 11.0 ms.)
usage: python profiles/tools/approx_er_iter_probe.py [iters]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from gsr_b200 import engine
from gsr_b200.metrics import _approx_er_on_graph
from gsr_b200.synthetic import SHAPES, rmat_graph_device

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 40
n, e, d, scale, seed = SHAPES["products"]
dev = torch.device("cuda:0")
ei = rmat_graph_device(n, e, scale, seed, dev)
g = engine.DeviceGraph(ei, n)
_approx_er_on_graph(g, k=8, max_cg_iters=3, projection="device")
for variant in ("default", "default"):
    out = []
    for cap in (iters, 2 * iters):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        _approx_er_on_graph(g, k=64, max_cg_iters=cap, projection="device")
        torch.cuda.synchronize(); out.append((time.perf_counter() - t0) * 1e3)
    print(f"variant {variant}: {out[0]:.1f} ms for {iters} iterations, {out[1]:.1f} ms for {2*iters}: {(out[1]-out[0])/iters:.3f} ms per CG iteration, setup+projection {2*out[0]-out[1]:.1f} ms", flush=True)
