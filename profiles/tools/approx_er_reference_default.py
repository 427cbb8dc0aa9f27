"""ApproxER with the REFERENCE's defaults (epsilon 0.3 -> k = 24 ln n / eps^2, NumPy PCG64 projection, rtol 1e-6, 500 its)
on the Roman-empire-shaped graph (BASELINE config 2): the case the survey timed at 183 s on the CPU reference."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import gsr_b200
from gsr_b200.synthetic import named_graph

ei, x, n = named_graph("roman_empire")
data = gsr_b200.Data(edge_index=torch.from_numpy(ei), x=torch.from_numpy(x), num_nodes=n)
for rep in range(2):
    sp = gsr_b200.GraphSparsifier(data, "cuda:0")
    torch.cuda.synchronize(); t0 = time.time()
    s = sp.compute_scores("approx_er")
    t1 = time.time()
    out = sp.sparsify("approx_er", 0.5)
    torch.cuda.synchronize(); t2 = time.time()
    print(f"roman-empire shape: n={n} E={ei.shape[1]} k={gsr_b200.metrics.jl_dimension(n, 0.3)}: compute_scores {1e3*(t1-t0):.0f} ms, sparsify {1e3*(t2-t1):.1f} ms, mean score {s.mean():.4f}")
