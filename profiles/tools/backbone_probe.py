"""Metric-backbone sparsification (Jaccard costs) on the Cora- and Roman-empire-shaped graphs; the reference runs
NetworkX all-pairs Dijkstra here."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import gsr_b200
from gsr_b200.synthetic import named_graph

for shape in ("cora", "roman_empire"):
    ei, x, n = named_graph(shape, with_features=False)
    sp = gsr_b200.GraphSparsifier(gsr_b200.Data(edge_index=torch.from_numpy(ei), num_nodes=n), "cuda:0")
    sp.compute_scores("jaccard")
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        out, stats = sp.sparsify_metric_backbone("jaccard")
        torch.cuda.synchronize(); dt = time.time() - t0
    print(f"{shape}: n={n} E={ei.shape[1]} backbone {dt*1e3:.1f} ms, kept {stats['retained_edges']} ({stats['retention_ratio']:.3f}), sweeps {stats['relaxation_sweeps']}")
