#!/usr/bin/env python
"""Benchmark of the edge-scoring sparsification hot path (BASELINE.json metric: edges scored/sec per method +
ApproxER sparsify ms at 1/2/4/8 B200 vs host CPU).

    python bench.py --gpus N --steps K --warmup W                 (N>1: launched under torchrun, one rank per GPU)
    python bench.py --impl reference ...                           (the reference's CPU path on the box's host cores)

A *step* is one pass of the hot path over the whole synthetic graph (BASELINE config 5: R-MAT scale 24, 268 M directed
edges, 128-d fp32 features): Jaccard, Adamic-Adar and feature-cosine scoring of every directed edge, each followed by
global top-50 % selection and edge_index compaction. Units per step = 3 x E edge scores. Jaccard and Adamic-Adar walk
the same neighbour lists, so the step takes both from ONE streaming pass (`gsp_jaccard_adamic_adar`, bit-identical to the
separate calls; GSP_BENCH_FUSED=0 times two passes); selection + mask + compaction is one library call per method
(`gsp_select_compact`).

  value    device-timed: graph CSR + features already resident in HBM, CUDA events around the K steps, max over ranks.
  e2e      the same step through the reference-facing API with HOST inputs, all copies inside the timed region:
           N = 1  data_host.to(cuda) -> GraphSparsifier(data, cuda) -> prefetch_scores -> compute_scores / sparsify
           N > 1  sharded_to_device(data_host, cuda, group) -> GraphSparsifier(data, cuda, group=group) -> the same calls;
                  every rank uploads 1/N of the inputs (replicas assembled over NVLink) and reads back ITS slice of the
                  score vectors and masks (`sp.local_range`), the kept edge_index is all-gathered on every device.
  roofline the dominant kernel by ALGORITHMIC bytes / CUDA-event time / measured copy peak, with the SURVEY 8d formula,
           the ncu DRAM traffic, every per-method line, the other BASELINE configs (1-4) and the per-rank spread beside it.

N > 1 (device-timed): FeatCos scores a contiguous edge slice per rank; Jaccard / Adamic-Adar are owner-sharded (each
undirected pair on exactly one rank; owners dealt to the ranks in cost order, GSP_BENCH_PARTITION=ranges: contiguous node
ranges) with every score stored by the scoring kernel straight into the slice of the rank that owns its position (NVLink
symmetric memory; NCCL reduce-scatter as the fallback); selection all-gathers 16 KB radix slots and two counters per rank,
mask + compaction are one emit per rank (`engine.ShardedSelect`). Fixed graph => "scaling": "strong".
Clocks and throttle reasons are sampled during the timed region through NVML (GSP_BENCH_CLOCKS=smi: the nvidia-smi CLI).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METHODS = ("jaccard", "adamic_adar", "feature_cosine")
RETENTION = 0.5
BOTH = "jaccard+adamic_adar"
METRIC = "edges scored/sec (Jaccard+AA+FeatCos scoring + top-k select)"
DTYPE = "int32 indices / f64 scores / f32 features"


def measured_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line).

    Default: NVML in this process (the interface nvidia-smi itself reads), one light query every 10 ms from a thread set
    up before the timed region starts — ~60 samples over a 2-GPU run where the CLI poll (`GSP_BENCH_CLOCKS=smi`,
    `-lms 100`) returns 4-7. Neither perturbs the measurement (A/B: `profiles/tools/ab_clock_sampler.sh`)."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None
        self.thread = None
        self.samples = []
        self.mode = os.environ.get("GSP_BENCH_CLOCKS", "nvml")

    # -- NVML thread ---------------------------------------------------------------------------------
    def _nvml_handle(self):
        import pynvml

        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(self.gpu_index).uuid)
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
        except Exception:
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.gpu_index)

    def _nvml_loop(self, nv, handle):
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)
                reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(handle)
                self.samples.append((float(sm), int(reasons)))
            except Exception:
                pass
            self._stop.wait(0.01)

    def start(self):
        if self.mode == "nvml":
            try:
                import threading

                nv, handle = self._nvml_handle()
                self._nv, self._handle = nv, handle
                self._max = float(nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM))
                self._stop = threading.Event()
                self.thread = threading.Thread(target=self._nvml_loop, args=(nv, handle), daemon=True)
                self.thread.start()
                return
            except Exception:
                self.thread = None
                self.mode = "smi"
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def mark(self):
        """Called when the timed region starts: earlier samples are dropped."""
        self._mark = len(self.samples)

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": self.mode}
        if self.thread is not None:
            self._stop.set()
            self.thread.join(timeout=2)
            nv = self._nv
            taken = self.samples[getattr(self, "_mark", 0):] or self.samples
            names = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                     ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap))
            if taken:
                out.update(sm_mhz=float(np.median([t[0] for t in taken])), sm_max_mhz=self._max, samples=len(taken),
                           reasons=sorted({name for name, bit in names for t in taken if t[1] & bit}))
            return out
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons, mx = [], set(), None
        try:
            for line in open(self.path):
                f = [t.strip() for t in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx = float(f[2])
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))
        return out


def cpu_sample(total_steps: int) -> dict:
    """Bounded CPU sample of the workload: one step of the SciPy path costs ~9 s at R-MAT scale 15 and ~30 s at scale 16
    (SpGEMM fill-in grows faster than the edge count), so the scale follows the number of steps the caller will time."""
    scale = int(os.environ.get("GSP_BENCH_CPU_SCALE", "16" if total_steps <= 4 else "15"))
    return dict(scale=scale, num_nodes=1 << scale, edges=(1 << scale) * 16, dim=128, seed=5)


def cpu_sample_description(c: dict) -> str:
    return (f"R-MAT scale {c['scale']} ({c['num_nodes']} nodes, {c['edges']} directed edges, {c['dim']}-d fp32), same generator "
            f"family as the GPU workload; one full step (Jaccard+AA+FeatCos scoring via SciPy SpGEMM/NumPy + argsort top-50%)")


def cpu_step_fn(c: dict):
    """(callable running one step on the CPU, kind): the UNMODIFIED reference (`GraphSparsifier.compute_scores` +
    `sparsify`, /root/reference through oracle/ref_loader.py) where it exists — the build container — else the oracle's
    SciPy/NumPy port of the same calls (the GPU box has no /root/reference)."""
    from gsr_b200.synthetic import features, rmat_graph

    ei = rmat_graph(c["num_nodes"], c["edges"], c["scale"], c["seed"])
    x = features(c["num_nodes"], c["dim"], c["seed"])
    try:
        from oracle import ref_loader
        if ref_loader.available():
            ref = ref_loader.load(stable=False)
            import gsr_b200
            data = gsr_b200.Data(edge_index=torch.from_numpy(ei), x=torch.from_numpy(x), num_nodes=c["num_nodes"])

            def step_ref():
                sp = ref.GraphSparsifier(data, "cpu")
                for m in METHODS:
                    sp.compute_scores(m)
                    sp.sparsify(m, RETENTION, return_mask=True)
            return step_ref, "reference"
    except Exception:
        pass
    from oracle import scipy_port as port
    adj = port.build_adjacency(ei, c["num_nodes"])

    def step_port():
        for m in METHODS:
            s = port.jaccard(adj) if m == "jaccard" else port.adamic_adar(adj) if m == "adamic_adar" else port.feature_cosine(adj, x)
            port.threshold_mask(s, c["edges"], RETENTION, stable=False)      # reference default argsort kind
    return step_port, "port"


def run_reference_arm(args) -> None:
    """`--impl reference`: the reference's CPU implementation on a bounded sample of the workload. SciPy's SpGEMM, the
    NumPy gathers and argsort of this path are single-threaded: cores = 1 (the box's core count is reported beside it)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    c = cpu_sample(args.steps + args.warmup)
    step, kind = cpu_step_fn(c)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = len(METHODS) * c["edges"] * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "edges/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": {"workload": workload_name(args), "cpu_sample": cpu_sample_description(c)},
        "cpu_baseline": {"value": value, "unit": "edges/s", "cores": 1, "host_cores_available": len(os.sched_getaffinity(0)),
                         "kind": kind, "sample": cpu_sample_description(c)},
        "e2e": {"value": value, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm: helpers
def workload_name(args) -> str:
    return (f"rmat scale {args.scale}: {1 << args.scale} nodes, {(1 << args.scale) * args.edge_factor} directed edges, "
            f"{args.dim}-d fp32 features; Jaccard/AA/FeatCos scoring + top-{int(RETENTION * 100)}% select + compaction")


def _timed(fn, dev, repeat=3):
    fn()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(repeat):
        fn()
    torch.cuda.synchronize(dev)
    return (time.perf_counter() - t0) / repeat * 1e3


def run_small_configs(dev):
    """BASELINE configs 1 and 2 through the public API (wall-clock ms per call, scores cached for the selection lines):
    C1 Jaccard-T keep 50 % on the Cora-shaped graph; C2 every threshold method (-T, -IT, -W) on the Roman-empire-shaped
    graph, ApproxER with the reference's own defaults (eps = 0.3 -> k = 2 674 columns, host PCG64 projection)."""
    import gsr_b200
    from gsr_b200.synthetic import named_graph

    out = {}
    for name, methods in (("cora", ("jaccard",)), ("roman_empire", ("jaccard", "adamic_adar", "feature_cosine", "approx_er"))):
        ei, x, n = named_graph(name)
        data = gsr_b200.Data(edge_index=torch.from_numpy(ei), x=torch.from_numpy(x), num_nodes=n).to(dev)
        row = {"nodes": n, "directed_edges": int(ei.shape[1]), "feature_dim": int(x.shape[1])}
        for m in methods:
            def score():
                sp = gsr_b200.GraphSparsifier(data, str(dev))
                sp.compute_scores(m)
                return sp
            reps = 1 if m == "approx_er" else 3
            ms = _timed(score, dev, reps)
            sp = score()
            row[m] = {"compute_scores_ms": ms, "edges_per_s": ei.shape[1] / (ms * 1e-3),
                      "sparsify_T_ms": _timed(lambda: sp.sparsify(m, RETENTION, return_mask=True), dev),
                      "sparsify_IT_ms": _timed(lambda: sp.sparsify(m, RETENTION, return_mask=True, keep_lowest=True), dev),
                      "sparsify_T_W_ms": _timed(lambda: sp.sparsify_with_weights(m, RETENTION), dev)}
        out[name] = row
    return out


def run_variants(dev):
    """Degree-aware and sampled Jaccard sparsification through the public API on the ogbn-arxiv-shaped graph (BASELINE
    config 3; the reference's own degree-aware loops are O(N*E) + O(E^2): 2.7 s at the Roman-empire shape, ~45 min here)."""
    import gsr_b200
    from gsr_b200.synthetic import SHAPES, rmat_graph_device

    n3, e3, _, scale3, seed3 = SHAPES["arxiv"]
    ei3 = rmat_graph_device(n3, e3, scale3, seed3, dev)
    sp = gsr_b200.GraphSparsifier(gsr_b200.Data(edge_index=ei3, num_nodes=n3), str(dev))
    sp.compute_scores("jaccard")
    out = {"workload": f"arxiv-shaped R-MAT: {n3} nodes, {e3} directed edges; Jaccard scores cached, retention 0.5"}
    for name, fn in (("threshold_ms", lambda: sp.sparsify("jaccard", 0.5, return_mask=True)),
                     ("degree_aware_ms", lambda: sp.sparsify_degree_aware("jaccard", 0.5, return_mask=True)),
                     ("sampled_numpy_choice_ms", lambda: sp.sparsify_sampled("jaccard", 0.5, return_mask=True)),
                     ("sampled_device_ms", lambda: sp.sparsify_sampled("jaccard", 0.5, return_mask=True, method="device"))):
        out[name] = _timed(fn, dev)
    return out


def run_approx_er(args, dev, rank, world, group):
    """ApproxER-T sparsify (scores + top-50 % select + compaction) on the products-shaped R-MAT graph; projection columns
    are split over the ranks and the per-edge partial sums all-reduced (NCCL). Device Philox projection (throughput
    mode; parity mode feeds the reference's NumPy matrix, tests/test_gpu_parity.py)."""
    from gsr_b200 import engine
    from gsr_b200.metrics import _approx_er_on_graph
    from gsr_b200.synthetic import SHAPES, rmat_graph_device

    torch.cuda.empty_cache()
    n4, e4, _, scale4, seed4 = SHAPES[args.er_shape]
    ei4 = rmat_graph_device(n4, e4, scale4, seed4, dev)
    g4 = engine.DeviceGraph(ei4, n4)
    k = 64
    _approx_er_on_graph(g4, k=8, max_cg_iters=3, projection="device", group=group)      # builds the lazy side structures
    if world > 1:
        torch.distributed.barrier(group)
    torch.cuda.synchronize(dev)
    t0, t1, t2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    t0.record()
    scores, iters = _approx_er_on_graph(g4, k=k, max_cg_iters=500, cg_tol=1e-6, projection="device", group=group,
                                        return_iters=True)
    t1.record()
    keep = int(e4 * RETENTION)
    engine.select_compact(scores, keep, False, ei4)
    t2.record()
    torch.cuda.synchronize(dev)
    score_ms, total_ms = t0.elapsed_time(t1), t0.elapsed_time(t2)
    tms = torch.tensor([score_ms, total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(tms, op=torch.distributed.ReduceOp.MAX, group=group)
    it = iters.float()
    k_loc = (k + world - 1) // world
    iters_run = int(iters.max().item())
    per_iter_ms = float(tms[0]) / max(iters_run, 1)
    bytes_reuse = 4.0 * g4.nnz + 8.0 * k_loc * 10 * n4
    bytes_noreuse = 4.0 * g4.nnz + 8.0 * k_loc * (g4.nnz + 9 * n4)
    peak, _ = measured_peak()
    return {"workload": f"{args.er_shape}-shaped R-MAT: {n4} nodes, {e4} directed edges; JLT k={k}, CG rtol 1e-6, <=500 iterations, reg 1e-6",
            "score_ms": float(tms[0]), "sparsify_ms": float(tms[1]), "edges_per_s": e4 / (float(tms[1]) * 1e-3),
            "cg_iterations": {"min": int(it.min()), "mean": float(it.mean()), "max": iters_run}, "columns_per_gpu": k_loc,
            "ms_per_cg_iteration": per_iter_ms,
            "roofline_per_iteration": {"alg_gb_perfect_reuse": bytes_reuse / 1e9, "alg_gb_no_reuse": bytes_noreuse / 1e9,
                                       "frac_perfect_reuse": bytes_reuse / (per_iter_ms * 1e-3) / 1e9 / peak,
                                       "frac_no_reuse": bytes_noreuse / (per_iter_ms * 1e-3) / 1e9 / peak},
            "max_degree": g4.max_degree, "projection": "device Philox normals / sqrt(k)"}


# ------------------------------------------------------------------------------------------ GPU arm
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=int, default=24, help="R-MAT scale (BASELINE config 5: 24)")
    ap.add_argument("--edge-factor", type=int, default=16, help="directed edges per node (16 -> 268 M at scale 24)")
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-approx-er", action="store_true", help="skip the ApproxER sparsify timing (BASELINE config 4)")
    ap.add_argument("--no-small-configs", action="store_true", help="skip the BASELINE config 1-3 latency lines")
    ap.add_argument("--er-shape", default="products")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = max(args.warmup, 1)   # the driver passes W; timing rules ask for >= 3 in reported runs

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import gsr_b200
    from gsr_b200 import _lib, engine, sharding
    from gsr_b200.sharded_sparsifier import sharded_to_device
    from gsr_b200.synthetic import rmat_graph_device

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    group = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's version / debug lines stay off stdout (one JSON line)
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    lib = _lib.load()

    n = 1 << args.scale
    e = n * args.edge_factor
    ei = rmat_graph_device(n, e, args.scale, seed=5, device=dev)
    gen = torch.Generator(device=dev); gen.manual_seed(1005)
    x = torch.randn((n, args.dim), dtype=torch.float32, device=dev, generator=gen)
    graph = engine.DeviceGraph(ei, n)
    assert graph.nnz == e and graph.symmetric and graph.input_canonical
    slice_len, slices = sharding.equal_slices(e, world)     # every rank's contiguous slice of canonical positions
    e_lo, e_hi = slices[rank]
    local = e_hi - e_lo
    # owners dealt in cost order (GSP_BENCH_PARTITION=ranges: contiguous node ranges, for A/B)
    if os.environ.get("GSP_BENCH_PARTITION", "deal") == "ranges":
        node_range = sharding.owner_node_ranges(graph, world)[rank]
    else:
        node_range = sharding.install_owner_deal(graph, world, rank)
    fused = os.environ.get("GSP_BENCH_FUSED", "1") != "0"
    full_scratch = torch.empty(slice_len * world * (2 if fused else 1), dtype=torch.float64, device=dev) if world > 1 else None
    peer = peer_j = None
    if world > 1 and os.environ.get("GSP_BENCH_EXCHANGE", "p2p") == "p2p":
        try:    # scores delivered by the scoring kernel itself through NVLink peer stores (symmetric memory)
            peer = sharding.PeerScoreSlices(e, group, dev)
            peer_j = sharding.PeerScoreSlices(e, group, dev) if fused else None
        except Exception as exc:   # symmetric memory unavailable: NCCL reduce-scatter of the full vector
            if rank == 0:
                print(f"[bench] peer scatter unavailable ({type(exc).__name__}: {exc}); using reduce-scatter", file=sys.stderr)
            peer = None
    num_keep = int(e * RETENTION)
    peer_used = peer is not None

    scores = torch.empty(slice_len if world > 1 else local, dtype=torch.float64, device=dev)
    scores_j = torch.empty(local, dtype=torch.float64, device=dev) if fused and world == 1 else None
    mask = torch.empty(local, dtype=torch.uint8, device=dev)
    ei_local = ei[:, e_lo:e_hi].contiguous() if world > 1 else ei
    kept_out = torch.empty((2, num_keep if world == 1 else local), dtype=torch.int64, device=dev)
    sharded_select = engine.ShardedSelect(dev, group) if world > 1 else None
    phases = ([BOTH] if fused else ["jaccard", "adamic_adar"]) + ["feature_cosine"]     # scoring launches of one step

    def new_events():
        return ({m: [torch.cuda.Event(enable_timing=True) for _ in range(2)] for m in phases},
                [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in METHODS])
    step_events = [new_events() for _ in range(args.steps)]
    spare_events = new_events()
    kernel_ms = {m: 0.0 for m in phases}
    kernel_ms["select+compact"] = 0.0
    kernel_ms_steps = {m: [] for m in kernel_ms}

    own_kernel_events = []

    def aa_weights():
        return graph.aa_node_weights_numpy()

    def normalized_features():
        # inside every step, replicated on every rank: 2.7 ms; all-gathering row-sharded results was measured at ~11 ms on
        # 8 GPUs (7/8 of 8.6 GB arrive at every rank)
        return graph.normalize_features(x)

    def score_single(m):
        """One metric's scores for this rank's slice of canonical positions (separate pass per metric)."""
        if world > 1 and m != "feature_cosine" and peer is not None:
            return sharding.owner_sharded_scores_p2p(graph, m, peer, node_range, aa_weights() if m == "adamic_adar" else None)[:local]
        if world > 1 and m != "feature_cosine":
            return sharding.owner_sharded_scores(graph, m, group, node_range, aa_weights() if m == "adamic_adar" else None,
                                                 scratch=full_scratch)[:local]
        if m == "jaccard":
            return graph.jaccard(e_lo, e_hi, out=scores[:local])
        if m == "adamic_adar":
            return graph.adamic_adar(aa_weights(), e_lo, e_hi, out=scores[:local])
        return graph.feature_cosine(normalized_features(), e_lo, e_hi, out=scores[:local])

    def score_both():
        """(jaccard, adamic_adar) slices from one streaming pass over the neighbour lists."""
        if world > 1 and peer is not None:
            pair = [torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)]
            own_kernel_events.append(pair)      # this rank's kernels alone, between the barriers (load balance)
            j, a = sharding.owner_sharded_jaccard_adamic_adar_p2p(graph, peer_j, peer, node_range, aa_weights(), kernel_events=pair)
            return j[:local], a[:local]
        if world > 1:
            j, a = sharding.owner_sharded_jaccard_adamic_adar(graph, group, node_range, aa_weights(), scratch=full_scratch)
            return j[:local], a[:local]
        return graph.jaccard_adamic_adar(aa_weights(), e_lo, e_hi, out_jaccard=scores_j, out_adamic_adar=scores[:local])

    def step(k):
        ev, ev_sel = step_events[k] if k is not None else spare_events
        sel = 0
        for m in phases:
            ev[m][0].record()
            outs = score_both() if m == BOTH else (score_single(m),)
            ev[m][1].record()
            for s_loc in outs:                     # every method: top-50 % select + mask + compaction of its own scores
                ev_sel[sel][0].record()
                if world > 1:    # distributed boundary search + fused mask / compaction; the kept count stays on the device
                    sharded_select(s_loc, num_keep, False, ei_local, mask=mask, out=kept_out)
                else:
                    engine.select_compact(s_loc, num_keep, False, ei_local, mask=mask, out=kept_out)
                ev_sel[sel][1].record()
                sel += 1

    def barrier():
        if world > 1:
            torch.distributed.barrier(group)
        torch.cuda.synchronize(dev)

    if peer is not None:   # one-off cross-check of the fused peer-store exchange against the NCCL reduce-scatter path
        ok = True
        for m in ("jaccard", "adamic_adar"):
            wts = aa_weights() if m == "adamic_adar" else None
            a = sharding.owner_sharded_scores_p2p(graph, m, peer, node_range, wts)[:local].clone()
            b = sharding.owner_sharded_scores(graph, m, group, node_range, wts, scratch=full_scratch)[:local]
            ok = ok and bool(torch.equal(a, b))
            if fused:
                fj, fa = score_both()
                ok = ok and bool(torch.equal(fj if m == "jaccard" else fa, b))
        flag = torch.tensor([1 if ok else 0], device=dev)
        torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN, group=group)
        if int(flag) == 0:
            if rank == 0:
                print("[bench] peer-store exchange disagrees with reduce-scatter; falling back", file=sys.stderr)
            peer = None
            peer_used = False
    for _ in range(max(args.warmup, 1)):
        step(None)
    own_kernel_events.clear()
    single_ms = {}
    if fused:   # each metric scored on its own (reported beside the fused pass; not part of the step)
        for m in ("jaccard", "adamic_adar"):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            score_single(m)
            barrier()
            a.record()
            for _ in range(args.steps):
                score_single(m)
            b.record()
            barrier()
            single_ms[m] = a.elapsed_time(b)
    # the reported number — K steps back to back, no host sync inside, max over ranks
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    if rank == 0:
        sampler.mark()
    launches0 = lib.gsp_launch_count()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for k in range(args.steps):
        step(k)
    t_end.record()
    barrier()
    for ev, ev_sel in step_events:
        for m in phases:
            kernel_ms_steps[m].append(ev[m][0].elapsed_time(ev[m][1]))
        kernel_ms_steps["select+compact"].append(sum(a.elapsed_time(b) for a, b in ev_sel))
    for m in kernel_ms:
        kernel_ms[m] = sum(kernel_ms_steps[m])
    launches = lib.gsp_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    elapsed_ms = torch.tensor([t_start.elapsed_time(t_end)], dtype=torch.float64, device=dev)
    rank_spread = None
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(elapsed_ms, op=dist.ReduceOp.MAX, group=group)
        dominant_local = max(phases, key=lambda m: kernel_ms[m])
        own = [a.elapsed_time(b) for a, b in own_kernel_events[-args.steps:]] if own_kernel_events else [kernel_ms[dominant_local] / args.steps]
        mine = torch.tensor([sum(own) / len(own)], dtype=torch.float64, device=dev)
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine, group=group)
        per_rank = [float(t) for t in allr]
        rank_spread = {"kernel": dominant_local, "what": "each rank's scoring kernels alone, between the exchange barriers (mean over the timed steps)",
                       "ms_per_rank": [round(v, 3) for v in per_rank], "min_ms": min(per_rank), "max_ms": max(per_rank)}
        for d in (kernel_ms, single_ms):
            for k in d:
                t = torch.tensor([d[k]], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
                d[k] = float(t)
    ms_per_step = float(elapsed_ms) / args.steps
    value = len(METHODS) * e / (ms_per_step * 1e-3)

    # graph statistics for the roofline accounting (rank 0; needs the graph, so before it is released below)
    s2 = graph.sum_degree_sq
    max_degree = graph.max_degree
    sum_min = common = 0.0
    if rank == 0:
        indptr_t, indices_t, _, rows_t = graph.export(with_data=False, with_rows=True)
        deg_t = indptr_t[1:] - indptr_t[:-1]
        sum_min = float(torch.minimum(deg_t[rows_t.long()], deg_t[indices_t.long()]).sum()) / 2.0
        del indptr_t, indices_t, rows_t, deg_t
        _, inter_t = graph.jaccard(return_counts=True)
        common = float(inter_t.sum(dtype=torch.int64)) / 2.0
        del inter_t
        torch.cuda.synchronize(dev)

    # ---- e2e: host inputs (pinned), H2D and D2H inside the timed region, through the reference-facing API only -------------
    e2e = None
    if not args.no_e2e:
        ei_host = ei.cpu().pin_memory()
        x_host = x.cpu().pin_memory()
        del graph, scores, scores_j, kept_out, ei_local, mask, full_scratch, peer, peer_j
        torch.cuda.empty_cache()
        host_data = gsr_b200.Data(edge_index=ei_host, x=x_host, num_nodes=n)
        h2d = (ei_host.numel() * 8 + x_host.numel() * 4) // world if world > 1 else ei_host.numel() * 8 + x_host.numel() * 4
        d2h = 0
        times, h2d_ms, breakdown = [], [], []
        import gc
        teardown_ms = []
        for it in range(1 + args.e2e_steps):
            gc.collect()
            barrier()
            t0 = time.perf_counter()
            ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
            ev0.record()
            if world == 1:
                data = host_data.to(dev, non_blocking=True)
                ev1.record()
                sp = gsr_b200.GraphSparsifier(data, str(dev))
            else:
                data = sharded_to_device(host_data, dev, group)
                ev1.record()
                sp = gsr_b200.GraphSparsifier(data, str(dev), group=group)
            marks = [("uploads queued", time.perf_counter())]
            sp.prefetch_scores(METHODS if fused else METHODS[2:], to_host=True)
            marks.append(("graph built, scoring + read-backs queued", time.perf_counter()))
            d2h = 0
            for m in METHODS:
                s = sp.compute_scores(m)                                  # np.ndarray fp64 on host (this rank's slice if N > 1)
                marks.append((m + ":scores_to_host", time.perf_counter()))
                out, msk = sp.sparsify(m, RETENTION, return_mask=True)    # Data (edge_index on device) + host bool mask
                marks.append((m + ":sparsify", time.perf_counter()))
                d2h += s.nbytes + msk.numel()
                del out, msk
            barrier()                      # every result is on the host (or, for the kept edge lists, on the device): the step ends here
            t1 = time.perf_counter()
            del sp, data, s                # teardown of the engine (cudaFree of the CSR ...) is reported beside the step, not in it
            gc.collect()
            torch.cuda.synchronize(dev)
            t2 = time.perf_counter()
            if it > 0:
                times.append(t1 - t0)
                teardown_ms.append((t2 - t1) * 1e3)
                h2d_ms.append(ev0.elapsed_time(ev1))   # the upload alone: tells a slow host link from a slow pipeline
                prev, row = t0, {}
                for name, t in marks:
                    row[name] = round((t - prev) * 1e3, 1)
                    prev = t
                breakdown.append(row)
        t_e2e = torch.tensor([float(np.median(times))], dtype=torch.float64, device=dev)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX, group=group)
        api = ("data_host.to(cuda) -> GraphSparsifier(data, cuda)" if world == 1 else
               "sharded_to_device(data_host, cuda, group) -> GraphSparsifier(data, cuda, group=group) [ShardedGraphSparsifier]")
        api += (" -> prefetch_scores(methods, to_host=True) -> compute_scores(m) [host fp64 ndarray] + sparsify(m, 0.5, "
                "return_mask=True) [host bool mask] for 3 metrics")
        e2e = {"value": len(METHODS) * e / float(t_e2e), "unit": "edges/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": float(t_e2e) * 1e3,
               "statistic": "median of the timed steps (max over ranks)", "ms_per_step_mean": sum(times) / len(times) * 1e3,
               "steps_ms": [round(t * 1e3, 1) for t in times], "h2d_ms": [round(t, 1) for t in h2d_ms],
               "h2d_gbs": h2d / (sum(h2d_ms) / len(h2d_ms) * 1e-3) / 1e9, "steps_breakdown_ms": breakdown[:3], "teardown_ms": [round(t, 1) for t in teardown_ms], "api": api,
               "note": ("bytes are per rank: every rank uploads 1/N of the inputs and reads back its slice of scores and masks"
                        if world > 1 else "single rank")}
        del host_data, ei_host, x_host
    else:
        del graph

    # ---- the other BASELINE configs -------------------------------------------------------------------------------------
    approx_er = None
    if not args.no_approx_er:
        del ei, x
        torch.cuda.empty_cache()
        approx_er = run_approx_er(args, dev, rank, world, group)
    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return
    small = variants = None
    if not args.no_small_configs and not args.no_e2e:
        small = run_small_configs(dev)
        variants = run_variants(dev)

    # ---- roofline -------------------------------------------------------------------------------------------------------
    peak, peak_src = measured_peak()
    frac_edges = local / e
    # Algorithmic bytes. The owner-hashed intersection streams, for every undirected pair, only the
    # SHORTER neighbour list: sum_pairs min(d_u, d_v) ids, plus the owner rows once (4E), neighbour metadata (16E) and the
    # fp64 output (8E); Adamic-Adar adds one 8-byte weight gather per common neighbour. SURVEY 8d's formula (4*S2 + 20E:
    # row(v) streamed once per directed edge) is reported beside it as frac_survey_formula.
    alg_bytes = {
        "jaccard": (4.0 * sum_min + 28.0 * e) * frac_edges,
        "adamic_adar": (4.0 * sum_min + 28.0 * e + 8.0 * common) * frac_edges,
        BOTH: (4.0 * sum_min + 36.0 * e + 8.0 * common) * frac_edges,
        "feature_cosine": (4.0 * args.dim * e + 4.0 * e + 8.0 * e) * frac_edges + 12.0 * args.dim * n,
        # P histogram passes are data dependent (2 when the boundary is one tie class, up to 6): counted at 6
        "select+compact": len(METHODS) * ((8.0 * local * 6) + (8.0 * local) + (8.0 * local + 16.0 * local + local) + 16.0 * num_keep / world),
    }
    survey_bytes = {"jaccard": (4.0 * s2 + 20.0 * e) * frac_edges, "adamic_adar": (4.0 * s2 + 20.0 * e + 8.0 * 2 * common) * frac_edges,
                    BOTH: (4.0 * s2 + 28.0 * e + 8.0 * 2 * common) * frac_edges}
    per_kernel = {}
    for k, ms in list(kernel_ms.items()) + list(single_ms.items()):
        avg = ms / args.steps
        gbs = alg_bytes[k] / (avg * 1e-3) / 1e9 if avg > 0 else 0.0
        per_kernel[k] = {"ms": avg, "alg_gb": alg_bytes[k] / 1e9, "achieved_gbs": gbs, "frac": gbs / peak,
                         "edges_per_s": ((2 if k == BOTH else 1) * local / (avg * 1e-3)) if k != "select+compact" else None,
                         "in_step": k in kernel_ms}
        if k in kernel_ms_steps:
            per_kernel[k]["ms_steps"] = [round(v, 3) for v in kernel_ms_steps[k]]
        if k in survey_bytes:
            per_kernel[k]["frac_survey_formula"] = survey_bytes[k] / (avg * 1e-3) / 1e9 / peak
    dominant = max(phases, key=lambda m: kernel_ms[m])
    roofline = {"bound": "hbm",
                "limiter": "instruction issue and the shared-memory pipe (ncu: issue slots 72 %, LSU shared wavefronts 65 % of cycles, "
                           "DRAM 9 % of peak): `frac` measures distance from a memory-bound kernel, not wasted bandwidth",
                "kernel": {BOTH: "cta_owner_kernel<2> (+warp_owner_kernel<2>): Jaccard and Adamic-Adar in one pass",
                           "jaccard": "cta_owner_kernel<0> (+warp_owner_kernel<0>)",
                           "adamic_adar": "cta_owner_kernel<1> (+warp_owner_kernel<1>)",
                           "feature_cosine": "featcos_kernel<float>"}[dominant],
                "achieved": per_kernel[dominant]["achieved_gbs"], "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": per_kernel[dominant]["frac"], "traffic": None,
                "byte_model": "owner schedule: 4*sum_pairs min(d_u,d_v) + 36*E + 8*T (DESIGN.md section 4)",
                "algorithmic_gb": per_kernel[dominant]["alg_gb"],
                "frac_survey_formula": per_kernel[dominant].get("frac_survey_formula"),
                "frac_step": sum(alg_bytes[k] for k in kernel_ms) / (ms_per_step * 1e-3) / 1e9 / peak,
                "per_method": per_kernel, "rank_spread": rank_spread,
                "approx_er": approx_er, "config1_config2": small, "config3_selection_variants": variants,
                "note": "duration = CUDA events on the launching stream around the scoring call, inside the timed steps"}
    try:   # DRAM traffic of the dominant kernel from the committed ncu capture (only valid for the workload it was taken on)
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        w = traffic["workload"]
        if (w["scale"], w["edge_factor"], w["dim"], w["n_gpus"]) == (args.scale, args.edge_factor, args.dim, world) and dominant in traffic:
            roofline["traffic"] = traffic[dominant]["dram_bytes_per_launch"] / 1e9
            roofline["traffic_unit"] = "GB per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, hub launch)"
            roofline["traffic_kernel"] = traffic[dominant]["kernel"]
            roofline["frac_dram"] = roofline["traffic"] / (per_kernel[dominant]["ms"] * 1e-3) / peak
    except Exception:
        pass

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:      # reported at N = 1 only
        c = cpu_sample(1)
        cpu_step, kind = cpu_step_fn(c)
        t0 = time.perf_counter()
        cpu_step()
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": len(METHODS) * c["edges"] / dt, "unit": "edges/s", "cores": 1, "kind": kind,
                        "host_cores_available": len(os.sched_getaffinity(0)), "seconds": dt, "sample": cpu_sample_description(c)}

    line = {
        "metric": METRIC, "value": value, "unit": "edges/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": {"workload": workload_name(args), "nodes": n, "directed_edges": e, "max_degree": max_degree,
                   "sum_degree_sq": s2, "sum_pairs_min_degree": sum_min, "common_neighbour_pairs": common, "retention": RETENTION, "l2": "inputs_larger_than_L2",
                   "scoring_passes": ("Jaccard + Adamic-Adar from one streaming pass (gsp_jaccard_adamic_adar), FeatCos" if fused
                                      else "one pass per method"),
                   "parallelism": (f"x{world}: owner-sharded Jaccard/AA, exchange = " + ("peer stores from the scoring kernel (NVLink symmetric memory)" if peer_used else "NCCL reduce-scatter") + ", edge-sliced FeatCos/select, CSR+features replicated") if world > 1 else "single GPU"},
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e,
        "gpu_launches": int(launches), "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
