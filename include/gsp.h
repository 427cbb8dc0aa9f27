/*
 * gsp.h — C ABI of libgsp.so, the B200 (sm_100a) edge-scoring sparsification engine.
 *
 * The reference (ilias-laoukili/gnn-sparsification-research) is pure Python: its "plugin
 * boundary" for this path is the class `GraphSparsifier` plus five `calculate_*_scores`
 * functions (reference src/sparsification/__init__.py:16-28), not an FFI. This header is the
 * C boundary a maintainer would bind from that Python layer (ctypes stub in INTEGRATION.md);
 * each entry point cites the reference statement(s) it replaces (paths relative to the
 * reference root).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only; no C++/torch types cross the boundary.
 *   - Every `d_*` pointer is DEVICE memory owned by the caller, on the graph's device. Outputs
 *     are caller-allocated. The library owns only `gsp_graph` and stream-ordered scratch.
 *   - `stream` is a `cudaStream_t` passed as `void*` (NULL = legacy default stream). All work
 *     is enqueued on it; calls return without synchronising unless stated otherwise.
 *   - Return value: 0 on success, non-zero `gsp_status` otherwise; `gsp_last_error()` returns
 *     a thread-local message. No exceptions cross the ABI. There is NO CPU fallback: without
 *     a CUDA device every compute entry point fails with GSP_ERR_CUDA.
 *   - Edge order everywhere is canonical CSR order (rows ascending, columns ascending inside a
 *     row, duplicates merged) == the order of SciPy's `adj.nonzero()` that every reference
 *     score vector uses. Scoring calls take a half-open range
 *     [e_begin, e_end) of canonical edge positions and write `e_end - e_begin` values, so a
 *     caller can shard edges over GPUs with the CSR replicated.
 */
#ifndef GSP_H_
#define GSP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSP_VERSION 100 /* major*100 + minor */

typedef enum gsp_status {
    GSP_OK = 0,
    GSP_ERR_INVALID = 1, /* bad argument (null pointer, negative size, index out of range, ...) */
    GSP_ERR_CUDA = 2,    /* a CUDA runtime call failed / no device */
    GSP_ERR_NOMEM = 3,
    GSP_ERR_UNSUPPORTED = 4
} gsp_status;

typedef struct gsp_graph gsp_graph; /* opaque: canonical CSR (+ transpose when asymmetric) on one device */

typedef struct gsp_graph_info {
    int64_t num_nodes;
    int64_t num_input_edges; /* columns of the edge_index handed in (reference `num_edges`, core.py:67) */
    int64_t nnz;             /* canonical entries == length of every score vector (adj.nnz) */
    int64_t num_undirected;  /* entries with row < col (ApproxER's m, metrics.py:239-242) */
    int64_t max_degree;
    double sum_degree_sq;    /* sum_u deg(u)^2: intersection work / roofline bytes */
    int32_t symmetric;       /* pattern symmetric => row(v) stands in for col(v) in Jaccard */
    int32_t input_canonical; /* input was already (row,col)-sorted and duplicate free */
    int32_t unit_weights;    /* every merged value == 1.0 */
    int32_t device;
} gsp_graph_info;

int gsp_version(void);
const char* gsp_last_error(void);
/* Number of CUDA kernels this library has launched in the process so far (monotonic; for benchmarks). */
uint64_t gsp_launch_count(void);
/* Device-memory hook. Before the first graph is created a host may hand the library its own allocator (the Python
 * layer passes torch's caching allocator): every device byte the library holds — graph arrays, lazily built side
 * structures, scratch — then comes from it, stays visible to the host framework and is recycled across graph builds
 * instead of paying cudaMalloc / cudaFree of gigabytes each time. alloc(bytes, device, stream) returns a device pointer
 * usable on `stream` (NULL on failure); free(ptr) may recycle the block for later allocations on that stream
 * (cudaFreeAsync semantics); gsp_graph_destroy synchronises the device first. Pass NULL, NULL (the default) for
 * cudaMalloc / cudaFree + the private scratch pool below. Not to be changed while library-owned memory is live. */
typedef void* (*gsp_alloc_fn)(size_t bytes, int device, void* stream);
typedef void (*gsp_free_fn)(void* ptr);
int gsp_set_allocator(gsp_alloc_fn alloc_fn, gsp_free_fn free_fn);

/* The library allocates its scratch from a private stream-ordered pool per device that keeps at most GSP_SCRATCH_KEEP_MB
 * (default 1024) of freed memory across synchronisation points; this hands all of it back to the driver (synchronises
 * the device). */
int gsp_trim_scratch(void);

/* ---- graph ----------------------------------------------------------------------------------
 * Replaces reference core.py:70-74: sp.csr_matrix((ones(E), (row, col)), shape=(n, n)).
 * d_row/d_col: int64[num_edges] (the two rows of PyG's edge_index). d_val: optional fp64
 * values (NULL = ones); duplicates are summed (data = multiplicity), explicit zeros dropped
 * like `adj.nonzero()`. Synchronises `stream` once (the merged nnz is needed on the host).
 * Fails with GSP_ERR_INVALID if an index is outside [0, num_nodes). */
int gsp_graph_create(int64_t num_nodes, int64_t num_edges, const int64_t* d_row, const int64_t* d_col,
                     const double* d_val, void* stream, gsp_graph** out);
void gsp_graph_destroy(gsp_graph* g);
int gsp_graph_get_info(const gsp_graph* g, gsp_graph_info* out);
/* Copies the canonical CSR out (any pointer may be NULL): indptr int64[n+1], indices int32[nnz],
 * data fp64[nnz], rows int32[nnz] (row id of every canonical position, i.e. adj.nonzero()[0]). */
int gsp_graph_export(const gsp_graph* g, int64_t* d_indptr, int32_t* d_indices, double* d_data, int32_t* d_rows,
                     void* stream);
/* int32[n] binarised row degrees (reference metrics.py:44,100). */
int gsp_graph_degrees(const gsp_graph* g, int32_t* d_deg, void* stream);
/* int32[nnz]: id of the undirected edge {u,v} each canonical position belongs to (rank of its
 * (min,max) orientation among row<col positions; -1 for self loops / unmatched directions). */
int gsp_graph_undirected_ids(const gsp_graph* g, int32_t* d_uid, void* stream);

/* ---- scoring --------------------------------------------------------------------------------
 * Jaccard — replaces calculate_jaccard_scores, reference metrics.py:43-64:
 *   I = |row(u) ∩ col(v)| (== (Ab@Ab)[u,v]), score = I / (deg u + deg v - I), 0 when the union is 0.
 * d_inter (optional) receives the exact integer intersection counts. */
int gsp_jaccard(const gsp_graph* g, int64_t e_begin, int64_t e_end, int32_t* d_inter, double* d_score, void* stream);

/* Adamic-Adar — replaces calculate_adamic_adar_scores, reference metrics.py:99-121:
 *   score = sum over x in row(u) ∩ row(v), DESCENDING x, of fl(w[x]*w[x]) (sequential fp64 adds: the
 *   association order SciPy's SpGEMM uses, needed for bit-equal scores / keep-masks).
 * d_node_w: fp64[n] node weights 1/sqrt(max(log(deg+1),1e-10)); NULL = computed on device by
 * gsp_aa_node_weights (CUDA libm; may differ from NumPy's SIMD log in the last bit). */
int gsp_adamic_adar(const gsp_graph* g, const double* d_node_w, int64_t e_begin, int64_t e_end, double* d_score,
                    void* stream);
int gsp_aa_node_weights(const gsp_graph* g, double* d_node_w, void* stream);
/* d_node_w[u] = d_table[deg(u)]: node weights from a caller-supplied per-degree table of max_degree + 1 entries (the host
 * evaluates the reference's NumPy expression once per distinct degree; keeps the libm-defined last bit). */
/* d_dst[i] = src[i] by a kernel (SM loads), not by the copy engine: src may be page-locked host memory (unified
 * addressing), so a small host-evaluated table reaches the device even while a large upload queued earlier owns the
 * copy engine. */
int gsp_copy_f64(const double* src, double* d_dst, int64_t count, void* stream);
int gsp_aa_node_weights_from_table(const gsp_graph* g, const double* d_table, int64_t table_len, double* d_node_w,
                                   void* stream);

/* Jaccard and Adamic-Adar in ONE streaming pass (callers that score a graph with both metrics — the reference's
 * ablation loop, src/experiments/ablation.py:220-270, and scripts/nb05_roman_empire/roman_empire_gpu.py:81-102 do):
 * the hit ballots the ordered Adamic-Adar sum needs also give the intersection count, so d_jaccard / d_inter come
 * at the price of the Adamic-Adar pass alone. Bit-identical to gsp_jaccard + gsp_adamic_adar (asymmetric graphs run
 * the two passes internally: their intersections differ). d_inter may be NULL. */
int gsp_jaccard_adamic_adar(const gsp_graph* g, const double* d_node_w, int64_t e_begin, int64_t e_end, int32_t* d_inter,
                            double* d_jaccard, double* d_adamic_adar, void* stream);

/* Owner-sharded variants for multi-GPU scoring of a SYMMETRIC graph: every undirected pair {u,v} is evaluated on
 * exactly one rank — the one whose node range [node_begin, node_end) holds the pair's owner (the endpoint with the
 * larger degree, ties: smaller id) — and its score is written at BOTH directed positions of full-length (nnz)
 * output arrays; positions of pairs owned elsewhere are not touched. The caller zero-fills the arrays and
 * reduce-scatters (sum) them over the ranks. gsp_owner_costs gives fp64[n] per-node work estimates for balancing
 * the node ranges. GSP_ERR_UNSUPPORTED for asymmetric graphs (shard those by edge range). */
int gsp_jaccard_owned(const gsp_graph* g, int64_t node_begin, int64_t node_end, int32_t* d_inter_full,
                      double* d_score_full, void* stream);
int gsp_adamic_adar_owned(const gsp_graph* g, const double* d_node_w, int64_t node_begin, int64_t node_end,
                          double* d_score_full, void* stream);
int gsp_jaccard_adamic_adar_owned(const gsp_graph* g, const double* d_node_w, int64_t node_begin, int64_t node_end,
                                  double* d_jaccard_full, double* d_adamic_adar_full, void* stream);
int gsp_owner_costs(const gsp_graph* g, double* d_cost, void* stream);
/* Dealt ownership — a finer unit of owner sharding than contiguous node ranges: d_owner_rank is uint8[num_nodes]
 * (copied into the graph), and from now on every *_owned / *_owned_scatter call on this handle evaluates only the
 * pairs whose owner o has node_begin <= o < node_end AND d_owner_rank[o] == rank. Dealing the owners in cost order
 * (gsp_owner_costs, snake order over the ranks) gives every rank the same mix of hub, medium and small owners, so
 * errors of the cost estimate cancel instead of piling up on the rank that holds the largest hubs. NULL clears the
 * deal. The plain scoring calls (gsp_jaccard, ...) ignore it. */
int gsp_graph_set_owner_deal(gsp_graph* g, const uint8_t* d_owner_rank, int32_t rank, void* stream);
/* Fused scoring + exchange: like the *_owned calls, but every score is stored by the scoring kernel directly into the
 * slice of the rank that owns its position — d_slices is a DEVICE array of `world` pointers, d_slices[k] = base of rank
 * k's fp64[slice_len] slice (position p lives at d_slices[p / slice_len][p % slice_len]); the pointers are typically
 * peer allocations mapped over NVLink (e.g. torch symmetric memory), so each score crosses the fabric once and no
 * reduce-scatter / zero-fill is needed. The caller brackets the call with a cross-rank barrier. */
int gsp_jaccard_owned_scatter(const gsp_graph* g, int64_t node_begin, int64_t node_end, double* const* d_slices,
                              int32_t world, int64_t slice_len, void* stream);
int gsp_adamic_adar_owned_scatter(const gsp_graph* g, const double* d_node_w, int64_t node_begin, int64_t node_end,
                                  double* const* d_slices, int32_t world, int64_t slice_len, void* stream);
int gsp_jaccard_adamic_adar_owned_scatter(const gsp_graph* g, const double* d_node_w, int64_t node_begin, int64_t node_end,
                                          double* const* d_jaccard_slices, double* const* d_adamic_adar_slices,
                                          int32_t world, int64_t slice_len, void* stream);

/* degree product — replaces reference core.py:167-172 (`degree` metric; raw-value row sums). */
int gsp_degree_product(const gsp_graph* g, int64_t e_begin, int64_t e_end, double* d_score, void* stream);

/* Feature cosine — replaces calculate_feature_cosine_scores, reference metrics.py:344-358.
 * Step 1 (per node): xhat = x / max(sqrt(pairwise_sum(x*x)), 1e-10)   [metrics.py:344-346]
 * Step 2 (per edge): score = (double) max(pairwise_sum(xhat_u * xhat_v), 0)   [metrics.py:351-358]
 * Arithmetic runs in the feature dtype with NumPy's pairwise-summation tree,
 * products rounded before they are summed, no FMA contraction. `ld` = row stride in elements. */
int gsp_featcos_normalize_f32(int64_t num_nodes, int32_t dim, const float* d_x, int64_t ld, float* d_xhat,
                              int64_t ld_out, void* stream);
int gsp_featcos_f32(const gsp_graph* g, const float* d_xhat, int32_t dim, int64_t ld, int64_t e_begin, int64_t e_end,
                    double* d_score, void* stream);
/* Packed fast path for fp32 features with dim in {32, 64, 96, 128}: d_packed is an opaque, accumulator-major copy of
 * the normalised rows (16-byte aligned, ld % 4 == 0) that only gsp_featcos_f32_packed reads — 128-bit loads, the same
 * arithmetic and association order, bit-identical scores. */
int gsp_featcos_normalize_f32_packed(int64_t num_nodes, int32_t dim, const float* d_x, int64_t ld, float* d_packed,
                                     int64_t ld_out, void* stream);
int gsp_featcos_f32_packed(const gsp_graph* g, const float* d_packed, int32_t dim, int64_t ld, int64_t e_begin,
                           int64_t e_end, double* d_score, void* stream);
int gsp_featcos_normalize_f64(int64_t num_nodes, int32_t dim, const double* d_x, int64_t ld, double* d_xhat,
                              int64_t ld_out, void* stream);
int gsp_featcos_f64(const gsp_graph* g, const double* d_xhat, int32_t dim, int64_t ld, int64_t e_begin,
                    int64_t e_end, double* d_score, void* stream);

/* ---- approximate effective resistance ----------------------------------------------------------
 * Replaces calculate_approx_effective_resistance_scores, reference metrics.py:232-298, for the
 * projection columns [0, k) of d_R (fp64 [m, ldr] row-major, m = num_undirected; the caller
 * slices columns to shard them over GPUs):  Y = B R;  Z = CG(L + reg*I, Y) column by column with
 * SciPy's cg semantics (x0 = 0, test ||r|| < rtol*||b|| before each update, at most max_iters
 * updates, the partial iterate is kept);  d_partial[e] = sum_j (Z[u,j] - Z[v,j])^2 for e in
 * [e_begin, e_end). The caller all-reduces partial sums over column shards and then applies
 * gsp_er_finalize (max(., 1e-10), metrics.py:296-297). d_iters (optional) int32[k] = CG updates per column. */
int gsp_approx_er_partial(const gsp_graph* g, const double* d_R, int64_t ldr, int32_t k, int32_t max_iters,
                          double rtol, double reg, int64_t e_begin, int64_t e_end, double* d_partial,
                          int32_t* d_iters, void* stream);
/* The same with the projection generated where it is used (throughput mode): R[e, c] = N(0,1) / sqrt(k_total) as a pure
 * function of (seed, e, col_begin + c) — Philox4x32-10, Box-Muller — so the fp64 [m, k] matrix (31.7 GB at the
 * products shape with k = 64) is never materialised. Columns [col_begin, col_begin + k) of k_total: ranks that split the
 * columns draw from one matrix. gsp_philox_projection writes those entries out (fp64 [m, k] row-major): feeding them to
 * gsp_approx_er_partial gives bit-identical results (tests). Not the reference's PCG64 stream: parity runs pass d_R. */
int gsp_approx_er_partial_philox(const gsp_graph* g, uint64_t seed, int32_t col_begin, int32_t k, int32_t k_total,
                                 int32_t max_iters, double rtol, double reg, int64_t e_begin, int64_t e_end,
                                 double* d_partial, int32_t* d_iters, void* stream);
int gsp_philox_projection(uint64_t seed, int64_t m, int32_t col_begin, int32_t k, int32_t k_total, double* d_R, void* stream);
int gsp_er_finalize(double* d_score, int64_t count, void* stream);
/* The batched solver on its own: X = CG(D - A + reg*I, RHS) for fp64 [n, k] row-major right-hand sides, same column-wise
 * SciPy cg semantics, symmetric graphs. Used by the exact effective resistance (reference metrics.py:158-175: the columns
 * of the Laplacian pseudo-inverse that the edges need, instead of a dense O(n^3) pinv) and by the algebraic connectivity
 * of large components (metrics.py:480-511: shift-and-invert Lanczos). d_iters (optional) int32[k]. */
int gsp_laplacian_solve(const gsp_graph* g, const double* d_rhs, int32_t k, int32_t max_iters, double rtol, double reg,
                        double* d_x, int32_t* d_iters, void* stream);

/* ---- metric backbone ---------------------------------------------------------------
 * Shortest-path lengths from the sources [src_begin, src_begin + src_count) to every node of a SYMMETRIC graph with
 * non-negative edge lengths d_weights (fp64[nnz], canonical order) — replaces the all-pairs Dijkstra of reference
 * metric_backbone.py:84-87. d_dist is fp64 [num_nodes, src_count] row-major (dist[v, s]); unreachable = +inf. In-place
 * (min,+) relaxation sweeps until a sweep changes nothing (at most max_rounds; *rounds_out, host memory, receives the
 * count). The fixpoint equals Dijkstra's lengths bit for bit (sums accumulate from the source; min is exact).
 * Synchronises the stream every few sweeps. */
int gsp_sssp_batch(const gsp_graph* g, const double* d_weights, int64_t src_begin, int32_t src_count, double* d_dist,
                   int32_t max_rounds, int32_t* rounds_out, void* stream);
/* Same relaxation from an explicit list of source nodes (d_sources int32[src_count], ids in [0, num_nodes)); column c of
 * d_dist holds the distances from d_sources[c]. With unit weights these are hop counts: the sampled
 * nx.shortest_path_length calls of reference metrics.py:361-442 (compute_geodesic_preservation). */
int gsp_sssp_sources(const gsp_graph* g, const double* d_weights, const int32_t* d_sources, int32_t src_count,
                     double* d_dist, int32_t max_rounds, int32_t* rounds_out, void* stream);

/* ---- selection ---------------------------------------------------------------------------------
 * Radix-histogram select with stable (score, position) tie-breaking — replaces the full argsort of
 * reference core.py:232-240 (and :446-451 with an exclusion mask). Keys are fp64 scores mapped to
 * order-preserving uint64. Phased so a multi-GPU caller can all-reduce the histograms:
 *
 *   gsp_select_begin(state, num_keep, keep_lowest)
 *   for pass in 0 .. GSP_SELECT_PASSES-1:
 *       gsp_select_histogram(scores, count, exclude, state, pass, hist)      (local slice)
 *       [all-reduce hist: uint64[GSP_SELECT_BINS], sum]
 *       gsp_select_pick(state, hist, pass)
 *   gsp_select_count_ties(scores, count, exclude, state, tie_count)           (local slice)
 *   [all-gather tie counts -> ties_before = sum over lower ranks, ties_total]
 *   gsp_select_write_mask(scores, count, exclude, state, ties_before, ties_total, mask)
 *
 * Contract (the reference run with a stable argsort): top-k keeps {s > t} plus the
 * HIGHEST-position members of {s == t}; keep_lowest keeps {s < t} plus the LOWEST-position members.
 * `state` is GSP_SELECT_STATE_BYTES of caller-owned device memory. Everything is stream-ordered; no
 * host synchronisation. gsp_select_mask runs the whole sequence for one GPU (and, knowing that no other rank holds
 * keys, stops the histogram rounds as soon as the boundary bucket turns out to be a single tie class).
 * num_keep must be in [0, number of non-excluded scores]. */
#define GSP_SELECT_BINS 2048
#define GSP_SELECT_PASSES 6
#define GSP_SELECT_STATE_BYTES 16384

int gsp_select_begin(void* d_state, int64_t num_keep, int keep_lowest, void* stream);
int gsp_select_histogram(const double* d_scores, int64_t count, const uint8_t* d_exclude, const void* d_state,
                         int pass, uint64_t* d_hist, void* stream);
int gsp_select_pick(void* d_state, const uint64_t* d_hist, int pass, void* stream);
int gsp_select_count_ties(const double* d_scores, int64_t count, const uint8_t* d_exclude, const void* d_state,
                          int64_t* d_tie_count, void* stream);
int gsp_select_write_mask(const double* d_scores, int64_t count, const uint8_t* d_exclude, const void* d_state,
                          const int64_t* d_ties_before, const int64_t* d_ties_total, int or_into, uint8_t* d_mask,
                          void* stream);
int gsp_select_mask(const double* d_scores, int64_t count, int64_t num_keep, int keep_lowest,
                    const uint8_t* d_exclude, int or_into, uint8_t* d_mask, void* stream);

/* Single-GPU select + mask + compaction (+ "-W" weights) in one call — what `GraphSparsifier.sparsify` needs
 * (reference core.py:232-245, roman_empire_gpu.py:248-256): the same boundary search as gsp_select_mask, then two
 * passes instead of four (per-block counts; mask bytes + kept edge_index columns + weights in one sweep). The extrema
 * of the kept scores are the boundary score and the best score, so the weights need no pass of their own. d_mask
 * (uint8[count]) and d_out_weight (fp32[out_ld]) may be NULL; d_edge_index is int64 [2, count] with row stride ld;
 * outputs as gsp_compact_edges. */
int gsp_select_compact(const double* d_scores, int64_t count, int64_t num_keep, int keep_lowest,
                       const int64_t* d_edge_index, int64_t ld, uint8_t* d_mask, int64_t* d_out_edge_index,
                       int64_t out_ld, float* d_out_weight, int invert_weights, int64_t* d_num_kept, void* stream);

/* Sharded select + mask + compaction (+ "-W" weights): the multi-GPU form of gsp_select_compact, one rank's contiguous
 * slice per call (rank order == canonical position order). Two small all-gathers replace the all-reduces of the phased
 * protocol above and carry what ends the rounds early and what the weights need:
 *
 *   gsp_select_begin(state, num_keep, keep_lowest)
 *   for pass in 0 .. GSP_SELECT_PASSES-1:
 *       gsp_select_histogram_slot(scores, count, state, pass, slot)   slot = uint64[GSP_SELECT_SLOT_WORDS]:
 *                                                                     histogram | ~min live key | max live key
 *       [all-gather slot -> slots[nranks][GSP_SELECT_SLOT_WORDS]]
 *       gsp_select_pick_slots(state, slots, nranks, pass)             sums the histograms; equal global extrema = the
 *                                                                     boundary bucket is one tie class: later passes no-ops
 *   gsp_select_tally(scores, count, state, scratch, totals)           totals = int64[2]: keys below the boundary, ties
 *   [all-gather totals -> rank_totals[nranks][2]]
 *   gsp_select_emit(scores, count, state, scratch, rank_totals, rank, nranks, edge_index, ld, mask, out, out_ld, w, ..)
 *
 * The emit writes this rank's mask slice and its kept columns in position order (the caller concatenates the ranks'
 * lists); weights use the boundary score and the globally best score (pass 0 extrema) as the kept set's extrema, as
 * gsp_select_compact does (reference core.py:232-245, roman_empire_gpu.py:248-256). `scratch` is
 * GSP_SELECT_SCRATCH_BYTES of caller-owned device memory shared by tally and emit. */
#define GSP_SELECT_SLOT_WORDS 2050
#define GSP_SELECT_SCRATCH_BYTES 18944
int gsp_select_histogram_slot(const double* d_scores, int64_t count, const void* d_state, int pass, uint64_t* d_slot,
                              void* stream);
int gsp_select_pick_slots(void* d_state, const uint64_t* d_slots, int32_t nranks, int pass, void* stream);
int gsp_select_tally(const double* d_scores, int64_t count, const void* d_state, void* d_scratch, int64_t* d_totals,
                     void* stream);
int gsp_select_emit(const double* d_scores, int64_t count, const void* d_state, const void* d_scratch,
                    const int64_t* d_rank_totals, int32_t rank, int32_t nranks, const int64_t* d_edge_index, int64_t ld,
                    uint8_t* d_mask, int64_t* d_out_edge_index, int64_t out_ld, float* d_out_weight, int invert_weights,
                    int64_t* d_num_kept, void* stream);

/* Degree-aware guarantee phase — replaces reference core.py:421-435: for every source node
 * (d_src = edge_index[0], positional, any order) mark its top min(min_per_node, out-degree) edges by
 * (score, position). d_mask (uint8[count]) is overwritten; d_num_marked (int64[1]) = |G|. The fill phase
 * (core.py:444-451) is gsp_select_mask with d_exclude = d_mask, or_into = 1. */
int gsp_degree_aware_guarantee(const int64_t* d_src, const double* d_scores, int64_t count, int64_t num_nodes,
                               int32_t min_per_node, uint8_t* d_mask, int64_t* d_num_marked, void* stream);

/* ---- compaction / edge weights -------------------------------------------------------------------
 * Replaces `edge_index[:, mask]` (reference core.py:242) and the "-W" weight derivation
 * (scripts/nb05_roman_empire/roman_empire_gpu.py:248-256): w = (s - min)/(max - min + 1e-8) over the kept
 * scores (1 - w when invert), evaluated in fp64 and rounded to fp32. d_edge_index is int64 [2, count]
 * (row stride `ld`); outputs keep the original column order. d_out_weight / d_scores may be NULL.
 * d_num_kept (int64[1]) receives the number of kept columns; out buffers must hold `out_ld` columns. */
int gsp_compact_edges(const int64_t* d_edge_index, int64_t ld, int64_t count, const uint8_t* d_mask,
                      const double* d_scores, int invert_weights, int64_t* d_out_edge_index, int64_t out_ld,
                      float* d_out_weight, int64_t* d_num_kept, void* stream);

/* ---- consumer side: GCN normalisation and propagation of the kept sub-graph ---------
 * The reference trains GCN / GCN* on the sparsified graph with GCNConv(cached=False, normalize=True) and the
 * optional "-W" edge weights (src/models/gnn.py:222-223,244), i.e. every layer of every forward pass re-runs
 * torch_geometric's gcn_norm (torch-geometric >= 2.3, pyproject.toml:31; not vendored) and a scatter-add
 * propagate on the same kept edges. gsp_gcn_norm restates gcn_norm(add_self_loops=True, improved=False,
 * flow="source_to_target"): self-loop edges are dropped from the list, one loop per node is appended (weight =
 * the node's LAST existing loop weight, else 1), deg[t] = sum of weights over edges INTO t (fp32, in list
 * order), out_weight = deg^-1/2[row] * w * deg^-1/2[col] (inf -> 0). d_weight may be NULL (all ones). Output
 * buffers hold num_edges + num_nodes entries; d_out_count (int64[1], device, optional) receives the count. */
int gsp_gcn_norm(int64_t num_nodes, int64_t num_edges, const int64_t* d_row, const int64_t* d_col, const float* d_weight,
                 int64_t* d_out_row, int64_t* d_out_col, float* d_out_weight, int64_t* d_out_count, void* stream);
/* Stable grouping of an edge list by target: d_indptr int64[num_nodes + 1], d_perm int64[num_edges] (edge positions,
 * targets ascending, list order kept inside a target). */
int gsp_target_order(int64_t num_nodes, int64_t num_edges, const int64_t* d_col, int64_t* d_indptr, int64_t* d_perm,
                     void* stream);
/* GCNConv's propagate (sum aggregation at the targets): d_out[t, :] = sum over the edges e into t, in list order, of
 * d_weight[e] * d_x[d_row[e], :] — fp32, multiply then add (no FMA), so the result equals a sequential CPU
 * scatter-add bit for bit and does not change from run to run. d_x: fp32 [num_nodes, dim] with row stride ldx. */
int gsp_gcn_propagate(int64_t num_nodes, const int64_t* d_indptr, const int64_t* d_perm, const int64_t* d_row,
                      const float* d_weight, const float* d_x, int32_t dim, int64_t ldx, float* d_out, int64_t ldo,
                      void* stream);

/* ---- topology of the (sparsified) graph ---------------------------------------------
 * Building blocks of reference src/sparsification/metrics.py:445-520 `compute_topology_metrics` (NetworkX there).
 * Symmetric graphs only (GSP_ERR_UNSUPPORTED otherwise).
 * gsp_node_triangles: d_inter = the int32[nnz] intersection counts of gsp_jaccard; d_pairs[v] (int64) = number of
 * ordered neighbour pairs of v that are adjacent = 2 * triangles through v, d_degree[v] (int32) = neighbours other
 * than v itself — the t and d of nx.clustering: c_v = t / (d (d - 1)) (metrics.py:468 nx.average_clustering).
 * Self loops are discounted the way NetworkX does (a node is not its own neighbour). */
int gsp_node_triangles(const gsp_graph* g, const int32_t* d_inter, int64_t* d_pairs, int32_t* d_degree, void* stream);
/* d_label[v] (int32) = smallest node id of v's connected component (metrics.py:471 nx.connected_components):
 * min-label hooking + pointer jumping; *rounds_out (host, optional) = sweeps used. Reads one flag back per sweep. */
int gsp_connected_components(const gsp_graph* g, int32_t* d_label, int32_t* rounds_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GSP_H_ */
