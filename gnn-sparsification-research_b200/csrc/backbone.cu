// backbone.cu — batched shortest-path lengths for the global metric backbone (a "next" row).
//
// Replaces the all-pairs Dijkstra of reference src/sparsification/metric_backbone.py:84-87
// (`nx.all_pairs_dijkstra_path_length`). An edge (u,v) is kept iff w_uv <= d(u,v) + eps, so what is needed is the
// shortest-path length between adjacent nodes. Distances for a batch of S sources live in an [n, S] fp64 matrix
// (sources contiguous) and are relaxed in place with a pull-style (min,+) sweep over the CSR rows — the same gather
// shape as the ApproxER SpMM: lane = source column, a neighbour row is one coalesced request. The fixpoint
// dist[v] = min_x fl(dist[x] + w_xv) is the value Dijkstra computes (sums accumulated from the source outward), and
// min is exact, so the result does not depend on the relaxation order. A sweep with no update ends the batch.
#include "common.cuh"

namespace gsp {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / kWarp;
constexpr int kCheckEvery = 4;   // sweeps between host reads of the "changed" flag

// column c starts from node src_begin + c, or from sources[c] when an explicit source list is given
__global__ void sssp_init_kernel(int64_t n, int64_t src_begin, const int32_t* __restrict__ sources, int S,
                                 double* __restrict__ dist) {
    const int64_t total = n * (int64_t)S;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = i / S;
        const int c = (int)(i - v * S);
        const int64_t src = sources ? (int64_t)sources[c] : src_begin + c;
        dist[i] = (v == src) ? 0.0 : __longlong_as_double(0x7ff0000000000000ll);
    }
}

__global__ void __launch_bounds__(kThreads)
sssp_relax_kernel(int64_t n, const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                  const double* __restrict__ weights, int S, double* dist, int* __restrict__ changed) {
    const int c = blockIdx.y * kWarp + lane_id();
    const bool col_ok = c < S;
    bool any_update = false;
    const int64_t warp = blockIdx.x * (int64_t)kWarps + (threadIdx.x >> 5);
    for (int64_t v = warp; v < n; v += (int64_t)gridDim.x * kWarps) {
        const int64_t p0 = indptr[v], p1 = indptr[v + 1];
        const double cur = col_ok ? dist[v * (int64_t)S + c] : 0.0;
        double best = cur;
        for (int64_t t0 = p0; t0 < p1; t0 += 4) {
            int32_t x[4];
            double w[4], dx[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                x[u] = t0 + u < p1 ? __ldg(indices + t0 + u) : -1;
                w[u] = t0 + u < p1 ? __ldg(weights + t0 + u) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) dx[u] = (col_ok && x[u] >= 0) ? dist[(int64_t)x[u] * S + c] : best;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (x[u] < 0) break;
                const double cand = __dadd_rn(dx[u], w[u]);
                best = cand < best ? cand : best;
            }
        }
        if (col_ok && best < cur) {
            dist[v * (int64_t)S + c] = best;
            any_update = true;
        }
    }
    if (__any_sync(0xffffffffu, any_update) && lane_id() == 0) atomicOr(changed, 1);
}

}  // namespace
}  // namespace gsp

using namespace gsp;

static int sssp_run(const Graph* g, const double* d_weights, int64_t src_begin, const int32_t* d_sources, int32_t src_count,
                    double* d_dist, int32_t max_rounds, int32_t* rounds_out, cudaStream_t s) {
    GSP_REQUIRE(max_rounds >= 1, "max_rounds must be >= 1");
    if (rounds_out) *rounds_out = 0;
    if (src_count == 0 || g->n == 0) return GSP_OK;
    GSP_REQUIRE(d_dist && (g->nnz == 0 || d_weights), "NULL argument");
    if (!g->symmetric) {
        set_error("batched shortest paths need a symmetric (undirected) graph");
        return GSP_ERR_UNSUPPORTED;
    }
    const int64_t n = g->n;
    const int S = src_count;
    sssp_init_kernel<<<grid_for(n * (int64_t)S, 256), 256, 0, s>>>(n, src_begin, d_sources, S, d_dist);
    GSP_CHECK_LAUNCH();
    Scratch<int> changed;
    GSP_CUDA_TRY(changed.alloc(1, s));
    const int strips = (S + kWarp - 1) / kWarp;
    int64_t row_blocks = (static_cast<int64_t>(kNumSMs) * 8 + strips - 1) / strips;
    const int64_t max_rb = (n + kWarps - 1) / kWarps;
    if (row_blocks > max_rb) row_blocks = max_rb;
    if (row_blocks < 1) row_blocks = 1;
    const dim3 grid((unsigned)row_blocks, (unsigned)strips);
    int rounds = 0;
    while (rounds < max_rounds) {
        GSP_CUDA_TRY(cudaMemsetAsync(changed.ptr, 0, sizeof(int), s));
        for (int r = 0; r < kCheckEvery && rounds < max_rounds; ++r, ++rounds) {
            sssp_relax_kernel<<<grid, kThreads, 0, s>>>(n, g->indptr, g->indices, d_weights, S, d_dist, changed.ptr);
            GSP_CHECK_LAUNCH();
        }
        // the flag covers the last kCheckEvery sweeps; a clean group means the previous state was already a fixpoint
        int host_changed = 0;
        GSP_CUDA_TRY(cudaMemcpyAsync(&host_changed, changed.ptr, sizeof(int), cudaMemcpyDeviceToHost, s));
        GSP_CUDA_TRY(cudaStreamSynchronize(s));
        if (!host_changed) break;
    }
    if (rounds_out) *rounds_out = rounds;
    return GSP_OK;
}

GSP_API int gsp_sssp_batch(const gsp_graph* gg, const double* d_weights, int64_t src_begin, int32_t src_count,
                           double* d_dist, int32_t max_rounds, int32_t* rounds_out, void* stream) {
    GSP_REQUIRE(gg != nullptr, "graph is NULL");
    const Graph* g = reinterpret_cast<const Graph*>(gg);
    GSP_REQUIRE(src_begin >= 0 && src_count >= 0 && src_begin + src_count <= g->n, "source range outside [0, num_nodes]");
    return sssp_run(g, d_weights, src_begin, nullptr, src_count, d_dist, max_rounds, rounds_out, as_stream(stream));
}

GSP_API int gsp_sssp_sources(const gsp_graph* gg, const double* d_weights, const int32_t* d_sources, int32_t src_count,
                             double* d_dist, int32_t max_rounds, int32_t* rounds_out, void* stream) {
    GSP_REQUIRE(gg != nullptr, "graph is NULL");
    GSP_REQUIRE(src_count >= 0 && (src_count == 0 || d_sources != nullptr), "source list is NULL");
    return sssp_run(reinterpret_cast<const Graph*>(gg), d_weights, 0, d_sources, src_count, d_dist, max_rounds, rounds_out,
                    as_stream(stream));
}
