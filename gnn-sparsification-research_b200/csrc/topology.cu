// topology.cu — output-graph topology metrics on the canonical CSR (a "next" row).
//
// The reference evaluates them through NetworkX (src/sparsification/metrics.py:445-520 `compute_topology_metrics`):
// `nx.average_clustering`, `nx.connected_components`. Both fall out of structures the engine already has:
//   * triangles through a node = half the sum of the intersection counts |N(v) ∩ N(w)| over its edges — the counts
//     the Jaccard pass returns (gsp_jaccard's d_inter), corrected for self loops (NetworkX drops the node itself from
//     its neighbour set);
//   * connected components = min-label hooking over the edge list with pointer jumping (Shiloach-Vishkin style):
//     O(log n) sweeps also on chain-like graphs whose diameter is in the thousands.
#include "common.cuh"

namespace gsp {
namespace {

constexpr int kThreads = 256;

// loop[v] = 1 when the row of v holds v itself (rows are sorted: binary search)
__global__ void loop_flags_kernel(int64_t n, const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                  uint8_t* __restrict__ loop) {
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
        int64_t lo = indptr[v], hi = indptr[v + 1];
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (indices[mid] < (int32_t)v) lo = mid + 1; else hi = mid;
        }
        loop[v] = (lo < indptr[v + 1] && indices[lo] == (int32_t)v) ? 1 : 0;
    }
}

// warp per row: pairs[v] = sum over neighbours w != v of (|row(v) ∩ row(w)| - loop[v] - loop[w]) = 2 * triangles(v);
// degree[v] = neighbours other than v
__global__ void __launch_bounds__(kThreads)
node_triangles_kernel(int64_t n, const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                      const int32_t* __restrict__ inter, const uint8_t* __restrict__ loop, int64_t* __restrict__ pairs,
                      int32_t* __restrict__ degree) {
    const int lane = lane_id();
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t v = warp; v < n; v += nwarps) {
        const int64_t p0 = indptr[v], p1 = indptr[v + 1];
        const int lv = loop[v];
        long long sum = 0;
        for (int64_t p = p0 + lane; p < p1; p += kWarp) {
            const int32_t w = indices[p];
            if (w != (int32_t)v) sum += (long long)inter[p] - lv - loop[w];
        }
        for (int off = 16; off; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
        if (lane == 0) {
            pairs[v] = sum;
            degree[v] = (int32_t)(p1 - p0) - lv;
        }
    }
}

__global__ void init_labels_kernel(int64_t n, int32_t* __restrict__ label) {
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) label[v] = (int32_t)v;
}

// hook the larger of the two endpoint labels under the smaller one (labels are node ids: label[x] <= x always)
__global__ void hook_kernel(int64_t nnz, const int32_t* __restrict__ rows, const int32_t* __restrict__ indices, int32_t* label,
                            int* __restrict__ changed) {
    bool any = false;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nnz; p += (int64_t)gridDim.x * blockDim.x) {
        const int32_t a = label[rows[p]], b = label[indices[p]];
        if (a == b) continue;
        atomicMin(label + max(a, b), min(a, b));
        any = true;
    }
    if (any) *changed = 1;
}

// pointer jumping: every node ends up pointing at the root of its tree (a node with label[r] == r)
__global__ void compress_kernel(int64_t n, int32_t* label) {
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
        int32_t c = label[v];
        while (true) {
            const int32_t up = label[c];
            if (up == c) break;
            c = up;
        }
        label[v] = c;
    }
}

}  // namespace
}  // namespace gsp

using namespace gsp;

static int check_symmetric(const Graph* g, const char* what) {
    GSP_REQUIRE(g != nullptr, "graph is NULL");
    if (!g->symmetric) {
        set_error("%s needs a symmetric (undirected) adjacency pattern", what);
        return GSP_ERR_UNSUPPORTED;
    }
    return GSP_OK;
}

GSP_API int gsp_node_triangles(const gsp_graph* gg, const int32_t* d_inter, int64_t* d_pairs, int32_t* d_degree, void* stream) {
    const Graph* g = reinterpret_cast<const Graph*>(gg);
    if (int rc = check_symmetric(g, "gsp_node_triangles")) return rc;
    if (g->n == 0) return GSP_OK;
    GSP_REQUIRE(d_pairs && d_degree && (g->nnz == 0 || d_inter), "NULL argument");
    cudaStream_t s = as_stream(stream);
    Scratch<uint8_t> loop;
    GSP_CUDA_TRY(loop.alloc(g->n, s));
    loop_flags_kernel<<<grid_for(g->n, kThreads), kThreads, 0, s>>>(g->n, g->indptr, g->indices, loop.ptr);
    GSP_CHECK_LAUNCH();
    node_triangles_kernel<<<grid_for(g->n, kThreads / kWarp, 8), kThreads, 0, s>>>(g->n, g->indptr, g->indices, d_inter, loop.ptr,
                                                                                  d_pairs, d_degree);
    GSP_CHECK_LAUNCH();
    return GSP_OK;
}

GSP_API int gsp_connected_components(const gsp_graph* gg, int32_t* d_label, int32_t* rounds_out, void* stream) {
    const Graph* g = reinterpret_cast<const Graph*>(gg);
    if (int rc = check_symmetric(g, "gsp_connected_components")) return rc;
    if (rounds_out) *rounds_out = 0;
    if (g->n == 0) return GSP_OK;
    GSP_REQUIRE(d_label != nullptr, "d_label is NULL");
    cudaStream_t s = as_stream(stream);
    init_labels_kernel<<<grid_for(g->n, kThreads), kThreads, 0, s>>>(g->n, d_label);
    GSP_CHECK_LAUNCH();
    if (g->nnz == 0) return GSP_OK;
    Scratch<int> changed;
    GSP_CUDA_TRY(changed.alloc(1, s));
    int rounds = 0;
    for (;; ++rounds) {
        GSP_REQUIRE(rounds < 4096, "connected components did not converge");
        GSP_CUDA_TRY(cudaMemsetAsync(changed.ptr, 0, sizeof(int), s));
        hook_kernel<<<grid_for(g->nnz, kThreads), kThreads, 0, s>>>(g->nnz, g->rows, g->indices, d_label, changed.ptr);
        GSP_CHECK_LAUNCH();
        compress_kernel<<<grid_for(g->n, kThreads), kThreads, 0, s>>>(g->n, d_label);
        GSP_CHECK_LAUNCH();
        int host_changed = 0;
        GSP_CUDA_TRY(cudaMemcpyAsync(&host_changed, changed.ptr, sizeof(int), cudaMemcpyDeviceToHost, s));
        GSP_CUDA_TRY(cudaStreamSynchronize(s));
        if (!host_changed) break;
    }
    if (rounds_out) *rounds_out = rounds + 1;
    return GSP_OK;
}
