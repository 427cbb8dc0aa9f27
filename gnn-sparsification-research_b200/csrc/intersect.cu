// intersect.cu — Jaccard / Adamic-Adar / degree-product edge scoring (kernel families K2, K3).
//
// Replaces the SciPy SpGEMM formulation of reference src/sparsification/metrics.py:43-64 (Jaccard,
// `(Ab @ Ab)[u, v]`) and :99-121 (Adamic-Adar, `(W W^T)[u, v]`) by direct intersection of the two
// sorted CSR neighbour lists of every canonical edge. The SpGEMM materialises every 2-hop pair
// (work sum_w d_w^2, memory nnz(A^2)); the per-edge intersection touches only the two lists.
//
// Work distribution: persistent CTAs (a multiple of the 148 SMs) whose warps claim chunks of
// consecutive canonical edges from a global counter, so a hub row's edges are spread over the whole
// chip while consecutive edges (same row u) still share row(u) in L1.
// Per edge the shorter list is streamed 32 ids per step (one coalesced request) and every lane
// lower-bounds its id in the longer list; the search window shrinks monotonically because both lists
// are sorted.
#include <cstdlib>

#include "common.cuh"

namespace gsp {
namespace {

constexpr int kThreads = 256;
constexpr int kWarpsPerBlock = kThreads / kWarp;
constexpr int kChunk = 8;  // consecutive edges claimed per counter bump

struct GraphView {
    const int64_t* indptr;
    const int32_t* indices;
    const int32_t* rows;
    const int64_t* bptr;   // list B of edge (u, v): column v of A for Jaccard (== indptr when symmetric)
    const int32_t* bidx;
};

__device__ __forceinline__ int lower_bound(const int32_t* __restrict__ list, int lo, int hi, int32_t x) {
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (__ldg(list + mid) < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// kMode 0: Jaccard (count + one fp64 divide).  kMode 1: Adamic-Adar (ordered fp64 accumulation).
template <int kMode>
__global__ void __launch_bounds__(kThreads)
intersect_kernel(GraphView g, int64_t e_begin, int64_t e_end, const double* __restrict__ node_w,
                 int32_t* __restrict__ inter_out, double* __restrict__ score_out, unsigned long long* counter) {
    const int lane = lane_id();
    for (;;) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(counter, (unsigned long long)kChunk);
        base = __shfl_sync(0xffffffffu, base, 0);
        int64_t e0 = e_begin + (int64_t)base;
        if (e0 >= e_end) break;
        int64_t e1 = e0 + kChunk < e_end ? e0 + kChunk : e_end;
        for (int64_t e = e0; e < e1; ++e) {
            const int32_t u = __ldg(g.rows + e), v = __ldg(g.indices + e);
            const int64_t a0 = __ldg(g.indptr + u), a1 = __ldg(g.indptr + u + 1);
            const int64_t b0 = __ldg(g.bptr + v), b1 = __ldg(g.bptr + v + 1);
            const int la = (int)(a1 - a0), lb = (int)(b1 - b0);
            const int32_t* S = g.indices + a0;
            const int32_t* L = g.bidx + b0;
            int ls = la, ll = lb;
            if (lb < la) { S = g.bidx + b0; L = g.indices + a0; ls = lb; ll = la; }

            if (kMode == 0) {
                int count = 0, lo = 0;
                for (int base_i = 0; base_i < ls && lo < ll; base_i += kWarp) {
                    int i = base_i + lane;
                    int pos = ll;
                    if (i < ls) {
                        int32_t x = __ldg(S + i);
                        pos = lower_bound(L, lo, ll, x);
                        count += (pos < ll && __ldg(L + pos) == x);
                    }
                    // ids after this step are larger than the step's last id: its bound is the new floor
                    int last = min(kWarp - 1, ls - 1 - base_i);
                    lo = __shfl_sync(0xffffffffu, pos, last);
                }
                count = __reduce_add_sync(0xffffffffu, count);
                if (lane == 0) {
                    // both degrees are ROW degrees of the binarised matrix (metrics.py:44,50)
                    const double du = (double)la;
                    const double dv = (double)(__ldg(g.indptr + v + 1) - __ldg(g.indptr + v));
                    const double uni = du + dv - (double)count;
                    if (inter_out) inter_out[e - e_begin] = count;
                    score_out[e - e_begin] = uni > 0.0 ? __ddiv_rn((double)count, uni) : 0.0;
                }
            } else {
                // common ids must be visited in DESCENDING order (SciPy SpGEMM's accumulation order)
                double acc = 0.0;
                int hi = ll;
                for (int base_i = 0; base_i < ls && hi > 0; base_i += kWarp) {
                    int i = ls - 1 - (base_i + lane);
                    int pos = 0;
                    bool hit = false;
                    double term = 0.0;
                    if (i >= 0) {
                        int32_t x = __ldg(S + i);
                        pos = lower_bound(L, 0, hi, x);
                        hit = pos < hi && __ldg(L + pos) == x;
                        if (hit) {
                            double w = __ldg(node_w + x);
                            term = __dmul_rn(w, w);
                        }
                    }
                    unsigned hits = __ballot_sync(0xffffffffu, hit);
                    while (hits) {  // lanes ascending == ids descending
                        int src = __ffs(hits) - 1;
                        acc = __dadd_rn(acc, __shfl_sync(0xffffffffu, term, src));
                        hits &= hits - 1;
                    }
                    int last = min(kWarp - 1, ls - 1 - base_i);
                    hi = __shfl_sync(0xffffffffu, pos, last);  // smaller ids lie below the smallest id's bound
                }
                if (lane == 0) score_out[e - e_begin] = acc;
            }
        }
    }
}

__global__ void aa_weights_kernel(int64_t n, const int64_t* __restrict__ indptr, double* __restrict__ w) {
    for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u < n; u += (int64_t)gridDim.x * blockDim.x) {
        double deg = (double)(indptr[u + 1] - indptr[u]);
        double lg = fmax(log(deg + 1.0), 1e-10);            // metrics.py:104-105
        w[u] = __ddiv_rn(1.0, __dsqrt_rn(lg));              // metrics.py:108
    }
}

__global__ void aa_weights_table_kernel(int64_t n, const int64_t* __restrict__ indptr, const double* __restrict__ table,
                                        int64_t table_len, double* __restrict__ w) {
    for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u < n; u += (int64_t)gridDim.x * blockDim.x) {
        const int64_t d = indptr[u + 1] - indptr[u];
        w[u] = table[d < table_len ? d : table_len - 1];
    }
}

__global__ void weighted_degree_kernel(int64_t n, const int64_t* __restrict__ indptr, const double* __restrict__ data,
                                       double* __restrict__ deg) {
    for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u < n; u += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        if (data) {
            for (int64_t p = indptr[u]; p < indptr[u + 1]; ++p) s = __dadd_rn(s, data[p]);
        } else {
            s = (double)(indptr[u + 1] - indptr[u]);
        }
        deg[u] = s;
    }
}

__global__ void degree_product_kernel(int64_t e_begin, int64_t e_end, const int32_t* __restrict__ rows,
                                      const int32_t* __restrict__ indices, const double* __restrict__ deg,
                                      double* __restrict__ out) {
    for (int64_t e = e_begin + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < e_end; e += (int64_t)gridDim.x * blockDim.x)
        out[e - e_begin] = __dmul_rn(deg[rows[e]], deg[indices[e]]);
}

int check_range(const Graph* g, int64_t e_begin, int64_t e_end) {
    GSP_REQUIRE(g != nullptr, "graph is NULL");
    GSP_REQUIRE(e_begin >= 0 && e_begin <= e_end && e_end <= g->nnz, "edge range outside [0, nnz]");
    return GSP_OK;
}

// GSP_INTERSECT=general forces the per-edge search kernel (tests cover both schedules)
bool use_owner_path(const Graph* g) {
    if (!g->symmetric) return false;
    const char* env = getenv("GSP_INTERSECT");
    return !(env && env[0] == 'g');
}

template <int kMode>
int launch_intersect(const Graph* g, int64_t e_begin, int64_t e_end, const double* node_w, int32_t* inter, double* score,
                     cudaStream_t s) {
    if (use_owner_path(g)) {
        Graph* gm = const_cast<Graph*>(g);
        return kMode == 0 ? owner_intersect_jaccard(gm, e_begin, e_end, 0, g->n, inter, score, s)
                          : owner_intersect_adamic_adar(gm, e_begin, e_end, 0, g->n, node_w, score, s);
    }
    GraphView view{g->indptr, g->indices, g->rows, g->indptr, g->indices};
    if (kMode == 0 && !g->symmetric) {
        view.bptr = g->tptr;
        view.bidx = g->tidx;
    }
    Scratch<unsigned long long> counter;
    GSP_CUDA_TRY(counter.alloc(1, s));
    GSP_CUDA_TRY(cudaMemsetAsync(counter.ptr, 0, sizeof(unsigned long long), s));
    const int64_t chunks = (e_end - e_begin + kChunk - 1) / kChunk;
    const int grid = grid_for(chunks, kWarpsPerBlock, 8);
    intersect_kernel<kMode><<<grid, kThreads, 0, s>>>(view, e_begin, e_end, node_w, inter, score, counter.ptr);
    GSP_CHECK_LAUNCH();
    return GSP_OK;
}

}  // namespace
}  // namespace gsp

using namespace gsp;

GSP_API int gsp_jaccard(const gsp_graph* gg, int64_t e_begin, int64_t e_end, int32_t* d_inter, double* d_score,
                        void* stream) {
    const Graph* g = reinterpret_cast<const Graph*>(gg);
    if (int rc = check_range(g, e_begin, e_end)) return rc;
    if (e_end == e_begin) return GSP_OK;
    GSP_REQUIRE(d_score != nullptr, "d_score is NULL");
    return launch_intersect<0>(g, e_begin, e_end, nullptr, d_inter, d_score, as_stream(stream));
}

GSP_API int gsp_aa_node_weights(const gsp_graph* gg, double* d_node_w, void* stream) {
    GSP_REQUIRE(gg && d_node_w, "NULL argument");
    const Graph* g = reinterpret_cast<const Graph*>(gg);
    if (g->n == 0) return GSP_OK;
    aa_weights_kernel<<<grid_for(g->n, 256), 256, 0, as_stream(stream)>>>(g->n, g->indptr, d_node_w);
    GSP_CHECK_LAUNCH();
    return GSP_OK;
}

// dst[i] = src[i] with SM loads: `src` may be page-locked HOST memory (unified addressing), which lets a small table reach
// the device while the copy engine is busy with a large upload queued earlier
__global__ void copy_f64_kernel(int64_t n, const double* __restrict__ src, double* __restrict__ dst) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

GSP_API int gsp_copy_f64(const double* src, double* d_dst, int64_t count, void* stream) {
    GSP_REQUIRE(count >= 0 && (count == 0 || (src && d_dst)), "NULL argument");
    if (count == 0) return GSP_OK;
    copy_f64_kernel<<<grid_for(count, 256, 2), 256, 0, as_stream(stream)>>>(count, src, d_dst);
    GSP_CHECK_LAUNCH();
    return GSP_OK;
}

GSP_API int gsp_aa_node_weights_from_table(const gsp_graph* gg, const double* d_table, int64_t table_len, double* d_node_w,
                                           void* stream) {
    GSP_REQUIRE(gg && d_table && d_node_w, "NULL argument");
    const Graph* g = reinterpret_cast<const Graph*>(gg);
    GSP_REQUIRE(table_len > g->max_degree, "table must have max_degree + 1 entries");
    if (g->n == 0) return GSP_OK;
    aa_weights_table_kernel<<<grid_for(g->n, 256), 256, 0, as_stream(stream)>>>(g->n, g->indptr, d_table, table_len, d_node_w);
    GSP_CHECK_LAUNCH();
    return GSP_OK;
}

GSP_API int gsp_adamic_adar(const gsp_graph* gg, const double* d_node_w, int64_t e_begin, int64_t e_end, double* d_score,
                            void* stream) {
    const Graph* g = reinterpret_cast<const Graph*>(gg);
    if (int rc = check_range(g, e_begin, e_end)) return rc;
    if (e_end == e_begin) return GSP_OK;
    GSP_REQUIRE(d_score != nullptr, "d_score is NULL");
    cudaStream_t s = as_stream(stream);
    Scratch<double> w;
    if (!d_node_w) {
        GSP_CUDA_TRY(w.alloc(g->n, s));
        if (int rc = gsp_aa_node_weights(gg, w.ptr, stream)) return rc;
        d_node_w = w.ptr;
    }
    return launch_intersect<1>(g, e_begin, e_end, d_node_w, nullptr, d_score, s);
}

GSP_API int gsp_jaccard_adamic_adar(const gsp_graph* gg, const double* d_node_w, int64_t e_begin, int64_t e_end,
                                    int32_t* d_inter, double* d_jaccard, double* d_adamic_adar, void* stream) {
    const Graph* g = reinterpret_cast<const Graph*>(gg);
    if (int rc = check_range(g, e_begin, e_end)) return rc;
    if (e_end == e_begin) return GSP_OK;
    GSP_REQUIRE(d_jaccard != nullptr && d_adamic_adar != nullptr, "output is NULL");
    if (!use_owner_path(g)) {
        // asymmetric pattern (Jaccard intersects row(u) with col(v), Adamic-Adar with row(v)) or the forced general
        // schedule: two passes
        if (int rc = gsp_jaccard(gg, e_begin, e_end, d_inter, d_jaccard, stream)) return rc;
        return gsp_adamic_adar(gg, d_node_w, e_begin, e_end, d_adamic_adar, stream);
    }
    cudaStream_t s = as_stream(stream);
    Scratch<double> w;
    if (!d_node_w) {
        GSP_CUDA_TRY(w.alloc(g->n, s));
        if (int rc = gsp_aa_node_weights(gg, w.ptr, stream)) return rc;
        d_node_w = w.ptr;
    }
    return owner_intersect_both(const_cast<Graph*>(g), e_begin, e_end, 0, g->n, d_node_w, d_inter, d_jaccard, d_adamic_adar, s);
}

GSP_API int gsp_degree_product(const gsp_graph* gg, int64_t e_begin, int64_t e_end, double* d_score, void* stream) {
    const Graph* g = reinterpret_cast<const Graph*>(gg);
    if (int rc = check_range(g, e_begin, e_end)) return rc;
    if (e_end == e_begin) return GSP_OK;
    GSP_REQUIRE(d_score != nullptr, "d_score is NULL");
    cudaStream_t s = as_stream(stream);
    Scratch<double> deg;
    GSP_CUDA_TRY(deg.alloc(g->n, s));
    weighted_degree_kernel<<<grid_for(g->n, 256), 256, 0, s>>>(g->n, g->indptr, g->data, deg.ptr);
    GSP_CHECK_LAUNCH();
    degree_product_kernel<<<grid_for(e_end - e_begin, 256), 256, 0, s>>>(e_begin, e_end, g->rows, g->indices, deg.ptr, d_score);
    GSP_CHECK_LAUNCH();
    return GSP_OK;
}

// ---- owner-sharded variants (multi-GPU): every undirected pair is evaluated on exactly one rank -----------------
static int check_owned(const Graph* g, int64_t node_begin, int64_t node_end, const void* out) {
    GSP_REQUIRE(g != nullptr, "graph is NULL");
    GSP_REQUIRE(node_begin >= 0 && node_begin <= node_end && node_end <= g->n, "owner range outside [0, num_nodes]");
    GSP_REQUIRE(g->nnz == 0 || out != nullptr, "output is NULL");
    if (!g->symmetric) {
        set_error("owner-sharded scoring needs a symmetric adjacency pattern; shard by edge range instead");
        return GSP_ERR_UNSUPPORTED;
    }
    return GSP_OK;
}

GSP_API int gsp_jaccard_owned(const gsp_graph* gg, int64_t node_begin, int64_t node_end, int32_t* d_inter_full,
                              double* d_score_full, void* stream) {
    const Graph* g = reinterpret_cast<const Graph*>(gg);
    if (int rc = check_owned(g, node_begin, node_end, d_score_full)) return rc;
    return owner_intersect_jaccard(const_cast<Graph*>(g), 0, g->nnz, node_begin, node_end, d_inter_full, d_score_full,
                                   as_stream(stream), true);
}

GSP_API int gsp_adamic_adar_owned(const gsp_graph* gg, const double* d_node_w, int64_t node_begin, int64_t node_end,
                                  double* d_score_full, void* stream) {
    const Graph* g = reinterpret_cast<const Graph*>(gg);
    if (int rc = check_owned(g, node_begin, node_end, d_score_full)) return rc;
    cudaStream_t s = as_stream(stream);
    Scratch<double> w;
    if (!d_node_w) {
        GSP_CUDA_TRY(w.alloc(g->n, s));
        if (int rc = gsp_aa_node_weights(gg, w.ptr, stream)) return rc;
        d_node_w = w.ptr;
    }
    return owner_intersect_adamic_adar(const_cast<Graph*>(g), 0, g->nnz, node_begin, node_end, d_node_w, d_score_full, s, true);
}

GSP_API int gsp_jaccard_adamic_adar_owned(const gsp_graph* gg, const double* d_node_w, int64_t node_begin, int64_t node_end,
                                          double* d_jaccard_full, double* d_adamic_adar_full, void* stream) {
    const Graph* g = reinterpret_cast<const Graph*>(gg);
    if (int rc = check_owned(g, node_begin, node_end, d_jaccard_full)) return rc;
    if (int rc = check_owned(g, node_begin, node_end, d_adamic_adar_full)) return rc;
    cudaStream_t s = as_stream(stream);
    Scratch<double> w;
    if (!d_node_w) {
        GSP_CUDA_TRY(w.alloc(g->n, s));
        if (int rc = gsp_aa_node_weights(gg, w.ptr, stream)) return rc;
        d_node_w = w.ptr;
    }
    return owner_intersect_both(const_cast<Graph*>(g), 0, g->nnz, node_begin, node_end, d_node_w, nullptr, d_jaccard_full,
                                d_adamic_adar_full, s, true);
}

GSP_API int gsp_graph_set_owner_deal(gsp_graph* gg, const uint8_t* d_owner_rank, int32_t rank, void* stream) {
    Graph* g = reinterpret_cast<Graph*>(gg);
    GSP_REQUIRE(g != nullptr, "graph is NULL");
    GSP_REQUIRE(d_owner_rank == nullptr || (rank >= 0 && rank < 256), "rank must be in [0, 256)");
    return set_owner_deal(g, d_owner_rank, rank, as_stream(stream));
}

// Peer-scatter variants: the scoring kernel itself delivers every score to the rank that owns its position
// (d_slices[k] = base of rank k's slice of `slice_len` positions, possibly peer memory mapped over NVLink).
static int check_scatter(const Graph* g, int64_t node_begin, int64_t node_end, double* const* d_slices, int32_t world,
                         int64_t slice_len) {
    if (int rc = check_owned(g, node_begin, node_end, d_slices)) return rc;
    GSP_REQUIRE(world >= 1 && slice_len >= 1 && (int64_t)world * slice_len >= g->nnz, "slices do not cover the positions");
    return GSP_OK;
}

GSP_API int gsp_jaccard_owned_scatter(const gsp_graph* gg, int64_t node_begin, int64_t node_end, double* const* d_slices,
                                      int32_t world, int64_t slice_len, void* stream) {
    const Graph* g = reinterpret_cast<const Graph*>(gg);
    if (int rc = check_scatter(g, node_begin, node_end, d_slices, world, slice_len)) return rc;
    return owner_intersect_scatter(const_cast<Graph*>(g), 0, node_begin, node_end, nullptr, d_slices, slice_len, as_stream(stream));
}

GSP_API int gsp_adamic_adar_owned_scatter(const gsp_graph* gg, const double* d_node_w, int64_t node_begin, int64_t node_end,
                                          double* const* d_slices, int32_t world, int64_t slice_len, void* stream) {
    const Graph* g = reinterpret_cast<const Graph*>(gg);
    if (int rc = check_scatter(g, node_begin, node_end, d_slices, world, slice_len)) return rc;
    cudaStream_t s = as_stream(stream);
    Scratch<double> w;
    if (!d_node_w) {
        GSP_CUDA_TRY(w.alloc(g->n, s));
        if (int rc = gsp_aa_node_weights(gg, w.ptr, stream)) return rc;
        d_node_w = w.ptr;
    }
    return owner_intersect_scatter(const_cast<Graph*>(g), 1, node_begin, node_end, d_node_w, d_slices, slice_len, s);
}

GSP_API int gsp_jaccard_adamic_adar_owned_scatter(const gsp_graph* gg, const double* d_node_w, int64_t node_begin,
                                                  int64_t node_end, double* const* d_jaccard_slices,
                                                  double* const* d_adamic_adar_slices, int32_t world, int64_t slice_len,
                                                  void* stream) {
    const Graph* g = reinterpret_cast<const Graph*>(gg);
    if (int rc = check_scatter(g, node_begin, node_end, d_jaccard_slices, world, slice_len)) return rc;
    if (int rc = check_scatter(g, node_begin, node_end, d_adamic_adar_slices, world, slice_len)) return rc;
    cudaStream_t s = as_stream(stream);
    Scratch<double> w;
    if (!d_node_w) {
        GSP_CUDA_TRY(w.alloc(g->n, s));
        if (int rc = gsp_aa_node_weights(gg, w.ptr, stream)) return rc;
        d_node_w = w.ptr;
    }
    return owner_intersect_scatter(const_cast<Graph*>(g), 2, node_begin, node_end, d_node_w, d_adamic_adar_slices, slice_len, s,
                                   d_jaccard_slices);
}

GSP_API int gsp_owner_costs(const gsp_graph* gg, double* d_cost, void* stream) {
    GSP_REQUIRE(gg && d_cost, "NULL argument");
    return owner_costs(reinterpret_cast<const Graph*>(gg), d_cost, as_stream(stream));
}
