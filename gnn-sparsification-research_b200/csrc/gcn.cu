// gcn.cu — GCN normalisation and propagation of the kept sub-graph (a "next" row: the consumer of the
// sparsifier's output).
//
// The reference trains GCN / GCN* on the sparsified graph with `GCNConv(..., cached=False, normalize=True)` and the
// optional "-W" edge weights (src/models/gnn.py:222-223,244; scripts/nb05_roman_empire/roman_empire_gpu.py:258-281), so
// every layer of every forward pass re-runs torch_geometric's `gcn_norm` and a scatter-add propagate over the same
// kept edges. These entry points do that work once, on the edge list the engine has just compacted:
//   gsp_gcn_norm       add_remaining_self_loops + D^-1/2 A D^-1/2 (torch_geometric/nn/conv/gcn_conv.py `gcn_norm`,
//                      flow = source_to_target: degrees are sums over TARGET nodes)
//   gsp_target_order   stable grouping of edges by target (indptr + permutation)
//   gsp_gcn_propagate  out[t] = sum over the edges into t, in edge order, of w_e * x[source_e]   (fp32, multiply then add)
// Sums run sequentially in edge order, which is what a CPU scatter_add does, so results are reproducible run to run
// (torch_geometric on a GPU accumulates with atomics in arbitrary order).
#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace gsp {
namespace {

constexpr int kThreads = 256;
constexpr int kLongRowDefault = 2048;   // target rows with more edges get a CTA each (GSP_GCN_LONG_ROW overrides, for tuning)
constexpr int kStageEdges = 128;
constexpr int kStageFeatures = 128;

__global__ void flag_non_loops_kernel(int64_t e, const int64_t* __restrict__ row, const int64_t* __restrict__ col,
                                      int64_t* __restrict__ flags, long long* __restrict__ last_loop) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < e; i += (int64_t)gridDim.x * blockDim.x) {
        const bool loop = row[i] == col[i];
        flags[i] = loop ? 0 : 1;
        if (loop) atomicMax(last_loop + row[i], (long long)i);   // the LAST loop edge of a node keeps its weight
    }
}

__global__ void fill_i64_kernel(int64_t count, long long value, long long* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) out[i] = value;
}

__global__ void list_long_rows_kernel(int64_t n, const int64_t* __restrict__ indptr, int long_row, int64_t* __restrict__ rows,
                                      unsigned long long* __restrict__ count) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
        if (indptr[t + 1] - indptr[t] > long_row) rows[atomicAdd(count, 1ull)] = t;
}

// non-loop edges keep their order at the front; one loop per node follows (weight: its last existing loop, else 1)
__global__ void emit_edges_kernel(int64_t n, int64_t e, const int64_t* __restrict__ row, const int64_t* __restrict__ col,
                                  const float* __restrict__ w, const int64_t* __restrict__ incl,
                                  const long long* __restrict__ last_loop, int64_t* __restrict__ out_row,
                                  int64_t* __restrict__ out_col, float* __restrict__ out_w, int64_t* __restrict__ out_count) {
    const int64_t kept = e ? incl[e - 1] : 0;
    const int64_t start = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = start; i < e; i += stride) {
        if (row[i] == col[i]) continue;
        const int64_t p = incl[i] - 1;
        out_row[p] = row[i];
        out_col[p] = col[i];
        out_w[p] = w ? w[i] : 1.0f;
    }
    for (int64_t v = start; v < n; v += stride) {
        out_row[kept + v] = v;
        out_col[kept + v] = v;
        const long long l = last_loop[v];
        out_w[kept + v] = (l >= 0 && w) ? w[l] : 1.0f;
    }
    if (start == 0 && out_count) *out_count = kept + n;
}

// deg[t] = (fp32 sum of the non-loop weights into t, in edge order) + loop weight; dinv = deg^-1/2, inf -> 0
__device__ __forceinline__ float inv_sqrt_degree(float deg) {
    float r = __fdiv_rn(1.0f, __fsqrt_rn(deg));
    return isinf(r) ? 0.0f : r;
}

__global__ void degree_kernel(int64_t n, int64_t kept, const int64_t* __restrict__ indptr, const int64_t* __restrict__ perm,
                              const float* __restrict__ w, int long_row, float* __restrict__ dinv) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        if (indptr[t + 1] - indptr[t] > long_row) continue;   // degree_long_kernel's rows
        float deg = 0.0f;
        for (int64_t k = indptr[t]; k < indptr[t + 1]; ++k) deg = __fadd_rn(deg, w[perm[k]]);
        dinv[t] = inv_sqrt_degree(__fadd_rn(deg, w[kept + t]));
    }
}

// Long rows: a warp fetches 256 weights at a time (eight gathers per lane in flight) and adds them in order.
__global__ void __launch_bounds__(kThreads)
degree_long_kernel(const int64_t* __restrict__ long_rows, const unsigned long long* __restrict__ num_long, int64_t kept,
                   const int64_t* __restrict__ indptr, const int64_t* __restrict__ perm, const float* __restrict__ w,
                   float* __restrict__ dinv) {
    const int lane = lane_id();
    const unsigned long long total = *num_long;
    const unsigned long long nwarps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    for (unsigned long long item = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) >> 5; item < total; item += nwarps) {
        const int64_t t = long_rows[item];
        const int64_t p0 = indptr[t], p1 = indptr[t + 1];
        float deg = 0.0f;
        for (int64_t base = p0; base < p1; base += 8 * kWarp) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int64_t p = base + u * kWarp + lane;
                v[u] = p < p1 ? w[perm[p]] : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int cnt = (int)min((int64_t)kWarp, p1 - (base + u * kWarp));
                for (int k = 0; k < cnt; ++k) deg = __fadd_rn(deg, __shfl_sync(0xffffffffu, v[u], k));
            }
        }
        if (lane == 0) dinv[t] = inv_sqrt_degree(__fadd_rn(deg, w[kept + t]));
    }
}

__global__ void scale_weights_kernel(int64_t count, const int64_t* __restrict__ row, const int64_t* __restrict__ col,
                                     const float* __restrict__ dinv, float* __restrict__ w) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
        w[i] = __fmul_rn(__fmul_rn(dinv[row[i]], w[i]), dinv[col[i]]);
}

__global__ void target_keys_kernel(int64_t e, const int64_t* __restrict__ col, uint32_t* __restrict__ keys,
                                   uint64_t* __restrict__ vals, int64_t* __restrict__ counts) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < e; i += (int64_t)gridDim.x * blockDim.x) {
        keys[i] = (uint32_t)col[i];
        vals[i] = (uint64_t)i;
        atomicAdd(reinterpret_cast<unsigned long long*>(counts + col[i]), 1ull);
    }
}

__global__ void shift_indptr_kernel(int64_t n, const int64_t* __restrict__ incl, int64_t* __restrict__ indptr) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i <= n; i += (int64_t)gridDim.x * blockDim.x)
        indptr[i] = i ? incl[i - 1] : 0;
}

// A warp takes 32 consecutive target rows at a time (dynamic claims): their edges are one contiguous stretch of the
// target-ordered stream, so the edge metadata (perm -> source, weight) is fetched for 32 EDGES at a time across the row
// boundaries (one dependent-load chain per 32 edges instead of one per row; the next chunk's perm is fetched a step
// ahead), and the gathered feature rows are fetched eight at a time. Lane j owns features 4j..4j+3 of a 128-feature strip
// (kVec) or features j, j+32, ... (scalar); the sum over the edges of a row stays sequential.
constexpr int kRowsPerClaim = 32;
constexpr int kGathers = 8;

template <bool kVec>
__global__ void __launch_bounds__(kThreads)
propagate_kernel(int64_t n, const int64_t* __restrict__ indptr, const int64_t* __restrict__ perm,
                 const int64_t* __restrict__ row, const float* __restrict__ w, const float* __restrict__ x, int d, int64_t ldx,
                 float* __restrict__ out, int64_t ldo, int long_row, unsigned long long* __restrict__ counter) {
    const int lane = lane_id();
    const int strip = kVec ? 128 : 32;
    const int64_t num_claims = (n + kRowsPerClaim - 1) / kRowsPerClaim;
    for (;;) {
        unsigned long long claim = 0;
        if (lane == 0) claim = atomicAdd(counter, 1ull);
        claim = __shfl_sync(0xffffffffu, claim, 0);
        if ((int64_t)claim >= num_claims) break;
        const int64_t t0 = (int64_t)claim * kRowsPerClaim, t1 = min(n, t0 + kRowsPerClaim);
        const int64_t a = indptr[min(t0 + lane, n)];          // lane i: start of row t0 + i
        const int64_t p_end = indptr[t1];
        for (int f0 = 0; f0 < d; f0 += strip) {
            const int f = f0 + (kVec ? 4 * lane : lane);
            int64_t chunk = -(int64_t)kWarp - 1;              // first stream position held in (src, we); none yet
            int64_t src = 0, ahead = 0;                       // ahead: perm of the chunk after the current one
            float we = 0.f;
            for (int r = 0; r < (int)(t1 - t0); ++r) {
                const int64_t s0 = __shfl_sync(0xffffffffu, a, r);
                const int64_t s1 = r + 1 < kRowsPerClaim ? __shfl_sync(0xffffffffu, a, (r + 1) & 31) : p_end;
                const int64_t e1 = (t0 + r + 1 == t1) ? p_end : s1;
                if (e1 - s0 > long_row) continue;            // propagate_long_kernel's rows
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int64_t p = s0; p < e1;) {
                    if (p >= chunk + kWarp || p < chunk) {   // fetch the metadata of the next 32 edges of the stream
                        int64_t e = 0;
                        const bool contiguous = p == chunk + kWarp;
                        chunk = p;
                        if (chunk + lane < p_end) e = contiguous ? ahead : perm[chunk + lane];
                        if (chunk + kWarp + lane < p_end) ahead = perm[chunk + kWarp + lane];
                        if (chunk + lane < p_end) {
                            src = row[e];
                            we = w[e];
                        }
                    }
                    const int k_begin = (int)(p - chunk);
                    const int k_end = (int)(min(e1, chunk + kWarp) - chunk);
                    for (int k0 = k_begin; k0 < k_end; k0 += kGathers) {   // gathers in flight, adds in edge order
                        float4 v[kGathers];
                        float ws[kGathers];
#pragma unroll
                        for (int u = 0; u < kGathers; ++u) {
                            const int64_t sidx = __shfl_sync(0xffffffffu, src, (k0 + u) & 31);
                            ws[u] = __shfl_sync(0xffffffffu, we, (k0 + u) & 31);
                            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (k0 + u < k_end && f < d) {
                                if (kVec) v[u] = __ldg(reinterpret_cast<const float4*>(x + sidx * ldx + f));
                                else v[u].x = __ldg(x + sidx * ldx + f);
                            }
                        }
#pragma unroll
                        for (int u = 0; u < kGathers; ++u) {
                            if (k0 + u >= k_end) break;
                            acc.x = __fadd_rn(acc.x, __fmul_rn(ws[u], v[u].x));
                            if (kVec) {
                                acc.y = __fadd_rn(acc.y, __fmul_rn(ws[u], v[u].y));
                                acc.z = __fadd_rn(acc.z, __fmul_rn(ws[u], v[u].z));
                                acc.w = __fadd_rn(acc.w, __fmul_rn(ws[u], v[u].w));
                            }
                        }
                    }
                    p = chunk + k_end;
                }
                if (f < d) {
                    if (kVec) *reinterpret_cast<float4*>(out + (t0 + r) * ldo + f) = acc;
                    else out[(t0 + r) * ldo + f] = acc.x;
                }
            }
        }
    }
}

// ---- long target rows ------------------------------------------------------------------------------------------
// The sum over the edges into one target is sequential, so a warp that walks a 100 k-edge hub row alone pays one gather
// latency per four edges (measured: the single longest row of the R-MAT-24 kept graph took 110 ms of a 137 ms launch).
// Rows longer than the long-row threshold go to one CTA each: all threads stage the gathered feature rows of 128 edges in shared
// memory (64 loads in flight per thread), then thread f adds feature f over the staged edges in order.

__global__ void __launch_bounds__(kThreads)
propagate_long_kernel(const int64_t* __restrict__ long_rows, const unsigned long long* __restrict__ num_long,
                      const int64_t* __restrict__ indptr, const int64_t* __restrict__ perm, const int64_t* __restrict__ row,
                      const float* __restrict__ w, const float* __restrict__ x, int d, int64_t ldx, float* __restrict__ out,
                      int64_t ldo) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* tile = reinterpret_cast<float*>(smem_raw);                          // [kStageEdges][kStageFeatures]
    long long* src_s = reinterpret_cast<long long*>(tile + kStageEdges * kStageFeatures);
    float* w_s = reinterpret_cast<float*>(src_s + kStageEdges);
    const unsigned long long total = *num_long;
    for (unsigned long long item = blockIdx.x; item < total; item += gridDim.x) {
        const int64_t t = long_rows[item];
        const int64_t p0 = indptr[t], p1 = indptr[t + 1];
        for (int f0 = 0; f0 < d; f0 += kStageFeatures) {
            const int fw = min(kStageFeatures, d - f0);
            float acc = 0.f;
            for (int64_t base = p0; base < p1; base += kStageEdges) {
                const int cnt = (int)min((int64_t)kStageEdges, p1 - base);
                __syncthreads();                                   // the previous stage has been consumed
                for (int k = threadIdx.x; k < cnt; k += kThreads) {
                    const int64_t e = perm[base + k];
                    src_s[k] = row[e];
                    w_s[k] = w[e];
                }
                __syncthreads();
                const int total_el = cnt * fw;
                const bool full = fw == kStageFeatures;            // the usual strip: shifts instead of divisions
                for (int j0 = threadIdx.x; j0 < total_el; j0 += 16 * kThreads) {
                    float v[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        const int j = j0 + u * kThreads;
                        const int k = full ? j >> 7 : j / fw, f = full ? j & 127 : j % fw;
                        v[u] = j < total_el ? __ldg(x + src_s[k] * ldx + f0 + f) : 0.f;
                    }
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        const int j = j0 + u * kThreads;
                        const int k = full ? j >> 7 : j / fw, f = full ? j & 127 : j % fw;
                        if (j < total_el) tile[k * kStageFeatures + f] = v[u];
                    }
                }
                __syncthreads();
                if ((int)threadIdx.x < fw)
                    for (int k = 0; k < cnt; ++k) acc = __fadd_rn(acc, __fmul_rn(w_s[k], tile[k * kStageFeatures + threadIdx.x]));
            }
            if ((int)threadIdx.x < fw) out[t * ldo + f0 + threadIdx.x] = acc;
        }
    }
}

// Pipelined form for 16-byte aligned feature rows: the gathered rows go global -> shared with cp.async (no registers),
// three 128-edge stages in flight, the edge metadata (perm -> source, weight) runs two stages further ahead, so the
// per-stage cost is the ordered accumulation itself (~1 us) instead of the ~8 memory latencies of the plain version
// (27 ms -> ~2 ms for the longest row of the R-MAT-24 kept graph). One CTA per SM (197 KB of shared memory).
constexpr int kPipeStages = 3;

__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src) {
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gmem_src) : "memory");
}

__global__ void __launch_bounds__(kThreads)
propagate_long_pipelined_kernel(const int64_t* __restrict__ long_rows, const unsigned long long* __restrict__ num_long,
                                const int64_t* __restrict__ indptr, const int64_t* __restrict__ perm,
                                const int64_t* __restrict__ row, const float* __restrict__ w, const float* __restrict__ x, int d,
                                int64_t ldx, float* __restrict__ out, int64_t ldo) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* tile = reinterpret_cast<float*>(smem_raw);                          // [kPipeStages][kStageEdges][kStageFeatures]
    long long* src_s = reinterpret_cast<long long*>(tile + kPipeStages * kStageEdges * kStageFeatures);   // [kPipeStages][kStageEdges]
    float* w_s = reinterpret_cast<float*>(src_s + kPipeStages * kStageEdges);  // [kPipeStages][kStageEdges]
    const int tid = threadIdx.x;
    const unsigned long long total = *num_long;
    for (unsigned long long item = blockIdx.x; item < total; item += gridDim.x) {
        const int64_t t = long_rows[item];
        const int64_t p0 = indptr[t], p1 = indptr[t + 1];
        const int num_stages = (int)((p1 - p0 + kStageEdges - 1) / kStageEdges);
        for (int f0 = 0; f0 < d; f0 += kStageFeatures) {
            const int fw = min(kStageFeatures, d - f0);          // multiple of 4 (caller checked)
            const int pieces = fw >> 2;                          // 16-byte pieces per feature row
            auto stage_count = [&](int st) { return (int)min((int64_t)kStageEdges, p1 - (p0 + (int64_t)st * kStageEdges)); };
            auto load_meta = [&](int st, long long& m_src, float& m_w) {      // threads < kStageEdges
                const int64_t p = p0 + (int64_t)st * kStageEdges + tid;
                if (tid < kStageEdges && st < num_stages && p < p1) {
                    const int64_t e = perm[p];
                    m_src = row[e];
                    m_w = w[e];
                }
            };
            auto store_meta = [&](int st, long long m_src, float m_w) {
                if (tid < kStageEdges) {
                    src_s[(st % kPipeStages) * kStageEdges + tid] = m_src;
                    w_s[(st % kPipeStages) * kStageEdges + tid] = m_w;
                }
            };
            auto issue_tile = [&](int st) {                                   // after the stage's metadata is visible
                if (st < num_stages) {
                    const int cnt = stage_count(st);
                    const long long* ss = src_s + (st % kPipeStages) * kStageEdges;
                    float* tl = tile + (size_t)(st % kPipeStages) * kStageEdges * kStageFeatures;
                    const bool full = pieces == kStageFeatures / 4;           // the usual strip: shifts instead of divisions
                    for (int j = tid; j < cnt * pieces; j += kThreads) {
                        const int k = full ? j >> 5 : j / pieces, c = j - k * pieces;
                        cp_async_16(tl + k * kStageFeatures + 4 * c, x + ss[k] * ldx + f0 + 4 * c);
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");          // one group per stage, empty ones included
            };
            __syncthreads();                                                  // previous row / strip is done with the buffers
            long long m_src = 0;
            float m_w = 0.f;
            for (int st = 0; st < 2; ++st) {                                  // prologue: metadata and tiles of stages 0, 1
                load_meta(st, m_src, m_w);
                store_meta(st, m_src, m_w);
            }
            __syncthreads();
            issue_tile(0);
            issue_tile(1);
            load_meta(2, m_src, m_w);                                         // in flight during the first accumulation
            float acc = 0.f;
            for (int st = 0; st < num_stages; ++st) {
                asm volatile("cp.async.wait_group 1;" ::: "memory");          // stage st has landed (st + 1 may be in flight)
                __syncthreads();
                if (tid < fw) {
                    const int cnt = stage_count(st);
                    const float* tl = tile + (size_t)(st % kPipeStages) * kStageEdges * kStageFeatures + tid;
                    const float* ws = w_s + (st % kPipeStages) * kStageEdges;
#pragma unroll 8
                    for (int k = 0; k < cnt; ++k) acc = __fadd_rn(acc, __fmul_rn(ws[k], tl[k * kStageFeatures]));
                }
                store_meta(st + 2, m_src, m_w);                               // slot of stage st - 1: free since the last barrier
                __syncthreads();                                              // metadata visible; tile st consumed
                issue_tile(st + 2);                                           // ... into the slot of stage st - 1
                load_meta(st + 3, m_src, m_w);
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            if (tid < fw) out[t * ldo + f0 + tid] = acc;
        }
    }
}

// The long-row kernel (one CTA per SM, bound by its ordered accumulation chains) and the streaming kernel (bound by HBM)
// write disjoint rows, so the former runs on a side stream forked from and joined to the caller's stream.
struct SideStream {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};
std::mutex g_side_mutex;
SideStream g_side[64];

int side_stream(SideStream** out) {
    int dev = 0;
    GSP_CUDA_TRY(cudaGetDevice(&dev));
    GSP_REQUIRE(dev >= 0 && dev < 64, "device index out of range");
    std::lock_guard<std::mutex> lock(g_side_mutex);
    SideStream& ss = g_side[dev];
    if (!ss.stream) {
        GSP_CUDA_TRY(cudaStreamCreateWithFlags(&ss.stream, cudaStreamNonBlocking));
        GSP_CUDA_TRY(cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming));
        GSP_CUDA_TRY(cudaEventCreateWithFlags(&ss.join, cudaEventDisableTiming));
    }
    *out = &ss;
    return GSP_OK;
}

int long_row_threshold() {
    static const int value = [] {
        const char* env = getenv("GSP_GCN_LONG_ROW");
        const int v = env ? atoi(env) : kLongRowDefault;
        return v >= 32 ? v : kLongRowDefault;
    }();
    return value;
}

int target_order(int64_t n, int64_t e, const int64_t* col, int64_t* indptr, int64_t* perm, cudaStream_t s) {
    Scratch<int64_t> counts, incl;
    GSP_CUDA_TRY(counts.alloc(n, s));
    GSP_CUDA_TRY(incl.alloc(n, s));
    GSP_CUDA_TRY(cudaMemsetAsync(counts.ptr, 0, (size_t)(n ? n : 1) * sizeof(int64_t), s));
    if (e > 0) {
        Scratch<uint32_t> keys, keys_sorted;
        Scratch<uint64_t> vals;
        GSP_CUDA_TRY(keys.alloc(e, s));
        GSP_CUDA_TRY(keys_sorted.alloc(e, s));
        GSP_CUDA_TRY(vals.alloc(e, s));
        target_keys_kernel<<<grid_for(e, kThreads), kThreads, 0, s>>>(e, col, keys.ptr, vals.ptr, counts.ptr);
        GSP_CHECK_LAUNCH();
        // LSD radix sort is stable: edges of one target keep their order
        if (int rc = sort_pairs_u32_u64(keys.ptr, keys_sorted.ptr, vals.ptr, reinterpret_cast<uint64_t*>(perm), e, s)) return rc;
    }
    if (n > 0) {
        if (int rc = inclusive_sum_i64(counts.ptr, incl.ptr, n, s)) return rc;
    }
    shift_indptr_kernel<<<grid_for(n + 1, kThreads), kThreads, 0, s>>>(n, incl.ptr, indptr);
    GSP_CHECK_LAUNCH();
    return GSP_OK;
}

}  // namespace
}  // namespace gsp

using namespace gsp;

GSP_API int gsp_target_order(int64_t num_nodes, int64_t num_edges, const int64_t* d_col, int64_t* d_indptr, int64_t* d_perm,
                             void* stream) {
    GSP_REQUIRE(num_nodes >= 0 && num_edges >= 0 && num_nodes < (int64_t(1) << 32), "bad sizes");
    GSP_REQUIRE(d_indptr != nullptr && (num_edges == 0 || (d_col && d_perm)), "NULL argument");
    return target_order(num_nodes, num_edges, d_col, d_indptr, d_perm, as_stream(stream));
}

GSP_API int gsp_gcn_norm(int64_t num_nodes, int64_t num_edges, const int64_t* d_row, const int64_t* d_col, const float* d_weight,
                         int64_t* d_out_row, int64_t* d_out_col, float* d_out_weight, int64_t* d_out_count, void* stream) {
    GSP_REQUIRE(num_nodes >= 0 && num_edges >= 0 && num_nodes < (int64_t(1) << 32), "bad sizes");
    GSP_REQUIRE(num_edges == 0 || (d_row && d_col), "edge list is NULL");
    GSP_REQUIRE(num_nodes + num_edges == 0 || (d_out_row && d_out_col && d_out_weight), "output is NULL");
    cudaStream_t s = as_stream(stream);
    const int64_t n = num_nodes, e = num_edges;
    Scratch<int64_t> flags, incl;
    Scratch<long long> last_loop;
    GSP_CUDA_TRY(flags.alloc(e, s));
    GSP_CUDA_TRY(incl.alloc(e, s));
    GSP_CUDA_TRY(last_loop.alloc(n, s));
    fill_i64_kernel<<<grid_for(n, kThreads), kThreads, 0, s>>>(n, -1, last_loop.ptr);
    GSP_CHECK_LAUNCH();
    int64_t kept = 0;
    if (e > 0) {
        flag_non_loops_kernel<<<grid_for(e, kThreads), kThreads, 0, s>>>(e, d_row, d_col, flags.ptr, last_loop.ptr);
        GSP_CHECK_LAUNCH();
        if (int rc = inclusive_sum_i64(flags.ptr, incl.ptr, e, s)) return rc;
        GSP_CUDA_TRY(cudaMemcpyAsync(&kept, incl.ptr + (e - 1), sizeof(int64_t), cudaMemcpyDeviceToHost, s));
        GSP_CUDA_TRY(cudaStreamSynchronize(s));
    }
    emit_edges_kernel<<<grid_for(e > n ? e : n, kThreads), kThreads, 0, s>>>(n, e, d_row, d_col, d_weight, incl.ptr, last_loop.ptr,
                                                                            d_out_row, d_out_col, d_out_weight, d_out_count);
    GSP_CHECK_LAUNCH();
    if (n == 0) return GSP_OK;
    // degrees over the targets of the non-loop edges, in edge order, then the loop term (it sits last in the list)
    Scratch<int64_t> indptr, perm;
    Scratch<float> dinv;
    GSP_CUDA_TRY(indptr.alloc(n + 1, s));
    GSP_CUDA_TRY(perm.alloc(kept, s));
    GSP_CUDA_TRY(dinv.alloc(n, s));
    if (int rc = target_order(n, kept, d_out_col, indptr.ptr, perm.ptr, s)) return rc;
    degree_kernel<<<grid_for(n, kThreads), kThreads, 0, s>>>(n, kept, indptr.ptr, perm.ptr, d_out_weight, long_row_threshold(), dinv.ptr);
    GSP_CHECK_LAUNCH();
    Scratch<int64_t> long_rows;
    Scratch<unsigned long long> num_long;
    GSP_CUDA_TRY(long_rows.alloc(n, s));
    GSP_CUDA_TRY(num_long.alloc(1, s));
    GSP_CUDA_TRY(cudaMemsetAsync(num_long.ptr, 0, sizeof(unsigned long long), s));
    list_long_rows_kernel<<<grid_for(n, kThreads), kThreads, 0, s>>>(n, indptr.ptr, long_row_threshold(), long_rows.ptr, num_long.ptr);
    GSP_CHECK_LAUNCH();
    degree_long_kernel<<<kNumSMs * 2, kThreads, 0, s>>>(long_rows.ptr, num_long.ptr, kept, indptr.ptr, perm.ptr, d_out_weight, dinv.ptr);
    GSP_CHECK_LAUNCH();
    scale_weights_kernel<<<grid_for(kept + n, kThreads), kThreads, 0, s>>>(kept + n, d_out_row, d_out_col, dinv.ptr, d_out_weight);
    GSP_CHECK_LAUNCH();
    return GSP_OK;
}

GSP_API int gsp_gcn_propagate(int64_t num_nodes, const int64_t* d_indptr, const int64_t* d_perm, const int64_t* d_row,
                              const float* d_weight, const float* d_x, int32_t dim, int64_t ldx, float* d_out, int64_t ldo,
                              void* stream) {
    GSP_REQUIRE(num_nodes >= 0 && dim >= 0, "bad sizes");
    if (num_nodes == 0 || dim == 0) return GSP_OK;
    GSP_REQUIRE(d_indptr && d_perm && d_row && d_weight && d_x && d_out, "NULL argument");
    GSP_REQUIRE(ldx >= dim && ldo >= dim, "leading dimension < dim");
    cudaStream_t s = as_stream(stream);
    const bool vec = dim % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(d_x) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(d_out) & 15) == 0;
    Scratch<int64_t> long_rows;
    Scratch<unsigned long long> num_long;     // [0]: number of long rows, [1]: claim counter of the streaming kernel
    GSP_CUDA_TRY(long_rows.alloc(num_nodes, s));
    GSP_CUDA_TRY(num_long.alloc(2, s));
    GSP_CUDA_TRY(cudaMemsetAsync(num_long.ptr, 0, 2 * sizeof(unsigned long long), s));
    list_long_rows_kernel<<<grid_for(num_nodes, kThreads), kThreads, 0, s>>>(num_nodes, d_indptr, long_row_threshold(), long_rows.ptr,
                                                                             num_long.ptr);
    GSP_CHECK_LAUNCH();
    SideStream* side = nullptr;
    if (int rc = side_stream(&side)) return rc;
    GSP_CUDA_TRY(cudaEventRecord(side->fork, s));
    GSP_CUDA_TRY(cudaStreamWaitEvent(side->stream, side->fork, 0));
    if (vec) {
        const size_t smem = (size_t)kPipeStages * (kStageEdges * kStageFeatures * sizeof(float) + kStageEdges * (sizeof(long long) + sizeof(float)));
        GSP_CUDA_TRY(cudaFuncSetAttribute(propagate_long_pipelined_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        propagate_long_pipelined_kernel<<<kNumSMs, kThreads, smem, side->stream>>>(long_rows.ptr, num_long.ptr, d_indptr, d_perm, d_row,
                                                                                  d_weight, d_x, dim, ldx, d_out, ldo);
    } else {
        const size_t smem = (size_t)kStageEdges * kStageFeatures * sizeof(float) + kStageEdges * (sizeof(long long) + sizeof(float));
        GSP_CUDA_TRY(cudaFuncSetAttribute(propagate_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        propagate_long_kernel<<<kNumSMs * 3, kThreads, smem, side->stream>>>(long_rows.ptr, num_long.ptr, d_indptr, d_perm, d_row,
                                                                            d_weight, d_x, dim, ldx, d_out, ldo);
    }
    GSP_CHECK_LAUNCH();
    GSP_CUDA_TRY(cudaEventRecord(side->join, side->stream));
    const int grid = grid_for((num_nodes + kRowsPerClaim - 1) / kRowsPerClaim, kThreads / kWarp, 8);
    if (vec)
        propagate_kernel<true><<<grid, kThreads, 0, s>>>(num_nodes, d_indptr, d_perm, d_row, d_weight, d_x, dim, ldx, d_out, ldo,
                                                        long_row_threshold(), num_long.ptr + 1);
    else
        propagate_kernel<false><<<grid, kThreads, 0, s>>>(num_nodes, d_indptr, d_perm, d_row, d_weight, d_x, dim, ldx, d_out, ldo,
                                                         long_row_threshold(), num_long.ptr + 1);
    GSP_CHECK_LAUNCH();
    GSP_CUDA_TRY(cudaStreamWaitEvent(s, side->join, 0));
    return GSP_OK;
}
