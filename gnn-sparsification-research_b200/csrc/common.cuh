// common.cuh — shared helpers for libgsp.so (sm_100a). Internal; the public surface is include/gsp.h.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/gsp.h"

#define GSP_API extern "C" __attribute__((visibility("default")))

namespace gsp {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this
constexpr int kWarp = 32;

void set_error(const char* fmt, ...);
void count_launch();

#define GSP_CUDA_TRY(expr)                                                                              \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess) {                                                                        \
            gsp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return GSP_ERR_CUDA;                                                                        \
        }                                                                                               \
    } while (0)

#define GSP_REQUIRE(cond, msg)                                  \
    do {                                                        \
        if (!(cond)) {                                          \
            gsp::set_error("invalid argument: %s", msg);        \
            return GSP_ERR_INVALID;                             \
        }                                                       \
    } while (0)

// Placed after EVERY kernel launch: checks the launch and counts it (gsp_launch_count feeds bench.py's gpu_launches).
#define GSP_CHECK_LAUNCH()                  \
    do {                                    \
        gsp::count_launch();                \
        GSP_CUDA_TRY(cudaGetLastError());   \
    } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// graph.cu: the library's private stream-ordered pool on the current device (created on first use). It keeps a bounded
// amount of freed scratch across synchronisation points (GSP_SCRATCH_KEEP_MB, default 1 GiB) and releases the rest to the
// driver; gsp_trim_scratch releases all of it. nullptr: pool creation failed, use the device's default pool.
cudaMemPool_t scratch_pool();

// Device memory of the library (graph arrays, lazily built side structures, scratch). A host may plug its own allocator
// in (gsp_set_allocator: the Python layer hands over torch's caching allocator), so that nothing the library holds is
// invisible to the host framework and repeated graph builds reuse blocks instead of paying cudaMalloc / cudaFree of
// gigabytes every time; without one: cudaMalloc / cudaFree for persistent arrays, the private pool for scratch.
void* device_alloc_bytes(size_t bytes, cudaStream_t s);   // nullptr on failure
void device_free_bytes(void* p);
bool custom_allocator();
template <typename T>
inline cudaError_t device_alloc(T** out, size_t bytes, cudaStream_t s) {
    *out = static_cast<T*>(device_alloc_bytes(bytes, s));
    return *out ? cudaSuccess : cudaErrorMemoryAllocation;
}
inline cudaError_t device_free(void* p) {
    device_free_bytes(p);
    return cudaSuccess;
}

// Stream-ordered scratch buffer (plugged allocator, else the private pool): freed on the same stream when it goes out of scope.
template <typename T>
struct Scratch {
    T* ptr = nullptr;
    cudaStream_t stream = nullptr;
    cudaError_t alloc(size_t count, cudaStream_t s) {
        stream = s;
        const size_t bytes = (count ? count : 1) * sizeof(T);
        if (custom_allocator()) {
            hosted = true;
            ptr = static_cast<T*>(device_alloc_bytes(bytes, s));
            return ptr ? cudaSuccess : cudaErrorMemoryAllocation;
        }
        if (cudaMemPool_t pool = scratch_pool()) return cudaMallocFromPoolAsync(reinterpret_cast<void**>(&ptr), bytes, pool, s);
        return cudaMallocAsync(reinterpret_cast<void**>(&ptr), bytes, s);
    }
    bool hosted = false;   // from the plugged allocator (its free is ordered on the allocation stream, like cudaFreeAsync)
    ~Scratch() {
        if (!ptr) return;
        if (hosted) device_free_bytes(ptr);
        else cudaFreeAsync(ptr, stream);
    }
    Scratch() = default;
    Scratch(const Scratch&) = delete;
    Scratch& operator=(const Scratch&) = delete;
};

// Library-owned canonical graph. Everything lives on `device`.
struct Graph {
    int device = 0;
    int64_t n = 0;
    int64_t num_input_edges = 0;
    int64_t nnz = 0;
    int64_t num_undirected = 0;
    int64_t max_degree = 0;
    double sum_degree_sq = 0.0;
    bool symmetric = false;
    bool input_canonical = false;
    bool unit_weights = true;
    int64_t* indptr = nullptr;   // [n+1]
    int32_t* indices = nullptr;  // [nnz] column ids, ascending inside a row
    int32_t* rows = nullptr;     // [nnz] row id per canonical position
    double* data = nullptr;      // [nnz] merged values (multiplicities); nullptr when unit_weights
    // transpose pattern (only when !symmetric): row v of A^T == column v of A
    int64_t* tptr = nullptr;
    int32_t* tidx = nullptr;
    int32_t* rev_off = nullptr;  // [nnz] offset of u inside row(v) for position (u,v) (-1 when (v,u) is absent)
    // lazily built
    int32_t* und_id = nullptr;   // [nnz]
    void* seg_items = nullptr;   // row segments of the Laplacian SpMM (approx_er.cu)
    int64_t* seg_incl = nullptr; // [n] inclusive prefix of segments per row
    int64_t num_seg_items = 0;
    void* owner_items = nullptr; // work items of the owner-hashed intersection (intersect_owner.cu)
    int64_t num_owner_items = 0; // medium-class items (first in the array)
    int64_t num_hub_items = 0;   // hub-class items (after them)
    bool owner_items_ready = false;
    // the work items whose owner lies in [owned_lo, owned_hi) (owner-sharded scoring: a rank would otherwise claim and
    // skip the items of all other ranks, 7/8 of ~6 x 10^5 claims on 8 GPUs); rebuilt when the range changes
    void* owned_items = nullptr;
    int64_t num_owned_items = 0, num_owned_hub_items = 0;
    int64_t owned_lo = -1, owned_hi = -1;
    bool owned_dealt = false;
    // dealt ownership (gsp_graph_set_owner_deal): owner o is evaluated by this process only when deal[o] == deal_rank
    uint8_t* deal = nullptr;
    int32_t deal_rank = -1;
};

// graph.cu: CUB inclusive scan wrapper shared by the lazily built side structures
int inclusive_sum_i64(const int64_t* in, int64_t* out, int64_t count, cudaStream_t s);
int sort_pairs_u32_u64(const uint32_t* keys_in, uint32_t* keys_out, const uint64_t* vals_in, uint64_t* vals_out, int64_t count,
                       cudaStream_t s);
// intersect_owner.cu: fast path for symmetric graphs (each undirected pair evaluated once at its owner)
// `dealt`: honour the graph's owner deal (gsp_graph_set_owner_deal) — the *_owned entry points only
int owner_intersect_jaccard(Graph* g, int64_t e_begin, int64_t e_end, int64_t owner_lo, int64_t owner_hi, int32_t* inter,
                            double* score, cudaStream_t s, bool dealt = false);
int owner_intersect_adamic_adar(Graph* g, int64_t e_begin, int64_t e_end, int64_t owner_lo, int64_t owner_hi,
                                const double* node_w, double* score, cudaStream_t s, bool dealt = false);
int set_owner_deal(Graph* g, const uint8_t* d_owner_rank, int32_t rank, cudaStream_t s);
int owner_costs(const Graph* g, double* cost, cudaStream_t s);
int owner_intersect_both(Graph* g, int64_t e_begin, int64_t e_end, int64_t owner_lo, int64_t owner_hi, const double* node_w,
                         int32_t* inter, double* jaccard, double* adamic_adar, cudaStream_t s, bool dealt = false);
int owner_intersect_scatter(Graph* g, int mode, int64_t owner_lo, int64_t owner_hi, const double* node_w,
                            double* const* slices, int64_t slice_len, cudaStream_t s, double* const* slices2 = nullptr);

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// Order-preserving map fp64 -> uint64 (ascending). -0.0 and +0.0 are made equal first because IEEE
// comparison (what argsort uses) treats them as ties.
__device__ __forceinline__ uint64_t ordered_key(double s) {
    if (s == 0.0) s = 0.0;
    uint64_t b = static_cast<uint64_t>(__double_as_longlong(s));
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

inline int grid_for(int64_t work_items, int per_block, int max_blocks_per_sm = 16) {
    int64_t blocks = (work_items + per_block - 1) / per_block;
    int64_t cap = static_cast<int64_t>(kNumSMs) * max_blocks_per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return static_cast<int>(blocks);
}

}  // namespace gsp
