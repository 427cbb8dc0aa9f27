// intersect_owner.cu — owner-hashed neighbour intersection for symmetric graphs (fast path of K2 / K3).
//
// Same results as intersect.cu (reference src/sparsification/metrics.py:43-64 and :99-121), different
// schedule. On a power-law graph almost every edge touches a high-degree endpoint, so per-edge searches
// of the short list in the long one cost sum_e min(d)*log(max(d)) dependent probes. Here every undirected
// pair {o, w} is evaluated ONCE, at its OWNER o = the endpoint with the larger degree (ties: smaller id):
// the owner's neighbour set is put into a shared-memory hash table once and the (shorter) list of every
// owned neighbour w is streamed through it with coalesced loads — sum over pairs of min(d) element tests,
// one smem probe each. While row(w) streams by, the position of o inside it is seen as well, so the
// score is written to both directed positions (o,w) and (w,o) (intersection, degrees and the
// descending-id accumulation order are all symmetric in the pair).
//
//   level A  rows with 1 <= deg <= 64: one warp per owner, 128-slot table per warp, neighbours in registers
//   level B  rows with deg > 64: CTA per (owner, 1024-neighbour chunk); the owner row is hashed in tiles of
//            8192 ids (16384 slots), tiles visited in DESCENDING id order so Adamic-Adar keeps SciPy's
//            accumulation order; per-neighbour partial state lives in shared memory between tiles
//
// An edge range [e_begin, e_end) (multi-GPU sharding) restricts the pairs to those with a directed position
// inside the range; only in-range positions are written.
#include <climits>
#include <mutex>

#include "common.cuh"

namespace gsp {
namespace {

constexpr int kWarpOwnerMax = 64;          // level A / level B split
constexpr int kATableSlots = 128;          // per-warp hash slots (load factor <= 0.5)
constexpr int kAThreads = 256;
constexpr int kAWarps = kAThreads / kWarp;
constexpr int kARowsPerClaim = 16;

constexpr int kBThreads = 512;
constexpr int kBWarps = kBThreads / kWarp;
constexpr int kBTile = 8192;               // owner ids per hash tile
constexpr int kBSlots = 2 * kBTile;
constexpr int kBChunk = 1024;              // neighbours per work item

__device__ __forceinline__ uint32_t hash_id(int32_t x) { return (uint32_t)x * 0x9E3779B1u; }

__device__ __forceinline__ void hash_insert(int32_t* slots, uint32_t mask, int shift, int32_t x) {
    uint32_t h = hash_id(x) >> shift;
    for (;;) {
        int32_t prev = atomicCAS(&slots[h], -1, x);
        if (prev == -1 || prev == x) return;
        h = (h + 1) & mask;
    }
}

__device__ __forceinline__ bool hash_contains(const int32_t* slots, uint32_t mask, int shift, int32_t x) {
    uint32_t h = hash_id(x) >> shift;
    for (;;) {
        int32_t s = slots[h];
        if (s == x) return true;
        if (s == -1) return false;
        h = (h + 1) & mask;
    }
}

// the pair {o, w} is evaluated at o unless w has the larger degree (ties: smaller id owns)
__device__ __forceinline__ bool other_owns(int d_w, int32_t w, int d_o, int32_t o) {
    return d_w > d_o || (d_w == d_o && w < o);
}

__device__ __forceinline__ int lower_bound_i32(const int32_t* __restrict__ list, int n, int32_t x) {
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (__ldg(list + mid) < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

struct RangeInfo {
    int64_t e_begin, e_end;
    int32_t row_lo, row_hi;   // rows holding the first / last position of the range
    bool full;
};

// Stream row(w)[s, e) through the hash table. kMode 0: returns the number of hits. kMode 1: continues the
// ordered fp64 accumulation (ids visited in descending order). rev receives the offset of `o` in row(w).
template <int kMode>
__device__ __forceinline__ void stream_row(const int32_t* __restrict__ row_w, int s, int e, const int32_t* slots,
                                           uint32_t mask, int shift, int32_t o, const double* __restrict__ node_w,
                                           int& count, double& acc, int& rev) {
    const int lane = lane_id();
    if (kMode == 0) {
        int c = 0;
        for (int base = s; base < e; base += kWarp) {
            const int i = base + lane;
            if (i < e) {
                const int32_t x = __ldg(row_w + i);
                c += hash_contains(slots, mask, shift, x);
                if (x == o) rev = i;
            }
        }
        count += __reduce_add_sync(0xffffffffu, c);
    } else {
        for (int base = e - 1; base >= s; base -= kWarp) {
            const int i = base - lane;   // lanes ascending == ids descending
            bool hit = false;
            double term = 0.0;
            if (i >= s) {
                const int32_t x = __ldg(row_w + i);
                hit = hash_contains(slots, mask, shift, x);
                if (x == o) rev = i;
                if (hit) {
                    const double w = __ldg(node_w + x);
                    term = __dmul_rn(w, w);
                }
            }
            unsigned hits = __ballot_sync(0xffffffffu, hit);
            while (hits) {
                const int src = __ffs(hits) - 1;
                acc = __dadd_rn(acc, __shfl_sync(0xffffffffu, term, src));
                hits &= hits - 1;
            }
        }
    }
    rev = __reduce_max_sync(0xffffffffu, rev);
}

template <int kMode>
__device__ __forceinline__ void write_pair(const RangeInfo& r, int64_t p1, int64_t p2, int d_o, int d_w, int count,
                                           double acc, int32_t* __restrict__ inter_out, double* __restrict__ score_out) {
    double score;
    if (kMode == 0) {
        const double uni = (double)d_o + (double)d_w - (double)count;
        score = uni > 0.0 ? __ddiv_rn((double)count, uni) : 0.0;
    } else {
        score = acc;
    }
    if (p1 >= r.e_begin && p1 < r.e_end) {
        score_out[p1 - r.e_begin] = score;
        if (kMode == 0 && inter_out) inter_out[p1 - r.e_begin] = count;
    }
    if (p2 != p1 && p2 >= r.e_begin && p2 < r.e_end) {
        score_out[p2 - r.e_begin] = score;
        if (kMode == 0 && inter_out) inter_out[p2 - r.e_begin] = count;
    }
}

// ---- level A: one warp per low-degree owner --------------------------------------------------------------
template <int kMode>
__global__ void __launch_bounds__(kAThreads)
warp_owner_kernel(int64_t n, const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, RangeInfo r,
                  const double* __restrict__ node_w, int32_t* __restrict__ inter_out, double* __restrict__ score_out,
                  unsigned long long* counter) {
    __shared__ int32_t tables[kAWarps][kATableSlots];
    const int lane = lane_id();
    int32_t* slots = tables[threadIdx.x >> 5];
    constexpr uint32_t mask = kATableSlots - 1;
    constexpr int shift = 32 - 7;
    for (;;) {
        unsigned long long first = 0;
        if (lane == 0) first = atomicAdd(counter, (unsigned long long)kARowsPerClaim);
        first = __shfl_sync(0xffffffffu, first, 0);
        if ((int64_t)first >= n) break;
        const int64_t last = min((int64_t)first + kARowsPerClaim, n);
        for (int64_t o64 = (int64_t)first; o64 < last; ++o64) {
            const int32_t o = (int32_t)o64;
            const int64_t a0 = __ldg(indptr + o), a1 = __ldg(indptr + o + 1);
            const int d_o = (int)(a1 - a0);
            if (d_o == 0 || d_o > kWarpOwnerMax) continue;
            const bool row_in_range = r.full || (a0 < r.e_end && a1 > r.e_begin);
            // neighbours in registers: lane holds elements lane and lane + 32
            const int32_t nb0 = lane < d_o ? __ldg(indices + a0 + lane) : INT_MAX;
            const int32_t nb1 = lane + 32 < d_o ? __ldg(indices + a0 + lane + 32) : INT_MAX;
            if (!row_in_range) {  // only neighbours whose own row meets the range matter: any of them?
                const bool any0 = lane < d_o && nb0 >= r.row_lo && nb0 <= r.row_hi;
                const bool any1 = lane + 32 < d_o && nb1 >= r.row_lo && nb1 <= r.row_hi;
                if (!__any_sync(0xffffffffu, any0 || any1)) continue;
            }
            __syncwarp();
            for (int i = lane; i < kATableSlots; i += kWarp) slots[i] = -1;
            __syncwarp();
            if (lane < d_o) hash_insert(slots, mask, shift, nb0);
            if (lane + 32 < d_o) hash_insert(slots, mask, shift, nb1);
            __syncwarp();
            for (int j = 0; j < d_o; ++j) {
                const int32_t w = __shfl_sync(0xffffffffu, j < 32 ? nb0 : nb1, j & 31);
                const int64_t b0 = __ldg(indptr + w);
                const int d_w = (int)(__ldg(indptr + w + 1) - b0);
                if (other_owns(d_w, w, d_o, o)) continue;
                const int64_t p1 = a0 + j;
                if (!r.full) {
                    const bool p1_in = p1 >= r.e_begin && p1 < r.e_end;
                    const bool w_in = b0 < r.e_end && b0 + d_w > r.e_begin;
                    if (!p1_in && !w_in) continue;
                }
                int count = 0, rev = -1;
                double acc = 0.0;
                stream_row<kMode>(indices + b0, 0, d_w, slots, mask, shift, o, node_w, count, acc, rev);
                if (lane == 0 && rev >= 0) write_pair<kMode>(r, p1, b0 + rev, d_o, d_w, count, acc, inter_out, score_out);
            }
        }
    }
}

// ---- level B: CTA per (owner, neighbour chunk), hash tiles in shared memory ----------------------------------
struct OwnerItem {
    int32_t owner;
    int32_t first;  // index of the chunk's first neighbour inside the owner's row
};

template <int kMode>
__global__ void __launch_bounds__(kBThreads)
cta_owner_kernel(const OwnerItem* __restrict__ items, int64_t num_items, const int64_t* __restrict__ indptr,
                 const int32_t* __restrict__ indices, RangeInfo r, const double* __restrict__ node_w,
                 int32_t* __restrict__ inter_out, double* __restrict__ score_out, unsigned long long* counter) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int32_t* slots = reinterpret_cast<int32_t*>(smem_raw);                          // [kBSlots]
    double* acc_s = reinterpret_cast<double*>(smem_raw + sizeof(int32_t) * kBSlots);  // [kBChunk]
    int32_t* cnt_s = reinterpret_cast<int32_t*>(acc_s + kBChunk);                   // [kBChunk]
    int32_t* rev_s = cnt_s + kBChunk;                                               // [kBChunk]
    __shared__ long long item_s;
    __shared__ int next_s;
    const int lane = lane_id();
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) item_s = (long long)atomicAdd(counter, 1ull);
        __syncthreads();
        const int64_t item = item_s;
        if (item >= num_items) break;
        const int32_t o = items[item].owner;
        const int j0 = items[item].first;
        const int64_t a0 = __ldg(indptr + o);
        const int d_o = (int)(__ldg(indptr + o + 1) - a0);
        const int nb = min(kBChunk, d_o - j0);
        const int32_t* row_o = indices + a0;
        if (!r.full) {
            const int64_t c0 = a0 + j0, c1 = c0 + nb;
            const bool chunk_in_range = c0 < r.e_end && c1 > r.e_begin;
            if (!chunk_in_range) {
                const int32_t w_first = __ldg(row_o + j0), w_last = __ldg(row_o + j0 + nb - 1);
                if (w_last < r.row_lo || w_first > r.row_hi) continue;
            }
        }
        for (int i = threadIdx.x; i < nb; i += kBThreads) {
            acc_s[i] = 0.0;
            cnt_s[i] = 0;
            rev_s[i] = -1;
        }
        const int num_tiles = (d_o + kBTile - 1) / kBTile;
        for (int t = num_tiles - 1; t >= 0; --t) {
            const int ts = t * kBTile, te = min(d_o, ts + kBTile);
            int cap = 64;
            while (cap < 2 * (te - ts)) cap <<= 1;
            const uint32_t mask = (uint32_t)cap - 1u;
            const int shift = 32 - (31 - __clz(cap));
            __syncthreads();  // previous tile's probes are done
            for (int i = threadIdx.x; i < cap; i += kBThreads) slots[i] = -1;
            if (threadIdx.x == 0) next_s = 0;
            __syncthreads();
            for (int i = ts + threadIdx.x; i < te; i += kBThreads) hash_insert(slots, mask, shift, __ldg(row_o + i));
            __syncthreads();
            const int32_t lo_id = t == 0 ? INT_MIN : __ldg(row_o + ts);
            const bool last_tile = t == num_tiles - 1;
            const int32_t hi_id = last_tile ? INT_MAX : __ldg(row_o + te);
            for (;;) {
                int i = 0;
                if (lane == 0) i = atomicAdd(&next_s, 1);
                i = __shfl_sync(0xffffffffu, i, 0);
                if (i >= nb) break;
                const int32_t w = __ldg(row_o + j0 + i);
                const int64_t b0 = __ldg(indptr + w);
                const int d_w = (int)(__ldg(indptr + w + 1) - b0);
                if (other_owns(d_w, w, d_o, o)) continue;
                if (!r.full) {
                    const int64_t p1 = a0 + j0 + i;
                    const bool p1_in = p1 >= r.e_begin && p1 < r.e_end;
                    const bool w_in = b0 < r.e_end && b0 + d_w > r.e_begin;
                    if (!p1_in && !w_in) continue;
                }
                const int32_t* row_w = indices + b0;
                const int s = t == 0 ? 0 : lower_bound_i32(row_w, d_w, lo_id);
                const int e = last_tile ? d_w : lower_bound_i32(row_w, d_w, hi_id);
                int count = 0, rev = -1;
                double acc = kMode == 1 ? acc_s[i] : 0.0;
                stream_row<kMode>(row_w, s, e, slots, mask, shift, o, node_w, count, acc, rev);
                if (lane == 0) {
                    if (kMode == 0) cnt_s[i] += count; else acc_s[i] = acc;
                    if (rev >= 0) rev_s[i] = rev;
                }
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < nb; i += kBThreads) {
            const int rev = rev_s[i];
            if (rev < 0) continue;  // pair owned by the neighbour, or outside the range
            const int32_t w = __ldg(row_o + j0 + i);
            const int64_t b0 = __ldg(indptr + w);
            const int d_w = (int)(__ldg(indptr + w + 1) - b0);
            write_pair<kMode>(r, a0 + j0 + i, b0 + rev, d_o, d_w, cnt_s[i], acc_s[i], inter_out, score_out);
        }
    }
}

constexpr size_t kBSmemBytes = sizeof(int32_t) * kBSlots + sizeof(double) * kBChunk + 2 * sizeof(int32_t) * kBChunk;

// ---- work items for level B (built once per graph) -----------------------------------------------------------
__global__ void count_items_kernel(int64_t n, const int64_t* __restrict__ indptr, int64_t* __restrict__ counts) {
    for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) {
        const int64_t d = indptr[o + 1] - indptr[o];
        counts[o] = d > kWarpOwnerMax ? (d + kBChunk - 1) / kBChunk : 0;
    }
}

__global__ void fill_items_kernel(int64_t n, const int64_t* __restrict__ indptr, const int64_t* __restrict__ incl,
                                  OwnerItem* __restrict__ items) {
    for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) {
        const int64_t d = indptr[o + 1] - indptr[o];
        if (d <= kWarpOwnerMax) continue;
        const int64_t c = (d + kBChunk - 1) / kBChunk;
        OwnerItem* dst = items + (incl[o] - c);
        for (int64_t k = 0; k < c; ++k) dst[k] = OwnerItem{(int32_t)o, (int32_t)(k * kBChunk)};
    }
}

__global__ void range_rows_kernel(const int32_t* __restrict__ rows, int64_t e_begin, int64_t e_end, int32_t* out) {
    out[0] = rows[e_begin];
    out[1] = rows[e_end - 1];
}

std::mutex g_items_mutex;

int ensure_items(Graph* g, cudaStream_t s) {
    std::lock_guard<std::mutex> lock(g_items_mutex);
    if (g->owner_items_ready) return GSP_OK;
    Scratch<int64_t> counts, incl;
    GSP_CUDA_TRY(counts.alloc(g->n, s));
    GSP_CUDA_TRY(incl.alloc(g->n, s));
    count_items_kernel<<<grid_for(g->n, 256), 256, 0, s>>>(g->n, g->indptr, counts.ptr);
    GSP_CHECK_LAUNCH();
    if (int rc = inclusive_sum_i64(counts.ptr, incl.ptr, g->n, s)) return rc;
    int64_t total = 0;
    GSP_CUDA_TRY(cudaMemcpyAsync(&total, incl.ptr + (g->n - 1), sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    GSP_CUDA_TRY(cudaStreamSynchronize(s));
    if (total > 0) {
        OwnerItem* items = nullptr;
        GSP_CUDA_TRY(cudaMalloc(&items, (size_t)total * sizeof(OwnerItem)));
        fill_items_kernel<<<grid_for(g->n, 256), 256, 0, s>>>(g->n, g->indptr, incl.ptr, items);
        GSP_CHECK_LAUNCH();
        GSP_CUDA_TRY(cudaStreamSynchronize(s));
        g->owner_items = items;
    }
    g->num_owner_items = total;
    g->owner_items_ready = true;
    return GSP_OK;
}

template <int kMode>
int launch(Graph* g, int64_t e_begin, int64_t e_end, const double* node_w, int32_t* inter, double* score, cudaStream_t s) {
    if (g->n == 0 || e_end == e_begin) return GSP_OK;
    if (int rc = ensure_items(g, s)) return rc;
    RangeInfo r{e_begin, e_end, 0, 0, e_begin == 0 && e_end == g->nnz};
    if (!r.full) {
        Scratch<int32_t> rr;
        GSP_CUDA_TRY(rr.alloc(2, s));
        range_rows_kernel<<<1, 1, 0, s>>>(g->rows, e_begin, e_end, rr.ptr);
        GSP_CHECK_LAUNCH();
        int32_t h[2];
        GSP_CUDA_TRY(cudaMemcpyAsync(h, rr.ptr, sizeof(h), cudaMemcpyDeviceToHost, s));
        GSP_CUDA_TRY(cudaStreamSynchronize(s));
        r.row_lo = h[0];
        r.row_hi = h[1];
    }
    Scratch<unsigned long long> counters;
    GSP_CUDA_TRY(counters.alloc(2, s));
    GSP_CUDA_TRY(cudaMemsetAsync(counters.ptr, 0, 2 * sizeof(unsigned long long), s));
    GSP_CUDA_TRY(cudaFuncSetAttribute(cta_owner_kernel<kMode>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBSmemBytes));
    // hubs first: their long work items should not land in the tail
    if (g->num_owner_items > 0) {
        int64_t blocks = g->num_owner_items < 2ll * kNumSMs ? g->num_owner_items : 2ll * kNumSMs;
        cta_owner_kernel<kMode><<<(int)blocks, kBThreads, kBSmemBytes, s>>>(
            reinterpret_cast<const OwnerItem*>(g->owner_items), g->num_owner_items, g->indptr, g->indices, r, node_w, inter,
            score, counters.ptr);
        GSP_CHECK_LAUNCH();
    }
    const int64_t claims = (g->n + kARowsPerClaim - 1) / kARowsPerClaim;
    warp_owner_kernel<kMode><<<grid_for(claims, kAWarps, 8), kAThreads, 0, s>>>(g->n, g->indptr, g->indices, r, node_w, inter,
                                                                            score, counters.ptr + 1);
    GSP_CHECK_LAUNCH();
    return GSP_OK;
}

}  // namespace

int owner_intersect_jaccard(Graph* g, int64_t e_begin, int64_t e_end, int32_t* inter, double* score, cudaStream_t s) {
    return launch<0>(g, e_begin, e_end, nullptr, inter, score, s);
}

int owner_intersect_adamic_adar(Graph* g, int64_t e_begin, int64_t e_end, const double* node_w, double* score,
                                cudaStream_t s) {
    return launch<1>(g, e_begin, e_end, node_w, nullptr, score, s);
}

}  // namespace gsp
