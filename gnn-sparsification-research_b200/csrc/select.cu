// select.cu — global top-k / bottom-k edge selection, degree-aware guarantee, compaction (K8, K9, K10).
//
// Replaces the full `np.argsort(scores)` of reference src/sparsification/core.py:232-240 (threshold /
// inverse-threshold), :421-451 (degree-aware) and the boolean-mask gather `edge_index[:, mask]` (:242) plus the
// "-W" weights of scripts/nb05_roman_empire/roman_empire_gpu.py:248-256.
//
// Selection is an MSB-first radix select on order-preserving 64-bit keys: 6 histogram passes (11,11,11,11,11,9
// bits) find the key t of the num_keep-th element and how many members of its tie class are needed; one counting
// pass and one writing pass then resolve ties by POSITION so the result equals a stable argsort:
// top-k keeps the highest positions of the tie class, bottom-k the lowest. No sort, no host synchronisation; the
// 16 KB histograms are the only data a multi-GPU caller has to all-reduce.
#include "common.cuh"

namespace gsp {
namespace {

constexpr int kThreads = 256;
constexpr int kBins = GSP_SELECT_BINS;
constexpr int kPasses = GSP_SELECT_PASSES;
constexpr int kMaxBlocks = 1184;  // 148 SMs x 8 resident CTAs
constexpr int kSlotWords = kBins + 2;   // sharded select: histogram + (~min, max) of the live keys, one slot per rank
static_assert(kBins == 2048, "pass layout assumes 11-bit digits");

__host__ __device__ __forceinline__ int pass_shift(int p) { return p < 5 ? 53 - 11 * p : 0; }
__host__ __device__ __forceinline__ int pass_bits(int p) { return p < 5 ? 11 : 9; }

struct SelectState {
    uint64_t prefix;      // digits fixed so far, in place
    int64_t remaining;    // rank still to resolve inside the current prefix bucket (1-based); after the last
                          // pass: how many members of the tie class are kept
    int64_t num_keep;
    int32_t keep_lowest;
    int32_t empty;        // num_keep == 0: keep nothing
    // single-GPU fast path (gsp_select_mask / gsp_select_compact; a multi-GPU caller would have to all-reduce these too):
    // the smallest / largest live key of the current pass. When they are equal the boundary bucket is ONE tie class
    // (Jaccard's and feature-cosine's exact zeros hold ~half of all edges), the key is known and the remaining
    // histogram passes exit at once; the smallest key of pass 0 is the best score (an extremum of the "-W" weights).
    int32_t local_only;
    int32_t resolved;
    unsigned long long live_min, live_max;
    unsigned long long best_key;
    int64_t block_ties[kMaxBlocks];
};
static_assert(sizeof(SelectState) <= 64 + 8 * kMaxBlocks, "state layout");

// keep_lowest selects the smallest keys; top-k selects the smallest INVERTED keys, so one code path serves both.
__device__ __forceinline__ uint64_t select_key(double s, int keep_lowest) {
    uint64_t k = ordered_key(s);
    return keep_lowest ? k : ~k;
}

__global__ void begin_kernel(SelectState* st, int64_t num_keep, int keep_lowest, int local_only) {
    st->prefix = 0;
    st->remaining = num_keep;
    st->num_keep = num_keep;
    st->keep_lowest = keep_lowest;
    st->empty = num_keep <= 0;
    st->local_only = local_only;
    st->resolved = 0;
    st->live_min = ~0ull;
    st->live_max = 0ull;
    st->best_key = ~0ull;
}

// One live key into the block histogram. Heavy tie classes (e.g. Jaccard's zeros) put whole warps on one bin: one atomic
// for the warp then; otherwise one per lane. (A two-round leader aggregation for the exponent-level passes, where a warp
// spreads over a handful of bins, cost more issue slots than the conflicts it removed: 1.2 ms instead of 0.65 ms per pass.)
__device__ __forceinline__ void tally_digit(unsigned int* sh, bool live, unsigned int digit) {
    const unsigned live_mask = __ballot_sync(0xffffffffu, live);
    if (live_mask == 0) return;
    const int leader = __ffs(live_mask) - 1;
    const unsigned int lead_digit = __shfl_sync(0xffffffffu, digit, leader);
    const unsigned same = __ballot_sync(0xffffffffu, live && digit == lead_digit);
    if (same == live_mask) {
        if ((threadIdx.x & 31) == leader) atomicAdd(&sh[lead_digit], (unsigned int)__popc(live_mask));
    } else if (live) {
        atomicAdd(&sh[digit], 1u);
    }
}

// kVec: 16-byte loads (scores 16-byte aligned), four per thread in flight.
template <bool kVec>
__global__ void __launch_bounds__(kThreads)
histogram_kernel(const double* __restrict__ scores, int64_t count, const uint8_t* __restrict__ exclude,
                 SelectState* __restrict__ st, int pass, unsigned long long* __restrict__ hist,
                 unsigned long long* __restrict__ ext) {
    if (st->resolved) return;   // the boundary key is known: nothing left to histogram
    __shared__ unsigned int sh[kBins];
    for (int i = threadIdx.x; i < kBins; i += kThreads) sh[i] = 0;
    __syncthreads();
    const int shift = pass_shift(pass), bits = pass_bits(pass);
    const int keep_lowest = st->keep_lowest;
    const bool extrema = st->local_only != 0 || ext != nullptr;
    const uint64_t prefix = st->prefix;
    const int hi_shift = shift + bits;  // bits above the current digit must match the prefix
    const unsigned int digit_mask = (1u << bits) - 1u;
    unsigned long long kmin = ~0ull, kmax = 0ull;
    constexpr int kUnroll = 4;
    constexpr int kPer = kVec ? 2 : 1;   // scores per load
    const int64_t stride = (int64_t)gridDim.x * kThreads * kUnroll * kPer;
    for (int64_t base = blockIdx.x * (int64_t)kThreads * kUnroll * kPer; base < count; base += stride) {
        double v[kUnroll][kPer];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int64_t i = base + ((int64_t)u * kThreads + threadIdx.x) * kPer;
            if (kVec) {
                double2 t = make_double2(0.0, 0.0);
                if (i + 1 < count) t = *reinterpret_cast<const double2*>(scores + i);
                else if (i < count) t.x = scores[i];
                v[u][0] = t.x;
                v[u][kPer - 1] = t.y;
            } else {
                v[u][0] = i < count ? scores[i] : 0.0;
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
#pragma unroll
            for (int e = 0; e < kPer; ++e) {
                const int64_t i = base + ((int64_t)u * kThreads + threadIdx.x) * kPer + e;
                bool live = i < count && !(exclude && exclude[i]);
                unsigned int digit = 0;
                if (live) {
                    const uint64_t k = select_key(v[u][e], keep_lowest);
                    if (hi_shift < 64 && (k >> hi_shift) != (prefix >> hi_shift)) live = false;
                    digit = (unsigned int)(k >> shift) & digit_mask;
                    if (live && extrema) {
                        kmin = k < kmin ? k : kmin;
                        kmax = k > kmax ? k : kmax;
                    }
                }
                tally_digit(sh, live, digit);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (1 << bits); i += kThreads)
        if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
    if (extrema) {
        for (int o = 16; o; o >>= 1) {
            const unsigned long long a = __shfl_xor_sync(0xffffffffu, kmin, o), b = __shfl_xor_sync(0xffffffffu, kmax, o);
            kmin = a < kmin ? a : kmin;
            kmax = b > kmax ? b : kmax;
        }
        if ((threadIdx.x & 31) == 0 && kmin <= kmax) {
            if (ext) {   // sharded caller: the slot travels with the histogram (zero-initialised, so the minimum is stored inverted)
                atomicMax(&ext[0], ~kmin);
                atomicMax(&ext[1], kmax);
            } else {
                atomicMin(&st->live_min, kmin);
                atomicMax(&st->live_max, kmax);
            }
        }
    }
}

// One warp: find the bucket holding the `remaining`-th smallest live key and descend into it.
// `slots` > 0: `hist` holds one [kBins histogram | ~min live key | max live key] slot per rank (all-gathered), summed here.
__global__ void pick_kernel(SelectState* st, const unsigned long long* __restrict__ hist, int pass, int slots) {
    if (st->empty || st->resolved) return;
    const int lane = threadIdx.x;
    const int nsum = slots > 0 ? slots : 1;
    const int stride = slots > 0 ? kSlotWords : 0;
    if (st->local_only || slots > 0) {
        unsigned long long lo = st->live_min, hi = st->live_max;
        if (slots > 0) {
            lo = ~0ull;
            hi = 0ull;
            for (int r = 0; r < slots; ++r) {
                const unsigned long long a = ~hist[(size_t)r * kSlotWords + kBins], b = hist[(size_t)r * kSlotWords + kBins + 1];
                lo = a < lo ? a : lo;
                hi = b > hi ? b : hi;
            }
        }
        __syncwarp();
        if (lane == 0) {
            if (pass == 0) st->best_key = lo;
            st->live_min = ~0ull;          // reset for the next pass
            st->live_max = 0ull;
        }
        if (lo == hi) {                    // every live key is the same: the boundary bucket is one tie class
            if (lane == 0) {
                st->prefix = lo;           // `remaining` already says how many of its members are kept
                st->resolved = 1;
            }
            return;
        }
    }
    const int bits = pass_bits(pass), shift = pass_shift(pass);
    const int nbins = 1 << bits, per = nbins / 32;
    unsigned long long local = 0;
    for (int r = 0; r < nsum; ++r)
        for (int i = 0; i < per; ++i) local += hist[(size_t)r * stride + lane * per + i];
    unsigned long long incl = local;
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const unsigned long long excl = incl - local;
    const unsigned long long total = __shfl_sync(0xffffffffu, incl, 31);
    unsigned long long want = (unsigned long long)st->remaining;
    if (want > total) want = total;  // caller asked for more than there is: keep everything
    if (total == 0) return;
    const bool mine = want > excl && want <= incl;
    if (mine) {
        unsigned long long run = excl;
        for (int i = 0; i < per; ++i) {
            unsigned long long h = 0;
            for (int r = 0; r < nsum; ++r) h += hist[(size_t)r * stride + lane * per + i];
            if (want <= run + h) {
                st->prefix |= (uint64_t)(lane * per + i) << shift;
                st->remaining = (int64_t)(want - run);
                break;
            }
            run += h;
        }
    }
}

__device__ __forceinline__ int64_t chunk_size(int64_t count, int blocks) {
    int64_t c = (count + blocks - 1) / blocks;
    return (c + kThreads - 1) / kThreads * kThreads;
}

__global__ void __launch_bounds__(kThreads)
count_ties_kernel(const double* __restrict__ scores, int64_t count, const uint8_t* __restrict__ exclude,
                  SelectState* st, unsigned long long* total) {
    __shared__ unsigned long long sh;
    if (threadIdx.x == 0) sh = 0;
    __syncthreads();
    const int keep_lowest = st->keep_lowest;
    const uint64_t t = st->prefix;
    const int64_t chunk = chunk_size(count, gridDim.x);
    const int64_t lo = blockIdx.x * chunk, hi = min(lo + chunk, count);
    unsigned long long c = 0;
    if (!st->empty) {
        for (int64_t i0 = lo + threadIdx.x; i0 < hi; i0 += 4 * kThreads) {
            double v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = i0 + u * kThreads < hi ? scores[i0 + u * kThreads] : 0.0;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t i = i0 + u * kThreads;
                c += i < hi && !(exclude && exclude[i]) && select_key(v[u], keep_lowest) == t;
            }
        }
    }
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&sh, c);
    __syncthreads();
    if (threadIdx.x == 0) {
        st->block_ties[blockIdx.x] = (int64_t)sh;
        if (sh) atomicAdd(total, sh);
    }
}

__global__ void __launch_bounds__(kThreads)
write_mask_kernel(const double* __restrict__ scores, int64_t count, const uint8_t* exclude /* may alias mask */,
                  const SelectState* __restrict__ st, const int64_t* __restrict__ ties_before_p,
                  const int64_t* __restrict__ ties_total_p, int or_into, uint8_t* mask) {
    __shared__ long long warp_sums[kThreads / 32];
    __shared__ long long block_base;
    const int keep_lowest = st->keep_lowest;
    const bool empty = st->empty;
    const uint64_t t = st->prefix;
    const int64_t need = st->remaining;
    const int64_t ties_total = *ties_total_p;
    // window of global tie ranks (position order) that are kept
    const int64_t win_lo = keep_lowest ? 0 : ties_total - need;
    const int64_t win_hi = keep_lowest ? need : ties_total;
    const bool all_ties = need >= ties_total;   // no rank bookkeeping needed
    const int64_t chunk = chunk_size(count, gridDim.x);
    const int64_t lo = blockIdx.x * chunk, hi = min(lo + chunk, count);

    if (!all_ties && !empty) {  // ties in earlier blocks (and on lower ranks)
        long long s = 0;
        for (int b = threadIdx.x; b < (int)blockIdx.x; b += kThreads) s += st->block_ties[b];
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            long long tot = *ties_before_p;
            for (int w = 0; w < kThreads / 32; ++w) tot += warp_sums[w];
            block_base = tot;
        }
        __syncthreads();
    }
    long long running = (!all_ties && !empty) ? block_base : 0;
    for (int64_t base = lo; base < hi; base += kThreads) {
        const int64_t i = base + threadIdx.x;
        const bool in = i < hi;
        bool excluded = false, below = false, tie = false;
        if (in && !empty) {
            excluded = exclude && exclude[i];
            if (!excluded) {
                uint64_t k = select_key(scores[i], keep_lowest);
                below = k < t;
                tie = k == t;
            }
        } else if (in) {
            excluded = exclude && exclude[i];
        }
        bool keep = below;
        if (all_ties) {
            keep = keep || tie;
        } else if (!empty) {
            // block-wide exclusive rank of this tie in position order
            const unsigned bal = __ballot_sync(0xffffffffu, tie);
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            __syncthreads();  // warp_sums reuse
            if (lane == 0) warp_sums[warp] = __popc(bal);
            __syncthreads();
            long long before = running;
            for (int w = 0; w < warp; ++w) before += warp_sums[w];
            long long tile_total = 0;
            for (int w = 0; w < kThreads / 32; ++w) tile_total += warp_sums[w];
            const long long rank = before + __popc(bal & ((1u << lane) - 1u));
            if (tie && rank >= win_lo && rank < win_hi) keep = true;
            running += tile_total;
        }
        if (in) {
            if (!or_into) mask[i] = keep ? 1 : 0;
            else if (keep) mask[i] = 1;  // union with what is already marked (excluded entries are never kept here)
        }
    }
}

// ---- degree-aware guarantee (reference core.py:421-435) -------------------------------------------------
__global__ void node_best_key_kernel(const int64_t* __restrict__ src, const double* __restrict__ scores, int64_t count,
                                     const uint8_t* __restrict__ mask, unsigned long long* __restrict__ best) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        if (mask[i]) continue;
        atomicMax(&best[src[i]], (unsigned long long)ordered_key(scores[i]) );
    }
}

__global__ void node_best_pos_kernel(const int64_t* __restrict__ src, const double* __restrict__ scores, int64_t count,
                                     const uint8_t* __restrict__ mask, const unsigned long long* __restrict__ best,
                                     long long* __restrict__ best_pos) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        if (mask[i]) continue;
        const int64_t v = src[i];
        if ((unsigned long long)ordered_key(scores[i]) == best[v]) atomicMax(&best_pos[v], (long long)i);
    }
}

__global__ void node_mark_kernel(int64_t num_nodes, long long* __restrict__ best_pos, unsigned long long* __restrict__ best,
                                 uint8_t* __restrict__ mask, unsigned long long* __restrict__ marked) {
    unsigned long long c = 0;
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < num_nodes; v += (int64_t)gridDim.x * blockDim.x) {
        const long long p = best_pos[v];
        if (p >= 0) {
            mask[p] = 1;
            ++c;
        }
        best_pos[v] = -1;  // reset for the next round
        best[v] = 0;
    }
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(marked, c);
}

__global__ void fill_i64_kernel(int64_t n, long long* p, long long v) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

// ---- compaction + "-W" weights ---------------------------------------------------------------------------
struct CompactScratch {
    unsigned long long min_key, max_key;
    long long block_counts[kMaxBlocks];
};

__device__ __forceinline__ double key_to_double(uint64_t k) {
    uint64_t b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

__global__ void __launch_bounds__(kThreads)
compact_count_kernel(int64_t count, const uint8_t* __restrict__ mask, const double* __restrict__ scores,
                     CompactScratch* sc) {
    __shared__ unsigned long long sh_cnt, sh_min, sh_max;
    if (threadIdx.x == 0) { sh_cnt = 0; sh_min = ~0ull; sh_max = 0; }
    __syncthreads();
    const int64_t chunk = chunk_size(count, gridDim.x);
    const int64_t lo = blockIdx.x * chunk, hi = min(lo + chunk, count);
    unsigned long long c = 0, mn = ~0ull, mx = 0;
    for (int64_t i = lo + threadIdx.x; i < hi; i += kThreads) {
        if (mask[i]) {
            ++c;
            if (scores) {
                uint64_t k = ordered_key(scores[i]);
                mn = k < mn ? k : mn;
                mx = k > mx ? k : mx;
            }
        }
    }
    for (int o = 16; o; o >>= 1) {
        c += __shfl_xor_sync(0xffffffffu, c, o);
        unsigned long long a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
        mn = a < mn ? a : mn;
        mx = b > mx ? b : mx;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&sh_cnt, c);
        atomicMin(&sh_min, mn);
        atomicMax(&sh_max, mx);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        sc->block_counts[blockIdx.x] = (long long)sh_cnt;
        if (scores && sh_cnt) {
            atomicMin(&sc->min_key, sh_min);
            atomicMax(&sc->max_key, sh_max);
        }
    }
}

__global__ void __launch_bounds__(kThreads)
compact_scatter_kernel(const int64_t* __restrict__ ei, int64_t ld, int64_t count, const uint8_t* __restrict__ mask,
                       const double* __restrict__ scores, int invert, const CompactScratch* __restrict__ sc,
                       int64_t* __restrict__ out_ei, int64_t out_ld, float* __restrict__ out_w,
                       int64_t* __restrict__ num_kept) {
    __shared__ long long warp_sums[kThreads / 32];
    __shared__ long long block_base;
    {
        long long s = 0;
        for (int b = threadIdx.x; b < (int)blockIdx.x; b += kThreads) s += sc->block_counts[b];
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            long long tot = 0;
            for (int w = 0; w < kThreads / 32; ++w) tot += warp_sums[w];
            block_base = tot;
            if (blockIdx.x == gridDim.x - 1 && num_kept) *num_kept = tot + sc->block_counts[blockIdx.x];
        }
        __syncthreads();
    }
    double mn = 0.0, denom = 1.0;
    if (out_w && scores) {
        mn = key_to_double(sc->min_key);
        const double mx = key_to_double(sc->max_key);
        denom = __dadd_rn(__dsub_rn(mx, mn), 1e-8);  // (mx - mn + 1e-8), roman_empire_gpu.py:251
    }
    const int64_t chunk = chunk_size(count, gridDim.x);
    const int64_t lo = blockIdx.x * chunk, hi = min(lo + chunk, count);
    long long running = block_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t base = lo; base < hi; base += kThreads) {
        const int64_t i = base + threadIdx.x;
        const bool keep = i < hi && mask[i];
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        __syncthreads();
        if (lane == 0) warp_sums[warp] = __popc(bal);
        __syncthreads();
        long long before = running, tile_total = 0;
        for (int w = 0; w < kThreads / 32; ++w) {
            if (w < warp) before += warp_sums[w];
            tile_total += warp_sums[w];
        }
        if (keep) {
            const long long p = before + __popc(bal & ((1u << lane) - 1u));
            if (p < out_ld) {
                out_ei[p] = ei[i];
                out_ei[out_ld + p] = ei[ld + i];
                if (out_w && scores) {
                    double w = __ddiv_rn(__dsub_rn(scores[i], mn), denom);
                    if (invert) w = __dsub_rn(1.0, w);
                    out_w[p] = (float)w;  // torch.tensor(norm, dtype=float32): round to nearest
                }
            }
        }
        running += tile_total;
    }
}

// ---- fused tail of the single-GPU select: tie counts, mask, compaction, "-W" weights in two passes --------------------
// Replaces count_ties + write_mask + compact_count + compact_scatter (four kernels, two of which re-read the mask) when
// the caller wants the kept edge list anyway (`GraphSparsifier.sparsify`): tally = per-block numbers of keys below the
// boundary key and equal to it (8 B / score), emit = mask bytes + kept edge_index columns (+ weights) in one sweep
// (8 B score + 16 B edge + 1 B mask + 16 B per kept column). Positions inside a 1024-score tile come from ONE block
// scan of the packed pair (below, ties): the number of kept ties before a position is a function of the tie rank alone.
struct FusedScratch {
    long long block_below[kMaxBlocks];
    long long block_ties[kMaxBlocks];
};
constexpr int kTile = 4 * kThreads;   // scores per block iteration, four consecutive ones per thread

__device__ __forceinline__ void load4(const double* __restrict__ scores, int64_t i, int64_t hi, bool vec, double (&v)[4]) {
    if (vec && i + 3 < hi) {
        const double2 a = *reinterpret_cast<const double2*>(scores + i), b = *reinterpret_cast<const double2*>(scores + i + 2);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = i + j < hi ? scores[i + j] : 0.0;
    }
}

__global__ void __launch_bounds__(kThreads)
fused_tally_kernel(const double* __restrict__ scores, int64_t count, const SelectState* __restrict__ st, FusedScratch* fs,
                   unsigned long long* __restrict__ totals) {
    __shared__ unsigned long long sh[2];
    if (threadIdx.x < 2) sh[threadIdx.x] = 0;
    __syncthreads();
    const int keep_lowest = st->keep_lowest;
    const uint64_t t = st->prefix;
    const bool vec = (reinterpret_cast<uintptr_t>(scores) & 15) == 0;
    int64_t chunk = (count + gridDim.x - 1) / gridDim.x;
    chunk = (chunk + kTile - 1) / kTile * kTile;
    const int64_t lo = blockIdx.x * chunk, hi = min(lo + chunk, count);
    unsigned long long nb = 0, nt = 0;
    if (!st->empty) {
        for (int64_t i0 = lo + 4 * (int64_t)threadIdx.x; i0 < hi; i0 += 2 * kTile) {   // two tiles in flight
            double v[2][4];
            load4(scores, i0, hi, vec, v[0]);
            load4(scores, i0 + kTile, hi, vec, v[1]);
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int64_t i = i0 + u * kTile + j;
                    const uint64_t k = select_key(v[u][j], keep_lowest);
                    nb += i < hi && k < t;
                    nt += i < hi && k == t;
                }
        }
    }
    for (int o = 16; o; o >>= 1) {
        nb += __shfl_xor_sync(0xffffffffu, nb, o);
        nt += __shfl_xor_sync(0xffffffffu, nt, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&sh[0], nb);
        atomicAdd(&sh[1], nt);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        fs->block_below[blockIdx.x] = (long long)sh[0];
        fs->block_ties[blockIdx.x] = (long long)sh[1];
        if (totals) {   // sharded caller: this rank's (below, ties), all-gathered before the emit
            if (sh[0]) atomicAdd(&totals[0], sh[0]);
            if (sh[1]) atomicAdd(&totals[1], sh[1]);
        }
    }
}

__global__ void __launch_bounds__(kThreads, 5)
fused_emit_kernel(const double* __restrict__ scores, int64_t count, const SelectState* __restrict__ st,
                  const FusedScratch* __restrict__ fs, const int64_t* __restrict__ ei, int64_t ld, int invert,
                  uint8_t* __restrict__ mask, int64_t* __restrict__ out_ei, int64_t out_ld, float* __restrict__ out_w,
                  int64_t* __restrict__ num_kept, const long long* __restrict__ rank_totals, int rank, int nranks) {
    __shared__ long long red[3][kThreads / 32];
    __shared__ unsigned int warp_tot[2][kThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int keep_lowest = st->keep_lowest;
    const bool empty = st->empty;
    const uint64_t t = st->prefix;
    const long long need = st->remaining;
    // ties / keys below the boundary in earlier blocks, and all ties
    long long below_before = 0, ties_before = 0, ties_total = 0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += kThreads) {
        const long long tb = fs->block_ties[b];
        ties_total += tb;
        if (b < (int)blockIdx.x) {
            ties_before += tb;
            below_before += fs->block_below[b];
        }
    }
    for (int o = 16; o; o >>= 1) {
        below_before += __shfl_xor_sync(0xffffffffu, below_before, o);
        ties_before += __shfl_xor_sync(0xffffffffu, ties_before, o);
        ties_total += __shfl_xor_sync(0xffffffffu, ties_total, o);
    }
    if (lane == 0) { red[0][warp] = below_before; red[1][warp] = ties_before; red[2][warp] = ties_total; }
    __syncthreads();
    below_before = ties_before = ties_total = 0;
    for (int w = 0; w < kThreads / 32; ++w) { below_before += red[0][w]; ties_before += red[1][w]; ties_total += red[2][w]; }
    long long ties_lower_ranks = 0;   // sharded: tie ranks are global (rank order == position order)
    if (rank_totals) {
        ties_total = 0;
        for (int r = 0; r < nranks; ++r) {
            const long long tr = rank_totals[2 * r + 1];
            ties_total += tr;
            if (r < rank) ties_lower_ranks += tr;
        }
    }
    // window of tie ranks (position order) that are kept: the highest ones for top-k, the lowest ones for keep_lowest
    const long long win_lo = keep_lowest ? 0 : max(ties_total - need, 0ll);
    const long long win_hi = keep_lowest ? min(need, ties_total) : ties_total;
    auto kept_ties_before = [&](long long rank) { return max(min(rank, win_hi) - win_lo, 0ll); };
    double mn = 0.0, denom = 1.0;
    if (out_w) {   // extrema of the kept scores: the boundary key and the best key (no pass over the kept set needed)
        const double s_t = key_to_double(keep_lowest ? t : ~t), s_best = key_to_double(keep_lowest ? st->best_key : ~st->best_key);
        mn = keep_lowest ? s_best : s_t;
        const double mx = keep_lowest ? s_t : s_best;
        denom = __dadd_rn(__dsub_rn(mx, mn), 1e-8);   // (mx - mn + 1e-8), roman_empire_gpu.py:251
    }
    const bool vec = (reinterpret_cast<uintptr_t>(scores) & 15) == 0;
    int64_t chunk = (count + gridDim.x - 1) / gridDim.x;
    chunk = (chunk + kTile - 1) / kTile * kTile;
    const int64_t lo = blockIdx.x * chunk, hi = min(lo + chunk, count);
    long long tie_base = ties_lower_ranks + ties_before;
    long long kept_base = below_before + (kept_ties_before(tie_base) - kept_ties_before(ties_lower_ranks));
    int buf = 0;
    for (int64_t base = lo; base < hi; base += kTile, buf ^= 1) {
        const int64_t i0 = base + 4 * (int64_t)threadIdx.x;
        double v[4];
        load4(scores, i0, hi, vec, v);
        bool below[4], tie[4];
        unsigned int nb = 0, nt = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint64_t k = select_key(v[j], keep_lowest);
            const bool in = i0 + j < hi && !empty;
            below[j] = in && k < t;
            tie[j] = in && k == t;
            nb += below[j];
            nt += tie[j];
        }
        // block-exclusive scan of (below << 16 | ties): a tile holds 1024 scores, neither half overflows
        unsigned int packed = (nb << 16) | nt, incl = packed;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        if (lane == 31) warp_tot[buf][warp] = incl;
        __syncthreads();
        unsigned int before = incl - packed, tile_tot = 0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) {
            const unsigned int wt = warp_tot[buf][w];
            if (w < warp) before += wt;
            tile_tot += wt;
        }
        long long rank = tie_base + (before & 0xffffu);                       // tie rank of this thread's first tie
        long long pos = kept_base + (before >> 16) + (kept_ties_before(rank) - kept_ties_before(tie_base));
        unsigned char m[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            bool keep = below[j];
            if (tie[j]) {
                keep = rank >= win_lo && rank < win_hi;
                ++rank;
            }
            m[j] = keep ? 1 : 0;
            if (keep) {
                if (pos < out_ld) {
                    out_ei[pos] = ei[i0 + j];
                    out_ei[out_ld + pos] = ei[ld + i0 + j];
                    if (out_w) {
                        double w = __ddiv_rn(__dsub_rn(v[j], mn), denom);
                        if (invert) w = __dsub_rn(1.0, w);
                        out_w[pos] = (float)w;  // torch.tensor(norm, dtype=float32): round to nearest
                    }
                }
                ++pos;
            }
        }
        if (mask) {
            if (i0 + 3 < hi && (reinterpret_cast<uintptr_t>(mask + i0) & 3) == 0) {
                *reinterpret_cast<uchar4*>(mask + i0) = make_uchar4(m[0], m[1], m[2], m[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (i0 + j < hi) mask[i0 + j] = m[j];
            }
        }
        const long long tile_ties = tile_tot & 0xffffu;
        kept_base += (tile_tot >> 16) + (kept_ties_before(tie_base + tile_ties) - kept_ties_before(tie_base));
        tie_base += tile_ties;
    }
    if (num_kept && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *num_kept = kept_base;
}

int blocks_for(int64_t count) {
    int64_t b = (count + 4 * kThreads - 1) / (4 * kThreads);
    if (b < 1) b = 1;
    if (b > kMaxBlocks) b = kMaxBlocks;
    return (int)b;
}

}  // namespace
}  // namespace gsp

using namespace gsp;

static_assert(sizeof(SelectState) <= GSP_SELECT_STATE_BYTES, "GSP_SELECT_STATE_BYTES too small");

GSP_API int gsp_select_begin(void* d_state, int64_t num_keep, int keep_lowest, void* stream) {
    GSP_REQUIRE(d_state != nullptr, "d_state is NULL");
    GSP_REQUIRE(num_keep >= 0, "num_keep must be >= 0");
    begin_kernel<<<1, 1, 0, as_stream(stream)>>>(reinterpret_cast<SelectState*>(d_state), num_keep, keep_lowest ? 1 : 0, 0);
    GSP_CHECK_LAUNCH();
    return GSP_OK;
}

namespace gsp {
namespace {
// begin + the six histogram / pick rounds on ONE device: live-key extrema on, so a boundary bucket that is a single tie
// class ends the rounds early
int select_boundary_local(const double* d_scores, int64_t count, const uint8_t* d_exclude, int64_t num_keep, int keep_lowest,
                          SelectState* state, uint64_t* hist, void* stream) {
    begin_kernel<<<1, 1, 0, as_stream(stream)>>>(state, num_keep, keep_lowest ? 1 : 0, 1);
    GSP_CHECK_LAUNCH();
    for (int p = 0; p < kPasses; ++p) {
        if (int rc = gsp_select_histogram(d_scores, count, d_exclude, state, p, hist, stream)) return rc;
        if (int rc = gsp_select_pick(state, hist, p, stream)) return rc;
    }
    return GSP_OK;
}
}  // namespace
}  // namespace gsp

GSP_API int gsp_select_histogram(const double* d_scores, int64_t count, const uint8_t* d_exclude, const void* d_state,
                                 int pass, uint64_t* d_hist, void* stream) {
    GSP_REQUIRE(d_state && d_hist, "NULL argument");
    GSP_REQUIRE(count >= 0 && (count == 0 || d_scores), "bad scores");
    GSP_REQUIRE(pass >= 0 && pass < kPasses, "pass out of range");
    cudaStream_t s = as_stream(stream);
    GSP_CUDA_TRY(cudaMemsetAsync(d_hist, 0, kBins * sizeof(uint64_t), s));
    if (count == 0) return GSP_OK;
    SelectState* st = reinterpret_cast<SelectState*>(const_cast<void*>(d_state));
    if ((reinterpret_cast<uintptr_t>(d_scores) & 15) == 0)
        histogram_kernel<true><<<grid_for(count, 8 * kThreads, 8), kThreads, 0, s>>>(d_scores, count, d_exclude, st, pass,
                                                                                   reinterpret_cast<unsigned long long*>(d_hist), nullptr);
    else
        histogram_kernel<false><<<grid_for(count, 8 * kThreads, 8), kThreads, 0, s>>>(d_scores, count, d_exclude, st, pass,
                                                                                    reinterpret_cast<unsigned long long*>(d_hist), nullptr);
    GSP_CHECK_LAUNCH();
    return GSP_OK;
}

GSP_API int gsp_select_pick(void* d_state, const uint64_t* d_hist, int pass, void* stream) {
    GSP_REQUIRE(d_state && d_hist, "NULL argument");
    GSP_REQUIRE(pass >= 0 && pass < kPasses, "pass out of range");
    pick_kernel<<<1, 32, 0, as_stream(stream)>>>(reinterpret_cast<SelectState*>(d_state),
                                                 reinterpret_cast<const unsigned long long*>(d_hist), pass, 0);
    GSP_CHECK_LAUNCH();
    return GSP_OK;
}

GSP_API int gsp_select_count_ties(const double* d_scores, int64_t count, const uint8_t* d_exclude, const void* d_state,
                                  int64_t* d_tie_count, void* stream) {
    GSP_REQUIRE(d_state && d_tie_count, "NULL argument");
    cudaStream_t s = as_stream(stream);
    GSP_CUDA_TRY(cudaMemsetAsync(d_tie_count, 0, sizeof(int64_t), s));
    if (count == 0) return GSP_OK;
    count_ties_kernel<<<blocks_for(count), kThreads, 0, s>>>(
        d_scores, count, d_exclude, reinterpret_cast<SelectState*>(const_cast<void*>(d_state)),
        reinterpret_cast<unsigned long long*>(d_tie_count));
    GSP_CHECK_LAUNCH();
    return GSP_OK;
}

GSP_API int gsp_select_write_mask(const double* d_scores, int64_t count, const uint8_t* d_exclude, const void* d_state,
                                  const int64_t* d_ties_before, const int64_t* d_ties_total, int or_into, uint8_t* d_mask,
                                  void* stream) {
    GSP_REQUIRE(d_state && d_ties_before && d_ties_total, "NULL argument");
    if (count == 0) return GSP_OK;
    GSP_REQUIRE(d_mask != nullptr, "d_mask is NULL");
    write_mask_kernel<<<blocks_for(count), kThreads, 0, as_stream(stream)>>>(
        d_scores, count, d_exclude, reinterpret_cast<const SelectState*>(d_state), d_ties_before, d_ties_total, or_into,
        d_mask);
    GSP_CHECK_LAUNCH();
    return GSP_OK;
}

GSP_API int gsp_select_mask(const double* d_scores, int64_t count, int64_t num_keep, int keep_lowest,
                            const uint8_t* d_exclude, int or_into, uint8_t* d_mask, void* stream) {
    GSP_REQUIRE(count >= 0 && num_keep >= 0, "negative size");
    if (count == 0) return GSP_OK;
    cudaStream_t s = as_stream(stream);
    Scratch<char> state;
    Scratch<uint64_t> hist;
    Scratch<int64_t> ties;  // [0] = total ties, [1] = 0 (ties before this shard)
    GSP_CUDA_TRY(state.alloc(GSP_SELECT_STATE_BYTES, s));
    GSP_CUDA_TRY(hist.alloc(kBins, s));
    GSP_CUDA_TRY(ties.alloc(2, s));
    GSP_CUDA_TRY(cudaMemsetAsync(ties.ptr, 0, 2 * sizeof(int64_t), s));
    if (int rc = select_boundary_local(d_scores, count, d_exclude, num_keep, keep_lowest, reinterpret_cast<SelectState*>(state.ptr),
                                       hist.ptr, stream)) return rc;
    if (int rc = gsp_select_count_ties(d_scores, count, d_exclude, state.ptr, ties.ptr, stream)) return rc;
    return gsp_select_write_mask(d_scores, count, d_exclude, state.ptr, ties.ptr + 1, ties.ptr, or_into, d_mask, stream);
}

GSP_API int gsp_select_compact(const double* d_scores, int64_t count, int64_t num_keep, int keep_lowest,
                               const int64_t* d_edge_index, int64_t ld, uint8_t* d_mask, int64_t* d_out_edge_index,
                               int64_t out_ld, float* d_out_weight, int invert_weights, int64_t* d_num_kept, void* stream) {
    GSP_REQUIRE(count >= 0 && num_keep >= 0 && ld >= count && out_ld >= 0, "bad sizes");
    cudaStream_t s = as_stream(stream);
    if (count == 0) {
        if (d_num_kept) GSP_CUDA_TRY(cudaMemsetAsync(d_num_kept, 0, sizeof(int64_t), s));
        return GSP_OK;
    }
    GSP_REQUIRE(d_scores && d_edge_index, "NULL argument");
    GSP_REQUIRE(out_ld == 0 || d_out_edge_index, "d_out_edge_index is NULL");
    Scratch<char> state;
    Scratch<uint64_t> hist;
    Scratch<FusedScratch> fs;
    GSP_CUDA_TRY(state.alloc(GSP_SELECT_STATE_BYTES, s));
    GSP_CUDA_TRY(hist.alloc(kBins, s));
    GSP_CUDA_TRY(fs.alloc(1, s));
    SelectState* st = reinterpret_cast<SelectState*>(state.ptr);
    if (int rc = select_boundary_local(d_scores, count, nullptr, num_keep, keep_lowest, st, hist.ptr, stream)) return rc;
    const int blocks = blocks_for(count);
    fused_tally_kernel<<<blocks, kThreads, 0, s>>>(d_scores, count, st, fs.ptr, nullptr);
    GSP_CHECK_LAUNCH();
    fused_emit_kernel<<<blocks, kThreads, 0, s>>>(d_scores, count, st, fs.ptr, d_edge_index, ld, invert_weights, d_mask,
                                                  d_out_edge_index, out_ld, d_out_weight, d_num_kept, nullptr, 0, 1);
    GSP_CHECK_LAUNCH();
    return GSP_OK;
}

// ---- sharded select + compaction (one rank's slice; the caller all-gathers the slots and the totals) ------------------
static_assert(kSlotWords == GSP_SELECT_SLOT_WORDS, "slot layout");
static_assert(sizeof(FusedScratch) <= GSP_SELECT_SCRATCH_BYTES, "GSP_SELECT_SCRATCH_BYTES too small");

GSP_API int gsp_select_histogram_slot(const double* d_scores, int64_t count, const void* d_state, int pass, uint64_t* d_slot,
                                      void* stream) {
    GSP_REQUIRE(d_state && d_slot, "NULL argument");
    GSP_REQUIRE(count >= 0 && (count == 0 || d_scores), "bad scores");
    GSP_REQUIRE(pass >= 0 && pass < kPasses, "pass out of range");
    cudaStream_t s = as_stream(stream);
    GSP_CUDA_TRY(cudaMemsetAsync(d_slot, 0, kSlotWords * sizeof(uint64_t), s));
    if (count == 0) return GSP_OK;
    SelectState* st = reinterpret_cast<SelectState*>(const_cast<void*>(d_state));
    unsigned long long* slot = reinterpret_cast<unsigned long long*>(d_slot);
    if ((reinterpret_cast<uintptr_t>(d_scores) & 15) == 0)
        histogram_kernel<true><<<grid_for(count, 8 * kThreads, 8), kThreads, 0, s>>>(d_scores, count, nullptr, st, pass, slot,
                                                                                   slot + kBins);
    else
        histogram_kernel<false><<<grid_for(count, 8 * kThreads, 8), kThreads, 0, s>>>(d_scores, count, nullptr, st, pass, slot,
                                                                                    slot + kBins);
    GSP_CHECK_LAUNCH();
    return GSP_OK;
}

GSP_API int gsp_select_pick_slots(void* d_state, const uint64_t* d_slots, int32_t nranks, int pass, void* stream) {
    GSP_REQUIRE(d_state && d_slots, "NULL argument");
    GSP_REQUIRE(nranks >= 1, "nranks must be >= 1");
    GSP_REQUIRE(pass >= 0 && pass < kPasses, "pass out of range");
    pick_kernel<<<1, 32, 0, as_stream(stream)>>>(reinterpret_cast<SelectState*>(d_state),
                                                 reinterpret_cast<const unsigned long long*>(d_slots), pass, nranks);
    GSP_CHECK_LAUNCH();
    return GSP_OK;
}

GSP_API int gsp_select_tally(const double* d_scores, int64_t count, const void* d_state, void* d_scratch, int64_t* d_totals,
                             void* stream) {
    GSP_REQUIRE(d_state && d_scratch && d_totals, "NULL argument");
    GSP_REQUIRE(count >= 0 && (count == 0 || d_scores), "bad scores");
    cudaStream_t s = as_stream(stream);
    GSP_CUDA_TRY(cudaMemsetAsync(d_totals, 0, 2 * sizeof(int64_t), s));
    if (count == 0) return GSP_OK;
    fused_tally_kernel<<<blocks_for(count), kThreads, 0, s>>>(d_scores, count, reinterpret_cast<const SelectState*>(d_state),
                                                              reinterpret_cast<FusedScratch*>(d_scratch),
                                                              reinterpret_cast<unsigned long long*>(d_totals));
    GSP_CHECK_LAUNCH();
    return GSP_OK;
}

GSP_API int gsp_select_emit(const double* d_scores, int64_t count, const void* d_state, const void* d_scratch,
                            const int64_t* d_rank_totals, int32_t rank, int32_t nranks, const int64_t* d_edge_index, int64_t ld,
                            uint8_t* d_mask, int64_t* d_out_edge_index, int64_t out_ld, float* d_out_weight, int invert_weights,
                            int64_t* d_num_kept, void* stream) {
    GSP_REQUIRE(count >= 0 && ld >= count && out_ld >= 0, "bad sizes");
    GSP_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank");
    cudaStream_t s = as_stream(stream);
    if (count == 0) {
        if (d_num_kept) GSP_CUDA_TRY(cudaMemsetAsync(d_num_kept, 0, sizeof(int64_t), s));
        return GSP_OK;
    }
    GSP_REQUIRE(d_scores && d_state && d_scratch && d_rank_totals && d_edge_index, "NULL argument");
    GSP_REQUIRE(out_ld == 0 || d_out_edge_index, "d_out_edge_index is NULL");
    fused_emit_kernel<<<blocks_for(count), kThreads, 0, s>>>(
        d_scores, count, reinterpret_cast<const SelectState*>(d_state), reinterpret_cast<const FusedScratch*>(d_scratch),
        d_edge_index, ld, invert_weights, d_mask, d_out_edge_index, out_ld, d_out_weight, d_num_kept,
        reinterpret_cast<const long long*>(d_rank_totals), rank, nranks);
    GSP_CHECK_LAUNCH();
    return GSP_OK;
}

GSP_API int gsp_degree_aware_guarantee(const int64_t* d_src, const double* d_scores, int64_t count, int64_t num_nodes,
                                       int32_t min_per_node, uint8_t* d_mask, int64_t* d_num_marked, void* stream) {
    GSP_REQUIRE(count >= 0 && num_nodes >= 0 && min_per_node >= 0, "negative size");
    GSP_REQUIRE(d_num_marked != nullptr, "d_num_marked is NULL");
    cudaStream_t s = as_stream(stream);
    GSP_CUDA_TRY(cudaMemsetAsync(d_num_marked, 0, sizeof(int64_t), s));
    if (count == 0) return GSP_OK;
    GSP_REQUIRE(d_src && d_scores && d_mask, "NULL argument");
    GSP_CUDA_TRY(cudaMemsetAsync(d_mask, 0, (size_t)count, s));
    if (num_nodes == 0 || min_per_node == 0) return GSP_OK;
    Scratch<unsigned long long> best;
    Scratch<long long> best_pos;
    GSP_CUDA_TRY(best.alloc(num_nodes, s));
    GSP_CUDA_TRY(best_pos.alloc(num_nodes, s));
    GSP_CUDA_TRY(cudaMemsetAsync(best.ptr, 0, (size_t)num_nodes * sizeof(unsigned long long), s));
    fill_i64_kernel<<<grid_for(num_nodes, 256), 256, 0, s>>>(num_nodes, best_pos.ptr, -1);
    GSP_CHECK_LAUNCH();
    const int grid = grid_for(count, 256);
    for (int round = 0; round < min_per_node; ++round) {
        node_best_key_kernel<<<grid, 256, 0, s>>>(d_src, d_scores, count, d_mask, best.ptr);
        GSP_CHECK_LAUNCH();
        node_best_pos_kernel<<<grid, 256, 0, s>>>(d_src, d_scores, count, d_mask, best.ptr, best_pos.ptr);
        GSP_CHECK_LAUNCH();
        node_mark_kernel<<<grid_for(num_nodes, 256), 256, 0, s>>>(num_nodes, best_pos.ptr, best.ptr, d_mask,
                                                                  reinterpret_cast<unsigned long long*>(d_num_marked));
        GSP_CHECK_LAUNCH();
    }
    return GSP_OK;
}

GSP_API int gsp_compact_edges(const int64_t* d_edge_index, int64_t ld, int64_t count, const uint8_t* d_mask,
                              const double* d_scores, int invert_weights, int64_t* d_out_edge_index, int64_t out_ld,
                              float* d_out_weight, int64_t* d_num_kept, void* stream) {
    GSP_REQUIRE(count >= 0 && ld >= count && out_ld >= 0, "bad sizes");
    cudaStream_t s = as_stream(stream);
    if (count == 0) {
        if (d_num_kept) GSP_CUDA_TRY(cudaMemsetAsync(d_num_kept, 0, sizeof(int64_t), s));
        return GSP_OK;
    }
    GSP_REQUIRE(d_edge_index && d_mask, "NULL argument");
    GSP_REQUIRE(out_ld == 0 || d_out_edge_index, "d_out_edge_index is NULL");
    Scratch<CompactScratch> sc;
    GSP_CUDA_TRY(sc.alloc(1, s));
    GSP_CUDA_TRY(cudaMemsetAsync(sc.ptr, 0, sizeof(CompactScratch), s));
    GSP_CUDA_TRY(cudaMemsetAsync(&sc.ptr->min_key, 0xff, sizeof(unsigned long long), s));
    const int blocks = blocks_for(count);
    const double* sc_scores = d_out_weight ? d_scores : nullptr;
    compact_count_kernel<<<blocks, kThreads, 0, s>>>(count, d_mask, sc_scores, sc.ptr);
    GSP_CHECK_LAUNCH();
    compact_scatter_kernel<<<blocks, kThreads, 0, s>>>(d_edge_index, ld, count, d_mask, sc_scores, invert_weights, sc.ptr,
                                                       d_out_edge_index, out_ld, d_out_weight, d_num_kept);
    GSP_CHECK_LAUNCH();
    return GSP_OK;
}
