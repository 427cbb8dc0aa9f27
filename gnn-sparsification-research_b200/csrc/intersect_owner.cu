// intersect_owner.cu — owner-hashed neighbour intersection for symmetric graphs (fast path of K2 / K3).
//
// Same results as intersect.cu (reference src/sparsification/metrics.py:43-64 and :99-121), different
// schedule. On a power-law graph almost every edge touches a high-degree endpoint, so per-edge searches
// of the short list in the long one cost sum_e min(d)*log(max(d)) dependent probes. Here every undirected
// pair {o, w} is evaluated ONCE, at its OWNER o = the endpoint with the larger degree (ties: smaller id):
// the owner's neighbour set is put into a shared-memory hash table once and the (shorter) list of every
// owned neighbour w is streamed through it with coalesced loads — sum over pairs of min(d) element tests,
// one smem probe each. The score is written to both directed positions (o,w) and (w,o) — the mirrored position
// comes from the reverse-offset array the graph build fills during its symmetry pass (intersection, degrees and
// the descending-id accumulation order are all symmetric in the pair).
//
//   level A  rows with 1 <= deg <= 64: one warp per owner, 128-slot table per warp, neighbours in registers
//   level B  rows with deg > 64: CTA per (owner, neighbour chunk). Two size classes of the same kernel:
//            64 < deg <= 1536: 256 threads, 4096 cuckoo slots, 512-neighbour chunks (~6 CTAs per SM);
//            deg > 1536: 768 threads, two CTAs per SM, the owner row hashed in tiles of 6144 ids (16384 slots),
//            tiles visited in DESCENDING id order so Adamic-Adar keeps SciPy's accumulation order. Neighbour
//            metadata is fetched once per item; each neighbour keeps a cursor into its row between tiles.
//
// An edge range [e_begin, e_end) restricts the pairs to those with a directed position inside the range (only in-range
// positions are written); an owner range [owner_lo, owner_hi) restricts them to the pairs those nodes own (multi-GPU:
// every pair evaluated on exactly one rank, full-length outputs reduce-scattered by the caller).
// The hub kernel sits exactly at 40 registers x 1536 threads per SM; experiments that added live state (cross-row
// prefetch) fell to one CTA per SM and ran 2x slower.
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace gsp {
namespace {

constexpr int kWarpOwnerMax = 64;          // level A / level B split
constexpr int kATableSlots = 128;          // per-warp hash slots (load factor <= 0.5)
constexpr int kAThreads = 256;
constexpr int kAWarps = kAThreads / kWarp;
constexpr int kARowsPerClaim = 16;
constexpr int kADealtRowsPerClaim = 32;   // dealt ownership: a claim is screened with one row per lane

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

// ---- level-B membership structure: two-table cuckoo hash in shared memory --------------------------------
// A lookup is exactly two independent shared-memory loads (no probe loop, no divergence): profiling the first
// linear-probing version showed ~50 % of all issued instructions in its probe loop at 3-8 active lanes.
// Keys that cannot be placed after kCuckooMaxKicks evictions go to a small stash that lookups scan only when it
// is non-empty; if even the stash overflows the tile is rebuilt with the next pair of multipliers.
constexpr int kJaccardDepth = 8;   // Jaccard streams long rows eight groups deep (counts need no ordering)
constexpr int kQueueStride = 68;   // doubles per warp in the Adamic-Adar hit queue (two 32-id groups + padding, 16-byte aligned)
constexpr int kCuckooMaxKicks = 64;
constexpr int kStashMax = 32;

struct Cuckoo {
    uint32_t t1, t2;        // shared-state-space byte addresses of the two tables (explicit ld.shared, no generic loads)
    uint32_t mul1, mul2;
    int shift;              // 32 - log2(slots per table)
    int stash_n;
    const int32_t* stash;
};

__device__ __forceinline__ int32_t lds_s32(uint32_t addr) {
    int32_t v;
    asm("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

__device__ __forceinline__ void cuckoo_insert(int32_t* t1, int32_t* t2, int shift, uint32_t mul1, uint32_t mul2, int32_t x,
                                              int32_t* stash, int* stash_n) {
    int which = 0;
    for (int it = 0; it < kCuckooMaxKicks; ++it) {
        int32_t* tab = which ? t2 : t1;
        const uint32_t h = ((uint32_t)x * (which ? mul2 : mul1)) >> shift;
        x = atomicExch(&tab[h], x);
        if (x == -1) return;
        which ^= 1;  // the evicted key moves to its slot in the other table
    }
    const int k = atomicAdd(stash_n, 1);
    if (k < kStashMax) stash[k] = x;
}

// x may be the INT_MIN sentinel of an out-of-row lane: it never equals a stored id or the empty marker (-1).
__device__ __forceinline__ bool cuckoo_contains(const Cuckoo& c, int32_t x) {
    const uint32_t ux = (uint32_t)x;
    const int32_t a = lds_s32(c.t1 + (((ux * c.mul1) >> (c.shift - 2)) & ~3u));
    const int32_t b = lds_s32(c.t2 + (((ux * c.mul2) >> (c.shift - 2)) & ~3u));
    bool f = (a == x) | (b == x);
    if (c.stash_n) {
        for (int k = 0; k < c.stash_n; ++k) f |= c.stash[k] == x;
    }
    return f;
}

// Stash-free lookups for the main streaming loops (the caller checked stash_n == 0): two loads, two compares.
__device__ __forceinline__ bool cuckoo_hit(const Cuckoo& c, int32_t x) {
    const uint32_t ux = (uint32_t)x;
    const int32_t a = lds_s32(c.t1 + (((ux * c.mul1) >> (c.shift - 2)) & ~3u));
    const int32_t b = lds_s32(c.t2 + (((ux * c.mul2) >> (c.shift - 2)) & ~3u));
    return (a == x) | (b == x);
}

// count += contains(x) as compare, compare-or, predicated add (the generic bool -> int path costs five instructions)
__device__ __forceinline__ void cuckoo_count(const Cuckoo& c, int32_t x, int& count) {
    const uint32_t ux = (uint32_t)x;
    const int32_t a = lds_s32(c.t1 + (((ux * c.mul1) >> (c.shift - 2)) & ~3u));
    const int32_t b = lds_s32(c.t2 + (((ux * c.mul2) >> (c.shift - 2)) & ~3u));
    asm("{\n\t.reg .pred p, q;\n\tsetp.eq.s32 q, %1, %3;\n\tsetp.eq.or.s32 p, %2, %3, q;\n\t@p add.s32 %0, %0, 1;\n\t}"
        : "+r"(count) : "r"(a), "r"(b), "r"(x));
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
    int32_t owner_lo, owner_hi;  // only pairs owned by nodes in [owner_lo, owner_hi) are evaluated (owner sharding)
    const uint8_t* deal;         // ... and, when set, only by owners dealt to this rank (deal[o] == deal_rank)
    int32_t deal_rank;
    // peer scatter (multi-GPU): position p goes to slices[p / slice_len][p % slice_len]; the pointers may be peer
    // memory mapped over NVLink, so every score crosses the fabric exactly once, straight from the scoring kernel
    double* const* slices;
    int64_t slice_len;
    double* const* slices2;   // Jaccard slices of the fused Jaccard + Adamic-Adar pass (kMode 2)
    double inv_slice_len;     // 1 / slice_len: the slice of a position without a 64-bit division
    const struct ScatterInfo* scatter;   // the same four values in device memory, for the out-of-line store routine
};
struct ScatterInfo {
    double* const* slices;
    double* const* slices2;
    int64_t slice_len;
    double inv_slice_len;
};

// p / slice_len for 0 <= p < 2^53: the fp64 estimate is off by at most one
__device__ __forceinline__ int64_t slice_of(int64_t slice_len, double inv_slice_len, int64_t p) {
    int64_t k = __double2ll_rd((double)p * inv_slice_len);
    if (k * slice_len > p) --k;
    else if ((k + 1) * slice_len <= p) ++k;
    return k;
}

// The peer stores of one pair (position p lives at slices[p / slice_len][p % slice_len]).
__device__ __forceinline__ void scatter_pair(const ScatterInfo* __restrict__ info, int64_t p1, int64_t p2, double score, double jac) {
    double* const* slices = info->slices;
    double* const* slices2 = info->slices2;
    const int64_t slice_len = info->slice_len;
    const double inv = info->inv_slice_len;
    const int64_t k1 = slice_of(slice_len, inv, p1);
    slices[k1][p1 - k1 * slice_len] = score;
    if (slices2) slices2[k1][p1 - k1 * slice_len] = jac;
    if (p2 != p1) {
        const int64_t k2 = slice_of(slice_len, inv, p2);
        slices[k2][p2 - k2 * slice_len] = score;
        if (slices2) slices2[k2][p2 - k2 * slice_len] = jac;
    }
}

// The whole per-row epilogue of a work item in peer-scatter mode, out of line (its own register allocation): the state
// arrays follow each other in shared memory (base | acc | len | cursor | cnt | top, `chunk` entries each).
__device__ __noinline__ void scatter_chunk(const ScatterInfo* __restrict__ info, const int32_t* __restrict__ rev_off,
                                           const long long* base_s, int chunk, int nb, int64_t p_first, int d_o, int mode) {
    const double* acc_s = reinterpret_cast<const double*>(base_s + chunk);
    const int32_t* len_s = reinterpret_cast<const int32_t*>(acc_s + chunk);
    const int32_t* cnt_s = len_s + 2 * chunk;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) {
        const int d_w = len_s[i];
        if (d_w < 0) continue;  // pair owned by the neighbour
        const int64_t p1 = p_first + i;
        const int64_t p2 = base_s[i] + __ldg(rev_off + p1);
        const int count = cnt_s[i];
        double jac = 0.0;
        if (mode != 1) {
            const double uni = (double)d_o + (double)d_w - (double)count;
            jac = uni > 0.0 ? __ddiv_rn((double)count, uni) : 0.0;
        }
        scatter_pair(info, p1, p2, mode == 0 ? jac : acc_s[i], jac);
    }
}

__global__ void fill_scatter_info_kernel(ScatterInfo* info, double* const* slices, double* const* slices2, int64_t slice_len) {
    info->slices = slices;
    info->slices2 = slices2;
    info->slice_len = slice_len;
    info->inv_slice_len = slice_len > 0 ? 1.0 / (double)slice_len : 0.0;
}

// Stream row(w)[s, e) through the hash table. kMode 0: returns the number of hits. kMode 1: continues the
// ordered fp64 accumulation (ids visited in descending order). kMode 2: both (the fused pass: the hit ballots the
// ordered sum needs anyway also give the count).
template <int kMode>
__device__ __forceinline__ void stream_row(const int32_t* __restrict__ row_w, int s, int e, const int32_t* slots,
                                           uint32_t mask, int shift, int32_t o, const double* __restrict__ node_w,
                                           int& count, double& acc) {
    const int lane = lane_id();
    if (kMode == 0) {
        int c = 0;
        for (int base = s; base < e; base += kWarp) {
            const int i = base + lane;
            if (i < e) c += hash_contains(slots, mask, shift, __ldg(row_w + i));
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
                if (hit) term = __ldg(node_w + x);   // node_w holds the squared weights
            }
            unsigned hits = __ballot_sync(0xffffffffu, hit);
            if (kMode == 2) count += __popc(hits);
            while (hits) {
                const int src = __ffs(hits) - 1;
                acc = __dadd_rn(acc, __shfl_sync(0xffffffffu, term, src));
                hits &= hits - 1;
            }
        }
    }
}

// score_out receives the Jaccard score (kMode 0) or the Adamic-Adar sum (kMode 1, 2); jaccard_out the Jaccard score of
// the fused pass (kMode 2).
template <int kMode, bool kScatter>
__device__ __forceinline__ void write_pair(const RangeInfo& r, int64_t p1, int64_t p2, int d_o, int d_w, int count,
                                           double acc, int32_t* __restrict__ inter_out, double* __restrict__ score_out,
                                           double* __restrict__ jaccard_out) {
    double jac = 0.0;
    if (kMode != 1) {
        const double uni = (double)d_o + (double)d_w - (double)count;
        jac = uni > 0.0 ? __ddiv_rn((double)count, uni) : 0.0;
    }
    const double score = kMode == 0 ? jac : acc;
    if (kScatter) {
        scatter_pair(r.scatter, p1, p2, score, jac);
        return;
    }
    if (p1 >= r.e_begin && p1 < r.e_end) {
        score_out[p1 - r.e_begin] = score;
        if (kMode == 2) jaccard_out[p1 - r.e_begin] = jac;
        if (kMode != 1 && inter_out) inter_out[p1 - r.e_begin] = count;
    }
    if (p2 != p1 && p2 >= r.e_begin && p2 < r.e_end) {
        score_out[p2 - r.e_begin] = score;
        if (kMode == 2) jaccard_out[p2 - r.e_begin] = jac;
        if (kMode != 1 && inter_out) inter_out[p2 - r.e_begin] = count;
    }
}

// ---- level A: one warp per low-degree owner --------------------------------------------------------------
// kDealt (dealt ownership: most rows of a claim belong to other ranks): a claim of 32 rows is screened with one row per lane,
// so a foreign row costs a share of one ballot instead of a dependent indptr round trip of the whole warp (one of eight
// ranks: 1.13 -> 0.73 ms). Without a deal the rows are visited one after the other: screening every claim was a net loss
// there (4.26 -> 4.56 ms), and so was sharing one loop between the two modes (5.0 ms).
template <int kMode, bool kScatter, bool kDealt>
__global__ void __launch_bounds__(kAThreads)
warp_owner_kernel(int64_t n, const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                  const int32_t* __restrict__ rev_off, RangeInfo r,
                  const double* __restrict__ node_w, int32_t* __restrict__ inter_out, double* __restrict__ score_out,
                  double* __restrict__ jaccard_out, unsigned long long* counter) {
    __shared__ int32_t tables[kAWarps][kATableSlots];
    const int lane = lane_id();
    int32_t* slots = tables[threadIdx.x >> 5];
    constexpr uint32_t mask = kATableSlots - 1;
    constexpr int shift = 32 - 7;
    // one owner row: hash its (<= 64) neighbours into the warp's table, stream the rows of the neighbours it owns
    auto score_owner = [&](int32_t o, int64_t a0, int d_o) {
        const int64_t a1 = a0 + d_o;
        const bool row_in_range = r.full || (a0 < r.e_end && a1 > r.e_begin);
        // neighbours in registers: lane holds elements lane and lane + 32
        const int32_t nb0 = lane < d_o ? __ldg(indices + a0 + lane) : INT_MAX;
        const int32_t nb1 = lane + 32 < d_o ? __ldg(indices + a0 + lane + 32) : INT_MAX;
        if (!row_in_range) {  // only neighbours whose own row meets the range matter: any of them?
            const bool any0 = lane < d_o && nb0 >= r.row_lo && nb0 <= r.row_hi;
            const bool any1 = lane + 32 < d_o && nb1 >= r.row_lo && nb1 <= r.row_hi;
            if (!__any_sync(0xffffffffu, any0 || any1)) return;
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
            int count = 0;
            double acc = 0.0;
            stream_row<kMode>(indices + b0, 0, d_w, slots, mask, shift, o, node_w, count, acc);
            if (lane == 0)
                write_pair<kMode, kScatter>(r, p1, b0 + __ldg(rev_off + p1), d_o, d_w, count, acc, inter_out, score_out, jaccard_out);
        }
    };
    constexpr int kRows = kDealt ? kADealtRowsPerClaim : kARowsPerClaim;
    for (;;) {
        unsigned long long first = 0;
        if (lane == 0) first = atomicAdd(counter, (unsigned long long)kRows);
        first = __shfl_sync(0xffffffffu, first, 0);
        first += (unsigned long long)r.owner_lo;
        if ((int64_t)first >= r.owner_hi) break;
        const int64_t last = min((int64_t)first + kRows, (int64_t)r.owner_hi);
        if constexpr (kDealt) {
            const int64_t cand = (int64_t)first + lane;
            int64_t cand_a0 = 0;
            int cand_d = 0;
            if (cand < last && r.deal[cand] == r.deal_rank) {
                cand_a0 = __ldg(indptr + cand);
                cand_d = (int)(__ldg(indptr + cand + 1) - cand_a0);
                if (cand_d > kWarpOwnerMax) cand_d = 0;
            }
            unsigned todo = __ballot_sync(0xffffffffu, cand_d > 0);
            while (todo) {
                const int pick = __ffs(todo) - 1;
                todo &= todo - 1;
                score_owner((int32_t)first + pick, __shfl_sync(0xffffffffu, cand_a0, pick), __shfl_sync(0xffffffffu, cand_d, pick));
            }
        } else {
            for (int64_t o64 = (int64_t)first; o64 < last; ++o64) {
                const int64_t a0 = __ldg(indptr + o64);
                const int d_o = (int)(__ldg(indptr + o64 + 1) - a0);
                if (d_o == 0 || d_o > kWarpOwnerMax) continue;
                score_owner((int32_t)o64, a0, d_o);
            }
        }
    }
}

// ---- level B: CTA per (owner, neighbour chunk), hash tiles in shared memory ----------------------------------
struct OwnerItem {
    int32_t owner;
    int32_t first;  // index of the chunk's first neighbour inside the owner's row
};

// Two size classes share one kernel; the launch picks block size, hash slots and chunk length.
struct OwnerClass {
    int slots;      // hash slots per tile (power of two); a tile holds slots/2 owner ids
    int chunk;      // neighbours per work item
    int threads;
};
// slots = both cuckoo tables together; a tile holds at most kTileLoad * slots owner ids (load factor 0.375)
constexpr int kMediumMaxDegree = 1536;
constexpr OwnerClass kMediumClass{4096, 512, 256};       // 16 KB tables + 16 KB state + 4 KB queues: ~6 CTAs / SM
constexpr OwnerClass kHubClass{16384, 1024, 768};        // 64 KB tables + 32 KB state + 13 KB queues: 2 CTAs / SM (40 regs x 1536 threads)
__host__ __device__ constexpr int tile_ids_for(int slots) { return slots / 8 * 3; }

__host__ __device__ inline size_t owner_smem_bytes(const OwnerClass& c, bool ordered_sum) {
    // slots | base(int64) | acc(double) | len | cursor | cnt | top | per-warp hit queues (Adamic-Adar only); chunk is even
    return sizeof(int32_t) * (size_t)c.slots + (size_t)c.chunk * (8 + 8 + 4 + 4 + 4 + 4) +
           (ordered_sum ? (size_t)(c.threads / kWarp) * kQueueStride * sizeof(double) : 0);
}

// Ordered accumulation of two consecutive 32-id groups (group 0 holds the larger ids): the hit lanes park their terms in
// a warp-private shared queue in descending-id order, then every lane replays the queue (broadcast loads) with the
// sequential fp64 adds the reference's SpGEMM performs — ~3 issue slots per hit instead of ~8 for a ballot/shuffle
// loop; one padding store, one pair of warp barriers and one replay loop serve up to 64 hits.
__device__ __forceinline__ void accumulate_hits2(bool hit0, double w0, bool hit1, double w1, double* queue, double& acc,
                                                 int& nhits) {
    const unsigned h0 = __ballot_sync(0xffffffffu, hit0);
    const unsigned h1 = __ballot_sync(0xffffffffu, hit1);
    if ((h0 | h1) == 0) return;
    const unsigned lt = (1u << lane_id()) - 1u;
    const int n0 = __popc(h0);
    const int n = n0 + __popc(h1);
    nhits += n;
    if (hit0) queue[__popc(h0 & lt)] = w0;
    if (hit1) queue[n0 + __popc(h1 & lt)] = w1;
    if (lane_id() < 3) queue[n + lane_id()] = 0.0;
    __syncwarp();
    for (int h = 0; h < n; h += 4) {
        const double2 a = *reinterpret_cast<const double2*>(queue + h);
        const double2 b = *reinterpret_cast<const double2*>(queue + h + 2);
        acc = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(acc, a.x), a.y), b.x), b.y);
    }
    __syncwarp();
}

// Process the part of row(w) below `cursor` that lies in the current tile's id range [lo_id, +inf), walking DOWN
// (ids descending: SciPy's Adamic-Adar accumulation order; Jaccard does not care). Several 32-id groups are loaded
// per round so a long row keeps several requests in flight. Out-of-row lanes carry an INT_MIN sentinel that fails
// every test. kBounded = false is the owner's lowest tile (no id bound: the rest of the row is consumed) — the only
// pass single-tile owners ever run. Returns the new cursor.
template <int kMode, bool kBounded>
__device__ __forceinline__ int stream_down(const int32_t* __restrict__ row_w, int cursor, int32_t lo_id,
                                           const Cuckoo& table, const double* __restrict__ node_w, double* queue,
                                           int& count, double& acc) {
    const int lane = lane_id();
    int c = 0;    // per-lane hits (kMode 0)
    int nh = 0;   // warp-uniform hits (kMode 2)
    if (!kBounded) {
        constexpr int D = kMode == 0 ? kJaccardDepth : 4;   // 32-id groups fetched per round
        int top = cursor - 1 - lane;
        if (table.stash_n == 0) {
            // full rounds: every lane of every group is inside the row, so no bound tests, no sentinels, no stash scan
            // (rows streamed by the CTA classes average several hundred ids: ~85 % of all groups take this loop)
            for (int rounds = cursor / (D * kWarp); rounds > 0; --rounds, top -= D * kWarp) {
                int32_t x[D];
#pragma unroll
                for (int k = 0; k < D; ++k) x[k] = __ldg(row_w + top - k * kWarp);
                if (kMode == 0) {
#pragma unroll
                    for (int k = 0; k < D; ++k) cuckoo_count(table, x[k], c);
                } else {
                    bool hit[D];
                    double w[D];
#pragma unroll
                    for (int k = 0; k < D; ++k) hit[k] = cuckoo_hit(table, x[k]);
#pragma unroll
                    for (int k = 0; k < D; ++k) w[k] = hit[k] ? __ldg(node_w + x[k]) : 0.0;
#pragma unroll
                    for (int k = 0; k < D; k += 2) accumulate_hits2(hit[k], w[k], hit[k + 1], w[k + 1], queue, acc, nh);
                }
            }
        }
        for (; top + lane >= 0; top -= D * kWarp) {
            int32_t x[D];
#pragma unroll
            for (int k = 0; k < D; ++k) x[k] = top - k * kWarp >= 0 ? __ldg(row_w + top - k * kWarp) : INT_MIN;
            if (kMode == 0) {
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    if (k > 0 && top + lane - k * kWarp < 0) break;   // whole group below the row start (warp-uniform)
                    c += cuckoo_contains(table, x[k]);
                }
            } else {
                // probe all four groups, then issue all weight gathers, then accumulate in order: the gather latency
                // of a round is paid once instead of once per group
                bool hit[D];
                double w[D];
#pragma unroll
                for (int k = 0; k < D; ++k) hit[k] = cuckoo_contains(table, x[k]);
#pragma unroll
                for (int k = 0; k < D; ++k) w[k] = hit[k] ? __ldg(node_w + x[k]) : 0.0;
#pragma unroll
                for (int k = 0; k < D; k += 2) accumulate_hits2(hit[k], w[k], hit[k + 1], w[k + 1], queue, acc, nh);
            }
        }
        cursor = 0;
    } else if (kMode != 0) {
        // Branch-free rounds of two groups: ids below the tile are not in its table, so both groups are probed and
        // accumulated wherever the boundary falls; the in-tile ids are a prefix of the 64 visited (ids descend).
        const bool no_stash = table.stash_n == 0;
        if (no_stash) {
            // While the 128 ids below the cursor all lie inside the tile (ids descend: the lowest of them decides, one
            // uniform load from the line the round reads anyway) they take the bound-free four-group round of the lowest
            // tile: ~36 instead of ~53 instructions per 32 ids (ncu source page, round 2); only the ragged end of the piece
            // goes through the two-group rounds below.
            constexpr int D = 4;
            while (cursor >= D * kWarp && __ldg(row_w + cursor - D * kWarp) >= lo_id) {
                const int top = cursor - 1 - lane;
                int32_t x[D];
#pragma unroll
                for (int k = 0; k < D; ++k) x[k] = __ldg(row_w + top - k * kWarp);
                bool hit[D];
                double w[D];
#pragma unroll
                for (int k = 0; k < D; ++k) hit[k] = cuckoo_hit(table, x[k]);
#pragma unroll
                for (int k = 0; k < D; ++k) w[k] = hit[k] ? __ldg(node_w + x[k]) : 0.0;
#pragma unroll
                for (int k = 0; k < D; k += 2) accumulate_hits2(hit[k], w[k], hit[k + 1], w[k + 1], queue, acc, nh);
                cursor -= D * kWarp;
            }
        }
        while (cursor > 0) {
            const int top = cursor - 1 - lane;
            const int32_t x0 = top >= 0 ? __ldg(row_w + top) : INT_MIN;
            const int32_t x1 = top - kWarp >= 0 ? __ldg(row_w + top - kWarp) : INT_MIN;
            bool h0, h1;
            if (no_stash) {
                h0 = cuckoo_hit(table, x0);
                h1 = cuckoo_hit(table, x1);
            } else {
                h0 = cuckoo_contains(table, x0);
                h1 = cuckoo_contains(table, x1);
            }
            const double w0 = h0 ? __ldg(node_w + x0) : 0.0, w1 = h1 ? __ldg(node_w + x1) : 0.0;
            const int used = __popc(__ballot_sync(0xffffffffu, x0 >= lo_id)) + __popc(__ballot_sync(0xffffffffu, x1 >= lo_id));
            accumulate_hits2(h0, w0, h1, w1, queue, acc, nh);
            cursor -= used;
            if (used < 2 * kWarp) break;   // ran off the tile (or the row: cursor == 0)
        }
    } else {
        bool done = false;
        int groups = 1;   // a tile usually holds a short piece of the row: probe one group before going four deep
        while (cursor > 0 && !done) {
            const int top = cursor - 1 - lane;
            int32_t x[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) x[k] = (k < groups && top - k * kWarp >= 0) ? __ldg(row_w + top - k * kWarp) : INT_MIN;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (done || k >= groups) break;
                const bool in_tile = x[k] >= lo_id;   // sentinel lanes fail (lo_id > INT_MIN)
                const bool hit = in_tile && cuckoo_contains(table, x[k]);
                c += hit;
                const unsigned inside = __ballot_sync(0xffffffffu, in_tile);
                if (inside != 0xffffffffu) {           // ran off the tile (or the row): stop after this group
                    done = true;
                    cursor = cursor - k * kWarp - __popc(inside);
                }
            }
            if (!done) cursor -= groups * kWarp;
            groups = 4;
        }
        if (cursor < 0) cursor = 0;
    }
    if (kMode == 0) count += __reduce_add_sync(0xffffffffu, c);
    if (kMode == 2) count += nh;
    return cursor;
}

template <int kMode, bool kScatter>
__device__ __forceinline__ void cta_owner_body(const OwnerItem* __restrict__ items, int64_t num_items, OwnerClass cls,
                                 const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                 const int32_t* __restrict__ rev_off, RangeInfo r,
                                 const double* __restrict__ node_w, int32_t* __restrict__ inter_out,
                                 double* __restrict__ score_out, double* __restrict__ jaccard_out,
                                 unsigned long long* counter) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int32_t* slots = reinterpret_cast<int32_t*>(smem_raw);
    long long* base_s = reinterpret_cast<long long*>(smem_raw + sizeof(int32_t) * (size_t)cls.slots);
    double* acc_s = reinterpret_cast<double*>(base_s + cls.chunk);
    int32_t* len_s = reinterpret_cast<int32_t*>(acc_s + cls.chunk);   // row length, -1 = pair not evaluated here
    int32_t* cur_s = len_s + cls.chunk;                               // unprocessed prefix of row(w)
    int32_t* cnt_s = cur_s + cls.chunk;
    int32_t* top_s = cnt_s + cls.chunk;                               // largest unprocessed id of row(w) (INT_MIN: none)
    double* queue = reinterpret_cast<double*>(top_s + cls.chunk) + (threadIdx.x >> 5) * kQueueStride;   // valid when kMode != 0
    __shared__ long long item_s;
    __shared__ int next_s;
    const int lane = lane_id();
    const int nthreads = blockDim.x;
    const int tile_ids = tile_ids_for(cls.slots);
    __shared__ int32_t stash_s[kStashMax];
    __shared__ int stash_n_s;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) item_s = (long long)atomicAdd(counter, 1ull);
        __syncthreads();
        const int64_t item = item_s;
        if (item >= num_items) break;
        const int32_t o = items[item].owner;
        if (o < r.owner_lo || o >= r.owner_hi) continue;
        const int j0 = items[item].first;
        const int64_t a0 = __ldg(indptr + o);
        const int d_o = (int)(__ldg(indptr + o + 1) - a0);
        const int nb = min(cls.chunk, d_o - j0);
        const int32_t* row_o = indices + a0;
        if (!r.full) {
            const int64_t c0 = a0 + j0, c1 = c0 + nb;
            if (!(c0 < r.e_end && c1 > r.e_begin)) {
                const int32_t w_first = __ldg(row_o + j0), w_last = __ldg(row_o + j0 + nb - 1);
                if (w_last < r.row_lo || w_first > r.row_hi) continue;
            }
        }
        // neighbour metadata once per item: row start, length (or -1 when the pair is evaluated elsewhere)
        for (int i = threadIdx.x; i < nb; i += nthreads) {
            const int32_t w = __ldg(row_o + j0 + i);
            const int64_t b0 = __ldg(indptr + w);
            const int d_w = (int)(__ldg(indptr + w + 1) - b0);
            bool skip = other_owns(d_w, w, d_o, o);
            if (!skip && !r.full) {
                const int64_t p1 = a0 + j0 + i;
                skip = !(p1 >= r.e_begin && p1 < r.e_end) && !(b0 < r.e_end && b0 + d_w > r.e_begin);
            }
            base_s[i] = b0;
            len_s[i] = skip ? -1 : d_w;
            cur_s[i] = d_w;
            top_s[i] = (skip || d_w == 0) ? INT_MIN : __ldg(indices + b0 + d_w - 1);
            cnt_s[i] = 0;
            acc_s[i] = 0.0;
        }
        const int num_tiles = (d_o + tile_ids - 1) / tile_ids;
        for (int t = num_tiles - 1; t >= 0; --t) {
            const int ts = t * tile_ids, te = min(d_o, ts + tile_ids);
            int cap = 32;  // slots per table: smallest power of two with (te - ts) <= 0.75 * cap
            while (3 * cap < 4 * (te - ts)) cap <<= 1;
            const int shift = 32 - (31 - __clz(cap));
            int32_t* t1 = slots;
            int32_t* t2 = slots + cap;
            uint32_t mul1 = 0x9E3779B1u, mul2 = 0x85EBCA77u;
            for (;;) {
                __syncthreads();  // previous tile's probes (and the metadata writes) are done
                for (int i = threadIdx.x; i < 2 * cap; i += nthreads) slots[i] = -1;
                if (threadIdx.x == 0) { next_s = 0; stash_n_s = 0; }
                __syncthreads();
                for (int i = ts + threadIdx.x; i < te; i += nthreads)
                    cuckoo_insert(t1, t2, shift, mul1, mul2, __ldg(row_o + i), stash_s, &stash_n_s);
                __syncthreads();
                if (stash_n_s <= kStashMax) break;
                mul1 += 0x3C6EF372u;  // stash overflow (practically never): rebuild with other multipliers (kept odd)
                mul2 += 0x1B873592u;
            }
            const Cuckoo table{(uint32_t)__cvta_generic_to_shared(t1), (uint32_t)__cvta_generic_to_shared(t2), mul1, mul2, shift,
                               stash_n_s, stash_s};
            const int32_t lo_id = t == 0 ? INT_MIN + 1 : __ldg(row_o + ts);
            for (;;) {
                int i = 0;
                if (lane == 0) i = atomicAdd(&next_s, 4);
                i = __shfl_sync(0xffffffffu, i, 0);
                if (i >= nb) break;
                const int i_end = min(i + 4, nb);
                for (; i < i_end; ++i) {
                    // rows with nothing left in this tile's id range cost one shared-memory read (multi-tile owners chop
                    // every neighbour row into pieces; most tiles miss most short rows)
                    if (top_s[i] < lo_id) continue;
                    const int cursor = cur_s[i];
                    int count = 0;
                    double acc = kMode != 0 ? acc_s[i] : 0.0;
                    const int32_t* row_w = indices + base_s[i];
                    const int new_cursor = t == 0 ? stream_down<kMode, false>(row_w, cursor, lo_id, table, node_w, queue, count, acc)
                                                  : stream_down<kMode, true>(row_w, cursor, lo_id, table, node_w, queue, count, acc);
                    if (lane == 0) {
                        if (kMode != 1) cnt_s[i] += count;
                        if (kMode != 0) acc_s[i] = acc;
                        if (t > 0) {   // the lowest tile consumes the rest of the row: nothing to carry over
                            cur_s[i] = new_cursor;
                            top_s[i] = new_cursor > 0 ? __ldg(row_w + new_cursor - 1) : INT_MIN;
                        }
                    }
                }
            }
        }
        __syncthreads();
        if constexpr (kScatter) {
            scatter_chunk(r.scatter, rev_off, base_s, cls.chunk, nb, a0 + j0, d_o, kMode);
        } else {
            for (int i = threadIdx.x; i < nb; i += nthreads) {
                if (len_s[i] < 0) continue;  // pair owned by the neighbour, or outside the range
                const int64_t p1 = a0 + j0 + i;
                write_pair<kMode, kScatter>(r, p1, base_s[i] + __ldg(rev_off + p1), d_o, len_s[i], cnt_s[i], acc_s[i], inter_out,
                                            score_out, jaccard_out);
            }
        }
    }
}

#define GSP_OWNER_PARAMS                                                                                              \
    const OwnerItem *__restrict__ items, int64_t num_items, OwnerClass cls, const int64_t *__restrict__ indptr,             \
        const int32_t *__restrict__ indices, const int32_t *__restrict__ rev_off, RangeInfo r,                              \
        const double *__restrict__ node_w, int32_t *__restrict__ inter_out, double *__restrict__ score_out,                 \
        double *__restrict__ jaccard_out, unsigned long long *counter
#define GSP_OWNER_ARGS items, num_items, cls, indptr, indices, rev_off, r, node_w, inter_out, score_out, jaccard_out, counter

// No launch bounds: the natural allocation is 32-40 registers (two 768-thread hub CTAs per SM); bounding every
// instantiation measured 5 % slower. The peer-scatter instantiations reach 40 only with their per-row epilogue out of
// line (`scatter_chunk`); inlined, the fused one took 44-47 and needed a cap that cost ~6 % of every launch.
template <int kMode, bool kScatter>
__global__ void cta_owner_kernel(GSP_OWNER_PARAMS) {
    cta_owner_body<kMode, kScatter>(GSP_OWNER_ARGS);
}

// ---- work items for level B (built once per graph) -----------------------------------------------------------
__device__ __forceinline__ void item_counts(int64_t d, int64_t& medium, int64_t& hub) {
    medium = (d > kWarpOwnerMax && d <= kMediumMaxDegree) ? (d + kMediumClass.chunk - 1) / kMediumClass.chunk : 0;
    hub = d > kMediumMaxDegree ? (d + kHubClass.chunk - 1) / kHubClass.chunk : 0;
}

__global__ void count_items_kernel(int64_t n, const int64_t* __restrict__ indptr, int64_t* __restrict__ medium,
                                   int64_t* __restrict__ hub) {
    for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x)
        item_counts(indptr[o + 1] - indptr[o], medium[o], hub[o]);
}

__global__ void fill_items_kernel(int64_t n, const int64_t* __restrict__ indptr, const int64_t* __restrict__ medium_incl,
                                  const int64_t* __restrict__ hub_incl, OwnerItem* __restrict__ medium_items,
                                  OwnerItem* __restrict__ hub_items) {
    for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) {
        int64_t m, h;
        item_counts(indptr[o + 1] - indptr[o], m, h);
        OwnerItem* dst = m ? medium_items + (medium_incl[o] - m) : hub_items + (hub_incl[o] - h);
        const int chunk = m ? kMediumClass.chunk : kHubClass.chunk;
        for (int64_t k = 0; k < m + h; ++k) dst[k] = OwnerItem{(int32_t)o, (int32_t)(k * chunk)};
    }
}

// Sort key of a work item: the id of the first neighbour its chunk covers. Items sorted by it make the kernel sweep the id
// space once with the chunks of ALL owners that cover an id window running at about the same time, so a neighbour row that
// several owners stream is fetched from HBM once and then served by L2 (without the sort every use came from DRAM: ncu
// measured 4*M bytes of DRAM reads).
__global__ void item_keys_kernel(int64_t count, const OwnerItem* __restrict__ items, const int64_t* __restrict__ indptr,
                                 const int32_t* __restrict__ indices, uint32_t* __restrict__ keys) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
        keys[i] = (uint32_t)indices[indptr[items[i].owner] + items[i].first];
}

int sort_items_by_window(OwnerItem* items, int64_t count, const Graph* g, cudaStream_t s) {
    if (count <= 1) return GSP_OK;
    static_assert(sizeof(OwnerItem) == sizeof(uint64_t), "items are sorted as 64-bit values");
    Scratch<uint32_t> keys, keys_sorted;
    Scratch<uint64_t> sorted;
    GSP_CUDA_TRY(keys.alloc(count, s));
    GSP_CUDA_TRY(keys_sorted.alloc(count, s));
    GSP_CUDA_TRY(sorted.alloc(count, s));
    item_keys_kernel<<<grid_for(count, 256), 256, 0, s>>>(count, items, g->indptr, g->indices, keys.ptr);
    GSP_CHECK_LAUNCH();
    if (int rc = sort_pairs_u32_u64(keys.ptr, keys_sorted.ptr, reinterpret_cast<const uint64_t*>(items), sorted.ptr, count, s)) return rc;
    GSP_CUDA_TRY(cudaMemcpyAsync(items, sorted.ptr, (size_t)count * sizeof(uint64_t), cudaMemcpyDeviceToDevice, s));
    return GSP_OK;
}

__global__ void range_rows_kernel(const int32_t* __restrict__ rows, int64_t e_begin, int64_t e_end, int32_t* out) {
    out[0] = rows[e_begin];
    out[1] = rows[e_end - 1];
}

// w -> w * w once per call: the Adamic-Adar term of a common neighbour is the rounded square of its weight, so the
// streaming kernels gather the term itself (same bits as squaring after the gather, one fp64 multiply less per hit)
__global__ void square_weights_kernel(int64_t n, const double* __restrict__ w, double* __restrict__ w2) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        w2[i] = __dmul_rn(w[i], w[i]);
}

std::mutex g_items_mutex;

int ensure_items(Graph* g, cudaStream_t s) {
    std::lock_guard<std::mutex> lock(g_items_mutex);
    if (g->owner_items_ready) return GSP_OK;
    Scratch<int64_t> cm, ch, im, ih;
    GSP_CUDA_TRY(cm.alloc(g->n, s));
    GSP_CUDA_TRY(ch.alloc(g->n, s));
    GSP_CUDA_TRY(im.alloc(g->n, s));
    GSP_CUDA_TRY(ih.alloc(g->n, s));
    count_items_kernel<<<grid_for(g->n, 256), 256, 0, s>>>(g->n, g->indptr, cm.ptr, ch.ptr);
    GSP_CHECK_LAUNCH();
    if (int rc = inclusive_sum_i64(cm.ptr, im.ptr, g->n, s)) return rc;
    if (int rc = inclusive_sum_i64(ch.ptr, ih.ptr, g->n, s)) return rc;
    int64_t totals[2] = {0, 0};
    GSP_CUDA_TRY(cudaMemcpyAsync(&totals[0], im.ptr + (g->n - 1), sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    GSP_CUDA_TRY(cudaMemcpyAsync(&totals[1], ih.ptr + (g->n - 1), sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    GSP_CUDA_TRY(cudaStreamSynchronize(s));
    if (totals[0] + totals[1] > 0) {
        OwnerItem* items = nullptr;   // medium items first, hub items after them
        GSP_CUDA_TRY(device_alloc(&items, (size_t)(totals[0] + totals[1]) * sizeof(OwnerItem), s));
        fill_items_kernel<<<grid_for(g->n, 256), 256, 0, s>>>(g->n, g->indptr, im.ptr, ih.ptr, items, items + totals[0]);
        GSP_CHECK_LAUNCH();
        const char* order = getenv("GSP_ITEM_ORDER");   // "owner" keeps the owner-major order (for A/B measurements)
        if (!(order && order[0] == 'o')) {
            if (int rc = sort_items_by_window(items, totals[0], g, s)) { device_free(items); return rc; }
            if (int rc = sort_items_by_window(items + totals[0], totals[1], g, s)) { device_free(items); return rc; }
        }
        GSP_CUDA_TRY(cudaStreamSynchronize(s));
        g->owner_items = items;
    }
    g->num_owner_items = totals[0];
    g->num_hub_items = totals[1];
    g->owner_items_ready = true;
    return GSP_OK;
}

// ---- the items of one owner range (owner-sharded scoring) ---------------------------------------------------------
__global__ void flag_owned_kernel(int64_t count, const OwnerItem* __restrict__ items, int32_t lo, int32_t hi,
                                  const uint8_t* __restrict__ deal, int32_t deal_rank, int64_t* __restrict__ flags) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t o = items[i].owner;
        flags[i] = o >= lo && o < hi && (!deal || deal[o] == deal_rank);
    }
}

__global__ void gather_owned_kernel(int64_t count, const OwnerItem* __restrict__ items, const int64_t* __restrict__ flags,
                                    const int64_t* __restrict__ incl, OwnerItem* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
        if (flags[i]) out[incl[i] - 1] = items[i];      // order kept: the id-window order of the full list
}

// Order-preserving filter of both item classes by owner range; cached for the last range used.
int ensure_owned_items(Graph* g, int64_t owner_lo, int64_t owner_hi, bool dealt, cudaStream_t s) {
    std::lock_guard<std::mutex> lock(g_items_mutex);
    if (g->owned_items && g->owned_lo == owner_lo && g->owned_hi == owner_hi && g->owned_dealt == dealt) return GSP_OK;
    const int64_t total = g->num_owner_items + g->num_hub_items;
    if (g->owned_items) {     // kernels of an earlier call (any stream) may still read the list of the previous range
        GSP_CUDA_TRY(cudaDeviceSynchronize());
        device_free(g->owned_items);
    }
    g->owned_items = nullptr;
    g->num_owned_items = g->num_owned_hub_items = 0;
    g->owned_lo = owner_lo;
    g->owned_hi = owner_hi;
    g->owned_dealt = dealt;
    if (total == 0) return GSP_OK;
    const OwnerItem* items = reinterpret_cast<const OwnerItem*>(g->owner_items);
    Scratch<int64_t> flags, incl;
    GSP_CUDA_TRY(flags.alloc(total, s));
    GSP_CUDA_TRY(incl.alloc(total, s));
    flag_owned_kernel<<<grid_for(total, 256), 256, 0, s>>>(total, items, (int32_t)owner_lo, (int32_t)owner_hi,
                                                           dealt ? g->deal : nullptr, g->deal_rank, flags.ptr);
    GSP_CHECK_LAUNCH();
    int64_t kept[2] = {0, 0};
    const int64_t counts[2] = {g->num_owner_items, g->num_hub_items};
    int64_t off = 0;
    for (int c = 0; c < 2; ++c) {      // medium items first, hub items after them (each class scanned on its own)
        if (counts[c] > 0) {
            if (int rc = inclusive_sum_i64(flags.ptr + off, incl.ptr + off, counts[c], s)) return rc;
            GSP_CUDA_TRY(cudaMemcpyAsync(&kept[c], incl.ptr + off + counts[c] - 1, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
        }
        off += counts[c];
    }
    GSP_CUDA_TRY(cudaStreamSynchronize(s));
    OwnerItem* out = nullptr;
    GSP_CUDA_TRY(device_alloc(&out, (size_t)(kept[0] + kept[1] ? kept[0] + kept[1] : 1) * sizeof(OwnerItem), s));
    off = 0;
    int64_t out_off = 0;
    for (int c = 0; c < 2; ++c) {
        if (counts[c] > 0) {
            gather_owned_kernel<<<grid_for(counts[c], 256), 256, 0, s>>>(counts[c], items + off, flags.ptr + off, incl.ptr + off, out + out_off);
            GSP_CHECK_LAUNCH();
        }
        off += counts[c];
        out_off += kept[c];
    }
    g->owned_items = out;
    g->num_owned_items = kept[0];
    g->num_owned_hub_items = kept[1];
    return GSP_OK;
}

// Debug/tuning knobs: GSP_HUB_CLASS="slots,chunk,threads,ctas_per_sm" overrides the hub launch configuration
// (the chunk must not exceed the value the work items were built with).
OwnerClass hub_class_from_env(int& ctas_per_sm) {
    OwnerClass c = kHubClass;
    if (const char* env = getenv("GSP_HUB_CLASS")) {
        int slots, chunk, threads, ctas;
        if (sscanf(env, "%d,%d,%d,%d", &slots, &chunk, &threads, &ctas) == 4 && chunk == kHubClass.chunk) {
            c = OwnerClass{slots, chunk, threads};
            ctas_per_sm = ctas;
        }
    }
    return c;
}

template <int kMode, bool kScatter>
int launch_class(const OwnerClass& cls, const OwnerItem* items, int64_t count, int ctas_per_sm, Graph* g, const RangeInfo& r,
                 const double* node_w, int32_t* inter, double* score, double* jaccard, unsigned long long* counter,
                 cudaStream_t s) {
    if (count <= 0) return GSP_OK;
    const size_t smem = owner_smem_bytes(cls, kMode != 0);
    auto kernel = cta_owner_kernel<kMode, kScatter>;
    GSP_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    int64_t blocks = (int64_t)kNumSMs * ctas_per_sm;
    if (blocks > count) blocks = count;
    kernel<<<(int)blocks, cls.threads, smem, s>>>(items, count, cls, g->indptr, g->indices, g->rev_off, r, node_w, inter, score, jaccard,
                                                  counter);
    GSP_CHECK_LAUNCH();
    return GSP_OK;
}

template <int kMode, bool kScatter>
int launch(Graph* g, int64_t e_begin, int64_t e_end, int64_t owner_lo, int64_t owner_hi, const double* node_w, int32_t* inter,
           double* score, double* jaccard, cudaStream_t s, bool dealt, double* const* slices = nullptr, int64_t slice_len = 0,
           double* const* slices2 = nullptr) {
    if (g->n == 0 || e_end == e_begin || owner_hi <= owner_lo) return GSP_OK;
    if (int rc = ensure_items(g, s)) return rc;
    RangeInfo r{e_begin, e_end, 0, 0, e_begin == 0 && e_end == g->nnz, (int32_t)owner_lo, (int32_t)owner_hi,
                dealt ? g->deal : nullptr, g->deal_rank, slices, slice_len, slices2, slice_len > 0 ? 1.0 / (double)slice_len : 0.0, nullptr};
    Scratch<ScatterInfo> scatter_info;
    if (kScatter) {
        GSP_CUDA_TRY(scatter_info.alloc(1, s));
        fill_scatter_info_kernel<<<1, 1, 0, s>>>(scatter_info.ptr, slices, kMode == 2 ? slices2 : nullptr, slice_len);
        GSP_CHECK_LAUNCH();
        r.scatter = scatter_info.ptr;
    }
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
    Scratch<double> squared;
    if (kMode != 0) {
        GSP_CUDA_TRY(squared.alloc(g->n, s));
        square_weights_kernel<<<grid_for(g->n, 256), 256, 0, s>>>(g->n, node_w, squared.ptr);
        GSP_CHECK_LAUNCH();
        node_w = squared.ptr;
    }
    Scratch<unsigned long long> counters;
    GSP_CUDA_TRY(counters.alloc(3, s));
    GSP_CUDA_TRY(cudaMemsetAsync(counters.ptr, 0, 3 * sizeof(unsigned long long), s));
    const OwnerItem* items = reinterpret_cast<const OwnerItem*>(g->owner_items);
    int64_t num_medium = g->num_owner_items, num_hub = g->num_hub_items;
    if (owner_lo > 0 || owner_hi < g->n || r.deal) {   // owner-sharded call: only this rank's items are claimed
        if (int rc = ensure_owned_items(g, owner_lo, owner_hi, r.deal != nullptr, s)) return rc;
        items = reinterpret_cast<const OwnerItem*>(g->owned_items);
        num_medium = g->num_owned_items;
        num_hub = g->num_owned_hub_items;
    }
    // hubs first: their long work items should not land in the tail
    int hub_ctas = 2;
    const OwnerClass hub = hub_class_from_env(hub_ctas);
    if (int rc = launch_class<kMode, kScatter>(hub, items + num_medium, num_hub, hub_ctas, g, r, node_w, inter, score,
                                     jaccard, counters.ptr, s)) return rc;
    if (int rc = launch_class<kMode, kScatter>(kMediumClass, items, num_medium, 7, g, r, node_w, inter, score, jaccard,
                                     counters.ptr + 1, s)) return rc;
    if (r.deal) {
        const int64_t claims = (owner_hi - owner_lo + kADealtRowsPerClaim - 1) / kADealtRowsPerClaim;
        warp_owner_kernel<kMode, kScatter, true><<<grid_for(claims, kAWarps, 8), kAThreads, 0, s>>>(
            g->n, g->indptr, g->indices, g->rev_off, r, node_w, inter, score, jaccard, counters.ptr + 2);
    } else {
        const int64_t claims = (owner_hi - owner_lo + kARowsPerClaim - 1) / kARowsPerClaim;
        warp_owner_kernel<kMode, kScatter, false><<<grid_for(claims, kAWarps, 8), kAThreads, 0, s>>>(
            g->n, g->indptr, g->indices, g->rev_off, r, node_w, inter, score, jaccard, counters.ptr + 2);
    }
    GSP_CHECK_LAUNCH();
    return GSP_OK;
}

}  // namespace

// Estimated work of every owner, in streamed-id equivalents: the ids of the owned neighbours' rows + a constant per
// neighbour; the rows of an owner that needs several hash tiles are chopped into one piece per tile (a fixed cost per
// piece) and their ids go through the costlier bounded rounds. Calibrated on 4 GPUs (R-MAT scale 24, fused pass): with
// the plain id count the rank that holds the largest hubs ran 54.8 ms against 46.6-47.6 ms for the others. The cost of a
// piece (~80 bookkeeping instructions + a 64-id round that a short piece fills only in part, ncu source view) was then
// fitted on the one owner no deal can spread: node 0 of R-MAT scale 24 (232 559 neighbours = 38 tiles, 8.8 M pieces of
// ~14 ids) kept its rank ~2.3 ms behind the others at 2 and at 8 GPUs with 24 per piece.
constexpr double kPieceCost = 110.0;
__global__ void owner_cost_kernel(int64_t n, const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                  double* __restrict__ cost) {
    const int lane = lane_id();
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    constexpr int hub_tile = tile_ids_for(kHubClass.slots);
    for (int64_t o = warp; o < n; o += nwarps) {
        const int64_t a0 = indptr[o], a1 = indptr[o + 1];
        const int d_o = (int)(a1 - a0);
        const int tiles = d_o > hub_tile ? (d_o + hub_tile - 1) / hub_tile : 1;
        const double per_id = tiles > 1 ? 1.3 : 1.0, per_row = 16.0 + (tiles > 1 ? kPieceCost * tiles : 0.0);
        double c = 0.0;
        for (int64_t p = a0 + lane; p < a1; p += kWarp) {
            const int32_t w = __ldg(indices + p);
            const int d_w = (int)(__ldg(indptr + w + 1) - __ldg(indptr + w));
            c += other_owns(d_w, w, d_o, (int32_t)o) ? 2.0 : per_id * (double)d_w + per_row;
        }
        for (int off = 16; off; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
        if (lane == 0) cost[o] = c + (d_o ? 8.0 : 0.0);
    }
}

// Dealt ownership: a per-node rank byte replaces contiguous node ranges as the unit of owner sharding, so a caller can
// hand every rank the same mix of hub / medium / small owners (e.g. owners sorted by cost and dealt in snake order).
int set_owner_deal(Graph* g, const uint8_t* d_owner_rank, int32_t rank, cudaStream_t s) {
    std::lock_guard<std::mutex> lock(g_items_mutex);
    if (g->owned_items || g->deal) GSP_CUDA_TRY(cudaDeviceSynchronize());   // earlier kernels may still read the old lists
    device_free(g->owned_items);
    g->owned_items = nullptr;
    g->num_owned_items = g->num_owned_hub_items = 0;
    g->owned_lo = g->owned_hi = -1;
    device_free(g->deal);
    g->deal = nullptr;
    g->deal_rank = -1;
    if (d_owner_rank == nullptr || g->n == 0) return GSP_OK;
    GSP_CUDA_TRY(device_alloc(&g->deal, (size_t)g->n, s));
    GSP_CUDA_TRY(cudaMemcpyAsync(g->deal, d_owner_rank, (size_t)g->n, cudaMemcpyDeviceToDevice, s));
    g->deal_rank = rank;
    return GSP_OK;
}

int owner_intersect_jaccard(Graph* g, int64_t e_begin, int64_t e_end, int64_t owner_lo, int64_t owner_hi, int32_t* inter,
                            double* score, cudaStream_t s, bool dealt) {
    return launch<0, false>(g, e_begin, e_end, owner_lo, owner_hi, nullptr, inter, score, nullptr, s, dealt);
}

int owner_intersect_adamic_adar(Graph* g, int64_t e_begin, int64_t e_end, int64_t owner_lo, int64_t owner_hi,
                                const double* node_w, double* score, cudaStream_t s, bool dealt) {
    return launch<1, false>(g, e_begin, e_end, owner_lo, owner_hi, node_w, nullptr, score, nullptr, s, dealt);
}

// one streaming pass, both scores (and optionally the counts)
int owner_intersect_both(Graph* g, int64_t e_begin, int64_t e_end, int64_t owner_lo, int64_t owner_hi, const double* node_w,
                         int32_t* inter, double* jaccard, double* adamic_adar, cudaStream_t s, bool dealt) {
    return launch<2, false>(g, e_begin, e_end, owner_lo, owner_hi, node_w, inter, adamic_adar, jaccard, s, dealt);
}

// mode 0: Jaccard into `slices`; 1: Adamic-Adar into `slices`; 2: Adamic-Adar into `slices`, Jaccard into `slices2`
int owner_intersect_scatter(Graph* g, int mode, int64_t owner_lo, int64_t owner_hi, const double* node_w,
                            double* const* slices, int64_t slice_len, cudaStream_t s, double* const* slices2) {
    if (mode == 0) return launch<0, true>(g, 0, g->nnz, owner_lo, owner_hi, nullptr, nullptr, nullptr, nullptr, s, true, slices, slice_len);
    if (mode == 1) return launch<1, true>(g, 0, g->nnz, owner_lo, owner_hi, node_w, nullptr, nullptr, nullptr, s, true, slices, slice_len);
    return launch<2, true>(g, 0, g->nnz, owner_lo, owner_hi, node_w, nullptr, nullptr, nullptr, s, true, slices, slice_len, slices2);
}

int owner_costs(const Graph* g, double* cost, cudaStream_t s) {
    if (g->n == 0) return GSP_OK;
    owner_cost_kernel<<<grid_for(g->n, 8, 8), 256, 0, s>>>(g->n, g->indptr, g->indices, cost);
    GSP_CHECK_LAUNCH();
    return GSP_OK;
}

}  // namespace gsp
