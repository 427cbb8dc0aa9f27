// approx_er.cu — approximate effective resistance: JL projection + batched Laplacian CG (K5, K6, K7).
//
// Replaces reference src/sparsification/metrics.py:232-298. The reference runs k independent
// scipy.sparse.linalg.cg solves one after another (k = 2 674 at the Roman-empire shape); here all
// right-hand sides of a column block advance together: vectors are [n, k] row-major so a CSR row
// gathers k-contiguous segments of p (256-byte coalesced requests per neighbour), the SpMM is fused
// with the p.q reduction and the axpy pair with the r.r reduction, and every column keeps SciPy's
// semantics on its own (x0 = 0; test ||r|| < rtol*||b|| BEFORE each update; at most max_iters updates;
// a column that stops is frozen while the others continue; ||b|| = 0 -> x = 0).
//
// Arithmetic is fp64 with separately rounded multiply/add (SciPy's `x += alpha*p` forms the product
// first). Row sums inside the SpMM follow csr_matvec's order (neighbours ascending, the diagonal at
// its sorted position). Column reductions are two-stage and deterministic (fixed block order), so a
// run is bit-reproducible; they are NOT BLAS ddot's order — ApproxER parity is the 1e-4 relative
// tolerance BASELINE.json states, not bit equality.
#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace gsp {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / kWarp;
constexpr int kCheckEvery = 8;  // CG iterations between host reads of the active-column count

struct CgColumns {        // per-column solver state, arrays of length k
    double* rho;          // r.r at the top of the current iteration
    double* rho_prev;
    double* alpha;
    double* beta;
    double* atol;
    int* active;
    int32_t* iters;
};

// Lane -> (row slot, column). With k >= 32 a warp works on one row and a strip of 32 columns (blockIdx.y picks the
// strip). With k < 32 (projection columns sharded over several GPUs leave 8 or 16 per rank) the columns are padded to
// kc = 2^ceil(log2 k) and a warp works on 32/kc rows at once, so no lane idles and a row's kc values are one
// contiguous request.
struct LaneMap {
    int c;        // column of this lane
    int sub;      // row slot inside the warp
    int rows;     // rows per warp pass
    int kc;       // padded columns per row slot (32 in strip mode)
    bool col_ok;
};

__device__ __forceinline__ LaneMap lane_map(int k) {
    LaneMap m;
    const int lane = lane_id();
    if (k >= kWarp) {
        m.c = blockIdx.y * kWarp + lane; m.sub = 0; m.rows = 1; m.kc = kWarp;
    } else {
        int kc = 1;
        while (kc < k) kc <<= 1;
        m.c = lane & (kc - 1); m.sub = lane / kc; m.rows = kWarp / kc; m.kc = kc;
    }
    m.col_ok = m.c < k;
    return m;
}

// ---- projection entries generated where they are used ----------------------------------------------------------------
// R[e, c] ~ N(0, 1) / sqrt(k) as a pure function of (seed, e, c): Philox4x32-10 keyed by the seed with the counter
// (e, c), two 53-bit uniforms, Box-Muller. Both endpoints of an edge evaluate the same entry, so the [m, k] fp64 matrix
// (31.7 GB at the products shape with k = 64) is never materialised; gsp_philox_projection writes the same entries out for
// tests and for callers that want to inspect the matrix.
struct Projection {
    const double* R;     // explicit matrix (parity mode: the reference's NumPy PCG64 draws), or nullptr
    int64_t ldr;
    uint64_t seed;       // generated mode
    int32_t col0;        // global index of local column 0 (column sharding)
    double scale;        // 1 / sqrt(total columns)
};

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}

__device__ __forceinline__ double philox_normal(uint64_t seed, int64_t e, int32_t c) {
    uint32_t ctr[4] = {(uint32_t)e, (uint32_t)((uint64_t)e >> 32), (uint32_t)c, 0x6a09e667u};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(ctr, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    // (0, 1) uniforms from 53 random bits each
    const double u1 = ((double)((((uint64_t)ctr[0] << 32) | ctr[1]) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    const double u2 = ((double)((((uint64_t)ctr[2] << 32) | ctr[3]) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

__device__ __forceinline__ double projection_entry(const Projection& pr, int64_t e, int32_t c) {
    return pr.R ? __ldg(pr.R + e * pr.ldr + c) : philox_normal(pr.seed, e, pr.col0 + c) * pr.scale;
}

__global__ void philox_projection_kernel(int64_t m, int32_t k, Projection pr, double* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m * k; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = projection_entry(pr, i / k, (int32_t)(i % k));
}

// Y[u, c] = sum over incident undirected edges e (ascending e) of +-R[e, c]   (metrics.py:260-275).
// Same row-segment work items as the SpMM (a hub row's 10^5 incident edges are spread over many warps); rows longer
// than one segment are finished by project_combine_kernel in segment order.
struct SegItem {
    int32_t row;
    int32_t seg;   // segment index inside the row
};
constexpr int kSeg = 512;

__global__ void __launch_bounds__(kThreads)
project_kernel(const SegItem* __restrict__ items, int64_t num_items, const int64_t* __restrict__ seg_incl,
               const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
               const int32_t* __restrict__ und_id, Projection pr, int k,
               double* __restrict__ Y, double* __restrict__ segpart) {
    const LaneMap m = lane_map(k);
    const int64_t warp = blockIdx.x * (int64_t)kWarps + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kWarps;
    for (int64_t base = warp * m.rows; base < num_items; base += nwarps * m.rows) {
        const int64_t it = base + m.sub;
        if (it >= num_items) continue;
        const int64_t u = items[it].row;
        const int seg = items[it].seg;
        const int64_t p0 = indptr[u] + (int64_t)seg * kSeg;
        const int64_t p1 = min(p0 + kSeg, indptr[u + 1]);
        double acc = 0.0;
        for (int64_t p = p0; p < p1; ++p) {
            const int32_t e = __ldg(und_id + p);
            if (e < 0) continue;                                    // self loop / unmatched direction
            const int32_t v = __ldg(indices + p);
            if (m.col_ok) {
                const double r = projection_entry(pr, e, m.c);
                acc = (u < v) ? __dadd_rn(acc, r) : __dsub_rn(acc, r);   // +1 * r / -1 * r are exact
            }
        }
        if (m.col_ok) {
            if (seg == 0) Y[u * (int64_t)k + m.c] = acc;
            else segpart[((u ? seg_incl[u - 1] : 0) - u + seg - 1) * k + m.c] = acc;
        }
    }
}

__global__ void __launch_bounds__(kThreads)
project_combine_kernel(int64_t n, const int64_t* __restrict__ seg_incl, const int64_t* __restrict__ indptr, int k,
                       double* __restrict__ Y, const double* __restrict__ segpart) {
    const LaneMap m = lane_map(k);
    const int64_t warp = blockIdx.x * (int64_t)kWarps + (threadIdx.x >> 5);
    for (int64_t base = warp * m.rows; base < n; base += (int64_t)gridDim.x * kWarps * m.rows) {
        const int64_t i = base + m.sub;
        if (i >= n || !m.col_ok || indptr[i + 1] - indptr[i] <= kSeg) continue;
        const int64_t before = i ? seg_incl[i - 1] : 0;
        const int64_t nseg = seg_incl[i] - before;
        double s = Y[i * (int64_t)k + m.c];
        for (int64_t sg = 1; sg < nseg; ++sg) s = __dadd_rn(s, segpart[(before - i + sg - 1) * k + m.c]);
        Y[i * (int64_t)k + m.c] = s;
    }
}

// Deterministic column reduction: row slots of a warp are combined with a fixed xor tree, warps of a block in fixed
// order through shared memory, and every block writes partial[blockIdx.x][c].
__device__ __forceinline__ void block_column_partial(double lane_sum, const LaneMap& m, int k, double* partial) {
    __shared__ double sh[kWarps][kWarp];
    for (int off = kWarp / 2; off >= m.kc; off >>= 1) lane_sum = __dadd_rn(lane_sum, __shfl_xor_sync(0xffffffffu, lane_sum, off));
    const int w = threadIdx.x >> 5, l = lane_id();
    sh[w][l] = lane_sum;
    __syncthreads();
    if (w == 0 && m.sub == 0 && m.col_ok) {
        double s = sh[0][l];
#pragma unroll
        for (int i = 1; i < kWarps; ++i) s = __dadd_rn(s, sh[i][l]);
        partial[(int64_t)blockIdx.x * k + m.c] = s;
    }
    __syncthreads();   // the staging array may be reused right away
}

// r = Y (aliased), x = 0, partial column sums of b^2.
__global__ void __launch_bounds__(kThreads)
init_kernel(int64_t n, int k, const double* __restrict__ r, double* __restrict__ x, double* __restrict__ partial) {
    const LaneMap m = lane_map(k);
    double s = 0.0;
    const int64_t warp = blockIdx.x * (int64_t)kWarps + (threadIdx.x >> 5);
    for (int64_t base = warp * m.rows; base < n; base += (int64_t)gridDim.x * kWarps * m.rows) {
        const int64_t i = base + m.sub;
        if (i < n && m.col_ok) {
            const double b = r[i * (int64_t)k + m.c];
            x[i * (int64_t)k + m.c] = 0.0;
            s = __dadd_rn(s, __dmul_rn(b, b));
        }
    }
    block_column_partial(s, m, k, partial);
}

// Column finalisation at the TOP of iteration `it`: rr = sum of partials; convergence test; beta.
// Sum of one column's block partials by one warp: lane l adds blocks l, l + 32, ... in order, then a fixed xor tree —
// deterministic, and ~40 dependent adds instead of the ~1200 of a single thread (18 us per launch, twice per iteration).
__device__ __forceinline__ double column_sum(const double* __restrict__ partial, int nblocks, int k, int c) {
    double s = 0.0;
    for (int b = lane_id(); b < nblocks; b += kWarp) s = __dadd_rn(s, partial[(int64_t)b * k + c]);
    for (int off = kWarp / 2; off; off >>= 1) s = __dadd_rn(s, __shfl_xor_sync(0xffffffffu, s, off));
    return s;
}

__global__ void top_kernel(int k, int nblocks, const double* __restrict__ partial, CgColumns cg, double rtol, int it,
                           int* __restrict__ num_active) {
    const int c = blockIdx.x * (blockDim.x / kWarp) + (threadIdx.x >> 5);   // one warp per column
    if (c >= k) return;
    const double rr = column_sum(partial, nblocks, k, c);
    if (lane_id() != 0) return;
    if (it < 0) {  // initialisation: rr = b.b
        const double bnrm = sqrt(rr);
        cg.atol[c] = __dmul_rn(rtol, bnrm);                        // atol = max(0, rtol*||b||)
        cg.rho[c] = rr;
        cg.rho_prev[c] = 0.0;
        cg.iters[c] = 0;
        cg.active[c] = bnrm != 0.0;                                // ||b|| == 0 -> return b (zeros)
        if (bnrm != 0.0) atomicAdd(num_active, 1);
        return;
    }
    if (!cg.active[c]) return;
    if (it > 0) {
        cg.rho_prev[c] = cg.rho[c];
        cg.rho[c] = rr;
    }
    if (sqrt(rr) < cg.atol[c]) {                                   // "Are we done?" before the update
        cg.active[c] = 0;
        atomicSub(num_active, 1);
        return;
    }
    cg.beta[c] = it > 0 ? __ddiv_rn(rr, cg.rho_prev[c]) : 0.0;
}

// p = r + beta * p   (first iteration: p = r)
__global__ void __launch_bounds__(kThreads)
direction_kernel(int64_t n, int k, const double* __restrict__ r, double* __restrict__ p, CgColumns cg, int it) {
    const LaneMap m = lane_map(k);
    if (!m.col_ok || !cg.active[m.c]) return;
    const double beta = cg.beta[m.c];
    const int64_t warp = blockIdx.x * (int64_t)kWarps + (threadIdx.x >> 5);
    for (int64_t base = warp * m.rows; base < n; base += (int64_t)gridDim.x * kWarps * m.rows) {
        const int64_t i = base + m.sub;
        if (i >= n) continue;
        const int64_t o = i * (int64_t)k + m.c;
        p[o] = it > 0 ? __dadd_rn(__dmul_rn(p[o], beta), r[o]) : r[o];
    }
}

// q = (D - A + reg I) p, fused with the p.q column partials.
//
// Work items are row SEGMENTS of at most kSeg neighbours (a row of degree d has max(1, ceil(d/kSeg)) of them), one
// row slot of a warp per item, so a hub row with 10^5 neighbours is spread over hundreds of warps instead of
// serialising one. Rows with a single segment (almost all) are finished here in csr_matvec's exact order (neighbours
// ascending, the diagonal at its sorted position). For longer rows segment 0 parks its partial sum in q[i] and the
// others in `segpart`; spmm_combine_kernel adds them in segment order. Indices and p values are fetched four deep.
__global__ void __launch_bounds__(kThreads)
spmm_dot_kernel(const SegItem* __restrict__ items, int64_t num_items, const int64_t* __restrict__ seg_incl,
                const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, const double* __restrict__ data,
                const double* __restrict__ diag, int k, const double* __restrict__ p, double* __restrict__ q,
                double* __restrict__ segpart, const int* __restrict__ active, double* __restrict__ partial) {
    const LaneMap m = lane_map(k);
    const bool col_ok = m.col_ok && active[m.c];
    double dot = 0.0;
    if (__any_sync(0xffffffffu, col_ok)) {  // a warp whose columns have all stopped does no work
        const int64_t warp = blockIdx.x * (int64_t)kWarps + (threadIdx.x >> 5);
        for (int64_t base = warp * m.rows; base < num_items; base += (int64_t)gridDim.x * kWarps * m.rows) {
            const int64_t it = base + m.sub;
            const bool item_ok = it < num_items;
            const int32_t i = item_ok ? items[it].row : 0;
            const int seg = item_ok ? items[it].seg : 0;
            const int64_t row0 = indptr[i], row1 = indptr[i + 1];
            const int64_t p0 = item_ok ? row0 + (int64_t)seg * kSeg : 0;
            const int64_t p1 = item_ok ? min(p0 + kSeg, row1) : 0;
            const bool single = row1 - row0 <= kSeg;
            const bool lane_ok = col_ok && item_ok;
            const double pi = (lane_ok && single) ? p[i * (int64_t)k + m.c] : 0.0;
            const double dterm = __dmul_rn(diag[i], pi);
            double s = 0.0;
            bool placed = !single;                  // multi-segment rows: the diagonal is added by the combine kernel
            int32_t jn[4];                          // ids of the next four-deep step, fetched one step ahead
#pragma unroll
            for (int u = 0; u < 4; ++u) jn[u] = p0 + u < p1 ? __ldg(indices + p0 + u) : -1;
            for (int64_t t0 = p0; __any_sync(0xffffffffu, t0 < p1); t0 += 4) {   // row slots may differ in length
                int32_t j[4];
                double pj[4], a[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    j[u] = jn[u];
                    pj[u] = (lane_ok && j[u] >= 0) ? __ldg(p + (int64_t)j[u] * k + m.c) : 0.0;
                    a[u] = (data && j[u] >= 0) ? __ldg(data + t0 + u) : 1.0;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) jn[u] = t0 + 4 + u < p1 ? __ldg(indices + t0 + 4 + u) : -1;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (j[u] < 0) break;
                    if (j[u] == i) {                // self loop: folded into diag (L_ii = deg - a_ii + reg)
                        if (!placed) { s = __dadd_rn(s, dterm); placed = true; }
                        continue;
                    }
                    if (!placed && j[u] > i) { s = __dadd_rn(s, dterm); placed = true; }
                    s = data ? __dadd_rn(s, __dmul_rn(-a[u], pj[u])) : __dsub_rn(s, pj[u]);
                }
            }
            if (!placed) s = __dadd_rn(s, dterm);
            if (lane_ok) {
                if (single) {
                    q[i * (int64_t)k + m.c] = s;
                    dot = __dadd_rn(dot, __dmul_rn(pi, s));
                } else if (seg == 0) {
                    q[i * (int64_t)k + m.c] = s;
                } else {   // extra segments of all rows are numbered consecutively: (items before row i) - i + seg - 1
                    const int64_t slot = (i ? seg_incl[i - 1] : 0) - i + seg - 1;
                    segpart[slot * k + m.c] = s;
                }
            }
        }
    }
    block_column_partial(dot, m, k, partial);
}

// Two adjacent columns per lane (k % 64 == 0): one warp covers a 64-column strip with 16-byte gathers, so the index
// list of a row is read once per 64 columns instead of once per 32 and half as many load instructions are issued.
// The neighbour ids of the next four-deep step are fetched while the current step's gathers are in flight (the first
// version serialised "load ids -> gather p -> consume" and ran at 30 % occupancy / 12 % of the L2->SM path, ncu).
template <bool kHasData, bool kPacked, int kBlocks, int kDepth>
__global__ void __launch_bounds__(kThreads, kBlocks)
spmm_dot2_kernel(const SegItem* __restrict__ items, int64_t num_items, const int64_t* __restrict__ seg_incl,
                 const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, const double* __restrict__ data,
                 const double* __restrict__ diag, int k, const double* __restrict__ p, double* __restrict__ q,
                 double* __restrict__ segpart, const int* __restrict__ active, double* __restrict__ partial) {
    // k >= 64: one row per warp, 64-column strips. k in {2,4,...,32}: k/2 lanes per row, 64/k rows per warp.
    const int lanes_per_row = kPacked ? k / 2 : kWarp;
    const int rows_per_warp = kWarp / lanes_per_row;
    const int sub = lane_id() / lanes_per_row;
    const int c = blockIdx.y * 64 + 2 * (lane_id() % lanes_per_row);
    const bool any_active = active[c] | active[c + 1];
    double dot0 = 0.0, dot1 = 0.0;
    if (__any_sync(0xffffffffu, any_active)) {
        const int64_t warp = blockIdx.x * (int64_t)kWarps + (threadIdx.x >> 5);
        for (int64_t base = warp * rows_per_warp; base < num_items; base += (int64_t)gridDim.x * kWarps * rows_per_warp) {
            const int64_t it = base + sub;
            const bool item_ok = it < num_items;
            const int32_t i = item_ok ? items[it].row : 0;
            const int seg = item_ok ? items[it].seg : 0;
            const int64_t row0 = indptr[i], row1 = indptr[i + 1];
            const int64_t p0 = item_ok ? row0 + (int64_t)seg * kSeg : 0;
            const int64_t p1 = item_ok ? min(p0 + kSeg, row1) : 0;
            const bool single = row1 - row0 <= kSeg;
            int32_t jn[kDepth];
#pragma unroll
            for (int u = 0; u < kDepth; ++u) jn[u] = p0 + u < p1 ? __ldg(indices + p0 + u) : -1;
            double2 pi = make_double2(0.0, 0.0);
            if (single && item_ok) pi = *reinterpret_cast<const double2*>(p + i * (int64_t)k + c);
            const double di = diag[i];
            const double d0 = __dmul_rn(di, pi.x), d1 = __dmul_rn(di, pi.y);
            double s0 = 0.0, s1 = 0.0;
            bool placed = !single;
            for (int64_t t0 = p0; kPacked ? __any_sync(0xffffffffu, t0 < p1) : t0 < p1; t0 += kDepth) {   // packed row slots differ in length
                int32_t j[kDepth];
                double2 pj[kDepth];
                double a[kDepth];
#pragma unroll
                for (int u = 0; u < kDepth; ++u) {
                    j[u] = jn[u];
                    pj[u] = j[u] >= 0 ? __ldg(reinterpret_cast<const double2*>(p + (int64_t)j[u] * k + c)) : make_double2(0.0, 0.0);
                    a[u] = (kHasData && j[u] >= 0) ? __ldg(data + t0 + u) : 1.0;
                }
#pragma unroll
                for (int u = 0; u < kDepth; ++u) jn[u] = t0 + kDepth + u < p1 ? __ldg(indices + t0 + kDepth + u) : -1;   // next step's ids
#pragma unroll
                for (int u = 0; u < kDepth; ++u) {
                    if (j[u] < 0) break;
                    if (j[u] == i) {
                        if (!placed) { s0 = __dadd_rn(s0, d0); s1 = __dadd_rn(s1, d1); placed = true; }
                        continue;
                    }
                    if (!placed && j[u] > i) { s0 = __dadd_rn(s0, d0); s1 = __dadd_rn(s1, d1); placed = true; }
                    if (kHasData) {
                        s0 = __dadd_rn(s0, __dmul_rn(-a[u], pj[u].x));
                        s1 = __dadd_rn(s1, __dmul_rn(-a[u], pj[u].y));
                    } else {
                        s0 = __dsub_rn(s0, pj[u].x);
                        s1 = __dsub_rn(s1, pj[u].y);
                    }
                }
            }
            if (!placed) { s0 = __dadd_rn(s0, d0); s1 = __dadd_rn(s1, d1); }
            if (item_ok) {
                if (single || seg == 0) {
                    *reinterpret_cast<double2*>(q + i * (int64_t)k + c) = make_double2(s0, s1);
                    if (single) {
                        dot0 = __dadd_rn(dot0, __dmul_rn(pi.x, s0));
                        dot1 = __dadd_rn(dot1, __dmul_rn(pi.y, s1));
                    }
                } else {
                    const int64_t slot = (i ? seg_incl[i - 1] : 0) - i + seg - 1;
                    *reinterpret_cast<double2*>(segpart + slot * k + c) = make_double2(s0, s1);
                }
            }
        }
    }
    LaneMap m{c, sub, rows_per_warp, lanes_per_row, true};
    block_column_partial(dot0, m, k, partial);
    m.c = c + 1;
    block_column_partial(dot1, m, k, partial);
}

// Rows longer than kSeg: q[i] = seg0 + diag*p_i + seg1 + seg2 + ...   and their share of p.q
__global__ void __launch_bounds__(kThreads)
spmm_combine_kernel(int64_t n, const int64_t* __restrict__ seg_incl, const int64_t* __restrict__ indptr,
                    const double* __restrict__ diag, int k, const double* __restrict__ p, double* __restrict__ q,
                    const double* __restrict__ segpart, const int* __restrict__ active, double* __restrict__ partial) {
    const LaneMap m = lane_map(k);
    const bool col_ok = m.col_ok && active[m.c];
    double dot = 0.0;
    const int64_t warp = blockIdx.x * (int64_t)kWarps + (threadIdx.x >> 5);
    for (int64_t base = warp * m.rows; base < n; base += (int64_t)gridDim.x * kWarps * m.rows) {
        const int64_t i = base + m.sub;
        if (i >= n || !col_ok || indptr[i + 1] - indptr[i] <= kSeg) continue;
        const int64_t before = i ? seg_incl[i - 1] : 0;
        const int64_t nseg = seg_incl[i] - before;
        const double pi = p[i * (int64_t)k + m.c];
        double s = __dadd_rn(q[i * (int64_t)k + m.c], __dmul_rn(diag[i], pi));
        const int64_t slot0 = before - i;
        for (int64_t sg = 1; sg < nseg; ++sg) s = __dadd_rn(s, segpart[(slot0 + sg - 1) * k + m.c]);
        q[i * (int64_t)k + m.c] = s;
        dot = __dadd_rn(dot, __dmul_rn(pi, s));
    }
    block_column_partial(dot, m, k, partial);
}

__global__ void seg_count_kernel(int64_t n, const int64_t* __restrict__ indptr, int64_t* __restrict__ counts) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t d = indptr[i + 1] - indptr[i];
        counts[i] = d <= kSeg ? 1 : (d + kSeg - 1) / kSeg;
    }
}

__global__ void seg_fill_kernel(int64_t n, const int64_t* __restrict__ incl, SegItem* __restrict__ items) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t before = i ? incl[i - 1] : 0;
        const int64_t c = incl[i] - before;
        for (int64_t sg = 0; sg < c; ++sg) items[before + sg] = SegItem{(int32_t)i, (int32_t)sg};
    }
}

__global__ void alpha_kernel(int k, int nblocks, const double* __restrict__ partial, CgColumns cg) {
    const int c = blockIdx.x * (blockDim.x / kWarp) + (threadIdx.x >> 5);   // one warp per column
    if (c >= k || !cg.active[c]) return;
    const double pq = column_sum(partial, nblocks, k, c);
    if (lane_id() != 0) return;
    cg.alpha[c] = __ddiv_rn(cg.rho[c], pq);
    cg.iters[c] += 1;
}

// x += alpha p ; r -= alpha q ; partial column sums of the new r.r
__global__ void __launch_bounds__(kThreads)
update_kernel(int64_t n, int k, const double* __restrict__ p, const double* __restrict__ q, double* __restrict__ x,
              double* __restrict__ r, CgColumns cg, double* __restrict__ partial) {
    const LaneMap m = lane_map(k);
    const bool col_ok = m.col_ok && cg.active[m.c];
    double s = 0.0;
    if (col_ok) {
        const double alpha = cg.alpha[m.c];
        const int64_t warp = blockIdx.x * (int64_t)kWarps + (threadIdx.x >> 5);
        for (int64_t base = warp * m.rows; base < n; base += (int64_t)gridDim.x * kWarps * m.rows) {
            const int64_t i = base + m.sub;
            if (i >= n) continue;
            const int64_t o = i * (int64_t)k + m.c;
            x[o] = __dadd_rn(x[o], __dmul_rn(alpha, p[o]));
            const double rn = __dsub_rn(r[o], __dmul_rn(alpha, q[o]));
            r[o] = rn;
            s = __dadd_rn(s, __dmul_rn(rn, rn));
        }
    }
    block_column_partial(s, m, k, partial);
}

__global__ void diag_kernel(int64_t n, const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                            const double* __restrict__ data, double reg, double* __restrict__ diag) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double deg = 0.0, self = 0.0;
        for (int64_t t = indptr[i]; t < indptr[i + 1]; ++t) {
            const double a = data ? data[t] : 1.0;
            deg = __dadd_rn(deg, a);
            if (indices[t] == i) self = a;
        }
        diag[i] = __dadd_rn(__dsub_rn(deg, self), reg);  // (D - A)_ii + 1e-6  (metrics.py:251-256)
    }
}

// partial[e] = sum_c (Z[u,c] - Z[v,c])^2 ; non-finite solver output counts as 0 (metrics.py:287-293)
__global__ void __launch_bounds__(kThreads)
resistance_kernel(int64_t e_begin, int64_t e_end, const int32_t* __restrict__ rows, const int32_t* __restrict__ indices,
                  const double* __restrict__ z, int k, double* __restrict__ out) {
    // lanes of a slot cover the columns (stride kc); small k packs 32/kc edges into one warp
    int kc = kWarp;
    if (k < kWarp) { kc = 1; while (kc < k) kc <<= 1; }
    const int lane = lane_id(), c0 = lane & (kc - 1), sub = lane / kc, per = kWarp / kc;
    const int64_t warp = blockIdx.x * (int64_t)kWarps + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kWarps;
    for (int64_t base = e_begin + warp * per; base < e_end; base += nwarps * per) {
        const int64_t e = base + sub;
        double s = 0.0;
        if (e < e_end) {
            const double* zu = z + (int64_t)__ldg(rows + e) * k;
            const double* zv = z + (int64_t)__ldg(indices + e) * k;
            for (int c = c0; c < k; c += kc) {
                double a = __ldg(zu + c), b = __ldg(zv + c);
                if (!isfinite(a)) a = 0.0;
                if (!isfinite(b)) b = 0.0;
                const double d = __dsub_rn(a, b);
                s = __dadd_rn(s, __dmul_rn(d, d));
            }
        }
        for (int o = kc / 2; o; o >>= 1) s = __dadd_rn(s, __shfl_xor_sync(0xffffffffu, s, o));
        if (e < e_end && c0 == 0) out[e - e_begin] = s;
    }
}

__global__ void er_finalize_kernel(int64_t n, double* s) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double v = s[i];
        if (!isfinite(v)) v = 1e-10;          // nan_to_num(nan=1e-10, posinf=1e-10, neginf=1e-10)
        s[i] = v > 1e-10 ? v : 1e-10;         // np.maximum(r_eff, 1e-10)
    }
}

std::mutex g_und_mutex;

int ensure_segments(Graph* g, cudaStream_t s) {
    std::lock_guard<std::mutex> lock(g_und_mutex);
    if (g->seg_items || g->n == 0) return GSP_OK;
    Scratch<int64_t> counts;
    GSP_CUDA_TRY(counts.alloc(g->n, s));
    int64_t* incl = nullptr;
    GSP_CUDA_TRY(device_alloc(&incl, (size_t)g->n * sizeof(int64_t), s));
    seg_count_kernel<<<grid_for(g->n, 256), 256, 0, s>>>(g->n, g->indptr, counts.ptr);
    GSP_CHECK_LAUNCH();
    if (int rc = inclusive_sum_i64(counts.ptr, incl, g->n, s)) { device_free(incl); return rc; }
    int64_t total = 0;
    GSP_CUDA_TRY(cudaMemcpyAsync(&total, incl + (g->n - 1), sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    GSP_CUDA_TRY(cudaStreamSynchronize(s));
    SegItem* items = nullptr;
    GSP_CUDA_TRY(device_alloc(&items, (size_t)total * sizeof(SegItem), s));
    seg_fill_kernel<<<grid_for(g->n, 256), 256, 0, s>>>(g->n, incl, items);
    GSP_CHECK_LAUNCH();
    GSP_CUDA_TRY(cudaStreamSynchronize(s));
    g->seg_items = items;
    g->seg_incl = incl;
    g->num_seg_items = total;
    return GSP_OK;
}

int ensure_und_id(Graph* g, void* stream) {
    std::lock_guard<std::mutex> lock(g_und_mutex);
    if (g->und_id || g->nnz == 0) return GSP_OK;
    int32_t* buf = nullptr;
    GSP_CUDA_TRY(device_alloc(&buf, (size_t)g->nnz * sizeof(int32_t), as_stream(stream)));
    int rc = gsp_graph_undirected_ids(reinterpret_cast<gsp_graph*>(g), buf, stream);
    if (rc) {
        device_free(buf);
        return rc;
    }
    g->und_id = buf;
    return GSP_OK;
}

}  // namespace
}  // namespace gsp

using namespace gsp;

// `d_rhs` / `d_solution` (both [n, k] row-major, optional): solve for an explicit right-hand side instead of the projected
// incidence columns and hand the solution out instead of the per-edge resistance sums (gsp_laplacian_solve).
static int approx_er_partial(const gsp_graph* gg, const Projection& pr, int32_t k, int32_t max_iters, double rtol, double reg,
                             int64_t e_begin, int64_t e_end, double* d_partial, int32_t* d_iters, void* stream,
                             const double* d_rhs = nullptr, double* d_solution = nullptr) {
    GSP_REQUIRE(gg != nullptr, "graph is NULL");
    Graph* g = const_cast<Graph*>(reinterpret_cast<const Graph*>(gg));
    GSP_REQUIRE(e_begin >= 0 && e_begin <= e_end && e_end <= g->nnz, "edge range outside [0, nnz]");
    GSP_REQUIRE(k >= 1 && max_iters >= 0, "bad k / max_iters");
    GSP_REQUIRE(e_end == e_begin || d_partial || d_solution, "NULL argument");
    if (!g->symmetric) {
        set_error("approximate effective resistance needs a symmetric adjacency pattern (reference metrics.py:208-209)");
        return GSP_ERR_UNSUPPORTED;
    }
    cudaStream_t s = as_stream(stream);
    const int64_t n = g->n;
    if (!d_rhs) {
        if (int rc = ensure_und_id(g, stream)) return rc;
    }
    if (int rc = ensure_segments(g, s)) return rc;
    const int64_t extra_segments = g->num_seg_items - n;   // > 0 when some row is longer than kSeg

    // Four-deep gathers at 4 CTAs / SM. Measured on the products shape (k = 64): 9.8-10.2 ms per CG iteration; eight-deep
    // gathers at 3 CTAs / SM 11.2 ms, four-deep at 5 CTAs / SM (48 registers, spills) 11.0 ms — more loads in flight do not
    // help: the gather moves nnz * 8k = 63 GB from L2 to the SMs per product (ncu: 25 GB of it from DRAM, L2 hit rate 48 %),
    // ~6.5 TB/s through the L2 fabric.
    const int spmm_ctas = 8;
    const int strips = k >= kWarp ? (k + kWarp - 1) / kWarp : 1;   // k < 32: several rows per warp instead of strips
    // row blocks: enough CTAs for >= 8 per SM over all strips, at most one warp-row each
    int64_t row_blocks = (static_cast<int64_t>(kNumSMs) * spmm_ctas + strips - 1) / strips;
    const int64_t max_rb = (n + kWarps - 1) / kWarps;
    if (row_blocks > max_rb) row_blocks = max_rb;
    if (row_blocks < 1) row_blocks = 1;
    const dim3 grid2d((unsigned)row_blocks, (unsigned)strips);
    const int nb = (int)row_blocks;

    const size_t vec = (size_t)n * (size_t)k;
    Scratch<double> x, r, p, q, partial, diag, cols;
    Scratch<int> flags;
    Scratch<int32_t> iters_local;
    GSP_CUDA_TRY(x.alloc(vec, s));
    GSP_CUDA_TRY(r.alloc(vec, s));
    GSP_CUDA_TRY(p.alloc(vec, s));
    GSP_CUDA_TRY(q.alloc(vec, s));
    GSP_CUDA_TRY(partial.alloc((size_t)2 * nb * k, s));   // second half: the combine kernel's p.q partials
    Scratch<double> segpart;
    GSP_CUDA_TRY(segpart.alloc((size_t)(extra_segments > 0 ? extra_segments : 1) * k, s));
    GSP_CUDA_TRY(diag.alloc(n, s));
    GSP_CUDA_TRY(cols.alloc((size_t)5 * k, s));
    GSP_CUDA_TRY(flags.alloc((size_t)k + 1, s));
    GSP_CUDA_TRY(iters_local.alloc(k, s));
    CgColumns cg{cols.ptr, cols.ptr + k, cols.ptr + 2 * (size_t)k, cols.ptr + 3 * (size_t)k, cols.ptr + 4 * (size_t)k,
                 flags.ptr, d_iters ? d_iters : iters_local.ptr};
    int* num_active = flags.ptr + k;
    GSP_CUDA_TRY(cudaMemsetAsync(num_active, 0, sizeof(int), s));

    diag_kernel<<<grid_for(n, 256), 256, 0, s>>>(n, g->indptr, g->indices, g->data, reg, diag.ptr);
    GSP_CHECK_LAUNCH();
    const SegItem* seg_items = reinterpret_cast<const SegItem*>(g->seg_items);
    if (d_rhs) {
        GSP_CUDA_TRY(cudaMemcpyAsync(r.ptr, d_rhs, vec * sizeof(double), cudaMemcpyDeviceToDevice, s));
    } else {
        project_kernel<<<grid2d, kThreads, 0, s>>>(seg_items, g->num_seg_items, g->seg_incl, g->indptr, g->indices, g->und_id, pr,
                                                   k, r.ptr, segpart.ptr);
        GSP_CHECK_LAUNCH();
        if (extra_segments > 0) {
            project_combine_kernel<<<grid2d, kThreads, 0, s>>>(n, g->seg_incl, g->indptr, k, r.ptr, segpart.ptr);
            GSP_CHECK_LAUNCH();
        }
    }
    init_kernel<<<grid2d, kThreads, 0, s>>>(n, k, r.ptr, x.ptr, partial.ptr);
    GSP_CHECK_LAUNCH();
    const int col_blocks = (k + 3) / 4;   // 128 threads = four columns (one warp each)
    top_kernel<<<col_blocks, 128, 0, s>>>(k, nb, partial.ptr, cg, rtol, -1, num_active);
    GSP_CHECK_LAUNCH();

    int host_active = 1;
    for (int it = 0; it < max_iters; ++it) {
        if (it % kCheckEvery == 0) {  // the only host round trip of the solve
            GSP_CUDA_TRY(cudaMemcpyAsync(&host_active, num_active, sizeof(int), cudaMemcpyDeviceToHost, s));
            GSP_CUDA_TRY(cudaStreamSynchronize(s));
            if (host_active <= 0) break;
        }
        top_kernel<<<col_blocks, 128, 0, s>>>(k, nb, partial.ptr, cg, rtol, it, num_active);
        GSP_CHECK_LAUNCH();
        direction_kernel<<<grid2d, kThreads, 0, s>>>(n, k, r.ptr, p.ptr, cg, it);
        GSP_CHECK_LAUNCH();
        const bool pow2_small = k >= 2 && k < 64 && (k & (k - 1)) == 0;
        if (k % 64 == 0 || pow2_small) {   // two columns per lane: 64-column strips, or 64/k rows per warp when k < 64
            const dim3 grid64((unsigned)row_blocks, (unsigned)(k >= 64 ? k / 64 : 1));
#define GSP_SPMM2(HAS_DATA, PACKED)                                                                                          \
    spmm_dot2_kernel<HAS_DATA, PACKED, 4, 4><<<grid64, kThreads, 0, s>>>(seg_items, g->num_seg_items, g->seg_incl, g->indptr, g->indices, \
                                                                  g->data, diag.ptr, k, p.ptr, q.ptr, segpart.ptr, cg.active,    \
                                                                  partial.ptr)
            if (g->data) { if (pow2_small) GSP_SPMM2(true, true); else GSP_SPMM2(true, false); }
            else { if (pow2_small) GSP_SPMM2(false, true); else GSP_SPMM2(false, false); }
#undef GSP_SPMM2
        } else {
            spmm_dot_kernel<<<grid2d, kThreads, 0, s>>>(seg_items, g->num_seg_items, g->seg_incl, g->indptr, g->indices, g->data,
                                                       diag.ptr, k, p.ptr, q.ptr, segpart.ptr, cg.active, partial.ptr);
        }
        GSP_CHECK_LAUNCH();
        if (extra_segments > 0) {
            spmm_combine_kernel<<<grid2d, kThreads, 0, s>>>(n, g->seg_incl, g->indptr, diag.ptr, k, p.ptr, q.ptr, segpart.ptr,
                                                           cg.active, partial.ptr + (size_t)nb * k);
            GSP_CHECK_LAUNCH();
        }
        alpha_kernel<<<col_blocks, 128, 0, s>>>(k, extra_segments > 0 ? 2 * nb : nb, partial.ptr, cg);
        GSP_CHECK_LAUNCH();
        update_kernel<<<grid2d, kThreads, 0, s>>>(n, k, p.ptr, q.ptr, x.ptr, r.ptr, cg, partial.ptr);
        GSP_CHECK_LAUNCH();
    }
    if (d_solution) {
        GSP_CUDA_TRY(cudaMemcpyAsync(d_solution, x.ptr, vec * sizeof(double), cudaMemcpyDeviceToDevice, s));
    } else if (e_end > e_begin) {
        const int64_t edges = e_end - e_begin;
        resistance_kernel<<<grid_for(edges, kWarps, 16), kThreads, 0, s>>>(e_begin, e_end, g->rows, g->indices, x.ptr, k,
                                                                         d_partial);
        GSP_CHECK_LAUNCH();
    }
    return GSP_OK;
}

GSP_API int gsp_approx_er_partial(const gsp_graph* gg, const double* d_R, int64_t ldr, int32_t k, int32_t max_iters,
                                  double rtol, double reg, int64_t e_begin, int64_t e_end, double* d_partial,
                                  int32_t* d_iters, void* stream) {
    GSP_REQUIRE(d_R != nullptr && ldr >= k, "projection matrix is NULL or ldr < k");
    const Projection pr{d_R, ldr, 0, 0, 1.0};
    return approx_er_partial(gg, pr, k, max_iters, rtol, reg, e_begin, e_end, d_partial, d_iters, stream);
}

GSP_API int gsp_approx_er_partial_philox(const gsp_graph* gg, uint64_t seed, int32_t col_begin, int32_t k, int32_t k_total,
                                         int32_t max_iters, double rtol, double reg, int64_t e_begin, int64_t e_end,
                                         double* d_partial, int32_t* d_iters, void* stream) {
    GSP_REQUIRE(col_begin >= 0 && k >= 1 && k_total >= col_begin + k, "bad column range");
    const Projection pr{nullptr, 0, seed, col_begin, 1.0 / sqrt((double)k_total)};
    return approx_er_partial(gg, pr, k, max_iters, rtol, reg, e_begin, e_end, d_partial, d_iters, stream);
}

GSP_API int gsp_laplacian_solve(const gsp_graph* gg, const double* d_rhs, int32_t k, int32_t max_iters, double rtol, double reg,
                                double* d_x, int32_t* d_iters, void* stream) {
    GSP_REQUIRE(d_rhs != nullptr && d_x != nullptr, "NULL argument");
    const Projection none{nullptr, 0, 0, 0, 1.0};
    return approx_er_partial(gg, none, k, max_iters, rtol, reg, 0, 0, nullptr, d_iters, stream, d_rhs, d_x);
}

GSP_API int gsp_philox_projection(uint64_t seed, int64_t m, int32_t col_begin, int32_t k, int32_t k_total, double* d_R,
                                  void* stream) {
    GSP_REQUIRE(m >= 0 && col_begin >= 0 && k >= 1 && k_total >= col_begin + k, "bad sizes");
    if (m == 0) return GSP_OK;
    GSP_REQUIRE(d_R != nullptr, "d_R is NULL");
    const Projection pr{nullptr, 0, seed, col_begin, 1.0 / sqrt((double)k_total)};
    philox_projection_kernel<<<grid_for(m * k, 256), 256, 0, as_stream(stream)>>>(m, k, pr, d_R);
    GSP_CHECK_LAUNCH();
    return GSP_OK;
}

GSP_API int gsp_er_finalize(double* d_score, int64_t count, void* stream) {
    GSP_REQUIRE(count >= 0 && (count == 0 || d_score), "bad arguments");
    if (count == 0) return GSP_OK;
    er_finalize_kernel<<<grid_for(count, 256), 256, 0, as_stream(stream)>>>(count, d_score);
    GSP_CHECK_LAUNCH();
    return GSP_OK;
}
