/*
 * gsp_oracle.c — TEST INFRASTRUCTURE ONLY (never linked into or called by the product).
 *
 * Plain-C, single-threaded restatement of the *per-edge semantics* of the reference's
 * sparsification engine (/root/reference/src/sparsification/{core,metrics}.py). Where the
 * reference gets a number out of SciPy SpGEMM / NumPy reductions, this file states the
 * arithmetic those libraries perform (operation order included) as explicit loops, so the
 * CUDA kernels can be checked bit-for-bit at sizes where the reference itself is too slow
 * (its SpGEMM materialises all 2-hop pairs).
 *
 * Pinning: tests/test_oracle_pin.py compares every function here with the live reference
 * (build container, numpy 2.3.5 / scipy 1.18.1) and with the committed fixtures in
 * tests/golden/ (generated from the live reference by oracle/make_golden.py).
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off: no FMA contraction, no reassociation)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define GSPO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------
 * Canonical CSR.  reference core.py:70-74 — sp.csr_matrix((ones(E), (row, col))):
 * rows ascending, columns ascending inside a row, duplicate (row, col) pairs summed so
 * data[] holds the multiplicity.  `val` may be NULL (all ones).
 * Returns nnz, or -1 when an index is outside [0, n).
 * ---------------------------------------------------------------------------------- */
typedef struct { int32_t c; int64_t o; double v; } gspo_cov;   /* column, input offset, value */

static int cmp_cov(const void* a, const void* b) {
    const gspo_cov* p = (const gspo_cov*)a; const gspo_cov* q = (const gspo_cov*)b;
    if (p->c != q->c) return (p->c > q->c) - (p->c < q->c);
    return (p->o > q->o) - (p->o < q->o);           /* ties in input order: a stable sort */
}

GSPO_API int64_t gspo_csr_build(int64_t n, int64_t E, const int64_t* row, const int64_t* col, const double* val,
                                int64_t* indptr, int32_t* indices, double* data) {
    int64_t* start = (int64_t*)calloc((size_t)n + 1, sizeof(int64_t));
    for (int64_t e = 0; e < E; ++e) {
        if (row[e] < 0 || row[e] >= n || col[e] < 0 || col[e] >= n) { free(start); return -1; }
        start[row[e] + 1]++;
    }
    for (int64_t i = 0; i < n; ++i) start[i + 1] += start[i];
    gspo_cov* tmp = (gspo_cov*)malloc((size_t)(E > 0 ? E : 1) * sizeof(gspo_cov));
    int64_t* cur = (int64_t*)malloc(((size_t)n + 1) * sizeof(int64_t));
    memcpy(cur, start, ((size_t)n + 1) * sizeof(int64_t));
    for (int64_t e = 0; e < E; ++e) {               /* bucket by row, input order kept */
        int64_t p = cur[row[e]]++;
        tmp[p].c = (int32_t)col[e]; tmp[p].o = e; tmp[p].v = val ? val[e] : 1.0;
    }
    int64_t nnz = 0;
    indptr[0] = 0;
    for (int64_t i = 0; i < n; ++i) {
        int64_t lo = start[i], hi = start[i + 1];
        if (hi - lo > 1) qsort(tmp + lo, (size_t)(hi - lo), sizeof(gspo_cov), cmp_cov);
        for (int64_t a = lo; a < hi; ++a) {         /* duplicates summed in input order */
            if (nnz > indptr[i] && indices[nnz - 1] == tmp[a].c) data[nnz - 1] += tmp[a].v;
            else { indices[nnz] = tmp[a].c; data[nnz] = tmp[a].v; ++nnz; }
        }
        indptr[i + 1] = nnz;
    }
    free(tmp); free(cur); free(start);
    return nnz;
}

/* Pattern transpose (CSC of the same matrix viewed as CSR of A^T); columns come out sorted. */
GSPO_API void gspo_transpose(int64_t n, const int64_t* indptr, const int32_t* indices,
                             int64_t* tptr, int32_t* tidx) {
    int64_t nnz = indptr[n];
    memset(tptr, 0, ((size_t)n + 1) * sizeof(int64_t));
    for (int64_t p = 0; p < nnz; ++p) tptr[indices[p] + 1]++;
    for (int64_t i = 0; i < n; ++i) tptr[i + 1] += tptr[i];
    int64_t* cur = (int64_t*)malloc(((size_t)n + 1) * sizeof(int64_t));
    memcpy(cur, tptr, ((size_t)n + 1) * sizeof(int64_t));
    for (int64_t i = 0; i < n; ++i)
        for (int64_t p = indptr[i]; p < indptr[i + 1]; ++p) tidx[cur[indices[p]]++] = (int32_t)i;
    free(cur);
}

/* ------------------------------------------------------------------------------------
 * Jaccard.  reference metrics.py:43-62, SURVEY App. A.1.
 *   I = (Ab @ Ab)[u, v] = |row_A(u) ∩ col_A(v)|   (col_A(v) = row_{A^T}(v); pass tptr/tidx = NULL
 *       for a symmetric pattern, where it equals row_A(v))
 *   U = deg[u] + deg[v] - I  with deg = row sums of the binarised matrix (both are ROW degrees)
 *   score = I / U when U > 0 else 0      (one IEEE fp64 divide)
 * ---------------------------------------------------------------------------------- */
GSPO_API void gspo_jaccard(int64_t n, const int64_t* indptr, const int32_t* indices,
                           const int64_t* tptr, const int32_t* tidx, int32_t* inter_out, double* score_out) {
    if (!tptr) { tptr = indptr; tidx = indices; }
    for (int64_t u = 0; u < n; ++u) {
        for (int64_t p = indptr[u]; p < indptr[u + 1]; ++p) {
            int64_t v = indices[p];
            int64_t a = indptr[u], ae = indptr[u + 1], b = tptr[v], be = tptr[v + 1];
            int32_t cnt = 0;
            while (a < ae && b < be) {
                int32_t x = indices[a], y = tidx[b];
                if (x == y) { ++cnt; ++a; ++b; } else if (x < y) ++a; else ++b;
            }
            double du = (double)(indptr[u + 1] - indptr[u]), dv = (double)(indptr[v + 1] - indptr[v]);
            double uni = du + dv - (double)cnt;
            if (inter_out) inter_out[p] = cnt;
            score_out[p] = uni > 0.0 ? (double)cnt / uni : 0.0;
        }
    }
}

/* ------------------------------------------------------------------------------------
 * Adamic-Adar.  reference metrics.py:99-119, SURVEY App. A.2.
 *   score[u, v] = (W W^T)[u, v] = sum over x in row(u) ∩ row(v) of w[x] * w[x]
 * SciPy's SpGEMM accumulates the terms in DESCENDING x (artefact of csr_matmat's linked
 * list + the unsorted product W = Ab @ diag(w)); each term is a rounded multiply followed
 * by a rounded add into an accumulator that starts at 0.0.  w[] is supplied by the caller
 * (NumPy expression of metrics.py:104-108) because its bits are libm-defined.
 * ---------------------------------------------------------------------------------- */
GSPO_API void gspo_adamic_adar(int64_t n, const int64_t* indptr, const int32_t* indices, const double* w,
                               double* out) {
    for (int64_t u = 0; u < n; ++u) {
        for (int64_t p = indptr[u]; p < indptr[u + 1]; ++p) {
            int64_t v = indices[p];
            int64_t a = indptr[u + 1] - 1, a0 = indptr[u], b = indptr[v + 1] - 1, b0 = indptr[v];
            double acc = 0.0;
            while (a >= a0 && b >= b0) {
                int32_t x = indices[a], y = indices[b];
                if (x == y) { double t = w[x] * w[x]; acc = acc + t; --a; --b; }
                else if (x > y) --a; else --b;
            }
            out[p] = acc;
        }
    }
}

/* reference core.py:167-172 — degrees[rows] * degrees[cols], degrees = row sums of raw data */
GSPO_API void gspo_degree_product(int64_t n, const int64_t* indptr, const int32_t* indices, const double* data,
                                  double* out) {
    double* deg = (double*)calloc((size_t)n + 1, sizeof(double));
    for (int64_t u = 0; u < n; ++u) {
        double s = 0.0;
        for (int64_t p = indptr[u]; p < indptr[u + 1]; ++p) s += data ? data[p] : 1.0;
        deg[u] = s;
    }
    for (int64_t u = 0; u < n; ++u)
        for (int64_t p = indptr[u]; p < indptr[u + 1]; ++p) out[p] = deg[u] * deg[indices[p]];
    free(deg);
}

/* ------------------------------------------------------------------------------------
 * NumPy pairwise summation (numpy/_core/src/umath/loops_utils.h.src, pairwise_sum_@TYPE@),
 * the tree behind np.sum(axis=1) / np.linalg.norm(axis=1) on a C-contiguous [rows, d] array
 * (reference metrics.py:344,351).  SURVEY App. A.3.
 * ---------------------------------------------------------------------------------- */
#define DEFINE_PAIRWISE(NAME, T)                                                         \
    static T NAME(const T* a, int64_t n) {                                               \
        if (n < 8) {                                                                     \
            T res = (T)0;                                                                \
            for (int64_t i = 0; i < n; ++i) res = res + a[i];                            \
            return res;                                                                  \
        } else if (n <= 128) {                                                           \
            T r[8];                                                                      \
            for (int j = 0; j < 8; ++j) r[j] = a[j];                                     \
            int64_t i;                                                                   \
            for (i = 8; i < n - (n % 8); i += 8)                                         \
                for (int j = 0; j < 8; ++j) r[j] = r[j] + a[i + j];                      \
            T res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));   \
            for (; i < n; ++i) res = res + a[i];                                         \
            return res;                                                                  \
        } else {                                                                         \
            int64_t n2 = n / 2;                                                          \
            n2 -= n2 % 8;                                                                \
            return NAME(a, n2) + NAME(a + n2, n - n2);                                   \
        }                                                                                \
    }
DEFINE_PAIRWISE(pw_f32, float)
DEFINE_PAIRWISE(pw_f64, double)

GSPO_API float gspo_pairwise_sum_f32(const float* a, int64_t n) { return pw_f32(a, n); }
GSPO_API double gspo_pairwise_sum_f64(const double* a, int64_t n) { return pw_f64(a, n); }

/* ------------------------------------------------------------------------------------
 * Feature cosine.  reference metrics.py:344-358 (+ core.py:182: runs in the dtype of data.x).
 *   nrm_i = max(sqrt(pw(x_i * x_i)), 1e-10);  xh_i = x_i / nrm_i  (elementwise divide)
 *   s = pw(xh_u * xh_v)  (products rounded first);  score = (double) max(s, 0)
 * ---------------------------------------------------------------------------------- */
#define DEFINE_FEATCOS(NAME, T, PW, SQRT, FLOOR)                                                     \
    GSPO_API void NAME(int64_t n, int64_t d, const T* x, const int64_t* indptr, const int32_t* indices, \
                       double* out, T* unit_out) {                                                   \
        T* unit = unit_out ? unit_out : (T*)malloc((size_t)(n * d > 0 ? n * d : 1) * sizeof(T));     \
        T* buf = (T*)malloc((size_t)(d > 0 ? d : 1) * sizeof(T));                                    \
        for (int64_t i = 0; i < n; ++i) {                                                            \
            for (int64_t f = 0; f < d; ++f) buf[f] = x[i * d + f] * x[i * d + f];                    \
            T nrm = SQRT(PW(buf, d));                                                                \
            if (!(nrm >= (T)FLOOR)) nrm = (T)FLOOR; /* np.maximum(norms, 1e-10) */                   \
            for (int64_t f = 0; f < d; ++f) unit[i * d + f] = x[i * d + f] / nrm;                    \
        }                                                                                            \
        for (int64_t u = 0; u < n; ++u)                                                              \
            for (int64_t p = indptr[u]; p < indptr[u + 1]; ++p) {                                    \
                int64_t v = indices[p];                                                              \
                for (int64_t f = 0; f < d; ++f) buf[f] = unit[u * d + f] * unit[v * d + f];          \
                T s = PW(buf, d);                                                                    \
                if (!(s >= (T)0)) s = (T)0;                                                          \
                out[p] = (double)s;                                                                  \
            }                                                                                        \
        free(buf);                                                                                   \
        if (!unit_out) free(unit);                                                                   \
    }
DEFINE_FEATCOS(gspo_featcos_f32, float, pw_f32, sqrtf, 1e-10)
DEFINE_FEATCOS(gspo_featcos_f64, double, pw_f64, sqrt, 1e-10)

/* ------------------------------------------------------------------------------------
 * Threshold selection.  reference core.py:232-240 under the stable-sort contract
 * (SURVEY App. A.4): order = stable ascending argsort(scores[0:nnz]);
 *   keep_lowest: order[:num_keep]   else: order[-num_keep:]
 * Python slicing quirks are part of the behaviour: order[-0:] is the WHOLE array, and
 * num_keep > nnz keeps everything.  mask has num_edges entries (positional aliasing,
 * SURVEY 8a-0); positions >= nnz are never set.
 * ---------------------------------------------------------------------------------- */
typedef struct { double s; int64_t i; } gspo_si;
static int cmp_si(const void* a, const void* b) {
    const gspo_si* p = (const gspo_si*)a; const gspo_si* q = (const gspo_si*)b;
    if (p->s < q->s) return -1;
    if (p->s > q->s) return 1;
    return (p->i > q->i) - (p->i < q->i);
}
static gspo_si* sorted_pairs(int64_t nnz, const double* scores) {
    gspo_si* o = (gspo_si*)malloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(gspo_si));
    for (int64_t i = 0; i < nnz; ++i) { o[i].s = scores[i]; o[i].i = i; }
    qsort(o, (size_t)nnz, sizeof(gspo_si), cmp_si);
    return o;
}

GSPO_API void gspo_select(int64_t nnz, const double* scores, int64_t num_edges, int64_t num_keep, int keep_lowest,
                          uint8_t* mask) {
    memset(mask, 0, (size_t)num_edges);
    gspo_si* o = sorted_pairs(nnz, scores);
    int64_t lo, hi;
    if (keep_lowest) { lo = 0; hi = num_keep < nnz ? num_keep : nnz; }
    else if (num_keep == 0) { lo = 0; hi = nnz; }            /* order[-0:] == order[:] */
    else { lo = nnz - num_keep; if (lo < 0) lo = 0; hi = nnz; }
    for (int64_t k = lo; k < hi; ++k) mask[o[k].i] = 1;
    free(o);
}

/* ------------------------------------------------------------------------------------
 * Degree-aware selection.  reference core.py:415-451, closed form of SURVEY App. A.5.
 *   G = for each source node (src = edge_index[0], positional), the last min(m, deg) entries of
 *       the stable ascending sort of its out-edge scores;
 *   if |G| < num_keep: add the first num_keep - |G| edges not in G along the reversed stable
 *       ascending order of all scores.
 * Requires nnz == num_edges (with duplicates the reference raises IndexError).
 * ---------------------------------------------------------------------------------- */
GSPO_API void gspo_degree_aware(int64_t num_edges, const int64_t* src, const double* scores, int64_t num_nodes,
                                int64_t num_keep, int64_t min_per_node, uint8_t* mask) {
    memset(mask, 0, (size_t)num_edges);
    int64_t* ptr = (int64_t*)calloc((size_t)num_nodes + 1, sizeof(int64_t));
    for (int64_t e = 0; e < num_edges; ++e) ptr[src[e] + 1]++;
    for (int64_t i = 0; i < num_nodes; ++i) ptr[i + 1] += ptr[i];
    int64_t* cur = (int64_t*)malloc(((size_t)num_nodes + 1) * sizeof(int64_t));
    memcpy(cur, ptr, ((size_t)num_nodes + 1) * sizeof(int64_t));
    gspo_si* grp = (gspo_si*)malloc((size_t)(num_edges > 0 ? num_edges : 1) * sizeof(gspo_si));
    for (int64_t e = 0; e < num_edges; ++e) { int64_t p = cur[src[e]]++; grp[p].s = scores[e]; grp[p].i = e; }
    int64_t have = 0;
    for (int64_t v = 0; v < num_nodes; ++v) {
        int64_t lo = ptr[v], hi = ptr[v + 1];
        if (hi == lo) continue;
        qsort(grp + lo, (size_t)(hi - lo), sizeof(gspo_si), cmp_si);
        int64_t k = min_per_node < hi - lo ? min_per_node : hi - lo;
        for (int64_t a = hi - k; a < hi; ++a) { if (!mask[grp[a].i]) { mask[grp[a].i] = 1; ++have; } }
    }
    if (have < num_keep) {
        gspo_si* o = sorted_pairs(num_edges, scores);
        for (int64_t k = num_edges - 1; k >= 0 && have < num_keep; --k)
            if (!mask[o[k].i]) { mask[o[k].i] = 1; ++have; }
        free(o);
    }
    free(grp); free(cur); free(ptr);
}

/* ------------------------------------------------------------------------------------
 * Approximate effective resistance.  reference metrics.py:232-298, SURVEY 3.3 / App. A.6.
 *   undirected edges = CSR positions with row < col, in CSR order (m of them)
 *   Y = B R:  Y[u,:] accumulates +R[e,:] / -R[e,:] over incident edges in ascending e
 *             (scipy csr_matvecs walks row u of B in column order)
 *   L_reg = diag(rowsum(data)) - A + reg*I   (raw data: multiplicities / weights)
 *   per column: scipy.sparse.linalg.cg(L_reg, y, x0=0, rtol, atol=0, maxiter) — test
 *     ||r|| < rtol*||b|| at the top of each iteration, no final test, ||b|| == 0 -> 0
 *   r_eff[p] = sum_j (Z[u,j] - Z[v,j])^2 ; max(., 1e-10)
 * Dot products / norms are plain sequential sums here (NumPy uses BLAS; not bit-matched —
 * ApproxER parity is a 1e-4 relative tolerance, BASELINE.json north_star).
 * R is [m, k] row-major (element (e, j) is PCG64 draw e*k + j).
 * ---------------------------------------------------------------------------------- */
static void lap_matvec(int64_t n, const int64_t* indptr, const int32_t* indices, const double* data,
                       const double* diag, const double* x, double* y) {
    for (int64_t i = 0; i < n; ++i) {
        double s = 0.0;
        int placed = 0;
        for (int64_t p = indptr[i]; p < indptr[i + 1]; ++p) {
            int64_t j = indices[p];
            double a = data ? data[p] : 1.0;
            if (j == i) { s += (diag[i] /* already has -a_ii folded */) * x[i]; placed = 1; continue; }
            if (!placed && j > i) { s += diag[i] * x[i]; placed = 1; }
            s += (-a) * x[j];
        }
        if (!placed) s += diag[i] * x[i];
        y[i] = s;
    }
}

GSPO_API int64_t gspo_count_upper(int64_t n, const int64_t* indptr, const int32_t* indices) {
    int64_t m = 0;
    for (int64_t u = 0; u < n; ++u)
        for (int64_t p = indptr[u]; p < indptr[u + 1]; ++p) m += (u < indices[p]);
    return m;
}

GSPO_API int gspo_approx_er(int64_t n, const int64_t* indptr, const int32_t* indices, const double* data,
                            int64_t k, const double* R, int64_t max_iters, double rtol, double reg,
                            double* out, int32_t* iters_out, double* z_out) {
    int64_t nnz = indptr[n];
    int64_t m = gspo_count_upper(n, indptr, indices);
    if (m == 0) { for (int64_t p = 0; p < nnz; ++p) out[p] = 0.0; return 0; }
    double* diag = (double*)malloc((size_t)n * sizeof(double));
    for (int64_t i = 0; i < n; ++i) {
        double deg = 0.0, self = 0.0;
        for (int64_t p = indptr[i]; p < indptr[i + 1]; ++p) {
            double a = data ? data[p] : 1.0;
            deg += a;
            if (indices[p] == i) self = a;
        }
        diag[i] = (deg - self) + reg;
    }
    double* Y = (double*)calloc((size_t)n * (size_t)k, sizeof(double));   /* [n, k] row-major */
    double* Z = z_out ? z_out : (double*)malloc((size_t)n * (size_t)k * sizeof(double));
    /* B @ R in the order csr_matvecs uses: row u of B lists its edges in ascending id */
    {
        /* edge id of each upper-triangle position, and per-node incident lists */
        int64_t e = 0;
        int64_t* cnt = (int64_t*)calloc((size_t)n + 1, sizeof(int64_t));
        for (int64_t u = 0; u < n; ++u)
            for (int64_t p = indptr[u]; p < indptr[u + 1]; ++p)
                if (u < indices[p]) { cnt[u + 1]++; cnt[indices[p] + 1]++; }
        for (int64_t i = 0; i < n; ++i) cnt[i + 1] += cnt[i];
        int64_t* cur = (int64_t*)malloc(((size_t)n + 1) * sizeof(int64_t));
        memcpy(cur, cnt, ((size_t)n + 1) * sizeof(int64_t));
        int64_t* inc_e = (int64_t*)malloc((size_t)(2 * m) * sizeof(int64_t));
        signed char* inc_s = (signed char*)malloc((size_t)(2 * m));
        for (int64_t u = 0; u < n; ++u)
            for (int64_t p = indptr[u]; p < indptr[u + 1]; ++p) {
                int64_t v = indices[p];
                if (u < v) {                    /* ascending e => lists come out sorted by e */
                    inc_e[cur[u]] = e; inc_s[cur[u]++] = 1;
                    inc_e[cur[v]] = e; inc_s[cur[v]++] = -1;
                    ++e;
                }
            }
        for (int64_t u = 0; u < n; ++u)
            for (int64_t q = cnt[u]; q < cnt[u + 1]; ++q) {
                const double* r = R + inc_e[q] * k;
                double sgn = (double)inc_s[q];
                double* y = Y + u * k;
                for (int64_t j = 0; j < k; ++j) y[j] += sgn * r[j];
            }
        free(inc_e); free(inc_s); free(cur); free(cnt);
    }
    double* b = (double*)malloc((size_t)n * sizeof(double));
    double* x = (double*)malloc((size_t)n * sizeof(double));
    double* r = (double*)malloc((size_t)n * sizeof(double));
    double* pv = (double*)malloc((size_t)n * sizeof(double));
    double* q = (double*)malloc((size_t)n * sizeof(double));
    for (int64_t j = 0; j < k; ++j) {
        double bb = 0.0;
        for (int64_t i = 0; i < n; ++i) { b[i] = Y[i * k + j]; bb += b[i] * b[i]; x[i] = 0.0; r[i] = b[i]; }
        double bnrm = sqrt(bb), atol = rtol * bnrm;
        int32_t it = 0;
        if (bnrm != 0.0) {
            double rho_prev = 0.0;
            for (it = 0; it < max_iters; ++it) {
                double rr = 0.0;
                for (int64_t i = 0; i < n; ++i) rr += r[i] * r[i];
                if (sqrt(rr) < atol) break;
                double rho = rr;
                if (it > 0) {
                    double beta = rho / rho_prev;
                    for (int64_t i = 0; i < n; ++i) { double t = pv[i] * beta; pv[i] = t + r[i]; }
                } else {
                    memcpy(pv, r, (size_t)n * sizeof(double));
                }
                lap_matvec(n, indptr, indices, data, diag, pv, q);
                double pq = 0.0;
                for (int64_t i = 0; i < n; ++i) pq += pv[i] * q[i];
                double alpha = rho / pq;
                for (int64_t i = 0; i < n; ++i) {
                    double t = alpha * pv[i]; x[i] = x[i] + t;
                    double s = alpha * q[i];  r[i] = r[i] - s;
                }
                rho_prev = rho;
            }
        }
        if (iters_out) iters_out[j] = it;
        for (int64_t i = 0; i < n; ++i) {
            double zi = x[i];
            if (isnan(zi) || isinf(zi)) zi = 0.0;   /* metrics.py:287-288 (applied only on failure there;
                                                       finite values are unaffected either way) */
            Z[i * k + j] = zi;
        }
    }
    for (int64_t u = 0; u < n; ++u)
        for (int64_t p = indptr[u]; p < indptr[u + 1]; ++p) {
            int64_t v = indices[p];
            double s = 0.0;
            for (int64_t j = 0; j < k; ++j) { double dlt = Z[u * k + j] - Z[v * k + j]; s += dlt * dlt; }
            if (isnan(s) || isinf(s)) s = 1e-10;
            out[p] = s > 1e-10 ? s : 1e-10;
        }
    free(b); free(x); free(r); free(pv); free(q); free(Y); free(diag);
    if (!z_out) free(Z);
    return 0;
}

/* ------------------------------------------------------------------------------------
 * Global metric backbone.  reference metric_backbone.py:58-112 (+ core.py:251-279).
 *   G = undirected graph over the positions with u < v, weight = min over duplicates of that
 *       orientation (the (v,u) positions do not contribute to G);
 *   dist = all-pairs shortest path lengths (Dijkstra, fp64 sums accumulated from the source);
 *   position idx is kept iff dist(u,v) is infinite or weights[idx] <= dist(u,v) + epsilon.
 * Binary-heap Dijkstra from every source; mask has E entries.
 * ---------------------------------------------------------------------------------- */
typedef struct { double d; int32_t v; } gspo_hn;

static void heap_push(gspo_hn* h, int64_t* sz, gspo_hn x) {
    int64_t i = (*sz)++;
    h[i] = x;
    while (i > 0) {
        int64_t p = (i - 1) / 2;
        if (h[p].d <= h[i].d) break;
        gspo_hn t = h[p]; h[p] = h[i]; h[i] = t; i = p;
    }
}
static gspo_hn heap_pop(gspo_hn* h, int64_t* sz) {
    gspo_hn top = h[0];
    h[0] = h[--(*sz)];
    int64_t i = 0;
    for (;;) {
        int64_t l = 2 * i + 1, r = l + 1, m = i;
        if (l < *sz && h[l].d < h[m].d) m = l;
        if (r < *sz && h[r].d < h[m].d) m = r;
        if (m == i) break;
        gspo_hn t = h[m]; h[m] = h[i]; h[i] = t; i = m;
    }
    return top;
}

GSPO_API int gspo_metric_backbone(int64_t n, int64_t E, const int64_t* row, const int64_t* col, const double* weights,
                                  double epsilon, uint8_t* mask) {
    /* undirected adjacency from the u < v positions, duplicates merged with min */
    int64_t* ptr = (int64_t*)calloc((size_t)n + 1, sizeof(int64_t));
    for (int64_t e = 0; e < E; ++e) if (row[e] < col[e]) { ptr[row[e] + 1]++; ptr[col[e] + 1]++; }
    for (int64_t i = 0; i < n; ++i) ptr[i + 1] += ptr[i];
    int64_t m2 = ptr[n];
    int32_t* adj = (int32_t*)malloc((size_t)(m2 > 0 ? m2 : 1) * sizeof(int32_t));
    double* aw = (double*)malloc((size_t)(m2 > 0 ? m2 : 1) * sizeof(double));
    int64_t* cur = (int64_t*)malloc(((size_t)n + 1) * sizeof(int64_t));
    memcpy(cur, ptr, ((size_t)n + 1) * sizeof(int64_t));
    for (int64_t e = 0; e < E; ++e) if (row[e] < col[e]) {
        adj[cur[row[e]]] = (int32_t)col[e]; aw[cur[row[e]]++] = weights[e];
        adj[cur[col[e]]] = (int32_t)row[e]; aw[cur[col[e]]++] = weights[e];
    }
    double* dist = (double*)malloc((size_t)n * sizeof(double));
    gspo_hn* heap = (gspo_hn*)malloc((size_t)(m2 + n + 1) * sizeof(gspo_hn));
    /* positions grouped by source so each Dijkstra serves all of a node's out-positions */
    int64_t* sptr = (int64_t*)calloc((size_t)n + 1, sizeof(int64_t));
    for (int64_t e = 0; e < E; ++e) sptr[row[e] + 1]++;
    for (int64_t i = 0; i < n; ++i) sptr[i + 1] += sptr[i];
    int64_t* spos = (int64_t*)malloc((size_t)(E > 0 ? E : 1) * sizeof(int64_t));
    memcpy(cur, sptr, ((size_t)n + 1) * sizeof(int64_t));
    for (int64_t e = 0; e < E; ++e) spos[cur[row[e]]++] = e;
    for (int64_t s = 0; s < n; ++s) {
        if (sptr[s + 1] == sptr[s]) continue;
        for (int64_t i = 0; i < n; ++i) dist[i] = INFINITY;
        int64_t sz = 0;
        dist[s] = 0.0;
        heap_push(heap, &sz, (gspo_hn){0.0, (int32_t)s});
        while (sz > 0) {
            gspo_hn t = heap_pop(heap, &sz);
            if (t.d > dist[t.v]) continue;
            for (int64_t p = ptr[t.v]; p < ptr[t.v + 1]; ++p) {
                double nd = t.d + aw[p];
                if (nd < dist[adj[p]]) { dist[adj[p]] = nd; heap_push(heap, &sz, (gspo_hn){nd, adj[p]}); }
            }
        }
        for (int64_t q = sptr[s]; q < sptr[s + 1]; ++q) {
            int64_t e = spos[q];
            double d = dist[col[e]];
            mask[e] = (isinf(d) || weights[e] <= d + epsilon) ? 1 : 0;
        }
    }
    free(spos); free(sptr); free(heap); free(dist); free(cur); free(aw); free(adj); free(ptr);
    return 0;
}
