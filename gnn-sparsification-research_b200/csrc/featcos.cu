// featcos.cu — feature-cosine edge scoring (kernel family K4).
//
// Replaces reference src/sparsification/metrics.py:344-358. The reference materialises two
// [E, d] gathers (`normalized[rows]`, `normalized[cols]`) and reduces their product with NumPy's
// pairwise summation; here each edge is a gathered row-dot evaluated in registers. This is a gather,
// not a dense contraction: no tensor cores. To reproduce the reference's keep-masks the arithmetic
// follows NumPy bit for bit: products rounded to the feature dtype, the 8-accumulator
// / 128-element-leaf pairwise tree, IEEE divide and sqrt, no FMA contraction.
//
// Mapping: 8 lanes per edge, lane j owns pairwise accumulator j (elements j, j+8, j+16, ... of a leaf),
// so one load instruction covers a 32-byte sector per edge and four edges share a warp; the three
// shuffle-xor steps reproduce ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) exactly (IEEE add is commutative).
#include "common.cuh"

namespace gsp {
namespace {

constexpr int kThreads = 256;
constexpr int kGroup = 8;  // lanes per edge == NumPy's accumulator count

template <typename T> struct Arith;
template <> struct Arith<float> {
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
    static __device__ __forceinline__ float floor_norm() { return 1e-10f; }  // (float)1e-10: np.maximum(norms, 1e-10)
};
template <> struct Arith<double> {
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double sqrt(double a) { return __dsqrt_rn(a); }
    static __device__ __forceinline__ double floor_norm() { return 1e-10; }
};

// One pairwise leaf (n <= 128) summed cooperatively by an 8-lane group; every lane returns the leaf sum.
template <typename T, typename Elem>
__device__ __forceinline__ T leaf_sum(Elem elem, int n, int j) {
    using A = Arith<T>;
    if (n < 8) {
        T res = T(0);
        for (int i = 0; i < n; ++i) res = A::add(res, elem(i));
        return res;
    }
    const int main_n = n - (n % 8);
    T r = elem(j);
#pragma unroll 4
    for (int i = 8; i < main_n; i += 8) r = A::add(r, elem(i + j));
    r = A::add(r, __shfl_xor_sync(0xffffffffu, r, 1));
    r = A::add(r, __shfl_xor_sync(0xffffffffu, r, 2));
    r = A::add(r, __shfl_xor_sync(0xffffffffu, r, 4));
    for (int i = main_n; i < n; ++i) r = A::add(r, elem(i));
    return r;
}

// NumPy pairwise_sum over n elements: recursion on halves rounded down to a multiple of 8, leaves <= 128.
template <typename T, typename Elem>
__device__ T pairwise_sum(Elem elem, int n, int j) {
    using A = Arith<T>;
    if (n <= 128) return leaf_sum<T>(elem, n, j);
    struct Frame { int off, n, state; T left; };
    Frame stack[26];
    int sp = 0;
    stack[sp++] = Frame{0, n, 0, T(0)};
    T ret = T(0);
    while (sp > 0) {
        Frame& f = stack[sp - 1];
        if (f.n <= 128) {
            const int off = f.off;
            ret = leaf_sum<T>([&](int i) { return elem(off + i); }, f.n, j);
            --sp;
            continue;
        }
        int n2 = f.n / 2;
        n2 -= n2 % 8;
        if (f.state == 0) {
            f.state = 1;
            stack[sp++] = Frame{f.off, n2, 0, T(0)};
        } else if (f.state == 1) {
            f.left = ret;
            f.state = 2;
            stack[sp++] = Frame{f.off + n2, f.n - n2, 0, T(0)};
        } else {
            ret = A::add(f.left, ret);
            --sp;
        }
    }
    return ret;
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
normalize_kernel(int64_t n, int dim, const T* __restrict__ x, int64_t ld, T* __restrict__ xhat, int64_t ld_out) {
    using A = Arith<T>;
    const int j = threadIdx.x & (kGroup - 1);
    const int64_t group = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / kGroup;
    const int64_t num_groups = (int64_t)gridDim.x * blockDim.x / kGroup;
    const int64_t rounds = (n + num_groups - 1) / num_groups;  // whole warp iterates together (shuffles)
    for (int64_t it = 0; it < rounds; ++it) {
        const int64_t node = it * num_groups + group;
        const bool live = node < n;
        const T* row = x + (live ? node : 0) * ld;
        T ss = pairwise_sum<T>([&](int i) { T v = row[i]; return A::mul(v, v); }, dim, j);
        T nrm = A::sqrt(ss);
        nrm = nrm < A::floor_norm() ? A::floor_norm() : nrm;  // np.maximum(norms, 1e-10); NaN propagates
        if (live) {
            T* out = xhat + node * ld_out;
            for (int i = j; i < dim; i += kGroup) out[i] = A::div(row[i], nrm);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
featcos_kernel(int64_t e_begin, int64_t e_end, const int32_t* __restrict__ rows, const int32_t* __restrict__ indices,
               const T* __restrict__ xhat, int dim, int64_t ld, double* __restrict__ out) {
    using A = Arith<T>;
    const int j = threadIdx.x & (kGroup - 1);
    const int64_t group = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / kGroup;
    const int64_t num_groups = (int64_t)gridDim.x * blockDim.x / kGroup;
    const int64_t count = e_end - e_begin;
    const int64_t rounds = (count + num_groups - 1) / num_groups;
    for (int64_t it = 0; it < rounds; ++it) {
        const int64_t k = it * num_groups + group;
        const bool live = k < count;
        const int64_t e = e_begin + (live ? k : 0);
        const T* xu = xhat + (int64_t)__ldg(rows + e) * ld;
        const T* xv = xhat + (int64_t)__ldg(indices + e) * ld;
        T s = pairwise_sum<T>([&](int i) { return A::mul(__ldg(xu + i), __ldg(xv + i)); }, dim, j);
        s = s < T(0) ? T(0) : s;  // np.maximum(scores, 0.0)
        if (live && j == 0) out[k] = (double)s;
    }
}

// ---- packed fp32 fast path (dim % 32 == 0, dim <= 128: one pairwise leaf) -----------------------------------------
// The normalised row is stored accumulator-major: packed[q*32 + j*4 + k] = xhat[j + 8*(4q + k)], i.e. the four
// consecutive chain elements (4q .. 4q+3) of NumPy accumulator j sit in one float4 and the eight lanes of an edge
// group read 128 contiguous bytes per load. Same arithmetic, same association order, a quarter of the load
// instructions (the scalar kernel is LSU-issue bound, not HBM bound).
template <int kQuads>
__global__ void __launch_bounds__(kThreads)
normalize_packed_kernel(int64_t n, const float* __restrict__ x, int64_t ld, float* __restrict__ packed, int64_t ld_out) {
    using A = Arith<float>;
    const int j = threadIdx.x & (kGroup - 1);
    const int64_t group = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / kGroup;
    const int64_t num_groups = (int64_t)gridDim.x * blockDim.x / kGroup;
    const int64_t rounds = (n + num_groups - 1) / num_groups;
    for (int64_t it = 0; it < rounds; ++it) {
        const int64_t node = it * num_groups + group;
        const bool live = node < n;
        const float* row = x + (live ? node : 0) * ld;
        float v[4 * kQuads];                                   // accumulator j's chain: elements j, j+8, j+16, ...
#pragma unroll
        for (int m = 0; m < 4 * kQuads; ++m) v[m] = __ldg(row + j + 8 * m);
        float r = A::mul(v[0], v[0]);                          // the leaf's sum of squares, NumPy's order
#pragma unroll
        for (int m = 1; m < 4 * kQuads; ++m) r = A::add(r, A::mul(v[m], v[m]));
        r = A::add(r, __shfl_xor_sync(0xffffffffu, r, 1));
        r = A::add(r, __shfl_xor_sync(0xffffffffu, r, 2));
        r = A::add(r, __shfl_xor_sync(0xffffffffu, r, 4));
        float nrm = A::sqrt(r);
        nrm = nrm < A::floor_norm() ? A::floor_norm() : nrm;
        if (live) {
            float4* out = reinterpret_cast<float4*>(packed + node * ld_out) + j;
#pragma unroll
            for (int q = 0; q < kQuads; ++q)
                out[q * 8] = make_float4(A::div(v[4 * q], nrm), A::div(v[4 * q + 1], nrm), A::div(v[4 * q + 2], nrm),
                                         A::div(v[4 * q + 3], nrm));
        }
    }
}

template <int kQuads>   // kQuads = dim / 32 float4 loads per lane and row
__global__ void __launch_bounds__(kThreads)
featcos_packed_kernel(int64_t e_begin, int64_t e_end, const int32_t* __restrict__ rows, const int32_t* __restrict__ indices,
                      const float* __restrict__ packed, int64_t ld, double* __restrict__ out) {
    using A = Arith<float>;
    const int j = threadIdx.x & (kGroup - 1);
    const int64_t group = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / kGroup;
    const int64_t num_groups = (int64_t)gridDim.x * blockDim.x / kGroup;
    const int64_t count = e_end - e_begin;
    const int64_t rounds = (count + num_groups - 1) / num_groups;
    for (int64_t it = 0; it < rounds; ++it) {
        const int64_t k = it * num_groups + group;
        const bool live = k < count;
        const int64_t e = e_begin + (live ? k : 0);
        const float4* pu = reinterpret_cast<const float4*>(packed + (int64_t)__ldg(rows + e) * ld) + j;
        const float4* pv = reinterpret_cast<const float4*>(packed + (int64_t)__ldg(indices + e) * ld) + j;
        float4 a[kQuads], b[kQuads];
#pragma unroll
        for (int q = 0; q < kQuads; ++q) {
            a[q] = __ldg(pu + q * 8);
            b[q] = __ldg(pv + q * 8);
        }
        float r = A::mul(a[0].x, b[0].x);                   // r[j] = a[j]
        r = A::add(r, A::mul(a[0].y, b[0].y));
        r = A::add(r, A::mul(a[0].z, b[0].z));
        r = A::add(r, A::mul(a[0].w, b[0].w));
#pragma unroll
        for (int q = 1; q < kQuads; ++q) {
            r = A::add(r, A::mul(a[q].x, b[q].x));
            r = A::add(r, A::mul(a[q].y, b[q].y));
            r = A::add(r, A::mul(a[q].z, b[q].z));
            r = A::add(r, A::mul(a[q].w, b[q].w));
        }
        r = A::add(r, __shfl_xor_sync(0xffffffffu, r, 1));
        r = A::add(r, __shfl_xor_sync(0xffffffffu, r, 2));
        r = A::add(r, __shfl_xor_sync(0xffffffffu, r, 4));
        r = r < 0.0f ? 0.0f : r;
        if (live && j == 0) out[k] = (double)r;
    }
}

inline bool packed_ok(int32_t dim, int64_t ld) { return dim >= 32 && dim <= 128 && dim % 32 == 0 && ld % 4 == 0; }

template <typename T>
int normalize(int64_t n, int32_t dim, const T* x, int64_t ld, T* xhat, int64_t ld_out, cudaStream_t s) {
    GSP_REQUIRE(n >= 0 && dim >= 0, "negative size");
    if (n == 0 || dim == 0) return GSP_OK;
    GSP_REQUIRE(x && xhat, "NULL feature pointer");
    GSP_REQUIRE(ld >= dim && ld_out >= dim, "row stride smaller than dim");
    normalize_kernel<T><<<grid_for(n * kGroup, kThreads, 8), kThreads, 0, s>>>(n, dim, x, ld, xhat, ld_out);
    GSP_CHECK_LAUNCH();
    return GSP_OK;
}

template <typename T>
int featcos(const Graph* g, const T* xhat, int32_t dim, int64_t ld, int64_t e_begin, int64_t e_end, double* out,
            cudaStream_t s) {
    GSP_REQUIRE(g != nullptr, "graph is NULL");
    GSP_REQUIRE(e_begin >= 0 && e_begin <= e_end && e_end <= g->nnz, "edge range outside [0, nnz]");
    if (e_begin == e_end) return GSP_OK;
    GSP_REQUIRE(out != nullptr, "d_score is NULL");
    GSP_REQUIRE(dim >= 0 && ld >= dim, "bad feature shape");
    GSP_REQUIRE(dim == 0 || xhat != nullptr, "d_xhat is NULL");
    featcos_kernel<T><<<grid_for((e_end - e_begin) * kGroup, kThreads, 8), kThreads, 0, s>>>(
        e_begin, e_end, g->rows, g->indices, xhat, dim, ld, out);
    GSP_CHECK_LAUNCH();
    return GSP_OK;
}

}  // namespace
}  // namespace gsp

using namespace gsp;

GSP_API int gsp_featcos_normalize_f32(int64_t num_nodes, int32_t dim, const float* d_x, int64_t ld, float* d_xhat,
                                      int64_t ld_out, void* stream) {
    return normalize<float>(num_nodes, dim, d_x, ld, d_xhat, ld_out, as_stream(stream));
}
GSP_API int gsp_featcos_normalize_f64(int64_t num_nodes, int32_t dim, const double* d_x, int64_t ld, double* d_xhat,
                                      int64_t ld_out, void* stream) {
    return normalize<double>(num_nodes, dim, d_x, ld, d_xhat, ld_out, as_stream(stream));
}
GSP_API int gsp_featcos_f32(const gsp_graph* g, const float* d_xhat, int32_t dim, int64_t ld, int64_t e_begin,
                            int64_t e_end, double* d_score, void* stream) {
    return featcos<float>(reinterpret_cast<const Graph*>(g), d_xhat, dim, ld, e_begin, e_end, d_score, as_stream(stream));
}
GSP_API int gsp_featcos_f64(const gsp_graph* g, const double* d_xhat, int32_t dim, int64_t ld, int64_t e_begin,
                            int64_t e_end, double* d_score, void* stream) {
    return featcos<double>(reinterpret_cast<const Graph*>(g), d_xhat, dim, ld, e_begin, e_end, d_score, as_stream(stream));
}

// Packed-layout variants (fp32, dim in {32, 64, 96, 128}): d_packed is an opaque accumulator-major copy of the
// normalised features, valid only as input of gsp_featcos_f32_packed. Results are bit-identical to the plain pair.
GSP_API int gsp_featcos_normalize_f32_packed(int64_t num_nodes, int32_t dim, const float* d_x, int64_t ld, float* d_packed,
                                             int64_t ld_out, void* stream) {
    GSP_REQUIRE(num_nodes >= 0, "negative size");
    GSP_REQUIRE(packed_ok(dim, ld_out) && ld >= dim && ld_out >= dim, "packed layout needs dim in {32,64,96,128} and ld_out % 4 == 0");
    if (num_nodes == 0) return GSP_OK;
    GSP_REQUIRE(d_x && d_packed, "NULL feature pointer");
    GSP_REQUIRE((reinterpret_cast<uintptr_t>(d_packed) & 15) == 0, "d_packed must be 16-byte aligned");
    const int grid = grid_for(num_nodes * kGroup, kThreads, 8);
    cudaStream_t s = as_stream(stream);
    switch (dim / 32) {
        case 1: normalize_packed_kernel<1><<<grid, kThreads, 0, s>>>(num_nodes, d_x, ld, d_packed, ld_out); break;
        case 2: normalize_packed_kernel<2><<<grid, kThreads, 0, s>>>(num_nodes, d_x, ld, d_packed, ld_out); break;
        case 3: normalize_packed_kernel<3><<<grid, kThreads, 0, s>>>(num_nodes, d_x, ld, d_packed, ld_out); break;
        default: normalize_packed_kernel<4><<<grid, kThreads, 0, s>>>(num_nodes, d_x, ld, d_packed, ld_out); break;
    }
    GSP_CHECK_LAUNCH();
    return GSP_OK;
}

GSP_API int gsp_featcos_f32_packed(const gsp_graph* gg, const float* d_packed, int32_t dim, int64_t ld, int64_t e_begin,
                                   int64_t e_end, double* d_score, void* stream) {
    const Graph* g = reinterpret_cast<const Graph*>(gg);
    GSP_REQUIRE(g != nullptr, "graph is NULL");
    GSP_REQUIRE(e_begin >= 0 && e_begin <= e_end && e_end <= g->nnz, "edge range outside [0, nnz]");
    GSP_REQUIRE(packed_ok(dim, ld) && ld >= dim, "packed layout needs dim in {32,64,96,128} and ld % 4 == 0");
    if (e_begin == e_end) return GSP_OK;
    GSP_REQUIRE(d_packed && d_score, "NULL argument");
    GSP_REQUIRE((reinterpret_cast<uintptr_t>(d_packed) & 15) == 0, "d_packed must be 16-byte aligned");
    cudaStream_t s = as_stream(stream);
    const int grid = grid_for((e_end - e_begin) * kGroup, kThreads, 8);
    switch (dim / 32) {
        case 1: featcos_packed_kernel<1><<<grid, kThreads, 0, s>>>(e_begin, e_end, g->rows, g->indices, d_packed, ld, d_score); break;
        case 2: featcos_packed_kernel<2><<<grid, kThreads, 0, s>>>(e_begin, e_end, g->rows, g->indices, d_packed, ld, d_score); break;
        case 3: featcos_packed_kernel<3><<<grid, kThreads, 0, s>>>(e_begin, e_end, g->rows, g->indices, d_packed, ld, d_score); break;
        default: featcos_packed_kernel<4><<<grid, kThreads, 0, s>>>(e_begin, e_end, g->rows, g->indices, d_packed, ld, d_score); break;
    }
    GSP_CHECK_LAUNCH();
    return GSP_OK;
}
