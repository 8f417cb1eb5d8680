// ppr_dense.cu -- exact PPNP's matrix  Pi = alpha (I - (1-alpha) A_hat)^-1  on the GPU.
//
// Replaces helpers.py:68-71 compute_ppr (dense fp64 np.linalg.inv on the host, O(n^3)) by a
// power iteration over all n right-hand sides at once:
//     Pi_0 = I,   Pi_{k+1} = (1-alpha) A_hat Pi_k + alpha I        (K -> inf limit = the inverse,
//     error (1-alpha)^K in the spectral norm because rho(A_hat) = 1)
// i.e. the APPNP recurrence with H = I.  Each step is a sparse x dense
// product that streams the n x n iterate once: 3 n^2 * 4 bytes per step (read, write, nothing
// else of size n^2), HBM-bound.  One CTA owns (row i, column tile); consecutive CTAs share the
// tile, so the tile's n x TILE slab of Pi_k stays in L2 while the rows that gather from it run.
//
// ppnp_ppr_dense_cheb runs the same fixed point with Chebyshev acceleration (the iteration matrix
// (1-alpha) A_hat is symmetric -- or similar to a symmetric matrix in 'rw' mode -- with its spectrum in
// [-rho, rho], rho = 1 - alpha):
//     x_1 = G x_0 + b,   x_{k+1} = w_{k+1} (G x_k + b) + (1 - w_{k+1}) x_{k-1},
//     w_2 = 1 / (1 - rho^2 / 2),   w_{k+1} = 1 / (1 - rho^2 w_k / 4)
// whose error falls like sigma^k, sigma = (1 - sqrt(1 - rho^2)) / rho = 0.627 for alpha = 0.1 against
// 0.9 for the plain iteration: ~40 steps instead of ~150 for the fp32 floor, same bytes per step (the
// x_{k-1} term is read where x_{k+1} is written, in place).
#include "common.cuh"

namespace ppnp {
namespace {

constexpr int PPR_THREADS = 128;
constexpr int PPR_CPT = 4;                        // columns per thread
constexpr int PPR_TILE = PPR_THREADS * PPR_CPT;   // 512 columns: n * 2 KB per slab
#ifndef PPNP_PPR_ROWS
#define PPNP_PPR_ROWS 8                           // rows per CTA of a step: one CTA per (row, tile) is 769 k CTAs of ~13 KB
#endif                                            // of traffic each at PubMed shape -- the block scheduler, not HBM, sets the pace
constexpr int PPR_ROWS = PPNP_PPR_ROWS;

// Pi_1 = (1-alpha) A_hat + alpha I, written densely (one CTA per row).
__global__ void __launch_bounds__(256)
ppr_init_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                const float* __restrict__ val, int64_t n, float alpha, float* __restrict__ P) {
    const int64_t i = blockIdx.x;
    float* row = P + i * n;
    for (int64_t c = threadIdx.x; c < n; c += blockDim.x) row[c] = 0.f;
    __syncthreads();
    const int b = indptr[i], e = indptr[i + 1];
    const float oma = 1.0f - alpha;
    bool diag_seen = false;
    for (int t = b + threadIdx.x; t < e; t += blockDim.x) {
        const int j = indices[t];
        float v = oma * val[t];
        if (j == (int)i) { v += alpha; diag_seen = true; }
        row[j] = v;
    }
    // A_hat always holds the diagonal (self loop), but stay correct if a caller passes a CSR without it
    const int any = __syncthreads_or(diag_seen ? 1 : 0);
    if (!any && threadIdx.x == 0) row[i] = alpha;
}

// K == 0: Pi = I
__global__ void ppr_identity_kernel(int64_t n, float* __restrict__ P) {
    const int64_t i = blockIdx.x;
    float* row = P + i * n;
    for (int64_t c = threadIdx.x; c < n; c += blockDim.x) row[c] = (c == i) ? 1.f : 0.f;
}

// CHEB: Pout holds x_{k-1} on entry (or the identity, implicitly, when prev_identity) and x_{k+1} on exit.
template <bool CHEB>
__global__ void __launch_bounds__(PPR_THREADS)
ppr_step_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                const float* __restrict__ val, int64_t n, float alpha,
                const float* __restrict__ Pin, float* Pout, float omega, int prev_identity) {
    const int64_t c0 = (int64_t)blockIdx.y * PPR_TILE + threadIdx.x;
    const int64_t i_end = ((int64_t)blockIdx.x + 1) * PPR_ROWS < n ? ((int64_t)blockIdx.x + 1) * PPR_ROWS : n;
    for (int64_t i = (int64_t)blockIdx.x * PPR_ROWS; i < i_end; ++i) {
    float acc[PPR_CPT];
#pragma unroll
    for (int k = 0; k < PPR_CPT; ++k) acc[k] = 0.f;
    const int b = indptr[i], e = indptr[i + 1];
    int t = b;
    // two edges per iteration: 2 * PPR_CPT independent loads in flight per thread
    for (; t + 1 < e; t += 2) {
        const int j0 = __ldg(indices + t), j1 = __ldg(indices + t + 1);
        const float w0 = __ldg(val + t), w1 = __ldg(val + t + 1);
        const float* r0 = Pin + (int64_t)j0 * n;
        const float* r1 = Pin + (int64_t)j1 * n;
        float x0[PPR_CPT], x1[PPR_CPT];
#pragma unroll
        for (int k = 0; k < PPR_CPT; ++k) {
            const int64_t c = c0 + k * PPR_THREADS;
            x0[k] = (c < n) ? __ldg(r0 + c) : 0.f;
            x1[k] = (c < n) ? __ldg(r1 + c) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < PPR_CPT; ++k) { acc[k] = fmaf(w0, x0[k], acc[k]); acc[k] = fmaf(w1, x1[k], acc[k]); }
    }
    if (t < e) {
        const int j0 = __ldg(indices + t);
        const float w0 = __ldg(val + t);
        const float* r0 = Pin + (int64_t)j0 * n;
#pragma unroll
        for (int k = 0; k < PPR_CPT; ++k) {
            const int64_t c = c0 + k * PPR_THREADS;
            if (c < n) acc[k] = fmaf(w0, __ldg(r0 + c), acc[k]);
        }
    }
    const float oma = 1.0f - alpha;
    float* orow = Pout + i * n;
#pragma unroll
    for (int k = 0; k < PPR_CPT; ++k) {
        const int64_t c = c0 + k * PPR_THREADS;
        if (c < n) {
            const float plain = oma * acc[k] + ((c == i) ? alpha : 0.f);
            if (CHEB) {
                const float prev = prev_identity ? ((c == i) ? 1.f : 0.f) : __ldcs(orow + c);
                __stcs(orow + c, fmaf(omega, plain, (1.0f - omega) * prev));
            } else {
                __stcs(orow + c, plain);
            }
        }
    }
    }
}

}  // namespace
}  // namespace ppnp

extern "C" int ppnp_ppr_dense(const int32_t* indptr, const int32_t* indices, const float* val, int64_t n,
                              float alpha, int32_t K, float* Pi, float* scratch, void* stream_) {
    using namespace ppnp;
    PPNP_REQUIRE(indptr && indices && val && Pi, "null pointer");
    PPNP_REQUIRE(n > 0 && n < ((int64_t)1 << 31), "0 < n < 2^31");
    PPNP_REQUIRE(K >= 0, "K >= 0");
    PPNP_REQUIRE(K <= 1 || (scratch != nullptr && scratch != Pi), "scratch buffer required for K > 1");
    cudaStream_t stream = as_stream(stream_);
    if (K == 0) {
        ppr_identity_kernel<<<(unsigned)n, 256, 0, stream>>>(n, Pi);
        PPNP_CHECK_LAUNCH("ppr_identity_kernel");
        return PPNP_OK;
    }
    // step k (1-based) writes Pi when (K - k) is even
    float* dst = ((K - 1) % 2 == 0) ? Pi : scratch;
    ppr_init_kernel<<<(unsigned)n, 256, 0, stream>>>(indptr, indices, val, n, alpha, dst);
    PPNP_CHECK_LAUNCH("ppr_init_kernel");
    const dim3 grid((unsigned)((n + PPR_ROWS - 1) / PPR_ROWS), (unsigned)((n + PPR_TILE - 1) / PPR_TILE));
    for (int k = 2; k <= K; ++k) {
        const float* src = dst;
        dst = ((K - k) % 2 == 0) ? Pi : scratch;
        ppr_step_kernel<false><<<grid, PPR_THREADS, 0, stream>>>(indptr, indices, val, n, alpha, src, dst, 1.0f, 0);
        PPNP_CHECK_LAUNCH("ppr_step_kernel");
    }
    return PPNP_OK;
}

extern "C" int ppnp_ppr_dense_cheb(const int32_t* indptr, const int32_t* indices, const float* val, int64_t n,
                                   float alpha, int32_t K, float* Pi, float* scratch, void* stream_) {
    using namespace ppnp;
    PPNP_REQUIRE(indptr && indices && val && Pi, "null pointer");
    PPNP_REQUIRE(n > 0 && n < ((int64_t)1 << 31), "0 < n < 2^31");
    PPNP_REQUIRE(K >= 1, "K >= 1");
    PPNP_REQUIRE(alpha > 0.f && alpha < 1.f, "0 < alpha < 1");
    PPNP_REQUIRE(K <= 1 || (scratch != nullptr && scratch != Pi), "scratch buffer required for K > 1");
    cudaStream_t stream = as_stream(stream_);
    // x_k lives in `a` for odd k and in `b` for even k; the result x_K must land in Pi
    float* a = (K % 2 == 1) ? Pi : scratch;
    float* b = (K % 2 == 1) ? scratch : Pi;
    ppr_init_kernel<<<(unsigned)n, 256, 0, stream>>>(indptr, indices, val, n, alpha, a);      // x_1 = G I + alpha I
    PPNP_CHECK_LAUNCH("ppr_init_kernel");
    const dim3 grid((unsigned)((n + PPR_ROWS - 1) / PPR_ROWS), (unsigned)((n + PPR_TILE - 1) / PPR_TILE));
    const double rho2 = (1.0 - (double)alpha) * (1.0 - (double)alpha);
    double w = 1.0;
    for (int k = 2; k <= K; ++k) {
        w = (k == 2) ? 1.0 / (1.0 - rho2 / 2.0) : 1.0 / (1.0 - rho2 * w / 4.0);
        const float* src = (k % 2 == 0) ? a : b;       // x_{k-1}
        float* dst = (k % 2 == 0) ? b : a;             // holds x_{k-2} (k = 2: nothing yet, x_0 = I is implicit)
        ppr_step_kernel<true><<<grid, PPR_THREADS, 0, stream>>>(indptr, indices, val, n, alpha, src, dst, (float)w, k == 2 ? 1 : 0);
        PPNP_CHECK_LAUNCH("ppr_step_kernel<cheb>");
    }
    return PPNP_OK;
}
