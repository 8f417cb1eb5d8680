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
#include "common.cuh"

namespace ppnp {
namespace {

constexpr int PPR_THREADS = 128;
constexpr int PPR_CPT = 4;                        // columns per thread
constexpr int PPR_TILE = PPR_THREADS * PPR_CPT;   // 512 columns: n * 2 KB per slab

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

__global__ void __launch_bounds__(PPR_THREADS)
ppr_step_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                const float* __restrict__ val, int64_t n, float alpha,
                const float* __restrict__ Pin, float* __restrict__ Pout) {
    const int64_t i = blockIdx.x;
    const int64_t c0 = (int64_t)blockIdx.y * PPR_TILE + threadIdx.x;
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
        if (c < n) __stcs(orow + c, oma * acc[k] + ((c == i) ? alpha : 0.f));
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
    const dim3 grid((unsigned)n, (unsigned)((n + PPR_TILE - 1) / PPR_TILE));
    for (int k = 2; k <= K; ++k) {
        const float* src = dst;
        dst = ((K - k) % 2 == 0) ? Pi : scratch;
        ppr_step_kernel<<<grid, PPR_THREADS, 0, stream>>>(indptr, indices, val, n, alpha, src, dst);
        PPNP_CHECK_LAUNCH("ppr_step_kernel");
    }
    return PPNP_OK;
}
