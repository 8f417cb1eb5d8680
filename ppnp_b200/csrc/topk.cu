// topk.cu -- batch-main.py's PPR sparsification on the GPU.
//
//   batch-main.py:115   thresh, _ = ppr.topk(k, axis=-1)      -> topk_thresh_kernel (k-th largest per row,
//                                                                 exact radix select on the fp32 bit pattern)
//   batch-main.py:116   ppr[ppr < thresh[:, -1]] = 0          -> topk_mask_kernel; thresh[:, -1] is [n] and
//                                                                 broadcasts along the LAST axis, so entry
//                                                                 (i, j) is compared with thresh[j]
//   (new) compaction of the masked matrix to CSR: the reference keeps the dense n x n buffer and reads
//   B x n mostly-zero rows per batch (batch-main.py:140-142); the compact form is what batch.cu consumes.
// All of it is HBM-bound fp32 compare/select work: each pass streams the dense matrix once.
#include "common.cuh"

namespace ppnp {
namespace {

__device__ __forceinline__ uint32_t f32_key(float x) {
    const uint32_t u = __float_as_uint(x);
    return u ^ ((u >> 31) ? 0xffffffffu : 0x80000000u);  // ascending uint order == ascending float order
}
__device__ __forceinline__ float key_f32(uint32_t k) {
    const uint32_t u = k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu);
    return __uint_as_float(u);
}

// One CTA per row.  4 passes of an 8-bit MSB-first radix select; the row (<= a few hundred KB) is
// re-read from L2, never from HBM, after the first pass.
__global__ void __launch_bounds__(256)
topk_thresh_kernel(const float* __restrict__ ppr, int64_t n_cols, int64_t ld, int k, float* __restrict__ thresh) {
    __shared__ unsigned hist[256];
    __shared__ uint32_t s_prefix, s_mask;
    __shared__ int s_k;
    const float* row = ppr + (int64_t)blockIdx.x * ld;
    if (threadIdx.x == 0) { s_prefix = 0; s_mask = 0; s_k = k; }
    for (int shift = 24; shift >= 0; shift -= 8) {
        hist[threadIdx.x] = 0;
        __syncthreads();
        const uint32_t prefix = s_prefix, mask = s_mask;
        // all 32 lanes of a warp run the same number of iterations (match_any needs them)
        const int64_t iters = (n_cols + 255) / 256;
        for (int64_t it = 0; it < iters; ++it) {
            const int64_t c = it * 256 + threadIdx.x;
            bool in = false;
            uint32_t bucket = 0;
            if (c < n_cols) {
                const uint32_t key = f32_key(__ldg(row + c));
                in = (key & mask) == prefix;
                bucket = (key >> shift) & 255u;
            }
            const unsigned act = __ballot_sync(0xffffffffu, in);
            if (in) {
                const unsigned peers = __match_any_sync(act, bucket);
                if ((threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&hist[bucket], (unsigned)__popc(peers));
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int need = s_k;
            int b = 255;
            for (; b > 0; --b) {
                const int h = (int)hist[b];
                if (h >= need) break;
                need -= h;
            }
            s_k = need;
            s_prefix = prefix | ((uint32_t)b << shift);
            s_mask = mask | (255u << shift);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) thresh[blockIdx.x] = key_f32(s_prefix);
}

__global__ void __launch_bounds__(256)
topk_mask_kernel(float* __restrict__ ppr, int64_t n_rows, int64_t n_cols, int64_t ld, const float* __restrict__ thresh) {
    const int64_t i = blockIdx.y;
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_cols || i >= n_rows) return;
    float* p = ppr + i * ld + j;
    const float v = __ldcs(p);
    if (v < __ldg(thresh + j)) *p = 0.f;
}

// one warp per row
__global__ void __launch_bounds__(256)
dense_row_nnz_kernel(const float* __restrict__ ppr, int64_t n_rows, int64_t n_cols, int64_t ld, int32_t* __restrict__ row_nnz) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n_rows) return;
    const float* row = ppr + i * ld;
    int cnt = 0;
    for (int64_t c0 = 0; c0 < n_cols; c0 += 32) {
        const int64_t c = c0 + lane;
        const bool nz = (c < n_cols) && (__ldcs(row + c) > 0.f);
        cnt += __popc(__ballot_sync(0xffffffffu, nz));
    }
    if (lane == 0) row_nnz[i] = cnt;
}

__global__ void __launch_bounds__(256)
dense_to_csr_kernel(const float* __restrict__ ppr, int64_t n_rows, int64_t n_cols, int64_t ld,
                    const int64_t* __restrict__ indptr, int32_t* __restrict__ indices, float* __restrict__ val) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n_rows) return;
    const float* row = ppr + i * ld;
    int64_t o = indptr[i];
    for (int64_t c0 = 0; c0 < n_cols; c0 += 32) {
        const int64_t c = c0 + lane;
        const float v = (c < n_cols) ? __ldcs(row + c) : 0.f;
        const bool nz = v > 0.f;
        const unsigned m = __ballot_sync(0xffffffffu, nz);
        if (nz) {
            const int64_t at = o + __popc(m & ((1u << lane) - 1u));
            indices[at] = (int32_t)c;
            val[at] = v;
        }
        o += __popc(m);
    }
}

}  // namespace
}  // namespace ppnp

extern "C" {

int ppnp_topk_thresh(const float* ppr, int64_t n_rows, int64_t n_cols, int64_t ld, int32_t k, float* thresh, void* stream) {
    using namespace ppnp;
    PPNP_REQUIRE(ppr && thresh, "null pointer");
    PPNP_REQUIRE(n_rows > 0 && n_cols > 0 && ld >= n_cols, "bad shape");
    PPNP_REQUIRE(k >= 1 && k <= n_cols, "k out of range (torch.topk raises as well)");
    topk_thresh_kernel<<<(unsigned)n_rows, 256, 0, as_stream(stream)>>>(ppr, n_cols, ld, k, thresh);
    PPNP_CHECK_LAUNCH("topk_thresh_kernel");
    return PPNP_OK;
}

int ppnp_topk_mask(float* ppr, int64_t n_rows, int64_t n_cols, int64_t ld, const float* thresh, void* stream) {
    using namespace ppnp;
    PPNP_REQUIRE(ppr && thresh, "null pointer");
    PPNP_REQUIRE(n_rows > 0 && n_cols > 0 && ld >= n_cols, "bad shape");
    // grid.y is limited to 65535: fold rows beyond that into several launches
    for (int64_t r0 = 0; r0 < n_rows; r0 += 65535) {
        const int64_t nr = (n_rows - r0 < 65535) ? n_rows - r0 : 65535;
        dim3 grid((unsigned)((n_cols + 255) / 256), (unsigned)nr);
        topk_mask_kernel<<<grid, 256, 0, as_stream(stream)>>>(ppr + r0 * ld, nr, n_cols, ld, thresh);
        PPNP_CHECK_LAUNCH("topk_mask_kernel");
    }
    return PPNP_OK;
}

int ppnp_dense_row_nnz(const float* ppr, int64_t n_rows, int64_t n_cols, int64_t ld, int32_t* row_nnz, void* stream) {
    using namespace ppnp;
    PPNP_REQUIRE(ppr && row_nnz, "null pointer");
    PPNP_REQUIRE(n_rows > 0 && n_cols > 0 && ld >= n_cols, "bad shape");
    dense_row_nnz_kernel<<<(unsigned)((n_rows + 7) / 8), 256, 0, as_stream(stream)>>>(ppr, n_rows, n_cols, ld, row_nnz);
    PPNP_CHECK_LAUNCH("dense_row_nnz_kernel");
    return PPNP_OK;
}

int ppnp_dense_to_csr(const float* ppr, int64_t n_rows, int64_t n_cols, int64_t ld, const int64_t* indptr,
                      int32_t* indices, float* val, void* stream) {
    using namespace ppnp;
    PPNP_REQUIRE(ppr && indptr && indices && val, "null pointer");
    PPNP_REQUIRE(n_rows > 0 && n_cols > 0 && ld >= n_cols, "bad shape");
    dense_to_csr_kernel<<<(unsigned)((n_rows + 7) / 8), 256, 0, as_stream(stream)>>>(ppr, n_rows, n_cols, ld, indptr, indices, val);
    PPNP_CHECK_LAUNCH("dense_to_csr_kernel");
    return PPNP_OK;
}

}  // extern "C"
