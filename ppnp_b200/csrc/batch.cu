// batch.cu -- batch-main.py's per-batch gather / propagate on the compact top-k PPR matrix.
//
//   batch-main.py:140  ppr_sub = model.ppr[idx_batch]          B x n dense gather         |
//   batch-main.py:141  sel = (ppr_sub > 0).any(dim=0)          support union              | -> batch_support_kernel
//   batch-main.py:142  ppr_sub = ppr_sub[:, sel]               column compaction          |    (marks only)
//   batch-main.py:146  logits = ppr_sub @ encoder(X[sel])      model.py:65                  -> batch_propagate_kernel
// The reference reads 3 * B * n * 4 bytes of mostly-zero rows per batch; here a batch touches the
// nnz of its rows once (8 B per kept entry) plus the m x C block of encoder outputs.
#include "common.cuh"

namespace ppnp {
namespace {

// one warp per batch row: mark the columns it keeps
__global__ void __launch_bounds__(256)
batch_support_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                     const int64_t* __restrict__ idx_batch, int64_t B, uint8_t* __restrict__ mark) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= B) return;
    const int64_t r = idx_batch[w];
    const int64_t b = indptr[r], e = indptr[r + 1];
    for (int64_t t = b + lane; t < e; t += 32) mark[__ldg(indices + t)] = 1;
}

// The three lines batch-main.py:140-142 in ONE launch (a batch is small: B x k kept entries, n columns): one CTA
// zeroes a byte flag per column in shared memory, marks the columns the batch rows keep, scans the flags and
// writes sel[n] (the bool mask the caller indexes X with) and colmap[n] (position of a column inside the mask,
// -1 outside it).  Replaces memset + mark + bool conversion + cumsum + subtract (5-7 launches per batch).
constexpr int BSC_THREADS = 1024;

__global__ void __launch_bounds__(BSC_THREADS)
batch_support_colmap_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                            const int64_t* __restrict__ idx_batch, int64_t B, int n_cols,
                            uint8_t* __restrict__ sel, int32_t* __restrict__ colmap, int32_t* __restrict__ m_out) {
    extern __shared__ __align__(16) uint8_t flag[];     // n_cols rounded up to 16 bytes
    __shared__ int warp_tot[BSC_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_words = (n_cols + 15) / 16 * 4;
    for (int i = tid; i < n_words; i += BSC_THREADS) reinterpret_cast<uint32_t*>(flag)[i] = 0u;
    __syncthreads();
    for (int64_t w = warp; w < B; w += BSC_THREADS / 32) {
        const int64_t r = idx_batch[w];
        const int64_t b = indptr[r], e = indptr[r + 1];
        for (int64_t t = b + lane; t < e; t += 32) flag[__ldg(indices + t)] = 1;    // benign same-value races
    }
    __syncthreads();
    // every thread owns a contiguous span of columns; exclusive scan of the span counts over the CTA
    const int per = (n_cols + BSC_THREADS - 1) / BSC_THREADS;
    const int lo = min(tid * per, n_cols), hi = min(lo + per, n_cols);
    int cnt = 0;
    for (int c = lo; c < hi; ++c) cnt += flag[c];
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int v = warp_tot[lane];
        int iv = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, iv, o);
            if (lane >= o) iv += u;
        }
        warp_tot[lane] = iv - v;      // exclusive
        if (lane == 31 && m_out != nullptr) *m_out = iv;
    }
    __syncthreads();
    int pos = warp_tot[warp] + inc - cnt;
    for (int c = lo; c < hi; ++c) {
        const uint8_t f = flag[c];
        sel[c] = f;
        colmap[c] = f ? pos++ : -1;
    }
}

// one warp per batch row; G lanes share one kept entry (G = pow2 >= C chunk), 32/G entries in flight
template <int G>
__global__ void __launch_bounds__(256)
batch_propagate_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                       const float* __restrict__ val, const int64_t* __restrict__ idx_batch, int64_t B,
                       const int32_t* __restrict__ colmap, const float* __restrict__ Hsub, int64_t ld_h, int C,
                       float* __restrict__ out, int64_t ld_out) {
    constexpr int EPW = 32 / G;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= B) return;
    const int sub = lane / G, lc = lane % G;
    const int64_t r = idx_batch[w];
    const int64_t b = indptr[r], e = indptr[r + 1];
    for (int c0 = 0; c0 < C; c0 += G) {
        const int c = c0 + lc;
        float acc = 0.f;
        for (int64_t t = b + sub; t < e; t += EPW) {
            const int col = __ldg(indices + t);
            const float v = __ldg(val + t);
            const int pos = __ldg(colmap + col);     // -1: column outside the mask (ppr_sub[:, mask] drops it)
            if (c < C && pos >= 0) acc = fmaf(v, __ldg(Hsub + (int64_t)pos * ld_h + c), acc);
        }
#pragma unroll
        for (int o = G; o < 32; o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (sub == 0 && c < C) out[w * ld_out + c] = acc;
    }
}

// adjoint: dHsub[colmap[col], :] += val * dlogits[b, :]
template <int G>
__global__ void __launch_bounds__(256)
batch_propagate_t_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                         const float* __restrict__ val, const int64_t* __restrict__ idx_batch, int64_t B,
                         const int32_t* __restrict__ colmap, const float* __restrict__ dlogits, int64_t ld_g, int C,
                         float* __restrict__ dH, int64_t ld_dh) {
    constexpr int EPW = 32 / G;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= B) return;
    const int sub = lane / G, lc = lane % G;
    const int64_t r = idx_batch[w];
    const int64_t b = indptr[r], e = indptr[r + 1];
    for (int c0 = 0; c0 < C; c0 += G) {
        const int c = c0 + lc;
        const float g = (c < C) ? __ldg(dlogits + w * ld_g + c) : 0.f;
        for (int64_t t = b + sub; t < e; t += EPW) {
            const int col = __ldg(indices + t);
            const float v = __ldg(val + t);
            const int pos = __ldg(colmap + col);
            if (c < C && pos >= 0) atomicAdd(dH + (int64_t)pos * ld_dh + c, v * g);
        }
    }
}

}  // namespace
}  // namespace ppnp

extern "C" {

int ppnp_batch_support(const int64_t* indptr, const int32_t* indices, const int64_t* idx_batch, int64_t B,
                       uint8_t* mark, void* stream) {
    using namespace ppnp;
    PPNP_REQUIRE(indptr && indices && idx_batch && mark, "null pointer");
    PPNP_REQUIRE(B > 0, "B > 0");
    batch_support_kernel<<<(unsigned)((B + 7) / 8), 256, 0, as_stream(stream)>>>(indptr, indices, idx_batch, B, mark);
    PPNP_CHECK_LAUNCH("batch_support_kernel");
    return PPNP_OK;
}

int ppnp_batch_support_colmap(const int64_t* indptr, const int32_t* indices, const int64_t* idx_batch, int64_t B,
                              int64_t n_cols, uint8_t* sel, int32_t* colmap, int32_t* m_out, void* stream) {
    using namespace ppnp;
    PPNP_REQUIRE(indptr && indices && idx_batch && sel && colmap, "null pointer");
    PPNP_REQUIRE(B > 0 && n_cols > 0, "B > 0 and n_cols > 0");
    const size_t smem = (size_t)((n_cols + 15) / 16) * 16;
    if (smem > 200 * 1024) {
        set_error("ppnp_batch_support_colmap keeps one byte per column in shared memory: n_cols <= %d", 200 * 1024);
        return PPNP_ENOTSUP;
    }
    static thread_local size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        int rc = check_cuda(cudaFuncSetAttribute(batch_support_colmap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                            "cudaFuncSetAttribute batch_support_colmap_kernel");
        if (rc) return rc;
        configured = smem;
    }
    batch_support_colmap_kernel<<<1, BSC_THREADS, smem, as_stream(stream)>>>(indptr, indices, idx_batch, B, (int)n_cols, sel,
                                                                            colmap, m_out);
    PPNP_CHECK_LAUNCH("batch_support_colmap_kernel");
    return PPNP_OK;
}

int ppnp_batch_propagate(const int64_t* indptr, const int32_t* indices, const float* val, const int64_t* idx_batch,
                         int64_t B, const int32_t* colmap, const float* Hsub, int64_t ld_h, int32_t C, float* out,
                         int64_t ld_out, int32_t transpose, void* stream_) {
    using namespace ppnp;
    PPNP_REQUIRE(indptr && indices && val && idx_batch && colmap && Hsub && out, "null pointer");
    PPNP_REQUIRE(B > 0 && C > 0 && ld_h >= C && ld_out >= C, "bad shape");
    cudaStream_t stream = as_stream(stream_);
    const unsigned grid = (unsigned)((B + 7) / 8);
#define PPNP_BP(G_)                                                                                                    \
    do {                                                                                                               \
        if (!transpose)                                                                                                \
            batch_propagate_kernel<G_><<<grid, 256, 0, stream>>>(indptr, indices, val, idx_batch, B, colmap, Hsub,     \
                                                                 ld_h, C, out, ld_out);                                \
        else                                                                                                           \
            batch_propagate_t_kernel<G_><<<grid, 256, 0, stream>>>(indptr, indices, val, idx_batch, B, colmap, Hsub,   \
                                                                   ld_h, C, out, ld_out);                              \
    } while (0)
    if (C <= 4) PPNP_BP(4);
    else if (C <= 8) PPNP_BP(8);
    else if (C <= 16) PPNP_BP(16);
    else PPNP_BP(32);
#undef PPNP_BP
    PPNP_CHECK_LAUNCH("batch_propagate_kernel");
    return PPNP_OK;
}

}  // extern "C"
