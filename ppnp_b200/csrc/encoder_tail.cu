// encoder_tail.cu -- the encoder's last linear layer fused with the propagation's row scaling
// (model.py:51 `nn.Linear(hidden_dim, n_classes)` feeding model.py:63).
//
//   forward   out[i, :] = s_i * (A[i, :] @ W^T + bias)          A: n x hidden, W: C x hidden (nn.Linear layout)
//   backward  dA[i, :]  = s_i * (dOut[i, :] @ W)
//             dW        = sum_i s_i * dOut[i, :]^T A[i, :]       (fixed-order two-stage reduction: deterministic)
//             dbias     = sum_i s_i * dOut[i, :]
//
// With s = D^-1/2 the forward writes Y0 = D^-1/2 H straight from the hidden activations: the K propagation
// steps then run value-free from the first one (PPNP_MODE_SYM_Y0) and the stored values of A_hat (4 bytes per
// edge: 8.4 GB at BASELINE config 5) are neither read nor kept.  H itself is never materialised.
// HBM-bound: reads n * hidden * 4 bytes, writes n * C * 4; W (<= 64 x 256 floats) lives in shared memory.
#include "common.cuh"

namespace ppnp {
namespace {

constexpr int ROWS = 32;       // rows of A per tile
constexpr int THREADS = 256;

// out tile = A tile (ROWS x hidden) @ W^T: thread t owns output columns c = t % CP .. step CP of rows r = t / CP .. step THREADS / CP
__global__ void __launch_bounds__(THREADS)
linear_rowscale_kernel(const float* __restrict__ A, int64_t n, int hidden, const float* __restrict__ W, const float* __restrict__ bias,
                       const float* __restrict__ scale, float* __restrict__ out, int64_t ld_out, int C) {
    extern __shared__ float sm[];
    const int hp = hidden + 1;                 // padded row: conflict-free column walks
    float* sW = sm;                            // [C][hp]
    float* sA = sm + (size_t)C * hp;           // [ROWS][hp]
    for (int i = threadIdx.x; i < C * hidden; i += THREADS) sW[(i / hidden) * hp + i % hidden] = __ldg(W + i);
    for (int64_t tile = blockIdx.x; tile * ROWS < n; tile += gridDim.x) {
        const int64_t r0 = tile * ROWS;
        __syncthreads();
        for (int i = threadIdx.x; i < ROWS * hidden; i += THREADS) {
            const int r = i / hidden, h = i % hidden;
            sA[r * hp + h] = (r0 + r < n) ? __ldcs(A + (r0 + r) * hidden + h) : 0.f;
        }
        __syncthreads();
        for (int o = threadIdx.x; o < ROWS * C; o += THREADS) {
            const int r = o / C, c = o % C;
            if (r0 + r >= n) continue;
            const float* a = sA + r * hp;
            const float* w = sW + c * hp;
            float acc = 0.f;
#pragma unroll 8
            for (int h = 0; h < hidden; ++h) acc = fmaf(a[h], w[h], acc);
            if (bias) acc += __ldg(bias + c);
            if (scale) acc *= __ldg(scale + r0 + r);
            out[(r0 + r) * ld_out + c] = acc;
        }
    }
}

// dA tile = s * (dOut tile @ W); per-CTA partial of dW (and dbias) over the CTA's tiles -> part[blockIdx.x]
__global__ void __launch_bounds__(THREADS)
linear_rowscale_bwd_kernel(const float* __restrict__ A, const float* __restrict__ dOut, int64_t ld_dout, int64_t n, int hidden, int C,
                           const float* __restrict__ W, const float* __restrict__ scale, float* __restrict__ dA,
                           float* __restrict__ part /* [grid][C * hidden + C] */) {
    extern __shared__ float sm[];
    const int hp = hidden + 1, cp = C + 1;
    float* sW = sm;                            // [C][hp]
    float* sA = sW + (size_t)C * hp;           // [ROWS][hp]
    float* sG = sA + (size_t)ROWS * hp;        // [ROWS][cp]   s_i * dOut
    for (int i = threadIdx.x; i < C * hidden; i += THREADS) sW[(i / hidden) * hp + i % hidden] = __ldg(W + i);
    // every thread keeps the partial sums of the dW entries it owns across all tiles of this CTA
    constexpr int MAXOWN = 64;                 // C * hidden <= 64 * 256 = 16384 = 64 * THREADS
    float own[MAXOWN];
#pragma unroll
    for (int q = 0; q < MAXOWN; ++q) own[q] = 0.f;
    float ownb = 0.f;                          // dbias entry threadIdx.x (< C)
    const int n_own = (C * hidden + THREADS - 1) / THREADS;
    for (int64_t tile = blockIdx.x; tile * ROWS < n; tile += gridDim.x) {
        const int64_t r0 = tile * ROWS;
        __syncthreads();
        for (int i = threadIdx.x; i < ROWS * hidden; i += THREADS) {
            const int r = i / hidden, h = i % hidden;
            sA[r * hp + h] = (r0 + r < n) ? __ldcs(A + (r0 + r) * hidden + h) : 0.f;
        }
        for (int i = threadIdx.x; i < ROWS * C; i += THREADS) {
            const int r = i / C, c = i % C;
            float g = 0.f;
            if (r0 + r < n) {
                g = __ldcs(dOut + (r0 + r) * ld_dout + c);
                if (scale) g *= __ldg(scale + r0 + r);
            }
            sG[r * cp + c] = g;
        }
        __syncthreads();
        if (dA != nullptr) {
            for (int o = threadIdx.x; o < ROWS * hidden; o += THREADS) {
                const int r = o / hidden, h = o % hidden;
                if (r0 + r >= n) continue;
                float acc = 0.f;
                for (int c = 0; c < C; ++c) acc = fmaf(sG[r * cp + c], sW[c * hp + h], acc);
                dA[(r0 + r) * hidden + h] = acc;
            }
        }
#pragma unroll
        for (int q = 0; q < MAXOWN; ++q) {
            if (q < n_own) {
                const int e = q * THREADS + threadIdx.x;
                if (e < C * hidden) {
                    const int c = e / hidden, h = e % hidden;
                    float acc = own[q];
#pragma unroll 8
                    for (int r = 0; r < ROWS; ++r) acc = fmaf(sG[r * cp + c], sA[r * hp + h], acc);
                    own[q] = acc;
                }
            }
        }
        if (threadIdx.x < C) {
            float acc = ownb;
            for (int r = 0; r < ROWS; ++r) acc += sG[r * cp + threadIdx.x];
            ownb = acc;
        }
    }
    float* mine = part + (size_t)blockIdx.x * (C * hidden + C);
#pragma unroll
    for (int q = 0; q < MAXOWN; ++q) {
        if (q < n_own) {
            const int e = q * THREADS + threadIdx.x;
            if (e < C * hidden) mine[e] = own[q];
        }
    }
    if (threadIdx.x < C) mine[C * hidden + threadIdx.x] = ownb;
}

// dW[e] = sum over CTAs of part[b][e], in CTA order (bit-reproducible)
__global__ void reduce_parts_kernel(const float* __restrict__ part, int n_parts, int len, int split, float* __restrict__ dW, float* __restrict__ dbias) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= len) return;
    float acc = 0.f;
    for (int b = 0; b < n_parts; ++b) acc += part[(size_t)b * len + e];
    if (e < split) dW[e] = acc;
    else if (dbias) dbias[e - split] = acc;
}

int tail_grid(int64_t n) {
    const int64_t tiles = (n + ROWS - 1) / ROWS;
    const int64_t cap = (int64_t)sm_count() * 4;
    return (int)(tiles < cap ? (tiles < 1 ? 1 : tiles) : cap);
}

}  // namespace
}  // namespace ppnp

extern "C" {

int ppnp_linear_rowscale(const float* A, int64_t n, int32_t hidden, const float* W, const float* bias, const float* scale, float* out,
                         int64_t ld_out, int32_t C, void* stream_) {
    using namespace ppnp;
    PPNP_REQUIRE(A && W && out, "null pointer");
    PPNP_REQUIRE(n > 0 && hidden >= 1 && hidden <= 256 && C >= 1 && C <= 64 && ld_out >= C, "need n > 0, hidden <= 256, C <= 64, ld_out >= C");
    const int smem = (C + ROWS) * (hidden + 1) * 4;
    static thread_local int configured = 0;
    if (smem > 48 * 1024 && configured < smem) {
        int rc = check_cuda(cudaFuncSetAttribute(linear_rowscale_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "smem opt-in");
        if (rc) return rc;
        configured = smem;
    }
    linear_rowscale_kernel<<<tail_grid(n), THREADS, smem, as_stream(stream_)>>>(A, n, hidden, W, bias, scale, out, ld_out, C);
    PPNP_CHECK_LAUNCH("linear_rowscale_kernel");
    return PPNP_OK;
}

int64_t ppnp_linear_rowscale_backward_workspace_bytes(int64_t n, int32_t hidden, int32_t C) {
    if (n <= 0 || hidden < 1 || C < 1) return 0;
    return (int64_t)ppnp::tail_grid(n) * ((int64_t)C * hidden + C) * 4;
}

int ppnp_linear_rowscale_backward(const float* A, const float* dOut, int64_t ld_dout, int64_t n, int32_t hidden, int32_t C, const float* W,
                                  const float* scale, float* dA, float* dW, float* dbias, void* workspace, int64_t workspace_bytes,
                                  void* stream_) {
    using namespace ppnp;
    PPNP_REQUIRE(A && dOut && W && dW, "null pointer");
    PPNP_REQUIRE(n > 0 && hidden >= 1 && hidden <= 256 && C >= 1 && C <= 64 && ld_dout >= C, "need n > 0, hidden <= 256, C <= 64, ld_dout >= C");
    PPNP_REQUIRE(workspace && workspace_bytes >= ppnp_linear_rowscale_backward_workspace_bytes(n, hidden, C), "workspace too small");
    const int grid = tail_grid(n);
    const int smem = ((C + ROWS) * (hidden + 1) + ROWS * (C + 1)) * 4;
    static thread_local int configured = 0;
    if (smem > 48 * 1024 && configured < smem) {
        int rc = check_cuda(cudaFuncSetAttribute(linear_rowscale_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "smem opt-in");
        if (rc) return rc;
        configured = smem;
    }
    cudaStream_t stream = as_stream(stream_);
    float* part = reinterpret_cast<float*>(workspace);
    linear_rowscale_bwd_kernel<<<grid, THREADS, smem, stream>>>(A, dOut, ld_dout, n, hidden, C, W, scale, dA, part);
    PPNP_CHECK_LAUNCH("linear_rowscale_bwd_kernel");
    const int len = C * hidden + C;
    reduce_parts_kernel<<<(len + 255) / 256, 256, 0, stream>>>(part, grid, len, C * hidden, dW, dbias);
    PPNP_CHECK_LAUNCH("reduce_parts_kernel");
    return PPNP_OK;
}

}  // extern "C"
