// gather_gemm.cu -- exact PPNP's propagation  out = Pi[idx, :] @ H  in fp32 (SIMT).
//
// Replaces model.py:63 (``self.ppr[idx] @ self.encoder(X)``: an ATen row gather that
// materialises |idx| x n, then a cuBLAS SGEMM) and model.py:65 (``ppr @ ...``) with one kernel
// that reads each needed row of Pi exactly once, straight from its place in the n x n matrix.
// This is the 1e-5 parity path; the bf16 tensor-core path is gather_gemm_tc.cu.
//
// Roofline: HBM-bound for the class counts the reference uses (C = 3..15): 2C/4 flop per byte
// of Pi.  The rows of Pi are streamed once (evict-first); H (n x C, <= a few MB) lives in L2 and
// is staged through shared memory one K-tile at a time so that 8*RPW rows share every load.
#include "common.cuh"

namespace ppnp {
namespace {

constexpr int GG_THREADS = 256;
constexpr int GG_KT = 128;  // k-values per shared-memory tile

template <int CP, int RPW>
__global__ void __launch_bounds__(GG_THREADS)
gather_gemm_f32_kernel(const float* __restrict__ Pi, int64_t ld_pi, const int64_t* __restrict__ idx,
                       int64_t m, int64_t n, const float* __restrict__ H, int64_t ld_h, int C, int c_base,
                       float* __restrict__ out, int64_t ld_out, int64_t k_per_split, int use_atomic) {
    __shared__ float Hs[GG_KT][CP + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row0 = ((int64_t)blockIdx.x * (GG_THREADS / 32) + warp) * RPW;
    const int64_t kbeg = (int64_t)blockIdx.y * k_per_split;
    const int64_t kend = (kbeg + k_per_split < n) ? kbeg + k_per_split : n;

    const float* prow[RPW];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        const int64_t rr = row0 + r;
        prow[r] = nullptr;
        if (rr < m) {
            const int64_t src = idx ? idx[rr] : rr;
            prow[r] = Pi + src * ld_pi;
        }
    }
    float acc[RPW][CP];
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
        for (int c = 0; c < CP; ++c) acc[r][c] = 0.f;

    for (int64_t k0 = kbeg; k0 < kend; k0 += GG_KT) {
        // stage H[k0 : k0+KT, c_base : c_base+CP] (zero padded)
        for (int t = threadIdx.x; t < GG_KT * CP; t += GG_THREADS) {
            const int kk = t / CP, c = t % CP;
            const int64_t k = k0 + kk;
            float h = 0.f;
            if (k < kend && c_base + c < C) h = __ldg(H + k * ld_h + c_base + c);
            Hs[kk][c] = h;
        }
        __syncthreads();
        float p[RPW][GG_KT / 32];
#pragma unroll
        for (int r = 0; r < RPW; ++r)
#pragma unroll
            for (int i = 0; i < GG_KT / 32; ++i) {
                const int64_t k = k0 + lane + 32 * i;
                p[r][i] = (prow[r] != nullptr && k < kend) ? __ldcs(prow[r] + k) : 0.f;
            }
#pragma unroll
        for (int i = 0; i < GG_KT / 32; ++i) {
#pragma unroll
            for (int c = 0; c < CP; ++c) {
                const float h = Hs[lane + 32 * i][c];
#pragma unroll
                for (int r = 0; r < RPW; ++r) acc[r][c] = fmaf(p[r][i], h, acc[r][c]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
#pragma unroll
        for (int c = 0; c < CP; ++c) {
            float v = acc[r][c];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            acc[r][c] = v;
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const int64_t rr = row0 + r;
            if (rr >= m) continue;
#pragma unroll
            for (int c = 0; c < CP; ++c) {
                if (c_base + c < C) {
                    float* o = out + rr * ld_out + c_base + c;
                    if (use_atomic) atomicAdd(o, acc[r][c]); else *o = acc[r][c];
                }
            }
        }
    }
}

// Wide-C variant (16 < C <= 64): lanes across the output columns, GW_R rows of Pi per warp in
// registers-as-accumulators.  The Pi tile of the warp and the H tile of the CTA live in shared memory;
// per 4 k-values a lane issues GW_R broadcast LDS.128 (Pi) + 4*CL LDS.32 (H) for 4*GW_R*CL FMAs, so the
// FP32 pipe, not shared memory, is the limit (the narrow kernel above needs one LDS per FMA).
constexpr int GW_R = 8;    // rows per warp
constexpr int GW_KT = 64;  // k-values per tile

template <int CL>          // columns per lane: 1 (C <= 32) or 2 (C <= 64)
__global__ void __launch_bounds__(GG_THREADS)
gather_gemm_f32_wide_kernel(const float* __restrict__ Pi, int64_t ld_pi, const int64_t* __restrict__ idx,
                            int64_t m, int64_t n, const float* __restrict__ H, int64_t ld_h, int C, int c_base,
                            float* __restrict__ out, int64_t ld_out, int64_t k_per_split, int use_atomic) {
    constexpr int CW = 32 * CL;
    __shared__ __align__(16) float Hs[GW_KT][CW];
    __shared__ __align__(16) float Ps[GG_THREADS / 32][GW_R][GW_KT];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row0 = ((int64_t)blockIdx.x * (GG_THREADS / 32) + warp) * GW_R;
    const int64_t kbeg = (int64_t)blockIdx.y * k_per_split;
    const int64_t kend = (kbeg + k_per_split < n) ? kbeg + k_per_split : n;
    const float* prow[GW_R];
#pragma unroll
    for (int r = 0; r < GW_R; ++r) {
        const int64_t rr = row0 + r;
        prow[r] = (rr < m) ? Pi + (idx ? idx[rr] : rr) * ld_pi : nullptr;
    }
    float acc[GW_R][CL];
#pragma unroll
    for (int r = 0; r < GW_R; ++r)
#pragma unroll
        for (int j = 0; j < CL; ++j) acc[r][j] = 0.f;

    for (int64_t k0 = kbeg; k0 < kend; k0 += GW_KT) {
        for (int t = threadIdx.x; t < GW_KT * CW; t += GG_THREADS) {
            const int kk = t / CW, c = t % CW;
            const int64_t k = k0 + kk;
            Hs[kk][c] = (k < kend && c_base + c < C) ? __ldg(H + k * ld_h + c_base + c) : 0.f;
        }
#pragma unroll
        for (int r = 0; r < GW_R; ++r)
#pragma unroll
            for (int i = 0; i < GW_KT / 32; ++i) {
                const int64_t k = k0 + lane + 32 * i;
                Ps[warp][r][lane + 32 * i] = (prow[r] != nullptr && k < kend) ? __ldcs(prow[r] + k) : 0.f;
            }
        __syncthreads();
#pragma unroll 4
        for (int kk = 0; kk < GW_KT; kk += 4) {
            float h[4][CL];
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int j = 0; j < CL; ++j) h[q][j] = Hs[kk + q][lane + 32 * j];
#pragma unroll
            for (int r = 0; r < GW_R; ++r) {
                const float4 p = *reinterpret_cast<const float4*>(&Ps[warp][r][kk]);
#pragma unroll
                for (int j = 0; j < CL; ++j) {
                    acc[r][j] = fmaf(p.x, h[0][j], acc[r][j]);
                    acc[r][j] = fmaf(p.y, h[1][j], acc[r][j]);
                    acc[r][j] = fmaf(p.z, h[2][j], acc[r][j]);
                    acc[r][j] = fmaf(p.w, h[3][j], acc[r][j]);
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < GW_R; ++r) {
        const int64_t rr = row0 + r;
        if (rr >= m) continue;
#pragma unroll
        for (int j = 0; j < CL; ++j) {
            const int c = c_base + lane + 32 * j;
            if (c < C) {
                float* o = out + rr * ld_out + c;
                if (use_atomic) atomicAdd(o, acc[r][j]); else *o = acc[r][j];
            }
        }
    }
}

// adjoint: out[k, c] = sum_r Pi[idx[r], k] * G[r, c]   (dH of model.py:63 as autograd computes it)
constexpr int GT_RT = 32;
template <int CP>
__global__ void __launch_bounds__(GG_THREADS)
gather_gemm_f32_t_kernel(const float* __restrict__ Pi, int64_t ld_pi, const int64_t* __restrict__ idx,
                         int64_t m, int64_t n, const float* __restrict__ G, int64_t ld_g, int C, int c_base,
                         float* __restrict__ out, int64_t ld_out, int64_t r_per_split, int use_atomic) {
    __shared__ float Gs[GT_RT][CP];
    __shared__ int64_t src[GT_RT];
    const int64_t k = (int64_t)blockIdx.x * GG_THREADS + threadIdx.x;
    const int64_t rbeg = (int64_t)blockIdx.y * r_per_split;
    const int64_t rend = (rbeg + r_per_split < m) ? rbeg + r_per_split : m;
    float acc[CP];
#pragma unroll
    for (int c = 0; c < CP; ++c) acc[c] = 0.f;
    for (int64_t r0 = rbeg; r0 < rend; r0 += GT_RT) {
        for (int t = threadIdx.x; t < GT_RT * CP; t += GG_THREADS) {
            const int rr = t / CP, c = t % CP;
            const int64_t r = r0 + rr;
            Gs[rr][c] = (r < rend && c_base + c < C) ? __ldg(G + r * ld_g + c_base + c) : 0.f;
        }
        if (threadIdx.x < GT_RT) {
            const int64_t r = r0 + threadIdx.x;
            src[threadIdx.x] = (r < rend) ? (idx ? idx[r] : r) : -1;
        }
        __syncthreads();
        if (k < n) {
            float p[GT_RT];
#pragma unroll
            for (int rr = 0; rr < GT_RT; ++rr) p[rr] = (src[rr] >= 0) ? __ldcs(Pi + src[rr] * ld_pi + k) : 0.f;
#pragma unroll
            for (int rr = 0; rr < GT_RT; ++rr)
#pragma unroll
                for (int c = 0; c < CP; ++c) acc[c] = fmaf(p[rr], Gs[rr][c], acc[c]);
        }
        __syncthreads();
    }
    if (k < n) {
#pragma unroll
        for (int c = 0; c < CP; ++c) {
            if (c_base + c < C) {
                float* o = out + k * ld_out + c_base + c;
                if (use_atomic) atomicAdd(o, acc[c]); else *o = acc[c];
            }
        }
    }
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, int64_t count) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const uint32_t u = __float_as_uint(__ldcs(src + i));
        uint32_t r;
        if ((u & 0x7f800000u) == 0x7f800000u) r = u >> 16 | ((u & 0xffffu) ? 0x40u : 0u);  // inf / nan
        else r = (u + 0x7fffu + ((u >> 16) & 1u)) >> 16;                                   // round to nearest even
        dst[i] = (uint16_t)r;
    }
}

template <int CP, int RPW>
int launch_fwd(const float* Pi, int64_t ld_pi, const int64_t* idx, int64_t m, int64_t n, const float* H,
               int64_t ld_h, int C, int c_base, float* out, int64_t ld_out, cudaStream_t stream) {
    const int64_t rows_per_cta = (GG_THREADS / 32) * RPW;
    const int64_t gx = (m + rows_per_cta - 1) / rows_per_cta;
    const int64_t ktiles = (n + GG_KT - 1) / GG_KT;
    int64_t want = (2 * (int64_t)sm_count() + gx - 1) / gx;  // K splits so that the grid fills the GPU twice
    if (want < 1) want = 1;
    if (want > ktiles) want = ktiles;
    const int64_t k_per_split = ((ktiles + want - 1) / want) * GG_KT;
    const int64_t gy = (n + k_per_split - 1) / k_per_split;
    const int use_atomic = gy > 1;
    if (use_atomic && c_base == 0) {
        int rc = check_cuda(cudaMemset2DAsync(out, ld_out * sizeof(float), 0, (size_t)C * sizeof(float), (size_t)m, stream), "memset out");
        if (rc) return rc;
    }
    gather_gemm_f32_kernel<CP, RPW><<<dim3((unsigned)gx, (unsigned)gy), GG_THREADS, 0, stream>>>(
        Pi, ld_pi, idx, m, n, H, ld_h, C, c_base, out, ld_out, k_per_split, use_atomic);
    PPNP_CHECK_LAUNCH("gather_gemm_f32_kernel");
    return PPNP_OK;
}

template <int CL>
int launch_fwd_wide(const float* Pi, int64_t ld_pi, const int64_t* idx, int64_t m, int64_t n, const float* H,
                    int64_t ld_h, int C, int c_base, float* out, int64_t ld_out, cudaStream_t stream) {
    const int64_t rows_per_cta = (GG_THREADS / 32) * GW_R;
    const int64_t gx = (m + rows_per_cta - 1) / rows_per_cta;
    const int64_t ktiles = (n + GW_KT - 1) / GW_KT;
    int64_t want = (2 * (int64_t)sm_count() + gx - 1) / gx;
    if (want < 1) want = 1;
    if (want > ktiles) want = ktiles;
    const int64_t k_per_split = ((ktiles + want - 1) / want) * GW_KT;
    const int64_t gy = (n + k_per_split - 1) / k_per_split;
    const int use_atomic = gy > 1;
    if (use_atomic && c_base == 0) {
        int rc = check_cuda(cudaMemset2DAsync(out, ld_out * sizeof(float), 0, (size_t)C * sizeof(float), (size_t)m, stream), "memset out");
        if (rc) return rc;
    }
    gather_gemm_f32_wide_kernel<CL><<<dim3((unsigned)gx, (unsigned)gy), GG_THREADS, 0, stream>>>(
        Pi, ld_pi, idx, m, n, H, ld_h, C, c_base, out, ld_out, k_per_split, use_atomic);
    PPNP_CHECK_LAUNCH("gather_gemm_f32_wide_kernel");
    return PPNP_OK;
}

template <int CP>
int launch_t(const float* Pi, int64_t ld_pi, const int64_t* idx, int64_t m, int64_t n, const float* G,
             int64_t ld_g, int C, int c_base, float* out, int64_t ld_out, cudaStream_t stream) {
    const int64_t gx = (n + GG_THREADS - 1) / GG_THREADS;
    const int64_t rtiles = (m + GT_RT - 1) / GT_RT;
    int64_t want = (2 * (int64_t)sm_count() + gx - 1) / gx;
    if (want < 1) want = 1;
    if (want > rtiles) want = rtiles;
    const int64_t r_per_split = ((rtiles + want - 1) / want) * GT_RT;
    const int64_t gy = (m + r_per_split - 1) / r_per_split;
    const int use_atomic = gy > 1;
    if (use_atomic && c_base == 0) {
        int rc = check_cuda(cudaMemset2DAsync(out, ld_out * sizeof(float), 0, (size_t)C * sizeof(float), (size_t)n, stream), "memset out");
        if (rc) return rc;
    }
    gather_gemm_f32_t_kernel<CP><<<dim3((unsigned)gx, (unsigned)gy), GG_THREADS, 0, stream>>>(
        Pi, ld_pi, idx, m, n, G, ld_g, C, c_base, out, ld_out, r_per_split, use_atomic);
    PPNP_CHECK_LAUNCH("gather_gemm_f32_t_kernel");
    return PPNP_OK;
}

}  // namespace
}  // namespace ppnp

extern "C" {

int ppnp_gather_gemm_f32(const float* Pi, int64_t ld_pi, const int64_t* idx, int64_t m, int64_t n,
                         const float* H, int64_t ld_h, int32_t C, float* out, int64_t ld_out,
                         int32_t transpose, void* stream_) {
    using namespace ppnp;
    PPNP_REQUIRE(Pi && H && out, "null pointer");
    PPNP_REQUIRE(m > 0 && n > 0 && C > 0, "m, n, C > 0");
    PPNP_REQUIRE(ld_pi >= n && ld_h >= C && ld_out >= C, "leading dimensions too small");
    cudaStream_t stream = as_stream(stream_);
    for (int c_base = 0; c_base < C; c_base += 64) {
        const int cw = (C - c_base < 64) ? C - c_base : 64;
        int rc;
        if (!transpose) {
            if (cw <= 4) rc = launch_fwd<4, 4>(Pi, ld_pi, idx, m, n, H, ld_h, C, c_base, out, ld_out, stream);
            else if (cw <= 8) rc = launch_fwd<8, 4>(Pi, ld_pi, idx, m, n, H, ld_h, C, c_base, out, ld_out, stream);
            else if (cw <= 16) rc = launch_fwd<16, 4>(Pi, ld_pi, idx, m, n, H, ld_h, C, c_base, out, ld_out, stream);
            else if (cw <= 32) rc = launch_fwd_wide<1>(Pi, ld_pi, idx, m, n, H, ld_h, C, c_base, out, ld_out, stream);
            else rc = launch_fwd_wide<2>(Pi, ld_pi, idx, m, n, H, ld_h, C, c_base, out, ld_out, stream);
        } else {
            if (cw <= 4) rc = launch_t<4>(Pi, ld_pi, idx, m, n, H, ld_h, C, c_base, out, ld_out, stream);
            else if (cw <= 8) rc = launch_t<8>(Pi, ld_pi, idx, m, n, H, ld_h, C, c_base, out, ld_out, stream);
            else if (cw <= 16) rc = launch_t<16>(Pi, ld_pi, idx, m, n, H, ld_h, C, c_base, out, ld_out, stream);
            else if (cw <= 32) rc = launch_t<32>(Pi, ld_pi, idx, m, n, H, ld_h, C, c_base, out, ld_out, stream);
            else rc = launch_t<64>(Pi, ld_pi, idx, m, n, H, ld_h, C, c_base, out, ld_out, stream);
        }
        if (rc) return rc;
    }
    return PPNP_OK;
}

int ppnp_f32_to_bf16(const float* src, void* dst, int64_t count, void* stream_) {
    using namespace ppnp;
    PPNP_REQUIRE(src && dst && count >= 0, "bad arguments");
    if (count == 0) return PPNP_OK;
    int64_t blocks = (count + 1023) / 1024;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    f32_to_bf16_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream_)>>>(src, reinterpret_cast<uint16_t*>(dst), count);
    PPNP_CHECK_LAUNCH("f32_to_bf16_kernel");
    return PPNP_OK;
}

}  // extern "C"
