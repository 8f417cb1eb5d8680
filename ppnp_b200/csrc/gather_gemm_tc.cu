// gather_gemm_tc.cu -- exact PPNP's dense contraction  out = Pi[idx, :] @ H  on the 5th-generation
// tensor cores (tcgen05.mma, bf16 x bf16 -> fp32 accumulators in TMEM), sm_100a only.
//
// Replaces model.py:63 / model.py:65 (ATen row gather + cuBLAS SGEMM) for the bf16 mode of
// BASELINE.json config 2.  The contraction is HBM-bound for every class count the reference uses
// (N = C <= 64: 2N/2 flop per byte of Pi against a ridge of ~200 flop/B), so the design goal is to
// stream the gathered rows of Pi at HBM speed and keep the tensor pipe off the critical path:
//   * one CTA per (128-row tile of idx, K-slice); K-slices give every SM work when |idx| is small
//     (main.py gathers 60..940 rows) and are reduced in a fixed order afterwards (deterministic);
//   * warps 0-3 are producers: the gather happens in the load itself -- 16-byte cp.async copies
//     from row idx[r] of Pi straight into the 128-byte-swizzled K-major shared-memory tile that
//     the UMMA descriptor expects (no |idx| x n intermediate, unlike the reference); 5 stages of
//     16 KB + B tile are in flight per SM;
//   * warp 4 issues tcgen05.mma (M=128, N=C padded to 16, K=16 per instruction) from one thread and
//     frees smem stages / publishes the accumulator through tcgen05.commit -> mbarrier;
//   * warps 0-3 then read the accumulator with tcgen05.ld (32 lanes x 32 bit, their own TMEM
//     quarter) and write fp32 rows.
// H is converted once per call to a K-major bf16 tile source Bt[N_pad][K_pad] (prep kernel).
#include "common.cuh"

namespace ppnp {
namespace {

constexpr int TC_BM = 128;          // rows per tile (UMMA M)
constexpr int TC_BK = 64;           // bf16 per k-block row = 128 B = one swizzle span
constexpr int TC_STAGES = 6;
constexpr int TC_LAG = 4;           // cp.async groups a producer keeps in flight behind the newest
constexpr int TC_PRODUCERS = 128;   // warps 0..3
constexpr int TC_THREADS = 160;     // + warp 4 (MMA issue, TMEM alloc)
constexpr int TC_A_STAGE = TC_BM * TC_BK * 2;  // 16384 B

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, 128-byte swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start address >> 4 [0,14), LBO >> 4 [16,30) (= 1 for swizzled K-major), SBO >> 4 [32,46) (8 rows x
// 128 B = 1024 B between row groups), version = 1 [46,48), layout SWIZZLE_128B = 2 [61,64).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 [4,6) = 1, A = B = BF16
// [7,10) / [10,13) = 1, both K-major (bits 15, 16 = 0), N >> 3 at [17,23), M >> 4 at [24,29).
__device__ __forceinline__ uint32_t umma_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// H[n x C] fp32 -> Bt[N_pad][K_pad] bf16, K-major, zero padded.
__global__ void tc_prep_b_kernel(const float* __restrict__ H, int64_t ld_h, int64_t n, int C, int c_base, int n_pad,
                                 int64_t k_pad, uint16_t* __restrict__ Bt) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= k_pad) return;
    for (int c = 0; c < n_pad; ++c) {
        float v = 0.f;
        if (k < n && c_base + c < C) v = __ldg(H + k * ld_h + c_base + c);
        const uint32_t u = __float_as_uint(v);
        Bt[(int64_t)c * k_pad + k] = (uint16_t)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
    }
}

// out[r, c] = sum over K-slices, in slice order
__global__ void tc_reduce_kernel(const float* __restrict__ ws, int splits, int64_t m, int64_t m_pad, int n_pad, int C,
                                 int c_base, float* __restrict__ out, int64_t ld_out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int cw = (C - c_base < n_pad) ? C - c_base : n_pad;
    if (t >= m * cw) return;
    const int64_t r = t / cw;
    const int c = (int)(t % cw);
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += ws[((int64_t)s * m_pad + r) * n_pad + c];
    out[r * ld_out + c_base + c] = acc;
}

template <int N_PAD>
__global__ void __launch_bounds__(TC_THREADS, 1)
gather_gemm_tc_kernel(const uint16_t* __restrict__ Pi, int64_t ld_pi, const int64_t* __restrict__ idx, int64_t m,
                      int64_t n, const uint16_t* __restrict__ Bt, int64_t k_pad, int kb_per_split, int kb_total,
                      float* __restrict__ dst, int64_t dst_ld, int64_t dst_split_stride, int C, int c_base) {
    constexpr int B_STAGE = N_PAD * TC_BK * 2;
    constexpr int TMEM_COLS = (N_PAD <= 32) ? 32 : 64;  // power of two >= 32
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;            // 1024 B alignment for the swizzle atoms
    const uint32_t sA = base;
    const uint32_t sB = sA + TC_STAGES * TC_A_STAGE;
    const uint32_t sBar = sB + TC_STAGES * B_STAGE;          // full[STAGES], empty[STAGES], accum
    const uint32_t sTmem = sBar + (2 * TC_STAGES + 1) * 8;
    uint8_t* gen_base = smem_raw + (base - raw);
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen_base + (sTmem - base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row0 = (int64_t)blockIdx.x * TC_BM;
    const int kb0 = (int)blockIdx.y * kb_per_split;
    int nkb = kb_total - kb0;
    if (nkb > kb_per_split) nkb = kb_per_split;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(sBar + 8 * s, TC_PRODUCERS);               // full: every producer thread arrives
            mbar_init(sBar + 8 * (TC_STAGES + s), 1);            // empty: one tcgen05.commit
        }
        mbar_init(sBar + 8 * (2 * TC_STAGES), 1);                // accumulator ready
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sTmem), "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        // ------------------------------------------------------------------ producers
        const int t = threadIdx.x;            // 0..127
        const int ch = t & 7;                 // 16-byte chunk of the 128-byte k-block row
        const uint16_t* arow[8];
        uint32_t aoff[8];
        bool avalid[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int R = j * 16 + (t >> 3);
            const int64_t gr = row0 + R;
            avalid[j] = gr < m;
            const int64_t src = avalid[j] ? (idx ? idx[gr] : gr) : 0;
            arow[j] = Pi + src * ld_pi + ch * 8;
            aoff[j] = (uint32_t)((R >> 3) * 1024 + (R & 7) * 128 + ((ch ^ (R & 7)) << 4));
        }
        constexpr int B_CHUNKS = N_PAD * 8;                       // 16-byte chunks of the B tile
        constexpr int B_PER_THREAD = (B_CHUNKS + TC_PRODUCERS - 1) / TC_PRODUCERS;

        for (int i = 0; i < nkb; ++i) {
            const int s = i % TC_STAGES;
            mbar_wait(sBar + 8 * (TC_STAGES + s), (((uint32_t)(i / TC_STAGES)) & 1u) ^ 1u);
            const int64_t k = (int64_t)(kb0 + i) * TC_BK + ch * 8;  // first element of my chunk
            int64_t left = n - k;                                   // valid elements from k on
            const uint32_t kbytes = left >= 8 ? 16u : (left > 0 ? (uint32_t)left * 2u : 0u);
            const uint32_t a_stage = sA + s * TC_A_STAGE;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                cp_async16(a_stage + aoff[j], arow[j] + (int64_t)(kb0 + i) * TC_BK, avalid[j] ? kbytes : 0u);
            const uint32_t b_stage = sB + s * B_STAGE;
#pragma unroll
            for (int q = 0; q < B_PER_THREAD; ++q) {
                const int cid = q * TC_PRODUCERS + t;
                if (cid < B_CHUNKS) {
                    const int c = cid >> 3, bc = cid & 7;
                    const uint32_t off = (uint32_t)((c >> 3) * 1024 + (c & 7) * 128 + ((bc ^ (c & 7)) << 4));
                    cp_async16(b_stage + off, Bt + (int64_t)c * k_pad + (int64_t)(kb0 + i) * TC_BK + bc * 8, 16u);
                }
            }
            cp_async_commit();
            if (i >= TC_LAG) {
                cp_async_wait<TC_LAG>();
                fence_proxy_async();
                mbar_arrive(sBar + 8 * ((i - TC_LAG) % TC_STAGES));
            }
        }
        cp_async_wait<0>();
        fence_proxy_async();
        for (int i = (nkb > TC_LAG ? nkb - TC_LAG : 0); i < nkb; ++i) mbar_arrive(sBar + 8 * (i % TC_STAGES));

        // ------------------------------------------------------------------ epilogue (same warps)
        mbar_wait(sBar + 8 * (2 * TC_STAGES), 0);
        tc_fence_after();
        const int R = warp * 32 + lane;
        const int64_t gr = row0 + R;
        float* drow = dst + (int64_t)blockIdx.y * dst_split_stride + gr * dst_ld;
#pragma unroll
        for (int c0 = 0; c0 < N_PAD; c0 += 16) {
            uint32_t r[16];
            tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
            tmem_ld_wait();
            if (gr < m) {
#pragma unroll
                for (int c = 0; c < 16; ++c)
                    if (c_base + c0 + c < C) drow[c0 + c] = __uint_as_float(r[c]);
            }
        }
        tc_fence_before();
    } else {
        // ------------------------------------------------------------------ MMA issuer (warp 4)
        const uint32_t idesc = umma_idesc_bf16(TC_BM, N_PAD);
        for (int i = 0; i < nkb; ++i) {
            const int s = i % TC_STAGES;
            mbar_wait(sBar + 8 * s, ((uint32_t)(i / TC_STAGES)) & 1u);
            tc_fence_after();
            fence_proxy_async();
            if (lane == 0) {
                const uint64_t adesc = umma_desc_sw128(sA + s * TC_A_STAGE);
                const uint64_t bdesc = umma_desc_sw128(sB + s * B_STAGE);
#pragma unroll
                for (int k = 0; k < TC_BK / 16; ++k)   // +32 B per UMMA_K = +2 in the encoded start address
                    umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (i > 0 || k > 0) ? 1u : 0u);
                umma_commit(sBar + 8 * (TC_STAGES + s));
                if (i == nkb - 1) umma_commit(sBar + 8 * (2 * TC_STAGES));
            }
            __syncwarp();
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

inline int64_t align256(int64_t x) { return (x + 255) & ~(int64_t)255; }

struct TcShape {
    int n_pad;       // output columns per launch, multiple of 16, <= 64
    int kb_total;    // 64-wide k-blocks
    int64_t k_pad;
    int64_t m_tiles, m_pad;
    int splits, kb_per_split;
};

inline TcShape tc_shape(int64_t m, int64_t n, int C) {
    TcShape s;
    const int cw = C < 64 ? C : 64;
    s.n_pad = ((cw + 15) / 16) * 16;
    s.kb_total = (int)((n + TC_BK - 1) / TC_BK);
    s.k_pad = (int64_t)s.kb_total * TC_BK;
    s.m_tiles = (m + TC_BM - 1) / TC_BM;
    s.m_pad = s.m_tiles * TC_BM;
    // One CTA per SM is resident (6 stages of smem), so the grid runs in waves of sm_count CTAs: pick the
    // number of K-slices that wastes the least of the last wave (155 row tiles x 2 slices = 2.09 waves
    // ran at 70 %; x 19 slices = 19.9 waves).  At least 8 k-blocks per slice keep the pipeline busy.
    const int64_t sms = sm_count();
    int64_t max_splits = s.kb_total / 8;
    if (max_splits < 1) max_splits = 1;
    if (max_splits > 64) max_splits = 64;
    int64_t want = 1;
    double best = -1.0;
    for (int64_t c = 1; c <= max_splits; ++c) {
        const int64_t per = (s.kb_total + c - 1) / c;
        const int64_t real = (s.kb_total + per - 1) / per;          // slices that actually exist
        const int64_t ctas = s.m_tiles * real;
        const int64_t waves = (ctas + sms - 1) / sms;
        // useful fraction of the waves, minus a charge per slice for the partial sums (written by the
        // epilogue, re-read by the reduce kernel): negligible at N_pad = 16, 0.13 ms of 0.21 at N_pad = 64
        const double w = (double)s.n_pad / 16.0;
        const double eff = (double)ctas / (double)(waves * sms) - 0.002 * w * w * (double)real;
        if (eff > best + 1e-9) { best = eff; want = c; }
    }
    s.kb_per_split = (int)((s.kb_total + want - 1) / want);
    s.splits = (s.kb_total + s.kb_per_split - 1) / s.kb_per_split;
    return s;
}

template <int N_PAD>
int launch_tc(const TcShape& s, const uint16_t* Pi, int64_t ld_pi, const int64_t* idx, int64_t m, int64_t n,
              const uint16_t* Bt, float* dst, int64_t dst_ld, int64_t dst_split_stride, int C, int c_base,
              cudaStream_t stream) {
    constexpr int B_STAGE = N_PAD * TC_BK * 2;
    const int smem = 1024 + TC_STAGES * (TC_A_STAGE + B_STAGE) + (2 * TC_STAGES + 1) * 8 + 16;
    auto k = gather_gemm_tc_kernel<N_PAD>;
    int rc = check_cuda(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "smem attribute");
    if (rc) return rc;
    dim3 grid((unsigned)s.m_tiles, (unsigned)s.splits);
    k<<<grid, TC_THREADS, smem, stream>>>(Pi, ld_pi, idx, m, n, Bt, s.k_pad, s.kb_per_split, s.kb_total, dst, dst_ld,
                                          dst_split_stride, C, c_base);
    PPNP_CHECK_LAUNCH("gather_gemm_tc_kernel");
    return PPNP_OK;
}

}  // namespace
}  // namespace ppnp

extern "C" {

int64_t ppnp_gather_gemm_bf16_workspace_bytes(int64_t m, int64_t n, int32_t C) {
    using namespace ppnp;
    if (m <= 0 || n <= 0 || C <= 0) return 256;
    const TcShape s = tc_shape(m, n, C);
    return align256((int64_t)s.n_pad * s.k_pad * 2) + align256((int64_t)s.splits * s.m_pad * s.n_pad * 4) + 256;
}

int ppnp_gather_gemm_bf16(const void* Pi_, int64_t ld_pi, const int64_t* idx, int64_t m, int64_t n, const float* H,
                          int64_t ld_h, int32_t C, float* out, int64_t ld_out, void* workspace,
                          int64_t workspace_bytes, void* stream_) {
    using namespace ppnp;
    PPNP_REQUIRE(Pi_ && H && out && workspace, "null pointer");
    PPNP_REQUIRE(m > 0 && n > 0 && C > 0, "m, n, C > 0");
    PPNP_REQUIRE(ld_pi >= n && ld_pi % 8 == 0, "bf16 Pi needs a leading dimension that is a multiple of 8 (16-byte rows)");
    PPNP_REQUIRE((reinterpret_cast<uintptr_t>(Pi_) & 15u) == 0, "Pi must be 16-byte aligned");
    PPNP_REQUIRE(ld_h >= C && ld_out >= C, "leading dimensions too small");
    PPNP_REQUIRE(workspace_bytes >= ppnp_gather_gemm_bf16_workspace_bytes(m, n, C), "workspace too small");
    cudaStream_t stream = as_stream(stream_);
    const uint16_t* Pi = reinterpret_cast<const uint16_t*>(Pi_);
    const TcShape s = tc_shape(m, n, C);
    uint16_t* Bt = reinterpret_cast<uint16_t*>(workspace);
    float* ws = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + align256((int64_t)s.n_pad * s.k_pad * 2));

    for (int c_base = 0; c_base < C; c_base += 64) {
        const int cw = (C - c_base < 64) ? C - c_base : 64;
        const int n_pad = ((cw + 15) / 16) * 16;
        tc_prep_b_kernel<<<(unsigned)((s.k_pad + 255) / 256), 256, 0, stream>>>(H, ld_h, n, C, c_base, n_pad, s.k_pad, Bt);
        PPNP_CHECK_LAUNCH("tc_prep_b_kernel");
        float* dst;
        int64_t dst_ld, dst_split;
        if (s.splits == 1) { dst = out + c_base; dst_ld = ld_out; dst_split = 0; }
        else { dst = ws; dst_ld = n_pad; dst_split = s.m_pad * (int64_t)n_pad; }
        // C bound for the epilogue: direct stores use the caller's C, workspace stores write all n_pad columns
        const int c_lim = (s.splits == 1) ? C : c_base + n_pad;
        int rc;
        switch (n_pad) {
            case 16: rc = launch_tc<16>(s, Pi, ld_pi, idx, m, n, Bt, dst, dst_ld, dst_split, c_lim, c_base, stream); break;
            case 32: rc = launch_tc<32>(s, Pi, ld_pi, idx, m, n, Bt, dst, dst_ld, dst_split, c_lim, c_base, stream); break;
            case 48: rc = launch_tc<48>(s, Pi, ld_pi, idx, m, n, Bt, dst, dst_ld, dst_split, c_lim, c_base, stream); break;
            default: rc = launch_tc<64>(s, Pi, ld_pi, idx, m, n, Bt, dst, dst_ld, dst_split, c_lim, c_base, stream); break;
        }
        if (rc) return rc;
        if (s.splits > 1) {
            const int64_t total = m * cw;
            tc_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(ws, s.splits, m, s.m_pad, n_pad, C, c_base, out, ld_out);
            PPNP_CHECK_LAUNCH("tc_reduce_kernel");
        }
    }
    return PPNP_OK;
}

}  // extern "C"
