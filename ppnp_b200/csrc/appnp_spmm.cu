// appnp_spmm.cu -- the APPNP propagation step, fused SpMM + teleport axpy, for sm_100a.
//
//   Z_{k+1} = (1 - alpha) * A_hat @ Z_k + alpha * H            (BASELINE.json north_star)
//
// HBM-bound gather/stream work: the kernel never forms a GEMM.  The
// adjacency arrives as an edge stream cut into fixed-size chunks (ppnp_b200/plan.py):
//   * a group of G lanes owns VEC * G consecutive features of a row (G * VEC * 4 B = the bytes
//     one gathered Z row contributes; F = 64 -> 16 lanes x float4 = one 256 B request),
//   * every group walks one chunk of `chunk_edges` edges: equal work per group whatever the
//     degree distribution (nnz-split load balancing; a hub row of 86 k edges is just 337
//     chunks running on 337 different groups),
//   * the last edge of a segment carries PPNP_FLAG; the group then either finishes the row
//     (epilogue a * acc + b * teleport, streamed store) or writes a partial sum that
//     fixup_kernel adds up in slot order (deterministic, no atomics),
//   * column indices, teleport rows and outputs are streamed with evict-first hints so that
//     the gathered Z rows (re-used across hub neighbourhoods) keep the 126 MB L2.
#include "common.cuh"

namespace ppnp {
namespace {

constexpr unsigned FULL = 0xffffffffu;

template <int G>
struct GroupCfg {
    // indices loaded per lane and batch: keep the per-group request >= 64 B (two full sectors)
    static constexpr int IPL = (G >= 16) ? 1 : (G == 8 ? 2 : 4);
    static constexpr int EB = G * IPL;  // edges per batch
};

template <int IPL>
struct IdxLoad;
template <>
struct IdxLoad<1> {
    __device__ __forceinline__ static void load(const int32_t* p, int (&r)[1]) { r[0] = __ldcs(p); }
    __device__ __forceinline__ static void loadf(const float* p, float (&r)[1]) { r[0] = __ldcs(p); }
};
template <>
struct IdxLoad<2> {
    __device__ __forceinline__ static void load(const int32_t* p, int (&r)[2]) {
        const int2 v = __ldcs(reinterpret_cast<const int2*>(p)); r[0] = v.x; r[1] = v.y;
    }
    __device__ __forceinline__ static void loadf(const float* p, float (&r)[2]) {
        const float2 v = __ldcs(reinterpret_cast<const float2*>(p)); r[0] = v.x; r[1] = v.y;
    }
};
template <>
struct IdxLoad<4> {
    __device__ __forceinline__ static void load(const int32_t* p, int (&r)[4]) {
        const int4 v = __ldcs(reinterpret_cast<const int4*>(p)); r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
    }
    __device__ __forceinline__ static void loadf(const float* p, float (&r)[4]) {
        const float4 v = __ldcs(reinterpret_cast<const float4*>(p)); r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
    }
};

template <int VEC, int G, bool HAS_VAL, int U>
__global__ void __launch_bounds__(256)
spmm_stream_kernel(const int32_t* __restrict__ cols, const float* __restrict__ vals,
                   const int32_t* __restrict__ seg_row, const int32_t* __restrict__ chunk_seg,
                   int64_t n_chunks, int chunk_edges,
                   const float* __restrict__ Zin, const float* __restrict__ T,
                   float* __restrict__ Zout, float* __restrict__ partial,
                   int64_t ld, int F, float alpha, int epi) {
    using V = Vec<VEC>;
    constexpr int IPL = GroupCfg<G>::IPL;
    constexpr int EB = GroupCfg<G>::EB;
    constexpr int GPW = 32 / G;  // groups per warp
    static_assert(EB % U == 0, "sub-batch must divide the batch");

    const int lane = threadIdx.x & 31;
    const int g = lane / G;
    const int lg = lane % G;
    const unsigned gshift = (unsigned)(g * G);
    const unsigned gmask = (G == 32) ? FULL : ((1u << G) - 1u);
    const unsigned lt = (1u << lg) - 1u;  // lower lanes of my group

    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t total_groups = (((int64_t)gridDim.x * blockDim.x) >> 5) * GPW;

    const int f = ((int)blockIdx.y * G + lg) * VEC;
    const bool active = f < F;

    for (int64_t c = warp_global * GPW + g; c < n_chunks; c += total_groups) {
        int s = __ldg(chunk_seg + c);
        const int64_t ebase = c * (int64_t)chunk_edges;
        V acc; acc.zero();
        float cnt = 0.f;

        for (int b = 0; b < chunk_edges; b += EB) {
            int raw[IPL];
            float w[IPL];
            IdxLoad<IPL>::load(cols + ebase + b + lg * IPL, raw);
            if (HAS_VAL) IdxLoad<IPL>::loadf(vals + ebase + b + lg * IPL, w);

            // position of every segment end of this batch in seg_row (edge order: lane-major)
            int pre = 0, tot = 0;
#pragma unroll
            for (int k = 0; k < IPL; ++k) {
                const unsigned m = (__ballot_sync(FULL, raw[k] < 0) >> gshift) & gmask;
                pre += __popc(m & lt);
                tot += __popc(m);
            }
            int segv[IPL];
#pragma unroll
            for (int k = 0; k < IPL; ++k) {
                segv[k] = 0;
                if (raw[k] < 0) { segv[k] = __ldcs(seg_row + s + pre); ++pre; }
            }
            s += tot;

#pragma unroll
            for (int u0 = 0; u0 < EB; u0 += U) {
                V v[U], t[U];
                int ru[U], sv[U];
                float wu[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int e = u0 + u;
                    ru[u] = __shfl_sync(FULL, raw[e % IPL], e / IPL, G);
                    sv[u] = __shfl_sync(FULL, segv[e % IPL], e / IPL, G);
                    if (HAS_VAL) wu[u] = __shfl_sync(FULL, w[e % IPL], e / IPL, G);
                    const int col = ru[u] & 0x7fffffff;
                    v[u].zero();
                    t[u].zero();
                    if (active && col != PPNP_NULL_COL) v[u] = V::load(Zin + (int64_t)col * ld + f);
                    if (active && ru[u] < 0 && sv[u] >= 0) t[u] = V::load_stream(T + (int64_t)sv[u] * ld + f);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (HAS_VAL) acc.fma(wu[u], v[u]); else acc.add(v[u]);
                    cnt += ((ru[u] & 0x7fffffff) != PPNP_NULL_COL) ? 1.f : 0.f;
                    if (ru[u] < 0) {
                        if (sv[u] < 0) {
                            const int64_t slot = sv[u] & 0x7fffffff;
                            if (active) acc.store(partial + slot * ld + f);
                        } else {
                            float a, bb;
                            epi_coef(epi, alpha, cnt, a, bb);
                            const V o = V::axpby(a, acc, bb, t[u]);
                            if (active) o.store_stream(Zout + (int64_t)sv[u] * ld + f);
                        }
                        acc.zero();
                        cnt = 0.f;
                    }
                }
            }
        }
    }
}

// Rows split over several segments: add the partial sums in slot order, then the epilogue.
template <int VEC, int G>
__global__ void __launch_bounds__(256)
fixup_kernel(const int32_t* __restrict__ fix_ptr, const int32_t* __restrict__ fix_row,
             const float* __restrict__ fix_deg, int64_t n_fix, const float* __restrict__ partial,
             const float* __restrict__ T, float* __restrict__ Zout, int64_t ld, int F, float alpha, int epi) {
    using V = Vec<VEC>;
    constexpr int GPW = 32 / G;
    constexpr int U = 8;
    const int lane = threadIdx.x & 31;
    const int g = lane / G;
    const int lg = lane % G;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t total_groups = (((int64_t)gridDim.x * blockDim.x) >> 5) * GPW;
    const int f = ((int)blockIdx.y * G + lg) * VEC;
    if (f >= F) return;
    for (int64_t q = warp_global * GPW + g; q < n_fix; q += total_groups) {
        const int s0 = __ldg(fix_ptr + q), s1 = __ldg(fix_ptr + q + 1);
        const int row = __ldg(fix_row + q);
        const V t = V::load_stream(T + (int64_t)row * ld + f);
        V acc; acc.zero();
        int s = s0;
        for (; s + U <= s1; s += U) {
            V p[U];
#pragma unroll
            for (int u = 0; u < U; ++u) p[u] = V::load_plain(partial + (int64_t)(s + u) * ld + f);
#pragma unroll
            for (int u = 0; u < U; ++u) acc.add(p[u]);
        }
        for (; s < s1; ++s) acc.add(V::load_plain(partial + (int64_t)s * ld + f));
        float a, bb;
        epi_coef(epi, alpha, __ldg(fix_deg + q), a, bb);
        V::axpby(a, acc, bb, t).store_stream(Zout + (int64_t)row * ld + f);
    }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline int pow2ceil(int x) { int p = 1; while (p < x) p <<= 1; return p; }

template <typename K>
int blocks_per_sm(K kernel, int threads) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, 0) != cudaSuccess || nb < 1) nb = 1;
    return nb;
}

template <int VEC, int G>
int launch_step(const ppnp_plan_t* p, const float* Zin, const float* T, float* Zout, float* partial,
                int64_t ld, int F, float alpha, int epi, bool use_vals, cudaStream_t stream) {
    constexpr int THREADS = 256;
    constexpr int U = (VEC == 4) ? 4 : ((GroupCfg<G>::EB >= 8) ? 8 : GroupCfg<G>::EB);
    constexpr int GPW = 32 / G;
    const int tiles = (F + G * VEC - 1) / (G * VEC);
    const int64_t groups_per_block = (THREADS / 32) * GPW;
    const int64_t need = (p->n_chunks + groups_per_block - 1) / groups_per_block;
    if (use_vals) {
        auto k = spmm_stream_kernel<VEC, G, true, U>;
        static thread_local int occ = 0;
        if (!occ) occ = blocks_per_sm(k, THREADS);
        const int64_t cap = (int64_t)sm_count() * occ;
        dim3 grid((unsigned)(need < cap ? need : cap), (unsigned)tiles);
        k<<<grid, THREADS, 0, stream>>>(p->cols, p->vals, p->seg_row, p->chunk_seg, p->n_chunks, p->chunk_edges,
                                        Zin, T, Zout, partial, ld, F, alpha, epi);
    } else {
        auto k = spmm_stream_kernel<VEC, G, false, U>;
        static thread_local int occ = 0;
        if (!occ) occ = blocks_per_sm(k, THREADS);
        const int64_t cap = (int64_t)sm_count() * occ;
        dim3 grid((unsigned)(need < cap ? need : cap), (unsigned)tiles);
        k<<<grid, THREADS, 0, stream>>>(p->cols, nullptr, p->seg_row, p->chunk_seg, p->n_chunks, p->chunk_edges,
                                        Zin, T, Zout, partial, ld, F, alpha, epi);
    }
    PPNP_CHECK_LAUNCH("spmm_stream_kernel");
    if (p->n_fix > 0) {
        const int64_t needf = (p->n_fix + groups_per_block - 1) / groups_per_block;
        const int64_t capf = (int64_t)sm_count() * 8;
        dim3 grid((unsigned)(needf < capf ? needf : capf), (unsigned)tiles);
        fixup_kernel<VEC, G><<<grid, THREADS, 0, stream>>>(p->fix_ptr, p->fix_row, p->fix_deg, p->n_fix, partial, T,
                                                           Zout, ld, F, alpha, epi);
        PPNP_CHECK_LAUNCH("fixup_kernel");
    }
    return PPNP_OK;
}

int dispatch_step(const ppnp_plan_t* p, const float* Zin, const float* T, float* Zout, float* partial,
                  int64_t ld, int F, float alpha, int epi, bool use_vals, cudaStream_t stream) {
    const bool vec4 = (F % 4 == 0) && (ld % 4 == 0) && aligned16(Zin) && aligned16(T) && aligned16(Zout) &&
                      (partial == nullptr || aligned16(partial));
#define PPNP_GO(V_, G_) return launch_step<V_, G_>(p, Zin, T, Zout, partial, ld, F, alpha, epi, use_vals, stream)
    if (vec4) {
        const int gl = pow2ceil(F / 4);
        switch (gl >= 32 ? 32 : gl) {
            case 1: PPNP_GO(4, 1);
            case 2: PPNP_GO(4, 2);
            case 4: PPNP_GO(4, 4);
            case 8: PPNP_GO(4, 8);
            case 16: PPNP_GO(4, 16);
            default: PPNP_GO(4, 32);
        }
    } else {
        const int gl = pow2ceil(F);
        switch (gl >= 32 ? 32 : gl) {
            case 1: PPNP_GO(1, 1);
            case 2: PPNP_GO(1, 2);
            case 4: PPNP_GO(1, 4);
            case 8: PPNP_GO(1, 8);
            case 16: PPNP_GO(1, 16);
            default: PPNP_GO(1, 32);
        }
    }
#undef PPNP_GO
}

int validate_plan(const ppnp_plan_t* p) {
    PPNP_REQUIRE(p != nullptr, "plan is null");
    PPNP_REQUIRE(p->n > 0 && p->n_chunks > 0 && p->n_edges > 0, "empty plan");
    PPNP_REQUIRE(p->chunk_edges > 0 && p->chunk_edges % 128 == 0, "chunk_edges must be a multiple of 128");
    PPNP_REQUIRE(p->n_edges == p->n_chunks * (int64_t)p->chunk_edges, "n_edges != n_chunks * chunk_edges");
    PPNP_REQUIRE(p->n_chunks % 32 == 0, "n_chunks must be a multiple of 32");
    PPNP_REQUIRE(p->cols && p->seg_row && p->chunk_seg, "plan arrays missing");
    PPNP_REQUIRE(aligned16(p->cols) && (p->vals == nullptr || aligned16(p->vals)), "plan arrays must be 16-byte aligned");
    PPNP_REQUIRE(p->n_fix == 0 || (p->fix_ptr && p->fix_row && p->fix_deg), "fix arrays missing");
    return PPNP_OK;
}

}  // namespace
}  // namespace ppnp

extern "C" {

int ppnp_spmm_step(const ppnp_plan_t* plan, const float* Zin, const float* T, float* Zout, float* partial,
                   int64_t ld, int32_t F, float alpha, int32_t epi, int32_t use_vals, void* stream) {
    using namespace ppnp;
    int rc = validate_plan(plan);
    if (rc) return rc;
    PPNP_REQUIRE(Zin && T && Zout, "null matrix pointer");
    PPNP_REQUIRE(Zin != Zout, "Zout must not alias Zin");
    PPNP_REQUIRE(F > 0 && ld >= F, "need 0 < F <= ld");
    PPNP_REQUIRE(plan->n_slots == 0 || partial != nullptr, "partial buffer required");
    PPNP_REQUIRE(!use_vals || plan->vals != nullptr, "use_vals needs plan->vals");
    PPNP_REQUIRE(epi >= PPNP_EPI_PLAIN && epi <= PPNP_EPI_RW, "bad epilogue");
    return dispatch_step(plan, Zin, T, Zout, partial, ld, F, alpha, epi, use_vals != 0, as_stream(stream));
}

int ppnp_appnp_propagate(const ppnp_plan_t* plan, const float* H, float* Z, float* scratch, float* partial,
                         int64_t ld, int32_t F, int32_t K, float alpha, int32_t mode, int32_t use_vals,
                         void* stream_) {
    using namespace ppnp;
    int rc = validate_plan(plan);
    if (rc) return rc;
    PPNP_REQUIRE(H && Z && scratch, "null matrix pointer");
    PPNP_REQUIRE(H != Z && H != scratch && Z != scratch, "H, Z, scratch must be distinct buffers");
    PPNP_REQUIRE(F > 0 && ld >= F, "need 0 < F <= ld");
    PPNP_REQUIRE(K >= 1, "K >= 1");
    PPNP_REQUIRE(plan->n_slots == 0 || partial != nullptr, "partial buffer required");
    PPNP_REQUIRE(mode == PPNP_MODE_SYM || mode == PPNP_MODE_RW, "bad mode");
    PPNP_REQUIRE(!(use_vals || (mode == PPNP_MODE_SYM)) || plan->vals != nullptr,
                 "plan->vals required (stored-value steps / first 'sym' step)");
    cudaStream_t stream = as_stream(stream_);
    const float* src = H;
    for (int k = 1; k <= K; ++k) {
        float* dst = ((K - k) % 2 == 0) ? Z : scratch;
        int epi;
        bool vals;
        if (use_vals) { epi = PPNP_EPI_PLAIN; vals = true; }
        else if (mode == PPNP_MODE_RW) { epi = PPNP_EPI_RW; vals = false; }
        else if (K == 1) { epi = PPNP_EPI_PLAIN; vals = true; }
        else if (k == 1) { epi = PPNP_EPI_Z2Y; vals = true; }
        else if (k == K) { epi = PPNP_EPI_Y2Z; vals = false; }
        else { epi = PPNP_EPI_Y; vals = false; }
        rc = dispatch_step(plan, src, H, dst, partial, ld, F, alpha, epi, vals, stream);
        if (rc) return rc;
        src = dst;
    }
    return PPNP_OK;
}

}  // extern "C"
