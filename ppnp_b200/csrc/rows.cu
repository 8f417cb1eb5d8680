// rows.cu -- row gather used by the partitioned propagation (ppnp_b200/dist.py): copy rows idx[i] of
// a row-major fp32 matrix into consecutive rows of another one.  The source may live on a PEER GPU
// (symmetric memory mapped over NVLink): the halo rows a shard needs are then pulled straight from
// their owner, de-duplicated, with no pack on the sender and no collective call -- or it is the
// local Z buffer and the kernel is the pack step in front of an NCCL send.
// Pure data movement, bound by NVLink (peer source) or HBM (local source): 2 * F * 4 bytes per row.
#include "common.cuh"

namespace ppnp {
namespace {

template <int VEC, int G>
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ src, int64_t ld_src, const int64_t* __restrict__ idx, int64_t n_rows,
                   int F, float* __restrict__ dst, int64_t ld_dst) {
    using V = Vec<VEC>;
    constexpr int GPW = 32 / G;
    constexpr int U = 8;                      // rows in flight per group
    const int lane = threadIdx.x & 31;
    const int g = lane / G, lg = lane % G;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t total_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    // A warp owns GPW * U consecutive destination rows; in round u its GPW groups write GPW
    // CONSECUTIVE rows, i.e. one contiguous GPW * G * VEC * 4-byte run per store instruction (full
    // 128-byte lines towards a peer over NVLink instead of 64-byte fragments).
    for (int f0 = 0; f0 < F; f0 += G * VEC) {
        const int f = f0 + lg * VEC;
        const bool active = f < F;
        for (int64_t r0 = warp_global * (GPW * U); r0 < n_rows; r0 += total_warps * (GPW * U)) {
            int64_t s[U];
            V v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t r = r0 + u * GPW + g;
                s[u] = (r < n_rows) ? __ldg(idx + r) : -1;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                v[u].zero();
                if (active && s[u] >= 0) v[u] = V::load_stream(src + s[u] * ld_src + f);
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (active && s[u] >= 0) v[u].store(dst + (r0 + u * GPW + g) * ld_dst + f);
        }
    }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace
}  // namespace ppnp

extern "C" int ppnp_gather_rows(const float* src, int64_t ld_src, const int64_t* idx, int64_t n_rows, int32_t F,
                                float* dst, int64_t ld_dst, void* stream_) {
    using namespace ppnp;
    PPNP_REQUIRE(src && dst && (idx || n_rows == 0), "null pointer");
    PPNP_REQUIRE(n_rows >= 0 && F > 0 && ld_src >= F && ld_dst >= F, "bad shape");
    if (n_rows == 0) return PPNP_OK;
    cudaStream_t stream = as_stream(stream_);
    const bool vec4 = (F % 4 == 0) && (ld_src % 4 == 0) && (ld_dst % 4 == 0) && aligned16(src) && aligned16(dst);
    const int64_t cap = (int64_t)sm_count() * 8;
#define PPNP_GR(V_, G_)                                                                                           \
    do {                                                                                                          \
        const int64_t groups_per_block = 8 * (32 / G_);                                                           \
        int64_t need = (n_rows + groups_per_block * 8 - 1) / (groups_per_block * 8);                              \
        if (need > cap) need = cap;                                                                               \
        gather_rows_kernel<V_, G_><<<(unsigned)need, 256, 0, stream>>>(src, ld_src, idx, n_rows, F, dst, ld_dst); \
    } while (0)
    if (vec4) {
        const int l = F / 4;
        if (l <= 1) PPNP_GR(4, 1); else if (l <= 2) PPNP_GR(4, 2); else if (l <= 4) PPNP_GR(4, 4);
        else if (l <= 8) PPNP_GR(4, 8); else if (l <= 16) PPNP_GR(4, 16); else PPNP_GR(4, 32);
    } else {
        if (F <= 1) PPNP_GR(1, 1); else if (F <= 2) PPNP_GR(1, 2); else if (F <= 4) PPNP_GR(1, 4);
        else if (F <= 8) PPNP_GR(1, 8); else if (F <= 16) PPNP_GR(1, 16); else PPNP_GR(1, 32);
    }
#undef PPNP_GR
    PPNP_CHECK_LAUNCH("gather_rows_kernel");
    return PPNP_OK;
}
