// appnp_tiled.cu -- the APPNP step for the HUB rows of a skewed graph: accumulators resident in shared
// memory, edges walked column window by column window, gathered rows re-used out of the SM's L1.
//
// Why (profiles/r02_*.md): the row-major stream kernel (appnp_spmm.cu) brings one 4*F-byte
// row across the L2 -> SM fabric per edge; on BASELINE config 4 that is 13.3 GB per step against 1.77 GB
// of algorithmic HBM bytes, and the fabric (~6300 B/clk for the whole chip) is what the step waits for.
// Bytes only disappear from that fabric when an SM re-uses a gathered row.  A power-law graph offers
// the re-use in one place: the rows of the highest degrees meet the columns of the highest degrees
// densely (config 4: the top 3 % of the rows hold 67 % of the edges; their 800-row x 512-column tiles
// hold tens of edges per column).  So:
//   * a CTA owns a GROUP of hub rows for the whole step; their accumulators (one slice of `W` floats per
//     row) live in shared memory -- no partial sums ever travel through L2/HBM;
//   * the feature dimension is cut into S = F / W slices, one CTA per (row group, slice): a narrower slice
//     means more rows per CTA and more columns per L1, i.e. more edges per gathered byte;
//   * inside the CTA every WARP owns a disjoint set of the rows (slots) and walks its own edge stream,
//     sorted by (column window, row, column): all warps sweep the column space in the same direction and a
//     column gathered by one warp is an L1 hit for the others while the window is current.  A light
//     pacing rule (a warp may run at most `slack` windows ahead of the slowest one; one shared-memory
//     word per warp, no barrier) keeps the sweep together;
//   * a warp consumes its stream in slabs of 32 edges, lane group g taking the g-th run of 32/NGW edges;
//     a piece (the edges of one row inside one window) that ends inside a run is added to its slot with a
//     plain LDS/FADD/STS (the slot belongs to this warp; a slab in which one slot ends twice is marked by the
//     plan and handled one lane group at a time), partial sums that continue are handed from run to run by
//     shuffles;
//   * at the end the CTA finishes its rows: out = a(deg) * acc + b(deg) * T, streamed out.
// The rows that are not hubs go through the row-major kernel as before (ppnp_appnp_propagate_tiled runs
// both per step).  Results equal the row-major kernel's up to the order of the fp32 additions.
#include "common.cuh"

namespace ppnp {
namespace {

constexpr unsigned FULL = 0xffffffffu;

template <int G>
__device__ __forceinline__ Vec<4> bcast_from_group(const Vec<4>& x, int q, int lg) {
    Vec<4> r;
    const int src = q * G + lg;
    r.v.x = __shfl_sync(FULL, x.v.x, src);
    r.v.y = __shfl_sync(FULL, x.v.y, src);
    r.v.z = __shfl_sync(FULL, x.v.z, src);
    r.v.w = __shfl_sync(FULL, x.v.w, src);
    return r;
}

__device__ __forceinline__ void slot_add(float* acc_smem, int slot, int W, int lg, const Vec<4>& x) {
    float4* p = reinterpret_cast<float4*>(acc_smem + slot * W) + lg;
    float4 a = *p;
    a.x += x.v.x; a.y += x.v.y; a.z += x.v.z; a.w += x.v.w;
    *p = a;
}

__device__ __forceinline__ int ld_stream_i32(const int32_t* p) { return __ldcs(p); }

// G lanes x float4 = one slice of W = 4 G floats.  NGW = 32 / G lane groups per warp, RUN = G edges per group and slab.
template <int G, bool HAS_VAL>
__global__ void __launch_bounds__(512, 1)
spmm_tiled_kernel(const int32_t* __restrict__ cols, const float* __restrict__ vals, const int2* __restrict__ slab_meta,
                  const int32_t* __restrict__ piece_slot, const int32_t* __restrict__ warp_slab_ptr,
                  const int32_t* __restrict__ cta_slot_ptr, const int32_t* __restrict__ slot_row,
                  const float* __restrict__ row_deg, const float* Zin, const float* T, float* Zout, int ld, float alpha,
                  int epi, int slack) {
    using V = Vec<4>;
    constexpr int W = 4 * G;
    constexpr int NGW = 32 / G;
    constexpr int RUN = 32 / NGW;            // == G
    constexpr int B = (RUN < 8) ? RUN : 8;   // gathers in flight per lane
    extern __shared__ __align__(16) float smem[];
    const int NW = blockDim.x >> 5;
    volatile int* wprog = reinterpret_cast<volatile int*>(smem);           // [32] window every warp is in
    float* acc_smem = smem + 32;

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int grp = lane / G, lg = lane % G;
    const int cta = blockIdx.x;
    const int foff = (int)blockIdx.y * W + lg * 4;
    const int slot0 = __ldg(cta_slot_ptr + cta), n_slots = __ldg(cta_slot_ptr + cta + 1) - slot0;

    for (int i = tid; i < n_slots * (W / 4); i += blockDim.x) reinterpret_cast<float4*>(acc_smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid < 32) wprog[tid] = (tid < NW) ? 0 : 0x7fffffff;
    __syncthreads();

    const char* zbase = reinterpret_cast<const char*>(Zin + foff);
    const unsigned row_bytes = (unsigned)ld * 4u;
    const int s0 = __ldg(warp_slab_ptr + cta * NW + w), s1 = __ldg(warp_slab_ptr + cta * NW + w + 1);

    V c_rep; c_rep.zero();      // partial sum of the piece that is open at the slab boundary (same in every lane group)
    V open; open.zero();        // this lane group's share of it, collected over slabs without any piece end
    bool open_dirty = false;
    int cur_win = 0;

    // software pipeline: slab words one slab ahead, slab meta two ahead
    int cw_n = 0, slot_n = 0;
    float val_n = 0.f;
    int2 meta_n = make_int2(0, 0), meta_nn = make_int2(0, 0);
    if (s0 < s1) {
        meta_n = __ldg(slab_meta + s0);
        if (s0 + 1 < s1) meta_nn = __ldg(slab_meta + s0 + 1);
        cw_n = ld_stream_i32(cols + (int64_t)s0 * 32 + lane);
        if (HAS_VAL) val_n = __ldcs(vals + (int64_t)s0 * 32 + lane);
        slot_n = __ldg(piece_slot + meta_n.x + lane);
    }
    const int base = grp * RUN;
    const unsigned below = (base == 0) ? 0u : ((1u << base) - 1u);
    const unsigned runmask = (RUN == 32) ? FULL : ((1u << RUN) - 1u);

#pragma unroll 1
    for (int s = s0; s < s1; ++s) {
        const int cw = cw_n, slotw = slot_n;
        const float valw = val_n;
        const int2 meta = meta_n;
        meta_n = meta_nn;
        if (s + 1 < s1) {
            cw_n = ld_stream_i32(cols + (int64_t)(s + 1) * 32 + lane);
            if (HAS_VAL) val_n = __ldcs(vals + (int64_t)(s + 1) * 32 + lane);
            slot_n = __ldg(piece_slot + meta_n.x + lane);
            if (s + 2 < s1) meta_nn = __ldg(slab_meta + s + 2);
        }
        const bool hz = (meta.y >> 30) & 1;      // some slot ends twice inside this slab: piece ends go one lane group at a time
        const int win = meta.y & 0x3fffffff;
        if (win != cur_win) {        // warp-uniform: entering another column window
            cur_win = win;
            if (lane == 0) wprog[w] = cur_win;
            const int need = cur_win - slack;
            if (need > 0) {
                while (true) {
                    int m = wprog[lane];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(FULL, m, o));
                    if (m >= need) break;
                    __nanosleep(200);
                }
            }
        }
        const unsigned flags = __ballot_sync(FULL, cw < 0);
        const unsigned rf = (flags >> base) & runmask;

        if (flags == 0) {
            // ---- no piece ends in this slab: every lane group adds its run to its open share
#pragma unroll
            for (int k0 = 0; k0 < RUN; k0 += B) {
                V v[B];
                float wv[B];
#pragma unroll
                for (int u = 0; u < B; ++u) {
                    const int col = __shfl_sync(FULL, cw, base + k0 + u);
                    if (HAS_VAL) wv[u] = __shfl_sync(FULL, valw, base + k0 + u);
                    v[u] = V::load(reinterpret_cast<const float*>(zbase + (uint64_t)(unsigned)col * row_bytes));
                }
#pragma unroll
                for (int u = 0; u < B; ++u) {
                    if (HAS_VAL) open.fma(wv[u], v[u]); else open.add(v[u]);
                }
            }
            open_dirty = true;
            continue;
        }

        // ---- general slab
        const int nb = __popc(flags & below);     // piece ends in front of my run = index of my first one in slotw
        V acc; acc.zero();
        V head; head.zero();
        int head_slot = 0;
        bool has = false;
#pragma unroll
        for (int k0 = 0; k0 < RUN; k0 += B) {
            V v[B];
            float wv[B];
#pragma unroll
            for (int u = 0; u < B; ++u) {
                const int col = __shfl_sync(FULL, cw, base + k0 + u) & 0x7fffffff;
                if (HAS_VAL) wv[u] = __shfl_sync(FULL, valw, base + k0 + u);
                v[u] = V::load(reinterpret_cast<const float*>(zbase + (uint64_t)(unsigned)col * row_bytes));
            }
#pragma unroll
            for (int u = 0; u < B; ++u) {
                const int k = k0 + u;
                if (HAS_VAL) acc.fma(wv[u], v[u]); else acc.add(v[u]);
                const int slot_k = __shfl_sync(FULL, slotw, (nb + __popc(rf & ((1u << k) - 1u))) & 31);
                const bool end_k = (rf >> k) & 1u;
                if (!hz) {
                    if (end_k && has) slot_add(acc_smem, slot_k, W, lg, acc);
                } else {
#pragma unroll 1
                    for (int q = 0; q < NGW; ++q) {
                        if (end_k && has && grp == q) slot_add(acc_smem, slot_k, W, lg, acc);
                        __syncwarp();
                    }
                }
                if (end_k) {
                    if (!has) { head = acc; head_slot = slot_k; has = true; }
                    acc.zero();
                }
            }
        }
        // what was open when the slab began: the replicated carry plus the shares of the flag-free slabs before
        if (open_dirty) {        // warp-uniform
#pragma unroll
            for (int o = G; o < 32; o <<= 1) open = V::shfl_xor_add(open, o);
            c_rep.add(open);
            open.zero();
            open_dirty = false;
        }
        // hand the open partial sums from run to run: group q receives c, passes on its tail (plus c when
        // no piece ended inside its run)
        V cin; cin.zero();
#pragma unroll
        for (int q = 0; q < NGW; ++q) {
            if (grp == q) cin = c_rep;
            V out = acc;
            if (!has) out.add(c_rep);
            c_rep = bcast_from_group<G>(out, q, lg);
        }
        head.add(cin);
        if (!hz) {
            if (has) slot_add(acc_smem, head_slot, W, lg, head);
        } else {
#pragma unroll 1
            for (int q = 0; q < NGW; ++q) {
                if (has && grp == q) slot_add(acc_smem, head_slot, W, lg, head);
                __syncwarp();
            }
        }
    }
    __syncwarp();
    if (lane == 0) wprog[w] = 0x7fffffff;
    __syncthreads();

    // ---- the CTA's rows: sum the parts of split rows, epilogue, stream out
    const int n_groups = (blockDim.x >> 5) * NGW;
    const int gid = w * NGW + grp;
    const char* tbase = reinterpret_cast<const char*>(T + foff);
    for (int sl = gid; sl < n_slots; sl += n_groups) {
        const int row = __ldg(slot_row + slot0 + sl);
        if (row < 0) continue;                         // continuation part of a split row
        V a = V::load_plain(acc_smem + sl * W + lg * 4);
        for (int t = sl + 1; t < n_slots && __ldg(slot_row + slot0 + t) < 0; ++t) a.add(V::load_plain(acc_smem + t * W + lg * 4));
        float ca, cb;
        epi_coef(epi, alpha, __ldg(row_deg + row), ca, cb);
        const V tv = V::load_stream(reinterpret_cast<const float*>(tbase + (uint64_t)(unsigned)row * row_bytes));
        const V o = V::axpby(ca, a, cb, tv);
        o.store_stream(Zout + (int64_t)row * ld + foff);
    }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <int G, bool HV>
int launch_tiled(const ppnp_tiled_plan_t* p, const float* Zin, const float* T, float* Zout, int64_t ld, int F, float alpha,
                 int epi, cudaStream_t stream) {
    constexpr int W = 4 * G;
    auto k = spmm_tiled_kernel<G, HV>;
    const int smem_bytes = (32 + p->slots_cap * W) * 4;   // slots_cap counts the spare slot
    static thread_local int configured = 0;
    if (configured != smem_bytes) {
        int rc = check_cuda(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes), "tiled kernel smem opt-in");
        if (rc) return rc;
        // the rest of the 228 KB stays L1: that is where the column windows live
        const int pct = (int)(((int64_t)(smem_bytes + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024));
        cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, pct > 100 ? 100 : pct);
        configured = smem_bytes;
    }
    dim3 grid((unsigned)p->n_ctas, (unsigned)(F / W));
    k<<<grid, p->warps_per_cta * 32, smem_bytes, stream>>>(p->cols, HV ? p->vals : nullptr, reinterpret_cast<const int2*>(p->slab_meta),
                                                           p->piece_slot, p->warp_slab_ptr, p->cta_slot_ptr, p->slot_row, p->row_deg,
                                                           Zin, T, Zout, (int)ld, alpha, epi, p->slack);
    PPNP_CHECK_LAUNCH("spmm_tiled_kernel");
    return PPNP_OK;
}

int validate_tiled(const ppnp_tiled_plan_t* p, int F, int W) {
    PPNP_REQUIRE(p != nullptr, "tiled plan is null");
    PPNP_REQUIRE(p->n > 0 && p->n_ctas > 0 && p->n_slabs > 0, "empty tiled plan");
    PPNP_REQUIRE(p->warps_per_cta >= 1 && p->warps_per_cta <= 16, "1..16 warps per CTA");
    PPNP_REQUIRE(p->cols && p->slab_meta && p->piece_slot && p->warp_slab_ptr && p->cta_slot_ptr && p->slot_row && p->row_deg,
                 "tiled plan arrays missing");
    PPNP_REQUIRE(W == 16 || W == 32 || W == 64, "slice width must be 16, 32 or 64 floats");
    PPNP_REQUIRE(F % W == 0, "F must be a multiple of the slice width");
    PPNP_REQUIRE(p->slots_cap > 0 && (32 + (int64_t)p->slots_cap * W) * 4 <= 200 * 1024, "slot accumulators exceed 200 KB of shared memory");
    PPNP_REQUIRE(p->slack >= 0, "slack >= 0");
    return PPNP_OK;
}

}  // namespace
}  // namespace ppnp

extern "C" {

int ppnp_spmm_step_tiled(const ppnp_tiled_plan_t* plan, const float* Zin, const float* T, float* Zout, int64_t ld, int32_t F,
                         int32_t slice_width, float alpha, int32_t epi, int32_t use_vals, void* stream_) {
    using namespace ppnp;
    int rc = validate_tiled(plan, F, slice_width);
    if (rc) return rc;
    PPNP_REQUIRE(Zin && T && Zout && Zin != Zout, "null or aliased matrix pointer");
    PPNP_REQUIRE(F > 0 && ld >= F && ld % 4 == 0 && ld < ((int64_t)1 << 30), "need 0 < F <= ld < 2^30, ld % 4 == 0");
    PPNP_REQUIRE(aligned16(Zin) && aligned16(T) && aligned16(Zout), "matrices must be 16-byte aligned");
    PPNP_REQUIRE(epi >= PPNP_EPI_PLAIN && epi <= PPNP_EPI_Y02Z, "bad epilogue");
    PPNP_REQUIRE(!use_vals || plan->vals != nullptr, "use_vals needs plan->vals");
    cudaStream_t stream = as_stream(stream_);
#define PPNP_TGO(G_) return use_vals ? launch_tiled<G_, true>(plan, Zin, T, Zout, ld, F, alpha, epi, stream) \
                                     : launch_tiled<G_, false>(plan, Zin, T, Zout, ld, F, alpha, epi, stream)
    switch (slice_width) {
        case 16: PPNP_TGO(4);
        case 32: PPNP_TGO(8);
        default: PPNP_TGO(16);
    }
#undef PPNP_TGO
}

}  // extern "C"
