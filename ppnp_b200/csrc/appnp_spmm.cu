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
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"

namespace ppnp {
namespace {

constexpr unsigned FULL = 0xffffffffu;

// tuning knobs (overridable at build time for experiments, see tools/)
#ifndef PPNP_SPMM_MINBLOCKS
#define PPNP_SPMM_MINBLOCKS 4   // resident 256-thread CTAs per SM the register budget is capped for
#endif
#ifndef PPNP_PERSISTENT_MAX_CHUNKS
#define PPNP_PERSISTENT_MAX_CHUNKS 4096   // <= 1 M edges: launch latency dominates, use the one-launch K-step kernel
#endif
#ifndef PPNP_SPMM_MINBLOCKS_WIDE
#define PPNP_SPMM_MINBLOCKS_WIDE 3   // same, for the variants that carry several index registers or values
#endif
#ifndef PPNP_SPMM_RING
#define PPNP_SPMM_RING 8        // gathers kept in flight per lane on the rolling path
#endif
#ifndef PPNP_SPMM_SEGPRED
#define PPNP_SPMM_SEGPRED 0     // 1: only the lanes that end a segment fetch their segment row (4-byte staging path)
#endif
#ifndef PPNP_SPMM_FEWENDS
#define PPNP_SPMM_FEWENDS 0     // 1: slabs with at most two segment ends per group keep the rolling gather ring (experimental)
#endif
#ifndef PPNP_SPMM_U4
#define PPNP_SPMM_U4 4          // float4 gathers issued back to back per group (VEC == 4)
#endif

// Finish a segment: either park the partial sum or apply the epilogue and stream the row out.
// Kept out of line so that the (rarely taken, per segment end) code exists once in the kernel.
// Fused halo push (partitioned propagation): a finished row that peers reference is written straight
// into their halo slots over NVLink from the epilogue -- compute and transfer in ONE kernel, the
// exchange overlaps the gathers row by row and nothing but a barrier is left between two steps.
// code = (peer << 28) | row slot in the peer's buffer; base[] = the peers' mappings of the output buffer.
struct PushArgs {
    const int32_t* ptr;    // [n + 1] per-row ranges into code[], or nullptr (no pushes)
    const int32_t* code;
    const int32_t* first;  // [n] -1: row is not pushed; >= 0: its only destination code; <= -2: several, use ptr/code
    float* base[PPNP_MAX_PEERS];
};

template <int VEC>
__device__ __forceinline__ void push_one(const Vec<VEC>& o, int code, int ld, int f, const PushArgs* pa) {
    o.store(pa->base[(code >> 28) & (PPNP_MAX_PEERS - 1)] + (int64_t)(code & 0x0fffffff) * ld + f);
}

template <int VEC>
__device__ __forceinline__ void push_row(const Vec<VEC>& o, int row, int ld, int f, const PushArgs* pa) {
    const int b = __ldg(pa->ptr + row), e = __ldg(pa->ptr + row + 1);
    for (int i = b; i < e; ++i) {
        const int code = __ldg(pa->code + i);
        o.store(pa->base[(code >> 28) & (PPNP_MAX_PEERS - 1)] + (int64_t)(code & 0x0fffffff) * ld + f);
    }
}

// Warp-uniform: no lane group of this warp has more than two segment ends in the slab.
template <int SR, int G>
__device__ __forceinline__ bool few_ends(const unsigned (&ends0)[SR]) {
    int c = 0;
#pragma unroll
    for (int r = 0; r < SR; ++r) c += __popc(ends0[r]);
    return __all_sync(FULL, c <= 2);
}

template <typename V, bool COHERENT>
__device__ __forceinline__ V gather_load(const float* p) {
    if (COHERENT) return V::load_cg(p);
    return V::load(p);
}

template <int VEC, bool PUSH>
__device__ __noinline__ void emit_segment(const Vec<VEC>& acc, const Vec<VEC>& t, int sv, float deg, bool active,
                                          float* Zout, float* __restrict__ partial, int ld, int f,
                                          float alpha, int epi, const float* __restrict__ row_deg,
                                          const PushArgs* pa, int pf) {
    if (!active) return;
    if (sv < 0) {
        acc.store(partial + (int64_t)(sv & 0x7fffffff) * ld + f);
    } else {
        float a, bb;
        if (row_deg != nullptr) deg = __ldg(row_deg + sv);   // stream holds only part of the row
        epi_coef(epi, alpha, deg, a, bb);
        const Vec<VEC> o = Vec<VEC>::axpby(a, acc, bb, t);
        o.store_stream(Zout + (int64_t)sv * ld + f);
        if (PUSH) {
            if (pf >= 0) push_one<VEC>(o, pf, ld, f, pa);          // the common case: one peer wants this row
            else if (pf < -1) push_row<VEC>(o, sv, ld, f, pa);     // several peers
        }
    }
}

// A slab is SE = SR * G consecutive edges of the chunk (16 edges, 32 for G = 32): SR index words per
// lane, word r of lane l being edge r * G + l.
//
// Index traffic never touches registers before it is needed: every lane cp.async's its own index
// words (and values) of slab j + 2*PD and its segment-row words of slab j + PD into private
// shared-memory slots and reads them back with LDS once `cp.async.wait_group` says they have
// landed -- a prefetch distance of PD slabs for both streams with no in-flight register to move or
// spill (a register-based pipeline stalls on exactly those moves, profiles/r01_prof_spmm4).
// A slab without any segment end (warp-wide) takes the rolling path: RING gathers stay in flight per
// lane, each accumulate immediately re-arms its register with the next gather.
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {   // L1-bypassing (.cg): 16 B only
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int G>
struct StageCfg {
    static constexpr int SR = (G >= 16) ? 1 : 16 / G;   // index words per lane and slab
    static constexpr int PD = (SR == 1) ? 2 : 1;        // prefetch distance in slabs
    static constexpr int ISLOTS = (PD == 2) ? 8 : 4;    // >= 2*PD + 1, power of two
    static constexpr int SSLOTS = (PD == 2) ? 4 : 2;    // >= PD + 1, power of two
    static constexpr int words_per_thread(bool has_val) { return SR * (ISLOTS * (has_val ? 2 : 1) + SSLOTS); }
};

// Lane-transposed streams (PPNP_PLAN_LANE_GROUP, G >= 4): the 4 index words a lane consumes over CPS
// consecutive slabs (a "quad") are contiguous in memory, so ONE 16-byte cp.async.cg per lane stages
// them -- a quarter of the copy instructions of the 4-byte form and none of its per-element L1
// data-pipe wavefronts (profiles/r01_prof_spmm5: LDGSTS.32 staging took a third of the pipe).
// Schedule (everything issued in iteration j-1 has landed at the top of iteration j):
//   quad Q (slabs Q*CPS ..) is requested in iteration Q*CPS - 2, read from iteration Q*CPS - 1 (ballot
//   for the segment rows of its first slab) to (Q+1)*CPS - 1; the ring holds NQ quads;
//   segment rows of slab j+1 are requested in iteration j, only by the lanes that end a segment.
template <int G>
struct Stage16Cfg {
    static constexpr int SR = (G >= 16) ? 1 : 16 / G;
    static constexpr int CPS = 4 / SR;                 // slabs per quad
    static constexpr int NQ = (CPS == 1) ? 4 : 2;      // quads in the ring (slot of Q is free again in iteration Q*CPS - 2)
    static constexpr int SSL = 2;                      // segment-row slots: slab j (read) and j + 1 (landing)
    static constexpr int words_per_thread(bool has_val) { return NQ * 4 * (has_val ? 2 : 1) + SSL * SR; }
};

// COHERENT: gathers bypass L1 (ld.global.cg).  Needed when the source was written earlier in the SAME
// launch (persistent K-step kernel): ld.global.nc / L1 hits could return the previous iterate.
template <int VEC, int G, bool HAS_VAL, int U, bool FULL_TILE, bool COHERENT, bool PUSH, int NT = 256, bool IDX16 = false>
__device__ __forceinline__ void
spmm_stream_body(const int32_t* __restrict__ cols, const float* __restrict__ vals,
                 const int32_t* __restrict__ seg_row, const int32_t* __restrict__ chunk_seg,
                 int64_t n_chunks, int chunk_edges,
                 const float* Zin, const float* T,
                 float* Zout, float* partial,   // T may alias Zout (PPNP_EPI_ACC)
                 int ld, int F, float alpha, int epi, const float* __restrict__ row_deg, const PushArgs* pa) {
    using V = Vec<VEC>;
    using SC = StageCfg<G>;
    constexpr int GPW = 32 / G;                     // groups per warp
    constexpr int SR = SC::SR;
    constexpr int SE = SR * G;                      // edges per slab
    constexpr int PD = SC::PD, ISLOTS = SC::ISLOTS, SSLOTS = SC::SSLOTS;
    using S16 = Stage16Cfg<(G >= 4) ? G : 4>;
    constexpr int CPS = S16::CPS, NQ = S16::NQ, SSL = S16::SSL;
    static_assert(!IDX16 || G >= 4, "lane-transposed staging needs lane groups of at least 4");
    constexpr int RING = (SE >= PPNP_SPMM_RING) ? PPNP_SPMM_RING : SE;
    static_assert(G % U == 0, "sub-batch must divide the register slab");
    static_assert(SE % RING == 0, "ring must divide the slab");

    // private staging slots, [slot][word r][thread]: conflict-free, no cross-thread visibility needed
    extern __shared__ int32_t stage[];
    const int tid = threadIdx.x;
    int32_t* s_idx = stage + tid;                                   // + (slot * SR + r) * NT
    int32_t* s_seg = stage + (IDX16 ? NQ * 4 * (HAS_VAL ? 2 : 1) : ISLOTS * SR) * NT + tid;   // + (slot * SR + r) * NT
    float* s_val = reinterpret_cast<float*>(stage + (ISLOTS + SSLOTS) * SR * NT) + tid;
    // IDX16: [quad slot][thread][4 words] for indices, then the same for values, then the segment rows
    int32_t* s_idx4 = stage + tid * 4;                              // + slot * NT * 4 + word
    float* s_val4 = reinterpret_cast<float*>(stage + NQ * 4 * NT) + tid * 4;

    const int lane = tid & 31;
    const int g = lane / G;
    const int lg = lane % G;
    // The groups of a warp run in lock step (same trip counts everywhere), so shuffles and ballots
    // use the full mask; the only group-divergent code is the per-segment emit, which has none.
    const int gshift = g * G;
    const unsigned gbits = (G == 32) ? FULL : ((1u << G) - 1u);
    const unsigned lt = (1u << lg) - 1u;  // lower lanes of my group (after shifting the ballot down)

    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + tid) >> 5;
    const int64_t total_groups = (((int64_t)gridDim.x * blockDim.x) >> 5) * GPW;

    const int f = ((int)blockIdx.y * G + lg) * VEC;
    const bool active = FULL_TILE ? true : (f < F);   // FULL_TILE: F == gridDim.y * G * VEC, no idle lanes
    // row addresses as base + col * row_bytes: one IMAD.WIDE.U32 per gathered row
    const char* zbase = reinterpret_cast<const char*>(Zin + f);
    const char* tbase = reinterpret_cast<const char*>(T + f);
    const unsigned row_bytes = (unsigned)ld * 4u;
    const int n_slabs = chunk_edges / SE;

    for (int64_t c = warp_global * GPW + g; c < n_chunks; c += total_groups) {
        int s = __ldg(chunk_seg + c);       // next segment whose row has not been requested yet
        const int32_t* cp = cols + c * (int64_t)chunk_edges + lg;
        const float* vp = HAS_VAL ? vals + c * (int64_t)chunk_edges + lg : nullptr;
        V acc; acc.zero();
        int seg_begin = 0;  // chunk-local position where the running segment started

        // ---- prologue
        const int32_t* cp16 = cols + c * (int64_t)chunk_edges + lg * 4;      // IDX16: + quad * CPS * SE
        const float* vp16 = HAS_VAL ? vals + c * (int64_t)chunk_edges + lg * 4 : nullptr;
        if constexpr (IDX16) {
            // quads that iteration 0 (and its look-ahead to slab 1) reads; then the segment rows of slab 0
#pragma unroll
            for (int Q = 0; Q * CPS < 2; ++Q) {
                if (Q * CPS < n_slabs) {
                    cp_async16(s_idx4 + (Q % NQ) * NT * 4, cp16 + Q * (CPS * SE));
                    if (HAS_VAL) cp_async16(s_val4 + (Q % NQ) * NT * 4, vp16 + Q * (CPS * SE));
                }
            }
            cp_async_commit();
            cp_async_wait<0>();
#pragma unroll
            for (int r = 0; r < SR; ++r) {
                const int raw = s_idx4[r];                       // quad 0, slab 0, word r
                const unsigned e = (__ballot_sync(FULL, raw < 0) >> gshift) & gbits;
                if (raw < 0) cp_async4(s_seg + r * NT, seg_row + s + __popc(e & lt));
                s += __popc(e);
            }
            cp_async_commit();
        } else {
        // indices of slabs 0 .. 2*PD-1, then segment rows of slabs 0 .. PD-1
#pragma unroll
        for (int t = 0; t < 2 * PD; ++t) {
            const int jt = (t < n_slabs) ? t : n_slabs - 1;
#pragma unroll
            for (int r = 0; r < SR; ++r) {
                cp_async4(s_idx + ((t % ISLOTS) * SR + r) * NT, cp + jt * SE + r * G);
                if (HAS_VAL) cp_async4(s_val + ((t % ISLOTS) * SR + r) * NT, vp + jt * SE + r * G);
            }
        }
        cp_async_commit();
        cp_async_wait<0>();
#pragma unroll
        for (int t = 0; t < PD; ++t) {
#pragma unroll
            for (int r = 0; r < SR; ++r) {
                const int raw = s_idx[((t % ISLOTS) * SR + r) * NT];
                unsigned e = (__ballot_sync(FULL, raw < 0) >> gshift) & gbits;
                if (t >= n_slabs) e = 0;
                if (!PPNP_SPMM_SEGPRED || raw < 0) cp_async4(s_seg + ((t % SSLOTS) * SR + r) * NT, seg_row + s + __popc(e & lt));
                s += __popc(e);
            }
        }
        cp_async_commit();
        cp_async_wait<0>();   // slab 0's segment rows are read in the first iteration
        }

#pragma unroll 1
        for (int j = 0; j < n_slabs; ++j) {
            int raw0[SR], segv0[SR];
            float w0[SR];
            unsigned ends0[SR];
            unsigned any_end = 0;
            if constexpr (IDX16) {
                cp_async_wait<0>();   // indices up to slab j+1 (and the quad requested last time), segment rows of slab j
                // request: the quad whose first slab is j+2, segment rows of slab j+1
                if ((j + 2) % CPS == 0 && j + 2 < n_slabs) {
                    const int Q = (j + 2) / CPS;
                    cp_async16(s_idx4 + (Q % NQ) * NT * 4, cp16 + Q * (CPS * SE));
                    if (HAS_VAL) cp_async16(s_val4 + (Q % NQ) * NT * 4, vp16 + Q * (CPS * SE));
                }
                {
                    const int jn = j + 1;
                    const int32_t* qn = s_idx4 + ((jn / CPS) % NQ) * NT * 4 + (jn % CPS) * SR;
#pragma unroll
                    for (int r = 0; r < SR; ++r) {
                        const int rawn = qn[r];
                        unsigned e = (__ballot_sync(FULL, rawn < 0) >> gshift) & gbits;
                        if (jn >= n_slabs) e = 0;          // past the chunk: stale words, never processed
                        if (rawn < 0 && jn < n_slabs) cp_async4(s_seg + ((jn % SSL) * SR + r) * NT, seg_row + s + __popc(e & lt));
                        s += __popc(e);
                    }
                }
                cp_async_commit();
                {
                    const int32_t* q0 = s_idx4 + ((j / CPS) % NQ) * NT * 4 + (j % CPS) * SR;
                    const float* v0 = s_val4 + ((j / CPS) % NQ) * NT * 4 + (j % CPS) * SR;
#pragma unroll
                    for (int r = 0; r < SR; ++r) {
                        raw0[r] = q0[r];
                        segv0[r] = s_seg[((j % SSL) * SR + r) * NT];
                        if (HAS_VAL) w0[r] = v0[r];
                        ends0[r] = (__ballot_sync(FULL, raw0[r] < 0) >> gshift) & gbits;
                        any_end |= ends0[r];
                    }
                }
            } else {
            // everything but the PD-1 newest groups has landed: indices up to slab j+PD, segment rows up to slab j
            cp_async_wait<PD - 1>();
            // request: indices of slab j+2PD, segment rows of slab j+PD
            {
                const int ji = (j + 2 * PD < n_slabs) ? j + 2 * PD : n_slabs - 1;
                const int si = (j + 2 * PD) & (ISLOTS - 1);
                const int sn = (j + PD) & (ISLOTS - 1), ss = (j + PD) & (SSLOTS - 1);
#pragma unroll
                for (int r = 0; r < SR; ++r) {
                    cp_async4(s_idx + (si * SR + r) * NT, cp + ji * SE + r * G);
                    if (HAS_VAL) cp_async4(s_val + (si * SR + r) * NT, vp + ji * SE + r * G);
                    const int rawn = s_idx[(sn * SR + r) * NT];
                    unsigned e = (__ballot_sync(FULL, rawn < 0) >> gshift) & gbits;
                    if (j + PD >= n_slabs) e = 0;      // past the chunk: the replayed slab is never processed
                    if (!PPNP_SPMM_SEGPRED || rawn < 0) cp_async4(s_seg + (ss * SR + r) * NT, seg_row + s + __popc(e & lt));
                    s += __popc(e);
                }
                cp_async_commit();
            }
            // the slab to process
            {
                const int s0 = j & (ISLOTS - 1), q0 = j & (SSLOTS - 1);
#pragma unroll
                for (int r = 0; r < SR; ++r) {
                    raw0[r] = s_idx[(s0 * SR + r) * NT];
                    segv0[r] = s_seg[(q0 * SR + r) * NT];
                    if (HAS_VAL) w0[r] = s_val[(s0 * SR + r) * NT];
                    ends0[r] = (__ballot_sync(FULL, raw0[r] < 0) >> gshift) & gbits;
                    any_end |= ends0[r];
                }
            }
            }

            if (!__any_sync(FULL, any_end != 0)) {
                // ---- rolling path: no segment end in this slab for any group of the warp
                V v[RING];
                float wv[RING];
#pragma unroll
                for (int e = 0; e < RING; ++e) {
                    const int col = __shfl_sync(FULL, raw0[e / G], e % G, G);
                    if (HAS_VAL) wv[e] = __shfl_sync(FULL, w0[e / G], e % G, G);
                    v[e].zero();
                    if (active) v[e] = gather_load<V, COHERENT>(reinterpret_cast<const float*>(zbase + (uint64_t)(unsigned)col * row_bytes));
                }
#pragma unroll
                for (int e = RING; e < SE; ++e) {
                    if (HAS_VAL) acc.fma(wv[e % RING], v[e % RING]); else acc.add(v[e % RING]);
                    const int col = __shfl_sync(FULL, raw0[e / G], e % G, G);
                    if (HAS_VAL) wv[e % RING] = __shfl_sync(FULL, w0[e / G], e % G, G);
                    if (active) v[e % RING] = gather_load<V, COHERENT>(reinterpret_cast<const float*>(zbase + (uint64_t)(unsigned)col * row_bytes));
                }
#pragma unroll
                for (int e = 0; e < RING; ++e) {
                    if (HAS_VAL) acc.fma(wv[e], v[e]); else acc.add(v[e]);
                }
            } else if (PPNP_SPMM_FEWENDS && G >= 4 && few_ends<SR, G>(ends0)) {
                // ---- rolling path with segment ends: at most two ends per group in this slab (mid-degree rows,
                // carved pieces).  The gather ring keeps rolling; an end is one group-uniform branch after its
                // accumulate.  Only a row that is finished here reads its teleport row, at that point.
                unsigned m = 0;                    // bit k: edge k of the slab ends a segment (k = r * G + lane)
#pragma unroll
                for (int r = 0; r < SR; ++r) m |= ends0[r] << (r * G);
                const int p1 = __ffs(m) - 1;       // >= 0: this branch is taken only when some group has an end,
                const unsigned m2 = m & (m - 1);   // but MY group may have none (p1 = -1)
                const int p2 = __ffs(m2) - 1;
                int sv1 = 0, sv2 = 0;
                {
                    const int q1 = p1 < 0 ? 0 : p1, q2 = p2 < 0 ? 0 : p2;
                    int a1 = segv0[0], a2 = segv0[0];
#pragma unroll
                    for (int r = 1; r < SR; ++r) { if (q1 / G == r) a1 = segv0[r]; if (q2 / G == r) a2 = segv0[r]; }
                    sv1 = __shfl_sync(FULL, a1, q1 % G, G);
                    sv2 = __shfl_sync(FULL, a2, q2 % G, G);
                }
                V v[RING];
                float wv[RING];
#pragma unroll
                for (int e = 0; e < RING; ++e) {
                    const int col = __shfl_sync(FULL, raw0[e / G], e % G, G) & 0x7fffffff;
                    if (HAS_VAL) wv[e] = __shfl_sync(FULL, w0[e / G], e % G, G);
                    v[e].zero();
                    if (active) v[e] = gather_load<V, COHERENT>(reinterpret_cast<const float*>(zbase + (uint64_t)(unsigned)col * row_bytes));
                }
#define PPNP_END_CHECK(K_)                                                                                         \
    if ((K_) == p1 || (K_) == p2) {                                                                                \
        const int sv = ((K_) == p1) ? sv1 : sv2;                                                                   \
        const int pos = j * SE + (K_);                                                                             \
        V t; t.zero();                                                                                             \
        int pf = -1;                                                                                               \
        if (sv >= 0) {                                                                                             \
            if (active) t = V::load_stream(reinterpret_cast<const float*>(tbase + (uint64_t)(unsigned)sv * row_bytes)); \
            if (PUSH) pf = __ldg(pa->first + sv);                                                                  \
        }                                                                                                          \
        {                                                                                                          \
            const V a2 = acc, t2 = t;                                                                              \
            emit_segment<VEC, PUSH>(a2, t2, sv, (float)(pos - seg_begin + 1), active, Zout, partial, ld, f, alpha, epi, row_deg, pa, pf); \
        }                                                                                                          \
        acc.zero();                                                                                                \
        seg_begin = pos + 1;                                                                                       \
    }
#pragma unroll
                for (int e = RING; e < SE; ++e) {
                    if (HAS_VAL) acc.fma(wv[e % RING], v[e % RING]); else acc.add(v[e % RING]);
                    PPNP_END_CHECK(e - RING)
                    const int col = __shfl_sync(FULL, raw0[e / G], e % G, G) & 0x7fffffff;
                    if (HAS_VAL) wv[e % RING] = __shfl_sync(FULL, w0[e / G], e % G, G);
                    if (active) v[e % RING] = gather_load<V, COHERENT>(reinterpret_cast<const float*>(zbase + (uint64_t)(unsigned)col * row_bytes));
                }
#pragma unroll
                for (int e = 0; e < RING; ++e) {
                    if (HAS_VAL) acc.fma(wv[(SE - RING + e) % RING], v[(SE - RING + e) % RING]); else acc.add(v[(SE - RING + e) % RING]);
                    PPNP_END_CHECK(SE - RING + e)
                }
#undef PPNP_END_CHECK
            } else {
#pragma unroll
              for (int r = 0; r < SR; ++r) {
#pragma unroll
                for (int u0 = 0; u0 < G; u0 += U) {
                    const unsigned sub = (ends0[r] >> u0) & ((1u << U) - 1u);
                    if (!__any_sync(FULL, sub != 0)) {   // warp-uniform: no group of this warp ends a segment here
                        V v[U];
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            const int col = __shfl_sync(FULL, raw0[r], u0 + u, G);
                            v[u].zero();
                            if (active) v[u] = gather_load<V, COHERENT>(reinterpret_cast<const float*>(zbase + (uint64_t)(unsigned)col * row_bytes));
                        }
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            if (HAS_VAL) acc.fma(__shfl_sync(FULL, w0[r], u0 + u, G), v[u]); else acc.add(v[u]);
                        }
                    } else {
                        // ---- general path: some of these edges finish a segment
                        V v[U], t[U];
                        int ru[U], sv[U], pf[U];
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            ru[u] = __shfl_sync(FULL, raw0[r], u0 + u, G);
                            sv[u] = __shfl_sync(FULL, segv0[r], u0 + u, G);
                            v[u].zero();
                            t[u].zero();
                            pf[u] = -1;
                            if (PUSH && ru[u] < 0 && sv[u] >= 0) pf[u] = __ldg(pa->first + sv[u]);
                            if (active) v[u] = gather_load<V, COHERENT>(reinterpret_cast<const float*>(zbase + (uint64_t)(unsigned)(ru[u] & 0x7fffffff) * row_bytes));
                            if (active && ru[u] < 0 && sv[u] >= 0) t[u] = V::load_stream(reinterpret_cast<const float*>(tbase + (uint64_t)(unsigned)sv[u] * row_bytes));
                        }
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            if (HAS_VAL) acc.fma(__shfl_sync(FULL, w0[r], u0 + u, G), v[u]); else acc.add(v[u]);
                            if (ru[u] < 0) {
                                const int pos = j * SE + r * G + u0 + u;
                                {   // copies: the out-of-line call takes references, acc itself must stay in registers
                                    const V a2 = acc, t2 = t[u];
                                    emit_segment<VEC, PUSH>(a2, t2, sv[u], (float)(pos - seg_begin + 1), active, Zout, partial, ld, f, alpha, epi, row_deg, pa, pf[u]);
                                }
                                acc.zero();
                                seg_begin = pos + 1;
                            }
                        }
                    }
                }
              }
            }
        }
        cp_async_wait<0>();   // nothing of this chunk may land after its slots are reused
    }
}

template <int VEC, int G, bool HAS_VAL, int U, bool FULL_TILE, bool PUSH, int NT = 256, bool IDX16 = false>
__global__ void __launch_bounds__(NT, (NT == 1024) ? 1 : ((G >= 16 && !HAS_VAL) ? PPNP_SPMM_MINBLOCKS : PPNP_SPMM_MINBLOCKS_WIDE))
spmm_stream_kernel(const int32_t* __restrict__ cols, const float* __restrict__ vals,
                   const int32_t* __restrict__ seg_row, const int32_t* __restrict__ chunk_seg,
                   int64_t n_chunks, int chunk_edges, const float* Zin, const float* T, float* Zout,
                   float* partial, int ld, int F, float alpha, int epi, const float* __restrict__ row_deg,
                   const __grid_constant__ PushArgs pa) {
    spmm_stream_body<VEC, G, HAS_VAL, U, FULL_TILE, false, PUSH, NT, IDX16>(cols, vals, seg_row, chunk_seg, n_chunks, chunk_edges,
                                                                      Zin, T, Zout, partial, ld, F, alpha, epi, row_deg, &pa);
}

// Rows split over several segments: add the partial sums in slot order, then the epilogue.
template <int VEC, int G, bool COHERENT>
__device__ __forceinline__ void
fixup_body(const int32_t* __restrict__ fix_ptr, const int32_t* __restrict__ fix_row,
           const float* __restrict__ fix_deg, int64_t n_fix, const float* partial,
           const float* T, float* Zout, int64_t ld, int F, float alpha, int epi, const PushArgs* pa) {
    using V = Vec<VEC>;
    constexpr int GPW = 32 / G;
    constexpr int U = 8;
    const int lane = threadIdx.x & 31;
    const int g = lane / G;
    const int lg = lane % G;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t total_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int f = ((int)blockIdx.y * G + lg) * VEC;
    const bool active = f < F;
    // One WARP per split row: its GPW lane groups add every GPW-th partial (U loads in flight each), a
    // fixed xor tree over the groups joins them -- a hub row with hundreds of partials is GPW x faster
    // than one group walking them, and the order of the additions is still fixed (bit-reproducible).
    for (int64_t q = warp_global; q < n_fix; q += total_warps) {
        const int s0 = __ldg(fix_ptr + q), s1 = __ldg(fix_ptr + q + 1);
        V acc; acc.zero();
        int s = s0 + g;
        for (; s + (U - 1) * GPW < s1; s += U * GPW) {
            V p[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                p[u].zero();
                if (active) p[u] = COHERENT ? V::load_cg(partial + (int64_t)(s + u * GPW) * ld + f) : V::load_plain(partial + (int64_t)(s + u * GPW) * ld + f);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) acc.add(p[u]);
        }
        for (; s < s1; s += GPW)
            if (active) acc.add(COHERENT ? V::load_cg(partial + (int64_t)s * ld + f) : V::load_plain(partial + (int64_t)s * ld + f));
        __syncwarp();
#pragma unroll
        for (int o = G; o < 32; o <<= 1) acc = V::shfl_xor_add(acc, o);
        if (g == 0 && active) {
            const int row = __ldg(fix_row + q);
            const V t = V::load_stream(T + (int64_t)row * ld + f);
            float a, bb;
            epi_coef(epi, alpha, __ldg(fix_deg + q), a, bb);
            const V o = V::axpby(a, acc, bb, t);
            o.store_stream(Zout + (int64_t)row * ld + f);
            if (pa->ptr != nullptr) push_row<VEC>(o, row, (int)ld, f, pa);
        }
    }
}

template <int VEC, int G>
__global__ void __launch_bounds__(256)
fixup_kernel(const int32_t* __restrict__ fix_ptr, const int32_t* __restrict__ fix_row,
             const float* __restrict__ fix_deg, int64_t n_fix, const float* __restrict__ partial,
             const float* T, float* Zout, int64_t ld, int F, float alpha, int epi,
             const __grid_constant__ PushArgs pa) {
    fixup_body<VEC, G, false>(fix_ptr, fix_row, fix_deg, n_fix, partial, T, Zout, ld, F, alpha, epi, &pa);
}

// All K steps in ONE cooperative launch (north_star: "all K iterations in one persistent launch where
// the graph fits"): grid-wide barriers replace 2K kernel launches, which is what a small graph
// (Cora-ML: 74 chunks) pays for.  Stored-value form in every step (the value stream of a graph this
// small is irrelevant), coherent gathers because the source of step k+1 was written in step k.
template <int VEC, int G, int U, bool FULL_TILE>
__global__ void __launch_bounds__(256, PPNP_SPMM_MINBLOCKS_WIDE)
appnp_persistent_kernel(const int32_t* __restrict__ cols, const float* __restrict__ vals,
                        const int32_t* __restrict__ seg_row, const int32_t* __restrict__ chunk_seg,
                        int64_t n_chunks, int chunk_edges,
                        const int32_t* __restrict__ fix_ptr, const int32_t* __restrict__ fix_row,
                        const float* __restrict__ fix_deg, int64_t n_fix,
                        const float* H, float* Z, float* S, float* partial, int ld, int F, int K, float alpha) {
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    __shared__ PushArgs no_push;
    if (threadIdx.x == 0) { no_push.ptr = nullptr; no_push.first = nullptr; }
    __syncthreads();
    const float* src = H;
    for (int k = 1; k <= K; ++k) {
        float* dst = ((K - k) % 2 == 0) ? Z : S;
        spmm_stream_body<VEC, G, true, U, FULL_TILE, true, false>(cols, vals, seg_row, chunk_seg, n_chunks, chunk_edges, src, H,
                                                            dst, partial, ld, F, alpha, PPNP_EPI_PLAIN, nullptr, &no_push);
        if (n_fix > 0) {
            grid.sync();
            fixup_body<VEC, G, true>(fix_ptr, fix_row, fix_deg, n_fix, partial, H, dst, ld, F, alpha, PPNP_EPI_PLAIN, &no_push);
        }
        grid.sync();
        src = dst;
    }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline int pow2ceil(int x) { int p = 1; while (p < x) p <<= 1; return p; }

template <typename K>
int blocks_per_sm(K kernel, int threads, int smem_bytes) {
    int nb = 0;
    if (smem_bytes > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, smem_bytes) != cudaSuccess || nb < 1) nb = 1;
    return nb;
}

template <int VEC, int G>
int launch_step(const ppnp_plan_t* p, const float* Zin, const float* T, float* Zout, float* partial,
                int64_t ld, int F, float alpha, int epi, bool use_vals, const PushArgs& pa, cudaStream_t stream) {
    constexpr int THREADS = 256;
    constexpr int U = (VEC == 4) ? ((G >= PPNP_SPMM_U4) ? PPNP_SPMM_U4 : G) : ((G >= 8) ? 8 : G);
    constexpr int GPW = 32 / G;
    const int tiles = (F + G * VEC - 1) / (G * VEC);
    const int64_t groups_per_block = (THREADS / 32) * GPW;
    const int64_t need = (p->n_chunks + groups_per_block - 1) / groups_per_block;
    const bool full_tile = (tiles * G * VEC == F);
#define PPNP_LAUNCH2(HV_, FT_, PU_)                                                                                \
    do {                                                                                                           \
        auto k = spmm_stream_kernel<VEC, G, HV_, U, FT_, PU_>;                                                     \
        static thread_local int occ = 0;                                                                           \
        const int smem_bytes = StageCfg<G>::words_per_thread(HV_) * THREADS * 4;                                   \
        if (!occ) occ = blocks_per_sm(k, THREADS, smem_bytes);                                                     \
        const int64_t cap = (int64_t)sm_count() * occ;                                                             \
        dim3 grid((unsigned)(need < cap ? need : cap), (unsigned)tiles);                                           \
        k<<<grid, THREADS, smem_bytes, stream>>>(p->cols, HV_ ? p->vals : nullptr, p->seg_row, p->chunk_seg, p->n_chunks, \
                                        p->chunk_edges, Zin, T, Zout, partial, (int)ld, F, alpha, epi, p->row_deg, pa); \
    } while (0)
#define PPNP_LAUNCH(HV_, FT_) do { if (pa.ptr != nullptr) PPNP_LAUNCH2(HV_, FT_, true); else PPNP_LAUNCH2(HV_, FT_, false); } while (0)
    // plan-selected variants (include/ppnp_b200.h PPNP_PLAN_*): 1024-thread CTAs (one per SM, all its warps on
    // consecutive chunks) and/or 16-byte staging of a lane-transposed stream.  Built for the two headline
    // widths only (F = 64: G = 16, F = 16: G = 4), whole feature tiles, no halo push.
    const int lane_group = PPNP_PLAN_LANE_GROUP(p->flags);
    const bool wide = (p->flags & PPNP_PLAN_WIDE_CTA) != 0;
    if (lane_group != 0 && lane_group != G) {
        set_error("plan is lane-transposed for groups of %d lanes, this feature width uses %d", lane_group, G);
        return PPNP_EINVAL;
    }
    constexpr bool HAS_VARIANTS = (VEC == 4) && (G == 4 || G == 16);
    bool launched = false;
    if constexpr (HAS_VARIANTS) {
        // with a halo push only the 256-thread lane-transposed form exists (the multi-GPU shards are not carved for L1)
        if ((wide || lane_group != 0) && full_tile && (pa.ptr == nullptr || (lane_group != 0 && !wide))) {
#define PPNP_LAUNCH3(HV_, NT_, I16_, PU_)                                                                          \
    do {                                                                                                           \
        auto k = spmm_stream_kernel<VEC, G, HV_, U, true, PU_, NT_, I16_>;                                         \
        static thread_local int occ = 0;                                                                           \
        const int smem_bytes = (I16_ ? Stage16Cfg<G>::words_per_thread(HV_) : StageCfg<G>::words_per_thread(HV_)) * NT_ * 4; \
        if (!occ) {                                                                                                \
            occ = blocks_per_sm(k, NT_, smem_bytes);                                                               \
            /* leave everything the staging does not need to the L1: the carved blocks live there */               \
            const int pct = (int)(((int64_t)(smem_bytes + 1024) * occ * 100 + 228 * 1024 - 1) / (228 * 1024));     \
            cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, pct > 100 ? 100 : pct);        \
        }                                                                                                          \
        const int64_t gpb = (NT_ / 32) * GPW;                                                                      \
        const int64_t need3 = (p->n_chunks + gpb - 1) / gpb;                                                       \
        const int64_t cap = (int64_t)sm_count() * occ;                                                             \
        dim3 grid((unsigned)(need3 < cap ? need3 : cap), (unsigned)tiles);                                         \
        k<<<grid, NT_, smem_bytes, stream>>>(p->cols, HV_ ? p->vals : nullptr, p->seg_row, p->chunk_seg, p->n_chunks,  \
                                       p->chunk_edges, Zin, T, Zout, partial, (int)ld, F, alpha, epi, p->row_deg, pa); \
    } while (0)
            if (pa.ptr != nullptr) {
                if (use_vals) PPNP_LAUNCH3(true, 256, true, true); else PPNP_LAUNCH3(false, 256, true, true);
            } else if (use_vals) {
                if (wide && lane_group) PPNP_LAUNCH3(true, 1024, true, false);
                else if (wide) PPNP_LAUNCH3(true, 1024, false, false);
                else PPNP_LAUNCH3(true, 256, true, false);
            } else {
                if (wide && lane_group) PPNP_LAUNCH3(false, 1024, true, false);
                else if (wide) PPNP_LAUNCH3(false, 1024, false, false);
                else PPNP_LAUNCH3(false, 256, true, false);
            }
#undef PPNP_LAUNCH3
            launched = true;
        }
    }
    if (!launched && lane_group != 0) {
        set_error("lane-transposed plans run only with F = 16 or 64 (whole tiles, 16-byte aligned; with a halo push: 256-thread CTAs)");
        return PPNP_ENOTSUP;
    }
    if (!launched) {
        if (use_vals) { if (full_tile) PPNP_LAUNCH(true, true); else PPNP_LAUNCH(true, false); }
        else          { if (full_tile) PPNP_LAUNCH(false, true); else PPNP_LAUNCH(false, false); }
    }
#undef PPNP_LAUNCH
#undef PPNP_LAUNCH2
    PPNP_CHECK_LAUNCH("spmm_stream_kernel");
    if (p->n_fix > 0) {
        const int64_t needf = (p->n_fix + (THREADS / 32) - 1) / (THREADS / 32);   // one warp per split row
        const int64_t capf = (int64_t)sm_count() * 8;
        dim3 grid((unsigned)(needf < capf ? needf : capf), (unsigned)tiles);
        fixup_kernel<VEC, G><<<grid, THREADS, 0, stream>>>(p->fix_ptr, p->fix_row, p->fix_deg, p->n_fix, partial, T,
                                                           Zout, ld, F, alpha, epi, pa);
        PPNP_CHECK_LAUNCH("fixup_kernel");
    }
    return PPNP_OK;
}

int dispatch_step(const ppnp_plan_t* p, const float* Zin, const float* T, float* Zout, float* partial,
                  int64_t ld, int F, float alpha, int epi, bool use_vals, const PushArgs& pa, cudaStream_t stream) {
    const bool vec4 = (F % 4 == 0) && (ld % 4 == 0) && aligned16(Zin) && aligned16(T) && aligned16(Zout) &&
                      (partial == nullptr || aligned16(partial));
#define PPNP_GO(V_, G_) return launch_step<V_, G_>(p, Zin, T, Zout, partial, ld, F, alpha, epi, use_vals, pa, stream)
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

template <int VEC, int G>
int launch_persistent(const ppnp_plan_t* p, const float* H, float* Z, float* S, float* partial, int64_t ld, int F,
                      int K, float alpha, cudaStream_t stream) {
    constexpr int THREADS = 256;
    constexpr int U = (VEC == 4) ? ((G >= PPNP_SPMM_U4) ? PPNP_SPMM_U4 : G) : ((G >= 8) ? 8 : G);
    constexpr int GPW = 32 / G;
    const int tiles = (F + G * VEC - 1) / (G * VEC);
    const int64_t groups_per_block = (THREADS / 32) * GPW;
    const int64_t need = (p->n_chunks + groups_per_block - 1) / groups_per_block;
    const bool full_tile = (tiles * G * VEC == F);
    const int smem_bytes = StageCfg<G>::words_per_thread(true) * THREADS * 4;
    void* kern = full_tile ? (void*)appnp_persistent_kernel<VEC, G, U, true> : (void*)appnp_persistent_kernel<VEC, G, U, false>;
    int occ = 0;
    if (smem_bytes > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem_bytes) != cudaSuccess || occ < 1) occ = 1;
    int64_t cap = (int64_t)sm_count() * occ / tiles;     // every block of the grid must be co-resident
    if (cap < 1) { set_error("persistent launch: feature tiles exceed the co-resident capacity"); return PPNP_ENOTSUP; }
    const int64_t gx = need < cap ? need : cap;
    int chunk_edges = p->chunk_edges, ldi = (int)ld;
    int64_t n_chunks = p->n_chunks, n_fix = p->n_fix;
    void* args[] = {(void*)&p->cols, (void*)&p->vals, (void*)&p->seg_row, (void*)&p->chunk_seg, &n_chunks, &chunk_edges,
                    (void*)&p->fix_ptr, (void*)&p->fix_row, (void*)&p->fix_deg, &n_fix,
                    (void*)&H, (void*)&Z, (void*)&S, (void*)&partial, &ldi, &F, &K, &alpha};
    int rc = check_cuda(cudaLaunchCooperativeKernel(kern, dim3((unsigned)gx, (unsigned)tiles), dim3(THREADS), args, smem_bytes, stream),
                        "cooperative launch appnp_persistent_kernel");
    return rc;
}

int dispatch_persistent(const ppnp_plan_t* p, const float* H, float* Z, float* S, float* partial, int64_t ld, int F,
                        int K, float alpha, cudaStream_t stream) {
    const bool vec4 = (F % 4 == 0) && (ld % 4 == 0) && aligned16(H) && aligned16(Z) && aligned16(S) &&
                      (partial == nullptr || aligned16(partial));
#define PPNP_GO(V_, G_) return launch_persistent<V_, G_>(p, H, Z, S, partial, ld, F, K, alpha, stream)
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

// PPNP_PERSISTENT=0 in the environment forces the per-step launches (A/B measurements)
bool persistent_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("PPNP_PERSISTENT"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
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
    PPNP_REQUIRE((p->flags & ~(PPNP_PLAN_WIDE_CTA | 0xff00)) == 0, "unknown plan flags");
    {
        const int lgp = PPNP_PLAN_LANE_GROUP(p->flags);
        PPNP_REQUIRE(lgp == 0 || lgp == 4 || lgp == 8 || lgp == 16 || lgp == 32, "lane group must be 4, 8, 16 or 32");
    }
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
    PPNP_REQUIRE(Zin != Zout || (epi & PPNP_EPI_INPLACE), "Zout must not alias Zin (unless PPNP_EPI_INPLACE)");
    PPNP_REQUIRE(F > 0 && ld >= F && ld < ((int64_t)1 << 30), "need 0 < F <= ld < 2^30");
    PPNP_REQUIRE(plan->n_slots == 0 || partial != nullptr, "partial buffer required");
    PPNP_REQUIRE(!use_vals || plan->vals != nullptr, "use_vals needs plan->vals");
    PPNP_REQUIRE((epi & 15) >= PPNP_EPI_PLAIN && (epi & 15) <= PPNP_EPI_Y02Z && (epi & ~63) == 0, "bad epilogue");
    PPNP_REQUIRE(!(epi & PPNP_EPI_ACC) || T == Zout, "PPNP_EPI_ACC adds to the output: pass T == Zout");
    PPNP_REQUIRE(!(epi & PPNP_EPI_INPLACE) || (epi & PPNP_EPI_ACC), "PPNP_EPI_INPLACE goes with PPNP_EPI_ACC");
    PushArgs none{};
    return dispatch_step(plan, Zin, T, Zout, partial, ld, F, alpha, epi, use_vals != 0, none, as_stream(stream));
}

int ppnp_appnp_propagate(const ppnp_plan_t* plan, const float* H, float* Z, float* scratch, float* partial,
                         int64_t ld, int32_t F, int32_t K, float alpha, int32_t mode, int32_t use_vals,
                         void* stream_) {
    using namespace ppnp;
    int rc = validate_plan(plan);
    if (rc) return rc;
    PPNP_REQUIRE(H && Z && scratch, "null matrix pointer");
    PPNP_REQUIRE(H != Z && H != scratch && Z != scratch, "H, Z, scratch must be distinct buffers");
    PPNP_REQUIRE(F > 0 && ld >= F && ld < ((int64_t)1 << 30), "need 0 < F <= ld < 2^30");
    PPNP_REQUIRE(K >= 1, "K >= 1");
    PPNP_REQUIRE(plan->n_slots == 0 || partial != nullptr, "partial buffer required");
    const bool per_step = (mode & PPNP_MODE_PER_STEP) != 0;
    mode &= ~PPNP_MODE_PER_STEP;
    PPNP_REQUIRE(mode == PPNP_MODE_SYM || mode == PPNP_MODE_RW || mode == PPNP_MODE_SYM_Y0, "bad mode");
    PPNP_REQUIRE(mode != PPNP_MODE_SYM_Y0 || !use_vals, "PPNP_MODE_SYM_Y0 is value-free");
    PPNP_REQUIRE(!(use_vals || (mode == PPNP_MODE_SYM)) || plan->vals != nullptr,
                 "plan->vals required (stored-value steps / first 'sym' step)");
    cudaStream_t stream = as_stream(stream_);
    // small graphs: all K steps in one cooperative launch (grid barriers instead of 2K launches)
    if (!per_step && K >= 2 && mode != PPNP_MODE_SYM_Y0 && plan->vals != nullptr && plan->row_deg == nullptr && plan->n_chunks <= PPNP_PERSISTENT_MAX_CHUNKS &&
        PPNP_PLAN_LANE_GROUP(plan->flags) == 0 && persistent_enabled()) {
        return dispatch_persistent(plan, H, Z, scratch, partial, ld, F, K, alpha, stream);
    }
    const float* src = H;
    for (int k = 1; k <= K; ++k) {
        float* dst = ((K - k) % 2 == 0) ? Z : scratch;
        int epi, vals;
        step_form(mode, use_vals, k, K, epi, vals);
        PushArgs none{};
        rc = dispatch_step(plan, src, H, dst, partial, ld, F, alpha, epi, vals != 0, none, stream);
        if (rc) return rc;
        src = dst;
    }
    return PPNP_OK;
}

int ppnp_spmm_step_push(const ppnp_plan_t* plan, const float* Zin, const float* T, float* Zout, float* partial,
                        int64_t ld, int32_t F, float alpha, int32_t epi, int32_t use_vals, const int32_t* push_ptr,
                        const int32_t* push_code, const int32_t* push_first, const void* const* peer_bases_host,
                        int32_t n_peers, void* stream) {
    using namespace ppnp;
    int rc = validate_plan(plan);
    if (rc) return rc;
    PPNP_REQUIRE(Zin && T && Zout, "null matrix pointer");
    PPNP_REQUIRE(Zin != Zout || (epi & PPNP_EPI_INPLACE), "Zout must not alias Zin (unless PPNP_EPI_INPLACE)");
    PPNP_REQUIRE(F > 0 && ld >= F && ld < ((int64_t)1 << 30), "need 0 < F <= ld < 2^30");
    PPNP_REQUIRE(plan->n_slots == 0 || partial != nullptr, "partial buffer required");
    PPNP_REQUIRE(!use_vals || plan->vals != nullptr, "use_vals needs plan->vals");
    PPNP_REQUIRE((epi & 15) >= PPNP_EPI_PLAIN && (epi & 15) <= PPNP_EPI_Y02Z && (epi & ~63) == 0, "bad epilogue");
    PPNP_REQUIRE(push_ptr == nullptr || (push_code && push_first && peer_bases_host && n_peers >= 1 && n_peers <= PPNP_MAX_PEERS),
                 "push lists need codes, the per-row summary and 1..PPNP_MAX_PEERS peer base pointers");
    PushArgs pa{};
    pa.ptr = push_ptr;
    pa.code = push_code;
    pa.first = push_ptr ? push_first : nullptr;
    for (int i = 0; i < PPNP_MAX_PEERS; ++i)
        pa.base[i] = (push_ptr && i < n_peers) ? reinterpret_cast<float*>(const_cast<void*>(peer_bases_host[i])) : nullptr;
    return dispatch_step(plan, Zin, T, Zout, partial, ld, F, alpha, epi, use_vals != 0, pa, as_stream(stream));
}

int ppnp_appnp_propagate_persistent(const ppnp_plan_t* plan, const float* H, float* Z, float* scratch, float* partial,
                                    int64_t ld, int32_t F, int32_t K, float alpha, void* stream_) {
    using namespace ppnp;
    int rc = validate_plan(plan);
    if (rc) return rc;
    PPNP_REQUIRE(H && Z && scratch, "null matrix pointer");
    PPNP_REQUIRE(H != Z && H != scratch && Z != scratch, "H, Z, scratch must be distinct buffers");
    PPNP_REQUIRE(F > 0 && ld >= F && ld < ((int64_t)1 << 30), "need 0 < F <= ld < 2^30");
    PPNP_REQUIRE(K >= 1, "K >= 1");
    PPNP_REQUIRE(plan->vals != nullptr, "the persistent kernel uses the stored values");
    PPNP_REQUIRE(plan->row_deg == nullptr, "partial-row streams are not supported here");
    PPNP_REQUIRE(PPNP_PLAN_LANE_GROUP(plan->flags) == 0, "lane-transposed streams are not supported here");
    PPNP_REQUIRE(plan->n_slots == 0 || partial != nullptr, "partial buffer required");
    return dispatch_persistent(plan, H, Z, scratch, partial, ld, F, K, alpha, as_stream(stream_));
}

}  // extern "C"
