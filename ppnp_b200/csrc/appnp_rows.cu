// appnp_rows.cu -- the APPNP step for rows of LOW degree: one lane group per row, straight off the CSR.
//
// The chunked edge stream (appnp_spmm.cu) buys load balance for hub rows with per-edge bookkeeping:
// segment flags, segment rows, partial slots.  A power-law graph has few hub rows and millions of rows with a
// handful of edges (BASELINE config 4: 97 % of the rows hold a third of the edges, median degree 1), and for
// those the bookkeeping IS the cost: ncu counts 17 warp instructions per edge in the stream kernel over the
// low-degree rows (profiles/r02_tiled.md).  Here a lane group of G lanes (G x float4 = the feature row) owns
// one row at a time: it reads the row's column ids straight from the CSR (one coalesced load of G ids, one
// shuffle per edge), keeps up to 8 gathers in flight, and finishes the row itself -- no flags, no segment
// table, no partial sums, no second kernel.  Rows are handed out in the caller's order (descending degree:
// the groups of a warp work on rows of nearly the same length, and a grid-stride deal gives every CTA the
// same mix).  The halo push of the partitioned form (ppnp_spmm_step_push) is the same epilogue code.
// Results are bit-identical run to run (one group adds a row's edges in CSR order).
#include "common.cuh"

namespace ppnp {
namespace {

constexpr unsigned FULL = 0xffffffffu;

struct RowsPush {
    const int32_t* ptr;    // [n + 1] per-row ranges into code[], or nullptr (no pushes)
    const int32_t* code;
    const int32_t* first;  // [n] -1: row is not pushed; >= 0: its only destination code; <= -2: several
    float* base[PPNP_MAX_PEERS];
};

template <int G, bool HAS_VAL, bool PUSH>
__global__ void __launch_bounds__(256, 4)
spmm_rows_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices, const float* __restrict__ vals,
                 const int32_t* __restrict__ rows, int64_t n_rows, const float* Zin, const float* T, float* Zout, int ld, int F,
                 float alpha, int epi, const __grid_constant__ RowsPush pa) {
    using V = Vec<4>;
    constexpr int GPW = 32 / G;
    constexpr int B = (G < 8) ? G : 8;         // gathers in flight per lane
    const int lane = threadIdx.x & 31;
    const int lg = lane % G;
    const int grp = lane / G;
    const int f = ((int)blockIdx.y * G + lg) * 4;
    const bool active = f < F;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t total_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const char* zbase = reinterpret_cast<const char*>(Zin + (active ? f : 0));
    const unsigned row_bytes = (unsigned)ld * 4u;

    // a warp takes GPW consecutive rows of the list per round (nearly equal degrees: the list is degree-sorted)
    for (int64_t r0 = warp_global * GPW; r0 < n_rows; r0 += total_warps * GPW) {
        const int64_t ri = r0 + grp;
        const bool have = ri < n_rows;
        const int row = have ? __ldg(rows + ri) : 0;
        int b = 0, e = 0;
        if (have) { b = __ldg(indptr + row); e = __ldg(indptr + row + 1); }
        V t; t.zero();
        int pf = -1;
        if (have && active) t = V::load_stream(T + (int64_t)row * ld + f);
        if (PUSH && have) pf = __ldg(pa.first + row);
        V acc; acc.zero();
        // all groups of the warp run the same number of rounds (shuffles need every lane): the longest row decides
        int len = e - b;
#pragma unroll
        for (int o = G; o < 32; o <<= 1) len = max(len, __shfl_xor_sync(FULL, len, o));
        for (int p = 0; p < len; p += G) {
            const int mine = b + p + lg;
            int idx = -1;
            float wv = 0.f;
            if (mine < e) {
                idx = __ldcs(indices + mine);
                if (HAS_VAL) wv = __ldcs(vals + mine);
            }
#pragma unroll
            for (int k0 = 0; k0 < G; k0 += B) {
                if (p + k0 >= len) break;           // warp-uniform
                V v[B];
                float w[B];
#pragma unroll
                for (int u = 0; u < B; ++u) {
                    const int col = __shfl_sync(FULL, idx, k0 + u, G);
                    if (HAS_VAL) w[u] = __shfl_sync(FULL, wv, k0 + u, G);
                    v[u].zero();
                    if (col >= 0 && active) v[u] = V::load(reinterpret_cast<const float*>(zbase + (uint64_t)(unsigned)col * row_bytes));
                }
#pragma unroll
                for (int u = 0; u < B; ++u) {
                    if (HAS_VAL) acc.fma(w[u], v[u]); else acc.add(v[u]);     // absent edges carry zeros
                }
            }
        }
        if (have && active) {
            float ca, cb;
            epi_coef(epi, alpha, (float)(e - b), ca, cb);
            const V o = V::axpby(ca, acc, cb, t);
            o.store_stream(Zout + (int64_t)row * ld + f);
            if (PUSH) {
                if (pf >= 0) {
                    o.store(pa.base[(pf >> 28) & (PPNP_MAX_PEERS - 1)] + (int64_t)(pf & 0x0fffffff) * ld + f);
                } else if (pf < -1) {
                    const int pb = __ldg(pa.ptr + row), pe = __ldg(pa.ptr + row + 1);
                    for (int i = pb; i < pe; ++i) {
                        const int code = __ldg(pa.code + i);
                        o.store(pa.base[(code >> 28) & (PPNP_MAX_PEERS - 1)] + (int64_t)(code & 0x0fffffff) * ld + f);
                    }
                }
            }
        }
    }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <int G>
int launch_rows(const int32_t* indptr, const int32_t* indices, const float* vals, const int32_t* rows, int64_t n_rows,
                const float* Zin, const float* T, float* Zout, int64_t ld, int F, float alpha, int epi, bool use_vals,
                const RowsPush& pa, cudaStream_t stream) {
    constexpr int THREADS = 256;
    constexpr int GPW = 32 / G;
    const int tiles = (F + G * 4 - 1) / (G * 4);
    const int64_t rows_per_block = (THREADS / 32) * GPW;
    const int64_t need = (n_rows + rows_per_block - 1) / rows_per_block;
    const int64_t cap = (int64_t)sm_count() * 8;
    dim3 grid((unsigned)(need < cap ? need : cap), (unsigned)tiles);
    const bool push = pa.ptr != nullptr;
#define PPNP_RGO(HV_, PU_) spmm_rows_kernel<G, HV_, PU_><<<grid, THREADS, 0, stream>>>(indptr, indices, vals, rows, n_rows, Zin, T, Zout, \
                                                                                      (int)ld, F, alpha, epi, pa)
    if (use_vals) { if (push) PPNP_RGO(true, true); else PPNP_RGO(true, false); }
    else          { if (push) PPNP_RGO(false, true); else PPNP_RGO(false, false); }
#undef PPNP_RGO
    PPNP_CHECK_LAUNCH("spmm_rows_kernel");
    return PPNP_OK;
}

inline int pow2ceil(int x) { int p = 1; while (p < x) p <<= 1; return p; }

}  // namespace
}  // namespace ppnp

extern "C" {

int ppnp_spmm_step_rows(const int32_t* indptr, const int32_t* indices, const float* vals, const int32_t* rows, int64_t n_rows,
                        int64_t n, const float* Zin, const float* T, float* Zout, int64_t ld, int32_t F, float alpha, int32_t epi,
                        int32_t use_vals, const int32_t* push_ptr, const int32_t* push_code, const int32_t* push_first,
                        const void* const* peer_bases_host, int32_t n_peers, void* stream_) {
    using namespace ppnp;
    PPNP_REQUIRE(indptr && indices && rows && n_rows >= 0 && n > 0, "CSR arrays and the row list are required");
    if (n_rows == 0) return PPNP_OK;
    PPNP_REQUIRE(Zin && T && Zout && Zin != Zout, "null or aliased matrix pointer");
    PPNP_REQUIRE(F > 0 && F % 4 == 0 && ld >= F && ld % 4 == 0 && ld < ((int64_t)1 << 30), "need F % 4 == 0, F <= ld < 2^30, ld % 4 == 0");
    PPNP_REQUIRE(aligned16(Zin) && aligned16(T) && aligned16(Zout), "matrices must be 16-byte aligned");
    PPNP_REQUIRE(epi >= PPNP_EPI_PLAIN && epi <= PPNP_EPI_Y02Z, "bad epilogue");
    PPNP_REQUIRE(!use_vals || vals != nullptr, "use_vals needs the stored values");
    PPNP_REQUIRE(push_ptr == nullptr || (push_code && push_first && peer_bases_host && n_peers >= 1 && n_peers <= PPNP_MAX_PEERS),
                 "push lists need codes, the per-row summary and 1..PPNP_MAX_PEERS peer base pointers");
    RowsPush pa{};
    pa.ptr = push_ptr;
    pa.code = push_code;
    pa.first = push_ptr ? push_first : nullptr;
    for (int i = 0; i < PPNP_MAX_PEERS; ++i)
        pa.base[i] = (push_ptr && i < n_peers) ? reinterpret_cast<float*>(const_cast<void*>(peer_bases_host[i])) : nullptr;
    cudaStream_t stream = as_stream(stream_);
    const int gl = pow2ceil(F / 4);
#define PPNP_RLAUNCH(G_) return launch_rows<G_>(indptr, indices, vals, rows, n_rows, Zin, T, Zout, ld, F, alpha, epi, use_vals != 0, pa, stream)
    switch (gl >= 32 ? 32 : gl) {
        case 1: PPNP_RLAUNCH(1);
        case 2: PPNP_RLAUNCH(2);
        case 4: PPNP_RLAUNCH(4);
        case 8: PPNP_RLAUNCH(8);
        case 16: PPNP_RLAUNCH(16);
        default: PPNP_RLAUNCH(32);
    }
#undef PPNP_RLAUNCH
}

}  // extern "C"

extern "C" int ppnp_appnp_propagate_parts(const ppnp_tiled_plan_t* tiled, const ppnp_plan_t* stream_plan, const ppnp_rows_plan_t* rows,
                                          const float* H, float* Z, float* scratch, float* partial, int64_t ld, int32_t F,
                                          int32_t slice_width, int32_t K, float alpha, int32_t mode, int32_t use_vals, void* stream) {
    using namespace ppnp;
    PPNP_REQUIRE(tiled || stream_plan || rows, "at least one part is required");
    PPNP_REQUIRE(H && Z && scratch && H != Z && H != scratch && Z != scratch, "H, Z, scratch must be distinct buffers");
    PPNP_REQUIRE(K >= 1, "K >= 1");
    mode &= ~PPNP_MODE_PER_STEP;      // this entry point always launches per step
    PPNP_REQUIRE(mode == PPNP_MODE_SYM || mode == PPNP_MODE_RW || mode == PPNP_MODE_SYM_Y0, "bad mode");
    const float* src = H;
    for (int k = 1; k <= K; ++k) {
        float* dst = ((K - k) % 2 == 0) ? Z : scratch;
        int epi, vals;
        step_form(mode, use_vals, k, K, epi, vals);
        int rc;
        if (tiled) {
            rc = ppnp_spmm_step_tiled(tiled, src, H, dst, ld, F, slice_width, alpha, epi, vals, stream);
            if (rc) return rc;
        }
        if (stream_plan) {
            rc = ppnp_spmm_step(stream_plan, src, H, dst, partial, ld, F, alpha, epi, vals, stream);
            if (rc) return rc;
        }
        if (rows) {
            rc = ppnp_spmm_step_rows(rows->indptr, rows->indices, rows->vals, rows->rows, rows->n_rows, rows->n, src, H, dst, ld, F, alpha,
                                     epi, vals, nullptr, nullptr, nullptr, nullptr, 0, stream);
            if (rc) return rc;
        }
        src = dst;
    }
    return PPNP_OK;
}
