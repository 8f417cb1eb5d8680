/*
 * ppnp_rmat.h -- counter-based R-MAT edge generator shared by the CUDA library
 * (device code), the C oracle (host code) and the bench harness.
 *
 * Synthetic-workload plumbing, not part of the reference: BASELINE.json configs
 * 4 and 5 name "synthetic RMAT" graphs, SURVEY.md section 8(d) fixes the recipe
 * (a,b,c,d) = (0.57,0.19,0.19,0.05), ids truncated to n, loops dropped,
 * symmetrised and de-duplicated afterwards.  Every edge is a pure function of
 * (seed, edge_id), so any rank / any device / the host generates the same edge
 * list for the same id range.
 */
#ifndef PPNP_RMAT_H
#define PPNP_RMAT_H

#include <stdint.h>

#if defined(__CUDACC__)
#define PPNP_HD __host__ __device__ __forceinline__
#else
#define PPNP_HD static inline
#endif

/* splitmix64 finaliser */
PPNP_HD uint64_t ppnp_mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

/* 16-bit cumulative thresholds of (a, a+b, a+b+c) = (0.57, 0.76, 0.95) */
#define PPNP_RMAT_T0 37356u
#define PPNP_RMAT_T1 49807u
#define PPNP_RMAT_T2 62259u

/* Edge `e` of the stream `seed`: `scale` quadrant choices, 4 per 64-bit hash. */
PPNP_HD void ppnp_rmat_edge(uint64_t seed, uint64_t e, int scale,
                            uint32_t* src, uint32_t* dst) {
    uint32_t s = 0, d = 0;
    uint64_t r = 0;
    const uint64_t base = ppnp_mix64(seed * 0xD1342543DE82EF95ULL + 0x632BE59BD9B4E019ULL) ^ (e * 0x9E3779B97F4A7C15ULL);
    for (int level = 0; level < scale; ++level) {
        if ((level & 3) == 0) r = ppnp_mix64(base + (uint64_t)(level >> 2) * 0xA0761D6478BD642FULL);
        const uint32_t u = (uint32_t)(r & 0xFFFFu);
        r >>= 16;
        /* quadrant: a -> (0,0)  b -> (0,1)  c -> (1,0)  d -> (1,1) */
        const uint32_t sb = (u >= PPNP_RMAT_T1) ? 1u : 0u;
        const uint32_t db = ((u >= PPNP_RMAT_T0 && u < PPNP_RMAT_T1) || u >= PPNP_RMAT_T2) ? 1u : 0u;
        s = (s << 1) | sb;
        d = (d << 1) | db;
    }
    *src = s;
    *dst = d;
}

#endif /* PPNP_RMAT_H */
