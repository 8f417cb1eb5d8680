/*
 * l1sim.c -- DESIGN INFRASTRUCTURE (not product code).
 * Per-SM LRU model of the L1 for the gather stream of one APPNP step: how many feature rows
 * cross the L2 -> SM fabric for a given edge stream (ppnp_b200/plan.py) and chunk -> SM schedule.
 *
 *   ./l1sim <cols.i32> <n_chunks> <chunk_edges> <n_rows> <l1_rows> <sms> <chunks_per_unit>
 *
 * cols.i32: plan.cols dumped as raw int32 (bit 31 = segment end).  Unit u (chunks_per_unit
 * consecutive chunks = what the CTAs resident on one SM walk at a time) runs on SM u % sms, as
 * the grid-stride loop of csrc/appnp_spmm.cu assigns it; each SM has its own LRU of l1_rows rows.
 * The edges of a unit are touched slab-interleaved (all its chunks advance together), as the
 * warps of an SM do.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct { int32_t *key, *prev, *next; int32_t head, tail, size, cap; int32_t* where; } lru_t;
/* where[row] = slot or -1, per SM: too big for 148 x n; use an open-addressing hash per SM instead */
typedef struct { int32_t* tab; int32_t mask; } hmap_t;
static inline uint32_t hsh(uint32_t x) { x *= 0x9E3779B1u; return x ^ (x >> 15); }
static int32_t hfind(const hmap_t* h, const int32_t* key, int32_t row) {
    uint32_t i = hsh((uint32_t)row) & h->mask;
    while (h->tab[i] != -1) { if (h->tab[i] >= 0 && key[h->tab[i]] == row) return h->tab[i]; i = (i + 1) & h->mask; }
    return -1;
}
static void hput(hmap_t* h, const int32_t* key, int32_t slot) {
    uint32_t i = hsh((uint32_t)key[slot]) & h->mask;
    while (h->tab[i] >= 0) i = (i + 1) & h->mask;
    h->tab[i] = slot;
}
static void hdel(hmap_t* h, const int32_t* key, int32_t slot) {
    uint32_t i = hsh((uint32_t)key[slot]) & h->mask;
    while (h->tab[i] != slot) i = (i + 1) & h->mask;
    h->tab[i] = -2;   /* tombstone */
}

int main(int argc, char** argv) {
    if (argc < 8) { fprintf(stderr, "usage: see header\n"); return 2; }
    const int64_t n_chunks = atoll(argv[2]); const int W = atoi(argv[3]);
    const int cap = atoi(argv[5]); const int sms = atoi(argv[6]); const int cpu_ = atoi(argv[7]);
    FILE* f = fopen(argv[1], "rb"); if (!f) { perror("cols"); return 1; }
    int32_t* cols = malloc(n_chunks * W * 4);
    if (fread(cols, 4, n_chunks * W, f) != (size_t)(n_chunks * W)) { fprintf(stderr, "short read\n"); return 1; }
    fclose(f);
    int64_t miss = 0, acc = 0, rebuilds = 0;
    const int64_t n_units = (n_chunks + cpu_ - 1) / cpu_;
    for (int sm = 0; sm < sms; ++sm) {
        int32_t* key = malloc(cap * 4); int32_t* prev = malloc(cap * 4); int32_t* next = malloc(cap * 4);
        hmap_t h; h.mask = 1; while (h.mask < 4 * cap) h.mask <<= 1; h.mask -= 1;
        h.tab = malloc((h.mask + 1) * 4); memset(h.tab, 0xff, (h.mask + 1) * 4);
        int32_t head = -1, tail = -1, size = 0; int64_t tomb = 0;
        for (int64_t u = sm; u < n_units; u += sms) {
            const int64_t c0 = u * cpu_, c1 = (c0 + cpu_ < n_chunks) ? c0 + cpu_ : n_chunks;
            for (int e0 = 0; e0 < W; e0 += 16)
                for (int64_t c = c0; c < c1; ++c)
                    for (int e = e0; e < e0 + 16; ++e) {
                        const int32_t row = cols[c * W + e] & 0x7fffffff;
                        ++acc;
                        int32_t s = hfind(&h, key, row);
                        if (s >= 0) {          /* hit: move to front */
                            if (s != head) {
                                int32_t p = prev[s], q = next[s];
                                next[p] = q; if (q >= 0) prev[q] = p; else tail = p;
                                prev[s] = -1; next[s] = head; prev[head] = s; head = s;
                            }
                            continue;
                        }
                        ++miss;
                        if (size == cap) {     /* evict the tail, reuse its slot */
                            s = tail; hdel(&h, key, s); ++tomb;
                            tail = prev[s]; if (tail >= 0) next[tail] = -1; else head = -1;
                        } else s = size++;
                        key[s] = row; prev[s] = -1; next[s] = head; if (head >= 0) prev[head] = s; else tail = s; head = s;
                        hput(&h, key, s);
                        if (tomb > cap) {      /* rebuild the hash without tombstones */
                            memset(h.tab, 0xff, (h.mask + 1) * 4);
                            for (int32_t t = head; t >= 0; t = next[t]) hput(&h, key, t);
                            tomb = 0; ++rebuilds;
                        }
                    }
        }
        free(key); free(prev); free(next); free(h.tab);
    }
    printf("edges %lld  L1 misses %lld (%.1f%% of edges)  -> L2->SM rows %.1f%% of the no-reuse figure\n",
           (long long)acc, (long long)miss, 100.0 * miss / acc, 100.0 * miss / acc);
    (void)rebuilds;
    return 0;
}
