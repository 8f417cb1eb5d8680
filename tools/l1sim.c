/*
 * l1sim.c -- DESIGN INFRASTRUCTURE (not product code).
 * Per-SM LRU model of the L1 for the gather stream of one APPNP step: how many feature rows
 * cross the L2 -> SM fabric for a given edge stream (ppnp_b200/plan.py) and chunk -> SM schedule.
 *
 *   ./l1sim <cols.i32> <n_chunks> <chunk_edges> <n_rows> <l1_rows> <sms> <chunks_per_unit> [<l2_rows>]
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

/* open-addressing hash per cache: row -> slot of the LRU list */
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

typedef struct { int32_t *key, *prev, *next; hmap_t h; int32_t head, tail, size, cap; int64_t tomb; } cache_t;

static void cache_init(cache_t* c, int cap) {
    c->key = malloc((size_t)cap * 4); c->prev = malloc((size_t)cap * 4); c->next = malloc((size_t)cap * 4);
    c->h.mask = 1; while (c->h.mask < 4 * (int64_t)cap) c->h.mask <<= 1; c->h.mask -= 1;
    c->h.tab = malloc(((size_t)c->h.mask + 1) * 4); memset(c->h.tab, 0xff, ((size_t)c->h.mask + 1) * 4);
    c->head = c->tail = -1; c->size = 0; c->cap = cap; c->tomb = 0;
}

/* returns 1 on miss */
static int cache_touch(cache_t* c, int32_t row) {
    int32_t s = hfind(&c->h, c->key, row);
    if (s >= 0) {
        if (s != c->head) {
            int32_t p = c->prev[s], q = c->next[s];
            c->next[p] = q; if (q >= 0) c->prev[q] = p; else c->tail = p;
            c->prev[s] = -1; c->next[s] = c->head; c->prev[c->head] = s; c->head = s;
        }
        return 0;
    }
    if (c->size == c->cap) {
        s = c->tail; hdel(&c->h, c->key, s); ++c->tomb;
        c->tail = c->prev[s]; if (c->tail >= 0) c->next[c->tail] = -1; else c->head = -1;
    } else s = c->size++;
    c->key[s] = row; c->prev[s] = -1; c->next[s] = c->head;
    if (c->head >= 0) c->prev[c->head] = s; else c->tail = s;
    c->head = s;
    hput(&c->h, c->key, s);
    if (c->tomb > c->cap) {
        memset(c->h.tab, 0xff, ((size_t)c->h.mask + 1) * 4);
        for (int32_t t = c->head; t >= 0; t = c->next[t]) hput(&c->h, c->key, t);
        c->tomb = 0;
    }
    return 1;
}

int main(int argc, char** argv) {
    if (argc < 8) { fprintf(stderr, "usage: see header\n"); return 2; }
    const int64_t n_chunks = atoll(argv[2]); const int W = atoi(argv[3]);
    const int cap = atoi(argv[5]); const int sms = atoi(argv[6]); const int cpu_ = atoi(argv[7]);
    const int l2_rows = argc > 8 ? atoi(argv[8]) : 0;     /* optional shared second level behind the per-SM caches */
    FILE* f = fopen(argv[1], "rb"); if (!f) { perror("cols"); return 1; }
    int32_t* cols = malloc((size_t)n_chunks * W * 4);
    if (fread(cols, 4, (size_t)n_chunks * W, f) != (size_t)(n_chunks * W)) { fprintf(stderr, "short read\n"); return 1; }
    fclose(f);
    cache_t* l1 = malloc(sizeof(cache_t) * sms);
    for (int s = 0; s < sms; ++s) cache_init(&l1[s], cap);
    cache_t l2; if (l2_rows > 0) cache_init(&l2, l2_rows);
    int64_t miss = 0, acc = 0, miss2 = 0;
    const int64_t n_units = (n_chunks + cpu_ - 1) / cpu_;
    /* units in stream order; unit u runs on SM u % sms (grid-stride assignment).  The shared level sees the
       first-level misses in that order -- units of one wave really run concurrently, which this ignores. */
    for (int64_t u = 0; u < n_units; ++u) {
        cache_t* c1 = &l1[u % sms];
        const int64_t c0 = u * cpu_, c1e = (c0 + cpu_ < n_chunks) ? c0 + cpu_ : n_chunks;
        for (int e0 = 0; e0 < W; e0 += 16)
            for (int64_t c = c0; c < c1e; ++c)
                for (int e = e0; e < e0 + 16; ++e) {
                    const int32_t row = cols[c * W + e] & 0x7fffffff;
                    ++acc;
                    if (cache_touch(c1, row)) {
                        ++miss;
                        if (l2_rows > 0) miss2 += cache_touch(&l2, row);
                    }
                }
    }
    printf("edges %lld  L1 misses %lld (%.1f%% of edges)  -> L2->SM rows %.1f%% of the no-reuse figure",
           (long long)acc, (long long)miss, 100.0 * miss / acc, 100.0 * miss / acc);
    if (l2_rows > 0) printf("  L2 misses %lld (%.1f%% of edges)", (long long)miss2, 100.0 * miss2 / acc);
    printf("\n");
    return 0;
}
