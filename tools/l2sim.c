/*
 * l2sim.c -- TEST / DESIGN INFRASTRUCTURE (not product code).
 * LRU model of the B200 L2 for the Z-row gather stream of one APPNP step on the
 * config-4 RMAT graph, used to choose the node ordering before spending GPU time.
 *
 *   ./l2sim <n> <raw_draws> <scale> <row_bytes> <cache_MB> <pollute 0|1>
 *
 * Orderings simulated: natural, degree-descending, hub-cluster (each hub followed
 * by its still-unplaced neighbours), random.  Output: gather misses and the DRAM
 * traffic they imply next to the compulsory figure of SURVEY.md section 8(d).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int64_t oracle_rmat_edges(uint64_t, int, int64_t, int64_t, int64_t, int32_t*, int32_t*);
int64_t oracle_sym_csr(int64_t, int64_t, const int32_t*, const int32_t*, int64_t*, int32_t*);

typedef struct { int32_t *prev, *next; uint8_t* in; int32_t head, tail; int64_t size, cap; } lru_t;

static void lru_init(lru_t* c, int64_t n, int64_t cap) {
    c->prev = malloc(n * 4); c->next = malloc(n * 4); c->in = calloc(n, 1);
    c->head = c->tail = -1; c->size = 0; c->cap = cap;
}
static void lru_unlink(lru_t* c, int32_t x) {
    int32_t p = c->prev[x], q = c->next[x];
    if (p >= 0) c->next[p] = q; else c->head = q;
    if (q >= 0) c->prev[q] = p; else c->tail = p;
}
static void lru_push(lru_t* c, int32_t x) {
    c->prev[x] = -1; c->next[x] = c->head;
    if (c->head >= 0) c->prev[c->head] = x; else c->tail = x;
    c->head = x;
}
/* returns 1 on miss */
static int lru_touch(lru_t* c, int32_t x) {
    if (c->in[x]) { lru_unlink(c, x); lru_push(c, x); return 0; }
    if (c->size == c->cap) { int32_t v = c->tail; lru_unlink(c, v); c->in[v] = 0; c->size--; }
    lru_push(c, x); c->in[x] = 1; c->size++;
    return 1;
}

static int cmp_i32_local(const int32_t* a, const int32_t* b) { return (*a > *b) - (*a < *b); }
static int64_t* g_deg;
static int cmp_deg_desc(const void* a, const void* b) {
    int32_t x = *(const int32_t*)a, y = *(const int32_t*)b;
    if (g_deg[x] != g_deg[y]) return (g_deg[x] < g_deg[y]) - (g_deg[x] > g_deg[y]);
    return (x > y) - (x < y);
}

/* order[p] = old id processed at position p.  Simulates gathers of (A+I) rows. */
static void simulate(const char* name, int64_t n, const int64_t* indptr, const int32_t* indices,
                     const int32_t* order, int64_t cap_rows, int pollute, int row_bytes) {
    /* ids in the cache are NEW labels only through identity: a relabel does not change which
       rows are touched together, only the processing order, so old ids are fine as keys.
       Pollution lines use ids n..n+2*cap (never re-touched). */
    lru_t c; lru_init(&c, n + 2 * n + 2, cap_rows);
    int64_t miss = 0, acc = 0, poll = n;
    for (int64_t p = 0; p < n; ++p) {
        const int32_t i = order[p];
        miss += lru_touch(&c, i); acc++;                     /* self loop */
        for (int64_t t = indptr[i]; t < indptr[i + 1]; ++t) { miss += lru_touch(&c, indices[t]); acc++; }
        if (pollute) { lru_touch(&c, (int32_t)poll++); lru_touch(&c, (int32_t)poll++); }
    }
    const double comp = (double)n * row_bytes;
    printf("%-14s accesses %lld  misses %lld (%.2f%%)  gather DRAM %.3f GB  (compulsory %.3f GB, x%.2f)\n",
           name, (long long)acc, (long long)miss, 100.0 * miss / acc, miss * (double)row_bytes / 1e9,
           comp / 1e9, miss * (double)row_bytes / comp);
    fflush(stdout);
    free(c.prev); free(c.next); free(c.in);
}

int main(int argc, char** argv) {
    const int64_t n = argc > 1 ? atoll(argv[1]) : 2000000;
    const int64_t raw = argc > 2 ? atoll(argv[2]) : 26400000;
    const int scale = argc > 3 ? atoi(argv[3]) : 21;
    const int row_bytes = argc > 4 ? atoi(argv[4]) : 256;
    const double cache_mb = argc > 5 ? atof(argv[5]) : 96.0;
    const int pollute = argc > 6 ? atoi(argv[6]) : 1;
    int32_t* src = malloc(raw * 4); int32_t* dst = malloc(raw * 4);
    int64_t m = oracle_rmat_edges(0, scale, n, 0, raw, src, dst);
    int64_t* indptr = malloc((n + 1) * 8); int32_t* indices = malloc(2 * m * 4);
    int64_t nnz = oracle_sym_csr(n, m, src, dst, indptr, indices);
    free(src); free(dst);
    int64_t* deg = malloc(n * 8); int64_t iso = 0, dmax = 0;
    for (int64_t i = 0; i < n; ++i) { deg[i] = indptr[i + 1] - indptr[i]; iso += deg[i] == 0; if (deg[i] > dmax) dmax = deg[i]; }
    printf("n %lld kept draws %lld nnz(A) %lld nnz(A_hat) %lld isolated %.1f%% max deg %lld\n",
           (long long)n, (long long)m, (long long)nnz, (long long)(nnz + n), 100.0 * iso / n, (long long)dmax);
    const int64_t cap_rows = (int64_t)(cache_mb * 1e6 / row_bytes);
    printf("cache %.0f MB = %lld rows of %d B, pollute=%d\n", cache_mb, (long long)cap_rows, row_bytes, pollute);

    int32_t* order = malloc(n * 4);
    for (int64_t i = 0; i < n; ++i) order[i] = (int32_t)i;
    simulate("natural", n, indptr, indices, order, cap_rows, pollute, row_bytes);

    g_deg = deg;
    qsort(order, n, 4, cmp_deg_desc);
    simulate("degree-desc", n, indptr, indices, order, cap_rows, pollute, row_bytes);

    /* hub-cluster: walk degree-desc; place vertex, then its unplaced neighbours of degree <= T */
    {
        int32_t* by_deg = malloc(n * 4); memcpy(by_deg, order, n * 4);
        for (int T = 2; T <= 32 && !getenv("L2SIM_FAST"); T *= 4) {
            uint8_t* placed = calloc(n, 1); int64_t p = 0;
            for (int64_t q = 0; q < n; ++q) {
                const int32_t h = by_deg[q];
                if (placed[h]) continue;
                placed[h] = 1; order[p++] = h;
                for (int64_t t = indptr[h]; t < indptr[h + 1]; ++t) {
                    const int32_t v = indices[t];
                    if (!placed[v] && deg[v] <= T) { placed[v] = 1; order[p++] = v; }
                }
            }
            char nm[32]; snprintf(nm, sizeof nm, "hubclust T=%d", T);
            simulate(nm, n, indptr, indices, order, cap_rows, pollute, row_bytes);
            free(placed);
        }
        free(by_deg);
    }

    /* 2D hub blocking on degree-desc labels: hub rows (top NH by degree) are processed
       column-block-major (virtual rows -> partials); other rows with the block that holds them.
       NH / BS lists come from the environment (L2SIM_NH, L2SIM_BS, comma separated). */
    {
        qsort(order, n, 4, cmp_deg_desc);              /* order[p] = old id with rank p */
        int32_t* rank = malloc(n * 4);
        for (int64_t p = 0; p < n; ++p) rank[order[p]] = (int32_t)p;
        int64_t NHs[8] = {20000, 50000, 100000, 200000}, BSs[8] = {125000, 250000, 500000};
        int nNH = 4, nBS = 3;
        const char* e1 = getenv("L2SIM_NH"); const char* e2 = getenv("L2SIM_BS");
        if (e1) { nNH = 0; char* t = strdup(e1); for (char* q = strtok(t, ","); q && nNH < 8; q = strtok(NULL, ",")) NHs[nNH++] = atoll(q); }
        if (e2) { nBS = 0; char* t = strdup(e2); for (char* q = strtok(t, ","); q && nBS < 8; q = strtok(NULL, ",")) BSs[nBS++] = atoll(q); }
        for (int a = 0; a < nNH; ++a) for (int bsi = 0; bsi < nBS; ++bsi) {
            const int64_t NH = NHs[a] < n ? NHs[a] : n, BS = BSs[bsi];
            const int64_t nb = (n + BS - 1) / BS;
            /* hub edge lists as column RANKS (self loop included), sorted: one pass per hub, then cursors */
            int64_t* hp = malloc((NH + 1) * 8); hp[0] = 0;
            for (int64_t p = 0; p < NH; ++p) hp[p + 1] = hp[p] + (indptr[order[p] + 1] - indptr[order[p]]) + 1;
            int32_t* hr = malloc(hp[NH] * 4);
#pragma omp parallel for schedule(dynamic, 64)
            for (int64_t p = 0; p < NH; ++p) {
                const int32_t h = order[p]; int64_t o = hp[p];
                hr[o++] = (int32_t)p;
                for (int64_t t = indptr[h]; t < indptr[h + 1]; ++t) hr[o++] = rank[indices[t]];
                qsort(hr + hp[p], hp[p + 1] - hp[p], 4, (int (*)(const void*, const void*))cmp_i32_local);
            }
            int64_t* cur = malloc(NH * 8); memcpy(cur, hp, NH * 8);
            lru_t c; lru_init(&c, 4 * n + 2, cap_rows);
            int64_t miss = 0, acc = 0, poll = n, nvirt = 0;
            for (int64_t b = 0; b < nb; ++b) {
                const int64_t lo = b * BS, hi = (b + 1) * BS < n ? (b + 1) * BS : n;
                for (int64_t p = 0; p < NH; ++p) {
                    int any = 0;
                    while (cur[p] < hp[p + 1] && hr[cur[p]] < hi) { miss += lru_touch(&c, order[hr[cur[p]]]); acc++; cur[p]++; any = 1; }
                    if (any) { nvirt++; if (pollute) { lru_touch(&c, (int32_t)poll++); if (poll >= 4 * n) poll = n; } }
                }
                for (int64_t p = (lo > NH ? lo : NH); p < hi; ++p) {
                    const int32_t i = order[p];
                    miss += lru_touch(&c, i); acc++;
                    for (int64_t t = indptr[i]; t < indptr[i + 1]; ++t) { miss += lru_touch(&c, indices[t]); acc++; }
                    if (pollute) { lru_touch(&c, (int32_t)poll++); if (poll >= 4 * n) poll = n; lru_touch(&c, (int32_t)poll++); if (poll >= 4 * n) poll = n; }
                }
            }
            printf("hub2d NH=%lld BS=%lld: accesses %lld misses %lld (%.2f%%) gather DRAM %.3f GB, virt rows %lld (partials %.3f GB w+r)\n",
                   (long long)NH, (long long)BS, (long long)acc, (long long)miss, 100.0 * miss / acc,
                   miss * (double)row_bytes / 1e9, (long long)nvirt, 2.0 * nvirt * row_bytes / 1e9);
            fflush(stdout);
            free(c.prev); free(c.next); free(c.in); free(hp); free(hr); free(cur);
        }
        free(rank);
    }

    /* random */
    {
        for (int64_t i = 0; i < n; ++i) order[i] = (int32_t)i;
        uint64_t s = 12345;
        for (int64_t i = n - 1; i > 0; --i) {
            s = s * 6364136223846793005ULL + 1442695040888963407ULL;
            int64_t j = (int64_t)((s >> 33) % (uint64_t)(i + 1));
            int32_t t = order[i]; order[i] = order[j]; order[j] = t;
        }
        simulate("random", n, indptr, indices, order, cap_rows, pollute, row_bytes);
    }
    return 0;
}
