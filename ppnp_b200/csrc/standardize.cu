// standardize.cu -- graph standardisation on the GPU: the step right before the propagation path
// (SURVEY.md section 8f, rank 1).  Replaces, for the unit-weight pipeline main.py:75 runs,
//   ppnp/data/sparsegraph.py:191-222  SparseGraph.standardize
//     :150-154 to_unweighted            every stored entry counts as 1 (weights are never read)
//     :127-148 to_undirected            pattern union of A and A^T          -> emit_keys + sort + unique
//     :381-395 remove_self_loops        diagonal entries dropped            -> emit_keys
//     :355-379 largest_connected_components (scipy weak components, largest kept)
//                                                                           -> hook / flatten / size / best
//     :300-352 create_subgraph          kept nodes in ascending order, relabelled -> scans + compact
// All of it is integer work on (row, column) keys: HBM-bound sort, select, scan and scatter passes;
// results are bit-exact with the reference (tests/golden/standardize_cases.npz).
//
// Every edge becomes the 64-bit key (row << 32) | column; a radix sort of the keys of A (and of A^T)
// followed by a unique pass IS the canonical CSR of the symmetrised pattern: sorted, de-duplicated,
// row pointers by binary search.  Components: lock-free union-find (link the larger root under the
// smaller with one CAS, path halving on the way), so a component's root is its smallest node -- the
// order in which scipy numbers components -- and a packed atomicMax over (size << 32 | root) picks
// the largest component with the reference's tie rule for up to 16 components
// (np.argsort(sizes)[::-1][0]: the last of equal sizes).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>

#include "common.cuh"

namespace ppnp {
namespace {

constexpr unsigned long long SENTINEL = ~0ull;   // sorts last; row part 0xffffffff is no valid node
constexpr int THREADS = 256;

inline int64_t align256(int64_t x) { return (x + 255) & ~(int64_t)255; }

inline unsigned grid_for(int64_t items, int per_sm = 8) {
    int64_t need = (items + THREADS - 1) / THREADS;
    const int64_t cap = (int64_t)sm_count() * per_sm;     // grid-stride loops: a few CTAs per SM
    if (need > cap) need = cap;
    if (need < 1) need = 1;
    return (unsigned)need;
}

__device__ __forceinline__ int64_t upper_bound_i64(const int64_t* __restrict__ a, int64_t lo, int64_t hi, int64_t key) {
    while (lo < hi) {   // first position with a[pos] > key
        const int64_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(a + mid) <= key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// One thread per stored entry: its key, and the key of the transposed entry.
__global__ void emit_keys_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, int64_t n,
                                 int64_t nnz, int undirected, int drop_loops, unsigned long long* __restrict__ keys,
                                 int64_t* __restrict__ status) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = upper_bound_i64(indptr, 0, n + 1, e) - 1;      // indptr[r] <= e < indptr[r+1]
        const int64_t c = __ldg(indices + e);
        unsigned long long k = SENTINEL, kt = SENTINEL;
        if (c < 0 || c >= n || r < 0 || r >= n) {
            atomicExch(reinterpret_cast<unsigned long long*>(status), 1ull);   // reported by the host wrapper
        } else if (r != c) {
            k = ((unsigned long long)r << 32) | (unsigned long long)c;
            kt = ((unsigned long long)c << 32) | (unsigned long long)r;
        } else if (!drop_loops) {
            k = ((unsigned long long)r << 32) | (unsigned long long)c;         // its transpose is itself
        }
        keys[e] = k;
        if (undirected) keys[nnz + e] = kt;
    }
}

// After the unique pass the sentinel, if any, is the last key: do not count it.
__global__ void drop_sentinel_kernel(const unsigned long long* __restrict__ ukeys, int64_t* __restrict__ m) {
    const int64_t v = *m;
    if (v > 0 && ukeys[v - 1] == SENTINEL) *m = v - 1;
}

// Row pointers of the sorted unique keys: indptr[r] = first key whose row is >= r.
__global__ void row_pointers_kernel(const unsigned long long* __restrict__ ukeys, const int64_t* __restrict__ m_dev,
                                    int64_t n, int64_t* __restrict__ indptr) {
    const int64_t m = *m_dev;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= n; r += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long key = (unsigned long long)r << 32;
        int64_t lo = 0, hi = m;
        while (lo < hi) {
            const int64_t mid = lo + ((hi - lo) >> 1);
            if (ukeys[mid] < key) lo = mid + 1; else hi = mid;
        }
        indptr[r] = lo;
    }
}

__global__ void init_parent_kernel(int32_t* __restrict__ parent, int32_t* __restrict__ size, int64_t n) {
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
        parent[v] = (int32_t)v;
        size[v] = 0;
    }
}

// Root of x with path halving.  Parents only ever point at smaller ids (ancestors), a non-root never
// becomes a root again, and links are made by CAS on roots only: the plain stores below race with
// nothing but other shortenings of the same path.
__device__ __forceinline__ int32_t find_root(volatile int32_t* parent, int32_t x) {
    int32_t p = parent[x];
    while (p != x) {
        const int32_t gp = parent[p];
        if (gp != p) parent[x] = gp;
        x = p;
        p = gp;
    }
    return x;
}

// One thread per key (u, v): union of the two endpoints (weak connectivity: direction is ignored).
__global__ void hook_kernel(const unsigned long long* __restrict__ ukeys, const int64_t* __restrict__ m_dev,
                            int32_t* parent, int undirected) {
    const int64_t m = *m_dev;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long k = ukeys[i];
        const int32_t u = (int32_t)(k >> 32), v = (int32_t)(k & 0xffffffffu);
        if (u == v || (undirected && u > v)) continue;      // loops join nothing; (v, u) is in the list as well
        while (true) {
            int32_t ru = find_root(parent, u), rv = find_root(parent, v);
            if (ru == rv) break;
            if (ru < rv) { const int32_t t = ru; ru = rv; rv = t; }      // ru: the larger root goes under the smaller
            if (atomicCAS(parent + ru, ru, rv) == ru) break;
        }
    }
}

// Component of every node = the root of its tree.  The forest is final here (all links were made by
// hook_kernel, a kernel boundary ago), and this pass only READS it: a walk that shortened paths while
// another thread stored its root into the same array could overwrite that root with a mere ancestor
// (round 1's flatten did exactly that and dropped 2 of CiteSeer's 2110 nodes on some runs).  Roots go
// to a separate array, so the result is a pure function of the forest whatever the interleaving.
__global__ void flatten_kernel(const int32_t* __restrict__ parent, int32_t* __restrict__ comp,
                               int32_t* __restrict__ size, int64_t n) {
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
        int32_t x = (int32_t)v, p = __ldg(parent + x);
        while (p != x) { x = p; p = __ldg(parent + x); }
        comp[v] = x;
        atomicAdd(size + x, 1);
    }
}

// (size << 32 | root) over the roots: the largest component, the larger root among equals.
__global__ void best_component_kernel(const int32_t* __restrict__ comp, const int32_t* __restrict__ size, int64_t n,
                                      unsigned long long* __restrict__ best) {
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
        if (comp[v] == (int32_t)v)
            atomicMax(best, ((unsigned long long)(unsigned)size[v] << 32) | (unsigned long long)(unsigned)v);
    }
}

// flag[v] = node is kept; kdeg[v] = its row length if kept.  Entry n of both is 0 (scan totals).
__global__ void keep_flags_kernel(const int32_t* __restrict__ comp, const unsigned long long* __restrict__ best,
                                  const int64_t* __restrict__ indptr_s, int64_t n, int select_lcc,
                                  int32_t* __restrict__ flag, int64_t* __restrict__ kdeg) {
    const int32_t root = select_lcc ? (int32_t)(*best & 0xffffffffu) : -1;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v <= n; v += (int64_t)gridDim.x * blockDim.x) {
        if (v == n) { flag[n] = 0; kdeg[n] = 0; continue; }
        const int keep = select_lcc ? (comp[v] == root) : 1;
        flag[v] = keep;
        kdeg[v] = keep ? (indptr_s[v + 1] - indptr_s[v]) : 0;
    }
}

// Kept nodes in ascending order, their row pointers, and the two totals.
__global__ void write_rows_kernel(const int32_t* __restrict__ flag, const int32_t* __restrict__ newid,
                                  const int64_t* __restrict__ kpos, int64_t n, int32_t* __restrict__ out_keep,
                                  int64_t* __restrict__ out_indptr, int64_t* __restrict__ out_counts) {
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v <= n; v += (int64_t)gridDim.x * blockDim.x) {
        if (v == n) {
            out_indptr[newid[n]] = kpos[n];      // newid[n] = number of kept nodes, kpos[n] = kept entries
            out_counts[0] = newid[n];
            out_counts[1] = kpos[n];
        } else if (flag[v]) {
            out_keep[newid[v]] = (int32_t)v;
            out_indptr[newid[v]] = kpos[v];
        }
    }
}

// One thread per key: entries of kept rows go to their place, columns relabelled.
__global__ void write_cols_kernel(const unsigned long long* __restrict__ ukeys, const int64_t* __restrict__ m_dev,
                                  const int32_t* __restrict__ flag, const int32_t* __restrict__ newid,
                                  const int64_t* __restrict__ indptr_s, const int64_t* __restrict__ kpos,
                                  int32_t* __restrict__ out_indices) {
    const int64_t m = *m_dev;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long k = ukeys[i];
        const int32_t u = (int32_t)(k >> 32), v = (int32_t)(k & 0xffffffffu);
        if (flag[u]) out_indices[kpos[u] + (i - indptr_s[u])] = newid[v];
    }
}

struct Workspace {
    unsigned long long *keys_a, *keys_b;
    int64_t *indptr_s, *kdeg, *kpos, *scalars;   // scalars: [0] number of unique keys, [1] best component (packed)
    int32_t *parent, *comp, *size, *flag, *newid;
    void* cub_tmp;
    size_t cub_bytes;
    int64_t total;
};

size_t cub_temp_bytes(int64_t n, int64_t n_keys) {
    size_t a = 0, b = 0, c = 0, d = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, a, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, n_keys, 0, 64);
    cub::DeviceSelect::Unique(nullptr, b, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (int64_t*)nullptr, n_keys);
    cub::DeviceScan::ExclusiveSum(nullptr, c, (const int32_t*)nullptr, (int32_t*)nullptr, n + 1);
    cub::DeviceScan::ExclusiveSum(nullptr, d, (const int64_t*)nullptr, (int64_t*)nullptr, n + 1);
    size_t m = a > b ? a : b;
    if (c > m) m = c;
    if (d > m) m = d;
    return m + 256;
}

Workspace carve(void* base, int64_t n, int64_t n_keys) {
    Workspace w{};
    char* p = reinterpret_cast<char*>(base);
    int64_t off = 0;
    auto take = [&](int64_t bytes) { char* q = p ? p + off : nullptr; off += align256(bytes); return q; };
    w.keys_a = reinterpret_cast<unsigned long long*>(take(8 * (n_keys > 0 ? n_keys : 1)));
    w.keys_b = reinterpret_cast<unsigned long long*>(take(8 * (n_keys > 0 ? n_keys : 1)));
    w.indptr_s = reinterpret_cast<int64_t*>(take(8 * (n + 1)));
    w.kdeg = reinterpret_cast<int64_t*>(take(8 * (n + 1)));
    w.kpos = reinterpret_cast<int64_t*>(take(8 * (n + 1)));
    w.scalars = reinterpret_cast<int64_t*>(take(16));
    w.parent = reinterpret_cast<int32_t*>(take(4 * n));
    w.comp = reinterpret_cast<int32_t*>(take(4 * n));
    w.size = reinterpret_cast<int32_t*>(take(4 * n));
    w.flag = reinterpret_cast<int32_t*>(take(4 * (n + 1)));
    w.newid = reinterpret_cast<int32_t*>(take(4 * (n + 1)));
    w.cub_bytes = cub_temp_bytes(n, n_keys);
    w.cub_tmp = take((int64_t)w.cub_bytes);
    w.total = off;
    return w;
}

}  // namespace
}  // namespace ppnp

extern "C" {

int64_t ppnp_graph_standardize_workspace_bytes(int64_t n, int64_t nnz, int32_t flags) {
    if (n <= 0 || nnz < 0) return 0;
    const int64_t n_keys = (flags & PPNP_STD_UNDIRECTED) ? 2 * nnz : nnz;
    return ppnp::carve(nullptr, n, n_keys).total;
}

int ppnp_graph_standardize(const int64_t* indptr, const int32_t* indices, int64_t n, int64_t nnz, int32_t flags,
                           int64_t* out_indptr, int32_t* out_indices, int32_t* out_keep, int64_t* out_counts,
                           void* workspace, int64_t workspace_bytes, void* stream_) {
    using namespace ppnp;
    PPNP_REQUIRE(n > 0 && n < ((int64_t)1 << 31) && nnz >= 0, "0 < n < 2^31, nnz >= 0");
    PPNP_REQUIRE((flags & ~(PPNP_STD_UNDIRECTED | PPNP_STD_NO_SELF_LOOPS | PPNP_STD_LCC)) == 0, "unknown flags");
    PPNP_REQUIRE(indptr && (indices || nnz == 0) && out_indptr && out_indices && out_keep && out_counts, "null pointer");
    PPNP_REQUIRE(workspace && workspace_bytes >= ppnp_graph_standardize_workspace_bytes(n, nnz, flags), "workspace too small");
    cudaStream_t stream = as_stream(stream_);
    const int undirected = (flags & PPNP_STD_UNDIRECTED) ? 1 : 0;
    const int drop_loops = (flags & PPNP_STD_NO_SELF_LOOPS) ? 1 : 0;
    const int select_lcc = (flags & PPNP_STD_LCC) ? 1 : 0;
    const int64_t n_keys = undirected ? 2 * nnz : nnz;
    Workspace w = carve(workspace, n, n_keys);
    int rc;

    rc = check_cuda(cudaMemsetAsync(out_counts, 0, 3 * sizeof(int64_t), stream), "memset counts");
    if (rc) return rc;
    rc = check_cuda(cudaMemsetAsync(w.scalars, 0, 16, stream), "memset scalars");
    if (rc) return rc;
    const unsigned long long* ukeys = w.keys_a;
    if (n_keys > 0) {
        emit_keys_kernel<<<grid_for(nnz), THREADS, 0, stream>>>(indptr, indices, n, nnz, undirected, drop_loops, w.keys_a,
                                                                out_counts + 2);
        PPNP_CHECK_LAUNCH("emit_keys_kernel");
        size_t bytes = w.cub_bytes;
        rc = check_cuda(cub::DeviceRadixSort::SortKeys(w.cub_tmp, bytes, w.keys_a, w.keys_b, n_keys, 0, 64, stream), "sort keys");
        if (rc) return rc;
        bytes = w.cub_bytes;
        rc = check_cuda(cub::DeviceSelect::Unique(w.cub_tmp, bytes, w.keys_b, w.keys_a, w.scalars, n_keys, stream), "unique keys");
        if (rc) return rc;
        drop_sentinel_kernel<<<1, 1, 0, stream>>>(w.keys_a, w.scalars);
        PPNP_CHECK_LAUNCH("drop_sentinel_kernel");
    }
    row_pointers_kernel<<<grid_for(n + 1), THREADS, 0, stream>>>(ukeys, w.scalars, n, w.indptr_s);
    PPNP_CHECK_LAUNCH("row_pointers_kernel");

    if (select_lcc) {
        init_parent_kernel<<<grid_for(n), THREADS, 0, stream>>>(w.parent, w.size, n);
        PPNP_CHECK_LAUNCH("init_parent_kernel");
        if (n_keys > 0) {
            hook_kernel<<<grid_for(n_keys), THREADS, 0, stream>>>(ukeys, w.scalars, w.parent, undirected);
            PPNP_CHECK_LAUNCH("hook_kernel");
        }
        flatten_kernel<<<grid_for(n), THREADS, 0, stream>>>(w.parent, w.comp, w.size, n);
        PPNP_CHECK_LAUNCH("flatten_kernel");
        best_component_kernel<<<grid_for(n), THREADS, 0, stream>>>(w.comp, w.size, n,
                                                                   reinterpret_cast<unsigned long long*>(w.scalars + 1));
        PPNP_CHECK_LAUNCH("best_component_kernel");
    }
    keep_flags_kernel<<<grid_for(n + 1), THREADS, 0, stream>>>(w.comp, reinterpret_cast<unsigned long long*>(w.scalars + 1),
                                                               w.indptr_s, n, select_lcc, w.flag, w.kdeg);
    PPNP_CHECK_LAUNCH("keep_flags_kernel");
    size_t bytes = w.cub_bytes;
    rc = check_cuda(cub::DeviceScan::ExclusiveSum(w.cub_tmp, bytes, w.flag, w.newid, n + 1, stream), "scan kept nodes");
    if (rc) return rc;
    bytes = w.cub_bytes;
    rc = check_cuda(cub::DeviceScan::ExclusiveSum(w.cub_tmp, bytes, w.kdeg, w.kpos, n + 1, stream), "scan kept entries");
    if (rc) return rc;
    write_rows_kernel<<<grid_for(n + 1), THREADS, 0, stream>>>(w.flag, w.newid, w.kpos, n, out_keep, out_indptr, out_counts);
    PPNP_CHECK_LAUNCH("write_rows_kernel");
    if (n_keys > 0) {
        write_cols_kernel<<<grid_for(n_keys), THREADS, 0, stream>>>(ukeys, w.scalars, w.flag, w.newid, w.indptr_s, w.kpos,
                                                                    out_indices);
        PPNP_CHECK_LAUNCH("write_cols_kernel");
    }
    return PPNP_OK;
}

}  // extern "C"
