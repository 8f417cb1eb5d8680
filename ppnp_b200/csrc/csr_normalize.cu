// csr_normalize.cu -- A_hat = D^-1/2 (A+I) D^-1/2 (or D^-1 (A+I)) as GPU kernels.
//
// Replaces helpers.py:58-66 calc_A_hat of the reference (scipy on the host):
//   A = adj + sp.eye(n)                    -> row_count_kernel + scan + fill_kernel (structure)
//   D = np.sum(A, axis=1).A1               -> row_count_kernel (fp64, same summation order)
//   D_inv @ A @ D_inv / D_inv @ A          -> fill_kernel, fp64 products in the reference's order
// Structure (indptr, indices) and D are bit-exact with scipy; fp64 values are bit-exact too
// (one IEEE product per factor, no contraction possible); fp32 values are that fp64 value
// rounded once.
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace ppnp {
namespace {

__device__ __forceinline__ int lower_bound_i32(const int32_t* __restrict__ a, int lo, int hi, int key) {
    while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        if (__ldg(a + mid) < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// One thread per row: entries of row i of adj + I, and D_i.
__global__ void row_count_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                 const float* __restrict__ data, int64_t n, int32_t* __restrict__ cnt,
                                 double* __restrict__ deg) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == n) { cnt[n] = 0; return; }
    const int b = indptr[i], e = indptr[i + 1];
    const int p = lower_bound_i32(indices, b, e, (int)i);
    const bool has_diag = (p < e) && (__ldg(indices + p) == (int)i);
    cnt[i] = (e - b) + (has_diag ? 0 : 1);
    double d;
    if (data == nullptr) {
        d = (double)(e - b) + 1.0;  // unit weights: integers, any order is exact
    } else {
        // scipy sums the merged row left to right (CSR mat-vec with ones)
        d = 0.0;
        for (int t = b; t < p; ++t) d += (double)__ldg(data + t);
        if (has_diag) { d += (double)__ldg(data + p) + 1.0; } else { d += 1.0; }
        for (int t = p + (has_diag ? 1 : 0); t < e; ++t) d += (double)__ldg(data + t);
    }
    deg[i] = d;
}

// One warp per row: write the merged row, its values and D^-1/2.
__global__ void fill_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                            const float* __restrict__ data, int64_t n, int mode,
                            const int32_t* __restrict__ out_indptr, int32_t* __restrict__ out_indices,
                            const double* __restrict__ deg, double* __restrict__ out_val64,
                            float* __restrict__ out_val32, float* __restrict__ out_dinv) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    const int b = indptr[i], e = indptr[i + 1];
    const int ob = out_indptr[i];
    const int p = lower_bound_i32(indices, b, e, (int)i);
    const bool has_diag = (p < e) && (__ldg(indices + p) == (int)i);
    const double di = deg[i];
    const double si = (mode == PPNP_MODE_SYM) ? 1.0 / sqrt(di) : 1.0 / di;
    const bool want_val = (out_val64 != nullptr) || (out_val32 != nullptr);
    if (lane == 0 && out_dinv != nullptr) out_dinv[i] = (float)si;
    for (int t = b + lane; t < e; t += 32) {
        const int j = __ldg(indices + t);
        double a = data ? (double)__ldg(data + t) : 1.0;
        int o;
        if (has_diag) { o = ob + (t - b); if (t == p) a += 1.0; }
        else          { o = ob + (t - b) + (t >= p ? 1 : 0); }
        out_indices[o] = j;
        if (want_val) {
            double v;
            if (mode == PPNP_MODE_SYM) { const double sj = 1.0 / sqrt(__ldg(deg + j)); v = (si * a) * sj; }
            else { v = si * a; }
            if (out_val64) out_val64[o] = v;
            if (out_val32) out_val32[o] = (float)v;
        }
    }
    if (!has_diag && lane == 0) {
        const int o = ob + (p - b);
        out_indices[o] = (int)i;
        if (want_val) {
            const double v = (mode == PPNP_MODE_SYM) ? (si * 1.0) * si : si * 1.0;
            if (out_val64) out_val64[o] = v;
            if (out_val32) out_val32[o] = (float)v;
        }
    }
}

inline int64_t align256(int64_t x) { return (x + 255) & ~(int64_t)255; }

}  // namespace
}  // namespace ppnp

extern "C" {

int64_t ppnp_csr_normalize_workspace_bytes(int64_t n) {
    size_t cub_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, (const int32_t*)nullptr, (int32_t*)nullptr, (int)(n + 1));
    return ppnp::align256(4 * (n + 1)) + ppnp::align256((int64_t)cub_bytes) + 256;
}

int ppnp_csr_normalize(const int32_t* indptr, const int32_t* indices, const float* data, int64_t n,
                       int64_t nnz, int32_t mode, int32_t* out_indptr, int32_t* out_indices,
                       double* out_deg, double* out_val64, float* out_val32, float* out_dinv,
                       void* workspace, int64_t workspace_bytes, void* stream_) {
    using namespace ppnp;
    PPNP_REQUIRE(n > 0 && nnz >= 0, "n > 0, nnz >= 0");
    PPNP_REQUIRE(nnz + n < ((int64_t)1 << 31), "nnz + n must fit int32 (shard larger graphs)");
    PPNP_REQUIRE(indptr && (indices || nnz == 0) && out_indptr && out_indices && out_deg, "null pointer");
    PPNP_REQUIRE(mode == PPNP_MODE_SYM || mode == PPNP_MODE_RW, "mode must be sym or rw");
    PPNP_REQUIRE(workspace && workspace_bytes >= ppnp_csr_normalize_workspace_bytes(n), "workspace too small");
    cudaStream_t stream = as_stream(stream_);
    int32_t* cnt = reinterpret_cast<int32_t*>(workspace);
    void* cub_tmp = reinterpret_cast<char*>(workspace) + align256(4 * (n + 1));
    size_t cub_bytes = (size_t)(workspace_bytes - align256(4 * (n + 1)));

    const int threads = 256;
    row_count_kernel<<<(unsigned)((n + 1 + threads - 1) / threads), threads, 0, stream>>>(indptr, indices, data, n, cnt, out_deg);
    PPNP_CHECK_LAUNCH("row_count_kernel");
    int rc = check_cuda(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, cnt, out_indptr, (int)(n + 1), stream), "scan");
    if (rc) return rc;
    const int64_t warps_per_block = threads / 32;
    fill_kernel<<<(unsigned)((n + warps_per_block - 1) / warps_per_block), threads, 0, stream>>>(
        indptr, indices, data, n, mode, out_indptr, out_indices, out_deg, out_val64, out_val32, out_dinv);
    PPNP_CHECK_LAUNCH("fill_kernel");
    return PPNP_OK;
}

}  // extern "C"
