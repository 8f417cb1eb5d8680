// plan_build.cu -- the edge-stream plan of the propagation kernel, built on the GPU.
//
// Same stream as ppnp_b200/plan.py build_stream_plan (+ lane_transpose): the listed rows of the CSR laid end to end in
// processing order, cut into chunks of `chunk_edges` edges; bit 31 of a column word marks the last edge of a segment,
// seg_row[s] is the row a segment finishes or PPNP_FLAG | slot for a row that is cut, chunk_seg[c] the first segment
// of chunk c, fix_* the rows that are cut.  plan.py does this with ~15 eager tensor passes over int64 temporaries of
// nnz entries (0.4 s for config 4, seconds per shard of config 5); here it is four short scans over the ROWS and ONE
// pass over the edges:
//   ppnp_plan_measure : L[i] = degree of the i-th listed row -> row_start = exclusive scan (stream position of the
//                       row); from it the number of segments / partial slots / cut rows per row and their scans;
//                       totals[] = {nnz, n_segs, n_slots, n_fix, rows without an edge}.
//   ppnp_plan_fill    : one CTA per chunk.  Two binary searches find the rows of the chunk's first and last edge, every
//                       thread then finds the row of its own edge inside that short range, copies the column (and
//                       value) to its -- optionally lane-transposed -- place, flags segment ends and writes the segment's
//                       seg_row word (every segment has exactly one last edge, so every word is written exactly
//                       once); one thread per cut row fills fix_ptr / fix_row / fix_deg.
// Reference anchor: the stream is a re-encoding of the CSR that helpers.py:58-63 (calc_A_hat) produces.
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace ppnp {
namespace {

constexpr int32_t SEG_FLAG = (int32_t)0x80000000;

__global__ void __launch_bounds__(256)
plan_degrees_kernel(const int64_t* __restrict__ indptr, const int64_t* __restrict__ order, int64_t m, int64_t* __restrict__ L) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > m) return;
    if (i == m) { L[i] = 0; return; }
    const int64_t r = order ? order[i] : i;
    L[i] = indptr[r + 1] - indptr[r];
}

__global__ void __launch_bounds__(256)
plan_pieces_kernel(const int64_t* __restrict__ row_start, int64_t m, int W, int32_t* __restrict__ pieces,
                   int32_t* __restrict__ slots, int32_t* __restrict__ cut, unsigned long long* __restrict__ n_empty) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > m) return;
    if (i == m) { pieces[i] = 0; slots[i] = 0; cut[i] = 0; return; }
    const int64_t a = row_start[i], b = row_start[i + 1];
    if (b <= a) { atomicAdd(n_empty, 1ull); pieces[i] = 0; slots[i] = 0; cut[i] = 0; return; }
    const int32_t p = (int32_t)((b - 1) / W - a / W + 1);
    pieces[i] = p;
    slots[i] = p > 1 ? p : 0;
    cut[i] = p > 1 ? 1 : 0;
}

__global__ void plan_totals_kernel(const int64_t* __restrict__ row_start, const int32_t* __restrict__ seg_first,
                                   const int32_t* __restrict__ slot_first, const int32_t* __restrict__ fix_first, int64_t m,
                                   const unsigned long long* __restrict__ n_empty, int64_t* __restrict__ totals) {
    totals[0] = row_start[m];
    totals[1] = seg_first[m];
    totals[2] = slot_first[m];
    totals[3] = fix_first[m];
    totals[4] = (int64_t)*n_empty;
}

// last row i in [lo, hi] with row_start[i] <= p   (rows are non-empty, so row_start is strictly increasing)
__device__ __forceinline__ int64_t row_of_position(const int64_t* __restrict__ row_start, int64_t lo, int64_t hi, int64_t p) {
    while (lo < hi) {
        const int64_t mid = (lo + hi + 1) >> 1;
        if (__ldg(row_start + mid) <= p) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// stored place of chunk-local edge t in a stream that is lane-transposed for groups of G lanes (plan.py lane_transpose)
__device__ __forceinline__ int lane_transposed_place(int t, int G) {
    const int SR = (G >= 16) ? 1 : 16 / G;
    const int SE = SR * G, CPS = 4 / SR;
    const int j = t / SE, r = (t % SE) / G, l = t % G;
    return (j / CPS) * (CPS * SE) + l * 4 + (j % CPS) * SR + r;
}

__global__ void __launch_bounds__(256)
plan_chunks_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, const float* __restrict__ vals,
                   const int64_t* __restrict__ order, int64_t m, int W, int G, const int64_t* __restrict__ row_start,
                   const int32_t* __restrict__ seg_first, const int32_t* __restrict__ slot_first, int64_t nnz,
                   int64_t n_chunks, int32_t n_segs, int32_t* __restrict__ cols, float* __restrict__ out_vals,
                   int32_t* __restrict__ seg_row, int32_t* __restrict__ chunk_seg) {
    __shared__ int64_t s_rows[2];
    for (int64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const int64_t p0 = c * (int64_t)W;
        __syncthreads();                         // s_rows of the previous chunk has been read by everyone
        if (p0 < nnz) {
            if (threadIdx.x == 0) s_rows[0] = row_of_position(row_start, 0, m - 1, p0);
            if (threadIdx.x == 32) {
                const int64_t p1 = (p0 + W < nnz ? p0 + W : nnz) - 1;
                s_rows[1] = row_of_position(row_start, 0, m - 1, p1);
            }
        }
        __syncthreads();
        if (p0 >= nnz) {                         // padding chunk: edges that point at row 0 and are never emitted
            for (int t = threadIdx.x; t < W; t += blockDim.x) {
                cols[p0 + t] = 0;
                if (out_vals) out_vals[p0 + t] = 0.f;
            }
            if (threadIdx.x == 0) chunk_seg[c] = n_segs;
            continue;
        }
        const int64_t i0 = s_rows[0], i1 = s_rows[1];
        if (threadIdx.x == 0) chunk_seg[c] = seg_first[i0] + (int32_t)(c - __ldg(row_start + i0) / W);
        for (int t = threadIdx.x; t < W; t += blockDim.x) {
            const int64_t p = p0 + t;
            const int place = G ? lane_transposed_place(t, G) : t;
            if (p >= nnz) {
                cols[p0 + place] = 0;
                if (out_vals) out_vals[p0 + place] = 0.f;
                continue;
            }
            const int64_t i = (i0 == i1) ? i0 : row_of_position(row_start, i0, i1, p);
            const int64_t a = __ldg(row_start + i), b = __ldg(row_start + i + 1);
            const int64_t r = order ? __ldg(order + i) : i;
            const int64_t src = __ldg(indptr + r) + (p - a);
            int32_t col = __ldg(indices + src);
            if (p == b - 1 || t == W - 1) {      // last edge of its segment: the row ends here or is cut by the chunk
                col |= SEG_FLAG;
                const int32_t q = (int32_t)(c - a / W);
                const int32_t s0 = __ldg(seg_first + i), pcs = __ldg(seg_first + i + 1) - s0;
                seg_row[s0 + q] = (pcs > 1) ? (SEG_FLAG | (__ldg(slot_first + i) + q)) : (int32_t)r;
            }
            cols[p0 + place] = col;
            if (out_vals) out_vals[p0 + place] = __ldg(vals + src);
        }
    }
}

__global__ void __launch_bounds__(256)
plan_fix_kernel(const int64_t* __restrict__ order, int64_t m, const int64_t* __restrict__ row_start,
                const int32_t* __restrict__ slot_first, const int32_t* __restrict__ fix_first, const float* __restrict__ row_deg,
                int32_t* __restrict__ fix_ptr, int32_t* __restrict__ fix_row, float* __restrict__ fix_deg) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > m) return;
    if (i == m) { fix_ptr[fix_first[m]] = slot_first[m]; return; }
    const int32_t f = fix_first[i];
    if (fix_first[i + 1] == f) return;           // not a cut row
    const int64_t r = order ? order[i] : i;
    fix_ptr[f] = slot_first[i];
    fix_row[f] = (int32_t)r;
    fix_deg[f] = row_deg ? row_deg[r] : (float)(row_start[i + 1] - row_start[i]);
}

size_t scan_temp_bytes(int64_t items) {
    size_t a = 0, b = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, a, (const int64_t*)nullptr, (int64_t*)nullptr, items);
    cub::DeviceScan::ExclusiveSum(nullptr, b, (const int32_t*)nullptr, (int32_t*)nullptr, items);
    return (a > b ? a : b) + 256;
}

}  // namespace
}  // namespace ppnp

extern "C" {

int64_t ppnp_plan_workspace_bytes(int64_t n_listed) {
    if (n_listed < 0) return -1;
    return (int64_t)ppnp::scan_temp_bytes(n_listed + 1);
}

int ppnp_plan_measure(const int64_t* indptr, const int64_t* order, int64_t n_listed, int32_t chunk_edges,
                      int64_t* row_start, int32_t* seg_first, int32_t* slot_first, int32_t* fix_first,
                      int64_t* totals, void* workspace, int64_t workspace_bytes, void* stream_) {
    using namespace ppnp;
    PPNP_REQUIRE(indptr && row_start && seg_first && slot_first && fix_first && totals && workspace, "null pointer");
    PPNP_REQUIRE(n_listed > 0, "no rows listed");
    PPNP_REQUIRE(chunk_edges > 0 && chunk_edges % 128 == 0, "chunk_edges must be a positive multiple of 128");
    const int64_t items = n_listed + 1;
    PPNP_REQUIRE(workspace_bytes >= (int64_t)scan_temp_bytes(items), "workspace too small (ppnp_plan_workspace_bytes)");
    cudaStream_t stream = as_stream(stream_);
    const unsigned grid = (unsigned)((items + 255) / 256);
    // the first 8 bytes of the workspace count the rows without an edge, CUB gets the rest
    unsigned long long* n_empty = reinterpret_cast<unsigned long long*>(workspace);
    void* cub_tmp = reinterpret_cast<char*>(workspace) + 256;
    size_t cub_bytes = (size_t)workspace_bytes - 256;
    int rc = check_cuda(cudaMemsetAsync(n_empty, 0, 8, stream), "cudaMemsetAsync");
    if (rc) return rc;
    plan_degrees_kernel<<<grid, 256, 0, stream>>>(indptr, order, n_listed, row_start);
    PPNP_CHECK_LAUNCH("plan_degrees_kernel");
    rc = check_cuda(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, row_start, row_start, items, stream), "scan of the row degrees");
    if (rc) return rc;
    plan_pieces_kernel<<<grid, 256, 0, stream>>>(row_start, n_listed, chunk_edges, seg_first, slot_first, fix_first, n_empty);
    PPNP_CHECK_LAUNCH("plan_pieces_kernel");
    rc = check_cuda(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, seg_first, seg_first, items, stream), "scan of the segment counts");
    if (rc) return rc;
    rc = check_cuda(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, slot_first, slot_first, items, stream), "scan of the slot counts");
    if (rc) return rc;
    rc = check_cuda(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, fix_first, fix_first, items, stream), "scan of the cut rows");
    if (rc) return rc;
    plan_totals_kernel<<<1, 1, 0, stream>>>(row_start, seg_first, slot_first, fix_first, n_listed, n_empty, totals);
    PPNP_CHECK_LAUNCH("plan_totals_kernel");
    return PPNP_OK;
}

int ppnp_plan_fill(const int64_t* indptr, const int32_t* indices, const float* vals, const int64_t* order,
                   int64_t n_listed, int32_t chunk_edges, int32_t lane_group, const int64_t* row_start,
                   const int32_t* seg_first, const int32_t* slot_first, const int32_t* fix_first, const float* row_deg,
                   int64_t nnz, int64_t n_chunks, int64_t n_segs, int64_t n_fix, int32_t* cols, float* out_vals,
                   int32_t* seg_row, int32_t* chunk_seg, int32_t* fix_ptr, int32_t* fix_row, float* fix_deg,
                   void* stream_) {
    using namespace ppnp;
    PPNP_REQUIRE(indptr && indices && row_start && seg_first && slot_first && fix_first, "null pointer");
    PPNP_REQUIRE(cols && seg_row && chunk_seg && fix_ptr, "null output pointer");
    PPNP_REQUIRE(n_fix == 0 || (fix_row && fix_deg), "fix arrays missing");
    PPNP_REQUIRE((vals == nullptr) == (out_vals == nullptr), "vals and out_vals go together");
    PPNP_REQUIRE(n_listed > 0 && nnz > 0, "empty stream");
    PPNP_REQUIRE(chunk_edges > 0 && chunk_edges % 128 == 0, "chunk_edges must be a positive multiple of 128");
    PPNP_REQUIRE(n_chunks * (int64_t)chunk_edges >= nnz, "n_chunks too small");
    PPNP_REQUIRE(n_segs > 0 && n_segs < ((int64_t)1 << 31), "segment count out of range");
    PPNP_REQUIRE(lane_group == 0 || lane_group == 4 || lane_group == 8 || lane_group == 16 || lane_group == 32,
                 "lane group must be 0 (linear), 4, 8, 16 or 32");
    if (lane_group) {
        const int SR = lane_group >= 16 ? 1 : 16 / lane_group;
        PPNP_REQUIRE((chunk_edges / (SR * lane_group)) % (4 / SR) == 0, "chunk_edges does not hold whole staging quads for this lane group");
    }
    cudaStream_t stream = as_stream(stream_);
    const int64_t cap = (int64_t)sm_count() * 32;
    plan_chunks_kernel<<<(unsigned)(n_chunks < cap ? n_chunks : cap), 256, 0, stream>>>(
        indptr, indices, vals, order, n_listed, chunk_edges, lane_group, row_start, seg_first, slot_first, nnz, n_chunks,
        (int32_t)n_segs, cols, out_vals, seg_row, chunk_seg);
    PPNP_CHECK_LAUNCH("plan_chunks_kernel");
    plan_fix_kernel<<<(unsigned)((n_listed + 1 + 255) / 256), 256, 0, stream>>>(order, n_listed, row_start, slot_first, fix_first,
                                                                              row_deg, fix_ptr, fix_row, fix_deg);
    PPNP_CHECK_LAUNCH("plan_fix_kernel");
    return PPNP_OK;
}

}  // extern "C"
