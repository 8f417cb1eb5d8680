// dsmem_probe.cu -- DESIGN TOOL (not product code): how many bytes per clock can an SM gather as random 256-byte rows
// (a) from an L2-resident table, (b) from the shared memory of the other CTAs of its thread-block cluster (DSMEM),
// (c) from a mix of both?  The question behind csrc/appnp_cluster.cu: the row-major SpMM of config 4 sits on the
// L2 -> SM throughput cap (~6300 B/clk chip-wide); rows served over the SM-to-SM network do not cross that fabric.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_build/dsmem_probe tools/dsmem_probe.cu
//   tools/_build/dsmem_probe            (prints one line per configuration)
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int NT = 1024;
constexpr int RING = 8;

__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

__device__ __forceinline__ float4 ld_dsmem(uint32_t local_saddr, uint32_t rank) {
    uint32_t a;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"(local_saddr), "r"(rank));
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ float4 ld_gmem(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// MODE 0: branch per load (two instructions, predicated); MODE 1: generic pointers for both spaces
template <int MODE>
__global__ void __launch_bounds__(NT, 1)
probe_kernel(const float* __restrict__ table, uint32_t table_rows, uint32_t smem_rows, int hot_pct, int local_too, int iters, float* out) {
    extern __shared__ __align__(16) float hot[];
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t csz = cluster.num_blocks();
    const uint32_t my_rank = cluster.block_rank();
    for (uint32_t i = threadIdx.x; i < smem_rows * 64; i += NT) hot[i] = (float)(i & 1023) * 1e-3f;
    cluster.sync();
    const int lane = threadIdx.x & 31, lg = lane & 15, g = lane >> 4;
    uint32_t seed = (blockIdx.x * NT + threadIdx.x) / 16 * 2654435761u + 12345u;   // same stream for the 16 lanes of a group
    (void)g;
    const uint32_t hot_base = (uint32_t)__cvta_generic_to_shared(hot) + lg * 16;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int it = 0; it < iters; ++it) {
        float4 v[RING];
#pragma unroll
        for (int e = 0; e < RING; ++e) {
            const uint32_t r = lcg(seed);
            const bool is_hot = (int)(r % 100u) < hot_pct;
            const uint32_t r2 = lcg(seed);
            if (MODE == 0) {
                if (is_hot) {
                    uint32_t owner = r2 % csz;
                    if (!local_too && csz > 1 && owner == my_rank) owner = (owner + 1) % csz;
                    v[e] = ld_dsmem(hot_base + ((r2 >> 4) % smem_rows) * 256u, owner);
                } else {
                    v[e] = ld_gmem(table + (size_t)(r2 % table_rows) * 64 + lg * 4);
                }
            } else {
                const float* p;
                if (is_hot) {
                    uint32_t owner = r2 % csz;
                    if (!local_too && csz > 1 && owner == my_rank) owner = (owner + 1) % csz;
                    p = cluster.map_shared_rank(hot + ((r2 >> 4) % smem_rows) * 64 + lg * 4, owner);
                } else {
                    p = table + (size_t)(r2 % table_rows) * 64 + lg * 4;
                }
                v[e] = *reinterpret_cast<const float4*>(p);
            }
        }
#pragma unroll
        for (int e = 0; e < RING; ++e) { acc.x += v[e].x; acc.y += v[e].y; acc.z += v[e].z; acc.w += v[e].w; }
    }
    cluster.sync();   // nobody leaves while a peer may still read its shared memory
    if (acc.x + acc.y + acc.z + acc.w == 123.456f) out[0] = acc.x;
}

template <int MODE>
static float run(int csz, int smem_kb, int hot_pct, int local_too, int iters, const float* table, uint32_t table_rows, float* out, int sms) {
    auto k = probe_kernel<MODE>;
    const int smem_bytes = smem_kb * 1024;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    if (csz > 8) CK(cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    int grid = (sms / csz) * csz;
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = smem_bytes;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = csz; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int max_clusters = 0;
    CK(cudaOccupancyMaxActiveClusters(&max_clusters, k, &cfg));
    if (max_clusters * csz < grid) { grid = max_clusters * csz; cfg.gridDim = dim3(grid); }
    const uint32_t smem_rows = smem_bytes / 256;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 2; ++w) CK(cudaLaunchKernelEx(&cfg, k, table, table_rows, smem_rows, hot_pct, local_too, iters, out));
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    const int reps = 5;
    for (int w = 0; w < reps; ++w) CK(cudaLaunchKernelEx(&cfg, k, table, table_rows, smem_rows, hot_pct, local_too, iters, out));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
    const double rows = (double)grid * (NT / 16) * (double)iters * RING;
    const double bytes = rows * 256.0;
    const double tbs = bytes / (ms * 1e-3) / 1e12;
    printf("mode %d cluster %2d smem %3d KB hot %3d%% local_too %d grid %3d: %.3f ms  %.2f TB/s  %.1f B/clk/SM @1.93GHz (%.1f from DSMEM, %.1f from L2)\n",
           MODE, csz, smem_kb, hot_pct, local_too, grid, ms, tbs, bytes / (ms * 1e-3) / 1.93e9 / grid,
           bytes * hot_pct / 100.0 / (ms * 1e-3) / 1.93e9 / grid, bytes * (100 - hot_pct) / 100.0 / (ms * 1e-3) / 1.93e9 / grid);
    fflush(stdout);
    return ms;
}

int main(int argc, char** argv) {
    int dev = 0;
    CK(cudaSetDevice(dev));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const uint32_t table_rows = (argc > 1) ? (uint32_t)atoi(argv[1]) : 131072;   // x 256 B = 32 MB: L2-resident
    float *table, *out;
    CK(cudaMalloc(&table, (size_t)table_rows * 256));
    CK(cudaMemset(table, 0, (size_t)table_rows * 256));
    CK(cudaMalloc(&out, 16));
    const int iters = 400;
    printf("SMs %d, table %u rows x 256 B\n", sms, table_rows);
    // (a) L2 only, by cluster size (placement changes with clusters)
    for (int csz : {1, 8}) run<0>(csz, 16, 0, 1, iters, table, table_rows, out, sms);
    // (b) DSMEM only
    for (int csz : {2, 4, 8, 16}) run<0>(csz, 192, 100, 0, iters, table, table_rows, out, sms);
    run<0>(8, 192, 100, 1, iters, table, table_rows, out, sms);
    run<1>(8, 192, 100, 0, iters, table, table_rows, out, sms);
    // (c) mixes
    for (int csz : {4, 8, 16})
        for (int pct : {15, 25, 35, 45}) run<0>(csz, 192, pct, 1, iters, table, table_rows, out, sms);
    for (int pct : {25, 35}) run<1>(8, 192, pct, 1, iters, table, table_rows, out, sms);
    return 0;
}
