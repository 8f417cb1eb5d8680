// rmat.cu -- synthetic-workload plumbing for BASELINE.json configs 4/5 (not reference code):
// R-MAT raw draws as 64-bit (src << 32 | dst) keys, both directions, ready for sort + unique.
#include "common.cuh"
#include "../../include/ppnp_rmat.h"

namespace ppnp {
namespace {
__global__ void rmat_keys_kernel(uint64_t seed, int scale, int64_t n, int64_t e0, int64_t count, int64_t* __restrict__ keys) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        uint32_t s, d;
        ppnp_rmat_edge(seed, (uint64_t)(e0 + i), scale, &s, &d);
        int64_t k0 = -1, k1 = -1;
        if (s != d && (int64_t)s < n && (int64_t)d < n) {
            k0 = ((int64_t)s << 32) | (int64_t)d;
            k1 = ((int64_t)d << 32) | (int64_t)s;
        }
        keys[2 * i] = k0;
        keys[2 * i + 1] = k1;
    }
}
}  // namespace
}  // namespace ppnp

extern "C" int ppnp_rmat_keys(uint64_t seed, int32_t scale, int64_t n, int64_t e0, int64_t e1, int64_t* out_keys, void* stream) {
    using namespace ppnp;
    PPNP_REQUIRE(out_keys && e1 >= e0 && scale > 0 && scale <= 31 && n > 0, "bad arguments");
    const int64_t count = e1 - e0;
    if (count == 0) return PPNP_OK;
    int64_t blocks = (count + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 32;
    if (blocks > cap) blocks = cap;
    rmat_keys_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(seed, scale, n, e0, count, out_keys);
    PPNP_CHECK_LAUNCH("rmat_keys_kernel");
    return PPNP_OK;
}
