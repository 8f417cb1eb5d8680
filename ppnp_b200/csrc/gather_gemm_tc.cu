// gather_gemm_tc.cu -- bf16 tensor-core path of the dense PPR apply (placeholder entry points;
// the tcgen05 kernel lands in a later commit of this round).
#include "common.cuh"

extern "C" {

int64_t ppnp_gather_gemm_bf16_workspace_bytes(int64_t m, int64_t n, int32_t C) {
    (void)m; (void)n; (void)C;
    return 256;
}

int ppnp_gather_gemm_bf16(const void*, int64_t, const int64_t*, int64_t, int64_t, const float*, int64_t, int32_t,
                          float*, int64_t, void*, int64_t, void*) {
    ppnp::set_error("ppnp_gather_gemm_bf16: tcgen05 kernel not built into this library yet");
    return PPNP_ENOTSUP;
}

}  // extern "C"
