// common.cuh -- error plumbing and small device helpers shared by every kernel file.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ppnp_b200.h"

namespace ppnp {

void set_error(const char* fmt, ...);

inline int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return PPNP_OK;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return PPNP_ECUDA;
}

#define PPNP_CHECK_LAUNCH(name)                                             \
    do {                                                                    \
        int rc__ = ::ppnp::check_cuda(cudaGetLastError(), "launch " name);  \
        if (rc__ != PPNP_OK) return rc__;                                   \
    } while (0)

#define PPNP_REQUIRE(cond, msg)                                             \
    do {                                                                    \
        if (!(cond)) {                                                      \
            ::ppnp::set_error("%s: requirement failed: %s", __func__, msg); \
            return PPNP_EINVAL;                                             \
        }                                                                   \
    } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();

// Streaming (evict-first) accesses -- __ldcs / __stcs -- are used for index, teleport and output
// streams so that they do not push the gathered Z rows out of the L2; see Vec<>::load_stream.
template <int N>
struct Vec;
template <>
struct Vec<1> {
    float x;
    __device__ __forceinline__ void zero() { x = 0.f; }
    __device__ __forceinline__ void add(const Vec& o) { x += o.x; }
    __device__ __forceinline__ void fma(float w, const Vec& o) { x = fmaf(w, o.x, x); }
    __device__ __forceinline__ static Vec load(const float* p) { Vec v; v.x = __ldg(p); return v; }
    __device__ __forceinline__ static Vec load_plain(const float* p) { Vec v; v.x = *p; return v; }
    __device__ __forceinline__ static Vec load_cg(const float* p) { Vec v; v.x = __ldcg(p); return v; }
    __device__ __forceinline__ static Vec load_stream(const float* p) { Vec v; v.x = __ldcs(p); return v; }
    __device__ __forceinline__ void store(float* p) const { *p = x; }
    __device__ __forceinline__ void store_stream(float* p) const { __stcs(p, x); }
    __device__ __forceinline__ static Vec axpby(float a, const Vec& u, float b, const Vec& v) {
        Vec r; r.x = fmaf(a, u.x, b * v.x); return r;
    }
    __device__ __forceinline__ static Vec shfl_xor_add(const Vec& a, int o) {
        Vec r; r.x = a.x + __shfl_xor_sync(0xffffffffu, a.x, o); return r;
    }
};
template <>
struct Vec<4> {
    float4 v;
    __device__ __forceinline__ void zero() { v = make_float4(0.f, 0.f, 0.f, 0.f); }
    __device__ __forceinline__ void add(const Vec& o) { v.x += o.v.x; v.y += o.v.y; v.z += o.v.z; v.w += o.v.w; }
    __device__ __forceinline__ void fma(float w, const Vec& o) {
        v.x = fmaf(w, o.v.x, v.x); v.y = fmaf(w, o.v.y, v.y); v.z = fmaf(w, o.v.z, v.z); v.w = fmaf(w, o.v.w, v.w);
    }
    __device__ __forceinline__ static Vec load(const float* p) { Vec r; r.v = __ldg(reinterpret_cast<const float4*>(p)); return r; }
    __device__ __forceinline__ static Vec load_plain(const float* p) { Vec r; r.v = *reinterpret_cast<const float4*>(p); return r; }
    __device__ __forceinline__ static Vec load_cg(const float* p) { Vec r; r.v = __ldcg(reinterpret_cast<const float4*>(p)); return r; }
    __device__ __forceinline__ static Vec load_stream(const float* p) { Vec r; r.v = __ldcs(reinterpret_cast<const float4*>(p)); return r; }
    __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = v; }
    __device__ __forceinline__ void store_stream(float* p) const { __stcs(reinterpret_cast<float4*>(p), v); }
    __device__ __forceinline__ static Vec axpby(float a, const Vec& u, float b, const Vec& w) {
        Vec r;
        r.v.x = fmaf(a, u.v.x, b * w.v.x); r.v.y = fmaf(a, u.v.y, b * w.v.y);
        r.v.z = fmaf(a, u.v.z, b * w.v.z); r.v.w = fmaf(a, u.v.w, b * w.v.w);
        return r;
    }
    __device__ __forceinline__ static Vec shfl_xor_add(const Vec& a, int o) {
        Vec r;
        r.v.x = a.v.x + __shfl_xor_sync(0xffffffffu, a.v.x, o); r.v.y = a.v.y + __shfl_xor_sync(0xffffffffu, a.v.y, o);
        r.v.z = a.v.z + __shfl_xor_sync(0xffffffffu, a.v.z, o); r.v.w = a.v.w + __shfl_xor_sync(0xffffffffu, a.v.w, o);
        return r;
    }
};

// epilogue coefficients, out = a * acc + b * teleport   (include/ppnp_b200.h PPNP_EPI_*)
__device__ __forceinline__ void epi_coef(int epi, float alpha, float deg, float& a, float& b) {
    const float oma = 1.0f - alpha;
    switch (epi & 15) {
        default:
        case PPNP_EPI_PLAIN: a = oma; b = alpha; break;
        // rsqrtf / __frcp_rn: one MUFU each (<= 2 ulp / correctly rounded); deg is a small integer
        case PPNP_EPI_Z2Y: { const float d = rsqrtf(deg); a = oma * d; b = alpha * d; } break;
        case PPNP_EPI_Y: a = oma * __frcp_rn(deg); b = alpha * rsqrtf(deg); break;
        case PPNP_EPI_Y2Z: a = oma * rsqrtf(deg); b = alpha; break;
        case PPNP_EPI_RW: a = oma * __frcp_rn(deg); b = alpha; break;
        case PPNP_EPI_Y02Z: a = oma * rsqrtf(deg); b = alpha * sqrtf(deg); break;
    }
    if (epi & PPNP_EPI_ACC) b = 1.0f;   // T is the output itself: add to what an earlier pass wrote
}

// Epilogue and value use of step k of K (1-based) of a propagation, shared by every K-step entry point.
inline void step_form(int mode, int use_vals, int k, int K, int& epi, int& vals) {
    if (mode == PPNP_MODE_SYM_Y0) { epi = (k == K) ? PPNP_EPI_Y02Z : PPNP_EPI_RW; vals = 0; }
    else if (use_vals) { epi = PPNP_EPI_PLAIN; vals = 1; }
    else if (mode == PPNP_MODE_RW) { epi = PPNP_EPI_RW; vals = 0; }
    else if (K == 1) { epi = PPNP_EPI_PLAIN; vals = 1; }
    else if (k == 1) { epi = PPNP_EPI_Z2Y; vals = 1; }
    else if (k == K) { epi = PPNP_EPI_Y2Z; vals = 0; }
    else { epi = PPNP_EPI_Y; vals = 0; }
}

}  // namespace ppnp
