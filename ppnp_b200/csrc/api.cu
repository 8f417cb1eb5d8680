// api.cu -- error string, version and device queries of the C ABI (include/ppnp_b200.h).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace ppnp {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static thread_local int cached_dev = -1;
    static thread_local int cached_sms = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
        cached_dev = dev;
        cached_sms = sms;
    }
    return cached_sms;
}

}  // namespace ppnp

extern "C" {

const char* ppnp_last_error(void) { return ppnp::g_err; }

int ppnp_version(void) { return 100; }

int ppnp_device_info(int32_t* sms, int32_t* cc_major, int32_t* cc_minor, int64_t* l2_bytes) {
    int dev = 0;
    int rc = ppnp::check_cuda(cudaGetDevice(&dev), "cudaGetDevice");
    if (rc) return rc;
    int v = 0;
    if (sms) { rc = ppnp::check_cuda(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev), "attr sm count"); if (rc) return rc; *sms = v; }
    if (cc_major) { rc = ppnp::check_cuda(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev), "attr cc major"); if (rc) return rc; *cc_major = v; }
    if (cc_minor) { rc = ppnp::check_cuda(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev), "attr cc minor"); if (rc) return rc; *cc_minor = v; }
    if (l2_bytes) { rc = ppnp::check_cuda(cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, dev), "attr l2"); if (rc) return rc; *l2_bytes = v; }
    return PPNP_OK;
}

}  // extern "C"
