// Version, error reporting and the architecture gate of libeegx.so.
#include <stdlib.h>

#include "eegx_common.h"

namespace eegx {

char* error_buffer() {
    static thread_local char buf[512] = {0};
    return buf;
}

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(error_buffer(), 512, fmt, ap);
    va_end(ap);
    return code;
}

bool pdl_enabled() {
    static const bool on = [] { const char* v = getenv("EEGX_PDL"); return v != nullptr && atoi(v) != 0; }();
    return on;
}

int require_sm100() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess)
        return set_error(EEGX_ERR_CUDA, "cudaGetDevice failed: %s", cudaGetErrorString(e));
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess)
        return set_error(EEGX_ERR_CUDA, "cudaDeviceGetAttribute failed: %s", cudaGetErrorString(e));
    if (major != 10)
        return set_error(EEGX_ERR_ARCH,
                         "libeegx is built for sm_100a only; device %d has compute capability %d.x "
                         "(there is no fallback path)", dev, major);
    return EEGX_OK;
}

}  // namespace eegx

extern "C" {

int eegx_version(void) { return EEGX_VERSION; }

const char* eegx_last_error(void) { return eegx::error_buffer(); }

int eegx_device_check(void) { return eegx::require_sm100(); }

}  // extern "C"
