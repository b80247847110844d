// Host-side helpers shared by every translation unit of libeegx.so.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "eegx.h"

namespace eegx {

// Thread-local message returned by eegx_last_error().
char* error_buffer();
int set_error(int code, const char* fmt, ...);

// 0 when the current device is compute capability 10.x, else EEGX_ERR_ARCH.
int require_sm100();

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

constexpr int kNumSMsB200 = 148;

}  // namespace eegx

#define EEGX_CUDA_CHECK(expr)                                                          \
    do {                                                                               \
        cudaError_t _e = (expr);                                                       \
        if (_e != cudaSuccess)                                                         \
            return eegx::set_error(EEGX_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,      \
                                   cudaGetErrorString(_e), __FILE__, __LINE__);        \
    } while (0)

#define EEGX_REQUIRE(cond, code, ...)                          \
    do {                                                       \
        if (!(cond)) return eegx::set_error(code, __VA_ARGS__); \
    } while (0)
