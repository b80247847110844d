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

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL), opt-in with EEGX_PDL=1.  Every libeegx kernel goes through
// eegx::launch() and starts with EEGX_PDL_SYNC(): `launch_dependents` lets the NEXT kernel's CTAs be
// scheduled while this grid is still running, `wait` holds this grid until everything it depends on has
// completed and flushed -- results are identical to plain stream order, only the launch gaps overlap.
// Measured on the B = 256 train step (4 region streams, whole step in one CUDA graph): 30.3-30.4 ms with
// the attribute set against 29.7-29.8 ms without -- the early-scheduled CTAs of one stream take SM slots
// from the kernels the other streams have in flight, which costs more than the ~1-2 us graph edges it
// hides.  So the attribute is OFF by default (the two instructions are no-ops then); the switch stays for
// single-stream callers.
namespace eegx {
bool pdl_enabled();
}
#ifdef __CUDACC__
#include <utility>
#ifdef EEGX_NO_GRIDDEP            // A/B build without the two instructions
#define EEGX_PDL_SYNC()
#define EEGX_PDL_TRIGGER()
#define EEGX_PDL_WAIT()
#else
#define EEGX_PDL_SYNC() asm volatile("griddepcontrol.launch_dependents;\n\tgriddepcontrol.wait;" ::: "memory")
#define EEGX_PDL_TRIGGER() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
#define EEGX_PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
#endif
namespace eegx {
template <typename... KArgs, typename... Args>
inline cudaError_t launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
}  // namespace eegx
#endif

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
