// Device helpers shared by the tuned DSP kernels (dsp_long.cu; dsp_tuned.cu keeps its own copies of the same
// few functions): complex arithmetic in registers, the 8-point FFT, MUFU log2, 1-D TMA bulk loads + mbarrier.
#pragma once
#include <cuda_runtime.h>

namespace eegx_dsp {

struct cf { float r, i; };
__device__ __forceinline__ cf cadd(cf a, cf b) { return {a.r + b.r, a.i + b.i}; }
__device__ __forceinline__ cf csub(cf a, cf b) { return {a.r - b.r, a.i - b.i}; }
__device__ __forceinline__ cf cmul(cf a, float wr, float wi) {
    return {fmaf(a.r, wr, -a.i * wi), fmaf(a.r, wi, a.i * wr)};
}
__device__ __forceinline__ cf mul_neg_i(cf a) { return {a.i, -a.r}; }   // a * (-i)

constexpr float RSQRT2 = 0.70710678118654752440f;

// In-place forward 8-point FFT (e^{-i...}), natural order in and out.
__device__ __forceinline__ void fft8(cf (&v)[8]) {
    cf e0 = cadd(v[0], v[4]), e1 = cadd(v[1], v[5]), e2 = cadd(v[2], v[6]), e3 = cadd(v[3], v[7]);
    cf d0 = csub(v[0], v[4]), d1 = csub(v[1], v[5]), d2 = csub(v[2], v[6]), d3 = csub(v[3], v[7]);
    cf o0 = d0;
    cf o1 = {(d1.r + d1.i) * RSQRT2, (d1.i - d1.r) * RSQRT2};     // * W8^1
    cf o2 = mul_neg_i(d2);                                         // * W8^2
    cf o3 = {(d3.i - d3.r) * RSQRT2, -(d3.r + d3.i) * RSQRT2};    // * W8^3
    cf s0 = cadd(e0, e2), s1 = csub(e0, e2), s2 = cadd(e1, e3), s3 = mul_neg_i(csub(e1, e3));
    v[0] = cadd(s0, s2); v[4] = csub(s0, s2); v[2] = cadd(s1, s3); v[6] = csub(s1, s3);
    cf t0 = cadd(o0, o2), t1 = csub(o0, o2), t2 = cadd(o1, o3), t3 = mul_neg_i(csub(o1, o3));
    v[1] = cadd(t0, t2); v[5] = csub(t0, t2); v[3] = cadd(t1, t3); v[7] = csub(t1, t3);
}

__device__ __forceinline__ float fast_log2(float x) {   // x >= 4*log_eps > 0: no denormal path
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra.uni WAIT_DONE;\n\t"
        "bra.uni WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

}  // namespace eegx_dsp
