// Device helpers shared by the fused glue kernels (fused_rowwise.cu, fused_bn.cu, attn_small.cu):
// bf16 <-> fp32 vector access, exact GELU and its derivative, counter-based dropout.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace eegx {

// ---------------------------------------------------------------- bf16 x 8 (16-byte) vectors
struct alignas(16) bf16x8 {
    __nv_bfloat162 h[4];
};

__device__ __forceinline__ void unpack8(const bf16x8& v, float (&f)[8]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = __bfloat1622float2(v.h[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}

__device__ __forceinline__ bf16x8 pack8(const float (&f)[8]) {
    bf16x8 v;
#pragma unroll
    for (int i = 0; i < 4; ++i) v.h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return v;
}

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&f)[8]) {
    unpack8(*reinterpret_cast<const bf16x8*>(p), f);
}

__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&f)[8]) {
    *reinterpret_cast<bf16x8*>(p) = pack8(f);
}

__device__ __forceinline__ void load8f(const float* p, float (&f)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// ---------------------------------------------------------------- activations (exact erf GELU)
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

__device__ __forceinline__ float gelu_grad_f(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
    const float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
    return fmaf(x, pdf, cdf);
}

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + __expf(-x)); }

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------- dropout
// Philox-4x32-10 keyed by (seed, step) read from device memory (so a CUDA graph replays with
// fresh masks after the step counter is bumped in-graph) and by the call site; the counter is the
// index of an 8-element group, and one call yields eight 16-bit uniforms (keep iff u16 >= p*65536).
// Forward and backward regenerate the same mask from (site, group index): no mask tensor in HBM.
struct DropoutCfg {
    const unsigned long long* state;   // [0] = seed, [1] = step; nullptr or p == 0 -> no dropout
    unsigned int site;
    float p;
};

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    constexpr unsigned int M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned int hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const unsigned int hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}

struct DropoutGen {
    uint2 key;
    unsigned int site, thresh;
    float scale;
    bool on;
    __device__ __forceinline__ DropoutGen(const DropoutCfg& c) {
        on = c.state != nullptr && c.p > 0.0f;
        site = c.site;
        thresh = 0;
        scale = 1.0f;
        key = make_uint2(0u, 0u);
        if (on) {
            const unsigned long long seed = c.state[0], step = c.state[1];
            key = make_uint2((unsigned int)seed ^ (unsigned int)(step * 0x9E3779B97F4A7C15ull >> 32),
                             (unsigned int)(seed >> 32) ^ (unsigned int)step);
            thresh = (unsigned int)(c.p * 65536.0f + 0.5f);
            scale = 1.0f / (1.0f - c.p);
        }
    }
    // multipliers (0 or 1/(1-p)) for the 8 elements of group `g`
    __device__ __forceinline__ void mask8(unsigned long long g, float (&m)[8]) const {
        if (!on) {
#pragma unroll
            for (int i = 0; i < 8; ++i) m[i] = 1.0f;
            return;
        }
        const uint4 r = philox4x32_10(make_uint4((unsigned int)g, (unsigned int)(g >> 32), site, 0x51ED270Bu), key);
        const unsigned int w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            m[2 * i] = (w[i] & 0xFFFFu) >= thresh ? scale : 0.0f;
            m[2 * i + 1] = (w[i] >> 16) >= thresh ? scale : 0.0f;
        }
    }
    // multipliers of elements idx, idx + 1 (idx even) of group `g`
    __device__ __forceinline__ void mask_pair(unsigned long long g, int idx, float& m0, float& m1) const {
        if (!on) {
            m0 = m1 = 1.0f;
            return;
        }
        const uint4 r = philox4x32_10(make_uint4((unsigned int)g, (unsigned int)(g >> 32), site, 0x51ED270Bu), key);
        const unsigned int w = idx < 4 ? (idx < 2 ? r.x : r.y) : (idx < 6 ? r.z : r.w);
        m0 = (w & 0xFFFFu) >= thresh ? scale : 0.0f;
        m1 = (w >> 16) >= thresh ? scale : 0.0f;
    }
    // multiplier of the single element idx of group `g`
    __device__ __forceinline__ float mask_one(unsigned long long g, int idx) const {
        float m0, m1;
        mask_pair(g, idx & ~1, m0, m1);
        return (idx & 1) ? m1 : m0;
    }
};

}  // namespace eegx
