// Fused global-norm gradient clipping + AdamW over flat fp32 buffers (HBM-bound).
//
// Replaces, for one optimizer step of EEGTrainer.train_epoch
// (main_model/src/training/trainer.py:101-113, optimizer wiring scripts/train.py:199-241):
//   torch.nn.utils.clip_grad_norm_(params, max_norm)   ->  eegx_sumsq_f32 (+ the coefficient
//                                                          computed on the device, no host sync)
//   AdamW.step()                                        ->  eegx_adamw_clip_f32
// AdamW follows torch.optim.AdamW (decoupled decay applied first, eps added after the bias-
// corrected sqrt); the reference's transformers.AdamW is removed upstream.
// Algorithmic bytes per parameter: 4 (grad-norm read) + 16 read + 12 written.
#include <cuda_bf16.h>

#include "eegx_common.h"

namespace {

constexpr int NT = 256;
constexpr int MAX_PARTIALS = 148 * 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// stage 1: one partial per CTA (fixed grid => fixed summation order => bit-stable)
__global__ void __launch_bounds__(NT) sumsq_partial_kernel(const float* __restrict__ g, long long n,
                                                           double* __restrict__ partials) {
    EEGX_PDL_SYNC();
    __shared__ float red[NT / 32];
    float acc = 0.0f;
    const long long n4 = n >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n4; i += (long long)gridDim.x * NT) {
        const float4 v = __ldg(g4 + i);
        acc = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, acc))));
    }
    if (blockIdx.x == 0)
        for (long long i = (n4 << 2) + threadIdx.x; i < n; i += NT) acc = fmaf(g[i], g[i], acc);
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = threadIdx.x < NT / 32 ? red[threadIdx.x] : 0.0f;
        t = warp_sum(t);
        if (threadIdx.x == 0) partials[blockIdx.x] = (double)t;
    }
}

// stage 2: single warp, fixed order, fp64; out[0] (+)= sum
__global__ void sumsq_final_kernel(const double* __restrict__ partials, int count, float* out, int accumulate) {
    EEGX_PDL_SYNC();
    double acc = 0.0;
    for (int i = threadIdx.x; i < count; i += 32) acc += partials[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) out[0] = (accumulate ? out[0] : 0.0f) + (float)acc;
}

__global__ void __launch_bounds__(NT)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             long long n, float lr, float beta1, float beta2, float eps, float wd, float bc1, float bc2_sqrt,
             const float* __restrict__ norm_sq, float max_norm, float grad_scale, __nv_bfloat16* __restrict__ w16) {
    EEGX_PDL_SYNC();
    // clip coefficient of clip_grad_norm_: min(1, max_norm / (||g|| + 1e-6)); grads may carry a
    // constant factor (grad_scale, e.g. 1/world_size) that is folded in here.
    float coef = grad_scale;
    if (norm_sq != nullptr) {
        const float total = sqrtf(norm_sq[0]) * grad_scale;
        const float c = max_norm / (total + 1e-6f);
        coef *= c < 1.0f ? c : 1.0f;
    }
    const float step_size = lr / bc1;
    const float decay = 1.0f - lr * wd;
    const long long n4 = n >> 2;
    float4* p4 = reinterpret_cast<float4*>(p);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    auto upd = [&](float& pp, float gg, float& mm, float& vv) {
        gg *= coef;
        pp *= decay;
        mm = fmaf(beta1, mm, (1.0f - beta1) * gg);
        vv = fmaf(beta2, vv, (1.0f - beta2) * gg * gg);
        const float denom = sqrtf(vv) / bc2_sqrt + eps;
        pp -= step_size * (mm / denom);
    };
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n4; i += (long long)gridDim.x * NT) {
        float4 pp = p4[i], mm = m4[i], vv = v4[i];
        const float4 gg = __ldg(g4 + i);
        upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y);
        upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
        p4[i] = pp; m4[i] = mm; v4[i] = vv;
        if (w16 != nullptr) {          // bf16 shadow of the updated weights: what the GEMMs read next step
            const __nv_bfloat162 lo = __floats2bfloat162_rn(pp.x, pp.y), hi = __floats2bfloat162_rn(pp.z, pp.w);
            uint2 o;
            o.x = *reinterpret_cast<const unsigned*>(&lo);
            o.y = *reinterpret_cast<const unsigned*>(&hi);
            reinterpret_cast<uint2*>(w16)[i] = o;
        }
    }
    if (blockIdx.x == 0)
        for (long long i = (n4 << 2) + threadIdx.x; i < n; i += NT) {
            upd(p[i], g[i], m[i], v[i]);
            if (w16 != nullptr) w16[i] = __float2bfloat16(p[i]);
        }
}

}  // namespace

extern "C" size_t eegx_sumsq_workspace_bytes(void) { return MAX_PARTIALS * sizeof(double); }

extern "C" int eegx_sumsq_f32(const float* g, int64_t n, float* out, int accumulate, void* workspace,
                              size_t workspace_bytes, void* stream) {
    if (int rc = eegx::require_sm100()) return rc;
    EEGX_REQUIRE(out && workspace, EEGX_ERR_ARG, "out/workspace must not be NULL");
    EEGX_REQUIRE(workspace_bytes >= MAX_PARTIALS * sizeof(double), EEGX_ERR_WORKSPACE,
                 "workspace too small: need %zu bytes", MAX_PARTIALS * sizeof(double));
    EEGX_REQUIRE(n >= 0 && (n == 0 || g), EEGX_ERR_ARG, "bad g/n");
    EEGX_REQUIRE(eegx::aligned16(g), EEGX_ERR_ALIGN, "g must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    long long blocks = (n / 4 + NT - 1) / NT;
    if (blocks < 1) blocks = 1;
    if (blocks > MAX_PARTIALS) blocks = MAX_PARTIALS;
    eegx::launch(sumsq_partial_kernel, (int)blocks, NT, 0, st, g, n, static_cast<double*>(workspace));
    eegx::launch(sumsq_final_kernel, 1, 32, 0, st, static_cast<const double*>(workspace), (int)blocks, out, accumulate);
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}

extern "C" int eegx_adamw_clip_f32(float* p, const float* g, float* m, float* v, int64_t n, float lr,
                                   float beta1, float beta2, float eps, float weight_decay, int64_t step,
                                   const float* grad_norm_sq, float max_norm, float grad_scale, void* w16,
                                   void* stream) {
    if (int rc = eegx::require_sm100()) return rc;
    EEGX_REQUIRE(n >= 0 && step >= 1, EEGX_ERR_ARG, "bad n/step");
    if (n == 0) return EEGX_OK;
    EEGX_REQUIRE(p && g && m && v, EEGX_ERR_ARG, "p/g/m/v must not be NULL");
    EEGX_REQUIRE(eegx::aligned16(p) && eegx::aligned16(g) && eegx::aligned16(m) && eegx::aligned16(v),
                 EEGX_ERR_ALIGN, "p/g/m/v must be 16-byte aligned");
    const float bc1 = 1.0f - powf(beta1, (float)step);
    const float bc2_sqrt = sqrtf(1.0f - powf(beta2, (float)step));
    long long blocks = (n / 4 + NT - 1) / NT;
    if (blocks < 1) blocks = 1;
    if (blocks > eegx::kNumSMsB200 * 8) blocks = eegx::kNumSMsB200 * 8;
    eegx::launch(adamw_kernel, (int)blocks, NT, 0, static_cast<cudaStream_t>(stream), 
        p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2_sqrt, grad_norm_sq, max_norm, grad_scale,
        static_cast<__nv_bfloat16*>(w16));
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}
