// Cross-entropy over the LM-head logits, forward and backward, one CTA per token row.
//
// Replaces, at the loss end of the train step (main_model/src/models/bart_decoder.py:41-48 ->
// transformers BartForConditionalGeneration.forward: lm_head + final_logits_bias +
// CrossEntropyLoss(ignore_index=-100)), the fp32 softmax / log-softmax / nll_loss chain over the
// (B*L) x 51271 logits.  The logits come out of the tcgen05 GEMM in bf16 (bias fused); forward
// is ONE read of them (online max / sum-exp per thread, fixed-order block reduction), backward is
// one read + one write producing d(logits) = (softmax - onehot) * coef in bf16 for the dgrad /
// wgrad GEMMs.  The fp32 (B*L, V) tensors of the library path (840 MB at B = 256) never exist.
#include "eegx_common.h"
#include "fused_common.cuh"

namespace {

using namespace eegx;
constexpr int CE_THREADS = 256;

__device__ __forceinline__ void online_add(float& m, float& s, float x) {
    if (x > m) {
        s = s * __expf(m - x) + 1.0f;
        m = x;
    } else {
        s += __expf(x - m);
    }
}

__device__ __forceinline__ void online_merge(float& m, float& s, float m2, float s2) {
    const float mm = fmaxf(m, m2);
    s = (m == -INFINITY ? 0.0f : s * __expf(m - mm)) + (m2 == -INFINITY ? 0.0f : s2 * __expf(m2 - mm));
    m = mm;
}

__global__ void __launch_bounds__(CE_THREADS)
ce_fwd_kernel(const __nv_bfloat16* __restrict__ logits, long long ld, const long long* __restrict__ labels,
              long long V, long long ignore_index, float* __restrict__ loss_rows, float* __restrict__ lse_out) {
    EEGX_PDL_SYNC();
    __shared__ float sm[CE_THREADS / 32], ss[CE_THREADS / 32];
    const long long r = blockIdx.x;
    const __nv_bfloat16* row = logits + r * ld;
    float m = -INFINITY, s = 0.0f;
    const long long v8 = V >> 3;
    for (long long g = threadIdx.x; g < v8; g += CE_THREADS) {
        float x[8];
        load8(row + g * 8, x);
#pragma unroll
        for (int e = 0; e < 8; ++e) online_add(m, s, x[e]);
    }
    for (long long c = (v8 << 3) + threadIdx.x; c < V; c += CE_THREADS) online_add(m, s, __bfloat162float(row[c]));
    // fixed-order reduction: lanes by xor tree, then warps in order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
        online_merge(m, s, m2, s2);
    }
    if ((threadIdx.x & 31) == 0) { sm[threadIdx.x >> 5] = m; ss[threadIdx.x >> 5] = s; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float M = sm[0], S = ss[0];
        for (int w = 1; w < CE_THREADS / 32; ++w) online_merge(M, S, sm[w], ss[w]);
        const float lse = M + logf(S);
        lse_out[r] = lse;
        const long long y = labels[r];
        loss_rows[r] = (y == ignore_index || y < 0 || y >= V) ? 0.0f : lse - __bfloat162float(row[y]);
    }
}

__global__ void __launch_bounds__(CE_THREADS)
ce_bwd_kernel(const __nv_bfloat16* __restrict__ logits, long long ld, const long long* __restrict__ labels,
              const float* __restrict__ lse, const float* __restrict__ coef, __nv_bfloat16* __restrict__ dlogits,
              long long V, long long ignore_index) {
    EEGX_PDL_SYNC();
    const long long r = blockIdx.x;
    const __nv_bfloat16* row = logits + r * ld;
    __nv_bfloat16* drow = dlogits + r * ld;
    const long long y = labels[r];
    const bool valid = !(y == ignore_index || y < 0 || y >= V);
    const float c = valid ? coef[0] : 0.0f, l = lse[r];
    const long long ld8 = ld >> 3;
    for (long long g = threadIdx.x; g < ld8; g += CE_THREADS) {
        float x[8], o[8];
        load8(row + g * 8, x);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const long long col = g * 8 + e;
            o[e] = (valid && col < V) ? (__expf(x[e] - l) - (col == y ? 1.0f : 0.0f)) * c : 0.0f;
        }
        store8(drow + g * 8, o);
    }
}

}  // namespace

extern "C" {

int eegx_ce_fwd_bf16(const void* logits, int64_t ld, const int64_t* labels, int64_t rows, int64_t V,
                     int64_t ignore_index, float* loss_rows, float* lse, void* stream) {
    if (int rc = eegx::require_sm100()) return rc;
    EEGX_REQUIRE(rows >= 0 && V >= 1 && ld >= V && (ld % 8) == 0, EEGX_ERR_SHAPE,
                 "cross-entropy: need ld >= V and ld a multiple of 8");
    if (rows == 0) return EEGX_OK;
    EEGX_REQUIRE(logits && labels && loss_rows && lse, EEGX_ERR_ARG, "cross-entropy: NULL pointer");
    EEGX_REQUIRE(eegx::aligned16(logits), EEGX_ERR_ALIGN, "cross-entropy: logits must be 16-byte aligned");
    eegx::launch(ce_fwd_kernel, (unsigned)rows, CE_THREADS, 0, static_cast<cudaStream_t>(stream), 
        static_cast<const __nv_bfloat16*>(logits), ld, reinterpret_cast<const long long*>(labels), V, ignore_index,
        loss_rows, lse);
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}

int eegx_ce_bwd_bf16(const void* logits, int64_t ld, const int64_t* labels, const float* lse, const float* coef,
                     void* dlogits, int64_t rows, int64_t V, int64_t ignore_index, void* stream) {
    if (int rc = eegx::require_sm100()) return rc;
    EEGX_REQUIRE(rows >= 0 && V >= 1 && ld >= V && (ld % 8) == 0, EEGX_ERR_SHAPE,
                 "cross-entropy: need ld >= V and ld a multiple of 8");
    if (rows == 0) return EEGX_OK;
    EEGX_REQUIRE(logits && labels && lse && coef && dlogits, EEGX_ERR_ARG, "cross-entropy bwd: NULL pointer");
    EEGX_REQUIRE(eegx::aligned16(logits) && eegx::aligned16(dlogits), EEGX_ERR_ALIGN,
                 "cross-entropy bwd: logits / dlogits must be 16-byte aligned");
    eegx::launch(ce_bwd_kernel, (unsigned)rows, CE_THREADS, 0, static_cast<cudaStream_t>(stream), 
        static_cast<const __nv_bfloat16*>(logits), ld, reinterpret_cast<const long long*>(labels), lse, coef,
        static_cast<__nv_bfloat16*>(dlogits), V, ignore_index);
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}

}  // extern "C"
