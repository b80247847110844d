// log-softmax + top-k of every logits row in one pass, for the beam-search step of generation.py.
//
// Replaces, per decode step of transformers' beam search (generation/utils.py GenerationMixin._beam_search: fp32
// log_softmax over the (batch*beams, V) logits, MinLength processor, add the running beam scores, torch.topk over
// the batch's beams*V accumulated scores), the three V-wide passes by ONE read of the logits: the best 2*beams
// continuations over all beams of a batch item are always among the best 2*beams of each single beam, so a per-row
// top-k (k = 2*beams) plus the row's log-sum-exp is all the bookkeeping needs; the merge over the beams is then a
// top-k over beams*k numbers.
//
//   out_val[r, j] = logits[r, idx_j] - logsumexp(logits[r, :])      (j-th largest, descending; ties: lower index first)
//   out_idx[r, j] = idx_j
//   banned >= 0: that token is excluded from the selection but NOT from the log-sum-exp (MinLengthLogitsProcessor
//   runs after log_softmax in the library).
//
// One CTA per row; every thread keeps a sorted top-k of its strided share in registers and an online (max, sum-exp)
// pair; k rounds of a block-wide arg-max over the threads' list heads produce the result in order.
#include <math.h>

#include "eegx_common.h"

namespace {

constexpr int TK_THREADS = 256;

struct Cand {
    float v;
    int i;
};
__device__ __forceinline__ bool better(float v, int i, float w, int j) { return v > w || (v == w && i < j); }

template <int KMAX>
__global__ void __launch_bounds__(TK_THREADS)
logsoftmax_topk_kernel(const float* __restrict__ logits, long long ld, long long V, int k, long long banned,
                       float* __restrict__ out_val, long long* __restrict__ out_idx) {
    EEGX_PDL_SYNC();
    __shared__ float s_m[TK_THREADS / 32], s_s[TK_THREADS / 32];
    __shared__ float s_v[TK_THREADS / 32];
    __shared__ int s_i[TK_THREADS / 32], s_t[TK_THREADS / 32];
    __shared__ int s_win;
    __shared__ float s_lse;
    const float* row = logits + (long long)blockIdx.x * ld;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    float tv[KMAX];
    int ti[KMAX];
#pragma unroll
    for (int j = 0; j < KMAX; ++j) { tv[j] = -INFINITY; ti[j] = 0x7fffffff; }
    float m = -INFINITY, s = 0.0f;
    for (long long c = tid; c < V; c += TK_THREADS) {
        const float x = row[c];
        if (x > m) { s = s * __expf(m - x) + 1.0f; m = x; }
        else s += __expf(x - m);
        if (c != banned && better(x, (int)c, tv[KMAX - 1], ti[KMAX - 1])) {
            tv[KMAX - 1] = x; ti[KMAX - 1] = (int)c;
#pragma unroll
            for (int j = KMAX - 1; j > 0; --j)
                if (better(tv[j], ti[j], tv[j - 1], ti[j - 1])) {
                    const float a = tv[j]; tv[j] = tv[j - 1]; tv[j - 1] = a;
                    const int b = ti[j]; ti[j] = ti[j - 1]; ti[j - 1] = b;
                }
        }
    }
    // log-sum-exp of the row: lanes by xor tree, warps in order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
        const float mm = fmaxf(m, m2);
        s = (m == -INFINITY ? 0.0f : s * __expf(m - mm)) + (m2 == -INFINITY ? 0.0f : s2 * __expf(m2 - mm));
        m = mm;
    }
    if (lane == 0) { s_m[warp] = m; s_s[warp] = s; }
    __syncthreads();
    if (tid == 0) {
        float M = s_m[0], S = s_s[0];
        for (int w = 1; w < TK_THREADS / 32; ++w) {
            const float mm = fmaxf(M, s_m[w]);
            S = (M == -INFINITY ? 0.0f : S * __expf(M - mm)) + (s_m[w] == -INFINITY ? 0.0f : s_s[w] * __expf(s_m[w] - mm));
            M = mm;
        }
        s_lse = M + logf(S);
    }
    // k rounds: block-wide arg-max over the heads of the per-thread lists; the winner pops its head
    for (int r = 0; r < k; ++r) {
        float v = tv[0];
        int i = ti[0], t = tid;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float v2 = __shfl_xor_sync(0xffffffffu, v, o);
            const int i2 = __shfl_xor_sync(0xffffffffu, i, o), t2 = __shfl_xor_sync(0xffffffffu, t, o);
            if (better(v2, i2, v, i)) { v = v2; i = i2; t = t2; }
        }
        if (lane == 0) { s_v[warp] = v; s_i[warp] = i; s_t[warp] = t; }
        __syncthreads();
        if (tid == 0) {
            float bv = s_v[0];
            int bi = s_i[0], bt = s_t[0];
            for (int w = 1; w < TK_THREADS / 32; ++w)
                if (better(s_v[w], s_i[w], bv, bi)) { bv = s_v[w]; bi = s_i[w]; bt = s_t[w]; }
            s_win = bt;
            out_val[(long long)blockIdx.x * k + r] = bv - s_lse;
            out_idx[(long long)blockIdx.x * k + r] = bi == 0x7fffffff ? 0 : bi;
        }
        __syncthreads();
        if (tid == s_win) {
#pragma unroll
            for (int j = 0; j < KMAX - 1; ++j) { tv[j] = tv[j + 1]; ti[j] = ti[j + 1]; }
            tv[KMAX - 1] = -INFINITY; ti[KMAX - 1] = 0x7fffffff;
        }
    }
}

}  // namespace

extern "C" int eegx_logsoftmax_topk_f32(const float* logits, int64_t ld, int64_t rows, int64_t V, int32_t k,
                                        int64_t banned, float* out_val, int64_t* out_idx, void* stream) {
    if (int rc = eegx::require_sm100()) return rc;
    EEGX_REQUIRE(rows >= 0 && V >= 1 && ld >= V && V < (1LL << 31) && rows < (1LL << 31), EEGX_ERR_SHAPE,
                 "logsoftmax_topk: need 1 <= V <= ld");
    EEGX_REQUIRE(k >= 1 && k <= 16 && k <= V, EEGX_ERR_ARG, "logsoftmax_topk: k must be in [1, min(16, V)]");
    if (rows == 0) return EEGX_OK;
    EEGX_REQUIRE(logits && out_val && out_idx, EEGX_ERR_ARG, "logsoftmax_topk: NULL pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (k <= 8)
        EEGX_CUDA_CHECK(eegx::launch(logsoftmax_topk_kernel<8>, (unsigned)rows, TK_THREADS, 0, st, logits, (long long)ld,
                                     (long long)V, (int)k, (long long)banned, out_val,
                                     reinterpret_cast<long long*>(out_idx)));
    else
        EEGX_CUDA_CHECK(eegx::launch(logsoftmax_topk_kernel<16>, (unsigned)rows, TK_THREADS, 0, st, logits, (long long)ld,
                                     (long long)V, (int)k, (long long)banned, out_val,
                                     reinterpret_cast<long long*>(out_idx)));
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}
