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
// One CTA per row.  Each WARP keeps one sorted top-k list spread over its lanes (lane j holds the j-th best so far)
// and a broadcast copy of its k-th entry as the admission threshold: an element costs a load, the online
// (max, sum-exp) update, one comparison and a ballot; only the ~k ln(n) elements that beat the threshold take the
// insertion path (a warp-wide shift by shuffles).  Per-thread lists would diverge on almost every element (each of
// the 32 lanes admits ~k/n of its own elements).  The 8 warp lists are merged by warp 0 with the same insertion.
#include <math.h>

#include "eegx_common.h"

namespace {

constexpr int TK_THREADS = 256;
constexpr int TK_WARPS = TK_THREADS / 32;
constexpr int NOIDX = 0x7fffffff;

__device__ __forceinline__ bool better(float v, int i, float w, int j) { return v > w || (v == w && i < j); }

// Insert (xv, xi) into the warp's sorted list (lane j < k holds entry j); all lanes call it with the same candidate.
__device__ __forceinline__ void warp_insert(float& lv, int& li, float xv, int xi, int k, int lane) {
    const unsigned ahead = __ballot_sync(0xffffffffu, lane < k && better(lv, li, xv, xi));
    const int pos = __popc(ahead);                                  // entries that stay in front of the candidate
    const float up_v = __shfl_up_sync(0xffffffffu, lv, 1);
    const int up_i = __shfl_up_sync(0xffffffffu, li, 1);
    if (lane == pos) { lv = xv; li = xi; }
    else if (lane > pos && lane < k) { lv = up_v; li = up_i; }
}

__global__ void __launch_bounds__(TK_THREADS)
logsoftmax_topk_kernel(const float* __restrict__ logits, long long ld, int V, int k, int banned, int vec_ok,
                       float* __restrict__ out_val, long long* __restrict__ out_idx) {
    EEGX_PDL_SYNC();
    __shared__ float s_m[TK_WARPS], s_s[TK_WARPS];
    __shared__ float s_v[TK_WARPS * 32];
    __shared__ int s_i[TK_WARPS * 32];
    // block-wide admission floor: the k-th best value of ANY warp is a lower bound of the row's k-th best, so the
    // warps publish theirs and skip everything below the highest one seen (a racy, possibly stale value is still
    // a valid bound; strict '<' keeps ties for the exact test)
    __shared__ volatile float s_floor;
    if (threadIdx.x == 0) s_floor = -INFINITY;
    __syncthreads();
    const float* row = logits + (long long)blockIdx.x * ld;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    float lv = -INFINITY;                                            // this lane's entry of the warp list
    int li = NOIDX;
    float thr_v = -INFINITY;                                         // entry k-1 (the worst admitted), all lanes
    int thr_i = NOIDX;
    float m = -INFINITY, s = 0.0f;
    // ---- main loop: four elements per thread and iteration (128-bit loads); one rescale of the running sum per
    //      group, and the admission test on the group's maximum first
    const int V4 = vec_ok ? (V >> 2) : 0;
    const int iters4 = (V4 + TK_THREADS - 1) / TK_THREADS;
    const float4* row4 = reinterpret_cast<const float4*>(row);
    for (int it = 0; it < iters4; ++it) {
        const int q = it * TK_THREADS + tid;
        const bool valid = q < V4;
        float4 v = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        if (valid) v = row4[q];
        const float mx = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
        if (valid) {
            if (mx > m) { s *= __expf(m - mx); m = mx; }
            s += (__expf(v.x - m) + __expf(v.y - m)) + (__expf(v.z - m) + __expf(v.w - m));
        }
        const float floor_v = fmaxf(thr_v, s_floor);
        unsigned cand = __ballot_sync(0xffffffffu, valid && mx >= floor_v);
        while (cand) {
            const int src = __ffs(cand) - 1;
            cand &= cand - 1;
            const int base = 4 * (it * TK_THREADS + (warp << 5) + src);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float xe = e == 0 ? v.x : (e == 1 ? v.y : (e == 2 ? v.z : v.w));
                const float xv = __shfl_sync(0xffffffffu, xe, src);
                const int xi = base + e;
                if (xi == banned || !better(xv, xi, thr_v, thr_i)) continue;
                warp_insert(lv, li, xv, xi, k, lane);
                thr_v = __shfl_sync(0xffffffffu, lv, k - 1);
                thr_i = __shfl_sync(0xffffffffu, li, k - 1);
                if (lane == 0 && thr_v > s_floor) s_floor = thr_v;
            }
        }
    }
    // ---- the remaining elements (all of them when the row is not 16-byte aligned), one per thread and iteration
    const int c0 = 4 * V4;
    const int iters = (V - c0 + TK_THREADS - 1) / TK_THREADS;
    for (int it = 0; it < iters; ++it) {
        const int c = c0 + it * TK_THREADS + tid;
        const bool valid = c < V;
        const float x = valid ? row[c] : -INFINITY;
        if (valid) {
            if (x > m) { s = s * __expf(m - x) + 1.0f; m = x; }
            else s += __expf(x - m);
        }
        unsigned cand = __ballot_sync(0xffffffffu, valid && c != banned && better(x, c, thr_v, thr_i));
        while (cand) {
            const int src = __ffs(cand) - 1;
            cand &= cand - 1;
            const float xv = __shfl_sync(0xffffffffu, x, src);
            const int xi = c0 + it * TK_THREADS + (warp << 5) + src;
            if (!better(xv, xi, thr_v, thr_i)) continue;             // the threshold rose since the ballot
            warp_insert(lv, li, xv, xi, k, lane);
            thr_v = __shfl_sync(0xffffffffu, lv, k - 1);
            thr_i = __shfl_sync(0xffffffffu, li, k - 1);
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
    s_v[tid] = lane < k ? lv : -INFINITY;
    s_i[tid] = lane < k ? li : NOIDX;
    __syncthreads();
    if (warp != 0) return;
    // warp 0 merges: its own list is the start, the other warps' entries are offered in order
    float M = s_m[0], S = s_s[0];
    for (int w = 1; w < TK_WARPS; ++w) {
        const float mm = fmaxf(M, s_m[w]);
        S = (M == -INFINITY ? 0.0f : S * __expf(M - mm)) + (s_m[w] == -INFINITY ? 0.0f : s_s[w] * __expf(s_m[w] - mm));
        M = mm;
    }
    const float lse = M + logf(S);
    for (int w = 1; w < TK_WARPS; ++w)
        for (int j = 0; j < k; ++j) {
            const float xv = s_v[w * 32 + j];
            const int xi = s_i[w * 32 + j];
            if (!better(xv, xi, thr_v, thr_i)) break;                // that list is sorted: nothing further can enter
            warp_insert(lv, li, xv, xi, k, lane);
            thr_v = __shfl_sync(0xffffffffu, lv, k - 1);
            thr_i = __shfl_sync(0xffffffffu, li, k - 1);
        }
    if (lane < k) {
        out_val[(long long)blockIdx.x * k + lane] = lv - lse;
        out_idx[(long long)blockIdx.x * k + lane] = li == NOIDX ? 0 : li;
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
    const int vec_ok = eegx::aligned16(logits) && (ld % 4) == 0;
    const int ban = banned >= 0 && banned < V ? (int)banned : -1;
    EEGX_CUDA_CHECK(eegx::launch(logsoftmax_topk_kernel, (unsigned)rows, TK_THREADS, 0, st, logits, (long long)ld,
                                 (int)V, (int)k, ban, vec_ok, out_val, reinterpret_cast<long long*>(out_idx)));
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}
