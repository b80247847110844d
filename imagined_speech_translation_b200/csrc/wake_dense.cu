// wake_model's dense head -- Linear(in, H, act) -> Linear(H, n_cls, softmax) -> categorical cross-entropy, trained by
// per-sample SGD -- as ONE persistent cooperative kernel (BASELINE config 5).
//
// What it replaces (reference, C++, fp64, one sample at a time):
//   wake_model/layers/linear.cpp:5-44    Linear::forward  (out = sum_j in_j * w_ij, + bias, activation / softmax)
//   wake_model/layers/linear.cpp:47-72   Linear::backward (dout *= act'(OUTPUT); dinput_j += w_ij * dout_i with the
//                                        weight BEFORE its update; w_ij -= lr * (in_j * dout_i); b_i -= lr * dout_i)
//   wake_model/layers/activations.h:12-41, 64-95 (sigmoid / tanh / relu, derivative evaluated on the layer OUTPUT,
//                                        softmax with max subtraction), wake_model/layers/losses.h:8-22
//   wake_model/train.cpp:98-117          the loop: forward, loss, delta = p - onehot, backward through both layers
//
// The update is sequential in the samples (sample s+1 sees the weights sample s wrote), so the parallelism is inside
// one sample: CTA c owns a contiguous slice of the H hidden rows -- the rows of W1, the matching COLUMNS of W2 -- and
// keeps them to itself for the whole launch.  Per sample there is ONE grid-wide barrier (the n_cls partial logits of
// every CTA must meet before the softmax); everything else is CTA-local:
//
//   h_own      = act(W1[own] . x_s + b1[own])                         (carried over from the previous iteration)
//   partial_c  = W2[:, own] . h_own                 -> global, grid.sync(), every CTA sums the G partials in order
//   p = softmax(z + b2), d2 = p - onehot, loss
//   dh_own     = W2[:, own]^T d2 (old W2);  W2[:, own] -= lr * (h_own * d2);  b2 -= lr * d2 (CTA 0)
//   d1_own     = dh_own * act'(h_own)
//   one pass over W1[own]:  dx += w * d1 (old w);  w -= lr * (x_s * d1);  acc_next += w * x_{s+1}   <- the forward of
//   the NEXT sample rides on the same read, so W1 is read once and written once per sample instead of 2 + 1.
//
// W1 (H x in fp64, 32 MB at H = 1024, in = 4096) stays L2-resident between samples; algorithmic traffic per sample is
// 16 * H * in bytes.  The multiplications / subtractions of the update use explicit round-to-nearest intrinsics in
// the reference's operation order (no FMA contraction), so one update step is bit-identical to the C++; the dot
// products are reduced in a fixed (thread, warp) order that differs from the reference's j = 0..in-1 chain, which is
// where the stated 1e-10 tolerance comes from.
#include <cooperative_groups.h>
#include <math.h>

#include "eegx_common.h"

namespace cg = cooperative_groups;

namespace {

using namespace eegx;

constexpr int WK_THREADS = 512;
constexpr int WK_WARPS = WK_THREADS / 32;
constexpr int WK_RC = 8;              // hidden rows handled together in one pass over j

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_SIGMOID = 2, ACT_TANH = 3 };

struct WakeParams {
    double* w1; double* b1; double* w2; double* b2;
    const double* x; const int* label;
    long long n; int in, hidden, ncls;
    double lr; int act; int train;
    double* loss; double* probs; double* dx;
    double* partial;                   // [2][G][ncls]
    double* dx_part;                   // [G][in] (only when dx != nullptr)
    int rows_per_cta;
};

__device__ __forceinline__ double act_fwd(double v, int act) {
    switch (act) {
        case ACT_RELU: return fmax(0.0, v);
        case ACT_SIGMOID: return 1.0 / (1.0 + exp(-v));
        case ACT_TANH: { const double a = exp(v), b = exp(-v); return (a - b) / (a + b); }
        default: return v;
    }
}
// the reference differentiates at the layer OUTPUT (linear.cpp:53-56 passes output_neurons[i].output), i.e. for
// sigmoid / tanh the activation is applied a second time -- kept as is
__device__ __forceinline__ double act_bwd(double out, int act) {
    switch (act) {
        case ACT_RELU: return out > 0.0 ? 1.0 : 0.0;
        case ACT_SIGMOID: { const double s = 1.0 / (1.0 + exp(-out)); return s * (1.0 - s); }
        case ACT_TANH: { const double a = exp(out), b = exp(-out); const double t = (a - b) / (a + b); return 1.0 - t * t; }
        default: return 1.0;
    }
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide reduction of WK_RC running sums held by every thread; result[r] valid for all threads after return
__device__ __forceinline__ void block_sum_rc(double (&acc)[WK_RC], double (*red)[WK_RC], double* result) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int r = 0; r < WK_RC; ++r) acc[r] = warp_sum_d(acc[r]);
    if (lane == 0) {
#pragma unroll
        for (int r = 0; r < WK_RC; ++r) red[warp][r] = acc[r];
    }
    __syncthreads();
    if (threadIdx.x < WK_RC) {
        double s = 0.0;
        for (int w = 0; w < WK_WARPS; ++w) s += red[w][threadIdx.x];
        result[threadIdx.x] = s;
    }
    __syncthreads();
}

__device__ __forceinline__ double block_max(double v, double* red) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double m = red[0];
    for (int w = 1; w < WK_WARPS; ++w) m = fmax(m, red[w]);
    __syncthreads();
    return m;
}
__device__ __forceinline__ double block_sum(double v, double* red) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    v = warp_sum_d(v);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < WK_WARPS; ++w) s += red[w];
    __syncthreads();
    return s;
}

__global__ void __launch_bounds__(WK_THREADS, 1)
wake_dense_kernel(const WakeParams p) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ double smem_d[];
    const int in = p.in, ncls = p.ncls, R = p.rows_per_cta;
    double* xs0 = smem_d;                       // [in]   current sample
    double* xs1 = xs0 + in;                     // [in]   next sample
    double* hpre = xs1 + in;                    // [R]    pre-activation (+bias) of the current sample, own rows
    double* hown = hpre + R;                    // [R]
    double* d1 = hown + R;                      // [R]
    double* zs = d1 + R;                        // [ncls] logits -> probabilities -> d2
    double* b2s = zs + ncls;                    // [ncls] private copy of b2 (every CTA applies the same updates)
    double* w2s = b2s + ncls;                   // [ncls][R] the own columns of W2, resident for the whole launch
    double* gsum = w2s + (size_t)ncls * R;      // [max(WK_THREADS, ncls)] partial-logit group sums
    double* red1 = gsum + max(WK_THREADS, ncls);   // [WK_WARPS]
    double* rcres = red1 + WK_WARPS;            // [WK_RC]
    double (*red)[WK_RC] = reinterpret_cast<double (*)[WK_RC]>(rcres + WK_RC);   // [WK_WARPS][WK_RC]

    const int G = gridDim.x, c = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r0 = min(c * R, p.hidden), r1 = min(r0 + R, p.hidden), nr = r1 - r0;

    for (int k = tid; k < ncls; k += WK_THREADS) b2s[k] = p.b2[k];
    for (int i = tid; i < ncls * nr; i += WK_THREADS) {
        const int k = i / nr, r = i - k * nr;
        w2s[k * R + r] = p.w2[(long long)k * p.hidden + r0 + r];
    }
    // ---- sample 0: stage x_0 and compute the own rows' pre-activations with a plain pass
    if (p.n > 0)
        for (int j = tid; j < in; j += WK_THREADS) xs0[j] = p.x[j];
    __syncthreads();
    for (int rb = 0; rb < nr; rb += WK_RC) {
        double acc[WK_RC];
#pragma unroll
        for (int r = 0; r < WK_RC; ++r) acc[r] = 0.0;
        const int cnt = min(WK_RC, nr - rb);
        for (int j = tid; j < in; j += WK_THREADS) {
            const double xj = xs0[j];
#pragma unroll
            for (int r = 0; r < WK_RC; ++r)
                if (r < cnt) acc[r] += p.w1[(long long)(r0 + rb + r) * in + j] * xj;
        }
        block_sum_rc(acc, red, rcres);
        if (tid < cnt) hpre[rb + tid] = rcres[tid] + p.b1[r0 + rb + tid];
        __syncthreads();
    }

    double* xc = xs0;
    double* xn = xs1;
    for (long long s = 0; s < p.n; ++s) {
        const bool has_next = s + 1 < p.n;
        if (has_next)
            for (int j = tid; j < in; j += WK_THREADS) xn[j] = p.x[(s + 1) * in + j];
        for (int r = tid; r < nr; r += WK_THREADS) hown[r] = act_fwd(hpre[r], p.act);
        __syncthreads();

        // ---- partial logits of the own columns of W2
        double* part = p.partial + ((s & 1) * (long long)G + c) * ncls;
        for (int k = tid; k < ncls; k += WK_THREADS) {
            double z = 0.0;
            const double* w2k = w2s + k * R;
            for (int r = 0; r < nr; ++r) z += hown[r] * w2k[r];
            part[k] = z;
        }
        grid.sync();

        // ---- every CTA: full logits, softmax, d2 (identical in all CTAs: same data, same order)
        const double* pall = p.partial + (s & 1) * (long long)G * ncls;
        // the G partials of logit k are summed by NG threads (contiguous CTA ranges, loads batched by the unroll),
        // then the NG group sums in order: a fixed order, the same in every CTA
        const int NG = max(1, min(WK_THREADS / ncls, 16)), gper = (G + NG - 1) / NG;
        for (int item = tid; item < NG * ncls; item += WK_THREADS) {
            const int grp = item / ncls, k = item - grp * ncls;
            const int g0 = grp * gper, g1 = min(G, g0 + gper);
            double a = 0.0;
#pragma unroll 8
            for (int g = g0; g < g1; ++g) a += pall[(long long)g * ncls + k];
            gsum[item] = a;
        }
        __syncthreads();
        double zmax = -INFINITY;
        for (int k = tid; k < ncls; k += WK_THREADS) {
            double z = gsum[k];
            for (int grp = 1; grp < NG; ++grp) z += gsum[grp * ncls + k];
            z += b2s[k];
            zs[k] = z;
            zmax = fmax(zmax, z);
        }
        zmax = block_max(zmax, red1);
        double esum = 0.0;
        for (int k = tid; k < ncls; k += WK_THREADS) {
            const double e = exp(zs[k] - zmax);
            zs[k] = e;
            esum += e;
        }
        esum = block_sum(esum, red1);
        const int y = p.label[s];
        for (int k = tid; k < ncls; k += WK_THREADS) {
            const double pr = zs[k] / esum;
            if (c == 0) {
                if (p.probs) p.probs[s * ncls + k] = pr;
                if (k == y && p.loss) p.loss[s] = -log(pr + 1e-15);
                if (k == 0 && p.loss && (y < 0 || y >= ncls)) p.loss[s] = nan("");     // the reference indexes out of bounds here
            }
            zs[k] = pr - (k == y ? 1.0 : 0.0);            // d2
        }
        __syncthreads();
        if (!p.train) {
            // inference: next sample's pre-activations with a plain pass, no updates
            if (has_next) {
                for (int rb = 0; rb < nr; rb += WK_RC) {
                    double acc[WK_RC];
#pragma unroll
                    for (int r = 0; r < WK_RC; ++r) acc[r] = 0.0;
                    const int cnt = min(WK_RC, nr - rb);
                    for (int j = tid; j < in; j += WK_THREADS) {
                        const double xj = xn[j];
#pragma unroll
                        for (int r = 0; r < WK_RC; ++r)
                            if (r < cnt) acc[r] += p.w1[(long long)(r0 + rb + r) * in + j] * xj;
                    }
                    block_sum_rc(acc, red, rcres);
                    if (tid < cnt) hpre[rb + tid] = rcres[tid] + p.b1[r0 + rb + tid];
                    __syncthreads();
                }
            }
            double* t = xc; xc = xn; xn = t;
            continue;
        }

        // ---- layer 2 backward on the own columns: dh with the OLD weights, then the update
        for (int r = warp; r < nr; r += WK_WARPS) {
            const double hr = hown[r];
            double dh = 0.0;
            for (int k = lane; k < ncls; k += 32) {
                double* wp = w2s + k * R + r;
                const double w = *wp, d = zs[k];
                dh += w * d;
                *wp = __dsub_rn(w, __dmul_rn(p.lr, __dmul_rn(hr, d)));
            }
            dh = warp_sum_d(dh);
            if (lane == 0) d1[r] = __dmul_rn(dh, act_bwd(hr, p.act));
        }
        for (int k = tid; k < ncls; k += WK_THREADS) b2s[k] = __dsub_rn(b2s[k], __dmul_rn(p.lr, zs[k]));
        __syncthreads();

        // ---- layer 1: one pass over the own rows of W1 = backward of sample s + forward of sample s+1
        const bool want_dx = p.dx != nullptr;
        double* dxp = want_dx ? p.dx_part + (long long)c * in : nullptr;
        for (int rb = 0; rb < nr; rb += WK_RC) {
            double acc[WK_RC], dr[WK_RC];
            const int cnt = min(WK_RC, nr - rb);
#pragma unroll
            for (int r = 0; r < WK_RC; ++r) {
                acc[r] = 0.0;
                dr[r] = r < cnt ? d1[rb + r] : 0.0;
            }
#pragma unroll 2
            for (int j = tid; j < in; j += WK_THREADS) {
                const double xj = xc[j], xnj = has_next ? xn[j] : 0.0;
                double dxa = 0.0;
#pragma unroll
                for (int r = 0; r < WK_RC; ++r) {
                    if (r < cnt) {
                        double* wp = p.w1 + (long long)(r0 + rb + r) * in + j;
                        double w = *wp;
                        dxa += w * dr[r];
                        w = __dsub_rn(w, __dmul_rn(p.lr, __dmul_rn(xj, dr[r])));
                        *wp = w;
                        acc[r] += w * xnj;
                    }
                }
                if (want_dx) dxp[j] = (rb == 0 ? 0.0 : dxp[j]) + dxa;
            }
            block_sum_rc(acc, red, rcres);
            if (tid < cnt) {
                const int i = r0 + rb + tid;
                const double b = __dsub_rn(p.b1[i], __dmul_rn(p.lr, d1[rb + tid]));
                p.b1[i] = b;
                hpre[rb + tid] = rcres[tid] + b;
            }
            __syncthreads();
        }
        if (want_dx) {
            if (nr == 0)
                for (int j = tid; j < in; j += WK_THREADS) dxp[j] = 0.0;
            grid.sync();
            // column slices of dx summed over the G partials in CTA order
            for (long long j = (long long)c * WK_THREADS + tid; j < in; j += (long long)G * WK_THREADS) {
                double a = 0.0;
                for (int g = 0; g < G; ++g) a += p.dx_part[(long long)g * in + j];
                p.dx[s * in + j] = a;
            }
            // the next iteration's grid.sync() (partial logits) orders these reads before dx_part is rewritten
        }
        double* t = xc; xc = xn; xn = t;
    }
    if (p.train) {
        __syncthreads();
        for (int i = tid; i < ncls * nr; i += WK_THREADS) {
            const int k = i / nr, r = i - k * nr;
            p.w2[(long long)k * p.hidden + r0 + r] = w2s[k * R + r];
        }
        if (c == 0)
            for (int k = tid; k < ncls; k += WK_THREADS) p.b2[k] = b2s[k];
    }
}

size_t smem_bytes(int in, int R, int ncls) {
    return sizeof(double) * ((size_t)2 * in + 3 * (size_t)R + 2 * (size_t)ncls + (size_t)ncls * R +
                             (size_t)(ncls > WK_THREADS ? ncls : WK_THREADS) + WK_WARPS + WK_RC + WK_WARPS * WK_RC);
}

int grid_for(int64_t hidden, int* rows_per_cta) {
    int dev = 0, sms = kNumSMsB200;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int G = (int)(hidden < sms ? hidden : sms);
    if (G < 1) G = 1;
    *rows_per_cta = (int)((hidden + G - 1) / G);
    return G;
}

}  // namespace

extern "C" {

size_t eegx_wake_dense_workspace_bytes(int64_t in, int64_t hidden, int64_t n_cls, int want_dx) {
    const size_t G = kNumSMsB200 * 2;         // upper bound on the grid (any sm_100 part has <= 2 * 148 SMs here)
    return sizeof(double) * (2 * G * (size_t)n_cls + (want_dx ? G * (size_t)in : 0)) + 256;
}

int eegx_wake_dense_f64(double* w1, double* b1, double* w2, double* b2, const double* x, const int32_t* label,
                        int64_t n, int64_t in, int64_t hidden, int64_t n_cls, double lr, int activation, int train,
                        double* loss, double* probs, double* dx, void* workspace, size_t workspace_bytes,
                        void* stream) {
    if (int rc = eegx::require_sm100()) return rc;
    EEGX_REQUIRE(n >= 0 && in >= 1 && hidden >= 1 && n_cls >= 1 && in < (1LL << 30) && hidden < (1LL << 30) &&
                 n_cls < (1LL << 30), EEGX_ERR_SHAPE, "wake_dense: need in, hidden, n_cls >= 1");
    EEGX_REQUIRE(activation >= ACT_NONE && activation <= ACT_TANH, EEGX_ERR_ARG,
                 "wake_dense: activation must be 0 (none), 1 (relu), 2 (sigmoid) or 3 (tanh)");
    if (n == 0) return EEGX_OK;
    EEGX_REQUIRE(w1 && b1 && w2 && b2 && x && label, EEGX_ERR_ARG, "wake_dense: NULL pointer");
    EEGX_REQUIRE(!dx || train, EEGX_ERR_ARG, "wake_dense: dx is produced by the backward pass (train = 1)");
    EEGX_REQUIRE(workspace && workspace_bytes >= eegx_wake_dense_workspace_bytes(in, hidden, n_cls, dx != nullptr),
                 EEGX_ERR_WORKSPACE, "wake_dense: workspace too small");
    int R = 0;
    const int G = grid_for(hidden, &R);
    const size_t smem = smem_bytes((int)in, R, (int)n_cls);
    EEGX_REQUIRE(smem <= 227 * 1024, EEGX_ERR_SHAPE,
                 "wake_dense: 2*in + n_cls*(rows_per_cta + 3) doubles (%zu bytes) exceed the 227 KB of shared memory", smem);
    EEGX_CUDA_CHECK(cudaFuncSetAttribute(wake_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    EEGX_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wake_dense_kernel, WK_THREADS, smem));
    EEGX_REQUIRE(per_sm >= 1, EEGX_ERR_CUDA, "wake_dense: kernel does not fit one CTA per SM");

    WakeParams p;
    p.w1 = w1; p.b1 = b1; p.w2 = w2; p.b2 = b2; p.x = x; p.label = label;
    p.n = n; p.in = (int)in; p.hidden = (int)hidden; p.ncls = (int)n_cls;
    p.lr = lr; p.act = activation; p.train = train;
    p.loss = loss; p.probs = probs; p.dx = dx;
    uintptr_t ws = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255);
    p.partial = reinterpret_cast<double*>(ws);
    p.dx_part = dx ? p.partial + 2 * (size_t)G * n_cls : nullptr;
    p.rows_per_cta = R;
    void* args[] = {&p};
    EEGX_CUDA_CHECK(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(wake_dense_kernel), dim3(G), dim3(WK_THREADS),
                                                args, smem, static_cast<cudaStream_t>(stream)));
    return EEGX_OK;
}

}  // extern "C"
