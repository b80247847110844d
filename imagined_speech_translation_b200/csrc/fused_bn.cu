// Glue of the CNN stack of Conv1DWithAttention (main_model/src/models/layers.py:142-178) on
// channels-last bf16 rows, forward and backward:
//   BatchNorm1d in train / eval mode over the valid rows (+ the 1x1-conv residual's BatchNorm)
//   + residual add + exact GELU + dropout + re-zeroing of the per-trial padding rows   (:142-174)
//   depthwise Conv1d k=5 (groups = channels)                                           (:157)
//   SqueezeExcite mean over time and the channel re-scaling + dropout                  (:177-178, :288-298)
//   (B, C, T) fp32 -> guarded channels-last bf16 (the transpose in front of conv1)     (:142)
//
// Layout ("guarded rows"): a (B, T, C) activation is stored as rows m = b*Tp + PAD + t of an
// (M = B*Tp) x C matrix, Tp = T + 2*PAD, whose other rows are zero, with PAD extra zero rows in
// front of m = 0 and after m = M-1.  A Conv1d is then a GEMM over overlapping rows (gemm_sm100.cu).
// Kernels take the pointer to row m = 0; "valid(m)" = (m mod Tp) in [PAD, PAD+T).
//
// All reductions over rows are two-stage with a fixed summation order (bit-stable, no atomics).
#include "eegx_common.h"
#include "fused_common.cuh"

namespace {

using namespace eegx;

struct RowGeom {
    long long M;   // G * B * Tp: all rows of the guarded buffer
    int Tp, lo, hi;  // valid rows: lo <= (m mod Tp) < hi
    long long Mg;  // rows per parameter group (B * Tp): rows [g*Mg, (g+1)*Mg) use parameter set g (the four
    int G;         // region encoders run as ONE guarded buffer of G*B trials; G = 1: a single module)
    __device__ __forceinline__ bool valid(long long m) const {
        const int r = (int)(m % Tp);
        return r >= lo && r < hi;
    }
};

constexpr int CR_THREADS = 256;   // 8 warps; a warp covers 64 columns (2 per lane)
constexpr int CR_COLS = 64;

// ---- column reduction skeleton: each lane owns 2 columns, warps stride over the rows of the
// CTA's slab, then the 8 warps are summed in order and the CTA partial is written.
struct NoPrep {
    __device__ __forceinline__ void operator()(int, int) const {}
};

// prep(c, grp) runs once per thread before the row loop (loop-invariant per-column parameters into registers)
template <int NACC, typename F, typename P = NoPrep>
__device__ __forceinline__ void col_reduce(const RowGeom& g, int C, float* __restrict__ part, F f, P prep = P()) {
    __shared__ float red[CR_THREADS / 32][NACC][CR_COLS];
    constexpr int NW = CR_THREADS / 32;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int c = blockIdx.x * CR_COLS + 2 * lane;
    // slabs never straddle a parameter group: gridDim.y = G * (slabs per group), group-major
    const int spg = gridDim.y / g.G;
    const int grp = blockIdx.y / spg;
    const long long rows_per = (g.Mg + spg - 1) / spg;
    const long long m0 = (long long)grp * g.Mg + (long long)(blockIdx.y - grp * spg) * rows_per;
    const long long mend = (long long)(grp + 1) * g.Mg;
    const long long m1 = m0 + rows_per < mend ? m0 + rows_per : mend;
    float acc[NACC][2];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k][0] = acc[k][1] = 0.0f;
    if (c < C && m0 + wid < m1) {
        prep(c, grp);
        // the position inside the trial is tracked incrementally (no 64-bit modulo per row)
        long long m = m0 + wid;
        int r = (int)(m % g.Tp);
        const int step = NW % g.Tp;
        const bool all_valid = g.lo == 0 && g.hi == g.Tp;
#pragma unroll 4
        for (; m < m1; m += NW) {
            if (all_valid || (r >= g.lo && r < g.hi)) f(m, c, grp, acc);
            r += step;
            if (r >= g.Tp) r -= g.Tp;
        }
    }
#pragma unroll
    for (int k = 0; k < NACC; ++k) {
        red[wid][k][2 * lane] = acc[k][0];
        red[wid][k][2 * lane + 1] = acc[k][1];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NACC * CR_COLS; i += CR_THREADS) {
        const int k = i / CR_COLS, cc = i % CR_COLS;
        if (blockIdx.x * CR_COLS + cc < C) {
            float t = 0.0f;
#pragma unroll
            for (int w = 0; w < NW; ++w) t += red[w][k][cc];
            part[((long long)blockIdx.y * NACC + k) * C + blockIdx.x * CR_COLS + cc] = t;
        }
    }
}

__device__ __forceinline__ float2 ld2(const __nv_bfloat16* p) {
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
}

// ------------------------------------------------------------------ BatchNorm statistics
__global__ void __launch_bounds__(CR_THREADS)
bn_stats_partial_kernel(const __nv_bfloat16* __restrict__ y, RowGeom g, int C, float* __restrict__ part) {
    EEGX_PDL_SYNC();
    col_reduce<2>(g, C, part, [&](long long m, int c, int, float (&acc)[2][2]) {
        const float2 v = ld2(y + m * C + c);
        acc[0][0] += v.x; acc[0][1] += v.y;
        acc[1][0] = fmaf(v.x, v.x, acc[1][0]); acc[1][1] = fmaf(v.y, v.y, acc[1][1]);
    });
}

// plain column sums of a (rows, C) bf16 matrix (bias gradients): same skeleton, every row valid
__global__ void __launch_bounds__(CR_THREADS)
colsum_partial_kernel(const __nv_bfloat16* __restrict__ y, long long ld, long long y_gstride, RowGeom g, int C,
                      float* __restrict__ part) {
    EEGX_PDL_SYNC();
    col_reduce<1>(g, C, part, [&](long long m, int c, int grp, float (&acc)[1][2]) {
        const float2 v = ld2(y + grp * y_gstride + (m - grp * g.Mg) * ld + c);
        acc[0][0] += v.x; acc[0][1] += v.y;
    });
}

// mean / rstd from the partials (fp64), and the running-statistics update of nn.BatchNorm1d
// (momentum; running_var takes the unbiased variance).
__global__ void bn_stats_final_kernel(const float* __restrict__ part, int nslabs, int C, double n_valid, float eps,
                                      float* __restrict__ mean, float* __restrict__ rstd,
                                      float* __restrict__ running_mean, float* __restrict__ running_var,
                                      long long running_gstride, float momentum) {
    EEGX_PDL_SYNC();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    // blockIdx.y = parameter group: its nslabs partials, its (C) slice of every statistics vector
    part += (long long)blockIdx.y * nslabs * 2 * C;
    mean += (long long)blockIdx.y * C;
    rstd += (long long)blockIdx.y * C;
    if (running_mean != nullptr) {
        running_mean += (long long)blockIdx.y * running_gstride;
        running_var += (long long)blockIdx.y * running_gstride;
    }
    double s = 0.0, q = 0.0;
    for (int b = 0; b < nslabs; ++b) {
        s += (double)part[((long long)b * 2 + 0) * C + c];
        q += (double)part[((long long)b * 2 + 1) * C + c];
    }
    const double mu = s / n_valid;
    double var = q / n_valid - mu * mu;
    if (var < 0.0) var = 0.0;
    mean[c] = (float)mu;
    rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean != nullptr) {
        const double unbiased = n_valid > 1.0 ? var * n_valid / (n_valid - 1.0) : var;
        running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * (float)mu;
        running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
}

// ------------------------------------------------------------------ BN (+ residual) + GELU + dropout
struct BnSide {
    const __nv_bfloat16* x;   // (M, C) rows from m = 0
    const float* mean;
    const float* rstd;
    const float* gamma;
    const float* beta;
};

// res_mode: 0 none, 1 identity (r.x added as is), 2 BatchNorm'd residual
__global__ void __launch_bounds__(256)
bn_act_fwd_kernel(BnSide a, BnSide r, int res_mode, __nv_bfloat16* __restrict__ out, RowGeom g, int pad, int C,
                  DropoutCfg dc) {
    EEGX_PDL_SYNC();
    const DropoutGen gen(dc);
    const int c8 = C >> 3;
    const long long total = (g.M + 2 * pad) * c8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long m = i / c8 - pad;
        const int c = (int)(i % c8) * 8;
        float o[8];
        if (m >= 0 && m < g.M && g.valid(m)) {
            float v[8], mu[8], rs[8], ga[8], be[8], msk[8];
            const int po = (int)(m / g.Mg) * C;          // this row's parameter group
            load8(a.x + m * C + c, v);
            load8f(a.mean + po + c, mu); load8f(a.rstd + po + c, rs); load8f(a.gamma + po + c, ga); load8f(a.beta + po + c, be);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = fmaf((v[e] - mu[e]) * rs[e], ga[e], be[e]);
            if (res_mode == 1) {
                load8(r.x + m * C + c, v);
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] += v[e];
            } else if (res_mode == 2) {
                load8(r.x + m * C + c, v);
                load8f(r.mean + po + c, mu); load8f(r.rstd + po + c, rs); load8f(r.gamma + po + c, ga); load8f(r.beta + po + c, be);
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] += fmaf((v[e] - mu[e]) * rs[e], ga[e], be[e]);
            }
            gen.mask8((unsigned long long)(m * C + c) >> 3, msk);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = gelu_f(o[e]) * msk[e];
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = 0.0f;
        }
        store8(out + m * C + c, o);
    }
}

// partial sums per channel: [0] sum dpre, [1] sum dpre * xhat_a, [2] sum dpre * xhat_r; dpre itself (the erf / exp /
// Philox work of this pass) is kept as bf16 in `dpre_out` so that the apply pass does not recompute it.  The
// per-column statistics and affine parameters are loop invariants held in registers.
__global__ void __launch_bounds__(CR_THREADS)
bn_act_bwd_reduce_kernel(BnSide a, BnSide r, int res_mode, const __nv_bfloat16* __restrict__ dout, RowGeom g, int C,
                         DropoutCfg dc, float* __restrict__ part, __nv_bfloat16* __restrict__ dpre_out) {
    EEGX_PDL_SYNC();
    const DropoutGen gen(dc);
    float mu_a[2], rs_a[2], ga_a[2], be_a[2], mu_r[2] = {0.f, 0.f}, rs_r[2] = {0.f, 0.f}, ga_r[2] = {0.f, 0.f}, be_r[2] = {0.f, 0.f};
    col_reduce<3>(g, C, part, [&](long long m, int c, int, float (&acc)[3][2]) {
        const float2 va = ld2(a.x + m * C + c);
        const float2 d = ld2(dout + m * C + c);
        const float ha0 = (va.x - mu_a[0]) * rs_a[0], ha1 = (va.y - mu_a[1]) * rs_a[1];
        float pre0 = fmaf(ha0, ga_a[0], be_a[0]), pre1 = fmaf(ha1, ga_a[1], be_a[1]);
        float hr0 = 0.0f, hr1 = 0.0f;
        if (res_mode == 1) {
            const float2 vr = ld2(r.x + m * C + c);
            pre0 += vr.x; pre1 += vr.y;
        } else if (res_mode == 2) {
            const float2 vr = ld2(r.x + m * C + c);
            hr0 = (vr.x - mu_r[0]) * rs_r[0]; hr1 = (vr.y - mu_r[1]) * rs_r[1];
            pre0 += fmaf(hr0, ga_r[0], be_r[0]); pre1 += fmaf(hr1, ga_r[1], be_r[1]);
        }
        float m0, m1;
        gen.mask_pair((unsigned long long)(m * C + c) >> 3, c & 7, m0, m1);
        const float dp0 = d.x * m0 * gelu_grad_f(pre0), dp1 = d.y * m1 * gelu_grad_f(pre1);
        *reinterpret_cast<__nv_bfloat162*>(dpre_out + m * C + c) = __floats2bfloat162_rn(dp0, dp1);
        acc[0][0] += dp0; acc[0][1] += dp1;
        acc[1][0] = fmaf(dp0, ha0, acc[1][0]); acc[1][1] = fmaf(dp1, ha1, acc[1][1]);
        acc[2][0] = fmaf(dp0, hr0, acc[2][0]); acc[2][1] = fmaf(dp1, hr1, acc[2][1]);
    }, [&](int c, int grp) {
        const int pc = grp * C + c;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            mu_a[e] = a.mean[pc + e]; rs_a[e] = a.rstd[pc + e]; ga_a[e] = a.gamma[pc + e]; be_a[e] = a.beta[pc + e];
            if (res_mode == 2) {
                mu_r[e] = r.mean[pc + e]; rs_r[e] = r.rstd[pc + e]; ga_r[e] = r.gamma[pc + e]; be_r[e] = r.beta[pc + e];
            }
        }
    });
}

// sums: (3, C) finished sums.  train != 0: batch-statistics backward; train == 0 (eval): the
// statistics are constants, dy = dpre * gamma * rstd.
// da / dr: rows m in [-pad, M + pad) are all written (zeros outside the valid rows).
__global__ void __launch_bounds__(256)
bn_act_bwd_apply_kernel(BnSide a, BnSide r, int res_mode, const __nv_bfloat16* __restrict__ dout,
                        const float* __restrict__ sums, float inv_n, int train, __nv_bfloat16* __restrict__ da,
                        __nv_bfloat16* __restrict__ dr, RowGeom g, int pad, int C, DropoutCfg dc) {
    EEGX_PDL_SYNC();
    const DropoutGen gen(dc);
    const int c2 = C >> 1;
    const long long total = (g.M + 2 * pad) * c2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long m = i / c2 - pad;
        const int c = (int)(i % c2) * 2;
        float oa[2] = {0.0f, 0.0f}, orr[2] = {0.0f, 0.0f};
        if (m >= 0 && m < g.M && g.valid(m)) {
            float dp[2], ha[2], hr[2] = {0.0f, 0.0f};
            const int grp = (int)(m / g.Mg), po = grp * C;
            const float* sg = sums + (long long)grp * 3 * C;       // sums: (G, 3, C)
            {   // dpre was left in `da` by the reduce pass (same element: read here, overwritten below)
                const float2 dpv = ld2(da + m * C + c);
                dp[0] = dpv.x; dp[1] = dpv.y;
                const float2 va = ld2(a.x + m * C + c);
                ha[0] = (va.x - a.mean[po + c]) * a.rstd[po + c];
                ha[1] = (va.y - a.mean[po + c + 1]) * a.rstd[po + c + 1];
                if (res_mode == 2) {
                    const float2 vr = ld2(r.x + m * C + c);
                    hr[0] = (vr.x - r.mean[po + c]) * r.rstd[po + c];
                    hr[1] = (vr.y - r.mean[po + c + 1]) * r.rstd[po + c + 1];
                }
            }
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const float s1 = train ? sg[c + e] * inv_n : 0.0f;
                const float s2a = train ? sg[C + c + e] * inv_n : 0.0f;
                oa[e] = a.gamma[po + c + e] * a.rstd[po + c + e] * (dp[e] - s1 - ha[e] * s2a);
                if (res_mode == 1) {
                    orr[e] = dp[e];
                } else if (res_mode == 2) {
                    const float s2r = train ? sg[2 * C + c + e] * inv_n : 0.0f;
                    orr[e] = r.gamma[po + c + e] * r.rstd[po + c + e] * (dp[e] - s1 - hr[e] * s2r);
                }
            }
        }
        *reinterpret_cast<__nv_bfloat162*>(da + m * C + c) = __floats2bfloat162_rn(oa[0], oa[1]);
        if (res_mode != 0) *reinterpret_cast<__nv_bfloat162*>(dr + m * C + c) = __floats2bfloat162_rn(orr[0], orr[1]);
    }
}

// out[k][c] = sum over slabs of part[slab][k][c]
// out[k][c] (+)= sum over slabs of part[slab][k][c], fixed order: a CTA owns 32 columns of one accumulator,
// its 8 warps each sum every 8th slab (coalesced rows), then the warps are added in order.
__global__ void __launch_bounds__(256)
sum_partials_kernel(const float* __restrict__ part, int nslabs, int nacc, int C, float* __restrict__ out,
                    long long out_gstride, int accumulate) {
    EEGX_PDL_SYNC();
    __shared__ float red[8][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane, k = blockIdx.y;
    // blockIdx.z = parameter group: its nslabs partials; its output block at out + z * out_gstride
    part += (long long)blockIdx.z * nslabs * nacc * C;
    out += (long long)blockIdx.z * out_gstride;
    float t = 0.0f;
    if (c < C)
        for (int b = wid; b < nslabs; b += 8) t += part[((long long)b * nacc + k) * C + c];
    red[wid][lane] = t;
    __syncthreads();
    if (wid == 0 && c < C) {
        float r = 0.0f;
#pragma unroll
        for (int w = 0; w < 8; ++w) r += red[w][lane];
        float* o = out + (long long)k * C + c;
        *o = accumulate ? *o + r : r;
    }
}

// dst[i] (+)= sum_s part[s][i]  (split-K weight-gradient partials folded into the gradient buffer)
__global__ void __launch_bounds__(256)
accumulate_partials_kernel(const float* __restrict__ part, int s, long long n4, long long per_group4,
                           long long dst_gstride4, float* __restrict__ dst, int accumulate) {
    // part: (s, n) contiguous with n = G * per_group; dst element (g, j) lives at g * dst_gstride + j
    EEGX_PDL_SYNC();
    const float4* p4 = reinterpret_cast<const float4*>(part);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const long long grp = i / per_group4;
        const long long di = grp * dst_gstride4 + (i - grp * per_group4);
        float4 a = accumulate ? d4[di] : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = 0; k < s; ++k) {
            const float4 v = __ldg(p4 + (long long)k * n4 + i);
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        d4[di] = a;
    }
}

// ------------------------------------------------------------------ depthwise Conv1d, k = 5
// w: (C, 5) fp32 (nn.Conv1d weight (C, 1, 5)), bias (C).  x, out: guarded rows.
__global__ void __launch_bounds__(256)
dwconv5_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                   __nv_bfloat16* __restrict__ out, RowGeom g, int pad, int C) {
    EEGX_PDL_SYNC();
    const int c8 = C >> 3;
    const long long total = (g.M + 2 * pad) * c8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long m = i / c8 - pad;
        const int c = (int)(i % c8) * 8;
        float o[8];
        if (m >= 0 && m < g.M && g.valid(m)) {
            const int po = (int)(m / g.Mg) * C;          // parameter group: w (G, C, 5), bias (G, C)
            load8f(bias + po + c, o);
#pragma unroll
            for (int tap = 0; tap < 5; ++tap) {
                float v[8];
                load8(x + (m + tap - 2) * C + c, v);     // rows m-2..m+2 exist (pad >= 2) and are zero outside trials
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = fmaf(v[e], w[(po + c + e) * 5 + tap], o[e]);
            }
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = 0.0f;
        }
        store8(out + m * C + c, o);
    }
}

// dx[m] = sum_tap w[tap] * dout[m - tap + 2]; dout is (M, C) from m = 0, zero on invalid rows
// (rows outside [0, M) are treated as zero); dx: guarded, all rows written, zero on invalid rows.
__global__ void __launch_bounds__(256)
dwconv5_bwd_data_kernel(const __nv_bfloat16* __restrict__ dout, const float* __restrict__ w,
                        __nv_bfloat16* __restrict__ dx, RowGeom g, int pad, int C) {
    EEGX_PDL_SYNC();
    const int c8 = C >> 3;
    const long long total = (g.M + 2 * pad) * c8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long m = i / c8 - pad;
        const int c = (int)(i % c8) * 8;
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = 0.0f;
        if (m >= 0 && m < g.M && g.valid(m)) {
            const int po = (int)(m / g.Mg) * C;
#pragma unroll
            for (int tap = 0; tap < 5; ++tap) {
                const long long mm = m - tap + 2;
                if (mm >= 0 && mm < g.M) {
                    float v[8];
                    load8(dout + mm * C + c, v);
#pragma unroll
                    for (int e = 0; e < 8; ++e) o[e] = fmaf(v[e], w[(po + c + e) * 5 + tap], o[e]);
                }
            }
        }
        store8(dx + m * C + c, o);
    }
}

// partials: [tap] sum_m dout[m] * x[m + tap - 2] (tap 0..4), [5] sum_m dout[m]
__global__ void __launch_bounds__(CR_THREADS)
dwconv5_bwd_weight_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ x, RowGeom g,
                          int C, float* __restrict__ part) {
    EEGX_PDL_SYNC();
    col_reduce<6>(g, C, part, [&](long long m, int c, int, float (&acc)[6][2]) {
        const float2 d = ld2(dout + m * C + c);
#pragma unroll
        for (int tap = 0; tap < 5; ++tap) {
            const float2 v = ld2(x + (m + tap - 2) * C + c);
            acc[tap][0] = fmaf(d.x, v.x, acc[tap][0]);
            acc[tap][1] = fmaf(d.y, v.y, acc[tap][1]);
        }
        acc[5][0] += d.x; acc[5][1] += d.y;
    });
}

// ------------------------------------------------------------------ SqueezeExcite pieces
// s[b][c] = (1/T) sum_t x[b, t, c]  (fp32); one CTA per (trial, 64 columns)
__global__ void __launch_bounds__(256)
group_mean_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ s, int Tp, int lo, int hi, int C) {
    EEGX_PDL_SYNC();
    __shared__ float red[8][CR_COLS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int c = blockIdx.x * CR_COLS + 2 * lane;
    const long long base = (long long)blockIdx.y * Tp;
    float a0 = 0.0f, a1 = 0.0f;
    if (c < C)
        for (int t = lo + wid; t < hi; t += 8) {
            const float2 v = ld2(x + (base + t) * C + c);
            a0 += v.x; a1 += v.y;
        }
    red[wid][2 * lane] = a0;
    red[wid][2 * lane + 1] = a1;
    __syncthreads();
    if (threadIdx.x < CR_COLS && blockIdx.x * CR_COLS + threadIdx.x < C) {
        float t = 0.0f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
        s[(long long)blockIdx.y * C + blockIdx.x * CR_COLS + threadIdx.x] = t / (float)(hi - lo);
    }
}

// ds (B, C) fp32 -> dx rows: dx[b, t, c] = ds[b, c] / T on valid rows (rows from m = 0; only valid rows written)
// fused with the SE scaling backward below via `dscale`.
// out[(b*T + t), c] = x[b, lo + t, c] * e[b, c] * dropout   (compact rows, bf16)
__global__ void __launch_bounds__(256)
se_scale_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ e, __nv_bfloat16* __restrict__ out,
                    long long B, int T, int Tp, int lo, int C, DropoutCfg dc) {
    EEGX_PDL_SYNC();
    const DropoutGen gen(dc);
    const int c8 = C >> 3;
    const long long total = B * T * c8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / c8;
        const int c = (int)(i % c8) * 8;
        const long long b = row / T;
        const int t = (int)(row % T);
        float v[8], ev[8], m[8], o[8];
        load8(x + (b * Tp + lo + t) * C + c, v);
        load8f(e + b * C + c, ev);
        gen.mask8((unsigned long long)i, m);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = v[k] * ev[k] * m[k];
        store8(out + row * C + c, o);
    }
}

// dx[b, lo + t, c] = dout[(b*T+t), c] * mask * e[b, c]   (valid rows only are written)
__global__ void __launch_bounds__(256)
se_scale_bwd_x_kernel(const __nv_bfloat16* __restrict__ dout, const float* __restrict__ e,
                      __nv_bfloat16* __restrict__ dx, long long B, int T, int Tp, int lo, int C, DropoutCfg dc) {
    EEGX_PDL_SYNC();
    const DropoutGen gen(dc);
    const int c8 = C >> 3;
    const long long total = B * T * c8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / c8;
        const int c = (int)(i % c8) * 8;
        const long long b = row / T;
        const int t = (int)(row % T);
        float d[8], ev[8], m[8], o[8];
        load8(dout + row * C + c, d);
        load8f(e + b * C + c, ev);
        gen.mask8((unsigned long long)i, m);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = d[k] * ev[k] * m[k];
        store8(dx + (b * Tp + lo + t) * C + c, o);
    }
}

// de[b][c] = sum_t dout[(b*T+t), c] * mask * x[b, lo + t, c]; one CTA per (trial, 64 columns)
__global__ void __launch_bounds__(256)
se_scale_bwd_e_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ x,
                      float* __restrict__ de, int T, int Tp, int lo, int C, DropoutCfg dc) {
    EEGX_PDL_SYNC();
    __shared__ float red[8][CR_COLS];
    const DropoutGen gen(dc);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int c = blockIdx.x * CR_COLS + 2 * lane;
    const long long b = blockIdx.y;
    const int c8 = C >> 3;
    float a0 = 0.0f, a1 = 0.0f;
    if (c < C)
        for (int t = wid; t < T; t += 8) {
            const float2 d = ld2(dout + (b * T + t) * C + c);
            const float2 v = ld2(x + (b * Tp + lo + t) * C + c);
            float m0, m1;
            gen.mask_pair((unsigned long long)((b * T + t) * c8 + (c >> 3)), c & 7, m0, m1);
            a0 = fmaf(d.x * m0, v.x, a0);
            a1 = fmaf(d.y * m1, v.y, a1);
        }
    red[wid][2 * lane] = a0;
    red[wid][2 * lane + 1] = a1;
    __syncthreads();
    if (threadIdx.x < CR_COLS && blockIdx.x * CR_COLS + threadIdx.x < C) {
        float t = 0.0f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
        de[b * C + blockIdx.x * CR_COLS + threadIdx.x] = t;
    }
}

// dx[b, lo + t, c] (+)= ds[b, c] / T on the valid rows (backward of group_mean)
__global__ void __launch_bounds__(256)
group_mean_bwd_kernel(const float* __restrict__ ds, __nv_bfloat16* __restrict__ dx, long long B, int T, int Tp,
                      int lo, int C, int accumulate) {
    EEGX_PDL_SYNC();
    const int c8 = C >> 3;
    const long long total = B * T * c8;
    const float inv_t = 1.0f / (float)T;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / c8;
        const int c = (int)(i % c8) * 8;
        const long long b = row / T;
        const int t = (int)(row % T);
        float d[8], o[8];
        load8f(ds + b * C + c, d);
        __nv_bfloat16* dst = dx + (b * Tp + lo + t) * C + c;
        if (accumulate) {
            load8(dst, o);
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = fmaf(d[k], inv_t, o[k]);
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = d[k] * inv_t;
        }
        store8(dst, o);
    }
}

// ------------------------------------------------------------------ (B, C, T) fp32 -> guarded channels-last bf16
// x: trial b at x + b * x_bstride, (C, T) contiguous.  out rows m in [-pad, M + pad) all written.
constexpr int TR_C = 64, TR_T = 32;
__global__ void __launch_bounds__(256)
nct_to_rows_kernel(const float* __restrict__ x, long long x_bstride, __nv_bfloat16* __restrict__ out, int T,
                   int Tp, int lo, int C) {
    EEGX_PDL_SYNC();
    __shared__ float tile[TR_C][TR_T + 1];
    const long long b = blockIdx.z;
    const int c0 = blockIdx.x * TR_C, t0 = blockIdx.y * TR_T;
    const float* src = x + b * x_bstride;
    for (int i = threadIdx.x; i < TR_C * TR_T; i += 256) {
        const int cc = i / TR_T, tt = i % TR_T;
        const int c = c0 + cc, t = t0 + tt;
        tile[cc][tt] = (c < C && t < T) ? src[(long long)c * T + t] : 0.0f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < TR_T * (TR_C / 2); i += 256) {
        const int tt = i / (TR_C / 2), cc = (i % (TR_C / 2)) * 2;
        const int c = c0 + cc, t = t0 + tt;
        if (c < C && t < T)
            *reinterpret_cast<__nv_bfloat162*>(out + (b * Tp + lo + t) * C + c) =
                __floats2bfloat162_rn(tile[cc][tt], tile[cc + 1][tt]);
    }
}

// zero the rows of a guarded buffer that are not valid (padding rows inside [0, M) and the guards)
__global__ void __launch_bounds__(256)
zero_invalid_rows_kernel(__nv_bfloat16* __restrict__ out, RowGeom g, int pad, int C) {
    EEGX_PDL_SYNC();
    const int c8 = C >> 3;
    const long long total = (g.M + 2 * pad) * c8;
    const float z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long m = i / c8 - pad;
        const int c = (int)(i % c8) * 8;
        if (!(m >= 0 && m < g.M && g.valid(m))) store8(out + m * C + c, z);
    }
}

// Conv1d weight gradient: the GEMM produces (s partials of) dW[co][tap * Cin + ci]; the parameter is laid out
// (Cout, Cin, k).  dst[co][ci][tap] (+)= sum_s part[s][co][tap * Cin + ci] -- partial folding, permutation and
// gradient accumulation in one pass.
__global__ void __launch_bounds__(256)
accumulate_conv_wgrad_kernel(const float* __restrict__ part, int s, long long n, int Cin, int k, float* __restrict__ dst,
                             int accumulate) {
    EEGX_PDL_SYNC();
    const long long per_co = (long long)Cin * k;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        // i indexes the destination (co, ci, tap): coalesced writes, strided (L2-resident) reads
        const long long co = i / per_co;
        const int rem = (int)(i - co * per_co);
        const int ci = rem / k, tap = rem - ci * k;
        const long long src = co * per_co + (long long)tap * Cin + ci;
        float a = accumulate ? dst[i] : 0.0f;
        for (int j = 0; j < s; ++j) a += __ldg(part + (long long)j * n + src);
        dst[i] = a;
    }
}

int ew_grid(long long n) {
    long long b = (n + 255) / 256;
    const long long cap = (long long)kNumSMsB200 * 8;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// slabs PER GROUP (rows Mg each); G * slabs <= 64 (the workspace bound)
int slabs_for(long long Mg, int C, int G = 1) {
    const int colblocks = (C + CR_COLS - 1) / CR_COLS;
    long long s = (4LL * kNumSMsB200 + (long long)colblocks * G - 1) / ((long long)colblocks * G);
    const long long max_by_rows = (Mg + 31) / 32;
    if (s > max_by_rows) s = max_by_rows;
    if (s > 64 / G) s = 64 / G;
    if (s < 1) s = 1;
    return (int)s;
}

}  // namespace

#define EEGX_GEOM_CHECK(name)                                                                                  \
    if (int rc = eegx::require_sm100()) return rc;                                                             \
    EEGX_REQUIRE(B >= 0 && T >= 1 && pad >= 2 && C >= 8 && (C % 8) == 0 && G >= 1 && G <= 64, EEGX_ERR_SHAPE,   \
                 name ": need T >= 1, pad >= 2, C a multiple of 8, 1 <= G <= 64");                             \
    const RowGeom g{(long long)G * B * (T + 2 * pad), (int)(T + 2 * pad), (int)pad, (int)(pad + T),            \
                    (long long)B * (T + 2 * pad), (int)G};                                                     \
    cudaStream_t st = static_cast<cudaStream_t>(stream);                                                       \
    (void)st

extern "C" {

size_t eegx_colreduce_workspace_bytes(int64_t C) { return (size_t)64 * 6 * (size_t)C * sizeof(float); }

int eegx_bn_stats_bf16(const void* y, int64_t G, int64_t B, int64_t T, int64_t pad, int64_t C, float eps, float* mean,
                       float* rstd, float* running_mean, float* running_var, int64_t running_gstride, float momentum,
                       void* workspace, size_t workspace_bytes, void* stream) {
    EEGX_GEOM_CHECK("bn_stats");
    if (B == 0) return EEGX_OK;
    EEGX_REQUIRE(y && mean && rstd && workspace, EEGX_ERR_ARG, "bn_stats: NULL pointer");
    EEGX_REQUIRE(workspace_bytes >= eegx_colreduce_workspace_bytes(C), EEGX_ERR_WORKSPACE, "bn_stats: workspace too small");
    const int slabs = slabs_for(g.Mg, (int)C, (int)G);
    float* part = static_cast<float*>(workspace);
    dim3 grid((unsigned)((C + CR_COLS - 1) / CR_COLS), (unsigned)(slabs * G));
    eegx::launch(bn_stats_partial_kernel, grid, CR_THREADS, 0, st, static_cast<const __nv_bfloat16*>(y), g, (int)C, part);
    eegx::launch(bn_stats_final_kernel, dim3((unsigned)((C + 127) / 128), (unsigned)G), 128, 0, st, part, slabs, (int)C,
                 (double)B * (double)T, eps, mean, rstd, running_mean, running_var, (long long)running_gstride, momentum);
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}

int eegx_accumulate_partials_f32(const float* part, int64_t s, int64_t G, int64_t n, float* dst, int64_t dst_gstride,
                                 int accumulate, void* stream) {
    if (int rc = eegx::require_sm100()) return rc;
    EEGX_REQUIRE(s >= 1 && G >= 1 && n >= 0 && (n % 4) == 0 && (dst_gstride % 4) == 0, EEGX_ERR_SHAPE,
                 "accumulate_partials: n and the group stride must be multiples of 4");
    if (n == 0) return EEGX_OK;
    EEGX_REQUIRE(part && dst, EEGX_ERR_ARG, "accumulate_partials: NULL pointer");
    EEGX_REQUIRE(eegx::aligned16(part) && eegx::aligned16(dst), EEGX_ERR_ALIGN, "accumulate_partials: 16-byte alignment");
    eegx::launch(accumulate_partials_kernel, ew_grid(G * n / 4), 256, 0, static_cast<cudaStream_t>(stream), part, (int)s,
                 (long long)(G * n / 4), (long long)(n / 4), (long long)((G > 1 ? dst_gstride : n) / 4), dst, accumulate);
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}

int eegx_accumulate_conv_wgrad_f32(const float* part, int64_t s, int64_t Cout, int64_t Cin, int64_t k, float* dst,
                                   int accumulate, void* stream) {
    if (int rc = eegx::require_sm100()) return rc;
    EEGX_REQUIRE(s >= 1 && Cout >= 0 && Cin >= 1 && k >= 1 && Cin * k < (1LL << 31), EEGX_ERR_SHAPE,
                 "accumulate_conv_wgrad: bad sizes");
    const long long n = Cout * Cin * k;
    if (n == 0) return EEGX_OK;
    EEGX_REQUIRE(part && dst, EEGX_ERR_ARG, "accumulate_conv_wgrad: NULL pointer");
    eegx::launch(accumulate_conv_wgrad_kernel, ew_grid(n), 256, 0, static_cast<cudaStream_t>(stream), part, (int)s, n, (int)Cin, (int)k,
                                                                                          dst, accumulate);
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}

int eegx_colsum_bf16(const void* y, int64_t ld, int64_t G, int64_t y_gstride, int64_t rows, int64_t C, float* out,
                     int64_t out_gstride, int accumulate, void* workspace, size_t workspace_bytes, void* stream) {
    if (int rc = eegx::require_sm100()) return rc;
    EEGX_REQUIRE(rows >= 0 && C >= 2 && (C % 2) == 0 && ld >= C && (ld % 2) == 0 && G >= 1 && G <= 64 &&
                     (y_gstride % 2) == 0, EEGX_ERR_SHAPE, "colsum: C, ld and the group stride must be even, ld >= C, 1 <= G <= 64");
    EEGX_REQUIRE(y && out && workspace, EEGX_ERR_ARG, "colsum: NULL pointer");
    EEGX_REQUIRE(workspace_bytes >= eegx_colreduce_workspace_bytes(C), EEGX_ERR_WORKSPACE, "colsum: workspace too small");
    const RowGeom g{(long long)rows * G, 1, 0, 1, (long long)rows, (int)G};     // rows per group, every row valid
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int slabs = slabs_for(g.Mg, (int)C, (int)G);
    float* part = static_cast<float*>(workspace);
    dim3 grid((unsigned)((C + CR_COLS - 1) / CR_COLS), (unsigned)(slabs * G));
    eegx::launch(colsum_partial_kernel, grid, CR_THREADS, 0, st, static_cast<const __nv_bfloat16*>(y), (long long)ld,
                 (long long)y_gstride, g, (int)C, part);
    eegx::launch(sum_partials_kernel, dim3((unsigned)((C + 31) / 32), 1, (unsigned)G), 256, 0, st, part, slabs, 1, (int)C, out,
                 (long long)out_gstride, accumulate);
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}

int eegx_bn_act_fwd_bf16(const void* ya, const float* mean_a, const float* rstd_a, const float* gamma_a,
                         const float* beta_a, const void* yr, const float* mean_r, const float* rstd_r,
                         const float* gamma_r, const float* beta_r, int res_mode, void* out, int64_t G, int64_t B, int64_t T,
                         int64_t pad, int64_t C, const uint64_t* rng_state, uint32_t site, float p, void* stream) {
    EEGX_GEOM_CHECK("bn_act_fwd");
    if (B == 0) return EEGX_OK;
    EEGX_REQUIRE(ya && mean_a && rstd_a && gamma_a && beta_a && out, EEGX_ERR_ARG, "bn_act_fwd: NULL pointer");
    EEGX_REQUIRE(res_mode == 0 || yr, EEGX_ERR_ARG, "bn_act_fwd: residual pointer missing");
    EEGX_REQUIRE(res_mode != 2 || (mean_r && rstd_r && gamma_r && beta_r), EEGX_ERR_ARG,
                 "bn_act_fwd: residual statistics missing");
    const BnSide a{static_cast<const __nv_bfloat16*>(ya), mean_a, rstd_a, gamma_a, beta_a};
    const BnSide r{static_cast<const __nv_bfloat16*>(yr), mean_r, rstd_r, gamma_r, beta_r};
    const DropoutCfg dc{reinterpret_cast<const unsigned long long*>(rng_state), site, p};
    eegx::launch(bn_act_fwd_kernel, ew_grid((g.M + 2 * pad) * (C / 8)), 256, 0, st, a, r, res_mode, static_cast<__nv_bfloat16*>(out),
                                                                          g, (int)pad, (int)C, dc);
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}

/* sums: (G, 3, C) fp32 out: [0] dbeta (both sides), [1] dgamma_a, [2] dgamma_r.  da / dr: guarded buffers
 * (pointer to row m = 0), every row written. */
int eegx_bn_act_bwd_bf16(const void* dout, const void* ya, const float* mean_a, const float* rstd_a,
                         const float* gamma_a, const float* beta_a, const void* yr, const float* mean_r,
                         const float* rstd_r, const float* gamma_r, const float* beta_r, int res_mode, int train,
                         void* da, void* dr, float* sums, void* workspace, size_t workspace_bytes, int64_t G, int64_t B,
                         int64_t T, int64_t pad, int64_t C, const uint64_t* rng_state, uint32_t site, float p,
                         void* stream) {
    EEGX_GEOM_CHECK("bn_act_bwd");
    if (B == 0) return EEGX_OK;
    EEGX_REQUIRE(dout && ya && mean_a && rstd_a && gamma_a && beta_a && da && sums && workspace, EEGX_ERR_ARG,
                 "bn_act_bwd: NULL pointer");
    EEGX_REQUIRE(res_mode == 0 || (yr && dr), EEGX_ERR_ARG, "bn_act_bwd: residual pointers missing");
    EEGX_REQUIRE(workspace_bytes >= eegx_colreduce_workspace_bytes(C), EEGX_ERR_WORKSPACE, "bn_act_bwd: workspace too small");
    const BnSide a{static_cast<const __nv_bfloat16*>(ya), mean_a, rstd_a, gamma_a, beta_a};
    const BnSide r{static_cast<const __nv_bfloat16*>(yr), mean_r, rstd_r, gamma_r, beta_r};
    const DropoutCfg dc{reinterpret_cast<const unsigned long long*>(rng_state), site, p};
    const int slabs = slabs_for(g.Mg, (int)C, (int)G);
    float* part = static_cast<float*>(workspace);
    dim3 grid((unsigned)((C + CR_COLS - 1) / CR_COLS), (unsigned)(slabs * G));
    eegx::launch(bn_act_bwd_reduce_kernel, grid, CR_THREADS, 0, st, a, r, res_mode, static_cast<const __nv_bfloat16*>(dout), g,
                                                          (int)C, dc, part, static_cast<__nv_bfloat16*>(da));
    eegx::launch(sum_partials_kernel, dim3((unsigned)((C + 31) / 32), 3, (unsigned)G), 256, 0, st, part, slabs, 3, (int)C, sums,
                 (long long)(3 * C), 0);
    eegx::launch(bn_act_bwd_apply_kernel, ew_grid((g.M + 2 * pad) * (C / 2)), 256, 0, st, 
        a, r, res_mode, static_cast<const __nv_bfloat16*>(dout), sums, 1.0f / (float)((double)B * (double)T), train,
        static_cast<__nv_bfloat16*>(da), static_cast<__nv_bfloat16*>(dr), g, (int)pad, (int)C, dc);
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}

int eegx_dwconv5_fwd_bf16(const void* x, const float* w, const float* bias, void* out, int64_t G, int64_t B, int64_t T,
                          int64_t pad, int64_t C, void* stream) {
    EEGX_GEOM_CHECK("dwconv5_fwd");
    if (B == 0) return EEGX_OK;
    EEGX_REQUIRE(x && w && bias && out, EEGX_ERR_ARG, "dwconv5_fwd: NULL pointer");
    eegx::launch(dwconv5_fwd_kernel, ew_grid((g.M + 2 * pad) * (C / 8)), 256, 0, st, 
        static_cast<const __nv_bfloat16*>(x), w, bias, static_cast<__nv_bfloat16*>(out), g, (int)pad, (int)C);
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}

/* dw: (G, C, 5) fp32, db: (G, C) fp32.  dwdb_scratch: (G, 6, C) fp32. */
int eegx_dwconv5_bwd_bf16(const void* dout, const void* x, const float* w, void* dx, float* dwdb_scratch,
                          void* workspace, size_t workspace_bytes, int64_t G, int64_t B, int64_t T, int64_t pad, int64_t C,
                          void* stream) {
    EEGX_GEOM_CHECK("dwconv5_bwd");
    if (B == 0) return EEGX_OK;
    EEGX_REQUIRE(dout && x && w && dx && dwdb_scratch && workspace, EEGX_ERR_ARG, "dwconv5_bwd: NULL pointer");
    EEGX_REQUIRE(workspace_bytes >= eegx_colreduce_workspace_bytes(C), EEGX_ERR_WORKSPACE, "dwconv5_bwd: workspace too small");
    eegx::launch(dwconv5_bwd_data_kernel, ew_grid((g.M + 2 * pad) * (C / 8)), 256, 0, st, 
        static_cast<const __nv_bfloat16*>(dout), w, static_cast<__nv_bfloat16*>(dx), g, (int)pad, (int)C);
    const int slabs = slabs_for(g.Mg, (int)C, (int)G);
    float* part = static_cast<float*>(workspace);
    dim3 grid((unsigned)((C + CR_COLS - 1) / CR_COLS), (unsigned)(slabs * G));
    eegx::launch(dwconv5_bwd_weight_kernel, grid, CR_THREADS, 0, st, static_cast<const __nv_bfloat16*>(dout),
                                                           static_cast<const __nv_bfloat16*>(x), g, (int)C, part);
    eegx::launch(sum_partials_kernel, dim3((unsigned)((C + 31) / 32), 6, (unsigned)G), 256, 0, st, part, slabs, 6, (int)C,
                 dwdb_scratch, (long long)(6 * C), 0);
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}

int eegx_group_mean_bf16(const void* x, float* s, int64_t B, int64_t T, int64_t pad, int64_t C, void* stream) {
    const int64_t G = 1;                  // per-trial kernel: the caller passes all G * B trials as B
    EEGX_GEOM_CHECK("group_mean");
    if (B == 0) return EEGX_OK;
    EEGX_REQUIRE(x && s, EEGX_ERR_ARG, "group_mean: NULL pointer");
    dim3 grid((unsigned)((C + CR_COLS - 1) / CR_COLS), (unsigned)B);
    eegx::launch(group_mean_kernel, grid, 256, 0, st, static_cast<const __nv_bfloat16*>(x), s, g.Tp, g.lo, g.hi, (int)C);
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}

int eegx_group_mean_bwd_bf16(const float* ds, void* dx, int64_t B, int64_t T, int64_t pad, int64_t C,
                             int accumulate, void* stream) {
    const int64_t G = 1;                  // per-trial kernel: the caller passes all G * B trials as B
    EEGX_GEOM_CHECK("group_mean_bwd");
    if (B == 0) return EEGX_OK;
    EEGX_REQUIRE(ds && dx, EEGX_ERR_ARG, "group_mean_bwd: NULL pointer");
    eegx::launch(group_mean_bwd_kernel, ew_grid(B * T * (C / 8)), 256, 0, st, ds, static_cast<__nv_bfloat16*>(dx), B, (int)T, g.Tp,
                                                                    g.lo, (int)C, accumulate);
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}

int eegx_se_scale_fwd_bf16(const void* x, const float* e, void* out, int64_t B, int64_t T, int64_t pad, int64_t C,
                           const uint64_t* rng_state, uint32_t site, float p, void* stream) {
    const int64_t G = 1;                  // per-trial kernel: the caller passes all G * B trials as B
    EEGX_GEOM_CHECK("se_scale_fwd");
    if (B == 0) return EEGX_OK;
    EEGX_REQUIRE(x && e && out, EEGX_ERR_ARG, "se_scale_fwd: NULL pointer");
    const DropoutCfg dc{reinterpret_cast<const unsigned long long*>(rng_state), site, p};
    eegx::launch(se_scale_fwd_kernel, ew_grid(B * T * (C / 8)), 256, 0, st, static_cast<const __nv_bfloat16*>(x), e,
                                                                  static_cast<__nv_bfloat16*>(out), B, (int)T, g.Tp, g.lo,
                                                                  (int)C, dc);
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}

/* dx: guarded rows pointer (m = 0), only the valid rows are written; de: (B, C) fp32. */
int eegx_se_scale_bwd_bf16(const void* dout, const void* x, const float* e, void* dx, float* de, int64_t B,
                           int64_t T, int64_t pad, int64_t C, const uint64_t* rng_state, uint32_t site, float p,
                           void* stream) {
    const int64_t G = 1;                  // per-trial kernel: the caller passes all G * B trials as B
    EEGX_GEOM_CHECK("se_scale_bwd");
    if (B == 0) return EEGX_OK;
    EEGX_REQUIRE(dout && x && e && dx && de, EEGX_ERR_ARG, "se_scale_bwd: NULL pointer");
    const DropoutCfg dc{reinterpret_cast<const unsigned long long*>(rng_state), site, p};
    eegx::launch(se_scale_bwd_x_kernel, ew_grid(B * T * (C / 8)), 256, 0, st, static_cast<const __nv_bfloat16*>(dout), e,
                                                                    static_cast<__nv_bfloat16*>(dx), B, (int)T, g.Tp,
                                                                    g.lo, (int)C, dc);
    dim3 grid((unsigned)((C + CR_COLS - 1) / CR_COLS), (unsigned)B);
    eegx::launch(se_scale_bwd_e_kernel, grid, 256, 0, st, static_cast<const __nv_bfloat16*>(dout),
                                                static_cast<const __nv_bfloat16*>(x), de, (int)T, g.Tp, g.lo, (int)C, dc);
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}

/* x: (B, C, T) fp32 with batch stride x_bstride (elements); out: guarded rows pointer (m = 0); every row of
 * the guarded buffer is written (zeros outside the valid rows). */
int eegx_nct_to_rows_bf16(const float* x, int64_t x_bstride, void* out, int64_t B, int64_t T, int64_t pad,
                          int64_t C, void* stream) {
    const int64_t G = 1;                  // per-trial kernel: the caller passes all G * B trials as B
    EEGX_GEOM_CHECK("nct_to_rows");
    if (B == 0) return EEGX_OK;
    EEGX_REQUIRE(x && out, EEGX_ERR_ARG, "nct_to_rows: NULL pointer");
    EEGX_REQUIRE(B <= 65535, EEGX_ERR_SHAPE, "nct_to_rows: B must be <= 65535");
    eegx::launch(zero_invalid_rows_kernel, ew_grid((g.M + 2 * pad) * (C / 8)), 256, 0, st, static_cast<__nv_bfloat16*>(out), g,
                                                                                  (int)pad, (int)C);
    dim3 grid((unsigned)((C + TR_C - 1) / TR_C), (unsigned)((T + TR_T - 1) / TR_T), (unsigned)B);
    eegx::launch(nct_to_rows_kernel, grid, 256, 0, st, x, x_bstride, static_cast<__nv_bfloat16*>(out), (int)T, g.Tp, g.lo, (int)C);
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}

}  // extern "C"
