// Row-wise / element-wise glue of the encoder, forward and backward, bf16 I/O with fp32 math.
// Each kernel is one read + one write of the activation (HBM-bound); together they replace the
// chains of ATen element-wise kernels behind
//   nn.LayerNorm (+ nn.GELU + nn.Dropout)        main_model/src/models/layers.py:61-71, 84-127, 232-242
//   FeedForwardNetwork's gelu(W1 x) * sigmoid(Wg x) + dropout      layers.py:311-317
//   the residual adds  x + dropout(f(x)),  x + 0.1 * cross(x)      layers.py:234-251
//   nn.GELU + nn.Dropout after a Linear                            brain_encoder.py:36-75
// Dropout masks are regenerated from (seed, step, site, element group) -- fused_common.cuh.
#include "eegx_common.h"
#include "fused_common.cuh"

namespace {

using namespace eegx;

constexpr int LN_WARPS = 4;
constexpr int LN_MAX_C = 2048;

// ------------------------------------------------------------------------------------------
// LayerNorm (+ GELU) (+ dropout): one warp per row, the row lives in registers.
// ------------------------------------------------------------------------------------------
template <int NVEC>
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
              __nv_bfloat16* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out,
              long long rows, long long rows_per_group, long long pstride, int C, float eps, int act, DropoutCfg dc) {
    EEGX_PDL_SYNC();
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * LN_WARPS;
    const DropoutGen gen(dc);
    const float inv_c = 1.0f / (float)C;
    for (long long row = warp; row < rows; row += nwarps) {
        float v[NVEC][8];
        float s = 0.0f;
#pragma unroll
        for (int i = 0; i < NVEC; ++i) {
            const int col = (lane + 32 * i) * 8;
            if (col < C) {
                load8(x + row * C + col, v[i]);
#pragma unroll
                for (int e = 0; e < 8; ++e) s += v[i][e];
            }
        }
        const float mean = warp_sum_f(s) * inv_c;
        float q = 0.0f;
#pragma unroll
        for (int i = 0; i < NVEC; ++i) {
            const int col = (lane + 32 * i) * 8;
            if (col < C) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float d = v[i][e] - mean;
                    q = fmaf(d, d, q);
                }
            }
        }
        const float rstd = rsqrtf(warp_sum_f(q) * inv_c + eps);
#pragma unroll
        for (int i = 0; i < NVEC; ++i) {
            const int col = (lane + 32 * i) * 8;
            if (col < C) {
                float g[8], b[8], m[8], o[8];
                const long long po = (row / rows_per_group) * pstride;      // this row's parameter set
                load8f(gamma + po + col, g);
                load8f(beta + po + col, b);
                gen.mask8((unsigned long long)(row * C + col) >> 3, m);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    float t = fmaf((v[i][e] - mean) * rstd, g[e], b[e]);
                    if (act) t = gelu_f(t);
                    o[e] = t * m[e];
                }
                store8(y + row * C + col, o);
            }
        }
        if (lane == 0) {
            mean_out[row] = mean;
            rstd_out[row] = rstd;
        }
    }
}

template <int NVEC>
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
              const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ mean_in,
              const float* __restrict__ rstd_in, __nv_bfloat16* __restrict__ dx, float* __restrict__ part,
              long long rows_per_group, long long pstride, int C, int act, DropoutCfg dc) {
    EEGX_PDL_SYNC();
    __shared__ float red[LN_WARPS][LN_MAX_C];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long warp = (long long)blockIdx.x * LN_WARPS + wid;
    const long long nwarps = (long long)gridDim.x * LN_WARPS;
    // blockIdx.y = parameter group: its rows, its gamma / beta, its block of partials
    const long long row_lo = (long long)blockIdx.y * rows_per_group, rows = row_lo + rows_per_group;
    gamma += (long long)blockIdx.y * pstride;
    beta += (long long)blockIdx.y * pstride;
    part += (long long)blockIdx.y * gridDim.x * 2 * C;
    const DropoutGen gen(dc);
    const float inv_c = 1.0f / (float)C;
    float dg[NVEC][8], db[NVEC][8];
#pragma unroll
    for (int i = 0; i < NVEC; ++i)
#pragma unroll
        for (int e = 0; e < 8; ++e) dg[i][e] = db[i][e] = 0.0f;

    for (long long row = row_lo + warp; row < rows; row += nwarps) {
        const float mean = mean_in[row], rstd = rstd_in[row];
        float xh[NVEC][8], dxh[NVEC][8];
        float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
        for (int i = 0; i < NVEC; ++i) {
            const int col = (lane + 32 * i) * 8;
            if (col < C) {
                float xv[8], d[8], g[8], m[8];
                load8(x + row * C + col, xv);
                load8(dy + row * C + col, d);
                load8f(gamma + col, g);
                gen.mask8((unsigned long long)(row * C + col) >> 3, m);
                float b[8];
                if (act) load8f(beta + col, b);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float h = (xv[e] - mean) * rstd;
                    float dp = d[e] * m[e];
                    if (act) dp *= gelu_grad_f(fmaf(h, g[e], b[e]));
                    dg[i][e] = fmaf(dp, h, dg[i][e]);
                    db[i][e] += dp;
                    const float t = dp * g[e];
                    xh[i][e] = h;
                    dxh[i][e] = t;
                    s1 += t;
                    s2 = fmaf(t, h, s2);
                }
            }
        }
        s1 = warp_sum_f(s1) * inv_c;
        s2 = warp_sum_f(s2) * inv_c;
#pragma unroll
        for (int i = 0; i < NVEC; ++i) {
            const int col = (lane + 32 * i) * 8;
            if (col < C) {
                float o[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = rstd * (dxh[i][e] - s1 - xh[i][e] * s2);
                store8(dx + row * C + col, o);
            }
        }
    }
    // CTA partials of dgamma then dbeta, summed over the warps in a fixed order
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
        __syncthreads();
#pragma unroll
        for (int i = 0; i < NVEC; ++i) {
            const int col = (lane + 32 * i) * 8;
            if (col < C) {
#pragma unroll
                for (int e = 0; e < 8; ++e) red[wid][col + e] = which == 0 ? dg[i][e] : db[i][e];
            }
        }
        __syncthreads();
        for (int c = threadIdx.x; c < C; c += LN_WARPS * 32) {
            float t = 0.0f;
#pragma unroll
            for (int w = 0; w < LN_WARPS; ++w) t += red[w][c];
            part[((long long)blockIdx.x * 2 + which) * C + c] = t;
        }
    }
}

// out_j[c] = sum_b part[b][j][c] in a fixed order (bit-stable): a CTA owns 32 columns, its 8 warps
// each sum every 8th block (coalesced 128-byte rows), then the warps are added in order.
__global__ void __launch_bounds__(256)
colsum_partials_kernel(const float* __restrict__ part, int nblocks, int nvec, int C,
                       float* __restrict__ out0, float* __restrict__ out1, long long out_gstride, int accumulate) {
    EEGX_PDL_SYNC();
    __shared__ float red[8][32];
    part += (long long)blockIdx.z * nblocks * nvec * C;      // blockIdx.z = parameter group
    out0 += (long long)blockIdx.z * out_gstride;
    out1 += (long long)blockIdx.z * out_gstride;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane, j = blockIdx.y;
    float t = 0.0f;
    if (c < C)
        for (int b = wid; b < nblocks; b += 8) t += part[((long long)b * nvec + j) * C + c];
    red[wid][lane] = t;
    __syncthreads();
    if (wid == 0 && c < C) {
        float r = 0.0f;
#pragma unroll
        for (int w = 0; w < 8; ++w) r += red[w][lane];
        float* o = (j == 0 ? out0 : out1) + c;
        *o = accumulate ? *o + r : r;
    }
}

// ------------------------------------------------------------------------------------------
// element-wise kernels on groups of 8
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
add_dropout_fwd_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                       __nv_bfloat16* __restrict__ out, long long n8, float scale, DropoutCfg dc) {
    EEGX_PDL_SYNC();
    const DropoutGen gen(dc);
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n8; g += (long long)gridDim.x * blockDim.x) {
        float av[8], bv[8], m[8], o[8];
        load8(a + g * 8, av);
        load8(b + g * 8, bv);
        gen.mask8((unsigned long long)g, m);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = fmaf(bv[e] * m[e], scale, av[e]);
        store8(out + g * 8, o);
    }
}

__global__ void __launch_bounds__(256)
dropout_scale_kernel(const __nv_bfloat16* __restrict__ din, __nv_bfloat16* __restrict__ dout, long long n8,
                     float scale, DropoutCfg dc) {
    EEGX_PDL_SYNC();
    const DropoutGen gen(dc);
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n8; g += (long long)gridDim.x * blockDim.x) {
        float v[8], m[8], o[8];
        load8(din + g * 8, v);
        gen.mask8((unsigned long long)g, m);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = v[e] * m[e] * scale;
        store8(dout + g * 8, o);
    }
}

__global__ void __launch_bounds__(256)
gelu_dropout_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, long long n8,
                        DropoutCfg dc) {
    EEGX_PDL_SYNC();
    const DropoutGen gen(dc);
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n8; g += (long long)gridDim.x * blockDim.x) {
        float v[8], m[8], o[8];
        load8(x + g * 8, v);
        gen.mask8((unsigned long long)g, m);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = gelu_f(v[e]) * m[e];
        store8(out + g * 8, o);
    }
}

__global__ void __launch_bounds__(256)
gelu_dropout_bwd_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ x,
                        __nv_bfloat16* __restrict__ dx, long long n8, DropoutCfg dc) {
    EEGX_PDL_SYNC();
    const DropoutGen gen(dc);
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n8; g += (long long)gridDim.x * blockDim.x) {
        float v[8], d[8], m[8], o[8];
        load8(x + g * 8, v);
        load8(dout + g * 8, d);
        gen.mask8((unsigned long long)g, m);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = d[e] * m[e] * gelu_grad_f(v[e]);
        store8(dx + g * 8, o);
    }
}

// ag: (rows, 2H) = [a | g];  out: (rows, H) = dropout(gelu(a) * sigmoid(g))
__global__ void __launch_bounds__(256)
glu_fwd_kernel(const __nv_bfloat16* __restrict__ ag, __nv_bfloat16* __restrict__ out, long long rows, int H,
               DropoutCfg dc) {
    EEGX_PDL_SYNC();
    const DropoutGen gen(dc);
    const int h8 = H >> 3;
    const long long n8 = rows * h8;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n8; g += (long long)gridDim.x * blockDim.x) {
        const long long r = g / h8;
        const int c = (int)(g - r * h8) * 8;
        float a[8], gt[8], m[8], o[8];
        load8(ag + r * 2 * H + c, a);
        load8(ag + r * 2 * H + H + c, gt);
        gen.mask8((unsigned long long)g, m);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = gelu_f(a[e]) * sigmoid_f(gt[e]) * m[e];
        store8(out + r * H + c, o);
    }
}

__global__ void __launch_bounds__(256)
glu_bwd_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ ag,
               __nv_bfloat16* __restrict__ dag, long long rows, int H, DropoutCfg dc) {
    EEGX_PDL_SYNC();
    const DropoutGen gen(dc);
    const int h8 = H >> 3;
    const long long n8 = rows * h8;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n8; g += (long long)gridDim.x * blockDim.x) {
        const long long r = g / h8;
        const int c = (int)(g - r * h8) * 8;
        float a[8], gt[8], d[8], m[8], da[8], dg[8];
        load8(ag + r * 2 * H + c, a);
        load8(ag + r * 2 * H + H + c, gt);
        load8(dout + r * H + c, d);
        gen.mask8((unsigned long long)g, m);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float s = sigmoid_f(gt[e]);
            const float t = d[e] * m[e];
            da[e] = t * s * gelu_grad_f(a[e]);
            dg[e] = t * gelu_f(a[e]) * s * (1.0f - s);
        }
        store8(dag + r * 2 * H + c, da);
        store8(dag + r * 2 * H + H + c, dg);
    }
}

// ------------------------------------------------------------------------------------------
// Token assembly in front of the attention stack (layers.py:214-225):
//   out[gb, s, :] = (s == 0 ? cls[g] : s < 4 ? temporal[g][s - 1] : h[gb, s - 4]) + pos[g][s],   g = gb / B
// for G parameter groups of B sequences each (G = 1: a single module).  The tokens are rounded to bf16 before the
// add, as torch.cat((tokens.to(bf16), h)) + pos does.  Backward: dh = dout[:, 4:] (below); d(pos) = column sums of
// dout over the B sequences of a group (eegx_colsum_bf16), d(cls) / d(temporal) = its first rows.
// ------------------------------------------------------------------------------------------
constexpr int N_TOK = 4;

__global__ void __launch_bounds__(256)
assemble_tokens_fwd_kernel(const __nv_bfloat16* __restrict__ h, const float* __restrict__ cls, long long cls_gs,
                           const float* __restrict__ temporal, long long tmp_gs, const float* __restrict__ pos,
                           long long pos_gs, __nv_bfloat16* __restrict__ out, long long GB, int B, int T, int d) {
    EEGX_PDL_SYNC();
    const int d8 = d >> 3, S = T + N_TOK;
    const long long total = GB * S * d8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / d8;
        const int c = (int)(i - row * d8) * 8;
        const long long gb = row / S;
        const int sidx = (int)(row - gb * S);
        const long long g = gb / B;
        float v[8], pe[8], o[8];
        if (sidx >= N_TOK) {
            load8(h + (gb * T + sidx - N_TOK) * d + c, v);
        } else {
            if (sidx == 0) load8f(cls + g * cls_gs + c, v);
            else load8f(temporal + g * tmp_gs + (long long)(sidx - 1) * d + c, v);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = __bfloat162float(__float2bfloat16(v[e]));
        }
        load8f(pos + g * pos_gs + (long long)sidx * d + c, pe);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = v[e] + pe[e];
        store8(out + row * d + c, o);
    }
}

// dh[gb, t, :] = dout[gb, t + 4, :]
__global__ void __launch_bounds__(256)
assemble_tokens_bwd_kernel(const __nv_bfloat16* __restrict__ dout, __nv_bfloat16* __restrict__ dh, long long GB, int T,
                           int d) {
    EEGX_PDL_SYNC();
    const int d8 = d >> 3, S = T + N_TOK;
    const long long total = GB * T * d8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / d8;
        const int c = (int)(i - row * d8) * 8;
        const long long gb = row / T;
        const int t = (int)(row - gb * T);
        *reinterpret_cast<uint4*>(dh + row * d + c) =
            *reinterpret_cast<const uint4*>(dout + (gb * S + t + N_TOK) * d + c);
    }
}

int ew_grid(long long n8) {
    long long b = (n8 + 255) / 256;
    const long long cap = (long long)kNumSMsB200 * 8;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// forward: one row per warp, as many CTAs as rows allow (memory bound: occupancy matters);
// backward: capped, every CTA leaves a (2, C) partial for the fixed-order finalize.
constexpr int LN_BWD_CTAS_PER_SM = 2;
int ln_grid(long long rows, bool bwd, int groups = 1) {     // CTAs per group (backward) / in total (forward)
    long long b = (rows + LN_WARPS - 1) / LN_WARPS;
    long long cap = (long long)kNumSMsB200 * (bwd ? LN_BWD_CTAS_PER_SM : 16);
    if (bwd) cap = cap / groups < 1 ? 1 : cap / groups;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

template <int NVEC>
void launch_ln_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                   long long rows, long long rpg, long long pstride, int C, float eps, int act, DropoutCfg dc,
                   cudaStream_t st) {
    eegx::launch(ln_fwd_kernel<NVEC>, ln_grid(rows, false), LN_WARPS * 32, 0, st,
        static_cast<const __nv_bfloat16*>(x), gamma, beta, static_cast<__nv_bfloat16*>(y), mean, rstd, rows, rpg, pstride,
        C, eps, act, dc);
}

template <int NVEC>
void launch_ln_bwd(const void* dy, const void* x, const float* gamma, const float* beta, const float* mean,
                   const float* rstd, void* dx, float* part, int grid, int groups, long long rpg, long long pstride,
                   int C, int act, DropoutCfg dc, cudaStream_t st) {
    eegx::launch(ln_bwd_kernel<NVEC>, dim3((unsigned)grid, (unsigned)groups), LN_WARPS * 32, 0, st,
        static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(x), gamma, beta, mean, rstd,
        static_cast<__nv_bfloat16*>(dx), part, rpg, pstride, C, act, dc);
}

}  // namespace

#define EEGX_EW_CHECK(n)                                                                            \
    if (int rc = eegx::require_sm100()) return rc;                                                   \
    EEGX_REQUIRE((n) >= 0 && ((n) % 8) == 0, EEGX_ERR_SHAPE, "element count must be a multiple of 8")

extern "C" {

int eegx_layernorm_fwd_bf16(const void* x, const float* gamma, const float* beta, void* y, float* mean,
                            float* rstd, int64_t rows, int64_t C, int64_t groups, int64_t param_stride, float eps, int act,
                            const uint64_t* rng_state, uint32_t site, float p, void* stream) {
    if (int rc = eegx::require_sm100()) return rc;
    EEGX_REQUIRE(rows >= 0 && C >= 8 && C <= LN_MAX_C && (C % 8) == 0, EEGX_ERR_SHAPE,
                 "layernorm: C must be a multiple of 8 in [8, %d]", LN_MAX_C);
    EEGX_REQUIRE(groups >= 1 && rows % groups == 0 && (param_stride % 4) == 0, EEGX_ERR_SHAPE,
                 "layernorm: rows must split evenly over the groups; parameter stride a multiple of 4");
    if (rows == 0) return EEGX_OK;
    const long long rpg = rows / groups, ps = groups > 1 ? param_stride : 0;
    EEGX_REQUIRE(x && gamma && beta && y && mean && rstd, EEGX_ERR_ARG, "layernorm: NULL pointer");
    EEGX_REQUIRE(eegx::aligned16(x) && eegx::aligned16(y) && eegx::aligned16(gamma) && eegx::aligned16(beta),
                 EEGX_ERR_ALIGN, "layernorm: pointers must be 16-byte aligned");
    const DropoutCfg dc{reinterpret_cast<const unsigned long long*>(rng_state), site, p};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int nvec = (int)((C + 255) / 256);
    switch (nvec) {
        case 1: launch_ln_fwd<1>(x, gamma, beta, y, mean, rstd, rows, rpg, ps, (int)C, eps, act, dc, st); break;
        case 2: launch_ln_fwd<2>(x, gamma, beta, y, mean, rstd, rows, rpg, ps, (int)C, eps, act, dc, st); break;
        case 3: launch_ln_fwd<3>(x, gamma, beta, y, mean, rstd, rows, rpg, ps, (int)C, eps, act, dc, st); break;
        case 4: launch_ln_fwd<4>(x, gamma, beta, y, mean, rstd, rows, rpg, ps, (int)C, eps, act, dc, st); break;
        case 5: case 6: launch_ln_fwd<6>(x, gamma, beta, y, mean, rstd, rows, rpg, ps, (int)C, eps, act, dc, st); break;
        default: launch_ln_fwd<8>(x, gamma, beta, y, mean, rstd, rows, rpg, ps, (int)C, eps, act, dc, st); break;
    }
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}

size_t eegx_layernorm_bwd_workspace_bytes(int64_t C) {
    return (size_t)kNumSMsB200 * LN_BWD_CTAS_PER_SM * 2 * (size_t)C * sizeof(float);
}

int eegx_layernorm_bwd_bf16(const void* dy, const void* x, const float* gamma, const float* beta,
                            const float* mean, const float* rstd, void* dx, float* dgamma, float* dbeta,
                            int accumulate, void* workspace, size_t workspace_bytes, int64_t rows, int64_t C,
                            int64_t groups, int64_t param_stride, int act, const uint64_t* rng_state, uint32_t site,
                            float p, void* stream) {
    if (int rc = eegx::require_sm100()) return rc;
    EEGX_REQUIRE(rows >= 0 && C >= 8 && C <= LN_MAX_C && (C % 8) == 0, EEGX_ERR_SHAPE,
                 "layernorm: C must be a multiple of 8 in [8, %d]", LN_MAX_C);
    EEGX_REQUIRE(groups >= 1 && groups <= kNumSMsB200 && rows % groups == 0 && (param_stride % 4) == 0, EEGX_ERR_SHAPE,
                 "layernorm bwd: rows must split evenly over the groups; parameter stride a multiple of 4");
    const long long rpg = rows / groups, ps = groups > 1 ? param_stride : 0;
    EEGX_REQUIRE(dy && x && gamma && beta && mean && rstd && dx && dgamma && dbeta && workspace, EEGX_ERR_ARG,
                 "layernorm bwd: NULL pointer");
    EEGX_REQUIRE(workspace_bytes >= eegx_layernorm_bwd_workspace_bytes(C), EEGX_ERR_WORKSPACE,
                 "layernorm bwd: workspace too small");
    EEGX_REQUIRE(eegx::aligned16(x) && eegx::aligned16(dy) && eegx::aligned16(dx) && eegx::aligned16(gamma) &&
                     eegx::aligned16(beta), EEGX_ERR_ALIGN, "layernorm bwd: pointers must be 16-byte aligned");
    const DropoutCfg dc{reinterpret_cast<const unsigned long long*>(rng_state), site, p};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = ln_grid(rpg, true, (int)groups);
    float* part = static_cast<float*>(workspace);
    const int nvec = (int)((C + 255) / 256);
    switch (nvec) {
        case 1: launch_ln_bwd<1>(dy, x, gamma, beta, mean, rstd, dx, part, grid, (int)groups, rpg, ps, (int)C, act, dc, st); break;
        case 2: launch_ln_bwd<2>(dy, x, gamma, beta, mean, rstd, dx, part, grid, (int)groups, rpg, ps, (int)C, act, dc, st); break;
        case 3: launch_ln_bwd<3>(dy, x, gamma, beta, mean, rstd, dx, part, grid, (int)groups, rpg, ps, (int)C, act, dc, st); break;
        case 4: launch_ln_bwd<4>(dy, x, gamma, beta, mean, rstd, dx, part, grid, (int)groups, rpg, ps, (int)C, act, dc, st); break;
        case 5: case 6: launch_ln_bwd<6>(dy, x, gamma, beta, mean, rstd, dx, part, grid, (int)groups, rpg, ps, (int)C, act, dc, st); break;
        default: launch_ln_bwd<8>(dy, x, gamma, beta, mean, rstd, dx, part, grid, (int)groups, rpg, ps, (int)C, act, dc, st); break;
    }
    eegx::launch(colsum_partials_kernel, dim3((unsigned)((C + 31) / 32), 2, (unsigned)groups), 256, 0, st, part, grid, 2, (int)C,
                 dgamma, dbeta, ps, accumulate);
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}

int eegx_add_dropout_fwd_bf16(const void* a, const void* b, void* out, int64_t n, float scale,
                              const uint64_t* rng_state, uint32_t site, float p, void* stream) {
    EEGX_EW_CHECK(n);
    if (n == 0) return EEGX_OK;
    EEGX_REQUIRE(a && b && out, EEGX_ERR_ARG, "add_dropout: NULL pointer");
    EEGX_REQUIRE(eegx::aligned16(a) && eegx::aligned16(b) && eegx::aligned16(out), EEGX_ERR_ALIGN,
                 "add_dropout: pointers must be 16-byte aligned");
    const DropoutCfg dc{reinterpret_cast<const unsigned long long*>(rng_state), site, p};
    eegx::launch(add_dropout_fwd_kernel, ew_grid(n / 8), 256, 0, static_cast<cudaStream_t>(stream), 
        static_cast<const __nv_bfloat16*>(a), static_cast<const __nv_bfloat16*>(b), static_cast<__nv_bfloat16*>(out),
        n / 8, scale, dc);
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}

int eegx_dropout_scale_bf16(const void* in, void* out, int64_t n, float scale, const uint64_t* rng_state,
                            uint32_t site, float p, void* stream) {
    EEGX_EW_CHECK(n);
    if (n == 0) return EEGX_OK;
    EEGX_REQUIRE(in && out, EEGX_ERR_ARG, "dropout_scale: NULL pointer");
    EEGX_REQUIRE(eegx::aligned16(in) && eegx::aligned16(out), EEGX_ERR_ALIGN,
                 "dropout_scale: pointers must be 16-byte aligned");
    const DropoutCfg dc{reinterpret_cast<const unsigned long long*>(rng_state), site, p};
    eegx::launch(dropout_scale_kernel, ew_grid(n / 8), 256, 0, static_cast<cudaStream_t>(stream), 
        static_cast<const __nv_bfloat16*>(in), static_cast<__nv_bfloat16*>(out), n / 8, scale, dc);
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}

int eegx_gelu_dropout_fwd_bf16(const void* x, void* out, int64_t n, const uint64_t* rng_state, uint32_t site,
                               float p, void* stream) {
    EEGX_EW_CHECK(n);
    if (n == 0) return EEGX_OK;
    EEGX_REQUIRE(x && out, EEGX_ERR_ARG, "gelu_dropout: NULL pointer");
    EEGX_REQUIRE(eegx::aligned16(x) && eegx::aligned16(out), EEGX_ERR_ALIGN,
                 "gelu_dropout: pointers must be 16-byte aligned");
    const DropoutCfg dc{reinterpret_cast<const unsigned long long*>(rng_state), site, p};
    eegx::launch(gelu_dropout_fwd_kernel, ew_grid(n / 8), 256, 0, static_cast<cudaStream_t>(stream), 
        static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(out), n / 8, dc);
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}

int eegx_gelu_dropout_bwd_bf16(const void* dout, const void* x, void* dx, int64_t n, const uint64_t* rng_state,
                               uint32_t site, float p, void* stream) {
    EEGX_EW_CHECK(n);
    if (n == 0) return EEGX_OK;
    EEGX_REQUIRE(dout && x && dx, EEGX_ERR_ARG, "gelu_dropout bwd: NULL pointer");
    EEGX_REQUIRE(eegx::aligned16(dout) && eegx::aligned16(x) && eegx::aligned16(dx), EEGX_ERR_ALIGN,
                 "gelu_dropout bwd: pointers must be 16-byte aligned");
    const DropoutCfg dc{reinterpret_cast<const unsigned long long*>(rng_state), site, p};
    eegx::launch(gelu_dropout_bwd_kernel, ew_grid(n / 8), 256, 0, static_cast<cudaStream_t>(stream), 
        static_cast<const __nv_bfloat16*>(dout), static_cast<const __nv_bfloat16*>(x),
        static_cast<__nv_bfloat16*>(dx), n / 8, dc);
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}

int eegx_glu_fwd_bf16(const void* ag, void* out, int64_t rows, int64_t H, const uint64_t* rng_state,
                      uint32_t site, float p, void* stream) {
    if (int rc = eegx::require_sm100()) return rc;
    EEGX_REQUIRE(rows >= 0 && H >= 8 && (H % 8) == 0, EEGX_ERR_SHAPE, "glu: H must be a multiple of 8");
    if (rows == 0) return EEGX_OK;
    EEGX_REQUIRE(ag && out, EEGX_ERR_ARG, "glu: NULL pointer");
    EEGX_REQUIRE(eegx::aligned16(ag) && eegx::aligned16(out), EEGX_ERR_ALIGN, "glu: pointers must be 16-byte aligned");
    const DropoutCfg dc{reinterpret_cast<const unsigned long long*>(rng_state), site, p};
    eegx::launch(glu_fwd_kernel, ew_grid(rows * (H / 8)), 256, 0, static_cast<cudaStream_t>(stream), 
        static_cast<const __nv_bfloat16*>(ag), static_cast<__nv_bfloat16*>(out), rows, (int)H, dc);
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}

int eegx_glu_bwd_bf16(const void* dout, const void* ag, void* dag, int64_t rows, int64_t H,
                      const uint64_t* rng_state, uint32_t site, float p, void* stream) {
    if (int rc = eegx::require_sm100()) return rc;
    EEGX_REQUIRE(rows >= 0 && H >= 8 && (H % 8) == 0, EEGX_ERR_SHAPE, "glu: H must be a multiple of 8");
    if (rows == 0) return EEGX_OK;
    EEGX_REQUIRE(dout && ag && dag, EEGX_ERR_ARG, "glu bwd: NULL pointer");
    EEGX_REQUIRE(eegx::aligned16(dout) && eegx::aligned16(ag) && eegx::aligned16(dag), EEGX_ERR_ALIGN,
                 "glu bwd: pointers must be 16-byte aligned");
    const DropoutCfg dc{reinterpret_cast<const unsigned long long*>(rng_state), site, p};
    eegx::launch(glu_bwd_kernel, ew_grid(rows * (H / 8)), 256, 0, static_cast<cudaStream_t>(stream), 
        static_cast<const __nv_bfloat16*>(dout), static_cast<const __nv_bfloat16*>(ag),
        static_cast<__nv_bfloat16*>(dag), rows, (int)H, dc);
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}

int eegx_assemble_tokens_fwd_bf16(const void* h, const float* cls, int64_t cls_gstride, const float* temporal,
                                  int64_t temporal_gstride, const float* pos, int64_t pos_gstride, void* out, int64_t G,
                                  int64_t B, int64_t T, int64_t d, void* stream) {
    if (int rc = eegx::require_sm100()) return rc;
    EEGX_REQUIRE(G >= 1 && B >= 0 && T >= 1 && d >= 8 && (d % 8) == 0, EEGX_ERR_SHAPE,
                 "assemble_tokens: need G >= 1, T >= 1, d a multiple of 8");
    if (B == 0) return EEGX_OK;
    EEGX_REQUIRE(h && cls && temporal && pos && out, EEGX_ERR_ARG, "assemble_tokens: NULL pointer");
    EEGX_REQUIRE(eegx::aligned16(h) && eegx::aligned16(out) && eegx::aligned16(cls) && eegx::aligned16(temporal) &&
                     eegx::aligned16(pos) && (cls_gstride % 4) == 0 && (temporal_gstride % 4) == 0 && (pos_gstride % 4) == 0,
                 EEGX_ERR_ALIGN, "assemble_tokens: pointers and group strides must be 16-byte aligned");
    eegx::launch(assemble_tokens_fwd_kernel, ew_grid(G * B * (T + N_TOK) * (d / 8)), 256, 0, static_cast<cudaStream_t>(stream),
                 static_cast<const __nv_bfloat16*>(h), cls, (long long)cls_gstride, temporal, (long long)temporal_gstride, pos,
                 (long long)pos_gstride, static_cast<__nv_bfloat16*>(out), (long long)(G * B), (int)B, (int)T, (int)d);
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}

int eegx_assemble_tokens_bwd_bf16(const void* dout, void* dh, int64_t GB, int64_t T, int64_t d, void* stream) {
    if (int rc = eegx::require_sm100()) return rc;
    EEGX_REQUIRE(GB >= 0 && T >= 1 && d >= 8 && (d % 8) == 0, EEGX_ERR_SHAPE, "assemble_tokens bwd: bad sizes");
    if (GB == 0) return EEGX_OK;
    EEGX_REQUIRE(dout && dh, EEGX_ERR_ARG, "assemble_tokens bwd: NULL pointer");
    EEGX_REQUIRE(eegx::aligned16(dout) && eegx::aligned16(dh), EEGX_ERR_ALIGN, "assemble_tokens bwd: 16-byte alignment");
    eegx::launch(assemble_tokens_bwd_kernel, ew_grid(GB * T * (d / 8)), 256, 0, static_cast<cudaStream_t>(stream),
                 static_cast<const __nv_bfloat16*>(dout), static_cast<__nv_bfloat16*>(dh), (long long)GB, (int)T, (int)d);
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}

}  // extern "C"
