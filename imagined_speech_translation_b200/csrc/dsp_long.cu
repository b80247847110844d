// Tuned fused DSP kernel for BASELINE config 4 (high-density, long window):
//   T = 4096, 65-tap FIR, n_fft = 1024, hop = 256.
//
//   x (rows, 4096) f32  ->  out (rows, 513, 17) f32,   rows = B * C
//
// Spec: SURVEY.md section 8(c) (the reference has no DSP code; dsp_generic.cu is the any-shape
// implementation of the same spec and the two are cross-checked on the GPU).
//
// One row per tile, 6 warps, two CTAs per SM, persistent over the rows:
//   load   one 16 KB cp.async.bulk per row (TMA, mbarrier), issued one tile ahead.
//   FIR    register-tiled direct form, 8 outputs per thread, taps as constant-bank operands,
//          written into a row that carries its own 512-sample reflections (no boundary logic later).
//   STFT   ONE WARP PER FRAME.  Real FFT-1024 = complex FFT-512 on z[n] = y[2n] + i y[2n+1], done as three
//          radix-8 passes (512 = 8 x 8 x 8): every lane runs two 8-point FFTs per pass in registers; the two
//          transposes between the passes go through a 4.5 KB per-warp patch of shared memory laid out so that
//          all of its accesses are bank-conflict free (pitch 72 for the first, XOR-swizzled 16-byte chunks for
//          the second).  The conjugate pairing Z[k] <-> Z[512-k] of the split step lives in lanes l and 32-l:
//          16 warp shuffles.  Window and second-pass twiddles come from shared tables laid out [j][lane]; the
//          first-pass and split-step twiddles stay in registers.  Power, log (MUFU lg2) and the z-score
//          partial sums are taken in registers.
//   stats  544 per-lane partials per row are combined in a fixed order in fp64 (bit-stable, no atomics).
//   store  coalesced 128-bit streaming stores.
//
// Algorithmic HBM bytes per row: 16,384 read + 34,884 written (6,562,304 B per 128-channel trial).
#include <math.h>

#include "dsp_device.cuh"
#include "dsp_plan.h"

namespace {

using namespace eegx_dsp;

constexpr int T = 4096;
constexpr int NF = 17;
constexpr int F = 513;
constexpr int ROW_OUT = F * NF;            // 8721
constexpr int NT = 192;
constexpr int NWARPS = NT / 32;
constexpr int REFL = 512;
constexpr int XS_FLOATS = 32 + T + 32 + 4;           // 16,656 B
constexpr int YS_FLOATS = REFL + T + REFL + 4;        // 20,496 B
constexpr int LS_FLOATS = 8728;                       // >= 8721 + 3, multiple of 4
constexpr int SCR_FLOATS = 2 * 8 * 72;                // per warp: 576 float2
constexpr int LANE_TABLE = 44;                        // per lane: twA[2][7] complex + split[2][4] complex
constexpr int SHARED_TABLE = 2 * 16 * 32 * 2;         // window[16][32] float2 | twB[16][32] float2

constexpr int OFF_XS = 0;
constexpr int OFF_YS = OFF_XS + XS_FLOATS;
constexpr int OFF_LS = OFF_YS + YS_FLOATS;
constexpr int OFF_SCR = OFF_LS + LS_FLOATS;
constexpr int OFF_TAB = OFF_SCR + NWARPS * SCR_FLOATS;
constexpr int OFF_STAT = OFF_TAB + SHARED_TABLE;
constexpr int OFF_BAR = OFF_STAT + NF * 32 * 2;
constexpr int SMEM_FLOATS = OFF_BAR + 2;
constexpr size_t SMEM_BYTES = SMEM_FLOATS * sizeof(float);
constexpr int CTAS_PER_SM = 2;
static_assert((OFF_YS % 4) == 0 && (OFF_LS % 4) == 0 && (OFF_SCR % 4) == 0 && (OFF_TAB % 4) == 0 &&
              (OFF_STAT % 2) == 0 && (OFF_BAR % 2) == 0, "alignment of the shared-memory regions");
static_assert(CTAS_PER_SM * (SMEM_BYTES + 1024) <= 227 * 1024, "two tiles per SM must fit");

struct LongArgs {
    const float* x;
    float* out;
    long long rows;
    const float* tables;       // [32][LANE_TABLE] per-lane constants, then SHARED_TABLE floats
    float log_eps4;            // 4 * log_eps (the FFT is kept scaled by 2)
    float z_eps;
    float taps_rev[65];        // taps_rev[d] = h[64 - d]
};

__global__ void __launch_bounds__(NT, CTAS_PER_SM) dsp_long_kernel(const __grid_constant__ LongArgs a) {
    extern __shared__ __align__(128) float smem[];
    float* xs = smem + OFF_XS;
    float* ys = smem + OFF_YS;
    float* Ls = smem + OFF_LS;
    float2* win_s = reinterpret_cast<float2*>(smem + OFF_TAB);
    float2* twb_s = win_s + 16 * 32;
    float2* stat = reinterpret_cast<float2*>(smem + OFF_STAT);
    const unsigned bar = smem_u32(smem + OFF_BAR);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float2* scr = reinterpret_cast<float2*>(smem + OFF_SCR + warp * SCR_FLOATS);

    // per-lane constants (fixed for the lifetime of the CTA)
    float twar[2][7], twai[2][7], spr[2][4], spi[2][4];
    {
        const float* tb = a.tables + lane * LANE_TABLE;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                twar[c][k] = __ldg(tb + (c * 7 + k) * 2);
                twai[c][k] = __ldg(tb + (c * 7 + k) * 2 + 1);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                spr[c][k] = __ldg(tb + 28 + (c * 4 + k) * 2);
                spi[c][k] = __ldg(tb + 28 + (c * 4 + k) * 2 + 1);
            }
        }
    }
    for (int i = tid; i < SHARED_TABLE; i += NT) smem[OFF_TAB + i] = __ldg(a.tables + 32 * LANE_TABLE + i);
    // FIR zero halos (32 samples each side of the row), written once
    if (tid < 64) xs[tid < 32 ? tid : T + tid] = 0.0f;
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    long long row = blockIdx.x;
    if (row < a.rows && tid == 0) {
        mbar_expect_tx(bar, (unsigned)(T * sizeof(float)));
        tma_load_1d(smem_u32(xs + 32), a.x + row * (long long)T, (unsigned)(T * sizeof(float)), bar);
    }
    unsigned phase = 0;

    for (; row < a.rows; row += gridDim.x) {
        mbar_wait(bar, phase);
        phase ^= 1;

        // ------------------------------ FIR ------------------------------
        // a thread owns outputs t = 8j .. 8j+7
        for (int j = tid; j < T / 8; j += NT) {
            const float4* src = reinterpret_cast<const float4*>(xs) + 2 * j;
            float acc[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = 0.0f;
#pragma unroll
            for (int sl = 0; sl < 18; ++sl) {
                const float4 v = src[sl];
                const float in[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = 4 * sl + u;   // input i feeds output e with tap d = i - e
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int d = i - e;
                        if (d >= 0 && d <= 64) acc[e] = fmaf(a.taps_rev[d], in[u], acc[e]);
                    }
                }
            }
            float4* dsty = reinterpret_cast<float4*>(ys + REFL + 8 * j);
            dsty[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
            dsty[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
            if (j <= REFL / 8) {                 // reflect copy on the left: index -t for t in [1, 512]
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int t = 8 * j + e;
                    if (t >= 1 && t <= REFL) ys[REFL - t] = acc[e];
                }
            }
            if (j >= (T - REFL - 1) / 8) {       // reflect copy on the right: index 2(T-1)-t for t in [T-513, T-2]
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int t = 8 * j + e;
                    if (t >= T - REFL - 1 && t <= T - 2) ys[2 * (T - 1) - t + REFL] = acc[e];
                }
            }
        }
        __syncthreads();
        // xs is free again: fetch the next row while the STFT runs
        if (tid == 0 && row + gridDim.x < a.rows) {
            mbar_expect_tx(bar, (unsigned)(T * sizeof(float)));
            tma_load_1d(smem_u32(xs + 32), a.x + (row + gridDim.x) * (long long)T, (unsigned)(T * sizeof(float)), bar);
        }

        // ------------------------------ STFT: one warp per frame ------------------------------
        const int ph = (int)(row & 3);           // the row sits in Ls with the 16-byte phase of its global address
#pragma unroll 1
        for (int m = warp; m < NF; m += NWARPS) {
            const float2* yseg = reinterpret_cast<const float2*>(ys + 256 * m);     // z[n] = yseg[n]
            cf v[2][8];
            // pass A: FFT over n1 of z[64 n1 + c], c = lane + 32 col
#pragma unroll
            for (int c = 0; c < 2; ++c) {
#pragma unroll
                for (int n1 = 0; n1 < 8; ++n1) {
                    const float2 y2 = yseg[64 * n1 + 32 * c + lane];
                    const float2 w2 = win_s[(c * 8 + n1) * 32 + lane];
                    v[c][n1] = {y2.x * w2.x, y2.y * w2.y};
                }
                fft8(v[c]);
#pragma unroll
                for (int k = 1; k < 8; ++k) v[c][k] = cmul(v[c][k], twar[c][k - 1], twai[c][k - 1]);   // W64^(n2 k1)
            }
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int k1 = 0; k1 < 8; ++k1) scr[k1 * 72 + 32 * c + lane] = make_float2(v[c][k1].r, v[c][k1].i);
            __syncwarp();
            // pass B: FFT over n2 of A[k1; 8 n2 + n3], n3 = lane & 7, k1 = (lane >> 3) + 4 col
            const int n3 = lane & 7;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int k1 = (lane >> 3) + 4 * c;
#pragma unroll
                for (int n2 = 0; n2 < 8; ++n2) {
                    const float2 t2 = scr[k1 * 72 + 8 * n2 + n3];
                    v[c][n2] = {t2.x, t2.y};
                }
            }
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int k1 = (lane >> 3) + 4 * c;
                fft8(v[c]);
#pragma unroll
                for (int k2 = 0; k2 < 8; ++k2) {
                    const float2 w2 = twb_s[(c * 8 + k2) * 32 + lane];                                  // W512^(n3 (k1 + 8 k2))
                    const cf b = cmul(v[c][k2], w2.x, w2.y);
                    const int q = k1 + 8 * k2;
                    scr[q * 8 + ((((n3 >> 1) ^ (q >> 1)) & 3) << 1) + (n3 & 1)] = make_float2(b.r, b.i);
                }
            }
            __syncwarp();
            // pass C: FFT over n3 of B[q; n3], q = lane + 32 col  ->  Z[q + 64 k3]
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int q = lane + 32 * c;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 t4 = *reinterpret_cast<const float4*>(scr + q * 8 + (((j ^ (q >> 1)) & 3) << 1));
                    v[c][2 * j] = {t4.x, t4.y};
                    v[c][2 * j + 1] = {t4.z, t4.w};
                }
                fft8(v[c]);
            }
            // conjugate partners: Z[512 - k] of k = q + 64 k3 (k3 < 4) is column (1 - col) of lane 32 - l at index
            // 7 - k3; lane 0 pairs with itself (column 0 at index 8 - k3, column 1 at 7 - k3)
            const int src_lane = (32 - lane) & 31;
            cf R[2][4];
#pragma unroll
            for (int k3 = 0; k3 < 4; ++k3) {
                const cf a0 = lane == 0 ? v[0][(8 - k3) & 7] : v[1][7 - k3];     // for the receiver's column 0
                const cf a1 = lane == 0 ? v[1][7 - k3] : v[0][7 - k3];           // for the receiver's column 1
                R[0][k3].r = __shfl_sync(0xffffffffu, a0.r, src_lane);
                R[0][k3].i = __shfl_sync(0xffffffffu, a0.i, src_lane);
                R[1][k3].r = __shfl_sync(0xffffffffu, a1.r, src_lane);
                R[1][k3].i = __shfl_sync(0xffffffffu, a1.i, src_lane);
            }
            float* Lcol = Ls + ph + m;
            float s1 = 0.0f, s2 = 0.0f;
            constexpr float LN2 = 0.69314718055994530942f;
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int k3 = 0; k3 < 4; ++k3) {
                    const int k = lane + 32 * c + 64 * k3;
                    const cf zk = v[c][k3], zm = R[c][k3];
                    const cf E = {zk.r + zm.r, zk.i - zm.i};
                    const cf D = {zk.r - zm.r, zk.i + zm.i};
                    const cf O = {D.i, -D.r};
                    const cf Tt = cmul(O, spr[c][k3], spi[c][k3]);
                    const cf A = cadd(E, Tt), Bc = csub(E, Tt);
                    const float pa = fmaf(A.r, A.r, fmaf(A.i, A.i, a.log_eps4));
                    const float pb = fmaf(Bc.r, Bc.r, fmaf(Bc.i, Bc.i, a.log_eps4));
                    const float la = fmaf(fast_log2(pa), LN2, -2.0f * LN2);
                    const float lb = fmaf(fast_log2(pb), LN2, -2.0f * LN2);
                    Lcol[k * NF] = la;
                    Lcol[(512 - k) * NF] = lb;
                    s1 += la + lb;
                    s2 = fmaf(la, la, fmaf(lb, lb, s2));
                }
            if (lane == 0) {   // bin 256 pairs with itself: |X[256]|^2 = |Z[256]|^2
                const cf zz = v[0][4];
                const float p = fmaf(4.0f * zz.r, zz.r, fmaf(4.0f * zz.i, zz.i, a.log_eps4));
                const float l = fmaf(fast_log2(p), LN2, -2.0f * LN2);
                Lcol[256 * NF] = l;
                s1 += l;
                s2 = fmaf(l, l, s2);
            }
            stat[m * 32 + lane] = make_float2(s1, s2);
        }
        __syncthreads();

        // ---------------- row statistics + normalise + store ----------------
        // every warp reduces the 544 partials itself, in the same fixed order: float within a lane, double across
        float p1 = 0.0f, p2 = 0.0f;
#pragma unroll
        for (int i = 0; i < NF; ++i) {
            const float2 p = stat[i * 32 + lane];
            p1 += p.x;
            p2 += p.y;
        }
        double d1 = (double)p1, d2 = (double)p2;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            d1 += __shfl_xor_sync(0xffffffffu, d1, o);
            d2 += __shfl_xor_sync(0xffffffffu, d2, o);
        }
        const double mean = d1 * (1.0 / (double)ROW_OUT);
        const double vard = d2 * (1.0 / (double)ROW_OUT) - mean * mean;
        const float var = vard > 0.0 ? (float)vard : 0.0f;
        const float inv = 1.0f / (sqrtf(var) + a.z_eps);
        const float cc = -(float)mean * inv;

        const float* src = Ls + ph;
        float* dst = a.out + row * (long long)ROW_OUT;
        const int head = (4 - ph) & 3;
        const int n4 = (ROW_OUT - head) >> 2;
        const int tail = ROW_OUT - head - 4 * n4;
        const float4* s4 = reinterpret_cast<const float4*>(src + head);
        float4* d4 = reinterpret_cast<float4*>(dst + head);
        for (int i = tid; i < n4; i += NT) {
            const float4 l = s4[i];
            __stcs(d4 + i, make_float4(fmaf(l.x, inv, cc), fmaf(l.y, inv, cc), fmaf(l.z, inv, cc), fmaf(l.w, inv, cc)));
        }
        if (tid < head) __stcs(dst + tid, fmaf(src[tid], inv, cc));
        if (tid < tail) __stcs(dst + head + 4 * n4 + tid, fmaf(src[head + 4 * n4 + tid], inv, cc));
        // the post-FIR barrier of the next iteration orders these reads of Ls / stat against the next STFT's writes
    }
}

}  // namespace

namespace eegx {

bool dsp_long_supported(const eegx_dsp_plan* p) {
    return p->n_fft == 1024 && p->hop == 256 && p->numtaps == 65 && p->T == T;
}

int dsp_long_table_floats() { return 32 * LANE_TABLE + SHARED_TABLE; }

// constant tables, computed in double
void dsp_long_fill_tables(float* t) {
    const double two_pi = 6.283185307179586476925286766559;
    for (int l = 0; l < 32; ++l) {
        float* tb = t + l * LANE_TABLE;
        for (int c = 0; c < 2; ++c) {
            const int n2 = (l >> 3) + 4 * c;                 // pass A: column 8 n2 + n3 = l + 32 c
            for (int k1 = 1; k1 < 8; ++k1) {
                const double ang = -two_pi * (double)(n2 * k1) / 64.0;
                tb[(c * 7 + (k1 - 1)) * 2] = (float)cos(ang);
                tb[(c * 7 + (k1 - 1)) * 2 + 1] = (float)sin(ang);
            }
            for (int k3 = 0; k3 < 4; ++k3) {                 // split step: W1024^k, k = l + 32 c + 64 k3
                const double ang = -two_pi * (double)(l + 32 * c + 64 * k3) / 1024.0;
                tb[28 + (c * 4 + k3) * 2] = (float)cos(ang);
                tb[28 + (c * 4 + k3) * 2 + 1] = (float)sin(ang);
            }
        }
    }
    float* win = t + 32 * LANE_TABLE;                        // [(c * 8 + n1)][lane] float2
    float* twb = win + 16 * 32 * 2;                          // [(c * 8 + k2)][lane] float2
    for (int c = 0; c < 2; ++c)
        for (int j = 0; j < 8; ++j)
            for (int l = 0; l < 32; ++l) {
                const int n = 2 * (64 * j + 32 * c + l);     // window of the two samples packed in z[64 n1 + col]
                win[((c * 8 + j) * 32 + l) * 2] = (float)(0.5 - 0.5 * cos(two_pi * n / 1024.0));
                win[((c * 8 + j) * 32 + l) * 2 + 1] = (float)(0.5 - 0.5 * cos(two_pi * (n + 1) / 1024.0));
                const int n3 = l & 7, k1 = (l >> 3) + 4 * c, q = k1 + 8 * j;
                const double ang = -two_pi * (double)((n3 * q) % 512) / 512.0;
                twb[((c * 8 + j) * 32 + l) * 2] = (float)cos(ang);
                twb[((c * 8 + j) * 32 + l) * 2 + 1] = (float)sin(ang);
            }
}

int launch_dsp_long(const eegx_dsp_plan* plan, const DspArgs& d, cudaStream_t st) {
    EEGX_REQUIRE(d.onsets == nullptr, EEGX_ERR_ARG, "tuned kernel takes pre-cut trials only");
    EEGX_REQUIRE(plan->d_lane_tables != nullptr, EEGX_ERR_ARG, "plan has no tuned tables");
    LongArgs a;
    a.x = d.x;
    a.out = d.out;
    a.rows = d.rows;
    a.tables = plan->d_lane_tables;
    a.log_eps4 = 4.0f * plan->log_eps;
    a.z_eps = plan->z_eps;
    for (int i = 0; i < 65; ++i) a.taps_rev[i] = plan->h_taps[64 - i];
    EEGX_CUDA_CHECK(cudaFuncSetAttribute(dsp_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    const long long max_ctas = (long long)CTAS_PER_SM * kNumSMsB200;
    const int grid = (int)(d.rows < max_ctas ? d.rows : max_ctas);
    dsp_long_kernel<<<grid, NT, SMEM_BYTES, st>>>(a);
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}

}  // namespace eegx
