// Packed-pair variant of the tuned fused DSP kernel (BASELINE config 2: T = 2048, 65-tap FIR,
// n_fft = 256, hop = 64; spec: SURVEY.md section 8(c)).
//
//   x (rows, 2048) f32  ->  out (rows, 129, 33) f32,   rows = B * C
//
// Same pipeline as dsp_tuned.cu (TMA bulk row loads one tile ahead -> FIR -> reflect pad -> Hann
// STFT -> log power -> fixed-order fp64 z-score statistics -> coalesced streaming stores), but a
// tile is TWO rows (A, B) that move through every arithmetic stage in lockstep as packed f32x2
// values (lo = row A, hi = row B): the FIR and the FFT butterflies issue add/mul/fma.rn.f32x2
// (SASS FADD2 / FMUL2 / FFMA2, taps and twiddles as scalar-broadcast operands), which halves the
// issue slots per sample.  Measured on the B200 (tools/ubench_fma.cu): FFMA2 has the same lane
// throughput as FFMA (127 vs 122 FMA/clk/SM) -- the win is instruction issue, which is what bounds
// the scalar kernel (ncu: issue slots 73 % busy, FMA pipe 57 %).
//
// Shared-memory layouts are pair-interleaved: element p of a row pair lives at floats (2p, 2p+1),
// and float offset `o` is stored at o + 4 * (o >> 5) (four pad floats per 32) which makes the
// three access patterns -- FIR windows at a 32-float stride, FIR results, and STFT frame loads at
// an 8-float stride -- all 128-bit bank-conflict free.
#include <math.h>

#include "dsp_plan.h"

namespace {

constexpr int T = 2048;
constexpr int NF = 33;
constexpr int F = 129;
constexpr int ROW_OUT = F * NF;          // 4257
constexpr int LS_PITCH = 4264;
constexpr int LANE_TABLE = 60;
constexpr int NT = 128;                  // threads per CTA: one 16-output FIR item per thread
constexpr int NGROUPS = NT / 8;
constexpr int NWARPS = NT / 32;
constexpr int NTASKS = 9;                // warp-tasks per tile: task k < 8 = frames k + 8q (q = group in warp), task 8 = frame 32
constexpr int ROUNDS = (NTASKS + NWARPS - 1) / NWARPS;

__host__ __device__ constexpr int padded(int o) { return o + 4 * (o >> 5); }

constexpr int XI_PAIRS = 32 + T + 32;            // zero halo of 32 pairs on each side
constexpr int YI_PAIRS = 128 + T + 128;          // reflect extension of 128 pairs on each side
constexpr int XI_FLOATS = padded(2 * XI_PAIRS);  // 4752
constexpr int YI_FLOATS = padded(2 * YI_PAIRS);  // 5184
constexpr int SCR_FLOATS = NGROUPS * 512;        // per group: two 256-float planes
constexpr int OFF_XS = 0;                                        // TMA staging: 2 dense rows
constexpr int OFF_XI = OFF_XS + 2 * T;                           // aliased with the STFT scratch
constexpr int OFF_SCR = OFF_XI;
constexpr int OFF_YI = OFF_XI + (XI_FLOATS > SCR_FLOATS ? XI_FLOATS : SCR_FLOATS);
constexpr int OFF_LS = OFF_YI + YI_FLOATS;
constexpr int OFF_STAT = OFF_LS + 2 * LS_PITCH;
constexpr int OFF_BAR = OFF_STAT + 2 * NF * 8 * 2;
constexpr int SMEM_FLOATS = OFF_BAR + 2;
constexpr size_t SMEM_BYTES = SMEM_FLOATS * sizeof(float);
static_assert((OFF_XI % 4) == 0 && (OFF_YI % 4) == 0 && (OFF_LS % 4) == 0 && (OFF_STAT % 2) == 0 && (OFF_BAR % 2) == 0,
              "alignment of the shared-memory regions");
static_assert(2 * (SMEM_BYTES + 1024) <= 227 * 1024, "two CTAs per SM must fit");

struct PairArgs {
    const float* x;
    float* out;
    long long rows;
    const float* lane_tables;  // [8][LANE_TABLE]
    float log_eps4;            // 4 * log_eps (the FFT is kept scaled by 2)
    float z_eps;
    float taps_rev[65];        // taps_rev[d] = h[64 - d]
};

// ---------------------------------------------------------------- packed f32x2 arithmetic
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(f2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2 bc(float x) { return pk(x, x); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f2 neg2(f2 a) { return a ^ 0x8000000080000000ull; }

struct c2 { f2 r, i; };   // a complex value of row A (low halves) and of row B (high halves)
__device__ __forceinline__ c2 cadd(c2 a, c2 b) { return {add2(a.r, b.r), add2(a.i, b.i)}; }
__device__ __forceinline__ c2 csub(c2 a, c2 b) { return {sub2(a.r, b.r), sub2(a.i, b.i)}; }
__device__ __forceinline__ c2 cmul(c2 a, float wr, float wi) {
    const f2 r = bc(wr), i = bc(wi);
    return {fma2(a.r, r, neg2(mul2(a.i, i))), fma2(a.r, i, mul2(a.i, r))};
}
__device__ __forceinline__ c2 mul_neg_i(c2 a) { return {a.i, neg2(a.r)}; }   // a * (-i)

constexpr float RSQRT2 = 0.70710678118654752440f;

// In-place forward 8-point FFT (e^{-i...}), natural order in and out.
__device__ __forceinline__ void fft8(c2 (&v)[8]) {
    const f2 rs = bc(RSQRT2);
    c2 e0 = cadd(v[0], v[4]), e1 = cadd(v[1], v[5]), e2 = cadd(v[2], v[6]), e3 = cadd(v[3], v[7]);
    c2 d0 = csub(v[0], v[4]), d1 = csub(v[1], v[5]), d2 = csub(v[2], v[6]), d3 = csub(v[3], v[7]);
    c2 o0 = d0;
    c2 o1 = {mul2(add2(d1.r, d1.i), rs), mul2(sub2(d1.i, d1.r), rs)};      // * W8^1
    c2 o2 = mul_neg_i(d2);                                                   // * W8^2
    c2 o3 = {mul2(sub2(d3.i, d3.r), rs), neg2(mul2(add2(d3.r, d3.i), rs))};  // * W8^3
    c2 s0 = cadd(e0, e2), s1 = csub(e0, e2), s2 = cadd(e1, e3), s3 = mul_neg_i(csub(e1, e3));
    v[0] = cadd(s0, s2); v[4] = csub(s0, s2); v[2] = cadd(s1, s3); v[6] = csub(s1, s3);
    c2 t0 = cadd(o0, o2), t1 = csub(o0, o2), t2 = cadd(o1, o3), t3 = mul_neg_i(csub(o1, o3));
    v[1] = cadd(t0, t2); v[5] = csub(t0, t2); v[3] = cadd(t1, t3); v[7] = csub(t1, t3);
}

// Forward 16-point FFT, natural order in and out.
__device__ __forceinline__ void fft16(const c2 (&c)[16], c2 (&out)[16]) {
    constexpr float C1 = 0.92387953251128675613f, S1 = 0.38268343236508977173f;  // cos/sin(pi/8)
    const f2 rs = bc(RSQRT2);
    c2 e[8], o[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        e[n] = cadd(c[n], c[n + 8]);
        o[n] = csub(c[n], c[n + 8]);
    }
    o[1] = cmul(o[1], C1, -S1);
    o[2] = {mul2(add2(o[2].r, o[2].i), rs), mul2(sub2(o[2].i, o[2].r), rs)};
    o[3] = cmul(o[3], S1, -C1);
    o[4] = mul_neg_i(o[4]);
    o[5] = cmul(o[5], -S1, -C1);
    o[6] = {mul2(sub2(o[6].i, o[6].r), rs), neg2(mul2(add2(o[6].r, o[6].i), rs))};
    o[7] = cmul(o[7], -C1, -S1);
    fft8(e);
    fft8(o);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        out[2 * k] = e[k];
        out[2 * k + 1] = o[k];
    }
}

__device__ __forceinline__ float fast_log2(float x) {   // x >= 4*log_eps > 0: no denormal path
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---- TMA (1-D bulk copy) + mbarrier helpers ----
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
        "WAIT_LOOP2:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra.uni WAIT_DONE2;\n\t"
        "bra.uni WAIT_LOOP2;\n\t"
        "WAIT_DONE2:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

__device__ __forceinline__ void issue_tile_loads(const PairArgs& a, float* xs, unsigned bar, long long row0) {
    const int nrows = (a.rows - row0) < 2 ? (int)(a.rows - row0) : 2;
    mbar_expect_tx(bar, (unsigned)(nrows * T * sizeof(float)));
    for (int r = 0; r < nrows; ++r)
        tma_load_1d(smem_u32(xs + r * T), a.x + (row0 + r) * (long long)T, (unsigned)(T * sizeof(float)), bar);
}

__global__ void __launch_bounds__(NT, 2)
dsp_pair_kernel(const __grid_constant__ PairArgs a) {
    extern __shared__ __align__(128) float smem[];
    float* xs = smem + OFF_XS;
    float* xi = smem + OFF_XI;
    float* scr = smem + OFF_SCR;
    float* yi = smem + OFF_YI;
    float* Ls = smem + OFF_LS;
    float2* stat = reinterpret_cast<float2*>(smem + OFF_STAT);
    const unsigned bar = smem_u32(smem + OFF_BAR);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = tid & 7, q = (tid >> 3) & 3;

    // per-lane constants (fixed for the lifetime of the CTA)
    float win[16], twr[2][7], twi[2][7], spr[8], spi[8];
    {
        const float* tb = a.lane_tables + g * LANE_TABLE;
#pragma unroll
        for (int i = 0; i < 16; ++i) win[i] = __ldg(tb + i);
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                twr[e][k] = __ldg(tb + 16 + (e * 7 + k) * 2);
                twi[e][k] = __ldg(tb + 16 + (e * 7 + k) * 2 + 1);
            }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            spr[k] = __ldg(tb + 44 + 2 * k);
            spi[k] = __ldg(tb + 44 + 2 * k + 1);
        }
    }

    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const long long ntiles = (a.rows + 1) / 2;
    long long tile = blockIdx.x;
    if (tile < ntiles && tid == 0) issue_tile_loads(a, xs, bar, tile * 2);
    unsigned phase = 0;

    for (; tile < ntiles; tile += gridDim.x) {
        const long long row0 = tile * 2;
        const bool has_b = row0 + 1 < a.rows;
        mbar_wait(bar, phase);
        phase ^= 1;

        // ---------------- interleave rows A / B into pairs (+ zero halos) ----------------
        // (the previous tile's store phase is done with the scratch region: see the barrier below)
        for (int u = tid; u < T / 4; u += NT) {
            const float4 va = *reinterpret_cast<const float4*>(xs + 4 * u);
            float4 vb = make_float4(0.f, 0.f, 0.f, 0.f);
            if (has_b) vb = *reinterpret_cast<const float4*>(xs + T + 4 * u);
            const int o = 64 + 8 * u;                       // float offset of pair 32 + 4u
            float* d = xi + padded(o);
            *reinterpret_cast<float4*>(d) = make_float4(va.x, vb.x, va.y, vb.y);
            *reinterpret_cast<float4*>(d + 4) = make_float4(va.z, vb.z, va.w, vb.w);
        }
        if (tid < 32) {                                     // 2 x 32 halo pairs = 2 x 16 chunks of 4 floats
            const int o = tid < 16 ? 4 * tid : 2 * (32 + T) + 4 * (tid - 16);
            *reinterpret_cast<float4*>(xi + padded(o)) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncthreads();
        // xs is free again: fetch the next tile while this one is processed
        if (tid == 0 && tile + gridDim.x < ntiles) issue_tile_loads(a, xs, bar, (tile + gridDim.x) * 2);

        // ------------------------------ FIR ------------------------------
        // thread owns outputs t = 16 tid .. 16 tid + 15 of both rows; streams 80 input pairs
        {
            f2 acc[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) acc[e] = 0ull;
            const float* src = xi + 36 * tid;               // padded(32 * tid)
#pragma unroll
            for (int s = 0; s < 40; ++s) {
                const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(src + 4 * s + 4 * (s >> 3));
                const f2 in[2] = {v.x, v.y};
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int i = 2 * s + u;                // input pair i feeds output e with tap d = i - e
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const int d = i - e;
                        if (d >= 0 && d <= 64) acc[e] = fma2(in[u], bc(a.taps_rev[d]), acc[e]);
                    }
                }
            }
            float* dsty = yi + 288 + 36 * tid;              // padded(2 * (128 + 16 tid))
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                ulonglong2 o;
                o.x = acc[2 * s];
                o.y = acc[2 * s + 1];
                *reinterpret_cast<ulonglong2*>(dsty + 4 * s) = o;
            }
            if (tid <= 8) {           // reflect copy on the left: index -t for t in [1, 128]
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const int t = 16 * tid + e;
                    if (t >= 1 && t <= 128) *reinterpret_cast<f2*>(yi + padded(2 * (128 - t))) = acc[e];
                }
            }
            if (tid >= 119) {         // reflect copy on the right: index 2(T-1)-t for t in [T-129, T-2]
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const int t = 16 * tid + e;
                    if (t >= T - 129 && t <= T - 2) *reinterpret_cast<f2*>(yi + padded(2 * (2 * (T - 1) - t + 128))) = acc[e];
                }
            }
        }
        __syncthreads();

        // ------------------------------ STFT ------------------------------
        float* myscr = scr + (tid >> 3) * 512;
        const int lds_lane = 8 * g + 4 * (g >> 2);           // lane part of padded(128 m + 64 aa + 8 g)
#pragma unroll 1
        for (int round = 0; round < ROUNDS; ++round) {
            // the four groups of a warp take frames 8 apart, which keeps the scattered log-power
            // stores (bank = g + frame) conflict free; the warp-uniform break skips empty tasks
            const int task = warp + NWARPS * round;
            if (task >= NTASKS) break;
            const int m = task < 8 ? task + 8 * q : 32;
            const bool valid = task < 8 || q == 0;
            const int mm = m;
            const float* yseg = yi + 144 * mm + lds_lane;

            c2 z0[8], z1[8];
#pragma unroll
            for (int aa = 0; aa < 8; ++aa) {
                const ulonglong2 u01 = *reinterpret_cast<const ulonglong2*>(yseg + 72 * aa);
                const ulonglong2 u23 = *reinterpret_cast<const ulonglong2*>(yseg + 72 * aa + 4);
                f2 v0, v1, v2, v3;
                if (aa < 4) {
                    v0 = mul2(u01.x, bc(win[aa * 4 + 0])); v1 = mul2(u01.y, bc(win[aa * 4 + 1]));
                    v2 = mul2(u23.x, bc(win[aa * 4 + 2])); v3 = mul2(u23.y, bc(win[aa * 4 + 3]));
                } else {   // hann[n + 128] = 1 - hann[n]
                    v0 = fma2(u01.x, bc(-win[(aa - 4) * 4 + 0]), u01.x); v1 = fma2(u01.y, bc(-win[(aa - 4) * 4 + 1]), u01.y);
                    v2 = fma2(u23.x, bc(-win[(aa - 4) * 4 + 2]), u23.x); v3 = fma2(u23.y, bc(-win[(aa - 4) * 4 + 3]), u23.y);
                }
                z0[aa] = {v0, v1};
                z1[aa] = {v2, v3};
            }
            fft8(z0);
            fft8(z1);
#pragma unroll
            for (int k = 1; k < 8; ++k) {
                z0[k] = cmul(z0[k], twr[0][k - 1], twi[0][k - 1]);
                z1[k] = cmul(z1[k], twr[1][k - 1], twi[1][k - 1]);
            }
            // 8 x 16 transpose through the group's two swizzled planes (plane 0: z0, plane 1: z1)
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                ulonglong2 w0, w1;
                w0.x = z0[k].r; w0.y = z0[k].i;
                w1.x = z1[k].r; w1.y = z1[k].i;
                *reinterpret_cast<ulonglong2*>(myscr + k * 32 + ((g ^ k) << 2)) = w0;
                *reinterpret_cast<ulonglong2*>(myscr + 256 + k * 32 + ((g ^ k) << 2)) = w1;
            }
            __syncwarp();
            c2 bb[16], Z[16];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const ulonglong2 w0 = *reinterpret_cast<const ulonglong2*>(myscr + g * 32 + ((jj ^ g) << 2));
                const ulonglong2 w1 = *reinterpret_cast<const ulonglong2*>(myscr + 256 + g * 32 + ((jj ^ g) << 2));
                bb[2 * jj] = {w0.x, w0.y};
                bb[2 * jj + 1] = {w1.x, w1.y};
            }
            fft16(bb, Z);   // Z[k2] = Zc[g + 8 k2]

            // conjugate partner: lane (8 - g) & 7 of the same group, index 15 - k2
            // (lane 0 pairs with itself at 16 - k2, so as a source it sends a rotated copy)
            const int src_lane = (lane & 24) | ((8 - g) & 7);
            c2 R[8];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const c2 own = Z[8 + jj];
                const c2 rot = Z[(9 + jj) & 15];
                const f2 sr = g == 0 ? rot.r : own.r;
                const f2 si = g == 0 ? rot.i : own.i;
                R[jj].r = __shfl_sync(0xffffffffu, sr, src_lane);
                R[jj].i = __shfl_sync(0xffffffffu, si, src_lane);
            }
            float* LrowA = Ls + (int)(row0 & 3) + mm;
            float* LrowB = Ls + LS_PITCH + (int)((row0 + 1) & 3) + mm;
            f2 s1 = 0ull, s2 = 0ull;
            constexpr float LN2 = 0.69314718055994530942f;
            const f2 eps4 = bc(a.log_eps4), ln2 = bc(LN2), off = bc(-2.0f * LN2);
            if (valid) {
#pragma unroll
                for (int k2 = 0; k2 < 8; ++k2) {
                    const c2 zk = Z[k2], zm = R[7 - k2];                     // R[j - 8] holds index j
                    const c2 E = {add2(zk.r, zm.r), sub2(zk.i, zm.i)};
                    const c2 D = {sub2(zk.r, zm.r), add2(zk.i, zm.i)};
                    const c2 O = {D.i, neg2(D.r)};
                    const c2 Tt = cmul(O, spr[k2], spi[k2]);
                    const c2 A = cadd(E, Tt), Bc = csub(E, Tt);
                    const f2 pa = fma2(A.r, A.r, fma2(A.i, A.i, eps4));
                    const f2 pb = fma2(Bc.r, Bc.r, fma2(Bc.i, Bc.i, eps4));
                    float paA, paB, pbA, pbB;
                    upk(pa, paA, paB);
                    upk(pb, pbA, pbB);
                    const f2 la = fma2(pk(fast_log2(paA), fast_log2(paB)), ln2, off);
                    const f2 lb = fma2(pk(fast_log2(pbA), fast_log2(pbB)), ln2, off);
                    float laA, laB, lbA, lbB;
                    upk(la, laA, laB);
                    upk(lb, lbA, lbB);
                    LrowA[(g + 8 * k2) * NF] = laA;
                    LrowB[(g + 8 * k2) * NF] = laB;
                    LrowA[(128 - g - 8 * k2) * NF] = lbA;
                    LrowB[(128 - g - 8 * k2) * NF] = lbB;
                    s1 = add2(s1, add2(la, lb));
                    s2 = fma2(la, la, fma2(lb, lb, s2));
                }
                if (g == 0) {   // bin 64 pairs with itself: |X[64]|^2 = |Zc[64]|^2
                    const c2 zz = Z[8];
                    const f2 four = bc(4.0f);
                    const f2 p = fma2(mul2(zz.r, four), zz.r, fma2(mul2(zz.i, four), zz.i, eps4));
                    float pA, pB;
                    upk(p, pA, pB);
                    const f2 l = fma2(pk(fast_log2(pA), fast_log2(pB)), ln2, off);
                    float lA, lB;
                    upk(l, lA, lB);
                    LrowA[64 * NF] = lA;
                    LrowB[64 * NF] = lB;
                    s1 = add2(s1, l);
                    s2 = fma2(l, l, s2);
                }
                float s1A, s1B, s2A, s2B;
                upk(s1, s1A, s1B);
                upk(s2, s2A, s2B);
                stat[(0 * NF + m) * 8 + g] = make_float2(s1A, s2A);
                stat[(1 * NF + m) * 8 + g] = make_float2(s1B, s2B);
            }
        }
        __syncthreads();

        // ---------------- row statistics + normalise + store ----------------
        const int nrows = has_b ? 2 : 1;
        for (int r = 0; r < nrows; ++r) {
            float p1 = 0.0f, p2 = 0.0f;
            for (int i = lane; i < NF * 8; i += 32) {
                const float2 p = stat[r * NF * 8 + i];
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
            const float c = -(float)mean * inv;

            const int ph = (int)((row0 + r) & 3);            // float phase of the global row start
            const float* src = Ls + r * LS_PITCH + ph;
            float* dst = a.out + (row0 + r) * (long long)ROW_OUT;
            const int head = (4 - ph) & 3;
            const int n4 = (ROW_OUT - head) >> 2;
            const int tail = ROW_OUT - head - 4 * n4;
            const float4* s4 = reinterpret_cast<const float4*>(src + head);
            float4* d4 = reinterpret_cast<float4*>(dst + head);
            for (int v = tid; v < n4; v += NT) {
                const float4 l = s4[v];
                __stcs(d4 + v, make_float4(fmaf(l.x, inv, c), fmaf(l.y, inv, c), fmaf(l.z, inv, c),
                                           fmaf(l.w, inv, c)));
            }
            if (tid < head) __stcs(dst + tid, fmaf(src[tid], inv, c));
            if (tid < tail) __stcs(dst + head + 4 * n4 + tid, fmaf(src[head + 4 * n4 + tid], inv, c));
        }
        // The interleave pass of the next iteration only touches the xi / scratch region (not Ls /
        // stat), and its barrier orders these reads against the next tile's STFT writes.
    }
}

}  // namespace

namespace eegx {

int launch_dsp_pair(const eegx_dsp_plan* plan, const DspArgs& d, cudaStream_t st) {
    EEGX_REQUIRE(d.onsets == nullptr, EEGX_ERR_ARG, "tuned kernel takes pre-cut trials only");
    EEGX_REQUIRE(plan->d_lane_tables != nullptr, EEGX_ERR_ARG, "plan has no tuned tables");
    PairArgs a;
    a.x = d.x;
    a.out = d.out;
    a.rows = d.rows;
    a.lane_tables = plan->d_lane_tables;
    a.log_eps4 = 4.0f * plan->log_eps;
    a.z_eps = plan->z_eps;
    for (int i = 0; i < 65; ++i) a.taps_rev[i] = plan->h_taps[64 - i];
    EEGX_CUDA_CHECK(cudaFuncSetAttribute(dsp_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    const long long ntiles = (a.rows + 1) / 2;
    const long long max_ctas = 2LL * kNumSMsB200;
    const int grid = (int)(ntiles < max_ctas ? ntiles : max_ctas);
    dsp_pair_kernel<<<grid, NT, SMEM_BYTES, st>>>(a);
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}

}  // namespace eegx
