// Tuned fused DSP kernel for BASELINE config 2: T = 2048, 65-tap FIR, n_fft = 256, hop = 64.
//
//   x (rows, 2048) f32  ->  out (rows, 129, 33) f32,   rows = B * C
//
// Spec: SURVEY.md section 8(c) (the reference has no DSP code; dsp_generic.cu is the
// any-shape implementation of the same spec and the two are cross-checked on the GPU).
//
// Design.  Persistent CTAs; a tile is ROWS consecutive rows per
// iteration.  Default: 1 row per tile, 96 threads, four CTAs per SM (measured faster than
// 2 rows / 192 threads / two CTAs per SM: more independent CTAs hide each other's barriers):
//
//   load   TMA: one cp.async.bulk (global -> shared, 8 KB) per row, completion on an
//          mbarrier, issued one tile ahead so it overlaps the STFT phase of the current
//          tile.  The zero halos the FIR needs are written once per CTA.
//   FIR    register-tiled direct form: a thread owns 8 consecutive outputs, streams the 72
//          inputs it needs through 18 LDS.128 and issues 520 FFMAs with the taps as
//          constant-bank operands.  Lanes alternate between the two rows and the row pitch
//          is 16 bytes off a multiple of 128, which makes those loads bank-conflict free in
//          a dense layout.  Outputs go to a reflect-extended row in shared memory, so the
//          STFT needs no boundary logic.
//   STFT   8 lanes per frame.  Real FFT-256 = complex FFT-128 on z[n] = y[2n] + i*y[2n+1]
//          = (8 lanes x two radix-8 FFTs in registers) -> twiddle -> 8x16 transpose through
//          an XOR-swizzled 1 KB shared-memory patch (conflict-free 128-bit accesses)
//          -> one 16-point FFT per lane.  The conjugate pairing Z[k] <-> Z[128-k] of the
//          split step lives in lanes g and 8-g: 16 warp shuffles per lane.  Power, log
//          (MUFU lg2) and the partial sums for the z-score are taken in registers.
//   stats  264 per-lane partials per row are combined in a fixed order in fp64 by one warp
//          per row (bit-stable, no atomics).
//   store  each row leaves as coalesced 128-bit streaming stores (the row is placed in
//          shared memory with the same 16-byte phase as its global address).
//
// Algorithmic HBM bytes per row: 8192 read + 17028 written; nothing else touches HBM.
#include <math.h>

#include "dsp_plan.h"

namespace {

constexpr int T = 2048;
constexpr int NF = 33;
constexpr int F = 129;
constexpr int ROW_OUT = F * NF;          // 4257
constexpr int XS_PITCH = 32 + T + 32 + 4;   // floats; 8464 B = 66 * 128 + 16
constexpr int YS_PITCH = 128 + T + 128 + 4; // floats; 9232 B = 72 * 128 + 16
constexpr int LS_PITCH = 4264;           // >= 4257 + 3, multiple of 4
constexpr int LANE_TABLE = 60;           // floats of per-lane constants

// Shared-memory carve-up for a tile of ROWS rows processed by NT threads.
template <int ROWS, int NT>
struct Cfg {
    static constexpr int NWARPS = NT / 32;
    static constexpr int NGROUPS = NT / 8;              // groups of 8 lanes (one frame each)
    static constexpr int NTASKS = 8 * ROWS + 1;         // warp-tasks per tile: 8 per row + leftover (m = 32)
    static constexpr int ROUNDS = (NTASKS + NWARPS - 1) / NWARPS;
    static constexpr int OFF_XS = 0;
    static constexpr int OFF_YS = OFF_XS + ROWS * XS_PITCH;
    static constexpr int OFF_LS = OFF_YS + ROWS * YS_PITCH;
    static constexpr int OFF_SCR = OFF_LS + ROWS * LS_PITCH;
    static constexpr int OFF_STAT = OFF_SCR + NGROUPS * 256;
    static constexpr int OFF_BAR = OFF_STAT + ROWS * NF * 8 * 2;   // 8-byte mbarrier
    static constexpr int SMEM_FLOATS = OFF_BAR + 2;
    static constexpr size_t SMEM_BYTES = SMEM_FLOATS * sizeof(float);
    static constexpr int CTAS_PER_SM = (int)((227 * 1024) / (SMEM_BYTES + 1024)) < (65536 / (NT * 168))
                                           ? (int)((227 * 1024) / (SMEM_BYTES + 1024))
                                           : (65536 / (NT * 168));
    static_assert((OFF_YS % 4) == 0 && (OFF_LS % 4) == 0 && (OFF_SCR % 4) == 0 && (OFF_STAT % 2) == 0 &&
                  (OFF_BAR % 2) == 0, "alignment of the shared-memory regions");
    static_assert(CTAS_PER_SM >= 1, "tile does not fit");
};

struct TunedArgs {
    const float* x;
    float* out;
    long long rows;
    const float* lane_tables;  // [8][LANE_TABLE]
    float log_eps4;            // 4 * log_eps (the FFT is kept scaled by 2)
    float z_eps;
    float taps_rev[65];        // taps_rev[d] = h[64 - d]
};

struct cf { float r, i; };
__device__ __forceinline__ cf cadd(cf a, cf b) { return {a.r + b.r, a.i + b.i}; }
__device__ __forceinline__ cf csub(cf a, cf b) { return {a.r - b.r, a.i - b.i}; }
__device__ __forceinline__ cf cmul(cf a, float wr, float wi) {
    return {fmaf(a.r, wr, -a.i * wi), fmaf(a.r, wi, a.i * wr)};
}
__device__ __forceinline__ cf mul_neg_i(cf a) { return {a.i, -a.r}; }   // a * (-i)

constexpr float RSQRT2 = 0.70710678118654752440f;

// In-place forward 8-point FFT (e^{-i...}), natural order in and out.
__device__ __forceinline__ void fft8(cf (&v)[8]) {
    // radix-2 DIF split: evens from sums, odds from twiddled differences
    cf e0 = cadd(v[0], v[4]), e1 = cadd(v[1], v[5]), e2 = cadd(v[2], v[6]), e3 = cadd(v[3], v[7]);
    cf d0 = csub(v[0], v[4]), d1 = csub(v[1], v[5]), d2 = csub(v[2], v[6]), d3 = csub(v[3], v[7]);
    cf o0 = d0;
    cf o1 = {(d1.r + d1.i) * RSQRT2, (d1.i - d1.r) * RSQRT2};     // * W8^1
    cf o2 = mul_neg_i(d2);                                         // * W8^2
    cf o3 = {(d3.i - d3.r) * RSQRT2, -(d3.r + d3.i) * RSQRT2};    // * W8^3
    // 4-point FFTs
    cf s0 = cadd(e0, e2), s1 = csub(e0, e2), s2 = cadd(e1, e3), s3 = mul_neg_i(csub(e1, e3));
    v[0] = cadd(s0, s2); v[4] = csub(s0, s2); v[2] = cadd(s1, s3); v[6] = csub(s1, s3);
    cf t0 = cadd(o0, o2), t1 = csub(o0, o2), t2 = cadd(o1, o3), t3 = mul_neg_i(csub(o1, o3));
    v[1] = cadd(t0, t2); v[5] = csub(t0, t2); v[3] = cadd(t1, t3); v[7] = csub(t1, t3);
}

// Forward 16-point FFT, natural order in and out.
__device__ __forceinline__ void fft16(const cf (&c)[16], cf (&out)[16]) {
    constexpr float C1 = 0.92387953251128675613f, S1 = 0.38268343236508977173f;  // cos/sin(pi/8)
    cf e[8], o[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        e[n] = cadd(c[n], c[n + 8]);
        o[n] = csub(c[n], c[n + 8]);
    }
    // o[n] *= W16^n
    o[1] = cmul(o[1], C1, -S1);
    o[2] = {(o[2].r + o[2].i) * RSQRT2, (o[2].i - o[2].r) * RSQRT2};
    o[3] = cmul(o[3], S1, -C1);
    o[4] = mul_neg_i(o[4]);
    o[5] = cmul(o[5], -S1, -C1);
    o[6] = {(o[6].i - o[6].r) * RSQRT2, -(o[6].r + o[6].i) * RSQRT2};
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


// ---- packed f32x2 arithmetic for the half-row FIR (PK variant) ----
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk2(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(f2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 bc2(float x) { return pk2(x, x); }

// ---- packed STFT (PS variant): one instruction works on two independent 8-point FFTs (or two bins) ----
struct c2 { f2 r, i; };   // two complex values: low halves = FFT (or bin) A, high halves = FFT (or bin) B
__device__ __forceinline__ c2 cadd(c2 a, c2 b) { return {add2(a.r, b.r), add2(a.i, b.i)}; }
__device__ __forceinline__ c2 csub(c2 a, c2 b) { return {sub2(a.r, b.r), sub2(a.i, b.i)}; }
__device__ __forceinline__ c2 cadd_mni(c2 x, c2 a) { return {add2(x.r, a.i), sub2(x.i, a.r)}; }   // x + (-i) a
__device__ __forceinline__ c2 csub_mni(c2 x, c2 a) { return {sub2(x.r, a.i), add2(x.i, a.r)}; }   // x - (-i) a
__device__ __forceinline__ c2 cmul2(c2 a, f2 wr, f2 wi) {                                           // a * (wr + i wi)
    return {sub2(mul2(a.r, wr), mul2(a.i, wi)), fma2(a.r, wi, mul2(a.i, wr))};
}
__device__ __forceinline__ void fft8p(c2 (&v)[8]) {
    const f2 rs = bc2(RSQRT2), nrs = bc2(-RSQRT2);
    const c2 e0 = cadd(v[0], v[4]), e1 = cadd(v[1], v[5]), e2 = cadd(v[2], v[6]), e3 = cadd(v[3], v[7]);
    const c2 d0 = csub(v[0], v[4]), d1 = csub(v[1], v[5]), d2 = csub(v[2], v[6]), d3 = csub(v[3], v[7]);
    const c2 o1 = {mul2(add2(d1.r, d1.i), rs), mul2(sub2(d1.i, d1.r), rs)};      // d1 * W8^1
    const c2 o3 = {mul2(sub2(d3.i, d3.r), rs), mul2(add2(d3.r, d3.i), nrs)};     // d3 * W8^3
    const c2 s0 = cadd(e0, e2), s1 = csub(e0, e2), s2 = cadd(e1, e3), s3 = csub(e1, e3);
    v[0] = cadd(s0, s2); v[4] = csub(s0, s2); v[2] = cadd_mni(s1, s3); v[6] = csub_mni(s1, s3);
    const c2 t0 = cadd_mni(d0, d2), t1 = csub_mni(d0, d2), t2 = cadd(o1, o3), t3 = csub(o1, o3);
    v[1] = cadd(t0, t2); v[5] = csub(t0, t2); v[3] = cadd_mni(t1, t3); v[7] = csub_mni(t1, t3);
}
// 16-point FFT of c[0..15] returned as pairs out[k] = (X[2k], X[2k + 1]): scalar even / odd split and odd-half
// twiddles, then the two 8-point FFTs packed
__device__ __forceinline__ void fft16p(const cf (&c)[16], c2 (&out)[8]) {
    constexpr float C1 = 0.92387953251128675613f, S1 = 0.38268343236508977173f;  // cos/sin(pi/8)
    cf e[8], o[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        e[n] = cadd(c[n], c[n + 8]);
        o[n] = csub(c[n], c[n + 8]);
    }
    o[1] = cmul(o[1], C1, -S1);
    o[2] = {(o[2].r + o[2].i) * RSQRT2, (o[2].i - o[2].r) * RSQRT2};
    o[3] = cmul(o[3], S1, -C1);
    o[4] = mul_neg_i(o[4]);
    o[5] = cmul(o[5], -S1, -C1);
    o[6] = {(o[6].i - o[6].r) * RSQRT2, -(o[6].r + o[6].i) * RSQRT2};
    o[7] = cmul(o[7], -C1, -S1);
#pragma unroll
    for (int n = 0; n < 8; ++n) out[n] = {pk2(e[n].r, o[n].r), pk2(e[n].i, o[n].i)};
    fft8p(out);
}
// position of extended sample p inside its aligned group of four: (0, 1, 2, 3) -> (0, 2, 1, 3)
__device__ __forceinline__ int perm4(int p) { return (p & ~3) | ((p & 1) << 1) | ((p >> 1) & 1); }

__host__ __device__ constexpr int padded32(int o) { return o + 4 * (o >> 5); }   // 4 pad floats per 32: conflict-free 128-bit access

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
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra.uni WAIT_DONE;\n\t"
        "bra.uni WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}


template <int ROWS>
__device__ __forceinline__ void issue_tile_loads(const TunedArgs& a, float* xs, unsigned bar, long long row0) {
    // one thread: arm the barrier with the byte count, then one bulk copy per row
    const int nrows = (a.rows - row0) < ROWS ? (int)(a.rows - row0) : ROWS;
    mbar_expect_tx(bar, (unsigned)(nrows * T * sizeof(float)));
    for (int r = 0; r < nrows; ++r)
        tma_load_1d(smem_u32(xs + r * XS_PITCH + 32), a.x + (row0 + r) * (long long)T,
                    (unsigned)(T * sizeof(float)), bar);
}

// PK (ROWS == 1 only): the FIR runs on packed f32x2 values -- the row's two halves (t, t + 1024) move through
// it as (lo, hi) pairs (FFMA2 with the taps as scalar-broadcast operands): half the FIR's issue slots.  The
// pairs are interleaved into the STFT scratch region (free during the FIR) with the padded32 layout.
//
// TC (ROWS == 1 only): the FIR runs on the tensor cores as a banded Toeplitz product, fp32-accurate through the
// three-term TF32 split (x_hi h_hi + x_lo h_hi + x_hi h_lo).  One mma.sync.m16n8k8 tile is 128 consecutive
// outputs, D[m][n] = y[t0 + 8m + n]; the contraction runs over the 72 inputs xs[t0 + 8m + u], u = 0..71 (9 k-tiles),
// B[u][n] = taps_rev[u - n] (a constant band: the B fragments are split once per thread and live in registers),
// A[m][u] = xs[t0 + 8m + u] -- the k-slots of a k-tile are permuted so that a thread's two A values of a row are
// adjacent samples: 17 conflict-free LDS.64 feed all 27 MMAs of a tile.  The FFMA pipe, which bounds the scalar
// FIR, is left to the STFT phase of the CTAs sharing the SM.
__device__ __forceinline__ void split_tf32(float v, unsigned& hi, unsigned& lo) {
    hi = (__float_as_uint(v) + 0x1000u) & 0xFFFFE000u;            // nearest 10-bit mantissa
    lo = __float_as_uint(v - __uint_as_float(hi));                  // exact remainder (its low bits are dropped by the MMA)
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3,
                                         unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// PS (scalar FIR only): the STFT runs on packed f32x2 values -- the lane's two 8-point FFTs, the two 8-point
// halves of its 16-point FFT and neighbouring bins of the split step move through FADD2 / FMUL2 / FFMA2 in
// lockstep; ys then holds every aligned group of four samples as (t, t+2, t+1, t+3) so that a 128-bit load is
// already the packed (re | re, im | im) pair.
// BS: bulk store.  Every warp normalises its share of the row IN PLACE in shared memory and hands it to the TMA
// store engine (cp.async.bulk shared -> global); the warp goes on to the next tile's FIR while the copy drains,
// and waits for the copy to have READ shared memory just before the next STFT overwrites it.  Replaces the
// LDS -> FFMA -> STG.CS loop whose warps sat on the store queue (24 % of the warp time for 7 % of the instructions).
template <int ROWS, int NT, bool PK, bool TC = false, bool PS = false, int MINB = 0, bool BS = false>
__global__ void __launch_bounds__(NT, MINB ? MINB : Cfg<ROWS, NT>::CTAS_PER_SM)
dsp_tuned_kernel(const __grid_constant__ TunedArgs a) {
    static_assert(!PS || (!PK && !TC), "packed STFT: scalar FIR only");
    static_assert(!PK || (ROWS == 1 && Cfg<ROWS, NT>::NGROUPS * 256 >= padded32(2 * (T / 2 + 64)) && NT >= 64),
                  "packed FIR: one row per tile, pairs must fit the STFT scratch");
    static_assert(!TC || (ROWS == 1 && !PK), "tensor-core FIR: one row per tile");
    using C = Cfg<ROWS, NT>;
    constexpr int NWARPS = C::NWARPS;
    extern __shared__ __align__(128) float smem[];
    float* xs = smem + C::OFF_XS;
    float* ys = smem + C::OFF_YS;
    float* Ls = smem + C::OFF_LS;
    float* scr = smem + C::OFF_SCR;
    float2* stat = reinterpret_cast<float2*>(smem + C::OFF_STAT);
    const unsigned bar = smem_u32(smem + C::OFF_BAR);

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

    // tensor-core FIR: this thread's B fragments (the tap band), split into TF32 hi / lo once
    const int mg = lane >> 2, mq = lane & 3;
    unsigned bh[TC ? 9 : 1][2], bl[TC ? 9 : 1][2];
    if constexpr (TC) {
#pragma unroll
        for (int kt = 0; kt < 9; ++kt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int d = 8 * kt + 2 * mq + e - mg;             // k-slot mq (+4) holds input offset 2 mq (+1)
                const float tv = (d >= 0 && d <= 64) ? a.taps_rev[d] : 0.0f;
                split_tf32(tv, bh[kt][e], bl[kt][e]);
            }
    }

    // FIR zero halos (32 samples each side of every row), written once.
    for (int i = tid; i < ROWS * 64; i += NT) {
        const int r = i >> 6, h = i & 63;
        xs[r * XS_PITCH + (h < 32 ? h : T + h)] = 0.0f;
    }
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const long long ntiles = (a.rows + ROWS - 1) / ROWS;
    long long tile = blockIdx.x;
    if (tile < ntiles && tid == 0) issue_tile_loads<ROWS>(a, xs, bar, tile * ROWS);
    unsigned phase = 0;

    for (; tile < ntiles; tile += gridDim.x) {
        const long long row0 = tile * ROWS;
        mbar_wait(bar, phase);
        phase ^= 1;

        if constexpr (PK) {
            // ---------------- interleave: pair p = (x[p - 32], x[p - 32 + T/2]), p in [0, T/2 + 64) ----------------
            float* xi = scr;
            for (int u = tid; u < (T / 2 + 64) / 4; u += NT) {
                const float4 va = *reinterpret_cast<const float4*>(xs + 4 * u);
                const float4 vb = *reinterpret_cast<const float4*>(xs + T / 2 + 4 * u);
                float* d = xi + padded32(8 * u);
                *reinterpret_cast<float4*>(d) = make_float4(va.x, vb.x, va.y, vb.y);
                *reinterpret_cast<float4*>(d + 4) = make_float4(va.z, vb.z, va.w, vb.w);
            }
            __syncthreads();
            // xs is free already: fetch the next tile while the FIR and the STFT run
            if (tid == 0 && tile + gridDim.x < ntiles) issue_tile_loads<ROWS>(a, xs, bar, (tile + gridDim.x) * ROWS);
            // ---------------- FIR: thread j < 64 owns pair-outputs t = 16 j .. 16 j + 15 ----------------
            if (tid < T / 32) {
                f2 acc[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) acc[e] = 0ull;
                const float* src = xi + 36 * tid;                       // padded32(32 * tid)
#pragma unroll
                for (int sl = 0; sl < 40; ++sl) {
                    const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(src + 4 * sl + 4 * (sl >> 3));
                    const f2 in[2] = {v.x, v.y};
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int i = 2 * sl + u;                       // input pair i feeds output e with tap d = i - e
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            const int d = i - e;
                            if (d >= 0 && d <= 64) acc[e] = fma2(in[u], pk2(a.taps_rev[d], a.taps_rev[d]), acc[e]);
                        }
                    }
                }
                float ya[16], yb[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) upk2(acc[e], ya[e], yb[e]);
                float* yrow = ys;
                float4* da = reinterpret_cast<float4*>(yrow + 128 + 16 * tid);
                float4* db = reinterpret_cast<float4*>(yrow + 128 + T / 2 + 16 * tid);
#pragma unroll
                for (int v4 = 0; v4 < 4; ++v4) {
                    da[v4] = make_float4(ya[4 * v4], ya[4 * v4 + 1], ya[4 * v4 + 2], ya[4 * v4 + 3]);
                    db[v4] = make_float4(yb[4 * v4], yb[4 * v4 + 1], yb[4 * v4 + 2], yb[4 * v4 + 3]);
                }
                if (tid <= 8) {           // reflect copy on the left: index -t for t in [1, 128]
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const int t = 16 * tid + e;
                        if (t >= 1 && t <= 128) yrow[128 - t] = ya[e];
                    }
                }
                if (tid >= 55) {          // reflect copy on the right: index 2(T-1)-t for t in [T-129, T-2]
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const int t = T / 2 + 16 * tid + e;
                        if (t >= T - 129 && t <= T - 2) yrow[2 * (T - 1) - t + 128] = yb[e];
                    }
                }
            }
        } else if constexpr (TC) {
        // ------------------------------ FIR on the tensor cores ------------------------------
        for (int blk = warp; blk < T / 128; blk += NWARPS) {
            const int t0 = 128 * blk;
            const float* src = xs + t0 + 8 * mg + 2 * mq;
            unsigned ah[17][2], al[17][2];
#pragma unroll
            for (int j = 0; j < 17; ++j) {
                const float2 v = *reinterpret_cast<const float2*>(src + 8 * j);
                split_tf32(v.x, ah[j][0], al[j][0]);
                split_tf32(v.y, ah[j][1], al[j][1]);
            }
            // six independent accumulator chains (3 terms x even / odd k-tiles): a single chain of 27 dependent
            // MMAs is bound by the MMA latency, not by the tensor pipe
            float c[6][4];
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
                for (int e = 0; e < 4; ++e) c[i][e] = 0.0f;
#pragma unroll
            for (int kt = 0; kt < 9; ++kt) {
                const int par = kt & 1;
                mma_tf32(c[0 + par], al[kt][0], al[kt + 8][0], al[kt][1], al[kt + 8][1], bh[kt][0], bh[kt][1]);
                mma_tf32(c[2 + par], ah[kt][0], ah[kt + 8][0], ah[kt][1], ah[kt + 8][1], bl[kt][0], bl[kt][1]);
                mma_tf32(c[4 + par], ah[kt][0], ah[kt + 8][0], ah[kt][1], ah[kt + 8][1], bh[kt][0], bh[kt][1]);
            }
            float acc[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)                            // small terms first
                acc[e] = ((c[0][e] + c[1][e]) + (c[2][e] + c[3][e])) + (c[4][e] + c[5][e]);
            float* yrow = ys;
#pragma unroll
            for (int h = 0; h < 2; ++h) {                           // rows mg and mg + 8 of the tile
                const int t = t0 + 8 * (mg + 8 * h) + 2 * mq;
                const float y0 = acc[2 * h], y1 = acc[2 * h + 1];
                *reinterpret_cast<float2*>(yrow + 128 + t) = make_float2(y0, y1);
                if (t <= 128) {                                     // reflect copy on the left: index -t for t in [1, 128]
                    if (t >= 1) yrow[128 - t] = y0;
                    if (t + 1 <= 128) yrow[128 - (t + 1)] = y1;
                }
                if (t + 1 >= T - 129) {                             // reflect copy on the right, t in [T-129, T-2]
                    if (t >= T - 129 && t <= T - 2) yrow[2 * (T - 1) - t + 128] = y0;
                    if (t + 1 <= T - 2) yrow[2 * (T - 1) - (t + 1) + 128] = y1;
                }
            }
        }
        } else {
        // ------------------------------ FIR ------------------------------
        // lanes alternate rows; a thread owns outputs t = 8j .. 8j+7 of its row
        for (int th = tid; th < ROWS * 256; th += NT) {
            const int r = ROWS == 1 ? 0 : (th % ROWS), j = ROWS == 1 ? th : (th / ROWS);
            const float4* src = reinterpret_cast<const float4*>(xs + r * XS_PITCH) + 2 * j;
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
            float* yrow = ys + r * YS_PITCH;
            float4* dsty = reinterpret_cast<float4*>(yrow + 128 + 8 * j);
            if constexpr (PS) {
                dsty[0] = make_float4(acc[0], acc[2], acc[1], acc[3]);
                dsty[1] = make_float4(acc[4], acc[6], acc[5], acc[7]);
            } else {
                dsty[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
                dsty[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
            }
            if (j <= 16) {          // reflect copy on the left: index -t for t in [1, 128]
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int t = 8 * j + e;
                    if (t >= 1 && t <= 128) yrow[PS ? perm4(128 - t) : 128 - t] = acc[e];
                }
            }
            if (j >= 239) {         // reflect copy on the right: index 2(T-1)-t for t in [T-129, T-2]
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int t = 8 * j + e;
                    if (t >= T - 129 && t <= T - 2) yrow[PS ? perm4(2 * (T - 1) - t + 128) : 2 * (T - 1) - t + 128] = acc[e];
                }
            }
        }
        }
        if constexpr (BS) {
            // the previous tile's bulk stores must have read Ls before this tile's STFT writes it
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncthreads();
        // xs is free again: fetch the next tile while the STFT runs
        if (!PK && tid == 0 && tile + gridDim.x < ntiles) issue_tile_loads<ROWS>(a, xs, bar, (tile + gridDim.x) * ROWS);

        // ------------------------------ STFT ------------------------------
        // rows sit in Ls with the same 16-byte phase as their global address
        float* myscr = scr + (tid >> 3) * 256;
#pragma unroll 1
        for (int round = 0; round < C::ROUNDS; ++round) {
            const int task = warp + NWARPS * round;
            if (task >= C::NTASKS) break;                    // warp-uniform
            // tasks 0..8*ROWS-1: row = task / 8, frames (task % 8) + 8 q;  last task: frame 32 of row q
            const bool full = task < 8 * ROWS;
            const bool valid = full || q < ROWS;
            const int r = full ? (task >> 3) : (q % ROWS);
            const int m = full ? (task & 7) + 8 * q : 32;
            const float* yseg = ys + r * YS_PITCH + m * 64;  // extended position 64 m

            if constexpr (PS) {
            // packed constants of this lane (built from the scalar tables; the scalar copies die)
            f2 winr[4], wini[4], ptwr[7], ptwi[7], pspr[4], pspi[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                winr[i] = pk2(win[4 * i + 0], win[4 * i + 2]);
                wini[i] = pk2(win[4 * i + 1], win[4 * i + 3]);
                pspr[i] = pk2(spr[2 * i], spr[2 * i + 1]);
                pspi[i] = pk2(spi[2 * i], spi[2 * i + 1]);
            }
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                ptwr[k] = pk2(twr[0][k], twr[1][k]);
                ptwi[k] = pk2(twi[0][k], twi[1][k]);
            }
            c2 P[8];
#pragma unroll
            for (int aa = 0; aa < 8; ++aa) {
                const ulonglong2 u = *reinterpret_cast<const ulonglong2*>(yseg + aa * 32 + 4 * g);
                if (aa < 4) {
                    P[aa] = {mul2(u.x, winr[aa]), mul2(u.y, wini[aa])};
                } else {   // hann[n + 128] = 1 - hann[n]
                    P[aa] = {sub2(u.x, mul2(u.x, winr[aa - 4])), sub2(u.y, mul2(u.y, wini[aa - 4]))};
                }
            }
            fft8p(P);
#pragma unroll
            for (int k = 1; k < 8; ++k) P[k] = cmul2(P[k], ptwr[k - 1], ptwi[k - 1]);
            // 8 x 16 transpose through the group's swizzled patch; an entry is (re A, re B, im A, im B)
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 8; ++k)
                *reinterpret_cast<ulonglong2*>(myscr + k * 32 + ((g ^ k) << 2)) = make_ulonglong2(P[k].r, P[k].i);
            __syncwarp();
            cf bb[16];
            c2 Zp[8];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const float4 v = *reinterpret_cast<const float4*>(myscr + g * 32 + ((jj ^ g) << 2));
                bb[2 * jj] = {v.x, v.z};
                bb[2 * jj + 1] = {v.y, v.w};
            }
            fft16p(bb, Zp);   // Zp[k] = (Zc[g + 8 (2k)], Zc[g + 8 (2k + 1)])
            cf Z[16];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                upk2(Zp[k].r, Z[2 * k].r, Z[2 * k + 1].r);
                upk2(Zp[k].i, Z[2 * k].i, Z[2 * k + 1].i);
            }
            const int src_lane = (lane & 24) | ((8 - g) & 7);
            cf R[8];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const cf own = Z[8 + jj];
                const cf rot = Z[(9 + jj) & 15];
                const float sr = g == 0 ? rot.r : own.r;
                const float si = g == 0 ? rot.i : own.i;
                R[jj].r = __shfl_sync(0xffffffffu, sr, src_lane);
                R[jj].i = __shfl_sync(0xffffffffu, si, src_lane);
            }
            float* Lrow = Ls + r * LS_PITCH + (int)((row0 + r) & 3) + m;
            constexpr float LN2 = 0.69314718055994530942f;
            if (valid) {
                const f2 eps2 = bc2(a.log_eps4), ln2p = bc2(LN2), cp = bc2(-2.0f * LN2);
                f2 S1 = bc2(0.0f), S2 = bc2(0.0f);
#pragma unroll
                for (int j = 0; j < 4; ++j) {                          // bins k2 = 2j | 2j + 1, split step packed
                    const c2 zk = Zp[j];
                    const c2 zm = {pk2(R[7 - 2 * j].r, R[6 - 2 * j].r), pk2(R[7 - 2 * j].i, R[6 - 2 * j].i)};
                    const f2 Er = add2(zk.r, zm.r), Ei = sub2(zk.i, zm.i);
                    const f2 Dr = sub2(zk.r, zm.r), Di = add2(zk.i, zm.i);
                    const f2 Ttr = fma2(Di, pspr[j], mul2(Dr, pspi[j]));         // Tt = (D.i - i D.r) * (spr + i spi)
                    const f2 Tti = sub2(mul2(Di, pspi[j]), mul2(Dr, pspr[j]));
                    const f2 Ar = add2(Er, Ttr), Ai = add2(Ei, Tti), Br = sub2(Er, Ttr), Bi = sub2(Ei, Tti);
                    const f2 pa = fma2(Ar, Ar, fma2(Ai, Ai, eps2));
                    const f2 pb = fma2(Br, Br, fma2(Bi, Bi, eps2));
                    float pa0, pa1, pb0, pb1;
                    upk2(pa, pa0, pa1);
                    upk2(pb, pb0, pb1);
                    const f2 la = fma2(pk2(fast_log2(pa0), fast_log2(pa1)), ln2p, cp);
                    const f2 lb = fma2(pk2(fast_log2(pb0), fast_log2(pb1)), ln2p, cp);
                    float la0, la1, lb0, lb1;
                    upk2(la, la0, la1);
                    upk2(lb, lb0, lb1);
                    Lrow[(g + 16 * j) * NF] = la0;
                    Lrow[(g + 16 * j + 8) * NF] = la1;
                    Lrow[(128 - g - 16 * j) * NF] = lb0;
                    Lrow[(120 - g - 16 * j) * NF] = lb1;
                    S1 = add2(S1, add2(la, lb));
                    S2 = fma2(la, la, fma2(lb, lb, S2));
                }
                float s1, s1b, s2, s2b;
                upk2(S1, s1, s1b);
                upk2(S2, s2, s2b);
                s1 += s1b;
                s2 += s2b;
                if (g == 0) {   // bin 64 pairs with itself: |X[64]|^2 = |Zc[64]|^2
                    const cf zz = Z[8];
                    const float p = fmaf(4.0f * zz.r, zz.r, fmaf(4.0f * zz.i, zz.i, a.log_eps4));
                    const float l = fmaf(fast_log2(p), LN2, -2.0f * LN2);
                    Lrow[64 * NF] = l;
                    s1 += l;
                    s2 = fmaf(l, l, s2);
                }
                stat[(r * NF + m) * 8 + g] = make_float2(s1, s2);
            }
            } else {
            cf z0[8], z1[8];
#pragma unroll
            for (int aa = 0; aa < 8; ++aa) {
                const float4 u = *reinterpret_cast<const float4*>(yseg + aa * 32 + 4 * g);
                float v0, v1, v2, v3;
                if (aa < 4) {
                    v0 = u.x * win[aa * 4 + 0]; v1 = u.y * win[aa * 4 + 1];
                    v2 = u.z * win[aa * 4 + 2]; v3 = u.w * win[aa * 4 + 3];
                } else {   // hann[n + 128] = 1 - hann[n]
                    v0 = fmaf(-u.x, win[(aa - 4) * 4 + 0], u.x); v1 = fmaf(-u.y, win[(aa - 4) * 4 + 1], u.y);
                    v2 = fmaf(-u.z, win[(aa - 4) * 4 + 2], u.z); v3 = fmaf(-u.w, win[(aa - 4) * 4 + 3], u.w);
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
            // 8 x 16 transpose through the group's swizzled patch
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 8; ++k)
                *reinterpret_cast<float4*>(myscr + k * 32 + ((g ^ k) << 2)) =
                    make_float4(z0[k].r, z0[k].i, z1[k].r, z1[k].i);
            __syncwarp();
            cf bb[16], Z[16];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const float4 v = *reinterpret_cast<const float4*>(myscr + g * 32 + ((jj ^ g) << 2));
                bb[2 * jj] = {v.x, v.y};
                bb[2 * jj + 1] = {v.z, v.w};
            }
            fft16(bb, Z);   // Z[k2] = Zc[g + 8 k2]

            // conjugate partner: lane (8 - g) & 7 of the same group, index 15 - k2
            // (lane 0 pairs with itself at 16 - k2, so as a source it sends a rotated copy)
            const int src_lane = (lane & 24) | ((8 - g) & 7);
            cf R[8];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const cf own = Z[8 + jj];
                const cf rot = Z[(9 + jj) & 15];
                const float sr = g == 0 ? rot.r : own.r;
                const float si = g == 0 ? rot.i : own.i;
                R[jj].r = __shfl_sync(0xffffffffu, sr, src_lane);
                R[jj].i = __shfl_sync(0xffffffffu, si, src_lane);
            }
            float* Lrow = Ls + r * LS_PITCH + (int)((row0 + r) & 3) + m;
            float s1 = 0.0f, s2 = 0.0f;
            constexpr float LN2 = 0.69314718055994530942f;
            if (valid) {
#pragma unroll
                for (int k2 = 0; k2 < 8; ++k2) {
                    const cf zk = Z[k2], zm = R[7 - k2];                     // R[j - 8] holds index j
                    const cf E = {zk.r + zm.r, zk.i - zm.i};
                    const cf D = {zk.r - zm.r, zk.i + zm.i};
                    const cf O = {D.i, -D.r};
                    const cf Tt = cmul(O, spr[k2], spi[k2]);
                    const cf A = cadd(E, Tt), Bc = csub(E, Tt);
                    const float pa = fmaf(A.r, A.r, fmaf(A.i, A.i, a.log_eps4));
                    const float pb = fmaf(Bc.r, Bc.r, fmaf(Bc.i, Bc.i, a.log_eps4));
                    const float la = fmaf(fast_log2(pa), LN2, -2.0f * LN2);
                    const float lb = fmaf(fast_log2(pb), LN2, -2.0f * LN2);
                    Lrow[(g + 8 * k2) * NF] = la;
                    Lrow[(128 - g - 8 * k2) * NF] = lb;
                    s1 += la + lb;
                    s2 = fmaf(la, la, fmaf(lb, lb, s2));
                }
                if (g == 0) {   // bin 64 pairs with itself: |X[64]|^2 = |Zc[64]|^2
                    const cf zz = Z[8];
                    const float p = fmaf(4.0f * zz.r, zz.r, fmaf(4.0f * zz.i, zz.i, a.log_eps4));
                    const float l = fmaf(fast_log2(p), LN2, -2.0f * LN2);
                    Lrow[64 * NF] = l;
                    s1 += l;
                    s2 = fmaf(l, l, s2);
                }
                stat[(r * NF + m) * 8 + g] = make_float2(s1, s2);
            }
            }   // !PS
        }
        __syncthreads();

        // ---------------- row statistics + normalise + store ----------------
        // Every warp reduces the 264 partials of a row itself, in the same fixed order
        // (bit-stable, no atomics, no extra barrier): float within a lane, double across lanes.
        const int nrows = (a.rows - row0) < ROWS ? (int)(a.rows - row0) : ROWS;
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
            if constexpr (BS) {
                // warp w owns float4s [w * per, (w + 1) * per) of the 16-byte aligned body
                constexpr int NW = NT / 32;
                const int per = (n4 + NW - 1) / NW;
                const int v0 = warp * per, v1 = (v0 + per < n4) ? v0 + per : n4;
                float4* w4 = const_cast<float4*>(s4);
                for (int v = v0 + lane; v < v1; v += 32) {
                    const float4 l = w4[v];
                    w4[v] = make_float4(fmaf(l.x, inv, c), fmaf(l.y, inv, c), fmaf(l.z, inv, c), fmaf(l.w, inv, c));
                }
                if (tid < head) __stcs(dst + tid, fmaf(src[tid], inv, c));
                if (tid < tail) __stcs(dst + head + 4 * n4 + tid, fmaf(src[head + 4 * n4 + tid], inv, c));
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // normalised values -> bulk-copy engine
                __syncwarp();
                if (lane == 0 && v1 > v0) {
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                 ::"l"(d4 + v0), "r"(smem_u32(w4 + v0)), "r"((unsigned)(16 * (v1 - v0))) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            } else {
            for (int v = tid; v < n4; v += NT) {
                const float4 l = s4[v];
                __stcs(d4 + v, make_float4(fmaf(l.x, inv, c), fmaf(l.y, inv, c), fmaf(l.z, inv, c),
                                           fmaf(l.w, inv, c)));
            }
            if (tid < head) __stcs(dst + tid, fmaf(src[tid], inv, c));
            if (tid < tail) __stcs(dst + head + 4 * n4 + tid, fmaf(src[head + 4 * n4 + tid], inv, c));
            }
        }
        // The post-FIR barrier of the next iteration orders these reads of Ls / stat against the
        // next tile's STFT writes.
    }
    if constexpr (BS) {
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // shared memory stays valid until read
    }
}

// per-lane constant tables, computed in double
void fill_lane_tables(float* t) {
    const double two_pi = 6.283185307179586476925286766559;
    for (int g = 0; g < 8; ++g) {
        float* tb = t + g * LANE_TABLE;
        for (int aa = 0; aa < 4; ++aa)
            for (int e = 0; e < 4; ++e) {
                const int n = 32 * aa + 4 * g + e;
                tb[aa * 4 + e] = (float)(0.5 - 0.5 * cos(two_pi * n / 256.0));
            }
        for (int e = 0; e < 2; ++e)
            for (int k = 1; k < 8; ++k) {
                const double ang = -two_pi * (double)((2 * g + e) * k) / 128.0;
                tb[16 + (e * 7 + (k - 1)) * 2] = (float)cos(ang);
                tb[16 + (e * 7 + (k - 1)) * 2 + 1] = (float)sin(ang);
            }
        for (int k2 = 0; k2 < 8; ++k2) {
            const double ang = -two_pi * (double)(g + 8 * k2) / 256.0;
            tb[44 + 2 * k2] = (float)cos(ang);
            tb[44 + 2 * k2 + 1] = (float)sin(ang);
        }
    }
}

}  // namespace

namespace eegx {

bool dsp_tuned_supported(const eegx_dsp_plan* p) {
    return p->n_fft == 256 && p->hop == 64 && p->numtaps == 65 && p->T == T;
}

int dsp_tuned_table_floats() { return 8 * LANE_TABLE; }
void dsp_tuned_fill_tables(float* host) { fill_lane_tables(host); }

template <int ROWS, int NT, bool PK = false, bool TC = false, bool PS = false, int MINB = 0, bool BS = false>
int launch_variant(const TunedArgs& a, cudaStream_t st) {
    using C = Cfg<ROWS, NT>;
    EEGX_CUDA_CHECK(cudaFuncSetAttribute(dsp_tuned_kernel<ROWS, NT, PK, TC, PS, MINB, BS>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
    const long long ntiles = (a.rows + ROWS - 1) / ROWS;
    const long long max_ctas = (long long)(MINB ? MINB : C::CTAS_PER_SM) * kNumSMsB200;
    const int grid = (int)(ntiles < max_ctas ? ntiles : max_ctas);
    dsp_tuned_kernel<ROWS, NT, PK, TC, PS, MINB, BS><<<grid, NT, C::SMEM_BYTES, st>>>(a);
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}

int launch_dsp_tuned(const eegx_dsp_plan* plan, const DspArgs& d, cudaStream_t st) {
    EEGX_REQUIRE(d.onsets == nullptr, EEGX_ERR_ARG, "tuned kernel takes pre-cut trials only");
    EEGX_REQUIRE(plan->d_lane_tables != nullptr, EEGX_ERR_ARG, "plan has no tuned tables");
    TunedArgs a;
    a.x = d.x;
    a.out = d.out;
    a.rows = d.rows;
    a.lane_tables = plan->d_lane_tables;
    a.log_eps4 = 4.0f * plan->log_eps;
    a.z_eps = plan->z_eps;
    for (int i = 0; i < 65; ++i) a.taps_rev[i] = plan->h_taps[64 - i];
    switch (plan->tuned_variant) {
        case 3: return launch_dsp_pair(plan, d, st);    // packed f32x2, 2 rows per tile (dsp_tuned2.cu)
        case 4: return launch_variant<1, 96, true>(a, st);   // 1 row per tile, packed half-row FIR
        case 5: return launch_variant<1, 96, false, true>(a, st);    // 1 row per tile, FIR as 3xTF32 Toeplitz MMAs
        case 6: return launch_variant<1, 128, false, true>(a, st);   // same, 4 warps (16 FIR blocks divide evenly)
        case 11: return launch_variant<1, 96, false, false, false, 0, true>(a, st);   // default kernel + bulk stores
        case 12: return launch_variant<1, 128, false, false, false, 0, true>(a, st);  // same, 4 warps (3 CTAs/SM)
        case 13: return launch_variant<2, 192, false, false, false, 0, true>(a, st);  // same, 2 rows per tile
        case 8: return launch_variant<1, 96, false, false, true>(a, st);   // scalar FIR + packed f32x2 STFT
        case 9: return launch_variant<1, 128, false, false, true, 4>(a, st);  // same, 4 warps x 4 CTAs/SM (128 registers)
        case 10: return launch_variant<1, 96, false, false, true, 5>(a, st);  // same, 3 warps x 5 CTAs/SM (136 registers)
        case 7: return launch_dsp_umma(plan, d, st);    // FIR as tcgen05 kind::tf32 Toeplitz MMAs, operands in tensor memory (dsp_umma.cu)
        case 1: return launch_variant<1, 96>(a, st);    // 4 CTAs/SM x 3 warps, 1 row per tile
        case 2: return launch_variant<2, 192>(a, st);   // 2 CTAs/SM x 6 warps, 2 rows per tile
        default: return launch_variant<1, 96>(a, st);
    }
}

}  // namespace eegx
