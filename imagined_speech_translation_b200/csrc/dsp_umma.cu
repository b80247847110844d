// Fused DSP kernel for BASELINE config 2 (T = 2048, 65-tap FIR, n_fft = 256, hop = 64) with the FIR on the
// 5th-generation tensor cores: the band-pass filter is a banded Toeplitz product issued as tcgen05.mma
// kind::tf32 with the signal windows in TENSOR MEMORY (A operand) and the tap band in shared memory (B operand).
//
//   x (rows, 2048) f32  ->  out (rows, 129, 33) f32,   rows = B * C
//
// Spec: SURVEY.md section 8(c), identical to dsp_tuned.cu / dsp_generic.cu (cross-checked on the GPU).
//
// Why: the scalar kernel (dsp_tuned.cu) is bound by the FP32 pipe / issue slots, and 65 of its ~125 lane-ops per
// sample are the direct-form FIR.  Here the FIR costs the CUDA cores ~9 instructions per sample (split + copies),
// the tensor pipe does the multiply-adds asynchronously while the same SM runs the STFT of the previous tile.
//
// One persistent CTA per SM, 11 warps, a tile = 4 signal rows = 128 tensor-memory lanes:
//
//   lane (s, b)   row s of the tile (0..3), block b (0..31) of 64 consecutive outputs
//   A[lane][k]    = x_s[64 b - 32 + k],  k = 0..127   (zero outside the row: "same" convolution)
//   B[k'][n']     = taps_rev[k' - n'],   k' = 0..95, n' = 0..31  (zero outside the 65-tap band), one 96 x 32 band
//                   serves both 32-wide halves of a block because the band is shift invariant
//   D[lane][n]    = y_s[64 b + n] = sum_k A[lane][32 j + k'] B[k'][n - 32 j],   j = n / 32
//
// fp32 accuracy through the three-term TF32 split: x = x_hi + x_lo, h = h_hi + h_lo (hi = nearest 10-bit
// mantissa, lo = exact remainder), y = x_lo h_hi + x_hi h_lo + x_hi h_hi accumulated in fp32 in tensor memory
// (small terms first); the dropped x_lo h_lo term is 2^-22 relative.  72 MMAs (128 x 32 x 8) per tile.
//
//   load     TMA: one cp.async.bulk (global -> shared, 8 KB) per row, issued one tile ahead
//   convert  warps 0-7: thread = lane; 16 LDS.128 of its window, split, tcgen05.st (hi -> columns 0..127,
//            lo -> 128..255)
//   FIR      one thread issues the 72 tcgen05.mma (A from tensor memory, B by shared-memory descriptor,
//            128-byte swizzle, K-major) + tcgen05.commit -> mbarrier; it runs while the CTA does the STFT of
//            the previous tile
//   readout  warps 0-7: tcgen05.ld of D (columns 256..319) -> reflect-extended rows in shared memory
//   STFT / statistics / store: as dsp_tuned.cu (8 lanes per frame, radix-8 x 16 register FFT, fp64 fixed-order
//            statistics, coalesced 128-bit streaming stores), four rows per pass.
//
// Algorithmic HBM bytes per row: 8192 read + 17028 written; nothing else touches HBM.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "dsp_device.cuh"
#include "dsp_plan.h"

namespace {

using namespace eegx_dsp;

constexpr int T = 2048;
constexpr int NF = 33;
constexpr int F = 129;
constexpr int ROW_OUT = F * NF;             // 4257
constexpr int ROWS = 4;                     // 4 rows x 32 blocks = 128 lanes
constexpr int STFT_WARPS = 11;              // 33 STFT warp-tasks per tile = 3 rounds exactly
constexpr int CTRL_WARP = STFT_WARPS;       // warp 11: TMA loads, MMA issue, bulk stores
constexpr int NT = 32 * (STFT_WARPS + 1);   // 384
constexpr int NGROUPS = 4 * STFT_WARPS;     // 8-lane groups of the STFT warps
constexpr int NTASKS = 8 * ROWS + 1;
constexpr int ROUNDS = (NTASKS + STFT_WARPS - 1) / STFT_WARPS;
constexpr int XS_PITCH = 32 + T + 32 + 4;   // floats; 8464 B = 66 * 128 + 16
constexpr int YS_PITCH = 128 + T + 128 + 4; // floats; 9232 B = 72 * 128 + 16
constexpr int LS_PITCH = 4264;
constexpr int LANE_TABLE = 60;

// tap band: N = 32 rows (n') x K = 96 (k'), tf32 = 4 bytes, K-major, 128-byte swizzle:
// three 32-wide k blocks of 32 rows x 128 B = 4096 B each
constexpr int BAND_N = 32;
constexpr int BAND_K = 96;
constexpr int BAND_BYTES = BAND_N * BAND_K * 4;   // 12288

// shared-memory carve-up (bytes)
constexpr int OFF_BHI = 0;
constexpr int OFF_BLO = OFF_BHI + BAND_BYTES;
constexpr int OFF_XS = OFF_BLO + BAND_BYTES;
constexpr int OFF_YS = OFF_XS + ROWS * XS_PITCH * 4;
constexpr int OFF_LS = OFF_YS + ROWS * YS_PITCH * 4;
constexpr int OFF_SCR = OFF_LS + ROWS * LS_PITCH * 4;
constexpr int OFF_STAT = OFF_SCR + NGROUPS * 1024;
constexpr int OFF_BAR = OFF_STAT + ROWS * NF * 8 * 8;
constexpr int SMEM_BYTES = OFF_BAR + 32;
static_assert(OFF_XS % 1024 == 0 && OFF_YS % 16 == 0 && OFF_LS % 16 == 0 && OFF_SCR % 16 == 0 && OFF_STAT % 8 == 0 &&
              OFF_BAR % 8 == 0, "alignment of the shared-memory regions");
static_assert(SMEM_BYTES + 1024 <= 227 * 1024, "tile does not fit");

// tensor-memory columns
constexpr unsigned TM_COLS = 512;
constexpr unsigned TM_AHI = 0, TM_ALO = 128, TM_D = 256;

struct UmmaArgs {
    const float* x;
    float* out;
    long long rows;
    const float* lane_tables;  // [8][LANE_TABLE]
    const float* taps;         // device, h[0..64]
    float log_eps4;
    float z_eps;
    long long* prof;           // EEGX_DSP_PROF=1: per-phase clock64 sums of CTA 0 / thread 0 (debug only), else NULL
};

#define PROF_MARK(i)                                              \
    do {                                                          \
        if (a.prof != nullptr && blockIdx.x == 0 && tid == 0) {   \
            const long long now_ = clock64();                     \
            prof_acc[i] += now_ - prof_t;                         \
            prof_t = now_;                                        \
        }                                                         \
    } while (0)

// ---- packed f32x2 arithmetic: one instruction works on two independent FFTs (or two bins) at once ----
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(f2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2 bc(float x) { return pk(x, x); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

struct c2 { f2 r, i; };   // two complex values: low halves = FFT (or bin) A, high halves = FFT (or bin) B
__device__ __forceinline__ c2 cadd(c2 a, c2 b) { return {add2(a.r, b.r), add2(a.i, b.i)}; }
__device__ __forceinline__ c2 csub(c2 a, c2 b) { return {sub2(a.r, b.r), sub2(a.i, b.i)}; }
__device__ __forceinline__ c2 cadd_mni(c2 x, c2 a) { return {add2(x.r, a.i), sub2(x.i, a.r)}; }   // x + (-i) a
__device__ __forceinline__ c2 csub_mni(c2 x, c2 a) { return {sub2(x.r, a.i), add2(x.i, a.r)}; }   // x - (-i) a
__device__ __forceinline__ c2 cmul2(c2 a, f2 wr, f2 wi) {                                           // a * (wr + i wi)
    return {sub2(mul2(a.r, wr), mul2(a.i, wi)), fma2(a.r, wi, mul2(a.i, wr))};
}

// Two forward 8-point FFTs in lockstep (e^{-i...}), in place, natural order in and out.
__device__ __forceinline__ void fft8p(c2 (&v)[8]) {
    const f2 rs = bc(RSQRT2), nrs = bc(-RSQRT2);
    const c2 e0 = cadd(v[0], v[4]), e1 = cadd(v[1], v[5]), e2 = cadd(v[2], v[6]), e3 = cadd(v[3], v[7]);
    const c2 d0 = csub(v[0], v[4]), d1 = csub(v[1], v[5]), d2 = csub(v[2], v[6]), d3 = csub(v[3], v[7]);
    const c2 o1 = {mul2(add2(d1.r, d1.i), rs), mul2(sub2(d1.i, d1.r), rs)};      // d1 * W8^1
    const c2 o3 = {mul2(sub2(d3.i, d3.r), rs), mul2(add2(d3.r, d3.i), nrs)};     // d3 * W8^3
    const c2 s0 = cadd(e0, e2), s1 = csub(e0, e2), s2 = cadd(e1, e3), s3 = csub(e1, e3);
    v[0] = cadd(s0, s2); v[4] = csub(s0, s2); v[2] = cadd_mni(s1, s3); v[6] = csub_mni(s1, s3);
    const c2 t0 = cadd_mni(d0, d2), t1 = csub_mni(d0, d2), t2 = cadd(o1, o3), t3 = csub(o1, o3);
    v[1] = cadd(t0, t2); v[5] = csub(t0, t2); v[3] = cadd_mni(t1, t3); v[7] = csub_mni(t1, t3);
}

// Forward 16-point FFT of c[0..15]; the result comes back as pairs: out[k] = (X[2k], X[2k + 1]).  The even /
// odd split and the odd half's twiddles are scalar, the two 8-point FFTs run packed.
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
    for (int n = 0; n < 8; ++n) out[n] = {pk(e[n].r, o[n].r), pk(e[n].i, o[n].i)};
    fft8p(out);
}

// ---- tcgen05 helpers ----
__device__ __forceinline__ void tmem_alloc(unsigned smem_dst, unsigned ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(unsigned taddr, unsigned ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem descriptor], kind::tf32, M = 128
__device__ __forceinline__ void umma_tf32_ts(unsigned tmem_d, unsigned tmem_a, unsigned long long bdesc, unsigned idesc,
                                             unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// K-major, 128-byte swizzle: 8-row atoms of 1024 B stacked along N -> SBO = 1024 B (LBO unused)
__device__ __forceinline__ unsigned long long make_smem_desc(unsigned saddr) {
    unsigned long long d = 0;
    d |= (unsigned long long)((saddr >> 4) & 0x3FFF);
    d |= (unsigned long long)((16u >> 4) & 0x3FFF) << 16;
    d |= (unsigned long long)((1024u >> 4) & 0x3FFF) << 32;
    d |= 1ull << 46;   // descriptor version (Blackwell)
    d |= 2ull << 61;   // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void tmem_st32(unsigned taddr, const unsigned (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
          "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
          "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(unsigned taddr, unsigned (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}

__device__ __forceinline__ void split_tf32(float v, unsigned& hi, unsigned& lo) {
    hi = (__float_as_uint(v) + 0x1000u) & 0xFFFFE000u;            // nearest 10-bit mantissa
    lo = __float_as_uint(v - __uint_as_float(hi));                  // exact remainder (its low bits are dropped by the MMA)
}

__device__ __forceinline__ void issue_tile_loads(const UmmaArgs& a, float* xs, unsigned bar, long long row0) {
    const int nrows = (a.rows - row0) < ROWS ? (int)(a.rows - row0) : ROWS;
    mbar_expect_tx(bar, (unsigned)(nrows * T * sizeof(float)));
    for (int r = 0; r < nrows; ++r)
        tma_load_1d(smem_u32(xs + r * XS_PITCH + 32), a.x + (row0 + r) * (long long)T, (unsigned)(T * sizeof(float)), bar);
}

// warps 0-7: signal windows -> tensor memory (hi / lo).  Warp w serves lane quarter w & 3 and columns
// [64 (w >> 2), +64) of the 128-wide window; lane i = (row i & 3, block 8 (w & 3) + (i >> 2)).
__device__ __forceinline__ void convert_tile(const float* xs, unsigned tmem_base, int warp, int lane) {
    const int q = warp & 3, half = warp >> 2;
    const int s = lane & 3, blk = 8 * q + (lane >> 2);
    const float* src = xs + s * XS_PITCH + 64 * blk + 64 * half;
    const unsigned tlane = tmem_base + ((unsigned)(32 * q) << 16);
#pragma unroll
    for (int c2 = 0; c2 < 2; ++c2) {
        unsigned hi[32], lo[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float4 v = *reinterpret_cast<const float4*>(src + 32 * c2 + 4 * c);
            split_tf32(v.x, hi[4 * c + 0], lo[4 * c + 0]);
            split_tf32(v.y, hi[4 * c + 1], lo[4 * c + 1]);
            split_tf32(v.z, hi[4 * c + 2], lo[4 * c + 2]);
            split_tf32(v.w, hi[4 * c + 3], lo[4 * c + 3]);
        }
        tmem_st32(tlane + TM_AHI + 64 * half + 32 * c2, hi);
        tmem_st32(tlane + TM_ALO + 64 * half + 32 * c2, lo);
    }
    tmem_st_wait();
}

// one thread: the 72 MMAs of a tile + commit
__device__ __forceinline__ void issue_fir(unsigned tmem_base, unsigned bhi_addr, unsigned blo_addr, unsigned bar) {
    // instruction descriptor: D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), both K-major, N >> 3, M >> 4
    constexpr unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(BAND_N >> 3) << 17) | ((128u >> 4) << 24);
#pragma unroll 1
    for (int j = 0; j < 2; ++j) {
        unsigned acc = 0;
#pragma unroll 1
        for (int pass = 0; pass < 3; ++pass) {            // x_lo h_hi, x_hi h_lo, x_hi h_hi: small terms first
            const unsigned a_col = tmem_base + (pass == 0 ? TM_ALO : TM_AHI) + 32 * j;
            const unsigned b_addr = pass == 1 ? blo_addr : bhi_addr;
#pragma unroll
            for (int ks = 0; ks < BAND_K / 8; ++ks) {
                const unsigned long long bdesc = make_smem_desc(b_addr + (ks >> 2) * 4096 + (ks & 3) * 32);
                umma_tf32_ts(tmem_base + TM_D + 32 * j, a_col + 8 * ks, bdesc, idesc, acc);
                acc = 1;
            }
        }
    }
    umma_commit(bar);
}

// position of extended sample p inside its aligned group of four: (0, 1, 2, 3) -> (0, 2, 1, 3)
__device__ __forceinline__ int perm4(int p) { return (p & ~3) | ((p & 1) << 1) | ((p >> 1) & 1); }

// warps 0-7: D -> reflect-extended rows in shared memory.  Warp w: lane quarter w & 3, columns [32 (w >> 2), +32).
__device__ __forceinline__ void readout_tile(float* ys, unsigned tmem_base, int warp, int lane) {
    const int q = warp & 3, half = warp >> 2;
    const int s = lane & 3, blk = 8 * q + (lane >> 2);
    unsigned r[32];
    tmem_ld32(tmem_base + ((unsigned)(32 * q) << 16) + TM_D + 32 * half, r);
    float* yrow = ys + s * YS_PITCH;
    const int t0 = 64 * blk + 32 * half;
    float4* dst = reinterpret_cast<float4*>(yrow + 128 + t0);
#pragma unroll
    for (int c = 0; c < 8; ++c)
        dst[c] = make_float4(__uint_as_float(r[4 * c]), __uint_as_float(r[4 * c + 2]), __uint_as_float(r[4 * c + 1]),
                             __uint_as_float(r[4 * c + 3]));      // (t, t+2, t+1, t+3): see the STFT loads
    if (t0 <= 128) {          // reflect copy on the left: index -t for t in [1, 128]
#pragma unroll
        for (int e = 0; e < 32; ++e) {
            const int t = t0 + e;
            if (t >= 1 && t <= 128) yrow[perm4(128 - t)] = __uint_as_float(r[e]);
        }
    }
    if (t0 + 31 >= T - 129) {   // reflect copy on the right: index 2(T-1)-t for t in [T-129, T-2]
#pragma unroll
        for (int e = 0; e < 32; ++e) {
            const int t = t0 + e;
            if (t >= T - 129 && t <= T - 2) yrow[perm4(2 * (T - 1) - t + 128)] = __uint_as_float(r[e]);
        }
    }
}

__device__ __forceinline__ bool elect_one() {
    unsigned pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// shared -> global bulk copy (TMA store engine), tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_store(float* dst, unsigned src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__global__ void __launch_bounds__(NT, 1) dsp_umma_kernel(const __grid_constant__ UmmaArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    // the swizzled tap band needs 1024-byte aligned atoms: round the dynamic window up (1 KB of slack is allocated)
    unsigned char* smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    float* xs = reinterpret_cast<float*>(smem_raw + OFF_XS);
    float* ys = reinterpret_cast<float*>(smem_raw + OFF_YS);
    float* Ls = reinterpret_cast<float*>(smem_raw + OFF_LS);
    float* scr = reinterpret_cast<float*>(smem_raw + OFF_SCR);
    float2* stat = reinterpret_cast<float2*>(smem_raw + OFF_STAT);
    const unsigned tma_bar = smem_u32(smem_raw + OFF_BAR);
    const unsigned mma_bar = smem_u32(smem_raw + OFF_BAR + 8);
    volatile unsigned* tmem_slot = reinterpret_cast<volatile unsigned*>(smem_raw + OFF_BAR + 16);
    const unsigned bhi_addr = smem_u32(smem_raw + OFF_BHI), blo_addr = smem_u32(smem_raw + OFF_BLO);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = tid & 7, q = (tid >> 3) & 3;
    const bool ctrl = warp == CTRL_WARP;

    // per-lane constants (fixed for the lifetime of the CTA)
    // packed: window pairs (samples 0 | 2 and 1 | 3 of a group of four), stage-1 twiddles of the lane's two
    // 8-point FFTs, split-step twiddles of two neighbouring bins
    f2 winr[4], wini[4], twr[7], twi[7], spr[4], spi[4];
    {
        const float* tb = a.lane_tables + g * LANE_TABLE;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            winr[i] = pk(__ldg(tb + 4 * i + 0), __ldg(tb + 4 * i + 2));
            wini[i] = pk(__ldg(tb + 4 * i + 1), __ldg(tb + 4 * i + 3));
        }
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            twr[k] = pk(__ldg(tb + 16 + k * 2), __ldg(tb + 16 + (7 + k) * 2));
            twi[k] = pk(__ldg(tb + 16 + k * 2 + 1), __ldg(tb + 16 + (7 + k) * 2 + 1));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            spr[j] = pk(__ldg(tb + 44 + 4 * j), __ldg(tb + 44 + 4 * j + 2));
            spi[j] = pk(__ldg(tb + 44 + 4 * j + 1), __ldg(tb + 44 + 4 * j + 3));
        }
    }

    // tap band, split into TF32 hi / lo, written in the swizzled K-major layout the MMA descriptor names
    for (int idx = tid; idx < BAND_N * BAND_K; idx += NT) {
        const int n = idx / BAND_K, k = idx - n * BAND_K;
        const int d = k - n;                                         // B[k][n] = taps_rev[k - n] = h[64 - (k - n)]
        const float tv = (d >= 0 && d <= 64) ? __ldg(a.taps + 64 - d) : 0.0f;
        unsigned hi, lo;
        split_tf32(tv, hi, lo);
        const int kk = k & 31;
        const int off = (k >> 5) * 4096 + (n >> 3) * 1024 + (n & 7) * 128 + ((((kk >> 2) ^ (n & 7)) & 7) << 4) + (kk & 3) * 4;
        *reinterpret_cast<unsigned*>(smem_raw + OFF_BHI + off) = hi;
        *reinterpret_cast<unsigned*>(smem_raw + OFF_BLO + off) = lo;
    }
    // FIR zero halos (32 samples each side of every row), written once.
    for (int i = tid; i < ROWS * 64; i += NT) {
        const int r = i >> 6, h = i & 63;
        xs[r * XS_PITCH + (h < 32 ? h : T + h)] = 0.0f;
    }
    if (tid == 0) {
        mbar_init(tma_bar, 1);
        mbar_init(mma_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(smem_u32(const_cast<unsigned*>(tmem_slot)), TM_COLS);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // band writes (generic proxy) -> MMA reads (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem_base = *tmem_slot;

    const long long ntiles = (a.rows + ROWS - 1) / ROWS;
    const long long stride = gridDim.x;
    long long tile = blockIdx.x;
    unsigned tma_phase = 0, mma_phase = 0;

    // prologue: first tile's loads, conversion and FIR
    if (tile < ntiles) {
        if (ctrl && elect_one()) issue_tile_loads(a, xs, tma_bar, tile * ROWS);
        if (warp < 8) {
            mbar_wait(tma_bar, tma_phase);
            convert_tile(xs, tmem_base, warp, lane);
            tc_fence_before();
        }
        tma_phase ^= 1;
        __syncthreads();
        if (ctrl) {
            tc_fence_after();
            if (elect_one()) {
                issue_fir(tmem_base, bhi_addr, blo_addr, mma_bar);
                if (tile + stride < ntiles) issue_tile_loads(a, xs, tma_bar, (tile + stride) * ROWS);
            }
            __syncwarp();
        }
    }

    long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long prof_t = clock64();
    for (; tile < ntiles; tile += stride) {
        const long long row0 = tile * ROWS;
        const bool has_next = tile + stride < ntiles;
        const int nrows = (a.rows - row0) < ROWS ? (int)(a.rows - row0) : ROWS;

        // ---------------- FIR result of this tile: tensor memory -> ys ----------------
        if (warp < 8) {
            mbar_wait(mma_bar, mma_phase);
            PROF_MARK(0);
            tc_fence_after();
            readout_tile(ys, tmem_base, warp, lane);
            tc_fence_before();
        }
        mma_phase ^= 1;
        __syncthreads();
        PROF_MARK(1);

        // ---------------- next tile: convert + start its FIR (runs under this tile's STFT) ----------------
        if (has_next) {
            if (warp < 8) {
                mbar_wait(tma_bar, tma_phase);
                PROF_MARK(2);
                convert_tile(xs, tmem_base, warp, lane);
                tc_fence_before();
            }
            tma_phase ^= 1;
        }
        // the previous tile's bulk stores must have read Ls before the STFT below overwrites it
        if (ctrl) bulk_wait_read();
        __syncthreads();
        PROF_MARK(3);

        if (ctrl) {
            // ---------------- control warp: start the next tile's FIR and loads ----------------
            if (has_next) {
                tc_fence_after();
                if (elect_one()) {
                    issue_fir(tmem_base, bhi_addr, blo_addr, mma_bar);
                    if (tile + 2 * stride < ntiles) issue_tile_loads(a, xs, tma_bar, (tile + 2 * stride) * ROWS);
                }
                __syncwarp();
            }
        } else {
        // ------------------------------ STFT (warps 0-10) ------------------------------
        float* myscr = scr + (tid >> 3) * 256;
#pragma unroll 1
        for (int round = 0; round < ROUNDS; ++round) {
            const int task = warp + STFT_WARPS * round;
            if (task >= NTASKS) break;                       // warp-uniform
            // tasks 0..8*ROWS-1: row = task / 8, frames (task % 8) + 8 q;  last task: frame 32 of row q
            const bool full = task < 8 * ROWS;
            const int r = full ? (task >> 3) : q;
            const int m = full ? (task & 7) + 8 * q : 32;
            const float* yseg = ys + r * YS_PITCH + m * 64;  // extended position 64 m

            // ys holds every aligned group of four samples as (t, t+2, t+1, t+3): a 128-bit load is the packed pair
            // (re | re, im | im) of the lane's two interleaved 8-point FFTs, z[n] = y[2n] + i y[2n+1]
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
            for (int k = 1; k < 8; ++k) P[k] = cmul2(P[k], twr[k - 1], twi[k - 1]);
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
                upk(Zp[k].r, Z[2 * k].r, Z[2 * k + 1].r);
                upk(Zp[k].i, Z[2 * k].i, Z[2 * k + 1].i);
            }

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
            constexpr float LN2 = 0.69314718055994530942f;
            const f2 eps2 = bc(a.log_eps4), ln2p = bc(LN2), cp = bc(-2.0f * LN2);
            f2 S1 = bc(0.0f), S2 = bc(0.0f);
#pragma unroll
            for (int j = 0; j < 4; ++j) {                              // bins k2 = 2j | 2j + 1, split step packed
                const c2 zk = Zp[j];
                const c2 zm = {pk(R[7 - 2 * j].r, R[6 - 2 * j].r), pk(R[7 - 2 * j].i, R[6 - 2 * j].i)};   // R[j' - 8] holds index j'
                const f2 Er = add2(zk.r, zm.r), Ei = sub2(zk.i, zm.i);
                const f2 Dr = sub2(zk.r, zm.r), Di = add2(zk.i, zm.i);
                // Tt = (D.i - i D.r) * (spr + i spi)
                const f2 Ttr = fma2(Di, spr[j], mul2(Dr, spi[j]));
                const f2 Tti = sub2(mul2(Di, spi[j]), mul2(Dr, spr[j]));
                const f2 Ar = add2(Er, Ttr), Ai = add2(Ei, Tti), Br = sub2(Er, Ttr), Bi = sub2(Ei, Tti);
                const f2 pa = fma2(Ar, Ar, fma2(Ai, Ai, eps2));
                const f2 pb = fma2(Br, Br, fma2(Bi, Bi, eps2));
                float pa0, pa1, pb0, pb1;
                upk(pa, pa0, pa1);
                upk(pb, pb0, pb1);
                const f2 la = fma2(pk(fast_log2(pa0), fast_log2(pa1)), ln2p, cp);
                const f2 lb = fma2(pk(fast_log2(pb0), fast_log2(pb1)), ln2p, cp);
                float la0, la1, lb0, lb1;
                upk(la, la0, la1);
                upk(lb, lb0, lb1);
                Lrow[(g + 16 * j) * NF] = la0;
                Lrow[(g + 16 * j + 8) * NF] = la1;
                Lrow[(128 - g - 16 * j) * NF] = lb0;
                Lrow[(120 - g - 16 * j) * NF] = lb1;
                S1 = add2(S1, add2(la, lb));
                S2 = fma2(la, la, fma2(lb, lb, S2));
            }
            float s1, s1b, s2, s2b;
            upk(S1, s1, s1b);
            upk(S2, s2, s2b);
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
        }
        PROF_MARK(5);
        __syncthreads();
        PROF_MARK(6);

        // ---------------- row statistics + normalise in place (all 12 warps: row w & 3, third w >> 2) ----------------
        // The 264 partials of a row are reduced in a fixed order (bit-stable, no atomics): float within a lane,
        // double across lanes; the three warps of a row compute the same values.
        {
            const int r = warp & 3, part = warp >> 2;
            if (r < nrows) {
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
                float* src = Ls + r * LS_PITCH + ph;
                float* dst = a.out + (row0 + r) * (long long)ROW_OUT;
                const int head = (4 - ph) & 3;
                const int n4 = (ROW_OUT - head) >> 2;
                const int tail = ROW_OUT - head - 4 * n4;
                float4* s4 = reinterpret_cast<float4*>(src + head);
                const int j = part * 32 + lane;
                for (int v = j; v < n4; v += 96) {
                    const float4 l = s4[v];
                    s4[v] = make_float4(fmaf(l.x, inv, c), fmaf(l.y, inv, c), fmaf(l.z, inv, c), fmaf(l.w, inv, c));
                }
                // the (at most 3 + 3) floats outside the 16-byte aligned body go out directly
                if (part == 0 && lane < head) __stcs(dst + lane, fmaf(src[lane], inv, c));
                if (part == 1 && lane < tail) __stcs(dst + head + 4 * n4 + lane, fmaf(src[head + 4 * n4 + lane], inv, c));
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // normalised rows -> bulk-copy engine
        }
        __syncthreads();
        // ---------------- store: one bulk copy per row, read asynchronously while the next tile proceeds ----------------
        if (ctrl && elect_one()) {
            for (int r = 0; r < nrows; ++r) {
                const int ph = (int)((row0 + r) & 3);
                const int head = (4 - ph) & 3;
                const int n4 = (ROW_OUT - head) >> 2;
                bulk_store(a.out + (row0 + r) * (long long)ROW_OUT + head, smem_u32(Ls + r * LS_PITCH + ph + head), 16u * n4);
            }
            bulk_commit();
        }
        PROF_MARK(7);
    }
    if (a.prof != nullptr && blockIdx.x == 0 && tid == 0)
        for (int i = 0; i < 8; ++i) a.prof[i] = prof_acc[i];

    if (ctrl) bulk_wait_all();      // shared memory must stay valid until the last bulk stores have read it
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, TM_COLS);
}

}  // namespace

namespace eegx {

int launch_dsp_umma(const eegx_dsp_plan* plan, const DspArgs& d, cudaStream_t st) {
    EEGX_REQUIRE(d.onsets == nullptr, EEGX_ERR_ARG, "tuned kernel takes pre-cut trials only");
    EEGX_REQUIRE(plan->d_lane_tables != nullptr, EEGX_ERR_ARG, "plan has no tuned tables");
    UmmaArgs a;
    a.x = d.x;
    a.out = d.out;
    a.rows = d.rows;
    a.lane_tables = plan->d_lane_tables;
    a.taps = d.taps;
    a.log_eps4 = 4.0f * plan->log_eps;
    a.z_eps = plan->z_eps;
    a.prof = nullptr;
    static const bool prof_on = getenv("EEGX_DSP_PROF") != nullptr;
    static long long* d_prof = nullptr;
    if (prof_on) {
        if (!d_prof) EEGX_CUDA_CHECK(cudaMalloc(&d_prof, 8 * sizeof(long long)));
        a.prof = d_prof;
    }
    EEGX_CUDA_CHECK(cudaFuncSetAttribute(dsp_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES + 1024));
    const long long ntiles = (a.rows + ROWS - 1) / ROWS;
    const int grid = (int)(ntiles < kNumSMsB200 ? ntiles : kNumSMsB200);
    dsp_umma_kernel<<<grid, NT, SMEM_BYTES + 1024, st>>>(a);
    EEGX_CUDA_CHECK(cudaGetLastError());
    if (prof_on) {   // debug only: synchronises
        long long h[8];
        EEGX_CUDA_CHECK(cudaMemcpy(h, d_prof, sizeof(h), cudaMemcpyDeviceToHost));
        fprintf(stderr, "[dsp_umma prof] clk: mma_wait %lld readout+sync %lld tma_wait %lld convert+sync %lld mma_issue %lld "
                "stft %lld stft_sync %lld stats+store %lld\n", h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
    }
    return EEGX_OK;
}

}  // namespace eegx
