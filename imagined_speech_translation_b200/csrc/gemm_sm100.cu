// bf16 GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM, operands
// staged by TMA) -- the contraction core of the encoder ("library-call
// sites that become sm_100a kernels": conv-as-GEMM, Linear, attention/FFN projections).
//
//   D[b] (M x N) = A[b] (M x K) * B[b] (N x K)^T  (+ bias[N]) (+ erf-GELU) (+ D[b] when accumulating)
//
// A and B are bf16, each either K-major (K contiguous; "row-major" M x K / N x K) or MN-major
// (stored K x M / K x N), selected per operand -- so the forward (x W^T), the data gradient
// (dy W) and the weight gradient (dy^T x) of a Linear layer all run on this one kernel without
// transposing anything in HBM.  D is bf16 or fp32, row-major.  fp32 accumulation.
//
// Structure (persistent, warp-specialised; one CTA per SM, 192 threads):
//   warp 0     TMA producer: cp.async.bulk.tensor (128B swizzle) into a STAGES-deep smem ring,
//              completion on "full" mbarriers.
//   warp 1     MMA issuer: one elected lane issues tcgen05.mma (128 x BLOCK_N x 16 per
//              instruction); tcgen05.commit releases smem slots ("empty") and publishes the
//              accumulator ("tmem_full").  Two accumulator stages in TMEM so the epilogue of
//              tile i overlaps the main loop of tile i+1.
//   warps 2-5  epilogue: tcgen05.ld (TMEM -> registers), bias / GELU / accumulate, convert,
//              128-bit global stores; then hand the TMEM stage back ("tmem_empty").
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "eegx_common.h"

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;          // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;
constexpr int NUM_EPI_WARPS = 4;
constexpr unsigned WATCHDOG_SPINS = 1u << 27;

struct GemmParams {
    long long M, N, K, batch;      // batch = inner * groups problems; problem bi = (g = bi / inner, s = bi % inner)
    long long inner;               // problems per group (split-K chunks); 1 for a plain grouped launch
    long long ldd, stride_d;       // D: row pitch, stride between the `inner` problems of a group
    long long stride_d_g;          // D stride between groups
    long long stride_bias_g;       // bias stride between groups (0: one bias for all)
    const float* bias;
    void* D;
    int out_f32;
    int epilogue;      // 0 none, 1 bias, 2 bias + gelu
    int accumulate;    // D += result (fp32 or bf16 read-modify-write)
    float alpha;
    int vec_ok;        // D rows are 16-byte addressable: coalesced 128-bit epilogue stores
    int debug;         // EEGX_GEMM_DEBUG (profiling experiments only): 1 = skip the global stores, 2 = skip the whole epilogue body
};

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Spin with a watchdog: a protocol bug traps (kernel error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > WATCHDOG_SPINS) asm volatile("trap;");
    }
}

// 4-D operand maps: (inner dim, rows, problem inside the group, group)
__device__ __forceinline__ void tma_load_4d(unsigned dst, const CUtensorMap* map, unsigned bar, int c0,
                                            int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

__device__ __forceinline__ void tmem_alloc(unsigned smem_dst, unsigned ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(unsigned taddr, unsigned ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc,
                                          unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma have completed.
__device__ __forceinline__ void umma_commit(unsigned bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// TMEM -> registers, 32 lanes x 32 columns.  The load is asynchronous: `tmem_ld32_issue` only starts it
// and `tmem_ld_wait` (which names the registers as in/out operands, so that no use can be scheduled in
// front of it) completes it -- this lets the epilogue fetch chunk c+1 while it converts and stores chunk c.
__device__ __forceinline__ void tmem_ld32_issue(unsigned taddr, unsigned (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait(unsigned (&r)[32]) {
    asm volatile(
        "tcgen05.wait::ld.sync.aligned;"
        : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
          "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
          "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
          "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
        :: "memory");
}

// Shared-memory matrix descriptor, 128-byte swizzle (cute::UMMA::SmemDescriptor, version 1).
//   K-major : rows of 128 B (64 bf16 along K), 8-row atoms of 1024 B stacked along M/N -> SBO = 1024 B.
//   MN-major: rows of 128 B (64 bf16 along M/N), one row per k; 8-k atoms of 1024 B stacked along K
//             -> SBO = 1024 B; the next 64-wide M/N block is a separate TMA box -> LBO = box bytes.
__device__ __forceinline__ unsigned long long make_smem_desc(unsigned saddr, unsigned lbo_bytes, unsigned sbo_bytes) {
    unsigned long long d = 0;
    d |= (unsigned long long)((saddr >> 4) & 0x3FFF);
    d |= (unsigned long long)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (unsigned long long)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= 1ull << 46;   // descriptor version (Blackwell)
    d |= 2ull << 61;   // SWIZZLE_128B
    return d;
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// ---------------------------------------------------------------------------------------- epilogue
// A warp owns 32 accumulator rows (one per lane in TMEM).  Writing them straight from that layout
// makes every store instruction touch 32 different cache lines, and measured on the B200 this -- not
// the tensor pipe -- bounded the K = 768 encoder GEMMs (9472 x 3072 x 768: 47 us with the stores,
// 31.6 us without).  So the tile is transposed through a small per-warp shared-memory patch
// (32 rows x 128 bytes, 16-byte row padding: conflict free both ways) and leaves as full 128-byte row
// segments: 4 lines per store instruction instead of 32.  TMEM reads are double buffered (the load
// of chunk c+1 is in flight while chunk c is converted).
constexpr int EPI_ROW_BYTES = 144;                       // 128 data + 16 pad
constexpr int EPI_WARP_BYTES = 32 * EPI_ROW_BYTES;       // 4608
constexpr int EPI_BYTES = NUM_EPI_WARPS * EPI_WARP_BYTES;

// alpha / bias / GELU on one 32-column chunk of this lane's row
__device__ __forceinline__ void epilogue_math(const GemmParams& p, const float* __restrict__ bias, const unsigned (&r)[32],
                                              float (&v)[32], long long col0) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * p.alpha;
    if (p.epilogue >= 1) {
        if (col0 + 32 <= p.N && (reinterpret_cast<uintptr_t>(bias) & 15) == 0) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + col0 + i));
                v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (col0 + i < p.N) v[i] += __ldg(bias + col0 + i);
        }
    }
    if (p.epilogue == 2) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
    }
}

// scalar fallback: this lane writes its own row (unaligned D / ragged N)
__device__ __forceinline__ void epilogue_store_direct(const GemmParams& p, const float (&v)[32], bool row_ok, long long row,
                                                      long long col0, long long d_off) {
    if (!row_ok || col0 >= p.N) return;
    const long long off = d_off + row * p.ldd + col0;
    if (p.out_f32) {
        float* d = reinterpret_cast<float*>(p.D) + off;
        for (int i = 0; i < 32; ++i)
            if (col0 + i < p.N) d[i] = p.accumulate ? d[i] + v[i] : v[i];
    } else {
        __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(p.D) + off;
        for (int i = 0; i < 32; ++i)
            if (col0 + i < p.N) d[i] = __float2bfloat16(p.accumulate ? __bfloat162float(d[i]) + v[i] : v[i]);
    }
}

__device__ __forceinline__ unsigned pack_bf16(float lo, float hi) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const unsigned*>(&h);
}

// this lane's 32 values -> its row of the patch, at byte offset `boff` (0 or 64 for bf16, 0 for fp32)
__device__ __forceinline__ void patch_write(unsigned char* patch, int lane, int boff, const float (&v)[32], bool f32) {
    uint4* dst = reinterpret_cast<uint4*>(patch + lane * EPI_ROW_BYTES + boff);
    if (f32) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            dst[i] = make_uint4(__float_as_uint(v[4 * i]), __float_as_uint(v[4 * i + 1]), __float_as_uint(v[4 * i + 2]),
                                __float_as_uint(v[4 * i + 3]));
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            dst[i] = make_uint4(pack_bf16(v[8 * i], v[8 * i + 1]), pack_bf16(v[8 * i + 2], v[8 * i + 3]),
                                pack_bf16(v[8 * i + 4], v[8 * i + 5]), pack_bf16(v[8 * i + 6], v[8 * i + 7]));
    }
}

// the patch (32 rows x 128 bytes) -> global: 8 lanes write one row's 128 bytes, 4 rows per instruction.
// With `accumulate` the eight old values are fetched first, all in flight together (a load -> add -> store
// chain per row exposed eight global-memory latencies per chunk and made the small weight-gradient GEMMs,
// which accumulate straight into the gradient buffers, ~3x slower than their forward twins).
__device__ __forceinline__ void patch_flush(const GemmParams& p, const unsigned char* patch, int lane, long long row0,
                                            long long col0, long long d_off) {
    const int piece = lane & 7;
    const int epp = p.out_f32 ? 4 : 8;                   // elements per 16-byte piece
    const long long col = col0 + piece * epp;
    const bool col_ok = col < p.N;
    const size_t es = p.out_f32 ? 4 : 2;
    unsigned char* base = static_cast<unsigned char*>(p.D) + (size_t)(d_off + col) * es;
    uint4 old[8];
    if (p.accumulate) {
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const long long row = row0 + 4 * it + (lane >> 3);
            old[it] = make_uint4(0u, 0u, 0u, 0u);
            if (row < p.M && col_ok) old[it] = *reinterpret_cast<const uint4*>(base + (size_t)row * p.ldd * es);
        }
    }
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int r = 4 * it + (lane >> 3);
        const long long row = row0 + r;
        if (row < p.M && col_ok) {
            uint4 val = *reinterpret_cast<const uint4*>(patch + r * EPI_ROW_BYTES + piece * 16);
            if (p.accumulate) {
                unsigned* w = &val.x;
                const unsigned* o = &old[it].x;
                if (p.out_f32) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) w[k] = __float_as_uint(__uint_as_float(w[k]) + __uint_as_float(o[k]));
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[k]));
                        const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&o[k]));
                        w[k] = pack_bf16(a.x + b.x, a.y + b.y);
                    }
                }
            }
            *reinterpret_cast<uint4*>(base + (size_t)row * p.ldd * es) = val;
        }
    }
}

// Epilogue of one accumulator tile for the 32 rows starting at row0 that this warp owns.
template <int BLOCK_N>
__device__ __forceinline__ void epilogue_rows(const GemmParams& p, unsigned taddr, long long row0, int lane, long long n_base,
                                              long long bi, unsigned char* patch) {
    const long long grp = bi / p.inner;
    const long long d_off = grp * p.stride_d_g + (bi - grp * p.inner) * p.stride_d;     // element offset of this problem's D
    const float* bias = p.bias + grp * p.stride_bias_g;
    constexpr int NC = BLOCK_N / 32;
    static_assert(NC % 2 == 0, "BLOCK_N must be a multiple of 64");
    if (p.debug == 2) return;
    const long long row = row0 + lane;
    const bool row_ok = row < p.M;
    const bool f32 = p.out_f32 != 0;
    unsigned r0[32], r1[32];
    float v[32];
    tmem_ld32_issue(taddr, r0);
    tmem_ld_wait(r0);
#pragma unroll 1
    for (int c = 0; c < NC; c += 2) {
        tmem_ld32_issue(taddr + (c + 1) * 32, r1);
        const long long col0 = n_base + c * 32;
        epilogue_math(p, bias, r0, v, col0);
        if (p.debug != 1) {
            if (!p.vec_ok) {
                epilogue_store_direct(p, v, row_ok, row, col0, d_off);
            } else {
                patch_write(patch, lane, 0, v, f32);
                if (f32) {
                    __syncwarp();
                    patch_flush(p, patch, lane, row0, col0, d_off);
                    __syncwarp();
                }
            }
        }
        tmem_ld_wait(r1);
        if (c + 2 < NC) tmem_ld32_issue(taddr + (c + 2) * 32, r0);
        epilogue_math(p, bias, r1, v, col0 + 32);
        if (p.debug != 1) {
            if (!p.vec_ok) {
                epilogue_store_direct(p, v, row_ok, row, col0 + 32, d_off);
            } else {
                patch_write(patch, lane, f32 ? 0 : 64, v, f32);
                __syncwarp();
                patch_flush(p, patch, lane, row0, f32 ? col0 + 32 : col0, d_off);
                __syncwarp();
            }
        }
        if (c + 2 < NC) tmem_ld_wait(r0);
    }
}

template <int BLOCK_N, int STAGES>
struct SmemLayout {
    static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
    static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int NUM_BARS = 2 * STAGES + 4;
    static constexpr int EPI_OFFSET = BAR_OFFSET + NUM_BARS * 8 + 16;     // per-warp epilogue transpose patches
    static constexpr int TOTAL = EPI_OFFSET + EPI_BYTES + 1024;          // + alignment slack
};

template <int BLOCK_N, int STAGES, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const GemmParams p) {
    using L = SmemLayout<BLOCK_N, STAGES>;
    EEGX_PDL_TRIGGER();        // the next kernel's CTAs may be scheduled behind this grid's
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte alignment required by the 128B swizzle atoms
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const unsigned smem_base = smem_u32(smem);
    const unsigned bar_base = smem_base + L::BAR_OFFSET;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tmem_full_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
    auto tmem_empty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
    unsigned* tmem_ptr_smem = reinterpret_cast<unsigned*>(smem + L::BAR_OFFSET + L::NUM_BARS * 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr unsigned TMEM_COLS = BLOCK_N <= 64 ? 128 : (BLOCK_N <= 128 ? 256 : 512);   // two accumulator stages, power of two

    const long long m_blocks = (p.M + BLOCK_M - 1) / BLOCK_M;
    const long long n_blocks = (p.N + BLOCK_N - 1) / BLOCK_N;
    const long long tiles_per_batch = m_blocks * n_blocks;
    const long long num_tiles = tiles_per_batch * p.batch;
    const int num_kb = (int)((p.K + BLOCK_K - 1) / BLOCK_K);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<unsigned long long>(&map_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<unsigned long long>(&map_b)) : "memory");
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tmem_full_bar(s), 1);
            mbar_init(tmem_empty_bar(s), NUM_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_ptr_smem), TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem_base = *tmem_ptr_smem;
    EEGX_PDL_WAIT();           // barrier init, TMEM allocation and descriptor prefetch above overlap the previous kernel's tail

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            unsigned phase = 0;
            for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const long long bi = tile / tiles_per_batch;
                const long long rem = tile - bi * tiles_per_batch;
                const int n_blk = (int)(rem / m_blocks), m_blk = (int)(rem % m_blocks);   // m fastest: B tile stays hot in L2
                const int cg = (int)(bi / p.inner), cs = (int)(bi - (long long)cg * p.inner);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    const unsigned a_dst = smem_base + stage * L::STAGE_BYTES;
                    const unsigned b_dst = a_dst + L::A_BYTES;
                    mbar_expect_tx(full_bar(stage), L::STAGE_BYTES);
                    if (!A_MN) {
                        tma_load_4d(a_dst, &map_a, full_bar(stage), kb * BLOCK_K, m_blk * BLOCK_M, cs, cg);
                    } else {
#pragma unroll
                        for (int h = 0; h < BLOCK_M / 64; ++h)
                            tma_load_4d(a_dst + h * (64 * BLOCK_K * 2), &map_a, full_bar(stage),
                                        m_blk * BLOCK_M + h * 64, kb * BLOCK_K, cs, cg);
                    }
                    if (!B_MN) {
                        tma_load_4d(b_dst, &map_b, full_bar(stage), kb * BLOCK_K, n_blk * BLOCK_N, cs, cg);
                    } else {
#pragma unroll
                        for (int h = 0; h < BLOCK_N / 64; ++h)
                            tma_load_4d(b_dst + h * (64 * BLOCK_K * 2), &map_b, full_bar(stage),
                                        n_blk * BLOCK_N + h * 64, kb * BLOCK_K, cs, cg);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            // instruction descriptor: D = F32, A = B = BF16, majors, N >> 3, M >> 4
            const unsigned idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) |
                                   ((B_MN ? 1u : 0u) << 16) | ((unsigned)(BLOCK_N >> 3) << 17) |
                                   ((unsigned)(BLOCK_M >> 4) << 24);
            int stage = 0;
            unsigned phase = 0;
            int acc = 0;
            unsigned acc_phase = 0;
            for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(tmem_empty_bar(acc), acc_phase ^ 1);
                tc_fence_after();
                const unsigned tmem_d = tmem_base + acc * BLOCK_N;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const unsigned a_addr = smem_base + stage * L::STAGE_BYTES;
                    const unsigned b_addr = a_addr + L::A_BYTES;
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        // K-major: +32 bytes per 16 elements inside the swizzle row.
                        // MN-major: +2 k-atoms of 1024 bytes.
                        const unsigned long long adesc =
                            A_MN ? make_smem_desc(a_addr + k * 2048, 64 * BLOCK_K * 2, 1024)
                                 : make_smem_desc(a_addr + k * 32, 16, 1024);
                        const unsigned long long bdesc =
                            B_MN ? make_smem_desc(b_addr + k * 2048, 64 * BLOCK_K * 2, 1024)
                                 : make_smem_desc(b_addr + k * 32, 16, 1024);
                        umma_bf16(tmem_d, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(empty_bar(stage));          // frees the smem slot when the MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(tmem_full_bar(acc));            // accumulator complete
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int quarter = warp & 3;                       // TMEM lane quarter this warp may access
        int acc = 0;
        unsigned acc_phase = 0;
        for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const long long bi = tile / tiles_per_batch;
            const long long rem = tile - bi * tiles_per_batch;
            const int n_blk = (int)(rem / m_blocks), m_blk = (int)(rem % m_blocks);
            mbar_wait(tmem_full_bar(acc), acc_phase);
            tc_fence_after();
            const long long row0 = (long long)m_blk * BLOCK_M + quarter * 32;
            const unsigned taddr = tmem_base + acc * BLOCK_N + ((unsigned)(quarter * 32) << 16);
            epilogue_rows<BLOCK_N>(p, taddr, row0, lane, (long long)n_blk * BLOCK_N, bi,
                                   smem + L::EPI_OFFSET + (warp - 2) * EPI_WARP_BYTES);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}


// ==================================================================================================
// CTA-pair variant (tcgen05 cta_group::2): a cluster of two CTAs computes a 256 x BLOCK_N tile.
// Each CTA stages its own 128 rows of A and HALF of the B tile (BLOCK_N / 2 rows), the leader CTA
// issues one tcgen05.mma.cta_group::2 per 16-wide k slice that reads both CTAs' shared memory and
// writes 128 accumulator rows into each CTA's TMEM.  Operand bytes fetched per MAC drop by a third
// against the 1-CTA 128 x 256 tile, which is what bounds this kernel: the L2 -> SM path, not the
// tensor pipe (B300_MICROARCH.md "LTS throughput cap").
//   full barrier   lives in the leader; the leader's producer arms it with the bytes of BOTH CTAs,
//                  and both CTAs' TMA loads complete on it (2-SM TMA form, peer bit cleared).
//   empty / tmem_full barriers exist in both CTAs; the leader's tcgen05.commit multicasts to both.
//   tmem_empty     lives in the leader; the epilogue warps of both CTAs arrive on it (the peer remotely).
// ==================================================================================================
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(unsigned dst, const CUtensorMap* map, unsigned leader_bar, int c0,
                                                int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
        ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(map)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
          "l"(0x1000000000000000ull)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(unsigned smem_dst, unsigned ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(unsigned taddr, unsigned ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc,
                                              unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// Arrive (once the previously issued MMAs retire) on the barrier at this offset in BOTH CTAs of the pair.
__device__ __forceinline__ void umma_commit_2sm(unsigned bar) {
    asm volatile(
        "{\n\t.reg .b16 m;\n\tmov.b16 m, 3;\n\t"
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}"
        ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta0(unsigned bar) {
    asm volatile(
        "{\n\t.reg .b32 rem;\n\t"
        "mapa.shared::cluster.u32 rem, %0, 0;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [rem];\n\t}" ::"r"(bar) : "memory");
}

template <int BLOCK_N, int STAGES>
struct SmemLayout2 {
    static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
    static constexpr int B_BYTES = (BLOCK_N / 2) * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int NUM_BARS = 2 * STAGES + 4;
    static constexpr int EPI_OFFSET = BAR_OFFSET + NUM_BARS * 8 + 16;
    static constexpr int TOTAL = EPI_OFFSET + EPI_BYTES + 1024;
};

template <int BLOCK_N, int STAGES, bool A_MN, bool B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                  const GemmParams p) {
    using L = SmemLayout2<BLOCK_N, STAGES>;
    constexpr int HALF_N = BLOCK_N / 2;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const unsigned smem_base = smem_u32(smem);
    const unsigned bar_base = smem_base + L::BAR_OFFSET;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tmem_full_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
    auto tmem_empty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
    unsigned* tmem_ptr_smem = reinterpret_cast<unsigned*>(smem + L::BAR_OFFSET + L::NUM_BARS * 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned rank = cluster_ctarank();
    const bool leader = rank == 0;
    constexpr unsigned TMEM_COLS = BLOCK_N <= 128 ? 256 : 512;   // two accumulator stages, power of two

    const long long m_blocks = (p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
    const long long n_blocks = (p.N + BLOCK_N - 1) / BLOCK_N;
    const long long tiles_per_batch = m_blocks * n_blocks;
    const long long num_tiles = tiles_per_batch * p.batch;
    const int num_kb = (int)((p.K + BLOCK_K - 1) / BLOCK_K);
    const long long cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<unsigned long long>(&map_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<unsigned long long>(&map_b)) : "memory");
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tmem_full_bar(s), 1);
            mbar_init(tmem_empty_bar(s), 2 * NUM_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc_2sm(smem_u32(tmem_ptr_smem), TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const unsigned tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            int stage = 0;
            unsigned phase = 0;
            for (long long tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                const long long bi = tile / tiles_per_batch;
                const long long rem = tile - bi * tiles_per_batch;
                const int n_blk = (int)(rem / m_blocks), m_blk = (int)(rem % m_blocks);
                const int m0 = m_blk * 2 * BLOCK_M + (int)rank * BLOCK_M;
                const int n0 = n_blk * BLOCK_N + (int)rank * HALF_N;
                const int cg = (int)(bi / p.inner), cs = (int)(bi - (long long)cg * p.inner);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    const unsigned a_dst = smem_base + stage * L::STAGE_BYTES;
                    const unsigned b_dst = a_dst + L::A_BYTES;
                    const unsigned lead_full = full_bar(stage) & 0xFEFFFFFFu;     // the leader's barrier (peer bit cleared)
                    if (leader) mbar_expect_tx(full_bar(stage), 2 * L::STAGE_BYTES);
                    if (!A_MN) {
                        tma_load_4d_2sm(a_dst, &map_a, lead_full, kb * BLOCK_K, m0, cs, cg);
                    } else {
#pragma unroll
                        for (int h = 0; h < BLOCK_M / 64; ++h)
                            tma_load_4d_2sm(a_dst + h * (64 * BLOCK_K * 2), &map_a, lead_full, m0 + h * 64, kb * BLOCK_K, cs, cg);
                    }
                    if (!B_MN) {
                        tma_load_4d_2sm(b_dst, &map_b, lead_full, kb * BLOCK_K, n0, cs, cg);
                    } else {
#pragma unroll
                        for (int h = 0; h < HALF_N / 64; ++h)
                            tma_load_4d_2sm(b_dst + h * (64 * BLOCK_K * 2), &map_b, lead_full, n0 + h * 64, kb * BLOCK_K, cs, cg);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader && lane == 0) {
            const unsigned idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) |
                                   ((B_MN ? 1u : 0u) << 16) | ((unsigned)(BLOCK_N >> 3) << 17) |
                                   ((unsigned)((2 * BLOCK_M) >> 4) << 24);
            int stage = 0;
            unsigned phase = 0;
            int acc = 0;
            unsigned acc_phase = 0;
            for (long long tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                mbar_wait(tmem_empty_bar(acc), acc_phase ^ 1);
                tc_fence_after();
                const unsigned tmem_d = tmem_base + acc * BLOCK_N;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const unsigned a_addr = smem_base + stage * L::STAGE_BYTES;
                    const unsigned b_addr = a_addr + L::A_BYTES;
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        const unsigned long long adesc =
                            A_MN ? make_smem_desc(a_addr + k * 2048, 64 * BLOCK_K * 2, 1024)
                                 : make_smem_desc(a_addr + k * 32, 16, 1024);
                        const unsigned long long bdesc =
                            B_MN ? make_smem_desc(b_addr + k * 2048, 64 * BLOCK_K * 2, 1024)
                                 : make_smem_desc(b_addr + k * 32, 16, 1024);
                        umma_bf16_2sm(tmem_d, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit_2sm(empty_bar(stage));
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit_2sm(tmem_full_bar(acc));
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue (warps 2..5, both CTAs) =====================
        const int quarter = warp & 3;
        int acc = 0;
        unsigned acc_phase = 0;
        for (long long tile = cluster_id; tile < num_tiles; tile += num_clusters) {
            const long long bi = tile / tiles_per_batch;
            const long long rem = tile - bi * tiles_per_batch;
            const int n_blk = (int)(rem / m_blocks), m_blk = (int)(rem % m_blocks);
            mbar_wait(tmem_full_bar(acc), acc_phase);
            tc_fence_after();
            const long long row0 = (long long)m_blk * 2 * BLOCK_M + (long long)rank * BLOCK_M + quarter * 32;
            const unsigned taddr = tmem_base + acc * BLOCK_N + ((unsigned)(quarter * 32) << 16);
            epilogue_rows<BLOCK_N>(p, taddr, row0, lane, (long long)n_blk * BLOCK_N, bi,
                                   smem + L::EPI_OFFSET + (warp - 2) * EPI_WARP_BYTES);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (leader) mbar_arrive(tmem_empty_bar(acc));
                else mbar_arrive_cta0(tmem_empty_bar(acc));
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) tmem_dealloc_2sm(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// 4-D bf16 tensor map: dims (inner, rows, batch, groups); box (box_inner, box_rows, 1, 1); 128B swizzle.
int make_map(CUtensorMap* map, const void* base, long long inner, long long rows, long long batch, long long groups,
             long long ld_elems, long long batch_stride_elems, long long group_stride_elems, int box_inner, int box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    EEGX_REQUIRE(enc != nullptr, EEGX_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    const long long bs = batch > 1 ? batch_stride_elems : ld_elems * rows;        // size-1 dims still need a legal stride
    const long long gs = groups > 1 ? group_stride_elems : bs * batch;
    cuuint64_t dims[4] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)batch, (cuuint64_t)groups};
    cuuint64_t strides[3] = {(cuuint64_t)ld_elems * 2, (cuuint64_t)bs * 2, (cuuint64_t)gs * 2};
    cuuint32_t box[4] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    EEGX_REQUIRE(r == CUDA_SUCCESS, EEGX_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d "
                 "(inner=%lld rows=%lld batch=%lld groups=%lld ld=%lld)", (int)r, inner, rows, batch, groups, ld_elems);
    return EEGX_OK;
}

template <int BLOCK_N, int STAGES, bool A_MN, bool B_MN>
int launch(const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p, int grid, cudaStream_t st) {
    using L = SmemLayout<BLOCK_N, STAGES>;
    auto kern = gemm_bf16_kernel<BLOCK_N, STAGES, A_MN, B_MN>;
    EEGX_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    EEGX_CUDA_CHECK(eegx::launch(kern, grid, NUM_THREADS, L::TOTAL, st, ma, mb, p));
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}

template <int BLOCK_N, int STAGES, bool A_MN, bool B_MN>
int launch2(const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p, int grid, cudaStream_t st) {
    using L = SmemLayout2<BLOCK_N, STAGES>;
    auto kern = gemm2_bf16_kernel<BLOCK_N, STAGES, A_MN, B_MN>;
    EEGX_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    kern<<<grid, NUM_THREADS, L::TOTAL, st>>>(ma, mb, p);
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}

template <int BLOCK_N, int STAGES>
int dispatch_major2(bool a_mn, bool b_mn, const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p,
                    int grid, cudaStream_t st) {
    if (!a_mn && !b_mn) return launch2<BLOCK_N, STAGES, false, false>(ma, mb, p, grid, st);
    if (!a_mn && b_mn) return launch2<BLOCK_N, STAGES, false, true>(ma, mb, p, grid, st);
    if (a_mn && !b_mn) return launch2<BLOCK_N, STAGES, true, false>(ma, mb, p, grid, st);
    return launch2<BLOCK_N, STAGES, true, true>(ma, mb, p, grid, st);
}

template <int BLOCK_N, int STAGES>
int dispatch_major(bool a_mn, bool b_mn, const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p,
                   int grid, cudaStream_t st) {
    if (!a_mn && !b_mn) return launch<BLOCK_N, STAGES, false, false>(ma, mb, p, grid, st);
    if (!a_mn && b_mn) return launch<BLOCK_N, STAGES, false, true>(ma, mb, p, grid, st);
    if (a_mn && !b_mn) return launch<BLOCK_N, STAGES, true, false>(ma, mb, p, grid, st);
    return launch<BLOCK_N, STAGES, true, true>(ma, mb, p, grid, st);
}

}  // namespace

extern "C" int eegx_gemm_bf16(const eegx_gemm_desc* d, const void* A, const void* B, const float* bias,
                              void* D, void* stream) {
    if (int rc = eegx::require_sm100()) return rc;
    EEGX_REQUIRE(d && A && B && D, EEGX_ERR_ARG, "desc/A/B/D must not be NULL");
    EEGX_REQUIRE(d->M > 0 && d->N > 0 && d->K > 0 && d->batch > 0, EEGX_ERR_SHAPE,
                 "bad sizes M=%lld N=%lld K=%lld batch=%lld", (long long)d->M, (long long)d->N,
                 (long long)d->K, (long long)d->batch);
    EEGX_REQUIRE(d->epilogue >= 0 && d->epilogue <= 2, EEGX_ERR_ARG, "bad epilogue %d", d->epilogue);
    EEGX_REQUIRE(d->epilogue == 0 || bias != nullptr, EEGX_ERR_ARG, "epilogue %d needs a bias", d->epilogue);
    EEGX_REQUIRE(eegx::aligned16(A) && eegx::aligned16(B) && eegx::aligned16(D), EEGX_ERR_ALIGN,
                 "A, B, D must be 16-byte aligned");
    EEGX_REQUIRE((d->lda % 8) == 0 && (d->ldb % 8) == 0 && (d->stride_a % 8) == 0 && (d->stride_b % 8) == 0,
                 EEGX_ERR_ALIGN, "lda/ldb/batch strides must be multiples of 8 elements (TMA: 16 bytes)");
    const long long groups = d->groups > 1 ? d->groups : 1;
    EEGX_REQUIRE(d->groups >= 0 && (groups == 1 || ((d->stride_a_g % 8) == 0 && (d->stride_b_g % 8) == 0)), EEGX_ERR_ALIGN,
                 "group strides must be multiples of 8 elements (TMA: 16 bytes)");
    const long long problems = d->batch * groups;
    // lda / ldb may be SMALLER than the contiguous extent: rows then overlap in memory, which is
    // how a channels-last Conv1d runs as an implicit-im2col GEMM (row r = k consecutive time steps).
    EEGX_REQUIRE(d->lda >= 8 && d->ldb >= 8 && d->ldd >= d->N, EEGX_ERR_SHAPE,
                 "leading dimensions too small (lda=%lld ldb=%lld ldd=%lld)", (long long)d->lda,
                 (long long)d->ldb, (long long)d->ldd);

    // Tile width by a wave-quantisation cost model: a launch takes `rounds` passes of the 148 SMs (74 CTA
    // pairs) over the tiles and a pass costs ~ BLOCK_N x a per-width efficiency factor (narrow tiles
    // move more operand bytes per MAC; measured with tools/bench_gemm.py).
    static const int env_2cta = [] { const char* v = getenv("EEGX_GEMM_2CTA"); return v ? atoi(v) : 3; }();
    const bool bm_ = d->b_mn_major != 0;
    auto rounds_for = [&](int bn, bool pr) {
        const long long mb_ = pr ? (d->M + 2 * BLOCK_M - 1) / (2 * BLOCK_M) : (d->M + BLOCK_M - 1) / BLOCK_M;
        const long long t = mb_ * ((d->N + bn - 1) / bn) * problems;
        const long long slots = pr ? eegx::kNumSMsB200 / 2 : eegx::kNumSMsB200;
        return (double)((t + slots - 1) / slots);
    };
    // CTA pairs (cta_group::2, 256-row tiles).  Stand-alone (tools/bench_gemm.py) they gain +16 % on the
    // 4096 x 51264 x 768 LM head and +5 % at 8192^3 but lose 3..6 % on a single 768-wide encoder GEMM.  Inside the
    // step the lock-step region encoders launch those GEMMs with four times the rows, and there the pairs win:
    // EEGX_GEMM_2CTA = 1 (large N / K only) 26.76 ms/step, 2 (always) 26.35, 3 (default: large N / K, or >= 16384 rows
    // over all problems of the launch) 26.16 ms/step at B = 256 (profiles/r2_gemm_pair_policy.txt).  0 = never.
    const bool want_pair = env_2cta != 0 && d->M > BLOCK_M &&
                           (env_2cta == 2 || d->N >= 8192 || d->K >= 4096 || (env_2cta == 3 && d->M * problems >= 16384));
    int block_n = 128;
    {
        const int cand[4] = {64, 128, 192, 256};
        const double eff[4] = {1.7, 1.35, 1.04, 1.0};
        double best = 1e300;
        for (int i = 0; i < 4; ++i) {
            const int bn = cand[i];
            if (bn == 64 && d->N > 64) continue;                       // 64 only for very narrow outputs
            if (bn > 64 && d->N <= 64) continue;
            // 192: stand-alone it removes the half-empty last pass of the N = 768 encoder GEMMs (+3..5 % over 256), but
            // inside the step, where four region streams share the SMs, it measured 0.3 ms slower: force_block_n only
            if (bn == 192) continue;
            const bool pr = want_pair && bn >= 128;
            const double c = rounds_for(bn, pr) * (bn + 16) * eff[i];
            if (c < best) { best = c; block_n = bn; }
        }
    }
    if (d->force_block_n == 64 || d->force_block_n == 128 || d->force_block_n == 192 || d->force_block_n == 256)
        block_n = d->force_block_n;
    const bool pair = want_pair && block_n >= 128 && !(block_n == 192 && bm_);
    const int b_box_rows = pair ? block_n / 2 : block_n;

    CUtensorMap ma, mb;
    int rc;
    if (!d->a_mn_major) rc = make_map(&ma, A, d->K, d->M, d->batch, groups, d->lda, d->stride_a, d->stride_a_g, BLOCK_K, BLOCK_M);
    else rc = make_map(&ma, A, d->M, d->K, d->batch, groups, d->lda, d->stride_a, d->stride_a_g, 64, BLOCK_K);
    if (rc) return rc;
    if (!d->b_mn_major) rc = make_map(&mb, B, d->K, d->N, d->batch, groups, d->ldb, d->stride_b, d->stride_b_g, BLOCK_K, b_box_rows);
    else rc = make_map(&mb, B, d->N, d->K, d->batch, groups, d->ldb, d->stride_b, d->stride_b_g, 64, BLOCK_K);
    if (rc) return rc;

    GemmParams p;
    p.M = d->M; p.N = d->N; p.K = d->K; p.batch = problems; p.inner = d->batch;
    p.ldd = d->ldd; p.stride_d = d->stride_d;
    p.stride_d_g = groups > 1 ? d->stride_d_g : 0;
    p.stride_bias_g = groups > 1 ? d->stride_bias_g : 0;
    p.bias = bias; p.D = D;
    p.out_f32 = d->out_f32; p.epilogue = d->epilogue; p.accumulate = d->accumulate;
    p.alpha = d->alpha;
    static const int env_debug = [] { const char* v = getenv("EEGX_GEMM_DEBUG"); return v ? atoi(v) : 0; }();
    p.debug = env_debug;
    {
        const long long es = d->out_f32 ? 4 : 2, epp = 16 / es;
        p.vec_ok = (d->ldd % epp) == 0 && (d->N % epp) == 0 && (d->batch == 1 || (d->stride_d % epp) == 0) &&
                   (groups == 1 || (d->stride_d_g % epp) == 0);
    }

    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool am = d->a_mn_major != 0, bm = d->b_mn_major != 0;
    const long long n_blocks = (d->N + block_n - 1) / block_n;
    if (pair) {
        const long long m_blocks2 = (d->M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
        const long long tiles2 = m_blocks2 * n_blocks * problems;
        const long long max_clusters = eegx::kNumSMsB200 / 2;
        const int grid2 = 2 * (int)(tiles2 < max_clusters ? tiles2 : max_clusters);
        if (block_n == 128) return dispatch_major2<128, 8>(am, bm, ma, mb, p, grid2, st);
        if (block_n == 192) {      // K-major B only (see above)
            if (am) return launch2<192, 7, true, false>(ma, mb, p, grid2, st);
            return launch2<192, 7, false, false>(ma, mb, p, grid2, st);
        }
        return dispatch_major2<256, 6>(am, bm, ma, mb, p, grid2, st);
    }
    const long long m_blocks = (d->M + BLOCK_M - 1) / BLOCK_M;
    const long long tiles = m_blocks * n_blocks * problems;
    const int grid = (int)(tiles < eegx::kNumSMsB200 ? tiles : eegx::kNumSMsB200);
    switch (block_n) {
        case 64: return dispatch_major<64, 8>(am, bm, ma, mb, p, grid, st);
        case 128: return dispatch_major<128, 6>(am, bm, ma, mb, p, grid, st);
        case 192: return dispatch_major<192, 5>(am, bm, ma, mb, p, grid, st);
        default: return dispatch_major<256, 4>(am, bm, ma, mb, p, grid, st);
    }
}
