// Fused multi-head attention core for short sequences (S_q, S_k <= 64), forward and backward.
//
// Replaces, inside nn.MultiheadAttention (main_model/src/models/layers.py:232-234, 245-251;
// brain_encoder.py:165-168) the chain  q k^T / sqrt(hd) -> softmax -> dropout -> (.) v  and its
// autograd graph.  With the STFT front-end the encoder's sequences are S = N_f + 4 = 37 tokens
// (4 in the fusion stage), so one (batch, head) pair fits a CTA: q, k, v tiles live in shared
// memory, the S x S scores live in registers, the probabilities never touch HBM.  The matmuls
// run on mma.sync m16n8k16 bf16 (warp-level tensor-core path; at these sizes a tcgen05 tile of
// M = 128 would be > 70 % padding), fp32 accumulate; backward recomputes P from the saved
// log-sum-exp (flash-attention style) and regenerates the dropout mask from (seed, step, site).
//
// q, k, v, o (and their gradients) are addressed as rows of (B*S, row_stride) matrices with head h
// at columns [h*hd, (h+1)*hd): the packed QKV projection output is consumed, and the packed dQKV
// is produced, without any transpose / split / contiguous copy.
#include "eegx_common.h"
#include "fused_common.cuh"

namespace {

using namespace eegx;
typedef __nv_bfloat16 bf16;

struct AttnArgs {
    const bf16 *q, *k, *v, *o, *d_o;
    bf16 *out, *dq, *dk, *dv;
    float* lse;        // (B, H, Sq)
    long long q_rs, k_rs, v_rs, o_rs, dq_rs, dk_rs, dv_rs;
    int B, H, Sq, Sk, causal;
    float scale;
    DropoutCfg dc;
};

__device__ __forceinline__ uint32_t lds32(const bf16* p) { return *reinterpret_cast<const uint32_t*>(p); }

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
}

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// B fragment (k = 16 rows of X starting at `row0`, n = 8 columns starting at `col0`) of a row-major
// shared-memory tile X[k][n], i.e. the transposed 8x8 loads.
__device__ __forceinline__ void ldsm_x2_trans(uint32_t& b0, uint32_t& b1, const bf16* tile, int ld, int row0,
                                              int col0, int lane) {
    const unsigned addr = (unsigned)__cvta_generic_to_shared(tile + (row0 + (lane & 15)) * ld + col0);
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n" : "=r"(b0), "=r"(b1) : "r"(addr));
}

// global (S rows x HD, row stride rs) -> shared tile (rows_pad x (HD + 8)), rows >= S zero filled.
// cp.async (LDGSTS) 16-byte copies: every thread's copies are in flight together and the caller waits
// once (cp_async_wait_all + __syncthreads) -- a register-staged loop exposed one global-memory latency
// per iteration, which dominated these short kernels.
template <int HD>
__device__ __forceinline__ void load_tile(bf16* dst, const bf16* src, long long rs, int S, int rows_pad) {
    constexpr int V = HD / 8, LD = HD + 8;
    for (int idx = threadIdx.x; idx < rows_pad * V; idx += blockDim.x) {
        const int r = idx / V, v = idx % V;
        const unsigned d = (unsigned)__cvta_generic_to_shared(dst + r * LD + v * 8);
        const bf16* g = src + (long long)(r < S ? r : 0) * rs + v * 8;
        const int bytes = r < S ? 16 : 0;                 // src-size 0: the 16 destination bytes are zero filled
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(g), "r"(bytes) : "memory");
    }
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// A fragment of rows [row0, row0+16), k-slice ks of a row-major tile
template <int LD>
__device__ __forceinline__ void load_a(uint32_t (&a)[4], const bf16* tile, int row0, int ks, int g, int t) {
    const bf16* p = tile + (row0 + g) * LD + ks * 16 + 2 * t;
    a[0] = lds32(p);
    a[1] = lds32(p + 8 * LD);
    a[2] = lds32(p + 8);
    a[3] = lds32(p + 8 * LD + 8);
}

// dropout multiplier of probability (i, j) of (batch*head) bh: group = 8 consecutive keys
__device__ __forceinline__ unsigned long long pgroup(int bh, int i, int j) {
    return ((unsigned long long)bh * 64ull + (unsigned long long)i) * 8ull + (unsigned long long)(j >> 3);
}

// ------------------------------------------------------------------------------------------ forward
template <int HD, int NT>
__global__ void __launch_bounds__(NT * 32)
attn_fwd_kernel(const AttnArgs a) {
    EEGX_PDL_SYNC();
    constexpr int LD = HD + 8, ROWS = NT * 16;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    bf16* Qs = reinterpret_cast<bf16*>(smem_raw);
    bf16* Ks = Qs + ROWS * LD;
    bf16* Vs = Ks + ROWS * LD;
    const int bh = blockIdx.x, b = bh / a.H, h = bh % a.H;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    load_tile<HD>(Qs, a.q + (long long)b * a.Sq * a.q_rs + h * HD, a.q_rs, a.Sq, ROWS);
    load_tile<HD>(Ks, a.k + (long long)b * a.Sk * a.k_rs + h * HD, a.k_rs, a.Sk, ROWS);
    load_tile<HD>(Vs, a.v + (long long)b * a.Sk * a.v_rs + h * HD, a.v_rs, a.Sk, ROWS);
    cp_async_wait_all();
    __syncthreads();
    const int r0 = warp * 16;
    if (r0 >= a.Sq) return;

    float s[2 * NT][4];
#pragma unroll
    for (int nt = 0; nt < 2 * NT; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) s[nt][e] = 0.0f;
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks) {
        uint32_t af[4];
        load_a<LD>(af, Qs, r0, ks, g, t);
#pragma unroll
        for (int nt = 0; nt < 2 * NT; ++nt) {
            const bf16* p = Ks + (nt * 8 + g) * LD + ks * 16 + 2 * t;
            mma16816(s[nt], af, lds32(p), lds32(p + 8));
        }
    }
    // scale, mask, softmax over the keys (rows i_lo = r0+g in regs 0,1; i_hi = r0+g+8 in regs 2,3)
    const int i_lo = r0 + g, i_hi = i_lo + 8;
    float mx_lo = -INFINITY, mx_hi = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 2 * NT; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int j = nt * 8 + 2 * t + (e & 1), i = e < 2 ? i_lo : i_hi;
            const bool ok = j < a.Sk && (!a.causal || j <= i);
            s[nt][e] = ok ? s[nt][e] * a.scale : -INFINITY;
            if (e < 2) mx_lo = fmaxf(mx_lo, s[nt][e]); else mx_hi = fmaxf(mx_hi, s[nt][e]);
        }
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
    float sum_lo = 0.0f, sum_hi = 0.0f;
#pragma unroll
    for (int nt = 0; nt < 2 * NT; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float p = __expf(s[nt][e] - (e < 2 ? mx_lo : mx_hi));
            s[nt][e] = p;
            if (e < 2) sum_lo += p; else sum_hi += p;
        }
    sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 1);
    sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 2);
    sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 1);
    sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 2);
    if (t == 0) {
        if (i_lo < a.Sq) a.lse[(long long)bh * a.Sq + i_lo] = mx_lo + __logf(sum_lo);
        if (i_hi < a.Sq) a.lse[(long long)bh * a.Sq + i_hi] = mx_hi + __logf(sum_hi);
    }
    const float inv_lo = 1.0f / sum_lo, inv_hi = 1.0f / sum_hi;
    const DropoutGen gen(a.dc);
#pragma unroll
    for (int nt = 0; nt < 2 * NT; ++nt) {
        float m0, m1, m2, m3;
        gen.mask_pair(pgroup(bh, i_lo, nt * 8), 2 * t, m0, m1);
        gen.mask_pair(pgroup(bh, i_hi, nt * 8), 2 * t, m2, m3);
        s[nt][0] *= inv_lo * m0; s[nt][1] *= inv_lo * m1;
        s[nt][2] *= inv_hi * m2; s[nt][3] *= inv_hi * m3;
    }
    // O = P V
    float acc[HD / 8][4];
#pragma unroll
    for (int nd = 0; nd < HD / 8; ++nd)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[nd][e] = 0.0f;
#pragma unroll
    for (int kk = 0; kk < NT; ++kk) {
        const uint32_t af[4] = {pack2(s[2 * kk][0], s[2 * kk][1]), pack2(s[2 * kk][2], s[2 * kk][3]),
                                pack2(s[2 * kk + 1][0], s[2 * kk + 1][1]), pack2(s[2 * kk + 1][2], s[2 * kk + 1][3])};
#pragma unroll
        for (int nd = 0; nd < HD / 8; ++nd) {
            uint32_t b0, b1;
            ldsm_x2_trans(b0, b1, Vs, LD, kk * 16, nd * 8, lane);
            mma16816(acc[nd], af, b0, b1);
        }
    }
    bf16* o_lo = a.out + ((long long)b * a.Sq + i_lo) * a.o_rs + h * HD + 2 * t;
    bf16* o_hi = a.out + ((long long)b * a.Sq + i_hi) * a.o_rs + h * HD + 2 * t;
#pragma unroll
    for (int nd = 0; nd < HD / 8; ++nd) {
        if (i_lo < a.Sq) *reinterpret_cast<uint32_t*>(o_lo + nd * 8) = pack2(acc[nd][0], acc[nd][1]);
        if (i_hi < a.Sq) *reinterpret_cast<uint32_t*>(o_hi + nd * 8) = pack2(acc[nd][2], acc[nd][3]);
    }
}

// ------------------------------------------------------------------------------------------ backward
template <int HD, int NT>
__device__ __forceinline__ void store_rows(bf16* base, long long rs, int row_lo, int S, int t, const float (&acc)[HD / 8][4]) {
    bf16* lo = base + (long long)row_lo * rs + 2 * t;
    bf16* hi = lo + 8 * rs;
#pragma unroll
    for (int nd = 0; nd < HD / 8; ++nd) {
        if (row_lo < S) *reinterpret_cast<uint32_t*>(lo + nd * 8) = pack2(acc[nd][0], acc[nd][1]);
        if (row_lo + 8 < S) *reinterpret_cast<uint32_t*>(hi + nd * 8) = pack2(acc[nd][2], acc[nd][3]);
    }
}

// 2*NT warps: warps [0, NT) produce dK / dV of their 16 keys (phase A), warps [NT, 2*NT) produce dQ of
// their 16 queries (phase B), concurrently, from the same shared-memory tiles.
template <int HD, int NT>
__global__ void __launch_bounds__(2 * NT * 32, 2)
attn_bwd_kernel(const AttnArgs a) {
    EEGX_PDL_SYNC();
    constexpr int LD = HD + 8, ROWS = NT * 16;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    bf16* Qs = reinterpret_cast<bf16*>(smem_raw);
    bf16* Ks = Qs + ROWS * LD;
    bf16* Vs = Ks + ROWS * LD;
    bf16* dOs = Vs + ROWS * LD;
    float* lse_s = reinterpret_cast<float*>(dOs + ROWS * LD);
    float* D_s = lse_s + ROWS;
    const int bh = blockIdx.x, b = bh / a.H, h = bh % a.H;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    load_tile<HD>(Qs, a.q + (long long)b * a.Sq * a.q_rs + h * HD, a.q_rs, a.Sq, ROWS);
    load_tile<HD>(Ks, a.k + (long long)b * a.Sk * a.k_rs + h * HD, a.k_rs, a.Sk, ROWS);
    load_tile<HD>(Vs, a.v + (long long)b * a.Sk * a.v_rs + h * HD, a.v_rs, a.Sk, ROWS);
    load_tile<HD>(dOs, a.d_o + (long long)b * a.Sq * a.o_rs + h * HD, a.o_rs, a.Sq, ROWS);
    for (int i = threadIdx.x; i < ROWS; i += blockDim.x) lse_s[i] = i < a.Sq ? a.lse[(long long)bh * a.Sq + i] : 0.0f;
    cp_async_wait_all();
    __syncthreads();
    // D_i = dO_i . O_i
    for (int i = warp; i < ROWS; i += 2 * NT) {
        float acc = 0.0f;
        if (i < a.Sq && lane < HD / 8) {
            float dv[8], ov[8];
            load8(dOs + i * LD + lane * 8, dv);
            load8(a.o + ((long long)b * a.Sq + i) * a.o_rs + h * HD + lane * 8, ov);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc = fmaf(dv[e], ov[e], acc);
        }
        acc = warp_sum_f(acc);
        if (lane == 0) D_s[i] = acc;
    }
    __syncthreads();
    const DropoutGen gen(a.dc);

    // ---- phase A: this warp owns keys j0 .. j0+15; transposed scores S^T (rows j, columns i)
    const int j0 = warp * 16;
    if (warp < NT && j0 < a.Sk) {
        float sT[2 * NT][4], dpT[2 * NT][4];
#pragma unroll
        for (int nt = 0; nt < 2 * NT; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) sT[nt][e] = dpT[nt][e] = 0.0f;
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) {
            uint32_t ak[4], av[4];
            load_a<LD>(ak, Ks, j0, ks, g, t);
            load_a<LD>(av, Vs, j0, ks, g, t);
#pragma unroll
            for (int nt = 0; nt < 2 * NT; ++nt) {
                const bf16* pq = Qs + (nt * 8 + g) * LD + ks * 16 + 2 * t;
                const bf16* pd = dOs + (nt * 8 + g) * LD + ks * 16 + 2 * t;
                mma16816(sT[nt], ak, lds32(pq), lds32(pq + 8));
                mma16816(dpT[nt], av, lds32(pd), lds32(pd + 8));
            }
        }
#pragma unroll
        for (int nt = 0; nt < 2 * NT; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = j0 + g + (e >= 2 ? 8 : 0), i = nt * 8 + 2 * t + (e & 1);
                const bool ok = j < a.Sk && i < a.Sq && (!a.causal || j <= i);
                const float p = ok ? __expf(sT[nt][e] * a.scale - lse_s[i]) : 0.0f;
                const float m = gen.mask_one(pgroup(bh, i, j), j & 7);
                sT[nt][e] = p * m;                                           // dropped probabilities (for dV)
                dpT[nt][e] = p * (dpT[nt][e] * m - D_s[i]) * a.scale;        // dS^T (scaled)
            }
        float acc[HD / 8][4];
        // dV = Pd^T dO
#pragma unroll
        for (int nd = 0; nd < HD / 8; ++nd)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[nd][e] = 0.0f;
#pragma unroll
        for (int kk = 0; kk < NT; ++kk) {
            const uint32_t af[4] = {pack2(sT[2 * kk][0], sT[2 * kk][1]), pack2(sT[2 * kk][2], sT[2 * kk][3]),
                                    pack2(sT[2 * kk + 1][0], sT[2 * kk + 1][1]), pack2(sT[2 * kk + 1][2], sT[2 * kk + 1][3])};
#pragma unroll
            for (int nd = 0; nd < HD / 8; ++nd) {
                uint32_t b0, b1;
                ldsm_x2_trans(b0, b1, dOs, LD, kk * 16, nd * 8, lane);
                mma16816(acc[nd], af, b0, b1);
            }
        }
        store_rows<HD, NT>(a.dv + (long long)b * a.Sk * a.dv_rs + h * HD, a.dv_rs, j0 + g, a.Sk, t, acc);
        // dK = dS^T Q
#pragma unroll
        for (int nd = 0; nd < HD / 8; ++nd)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[nd][e] = 0.0f;
#pragma unroll
        for (int kk = 0; kk < NT; ++kk) {
            const uint32_t af[4] = {pack2(dpT[2 * kk][0], dpT[2 * kk][1]), pack2(dpT[2 * kk][2], dpT[2 * kk][3]),
                                    pack2(dpT[2 * kk + 1][0], dpT[2 * kk + 1][1]), pack2(dpT[2 * kk + 1][2], dpT[2 * kk + 1][3])};
#pragma unroll
            for (int nd = 0; nd < HD / 8; ++nd) {
                uint32_t b0, b1;
                ldsm_x2_trans(b0, b1, Qs, LD, kk * 16, nd * 8, lane);
                mma16816(acc[nd], af, b0, b1);
            }
        }
        store_rows<HD, NT>(a.dk + (long long)b * a.Sk * a.dk_rs + h * HD, a.dk_rs, j0 + g, a.Sk, t, acc);
    }

    // ---- phase B: this warp owns queries i0 .. i0+15; scores S (rows i, columns j); dQ = dS K
    const int i0 = (warp - NT) * 16;
    if (warp >= NT && i0 < a.Sq) {
        float s[2 * NT][4], dp[2 * NT][4];
#pragma unroll
        for (int nt = 0; nt < 2 * NT; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) s[nt][e] = dp[nt][e] = 0.0f;
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) {
            uint32_t aq[4], ad[4];
            load_a<LD>(aq, Qs, i0, ks, g, t);
            load_a<LD>(ad, dOs, i0, ks, g, t);
#pragma unroll
            for (int nt = 0; nt < 2 * NT; ++nt) {
                const bf16* pk = Ks + (nt * 8 + g) * LD + ks * 16 + 2 * t;
                const bf16* pv = Vs + (nt * 8 + g) * LD + ks * 16 + 2 * t;
                mma16816(s[nt], aq, lds32(pk), lds32(pk + 8));
                mma16816(dp[nt], ad, lds32(pv), lds32(pv + 8));
            }
        }
#pragma unroll
        for (int nt = 0; nt < 2 * NT; ++nt) {
            float m[4];
            gen.mask_pair(pgroup(bh, i0 + g, nt * 8), 2 * t, m[0], m[1]);
            gen.mask_pair(pgroup(bh, i0 + g + 8, nt * 8), 2 * t, m[2], m[3]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int i = i0 + g + (e >= 2 ? 8 : 0), j = nt * 8 + 2 * t + (e & 1);
                const bool ok = j < a.Sk && i < a.Sq && (!a.causal || j <= i);
                const float p = ok ? __expf(s[nt][e] * a.scale - lse_s[i]) : 0.0f;
                s[nt][e] = p * (dp[nt][e] * m[e] - D_s[i]) * a.scale;
            }
        }
        float acc[HD / 8][4];
#pragma unroll
        for (int nd = 0; nd < HD / 8; ++nd)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[nd][e] = 0.0f;
#pragma unroll
        for (int kk = 0; kk < NT; ++kk) {
            const uint32_t af[4] = {pack2(s[2 * kk][0], s[2 * kk][1]), pack2(s[2 * kk][2], s[2 * kk][3]),
                                    pack2(s[2 * kk + 1][0], s[2 * kk + 1][1]), pack2(s[2 * kk + 1][2], s[2 * kk + 1][3])};
#pragma unroll
            for (int nd = 0; nd < HD / 8; ++nd) {
                uint32_t b0, b1;
                ldsm_x2_trans(b0, b1, Ks, LD, kk * 16, nd * 8, lane);
                mma16816(acc[nd], af, b0, b1);
            }
        }
        store_rows<HD, NT>(a.dq + (long long)b * a.Sq * a.dq_rs + h * HD, a.dq_rs, i0 + g, a.Sq, t, acc);
    }
}

template <int HD, int NT>
int launch(const AttnArgs& a, bool backward, cudaStream_t st) {
    constexpr int LD = HD + 8, ROWS = NT * 16;
    const size_t smem = backward ? (size_t)4 * ROWS * LD * sizeof(bf16) + 2 * ROWS * sizeof(float)
                                 : (size_t)3 * ROWS * LD * sizeof(bf16);
    if (backward) {
        EEGX_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_kernel<HD, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        eegx::launch(attn_bwd_kernel<HD, NT>, a.B * a.H, 2 * NT * 32, smem, st, a);
    } else {
        EEGX_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_kernel<HD, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        eegx::launch(attn_fwd_kernel<HD, NT>, a.B * a.H, NT * 32, smem, st, a);
    }
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}

template <int HD>
int dispatch_nt(const AttnArgs& a, bool backward, cudaStream_t st) {
    const int smax = a.Sq > a.Sk ? a.Sq : a.Sk;
    if (smax <= 16) return launch<HD, 1>(a, backward, st);
    if (smax <= 32) return launch<HD, 2>(a, backward, st);
    if (smax <= 48) return launch<HD, 3>(a, backward, st);
    return launch<HD, 4>(a, backward, st);
}

int dispatch(const AttnArgs& a, int hd, bool backward, cudaStream_t st) {
    switch (hd) {
        case 64: return dispatch_nt<64>(a, backward, st);
        case 96: return dispatch_nt<96>(a, backward, st);
        case 128: return dispatch_nt<128>(a, backward, st);
        case 192: return dispatch_nt<192>(a, backward, st);
        default: return eegx::set_error(EEGX_ERR_SHAPE, "attention: head_dim %d not in {64, 96, 128, 192}", hd);
    }
}

int check_desc(const eegx_attn_desc* d) {
    EEGX_REQUIRE(d != nullptr, EEGX_ERR_ARG, "attention: NULL descriptor");
    EEGX_REQUIRE(d->B >= 0 && d->H >= 1 && d->Sq >= 1 && d->Sk >= 1 && d->Sq <= 64 && d->Sk <= 64, EEGX_ERR_SHAPE,
                 "attention: this kernel handles 1 <= S_q, S_k <= 64 (got %lld, %lld)", (long long)d->Sq, (long long)d->Sk);
    EEGX_REQUIRE(d->B * d->H < (1LL << 31), EEGX_ERR_SHAPE, "attention: B * H too large");
    EEGX_REQUIRE((d->q_rs % 8) == 0 && (d->k_rs % 8) == 0 && (d->v_rs % 8) == 0 && (d->o_rs % 8) == 0, EEGX_ERR_ALIGN,
                 "attention: row strides must be multiples of 8 elements");
    return EEGX_OK;
}

}  // namespace

extern "C" {

int eegx_attn_fwd_bf16(const eegx_attn_desc* d, const void* q, const void* k, const void* v, void* o, float* lse,
                       const uint64_t* rng_state, uint32_t site, float p, void* stream) {
    if (int rc = eegx::require_sm100()) return rc;
    if (int rc = check_desc(d)) return rc;
    if (d->B == 0) return EEGX_OK;
    EEGX_REQUIRE(q && k && v && o && lse, EEGX_ERR_ARG, "attention fwd: NULL pointer");
    EEGX_REQUIRE(eegx::aligned16(q) && eegx::aligned16(k) && eegx::aligned16(v) && eegx::aligned16(o), EEGX_ERR_ALIGN,
                 "attention fwd: pointers must be 16-byte aligned");
    AttnArgs a{};
    a.q = static_cast<const bf16*>(q); a.k = static_cast<const bf16*>(k); a.v = static_cast<const bf16*>(v);
    a.out = static_cast<bf16*>(o); a.lse = lse;
    a.q_rs = d->q_rs; a.k_rs = d->k_rs; a.v_rs = d->v_rs; a.o_rs = d->o_rs;
    a.B = (int)d->B; a.H = (int)d->H; a.Sq = (int)d->Sq; a.Sk = (int)d->Sk; a.causal = d->causal; a.scale = d->scale;
    a.dc = DropoutCfg{reinterpret_cast<const unsigned long long*>(rng_state), site, p};
    return dispatch(a, (int)d->hd, false, static_cast<cudaStream_t>(stream));
}

int eegx_attn_bwd_bf16(const eegx_attn_desc* d, const void* q, const void* k, const void* v, const void* o,
                       const void* d_o, const float* lse, void* dq, void* dk, void* dv, int64_t dq_rs, int64_t dk_rs,
                       int64_t dv_rs, const uint64_t* rng_state, uint32_t site, float p, void* stream) {
    if (int rc = eegx::require_sm100()) return rc;
    if (int rc = check_desc(d)) return rc;
    if (d->B == 0) return EEGX_OK;
    EEGX_REQUIRE(q && k && v && o && d_o && lse && dq && dk && dv, EEGX_ERR_ARG, "attention bwd: NULL pointer");
    EEGX_REQUIRE(eegx::aligned16(q) && eegx::aligned16(k) && eegx::aligned16(v) && eegx::aligned16(o) &&
                     eegx::aligned16(d_o), EEGX_ERR_ALIGN, "attention bwd: pointers must be 16-byte aligned");
    EEGX_REQUIRE((dq_rs % 2) == 0 && (dk_rs % 2) == 0 && (dv_rs % 2) == 0, EEGX_ERR_ALIGN,
                 "attention bwd: gradient row strides must be even");
    AttnArgs a{};
    a.q = static_cast<const bf16*>(q); a.k = static_cast<const bf16*>(k); a.v = static_cast<const bf16*>(v);
    a.o = static_cast<const bf16*>(o); a.d_o = static_cast<const bf16*>(d_o);
    a.lse = const_cast<float*>(lse);
    a.dq = static_cast<bf16*>(dq); a.dk = static_cast<bf16*>(dk); a.dv = static_cast<bf16*>(dv);
    a.q_rs = d->q_rs; a.k_rs = d->k_rs; a.v_rs = d->v_rs; a.o_rs = d->o_rs;
    a.dq_rs = dq_rs; a.dk_rs = dk_rs; a.dv_rs = dv_rs;
    a.B = (int)d->B; a.H = (int)d->H; a.Sq = (int)d->Sq; a.Sk = (int)d->Sk; a.causal = d->causal; a.scale = d->scale;
    a.dc = DropoutCfg{reinterpret_cast<const unsigned long long*>(rng_state), site, p};
    return dispatch(a, (int)d->hd, true, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
