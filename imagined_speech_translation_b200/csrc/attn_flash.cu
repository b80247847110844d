// Fused multi-head attention core for LONG sequences (any S_q, S_k; head_dim in {64, 96, 128, 192}), forward
// and backward, flash style: the S_q x S_k scores and probabilities never exist in HBM.
//
// Replaces, inside nn.MultiheadAttention at the reference's actual encoder shapes (main_model/src/models/
// layers.py:230-251 with S = T + 4 = 1655 on the real data, 2052 / 4100 on the raw 2048 / 4096-sample trials) the
// chain  q k^T / sqrt(hd) -> softmax -> dropout -> (.) v  and its autograd graph.  attn_small.cu keeps the
// S <= 64 case (one CTA per (batch, head)).
//
//   forward   one CTA = 64 queries of one (batch, head), 4 warps x 16 rows; keys / values stream through shared
//             memory in tiles of 64 (cp.async); online softmax in the log2 domain (running max m, running sum l,
//             output accumulator rescaled when m moves); dropout multiplies the probabilities that enter P V but
//             not l (dropout(softmax(.)) semantics).  Saves lse = m + log l per row.
//   backward  recomputes P from lse; two kernels so that every gradient has one owner and a fixed summation order
//             (bit-reproducible, no atomics):
//               dQ     one CTA = 64 queries, loops over the key tiles; also writes D_i = dO_i . O_i to `dsum`
//               dK,dV  one CTA = 64 keys, loops over the query tiles; warps 0-3 accumulate dV, warps 4-7 dK
//                      (each holds one HD-wide accumulator: at hd = 192 two would not fit the register file)
//   matmuls   mma.sync m16n8k16 bf16, fp32 accumulate (the warp-level tensor-core path; a tcgen05 version needs
//             S, dP, dK, dV accumulators in TMEM at once -- 640 columns at hd = 192 -- and is future work).
//
// q, k, v, o (and their gradients) are rows of (B*S, row_stride) matrices with head h at columns
// [h*hd, (h+1)*hd): the packed QKV projection is consumed, and the packed dQKV produced, in place.
#include "eegx_common.h"
#include "fused_common.cuh"

namespace {

using namespace eegx;
typedef __nv_bfloat16 bf16;

constexpr int BM = 64;             // rows of the resident tile (queries in fwd / dQ, keys in dK,dV)
constexpr int BN = 64;             // rows of the streamed tile
constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;

struct FlashArgs {
    const bf16 *q, *k, *v, *o, *d_o;
    bf16 *out, *dq, *dk, *dv;
    float *lse, *dsum;          // (B, H, Sq)
    long long q_rs, k_rs, v_rs, o_rs, dq_rs, dk_rs, dv_rs;
    int B, H, Sq, Sk, nkg;      // nkg = ceil(Sk / 8): dropout groups per query row
    float scale, scale_log2;
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
__device__ __forceinline__ void ldsm_x2_trans(uint32_t& b0, uint32_t& b1, const bf16* tile, int ld, int row0,
                                              int col0, int lane) {
    const unsigned addr = (unsigned)__cvta_generic_to_shared(tile + (row0 + (lane & 15)) * ld + col0);
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n" : "=r"(b0), "=r"(b1) : "r"(addr));
}
// rows [row0, row0 + 64) of a (S x HD, row stride rs) matrix -> shared tile (64 x (HD + 8)), rows >= S zero filled
template <int HD>
__device__ __forceinline__ void load_tile(bf16* dst, const bf16* src, long long rs, int row0, int S) {
    constexpr int V = HD / 8, LD = HD + 8;
    for (int idx = threadIdx.x; idx < 64 * V; idx += blockDim.x) {
        const int r = idx / V, v = idx % V;
        const unsigned d = (unsigned)__cvta_generic_to_shared(dst + r * LD + v * 8);
        const int gr = row0 + r;
        const bf16* g = src + (long long)(gr < S ? gr : 0) * rs + v * 8;
        const int bytes = gr < S ? 16 : 0;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(g), "r"(bytes) : "memory");
    }
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int LD>
__device__ __forceinline__ void load_a(uint32_t (&a)[4], const bf16* tile, int row0, int ks, int g, int t) {
    const bf16* p = tile + (row0 + g) * LD + ks * 16 + 2 * t;
    a[0] = lds32(p);
    a[1] = lds32(p + 8 * LD);
    a[2] = lds32(p + 8);
    a[3] = lds32(p + 8 * LD + 8);
}

// dropout group of probability (i, j) of (batch*head) bh: 8 consecutive keys of one query row
__device__ __forceinline__ unsigned long long pgroup(const FlashArgs& a, int bh, int i, int j) {
    return ((unsigned long long)bh * (unsigned long long)a.Sq + (unsigned long long)i) * (unsigned long long)a.nkg +
           (unsigned long long)(j >> 3);
}

template <int HD>
__device__ __forceinline__ void store_rows(bf16* base, long long rs, int row_lo, int S, int t, const float (&acc)[HD / 8][4],
                                           float f_lo, float f_hi) {
    bf16* lo = base + (long long)row_lo * rs + 2 * t;
    bf16* hi = lo + 8 * rs;
#pragma unroll
    for (int nd = 0; nd < HD / 8; ++nd) {
        if (row_lo < S) *reinterpret_cast<uint32_t*>(lo + nd * 8) = pack2(acc[nd][0] * f_lo, acc[nd][1] * f_lo);
        if (row_lo + 8 < S) *reinterpret_cast<uint32_t*>(hi + nd * 8) = pack2(acc[nd][2] * f_hi, acc[nd][3] * f_hi);
    }
}

// ------------------------------------------------------------------------------------------ forward
template <int HD>
__global__ void __launch_bounds__(128) flash_fwd_kernel(const FlashArgs a) {
    EEGX_PDL_SYNC();
    constexpr int LD = HD + 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    bf16* Qs = reinterpret_cast<bf16*>(smem_raw);
    bf16* Ks = Qs + BM * LD;
    bf16* Vs = Ks + BN * LD;
    const int q0 = blockIdx.x * BM, bh = blockIdx.y, b = bh / a.H, h = bh % a.H;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const bf16* kbase = a.k + (long long)b * a.Sk * a.k_rs + h * HD;
    const bf16* vbase = a.v + (long long)b * a.Sk * a.v_rs + h * HD;
    load_tile<HD>(Qs, a.q + (long long)b * a.Sq * a.q_rs + h * HD, a.q_rs, q0, a.Sq);

    const int r0 = warp * 16;
    const int i_lo = q0 + r0 + g, i_hi = i_lo + 8;
    float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.0f, l_hi = 0.0f;
    float acc[HD / 8][4];
#pragma unroll
    for (int nd = 0; nd < HD / 8; ++nd)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[nd][e] = 0.0f;
    const DropoutGen gen(a.dc);

    for (int j0 = 0; j0 < a.Sk; j0 += BN) {
        __syncthreads();                                  // the previous K / V tile is consumed
        load_tile<HD>(Ks, kbase, a.k_rs, j0, a.Sk);
        load_tile<HD>(Vs, vbase, a.v_rs, j0, a.Sk);
        cp_async_wait_all();
        __syncthreads();
        float s[BN / 8][4];
#pragma unroll
        for (int nt = 0; nt < BN / 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) s[nt][e] = 0.0f;
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) {
            uint32_t af[4];
            load_a<LD>(af, Qs, r0, ks, g, t);
#pragma unroll
            for (int nt = 0; nt < BN / 8; ++nt) {
                const bf16* p = Ks + (nt * 8 + g) * LD + ks * 16 + 2 * t;
                mma16816(s[nt], af, lds32(p), lds32(p + 8));
            }
        }
        float mx_lo = -INFINITY, mx_hi = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < BN / 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = j0 + nt * 8 + 2 * t + (e & 1);
                s[nt][e] = j < a.Sk ? s[nt][e] * a.scale_log2 : -INFINITY;
                if (e < 2) mx_lo = fmaxf(mx_lo, s[nt][e]); else mx_hi = fmaxf(mx_hi, s[nt][e]);
            }
        mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
        mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
        mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
        mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
        const float mn_lo = fmaxf(m_lo, mx_lo), mn_hi = fmaxf(m_hi, mx_hi);   // finite: every tile holds a valid key
        const float c_lo = exp2f(m_lo - mn_lo), c_hi = exp2f(m_hi - mn_hi);
        m_lo = mn_lo; m_hi = mn_hi;
        l_lo *= c_lo; l_hi *= c_hi;
#pragma unroll
        for (int nd = 0; nd < HD / 8; ++nd) {
            acc[nd][0] *= c_lo; acc[nd][1] *= c_lo;
            acc[nd][2] *= c_hi; acc[nd][3] *= c_hi;
        }
#pragma unroll
        for (int nt = 0; nt < BN / 8; ++nt) {
            const float p0 = exp2f(s[nt][0] - m_lo), p1 = exp2f(s[nt][1] - m_lo);
            const float p2 = exp2f(s[nt][2] - m_hi), p3 = exp2f(s[nt][3] - m_hi);
            l_lo += p0 + p1;
            l_hi += p2 + p3;
            float k0, k1, k2, k3;
            gen.mask_pair(pgroup(a, bh, i_lo, j0 + nt * 8), 2 * t, k0, k1);
            gen.mask_pair(pgroup(a, bh, i_hi, j0 + nt * 8), 2 * t, k2, k3);
            s[nt][0] = p0 * k0; s[nt][1] = p1 * k1; s[nt][2] = p2 * k2; s[nt][3] = p3 * k3;
        }
#pragma unroll
        for (int kk = 0; kk < BN / 16; ++kk) {
            const uint32_t af[4] = {pack2(s[2 * kk][0], s[2 * kk][1]), pack2(s[2 * kk][2], s[2 * kk][3]),
                                    pack2(s[2 * kk + 1][0], s[2 * kk + 1][1]), pack2(s[2 * kk + 1][2], s[2 * kk + 1][3])};
#pragma unroll
            for (int nd = 0; nd < HD / 8; ++nd) {
                uint32_t b0, b1;
                ldsm_x2_trans(b0, b1, Vs, LD, kk * 16, nd * 8, lane);
                mma16816(acc[nd], af, b0, b1);
            }
        }
    }
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
    if (t == 0) {
        if (i_lo < a.Sq) a.lse[(long long)bh * a.Sq + i_lo] = (m_lo + log2f(l_lo)) * LN2;
        if (i_hi < a.Sq) a.lse[(long long)bh * a.Sq + i_hi] = (m_hi + log2f(l_hi)) * LN2;
    }
    store_rows<HD>(a.out + (long long)b * a.Sq * a.o_rs + h * HD, a.o_rs, i_lo, a.Sq, t, acc, 1.0f / l_lo, 1.0f / l_hi);
}

// ------------------------------------------------------------------------------------------ backward: dQ (+ D)
template <int HD>
__global__ void __launch_bounds__(128) flash_bwd_dq_kernel(const FlashArgs a) {
    EEGX_PDL_SYNC();
    constexpr int LD = HD + 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    bf16* Qs = reinterpret_cast<bf16*>(smem_raw);
    bf16* dOs = Qs + BM * LD;
    bf16* Ks = dOs + BM * LD;
    bf16* Vs = Ks + BN * LD;
    float* lse_s = reinterpret_cast<float*>(Vs + BN * LD);
    float* D_s = lse_s + BM;
    const int q0 = blockIdx.x * BM, bh = blockIdx.y, b = bh / a.H, h = bh % a.H;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const bf16* kbase = a.k + (long long)b * a.Sk * a.k_rs + h * HD;
    const bf16* vbase = a.v + (long long)b * a.Sk * a.v_rs + h * HD;
    load_tile<HD>(Qs, a.q + (long long)b * a.Sq * a.q_rs + h * HD, a.q_rs, q0, a.Sq);
    load_tile<HD>(dOs, a.d_o + (long long)b * a.Sq * a.o_rs + h * HD, a.o_rs, q0, a.Sq);
    for (int i = threadIdx.x; i < BM; i += blockDim.x)
        lse_s[i] = q0 + i < a.Sq ? a.lse[(long long)bh * a.Sq + q0 + i] * LOG2E : 0.0f;
    cp_async_wait_all();
    __syncthreads();
    for (int i = warp; i < BM; i += 4) {                  // D_i = dO_i . O_i
        float d = 0.0f;
        if (q0 + i < a.Sq && lane < HD / 8) {
            float dv[8], ov[8];
            load8(dOs + i * LD + lane * 8, dv);
            load8(a.o + ((long long)b * a.Sq + q0 + i) * a.o_rs + h * HD + lane * 8, ov);
#pragma unroll
            for (int e = 0; e < 8; ++e) d = fmaf(dv[e], ov[e], d);
        }
        d = warp_sum_f(d);
        if (lane == 0) {
            D_s[i] = d;
            if (q0 + i < a.Sq) a.dsum[(long long)bh * a.Sq + q0 + i] = d;
        }
    }
    const int r0 = warp * 16;
    const int i_lo = q0 + r0 + g, i_hi = i_lo + 8;
    float acc[HD / 8][4];
#pragma unroll
    for (int nd = 0; nd < HD / 8; ++nd)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[nd][e] = 0.0f;
    const DropoutGen gen(a.dc);

    for (int j0 = 0; j0 < a.Sk; j0 += BN) {
        __syncthreads();                                  // previous tile consumed; D_s visible on the first pass
        load_tile<HD>(Ks, kbase, a.k_rs, j0, a.Sk);
        load_tile<HD>(Vs, vbase, a.v_rs, j0, a.Sk);
        cp_async_wait_all();
        __syncthreads();
        float s[BN / 8][4], dp[BN / 8][4];
#pragma unroll
        for (int nt = 0; nt < BN / 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) s[nt][e] = dp[nt][e] = 0.0f;
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) {
            uint32_t aq[4], ad[4];
            load_a<LD>(aq, Qs, r0, ks, g, t);
            load_a<LD>(ad, dOs, r0, ks, g, t);
#pragma unroll
            for (int nt = 0; nt < BN / 8; ++nt) {
                const bf16* pk = Ks + (nt * 8 + g) * LD + ks * 16 + 2 * t;
                const bf16* pv = Vs + (nt * 8 + g) * LD + ks * 16 + 2 * t;
                mma16816(s[nt], aq, lds32(pk), lds32(pk + 8));
                mma16816(dp[nt], ad, lds32(pv), lds32(pv + 8));
            }
        }
        const float ls_lo = lse_s[r0 + g], ls_hi = lse_s[r0 + g + 8], D_lo = D_s[r0 + g], D_hi = D_s[r0 + g + 8];
#pragma unroll
        for (int nt = 0; nt < BN / 8; ++nt) {
            float k[4];
            gen.mask_pair(pgroup(a, bh, i_lo, j0 + nt * 8), 2 * t, k[0], k[1]);
            gen.mask_pair(pgroup(a, bh, i_hi, j0 + nt * 8), 2 * t, k[2], k[3]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = j0 + nt * 8 + 2 * t + (e & 1);
                const float p = j < a.Sk ? exp2f(fmaf(s[nt][e], a.scale_log2, -(e < 2 ? ls_lo : ls_hi))) : 0.0f;
                s[nt][e] = p * (dp[nt][e] * k[e] - (e < 2 ? D_lo : D_hi)) * a.scale;
            }
        }
#pragma unroll
        for (int kk = 0; kk < BN / 16; ++kk) {
            const uint32_t af[4] = {pack2(s[2 * kk][0], s[2 * kk][1]), pack2(s[2 * kk][2], s[2 * kk][3]),
                                    pack2(s[2 * kk + 1][0], s[2 * kk + 1][1]), pack2(s[2 * kk + 1][2], s[2 * kk + 1][3])};
#pragma unroll
            for (int nd = 0; nd < HD / 8; ++nd) {
                uint32_t b0, b1;
                ldsm_x2_trans(b0, b1, Ks, LD, kk * 16, nd * 8, lane);
                mma16816(acc[nd], af, b0, b1);
            }
        }
    }
    store_rows<HD>(a.dq + (long long)b * a.Sq * a.dq_rs + h * HD, a.dq_rs, i_lo, a.Sq, t, acc, 1.0f, 1.0f);
}

// ------------------------------------------------------------------------------------------ backward: dK, dV
template <int HD>
__global__ void __launch_bounds__(256) flash_bwd_dkv_kernel(const FlashArgs a) {
    EEGX_PDL_SYNC();
    constexpr int LD = HD + 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    bf16* Ks = reinterpret_cast<bf16*>(smem_raw);
    bf16* Vs = Ks + BM * LD;
    bf16* Qs = Vs + BM * LD;
    bf16* dOs = Qs + BN * LD;
    float* lse_s = reinterpret_cast<float*>(dOs + BN * LD);
    float* D_s = lse_s + BN;
    const int k0 = blockIdx.x * BM, bh = blockIdx.y, b = bh / a.H, h = bh % a.H;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const bf16* qbase = a.q + (long long)b * a.Sq * a.q_rs + h * HD;
    const bf16* dobase = a.d_o + (long long)b * a.Sq * a.o_rs + h * HD;
    load_tile<HD>(Ks, a.k + (long long)b * a.Sk * a.k_rs + h * HD, a.k_rs, k0, a.Sk);
    load_tile<HD>(Vs, a.v + (long long)b * a.Sk * a.v_rs + h * HD, a.v_rs, k0, a.Sk);

    const bool is_dk = warp >= 4;
    const int r0 = (warp & 3) * 16;                       // this warp's 16 keys inside the tile
    const int j_lo = k0 + r0 + g, j_hi = j_lo + 8;
    float acc[HD / 8][4];
#pragma unroll
    for (int nd = 0; nd < HD / 8; ++nd)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[nd][e] = 0.0f;
    const DropoutGen gen(a.dc);

    for (int i0 = 0; i0 < a.Sq; i0 += BN) {
        __syncthreads();
        load_tile<HD>(Qs, qbase, a.q_rs, i0, a.Sq);
        load_tile<HD>(dOs, dobase, a.o_rs, i0, a.Sq);
        for (int i = threadIdx.x; i < BN; i += blockDim.x) {
            const bool ok = i0 + i < a.Sq;
            lse_s[i] = ok ? a.lse[(long long)bh * a.Sq + i0 + i] * LOG2E : 0.0f;
            D_s[i] = ok ? a.dsum[(long long)bh * a.Sq + i0 + i] : 0.0f;
        }
        cp_async_wait_all();
        __syncthreads();
        // transposed scores: rows = this warp's keys, columns = the 64 queries of the tile
        float sT[BN / 8][4], dpT[BN / 8][4];
#pragma unroll
        for (int nt = 0; nt < BN / 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) sT[nt][e] = dpT[nt][e] = 0.0f;
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) {
            uint32_t ak[4], av[4];
            load_a<LD>(ak, Ks, r0, ks, g, t);
            if (is_dk) load_a<LD>(av, Vs, r0, ks, g, t);
#pragma unroll
            for (int nt = 0; nt < BN / 8; ++nt) {
                const bf16* pq = Qs + (nt * 8 + g) * LD + ks * 16 + 2 * t;
                mma16816(sT[nt], ak, lds32(pq), lds32(pq + 8));
                if (is_dk) {
                    const bf16* pd = dOs + (nt * 8 + g) * LD + ks * 16 + 2 * t;
                    mma16816(dpT[nt], av, lds32(pd), lds32(pd + 8));
                }
            }
        }
#pragma unroll
        for (int nt = 0; nt < BN / 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = e < 2 ? j_lo : j_hi, il = nt * 8 + 2 * t + (e & 1), i = i0 + il;
                const bool ok = j < a.Sk && i < a.Sq;
                const float p = ok ? exp2f(fmaf(sT[nt][e], a.scale_log2, -lse_s[il])) : 0.0f;
                const float m = gen.mask_one(pgroup(a, bh, i, j), j & 7);
                sT[nt][e] = is_dk ? p * (dpT[nt][e] * m - D_s[il]) * a.scale      // dS^T
                                  : p * m;                                       // dropped probabilities
            }
        const bf16* Bt = is_dk ? Qs : dOs;                // dK = dS^T Q,  dV = Pd^T dO
#pragma unroll
        for (int kk = 0; kk < BN / 16; ++kk) {
            const uint32_t af[4] = {pack2(sT[2 * kk][0], sT[2 * kk][1]), pack2(sT[2 * kk][2], sT[2 * kk][3]),
                                    pack2(sT[2 * kk + 1][0], sT[2 * kk + 1][1]), pack2(sT[2 * kk + 1][2], sT[2 * kk + 1][3])};
#pragma unroll
            for (int nd = 0; nd < HD / 8; ++nd) {
                uint32_t b0, b1;
                ldsm_x2_trans(b0, b1, Bt, LD, kk * 16, nd * 8, lane);
                mma16816(acc[nd], af, b0, b1);
            }
        }
    }
    if (is_dk) store_rows<HD>(a.dk + (long long)b * a.Sk * a.dk_rs + h * HD, a.dk_rs, j_lo, a.Sk, t, acc, 1.0f, 1.0f);
    else store_rows<HD>(a.dv + (long long)b * a.Sk * a.dv_rs + h * HD, a.dv_rs, j_lo, a.Sk, t, acc, 1.0f, 1.0f);
}

template <int HD>
int launch_fwd(const FlashArgs& a, cudaStream_t st) {
    constexpr int LD = HD + 8;
    const size_t smem = (size_t)(BM + 2 * BN) * LD * sizeof(bf16);
    EEGX_CUDA_CHECK(cudaFuncSetAttribute(flash_fwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    eegx::launch(flash_fwd_kernel<HD>, dim3((a.Sq + BM - 1) / BM, a.B * a.H), 128, smem, st, a);
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}

template <int HD>
int launch_bwd(const FlashArgs& a, cudaStream_t st) {
    constexpr int LD = HD + 8;
    const size_t smem = (size_t)(2 * BM + 2 * BN) * LD * sizeof(bf16) + 2 * 64 * sizeof(float);
    EEGX_CUDA_CHECK(cudaFuncSetAttribute(flash_bwd_dq_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    EEGX_CUDA_CHECK(cudaFuncSetAttribute(flash_bwd_dkv_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    eegx::launch(flash_bwd_dq_kernel<HD>, dim3((a.Sq + BM - 1) / BM, a.B * a.H), 128, smem, st, a);   // writes dsum
    EEGX_CUDA_CHECK(cudaGetLastError());
    eegx::launch(flash_bwd_dkv_kernel<HD>, dim3((a.Sk + BM - 1) / BM, a.B * a.H), 256, smem, st, a);  // reads dsum
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}

int check_desc(const eegx_attn_desc* d) {
    EEGX_REQUIRE(d != nullptr, EEGX_ERR_ARG, "flash attention: NULL descriptor");
    EEGX_REQUIRE(d->B >= 0 && d->H >= 1 && d->Sq >= 1 && d->Sk >= 1, EEGX_ERR_SHAPE, "flash attention: bad sizes");
    EEGX_REQUIRE(d->Sq < (1LL << 24) && d->Sk < (1LL << 24) && d->B * d->H <= 65535, EEGX_ERR_SHAPE,
                 "flash attention: B * H must be <= 65535 (grid.y) and S < 2^24");
    EEGX_REQUIRE(!d->causal, EEGX_ERR_ARG, "flash attention: causal masking is not implemented (the encoder has none)");
    EEGX_REQUIRE(d->hd == 64 || d->hd == 96 || d->hd == 128 || d->hd == 192, EEGX_ERR_SHAPE,
                 "flash attention: head_dim %lld not in {64, 96, 128, 192}", (long long)d->hd);
    EEGX_REQUIRE((d->q_rs % 8) == 0 && (d->k_rs % 8) == 0 && (d->v_rs % 8) == 0 && (d->o_rs % 8) == 0, EEGX_ERR_ALIGN,
                 "flash attention: row strides must be multiples of 8 elements");
    return EEGX_OK;
}

FlashArgs base_args(const eegx_attn_desc* d, const uint64_t* rng_state, uint32_t site, float p) {
    FlashArgs a{};
    a.q_rs = d->q_rs; a.k_rs = d->k_rs; a.v_rs = d->v_rs; a.o_rs = d->o_rs;
    a.B = (int)d->B; a.H = (int)d->H; a.Sq = (int)d->Sq; a.Sk = (int)d->Sk;
    a.nkg = (int)((d->Sk + 7) / 8);
    a.scale = d->scale;
    a.scale_log2 = d->scale * LOG2E;
    a.dc = DropoutCfg{reinterpret_cast<const unsigned long long*>(rng_state), site, p};
    return a;
}

}  // namespace

extern "C" {

int eegx_attn_flash_fwd_bf16(const eegx_attn_desc* d, const void* q, const void* k, const void* v, void* o, float* lse,
                             const uint64_t* rng_state, uint32_t site, float p, void* stream) {
    if (int rc = eegx::require_sm100()) return rc;
    if (int rc = check_desc(d)) return rc;
    if (d->B == 0) return EEGX_OK;
    EEGX_REQUIRE(q && k && v && o && lse, EEGX_ERR_ARG, "flash attention fwd: NULL pointer");
    EEGX_REQUIRE(eegx::aligned16(q) && eegx::aligned16(k) && eegx::aligned16(v) && eegx::aligned16(o), EEGX_ERR_ALIGN,
                 "flash attention fwd: pointers must be 16-byte aligned");
    FlashArgs a = base_args(d, rng_state, site, p);
    a.q = static_cast<const bf16*>(q); a.k = static_cast<const bf16*>(k); a.v = static_cast<const bf16*>(v);
    a.out = static_cast<bf16*>(o); a.lse = lse;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (d->hd) {
        case 64: return launch_fwd<64>(a, st);
        case 96: return launch_fwd<96>(a, st);
        case 128: return launch_fwd<128>(a, st);
        default: return launch_fwd<192>(a, st);
    }
}

int eegx_attn_flash_bwd_bf16(const eegx_attn_desc* d, const void* q, const void* k, const void* v, const void* o,
                             const void* d_o, const float* lse, float* dsum, void* dq, void* dk, void* dv, int64_t dq_rs,
                             int64_t dk_rs, int64_t dv_rs, const uint64_t* rng_state, uint32_t site, float p,
                             void* stream) {
    if (int rc = eegx::require_sm100()) return rc;
    if (int rc = check_desc(d)) return rc;
    if (d->B == 0) return EEGX_OK;
    EEGX_REQUIRE(q && k && v && o && d_o && lse && dsum && dq && dk && dv, EEGX_ERR_ARG, "flash attention bwd: NULL pointer");
    EEGX_REQUIRE(eegx::aligned16(q) && eegx::aligned16(k) && eegx::aligned16(v) && eegx::aligned16(o) &&
                     eegx::aligned16(d_o), EEGX_ERR_ALIGN, "flash attention bwd: pointers must be 16-byte aligned");
    EEGX_REQUIRE((dq_rs % 2) == 0 && (dk_rs % 2) == 0 && (dv_rs % 2) == 0, EEGX_ERR_ALIGN,
                 "flash attention bwd: gradient row strides must be even");
    FlashArgs a = base_args(d, rng_state, site, p);
    a.q = static_cast<const bf16*>(q); a.k = static_cast<const bf16*>(k); a.v = static_cast<const bf16*>(v);
    a.o = static_cast<const bf16*>(o); a.d_o = static_cast<const bf16*>(d_o);
    a.lse = const_cast<float*>(lse); a.dsum = dsum;
    a.dq = static_cast<bf16*>(dq); a.dk = static_cast<bf16*>(dk); a.dv = static_cast<bf16*>(dv);
    a.dq_rs = dq_rs; a.dk_rs = dk_rs; a.dv_rs = dv_rs;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (d->hd) {
        case 64: return launch_bwd<64>(a, st);
        case 96: return launch_bwd<96>(a, st);
        case 128: return launch_bwd<128>(a, st);
        default: return launch_bwd<192>(a, st);
    }
}

}  // extern "C"
