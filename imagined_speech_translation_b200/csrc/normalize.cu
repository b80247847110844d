// Reference-actual normalisation on the GPU (HBM-bound, one read + one write).
//
// Replaces, for a whole batch, what EEGDataset does per trial on the host:
//   _process_raw_eeg       main_model/src/data/dataset.py:172-191  (nan_to_num)
//   _normalize_eeg_sample  main_model/src/data/dataset.py:193-225  (region gather,
//                          RobustScaler.transform == (x - center) / scale, and the
//                          per-channel z-score fallback at :213-216)
// Algorithmic bytes per trial: 4*C_out*T read (only the gathered rows) + 4*C_out*T
// written; the index / center / scale vectors are O(C) and stay in L1/L2.
#include "eegx_common.h"

namespace {

__device__ __forceinline__ float clean(float v) {
    // np.nan_to_num(nan=0.0, posinf=10.0, neginf=-10.0)  (dataset.py:185)
    if (v != v) return 0.0f;
    if (v == __int_as_float(0x7f800000)) return 10.0f;
    if (v == __int_as_float(0xff800000)) return -10.0f;
    return v;
}

__device__ __forceinline__ float4 ld_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

// One thread handles UNROLL float4s of one output row per iteration; rows are
// (trial b, output channel j).  T4 = T / 4.
template <int UNROLL>
__global__ void __launch_bounds__(256)
normalize_vec4_kernel(const float* __restrict__ x, const int32_t* __restrict__ ch_idx,
                      const float* __restrict__ center, const float* __restrict__ scale,
                      float* __restrict__ out, const int64_t* __restrict__ out_off,
                      const int64_t* __restrict__ out_bstride, int64_t rows, int C_in, int C_out,
                      int T4) {
    EEGX_PDL_SYNC();
    const int chunks_per_row = (T4 + 256 * UNROLL - 1) / (256 * UNROLL);
    const int64_t total_chunks = rows * chunks_per_row;
    for (int64_t chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
        const int64_t row = chunk / chunks_per_row;
        const int part = (int)(chunk - row * chunks_per_row);
        const int64_t b = row / C_out;
        const int j = (int)(row - b * C_out);
        const int src_c = ch_idx ? ch_idx[j] : j;
        const float c = center ? center[j] : 0.0f;
        const float s = scale ? scale[j] : 1.0f;
        const float4* src = reinterpret_cast<const float4*>(x + (b * C_in + src_c) * (int64_t)T4 * 4);
        float* dst_row = out_off ? out + out_off[j] + b * out_bstride[j]
                                 : out + row * (int64_t)T4 * 4;
        float4* dst = reinterpret_cast<float4*>(dst_row);
        const int base = part * 256 * UNROLL + threadIdx.x;
        float4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int i = base + u * 256;
            if (i < T4) v[u] = ld_stream(src + i);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int i = base + u * 256;
            if (i < T4) {
                float4 r;
                r.x = (clean(v[u].x) - c) / s;
                r.y = (clean(v[u].y) - c) / s;
                r.z = (clean(v[u].z) - c) / s;
                r.w = (clean(v[u].w) - c) / s;
                __stcs(dst + i, r);
            }
        }
    }
}

// Scalar variant for T % 4 != 0 or unaligned rows (the real data has T = 1651).
__global__ void __launch_bounds__(256)
normalize_scalar_kernel(const float* __restrict__ x, const int32_t* __restrict__ ch_idx,
                        const float* __restrict__ center, const float* __restrict__ scale,
                        float* __restrict__ out, const int64_t* __restrict__ out_off,
                        const int64_t* __restrict__ out_bstride, int64_t rows, int C_in, int C_out,
                        int T) {
    EEGX_PDL_SYNC();
    for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
        const int64_t b = row / C_out;
        const int j = (int)(row - b * C_out);
        const int src_c = ch_idx ? ch_idx[j] : j;
        const float c = center ? center[j] : 0.0f;
        const float s = scale ? scale[j] : 1.0f;
        const float* src = x + (b * C_in + src_c) * (int64_t)T;
        float* dst = out_off ? out + out_off[j] + b * out_bstride[j] : out + row * (int64_t)T;
        for (int t = threadIdx.x; t < T; t += blockDim.x) dst[t] = (clean(__ldg(src + t)) - c) / s;
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Fixed-order block reduction (no float atomics: bit-stable run to run).
__device__ float block_sum(float v, float* red /* >= 33 floats */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        float t = lane < nwarps ? red[lane] : 0.0f;
        t = warp_sum(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

// One CTA per (trial, output channel) row; the row is staged in shared memory so
// HBM sees one read and one write.  (x - mean_t) / (std_t + 1e-8), population std.
__global__ void __launch_bounds__(256)
zscore_time_kernel(const float* __restrict__ x, const int32_t* __restrict__ ch_idx,
                   float* __restrict__ out, const int64_t* __restrict__ out_off,
                   const int64_t* __restrict__ out_bstride, int64_t rows, int C_in, int C_out,
                   int T) {
    EEGX_PDL_SYNC();
    extern __shared__ float row_s[];
    __shared__ float red[33];
    for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
        const int64_t b = row / C_out;
        const int j = (int)(row - b * C_out);
        const int src_c = ch_idx ? ch_idx[j] : j;
        const float* src = x + (b * C_in + src_c) * (int64_t)T;
        float* dst = out_off ? out + out_off[j] + b * out_bstride[j] : out + row * (int64_t)T;
        float acc = 0.0f;
        for (int t = threadIdx.x; t < T; t += blockDim.x) {
            const float v = clean(__ldg(src + t));
            row_s[t] = v;
            acc += v;
        }
        const float mean = block_sum(acc, red) / (float)T;
        float dev = 0.0f;
        for (int t = threadIdx.x; t < T; t += blockDim.x) {
            const float d = row_s[t] - mean;
            dev = fmaf(d, d, dev);
        }
        const float var = block_sum(dev, red) / (float)T;
        const float denom = sqrtf(var) + 1e-8f;
        for (int t = threadIdx.x; t < T; t += blockDim.x) dst[t] = (row_s[t] - mean) / denom;
        __syncthreads();
    }
}

int check_common(const float* x, float* out, int64_t B, int64_t C_in, int64_t C_out, int64_t T) {
    EEGX_REQUIRE(B == 0 || (x && out), EEGX_ERR_ARG, "x/out must not be NULL");
    EEGX_REQUIRE(B >= 0 && C_in > 0 && C_out > 0 && T > 0, EEGX_ERR_SHAPE,
                 "bad sizes B=%lld C_in=%lld C_out=%lld T=%lld", (long long)B, (long long)C_in,
                 (long long)C_out, (long long)T);
    EEGX_REQUIRE(C_in < (1 << 30) && C_out < (1 << 30) && T < (1 << 30), EEGX_ERR_SHAPE,
                 "sizes too large");
    return EEGX_OK;
}

}  // namespace

extern "C" int eegx_normalize_f32(const float* x, const int32_t* ch_idx, const float* center,
                                  const float* scale, float* out, const int64_t* out_off,
                                  const int64_t* out_bstride, int64_t B, int64_t C_in,
                                  int64_t C_out, int64_t T, void* stream) {
    if (int rc = eegx::require_sm100()) return rc;
    if (int rc = check_common(x, out, B, C_in, C_out, T)) return rc;
    EEGX_REQUIRE((center == nullptr) == (scale == nullptr), EEGX_ERR_ARG,
                 "center and scale must be both given or both NULL");
    EEGX_REQUIRE((out_off == nullptr) == (out_bstride == nullptr), EEGX_ERR_ARG,
                 "out_off and out_bstride must be both given or both NULL");
    EEGX_REQUIRE(ch_idx != nullptr || C_in == C_out, EEGX_ERR_ARG,
                 "ch_idx == NULL needs C_in == C_out");
    if (B == 0) return EEGX_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t rows = B * C_out;
    // vec4 path: every row start must be 16-byte aligned.  With out_off the caller
    // promises nothing, so only the dense layout takes it.
    const bool vec = (T % 4 == 0) && eegx::aligned16(x) && eegx::aligned16(out) && out_off == nullptr;
    if (vec) {
        constexpr int UNROLL = 2;
        const int T4 = (int)(T / 4);
        const int chunks_per_row = (T4 + 256 * UNROLL - 1) / (256 * UNROLL);
        const int64_t total = rows * chunks_per_row;
        const int grid = (int)(total < (int64_t)eegx::kNumSMsB200 * 32 ? total
                                                                        : (int64_t)eegx::kNumSMsB200 * 32);
        eegx::launch(normalize_vec4_kernel<UNROLL>, grid, 256, 0, st, x, ch_idx, center, scale, out, out_off,
                                                            out_bstride, rows, (int)C_in, (int)C_out,
                                                            T4);
    } else {
        const int grid = (int)(rows < (int64_t)eegx::kNumSMsB200 * 16 ? rows
                                                                       : (int64_t)eegx::kNumSMsB200 * 16);
        eegx::launch(normalize_scalar_kernel, grid, 256, 0, st, x, ch_idx, center, scale, out, out_off,
                                                      out_bstride, rows, (int)C_in, (int)C_out,
                                                      (int)T);
    }
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}

extern "C" int eegx_zscore_time_f32(const float* x, const int32_t* ch_idx, float* out,
                                    const int64_t* out_off, const int64_t* out_bstride, int64_t B,
                                    int64_t C_in, int64_t C_out, int64_t T, void* stream) {
    if (int rc = eegx::require_sm100()) return rc;
    if (int rc = check_common(x, out, B, C_in, C_out, T)) return rc;
    EEGX_REQUIRE((out_off == nullptr) == (out_bstride == nullptr), EEGX_ERR_ARG,
                 "out_off and out_bstride must be both given or both NULL");
    EEGX_REQUIRE(ch_idx != nullptr || C_in == C_out, EEGX_ERR_ARG,
                 "ch_idx == NULL needs C_in == C_out");
    EEGX_REQUIRE(T * 4 <= 200 * 1024, EEGX_ERR_SHAPE, "T=%lld too long for the row-in-smem kernel",
                 (long long)T);
    if (B == 0) return EEGX_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t rows = B * C_out;
    const size_t smem = (size_t)T * sizeof(float);
    if (smem > 48 * 1024)
        EEGX_CUDA_CHECK(cudaFuncSetAttribute(zscore_time_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)(rows < (int64_t)eegx::kNumSMsB200 * 8 ? rows : (int64_t)eegx::kNumSMsB200 * 8);
    eegx::launch(zscore_time_kernel, grid, 256, smem, st, x, ch_idx, out, out_off, out_bstride, rows,
                                                (int)C_in, (int)C_out, (int)T);
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}
