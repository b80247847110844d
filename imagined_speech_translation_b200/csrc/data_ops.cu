// Data-side pieces of the hot path that the reference runs in numpy / sklearn on the host:
//
//   eegx_robust_fit_f32   RobustScaler(quantile_range=(q_lo, q_hi)).fit of
//                         EEGDataset._initialize_scalers_efficiently (main_model/src/data/dataset.py:102-151):
//                         per channel center = median, scale = q_hi - q_lo percentile over the concatenated
//                         fit samples (numpy 'linear' percentile: interpolation between order statistics),
//                         zero scales replaced by 1 (sklearn _handle_zeros_in_scale).  Exact selection: an
//                         MSB-first 8-bit radix select on order-preserving integer keys, one CTA per channel,
//                         integer shared-memory histograms (deterministic).
//   eegx_region_std_f32   np.std of a whole (C_r, T) region per trial (population), two fixed-order stages.
//   eegx_augment_f32      EEGDataset._augment_eeg_regions (dataset.py:227-261) for a batch: optional
//                         Gaussian noise (sigma per trial), amplitude scaling (factor per trial) and a
//                         circular time shift (per trial), one pass.  The Bernoulli decisions / factors /
//                         shifts are per-trial inputs (drawn by the caller), the noise is Philox + Box-Muller.
#include <math.h>

#include "eegx_common.h"
#include "fused_common.cuh"

namespace {

using namespace eegx;

__device__ __forceinline__ unsigned key_of(float v) {
    const unsigned u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);       // monotone: a < b  <=>  key(a) < key(b)
}
__device__ __forceinline__ float float_of(unsigned k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

// k-th smallest (0-based) of row[0..n): four 8-bit passes from the most significant digit
__device__ float radix_select(const float* __restrict__ row, long long n, long long k, unsigned* hist /*256*/,
                              unsigned* bcast /*2*/) {
    unsigned prefix = 0, mask = 0;
    for (int pass = 3; pass >= 0; --pass) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        const int sh = 8 * pass;
        for (long long i = threadIdx.x; i < n; i += blockDim.x) {
            const unsigned key = key_of(row[i]);
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> sh) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long cum = 0;
            unsigned d = 0;
            for (; d < 256; ++d) {
                if (cum + hist[d] > (unsigned long long)k) break;
                cum += hist[d];
            }
            bcast[0] = d;
            bcast[1] = (unsigned)cum;
        }
        __syncthreads();
        prefix |= bcast[0] << sh;
        mask |= 255u << sh;
        k -= bcast[1];
        __syncthreads();
    }
    return float_of(prefix);
}

__device__ double percentile_linear(const float* row, long long n, double q, unsigned* hist, unsigned* bcast) {
    const double pos = q * 0.01 * (double)(n - 1);
    const long long lo = (long long)floor(pos), hi = lo + 1 < n ? lo + 1 : n - 1;
    const double a = (double)radix_select(row, n, lo, hist, bcast);
    const double frac = pos - (double)lo;
    if (frac == 0.0 || hi == lo) return a;
    const double b = (double)radix_select(row, n, hi, hist, bcast);
    return a + (b - a) * frac;
}

__global__ void __launch_bounds__(256)
robust_fit_kernel(const float* __restrict__ x, long long n, float q_lo, float q_hi, float* __restrict__ center,
                  float* __restrict__ scale) {
    EEGX_PDL_SYNC();
    __shared__ unsigned hist[256];
    __shared__ unsigned bcast[2];
    const float* row = x + (long long)blockIdx.x * n;
    const double med = percentile_linear(row, n, 50.0, hist, bcast);
    const double lo = percentile_linear(row, n, (double)q_lo, hist, bcast);
    const double hi = percentile_linear(row, n, (double)q_hi, hist, bcast);
    if (threadIdx.x == 0) {
        center[blockIdx.x] = (float)med;
        const float s = (float)(hi - lo);
        scale[blockIdx.x] = (s == 0.0f || fabsf(s) < 10.0f * 1.1920929e-07f) ? 1.0f : s;   // _handle_zeros_in_scale
    }
}

// population std of each row of an (B, n) matrix; one CTA per row, fp64 fixed-order reduction
__global__ void __launch_bounds__(256)
row_std_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
    EEGX_PDL_SYNC();
    __shared__ double rs[8], rq[8];
    const float* row = x + (long long)blockIdx.x * n;
    double s = 0.0, q = 0.0;
    for (long long i = threadIdx.x; i < n; i += 256) {
        const double v = (double)row[i];
        s += v;
        q += v * v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if ((threadIdx.x & 31) == 0) { rs[threadIdx.x >> 5] = s; rq[threadIdx.x >> 5] = q; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double S = 0.0, Q = 0.0;
        for (int w = 0; w < 8; ++w) { S += rs[w]; Q += rq[w]; }
        const double mean = S / (double)n;
        double var = Q / (double)n - mean * mean;
        if (var < 0.0) var = 0.0;
        out[blockIdx.x] = (float)sqrt(var);
    }
}

// out[b, c, t] = scale[b] * (x[b, c, (t - shift[b]) mod T] + sigma[b] * N(0, 1))
__global__ void __launch_bounds__(256)
augment_kernel(const float* __restrict__ x, float* __restrict__ out, long long B, int C, int T,
               const float* __restrict__ sigma, const float* __restrict__ scale, const int* __restrict__ shift,
               DropoutCfg dc) {
    EEGX_PDL_SYNC();
    const long long per = (long long)C * T;
    const long long total = B * per;
    uint2 key = make_uint2(0u, 0u);
    if (dc.state != nullptr) {
        const unsigned long long seed = dc.state[0], step = dc.state[1];
        key = make_uint2((unsigned)seed ^ (unsigned)(step * 0x9E3779B97F4A7C15ull >> 32), (unsigned)(seed >> 32) ^ (unsigned)step);
    }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / per;
        const long long rem = i - b * per;
        const int c = (int)(rem / T), t = (int)(rem - (long long)c * T);
        int src_t = (t - shift[b]) % T;
        if (src_t < 0) src_t += T;
        const long long src = b * per + (long long)c * T + src_t;
        float v = x[src];
        const float sg = sigma[b];
        if (sg > 0.0f) {
            // one Philox call per source element (keyed by the SOURCE index: the noise rolls with the signal)
            const uint4 r = philox4x32_10(make_uint4((unsigned)src, (unsigned)((unsigned long long)src >> 32), dc.site, 0xA06E17u), key);
            const float u1 = ((float)(r.x >> 8) + 0.5f) * (1.0f / 16777216.0f);
            const float u2 = ((float)(r.y >> 8) + 0.5f) * (1.0f / 16777216.0f);
            v += sg * sqrtf(-2.0f * __logf(u1)) * __cosf(6.283185307179586f * u2);
        }
        out[i] = v * scale[b];
    }
}

}  // namespace

extern "C" {

int eegx_robust_fit_f32(const float* x, int64_t C, int64_t n, float q_lo, float q_hi, float* center, float* scale,
                        void* stream) {
    if (int rc = eegx::require_sm100()) return rc;
    EEGX_REQUIRE(C >= 0 && n >= 1 && C < (1LL << 31), EEGX_ERR_SHAPE, "robust_fit: need n >= 1");
    EEGX_REQUIRE(q_lo >= 0.0f && q_hi <= 100.0f && q_lo < q_hi, EEGX_ERR_ARG, "robust_fit: need 0 <= q_lo < q_hi <= 100");
    if (C == 0) return EEGX_OK;
    EEGX_REQUIRE(x && center && scale, EEGX_ERR_ARG, "robust_fit: NULL pointer");
    eegx::launch(robust_fit_kernel, (unsigned)C, 256, 0, static_cast<cudaStream_t>(stream), x, n, q_lo, q_hi, center, scale);
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}

int eegx_region_std_f32(const float* x, int64_t B, int64_t n, float* out, void* stream) {
    if (int rc = eegx::require_sm100()) return rc;
    EEGX_REQUIRE(B >= 0 && n >= 1 && B < (1LL << 31), EEGX_ERR_SHAPE, "region_std: need n >= 1");
    if (B == 0) return EEGX_OK;
    EEGX_REQUIRE(x && out, EEGX_ERR_ARG, "region_std: NULL pointer");
    eegx::launch(row_std_kernel, (unsigned)B, 256, 0, static_cast<cudaStream_t>(stream), x, n, out);
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}

int eegx_augment_f32(const float* x, float* out, int64_t B, int64_t C, int64_t T, const float* sigma,
                     const float* scale, const int32_t* shift, const uint64_t* rng_state, uint32_t site, void* stream) {
    if (int rc = eegx::require_sm100()) return rc;
    EEGX_REQUIRE(B >= 0 && C >= 1 && T >= 1 && C * T < (1LL << 31), EEGX_ERR_SHAPE, "augment: bad sizes");
    if (B == 0) return EEGX_OK;
    EEGX_REQUIRE(x && out && sigma && scale && shift, EEGX_ERR_ARG, "augment: NULL pointer");
    EEGX_REQUIRE(x != out, EEGX_ERR_ARG, "augment: out must not alias x (circular shift)");
    const DropoutCfg dc{reinterpret_cast<const unsigned long long*>(rng_state), site, 0.0f};
    long long blocks = (B * C * T + 255) / 256;
    const long long cap = (long long)kNumSMsB200 * 8;
    if (blocks > cap) blocks = cap;
    eegx::launch(augment_kernel, (int)blocks, 256, 0, static_cast<cudaStream_t>(stream), x, out, B, (int)C, (int)T, sigma, scale,
                                                                              reinterpret_cast<const int*>(shift), dc);
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}

}  // extern "C"
