// Generic fused DSP kernel: any power-of-two n_fft, any hop, any odd tap count.
//
// Spec (the reference has no DSP code of its own):
//   window -> FIR "same" (zero pad) -> STFT (center, reflect pad, periodic Hann,
//   one-sided) -> log(|X|^2 + log_eps) -> per-(trial, channel) z-score.
//
// One CTA owns one (trial, channel) row at a time (persistent grid-stride loop):
// the T input samples are read from HBM once, everything in between lives in
// shared memory, and the F*N_f normalised values are written once, coalesced.
// This kernel favours generality; BASELINE config 2 (n_fft 256 / hop 64 / 65 taps)
// dispatches to the tuned kernel in dsp_tuned.cu, and the two are checked against
// each other and the oracle in tests/test_dsp_gpu.py.
#include "dsp_plan.h"

namespace {

constexpr int NT = 256;
constexpr int NWARPS = NT / 32;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        float t = lane < NWARPS ? red[lane] : 0.0f;
        t = warp_sum(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

__device__ __forceinline__ int reflect(int i, int T) {
    // torch.stft(center=True, pad_mode='reflect'): -i for i < 0, 2(T-1)-i for i >= T
    i = i < 0 ? -i : i;
    return i >= T ? 2 * (T - 1) - i : i;
}

__global__ void __launch_bounds__(NT) dsp_generic_kernel(const eegx::DspArgs a) {
    extern __shared__ __align__(16) float smem[];
    __shared__ float red[33];
    const int T = a.T, K = a.numtaps, P = (K - 1) / 2, N = a.n_fft, M = N / 2;
    const int F = a.F, NF = a.n_frames;
    float* taps_s = smem;                               // K (padded to 132)
    float* win_s = taps_s + 132;                        // N
    float2* tw_s = reinterpret_cast<float2*>(win_s + N);  // N/2 complex
    float* xs = reinterpret_cast<float*>(tw_s + M);     // T + 2P (+pad)
    float* ys = xs + ((T + 2 * P + 3) & ~3);            // T
    float* Ls = ys + ((T + 3) & ~3);                    // F * NF
    float2* scratch = reinterpret_cast<float2*>(Ls + ((F * NF + 3) & ~3));  // NWARPS * M complex

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < K; i += NT) taps_s[i] = a.taps[i];
    for (int i = tid; i < N; i += NT) win_s[i] = a.window[i];
    for (int i = tid; i < M; i += NT) tw_s[i] = a.twiddle[i];

    for (int64_t row = blockIdx.x; row < a.rows; row += gridDim.x) {
        const int64_t b = row / a.C;
        const int c = (int)(row - b * a.C);
        const float* src = a.onsets ? a.x + (int64_t)c * a.rec_len + a.onsets[b]
                                    : a.x + row * (int64_t)T;
        __syncthreads();  // previous row fully consumed (and tables visible on the first pass)
        for (int i = tid; i < T + 2 * P; i += NT)
            xs[i] = (i >= P && i < P + T) ? __ldg(src + (i - P)) : 0.0f;
        __syncthreads();

        // FIR: y[t] = sum_k h[k] * x[t + P - k]; xs is shifted by P.
        for (int t = tid; t < T; t += NT) {
            float acc = 0.0f;
            const float* xp = xs + t + 2 * P;
            for (int k = 0; k < K; ++k) acc = fmaf(taps_s[k], xp[-k], acc);
            ys[t] = acc;
        }
        __syncthreads();

        // STFT: one warp per frame, radix-2 DIT in shared memory on the packed
        // complex sequence z[n] = y[2n] + i*y[2n+1] (real FFT of length N via a
        // complex FFT of length N/2 plus the split step).
        float2* z = scratch + warp * M;
        for (int m = warp; m < NF; m += NWARPS) {
            const int s = m * a.hop - M;
            for (int n = lane; n < M; n += 32) {
                const int i0 = reflect(s + 2 * n, T), i1 = reflect(s + 2 * n + 1, T);
                const int r = (int)(__brev((unsigned)n) >> (32 - a.log2_m));
                z[r] = make_float2(ys[i0] * win_s[2 * n], ys[i1] * win_s[2 * n + 1]);
            }
            __syncwarp();
            for (int hs = 1; hs < M; hs <<= 1) {
                const int tw_stride = M / hs;  // table holds exp(-2*pi*i*k/N): W_M^p = tw[2p]
                for (int j = lane; j < M / 2; j += 32) {
                    const int pos = j & (hs - 1);
                    const int i0 = ((j - pos) << 1) + pos, i1 = i0 + hs;
                    const float2 w = tw_s[pos * tw_stride];
                    const float2 u = z[i0], v = z[i1];
                    const float tr = fmaf(w.x, v.x, -w.y * v.y);
                    const float ti = fmaf(w.x, v.y, w.y * v.x);
                    z[i0] = make_float2(u.x + tr, u.y + ti);
                    z[i1] = make_float2(u.x - tr, u.y - ti);
                }
                __syncwarp();
            }
            // split step + power + log.  Pairs (k, M-k), k = 0..M/2.
            for (int k = lane; k <= M / 2; k += 32) {
                if (k == 0) {
                    const float2 z0 = z[0];
                    const float x0 = z0.x + z0.y, xm = z0.x - z0.y;
                    Ls[0 * NF + m] = logf(fmaf(x0, x0, a.log_eps));
                    Ls[M * NF + m] = logf(fmaf(xm, xm, a.log_eps));
                } else {
                    const float2 zk = z[k], zm = z[M - k];
                    const float er = 0.5f * (zk.x + zm.x), ei = 0.5f * (zk.y - zm.y);
                    const float dr = zk.x - zm.x, di = zk.y + zm.y;
                    const float orr = 0.5f * di, oi = -0.5f * dr;   // Xo = -i/2 * D
                    const float2 w = tw_s[k];
                    const float tr = fmaf(w.x, orr, -w.y * oi);
                    const float ti = fmaf(w.x, oi, w.y * orr);
                    const float ar = er + tr, ai = ei + ti;         // X[k]
                    const float br = er - tr, bi = ei - ti;         // conj(X[M-k])
                    Ls[k * NF + m] = logf(fmaf(ar, ar, fmaf(ai, ai, a.log_eps)));
                    Ls[(M - k) * NF + m] = logf(fmaf(br, br, fmaf(bi, bi, a.log_eps)));
                }
            }
            __syncwarp();
        }
        __syncthreads();

        // z-score over all F*NF values (two passes over shared memory).
        const int n_out = F * NF;
        float acc = 0.0f;
        for (int i = tid; i < n_out; i += NT) acc += Ls[i];
        const float mean = block_sum(acc, red) / (float)n_out;
        float dev = 0.0f;
        for (int i = tid; i < n_out; i += NT) {
            const float d = Ls[i] - mean;
            dev = fmaf(d, d, dev);
        }
        const float var = block_sum(dev, red) / (float)n_out;
        const float inv = 1.0f / (sqrtf(var) + a.z_eps);
        float* dst = a.out + row * (int64_t)n_out;
        for (int i = tid; i < n_out; i += NT) __stcs(dst + i, (Ls[i] - mean) * inv);
    }
}

}  // namespace

namespace eegx {

size_t dsp_generic_smem_bytes(int T, int n_fft, int hop, int numtaps) {
    const int P = (numtaps - 1) / 2, M = n_fft / 2, F = M + 1, NF = 1 + T / hop;
    size_t floats = 132 + n_fft + 2 * (size_t)M + ((T + 2 * P + 3) & ~3) + ((T + 3) & ~3) +
                    ((F * NF + 3) & ~3) + 2 * (size_t)NWARPS * M;
    return floats * sizeof(float);
}

int launch_dsp_generic(const eegx_dsp_plan* plan, const DspArgs& a, cudaStream_t st) {
    const size_t smem = plan->smem_generic;
    EEGX_CUDA_CHECK(cudaFuncSetAttribute(dsp_generic_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    int64_t grid = (int64_t)kNumSMsB200 * per_sm;
    if (grid > a.rows) grid = a.rows;
    dsp_generic_kernel<<<(int)grid, NT, smem, st>>>(a);
    EEGX_CUDA_CHECK(cudaGetLastError());
    return EEGX_OK;
}

}  // namespace eegx
