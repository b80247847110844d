// Micro-benchmark: scalar FFMA vs packed fma.rn.f32x2 (FFMA2) throughput on sm_100a.
// Decides whether the DSP kernel's FIR / FFT inner loops should be written with f32x2 ops.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_fma tools/ubench_fma.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;
constexpr int ACC = 8;

__global__ void k_ffma(float* out, float a, float b) {
    float acc[ACC];
#pragma unroll
    for (int i = 0; i < ACC; ++i) acc[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ACC; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ffma2(float* out, float a, float b) {
    unsigned long long acc[ACC];
    unsigned long long av, bv;
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(av) : "f"(a));
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(bv) : "f"(b));
#pragma unroll
    for (int i = 0; i < ACC; ++i) {
        float x = threadIdx.x * 1e-3f + i;
        asm volatile("mov.b64 %0, {%1, %1};" : "=l"(acc[i]) : "f"(x));
    }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ACC; ++i)
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(av), "l"(bv));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ACC; ++i) {
        float lo, hi;
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i]));
        s += lo + hi;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_lg2(float* out, float a) {
    float acc[ACC];
#pragma unroll
    for (int i = 0; i < ACC; ++i) acc[i] = threadIdx.x + 2.0f + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ACC; ++i) acc[i] = __log2f(acc[i]) + a;
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_shfl(float* out) {
    float acc[ACC];
#pragma unroll
    for (int i = 0; i < ACC; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ACC; ++i) acc[i] = __shfl_xor_sync(0xffffffffu, acc[i], 1 + (i & 3));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_lds128(float* out) {
    __shared__ float4 buf[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    float4 acc = make_float4(0, 0, 0, 0);
    int idx = threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ACC; ++i) {
            float4 v = buf[(idx + i * 32) & 1023];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        idx += 7;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}


__global__ void k_fadd(float* out, float a) {
    float acc[ACC];
#pragma unroll
    for (int i = 0; i < ACC; ++i) acc[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ACC; ++i) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(acc[i]) : "f"(a));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 1 FFMA : 1 FADD interleaved on independent accumulators (do they share a pipe?)
__global__ void k_mix_ffma_fadd(float* out, float a, float b) {
    float acc[ACC], acd[ACC];
#pragma unroll
    for (int i = 0; i < ACC; ++i) { acc[i] = threadIdx.x * 1e-3f + i; acd[i] = i; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ACC; ++i) {
            asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(acc[i]) : "f"(a), "f"(b));
            asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(acd[i]) : "f"(b));
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ACC; ++i) s += acc[i] + acd[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 2 FFMA : 1 DFMA interleaved (is the fp64 pipe usable as extra FMA capacity?)
__global__ void k_mix_ffma_dfma(float* out, float a, float b, double c) {
    float acc[ACC];
    double dcc[ACC / 2];
#pragma unroll
    for (int i = 0; i < ACC; ++i) acc[i] = threadIdx.x * 1e-3f + i;
#pragma unroll
    for (int i = 0; i < ACC / 2; ++i) dcc[i] = i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ACC; ++i) {
            asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(acc[i]) : "f"(a), "f"(b));
            if ((i & 1) == 0) asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(dcc[i / 2]) : "d"(c));
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ACC; ++i) s += acc[i];
#pragma unroll
    for (int i = 0; i < ACC / 2; ++i) s += (float)dcc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dfma(float* out, double c) {
    double dcc[ACC];
#pragma unroll
    for (int i = 0; i < ACC; ++i) dcc[i] = i + threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ACC; ++i) asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(dcc[i]) : "d"(c));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ACC; ++i) s += (float)dcc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FFMA2 with both multiplicands in registers + 1 LDS.64 per 4 FFMA2 (constant pairs fetched from smem)
__global__ void k_ffma2_lds(float* out, float a) {
    __shared__ unsigned long long tab[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        unsigned long long v; float f = 1.0f + i * 1e-6f;
        asm volatile("mov.b64 %0, {%1, %1};" : "=l"(v) : "f"(f));
        tab[i] = v;
    }
    __syncthreads();
    unsigned long long acc[ACC], bv;
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(bv) : "f"(a));
#pragma unroll
    for (int i = 0; i < ACC; ++i) { float x = threadIdx.x * 1e-3f + i; asm volatile("mov.b64 %0, {%1, %1};" : "=l"(acc[i]) : "f"(x)); }
    int idx = threadIdx.x & 7;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ACC; i += 4) {
            const unsigned long long w = tab[(idx + i) & 255];
#pragma unroll
            for (int j = 0; j < 4; ++j) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i + j]) : "l"(w), "l"(bv));
        }
        idx += 3;
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ACC; ++i) { float lo, hi; asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i])); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < 5; ++i) f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / 5;
}

int main() {
    int sms = 0, clk = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int blocks = sms * 8, threads = 256;
    float* out; cudaMalloc(&out, sizeof(float) * blocks * threads);
    const double lanes = (double)blocks * threads * ITERS * ACC;
    float t1 = time_ms([&] { k_ffma<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
    float t2 = time_ms([&] { k_ffma2<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
    float t3 = time_ms([&] { k_lg2<<<blocks, threads>>>(out, 3.0f); });
    float t4 = time_ms([&] { k_shfl<<<blocks, threads>>>(out); });
    float t5 = time_ms([&] { k_lds128<<<blocks, threads>>>(out); });
    printf("SMs %d clock %d kHz\n", sms, clk);
    printf("FFMA : %.3f ms  %.2f Tfma/s   (%.1f fma/clk/SM at max clock)\n", t1, lanes / t1 * 1e-9, lanes / (t1 * 1e-3) / sms / (clk * 1e3));
    printf("FFMA2: %.3f ms  %.2f Tfma/s   (%.1f fma/clk/SM at max clock)\n", t2, 2 * lanes / t2 * 1e-9, 2 * lanes / (t2 * 1e-3) / sms / (clk * 1e3));
    printf("LG2  : %.3f ms  %.2f Tops/s   (%.1f /clk/SM)\n", t3, lanes / t3 * 1e-9, lanes / (t3 * 1e-3) / sms / (clk * 1e3));
    printf("SHFL : %.3f ms  %.2f Tlane/s  (%.1f lanes/clk/SM)\n", t4, lanes / t4 * 1e-9, lanes / (t4 * 1e-3) / sms / (clk * 1e3));
    printf("LDS128: %.3f ms %.2f TB/s     (%.1f B/clk/SM)\n", t5, 16 * lanes / t5 * 1e-9, 16 * lanes / (t5 * 1e-3) / sms / (clk * 1e3));
    float t6 = time_ms([&] { k_fadd<<<blocks, threads>>>(out, 0.5f); });
    float t7 = time_ms([&] { k_mix_ffma_fadd<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
    float t8 = time_ms([&] { k_dfma<<<blocks, threads>>>(out, 1.0000001); });
    float t9 = time_ms([&] { k_mix_ffma_dfma<<<blocks, threads>>>(out, 1.0001f, 0.5f, 1.0000001); });
    float t10 = time_ms([&] { k_ffma2_lds<<<blocks, threads>>>(out, 0.5f); });
    printf("FADD : %.3f ms  (%.1f add/clk/SM)\n", t6, lanes / (t6 * 1e-3) / sms / (clk * 1e3));
    printf("FFMA+FADD 1:1: %.3f ms  (%.1f fp-ops/clk/SM; 2 ops per pair)\n", t7, 2 * lanes / (t7 * 1e-3) / sms / (clk * 1e3));
    printf("DFMA : %.3f ms  (%.1f dfma/clk/SM)\n", t8, lanes / (t8 * 1e-3) / sms / (clk * 1e3));
    printf("FFMA+DFMA 2:1: %.3f ms  (%.1f ffma/clk/SM + %.1f dfma/clk/SM)\n", t9, lanes / (t9 * 1e-3) / sms / (clk * 1e3), 0.5 * lanes / (t9 * 1e-3) / sms / (clk * 1e3));
    printf("FFMA2 + LDS.64 per 4: %.3f ms  (%.1f fma/clk/SM)\n", t10, 2 * lanes / (t10 * 1e-3) / sms / (clk * 1e3));
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
