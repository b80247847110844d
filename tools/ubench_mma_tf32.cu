// Throughput of the legacy warp-level tensor path on sm_100a: mma.sync.m16n8k8 TF32 (and m16n8k16 bf16 for scale),
// measured as warp-MMAs per clock per SM with 4 / 8 / 16 resident warps per SM and 4 independent accumulator chains.
// Decides whether the DSP kernel's 65-tap FIR can move to a 3xTF32 Toeplitz product on mma.sync.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_mma_tf32 tools/ubench_mma_tf32.cu && ./ubench_mma_tf32
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int KIND>
__global__ void bench(float* out, int iters, long long* cycles) {
    float c[4][4] = {};
    unsigned a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (KIND == 0) mma_tf32(c[j], a0, a1, a2, a3, b0, b1);
            else mma_bf16(c[j], a0, a1, a2, a3, b0, b1);
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int j = 0; j < 4; ++j) for (int e = 0; e < 4; ++e) s += c[j][e];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(float));
    cudaMallocManaged(&cyc, sizeof(long long));
    const int iters = 20000;
    for (int kind = 0; kind < 2; ++kind)
        for (int warps : {4, 8, 16, 32}) {
            for (int rep = 0; rep < 2; ++rep) {
                if (kind == 0) bench<0><<<148, warps * 32>>>(out, iters, cyc);
                else bench<1><<<148, warps * 32>>>(out, iters, cyc);
                cudaDeviceSynchronize();
            }
            const double mma_per_clk_sm = (double)warps * iters * 4 / (double)*cyc;
            const double macs = kind == 0 ? 16.0 * 8 * 8 : 16.0 * 8 * 16;
            printf("%s warps/SM %2d: %.3f warp-MMA/clk/SM = %.0f MAC/clk/SM  (~%.0f TFLOP/s at 1.9 GHz x 148 SMs)\n",
                   kind == 0 ? "tf32 m16n8k8 " : "bf16 m16n8k16", warps, mma_per_clk_sm, mma_per_clk_sm * macs,
                   mma_per_clk_sm * macs * 2 * 1.9e9 * 148 / 1e12);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
