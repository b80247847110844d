// Issue / execution rate of tcgen05.mma kind::tf32 with small N on sm_100a: cycles per MMA (M = 128, K = 8) for
// N = 16..256, A from tensor memory (.ts form) or from shared memory (.ss form), measured on one CTA as
// clock64 from the first issue to the mbarrier flip of the closing tcgen05.commit.  Decides the tile shape of the
// DSP kernel's Toeplitz FIR (csrc/dsp_umma.cu).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_umma_tf32 tools/ubench_umma_tf32.cu && ./ubench_umma_tf32
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long make_desc(unsigned saddr) {
    unsigned long long d = 0;
    d |= (unsigned long long)((saddr >> 4) & 0x3FFF);
    d |= (unsigned long long)(1) << 16;
    d |= (unsigned long long)(1024 >> 4) << 32;
    d |= 1ull << 46;
    d |= 2ull << 61;
    return d;
}

template <bool TS, bool F16>
__global__ void __launch_bounds__(128, 1) bench(int N, int iters, int nacc, long long* cycles) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ unsigned tmem_slot;
    __shared__ __align__(8) unsigned long long bar;
    unsigned char* base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<unsigned*>(base)[i] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tm = tmem_slot;
    if (threadIdx.x == 0) {
        const unsigned fmt = F16 ? 1u : 2u;   // bf16 / tf32
        const unsigned idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((unsigned)(N >> 3) << 17) | ((128u >> 4) << 24);
        const unsigned a_smem = smem_u32(base), b_smem = smem_u32(base) + 32 * 1024;
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const unsigned ks = i & 3;
            const unsigned long long bdesc = make_desc(b_smem + ks * 32);
            if (TS) {
                if (F16)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                                 ::"r"(tm + 256 + 32 * (i % nacc)), "r"(tm + 8 * ks), "l"(bdesc), "r"(idesc), "r"(i) : "memory");
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                                 ::"r"(tm + 256 + 32 * (i % nacc)), "r"(tm + 8 * ks), "l"(bdesc), "r"(idesc), "r"(i) : "memory");
            } else {
                const unsigned long long adesc = make_desc(a_smem + ks * 32);
                if (F16)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(tm + 256 + 32 * (i % nacc)), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(i) : "memory");
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(tm + 256 + 32 * (i % nacc)), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(i) : "memory");
            }
        }
        const long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile(
            "{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra.uni WD;\n\tbra.uni WL;\n\tWD:\n\t}"
            ::"r"(smem_u32(&bar)) : "memory");
        const long long t2 = clock64();
        cycles[0] = t1 - t0;
        cycles[1] = t2 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

int main() {
    long long* cyc;
    cudaMallocManaged(&cyc, 2 * sizeof(long long));
    const int iters = 2000;
    const int smem = 65 * 1024;
    cudaFuncSetAttribute(bench<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(bench<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(bench<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(bench<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int nacc : {1, 2, 4, 6})
    for (int mode = 0; mode < (nacc == 1 ? 4 : 1); ++mode)
        for (int N : {16, 32, 64, 128, 256}) {
            if (nacc > 1 && N > 32) continue;
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) bench<true, false><<<1, 128, smem>>>(N, iters, nacc, cyc);
                if (mode == 1) bench<false, false><<<1, 128, smem>>>(N, iters, nacc, cyc);
                if (mode == 2) bench<true, true><<<1, 128, smem>>>(N, iters, nacc, cyc);
                if (mode == 3) bench<false, true><<<1, 128, smem>>>(N, iters, nacc, cyc);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
            }
            const char* names[4] = {"tf32 A=tmem", "tf32 A=smem", "bf16 A=tmem", "bf16 A=smem"};
            const double kk = mode < 2 ? 8.0 : 16.0;
            printf("%s acc=%d M=128 N=%3d K=%2.0f: issue %.1f clk/MMA, complete %.1f clk/MMA  (%.0f MAC/clk/SM)\n", names[mode], nacc, N, kk,
                   (double)cyc[0] / iters, (double)cyc[1] / iters, 128.0 * N * kk * iters / (double)cyc[1]);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
