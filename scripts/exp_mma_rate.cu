// Experiment: issue rate of tcgen05.mma.kind::f16 (bf16, M = 128, cta_group::1) from shared-memory operands for
// N = 64 / 128 / 256, one CTA per SM on every SM, operands resident (no TMA traffic): clocks per MMA.
// If N = 128 takes clearly more than 64 clk, the small-N layers are bound by operand fetch from shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o exp_mma_rate scripts/exp_mma_rate.cu && ./exp_mma_rate
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

template <int N, bool A_TMEM>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int iters) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (128 * 128 * 2 + 256 * 128 * 2) / 4; i += 128) ((uint32_t*)smem)[i] = 0;   // zeros: no NaN slow paths
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_ptr;
    if (tid == 0) {
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        // two A tiles (hi, lo) of 128 x 64 and two B tiles of N x 64, 4 K-steps each: the 3-pass pattern of the conv kernel
        const uint32_t sa = smem_u32(smem), sb = sa + 2 * 128 * 128;
        const uint64_t ah = make_desc(sa), al = make_desc(sa + 128 * 128), bh = make_desc(sb), bl = make_desc(sb + N * 128);
        const uint32_t a_tm = tmem + 256;           // A staging columns when A_TMEM
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint64_t ko = (uint64_t)((k * 32) >> 4);
                if (A_TMEM) {
                    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(a_tm + (uint32_t)((it & 7) * 16)), "l"(ah + ko) : "memory");
                    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(a_tm + (uint32_t)((it & 7) * 16 + 8)), "l"(al + ko) : "memory");
                    const uint32_t ta = a_tm + (uint32_t)((it & 7) * 16);
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                                 ::"r"(tmem), "r"(ta), "l"(bh + ko), "r"(idesc), "r"(1u) : "memory");
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                                 ::"r"(tmem), "r"(ta), "l"(bl + ko), "r"(idesc), "r"(1u) : "memory");
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                                 ::"r"(tmem), "r"(ta + 8), "l"(bh + ko), "r"(idesc), "r"(1u) : "memory");
                } else {
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(tmem), "l"(ah + ko), "l"(bh + ko), "r"(idesc), "r"(1u) : "memory");
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(tmem), "l"(ah + ko), "l"(bl + ko), "r"(idesc), "r"(1u) : "memory");
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(tmem), "l"(al + ko), "l"(bh + ko), "r"(idesc), "r"(1u) : "memory");
                }
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
            if (clock64() - t0 > 4000000000ll) asm volatile("trap;");
        }
        out[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

template <int N, bool A_TMEM>
void run(const char* name) {
    long long* d;
    cudaMalloc(&d, 148 * sizeof(long long));
    const int smem = 2 * 128 * 128 + 2 * 256 * 128 + 1024, iters = 2000;
    cudaFuncSetAttribute(rate_kernel<N, A_TMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int rep = 0; rep < 2; ++rep) rate_kernel<N, A_TMEM><<<148, 128, smem>>>(d, iters);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); exit(1); }
    long long h[148];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += (double)h[i];
    avg /= 148;
    const double per_mma = avg / (iters * 12.0);
    printf("%-28s %7.1f clk per MMA (math alone: %d clk) -> %.0f %% of the tensor peak\n", name, per_mma, N / 2, 100.0 * (N / 2) / per_mma);
    cudaFree(d);
}

int main() {
    run<64, false>("N= 64 A,B from smem");
    run<128, false>("N=128 A,B from smem");
    run<256, false>("N=256 A,B from smem");
    run<64, true>("N= 64 A via tcgen05.cp/TMEM");
    run<128, true>("N=128 A via tcgen05.cp/TMEM");
    run<256, true>("N=256 A via tcgen05.cp/TMEM");
    return 0;
}
