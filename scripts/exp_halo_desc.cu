// Experiment: can a tcgen05 K-major SWIZZLE_128B operand descriptor start at an arbitrary 128-byte row of a
// 1024-byte-aligned, absolutely-swizzled buffer, with a stride between 8-row groups that is NOT a multiple of 1024 B?
// (That is what reusing one activation halo tile for all 9 taps of a 3x3 conv needs: tap (dy,dx) of a 16x8 pixel tile
// inside an 18x10 halo starts at pixel dy*10+dx, and the 8-pixel row groups are 10 pixels = 1280 B apart.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -o exp_halo_desc scripts/exp_halo_desc.cu && ./exp_halo_desc
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

using bf16 = __nv_bfloat16;
constexpr int PIX = 512, N = 64, M = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int ROWB>
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t base_offset) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_offset & 7) << 49;
    d |= (uint64_t)(ROWB == 128 ? 2 : 4) << 61;
    return d;
}

// a_src: [PIX][K] bf16 (pixel-major), b_src: [N][K]; out: [M][N] fp32
// ROWB = bytes per operand row = swizzle width (128: K = 64 bf16, SWIZZLE_128B; 64: K = 32, SWIZZLE_64B).  MT = the
// shifted operand is the A (M = 128 rows) or the B (N = 256 rows) operand.
template <int ROWB, bool SHIFT_B>
__global__ void __launch_bounds__(128, 1) exp_kernel(const bf16* a_src, const bf16* b_src, float* out, int off_px, int sbo_bytes, int base_offset) {
    constexpr int K = ROWB / 2, CH = ROWB / 16, NS = SHIFT_B ? 256 : N;
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sa = smem;                       // PIX rows of ROWB bytes, swizzled by ABSOLUTE address (what TMA writes)
    uint8_t* sb = smem + PIX * ROWB;          // the other operand: 128 (or N) rows, unshifted
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    const int tid = threadIdx.x, warp = tid >> 5;
    // absolute-address swizzle: 16-byte chunk index ^= address bits [7..] (3 bits for 128B rows, 2 bits for 64B rows)
    for (int i = tid; i < PIX * CH; i += 128) {
        const int p = i / CH, ch = i % CH;
        const uint32_t row = (uint32_t)(p * ROWB);
        const uint32_t x = ROWB == 128 ? ((row >> 7) & 7) : ((row >> 7) & 3);
        *reinterpret_cast<uint4*>(sa + row + ((ch ^ x) << 4)) = *reinterpret_cast<const uint4*>(a_src + p * K + ch * 8);
    }
    for (int i = tid; i < 128 * CH; i += 128) {
        const int p = i / CH, ch = i % CH;
        const uint32_t row = (uint32_t)(p * ROWB);
        const uint32_t x = ROWB == 128 ? ((row >> 7) & 7) : ((row >> 7) & 3);
        *reinterpret_cast<uint4*>(sb + row + ((ch ^ x) << 4)) = *reinterpret_cast<const uint4*>(b_src + p * K + ch * 8);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the MMA (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_ptr;
    if (tid == 0) {
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NS >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint64_t dshift = make_desc<ROWB>(smem_u32(sa) + off_px * ROWB, sbo_bytes, base_offset);
        const uint64_t dplain = make_desc<ROWB>(smem_u32(sb), 8 * ROWB, 0);
        for (int k = 0; k < K / 16; ++k) {
            const uint64_t koff = (uint64_t)((k * 32) >> 4);
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                ::"r"(tmem), "l"((SHIFT_B ? dplain : dshift) + koff), "l"((SHIFT_B ? dshift : dplain) + koff), "r"(idesc), "r"(k ? 1u : 0u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // everyone waits for the MMAs
    {
        uint32_t ok = 0;
        long long t0 = clock64();
        while (!ok) {
            asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
            if (clock64() - t0 > 2000000000ll) { asm volatile("trap;"); }
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = tid;                         // warp w may read TMEM lanes 32w..32w+31
    for (int c0 = 0; c0 < NS; c0 += 32) {
        uint32_t r[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
              "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
              "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) out[row * NS + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
    }
}

template <int ROWB, bool SHIFT_B>
int run_cases(const char* title, const int (*cases)[3], int n_cases) {
    constexpr int K = ROWB / 2, NS = SHIFT_B ? 256 : N;
    std::vector<bf16> a(PIX * K), b(128 * K);
    std::vector<float> af(PIX * K), bf(128 * K);
    srand(1);
    for (int i = 0; i < PIX * K; ++i) { float v = (rand() % 17 - 8) / 8.0f; a[i] = __float2bfloat16(v); af[i] = v; }
    for (int i = 0; i < 128 * K; ++i) { float v = (rand() % 13 - 6) / 4.0f; b[i] = __float2bfloat16(v); bf[i] = v; }
    bf16 *da, *db; float* dout;
    cudaMalloc(&da, a.size() * 2); cudaMalloc(&db, b.size() * 2); cudaMalloc(&dout, M * 256 * 4);
    cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
    const int smem = PIX * ROWB + 128 * ROWB + 1024;
    cudaFuncSetAttribute(exp_kernel<ROWB, SHIFT_B>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    std::vector<float> out(M * 256);
    printf("%s\n", title);
    for (int ci = 0; ci < n_cases; ++ci) {
        const int off = cases[ci][0], sbo = cases[ci][1], bo = cases[ci][2];
        cudaMemset(dout, 0, M * 256 * 4);
        exp_kernel<ROWB, SHIFT_B><<<1, 128, smem>>>(da, db, dout, off, sbo, bo);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("off=%d sbo=%d bo=%d: CUDA error %s\n", off, sbo, bo, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(out.data(), dout, M * 256 * 4, cudaMemcpyDeviceToHost);
        int bad = 0; double maxerr = 0;
        // shifted operand row r lives at pixel off + (r / 8) * (sbo / ROWB) + r % 8
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < NS; ++n) {
                const int rs = SHIFT_B ? n : m, ro = SHIFT_B ? m : n;
                const int p = off + (rs / 8) * (sbo / ROWB) + (rs % 8);
                float ref = 0;
                for (int k = 0; k < K; ++k) ref += af[p * K + k] * bf[ro * K + k];
                const double err = fabs(ref - out[m * NS + n]);
                if (err > 1e-3) ++bad;
                if (err > maxerr) maxerr = err;
            }
        printf("  off_px=%2d sbo=%4d base_offset=%d : %s (bad=%d, max err %.3g)\n", off, sbo, bo, bad ? "MISMATCH" : "ok", bad, maxerr);
    }
    cudaFree(da); cudaFree(db); cudaFree(dout);
    return 0;
}

int main() {
    const int c128[][3] = {{0, 1024, 0}, {1, 1024, 0}, {1, 1024, 1}, {3, 1024, 0}, {0, 1280, 0}, {1, 1280, 0}, {11, 1280, 0}, {11, 1280, 3},
                           {22, 1280, 0}, {5, 2048, 0}};
    if (run_cases<128, false>("A operand (M = 128 rows) shifted, 128 B rows / SWIZZLE_128B", c128, 10)) return 1;
    if (run_cases<128, true>("B operand (N = 256 rows) shifted, 128 B rows / SWIZZLE_128B", c128, 9)) return 1;
    const int c64[][3] = {{0, 512, 0}, {1, 512, 0}, {3, 512, 0}, {0, 640, 0}, {1, 640, 0}, {11, 640, 0}, {22, 640, 0}, {11, 640, 1}, {2, 1024, 0}};
    if (run_cases<64, false>("A operand (M = 128 rows) shifted, 64 B rows / SWIZZLE_64B", c64, 9)) return 1;
    if (run_cases<64, true>("B operand (N = 256 rows) shifted, 64 B rows / SWIZZLE_64B", c64, 8)) return 1;
    return 0;
}
