// Shared helpers for libsis_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>
#include "../../include/sis_b200.h"

namespace sis {

void set_error(const char* fmt, ...);
extern std::atomic<unsigned long long> g_launches;
// One word of mapped, portable pinned host memory shared by every bounded device-side wait of the library: a watchdog
// writes its code there before it traps, so the host can still read WHICH wait gave up after the trap has poisoned the
// CUDA context (sis_watchdog_code).  Null when the allocation failed (the kernels then only trap).
unsigned int* watchdog_word();
inline void count_launch(int n = 1) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

#define SIS_CHECK_CUDA(expr)                                                                      \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            sis::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return SIS_ERR_CUDA;                                                                  \
        }                                                                                         \
    } while (0)

#define SIS_CHECK_LAUNCH()                                                                        \
    do {                                                                                          \
        cudaError_t _e = cudaGetLastError();                                                      \
        if (_e != cudaSuccess) {                                                                  \
            sis::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
            return SIS_ERR_CUDA;                                                                  \
        }                                                                                         \
        sis::count_launch();                                                                      \
    } while (0)

#define SIS_REQUIRE(cond, ...)                                                                    \
    do {                                                                                          \
        if (!(cond)) {                                                                            \
            sis::set_error(__VA_ARGS__);                                                          \
            return SIS_ERR_INVALID;                                                               \
        }                                                                                         \
    } while (0)

#define SIS_PROPAGATE(expr)                                                                       \
    do {                                                                                          \
        int _s = (expr);                                                                          \
        if (_s != SIS_OK) return _s;                                                              \
    } while (0)

constexpr int kNumSMs = 148;

// Optional per-kernel-category timing with CUDA events on the launching stream (bench.py's roofline leg).
enum ProfCat { PROF_MAPPING = 0, PROF_CONV_TC = 1, PROF_BLUR_SPLIT = 2, PROF_TORGB = 3, PROF_CONV_SIMT = 4,
               PROF_LABEL = 5, PROF_BLUR_SIMT = 6, PROF_OTHER = 7, PROF_CONV_TC_NARROW = 8, PROF_NUM = 9 };
extern bool g_prof_on;
void prof_begin(int cat, cudaStream_t stream);
void prof_end(int cat, cudaStream_t stream);
struct ProfScope {
    int cat; cudaStream_t stream; bool on;
    ProfScope(int c, cudaStream_t s) : cat(c), stream(s), on(g_prof_on) { if (on) prof_begin(cat, stream); }
    ~ProfScope() { if (on) prof_end(cat, stream); }
};

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

__device__ __forceinline__ float lrelu_scale(float x, float alpha, float scale) {
    // (x > 0 ? x : x*alpha) * scale with the reference's rounding sequence
    // (fused_bias_act_kernel.cu:39,47): two separately rounded multiplies, never an FMA.
    float y = (x > 0.0f) ? x : __fmul_rn(x, alpha);
    return __fmul_rn(y, scale);
}

// packed fp32x2 arithmetic (sm_100: FFMA2 / FADD2 / FMUL2): two IEEE fp32 operations per instruction, per-lane results
// identical to the scalar instructions
using u64_t = unsigned long long;
__device__ __forceinline__ float2 pk_fma(float2 a, float2 b, float2 c) {
    float2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(reinterpret_cast<u64_t&>(d)) : "l"(reinterpret_cast<u64_t&>(a)), "l"(reinterpret_cast<u64_t&>(b)), "l"(reinterpret_cast<u64_t&>(c)));
    return d;
}
__device__ __forceinline__ float2 pk_add(float2 a, float2 b) {
    float2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<u64_t&>(d)) : "l"(reinterpret_cast<u64_t&>(a)), "l"(reinterpret_cast<u64_t&>(b)));
    return d;
}
__device__ __forceinline__ float2 pk_sub(float2 a, float2 b) {
    float2 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<u64_t&>(d)) : "l"(reinterpret_cast<u64_t&>(a)), "l"(reinterpret_cast<u64_t&>(b)));
    return d;
}
__device__ __forceinline__ float2 pk_mul(float2 a, float2 b) {
    float2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<u64_t&>(d)) : "l"(reinterpret_cast<u64_t&>(a)), "l"(reinterpret_cast<u64_t&>(b)));
    return d;
}

// 128-bit streaming accesses: data touched once goes around L1.
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_f4(float4* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

}  // namespace sis
