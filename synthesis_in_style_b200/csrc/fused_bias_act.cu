// fused bias + activation (+ scale).  Replaces fused_bias_act_kernel
// (scf/networks/stylegan2/op/fused_bias_act_kernel.cu:18-98).
//
// HBM-bound: 2*N*sizeof(T) + C*sizeof(T) algorithmic bytes.  fp32 fast path: 128-bit streaming loads /
// stores, 4 independent vectors in flight per thread, one bias lookup per vector (a float4 never straddles
// a channel when step_b % 4 == 0), grid = a multiple of the SM count (grid-stride).
#include "common.cuh"

namespace sis {

template <typename T>
__device__ __forceinline__ T act_apply(T x, T ref, bool use_ref, int code, T alpha, T scale);

template <>
__device__ __forceinline__ float act_apply<float>(float x, float ref, bool use_ref, int code, float alpha,
                                                  float scale) {
    float y;
    switch (code) {
        case 12: case 32: y = 0.0f; break;
        case 30: y = (x > 0.0f) ? x : __fmul_rn(x, alpha); break;
        case 31: y = (ref > 0.0f) ? x : __fmul_rn(x, alpha); break;
        default: y = x; break;
    }
    return __fmul_rn(y, scale);
}
template <>
__device__ __forceinline__ double act_apply<double>(double x, double ref, bool use_ref, int code, double alpha,
                                                    double scale) {
    double y;
    switch (code) {
        case 12: case 32: y = 0.0; break;
        case 30: y = (x > 0.0) ? x : x * alpha; break;
        case 31: y = (ref > 0.0) ? x : x * alpha; break;
        default: y = x; break;
    }
    return y * scale;
}
template <>
__device__ __forceinline__ __half act_apply<__half>(__half x, __half ref, bool use_ref, int code, __half alpha,
                                                    __half scale) {
    // the reference instantiates the kernel with scalar_t = c10::Half: every op rounds to half.
    __half zero = __float2half(0.0f);
    __half y;
    switch (code) {
        case 12: case 32: y = zero; break;
        case 30: y = __hgt(x, zero) ? x : __hmul(x, alpha); break;
        case 31: y = __hgt(ref, zero) ? x : __hmul(x, alpha); break;
        default: y = x; break;
    }
    return __hmul(y, scale);
}

template <typename T> __device__ __forceinline__ T add_t(T a, T b) { return a + b; }
template <> __device__ __forceinline__ float add_t<float>(float a, float b) { return __fadd_rn(a, b); }
template <> __device__ __forceinline__ __half add_t<__half>(__half a, __half b) { return __hadd(a, b); }

template <typename T>
__global__ void __launch_bounds__(256) fused_bias_act_generic_kernel(
    T* __restrict__ out, const T* __restrict__ x, const T* __restrict__ b, const T* __restrict__ ref, int code,
    T alpha, T scale, int64_t size_x, int64_t step_b, int64_t size_b) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < size_x; i += stride) {
        T v = x[i];
        if (size_b) v = add_t<T>(v, b[(i / step_b) % size_b]);
        T r = ref ? ref[i] : v;
        out[i] = act_apply<T>(v, r, ref != nullptr, code, alpha, scale);
    }
}

constexpr int kVecPerThread = 4;

__global__ void __launch_bounds__(256) fused_bias_act_f32x4_kernel(
    float4* __restrict__ out, const float4* __restrict__ x, const float* __restrict__ b,
    const float4* __restrict__ ref, int code, float alpha, float scale, int64_t n_vec, int64_t step_b_vec,
    int64_t size_b) {
    const int64_t tile = (int64_t)blockDim.x * kVecPerThread;
    for (int64_t base = (int64_t)blockIdx.x * tile; base < n_vec; base += (int64_t)gridDim.x * tile) {
        float4 v[kVecPerThread];
        float4 r[kVecPerThread];
        float bias[kVecPerThread];
#pragma unroll
        for (int j = 0; j < kVecPerThread; ++j) {
            int64_t i = base + (int64_t)j * blockDim.x + threadIdx.x;
            if (i < n_vec) {
                v[j] = ld_stream_f4(x + i);
                if (ref) r[j] = ld_stream_f4(ref + i);
                bias[j] = size_b ? __ldg(b + (i / step_b_vec) % size_b) : 0.0f;
            }
        }
#pragma unroll
        for (int j = 0; j < kVecPerThread; ++j) {
            int64_t i = base + (int64_t)j * blockDim.x + threadIdx.x;
            if (i < n_vec) {
                float4 a = v[j];
                if (size_b) {
                    a.x = __fadd_rn(a.x, bias[j]); a.y = __fadd_rn(a.y, bias[j]);
                    a.z = __fadd_rn(a.z, bias[j]); a.w = __fadd_rn(a.w, bias[j]);
                }
                float4 rr = ref ? r[j] : a;
                float4 o;
                o.x = act_apply<float>(a.x, rr.x, ref != nullptr, code, alpha, scale);
                o.y = act_apply<float>(a.y, rr.y, ref != nullptr, code, alpha, scale);
                o.z = act_apply<float>(a.z, rr.z, ref != nullptr, code, alpha, scale);
                o.w = act_apply<float>(a.w, rr.w, ref != nullptr, code, alpha, scale);
                st_stream_f4(out + i, o);
            }
        }
    }
}

static int grid_for(int64_t work_items, int64_t per_block) {
    int64_t blocks = ceil_div64(work_items, per_block);
    int64_t cap = (int64_t)kNumSMs * 8;  // 8 resident 256-thread blocks per SM
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace sis

using namespace sis;

extern "C" int sis_fused_bias_act(void* d_out, const void* d_x, const void* d_bias, const void* d_ref, int dtype,
                                  int64_t size_x, int64_t step_b, int64_t size_b, int act, int grad, float alpha,
                                  float scale, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIS_REQUIRE(size_x >= 0 && size_b >= 0, "fused_bias_act: negative size");
    if (size_x == 0) return SIS_OK;
    SIS_REQUIRE(d_out && d_x, "fused_bias_act: input must be a CUDA tensor (null pointer)");
    SIS_REQUIRE(size_b == 0 || d_bias, "fused_bias_act: bias must be a CUDA tensor (null pointer)");
    SIS_REQUIRE(step_b >= 1, "fused_bias_act: step_b must be >= 1");
    const int code = act * 10 + grad;
    if (dtype == SIS_F32) {
        const bool aligned = (((uintptr_t)d_out | (uintptr_t)d_x | (uintptr_t)(d_ref ? d_ref : d_x)) & 15) == 0;
        if (aligned && size_x % 4 == 0 && (size_b == 0 || step_b % 4 == 0)) {
            int64_t n_vec = size_x / 4;
            int grid = grid_for(n_vec, 256 * kVecPerThread);
            fused_bias_act_f32x4_kernel<<<grid, 256, 0, stream>>>(
                (float4*)d_out, (const float4*)d_x, (const float*)d_bias, (const float4*)d_ref, code, alpha, scale,
                n_vec, size_b ? step_b / 4 : 1, size_b);
        } else {
            int grid = grid_for(size_x, 256);
            fused_bias_act_generic_kernel<float><<<grid, 256, 0, stream>>>(
                (float*)d_out, (const float*)d_x, (const float*)d_bias, (const float*)d_ref, code, alpha, scale,
                size_x, step_b, size_b);
        }
    } else if (dtype == SIS_F64) {
        int grid = grid_for(size_x, 256);
        fused_bias_act_generic_kernel<double><<<grid, 256, 0, stream>>>(
            (double*)d_out, (const double*)d_x, (const double*)d_bias, (const double*)d_ref, code, (double)alpha,
            (double)scale, size_x, step_b, size_b);
    } else if (dtype == SIS_F16) {
        int grid = grid_for(size_x, 256);
        fused_bias_act_generic_kernel<__half><<<grid, 256, 0, stream>>>(
            (__half*)d_out, (const __half*)d_x, (const __half*)d_bias, (const __half*)d_ref, code,
            __float2half(alpha), __float2half(scale), size_x, step_b, size_b);
    } else {
        set_error("fused_bias_act: unsupported dtype %d", dtype);
        return SIS_ERR_UNSUPPORTED;
    }
    SIS_CHECK_LAUNCH();
    return SIS_OK;
}
