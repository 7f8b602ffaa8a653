// Small dense pieces of the generator: PixelNorm, the style MLP, latent assembly (truncation / style mixing),
// the per-layer modulation linears and the demodulation factors.
//   PixelNorm           scf/networks/stylegan2/model.py:15-20
//   EqualLinear         model.py:133-162      (scale folded into the weights at prepare time, as the
//                                              reference does per call: `self.weight * self.scale`)
//   truncation / mixing model.py:502-528
//   modulation          model.py:227,240      s = EqualLinear(style_dim -> Cin, bias_init 1)(w)
//   demodulation        model.py:243-245      d[b,o] = rsqrt(sum_{i,ky,kx} (scale*W*s)^2 + 1e-8)
//                                             = rsqrt(sum_i s[b,i]^2 * Wsq[o,i] + 1e-8),  Wsq = sum_k (scale*W)^2
// All of this is < 0.02 % of the generator's FLOPs: fp32 FMA, latency-bound, one
// batched launch per stage.
#include "common.cuh"
#include "kernels.h"

namespace sis {

__global__ void __launch_bounds__(128) pixel_norm_kernel(float* __restrict__ out, const float* __restrict__ z,
                                                         int dim) {
    // one block per row: z * rsqrt(mean(z^2) + 1e-8)
    const float* row = z + (int64_t)blockIdx.x * dim;
    float acc = 0.0f;
    for (int i = threadIdx.x; i < dim; i += blockDim.x) acc += row[i] * row[i];
    __shared__ float red[4];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    float tot = red[0] + red[1] + red[2] + red[3];
    float r = rsqrtf(tot / (float)dim + 1e-8f);
    for (int i = threadIdx.x; i < dim; i += blockDim.x) out[(int64_t)blockIdx.x * dim + i] = row[i] * r;
}

// C[m, n] = epi( sum_k A'[m,k] * W[n,k] ), A' = A or A*A.  Batched over blockIdx.z through `jobs`.
// Tile 32 (m) x 64 (n) x 32 (k), 256 threads, 2x4 outputs per thread.
constexpr int LBM = 32, LBN = 64, LBK = 32;

__global__ void __launch_bounds__(256) linear_nt_kernel(const LinearJob* __restrict__ jobs) {
    const LinearJob job = jobs[blockIdx.z];
    const int m0 = blockIdx.y * LBM, n0 = blockIdx.x * LBN;
    if (m0 >= job.M || n0 >= job.N) return;
    __shared__ float sa[LBK][LBM + 1];
    __shared__ float sw[LBK][LBN + 1];
    const int tid = threadIdx.x;
    const int tm = (tid / 16) * 2;  // 16 row groups of 2
    const int tn = (tid % 16) * 4;  // 16 col groups of 4
    float acc[2][4] = {};
    for (int k0 = 0; k0 < job.K; k0 += LBK) {
        for (int i = tid; i < LBM * LBK; i += 256) {
            int r = i / LBK, c = i % LBK;
            float v = 0.0f;
            if (m0 + r < job.M && k0 + c < job.K) {
                v = job.A[(int64_t)(m0 + r) * job.lda + k0 + c];
                if (job.square_a) v = v * v;
            }
            sa[c][r] = v;
        }
        for (int i = tid; i < LBN * LBK; i += 256) {
            int r = i / LBK, c = i % LBK;
            float v = 0.0f;
            if (n0 + r < job.N && k0 + c < job.K) v = job.W[(int64_t)(n0 + r) * job.K + k0 + c];
            sw[c][r] = v;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < LBK; ++k) {
            float a0 = sa[k][tm], a1 = sa[k][tm + 1];
            float w0 = sw[k][tn], w1 = sw[k][tn + 1], w2 = sw[k][tn + 2], w3 = sw[k][tn + 3];
            acc[0][0] = fmaf(a0, w0, acc[0][0]); acc[0][1] = fmaf(a0, w1, acc[0][1]);
            acc[0][2] = fmaf(a0, w2, acc[0][2]); acc[0][3] = fmaf(a0, w3, acc[0][3]);
            acc[1][0] = fmaf(a1, w0, acc[1][0]); acc[1][1] = fmaf(a1, w1, acc[1][1]);
            acc[1][2] = fmaf(a1, w2, acc[1][2]); acc[1][3] = fmaf(a1, w3, acc[1][3]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        int m = m0 + tm + i;
        if (m >= job.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tn + j;
            if (n >= job.N) continue;
            float v = acc[i][j];
            if (job.epilogue == LINEAR_EPI_BIAS) {
                if (job.bias) v = __fadd_rn(v, job.bias[n]);
            } else if (job.epilogue == LINEAR_EPI_BIAS_LRELU) {
                if (job.bias) v = __fadd_rn(v, job.bias[n]);
                v = lrelu_scale(v, 0.2f, 1.41421356237309504880f);
            } else {  // LINEAR_EPI_RSQRT_EPS
                v = rsqrtf(v + 1e-8f);
            }
            job.C[(int64_t)m * job.ldc + n] = v;
        }
    }
}

// Small-M variant (M <= 32 rows per block, K % 4 == 0): the GEMMs of this path have M = batch (32) and N, K <= 512, so
// the work is streaming the weight matrix once.  One warp = one output column: its weight row is read with
// coalesced 128-bit loads, the activation tile [32][K] sits in shared memory, every lane keeps 32 row accumulators
// and the 32x32 (row, lane) partials are transposed through shared memory so that lane m finishes row m.
constexpr int SM_COLS = 8;      // output columns (= warps) per block
constexpr int SM_ROWS = 32;     // rows per block

__global__ void __launch_bounds__(32 * SM_COLS) linear_small_m_kernel(const LinearJob* __restrict__ jobs) {
    const LinearJob job = jobs[blockIdx.z];
    const int m0 = blockIdx.y * SM_ROWS, n0 = blockIdx.x * SM_COLS;
    if (m0 >= job.M || n0 >= job.N) return;
    extern __shared__ __align__(16) float smem_lin[];
    float* xs = smem_lin;                                  // [32][K]
    float* red = smem_lin + (size_t)SM_ROWS * job.K;       // [SM_COLS][32][33]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = job.K, K4 = K >> 2;
    for (int i = tid; i < SM_ROWS * K4; i += 32 * SM_COLS) {
        const int r = i / K4, c4 = i - r * K4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m0 + r < job.M) {
            const float* src = job.A + (int64_t)(m0 + r) * job.lda + c4 * 4;
            if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) v = __ldg(reinterpret_cast<const float4*>(src));   // one 128-bit load
            else v = make_float4(src[0], src[1], src[2], src[3]);
            if (job.square_a) { v.x *= v.x; v.y *= v.y; v.z *= v.z; v.w *= v.w; }
        }
        reinterpret_cast<float4*>(xs)[i] = v;
    }
    __syncthreads();
    const int n = n0 + warp;
    float acc[SM_ROWS];
#pragma unroll
    for (int m = 0; m < SM_ROWS; ++m) acc[m] = 0.0f;
    if (n < job.N) {
        const float4* wrow = reinterpret_cast<const float4*>(job.W + (int64_t)n * K);
        for (int c4 = lane; c4 < K4; c4 += 32) {
            const float4 w = __ldg(wrow + c4);
#pragma unroll
            for (int m = 0; m < SM_ROWS; ++m) {
                const float4 x = reinterpret_cast<const float4*>(xs)[m * K4 + c4];
                acc[m] = fmaf(w.x, x.x, acc[m]); acc[m] = fmaf(w.y, x.y, acc[m]);
                acc[m] = fmaf(w.z, x.z, acc[m]); acc[m] = fmaf(w.w, x.w, acc[m]);
            }
        }
    }
    float* myred = red + (size_t)warp * 32 * 33;
#pragma unroll
    for (int m = 0; m < SM_ROWS; ++m) myred[m * 33 + lane] = acc[m];
    __syncwarp();
    if (n < job.N && m0 + lane < job.M) {
        float v = 0.0f;
#pragma unroll
        for (int j = 0; j < 32; ++j) v += myred[lane * 33 + j];
        if (job.epilogue == LINEAR_EPI_BIAS) {
            if (job.bias) v = __fadd_rn(v, job.bias[n]);
        } else if (job.epilogue == LINEAR_EPI_BIAS_LRELU) {
            if (job.bias) v = __fadd_rn(v, job.bias[n]);
            v = lrelu_scale(v, 0.2f, 1.41421356237309504880f);
        } else {
            v = rsqrtf(v + 1e-8f);
        }
        job.C[(int64_t)(m0 + lane) * job.ldc + n] = v;
    }
}

// latent[b, l, :] (model.py:502-528).  truncation uses the reference's separately rounded sub, mul, add.
__global__ void __launch_bounds__(128) assemble_latent_kernel(float* __restrict__ latent, const float* __restrict__ w0,
                                                              const float* __restrict__ w1, int wplus, int inject_index,
                                                              float truncation, const float* __restrict__ tlat,
                                                              int tlat_rows, int n_latent, int dim) {
    const int b = blockIdx.y, l = blockIdx.x;
    const float* src;
    if (wplus) src = w0 + ((int64_t)b * n_latent + l) * dim;
    else src = ((w1 == nullptr || l < inject_index) ? w0 : w1) + (int64_t)b * dim;
    float* dst = latent + ((int64_t)b * n_latent + l) * dim;
    const float* t = tlat ? tlat + (tlat_rows > 1 ? (int64_t)b * dim : 0) : nullptr;
    for (int i = threadIdx.x; i < dim; i += blockDim.x) {
        float v = src[i];
        if (truncation < 1.0f && t) v = __fadd_rn(t[i], __fmul_rn(truncation, __fsub_rn(v, t[i])));
        dst[i] = v;
    }
}

// per-tensor weight preparation ---------------------------------------------------------------------------
__global__ void scale_copy_kernel(float* __restrict__ out, const float* __restrict__ in, float scale, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = __fmul_rn(in[i], scale);
}

// Wsq[o,i] = sum_t (scale*W[o,i,t])^2
__global__ void weight_sq_kernel(float* __restrict__ wsq, const float* __restrict__ w, float scale, int64_t n_oi,
                                 int taps) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_oi; i += (int64_t)gridDim.x * blockDim.x) {
        float acc = 0.0f;
        for (int t = 0; t < taps; ++t) {
            float v = __fmul_rn(w[i * taps + t], scale);
            acc = fmaf(v, v, acc);
        }
        wsq[i] = acc;
    }
}

// out[o,i,ky,kx] = scale * w[o,i,2-ky,2-kx] : correlation taps of the stride-2 transposed conv seen as a
// pad-2 correlation over the zero-inserted input (F.conv_transpose2d, model.py:259).
__global__ void scale_flip3x3_kernel(float* __restrict__ out, const float* __restrict__ in, float scale, int64_t n_oi) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_oi * 9; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t oi = i / 9;
        int t = (int)(i - oi * 9);
        out[i] = __fmul_rn(in[oi * 9 + (8 - t)], scale);
    }
}

int launch_pixel_norm(float* out, const float* z, int64_t rows, int dim, cudaStream_t stream) {
    if (rows == 0) return SIS_OK;
    pixel_norm_kernel<<<(unsigned)rows, 128, 0, stream>>>(out, z, dim);
    SIS_CHECK_LAUNCH();
    return SIS_OK;
}

int launch_linear_jobs(const LinearJob* d_jobs, int n_jobs, int max_m, int max_n, int max_k, bool small_m_ok, cudaStream_t stream) {
    if (n_jobs == 0 || max_m == 0 || max_n == 0) return SIS_OK;
    const size_t smem = ((size_t)SM_ROWS * max_k + (size_t)SM_COLS * 32 * 33) * sizeof(float);
    if (small_m_ok && smem <= 200 * 1024) {
        static bool configured = false;
        if (!configured) {
            SIS_CHECK_CUDA(cudaFuncSetAttribute(linear_small_m_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            configured = true;
        }
        dim3 grid(ceil_div(max_n, SM_COLS), ceil_div(max_m, SM_ROWS), n_jobs);
        linear_small_m_kernel<<<grid, 32 * SM_COLS, smem, stream>>>(d_jobs);
    } else {
        dim3 grid(ceil_div(max_n, LBN), ceil_div(max_m, LBM), n_jobs);
        linear_nt_kernel<<<grid, 256, 0, stream>>>(d_jobs);
    }
    SIS_CHECK_LAUNCH();
    return SIS_OK;
}

int launch_assemble_latent(float* latent, const float* w0, const float* w1, int wplus, int inject_index,
                           float truncation, const float* tlat, int tlat_rows, int batch, int n_latent, int dim,
                           cudaStream_t stream) {
    dim3 grid(n_latent, batch);
    assemble_latent_kernel<<<grid, 128, 0, stream>>>(latent, w0, w1, wplus, inject_index, truncation, tlat, tlat_rows,
                                                     n_latent, dim);
    SIS_CHECK_LAUNCH();
    return SIS_OK;
}

static int small_grid(int64_t n) {
    int64_t b = ceil_div64(n, 256);
    int64_t cap = (int64_t)kNumSMs * 8;
    return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

int launch_scale_copy(float* out, const float* in, float scale, int64_t n, cudaStream_t stream) {
    scale_copy_kernel<<<small_grid(n), 256, 0, stream>>>(out, in, scale, n);
    SIS_CHECK_LAUNCH();
    return SIS_OK;
}
int launch_weight_sq(float* wsq, const float* w, float scale, int64_t n_oi, int taps, cudaStream_t stream) {
    weight_sq_kernel<<<small_grid(n_oi), 256, 0, stream>>>(wsq, w, scale, n_oi, taps);
    SIS_CHECK_LAUNCH();
    return SIS_OK;
}
int launch_scale_flip3x3(float* out, const float* in, float scale, int64_t n_oi, cudaStream_t stream) {
    scale_flip3x3_kernel<<<small_grid(n_oi * 9), 256, 0, stream>>>(out, in, scale, n_oi);
    SIS_CHECK_LAUNCH();
    return SIS_OK;
}

}  // namespace sis
