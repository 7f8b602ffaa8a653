// Memory-bound generator pieces around the modulated conv (all HBM-bound; roofline = bytes / HBM GB/s):
//   * ConstantInput repeat                        model.py:295-305
//   * Blur(4x4, pad (1,1)) + NoiseInjection + FusedLeakyReLU in ONE pass over the (2H+1)^2 transposed-conv
//     output                                      model.py:76-92, 260-262, 287-292, 336-342
//       algorithmic bytes: (B*C*(2H+1)^2 + B*C*(2H)^2)*4 + (2H)^2*4 + C*4   (+ the bf16 hi/lo operand planes
//       of the next conv when the tcgen05 path asks for them)
//   * ToRGB: 1x1 modulated conv (no demod) + bias + Upsample(skip) (upfirdn2d up=2, pad (2,1)) + add in ONE
//     pass                                        model.py:34-52, 345-364
//       algorithmic bytes: B*C*H^2*4 + B*3*(H/2)^2*4 + B*3*H^2*4
#include "common.cuh"
#include "kernels.h"
#include "torgb_common.cuh"

namespace sis {

__global__ void __launch_bounds__(256) const_input_kernel(float* __restrict__ out, const float* __restrict__ inp,
                                                          int64_t per_sample, int64_t total) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = inp[i % per_sample];
}

int launch_const_input(float* out, const float* inp, int64_t per_sample, int batch, cudaStream_t stream) {
    int64_t total = per_sample * batch;
    int grid = (int)(ceil_div64(total, 256) < 1184 ? ceil_div64(total, 256) : 1184);
    const_input_kernel<<<grid, 256, 0, stream>>>(out, inp, per_sample, total);
    SIS_CHECK_LAUNCH();
    return SIS_OK;
}

// ------------------------------------------------------------------------------------------ blur + act
// in: [planes = B*C][IH][IW] fp32 (IH = OH + 1: the transposed-conv output), out: [planes][OH][OW].
// out[y,x] = lrelu( sum_{ky,kx} kflip[ky][kx] * inpad[y+ky][x+kx] + nw*noise[b,y,x] + bias[c] ) * sqrt(2),
// inpad = in padded by 1 on every side.  FMA chain order = the reference kernel's (y outer, x inner).
constexpr int BTH = 16, BTW = 64;
constexpr int BIN_H = BTH + 3, BIN_W = BTW + 3, BIN_WP = BIN_W + 1;

__global__ void __launch_bounds__(256) blur_noise_act_kernel(BlurActArgs a) {
    __shared__ float sk[4][4];
    __shared__ float sx[BIN_H][BIN_WP];
    const int tid = threadIdx.x;
    if (tid < 16) sk[tid / 4][tid % 4] = a.blur_k[(3 - tid / 4) * 4 + (3 - tid % 4)];
    const int tiles_x = (a.OW + BTW - 1) / BTW, tiles_y = (a.OH + BTH - 1) / BTH;
    const int64_t tiles_per_plane = (int64_t)tiles_x * tiles_y;
    const int64_t total_tiles = tiles_per_plane * a.planes;
    const int tx = tid % (BTW / 4), ty = tid / (BTW / 4);
    for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int64_t plane = tile / tiles_per_plane;
        const int trem = (int)(tile - plane * tiles_per_plane);
        const int y0 = (trem / tiles_x) * BTH, x0 = (trem % tiles_x) * BTW;
        const float* src = a.in + plane * (int64_t)a.IH * a.IW;
        __syncthreads();
        for (int i = tid; i < BIN_H * BIN_W; i += 256) {
            int ry = i / BIN_W, rx = i - ry * BIN_W;
            int iy = y0 + ry - 1, ix = x0 + rx - 1;
            float v = 0.0f;
            if (iy >= 0 && ix >= 0 && iy < a.IH && ix < a.IW) v = __ldg(src + (int64_t)iy * a.IW + ix);
            sx[ry][rx] = v;
        }
        __syncthreads();
        const int oy = y0 + ty;
        if (oy < a.OH) {
            const int b = (int)(plane / a.C), c = (int)(plane % a.C);
            const float bias = a.bias ? a.bias[c] : 0.0f;
            const float* nz = a.noise ? a.noise + (int64_t)b * a.noise_bstride + (int64_t)oy * a.OW : nullptr;
            float res[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int rx = tx * 4 + j;
                float v = 0.0f;
#pragma unroll
                for (int ky = 0; ky < 4; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 4; ++kx) v = __fmaf_rn(sx[ty + ky][rx + kx], sk[ky][kx], v);
                const int ox = x0 + rx;
                if (ox < a.OW && a.act) {
                    if (nz) v = __fadd_rn(v, __fmul_rn(a.noise_w, nz[ox]));
                    v = __fadd_rn(v, bias);
                    v = lrelu_scale(v, 0.2f, 1.41421356237309504880f);
                }
                res[j] = v;
            }
            float* dst = a.out + (plane * a.OH + oy) * (int64_t)a.OW + x0 + tx * 4;
            const int ox = x0 + tx * 4;
            if (ox + 3 < a.OW && ((((uintptr_t)dst) & 15) == 0)) {
                *reinterpret_cast<float4*>(dst) = make_float4(res[0], res[1], res[2], res[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (ox + j < a.OW) dst[j] = res[j];
            }
        }
    }
}

int launch_blur_noise_act(const BlurActArgs& a, cudaStream_t stream) {
    int64_t tiles = (int64_t)ceil_div(a.OW, BTW) * ceil_div(a.OH, BTH) * a.planes;
    int64_t cap = (int64_t)kNumSMs * 8;
    int grid = (int)(tiles < cap ? tiles : cap);
    blur_noise_act_kernel<<<grid, 256, 0, stream>>>(a);
    SIS_CHECK_LAUNCH();
    return SIS_OK;
}

// ----------------------------------------------------------------------------------------------- ToRGB
// One thread = 4 consecutive pixels of one sample, one channel slice; 8 slices per block reduced through
// shared memory in fixed order.  x is read exactly once with 128-bit loads.
constexpr int RGB_SLICES = 8;

// blockDim = (32, slices): 8 channel slices per block for large maps, 32 for the 4^2..32^2 maps whose few blocks would
// otherwise walk 64 channels per lane serially (latency-bound: 25 us per launch).
__global__ void __launch_bounds__(1024) torgb_kernel(ToRgbArgs a) {
    extern __shared__ float smem[];
    float* swr = smem;                                  // [3][C]  scale*W
    float* red = smem + 3 * a.C;                        // [SLICES][3][4][32]
    const int lane = threadIdx.x, slice = threadIdx.y, nsl = blockDim.y;
    const int tid = slice * 32 + lane;
    for (int i = tid; i < 3 * a.C; i += 32 * nsl) swr[i] = a.w[i];
    __syncthreads();
    const int64_t hw = (int64_t)a.H * a.W;
    const int64_t quads_per_sample = hw / 4;
    const int64_t total_quads = quads_per_sample * a.batch;
    const int cps = (a.C + nsl - 1) / nsl;
    const int c_begin = slice * cps, c_end = min(a.C, c_begin + cps);

    for (int64_t q0 = (int64_t)blockIdx.x * 32; q0 < total_quads; q0 += (int64_t)gridDim.x * 32) {
        const int64_t q = q0 + lane;
        const bool valid = q < total_quads;
        const int b = valid ? (int)(q / quads_per_sample) : 0;
        const int64_t pix = valid ? (q - (int64_t)b * quads_per_sample) * 4 : 0;
        float acc[3][4] = {};
        if (valid) {
            const float* xb = a.x + ((int64_t)b * a.C) * hw + pix;
            const float* sb = a.s + (int64_t)b * a.C;
#pragma unroll 8
            for (int c = c_begin; c < c_end; ++c) {
                const float4 v = ld_stream_f4(reinterpret_cast<const float4*>(xb + (int64_t)c * hw));
                const float s = __ldg(sb + c);
                const float w0 = __fmul_rn(swr[c], s), w1 = __fmul_rn(swr[a.C + c], s), w2 = __fmul_rn(swr[2 * a.C + c], s);
                acc[0][0] = fmaf(w0, v.x, acc[0][0]); acc[0][1] = fmaf(w0, v.y, acc[0][1]);
                acc[0][2] = fmaf(w0, v.z, acc[0][2]); acc[0][3] = fmaf(w0, v.w, acc[0][3]);
                acc[1][0] = fmaf(w1, v.x, acc[1][0]); acc[1][1] = fmaf(w1, v.y, acc[1][1]);
                acc[1][2] = fmaf(w1, v.z, acc[1][2]); acc[1][3] = fmaf(w1, v.w, acc[1][3]);
                acc[2][0] = fmaf(w2, v.x, acc[2][0]); acc[2][1] = fmaf(w2, v.y, acc[2][1]);
                acc[2][2] = fmaf(w2, v.z, acc[2][2]); acc[2][3] = fmaf(w2, v.w, acc[2][3]);
            }
        }
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int p = 0; p < 4; ++p) red[((slice * 3 + j) * 4 + p) * 32 + lane] = acc[j][p];
        __syncthreads();
        // 96 (j, p, lane) outputs per... : 3*4*32 = 384 results, 256 threads -> strided
        for (int r = tid; r < 3 * 4 * 32; r += 32 * nsl) {
            const int ln = r % 32, p = (r / 32) % 4, j = r / 128;
            const int64_t qq = q0 + ln;
            if (qq >= total_quads) continue;
            float v = 0.0f;
            for (int s = 0; s < nsl; ++s) v += red[((s * 3 + j) * 4 + p) * 32 + ln];
            const int bb = (int)(qq / quads_per_sample);
            const int64_t px = (qq - (int64_t)bb * quads_per_sample) * 4 + p;
            const int y = (int)(px / a.W), x = (int)(px % a.W);
            v = __fadd_rn(v, a.bias[j]);
            if (a.skip) {
                // Upsample: upfirdn2d(skip, k*4, up=2, pad=(2,1)) (model.py:34-52): mid = o - 1
                const int SH = a.H / 2, SW = a.W / 2;
                const int mid_y = y - 1, mid_x = x - 1;
                const int iy0 = (mid_y < 0) ? -1 : (mid_y >> 1), ix0 = (mid_x < 0) ? -1 : (mid_x >> 1);
                const int ky0 = (iy0 + 1) * 2 - mid_y - 1, kx0 = (ix0 + 1) * 2 - mid_x - 1;
                const float* sp = a.skip + ((int64_t)bb * 3 + j) * SH * SW;
                float u = 0.0f;
#pragma unroll
                for (int yy = 0; yy < 2; ++yy)
#pragma unroll
                    for (int xx = 0; xx < 2; ++xx) {
                        const int iy = iy0 + yy, ix = ix0 + xx;
                        float sv = 0.0f;
                        if (iy >= 0 && ix >= 0 && iy < SH && ix < SW) sv = __ldg(sp + (int64_t)iy * SW + ix);
                        const int ky = ky0 + yy * 2, kx = kx0 + xx * 2;
                        u = __fmaf_rn(sv, a.up_k[(3 - ky) * 4 + (3 - kx)], u);
                    }
                v = __fadd_rn(v, u);
            }
            a.out[((int64_t)bb * 3 + j) * hw + px] = v;
        }
        __syncthreads();
    }
}

// Wide variant for large maps: one thread = 4 consecutive pixels x ALL channels (no channel slicing, no barrier in
// the main loop); a block reads 4 KB contiguous per channel plane; 128-bit loads and stores.
__global__ void __launch_bounds__(256) torgb_wide_kernel(ToRgbArgs a) {
    extern __shared__ float smem[];
    float* swr = smem;                                  // [3][C]  scale*W
    const int tid = threadIdx.x;
    for (int i = tid; i < 3 * a.C; i += 256) swr[i] = a.w[i];
    __syncthreads();
    const int64_t hw = (int64_t)a.H * a.W;
    const int64_t quads_per_sample = hw >> 2;
    const int64_t q = (int64_t)blockIdx.x * 256 + tid;
    if (q >= quads_per_sample * a.batch) return;
    const int b = (int)(q / quads_per_sample);
    const int64_t pix = (q - (int64_t)b * quads_per_sample) << 2;
    const float* xb = a.x + ((int64_t)b * a.C) * hw + pix;
    const float* sb = a.s + (int64_t)b * a.C;
    float acc[3][4] = {};
#pragma unroll 8
    for (int c = 0; c < a.C; ++c) {
        const float4 v = ld_stream_f4(reinterpret_cast<const float4*>(xb + (int64_t)c * hw));
        const float s = __ldg(sb + c);
        const float w0 = __fmul_rn(swr[c], s), w1 = __fmul_rn(swr[a.C + c], s), w2 = __fmul_rn(swr[2 * a.C + c], s);
        acc[0][0] = fmaf(w0, v.x, acc[0][0]); acc[0][1] = fmaf(w0, v.y, acc[0][1]);
        acc[0][2] = fmaf(w0, v.z, acc[0][2]); acc[0][3] = fmaf(w0, v.w, acc[0][3]);
        acc[1][0] = fmaf(w1, v.x, acc[1][0]); acc[1][1] = fmaf(w1, v.y, acc[1][1]);
        acc[1][2] = fmaf(w1, v.z, acc[1][2]); acc[1][3] = fmaf(w1, v.w, acc[1][3]);
        acc[2][0] = fmaf(w2, v.x, acc[2][0]); acc[2][1] = fmaf(w2, v.y, acc[2][1]);
        acc[2][2] = fmaf(w2, v.z, acc[2][2]); acc[2][3] = fmaf(w2, v.w, acc[2][3]);
    }
    const int y = (int)(pix / a.W), x = (int)(pix - (int64_t)y * a.W);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        float o[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            float v = __fadd_rn(acc[j][p], a.bias[j]);
            if (a.skip) v = __fadd_rn(v, torgb_skip_tap(a, a.skip + ((int64_t)b * 3 + j) * (hw >> 2), y, x + p));
            o[p] = v;
        }
        *reinterpret_cast<float4*>(a.out + ((int64_t)b * 3 + j) * hw + pix) = make_float4(o[0], o[1], o[2], o[3]);
    }
}

int launch_torgb(const ToRgbArgs& a, cudaStream_t stream) {
    SIS_REQUIRE((a.H * a.W) % 4 == 0, "torgb: H*W must be a multiple of 4");
    int64_t quads = (int64_t)a.H * a.W / 4 * a.batch;
    if (a.W % 4 == 0 && quads >= (int64_t)kNumSMs * 512 && 3 * a.C * sizeof(float) <= 48 * 1024) {
        torgb_wide_kernel<<<(unsigned)ceil_div64(quads, 256), 256, 3 * a.C * sizeof(float), stream>>>(a);
        SIS_CHECK_LAUNCH();
        return SIS_OK;
    }
    int64_t blocks = ceil_div64(quads, 32);
    int64_t cap = (int64_t)kNumSMs * 8;
    int grid = (int)(blocks < cap ? blocks : cap);
    const int slices = blocks < kNumSMs ? 32 : RGB_SLICES;
    size_t smem = (size_t)(3 * a.C + slices * 3 * 4 * 32) * sizeof(float);
    if (smem > 48 * 1024) {
        static bool configured = false;
        if (!configured) {
            SIS_CHECK_CUDA(cudaFuncSetAttribute(torgb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            configured = true;
        }
    }
    torgb_kernel<<<grid, dim3(32, slices), smem, stream>>>(a);
    SIS_CHECK_LAUNCH();
    return SIS_OK;
}

// ------------------------------------------------------------------------------------------ make_image
__global__ void __launch_bounds__(256) make_image_kernel(uint8_t* __restrict__ out, const float* __restrict__ img,
                                                         int64_t hw, int64_t total_px) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_px; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / hw, p = i - b * hw;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v = img[(b * 3 + c) * hw + p];
            v = fminf(fmaxf(v, -1.0f), 1.0f);
            v = __fmul_rn(__fdiv_rn(__fadd_rn(v, 1.0f), 2.0f), 255.0f);
            out[i * 3 + c] = (uint8_t)(int)v;  // truncation, as Tensor.type(uint8)
        }
    }
}

}  // namespace sis

using namespace sis;

extern "C" int sis_make_image_u8(const float* d_image, int batch, int size, uint8_t* d_out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIS_REQUIRE(batch >= 0 && size >= 0, "make_image: negative size");
    int64_t hw = (int64_t)size * size, total = hw * batch;
    if (total == 0) return SIS_OK;
    SIS_REQUIRE(d_image && d_out, "make_image: null pointer");
    int64_t blocks = ceil_div64(total, 256);
    int64_t cap = (int64_t)kNumSMs * 8;
    make_image_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, stream>>>(d_out, d_image, hw, total);
    SIS_CHECK_LAUNCH();
    return SIS_OK;
}
