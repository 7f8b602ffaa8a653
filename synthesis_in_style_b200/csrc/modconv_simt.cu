// Modulated 3x3 convolution, fp32 CUDA-core path (sis_precision FP32).
//   ModulatedConv2d.forward  scf/networks/stylegan2/model.py:237-278
// The reference materialises per-sample weights (scale*W*s*demod) and runs a grouped cuDNN conv.  Here the
// algebraically equal shared-weight form is used (SURVEY.md §7 "Per-sample weights"):
//     y[b,o,p] = d[b,o] * sum_{i,t} (scale*W[o,i,t]) * (s[b,i] * x[b,i,p+t])
// The stride-2 transposed conv (model.py:251-259) is the same kernel run as a pad-2 correlation with flipped
// taps over the zero-inserted input (virtual size 2H-1, output 2H+1).
// This path is the exact-precision fallback / debugging reference for the tcgen05 path; it is tiled for reuse
// (64 output channels x 16x16 pixels per block, 8x8 register tile) but is FP32-FMA bound by design.
#include "common.cuh"
#include "kernels.h"

namespace sis {

constexpr int TO = 64, TH = 16, TW = 16, CI = 8;
constexpr int MAXPAD = 2;
constexpr int IN_H = TH + 2 * MAXPAD, IN_W = TW + 2 * MAXPAD, IN_WP = IN_W + 1;

__global__ void __launch_bounds__(256) modconv3x3_simt_kernel(ModConvSimtArgs a) {
    __shared__ float sx[CI][IN_H][IN_WP];
    __shared__ __align__(16) float sw[CI][9][TO];

    const int tiles_x = (a.OW + TW - 1) / TW;
    const int tile_y0 = (blockIdx.x / tiles_x) * TH, tile_x0 = (blockIdx.x % tiles_x) * TW;
    const int o0 = blockIdx.y * TO;
    const int b = blockIdx.z;
    const int tid = threadIdx.x;
    const int to = tid >> 5;           // warp id -> 8 output channels
    const int pg = tid & 31;
    const int row = pg >> 1, col0 = (pg & 1) * 8;
    const int span = 2 * a.pad + 1;    // taps are always 3x3; pad 1 (plain) or 2 (transposed)
    (void)span;
    const int in_rows = TH + 2, in_cols = TW + 2;
    const int VH = a.zero_insert ? 2 * a.H - 1 : a.H, VW = a.zero_insert ? 2 * a.W - 1 : a.W;

    float acc[8][8];
#pragma unroll
    for (int c = 0; c < 8; ++c)
#pragma unroll
        for (int p = 0; p < 8; ++p) acc[c][p] = 0.0f;

    const float* xb = a.x + (int64_t)b * a.Cin * a.H * a.W;
    const float* sb = a.s + (int64_t)b * a.Cin;

    for (int c0 = 0; c0 < a.Cin; c0 += CI) {
        __syncthreads();
        // input tile (virtual coords), scaled by the style
        for (int i = tid; i < CI * in_rows * in_cols; i += 256) {
            int ci = i / (in_rows * in_cols);
            int r = i - ci * (in_rows * in_cols);
            int ry = r / in_cols, rx = r - ry * in_cols;
            int vy = tile_y0 + ry - a.pad, vx = tile_x0 + rx - a.pad;
            float v = 0.0f;
            if (c0 + ci < a.Cin && vy >= 0 && vx >= 0 && vy < VH && vx < VW) {
                if (a.zero_insert) {
                    if (((vy | vx) & 1) == 0)
                        v = xb[((int64_t)(c0 + ci) * a.H + (vy >> 1)) * a.W + (vx >> 1)];
                } else {
                    v = xb[((int64_t)(c0 + ci) * a.H + vy) * a.W + vx];
                }
                v *= sb[c0 + ci];
            }
            sx[ci][ry][rx] = v;
        }
        // weights [Cout][Cin][9] -> sw[ci][tap][o]
        for (int i = tid; i < TO * CI * 9; i += 256) {
            int o = i / (CI * 9);
            int r = i - o * (CI * 9);
            int ci = r / 9, t = r - ci * 9;
            float v = 0.0f;
            if (o0 + o < a.Cout && c0 + ci < a.Cin) v = a.w[((int64_t)(o0 + o) * a.Cin + c0 + ci) * 9 + t];
            sw[ci][t][o] = v;
        }
        __syncthreads();
#pragma unroll 1
        for (int ci = 0; ci < CI; ++ci) {
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                float in[10];
#pragma unroll
                for (int j = 0; j < 10; ++j) in[j] = sx[ci][row + ky][col0 + j];
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float4 w0 = *reinterpret_cast<const float4*>(&sw[ci][ky * 3 + kx][to * 8]);
                    const float4 w1 = *reinterpret_cast<const float4*>(&sw[ci][ky * 3 + kx][to * 8 + 4]);
                    const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                    for (int c = 0; c < 8; ++c)
#pragma unroll
                        for (int p = 0; p < 8; ++p) acc[c][p] = fmaf(wv[c], in[p + kx], acc[c][p]);
                }
            }
        }
    }

    const int oy = tile_y0 + row;
    if (oy >= a.OH) return;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int o = o0 + to * 8 + c;
        if (o >= a.Cout) continue;
        const float d = a.d ? a.d[(int64_t)b * a.Cout + o] : 1.0f;
        const float bias = (a.fuse_act && a.bias) ? a.bias[o] : 0.0f;
        float* dst = a.out + (((int64_t)b * a.Cout + o) * a.OH + oy) * a.OW;
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            const int ox = tile_x0 + col0 + p;
            if (ox >= a.OW) continue;
            float v = __fmul_rn(acc[c][p], d);
            if (a.fuse_act) {
                // NoiseInjection (model.py:292) then FusedLeakyReLU (fused_bias_act_kernel.cu:26-47)
                if (a.noise) v = __fadd_rn(v, __fmul_rn(a.noise_w, a.noise[(int64_t)b * a.noise_bstride + (int64_t)oy * a.OW + ox]));
                v = __fadd_rn(v, bias);
                v = lrelu_scale(v, 0.2f, 1.41421356237309504880f);
            }
            dst[ox] = v;
        }
    }
}

int launch_modconv3x3_simt(const ModConvSimtArgs& a, int batch, cudaStream_t stream) {
    SIS_REQUIRE(a.pad == 1 || a.pad == 2, "modconv simt: pad must be 1 or 2");
    dim3 grid(ceil_div(a.OW, TW) * ceil_div(a.OH, TH), ceil_div(a.Cout, TO), batch);
    modconv3x3_simt_kernel<<<grid, 256, 0, stream>>>(a);
    SIS_CHECK_LAUNCH();
    return SIS_OK;
}

}  // namespace sis
