// Shared by the ToRGB kernels (epilogue.cu) and the fused ToRGB + labelling kernel (labelling.cu).
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace sis {

// One output pixel of Upsample(skip): upfirdn2d(skip, k*4, up=2, pad=(2,1)) (model.py:34-52): mid = o - 1, 2x2 taps,
// FMA chain in the reference kernel's order.  `sp` = plane [H/2][W/2] of the previous skip image.
__device__ __forceinline__ float torgb_skip_tap(const ToRgbArgs& a, const float* sp, int y, int x) {
    const int SH = a.H / 2, SW = a.W / 2;
    const int mid_y = y - 1, mid_x = x - 1;
    const int iy0 = (mid_y < 0) ? -1 : (mid_y >> 1), ix0 = (mid_x < 0) ? -1 : (mid_x >> 1);
    const int ky0 = (iy0 + 1) * 2 - mid_y - 1, kx0 = (ix0 + 1) * 2 - mid_x - 1;
    float u = 0.0f;
#pragma unroll
    for (int yy = 0; yy < 2; ++yy)
#pragma unroll
        for (int xx = 0; xx < 2; ++xx) {
            const int iy = iy0 + yy, ix = ix0 + xx;
            float sv = 0.0f;
            if (iy >= 0 && ix >= 0 && iy < SH && ix < SW) sv = __ldg(sp + (int64_t)iy * SW + ix);
            const int ky = ky0 + yy * 2, kx = kx0 + xx * 2;
            u = __fmaf_rn(sv, __ldg(a.up_k + (3 - ky) * 4 + (3 - kx)), u);
        }
    return u;
}

}  // namespace sis
