// upfirdn2d: zero-insert upsample -> pad -> 2-D FIR (true convolution, flipped taps) -> decimate.
// Replaces upfirdn2d_kernel / upfirdn2d_op (scf/networks/stylegan2/op/upfirdn2d_kernel.cu:52-272).
//
// Index math per output (same as the reference kernel, :112-121):
//   mid = o*down + up - 1 - pad0 ; in0 = floor(mid / up) ; k0 = (in0+1)*up - mid - 1
//   v = sum_{y} sum_{x} in[in0y + y][in0x + x] * kflip[k0y + y*up][k0x + x*up]      (y outer, x inner,
//   accumulated as an FMA chain from 0, which is what nvcc makes of the reference's `v += a*b`).
//
// Two paths, both sm_100a CUDA:
//   * tiled fp32 path for minor == 1 and the reference's six (up, down, K) modes: per-plane output tile
//     32x64, input tile + halo staged in shared memory with coalesced row loads, 4 consecutive outputs per
//     thread with 128-bit stores.  HBM-bound: (N_in + N_out)*4 algorithmic bytes.
//   * generic path for every other (dtype, minor, up, down, kernel <= 32x32): one thread per output.
#include "common.cuh"

namespace sis {

__host__ __device__ __forceinline__ int floor_div(int a, int b) {
    int c = a / b;
    if (c * b > a) c--;
    return c;
}

struct UpfirdnParams {
    int up_x, up_y, down_x, down_y, pad_x0, pad_y0;
    int in_h, in_w, minor, kernel_h, kernel_w, out_h, out_w;
    int64_t major;
};

constexpr int kMaxTaps = 32;

// ---------------------------------------------------------------------------------------------- generic
template <typename T> struct Acc;
template <> struct Acc<float> {
    using type = float;
    static __device__ __forceinline__ float mac(float v, float a, float k) { return __fmaf_rn(a, k, v); }
};
template <> struct Acc<double> {
    // the reference stages input and taps in `volatile float` shared memory (upfirdn2d_kernel.cu:56-57):
    // products are float, the running sum is double.
    using type = double;
    static __device__ __forceinline__ double mac(double v, float a, float k) { return v + (double)__fmul_rn(a, k); }
};
template <> struct Acc<__half> {
    using type = __half;
    static __device__ __forceinline__ __half mac(__half v, float a, float k) {
        __half p = __float2half_rn(__fmul_rn(a, k));
        return __float2half_rn(__fadd_rn(__half2float(v), __half2float(p)));
    }
};
template <typename T> __device__ __forceinline__ float to_f(T v) { return (float)v; }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T zero_of() { return (T)0; }
template <> __device__ __forceinline__ __half zero_of<__half>() { return __float2half(0.0f); }

template <typename T>
__global__ void __launch_bounds__(256) upfirdn2d_generic_kernel(T* __restrict__ out, const T* __restrict__ in,
                                                                const T* __restrict__ kernel, UpfirdnParams p) {
    __shared__ float sk[kMaxTaps * kMaxTaps];
    for (int t = threadIdx.x; t < p.kernel_h * p.kernel_w; t += blockDim.x) {
        int ky = t / p.kernel_w, kx = t - ky * p.kernel_w;
        sk[t] = to_f<T>(kernel[(p.kernel_h - 1 - ky) * p.kernel_w + (p.kernel_w - 1 - kx)]);  // flipped
    }
    __syncthreads();
    const int64_t total = p.major * p.out_h * p.out_w * p.minor;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        int m = (int)(idx % p.minor);
        int64_t r = idx / p.minor;
        int ox = (int)(r % p.out_w); r /= p.out_w;
        int oy = (int)(r % p.out_h);
        int64_t mj = r / p.out_h;
        int mid_x = ox * p.down_x + p.up_x - 1 - p.pad_x0;
        int mid_y = oy * p.down_y + p.up_y - 1 - p.pad_y0;
        int in_x0 = floor_div(mid_x, p.up_x), in_y0 = floor_div(mid_y, p.up_y);
        int kx0 = (in_x0 + 1) * p.up_x - mid_x - 1, ky0 = (in_y0 + 1) * p.up_y - mid_y - 1;
        typename Acc<T>::type v = zero_of<T>();
        for (int y = 0, ky = ky0; ky < p.kernel_h; ++y, ky += p.up_y) {
            int iy = in_y0 + y;
            for (int x = 0, kx = kx0; kx < p.kernel_w; ++x, kx += p.up_x) {
                int ix = in_x0 + x;
                float a = 0.0f;
                if (ix >= 0 && iy >= 0 && ix < p.in_w && iy < p.in_h)
                    a = to_f<T>(in[((mj * p.in_h + iy) * p.in_w + ix) * p.minor + m]);
                v = Acc<T>::mac(v, a, sk[ky * p.kernel_w + kx]);
            }
        }
        out[idx] = v;
    }
}

// ------------------------------------------------------------------------------------------ small planes
// Planes of at most 16 x 16 outputs (the 4^2 ... 16^2 maps; minor == 1, fp32, any up / down / taps).  A 64-wide tile per
// plane pair would leave most of a block idle there, and one thread per output straight from global memory is bound by
// its 64-bit index arithmetic and per-tap bounds checks.  Here a block owns PB consecutive
// planes: their inputs are ONE contiguous span of global memory (coalesced loads) scattered into zero-padded planes in
// shared memory, so the tap loops carry no bounds checks and all index arithmetic is 32-bit; the outputs of the PB planes
// are one contiguous span again.  Same FMA chain per output as the generic kernel (real taps only, y outer, x inner).
struct SmallPlaneGeom {
    int px0, py0, pw, ph;        // padded input window: columns px0 .. px0+pw-1, rows py0 .. py0+ph-1 (zero outside the plane)
    int planes_per_block;
};

__global__ void __launch_bounds__(256) upfirdn2d_small_plane_kernel(float* __restrict__ out, const float* __restrict__ in,
                                                                   const float* __restrict__ kernel, UpfirdnParams p,
                                                                   SmallPlaneGeom g) {
    extern __shared__ float sp_smem[];
    float* sk = sp_smem;                                   // [kernel_h][kernel_w], flipped
    float* sin = sp_smem + p.kernel_h * p.kernel_w;        // [PB][ph][pw]
    const int tid = threadIdx.x;
    const int plane_in = p.in_h * p.in_w, plane_pad = g.ph * g.pw, plane_out = p.out_h * p.out_w;
    for (int t = tid; t < p.kernel_h * p.kernel_w; t += 256) {
        int ky = t / p.kernel_w, kx = t - ky * p.kernel_w;
        sk[t] = kernel[(p.kernel_h - 1 - ky) * p.kernel_w + (p.kernel_w - 1 - kx)];
    }
    const int64_t groups = (p.major + g.planes_per_block - 1) / g.planes_per_block;
    for (int64_t grp = blockIdx.x; grp < groups; grp += gridDim.x) {
        const int64_t plane0 = grp * g.planes_per_block;
        const int np = (int)(p.major - plane0 < g.planes_per_block ? p.major - plane0 : g.planes_per_block);
        __syncthreads();                                   // previous group's reads are done
        for (int i = tid; i < np * plane_pad; i += 256) sin[i] = 0.0f;
        __syncthreads();
        const float* src = in + plane0 * plane_in;
        for (int i = tid; i < np * plane_in; i += 256) {
            const int pl = i / plane_in, r = i - pl * plane_in;
            const int iy = r / p.in_w, ix = r - iy * p.in_w;
            const int sy = iy - g.py0, sx = ix - g.px0;
            if (sy >= 0 && sy < g.ph && sx >= 0 && sx < g.pw) sin[pl * plane_pad + sy * g.pw + sx] = src[i];
        }
        __syncthreads();
        float* dst = out + plane0 * plane_out;
        for (int o = tid; o < np * plane_out; o += 256) {
            const int pl = o / plane_out, r = o - pl * plane_out;
            const int oy = r / p.out_w, ox = r - oy * p.out_w;
            const int mid_x = ox * p.down_x + p.up_x - 1 - p.pad_x0, mid_y = oy * p.down_y + p.up_y - 1 - p.pad_y0;
            const int in_x0 = floor_div(mid_x, p.up_x), in_y0 = floor_div(mid_y, p.up_y);
            const int kx0 = (in_x0 + 1) * p.up_x - mid_x - 1, ky0 = (in_y0 + 1) * p.up_y - mid_y - 1;
            const float* row = sin + pl * plane_pad + (in_y0 - g.py0) * g.pw + (in_x0 - g.px0);
            float v = 0.0f;
            for (int ky = ky0; ky < p.kernel_h; ky += p.up_y, row += g.pw) {
                const float* a = row;
                const float* kr = sk + ky * p.kernel_w;
                for (int kx = kx0; kx < p.kernel_w; kx += p.up_x, ++a) v = __fmaf_rn(*a, kr[kx], v);
            }
            dst[o] = v;
        }
    }
}

// The same plane-group scheme for kernels of at most 4 x 4 taps (every mode the reference builds): UP / DOWN are template
// parameters, the flipped taps sit in registers padded with zeros to 4 x 4 (a zero tap leaves the FMA chain's value
// unchanged, so the bits stay the reference's), the tap loops have constant trip counts and no bounds checks, and a
// thread produces four neighbouring outputs of a row from one set of row / plane indices: ~60 instructions per output
// instead of ~250 (measured at 32^2, B*C = 16 k planes: 0.45 -> 1.1 TB/s, level with the tiled kernel there, which keeps 32^2).
template <int UP, int DOWN>
__global__ void __launch_bounds__(256) upfirdn2d_small_plane4_kernel(float* __restrict__ out, const float* __restrict__ in,
                                                                    const float* __restrict__ kernel, UpfirdnParams p,
                                                                    SmallPlaneGeom g) {
    extern __shared__ float sp_smem[];
    float* sin = sp_smem;                                  // [PB][ph][pw]
    constexpr int T = (4 + UP - 1) / UP;                   // taps per axis and output
    const int tid = threadIdx.x;
    const int plane_in = p.in_h * p.in_w, plane_pad = g.ph * g.pw, plane_out = p.out_h * p.out_w;
    float kf[4][4];                                        // flipped, zero-padded
#pragma unroll
    for (int ky = 0; ky < 4; ++ky)
#pragma unroll
        for (int kx = 0; kx < 4; ++kx)
            kf[ky][kx] = (ky < p.kernel_h && kx < p.kernel_w) ? kernel[(p.kernel_h - 1 - ky) * p.kernel_w + (p.kernel_w - 1 - kx)] : 0.0f;
    const int chunks = (p.out_w + 3) >> 2, items_per_plane = p.out_h * chunks;
    const int64_t groups = (p.major + g.planes_per_block - 1) / g.planes_per_block;
    for (int64_t grp = blockIdx.x; grp < groups; grp += gridDim.x) {
        const int64_t plane0 = grp * g.planes_per_block;
        const int np = (int)(p.major - plane0 < g.planes_per_block ? p.major - plane0 : g.planes_per_block);
        __syncthreads();
        for (int i = tid; i < np * plane_pad; i += 256) sin[i] = 0.0f;
        __syncthreads();
        const float* src = in + plane0 * plane_in;
        for (int i = tid; i < np * plane_in; i += 256) {
            const int pl = i / plane_in, r = i - pl * plane_in;
            const int iy = r / p.in_w, ix = r - iy * p.in_w;
            const int sy = iy - g.py0, sx = ix - g.px0;
            if (sy >= 0 && sy < g.ph && sx >= 0 && sx < g.pw) sin[pl * plane_pad + sy * g.pw + sx] = src[i];
        }
        __syncthreads();
        float* dst = out + plane0 * plane_out;
        for (int it = tid; it < np * items_per_plane; it += 256) {
            const int pl = it / items_per_plane, r = it - pl * items_per_plane;
            const int oy = r / chunks, ox0 = (r - oy * chunks) * 4;
            const int mid_y = oy * DOWN + UP - 1 - p.pad_y0, in_y0 = floor_div(mid_y, UP), ky0 = (in_y0 + 1) * UP - mid_y - 1;
            const float* plane = sin + pl * plane_pad + (in_y0 - g.py0) * g.pw - g.px0;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int mid_x = (ox0 + j) * DOWN + UP - 1 - p.pad_x0, in_x0 = floor_div(mid_x, UP), kx0 = (in_x0 + 1) * UP - mid_x - 1;
                // outputs past the row end read the (zero-padded) window of the last real output and are not stored
                const float* a = plane + (ox0 + j < p.out_w ? in_x0 : g.px0);
                float acc = 0.0f;
#pragma unroll
                for (int ty = 0; ty < T; ++ty)
#pragma unroll
                    for (int tx = 0; tx < T; ++tx) {
                        float w = 0.0f;
#pragma unroll
                        for (int sy = 0; sy < UP; ++sy)
#pragma unroll
                            for (int sx = 0; sx < UP; ++sx)
                                if (ty * UP + sy < 4 && tx * UP + sx < 4 && ky0 == sy && kx0 == sx) w = kf[ty * UP + sy][tx * UP + sx];
                        if (ty * UP + ky0 < p.kernel_h && tx * UP + kx0 < p.kernel_w)     // the reference's loops stop at the real taps
                            acc = __fmaf_rn(a[ty * g.pw + tx], w, acc);
                    }
                v[j] = acc;
            }
            float* o = dst + pl * plane_out + oy * p.out_w + ox0;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (ox0 + j < p.out_w) o[j] = v[j];
        }
    }
}

// false when the padded planes do not fit (the caller falls through to the other paths)
static bool launch_small_plane(float* out, const float* in, const float* kernel, const UpfirdnParams& p, cudaStream_t stream) {
    auto first = [](int o, int down, int up, int pad0) { return floor_div(o * down + up - 1 - pad0, up); };
    const bool k4 = p.kernel_h <= 4 && p.kernel_w <= 4 && p.up_x == p.up_y && p.down_x == p.down_y &&
                    ((p.up_x == 1 && p.down_x == 1) || (p.up_x == 2 && p.down_x == 1) || (p.up_x == 1 && p.down_x == 2));
    SmallPlaneGeom g;
    g.px0 = first(0, p.down_x, p.up_x, p.pad_x0);
    g.py0 = first(0, p.down_y, p.up_y, p.pad_y0);
    g.pw = first(p.out_w - 1, p.down_x, p.up_x, p.pad_x0) + ceil_div(k4 ? 4 : p.kernel_w, p.up_x) - g.px0;
    g.ph = first(p.out_h - 1, p.down_y, p.up_y, p.pad_y0) + ceil_div(k4 ? 4 : p.kernel_h, p.up_y) - g.py0;
    const int64_t plane_bytes = (int64_t)g.pw * g.ph * 4, budget = 40 * 1024;
    if (plane_bytes > budget) return false;
    int64_t ppb = budget / plane_bytes;
    const int64_t want_blocks = (int64_t)kNumSMs * 4;       // keep every SM busy before growing the groups
    if (ppb * want_blocks > p.major) ppb = p.major / want_blocks;
    if (ppb < 1) ppb = 1;
    g.planes_per_block = (int)ppb;
    const int64_t groups = (p.major + ppb - 1) / ppb;
    const int64_t cap = (int64_t)kNumSMs * 5;
    const int grid = (int)(groups < cap ? groups : cap);
    if (k4) {
        const int smem = (int)(ppb * plane_bytes);
        if (p.up_x == 1 && p.down_x == 1) upfirdn2d_small_plane4_kernel<1, 1><<<grid, 256, smem, stream>>>(out, in, kernel, p, g);
        else if (p.up_x == 2) upfirdn2d_small_plane4_kernel<2, 1><<<grid, 256, smem, stream>>>(out, in, kernel, p, g);
        else upfirdn2d_small_plane4_kernel<1, 2><<<grid, 256, smem, stream>>>(out, in, kernel, p, g);
        return true;
    }
    const int smem = (int)(p.kernel_h * p.kernel_w * 4 + ppb * plane_bytes);
    upfirdn2d_small_plane_kernel<<<grid, 256, smem, stream>>>(out, in, kernel, p, g);
    return true;
}

// ------------------------------------------------------------------------------------------------ tiled
// One block = one TH x 64 output tile of TWO planes (minor == 1), interleaved in shared memory as float2 so that every
// shared-memory access is an LDS.64 and every multiply-add is a packed FFMA2 (two planes per instruction).  K = padded
// (template) tap count as in the reference's mode table; taps beyond the real kernel are zero, so each plane's FMA chain
// has the reference's order (y outer, x inner, from 0) and results stay bit-identical to the reference kernel.
// A thread owns 4 consecutive output columns x 2 rows; for UP == 1 the (3*DOWN + K)-wide input span of each needed row
// is loaded once into registers and shared by the 4 columns and both rows.  All index arithmetic is per thread / per
// row, none per tap: the first version of this kernel was instruction-bound at 31 % of HBM peak.
using u64 = unsigned long long;
__device__ __forceinline__ float2 up_ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(reinterpret_cast<u64&>(d)) : "l"(reinterpret_cast<u64&>(a)), "l"(reinterpret_cast<u64&>(b)), "l"(reinterpret_cast<u64&>(c)));
    return d;
}

template <int UP, int DOWN, int K, int TH>
__global__ void __launch_bounds__(256) upfirdn2d_pair_kernel(float* __restrict__ out, const float* __restrict__ in,
                                                             const float* __restrict__ kernel, UpfirdnParams p,
                                                             int tiles_x, int tiles_y) {
    constexpr int TW = 64, RPT = TH / 16;                    // 16 column groups x 16 row groups, RPT rows per thread
    constexpr int TIN_H = ((TH - 1) * DOWN + K - 1) / UP + 1;
    constexpr int TIN_W = ((TW - 1) * DOWN + K - 1) / UP + 1;
    constexpr int TAPS = K / UP;
    // Lane t of a half-warp starts its 4 outputs at input column 4*DOWN/UP * t: a 32 B (or 16 / 64 B) lane stride would
    // serialise the LDS.64 on 4 (2 / 8) banks.  One padding slot every PADG columns makes the stride odd in 8-byte units.
    constexpr int PADG = 4 * DOWN / UP;
    constexpr int TIN_WP = TIN_W + TIN_W / PADG + 1;
    __shared__ float sk[K][K];
    constexpr int NBUF = (DOWN == 2) ? 1 : 2;               // the decimating tile (35 KB) is single-buffered
    extern __shared__ __align__(16) float2 sx_dyn[];         // [NBUF][TIN_H][TIN_WP], double-buffered: tile i+1 is in flight
    float2 (*sxbuf)[TIN_H][TIN_WP] = reinterpret_cast<float2 (*)[TIN_H][TIN_WP]>(sx_dyn);   // while tile i is computed
#define SXC(c) ((c) + (c) / PADG)

    for (int t = threadIdx.x; t < K * K; t += 256) {
        int ky = t / K, kx = t - ky * K;
        float v = 0.0f;
        if (kx < p.kernel_w && ky < p.kernel_h) v = kernel[(p.kernel_h - 1 - ky) * p.kernel_w + (p.kernel_w - 1 - kx)];
        sk[ky][kx] = v;
    }
    const int64_t pairs = (p.major + 1) / 2;
    const int64_t tiles_per_pair = (int64_t)tiles_x * tiles_y;
    const int64_t total_tiles = tiles_per_pair * pairs;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t in_plane = (int64_t)p.in_h * p.in_w, out_plane = (int64_t)p.out_h * p.out_w;

    // stage one tile: warp w takes rows w, w+8, ...; lanes take consecutive columns (coalesced 4-byte cp.async into the
    // interleaved float2 slots, fire-and-forget; out-of-range elements are zero-filled with plain stores)
    auto stage = [&](int64_t tile, int buf) {
        const int64_t pair = tile / tiles_per_pair;
        const int trem = (int)(tile - pair * tiles_per_pair);
        const int tile_out_y = (trem / tiles_x) * TH, tile_out_x = (trem % tiles_x) * TW;
        const int tile_in_x = floor_div(tile_out_x * DOWN + UP - 1 - p.pad_x0, UP);
        const int tile_in_y = floor_div(tile_out_y * DOWN + UP - 1 - p.pad_y0, UP);
        const int64_t plane0 = pair * 2;
        const float* src0 = in + plane0 * in_plane;
        const float* src1 = (plane0 + 1 < p.major) ? src0 + in_plane : src0;
        const uint32_t sx_u32 = (uint32_t)__cvta_generic_to_shared(&sxbuf[buf][0][0]);
        for (int ry = warp; ry < TIN_H; ry += 8) {
            const int iy = ry + tile_in_y;
            const bool rowok = iy >= 0 && iy < p.in_h;
            const float* r0 = src0 + (int64_t)iy * p.in_w + tile_in_x;
            const float* r1 = src1 + (int64_t)iy * p.in_w + tile_in_x;
#pragma unroll
            for (int rx = lane; rx < TIN_W; rx += 32) {
                const int ix = rx + tile_in_x;
                const uint32_t dst = sx_u32 + (uint32_t)(ry * TIN_WP + SXC(rx)) * 8u;
                if (rowok && ix >= 0 && ix < p.in_w) {
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(r0 + rx) : "memory");
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 4u), "l"(r1 + rx) : "memory");
                } else {
                    asm volatile("st.shared.v2.f32 [%0], {%1, %1};" ::"r"(dst), "f"(0.0f) : "memory");
                }
            }
        }
    };

    int buf = 0;
    if (NBUF == 2 && (int64_t)blockIdx.x < total_tiles) stage(blockIdx.x, 0);
    for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, buf ^= (NBUF - 1)) {
        const int64_t pair = tile / tiles_per_pair;
        const int trem = (int)(tile - pair * tiles_per_pair);
        const int tile_out_y = (trem / tiles_x) * TH, tile_out_x = (trem % tiles_x) * TW;
        const int tile_mid_x = tile_out_x * DOWN + UP - 1 - p.pad_x0;
        const int tile_mid_y = tile_out_y * DOWN + UP - 1 - p.pad_y0;
        const int tile_in_x = floor_div(tile_mid_x, UP), tile_in_y = floor_div(tile_mid_y, UP);
        const int64_t plane0 = pair * 2;
        const bool has1 = plane0 + 1 < p.major;
        if (NBUF == 1) {
            __syncthreads();  // everybody is done computing the previous tile
            stage(tile, 0);
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();      // this tile's data has landed; everybody is done with the other buffer
        if (NBUF == 2 && tile + gridDim.x < total_tiles) stage(tile + gridDim.x, buf ^ 1);
        float2 (*sx)[TIN_WP] = sxbuf[buf];

        float2 acc[RPT][4];
        if (UP == 1) {
            // register window: rows [rel_y0, rel_y0 + (RPT-1)*DOWN + K), columns [rel_x0, rel_x0 + 3*DOWN + K)
            constexpr int SPAN = 3 * DOWN + K, ROWS = (RPT - 1) * DOWN + K;
            const int rel_x0 = tx * 4 * DOWN, rel_y0 = ty * RPT * DOWN;    // UP == 1: mid == in, kernel phase 0
            float wreg[K][K];
#pragma unroll
            for (int ky = 0; ky < K; ++ky)
#pragma unroll
                for (int kx = 0; kx < K; ++kx) wreg[ky][kx] = sk[ky][kx];
#pragma unroll
            for (int r = 0; r < RPT; ++r)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[r][j] = make_float2(0.0f, 0.0f);
#pragma unroll
            for (int wy = 0; wy < ROWS; ++wy) {
                float2 row[SPAN];
#pragma unroll
                for (int c = 0; c < SPAN; ++c) row[c] = sx[rel_y0 + wy][SXC(rel_x0) + SXC(c)];   // rel_x0 % PADG == 0
#pragma unroll
                for (int r = 0; r < RPT; ++r) {
                    const int ky = wy - r * DOWN;              // compile-time after unrolling
                    if (ky >= 0 && ky < K) {
#pragma unroll
                        for (int kx = 0; kx < K; ++kx) {
                            const float w = wreg[ky][kx];
#pragma unroll
                            for (int j = 0; j < 4; ++j) acc[r][j] = up_ffma2(row[j * DOWN + kx], make_float2(w, w), acc[r][j]);
                        }
                    }
                }
            }
        } else {
            // UP > 1: TAPS x TAPS taps per output, phase pattern periodic in the output coordinate
            int rel_x[4], kx0[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int mid_x = tile_mid_x + (tx * 4 + j) * DOWN;
                const int in_x = floor_div(mid_x, UP);
                rel_x[j] = in_x - tile_in_x;
                kx0[j] = (in_x + 1) * UP - mid_x - 1;
            }
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                const int mid_y = tile_mid_y + (ty * RPT + r) * DOWN;
                const int in_y = floor_div(mid_y, UP);
                const int rel_y = in_y - tile_in_y, ky0 = (in_y + 1) * UP - mid_y - 1;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float2 v = make_float2(0.0f, 0.0f);
#pragma unroll
                    for (int y = 0; y < TAPS; ++y)
#pragma unroll
                        for (int x = 0; x < TAPS; ++x) {
                            const float w = sk[ky0 + y * UP][kx0[j] + x * UP];
                            v = up_ffma2(sx[rel_y + y][SXC(rel_x[j] + x)], make_float2(w, w), v);
                        }
                    acc[r][j] = v;
                }
            }
        }
        // NOTE on ordering: for UP == 1 the accumulation runs over wy (= ky for row r) outer and kx inner, i.e. the
        // reference's y-outer / x-inner chain per output.
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int oy = tile_out_y + ty * RPT + r, ox = tile_out_x + tx * 4;
            if (oy >= p.out_h) continue;
            float* d0 = out + plane0 * out_plane + (int64_t)oy * p.out_w + ox;
            float* d1 = d0 + out_plane;
            if (ox + 3 < p.out_w && ((((uintptr_t)d0) & 15) == 0) && ((((uintptr_t)d1) & 15) == 0)) {
                st_stream_f4((float4*)d0, make_float4(acc[r][0].x, acc[r][1].x, acc[r][2].x, acc[r][3].x));
                if (has1) st_stream_f4((float4*)d1, make_float4(acc[r][0].y, acc[r][1].y, acc[r][2].y, acc[r][3].y));
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (ox + j < p.out_w) {
                        d0[j] = acc[r][j].x;
                        if (has1) d1[j] = acc[r][j].y;
                    }
            }
        }
    }
}

#undef SXC

template <int UP, int DOWN, int K>
static void launch_tiled(float* out, const float* in, const float* kernel, const UpfirdnParams& p,
                         cudaStream_t stream) {
    // tall tiles for the up == down == 1 FIR (4 rows per thread: the per-tile overheads and the window loads are
    // amortised over 32 outputs per thread); 32 rows for the interpolating modes, 16 for the decimating one
    constexpr int TH = (DOWN == 2) ? 16 : (UP == 1 ? 64 : 32), TW = 64;
    constexpr int TIN_H = ((TH - 1) * DOWN + K - 1) / UP + 1, TIN_W = ((TW - 1) * DOWN + K - 1) / UP + 1;
    constexpr int PADG = 4 * DOWN / UP, TIN_WP = TIN_W + TIN_W / PADG + 1, NBUF = (DOWN == 2) ? 1 : 2;
    constexpr int SMEM = NBUF * TIN_H * TIN_WP * 8;
    auto kern = upfirdn2d_pair_kernel<UP, DOWN, K, TH>;
    static int blocks_per_sm = 0;
    if (!blocks_per_sm) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, 256, SMEM) != cudaSuccess || blocks_per_sm < 1) blocks_per_sm = 1;
    }
    int tiles_x = ceil_div(p.out_w, TW), tiles_y = ceil_div(p.out_h, TH);
    int64_t total = (int64_t)tiles_x * tiles_y * ((p.major + 1) / 2);
    int64_t cap = (int64_t)kNumSMs * blocks_per_sm;      // exactly one resident wave of persistent blocks
    int grid = (int)(total < cap ? total : cap);
    kern<<<grid, 256, SMEM, stream>>>(out, in, kernel, p, tiles_x, tiles_y);
}

}  // namespace sis

using namespace sis;

extern "C" int sis_upfirdn2d_out_size(int in_size, int up, int down, int pad0, int pad1, int ksize) {
    // upfirdn2d_kernel.cu:168-169
    return (in_size * up + pad0 + pad1 - ksize + down) / down;
}

extern "C" int sis_upfirdn2d(void* d_out, const void* d_x, const void* d_kernel, int dtype, int64_t major, int in_h,
                             int in_w, int minor, int kernel_h, int kernel_w, int up_x, int up_y, int down_x,
                             int down_y, int pad_x0, int pad_x1, int pad_y0, int pad_y1, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIS_REQUIRE(up_x >= 1 && up_y >= 1 && down_x >= 1 && down_y >= 1, "upfirdn2d: up/down must be >= 1");
    SIS_REQUIRE(kernel_h >= 1 && kernel_w >= 1 && kernel_h <= kMaxTaps && kernel_w <= kMaxTaps,
                "upfirdn2d: kernel must be between 1x1 and %dx%d (got %dx%d)", kMaxTaps, kMaxTaps, kernel_h, kernel_w);
    SIS_REQUIRE(major >= 0 && in_h >= 0 && in_w >= 0 && minor >= 0, "upfirdn2d: negative size");
    UpfirdnParams p;
    p.up_x = up_x; p.up_y = up_y; p.down_x = down_x; p.down_y = down_y; p.pad_x0 = pad_x0; p.pad_y0 = pad_y0;
    p.in_h = in_h; p.in_w = in_w; p.minor = minor; p.kernel_h = kernel_h; p.kernel_w = kernel_w; p.major = major;
    p.out_h = sis_upfirdn2d_out_size(in_h, up_y, down_y, pad_y0, pad_y1, kernel_h);
    p.out_w = sis_upfirdn2d_out_size(in_w, up_x, down_x, pad_x0, pad_x1, kernel_w);
    SIS_REQUIRE(p.out_h >= 0 && p.out_w >= 0, "upfirdn2d: negative output size %dx%d", p.out_h, p.out_w);
    const int64_t total = major * p.out_h * p.out_w * minor;
    if (total == 0) return SIS_OK;
    SIS_REQUIRE(d_out && d_kernel, "upfirdn2d: kernel must be a CUDA tensor (null pointer)");
    SIS_REQUIRE(d_x || (int64_t)in_h * in_w * minor * major == 0, "upfirdn2d: input must be a CUDA tensor (null pointer)");

    bool done = false;
    // Small planes (the 4^2 ... 16^2 maps): a 64-wide tile per plane pair would leave most of a block idle; they take the
    // shared-memory plane-group kernel (same y-outer / x-inner FMA chain over the real taps, i.e. the same bits).  At 32^2
    // the tiled kernel is still slightly faster (measured 1.2-1.3 vs 1.0-1.1 TB/s: both instruction-bound there).
    const bool small_plane = (int64_t)p.out_h * p.out_w <= 16 * 16;
    if (small_plane && dtype == SIS_F32 && minor == 1 && in_h > 0 && in_w > 0)
        done = launch_small_plane((float*)d_out, (const float*)d_x, (const float*)d_kernel, p, stream);
    if (!done && dtype == SIS_F32 && minor == 1 && up_x == up_y && down_x == down_y && in_h > 0 && in_w > 0 && !small_plane) {
        float* o = (float*)d_out; const float* x = (const float*)d_x; const float* k = (const float*)d_kernel;
        const int kmax = kernel_h > kernel_w ? kernel_h : kernel_w;
        done = true;
        // the reference's mode table (upfirdn2d_kernel.cu:177-211); later matches override earlier ones there.
        if (up_x == 1 && down_x == 1 && kmax <= 3) launch_tiled<1, 1, 3>(o, x, k, p, stream);
        else if (up_x == 1 && down_x == 1 && kmax <= 4) launch_tiled<1, 1, 4>(o, x, k, p, stream);
        else if (up_x == 2 && down_x == 1 && kmax <= 2) launch_tiled<2, 1, 2>(o, x, k, p, stream);
        else if (up_x == 2 && down_x == 1 && kmax <= 4) launch_tiled<2, 1, 4>(o, x, k, p, stream);
        else if (up_x == 1 && down_x == 2 && kmax <= 4) launch_tiled<1, 2, 4>(o, x, k, p, stream);  // modes 5 and 6
        else done = false;
    }
    if (!done) {
        int64_t blocks = ceil_div64(total, 256);
        int64_t cap = (int64_t)kNumSMs * 8;
        int grid = (int)(blocks < cap ? blocks : cap);
        if (dtype == SIS_F32)
            upfirdn2d_generic_kernel<float><<<grid, 256, 0, stream>>>((float*)d_out, (const float*)d_x, (const float*)d_kernel, p);
        else if (dtype == SIS_F64)
            upfirdn2d_generic_kernel<double><<<grid, 256, 0, stream>>>((double*)d_out, (const double*)d_x, (const double*)d_kernel, p);
        else if (dtype == SIS_F16)
            upfirdn2d_generic_kernel<__half><<<grid, 256, 0, stream>>>((__half*)d_out, (const __half*)d_x, (const __half*)d_kernel, p);
        else {
            set_error("upfirdn2d: unsupported dtype %d", dtype);
            return SIS_ERR_UNSUPPORTED;
        }
    }
    SIS_CHECK_LAUNCH();
    return SIS_OK;
}
