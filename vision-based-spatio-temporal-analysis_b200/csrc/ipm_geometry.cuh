// ipm_geometry.cuh -- per-cell projection arithmetic shared by every kernel of the library.
//
// Follows, op for op and rounding for rounding, the fp32 chain the reference executes on the
// CPU; each step cites the reference line it replaces.  All arithmetic uses
// the explicit _rn intrinsics so nvcc cannot contract or re-associate anything.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bevipm {

// What one BEV cell needs from one view: the NW texel index, the four bilinear weights and
// which taps exist.  32 bytes so a warp-uniform read is two LDS.128 broadcasts.
struct __align__(16) CellTap {
    // first 16 bytes + next 8: all the fused kernel reads per step (LDS.128 + LDS.64)
    int off16;             // NW texel offset from the view's base in 16-byte units (NHWC fast path only)
    int flags;             // bit0..3: NW, NE, SW, SE inside the map; bit4: non-finite coordinate
    float nw, ne, sw, se;  // ATen grid-sampler weights
    int x0, y0;            // NW texel (clamped to [-2, size] so the int conversion is always defined)
};
static_assert(sizeof(CellTap) == 32, "CellTap must stay 32 bytes");

constexpr int kTapMask = 15;
constexpr int kNonFinite = 16;

// geometry.py:60-63   G = [r1 r2 t];  H = K @ G   (aten::mm, K=3: k-ordered fma chain)
__device__ __forceinline__ void homography(const float* __restrict__ K, const float* __restrict__ Rt34,
                                           float* H) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int gj = (j == 2) ? 3 : j;
            float acc = __fmul_rn(K[i * 3 + 0], Rt34[0 * 4 + gj]);
            acc = __fmaf_rn(K[i * 3 + 1], Rt34[1 * 4 + gj], acc);
            acc = __fmaf_rn(K[i * 3 + 2], Rt34[2 * 4 + gj], acc);
            H[i * 3 + j] = acc;
        }
    }
}

// geometry.py:144-158 + grid_sampler's align_corners=False un-normalisation:
// world ground point (x, y) -> feature-pixel sample position (ix, iy).
// kx, ky = 0: the reference's grid_sample branch (the chain below).  kx = Wf/(Wf-1), ky = Hf/(Hf-1): the sample
// position of the reference's kornia branch (geometry.py:124-141 through kornia's warp_perspective: pixel grid
// normalised with (size-1), sampled with align_corners=False) -- a from-spec compatibility mode, parity unpinned.
__device__ __forceinline__ void cell_coord(const float* H, float x, float y, float sw, float sh,
                                           float Wf, float Hf, float& ix, float& iy, float kx = 0.0f, float ky = 0.0f) {
    // :145  uvw = H @ [x; y; 1]
    const float r0 = __fmaf_rn(H[2], 1.0f, __fmaf_rn(H[1], y, __fmul_rn(H[0], x)));
    const float r1 = __fmaf_rn(H[5], 1.0f, __fmaf_rn(H[4], y, __fmul_rn(H[3], x)));
    const float r2 = __fmaf_rn(H[8], 1.0f, __fmaf_rn(H[7], y, __fmul_rn(H[6], x)));
    // :146-147  w_safe = where(|w| < 1e-6, 1, w)      (no w < 0 cull, on purpose)
    const float w = (fabsf(r2) < 1e-6f) ? 1.0f : r2;
    // :148-149  IEEE division
    const float u = __fdiv_rn(r0, w);
    const float v = __fdiv_rn(r1, w);
    // :151-155  image px -> feature px
    const float fx = __fmul_rn(u, sw);
    const float fy = __fmul_rn(v, sh);
    if (kx != 0.0f) {  // kornia: source pixel p -> 2p/(size-1)-1 -> grid_sample(align_corners=False): p * size/(size-1) - 0.5
        ix = __fmaf_rn(fx, kx, -0.5f);
        iy = __fmaf_rn(fy, ky, -0.5f);
        return;
    }
    // :156-158  (p + 0.5) / size * 2 - 1, four separately rounded ops, true division
    const float nx = __fsub_rn(__fmul_rn(__fdiv_rn(__fadd_rn(fx, 0.5f), Wf), 2.0f), 1.0f);
    const float ny = __fsub_rn(__fmul_rn(__fdiv_rn(__fadd_rn(fy, 0.5f), Hf), 2.0f), 1.0f);
    // ATen GridSampler.h:27-36  ((n + 1) * size - 1) / 2  ==  fma(n + 1, size / 2, -0.5) on the CPU
    ix = __fmaf_rn(__fadd_rn(nx, 1.0f), __fmul_rn(Wf, 0.5f), -0.5f);
    iy = __fmaf_rn(__fadd_rn(ny, 1.0f), __fmul_rn(Hf, 0.5f), -0.5f);
}

// Bilinear set-up of ATen's grid sampler: floor, distances, weight products, per-tap bounds.
__device__ __forceinline__ CellTap make_tap(float ix, float iy, int Wf, int Hf, int fsy16 = 0, int fsx16 = 0) {
    CellTap t;
    const float x0 = floorf(ix), y0 = floorf(iy);
    const float wx = __fsub_rn(ix, x0), ex = __fsub_rn(1.0f, wx);
    const float wy = __fsub_rn(iy, y0), ey = __fsub_rn(1.0f, wy);
    t.nw = __fmul_rn(ey, ex);
    t.ne = __fmul_rn(ey, wx);
    t.sw = __fmul_rn(wy, ex);
    t.se = __fmul_rn(wy, wx);
    const float Wm = (float)Wf, Hm = (float)Hf;
    // |c| <= FLT_MAX is false for NaN and +-inf: those cells turn NaN in the reference (0 * NaN)
    const bool finite = (fabsf(ix) <= 3.402823466e+38f) && (fabsf(iy) <= 3.402823466e+38f);
    const bool xw = (x0 > -1.0f) && (x0 < Wm);
    const bool xe = (x0 + 1.0f > -1.0f) && (x0 + 1.0f < Wm);
    const bool yn = (y0 > -1.0f) && (y0 < Hm);
    const bool ys = (y0 + 1.0f > -1.0f) && (y0 + 1.0f < Hm);
    int f = (xw && yn ? 1 : 0) | (xe && yn ? 2 : 0) | (xw && ys ? 4 : 0) | (xe && ys ? 8 : 0);
    t.flags = finite ? f : kNonFinite;
    t.x0 = (int)fminf(fmaxf(x0, -2.0f), Wm);
    t.y0 = (int)fminf(fmaxf(y0, -2.0f), Hm);
    if (!finite) { t.x0 = -2; t.y0 = -2; }
    t.off16 = t.y0 * fsy16 + t.x0 * fsx16;
    return t;
}

}  // namespace bevipm
