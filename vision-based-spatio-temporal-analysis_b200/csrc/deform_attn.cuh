// deform_attn.cuh -- multi-view / multi-head deformable-attention sampling (Phase-2 BEV fusion).
//
// The reference has only a placeholder for this (models/fusion/fusion.py:25-36, roadmap README.md:57-63);
// the semantics implemented here are the published ones of Deformable-DETR's MSDeformAttn, which
// MVDeTr uses with one "level" per camera view:
//
//   out[b,q,m,:] = sum_l sum_p  A[b,q,m,l,p] * bilinear(value_l[b,:,m,:], loc[b,q,m,l,p])
//
// value  [B, S, M, D]   S = sum_l H_l*W_l, level l starts at level_start[l], row-major (y, x)
// loc    [B, Q, M, L, P, 2]  (x, y) normalised to [0,1]; pixel = loc*size - 0.5  (grid_sample,
//                            align_corners=False), taps outside the map read as zero
// A      [B, Q, M, L, P]     attention weights (already soft-maxed by the caller)
// out    [B, Q, M*D]
//
// Mapping: one lane owns 16 bytes of one head's D channels (LPH = D*elem/16 lanes per head), a warp
// owns 32/LPH (query, head) pairs; each lane walks the L*P samples of its pair, reads its four taps
// as 16-byte vectors (the LPH lanes of a head read one contiguous D*elem-byte run per tap) and
// accumulates in fp32 registers.  Memory-bound gather: no tensor cores.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ipm_fused.cuh"

namespace bevipm {

struct DeformParams {
    const void* value;
    const int* shapes;          // [L,2] (H, W)
    const long long* start;     // [L]
    const float* loc;
    const float* attn;
    void* out;
    int B, Q, M, D, L, P;
    long long S;
};

template <typename TIn, typename TOut, int LPH>
__global__ void __launch_bounds__(256) deform_attn_kernel(const DeformParams p) {
    using VT = VecTraits<TIn>;
    constexpr int VE = VT::VE, PR = VT::P;
    constexpr int PAIRS = 32 / LPH;  // (query, head) pairs per warp
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long pair = warp * PAIRS + lane / LPH;  // flat (b, q, m)
    const long long npairs = (long long)p.B * p.Q * p.M;
    if (pair >= npairs) return;
    const int sub = lane % LPH;                         // which 16-byte slice of the head's channels
    const int m = (int)(pair % p.M);
    const long long bq = pair / p.M;
    const int b = (int)(bq / p.Q);
    const int LP = p.L * p.P;
    const float* loc = p.loc + pair * LP * 2;
    const float* aw = p.attn + pair * LP;
    const TIn* vb = reinterpret_cast<const TIn*>(p.value) + ((long long)b * p.S * p.M + m) * p.D + sub * VE;
    const long long row_stride = (long long)p.M * p.D;  // elements between consecutive spatial positions

    float2 acc[PR];
#pragma unroll
    for (int e = 0; e < PR; ++e) acc[e] = make_float2(0.0f, 0.0f);

    for (int l = 0; l < p.L; ++l) {
        const int H = __ldg(p.shapes + 2 * l), W = __ldg(p.shapes + 2 * l + 1);
        const TIn* lv = vb + __ldg(p.start + l) * row_stride;
        for (int pt = 0; pt < p.P; ++pt) {
            const float2 xy = __ldg(reinterpret_cast<const float2*>(loc) + l * p.P + pt);
            const float a = __ldg(aw + l * p.P + pt);
            // grid_sample un-normalisation with align_corners=False: ((2*loc-1 + 1) * size - 1) / 2
            const float x = __fmaf_rn(xy.x, (float)W, -0.5f);
            const float y = __fmaf_rn(xy.y, (float)H, -0.5f);
            if (!(y > -1.0f && x > -1.0f && y < (float)H && x < (float)W)) continue;  // all four taps outside
            const float x0f = floorf(x), y0f = floorf(y);
            const float lx = x - x0f, ly = y - y0f, hx = 1.0f - lx, hy = 1.0f - ly;
            const int x0 = (int)x0f, y0 = (int)y0f;
            const bool xw = x0 >= 0, xe = x0 + 1 <= W - 1, yn = y0 >= 0, ys = y0 + 1 <= H - 1;
            const TIn* t00 = lv + ((long long)y0 * W + x0) * row_stride;
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
            uint4 raw[4][1];
            raw[0][0] = (yn && xw) ? ldg16(reinterpret_cast<const uint4*>(t00)) : z;
            raw[1][0] = (yn && xe) ? ldg16(reinterpret_cast<const uint4*>(t00 + row_stride)) : z;
            raw[2][0] = (ys && xw) ? ldg16(reinterpret_cast<const uint4*>(t00 + (long long)W * row_stride)) : z;
            raw[3][0] = (ys && xe) ? ldg16(reinterpret_cast<const uint4*>(t00 + (long long)(W + 1) * row_stride)) : z;
            StepHdr h;
            h.off16 = 0; h.flags = kTapMask;
            h.nw = hy * hx; h.ne = hy * lx; h.sw = ly * hx; h.se = ly * lx;
            float2 o[1][PR];
            blend<TIn, 1>(raw, h, o);
            const float2 a2 = make_float2(a, a);
#pragma unroll
            for (int e = 0; e < PR; ++e) acc[e] = __ffma2_rn(o[0][e], a2, acc[e]);
        }
    }
    TOut* op = reinterpret_cast<TOut*>(p.out) + pair * p.D + sub * VE;
    store_pairs<TOut, PR>(op, acc);
}

}  // namespace bevipm
