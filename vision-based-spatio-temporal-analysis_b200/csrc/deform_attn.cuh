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
// Mapping: one lane owns 16 bytes of one head's D channels (LPH = D*elem/16 lanes per head), a warp owns
// 32/LPH (query, head) pairs.  The L*P samples of a pair are walked in rounds of LPH: in a round every lane of
// the head group projects ONE sample (un-normalisation, floor, the four bilinear weights times the attention
// weight, tap index clamped into the map with weight 0 for taps outside it), then the group takes the LPH
// samples one by one from its lanes by shuffle, reads the four taps as 16-byte vectors (the LPH lanes of a head
// read one contiguous D*elem-byte run per tap) and accumulates  acc += tap * (w_tap * A)  in fp32 registers.
// No branch depends on the data, so the 4 * LPH loads of a round are in flight together; the next round's
// locations and weights are fetched one round ahead.  Memory-bound gather: no tensor cores.
// (ncu on the first, sample-by-sample form: 62 % of the stall samples on two dependent global loads per sample,
//  149 instructions per sample and warp; profiles/r01_notes.md.)
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
    const long long npairs = (long long)p.B * p.Q * p.M;
    long long pair = warp * PAIRS + lane / LPH;  // flat (b, q, m)
    const bool live = pair < npairs;             // dead groups of the last warp walk pair 0 and store nothing
    if (!live) pair = 0;
    const int sub = lane % LPH;                  // which 16-byte slice of the head's channels / which sample of a round
    const int m = (int)(pair % p.M);
    const int b = (int)(pair / p.M / p.Q);
    const int LP = p.L * p.P;
    const float2* loc = reinterpret_cast<const float2*>(p.loc) + pair * LP;
    const float* aw = p.attn + pair * LP;
    // this lane's slice of head m at spatial position 0 of frame b, in 16-byte units; one position further = M * LPH units
    const uint4* vb = reinterpret_cast<const uint4*>(reinterpret_cast<const TIn*>(p.value) + ((long long)b * p.S * p.M + m) * p.D) + sub;
    const int pos_stride = p.M * LPH;

    float2 acc[PR];
#pragma unroll
    for (int e = 0; e < PR; ++e) acc[e] = make_float2(0.0f, 0.0f);

    // the sample this lane projects in round r is s = r * LPH + sub
    float2 xy = make_float2(0.0f, 0.0f);
    float a = 0.0f;
    if (sub < LP) { xy = __ldg(loc + sub); a = __ldg(aw + sub); }
    for (int s0 = 0; s0 < LP; s0 += LPH) {
        const int s = s0 + sub;
        const bool mine = s < LP;
        const int l = mine ? s / p.P : 0;
        const int H = __ldg(p.shapes + 2 * l), W = __ldg(p.shapes + 2 * l + 1);
        const long long st = __ldg(p.start + l);
        const float2 xy_c = xy;
        const float a_c = mine ? a : 0.0f;
        if (s + LPH < LP) { xy = __ldg(loc + s + LPH); a = __ldg(aw + s + LPH); }  // next round, one round ahead
        // grid_sample un-normalisation with align_corners=False: ((2*loc-1 + 1) * size - 1) / 2
        const float x = __fmaf_rn(xy_c.x, (float)W, -0.5f);
        const float y = __fmaf_rn(xy_c.y, (float)H, -0.5f);
        const bool inside = y > -1.0f && x > -1.0f && y < (float)H && x < (float)W;  // false for NaN too
        const float xs_ = inside ? x : 0.0f, ys_ = inside ? y : 0.0f;
        const float x0f = floorf(xs_), y0f = floorf(ys_);
        const float lx = xs_ - x0f, ly = ys_ - y0f, hx = 1.0f - lx, hy = 1.0f - ly;
        const int x0 = (int)x0f, y0 = (int)y0f;
        const bool xw = x0 >= 0, xe = x0 + 1 <= W - 1, yn = y0 >= 0, ysb = y0 + 1 <= H - 1;
        // samples outside every map (and non-finite locations: every comparison above is false) contribute nothing
        const float aa = inside ? a_c : 0.0f;
        const float w00 = (yn && xw) ? hy * hx * aa : 0.0f, w01 = (yn && xe) ? hy * lx * aa : 0.0f;
        const float w10 = (ysb && xw) ? ly * hx * aa : 0.0f, w11 = (ysb && xe) ? ly * lx * aa : 0.0f;
        // clamped taps: NW at (yc, xc), the others dx / dy positions further (0 when clamped onto NW's column / row)
        const int xc = max(x0, 0), yc = max(y0, 0);
        const int dx = (min(x0 + 1, W - 1) - xc) * pos_stride;
        const int dy = (min(y0 + 1, H - 1) - yc) * W * pos_stride;
        const int base = (int)((st + (long long)yc * W + xc) * pos_stride);   // < 2^31 16-byte units: checked by the launcher
#pragma unroll
        for (int k = 0; k < LPH; ++k) {
            // take sample s0 + k from lane k of this head group (rounds past the end carry weight 0)
            const int kb = __shfl_sync(0xffffffffu, base, k, LPH);
            const int kdx = __shfl_sync(0xffffffffu, dx, k, LPH), kdy = __shfl_sync(0xffffffffu, dy, k, LPH);
            const float k00 = __shfl_sync(0xffffffffu, w00, k, LPH), k01 = __shfl_sync(0xffffffffu, w01, k, LPH);
            const float k10 = __shfl_sync(0xffffffffu, w10, k, LPH), k11 = __shfl_sync(0xffffffffu, w11, k, LPH);
            const uint4* t = vb + kb;
            float2 f[4][PR];
            VT::unpack(ldg16(t), f[0]);
            VT::unpack(ldg16(t + kdx), f[1]);
            VT::unpack(ldg16(t + kdy), f[2]);
            VT::unpack(ldg16(t + kdy + kdx), f[3]);
#pragma unroll
            for (int e = 0; e < PR; ++e) {
                acc[e] = __ffma2_rn(f[0][e], make_float2(k00, k00), acc[e]);
                acc[e] = __ffma2_rn(f[1][e], make_float2(k01, k01), acc[e]);
                acc[e] = __ffma2_rn(f[2][e], make_float2(k10, k10), acc[e]);
                acc[e] = __ffma2_rn(f[3][e], make_float2(k11, k11), acc[e]);
            }
        }
    }
    if (live) {
        TOut* op = reinterpret_cast<TOut*>(p.out) + pair * p.D + sub * VE;
        store_pairs<TOut, PR>(op, acc);
    }
}

}  // namespace bevipm
