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

#include "ipm_aux.cuh"
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

namespace bevipm {

// ---- backward ---------------------------------------------------------------------------------------------------------
// d out / d value (scatter, fp32 atomics), d out / d sampling_locations and d out / d attention_weights (Deformable-DETR's
// ms_deform_attn backward, restated from its published formulas).  Same mapping as the forward: LPH lanes own one (query,
// head) pair, a lane owns 16 bytes of the head's channels.  Per sample the four taps are re-read, the lane forms its part of
//     t   = sum_d g_d * (w00 v00 + w01 v01 + w10 v10 + w11 v11)_d          -> grad_attn = t (with the attention weight left out)
//     gx  = sum_d g_d * (hy (v01 - v00) + ly (v11 - v10))_d * W * A         -> grad_loc.x
//     gy  = sum_d g_d * (hx (v10 - v00) + lx (v11 - v01))_d * H * A         -> grad_loc.y
// (taps outside the map count as zero), the LPH lanes add their parts by shuffle, lane 0 of the group stores the three
// numbers, and every lane adds  g_d * w_tap * A  into grad_value at its four taps (red.global.add.v4.f32).
struct DeformBwdParams {
    DeformParams f;            // forward tensors (out unused)
    const void* gout;          // [B,Q,M*D] TG
    float* gvalue;             // [B,S,M,D] f32, pre-zeroed
    float* gloc;               // [B,Q,M,L,P,2] f32
    float* gattn;              // [B,Q,M,L,P] f32
};

template <typename TIn, typename TG, int LPH>
__global__ void __launch_bounds__(256) deform_attn_bwd_kernel(const DeformBwdParams bp) {
    using VT = VecTraits<TIn>;
    constexpr int VE = VT::VE, PR = VT::P;
    constexpr int PAIRS = 32 / LPH;
    const DeformParams& p = bp.f;
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long npairs = (long long)p.B * p.Q * p.M;
    long long pair = warp * PAIRS + lane / LPH;
    const bool live = pair < npairs;
    if (!live) pair = 0;
    const int sub = lane % LPH;
    const int m = (int)(pair % p.M);
    const int b = (int)(pair / p.M / p.Q);
    const int LP = p.L * p.P;
    const float2* loc = reinterpret_cast<const float2*>(p.loc) + pair * LP;
    const float* aw = p.attn + pair * LP;
    const long long head0 = ((long long)b * p.S * p.M + m) * p.D + sub * VE;   // element offset of this lane's channels at position 0
    const TIn* vb = reinterpret_cast<const TIn*>(p.value) + head0;
    float* gvb = bp.gvalue + head0;
    const long long pos_stride = (long long)p.M * p.D;                          // elements per spatial position

    // this lane's slice of dL/dout
    float g[VE];
    {
        const TG* gp = reinterpret_cast<const TG*>(bp.gout) + pair * p.D + sub * VE;
#pragma unroll
        for (int e = 0; e < VE; ++e) g[e] = live ? load_f32(gp + e) : 0.0f;
    }
    for (int s = 0; s < LP; ++s) {
        const int l = s / p.P;
        const int H = __ldg(p.shapes + 2 * l), W = __ldg(p.shapes + 2 * l + 1);
        const long long st = __ldg(p.start + l);
        const float2 xy = __ldg(loc + s);
        const float a = __ldg(aw + s);
        const float x = __fmaf_rn(xy.x, (float)W, -0.5f);
        const float y = __fmaf_rn(xy.y, (float)H, -0.5f);
        const bool inside = y > -1.0f && x > -1.0f && y < (float)H && x < (float)W;
        float t = 0.0f, gx = 0.0f, gy = 0.0f;
        if (inside) {   // (uniform within the head group: all its lanes share the sample)
            const float x0f = floorf(x), y0f = floorf(y);
            const float lx = x - x0f, ly = y - y0f, hx = 1.0f - lx, hy = 1.0f - ly;
            const int x0 = (int)x0f, y0 = (int)y0f;
            const bool xw = x0 >= 0, xe = x0 + 1 <= W - 1, yn = y0 >= 0, ysb = y0 + 1 <= H - 1;
            const bool ok[4] = {yn && xw, yn && xe, ysb && xw, ysb && xe};
            const float wt[4] = {hy * hx, hy * lx, ly * hx, ly * lx};
            float v[4][VE];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const long long pos = (st + (long long)(y0 + (k >> 1)) * W + (x0 + (k & 1))) * pos_stride;
                float2 f[PR];
#pragma unroll
                for (int e = 0; e < PR; ++e) f[e] = make_float2(0.0f, 0.0f);
                if (ok[k]) VT::unpack(ldg16(reinterpret_cast<const uint4*>(vb + pos)), f);   // one 16-byte load per tap and lane
#pragma unroll
                for (int e = 0; e < PR; ++e) { v[k][2 * e] = f[e].x; v[k][2 * e + 1] = f[e].y; }
                if (ok[k] && live) {
                    const float c = wt[k] * a;
#pragma unroll
                    for (int e = 0; e < VE; e += 4)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(gvb + pos + e), "f"(g[e] * c), "f"(g[e + 1] * c),
                                     "f"(g[e + 2] * c), "f"(g[e + 3] * c) : "memory");
                }
            }
#pragma unroll
            for (int e = 0; e < VE; ++e) {
                t += g[e] * (wt[0] * v[0][e] + wt[1] * v[1][e] + wt[2] * v[2][e] + wt[3] * v[3][e]);
                gx += g[e] * (hy * (v[1][e] - v[0][e]) + ly * (v[3][e] - v[2][e]));
                gy += g[e] * (hx * (v[2][e] - v[0][e]) + lx * (v[3][e] - v[1][e]));
            }
        }
#pragma unroll
        for (int o = LPH / 2; o > 0; o >>= 1) {
            t += __shfl_xor_sync(0xffffffffu, t, o, LPH);
            gx += __shfl_xor_sync(0xffffffffu, gx, o, LPH);
            gy += __shfl_xor_sync(0xffffffffu, gy, o, LPH);
        }
        if (live && sub == 0) {
            bp.gattn[pair * LP + s] = t;
            reinterpret_cast<float2*>(bp.gloc)[pair * LP + s] = make_float2(gx * (float)W * a, gy * (float)H * a);
        }
    }
}

}  // namespace bevipm
