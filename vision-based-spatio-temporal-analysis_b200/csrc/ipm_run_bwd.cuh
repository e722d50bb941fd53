// ipm_run_bwd.cuh -- backward of the fused warp-and-fuse w.r.t. the features, in the run kernel's walking order.
//
// Autograd of geometry.py:161 (grid_sample, bilinear, zeros) followed by fusion.py:18-21: every BEV cell scatters
// w_tap * dL/dBEV (divided by V for the mean) to its in-map taps.  The generic kernel (ipm_aux.cuh) issues one
// 16-byte atomic per (cell, view, tap, 4 channels) and sits on the L2 atomic units (0.43 ms at BASELINE config 0,
// 4x the forward).  Here the contributions of consecutive cells of a row that fall into the SAME 2x2 texel block
// (1.76x .. 2.4x of them, tools/reuse_stats.py) are summed in registers first and flushed with four atomics when the
// row leaves the block: same tables as the forward (run_build_tables), out-of-map taps marked instead of redirected.
// Sums are re-associated (atomics already are): tolerance 1e-5 relative against the reference's autograd, as before.
#pragma once
#include "ipm_run.cuh"

namespace bevipm {

// FwdParams here: feats = grad_feats (fp32, pre-zeroed, channels-last, strides fs_*), out = grad_out (TG, strides os_*).
template <typename TG, int CELLS, int NW>
__global__ void __launch_bounds__(NW * 32) warp_fuse_run_bwd_kernel(const FwdParams p) {
    constexpr int R = NW;  // one warp per row segment
    extern __shared__ __align__(128) unsigned char smem_run[];  // (its own symbol: the other kernels declare theirs 16-byte aligned)
    const int V = p.V;
    const int seg_bytes = run_seg_bytes(V, CELLS);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int b = blockIdx.z;
    const int ty = blockIdx.x / p.tiles_x, tx = blockIdx.x - ty * p.tiles_x;
    const int i = ty * R + warp, j0 = tx * CELLS;
    float4* wts = reinterpret_cast<float4*>(smem_run + warp * seg_bytes);
    int4* loads = reinterpret_cast<int4*>(smem_run + warp * seg_bytes + V * CELLS * 16);
    int* ml = reinterpret_cast<int*>(smem_run + warp * seg_bytes + V * CELLS * 16 + (V * CELLS + 8) * 16);
    float* sH = reinterpret_cast<float*>(smem_run + run_tables_bytes(V, CELLS, R));

    if (tid < V) {
        float H[9];
        homography(p.K + 9 * (b * V + tid), p.Rt + 12 * (b * V + tid), H);
#pragma unroll
        for (int q = 0; q < 3; ++q)
            reinterpret_cast<float4*>(sH + 12 * tid)[q] = make_float4(H[3 * q], H[3 * q + 1], H[3 * q + 2], 0.0f);
    }
    __syncthreads();
    run_build_tables<CELLS, false, true>(p, V, i, j0, lane, (int)(p.fs_v / 4), sH, wts, loads, ml);
    __syncwarp();
    if (i >= p.Hb) return;

    const uint32_t s_wts = (uint32_t)__cvta_generic_to_shared(wts);
    const uint32_t s_loads = (uint32_t)__cvta_generic_to_shared(loads);
    const uint32_t s_meta = (uint32_t)__cvta_generic_to_shared(ml);
    const int nviews = __shfl_sync(0xffffffffu, lds4i(s_meta + 8 * V), 0);
    if (nviews == 0) return;
    const float Vf = (float)V;
    const int chunks = p.C / 128;  // 32 lanes x 4 fp32 channels
    const TG* grow = reinterpret_cast<const TG*>(p.out) + (long long)b * p.os_b + (long long)i * p.os_y + (long long)j0 * p.os_x;
    float4* gfb = reinterpret_cast<float4*>(reinterpret_cast<float*>(const_cast<void*>(p.feats)) + (long long)b * p.fs_b);

    auto load_g = [&](const TG* cellp, float2 (&g)[2]) {  // 4 channels of one cell's dL/dBEV
        if constexpr (sizeof(TG) == 4) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(cellp));
            g[0] = make_float2(t.x, t.y); g[1] = make_float2(t.z, t.w);
        } else {
            const uint2 t = __ldg(reinterpret_cast<const uint2*>(cellp));
            g[0] = make_float2(__uint_as_float(t.x << 16), __uint_as_float(t.x & 0xffff0000u));
            g[1] = make_float2(__uint_as_float(t.y << 16), __uint_as_float(t.y & 0xffff0000u));
        }
    };

    for (int k = 0; k < chunks; ++k) {
        const int ch = k * 128 + lane * 4;
        float4* gbase = gfb + (k * 32 + lane);  // this lane's 16 bytes of texel 0 of view 0
        float2 g[CELLS][2];
        if (p.mode != 3) {
#pragma unroll
            for (int c = 0; c < CELLS; ++c) {
                g[c][0] = g[c][1] = make_float2(0.0f, 0.0f);
                if (j0 + c < p.Wb) {
                    load_g(grow + (long long)c * p.os_x + ch, g[c]);
                    if (p.mode == 1) div_exact_vec<2>(g[c], Vf, p.rcpV);  // mean: d/dx (sum / V)
                }
            }
        }
        uint32_t lp = s_loads;
        for (int vi = 0; vi < nviews; ++vi) {
            const int v = lds4i(s_meta + 4 * (V + vi));
            const unsigned m = (unsigned)__shfl_sync(0xffffffffu, lds4i(s_meta + 4 * v), 0);  // warp-uniform
            const uint32_t wv = s_wts + v * (CELLS * 16);
            if (p.mode == 3) {  // per-view maps: every view has its own dL/dBEV
#pragma unroll
                for (int c = 0; c < CELLS; ++c) {
                    g[c][0] = g[c][1] = make_float2(0.0f, 0.0f);
                    if (((m >> c) & 1u) && j0 + c < p.Wb) load_g(grow + (long long)v * p.os_v + (long long)c * p.os_x + ch, g[c]);
                }
            }
            float2 ga[4][2];
#pragma unroll
            for (int t = 0; t < 4; ++t) ga[t][0] = ga[t][1] = make_float2(0.0f, 0.0f);
            int4 o = make_int4(-1, -1, -1, -1);
            auto flush = [&]() {  // the block's four taps: one 16-byte atomic each (out-of-map taps carry -1)
                if (o.x >= 0) atomicAdd(gbase + o.x, make_float4(ga[0][0].x, ga[0][0].y, ga[0][1].x, ga[0][1].y));
                if (o.y >= 0) atomicAdd(gbase + o.y, make_float4(ga[1][0].x, ga[1][0].y, ga[1][1].x, ga[1][1].y));
                if (o.z >= 0) atomicAdd(gbase + o.z, make_float4(ga[2][0].x, ga[2][0].y, ga[2][1].x, ga[2][1].y));
                if (o.w >= 0) atomicAdd(gbase + o.w, make_float4(ga[3][0].x, ga[3][0].y, ga[3][1].x, ga[3][1].y));
            };
#pragma unroll
            for (int c = 0; c < CELLS; ++c) {
                if ((m >> c) & 1u) {
                    if ((m >> (16 + c)) & 1u) {  // the row enters a new 2x2 block: flush the one before it
                        flush();
                        o = lds16i(lp);
                        lp += 16;
#pragma unroll
                        for (int t = 0; t < 4; ++t) ga[t][0] = ga[t][1] = make_float2(0.0f, 0.0f);
                    }
                    const float4 w = lds16f(wv + c * 16);
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        ga[0][q] = __ffma2_rn(g[c][q], make_float2(w.x, w.x), ga[0][q]);
                        ga[1][q] = __ffma2_rn(g[c][q], make_float2(w.y, w.y), ga[1][q]);
                        ga[2][q] = __ffma2_rn(g[c][q], make_float2(w.z, w.z), ga[2][q]);
                        ga[3][q] = __ffma2_rn(g[c][q], make_float2(w.w, w.w), ga[3][q]);
                    }
                }
            }
            flush();
        }
    }
}

}  // namespace bevipm
