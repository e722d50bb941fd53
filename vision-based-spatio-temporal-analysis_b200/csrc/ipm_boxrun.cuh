// ipm_boxrun.cuh -- the run kernel (ipm_run.cuh) with its per-warp ring filled by the TMA unit: ONE tensor-map
// [256 channels x 2 x 2 texels] box copy per reload instead of four per-lane cp.async requests.
//
// Why: the run kernel's memory side is bound by the LSU, not by bytes -- 18 M LDGSTS warp-instructions per BASELINE
// config-1 launch at 8 cycles each per SM (B300_MICROARCH.md: "LDGSTS rt 8.0 cyc/op") are 0.5 ms on their own, and each
// reload spends ~17 issue slots on four 64-bit tap addresses, four LDGSTS and ring bookkeeping.  A box copy is issued by one
// elected lane (descriptor + five coordinates), does not touch the LSU / L1 at all, lands the 2x2 block as NW, NE, SW, SE
// at +0 / +512 / +1024 / +1536 -- the ring layout the walk already reads -- and ZERO-FILLS the taps outside the map: the
// reference's zero padding itself (geometry.py:161), so the stand-in addresses of the cp.async form and its Inf * 0
// deviation on map borders are gone.  Completion is counted in bytes on one mbarrier per ring stage and warp; there is
// still no cross-warp synchronisation of any kind.  (Round 1 tried four 512-byte bulk copies per reload: 60 issue slots per
// reload, 1.7x slower; the box needs one instruction.)
//
// Everything else -- phase A tables in walking order, view-major walk with the unpacked block kept in registers, frame
// groups sharing tables, exact mean division, epilogue -- is the run kernel's; see ipm_run.cuh.
#pragma once
#include <cuda.h>

#include "ipm_run.cuh"
#include "ipm_staged.cuh"

namespace bevipm {

template <typename TIn, typename TOut, int CELLS, int NW, int MAXREG, int DEPTH, int KMODE>
__global__ void __maxnreg__(MAXREG) warp_fuse_boxrun_kernel(const FwdParams p, int fpc, const __grid_constant__ CUtensorMap tmap) {
    static_assert(DEPTH >= 2 && DEPTH <= 8 && (DEPTH & (DEPTH - 1)) == 0, "ring depth: a power of two");
    static_assert(KMODE == KM_ACC || KMODE == KM_MAX, "sum / mean / max");
    using VT = VecTraits<TIn>;
    constexpr int VE = VT::VE, P = VT::P;
    constexpr int R = NW, NT = NW * 32;
    constexpr int ILP = (MAXREG <= 128 && P > 2 && CELLS >= 8) ? 2 : P;
    extern __shared__ __align__(128) unsigned char smem_box[];

    const int V = p.V;
    const int seg_bytes = run_seg_bytes(V, CELLS);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int ty = blockIdx.x / p.tiles_x, tx = blockIdx.x - ty * p.tiles_x;
    const int i0 = ty * R, j0 = tx * CELLS;
    const int b0 = blockIdx.z * fpc, b1 = min(p.B, b0 + fpc);
    const int r = warp, i = i0 + r;
    const int chunks = (p.C + 32 * VE - 1) / (32 * VE);
    const int last_vec = p.C / VE - 1;
    const float Vf = (float)V;
    unsigned char* seg = smem_box + r * seg_bytes;
    float4* seg_wts = reinterpret_cast<float4*>(seg);
    int4* seg_loads = reinterpret_cast<int4*>(seg + V * CELLS * 16);
    int* seg_meta = reinterpret_cast<int*>(seg + V * CELLS * 16 + (V * CELLS + 8) * 16);
    const uint32_t s_wts = (uint32_t)__cvta_generic_to_shared(seg_wts);
    const uint32_t s_loads = (uint32_t)__cvta_generic_to_shared(seg_loads);
    const uint32_t s_meta = (uint32_t)__cvta_generic_to_shared(seg_meta);
    const uint32_t ring0 = (uint32_t)__cvta_generic_to_shared(smem_box) + run_tables_bytes(V, CELLS, R) + warp * (DEPTH * 2048);
    uint32_t ring = ring0 + lane * 16;  // this lane's 16 bytes of stage 0 / tap 0
    asm volatile("" : "+r"(ring));
    uint32_t bars = (uint32_t)__cvta_generic_to_shared(smem_box) + run_tables_bytes(V, CELLS, R) + NW * (DEPTH * 2048) + V * 48 + warp * (DEPTH * 8);
    uint32_t n_issue = 0, n_cons = 0;  // ring entries issued / consumed so far (stage = count % DEPTH, parity = (count / DEPTH) & 1)
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < DEPTH; ++s) mbar_init(bars + s * 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("" : "+r"(bars));

    for (int b = b0; b < b1;) {
        if (b > b0) __syncthreads();
        float* sH = reinterpret_cast<float*>(smem_box + run_tables_bytes(V, CELLS, R) + NW * (DEPTH * 2048));
        if (tid < V) {
            float H[9];
            homography(p.K + 9 * (b * V + tid), p.Rt + 12 * (b * V + tid), H);
#pragma unroll
            for (int q = 0; q < 3; ++q) reinterpret_cast<float4*>(sH + 12 * tid)[q] = make_float4(H[3 * q], H[3 * q + 1], H[3 * q + 2], 0.0f);
        }
        __syncthreads();
        run_build_tables<CELLS, KMODE == KM_MAX, false, false, true>(p, V, i, j0, lane, 0, sH, seg_wts, seg_loads, seg_meta);
        __syncwarp();

        // ---- the run of frames b .. e-1 shares these tables: same calibration, bit for bit -----------------------
        int e = b1;
        if (b + 1 < b1) {
            bool differs = false;
            const int per = 21 * V, n = per * (b1 - b - 1);
            for (int z = tid; z < n; z += NT) {
                const int f = z / per, q = z - f * per;
                const float* cur = q < 9 * V ? p.K + (size_t)(b + 1 + f) * 9 * V + q : p.Rt + (size_t)(b + 1 + f) * 12 * V + (q - 9 * V);
                const float* prv = q < 9 * V ? cur - 9 * V : cur - 12 * V;
                differs |= __float_as_uint(__ldg(cur)) != __float_as_uint(__ldg(prv));
            }
            if (__syncthreads_or(differs)) {
                for (e = b + 1; e < b1; ++e) {
                    bool d = false;
                    for (int q = tid; q < per; q += NT) {
                        const float* cur = q < 9 * V ? p.K + (size_t)e * 9 * V + q : p.Rt + (size_t)e * 12 * V + (q - 9 * V);
                        const float* prv = q < 9 * V ? cur - 9 * V : cur - 12 * V;
                        d |= __float_as_uint(__ldg(cur)) != __float_as_uint(__ldg(prv));
                    }
                    if (__syncthreads_or(d)) break;
                }
            }
        }
        const int b_run = b;
        const int n_items = (e - b) * chunks;  // this warp's (frame, chunk) items, frame-major
        b = e;
        if (i >= p.Hb || n_items <= 0) continue;
        const int nviews = __shfl_sync(0xffffffffu, lds4i(s_meta + 8 * V), 0);

        // one box copy: the 2x2 block of entry `o` of item (frame fb, channel offset c0) into the next free stage
        auto issue = [&](const int4& o, int c0, int fb) {
            __syncwarp();  // every lane has read the stage this copy overwrites
            if (lane == 0) {
                const uint32_t st = n_issue & (DEPTH - 1), bar = bars + st * 8;
                mbar_expect_tx(bar, 2048);
                tma_load_5d(ring0 + st * 2048, &tmap, c0, (int)(short)(o.x & 0xffff), o.x >> 16, o.y, fb, bar);
            }
            ++n_issue;
        };
        auto prime = [&](int c0, int fb) {
#pragma unroll
            for (int s = 0; s < DEPTH - 1; ++s) {
                const int4 o = lds16i(s_loads + s * 16);
                if (o.y >= 0) issue(o, c0, fb);
            }
        };
        int fi_c = 0, k_c = 0;
        if (nviews > 0) prime(0, b_run);
        for (int it = 0; it < n_items; ++it) {
            float2 acc[CELLS][P];
#pragma unroll
            for (int c = 0; c < CELLS; ++c)
#pragma unroll
                for (int q = 0; q < P; ++q) acc[c][q] = (KMODE == KM_MAX) ? make_float2(-INFINITY, -INFINITY) : make_float2(0.0f, 0.0f);
            const bool lok = k_c * 32 + lane <= last_vec;
            if (nviews > 0) {
                float2 cur[4][P];
#pragma unroll
                for (int tap = 0; tap < 4; ++tap)
#pragma unroll
                    for (int q = 0; q < P; ++q) cur[tap][q] = make_float2(0.0f, 0.0f);
                const int c0 = k_c * 32 * VE, fb = b_run + fi_c;
                uint32_t lp = s_loads + (DEPTH - 1) * 16;  // the entry the next reload hands to the copy engine
                for (int vi = 0; vi < nviews; ++vi) {
                    const int v = lds4i(s_meta + 4 * (V + vi));
                    const unsigned m = (unsigned)__shfl_sync(0xffffffffu, lds4i(s_meta + 4 * v), 0);
                    const uint32_t wv = s_wts + v * (CELLS * 16);
                    float4 wn = lds16f(wv);
                    bool rl = (m >> 16) & 1u;
#pragma unroll
                    for (int c = 0; c < CELLS; ++c) {
                        const float4 w = wn;
                        if (c + 1 < CELLS) wn = lds16f(wv + (c + 1) * 16);
                        const bool seen = (m >> c) & 1u;
                        const bool rl_now = rl;
                        if (c + 1 < CELLS) rl = (m >> (17 + c)) & 1u;
                        if (rl_now) {  // the row leaves the block held in `cur`
                            const uint32_t stc = n_cons & (DEPTH - 1);
                            mbar_wait(bars + stc * 8, (n_cons / DEPTH) & 1u);  // the block's 2048 bytes have landed
                            const uint32_t sr = ring + stc * 2048;
                            uint4 nxt[4];
                            nxt[0] = lds16(sr); nxt[1] = lds16(sr + 512);
                            nxt[2] = lds16(sr + 1024); nxt[3] = lds16(sr + 1536);
                            const int4 o = lds16i(lp);
#pragma unroll
                            for (int tap = 0; tap < 4; ++tap) VT::unpack(nxt[tap], cur[tap]);
                            ++n_cons;
                            if (o.y >= 0) issue(o, c0, fb);  // the entry DEPTH-1 ahead goes into the stage read at the previous reload
                            lp += 16;
                        }
#pragma unroll
                        for (int q0 = 0; q0 < P; q0 += ILP) {
                            float2 sv[ILP];
#pragma unroll
                            for (int q = 0; q < ILP; ++q) sv[q] = __fmul2_rn(cur[0][q0 + q], make_float2(w.x, w.x));
#pragma unroll
                            for (int q = 0; q < ILP; ++q) sv[q] = __ffma2_rn(cur[1][q0 + q], make_float2(w.y, w.y), sv[q]);
#pragma unroll
                            for (int q = 0; q < ILP; ++q) sv[q] = __ffma2_rn(cur[2][q0 + q], make_float2(w.z, w.z), sv[q]);
#pragma unroll
                            for (int q = 0; q < ILP; ++q) sv[q] = __ffma2_rn(cur[3][q0 + q], make_float2(w.w, w.w), sv[q]);
#pragma unroll
                            for (int q = 0; q < ILP; ++q)
                                if (seen) {
                                    if constexpr (KMODE == KM_MAX) {
                                        float2& mx = acc[c][q0 + q];
                                        mx.x = max_nan(mx.x, sv[q].x);
                                        mx.y = max_nan(mx.y, sv[q].y);
                                    } else {
                                        acc[c][q0 + q] = __fadd2_rn(acc[c][q0 + q], sv[q]);
                                    }
                                }
                        }
                    }
                }
            }
            const int k_this = k_c, fi_this = fi_c;
            if (++k_c >= chunks) { k_c = 0; ++fi_c; }
            if (nviews > 0 && it + 1 < n_items) prime(k_c * 32 * VE, b_run + fi_c);  // the next item's first blocks fly during the epilogue

            if constexpr (KMODE == KM_MAX) {
                const unsigned every = (unsigned)__shfl_sync(0xffffffffu, lds4i(s_meta + 8 * V + 8), 0);
#pragma unroll
                for (int c = 0; c < CELLS; ++c)
                    if (!((every >> c) & 1u)) {
#pragma unroll
                        for (int q = 0; q < P; ++q) {
                            acc[c][q].x = max_nan(acc[c][q].x, 0.0f);
                            acc[c][q].y = max_nan(acc[c][q].y, 0.0f);
                        }
                    }
            } else if (p.mode == 1) {
                float2 t[CELLS];
#pragma unroll
                for (int c = 0; c < CELLS; ++c) {
                    float2 u = acc[c][0];
#pragma unroll
                    for (int q = 1; q < P; ++q) u = __fadd2_rn(u, acc[c][q]);
                    t[c] = u;
                }
#pragma unroll
                for (int w = 1; w < CELLS; w *= 2)
#pragma unroll
                    for (int c = 0; c + w < CELLS; c += 2 * w) t[c] = __fadd2_rn(t[c], t[c + w]);
                const float tot = __fadd_rn(t[0].x, t[0].y);
                if (fabsf(tot) <= 3.402823466e+38f) {
                    const float rr = p.rcpV;
#pragma unroll
                    for (int c = 0; c < CELLS; ++c)
#pragma unroll
                        for (int q = 0; q < P; ++q) {
                            const float2 qq = __fmul2_rn(acc[c][q], make_float2(rr, rr));
                            const float2 rem = __ffma2_rn(qq, make_float2(-Vf, -Vf), acc[c][q]);
                            acc[c][q] = __ffma2_rn(rem, make_float2(rr, rr), qq);
                        }
                } else {
#pragma unroll
                    for (int c = 0; c < CELLS; ++c)
#pragma unroll
                        for (int q = 0; q < P; ++q) {
                            acc[c][q].x = __fdiv_rn(acc[c][q].x, Vf);
                            acc[c][q].y = __fdiv_rn(acc[c][q].y, Vf);
                        }
                }
            }
            TOut* oc = reinterpret_cast<TOut*>(p.out) + (long long)(b_run + fi_this) * p.os_b + (long long)i * p.os_y + (long long)j0 * p.os_x +
                       (k_this * 32 + lane) * VE;
            if (!lok) continue;
            if (j0 + CELLS <= p.Wb) {
#pragma unroll
                for (int c = 0; c < CELLS; ++c) {
                    store_pairs<TOut, P>(oc, acc[c]);
                    oc += p.os_x;
                }
            } else {
#pragma unroll
                for (int c = 0; c < CELLS; ++c) {
                    if (j0 + c < p.Wb) store_pairs<TOut, P>(oc, acc[c]);
                    oc += p.os_x;
                }
            }
        }
    }
}

}  // namespace bevipm
