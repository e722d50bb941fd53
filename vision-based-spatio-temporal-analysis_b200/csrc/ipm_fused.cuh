// ipm_fused.cuh -- the fused warp-and-fuse gather kernel (NHWC fast path) for sm_100a.
//
// Replaces, in one launch, geometry.py:120-162 (per-view projection + bilinear zero-padded
// sampling) and fusion.py:17-22 (sum / mean / max over views) of the reference; NONE mode
// writes the per-view maps GeometryTransformer returns (geometry.py:163).
//
// Work decomposition (see DESIGN.md "Kernel"):
//   CTA   = a TH x CELLS patch of BEV cells x `chunks_per_cta` channel chunks, all V views
//   phase A: every (view, cell) of the patch is projected ONCE (cell_coord + make_tap) by one
//            thread and parked in shared memory as a 32-byte CellTap
//   phase B: a warp owns (patch row, channel chunk); lane l owns NV 16-byte channel vectors
//            (vector n = channels [chunk0 + (n*32 + l)*VE, +VE)), so every tap is read by the
//            warp as NV fully coalesced 512-byte requests.  The warp walks its CELLS cells
//            view by view, accumulating across views in registers, and writes each BEV cell
//            exactly once with 16-byte stores.
//   REUSE: while walking a row the 2x2 texel block is kept in registers; because the cell is
//            warp-uniform the "did the block move?" test is a uniform branch, and a one-texel
//            move reloads only the new column/row.  This cuts L1 requests 2-5x when the BEV
//            grid is denser than the source map (BASELINE config 3).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ipm_geometry.cuh"

namespace bevipm {

struct FwdParams {
    const void* feats;
    void* out;
    const float* K;
    const float* Rt;
    const float* xs;
    const float* ys;
    int B, V, C, Hf, Wf, Hb, Wb;
    float sw, sh;   // (float)(Wf / (double)img_w), (float)(Hf / (double)img_h)   geometry.py:151-152
    int mode;       // bevipm_mode
    long long fs_b, fs_v, fs_c, fs_y, fs_x;
    long long os_b, os_v, os_c, os_y, os_x;
    int tiles_x, tiles_y;
    int chunks;          // channel chunks in total
    int chunks_per_cta;  // channel chunks one CTA walks
};

enum { KM_ACC = 0, KM_MAX = 1, KM_NONE = 2 };

// ---- 16-byte vector <-> fp32 pairs -------------------------------------------------------
template <typename T> struct VecTraits;
template <> struct VecTraits<float> {
    static constexpr int VE = 4;  // elements per 16 bytes
    static constexpr int P = 2;   // float2 pairs per vector
    __device__ static __forceinline__ void unpack(const uint4& r, float2 (&f)[2]) {
        f[0] = make_float2(__uint_as_float(r.x), __uint_as_float(r.y));
        f[1] = make_float2(__uint_as_float(r.z), __uint_as_float(r.w));
    }
};
template <> struct VecTraits<__nv_bfloat16> {
    static constexpr int VE = 8;
    static constexpr int P = 4;
    // bf16 -> fp32 is exact: the bf16 bits are the high half of the fp32 bits
    __device__ static __forceinline__ void unpack(const uint4& r, float2 (&f)[4]) {
        f[0] = make_float2(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xffff0000u));
        f[1] = make_float2(__uint_as_float(r.y << 16), __uint_as_float(r.y & 0xffff0000u));
        f[2] = make_float2(__uint_as_float(r.z << 16), __uint_as_float(r.z & 0xffff0000u));
        f[3] = make_float2(__uint_as_float(r.w << 16), __uint_as_float(r.w & 0xffff0000u));
    }
};

__device__ __forceinline__ uint4 ldg16(const void* p) {
    return __ldg(reinterpret_cast<const uint4*>(p));
}

// P float2 pairs -> global memory as TOut, streaming (written once, never re-read by us)
template <typename TOut, int P>
__device__ __forceinline__ void store_pairs(TOut* dst, const float2 (&f)[P]) {
    if constexpr (sizeof(TOut) == 4) {
#pragma unroll
        for (int k = 0; k < P; k += 2)
            __stcs(reinterpret_cast<float4*>(dst) + (k >> 1), make_float4(f[k].x, f[k].y, f[k + 1].x, f[k + 1].y));
    } else {
        uint32_t w[P];
#pragma unroll
        for (int k = 0; k < P; ++k) {
            __nv_bfloat162 h = __float22bfloat162_rn(f[k]);
            w[k] = *reinterpret_cast<uint32_t*>(&h);
        }
        if constexpr (P == 4) {
            __stcs(reinterpret_cast<uint4*>(dst), make_uint4(w[0], w[1], w[2], w[3]));
        } else {
            __stcs(reinterpret_cast<uint2*>(dst), make_uint2(w[0], w[1]));
        }
    }
}

// ---- phase A: project the patch ------------------------------------------------------------
__device__ __forceinline__ void project_patch(const FwdParams& p, int b, int i0, int j0, int TH, int TW,
                                              CellTap* taps, float* sH) {
    const int tid = threadIdx.x;
    if (tid < p.V) homography(p.K + 9 * (b * p.V + tid), p.Rt + 12 * (b * p.V + tid), sH + 9 * tid);
    __syncthreads();
    const int per_view = TH * TW;
    const float Wm = (float)p.Wf, Hm = (float)p.Hf;
    for (int idx = tid; idx < p.V * per_view; idx += blockDim.x) {
        const int v = idx / per_view;
        const int rem = idx - v * per_view;
        const int r = rem / TW, q = rem - r * TW;
        const int i = i0 + r, j = j0 + q;
        CellTap t;
        if (i < p.Hb && j < p.Wb) {
            float ix, iy;
            cell_coord(sH + 9 * v, __ldg(p.xs + j), __ldg(p.ys + i), p.sw, p.sh, Wm, Hm, ix, iy);
            t = make_tap(ix, iy, p.Wf, p.Hf);
        } else {
            t.x0 = t.y0 = -2; t.nw = t.ne = t.sw = t.se = 0.0f; t.flags = 0; t.pad = 0;
        }
        taps[idx] = t;
    }
    __syncthreads();
}

// ---- phase B helpers --------------------------------------------------------------------
// out_v = fma(SE,se, fma(SW,sw, fma(NE,ne, NW*nw)))  -- ATen's interpolation order, two channels
// per instruction with the sm_100 packed-fp32 pipe (FMUL2 / FFMA2: IEEE fp32 per element).
template <typename TIn, int NV>
__device__ __forceinline__ void blend(const uint4 (&raw)[4][NV], const CellTap& t,
                                      float2 (&o)[NV][VecTraits<TIn>::P]) {
    constexpr int P = VecTraits<TIn>::P;
    const float2 nw = make_float2(t.nw, t.nw), ne = make_float2(t.ne, t.ne);
    const float2 sw = make_float2(t.sw, t.sw), se = make_float2(t.se, t.se);
#pragma unroll
    for (int n = 0; n < NV; ++n) {
        float2 a[P], bq[P], c[P], d[P];
        VecTraits<TIn>::unpack(raw[0][n], a);
        VecTraits<TIn>::unpack(raw[1][n], bq);
        VecTraits<TIn>::unpack(raw[2][n], c);
        VecTraits<TIn>::unpack(raw[3][n], d);
#pragma unroll
        for (int k = 0; k < P; ++k) {
            float2 r = __fmul2_rn(a[k], nw);
            r = __ffma2_rn(bq[k], ne, r);
            r = __ffma2_rn(c[k], sw, r);
            r = __ffma2_rn(d[k], se, r);
            o[n][k] = r;
        }
    }
}

template <typename TIn, typename TOut, int NV, int CELLS, int TH, int KMODE, bool REUSE>
__global__ void __launch_bounds__(256) warp_fuse_nhwc_kernel(const FwdParams p) {
    using VT = VecTraits<TIn>;
    constexpr int VE = VT::VE, P = VT::P;
    constexpr int CH_CHUNK = 32 * NV * VE;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CellTap* taps = reinterpret_cast<CellTap*>(smem_raw);
    float* sH = reinterpret_cast<float*>(taps + p.V * TH * CELLS);

    const int b = blockIdx.z;
    const int ty = blockIdx.x / p.tiles_x, tx = blockIdx.x - ty * p.tiles_x;
    const int i0 = ty * TH, j0 = tx * CELLS;
    project_patch(p, b, i0, j0, TH, CELLS, taps, sH);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = blockDim.x >> 5;
    const TIn* fb = reinterpret_cast<const TIn*>(p.feats) + (long long)b * p.fs_b;
    TOut* ob = reinterpret_cast<TOut*>(p.out) + (long long)b * p.os_b;
    const float qnan = __int_as_float(0x7fc00000);

    const int items = TH * p.chunks_per_cta;
    for (int item = warp; item < items; item += nwarps) {
        const int r = item % TH;
        const int k = blockIdx.y * p.chunks_per_cta + item / TH;
        const int i = i0 + r;
        if (i >= p.Hb || k >= p.chunks) continue;  // warp-uniform
        int cvec[NV];
        bool cok[NV];
#pragma unroll
        for (int n = 0; n < NV; ++n) {
            cvec[n] = k * CH_CHUNK + (n * 32 + lane) * VE;
            cok[n] = cvec[n] < p.C;
        }

        float2 acc[CELLS][NV][P];
        if constexpr (KMODE != KM_NONE) {
            const float init = (KMODE == KM_MAX) ? -INFINITY : 0.0f;
#pragma unroll
            for (int q = 0; q < CELLS; ++q)
#pragma unroll
                for (int n = 0; n < NV; ++n)
#pragma unroll
                    for (int e = 0; e < P; ++e) acc[q][n][e] = make_float2(init, init);
        }

        for (int v = 0; v < p.V; ++v) {
            const TIn* fv = fb + (long long)v * p.fs_v;
            const CellTap* row = taps + (v * TH + r) * CELLS;
            uint4 raw[4][NV];
            int cx = -(1 << 24), cy = -(1 << 24);  // origin of the cached 2x2 block (REUSE); "none yet"
#pragma unroll
            for (int q = 0; q < CELLS; ++q) {
                const CellTap t = row[q];  // warp-uniform broadcast, 2 x LDS.128
                float2 o[NV][P];
                bool have = false;
                if (t.flags & kTapMask) {
                    const TIn* base = fv + (long long)t.y0 * p.fs_y + (long long)t.x0 * p.fs_x;
                    // tap k lives at base + (k&1)*fs_x + (k>>1)*fs_y
                    auto load_tap = [&](int tap) {
                        const bool ok = (t.flags >> tap) & 1;
                        const TIn* tp = base + (tap & 1 ? p.fs_x : 0) + (tap & 2 ? p.fs_y : 0);
#pragma unroll
                        for (int n = 0; n < NV; ++n)
                            raw[tap][n] = (ok && cok[n]) ? ldg16(tp + cvec[n]) : make_uint4(0, 0, 0, 0);
                    };
                    if constexpr (REUSE) {
                        const int dx = t.x0 - cx, dy = t.y0 - cy;
                        if (dx == 0 && dy == 0) {
                            // same 2x2 block: nothing to fetch
                        } else if (dy == 0 && dx == 1) {
#pragma unroll
                            for (int n = 0; n < NV; ++n) { raw[0][n] = raw[1][n]; raw[2][n] = raw[3][n]; }
                            load_tap(1); load_tap(3);
                        } else if (dy == 0 && dx == -1) {
#pragma unroll
                            for (int n = 0; n < NV; ++n) { raw[1][n] = raw[0][n]; raw[3][n] = raw[2][n]; }
                            load_tap(0); load_tap(2);
                        } else if (dx == 0 && dy == 1) {
#pragma unroll
                            for (int n = 0; n < NV; ++n) { raw[0][n] = raw[2][n]; raw[1][n] = raw[3][n]; }
                            load_tap(2); load_tap(3);
                        } else if (dx == 0 && dy == -1) {
#pragma unroll
                            for (int n = 0; n < NV; ++n) { raw[2][n] = raw[0][n]; raw[3][n] = raw[1][n]; }
                            load_tap(0); load_tap(1);
                        } else {
                            load_tap(0); load_tap(1); load_tap(2); load_tap(3);
                        }
                        cx = t.x0; cy = t.y0;
                    } else {
                        load_tap(0); load_tap(1); load_tap(2); load_tap(3);
                    }
                    blend<TIn, NV>(raw, t, o);
                    have = true;
                } else if (t.flags & kNonFinite) {
                    // reference: weights are NaN and 0 * NaN = NaN reaches every channel
#pragma unroll
                    for (int n = 0; n < NV; ++n)
#pragma unroll
                        for (int e = 0; e < P; ++e) o[n][e] = make_float2(qnan, qnan);
                    have = true;
                }
                if constexpr (KMODE == KM_ACC) {
                    // fusion.py:18-21  sequential fp32 accumulation over views (a view that misses
                    // the cell contributes exactly +0: skipped)
                    if (have) {
#pragma unroll
                        for (int n = 0; n < NV; ++n)
#pragma unroll
                            for (int e = 0; e < P; ++e) acc[q][n][e] = __fadd2_rn(acc[q][n][e], o[n][e]);
                    }
                } else if constexpr (KMODE == KM_MAX) {
                    // fusion.py:22  the zeros of out-of-view cells take part; NaN propagates
#pragma unroll
                    for (int n = 0; n < NV; ++n)
#pragma unroll
                        for (int e = 0; e < P; ++e) {
                            const float sx = have ? o[n][e].x : 0.0f, sy = have ? o[n][e].y : 0.0f;
                            float2& m = acc[q][n][e];
                            m.x = (sx > m.x || sx != sx) ? sx : m.x;
                            m.y = (sy > m.y || sy != sy) ? sy : m.y;
                        }
                } else {
                    // per-view maps (geometry.py:162): written straight out, zero where the view misses
                    const int j = j0 + q;
                    if (j < p.Wb) {
                        TOut* oc = ob + (long long)v * p.os_v + (long long)i * p.os_y + (long long)j * p.os_x;
#pragma unroll
                        for (int n = 0; n < NV; ++n) {
                            if (!cok[n]) continue;
                            float2 z[P];
#pragma unroll
                            for (int e = 0; e < P; ++e) z[e] = have ? o[n][e] : make_float2(0.0f, 0.0f);
                            store_pairs<TOut, P>(oc + cvec[n], z);
                        }
                    }
                }
            }
        }

        if constexpr (KMODE != KM_NONE) {
            const float Vf = (float)p.V;
#pragma unroll
            for (int q = 0; q < CELLS; ++q) {
                const int j = j0 + q;
                if (j >= p.Wb) continue;
                TOut* oc = ob + (long long)i * p.os_y + (long long)j * p.os_x;
#pragma unroll
                for (int n = 0; n < NV; ++n) {
                    if (!cok[n]) continue;
                    if (KMODE == KM_ACC && p.mode == 1 /* BEVIPM_MEAN: sum / V, IEEE division */) {
#pragma unroll
                        for (int e = 0; e < P; ++e) {
                            acc[q][n][e].x = __fdiv_rn(acc[q][n][e].x, Vf);
                            acc[q][n][e].y = __fdiv_rn(acc[q][n][e].y, Vf);
                        }
                    }
                    store_pairs<TOut, P>(oc + cvec[n], acc[q][n]);
                }
            }
        }
    }
}

}  // namespace bevipm
