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
//            warp as NV fully coalesced 512-byte requests.  The warp walks (view, cell) steps,
//            accumulating across views in registers, and writes each BEV cell exactly once with
//            16-byte stores.  The taps of step s+1 are requested before step s is blended
//            (two register buffers), so the loads of one step hide behind the math of the last.
//   The cell is warp-uniform, so "are all four taps inside the map?" is a uniform branch:
//   interior cells take loads without predicates, border cells a predicated zero-filling path,
//   cells a view does not see cost two LDS and a branch.
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
    int chunks_per_cta;  // channel chunks one CTA walks (strided / backward kernels)
    int chunk_groups;    // fused kernel: channel chunks == tile groups along C
    int total_tiles;     // fused kernel: tiles_x * tiles_y * chunk_groups * B
    int fsy16, fsx16;    // fs_y, fs_x in 16-byte units (fast path)
    float rcpV;          // RN(1 / V) for the exact mean division
    float kx, ky;        // 0 = reference grid_sample geometry; Wf/(Wf-1), Hf/(Hf-1) = kornia-compatible sample positions
    // KM_RED (view sharding over peer memory): BEV rows [q * slab_rows, (q + 1) * slab_rows) live in slab[q], the
    // fp32 buffer of the rank that owns them (its own memory or a peer's, mapped over NVLink); strides os_* are a slab's
    void* slab[16];
    int slab_rows;
    int slab_put;        // 0: add into the slab (bulk reduction); 1: plain bulk store (every rank has its own receive buffer at the owner)
    // table cache across launches (run kernel, PLAN instantiations; ipm_run.cuh): a device buffer holding the phase-A tables of
    // every row segment for ONE calibration, and the key of everything else they depend on (shapes, strides, axes, kernel shape)
    void* plan;
    unsigned long long plan_key;
};

// Header of a table cache (plan): 1 KB.  key == 0: empty (the first launch that finds it empty fills the tables from its frame 0
// and its last CTA publishes the header); otherwise the tables are valid for exactly this key and this calibration, bit for bit.
struct PlanHeader {
    unsigned long long key;
    unsigned int done;        // CTAs that have written their tables (reset by the publishing CTA)
    unsigned int pad;
    unsigned int calib[21 * 32 + 4];  // K (9 V) and Rt34 (12 V) of the calibration the tables were built from, as bit patterns
};
constexpr int kPlanHeaderBytes = 4096;

enum { KM_ACC = 0, KM_MAX = 1, KM_NONE = 2, KM_PROBE = 3 /* timing probe: loads only, no blend */,
       KM_RED = 4 /* sum over this rank's views, ADDED (red.global.add) into the row slab of the owning rank */ };

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

__device__ __forceinline__ uint4 ldg16(const uint4* p) { return __ldg(p); }

// torch.max semantics in one instruction: NaN if either operand is NaN (fusion.py:22 propagates NaN)
__device__ __forceinline__ float max_nan(float a, float b) {
    float d;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
    return d;
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

// P float2 pairs ADDED to fp32 global memory (local or peer) with 16-byte reductions: the partial sum of one rank's views
// goes straight into the owner's slab, it never lands in this rank's HBM
template <int P>
__device__ __forceinline__ void red_pairs(float* dst, const float2 (&f)[P]) {
#pragma unroll
    for (int k = 0; k < P; k += 2)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 2 * k), "f"(f[k].x), "f"(f[k].y), "f"(f[k + 1].x), "f"(f[k + 1].y) : "memory");
}

// x / V, correctly rounded, in three FMA-pipe ops instead of the ~10-instruction IEEE division
// sequence: q = RN(x * r), e = x - q * V (exact in one fma, denormals included: flush-to-zero is off),
// result = RN(q + e * r) with r = RN(1/V) -- Markstein's correction step, which returns the IEEE
// quotient for every finite x (checked on the device against true division from denormals to
// FLT_MAX, tests/test_gpu_parity.py::test_mean_division_is_ieee).  +-Inf would turn into NaN
// (Inf - Inf), so a vector with any non-finite element takes the library division.
template <int P>
__device__ __forceinline__ void div_exact_vec(float2 (&a)[P], float Vf, float r) {
    float mag = 0.0f;
#pragma unroll
    for (int e = 0; e < P; ++e) mag = __fadd_rn(__fadd_rn(mag, fabsf(a[e].x)), fabsf(a[e].y));
    if (mag <= 3.402823466e+38f) {  // every element finite and the sum did not overflow
        const float2 r2 = make_float2(r, r), nV = make_float2(-Vf, -Vf);
#pragma unroll
        for (int e = 0; e < P; ++e) {
            const float2 q = __fmul2_rn(a[e], r2);
            const float2 rem = __ffma2_rn(q, nV, a[e]);
            a[e] = __ffma2_rn(rem, r2, q);
        }
    } else {
#pragma unroll
        for (int e = 0; e < P; ++e) {
            a[e].x = __fdiv_rn(a[e].x, Vf);
            a[e].y = __fdiv_rn(a[e].y, Vf);
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
            cell_coord(sH + 9 * v, __ldg(p.xs + j), __ldg(p.ys + i), p.sw, p.sh, Wm, Hm, ix, iy, p.kx, p.ky);
            t = make_tap(ix, iy, p.Wf, p.Hf, p.fsy16, p.fsx16);
        } else {
            t.x0 = t.y0 = -2; t.nw = t.ne = t.sw = t.se = 0.0f; t.flags = 0; t.off16 = 0;
        }
        taps[idx] = t;
    }
    __syncthreads();
}

// ---- phase B helpers --------------------------------------------------------------------
// What phase B keeps in registers for one (view, cell) step: the first 24 bytes of the CellTap.
struct StepHdr {
    int off16, flags;
    float nw, ne, sw, se;
};

__device__ __forceinline__ StepHdr read_hdr(const CellTap* t) {
    const int4 a = *reinterpret_cast<const int4*>(t);
    const float2 b = *reinterpret_cast<const float2*>(reinterpret_cast<const char*>(t) + 16);
    StepHdr h;
    h.off16 = a.x; h.flags = a.y; h.nw = __int_as_float(a.z); h.ne = __int_as_float(a.w); h.sw = b.x; h.se = b.y;
    return h;
}

// Request the four taps of one (view, cell) step.  `vb` points at this lane's first vector of the
// view; tap k lives at vb[off16 + (k&1)*dx16 + (k>>1)*dy16], vector n 32 vectors further.  Indices
// are 32-bit so each address is one IMAD.WIDE.
template <int NV>
__device__ __forceinline__ void request_taps(uint4 (&raw)[4][NV], const uint4* vb, const StepHdr& h, bool full,
                                             const bool (&cok)[NV], int dx16, int dy16) {
    const int tm = h.flags & kTapMask;
    if (tm == kTapMask && full) {  // interior cell, full channel chunk: no predicates, no zero fill
        const uint4* p0 = vb + h.off16;
        const uint4* p1 = vb + (h.off16 + dx16);
        const uint4* p2 = vb + (h.off16 + dy16);
        const uint4* p3 = vb + (h.off16 + dy16 + dx16);
#pragma unroll
        for (int n = 0; n < NV; ++n) {
            raw[0][n] = ldg16(p0 + n * 32);
            raw[1][n] = ldg16(p1 + n * 32);
            raw[2][n] = ldg16(p2 + n * 32);
            raw[3][n] = ldg16(p3 + n * 32);
        }
    } else if (tm) {               // border cell (taps outside the map read as zero) or partial chunk
#pragma unroll
        for (int tap = 0; tap < 4; ++tap) {
            const bool ok = (tm >> tap) & 1;
            const uint4* tp = vb + (h.off16 + ((tap & 1) ? dx16 : 0) + ((tap & 2) ? dy16 : 0));
#pragma unroll
            for (int n = 0; n < NV; ++n)
                raw[tap][n] = (ok && cok[n]) ? ldg16(tp + n * 32) : make_uint4(0u, 0u, 0u, 0u);
        }
    }
}

// out_v = fma(SE,se, fma(SW,sw, fma(NE,ne, NW*nw)))  -- ATen's interpolation order, two channels
// per instruction with the sm_100 packed-fp32 pipe (FMUL2 / FFMA2: IEEE fp32 per element).
template <typename TIn, int NV>
__device__ __forceinline__ void blend(const uint4 (&raw)[4][NV], const StepHdr& t,
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

// Fold one step's per-view value into the cell's accumulator / write it out (NONE).
template <typename TIn, typename TOut, int NV, int KMODE>
__device__ __forceinline__ void consume(const uint4 (&raw)[4][NV], const StepHdr& t,
                                        float2 (&acc)[NV][VecTraits<TIn>::P], TOut* oc, const int (&cvec)[NV],
                                        const bool (&cok)[NV], bool store_ok) {
    constexpr int P = VecTraits<TIn>::P;
    if constexpr (KMODE == KM_PROBE) {
        // measurement aid (never dispatched by the public modes): fold the raw taps with integer adds
        if (t.flags & kTapMask) {
#pragma unroll
            for (int n = 0; n < NV; ++n) {
                const uint4 s = raw[0][n], u = raw[1][n], w = raw[2][n], z = raw[3][n];
                acc[n][0].x = __uint_as_float(__float_as_uint(acc[n][0].x) + s.x + u.x + w.x + z.x);
                acc[n][0].y = __uint_as_float(__float_as_uint(acc[n][0].y) + s.y + u.y + w.y + z.y);
                acc[n][1].x = __uint_as_float(__float_as_uint(acc[n][1].x) + s.z + u.z + w.z + z.z);
                acc[n][1].y = __uint_as_float(__float_as_uint(acc[n][1].y) + s.w + u.w + w.w + z.w);
            }
        }
        return;
    }
    float2 o[NV][P];
    bool have = false;
    if (t.flags & kTapMask) {
        blend<TIn, NV>(raw, t, o);
        have = true;
    } else if (t.flags & kNonFinite) {
        // reference: weights are NaN and 0 * NaN = NaN reaches every channel
        const float qnan = __int_as_float(0x7fc00000);
#pragma unroll
        for (int n = 0; n < NV; ++n)
#pragma unroll
            for (int e = 0; e < P; ++e) o[n][e] = make_float2(qnan, qnan);
        have = true;
    }
    if constexpr (KMODE == KM_ACC) {
        // fusion.py:18-21  sequential fp32 accumulation over views (a view that misses the cell
        // contributes exactly +0: skipped)
        if (have) {
#pragma unroll
            for (int n = 0; n < NV; ++n)
#pragma unroll
                for (int e = 0; e < P; ++e) acc[n][e] = __fadd2_rn(acc[n][e], o[n][e]);
        }
    } else if constexpr (KMODE == KM_MAX) {
        // fusion.py:22  the zeros of out-of-view cells take part; NaN propagates
#pragma unroll
        for (int n = 0; n < NV; ++n)
#pragma unroll
            for (int e = 0; e < P; ++e) {
                const float sx = have ? o[n][e].x : 0.0f, sy = have ? o[n][e].y : 0.0f;
                float2& m = acc[n][e];
                m.x = (sx > m.x || sx != sx) ? sx : m.x;
                m.y = (sy > m.y || sy != sy) ? sy : m.y;
            }
    } else {
        // per-view maps (geometry.py:162): written straight out, zero where the view misses
        if (store_ok) {
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

template <typename TIn, typename TOut, int NV, int CELLS, int TH, int KMODE, int MINB, bool PIPE>
__global__ void __launch_bounds__(256, MINB) warp_fuse_nhwc_kernel(const FwdParams p) {
    using VT = VecTraits<TIn>;
    constexpr int VE = VT::VE, P = VT::P;
    constexpr int CH_CHUNK = 32 * NV * VE;
    static_assert(!PIPE || CELLS % 2 == 0, "the two tap buffers alternate per cell");
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
    const int dx16 = p.fsx16, dy16 = p.fsy16;
    const long long fsv16 = p.fs_v / VE;

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
        const bool full = (k + 1) * CH_CHUNK <= p.C;
        // this lane's first vector of view 0
        const uint4* vb = reinterpret_cast<const uint4*>(fb) + (k * (CH_CHUNK / VE) + lane);
        TOut* orow = ob + (long long)i * p.os_y + (long long)j0 * p.os_x;

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

        const CellTap* row = taps + r * CELLS;  // view 0; + TH*CELLS per view
        if constexpr (PIPE) {
            // two register buffers: step s = v*CELLS + q uses buffer (q & 1); CELLS is even
            uint4 raw[2][4][NV];
            StepHdr hdr[2];
            hdr[0] = read_hdr(row);
            request_taps<NV>(raw[0], vb, hdr[0], full, cok, dx16, dy16);
            for (int v = 0; v < p.V; ++v) {
#pragma unroll
                for (int q = 0; q < CELLS; ++q) {
                    const int cur = q & 1, nxt = cur ^ 1;
                    if (q + 1 < CELLS) {
                        hdr[nxt] = read_hdr(row + q + 1);
                        request_taps<NV>(raw[nxt], vb, hdr[nxt], full, cok, dx16, dy16);
                    } else if (v + 1 < p.V) {
                        hdr[nxt] = read_hdr(row + TH * CELLS);
                        request_taps<NV>(raw[nxt], vb + fsv16, hdr[nxt], full, cok, dx16, dy16);
                    }
                    consume<TIn, TOut, NV, KMODE>(raw[cur], hdr[cur], acc[q], orow + (long long)v * p.os_v + (long long)q * p.os_x,
                                                  cvec, cok, j0 + q < p.Wb);
                }
                row += TH * CELLS;
                vb += fsv16;
            }
        } else {
            for (int v = 0; v < p.V; ++v) {
                // Walking a BEV row, consecutive cells often fall into the SAME 2x2 texel block when the
                // BEV grid is denser than the source map: then the registers already hold the taps and
                // the step costs no L1 request at all (warp-uniform test).
                uint4 raw[4][NV];
                int held_off = 0, held_flags = 0;  // flags 0 = nothing held
#pragma unroll
                for (int q = 0; q < CELLS; ++q) {
                    const StepHdr h = read_hdr(row + q);
                    if (CELLS == 1 || h.off16 != held_off || h.flags != held_flags) {
                        request_taps<NV>(raw, vb, h, full, cok, dx16, dy16);
                        if (h.flags & kTapMask) { held_off = h.off16; held_flags = h.flags; }
                    }
                    consume<TIn, TOut, NV, KMODE>(raw, h, acc[q], orow + (long long)v * p.os_v + (long long)q * p.os_x, cvec, cok,
                                                  j0 + q < p.Wb);
                }
                row += TH * CELLS;
                vb += fsv16;
            }
        }

        if constexpr (KMODE != KM_NONE) {
            const float Vf = (float)p.V;  // (KM_PROBE stores its folded bits the same way)
#pragma unroll
            for (int q = 0; q < CELLS; ++q) {
                if (j0 + q >= p.Wb) continue;
                TOut* oc = orow + (long long)q * p.os_x;
#pragma unroll
                for (int n = 0; n < NV; ++n) {
                    if (!cok[n]) continue;
                    if (KMODE == KM_ACC && p.mode == 1 /* BEVIPM_MEAN: sum / V, IEEE quotient */)
                        div_exact_vec<P>(acc[q][n], Vf, p.rcpV);
                    store_pairs<TOut, P>(oc + cvec[n], acc[q][n]);
                }
            }
        }
    }
}

}  // namespace bevipm
