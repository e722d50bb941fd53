// ipm_list.cuh -- branch-free, software-pipelined form of the fused warp-and-fuse kernel.
//
// Same arithmetic as ipm_fused.cuh (geometry.py:120-162 + fusion.py:17-22 of the reference), other
// control structure.  ncu on the first kernel (profiles/) showed the L1 request pipe and the issue
// slots each worth ~0.5 ms on BASELINE config 1 and ADDING instead of overlapping: a warp
// alternates "request taps -> wait -> blend", and ptxas tags every LDG of the kernel with one
// scoreboard slot, so a second batch in flight does not help -- waiting for the older batch waits
// for the younger one too.  Here each warp compacts its BEV row into a LIST of live steps and walks
// it with straight-line code:
//
//   phase A (per warp, no block barrier): lane = (cell, view) pair of the warp's row; project,
//            ballot the pairs some view actually sees, store one 32-byte StepRec per live pair in
//            walking order (cell-major, views ascending: the reference's accumulation order).
//            A tap outside the map gets weight 0 and the address of one of the cell's in-map taps
//            (it contributes exactly +0 for finite features); cells no view sees are not steps at
//            all, their zeros are written after the walk.
//   phase B: for s in steps: [wait batch s] -> request taps of s+1 -> blend s.  The request's
//            addresses are made to depend on a register of batch s (an opaque AND with 0), which
//            pins the order "wait, then request": one batch is outstanding at every wait, and it
//            flies while the previous one is blended.  No predicates; the only branch is
//            "last view of this cell -> divide, store, reset".
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ipm_fused.cuh"

namespace bevipm {

struct __align__(16) StepRec {
    int off[4];  // tap offsets from the lane's chunk base, in 16-byte units (view offset included)
    float w[4];  // nw, ne, sw, se (0 for taps outside the map, NaN for non-finite sample positions)
};
static_assert(sizeof(StepRec) == 32, "StepRec is 32 bytes");

constexpr int kListCells = 8;   // cells per warp row (fewer when V > 8)
constexpr int kListMaxSteps = 64;
// per-warp shared memory: records (+2 trailing copies so the walk may request one step ahead) and one
// meta word per step: bits 0..7 cell index in the row, 8 last step of its cell, 9 cell misses >= 1 view
constexpr int kListWarpBytes = (kListMaxSteps + 2) * (int)sizeof(StepRec) + (((kListMaxSteps + 2) * 4 + 15) / 16) * 16;
static_assert(kListWarpBytes % 16 == 0, "every warp's records start 16-byte aligned");

template <typename TIn, typename TOut, int NV, int KMODE, bool FULL>
__device__ __forceinline__ void walk_steps(const FwdParams& p, const StepRec* steps, const int* meta, int n,
                                           const uint4* lbase, TOut* orow, int k, int lane) {
    using VT = VecTraits<TIn>;
    constexpr int VE = VT::VE, P = VT::P;
    constexpr int CH_CHUNK = 32 * NV * VE;
    int cvec[NV];
    bool cok[NV];
#pragma unroll
    for (int nn = 0; nn < NV; ++nn) {
        cvec[nn] = k * CH_CHUNK + (nn * 32 + lane) * VE;
        cok[nn] = FULL || cvec[nn] < p.C;
    }
    const float Vf = (float)p.V;
    const float init = (KMODE == KM_MAX) ? -INFINITY : 0.0f;
    float2 acc[NV][P];
#pragma unroll
    for (int nn = 0; nn < NV; ++nn)
#pragma unroll
        for (int e = 0; e < P; ++e) acc[nn][e] = make_float2(init, init);

    auto request = [&](uint4 (&raw)[4][NV], const StepRec& r, unsigned after) {
        unsigned z;
        asm volatile("and.b32 %0, %1, 0;" : "=r"(z) : "r"(after));  // opaque zero: orders this batch after `after`
        const uint4* lb = lbase + z;
#pragma unroll
        for (int tap = 0; tap < 4; ++tap) {
            const uint4* tp = lb + r.off[tap];
#pragma unroll
            for (int nn = 0; nn < NV; ++nn) {
                if (FULL) raw[tap][nn] = ldg16(tp + nn * 32);
                else raw[tap][nn] = cok[nn] ? ldg16(tp + nn * 32) : make_uint4(0u, 0u, 0u, 0u);
            }
        }
    };
    auto finish = [&](const StepRec& r, int m, const uint4 (&raw)[4][NV]) {
        StepHdr h;
        h.off16 = 0; h.flags = kTapMask; h.nw = r.w[0]; h.ne = r.w[1]; h.sw = r.w[2]; h.se = r.w[3];
        float2 o[NV][P];
        blend<TIn, NV>(raw, h, o);
#pragma unroll
        for (int nn = 0; nn < NV; ++nn)
#pragma unroll
            for (int e = 0; e < P; ++e) {
                if constexpr (KMODE == KM_MAX) {
                    float2& mx = acc[nn][e];
                    mx.x = max_nan(mx.x, o[nn][e].x);
                    mx.y = max_nan(mx.y, o[nn][e].y);
                } else {
                    acc[nn][e] = __fadd2_rn(acc[nn][e], o[nn][e]);  // fusion.py:18-21, view order kept
                }
            }
        if (m & 0x100) {  // last view of this cell: finish, store once, reset   (warp-uniform)
            TOut* oc = orow + (long long)(m & 0xff) * p.os_x;
#pragma unroll
            for (int nn = 0; nn < NV; ++nn) {
                if constexpr (KMODE == KM_MAX) {
                    if (m & 0x200) {  // fusion.py:22: the zeros of views that miss the cell take part
#pragma unroll
                        for (int e = 0; e < P; ++e) {
                            acc[nn][e].x = max_nan(acc[nn][e].x, 0.0f);
                            acc[nn][e].y = max_nan(acc[nn][e].y, 0.0f);
                        }
                    }
                } else if (p.mode == 1) {
                    div_exact_vec<P>(acc[nn], Vf, p.rcpV);  // BEVIPM_MEAN: sum / V, IEEE quotient
                }
                if (cok[nn]) store_pairs<TOut, P>(oc + cvec[nn], acc[nn]);
#pragma unroll
                for (int e = 0; e < P; ++e) acc[nn][e] = make_float2(init, init);
            }
        }
    };

    uint4 rawA[4][NV], rawB[4][NV];
    StepRec recA = steps[0], recB;
    int mA = meta[0], mB;
    request(rawA, recA, 0u);
    for (int s = 0; s < n; s += 2) {
        recB = steps[s + 1];
        mB = meta[s + 1];
        request(rawB, recB, rawA[3][NV - 1].w);
        finish(recA, mA, rawA);
        if (s + 1 >= n) break;
        recA = steps[s + 2];
        mA = meta[s + 2];
        request(rawA, recA, rawB[3][NV - 1].w);
        finish(recB, mB, rawB);
    }
}

template <typename TIn, typename TOut, int NV, int KMODE, int NWARPS, int MINB>
__global__ void __launch_bounds__(NWARPS * 32, MINB) warp_fuse_list_kernel(const FwdParams p, int tw) {
    using VT = VecTraits<TIn>;
    constexpr int VE = VT::VE, P = VT::P;
    constexpr int CH_CHUNK = 32 * NV * VE;
    static_assert(KMODE == KM_ACC || KMODE == KM_MAX, "per-view output stays with the tile kernel");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    StepRec* steps = reinterpret_cast<StepRec*>(smem_raw + warp * kListWarpBytes);
    int* meta = reinterpret_cast<int*>(steps + kListMaxSteps + 2);

    const int b = blockIdx.z, k = blockIdx.y;
    const int ty = blockIdx.x / p.tiles_x, tx = blockIdx.x - ty * p.tiles_x;
    const int i = ty * NWARPS + warp, j0 = tx * tw;
    if (i >= p.Hb) return;  // whole warp; there is no block-level barrier in this kernel

    // ---- phase A: this warp's row -> compact step list ---------------------------------------------
    const int V = p.V, E = tw * V;  // E <= 64
    const int fsv16 = (int)(p.fs_v / VE);
    unsigned long long live = 0, in_row = 0;
    CellTap mine[2];
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        const int e = lane + 32 * pass;
        const int q = e / V, v = e - q * V;
        CellTap t;
        t.flags = 0;
        bool in_grid = false;
        if (e < E && j0 + q < p.Wb) {
            float H[9], ix, iy;
            homography(p.K + 9 * (b * V + v), p.Rt + 12 * (b * V + v), H);
            cell_coord(H, __ldg(p.xs + j0 + q), __ldg(p.ys + i), p.sw, p.sh, (float)p.Wf, (float)p.Hf, ix, iy, p.kx, p.ky);
            t = make_tap(ix, iy, p.Wf, p.Hf, p.fsy16, p.fsx16);
            in_grid = true;
        }
        mine[pass] = t;
        live |= (unsigned long long)__ballot_sync(0xffffffffu, in_grid && t.flags != 0) << (32 * pass);
        in_row |= (unsigned long long)__ballot_sync(0xffffffffu, in_grid) << (32 * pass);
    }
    const unsigned long long vmask = (1ull << V) - 1ull;  // V <= 32
    const int n = __popcll(live);
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        const int e = lane + 32 * pass;
        if (e < E && ((live >> e) & 1ull)) {
            const int q = e / V, v = e - q * V;
            const CellTap t = mine[pass];
            const int pos = __popcll(live & ((1ull << e) - 1ull));
            const unsigned long long cell_bits = (live >> (q * V)) & vmask;
            const bool last = (cell_bits >> (v + 1)) == 0;
            const bool misses = cell_bits != vmask;
            StepRec r;
            const int vo = v * fsv16;
            const bool nf = (t.flags & kNonFinite) != 0;
            const float qnan = __int_as_float(0x7fc00000);
            // an in-map tap of this cell: where the out-of-map ones (weight 0) are pointed
            const int tm = t.flags & kTapMask;
            const int first = tm ? (__ffs(tm) - 1) : 0;
            const int safe = nf ? vo : vo + t.off16 + ((first & 1) ? p.fsx16 : 0) + ((first & 2) ? p.fsy16 : 0);
            const float w[4] = {t.nw, t.ne, t.sw, t.se};
#pragma unroll
            for (int tap = 0; tap < 4; ++tap) {
                const bool ok = (tm >> tap) & 1;
                r.off[tap] = ok ? vo + t.off16 + ((tap & 1) ? p.fsx16 : 0) + ((tap & 2) ? p.fsy16 : 0) : safe;
                r.w[tap] = nf ? qnan : (ok ? w[tap] : 0.0f);
            }
            steps[pos] = r;
            meta[pos] = q | (last ? 0x100 : 0) | (misses ? 0x200 : 0);
            if (pos == n - 1) {  // trailing copies: requested by the walk, never blended
                steps[n] = r;
                steps[n + 1] = r;
                meta[n] = meta[n + 1] = 0;
            }
        }
    }
    __syncwarp();

    const TIn* fb = reinterpret_cast<const TIn*>(p.feats) + (long long)b * p.fs_b;
    const uint4* lbase = reinterpret_cast<const uint4*>(fb) + (k * (CH_CHUNK / VE) + lane);
    TOut* orow = reinterpret_cast<TOut*>(p.out) + (long long)b * p.os_b + (long long)i * p.os_y + (long long)j0 * p.os_x;
    const bool full = (k + 1) * CH_CHUNK <= p.C;

    // ---- phase B: straight-line pipelined walk --------------------------------------------------------
    if (n > 0) {
        if (full) walk_steps<TIn, TOut, NV, KMODE, true>(p, steps, meta, n, lbase, orow, k, lane);
        else walk_steps<TIn, TOut, NV, KMODE, false>(p, steps, meta, n, lbase, orow, k, lane);
    }

    // ---- cells of this row that no view sees: exact zeros (sum, mean and max alike) ---------------------
    for (int q = 0; q < tw; ++q) {
        if (!((in_row >> (q * V)) & 1ull)) break;            // past the grid edge
        if (((live >> (q * V)) & vmask) != 0) continue;      // walked above
        TOut* oc = orow + (long long)q * p.os_x;
        float2 z[P];
#pragma unroll
        for (int e = 0; e < P; ++e) z[e] = make_float2(0.0f, 0.0f);
#pragma unroll
        for (int nn = 0; nn < NV; ++nn) {
            const int c = k * CH_CHUNK + (nn * 32 + lane) * VE;
            if (c < p.C) store_pairs<TOut, P>(oc + c, z);
        }
    }
}

}  // namespace bevipm
