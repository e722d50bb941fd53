// ipm_list.cuh -- branch-free, software-pipelined form of the fused warp-and-fuse kernel.
//
// Same arithmetic as ipm_fused.cuh (geometry.py:120-162 + fusion.py:17-22 of the reference), other
// control structure.  ncu on the first kernel showed (profiles/): the L1 request pipe and the issue
// slots are each worth ~0.5 ms on BASELINE config 1, and they ADD instead of overlapping, because
// a warp alternates "request taps -> wait -> blend" and ptxas tags every LDG with one scoreboard
// slot as soon as the requests sit under (warp-uniform) branches.  Here each warp first compacts
// its row into a LIST of steps and then runs a straight-line loop over it:
//
//   phase A (per warp, no block barrier): lane = (cell, view) pair of the warp's BEV row; project,
//            ballot the pairs some view actually sees, and store one 64-byte StepRec per live pair
//            in walking order.  Taps outside the map are redirected to a zero page (so they read
//            as exact zeros with their true weight, the reference's 0 * w), cells no view sees get
//            one all-zero step, a trailing pair of dummy steps lets the loop over-request safely.
//   phase B: for s in steps: request taps of s+1 (16-byte x NV per lane, unconditional) ; blend s.
//            No predicate, no branch except "last step of this cell -> divide, store, reset".
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ipm_fused.cuh"

namespace bevipm {

// 4 KB of zeros: where out-of-map taps are read from.  A lane reads at most NV*512 bytes past its base.
__device__ __align__(16) unsigned char g_zero_page[4096];

struct __align__(16) StepRec {
    long long off[4];  // byte offset of each tap from the frame's chunk base (lane 0, vector 0)
    float w[4];        // nw, ne, sw, se
    int meta;          // bits 0..7 cell index in the row, 8 last step of the cell, 9 cell misses >= 1 view,
                       // 16..23 view index
    int pad[3];
};
static_assert(sizeof(StepRec) == 64, "StepRec is 64 bytes");

constexpr int kListCells = 8;   // cells per warp row (fewer when V > 8)
constexpr int kListMaxSteps = 64;

template <typename TIn, typename TOut, int NV, int KMODE, int NWARPS, int MINB>
__global__ void __launch_bounds__(NWARPS * 32, MINB) warp_fuse_list_kernel(const FwdParams p, int tw) {
    using VT = VecTraits<TIn>;
    constexpr int VE = VT::VE, P = VT::P;
    constexpr int CH_CHUNK = 32 * NV * VE;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    StepRec* steps = reinterpret_cast<StepRec*>(smem_raw) + warp * (kListMaxSteps + 2);

    const int b = blockIdx.z, k = blockIdx.y;
    const int ty = blockIdx.x / p.tiles_x, tx = blockIdx.x - ty * p.tiles_x;
    const int i = ty * NWARPS + warp, j0 = tx * tw;
    if (i >= p.Hb) return;  // whole warp; there is no block-level barrier in this kernel

    const char* cbase = reinterpret_cast<const char*>(reinterpret_cast<const TIn*>(p.feats) + (long long)b * p.fs_b +
                                                      (long long)k * CH_CHUNK);
    const long long zero_off = reinterpret_cast<const char*>(g_zero_page) - cbase;
    const long long fsv_b = p.fs_v * (long long)sizeof(TIn);

    // ---- phase A: this warp's row -> compact step list ---------------------------------------------
    const int V = p.V, E = tw * V;  // E <= 64
    unsigned long long live = 0;
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
            cell_coord(H, __ldg(p.xs + j0 + q), __ldg(p.ys + i), p.sw, p.sh, (float)p.Wf, (float)p.Hf, ix, iy);
            t = make_tap(ix, iy, p.Wf, p.Hf, p.fsy16, p.fsx16);
            in_grid = true;
        }
        mine[pass] = t;
        // NONE writes every view's map, so every in-grid pair is a step; the fusing modes only walk
        // the pairs a view actually sees
        const bool act = in_grid && (KMODE == KM_NONE || t.flags != 0);
        live |= (unsigned long long)__ballot_sync(0xffffffffu, act) << (32 * pass);
    }
    const unsigned long long vmask = (V >= 64) ? ~0ull : ((1ull << V) - 1ull);
    unsigned long long full_mask = live;
    // cells that no view sees still owe an all-zero output: their view-0 pair becomes a zero step
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        const int e = lane + 32 * pass;
        const int q = e / V, v = e - q * V;
        const bool dummy = (e < E) && (j0 + q < p.Wb) && v == 0 && ((live >> (q * V)) & vmask) == 0;
        full_mask |= (unsigned long long)__ballot_sync(0xffffffffu, dummy) << (32 * pass);
    }
    const int n = __popcll(full_mask);
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        const int e = lane + 32 * pass;
        if (e < E && ((full_mask >> e) & 1ull)) {
            const int q = e / V, v = e - q * V;
            const CellTap t = mine[pass];
            const int pos = __popcll(full_mask & ((1ull << e) - 1ull));
            const unsigned long long cell_bits = (full_mask >> (q * V)) & vmask;
            const bool last = (cell_bits >> (v + 1)) == 0;
            const bool misses = ((live >> (q * V)) & vmask) != vmask;
            StepRec r;
            const long long vb = (long long)v * fsv_b;
#pragma unroll
            for (int tap = 0; tap < 4; ++tap) {
                const bool ok = (t.flags >> tap) & 1;
                const long long o = vb + 16ll * (t.off16 + ((tap & 1) ? p.fsx16 : 0) + ((tap & 2) ? p.fsy16 : 0));
                r.off[tap] = ok ? o : zero_off;
            }
            const bool nf = (t.flags & kNonFinite) != 0;
            const bool none = (t.flags & (kTapMask | kNonFinite)) == 0;  // zero step
            const float qnan = __int_as_float(0x7fc00000);
            r.w[0] = nf ? qnan : (none ? 0.0f : t.nw);
            r.w[1] = nf ? qnan : (none ? 0.0f : t.ne);
            r.w[2] = nf ? qnan : (none ? 0.0f : t.sw);
            r.w[3] = nf ? qnan : (none ? 0.0f : t.se);
            r.meta = q | (last ? 0x100 : 0) | (misses ? 0x200 : 0) | (v << 16);
            r.pad[0] = r.pad[1] = r.pad[2] = 0;
            steps[pos] = r;
        }
    }
    if (lane < 2) {  // two trailing dummies: the loop requests one step ahead without testing
        StepRec r;
        r.off[0] = r.off[1] = r.off[2] = r.off[3] = zero_off;
        r.w[0] = r.w[1] = r.w[2] = r.w[3] = 0.0f;
        r.meta = 0; r.pad[0] = r.pad[1] = r.pad[2] = 0;
        steps[n + lane] = r;
    }
    __syncwarp();
    if (n == 0) return;

    // ---- phase B: straight-line pipelined walk --------------------------------------------------------
    int cvec[NV];
    bool cok[NV];
#pragma unroll
    for (int nn = 0; nn < NV; ++nn) {
        cvec[nn] = k * CH_CHUNK + (nn * 32 + lane) * VE;
        cok[nn] = cvec[nn] < p.C;
    }
    // partial last chunk: lanes past C read the zero page instead (their results are never stored)
    const char* lbase = cbase + lane * 16;
    TOut* orow = reinterpret_cast<TOut*>(p.out) + (long long)b * p.os_b + (long long)i * p.os_y + (long long)j0 * p.os_x;
    const float Vf = (float)V;

    float2 acc[NV][P];
    const float init = (KMODE == KM_MAX) ? -INFINITY : 0.0f;
#pragma unroll
    for (int nn = 0; nn < NV; ++nn)
#pragma unroll
        for (int e = 0; e < P; ++e) acc[nn][e] = make_float2(init, init);

    // ptxas tags every LDG of this kernel with ONE scoreboard slot, so "wait for the older batch" would
    // also wait for a younger one already in flight.  `after` makes the addresses of the next batch
    // depend (through an opaque AND with 0) on a register of the batch being waited for: the wait
    // happens first, with a single batch outstanding, then the next batch is requested and flies
    // while the current one is blended.
    auto request = [&](uint4 (&raw)[4][NV], const StepRec& r, unsigned after) {
        unsigned z;
        asm volatile("and.b32 %0, %1, 0;" : "=r"(z) : "r"(after));
        const char* lb = lbase + z;
#pragma unroll
        for (int tap = 0; tap < 4; ++tap) {
            const char* tp = lb + r.off[tap];
#pragma unroll
            for (int nn = 0; nn < NV; ++nn) {
                const char* a = cok[nn] ? tp + nn * 512 : reinterpret_cast<const char*>(g_zero_page);
                raw[tap][nn] = ldg16(reinterpret_cast<const uint4*>(a));
            }
        }
    };
    auto finish = [&](const StepRec& r, const uint4 (&raw)[4][NV]) {
        StepHdr h;
        h.off16 = 0; h.flags = kTapMask; h.nw = r.w[0]; h.ne = r.w[1]; h.sw = r.w[2]; h.se = r.w[3];
        float2 o[NV][P];
        blend<TIn, NV>(raw, h, o);
        const int q = r.meta & 0xff;
        if constexpr (KMODE == KM_NONE) {
            TOut* oc = orow + (long long)((r.meta >> 16) & 0xff) * p.os_v + (long long)q * p.os_x;
#pragma unroll
            for (int nn = 0; nn < NV; ++nn)
                if (cok[nn]) store_pairs<TOut, P>(oc + cvec[nn], o[nn]);
            return;
        } else {
#pragma unroll
            for (int nn = 0; nn < NV; ++nn)
#pragma unroll
                for (int e = 0; e < P; ++e) {
                    if constexpr (KMODE == KM_MAX) {
                        float2& m = acc[nn][e];
                        m.x = (o[nn][e].x > m.x || o[nn][e].x != o[nn][e].x) ? o[nn][e].x : m.x;
                        m.y = (o[nn][e].y > m.y || o[nn][e].y != o[nn][e].y) ? o[nn][e].y : m.y;
                    } else {
                        acc[nn][e] = __fadd2_rn(acc[nn][e], o[nn][e]);  // fusion.py:18-21, view order kept
                    }
                }
            if (r.meta & 0x100) {  // last view of this cell: finish, store once, reset   (warp-uniform)
                TOut* oc = orow + (long long)q * p.os_x;
#pragma unroll
                for (int nn = 0; nn < NV; ++nn) {
#pragma unroll
                    for (int e = 0; e < P; ++e) {
                        if constexpr (KMODE == KM_MAX) {
                            if (r.meta & 0x200) {  // fusion.py:22: the zeros of views that miss the cell take part
                                acc[nn][e].x = (0.0f > acc[nn][e].x) ? 0.0f : acc[nn][e].x;
                                acc[nn][e].y = (0.0f > acc[nn][e].y) ? 0.0f : acc[nn][e].y;
                            }
                        } else if (p.mode == 1) {
                            acc[nn][e].x = div_exact(acc[nn][e].x, Vf, p.rcpV);
                            acc[nn][e].y = div_exact(acc[nn][e].y, Vf, p.rcpV);
                        }
                    }
                    if (cok[nn]) store_pairs<TOut, P>(oc + cvec[nn], acc[nn]);
#pragma unroll
                    for (int e = 0; e < P; ++e) acc[nn][e] = make_float2(init, init);
                }
            }
        }
    };

    uint4 rawA[4][NV], rawB[4][NV];
    StepRec recA = steps[0], recB;
    request(rawA, recA, 0u);
    for (int s = 0; s < n; s += 2) {
        recB = steps[s + 1];
        request(rawB, recB, rawA[3][NV - 1].w);
        finish(recA, rawA);
        if (s + 1 >= n) break;
        recA = steps[s + 2];
        request(rawA, recA, rawB[3][NV - 1].w);
        finish(recB, rawB);
    }
}

}  // namespace bevipm
