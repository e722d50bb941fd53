// ipm_run.cuh -- "run" form of the fused warp-and-fuse kernel: taps are re-used in registers.
//
// Same arithmetic as ipm_list.cuh / ipm_fused.cuh (geometry.py:120-162 + fusion.py:17-22 of the
// reference), other walking order.  ncu on the list kernel (profiles/r01_notes.md) showed the two
// scarce units to be the L1 request path (4 taps x texel bytes per cell-view, 6.3x the HBM bytes) and
// the issue slots (bf16: 32 of the 56 instructions of a vector-visit only unpack taps).  Walking a BEV
// row, consecutive cells of one view fall into the SAME 2x2 texel block 1.76x (BASELINE config 1) to
// 2.4x (config 2) on average (tools/reuse_stats.py), so this kernel walks VIEW-major and keeps the
// unpacked block in registers for as long as the row stays inside it:
//
//   phase A (the first warp of every row segment; no block barrier when one warp owns a segment):
//            lane = (view, cell); project once (homographies come from a table built once per CTA), mark
//            the cell-views some view sees and, among them, the ones that START a new 2x2 block
//            ("reloads": the cell before is not seen or sits in another block).  Ballot prefix sums put
//            three tables of the segment into shared memory, directly in walking order (views ascending,
//            cells ascending): per-(view, cell) blend weights, the LOAD LIST (tap offsets of every reload)
//            and the list of views that see the segment with their seen / reload bit masks.
//   phase B (per warp = row segment x 512-byte channel chunk, chunks one after the other): accumulators
//            of all CELLS cells live in registers; for each view of the view list, for each cell
//            (unrolled, so the accumulator index is static): on a reload bit, read the oldest block of
//            the warp's cp.async ring, unpack it once into the `cur` registers and hand the load-list
//            entry DEPTH-1 ahead to the copy engine; then blend `cur` with the cell's weights and add
//            to the cell's accumulator (predicated on "seen").  Per cell the views are still added in
//            ascending order: the reference's accumulation order.
//   A tap outside the map gets weight 0 and the address of one of the block's in-map taps; the walk zeroes the
//   registers of such taps after the unpack (views that have a border block in the segment carry a flag and
//   per-tap cell masks, everybody else pays one uniform test per view), so the tap contributes w * 0 = +0
//   whatever the stand-in texel holds: the reference's zero padding, also next to +-Inf / NaN features.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "ipm_fused.cuh"

namespace bevipm {

constexpr int kRunMaxViews = 32;

// shared-memory bytes of the R row segments' tables, rounded up so that the rings behind them start on a 128-byte
// boundary: a warp's 512-byte ring access that straddles bank rows costs a fifth wavefront (measured: +14 % on c1)
__host__ __device__ constexpr int run_tables_bytes(int V, int cells, int R);
// shared-memory bytes of one row segment
__host__ __device__ constexpr int run_seg_bytes(int V, int cells) {
    return V * cells * 16                    // blend weights (nw, ne, sw, se) of every (view, cell)
           + (V * cells + 8) * 16            // the load list (+8: entries the walk reads ahead but never copies)
           + ((V * 20 + 12 + 15) / 16) * 16;  // per-view masks, view list, totals, cells every view sees, half-reload masks, border-tap masks
}

__host__ __device__ constexpr int run_tables_bytes(int V, int cells, int R) { return (R * run_seg_bytes(V, cells) + 127) / 128 * 128; }

// ---- async-copy ring helpers ------------------------------------------------------------------------
template <bool CA>
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    if constexpr (CA) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    else asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint4 lds16(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts16z(uint32_t addr) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(addr), "r"(0u) : "memory");
}
__device__ __forceinline__ int4 lds16i(uint32_t addr) {
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ int lds4i(uint32_t addr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ float4 lds16f(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
// lane base + 16 * off as ONE mad.wide (the base is made opaque so ptxas keeps it in a register pair
// instead of rebuilding it from blockIdx / threadIdx at every use)
__device__ __forceinline__ const uint4* tap_ptr(unsigned long long base, int off16) {
    unsigned long long a;
    asm("mad.wide.s32 %0, %1, 16, %2;" : "=l"(a) : "r"(off16), "l"(base));
    return reinterpret_cast<const uint4*>(a);
}
// the four taps of one load-list entry -> one ring stage (tap t at +512 t), 16 bytes per lane each
template <bool CA>
__device__ __forceinline__ void run_copy4(uint32_t stage, unsigned long long base, const int4& o) {
    cp_async16<CA>(stage, tap_ptr(base, o.x));
    cp_async16<CA>(stage + 512, tap_ptr(base, o.y));
    cp_async16<CA>(stage + 1024, tap_ptr(base, o.z));
    cp_async16<CA>(stage + 1536, tap_ptr(base, o.w));
}

// a load-list entry -> one ring stage: four taps, or (HALF entries: z = -2) the two taps of the new column
template <bool CA, bool HALF>
__device__ __forceinline__ void run_copy_entry(uint32_t stage, unsigned long long base, const int4& o) {
    if (HALF && o.z < -1) {
        cp_async16<CA>(stage, tap_ptr(base, o.x));
        cp_async16<CA>(stage + 512, tap_ptr(base, o.y));
    } else {
        run_copy4<CA>(stage, base, o);
    }
}

// ---- phase A of the run kernels: one warp builds one row segment's tables, directly in walking order -------------
// lane = (view of this pass, cell); a ballot gives every reload its place in the load list and every view that sees
// the segment its place in the view list (lane order = views ascending, cells ascending).  sH: the V homographies,
// rows padded to 4 floats.  MARK_INVALID (backward): out-of-map taps carry offset -1 instead of a stand-in address.
// HALF: a reload whose block is the previous one shifted by exactly one texel in x (both blocks fully inside the map)
// becomes a HALF entry: only the new column's two taps are listed ({top, bottom, -2, 0}), the walk keeps the other
// column in registers; per view, ml[2V+3+v] = half bits | (moved-to-the-east bits << 16).
// BOX: the load list holds texel coordinates instead of tap offsets ({x0 | y0 << 16, view, -, -}; end of list: view -1),
// for rings that are filled by one tensor-map [2 x 2] box copy per reload (ipm_boxrun.cuh).
// Border blocks (forward walk over tap offsets only, i.e. neither MARK_INVALID nor BOX): bit 16 of a view's mask word -- the
// reload bit of cell 0, which always equals its seen bit -- says instead "some reload of this view has an out-of-map tap", and
// ml[3V+3+2v] / ml[3V+4+2v] hold, per tap (NW | NE << 16, SW | SE << 16), the cells whose reload must zero that tap.
template <int CELLS, bool WANT_ALL_SEEN, bool MARK_INVALID, bool HALF = false, bool BOX = false>
__device__ __forceinline__ void run_build_tables(const FwdParams& p, int V, int i, int j0, int lane, int fsv16, const float* sH,
                                                 float4* wts, int4* loads, int* ml) {
    constexpr int GPW = 32 / CELLS;
    constexpr unsigned CMASK = (1u << CELLS) - 1u;
    const int gl = lane / CELLS, c = lane - gl * CELLS;
    const unsigned lt = (1u << lane) - 1u;
    int nloads = 0, nseen = 0;  // warp-uniform running totals
    unsigned all_seen = CMASK;  // cells every view sees
    for (int v0 = 0; v0 < V; v0 += GPW) {
        const int v = v0 + gl;
        const bool active = gl < GPW && v < V;
        const int j = j0 + c;
        CellTap t;
        t.flags = 0; t.x0 = t.y0 = -2; t.off16 = 0; t.nw = t.ne = t.sw = t.se = 0.0f;
        if (active && i < p.Hb && j < p.Wb) {
            float H[9], ix, iy;
            const float4* hv = reinterpret_cast<const float4*>(sH + 12 * v);
            const float4 h0 = hv[0], h1 = hv[1], h2 = hv[2];
            H[0] = h0.x; H[1] = h0.y; H[2] = h0.z; H[3] = h1.x; H[4] = h1.y; H[5] = h1.z; H[6] = h2.x; H[7] = h2.y; H[8] = h2.z;
            cell_coord(H, __ldg(p.xs + j), __ldg(p.ys + i), p.sw, p.sh, (float)p.Wf, (float)p.Hf, ix, iy, p.kx, p.ky);
            t = make_tap(ix, iy, p.Wf, p.Hf, p.fsy16, p.fsx16);
        }
        const bool seen = t.flags != 0;
        const unsigned seen_b = __ballot_sync(0xffffffffu, seen);
        const int px0 = __shfl_up_sync(0xffffffffu, t.x0, 1), py0 = __shfl_up_sync(0xffffffffu, t.y0, 1);
        const bool prev_seen = lane > 0 && ((seen_b >> (lane - 1)) & 1u);
        const bool same = c > 0 && prev_seen && px0 == t.x0 && py0 == t.y0;
        const bool reload = seen && !same;  // the row enters a new 2x2 block here
        const unsigned reload_b = __ballot_sync(0xffffffffu, reload);
        bool half = false, east = false;
        unsigned half_b = 0, east_b = 0;
        if constexpr (HALF) {
            const int pfl = __shfl_up_sync(0xffffffffu, t.flags, 1);
            const int dxs = t.x0 - px0;
            half = reload && c > 0 && prev_seen && py0 == t.y0 && (dxs == 1 || dxs == -1) && t.flags == kTapMask && pfl == kTapMask;
            east = half && dxs == 1;
            half_b = __ballot_sync(0xffffffffu, half);
            east_b = __ballot_sync(0xffffffffu, east);
        }
        const int shift = gl * CELLS;
        unsigned seen_c = (seen_b >> shift) & CMASK, reload_c = (reload_b >> shift) & CMASK;
        unsigned bz0 = 0, bz1 = 0;
        if constexpr (!MARK_INVALID && !BOX) {
            const unsigned inv = (reload && !(t.flags & kNonFinite)) ? (~(unsigned)t.flags & (unsigned)kTapMask) : 0u;
            if (__any_sync(0xffffffffu, inv != 0)) {  // warp-uniform, rare: only segments on the rim of a view's footprint
                const unsigned i0 = (__ballot_sync(0xffffffffu, inv & 1u) >> shift) & CMASK, i1 = (__ballot_sync(0xffffffffu, inv & 2u) >> shift) & CMASK;
                const unsigned i2 = (__ballot_sync(0xffffffffu, inv & 4u) >> shift) & CMASK, i3 = (__ballot_sync(0xffffffffu, inv & 8u) >> shift) & CMASK;
                bz0 = i0 | (i1 << 16);
                bz1 = i2 | (i3 << 16);
            }
            reload_c = (reload_c & ~1u) | ((bz0 | bz1) ? 1u : 0u);  // bit 16 of the mask word: the view has border blocks
        }
        const bool lead_seen = active && c == 0 && seen_c != 0;
        const unsigned lead_b = __ballot_sync(0xffffffffu, lead_seen);
        if (active) {
            const int vo = v * fsv16;
            const bool nf = (t.flags & kNonFinite) != 0;
            const float qnan = __int_as_float(0x7fc00000);
            const int tm = t.flags & kTapMask;
            const int first = tm ? (__ffs(tm) - 1) : 0;
            // an in-map tap of this block: where the out-of-map ones (weight 0) are pointed
            const int safe = (nf || !tm) ? vo : vo + t.off16 + ((first & 1) ? p.fsx16 : 0) + ((first & 2) ? p.fsy16 : 0);
            const float w[4] = {t.nw, t.ne, t.sw, t.se};
            int off[4];
            float ww[4];
#pragma unroll
            for (int tap = 0; tap < 4; ++tap) {
                const bool ok = (tm >> tap) & 1;
                off[tap] = ok ? vo + t.off16 + ((tap & 1) ? p.fsx16 : 0) + ((tap & 2) ? p.fsy16 : 0) : (MARK_INVALID ? -1 : safe);
                ww[tap] = nf ? qnan : (ok ? w[tap] : 0.0f);
            }
            wts[v * CELLS + c] = make_float4(ww[0], ww[1], ww[2], ww[3]);
            if (reload) {
                int4 entry = make_int4(off[0], off[1], off[2], off[3]);
                if (HALF && half) entry = east ? make_int4(off[1], off[3], -2, 0) : make_int4(off[0], off[2], -2, 0);  // the new column
                if (BOX) entry = make_int4((t.x0 & 0xffff) | (t.y0 << 16), v, 0, 0);
                loads[nloads + __popc(reload_b & lt)] = entry;
            }
            if (c == 0) {
                ml[v] = (int)(seen_c | (reload_c << 16));
                if (!MARK_INVALID && !BOX) { ml[3 * V + 3 + 2 * v] = (int)bz0; ml[3 * V + 4 + 2 * v] = (int)bz1; }
                if (HALF) ml[2 * V + 3 + v] = (int)(((half_b >> shift) & CMASK) | (((east_b >> shift) & CMASK) << 16));
            }
            if (lead_seen) ml[V + nseen + __popc(lead_b & lt)] = v;
        }
        nloads += __popc(reload_b);
        nseen += __popc(lead_b);
        if (WANT_ALL_SEEN) {
#pragma unroll
            for (int gq = 0; gq < GPW; ++gq)
                if (v0 + gq < V) all_seen &= (seen_b >> (gq * CELLS)) & CMASK;
        }
    }
    if (lane == 0) { ml[2 * V] = nseen; ml[2 * V + 1] = nloads; ml[2 * V + 2] = (int)all_seen; }
    if (lane < 8) loads[nloads + lane] = make_int4(-1, -1, 0, 0);  // end of list
}

// ---- TMA form of the ring (template flag TMA): one elected lane hands the copy engine four 512-byte bulk copies per
// load-list entry (cp.async.bulk global -> shared, completion counted on the stage's mbarrier in bytes); the warp
// waits on the barrier's phase parity instead of a cp.async group.  No LSU / L1 tag work for the loads at all.
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy512(uint32_t dst, const void* src, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], 512, [%2];" ::"r"(dst), "l"(src), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (int spin = 0; !done; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (spin > (1 << 24)) __trap();  // a lost copy must not hang the device
    }
}

// what a table cache is compared with / filled from: K (9 V) and Rt34 (12 V) of frame b, then the ends of the two BEV axes
__device__ __forceinline__ const float* plan_source(const FwdParams& p, int V, int b, int q) {
    if (q < 9 * V) return p.K + (size_t)b * 9 * V + q;
    if (q < 21 * V) return p.Rt + (size_t)b * 12 * V + (q - 9 * V);
    q -= 21 * V;
    return q == 0 ? p.xs : (q == 1 ? p.xs + (p.Wb - 1) : (q == 2 ? p.ys : p.ys + (p.Hb - 1)));
}

// A per-warp ring of DEPTH 2-KB stages in shared memory is filled by cp.async (.ca when CA, else .cg), DEPTH-1
// blocks in flight per warp; every lane reads back exactly the 16 bytes it copied, so no barrier is involved.
// One CTA walks `fpc` consecutive frames of its tile and re-uses the phase A tables for as long as the
// calibration of the next frame equals (bit for bit) the one the tables were built from -- static cameras,
// the normal case (wildtrack_loader.py:291-293 reads one calibration per camera).
// PROBE (timing aid, results are NOT the fusion): 1 = no copies are issued (instruction side alone), 2 = copies and
// unpack but no blend (memory side alone)
// KMODE: KM_ACC = sum / mean (fusion.py:18-21), KM_MAX = max over views, zeros of views that miss a cell included (fusion.py:22),
// KM_NONE = the per-view maps GeometryTransformer returns (geometry.py:162-163; what ConcatFusion reshapes): every (view, cell)
// result is stored as soon as it is blended, zeros where a view does not see a cell; no accumulators at all.
template <typename TIn, typename TOut, int CELLS, int NW, int KSPLIT, int MAXREG, int DEPTH, bool CA, int PROBE = 0, int KMODE = KM_ACC, bool TMA = false,
          bool HALF = false, bool PLAN = false>
__global__ void __maxnreg__(MAXREG) warp_fuse_run_kernel(const FwdParams p, int fpc) {
    static_assert(DEPTH >= 2 && DEPTH <= 8, "ring depth");
    static_assert(!TMA || (DEPTH & (DEPTH - 1)) == 0, "the TMA ring indexes its stages with a mask");
    static_assert(!HALF || (!TMA && PROBE == 0), "half reloads are built for the cp.async ring only");
    using VT = VecTraits<TIn>;
    constexpr int VE = VT::VE, P = VT::P;
    constexpr int R = NW / KSPLIT;  // row segments per CTA
    constexpr int NT = NW * 32;
    constexpr int ILP = (KMODE == KM_NONE) ? P : ((MAXREG <= 128 && P > 2 && CELLS >= 8) ? 2 : P);  // at 128 registers there is room for two chains in flight, not four
    static_assert(CELLS >= 2 && CELLS <= 16, "cells per segment");
    static_assert(NW % KSPLIT == 0 && (KSPLIT == 1 || KSPLIT == 2 || KSPLIT == 4), "warps per row segment");
    extern __shared__ __align__(128) unsigned char smem_run[];  // (its own symbol: the other kernels declare theirs 16-byte aligned)

    const int V = p.V;
    const int seg_bytes = run_seg_bytes(V, CELLS);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // tells the compiler it is warp-uniform
    const int ty = blockIdx.x / p.tiles_x, tx = blockIdx.x - ty * p.tiles_x;
    const int i0 = ty * R, j0 = tx * CELLS;
    const int b0 = blockIdx.z * fpc, b1 = min(p.B, b0 + fpc);

    auto seg_wts = [&](int r) { return reinterpret_cast<float4*>(smem_run + r * seg_bytes); };
    auto seg_loads = [&](int r) { return reinterpret_cast<int4*>(smem_run + r * seg_bytes + V * CELLS * 16); };
    auto seg_meta = [&](int r) { return reinterpret_cast<int*>(smem_run + r * seg_bytes + V * CELLS * 16 + (V * CELLS + 8) * 16); };
    // meta: [0, V) per-view mask (seen | reload << 16), [V, 2V) the views that see the segment, [2V] their
    // number, [2V+1] entries of the load list, [2V+2] mask of the cells EVERY view sees (max fusion)

    const int fsv16 = (int)(p.fs_v / VE);
    const int r = warp / KSPLIT, kk = warp - r * KSPLIT;
    const int i = i0 + r;
    const int chunks = (p.C + 32 * VE - 1) / (32 * VE);  // the last one may be partial: its spare lanes re-read the last
    const int last_vec = p.C / VE - 1;                    // valid 16 bytes of the texel and store nothing
    const float Vf = (float)V;
    const uint32_t s_wts = (uint32_t)__cvta_generic_to_shared(seg_wts(r));
    const uint32_t s_loads = (uint32_t)__cvta_generic_to_shared(seg_loads(r));
    const uint32_t s_meta = (uint32_t)__cvta_generic_to_shared(seg_meta(r));
    // this lane's 16 bytes of stage 0 / tap 0 in the warp's ring (stage = 2 KB, tap = 512 B)
    uint32_t ring = (uint32_t)__cvta_generic_to_shared(smem_run) + run_tables_bytes(V, CELLS, R) + warp * (DEPTH * 2048) + lane * 16;
    asm volatile("" : "+r"(ring));  // opaque: one register, not re-derived from %tid at every reload
    // TMA: one mbarrier per stage of this warp's ring, behind the homography table; entries issued / consumed so far
    uint32_t bars = (uint32_t)__cvta_generic_to_shared(smem_run) + run_tables_bytes(V, CELLS, R) + NW * (DEPTH * 2048) + V * 48 + warp * (DEPTH * 8);
    uint32_t n_issue = 0, n_cons = 0;
    if constexpr (TMA) {
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < DEPTH; ++s) mbar_init(bars + s * 8, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("" : "+r"(bars));
    }

    const int cpw = (chunks - kk + KSPLIT - 1) / KSPLIT;  // 512-byte chunks this warp walks per frame
    bool plan_filled = false;  // PLAN: this CTA wrote its segments' tables into an empty cache

    for (int b = b0; b < b1;) {
        if (b > b0) __syncthreads();  // every warp is done with the previous run's tables
        // ---- PLAN: the tables of this calibration may be waiting in the cache (static cameras: the `_grid_cache` the reference
        // declares, geometry.py:22).  Valid iff the key matches and K, Rt34 of this frame equal the cached ones bit for bit.
        bool use_plan = false, fill_plan = false;
        if constexpr (PLAN) {
            const PlanHeader* ph = reinterpret_cast<const PlanHeader*>(p.plan);
            // the copy of this segment's cached tables starts before the header has been checked: one memory round trip instead
            // of two (whatever an empty or foreign cache holds is overwritten by the build below)
            if (kk == 0) {
                const uint4* cached = reinterpret_cast<const uint4*>(static_cast<const unsigned char*>(p.plan) + kPlanHeaderBytes +
                                                                     ((size_t)(ty * R + r) * p.tiles_x + tx) * (size_t)seg_bytes);
                const uint32_t mine = (uint32_t)__cvta_generic_to_shared(smem_run + r * seg_bytes);
                for (int q = lane; q < seg_bytes / 16; q += 32) cp_async16<false>(mine + q * 16, cached + q);
                cp_async_commit();
            }
            const unsigned long long key = *reinterpret_cast<const volatile unsigned long long*>(&ph->key);
            bool differs = key != p.plan_key;
            if (!differs)
                for (int q = tid; q < 21 * V + 4; q += NT) differs |= __float_as_uint(__ldg(plan_source(p, V, b, q))) != ph->calib[q];
            use_plan = !__syncthreads_or(differs);
            // an empty cache is filled from frame 0, by the CTAs that walk it (all of them see it empty: the header is published
            // only after every one of them has finished)
            fill_plan = __syncthreads_and(key == 0ull) && b == 0;
            plan_filled |= fill_plan;
            if (kk == 0) {
                cp_async_wait<0>();
                __syncwarp();
            }
        }
        // ---- the V homographies of this frame, once per CTA (geometry.py:60-63): rows padded to 4 floats ------------
        float* sH = reinterpret_cast<float*>(smem_run + run_tables_bytes(V, CELLS, R) + NW * (DEPTH * 2048));
        if (!use_plan) {
            if (tid < V) {
                float H[9];
                homography(p.K + 9 * (b * V + tid), p.Rt + 12 * (b * V + tid), H);
#pragma unroll
                for (int q = 0; q < 3; ++q)
                    reinterpret_cast<float4*>(sH + 12 * tid)[q] = make_float4(H[3 * q], H[3 * q + 1], H[3 * q + 2], 0.0f);
            }
            __syncthreads();
        }
        // ---- phase A: the first warp of every row segment builds the segment's tables, directly in walking order --
        // lane = (view of this pass, cell); a ballot gives every reload its place in the load list and every view
        // that sees the segment its place in the view list (lane order = views ascending, cells ascending).
        if (kk == 0) {
            if constexpr (PLAN) {
                // this segment's tables in the cache: [header][segment (tile row ty * R + r) * tiles_x + tx] of seg_bytes each
                uint4* cached = reinterpret_cast<uint4*>(static_cast<unsigned char*>(p.plan) + kPlanHeaderBytes +
                                                         ((size_t)(ty * R + r) * p.tiles_x + tx) * (size_t)seg_bytes);
                uint4* mine = reinterpret_cast<uint4*>(smem_run + r * seg_bytes);
                if (!use_plan) {
                    run_build_tables<CELLS, KMODE == KM_MAX, false, HALF>(p, V, i, j0, lane, fsv16, sH, seg_wts(r), seg_loads(r), seg_meta(r));
                    if (fill_plan) {
                        __syncwarp();
                        for (int q = lane; q < seg_bytes / 16; q += 32) cached[q] = mine[q];
                    }
                }
            } else {
                run_build_tables<CELLS, KMODE == KM_MAX, false, HALF>(p, V, i, j0, lane, fsv16, sH, seg_wts(r), seg_loads(r), seg_meta(r));
            }
        }
        if (KSPLIT > 1) __syncthreads();
        else __syncwarp();

        // ---- the run of frames b .. e-1 shares these tables: same calibration, bit for bit -----------------------
        int e = b1;
        if (b + 1 < b1) {
            // all comparisons of the group in one round (the loads overlap); the usual answer is "no frame differs"
            bool differs = false;
            const int per = 21 * V, n = per * (b1 - b - 1);
            for (int z = tid; z < n; z += NT) {
                const int f = z / per, q = z - f * per;  // frame b + 1 + f against its predecessor
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
        const int n_items = (e - b) * cpw;  // this warp's (frame, chunk) items, frame-major
        b = e;

        // ---- phase B: per warp, one row segment x one 512-byte channel chunk of one frame at a time ----------
        if (i >= p.Hb || n_items <= 0) continue;
        const int nviews = __shfl_sync(0xffffffffu, lds4i(s_meta + 8 * V), 0);
        int fi_c = 0, k_c = kk;  // frame (within the run) and chunk of the item being blended
        // lane base of an item's chunk
        auto item_base = [&](int fi, int k) {
            return reinterpret_cast<unsigned long long>(
                reinterpret_cast<const uint4*>(reinterpret_cast<const TIn*>(p.feats) + (long long)(b_run + fi) * p.fs_b) + min(k * 32 + lane, last_vec));
        };
        // entries 0 .. DEPTH-2 of the load list start flying: one commit group per entry (past the end of the list
        // the entries carry x = -1: nothing is copied, the group is empty).  Called before the first item and, for
        // every later item, right after the walk of the one before it: the copies fly during that item's epilogue.
        // TMA: lane 0 arms the stage's barrier with the entry's 2048 bytes and issues the four bulk copies (its own `ring`
        // and lane base are the warp's: lane offset 0); the stage was read by every lane at an earlier reload
        auto issue_tma = [&](unsigned long long base, const int4& o) {
            __syncwarp();
            if (lane == 0) {
                const uint32_t st = n_issue & (DEPTH - 1), bar = bars + st * 8, dst = ring + st * 2048;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(bar, 2048);
                bulk_copy512(dst, tap_ptr(base, o.x), bar);
                bulk_copy512(dst + 512, tap_ptr(base, o.y), bar);
                bulk_copy512(dst + 1024, tap_ptr(base, o.z), bar);
                bulk_copy512(dst + 1536, tap_ptr(base, o.w), bar);
            }
            ++n_issue;
        };
        auto prime = [&](unsigned long long base) {
#pragma unroll
            for (int s = 0; s < DEPTH - 1; ++s) {
                const int4 o = lds16i(s_loads + s * 16);
                if constexpr (TMA) {
                    if (PROBE != 1 && o.x >= 0) issue_tma(base, o);
                } else {
                    if (PROBE != 1 && o.x >= 0) run_copy_entry<CA, HALF>(ring + s * 2048, base, o);
                    cp_async_commit();
                }
            }
        };
        unsigned long long lbase = item_base(0, kk);
        if (nviews > 0) prime(lbase);
        for (int it = 0; it < n_items; ++it) {
            float2 acc[CELLS][P];
#pragma unroll
            for (int c = 0; c < CELLS; ++c)
#pragma unroll
                for (int q = 0; q < P; ++q) acc[c][q] = (KMODE == KM_MAX) ? make_float2(-INFINITY, -INFINITY) : make_float2(0.0f, 0.0f);
            const bool lok = k_c * 32 + lane <= last_vec;  // this lane holds real channels of this item's chunk
            // KM_NONE: this item's 16 bytes of view 0, cell 0 in the per-view output
            TOut* ovb = reinterpret_cast<TOut*>(p.out) + (long long)(b_run + fi_c) * p.os_b + (long long)i * p.os_y + (long long)j0 * p.os_x +
                        (k_c * 32 + lane) * VE;

            if (nviews > 0) {
                float2 cur[4][P];
#pragma unroll
                for (int tap = 0; tap < 4; ++tap)
#pragma unroll
                    for (int q = 0; q < P; ++q) cur[tap][q] = make_float2(0.0f, 0.0f);
                asm volatile("" : "+l"(lbase));  // opaque: the lane base stays in its register pair
                uint32_t st_rd = ring, st_wr = ring + (DEPTH - 1) * 2048;
                uint32_t lp = s_loads + (DEPTH - 1) * 16;  // the entry the next reload hands to the copy engine
                for (int vi = 0; vi < nviews; ++vi) {
                    const int v = lds4i(s_meta + 4 * (V + vi));
                    const unsigned m = (unsigned)__shfl_sync(0xffffffffu, lds4i(s_meta + 4 * v), 0);  // warp-uniform
                    const uint32_t wv = s_wts + v * (CELLS * 16);
                    unsigned m2 = 0;  // HALF: which reloads only bring a new column, and on which side
                    if constexpr (HALF) m2 = (unsigned)__shfl_sync(0xffffffffu, lds4i(s_meta + 4 * (2 * V + 3 + v)), 0);
                    float4 wn = lds16f(wv);
                    // A cell the view does not see is never a reload; its blend runs on whatever `cur` holds
                    // and is simply not added (predicated), so the only branch per cell is "reload?".
                    bool rl = m & 1u;                    // cell 0 starts a block iff the view sees it
                    const bool bview = (m >> 16) & 1u;  // some block of this view hangs over the map's border (warp-uniform, rare)
                    // the reference pads with zeros (geometry.py:161 padding_mode='zeros'): before a border block is read back from
                    // the ring, every lane overwrites its 16 bytes of the out-of-map taps (stand-in texels) with zeros, so whatever
                    // the stand-in holds never reaches the blend (w * 0 = +0).  A lane reads back only bytes it wrote itself.
                    auto patch_border = [&](int c, uint32_t st) {
                        const int z0 = lds4i(s_meta + 4 * (3 * V + 3 + 2 * v)), z1 = lds4i(s_meta + 4 * (3 * V + 4 + 2 * v));
                        if ((z0 >> c) & 1) sts16z(st);
                        if ((z0 >> (16 + c)) & 1) sts16z(st + 512);
                        if ((z1 >> c) & 1) sts16z(st + 1024);
                        if ((z1 >> (16 + c)) & 1) sts16z(st + 1536);
                    };
                    // two copies of the walk: views with border blocks (rare) patch the ring before every read-back, everybody else
                    // runs the copy without the test
                    auto walk_cells = [&](auto border_tag) {
                    constexpr bool BORDER = decltype(border_tag)::value;
#pragma unroll
                    for (int c = 0; c < CELLS; ++c) {
                        const float4 w = wn;
                        if (c + 1 < CELLS) wn = lds16f(wv + (c + 1) * 16);  // one cell ahead of its use
                        const bool seen = (m >> c) & 1u;                       // warp-uniform
                        const bool rl_now = rl;
                        if (c + 1 < CELLS) rl = (m >> (17 + c)) & 1u;          // decided one cell early
                        if (rl_now) {                                          // the row leaves the block held in `cur`
                            if constexpr (TMA) {
                                const uint32_t stc = n_cons & (DEPTH - 1);
                                mbar_wait(bars + stc * 8, (n_cons / DEPTH) & 1u);  // the entry's 2048 bytes have landed
                                const uint32_t sr = ring + stc * 2048;
                                if constexpr (BORDER) patch_border(c, sr);
                                uint4 nxt[4];
                                nxt[0] = lds16(sr); nxt[1] = lds16(sr + 512);
                                nxt[2] = lds16(sr + 1024); nxt[3] = lds16(sr + 1536);
                                const int4 o = lds16i(lp);
#pragma unroll
                                for (int tap = 0; tap < 4; ++tap) VT::unpack(nxt[tap], cur[tap]);
                                ++n_cons;
                                // the entry DEPTH-1 ahead goes into the stage unpacked at the previous reload
                                if (PROBE != 1 && o.x >= 0) issue_tma(lbase, o);
                                lp += 16;
                            } else if (HALF && ((m2 >> c) & 1u)) {
                            // the block moved by one texel in x: keep the shared column, unpack the new one
                            cp_async_wait<DEPTH - 2>();
                            const uint4 n0 = lds16(st_rd), n1 = lds16(st_rd + 512);
                            const int4 o = lds16i(lp);
                            if ((m2 >> (16 + c)) & 1u) {  // to the east: old NE / SE become NW / SW
#pragma unroll
                                for (int q = 0; q < P; ++q) { cur[0][q] = cur[1][q]; cur[2][q] = cur[3][q]; }
                                VT::unpack(n0, cur[1]);
                                VT::unpack(n1, cur[3]);
                            } else {                       // to the west
#pragma unroll
                                for (int q = 0; q < P; ++q) { cur[1][q] = cur[0][q]; cur[3][q] = cur[2][q]; }
                                VT::unpack(n0, cur[0]);
                                VT::unpack(n1, cur[2]);
                            }
                            if (o.x >= 0) run_copy_entry<CA, HALF>(st_wr, lbase, o);
                            cp_async_commit();
                            lp += 16;
                            st_wr = st_rd;
                            st_rd = (st_rd == ring + (DEPTH - 1) * 2048) ? ring : st_rd + 2048;
                            } else {
                            cp_async_wait<DEPTH - 2>();  // the oldest entry has landed (a lane reads back its own bytes)
                            if constexpr (BORDER) patch_border(c, st_rd);
                            uint4 nxt[4];
                            nxt[0] = lds16(st_rd); nxt[1] = lds16(st_rd + 512);
                            nxt[2] = lds16(st_rd + 1024); nxt[3] = lds16(st_rd + 1536);
                            const int4 o = lds16i(lp);
#pragma unroll
                            for (int tap = 0; tap < 4; ++tap) VT::unpack(nxt[tap], cur[tap]);
                            // the entry DEPTH-1 ahead goes into the stage unpacked at the previous reload
                            if (PROBE != 1 && o.x >= 0) run_copy_entry<CA, HALF>(st_wr, lbase, o);
                            cp_async_commit();
                            lp += 16;
                            st_wr = st_rd;
                            st_rd = (st_rd == ring + (DEPTH - 1) * 2048) ? ring : st_rd + 2048;
                            }
                        }
                        if constexpr (PROBE == 2) {
                            if (rl_now) {
#pragma unroll
                                for (int q = 0; q < P; ++q) {
                                    acc[c][q].x = __uint_as_float(__float_as_uint(acc[c][q].x) ^ __float_as_uint(cur[0][q].x) ^ __float_as_uint(cur[1][q].y));
                                    acc[c][q].y = __uint_as_float(__float_as_uint(acc[c][q].y) ^ __float_as_uint(cur[2][q].x) ^ __float_as_uint(cur[3][q].y));
                                }
                            }
                        } else {
                            // out_v = fma(SE,se, fma(SW,sw, fma(NE,ne, NW*nw)))   ATen's interpolation order; ILP of
                            // the P independent chains are written interleaved so an instruction does not wait
                            // on the one right before it
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
                                if constexpr (KMODE == KM_NONE) {  // geometry.py:162: the view's map, zero where it does not see the cell
                                    if (lok && j0 + c < p.Wb) {
                                        float2 z[P];
#pragma unroll
                                        for (int q = 0; q < P; ++q) z[q] = seen ? sv[q] : make_float2(0.0f, 0.0f);
                                        store_pairs<TOut, P>(ovb + (long long)v * p.os_v + (long long)c * p.os_x, z);
                                    }
                                } else {
#pragma unroll
                                for (int q = 0; q < ILP; ++q)
                                    if (seen) {
                                        if constexpr (KMODE == KM_MAX) {  // fusion.py:22, NaN propagates like torch.max
                                            float2& mx = acc[c][q0 + q];
                                            mx.x = max_nan(mx.x, sv[q].x);
                                            mx.y = max_nan(mx.y, sv[q].y);
                                        } else {
                                            acc[c][q0 + q] = __fadd2_rn(acc[c][q0 + q], sv[q]);  // fusion.py:18-21, views ascending per cell
                                        }
                                    }
                                }
                            }
                        }
                    }
                    };
                    if (!bview) walk_cells(std::false_type{});   // (this order keeps the common copy first in the code)
                    else walk_cells(std::true_type{});
                }
                if constexpr (!TMA) cp_async_wait<0>();  // (only empty groups are left) the ring restarts with the next item
            }
            // the item after this one: its first blocks fly while this item is divided, packed and stored
            const int k_this = k_c, fi_this = fi_c;
            k_c += KSPLIT;
            if (k_c >= chunks) { k_c = kk; ++fi_c; }
            if (nviews > 0 && it + 1 < n_items) {
                lbase = item_base(fi_c, k_c);
                prime(lbase);
            }

            if constexpr (KMODE == KM_NONE) {
                // views that do not see the segment at all: their part of the per-view output is zero
                float2 z[P];
#pragma unroll
                for (int q = 0; q < P; ++q) z[q] = make_float2(0.0f, 0.0f);
                TOut* ob0 = reinterpret_cast<TOut*>(p.out) + (long long)(b_run + fi_this) * p.os_b + (long long)i * p.os_y + (long long)j0 * p.os_x +
                            (k_this * 32 + lane) * VE;
                for (int v = 0; v < V; ++v) {
                    if ((lds4i(s_meta + 4 * v) & 0xffff) != 0) continue;
#pragma unroll
                    for (int c = 0; c < CELLS; ++c)
                        if (lok && j0 + c < p.Wb) store_pairs<TOut, P>(ob0 + (long long)v * p.os_v + (long long)c * p.os_x, z);
                }
                continue;
            }
            // ---- epilogue: mean division (IEEE quotient) and one 16-byte store per cell ---------------------------
            if constexpr (KMODE == KM_MAX) {
                // a view that misses the cell contributes its zero padding to the maximum (geometry.py:94 + fusion.py:22)
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
                // Markstein's 3-op division is exact for finite sums (ipm_fused.cuh div_exact_vec); +-Inf would
                // turn into NaN.  One test per chunk: the packed sum of all CELLS x 8 accumulators is finite
                // only if every one of them is (Inf - Inf = NaN; a finite overflow merely takes the slow path).
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
                    const float r = p.rcpV;
#pragma unroll
                    for (int c = 0; c < CELLS; ++c)
#pragma unroll
                        for (int q = 0; q < P; ++q) {
                            const float2 qq = __fmul2_rn(acc[c][q], make_float2(r, r));
                            const float2 rem = __ffma2_rn(qq, make_float2(-Vf, -Vf), acc[c][q]);
                            acc[c][q] = __ffma2_rn(rem, make_float2(r, r), qq);
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
            if constexpr (KMODE == KM_RED) {
                // View sharding: this rank's partial sum goes into the slab of the rank that owns BEV row i (its own memory
                // or a peer's over NVLink).  The warp parks its 8 cells x one chunk of fp32 sums in shared memory and hands
                // them to the copy engine as bulk reductions (cp.reduce.async.bulk ... add.f32): kilobyte packets over the
                // link instead of one 16-byte atomic per lane, and the kernel does not wait for them.
                constexpr int CB = 32 * VE * 4;  // bytes of one cell's chunk of fp32 sums (512 for fp32, 1024 for bf16 features)
                unsigned char* stg = smem_run + ((run_tables_bytes(V, CELLS, R) + NW * (DEPTH * 2048) + V * 48 + NW * DEPTH * 8 + 127) / 128) * 128 +
                                     warp * (CELLS * CB);
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the previous item's sums have left
                __syncwarp();
#pragma unroll
                for (int c = 0; c < CELLS; ++c)
#pragma unroll
                    for (int q = 0; q < P; q += 2)
                        *reinterpret_cast<float4*>(stg + c * CB + lane * (VE * 4) + q * 8) = make_float4(acc[c][q].x, acc[c][q].y, acc[c][q + 1].x, acc[c][q + 1].y);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    const int owner = i / p.slab_rows;
                    float* oc = reinterpret_cast<float*>(p.slab[owner]) + (long long)(b_run + fi_this) * p.os_b +
                                (long long)(i - owner * p.slab_rows) * p.os_y + (long long)j0 * p.os_x + k_this * 32 * VE;
                    const int ch = min(32 * VE, p.C - k_this * 32 * VE);       // channels of this chunk that exist
                    const int ncell = min(CELLS, p.Wb - j0);
                    const uint32_t src = (uint32_t)__cvta_generic_to_shared(stg);
                    const bool run = ch * 4 == CB && p.os_x == 32 * VE;          // one chunk per texel: the cells are contiguous
                    const int nop = run ? 1 : ncell, sz = run ? ncell * CB : ch * 4;
                    if (p.slab_put) {   // this rank's own receive buffer at the owner: plain stores, summed by the owner afterwards
                        for (int c = 0; c < nop; ++c)
                            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(oc + (long long)c * p.os_x), "r"(src + c * CB), "r"(sz) : "memory");
                    } else {
                        for (int c = 0; c < nop; ++c)
                            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(oc + (long long)c * p.os_x), "r"(src + c * CB), "r"(sz) : "memory");
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                continue;
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
    if constexpr (KMODE == KM_RED) {
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // every bulk reduction of this warp has been performed
    }
    if constexpr (PLAN) {
        // the last of the CTAs that filled an empty cache publishes its header: calibration first, key last
        if (plan_filled) {
            __syncthreads();
            if (tid == 0) {
                PlanHeader* ph = reinterpret_cast<PlanHeader*>(p.plan);
                __threadfence();
                if (atomicAdd(&ph->done, 1u) == gridDim.x - 1) {
                    for (int q = 0; q < 21 * V + 4; ++q) ph->calib[q] = __float_as_uint(__ldg(plan_source(p, V, 0, q)));
                    ph->done = 0;
                    __threadfence();
                    *reinterpret_cast<volatile unsigned long long*>(&ph->key) = p.plan_key;
                }
            }
        }
    }
}

}  // namespace bevipm
