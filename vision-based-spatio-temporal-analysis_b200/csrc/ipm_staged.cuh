// ipm_staged.cuh -- "staged" form of the fused warp-and-fuse kernel: source tiles are staged in shared memory by TMA.
//
// Same arithmetic as ipm_run.cuh (geometry.py:120-162 + fusion.py:17-22 of the reference, the op chain of
// ipm_geometry.cuh), other data path.  The run kernel fetches every 2x2 texel block a row segment enters with four
// per-lane cp.async requests through L1 (2.27 texel fetches per cell-view on BASELINE config 1, 8.3 GB of L2->L1
// traffic per launch, 17 issue slots of ring/address work per reload).  Here a CTA owns a tile of NW BEV rows x 8 cells
// and, per view and 512-byte channel chunk, ONE elected thread asks the TMA unit for the part of the source map the
// tile's taps fall into; all warps of the CTA then blend from shared memory with LDS.128:
//
//   phase A (once per tile and run of frames with equal calibration): every (row, view, cell) is projected once
//            (cell_coord + make_tap); per (BEV row, view) the taps' texel rows and, per texel row, the span of x they
//            touch are collected with shared-memory atomics.  Per view the rows are then grouped: the whole tile if the
//            merged spans fit `cap` bytes, else its halves, quarters, ... single BEV rows; a row whose own spans do
//            not fit (the BEV is coarser than the source map there) is staged as one [2 x 2] box per 2x2 block it
//            enters.  A group is one STAGE: a list of row copies, tensor-map boxes of [256 channels x BW texels x 1
//            row], BW quantised to the widths the launcher encoded maps for.  Out-of-map parts of a box are
//            zero-filled by the TMA unit: exactly the reference's zero padding (grid_sample padding_mode='zeros': a
//            tap outside the map contributes 0 * w), so there is no tap mask, no stand-in address and no special
//            case for non-finite features.
//            The stages of one (frame, channel chunk) item get STATIC places in a ring of shared memory (sequential
//            placement, wrapping to offset 0) and, each, the distance back to the last stage that used any of its
//            bytes before: the whole copy schedule is periodic and known before the first copy is issued.
//   phase B: the stage list is walked once per item.  Stage g of the run uses the full (transaction bytes) / empty (one
//            arrival per warp) mbarrier pair g % 64, phase parity (g / 64) & 1.  The warps take turns at issuing: before
//            stage g is walked, warp g % NW arms every stage whose predecessor in the ring was stage g - LAG or
//            older (it waits for that predecessor's empty barrier first), so copies run as far ahead as the ring
//            holds.  A warp owns one BEV row segment of 8 cells, keeps the 8 cells' accumulators in registers and
//            walks a stage like the run kernel walks a view: on a reload bit it reads the 2x2 block from the stage's
//            place (4 x LDS.128), unpacks it once, and blends it for as many cells as stay in the block.  Per cell
//            the views are still added in ascending order: the reference's accumulation order.
//            Every warp waits for and releases every stage, also those that hold nothing for its row: that keeps all
//            warps within one ring revolution of each other, which is what makes parity waits unambiguous.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ipm_run.cuh"

namespace bevipm {

constexpr int kStCells = 8;      // cells per row segment (one warp)
constexpr int kStRowsG = 32;     // texel rows a group stage may span
constexpr int kStRowsR = 12;     // texel rows one BEV row's spans may cover (else the row is staged block by block)
constexpr int kStNumW = 13;      // span widths with a tensor map of their own
constexpr int kStBlockMap = 13;  // the [2 x 2] block map
constexpr int kStNumMaps = 14;
constexpr int kStSched = 256;     // entries of the copy schedule (one period)
constexpr int kStBack = 64;       // how far back the schedule looks for a stage's predecessor in the ring; also the
                                  // number of mbarrier pairs: stage g uses pair g % 64 with phase parity (g / 64) & 1

// span width (texels, >= 1) -> index of the narrowest map that covers it; -1: wider than any map
__host__ __device__ __forceinline__ int st_width_index(int w) {
    if (w <= 16) return (w - 1) >> 1;
    if (w <= 20) return 8;
    if (w <= 24) return 9;
    if (w <= 32) return 10;
    if (w <= 48) return 11;
    if (w <= 64) return 12;
    return -1;
}
__host__ __device__ __forceinline__ int st_width(int idx) {
    return idx < 8 ? 2 * (idx + 1) : (idx == 8 ? 20 : (idx == 9 ? 24 : (idx == 10 ? 32 : (idx == 11 ? 48 : 64))));
}

struct alignas(64) StagedMaps {
    CUtensorMap m[kStNumMaps];
};

// shared-memory layout of one CTA: persistent tables, then the ring (phase A's scratch aliases the ring)
struct StagedSmem {
    int wts, ent, wst, sdesc, sched, ops, bars, misc, sH, ring;    // persistent
    int xy, mask, yr, rsp, gtab, ginfo, vst, vcnt, scratch_end;    // scratch, inside the ring
    int nst_max, max_ops, gt;
    __host__ __device__ StagedSmem(int V, int R) {
        auto up = [](int x, int a) { return (x + a - 1) / a * a; };
        nst_max = V * R;                                 // a view is at most R stages (one per BEV row)
        gt = kStRowsG > R * 16 ? kStRowsG : R * 16;      // planned texel rows per view: R/2 pairs x 32 rows at most
        max_ops = V * gt;
        int o = 0;
        wts = o; o += R * V * kStCells * 16;             // float4 (nw, ne, sw, se) per (row, view, cell)
        ent = o; o += R * V * kStCells * 8;              // int2 (byte offset of the NW tap, of the SW tap) inside the stage
        wst = o; o += up(R * nst_max * 4, 16);           // per (row, stage): seen | reload << 8 | view << 24; 0 = not this row's
        sdesc = o; o += nst_max * 16;                    // int4 {bytes, -, first op | ops << 16, view | row mask << 8}
        sched = o; o += kStSched * 8;                    // int2 {ring offset / 128 | stages back to the predecessor << 16, bytes} per stage of one period
        ops = o; o += max_ops * 8;                       // int2 {x | y << 16, stage offset / 16 | map << 16}
        bars = o; o += 2 * kStBack * 8;                  // full[64], empty[64]
        misc = o; o += 128;                              // [0] stages, [1 + r] cells of row r every view sees
        sH = o; o += V * 48;                             // homographies, rows padded to 4 floats
        ring = up(o, 128);
        o = ring;
        xy = o; o += R * V * kStCells * 4;               // x0 | y0 << 16
        mask = o; o += up(R * V * 4, 16);                // seen | reload << 16 per (row, view)
        yr = o; o += up(2 * R * V * 4, 16);              // min / max texel row per (row, view)
        rsp = o; o += R * V * kStRowsR * 8;              // x spans per (row, view, texel row): lo, hi
        gtab = o; o += V * gt * 4;                       // planned texel rows: stage offset in texels | map << 12 | x lo << 16
        ginfo = o; o += R * V * 16;                      // per (row, view): kind, first gtab row of its group, group's first texel row, real reloads
        vst = o; o += V * R * 32;                        // per (view, stage of the view): row mask, bytes, ops, kind, gtab row, first texel row, texel rows, row
        vcnt = o; o += up(V * 8, 16);                    // per view: stages, ops
        scratch_end = o;
    }
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const void* map, int c0, int c1, int c2, int c3, int c4, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(bar) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ int2 lds8i(uint32_t addr) {
    int2 v;
    asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}

enum { ST_SPANS = 1, ST_BLOCKS = 2 };

// ---- phase A ------------------------------------------------------------------------------------------------------
// Builds, for the tile at (i0, j0) of frame b: blend weights, the stage list with its row copies, ring places and
// predecessor distances, every (row, view, cell)'s tap offsets inside its stage and every warp's per-stage word.
template <int NW, bool WANT_ALL_SEEN>
__device__ __forceinline__ void staged_build(const FwdParams& p, const StagedSmem& L, unsigned char* sm, int cap, int i0, int j0, int b) {
    constexpr int CELLS = kStCells, GPW = 32 / CELLS, NT = NW * 32;
    constexpr unsigned CMASK = (1u << CELLS) - 1u;
    static_assert(NW == 4 || NW == 8 || NW == 16, "rows per tile");
    static_assert(kStBack == 64, "barrier pair = stage & 63, parity = stage >> 6");
    const int V = p.V, GT = L.gt;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gl = lane / CELLS, c = lane - gl * CELLS;
    const unsigned lt = (1u << lane) - 1u;
    float4* wts = reinterpret_cast<float4*>(sm + L.wts);
    int2* ent = reinterpret_cast<int2*>(sm + L.ent);
    unsigned* wst = reinterpret_cast<unsigned*>(sm + L.wst);
    int4* sdesc = reinterpret_cast<int4*>(sm + L.sdesc);
    int2* ops = reinterpret_cast<int2*>(sm + L.ops);
    int* misc = reinterpret_cast<int*>(sm + L.misc);
    float* sH = reinterpret_cast<float*>(sm + L.sH);
    int* xy = reinterpret_cast<int*>(sm + L.xy);
    unsigned* mask = reinterpret_cast<unsigned*>(sm + L.mask);
    int* yr = reinterpret_cast<int*>(sm + L.yr);        // [r * V + v] lo, [NW * V + ...] hi
    int* rsp = reinterpret_cast<int*>(sm + L.rsp);      // [((r * V + v) * 12 + row) * 2] lo, + 1 hi
    unsigned* gtab = reinterpret_cast<unsigned*>(sm + L.gtab);
    int4* ginfo = reinterpret_cast<int4*>(sm + L.ginfo);
    int* vst = reinterpret_cast<int*>(sm + L.vst);      // [(v * NW + k) * 8 + ...]
    int* vcnt = reinterpret_cast<int*>(sm + L.vcnt);    // [v * 2 + {stages, ops}]
    constexpr int IMAX = 0x7fffffff, IMIN = (int)0x80000000;

    // ---- P0: homographies (geometry.py:60-63), scratch initialisation -------------------------------------------
    if (tid < V) {
        float H[9];
        homography(p.K + 9 * (b * V + tid), p.Rt + 12 * (b * V + tid), H);
#pragma unroll
        for (int q = 0; q < 3; ++q) reinterpret_cast<float4*>(sH + 12 * tid)[q] = make_float4(H[3 * q], H[3 * q + 1], H[3 * q + 2], 0.0f);
    }
    for (int z = tid; z < NW * V * kStRowsR; z += NT) { rsp[2 * z] = IMAX; rsp[2 * z + 1] = IMIN; }
    __syncthreads();

    // ---- P1: project this warp's row; weights, masks, the row's texel-row range and spans per view -----------------
    const int r = warp, i = i0 + r;
    unsigned all_seen = CMASK;
    for (int v0 = 0; v0 < V; v0 += GPW) {
        const int v = v0 + gl;
        const bool active = v < V;
        const int j = j0 + c;
        CellTap t;
        t.flags = 0; t.x0 = t.y0 = -2; t.off16 = 0; t.nw = t.ne = t.sw = t.se = 0.0f;
        if (active && i < p.Hb && j < p.Wb) {
            float H[9], ix, iy;
            const float4* hv = reinterpret_cast<const float4*>(sH + 12 * v);
            const float4 h0 = hv[0], h1 = hv[1], h2 = hv[2];
            H[0] = h0.x; H[1] = h0.y; H[2] = h0.z; H[3] = h1.x; H[4] = h1.y; H[5] = h1.z; H[6] = h2.x; H[7] = h2.y; H[8] = h2.z;
            cell_coord(H, __ldg(p.xs + j), __ldg(p.ys + i), p.sw, p.sh, (float)p.Wf, (float)p.Hf, ix, iy, p.kx, p.ky);
            t = make_tap(ix, iy, p.Wf, p.Hf);
        }
        const bool seen = t.flags != 0;                  // some tap inside the map, or a non-finite position (NaN result)
        const bool real = (t.flags & kTapMask) != 0;     // has texels to stage
        const unsigned seen_b = __ballot_sync(0xffffffffu, seen);
        const int px0 = __shfl_up_sync(0xffffffffu, t.x0, 1), py0 = __shfl_up_sync(0xffffffffu, t.y0, 1);
        const bool prev_seen = lane > 0 && ((seen_b >> (lane - 1)) & 1u);
        const bool same = c > 0 && prev_seen && px0 == t.x0 && py0 == t.y0;
        const bool reload = seen && !same;               // the row enters a new 2x2 block here
        const unsigned reload_b = __ballot_sync(0xffffffffu, reload);
        int ymin = real ? t.y0 : IMAX, ymax = real ? t.y0 : IMIN;
#pragma unroll
        for (int o = 1; o < CELLS; o *= 2) {
            ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
            ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
        }
        const int shift = gl * CELLS;
        if (active) {
            const bool nf = (t.flags & kNonFinite) != 0;
            const float qnan = __int_as_float(0x7fc00000);
            const int e = (r * V + v) * CELLS + c;
            wts[e] = nf ? make_float4(qnan, qnan, qnan, qnan) : make_float4(t.nw, t.ne, t.sw, t.se);
            xy[e] = (t.x0 & 0xffff) | (t.y0 << 16);
            if (c == 0) {
                mask[r * V + v] = ((seen_b >> shift) & CMASK) | (((reload_b >> shift) & CMASK) << 16);
                yr[r * V + v] = ymin;
                yr[NW * V + r * V + v] = ymax;
            }
            if (real && ymax - ymin + 2 <= kStRowsR) {
                int* sp = rsp + ((r * V + v) * kStRowsR + (t.y0 - ymin)) * 2;
                atomicMin(sp, t.x0); atomicMax(sp + 1, t.x0 + 1);
                atomicMin(sp + 2, t.x0); atomicMax(sp + 3, t.x0 + 1);
            }
        }
        if (WANT_ALL_SEEN) {
#pragma unroll
            for (int gq = 0; gq < GPW; ++gq)
                if (v0 + gq < V) all_seen &= (seen_b >> (gq * CELLS)) & CMASK;
        }
    }
    if (lane == 0) misc[1 + r] = (int)all_seen;
    __syncthreads();

    // ---- P3: group the rows of every view into stages.  Warp w owns views w, w + NW, ...; lane = texel row -------------
    for (int v = warp; v < V; v += NW) {
        const bool rl = lane < NW;
        const int my_lo = rl ? yr[lane * V + v] : IMAX, my_hi = rl ? yr[NW * V + lane * V + v] : IMIN;
        const unsigned rowseen = __ballot_sync(0xffffffffu, rl && (mask[(rl ? lane : 0) * V + v] & 0xffffu) != 0);
        const unsigned realrows = __ballot_sync(0xffffffffu, rl && my_lo <= my_hi);
        const unsigned spanable = __ballot_sync(0xffffffffu, rl && my_lo <= my_hi && my_hi - my_lo + 2 <= kStRowsR);
        int nst_v = 0, nops_v = 0, gbase = 0;
        for (int r0 = 0; r0 < NW;) {
            int len = NW;
            while (len > 1 && (r0 & (len - 1))) len >>= 1;
            for (;;) {
                const unsigned gm = ((len >= 32 ? 0u : (1u << len)) - 1u) << r0;
                const unsigned gseen = rowseen & gm, greal = realrows & gm;
                if (!gseen) break;                                         // no row of the group samples this view: no stage
                if (len > 1 && (greal & ~spanable)) { len >>= 1; continue; }  // a row too tall for spans: split
                int kind = ST_SPANS, nops = 0, bytes = 0, gmin = 0, nrows = 0;
                if (greal && (len > 1 || (spanable >> r0) & 1u)) {
                    // merged spans of the group's rows: texel row `lane` of the group
                    int a = (greal >> lane) & 1u ? my_lo : IMAX, z = (greal >> lane) & 1u ? my_hi : IMIN;
#pragma unroll
                    for (int o = 1; o < 32; o *= 2) {
                        a = min(a, __shfl_xor_sync(0xffffffffu, a, o));
                        z = max(z, __shfl_xor_sync(0xffffffffu, z, o));
                    }
                    gmin = a;
                    nrows = z - a + 2;
                    bool ok = nrows <= kStRowsG;
                    int lo = IMAX, hi = IMIN;
                    if (ok) {
                        for (unsigned rem = greal; rem;) {
                            const int rr = __ffs(rem) - 1;
                            rem &= rem - 1;
                            const int idx = lane - (__shfl_sync(0xffffffffu, my_lo, rr) - gmin);
                            if (idx >= 0 && idx < kStRowsR) {
                                const int* sp = rsp + ((rr * V + v) * kStRowsR + idx) * 2;
                                lo = min(lo, sp[0]); hi = max(hi, sp[1]);
                            }
                        }
                    }
                    const int w = (lo <= hi) ? hi - lo + 1 : 0;
                    const int widx = w ? st_width_index(w) : -1;
                    const bool bad = w && widx < 0;
                    const int wq = (w && widx >= 0) ? st_width(widx) : 0;
                    int incl = wq;
#pragma unroll
                    for (int o = 1; o < 32; o *= 2) {
                        const int u = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += u;
                    }
                    const int total = __shfl_sync(0xffffffffu, incl, 31);
                    ok = ok && !__any_sync(0xffffffffu, bad) && total * 512 <= cap;
                    if (!ok) {
                        if (len > 1) { len >>= 1; continue; }
                        kind = ST_BLOCKS;
                    } else {
                        if (lane < nrows) gtab[v * GT + gbase + lane] = (unsigned)(incl - wq) | ((unsigned)(w ? widx : 15) << 12) | ((unsigned)(lo & 0xffff) << 16);
                        nops = __popc(__ballot_sync(0xffffffffu, w > 0));
                        bytes = total * 512;
                    }
                } else if (greal) {
                    kind = ST_BLOCKS;                                      // a single row too tall for spans
                }
                // (a group whose rows only hold non-finite sample positions: a stage without copies, kind SPANS, 0 bytes)
                unsigned rrl = 0;
                if (kind == ST_BLOCKS) {
                    // reloads of cells that have texels (a non-finite sample position has none): one [2 x 2] box each
                    const unsigned m = mask[r0 * V + v];
                    const int q = xy[(r0 * V + v) * CELLS + (lane & (CELLS - 1))];
                    rrl = __ballot_sync(0xffffffffu, lane < CELLS && ((m >> (16 + lane)) & 1u) && (int)(short)(q & 0xffff) != -2);
                    nops = __popc(rrl);
                    bytes = nops * 2048;
                    nrows = 0;
                }
                if (lane < NW && ((gseen >> lane) & 1u)) ginfo[lane * V + v] = make_int4(kind, gbase, gmin, (int)rrl);
                if (lane == 0) {
                    int* st = vst + (v * NW + nst_v) * 8;
                    st[0] = (int)gseen; st[1] = bytes; st[2] = nops; st[3] = kind; st[4] = gbase; st[5] = gmin; st[6] = nrows; st[7] = r0;
                }
                gbase += nrows;
                ++nst_v;
                nops_v += nops;
                break;
            }
            r0 += len;
        }
        if (lane == 0) { vcnt[2 * v] = nst_v; vcnt[2 * v + 1] = nops_v; }
    }
    __syncthreads();

    // ---- P4: emit the stage list and the row copies (views ascending: the accumulation order) --------------------
    int nst;
    {
        int my_st = 0, my_op = 0;  // lane = view: exclusive prefix sums of stages and ops
        if (lane < V) { my_st = vcnt[2 * lane]; my_op = vcnt[2 * lane + 1]; }
        int inc_st = my_st, inc_op = my_op;
#pragma unroll
        for (int o = 1; o < 32; o *= 2) {
            const int a = __shfl_up_sync(0xffffffffu, inc_st, o), bq = __shfl_up_sync(0xffffffffu, inc_op, o);
            if (lane >= o) { inc_st += a; inc_op += bq; }
        }
        nst = __shfl_sync(0xffffffffu, inc_st, 31);
        if (tid == 0) misc[0] = nst;
        for (int v = warp; v < V; v += NW) {
            int s = __shfl_sync(0xffffffffu, inc_st - my_st, v), o = __shfl_sync(0xffffffffu, inc_op - my_op, v);
            const int nv = __shfl_sync(0xffffffffu, my_st, v);
            for (int k = 0; k < nv; ++k) {
                const int* st = vst + (v * NW + k) * 8;
                const int nops = st[2], kind = st[3];
                if (kind == ST_SPANS) {
                    unsigned e = 15u << 12;
                    if (lane < st[6]) e = gtab[v * GT + st[4] + lane];
                    const int widx = (e >> 12) & 15;
                    const unsigned has = __ballot_sync(0xffffffffu, widx != 15);
                    if (widx != 15) {
                        const int x = (int)(short)(e >> 16), y = st[5] + lane;
                        ops[o + __popc(has & lt)] = make_int2((x & 0xffff) | (y << 16), (int)((e & 0xfffu) * 32u) | (widx << 16));
                    }
                } else {
                    const unsigned rrl = (unsigned)ginfo[st[7] * V + v].w;
                    if (lane < CELLS && ((rrl >> lane) & 1u))
                        ops[o + __popc(rrl & lt)] = make_int2(xy[(st[7] * V + v) * CELLS + lane], (__popc(rrl & lt) * 128) | (kStBlockMap << 16));
                }
                if (lane == 0) sdesc[s + k] = make_int4(st[1], 0, o | (nops << 16), v | (st[0] << 8));
                o += nops;
            }
        }
    }
    // ---- P5: every (row, view, cell)'s tap offsets inside its stage --------------------------------------------------
    for (int v0 = 0; v0 < V; v0 += GPW) {
        const int v = v0 + gl;
        if (v < V) {
            const int e = (r * V + v) * CELLS + c;
            const int q = xy[e];
            const int x0 = (int)(short)(q & 0xffff), y0 = q >> 16;
            const unsigned m = mask[r * V + v];
            const bool real = ((m >> c) & 1u) && x0 != -2;
            int2 o = make_int2(0, 0);
            if (real) {
                const int4 gi = ginfo[r * V + v];
                if (gi.x == ST_SPANS) {
                    const unsigned* tr = gtab + v * GT + gi.y + (y0 - gi.z);
                    const unsigned a = tr[0], bq = tr[1];
                    o.x = ((int)(a & 0xfffu) + x0 - (int)(short)(a >> 16)) * 512;
                    o.y = ((int)(bq & 0xfffu) + x0 - (int)(short)(bq >> 16)) * 512;
                } else {
                    const int k = __popc((unsigned)gi.w & ((2u << c) - 1u)) - 1;  // the block of the last reload at or before this cell
                    o.x = k * 2048;
                    o.y = k * 2048 + 1024;
                }
            }
            ent[e] = o;
        }
    }
    __syncthreads();

    // ---- P6: every warp's per-stage words ---------------------------------------------------------------------------
    for (int s = lane; s < nst; s += 32) {
        const int w = sdesc[s].w;
        const int v = w & 0xff;
        const unsigned m = mask[r * V + v];
        const bool mine = ((w >> (8 + r)) & 1) && (m & 0xffffu);
        wst[r * L.nst_max + s] = mine ? ((m & 0xffu) | (((m >> 16) & 0xffu) << 8) | ((unsigned)v << 24)) : 0u;
    }
    // the scratch arrays were written and read through the generic proxy; the TMA unit (async proxy) writes the ring next
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
}

// ---- the copy schedule of one run (warp 0) ---------------------------------------------------------------------------
// The PL = P * nst stages of P consecutive items get consecutive places in the ring (a stage is contiguous: one that does
// not fit before the end starts again at offset 0), then every stage learns how many stages back its PREDECESSOR is: the
// last stage before it whose bytes it overwrites, which every warp must have left before the copy may be issued.  The
// schedule repeats with period PL (P = all items of the run when they fit the table: then there is no repetition at
// all).  Looking back at most kStBack stages and never letting a predecessor be older than the one of the stage before
// are both on the safe side: a later stage is waited for than strictly necessary.
__device__ __forceinline__ void staged_schedule(const StagedSmem& L, unsigned char* sm, int nst, int PL, int ring_bytes, int lane) {
    int2* sched = reinterpret_cast<int2*>(sm + L.sched);
    const int4* sdesc = reinterpret_cast<const int4*>(sm + L.sdesc);
    if (lane == 0) {
        int pos = 0, s = 0;
        for (int q = 0; q < PL; ++q) {
            const int bts = sdesc[s].x;
            const int al = (bts + 127) & ~127;
            if (pos + al > ring_bytes) pos = 0;
            sched[q] = make_int2(pos >> 7, bts);
            pos += al;
            if (++s == nst) s = 0;
        }
    }
    __syncwarp();
    const int kmax = min(PL, kStBack);
    for (int q = lane; q < PL; q += 32) {
        const int2 me = sched[q];
        const int so = (me.x & 0xffff) << 7, sb = me.y;
        int dback = kmax;
        if (sb > 0) {
            for (int k = 1; k < kmax; ++k) {
                int q2 = q - k;
                if (q2 < 0) q2 += PL;
                const int2 o = sched[q2];
                const int qo = (o.x & 0xffff) << 7;
                if (o.y > 0 && qo < so + sb && so < qo + o.y) { dback = k; break; }
            }
        }
        sched[q].x = (me.x & 0xffff) | (dback << 16);
    }
    __syncwarp();
    if (lane == 0) {  // predecessors in non-decreasing order, also across the period boundary
        int prev = sched[PL - 1].x >> 16;
        for (int pass = 0; pass < 2; ++pass)
            for (int q = 0; q < PL; ++q) {
                const int x = sched[q].x;
                int d = x >> 16;
                if (d > prev + 1) { d = prev + 1; sched[q].x = (x & 0xffff) | (d << 16); }
                prev = d;
            }
    }
}

// ---- the kernel ---------------------------------------------------------------------------------------------------
// KMODE: KM_ACC = sum / mean (fusion.py:18-21), KM_MAX = max over views with the zeros of views that miss a cell
// (fusion.py:22).  PROBE (timing aids, results are NOT the fusion): 1 = no TMA copies are issued (instruction side alone),
// 2 = copies but no blend (memory side alone).
// ring_bytes: size of the stage ring; cap: largest stage (bytes); lag: a stage is armed `lag` stages after its
// predecessor in the ring was walked (1 = as soon as possible: the arming warp then waits for the slowest warp).
template <typename TIn, typename TOut, int NW, int MAXREG, int KMODE, int PROBE = 0>
__global__ void __maxnreg__(MAXREG) warp_fuse_staged_kernel(const FwdParams p, int fpc, int ring_bytes, int cap, int lag,
                                                                 const __grid_constant__ StagedMaps maps, unsigned char* dump) {
    using VT = VecTraits<TIn>;
    constexpr int VE = VT::VE, P = VT::P, CELLS = kStCells, NT = NW * 32;
    constexpr int ILP = (P > 2) ? 2 : P;
    extern __shared__ __align__(128) unsigned char smem_st[];
    const int V = p.V;
    const StagedSmem L(V, NW);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int ty = blockIdx.x / p.tiles_x, tx = blockIdx.x - ty * p.tiles_x;
    const int i0 = ty * NW, j0 = tx * CELLS;
    const int b0 = blockIdx.z * fpc, b1 = min(p.B, b0 + fpc);
    const int r = warp, i = i0 + r;
    const int chunks = (p.C + 32 * VE - 1) / (32 * VE);
    const int last_vec = p.C / VE - 1;
    const float Vf = (float)V;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem_st);
    const uint32_t s_wts = sbase + L.wts + r * V * CELLS * 16, s_ent = sbase + L.ent + r * V * CELLS * 8;
    const uint32_t s_wst = sbase + L.wst + r * L.nst_max * 4, s_sdesc = sbase + L.sdesc, s_ops = sbase + L.ops;
    const uint32_t s_full = sbase + L.bars, s_empty = s_full + kStBack * 8;
    const uint32_t s_ring = sbase + L.ring;
    uint32_t lring = s_ring + lane * 16;
    asm volatile("" : "+r"(lring));

    for (int b = b0; b < b1;) {
        if (b > b0) {
            __syncthreads();  // every warp is done with the previous run's tables and ring
            for (int s = tid; s < 2 * kStBack; s += NT) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(s_full + s * 8) : "memory");
            __syncthreads();
        }
        for (int s = tid; s < kStBack; s += NT) { mbar_init(s_full + s * 8, 1); mbar_init(s_empty + s * 8, NW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        staged_build<NW, KMODE == KM_MAX>(p, L, smem_st, cap, i0, j0, b);
        // ---- the run of frames b .. e-1 shares these tables: same calibration, bit for bit (static cameras) ----------
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
        const int n_items = (e - b) * chunks;  // (frame, chunk) items, frame-major
        b = e;
        const int nst = __shfl_sync(0xffffffffu, lds4i(sbase + L.misc), 0);
        const int total = n_items * nst;  // stages of this run
        // the copy schedule: one period = as many whole items as the table holds (normally the whole run)
        const int PL = nst > 0 ? max(1, min(n_items, kStSched / nst)) * nst : 0;
        if (warp == 0 && nst > 0) staged_schedule(L, smem_st, nst, PL, ring_bytes, lane);
        __syncthreads();
        if (dump) {  // development aid (BEVIPM_ST_DUMP): the tables and the copy schedule of every tile, no phase B
            unsigned char* dst = dump + ((size_t)blockIdx.z * gridDim.x + blockIdx.x) * (size_t)L.ring;
            for (int z = tid * 4; z < L.ring; z += NT * 4) *reinterpret_cast<int*>(dst + z) = *reinterpret_cast<const int*>(smem_st + z);
            return;
        }

        const uint32_t s_sched = sbase + L.sched;
        if (PROBE == 1) {  // no copies: blend zeros instead of whatever the ring holds (NaNs would take the slow division)
            for (int z = tid * 16; z < ring_bytes; z += NT * 16) *reinterpret_cast<uint4*>(smem_st + L.ring + z) = make_uint4(0u, 0u, 0u, 0u);
            __syncthreads();
        }

        // arm stage ga = (item ia, stage sa), schedule entry qa: warp-uniform; one elected lane talks to the TMA unit
        auto arm = [&](int ga, int qa) {
            const int ia = ga / nst, sa = ga - ia * nst;
            const int4 sd = lds16i(s_sdesc + sa * 16);
            const int2 sc = lds8i(s_sched + qa * 8);
            const int pred = ga - (sc.x >> 16);
            // the last stage that used these bytes: every warp must have left it.  (That stage is at most 64 back, so
            // this also frees the mbarrier pair of stage ga - 64, which is the one stage ga uses.)
            if (pred >= 0) mbar_wait(s_empty + (pred & (kStBack - 1)) * 8, (uint32_t)(pred >> 6) & 1u);
            if (elect_one()) {
                const uint32_t bar = s_full + (ga & (kStBack - 1)) * 8, dst0 = s_ring + ((uint32_t)(sc.x & 0xffff) << 7);
                mbar_expect_tx(bar, sd.x);
                if (PROBE != 1) {
                    const int o0 = sd.z & 0xffff, n = sd.z >> 16;
                    const int fim = ia / chunks, km = ia - fim * chunks;
                    const int c0 = km * 32 * VE, bb = b_run + fim;
                    for (int q = 0; q < n; ++q) {
                        const int2 op = lds8i(s_ops + (o0 + q) * 8);
                        const int x = (int)(short)(op.x & 0xffff), y = op.x >> 16;
                        tma_load_5d(dst0 + (uint32_t)(op.y & 0xffff) * 16u, &maps.m[op.y >> 16], c0, x, y, sd.w & 0xff, bb, bar);
                    }
                } else {
                    asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(sd.x) : "memory");
                }
            }
            __syncwarp();
        };
        // when stage ga is due: `lag` stages after its predecessor in the ring was walked, at the latest when it is walked itself
        auto due = [&](int ga, int qa) { return ga < total ? min(ga, ga - (lds4i(s_sched + qa * 8) >> 16) + lag) : 0x7fffffff; };

        int ga = 0, qa = 0;           // next stage to arm: global index, schedule entry
        int at = nst > 0 ? due(0, 0) : 0x7fffffff;
        int g = 0, qg = 0;            // stage being walked: global index, schedule entry
        float2 cur[4][P];
#pragma unroll
        for (int tap = 0; tap < 4; ++tap)
#pragma unroll
            for (int q = 0; q < P; ++q) cur[tap][q] = make_float2(0.0f, 0.0f);
        for (int it = 0; it < n_items; ++it) {
            float2 acc[CELLS][P];
#pragma unroll
            for (int c = 0; c < CELLS; ++c)
#pragma unroll
                for (int q = 0; q < P; ++q) acc[c][q] = (KMODE == KM_MAX) ? make_float2(-INFINITY, -INFINITY) : make_float2(0.0f, 0.0f);

            for (int s = 0; s < nst; ++s) {
                while (at <= g) {  // every warp follows the schedule; the warp on duty issues
                    if ((g & (NW - 1)) == warp) arm(ga, qa);
                    ++ga;
                    if (++qa == PL) qa = 0;
                    at = due(ga, qa);
                }
                const unsigned m = (unsigned)__shfl_sync(0xffffffffu, lds4i(s_wst + s * 4), 0);  // seen | reload << 8 | view << 24
                mbar_wait(s_full + (g & (kStBack - 1)) * 8, (uint32_t)(g >> 6) & 1u);  // the stage's bytes have landed
                if (m && PROBE != 2) {
                    const int v = (int)(m >> 24);
                    const uint32_t sb = lring + (((uint32_t)lds4i(s_sched + qg * 8) & 0xffffu) << 7);
                    const uint32_t wv = s_wts + v * (CELLS * 16), ev = s_ent + v * (CELLS * 8);
                    float4 wn = lds16f(wv);
#pragma unroll
                    for (int c = 0; c < CELLS; ++c) {
                        const float4 w = wn;
                        if (c + 1 < CELLS) wn = lds16f(wv + (c + 1) * 16);
                        if (!((m >> c) & 1u)) continue;  // the view does not see this cell (warp-uniform)
                        if ((m >> (8 + c)) & 1u) {       // the row leaves the block held in `cur`
                            const int2 o = lds8i(ev + c * 8);
                            const uint32_t a0 = sb + (uint32_t)o.x, a1 = sb + (uint32_t)o.y;
                            uint4 nxt[4];
                            nxt[0] = lds16(a0); nxt[1] = lds16(a0 + 512);
                            nxt[2] = lds16(a1); nxt[3] = lds16(a1 + 512);
#pragma unroll
                            for (int tap = 0; tap < 4; ++tap) VT::unpack(nxt[tap], cur[tap]);
                        }
                        // out_v = fma(SE,se, fma(SW,sw, fma(NE,ne, NW*nw)))   ATen's interpolation order
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
                            for (int q = 0; q < ILP; ++q) {
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
                __syncwarp();
                if (lane == 0) mbar_arrive(s_empty + (g & (kStBack - 1)) * 8);  // this warp is done with the stage's bytes
                ++g;
                if (++qg == PL) qg = 0;
            }

            // ---- epilogue: mean division (IEEE quotient) and one 16-byte store per cell ---------------------------
            const int fi_this = it / chunks, k_this = it - fi_this * chunks;
            const bool lok = k_this * 32 + lane <= last_vec;
            if (i >= p.Hb) continue;
            if constexpr (KMODE == KM_MAX) {
                const unsigned every = (unsigned)__shfl_sync(0xffffffffu, lds4i(sbase + L.misc + 4 * (1 + r)), 0);
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
            if (!lok) continue;
            TOut* oc = reinterpret_cast<TOut*>(p.out) + (long long)(b_run + fi_this) * p.os_b + (long long)i * p.os_y + (long long)j0 * p.os_x +
                       (k_this * 32 + lane) * VE;
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
