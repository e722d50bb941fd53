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
//            (cell_coord + make_tap); per view the taps' texel rows and, per texel row, the span of x they touch are
//            collected with shared-memory atomics, for the whole tile and for every BEV row on its own.  A view whose
//            tile-level spans fit one ring slot becomes ONE stage (a list of row copies: tensor-map boxes of
//            [256 channels x BW texels x 1 row], BW quantised to the widths the launcher encoded maps for); otherwise
//            every BEV row of the tile becomes its own stage, from its own spans or -- when even those do not fit (the
//            BEV is coarser than the source map there) -- from one [256 x 2 x 2] box per 2x2 block the row enters.
//            Out-of-map parts of a box are zero-filled by the TMA unit: exactly the reference's zero padding
//            (grid_sample padding_mode='zeros': a tap outside the map contributes 0 * w), so there is no tap mask,
//            no stand-in address and no special case for non-finite features.
//   phase B: the stage list of the tile is walked once per (frame, channel chunk) item.  A ring of D slots of S bytes
//            in shared memory holds the stages in flight; full[slot] (transaction bytes) / empty[slot] (one arrival
//            per warp) mbarriers order TMA writes and LDS reads.  The warps take turns at issuing: warp n % NW arms
//            stage n + LOOK (LOOK = D - 2: it waits for the slowest warp to leave stage n - 2, not n - 1) when it starts stage n.  A warp owns one BEV row segment of 8 cells, keeps the 8 cells'
//            accumulators in registers and walks a stage like the run kernel walks a view: on a reload bit it reads
//            the 2x2 block from the slot (4 x LDS.128), unpacks it once, and blends it for as many cells as stay in
//            the block.  Per cell the views are still added in ascending order: the reference's accumulation order.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ipm_run.cuh"

namespace bevipm {

constexpr int kStCells = 8;    // cells per row segment (one warp)
constexpr int kStRowsT = 32;   // texel rows a tile-level stage may span
constexpr int kStRowsR = 12;   // texel rows a row-level stage may span
constexpr int kStNumW = 13;    // span widths with a tensor map of their own
constexpr int kStBlockMap = 13;  // the [2 x 2] block map
constexpr int kStNumMaps = 14;

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
    int wts, ent, mask, stg, ops, misc, sH, bars, ring, total;
    int xy, yt, yr, tsp, rsp, trow, rrow, vinfo, rinfo, scratch_end;
    int max_ops;
    __host__ __device__ StagedSmem(int V, int R, int S, int D) {
        auto up = [](int x, int a) { return (x + a - 1) / a * a; };
        int o = 0;
        wts = o; o += R * V * kStCells * 16;             // float4 (nw, ne, sw, se) per (row, view, cell)
        ent = o; o += R * V * kStCells * 8;              // int2 (byte offset of the NW tap, of the SW tap) inside the slot
        mask = o; o += up(R * V * 4, 16);                // seen | reload << 16 per (row, view)
        stg = o; o += V * R * 16;                        // int4 {view, row mask, bytes, first op | ops << 16}
        max_ops = V * (kStRowsT > R * kStRowsR ? kStRowsT : R * kStRowsR);
        ops = o; o += max_ops * 8;                       // int2 {x | y << 16, slot offset / 16 | map << 16}
        misc = o; o += 128;                              // [0] stages, [1 + r] cells of row r every view sees
        sH = o; o += V * 48;                             // homographies, rows padded to 4 floats
        bars = o; o += up(2 * D * 8, 128);               // full[D], empty[D]
        ring = up(o, 128);
        total = ring + D * S;
        o = ring;
        xy = o; o += R * V * kStCells * 4;               // x0 | y0 << 16
        yt = o; o += up(2 * V * 4, 16);                  // tile-level min / max texel row per view
        yr = o; o += up(2 * R * V * 4, 16);              // row-level
        tsp = o; o += V * kStRowsT * 8;                  // tile-level x spans per (view, texel row): lo, hi
        rsp = o; o += R * V * kStRowsR * 8;              // row-level
        trow = o; o += V * kStRowsT * 4;                 // planned rows: slot offset in texels | map << 12 | x lo << 16
        rrow = o; o += R * V * kStRowsR * 4;
        vinfo = o; o += V * 32;                          // per view: kind, stages, ops, bytes, row mask
        rinfo = o; o += R * V * 16;                      // per (row, view): kind, ops, bytes, real-reload mask
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

enum { ST_NONE = 0, ST_TILE = 1, ST_ROWS = 2, ST_SPANS = 1, ST_BLOCKS = 2 };

// ---- phase A ------------------------------------------------------------------------------------------------------
// Builds, for the tile at (i0, j0) of frame b: blend weights, reload masks, the stage list with its row copies and
// every (row, view, cell)'s tap offsets inside its stage.  Five block barriers; the scratch arrays live in the ring.
template <int NW, bool WANT_ALL_SEEN>
__device__ __forceinline__ void staged_build(const FwdParams& p, const StagedSmem& L, unsigned char* sm, int S, int i0, int j0, int b) {
    constexpr int CELLS = kStCells, GPW = 32 / CELLS, NT = NW * 32;
    constexpr unsigned CMASK = (1u << CELLS) - 1u;
    const int V = p.V;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gl = lane / CELLS, c = lane - gl * CELLS;
    const unsigned lt = (1u << lane) - 1u;
    float4* wts = reinterpret_cast<float4*>(sm + L.wts);
    int2* ent = reinterpret_cast<int2*>(sm + L.ent);
    unsigned* mask = reinterpret_cast<unsigned*>(sm + L.mask);
    int4* stg = reinterpret_cast<int4*>(sm + L.stg);
    int2* ops = reinterpret_cast<int2*>(sm + L.ops);
    int* misc = reinterpret_cast<int*>(sm + L.misc);
    float* sH = reinterpret_cast<float*>(sm + L.sH);
    int* xy = reinterpret_cast<int*>(sm + L.xy);
    int* yt = reinterpret_cast<int*>(sm + L.yt);        // [v] lo, [V + v] hi
    int* yr = reinterpret_cast<int*>(sm + L.yr);        // [r * V + v] lo, [R * V + ...] hi
    int* tsp = reinterpret_cast<int*>(sm + L.tsp);      // [(v * 32 + row) * 2] lo, + 1 hi
    int* rsp = reinterpret_cast<int*>(sm + L.rsp);      // [((r * V + v) * 12 + row) * 2]
    unsigned* trow = reinterpret_cast<unsigned*>(sm + L.trow);
    unsigned* rrow = reinterpret_cast<unsigned*>(sm + L.rrow);
    int* vinfo = reinterpret_cast<int*>(sm + L.vinfo);  // [v * 8 + {kind, stages, ops, bytes, row mask}]
    int* rinfo = reinterpret_cast<int*>(sm + L.rinfo);  // [(r * V + v) * 4 + {kind, ops, bytes, real reloads}]
    constexpr int IMAX = 0x7fffffff, IMIN = (int)0x80000000;

    // ---- P0: homographies (geometry.py:60-63), scratch initialisation -------------------------------------------
    if (tid < V) {
        float H[9];
        homography(p.K + 9 * (b * V + tid), p.Rt + 12 * (b * V + tid), H);
#pragma unroll
        for (int q = 0; q < 3; ++q) reinterpret_cast<float4*>(sH + 12 * tid)[q] = make_float4(H[3 * q], H[3 * q + 1], H[3 * q + 2], 0.0f);
    }
    for (int z = tid; z < V; z += NT) { yt[z] = IMAX; yt[V + z] = IMIN; }
    for (int z = tid; z < V * kStRowsT; z += NT) { tsp[2 * z] = IMAX; tsp[2 * z + 1] = IMIN; }
    for (int z = tid; z < NW * V * kStRowsR; z += NT) { rsp[2 * z] = IMAX; rsp[2 * z + 1] = IMIN; }
    __syncthreads();

    // ---- P1: project this warp's row; weights, masks, row-level texel-row range and spans --------------------------
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
                if (ymin <= ymax) { atomicMin(yt + v, ymin); atomicMax(yt + V + v, ymax); }
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

    // ---- P2: tile-level spans ---------------------------------------------------------------------------------------
    for (int v0 = 0; v0 < V; v0 += GPW) {
        const int v = v0 + gl;
        if (v < V) {
            const int q = xy[(r * V + v) * CELLS + c];
            const int x0 = (int)(short)(q & 0xffff), y0 = q >> 16;
            const bool real = ((mask[r * V + v] >> c) & 1u) && x0 != -2;
            const int lo = yt[v], hi = yt[V + v];
            if (real && hi - lo + 2 <= kStRowsT) {
                int* sp = tsp + (v * kStRowsT + (y0 - lo)) * 2;
                atomicMin(sp, x0); atomicMax(sp + 1, x0 + 1);
                atomicMin(sp + 2, x0); atomicMax(sp + 3, x0 + 1);
            }
        }
    }
    __syncthreads();

    // ---- P3: plan.  Warp w owns views w, w + NW, ...: lane = texel row of the view's range ---------------------------
    auto plan_rows = [&](const int* sp, int nrows, bool ok_in, unsigned* out, int& nops, int& bytes) -> bool {
        // sp: spans of the `nrows` texel rows; out[row] = slot offset in texels | map << 12 | lo << 16 (map 15: no copy)
        int lo = IMAX, hi = IMIN;
        if (ok_in && lane < nrows) { lo = sp[2 * lane]; hi = sp[2 * lane + 1]; }
        const int w = (lo <= hi) ? hi - lo + 1 : 0;
        const int idx = w ? st_width_index(w) : -1;
        const bool bad = w && idx < 0;
        const int wq = (w && idx >= 0) ? st_width(idx) : 0;
        int incl = wq;
#pragma unroll
        for (int o = 1; o < 32; o *= 2) {
            const int u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        const bool ok = ok_in && !__any_sync(0xffffffffu, bad) && total * 512 <= S;
        if (ok && lane < nrows) out[lane] = (unsigned)(incl - wq) | ((unsigned)(w ? idx : 15) << 12) | ((unsigned)(lo & 0xffff) << 16);
        nops = __popc(__ballot_sync(0xffffffffu, w > 0));
        bytes = total * 512;
        return ok;
    };
    for (int v = warp; v < V; v += NW) {
        const int tlo = yt[v], thi = yt[V + v];
        const bool tile_range = tlo <= thi && thi - tlo + 2 <= kStRowsT;
        const unsigned rowmask = __ballot_sync(0xffffffffu, lane < NW && (mask[(lane < NW ? lane : 0) * V + v] & 0xffffu) != 0);
        int nops = 0, bytes = 0;
        const bool tile_ok = plan_rows(tsp + v * kStRowsT * 2, kStRowsT, tile_range, trow + v * kStRowsT, nops, bytes);
        int kind = ST_NONE, nst = 0;
        if (tile_ok) {
            kind = ST_TILE; nst = 1;
        } else if (rowmask) {
            kind = ST_ROWS; nops = 0; bytes = 0;
            for (int rr = 0; rr < NW; ++rr) {
                int* ri = rinfo + (rr * V + v) * 4;
                if (!((rowmask >> rr) & 1u)) { if (lane == 0) ri[0] = ST_NONE; continue; }
                const int rlo = yr[rr * V + v], rhi = yr[NW * V + rr * V + v];
                const bool any_real = rlo <= rhi;
                int rn = 0, rb = 0;
                const bool row_ok = plan_rows(rsp + (rr * V + v) * kStRowsR * 2, kStRowsR, any_real && rhi - rlo + 2 <= kStRowsR,
                                              rrow + (rr * V + v) * kStRowsR, rn, rb);
                // reloads of cells that have texels (a non-finite sample position has none)
                const unsigned m = mask[rr * V + v];
                const int q = xy[(rr * V + v) * CELLS + (lane & (CELLS - 1))];
                const unsigned rrl = __ballot_sync(0xffffffffu, lane < CELLS && ((m >> (16 + lane)) & 1u) && (int)(short)(q & 0xffff) != -2);
                int rk = ST_SPANS;
                if (!any_real) { rn = 0; rb = 0; }                       // only NaN cells: a stage without copies
                else if (!row_ok) { rk = ST_BLOCKS; rn = __popc(rrl); rb = rn * 2048; }
                if (lane == 0) { ri[0] = rk; ri[1] = rn; ri[2] = rb; ri[3] = (int)rrl; }
                ++nst; nops += rn;
            }
        }
        if (lane == 0) {
            int* vi = vinfo + v * 8;
            vi[0] = kind; vi[1] = nst; vi[2] = nops; vi[3] = bytes; vi[4] = (int)rowmask;
        }
    }
    __syncthreads();

    // ---- P4: emit the stage list and the row copies (views ascending: the accumulation order) --------------------
    {
        int my_st = 0, my_op = 0;  // lane = view: exclusive prefix sums of stages and ops
        if (lane < V) { my_st = vinfo[lane * 8 + 1]; my_op = vinfo[lane * 8 + 2]; }
        int inc_st = my_st, inc_op = my_op;
#pragma unroll
        for (int o = 1; o < 32; o *= 2) {
            const int a = __shfl_up_sync(0xffffffffu, inc_st, o), bq = __shfl_up_sync(0xffffffffu, inc_op, o);
            if (lane >= o) { inc_st += a; inc_op += bq; }
        }
        if (tid == 31) misc[0] = inc_st;  // stages of the tile
        for (int v = warp; v < V; v += NW) {
            int s = __shfl_sync(0xffffffffu, inc_st - my_st, v), o = __shfl_sync(0xffffffffu, inc_op - my_op, v);
            const int* vi = vinfo + v * 8;
            const int kind = vi[0];
            if (kind == ST_TILE) {
                const int tlo = yt[v];
                const unsigned e = trow[v * kStRowsT + lane];
                const int idx = (e >> 12) & 15;
                const unsigned has = __ballot_sync(0xffffffffu, idx != 15);
                if (idx != 15) {
                    const int x = (int)(short)(e >> 16), y = tlo + lane;
                    ops[o + __popc(has & lt)] = make_int2((x & 0xffff) | (y << 16), (int)((e & 0xfffu) * 32u) | (idx << 16));
                }
                if (lane == 0) stg[s] = make_int4(v, vi[4], vi[3], o | (vi[2] << 16));
            } else if (kind == ST_ROWS) {
                const unsigned rowmask = (unsigned)vi[4];
                for (int rr = 0; rr < NW; ++rr) {
                    if (!((rowmask >> rr) & 1u)) continue;
                    const int* ri = rinfo + (rr * V + v) * 4;
                    const int rk = ri[0], rn = ri[1];
                    if (rk == ST_SPANS && rn > 0) {
                        const int rlo = yr[rr * V + v];
                        unsigned e = 15u << 12;
                        if (lane < kStRowsR) e = rrow[(rr * V + v) * kStRowsR + lane];
                        const int idx = (e >> 12) & 15;
                        const unsigned has = __ballot_sync(0xffffffffu, idx != 15);
                        if (idx != 15) {
                            const int x = (int)(short)(e >> 16), y = rlo + lane;
                            ops[o + __popc(has & lt)] = make_int2((x & 0xffff) | (y << 16), (int)((e & 0xfffu) * 32u) | (idx << 16));
                        }
                    } else if (rk == ST_BLOCKS) {
                        const unsigned rrl = (unsigned)ri[3];
                        if (lane < CELLS && ((rrl >> lane) & 1u)) {
                            const int q = xy[(rr * V + v) * CELLS + lane];
                            ops[o + __popc(rrl & lt)] = make_int2(q, (__popc(rrl & lt) * 128) | (kStBlockMap << 16));
                        }
                    }
                    if (lane == 0) stg[s] = make_int4(v, (int)(1u << rr), ri[2], o | (rn << 16));
                    ++s; o += rn;
                }
            }
        }
    }
    // ---- P5: every (row, view, cell)'s tap offsets inside its stage's slot ------------------------------------------
    for (int v0 = 0; v0 < V; v0 += GPW) {
        const int v = v0 + gl;
        if (v < V) {
            const int e = (r * V + v) * CELLS + c;
            const int q = xy[e];
            const int x0 = (int)(short)(q & 0xffff), y0 = q >> 16;
            const bool real = ((mask[r * V + v] >> c) & 1u) && x0 != -2;
            int2 o = make_int2(0, 0);
            if (real) {
                const int kind = vinfo[v * 8];
                if (kind == ST_TILE) {
                    const unsigned* tr = trow + v * kStRowsT + (y0 - yt[v]);
                    const unsigned a = tr[0], bq = tr[1];
                    o.x = ((int)(a & 0xfffu) + x0 - (int)(short)(a >> 16)) * 512;
                    o.y = ((int)(bq & 0xfffu) + x0 - (int)(short)(bq >> 16)) * 512;
                } else {
                    const int* ri = rinfo + (r * V + v) * 4;
                    if (ri[0] == ST_SPANS) {
                        const unsigned* tr = rrow + (r * V + v) * kStRowsR + (y0 - yr[r * V + v]);
                        const unsigned a = tr[0], bq = tr[1];
                        o.x = ((int)(a & 0xfffu) + x0 - (int)(short)(a >> 16)) * 512;
                        o.y = ((int)(bq & 0xfffu) + x0 - (int)(short)(bq >> 16)) * 512;
                    } else {
                        const unsigned rrl = (unsigned)ri[3];
                        const int k = __popc(rrl & ((2u << c) - 1u)) - 1;  // the block of the last reload at or before this cell
                        o.x = k * 2048;
                        o.y = k * 2048 + 1024;
                    }
                }
            }
            ent[e] = o;
        }
    }
    // the scratch arrays were written and read through the generic proxy; the TMA unit (async proxy) writes the ring next
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
}

// ---- the kernel ---------------------------------------------------------------------------------------------------
// KMODE: KM_ACC = sum / mean (fusion.py:18-21), KM_MAX = max over views with the zeros of views that miss a cell
// (fusion.py:22).  PROBE (timing aid, results are NOT the fusion): 1 = no TMA copies are issued.
template <typename TIn, typename TOut, int NW, int MAXREG, int KMODE, int PROBE = 0>
__global__ void __maxnreg__(MAXREG) warp_fuse_staged_kernel(const FwdParams p, int fpc, int S, int D, int LOOK, const __grid_constant__ StagedMaps maps,
                                                                 unsigned char* dump) {
    using VT = VecTraits<TIn>;
    constexpr int VE = VT::VE, P = VT::P, CELLS = kStCells, NT = NW * 32;
    constexpr int ILP = (P > 2) ? 2 : P;
    extern __shared__ __align__(128) unsigned char smem_st[];
    const int V = p.V;
    const StagedSmem L(V, NW, S, D);
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
    const uint32_t s_mask = sbase + L.mask + r * V * 4, s_stg = sbase + L.stg, s_ops = sbase + L.ops;
    const uint32_t s_full = sbase + L.bars, s_empty = s_full + D * 8;
    const uint32_t s_ring = sbase + L.ring;
    uint32_t lring = s_ring + lane * 16;
    asm volatile("" : "+r"(lring));

    for (int b = b0; b < b1;) {
        if (b > b0) __syncthreads();  // every warp is done with the previous run's tables and slots
        if (tid == 0) {
            if (b > b0)
                for (int s = 0; s < 2 * D; ++s) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(s_full + s * 8) : "memory");
            for (int s = 0; s < D; ++s) { mbar_init(s_full + s * 8, 1); mbar_init(s_empty + s * 8, NW); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        staged_build<NW, KMODE == KM_MAX>(p, L, smem_st, S, i0, j0, b);
        if (dump) {  // development aid (BEVIPM_ST_DUMP): the tables of every tile, no phase B
            unsigned char* dst = dump + ((size_t)blockIdx.z * gridDim.x + blockIdx.x) * (size_t)L.ring;
            for (int z = tid * 4; z < L.ring; z += NT * 4) *reinterpret_cast<int*>(dst + z) = *reinterpret_cast<const int*>(smem_st + z);
            return;
        }

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
        const unsigned total = (unsigned)n_items * (unsigned)nst;  // stages of this run

        // arm stage m (warp-uniform; one elected lane talks to the TMA unit)
        auto arm = [&](unsigned m) {
            if (m >= total) return;
            const unsigned use = m / (unsigned)D, slot = m - use * (unsigned)D;
            if (use > 0) mbar_wait(s_empty + slot * 8, (use - 1u) & 1u);  // every warp has left the stage this slot held
            const unsigned im = m / (unsigned)nst, sm_ = m - im * (unsigned)nst;
            const unsigned fim = im / (unsigned)chunks, km = im - fim * (unsigned)chunks;
            const int4 sd = lds16i(s_stg + sm_ * 16);
            if (elect_one()) {
                const uint32_t bar = s_full + slot * 8, dst0 = s_ring + slot * (unsigned)S;
                mbar_expect_tx(bar, sd.z);
                if (PROBE != 1) {
                    const int o0 = sd.w & 0xffff, n = sd.w >> 16;
                    const int c0 = (int)km * 32 * VE, bb = b_run + (int)fim;
                    for (int q = 0; q < n; ++q) {
                        const int2 op = lds8i(s_ops + (o0 + q) * 8);
                        const int x = (int)(short)(op.x & 0xffff), y = op.x >> 16;
                        tma_load_5d(dst0 + (uint32_t)(op.y & 0xffff) * 16u, &maps.m[op.y >> 16], c0, x, y, sd.x, bb, bar);
                    }
                } else {
                    asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(sd.z) : "memory");
                }
            }
            __syncwarp();
        };

        if (nst > 0 && warp == 0)
            for (int m = 0; m < LOOK; ++m) arm((unsigned)m);

        unsigned n = 0;       // stage counter of the run
        int slot = 0;         // n % D
        unsigned phase = 0;   // (n / D) & 1
        int duty = 0;         // n % NW
        int fi_c = 0, k_c = 0;
        for (int it = 0; it < n_items; ++it) {
            float2 acc[CELLS][P];
#pragma unroll
            for (int c = 0; c < CELLS; ++c)
#pragma unroll
                for (int q = 0; q < P; ++q) acc[c][q] = (KMODE == KM_MAX) ? make_float2(-INFINITY, -INFINITY) : make_float2(0.0f, 0.0f);
            float2 cur[4][P];
#pragma unroll
            for (int tap = 0; tap < 4; ++tap)
#pragma unroll
                for (int q = 0; q < P; ++q) cur[tap][q] = make_float2(0.0f, 0.0f);

            for (int s = 0; s < nst; ++s) {
                if (duty == warp) arm(n + (unsigned)LOOK);
                const int4 sd = lds16i(s_stg + s * 16);
                const int v = __shfl_sync(0xffffffffu, sd.x, 0);
                const unsigned m = (unsigned)__shfl_sync(0xffffffffu, lds4i(s_mask + 4 * v), 0);
                const bool mine = ((__shfl_sync(0xffffffffu, sd.y, 0) >> r) & 1) && (m & 0xffffu);
                // Every warp waits for every stage, also the ones that hold nothing for its row: a warp that skipped the wait
                // could run a whole ring revolution ahead, where the parity of its next wait on this slot would alias an
                // older phase (and its `empty` arrival would be counted for the stage before).
                mbar_wait(s_full + slot * 8, phase);  // the stage's bytes have landed
                if (mine) {
                    const uint32_t sb = lring + (uint32_t)slot * (uint32_t)S;
                    const uint32_t wv = s_wts + v * (CELLS * 16), ev = s_ent + v * (CELLS * 8);
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
                __syncwarp();
                if (lane == 0) mbar_arrive(s_empty + slot * 8);  // this warp is done with the slot
                ++n;
                if (++slot == D) { slot = 0; phase ^= 1u; }
                if (++duty == NW) duty = 0;
            }

            // ---- epilogue: mean division (IEEE quotient) and one 16-byte store per cell ---------------------------
            const int k_this = k_c, fi_this = fi_c;
            if (++k_c == chunks) { k_c = 0; ++fi_c; }
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
