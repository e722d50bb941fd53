// bevipm_proj.cu -- the per-view 1x1 projection of FoldedConcatProjIPM as a hand-written tcgen05 GEMM.
//
// BEVNet concatenates the V warped maps and applies nn.Conv2d(V*C, Co, 1) (model_wrapper.py:68-73).  The warp is linear
// per channel, so the projection can run on the small SOURCE maps, view by view, before the fused SUM kernel
// (bevipm/modules.py FoldedConcatProjIPM):   g[b,v,y,x,:] = W_v . f[b,v,y,x,:]   with W_v = proj.weight[:, v*C:(v+1)*C].
// That is V*B independent GEMMs  [Hf*Wf, C] x [C, Co]  on channels-last fp32 maps -- read-once operands of 1.2 GB per
// frame at wildtrack.yaml's sizes, so the kernel is HBM-bound by design and the tensor cores only have to keep up:
//
//   * one CTA = one 128-row tile of one (frame, view) map and all Co output channels; 2 CTAs per SM when Co <= 128;
//   * warp 0 (one elected lane) streams k-blocks of 32 channels through a ring of shared-memory stages with TMA
//     (cp.async.bulk.tensor.3d, 128-byte swizzle: A box [32 ch x 128 rows], B box [32 ch x 1 view x Co rows of W]);
//     rows past the end of the map and channels past C are zero-filled by the copy engine;
//   * warp 1 (one elected lane) issues tcgen05.mma.kind::tf32 (M = 128, N = Co, K = 8 per instruction) from shared-
//     memory descriptors into a TMEM accumulator and frees a stage with tcgen05.commit when its MMAs have read it;
//   * fp32-grade mode (passes = 3): two converter warps split every landed A tile in place into its TF32 head
//     (x & 0xffffe000) and an exact fp32 remainder (x - head) in a second buffer; the weights come pre-split from the
//     host side (W_hi, W_lo); three MMAs per k-step accumulate  hi*hi + lo*hi + hi*lo  (error ~2^-21 per product, the
//     usual 3xTF32 scheme), so the result is comparable with an fp32 FMA loop.  passes = 1 feeds the raw fp32 bits
//     (the tensor core reads the TF32 head itself): what cuDNN does for the reference's Conv2d under torch's default
//     allow_tf32 = True;
//   * warps 2-5 read the accumulator back (tcgen05.ld 32x32b), stage it in the (now idle) ring with the 128-byte
//     swizzle and one lane writes it with TMA stores (clipped at the end of the map).
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bevipm.h"

namespace bevipm {
void note_launch(int variant);
int set_error(int code, const char* msg);

namespace {

constexpr int kBM = 128;  // rows of one tile = UMMA M
constexpr int kBK = 32;   // fp32 channels per k-block: one 128-byte swizzle row
constexpr int kUK = 8;    // K of one tcgen05.mma.kind::tf32
constexpr int kThreads = 256;  // warp 0 TMA, 1 MMA, 2-5 epilogue, 6-7 A-tile converters (passes = 3)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pj_mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void pj_mbar_expect_tx(uint32_t bar, int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pj_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void pj_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (int spin = 0; !done; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (spin > (1 << 26)) __trap();  // a lost copy or MMA must not hang the device
    }
}
__device__ __forceinline__ void pj_tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void pj_tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ bool pj_elect() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// shared-memory matrix descriptor, K-major, 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart
// (start address >> 4 | LBO 1 (ignored with a swizzle) | SBO 1024 >> 4 | descriptor version 1 | SWIZZLE_128B)
__device__ __forceinline__ uint64_t pj_desc(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3fffu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void pj_mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void pj_commit(uint32_t bar) {  // arrives when every MMA issued so far has completed
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

struct ProjArgs {
    int kblocks;     // ceil(C / 32)
    int tiles_m;     // ceil(rows / 128)
    int tiles;       // B*V * tiles_m
    int V;           // views (the weight of view bv % V)
    int Co;          // output channels = UMMA N (multiple of 16, <= 256)
    int stages;      // ring depth
    int stage_bytes; // bytes of one stage (A [+ A_lo] + B [+ B_lo])
    int passes;      // 1: TF32, 3: split operands (fp32-grade)
    int tmem_cols;   // power of two >= 2 * Co: two accumulators
    int acc_cols;    // columns between the two accumulators
    int out_off;     // byte offset of the epilogue's staging area behind the ring
    int burst;       // k-blocks the producer requests back to back
    int out_sub;     // [128 rows x 32 columns] sub-tiles the staging area holds (the tile leaves in rounds of that many)
    uint32_t idesc;  // tcgen05 instruction descriptor
};

// ma: x as [C, rows, B*V]; mb / mbl: W_hi / W_lo as [C, V, Co]; md: out as [Co, rows, B*V] (all fp32, 128-byte swizzle).
// Persistent: one CTA per SM walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...; the k-block ring runs on across tile borders
// and the accumulator is double-buffered in TMEM, so the epilogue of tile i overlaps the loads and MMAs of tile i + 1.
__global__ void __launch_bounds__(kThreads, 1) proj1x1_tf32_kernel(const __grid_constant__ CUtensorMap ma, const __grid_constant__ CUtensorMap mb,
                                                                   const __grid_constant__ CUtensorMap mbl, const __grid_constant__ CUtensorMap md,
                                                                   const ProjArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_pj[];
    // full[8], ready[8] (A tile split), empty[8]; accumulator full[2] / empty[2]
    __shared__ __align__(8) unsigned long long bars[3 * 8 + 4];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t ring = (smem_u32(smem_pj) + 1023u) & ~1023u;  // the swizzle pattern repeats every 1024 bytes
    const uint32_t full0 = smem_u32(&bars[0]), ready0 = smem_u32(&bars[8]), empty0 = smem_u32(&bars[16]);
    const uint32_t accf0 = smem_u32(&bars[24]), acce0 = smem_u32(&bars[26]);
    const bool split = a.passes == 3;
    const int a_bytes = kBM * 128, b_bytes = a.Co * 128;
    const int off_alo = a_bytes, off_b = split ? 2 * a_bytes : a_bytes, off_blo = off_b + b_bytes;

    if (tid == 0) {
        for (int s = 0; s < a.stages; ++s) {
            pj_mbar_init(full0 + 8 * s, 1);
            pj_mbar_init(ready0 + 8 * s, 64);  // every lane of the two converter warps
            pj_mbar_init(empty0 + 8 * s, 1);
        }
        for (int q = 0; q < 2; ++q) {
            pj_mbar_init(accf0 + 8 * q, 1);    // tcgen05.commit of the tile's last MMA
            pj_mbar_init(acce0 + 8 * q, 128);  // every epilogue thread, after its last tcgen05.ld of the tile
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: allocated (and later freed) by the MMA warp
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)a.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;

    if (warp == 0) {
        // ===== TMA producer =====
        if (pj_elect()) {
            const int bytes = a_bytes + (split ? 2 : 1) * b_bytes;
            int g = 0;  // k-blocks issued so far, over all tiles of this CTA
            for (int tile = blockIdx.x; tile < a.tiles; tile += gridDim.x) {
                const int bv = tile / a.tiles_m, m0 = (tile - bv * a.tiles_m) * kBM, v = bv % a.V;
                // k-blocks are requested in bursts: the 128-byte pieces of `burst` consecutive k-blocks are neighbours in every row
                // of the map, and asking for them back to back lets DRAM serve them from one open page
                for (int kb0 = 0; kb0 < a.kblocks; kb0 += a.burst) {
                    const int nb = min(a.burst, a.kblocks - kb0);
                    for (int i = 0; i < nb; ++i) {
                        const int s = (g + i) % a.stages, n = (g + i) / a.stages;
                        if (n > 0) pj_mbar_wait(empty0 + 8 * s, (n - 1) & 1);
                    }
                    for (int i = 0; i < nb; ++i, ++g) {
                        const int s = g % a.stages;
                        const uint32_t st = ring + s * a.stage_bytes, bar = full0 + 8 * s;
                        pj_mbar_expect_tx(bar, bytes);
                        pj_tma_load_3d(st, &ma, (kb0 + i) * kBK, m0, bv, bar);
                        pj_tma_load_3d(st + off_b, &mb, (kb0 + i) * kBK, v, 0, bar);
                        if (split) pj_tma_load_3d(st + off_blo, &mbl, (kb0 + i) * kBK, v, 0, bar);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        int g = 0, t = 0;
        for (int tile = blockIdx.x; tile < a.tiles; tile += gridDim.x, ++t) {
            const int q = t & 1, use = t >> 1;                       // accumulator buffer and how often it was used before
            if (use > 0) pj_mbar_wait(acce0 + 8 * q, (use - 1) & 1);  // the epilogue has read the tile before last out of it
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t acc = tmem + (uint32_t)(q * a.acc_cols);
            for (int kb = 0; kb < a.kblocks; ++kb, ++g) {
                const int s = g % a.stages, n = g / a.stages;
                pj_mbar_wait((split ? ready0 : full0) + 8 * s, n & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (pj_elect()) {
                    const uint32_t st = ring + s * a.stage_bytes;
                    const uint64_t da = pj_desc(st), dal = pj_desc(st + off_alo), db = pj_desc(st + off_b), dbl = pj_desc(st + off_blo);
#pragma unroll
                    for (int k = 0; k < kBK / kUK; ++k) {
                        const uint64_t adv = (uint64_t)(k * kUK * 4 >> 4);  // 32 bytes along K inside the swizzle row
                        pj_mma_tf32(acc, da + adv, db + adv, a.idesc, (kb | k) != 0);
                        if (split) {
                            pj_mma_tf32(acc, dal + adv, db + adv, a.idesc, 1u);
                            pj_mma_tf32(acc, da + adv, dbl + adv, a.idesc, 1u);
                        }
                    }
                    pj_commit(empty0 + 8 * s);                            // the stage is free when these MMAs have read it
                    if (kb == a.kblocks - 1) pj_commit(accf0 + 8 * q);    // the accumulator is complete
                }
                __syncwarp();
            }
        }
    } else if (warp >= 6) {
        // ===== A-tile converters (fp32-grade mode): head in place, remainder beside it =====
        if (split) {
            const int ct = tid - 6 * 32;  // 0 .. 63
            int g = 0;
            for (int tile = blockIdx.x; tile < a.tiles; tile += gridDim.x) {
                for (int kb = 0; kb < a.kblocks; ++kb, ++g) {
                    const int s = g % a.stages, n = g / a.stages;
                    pj_mbar_wait(full0 + 8 * s, n & 1);
                    unsigned char* st = smem_pj + (ring - smem_u32(smem_pj)) + (size_t)s * a.stage_bytes;
                    float4* hi = reinterpret_cast<float4*>(st);
                    float4* lo = reinterpret_cast<float4*>(st + off_alo);
#pragma unroll 4
                    for (int e = ct; e < kBM * 128 / 16; e += 64) {  // element-wise: the swizzled order does not matter
                        const float4 x = hi[e];
                        float4 h, l;
                        h.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u); l.x = __fsub_rn(x.x, h.x);
                        h.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u); l.y = __fsub_rn(x.y, h.y);
                        h.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u); l.z = __fsub_rn(x.z, h.z);
                        h.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u); l.w = __fsub_rn(x.w, h.w);
                        hi[e] = h;
                        lo[e] = l;
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> the tensor core's reads
                    pj_mbar_arrive(ready0 + 8 * s);
                }
            }
        }
    } else {
        // ===== epilogue: TMEM -> registers -> swizzled staging behind the ring -> TMA store =====
        const int q4 = warp & 3;          // the TMEM lane quarter this warp may read
        const int row = q4 * 32 + lane;   // row of the tile
        const uint32_t stage_out = ring + (uint32_t)a.out_off;
        int t = 0;
        for (int tile = blockIdx.x; tile < a.tiles; tile += gridDim.x, ++t) {
            const int bv = tile / a.tiles_m, m0 = (tile - bv * a.tiles_m) * kBM;
            const int q = t & 1, use = t >> 1;
            pj_mbar_wait(accf0 + 8 * q, use & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // the tile leaves in rounds of out_sub 32-column groups (the staging area holds that many [128 x 128 byte] sub-tiles)
            for (int j0 = 0; j0 < a.Co; j0 += 32 * a.out_sub) {
                const int j1 = min(a.Co, j0 + 32 * a.out_sub);
                // the staging area is free again when the previous round's stores have read it
                if (t > 0 || j0 > 0) {
                    if (warp == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                }
                for (int j = j0; j < j1; j += 32) {
                    uint32_t r[32];
                    const uint32_t taddr = tmem + (uint32_t)(q * a.acc_cols) + ((uint32_t)(q4 * 32) << 16) + (uint32_t)j;
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
                        "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
                          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
                          "=r"(r[31])
                        : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    // sub-tile: [128 rows][32 columns = 128 bytes], 16-byte piece c of row `row` at piece c ^ (row & 7)
                    const uint32_t sub = stage_out + (uint32_t)((j - j0) / 32) * (kBM * 128) + (uint32_t)row * 128;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const uint32_t dst = sub + (uint32_t)((c ^ (row & 7)) << 4);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(r[4 * c]), "r"(r[4 * c + 1]), "r"(r[4 * c + 2]), "r"(r[4 * c + 3]) : "memory");
                    }
                }
                if (j1 == a.Co) {
                    // this thread is done with the accumulator: the MMA warp may start the tile after next in it
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    pj_mbar_arrive(acce0 + 8 * q);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("bar.sync 1, 128;" ::: "memory");  // the four epilogue warps: the round is staged
                if (warp == 2 && lane == 0) {
                    for (int j = j0; j < j1; j += 32) pj_tma_store_3d(&md, stage_out + (uint32_t)((j - j0) / 32) * (kBM * 128), j, m0, bv);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
        }
        if (warp == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the staging must outlive the copies' reads
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)a.tmem_cols) : "memory");
    }
}

// ---- Co == 128: the transposed form ------------------------------------------------------------------------------------
// D^T[o, r] = sum_c W[o, c] x[r, c]: the weights are the M = 128 operand, 256 rows of the map the N = 256 operand.  A
// tcgen05.mma.kind::tf32 of shape m128 n128 k8 reads 8 KB of operands for 0.26 MFLOP and takes ~128 cycles (measured: tensor
// pipe 54 % busy in the 128-row kernel): it is bound by its shared-memory operand reads.  m128 n256 k8 does twice the work in
// the same time with 12 KB, and the weight tile is loaded once per 256 rows instead of once per 128.
// The accumulator comes out transposed (TMEM lane = output channel, column = row of the map): every epilogue warp owns 32
// channels = one 128-byte column group of the result, transposes through its own staging sub-tile with conflict-free 4-byte
// stores and sends it off with its own TMA store -- no barrier between the epilogue warps.
// Work is handed out in units of 128 rows (a CTA takes a contiguous range of units and pairs them up; a range's odd unit or a
// map's last unit runs as a single, N = 128), so the grid stays balanced when the number of 256-row tiles is just over a
// multiple of the SM count (7 x 127 = 889 = 6 x 148 + 1 at wildtrack.yaml).
struct ProjTArgs {
    int kblocks, tiles_m, units, V;
    int stages, stage_bytes, passes;
    int sr;          // rows per epilogue round (each epilogue warp stages [sr rows x 32 channels] at a time)
    int out_off;     // byte offset of the staging area behind the ring
    uint32_t idesc1, idesc2;  // N = 128 (single unit) / N = 256 (pair)
};

constexpr int kTThreads = 320;  // warp 0 TMA, 1 MMA, 2-5 epilogue, 6-9 converters (split mode)

// mx2 / mx1: x as [C, rows, B*V] with boxes of 256 / 128 rows; mw / mwl: W_hi / W_lo as [C, V, 128]; md: out as [128, rows, B*V], box [32, sr]
__global__ void __launch_bounds__(kTThreads, 1) proj1x1_t_kernel(const __grid_constant__ CUtensorMap mx2, const __grid_constant__ CUtensorMap mx1,
                                                                 const __grid_constant__ CUtensorMap mw, const __grid_constant__ CUtensorMap mwl,
                                                                 const __grid_constant__ CUtensorMap md, const ProjTArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_pj[];
    __shared__ __align__(8) unsigned long long bars[3 * 8 + 4];  // full[8], ready[8], empty[8]; accumulator full[2] / empty[2]
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t ring = (smem_u32(smem_pj) + 1023u) & ~1023u;
    const uint32_t full0 = smem_u32(&bars[0]), ready0 = smem_u32(&bars[8]), empty0 = smem_u32(&bars[16]);
    const uint32_t accf0 = smem_u32(&bars[24]), acce0 = smem_u32(&bars[26]);
    const bool split = a.passes == 3;
    // stage: x (256 rows x 128 B) [, x remainders], W (128 x 128 B) [, W remainders]
    const int x_bytes = 256 * 128, w_bytes = 128 * 128;
    const int off_xlo = x_bytes, off_w = split ? 2 * x_bytes : x_bytes, off_wlo = off_w + w_bytes;
    const int u0 = (int)((long long)blockIdx.x * a.units / gridDim.x), u1 = (int)((long long)(blockIdx.x + 1) * a.units / gridDim.x);
    // the units of this CTA, as pairs where two neighbours lie in the same map: every role walks the same sequence
    auto halves_at = [&](int u) { return ((u % a.tiles_m) + 1 < a.tiles_m && u + 1 < u1) ? 2 : 1; };

    if (tid == 0) {
        for (int s = 0; s < a.stages; ++s) {
            pj_mbar_init(full0 + 8 * s, 1);
            pj_mbar_init(ready0 + 8 * s, 128);  // every lane of the four converter warps
            pj_mbar_init(empty0 + 8 * s, 1);
        }
        for (int q = 0; q < 2; ++q) {
            pj_mbar_init(accf0 + 8 * q, 1);
            pj_mbar_init(acce0 + 8 * q, 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;

    if (warp == 0) {
        // ===== TMA producer =====
        if (pj_elect()) {
            int g = 0;
            for (int u = u0; u < u1;) {
                const int halves = halves_at(u);
                const int bv = u / a.tiles_m, m0 = (u - bv * a.tiles_m) * kBM, v = bv % a.V;
                const int bytes = halves * (kBM * 128) + (split ? 2 : 1) * w_bytes;
                for (int kb = 0; kb < a.kblocks; ++kb, ++g) {
                    const int s = g % a.stages, n = g / a.stages;
                    if (n > 0) pj_mbar_wait(empty0 + 8 * s, (n - 1) & 1);
                    const uint32_t st = ring + s * a.stage_bytes, bar = full0 + 8 * s;
                    pj_mbar_expect_tx(bar, bytes);
                    pj_tma_load_3d(st, halves == 2 ? &mx2 : &mx1, kb * kBK, m0, bv, bar);
                    pj_tma_load_3d(st + off_w, &mw, kb * kBK, v, 0, bar);
                    if (split) pj_tma_load_3d(st + off_wlo, &mwl, kb * kBK, v, 0, bar);
                }
                u += halves;
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        int g = 0, t = 0;
        for (int u = u0; u < u1; ++t) {
            const int halves = halves_at(u);
            u += halves;
            const int q = t & 1, use = t >> 1;
            if (use > 0) pj_mbar_wait(acce0 + 8 * q, (use - 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t acc = tmem + (uint32_t)(q * 256);
            const uint32_t idesc = halves == 2 ? a.idesc2 : a.idesc1;
            for (int kb = 0; kb < a.kblocks; ++kb, ++g) {
                const int s = g % a.stages, n = g / a.stages;
                pj_mbar_wait((split ? ready0 : full0) + 8 * s, n & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (pj_elect()) {
                    const uint32_t st = ring + s * a.stage_bytes;
                    const uint64_t dx = pj_desc(st), dxl = pj_desc(st + off_xlo), dw = pj_desc(st + off_w), dwl = pj_desc(st + off_wlo);
#pragma unroll
                    for (int k = 0; k < kBK / kUK; ++k) {
                        const uint64_t adv = (uint64_t)(k * kUK * 4 >> 4);
                        pj_mma_tf32(acc, dw + adv, dx + adv, idesc, (kb | k) != 0);
                        if (split) {
                            pj_mma_tf32(acc, dwl + adv, dx + adv, idesc, 1u);
                            pj_mma_tf32(acc, dw + adv, dxl + adv, idesc, 1u);
                        }
                    }
                    pj_commit(empty0 + 8 * s);
                    if (kb == a.kblocks - 1) pj_commit(accf0 + 8 * q);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 6) {
        // ===== x-tile converters (fp32-grade mode): head in place, remainder beside it =====
        if (split) {
            const int ct = tid - 6 * 32;  // 0 .. 127
            int g = 0;
            for (int u = u0; u < u1;) {
                const int halves = halves_at(u);
                u += halves;
                for (int kb = 0; kb < a.kblocks; ++kb, ++g) {
                    const int s = g % a.stages, n = g / a.stages;
                    pj_mbar_wait(full0 + 8 * s, n & 1);
                    unsigned char* st = smem_pj + (ring - smem_u32(smem_pj)) + (size_t)s * a.stage_bytes;
                    float4* hi = reinterpret_cast<float4*>(st);
                    float4* lo = reinterpret_cast<float4*>(st + off_xlo);
#pragma unroll 4
                    for (int e = ct; e < halves * (kBM * 128 / 16); e += 128) {
                        const float4 x = hi[e];
                        float4 h, l;
                        h.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u); l.x = __fsub_rn(x.x, h.x);
                        h.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u); l.y = __fsub_rn(x.y, h.y);
                        h.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u); l.z = __fsub_rn(x.z, h.z);
                        h.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u); l.w = __fsub_rn(x.w, h.w);
                        hi[e] = h;
                        lo[e] = l;
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    pj_mbar_arrive(ready0 + 8 * s);
                }
            }
        }
    } else {
        // ===== epilogue: this warp's 32 output channels (TMEM lanes) of every row (TMEM column) =====
        const int q4 = warp & 3;
        const uint32_t my_sub = ring + (uint32_t)a.out_off + (uint32_t)q4 * (uint32_t)(a.sr * 128);  // [sr rows][32 channels = 128 bytes], swizzled
        const uint32_t lane_word = (uint32_t)(lane & 3) * 4u, lane_piece = (uint32_t)(lane >> 2);
        int t = 0;
        bool stored = false;
        for (int u = u0; u < u1; ++t) {
            const int halves = halves_at(u);
            const int bv = u / a.tiles_m, m0 = (u - bv * a.tiles_m) * kBM;
            u += halves;
            const int q = t & 1, use = t >> 1;
            pj_mbar_wait(accf0 + 8 * q, use & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int rows_t = halves * kBM;
            for (int r0 = 0; r0 < rows_t; r0 += a.sr) {
                if (stored) {  // my staging sub-tile is free again when my previous store has read it
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    __syncwarp();
                }
                for (int jj = 0; jj < a.sr; jj += 32) {
                    uint32_t r[32];
                    const uint32_t taddr = tmem + (uint32_t)(q * 256) + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(r0 + jj);
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
                        "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
                          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
                          "=r"(r[31])
                        : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    // r[i] = result of (row r0 + jj + i, channel 32 q4 + lane): word (lane & 3) of 16-byte piece (lane >> 2) ^ (row & 7)
                    const uint32_t rowbase = my_sub + (uint32_t)jj * 128u + lane_word;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const uint32_t dst = rowbase + (uint32_t)i * 128u + ((lane_piece ^ (uint32_t)(i & 7)) << 4);
                        asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst), "r"(r[i]) : "memory");
                    }
                }
                if (r0 + a.sr >= rows_t) {  // the accumulator has been read: the MMA warp may start the tile after next in it
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    pj_mbar_arrive(acce0 + 8 * q);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    pj_tma_store_3d(&md, my_sub, q4 * 32, m0 + r0, bv);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                stored = true;
            }
        }
        if (stored && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

PFN_cuTensorMapEncodeTiled pj_encode_fn() {
    static PFN_cuTensorMapEncodeTiled fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
        return reinterpret_cast<PFN_cuTensorMapEncodeTiled>(f);
    }();
    return fn;
}

int pj_fail(int code, const char* fmt, ...) {
    char buf[256];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    return set_error(code, buf);
}

// a 3-D fp32 map, 128-byte swizzle: dims / strides innermost first (stride of dim 0 is 4 bytes)
int pj_map(CUtensorMap* m, const void* base, cuuint64_t d0, cuuint64_t d1, cuuint64_t d2, cuuint64_t s1_bytes, cuuint64_t s2_bytes,
           cuuint32_t b0, cuuint32_t b1, cuuint32_t b2, const char* what) {
    PFN_cuTensorMapEncodeTiled enc = pj_encode_fn();
    if (!enc) return pj_fail(BEVIPM_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[3] = {d0, d1, d2};
    const cuuint64_t strides[2] = {s1_bytes, s2_bytes};
    const cuuint32_t box[3] = {b0, b1, b2};
    const cuuint32_t ones[3] = {1, 1, 1};
    const CUresult rc = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return pj_fail(BEVIPM_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled failed (%d) for %s", (int)rc, what);
    return 0;
}

}  // namespace
}  // namespace bevipm

extern "C" int bevipm_proj1x1(const float* x, const float* w_hi, const float* w_lo, float* out, int32_t BV, int32_t V, int64_t rows, int32_t C,
                              int32_t Co, int64_t x_row_stride, int64_t x_map_stride, int64_t out_row_stride, int64_t out_map_stride, int32_t passes,
                              void* stream) {
    using namespace bevipm;
    if (!x || !w_hi || !out || BV <= 0 || V <= 0 || rows <= 0 || C <= 0) return pj_fail(BEVIPM_ERR_BAD_ARG, "proj1x1: null pointer or empty extent");
    if (passes != 1 && passes != 3) return pj_fail(BEVIPM_ERR_BAD_ARG, "proj1x1: passes must be 1 (TF32) or 3 (split operands), got %d", passes);
    if (passes == 3 && !w_lo) return pj_fail(BEVIPM_ERR_BAD_ARG, "proj1x1: the split mode needs the weight remainders (w_lo)");
    if (Co < 16 || Co > 256 || Co % 16) return pj_fail(BEVIPM_ERR_UNSUPPORTED, "proj1x1: Co = %d (a multiple of 16 in [16, 256] is required)", Co);
    if (C % 4 || x_row_stride % 4 || x_map_stride % 4 || x_row_stride < C || out_row_stride % 4 || out_map_stride % 4 || out_row_stride < Co)
        return pj_fail(BEVIPM_ERR_UNSUPPORTED, "proj1x1: C and the strides of x / out must be multiples of 4 elements (16-byte rule of the copy engine)");
    if (((uintptr_t)x | (uintptr_t)w_hi | (uintptr_t)w_lo | (uintptr_t)out) & 15) return pj_fail(BEVIPM_ERR_BAD_ARG, "proj1x1: pointers must be 16-byte aligned");
    if (rows > (1ll << 31) - 256) return pj_fail(BEVIPM_ERR_UNSUPPORTED, "proj1x1: map too large");
    int rc;
    const cuuint64_t big = 1ull << 30;  // legal stride for an extent-1 outer dimension
    if (Co == 128 && !getenv("BEVIPM_PJ_GENERAL")) {
        // the transposed form (see proj1x1_t_kernel): weights as the M operand, 256 rows of the map as N
        CUtensorMap mx2, mx1, mw, mwl, mo;
        ProjTArgs t;
        t.passes = passes;
        t.stage_bytes = (passes == 3 ? 2 : 1) * (256 * 128 + 128 * 128);
        t.sr = 64;  // rows per epilogue round: a 32 KB staging area leaves four 48 KB stages (one pass) / two 96 KB stages (split)
        if (const char* e = getenv("BEVIPM_PJ_SR")) t.sr = atoi(e) == 64 ? 64 : (atoi(e) == 32 ? 32 : 128);
        const int stage_out = 4 * t.sr * 128;
        t.stages = std::min((227 * 1024 - 3 * 1024 - stage_out) / t.stage_bytes, 8);
        if (t.stages < 2) return pj_fail(BEVIPM_ERR_UNSUPPORTED, "proj1x1: no room for two stages");
        t.out_off = t.stages * t.stage_bytes;
        if ((rc = pj_map(&mx2, x, (cuuint64_t)C, (cuuint64_t)rows, (cuuint64_t)BV, (cuuint64_t)x_row_stride * 4, BV > 1 ? (cuuint64_t)x_map_stride * 4 : big * 16,
                         kBK, 256, 1, "x (256 rows)")))
            return rc;
        if ((rc = pj_map(&mx1, x, (cuuint64_t)C, (cuuint64_t)rows, (cuuint64_t)BV, (cuuint64_t)x_row_stride * 4, BV > 1 ? (cuuint64_t)x_map_stride * 4 : big * 16,
                         kBK, 128, 1, "x (128 rows)")))
            return rc;
        if ((rc = pj_map(&mw, w_hi, (cuuint64_t)C, (cuuint64_t)V, 128, (cuuint64_t)C * 4, (cuuint64_t)V * C * 4, kBK, 1, 128, "w_hi"))) return rc;
        if ((rc = pj_map(&mwl, passes == 3 ? w_lo : w_hi, (cuuint64_t)C, (cuuint64_t)V, 128, (cuuint64_t)C * 4, (cuuint64_t)V * C * 4, kBK, 1, 128, "w_lo"))) return rc;
        if ((rc = pj_map(&mo, out, 128, (cuuint64_t)rows, (cuuint64_t)BV, (cuuint64_t)out_row_stride * 4, BV > 1 ? (cuuint64_t)out_map_stride * 4 : big * 16, 32,
                         (cuuint32_t)t.sr, 1, "out")))
            return rc;
        t.kblocks = (C + kBK - 1) / kBK;
        t.tiles_m = (int)((rows + kBM - 1) / kBM);
        const long long units = (long long)BV * t.tiles_m;
        if (units > 0x7fffffffLL) return pj_fail(BEVIPM_ERR_UNSUPPORTED, "proj1x1: too many tiles");
        t.units = (int)units;
        t.V = V;
        t.idesc1 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        t.idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const int smem = t.out_off + stage_out + 1024;
        if (cudaFuncSetAttribute(proj1x1_t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
            return pj_fail(BEVIPM_ERR_CUDA, "proj1x1: cudaFuncSetAttribute(%d bytes)", smem);
        int dev = 0, sms = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const long long grid = std::min<long long>((units + 1) / 2, sms);
        proj1x1_t_kernel<<<(unsigned)grid, kTThreads, smem, (cudaStream_t)stream>>>(mx2, mx1, mw, mwl, mo, t);
        if (cudaGetLastError() != cudaSuccess) return pj_fail(BEVIPM_ERR_CUDA, "proj1x1: kernel launch failed");
        note_launch(71);
        return 0;
    }
    CUtensorMap ma, mb, mbl, md;
    if ((rc = pj_map(&ma, x, (cuuint64_t)C, (cuuint64_t)rows, (cuuint64_t)BV, (cuuint64_t)x_row_stride * 4, BV > 1 ? (cuuint64_t)x_map_stride * 4 : big * 16,
                     kBK, kBM, 1, "x")))
        return rc;
    if ((rc = pj_map(&mb, w_hi, (cuuint64_t)C, (cuuint64_t)V, (cuuint64_t)Co, (cuuint64_t)C * 4, (cuuint64_t)V * C * 4, kBK, 1, (cuuint32_t)Co, "w_hi"))) return rc;
    if ((rc = pj_map(&mbl, passes == 3 ? w_lo : w_hi, (cuuint64_t)C, (cuuint64_t)V, (cuuint64_t)Co, (cuuint64_t)C * 4, (cuuint64_t)V * C * 4, kBK, 1,
                     (cuuint32_t)Co, "w_lo")))
        return rc;
    if ((rc = pj_map(&md, out, (cuuint64_t)Co, (cuuint64_t)rows, (cuuint64_t)BV, (cuuint64_t)out_row_stride * 4, BV > 1 ? (cuuint64_t)out_map_stride * 4 : big * 16, 32, kBM, 1,
                     "out")))
        return rc;
    ProjArgs a;
    a.kblocks = (C + kBK - 1) / kBK;
    a.tiles_m = (int)((rows + kBM - 1) / kBM);
    const long long tiles = (long long)BV * a.tiles_m;
    if (tiles > 0x7fffffffLL) return pj_fail(BEVIPM_ERR_UNSUPPORTED, "proj1x1: too many tiles");
    a.tiles = (int)tiles;
    a.V = V;
    a.Co = Co;
    a.passes = passes;
    a.stage_bytes = (passes == 3 ? 2 : 1) * (kBM * 128 + Co * 128);
    // one CTA per SM: the ring takes what the epilogue's staging area leaves of the 227 KB; the staging area shrinks (the tile
    // then leaves in more rounds) until at least three stages fit, or two if that is all there is
    int out_bytes = 0;
    auto fit = [&](int want) {
        for (int sub = std::min((Co + 31) / 32, 4); sub >= 1; sub /= 2) {
            const int ob = sub * kBM * 128;
            const int budget = 227 * 1024 - 1024 /* driver */ - 1024 /* alignment slack */ - 1024 /* barriers */ - ob;
            const int stages = std::min(budget / a.stage_bytes, 8);
            if (stages >= want) {
                a.out_sub = sub; a.stages = stages; out_bytes = ob;
                return true;
            }
        }
        return false;
    };
    if (!fit(3) && !fit(2)) return pj_fail(BEVIPM_ERR_UNSUPPORTED, "proj1x1: Co = %d does not leave room for two stages", Co);
    a.burst = 1;
    if (const char* e = getenv("BEVIPM_PJ_BURST")) a.burst = std::max(1, atoi(e));
    a.burst = std::min(a.burst, a.stages - 1);
    a.out_off = a.stages * a.stage_bytes;
    a.acc_cols = (Co + 31) / 32 * 32;      // (the epilogue reads whole groups of 32 columns)
    a.tmem_cols = 32;
    while (a.tmem_cols < 2 * a.acc_cols) a.tmem_cols *= 2;
    // instruction descriptor: D fp32 (1 << 4), A and B TF32 (2 << 7, 2 << 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
    a.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(Co >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
    const int smem = a.out_off + out_bytes + 1024;  // + alignment slack
    if (cudaFuncSetAttribute(proj1x1_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
        return pj_fail(BEVIPM_ERR_CUDA, "proj1x1: cudaFuncSetAttribute(%d bytes)", smem);
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long grid = tiles < sms ? tiles : sms;
    proj1x1_tf32_kernel<<<(unsigned)grid, kThreads, smem, (cudaStream_t)stream>>>(ma, mb, mbl, md, a);
    if (cudaGetLastError() != cudaSuccess) return pj_fail(BEVIPM_ERR_CUDA, "proj1x1: kernel launch failed");
    note_launch(70);
    return 0;
}
