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
#include <cstdarg>
#include <cstdio>
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
    int V;           // views (the weight of view bv % V)
    int Co;          // output channels = UMMA N (multiple of 16, <= 256)
    int stages;      // ring depth
    int stage_bytes; // bytes of one stage (A [+ A_lo] + B [+ B_lo])
    int passes;      // 1: TF32, 3: split operands (fp32-grade)
    int tmem_cols;   // power of two >= max(32, Co)
    uint32_t idesc;  // tcgen05 instruction descriptor
};

// ma: x as [C, rows, B*V]; mb / mbl: W_hi / W_lo as [C, V, Co]; md: out as [Co, rows, B*V] (all fp32, 128-byte swizzle)
__global__ void __launch_bounds__(kThreads, 2) proj1x1_tf32_kernel(const __grid_constant__ CUtensorMap ma, const __grid_constant__ CUtensorMap mb,
                                                                const __grid_constant__ CUtensorMap mbl, const __grid_constant__ CUtensorMap md,
                                                                const ProjArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_pj[];
    __shared__ __align__(8) unsigned long long bars[3 * 8 + 1];  // full[8], ready[8] (A tile split), empty[8], accumulator done
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t ring = (smem_u32(smem_pj) + 1023u) & ~1023u;  // the swizzle pattern repeats every 1024 bytes
    const uint32_t full0 = smem_u32(&bars[0]), ready0 = smem_u32(&bars[8]), empty0 = smem_u32(&bars[16]), accb = smem_u32(&bars[24]);
    const int bv = blockIdx.x / a.tiles_m, mt = blockIdx.x - bv * a.tiles_m;
    const int m0 = mt * kBM, v = bv % a.V;
    const bool split = a.passes == 3;
    const int a_bytes = kBM * 128, b_bytes = a.Co * 128;
    const int off_alo = a_bytes, off_b = split ? 2 * a_bytes : a_bytes, off_blo = off_b + b_bytes;

    if (tid == 0) {
        for (int s = 0; s < a.stages; ++s) {
            pj_mbar_init(full0 + 8 * s, 1);
            pj_mbar_init(ready0 + 8 * s, 64);  // every lane of the two converter warps
            pj_mbar_init(empty0 + 8 * s, 1);
        }
        pj_mbar_init(accb, 1);
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
            for (int kb = 0; kb < a.kblocks; ++kb) {
                const int s = kb % a.stages, n = kb / a.stages;
                if (n > 0) pj_mbar_wait(empty0 + 8 * s, (n - 1) & 1);
                const uint32_t st = ring + s * a.stage_bytes, bar = full0 + 8 * s;
                pj_mbar_expect_tx(bar, bytes);
                pj_tma_load_3d(st, &ma, kb * kBK, m0, bv, bar);
                pj_tma_load_3d(st + off_b, &mb, kb * kBK, v, 0, bar);
                if (split) pj_tma_load_3d(st + off_blo, &mbl, kb * kBK, v, 0, bar);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        for (int kb = 0; kb < a.kblocks; ++kb) {
            const int s = kb % a.stages, n = kb / a.stages;
            pj_mbar_wait((split ? ready0 : full0) + 8 * s, n & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (pj_elect()) {
                const uint32_t st = ring + s * a.stage_bytes;
                const uint64_t da = pj_desc(st), dal = pj_desc(st + off_alo), db = pj_desc(st + off_b), dbl = pj_desc(st + off_blo);
#pragma unroll
                for (int k = 0; k < kBK / kUK; ++k) {
                    const uint64_t adv = (uint64_t)(k * kUK * 4 >> 4);  // 32 bytes along K inside the swizzle row
                    pj_mma_tf32(tmem, da + adv, db + adv, a.idesc, (kb | k) != 0);
                    if (split) {
                        pj_mma_tf32(tmem, dal + adv, db + adv, a.idesc, 1u);
                        pj_mma_tf32(tmem, da + adv, dbl + adv, a.idesc, 1u);
                    }
                }
                pj_commit(empty0 + 8 * s);                      // the stage is free when these MMAs have read it
                if (kb == a.kblocks - 1) pj_commit(accb);       // the accumulator is complete
            }
            __syncwarp();
        }
    } else if (warp >= 6) {
        // ===== A-tile converters (fp32-grade mode): head in place, remainder beside it =====
        if (split) {
            const int ct = tid - 6 * 32;  // 0 .. 63
            for (int kb = 0; kb < a.kblocks; ++kb) {
                const int s = kb % a.stages, n = kb / a.stages;
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
    } else {
        // ===== epilogue: TMEM -> registers -> swizzled staging in the idle ring -> TMA store =====
        const int q = warp & 3;          // the TMEM lane quarter this warp may read
        const int row = q * 32 + lane;   // row of the tile
        pj_mbar_wait(accb, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int j = 0; j < a.Co; j += 32) {
            uint32_t r[32];
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)j;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
                "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                  "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
                  "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
                  "=r"(r[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            // sub-tile j/32: [128 rows][32 columns = 128 bytes], 16-byte piece c of row `row` at piece c ^ (row & 7)
            const uint32_t sub = ring + (uint32_t)(j / 32) * (kBM * 128) + (uint32_t)row * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint32_t dst = sub + (uint32_t)((c ^ (row & 7)) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(r[4 * c]), "r"(r[4 * c + 1]), "r"(r[4 * c + 2]), "r"(r[4 * c + 3]) : "memory");
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");  // the four epilogue warps
        if (warp == 2 && pj_elect()) {
            for (int j = 0; j < a.Co; j += 32) pj_tma_store_3d(&md, ring + (uint32_t)(j / 32) * (kBM * 128), j, m0, bv);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the staging must outlive the copies' reads
        }
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)a.tmem_cols) : "memory");
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
    CUtensorMap ma, mb, mbl, md;
    int rc;
    const cuuint64_t big = 1ull << 30;  // legal stride for an extent-1 outer dimension
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
    a.V = V;
    a.Co = Co;
    a.passes = passes;
    a.stage_bytes = (passes == 3 ? 2 : 1) * (kBM * 128 + Co * 128);
    const int out_bytes = ((Co + 31) / 32) * kBM * 128;
    // two CTAs per SM when they fit (the second one hides the first one's pipeline fill and epilogue)
    const int budget2 = (227 * 1024 - 2 * 1024) / 2 - 2048, budget1 = 227 * 1024 - 1024 - 2048;
    int budget = (2 * a.stage_bytes <= budget2 && out_bytes <= budget2 && Co <= 256) ? budget2 : budget1;
    a.stages = budget / a.stage_bytes;
    if (a.stages > 8) a.stages = 8;
    if (a.stages < 2) return pj_fail(BEVIPM_ERR_UNSUPPORTED, "proj1x1: Co = %d does not leave room for two stages", Co);
    a.tmem_cols = 32;
    while (a.tmem_cols < Co) a.tmem_cols *= 2;
    // instruction descriptor: D fp32 (1 << 4), A and B TF32 (2 << 7, 2 << 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
    a.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(Co >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
    int smem = a.stages * a.stage_bytes;
    if (smem < out_bytes) smem = out_bytes;
    smem += 1024;  // alignment slack
    if (cudaFuncSetAttribute(proj1x1_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
        return pj_fail(BEVIPM_ERR_CUDA, "proj1x1: cudaFuncSetAttribute(%d bytes)", smem);
    const long long grid = (long long)BV * a.tiles_m;
    if (grid > 0x7fffffffLL) return pj_fail(BEVIPM_ERR_UNSUPPORTED, "proj1x1: grid too large");
    proj1x1_tf32_kernel<<<(unsigned)grid, kThreads, smem, (cudaStream_t)stream>>>(ma, mb, mbl, md, a);
    if (cudaGetLastError() != cudaSuccess) return pj_fail(BEVIPM_ERR_CUDA, "proj1x1: kernel launch failed");
    note_launch(70);
    return 0;
}
