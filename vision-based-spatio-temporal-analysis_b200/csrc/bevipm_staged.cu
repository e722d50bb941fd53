// bevipm_staged.cu -- host side of the TMA-staged fused kernel: tensor maps over the caller's feature tensor
// (encoded through the driver entry point, cached per thread while the tensor stays the same), ring sizing, launch.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <cuda.h>
#include <cudaTypedefs.h>

#include "../../include/bevipm.h"
#include "ipm_staged.cuh"
#include "ipm_boxrun.cuh"
#include "staged_api.h"

namespace bevipm {
bool staged_supported(const FwdParams& p, bool in_bf16);
namespace {

int st_fail(char* err, size_t n, int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    if (err && n) vsnprintf(err, n, fmt, ap);
    va_end(ap);
    return code;
}

PFN_cuTensorMapEncodeTiled encode_fn() {
    static PFN_cuTensorMapEncodeTiled fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return reinterpret_cast<PFN_cuTensorMapEncodeTiled>(f);
    }();
    return fn;
}

// the maps of one feature tensor: thread-local, rebuilt when pointer, extents, strides or type change
struct MapKey {
    const void* feats;
    int C, Wf, Hf, V, B, bf16;
    long long fs_x, fs_y, fs_v, fs_b;
    bool operator==(const MapKey& o) const { return memcmp(this, &o, sizeof(MapKey)) == 0; }
};
thread_local MapKey g_key;
thread_local StagedMaps g_maps;
thread_local bool g_have = false;

int build_maps(const FwdParams& p, bool bf16, char* err, size_t errlen) {
    MapKey k;
    memset(&k, 0, sizeof(k));
    k.feats = p.feats; k.C = p.C; k.Wf = p.Wf; k.Hf = p.Hf; k.V = p.V; k.B = p.B; k.bf16 = bf16;
    k.fs_x = p.fs_x; k.fs_y = p.fs_y; k.fs_v = p.fs_v; k.fs_b = p.fs_b;
    if (g_have && g_key == k) return 0;
    g_have = false;
    PFN_cuTensorMapEncodeTiled enc = encode_fn();
    if (!enc) return st_fail(err, errlen, BEVIPM_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled is not available from this driver");
    const size_t es = bf16 ? 2 : 4;
    const cuuint32_t chunk = bf16 ? 256 : 128;  // one 512-byte channel chunk
    // dims innermost first: channels, x, y, view, frame.  A dimension of extent 1 still needs a legal stride.
    const cuuint64_t dims[5] = {(cuuint64_t)p.C, (cuuint64_t)p.Wf, (cuuint64_t)p.Hf, (cuuint64_t)p.V, (cuuint64_t)p.B};
    const long long view_bytes = (long long)p.fs_v * (long long)es, frame_bytes = (long long)p.fs_b * (long long)es;
    const cuuint64_t strides[4] = {(cuuint64_t)(p.fs_x * (long long)es), (cuuint64_t)(p.fs_y * (long long)es),
                                   (cuuint64_t)(p.V > 1 ? view_bytes : (long long)p.fs_y * (long long)es * p.Hf),
                                   (cuuint64_t)(p.B > 1 ? frame_bytes : (long long)p.fs_y * (long long)es * p.Hf * p.V)};
    const cuuint32_t ones[5] = {1, 1, 1, 1, 1};
    for (int q = 0; q < kStNumMaps; ++q) {
        const cuuint32_t box[5] = {chunk, (cuuint32_t)(q == kStBlockMap ? 2 : st_width(q)), (cuuint32_t)(q == kStBlockMap ? 2 : 1), 1, 1};
        const CUresult rc = enc(&g_maps.m[q], bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5,
                                const_cast<void*>(p.feats), dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS)
            return st_fail(err, errlen, BEVIPM_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled failed (%d) for box width %u", (int)rc, box[1]);
    }
    g_key = k;
    g_have = true;
    return 0;
}

int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e && *e ? atoi(e) : dflt;
}

template <typename TIn, typename TOut, int NW, int KMODE, int PROBE>
int launch_t(FwdParams p, int ctas_per_sm, cudaStream_t st, char* err, size_t errlen) {
    constexpr int VE = VecTraits<TIn>::VE;
    p.tiles_x = (p.Wb + kStCells - 1) / kStCells;
    p.tiles_y = (p.Hb + NW - 1) / NW;
    p.fsy16 = (int)(p.fs_y / VE);
    p.fsx16 = (int)(p.fs_x / VE);
    p.rcpV = 1.0f / (float)p.V;
    // shared memory: the tables, then the stage ring with whatever is left of this CTA's share of the SM (the driver
    // keeps 1 KB per resident CTA); the ring must hold phase A's scratch arrays and two stages of the largest kind
    const StagedSmem L(p.V, NW);
    int ring = ((227 * 1024 - ctas_per_sm * 1024) / ctas_per_sm - L.ring) & ~511;
    ring = env_int("BEVIPM_ST_RING", ring);
    const int scratch = L.scratch_end - L.ring;
    if (ring < scratch) ring = (scratch + 511) & ~511;
    if (ring < 32 * 1024) ring = 32 * 1024;                  // many views: fewer CTAs per SM rather than no kernel
    if (L.ring + ring > 227 * 1024)
        return st_fail(err, errlen, BEVIPM_ERR_UNSUPPORTED, "staged kernel: V=%d needs %d bytes of shared memory", p.V, L.ring + ring);
    int cap = env_int("BEVIPM_ST_CAP", ring / 2);            // largest stage: half the ring, so that the next one can fly
    cap = std::max(16 * 1024, std::min(cap, ring)) & ~511;   // (a BEV row staged block by block takes up to 8 x 2 KB)
    const int lag = std::max(1, std::min(env_int("BEVIPM_ST_LAG", 2), 8));
    const int smem = L.ring + ring;
    auto kern = warp_fuse_staged_kernel<TIn, TOut, NW, 128, KMODE, PROBE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return st_fail(err, errlen, BEVIPM_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    // frames per CTA: phase A is shared by consecutive frames with the same calibration; keep >= 8 CTA waves
    int fpc = 1;
    {
        const long long tiles = (long long)p.tiles_x * p.tiles_y, slots = 148LL * ctas_per_sm;
        while (fpc < 8 && fpc * 2 <= p.B && tiles * ((p.B + fpc * 2 - 1) / (fpc * 2)) >= 8 * slots) fpc *= 2;
        fpc = std::max(1, std::min(env_int("BEVIPM_RUN_FPC", fpc), p.B));
    }
    dim3 grid(p.tiles_x * p.tiles_y, 1, (p.B + fpc - 1) / fpc);
    if (const char* path = getenv("BEVIPM_ST_DUMP")) {
        // development aid: phase A only, the shared-memory tables of every tile go to a file (tools/staged_dump.py reads it)
        const size_t per = (size_t)L.ring, n = (size_t)grid.x * grid.z * per;
        unsigned char* d = nullptr;
        if (cudaMalloc(&d, n) != cudaSuccess) return st_fail(err, errlen, BEVIPM_ERR_CUDA, "dump buffer");
        cudaMemsetAsync(d, 0, n, st);
        kern<<<grid, NW * 32, smem, st>>>(p, fpc, ring, cap, lag, g_maps, d);
        cudaStreamSynchronize(st);
        unsigned char* h = (unsigned char*)malloc(n);
        cudaMemcpy(h, d, n, cudaMemcpyDeviceToHost);
        cudaFree(d);
        if (FILE* f = fopen(path, "wb")) {
            const int hdr[16] = {p.V, NW, ring, cap, L.ring, (int)grid.x, (int)grid.z, p.tiles_x, p.tiles_y, fpc, p.Hb, p.Wb, p.Hf, p.Wf, p.C, (int)sizeof(TIn)};
            fwrite(hdr, sizeof(hdr), 1, f);
            fwrite(h, 1, n, f);
            fclose(f);
        }
        free(h);
        e = cudaGetLastError();
        if (e != cudaSuccess) return st_fail(err, errlen, BEVIPM_ERR_CUDA, "staged kernel (dump): %s", cudaGetErrorString(e));
        return 0;
    }
    kern<<<grid, NW * 32, smem, st>>>(p, fpc, ring, cap, lag, g_maps, nullptr);
    e = cudaGetLastError();
    if (e != cudaSuccess) return st_fail(err, errlen, BEVIPM_ERR_CUDA, "staged kernel launch: %s", cudaGetErrorString(e));
    return 0;
}

template <typename TIn, typename TOut, int NW>
int launch_mode(const FwdParams& p, int cps, int probe, cudaStream_t st, char* err, size_t errlen) {
    if (p.mode == BEVIPM_MAX) return launch_t<TIn, TOut, NW, KM_MAX, 0>(p, cps, st, err, errlen);
    if (probe == 1) return launch_t<TIn, TOut, NW, KM_ACC, 1>(p, cps, st, err, errlen);
    if (probe == 2) return launch_t<TIn, TOut, NW, KM_ACC, 2>(p, cps, st, err, errlen);
    return launch_t<TIn, TOut, NW, KM_ACC, 0>(p, cps, st, err, errlen);
}

// ---- run kernel with the TMA-box ring (ipm_boxrun.cuh) ----------------------------------------------------------------
template <typename TIn, typename TOut, int MAXREG, int KMODE>
int launch_box_t(FwdParams p, cudaStream_t st, char* err, size_t errlen) {
    constexpr int VE = VecTraits<TIn>::VE, CELLS = 8, NW = 4, DEPTH = 4;
    p.tiles_x = (p.Wb + CELLS - 1) / CELLS;
    p.tiles_y = (p.Hb + NW - 1) / NW;
    p.fsy16 = (int)(p.fs_y / VE);
    p.fsx16 = (int)(p.fs_x / VE);
    p.rcpV = 1.0f / (float)p.V;
    auto kern = warp_fuse_boxrun_kernel<TIn, TOut, CELLS, NW, MAXREG, DEPTH, KMODE>;
    const size_t smem = (size_t)run_tables_bytes(p.V, CELLS, NW) + (size_t)NW * DEPTH * 2048 + (size_t)p.V * 48 + (size_t)NW * DEPTH * 8;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return st_fail(err, errlen, BEVIPM_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    int fpc = 1;
    {
        const long long tiles = (long long)p.tiles_x * p.tiles_y, slots = 148LL * (65536 / (MAXREG * 32 * NW));
        while (fpc < 8 && fpc * 2 <= p.B && tiles * ((p.B + fpc * 2 - 1) / (fpc * 2)) >= 16 * slots) fpc *= 2;
        fpc = std::max(1, std::min(env_int("BEVIPM_RUN_FPC", fpc), p.B));
    }
    dim3 grid(p.tiles_x * p.tiles_y, 1, (p.B + fpc - 1) / fpc);
    kern<<<grid, NW * 32, smem, st>>>(p, fpc, g_maps.m[kStBlockMap]);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return st_fail(err, errlen, BEVIPM_ERR_CUDA, "box-ring kernel launch: %s", cudaGetErrorString(e));
    return 0;
}

template <typename TIn, typename TOut, int MAXREG>
int launch_box_mode(const FwdParams& p, cudaStream_t st, char* err, size_t errlen) {
    if (p.mode == BEVIPM_MAX) return launch_box_t<TIn, TOut, MAXREG, KM_MAX>(p, st, err, errlen);
    return launch_box_t<TIn, TOut, MAXREG, KM_ACC>(p, st, err, errlen);
}

}  // namespace

int launch_boxrun(FwdParams p, bool in_bf16, bool out_bf16, cudaStream_t st, char* err, size_t errlen) {
    if (!staged_supported(p, in_bf16)) return st_fail(err, errlen, BEVIPM_ERR_UNSUPPORTED, "box-ring kernel: strides / extents / mode not supported");
    if (int rc = build_maps(p, in_bf16, err, errlen)) return rc;
    if (!in_bf16 && !out_bf16) return launch_box_mode<float, float, 96>(p, st, err, errlen);
    if (in_bf16 && out_bf16) return launch_box_mode<__nv_bfloat16, __nv_bfloat16, 128>(p, st, err, errlen);
    if (in_bf16 && !out_bf16) return launch_box_mode<__nv_bfloat16, float, 128>(p, st, err, errlen);
    return st_fail(err, errlen, BEVIPM_ERR_UNSUPPORTED, "box-ring kernel: fp32 features with bf16 output are not built");
}

bool staged_supported(const FwdParams& p, bool in_bf16) {
    const long long es = in_bf16 ? 2 : 4;
    if (p.mode == BEVIPM_NONE) return false;                       // per-view maps stay on the run kernel (write-bound)
    if (p.V > 32 || p.Wf > 32000 || p.Hf > 32000) return false;    // 16-bit texel coordinates in the tables
    if (p.fs_c != 1) return false;
    const long long s[4] = {p.fs_x * es, p.fs_y * es, p.fs_v * es, p.fs_b * es};
    for (int q = 0; q < 4; ++q) {
        if (q == 2 && p.V == 1) continue;
        if (q == 3 && p.B == 1) continue;
        if (s[q] <= 0 || (s[q] & 15) || s[q] >= (1LL << 40)) return false;  // TMA: positive multiples of 16 bytes
    }
    if ((reinterpret_cast<uintptr_t>(p.feats) & 15) != 0) return false;
    return encode_fn() != nullptr;
}

int launch_staged(FwdParams p, bool in_bf16, bool out_bf16, int shape, int probe, cudaStream_t st, char* err, size_t errlen) {
    if (!staged_supported(p, in_bf16)) return st_fail(err, errlen, BEVIPM_ERR_UNSUPPORTED, "staged kernel: strides / extents / mode not supported");
    if (int rc = build_maps(p, in_bf16, err, errlen)) return rc;
    // 8-row tiles: 2 CTAs per SM; 4-row tiles: 4 CTAs per SM
    const int cps = env_int("BEVIPM_ST_CPS", shape == 1 ? 4 : 2);
    if (cps < 1 || cps > 8) return st_fail(err, errlen, BEVIPM_ERR_BAD_ARG, "staged kernel: %d CTAs per SM", cps);
#define BEVIPM_ST_GO(TI, TO)                                                                       \
    return shape == 1 ? launch_mode<TI, TO, 4>(p, cps, probe, st, err, errlen) : launch_mode<TI, TO, 8>(p, cps, probe, st, err, errlen)
    if (!in_bf16 && !out_bf16) BEVIPM_ST_GO(float, float);
    if (in_bf16 && out_bf16) BEVIPM_ST_GO(__nv_bfloat16, __nv_bfloat16);
    if (in_bf16 && !out_bf16) BEVIPM_ST_GO(__nv_bfloat16, float);
#undef BEVIPM_ST_GO
    return st_fail(err, errlen, BEVIPM_ERR_UNSUPPORTED, "staged kernel: fp32 features with bf16 output are not built");
}

}  // namespace bevipm
