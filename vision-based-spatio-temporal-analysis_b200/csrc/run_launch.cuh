// run_launch.cuh -- host-side launcher of the run kernels (ipm_run.cuh), shared by the translation unit of the default
// instantiations (bevipm_run.cu) and the one of the sweep variants (bevipm_api.cu).
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "../../include/bevipm.h"
#include "ipm_run.cuh"

namespace bevipm {

int set_error(int code, const char* msg);  // bevipm_api.cu: thread-local message behind bevipm_last_error()
void count_launch();                       // bevipm_api.cu: the library's launch counter

// the default run kernels, compiled in bevipm_run.cu: kmode KM_ACC (sum / mean) at 96 or 128 registers, KM_MAX (128), KM_NONE (96)
int launch_run_default(const FwdParams& p, bool in_bf16, bool out_bf16, int kmode, int maxreg, cudaStream_t st);
// the same kernels reading / filling a table cache (p.plan, p.plan_key), compiled in bevipm_run_plan.cu
int launch_run_planned(const FwdParams& p, bool in_bf16, bool out_bf16, int kmode, int maxreg, cudaStream_t st);
// bytes of a table cache for this launch: header + one table set per row segment (rows padded to the CTA's four)
inline size_t run_plan_bytes(int V, int Hb, int Wb) {
    return (size_t)kPlanHeaderBytes + (size_t)((Wb + 7) / 8) * (size_t)((Hb + 3) / 4 * 4) * (size_t)run_seg_bytes(V, 8);
}

// Warps per row segment of the default kernels.  One warp per segment (it walks all channel chunks itself) has the least
// overhead, but its CTAs are long: with fewer than ~6 waves of them the ragged end of the launch costs more than sharing a
// segment's tables between two warps that take every other chunk (measured, fp32 512 ch = 4 chunks: 1 frame 0.0885 -> 0.0816 ms,
// 2 frames 0.0773 -> 0.0758 per frame, 8 frames 0.0691 -> 0.0699, 64 frames 0.0715 -> 0.0722; single-chunk maps lose 45 %).
template <typename TIn>
int pick_ksplit(const FwdParams& p, int maxreg) {
    if (const char* e = getenv("BEVIPM_RUN_KSPLIT")) return atoi(e) == 2 ? 2 : 1;  // development switch
    constexpr int VE = VecTraits<TIn>::VE;
    const int chunks = (p.C + 32 * VE - 1) / (32 * VE);
    if (chunks < 2) return 1;
    const long long tiles = (long long)((p.Wb + 7) / 8) * ((p.Hb + 3) / 4);
    const long long slots = 148LL * (65536 / (maxreg * 128));
    return tiles * p.B < 6 * slots ? 2 : 1;
}

template <typename TIn>
bool run_kernel_ok(const FwdParams& p) {
    constexpr int VE = VecTraits<TIn>::VE;
    if (p.V > kRunMaxViews) return false;
    return (long long)p.V * (p.fs_v / VE) + (long long)(p.Hf + 2) * (p.fs_y / VE) + (long long)(p.Wf + 2) * (p.fs_x / VE) <= 0x7fffffffLL;
}

template <typename TIn, typename TOut, int CELLS, int NW, int KSPLIT, int MAXREG, int DEPTH, bool CA, int PROBE = 0, int KMODE = KM_ACC, bool TMA = false,
          bool HALF = false, bool PLAN = false>
int launch_run(FwdParams p, cudaStream_t st) {
    constexpr int VE = VecTraits<TIn>::VE;
    constexpr int R = NW / KSPLIT;
    char msg[200];
    if (TMA && p.C % (32 * VE)) {
        snprintf(msg, sizeof(msg), "the TMA ring copies whole 512-byte chunks: C must be a multiple of %d", 32 * VE);
        return set_error(BEVIPM_ERR_UNSUPPORTED, msg);
    }
    if (!run_kernel_ok<TIn>(p) || (p.mode == BEVIPM_MAX) != (KMODE == KM_MAX) || (p.mode == BEVIPM_NONE) != (KMODE == KM_NONE)) {
        snprintf(msg, sizeof(msg), "run kernel: needs V <= %d, 32-bit tap offsets, and the variant of the fusion mode", kRunMaxViews);
        return set_error(BEVIPM_ERR_UNSUPPORTED, msg);
    }
    p.tiles_x = (p.Wb + CELLS - 1) / CELLS;
    p.tiles_y = (p.Hb + R - 1) / R;
    p.fsy16 = (int)(p.fs_y / VE);
    p.fsx16 = (int)(p.fs_x / VE);
    p.rcpV = 1.0f / (float)p.V;
    auto kern = warp_fuse_run_kernel<TIn, TOut, CELLS, NW, KSPLIT, MAXREG, DEPTH, CA, PROBE, KMODE, TMA, HALF, PLAN>;
    const size_t smem = (size_t)run_tables_bytes(p.V, CELLS, R) + (size_t)NW * DEPTH * 2048 + (size_t)p.V * 48 + (size_t)NW * DEPTH * 8;  // tables, rings, homographies, ring barriers
    if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return set_error(BEVIPM_ERR_CUDA, cudaGetErrorString(cudaGetLastError()));
    // frames per CTA: the phase A tables are shared by consecutive frames with the same calibration; keep
    // at least ~16 CTA waves so the tail stays small
    int fpc = 1;
    {
        const long long tiles = (long long)p.tiles_x * p.tiles_y;
        const long long slots = 148LL * (65536 / (MAXREG * 32 * NW));
        while (fpc < 8 && fpc * 2 <= p.B && tiles * ((p.B + fpc * 2 - 1) / (fpc * 2)) >= 16 * slots) fpc *= 2;
        if (const char* e = getenv("BEVIPM_RUN_FPC")) fpc = std::max(1, std::min(atoi(e), p.B));
    }
    dim3 grid(p.tiles_x * p.tiles_y, 1, (p.B + fpc - 1) / fpc);
    kern<<<grid, NW * 32, smem, st>>>(p, fpc);
    count_launch();
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(BEVIPM_ERR_CUDA, cudaGetErrorString(e));
    return 0;
}

}  // namespace bevipm
