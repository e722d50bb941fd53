// bevipm_run.cu -- the DEFAULT instantiations of the run kernel (ipm_run.cuh), in a translation unit of their own:
//   sum / mean: fp32 features at 96 registers (variant 32), bf16 features at 128 (variant 33), each with one or two warps
//   per row segment (pick_ksplit);
//   max (fusion.py:22) at 128 registers; per-view maps (geometry.py:162-163, what ConcatFusion reshapes) at 96.
// Everything else of the run-kernel family (ring depth, cells per segment, TMA ring, half reloads, timing probes) is a
// sweep variant compiled in bevipm_api.cu.  Keeping the kernels every default launch runs apart from ~150 sweep
// instantiations makes their code generation independent of what else is being swept (the same source gave 4160 or 4120
// instructions and +-3 % on config 1 depending on its neighbours), and a change to them rebuilds in under a minute.
#include "run_launch.cuh"

// bevipm_run_plan.cu compiles this file a second time with BEVIPM_RUN_PLAN = 1: the same kernels reading / filling a table cache
#ifndef BEVIPM_RUN_PLAN
#define BEVIPM_RUN_PLAN 0
#endif
#if BEVIPM_RUN_PLAN
#define BEVIPM_RUN_ENTRY launch_run_planned
#else
#define BEVIPM_RUN_ENTRY launch_run_default
#endif

namespace bevipm {
namespace {

constexpr bool kPlan = BEVIPM_RUN_PLAN != 0;

template <typename TIn, typename TOut>
int launch_typed(const FwdParams& p, int kmode, int maxreg, cudaStream_t st) {
    if (kmode == KM_MAX)
        return pick_ksplit<TIn>(p, 128) == 2 ? launch_run<TIn, TOut, 8, 4, 2, 128, 4, false, 0, KM_MAX, false, false, kPlan>(p, st)
                                             : launch_run<TIn, TOut, 8, 4, 1, 128, 4, false, 0, KM_MAX, false, false, kPlan>(p, st);
    if (kmode == KM_NONE)
        return pick_ksplit<TIn>(p, 96) == 2 ? launch_run<TIn, TOut, 8, 4, 2, 96, 4, false, 0, KM_NONE, false, false, kPlan>(p, st)
                                            : launch_run<TIn, TOut, 8, 4, 1, 96, 4, false, 0, KM_NONE, false, false, kPlan>(p, st);
    const int ks = pick_ksplit<TIn>(p, maxreg <= 96 ? 96 : 128);
    if (maxreg <= 96) return ks == 2 ? launch_run<TIn, TOut, 8, 4, 2, 96, 4, false, 0, KM_ACC, false, false, kPlan>(p, st) : launch_run<TIn, TOut, 8, 4, 1, 96, 4, false, 0, KM_ACC, false, false, kPlan>(p, st);
    return ks == 2 ? launch_run<TIn, TOut, 8, 4, 2, 128, 4, false, 0, KM_ACC, false, false, kPlan>(p, st) : launch_run<TIn, TOut, 8, 4, 1, 128, 4, false, 0, KM_ACC, false, false, kPlan>(p, st);
}

}  // namespace

int BEVIPM_RUN_ENTRY(const FwdParams& p, bool in_bf16, bool out_bf16, int kmode, int maxreg, cudaStream_t st) {
    if (!in_bf16 && !out_bf16) return launch_typed<float, float>(p, kmode, maxreg, st);
    if (in_bf16 && out_bf16) return launch_typed<__nv_bfloat16, __nv_bfloat16>(p, kmode, maxreg, st);
    if (in_bf16 && !out_bf16) return launch_typed<__nv_bfloat16, float>(p, kmode, maxreg, st);
    return set_error(BEVIPM_ERR_UNSUPPORTED, "fp32 features with a bf16 result: no kernel is built for this pair");
}

}  // namespace bevipm
