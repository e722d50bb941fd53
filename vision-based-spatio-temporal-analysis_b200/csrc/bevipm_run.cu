// bevipm_run.cu -- the DEFAULT instantiations of the run kernel (ipm_run.cuh), in a translation unit of their own:
//   sum / mean: fp32 features at 96 registers (variant 32), bf16 features at 128 (variant 33);
//   max (fusion.py:22) at 128 registers; per-view maps (geometry.py:162-163, what ConcatFusion reshapes) at 96.
// Everything else of the run-kernel family (ring depth, cells per segment, TMA ring, half reloads, timing probes) is a
// sweep variant compiled in bevipm_api.cu.  Keeping the kernels every default launch runs apart from ~150 sweep
// instantiations makes their code generation independent of what else is being swept (the same source gave 4160 or 4120
// instructions and +-3 % on config 1 depending on its neighbours), and a change to them rebuilds in under a minute.
#include "run_launch.cuh"

namespace bevipm {
namespace {

template <typename TIn, typename TOut>
int launch_typed(const FwdParams& p, int kmode, int maxreg, cudaStream_t st) {
    if (kmode == KM_MAX) return launch_run<TIn, TOut, 8, 4, 1, 128, 4, false, 0, KM_MAX>(p, st);
    if (kmode == KM_NONE) return launch_run<TIn, TOut, 8, 4, 1, 96, 4, false, 0, KM_NONE>(p, st);
    if (maxreg <= 96) return launch_run<TIn, TOut, 8, 4, 1, 96, 4, false>(p, st);
    return launch_run<TIn, TOut, 8, 4, 1, 128, 4, false>(p, st);
}

}  // namespace

int launch_run_default(const FwdParams& p, bool in_bf16, bool out_bf16, int kmode, int maxreg, cudaStream_t st) {
    if (!in_bf16 && !out_bf16) return launch_typed<float, float>(p, kmode, maxreg, st);
    if (in_bf16 && out_bf16) return launch_typed<__nv_bfloat16, __nv_bfloat16>(p, kmode, maxreg, st);
    if (in_bf16 && !out_bf16) return launch_typed<__nv_bfloat16, float>(p, kmode, maxreg, st);
    return set_error(BEVIPM_ERR_UNSUPPORTED, "fp32 features with a bf16 result: no kernel is built for this pair");
}

}  // namespace bevipm
