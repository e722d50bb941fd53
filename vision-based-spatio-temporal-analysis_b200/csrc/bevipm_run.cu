// bevipm_run.cu -- the DEFAULT instantiations of the run kernel (ipm_run.cuh), in a translation unit of their own:
//   sum / mean: fp32 features at 96 registers (variant 32), bf16 features at 128 (variant 33), each with one or two warps
//   per row segment (pick_ksplit);
//   max (fusion.py:22) at 128 registers; per-view maps (geometry.py:162-163, what ConcatFusion reshapes) at 96.
// Everything else of the run-kernel family (ring depth, cells per segment, TMA ring, half reloads, timing probes) is a
// sweep variant compiled in bevipm_api.cu.  Keeping the kernels every default launch runs apart from ~150 sweep
// instantiations makes their code generation independent of what else is being swept (the same source gave 4160 or 4120
// instructions and +-3 % on config 1 depending on its neighbours), and a change to them rebuilds in under a minute.
#include "run_launch.cuh"

namespace bevipm {
namespace {

// Warps per row segment for the sum / mean walk.  One warp per segment (it walks all channel chunks itself) has the least
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

template <typename TIn, typename TOut>
int launch_typed(const FwdParams& p, int kmode, int maxreg, cudaStream_t st) {
    if (kmode == KM_MAX)
        return pick_ksplit<TIn>(p, 128) == 2 ? launch_run<TIn, TOut, 8, 4, 2, 128, 4, false, 0, KM_MAX>(p, st)
                                             : launch_run<TIn, TOut, 8, 4, 1, 128, 4, false, 0, KM_MAX>(p, st);
    if (kmode == KM_NONE)
        return pick_ksplit<TIn>(p, 96) == 2 ? launch_run<TIn, TOut, 8, 4, 2, 96, 4, false, 0, KM_NONE>(p, st)
                                            : launch_run<TIn, TOut, 8, 4, 1, 96, 4, false, 0, KM_NONE>(p, st);
    const int ks = pick_ksplit<TIn>(p, maxreg <= 96 ? 96 : 128);
    if (maxreg <= 96) return ks == 2 ? launch_run<TIn, TOut, 8, 4, 2, 96, 4, false>(p, st) : launch_run<TIn, TOut, 8, 4, 1, 96, 4, false>(p, st);
    return ks == 2 ? launch_run<TIn, TOut, 8, 4, 2, 128, 4, false>(p, st) : launch_run<TIn, TOut, 8, 4, 1, 128, 4, false>(p, st);
}

}  // namespace

int launch_run_default(const FwdParams& p, bool in_bf16, bool out_bf16, int kmode, int maxreg, cudaStream_t st) {
    if (!in_bf16 && !out_bf16) return launch_typed<float, float>(p, kmode, maxreg, st);
    if (in_bf16 && out_bf16) return launch_typed<__nv_bfloat16, __nv_bfloat16>(p, kmode, maxreg, st);
    if (in_bf16 && !out_bf16) return launch_typed<__nv_bfloat16, float>(p, kmode, maxreg, st);
    return set_error(BEVIPM_ERR_UNSUPPORTED, "fp32 features with a bf16 result: no kernel is built for this pair");
}

}  // namespace bevipm
