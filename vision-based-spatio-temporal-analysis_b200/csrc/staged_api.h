// staged_api.h -- launcher of the TMA-staged fused kernel (ipm_staged.cuh), compiled in its own translation unit.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "ipm_fused.cuh"

namespace bevipm {

// Launch warp_fuse_staged_kernel for `p` (fast-path strides already checked by the caller).
//   in_bf16 / out_bf16: element types; shape: 0 = 8 rows per tile (2 CTAs per SM), 1 = 4 rows per tile; probe: timing aid.
// Returns 0, or a negative bevipm_status with a message in err (BEVIPM_ERR_UNSUPPORTED when the shape cannot take the
// staged kernel: the caller falls back to the run kernel when the variant was not forced).
int launch_staged(FwdParams p, bool in_bf16, bool out_bf16, int shape, int probe, cudaStream_t st, char* err, size_t errlen);

// The run kernel with its ring filled by one tensor-map [2 x 2] box copy per reload (ipm_boxrun.cuh); sum / mean / max.
int launch_boxrun(FwdParams p, bool in_bf16, bool out_bf16, cudaStream_t st, char* err, size_t errlen);

// Can the staged kernel take this launch at all (TMA stride rules, map sizes, fusion mode)?
bool staged_supported(const FwdParams& p, bool in_bf16);

}  // namespace bevipm
