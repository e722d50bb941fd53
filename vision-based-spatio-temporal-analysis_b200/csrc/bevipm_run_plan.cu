// bevipm_run_plan.cu -- the default run kernels once more, as PLAN instantiations: they read the phase-A tables of every row
// segment from a device-side cache keyed by calibration (and fill an empty cache from frame 0), see ipm_run.cuh / include/bevipm.h
// (bevipm_plan_bytes, bevipm_warp_fuse_fwd_planned).  Separate translation unit: the code of the plain kernels stays untouched.
#define BEVIPM_RUN_PLAN 1
#include "bevipm_run.cu"
