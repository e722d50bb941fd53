#!/usr/bin/env python
"""A/B timing of fused-kernel variants under development switches, interleaved repeats on one box.

usage: ab_env.py <workload> <spec,spec,...> [repeats] [iters]
  spec = variant[:NAME=VALUE[:NAME=VALUE...]]   e.g.  33  50:BEVIPM_ST_S=32768:BEVIPM_ST_D=3  51
Prints best and median ms per spec (CUDA events, inputs resident in HBM, see sweep_variants.time_variant)."""
import os
import statistics
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "vision-based-spatio-temporal-analysis_b200"), str(ROOT / "tools")):
    sys.path.insert(0, p)
import sweep_variants as sv  # noqa: E402
from bevipm import rig  # noqa: E402

wl = rig.WORKLOADS[sys.argv[1]]
specs = sys.argv[2].split(",")
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 50
SWITCHES = ("BEVIPM_ST_RING", "BEVIPM_ST_CAP", "BEVIPM_ST_LAG", "BEVIPM_ST_CPS", "BEVIPM_RUN_FPC")
res = {s: [] for s in specs}
for _ in range(reps):
    for s in specs:
        parts = s.split(":")
        for k in SWITCHES:
            os.environ.pop(k, None)
        for kv in parts[1:]:
            k, v = kv.split("=")
            os.environ[k] = v
        r = sv.time_variant(wl, int(parts[0]), iters=iters)
        if "error" in r:
            print(s, r["error"], flush=True)
        res[s].append(r.get("ms", float("nan")))
for s in specs:
    print(f"{wl.name} {s:48s}: best {min(res[s]):.4f} ms  median {statistics.median(res[s]):.4f} ms  {['%.4f' % x for x in res[s]]}", flush=True)
