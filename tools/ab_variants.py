#!/usr/bin/env python
"""A/B timing of fused-kernel variants with interleaved repeats (CUDA events, inputs resident in HBM).
usage: ab_variants.py <workload> <v1,v2,...> [repeats] [iters]   -> best and median ms per variant"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "vision-based-spatio-temporal-analysis_b200"), str(ROOT / "tools")):
    sys.path.insert(0, p)
import statistics
import sweep_variants as sv
from bevipm import rig

wl = rig.WORKLOADS[sys.argv[1]]
vs = [int(x) for x in sys.argv[2].split(",")]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 100
res = {v: [] for v in vs}
for _ in range(reps):
    for v in vs:
        r = sv.time_variant(wl, v, iters=iters)
        res[v].append(r.get("ms", float("nan")))
for v in vs:
    print(f"{wl.name} variant {v:3d}: best {min(res[v]):.4f} ms  median {statistics.median(res[v]):.4f} ms  {['%.4f' % x for x in res[v]]}")
