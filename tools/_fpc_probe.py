#!/usr/bin/env python
"""Development probe: ms per frame of the default kernel for a BASELINE shape at `frames` frames per launch, forcing the number of
frames a CTA walks with one set of phase-A tables (BEVIPM_RUN_FPC).  usage: _fpc_probe.py <workload> <frames> [channels]"""
import dataclasses
import os
import sys
from pathlib import Path

R = str(Path(__file__).resolve().parents[1])
sys.path[:0] = [R, R + "/vision-based-spatio-temporal-analysis_b200", R + "/tools"]
import sweep_variants as sv  # noqa: E402
from bevipm import rig  # noqa: E402

base = rig.WORKLOADS[sys.argv[1]]
frames = int(sys.argv[2])
kw = {"name": f"{base.name}x{frames}", "frames": frames}
if len(sys.argv) > 3:
    kw["channels"] = int(sys.argv[3])
wl = dataclasses.replace(base, **kw)
for fpc in ("auto", 1, 2, 4, 8):
    os.environ.pop("BEVIPM_RUN_FPC", None)
    if fpc != "auto":
        if fpc > frames:
            continue
        os.environ["BEVIPM_RUN_FPC"] = str(fpc)
    r = sv.time_variant(wl, 0, iters=10 if frames > 16 else 20)
    print(wl.name, "C", wl.channels, "fpc", fpc, "ms/frame %.5f" % (r["ms"] / frames), flush=True)
