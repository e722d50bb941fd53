#!/usr/bin/env python
"""Time the fused forward in every fusion mode on a BASELINE shape (CUDA events, inputs resident in HBM):
sum / mean / max write one BEV map, none (= concat, what BEVNet's ConcatFusion consumes) writes V maps.
usage: bench_modes.py <workload> [variant]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "vision-based-spatio-temporal-analysis_b200"), str(ROOT / "tools")):
    sys.path.insert(0, p)
import sweep_variants as sv
from bevipm import rig

wl = rig.WORKLOADS[sys.argv[1]]
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 0
for mode in ("mean", "sum", "max", "none"):
    r = sv.time_variant(wl, variant, mode=mode, iters=40)
    if "ms" in r:
        print(f"{wl.name} {mode:5s} variant {variant}: {r['ms']:.4f} ms  {r['alg_gbs']:.0f} GB/s algorithmic ({r['b_alg_frame'] / 1e6:.1f} MB/frame)")
    else:
        print(wl.name, mode, r)
