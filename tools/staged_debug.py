#!/usr/bin/env python
"""Smallest staged-kernel launch against the oracle (development aid): staged_debug.py <variant> [V C Hf Wf Hb Wb B]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "vision-based-spatio-temporal-analysis_b200"), str(ROOT / "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
from oracle import ipm_oracle as orc
from test_gpu_parity import _rig_case, _run, _same

variant = int(sys.argv[1])
a = [int(x) for x in sys.argv[2:]]
V, C, Hf, Wf, Hb, Wb, B = (a + [3, 128, 20, 33, 19, 45, 1][len(a):])[:7]
feats, K, Rt, xs, ys, img = _rig_case(B, V, C, (Hf, Wf), (Hb, Wb), seed=5)
want = orc.warp_fuse(feats, K, Rt, xs, ys, img, "mean")
out = _run(feats, K, Rt, xs, ys, img, "mean", True, variant=variant).cpu().numpy()
same = _same(out, want)
print("variant", variant, "shape", (V, C, Hf, Wf, Hb, Wb, B), "bit-exact:", same, "max abs diff", float(np.nanmax(np.abs(out - want))),
      "mismatches", int((out != want).sum()), "of", out.size, flush=True)
if not same and variant != 52:
    bad = np.argwhere(out != want)
    print("first mismatches (b, c, i, j):", bad[:8].tolist())
    print("rows with mismatches:", sorted(set(bad[:, 2].tolist()))[:40], "cols:", sorted(set(bad[:, 3].tolist()))[:60])
