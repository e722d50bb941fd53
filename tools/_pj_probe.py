#!/usr/bin/env python
"""Development probe: the projection GEMM alone at wildtrack.yaml's shape under BEVIPM_PJ_* switches.  usage: _pj_probe.py [passes]"""
import os
import sys
from pathlib import Path

R = str(Path(__file__).resolve().parents[1])
sys.path[:0] = [R, R + "/vision-based-spatio-temporal-analysis_b200"]
import torch  # noqa: E402
from bevipm import ops  # noqa: E402

passes = int(sys.argv[1]) if len(sys.argv) > 1 else 1
V, rows, C, Co = 7, 135 * 240, 1280, 128
g = torch.Generator(device="cuda").manual_seed(0)
xs = [torch.randn(V, rows, C, device="cuda", generator=g) for _ in range(3)]
W = torch.randn(Co, V, C, device="cuda", generator=g) / C ** 0.5
alg = xs[0].numel() * 4 + W.numel() * 4 + V * rows * Co * 4
for i in range(3):
    ops.proj1x1(xs[i], W, passes)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(30):
    ops.proj1x1(xs[i % 3], W, passes)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 30
print({k: v for k, v in os.environ.items() if k.startswith("BEVIPM_PJ")}, "passes", passes, "ms %.4f" % ms, "GB/s %.0f" % (alg / ms / 1e6), flush=True)
