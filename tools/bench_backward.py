#!/usr/bin/env python
"""Time forward and backward (grad w.r.t. the features) of the fused op on a BASELINE shape, CUDA events.
usage: bench_backward.py <workload> [mode]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "vision-based-spatio-temporal-analysis_b200")):
    sys.path.insert(0, p)
import torch
from bevipm import _lib, ops, rig

wl = rig.WORKLOADS[sys.argv[1]]
mode = sys.argv[2] if len(sys.argv) > 2 else wl.fusion
dev = "cuda:0"
B, V, C = min(wl.frames, 2), wl.views, wl.channels
K, Rt = rig.look_at_rig(V, 0)
Kd = K[None].expand(B, -1, -1, -1).contiguous().to(dev)
Rd = Rt[None, :, :3, :].expand(B, -1, -1, -1).contiguous().to(dev)
xs, ys = rig.ground_axes(*wl.bev_hw, wl.bounds)
xd, yd = xs.to(dev), ys.to(dev)
f = torch.randn(B, V, *wl.feat_hw, C, device=dev).permute(0, 1, 4, 2, 3).requires_grad_(True)
out = ops.warp_fuse(f, Kd, Rd, xd, yd, wl.img_size[0], wl.img_size[1], _lib.MODES[mode], False, 0)
cot = torch.randn_like(out)


def t(fn, n=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def fwd():
    return ops.warp_fuse(f, Kd, Rd, xd, yd, wl.img_size[0], wl.img_size[1], _lib.MODES[mode], False, 0)


def fwd_bwd():
    f.grad = None
    o = fwd()
    o.backward(cot)


tf = t(fwd)
tb = t(fwd_bwd)
print(f"{wl.name} B={B} mode={mode}: forward {tf:.3f} ms, forward+backward {tb:.3f} ms (backward incl. zero-fill of the gradient ~ {tb - tf:.3f} ms)")
