#!/usr/bin/env python
"""Time forward and backward (grad w.r.t. the features) of the fused op on a BASELINE shape, CUDA events, and beside it the
reference's own op chain on ATen's CUDA kernels (oracle/torch_chain.py: one grid_sample per view + the view reduction, whose
backward is ATen's grid_sampler_2d_backward + index_put / sum backward) on the same GPU and the same inputs.
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

# the stock-torch arm: same chain as geometry.py:142-162 + fusion.py:17-22, autograd through ATen
from oracle import torch_chain as tc  # noqa: E402  (development tool: the checker's restatement, never the product path)
f2 = f.detach().float().contiguous().requires_grad_(True)          # NCHW-contiguous, what grid_sample wants
K4 = Kd
R4 = torch.cat([Rd, torch.tensor([0.0, 0.0, 0.0, 1.0], device=dev).expand(B, V, 1, 4)], dim=2)


def aten_fwd():
    return tc.fuse(tc.warp_views(f2, K4, R4, xd, yd, wl.img_size), mode if mode != "none" else "concat")


def aten_fwd_bwd():
    f2.grad = None
    o = aten_fwd()
    o.backward(cot.reshape(o.shape) if cot.numel() == o.numel() else torch.ones_like(o))


with torch.no_grad():
    af = t(aten_fwd, 3)
ab = t(aten_fwd_bwd, 3)
g_ours = torch.autograd.grad(fwd(), f, cot)[0]
o2 = aten_fwd()
g_aten = torch.autograd.grad(o2, f2, cot.reshape(o2.shape))[0]
rel = float((g_ours - g_aten).abs().max() / g_aten.abs().max())
print(f"{wl.name} B={B} mode={mode}: ATen chain forward {af:.3f} ms, forward+backward {ab:.3f} ms (backward ~ {ab - af:.3f} ms); "
      f"ours vs ATen backward: {(ab - af) / max(tb - tf, 1e-9):.1f}x faster, gradients differ by {rel:.2e} (max-normalised)")
