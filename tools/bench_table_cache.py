#!/usr/bin/env python
"""The default fused kernel with and without the table cache (bevipm_warp_fuse_fwd_planned: phase-A tables of every row segment
kept on the device between calls, static cameras) on the BASELINE shapes.  CUDA events, inputs resident in HBM and rotated
through enough buffers to stay out of L2 (tools/sweep_variants.py conventions).  Prints one JSON object.
usage: bench_table_cache.py [c1,c2,c3,c5]"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "vision-based-spatio-temporal-analysis_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from bevipm import _lib, ops, rig  # noqa: E402

dev = "cuda:0"
peak = 6543.1
try:
    peak = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    pass
res = {"peak_gbs": peak}
for name in (sys.argv[1] if len(sys.argv) > 1 else "c1,c2,c3").split(","):
    wl = rig.WORKLOADS[name]
    B, V, C = wl.frames, wl.views, wl.channels
    tdt = torch.bfloat16 if wl.dtype == "bf16" else torch.float32
    obf = wl.out_dtype == "bf16"
    K, Rt = rig.look_at_rig(V, 0)
    Kd = K[None].expand(B, -1, -1, -1).contiguous().to(dev)
    Rd = Rt[None, :, :3, :].expand(B, -1, -1, -1).contiguous().to(dev)
    xs, ys = rig.ground_axes(*wl.bev_hw, wl.bounds)
    xd, yd = xs.to(dev), ys.to(dev)
    in_bytes = B * V * wl.feat_hw[0] * wl.feat_hw[1] * C * wl.feat_elem_bytes
    nbuf = max(1, min(8, int(300e6 // in_bytes) + 1))
    g = torch.Generator(device=dev).manual_seed(0)
    bufs = []
    for _ in range(nbuf):
        f = torch.empty((B, V, *wl.feat_hw, C), device=dev, dtype=tdt)
        for b in range(B):
            f[b] = torch.randn((V, *wl.feat_hw, C), device=dev, generator=g).to(tdt)
        bufs.append(f.permute(0, 1, 4, 2, 3))
    m = _lib.MODES[wl.fusion]
    plan = ops.new_plan(V, wl.bev_hw, dev)
    plain = lambda i: ops.warp_fuse(bufs[i % nbuf], Kd, Rd, xd, yd, wl.img_size[0], wl.img_size[1], m, obf, 0)
    planned = lambda i: ops.warp_fuse_planned(bufs[i % nbuf], Kd, Rd, xd, yd, wl.img_size[0], wl.img_size[1], m, obf, plan)

    def timed(fn, iters=60):
        for i in range(4):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    same = bool(torch.equal(plain(0), planned(0)))
    t = {"plain_ms": [], "cached_ms": []}
    for _ in range(3):   # interleaved repeats
        t["plain_ms"].append(timed(plain))
        t["cached_ms"].append(timed(planned))
    ix, iy = ops.sample_coords(Kd[:1], Rd[:1], xd, yd, wl.feat_hw, wl.img_size)
    alg = rig.algorithmic_bytes(ix[0].cpu().numpy(), iy[0].cpu().numpy(), wl.feat_hw, C, wl.feat_elem_bytes, 2 if obf else 4, per_view_out=False)
    bp, bc = min(t["plain_ms"]), min(t["cached_ms"])
    res[name] = {"plain_ms": bp, "cached_ms": bc, "speedup": bp / bc, "identical_result": same, "cache_bytes": int(plan.numel()),
                 "algorithmic_bytes_per_launch": alg["b_alg"] * B, "plain_frac": alg["b_alg"] * B / bp / 1e6 / peak,
                 "cached_frac": alg["b_alg"] * B / bc / 1e6 / peak, "runs": t}
    del bufs
    torch.cuda.empty_cache()
print(json.dumps(res))
