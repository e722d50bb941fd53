#!/usr/bin/env python
"""Time every fused-kernel variant on the BASELINE shapes (CUDA events, inputs resident in HBM).

    python tools/sweep_variants.py [--workloads c1,c2,c3] [--out gpurun_out/sweep.json]

For the batch-1 shapes (c1, c3: working set near or below the 126 MB L2) the launches rotate through
enough distinct feature buffers that one pass touches > 2 x L2 before a buffer is reused.
"""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "vision-based-spatio-temporal-analysis_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

from bevipm import _lib, ops, rig  # noqa: E402


def time_variant(wl, variant, iters=40, dev="cuda:0", mode=None, out_bf16=None):
    B, V, C = wl.frames, wl.views, wl.channels
    tdt = torch.bfloat16 if wl.dtype == "bf16" else torch.float32
    obf = (wl.out_dtype == "bf16") if out_bf16 is None else out_bf16
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
    m = _lib.MODES[mode or wl.fusion]
    try:
        for i in range(3):
            ops.warp_fuse(bufs[i % nbuf], Kd, Rd, xd, yd, wl.img_size[0], wl.img_size[1], m, obf, variant)
    except RuntimeError as e:
        return {"variant": variant, "error": str(e)}
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        ops.warp_fuse(bufs[i % nbuf], Kd, Rd, xd, yd, wl.img_size[0], wl.img_size[1], m, obf, variant)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    ix, iy = ops.sample_coords(Kd[:1], Rd[:1], xd, yd, wl.feat_hw, wl.img_size)
    alg = rig.algorithmic_bytes(ix[0].cpu().numpy(), iy[0].cpu().numpy(), wl.feat_hw, C, wl.feat_elem_bytes,
                                2 if obf else 4, per_view_out=(m == _lib.NONE))
    gbs = alg["b_alg"] * B / (ms * 1e-3) / 1e9
    return {"variant": variant, "ms": ms, "frames_per_s": B / (ms * 1e-3), "alg_gbs": gbs, "buffers": nbuf,
            "b_alg_frame": alg["b_alg"]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="c1,c2,c3")
    ap.add_argument("--variants", default="1,2,3,4,5,6,7,8,9,10")
    ap.add_argument("--out", default="gpurun_out/sweep.json")
    args = ap.parse_args()
    res = {}
    for name in args.workloads.split(","):
        wl = rig.WORKLOADS[name]
        rows = []
        for v in [int(x) for x in args.variants.split(",")]:
            r = time_variant(wl, v)
            rows.append(r)
            print(name, r, flush=True)
        res[name] = rows
        torch.cuda.empty_cache()
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    Path(args.out).write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
