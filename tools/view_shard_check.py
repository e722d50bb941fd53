#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/view_shard_check.py : view-sharded fusion over NCCL == single-GPU fused result
(fp32 re-association only), on real CUDA ranks.  Prints one line per mode from rank 0."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "vision-based-spatio-temporal-analysis_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch
import torch.distributed as dist

from bevipm import _lib, ops, rig, sharding


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    V, C, fhw, bhw, B = 7, 64, (68, 120), (60, 180), 2
    K, Rt = rig.look_at_rig(V, 0)
    Kd = K[None].expand(B, -1, -1, -1).contiguous().to(dev)
    Rd = Rt[None, :, :3, :].expand(B, -1, -1, -1).contiguous().to(dev)
    xs, ys = rig.ground_axes(*bhw, rig.WILDTRACK_BOUNDS)
    xd, yd = xs.to(dev), ys.to(dev)
    feats = torch.randn(B, V, *fhw, C, device=dev, generator=torch.Generator(device=dev).manual_seed(0)).permute(0, 1, 4, 2, 3)
    img = rig.WILDTRACK_IMG_SIZE
    for mode in ("sum", "mean", "max"):
        full = ops.warp_fuse(feats, Kd, Rd, xd, yd, img[0], img[1], _lib.MODES[mode], False, 0)
        red = "max" if mode == "max" else "sum"

        def partial(ids):
            sl = slice(ids[0], ids[-1] + 1)
            return ops.warp_fuse(feats[:, sl], Kd[:, sl].contiguous(), Rd[:, sl].contiguous(), xd, yd, img[0], img[1],
                                 _lib.MODES[red], False, 0)

        out = sharding.ViewShardedFusion(V, mode)(partial, tuple(full.shape), dev)
        err = float((out - full).abs().max() / full.abs().max())
        ok = torch.equal(out, full) if mode == "max" else err <= 1e-5
        if rank == 0:
            print(f"view-sharded {mode} over {world} ranks: max-normalised diff {err:.2e} {'OK' if ok else 'FAIL'}", flush=True)
        assert ok
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
