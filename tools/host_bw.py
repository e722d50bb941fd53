#!/usr/bin/env python
"""Host-memory bandwidth ceiling of the box for the end-to-end path (pinned H2D + D2H, all GPUs at once).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/host_bw.py [--mb 1024]

Every rank streams a pinned host buffer to its GPU and another one back, on two streams, for a few seconds, exactly the
traffic pattern of bevipm_warp_fuse_host (uploads of the sampled source-row spans, download of the BEV) without any
kernel in between.  Rank 0 prints the aggregate GB/s in each direction and both at once: what `e2e` can reach at most
on this box, whatever the kernels do.  NUMA: the box's topology is printed beside it (nvidia-smi topo, lscpu)."""
import argparse
import json
import os
import subprocess
import time

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=1024)
    ap.add_argument("--seconds", type=float, default=2.0)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.mb << 20
    h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True).fill_(1)
    h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d_in = torch.empty(n, dtype=torch.uint8, device=dev)
    d_out = torch.empty(n, dtype=torch.uint8, device=dev).fill_(2)
    s_up, s_dn = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def run(up: bool, down: bool):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t0 < args.seconds:
            if up:
                with torch.cuda.stream(s_up):
                    d_in.copy_(h_in, non_blocking=True)
            if down:
                with torch.cuda.stream(s_dn):
                    h_out.copy_(d_out, non_blocking=True)
            s_up.synchronize()
            s_dn.synchronize()
            reps += 1
        dt = time.perf_counter() - t0
        gbs = torch.tensor([reps * n * (int(up) + int(down)) / dt / 1e9], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(gbs)
        return float(gbs.item())

    res = {"gpus": world, "buffer_mb": args.mb, "h2d_gbs": run(True, False), "d2h_gbs": run(False, True), "both_gbs": run(True, True)}
    if rank == 0:
        try:
            res["numa_nodes"] = subprocess.run("lscpu | grep -i 'numa node(s)'", shell=True, capture_output=True, text=True).stdout.strip()
            res["topo"] = subprocess.run("nvidia-smi topo -m | head -3 | cut -c1-200", shell=True, capture_output=True, text=True).stdout.strip()[-160:]
        except Exception:
            pass
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
