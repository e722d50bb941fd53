"""View sharding on real GPUs (BASELINE configs[2]): 2 ranks, NCCL + peer memory.  Skipped with fewer than 2 devices.

Every form must reproduce the oracle's fused result; the sum order differs from the sequential view order (partials are
summed across ranks), so the bound is the north star's 1e-5 relative, not bit-exactness."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from bevipm import _lib, ops, rig, sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        V, C, fhw, bhw, B = 7, 128, (40, 64), (36, 96), 2
        K, Rt = rig.look_at_rig(V, 0)
        feats = torch.randn(B, V, C, *fhw, generator=torch.Generator().manual_seed(0))
        xs, ys = rig.ground_axes(*bhw, rig.WILDTRACK_BOUNDS)
        img = rig.WILDTRACK_IMG_SIZE
        ids = sharding.view_assignment(V, world)[rank]
        f_r = feats[:, ids].to(dev).permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)
        K_r = K[ids][None].expand(B, -1, -1, -1).contiguous().to(dev)
        R_r = Rt[ids, :3, :][None].expand(B, -1, -1, -1).contiguous().to(dev)
        xd, yd = xs.to(dev), ys.to(dev)
        rows = sharding.slab_rows(bhw[0], world)
        res = {}
        # (1) all-reduce: the replicated BEV
        fuse = sharding.ViewShardedFusion(V, "mean")
        full = fuse(lambda v: ops.warp_fuse(f_r, K_r, R_r, xd, yd, img[0], img[1], _lib.SUM, False, 0), (B, C, *bhw), dev)
        res["allreduce"] = full.cpu().numpy()
        # (2) reduce-scatter on a side stream, two frames in flight
        rs = sharding.ReduceScatterFusion(V, bhw, C, "mean", device=dev)
        tickets = [rs.submit(ops.warp_fuse(f_r, K_r, R_r, xd, yd, img[0], img[1], _lib.SUM, False, 0)) for _ in range(3)]
        slabs = [rs.wait(t) for t in tickets]
        assert all(torch.equal(slabs[0], s) for s in slabs[1:])
        res["reduce_scatter"] = slabs[0].cpu().numpy()
        # (3) fused: partial sums sent into the owners' buffers through peer memory: stores + owner-side sum, and adds
        ps = sharding.PeerSlabFusion(V, bhw, C, frames=B, mode="mean", device=dev, put=True)
        tickets = [ps.submit(f_r, K_r, R_r, xd, yd, img) for _ in range(5)]      # both buffer sets, pipelined
        outs = [ps.wait(t).clone() for t in tickets]
        for o in outs[1:]:
            assert torch.equal(o, outs[0])                                       # stores + ordered sum: reproducible bit for bit
        res["peer_slab_put"] = outs[0].cpu().numpy()
        pa = sharding.PeerSlabFusion(V, bhw, C, frames=B, mode="mean", device=dev, put=False)
        outs = [pa.run(f_r, K_r, R_r, xd, yd, img).clone() for _ in range(4)]
        for o in outs[1:]:
            assert torch.allclose(o, outs[0], rtol=1e-6, atol=1e-7)              # (reductions: the add order may differ run to run)
        res["peer_slab_add"] = outs[0].cpu().numpy()
        res["nvlink_bytes"] = ps.bytes_over_nvlink_per_call()
        # (4) the same exchange on the copy engines (DMA peer copies on a second stream)
        ce = sharding.CopyEngineSlabFusion(V, bhw, C, frames=B, mode="mean", device=dev)
        tickets = [ce.submit(ops.warp_fuse(f_r, K_r, R_r, xd, yd, img[0], img[1], _lib.SUM, False, 0)) for _ in range(5)]
        outs = [ce.wait(t).clone() for t in tickets]
        for o in outs:
            assert torch.equal(o, torch.from_numpy(res["peer_slab_put"]).to(dev))   # same ordered sum, bit for bit
        res["copy_engine"] = outs[0].cpu().numpy()
        torch.cuda.synchronize()
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **res)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_view_sharding_two_gpus(tmp_path):
    import torch.multiprocessing as mp
    from bevipm import rig, sharding
    from oracle import ipm_oracle as orc
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    V, C, fhw, bhw, B = 7, 128, (40, 64), (36, 96), 2
    K, Rt = rig.look_at_rig(V, 0)
    feats = torch.randn(B, V, C, *fhw, generator=torch.Generator().manual_seed(0)).numpy()
    xs, ys = rig.ground_axes(*bhw, rig.WILDTRACK_BOUNDS)
    Kb = np.ascontiguousarray(np.broadcast_to(K.numpy(), (B, V, 3, 3)))
    Rb = np.ascontiguousarray(np.broadcast_to(Rt.numpy(), (B, V, 4, 4)))
    want = orc.warp_fuse(feats, Kb, Rb, xs.numpy(), ys.numpy(), rig.WILDTRACK_IMG_SIZE, "mean")
    tol = 1e-5 * np.abs(want).max()
    rows = sharding.slab_rows(bhw[0], world)
    for rank in range(world):
        z = np.load(tmp_path / f"rank{rank}.npz")
        assert np.abs(z["allreduce"] - want).max() <= tol
        lo, hi = rank * rows, min(bhw[0], (rank + 1) * rows)
        for name in ("reduce_scatter", "peer_slab_put", "peer_slab_add", "copy_engine"):
            got = z[name][:, :, : hi - lo]
            assert got.shape == want[:, :, lo:hi].shape, name
            assert np.abs(got - want[:, :, lo:hi]).max() <= tol, name
        assert int(z["nvlink_bytes"]) == B * (bhw[0] - (hi - lo)) * bhw[1] * C * 4
