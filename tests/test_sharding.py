"""N > 1 host logic on CPU ranks (gloo, world_size 2 and 3): view sharding + all-reduce and frame
sharding, with the oracle standing in for the per-rank CUDA warp."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bevipm import sharding


def test_assignments():
    assert sharding.view_assignment(7, 1) == [[0, 1, 2, 3, 4, 5, 6]]
    assert sharding.view_assignment(7, 2) == [[0, 1, 2, 3], [4, 5, 6]]
    assert [len(v) for v in sharding.view_assignment(7, 4)] == [2, 2, 2, 1]
    assert sharding.view_assignment(7, 8)[-1] == []
    for n, w in ((64, 8), (8, 3), (5, 8), (1, 1)):
        blocks = sharding.frame_assignment(n, w)
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
        sizes = [hi - lo for lo, hi in blocks]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.block_range(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _case():
    from bevipm import rig
    V = 7
    K, Rt = rig.look_at_rig(V, 0)
    feats = torch.randn(2, V, 4, 27, 48, generator=torch.Generator().manual_seed(0)).numpy()
    xs, ys = rig.ground_axes(24, 72, rig.WILDTRACK_BOUNDS)
    Kb = np.ascontiguousarray(np.broadcast_to(K.numpy(), (2, V, 3, 3)))
    Rb = np.ascontiguousarray(np.broadcast_to(Rt.numpy(), (2, V, 4, 4)))
    return feats, Kb, Rb, xs.numpy(), ys.numpy(), rig.WILDTRACK_IMG_SIZE


def _worker(rank, world, port, mode, q):
    from oracle import ipm_oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        feats, K, Rt, xs, ys, img = _case()
        V = feats.shape[1]

        def partial(view_ids):
            ids = list(view_ids)
            red = "max" if mode == "max" else "sum"
            return torch.from_numpy(orc.warp_fuse(np.ascontiguousarray(feats[:, ids]), np.ascontiguousarray(K[:, ids]),
                                                  np.ascontiguousarray(Rt[:, ids]), xs, ys, img, red))

        fuse = sharding.ViewShardedFusion(V, mode)
        out = fuse(partial, (2, 4, 24, 72), torch.device("cpu"))
        # frame sharding: each rank warps its frame block with every view, then all ranks assemble
        lo, hi = sharding.block_range(feats.shape[0], rank, world)
        if hi > lo:
            mine = torch.from_numpy(orc.warp_fuse(feats[lo:hi], K[lo:hi], Rt[lo:hi], xs, ys, img, mode))
        else:
            mine = torch.zeros((0, 4, 24, 72))
        gathered = sharding.gather_frames(mine, feats.shape[0])
        if rank == 0:
            q.put((out.numpy(), gathered.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("mode", ["mean", "sum", "max"])
def test_view_and_frame_sharding_over_gloo(world, mode):
    from oracle import ipm_oracle as orc
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    sharded, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    feats, K, Rt, xs, ys, img = _case()
    want = orc.warp_fuse(feats, K, Rt, xs, ys, img, mode)
    assert np.array_equal(gathered, want)                  # frame sharding changes nothing
    if mode == "max":
        assert np.array_equal(sharded, want)               # max is order-free
    else:
        # the partial sums re-associate the 7-term fp32 sum: a few ulp, inside the 1e-5 budget
        assert np.abs(sharded - want).max() <= 1e-5 * np.abs(want).max()


def test_view_sharding_more_ranks_than_views_identity():
    # a rank with no cameras must not disturb max (identity -inf) nor sum (identity 0)
    for mode, fill in (("max", float("-inf")), ("sum", 0.0)):
        assert sharding.view_assignment(2, 3)[2] == []
    f = sharding.ViewShardedFusion(7, "mean")
    x = torch.full((1, 1, 2, 2), 14.0)
    assert torch.equal(f.finish(x), torch.full((1, 1, 2, 2), 2.0))
