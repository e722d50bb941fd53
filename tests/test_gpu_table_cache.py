"""Table cache across launches (include/bevipm.h: bevipm_plan_bytes / bevipm_warp_fuse_fwd_planned; SURVEY.md 8(f) N4, the
`_grid_cache` the reference declares at geometry.py:22): whatever the state of the cache, the result must equal the oracle
(= the plain entry) bit for bit -- empty cache, filled cache, another calibration, mixed calibrations in one batch, reset."""
import numpy as np
import pytest
import torch

from oracle import ipm_oracle as orc
from test_gpu_parity import _dev_inputs, _rig_case, _same

pytestmark = pytest.mark.gpu


def _planned(feats, K, Rt, xs, ys, img, mode, plan, dtype=torch.float32, out_bf16=False):
    from bevipm import _lib, ops
    f, Kd, Rd, xd, yd = _dev_inputs(feats, K, Rt, xs, ys, True, dtype)
    out = ops.warp_fuse_planned(f, Kd, Rd, xd, yd, int(img[0]), int(img[1]), _lib.MODES[mode], out_bf16, plan)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def _header_key(plan):
    return int(plan[:8].view(torch.int64).item())


@pytest.mark.parametrize("mode", ["mean", "sum", "max", "none"])
@pytest.mark.parametrize("C,dtype", [(128, torch.float32), (512, torch.float32), (384, torch.float32), (512, torch.bfloat16), (264, torch.bfloat16)])
def test_planned_equals_oracle_in_every_cache_state(mode, C, dtype):
    from bevipm import ops
    B, V, fhw, bhw = 3, 5, (31, 53), (37, 91)
    feats, K, Rt, xs, ys, img = _rig_case(B, V, C, fhw, bhw, seed=5)
    f = torch.from_numpy(feats).to(dtype).float().numpy()
    want = orc.warp_fuse(f, K, Rt, xs, ys, img, mode)
    plan = ops.new_plan(V, bhw, "cuda")
    assert _header_key(plan) == 0
    out1 = _planned(f, K, Rt, xs, ys, img, mode, plan, dtype)          # empty cache: builds, fills, publishes
    assert _same(out1, want)
    assert _header_key(plan) != 0
    snapshot = plan.clone()
    out2 = _planned(f, K, Rt, xs, ys, img, mode, plan, dtype)          # filled cache: tables are copied
    assert _same(out2, want)
    assert torch.equal(plan, snapshot)                                  # a valid cache is never written again
    # another calibration: the cache does not match, the frames are computed as before, the cache stays as it is
    K2, Rt2 = K.copy(), Rt.copy()
    K2[:, :, 0, 0] *= 1.02
    want2 = orc.warp_fuse(f, K2, Rt2, xs, ys, img, mode)
    assert _same(_planned(f, K2, Rt2, xs, ys, img, mode, plan, dtype), want2)
    assert torch.equal(plan, snapshot)
    # a batch that mixes the cached calibration (frames 0, 2) with another one (frame 1)
    K3, Rt3 = K.copy(), Rt.copy()
    K3[1] = K2[1]
    want3 = orc.warp_fuse(f, K3, Rt3, xs, ys, img, mode)
    assert _same(_planned(f, K3, Rt3, xs, ys, img, mode, plan, dtype), want3)
    # re-armed for the new calibration
    plan.zero_()
    assert _same(_planned(f, K2, Rt2, xs, ys, img, mode, plan, dtype), want2)
    assert _header_key(plan) != 0 and not torch.equal(plan, snapshot)
    assert _same(_planned(f, K2, Rt2, xs, ys, img, mode, plan, dtype), want2)


def test_a_cache_filled_by_one_shape_is_not_used_by_another():
    """Same views and BEV grid (same cache size), other feature-map shape / channel count / axes: the key or the axis ends differ,
    the launch computes its own tables."""
    from bevipm import ops, rig
    V, bhw = 4, (20, 40)
    plan = ops.new_plan(V, bhw, "cuda")
    feats, K, Rt, xs, ys, img = _rig_case(1, V, 128, (24, 40), bhw, seed=2)
    want = orc.warp_fuse(feats, K, Rt, xs, ys, img, "mean")
    assert _same(_planned(feats, K, Rt, xs, ys, img, "mean", plan), want)
    key = _header_key(plan)
    for C, fhw in ((256, (24, 40)), (128, (30, 40))):
        f2, K2, Rt2, xs2, ys2, img2 = _rig_case(1, V, C, fhw, bhw, seed=2)
        assert _same(_planned(f2, K2, Rt2, xs2, ys2, img2, "mean", plan), orc.warp_fuse(f2, K2, Rt2, xs2, ys2, img2, "mean"))
        assert _header_key(plan) == key
    xs3, ys3 = rig.ground_axes(bhw[0], bhw[1], (-3.0, 9.0, -9.0, 26.0))   # other bounds: other axes, same length
    xs3, ys3 = xs3.numpy(), ys3.numpy()
    assert _same(_planned(feats, K, Rt, xs3, ys3, img, "mean", plan), orc.warp_fuse(feats, K, Rt, xs3, ys3, img, "mean"))


def test_module_with_table_cache_matches_module_without_and_trains():
    import bevipm
    from bevipm import rig
    torch.manual_seed(0)
    B, V, C, fhw, bhw = 2, 7, 128, (27, 48), (24, 72)
    K, Rt = rig.look_at_rig(V, 0)
    K = K[None].expand(B, -1, -1, -1).contiguous().cuda()
    Rt = Rt[None].expand(B, -1, -1, -1).contiguous().cuda()
    plain = bevipm.FusedIPM(bhw[0], bhw[1], rig.WILDTRACK_BOUNDS, fusion="mean").cuda()
    cached = bevipm.FusedIPM(bhw[0], bhw[1], rig.WILDTRACK_BOUNDS, fusion="mean", cache_tables=True).cuda()
    for step in range(3):
        x = torch.randn(B, V, C, *fhw, device="cuda")
        x1, x2 = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
        a = plain(x1, K, Rt, img_size=rig.WILDTRACK_IMG_SIZE)
        b = cached(x2, K, Rt, img_size=rig.WILDTRACK_IMG_SIZE)
        assert torch.equal(a, b)
        a.square().sum().backward()
        b.square().sum().backward()
        assert torch.allclose(x1.grad, x2.grad, rtol=1e-5, atol=1e-6)
    assert len(cached._plans) == 1 and int(next(iter(cached._plans.values()))[:8].view(torch.int64).item()) != 0
    cached.reset_table_cache()
    assert int(next(iter(cached._plans.values()))[:8].view(torch.int64).item()) == 0
    gt = bevipm.GeometryTransformer(bhw[0], bhw[1], rig.WILDTRACK_BOUNDS, cache_tables=True).cuda()
    gp = bevipm.GeometryTransformer(bhw[0], bhw[1], rig.WILDTRACK_BOUNDS).cuda()
    x = torch.randn(B, V, C, *fhw, device="cuda")
    for _ in range(2):
        assert torch.equal(gt(x, K, Rt, img_size=rig.WILDTRACK_IMG_SIZE), gp(x, K, Rt, img_size=rig.WILDTRACK_IMG_SIZE))


def test_planned_entry_rejects_what_it_cannot_take():
    from bevipm import _lib, ops
    feats, K, Rt, xs, ys, img = _rig_case(1, 3, 128, (20, 33), (19, 45), seed=1)
    f, Kd, Rd, xd, yd = _dev_inputs(feats, K, Rt, xs, ys, False, torch.float32)      # NCHW: not the run kernel
    plan = ops.new_plan(3, (19, 45), "cuda")
    with pytest.raises(RuntimeError):
        ops.warp_fuse_planned(f, Kd, Rd, xd, yd, int(img[0]), int(img[1]), _lib.MEAN, False, plan)
    f, Kd, Rd, xd, yd = _dev_inputs(feats, K, Rt, xs, ys, True, torch.float32)
    with pytest.raises(RuntimeError):
        ops.warp_fuse_planned(f, Kd, Rd, xd, yd, int(img[0]), int(img[1]), _lib.MEAN, False, plan[:1024].contiguous())   # too small


@pytest.mark.parametrize("wl", ["c1", "c3"])
def test_planned_full_size(wl):
    """BASELINE configs 0 and 2 at full size through the cache (fill, then use), against the plain entry."""
    from bevipm import _lib, ops, rig
    w = rig.WORKLOADS[wl]
    K, Rt = rig.look_at_rig(w.views, 0)
    Kd = K[None].contiguous().cuda()
    Rd = Rt[None, :, :3, :].contiguous().cuda()
    xs, ys = rig.ground_axes(*w.bev_hw, w.bounds)
    xd, yd = xs.cuda(), ys.cuda()
    f = torch.randn(1, w.views, *w.feat_hw, w.channels, device="cuda").permute(0, 1, 4, 2, 3)
    plan = ops.new_plan(w.views, w.bev_hw, "cuda")
    want = ops.warp_fuse(f, Kd, Rd, xd, yd, w.img_size[0], w.img_size[1], _lib.MODES[w.fusion], False, 0)
    for _ in range(3):
        out = ops.warp_fuse_planned(f, Kd, Rd, xd, yd, w.img_size[0], w.img_size[1], _lib.MODES[w.fusion], False, plan)
        assert torch.equal(out, want)
