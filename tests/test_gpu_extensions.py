"""Round-2 additions around the fused path: validity-mask counts, mean over the views that see a cell, SimpleFusion's
backward on our kernels, max fusion under autograd, autocast behaviour, fake-tensor strides."""
import numpy as np
import pytest
import torch

from oracle import ipm_oracle as orc
from test_gpu_parity import DEV, _dev_inputs, _rig_case, _run, _same

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(2, 7, (31, 53), (37, 91)), (1, 3, (20, 33), (19, 45)), (1, 7, (135, 240), (120, 360))])
def test_valid_count_matches_oracle(shape):
    from bevipm import ops
    B, V, fhw, bhw = shape
    feats, K, Rt, xs, ys, img = _rig_case(B, V, 4, fhw, bhw, seed=3)
    f, Kd, Rd, xd, yd = _dev_inputs(feats, K, Rt, xs, ys, True)
    cnt = ops.valid_count(Kd, Rd, xd, yd, fhw, (int(img[0]), int(img[1])))
    want = orc.valid_count(K, Rt, xs, ys, fhw, img)
    assert cnt.dtype == torch.int32 and tuple(cnt.shape) == (B, *bhw)
    assert np.array_equal(cnt.cpu().numpy(), want)
    assert 0 < want.max() <= V and want.min() == 0 or want.min() >= 0


def test_valid_count_on_golden_geometry(golden):
    """The goldens' degenerate calibration: a view entirely out of frame, cells behind a camera, non-finite positions."""
    from bevipm import ops
    from test_gpu_parity import _bcast
    for case in ("degenerate", "non_finite", "w_guard"):
        z = golden(case)
        feats, K, Rt = _bcast(z)
        f, Kd, Rd, xd, yd = _dev_inputs(feats, K, Rt, z["xs"], z["ys"], False)
        cnt = ops.valid_count(Kd, Rd, xd, yd, feats.shape[-2:], tuple(int(v) for v in z["img_size"]))
        assert np.array_equal(cnt.cpu().numpy(), orc.valid_count(K, Rt, z["xs"], z["ys"], feats.shape[-2:], tuple(z["img_size"]))), case


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_mean_over_valid_views_and_count_output(dtype):
    """FusedIPM(fusion='mean_valid', return_valid=True): sum / max(count, 1), bit-exact against the oracle's restatement;
    the default 'mean' keeps dividing by V (fusion.py:20-21) and only gains the count output."""
    from bevipm import modules
    feats, K, Rt, xs, ys, img = _rig_case(2, 7, 136 if dtype == torch.float32 else 264, (31, 53), (37, 91), seed=9)
    f32 = torch.from_numpy(feats).to(dtype).float().numpy()
    fd = torch.from_numpy(feats).to(DEV).to(dtype)
    Kt, Rtt = torch.from_numpy(K).to(DEV), torch.from_numpy(Rt).to(DEV)
    m = modules.FusedIPM(37, 91, (-24.0, 24.0, -7.2, 7.2), fusion="mean_valid", return_valid=True, layout="channels_last")
    out, cnt = m(fd, Kt, Rtt, img_size=img)
    want, wcnt = orc.warp_fuse_mean_valid(f32, K, Rt, xs, ys, img)
    assert np.array_equal(cnt.cpu().numpy(), wcnt)
    assert _same(out.cpu().numpy(), want)
    m2 = modules.FusedIPM(37, 91, (-24.0, 24.0, -7.2, 7.2), fusion="mean", return_valid=True, layout="channels_last")
    out2, cnt2 = m2(fd, Kt, Rtt, img_size=img)
    assert _same(out2.cpu().numpy(), orc.warp_fuse(f32, K, Rt, xs, ys, img, "mean"))
    assert torch.equal(cnt2, cnt)
    # where every view sees the cell both means agree exactly; where some do not, mean_valid is the larger in magnitude
    full = (cnt == 7)[:, None].expand_as(out)
    assert torch.equal(out[full], out2[full])


@pytest.mark.parametrize("mode", ["sum", "mean", "max"])
@pytest.mark.parametrize("channels_last", [False, True])
def test_simple_fusion_backward_on_our_kernels(mode, channels_last):
    """SimpleFusion (fusion.py:11-22) forward AND backward run in libbevipm.so; gradients equal torch's own."""
    from bevipm import _lib, modules
    torch.manual_seed(5)
    x = torch.randn(2, 5, 24, 9, 13, device=DEV)
    x[0, 1] = x[0, 3]                                   # ties: torch.max sends the gradient to the first arg-max
    if channels_last:
        x = x.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)
    a = x.clone().requires_grad_(True)
    b = x.clone().requires_grad_(True)
    before = _lib.launch_count()
    ya = modules.SimpleFusion(mode)(a)
    cot = torch.randn_like(ya)
    (ya * cot).sum().backward()
    assert _lib.launch_count() - before >= 2            # forward + backward kernels of the library
    if mode == "max":
        yb = b.max(1).values
    else:   # the reference's CPU reduction adds the views in order; torch's CUDA sum re-associates
        yb = b[:, 0]
        for v in range(1, b.shape[1]):
            yb = yb + b[:, v]
        if mode == "mean":
            yb = yb / torch.tensor(float(b.shape[1]), device=DEV)
    (yb * cot).sum().backward()
    assert torch.equal(ya, yb)
    if mode == "max":
        # same total gradient per cell, all of it on one arg-max view (torch's CUDA tie-break is not specified: compare sums
        # and the untied cells)
        assert torch.equal(a.grad.sum(1), b.grad.sum(1))
        untied = torch.ones_like(a.grad, dtype=torch.bool)
        untied[0] = False
        assert torch.equal(a.grad[untied], b.grad[untied])
        assert torch.equal(a.grad[0, 3], torch.zeros_like(a.grad[0, 3]))   # the tie goes to the FIRST view (1), never to 3
    else:
        assert torch.equal(a.grad, b.grad)


def test_fused_max_is_differentiable():
    """FusedIPM(fusion='max') under autograd (fusion.py:22 is differentiable in the reference): per-view maps + our max
    reduction, gradient = the oracle's per-view backward applied to the arg-max routing."""
    from bevipm import modules
    feats, K, Rt, xs, ys, img = _rig_case(1, 4, 8, (20, 33), (19, 45), seed=12)
    fd = torch.from_numpy(feats).to(DEV).requires_grad_(True)
    Kt, Rtt = torch.from_numpy(K).to(DEV), torch.from_numpy(Rt).to(DEV)
    m = modules.FusedIPM(19, 45, (-24.0, 24.0, -7.2, 7.2), fusion="max", layout="keep")
    out = m(fd, Kt, Rtt, img_size=img)
    assert _same(out.detach().cpu().numpy(), orc.warp_fuse(feats, K, Rt, xs, ys, img, "max"))
    cot = torch.randn_like(out)
    (out * cot).sum().backward()
    # reference autograd on the oracle's per-view maps
    pv = torch.from_numpy(orc.warp_fuse(feats, K, Rt, xs, ys, img, "none")).to(DEV).requires_grad_(True)
    (pv.max(1).values * cot).sum().backward()
    want = orc.warp_fuse_bwd(pv.grad.cpu().numpy(), K, Rt, xs, ys, feats.shape, img, "none")
    got = fd.grad.cpu().numpy()
    assert np.abs(got - want).max() <= 1e-5 * max(np.abs(want).max(), 1e-30)


def test_autocast_fp16_forward_backward_like_grid_sampler():
    """train.py:238-247 runs the model under autocast(fp16): grid_sampler's autocast policy is fp32, so fp16 features are
    up-cast, the result is fp32 and equals the fp32 path on the up-cast features; the gradient comes back in fp16."""
    from bevipm import modules
    feats, K, Rt, xs, ys, img = _rig_case(2, 5, 64, (20, 33), (19, 45), seed=14)
    f16 = torch.from_numpy(feats).to(DEV).half().requires_grad_(True)
    Kt, Rtt = torch.from_numpy(K).to(DEV), torch.from_numpy(Rt).to(DEV)
    m = modules.FusedIPM(19, 45, (-24.0, 24.0, -7.2, 7.2), fusion="mean")
    with torch.autocast("cuda", dtype=torch.float16):
        out = m(f16, Kt, Rtt, img_size=img)
    assert out.dtype == torch.float32
    up = f16.detach().float().cpu().numpy()
    assert _same(out.detach().cpu().numpy(), orc.warp_fuse(up, K, Rt, xs, ys, img, "mean"))
    cot = torch.randn_like(out)
    (out * cot).sum().backward()
    assert f16.grad is not None and f16.grad.dtype == torch.float16
    want = orc.warp_fuse_bwd(cot.cpu().numpy(), K, Rt, xs, ys, feats.shape, img, "mean")
    assert np.abs(f16.grad.float().cpu().numpy() - want).max() <= 2e-3 * np.abs(want).max()   # one fp16 rounding of the gradient


@pytest.mark.parametrize("mode", ["mean", "none"])
def test_fake_kernels_report_the_real_strides(mode):
    """torch.compile / opcheck plan with the fake kernels: their strides must be the real ops' (channels-last results for
    channels-last features)."""
    from bevipm import _lib, ops
    feats, K, Rt, xs, ys, img = _rig_case(1, 3, 16, (20, 33), (19, 45), seed=15)
    for cl in (False, True):
        f, Kd, Rd, xd, yd = _dev_inputs(feats, K, Rt, xs, ys, cl)
        args = (f, Kd, Rd, xd, yd, int(img[0]), int(img[1]), _lib.MODES[mode], False, 0)
        real = ops.warp_fuse(*args)
        torch.library.opcheck(ops.warp_fuse, args, test_utils=("test_schema", "test_faketensor"))
        from torch._subclasses.fake_tensor import FakeTensorMode
        with FakeTensorMode() as fm:
            fake = ops.warp_fuse(*[fm.from_tensor(a) if isinstance(a, torch.Tensor) else a for a in args])
        assert tuple(fake.shape) == tuple(real.shape) and fake.stride() == real.stride()
