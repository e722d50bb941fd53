"""The CPU oracle must reproduce the reference module's own outputs bit for bit.

Golden vectors: tests/golden/*.npz, produced by tests/golden/make_golden.py from
/root/reference/project/models/fusion/{geometry,fusion}.py.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN_CASES
from oracle import ipm_oracle as orc


def _broadcast_calib(z):
    feats = z["feats"]
    B, V = feats.shape[:2]
    K = np.broadcast_to(z["K"], (B, V) + z["K"].shape[-2:])
    Rt = np.broadcast_to(z["Rt"], (B, V) + z["Rt"].shape[-2:])
    return feats, K, Rt


def _same(a, b):
    """Bit-exact up to NaN payload: same NaN positions, identical everywhere else."""
    na, nb = np.isnan(a), np.isnan(b)
    return np.array_equal(na, nb) and np.array_equal(a[~na], b[~nb])


def _same_outside_sum_tail(a, b):
    """Fused sum/mean: bit-exact, except where ATen's CPU sum kernel itself changes its order.

    aten::sum over the view axis adds sequentially (v0+v1)+v2+... for blocks of 4 SIMD vectors
    (64 floats with AVX-512) and switches to four interleaved partial sums for the leftover
    (< 64) elements at the end of each frame's C*Hb*Wb run (SumKernel.cpp row_sum, ilp_factor 4).
    That tail order depends on the host's vector width and thread split, not on the reference's
    algorithm, so the oracle keeps the sequential order everywhere; tail elements may differ by
    an ulp.  C*Hb*Wb is a multiple of 64 at every BASELINE.json shape, i.e. no tail there.
    """
    na, nb = np.isnan(a), np.isnan(b)
    if not np.array_equal(na, nb):
        return False
    B = a.shape[0]
    fa, fb = a.reshape(B, -1), b.reshape(B, -1)
    inner = fa.shape[1]
    body = (inner // 64) * 64
    ok_body = np.array_equal(fa[:, :body][~na.reshape(B, -1)[:, :body]], fb[:, :body][~nb.reshape(B, -1)[:, :body]])
    ta, tb = fa[:, body:], fb[:, body:]
    fin = ~np.isnan(ta)
    ok_tail = np.all(np.abs(ta[fin] - tb[fin]) <= 4 * np.spacing(np.abs(tb[fin]).astype(np.float32)) + 1e-7)
    return ok_body and ok_tail


@pytest.mark.parametrize("case", GOLDEN_CASES)
@pytest.mark.parametrize("mode", ["none", "sum", "mean", "max"])
def test_oracle_matches_reference_golden(golden, case, mode):
    z = golden(case)
    feats, K, Rt = _broadcast_calib(z)
    out = orc.warp_fuse(feats, K, Rt, z["xs"], z["ys"], tuple(z["img_size"]), mode)
    assert out.shape == z["out_" + mode].shape
    if mode in ("sum", "mean"):
        assert _same_outside_sum_tail(out, z["out_" + mode])
    else:
        assert _same(out, z["out_" + mode])


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_concat_is_view_major_per_view(golden, case):
    # fusion.py:43-46 -- channel index v*C + c
    z = golden(case)
    B, V, C, Hb, Wb = z["out_none"].shape
    assert _same(z["out_concat"], z["out_none"].reshape(B, V * C, Hb, Wb))


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_oracle_strided_layouts_agree(golden, case):
    # channels-last features / channels-last output are the layouts the CUDA path uses
    z = golden(case)
    feats, K, Rt = _broadcast_calib(z)
    nhwc = np.ascontiguousarray(feats.transpose(0, 1, 3, 4, 2)).transpose(0, 1, 4, 2, 3)
    assert nhwc.strides != feats.strides
    out = orc.warp_fuse(nhwc, K, Rt, z["xs"], z["ys"], tuple(z["img_size"]), "mean", channels_last_out=True)
    assert _same(out, orc.warp_fuse(feats, K, Rt, z["xs"], z["ys"], tuple(z["img_size"]), "mean"))
    out = orc.warp_fuse(nhwc, K, Rt, z["xs"], z["ys"], tuple(z["img_size"]), "none", channels_last_out=True)
    assert _same(out, z["out_none"])


@pytest.mark.parametrize("case", ["rig_small", "seven_views_3x4", "degenerate", "w_guard"])
@pytest.mark.parametrize("mode", ["sum", "mean", "none"])
def test_oracle_backward_matches_autograd(golden, case, mode):
    z = golden(case)
    feats, K, Rt = _broadcast_calib(z)
    cot = z["cotangent_none"] if mode == "none" else z["cotangent"]
    g = orc.warp_fuse_bwd(cot, K, Rt, z["xs"], z["ys"], feats.shape, tuple(z["img_size"]), mode)
    ref = z["grad_" + mode]
    # torch's backward accumulates in a different order: tolerance, not bit-exact
    scale = np.abs(ref).max()
    assert np.abs(g - ref).max() <= 1e-5 * scale


def test_oracle_threads_do_not_change_bits(golden):
    z = golden("rig_small")
    feats, K, Rt = _broadcast_calib(z)
    a = orc.warp_fuse(feats, K, Rt, z["xs"], z["ys"], tuple(z["img_size"]), "mean", nthreads=1)
    b = orc.warp_fuse(feats, K, Rt, z["xs"], z["ys"], tuple(z["img_size"]), "mean", nthreads=0)
    assert np.array_equal(a, b)


def test_oracle_homography_golden(golden):
    # geometry.py:33-64 on the rig: H = K [r1 r2 t], k-ordered fma chain
    z = golden("rig_small")
    K, Rt = z["K"][0, 0], z["Rt"][0, 0]
    H = orc.homography(K, Rt)
    G = np.stack([Rt[:3, 0], Rt[:3, 1], Rt[:3, 3]], axis=1).astype(np.float64)
    assert np.allclose(H, K.astype(np.float64) @ G, rtol=1e-6, atol=1e-4)


@pytest.mark.skipif(not os.path.isdir("/root/reference/project"), reason="reference tree not mounted")
@pytest.mark.parametrize("shape", [(1, 7, 8, (135, 240), (120, 360)), (1, 7, 4, (270, 480), (480, 1440))])
def test_oracle_vs_live_reference_full_spatial_size(shape):
    """BASELINE configs 1 and 3 at full spatial size (fewer channels), against the reference run live."""
    import sys
    import torch
    sys.path.insert(0, "/root/reference/project")
    from models.fusion.geometry import GeometryTransformer
    from models.fusion.fusion import SimpleFusion
    from bevipm import rig
    B, V, C, fhw, bhw = shape
    K, Rt = rig.look_at_rig(V, 0)
    K, Rt = K[None].contiguous(), Rt[None].contiguous()
    feats = torch.randn(B, V, C, *fhw, generator=torch.Generator().manual_seed(0))
    geom = GeometryTransformer(bhw[0], bhw[1], rig.WILDTRACK_BOUNDS, "kornia")
    with torch.no_grad():
        ref = SimpleFusion("mean")(geom(feats, K, Rt, img_size=rig.WILDTRACK_IMG_SIZE)).numpy()
    xs, ys = rig.ground_axes(bhw[0], bhw[1], rig.WILDTRACK_BOUNDS)
    out = orc.warp_fuse(feats.numpy(), K.numpy(), Rt.numpy(), xs.numpy(), ys.numpy(),
                        rig.WILDTRACK_IMG_SIZE, "mean")
    assert np.array_equal(out, ref)


@pytest.mark.parametrize("case", ["rig_small", "seven_views_3x4", "degenerate", "w_guard"])
def test_valid_count_restatement_agrees_with_the_reference_per_view_maps(golden, case):
    """The validity-count restatement (oracle.valid_count) against the REFERENCE's own per-view maps (golden out_none, made
    by the imported reference module): a view counts for a cell iff the reference's map can be non-zero there, i.e. iff some
    bilinear tap lies inside the feature map.  With random features a seen cell is non-zero in some channel."""
    z = golden(case)
    feats = z["feats"]
    B, V = feats.shape[:2]
    K = np.ascontiguousarray(np.broadcast_to(z["K"], (B, V) + z["K"].shape[-2:]))
    Rt = np.ascontiguousarray(np.broadcast_to(z["Rt"], (B, V) + z["Rt"].shape[-2:]))
    cnt = orc.valid_count(K, Rt, z["xs"], z["ys"], feats.shape[-2:], tuple(z["img_size"]))
    nonzero_views = (z["out_none"] != 0).any(axis=2).sum(axis=1)           # [B,Hb,Wb] from the reference's output
    assert cnt.shape == nonzero_views.shape
    assert (cnt >= nonzero_views).all()                                    # an unseen view is exactly zero in the reference
    assert (cnt == nonzero_views).mean() > 0.995                           # (a seen cell can still blend to 0: weight 0 taps)
    mv, c2 = orc.warp_fuse_mean_valid(feats, K, Rt, z["xs"], z["ys"], tuple(z["img_size"]))
    s = orc.warp_fuse(feats, K, Rt, z["xs"], z["ys"], tuple(z["img_size"]), "sum")
    full = (c2 == V)[:, None].repeat(feats.shape[2], axis=1)
    assert np.array_equal(mv[full], orc.warp_fuse(feats, K, Rt, z["xs"], z["ys"], tuple(z["img_size"]), "mean")[full])
    assert np.array_equal(mv[~full & (np.broadcast_to(c2[:, None], s.shape) <= 1)], s[~full & (np.broadcast_to(c2[:, None], s.shape) <= 1)])
