"""Parity of the TMA-staged fused kernel (csrc/ipm_staged.cuh, variants 50 / 51) against the oracle.

The kernel stages source tiles in shared memory through tensor-map TMA copies whose out-of-map parts are
zero-filled by the hardware: the reference's zero padding itself (geometry.py:161, padding_mode='zeros'), so
unlike the run kernel there is no documented deviation for non-finite features -- every comparison here is
bit-exact (tolerance 0 <= the 1e-5 the north star allows), NaN positions included.
"""
import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES
from oracle import ipm_oracle as orc
from test_gpu_parity import DEV, _bcast, _rig_case, _run, _same

pytestmark = pytest.mark.gpu

STAGED = [50, 51, 55]   # staged tiles: 8-row (2 CTAs / SM), 4-row (4 CTAs / SM); 55: run kernel with the TMA-box ring


@pytest.mark.parametrize("variant", STAGED)
@pytest.mark.parametrize("mode", ["mean", "sum", "max"])
def test_staged_fp32_ragged(variant, mode):
    # Hb, Wb not multiples of the tile; C = 136 leaves a partial 512-byte chunk (TMA zero-fills past C)
    feats, K, Rt, xs, ys, img = _rig_case(2, 7, 136, (31, 53), (37, 91), seed=2)
    want = orc.warp_fuse(feats, K, Rt, xs, ys, img, mode)
    out = _run(feats, K, Rt, xs, ys, img, mode, True, variant=variant).cpu().numpy()
    assert _same(out, want)
    from bevipm import _lib
    assert int(_lib.load().bevipm_last_variant()) == variant


@pytest.mark.parametrize("variant", STAGED)
@pytest.mark.parametrize("out_bf16", [False, True])
@pytest.mark.parametrize("mode", ["mean", "max"])
def test_staged_bf16_ragged(variant, out_bf16, mode):
    feats, K, Rt, xs, ys, img = _rig_case(2, 5, 776, (31, 53), (37, 91), seed=4)   # three whole bf16 chunks + 8 channels
    fb = torch.from_numpy(feats).bfloat16()
    want = torch.from_numpy(orc.warp_fuse(fb.float().numpy(), K, Rt, xs, ys, img, mode))
    out = _run(fb.float().numpy(), K, Rt, xs, ys, img, mode, True, dtype=torch.bfloat16, out_bf16=out_bf16,
               variant=variant).cpu()
    if out_bf16:
        assert out.dtype == torch.bfloat16
        assert torch.equal(out, want.bfloat16())
    else:
        assert _same(out.numpy(), want.numpy())


@pytest.mark.parametrize("variant", STAGED)
@pytest.mark.parametrize("case", GOLDEN_CASES)
@pytest.mark.parametrize("mode", ["mean", "max"])
def test_staged_on_golden_geometry(golden, case, variant, mode):
    """The reference goldens' calibration (views out of frame, cells behind a camera, the |w| guard, non-finite
    sample positions) at the goldens' own channel count (a single partial chunk)."""
    z = golden(case)
    feats, K, Rt = _bcast(z)
    C0 = feats.shape[2]
    Ct = -(-C0 // 4) * 4   # one 16-byte vector of channels at least (fast-path rule); channel c = golden channel c % C0
    tiled = np.ascontiguousarray(np.tile(feats, (1, 1, 4, 1, 1))[:, :, :Ct])
    want = orc.warp_fuse(tiled, K, Rt, z["xs"], z["ys"], tuple(z["img_size"]), mode)
    out = _run(tiled, K, Rt, z["xs"], z["ys"], z["img_size"], mode, True, variant=variant).cpu().numpy()
    assert _same(out, want)


@pytest.mark.parametrize("fpc", [1, 2, 3, 8])
@pytest.mark.parametrize("variant", STAGED)
def test_staged_frame_runs_share_tables_only_when_calibration_repeats(monkeypatch, fpc, variant):
    monkeypatch.setenv("BEVIPM_RUN_FPC", str(fpc))
    from bevipm import rig
    B, V, C, fhw, bhw = 5, 4, 256, (20, 33), (19, 45)
    feats = torch.randn(B, V, C, *fhw, generator=torch.Generator().manual_seed(21)).numpy()
    K = np.zeros((B, V, 3, 3), np.float32)
    Rt = np.zeros((B, V, 4, 4), np.float32)
    for b, seed in enumerate([3, 3, 4, 5, 5]):
        k, r = rig.look_at_rig(V, seed)
        K[b], Rt[b] = k.numpy(), r.numpy()
    Rt[4, 2, 0, 3] = np.nextafter(Rt[4, 2, 0, 3], np.float32(np.inf))
    xs, ys = rig.ground_axes(bhw[0], bhw[1], rig.WILDTRACK_BOUNDS)
    want = orc.warp_fuse(feats, K, Rt, xs.numpy(), ys.numpy(), rig.WILDTRACK_IMG_SIZE, "mean")
    out = _run(feats, K, Rt, xs.numpy(), ys.numpy(), rig.WILDTRACK_IMG_SIZE, "mean", True, variant=variant).cpu().numpy()
    assert _same(out, want)
    fb = torch.from_numpy(feats).bfloat16().float().numpy()
    wantb = orc.warp_fuse(fb, K, Rt, xs.numpy(), ys.numpy(), rig.WILDTRACK_IMG_SIZE, "sum")
    outb = _run(fb, K, Rt, xs.numpy(), ys.numpy(), rig.WILDTRACK_IMG_SIZE, "sum", True, dtype=torch.bfloat16,
                variant=variant).cpu().numpy()
    assert _same(outb, wantb)


@pytest.mark.parametrize("variant", STAGED)
def test_staged_non_finite_features_match_the_reference_exactly(variant):
    """Inf / NaN features, including texels on the map border next to out-of-map taps: TMA zero fill is the
    reference's zero padding (0 * w = 0, never Inf * 0), so the result equals the oracle everywhere."""
    feats, K, Rt, xs, ys, img = _rig_case(1, 3, 128, (20, 33), (19, 45), seed=31)
    feats[0, 1, 5, 7, 11] = np.inf
    feats[0, 2, 9, 3, 20] = -np.inf
    feats[0, 0, 64, 10, 10] = np.nan
    feats[0, :, 3, 0, :] = np.inf       # the whole top border row of channel 3
    feats[0, :, 4, :, 0] = -np.inf      # the left border column of channel 4
    feats[0, :, 5, -1, -1] = np.inf
    for mode in ("mean", "sum", "max"):
        want = orc.warp_fuse(feats, K, Rt, xs, ys, img, mode)
        out = _run(feats, K, Rt, xs, ys, img, mode, True, variant=variant).cpu().numpy()
        assert np.isinf(want).any()
        assert _same(out, want), mode


@pytest.mark.parametrize("variant", STAGED)
@pytest.mark.parametrize("C,dtype", [(136, torch.float32), (64, torch.float32), (4, torch.float32), (264, torch.bfloat16), (8, torch.bfloat16)])
def test_staged_partial_channel_chunks(variant, C, dtype):
    feats, K, Rt, xs, ys, img = _rig_case(2, 5, C, (31, 53), (37, 91), seed=41)
    f = torch.from_numpy(feats).to(dtype).float().numpy()
    want = orc.warp_fuse(f, K, Rt, xs, ys, img, "mean")
    out = _run(f, K, Rt, xs, ys, img, "mean", True, dtype=dtype, variant=variant).cpu().numpy()
    assert _same(out, want)


@pytest.mark.parametrize("views", [1, 2, 9, 16, 17, 32])
def test_staged_view_counts(views):
    feats, K, Rt, xs, ys, img = _rig_case(1, views, 128, (20, 33), (19, 45), seed=10 + views)
    want = orc.warp_fuse(feats, K, Rt, xs, ys, img, "mean")
    for variant in STAGED:
        out = _run(feats, K, Rt, xs, ys, img, "mean", True, variant=variant).cpu().numpy()
        assert _same(out, want), variant


@pytest.mark.parametrize("variant", STAGED)
@pytest.mark.parametrize("fhw,bhw", [((120, 200), (10, 30)),     # BEV far coarser than the source: one 2x2 box per cell
                                     ((64, 96), (33, 70)),       # mixed: tile stages, row stages and boxes
                                     ((9, 12), (60, 150)),       # BEV far denser than a tiny source map
                                     ((200, 300), (24, 64))])
@pytest.mark.parametrize("ring,cap,lag", [(32768, 16384, 1), (49152, 16384, 2), (65536, 32768, 3), (98304, 98304, 2)])
def test_staged_every_stage_kind(monkeypatch, variant, fhw, bhw, ring, cap, lag):
    """Footprints from far smaller to far larger than the largest stage: views are staged per tile, per group of BEV
    rows, per row or per 2x2 block (ipm_staged.cuh, phase A); ring size, stage cap and arming lag move the
    boundaries and the copy schedule."""
    monkeypatch.setenv("BEVIPM_ST_RING", str(ring))
    monkeypatch.setenv("BEVIPM_ST_CAP", str(cap))
    monkeypatch.setenv("BEVIPM_ST_LAG", str(lag))
    feats, K, Rt, xs, ys, img = _rig_case(1, 7, 128, fhw, bhw, seed=sum(fhw) + sum(bhw))
    want = orc.warp_fuse(feats, K, Rt, xs, ys, img, "mean")
    out = _run(feats, K, Rt, xs, ys, img, "mean", True, variant=variant).cpu().numpy()
    assert _same(out, want)


@pytest.mark.parametrize("seed", range(16))
def test_staged_random_shapes(seed):
    rng = np.random.RandomState(2000 + seed)
    B, V = int(rng.randint(1, 4)), int(rng.choice([1, 2, 3, 5, 7, 8, 9, 16, 17]))
    bf16 = bool(rng.randint(0, 2))
    C = int(rng.choice([256, 512, 264] if bf16 else [128, 256, 384, 132]))
    fhw = (int(rng.randint(8, 80)), int(rng.randint(8, 120)))
    bhw = (int(rng.randint(5, 50)), int(rng.randint(5, 100)))
    mode = str(rng.choice(["mean", "sum", "max"]))
    variant = int(rng.choice(STAGED))
    feats, K, Rt, xs, ys, img = _rig_case(B, V, C, fhw, bhw, seed=seed)
    if bf16:
        feats = torch.from_numpy(feats).bfloat16().float().numpy()
    want = orc.warp_fuse(feats, K, Rt, xs, ys, img, mode)
    out = _run(feats, K, Rt, xs, ys, img, mode, True, dtype=torch.bfloat16 if bf16 else torch.float32, variant=variant).cpu().numpy()
    assert _same(out, want), (B, V, C, fhw, bhw, mode, bf16, variant)


def test_staged_strided_window_of_a_larger_tensor():
    """The tensor maps carry the caller's strides: a [.., 4:24, 3:36, :] window of a larger channels-last tensor."""
    from bevipm import _lib, ops, modules
    feats, K, Rt, xs, ys, img = _rig_case(2, 3, 128, (20, 33), (19, 45), seed=77)
    big = torch.full((2, 3, 28, 40, 160), float("nan"), device=DEV)
    big[:, :, 4:24, 3:36, 16:144] = torch.from_numpy(feats).to(DEV).permute(0, 1, 3, 4, 2)
    win = big[:, :, 4:24, 3:36, 16:144].permute(0, 1, 4, 2, 3)
    Kd, Rd = modules.pack_calibration(torch.from_numpy(K), torch.from_numpy(Rt), 2, 3, torch.device(DEV))
    xd, yd = torch.from_numpy(xs).to(DEV), torch.from_numpy(ys).to(DEV)
    want = orc.warp_fuse(feats, K, Rt, xs, ys, img, "mean")
    for variant in STAGED:
        out = ops.warp_fuse(win, Kd, Rd, xd, yd, int(img[0]), int(img[1]), _lib.MEAN, False, variant)
        assert _same(out.cpu().numpy(), want), variant


@pytest.mark.parametrize("name", ["c1", "c3"])
@pytest.mark.parametrize("variant", STAGED)
def test_staged_full_size_fp32_vs_oracle(name, variant):
    from bevipm import rig
    wl = rig.WORKLOADS[name]
    feats, K, Rt, xs, ys, img = _rig_case(1, wl.views, wl.channels, wl.feat_hw, wl.bev_hw, seed=0)
    nhwc = np.ascontiguousarray(feats.transpose(0, 1, 3, 4, 2)).transpose(0, 1, 4, 2, 3)
    want = orc.warp_fuse(nhwc, K, Rt, xs, ys, img, "mean", channels_last_out=True)
    out = _run(feats, K, Rt, xs, ys, img, "mean", True, variant=variant).cpu().numpy()
    assert _same(out, want)


@pytest.mark.parametrize("variant", [0] + STAGED)
def test_staged_full_size_c2_all_frames_vs_oracle(variant):
    """BASELINE config 1 (the bench workload), every one of the 8 frames against the oracle, bit-exact, plus the
    bf16 output = rounding of the fp32 output.  variant 0 = the default dispatch (run kernel)."""
    from bevipm import _lib, ops, rig
    wl = rig.WORKLOADS["c2"]
    B, V, C = wl.frames, wl.views, wl.channels
    g = torch.Generator(device=DEV).manual_seed(0)
    f = torch.randn(B, V, *wl.feat_hw, C, device=DEV, generator=g).bfloat16().permute(0, 1, 4, 2, 3)
    K, Rt = rig.look_at_rig(V, 0)
    Kd = K[None].expand(B, -1, -1, -1).contiguous().to(DEV)
    Rd = Rt[None, :, :3, :].expand(B, -1, -1, -1).contiguous().to(DEV)
    xs, ys = rig.ground_axes(*wl.bev_hw, wl.bounds)
    xd, yd = xs.to(DEV), ys.to(DEV)
    img = wl.img_size
    out = ops.warp_fuse(f, Kd, Rd, xd, yd, img[0], img[1], _lib.MEAN, False, variant)
    outb = ops.warp_fuse(f, Kd, Rd, xd, yd, img[0], img[1], _lib.MEAN, True, variant)
    assert torch.equal(outb, out.bfloat16())
    for b in range(B):
        fb = f[b:b + 1].float().cpu().numpy()
        want = orc.warp_fuse(fb, K[None].numpy(), Rt[None].numpy(), xs.numpy(), ys.numpy(), img, "mean")
        assert _same(out[b:b + 1].cpu().numpy(), want), b


def test_full_size_c5_clip_every_frame_vs_oracle():
    """BASELINE config 4 (c5): a 64-frame clip of c1's shape (7 views x 512 ch fp32) in ONE launch of the default kernel,
    every frame against the oracle, bit-exact.  Frames 20..29 use another rig and frame 40 a one-bit-different extrinsic:
    the frame groups must rebuild their tables exactly there."""
    from bevipm import _lib, ops, rig
    wl = rig.WORKLOADS["c5"]
    B, V, C = wl.frames, wl.views, wl.channels
    g = torch.Generator(device=DEV).manual_seed(7)
    f = torch.empty((B, V, *wl.feat_hw, C), device=DEV)
    for b in range(B):
        f[b] = torch.randn((V, *wl.feat_hw, C), device=DEV, generator=g)
    f = f.permute(0, 1, 4, 2, 3)
    K0, R0 = rig.look_at_rig(V, 0)
    K1, R1 = rig.look_at_rig(V, 1)
    K = K0[None].repeat(B, 1, 1, 1)
    Rt = R0[None].repeat(B, 1, 1, 1)
    K[20:30], Rt[20:30] = K1, R1
    Rt[40, 3, 1, 3] = torch.nextafter(Rt[40, 3, 1, 3], torch.tensor(float("inf")))
    Kd, Rd = K.contiguous().to(DEV), Rt[:, :, :3, :].contiguous().to(DEV)
    xs, ys = rig.ground_axes(*wl.bev_hw, wl.bounds)
    out = ops.warp_fuse(f, Kd, Rd, xs.to(DEV), ys.to(DEV), wl.img_size[0], wl.img_size[1], _lib.MEAN, False, 0)
    torch.cuda.synchronize()
    assert out.shape == (B, C, *wl.bev_hw)
    for b in range(B):
        fb = f[b:b + 1].cpu().numpy()                      # channels-last strides travel with the array
        want = orc.warp_fuse(fb, K[b:b + 1].numpy(), Rt[b:b + 1].numpy(), xs.numpy(), ys.numpy(), wl.img_size, "mean", channels_last_out=True)
        assert _same(out[b:b + 1].cpu().numpy(), want), b
