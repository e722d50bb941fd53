"""Parity of the CUDA path (through the C ABI) against the oracle and the reference's goldens.

fp32: bit-exact (numeric equality, NaN positions equal) -- tolerance 0, well inside the <= 1e-5
the north star allows.  bf16 features: the fp32 result is bit-exact against the oracle fed the
up-cast features (what CUDA autocast does for grid_sampler); a bf16 OUTPUT is that result rounded
to nearest-even, so the bound vs fp32 is one bf16 rounding, 2^-8 relative (<= 1e-2).
"""
import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES
from oracle import ipm_oracle as orc

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _same(a: np.ndarray, b: np.ndarray) -> bool:
    na, nb = np.isnan(a), np.isnan(b)
    return a.shape == b.shape and np.array_equal(na, nb) and np.array_equal(a[~na], b[~nb])


def _bcast(z):
    feats = z["feats"]
    B, V = feats.shape[:2]
    K = np.ascontiguousarray(np.broadcast_to(z["K"], (B, V) + z["K"].shape[-2:]))
    Rt = np.ascontiguousarray(np.broadcast_to(z["Rt"], (B, V) + z["Rt"].shape[-2:]))
    return feats, K, Rt


def _dev_inputs(feats, K, Rt, xs, ys, channels_last, dtype=torch.float32):
    from bevipm import modules
    f = torch.from_numpy(np.ascontiguousarray(feats)).to(DEV).to(dtype)
    if channels_last:
        f = f.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)
    B, V = f.shape[:2]
    Kd, Rd = modules.pack_calibration(torch.from_numpy(K), torch.from_numpy(Rt), B, V, torch.device(DEV))
    return f, Kd, Rd, torch.from_numpy(np.asarray(xs, np.float32)).to(DEV), torch.from_numpy(np.asarray(ys, np.float32)).to(DEV)


def _run(feats, K, Rt, xs, ys, img_size, mode, channels_last, dtype=torch.float32, out_bf16=False, variant=0):
    from bevipm import _lib, ops
    f, Kd, Rd, xd, yd = _dev_inputs(feats, K, Rt, xs, ys, channels_last, dtype)
    out = ops.warp_fuse(f, Kd, Rd, xd, yd, int(img_size[0]), int(img_size[1]), _lib.MODES[mode], out_bf16, variant)
    torch.cuda.synchronize()
    return out


def _rig_case(B, V, C, fhw, bhw, seed=0):
    from bevipm import rig
    K, Rt = rig.look_at_rig(V, seed)
    K = np.ascontiguousarray(np.broadcast_to(K.numpy(), (B, V, 3, 3)))
    Rt = np.ascontiguousarray(np.broadcast_to(Rt.numpy(), (B, V, 4, 4)))
    feats = torch.randn(B, V, C, *fhw, generator=torch.Generator().manual_seed(seed)).numpy()
    xs, ys = rig.ground_axes(bhw[0], bhw[1], rig.WILDTRACK_BOUNDS)
    return feats, K, Rt, xs.numpy(), ys.numpy(), rig.WILDTRACK_IMG_SIZE


# ---- golden vectors of the reference module ------------------------------------------------------

@pytest.mark.parametrize("case", GOLDEN_CASES)
@pytest.mark.parametrize("mode", ["none", "sum", "mean", "max"])
@pytest.mark.parametrize("channels_last", [False, True])
def test_golden_fp32(golden, case, mode, channels_last):
    z = golden(case)
    feats, K, Rt = _bcast(z)
    out = _run(feats, K, Rt, z["xs"], z["ys"], z["img_size"], mode, channels_last).cpu().numpy()
    # the oracle is pinned to these goldens bit-for-bit (tests/test_oracle.py); sum/mean goldens carry
    # ATen's vector-tail reordering, so the fused modes are compared with the oracle's sequential order
    want = z["out_none"] if mode == "none" else orc.warp_fuse(feats, K, Rt, z["xs"], z["ys"], tuple(z["img_size"]), mode)
    assert _same(out, want)
    if mode == "max":
        assert _same(out, z["out_max"])


@pytest.mark.parametrize("case", ["rig_small", "seven_views_3x4", "degenerate", "w_guard"])
@pytest.mark.parametrize("mode", ["sum", "mean", "none"])
@pytest.mark.parametrize("channels_last", [False, True])
def test_golden_backward(golden, case, mode, channels_last):
    from bevipm import _lib, ops
    z = golden(case)
    feats, K, Rt = _bcast(z)
    f, Kd, Rd, xd, yd = _dev_inputs(feats, K, Rt, z["xs"], z["ys"], channels_last)
    f.requires_grad_(True)
    out = ops.warp_fuse(f, Kd, Rd, xd, yd, int(z["img_size"][0]), int(z["img_size"][1]), _lib.MODES[mode], False, 0)
    cot = torch.from_numpy(z["cotangent_none"] if mode == "none" else z["cotangent"]).to(DEV)
    (out * cot).sum().backward()
    got = f.grad.cpu().numpy()
    ref = z["grad_" + mode]
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max()   # atomics: order differs, tolerance 1e-5 rel


@pytest.mark.parametrize("mode", ["sum", "mean", "none"])
@pytest.mark.parametrize("generic", [False, True])
def test_backward_run_kernel_and_generic_kernel_vs_oracle(monkeypatch, mode, generic):
    """Gradient w.r.t. the features on a shape the run-kernel backward takes (channels-last, C = 256): both the
    run-kernel backward (block-wise pre-reduction, then atomics) and the generic one must match the oracle's
    autograd restatement within 1e-5 relative (sums are re-associated by the atomics anyway)."""
    from bevipm import _lib, ops
    monkeypatch.setenv("BEVIPM_BWD_GENERIC", "1" if generic else "0")
    feats, K, Rt, xs, ys, img = _rig_case(2, 7, 256, (31, 53), (37, 91), seed=29)
    f, Kd, Rd, xd, yd = _dev_inputs(feats, K, Rt, xs, ys, True)
    f.requires_grad_(True)
    out = ops.warp_fuse(f, Kd, Rd, xd, yd, int(img[0]), int(img[1]), _lib.MODES[mode], False, 0)
    torch.manual_seed(3)
    cot = torch.randn_like(out)   # same (channels-last) strides as the output: what the fast paths take
    (out * cot).sum().backward()
    got = f.grad.cpu().numpy()
    want = orc.warp_fuse_bwd(np.ascontiguousarray(cot.cpu().numpy()), K, Rt, xs, ys, feats.shape, img, mode)
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()
    assert np.array_equal(got == 0, want == 0) or np.abs(got[(got == 0) != (want == 0)]).max() <= 1e-5 * np.abs(want).max()


# ---- every fused-kernel variant against the oracle, ragged shapes --------------------------------

@pytest.mark.parametrize("variant", list(range(1, 11)) + list(range(20, 28)))
@pytest.mark.parametrize("mode", ["mean", "sum"])
def test_variants_fp32_ragged(variant, mode):
    # Hb, Wb not multiples of any patch; C = 136 leaves a partial channel chunk
    feats, K, Rt, xs, ys, img = _rig_case(2, 5, 136, (31, 53), (37, 91), seed=2)
    want = orc.warp_fuse(feats, K, Rt, xs, ys, img, mode)
    out = _run(feats, K, Rt, xs, ys, img, mode, True, variant=variant).cpu().numpy()
    assert _same(out, want)


@pytest.mark.parametrize("variant", list(range(1, 11)) + list(range(20, 28)))
@pytest.mark.parametrize("out_bf16", [False, True])
def test_variants_bf16_ragged(variant, out_bf16):
    feats, K, Rt, xs, ys, img = _rig_case(2, 5, 264, (31, 53), (37, 91), seed=4)
    fb = torch.from_numpy(feats).bfloat16()
    want = torch.from_numpy(orc.warp_fuse(fb.float().numpy(), K, Rt, xs, ys, img, "mean"))
    out = _run(fb.float().numpy(), K, Rt, xs, ys, img, "mean", True, dtype=torch.bfloat16, out_bf16=out_bf16,
               variant=variant).cpu()
    if out_bf16:
        assert out.dtype == torch.bfloat16
        assert torch.equal(out, want.bfloat16())
        rel = (out.float() - want).abs().max() / want.abs().max()
        assert rel <= 1e-2
    else:
        assert _same(out.numpy(), want.numpy())


RUN_VARIANTS = list(range(30, 40))   # run kernel: {cells per segment, warps per segment, min CTAs/SM}


@pytest.mark.parametrize("variant", RUN_VARIANTS)
@pytest.mark.parametrize("mode", ["mean", "sum"])
def test_run_variants_fp32_ragged(variant, mode):
    # Hb, Wb not multiples of any segment; C = 256: two whole 512-byte chunks (what the run kernel needs)
    feats, K, Rt, xs, ys, img = _rig_case(2, 7, 256, (31, 53), (37, 91), seed=2)
    want = orc.warp_fuse(feats, K, Rt, xs, ys, img, mode)
    out = _run(feats, K, Rt, xs, ys, img, mode, True, variant=variant).cpu().numpy()
    assert _same(out, want)


@pytest.mark.parametrize("variant", RUN_VARIANTS)
@pytest.mark.parametrize("out_bf16", [False, True])
def test_run_variants_bf16_ragged(variant, out_bf16):
    feats, K, Rt, xs, ys, img = _rig_case(2, 5, 768, (31, 53), (37, 91), seed=4)   # three bf16 chunks
    fb = torch.from_numpy(feats).bfloat16()
    want = torch.from_numpy(orc.warp_fuse(fb.float().numpy(), K, Rt, xs, ys, img, "mean"))
    out = _run(fb.float().numpy(), K, Rt, xs, ys, img, "mean", True, dtype=torch.bfloat16, out_bf16=out_bf16,
               variant=variant).cpu()
    if out_bf16:
        assert torch.equal(out, want.bfloat16())
    else:
        assert _same(out.numpy(), want.numpy())


@pytest.mark.parametrize("variant", [30, 33, 35, 37])
@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_run_kernel_on_golden_geometry(golden, case, variant):
    """The reference goldens' calibration (views out of frame, cells behind a camera, the |w| guard,
    non-finite sample positions) with the features tiled up to one whole channel chunk: channel c of
    the result must equal channel c % C0 of the oracle's."""
    z = golden(case)
    feats, K, Rt = _bcast(z)
    C0 = feats.shape[2]
    reps = -(-128 // C0)
    tiled = np.ascontiguousarray(np.tile(feats, (1, 1, reps, 1, 1))[:, :, :128])
    want = orc.warp_fuse(tiled, K, Rt, z["xs"], z["ys"], tuple(z["img_size"]), "mean")
    out = _run(tiled, K, Rt, z["xs"], z["ys"], z["img_size"], "mean", True, variant=variant).cpu().numpy()
    assert _same(out, want)
    assert _same(out[:, :C0], orc.warp_fuse(feats, K, Rt, z["xs"], z["ys"], tuple(z["img_size"]), "mean"))


@pytest.mark.parametrize("fpc", [1, 2, 3, 8])
@pytest.mark.parametrize("variant", [31, 33, 34])
def test_run_kernel_frame_runs_share_tables_only_when_calibration_repeats(monkeypatch, fpc, variant):
    """One CTA walks `fpc` frames and keeps its tap tables while the calibration repeats bit for bit: frames
    0-1 share one rig, 2 has another, 3-4 a third (and 4 differs from 3 in a single extrinsic bit)."""
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
    if variant == 33:   # the other fusion modes of the run kernel walk the same shared tables (default dispatch)
        for mode in ("max", "none"):
            wantm = orc.warp_fuse(fb, K, Rt, xs.numpy(), ys.numpy(), rig.WILDTRACK_IMG_SIZE, mode)
            outm = _run(fb, K, Rt, xs.numpy(), ys.numpy(), rig.WILDTRACK_IMG_SIZE, mode, True, dtype=torch.bfloat16).cpu().numpy()
            assert _same(outm, wantm), mode


def test_run_kernel_non_finite_features_take_the_exact_division():
    """+-Inf sums make the 3-op mean division return NaN; the per-chunk finiteness test must route them to
    the IEEE division (Inf / V = Inf), exactly like the oracle."""
    feats, K, Rt, xs, ys, img = _rig_case(1, 3, 128, (20, 33), (19, 45), seed=31)
    feats[0, 1, 5, 7, 11] = np.inf
    feats[0, 2, 9, 3, 20] = -np.inf
    feats[0, 0, 64, 10, 10] = np.nan
    want = orc.warp_fuse(feats, K, Rt, xs, ys, img, "mean")
    assert np.isinf(want).any()
    for variant in (31, 33):
        out = _run(feats, K, Rt, xs, ys, img, "mean", True, variant=variant).cpu().numpy()
        assert _same(out, want)


@pytest.mark.parametrize("variant", [0, 31, 32, 33, 37])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_run_kernel_non_finite_features_on_the_map_border_match_the_reference_exactly(variant, dtype):
    """Blocks that hang over the map's border: the reference pads with zeros (geometry.py:161), so an out-of-map tap
    contributes w * 0 = +0 even when its in-map neighbours hold +-Inf / NaN.  The run kernel copies a stand-in texel
    for such taps and clears their registers after the unpack; the result must equal the oracle everywhere (round 1
    returned NaN where the reference returns +-Inf)."""
    feats, K, Rt, xs, ys, img = _rig_case(2, 3, 256, (20, 33), (19, 45), seed=31)
    feats[0, 1, 5, 7, 11] = np.inf
    feats[0, 2, 9, 3, 20] = -np.inf
    feats[0, 0, 64, 10, 10] = np.nan
    feats[:, :, 3, 0, :] = np.inf       # the whole top border row of channel 3
    feats[:, :, 4, :, 0] = -np.inf      # the left border column of channel 4
    feats[:, :, 5, -1, -1] = np.inf
    feats[1, :, 6, -1, :] = -np.inf     # bottom row
    feats[1, :, 7, :, -1] = np.nan      # right column
    f = torch.from_numpy(feats).to(dtype).float().numpy()
    for mode in ("mean", "sum", "max", "none"):
        want = orc.warp_fuse(f, K, Rt, xs, ys, img, mode)
        assert np.isinf(want).any()
        out = _run(f, K, Rt, xs, ys, img, mode, True, dtype=dtype, variant=variant).cpu().numpy()
        assert _same(out, want), mode


@pytest.mark.parametrize("ksplit", ["1", "2"])
@pytest.mark.parametrize("C,dtype", [(512, torch.float32), (384, torch.float32), (132, torch.float32), (512, torch.bfloat16), (264, torch.bfloat16)])
def test_run_kernel_one_or_two_warps_per_row_segment(monkeypatch, ksplit, C, dtype):
    """The default sum / mean kernel runs with one warp per row segment (large grids) or two that share the segment's tables
    and take every other 512-byte chunk (small grids, csrc/bevipm_run.cu pick_ksplit): both forms, forced through the
    development switch, on whole, odd and partial chunk counts, with a calibration change inside the frame group."""
    monkeypatch.setenv("BEVIPM_RUN_KSPLIT", ksplit)
    feats, K, Rt, xs, ys, img = _rig_case(3, 5, C, (31, 53), (37, 91), seed=17)
    K, Rt = K.copy(), Rt.copy()
    K[2, :, 0, 0] *= 1.03                      # frame 2 has its own calibration
    f = torch.from_numpy(feats).to(dtype).float().numpy()
    for mode in ("mean", "sum"):
        want = orc.warp_fuse(f, K, Rt, xs, ys, img, mode)
        out = _run(f, K, Rt, xs, ys, img, mode, True, dtype=dtype).cpu().numpy()
        assert _same(out, want), (mode, ksplit)


@pytest.mark.parametrize("seed", range(12))
def test_default_dispatch_random_shapes(seed):
    """Whatever kernel the library picks (variant 0) for a random shape and mode must equal the oracle: whole and
    partial channel chunks, 1..17 views, BEV grids smaller and larger than the source maps."""
    rng = np.random.RandomState(1000 + seed)
    B, V = int(rng.randint(1, 4)), int(rng.choice([1, 2, 3, 5, 7, 8, 9, 16, 17]))
    bf16 = bool(rng.randint(0, 2))
    C = int(rng.choice([256, 512, 264] if bf16 else [128, 256, 384, 132]))
    fhw = (int(rng.randint(8, 40)), int(rng.randint(8, 60)))
    bhw = (int(rng.randint(5, 50)), int(rng.randint(5, 100)))
    mode = str(rng.choice(["mean", "sum", "max"]))
    feats, K, Rt, xs, ys, img = _rig_case(B, V, C, fhw, bhw, seed=seed)
    if bf16:
        feats = torch.from_numpy(feats).bfloat16().float().numpy()
    want = orc.warp_fuse(feats, K, Rt, xs, ys, img, mode)
    out = _run(feats, K, Rt, xs, ys, img, mode, True, dtype=torch.bfloat16 if bf16 else torch.float32).cpu().numpy()
    assert _same(out, want), (B, V, C, fhw, bhw, mode, bf16)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("variant", [0, 21, 1])
def test_max_fusion_all_three_kernels(dtype, variant):
    """fusion.py:22 on a run-kernel-eligible shape: variant 0 = run kernel, 21 = list kernel, 1 = tile kernel.
    The rig has views that miss cells (their zero padding takes part in the maximum) and cells no view sees."""
    feats, K, Rt, xs, ys, img = _rig_case(2, 7, 256, (31, 53), (37, 91), seed=23)
    feats = feats - 1.5   # mostly negative values: the zeros of missing views must win there
    f = torch.from_numpy(feats).to(dtype).float().numpy()
    want = orc.warp_fuse(f, K, Rt, xs, ys, img, "max")
    assert (want == 0).any() and (want < 0).any() and (want > 0).any()
    out = _run(f, K, Rt, xs, ys, img, "max", True, dtype=dtype, variant=variant).cpu().numpy()
    assert _same(out, want)


@pytest.mark.parametrize("dtype,out_bf16", [(torch.float32, False), (torch.bfloat16, False), (torch.bfloat16, True)])
@pytest.mark.parametrize("variant", [0, 1])
def test_per_view_output_run_and_tile_kernels(dtype, out_bf16, variant):
    """geometry.py:162-163 on a run-kernel-eligible shape: variant 0 = run kernel (store as you go), 1 = tile kernel.
    Views that miss cells or whole segments must come out as exact zeros."""
    feats, K, Rt, xs, ys, img = _rig_case(2, 7, 256, (31, 53), (37, 91), seed=37)
    f = torch.from_numpy(feats).to(dtype).float().numpy()
    want = orc.warp_fuse(f, K, Rt, xs, ys, img, "none")
    assert (want == 0).mean() > 0.05
    out = _run(f, K, Rt, xs, ys, img, "none", True, dtype=dtype, out_bf16=out_bf16, variant=variant).cpu()
    if out_bf16:
        assert torch.equal(out, torch.from_numpy(want).bfloat16())
    else:
        assert _same(out.numpy(), want)


@pytest.mark.parametrize("variant", [0, 31, 33, 35, 37])
@pytest.mark.parametrize("C,dtype", [(136, torch.float32), (64, torch.float32), (4, torch.float32), (264, torch.bfloat16), (8, torch.bfloat16)])
@pytest.mark.parametrize("mode", ["mean", "max", "none"])
def test_run_kernel_partial_channel_chunks(variant, C, dtype, mode):
    """C is not a multiple of one 512-byte chunk (the reference's sanity config has C = 64): the spare lanes of the
    last chunk re-read valid channels and store nothing."""
    if mode != "mean" and variant != 0:
        pytest.skip("forced run-kernel variants are the sum/mean instantiations")
    feats, K, Rt, xs, ys, img = _rig_case(2, 5, C, (31, 53), (37, 91), seed=41)
    f = torch.from_numpy(feats).to(dtype).float().numpy()
    want = orc.warp_fuse(f, K, Rt, xs, ys, img, mode)
    out = _run(f, K, Rt, xs, ys, img, mode, True, dtype=dtype, variant=variant).cpu().numpy()
    assert _same(out, want)
    from bevipm import _lib
    assert 30 <= int(_lib.load().bevipm_last_variant()) <= 39   # really the run kernel


@pytest.mark.parametrize("views", [1, 2, 9, 16, 17, 32])
def test_run_kernel_view_counts(views):
    feats, K, Rt, xs, ys, img = _rig_case(1, views, 128, (20, 33), (19, 45), seed=10 + views)
    want = orc.warp_fuse(feats, K, Rt, xs, ys, img, "mean")
    for variant in (30, 33, 35):
        out = _run(feats, K, Rt, xs, ys, img, "mean", True, variant=variant).cpu().numpy()
        assert _same(out, want), variant


@pytest.mark.parametrize("mode", ["max", "none"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_max_and_per_view_fast_path(mode, dtype):
    feats, K, Rt, xs, ys, img = _rig_case(1, 7, 72, (34, 60), (30, 90), seed=6)
    f = torch.from_numpy(feats).to(dtype).float().numpy()
    want = orc.warp_fuse(f, K, Rt, xs, ys, img, mode)
    out = _run(f, K, Rt, xs, ys, img, mode, True, dtype=dtype).cpu().numpy()
    assert _same(out, want)


def test_strided_kernel_nchw_bf16():
    feats, K, Rt, xs, ys, img = _rig_case(1, 3, 7, (20, 33), (19, 45), seed=8)
    f = torch.from_numpy(feats).bfloat16().float().numpy()
    want = orc.warp_fuse(f, K, Rt, xs, ys, img, "mean")
    out = _run(f, K, Rt, xs, ys, img, "mean", False, dtype=torch.bfloat16).cpu().numpy()
    assert _same(out, want)


# ---- BASELINE.json shapes at full size ----------------------------------------------------------------

@pytest.mark.parametrize("name", ["c1", "c3"])
def test_full_size_fp32_vs_oracle(name):
    from bevipm import rig
    wl = rig.WORKLOADS[name]
    feats, K, Rt, xs, ys, img = _rig_case(1, wl.views, wl.channels, wl.feat_hw, wl.bev_hw, seed=0)
    nhwc = np.ascontiguousarray(feats.transpose(0, 1, 3, 4, 2)).transpose(0, 1, 4, 2, 3)
    want = orc.warp_fuse(nhwc, K, Rt, xs, ys, img, "mean", channels_last_out=True)
    out = _run(feats, K, Rt, xs, ys, img, "mean", True).cpu().numpy()
    assert _same(out, want)
    assert np.count_nonzero(out) > 0.5 * out.size


@pytest.mark.parametrize("variant", [0, 31, 33])
def test_full_size_c2_bf16_frames_and_properties(variant):
    """BASELINE config 1 (the bench workload): 8 frames x 7 views x 1024 ch bf16.  One frame is checked
    against the oracle; the whole batch through size-independent properties.  variant 0 = the default
    (list kernel), 31 / 33 = the run kernel (4 warps per row segment / one warp walking all chunks)."""
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
    run = lambda t, mode="mean", obf=False: ops.warp_fuse(t, Kd[:t.shape[0]], Rd[:t.shape[0]], xd, yd, img[0], img[1],
                                                          _lib.MODES[mode], obf, variant if mode != "none" else 0)
    out = run(f)
    assert out.shape == (B, C, *wl.bev_hw) and out.dtype == torch.float32
    # (1) frame 3 against the oracle, bit-exact
    f3 = f[3:4].float().cpu().numpy()
    want = orc.warp_fuse(f3, K[None].numpy(), Rt[None].numpy(), xs.numpy(), ys.numpy(), img, "mean")
    assert _same(out[3:4].cpu().numpy(), want)
    # (2) bf16 output is the rounding of the fp32 output
    assert torch.equal(run(f, obf=True), out.bfloat16())
    # (3) mean == sum / V exactly (IEEE division of the same accumulator)
    s = run(f, "sum")
    assert torch.equal(out, s / torch.tensor(float(V), device=DEV))   # tensor divisor: IEEE division
    # (4) homogeneity: scaling the features by 2 scales the BEV by 2 exactly (power of two)
    assert torch.equal(run(f * 2), out * 2)
    # (5) frames are independent: permuting frames permutes the output
    perm = torch.tensor([5, 0, 7, 2, 1, 6, 3, 4], device=DEV)
    assert torch.equal(run(f[perm]), out[perm])
    # (6) all-ones features: every cell = (in-bounds tap weight summed over views) / V, identical across channels
    ones = torch.ones_like(f[:1])
    cov = run(ones)
    assert torch.equal(cov[:, :1].expand_as(cov), cov)
    assert float(cov.max()) <= 1.0 + 1e-6 and float(cov.min()) >= 0.0
    # (7) fused mean == our per-view maps reduced sequentially
    pv = run(f[:1], "none")
    acc = pv[:, 0].clone()
    for v in range(1, V):
        acc += pv[:, v]
    assert torch.equal(out[:1], acc / torch.tensor(float(V), device=DEV))


# ---- properties with known answers ----------------------------------------------------------------------

def test_ramp_features_return_coordinates():
    """feature = x (resp. y) ramp  =>  output = ix (resp. iy) wherever all four taps are in bounds."""
    from bevipm import _lib, ops, rig
    V, Hf, Wf, Hb, Wb = 3, 40, 64, 36, 100
    K, Rt = rig.look_at_rig(V, 1)
    xs, ys = rig.ground_axes(Hb, Wb, rig.WILDTRACK_BOUNDS)
    ramp = torch.zeros(1, V, 4, Hf, Wf)
    ramp[:, :, 0] = torch.arange(Wf, dtype=torch.float32)[None, :]
    ramp[:, :, 1] = torch.arange(Hf, dtype=torch.float32)[:, None]
    ramp[:, :, 2] = 1.0
    f = ramp.to(DEV).permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)
    Kd, Rd = K[None].contiguous().to(DEV), Rt[None, :, :3].contiguous().to(DEV)
    pv = ops.warp_fuse(f, Kd, Rd, xs.to(DEV), ys.to(DEV), 1080, 1920, _lib.NONE, False, 0)
    ix, iy = ops.sample_coords(Kd, Rd, xs.to(DEV), ys.to(DEV), (Hf, Wf), (1080, 1920))
    inside = (ix >= 0) & (ix <= Wf - 1) & (iy >= 0) & (iy <= Hf - 1)
    assert inside.any()
    assert torch.allclose(pv[:, :, 0][inside], ix[inside], atol=2e-4)
    assert torch.allclose(pv[:, :, 1][inside], iy[inside], atol=2e-4)
    assert torch.allclose(pv[:, :, 2][inside], torch.ones_like(ix[inside]), atol=1e-6)
    outside = (ix < -1) | (ix > Wf) | (iy < -1) | (iy > Hf)
    assert outside.any() and float(pv[:, :, 2][outside].abs().max()) == 0.0


def test_sample_coords_match_oracle():
    from bevipm import ops, rig
    K, Rt = rig.look_at_rig(7, 0)
    xs, ys = rig.ground_axes(120, 360, rig.WILDTRACK_BOUNDS)
    ix, iy = ops.sample_coords(K[None].contiguous().to(DEV), Rt[None, :, :3].contiguous().to(DEV), xs.to(DEV), ys.to(DEV),
                               (135, 240), (1080, 1920))
    ox, oy = orc.coords(K[None].numpy(), Rt[None].numpy(), xs.numpy(), ys.numpy(), (135, 240), (1080, 1920))
    assert np.array_equal(ix.cpu().numpy(), ox) and np.array_equal(iy.cpu().numpy(), oy)


def test_view_permutation_sum():
    feats, K, Rt, xs, ys, img = _rig_case(1, 7, 16, (34, 60), (30, 90), seed=9)
    a = _run(feats, K, Rt, xs, ys, img, "sum", True).cpu().numpy()
    p = np.array([3, 0, 6, 1, 5, 2, 4])
    b = _run(feats[:, p], K[:, p], Rt[:, p], xs, ys, img, "sum", True).cpu().numpy()
    assert np.abs(a - b).max() <= 1e-5 * np.abs(a).max()       # fp32 re-association only
    m1 = _run(feats, K, Rt, xs, ys, img, "max", True).cpu().numpy()
    m2 = _run(feats[:, p], K[:, p], Rt[:, p], xs, ys, img, "max", True).cpu().numpy()
    assert np.array_equal(m1, m2)                                # max is order-free


# ---- module surface (the drop-in boundary) ---------------------------------------------------------------

def test_modules_match_reference_golden_and_calibration_forms(golden):
    import bevipm
    z = golden("rig_small")
    feats = torch.from_numpy(z["feats"]).to(DEV)
    K, Rt = torch.from_numpy(z["K"]).to(DEV), torch.from_numpy(z["Rt"]).to(DEV)
    B, V = feats.shape[:2]
    bounds = tuple(float(x) for x in z["bounds"])
    bev_h, bev_w = (int(x) for x in z["bev_hw"])
    img = tuple(int(x) for x in z["img_size"])
    geom = bevipm.GeometryTransformer(bev_h, bev_w, bounds, warp_impl="kornia").to(DEV)
    assert list(geom.state_dict().keys()) == [] and list(geom.parameters()) == []
    pv = geom(feats, K, Rt, img_size=img)
    assert pv.dtype == torch.float32 and _same(pv.cpu().numpy(), z["out_none"])
    # nested lists (what the loader hands over, wildtrack_loader.py:389-401)
    pv2 = geom(feats, [[K[b, v] for v in range(V)] for b in range(B)], [[Rt[b, v] for v in range(V)] for b in range(B)], img_size=img)
    assert torch.equal(pv, pv2)
    assert _same(bevipm.ConcatFusion()(pv).cpu().numpy(), z["out_concat"])
    for mode in ("sum", "mean", "max"):
        want = orc.warp_fuse(z["feats"], z["K"], z["Rt"], z["xs"], z["ys"], img, mode)
        assert _same(bevipm.SimpleFusion(mode)(pv).cpu().numpy(), want)
        fused = bevipm.FusedIPM(bev_h, bev_w, bounds, fusion=mode).to(DEV)
        assert _same(fused(feats, K, Rt, img_size=img).cpu().numpy(), want)
    cat = bevipm.FusedIPM(bev_h, bev_w, bounds, fusion="concat").to(DEV)(feats, K, Rt, img_size=img)
    assert _same(cat.cpu().numpy(), z["out_concat"])
    # shared calibration forms: [V,3,3]/[V,4,4] and a single 2-D pair (geometry.py:100-103)
    a = geom(feats, K[0], Rt[0], img_size=img)
    b = geom(feats, K[0][None].expand(B, -1, -1, -1), Rt[0][None].expand(B, -1, -1, -1), img_size=img)
    assert torch.equal(a, b)
    c = geom(feats, K[0, 0], Rt[0, 0], img_size=img)
    d = geom(feats, K[0, 0].expand(B, V, -1, -1), Rt[0, 0].expand(B, V, -1, -1), img_size=img)
    assert torch.equal(c, d)
    # fp16 features behave like grid_sampler under autocast: fp32 compute, fp32 result
    h = geom(feats.half(), K, Rt, img_size=img)
    assert h.dtype == torch.float32


def test_cpu_tensors_are_rejected():
    import bevipm
    geom = bevipm.FusedIPM(8, 8, (-1.0, 1.0, -1.0, 1.0))
    with pytest.raises(RuntimeError):
        geom(torch.zeros(1, 1, 4, 4, 4), torch.eye(3), torch.eye(4))


def test_autograd_through_fused_module_nchw_input():
    import bevipm
    feats, K, Rt, xs, ys, img = _rig_case(1, 3, 8, (20, 33), (19, 45), seed=11)
    f = torch.from_numpy(feats).to(DEV).requires_grad_(True)
    mod = bevipm.FusedIPM(19, 45, bevipm.rig.WILDTRACK_BOUNDS, fusion="mean").to(DEV)
    out = mod(f, torch.from_numpy(K).to(DEV), torch.from_numpy(Rt).to(DEV), img_size=img)
    cot = torch.randn(out.shape, device=DEV, generator=torch.Generator(device=DEV).manual_seed(0))
    (out * cot).sum().backward()
    want = orc.warp_fuse_bwd(cot.cpu().numpy(), K, Rt, xs, ys, feats.shape, img, "mean")
    assert f.grad is not None and f.grad.shape == f.shape
    assert np.abs(f.grad.cpu().numpy() - want).max() <= 1e-5 * np.abs(want).max()


def test_layout_prepass_and_view_reduction_kernels():
    from bevipm import ops
    g = torch.Generator(device=DEV).manual_seed(3)
    for dt in (torch.float32, torch.bfloat16):
        x = torch.randn(2, 3, 37, 19, 23, device=DEV, generator=g).to(dt)
        y = ops.to_channels_last5(x)
        assert y.stride(2) == 1 and torch.equal(x, y)
    pv = torch.randn(2, 5, 6, 11, 13, device=DEV, generator=g)
    seq = pv[:, 0].clone()
    for v in range(1, 5):
        seq = seq + pv[:, v]
    assert torch.equal(ops.fuse_views(pv, "sum"), seq)
    assert torch.equal(ops.fuse_views(pv, "mean"), seq / torch.tensor(5.0, device=DEV))
    assert torch.equal(ops.fuse_views(pv, "max"), pv.max(dim=1).values)


@pytest.mark.parametrize("pinned,gather", [(True, "1"), (True, "0"), (False, "1")])
def test_host_buffer_entry_matches_oracle(monkeypatch, pinned, gather):
    """Pinned features are pulled over PCIe by the span gather kernel; BEVIPM_HOST_GATHER=0 and pageable memory take the
    banded 2-D copies.  Repeated calls with the same calibration re-use the cached span table; a new calibration rebuilds it."""
    from bevipm import ops
    monkeypatch.setenv("BEVIPM_HOST_GATHER", gather)
    feats, K, Rt, xs, ys, img = _rig_case(3, 4, 32, (27, 48), (24, 72), seed=13)
    host = torch.from_numpy(np.ascontiguousarray(feats.transpose(0, 1, 3, 4, 2)))
    if pinned:
        host = host.pin_memory()
    want = orc.warp_fuse(feats, K, Rt, xs, ys, img, "mean")
    for _ in range(2):
        out = ops.warp_fuse_host(host, torch.from_numpy(K), torch.from_numpy(Rt[:, :, :3, :]).contiguous(),
                                 torch.from_numpy(xs), torch.from_numpy(ys), img, "mean")
        assert _same(out.numpy().transpose(0, 3, 1, 2), want)
    feats2, K2, Rt2, _, _, _ = _rig_case(3, 4, 32, (27, 48), (24, 72), seed=14)
    out = ops.warp_fuse_host(host, torch.from_numpy(K2), torch.from_numpy(Rt2[:, :, :3, :]).contiguous(),
                             torch.from_numpy(xs), torch.from_numpy(ys), img, "mean")
    assert _same(out.numpy().transpose(0, 3, 1, 2), orc.warp_fuse(feats, K2, Rt2, xs, ys, img, "mean"))


@pytest.mark.parametrize("dense", ["0", "0.8", "0.4"])
def test_host_buffer_entry_with_per_frame_calibration_and_modes(monkeypatch, dense):
    """Every frame has its own cameras (its own spans, bitmaps and copy-engine rows), an odd number of frames goes through the
    two staging buffers, and the per-view output mode takes the same upload path: gather kernel alone and with the dense
    rows of each view uploaded by the copy engine beside it."""
    from bevipm import ops
    monkeypatch.setenv("BEVIPM_HOST_DMA_DENSE", dense)
    B, V, C, fhw, bhw = 5, 4, 64, (27, 48), (24, 72)
    feats, K, Rt, xs, ys, img = _rig_case(B, V, C, fhw, bhw, seed=23)
    K, Rt = K.copy(), Rt.copy()
    for b in range(1, B):
        Kb, Rb = _rig_case(1, V, C, fhw, bhw, seed=23 + b)[1:3]
        K[b], Rt[b] = Kb[0], Rb[0]
    host = torch.from_numpy(np.ascontiguousarray(feats.transpose(0, 1, 3, 4, 2))).pin_memory()
    for mode in ("mean", "none"):
        want = orc.warp_fuse(feats, K, Rt, xs, ys, img, mode)
        for _ in range(2):
            out = ops.warp_fuse_host(host, torch.from_numpy(K), torch.from_numpy(Rt[:, :, :3, :]).contiguous(),
                                     torch.from_numpy(xs), torch.from_numpy(ys), img, mode).numpy()
            got = out.transpose(0, 3, 1, 2) if mode == "mean" else out.transpose(0, 1, 4, 2, 3)
            assert _same(got, want), (mode, dense)


def test_host_buffer_entry_never_reads_unsampled_staging_memory(monkeypatch):
    """The device staging arena is only partially written (sampled spans): with the arena poisoned with NaN patterns
    (BEVIPM_HOST_POISON) the result must still be the oracle's -- no kernel reads a texel that was not uploaded."""
    from bevipm import _lib, ops
    _lib.load().bevipm_host_release()            # force a fresh (poisoned) arena
    monkeypatch.setenv("BEVIPM_HOST_POISON", "1")
    for dtype in (torch.float32, torch.bfloat16):
        feats, K, Rt, xs, ys, img = _rig_case(2, 7, 256, (40, 64), (36, 100), seed=19)
        f = torch.from_numpy(feats).to(dtype)
        host = f.permute(0, 1, 3, 4, 2).contiguous().pin_memory()
        for mode in ("mean", "max"):
            out = ops.warp_fuse_host(host, torch.from_numpy(K), torch.from_numpy(Rt[:, :, :3, :]).contiguous(),
                                     torch.from_numpy(xs), torch.from_numpy(ys), img, mode, out_dtype=torch.float32)
            want = orc.warp_fuse(f.float().numpy(), K, Rt, xs, ys, img, mode)
            assert _same(out.numpy().transpose(0, 3, 1, 2), want), (dtype, mode)
        _lib.load().bevipm_host_release()
    monkeypatch.delenv("BEVIPM_HOST_POISON")


def test_host_buffer_entry_uploads_only_sampled_texels(monkeypatch):
    """The host entry uploads only texels some BEV cell samples: from pinned memory exactly those when the gather kernel works
    alone (a per-row bitmap drives it), those plus the holes of the dense rows the copy engine takes in the default mode, and
    from pageable memory the row spans that hold them.  Every texel NO cell samples is poisoned with NaN in the HOST buffer:
    the result must still be the oracle's on the clean features, and the bytes copied must be exactly the sampled texels
    (gather alone) / lie between them and the sampled rows (default, pageable)."""
    from bevipm import _lib, ops
    B, V, C, fhw, bhw = 2, 5, 32, (40, 64), (24, 72)
    feats, K, Rt, xs, ys, img = _rig_case(B, V, C, fhw, bhw, seed=17)
    want = orc.warp_fuse(feats, K, Rt, xs, ys, img, "mean")
    ix, iy = orc.coords(K[:1], Rt[:1], xs, ys, fhw, img)
    ix, iy = ix.reshape(V, *bhw), iy.reshape(V, *bhw)
    nhwc = np.ascontiguousarray(feats.transpose(0, 1, 3, 4, 2))
    touched, row_texels = 0, 0
    for v in range(V):
        x0, y0 = np.floor(ix[v]).astype(np.int64), np.floor(iy[v]).astype(np.int64)
        hit = np.zeros(fhw, dtype=bool)
        for dy in (0, 1):
            for dx in (0, 1):
                ok = (x0 + dx >= 0) & (x0 + dx < fhw[1]) & (y0 + dy >= 0) & (y0 + dy < fhw[0])
                hit[(y0 + dy)[ok], (x0 + dx)[ok]] = True
        nhwc[:, v][:, ~hit] = np.nan
        touched += int(hit.sum())
        row_texels += int(hit.any(axis=1).sum()) * fhw[1]
    assert touched < row_texels < V * fhw[0] * fhw[1]
    host = torch.from_numpy(nhwc).pin_memory()
    calib_bytes = (B * V * 21 + bhw[0] + bhw[1]) * 4
    monkeypatch.setenv("BEVIPM_HOST_DMA_DENSE", "0")          # the gather kernel alone
    out = ops.warp_fuse_host(host, torch.from_numpy(K), torch.from_numpy(Rt[:, :, :3, :]).contiguous(),
                             torch.from_numpy(xs), torch.from_numpy(ys), img, "mean")
    assert _same(out.numpy().transpose(0, 3, 1, 2), want)
    copied = int(_lib.load().bevipm_host_last_h2d_bytes()) - calib_bytes
    assert copied == B * touched * C * 4
    for dense in ("0.8", "0.3"):                              # default: dense rows through the copy engine; and a low threshold
        monkeypatch.setenv("BEVIPM_HOST_DMA_DENSE", dense)
        out = ops.warp_fuse_host(host, torch.from_numpy(K), torch.from_numpy(Rt[:, :, :3, :]).contiguous(),
                                 torch.from_numpy(xs), torch.from_numpy(ys), img, "mean")
        assert _same(out.numpy().transpose(0, 3, 1, 2), want)
        copied = int(_lib.load().bevipm_host_last_h2d_bytes()) - calib_bytes
        assert B * touched * C * 4 <= copied <= B * row_texels * C * 4
    monkeypatch.delenv("BEVIPM_HOST_DMA_DENSE")
    pageable = torch.from_numpy(nhwc.copy())
    out = ops.warp_fuse_host(pageable, torch.from_numpy(K), torch.from_numpy(Rt[:, :, :3, :]).contiguous(),
                             torch.from_numpy(xs), torch.from_numpy(ys), img, "mean")
    assert _same(out.numpy().transpose(0, 3, 1, 2), want)
    copied = int(_lib.load().bevipm_host_last_h2d_bytes()) - calib_bytes
    assert B * touched * C * 4 <= copied <= B * row_texels * C * 4


def test_library_was_the_thing_that_ran():
    from bevipm import _lib
    before = _lib.launch_count()
    feats, K, Rt, xs, ys, img = _rig_case(1, 2, 8, (12, 16), (10, 20), seed=1)
    _run(feats, K, Rt, xs, ys, img, "mean", True)
    assert _lib.launch_count() == before + 1


@pytest.mark.parametrize("variant", [0, 7])
def test_mean_division_is_ieee(variant):
    """mean = sum / V must be the IEEE quotient for every accumulator value: the 3-op exact division in
    the kernels is checked against true division from the denormal range up to near FLT_MAX, and with
    +-Inf / NaN accumulators (which must take the guarded library-division path)."""
    feats, K, Rt, xs, ys, img = _rig_case(1, 7, 16, (34, 60), (30, 90), seed=21)
    V = torch.tensor(7.0, device=DEV)
    for scale in (2.0 ** -140, 2.0 ** -126, 2.0 ** -110, 2.0 ** -60, 1.0, 2.0 ** 60, 2.0 ** 120):
        f = (feats.astype(np.float64) * scale).astype(np.float32)
        s = _run(f, K, Rt, xs, ys, img, "sum", True, variant=variant)
        m = _run(f, K, Rt, xs, ys, img, "mean", True, variant=variant)
        assert torch.equal(m, s / V), scale
        assert float(s.abs().max()) > 0
    f = feats.copy()
    f[0, :, 3, 10:20, 10:30] = np.inf
    f[0, :, 5, 5:9, 40:50] = -np.inf
    f[0, :, 7, 20:25, 20:25] = np.nan
    s = _run(f, K, Rt, xs, ys, img, "sum", True, variant=variant)
    m = _run(f, K, Rt, xs, ys, img, "mean", True, variant=variant)
    want = s / V
    assert torch.equal(torch.isnan(m), torch.isnan(want))
    assert torch.equal(m[~torch.isnan(m)], want[~torch.isnan(want)])
    assert bool(torch.isinf(m).any()) and bool(torch.isnan(m).any())


# ---- kornia-compatible geometry (compat mode, parity unpinned: kornia is not installed anywhere we can run) --------

@pytest.mark.parametrize("fusion", ["none", "mean"])
def test_emulate_kornia_matches_from_spec_restatement(fusion):
    """bevipm's emulate_kornia=True against oracle/kornia_chain.py (kornia's published warp_perspective algorithm over
    the M of geometry.py:126-133).  Two different fp32 routes to the same sample positions (ours: H, then p*size/(size-1)
    - 0.5; kornia's: normalised homographies and a matrix inverse), so the check is a tolerance on smooth features."""
    import bevipm
    from bevipm import rig
    from oracle import kornia_chain
    B, V, C, fhw, bhw = 1, 4, 8, (34, 60), (30, 90)
    K, Rt = rig.look_at_rig(V, 5)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, fhw[0]), torch.linspace(0, 1, fhw[1]), indexing="ij")
    base = torch.stack([torch.sin(3 * xx + c) * torch.cos(2 * yy - 0.3 * c) for c in range(C)])      # smooth maps
    feats = torch.stack([base * (1 + 0.1 * v) for v in range(V)])[None].contiguous()
    want = kornia_chain.warp_views(feats, K[None], Rt[None], bhw, rig.WILDTRACK_BOUNDS, rig.WILDTRACK_IMG_SIZE)
    if fusion == "mean":
        want = want.mean(dim=1)
    mod = bevipm.FusedIPM(bhw[0], bhw[1], rig.WILDTRACK_BOUNDS, fusion=fusion if fusion != "none" else "concat",
                          warp_impl="kornia", emulate_kornia=True).to(DEV)
    out = mod(feats.to(DEV), K[None].to(DEV), Rt[None].to(DEV), img_size=rig.WILDTRACK_IMG_SIZE).cpu()
    out = out.reshape(want.shape)
    # interior agreement; a cell whose sample sits within ~1e-3 px of the map border may flip a tap in or out
    diff = (out - want).abs()
    assert float(diff.median()) <= 1e-5
    assert float((diff > 2e-3 * float(want.abs().max())).float().mean()) <= 2e-3
    # and it really is a different geometry from the grid_sample branch (half a BEV cell + the (size-1) scaling)
    ref = bevipm.FusedIPM(bhw[0], bhw[1], rig.WILDTRACK_BOUNDS, fusion=fusion if fusion != "none" else "concat",
                          warp_impl="kornia").to(DEV)(feats.to(DEV), K[None].to(DEV), Rt[None].to(DEV),
                                                       img_size=rig.WILDTRACK_IMG_SIZE).cpu().reshape(want.shape)
    assert float((ref - want).abs().median()) > 10 * float(diff.median()) + 1e-6


# ---- no stray reads: the features live inside a NaN-filled allocation -------------------------------------------------

@pytest.mark.parametrize("C,dtype", [(256, torch.float32), (136, torch.float32), (512, torch.bfloat16), (264, torch.bfloat16)])
@pytest.mark.parametrize("mode", ["mean", "max", "none"])
def test_features_inside_a_nan_halo(C, dtype, mode):
    """The logical feature tensor is a window of a larger allocation filled with NaN (one texel of halo on every side
    of every map, one 16-byte vector of halo on both sides of the channels).  Every tap the kernels blend -- also the
    weight-0 stand-ins of out-of-map taps and the spare lanes of a partial chunk -- multiplies what it read, so a single
    read outside the window would put a NaN into the output."""
    from bevipm import _lib, ops
    B, V, fhw, bhw = 2, 5, (31, 53), (37, 91)
    feats, K, Rt, xs, ys, img = _rig_case(B, V, C, fhw, bhw, seed=43)
    ve = 4 if dtype == torch.float32 else 8
    f = torch.from_numpy(feats).to(dtype)
    big = torch.full((B, V, fhw[0] + 2, fhw[1] + 2, C + 2 * ve), float("nan"), dtype=dtype, device=DEV)
    win = big[:, :, 1:-1, 1:-1, ve:ve + C]
    win.copy_(f.permute(0, 1, 3, 4, 2).to(DEV))
    fv = win.permute(0, 1, 4, 2, 3)   # logical [B,V,C,Hf,Wf], channel stride 1, the other strides those of the big tensor
    _, Kd, Rd, xd, yd = _dev_inputs(feats[:, :, :1], K, Rt, xs, ys, True)
    out = ops.warp_fuse(fv, Kd, Rd, xd, yd, int(img[0]), int(img[1]), _lib.MODES[mode], False, 0)
    torch.cuda.synchronize()
    assert 30 <= int(_lib.load().bevipm_last_variant()) <= 39   # the run kernel took it (fast path on a strided window)
    want = orc.warp_fuse(f.float().numpy(), K, Rt, xs, ys, img, mode)
    assert _same(out.cpu().numpy(), want)
