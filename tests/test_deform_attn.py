"""Deformable-attention sampling kernel vs the from-spec oracle.  The reference only has a placeholder (fusion.py:25-36), so
this row cannot be pinned to it; the oracle is pinned to golden vectors of an independent published implementation instead
(`transformers`' multi_scale_deformable_attention, tests/golden/make_deform_golden.py).  Tolerances: fp32 <= 1e-5,
bf16 <= 1e-2, max-normalised."""
import numpy as np
import pytest
import torch

from oracle import deform_attn_oracle as dorc


def _case(B, Q, M, D, shapes, P, seed, spread=0.6):
    g = torch.Generator().manual_seed(seed)
    S = sum(h * w for h, w in shapes)
    value = torch.randn(B, S, M, D, generator=g)
    L = len(shapes)
    loc = 0.5 + spread * (torch.rand(B, Q, M, L, P, 2, generator=g) - 0.5) * 2     # some land outside [0,1]
    aw = torch.softmax(torch.randn(B, Q, M, L * P, generator=g), dim=-1).view(B, Q, M, L, P)
    return value, loc, aw


GOLD = np.load(__import__("pathlib").Path(__file__).resolve().parent / "golden" / "deform_attn_hf.npz")
HF_CASES = ["views7", "ragged", "wide_head"]


def _gold(name):
    g = {k.split(".", 1)[1]: torch.from_numpy(GOLD[k]) for k in GOLD.files if k.startswith(name + ".")}
    g["shapes_list"] = [tuple(int(x) for x in hw) for hw in g["shapes"]]
    return g


@pytest.mark.parametrize("name", HF_CASES)
def test_oracle_is_pinned_to_the_published_implementation(name):
    """The reference has no deformable attention (fusion.py:25-36 is a placeholder).  The from-spec oracle is pinned to an
    independent published implementation instead: golden vectors produced by `transformers`' unmodified
    `multi_scale_deformable_attention` (the PyTorch port of Deformable-DETR's ms_deform_attn_core_pytorch;
    tests/golden/make_deform_golden.py).  Both oracle forms must reproduce them: the grid_sample form to fp32 rounding, the
    float64 loops to 1e-5."""
    g = _gold(name)
    a = dorc.deform_attn_grid_sample(g["value"], g["shapes_list"], g["loc"], g["aw"])
    assert float((a - g["out"]).abs().max()) <= 1e-6 * float(g["out"].abs().max())
    if name != "views7":   # (the loops are slow)
        b = dorc.deform_attn_scalar(g["value"], g["shapes_list"], g["loc"], g["aw"])
        assert float((b - g["out"]).abs().max()) <= 1e-5 * float(g["out"].abs().max())


@pytest.mark.gpu
@pytest.mark.parametrize("name", HF_CASES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_kernel_against_the_published_implementation_golden(name, dtype):
    """Forward and backward of our kernels against the `transformers` golden vectors: fp32 <= 1e-5 (forward) / 1e-4 (gradients,
    atomics re-associate), bf16 values <= 1e-2."""
    from bevipm import ops
    g = _gold(name)
    sh = g["shapes"].to(torch.int32).cuda()
    start = torch.tensor(np.concatenate([[0], np.cumsum([h * w for h, w in g["shapes_list"]])[:-1]]), dtype=torch.int64).cuda()
    v = g["value"].to(dtype).cuda().requires_grad_(True)
    l = g["loc"].cuda().requires_grad_(True)
    a = g["aw"].cuda().requires_grad_(True)
    out = ops.deform_attn(v, sh, start, l, a, out_dtype=torch.float32)
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert float((out.detach().cpu() - g["out"]).abs().max()) <= tol * float(g["out"].abs().max())
    (out * g["cot"].cuda()).sum().backward()
    gt = 1e-4 if dtype == torch.float32 else 2e-2
    for got, want in ((v.grad.float().cpu(), g["g_value"]), (l.grad.cpu(), g["g_loc"]), (a.grad.cpu(), g["g_aw"])):
        assert float((got - want).abs().max()) <= gt * float(want.abs().max())


def test_two_oracle_forms_agree():
    shapes = [(5, 7), (4, 6), (3, 3)]
    value, loc, aw = _case(1, 6, 2, 4, shapes, 2, seed=0)
    a = dorc.deform_attn_grid_sample(value, shapes, loc, aw)
    b = dorc.deform_attn_scalar(value, shapes, loc, aw)
    assert float((a - b).abs().max()) <= 1e-5 * float(b.abs().max())


def test_oracle_known_answers():
    # constant value map, all samples strictly inside -> output = sum of attention weights = 1
    shapes = [(6, 8)]
    value = torch.ones(1, 48, 1, 4)
    loc = torch.full((1, 3, 1, 1, 2, 2), 0.5)
    aw = torch.tensor([0.25, 0.75]).view(1, 1, 1, 1, 2).expand(1, 3, 1, 1, 2)
    out = dorc.deform_attn_grid_sample(value, shapes, loc, aw)
    assert torch.allclose(out, torch.ones_like(out), atol=1e-6)
    # everything far outside -> zero
    out = dorc.deform_attn_grid_sample(value, shapes, loc + 5.0, aw)
    assert float(out.abs().max()) == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("M,D", [(8, 32), (3, 8), (1, 64), (2, 128)])
def test_kernel_matches_oracle(dtype, tol, M, D):
    from bevipm import ops
    shapes = [(9, 13), (7, 10), (5, 5), (12, 4)]
    value, loc, aw = _case(2, 37, M, D, shapes, 4, seed=3)
    value = value.to(dtype)
    want = dorc.deform_attn_grid_sample(value.float(), shapes, loc, aw)
    sh = torch.tensor(shapes, dtype=torch.int32)
    start = torch.tensor(np.concatenate([[0], np.cumsum([h * w for h, w in shapes])[:-1]]), dtype=torch.int64)
    out = ops.deform_attn(value.cuda(), sh.cuda(), start.cuda(), loc.cuda(), aw.cuda(), out_dtype=torch.float32).cpu()
    assert out.shape == want.shape
    assert float((out - want).abs().max()) <= tol * float(want.abs().max())
    if dtype == torch.bfloat16:
        # bf16 value in fp32 arithmetic: as tight as the fp32 case against the oracle fed the same rounded values
        assert float((out - want).abs().max()) <= 1e-5 * float(want.abs().max())


@pytest.mark.gpu
def test_kernel_config4_shape_properties():
    """BASELINE config 3: 120x360 queries, 8 heads x 4 points x 7 views, 256 ch bf16 (views as 135x240 maps)."""
    from bevipm import ops
    B, Q, M, D, L, P = 1, 120 * 360, 8, 32, 7, 4
    shapes = [(135, 240)] * L
    S = sum(h * w for h, w in shapes)
    g = torch.Generator(device="cuda").manual_seed(0)
    value = torch.randn(B, S, M, D, device="cuda", generator=g).bfloat16()
    loc = torch.rand(B, Q, M, L, P, 2, device="cuda", generator=g) * 1.2 - 0.1
    aw = torch.softmax(torch.randn(B, Q, M, L * P, device="cuda", generator=g), dim=-1).view(B, Q, M, L, P)
    sh = torch.tensor(shapes, dtype=torch.int32, device="cuda")
    start = torch.arange(L, device="cuda", dtype=torch.int64) * (135 * 240)
    out = ops.deform_attn(value, sh, start, loc, aw)
    assert out.shape == (B, Q, M * D) and out.dtype == torch.bfloat16
    # a slice of queries against the oracle
    sl = slice(1000, 1256)
    want = dorc.deform_attn_grid_sample(value.float().cpu(), shapes, loc[:, sl].cpu(), aw[:, sl].cpu())
    assert float((out[:, sl].float().cpu() - want).abs().max()) <= 1e-2 * float(want.abs().max())
    # linear in the attention weights, linear in the values (power-of-two scale: exact in fp32 output)
    o32 = ops.deform_attn(value, sh, start, loc, aw, out_dtype=torch.float32)
    assert torch.equal(ops.deform_attn(value * 2, sh, start, loc, aw, out_dtype=torch.float32), o32 * 2)
    assert torch.equal(ops.deform_attn(value, sh, start, loc, aw * 0.5, out_dtype=torch.float32), o32 * 0.5)
    # constant value maps and in-map samples -> every channel equals the sum of the weights (= 1)
    ones = torch.ones_like(value)
    inside = loc.clamp(0.05, 0.95)
    c = ops.deform_attn(ones, sh, start, inside, aw, out_dtype=torch.float32)
    assert float((c - 1).abs().max()) <= 1e-5


def test_cpu_tensors_rejected():
    from bevipm import ops
    with pytest.raises(RuntimeError):
        ops.deform_attn(torch.zeros(1, 4, 1, 8), torch.tensor([[2, 2]]), torch.tensor([0]), torch.zeros(1, 1, 1, 1, 1, 2),
                        torch.ones(1, 1, 1, 1, 1))


@pytest.mark.gpu
def test_fusion_module_matches_oracle_composition():
    """DeformAttnFusion = torch projections + our sampling kernel; the same projections on the CPU with the
    oracle doing the sampling must agree."""
    import bevipm
    torch.manual_seed(0)
    B, V, C, H, W = 2, 3, 64, 10, 14
    mod = bevipm.DeformAttnFusion(C, V, heads=4, points=2)
    with torch.no_grad():
        mod.sampling_offsets.weight.normal_(0, 0.02)
        mod.attention_weights.weight.normal_(0, 0.2)
    maps = torch.randn(B, V, C, H, W)
    out = mod.cuda()(maps.cuda()).cpu()
    assert out.shape == (B, C, H, W)
    mod = mod.cpu()
    with torch.no_grad():
        M, P, D = 4, 2, C // 4
        m2 = maps.permute(0, 1, 3, 4, 2)
        query = m2.mean(dim=1).reshape(B, H * W, C)
        value = mod.value_proj(m2.reshape(B, V * H * W, C)).view(B, V * H * W, M, D)
        off = mod.sampling_offsets(query).view(B, H * W, M, V, P, 2)
        aw = torch.softmax(mod.attention_weights(query).view(B, H * W, M, V * P), -1).view(B, H * W, M, V, P)
        ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
        ref = torch.stack([(xs + 0.5) / W, (ys + 0.5) / H], -1).reshape(1, H * W, 1, 1, 1, 2).float()
        loc = ref + off / torch.tensor([W, H], dtype=torch.float32)
        smp = dorc.deform_attn_grid_sample(value, [(H, W)] * V, loc, aw)
        want = mod.output_proj(smp).view(B, H, W, C).permute(0, 3, 1, 2)
    assert float((out - want).abs().max()) <= 2e-4 * float(want.abs().max())   # cuBLAS vs CPU GEMM in the projections


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,D", [(8, 32), (3, 8), (2, 128)])
def test_backward_matches_autograd_of_the_oracle(dtype, M, D):
    """d/d value, d/d sampling_locations, d/d attention_weights of our kernel against torch autograd through the from-spec
    oracle (grid_sample form).  Parity unpinned by the reference (placeholder only); tolerance 1e-4 max-normalised (fp32
    atomics re-associate grad_value; bf16 value: the gradient w.r.t. the up-cast value, rounded once to bf16)."""
    from bevipm import ops
    shapes = [(9, 13), (7, 10), (5, 5), (12, 4)]
    value, loc, aw = _case(2, 37, M, D, shapes, 4, seed=11, spread=0.55)
    value = value.to(dtype)
    sh = torch.tensor(shapes, dtype=torch.int32)
    start = torch.tensor(np.concatenate([[0], np.cumsum([h * w for h, w in shapes])[:-1]]), dtype=torch.int64)
    g = torch.Generator().manual_seed(5)
    cot = torch.randn(2, 37, M * D, generator=g)
    # oracle side
    v0 = value.float().clone().requires_grad_(True)
    l0 = loc.clone().requires_grad_(True)
    a0 = aw.clone().requires_grad_(True)
    (dorc.deform_attn_grid_sample(v0, shapes, l0, a0) * cot).sum().backward()
    # ours
    v1 = value.cuda().requires_grad_(True)
    l1 = loc.cuda().requires_grad_(True)
    a1 = aw.cuda().requires_grad_(True)
    out = ops.deform_attn(v1, sh.cuda(), start.cuda(), l1, a1, out_dtype=torch.float32)
    (out * cot.cuda()).sum().backward()
    tol_v = 1e-4 if dtype == torch.float32 else 1e-2
    for name, got, want, tol in (("value", v1.grad.float().cpu(), v0.grad, tol_v), ("loc", l1.grad.cpu(), l0.grad, 1e-4), ("attn", a1.grad.cpu(), a0.grad, 1e-4)):
        assert got.shape == want.shape, name
        assert float((got - want).abs().max()) <= tol * float(want.abs().max()), name


@pytest.mark.gpu
def test_deform_attn_fusion_module_trains():
    """DeformAttnFusion is differentiable end to end: gradients reach the per-view maps and all four projections."""
    import bevipm
    torch.manual_seed(0)
    mod = bevipm.DeformAttnFusion(channels=32, views=3, heads=4, points=2).cuda()
    with torch.no_grad():
        mod.attention_weights.weight.normal_(0, 0.02)
        mod.sampling_offsets.weight.normal_(0, 0.02)
    x = torch.randn(1, 3, 32, 10, 14, device="cuda", requires_grad=True)
    y = mod(x)
    assert y.shape == (1, 32, 10, 14)
    y.square().mean().backward()
    assert x.grad is not None and float(x.grad.abs().max()) > 0
    for name, prm in mod.named_parameters():
        assert prm.grad is not None and torch.isfinite(prm.grad).all(), name
    assert float(mod.sampling_offsets.weight.grad.abs().max()) > 0 and float(mod.attention_weights.weight.grad.abs().max()) > 0


# ---- the fused, image-space form: queries sample the SOURCE maps around their IPM position (no per-view BEV maps) -------

def _rig(B, V, dev="cuda"):
    from bevipm import rig
    K, Rt = rig.look_at_rig(V, 3)
    return K[None].expand(B, -1, -1, -1).contiguous().to(dev), Rt[None].expand(B, -1, -1, -1).contiguous().to(dev)


def _identity_(mod):
    with torch.no_grad():
        for lin in (mod.value_proj, mod.output_proj):
            lin.weight.copy_(torch.eye(lin.weight.shape[0]))
            lin.bias.zero_()
        mod.sampling_offsets.weight.zero_()
        mod.sampling_offsets.bias.zero_()
        mod.attention_weights.weight.zero_()
        mod.attention_weights.bias.zero_()


@pytest.mark.gpu
@pytest.mark.parametrize("mask_invalid", [False, True])
def test_image_space_fusion_with_identity_parameters_is_the_reference_mean(mask_invalid):
    """Zero offsets, uniform weights, identity projections: the module must reproduce GeometryTransformer -> SimpleFusion('mean')
    (geometry.py:142-162 + fusion.py:20-21), i.e. the oracle-pinned fused kernel, up to the rounding of the two bilinear forms
    (1e-4 of the largest value); with mask_invalid the mean over the views that see the cell (fusion='mean_valid')."""
    import bevipm
    from bevipm import rig
    torch.manual_seed(0)
    B, V, C, fhw, bhw = 2, 5, 64, (27, 48), (24, 72)
    K, Rt = _rig(B, V)
    feats = torch.randn(B, V, C, *fhw, device="cuda")
    mod = bevipm.ImageSpaceDeformAttnFusion(bhw[0], bhw[1], rig.WILDTRACK_BOUNDS, C, V, heads=4, points=3, mask_invalid=mask_invalid).cuda()
    _identity_(mod)
    with torch.no_grad():
        out = mod(feats, K, Rt, img_size=rig.WILDTRACK_IMG_SIZE)
        want = bevipm.FusedIPM(bhw[0], bhw[1], rig.WILDTRACK_BOUNDS, fusion="mean_valid" if mask_invalid else "mean").cuda()(
            feats, K, Rt, img_size=rig.WILDTRACK_IMG_SIZE)
    assert out.shape == want.shape == (B, C, *bhw)
    Kp, Rp = bevipm.pack_calibration(K, Rt, B, V, "cuda")
    _, seen = mod.reference_points(Kp, Rp, fhw, rig.WILDTRACK_IMG_SIZE)
    assert 0.05 < float(seen.float().mean()) < 0.95          # the rig has views that miss cells: masking matters
    assert float((out - want).abs().max()) <= 1e-4 * float(want.abs().max())


@pytest.mark.gpu
def test_image_space_fusion_matches_oracle_composition():
    """Random parameters: the same projections on the CPU, the oracle's sample positions (pinned to the reference) as
    reference points and the from-spec deformable-attention oracle doing the sampling."""
    import bevipm
    from bevipm import rig
    from oracle import ipm_oracle as orc
    torch.manual_seed(1)
    B, V, C, fhw, bhw = 1, 3, 32, (20, 33), (12, 30)
    M, P, D = 4, 2, 8
    K, Rt = _rig(B, V, "cpu")
    feats = torch.randn(B, V, C, *fhw)
    mod = bevipm.ImageSpaceDeformAttnFusion(bhw[0], bhw[1], rig.WILDTRACK_BOUNDS, C, V, heads=M, points=P)
    with torch.no_grad():
        mod.sampling_offsets.weight.normal_(0, 0.05)
        mod.attention_weights.weight.normal_(0, 0.2)
    out = mod.cuda()(feats.cuda(), K.cuda(), Rt.cuda(), img_size=rig.WILDTRACK_IMG_SIZE).detach().cpu()
    mod = mod.cpu()
    xs, ys = rig.ground_axes(bhw[0], bhw[1], rig.WILDTRACK_BOUNDS)
    Kn, Rn = K.numpy(), Rt.numpy()
    with torch.no_grad():
        Q, Hf, Wf = bhw[0] * bhw[1], fhw[0], fhw[1]
        bev = torch.from_numpy(orc.warp_fuse(feats.numpy(), Kn, Rn, xs.numpy(), ys.numpy(), rig.WILDTRACK_IMG_SIZE, "mean"))
        query = bev.permute(0, 2, 3, 1).reshape(B, Q, C)
        ix, iy = orc.coords(Kn, Rn, xs.numpy(), ys.numpy(), fhw, rig.WILDTRACK_IMG_SIZE)
        ix, iy = torch.from_numpy(np.asarray(ix)).view(B, V, Q), torch.from_numpy(np.asarray(iy)).view(B, V, Q)
        ok = torch.isfinite(ix) & torch.isfinite(iy)
        ref = torch.stack([torch.where(ok, (ix + 0.5) / Wf, torch.full_like(ix, -1e3)),
                           torch.where(ok, (iy + 0.5) / Hf, torch.full_like(iy, -1e3))], -1).permute(0, 2, 1, 3)   # [B,Q,V,2]
        value = mod.value_proj(feats.permute(0, 1, 3, 4, 2).reshape(B, V * Hf * Wf, C)).view(B, V * Hf * Wf, M, D)
        off = mod.sampling_offsets(query).view(B, Q, M, V, P, 2)
        loc = ref.view(B, Q, 1, V, 1, 2) + off / torch.tensor([Wf, Hf], dtype=torch.float32)
        aw = torch.softmax(mod.attention_weights(query).view(B, Q, M, V * P), -1).view(B, Q, M, V, P)
        smp = dorc.deform_attn_grid_sample(value, [(Hf, Wf)] * V, loc, aw)
        want = mod.output_proj(smp).view(B, bhw[0], bhw[1], C).permute(0, 3, 1, 2)
    assert float((out - want).abs().max()) <= 3e-4 * float(want.abs().max())


@pytest.mark.gpu
def test_image_space_fusion_trains():
    import bevipm
    from bevipm import rig
    torch.manual_seed(0)
    B, V, C = 1, 3, 32
    K, Rt = _rig(B, V)
    mod = bevipm.ImageSpaceDeformAttnFusion(10, 14, rig.WILDTRACK_BOUNDS, C, V, heads=4, points=2).cuda()
    with torch.no_grad():
        mod.attention_weights.weight.normal_(0, 0.02)
        mod.sampling_offsets.weight.normal_(0, 0.02)
    x = torch.randn(B, V, C, 12, 20, device="cuda", requires_grad=True)
    y = mod(x, K, Rt, img_size=rig.WILDTRACK_IMG_SIZE)
    assert y.shape == (B, C, 10, 14)
    y.square().mean().backward()
    assert x.grad is not None and float(x.grad.abs().max()) > 0
    for name, prm in mod.named_parameters():
        assert prm.grad is not None and torch.isfinite(prm.grad).all(), name
    assert float(mod.sampling_offsets.weight.grad.abs().max()) > 0 and float(mod.attention_weights.weight.grad.abs().max()) > 0
