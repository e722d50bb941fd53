"""FoldedConcatProjIPM: the 1x1 projection of BEVNet folded in front of the warp must equal
GeometryTransformer -> ConcatFusion -> proj (model_wrapper.py:68-73) up to fp32 reassociation, forward and backward."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("co", [128, 24, 48, 256])
@pytest.mark.parametrize("bias", [True, False])
def test_folded_projection_equals_concat_then_conv(monkeypatch, co, bias):
    import bevipm
    from bevipm import rig
    torch.manual_seed(0)
    # cuDNN convolutions default to TF32 (1e-3 relative); compare both orders in true fp32
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    monkeypatch.setattr(torch.backends.cuda.matmul, "allow_tf32", False)
    B, V, C, fhw, bhw = 2, 3, 16, (27, 48), (24, 72)
    K, Rt = rig.look_at_rig(V, 1)
    K = K[None].expand(B, -1, -1, -1).contiguous().to(DEV)
    Rt = Rt[None].expand(B, -1, -1, -1).contiguous().to(DEV)
    feats = torch.randn(B, V, C, *fhw, device=DEV)
    proj = torch.nn.Conv2d(V * C, co, kernel_size=1, bias=bias).to(DEV)

    f1 = feats.clone().requires_grad_(True)
    geom = bevipm.GeometryTransformer(bhw[0], bhw[1], rig.WILDTRACK_BOUNDS, warp_impl="kornia").to(DEV)
    ref = proj(bevipm.ConcatFusion()(geom(f1, K, Rt, img_size=rig.WILDTRACK_IMG_SIZE)))      # the reference's order
    cot = torch.randn_like(ref)
    (ref * cot).sum().backward()
    gw_ref, gf_ref = proj.weight.grad.clone(), f1.grad.clone()
    gb_ref = proj.bias.grad.clone() if bias else None
    proj.zero_grad()

    f2 = feats.clone().requires_grad_(True)
    folded = bevipm.FoldedConcatProjIPM(bhw[0], bhw[1], rig.WILDTRACK_BOUNDS, proj, views=V).to(DEV)
    out = folded(f2, K, Rt, img_size=rig.WILDTRACK_IMG_SIZE)
    assert out.shape == ref.shape == (B, co, *bhw)
    tol = lambda a: 1e-4 * float(a.abs().max())
    assert float((out - ref).detach().abs().max()) <= tol(ref.detach())
    (out * cot).sum().backward()
    assert float((proj.weight.grad - gw_ref).abs().max()) <= tol(gw_ref)
    assert float((f2.grad - gf_ref).abs().max()) <= tol(gf_ref)
    if bias:
        assert float((proj.bias.grad - gb_ref).abs().max()) <= tol(gb_ref)
    assert len(list(folded.state_dict().keys())) == (2 if bias else 1)   # only the projection's own parameters
    assert folded.last_gemm == ("tcgen05" if co % 16 == 0 else "cublas")


def _proj_case(BV, V, rows, C, Co, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    x = torch.randn(BV, rows, C, device=DEV, generator=g)
    w = torch.randn(Co, V, C, device=DEV, generator=g) / C ** 0.5
    want = torch.einsum("bvrc,ovc->bvro", x.double().view(BV // V, V, rows, C), w.double()).reshape(BV, rows, Co)
    return x, w, want


@pytest.mark.parametrize("BV,V,rows,C,Co", [
    (7, 7, 135 * 240, 1280, 128),     # wildtrack.yaml: FEAT_DIM 1280 -> 128, one frame
    (6, 3, 1000, 96, 16),             # ragged last tile, small N
    (2, 1, 128, 32, 32),              # exactly one tile, one k-block
    (4, 2, 129, 36, 48),              # one row past a tile; C not a multiple of the k-block (zero-filled by the copy engine)
    (1, 1, 77, 4, 256),               # a single map (extent-1 outer dimension), widest N, one partial k-block
    (5, 5, 300, 520, 240),            # N not a power of two
    (3, 3, 129, 64, 128),             # transposed form: a pair and a single per map (2 units), second unit one row
    (2, 2, 128 * 5 + 17, 96, 128),    # transposed form: odd number of units per map, ragged last unit
    (1, 1, 100, 32, 128),             # transposed form: a single unit smaller than its box
    (3, 3, 4099, 64, 80),             # N = 80: the last 32-column store is clipped at 80
])
@pytest.mark.parametrize("passes", [3, 1])
def test_proj1x1_kernel_against_float64(BV, V, rows, C, Co, passes):
    """The tcgen05 GEMM (csrc/bevipm_proj.cu) against a float64 einsum: split operands within 3e-5 of the largest value at
    K = 1280 (the tensor core's fp32 accumulation is not round-to-nearest; measured 1.1e-5), one TF32 pass within 2e-3."""
    from bevipm import ops
    x, w, want = _proj_case(BV, V, rows, C, Co, seed=BV + rows)
    out = ops.proj1x1(x, w, passes)
    torch.cuda.synchronize()
    err = float((out.double() - want).abs().max() / want.abs().max())
    assert err <= (3e-5 if passes == 3 else 2e-3), err
    if passes == 3:   # and better than the single pass by orders of magnitude
        err1 = float((ops.proj1x1(x, w, 1).double() - want).abs().max() / want.abs().max())
        assert C < 16 or err < err1 / 8, (err, err1)


def test_proj1x1_strided_views_and_slices():
    """x rows / maps with padding between them and an output that is a channel slice of a wider tensor (how the input
    gradient is written, 256 columns per launch)."""
    from bevipm import ops
    BV, V, rows, C, Co = 4, 2, 500, 64, 32
    x, w, want = _proj_case(BV, V, rows, C, Co, seed=3)
    xp = torch.zeros(BV, rows + 3, C + 8, device=DEV)
    xp[:, :rows, :C] = x
    wide = torch.full((BV, rows, 3 * Co), 7.0, device=DEV)
    ops._proj1x1_raw(xp[:, :rows, :C], w, wide[:, :, Co:2 * Co], 3)
    torch.cuda.synchronize()
    assert float((wide[:, :, Co:2 * Co].double() - want).abs().max() / want.abs().max()) <= 3e-5
    assert bool((wide[:, :, :Co] == 7.0).all()) and bool((wide[:, :, 2 * Co:] == 7.0).all())   # neighbours untouched


@pytest.mark.parametrize("C,Co", [(96, 32), (520, 128), (64, 16)])
def test_proj1x1_gradients(C, Co):
    from bevipm import ops
    BV, V, rows = 4, 2, 700
    x, w, _ = _proj_case(BV, V, rows, C, Co, seed=11)
    x1, w1 = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    cot = torch.randn(BV, rows, Co, device=DEV)
    (ops.proj1x1(x1, w1, 3) * cot).sum().backward()
    x2, w2 = x.double().requires_grad_(True), w.double().requires_grad_(True)
    (torch.einsum("bvrc,ovc->bvro", x2.view(BV // V, V, rows, C), w2).reshape(BV, rows, Co) * cot.double()).sum().backward()
    assert float((x1.grad.double() - x2.grad).abs().max() / x2.grad.abs().max()) <= 3e-5
    assert float((w1.grad.double() - w2.grad).abs().max() / w2.grad.abs().max()) <= 1e-4


def test_proj1x1_rejects_what_it_cannot_run():
    from bevipm import ops
    x = torch.randn(2, 64, 32, device=DEV)
    with pytest.raises(ValueError):
        ops.proj1x1(x, torch.randn(24, 1, 32, device=DEV))          # Co not a multiple of 16
    with pytest.raises(ValueError):
        ops.proj1x1(x[:, :, :30], torch.randn(16, 1, 30, device=DEV))  # C not a multiple of 4
    with pytest.raises(RuntimeError):
        ops.proj1x1(x.cpu(), torch.randn(16, 1, 32))
