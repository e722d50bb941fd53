"""FoldedConcatProjIPM: the 1x1 projection of BEVNet folded in front of the warp must equal
GeometryTransformer -> ConcatFusion -> proj (model_wrapper.py:68-73) up to fp32 reassociation, forward and backward."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("co", [128, 24])
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
