"""Generate tests/golden/*.npz by running the REFERENCE module itself (CPU fp32).

Run in the build container only (needs /root/reference, which is absent on the GPU box):

    python tests/golden/make_golden.py

Every case stores the seeded inputs and what
  GeometryTransformer(...)(feats, K, Rt, img_size)        geometry.py:80-163  -> "none"
  SimpleFusion(mode)(that)                                fusion.py:17-22     -> "sum" / "mean" / "max"
  ConcatFusion()(that)                                    fusion.py:43-46     -> "concat"
returned, plus autograd gradients w.r.t. the features for mean and sum (the path train.py:243
exercises).  The reference is imported unmodified; kornia is not installed so warp_impl='kornia'
(what BEVNet asks for, model_wrapper.py:42) takes the grid_sample branch, geometry.py:142-162.
"""
import math
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
sys.path.insert(0, "/root/reference/project")
sys.path.insert(0, str(ROOT / "vision-based-spatio-temporal-analysis_b200"))

from models.fusion.geometry import GeometryTransformer  # noqa: E402  (reference)
from models.fusion.fusion import SimpleFusion, ConcatFusion  # noqa: E402  (reference)
from bevipm import rig  # noqa: E402


def run_case(name, feats, K, Rt, bev_hw, bounds, img_size, extra=None):
    geom = GeometryTransformer(bev_hw[0], bev_hw[1], bounds, warp_impl="kornia")
    feats = feats.clone().requires_grad_(True)
    per_view = geom(feats, K, Rt, img_size=img_size)
    out = {"none": per_view.detach().numpy()}
    grads = {}
    for mode in ("sum", "mean", "max"):
        out[mode] = SimpleFusion(mode)(per_view).detach().numpy()
    out["concat"] = ConcatFusion()(per_view).detach().numpy()
    # gradients: d/dfeats of <fused, gw> with a seeded cotangent
    g = torch.Generator().manual_seed(1234)
    gw = torch.randn(out["mean"].shape, generator=g)
    for mode in ("sum", "mean"):
        (grad,) = torch.autograd.grad((SimpleFusion(mode)(per_view) * gw).sum(), feats, retain_graph=True)
        grads["grad_" + mode] = grad.numpy()
    gwv = torch.randn(per_view.shape, generator=g)
    (grad,) = torch.autograd.grad((per_view * gwv).sum(), feats)
    grads["grad_none"] = grad.numpy()
    xs = geom.ground_grid[0, :, 0].numpy()
    ys = geom.ground_grid[:, 0, 1].numpy()
    np.savez_compressed(
        HERE / f"{name}.npz",
        feats=feats.detach().numpy(), K=np.asarray(K), Rt=np.asarray(Rt), xs=xs, ys=ys,
        bounds=np.asarray(bounds, np.float64), img_size=np.asarray(img_size, np.int64),
        bev_hw=np.asarray(bev_hw, np.int64), cotangent=gw.numpy(), cotangent_none=gwv.numpy(),
        **{"out_" + k: v for k, v in out.items()}, **grads, **(extra or {}))
    print(f"{name}: feats {tuple(feats.shape)} -> bev {bev_hw}, "
          f"nonzero {np.count_nonzero(out['none']) / out['none'].size:.2f}")


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)
    bounds = rig.WILDTRACK_BOUNDS

    # 1. Wildtrack-like rig, small: per-frame calibration [B,V,3,3] / [B,V,4,4]
    K, Rt = rig.look_at_rig(3, seed=0)
    K2 = torch.stack([K, K * torch.tensor([[1.1, 1, 1], [1, 1.1, 1], [1, 1, 1]])])
    Rt2 = torch.stack([Rt, Rt.clone()])
    Rt2[1, :, :3, 3] += 0.25
    feats = torch.randn(2, 3, 6, 27, 48)
    run_case("rig_small", feats, K2, Rt2, (24, 72), bounds, (1080, 1920))

    # 2. seven views, odd channel count, 3x4 extrinsics, img_size != default
    K, Rt = rig.look_at_rig(7, seed=3)
    K = K.clone()
    K[:, :2] *= 0.5  # calibration for a 540x960 image
    feats = torch.randn(1, 7, 5, 17, 30)
    run_case("seven_views_3x4", feats, K[None], Rt[None, :, :3, :], (20, 60), bounds, (540, 960))

    # 3. degenerate geometry: a view entirely out of frame, a camera inside the patch looking
    #    along the ground (cells behind it, |w| tiny along its principal plane), identity-like K
    K, Rt = rig.look_at_rig(4, seed=5)
    K, Rt = K.clone(), Rt.clone()
    Rt[0, :3, 3] += torch.tensor([1.0e4, 0.0, 0.0])       # view 0: projects far outside
    R = torch.tensor([[0.0, -1.0, 0.0], [0.0, 0.0, -1.0], [1.0, 0.0, 0.0]])  # looks along +x
    C = torch.tensor([0.0, 0.0, 1.5])
    Rt[1] = torch.eye(4)
    Rt[1, :3, :3] = R
    Rt[1, :3, 3] = -R @ C
    K[2] = torch.tensor([[40.0, 0.0, 960.0], [0.0, 40.0, 540.0], [0.0, 0.0, 1.0]])
    Rt[3] = torch.eye(4)                                  # w == 1 everywhere, u = x, v = y
    K[3] = torch.tensor([[40.0, 0.0, 960.0], [0.0, 75.0, 540.0], [0.0, 0.0, 1.0]])
    feats = torch.randn(1, 4, 4, 34, 60)
    run_case("degenerate", feats, K[None], Rt[None], (30, 90), bounds, (1080, 1920))

    # 4. w exactly 0 / below the 1e-6 guard: H row 2 = (0, 0, 5e-7) -> w_safe = 1
    K = torch.eye(3)[None, None].repeat(1, 2, 1, 1)
    Rt = torch.eye(4)[None, None].repeat(1, 2, 1, 1)
    K[0, 0] = torch.tensor([[30.0, 0.0, 700.0], [0.0, 30.0, 300.0], [0.0, 0.0, 5e-7]])
    K[0, 1] = torch.tensor([[30.0, 0.0, 700.0], [0.0, 30.0, 300.0], [0.0, 0.0, 0.5]])
    feats = torch.randn(1, 2, 3, 40, 64)
    run_case("w_guard", feats, K, Rt, (16, 40), (-20.0, 20.0, -8.0, 8.0), (640, 1024))

    # 5. non-finite coordinates: focal overflow -> inf/NaN sample positions
    K = torch.eye(3)[None, None].repeat(1, 2, 1, 1)
    Rt = torch.eye(4)[None, None].repeat(1, 2, 1, 1)
    K[0, 0] = torch.tensor([[3.0e38, 0.0, 0.0], [0.0, 3.0e38, 0.0], [0.0, 0.0, 1.0]])
    K[0, 1] = torch.tensor([[25.0, 0.0, 500.0], [0.0, 25.0, 200.0], [0.0, 0.0, 1.0]])
    feats = torch.randn(1, 2, 2, 20, 32)
    run_case("non_finite", feats, K, Rt, (8, 20), (-20.0, 20.0, -8.0, 8.0), (640, 1024))


if __name__ == "__main__":
    main()
