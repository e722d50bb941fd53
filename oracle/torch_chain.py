"""The reference's op chain re-stated with the SAME ATen CPU operators (TEST INFRASTRUCTURE ONLY).

/root/reference does not travel to the GPU box, so this file re-expresses, in our own words,
what geometry.py:142-162 and fusion.py:17-22 ask ATen to do -- aten::mm for the homography and
the projection, the pointwise chain, aten::grid_sampler_2d (bilinear, zeros,
align_corners=False), aten::sum/mean/amax over the view axis -- so that

  * tests can cross-check the C oracle against ATen on any host (tests/test_oracle.py), and
  * `bench.py --impl reference` can report, next to the multi-threaded C port, what the
    reference's own library calls cost on the box's CPU (informational key `torch_cpu_chain`), and
  * `bench.py` can time the same chain on ATen's CUDA kernels on the very B200 our kernel runs on (informational key
    `torch_gpu_chain`: the stock-torch arm of BASELINE.md 4.4, ~175 launches per frame).

It is never imported by the product package.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def ground_homography(K: torch.Tensor, Rt: torch.Tensor) -> torch.Tensor:
    """geometry.py:60-63: columns r1, r2, t of the extrinsics, times K."""
    cols = Rt[:3, [0, 1, 3]]
    return K[:3, :3] @ cols


def warp_views(feats: torch.Tensor, K: torch.Tensor, Rt: torch.Tensor, xs: torch.Tensor, ys: torch.Tensor,
               img_size) -> torch.Tensor:
    """feats [B,V,C,Hf,Wf] fp32 -> per-view BEV maps [B,V,C,Hb,Wb] (one grid_sample per view), on the device of `feats`
    (CPU = the reference's path in this image; CUDA = the same chain on ATen's CUDA kernels, the stock-torch arm)."""
    B, V, C, Hf, Wf = feats.shape
    Hb, Wb = ys.numel(), xs.numel()
    img_h, img_w = img_size
    dev = feats.device
    K, Rt, xs, ys = K.to(dev), Rt.to(dev), xs.to(dev), ys.to(dev)
    yy, xx = torch.meshgrid(ys, xs, indexing="ij")
    pts = torch.stack([xx, yy, torch.ones_like(xx)], dim=-1).reshape(-1, 3).T      # [3, Hb*Wb]
    out = torch.zeros(B, V, C, Hb, Wb, device=dev)
    for b in range(B):
        for v in range(V):
            uvw = ground_homography(K[b, v], Rt[b, v]) @ pts
            w = uvw[2:3]
            w = torch.where(w.abs() < 1e-6, torch.ones_like(w), w)
            px = torch.stack([(uvw[0:1] / w).squeeze(0), (uvw[1:2] / w).squeeze(0)], dim=1).reshape(Hb, Wb, 2)
            px[..., 0] = px[..., 0] * (Wf / float(img_w))
            px[..., 1] = px[..., 1] * (Hf / float(img_h))
            px[..., 0] = (px[..., 0] + 0.5) / Wf * 2.0 - 1.0
            px[..., 1] = (px[..., 1] + 0.5) / Hf * 2.0 - 1.0
            out[b, v] = F.grid_sample(feats[b, v][None], px[None], mode="bilinear", padding_mode="zeros",
                                      align_corners=False)[0]
    return out


def fuse(per_view: torch.Tensor, mode: str) -> torch.Tensor:
    """fusion.py:17-22 / :43-46."""
    if mode == "sum":
        return per_view.sum(dim=1)
    if mode == "mean":
        return per_view.mean(dim=1)
    if mode == "max":
        return per_view.max(dim=1).values
    if mode == "concat":
        B, V, C, H, W = per_view.shape
        return per_view.reshape(B, V * C, H, W)
    return per_view


def warp_fuse(feats, K, Rt, xs, ys, img_size, mode="mean"):
    with torch.no_grad():
        return fuse(warp_views(feats, K, Rt, xs, ys, img_size), mode)
