"""From-spec CPU oracle of multi-scale deformable attention (TEST INFRASTRUCTURE ONLY).

PARITY UNPINNED BY THE REFERENCE: /root/reference has only a placeholder for attention fusion
(project/models/fusion/fusion.py:25-36 prints once and returns the mean); there is no reference
implementation, test or golden vector to pin against.  What is restated here is the published algorithm
of Deformable-DETR's MSDeformAttn (`ms_deform_attn_core_pytorch`: per level, grid_sample(bilinear,
zeros, align_corners=False) at 2*loc-1, times the attention weights, summed over levels x points), which
MVDeTr applies with one level per camera view.  Two independent forms are kept so they can check each
other: the grid_sample form and a scalar numpy form of the CUDA reference kernel's formula
(h_im = loc_y*H - 0.5, per-tap bounds, w1*v1 + w2*v2 + w3*v3 + w4*v4, times the weight).

PINNED TO A PUBLISHED THIRD-PARTY IMPLEMENTATION instead: tests/golden/deform_attn_hf.npz holds inputs, outputs and autograd
gradients of `transformers` 5.5.0's `multi_scale_deformable_attention` (models/mask2former/modeling_mask2former.py, the
PyTorch port of the algorithm above), produced by tests/golden/make_deform_golden.py from the unmodified installed package;
tests/test_deform_attn.py::test_oracle_is_pinned_to_the_published_implementation checks both forms against it.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def deform_attn_grid_sample(value, spatial_shapes, sampling_locations, attention_weights):
    """value [B,S,M,D], spatial_shapes list[(H,W)], loc [B,Q,M,L,P,2], attn [B,Q,M,L,P] -> [B,Q,M*D] (fp32 CPU)."""
    value = value.float()
    B, S, M, D = value.shape
    _, Q, _, L, P, _ = sampling_locations.shape
    parts = value.split([h * w for h, w in spatial_shapes], dim=1)
    grids = 2 * sampling_locations.float() - 1
    sampled = []
    for lid, (h, w) in enumerate(spatial_shapes):
        v = parts[lid].flatten(2).transpose(1, 2).reshape(B * M, D, h, w)
        g = grids[:, :, :, lid].transpose(1, 2).flatten(0, 1)            # [B*M, Q, P, 2]
        sampled.append(F.grid_sample(v, g, mode="bilinear", padding_mode="zeros", align_corners=False))
    aw = attention_weights.float().transpose(1, 2).reshape(B * M, 1, Q, L * P)
    out = (torch.stack(sampled, dim=-2).flatten(-2) * aw).sum(-1).view(B, M * D, Q)
    return out.transpose(1, 2).contiguous()


def deform_attn_scalar(value, spatial_shapes, sampling_locations, attention_weights):
    """Same, as explicit loops in float64 (small cases only)."""
    v = value.double().numpy()
    loc = sampling_locations.double().numpy()
    aw = attention_weights.double().numpy()
    B, S, M, D = v.shape
    _, Q, _, L, P, _ = loc.shape
    starts = np.concatenate([[0], np.cumsum([h * w for h, w in spatial_shapes])])
    out = np.zeros((B, Q, M, D))
    for b in range(B):
        for q in range(Q):
            for m in range(M):
                for l, (H, W) in enumerate(spatial_shapes):
                    lv = v[b, starts[l]:starts[l + 1], m].reshape(H, W, D)
                    for p in range(P):
                        x = loc[b, q, m, l, p, 0] * W - 0.5
                        y = loc[b, q, m, l, p, 1] * H - 0.5
                        if not (y > -1 and x > -1 and y < H and x < W):
                            continue
                        x0, y0 = int(np.floor(x)), int(np.floor(y))
                        lx, ly = x - x0, y - y0
                        acc = np.zeros(D)
                        for dy, dx, wgt in ((0, 0, (1 - ly) * (1 - lx)), (0, 1, (1 - ly) * lx), (1, 0, ly * (1 - lx)), (1, 1, ly * lx)):
                            yy, xx = y0 + dy, x0 + dx
                            if 0 <= yy < H and 0 <= xx < W:
                                acc += wgt * lv[yy, xx]
                        out[b, q, m] += aw[b, q, m, l, p] * acc
    return torch.from_numpy(out.reshape(B, Q, M * D)).float()
