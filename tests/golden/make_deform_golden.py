#!/usr/bin/env python
"""Golden vectors for the Phase-2 deformable-attention sampling.

The reference has no implementation of this row (fusion.py:25-36 is a placeholder), so the pin is an independent, published
one: `transformers` (HuggingFace, version recorded in the fixture) ships the PyTorch port of Deformable-DETR's
`ms_deform_attn_core_pytorch` as `multi_scale_deformable_attention` -- per level F.grid_sample(bilinear, zeros,
align_corners=False) on 2*loc-1, weighted sum over levels x points.  This script calls THAT function (unmodified, imported
from the installed package) on seeded inputs and freezes inputs, output and autograd gradients.

    python tests/golden/make_deform_golden.py        # writes tests/golden/deform_attn_hf.npz
"""
from pathlib import Path

import numpy as np
import torch
import transformers
from transformers.models.mask2former.modeling_mask2former import multi_scale_deformable_attention as hf_msda

OUT = Path(__file__).resolve().parent / "deform_attn_hf.npz"


def case(B, Q, M, D, shapes, P, seed, spread):
    g = torch.Generator().manual_seed(seed)
    S = sum(h * w for h, w in shapes)
    L = len(shapes)
    value = torch.randn(B, S, M, D, generator=g)
    loc = 0.5 + spread * (torch.rand(B, Q, M, L, P, 2, generator=g) - 0.5) * 2          # some samples land outside [0, 1]
    aw = torch.softmax(torch.randn(B, Q, M, L * P, generator=g), -1).view(B, Q, M, L, P)
    cot = torch.randn(B, Q, M * D, generator=g)
    v, l, a = value.clone().requires_grad_(True), loc.clone().requires_grad_(True), aw.clone().requires_grad_(True)
    out = hf_msda(v, shapes, l, a)
    (out * cot).sum().backward()
    return {"value": value, "loc": loc, "aw": aw, "cot": cot, "out": out.detach(), "g_value": v.grad, "g_loc": l.grad, "g_aw": a.grad,
            "shapes": torch.tensor(shapes)}


CASES = {
    "views7": dict(B=1, Q=40, M=8, D=32, shapes=[(6, 10)] * 7, P=4, seed=1, spread=0.55),       # config 3's structure, small maps
    "ragged": dict(B=2, Q=37, M=3, D=8, shapes=[(9, 13), (7, 10), (5, 5), (12, 4)], P=4, seed=2, spread=0.6),
    "wide_head": dict(B=1, Q=19, M=2, D=128, shapes=[(6, 7), (3, 11)], P=2, seed=3, spread=0.7),
}

if __name__ == "__main__":
    blob = {"transformers_version": np.array(transformers.__version__), "torch_version": np.array(torch.__version__)}
    for name, kw in CASES.items():
        for k, t in case(**kw).items():
            blob[f"{name}.{k}"] = t.numpy()
    np.savez_compressed(OUT, **blob)
    print(OUT, OUT.stat().st_size, "bytes")
