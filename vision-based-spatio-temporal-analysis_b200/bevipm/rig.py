"""Synthetic Wildtrack-shaped inputs for the IPM warp+fuse path.

The reference has no synthetic data generator (its loader reads the real
Wildtrack folder, /root/reference/project/data/wildtrack_loader.py:154-247).
This module builds the stand-in that SURVEY.md section 8(d) specifies: seven
look-at pinhole cameras around the ground rectangle of
/root/reference/project/configs/wildtrack.yaml:12-13, producing the same
calibration *format* the loader hands to the model (3x3 intrinsics in image
pixels, 4x4 world->camera extrinsics in metres).

Nothing here touches the GPU; it is host-side setup shared by tests, bench.py
and the smoke entry.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Tuple

import numpy as np
import torch

WILDTRACK_BOUNDS = (-24.0, 24.0, -7.2, 7.2)   # wildtrack.yaml:13  (x_min, x_max, y_min, y_max)
WILDTRACK_IMG_SIZE = (1080, 1920)             # geometry.py:83 default img_size (H, W)


@dataclass(frozen=True)
class Workload:
    """One BASELINE.json configuration of the hot path (shapes only)."""
    name: str
    frames: int          # B
    views: int           # V
    channels: int        # C
    feat_hw: Tuple[int, int]
    bev_hw: Tuple[int, int]
    dtype: str           # "f32" | "bf16"  (feature storage type)
    out_dtype: str       # "f32" | "bf16"
    bounds: Tuple[float, float, float, float] = WILDTRACK_BOUNDS
    img_size: Tuple[int, int] = WILDTRACK_IMG_SIZE
    fusion: str = "mean"

    @property
    def feat_elem_bytes(self) -> int:
        return 4 if self.dtype == "f32" else 2

    @property
    def out_elem_bytes(self) -> int:
        return 4 if self.out_dtype == "f32" else 2


# BASELINE.json "configs" 0..2 and 4 (config 3 is the deformable-attention follow-on).
WORKLOADS = {
    "c1": Workload("c1", 1, 7, 512, (135, 240), (120, 360), "f32", "f32"),
    "c2": Workload("c2", 8, 7, 1024, (135, 240), (120, 360), "bf16", "bf16"),
    "c3": Workload("c3", 1, 7, 128, (270, 480), (480, 1440), "f32", "f32"),
    "c5": Workload("c5", 64, 7, 512, (135, 240), (120, 360), "f32", "f32"),
}


def look_at_rig(views: int = 7, seed: int = 0):
    """K [V,3,3] and Rt [V,4,4] (float32) for `views` cameras ringed around the ground patch.

    Camera v sits at (30 cos th, 12 sin th, 3 + 2u) with th = 2 pi v / views + 0.3 u and looks
    at a ground target (10u-5, 4u-2, 0); principal point (960, 540), focal 1000 + 800u
    (SURVEY.md 8(d)).  `u` are successive draws of a seeded CPU torch.Generator.
    """
    g = torch.Generator().manual_seed(seed)

    def u() -> float:
        return float(torch.rand((), generator=g))

    Ks, Rts = [], []
    up = np.array([0.0, 0.0, 1.0])
    for v in range(views):
        th = 2.0 * math.pi * v / views + 0.3 * u()
        cam = np.array([30.0 * math.cos(th), 12.0 * math.sin(th), 3.0 + 2.0 * u()])
        tgt = np.array([10.0 * u() - 5.0, 4.0 * u() - 2.0, 0.0])
        f = 1000.0 + 800.0 * u()
        z = tgt - cam
        z /= np.linalg.norm(z)
        x = np.cross(z, up)
        x /= np.linalg.norm(x)
        y = np.cross(z, x)
        R = np.stack([x, y, z], axis=0)
        t = -R @ cam
        Rt = np.eye(4)
        Rt[:3, :3] = R
        Rt[:3, 3] = t
        K = np.array([[f, 0.0, 960.0], [0.0, f, 540.0], [0.0, 0.0, 1.0]])
        Ks.append(K)
        Rts.append(Rt)
    K = torch.from_numpy(np.stack(Ks)).to(torch.float32)
    Rt = torch.from_numpy(np.stack(Rts)).to(torch.float32)
    return K, Rt


def ground_axes(bev_h: int, bev_w: int, bounds) -> Tuple[torch.Tensor, torch.Tensor]:
    """Cell-centre world coordinates xs[Wb], ys[Hb] exactly as geometry.py:24-27 forms them.

    They must come from torch.linspace (not the closed form): 152 of 360 x entries differ from
    x_min + (j + 1/2) res_x by one ulp.
    """
    x_min, x_max, y_min, y_max = bounds
    res_x = (x_max - x_min) / bev_w
    res_y = (y_max - y_min) / bev_h
    xs = torch.linspace(x_min + 0.5 * res_x, x_max - 0.5 * res_x, bev_w)
    ys = torch.linspace(y_min + 0.5 * res_y, y_max - 0.5 * res_y, bev_h)
    return xs, ys


def synthetic_features(wl: Workload, seed: int = 0, frames: int | None = None) -> torch.Tensor:
    """N(0,1) features, logical shape [B,V,C,Hf,Wf], NCHW-contiguous fp32 on the CPU."""
    g = torch.Generator().manual_seed(seed)
    B = wl.frames if frames is None else frames
    return torch.randn(B, wl.views, wl.channels, *wl.feat_hw, generator=g)


def algorithmic_bytes(ix: np.ndarray, iy: np.ndarray, feat_hw, channels: int,
                      feat_elem_bytes: int, out_elem_bytes: int, per_view_out: bool = False) -> dict:
    """SURVEY.md 8(d) byte model for ONE frame, recomputed from the actual sample coordinates.

    ix, iy: float32 [V, Hb, Wb] source coordinates in feature pixels (the oracle's step (7)).
    B_alg = bytes(out BEV) + sum_v (unique source texels hit by >= 1 in-bounds tap) * C * elem.
    B_full = bytes(all input) + bytes(out).
    """
    V, Hb, Wb = ix.shape
    Hf, Wf = feat_hw
    touched = 0
    taps_in = 0
    for v in range(V):
        fin = np.isfinite(ix[v]) & np.isfinite(iy[v])
        x0 = np.floor(np.where(fin, ix[v], -10.0)).astype(np.int64)
        y0 = np.floor(np.where(fin, iy[v], -10.0)).astype(np.int64)
        hit = np.zeros((Hf, Wf), dtype=bool)
        for dy in (0, 1):
            for dx in (0, 1):
                xx, yy = x0 + dx, y0 + dy
                ok = fin & (xx >= 0) & (xx < Wf) & (yy >= 0) & (yy < Hf)
                hit[yy[ok], xx[ok]] = True
                taps_in += int(ok.sum())
        touched += int(hit.sum())
    out_maps = V if per_view_out else 1
    out_bytes = out_maps * channels * Hb * Wb * out_elem_bytes
    in_touched = touched * channels * feat_elem_bytes
    in_full = V * channels * Hf * Wf * feat_elem_bytes
    return {
        "out_bytes": out_bytes,
        "touched_texels": touched,
        "touched_in_bytes": in_touched,
        "taps_in_bounds": taps_in,
        "taps_total": 4 * V * Hb * Wb,
        "b_alg": out_bytes + in_touched,
        "b_full": out_bytes + in_full,
    }
