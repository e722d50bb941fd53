"""TEST INFRASTRUCTURE ONLY -- from-spec restatement of the reference's kornia branch, parity UNPINNED.

/root/reference/project/models/fusion/geometry.py:124-141 calls kornia.geometry.transform.warp_perspective when
kornia is installed.  kornia is a third-party dependency that is absent from this image (and unpinned by the
reference: it is not even listed in its README), so neither the reference's branch nor kornia itself can be executed
here.  This module restates kornia's published algorithm for warp_perspective (normalize_homography with (size-1),
inverse, create_meshgrid(normalized) = linspace(-1, 1), transform_points, F.grid_sample) over the same M the
reference builds, in plain torch on the CPU.  It is the checker for the `emulate_kornia=True` compatibility mode of
bevipm.GeometryTransformer / FusedIPM; tolerance-based, never bit-exact, and it says nothing about what a particular
kornia release does beyond that algorithm.
"""
import torch
import torch.nn.functional as F


def _normal_transform_pixel(h: int, w: int) -> torch.Tensor:
    eps = 1e-14
    wd = eps if w == 1 else float(w - 1)
    hd = eps if h == 1 else float(h - 1)
    return torch.tensor([[2.0 / wd, 0.0, -1.0], [0.0, 2.0 / hd, -1.0], [0.0, 0.0, 1.0]], dtype=torch.float32)


def _warp_perspective(src: torch.Tensor, M: torch.Tensor, dsize) -> torch.Tensor:
    """src [1,C,H,W], M [3,3] (dst_pix <- src_pix), bilinear / zeros / align_corners=False."""
    _, _, H, W = src.shape
    h_out, w_out = dsize
    src_norm_trans_src_pix = _normal_transform_pixel(H, W)
    dst_norm_trans_dst_pix = _normal_transform_pixel(h_out, w_out)
    dst_norm_trans_src_norm = dst_norm_trans_dst_pix @ (M @ torch.linalg.inv(src_norm_trans_src_pix))
    src_norm_trans_dst_norm = torch.linalg.inv(dst_norm_trans_src_norm)
    xs = torch.linspace(-1.0, 1.0, w_out)
    ys = torch.linspace(-1.0, 1.0, h_out)
    gy, gx = torch.meshgrid(ys, xs, indexing="ij")
    pts = torch.stack([gx, gy, torch.ones_like(gx)], dim=-1)                   # [h,w,3]
    out = pts @ src_norm_trans_dst_norm.T
    z = out[..., 2:3]
    scale = torch.where(z.abs() > 1e-8, 1.0 / (z + 1e-8), torch.ones_like(z))
    grid = (scale * out[..., :2]).unsqueeze(0)
    return F.grid_sample(src, grid, mode="bilinear", padding_mode="zeros", align_corners=False)


def warp_views(feats: torch.Tensor, K: torch.Tensor, Rt: torch.Tensor, bev_hw, bounds, img_size) -> torch.Tensor:
    """feats [B,V,C,Hf,Wf] fp32 CPU, K [B,V,3,3], Rt [B,V,4,4] -> [B,V,C,Hb,Wb] as geometry.py:124-141 would fill it."""
    B, V, C, Hf, Wf = feats.shape
    Hb, Wb = bev_hw
    H_img, W_img = img_size
    min_x, max_x, min_y, max_y = bounds
    res_x, res_y = (max_x - min_x) / Wb, (max_y - min_y) / Hb
    out = torch.zeros(B, V, C, Hb, Wb)
    for b in range(B):
        for v in range(V):
            R, t = Rt[b, v, :3, :3], Rt[b, v, :3, 3]
            G = torch.stack([R[:, 0], R[:, 1], t], dim=1)
            H_w2i = K[b, v] @ G                                                  # geometry.py:60-63
            H_i2w = torch.linalg.inv(H_w2i)                                      # :66-78
            S = torch.tensor([[W_img / float(Wf), 0.0, 0.0], [0.0, H_img / float(Hf), 0.0], [0.0, 0.0, 1.0]])
            A = torch.tensor([[1.0 / res_x, 0.0, -min_x / res_x], [0.0, 1.0 / res_y, -min_y / res_y], [0.0, 0.0, 1.0]])
            M = A @ H_i2w @ S                                                    # :126-133
            out[b, v] = _warp_perspective(feats[b, v].unsqueeze(0), M, (Hb, Wb)).squeeze(0)
    return out
