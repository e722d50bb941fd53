"""Host-side mirror of the reference's projection / fusion modules, over the CUDA library.

Same class names, constructor arguments, forward signatures and error behaviour as
/root/reference/project/models/fusion/geometry.py and fusion.py, so BEVNet
(model_wrapper.py:42-43, :68-69) takes them unchanged:

    model.geom   = bevipm.GeometryTransformer(bev_h, bev_w, bounds, warp_impl)   # per-view maps
    model.fusion = bevipm.ConcatFusion()                                         # as wired today
or, fused (no [B,V,C,Hb,Wb] tensor ever exists):
    model.geom   = bevipm.FusedIPM(bev_h, bev_w, bounds, fusion="mean")
    model.fusion = torch.nn.Identity()

No parameters and no persistent buffers are added (state_dict keys unchanged, geometry.py:21).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.nn as nn

from . import _lib, ops


def _as_rt34(Rt: torch.Tensor) -> torch.Tensor:
    """geometry.py:41-59: 4x4 / 3x4 -> [R|t]; 3x3 -> [R|0]; anything else -> [I|0]."""
    if Rt.dim() == 2:
        if tuple(Rt.shape) == (4, 4):
            return Rt[:3, :]
        if tuple(Rt.shape) == (3, 4):
            return Rt
        if tuple(Rt.shape) == (3, 3):
            return torch.cat([Rt, Rt.new_zeros(3, 1)], dim=1)
    return torch.eye(3, 4, device=Rt.device if isinstance(Rt, torch.Tensor) else None)


def _as_k33(K: torch.Tensor) -> torch.Tensor:
    """geometry.py:35-40: top-left 3x3, or diag(1000, 1000, 1) when K is not at least 3x3."""
    if K.dim() != 2 or K.shape[0] < 3 or K.shape[1] < 3:
        k = torch.eye(3, device=K.device)
        k[0, 0] = k[1, 1] = 1000.0
        return k
    return K[:3, :3]


def _select(x, b: int, v: int, V: int, is_k: bool):
    """get_K / get_Rt of geometry.py:96-118."""
    if isinstance(x, torch.Tensor):
        if x.dim() == 4:
            return x[b, v]
        if x.dim() == 3:
            return x[v] if x.shape[0] == V else x[b]
        if x.dim() == 2:
            return x[:3, :3] if is_k else x[:4, :4]
        return torch.eye(3 if is_k else 4)
    return torch.as_tensor(x[b][v])


def pack_calibration(intrinsics, extrinsics, B: int, V: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
    """Every calibration form GeometryTransformer.forward accepts (geometry.py:96-118) ->
    dense float32 K [B,V,3,3] and Rt34 [B,V,3,4] on `device`."""
    def fast(x, tails):
        return isinstance(x, torch.Tensor) and x.dim() == 4 and tuple(x.shape[:2]) == (B, V) and tuple(x.shape[2:]) in tails

    if fast(intrinsics, {(3, 3)}):
        K = intrinsics
    else:
        K = torch.stack([torch.stack([_as_k33(_select(intrinsics, b, v, V, True)).to(device) for v in range(V)])
                         for b in range(B)])
    if fast(extrinsics, {(4, 4), (3, 4)}):
        Rt = extrinsics[..., :3, :]
    else:
        Rt = torch.stack([torch.stack([_as_rt34(_select(extrinsics, b, v, V, False)).to(device) for v in range(V)])
                          for b in range(B)])
    K = K.to(device=device, dtype=torch.float32).contiguous()
    Rt = Rt.to(device=device, dtype=torch.float32).contiguous()
    return K, Rt


_KORNIA_NOTE = False


def _resolve_kornia(warp_impl: str, emulate_kornia) -> bool:
    """The reference takes its kornia branch iff warp_impl == 'kornia' AND kornia imports (geometry.py:5-9, :124); BEVNet asks
    for 'kornia' (model_wrapper.py:42).  emulate_kornia=None mirrors that decision: the kornia-compatible sample positions are
    used exactly when `import kornia` works in this environment.  True / False force either geometry."""
    global _KORNIA_NOTE
    if warp_impl != "kornia":
        return False
    if emulate_kornia is None:
        try:
            import kornia  # noqa: F401
            emulate_kornia = True
        except Exception:
            emulate_kornia = False
    if emulate_kornia and not _KORNIA_NOTE:
        _KORNIA_NOTE = True
        import warnings
        warnings.warn("bevipm: warp_impl='kornia': using the kornia-compatible sample positions "
                      "(BEV pixel at the cell corner, source pixel p read at p*size/(size-1) - 0.5).  This mode restates kornia's "
                      "published warp_perspective and is NOT pinned against a kornia build; pass emulate_kornia=False for the "
                      "reference's grid_sample geometry.", stacklevel=3)
    return bool(emulate_kornia)


class _IPMBase(nn.Module):
    def __init__(self, bev_h: int, bev_w: int, bev_bounds: tuple):
        super().__init__()
        self.bev_h = bev_h
        self.bev_w = bev_w
        self.bounds = bev_bounds  # (x_min, x_max, y_min, y_max)
        self.res_x = (bev_bounds[1] - bev_bounds[0]) / bev_w
        self.res_y = (bev_bounds[3] - bev_bounds[2]) / bev_h
        self.register_buffer("ground_grid", self._create_ground_grid(), persistent=False)
        self._axes = {}
        # static cameras (wildtrack_loader.py:291-293): keep the calibration-derived tables of the fused kernel on the device
        # between calls -- the `_grid_cache` the reference declares and never fills (geometry.py:22).  Opt-in (subclasses take
        # `cache_tables=True`); results are identical either way, the kernel verifies the calibration on the device.
        self.cache_tables = False
        self._plans = {}

    def _create_ground_grid(self) -> torch.Tensor:
        # geometry.py:24-31 -- the linspace values themselves are part of the contract (rig.ground_axes)
        min_x, max_x, min_y, max_y = self.bounds
        xs = torch.linspace(min_x + 0.5 * self.res_x, max_x - 0.5 * self.res_x, self.bev_w)
        ys = torch.linspace(min_y + 0.5 * self.res_y, max_y - 0.5 * self.res_y, self.bev_h)
        yy, xx = torch.meshgrid(ys, xs, indexing="ij")
        return torch.stack([xx, yy, torch.ones_like(xx)], dim=-1)

    def _ground_axes(self, device):
        key = str(device)
        if key not in self._axes:
            g = self.ground_grid
            self._axes[key] = (g[0, :, 0].to(device=device, dtype=torch.float32).contiguous(),
                               g[:, 0, 1].to(device=device, dtype=torch.float32).contiguous())
        return self._axes[key]

    def _corner_axes(self, device):
        """BEV cell CORNERS x_min + j*res_x, y_min + i*res_y: where the reference's kornia branch puts BEV pixel (i, j)
        (geometry.py:129-131, A_w2bev maps world -> BEV pixel without a half-cell offset)."""
        key = "corner:" + str(device)
        if key not in self._axes:
            min_x, _, min_y, _ = self.bounds
            xs = (torch.arange(self.bev_w, dtype=torch.float64) * self.res_x + min_x).to(torch.float32)
            ys = (torch.arange(self.bev_h, dtype=torch.float64) * self.res_y + min_y).to(torch.float32)
            self._axes[key] = (xs.to(device).contiguous(), ys.to(device).contiguous())
        return self._axes[key]

    def _run(self, feats, intrinsics, extrinsics, img_size, mode: int, out_bf16: bool, variant: int, layout: str,
             kornia_geometry: bool = False):
        if feats.dim() != 5:
            raise ValueError("feats must be [B,V,C,Hf,Wf]")
        if not feats.is_cuda:
            raise RuntimeError("bevipm runs on CUDA tensors only: there is no CPU implementation of this path")
        B, V = feats.shape[:2]
        if feats.dtype == torch.float16:
            feats = feats.float()  # grid_sampler's autocast policy is fp32 (the reference under train.py:239)
        K, Rt = pack_calibration(intrinsics, extrinsics, B, V, feats.device)
        xs, ys = self._corner_axes(feats.device) if kornia_geometry else self._ground_axes(feats.device)
        if kornia_geometry:
            # geometry.py:134-136: a view whose feature->BEV matrix M is singular takes the grid_sample geometry instead
            sing = self._kornia_singular(K, Rt, feats.shape[-2:], img_size)
            if bool(sing.any()):   # (one small read-back, only in the kornia-compatible mode)
                return self._run_mixed(feats, K, Rt, sing, img_size, mode, out_bf16, variant, layout)
            mode = mode | (_lib.FLAG_KORNIA_GEOMETRY << 8)
        if layout == "channels_last" or (layout == "auto" and feats.shape[2] % (4 if feats.dtype == torch.float32 else 8) == 0):
            feats = ops.to_channels_last5(feats)
        H_img, W_img = img_size
        if self.cache_tables and variant == 0 and ops.planned_ok(feats):
            key = (str(feats.device), feats.dtype, tuple(feats.shape[1:]), tuple(feats.stride()[1:]), int(mode), int(H_img), int(W_img))
            plan = self._plans.get(key)
            if plan is None:
                plan = self._plans[key] = ops.new_plan(V, (self.bev_h, self.bev_w), feats.device)
            return ops.warp_fuse_planned(feats, K, Rt, xs, ys, int(H_img), int(W_img), mode, out_bf16, plan)
        return ops.warp_fuse(feats, K, Rt, xs, ys, int(H_img), int(W_img), mode, out_bf16, variant)

    def reset_table_cache(self) -> None:
        """Forget the cached tables (call after the cameras moved: a cache is bound to the first calibration it sees; until
        then frames with another calibration are still computed correctly, just without the cache)."""
        for plan in self._plans.values():
            plan.zero_()


    def _kornia_singular(self, K, Rt34, feat_hw, img_size) -> torch.Tensor:
        """bool [B,V]: |det(M)| < 1e-8 or not finite for M = A_w2bev @ H_img2world @ S_feat2img (geometry.py:125-134)."""
        Hf, Wf = feat_hw
        H_img, W_img = img_size
        G = torch.stack([Rt34[..., 0], Rt34[..., 1], Rt34[..., 3]], dim=-1)
        H = K @ G
        det = torch.linalg.det(H)
        bad = ~torch.isfinite(det) | (det.abs() < 1e-8)
        Hinv = torch.where(bad[..., None, None], torch.linalg.pinv(H), torch.linalg.inv(torch.where(bad[..., None, None], torch.eye(3, device=H.device), H)))
        S = torch.tensor([[W_img / float(Wf), 0.0, 0.0], [0.0, H_img / float(Hf), 0.0], [0.0, 0.0, 1.0]], device=H.device)
        min_x, _, min_y, _ = self.bounds
        A = torch.tensor([[1.0 / self.res_x, 0.0, -min_x / self.res_x], [0.0, 1.0 / self.res_y, -min_y / self.res_y], [0.0, 0.0, 1.0]],
                         device=H.device)
        detM = torch.linalg.det(A @ Hinv @ S)
        return ~torch.isfinite(detM) | (detM.abs() < 1e-8)

    def _run_mixed(self, feats, K, Rt, sing, img_size, mode, out_bf16, variant, layout):
        """Kornia-compatible geometry for the regular views, grid_sample geometry for the singular ones (rare): two per-view
        launches, a select, and the sequential view reduction of fusion.py:17-22 on our kernel."""
        if layout == "channels_last" or (layout == "auto" and feats.shape[2] % (4 if feats.dtype == torch.float32 else 8) == 0):
            feats = ops.to_channels_last5(feats)
        H_img, W_img = int(img_size[0]), int(img_size[1])
        xs_k, ys_k = self._corner_axes(feats.device)
        xs_g, ys_g = self._ground_axes(feats.device)
        pv_k = ops.warp_fuse(feats, K, Rt, xs_k, ys_k, H_img, W_img, _lib.NONE | (_lib.FLAG_KORNIA_GEOMETRY << 8), False, 0)
        pv_g = ops.warp_fuse(feats, K, Rt, xs_g, ys_g, H_img, W_img, _lib.NONE, False, 0)
        pv = torch.where(sing[:, :, None, None, None], pv_g, pv_k)
        if mode == _lib.NONE:
            return pv
        name = {v: k for k, v in _lib.MODES.items() if k != "concat"}[mode]
        out = ops.fuse_views_autograd(pv, name) if pv.requires_grad else ops.fuse_views(pv, name)
        return out.bfloat16() if out_bf16 else out


class GeometryTransformer(_IPMBase):
    """Per-view IPM warp, drop-in for geometry.py:12-163 (grid_sample semantics)."""

    def __init__(self, bev_h: int, bev_w: int, bev_bounds: tuple, warp_impl: str = "grid_sample", layout: str = "keep",
                 emulate_kornia: bool | None = None, cache_tables: bool = False):
        super().__init__(bev_h, bev_w, bev_bounds)
        self.cache_tables = cache_tables
        # geometry.py:20 -- both names are accepted; 'kornia' executes the grid_sample branch in any
        # environment without kornia (this image) and the kornia branch where kornia imports: emulate_kornia=None
        # (default) mirrors exactly that decision (_resolve_kornia).  The kornia-compatible mode reproduces the sample
        # positions of geometry.py:124-141: BEV pixel j at the cell corner, source pixel p read at
        # p*size/(size-1) - 0.5; a view whose matrix M is singular (|det| < 1e-8, geometry.py:134-136) falls back
        # to the grid_sample geometry like the reference.  From kornia's published algorithm: parity unpinned.
        self.warp_impl = warp_impl if warp_impl in ("grid_sample", "kornia") else "grid_sample"
        self.emulate_kornia = _resolve_kornia(self.warp_impl, emulate_kornia)
        self.layout = layout
        self._grid_cache = {}

    @staticmethod
    def _compute_homography(K: torch.Tensor, Rt: torch.Tensor) -> torch.Tensor:
        """H = K [r1 r2 t] (geometry.py:33-64); tiny host-side helper, called by BEVNet's unused loss."""
        K = _as_k33(K)
        Rt34 = _as_rt34(Rt).to(K.device)
        G = torch.cat([Rt34[:, 0:1], Rt34[:, 1:2], Rt34[:, 3:4]], dim=1)
        return K @ G

    @staticmethod
    def _compute_img_to_world_homography(K: torch.Tensor, Rt: torch.Tensor) -> torch.Tensor:
        """inverse with pinv fallback (geometry.py:66-78)."""
        H = GeometryTransformer._compute_homography(K, Rt)
        try:
            det = torch.det(H)
        except Exception:
            det = torch.tensor(float("nan"), device=H.device)
        if torch.isnan(det) or torch.isinf(det) or det.abs().item() < 1e-8:
            return torch.linalg.pinv(H)
        try:
            return torch.linalg.inv(H)
        except Exception:
            return torch.linalg.pinv(H)

    def forward(self, feats: torch.Tensor, intrinsics, extrinsics,
                img_size: Tuple[int, int] = (1080, 1920)) -> torch.Tensor:
        """feats [B,V,C,Hf,Wf] -> [B,V,C,Hb,Wb] float32 (geometry.py:94: fp32 whatever the input)."""
        return self._run(feats, intrinsics, extrinsics, img_size, _lib.NONE, False, 0, self.layout, self.emulate_kornia)


class FusedIPM(_IPMBase):
    """forward(features, calibration) -> BEV [B,C,Hb,Wb]: GeometryTransformer + SimpleFusion in one
    kernel launch (geometry.py:80-163 followed by fusion.py:17-22), or ConcatFusion's
    [B,V*C,Hb,Wb] for fusion='concat'."""

    def __init__(self, bev_h: int, bev_w: int, bev_bounds: tuple, fusion: str = "mean",
                 warp_impl: str = "grid_sample", out_dtype: torch.dtype = torch.float32,
                 layout: str = "auto", variant: int = 0, emulate_kornia: bool | None = None, return_valid: bool = False,
                 cache_tables: bool = False):
        super().__init__(bev_h, bev_w, bev_bounds)
        self.cache_tables = cache_tables   # static cameras: keep the calibration-derived tables on the device between calls
        # "mean_valid": mean over the views that SEE a cell (opt-in extension; the reference's "mean" divides by V)
        assert fusion in ("sum", "mean", "max", "concat", "none", "mean_valid")
        assert out_dtype in (torch.float32, torch.bfloat16)
        assert layout in ("auto", "keep", "channels_last")
        self.fusion = fusion
        self.warp_impl = warp_impl if warp_impl in ("grid_sample", "kornia") else "grid_sample"
        self.emulate_kornia = _resolve_kornia(self.warp_impl, emulate_kornia)   # see GeometryTransformer
        self.out_dtype = out_dtype
        self.layout = layout   # "auto": NCHW-contiguous features go through our transpose pre-pass
        self.variant = variant
        self.return_valid = return_valid   # also return count [B,Hb,Wb] int32: views that see each cell

    def valid_count(self, feats: torch.Tensor, intrinsics, extrinsics, img_size: Tuple[int, int] = (1080, 1920)) -> torch.Tensor:
        """count [B,Hb,Wb] int32 of the views whose sample position has a bilinear tap inside the feature map."""
        B, V = feats.shape[:2]
        K, Rt = pack_calibration(intrinsics, extrinsics, B, V, feats.device)
        xs, ys = self._corner_axes(feats.device) if self.emulate_kornia else self._ground_axes(feats.device)
        return ops.valid_count(K, Rt, xs, ys, tuple(feats.shape[-2:]), (int(img_size[0]), int(img_size[1])),
                               _lib.FLAG_KORNIA_GEOMETRY if self.emulate_kornia else 0)

    def forward(self, feats: torch.Tensor, intrinsics, extrinsics,
                img_size: Tuple[int, int] = (1080, 1920)):
        fusion = self.fusion
        needs_grad = feats.requires_grad and torch.is_grad_enabled()
        if fusion == "max" and needs_grad:
            # the fused max kernel has no backward: per-view maps (our forward + backward) followed by our max reduction,
            # whose backward routes the gradient to the first arg-max view like torch.max (fusion.py:22 is differentiable)
            pv = self._run(feats, intrinsics, extrinsics, img_size, _lib.NONE, False, 0, self.layout, self.emulate_kornia)
            out = ops.fuse_views_autograd(pv, "max")
            out = out.to(self.out_dtype)
        elif fusion == "mean_valid":
            out = self._run(feats, intrinsics, extrinsics, img_size, _lib.SUM, False, self.variant, self.layout, self.emulate_kornia)
            cnt = self.valid_count(feats, intrinsics, extrinsics, img_size)
            if needs_grad:
                out = out / cnt.clamp(min=1).to(out.dtype).unsqueeze(1)
            else:
                out = ops.divide_by_count_(out, cnt)
            out = out.to(self.out_dtype)
            return (out, cnt) if self.return_valid else out
        else:
            out = self._run(feats, intrinsics, extrinsics, img_size, _lib.MODES[fusion],
                            self.out_dtype == torch.bfloat16, self.variant, self.layout, self.emulate_kornia)
        if fusion == "concat":
            B, V, C, Hb, Wb = out.shape
            out = out.reshape(B, V * C, Hb, Wb)   # fusion.py:45-46, channel index v*C + c
        if self.return_valid:
            return out, self.valid_count(feats, intrinsics, extrinsics, img_size)
        return out


class FoldedConcatProjIPM(_IPMBase):
    """GeometryTransformer -> ConcatFusion -> 1x1 projection of BEVNet (model_wrapper.py:68-73) with the
    projection folded in FRONT of the warp.

    The warp is linear per channel, so  proj(concat_v(warp_v(f_v))) = sum_v warp_v(W_v f_v) + bias  with
    W_v = proj.weight[:, v*C:(v+1)*C]: the per-view 1x1 convolution runs on the small source maps (one launch of the
    hand-written tcgen05 GEMM, csrc/bevipm_proj.cu), the fused SUM kernel then gathers `out_channels` instead of V*C
    channels per cell and the [B,V*C,Hb,Wb] tensor never exists (7 x 1280 x 120 x 360 fp32 = 1.55 GB per frame at
    wildtrack.yaml).  A reassociation of the reference's arithmetic: equal within fp32 rounding (tests: 1e-4 relative), not
    bit-exact.

    precision: "fp32" (default) = split-operand TF32 (three MMAs per step, ~1e-6 relative: comparable with the fp32 Conv2d the
    reference runs on the CPU), "tf32" = one TF32 pass (what cuDNN runs for the reference's Conv2d on a GPU under torch's
    default `torch.backends.cudnn.allow_tf32 = True`, ~1e-3), "auto" = "tf32" iff that torch switch is on.
    gemm: "auto" (default) = our tcgen05 kernel whenever it takes the shape (out_channels a multiple of 16 up to 256, C a
    multiple of 4), torch.einsum (cuBLAS) otherwise; "tcgen05" = our kernel or an error; "cublas" = the round-1 path, kept for
    comparison.  `last_gemm` names the one the last forward used.

    `proj` is the nn.Conv2d(V*C, out_channels, 1) the reference builds lazily (model_wrapper.py:70-72); it stays
    the owner of weight and bias, so checkpoints load unchanged and gradients reach it through the GEMM.
    """

    def __init__(self, bev_h: int, bev_w: int, bev_bounds: tuple, proj: nn.Conv2d, views: int, variant: int = 0,
                 precision: str = "fp32", gemm: str = "auto"):
        super().__init__(bev_h, bev_w, bev_bounds)
        assert proj.kernel_size == (1, 1) and proj.in_channels % views == 0
        assert precision in ("fp32", "tf32", "auto") and gemm in ("auto", "tcgen05", "cublas")
        self.proj = proj
        self.views = views
        self.variant = variant
        self.precision = precision
        self.gemm = gemm
        self.last_gemm = None

    def _passes(self) -> int:
        if self.precision == "auto":
            return 1 if torch.backends.cudnn.allow_tf32 else 3
        return 1 if self.precision == "tf32" else 3

    def forward(self, feats: torch.Tensor, intrinsics, extrinsics,
                img_size: Tuple[int, int] = (1080, 1920)) -> torch.Tensor:
        B, V, C, Hf, Wf = feats.shape
        assert V == self.views and V * C == self.proj.in_channels
        Co = self.proj.out_channels
        W = self.proj.weight.view(Co, V, C).to(torch.float32)                      # [Co,V,C]
        x = feats.to(torch.float32).permute(0, 1, 3, 4, 2)                        # [B,V,Hf,Wf,C] (view, no copy if channels-last)
        use_ours = self.gemm == "tcgen05" or (self.gemm == "auto" and ops.proj1x1_supported(C, Co))
        self.last_gemm = "tcgen05" if use_ours else "cublas"
        if not use_ours:
            g = torch.einsum("bvhwc,ovc->bvhwo", x, W)                             # per-view 1x1 conv, channels-last result
        else:
            x = x.contiguous()                                                     # NCHW input: one transposing copy
            g = ops.proj1x1(x.view(B * V, Hf * Wf, C), W, self._passes()).view(B, V, Hf, Wf, Co)
        g = g.permute(0, 1, 4, 2, 3)                                               # logical [B,V,Co,Hf,Wf], NHWC in memory
        out = self._run(g, intrinsics, extrinsics, img_size, _lib.SUM, False, self.variant, "keep")
        if self.proj.bias is not None:
            out = out + self.proj.bias.view(1, Co, 1, 1)
        return out


class FusionModule(nn.Module):
    def forward(self, bev_maps: torch.Tensor) -> torch.Tensor:
        """bev_maps: Tensor[B, V, C, H, W] -> Tensor[B, C, H, W]   (fusion.py:5-8)"""
        raise NotImplementedError


class SimpleFusion(FusionModule):
    """fusion.py:11-22 on materialised per-view maps (our reduction kernels, forward and backward; mean divides by V,
    max sends the gradient to the first view holding the maximum like torch.max)."""

    def __init__(self, mode: str = "sum"):
        super().__init__()
        assert mode in ("sum", "mean", "max")
        self.mode = mode

    def forward(self, bev_maps: torch.Tensor) -> torch.Tensor:
        if bev_maps.requires_grad and torch.is_grad_enabled():
            return ops.fuse_views_autograd(bev_maps, self.mode)   # forward and backward on our kernels
        return ops.fuse_views(bev_maps, self.mode)


class AttentionFusion(FusionModule):
    """fusion.py:25-36: the reference's placeholder (announces itself once, returns the mean)."""

    def __init__(self):
        super().__init__()
        self._warned = False
        self._mean = SimpleFusion("mean")

    def forward(self, bev_maps: torch.Tensor) -> torch.Tensor:
        if not self._warned:
            print("[AttentionFusion] Placeholder only. Not implemented.")
            self._warned = True
        return self._mean(bev_maps)


class ConcatFusion(FusionModule):
    """fusion.py:39-46: [B,V,C,H,W] -> [B,V*C,H,W], view-major channels."""

    def forward(self, bev_maps: torch.Tensor) -> torch.Tensor:
        B, V, C, H, W = bev_maps.shape
        return bev_maps.reshape(B, V * C, H, W)


class DeformAttnFusion(FusionModule):
    """Phase-2 BEV fusion (the slot of the reference's AttentionFusion placeholder, fusion.py:25-36):
    every BEV cell is a query that samples `points` locations per head from EACH view's BEV-warped map
    around its own position (MVDeTr-style deformable attention, one level per camera view).

    forward(bev_maps [B,V,C,H,W]) -> [B,C,H,W].  The projections are ordinary nn.Linear layers; the
    sampling itself (B*H*W*heads*V*points bilinear gathers) runs in `bevipm_deform_attn_fwd`.
    Trainable end to end: the sampling's backward (d value, d locations, d weights) runs in `bevipm_deform_attn_bwd`.
    """

    def __init__(self, channels: int, views: int, heads: int = 8, points: int = 4):
        super().__init__()
        assert channels % heads == 0
        self.channels, self.views, self.heads, self.points = channels, views, heads, points
        self.sampling_offsets = nn.Linear(channels, heads * views * points * 2)
        self.attention_weights = nn.Linear(channels, heads * views * points)
        self.value_proj = nn.Linear(channels, channels)
        self.output_proj = nn.Linear(channels, channels)
        nn.init.zeros_(self.sampling_offsets.weight)
        # Deformable-DETR's initialisation: points fan out on a ring around the query
        import math
        th = torch.arange(heads, dtype=torch.float32) * (2.0 * math.pi / heads)
        grid = torch.stack([th.cos(), th.sin()], -1)
        grid = (grid / grid.abs().max(-1, keepdim=True)[0]).view(heads, 1, 1, 2).repeat(1, views, points, 1)
        for i in range(points):
            grid[:, :, i, :] *= i + 1
        with torch.no_grad():
            self.sampling_offsets.bias.copy_(grid.reshape(-1))
        nn.init.zeros_(self.attention_weights.weight)
        nn.init.zeros_(self.attention_weights.bias)

    def forward(self, bev_maps: torch.Tensor) -> torch.Tensor:
        B, V, C, H, W = bev_maps.shape
        assert V == self.views and C == self.channels
        M, P, D = self.heads, self.points, C // self.heads
        maps = bev_maps.permute(0, 1, 3, 4, 2)                                 # [B,V,H,W,C]
        query = maps.mean(dim=1).reshape(B, H * W, C)                          # mean-fused BEV as the query
        value = self.value_proj(maps.reshape(B, V * H * W, C)).view(B, V * H * W, M, D)
        off = self.sampling_offsets(query).view(B, H * W, M, V, P, 2)
        aw = torch.softmax(self.attention_weights(query).view(B, H * W, M, V * P), -1).view(B, H * W, M, V, P)
        ys, xs = torch.meshgrid(torch.arange(H, device=maps.device), torch.arange(W, device=maps.device), indexing="ij")
        ref = torch.stack([(xs + 0.5) / W, (ys + 0.5) / H], -1).reshape(1, H * W, 1, 1, 1, 2).to(off.dtype)
        loc = ref + off / torch.tensor([W, H], device=maps.device, dtype=off.dtype)
        shapes = torch.tensor([[H, W]] * V, dtype=torch.int32, device=maps.device)
        start = torch.arange(V, device=maps.device, dtype=torch.int64) * (H * W)
        if value.dtype not in (torch.float32, torch.bfloat16):
            value = value.float()
        out = ops.deform_attn(value, shapes, start, loc, aw, out_dtype=value.dtype)  # [B,Q,C]
        out = self.output_proj(out.to(query.dtype))
        return out.view(B, H, W, C).permute(0, 3, 1, 2)


class ImageSpaceDeformAttnFusion(_IPMBase):
    """Phase-2 fusion WITHOUT the per-view BEV maps: every BEV cell is a query that samples the SOURCE feature maps of all
    views around the cell's own inverse-perspective position (the fused, image-space form of DeformAttnFusion).

    forward(feats [B,V,C,Hf,Wf], intrinsics, extrinsics, img_size) -> [B,C,Hb,Wb] -- GeometryTransformer's signature, so it
    replaces `geom` + `fusion` of BEVNet together (model_wrapper.py:68-69) like FusedIPM does.
      reference point of cell q in view v = the position the reference's warp samples (geometry.py:142-160; our
        `bevipm_sample_coords`), as a normalised location of the deformable-attention kernel;
      query = the mean-fused BEV of the fused warp kernel; learned offsets are in SOURCE texels (per head / view / point);
      value = value_proj(features) on the source maps (V*Hf*Wf texels: never warped, never materialised on the BEV grid);
      weights: softmax over views x points per head; `mask_invalid=True` takes the softmax over the views that see the cell
        only (the attention counterpart of fusion="mean_valid"); the default keeps every view, a view that misses the cell
        contributes its zero padding like in the reference's mean (fusion.py:20-21).
    With zero offsets, uniform weights and identity projections the result IS the reference's GeometryTransformer ->
    SimpleFusion('mean') (tests/test_deform_attn.py pins the module to the fused kernel that way).  Forward and backward run on
    our kernels (bevipm_warp_fuse_*, bevipm_sample_coords, bevipm_deform_attn_*); the projections are nn.Linear.
    """

    def __init__(self, bev_h: int, bev_w: int, bev_bounds: tuple, channels: int, views: int, heads: int = 8, points: int = 4,
                 mask_invalid: bool = False):
        super().__init__(bev_h, bev_w, bev_bounds)
        assert channels % heads == 0
        self.channels, self.views, self.heads, self.points = channels, views, heads, points
        self.mask_invalid = mask_invalid
        self.sampling_offsets = nn.Linear(channels, heads * views * points * 2)
        self.attention_weights = nn.Linear(channels, heads * views * points)
        self.value_proj = nn.Linear(channels, channels)
        self.output_proj = nn.Linear(channels, channels)
        nn.init.zeros_(self.sampling_offsets.weight)
        import math
        th = torch.arange(heads, dtype=torch.float32) * (2.0 * math.pi / heads)
        grid = torch.stack([th.cos(), th.sin()], -1)
        grid = (grid / grid.abs().max(-1, keepdim=True)[0]).view(heads, 1, 1, 2).repeat(1, views, points, 1)
        for i in range(points):
            grid[:, :, i, :] *= i + 1
        with torch.no_grad():
            self.sampling_offsets.bias.copy_(grid.reshape(-1))
        nn.init.zeros_(self.attention_weights.weight)
        nn.init.zeros_(self.attention_weights.bias)

    def reference_points(self, K: torch.Tensor, Rt: torch.Tensor, feat_hw, img_size):
        """loc [B,Q,V,2] (x, y in [0,1] of each view's map; far outside where the position is not finite) and seen [B,Q,V]."""
        Hf, Wf = feat_hw
        xs, ys = self._ground_axes(K.device)
        ix, iy = ops.sample_coords(K, Rt, xs, ys, (Hf, Wf), img_size)                       # [B,V,Hb,Wb], source texel units
        ok = torch.isfinite(ix) & torch.isfinite(iy)
        seen = ok & (ix > -1) & (ix < Wf) & (iy > -1) & (iy < Hf)                           # some tap of the cell is in the map
        lx = torch.where(ok, (ix + 0.5) / Wf, torch.full_like(ix, -1e3))
        ly = torch.where(ok, (iy + 0.5) / Hf, torch.full_like(iy, -1e3))
        B, V = ix.shape[:2]
        loc = torch.stack([lx, ly], -1).view(B, V, -1, 2).permute(0, 2, 1, 3)                # [B,Q,V,2]
        return loc, seen.view(B, V, -1).permute(0, 2, 1)

    def forward(self, feats: torch.Tensor, intrinsics, extrinsics, img_size: Tuple[int, int] = (1080, 1920)) -> torch.Tensor:
        if feats.dim() != 5:
            raise ValueError("feats must be [B,V,C,Hf,Wf]")
        if not feats.is_cuda:
            raise RuntimeError("bevipm runs on CUDA tensors only: there is no CPU implementation of this path")
        B, V, C, Hf, Wf = feats.shape
        assert V == self.views and C == self.channels
        M, P, D = self.heads, self.points, C // self.heads
        Hb, Wb = self.bev_h, self.bev_w
        Q = Hb * Wb
        if feats.dtype == torch.float16:
            feats = feats.float()
        feats = ops.to_channels_last5(feats)                                               # [B,V,Hf,Wf,C] in memory
        K, Rt = pack_calibration(intrinsics, extrinsics, B, V, feats.device)
        bev = self._run(feats, K, Rt, img_size, _lib.MEAN, False, 0, "keep")               # fused warp + mean: the query
        query = bev.permute(0, 2, 3, 1).reshape(B, Q, C).to(self.value_proj.weight.dtype)
        with torch.no_grad():
            ref, seen = self.reference_points(K, Rt, (Hf, Wf), img_size)
        src = feats.permute(0, 1, 3, 4, 2).reshape(B, V * Hf * Wf, C)
        value = self.value_proj(src.to(self.value_proj.weight.dtype)).view(B, V * Hf * Wf, M, D)
        off = self.sampling_offsets(query).view(B, Q, M, V, P, 2)
        loc = ref.view(B, Q, 1, V, 1, 2).to(off.dtype) + off / torch.tensor([Wf, Hf], device=off.device, dtype=off.dtype)
        logits = self.attention_weights(query).view(B, Q, M, V, P)
        if self.mask_invalid:
            logits = logits.masked_fill(~seen.view(B, Q, 1, V, 1), float("-inf"))
            aw = torch.softmax(logits.reshape(B, Q, M, V * P), -1)
            aw = torch.nan_to_num(aw, nan=0.0).view(B, Q, M, V, P)                          # cells no view sees: all weights 0
        else:
            aw = torch.softmax(logits.reshape(B, Q, M, V * P), -1).view(B, Q, M, V, P)
        shapes = torch.tensor([[Hf, Wf]] * V, dtype=torch.int32, device=feats.device)
        start = torch.arange(V, device=feats.device, dtype=torch.int64) * (Hf * Wf)
        if value.dtype not in (torch.float32, torch.bfloat16):
            value = value.float()
        out = ops.deform_attn(value, shapes, start, loc, aw, out_dtype=value.dtype)         # [B,Q,C]
        out = self.output_proj(out.to(query.dtype))
        return out.view(B, Hb, Wb, C).permute(0, 3, 1, 2)
