"""`bevipm::warp_fuse` -- thin PyTorch custom op over the C ABI (include/bevipm.h).

torch supplies device memory, the current stream and autograd bookkeeping; every FLOP and byte
of the path runs in libbevipm.so.  CUDA tensors only: the op is registered for device type
"cuda" and nothing else, so CPU tensors raise instead of silently taking another route.
"""
from __future__ import annotations

import ctypes
from typing import Sequence, Tuple

import torch

from . import _lib
from ._lib import Desc, MODES

_DT = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}


def _stream_ptr(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _ptr(t: torch.Tensor) -> ctypes.c_void_p:
    return ctypes.c_void_p(t.data_ptr())


def _fill_desc(feat_shape, feat_strides, out_strides, bev_hw, img_hw, mode, in_dt, out_dt, variant, flags=0) -> Desc:
    B, V, C, Hf, Wf = feat_shape
    d = Desc()
    d.B, d.V, d.C, d.Hf, d.Wf = B, V, C, Hf, Wf
    d.Hb, d.Wb = bev_hw
    d.img_h, d.img_w = img_hw
    d.mode, d.in_dtype, d.out_dtype, d.variant = mode, in_dt, out_dt, variant
    d.flags = flags
    d.fs_b, d.fs_v, d.fs_c, d.fs_y, d.fs_x = feat_strides
    d.os_b, d.os_v, d.os_c, d.os_y, d.os_x = out_strides
    return d


def _out_strides5(out: torch.Tensor, per_view: bool):
    s = out.stride()
    return tuple(s) if per_view else (s[0], 0, s[1], s[2], s[3])


def _is_channels_last5(t: torch.Tensor) -> bool:
    return t.stride(2) == 1 or t.shape[2] == 1


def _alloc_out(feats: torch.Tensor, Hb: int, Wb: int, per_view: bool, dtype, channels_last: bool):
    B, V, C = feats.shape[:3]
    if channels_last:
        if per_view:
            return torch.empty((B, V, Hb, Wb, C), device=feats.device, dtype=dtype).permute(0, 1, 4, 2, 3)
        return torch.empty((B, Hb, Wb, C), device=feats.device, dtype=dtype).permute(0, 3, 1, 2)
    shape = (B, V, C, Hb, Wb) if per_view else (B, C, Hb, Wb)
    return torch.empty(shape, device=feats.device, dtype=dtype)


def _check_calib(feats, K, Rt34, xs, ys):
    B, V = feats.shape[:2]
    for name, t, shape in (("K", K, (B, V, 3, 3)), ("Rt34", Rt34, (B, V, 3, 4))):
        if tuple(t.shape) != shape or t.dtype != torch.float32 or not t.is_contiguous() or t.device != feats.device:
            raise ValueError(f"{name} must be a contiguous float32 {shape} tensor on {feats.device}")
    for name, t in (("xs", xs), ("ys", ys)):
        if t.dim() != 1 or t.dtype != torch.float32 or not t.is_contiguous() or t.device != feats.device:
            raise ValueError(f"{name} must be a contiguous 1-D float32 tensor on {feats.device}")


@torch.library.custom_op("bevipm::warp_fuse", mutates_args=(), device_types="cuda")
def warp_fuse(feats: torch.Tensor, K: torch.Tensor, Rt34: torch.Tensor, xs: torch.Tensor, ys: torch.Tensor,
              img_h: int, img_w: int, mode: int, out_bf16: bool, variant: int) -> torch.Tensor:
    """Fused IPM warp + view fusion (geometry.py:120-162 + fusion.py:17-22 of the reference).

    feats [B,V,C,Hf,Wf] float32 / bfloat16, any strides (channels-last = the fast kernel).
    Returns [B,C,Hb,Wb] (sum / mean / max) or [B,V,C,Hb,Wb] (mode NONE), channels-last in memory
    when the features are, fp32 unless out_bf16.
    """
    if feats.dim() != 5:
        raise ValueError("feats must be [B,V,C,Hf,Wf]")
    if feats.dtype not in _DT:
        raise TypeError(f"feats dtype {feats.dtype} is not supported (float32 / bfloat16)")
    _check_calib(feats, K, Rt34, xs, ys)
    L = _lib.load()
    mode, flags = mode & 0xff, mode >> 8   # bevipm_desc.flags ride in the high bits of `mode` (see modules._IPMBase._run)
    per_view = mode == _lib.NONE
    Hb, Wb = ys.numel(), xs.numel()
    out_dtype = torch.bfloat16 if out_bf16 else torch.float32
    with torch.cuda.device(feats.device):
        out = _alloc_out(feats, Hb, Wb, per_view, out_dtype, _is_channels_last5(feats))
        d = _fill_desc(feats.shape, feats.stride(), _out_strides5(out, per_view), (Hb, Wb), (img_h, img_w), mode,
                       _DT[feats.dtype], _DT[out_dtype], variant, flags)
        _lib.check(L.bevipm_warp_fuse_fwd(ctypes.byref(d), _ptr(feats), _ptr(K), _ptr(Rt34), _ptr(xs), _ptr(ys),
                                          _ptr(out), ctypes.c_void_p(_stream_ptr(feats.device))))
    return out


@warp_fuse.register_fake
def _(feats, K, Rt34, xs, ys, img_h, img_w, mode, out_bf16, variant):
    B, V, C = feats.shape[:3]
    Hb, Wb = ys.numel(), xs.numel()
    dt = torch.bfloat16 if out_bf16 else torch.float32
    # same strides as the real op: channels-last in memory when the features are (_alloc_out)
    per_view = (mode & 0xff) == _lib.NONE
    if _is_channels_last5(feats):
        if per_view:
            return feats.new_empty((B, V, Hb, Wb, C), dtype=dt).permute(0, 1, 4, 2, 3)
        return feats.new_empty((B, Hb, Wb, C), dtype=dt).permute(0, 3, 1, 2)
    return feats.new_empty((B, V, C, Hb, Wb) if per_view else (B, C, Hb, Wb), dtype=dt)


@torch.library.custom_op("bevipm::warp_fuse_bwd", mutates_args=(), device_types="cuda")
def warp_fuse_bwd(grad_out: torch.Tensor, K: torch.Tensor, Rt34: torch.Tensor, xs: torch.Tensor, ys: torch.Tensor,
                  feat_shape: Sequence[int], channels_last: bool, img_h: int, img_w: int, mode: int) -> torch.Tensor:
    """fp32 gradient w.r.t. the features (scatter of the forward taps; train.py:243 reaches it)."""
    L = _lib.load()
    B, V, C, Hf, Wf = feat_shape
    mode, flags = mode & 0xff, mode >> 8
    per_view = mode == _lib.NONE
    if grad_out.dtype not in _DT:
        grad_out = grad_out.float()
    if channels_last and grad_out.stride(-3 if not per_view else 2) != 1:
        # channels-last features want a channels-last dL/dBEV (the vectorised kernels); one transpose of the BEV-sized
        # gradient is cheaper than the strided scatter
        perm = (0, 1, 3, 4, 2) if per_view else (0, 2, 3, 1)
        inv = (0, 1, 4, 2, 3) if per_view else (0, 3, 1, 2)
        grad_out = grad_out.permute(*perm).contiguous().permute(*inv)
    with torch.cuda.device(grad_out.device):
        if channels_last:
            g = torch.zeros((B, V, Hf, Wf, C), device=grad_out.device, dtype=torch.float32).permute(0, 1, 4, 2, 3)
        else:
            g = torch.zeros((B, V, C, Hf, Wf), device=grad_out.device, dtype=torch.float32)
        d = _fill_desc(feat_shape, g.stride(), _out_strides5(grad_out, per_view), (ys.numel(), xs.numel()),
                       (img_h, img_w), mode, _lib.F32, _DT[grad_out.dtype], 0, flags)
        _lib.check(L.bevipm_warp_fuse_bwd(ctypes.byref(d), _ptr(grad_out), _ptr(K), _ptr(Rt34), _ptr(xs), _ptr(ys),
                                          _ptr(g), ctypes.c_void_p(_stream_ptr(grad_out.device))))
    return g


@warp_fuse_bwd.register_fake
def _(grad_out, K, Rt34, xs, ys, feat_shape, channels_last, img_h, img_w, mode):
    B, V, C, Hf, Wf = feat_shape
    if channels_last:
        return grad_out.new_empty((B, V, Hf, Wf, C), dtype=torch.float32).permute(0, 1, 4, 2, 3)
    return grad_out.new_empty((B, V, C, Hf, Wf), dtype=torch.float32)


def _setup_ctx(ctx, inputs, output):
    feats, K, Rt34, xs, ys, img_h, img_w, mode, out_bf16, variant = inputs
    ctx.save_for_backward(K, Rt34, xs, ys)
    ctx.feat_shape = tuple(feats.shape)
    ctx.feat_dtype = feats.dtype
    ctx.channels_last = _is_channels_last5(feats)
    ctx.args = (img_h, img_w, mode)


def _backward(ctx, grad_out):
    K, Rt34, xs, ys = ctx.saved_tensors
    img_h, img_w, mode = ctx.args
    g = warp_fuse_bwd(grad_out, K, Rt34, xs, ys, list(ctx.feat_shape), ctx.channels_last, img_h, img_w, mode)
    return (g.to(ctx.feat_dtype),) + (None,) * 9


warp_fuse.register_autograd(_backward, setup_context=_setup_ctx)


# ---- the same op over a table cache (static cameras: SURVEY.md 8(f) N4, the reference's unused `_grid_cache`) --------

def plan_bytes(V: int, bev_hw: Tuple[int, int]) -> int:
    """Bytes of a table cache for V views and a [Hb, Wb] grid (include/bevipm.h: bevipm_plan_bytes)."""
    d = _fill_desc((1, V, 8, 2, 2), (0,) * 5, (0,) * 5, bev_hw, (2, 2), 0, 0, 0, 0)
    n = int(_lib.load().bevipm_plan_bytes(ctypes.byref(d)))
    if n < 0:
        raise RuntimeError(_lib.load().bevipm_last_error().decode())
    return n


def new_plan(V: int, bev_hw: Tuple[int, int], device) -> torch.Tensor:
    """An empty (zeroed) table cache; `plan.zero_()` re-arms it for another calibration."""
    return torch.zeros(plan_bytes(V, bev_hw), dtype=torch.uint8, device=device)


# (`plan` is declared read-only although the first call fills it: a cache whose content never changes the result is not a
#  mutation the dispatcher has to order anything around, and a functional schema is what autograd registration needs)
@torch.library.custom_op("bevipm::warp_fuse_planned", mutates_args=(), device_types="cuda")
def warp_fuse_planned(feats: torch.Tensor, K: torch.Tensor, Rt34: torch.Tensor, xs: torch.Tensor, ys: torch.Tensor,
                      img_h: int, img_w: int, mode: int, out_bf16: bool, plan: torch.Tensor) -> torch.Tensor:
    """warp_fuse for launches that take the default run kernel (channels-last features, V <= 32), with the calibration-
    derived tables of every row segment kept in `plan` between calls: the first call fills an empty cache from its frame 0,
    later frames with the same calibration (compared on the device, bit for bit) copy their tables instead of computing
    them.  The result equals warp_fuse's bit for bit in every case."""
    if feats.dim() != 5:
        raise ValueError("feats must be [B,V,C,Hf,Wf]")
    if feats.dtype not in _DT:
        raise TypeError(f"feats dtype {feats.dtype} is not supported (float32 / bfloat16)")
    _check_calib(feats, K, Rt34, xs, ys)
    if plan.dtype != torch.uint8 or not plan.is_contiguous() or plan.device != feats.device:
        raise ValueError("plan must be a contiguous uint8 tensor on the features' device (ops.new_plan)")
    L = _lib.load()
    mode, flags = mode & 0xff, mode >> 8
    per_view = mode == _lib.NONE
    Hb, Wb = ys.numel(), xs.numel()
    out_dtype = torch.bfloat16 if out_bf16 else torch.float32
    with torch.cuda.device(feats.device):
        out = _alloc_out(feats, Hb, Wb, per_view, out_dtype, _is_channels_last5(feats))
        d = _fill_desc(feats.shape, feats.stride(), _out_strides5(out, per_view), (Hb, Wb), (img_h, img_w), mode,
                       _DT[feats.dtype], _DT[out_dtype], 0, flags)
        _lib.check(L.bevipm_warp_fuse_fwd_planned(ctypes.byref(d), _ptr(feats), _ptr(K), _ptr(Rt34), _ptr(xs), _ptr(ys),
                                                  _ptr(out), _ptr(plan), plan.numel(), ctypes.c_void_p(_stream_ptr(feats.device))))
    return out


@warp_fuse_planned.register_fake
def _(feats, K, Rt34, xs, ys, img_h, img_w, mode, out_bf16, plan):
    B, V, C = feats.shape[:3]
    Hb, Wb = ys.numel(), xs.numel()
    dt = torch.bfloat16 if out_bf16 else torch.float32
    per_view = (mode & 0xff) == _lib.NONE
    if _is_channels_last5(feats):
        if per_view:
            return feats.new_empty((B, V, Hb, Wb, C), dtype=dt).permute(0, 1, 4, 2, 3)
        return feats.new_empty((B, Hb, Wb, C), dtype=dt).permute(0, 3, 1, 2)
    return feats.new_empty((B, V, C, Hb, Wb) if per_view else (B, C, Hb, Wb), dtype=dt)


def _setup_ctx_planned(ctx, inputs, output):
    _setup_ctx(ctx, inputs[:9] + (0,), output)


warp_fuse_planned.register_autograd(_backward, setup_context=_setup_ctx_planned)


def planned_ok(feats: torch.Tensor) -> bool:
    """Does this launch take the kernels a table cache belongs to (channels-last 16-byte vectors, V <= 32)?"""
    ve = 4 if feats.dtype == torch.float32 else 8
    return (feats.dim() == 5 and feats.dtype in _DT and feats.shape[1] <= 32 and _is_channels_last5(feats) and feats.shape[2] % ve == 0
            and all(st % ve == 0 for st in (feats.stride(0), feats.stride(1), feats.stride(3), feats.stride(4))) and feats.data_ptr() % 16 == 0)


# ---- the remaining entry points, as plain functions ----------------------------------------------

def sample_coords(K: torch.Tensor, Rt34: torch.Tensor, xs: torch.Tensor, ys: torch.Tensor,
                  feat_hw: Tuple[int, int], img_size: Tuple[int, int]):
    """ix, iy [B,V,Hb,Wb]: feature-pixel sample position of every BEV cell (device tensors)."""
    L = _lib.load()
    B, V = K.shape[:2]
    Hb, Wb = ys.numel(), xs.numel()
    ix = torch.empty((B, V, Hb, Wb), device=K.device, dtype=torch.float32)
    iy = torch.empty_like(ix)
    d = _fill_desc((B, V, 1, feat_hw[0], feat_hw[1]), (0,) * 5, (0,) * 5, (Hb, Wb), img_size, 0, 0, 0, 0)
    with torch.cuda.device(K.device):
        _lib.check(L.bevipm_sample_coords(ctypes.byref(d), _ptr(K), _ptr(Rt34), _ptr(xs), _ptr(ys), _ptr(ix), _ptr(iy),
                                          ctypes.c_void_p(_stream_ptr(K.device))))
    return ix, iy


class _ToChannelsLast(torch.autograd.Function):
    """Layout change only: the logical tensor is unchanged, so the gradient passes straight through."""

    @staticmethod
    def forward(ctx, feats):
        return _to_channels_last5_impl(feats)

    @staticmethod
    def backward(ctx, grad):
        return grad


def to_channels_last5(feats: torch.Tensor) -> torch.Tensor:
    """[B,V,C,H,W] NCHW-contiguous -> same logical tensor stored [B,V,H,W,C] (our transpose kernel)."""
    if feats.dim() != 5 or not feats.is_cuda:
        raise ValueError("to_channels_last5 wants a CUDA [B,V,C,H,W] tensor")
    if _is_channels_last5(feats):
        return feats
    if feats.dtype not in _DT:
        raise TypeError(f"dtype {feats.dtype} is not supported")
    if feats.requires_grad and torch.is_grad_enabled():
        return _ToChannelsLast.apply(feats)
    return _to_channels_last5_impl(feats)


def _to_channels_last5_impl(feats: torch.Tensor) -> torch.Tensor:
    src = feats.detach().contiguous()
    B, V, C, H, W = src.shape
    dst = torch.empty((B, V, H, W, C), device=src.device, dtype=src.dtype)
    with torch.cuda.device(src.device):
        _lib.check(_lib.load().bevipm_nchw_to_nhwc(_ptr(src), _ptr(dst), B * V, C, H, W, _DT[src.dtype],
                                                   ctypes.c_void_p(_stream_ptr(src.device))))
    return dst.permute(0, 1, 4, 2, 3)


def fuse_views(bev_maps: torch.Tensor, mode: str, out_dtype=None) -> torch.Tensor:
    """SimpleFusion on materialised per-view maps [B,V,C,H,W] -> [B,C,H,W] (fusion.py:17-22)."""
    if bev_maps.dim() != 5 or not bev_maps.is_cuda:
        raise ValueError("fuse_views wants a CUDA [B,V,C,H,W] tensor")
    if bev_maps.dtype not in _DT:
        raise TypeError(f"dtype {bev_maps.dtype} is not supported")
    B, V, C, H, W = bev_maps.shape
    cl = _is_channels_last5(bev_maps) and not bev_maps.is_contiguous()
    src = bev_maps.permute(0, 1, 3, 4, 2) if cl else bev_maps
    src = src.contiguous()
    out_dtype = out_dtype or bev_maps.dtype
    out = torch.empty((B, H, W, C) if cl else (B, C, H, W), device=src.device, dtype=out_dtype)
    with torch.cuda.device(src.device):
        _lib.check(_lib.load().bevipm_fuse_views(_ptr(src), _ptr(out), B, V, C * H * W, MODES[mode], _DT[src.dtype],
                                                 _DT[out_dtype], ctypes.c_void_p(_stream_ptr(src.device))))
    return out.permute(0, 3, 1, 2) if cl else out


def fuse_views_bwd(grad_out: torch.Tensor, mode: str, shape, channels_last: bool, bev_maps: torch.Tensor | None = None) -> torch.Tensor:
    """Gradient of fuse_views w.r.t. the per-view maps (autograd of fusion.py:17-22): [B,V,C,H,W] float32, channels-last
    in memory when the maps were.  max needs the forward input `bev_maps` and sends the gradient to the first view holding
    the maximum (torch.max's rule)."""
    B, V, C, H, W = shape
    cl = channels_last
    src = None
    if mode == "max":
        src = (bev_maps.permute(0, 1, 3, 4, 2) if cl else bev_maps).contiguous()
    g = (grad_out.permute(0, 2, 3, 1) if cl else grad_out).to(torch.float32).contiguous()
    gin = torch.empty((B, V, H, W, C) if cl else (B, V, C, H, W), device=g.device, dtype=torch.float32)
    with torch.cuda.device(g.device):
        _lib.check(_lib.load().bevipm_fuse_views_bwd(_ptr(src) if src is not None else None, _ptr(g), _ptr(gin), B, V, C * H * W,
                                                     MODES[mode], _DT[src.dtype] if src is not None else _lib.F32,
                                                     ctypes.c_void_p(_stream_ptr(g.device))))
    return gin.permute(0, 1, 4, 2, 3) if cl else gin


class _FuseViews(torch.autograd.Function):
    """SimpleFusion on our kernels in both directions (bevipm_fuse_views / bevipm_fuse_views_bwd)."""

    @staticmethod
    def forward(ctx, bev_maps, mode):
        ctx.mode = mode
        ctx.meta = (tuple(bev_maps.shape), _is_channels_last5(bev_maps) and not bev_maps.is_contiguous(), bev_maps.dtype)
        if mode == "max":
            ctx.save_for_backward(bev_maps)
        return fuse_views(bev_maps, mode)

    @staticmethod
    def backward(ctx, grad):
        shape, cl, dtype = ctx.meta
        src = ctx.saved_tensors[0].detach() if ctx.mode == "max" else None
        return fuse_views_bwd(grad, ctx.mode, shape, cl, src).to(dtype), None


def fuse_views_autograd(bev_maps: torch.Tensor, mode: str) -> torch.Tensor:
    return _FuseViews.apply(bev_maps, mode)


def valid_count(K: torch.Tensor, Rt34: torch.Tensor, xs: torch.Tensor, ys: torch.Tensor, feat_hw: Tuple[int, int],
                img_size: Tuple[int, int], flags: int = 0) -> torch.Tensor:
    """count [B,Hb,Wb] int32: how many views see each BEV cell (at least one bilinear tap inside the feature map).
    The validity-mask count of the north star; an extension (the reference's mean divides by V, fusion.py:20-21)."""
    B, V = K.shape[:2]
    Hb, Wb = ys.numel(), xs.numel()
    cnt = torch.empty((B, Hb, Wb), device=K.device, dtype=torch.int32)
    d = _fill_desc((B, V, 1, feat_hw[0], feat_hw[1]), (0,) * 5, (0,) * 5, (Hb, Wb), img_size, 0, 0, 0, 0, flags)
    with torch.cuda.device(K.device):
        _lib.check(_lib.load().bevipm_valid_count(ctypes.byref(d), _ptr(K), _ptr(Rt34), _ptr(xs), _ptr(ys), _ptr(cnt),
                                                  ctypes.c_void_p(_stream_ptr(K.device))))
    return cnt


def divide_by_count_(bev_sum: torch.Tensor, count: torch.Tensor) -> torch.Tensor:
    """In place: bev_sum [B,C,Hb,Wb] float32 (a SUM-mode result) /= max(count, 1): the mean over the views that see each
    cell.  Opt-in extension, NOT the reference's mean."""
    if bev_sum.dtype != torch.float32 or bev_sum.dim() != 4:
        raise ValueError("divide_by_count_ wants the float32 [B,C,Hb,Wb] result of sum fusion")
    B, C, Hb, Wb = bev_sum.shape
    s = bev_sum.stride()
    d = _fill_desc((B, 1, C, 1, 1), (0,) * 5, (s[0], 0, s[1], s[2], s[3]), (Hb, Wb), (1, 1), 0, 0, 0, 0)
    with torch.cuda.device(bev_sum.device):
        _lib.check(_lib.load().bevipm_divide_by_count(ctypes.byref(d), _ptr(bev_sum), _ptr(count.contiguous()),
                                                      ctypes.c_void_p(_stream_ptr(bev_sum.device))))
    return bev_sum


def warp_fuse_host(feats: torch.Tensor, K: torch.Tensor, Rt34: torch.Tensor, xs: torch.Tensor, ys: torch.Tensor,
                   img_size: Tuple[int, int], mode: str = "mean", out: torch.Tensor | None = None,
                   out_dtype=None, variant: int = 0) -> torch.Tensor:
    """Host-buffer call: feats [B,V,Hf,Wf,C] on the HOST (pinned for full PCIe rate) -> out [B,Hb,Wb,C]
    on the host.  H2D, kernel and D2H are pipelined frame by frame inside the library."""
    if feats.is_cuda or feats.dim() != 5 or not feats.is_contiguous():
        raise ValueError("warp_fuse_host wants a contiguous host tensor [B,V,Hf,Wf,C]")
    B, V, Hf, Wf, C = feats.shape
    Hb, Wb = ys.numel(), xs.numel()
    m = MODES[mode]
    out_dtype = out_dtype or feats.dtype
    shape = (B, V, Hb, Wb, C) if m == _lib.NONE else (B, Hb, Wb, C)
    if out is None:
        out = torch.empty(shape, dtype=out_dtype, pin_memory=True)
    if tuple(out.shape) != shape or out.dtype != out_dtype or not out.is_contiguous() or out.is_cuda:
        raise ValueError(f"out must be a contiguous host {out_dtype} tensor of shape {shape}")
    Kc, Rc = K.contiguous().float().cpu(), Rt34.contiguous().float().cpu()
    xc, yc = xs.contiguous().float().cpu(), ys.contiguous().float().cpu()
    d = _fill_desc((B, V, C, Hf, Wf), (0,) * 5, (0,) * 5, (Hb, Wb), img_size, m, _DT[feats.dtype], _DT[out_dtype], variant)
    _lib.check(_lib.load().bevipm_warp_fuse_host(ctypes.byref(d), _ptr(feats), _ptr(Kc), _ptr(Rc), _ptr(xc), _ptr(yc),
                                                 _ptr(out)))
    return out


# ---- Phase-2 follow-on: deformable-attention sampling ---------------------------------------------------

def _deform_fwd(value, shp, start, loc, aw, out_dtype):
    B, S, M, D = value.shape
    _, Q, _, Lv, P, _ = loc.shape
    dev = value.device
    out = torch.empty((B, Q, M * D), device=dev, dtype=out_dtype)
    d = _lib.DeformDesc()
    d.B, d.Q, d.M, d.D, d.L, d.P = B, Q, M, D, Lv, P
    d.value_dtype, d.out_dtype, d.S = _DT[value.dtype], _DT[out_dtype], S
    with torch.cuda.device(dev):
        _lib.check(_lib.load().bevipm_deform_attn_fwd(ctypes.byref(d), _ptr(value), _ptr(shp), _ptr(start), _ptr(loc), _ptr(aw),
                                                      _ptr(out), ctypes.c_void_p(_stream_ptr(dev))))
    return out


class _DeformAttn(torch.autograd.Function):
    """Forward and backward of the sampling in libbevipm.so (bevipm_deform_attn_fwd / _bwd)."""

    @staticmethod
    def forward(ctx, value, shp, start, loc, aw, out_dtype):
        ctx.save_for_backward(value, shp, start, loc, aw)
        ctx.out_dtype = out_dtype
        return _deform_fwd(value, shp, start, loc, aw, out_dtype)

    @staticmethod
    def backward(ctx, grad_out):
        value, shp, start, loc, aw = ctx.saved_tensors
        B, S, M, D = value.shape
        _, Q, _, Lv, P, _ = loc.shape
        dev = value.device
        g = grad_out.contiguous()
        if g.dtype not in _DT:
            g = g.float()
        gv = torch.zeros((B, S, M, D), device=dev, dtype=torch.float32)
        gl = torch.empty(loc.shape, device=dev, dtype=torch.float32)
        ga = torch.empty(aw.shape, device=dev, dtype=torch.float32)
        d = _lib.DeformDesc()
        d.B, d.Q, d.M, d.D, d.L, d.P = B, Q, M, D, Lv, P
        d.value_dtype, d.out_dtype, d.S = _DT[value.dtype], _DT[g.dtype], S
        with torch.cuda.device(dev):
            _lib.check(_lib.load().bevipm_deform_attn_bwd(ctypes.byref(d), _ptr(value), _ptr(shp), _ptr(start), _ptr(loc), _ptr(aw), _ptr(g),
                                                          _ptr(gv), _ptr(gl), _ptr(ga), ctypes.c_void_p(_stream_ptr(dev))))
        return gv.to(value.dtype), None, None, gl, ga, None


def deform_attn(value: torch.Tensor, spatial_shapes: torch.Tensor, level_start_index: torch.Tensor,
                sampling_locations: torch.Tensor, attention_weights: torch.Tensor, out_dtype=None) -> torch.Tensor:
    """MSDeformAttn (Deformable-DETR semantics, one level per camera view), forward and backward on our kernels.

    value [B,S,M,D] float32/bfloat16, spatial_shapes [L,2] (H,W), level_start_index [L],
    sampling_locations [B,Q,M,L,P,2] in [0,1], attention_weights [B,Q,M,L,P]  ->  [B,Q,M*D].
    Differentiable w.r.t. value, sampling_locations and attention_weights; CUDA tensors only.
    """
    if not value.is_cuda:
        raise RuntimeError("bevipm runs on CUDA tensors only: there is no CPU implementation of this path")
    if value.dim() != 4 or sampling_locations.dim() != 6 or attention_weights.dim() != 5:
        raise ValueError("expected value [B,S,M,D], sampling_locations [B,Q,M,L,P,2], attention_weights [B,Q,M,L,P]")
    if value.dtype not in _DT:
        raise TypeError(f"value dtype {value.dtype} is not supported (float32 / bfloat16)")
    B, S, M, D = value.shape
    _, Q, M2, Lv, P, two = sampling_locations.shape
    if M2 != M or two != 2 or tuple(attention_weights.shape) != (B, Q, M, Lv, P) or spatial_shapes.shape != (Lv, 2):
        raise ValueError("inconsistent deformable-attention shapes")
    dev = value.device
    value = value.contiguous()
    loc = sampling_locations.to(device=dev, dtype=torch.float32).contiguous()
    aw = attention_weights.to(device=dev, dtype=torch.float32).contiguous()
    shp = spatial_shapes.to(device=dev, dtype=torch.int32).contiguous()
    start = level_start_index.to(device=dev, dtype=torch.int64).contiguous()
    out_dtype = out_dtype or value.dtype
    if torch.is_grad_enabled() and (value.requires_grad or loc.requires_grad or aw.requires_grad):
        return _DeformAttn.apply(value, shp, start, loc, aw, out_dtype)
    return _deform_fwd(value, shp, start, loc, aw, out_dtype)


# ---- BEVNet's 1x1 projection on the source maps: hand-written tcgen05 GEMM (csrc/bevipm_proj.cu) ----------------------

def split_tf32(w: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """w = hi + lo with hi = w's TF32 head (the low 13 mantissa bits cleared) and lo the exact fp32 remainder."""
    w = w.contiguous().float()
    hi = (w.view(torch.int32) & -8192).view(torch.float32)
    return hi, (w - hi).contiguous()


def proj1x1_supported(C: int, Co: int) -> bool:
    return C % 4 == 0 and Co % 16 == 0 and 16 <= Co <= 256


def _proj1x1_raw(x: torch.Tensor, w: torch.Tensor, out: torch.Tensor, passes: int) -> torch.Tensor:
    """x [BV,rows,C] fp32 (unit channel stride), w [Co,V,C] fp32, out [BV,rows,Co] fp32 (unit channel stride; may be a
    channel slice of a wider tensor): out[m,r,:] = w[:, m % V, :] @ x[m,r,:]."""
    BV, rows, C = x.shape
    Co, V, C2 = w.shape
    assert C2 == C and out.shape == (BV, rows, Co) and x.stride(2) == 1 and out.stride(2) == 1 and BV % V == 0
    if passes == 3:
        w_hi, w_lo = split_tf32(w)
    else:
        w_hi, w_lo = w.contiguous().float(), None
    dev = x.device
    with torch.cuda.device(dev):
        _lib.check(_lib.load().bevipm_proj1x1(_ptr(x), _ptr(w_hi), _ptr(w_lo) if w_lo is not None else None, _ptr(out), BV, V, rows, C, Co,
                                              x.stride(1), x.stride(0), out.stride(1), out.stride(0), passes,
                                              ctypes.c_void_p(_stream_ptr(dev))))
    return out


class _Proj1x1(torch.autograd.Function):
    """Forward and input gradient on the tcgen05 kernel; the weight gradient (a reduction over all texels, both operands
    MN-major) is one torch.einsum."""

    @staticmethod
    def forward(ctx, x, w, passes):
        ctx.save_for_backward(x, w)
        ctx.passes = passes
        BV, rows, _ = x.shape
        out = torch.empty((BV, rows, w.shape[0]), device=x.device, dtype=torch.float32)
        return _proj1x1_raw(x, w, out, passes)

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        Co, V, C = w.shape
        BV, rows, _ = x.shape
        g = g.contiguous()
        gx = gw = None
        if ctx.needs_input_grad[0]:
            gx = torch.empty_like(x, memory_format=torch.contiguous_format)
            wt = w.permute(2, 1, 0).contiguous()                          # [C, V, Co]: the transposed weights of every view
            if proj1x1_supported(Co, 16):
                for n0 in range(0, C, 256):                               # N of one launch is at most 256
                    n1 = min(C, n0 + 256)
                    if (n1 - n0) % 16:
                        gx[:, :, n0:n1] = torch.einsum("mvro,ovc->mvrc", g.view(BV // V, V, rows, Co), w[:, :, n0:n1]).reshape(BV, rows, n1 - n0)
                    else:
                        _proj1x1_raw(g, wt[n0:n1], gx[:, :, n0:n1], ctx.passes)
            else:
                gx = torch.einsum("mvro,ovc->mvrc", g.view(BV // V, V, rows, Co), w).reshape(BV, rows, C)
        if ctx.needs_input_grad[1]:
            gw = torch.einsum("mvro,mvrc->ovc", g.view(BV // V, V, rows, Co), x.reshape(BV // V, V, rows, C))
        return gx, gw, None


def proj1x1(x: torch.Tensor, w: torch.Tensor, passes: int = 3) -> torch.Tensor:
    """Per-view 1x1 projection on channels-last source maps: x [BV,rows,C] fp32, w [Co,V,C] -> [BV,rows,Co] fp32.
    passes = 1: one TF32 pass; 3: split operands (fp32-grade).  model_wrapper.py:70-73 folded in front of the warp."""
    if not x.is_cuda:
        raise RuntimeError("bevipm runs on CUDA tensors only: there is no CPU implementation of this path")
    if x.dtype != torch.float32 or x.dim() != 3 or w.dim() != 3 or x.stride(2) != 1:
        raise ValueError("proj1x1 expects x [BV,rows,C] float32 with unit channel stride and w [Co,V,C]")
    if not proj1x1_supported(x.shape[2], w.shape[0]):
        raise ValueError(f"proj1x1: C={x.shape[2]} must be a multiple of 4 and Co={w.shape[0]} a multiple of 16 in [16, 256]")
    w = w.to(device=x.device, dtype=torch.float32)
    if torch.is_grad_enabled() and (x.requires_grad or w.requires_grad):
        return _Proj1x1.apply(x, w, passes)
    out = torch.empty((x.shape[0], x.shape[1], w.shape[0]), device=x.device, dtype=torch.float32)
    return _proj1x1_raw(x, w, out, passes)
