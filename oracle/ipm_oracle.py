"""ctypes front-end of oracle/ipm_oracle.c (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does; it fails loudly without its CUDA library.

Restates /root/reference/project/models/fusion/geometry.py:80-163 (grid_sample branch) and
fusion.py:11-46 on numpy arrays; see the C file header for the op-by-op citation.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "libipm_oracle.so"
_SRC = _HERE / "ipm_oracle.c"

MODES = {"sum": 0, "mean": 1, "max": 2, "none": 3, "concat": 3}


class _Desc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ("B", "V", "C", "Hf", "Wf", "Hb", "Wb", "img_h", "img_w", "mode")] + \
               [(n, ctypes.c_int64) for n in
                ("fs_b", "fs_v", "fs_c", "fs_y", "fs_x", "os_b", "os_v", "os_c", "os_y", "os_x")]


def build(force: bool = False) -> Path:
    """Compile the C oracle with the recipe in oracle/Makefile (gcc is in the image)."""
    if force or not _SO.exists() or _SO.stat().st_mtime < _SRC.stat().st_mtime:
        subprocess.run(["make", "-s", "-B", "-C", str(_HERE), "libipm_oracle.so"], check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(str(_SO))
        fp = ctypes.POINTER(ctypes.c_float)
        dp = ctypes.POINTER(_Desc)
        L.ipm_oracle_homography.argtypes = [fp, fp, fp]
        L.ipm_oracle_homography.restype = None
        L.ipm_oracle_coords.argtypes = [dp, fp, fp, fp, fp, fp, fp]
        L.ipm_oracle_warp_fuse.argtypes = [dp, fp, fp, fp, fp, fp, fp, ctypes.c_int]
        L.ipm_oracle_warp_fuse_bwd.argtypes = [dp, fp, fp, fp, fp, fp, fp]
        L.ipm_oracle_num_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def rt34(extrinsics) -> np.ndarray:
    """[...,4,4] / [...,3,4] / [...,3,3] extrinsics -> [...,3,4] (geometry.py:41-52)."""
    e = np.asarray(extrinsics, dtype=np.float32)
    if e.shape[-2:] == (4, 4):
        return np.ascontiguousarray(e[..., :3, :])
    if e.shape[-2:] == (3, 4):
        return np.ascontiguousarray(e)
    if e.shape[-2:] == (3, 3):
        z = np.zeros(e.shape[:-1] + (1,), np.float32)
        return np.ascontiguousarray(np.concatenate([e, z], axis=-1))
    raise ValueError(f"unsupported extrinsics shape {e.shape}")


def _estrides(a: np.ndarray):
    return [s // a.itemsize for s in a.strides]


def _desc(B, V, C, Hf, Wf, Hb, Wb, img_size, mode, fstr, ostr) -> _Desc:
    d = _Desc()
    d.B, d.V, d.C, d.Hf, d.Wf, d.Hb, d.Wb = B, V, C, Hf, Wf, Hb, Wb
    d.img_h, d.img_w = int(img_size[0]), int(img_size[1])
    d.mode = MODES[mode]
    d.fs_b, d.fs_v, d.fs_c, d.fs_y, d.fs_x = fstr
    d.os_b, d.os_v, d.os_c, d.os_y, d.os_x = ostr
    return d


def homography(K, Rt) -> np.ndarray:
    K = _f32(np.asarray(K)[:3, :3])
    R = rt34(Rt)
    H = np.empty((3, 3), np.float32)
    lib().ipm_oracle_homography(_ptr(K), _ptr(R), _ptr(H))
    return H


def coords(K, Rt, xs, ys, feat_hw, img_size):
    """ix, iy [B,V,Hb,Wb] float32: the feature-pixel sample position of every BEV cell."""
    K = _f32(K)
    R = rt34(Rt)
    B, V = K.shape[:2]
    xs, ys = _f32(xs), _f32(ys)
    Hb, Wb = len(ys), len(xs)
    d = _desc(B, V, 1, feat_hw[0], feat_hw[1], Hb, Wb, img_size, "sum", (0,) * 5, (0,) * 5)
    ix = np.empty((B, V, Hb, Wb), np.float32)
    iy = np.empty_like(ix)
    lib().ipm_oracle_coords(ctypes.byref(d), _ptr(K), _ptr(R), _ptr(xs), _ptr(ys), _ptr(ix), _ptr(iy))
    return ix, iy


def warp_fuse(feats, K, Rt, xs, ys, img_size, mode="mean", nthreads=0, channels_last_out=False):
    """feats [B,V,C,Hf,Wf] fp32 (any numpy strides) -> [B,C,Hb,Wb] or [B,V,C,Hb,Wb] ('none')."""
    feats = np.asarray(feats)
    assert feats.dtype == np.float32 and feats.ndim == 5
    K = _f32(K)
    R = rt34(Rt)
    B, V, C, Hf, Wf = feats.shape
    assert K.shape == (B, V, 3, 3) and R.shape == (B, V, 3, 4)
    xs, ys = _f32(xs), _f32(ys)
    Hb, Wb = len(ys), len(xs)
    per_view = MODES[mode] == 3
    if per_view:
        if channels_last_out:
            out = np.empty((B, V, Hb, Wb, C), np.float32).transpose(0, 1, 4, 2, 3)
        else:
            out = np.empty((B, V, C, Hb, Wb), np.float32)
        ostr = _estrides(out)
    else:
        if channels_last_out:
            out = np.empty((B, Hb, Wb, C), np.float32).transpose(0, 3, 1, 2)
        else:
            out = np.empty((B, C, Hb, Wb), np.float32)
        s = _estrides(out)
        ostr = [s[0], 0, s[1], s[2], s[3]]
    d = _desc(B, V, C, Hf, Wf, Hb, Wb, img_size, mode, _estrides(feats), ostr)
    rc = lib().ipm_oracle_warp_fuse(ctypes.byref(d), _ptr(feats), _ptr(K), _ptr(R), _ptr(xs), _ptr(ys),
                                    _ptr(out), int(nthreads))
    if rc != 0:
        raise RuntimeError(f"ipm_oracle_warp_fuse rc={rc}")
    return out


def warp_fuse_bwd(gout, K, Rt, xs, ys, feat_shape, img_size, mode="mean"):
    """grad wrt feats [B,V,C,Hf,Wf] given grad of the forward output (contiguous NCHW)."""
    gout = _f32(gout)
    K = _f32(K)
    R = rt34(Rt)
    B, V, C, Hf, Wf = feat_shape
    xs, ys = _f32(xs), _f32(ys)
    Hb, Wb = len(ys), len(xs)
    g = np.zeros(feat_shape, np.float32)
    s = _estrides(gout)
    ostr = s if MODES[mode] == 3 else [s[0], 0, s[1], s[2], s[3]]
    d = _desc(B, V, C, Hf, Wf, Hb, Wb, img_size, mode, _estrides(g), ostr)
    rc = lib().ipm_oracle_warp_fuse_bwd(ctypes.byref(d), _ptr(gout), _ptr(K), _ptr(R), _ptr(xs),
                                        _ptr(ys), _ptr(g))
    if rc != 0:
        raise RuntimeError(f"ipm_oracle_warp_fuse_bwd rc={rc}")
    return g


def num_threads() -> int:
    return int(lib().ipm_oracle_num_threads())


def valid_count(K, Rt, xs, ys, feat_hw, img_size) -> np.ndarray:
    """count [B,Hb,Wb] int32: views whose sample position has at least one bilinear tap inside the feature map, i.e. the
    cells where grid_sample (geometry.py:161: zeros padding, align_corners=False) reads anything but padding.  The
    reference itself has no such output (its mean divides by V, fusion.py:20-21): the restatement of the north star's
    "validity-mask counts", from the same coordinates as the warp."""
    ix, iy = coords(K, Rt, xs, ys, feat_hw, img_size)
    Hf, Wf = feat_hw
    fin = np.isfinite(ix) & np.isfinite(iy)
    x0 = np.floor(np.where(fin, ix, -5.0))
    y0 = np.floor(np.where(fin, iy, -5.0))
    seen = fin & (x0 >= -1) & (x0 <= Wf - 1) & (y0 >= -1) & (y0 <= Hf - 1)
    return seen.sum(axis=1).astype(np.int32)


def warp_fuse_mean_valid(feats, K, Rt, xs, ys, img_size):
    """Opt-in extension: the sequential fp32 sum over views (fusion.py:18) divided by max(valid_count, 1) (IEEE fp32)."""
    s = warp_fuse(feats, K, Rt, xs, ys, img_size, "sum")
    cnt = valid_count(K, Rt, xs, ys, feats.shape[-2:], img_size)
    return (s / np.maximum(cnt, 1).astype(np.float32)[:, None]).astype(np.float32), cnt
