"""ctypes binding of include/bevipm.h -- the stub INTEGRATION.md shows, shipped.

There is no fallback: if libbevipm.so is missing or was built without CUDA the import of the
product path fails with the reason and the build command.
"""
from __future__ import annotations

import ctypes
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = Path(__import__("os").environ.get("BEVIPM_LIB", PKG / "libbevipm.so"))  # BEVIPM_LIB: A/B of two builds (development aid)

F32, BF16 = 0, 1
FLAG_KORNIA_GEOMETRY = 1
FLAG_SLAB_PUT = 2
SUM, MEAN, MAX, NONE = 0, 1, 2, 3
MODES = {"sum": SUM, "mean": MEAN, "max": MAX, "none": NONE, "concat": NONE}


class Desc(ctypes.Structure):
    """struct bevipm_desc (include/bevipm.h)."""
    _fields_ = [(n, ctypes.c_int32) for n in
                ("B", "V", "C", "Hf", "Wf", "Hb", "Wb", "img_h", "img_w", "mode", "in_dtype", "out_dtype",
                 "variant", "flags")] + \
               [(n, ctypes.c_int64) for n in
                ("fs_b", "fs_v", "fs_c", "fs_y", "fs_x", "os_b", "os_v", "os_c", "os_y", "os_x")]


EXPORTS = (
    "bevipm_version", "bevipm_last_error", "bevipm_launch_count", "bevipm_warp_fuse_fwd",
    "bevipm_warp_fuse_bwd", "bevipm_sample_coords", "bevipm_nchw_to_nhwc", "bevipm_fuse_views",
    "bevipm_warp_fuse_host", "bevipm_host_release", "bevipm_deform_attn_fwd", "bevipm_last_variant", "bevipm_host_last_h2d_bytes",
    "bevipm_fuse_views_bwd", "bevipm_valid_count", "bevipm_divide_by_count", "bevipm_warp_fuse_red", "bevipm_slab_finish", "bevipm_deform_attn_bwd",
    "bevipm_proj1x1", "bevipm_plan_bytes", "bevipm_warp_fuse_fwd_planned",
)

class DeformDesc(ctypes.Structure):
    """struct bevipm_deform_desc (include/bevipm.h)."""
    _fields_ = [(n, ctypes.c_int32) for n in ("B", "Q", "M", "D", "L", "P", "value_dtype", "out_dtype")] + [("S", ctypes.c_int64)]


_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA library is the only implementation of this path. "
            "Build it with `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc).")
    L = ctypes.CDLL(str(LIB_PATH))
    vp, fp, dp = ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(Desc)
    L.bevipm_version.restype = ctypes.c_int
    L.bevipm_last_error.restype = ctypes.c_char_p
    L.bevipm_launch_count.restype = ctypes.c_int64
    L.bevipm_last_variant.restype = ctypes.c_int32
    L.bevipm_host_last_h2d_bytes.restype = ctypes.c_int64
    L.bevipm_warp_fuse_fwd.argtypes = [dp, vp, fp, fp, fp, fp, vp, vp]
    L.bevipm_warp_fuse_bwd.argtypes = [dp, vp, fp, fp, fp, fp, vp, vp]
    L.bevipm_sample_coords.argtypes = [dp, fp, fp, fp, fp, fp, fp, vp]
    L.bevipm_nchw_to_nhwc.argtypes = [vp, vp, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                      ctypes.c_int32, vp]
    L.bevipm_fuse_views.argtypes = [vp, vp, ctypes.c_int64, ctypes.c_int32, ctypes.c_int64, ctypes.c_int32,
                                    ctypes.c_int32, ctypes.c_int32, vp]
    L.bevipm_fuse_views_bwd.argtypes = [vp, fp, fp, ctypes.c_int64, ctypes.c_int32, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, vp]
    L.bevipm_valid_count.argtypes = [dp, fp, fp, fp, fp, vp, vp]
    L.bevipm_divide_by_count.argtypes = [dp, vp, vp, vp]
    L.bevipm_warp_fuse_red.argtypes = [dp, vp, fp, fp, fp, fp, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int32, ctypes.c_int32, vp]
    L.bevipm_slab_finish.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int32, vp, ctypes.c_int64, ctypes.c_float, vp]
    L.bevipm_warp_fuse_host.argtypes = [dp, vp, fp, fp, fp, fp, vp]
    L.bevipm_host_release.restype = None
    L.bevipm_deform_attn_fwd.argtypes = [ctypes.POINTER(DeformDesc), vp, vp, vp, fp, fp, vp, vp]
    L.bevipm_deform_attn_fwd.restype = ctypes.c_int
    L.bevipm_deform_attn_bwd.argtypes = [ctypes.POINTER(DeformDesc), vp, vp, vp, fp, fp, vp, fp, fp, fp, vp]
    L.bevipm_deform_attn_bwd.restype = ctypes.c_int
    i32, i64 = ctypes.c_int32, ctypes.c_int64
    if hasattr(L, "bevipm_plan_bytes"):
        L.bevipm_plan_bytes.argtypes = [dp]
        L.bevipm_plan_bytes.restype = ctypes.c_int64
        L.bevipm_warp_fuse_fwd_planned.argtypes = [dp, vp, fp, fp, fp, fp, vp, vp, ctypes.c_int64, vp]
        L.bevipm_warp_fuse_fwd_planned.restype = ctypes.c_int
    if hasattr(L, "bevipm_proj1x1"):   # (absent only from an older build loaded through BEVIPM_LIB for an A/B)
        L.bevipm_proj1x1.argtypes = [fp, fp, fp, fp, i32, i32, i64, i32, i32, i64, i64, i64, i64, i32, vp]
        L.bevipm_proj1x1.restype = ctypes.c_int
    for name in ("bevipm_warp_fuse_fwd", "bevipm_warp_fuse_bwd", "bevipm_sample_coords", "bevipm_nchw_to_nhwc",
                 "bevipm_fuse_views", "bevipm_warp_fuse_host", "bevipm_fuse_views_bwd", "bevipm_valid_count", "bevipm_divide_by_count", "bevipm_warp_fuse_red", "bevipm_slab_finish"):
        getattr(L, name).restype = ctypes.c_int
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        msg = load().bevipm_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"bevipm error {rc}: {msg}")


def launch_count() -> int:
    return int(load().bevipm_launch_count())


def variant_name(v: int) -> str:
    """Kernel behind a `variant` number of bevipm_desc (see csrc/bevipm_api.cu dispatch_fused)."""
    if v == -1:
        return "warp_fuse_strided_kernel"
    if 1 <= v <= 14:
        return f"warp_fuse_nhwc_kernel (tile kernel, variant {v})"
    if 20 <= v <= 27:
        return f"warp_fuse_list_kernel (variant {v})"
    if 30 <= v <= 41:
        return f"warp_fuse_run_kernel (variant {v})"
    if v == 55:
        return "warp_fuse_boxrun_kernel (run kernel, ring filled by TMA 2x2 box copies, variant 55)"
    if v == 71:
        return "proj1x1_t_kernel (tcgen05 TF32 GEMM of the folded 1x1 projection, transposed form for 128 output channels, variant 71)"
    if v == 70:
        return "proj1x1_tf32_kernel (tcgen05 TF32 GEMM of the folded 1x1 projection, variant 70)"
    if v == 60:
        return "warp_fuse_run_kernel<KM_RED> (partial sums added into peer slabs)"
    if 50 <= v <= 53:
        return f"warp_fuse_staged_kernel (TMA-staged tiles, variant {v})"
    return f"variant {v}"
