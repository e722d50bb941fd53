#!/usr/bin/env python
"""BEVNet's view projection + concat + 1x1 projection at the reference's own yaml shape (wildtrack.yaml: 7 views,
FEAT_DIM 1280, BEV 120x360, projection to 128 channels) on one GPU:
  unfolded: bevipm.GeometryTransformer (per-view maps) -> ConcatFusion -> nn.Conv2d 1x1   (model_wrapper.py:68-73)
  folded  : bevipm.FoldedConcatProjIPM (per-view GEMM on the source maps, then the fused SUM kernel), with the GEMM on
            our tcgen05 kernel (split operands = fp32-grade, or one TF32 pass) or on cuBLAS (torch.einsum)
and the per-view GEMM alone, [7 x 32400, 1280] x [1280, 128]: ms, algorithmic GB/s (x read once + W once + out written
once) against the measured copy bandwidth, and the distance to a float64 result.
Prints one JSON object (CUDA events, 3 warm-ups; the GEMM inputs rotate through 3 buffers = 3.5 GB > L2)."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "vision-based-spatio-temporal-analysis_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bevipm  # noqa: E402
from bevipm import ops, rig  # noqa: E402

dev = "cuda:0"
B, V, C, Co, fhw, bhw = 1, 7, 1280, 128, (135, 240), (120, 360)
K, Rt = rig.look_at_rig(V, 0)
K, Rt = K[None].to(dev), Rt[None].to(dev)
g = torch.Generator(device=dev).manual_seed(0)
feats = [torch.randn(B, V, *fhw, C, device=dev, generator=g).permute(0, 1, 4, 2, 3) for _ in range(3)]   # channels-last in memory
proj = torch.nn.Conv2d(V * C, Co, 1).to(dev)
geom = bevipm.GeometryTransformer(*bhw, rig.WILDTRACK_BOUNDS, warp_impl="kornia", emulate_kornia=False).to(dev)
cat = bevipm.ConcatFusion()


def timeit(fn, iters=20):
    with torch.no_grad():
        for i in range(3):
            out = fn(i)
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            out = fn(i)
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, torch.cuda.max_memory_allocated() / 1e9, out


peak = 6543.1
try:
    peak = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    pass

res = {"workload": "wildtrack.yaml projection: 7 views x 1280 ch x 135x240 fp32 -> 128 ch, BEV 120x360, 1 frame", "peak_gbs": peak}

# ---- the GEMM alone -------------------------------------------------------------------------------------------------
W = proj.weight.detach().view(Co, V, C).contiguous()
xs = [f.permute(0, 1, 3, 4, 2).reshape(B * V, fhw[0] * fhw[1], C) for f in feats]
want = torch.einsum("vrc,ovc->vro", xs[0][:, :4096].double(), W.double())
alg = xs[0].numel() * 4 + W.numel() * 4 + B * V * fhw[0] * fhw[1] * Co * 4
gemm = {}
for name, fn in (("tcgen05_split3", lambda i: ops.proj1x1(xs[i % 3], W, 3)),
                 ("tcgen05_tf32", lambda i: ops.proj1x1(xs[i % 3], W, 1)),
                 ("cublas_fp32", lambda i: torch.einsum("vrc,ovc->vro", xs[i % 3], W)),
                 ("cublas_tf32", lambda i: torch.einsum("vrc,ovc->vro", xs[i % 3], W))):
    tf32 = name == "cublas_tf32"
    torch.backends.cuda.matmul.allow_tf32 = tf32
    ms, _, out = timeit(fn, 30)
    out0 = fn(0)
    err = float((out0[:, :4096].double() - want).abs().max() / want.abs().max())
    gemm[name] = {"ms": round(ms, 4), "alg_gbs": round(alg / ms / 1e6, 1), "frac_of_measured_hbm": round(alg / ms / 1e6 / peak, 3),
                  "tflops": round(2.0 * B * V * fhw[0] * fhw[1] * C * Co / ms / 1e9, 1), "max_rel_err_vs_float64": err}
torch.backends.cuda.matmul.allow_tf32 = False
res["gemm_alone"] = gemm
res["gemm_algorithmic_bytes"] = alg

# ---- the module ------------------------------------------------------------------------------------------------------
mods = {}
for name, kw in (("folded_tcgen05_split3", dict(precision="fp32", gemm="tcgen05")), ("folded_tcgen05_tf32", dict(precision="tf32", gemm="tcgen05")),
                 ("folded_cublas_fp32", dict(gemm="cublas"))):
    m = bevipm.FoldedConcatProjIPM(*bhw, rig.WILDTRACK_BOUNDS, proj, views=V, **kw).to(dev)
    ms, mem, out = timeit(lambda i: m(feats[i % 3], K, Rt, img_size=rig.WILDTRACK_IMG_SIZE))
    mods[name] = {"ms_per_frame": round(ms, 4), "peak_mem_gb": round(mem, 2), "out": out}
for tf32 in (False, True):
    torch.backends.cudnn.allow_tf32 = tf32
    ms, mem, out = timeit(lambda i: proj(cat(geom(feats[i % 3], K, Rt, img_size=rig.WILDTRACK_IMG_SIZE))))
    mods["unfolded_conv_tf32" if tf32 else "unfolded_conv_fp32"] = {"ms_per_frame": round(ms, 4), "peak_mem_gb": round(mem, 2), "out": out}
ref = mods["unfolded_conv_fp32"]["out"]
for k, v in mods.items():
    o = v.pop("out")
    v["max_rel_diff_vs_unfolded_fp32"] = float((o - ref).abs().max() / ref.abs().max())
    v["speedup_vs_unfolded_fp32"] = round(mods["unfolded_conv_fp32"]["ms_per_frame"] / v["ms_per_frame"], 2)
res["module"] = mods
print(json.dumps(res))
