#!/usr/bin/env python
"""BEVNet's view projection + concat + 1x1 projection at the reference's own yaml shape (wildtrack.yaml: 7 views,
FEAT_DIM 1280, BEV 120x360, projection to 128 channels), two ways on one GPU:
  unfolded: bevipm.GeometryTransformer (per-view maps) -> ConcatFusion -> nn.Conv2d 1x1   (model_wrapper.py:68-73)
  folded  : bevipm.FoldedConcatProjIPM (per-view GEMM on the source maps, then the fused SUM kernel)
Prints ms per frame (CUDA events) and the peak memory of each."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "vision-based-spatio-temporal-analysis_b200")):
    sys.path.insert(0, p)
import torch
import bevipm
from bevipm import rig

dev = "cuda:0"
B, V, C, Co, fhw, bhw = 1, 7, 1280, 128, (135, 240), (120, 360)
K, Rt = rig.look_at_rig(V, 0)
K, Rt = K[None].to(dev), Rt[None].to(dev)
feats = torch.randn(B, V, *fhw, C, device=dev).permute(0, 1, 4, 2, 3)   # channels-last in memory
proj = torch.nn.Conv2d(V * C, Co, 1).to(dev)
geom = bevipm.GeometryTransformer(*bhw, rig.WILDTRACK_BOUNDS, warp_impl="kornia").to(dev)
cat = bevipm.ConcatFusion()
folded = bevipm.FoldedConcatProjIPM(*bhw, rig.WILDTRACK_BOUNDS, proj, views=V).to(dev)


def unfolded_fwd():
    return proj(cat(geom(feats, K, Rt, img_size=rig.WILDTRACK_IMG_SIZE)))


def folded_fwd():
    return folded(feats, K, Rt, img_size=rig.WILDTRACK_IMG_SIZE)


def timeit(fn, iters=20):
    with torch.no_grad():
        for _ in range(3):
            out = fn()
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, torch.cuda.max_memory_allocated() / 1e9, out


for tf32 in (False, True):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a_ms, a_mem, a = timeit(unfolded_fwd)
    b_ms, b_mem, b = timeit(folded_fwd)
    rel = float((a - b).abs().max() / a.abs().max())
    print(f"tf32={tf32}: unfolded {a_ms:.3f} ms/frame (peak {a_mem:.2f} GB)   folded {b_ms:.3f} ms/frame (peak {b_mem:.2f} GB)   "
          f"speed-up {a_ms / b_ms:.2f}x   max rel diff {rel:.2e}")
