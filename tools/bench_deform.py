#!/usr/bin/env python
"""Time the deformable-attention sampling kernel at BASELINE config 3 (120x360 queries, 8 heads x 4 points x
7 views, 256 ch bf16, views as 135x240 maps).  CUDA events, value resident in HBM."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "vision-based-spatio-temporal-analysis_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402

from bevipm import ops  # noqa: E402


def main():
    B, Q, M, D, L, P = 1, 120 * 360, 8, 32, 7, 4
    H, W = 135, 240
    g = torch.Generator(device="cuda").manual_seed(0)
    nbuf = 3
    values = [torch.randn(B, L * H * W, M, D, device="cuda", generator=g).bfloat16() for _ in range(nbuf)]
    # samples clustered around each query's own position in every view (what a trained module produces)
    qy, qx = torch.meshgrid(torch.arange(120, device="cuda"), torch.arange(360, device="cuda"), indexing="ij")
    ref = torch.stack([(qx + 0.5) / 360, (qy + 0.5) / 120], -1).reshape(1, Q, 1, 1, 1, 2)
    loc = (ref + 0.02 * torch.randn(B, Q, M, L, P, 2, device="cuda", generator=g)).float().contiguous()
    aw = torch.softmax(torch.randn(B, Q, M, L * P, device="cuda", generator=g), -1).view(B, Q, M, L, P).contiguous()
    shapes = torch.tensor([[H, W]] * L, dtype=torch.int32, device="cuda")
    start = torch.arange(L, device="cuda", dtype=torch.int64) * (H * W)
    for i in range(5):
        ops.deform_attn(values[i % nbuf], shapes, start, loc, aw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 100
    e0.record()
    for i in range(iters):
        ops.deform_attn(values[i % nbuf], shapes, start, loc, aw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    nbytes = values[0].numel() * 2 + loc.numel() * 4 + aw.numel() * 4 + B * Q * M * D * 2
    # forward + backward (grad value fp32 zero-fill + atomics, grad loc, grad attn) minus forward
    vg = [v.clone().requires_grad_(True) for v in values]
    lg, ag = loc.clone().requires_grad_(True), aw.clone().requires_grad_(True)
    cot = torch.randn(B, Q, M * D, device="cuda", generator=g).bfloat16()

    def fb(i):
        out = ops.deform_attn(vg[i % nbuf], shapes, start, lg, ag)
        out.backward(cot)
        vg[i % nbuf].grad = None
        lg.grad = None
        ag.grad = None

    for i in range(3):
        fb(i)
    torch.cuda.synchronize()
    e0.record()
    for i in range(30):
        fb(i)
    e1.record()
    torch.cuda.synchronize()
    ms_fb = e0.elapsed_time(e1) / 30
    peak = 6543.1
    try:
        peak = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
    except Exception:
        pass
    print(json.dumps({"workload": "c4 deformable attention 120x360 queries, 8 heads x 4 pts x 7 views, 256 ch bf16",
                      "ms": ms, "frames_per_s": B / (ms * 1e-3), "bytes_full_per_frame": nbytes,
                      "gbs_full": nbytes / (ms * 1e-3) / 1e9, "roofline_frac_of_measured_hbm": nbytes / (ms * 1e-3) / 1e9 / peak,
                      "tap_bytes_through_l1": B * Q * M * L * P * 4 * D * 2,
                      "forward_plus_backward_ms": ms_fb, "backward_ms": ms_fb - ms}))


if __name__ == "__main__":
    main()
