"""Multi-GPU partitioning of the warp+fuse path: one process per GPU, torch.distributed plumbing.

The reference is single-process (train.py:114, inference.py:24) -- this is new work defined by
BASELINE.json / SURVEY.md 8(e):

  frames : frames of a batch / temporal clip are independent -> contiguous blocks per rank,
           NO data-path collective (optional all_gather when one rank needs every BEV).
  views  : out = reduce_v warp_v(f_v) -> rank r warps its camera subset into a partial BEV
           (sum, fp32) and ONE all-reduce(sum) over NCCL/NVLink finishes the fusion; mean divides
           by the GLOBAL view count afterwards (fusion.py:20-21 divides by V, not by coverage);
           max uses all-reduce(max) (idle ranks contribute -inf).

The partial warp is injected as a callable so the sharding logic is testable on CPU ranks over
gloo with the oracle standing in (tests/test_sharding.py); the product passes FusedIPM.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def block_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) of n items for `rank` (the first n % world ranks get one more)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    q, r = divmod(n, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def view_assignment(views: int, world: int) -> List[List[int]]:
    """Which cameras each rank warps.  7 views -> 4+3 (2 ranks), 2+2+2+1 (4), 1 x 7 + one idle rank (8)."""
    return [list(range(*block_range(views, r, world))) for r in range(world)]


def frame_assignment(frames: int, world: int) -> List[Tuple[int, int]]:
    return [block_range(frames, r, world) for r in range(world)]


class ViewShardedFusion:
    """Fuse over views held by different ranks.

    partial_fn(view_ids) -> Tensor [B,C,Hb,Wb] fp32: this rank's SUM (or MAX) over `view_ids`;
    it is not called when the rank has no views (more ranks than cameras).
    """

    def __init__(self, views: int, mode: str = "mean", group=None):
        if mode not in ("sum", "mean", "max"):
            raise ValueError("view sharding supports sum / mean / max")
        self.views = views
        self.mode = mode
        self.group = group

    def __call__(self, partial_fn: Callable[[Sequence[int]], torch.Tensor], out_shape, device, async_op: bool = False):
        world = dist.get_world_size(self.group)
        rank = dist.get_rank(self.group)
        mine = view_assignment(self.views, world)[rank]
        if mine:
            part = partial_fn(mine)
            if tuple(part.shape) != tuple(out_shape):
                raise ValueError(f"partial has shape {tuple(part.shape)}, expected {tuple(out_shape)}")
        else:
            # identity element of the reduction (a rank without cameras adds no view: the zeros of
            # fusion.py:22 come from real views that miss a cell and are already inside their maps)
            fill = float("-inf") if self.mode == "max" else 0.0
            part = torch.full(out_shape, fill, device=device, dtype=torch.float32)
        buf = _dense_view(part)
        op = dist.ReduceOp.MAX if self.mode == "max" else dist.ReduceOp.SUM
        work = dist.all_reduce(buf, op=op, group=self.group, async_op=async_op)
        if async_op:
            return part, work
        return self.finish(part)

    def finish(self, reduced: torch.Tensor) -> torch.Tensor:
        if self.mode == "mean":
            reduced.div_(float(self.views))
        return reduced


def _dense_view(t: torch.Tensor) -> torch.Tensor:
    """A contiguous alias of a dense-but-permuted tensor (channels-last output), so the collective
    works on the storage in place instead of on a copy."""
    if t.is_contiguous():
        return t
    order = sorted(range(t.dim()), key=lambda k: -t.stride(k))
    p = t.permute(order)
    if not p.is_contiguous():
        raise ValueError("partial BEV must be dense")
    return p


def gather_frames(local: torch.Tensor, frames: int, group=None) -> torch.Tensor:
    """Optional: assemble every rank's frame block [b_r, ...] into [frames, ...] on all ranks."""
    world = dist.get_world_size(group)
    sizes = [hi - lo for lo, hi in frame_assignment(frames, world)]
    pad = max(sizes)
    buf = local.new_zeros((pad,) + tuple(local.shape[1:]))
    buf[: local.shape[0]] = local
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf, group=group)
    return torch.cat([o[:n] for o, n in zip(outs, sizes)], dim=0)
