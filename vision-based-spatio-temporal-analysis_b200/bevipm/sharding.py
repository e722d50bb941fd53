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
            reduced.div_(torch.tensor(float(self.views), device=reduced.device))   # tensor divisor: IEEE division on CUDA too
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


# ---- view sharding that leaves the BEV sharded by rows --------------------------------------------------
# The all-reduce above replicates the whole fp32 BEV on every rank: 2 (N-1)/N x 354 MB per rank and frame at BASELINE
# configs[2], more than the single-GPU kernel reads and writes.  The consumer of a BEV grid (detector head, the next
# fusion stage) is itself row-separable, so the two forms below leave rank r with rows [r * Hb/N, (r+1) * Hb/N) only:
# (N-1)/N x 354 MB per rank and frame cross NVLink, once.

def slab_rows(bev_h: int, world: int) -> int:
    """BEV rows per rank (the last slab is padded when world does not divide bev_h)."""
    return -(-bev_h // world)


class ReduceScatterFusion:
    """Partial BEV per rank -> NCCL reduce-scatter -> this rank's row slab, on a side stream.

    submit(part) queues the collective behind the kernel that produced `part` and returns at once, so the warp of frame
    t+1 (current stream) overlaps the reduce-scatter of frame t (side stream); wait(ticket) returns the finished slab
    [B,C,rows,Wb] fp32 (channels-last in memory) and orders the current stream after it.
    """

    def __init__(self, views: int, bev_hw, channels: int, mode: str = "mean", group=None, device=None):
        if mode not in ("sum", "mean"):
            raise ValueError("reduce-scatter view sharding supports sum / mean")
        self.views, self.mode, self.group = views, mode, group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.Hb, self.Wb = bev_hw
        self.C = channels
        self.rows = slab_rows(self.Hb, self.world)
        self.device = device
        self.side = torch.cuda.Stream(device=device)
        self._div = None

    def submit(self, part: torch.Tensor):
        """part: this rank's partial sum [B,C,Hb,Wb] fp32, channels-last in memory (what ops.warp_fuse returns)."""
        B = part.shape[0]
        mem = part.permute(0, 2, 3, 1)                       # [B,Hb,Wb,C] as it lies in memory
        if not mem.is_contiguous():
            raise ValueError("partial BEV must be channels-last")
        pad = self.rows * self.world - self.Hb
        if pad:
            mem = torch.cat([mem, mem.new_zeros(B, pad, self.Wb, self.C)], dim=1)
        out = torch.empty((B, self.rows, self.Wb, self.C), device=part.device, dtype=torch.float32)
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(self.side):
            self.side.wait_event(ready)
            for b in range(B):                               # a frame's row slabs are contiguous: one collective per frame
                dist.reduce_scatter_tensor(out[b], mem[b], op=dist.ReduceOp.SUM, group=self.group)
            if self.mode == "mean":
                if self._div is None:
                    self._div = torch.tensor(float(self.views), device=part.device)
                out.div_(self._div)                          # tensor divisor: IEEE division (fusion.py:20-21 divides by V)
            done = torch.cuda.Event()
            done.record()
        mem.record_stream(self.side)
        return out, done

    def wait(self, ticket) -> torch.Tensor:
        out, done = ticket
        torch.cuda.current_stream().wait_event(done)
        out.record_stream(torch.cuda.current_stream())
        return out.permute(0, 3, 1, 2)                       # logical [B,C,rows,Wb]


class PeerSlabFusion:
    """Fused compute + exchange: every rank's warp kernel sends its partial sum straight to the owners of the BEV rows
    through peer memory (bevipm_warp_fuse_red on NVLink-mapped pointers from torch's symmetric memory), as kilobyte bulk
    copies out of shared memory, tile by tile while it warps: the partial BEV is never written to or re-read from local HBM.

      put=True  (default): every rank has a private receive buffer at each owner; the kernel STORES (cp.async.bulk), the
                owner then adds the buffers in rank order and divides (bevipm_slab_finish, on a side stream so it overlaps
                the next frame's warp).  No atomics, no zeroing, reproducible sums.
      put=False: all ranks ADD into one slab per owner (cp.reduce.async.bulk add.f32); the owner zeroes it between uses.

    Two buffer sets alternate; one cross-rank barrier per frame.
    submit(...) -> ticket, wait(ticket) -> this rank's rows [B,C,rows,Wb] fp32; run(...) = wait(submit(...)).
    """

    def __init__(self, views: int, bev_hw, channels: int, frames: int = 1, mode: str = "mean", group=None, device=None, put: bool = True):
        import torch.distributed._symmetric_memory as symm_mem
        if mode not in ("sum", "mean"):
            raise ValueError("peer-slab view sharding supports sum / mean")
        self.views, self.mode, self.put = views, mode, put
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.Hb, self.Wb = bev_hw
        self.C, self.B = channels, frames
        self.rows = slab_rows(self.Hb, self.world)
        self.slab_elems = frames * self.rows * self.Wb * channels
        self.nsrc = self.world if put else 1
        self.sources = [r for r, ids in enumerate(view_assignment(views, self.world)) if ids]   # ranks that hold cameras
        self.buf = symm_mem.empty((2, self.nsrc, frames, self.rows, self.Wb, channels), dtype=torch.float32, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.buf.zero_()
        self.turn = 0
        self.side = torch.cuda.Stream(device=device)
        self._last_finish = None
        self._div = torch.tensor(float(views), device=device)
        self.hdl.barrier()

    def bytes_over_nvlink_per_call(self) -> int:
        """fp32 bytes this rank sends into OTHER ranks' buffers per call (what crosses NVLink)."""
        own = min(self.rows, max(0, self.Hb - self.rank * self.rows))
        return self.B * (self.Hb - own) * self.Wb * self.C * 4

    def submit(self, feats_r, K_r, Rt_r, xs, ys, img_size):
        import ctypes
        from . import _lib, ops
        k = self.turn
        self.turn ^= 1
        main = torch.cuda.current_stream()
        L = _lib.load()
        # Buffer set k was last read (put: by the owners' finish kernels; add: zeroed) two calls ago: the barrier of the call
        # in between ordered that before any rank's writes of this call.
        if feats_r is not None and feats_r.shape[1] > 0:
            if not ops._is_channels_last5(feats_r):
                raise ValueError("PeerSlabFusion wants channels-last features")
            d = ops._fill_desc(feats_r.shape, feats_r.stride(), (self.rows * self.Wb * self.C, 0, 1, self.Wb * self.C, self.C),
                               (self.Hb, self.Wb), (int(img_size[0]), int(img_size[1])), _lib.SUM, ops._DT[feats_r.dtype], _lib.F32, 0,
                               _lib.FLAG_SLAB_PUT if self.put else 0)
            off = ((k * self.nsrc + (self.rank if self.put else 0)) * self.slab_elems) * 4
            arr = (ctypes.c_void_p * self.world)(*[p + off for p in self.ptrs])
            _lib.check(L.bevipm_warp_fuse_red(ctypes.byref(d), ops._ptr(feats_r), ops._ptr(K_r), ops._ptr(Rt_r), ops._ptr(xs),
                                              ops._ptr(ys), arr, self.world, self.rows, ctypes.c_void_p(main.cuda_stream)))
        if self._last_finish is not None:
            main.wait_event(self._last_finish)                # my previous finish has read its buffers before others may pass on
        self.hdl.barrier()                                    # every rank's partial sums for this frame have landed
        if not self.put:
            slab = self.buf[k, 0]
            out = slab.permute(0, 3, 1, 2)
            out = out / self._div if self.mode == "mean" else out.clone()
            slab.zero_()                                      # ready for its next turn (two calls from now)
            return out, None
        landed = torch.cuda.Event()
        landed.record(main)
        out = torch.empty((self.B, self.rows, self.Wb, self.C), device=self.buf.device, dtype=torch.float32)
        with torch.cuda.stream(self.side):
            self.side.wait_event(landed)
            base = self.buf.data_ptr() + k * self.nsrc * self.slab_elems * 4
            arr = (ctypes.c_void_p * len(self.sources))(*[base + r * self.slab_elems * 4 for r in self.sources])
            _lib.check(L.bevipm_slab_finish(arr, len(self.sources), ops._ptr(out), self.slab_elems,
                                            float(self.views) if self.mode == "mean" else 1.0, ctypes.c_void_p(self.side.cuda_stream)))
            done = torch.cuda.Event()
            done.record(self.side)
        out.record_stream(self.side)
        self._last_finish = done
        return out.permute(0, 3, 1, 2), done                  # logical [B,C,rows,Wb]

    def wait(self, ticket) -> torch.Tensor:
        out, done = ticket
        if done is not None:
            torch.cuda.current_stream().wait_event(done)
        return out

    def run(self, feats_r, K_r, Rt_r, xs, ys, img_size):
        return self.wait(self.submit(feats_r, K_r, Rt_r, xs, ys, img_size))


class CopyEngineSlabFusion:
    """Row-slab view sharding with the exchange on the copy engines: the ordinary fused SUM kernel writes this rank's
    partial BEV to local HBM, DMA copies (peer-to-peer over NVLink, a second stream) push every other owner's rows into
    this rank's receive buffer there, one cross-rank barrier, and the owner adds the buffers in rank order and divides
    (bevipm_slab_finish).  Copies and owner-side sums of frame t overlap the warp of frame t+1; nothing but the warp kernel
    runs on the SMs of the compute stream.  Same results as PeerSlabFusion(put=True), bit for bit.

    submit(part) -> ticket with part = this rank's partial sum [B,C,Hb,Wb] fp32 channels-last (ops.warp_fuse, mode SUM) or
    None for a rank without cameras; wait(ticket) -> this rank's rows [B,C,rows,Wb].
    """

    def __init__(self, views: int, bev_hw, channels: int, frames: int = 1, mode: str = "mean", group=None, device=None):
        import torch.distributed._symmetric_memory as symm_mem
        if mode not in ("sum", "mean"):
            raise ValueError("copy-engine view sharding supports sum / mean")
        self.views, self.mode = views, mode
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.Hb, self.Wb = bev_hw
        self.C, self.B = channels, frames
        self.rows = slab_rows(self.Hb, self.world)
        if frames != 1 and self.rows * self.world != self.Hb:
            raise ValueError("several frames per call need bev_h divisible by the number of ranks")
        self.slab_elems = frames * self.rows * self.Wb * channels
        self.sources = [r for r, ids in enumerate(view_assignment(views, self.world)) if ids]
        self.buf = symm_mem.empty((2, self.world, frames, self.rows, self.Wb, channels), dtype=torch.float32, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        self.buf.zero_()
        # peer views of MY receive slot at every owner, both buffer sets
        self.peer = [[self.hdl.get_buffer(q, (frames, self.rows, self.Wb, channels), torch.float32,
                                          (k * self.world + self.rank) * self.slab_elems) for q in range(self.world)] for k in range(2)]
        self.turn = 0
        self.side = torch.cuda.Stream(device=device)
        self._last_finish = None
        self.hdl.barrier()

    def bytes_over_nvlink_per_call(self) -> int:
        own = min(self.rows, max(0, self.Hb - self.rank * self.rows))
        return self.B * (self.Hb - own) * self.Wb * self.C * 4

    def submit(self, part):
        import ctypes
        from . import _lib, ops
        k = self.turn
        self.turn ^= 1
        main = torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(main)
        out = torch.empty((self.B, self.rows, self.Wb, self.C), device=self.buf.device, dtype=torch.float32)
        with torch.cuda.stream(self.side):
            self.side.wait_event(ready)
            if part is not None:
                mem = part.permute(0, 2, 3, 1)                 # [B,Hb,Wb,C] as it lies in memory
                for q in range(self.world):
                    lo, hi = q * self.rows, min(self.Hb, (q + 1) * self.rows)
                    if hi > lo:                                # (my own rows go through my own slot too: one code path)
                        self.peer[k][q][:, : hi - lo].copy_(mem[:, lo:hi], non_blocking=True)
                part.record_stream(self.side)
            self.hdl.barrier()                                 # every rank's copies for this frame have landed
            base = self.buf.data_ptr() + k * self.world * self.slab_elems * 4
            arr = (ctypes.c_void_p * len(self.sources))(*[base + r * self.slab_elems * 4 for r in self.sources])
            _lib.check(_lib.load().bevipm_slab_finish(arr, len(self.sources), ops._ptr(out), self.slab_elems,
                                                      float(self.views) if self.mode == "mean" else 1.0,
                                                      ctypes.c_void_p(self.side.cuda_stream)))
            done = torch.cuda.Event()
            done.record(self.side)
        out.record_stream(self.side)
        return out.permute(0, 3, 1, 2), done

    def wait(self, ticket) -> torch.Tensor:
        out, done = ticket
        torch.cuda.current_stream().wait_event(done)
        return out
