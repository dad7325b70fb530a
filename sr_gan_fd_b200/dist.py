"""Data-parallel training of the generator: one process per GPU, one logical all-reduce of the flat gradient buffer.
The backward pass announces gradient buckets as they become ready; the reducer either all-reduces them right away on a
side stream (overlapped with the remaining kernels) or -- the default, faster on B200 -- coalesces them into a single
all-reduce at the end of backward.

The reference has no distributed code; this is the new capability BASELINE.json config 3 asks for.
The generator has no BatchNorm / dropout / buffers, so averaging gradients over ranks is mathematically identical to
the reference's mean-reduced L1 over the global batch (``nn.L1Loss()`` default, ``ESRGAN/train_rrdbnet.py:188-190``).

Mechanics: ``b200sr_backward`` announces (host callback) each contiguous range of the flat gradient buffer whose
producing kernels have all been enqueued -- tail convs first, then one bucket per RRDB in reverse order, conv1 last.
For every bucket we record an event on the compute stream, make the communication stream wait on it and enqueue an
averaging all-reduce there; when backward has enqueued everything the compute stream waits for the outstanding
collectives.  No host synchronisation anywhere.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist

from .function import GeneratorRuntime


class GradBucketReducer:
    """Averages gradient buckets over the process group as they become ready.

    Works on CUDA tensors (NCCL, side stream) and on CPU tensors (gloo; used by the world_size-2 CPU tests)."""

    def __init__(self, process_group=None, average: bool = True, min_bucket_numel: int = 0) -> None:
        self.group = process_group
        self.average = average
        self.min_bucket_numel = min_bucket_numel
        self.world_size = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self._works: List = []
        self._comm_stream: Optional[torch.cuda.Stream] = None
        self._pending = None  # (flat, off, cnt) coalescing of small adjacent buckets
        self.buckets_seen: List[tuple] = []

    # ---- called from inside backward -------------------------------------------------------------------------
    def bucket_ready(self, flat: torch.Tensor, offset: int, count: int) -> None:
        self.buckets_seen.append((offset, count))
        if self.world_size == 1:
            return
        if self._pending is not None:
            p_flat, p_off, p_cnt = self._pending
            if p_flat is flat and offset + count == p_off:      # buckets arrive in descending address order
                offset, count = offset, count + p_cnt
                self._pending = None
            elif p_flat is flat and p_off + p_cnt == offset:
                offset, count = p_off, p_cnt + count
                self._pending = None
            else:
                self._launch(*self._pending)
                self._pending = None
        if count < self.min_bucket_numel:
            self._pending = (flat, offset, count)
            return
        self._launch(flat, offset, count)

    def _launch(self, flat: torch.Tensor, offset: int, count: int) -> None:
        view = flat[offset:offset + count]
        op = dist.ReduceOp.AVG if (self.average and flat.is_cuda) else dist.ReduceOp.SUM
        if flat.is_cuda:
            if self._comm_stream is None:
                self._comm_stream = torch.cuda.Stream(device=flat.device)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(flat.device))
            self._comm_stream.wait_event(ev)
            with torch.cuda.stream(self._comm_stream):
                work = dist.all_reduce(view, op=op, group=self.group, async_op=True)
            self._works.append((work, None))
        else:
            work = dist.all_reduce(view, op=op, group=self.group, async_op=True)
            self._works.append((work, view if self.average else None))

    def reduce_all(self, flat: torch.Tensor) -> None:
        """Coalesced mode: ONE all-reduce of the whole flat gradient buffer once backward has enqueued everything.  No
        bucket callback is installed, so the backward pass keeps its merged gradient unpack and two-stream wgrad overlap."""
        self.buckets_seen.append((0, flat.numel()))
        if self.world_size == 1:
            return
        self._launch(flat, 0, flat.numel())
        self.finish(flat)

    def finish(self, flat: Optional[torch.Tensor] = None) -> None:
        """Make the compute stream wait for every outstanding bucket (called once backward has enqueued all work)."""
        if self._pending is not None:
            self._launch(*self._pending)
            self._pending = None
        for work, cpu_view in self._works:
            work.wait()
            if cpu_view is not None:
                cpu_view.div_(self.world_size)
        if self._comm_stream is not None and flat is not None and flat.is_cuda:
            torch.cuda.current_stream(flat.device).wait_stream(self._comm_stream)
            flat.record_stream(self._comm_stream)
        self._works.clear()


def broadcast_parameters(module: torch.nn.Module, src: int = 0, process_group=None) -> None:
    """Identical replicas at start (the alternative to seeding every rank the same way)."""
    if not dist.is_initialized() or dist.get_world_size(process_group) == 1:
        return
    for p in module.parameters():
        dist.broadcast(p.data, src=src, group=process_group)


def make_data_parallel(module, process_group=None, min_bucket_numel: Optional[int] = None) -> GradBucketReducer:
    """Attach gradient averaging to a drop-in generator.  Returns the reducer (for inspection).

    ``min_bucket_numel=None`` (default): adjacent buckets are coalesced into ONE all-reduce issued when backward has
    enqueued everything.  Measured on B200 this beats overlapping: the backward pass fills every SM with one-CTA-per-SM
    kernels, so NCCL kernels launched in the middle of it only take SMs away from the weight-gradient launches
    (8 GPUs: 8730 vs 8537 img/s; 2 GPUs: 2193 vs 2146).  Pass a number (elements) to all-reduce buckets of at least
    that size as soon as they are ready instead (0 = every bucket the backward pass announces)."""
    import os
    if "B200SR_DP_MIN_BUCKET" in os.environ:  # experiment switch: all-reduce buckets of at least this many elements early
        min_bucket_numel = int(float(os.environ["B200SR_DP_MIN_BUCKET"]))
    rt: GeneratorRuntime = module._runtime()
    if min_bucket_numel is None:
        reducer = GradBucketReducer(process_group, average=True, min_bucket_numel=1 << 62)
        rt.grad_bucket_hook = None          # backward is not told about buckets: no per-bucket unpack launches / stream joins
        rt.grad_done_hook = reducer.reduce_all
        return reducer
    reducer = GradBucketReducer(process_group, average=True, min_bucket_numel=min_bucket_numel)
    rt.grad_bucket_hook = reducer.bucket_ready
    rt.grad_done_hook = reducer.finish
    return reducer
