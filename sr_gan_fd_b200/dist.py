"""Data-parallel training of the generator: one process per GPU, one logical all-reduce of the flat gradient buffer.
Default: ONE all-reduce of the whole buffer when backward has enqueued its last kernel (no bucket callback is installed, so
backward keeps its merged gradient unpack).  Optional (``overlap=True``): the backward pass announces gradient buckets as they
become ready and the reducer all-reduces them right away on a few-CTA, high-priority NCCL communicator.

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


def _overlap_group(process_group=None, max_ctas: int = 4):
    """A dedicated NCCL communicator for the gradient buckets: few CTAs (the buckets only need ~30 GB/s to hide behind the
    weight-gradient kernels, and every SM NCCL holds is one the backward pass loses) on a HIGH-PRIORITY stream, so that its
    blocks are scheduled as soon as running weight-gradient CTAs retire instead of queueing behind the next launch."""
    try:
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        opts.config.max_ctas = max_ctas
        opts.config.min_ctas = 1
        ranks = dist.get_process_group_ranks(process_group) if process_group is not None else list(range(dist.get_world_size()))
        return dist.new_group(ranks=ranks, backend="nccl", pg_options=opts)
    except Exception:
        return process_group


def make_data_parallel(module, process_group=None, min_bucket_numel: Optional[int] = None,
                       overlap: Optional[bool] = None) -> GradBucketReducer:
    """Attach gradient averaging to a drop-in generator.  Returns the reducer (for inspection).

    ``overlap=False``: ONE all-reduce of the whole flat gradient buffer, issued when backward has enqueued everything (no
    bucket callback: backward keeps its merged gradient unpack).  ``overlap=True``: backward announces buckets of
    ``B200SR_DP_BUCKET_RRDBS`` (default 6) RRDBs as their weight-gradient kernels are enqueued; each is all-reduced right away
    on a few-CTA, high-priority NCCL communicator while the remaining weight-gradient kernels run, and only the last small
    bucket (two RRDBs + conv1) is left exposed.  Default: ``B200SR_DP_OVERLAP`` (0: coalesced -- measured on 2 x B200 the single
    all-reduce costs 0.12 ms per step over one GPU, the overlapped buckets 0.21 ms: their NCCL kernels take SMs from the
    weight-gradient launches and backward needs one unpack launch + stream join per bucket).  ``min_bucket_numel``
    (elements) additionally coalesces adjacent announced buckets below that size."""
    import os
    if overlap is None:
        overlap = os.environ.get("B200SR_DP_OVERLAP", "0") == "1"
    if "B200SR_DP_MIN_BUCKET" in os.environ:  # experiment switch
        min_bucket_numel = int(float(os.environ["B200SR_DP_MIN_BUCKET"]))
    rt: GeneratorRuntime = module._runtime()
    cuda = next(module.parameters()).is_cuda
    if not overlap:
        reducer = GradBucketReducer(process_group, average=True, min_bucket_numel=1 << 62)
        rt.grad_bucket_hook = None          # backward is not told about buckets: no per-bucket unpack launches / stream joins
        rt.grad_bucket_rrdbs = 0
        rt.grad_done_hook = reducer.reduce_all
        return reducer
    group = process_group
    if cuda and dist.is_initialized() and dist.get_world_size(process_group) > 1:
        group = _overlap_group(process_group, int(os.environ.get("B200SR_DP_MAX_CTAS", "4")))
    reducer = GradBucketReducer(group, average=True, min_bucket_numel=min_bucket_numel or 0)
    rt.grad_bucket_rrdbs = int(os.environ.get("B200SR_DP_BUCKET_RRDBS", "6"))
    rt.grad_bucket_hook = reducer.bucket_ready
    rt.grad_done_hook = reducer.finish
    return reducer
