"""Replica data parallelism for the fine-tune step: one process per GPU, identical
replicas, ONE sum all-reduce of the flat bf16 gradient buffer per optimiser step, issued bucket by
bucket as the backward finishes each decoder layer so it overlaps the remaining backward.  The
reference shards with FSDP (M/main_finetune.py:241-263: all-gather + reduce-scatter per block); with
13.5 GB of bf16 weights per replica on a 180 GB part no sharding is needed and the only exchange left
is the gradient sum.  The 1/world mean is folded into the loss gradient by the caller, so the
collective is a plain SUM.  Global-norm clipping happens after the all-reduce on identical
replicas -> no further collective (the reference needs two, M/util/clip_grad.py:174-181).

On CUDA the collective runs on a side stream behind an event recorded on the compute stream; on CPU
tensors (gloo, used by the world_size-2 tests) it falls back to async work handles."""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


class BucketAllReduce:
    def __init__(self, flat: torch.Tensor, group=None, comm_stream: Optional["torch.cuda.Stream"] = None):
        self.flat, self.group, self.cs = flat, group, comm_stream
        self.handles: List = []
        self.covered: List = []

    def on_bucket(self, lo: int, hi: int) -> None:
        if hi <= lo:
            return
        self.covered.append((lo, hi))
        buf = self.flat[lo:hi]
        if self.flat.is_cuda:
            ev = torch.cuda.Event()
            ev.record()
            self.cs.wait_event(ev)
            with torch.cuda.stream(self.cs):
                dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
        else:
            self.handles.append(dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self) -> None:
        if self.flat.is_cuda:
            torch.cuda.current_stream().wait_stream(self.cs)
        for h in self.handles:
            h.wait()
        self.handles = []


class BucketReduceScatter:
    """Gradient exchange for the sharded optimiser (Engine.enable_sharded_optimizer): every decay bucket
    [lo, hi) is REDUCE-SCATTERED in place -- rank r ends up with the summed gradients of its slice
    [lo + r*n, lo + (r+1)*n), n = (hi - lo) / world, and nothing else of the bucket is valid afterwards --
    while the small no-decay tail [decay_numel, ...) is all-reduced (its update is replicated).  Half the
    bytes of BucketAllReduce during the backward; the other half of an all-reduce (the all-gather) moves
    the updated WEIGHTS after the optimiser step instead of the gradients.

    CPU tensors (gloo, used by the world_size-2 tests): gloo has no reduce-scatter, so the bucket is
    all-reduced, which leaves the owned slice with the same values."""

    def __init__(self, flat: torch.Tensor, decay_numel: int, world: int, rank: int, group=None,
                 comm_stream: Optional["torch.cuda.Stream"] = None):
        self.flat, self.decay_numel, self.world, self.rank = flat, decay_numel, world, rank
        self.group, self.cs = group, comm_stream
        self.handles: List = []
        self.covered: List = []

    def owned(self, lo: int, hi: int):
        n = (hi - lo) // self.world
        return lo + self.rank * n, lo + (self.rank + 1) * n

    def _exchange(self, lo: int, hi: int, async_op: bool):
        buf = self.flat[lo:hi]
        if lo >= self.decay_numel or not self.flat.is_cuda:
            return dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
        if (hi - lo) % self.world:
            raise ValueError(f"bucket [{lo}, {hi}) does not split over {self.world} ranks")
        a, b = self.owned(lo, hi)
        return dist.reduce_scatter_tensor(self.flat[a:b], buf, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)

    def on_bucket(self, lo: int, hi: int) -> None:
        if hi <= lo:
            return
        if lo < self.decay_numel < hi:
            raise ValueError("a bucket must not straddle the decay / no-decay boundary")
        self.covered.append((lo, hi))
        if self.flat.is_cuda:
            ev = torch.cuda.Event()
            ev.record()
            self.cs.wait_event(ev)
            with torch.cuda.stream(self.cs):
                self._exchange(lo, hi, False)
        else:
            self.handles.append(self._exchange(lo, hi, True))

    def finish(self) -> None:
        if self.flat.is_cuda:
            torch.cuda.current_stream().wait_stream(self.cs)
        for h in self.handles:
            h.wait()
        self.handles = []
