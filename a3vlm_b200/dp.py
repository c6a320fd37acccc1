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
