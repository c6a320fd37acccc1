"""Fine-tune step plumbing with the reference's call shapes (model/accessory/engine_finetune.py:44-79,
model/accessory/util/misc.py:296-321): an optimiser object and a loss scaler that drive the engine's
fused all-reduce / global-norm clip / AdamW kernels instead of FSDP + apex FusedAdam."""
from __future__ import annotations

import contextlib
import math
from typing import Optional

import torch


class AdamW(torch.optim.Optimizer):
    """torch.optim.AdamW-shaped front end (M/main_finetune.py:284-285: AdamW(param_groups, lr,
    betas=(0.9, 0.95))) over the engine's flat fp32 master / m / v buffers.  The decay / no-decay split
    of M/util/misc.py:586-599 is built into the flat layout; `weight_decay` applies to the decay
    region only.  `step(max_norm=...)` fuses the global-L2 clip of M/util/clip_grad.py."""

    def __init__(self, model, lr: float = 1e-3, betas=(0.9, 0.95), eps: float = 1e-8, weight_decay: float = 0.0,
                 pipelined: Optional[bool] = None):
        """pipelined=True: the update (norm, clip, AdamW) runs on a side stream, chunk by chunk in the order
        the forward consumes the weights, underneath the NEXT step's forward (Engine.overlap_optimizer;
        with data parallelism the sharded optimiser always runs this way and all-gathers each chunk after it).
        Same arithmetic, bit for bit.  The model's own entry points (forward, forward_inference,
        state_dict, parameters / named_parameters, load_state_dict, checkpoint save / load) wait for the
        chunks they touch; code that reads a parameter tensor it kept a handle to must call
        `model.llma.engine.wait_weights()` first.  None leaves the engine's setting unchanged."""
        llma = getattr(model, "llma", model)
        self.engine = llma.engine
        if pipelined is not None:
            self.engine.overlap_optimizer = bool(pipelined)
        params = [p for p in llma.parameters() if p.requires_grad]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None, max_norm: Optional[float] = None):
        g = self.param_groups[0]
        self.last_norm_sq = self.engine.optimizer_step(g["lr"], g["weight_decay"], max_norm or 0.0, g["betas"], g["eps"])
        return None

    def grad_norm(self) -> torch.Tensor:
        """Pre-clip global gradient norm of the last step (device tensor; reading it syncs)."""
        return self.engine.grad_norm_sq().sqrt()


class NativeScalerWithGradNormCount:
    """M/util/misc.py:296-321 for bf16 (no loss scaling needed -- the reference's ShardedGradScaler is
    only enabled for fp16, :299-300): backward, inside no_sync() on accumulation micro-steps, then
    clip + step."""

    state_dict_key = "amp_scaler"

    def __init__(self, args=None):
        pass

    def __call__(self, loss, optimizer, model, clip_grad=None, parameters=None, create_graph=False, update_grad=True):
        """Same positional order as the reference (misc.py:302).  `clip_grad=None` disables clipping (the
        reference clips at 1e8) but the pre-clip norm is still returned; `parameters` is accepted and unused,
        as there."""
        if create_graph:
            raise NotImplementedError("second-order gradients are not built (the fine-tune loop never asks for them)")
        llma = getattr(model, "llma", model)
        if update_grad:
            loss.backward()
            optimizer.step(max_norm=clip_grad)
            return optimizer.grad_norm()
        with no_sync(llma):
            loss.backward()
        return None

    def state_dict(self):
        return {}

    def load_state_dict(self, state_dict):
        pass


@contextlib.contextmanager
def no_sync(llma):
    """FSDP.no_sync() equivalent (M/util/misc.py:311-313): skip the gradient all-reduce."""
    prev = llma.sync_grads
    llma.sync_grads = False
    try:
        yield
    finally:
        llma.sync_grads = prev


def adjust_learning_rate(optimizer, it, args) -> float:
    """M/util/lr_sched.py:3-20 (per-iteration schedule used by engine_finetune.py:37-41): linear warm-up over
    `args.warmup_iters`, half-cycle cosine from `args.lr` down to `args.min_lr` until `args.lr_decay_iters`,
    `min_lr` afterwards; honours a per-group `lr_scale`.  Host arithmetic, kept for standalone use."""
    if it < args.warmup_iters:
        lr = args.lr * it / args.warmup_iters
    elif it > args.lr_decay_iters:
        lr = args.min_lr
    else:
        ratio = (it - args.warmup_iters) / (args.lr_decay_iters - args.warmup_iters)
        assert 0 <= ratio <= 1
        lr = args.min_lr + (args.lr - args.min_lr) * 0.5 * (1.0 + math.cos(math.pi * ratio))
    for group in optimizer.param_groups:
        group["lr"] = lr * group["lr_scale"] if "lr_scale" in group else lr
    return lr


def adjust_learning_rate_epoch(optimizer, epoch, args) -> float:
    """M/util/lr_sched.py:23-35: the same shape over fractional epochs (warmup_epochs, epochs)."""
    if epoch < args.warmup_epochs:
        lr = args.lr * epoch / args.warmup_epochs
    else:
        lr = args.min_lr + (args.lr - args.min_lr) * 0.5 * \
            (1.0 + math.cos(math.pi * (epoch - args.warmup_epochs) / (args.epochs - args.warmup_epochs)))
    for group in optimizer.param_groups:
        group["lr"] = lr * group["lr_scale"] if "lr_scale" in group else lr
    return lr
