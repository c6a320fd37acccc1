"""`llama_type = "llama_b200"`: the B200-native inner model behind the reference's plugin seam.

The reference picks its inner LLM by module name (model/accessory/model/meta.py:30-32:
`importlib.import_module(f"accessory.model.LLM.{llama_type}")` and reads `ModelArgs`, `Transformer`
from it).  This module exports the same two names with the contract of SURVEY 8b, restating
LLM/llama_ens5.py narrowed to CLIP ViT-L/14 + visual_proj + Llama (single view, `image_size`
configurable), with every tensor op executed by the sm_100a kernels of liba3vlm_b200.so through
`a3vlm_b200.engine.Engine`.  Parameters keep the reference's names and shapes (they are views into
the engine's flat HBM buffers), so `M/util/tensor_parallel.py` loads / saves checkpoints unchanged.
There is no CPU path: constructing `Transformer` without a CUDA device raises.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import os

import torch
import torch.nn as nn

from ... import _lib
from ...checkpoint import holder_class_for
from ...engine import Engine

bf16 = torch.bfloat16


@dataclass
class ModelArgs:
    # --- same fields and defaults as LLM/llama_ens5.py:33-50
    dim: int = 5120
    n_layers: int = 40
    n_heads: int = 40
    n_kv_heads: Optional[int] = None
    vocab_size: int = -1  # defined later by tokenizer
    multiple_of: int = 256
    ffn_dim_multiplier: Optional[float] = None
    norm_eps: float = 1e-5
    rope_theta: float = 10000
    max_batch_size: int = 32
    max_seq_len: int = 2048
    rope_scaling: Optional[float] = None
    load_pretrained_visual_encoder: bool = False
    # --- vision geometry (the reference hard-codes open_clip ViT-L-14 at 224 px, llama_ens5.py:296-301,336)
    image_size: int = 336   # side of ONE ViT view
    image_views: int = 1    # 1: single view of image_size; 5: the reference's tiling of a (2*image_size)^2 image
    vit_patch: int = 14
    vit_width: int = 1024
    vit_layers: int = 24
    vit_heads: int = 16
    vit_mlp: int = 4096
    vit_quick_gelu: Optional[bool] = None  # None: QuickGELU iff OpenAI weights are loaded (llama_ens5.py:297-299)


@dataclass
class _Vision:
    image_size: int
    patch: int
    width: int
    layers: int
    heads: int
    mlp: int
    quick_gelu: bool
    views: int = 1
    ln_eps: float = 1e-5
    rgbd: bool = False  # second (depth) image stream with its own tags: LLM/llama_ens5_2images.py

    @property
    def grid(self):
        return self.image_size // self.patch

    @property
    def tokens(self):
        return self.grid * self.grid + 1


class _Holder(nn.Module):
    """Parameter container: only exists so that state-dict names match the reference's."""


class TransformerBlock(_Holder):
    """Unit the reference uses for FSDP wrapping / activation checkpointing
    (M/main_finetune.py:234,268-276 take `type(model.llma.layers[0])`)."""


def _register(root: nn.Module, dotted: str, tensor: torch.Tensor, requires_grad: bool) -> nn.Parameter:
    parts = dotted.split(".")
    mod = root
    for k, p in enumerate(parts[:-1]):
        child = mod._modules.get(p)
        if child is None:
            # the module that owns a model-parallel weight carries the fairscale layer's class name, so
            # the reference's checkpoint loader can resolve its split dimension (checkpoint.py)
            leaf_cls = holder_class_for(dotted) if k == len(parts) - 2 else None
            child = (leaf_cls or _Holder)()
            mod.add_module(p, child)
        mod = child
    prm = nn.Parameter(tensor, requires_grad=requires_grad)
    mod.register_parameter(parts[-1], prm)
    return prm


class _TrainFn(torch.autograd.Function):
    """Autograd boundary around the engine: forward runs the hand-written forward (optionally with
    the fused shifted cross-entropy), backward runs the hand-written backward and publishes the flat
    bf16 gradient views as `param.grad` (no per-parameter copies)."""

    @staticmethod
    def forward(ctx, model: "Transformer", examples, image, labels, anchor):
        eng = model.engine
        ws, logits = eng.forward_train(examples, image, labels)
        # ws.examples: the dense copy of the ids the forward used (a trimmed column view would be mis-strided)
        ctx.model, ctx.ws, ctx.examples, ctx.fused = model, ws, ws.examples, labels is not None
        ctx.generation = ws.generation
        if labels is not None:
            return ws.loss[0].clone()
        B, T = examples.shape
        return logits.view(B, T, -1).clone()  # the workspace is reused by the next call

    @staticmethod
    def backward(ctx, grad_out):
        model, ws = ctx.model, ctx.ws
        eng = model.engine
        if ws.generation != ctx.generation:
            # the engine keeps ONE activation workspace per shape: a later forward (another micro-batch, or a
            # no_grad evaluation through loss_forward) has overwritten what this backward would read
            raise RuntimeError(
                "llama_b200: backward() after another forward of the same shape -- the activation workspace is "
                "reused, so run forward -> backward per micro-batch (gradients accumulate across backward calls; "
                "this is the order of M/engine_finetune.py:44-68)")
        params = model._trainable
        accumulate = params[0].grad is not None  # zero_grad(set_to_none=True) -> overwrite
        if ctx.fused:
            scale = float(grad_out)  # d loss (e.g. 1/accum_iter, M/engine_finetune.py:60)
            dlogits = ws.logits
        else:
            scale = 1.0
            dlogits = grad_out.reshape(ws.B * ws.T, -1).to(bf16).contiguous()
        hook = finish = None
        if model.dp_group is not None:
            # the mean over the ranks is folded into every micro-step's gradient, also the ones accumulated
            # inside no_sync() (FSDP divides the whole accumulated gradient when it finally reduces it)
            scale = scale / model.dp_world
            if model.sync_grads:
                hook, finish = eng.allreduce_hook(model.dp_group)
        eng.backward(ws, ctx.examples, dlogits, accumulate=accumulate, on_bucket=hook, scale=scale)
        if finish is not None:
            finish()
        gv = eng.grad_views()
        for name, prm in model._trainable_named:
            prm.grad = gv[name]
        return None, None, None, None, None


class Transformer(nn.Module):
    _RGBD = False  # llama_b200_2images.Transformer flips this (the reference's llama_ens5_2images build)
    _FROZEN_TAGS = ("start_depth_img", "end_depth_img")

    def __init__(self, args: ModelArgs, with_visual: bool = False):
        super().__init__()
        if not torch.cuda.is_available():
            raise _lib.A3Error("llama_b200 needs a B200 (sm_100a) CUDA device; there is no CPU fallback")
        self.args = args
        self.vocab_size = args.vocab_size
        self.n_layers = args.n_layers
        dev = torch.empty(0).device  # honours default_tensor_type(device=...) like the reference build
        if dev.type != "cuda":
            dev = torch.device("cuda", torch.cuda.current_device())
        vis = None
        if with_visual:
            quick = args.vit_quick_gelu if args.vit_quick_gelu is not None else args.load_pretrained_visual_encoder
            if args.image_views not in (1, 5):
                raise ValueError("image_views must be 1 (single view) or 5 (reference tiling, llama_ens5.py:383-385)")
            vis = _Vision(args.image_size, args.vit_patch, args.vit_width, args.vit_layers, args.vit_heads,
                          args.vit_mlp, bool(quick), views=args.image_views, rgbd=self._RGBD)
        self.engine = Engine(args, vis, dev)
        # views x (start_img + N ViT tokens + end_img); the RGB-D build counts both streams in image_words
        # and one in visual_image_words (LLM/llama_ens5_2images.py:335-336)
        self.image_words = self.engine.image_words
        self.visual_image_words = self.engine.visual_image_words
        self.cache_image_words = 0
        if with_visual:
            # the size the data pipeline must deliver (M/main_finetune.py:295 reads model.llma.image_size)
            self.image_size = args.image_size * (2 if args.image_views == 5 else 1)
        self.layers = nn.ModuleList([TransformerBlock() for _ in range(args.n_layers)])
        views = self.engine.param_views()
        self._trainable_named: List = []
        for name, t in views.items():
            frozen = name.startswith("clip.")
            prm = _register(self, name, t, requires_grad=not frozen)
            # the depth tags are trainable by name (llama_ens5_2images.py:343-344) but the training forward
            # never reaches them (meta.py:251), so like in the reference their .grad stays None
            if not frozen and name not in self._FROZEN_TAGS:
                self._trainable_named.append((name, prm))
        self._trainable = [p for _, p in self._trainable_named]
        self._anchor = torch.zeros((), device=dev, requires_grad=True)
        self.dp_group, self.dp_world, self.sync_grads = None, 1, True
        self.reset_parameters()

    # ------------------------------------------------------------------ init / bookkeeping
    @torch.no_grad()
    def reset_parameters(self, seed: Optional[int] = None) -> None:
        """Random init of the same families the reference uses: kaiming_uniform_(a=sqrt(5)) for the
        decoder linears / embedding (llama_ens5.py:28), ones for norms, U(0,1) tags (:338-339),
        width^-0.5 N(0,1) for the ViT class / positional embeddings (open_clip)."""
        self.engine.wait_weights()
        g = torch.Generator(device=self.engine.device)
        if seed is not None:
            g.manual_seed(seed)
        for name, p in self.named_parameters():
            if name in ("start_img", "end_img") + self._FROZEN_TAGS:
                p.copy_(torch.rand(p.shape, device=p.device, generator=g))
            elif name.endswith("norm.weight") or ".ln_" in name and name.endswith(".weight") or name == "visual_proj.1.weight":
                p.fill_(1.0)
            elif name.endswith(".bias") or name.endswith("_bias"):
                p.zero_()
            elif name.endswith("class_embedding") or name.endswith("positional_embedding"):
                p.copy_(torch.randn(p.shape, device=p.device, generator=g) * (p.shape[-1] ** -0.5))
            else:
                fan_in = p[0].numel()
                bound = fan_in ** -0.5
                p.copy_((torch.rand(p.shape, device=p.device, generator=g) * 2 - 1) * bound)
        self.engine.weights_changed()

    def _check_alias(self) -> None:
        prm = self._trainable[0]
        if prm.data_ptr() != self.engine.w.flat.data_ptr() or prm.dtype != bf16:
            raise _lib.A3Error(
                "llama_b200 parameters must stay bf16 views of the engine's flat buffer: do not call "
                ".float()/.to(other device) on the module (fp32 masters live in a3vlm_b200.train.AdamW)")

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        self.engine.wait_weights()  # never copy underneath a weight update still running on a side stream
        out = super().load_state_dict(state_dict, strict=strict, assign=False)
        self._check_alias()
        self.engine.weights_changed()
        return out

    def _load_from_state_dict(self, *a, **k):  # pragma: no cover - containers hold the parameters
        return super()._load_from_state_dict(*a, **k)

    def get_trainable_params(self, pretrain_stage: bool = False) -> Dict[str, nn.Parameter]:
        """llama_ens5.py:342-349: everything except the frozen encoders."""
        return {n: p for n, p in self.named_parameters() if not n.startswith("clip.")}

    def get_basic_block_classes(self):
        return [TransformerBlock]

    def get_quant_blocklist(self) -> List[str]:
        return [n for n, _ in self.named_modules() if n.startswith("clip.") or n.startswith("visual_proj.")]

    def enable_data_parallel(self, group=None, shard_optimizer: bool = True) -> None:
        """Replica data parallelism: the gradients are exchanged bucket by bucket (one per decoder
        layer) on a side stream while the backward is still running, 1/world folded into the loss gradient.
        shard_optimizer (default): rank r owns the r-th 1/N of every chunk -- the exchange is a reduce-scatter,
        the fp32 AdamW state exists for the owned slices only, and the updated bf16 weights are all-gathered
        chunk by chunk behind the next forward (Engine.enable_sharded_optimizer).  False: plain all-reduce and a
        replicated optimiser."""
        import torch.distributed as dist

        self.dp_group = group if group is not None else dist.group.WORLD
        self.dp_world = dist.get_world_size(self.dp_group)
        if self.dp_world > 1:
            # the CTA-pair GEMM claims its tiles at run time, so pairs whose SMs also run collective CTAs take fewer
            # (267.6 vs 273.8 ms per step at 2 GPUs); the library reads the switch at its first GEMM launch
            os.environ.setdefault("A3_GEMM_DYNAMIC", "1")
        if shard_optimizer and self.dp_world > 1:
            self.engine.enable_sharded_optimizer(self.dp_group)

    def set_activation_recompute(self, level: int) -> None:
        """The reference's `--checkpointing` (M/main_finetune.py:268-276 wraps every TransformerBlock in an
        activation checkpoint).  0: keep all activations (default); 1: selective -- norm outputs and the gate
        output are rebuilt in the backward by row kernels; 2: per-block recompute as in the reference (only the
        block inputs are kept, the backward re-runs each block's forward).  Same gradients bit for bit."""
        if level not in (0, 1, 2):
            raise ValueError("activation recompute level must be 0, 1 or 2")
        self.engine.recompute = int(level)

    def state_dict(self, *args, **kwargs):
        self.engine.wait_weights()  # a pending weight exchange / pipelined update must land first
        return super().state_dict(*args, **kwargs)

    def named_parameters(self, *args, **kwargs):
        self.engine.wait_weights()  # whoever asks for the parameters sees the finished update
        return super().named_parameters(*args, **kwargs)

    # ------------------------------------------------------------------ reference-facing forwards
    def forward(self, examples: torch.Tensor, image: Optional[torch.Tensor] = None, depth_imgs=None):
        """LLM/llama_ens5.py:461-487: (B,T) int64 ids [+ (B,3,S,S) image] -> (B,T,V) bf16 logits of the
        text positions.  Differentiable w.r.t. every trainable parameter."""
        if depth_imgs is not None:
            # dead code in the reference: MetaModel.forward drops depth_imgs (M/model/meta.py:251), and
            # llama_ens5_2images.forward would return T + depth_words logit rows (:507 slices by the RGB
            # count only), which the loss at meta.py:256-262 cannot consume
            raise NotImplementedError("the training forward takes the RGB stream only (M/model/meta.py:251)")
        self._check_alias()
        self.engine.destroy_kv_cache()  # training always disables the kv cache (:462)
        examples = examples.to(self.engine.device)
        if image is not None:
            image = image.to(self.engine.device)
        if torch.is_grad_enabled():
            return _TrainFn.apply(self, examples, image, None, self._anchor)
        _, logits = self.engine.forward_train(examples, image)
        return logits.view(examples.shape[0], examples.shape[1], -1).clone()

    def loss_forward(self, examples: torch.Tensor, labels: torch.Tensor, image=None) -> torch.Tensor:
        """forward + the shifted CrossEntropyLoss(ignore_index=0) of M/model/meta.py:256-262 fused behind
        one autograd node, so the (B,T,V) logits never leave the workspace."""
        self._check_alias()
        self.engine.destroy_kv_cache()
        dev = self.engine.device
        examples, labels = examples.to(dev), labels.to(dev)
        if image is not None:
            image = image.to(dev)
        if torch.is_grad_enabled():
            return _TrainFn.apply(self, examples, image, labels, self._anchor)
        ws, _ = self.engine.forward_train(examples, image, labels, want_grad=False)
        return ws.loss[0].clone()

    @torch.inference_mode()
    def forward_inference(self, tokens: torch.Tensor, start_pos: int, image=None, depth_images=None) -> torch.Tensor:
        """LLM/llama_ens5.py:490-531: fp32 (B,V) logits of the last position; KV cache keyed on
        start_pos == 0; decode positions offset by cache_image_words (:520).  `depth_images` is used by the
        RGB-D build only (LLM/llama_ens5_2images.py:511-547)."""
        dev = self.engine.device
        tokens = tokens.to(dev)
        if image is not None:
            image = image.to(dev)
        if depth_images is not None:
            depth_images = depth_images.to(dev)
        out = self.engine.forward_inference(tokens, start_pos, image, depth_images)
        self.cache_image_words = self.engine.cache_image_words
        return out

    def _allocate_kv_cache(self, max_batch_size: int) -> None:
        self.engine.allocate_kv_cache(max_batch_size)

    def _destroy_kv_cache(self) -> None:
        self.engine.destroy_kv_cache()
