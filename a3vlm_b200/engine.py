"""Execution engine of the hot path: owns the flat weight / gradient / optimiser buffers in HBM and
the activation workspaces, and sequences the sm_100a kernels (through `ops`) for

  * the frozen CLIP ViT tower + visual_proj      (reference LLM/llama_ens5.py:351-375,450-452)
  * the Llama decoder training forward + loss    (LLM/llama_ens5.py:461-487, M/model/meta.py:251-262)
  * the hand-written backward of all of it       (autograd of the above in the reference)
  * grad all-reduce + global-norm clip + AdamW   (M/util/misc.py:302-315, M/util/clip_grad.py:187-193,
                                                  M/main_finetune.py:284-285)
  * prefill / decode with a KV cache             (LLM/llama_ens5.py:490-531)

No autograd graph is built: every activation the backward needs is kept in a pre-sized workspace and
the backward is an explicit kernel sequence.  torch is used for memory, streams and the NCCL
process group only.  Layout in HBM: one bf16 flat buffer for the trainable
weights with wq|wk|wv and w1|w3 of a layer adjacent (so the fused QKV / gate-up GEMMs read one
[N][K] matrix while the reference's parameter names stay views of it), one bf16 flat gradient buffer
of the same layout (bucketed per layer for the all-reduce), fp32 master / m / v flats for AdamW.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional, Tuple

import torch

from . import ops

bf16 = torch.bfloat16


def ffn_hidden_dim(dim: int, multiple_of: int, ffn_dim_multiplier) -> int:
    """LLM/llama_ens5.py:196-200 with hidden_dim = 4 * dim (:228)."""
    hidden = int(2 * (4 * dim) / 3)
    if ffn_dim_multiplier is not None:
        hidden = int(ffn_dim_multiplier * hidden)
    return multiple_of * ((hidden + multiple_of - 1) // multiple_of)


def build_layout(args, vis):
    """Pure function (no device): the ordered (name, shape) lists of the flat buffers.
    Returns (decay, nodecay, frozen, kpad).  Trainable order = decay + nodecay:
      visual_proj.{0,1}.weight, start/end_img | tok_embeddings | per layer: wqkv, wo, w13, w2 | output ||
      per layer: attention_norm, ffn_norm | norm | visual_proj.{0,1}.bias
    i.e. in the order the forward consumes them (the pipelined / sharded optimiser updates and all-gathers
    the chunks in that order while the next forward is already running).  wq|wk|wv and w1|w3 are adjacent rows of one
    matrix, so the reference-named parameters are views.  `||` = decay / no-decay boundary."""
    D, V = args.dim, args.vocab_size
    F = ffn_hidden_dim(D, args.multiple_of, args.ffn_dim_multiplier)
    n_kv = args.n_heads if getattr(args, "n_kv_heads", None) is None else args.n_kv_heads
    Dkv = n_kv * (D // args.n_heads)  # grouped-query attention: wk / wv have n_kv_heads heads (llama_ens5.py:76-90)
    decay: List[Tuple[str, Tuple[int, ...]]] = []
    nodecay: List[Tuple[str, Tuple[int, ...]]] = []
    if vis is not None:
        # visual_proj.1 is a LayerNorm, but its weight's NAME does not end in "norm.weight", so the
        # reference's add_weight_decay (M/util/misc.py:591) puts it in the decay group -- kept faithfully
        decay += [("visual_proj.0.weight", (D, vis.width)), ("visual_proj.1.weight", (D,)),
                  ("start_img", (1, 1, D)), ("end_img", (1, 1, D))]
    decay.append(("tok_embeddings.weight", (V, D)))
    for i in range(args.n_layers):
        p = f"layers.{i}."
        decay += [(p + "wqkv", (D + 2 * Dkv, D)), (p + "attention.wo.weight", (D, D)),
                  (p + "w13", (2 * F, D)), (p + "feed_forward.w2.weight", (D, F))]
        nodecay += [(p + "attention_norm.weight", (D,)), (p + "ffn_norm.weight", (D,))]
    decay.append(("output.weight", (V, D)))
    nodecay.append(("norm.weight", (D,)))
    frozen: List[Tuple[str, Tuple[int, ...]]] = []
    kpad = 0
    if vis is not None:
        v, c = vis, "clip.visual."
        Wv = v.width
        # conv1 re-laid for the im2col-free patch embedding: every (c) window zero-padded to 16 x 16 -> K = 768
        kpad = 3 * 16 * 16
        nodecay += [("visual_proj.0.bias", (D,)), ("visual_proj.1.bias", (D,))]
        frozen += [(c + "conv1.weight", (Wv, 3, v.patch, v.patch)), (c + "conv1_padded", (Wv, kpad)),
                   (c + "class_embedding", (Wv,)), (c + "positional_embedding", (v.tokens, Wv)),
                   (c + "ln_pre.weight", (Wv,)), (c + "ln_pre.bias", (Wv,)),
                   (c + "ln_post.weight", (Wv,)), (c + "ln_post.bias", (Wv,))]
        for i in range(v.layers):
            r = f"{c}transformer.resblocks.{i}."
            frozen += [(r + "ln_1.weight", (Wv,)), (r + "ln_1.bias", (Wv,)),
                       (r + "attn.in_proj_weight", (3 * Wv, Wv)), (r + "attn.in_proj_bias", (3 * Wv,)),
                       (r + "attn.out_proj.weight", (Wv, Wv)), (r + "attn.out_proj.bias", (Wv,)),
                       (r + "ln_2.weight", (Wv,)), (r + "ln_2.bias", (Wv,)),
                       (r + "mlp.c_fc.weight", (v.mlp, Wv)), (r + "mlp.c_fc.bias", (v.mlp,)),
                       (r + "mlp.c_proj.weight", (Wv, v.mlp)), (r + "mlp.c_proj.bias", (Wv,))]
        if getattr(v, "rgbd", False):
            # LLM/llama_ens5_2images.py:343-344.  They live with the frozen tensors: the reference's
            # training forward never sees the depth stream (M/model/meta.py:251), so these parameters
            # never receive a gradient and its optimiser skips them -- no master / m / v needed.
            frozen += [("start_depth_img", (1, 1, D)), ("end_depth_img", (1, 1, D))]
    return decay, nodecay, frozen, kpad


SHARD_ALIGN = 512  # 8 ranks x 64 elements: the decay region splits evenly over 1 / 2 / 4 / 8 ranks


def flat_offsets(layout, align: int = 64, segment_after: Optional[str] = None):
    """name -> (offset, numel, shape) with every tensor starting on a 128-byte boundary; total.
    After `segment_after` (the last decay parameter) the offset is rounded up to SHARD_ALIGN so the
    decay region can be sharded evenly across data-parallel ranks."""
    offsets: Dict[str, Tuple[int, int, Tuple[int, ...]]] = {}
    off = 0
    for name, shape in layout:
        n = int(math.prod(shape))
        offsets[name] = (off, n, tuple(shape))
        off += (n + align - 1) // align * align
        if name == segment_after:
            off = (off + SHARD_ALIGN - 1) // SHARD_ALIGN * SHARD_ALIGN
    return offsets, off


def bucket_plan(offsets, numel: int, n_layers: int, decay_numel: Optional[int] = None):
    """Flat element ranges in the order the backward completes them (the gradient-exchange buckets):
    LM head, decoder layers last to first, then projector + tags + embedding (the head of the buffer),
    then everything after the LM head (all norm weights, projector biases).  Together they tile
    [0, numel) exactly once.  With `decay_numel` the LM-head bucket runs to the end of the decay region
    (alignment padding included), so the decay buckets are exactly optimizer_chunks() in reverse."""
    rng = lambda n: (offsets[n][0], offsets[n][0] + offsets[n][1])
    head_hi = rng("output.weight")[1] if decay_numel is None else decay_numel
    plan = [(rng("output.weight")[0], head_hi)]
    prev_lo = rng("output.weight")[0]
    for i in reversed(range(n_layers)):
        lo = rng(f"layers.{i}.wqkv")[0]
        plan.append((lo, prev_lo))  # up to the next bucket's start: alignment padding is covered once
        prev_lo = lo
    plan.append((0, prev_lo))
    plan.append((head_hi, numel))
    return plan


def optimizer_chunks(offsets, decay_numel: int, n_layers: int):
    """Flat ranges of the decay region in the order the forward consumes them: projector + tags +
    embedding, one per decoder layer, LM head.  They tile [0, decay_numel) exactly once (the pipelined
    AdamW updates one chunk per launch and publishes an event after each)."""
    cuts = [0] + [offsets[f"layers.{i}.wqkv"][0] for i in range(n_layers)] + [offsets["output.weight"][0], decay_numel]
    return [(lo, hi) for lo, hi in zip(cuts[:-1], cuts[1:]) if hi > lo]


def owned_ranges(chunks, decay_numel: int, numel: int, world: int = 1, rank: int = 0):
    """[(lo, hi, decayed)]: rank's 1/world slice of every decay chunk, in order, then the no-decay tail
    (replicated).  For world = 1 the chunks themselves."""
    out = []
    for lo, hi in chunks:
        n = (hi - lo) // world
        out.append((lo + rank * n, lo + (rank + 1) * n, True))
    out.append((decay_numel, numel, False))
    return out


class FlatParams:
    """Flat bf16 storage for a set of named parameters.  `layout` is an ordered list of
    (name, shape); views[name] aliases the flat buffer."""

    def __init__(self, layout: List[Tuple[str, Tuple[int, ...]]], device, align: int = 64,
                 segment_after: Optional[str] = None):
        self.offsets, off = flat_offsets(layout, align, segment_after)
        self.numel = off
        self.flat = torch.zeros(max(off, align), device=device, dtype=bf16)
        self.views = {n: self.flat[o:o + k].view(s) for n, (o, k, s) in self.offsets.items()}

    def like(self, dtype) -> torch.Tensor:
        return torch.zeros(self.flat.numel(), device=self.flat.device, dtype=dtype)


class Workspace:
    """Activation + gradient scratch for one (B, T, W) shape.  Sized once, reused every step."""

    def __init__(self, eng: "Engine", B: int, T: int, W: int, train: bool):
        a, dev = eng.args, eng.device
        D, F, V, NL = a.dim, eng.ffn_dim, a.vocab_size, a.n_layers
        L = T + W
        M = B * L
        self.B, self.T, self.W, self.L, self.M = B, T, W, L, M
        e = lambda *s, dt=bf16: torch.empty(*s, device=dev, dtype=dt)
        keep = NL if train else 1
        # Activation recompute (the reference checkpoints every TransformerBlock, M/main_finetune.py:268-276):
        #   0  keep everything the backward reads (fastest; 7B at batch 8 needs ~28 GB)
        #   1  selective: the norm outputs xn / hn (+ rstd) and the gate output y are rebuilt in the backward
        #      from x / h / gu by three HBM-bound row kernels per layer (-25 % activation memory)
        #   2  full block recompute as in the reference: only the block inputs x[i] are kept, the backward
        #      re-runs the block's forward (without its last GEMM) into one layer's worth of scratch
        self.recompute = rc = (eng.recompute if train else 0)
        self.generation = 0  # bumped by every forward_train into this workspace (stale-backward guard)
        k_sel = 1 if rc >= 1 else keep   # xn, hn, y, rstd
        k_blk = 1 if rc >= 2 else keep   # qkv, o, h, gu, lse
        self.x = [e(M, D) for _ in range(keep + 1)]
        self.xn = [e(M, D) for _ in range(k_sel)]
        self.qkv = [e(M, 3 * D) for _ in range(k_blk)]
        self.o = [e(M, D) for _ in range(k_blk)]
        self.h = [e(M, D) for _ in range(k_blk)]
        self.hn = [e(M, D) for _ in range(k_sel)]
        self.gu = [e(M, 2 * F) for _ in range(k_blk)]
        # swiglu output: kept per layer when training without recompute (the w2 wgrad needs it), one
        # scratch buffer otherwise
        self.y = [e(M, F) for _ in range(k_sel)]
        self.lse = [e(B, a.n_heads, L, dt=torch.float32) for _ in range(k_blk)]
        self.rstd1 = [e(M, dt=torch.float32) for _ in range(k_sel)]
        self.rstd2 = [e(M, dt=torch.float32) for _ in range(k_sel)]
        if train:
            R = B * T
            self.fn = e(R, D)
            self.rstd_f = e(R, dt=torch.float32)
            self.logits = e(R, V)
            self.shift = e(R, dt=torch.int64)
            self.count = torch.zeros(1, device=dev, dtype=torch.int32)
            self.row_loss = e(R, dt=torch.float32)
            self.loss = torch.zeros(1, device=dev, dtype=torch.float32)
            self.dfn = e(R, D)
            self.d = e(M, D)        # running gradient of the residual stream
            self.dtmp = e(M, D)     # dgrad outputs into the norms / attention
            self.dy = e(M, F)
            self.dgu = e(M, 2 * F)
            self.dqkv = e(M, 3 * D)
            self.dq_acc = e(M, D, dt=torch.float32)
            self.delta = e(B, a.n_heads, L, dt=torch.float32)
            self.nparts = 2 * eng.n_sms  # rmsnorm_bwd runs 2 CTAs/SM: one resident wave
            self.part2 = e(self.nparts, D, dt=torch.float32)
            self.tag_grad = e(2, D, dt=torch.float32)
        if W:
            streams = W // eng.visual_image_words  # 2 when the depth stream is spliced in as well
            self.vit = eng._vit_workspace(B * eng.views * streams, train)
            if eng.views > 1:
                S = eng.vis.image_size
                self.vit["views"] = torch.empty(streams * eng.views * B, 3, S, S, device=dev, dtype=torch.float32)


class Engine:
    def __init__(self, args, vis, device):
        """args: ModelArgs-like (dim, n_layers, n_heads, n_kv_heads, vocab_size, multiple_of,
        ffn_dim_multiplier, norm_eps, rope_theta, rope_scaling, max_seq_len); vis: vision geometry or
        None."""
        self.args, self.vis, self.device = args, vis, torch.device(device)
        if self.device.type != "cuda":
            raise ops._lib.A3Error("a3vlm_b200.Engine needs a CUDA (sm_100a) device: there is no CPU path")
        ops._L()  # fail loudly now if the native library is missing
        a = args
        # grouped-query attention (llama_ens5.py:66,76-90,131-133): wk / wv and the KV cache carry n_kv_heads heads; the
        # prefill / training attention kernels run on the repeated heads (a3_repeat_kv, the reference's own repeat_kv
        # materialisation), the decode kernels index the group head directly
        self.n_kv = a.n_heads if a.n_kv_heads is None else a.n_kv_heads
        if self.n_kv <= 0 or a.n_heads % self.n_kv:
            raise ValueError("n_heads must be a multiple of n_kv_heads")
        self.n_rep = a.n_heads // self.n_kv
        if a.rope_scaling is not None:
            raise NotImplementedError("rope_scaling: direction cannot be verified from the reference tree")
        self.head_dim = a.dim // a.n_heads
        if self.head_dim != 128:
            raise NotImplementedError("decoder head_dim must be 128 (7B / 13B)")
        self.kv_dim = self.n_kv * self.head_dim
        self._gqa_scratch: Dict[Tuple, torch.Tensor] = {}
        self.ffn_dim = ffn_hidden_dim(a.dim, a.multiple_of, a.ffn_dim_multiplier)
        self.fuse_swiglu = self.ffn_dim % 128 == 0      # forward: gate in the w1|w3 epilogue (-40 us / layer)
        # backward: gate backward as the epilogue of the w2 dgrad (dy never reaches HBM), A3_FUSE_DSWIGLU=1: 675 us per layer
        # against 489 + 138 unfused -- the epilogue is a latency chain as long as the main loop at K 4096.  Prefetching the
        # saved g | u (and the residual rows of the generic epilogue) one chunk ahead was built in round 2: 568 us in
        # isolation, but at 165 registers the GEMMs of a live step dropped from 0.82 to 0.77 of burst and the step did not
        # move (272.9 vs 273.1 / 273.4 ms, power-capped) -- reverted; plain dgrad + a3_swiglu_bwd stays the default
        self.fuse_dswiglu = os.environ.get("A3_FUSE_DSWIGLU", "0") == "1"
        self.n_sms = torch.cuda.get_device_properties(self.device).multi_processor_count
        self.rope = ops.rope_table(self.head_dim, a.max_seq_len * 2, a.rope_theta, device=self.device)
        # one view: start + N + end; reference tiling: 5 views of a (2S x 2S) image (llama_ens5.py:383-385)
        self.views = getattr(vis, "views", 1) if vis is not None else 1
        # RGB-D build (LLM/llama_ens5_2images.py:335-336): a depth stream of the same geometry follows the
        # RGB stream at inference; image_words counts both, visual_image_words one of them
        self.rgbd = bool(getattr(vis, "rgbd", False)) if vis is not None else False
        self.visual_image_words = self.views * (vis.tokens + 2) if vis is not None else 0
        self.image_words = self.visual_image_words * (2 if self.rgbd else 1)
        self.scale = 1.0 / math.sqrt(self.head_dim)
        self._build_layouts()
        self._ws: Dict[Tuple, Workspace] = {}
        self.master = self.m = self.v = None
        self.opt_step = 0
        self.norm_sq = torch.zeros(1, device=self.device, dtype=torch.float32)
        self.kcache: List[torch.Tensor] = []
        self.vcache: List[torch.Tensor] = []
        self.cache_B = 0
        self.cache_image_words = 0
        self.dec = None
        self.use_decode_graph = True
        self.decode_fused_max_b = 1  # batch 1: fused CUDA-core weight streaming (3.05 ms); larger: swapped-operand tcgen05 (B=4 3.9 ms vs 5.9 fused)
        self.comm_stream = None
        # sharded optimiser (ZeRO-1 style, enable_sharded_optimizer): this rank's slice of the decay region
        self.shard_group, self.shard_world, self.shard_rank = None, 1, 0
        # weight ranges still being written on a side stream (pipelined AdamW chunks, all-gathered chunks), in
        # forward-consumption order: (lo, hi, event).  wait_weights(upto) makes the compute stream wait for
        # those that start below `upto`.
        self._pending: List[Tuple[int, int, "torch.cuda.Event"]] = []
        # pipelined optimiser step (single replica / unsharded state): sumsq + AdamW run on a side stream,
        # chunk by chunk in forward-consumption order, underneath the NEXT step's ViT + decoder forward
        self.overlap_optimizer = False
        # activation recompute level for the training workspaces (Workspace): 0 none, 1 selective, 2 per block
        self.recompute = 0
        self.overlap_adamw_blocks_per_sm = 4  # measured: 4 CTAs/SM 275.8-277.5 ms/step, 1: 279.9, 16: 279.4, in line: 281-285
        self.opt_stream = None
        self.norm_ready = None

    # ------------------------------------------------------------------------------------ layouts
    def _build_layouts(self):
        decay, nodecay, frozen, self.kpad = build_layout(self.args, self.vis)
        self._segment_after = decay[-1][0]
        self.w = FlatParams(decay + nodecay, self.device, segment_after=self._segment_after)
        self.decay_numel = self.w.offsets[nodecay[0][0]][0]  # no-decay region starts here (multiple of 512)
        self.fz = FlatParams(frozen, self.device) if frozen else None
        self.g: Optional[FlatParams] = None  # gradients: allocated on first training use
        self._layout = decay + nodecay
        self._buckets = bucket_plan(self.w.offsets, self.w.numel, self.args.n_layers, self.decay_numel)
        self._opt_chunks = optimizer_chunks(self.w.offsets, self.decay_numel, self.args.n_layers)

    def param_views(self) -> Dict[str, torch.Tensor]:
        """Reference-named views (SURVEY 8b state-dict names) into the flat buffers."""
        D, F = self.args.dim, self.ffn_dim
        out: Dict[str, torch.Tensor] = {}
        for name, t in self.w.views.items():
            if name.endswith(".wqkv"):
                p = name[:-4] + "attention."
                out[p + "wq.weight"], out[p + "wk.weight"], out[p + "wv.weight"] = t[:D], t[D:D + self.kv_dim], t[D + self.kv_dim:]
            elif name.endswith(".w13"):
                p = name[:-3] + "feed_forward."
                out[p + "w1.weight"], out[p + "w3.weight"] = t[:F], t[F:]
            else:
                out[name] = t
        if self.fz is not None:
            for name, t in self.fz.views.items():
                if not name.endswith("conv1_padded"):
                    out[name] = t
        return out

    def grad_views(self) -> Dict[str, torch.Tensor]:
        self._ensure_grads()
        D, F = self.args.dim, self.ffn_dim
        out: Dict[str, torch.Tensor] = {}
        for name, t in self.g.views.items():
            if name.endswith(".wqkv"):
                p = name[:-4] + "attention."
                out[p + "wq.weight"], out[p + "wk.weight"], out[p + "wv.weight"] = t[:D], t[D:D + self.kv_dim], t[D + self.kv_dim:]
            elif name.endswith(".w13"):
                p = name[:-3] + "feed_forward."
                out[p + "w1.weight"], out[p + "w3.weight"] = t[:F], t[F:]
            else:
                out[name] = t
        return out

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        self.wait_weights()
        views = self.param_views()
        missing = [k for k in views if k not in sd]
        if missing:
            raise KeyError(f"missing parameters: {missing[:5]} ... ({len(missing)})")
        for k, v in views.items():
            v.copy_(sd[k].to(device=self.device, dtype=bf16).view(v.shape))
        self.weights_changed()

    def weights_changed(self) -> None:
        """Re-derive the kernel-side layouts from the reference-named weights (conv1 as a K-major [1024][768]
        matrix, k = c * 256 + ky * 16 + kx, the 14 x 14 windows zero-padded to 16 x 16 for the TMA patch loader)
        and invalidate fp32 masters."""
        if self.fz is not None:
            c = "clip.visual."
            v = self.vis
            if v.patch > 16:
                raise NotImplementedError("patch sizes above 16 are not built (CLIP ViT-L/14)")
            pad = self.fz.views[c + "conv1_padded"]
            pad.zero_()
            pad.view(v.width, 3, 16, 16)[:, :, :v.patch, :v.patch] = self.fz.views[c + "conv1.weight"]
        self.master = None

    # ------------------------------------------------------------------------------------ workspaces
    def _vit_workspace(self, B: int, train: bool):
        v, D, dev = self.vis, self.args.dim, self.device
        g2, N, Wv = v.grid * v.grid, v.tokens, v.width
        e = lambda *s, dt=bf16: torch.empty(*s, device=dev, dtype=dt)
        ws = {"img16": e(B, 3, v.image_size, v.grid * 16), "po": e(B * g2, Wv), "x": e(B * N, Wv), "y": e(B * N, Wv),
              "qkv": e(B * N, 3 * Wv), "o": e(B * N, Wv), "fc": e(B * N, v.mlp), "feats": e(B * N, Wv),
              "pre": e(B * N, D), "tok": e(B * N, D), "mean": e(B * N, dt=torch.float32),
              "rstd": e(B * N, dt=torch.float32)}
        if train:
            ws["dtok"] = e(B * N, D)
            ws["dpre"] = e(B * N, D)
            ws["pw"] = e(self.n_sms, D, dt=torch.float32)
            ws["pb"] = e(self.n_sms, D, dt=torch.float32)
        return ws

    def workspace(self, B: int, T: int, W: int, train: bool) -> Workspace:
        key = (B, T, W, train, self.recompute if train else 0)
        ws = self._ws.get(key)
        if ws is None:
            # one live shape per mode: drop workspaces of other shapes (they are GBs at 7B)
            for k in [k for k in self._ws if k[3] == train]:
                del self._ws[k]
            ws = self._ws[key] = Workspace(self, B, T, W, train)
        return ws

    def _ensure_grads(self):
        if self.g is None:
            self.g = FlatParams(self._layout, self.device, segment_after=self._segment_after)

    # ------------------------------------------------------------------------------------ ViT + projector
    def vit_forward(self, image: torch.Tensor, vw, depth: Optional[torch.Tensor] = None) -> torch.Tensor:
        """clip_encode_image (LLM/llama_ens5.py:351-375) + visual_proj (:330-333,451), frozen tower.
        image (B,3,S,S) fp32 or bf16 on the device -> image tokens [B*N][D] (vw['tok']).  With `depth`
        (RGB-D build) both streams go through the tower as one batch, RGB views first."""
        v, f = self.vis, self.fz.views
        c = "clip.visual."
        B0 = image.shape[0]
        S_in = v.image_size * (2 if self.views == 5 else 1)
        N, Wv, g2 = v.tokens, v.width, v.grid * v.grid
        streams = [image] if depth is None else [image, depth]
        for s, im in enumerate(streams):
            if tuple(im.shape) != (B0, 3, S_in, S_in):
                raise ValueError(f"image must be ({B0},3,{S_in},{S_in}), got {tuple(im.shape)}")
            if im.dtype not in (torch.float32, bf16):
                im = im.float()
            im = im.contiguous()
            if self.views == 5:  # [global bicubic | TL | TR | BL | BR] x B, view-major (llama_ens5.py:383-385)
                ops.build_views(im, vw["views"][s * 5 * B0:(s + 1) * 5 * B0], B0, v.image_size)
            else:  # the autocast cast of the conv input, patch columns on a 16-element pitch for the TMA loader
                ops.vit_pad_cast(im, vw["img16"][s * B0:(s + 1) * B0], B0, v.image_size, v.patch)
        B = len(streams) * self.views * B0
        if self.views == 5:
            ops.vit_pad_cast(vw["views"], vw["img16"], B, v.image_size, v.patch)
        act = ops.ACT_QUICKGELU if v.quick_gelu else ops.ACT_GELU
        # conv1 without im2col: 4-D TMA boxes over the NCHW pixels land in the UMMA layout (a3_vit_patch_embed)
        ops.vit_patch_embed(vw["img16"], f[c + "conv1_padded"], vw["po"], B, v.image_size, v.patch)
        ops.vit_assemble_lnpre(vw["po"], f[c + "class_embedding"], f[c + "positional_embedding"],
                               f[c + "ln_pre.weight"], f[c + "ln_pre.bias"], vw["x"], B, g2, Wv, v.ln_eps)
        x, y, qkv, o, fc = vw["x"], vw["y"], vw["qkv"], vw["o"], vw["fc"]
        hd = Wv // v.heads
        sc = 1.0 / math.sqrt(hd)
        for i in range(v.layers):
            r = f"{c}transformer.resblocks.{i}."
            ops.layernorm_fwd(x, f[r + "ln_1.weight"], f[r + "ln_1.bias"], v.ln_eps, out=y)
            ops.gemm(y, f[r + "attn.in_proj_weight"], qkv, bias=f[r + "attn.in_proj_bias"])
            ops.attn_fwd(qkv[:, :Wv], qkv[:, Wv:2 * Wv], qkv[:, 2 * Wv:], B, v.heads, N, N, hd, sc, False, out=o)
            ops.gemm(o, f[r + "attn.out_proj.weight"], x, bias=f[r + "attn.out_proj.bias"], residual=x)
            ops.layernorm_fwd(x, f[r + "ln_2.weight"], f[r + "ln_2.bias"], v.ln_eps, out=y)
            ops.gemm(y, f[r + "mlp.c_fc.weight"], fc, bias=f[r + "mlp.c_fc.bias"], act=act)
            ops.gemm(fc, f[r + "mlp.c_proj.weight"], x, bias=f[r + "mlp.c_proj.bias"], residual=x)
        ops.layernorm_fwd(x, f[c + "ln_post.weight"], f[c + "ln_post.bias"], v.ln_eps, out=vw["feats"])
        w = self.w.views
        self.wait_weights(self._range("end_img")[1])
        ops.gemm(vw["feats"], w["visual_proj.0.weight"], vw["pre"], bias=w["visual_proj.0.bias"])
        ops.layernorm_fwd(vw["pre"], w["visual_proj.1.weight"], w["visual_proj.1.bias"], 1e-5, out=vw["tok"],
                          mean=vw["mean"], rstd=vw["rstd"])
        return vw["tok"]

    def _scratch(self, name: str, rows: int, cols: int) -> torch.Tensor:
        key = (name, rows, cols)
        t = self._gqa_scratch.get(key)
        if t is None:
            t = self._gqa_scratch[key] = torch.empty(rows, cols, device=self.device, dtype=bf16)
        return t

    # ------------------------------------------------------------------------------------ decoder layer
    def _layer_fwd(self, i, x, xn, qkv, o, h, hn, gu, y, out, lse, rstd1, rstd2, B, L, pos0=0,
                   kv=None, skip_out=False):
        a, w, D = self.args, self.w.views, self.args.dim
        p = f"layers.{i}."
        ops.rmsnorm_fwd(x, w[p + "attention_norm.weight"], a.norm_eps, out=xn, rstd=rstd1)
        Dkv = self.kv_dim
        if self.n_rep == 1:
            ops.gemm(xn, w[p + "wqkv"], qkv, rope=self.rope, rope_cols=2 * D, rope_period=L, rope_pos0=pos0)
            kc, vc = qkv[:, D:2 * D], qkv[:, 2 * D:]
        else:  # compact q | k | v (Hkv key / value heads), then the repeated heads the attention kernels read
            qc = self._scratch("qkv_c", qkv.shape[0], D + 2 * Dkv)
            ops.gemm(xn, w[p + "wqkv"], qc, rope=self.rope, rope_cols=D + Dkv, rope_period=L, rope_pos0=pos0)
            kc, vc = qc[:, D:D + Dkv], qc[:, D + Dkv:]
            ops.repeat_kv(qc, qkv, a.n_heads, self.n_kv, self.head_dim)
        q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
        if kv is not None:  # LLM/llama_ens5.py:121-127
            ops.kv_append(kc, vc, self.kcache[i], self.vcache[i], B, L, self.n_kv, self.head_dim, pos0,
                          a.max_seq_len)
        if kv is not None and pos0 > 0:  # multi-token step against the cache, right-aligned causal
            Wd = a.n_heads * self.head_dim
            if self.n_rep == 1:
                kfull, vfull, bstride = self.kcache[i].view(-1, Wd), self.vcache[i].view(-1, Wd), a.max_seq_len * Wd
            else:  # rare path (a multi-token append after the prefill): repeat the cached group heads, as the reference does
                kfull = self.kcache[i][:, :pos0 + L].repeat_interleave(self.n_rep, dim=2).reshape(-1, Wd)
                vfull = self.vcache[i][:, :pos0 + L].repeat_interleave(self.n_rep, dim=2).reshape(-1, Wd)
                bstride = (pos0 + L) * Wd
            ops.attn_fwd(q, kfull, vfull, B, a.n_heads, L,
                         pos0 + L, self.head_dim, self.scale, True, out=o, lse=lse,
                         kv_bstride=bstride)
        else:
            ops.attn_fwd(q, k, v, B, a.n_heads, L, L, self.head_dim, self.scale, True, out=o, lse=lse)
        ops.gemm(o, w[p + "attention.wo.weight"], h, residual=x)
        ops.rmsnorm_fwd(h, w[p + "ffn_norm.weight"], a.norm_eps, out=hn, rstd=rstd2)
        if self.fuse_swiglu:  # gate fused into the w1|w3 projection's epilogue (one launch, y never re-read)
            ops.gemm_swiglu(hn, w[p + "w13"], gu, y)
        else:
            ops.gemm(hn, w[p + "w13"], gu)
            ops.swiglu_fwd(gu, y)
        if not skip_out:  # the backward's block recompute already holds the block output x[i + 1]
            ops.gemm(y, w[p + "feed_forward.w2.weight"], out, residual=h)

    # ------------------------------------------------------------------------------------ training forward
    def forward_train(self, examples: torch.Tensor, image: Optional[torch.Tensor], labels=None,
                      grad_scale: float = 1.0, want_grad: bool = True):
        """Transformer.forward (LLM/llama_ens5.py:461-487) keeping what the backward needs.
        Returns (workspace, logits [B*T][V] view).  With labels: also runs the shifted cross entropy
        of M/model/meta.py:256-262 (loss in ws.loss; logits are overwritten by dlogits)."""
        a, w = self.args, self.w.views
        B, T = examples.shape
        # the training forward splices the RGB stream only, also in the RGB-D build (M/model/meta.py:251)
        W = self.visual_image_words if image is not None else 0
        ws = self.workspace(B, T, W, True)
        ws.generation += 1
        L, D = ws.L, a.dim
        if L > 2 * a.max_seq_len:
            raise ValueError("sequence longer than the RoPE table (2 * max_seq_len)")
        # the backward scatters into the embedding by these ids: keep the dense copy with the workspace (a
        # trimmed `examples[:, :k]` view handed to the kernels as a bare pointer would be read with the wrong stride)
        examples = ws.examples = examples.contiguous()
        img = self.vit_forward(image, ws.vit) if W else None
        self.wait_weights(self._range("tok_embeddings.weight")[1])
        ops.embed_splice(examples, w["tok_embeddings.weight"], img, w["start_img"] if W else None,
                         w["end_img"] if W else None, ws.x[0], B, T, self.vis.tokens if W else 0, D, self.views)
        ks, kb = (lambda i: 0 if ws.recompute >= 1 else i), (lambda i: 0 if ws.recompute >= 2 else i)
        for i in range(a.n_layers):
            self.wait_weights(self._range(f"layers.{i}.feed_forward.w2.weight")[1])
            self._layer_fwd(i, ws.x[i], ws.xn[ks(i)], ws.qkv[kb(i)], ws.o[kb(i)], ws.h[kb(i)], ws.hn[ks(i)],
                            ws.gu[kb(i)], ws.y[ks(i)], ws.x[i + 1], ws.lse[kb(i)], ws.rstd1[ks(i)],
                            ws.rstd2[ks(i)], B, L)
        self.wait_weights()
        xf = ws.x[a.n_layers].view(B, L, D)
        # final norm on the text rows only (llama_ens5.py:485-486): one grouped launch over [B][T] rows
        ops.rmsnorm_fwd(xf[:, W:], w["norm.weight"], a.norm_eps, out=ws.fn, rstd=ws.rstd_f)
        ops.gemm(ws.fn, w["output.weight"], ws.logits)
        if labels is not None:
            ops.shift_labels(labels, ws.shift, ws.count)
            ops.cross_entropy(ws.logits, ws.shift, ws.row_loss, ws.count, grad_scale, want_grad)
            ops.loss_reduce(ws.row_loss, ws.count, ws.loss)
        return ws, ws.logits

    # ------------------------------------------------------------------------------------ backward
    def backward(self, ws: Workspace, examples: torch.Tensor, dlogits: torch.Tensor,
                 accumulate: bool = False, on_bucket=None, scale: float = 1.0) -> None:
        """Backward of forward_train given dlogits [B*T][V] (bf16).  Weight gradients go to the flat
        bf16 gradient buffer (overwritten, or added to when `accumulate`).  `on_bucket(lo, hi)` is
        called with the flat element range whose gradients are complete (all-reduce overlap).
        `scale` multiplies dlogits on the fly (the upstream d loss, and 1/world for data parallel)."""
        self._ensure_grads()
        a, w, g = self.args, self.w.views, self.g.views
        B, T, W, L, M, D, F = ws.B, ws.T, ws.W, ws.L, ws.M, a.dim, self.ffn_dim
        NL, acc = a.n_layers, accumulate
        examples = examples.contiguous()  # see forward_train: never index a strided view through a bare pointer
        wg = lambda dy, x, out: ops.gemm(dy, x, out, a_mn=True, b_mn=True, accumulate=acc)   # dW = dy^T x
        dg = lambda dy, wt, out: ops.gemm(dy, wt, out, b_mn=True)                             # dx = dy W
        np_ = ws.nparts
        # LM head + final norm
        ops.gemm(dlogits, ws.fn, g["output.weight"], a_mn=True, b_mn=True, accumulate=acc, alpha=scale)
        ops.gemm(dlogits, w["output.weight"], ws.dfn, b_mn=True, alpha=scale)
        bk = iter(self._buckets)
        if on_bucket:
            on_bucket(*next(bk))
        ws.d.zero_()
        xf, dv = ws.x[NL].view(B, L, D), ws.d.view(B, L, D)
        ops.rmsnorm_bwd(ws.dfn, xf[:, W:], w["norm.weight"], ws.rstd_f, dv[:, W:], ws.part2)
        ops.colsum_partials(ws.part2, g["norm.weight"], accumulate=acc)
        d, t = ws.d, ws.dtmp
        rc = ws.recompute
        for i in reversed(range(NL)):
            p = f"layers.{i}."
            ks, kb = (0 if rc >= 1 else i), (0 if rc >= 2 else i)
            if rc >= 2:    # the reference's per-block activation checkpoint: re-run the block (minus its last GEMM)
                self._layer_fwd(i, ws.x[i], ws.xn[0], ws.qkv[0], ws.o[0], ws.h[0], ws.hn[0], ws.gu[0], ws.y[0],
                                None, ws.lse[0], ws.rstd1[0], ws.rstd2[0], B, L, skip_out=True)
            elif rc == 1:  # selective: rebuild the norm outputs and the gate output (three row kernels)
                ops.swiglu_fwd(ws.gu[i], ws.y[0])
                ops.rmsnorm_fwd(ws.h[i], w[p + "ffn_norm.weight"], a.norm_eps, out=ws.hn[0], rstd=ws.rstd2[0])
                ops.rmsnorm_fwd(ws.x[i], w[p + "attention_norm.weight"], a.norm_eps, out=ws.xn[0], rstd=ws.rstd1[0])
            wg(d, ws.y[ks], g[p + "feed_forward.w2.weight"])
            if self.fuse_dswiglu:  # gate backward fused behind the w2 dgrad: dy stays in TMEM / registers
                ops.gemm_dswiglu(d, w[p + "feed_forward.w2.weight"], ws.gu[kb], ws.dgu)
            else:
                dg(d, w[p + "feed_forward.w2.weight"], ws.dy)
                ops.swiglu_bwd(ws.gu[kb], ws.dy, ws.dgu)
            wg(ws.dgu, ws.hn[ks], g[p + "w13"])
            dg(ws.dgu, w[p + "w13"], t)
            ops.rmsnorm_bwd(t, ws.h[kb], w[p + "ffn_norm.weight"], ws.rstd2[ks], d, ws.part2, add_to_dx=True)
            ops.colsum_partials(ws.part2, g[p + "ffn_norm.weight"], accumulate=acc)
            wg(d, ws.o[kb], g[p + "attention.wo.weight"])
            dg(d, w[p + "attention.wo.weight"], t)
            qkv, dqkv = ws.qkv[kb], ws.dqkv
            ops.attn_bwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], ws.o[kb], t, ws.lse[kb], ws.delta, ws.dq_acc,
                         dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:], B, a.n_heads, L, self.head_dim,
                         self.scale, True, rope=self.rope)  # backward of RoPE fused into the dq / dk stores
            if self.n_rep > 1:  # backward of repeat_kv: sum the copies of every group head
                dqkv = ops.reduce_kv(dqkv, self._scratch("dqkv_c", dqkv.shape[0], D + 2 * self.kv_dim), a.n_heads,
                                     self.n_kv, self.head_dim)
            wg(dqkv, ws.xn[ks], g[p + "wqkv"])
            dg(dqkv, w[p + "wqkv"], t)
            ops.rmsnorm_bwd(t, ws.x[i], w[p + "attention_norm.weight"], ws.rstd1[ks], d, ws.part2, add_to_dx=True)
            ops.colsum_partials(ws.part2, g[p + "attention_norm.weight"], accumulate=acc)
            if on_bucket:
                on_bucket(*next(bk))
        # embedding scatter-add, tags, projector
        if not acc:
            g["tok_embeddings.weight"].zero_()
        vw = ws.vit if W else None
        ops.embed_splice_bwd(examples, d, g["tok_embeddings.weight"], vw["dtok"] if W else None,
                             ws.tag_grad[0] if W else None, ws.tag_grad[1] if W else None, B, T,
                             self.vis.tokens if W else 0, D, self.views)
        if W:
            ops.colsum_partials(ws.tag_grad[0:1], g["start_img"].view(-1), accumulate=acc)
            ops.colsum_partials(ws.tag_grad[1:2], g["end_img"].view(-1), accumulate=acc)
            ops.layernorm_bwd(vw["dtok"], vw["pre"], w["visual_proj.1.weight"], vw["mean"], vw["rstd"],
                              vw["dpre"], vw["pw"], vw["pb"])
            ops.colsum_partials(vw["pw"], g["visual_proj.1.weight"], accumulate=acc)
            ops.colsum_partials(vw["pb"], g["visual_proj.1.bias"], accumulate=acc)
            wg(vw["dpre"], vw["feats"], g["visual_proj.0.weight"])
            ops.colsum_bf16(vw["dpre"], g["visual_proj.0.bias"], accumulate=acc)
        elif not acc and "start_img" in g:
            for n in ("start_img", "end_img", "visual_proj.0.weight", "visual_proj.0.bias",
                      "visual_proj.1.weight", "visual_proj.1.bias"):
                g[n].zero_()
        if on_bucket:
            for lo_hi in bk:
                on_bucket(*lo_hi)

    def _range(self, name: str) -> Tuple[int, int]:
        o, n, _ = self.w.offsets[name]
        return o, o + n

    # ------------------------------------------------------------------------------------ optimiser
    def enable_sharded_optimizer(self, group=None) -> None:
        """ZeRO-2 style sharding over the data-parallel ranks (the reference shards too: FSDP
        SHARD_GRAD_OP, M/main_finetune.py:256-260).  Ownership is sliced PER CHUNK: the decay region is cut
        into chunks in forward-consumption order (projector + embedding, one per decoder layer, LM head) and
        rank r owns the r-th 1/N of every chunk.  Then
          * the backward's gradient exchange is one in-place REDUCE-SCATTER per chunk (each rank receives
            only the summed gradients of its slice: half the bytes of an all-reduce),
          * fp32 master / m / v exist for the owned slices only (13B: 26 + 26 + 156/8 GB per GPU),
          * the step updates the owned slice of a chunk and ALL-GATHERs the chunk's bf16 weights in place,
            chunk by chunk on the side stream, so the next forward (which waits per chunk) starts while
            later chunks are still being updated and exchanged.
        The tiny no-decay tail (norm weights, biases) is all-reduced and updated redundantly on every rank."""
        import torch.distributed as dist

        self.shard_group = group if group is not None else dist.group.WORLD
        self.shard_world = dist.get_world_size(self.shard_group)
        self.shard_rank = dist.get_rank(self.shard_group)
        for lo, hi in self._opt_chunks:
            if (hi - lo) % (self.shard_world * 8):
                raise ValueError(f"chunk [{lo}, {hi}) does not split into {self.shard_world} 16-byte aligned slices")
        self.master = self.m = self.v = None
        if self.comm_stream is None:
            self.comm_stream = torch.cuda.Stream(device=self.device)

    def _opt_ranges(self):
        """[(lo, hi, decayed)] flat ranges whose optimiser state this rank holds: its slice of every decay
        chunk in forward order (the whole chunk for a single replica), then the no-decay tail."""
        return owned_ranges(self._opt_chunks, self.decay_numel, self.w.flat.numel(), self.shard_world, self.shard_rank)

    def _ensure_opt(self):
        if self.master is None:
            self.wait_weights()
            self.master = [self.w.flat[lo:hi].float() for lo, hi, _ in self._opt_ranges()]
        if self.m is None:
            self.m = [torch.zeros_like(t) for t in self.master]
            self.v = [torch.zeros_like(t) for t in self.master]

    def optimizer_step(self, lr: float, weight_decay: float = 0.0, max_norm: float = 0.0,
                       betas=(0.9, 0.95), eps: float = 1e-8) -> torch.Tensor:
        """Global-L2 clip (M/util/clip_grad.py:187-193) + AdamW on fp32 masters with the bf16 weights
        rewritten in the same pass (M/main_finetune.py:284-285; M/util/misc.py:586-599 decay groups).
        Returns the device tensor holding the squared gradient norm (no host sync; read it through
        grad_norm_sq()).

        Schedule.  In line on the current stream for a single replica unless `overlap_optimizer`; otherwise
        (and always with the sharded optimiser) on a side stream behind an event that marks the end of the
        backward: norm, the small no-decay tail first (layer 0 needs its norm weight at once), then the decay
        chunks in forward-consumption order -- AdamW on the owned range, all-gather of the chunk when sharded,
        an event after each.  The caller's next forward waits per chunk (wait_weights), so the frozen ViT
        tower and the early layers run while later chunks are still being updated: AdamW is HBM-bound
        (28 B/param), the forward tensor-bound, and a small AdamW grid fits beside the resident GEMM CTA.
        Hazards: the forward's last wait_weights() (before the LM head) leaves the update fully consumed
        before any backward can overwrite the gradients or read the weights for dgrad."""
        self._ensure_grads()
        self._ensure_opt()
        self.opt_step += 1
        self.wait_weights()  # an earlier update nobody consumed
        sharded = self.shard_world > 1
        side = sharded or self.overlap_optimizer
        ranges = self._opt_ranges()
        b1, b2 = betas
        if side:
            st = self.comm_stream if sharded else self.opt_stream
            if st is None:
                st = self.opt_stream = torch.cuda.Stream(device=self.device)
            done = torch.cuda.Event()
            done.record()
            st.wait_event(done)
        else:
            st = torch.cuda.current_stream()
        pend = []
        if side:
            ops.set_adamw_blocks_per_sm(self.overlap_adamw_blocks_per_sm)
        try:
            with torch.cuda.stream(st):
                self.norm_sq.zero_()
                if sharded:
                    import torch.distributed as dist

                    for lo, hi, decayed in ranges:  # owned slices hold fully reduced gradients
                        if decayed and hi > lo:
                            ops.sumsq_bf16(self.g.flat[lo:hi], self.norm_sq)
                    dist.all_reduce(self.norm_sq, op=dist.ReduceOp.SUM, group=self.shard_group)
                    tlo, thi, _ = ranges[-1]  # the tail was all-reduced: identical on every rank, counted once
                    if thi > tlo:
                        ops.sumsq_bf16(self.g.flat[tlo:thi], self.norm_sq)
                else:
                    ops.sumsq_bf16(self.g.flat, self.norm_sq)
                if side:
                    self.norm_ready = torch.cuda.Event()
                    self.norm_ready.record(st)
                else:
                    self.norm_ready = None
                order = [len(ranges) - 1] + list(range(len(ranges) - 1))  # tail first, then forward order
                for k in order:
                    lo, hi, decayed = ranges[k]
                    if hi > lo:
                        ops.adamw(self.master[k], self.w.flat[lo:hi], self.g.flat[lo:hi], self.m[k], self.v[k], lr, b1, b2,
                                  eps, weight_decay if decayed else 0.0, self.opt_step, self.norm_sq, max_norm)
                    clo, chi = (self._opt_chunks[k] if decayed else (-1, 0))
                    if sharded and decayed:
                        dist.all_gather_into_tensor(self.w.flat[clo:chi], self.w.flat[lo:hi], group=self.shard_group)
                    if side:
                        ev = torch.cuda.Event()
                        ev.record(st)
                        pend.append((clo, chi, ev))
        finally:
            if side:
                ops.set_adamw_blocks_per_sm(16)
        self._pending = pend
        return self.norm_sq

    def grad_norm_sq(self) -> torch.Tensor:
        """Squared global gradient norm of the last optimiser step, safe to read on the current stream."""
        if self.norm_ready is not None:
            torch.cuda.current_stream().wait_event(self.norm_ready)
        return self.norm_sq

    def optimizer_state_dict(self) -> Dict:
        """This rank's AdamW state (fp32 master / m / v of the flat ranges it owns, step count) on the host
        -- the `consolidated.*.optimizer.pth` payload of checkpoint.save_checkpoint (M/util/misc.py:390-406).
        With the sharded optimiser every rank owns a different slice of every chunk."""
        self._ensure_opt()
        self.wait_weights()
        cpu = lambda ts: [t.detach().to("cpu") for t in ts]
        return {"step": self.opt_step, "ranges": [(lo, hi, d) for lo, hi, d in self._opt_ranges()],
                "shard_world": self.shard_world, "shard_rank": self.shard_rank,
                "master": cpu(self.master), "m": cpu(self.m), "v": cpu(self.v)}

    def load_optimizer_state_dict(self, sd: Dict) -> None:
        """Inverse of optimizer_state_dict (resume, M/util/misc.py:510-523).  The sharding must match the
        one the state was saved with."""
        ranges = [(lo, hi, d) for lo, hi, d in self._opt_ranges()]
        if [tuple(r) for r in sd["ranges"]] != ranges:
            raise ValueError("optimizer state was saved for different flat ranges than this engine owns "
                             "(different model geometry or data-parallel sharding)")
        to = lambda ts: [t.to(device=self.device, dtype=torch.float32).contiguous() for t in ts]
        self.wait_weights()
        self.master, self.m, self.v = to(sd["master"]), to(sd["m"]), to(sd["v"])
        self.opt_step = int(sd["step"])

    def wait_weights(self, upto: Optional[int] = None) -> None:
        """Make the current stream wait until the weights at flat offsets [0, upto) written on a side
        stream (pipelined AdamW chunks / all-gathered chunks) have landed -- all of them when upto is None.
        No-op without a pending update."""
        if not self._pending:
            return
        cur = torch.cuda.current_stream()
        while self._pending and (upto is None or self._pending[0][0] < upto):
            _, _, ev = self._pending.pop(0)
            cur.wait_event(ev)

    # ------------------------------------------------------------------------------------ data-parallel
    def allreduce_hook(self, group=None):
        """(on_bucket, finish) for the backward.  Replicated optimiser: bucketed all-reduce
        (dp.BucketAllReduce).  Sharded optimiser: reduce-scatter of every decay chunk to its owners and an
        all-reduce of the small tail (dp.BucketReduceScatter)."""
        from .dp import BucketAllReduce, BucketReduceScatter

        self._ensure_grads()
        if self.comm_stream is None:
            self.comm_stream = torch.cuda.Stream(device=self.device)
        if self.shard_world > 1:
            r = BucketReduceScatter(self.g.flat, self.decay_numel, self.shard_world, self.shard_rank, group, self.comm_stream)
        else:
            r = BucketAllReduce(self.g.flat, group, self.comm_stream)
        return r.on_bucket, r.finish

    # ------------------------------------------------------------------------------------ inference
    def allocate_kv_cache(self, B: int) -> None:
        """LLM/llama_ens5.py:171-176,533-535: (B, max_seq_len, H, 128) x 2 per layer, reused while the
        shape is unchanged."""
        a = self.args
        if self.cache_B == B and self.kcache:
            return
        self.kcache = self.vcache = []
        shape = (B, a.max_seq_len, self.n_kv, self.head_dim)
        self.kcache = [torch.empty(shape, device=self.device, dtype=bf16) for _ in range(a.n_layers)]
        self.vcache = [torch.empty(shape, device=self.device, dtype=bf16) for _ in range(a.n_layers)]
        self.cache_B = B
        self.dec = None

    def destroy_kv_cache(self) -> None:
        self.kcache, self.vcache, self.cache_B, self.dec = [], [], 0, None

    def forward_inference(self, tokens: torch.Tensor, start_pos: int, image=None, depth=None) -> torch.Tensor:
        """Transformer.forward_inference (LLM/llama_ens5.py:490-531): fp32 last-position logits (B,V).
        RGB-D build (LLM/llama_ens5_2images.py:511-547): the prefill takes both image streams or none."""
        a, w = self.args, self.w.views
        if self.rgbd and (image is None or depth is None):
            image = depth = None  # llama_ens5_2images.py:518: `if image is not None and depth_images is not None`
        if not self.rgbd:
            depth = None
        B, l = tokens.shape
        D, V = a.dim, a.vocab_size
        self.wait_weights()
        if start_pos == 0:
            self.allocate_kv_cache(B)
        elif B != self.cache_B:
            raise ValueError("batch size changed between prefill and decode")
        if image is not None:
            assert start_pos == 0  # llama_ens5.py:499
            self.cache_image_words = self.image_words
        elif start_pos == 0:
            self.cache_image_words = 0
        else:
            start_pos = start_pos + self.cache_image_words  # llama_ens5.py:520
        if l == 1 and start_pos > 0:
            return self._decode_step(tokens, start_pos)
        W = self.image_words if image is not None else 0
        L = l + W
        if start_pos + L > a.max_seq_len:
            raise ValueError(f"sequence {start_pos + L} exceeds max_seq_len {a.max_seq_len}")
        ws = self.workspace(B, l, W, False)
        tokens = tokens.contiguous()
        img = self.vit_forward(image, ws.vit, depth) if W else None
        if W and depth is not None:
            fz = self.fz.views
            ops.embed_splice_rgbd(tokens, w["tok_embeddings.weight"], img, w["start_img"], w["end_img"],
                                  fz["start_depth_img"], fz["end_depth_img"], ws.x[0], B, l, self.vis.tokens, D,
                                  self.views)
        else:
            ops.embed_splice(tokens, w["tok_embeddings.weight"], img, w["start_img"] if W else None,
                             w["end_img"] if W else None, ws.x[0], B, l, self.vis.tokens if W else 0, D, self.views)
        xa, xb = ws.x[0], ws.x[1]
        for i in range(a.n_layers):
            self._layer_fwd(i, xa, ws.xn[0], ws.qkv[0], ws.o[0], ws.h[0], ws.hn[0], ws.gu[0], ws.y[0], xb,
                            None, None, None, B, L, pos0=start_pos, kv=True)
            xa, xb = xb, xa
        last = xa.view(B, L, D)[:, L - 1]  # (B, D) strided view: last position of every sample
        fn = ops.rmsnorm_fwd(last, w["norm.weight"], a.norm_eps)
        return self._lm_head(fn)

    def _lm_head(self, fn: torch.Tensor) -> torch.Tensor:
        w, B = self.w.views, fn.shape[0]
        if B <= 8:
            return ops.decode_linear(fn, w["output.weight"],
                                     torch.empty(B, self.args.vocab_size, device=self.device, dtype=torch.float32),
                                     out_fp32=True)
        return ops.gemm(fn, w["output.weight"]).float()

    def _decode_step(self, tokens: torch.Tensor, pos: int) -> torch.Tensor:
        """One-token step (q_len = 1): weight-streaming linears + fused RoPE/KV-append + split-KV
        attention over the cache.  The kernel sequence takes the position and the token ids from
        device memory, so after one eager step it is captured once per batch size and replayed as a
        single CUDA graph (~330 launches per token would otherwise be launch-bound)."""
        a = self.args
        B, D, F = tokens.shape[0], a.dim, self.ffn_dim
        if pos + 1 > a.max_seq_len:
            raise ValueError("KV cache overflow")
        if self.dec is None or self.dec["B"] != B:
            e = lambda *s, dt=bf16: torch.empty(*s, device=self.device, dtype=dt)
            self.dec = {"B": B, "x": e(B, D), "x2": e(B, D), "xn": e(B, D), "qkv": e(B, D + 2 * self.kv_dim), "o": e(B, D),
                        "h": e(B, D), "gu": e(B, 2 * F), "y": e(B, F),
                        "scratch": torch.zeros(B * a.n_heads * (32 * (self.head_dim + 2) + 1), device=self.device,
                                               dtype=torch.float32),
                        "tok": torch.zeros(B, 1, device=self.device, dtype=torch.int64),
                        "pos": torch.zeros(1, device=self.device, dtype=torch.int32),
                        "logits": e(B, a.vocab_size, dt=torch.float32), "graph": None, "eager_steps": 0}
            if B > self.decode_fused_max_b:
                self.dec["logits_bf"] = e(B, a.vocab_size)
        d = self.dec
        d["tok"].copy_(tokens.reshape(B, 1))
        d["pos"].fill_(pos)
        if d["graph"] is not None:
            d["graph"].replay()
            return d["logits"]
        if self.use_decode_graph and d["eager_steps"] >= 1:
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._decode_kernels(d)
            d["graph"] = g
            g.replay()
            return d["logits"]
        d["eager_steps"] += 1
        self._decode_kernels(d)
        return d["logits"]

    def _decode_kernels(self, d) -> None:
        a, w = self.args, self.w.views
        B, D, F = d["B"], a.dim, self.ffn_dim
        ops.embed_splice(d["tok"], w["tok_embeddings.weight"], None, None, None, d["x"], B, 1, 0, D)
        x, x2 = d["x"], d["x2"]
        if B <= self.decode_fused_max_b:
            # five launches per layer: every norm / activation / RoPE / cache write is a prologue or
            # epilogue of the weight-streaming projection that consumes or produces it
            gated = os.environ.get("A3_DEC_GATED", "1") != "0"  # 0: round-1 form (gate as the prologue of w2), for A/B timing
            for i in range(a.n_layers):
                p = f"layers.{i}."
                ops.decode_linear(x, w[p + "wqkv"], d["qkv"], norm_w=w[p + "attention_norm.weight"], eps=a.norm_eps,
                                  rope=self.rope, n_heads=a.n_heads, n_kv_heads=self.n_kv, kcache=self.kcache[i],
                                  vcache=self.vcache[i], pos_dev=d["pos"], max_seq=a.max_seq_len)
                ops.attn_decode(d["qkv"][:, :D], self.kcache[i], self.vcache[i], d["o"], d["scratch"], B,
                                a.n_heads, self.head_dim, a.max_seq_len, a.max_seq_len, self.scale, pos_dev=d["pos"],
                                n_kv_heads=self.n_kv)
                ops.decode_linear(d["o"], w[p + "attention.wo.weight"], d["h"], residual=x)
                if gated:  # the gate is the epilogue of the w1|w3 projection: gu never reaches HBM
                    ops.decode_linear_gated(d["h"], w[p + "w13"], d["y"], norm_w=w[p + "ffn_norm.weight"], eps=a.norm_eps)
                    ops.decode_linear(d["y"], w[p + "feed_forward.w2.weight"], x2, residual=d["h"])
                else:
                    ops.decode_linear(d["h"], w[p + "w13"], d["gu"], norm_w=w[p + "ffn_norm.weight"], eps=a.norm_eps)
                    ops.decode_linear(d["gu"], w[p + "feed_forward.w2.weight"], x2, K=F, swiglu=True, residual=d["h"])
                x, x2 = x2, x
            ops.decode_linear(x, w["output.weight"], d["logits"], norm_w=w["norm.weight"], eps=a.norm_eps, out_fp32=True)
            return
        # larger decode batches: swapped-operand tcgen05 projection (weights are the 128-row MMA operand,
        # split-K over a cluster) + the small row kernels
        if B <= 64:
            lin = lambda x, wt, out, residual=None: ops.decode_gemm(x, wt, out, residual=residual)
        else:
            lin = lambda x, wt, out, residual=None: ops.gemm(x, wt, out, residual=residual)
        for i in range(a.n_layers):
            p = f"layers.{i}."
            ops.rmsnorm_fwd(x, w[p + "attention_norm.weight"], a.norm_eps, out=d["xn"])
            lin(d["xn"], w[p + "wqkv"], d["qkv"])
            ops.decode_rope_kv(d["qkv"], self.kcache[i], self.vcache[i], B, a.n_heads, self.head_dim, self.rope,
                               0, a.max_seq_len, pos_dev=d["pos"], n_kv_heads=self.n_kv)
            ops.attn_decode(d["qkv"][:, :D], self.kcache[i], self.vcache[i], d["o"], d["scratch"], B,
                            a.n_heads, self.head_dim, a.max_seq_len, a.max_seq_len, self.scale, pos_dev=d["pos"],
                            n_kv_heads=self.n_kv)
            lin(d["o"], w[p + "attention.wo.weight"], d["h"], residual=x)
            ops.rmsnorm_fwd(d["h"], w[p + "ffn_norm.weight"], a.norm_eps, out=d["xn"])
            lin(d["xn"], w[p + "w13"], d["gu"])
            ops.swiglu_fwd(d["gu"], d["y"])
            lin(d["y"], w[p + "feed_forward.w2.weight"], x2, residual=d["h"])
            x, x2 = x2, x
        fn = ops.rmsnorm_fwd(x, w["norm.weight"], a.norm_eps, out=d["xn"])
        if B <= 64:
            ops.decode_gemm(fn, w["output.weight"], d["logits"], out_fp32=True)
        else:
            ops.gemm(fn, w["output.weight"], d["logits_bf"])
            d["logits"].copy_(d["logits_bf"])  # .float() of the bf16 result (llama_ens5.py:530-531)
