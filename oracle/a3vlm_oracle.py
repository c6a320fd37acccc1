"""CPU oracle for the A3VLM hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may import
this file; nothing under a3vlm_b200/ does (the product path fails loudly without its CUDA library).

What it is: a plain-PyTorch restatement, runnable on CPU, of the reference's multimodal
forward / loss / greedy-generate / AdamW step for the path SURVEY.md section 8 scopes:
CLIP ViT-L/14 tower -> visual_proj -> start/end tags + splice -> Llama decoder -> logits.
Every function cites the reference file:line it follows (paths relative to /root/reference;
`M/` = model/accessory/, `LLM/` = model/accessory/model/LLM/).

PARITY PINNING.  The reference ships no tests, golden vectors or fixtures for this path (SURVEY
section 4) and cannot be imported as shipped (LLM/llama.py and configs/global_configs.py are absent
from the tree; fairscale / open_clip / apex / dacite are not installed).  Pins (tests/test_reference_pins.py,
tests/test_oracle.py):
  * OUTPUTS OF THE REFERENCE'S OWN SOURCE executed in the authoring container with the missing third-party
    names stubbed (tests/golden/make_reference_golden.py lists what ran unmodified and what was stubbed;
    vectors committed as tests/golden/reference_exec.pt): components.RMSNorm (bit-exact); llama_ens5
    Attention (SDPA branch, KV cache, right-aligned mask) / FeedForward / TransformerBlock /
    Transformer.forward / forward_inference in fp32 (2e-4) and run in bf16 (rounding points); the image
    splice, `image_words` logit slice and decode offset; clip_encode_image's glue over a tower built from
    torch.nn primitives; MetaModel.forward's trimmed shifted loss; MetaModel.generate's greedy loop
    (token-exact); MetaModel.sample_top_p under a fixed seed;
  * what that execution had to stub stays pinned independently: the RoPE helpers of the absent LLM/llama.py
    and the decoder as a whole against `transformers.LlamaForCausalLM` after the q/k row permutation the
    reference's own exporter documents (M/tools/convert_weights_to_hf.py:209-219); the open_clip tower
    against `transformers.CLIPVisionModel` (+ post_layernorm on all tokens), exact GELU (open_clip built
    with pretrained=None, LLM/llama_ens5.py:299) and QuickGELU (:297);
  * committed golden vectors of the whole multimodal path under tests/golden/tiny_e2e.pt
    (tests/golden/make_golden.py).
Not pinned by the reference itself (stated in DESIGN.md section 2): flash-attn (GPU-only; same function as
the SDPA branch that was executed), the ConvNeXt / DINOv2 / Q-Former encoders (out of scope), the bicubic
global view (torch's own F.interpolate in fp16, as the reference calls it).

Third-party arithmetic restated from published algorithms (un-vendored, unpinned -- the reference
has no requirements file): open_clip `VisionTransformer` (ViT-L-14), Meta Llama `precompute_freqs_cis
/ apply_rotary_emb / repeat_kv` (the absent LLM/llama.py), fairscale Column/RowParallelLinear and
ParallelEmbedding at model-parallel size 1 (== nn.Linear(bias=False) / nn.Embedding).

PRECISION.  `act_dtype=torch.bfloat16` reproduces the rounding points of the bf16 inference path
(`model.bfloat16()` + autocast, M/eval_affordance_v2.py:282,338) listed in SURVEY A.9: every Linear
output rounds to bf16, RMSNorm normalises in fp32, rounds, then multiplies by the weight, RoPE is
fp32 with one rounding, silu and the gate product round separately, residual adds are bf16, softmax
is fp32, LayerNorm is fp32 inside with a bf16 result, last-position logits are `.float()` of a
bf16 GEMM.  `act_dtype=torch.float32` is the unrounded mathematical model used for the HF pins.
The fine-tune path uses the same rounding points (bf16 residual stream); under CUDA autocast with
fp32 master weights the reference's stream would be promoted to fp32 by `torch.cat` with the fp32
LayerNorm output -- documented in DESIGN.md as a known, deliberate difference.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

bf16 = torch.bfloat16


# =================================================================================== configuration
@dataclass
class ModelArgs:
    """LLM/llama_ens5.py:33-50 (same field names and defaults)."""
    dim: int = 5120
    n_layers: int = 40
    n_heads: int = 40
    n_kv_heads: Optional[int] = None
    vocab_size: int = -1
    multiple_of: int = 256
    ffn_dim_multiplier: Optional[float] = None
    norm_eps: float = 1e-5
    rope_theta: float = 10000
    max_batch_size: int = 32
    max_seq_len: int = 2048
    rope_scaling: Optional[float] = None
    load_pretrained_visual_encoder: bool = False


@dataclass
class VisionArgs:
    """open_clip ViT-L-14 visual tower geometry (built at LLM/llama_ens5.py:296-301).  The reference
    hard-codes 224 px views; BASELINE.json's 336 / 448 single-view geometries need image_size as a
    parameter."""
    image_size: int = 336
    patch: int = 14
    width: int = 1024
    layers: int = 24
    heads: int = 16
    mlp: int = 4096
    quick_gelu: bool = False  # pretrained=None -> nn.GELU (exact erf); OpenAI weights -> QuickGELU
    ln_eps: float = 1e-5
    views: int = 1  # 5: the reference's tiling of a (2*image_size)^2 input (LLM/llama_ens5.py:383-385)

    @property
    def grid(self) -> int:
        return self.image_size // self.patch

    @property
    def tokens(self) -> int:
        return self.grid * self.grid + 1


def ffn_hidden_dim(dim: int, multiple_of: int, ffn_dim_multiplier: Optional[float]) -> int:
    """LLM/llama_ens5.py:196-200 with hidden_dim = 4*dim passed at :228."""
    hidden = int(2 * (4 * dim) / 3)
    if ffn_dim_multiplier is not None:
        hidden = int(ffn_dim_multiplier * hidden)
    return multiple_of * ((hidden + multiple_of - 1) // multiple_of)


# =================================================================================== RoPE helpers
def precompute_freqs_cis(dim: int, end: int, theta: float = 10000.0, scaling=None) -> torch.Tensor:
    """Absent LLM/llama.py; call site LLM/llama_ens5.py:271-274.  Restated from Meta's public Llama
    reference.  `scaling` other than None is unverifiable from the tree -> refused."""
    if scaling is not None:
        raise NotImplementedError("rope_scaling: direction cannot be verified from the reference tree")
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
    t = torch.arange(end).float()
    freqs = torch.outer(t, freqs)
    return torch.polar(torch.ones_like(freqs), freqs)  # complex64 (end, dim/2)


def apply_rotary_emb(xq: torch.Tensor, xk: torch.Tensor, freqs_cis: torch.Tensor):
    """Absent LLM/llama.py; call site LLM/llama_ens5.py:118.  Interleaved pairs (x[2j], x[2j+1]),
    fp32 math, one rounding back to the input dtype."""
    xq_ = torch.view_as_complex(xq.float().reshape(*xq.shape[:-1], -1, 2))
    xk_ = torch.view_as_complex(xk.float().reshape(*xk.shape[:-1], -1, 2))
    f = freqs_cis.view(1, xq_.shape[1], 1, xq_.shape[-1])  # reshape_for_broadcast
    xq_out = torch.view_as_real(xq_ * f).flatten(3)
    xk_out = torch.view_as_real(xk_ * f).flatten(3)
    return xq_out.type_as(xq), xk_out.type_as(xk)


def repeat_kv(x: torch.Tensor, n_rep: int) -> torch.Tensor:
    """Absent LLM/llama.py; call site LLM/llama_ens5.py:152-153."""
    if n_rep == 1:
        return x
    b, s, h, d = x.shape
    return x[:, :, :, None, :].expand(b, s, h, n_rep, d).reshape(b, s, h * n_rep, d)


def make_causal_mask(q_len: int, kv_len: int) -> torch.Tensor:
    """LLM/llama_ens5.py:181-185: right-aligned causal boolean mask (True = visible)."""
    q_idx = torch.arange(q_len) - q_len
    kv_idx = torch.arange(kv_len) - kv_len
    return q_idx.view(-1, 1) >= kv_idx.view(1, -1)


# =================================================================================== primitive ops
def linear(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], act_dtype, native=False):
    """nn.Linear under autocast: bf16 operands, fp32 accumulate, ONE rounding of the result
    (bias added before the rounding, as cuBLAS's epilogue does)."""
    if act_dtype == torch.float32:
        return F.linear(x.float(), w.float(), None if b is None else b.float())
    if native:  # the CPU-baseline timing path: let oneDNN/AMX run the bf16 GEMM directly
        return F.linear(x.to(bf16), w.to(bf16), None if b is None else b.to(bf16))
    return F.linear(x.to(bf16).float(), w.to(bf16).float(), None if b is None else b.to(bf16).float()).to(bf16)


def rmsnorm(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    """M/model/components.py:39-53: `_norm(x.float()).type_as(x) * weight`."""
    xf = x.float()
    out = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)).type_as(x)
    return out * weight.to(x.dtype)


def layernorm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float) -> torch.Tensor:
    """open_clip LayerNorm subclass / nn.LayerNorm: fp32 statistics, result cast back to x.dtype."""
    return F.layer_norm(x.float(), (x.shape[-1],), w.float(), b.float(), eps).to(x.dtype)


def attention_core(q, k, v, mask: Optional[torch.Tensor], scale: float) -> torch.Tensor:
    """Exact softmax attention == both branches of LLM/llama_ens5.py:142-167 (flash_attn_func and
    F.scaled_dot_product_attention compute the same function).  q (B,Tq,H,d), k/v (B,Tk,H,d);
    fp32 scores / softmax, output in q.dtype."""
    qf, kf, vf = q.float().transpose(1, 2), k.float().transpose(1, 2), v.float().transpose(1, 2)
    s = torch.matmul(qf, kf.transpose(-1, -2)) * scale
    if mask is not None:
        s = s.masked_fill(~mask, float("-inf"))
    p = torch.softmax(s, dim=-1)
    return torch.matmul(p, vf).transpose(1, 2).to(q.dtype)


def gelu(x: torch.Tensor, quick: bool) -> torch.Tensor:
    if quick:
        return (x.float() * torch.sigmoid(1.702 * x.float())).to(x.dtype)
    return F.gelu(x.float()).to(x.dtype)


# =================================================================================== parameters
def init_params(args: ModelArgs, vis: Optional[VisionArgs], seed: int = 2, dtype=bf16,
                init: str = "normal", output_scale: float = 1.0, rgbd: bool = False) -> Dict[str, torch.Tensor]:
    """Random-init state dict with the reference's parameter names and shapes (SURVEY 8b;
    M/tools/convert_weights_to_hf.py:101-115; M/util/param_group.py:73-93; LLM/llama_ens5.py:330-339).
    init="normal": N(0, 0.02) linears/embeddings (bench default); "kaiming": the
    reference's kaiming_uniform_(a=sqrt(5)) (LLM/llama_ens5.py:28).  Norm weights = 1."""
    g = torch.Generator().manual_seed(seed)
    D, V = args.dim, args.vocab_size
    Fh = ffn_hidden_dim(D, args.multiple_of, args.ffn_dim_multiplier)
    n_kv = args.n_heads if args.n_kv_heads is None else args.n_kv_heads
    hd = D // args.n_heads

    def lin(out_f, in_f, std=0.02):
        if init == "kaiming":
            bound = 1.0 / math.sqrt(in_f)
            return ((torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound).to(dtype)
        return (torch.randn(out_f, in_f, generator=g) * std).to(dtype)

    sd: Dict[str, torch.Tensor] = {}
    sd["tok_embeddings.weight"] = lin(V, D)
    for i in range(args.n_layers):
        p = f"layers.{i}."
        sd[p + "attention.wq.weight"] = lin(D, D)
        sd[p + "attention.wk.weight"] = lin(n_kv * hd, D)
        sd[p + "attention.wv.weight"] = lin(n_kv * hd, D)
        sd[p + "attention.wo.weight"] = lin(D, D)
        sd[p + "feed_forward.w1.weight"] = lin(Fh, D)
        sd[p + "feed_forward.w2.weight"] = lin(D, Fh)
        sd[p + "feed_forward.w3.weight"] = lin(Fh, D)
        sd[p + "attention_norm.weight"] = torch.ones(D, dtype=dtype)
        sd[p + "ffn_norm.weight"] = torch.ones(D, dtype=dtype)
    sd["norm.weight"] = torch.ones(D, dtype=dtype)
    sd["output.weight"] = (lin(V, D).float() * output_scale).to(dtype)
    if vis is not None:
        Wv, c = vis.width, "clip.visual."
        sc = Wv ** -0.5
        sd[c + "conv1.weight"] = (torch.randn(Wv, 3, vis.patch, vis.patch, generator=g) * 0.02).to(dtype)
        sd[c + "class_embedding"] = (torch.randn(Wv, generator=g) * sc).to(dtype)
        sd[c + "positional_embedding"] = (torch.randn(vis.tokens, Wv, generator=g) * sc).to(dtype)
        for n in ("ln_pre", "ln_post"):
            sd[c + n + ".weight"] = torch.ones(Wv, dtype=dtype)
            sd[c + n + ".bias"] = torch.zeros(Wv, dtype=dtype)
        for i in range(vis.layers):
            r = f"{c}transformer.resblocks.{i}."
            for n in ("ln_1", "ln_2"):
                sd[r + n + ".weight"] = torch.ones(Wv, dtype=dtype)
                sd[r + n + ".bias"] = torch.zeros(Wv, dtype=dtype)
            sd[r + "attn.in_proj_weight"] = lin(3 * Wv, Wv)
            sd[r + "attn.in_proj_bias"] = (torch.randn(3 * Wv, generator=g) * 0.01).to(dtype)
            sd[r + "attn.out_proj.weight"] = lin(Wv, Wv)
            sd[r + "attn.out_proj.bias"] = (torch.randn(Wv, generator=g) * 0.01).to(dtype)
            sd[r + "mlp.c_fc.weight"] = lin(vis.mlp, Wv)
            sd[r + "mlp.c_fc.bias"] = (torch.randn(vis.mlp, generator=g) * 0.01).to(dtype)
            sd[r + "mlp.c_proj.weight"] = lin(Wv, vis.mlp)
            sd[r + "mlp.c_proj.bias"] = (torch.randn(Wv, generator=g) * 0.01).to(dtype)
        sd["visual_proj.0.weight"] = lin(D, Wv)
        sd["visual_proj.0.bias"] = (torch.randn(D, generator=g) * 0.01).to(dtype)
        sd["visual_proj.1.weight"] = torch.ones(D, dtype=dtype)
        sd["visual_proj.1.bias"] = torch.zeros(D, dtype=dtype)
        sd["start_img"] = torch.rand(1, 1, D, generator=g).to(dtype)  # LLM/llama_ens5.py:338-339
        sd["end_img"] = torch.rand(1, 1, D, generator=g).to(dtype)
        if rgbd:  # LLM/llama_ens5_2images.py:343-344 (drawn last: seeds of the RGB-only build are unchanged)
            sd["start_depth_img"] = torch.rand(1, 1, D, generator=g).to(dtype)
            sd["end_depth_img"] = torch.rand(1, 1, D, generator=g).to(dtype)
    return sd


# =================================================================================== the model
class OracleTransformer:
    """Restates `Transformer` of LLM/llama_ens5.py:252-539 narrowed to CLIP ViT + visual_proj + Llama
    (ConvNeXt / DINOv2 / Q-Former branches are out of scope), single-view."""

    def __init__(self, args: ModelArgs, vis: Optional[VisionArgs], params: Dict[str, torch.Tensor],
                 act_dtype=bf16, native_linear: bool = False, rgbd: bool = False, stream_dtype=None):
        """rgbd=True restates LLM/llama_ens5_2images.py: a second (depth) image stream with its own
        start / end tags, image_words doubled (:335-344).
        stream_dtype=torch.float32 (with act_dtype bf16) restates the reference's TRAINING numerics under CUDA
        autocast with fp32 master weights (M/util/tensor_type.py:60-66, M/engine_finetune.py:44-49): embedding rows
        and the nn.LayerNorm output of visual_proj are fp32, so the residual stream (`h = x + linear(...)`, fp32 +
        bf16 -> fp32) and the RMSNorm outputs stay fp32, while every Linear still runs on bf16 operands and rounds
        its output to bf16.  Default (None = act_dtype): the bf16 stream of the inference path, which
        is what the CUDA path keeps in training too."""
        self.args, self.vis, self.p = args, vis, params
        self.rgbd = rgbd
        self.act_dtype, self.native = act_dtype, native_linear
        self.stream_dtype = act_dtype if stream_dtype is None else stream_dtype
        self.head_dim = args.dim // args.n_heads
        self.n_kv_heads = args.n_heads if args.n_kv_heads is None else args.n_kv_heads
        self.freqs_cis = precompute_freqs_cis(self.head_dim, args.max_seq_len * 2, args.rope_theta,
                                              args.rope_scaling)  # LLM/llama_ens5.py:271-274
        # per view: start + N + end (LLM/llama_ens5.py:335 is (32 + 257 + 2) * 5 with the Q-Former tokens)
        self.image_words = vis.views * (vis.tokens + 2) if vis is not None else 0
        self.visual_image_words = self.image_words  # LLM/llama_ens5_2images.py:336
        if rgbd:
            self.image_words *= 2  # :335
        self.image_size = (vis.image_size * (2 if vis.views == 5 else 1)) if vis is not None else None
        self.cache_image_words = 0
        self.k_cache: List[Optional[torch.Tensor]] = [None] * args.n_layers
        self.v_cache: List[Optional[torch.Tensor]] = [None] * args.n_layers

    # ---- helpers
    def _lin(self, x, name, bias=False):
        return linear(x, self.p[name + ".weight"], self.p[name + ".bias"] if bias else None,
                      self.act_dtype, self.native)

    def _cast(self, t):
        return t.to(self.act_dtype)

    # ---- vision: LLM/llama_ens5.py:351-375 (clip_encode_image), open_clip ResidualAttentionBlock
    @torch.no_grad()
    def clip_encode_image(self, image: torch.Tensor) -> torch.Tensor:
        v, c = self.vis, "clip.visual."
        B = image.shape[0]
        x = self._cast(image)
        # conv1 (k14 s14, no bias) == GEMM over (c, ky, kx)-flattened patches  (:354-357)
        patches = F.unfold(x.float(), v.patch, stride=v.patch).transpose(1, 2)  # (B, g*g, 3*P*P)
        w = self.p[c + "conv1.weight"].reshape(v.width, -1)
        x = linear(self._cast(patches), w, None, self.act_dtype, self.native)
        cls = self._cast(self.p[c + "class_embedding"]).expand(B, 1, v.width)
        x = torch.cat([cls, x], dim=1)  # (:358-361)
        x = x + self._cast(self.p[c + "positional_embedding"])  # (:362)
        x = layernorm(x, self.p[c + "ln_pre.weight"], self.p[c + "ln_pre.bias"], v.ln_eps)  # (:363)
        hd = v.width // v.heads
        for i in range(v.layers):  # (:366) transformer of ResidualAttentionBlocks, pre-LN
            r = f"{c}transformer.resblocks.{i}."
            y = layernorm(x, self.p[r + "ln_1.weight"], self.p[r + "ln_1.bias"], v.ln_eps)
            qkv = linear(y, self.p[r + "attn.in_proj_weight"], self.p[r + "attn.in_proj_bias"],
                         self.act_dtype, self.native)
            q, k, vv = qkv.split(v.width, dim=-1)  # packed rows: [0:W]=q, [W:2W]=k, [2W:3W]=v
            N = x.shape[1]
            o = attention_core(q.reshape(B, N, v.heads, hd), k.reshape(B, N, v.heads, hd),
                               vv.reshape(B, N, v.heads, hd), None, 1.0 / math.sqrt(hd))
            x = x + self._lin(o.reshape(B, N, v.width), r + "attn.out_proj", bias=True)
            y = layernorm(x, self.p[r + "ln_2.weight"], self.p[r + "ln_2.bias"], v.ln_eps)
            y = gelu(self._lin(y, r + "mlp.c_fc", bias=True), v.quick_gelu)
            x = x + self._lin(y, r + "mlp.c_proj", bias=True)
        # ln_post on ALL tokens, no 768-d projection (:370-373)
        return layernorm(x, self.p[c + "ln_post.weight"], self.p[c + "ln_post.bias"], v.ln_eps)

    def encode_image(self, image: torch.Tensor) -> torch.Tensor:
        """LLM/llama_ens5.py:377-458 narrowed to the CLIP branch + visual_proj (:330-333,451):
        Linear(C_in -> D) + LayerNorm(D); C_in = 1024 in the ViT-only build."""
        if self.vis.views == 5:
            # :383-385 -- global view = bicubic down-sample computed in fp16, then the four quadrants,
            # concatenated on the batch dimension (view-major)
            V = self.vis.image_size
            image_g = F.interpolate(image.half(), size=(V, V), mode="bicubic").to(image)
            parts = [image[..., :V, :V], image[..., :V, V:], image[..., V:, :V], image[..., V:, V:]]
            image = torch.cat([image_g] + parts, dim=0)
        feats = self.clip_encode_image(image)
        y = self._lin(feats, "visual_proj.0", bias=True)
        return layernorm(y.to(self.stream_dtype), self.p["visual_proj.1.weight"], self.p["visual_proj.1.bias"], 1e-5)

    def _wrap_views(self, image: torch.Tensor, start: str, end: str) -> torch.Tensor:
        B = image.shape[0]
        views = list(torch.chunk(self.encode_image(image), self.vis.views))  # :457 one (B, N, D) per view
        st, en = self.p[start].to(self.stream_dtype).expand(B, -1, -1), self.p[end].to(self.stream_dtype).expand(B, -1, -1)
        toks = torch.cat([torch.cat((st, v, en), dim=1) for v in views], dim=1)  # :471-477
        assert toks.shape[1] == self.visual_image_words, f"{toks.shape[1]} v.s. {self.visual_image_words}"  # :478
        return toks

    def _splice(self, h: torch.Tensor, image: torch.Tensor, depth: Optional[torch.Tensor] = None) -> torch.Tensor:
        """LLM/llama_ens5.py:469-479 / :500-511: BOS | start_img | image tokens | end_img | text.
        With `depth` (LLM/llama_ens5_2images.py:487-501,531-544): the depth views, wrapped in
        start_depth_img / end_depth_img, follow the RGB views."""
        toks = self._wrap_views(image, "start_img", "end_img")
        if depth is not None:
            toks = torch.cat((toks, self._wrap_views(depth, "start_depth_img", "end_depth_img")), dim=1)
        return torch.cat((h[:, :1], toks, h[:, 1:]), dim=1).to(h.dtype)

    # ---- decoder block: LLM/llama_ens5.py:97-169 (Attention), :213-217 (FFN), :237-249 (Block)
    def _block(self, i: int, x: torch.Tensor, start_pos: int, freqs_cis, causal: bool, use_cache: bool):
        p = f"layers.{i}."
        B, L, D = x.shape
        H, Hkv, hd = self.args.n_heads, self.n_kv_heads, self.head_dim
        xn = rmsnorm(x, self.p[p + "attention_norm.weight"], self.args.norm_eps)
        xq = self._lin(xn, p + "attention.wq").view(B, L, H, hd)
        xk = self._lin(xn, p + "attention.wk").view(B, L, Hkv, hd)
        xv = self._lin(xn, p + "attention.wv").view(B, L, Hkv, hd)
        xq, xk = apply_rotary_emb(xq, xk, freqs_cis)
        if use_cache:  # :121-129
            self.k_cache[i][:B, start_pos:start_pos + L] = xk
            self.v_cache[i][:B, start_pos:start_pos + L] = xv
            keys, values = self.k_cache[i][:B, :start_pos + L], self.v_cache[i][:B, :start_pos + L]
        else:
            keys, values = xk, xv
        keys, values = repeat_kv(keys, H // Hkv), repeat_kv(values, H // Hkv)
        mask = make_causal_mask(L, keys.shape[1]) if causal else None  # :160-161
        o = attention_core(xq, keys, values, mask, 1.0 / math.sqrt(hd)).reshape(B, L, D)
        h = x + self._lin(o, p + "attention.wo")  # :241
        hn = rmsnorm(h, self.p[p + "ffn_norm.weight"], self.args.norm_eps)
        g, u = self._lin(hn, p + "feed_forward.w1"), self._lin(hn, p + "feed_forward.w3")
        y = F.silu(g) * u  # :213-214: silu rounds, the product rounds
        return h + self._lin(y, p + "feed_forward.w2")  # :238

    def _embed(self, tokens: torch.Tensor) -> torch.Tensor:
        return F.embedding(tokens, self.p["tok_embeddings.weight"]).to(self.stream_dtype)

    # ---- training forward: LLM/llama_ens5.py:461-487
    def forward(self, examples: torch.Tensor, image: Optional[torch.Tensor] = None) -> torch.Tensor:
        self.k_cache = [None] * self.args.n_layers  # :462 training always disables the kv cache
        self.v_cache = [None] * self.args.n_layers
        h = self._embed(examples)
        image_words = 0
        if image is not None:
            h = self._splice(h, image)
            # the training forward only ever sees the RGB stream, also in the 2images build
            # (M/model/meta.py:251 calls self.llma(examples, images) without the depth images)
            image_words = self.visual_image_words
        L = h.shape[1]
        freqs = self.freqs_cis[:L]
        for i in range(self.args.n_layers):
            h = self._block(i, h, 0, freqs, True, False)
        h = rmsnorm(h, self.p["norm.weight"], self.args.norm_eps)
        return self._lin(h[:, image_words:, :], "output")  # :486 text positions only

    # ---- inference forward with KV cache: LLM/llama_ens5.py:490-531
    @torch.no_grad()
    def forward_inference(self, tokens: torch.Tensor, start_pos: int, image=None, depth_images=None) -> torch.Tensor:
        B, L = tokens.shape
        if self.rgbd and (image is None or depth_images is None):
            image = depth_images = None  # LLM/llama_ens5_2images.py:518: both streams or neither
        if not self.rgbd:
            depth_images = None
        a = self.args
        if start_pos == 0:  # :493-494, :533-539
            shape = (B, a.max_seq_len, self.n_kv_heads, self.head_dim)
            for i in range(a.n_layers):
                if self.k_cache[i] is None or self.k_cache[i].shape != shape:
                    self.k_cache[i] = torch.zeros(shape, dtype=self.act_dtype)
                    self.v_cache[i] = torch.zeros(shape, dtype=self.act_dtype)
        h = self._embed(tokens)
        if image is not None:
            assert start_pos == 0  # :499
            h = self._splice(h, image, depth_images)
            self.cache_image_words = self.image_words  # :546-547 (both streams) / llama_ens5.py:508-509
            L = h.shape[1]
            freqs = self.freqs_cis[0:L]
        elif start_pos == 0:
            self.cache_image_words = 0
            freqs = self.freqs_cis[0:L]
        else:
            start_pos = start_pos + self.cache_image_words  # :520
            freqs = self.freqs_cis[start_pos:start_pos + L]
        causal = L != 1  # :525
        for i in range(a.n_layers):
            h = self._block(i, h, start_pos, freqs, causal, True)
        h = rmsnorm(h, self.p["norm.weight"], a.norm_eps)
        return self._lin(h[:, -1, :], "output").float()  # :529-531


# =================================================================================== MetaModel pieces
def trim_trailing_pad(examples: torch.Tensor, labels: torch.Tensor):
    """M/model/meta.py:235-249: drop trailing columns whose labels are all 0; keep 3 if all are."""
    non_zero = torch.count_nonzero(labels, dim=0)
    pos = non_zero.shape[0] - 1
    while pos >= 0 and non_zero[pos] == 0:
        pos -= 1
    if pos == -1:
        pos = 2
    return examples[:, :pos + 1], labels[:, :pos + 1]


def meta_forward_loss(model: OracleTransformer, examples, labels, images=None) -> torch.Tensor:
    """M/model/meta.py:234-263: trim, forward, shift, CrossEntropyLoss(ignore_index=0) (mean)."""
    examples, labels = trim_trailing_pad(examples, labels)
    output = model.forward(examples, images)
    output = output[:, :-1, :]
    labels = labels[:, 1:]
    if labels.sum() == 0:
        return output.float().mean() * 0
    V = output.shape[-1]
    return F.cross_entropy(output.reshape(-1, V).float(), labels.flatten(), ignore_index=0)


def sample_top_p(probs: torch.Tensor, p: float, generator=None) -> torch.Tensor:
    """M/model/meta.py:568-583."""
    probs_sort, probs_idx = torch.sort(probs, dim=-1, descending=True)
    probs_sum = torch.cumsum(probs_sort, dim=-1)
    mask = probs_sum - probs_sort > p
    probs_sort[mask] = 0.0
    probs_sort.div_(probs_sort.sum(dim=-1, keepdim=True))
    next_token = torch.multinomial(probs_sort, num_samples=1, generator=generator)
    return torch.gather(probs_idx, -1, next_token)


def sample_top_p_u(probs: torch.Tensor, p: float, u: torch.Tensor):
    """M/model/meta.py:568-583 with the draw made explicit: torch.multinomial samples index j of the
    renormalised, masked, sorted distribution with probability probs_sort[j]; here that draw is the
    inverse CDF at a caller-supplied uniform u in [0,1) (the reference's random stream cannot be
    reproduced across devices, its distribution can).  Ties are ordered by token id (stable sort).
    Returns (token ids (B,), cdf_lo (B,), cdf_hi (B,)): the chosen token's interval of the CDF."""
    probs_sort, probs_idx = torch.sort(probs, dim=-1, descending=True, stable=True)
    probs_sum = torch.cumsum(probs_sort, dim=-1)
    mask = probs_sum - probs_sort > p
    probs_sort = probs_sort.masked_fill(mask, 0.0)
    probs_sort = probs_sort / probs_sort.sum(dim=-1, keepdim=True)
    cdf = torch.cumsum(probs_sort, dim=-1)
    j = (cdf > u.reshape(-1, 1)).float().argmax(dim=-1, keepdim=True)  # first position with cdf > u
    hi = torch.gather(cdf, -1, j).reshape(-1)
    lo = torch.where(j.reshape(-1) > 0, torch.gather(cdf, -1, (j - 1).clamp(min=0)).reshape(-1), torch.zeros_like(hi))
    return torch.gather(probs_idx, -1, j).reshape(-1), lo, hi


def top_p_token_interval(probs: torch.Tensor, p: float, token: torch.Tensor):
    """CDF interval [lo, hi) that `sample_top_p_u` assigns to `token` (B,) -- (0, 0) when the token is
    outside the nucleus.  Lets a test accept any token whose interval contains u up to float rounding."""
    probs_sort, probs_idx = torch.sort(probs, dim=-1, descending=True, stable=True)
    probs_sum = torch.cumsum(probs_sort, dim=-1)
    probs_sort = probs_sort.masked_fill(probs_sum - probs_sort > p, 0.0)
    probs_sort = probs_sort / probs_sort.sum(dim=-1, keepdim=True)
    cdf = torch.cumsum(probs_sort, dim=-1)
    j = (probs_idx == token.reshape(-1, 1)).float().argmax(dim=-1, keepdim=True)
    hi = torch.gather(cdf, -1, j).reshape(-1)
    lo = hi - torch.gather(probs_sort, -1, j).reshape(-1)
    return lo, hi


@torch.no_grad()
def generate_tokens(model: OracleTransformer, prompt_tokens: Sequence[Sequence[int]], images=None,
                    max_gen_len: int = 512, temperature: float = 0.0, top_p: float = 0.95,
                    stop_sequences: Iterable[Sequence[int]] = (), eos_id: int = 2,
                    max_batch_size: Optional[int] = None, depth_images=None, uniforms: Optional[torch.Tensor] = None):
    """Token-level restatement of MetaModel.generate, M/model/meta.py:379-485 (the tokenizer
    encode / decode at :417-418 and :481-484 stay outside: they are CPU string work kept as-is).
    Returns (per-row generated token lists, tokens tensor, stop_pos)."""
    bsz = len(prompt_tokens)
    a = model.args
    assert bsz <= (a.max_batch_size if max_batch_size is None else max_batch_size)  # :415
    prompt_tokens = [list(t) for t in prompt_tokens]
    min_prompt = min(len(t) for t in prompt_tokens)
    max_prompt = max(len(t) for t in prompt_tokens)
    max_seq_len = a.max_seq_len
    if images is not None:
        max_seq_len -= model.image_words  # :423-425
    total_len = min(max_seq_len, max_gen_len + max_prompt)  # :427
    for k, t in enumerate(prompt_tokens):
        prompt_tokens[k] = t[-(max_seq_len - max_gen_len):]  # :428-429 left truncation
    tokens = torch.full((bsz, total_len), 0, dtype=torch.long)
    input_text_mask = torch.full((bsz, total_len), False)
    for k, t in enumerate(prompt_tokens):
        tokens[k, :len(t)] = torch.tensor(t).long()
        input_text_mask[k, :len(t)] = True
    start_pos, prev_pos = min_prompt, 0
    l_stop = [torch.tensor([eos_id])] + [torch.tensor(list(s)) for s in stop_sequences]  # :439-442
    stopped = torch.zeros(bsz, dtype=torch.bool)
    stop_pos = torch.full((bsz,), start_pos + 1)
    for cur_pos in range(start_pos, total_len):
        if depth_images is not None:  # :447-451
            if prev_pos == 0:
                logits = model.forward_inference(tokens[:, prev_pos:cur_pos], prev_pos, images, depth_images).float()
            else:
                logits = model.forward_inference(tokens[:, prev_pos:cur_pos], prev_pos).float()
        else:
            logits = model.forward_inference(tokens[:, prev_pos:cur_pos], prev_pos,
                                             images if prev_pos == 0 else None).float()
        if temperature > 0 and uniforms is not None:  # uniforms (total_len, bsz): row cur_pos drives this step
            next_token = sample_top_p_u(torch.softmax(logits / temperature, dim=-1), top_p, uniforms[cur_pos])[0]
        elif temperature > 0:
            next_token = sample_top_p(torch.softmax(logits / temperature, dim=-1), top_p)
        else:
            next_token = torch.argmax(logits, dim=-1)
        next_token = next_token.reshape(-1)
        next_token = torch.where(input_text_mask[:, cur_pos], tokens[:, cur_pos], next_token)  # :463-465
        tokens[:, cur_pos] = next_token
        stop_pos = torch.where(stopped, stop_pos, cur_pos + 1)
        for st in l_stop:  # :468-475
            if cur_pos + 1 - len(st) >= 0:
                cond1 = (tokens[:, cur_pos + 1 - len(st):cur_pos + 1] == st.unsqueeze(0)).all(dim=-1)
                cond2 = ~input_text_mask[:, cur_pos]
                new_stop = cond1 & cond2 & (~stopped)
                stop_pos = torch.where(new_stop, cur_pos + 1 - len(st), stop_pos)
                stopped = torch.logical_or(new_stop, stopped)
        if stopped.all():
            break
        prev_pos = cur_pos
    out = [tokens[i, len(prompt_tokens[i]):stop_pos[i].item()].tolist() for i in range(bsz)]
    return out, tokens, stop_pos


# =================================================================================== optimiser
def is_no_decay(name: str) -> bool:
    """M/util/misc.py:586-599 add_weight_decay: no decay for *.bias and *norm.weight."""
    return name.endswith(".bias") or name.endswith("norm.weight")


def clip_coef(total_norm: float, max_norm: float) -> float:
    """M/util/clip_grad.py:187-193."""
    return min(1.0, max_norm / (total_norm + 1e-6))


def adamw_step(master: Dict[str, torch.Tensor], grads: Dict[str, torch.Tensor],
               state: Dict[str, Tuple[torch.Tensor, torch.Tensor]], step: int, lr: float,
               weight_decay: float, max_norm: float, betas=(0.9, 0.95), eps=1e-8) -> float:
    """One optimiser step of the fine-tune loop (M/main_finetune.py:284-285 AdamW(betas=(0.9,0.95));
    M/util/misc.py:302-315 clip then step): global L2 norm over all grads in fp32, clip, decoupled
    AdamW on fp32 masters (M/util/tensor_type.py:60-66).  Returns the pre-clip grad norm."""
    total = math.sqrt(sum(float(g.float().pow(2).sum()) for g in grads.values()))
    coef = clip_coef(total, max_norm) if max_norm > 0 else 1.0
    b1, b2 = betas
    for n, p in master.items():
        g = grads[n].float() * coef
        m, v = state[n]
        wd = 0.0 if is_no_decay(n) else weight_decay
        p.mul_(1 - lr * wd)
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = v.sqrt() / math.sqrt(1 - b2 ** step) + eps
        p.addcdiv_(m, denom, value=-lr / (1 - b1 ** step))
    return total
