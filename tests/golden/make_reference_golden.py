"""Golden vectors produced by EXECUTING THE REFERENCE'S OWN SOURCE (read from /root/reference at
generation time only; nothing of it is copied here) -> tests/golden/reference_exec.pt.

    python tests/golden/make_reference_golden.py            # needs /root/reference (this container)

SURVEY 8c: the reference cannot be imported as shipped (fairscale / open_clip / apex / dacite are not
installed, `LLM/llama.py` and `configs/global_configs.py` are absent from the tree, the vision branch
pulls DINOv2 from the network).  What CAN be done is to import its modules with the missing third-party
names stubbed, and run its own classes / methods on CPU:

  executed unmodified (the thing being pinned)
    model/accessory/model/components.py        RMSNorm
    model/accessory/model/LLM/llama_ens5.py    Attention (SDPA branch + KV cache + right-aligned mask),
                                               FeedForward, TransformerBlock, Transformer.__init__ /
                                               forward / forward_inference / _allocate_kv_cache (text-only
                                               build), the image splice + logit slice of forward /
                                               forward_inference, clip_encode_image
    model/accessory/model/LLM/llama_ens5_2images.py  forward_inference with image + depth_images (RGB-D splice)
    model/accessory/model/meta.py              MetaModel.forward (pad trimming, shifted CE ignore_index 0),
                                               MetaModel.generate (force-feeding, stops, truncation),
                                               MetaModel.sample_top_p, compute_logits, evaluate_examples,
                                               stream_generate
    model/accessory/model/tokenizer.py         Tokenizer (SentencePiece branch) on tests/golden/toy_spm.model
    model/accessory/data/transform.py          get_transform (padded_resize / resized_center_crop / random_resized_crop)
    model/accessory/util/lr_sched.py           adjust_learning_rate, adjust_learning_rate_epoch
    model/accessory/util/misc.py               add_weight_decay (function source executed on its own)

  stubbed (stated here and in DESIGN.md section 2)
    fairscale ColumnParallelLinear / RowParallelLinear / ParallelEmbedding -> nn.Linear / nn.Embedding
        (what they are at model-parallel size 1), fs_init.get_model_parallel_* -> 1 / 0
    accessory.configs.global_configs (absent file)  -> USE_FLASH_ATTENTION = False (the SDPA branch;
        flash-attn needs a GPU and computes the same function)
    accessory.model.LLM.llama (absent file)         -> precompute_freqs_cis / apply_rotary_emb / repeat_kv
        from oracle/a3vlm_oracle.py (SURVEY A.5 restatement; pinned separately against HF Llama)
    open_clip (not installed)                       -> a ViT tower assembled from torch.nn primitives
        (nn.Conv2d, nn.LayerNorm, nn.MultiheadAttention seq-first, Linear-GELU-Linear) under the attribute
        names the reference reads (`clip.visual.conv1 / class_embedding / positional_embedding / ln_pre /
        transformer / ln_post`): pins the reference's glue in clip_encode_image and gives the oracle's
        block arithmetic a third independent implementation next to HF CLIP
    the four-encoder `encode_image` (ConvNeXt / DINOv2 / Q-Former need the network) -> replaced on the
        instance by a function returning given per-view token tensors, so the reference's OWN splice,
        `image_words` assert, logit slice and decode offset run on them
    dacite, accessory.util.{misc, tensor_parallel, tensor_type}, the tokenizer -> inert stubs (only
        MetaModel's methods are used, called on an object built with __new__); a toy tokenizer whose text is
        space-separated token ids; `Tensor.cuda()` -> identity (no GPU in this container)
"""
import importlib
import importlib.util
import os
import sys
import types
from types import SimpleNamespace

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/model"
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import a3vlm_oracle as O  # noqa: E402


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _DenseLinear(nn.Linear):
    """fairscale Column/RowParallelLinear at model-parallel size 1."""

    def __init__(self, in_features, out_features, bias=True, gather_output=True, input_is_parallel=False,
                 init_method=None, **kw):
        super().__init__(in_features, out_features, bias=bias)
        if init_method is not None:
            init_method(self.weight)


class _DenseEmbedding(nn.Embedding):
    def __init__(self, num_embeddings, embedding_dim, init_method=None, **kw):
        super().__init__(num_embeddings, embedding_dim)
        if init_method is not None:
            init_method(self.weight)


def install_stubs():
    fs, fsnn, fsmp = _mod("fairscale"), _mod("fairscale.nn"), _mod("fairscale.nn.model_parallel")
    init = _mod("fairscale.nn.model_parallel.initialize", get_model_parallel_world_size=lambda: 1,
                get_model_parallel_rank=lambda: 0, get_model_parallel_group=lambda: None,
                get_data_parallel_rank=lambda: 0, get_data_parallel_world_size=lambda: 1)
    layers = _mod("fairscale.nn.model_parallel.layers", ColumnParallelLinear=_DenseLinear,
                  RowParallelLinear=_DenseLinear, ParallelEmbedding=_DenseEmbedding)
    fs.nn, fsnn.model_parallel, fsmp.initialize, fsmp.layers = fsnn, fsmp, init, layers
    _mod("open_clip")
    _mod("dacite")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import accessory.configs  # namespace package of the reference tree

    accessory.configs.global_configs = _mod("accessory.configs.global_configs", USE_FLASH_ATTENTION=False)
    _mod("accessory.model.LLM.llama", precompute_freqs_cis=O.precompute_freqs_cis,
         apply_rotary_emb=O.apply_rotary_emb, repeat_kv=O.repeat_kv)
    util = _mod("accessory.util")
    util.misc = _mod("accessory.util.misc")
    util.tensor_parallel = _mod("accessory.util.tensor_parallel")
    util.tensor_type = _mod("accessory.util.tensor_type", default_tensor_type=None)
    _mod("accessory.model.tokenizer", Tokenizer=object, probe_tokenizer_path_from_pretrained=None)


class ToyTokenizer:
    """Text = space-separated token ids, so the reference's string-level `generate` can be driven and read
    back token by token."""
    bos_id, eos_id, n_words = 1, 2, 512

    def encode(self, s, bos, eos):
        t = [int(x) for x in s.split()]
        return ([self.bos_id] if bos else []) + t + ([self.eos_id] if eos else [])

    def encode_segment(self, s):
        return self.encode(s, False, False)

    def encode_wo_prefix_space(self, s):
        return self.encode(s, False, False)

    def decode(self, t):
        return " ".join(str(x) for x in t)


class TorchViT(nn.Module):
    """open_clip VisionTransformer (seq-first releases) assembled from torch.nn primitives, under the
    attribute names LLM/llama_ens5.py:354-370 reads."""

    class Block(nn.Module):
        def __init__(self, w, h, mlp, quick):
            super().__init__()
            self.ln_1, self.ln_2 = nn.LayerNorm(w), nn.LayerNorm(w)
            self.attn = nn.MultiheadAttention(w, h)  # seq-first (L, N, D)
            self.c_fc, self.c_proj, self.quick = nn.Linear(w, mlp), nn.Linear(mlp, w), quick

        def forward(self, x):
            y = self.ln_1(x)
            x = x + self.attn(y, y, y, need_weights=False)[0]
            y = self.c_fc(self.ln_2(x))
            y = y * torch.sigmoid(1.702 * y) if self.quick else nn.functional.gelu(y)
            return x + self.c_proj(y)

    def __init__(self, v: "O.VisionArgs", sd):
        super().__init__()
        c = "clip.visual."
        self.conv1 = nn.Conv2d(3, v.width, v.patch, v.patch, bias=False)
        self.class_embedding = nn.Parameter(sd[c + "class_embedding"].float())
        self.positional_embedding = nn.Parameter(sd[c + "positional_embedding"].float())
        self.ln_pre, self.ln_post = nn.LayerNorm(v.width), nn.LayerNorm(v.width)
        self.transformer = nn.Sequential(*[TorchViT.Block(v.width, v.heads, v.mlp, v.quick_gelu) for _ in range(v.layers)])
        with torch.no_grad():
            self.conv1.weight.copy_(sd[c + "conv1.weight"].float())
            for n in ("ln_pre", "ln_post"):
                getattr(self, n).weight.copy_(sd[c + n + ".weight"].float())
                getattr(self, n).bias.copy_(sd[c + n + ".bias"].float())
            for i, b in enumerate(self.transformer):
                r = f"{c}transformer.resblocks.{i}."
                for n in ("ln_1", "ln_2"):
                    getattr(b, n).weight.copy_(sd[r + n + ".weight"].float())
                    getattr(b, n).bias.copy_(sd[r + n + ".bias"].float())
                b.attn.in_proj_weight.copy_(sd[r + "attn.in_proj_weight"].float())
                b.attn.in_proj_bias.copy_(sd[r + "attn.in_proj_bias"].float())
                b.attn.out_proj.weight.copy_(sd[r + "attn.out_proj.weight"].float())
                b.attn.out_proj.bias.copy_(sd[r + "attn.out_proj.bias"].float())
                b.c_fc.weight.copy_(sd[r + "mlp.c_fc.weight"].float())
                b.c_fc.bias.copy_(sd[r + "mlp.c_fc.bias"].float())
                b.c_proj.weight.copy_(sd[r + "mlp.c_proj.weight"].float())
                b.c_proj.bias.copy_(sd[r + "mlp.c_proj.bias"].float())


def synthetic_image(w, h):
    """Deterministic RGB test card (gradients + a checker), identical wherever it is rebuilt."""
    import numpy as np
    from PIL import Image

    y, x = np.mgrid[0:h, 0:w]
    rgb = np.stack([(x * 255) // max(w - 1, 1), (y * 255) // max(h - 1, 1), ((x // 4 + y // 4) % 2) * 200 + 20], axis=-1)
    return Image.fromarray(rgb.astype(np.uint8), "RGB")


def tiny_cfg():
    a = O.ModelArgs(dim=256, n_layers=2, n_heads=2, vocab_size=512, multiple_of=64, max_seq_len=64, max_batch_size=4)
    v = O.VisionArgs(image_size=28, patch=14, width=128, layers=2, heads=2, mlp=256)
    return a, v


def build_reference_decoder(L, a, sd, dtype):
    """The reference's text-only Transformer with the oracle's weights."""
    args = L.ModelArgs(dim=a.dim, n_layers=a.n_layers, n_heads=a.n_heads, vocab_size=a.vocab_size,
                       multiple_of=a.multiple_of, norm_eps=a.norm_eps, max_seq_len=a.max_seq_len,
                       max_batch_size=a.max_batch_size)
    ref = L.Transformer(args, with_visual=False)
    own = {k: v_.float() for k, v_ in sd.items() if not k.startswith(("clip.", "visual_proj", "start_img", "end_img"))}
    res = ref.load_state_dict(own, strict=False)
    assert not res.unexpected_keys and not res.missing_keys, res
    return ref.to(dtype).eval()


def main():
    assert os.path.isdir(REF), "needs the reference tree at /root/reference"
    install_stubs()
    torch.manual_seed(0)
    L = importlib.import_module("accessory.model.LLM.llama_ens5")
    C = importlib.import_module("accessory.model.components")
    M = importlib.import_module("accessory.model.meta")
    a, v = tiny_cfg()
    sd = O.init_params(a, v, seed=17, output_scale=8.0)
    g = torch.Generator().manual_seed(3)
    out = {"seed": 17, "output_scale": 8.0}

    # ---- RMSNorm (components.py:10-53), fp32 and bf16 inputs, eps as the model sets it
    x = torch.randn(3, 7, a.dim, generator=g) * 3
    w = torch.rand(a.dim, generator=g) + 0.5
    rn = C.RMSNorm(a.dim, eps=a.norm_eps)
    with torch.no_grad():
        rn.weight.copy_(w)
        out["rms_x"], out["rms_w"] = x, w
        out["rms_fp32"] = rn(x)
        out["rms_bf16"] = rn.to(torch.bfloat16)(x.to(torch.bfloat16))

    # ---- decoder, text only: forward, prefill + decode through the KV cache, multi-token step on the cache
    tok = torch.randint(3, a.vocab_size, (2, 11), generator=g)
    tok[:, 0] = 1
    out["tokens"] = tok
    for name, dt in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
        ref = build_reference_decoder(L, a, sd, dt)
        with torch.no_grad():
            out[f"fwd_{name}"] = ref(tok).float()
        out[f"prefill_{name}"] = ref.forward_inference(tok[:, :6], 0).clone()
        out[f"decode1_{name}"] = ref.forward_inference(tok[:, 6:7], 6).clone()
        out[f"decode2_{name}"] = ref.forward_inference(tok[:, 7:8], 7).clone()
        out[f"chunk_{name}"] = ref.forward_inference(tok[:, 8:11], 8).clone()  # l > 1 against the cache: right-aligned mask

    # ---- image splice / logit slice / decode offset (llama_ens5.py:469-486, 500-520) on given view tokens
    ref = build_reference_decoder(L, a, sd, torch.float32)
    n_views, n_tok = 5, 5  # = VisionArgs(image_size=28, views=5): 2 x 2 patches + cls
    views = [torch.randn(2, n_tok, a.dim, generator=g) for _ in range(n_views)]
    ref.start_img = nn.Parameter(sd["start_img"].float())
    ref.end_img = nn.Parameter(sd["end_img"].float())
    ref.image_words = n_views * (n_tok + 2)
    ref.encode_image = lambda image: [t.clone() for t in views]
    fake_image = torch.zeros(2, 3, 4, 4)
    with torch.no_grad():
        out["img_views"] = torch.stack(views)
        out["img_fwd_fp32"] = ref(tok, fake_image).float()
    out["img_prefill_fp32"] = ref.forward_inference(tok[:, :5], 0, fake_image).clone()
    out["img_cache_image_words"] = int(ref.cache_image_words)
    out["img_decode_fp32"] = ref.forward_inference(tok[:, 5:6], 5).clone()

    # ---- RGB-D build (LLM/llama_ens5_2images.py:511-547): both streams spliced with their own tags; a lone RGB
    # image is ignored at inference (:518)
    L2 = importlib.import_module("accessory.model.LLM.llama_ens5_2images")
    a_rgbd = O.ModelArgs(**{**vars(a), "max_seq_len": 128})  # 1 + 70 image words + text
    ref2 = build_reference_decoder(L2, a_rgbd, sd, torch.float32)
    g2 = torch.Generator().manual_seed(9)
    depth_views = [torch.randn(2, n_tok, a.dim, generator=g2) for _ in range(n_views)]
    tags = {k: torch.rand(1, 1, a.dim, generator=g2) for k in ("start_depth_img", "end_depth_img")}
    ref2.start_img, ref2.end_img = nn.Parameter(sd["start_img"].float()), nn.Parameter(sd["end_img"].float())
    ref2.start_depth_img, ref2.end_depth_img = nn.Parameter(tags["start_depth_img"]), nn.Parameter(tags["end_depth_img"])
    ref2.visual_image_words = n_views * (n_tok + 2)
    ref2.image_words = 2 * ref2.visual_image_words
    rgb_marker, depth_marker = torch.zeros(2, 3, 4, 4), torch.ones(2, 3, 4, 4)
    ref2.encode_image = lambda image: [t.clone() for t in (depth_views if bool(image.flatten()[0] == 1) else views)]
    out["rgbd_depth_views"] = torch.stack(depth_views)
    out["rgbd_tags"] = tags
    out["rgbd_prefill_fp32"] = ref2.forward_inference(tok[:, :5], 0, rgb_marker, depth_marker).clone()
    out["rgbd_cache_image_words"] = int(ref2.cache_image_words)
    out["rgbd_decode_fp32"] = ref2.forward_inference(tok[:, 5:6], 5).clone()
    out["rgbd_lone_rgb_prefill_fp32"] = ref2.forward_inference(tok[:, :5], 0, rgb_marker).clone()
    out["rgbd_lone_cache_image_words"] = int(ref2.cache_image_words)

    # ---- clip_encode_image glue (llama_ens5.py:351-375) over a torch-primitive tower
    img = torch.randn(2, 3, v.image_size, v.image_size, generator=g)
    holder = SimpleNamespace(clip=SimpleNamespace(visual=TorchViT(v, sd).eval()))
    out["vit_image"] = img
    out["vit_tokens_fp32"] = L.Transformer.clip_encode_image(holder, img)

    # ---- MetaModel.forward (meta.py:234-263): pad trimming + shifted CE(ignore_index=0), text-only inner model
    meta = M.MetaModel.__new__(M.MetaModel)
    nn.Module.__init__(meta)
    meta.llma = build_reference_decoder(L, a, sd, torch.float32)
    meta.tokenizer = ToyTokenizer()
    meta.criterion = torch.nn.CrossEntropyLoss(ignore_index=0)
    ex = torch.randint(3, a.vocab_size, (3, 14), generator=g)
    ex[:, 0] = 1
    lab = ex.clone()
    lab[:, :5] = 0
    ex[:, 10:], lab[:, 10:] = 0, 0          # padded tail, trimmed by the reference
    ex[2, 8:], lab[2, 8:] = 0, 0            # ragged
    with torch.no_grad():
        loss, extra = meta(ex, lab)
    out["loss_examples"], out["loss_labels"], out["loss_fp32"] = ex, lab, loss.clone()
    assert extra == {}

    # ---- MetaModel.generate (meta.py:379-485), greedy, through the reference's own loop
    cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a_, **k_: self  # no GPU here; the loop itself is device-agnostic
    try:
        cases = [dict(prompts=["17 33 49", "100 101 102 103 104"], max_gen_len=8, stops=[]),
                 dict(prompts=["5 6", "9 10 11 12", "7"], max_gen_len=12, stops=["165 462", "247", "409"]),  # every row stops
                 dict(prompts=[" ".join(str(3 + i) for i in range(70)), "4"], max_gen_len=6, stops=[])]  # left truncation
        gens = []
        inner = meta.llma.forward_inference
        for c in cases:
            margins = []  # instrumentation only: the reference's top-2 logit gap at every call (tie detection)

            def recording(tokens, start_pos, image=None):
                lg = inner(tokens, start_pos, image)
                top2 = lg.float().topk(2, dim=-1).values
                margins.append(torch.stack([top2[:, 0] - top2[:, 1], top2[:, 0].abs()], dim=-1))
                return lg

            meta.llma.forward_inference = recording
            res = meta.generate(c["prompts"], None, None, c["max_gen_len"], 0.0, 0.95, c["stops"])
            gens.append(dict(c, generated=[[int(t) for t in r.split()] for r in res], margins=torch.stack(margins)))
        meta.llma.forward_inference = inner
        out["generate_cases"] = gens
        # ---- generate with images (budget max_seq_len - image_words, image only in the prefill) and with
        # sampling (temperature 0.8, top-p 0.9) under a fixed torch seed
        meta_img = M.MetaModel.__new__(M.MetaModel)
        nn.Module.__init__(meta_img)
        meta_img.llma, meta_img.tokenizer = ref, ToyTokenizer()   # `ref`: the decoder with the patched encode_image above
        res = meta_img.generate(["17 33 49", "100 101 102 103 104 105 106"], fake_image, None, 20, 0.0, 0.95, [])
        out["generate_image_case"] = dict(prompts=["17 33 49", "100 101 102 103 104 105 106"], max_gen_len=20,
                                          generated=[[int(t) for t in r.split()] for r in res])
        torch.manual_seed(4321)
        res = meta.generate(["17 33 49", "100 101 102 103 104"], None, None, 10, 0.8, 0.9, [])
        out["generate_sampled_case"] = dict(prompts=["17 33 49", "100 101 102 103 104"], max_gen_len=10, temperature=0.8,
                                            top_p=0.9, seed=4321, generated=[[int(t) for t in r.split()] for r in res])

        # ---- scoring API (meta.py:265-377): compute_logits / evaluate_examples on token lists and on text
        exs = [[1, 17, 33, 49, 80, 81], [1, 100, 101, 102], [1, 9, 10, 11, 12, 13, 14, 15]]
        ctx = [[1, 17, 33], [1], [1, 9, 10, 11, 12]]
        cl = meta.compute_logits(exs)
        ev_all = meta.evaluate_examples(exs)
        ev_ctx = meta.evaluate_examples(exs, ctx)
        ev_txt = meta.evaluate_examples(["17 33 49", "100 101"], ["17", "100"], None, True, True)
        strip = lambda r: {k: ([t.clone() for t in v_] if k == "non_context_logits" else list(v_)) for k, v_ in r.items()}
        out["score_cases"] = dict(examples=exs, contexts=ctx, compute_logits=[t.clone() for t in cl], all=strip(ev_all),
                                  ctx=strip(ev_ctx), txt=strip(ev_txt))
        # ---- stream_generate (meta.py:487-566): greedy, a stop symbol that cuts the text, and a length-bound run
        out["stream_cases"] = []
        for prompt, max_gen, stops_ in (("17 33 49", 6, ()), ("100 101 102 103 104", 8, ("483 483",)), ("5 6", 5, ("999",))):
            chunks = list(meta.stream_generate(prompt, None, max_gen, 0.0, 0.95, stops_))
            out["stream_cases"].append(dict(prompt=prompt, max_gen_len=max_gen, stops=stops_, chunks=chunks))
    finally:
        torch.Tensor.cuda = cuda

    # ---- MetaModel.sample_top_p (meta.py:568-583) under a fixed torch seed
    probs = torch.softmax(torch.randn(4, 97, generator=g) * 2, dim=-1)
    torch.manual_seed(1234)
    out["topp_probs"], out["topp_p"] = probs, 0.75
    out["topp_tokens_seed1234"] = meta.sample_top_p(probs.clone(), 0.75)

    # ---- host-side training arithmetic: util/lr_sched.py (imported as is) and misc.add_weight_decay (its
    # function source executed on its own: the rest of misc.py needs FSDP / fairscale at import time)
    import ast

    spec = importlib.util.spec_from_file_location("ref_lr_sched", os.path.join(REF, "accessory", "util", "lr_sched.py"))
    S = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(S)
    sched_args = SimpleNamespace(lr=2e-5, min_lr=1e-6, warmup_iters=50, lr_decay_iters=1000, warmup_epochs=0.03, epochs=3)
    opt = SimpleNamespace(param_groups=[{"lr": 0.0}, {"lr": 0.0, "lr_scale": 0.5}])
    its = [0, 1, 25, 50, 51, 400, 999, 1000, 1001, 5000]
    out["lr_args"] = vars(sched_args)
    out["lr_iter"] = [(it, S.adjust_learning_rate(opt, it, sched_args), opt.param_groups[1]["lr"]) for it in its]
    eps_ = [0.0, 0.01, 0.03, 0.5, 1.7, 3.0]
    out["lr_epoch"] = [(e, S.adjust_learning_rate_epoch(opt, e, sched_args)) for e in eps_]
    src = open(os.path.join(REF, "accessory", "util", "misc.py")).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "add_weight_decay")
    ns = {}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "misc.py:add_weight_decay", "exec"), ns)
    holder = nn.Module()
    names = [k for k in sd]  # every parameter name of the path, frozen ViT included (requires_grad False)
    params = {}
    for k in names:
        prm = nn.Parameter(torch.zeros(1), requires_grad=not k.startswith("clip."))
        params[k] = prm
    holder.named_parameters = lambda: list(params.items())
    groups = ns["add_weight_decay"](holder, 0.1)
    ids = {id(p_): k for k, p_ in params.items()}
    out["wd_no_decay"] = sorted(ids[id(p_)] for p_ in groups[0]["params"])
    out["wd_decay"] = sorted(ids[id(p_)] for p_ in groups[1]["params"])
    assert groups[0]["weight_decay"] == 0.0 and groups[1]["weight_decay"] == 0.1

    # ---- Tokenizer (model/tokenizer.py:15-131), the real class on a tiny committed SentencePiece model
    # (tests/golden/toy_spm.model, trained offline on a few sentences; identity normalisation keeps it ~2 KB)
    del sys.modules["accessory.model.tokenizer"]  # drop the inert stub used for meta.py: import the real module
    TK = importlib.import_module("accessory.model.tokenizer")
    rt = TK.Tokenizer(os.path.join(HERE, "toy_spm.model"))
    texts = ["Hi my darling", "the robot opens the drawer", " leading space", "###", "\n###", "rotate the door ### Assistant:", ""]
    out["tok_cases"] = {
        "n_words": rt.n_words, "bos_id": rt.bos_id, "eos_id": rt.eos_id, "need_space_before_segment": rt.need_space_before_segment,
        "encode": [(t, b, e, rt.encode(t, b, e)) for t in texts for b, e in ((True, False), (False, True), (False, False))],
        "encode_segment": [(t, rt.encode_segment(t)) for t in texts if t],
        "encode_wo_prefix_space": [(t, rt.encode_wo_prefix_space(t)) for t in texts if t.strip()],
        "decode": [(ids, rt.decode(ids)) for ids in ([85, 7, 69, 32], rt.encode("the robot opens the drawer", False, False))],
    }

    # ---- image pre-processing (data/transform.py:37-79) on synthetic PIL images
    spec = importlib.util.spec_from_file_location("ref_transform", os.path.join(REF, "accessory", "data", "transform.py"))
    TR = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(TR)
    out["transform_cases"] = []
    for (w_, h_) in ((50, 37), (31, 64), (40, 40)):
        pil = synthetic_image(w_, h_)
        for kind in ("padded_resize", "resized_center_crop", "random_resized_crop"):
            torch.manual_seed(7)
            out["transform_cases"].append((w_, h_, kind, 28, TR.get_transform(kind, 28)(pil)))

    path = os.path.join(HERE, "reference_exec.pt")
    torch.save(out, path)
    print("wrote", path, {k: (tuple(t.shape) if torch.is_tensor(t) else t) for k, t in out.items() if k != "generate_cases"})
    print("generate:", [c["generated"] for c in out["generate_cases"]])


if __name__ == "__main__":
    main()
