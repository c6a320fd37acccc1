"""The oracle (and, on a GPU, the CUDA path itself) against outputs of THE REFERENCE'S OWN SOURCE executed
in the authoring container -- tests/golden/reference_exec.pt, produced by
tests/golden/make_reference_golden.py (its docstring lists exactly what ran unmodified and what was
stubbed).  fp32 vectors: the oracle's mathematics vs the reference's (tolerance = fp32 summation order,
SDPA vs an explicit softmax).  bf16 vectors: the rounding points vs the reference module run in
bf16 (tolerance = a few bf16 ulps: CPU SDPA / matmul accumulate differently)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
bf16 = torch.bfloat16


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(ROOT, "tests", "golden", "reference_exec.pt"))


@pytest.fixture(scope="module")
def cfg(gold):
    import a3vlm_oracle as O
    import make_reference_golden as G

    a, v = G.tiny_cfg()
    sd = O.init_params(a, v, seed=gold["seed"], output_scale=gold["output_scale"])
    return a, v, sd


def close(got, ref, atol, rtol=0.0):
    torch.testing.assert_close(got.float(), ref.float(), atol=atol, rtol=rtol)


def test_rmsnorm_is_the_reference_class_bit_for_bit(gold):
    import a3vlm_oracle as O

    assert torch.equal(O.rmsnorm(gold["rms_x"], gold["rms_w"], 1e-5), gold["rms_fp32"])
    got = O.rmsnorm(gold["rms_x"].to(bf16), gold["rms_w"].to(bf16), 1e-5)
    assert got.dtype == bf16 and torch.equal(got, gold["rms_bf16"])  # normalise in fp32, round, then scale


def test_decoder_fp32_matches_reference_transformer(gold, cfg):
    """Attention (SDPA branch, KV cache, right-aligned mask), FeedForward, TransformerBlock, Transformer.forward /
    forward_inference of LLM/llama_ens5.py, text only."""
    import a3vlm_oracle as O

    a, _, sd = cfg
    orc = O.OracleTransformer(a, None, sd, act_dtype=torch.float32)
    tok = gold["tokens"]
    close(orc.forward(tok), gold["fwd_fp32"], 2e-4)
    close(orc.forward_inference(tok[:, :6], 0), gold["prefill_fp32"], 2e-4)
    close(orc.forward_inference(tok[:, 6:7], 6), gold["decode1_fp32"], 2e-4)
    close(orc.forward_inference(tok[:, 7:8], 7), gold["decode2_fp32"], 2e-4)
    close(orc.forward_inference(tok[:, 8:11], 8), gold["chunk_fp32"], 2e-4)  # l > 1 on the cache: mask alignment


def test_decoder_bf16_rounding_points_match_reference_run_in_bf16(gold, cfg):
    import a3vlm_oracle as O

    a, _, sd = cfg
    orc = O.OracleTransformer(a, None, sd)  # act_dtype bf16
    tok = gold["tokens"]
    pairs = [(orc.forward(tok), gold["fwd_bf16"]), (orc.forward_inference(tok[:, :6], 0), gold["prefill_bf16"]),
             (orc.forward_inference(tok[:, 6:7], 6), gold["decode1_bf16"]),
             (orc.forward_inference(tok[:, 7:8], 7), gold["decode2_bf16"]),
             (orc.forward_inference(tok[:, 8:11], 8), gold["chunk_bf16"])]
    for got, ref in pairs:
        got, ref = got.float(), ref.float()
        rel = ((got - ref).norm() / ref.norm()).item()
        assert rel <= 1e-2, rel
        assert (got - ref).abs().max().item() <= 4 * 2.0 ** -8 * ref.abs().max().item()
        # and both are the same distance from the unrounded model (fp32 vectors of the same reference code)
    exact = gold["fwd_fp32"]
    e_ref = ((gold["fwd_bf16"].float() - exact).norm() / exact.norm()).item()
    e_got = ((pairs[0][0].float() - exact).norm() / exact.norm()).item()
    assert e_got <= 1.25 * e_ref + 1e-3, (e_got, e_ref)


def test_image_splice_logit_slice_and_decode_offset_match_reference(gold, cfg):
    """llama_ens5.py:469-486,500-520 run on given per-view tokens: BOS | 5 x (start, tokens, end) | text, logits
    for the text positions only, `cache_image_words` added to later start positions."""
    import a3vlm_oracle as O

    a, _, sd = cfg
    v5 = O.VisionArgs(image_size=28, patch=14, width=128, layers=2, heads=2, mlp=256, views=5)
    orc = O.OracleTransformer(a, v5, sd, act_dtype=torch.float32)
    views = gold["img_views"]  # (5, B, 5, D)
    orc.encode_image = lambda image: torch.cat(list(views), dim=0)  # view-major, as :455-457 returns it
    tok = gold["tokens"]
    fake = torch.zeros(2, 3, 4, 4)
    assert orc.image_words == 35
    close(orc.forward(tok, fake), gold["img_fwd_fp32"], 2e-4)
    close(orc.forward_inference(tok[:, :5], 0, fake), gold["img_prefill_fp32"], 2e-4)
    assert orc.cache_image_words == gold["img_cache_image_words"] == 35
    close(orc.forward_inference(tok[:, 5:6], 5), gold["img_decode_fp32"], 2e-4)


def test_rgbd_splice_matches_reference_2images(gold, cfg):
    """LLM/llama_ens5_2images.py:511-547: BOS | RGB views | depth views (own tags) | text; decode offset counts
    both streams; a lone RGB image is ignored at inference (:518)."""
    import a3vlm_oracle as O

    a, _, sd = cfg
    v5 = O.VisionArgs(image_size=28, patch=14, width=128, layers=2, heads=2, mlp=256, views=5)
    p = dict(sd)
    p.update(gold["rgbd_tags"])
    orc = O.OracleTransformer(O.ModelArgs(**{**vars(a), "max_seq_len": 128}), v5, p, act_dtype=torch.float32, rgbd=True)
    rgb, depth = gold["img_views"], gold["rgbd_depth_views"]
    rgb_marker, depth_marker = torch.zeros(2, 3, 4, 4), torch.ones(2, 3, 4, 4)
    orc.encode_image = lambda image: torch.cat(list(depth if bool(image.flatten()[0] == 1) else rgb), dim=0)
    tok = gold["tokens"]
    assert orc.image_words == 70 and orc.visual_image_words == 35
    close(orc.forward_inference(tok[:, :5], 0, rgb_marker, depth_marker), gold["rgbd_prefill_fp32"], 2e-4)
    assert orc.cache_image_words == gold["rgbd_cache_image_words"] == 70
    close(orc.forward_inference(tok[:, 5:6], 5), gold["rgbd_decode_fp32"], 2e-4)
    close(orc.forward_inference(tok[:, :5], 0, rgb_marker), gold["rgbd_lone_rgb_prefill_fp32"], 2e-4)
    assert orc.cache_image_words == gold["rgbd_lone_cache_image_words"] == 0


def test_vit_tower_matches_reference_glue_over_torch_primitives(gold, cfg):
    """clip_encode_image (llama_ens5.py:351-375) executed over conv / LayerNorm / nn.MultiheadAttention."""
    import a3vlm_oracle as O

    a, v, sd = cfg
    orc = O.OracleTransformer(a, v, sd, act_dtype=torch.float32)
    close(orc.clip_encode_image(gold["vit_image"]), gold["vit_tokens_fp32"], 1e-4)


def test_loss_matches_reference_metamodel_forward(gold, cfg):
    import a3vlm_oracle as O

    a, _, sd = cfg
    orc = O.OracleTransformer(a, None, sd, act_dtype=torch.float32)
    loss = O.meta_forward_loss(orc, gold["loss_examples"], gold["loss_labels"])
    close(loss, gold["loss_fp32"], 1e-5, 1e-5)
    ex, lab = O.trim_trailing_pad(gold["loss_examples"], gold["loss_labels"])
    assert ex.shape[1] == 10  # the reference dropped the all-pad tail columns


def test_generate_matches_reference_metamodel_generate(gold, cfg):
    """Greedy ids from the reference's own loop (meta.py:379-485): unequal prompts, stops that end every row
    early, left truncation with a row that keeps generating while the long prompt is force-fed."""
    import a3vlm_oracle as O

    a, _, sd = cfg
    orc = O.OracleTransformer(a, None, sd, act_dtype=torch.float32)
    for c in gold["generate_cases"]:
        prompts = [[1] + [int(x) for x in p.split()] for p in c["prompts"]]
        stops = [[int(x) for x in s.split()] for s in c["stops"]] * 2  # encode_segment + encode_wo_prefix_space
        out, _, _ = O.generate_tokens(orc, prompts, max_gen_len=c["max_gen_len"], eos_id=2, stop_sequences=stops)
        assert out == c["generated"], (c["prompts"], out, c["generated"])


def test_sample_top_p_matches_reference_under_the_same_seed(gold):
    import a3vlm_oracle as O

    torch.manual_seed(1234)
    got = O.sample_top_p(gold["topp_probs"].clone(), gold["topp_p"])
    assert torch.equal(got, gold["topp_tokens_seed1234"])
    # and the explicit-uniform form draws from the same support: every reference draw has positive mass in it
    lo, hi = O.top_p_token_interval(gold["topp_probs"].double(), gold["topp_p"], gold["topp_tokens_seed1234"].reshape(-1))
    assert bool((hi > lo).all())


@pytest.mark.gpu
def test_cuda_decoder_matches_reference_run_in_bf16(gold, cfg):
    """The CUDA path itself (text-only llama_b200 through MetaModel) against the reference module's bf16
    outputs -- no oracle in between.  Tolerances as tests/test_e2e_gpu.py."""
    from a3vlm_b200.model import MetaModel
    from test_e2e_gpu import check

    a, _, sd = cfg
    text = {k: t for k, t in sd.items() if not k.startswith(("clip.", "visual_proj", "start_img", "end_img"))}
    m = MetaModel("llama_b200", dict(dim=a.dim, n_layers=a.n_layers, n_heads=a.n_heads, vocab_size=a.vocab_size,
                                     multiple_of=a.multiple_of, norm_eps=a.norm_eps), None, with_visual=False,
                  max_seq_len=a.max_seq_len, max_batch_size=4)
    m.llma.load_state_dict(text)
    tok = gold["tokens"]
    with torch.no_grad():
        check("cuda vs reference fwd", m.llma(tok), gold["fwd_bf16"])
    check("cuda vs reference prefill", m.llma.forward_inference(tok[:, :6], 0), gold["prefill_bf16"])
    check("cuda vs reference decode 1", m.llma.forward_inference(tok[:, 6:7], 6), gold["decode1_bf16"])
    check("cuda vs reference decode 2", m.llma.forward_inference(tok[:, 7:8], 7), gold["decode2_bf16"])
    check("cuda vs reference 3-token step on the cache", m.llma.forward_inference(tok[:, 8:11], 8), gold["chunk_bf16"])
    # greedy ids of the reference's generate loop (run in fp32): identical, except that a row may part ways at a
    # step where the REFERENCE's own top-2 logits are within bf16 rounding of each other (recorded per call)
    for c in gold["generate_cases"]:
        prompts = [[1] + [int(x) for x in p.split()] for p in c["prompts"]]
        stops = [[int(x) for x in s.split()] for s in c["stops"]] * 2
        out, _, _ = m.generate_tokens(prompts, None, max_gen_len=c["max_gen_len"], eos_id=2, stop_sequences=stops)
        min_prompt = min(len(q) for q in prompts)
        for r, (g_, e_) in enumerate(zip(out, c["generated"])):
            k = next((i for i, (x, y) in enumerate(zip(g_, e_)) if x != y), None)
            if k is None:
                assert len(g_) == len(e_)
                continue
            margin, mag = c["margins"][len(prompts[r]) + k - min_prompt][r].tolist()
            print(f"row {r} diverges at generated token {k}: reference top-2 margin {margin:.4f} at logit {mag:.2f}")
            assert margin <= 4 * 2.0 ** -8 * max(mag, 1.0), (r, k, margin, mag, g_, e_)


def test_lr_schedule_matches_reference_lr_sched(gold):
    from types import SimpleNamespace

    from a3vlm_b200.train import adjust_learning_rate, adjust_learning_rate_epoch

    args = SimpleNamespace(**gold["lr_args"])
    opt = SimpleNamespace(param_groups=[{"lr": 0.0}, {"lr": 0.0, "lr_scale": 0.5}])
    for it, lr, scaled in gold["lr_iter"]:
        assert adjust_learning_rate(opt, it, args) == lr and opt.param_groups[0]["lr"] == lr
        assert opt.param_groups[1]["lr"] == scaled
    for ep, lr in gold["lr_epoch"]:
        assert adjust_learning_rate_epoch(opt, ep, args) == lr


def test_weight_decay_groups_match_reference_add_weight_decay(gold, cfg):
    """misc.add_weight_decay run on the path's parameter names: the engine's decay / no-decay regions hold
    exactly the same sets (incl. the reference's quirk that visual_proj.1 -- a LayerNorm -- IS decayed because
    its name does not end in "norm.weight")."""
    import a3vlm_oracle as O
    from a3vlm_b200.engine import build_layout

    a, v, _ = cfg
    decay, nodecay, frozen, _ = build_layout(a, v)

    def ref_names(layout):
        D = a.dim
        out = []
        for n, _shape in layout:
            if n.endswith(".wqkv"):
                out += [n[:-4] + f"attention.{x}.weight" for x in ("wq", "wk", "wv")]
            elif n.endswith(".w13"):
                out += [n[:-3] + f"feed_forward.{x}.weight" for x in ("w1", "w3")]
            else:
                out.append(n)
        return sorted(out)

    assert ref_names(decay) == gold["wd_decay"]
    assert ref_names(nodecay) == gold["wd_no_decay"]
    assert "visual_proj.1.weight" in gold["wd_decay"] and "visual_proj.1.bias" in gold["wd_no_decay"]
    assert all(O.is_no_decay(n) for n in gold["wd_no_decay"]) and not any(O.is_no_decay(n) for n in gold["wd_decay"])
    assert not any(n.startswith("clip.") for n in gold["wd_decay"] + gold["wd_no_decay"])  # frozen tower: no group


@pytest.mark.gpu
def test_cuda_vit_tower_matches_reference_glue(gold, cfg):
    """The CUDA ViT tower (patchify + patch-embed GEMM + cls / pos / ln_pre + blocks + ln_post) against
    clip_encode_image of the reference executed over torch.nn primitives (fp32): bf16 tolerance."""
    from a3vlm_b200.model import MetaModel
    from test_e2e_gpu import check

    a, v, sd = cfg
    m = MetaModel("llama_b200", dict(dim=a.dim, n_layers=a.n_layers, n_heads=a.n_heads, vocab_size=a.vocab_size,
                                     multiple_of=a.multiple_of, image_size=v.image_size, vit_width=v.width,
                                     vit_layers=v.layers, vit_heads=v.heads, vit_mlp=v.mlp), None, with_visual=True,
                  max_seq_len=a.max_seq_len, max_batch_size=4)
    m.llma.load_state_dict(sd)
    eng = m.llma.engine
    vw = eng._vit_workspace(2, False)
    eng.vit_forward(gold["vit_image"].cuda(), vw)
    check("cuda ViT vs reference glue", vw["feats"].view(2, v.tokens, v.width), gold["vit_tokens_fp32"], rel=1e-2, ulps=6)


def test_tokenizer_wrapper_matches_reference_tokenizer_class(gold, tmp_path):
    """a3vlm_b200.model.Tokenizer against the reference's Tokenizer class run on the same tiny SentencePiece
    model (tests/golden/toy_spm.model): ids with / without bos / eos, mid-text segments, glued segments (the
    stop-sequence encodings generate() uses, meta.py:440-441), decode, vocabulary size, save."""
    from a3vlm_b200.model.tokenizer import Tokenizer

    c = gold["tok_cases"]
    t = Tokenizer(os.path.join(ROOT, "tests", "golden", "toy_spm.model"))
    assert (t.n_words, t.bos_id, t.eos_id, t.need_space_before_segment) == (
        c["n_words"], c["bos_id"], c["eos_id"], c["need_space_before_segment"])
    assert t.tokenizer_type == "spm"
    for text, bos, eos, ids in c["encode"]:
        assert t.encode(text, bos, eos) == ids, (text, bos, eos)
    for text, ids in c["encode_segment"]:
        assert t.encode_segment(text) == ids, text
    for text, ids in c["encode_wo_prefix_space"]:
        assert t.encode_wo_prefix_space(text) == ids, text
    for ids, text in c["decode"]:
        assert t.decode(ids) == text
    with pytest.raises(TypeError):
        t.encode(["not", "a", "string"], True, False)
    t.save(str(tmp_path))
    again = Tokenizer(str(tmp_path / "tokenizer.model"))
    assert again.encode("the robot opens the drawer", True, True) == t.encode("the robot opens the drawer", True, True)


def test_image_transform_matches_reference_get_transform(gold):
    """a3vlm_b200.data.get_transform against the reference's data/transform.py run on the same synthetic
    images: bit-identical tensors for all three pipelines (the random crop under the same torch seed)."""
    import make_reference_golden as G
    from a3vlm_b200.data import get_transform

    for w, h, kind, size, ref in gold["transform_cases"]:
        torch.manual_seed(7)
        got = get_transform(kind, size)(G.synthetic_image(w, h))
        assert got.shape == (3, size, size) and torch.equal(got, ref), (w, h, kind)
    with pytest.raises(ValueError):
        get_transform("no_such_pipeline")


class _OracleLLM(torch.nn.Module):
    """CPU stand-in for the inner model behind our MetaModel's host-side methods: the fp32 oracle."""

    def __init__(self, orc):
        super().__init__()
        self.orc, self.args, self.image_words = orc, orc.args, 0
        self.w = torch.nn.Parameter(torch.zeros(1))

    def forward(self, tokens, image=None):
        return self.orc.forward(tokens, image)

    def forward_inference(self, tokens, start_pos, image=None):
        return self.orc.forward_inference(tokens, start_pos, image)


def test_scoring_api_matches_reference_compute_logits_and_evaluate_examples(gold, cfg):
    """OUR MetaModel.compute_logits / evaluate_examples (host logic; the oracle supplies the logits) against the
    reference's methods executed on its own decoder: same keys, context handling, "ppl" = mean token loss,
    `max_equal`, text inputs tokenised with (bos, eos) for examples and (bos, False) for contexts."""
    import a3vlm_oracle as O
    import make_reference_golden as G
    from a3vlm_b200.model.meta import MetaModel

    a, _, sd = cfg
    m = MetaModel.__new__(MetaModel)
    torch.nn.Module.__init__(m)
    m.llma = _OracleLLM(O.OracleTransformer(a, None, sd, act_dtype=torch.float32))
    m.tokenizer = G.ToyTokenizer()
    c = gold["score_cases"]
    got = m.compute_logits(c["examples"])
    assert [tuple(t.shape) for t in got] == [tuple(t.shape) for t in c["compute_logits"]]
    for g_, r_ in zip(got, c["compute_logits"]):
        assert g_.dtype == torch.float32
        close(g_, r_, 2e-4)
    runs = {"all": m.evaluate_examples(c["examples"]), "ctx": m.evaluate_examples(c["examples"], c["contexts"]),
            "txt": m.evaluate_examples(["17 33 49", "100 101"], ["17", "100"], None, True, True)}
    for name, res in runs.items():
        ref = c[name]
        assert set(res) == set(ref) == {"log_likelihood", "ppl", "max_equal", "non_context_logits"}
        assert res["max_equal"] == ref["max_equal"]
        for k in ("log_likelihood", "ppl"):
            assert len(res[k]) == len(ref[k])
            for x, y in zip(res[k], ref[k]):
                assert abs(x - y) <= 1e-4 * max(1.0, abs(y)), (name, k, x, y)
        for x, y in zip(res["non_context_logits"], ref["non_context_logits"]):
            close(x, y, 2e-4)
    with pytest.raises(AssertionError):
        m.evaluate_examples([[1, 2, 3]], [[1, 5]])  # a context must be a prefix of its example
    with pytest.raises(ValueError):
        m.compute_logits("a single string")


def test_stream_generate_matches_reference_stream_generate(gold, cfg):
    """OUR MetaModel.stream_generate (host loop; the fp32 oracle supplies the logits and torch.argmax stands in
    for the argmax kernel) yields the same chunk sequence as the reference's generator: plain greedy run, a stop
    symbol that truncates the text and ends the stream, a run bounded by max_gen_len."""
    import a3vlm_oracle as O
    import make_reference_golden as G
    from a3vlm_b200.model.meta import MetaModel

    a, _, sd = cfg
    m = MetaModel.__new__(MetaModel)
    torch.nn.Module.__init__(m)
    m.llma = _OracleLLM(O.OracleTransformer(a, None, sd, act_dtype=torch.float32))
    m.tokenizer = G.ToyTokenizer()
    m._argmax = lambda logits: torch.argmax(logits, dim=-1)
    for c in gold["stream_cases"]:
        got = list(m.stream_generate(c["prompt"], None, c["max_gen_len"], 0.0, 0.95, c["stops"]))
        assert got == c["chunks"], (c["prompt"], got, c["chunks"])


def test_metamodel_forward_host_path_matches_reference_loss(gold, cfg):
    """OUR MetaModel.forward on its un-fused path (an inner model without `loss_forward`: trim trailing pad columns,
    shift, CrossEntropyLoss(ignore_index=0)) against the reference's MetaModel.forward run on its own decoder."""
    import a3vlm_oracle as O
    from a3vlm_b200.model.meta import MetaModel

    a, _, sd = cfg
    m = MetaModel.__new__(MetaModel)
    torch.nn.Module.__init__(m)
    m.llma = _OracleLLM(O.OracleTransformer(a, None, sd, act_dtype=torch.float32))
    m.criterion = torch.nn.CrossEntropyLoss(ignore_index=0)
    loss, extra = m(gold["loss_examples"], gold["loss_labels"])
    assert extra == {}
    close(loss, gold["loss_fp32"], 1e-5, 1e-5)


def test_generate_with_images_and_with_sampling_match_reference(gold, cfg):
    """(1) With images the reference shrinks the budget by image_words (meta.py:423-425) and feeds the image to
    the prefill only; (2) with temperature > 0 the reference's loop draws through sample_top_p -- under the same
    torch seed the oracle's loop consumes the generator identically, so the ids are equal."""
    import a3vlm_oracle as O

    a, _, sd = cfg
    v5 = O.VisionArgs(image_size=28, patch=14, width=128, layers=2, heads=2, mlp=256, views=5)
    orc = O.OracleTransformer(a, v5, sd, act_dtype=torch.float32)
    views = gold["img_views"]
    orc.encode_image = lambda image: torch.cat(list(views), dim=0)
    c = gold["generate_image_case"]
    prompts = [[1] + [int(x) for x in p.split()] for p in c["prompts"]]
    out, toks, _ = O.generate_tokens(orc, prompts, images=torch.zeros(2, 3, 4, 4), max_gen_len=c["max_gen_len"], eos_id=2)
    assert out == c["generated"]
    assert toks.shape[1] == min(a.max_seq_len - 35, c["max_gen_len"] + 8)  # the budget the reference applied

    txt = O.OracleTransformer(a, None, sd, act_dtype=torch.float32)
    c = gold["generate_sampled_case"]
    prompts = [[1] + [int(x) for x in p.split()] for p in c["prompts"]]
    torch.manual_seed(c["seed"])
    out, _, _ = O.generate_tokens(txt, prompts, max_gen_len=c["max_gen_len"], temperature=c["temperature"], top_p=c["top_p"], eos_id=2)
    assert out == c["generated"]


def test_reference_metamodel_reaches_the_real_plugin():
    """The literal one-line stub of INTEGRATION.md (`accessory/model/LLM/llama_b200.py` re-exporting our
    ModelArgs / Transformer), resolved by the reference's UNMODIFIED MetaModel.__init__ through its importlib seam
    (meta.py:30-32): the reference builds our ModelArgs from its JSON config and calls our real constructor,
    which refuses to build without a CUDA device -- there is no CPU fallback behind the seam.  Runs where
    /root/reference exists (this container); the GPU half is tests/test_seam_gpu.py."""
    import importlib
    import json
    import tempfile
    import types

    if not os.path.isdir("/root/reference/model"):
        pytest.skip("reference tree not present")
    if torch.cuda.is_available():
        pytest.skip("CPU-side half of the seam proof")
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_seam_trace as MS
    from a3vlm_b200 import _lib
    from a3vlm_b200.model.LLM import llama_b200 as ours

    M = MS.install_seam(ours.Transformer)
    stub = types.ModuleType("accessory.model.LLM.llama_b200")
    exec("from a3vlm_b200.model.LLM.llama_b200 import ModelArgs, Transformer  # noqa: F401", stub.__dict__)
    sys.modules["accessory.model.LLM.llama_b200"] = stub
    assert importlib.import_module("accessory.model.LLM.llama_b200").Transformer is ours.Transformer
    seen = {}
    real_init = ours.Transformer.__init__

    def spy(self, args, with_visual=False):
        seen["args"], seen["with_visual"] = args, with_visual
        return real_init(self, args, with_visual=with_visual)

    ours.Transformer.__init__ = spy
    try:
        with tempfile.TemporaryDirectory() as td:
            cfg = os.path.join(td, "params.json")
            json.dump(MS.seam_cfg(), open(cfg, "w"))
            with pytest.raises(_lib.A3Error, match="no CPU fallback"):
                M.MetaModel("llama_b200", cfg, os.path.join(td, "tokenizer.model"), with_visual=True, max_seq_len=64)
    finally:
        ours.Transformer.__init__ = real_init
    assert isinstance(seen["args"], ours.ModelArgs) and seen["with_visual"] is True
    assert (seen["args"].dim, seen["args"].max_seq_len, seen["args"].max_batch_size, seen["args"].vocab_size) == (256, 64, 32, 512)
    # and the recorded trace the GPU half replays was made from the same constructor arguments
    tr = torch.load(os.path.join(ROOT, "tests", "golden", "reference_seam_trace.pt"), weights_only=False)
    import dataclasses
    assert tr["ctor"]["model_args"] == dataclasses.asdict(seen["args"])
