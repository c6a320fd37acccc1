"""Checkpoint loader / saver (a3vlm_b200/checkpoint.py).

CPU part: the format logic restated from the reference's util/tensor_parallel.py -- file-name patterns,
format / shard-count inference and its error cases, split dimensions, streaming the shards of 1- and
2-way model-parallel `consolidated` and `meta_ori` checkpoints into FUSED targets (wq|wk|wv and w1|w3 as
row ranges of one matrix, like the engine's flat buffers), `consolidated_diff` addition, unequal
replicas, missing / unexpected keys.  The expected result is what the reference computes:
`torch.cat(shards, dim)` per tensor (tensor_parallel.py:113-114).

GPU part: save_checkpoint -> from_pretrained / resume round trips on the tiny model (bit-exact weights
and logits; the optimiser continues exactly)."""
import json
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
bf16 = torch.bfloat16


def _tiny_sd(rgbd=False, seed=31):
    import a3vlm_oracle as O

    a = O.ModelArgs(dim=256, n_layers=2, n_heads=2, vocab_size=512, multiple_of=64, max_seq_len=64, max_batch_size=4)
    v = O.VisionArgs(image_size=28, patch=14, width=128, layers=2, heads=2, mlp=256)
    return a, v, O.init_params(a, v, seed=seed, output_scale=8.0, rgbd=rgbd)


def _split(sd, mp, prefix="llma."):
    """What a model-parallel run of the reference would have saved: rank r's chunk of every
    tensor-parallel weight, replicas of the rest."""
    from a3vlm_b200.checkpoint import weight_parallel_dim

    shards = [{} for _ in range(mp)]
    for k, t in sd.items():
        d = weight_parallel_dim(k)
        pieces = torch.chunk(t, mp, dim=d) if d is not None else [t] * mp
        for r in range(mp):
            shards[r][prefix + k] = pieces[r].clone()
    return shards


def _fused_targets(sd):
    """Targets laid out like the engine's flat buffers: q|k|v and w1|w3 are views of one matrix."""
    targets, fused = {}, {}
    for k, t in sd.items():
        if k.endswith("attention.wq.weight"):
            p = k[:-len("attention.wq.weight")]
            D = t.shape[0]
            m = torch.zeros(3 * D, t.shape[1], dtype=bf16)
            fused[p + "wqkv"] = m
            for j, n in enumerate(("wq", "wk", "wv")):
                targets[f"{p}attention.{n}.weight"] = m[j * D:(j + 1) * D]
        elif k.endswith("feed_forward.w1.weight"):
            p = k[:-len("feed_forward.w1.weight")]
            F = t.shape[0]
            m = torch.zeros(2 * F, t.shape[1], dtype=bf16)
            fused[p + "w13"] = m
            targets[p + "feed_forward.w1.weight"], targets[p + "feed_forward.w3.weight"] = m[:F], m[F:]
    for k, t in sd.items():
        if k not in targets:
            targets[k] = torch.zeros_like(t)
    return targets, fused


def test_split_dimensions_follow_the_reference_layer_classes():
    from a3vlm_b200 import checkpoint as C

    dims = C.get_weight_parallel_dim(["llma.layers.3.attention.wq.weight", "llma.layers.3.attention.wo.weight",
                                      "llma.layers.0.feed_forward.w1.weight", "llma.layers.0.feed_forward.w2.weight",
                                      "llma.layers.0.feed_forward.w3.weight", "llma.tok_embeddings.weight",
                                      "llma.output.weight", "llma.norm.weight", "llma.layers.0.attention_norm.weight",
                                      "llma.visual_proj.0.weight", "llma.clip.visual.conv1.weight", "llma.start_img"])
    assert dims == {"llma.layers.3.attention.wq.weight": 0, "llma.layers.3.attention.wo.weight": 1,
                    "llma.layers.0.feed_forward.w1.weight": 0, "llma.layers.0.feed_forward.w2.weight": 1,
                    "llma.layers.0.feed_forward.w3.weight": 0, "llma.tok_embeddings.weight": 1, "llma.output.weight": 0}
    # the same table the reference keeps per layer class (tensor_parallel.py:35-39)
    assert [(c.__name__, d) for c, d in C.MODEL_PARALLEL_MODULES] == [
        ("ColumnParallelLinear", {"weight": 0, "bias": 0}), ("RowParallelLinear", {"weight": 1, "bias": -1}),
        ("ParallelEmbedding", {"weight": 1})]
    assert C.holder_class_for("layers.0.attention.wv.weight") is C.ColumnParallelLinear
    assert C.holder_class_for("layers.0.ffn_norm.weight") is None


def test_format_inference_and_errors(tmp_path):
    from a3vlm_b200 import checkpoint as C

    d = tmp_path / "a"
    d.mkdir()
    for r in range(2):
        torch.save({}, d / f"consolidated.{r:02d}-of-02.model.pth")
    assert C.infer_checkpoint_format_and_mp_size(str(d)) == ("consolidated", 2)
    m = tmp_path / "m"
    m.mkdir()
    torch.save({}, m / "consolidated.00.pth")
    assert C.infer_checkpoint_format_and_mp_size(str(m)) == ("meta_ori", 1)
    torch.save({}, m / "consolidated.00-of-01.model-diff.pth")
    with pytest.raises(NotImplementedError, match="Multiple matched"):
        C.infer_checkpoint_format_and_mp_size(str(m))
    e = tmp_path / "e"
    e.mkdir()
    (e / "readme.txt").write_text("x")
    with pytest.raises(NotImplementedError, match="do not match"):
        C.infer_checkpoint_format_and_mp_size(str(e))
    with pytest.raises(NotImplementedError, match="valid folder"):
        C.infer_checkpoint_format_and_mp_size(str(e / "readme.txt"))
    h = tmp_path / "h"
    h.mkdir()
    torch.save({}, h / "consolidated.00-of-02.model.pth")
    torch.save({}, h / "consolidated.01-of-03.model.pth")  # 2 files, neither 00-of-02 + 01-of-02
    with pytest.raises(NotImplementedError, match="expected file"):
        C.infer_checkpoint_format_and_mp_size(str(h))
    assert C.get_tensor_parallel_shards_file_name("consolidated_diff", 2)[1] == "consolidated.01-of-02.model-diff.pth"


@pytest.mark.parametrize("mp", [1, 2, 4])
@pytest.mark.parametrize("fmt", ["consolidated", "meta_ori"])
def test_streaming_merge_equals_reference_concatenation(tmp_path, mp, fmt):
    from a3vlm_b200 import checkpoint as C

    _, _, sd = _tiny_sd()
    if fmt == "meta_ori":  # Meta's files: no prefix, no wrapper, LLM tensors only
        sd = {k: v for k, v in sd.items() if not k.startswith(("clip.", "visual_proj", "start_img", "end_img"))}
    shards = _split(sd, mp, prefix="" if fmt == "meta_ori" else "llma.")
    for r, sh in enumerate(shards):
        fn = C.get_tensor_parallel_shards_file_name(fmt, mp)[r]
        torch.save(sh if fmt == "meta_ori" or r % 2 else {"model": sh}, tmp_path / fn)  # wrapper is optional
    targets, fused = _fused_targets(sd)
    res = C.load_into({"llma." + k: t for k, t in targets.items()}, str(tmp_path))
    assert res == {"missing_keys": [], "unexpected_keys": []}
    for k, t in sd.items():  # == torch.cat(param_shards, dim) of the reference, == the original tensor
        assert torch.equal(targets[k], t), k
    for k, m in fused.items():  # and the fused kernel-side matrices hold q|k|v / w1|w3
        p = k.rsplit(".", 1)[0] + "."
        if k.endswith("wqkv"):
            want = torch.cat([sd[p + f"attention.{n}.weight"] for n in ("wq", "wk", "wv")])
        else:
            want = torch.cat([sd[p + "feed_forward.w1.weight"], sd[p + "feed_forward.w3.weight"]])
        assert torch.equal(m, want), k


def test_diff_checkpoint_missing_unexpected_and_unequal_replicas(tmp_path, capsys):
    from a3vlm_b200 import checkpoint as C

    _, _, sd = _tiny_sd()
    base, diff = tmp_path / "base", tmp_path / "diff"
    base.mkdir()
    diff.mkdir()
    shards = _split(sd, 2)
    del shards[0]["llma.norm.weight"], shards[1]["llma.norm.weight"]          # a missing key
    shards[0]["llma.not_a_parameter"] = torch.zeros(3)                         # an unexpected key (discarded)
    shards[1]["llma.layers.0.ffn_norm.weight"] = shards[1]["llma.layers.0.ffn_norm.weight"] + 1  # unequal replica
    for r, sh in enumerate(shards):
        torch.save({"model": sh}, base / f"consolidated.{r:02d}-of-02.model.pth")
    delta = {"llma.layers.1.attention.wo.weight": torch.full_like(sd["layers.1.attention.wo.weight"], 0.5),
             "llma.norm.weight": torch.full_like(sd["norm.weight"], 3.0)}  # not loaded before: set, not added
    torch.save({"model": delta}, diff / "consolidated.00-of-01.model-diff.pth")
    targets, _ = _fused_targets(sd)
    named = {"llma." + k: t for k, t in targets.items()}
    with pytest.raises(AssertionError):
        C.load_into(dict(named), [str(diff)])  # a diff cannot come first (tensor_parallel.py:455-457)
    res = C.load_into(named, [str(base)])
    out = capsys.readouterr().out
    assert res["missing_keys"] == ["llma.norm.weight"] and res["unexpected_keys"] == []
    assert "discard unexpected parameter: llma.not_a_parameter" in out
    assert "unequal replicas" in out and "layers.0.ffn_norm.weight" in out
    res = C.load_into(named, [str(base), str(diff)])
    assert res == {"missing_keys": [], "unexpected_keys": []}
    want = (sd["layers.1.attention.wo.weight"] + torch.full_like(sd["layers.1.attention.wo.weight"], 0.5))
    assert torch.equal(targets["layers.1.attention.wo.weight"], want)          # added in the parameter dtype
    assert torch.equal(targets["norm.weight"], torch.full_like(sd["norm.weight"], 3.0))
    assert torch.equal(targets["layers.0.ffn_norm.weight"], sd["layers.0.ffn_norm.weight"])  # shard 0 wins


def test_shard_shape_errors_are_loud(tmp_path):
    from a3vlm_b200 import checkpoint as C

    with pytest.raises(ValueError, match="does not split"):
        C.tensor_load_shard(torch.zeros(5, 4), 0, 2, 0, torch.zeros(2, 4))
    with pytest.raises(ValueError, match="shape mismatch"):
        C.tensor_load_shard(torch.zeros(4, 4), 0, 2, 1, torch.zeros(3, 4))
    t = torch.zeros(4, 6)
    C.tensor_load_shard(t, 1, 3, 2, torch.ones(4, 2))
    assert t[:, 4:].eq(1).all() and t[:, :4].eq(0).all()
    C.tensor_load_shard(t, -1, 3, 1, torch.full((4, 6), 7.0))  # replicated: only shard 0 writes
    assert t[:, :4].eq(0).all()


def test_latest_checkpoint_ordering(tmp_path):
    from a3vlm_b200 import checkpoint as C

    for n in ("epoch0", "epoch1-iter500", "epoch1-iter1500", "epoch0-iter9999", "notes"):
        (tmp_path / n).mkdir()
    assert C.get_latest_checkpoint_path(str(tmp_path)) == "epoch1-iter1500"
    (tmp_path / "epoch1").mkdir()  # a finished epoch outranks its iterations (misc.py:455-457)
    assert C.get_latest_checkpoint_path(str(tmp_path)) == "epoch1"
    assert C.checkpoint_dir_name(3, 20) == "epoch3-iter20" and C.checkpoint_dir_name(3) == "epoch3"


# ------------------------------------------------------------------------------------------ GPU round trips
def _build(a, v, sd, llama_type="llama_b200"):
    from a3vlm_b200.model import MetaModel

    cfg = dict(dim=a.dim, n_layers=a.n_layers, n_heads=a.n_heads, vocab_size=a.vocab_size, multiple_of=a.multiple_of,
               image_size=v.image_size, vit_width=v.width, vit_layers=v.layers, vit_heads=v.heads, vit_mlp=v.mlp)
    m = MetaModel(llama_type, cfg, None, with_visual=True, max_seq_len=a.max_seq_len, max_batch_size=4)
    m.llma.load_state_dict(sd)
    return m


@pytest.mark.gpu
@pytest.mark.parametrize("llama_type", ["llama_b200", "llama_b200_2images"])
def test_save_then_from_pretrained_is_bit_exact(tmp_path, llama_type):
    from types import SimpleNamespace

    from a3vlm_b200 import checkpoint as C
    from a3vlm_b200.model import MetaModel

    a, v, sd = _tiny_sd(rgbd=llama_type.endswith("2images"))
    m = _build(a, v, sd, llama_type)
    d = C.save_checkpoint(str(tmp_path), SimpleNamespace(precision="bf16"), m, epoch=0, iteration=7)
    assert sorted(os.listdir(d)) == ["config.json", "consolidated.00-of-01.model.pth", "consolidated.00-of-01.other.pth",
                                     "meta.json", "rank-specific-00000-of-00001.pth"]
    assert json.load(open(os.path.join(d, "meta.json"))) == {"llama_type": llama_type}
    saved = torch.load(os.path.join(d, "consolidated.00-of-01.model.pth"))["model"]
    assert set(saved) == {"llma." + k for k in sd} and all(torch.equal(saved["llma." + k], sd[k]) for k in sd)
    # re-save as a 2-way model-parallel checkpoint, as the reference would have written it
    mp_dir = tmp_path / "mp2"
    mp_dir.mkdir()
    for r, sh in enumerate(_split(sd, 2)):
        torch.save({"model": sh}, mp_dir / f"consolidated.{r:02d}-of-02.model.pth")
    for f in ("config.json", "meta.json"):
        (mp_dir / f).write_text(open(os.path.join(d, f)).read())
    g = torch.Generator().manual_seed(4)
    tok = torch.randint(3, 512, (2, 7), generator=g)
    tok[:, 0] = 1
    img = torch.randn(2, 3, 28, 28, generator=g)
    ref = m.llma.forward_inference(tok, 0, img, *([img.flip(-1)] if llama_type.endswith("2images") else [])).clone()
    for path in (d, str(mp_dir)):
        m2 = MetaModel.from_pretrained(path, with_visual=True, max_seq_len=a.max_seq_len, max_batch_size=4)
        assert m2.llama_type == llama_type and not m2.training
        got = m2.llma.state_dict()
        assert set(got) == set(sd) and all(torch.equal(got[k].cpu(), sd[k]) for k in sd)
        out = m2.llma.forward_inference(tok, 0, img, *([img.flip(-1)] if llama_type.endswith("2images") else []))
        assert torch.equal(out, ref)
        del m2


@pytest.mark.gpu
def test_resume_continues_the_optimiser_exactly(tmp_path):
    from types import SimpleNamespace

    from a3vlm_b200 import checkpoint as C
    from a3vlm_b200.train import AdamW, NativeScalerWithGradNormCount

    a, v, sd = _tiny_sd()
    g = torch.Generator().manual_seed(6)
    tok = torch.randint(3, 512, (2, 9), generator=g)
    tok[:, 0] = 1
    lab = tok.clone()
    lab[:, :4] = 0
    img = torch.randn(2, 3, 28, 28, generator=g)

    def step(m, opt):
        m.zero_grad(set_to_none=True)
        loss, _ = m(tok.cuda(), lab.cuda(), images=img.cuda())
        NativeScalerWithGradNormCount()(loss, opt, m, clip_grad=0.5, update_grad=True)
        return loss.item()

    m = _build(a, v, sd)
    opt = AdamW(m, lr=1e-3, weight_decay=0.1)
    step(m, opt)
    d = C.save_checkpoint(str(tmp_path), SimpleNamespace(precision="bf16"), m, opt, NativeScalerWithGradNormCount(),
                          {"pos": 5}, epoch=1, iteration=3)
    assert "consolidated.00-of-01.optimizer.pth" in os.listdir(d)
    l2 = step(m, opt)
    want = {k: t.clone() for k, t in m.llma.state_dict().items()}

    m2 = _build(a, v, sd)  # fresh weights; everything else comes from the checkpoint directory's parent
    opt2 = AdamW(m2, lr=1e-3, weight_decay=0.1)
    state = {}
    ds = SimpleNamespace(load_state_dict=lambda s: state.update(s))
    assert C.resume(str(tmp_path), m2, opt2, NativeScalerWithGradNormCount(), ds) == [1, 4]
    assert state == {"pos": 5} and m2.llma.engine.opt_step == 1
    l2b = step(m2, opt2)
    assert l2b == l2
    got = m2.llma.state_dict()
    for k in want:  # bit-exact up to the bf16 atomics of the embedding scatter-add (order-dependent last bit)
        d = (got[k].float() - want[k].float()).abs()
        assert (d <= 2.0 ** -7 * want[k].float().abs()).all(), k
        assert (d == 0).float().mean() > 0.999, k


def test_from_pretrained_refuses_what_it_does_not_implement(tmp_path):
    """Loud errors instead of silent differences (no CUDA needed: all raised before the model is built)."""
    from a3vlm_b200.model import MetaModel

    with pytest.raises(ValueError, match="pretrained_path"):
        MetaModel.from_pretrained([])
    with pytest.raises(NotImplementedError, match="hf://"):
        MetaModel.from_pretrained("hf://Alpha-VLLM/whatever")
    with pytest.raises(NotImplementedError, match="quant"):
        MetaModel.from_pretrained(str(tmp_path), llama_type="llama_b200", quant=True)
    with pytest.raises(NotImplementedError, match="bf16"):
        MetaModel.from_pretrained(str(tmp_path), llama_type="llama_b200", dtype=torch.float16)
    with pytest.raises(ValueError, match="llama_type"):
        MetaModel.from_pretrained(str(tmp_path))  # no meta.json to read it from
    (tmp_path / "meta.json").write_text(json.dumps({"llama_type": "llama_ens5"}))
    with pytest.raises(FileNotFoundError, match="tokenizer"):
        MetaModel.from_pretrained(str(tmp_path))  # reference plugin name is mapped, but nothing to tokenise with
