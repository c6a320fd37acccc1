"""CPU tests of the host-side logic that does not need the CUDA library: flat-buffer layout and
bucket plan, the data-parallel gradient exchange (world_size 2 over gloo), rope table, the FLOP model
of bench.py against SURVEY 8d, and the host-side set-up of MetaModel.generate."""
import os
import sys
from types import SimpleNamespace

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def _args(**kw):
    d = dict(dim=256, n_layers=3, n_heads=2, n_kv_heads=None, vocab_size=512, multiple_of=64, ffn_dim_multiplier=None)
    d.update(kw)
    return SimpleNamespace(**d)


def _vis():
    return SimpleNamespace(image_size=28, patch=14, width=128, layers=2, heads=2, mlp=256, tokens=5, grid=2)


def test_layout_matches_reference_names_and_ffn_rule():
    import a3vlm_oracle as O
    from a3vlm_b200.engine import build_layout, ffn_hidden_dim, flat_offsets

    assert ffn_hidden_dim(4096, 256, None) == 11008 and ffn_hidden_dim(5120, 256, None) == 13824
    a, v = _args(), _vis()
    decay, nodecay, frozen, kpad = build_layout(a, v)
    assert kpad == 768  # 3 channels x (14 x 14 window zero-padded to 16 x 16)
    ref = O.init_params(O.ModelArgs(dim=256, n_layers=3, n_heads=2, vocab_size=512, multiple_of=64),
                        O.VisionArgs(image_size=28, patch=14, width=128, layers=2, heads=2, mlp=256))
    ours = {}
    for n, s in decay + nodecay + frozen:
        if n.endswith(".wqkv"):
            for k in ("wq", "wk", "wv"):
                ours[n[:-4] + f"attention.{k}.weight"] = (s[0] // 3, s[1])
        elif n.endswith(".w13"):
            for k in ("w1", "w3"):
                ours[n[:-3] + f"feed_forward.{k}.weight"] = (s[0] // 2, s[1])
        elif not n.endswith("conv1_padded"):
            ours[n] = tuple(s)
    assert set(ours) == set(ref)
    for k, t in ref.items():
        assert tuple(t.shape) == ours[k], k
    # decay / no-decay split of M/util/misc.py:586-599
    assert all(O.is_no_decay(n) for n, _ in nodecay)
    assert not any(O.is_no_decay(n) for n, _ in decay)
    offs, total = flat_offsets(decay + nodecay)
    assert all(o % 64 == 0 for o, _, _ in offs.values()) and total % 64 == 0


def test_bucket_plan_tiles_the_flat_buffer_exactly_once():
    from a3vlm_b200.engine import bucket_plan, build_layout, flat_offsets

    for vis in (None, _vis()):
        a = _args()
        decay, nodecay, _, _ = build_layout(a, vis)
        offs, total = flat_offsets(decay + nodecay)
        plan = bucket_plan(offs, total, a.n_layers)
        assert len(plan) == a.n_layers + 3
        cover = torch.zeros(total, dtype=torch.int32)
        for lo, hi in plan:
            assert 0 <= lo < hi <= total
            cover[lo:hi] += 1
        assert bool((cover == 1).all())


def _dp_worker(rank, world, port, total, plan, out, decay_numel=None, chunks=None):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from a3vlm_b200.dp import BucketAllReduce, BucketReduceScatter
    from a3vlm_b200.engine import owned_ranges

    g = torch.Generator().manual_seed(rank)
    # local gradients already carry the 1/world factor (folded into the loss gradient)
    local = torch.randn(total, generator=g) / world
    expect = sum(torch.randn(total, generator=torch.Generator().manual_seed(k)) for k in range(world)) / world
    flat = local.clone()
    r = BucketAllReduce(flat, None, None)
    for lo, hi in plan:
        r.on_bucket(lo, hi)
    r.finish()
    ok = torch.allclose(flat, expect, atol=1e-6) and sorted(r.covered) == sorted(plan)
    if decay_numel is not None:
        # sharded optimiser: after the exchange every range this rank OWNS (its slice of each chunk + the
        # replicated tail) holds the mean gradient; the ranks' slices tile every chunk
        flat = local.clone()
        rs = BucketReduceScatter(flat, decay_numel, world, rank, None, None)
        for lo, hi in plan:
            rs.on_bucket(lo, hi)
        rs.finish()
        mine = owned_ranges(chunks, decay_numel, total, world, rank)
        for (lo, hi, decayed), (clo, chi) in zip(mine, chunks + [(decay_numel, total)]):
            ok = ok and torch.allclose(flat[lo:hi], expect[lo:hi], atol=1e-6)
            ok = ok and ((lo, hi) == rs.owned(clo, chi) if decayed else (lo, hi) == (decay_numel, total))
        ok = ok and sorted(rs.covered) == sorted(plan)
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_gradient_exchange_world2_gloo():
    """N > 1 path on CPU: two ranks, the bucketed SUM all-reduce over the engine's bucket plan yields
    the mean gradient on both ranks; the sharded variant (reduce-scatter per chunk, all-reduce of the
    tail) yields it on every range a rank owns."""
    from a3vlm_b200.engine import bucket_plan, build_layout, flat_offsets, optimizer_chunks

    a = _args()
    decay, nodecay, _, _ = build_layout(a, _vis())
    offs, total = flat_offsets(decay + nodecay, segment_after=decay[-1][0])
    decay_numel = offs[nodecay[0][0]][0]
    plan = bucket_plan(offs, total, a.n_layers, decay_numel)
    chunks = optimizer_chunks(offs, decay_numel, a.n_layers)
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_dp_worker, args=(2, port, total, plan, out, decay_numel, chunks), nprocs=2, join=True)
    assert out[0] and out[1]


def test_optimizer_chunks_and_owned_slices_tile_the_decay_region():
    """The pipelined / sharded optimiser's plan: chunks in forward-consumption order tile [0, decay_numel),
    are the decay buckets of the backward in reverse, and for 1 / 2 / 4 / 8 ranks the ranks' 16-byte aligned
    slices tile every chunk exactly once."""
    from a3vlm_b200.engine import bucket_plan, build_layout, flat_offsets, optimizer_chunks, owned_ranges

    for vis in (None, _vis()):
        a = _args()
        decay, nodecay, _, _ = build_layout(a, vis)
        offs, total = flat_offsets(decay + nodecay, segment_after=decay[-1][0])
        decay_numel = offs[nodecay[0][0]][0]
        chunks = optimizer_chunks(offs, decay_numel, a.n_layers)
        assert chunks[0][0] == 0 and chunks[-1][1] == decay_numel
        assert all(x[1] == y[0] for x, y in zip(chunks[:-1], chunks[1:]))
        assert chunks[1][0] == offs["layers.0.wqkv"][0] and chunks[-1][0] == offs["output.weight"][0]
        plan = bucket_plan(offs, total, a.n_layers, decay_numel)
        assert plan[:-1] == chunks[::-1] and plan[-1] == (decay_numel, total)
        for world in (1, 2, 4, 8):
            cover = torch.zeros(total, dtype=torch.int32)
            for rank in range(world):
                mine = owned_ranges(chunks, decay_numel, total, world, rank)
                assert mine[-1] == (decay_numel, total, False) and len(mine) == len(chunks) + 1
                for lo, hi, decayed in mine[:-1]:
                    assert decayed and lo % 8 == 0 and (hi - lo) % 8 == 0
                    cover[lo:hi] += 1
            assert bool((cover[:decay_numel] == 1).all())


def test_rope_table_is_the_reference_expression():
    import a3vlm_oracle as O
    from a3vlm_b200.ops import rope_table

    t = rope_table(128, 64, 10000.0)
    ref = torch.view_as_real(O.precompute_freqs_cis(128, 64, 10000.0))
    assert torch.equal(t, ref) and t.shape == (64, 64, 2)


def test_bench_flop_model_matches_survey():
    import bench

    fl = bench.flops_per_sample(bench.MODELS["7b"], 336, 256)
    assert fl["L"] == 835 and fl["N"] == 577 and fl["F"] == 11008
    assert abs(fl["fwd"] / 1e12 - 11.45) < 0.05      # 11.45 TFLOP / sample forward
    assert abs(fl["total"] / 1e12 - 33.6) < 0.1       # 33.6 TFLOP / sample fwd + bwd


class _FakeLLM(torch.nn.Module):
    """Deterministic stand-in for the inner model: next token = (last token * 7 + 3) % V."""

    def __init__(self, V=50, max_seq_len=32):
        super().__init__()
        self.args = SimpleNamespace(max_seq_len=max_seq_len, max_batch_size=4)
        self.image_words = 0
        self.w = torch.nn.Parameter(torch.zeros(1))
        self.V = V
        self.calls = []

    def forward_inference(self, tokens, start_pos, image=None):
        self.calls.append((tuple(tokens.shape), start_pos))
        nxt = (tokens[:, -1] * 7 + 3) % self.V
        return torch.nn.functional.one_hot(nxt, self.V).float() * 10


def test_generation_plan_matches_oracle_restatement():
    """meta.py:419-442 on the host: budget, left truncation, token matrix and stop list of our
    `plan_generation` equal what the oracle's restatement of `generate` sets up (the loop itself runs on
    the device kernels and is compared with the oracle in tests/test_generate_gpu.py)."""
    import a3vlm_oracle as O
    from a3vlm_b200.model.meta import plan_generation

    cases = [dict(prompts=[[1, 5, 6], [1, 9, 10, 11, 12]], max_seq_len=32, image_words=0, max_gen_len=6),
             dict(prompts=[[1] + list(range(3, 30)), [1, 4]], max_seq_len=24, image_words=7, max_gen_len=5),  # truncation
             dict(prompts=[[1, 2, 3]], max_seq_len=16, image_words=0, max_gen_len=64)]                      # budget-bound
    for c in cases:
        plan = plan_generation(c["prompts"], c["max_seq_len"], c["image_words"], c["max_gen_len"], 2, [[38, 19], [7]])
        calls = []

        class Fake:
            args = O.ModelArgs(max_seq_len=c["max_seq_len"], max_batch_size=4)
            image_words = c["image_words"]

            def forward_inference(self, tokens, start_pos, image=None):
                calls.append((tuple(tokens.shape), start_pos))
                return torch.nn.functional.one_hot((tokens[:, -1] * 7 + 3) % 50, 50).float()

        img = torch.zeros(1) if c["image_words"] else None
        _, rtoks, _ = O.generate_tokens(Fake(), c["prompts"], images=img, max_gen_len=c["max_gen_len"], eos_id=2,
                                        stop_sequences=[[38, 19], [7]])
        assert plan["total_len"] == rtoks.shape[1]
        assert plan["start_pos"] == min(len(t) for t in c["prompts"])
        for b, t in enumerate(plan["prompts"]):
            assert rtoks[b, :len(t)].tolist() == t  # the oracle wrote the same (truncated) prompt
            assert plan["tokens"][b, :len(t)].tolist() == t and plan["tokens"][b, len(t):].sum() == 0
            assert int(plan["prompt_len"][b]) == len(t)
        assert plan["stops"] == [[2], [38, 19], [7]]
        assert plan["stop_flat"].tolist() == [2, 38, 19, 7] and plan["stop_offsets"].tolist() == [0, 1, 3, 4]
        if calls:
            assert calls[0] == ((len(c["prompts"]), plan["start_pos"]), 0)


def test_reference_ensemble_checkpoint_is_refused_clearly(tmp_path):
    """a released llama_ens5 checkpoint (Q-Former + 5632-wide visual_proj) must raise an explicit
    unsupported-encoder error, not a generic shape mismatch deep in the loader."""
    import torch
    from a3vlm_b200 import checkpoint as ckpt

    d = tmp_path / "ens"
    d.mkdir()
    torch.save({"model": {"llma.visual_proj.0.weight": torch.zeros(8, 5632), "llma.qformer_proj.weight": torch.zeros(4, 4),
                          "llma.tok_embeddings.weight": torch.zeros(16, 8)}}, d / "consolidated.00-of-01.model.pth")
    with pytest.raises(NotImplementedError, match="four-encoder"):
        ckpt.check_vision_branch(str(d), 1024)
    ok = tmp_path / "vit_only"
    ok.mkdir()
    torch.save({"model": {"llma.visual_proj.0.weight": torch.zeros(8, 1024)}}, ok / "consolidated.00-of-01.model.pth")
    ckpt.check_vision_branch(str(ok), 1024)


def test_grouped_query_layout_is_pure_shape_arithmetic():
    """n_kv_heads < n_heads: wk / wv rows shrink to n_kv_heads * head_dim inside the fused q|k|v matrix (llama_ens5.py:76-90)."""
    import types

    from a3vlm_b200.engine import build_layout

    a = types.SimpleNamespace(dim=512, n_layers=2, n_heads=4, n_kv_heads=1, vocab_size=512, multiple_of=64, ffn_dim_multiplier=None)
    decay, nodecay, frozen, kpad = build_layout(a, None)
    shapes = dict(decay + nodecay)
    assert shapes["layers.0.wqkv"] == (512 + 2 * 128, 512) and shapes["layers.1.attention.wo.weight"] == (512, 512)
    a.n_kv_heads = None
    assert dict(build_layout(a, None)[0])["layers.0.wqkv"] == (3 * 512, 512)
