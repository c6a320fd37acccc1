"""CPU tests of the host-side logic that does not need the CUDA library: flat-buffer layout and
bucket plan, the data-parallel gradient exchange (world_size 2 over gloo), rope table, the FLOP model
of bench.py against SURVEY 8d, and MetaModel.generate's control flow driven by a fake inner model."""
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
    assert kpad == 592
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


def _dp_worker(rank, world, port, total, plan, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from a3vlm_b200.dp import BucketAllReduce

    g = torch.Generator().manual_seed(rank)
    # local gradients already carry the 1/world factor (folded into the loss gradient)
    local = torch.randn(total, generator=g) / world
    flat = local.clone()
    r = BucketAllReduce(flat, None, None)
    for lo, hi in plan:
        r.on_bucket(lo, hi)
    r.finish()
    expect = sum(torch.randn(total, generator=torch.Generator().manual_seed(k)) for k in range(world)) / world
    ok = torch.allclose(flat, expect, atol=1e-6) and sorted(r.covered) == sorted(plan)
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_gradient_exchange_world2_gloo():
    """N > 1 path on CPU: two ranks, the bucketed SUM all-reduce over the engine's bucket plan yields
    the mean gradient on both ranks."""
    from a3vlm_b200.engine import bucket_plan, build_layout, flat_offsets

    a = _args()
    decay, nodecay, _, _ = build_layout(a, _vis())
    offs, total = flat_offsets(decay + nodecay)
    plan = bucket_plan(offs, total, a.n_layers)
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_dp_worker, args=(2, port, total, plan, out), nprocs=2, join=True)
    assert out[0] and out[1]


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


def test_generate_control_flow_matches_oracle_restatement():
    """meta.py:379-485 on the host: our MetaModel.generate_tokens and the oracle's restatement walk the
    same prompts / stops identically when driven by the same (fake) logits."""
    import a3vlm_oracle as O
    from a3vlm_b200.model.meta import MetaModel

    m = MetaModel.__new__(MetaModel)
    torch.nn.Module.__init__(m)
    m.llma = _FakeLLM()
    m.tokenizer = None
    m._argmax = lambda logits: torch.argmax(logits, dim=-1)  # CPU stand-in for the argmax kernel
    prompts = [[1, 5, 6], [1, 9, 10, 11, 12]]
    out, toks, sp = m.generate_tokens(prompts, None, max_gen_len=6, eos_id=2, stop_sequences=[[38, 19]])
    fake = _FakeLLM()
    orc = SimpleNamespace(args=O.ModelArgs(max_seq_len=32, max_batch_size=4), image_words=0,
                          forward_inference=fake.forward_inference)
    ref, rtoks, rsp = O.generate_tokens(orc, prompts, max_gen_len=6, eos_id=2, stop_sequences=[[38, 19]])
    assert out == ref and torch.equal(toks, rtoks) and torch.equal(sp, rsp)
    assert m.llma.calls == fake.calls
    assert m.llma.calls[0] == ((2, 3), 0) and all(c[0] == (2, 1) for c in m.llma.calls[1:])
