"""Edge cases of the reference-facing API on the CUDA path, each against the oracle's restatement of the
reference behaviour: nothing to predict (meta.py:244-247,259-260), batch 1 / one-token prompts, generation
budgets (max_gen_len 0, left truncation with image words, meta.py:423-429), cache overflow, batch-size changes
between prefill and decode."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


@pytest.fixture(scope="module")
def pair():
    import a3vlm_oracle as O
    from test_generate_gpu import _build, _tiny

    a, v, sd = _tiny(False, seed=29)
    return _build("llama_b200", a, v, sd), O.OracleTransformer(a, v, sd), a, v


def test_nothing_to_predict_gives_zero_loss_and_zero_gradients(pair, capsys):
    import a3vlm_oracle as O

    m, orc, a, v = pair
    g = torch.Generator().manual_seed(1)
    tok = torch.randint(3, 512, (2, 9), generator=g)
    tok[:, 0] = 1
    lab = torch.zeros_like(tok)
    img = torch.randn(2, 3, 28, 28, generator=g)
    with torch.no_grad():
        ref = O.meta_forward_loss(orc, tok, lab, img)
    assert ref.item() == 0.0
    m.zero_grad(set_to_none=True)
    loss, _ = m(tok.cuda(), lab.cuda(), images=img.cuda())
    assert "nothing to predict" in capsys.readouterr().out
    assert loss.item() == 0.0
    loss.backward()
    for name, p in m.llma.named_parameters():
        if p.grad is not None:
            assert float(p.grad.float().abs().max()) == 0.0, name


def test_batch_one_single_token_prompt_with_image(pair):
    import a3vlm_oracle as O
    from test_e2e_gpu import check

    m, orc, a, v = pair
    g = torch.Generator().manual_seed(2)
    img = torch.randn(1, 3, 28, 28, generator=g)
    one = torch.tensor([[1]])
    check("prefill of a lone BOS + image", m.llma.forward_inference(one, 0, img), orc.forward_inference(one, 0, img))
    nxt = torch.tensor([[77]])
    check("decode after it", m.llma.forward_inference(nxt, 1), orc.forward_inference(nxt, 1))
    with torch.no_grad():
        check("training forward, T = 1", m.llma(one, img), orc.forward(one, img))
        check("text only, T = 2, B = 1", m.llma(torch.tensor([[1, 5]])), orc.forward(torch.tensor([[1, 5]])))


def test_generation_budget_edges_match_oracle(pair):
    import a3vlm_oracle as O

    m, orc, a, v = pair
    g = torch.Generator().manual_seed(3)
    img = torch.randn(2, 3, 28, 28, generator=g)
    # max_gen_len = 0: the loop only force-feeds the longer prompt; nothing is generated for it
    prompts = [[1, 17], [1, 100, 101, 102, 103]]
    ref, rtoks, rsp = O.generate_tokens(orc, prompts, images=img, max_gen_len=0, eos_id=-1)
    out, toks, sp = m.generate_tokens(prompts, img, max_gen_len=0, eos_id=-1)
    assert [len(x) for x in out] == [len(x) for x in ref] and out[1] == ref[1] == []
    assert toks.shape == rtoks.shape and torch.equal(sp.cpu(), rsp)
    # a prompt longer than the budget (max_seq_len 64 - 7 image words - max_gen_len 40 = 17 tokens) is cut from the left
    long = [1] + [3 + (i * 7) % 500 for i in range(40)]
    ref, rtoks, _ = O.generate_tokens(orc, [long, [1, 9]], images=img, max_gen_len=40, eos_id=-1)
    out, toks, _ = m.generate_tokens([long, [1, 9]], img, max_gen_len=40, eos_id=-1)
    assert toks.shape == rtoks.shape == (2, 57)
    assert torch.equal(toks[0, :17].cpu(), rtoks[0, :17]) and toks[0, 0].item() == long[-17]  # same truncated prompt
    assert len(out[0]) == len(ref[0]) == 40


def test_cache_overflow_and_batch_change_are_loud(pair):
    m, orc, a, v = pair
    tok = torch.ones(2, 60, dtype=torch.long)
    img = torch.zeros(2, 3, 28, 28)
    with pytest.raises(ValueError, match="exceeds max_seq_len"):
        m.llma.forward_inference(tok, 0, img)  # 60 + 7 image words > 64
    m.llma.forward_inference(tok[:, :8], 0, img)
    with pytest.raises(ValueError, match="batch size changed"):
        m.llma.forward_inference(tok[:1, :1], 8)
    with pytest.raises(ValueError):
        m.llma.forward_inference(tok[:, :1], 64)  # position 64 + 7 is outside the cache
    with pytest.raises(AssertionError):
        m.generate_tokens([[1]] * 5, None, max_gen_len=2, eos_id=2)  # bsz > max_batch_size (meta.py:415)
