"""Drop-in proof, second half (GPU): replay what the reference's UNMODIFIED `MetaModel` asked of the plugin
(tests/golden/reference_seam_trace.pt, recorded by tests/golden/make_seam_trace.py from the reference's own
`__init__` / `forward` / `generate` with `llama_type="llama_b200"`) against the real CUDA
`a3vlm_b200.model.LLM.llama_b200.Transformer`:

  * the constructor is called with the `ModelArgs` the REFERENCE built from its JSON config (meta.py:35-48);
  * every recorded `forward(examples, image)` / `forward_inference(tokens, start_pos, image)` call is issued
    with the recorded arguments, in order, and the answer is compared with the recorded (oracle) answer;
  * what the reference derived from the answers (loss, generated strings, image_words, trainable set) is
    re-derived from the CUDA answers.

Tolerances: logits rel-L2 <= 2e-2 and <= 6 bf16 ulps of the largest logit (two bf16 paths, see
test_e2e_gpu.py); argmax must agree wherever the recorded top-2 margin exceeds 4 bf16 ulps of the top logit
(the oracle's own logits are bf16: smaller margins are ties of the reference itself)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def rel_l2(got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    return ((got - ref).norm() / ref.norm().clamp_min(1e-12)).item()


@pytest.fixture(scope="module")
def trace():
    return torch.load(os.path.join(ROOT, "tests", "golden", "reference_seam_trace.pt"), weights_only=False)


@pytest.fixture(scope="module")
def plugin(trace):
    import a3vlm_oracle as O
    import make_seam_trace as MS
    from a3vlm_b200.model.LLM import llama_b200 as P

    assert trace["ctor"]["args_type"] == "a3vlm_b200.model.LLM.llama_b200.ModelArgs"
    margs = P.ModelArgs(**trace["ctor"]["model_args"])          # exactly what the reference passed
    model = P.Transformer(margs, with_visual=trace["ctor"]["with_visual"])
    a, v = MS.oracle_args(margs)
    sd = O.init_params(a, v, seed=trace["seed"], output_scale=trace["output_scale"])
    model.load_state_dict(sd)
    return model, margs


def test_reference_ctor_contract(trace, plugin):
    model, margs = plugin
    assert model.image_words == trace["image_words"]
    assert model.args is margs and model.image_size == margs.image_size
    got = sorted("llma." + n for n in model.get_trainable_params().keys())
    assert got == trace["trainable"]
    # the reference flips requires_grad by name (meta.py:228-232): same names, same flags
    flags = {"llma." + n: bool(p.requires_grad) for n, p in model.named_parameters()}
    assert flags == trace["requires_grad"]
    assert [c.__name__ for c in model.get_basic_block_classes()] == ["TransformerBlock"]
    assert type(model.layers[0]).__name__ == "TransformerBlock" and len(model.layers) == margs.n_layers


def test_replay_reference_calls(trace, plugin):
    model, _ = plugin
    worst = 0.0
    for k, c in enumerate(trace["calls"]):
        if c["fn"] == "forward":
            with torch.no_grad():
                got = model(c["examples"], c["image"])
        else:
            got = model.forward_inference(c["tokens"], c["start_pos"], c["image"])
            assert got.dtype == torch.float32 and got.shape == c["logits"].shape
        ref = c["logits"]
        r = rel_l2(got, ref)
        worst = max(worst, r)
        mx = (got.float().cpu() - ref).abs().max().item()
        assert r <= 2e-2, (k, c["fn"], r)
        assert mx <= 6 * 2.0 ** -8 * ref.abs().max().item(), (k, c["fn"], mx)
        if c["fn"] == "forward_inference":
            top = ref.max(dim=-1).values.abs().clamp_min(1.0)
            decided = c["margin"] > 4 * 2.0 ** -8 * top
            same = got.float().cpu().argmax(-1) == ref.argmax(-1)
            assert bool((same | ~decided).all()), (k, c["start_pos"], c["margin"], got.float().cpu().argmax(-1), ref.argmax(-1))
    print(f"replayed {len(trace['calls'])} reference calls, worst rel-L2 {worst:.3e}")


def test_reference_results_rederived(trace, plugin):
    """The reference's own post-processing of the answers, applied to the CUDA answers: the loss of
    meta.py:256-262 and the strings of meta.py:479-484."""
    from a3vlm_b200.model import MetaModel
    import make_reference_golden as RG

    model, margs = plugin
    fc = trace["forward_case"]
    call = trace["calls"][fc["call"]]
    assert call["examples"].shape[1] == 9  # the reference trimmed the 3-column pad tail before calling the plugin
    with torch.no_grad():
        out = model(call["examples"], call["image"]).float().cpu()
    labels = fc["labels"][:, :9]
    loss = torch.nn.functional.cross_entropy(out[:, :-1].reshape(-1, out.shape[-1]), labels[:, 1:].flatten(), ignore_index=0)
    assert abs(loss.item() - fc["loss"].item()) <= 2e-3 * abs(fc["loss"].item()), (loss.item(), fc["loss"].item())

    # our facade over the same inner model must give the strings the reference produced
    cfg = {k: v for k, v in trace["ctor"]["model_args"].items() if k not in ("max_seq_len", "max_batch_size")}
    facade = MetaModel("llama_b200", cfg, None, with_visual=True, max_seq_len=margs.max_seq_len, max_batch_size=32)
    facade.llma.load_state_dict(model.state_dict())
    facade.tokenizer = RG.ToyTokenizer()
    floss, _ = facade(fc["examples"].cuda(), fc["labels"].cuda(), images=fc["images"].cuda())
    assert abs(floss.item() - fc["loss"].item()) <= 2e-3 * abs(fc["loss"].item())
    for case in (trace["generate_case"], trace["generate_text_case"]):
        res = facade.generate(case["prompts"], case.get("images"), None, case["max_gen_len"], 0.0, 0.95, [])
        lo, hi = case["calls"]
        calls = trace["calls"][lo:hi]
        min_prompt = min(len(p.split()) + 1 for p in case["prompts"])
        for r, (g, e) in enumerate(zip(res, case["result"])):
            if g == e:
                continue
            gt, et = g.split(), e.split()
            k = next(i for i, (x, y) in enumerate(zip(gt, et)) if x != y)
            step = len(case["prompts"][r].split()) + 1 + k - min_prompt  # the forward_inference call that chose token k
            c = calls[step]
            top = c["logits"][r].max().abs().clamp_min(1.0).item()
            print(f"row {r} parts ways at generated token {k}: reference margin {c['margin'][r].item():.4f} at logit {top:.2f}")
            assert c["margin"][r].item() <= 4 * 2.0 ** -8 * top, (r, k, g, e)
