"""BASELINE config 1 / 2 at FULL size on the GPU box: one 336x336 image + 32-token prompt, ViT-L/14 +
Llama-2-7B geometry, random init, greedy decode of 32 tokens through the reference-facing
`MetaModel.generate_tokens` (device-side generation loop, CUDA-graph decode) against the CPU oracle with
the SAME weights (generated on the GPU, copied to the host once).

Why the weights are not simply scaled: with random weights the 32 000 logits are close to i.i.d. Gaussian
(std ~1.3-1.8), so the top-2 gap is ~exponential with mean ~0.2 std and is 0 in ~10 % of the steps once the
reference rounds the logits to bf16 (M/model/LLM/llama_ens5.py:530-531).  Both the gap and the CUDA-vs-oracle
rounding noise (rel. 4e-2 after 32 bf16 layers, max |diff| ~0.3) are LINEAR in `output.weight`, so no
`output_scale` can separate them (SURVEY 8d assumed an absolute noise floor).  Hence two parts:

  1. natural random weights: the oracle runs free; the CUDA path is teacher-forced on the oracle's tokens
     so every one of the 32 steps is checked -- top-1 must agree except where the oracle's own top-2 margin
     is within the logit difference observed at that step (a tie of the reference itself); the free-running
     CUDA ids must equal the oracle's up to the first such tie.
  2. a peaked LM head that makes the decision margin >= 10x the observed logit noise at EVERY step:
     output.weight = W_rand + alpha * tok_embeddings[perm] with the residual branches (wo, w2) scaled by gamma so
     the current token's embedding stays visible in the final hidden state; the next token is then decided with
     a margin of several logit-std while every layer, the KV cache (positions 611 .. 643) and the LM head still
     run at full size.  Free-running ids must be EXACT for all 32 tokens, margin and noise are printed and asserted.
"""
import os
import sys
import time

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
bf16 = torch.bfloat16
NG = 32
CFG = dict(dim=4096, n_layers=32, n_heads=32, vocab_size=32000, multiple_of=256, norm_eps=1e-5)


def _run_pair(model, orc, prompt, img):
    """Oracle free-running greedy (records its logits), CUDA free-running through generate_tokens (records its
    logits), CUDA teacher-forced on the oracle's tokens.  Returns per-step records."""
    import a3vlm_oracle as O

    seen = []
    real = orc.forward_inference

    def rec(tokens, start_pos, image=None):
        lg = real(tokens, start_pos, image)
        seen.append(lg.float().clone())
        return lg

    orc.forward_inference = rec
    t0 = time.time()
    ref_gen, _, _ = O.generate_tokens(orc, [prompt], images=img, max_gen_len=NG, eos_id=-1)
    t_cpu = time.time() - t0
    orc.forward_inference = real
    got_gen, _, _ = model.generate_tokens([prompt], img, max_gen_len=NG, eos_id=-1)
    # teacher-forced CUDA logits on the oracle's tokens
    P = len(prompt)
    tf = [model.llma.forward_inference(torch.tensor([prompt]), 0, img).float().cpu()]
    for i in range(NG - 1):
        tf.append(model.llma.forward_inference(torch.tensor([[ref_gen[0][i]]]), P + i).float().cpu())
    steps = []
    for i, (r, c) in enumerate(zip(seen, tf)):
        top2 = r[0].topk(2).values
        steps.append(dict(step=i, ref_top1=int(r.argmax(-1)), cuda_top1=int(c.argmax(-1)), margin=float(top2[0] - top2[1]),
                          diff=float((r - c).abs().max()), std=float(r.std())))
    return ref_gen[0], got_gen[0], steps, t_cpu


def test_7b_greedy_32_tokens_full_size():
    import a3vlm_oracle as O
    from a3vlm_b200.model import MetaModel

    torch.set_num_threads(os.cpu_count() or 1)
    a = O.ModelArgs(**CFG, max_seq_len=1024, max_batch_size=1)
    v = O.VisionArgs(image_size=336)
    model = MetaModel("llama_b200", {**CFG, "image_size": 336}, None, with_visual=True, max_seq_len=1024, max_batch_size=1)
    model.llma.reset_parameters(seed=2)  # kaiming-uniform family of the reference (llama_ens5.py:28), on the device
    t0 = time.time()
    sd = {k: t.detach().cpu() for k, t in model.llma.state_dict().items()}
    print(f"7B weights to the host: {time.time() - t0:.1f}s, {os.cpu_count()} cores")
    g = torch.Generator().manual_seed(1)
    prompt = torch.randint(3, 32000, (32,), generator=g).tolist()
    prompt[0] = 1
    img = torch.randn(1, 3, 336, 336, generator=g)

    # ---------------- part 1: natural random weights, tie-aware
    orc = O.OracleTransformer(a, v, sd, native_linear=True)
    ref, got, steps, t_cpu = _run_pair(model, orc, prompt, img)
    ties = []
    for s in steps:
        if s["ref_top1"] != s["cuda_top1"]:
            assert s["margin"] <= s["diff"], ("top-1 differs at a step the oracle decided clearly", s)
            ties.append(s["step"])
    agree = sum(s["ref_top1"] == s["cuda_top1"] for s in steps)
    first_tie = next((s["step"] for s in steps if s["margin"] <= s["diff"]), NG)
    prefix = next((i for i, (x, y) in enumerate(zip(got, ref)) if x != y), NG)
    print(f"part 1 (natural weights): teacher-forced top-1 agreement {agree}/{NG} (misses at oracle near-ties {ties}); "
          f"free-running prefix match {prefix}, first oracle near-tie at step {first_tie}; logit std {steps[0]['std']:.2f}, "
          f"max |cuda - oracle| {max(s['diff'] for s in steps):.3f}; oracle {t_cpu:.1f}s on the host")
    assert agree >= NG - 8, steps           # near-ties are ~10-20 % of the steps with bf16-rounded random logits
    assert prefix >= min(first_tie, NG), (prefix, first_tie, got, ref)

    # ---------------- part 2: peaked LM head, margin >= 10x noise, token-exact free running
    gamma, alpha = 0.02, 0.25
    views = model.llma.engine.param_views()
    perm = torch.randperm(32000, generator=torch.Generator().manual_seed(7))
    with torch.no_grad():
        for i in range(CFG["n_layers"]):
            for n in (f"layers.{i}.attention.wo.weight", f"layers.{i}.feed_forward.w2.weight"):
                views[n].mul_(gamma)
                sd[n] = views[n].detach().cpu()
        # kaiming-uniform rows have norm 1/sqrt(3); N(0, 0.02)-sized embeddings keep alpha comparable to the calibration
        emb = views["tok_embeddings.weight"]
        emb.copy_((emb.float() * (0.02 * (3 * CFG["dim"]) ** 0.5)).to(bf16))
        views["output.weight"].copy_((views["output.weight"].float() * (0.02 * (3 * CFG["dim"]) ** 0.5)
                                      + alpha * emb[perm.cuda()].float()).to(bf16))
        sd["tok_embeddings.weight"] = emb.detach().cpu()
        sd["output.weight"] = views["output.weight"].detach().cpu()
    model.llma.engine.weights_changed()
    model.llma.engine.destroy_kv_cache()
    orc = O.OracleTransformer(a, v, sd, native_linear=True)
    ref, got, steps, t_cpu = _run_pair(model, orc, prompt, img)
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(32000)
    min_margin = min(s["margin"] for s in steps)
    noise = max(s["diff"] for s in steps)
    print(f"part 2 (peaked head, gamma {gamma}, alpha {alpha}): min top-2 margin {min_margin:.3f}, max |cuda - oracle| logit "
          f"difference {noise:.4f} (ratio {min_margin / max(noise, 1e-9):.1f}), logit std {steps[0]['std']:.2f}")
    print("oracle ids:", ref)
    print("cuda   ids:", got)
    assert min_margin >= 10 * noise, (min_margin, noise)
    assert got == ref and len(got) == NG
    assert len(set(got)) > NG // 2  # a walk through the vocabulary, not a fixed point
