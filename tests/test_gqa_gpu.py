"""Grouped-query attention (n_kv_heads < n_heads; reference LLM/llama_ens5.py:66,76-90,121-133: wk / wv and the KV
cache carry n_kv_heads heads, `repeat_kv` expands them for the attention product).  Unused by the 7B / 13B configs of
BASELINE.json, part of row A1 of the scope table (weights wk / wv of shape (Hkv * 128, D)).

Kernel level: the repeat / reduce glue, the decode attention reading group heads, the decode-step RoPE + cache append and
the fused QKV decode projection with Hkv key / value heads -- each against a torch restatement.
Model level: a 3-layer model with 4 query heads over 2 key / value heads against the CPU oracle with the same weights:
logits, loss, per-parameter gradients (incl. wk / wv, whose gradient is the sum over the group), prefill / decode
through the cache, greedy generation at batch 1 (fused weight-streaming decode) and batch 3 (tcgen05 decode)."""
import math
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
bf16 = torch.bfloat16


def dev():
    return torch.device("cuda", 0)


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device=dev()) * scale).to(bf16)


@pytest.fixture(scope="module")
def ops():
    from a3vlm_b200 import ops as _ops

    return _ops


def rel_l2(got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    return ((got - ref).norm() / ref.norm().clamp_min(1e-12)).item()


@pytest.mark.parametrize("H,Hkv", [(4, 2), (8, 1), (4, 4)])
def test_repeat_and_reduce_kv(ops, H, Hkv):
    torch.manual_seed(H * 10 + Hkv)
    rows, hd = 77, 128
    D, Dkv, n_rep = H * hd, Hkv * hd, H // Hkv
    c = rnd(rows, D + 2 * Dkv)
    full = torch.empty(rows, 3 * D, device=dev(), dtype=bf16)
    ops.repeat_kv(c, full, H, Hkv, hd)
    q, k, v = c[:, :D], c[:, D:D + Dkv].view(rows, Hkv, hd), c[:, D + Dkv:].view(rows, Hkv, hd)
    ref = torch.cat([q, k.repeat_interleave(n_rep, 1).reshape(rows, D), v.repeat_interleave(n_rep, 1).reshape(rows, D)], 1)
    assert torch.equal(full, ref)
    d = rnd(rows, 3 * D)
    dc = torch.empty(rows, D + 2 * Dkv, device=dev(), dtype=bf16)
    ops.reduce_kv(d, dc, H, Hkv, hd)
    dk = d[:, D:2 * D].float().view(rows, Hkv, n_rep, hd).sum(2).reshape(rows, Dkv)
    dv = d[:, 2 * D:].float().view(rows, Hkv, n_rep, hd).sum(2).reshape(rows, Dkv)
    assert torch.equal(dc[:, :D], d[:, :D])
    assert torch.equal(dc[:, D:D + Dkv], dk.to(bf16)) and torch.equal(dc[:, D + Dkv:], dv.to(bf16))


def test_decode_kernels_with_group_heads(ops):
    torch.manual_seed(5)
    B, H, Hkv, hd, max_seq, pos = 3, 4, 2, 128, 64, 21
    D, Dkv, n_rep = H * hd, Hkv * hd, H // Hkv
    table = ops.rope_table(128, 128, device=dev())
    qkv = rnd(B, D + 2 * Dkv)
    kc, vc = rnd(B, max_seq, Hkv, hd), rnd(B, max_seq, Hkv, hd)
    kc0, vc0 = kc.clone(), vc.clone()
    pos_dev = torch.tensor([pos], device=dev(), dtype=torch.int32)
    ref = qkv.clone()
    ops.rope_apply(ref[:, :D + Dkv], H + Hkv, hd, table, 1, pos)
    got = qkv.clone()
    ops.decode_rope_kv(got, kc, vc, B, H, hd, table, 0, max_seq, pos_dev=pos_dev, n_kv_heads=Hkv)
    assert torch.equal(got, ref)
    assert torch.equal(kc[:, pos].reshape(B, Dkv), ref[:, D:D + Dkv]) and torch.equal(vc[:, pos].reshape(B, Dkv), ref[:, D + Dkv:])
    kc0[:, pos], vc0[:, pos] = kc[:, pos], vc[:, pos]
    assert torch.equal(kc, kc0) and torch.equal(vc, vc0)  # nothing else was touched
    # decode attention: query head h reads group head h // n_rep
    out = torch.empty(B, D, device=dev(), dtype=bf16)
    scratch = torch.zeros(B * H * (32 * 130 + 1), device=dev())
    scale = 1.0 / math.sqrt(hd)
    ops.attn_decode(got[:, :D], kc, vc, out, scratch, B, H, hd, max_seq, max_seq, scale, pos_dev=pos_dev, n_kv_heads=Hkv)
    qf = got[:, :D].float().view(B, H, hd)
    kf = kc[:, :pos + 1].float().repeat_interleave(n_rep, 2)  # [B, ctx, H, hd]
    vf = vc[:, :pos + 1].float().repeat_interleave(n_rep, 2)
    p = torch.softmax(torch.einsum("bhd,bthd->bht", qf, kf) * scale, -1)
    ref_o = torch.einsum("bht,bthd->bhd", p, vf).reshape(B, D)
    torch.testing.assert_close(out.float(), ref_o, rtol=2 ** -7, atol=2e-3)
    # fused QKV decode projection (RMSNorm prologue, RoPE + cache append epilogue) == the unfused sequence
    K = 256
    x, nw, w = rnd(B, K), (1 + 0.1 * torch.randn(K, device=dev())).to(bf16), rnd(D + 2 * Dkv, K, scale=0.06)
    kc1, vc1 = torch.zeros_like(kc), torch.zeros_like(vc)
    kc2, vc2 = torch.zeros_like(kc), torch.zeros_like(vc)
    ref = ops.linear_small_m(ops.rmsnorm_fwd(x, nw, 1e-5), w)
    ops.decode_rope_kv(ref, kc2, vc2, B, H, hd, table, pos, max_seq, n_kv_heads=Hkv)
    got = torch.empty_like(ref)
    ops.decode_linear(x, w, got, norm_w=nw, eps=1e-5, rope=table, n_heads=H, n_kv_heads=Hkv, kcache=kc1, vcache=vc1,
                      pos_dev=pos_dev, max_seq=max_seq)
    torch.testing.assert_close(got.float(), ref.float(), rtol=2 ** -7, atol=4e-3)
    torch.testing.assert_close(kc1.float(), kc2.float(), rtol=2 ** -7, atol=4e-3)
    torch.testing.assert_close(vc1.float(), vc2.float(), rtol=2 ** -7, atol=4e-3)
    assert torch.equal(kc1[:, pos].reshape(B, Dkv), got[:, D:D + Dkv]) and torch.equal(vc1[:, pos].reshape(B, Dkv), got[:, D + Dkv:])


def test_gqa_model_matches_oracle():
    import a3vlm_oracle as O
    from a3vlm_b200.model import MetaModel
    from a3vlm_b200.train import AdamW, NativeScalerWithGradNormCount

    a = O.ModelArgs(dim=512, n_layers=3, n_heads=4, n_kv_heads=2, vocab_size=1024, multiple_of=128, max_seq_len=128,
                    max_batch_size=8)
    v = O.VisionArgs(image_size=56, patch=14, width=256, layers=2, heads=4, mlp=512, quick_gelu=True)
    sd = O.init_params(a, v, seed=31, output_scale=6.0)
    assert sd["layers.0.attention.wk.weight"].shape == (256, 512) and sd["layers.0.attention.wq.weight"].shape == (512, 512)
    cfg = dict(dim=a.dim, n_layers=a.n_layers, n_heads=a.n_heads, n_kv_heads=a.n_kv_heads, vocab_size=a.vocab_size,
               multiple_of=a.multiple_of, norm_eps=a.norm_eps, image_size=v.image_size, vit_patch=v.patch, vit_width=v.width,
               vit_layers=v.layers, vit_heads=v.heads, vit_mlp=v.mlp, vit_quick_gelu=v.quick_gelu)
    m = MetaModel("llama_b200", cfg, None, with_visual=True, max_seq_len=a.max_seq_len, max_batch_size=a.max_batch_size)
    missing, unexpected = m.llma.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    assert m.llma.state_dict()["layers.1.attention.wv.weight"].shape == (256, 512)
    g = torch.Generator().manual_seed(4)
    B, T = 3, 36
    tok = torch.randint(3, 1024, (B, T), generator=g)
    tok[:, 0] = 1
    lab = tok.clone()
    lab[:, :9] = 0
    img = torch.randn(B, 3, 56, 56, generator=g)
    train = [k for k in sd if not k.startswith("clip.")]
    for k in train:
        sd[k].requires_grad_(True)
    orc = O.OracleTransformer(a, v, sd)
    ref_loss = O.meta_forward_loss(orc, tok, lab, img)
    ref_loss.backward()
    exact = O.OracleTransformer(a, v, sd, act_dtype=torch.float32)
    with torch.no_grad():
        ref_logits, exact_logits = orc.forward(tok, img), exact.forward(tok, img)
        got_logits = m.llma(tok, img)
    e_ref, e_got = rel_l2(ref_logits, exact_logits), rel_l2(got_logits, exact_logits)
    print(f"gqa: |oracle_bf16 - fp32| = {e_ref:.3e}, |cuda - fp32| = {e_got:.3e}, |cuda - oracle_bf16| = {rel_l2(got_logits, ref_logits):.3e}")
    assert e_got <= 1.25 * e_ref + 1e-3
    assert rel_l2(got_logits, ref_logits) <= 2e-2
    m.zero_grad(set_to_none=True)
    loss, _ = m(tok.cuda(), lab.cuda(), images=img.cuda())
    assert abs(loss.item() - ref_loss.item()) <= 2e-3 * abs(ref_loss.item())
    opt = AdamW(m, lr=1e-4, betas=(0.9, 0.95), weight_decay=0.0)
    norm = NativeScalerWithGradNormCount()(loss, opt, m, clip_grad=1.0, update_grad=True)
    ref_total = math.sqrt(sum(float(sd[k].grad.float().pow(2).sum()) for k in train))
    assert abs(norm.item() - ref_total) <= 1e-2 * ref_total, (norm.item(), ref_total)
    named = dict(m.llma.named_parameters())
    for k in ("layers.0.attention.wk.weight", "layers.2.attention.wv.weight", "layers.1.attention.wq.weight",
              "layers.0.attention.wo.weight", "layers.1.feed_forward.w2.weight", "output.weight", "visual_proj.0.weight"):
        r = rel_l2(named[k].grad, sd[k].grad)
        print(f"gqa grad {k}: rel l2 {r:.3e}")
        assert r <= 3e-2, (k, r)
    for k in train:
        sd[k].requires_grad_(False)

    # prefill / decode through the Hkv-head cache and greedy generation, batch 1 (fused decode) and batch 3 (tcgen05)
    m2 = MetaModel("llama_b200", cfg, None, with_visual=True, max_seq_len=a.max_seq_len, max_batch_size=a.max_batch_size)
    m2.llma.load_state_dict(sd, strict=True)
    orc = O.OracleTransformer(a, v, sd)
    assert m2.llma.engine.n_kv == 2
    for prompts, im in (([[1, 5, 9, 11]], img[:1]), ([[1, 5, 9], [1, 7, 8, 9, 10], [1, 3]], img)):
        seen = []
        real_fi = orc.forward_inference

        def recording_fi(tokens, start_pos, image=None, _real=real_fi, _seen=seen):
            lg = _real(tokens, start_pos, image)
            _seen.append(lg.clone())
            return lg

        orc.forward_inference = recording_fi
        ref_gen, _, _ = O.generate_tokens(orc, prompts, images=im, max_gen_len=8, eos_id=2)
        orc.forward_inference = real_fi
        got_gen, _, _ = m2.generate_tokens(prompts, im, max_gen_len=8, eos_id=2)
        assert m2.llma.engine.kcache[0].shape[2] == 2
        min_prompt = min(len(q) for q in prompts)
        for r, (gg, ee) in enumerate(zip(got_gen, ref_gen)):
            if gg == ee:
                continue
            k = next(i for i, (x, y) in enumerate(zip(gg, ee)) if x != y)
            top2 = seen[len(prompts[r]) + k - min_prompt][r].topk(2).values
            margin, mag = (top2[0] - top2[1]).item(), top2[0].abs().item()
            assert margin <= 4 * 2.0 ** -8 * max(mag, 1.0), (r, k, margin, mag, gg, ee)
        # teacher-forced logits of the first decode step against the oracle
        P = len(prompts[0])
        if len(prompts) == 1:
            m2.llma.engine.destroy_kv_cache()
            orc.destroy_kv_cache() if hasattr(orc, "destroy_kv_cache") else None
            t0 = torch.tensor(prompts)
            r0 = orc.forward_inference(t0, 0, im)
            g0 = m2.llma.forward_inference(t0, 0, im)
            nxt = r0.argmax(-1, keepdim=True)
            r1 = orc.forward_inference(nxt, P)
            g1 = m2.llma.forward_inference(nxt, P)
            assert rel_l2(g0, r0) <= 2e-2 and rel_l2(g1, r1) <= 2e-2, (rel_l2(g0, r0), rel_l2(g1, r1))
