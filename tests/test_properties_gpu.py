"""Size-independent properties checked at BASELINE.json's FULL sizes (7B decoder shapes: B 8, L 835,
M = 6680 rows, D 4096, F 11008, H 32 x 128; ViT-L/14@336: N 577, 16 x 64), where running the CPU
oracle is too slow for a test: linearity of the tcgen05 GEMMs, softmax rows summing to one, causality,
scale invariance of RMSNorm, RoPE as an isometry with an exact inverse, the gated-SiLU fusion being the
composition of its parts, AdamW fixed points, split-KV decode == flash prefill."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16
B, L, D, F, H, HD = 8, 835, 4096, 11008, 32, 128
M = B * L


@pytest.fixture(scope="module")
def ops():
    from a3vlm_b200 import ops as o

    o._L()
    return o


def dev():
    return torch.device("cuda:0")


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device=dev()) * scale).to(bf16)


def test_gemm_linearity_and_transpose_forms_full_size(ops):
    torch.manual_seed(0)
    x1, x2, w = rnd(M, D), rnd(M, D), rnd(3 * D, D, scale=0.02)
    y1, y2 = ops.gemm(x1, w, out_fp32=True), ops.gemm(x2, w, out_fp32=True)
    xs = (x1.float() + x2.float()).to(bf16)  # exact when representable: use x2 = x1 * 2^k below instead
    y12 = ops.gemm(x1, w, out_fp32=True, alpha=3.0)
    torch.testing.assert_close(y12, 3.0 * y1, rtol=1e-6, atol=1e-6)  # alpha is applied to the fp32 accumulator
    y4 = ops.gemm((x1.float() * 4).to(bf16), w, out_fp32=True)        # power-of-two scaling is exact in bf16
    assert torch.equal(y4, 4 * y1)
    # C += A B^T accumulates: two halves of K give the whole
    h = D // 2
    acc = ops.gemm(x1[:, :h], w[:, :h], out_fp32=True)
    acc = ops.gemm(x1[:, h:], w[:, h:], out=acc, out_fp32=True, accumulate=True)
    torch.testing.assert_close(acc, y1, rtol=1e-4, atol=2e-3)
    # dgrad / wgrad forms against the forward form on transposed copies (same products, same fp32 sums)
    dy = rnd(M, 3 * D, scale=0.1)
    dx = ops.gemm(dy, w, b_mn=True, out_fp32=True)                      # dy @ W
    dx_ref = ops.gemm(dy, w.t().contiguous(), out_fp32=True)
    torch.testing.assert_close(dx, dx_ref, rtol=1e-4, atol=2e-3)
    dw = ops.gemm(dy, x1, a_mn=True, b_mn=True, out_fp32=True)          # dy^T @ x
    dw_ref = ops.gemm(dy.t().contiguous(), x1.t().contiguous(), out_fp32=True)
    torch.testing.assert_close(dw, dw_ref, rtol=1e-4, atol=5e-3)
    del y2, xs


def test_attention_rows_sum_to_one_and_causality_full_size(ops):
    torch.manual_seed(1)
    W = H * HD
    qkv = rnd(M, 3 * W)
    qkv[:, 2 * W:] = 1.0  # V = 1 -> every output element is sum_j P_ij = 1
    sc = 1 / math.sqrt(HD)
    lse = torch.empty(B, H, L, device=dev())
    o = ops.attn_fwd(qkv[:, :W], qkv[:, W:2 * W], qkv[:, 2 * W:], B, H, L, L, HD, sc, True, lse=lse)
    assert (o.float() - 1).abs().max().item() <= 2 ** -7
    # causality: perturbing the last 100 tokens leaves the first L-100 outputs bit-identical
    qkv = rnd(M, 3 * W)
    o1 = ops.attn_fwd(qkv[:, :W], qkv[:, W:2 * W], qkv[:, 2 * W:], B, H, L, L, HD, sc, True).clone()
    q3 = qkv.view(B, L, 3 * W).clone()
    q3[:, L - 100:] = rnd(B, 100, 3 * W)
    q2 = q3.view(M, 3 * W)
    o2 = ops.attn_fwd(q2[:, :W], q2[:, W:2 * W], q2[:, 2 * W:], B, H, L, L, HD, sc, True)
    assert torch.equal(o1.view(B, L, W)[:, :L - 100], o2.view(B, L, W)[:, :L - 100])
    # ViT shape, no mask, V = 1
    Bv, Hv, N, hd = 8, 16, 577, 64
    Wv = Hv * hd
    x = rnd(Bv * N, 3 * Wv)
    x[:, 2 * Wv:] = 1.0
    ov = ops.attn_fwd(x[:, :Wv], x[:, Wv:2 * Wv], x[:, 2 * Wv:], Bv, Hv, N, N, hd, 0.125, False)
    assert (ov.float() - 1).abs().max().item() <= 2 ** -7
    # backward with V = const: dQ = dK = 0 when dO is orthogonal... simpler invariant: sum_j dS_ij = 0
    # => dK summed over keys weighted by nothing is not zero; instead check dV = P^T dO against fwd:
    # for dO = 1 and V arbitrary, dV[j] = sum_i P_ij (column sums of P), and sum_j dV[j] = #queries
    qkv = rnd(M, 3 * W)
    o = ops.attn_fwd(qkv[:, :W], qkv[:, W:2 * W], qkv[:, 2 * W:], B, H, L, L, HD, sc, True, lse=lse)
    dout = torch.ones(M, W, device=dev(), dtype=bf16)
    dqkv = torch.empty_like(qkv)
    delta = torch.empty(B, H, L, device=dev())
    dq_acc = torch.empty(M, W, device=dev())
    ops.attn_bwd(qkv[:, :W], qkv[:, W:2 * W], qkv[:, 2 * W:], o, dout, lse, delta, dq_acc, dqkv[:, :W],
                 dqkv[:, W:2 * W], dqkv[:, 2 * W:], B, H, L, HD, sc, True)
    dv = dqkv[:, 2 * W:].float().view(B, L, H, HD)
    colsum = dv.sum(1)  # (B, H, HD): total probability mass = L per (b, h, d)
    torch.testing.assert_close(colsum, torch.full_like(colsum, float(L)), rtol=2e-2, atol=0)


def test_rmsnorm_scale_invariance_and_rope_isometry_full_size(ops):
    torch.manual_seed(2)
    x, w = rnd(M, D), (1 + 0.1 * torch.randn(D, device=dev())).to(bf16)
    y1 = ops.rmsnorm_fwd(x, w, 0.0)
    y2 = ops.rmsnorm_fwd((x.float() * 8).to(bf16), w, 0.0)  # eps = 0: exactly scale invariant
    assert torch.equal(y1, y2)
    table = ops.rope_table(HD, 2 * 1024, device=dev())
    q = rnd(M, D)
    r = q.clone()
    ops.rope_apply(r, H, HD, table, L, 0)
    n0, n1 = q.float().view(M, H, HD // 2, 2).norm(dim=-1), r.float().view(M, H, HD // 2, 2).norm(dim=-1)
    torch.testing.assert_close(n1, n0, rtol=2 ** -6, atol=2e-2)  # rotations preserve each pair's length
    assert torch.equal(r.view(B, L, D)[:, 0], q.view(B, L, D)[:, 0])  # position 0 is the identity
    ops.rope_apply(r, H, HD, table, L, 0, inverse=True)
    torch.testing.assert_close(r.float(), q.float(), rtol=2 ** -6, atol=3e-2)
    # the GEMM-fused rotation equals GEMM then stand-alone rotation
    wq = rnd(2 * D, D, scale=0.02)
    fused = ops.gemm(x, wq, rope=table, rope_cols=2 * D, rope_period=L)
    plain = ops.gemm(x, wq)
    ops.rope_apply(plain, 2 * H, HD, table, L, 0)
    assert (fused != plain).float().mean().item() < 1e-3  # identical arithmetic; rare 1-ulp flips from a different tile shape


def test_gate_fusion_is_composition_full_size(ops):
    torch.manual_seed(3)
    x, w13 = rnd(M, D), rnd(2 * F, D, scale=0.02)
    gu, y = torch.empty(M, 2 * F, device=dev(), dtype=bf16), torch.empty(M, F, device=dev(), dtype=bf16)
    ops.gemm_swiglu(x, w13, gu, y)
    assert torch.equal(y, ops.swiglu_fwd(gu))
    ref = ops.gemm(x, w13)
    assert (gu != ref).float().mean().item() < 2e-3


def test_adamw_fixed_point_and_sumsq_full_size(ops):
    n = 202_375_168  # one 7B decoder layer's parameters
    master = torch.randn(n, device=dev())
    keep = master.clone()
    g = torch.zeros(n, device=dev(), dtype=bf16)
    m, v, wb = torch.zeros(n, device=dev()), torch.zeros(n, device=dev()), torch.empty(n, device=dev(), dtype=bf16)
    nsq = torch.zeros(1, device=dev())
    ops.sumsq_bf16(g, nsq)
    assert nsq.item() == 0.0
    ops.adamw(master, wb, g, m, v, 1e-3, 0.9, 0.95, 1e-8, 0.0, 1, nsq, 8.0)  # zero grad, no decay: nothing moves
    assert torch.equal(master, keep) and torch.equal(wb, keep.to(bf16)) and not m.any() and not v.any()
    g.fill_(0.5)
    nsq.zero_()
    ops.sumsq_bf16(g, nsq)
    assert abs(nsq.item() - 0.25 * n) <= 1e-5 * 0.25 * n
    ops.adamw(master, wb, g, m, v, 1e-3, 0.9, 0.95, 1e-8, 0.0, 1, None, 0.0)
    # step 1 with constant gradient: update = -lr * g / (|g| + eps) = -lr for every element
    torch.testing.assert_close(master, keep - 1e-3, rtol=0, atol=2e-7)


def test_decode_attention_equals_flash_prefill_full_size(ops):
    torch.manual_seed(4)
    Bq, ctx, max_seq = 4, 900, 1024
    W = H * HD
    kc, vc = rnd(Bq, max_seq, H, HD), rnd(Bq, max_seq, H, HD)
    q = rnd(Bq, W)
    out = torch.empty(Bq, W, device=dev(), dtype=bf16)
    scratch = torch.zeros(Bq * H * (32 * 130 + 1), device=dev())
    sc = 1 / math.sqrt(HD)
    ops.attn_decode(q, kc, vc, out, scratch, Bq, H, HD, ctx, max_seq, sc)
    ref = ops.attn_fwd(q, kc.view(-1, W), vc.view(-1, W), Bq, H, 1, ctx, HD, sc, True, kv_bstride=max_seq * W)
    torch.testing.assert_close(out.float(), ref.float(), rtol=2 ** -6, atol=6e-3)
