"""Kernel-level parity on a B200: every C-ABI entry point against a plain PyTorch fp32 restatement of
the same op on the same bf16-rounded inputs.  Tolerances: fp32 outputs rtol 1e-3 / atol 1e-4 scaled by
the reduction length where sums are involved; bf16 outputs within one bf16 ulp of the fp32 result
(rtol 2^-7) -- the north star's 1e-3 is tighter than bf16 resolution."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

bf16 = torch.bfloat16


@pytest.fixture(scope="module")
def ops():
    from a3vlm_b200 import ops as o

    o._L()
    return o


def dev():
    return torch.device("cuda:0")


def rnd(*shape, scale=1.0, seed=None):
    if seed is not None:
        torch.manual_seed(seed)
    return (torch.randn(*shape, device=dev()) * scale).to(bf16)


def close_bf16(got, ref, rtol=2 ** -7, atol=1e-3):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    bad = err > tol
    assert not bad.any(), f"max err {err.max().item():.4g} at ref {ref[bad][:4].tolist()} got {got[bad][:4].tolist()} ({bad.sum().item()} bad of {bad.numel()})"


# ------------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (200, 264, 200), (835, 1024, 592), (1091, 512, 1024)])
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_gemm_fp32_out(ops, M, N, K, a_mn, b_mn):
    torch.manual_seed(M + N + K)
    if a_mn:
        M = (M + 7) // 8 * 8  # MN-major A is stored [K][M]: its row stride M must be 16-byte aligned
    a, b = rnd(M, K), rnd(N, K)
    ref = a.float() @ b.float().t()
    A = a.t().contiguous() if a_mn else a
    Bm = b.t().contiguous() if b_mn else b
    c = ops.gemm(A, Bm, a_mn=a_mn, b_mn=b_mn, out_fp32=True)
    torch.testing.assert_close(c, ref, rtol=1e-3, atol=1e-4 * math.sqrt(K))


@pytest.mark.parametrize("M", [1, 9, 16, 64])
def test_gemm_small_m_mode(ops, M):
    """Decode-time shapes: one short M block, narrow N tiles, only the live rows of A staged."""
    torch.manual_seed(M)
    N, K = 1504, 1024
    a, w, res = rnd(M, K), rnd(N, K, scale=0.05), rnd(M, N)
    ref = a.float() @ w.float().t()
    c = ops.gemm(a, w, out_fp32=True)
    torch.testing.assert_close(c, ref, rtol=1e-3, atol=1e-4 * math.sqrt(K))
    c = ops.gemm(a, w, residual=res)
    close_bf16(c, ref.to(bf16).float() + res.float(), atol=2e-2)


@pytest.mark.parametrize("M,N,K", [(1, 1000, 512), (5, 12288, 4096), (16, 4096, 4096), (17, 4096, 11008), (40, 512, 256), (64, 32000, 4096)])
def test_decode_gemm_swapped_operands(ops, M, N, K):
    """tcgen05 decode projection (weights as the 128-row operand, split-K over a cluster + DSMEM reduce)."""
    torch.manual_seed(M + N)
    x, w, res = rnd(M, K), rnd(N, K, scale=0.03), rnd(M, N)
    ref = x.float() @ w.float().t()
    y = torch.full((M, N), float("nan"), device=dev())
    ops.decode_gemm(x, w, y, out_fp32=True)
    close_bf16(y, ref, atol=4e-3)
    assert torch.equal(y, y.to(bf16).float())
    y2 = torch.full((M, N), float("nan"), device=dev(), dtype=bf16)
    ops.decode_gemm(x, w, y2, residual=res)
    close_bf16(y2, ref.to(bf16).float() + res.float(), atol=3e-2)


def test_gemm_epilogue_bias_gelu_residual(ops):
    torch.manual_seed(0)
    M, N, K = 577, 1024, 512
    a, w, bias, res = rnd(M, K), rnd(N, K, scale=0.05), rnd(N), rnd(M, N)
    lin = (a.float() @ w.float().t() + bias.float()).to(bf16)
    for act, fn in [(ops.ACT_GELU, lambda t: torch.nn.functional.gelu(t.float())),
                    (ops.ACT_QUICKGELU, lambda t: t.float() * torch.sigmoid(1.702 * t.float()))]:
        c = ops.gemm(a, w, bias=bias, act=act)
        close_bf16(c, fn(lin), rtol=2 ** -6, atol=2e-3)  # a 1-ulp flip of the pre-activation moves the output 1 ulp
    c = ops.gemm(a, w, bias=bias, residual=res)
    close_bf16(c, lin.float() + res.float(), atol=2e-2)
    # accumulate (fp32 and bf16)
    c0 = torch.randn(M, N, device=dev())
    c = ops.gemm(a, w, out=c0.clone(), out_fp32=True, accumulate=True, alpha=0.5)
    torch.testing.assert_close(c, c0 + 0.5 * (a.float() @ w.float().t()), rtol=1e-3, atol=2e-3)


def test_gemm_rope_epilogue(ops):
    torch.manual_seed(1)
    B, L, D, H = 2, 75, 256, 2
    x, w = rnd(B * L, D), rnd(3 * D, D, scale=0.06)
    table = ops.rope_table(128, 256, device=dev())
    qkv = ops.gemm(x, w, rope=table, rope_cols=2 * D, rope_period=L, rope_pos0=3)
    lin = (x.float() @ w.float().t()).to(bf16)
    ref = lin.clone().float()
    qk = lin[:, : 2 * D].float().reshape(B, L, 2 * H, 64, 2)
    cs = table[3 : 3 + L].reshape(1, L, 1, 64, 2)
    re = qk[..., 0] * cs[..., 0] - qk[..., 1] * cs[..., 1]
    im = qk[..., 0] * cs[..., 1] + qk[..., 1] * cs[..., 0]
    ref[:, : 2 * D] = torch.stack([re, im], -1).reshape(B * L, 2 * D)
    close_bf16(qkv, ref, atol=4e-3)
    # strided standalone rope (inverse undoes forward up to bf16 rounding)
    y = lin.clone()
    ops.rope_apply(y[:, : 2 * D], 2 * H, 128, table, L, 3)
    close_bf16(y, ref, atol=4e-3)
    ops.rope_apply(y[:, : 2 * D], 2 * H, 128, table, L, 3, inverse=True)
    close_bf16(y, lin, rtol=2 ** -6, atol=2e-2)


@pytest.mark.parametrize("M", [1, 3, 8, 16])
def test_linear_small_m(ops, M):
    torch.manual_seed(M)
    N, K = 1000, 512
    x, w, res = rnd(M, K), rnd(N, K, scale=0.05), rnd(M, N)
    ref = x.float() @ w.float().t()
    y = ops.linear_small_m(x, w, out_fp32=True)
    close_bf16(y, ref, atol=2e-3)
    assert torch.equal(y, y.to(bf16).float())  # logits are a bf16 result cast up (llama_ens5.py:530-531)
    y = ops.linear_small_m(x, w, residual=res)
    close_bf16(y, ref.to(bf16).float() + res.float(), atol=2e-2)


# ------------------------------------------------------------------------------------------- norms
def rms_ref(x, w, eps):
    xf = x.float()
    n = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)).to(bf16)
    return n * w


@pytest.mark.parametrize("rows,D", [(5, 256), (835, 4096), (64, 5120)])
def test_rmsnorm_fwd_exact(ops, rows, D):
    x, w = rnd(rows, D, seed=rows), (1 + 0.1 * torch.randn(D, device=dev())).to(bf16)
    rstd = torch.empty(rows, device=dev())
    y = ops.rmsnorm_fwd(x, w, 1e-5, rstd=rstd)
    ref = rms_ref(x, w, 1e-5)
    # same rounding points as components.py:52-53 -> equal up to rare 1-ulp flips from the reduction order
    mism = (y != ref).float().mean().item()
    assert mism < 2e-3, mism
    close_bf16(y, ref)
    torch.testing.assert_close(rstd, torch.rsqrt(x.float().pow(2).mean(-1) + 1e-5), rtol=1e-5, atol=1e-6)


def test_rmsnorm_bwd(ops):
    rows, D = 300, 1024
    x, w, dy, prev = rnd(rows, D, seed=3), (1 + 0.1 * torch.randn(D, device=dev())).to(bf16), rnd(rows, D), rnd(rows, D)
    rstd = torch.empty(rows, device=dev())
    ops.rmsnorm_fwd(x, w, 1e-5, rstd=rstd)
    xf = x.float().requires_grad_(True)
    wf = w.float().requires_grad_(True)
    yf = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5) * wf
    yf.backward(dy.float())
    parts = torch.empty(148, D, device=dev())
    dx = prev.clone()
    ops.rmsnorm_bwd(dy, x, w, rstd, dx, parts, add_to_dx=True)
    close_bf16(dx, xf.grad.to(bf16).float() + prev.float(), rtol=2 ** -6, atol=2e-2)
    dw = torch.empty(D, device=dev())
    ops.colsum_partials(parts, dw)
    torch.testing.assert_close(dw, wf.grad, rtol=2e-2, atol=0.15)  # dw uses the bf16-rounded normalised x
    dx2 = torch.empty_like(x)
    ops.rmsnorm_bwd(dy, x, w, rstd, dx2, parts)
    close_bf16(dx2, xf.grad, atol=2e-3)


@pytest.mark.parametrize("rows,D", [(7, 128), (577, 1024), (100, 4096)])
def test_layernorm_fwd_bwd(ops, rows, D):
    x, dy = rnd(rows, D, seed=rows, scale=2.0), rnd(rows, D)
    w, b = (1 + 0.1 * torch.randn(D, device=dev())).to(bf16), (0.1 * torch.randn(D, device=dev())).to(bf16)
    mean, rstd = torch.empty(rows, device=dev()), torch.empty(rows, device=dev())
    y = ops.layernorm_fwd(x, w, b, 1e-5, mean=mean, rstd=rstd)
    xf, wf, bf_ = x.float().requires_grad_(True), w.float().requires_grad_(True), b.float().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xf, (D,), wf, bf_, 1e-5)
    close_bf16(y, ref, atol=2e-3)
    ref.backward(dy.float())
    dx = torch.empty_like(x)
    pw, pb = torch.empty(64, D, device=dev()), torch.empty(64, D, device=dev())
    ops.layernorm_bwd(dy, x, w, mean, rstd, dx, pw, pb)
    close_bf16(dx, xf.grad, atol=3e-3)
    dw, db = torch.empty(D, device=dev()), torch.empty(D, device=dev())
    ops.colsum_partials(pw, dw)
    ops.colsum_partials(pb, db)
    torch.testing.assert_close(dw, wf.grad, rtol=1e-3, atol=1e-2)
    torch.testing.assert_close(db, bf_.grad, rtol=1e-3, atol=1e-2)
    s = torch.empty(D, device=dev(), dtype=bf16)
    ops.colsum_bf16(dy, s)
    close_bf16(s, dy.float().sum(0), atol=2e-2)


# ------------------------------------------------------------------------------------------- swiglu
def test_swiglu(ops):
    rows, F = 130, 1376
    gu, dy = rnd(rows, 2 * F, seed=5), rnd(rows, F)
    y = ops.swiglu_fwd(gu)
    g, u = gu[:, :F], gu[:, F:]
    ref = torch.nn.functional.silu(g.float()).to(bf16) * u
    close_bf16(y, ref)
    gf, uf = g.float().requires_grad_(True), u.float().requires_grad_(True)
    (torch.nn.functional.silu(gf) * uf).backward(dy.float())
    dgu = torch.empty_like(gu)
    ops.swiglu_bwd(gu, dy, dgu)
    close_bf16(dgu[:, :F], gf.grad, rtol=2 ** -6, atol=2e-2)
    close_bf16(dgu[:, F:], uf.grad, rtol=2 ** -6, atol=2e-2)


@pytest.mark.parametrize("M,F", [(200, 640), (835, 640), (1000, 1160)])
def test_gemm_fused_swiglu_fwd_bwd(ops, M, F):
    """The gate fused into the tcgen05 epilogues == the unfused kernel sequences it replaces (single-CTA kernel at
    M 200, the CTA-pair kernel with its coalesced staged gate-backward epilogue otherwise; F 1160 = 128 * 9 + 8 leaves a
    partial 16-byte piece pattern in the last column tile of the backward)."""
    torch.manual_seed(M)
    K = 256
    x, w13, w2, d = rnd(M, K), rnd(2 * F, K, scale=0.06), rnd(K, F, scale=0.05), rnd(M, K)
    gu_ref = ops.gemm(x, w13)
    y_ref = ops.swiglu_fwd(gu_ref)
    if F % 128 == 0:  # the fused forward needs whole 128-feature tiles
        gu, y = torch.empty_like(gu_ref), torch.empty_like(y_ref)
        ops.gemm_swiglu(x, w13, gu, y)
        close_bf16(gu, gu_ref, atol=2e-3)  # same MMAs, different tile shape -> same values up to accumulation order
        close_bf16(y, ops.swiglu_fwd(gu), rtol=0, atol=0)  # y is exactly the gate of the gu it wrote
    else:
        gu = gu_ref
    dy = ops.gemm(d, w2, b_mn=True)
    dgu_ref = torch.empty_like(gu)
    ops.swiglu_bwd(gu, dy, dgu_ref)
    dgu = torch.empty_like(gu)
    ops.gemm_dswiglu(d, w2, gu, dgu)
    close_bf16(dgu, dgu_ref, rtol=2 ** -6, atol=4e-3)


# ------------------------------------------------------------------------------------------- attention
def attn_ref(q, k, v, causal, scale):
    # q [B,Tq,H,d], k/v [B,Tk,H,d] fp32; right-aligned causal (llama_ens5.py:181-185)
    B, Tq, H, d = q.shape
    Tk = k.shape[1]
    s = torch.einsum("bqhd,bkhd->bhqk", q, k) * scale
    if causal:
        qi = torch.arange(Tq, device=q.device)[:, None] + (Tk - Tq)
        ki = torch.arange(Tk, device=q.device)[None, :]
        s = s.masked_fill(ki > qi, float("-inf"))
    lse = torch.logsumexp(s, -1)
    p = torch.softmax(s, -1)
    return torch.einsum("bhqk,bkhd->bqhd", p, v), lse


@pytest.mark.parametrize("B,H,Tq,Tk,hd,causal", [
    (2, 2, 128, 128, 128, 1), (1, 3, 200, 200, 128, 1), (2, 2, 835, 835, 128, 1),
    (1, 2, 70, 300, 128, 1), (2, 4, 257, 257, 64, 0), (1, 16, 577, 577, 64, 0), (1, 2, 5, 5, 128, 1),
    (2, 2, 256, 256, 128, 1), (1, 1, 129, 129, 128, 1), (1, 2, 1603, 1603, 128, 1), (1, 3, 300, 300, 64, 0),
    (2, 2, 384, 384, 128, 0), (1, 2, 3, 700, 128, 1), (1, 2, 300, 1000, 128, 1), (3, 2, 64, 64, 64, 1)])
def test_attn_fwd(ops, B, H, Tq, Tk, hd, causal):
    torch.manual_seed(Tq + hd)
    W = H * hd
    qkv = rnd(B * Tq, 3 * W) if Tq == Tk else None
    if qkv is not None:
        q, k, v = qkv[:, :W], qkv[:, W:2 * W], qkv[:, 2 * W:]
    else:
        q, k, v = rnd(B * Tq, W), rnd(B * Tk, W), rnd(B * Tk, W)
    lse = torch.empty(B, H, Tq, device=dev())
    scale = 1 / math.sqrt(hd)
    o = ops.attn_fwd(q, k, v, B, H, Tq, Tk, hd, scale, causal, lse=lse)
    ref, lse_ref = attn_ref(q.float().reshape(B, Tq, H, hd), k.float().reshape(B, Tk, H, hd),
                            v.float().reshape(B, Tk, H, hd), causal, scale)
    close_bf16(o.reshape(B, Tq, H, hd), ref, rtol=2 ** -6, atol=6e-3)  # P is rounded to bf16 before PV
    torch.testing.assert_close(lse, lse_ref, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("mag", [4.0, 12.0])
def test_attn_fwd_growing_maxima(ops, mag):
    """Scores with a large spread whose row maxima keep growing along the keys (later keys are scaled up), so the
    lazy-rescale path of the warp-specialised kernel (running maximum moves only past 2^8, O rescaled by the
    correction warps) is exercised tile after tile."""
    torch.manual_seed(int(mag))
    B, H, T, hd = 2, 2, 700, 128
    W = H * hd
    q, v = rnd(B * T, W), rnd(B * T, W)
    ramp = torch.linspace(0.2, mag, T, device=dev()).repeat(B)[:, None]
    k = (rnd(B * T, W).float() * ramp).to(torch.bfloat16)
    lse = torch.empty(B, H, T, device=dev())
    for causal in (1, 0):
        o = ops.attn_fwd(q, k, v, B, H, T, T, hd, 1 / math.sqrt(hd), causal, lse=lse)
        ref, lse_ref = attn_ref(q.float().reshape(B, T, H, hd), k.float().reshape(B, T, H, hd),
                                v.float().reshape(B, T, H, hd), causal, 1 / math.sqrt(hd))
        close_bf16(o.reshape(B, T, H, hd), ref, rtol=2 ** -6, atol=6e-3)
        torch.testing.assert_close(lse, lse_ref, rtol=1e-4, atol=2e-4)


@pytest.mark.parametrize("B,H,T,causal", [(1, 2, 128, 1), (2, 2, 200, 1), (1, 2, 835, 1), (1, 1, 64, 0)])
def test_attn_bwd(ops, B, H, T, causal):
    torch.manual_seed(T)
    hd, W = 128, H * 128
    qkv, dout = rnd(B * T, 3 * W), rnd(B * T, W)
    q, k, v = qkv[:, :W], qkv[:, W:2 * W], qkv[:, 2 * W:]
    scale = 1 / math.sqrt(hd)
    lse = torch.empty(B, H, T, device=dev())
    o = ops.attn_fwd(q, k, v, B, H, T, T, hd, scale, causal, lse=lse)
    qf, kf, vf = (t.float().reshape(B, T, H, hd).requires_grad_(True) for t in (q, k, v))
    ref, _ = attn_ref(qf, kf, vf, causal, scale)
    ref.backward(dout.float().reshape(B, T, H, hd))
    dqkv = torch.full_like(qkv, float("nan"))
    delta = torch.empty(B, H, T, device=dev())
    dq_acc = torch.empty(B * T, W, device=dev())
    ops.attn_bwd(q, k, v, o, dout, lse, delta, dq_acc, dqkv[:, :W], dqkv[:, W:2 * W], dqkv[:, 2 * W:], B, H, T,
                 hd, scale, causal)
    def cmp(dqkv, refs):
        for name, got, g in zip(("dq", "dk", "dv"), (dqkv[:, :W], dqkv[:, W:2 * W], dqkv[:, 2 * W:]), refs):
            got = got.float().reshape(B, T, H, hd)
            err = (got - g).abs().max().item()
            assert err <= 2e-2 * max(1.0, g.abs().max().item()), (name, err, g.abs().max().item())
            rel = ((got - g).norm() / g.norm()).item()
            assert rel < 1e-2, (name, rel)

    cmp(dqkv, (qf.grad, kf.grad, vf.grad))
    # fused backward of apply_rotary_emb: dq, dk rotated by -angle(position) == a3_rope_apply(inverse)
    table = ops.rope_table(128, 1024, device=dev())
    d2 = torch.full_like(qkv, float("nan"))
    ops.attn_bwd(q, k, v, o, dout, lse, delta, dq_acc, d2[:, :W], d2[:, W:2 * W], d2[:, 2 * W:], B, H, T, hd, scale,
                 causal, rope=table)
    ref = dqkv.clone()
    ops.rope_apply(ref[:, :2 * W], 2 * H, hd, table, T, 0, inverse=True)
    close_bf16(d2, ref, rtol=2 ** -6, atol=2e-2 * max(1.0, ref.float().abs().max().item()) * 0.25)


@pytest.mark.parametrize("B,H,ctx", [(1, 32, 836), (4, 4, 17), (2, 2, 1)])
def test_attn_decode_and_kv_append(ops, B, H, ctx):
    torch.manual_seed(ctx)
    hd, max_seq, W = 128, 1024, H * 128
    kc, vc = torch.zeros(B, max_seq, H, hd, device=dev(), dtype=bf16), torch.zeros(B, max_seq, H, hd, device=dev(), dtype=bf16)
    kv = rnd(B * ctx, 3 * W)
    ops.kv_append(kv[:, W:2 * W], kv[:, 2 * W:], kc, vc, B, ctx, H, hd, 0, max_seq)
    assert torch.equal(kc[:, :ctx].reshape(B * ctx, W), kv[:, W:2 * W])
    assert torch.equal(vc[:, :ctx].reshape(B * ctx, W), kv[:, 2 * W:])
    q = rnd(B, W)
    out = torch.empty(B, W, device=dev(), dtype=bf16)
    scratch = torch.zeros(B * H * (32 * 130 + 1), device=dev())
    scale = 1 / math.sqrt(hd)
    ops.attn_decode(q, kc, vc, out, scratch, B, H, hd, ctx, max_seq, scale)
    ref, _ = attn_ref(q.float().reshape(B, 1, H, hd), kc[:, :ctx].float(), vc[:, :ctx].float(), 0, scale)
    close_bf16(out.reshape(B, 1, H, hd), ref, atol=3e-3)
    # the flash kernel with Tq=1 (prefill-with-cache shape) agrees as well
    o2 = ops.attn_fwd(q, kc.reshape(B * max_seq, W), vc.reshape(B * max_seq, W), B, H, 1, ctx, hd, scale, 1,
                      kv_bstride=max_seq * W)
    close_bf16(o2.reshape(B, 1, H, hd), ref, rtol=2 ** -6, atol=6e-3)


def test_decode_rope_kv_and_device_pos(ops):
    torch.manual_seed(12)
    B, H, hd, max_seq, pos = 3, 4, 128, 64, 17
    W = H * hd
    qkv = rnd(B, 3 * W)
    table = ops.rope_table(128, 128, device=dev())
    kc, vc = torch.zeros(B, max_seq, H, hd, device=dev(), dtype=bf16), torch.zeros(B, max_seq, H, hd, device=dev(), dtype=bf16)
    ref = qkv.clone()
    ops.rope_apply(ref[:, :2 * W], 2 * H, hd, table, 1, pos)
    got = qkv.clone()
    pos_dev = torch.tensor([pos], device=dev(), dtype=torch.int32)
    ops.decode_rope_kv(got, kc, vc, B, H, hd, table, 0, max_seq, pos_dev=pos_dev)
    assert torch.equal(got, ref)
    assert torch.equal(kc[:, pos].reshape(B, W), ref[:, W:2 * W]) and torch.equal(vc[:, pos].reshape(B, W), ref[:, 2 * W:])
    assert (kc[:, :pos] == 0).all() and (kc[:, pos + 1:] == 0).all()
    got2 = qkv.clone()
    ops.decode_rope_kv(got2, kc, vc, B, H, hd, table, pos, max_seq)  # host-side position gives the same
    assert torch.equal(got2, ref)
    # attention with the context length read from device memory == host-side ctx
    kc.copy_(rnd(B, max_seq, H, hd)); vc.copy_(rnd(B, max_seq, H, hd))
    o1, o2 = torch.empty(B, W, device=dev(), dtype=bf16), torch.empty(B, W, device=dev(), dtype=bf16)
    scratch = torch.zeros(B * H * (32 * 130 + 1), device=dev())
    ops.attn_decode(got[:, :W], kc, vc, o1, scratch, B, H, hd, pos + 1, max_seq, 0.088)
    ops.attn_decode(got[:, :W], kc, vc, o2, scratch, B, H, hd, max_seq, max_seq, 0.088, pos_dev=pos_dev)
    close_bf16(o2, o1, atol=2e-3)


@pytest.mark.parametrize("M", [1, 3, 8])
def test_decode_linear_fused(ops, M):
    """One-launch decode projections: RMSNorm prologue, SwiGLU prologue, residual / fp32 epilogues and the
    RoPE + KV-append epilogue, against the unfused kernels they replace."""
    torch.manual_seed(40 + M)
    H, hd, K, F, max_seq, pos = 2, 128, 256, 704, 32, 9
    D = H * hd
    x, nw = rnd(M, K), (1 + 0.1 * torch.randn(K, device=dev())).to(bf16)
    wqkv, w2, res = rnd(3 * D, K, scale=0.06), rnd(K, F, scale=0.05), rnd(M, K)
    table = ops.rope_table(128, 64, device=dev())
    pos_dev = torch.tensor([pos], device=dev(), dtype=torch.int32)
    # norm prologue + rope/append epilogue == rmsnorm -> linear -> decode_rope_kv
    kc, vc = (torch.zeros(M, max_seq, H, hd, device=dev(), dtype=bf16) for _ in range(2))
    kc2, vc2 = torch.zeros_like(kc), torch.zeros_like(vc)
    ref = ops.linear_small_m(ops.rmsnorm_fwd(x, nw, 1e-5), wqkv)
    ops.decode_rope_kv(ref, kc2, vc2, M, H, hd, table, pos, max_seq)
    got = torch.empty_like(ref)
    ops.decode_linear(x, wqkv, got, norm_w=nw, eps=1e-5, rope=table, n_heads=H, kcache=kc, vcache=vc,
                      pos_dev=pos_dev, max_seq=max_seq)
    close_bf16(got, ref, atol=4e-3)
    close_bf16(kc, kc2, atol=4e-3)
    close_bf16(vc, vc2, atol=4e-3)
    assert torch.equal(kc[:, pos].reshape(M, D), got[:, D:2 * D]) and torch.equal(vc[:, pos].reshape(M, D), got[:, 2 * D:])
    # swiglu prologue + residual == swiglu -> linear(+residual)
    gu = rnd(M, 2 * F)
    ref = ops.linear_small_m(ops.swiglu_fwd(gu), w2, residual=res)
    got = torch.empty_like(ref)
    ops.decode_linear(gu, w2, got, K=F, swiglu=True, residual=res)
    close_bf16(got, ref, atol=4e-3)
    # gate as the epilogue of the w1|w3 projection == rmsnorm -> linear -> swiglu, bit for bit (same dot-product order)
    w13 = rnd(2 * F, K, scale=0.06)
    ref = ops.swiglu_fwd(ops.linear_small_m(ops.rmsnorm_fwd(x, nw, 1e-5), w13))
    ref2 = torch.empty(M, 2 * F, device=dev(), dtype=bf16)
    ops.decode_linear(x, w13, ref2, norm_w=nw, eps=1e-5)
    got = torch.empty(M, F, device=dev(), dtype=bf16)
    ops.decode_linear_gated(x, w13, got, norm_w=nw, eps=1e-5)
    assert torch.equal(got, ops.swiglu_fwd(ref2))
    close_bf16(got, ref, atol=4e-3)
    # norm prologue + fp32 logits
    wo = rnd(1000, K, scale=0.05)
    ref = ops.linear_small_m(ops.rmsnorm_fwd(x, nw, 1e-5), wo, out_fp32=True)
    got = torch.empty_like(ref)
    ops.decode_linear(x, wo, got, norm_w=nw, eps=1e-5, out_fp32=True)
    close_bf16(got, ref, atol=2e-3)
    assert torch.equal(got, got.to(bf16).float())


# ------------------------------------------------------------------------------------------- embed / ViT front
def test_embed_splice_and_bwd(ops):
    torch.manual_seed(7)
    B, T, Nimg, D, V = 2, 9, 5, 64, 50
    E, img, st, en = rnd(V, D), rnd(B, Nimg, D), rnd(D), rnd(D)
    tok = torch.randint(1, V, (B, T), device=dev())
    W = Nimg + 2
    h = torch.empty(B, T + W, D, device=dev(), dtype=bf16)
    ops.embed_splice(tok, E, img, st, en, h, B, T, Nimg, D)
    emb = E[tok]
    ref = torch.cat([emb[:, :1], st.expand(B, 1, D), img, en.expand(B, 1, D), emb[:, 1:]], 1)
    assert torch.equal(h, ref)
    h2 = torch.empty(B, T, D, device=dev(), dtype=bf16)
    ops.embed_splice(tok, E, None, None, None, h2, B, T, 0, D)
    assert torch.equal(h2, emb)
    dh = rnd(B, T + W, D)
    dE, dimg = torch.zeros(V, D, device=dev(), dtype=bf16), torch.empty(B, Nimg, D, device=dev(), dtype=bf16)
    ds, de = torch.empty(D, device=dev()), torch.empty(D, device=dev())
    ops.embed_splice_bwd(tok, dh, dE, dimg, ds, de, B, T, Nimg, D)
    assert torch.equal(dimg, dh[:, 2:W])
    torch.testing.assert_close(ds, dh[:, 1].float().sum(0))
    torch.testing.assert_close(de, dh[:, W].float().sum(0))
    refE = torch.zeros(V, D, device=dev())
    text = torch.cat([dh[:, :1], dh[:, W + 1:]], 1).float().reshape(B * T, D)
    refE.index_add_(0, tok.reshape(-1), text)
    close_bf16(dE, refE, rtol=2 ** -6, atol=3e-2)


def test_build_views_and_multiview_splice(ops):
    """The reference's 5-view tiling (llama_ens5.py:383-385) and the per-view start/end wrapping (:471-477)."""
    torch.manual_seed(14)
    B, V = 2, 28
    img = torch.randn(B, 3, 2 * V, 2 * V, device=dev())
    views = torch.empty(5 * B, 3, V, V, device=dev())
    ops.build_views(img, views, B, V)
    g = torch.nn.functional.interpolate(img.half(), size=(V, V), mode="bicubic").to(img)
    ref = torch.cat([g, img[..., :V, :V], img[..., :V, V:], img[..., V:, :V], img[..., V:, V:]], 0)
    assert torch.equal(views[B:], ref[B:])
    assert (views[:B] - ref[:B]).abs().max().item() <= 2 ** -10 * ref[:B].abs().max().item()  # <= 1 fp16 ulp
    vb = torch.empty_like(views)
    ops.build_views(img.to(bf16), vb, B, V)
    assert torch.equal(vb[B:], ref[B:].to(bf16).float())
    # multi-view splice: BOS | (start, N tokens, end) x 5 | text, image tokens view-major
    T, N, D, Vc = 6, 3, 64, 40
    E, toks, st, en = rnd(Vc, D), rnd(5 * B, N, D), rnd(D), rnd(D)
    tok = torch.randint(1, Vc, (B, T), device=dev())
    W = 5 * (N + 2)
    h = torch.empty(B, T + W, D, device=dev(), dtype=bf16)
    ops.embed_splice(tok, E, toks, st, en, h, B, T, N, D, n_views=5)
    emb = E[tok]
    chunks = [torch.cat([st.expand(B, 1, D), toks[v * B:(v + 1) * B], en.expand(B, 1, D)], 1) for v in range(5)]
    ref = torch.cat([emb[:, :1]] + chunks + [emb[:, 1:]], 1)
    assert torch.equal(h, ref)
    dh = rnd(B, T + W, D)
    dE, dimg = torch.zeros(Vc, D, device=dev(), dtype=bf16), torch.empty(5 * B, N, D, device=dev(), dtype=bf16)
    ds, de = torch.empty(D, device=dev()), torch.empty(D, device=dev())
    ops.embed_splice_bwd(tok, dh, dE, dimg, ds, de, B, T, N, D, n_views=5)
    for v in range(5):
        base = 1 + v * (N + 2)
        assert torch.equal(dimg[v * B:(v + 1) * B], dh[:, base + 1:base + 1 + N])
    starts = torch.stack([dh[:, 1 + v * (N + 2)] for v in range(5)], 0).float().sum((0, 1))
    ends = torch.stack([dh[:, (v + 1) * (N + 2)] for v in range(5)], 0).float().sum((0, 1))
    torch.testing.assert_close(ds, starts)
    torch.testing.assert_close(de, ends)


def test_vit_front(ops):
    torch.manual_seed(8)
    B, S, P, D = 2, 56, 14, 128
    g, K, Kpad = S // P, 3 * P * P, 592
    img = torch.randn(B, 3, S, S, device=dev())
    A = torch.empty(B * g * g, Kpad, device=dev(), dtype=bf16)
    ops.vit_patchify(img, A, B, S, P, Kpad)
    ref = torch.nn.functional.unfold(img, P, stride=P).transpose(1, 2).reshape(B * g * g, K).to(bf16)
    assert torch.equal(A[:, :K], ref) and (A[:, K:] == 0).all()
    A2 = torch.empty_like(A)
    ops.vit_patchify(img.to(bf16), A2, B, S, P, Kpad)
    assert torch.equal(A, A2)
    po, cls, pos = rnd(B * g * g, D), rnd(D), rnd(g * g + 1, D)
    lw, lb = (1 + 0.1 * torch.randn(D, device=dev())).to(bf16), (0.1 * torch.randn(D, device=dev())).to(bf16)
    x = torch.empty(B, g * g + 1, D, device=dev(), dtype=bf16)
    ops.vit_assemble_lnpre(po, cls, pos, lw, lb, x, B, g * g, D, 1e-5)
    t = torch.cat([cls.expand(B, 1, D), po.reshape(B, g * g, D)], 1)
    t = (t.float() + pos.float()).to(bf16)
    ref = torch.nn.functional.layer_norm(t.float(), (D,), lw.float(), lb.float(), 1e-5)
    close_bf16(x, ref, atol=2e-3)


# ------------------------------------------------------------------------------------------- loss / argmax / optimiser
def test_cross_entropy_and_argmax(ops):
    torch.manual_seed(9)
    rows, V = 37, 32000
    logits = rnd(rows, V, scale=2.0)
    labels = torch.randint(1, V, (rows,), device=dev())
    labels[::5] = 0
    count = torch.zeros(1, device=dev(), dtype=torch.int32)
    ops.count_nonzero(labels, count)
    assert count.item() == (labels != 0).sum().item()
    lf = logits.float().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(lf, labels, ignore_index=0)
    ref.backward()
    row_loss, out = torch.empty(rows, device=dev()), torch.empty(1, device=dev())
    work = logits.clone()
    ops.cross_entropy(work, labels, row_loss, count, 1.0, True)
    ops.loss_reduce(row_loss, count, out)
    torch.testing.assert_close(out[0], ref.detach(), rtol=1e-4, atol=1e-4)
    close_bf16(work, lf.grad, rtol=2 ** -7, atol=1e-6)
    fl = torch.randn(5, V, device=dev())
    fl[0, 100] = fl[0, 30000] = 50.0  # tie -> lowest index
    am = torch.empty(5, device=dev(), dtype=torch.int64)
    ops.argmax(fl, am)
    assert torch.equal(am, fl.argmax(-1)) and am[0].item() == 100


def test_adamw_and_sumsq(ops):
    torch.manual_seed(10)
    n = 1_000_003
    n8 = n + (-n) % 8
    master = torch.randn(n8, device=dev())[:n]
    g = (torch.randn(n8, device=dev()) * 0.1).to(bf16)[:n]
    p_ref = torch.nn.Parameter(master.clone())
    opt = torch.optim.AdamW([p_ref], lr=1e-2, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1)
    m, v, wb = torch.zeros(n, device=dev()), torch.zeros(n, device=dev()), torch.empty(n, device=dev(), dtype=bf16)
    nsq = torch.zeros(1, device=dev())
    max_norm = 8.0
    for step in (1, 2, 3):
        nsq.zero_()
        ops.sumsq_bf16(g, nsq)
        torch.testing.assert_close(nsq[0], g.float().pow(2).sum(), rtol=1e-4, atol=1e-3)
        coef = min(1.0, max_norm / (math.sqrt(nsq.item()) + 1e-6))
        assert coef < 1.0
        p_ref.grad = g.float() * coef
        opt.step()
        ops.adamw(master, wb, g, m, v, 1e-2, 0.9, 0.95, 1e-8, 0.1, step, nsq, max_norm)
        torch.testing.assert_close(master, p_ref.detach(), rtol=2e-5, atol=2e-6)
        assert torch.equal(wb, master.to(bf16))
