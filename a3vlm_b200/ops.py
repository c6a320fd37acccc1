"""Thin tensor-level wrappers over the C ABI (include/a3vlm_b200.h): PyTorch tensors in, PyTorch
tensors out, the current CUDA stream, no allocation inside the library.  Every function here
launches hand-written sm_100a kernels; there is no torch fallback -- a missing library raises
`A3LibraryError`, a non-CUDA tensor raises `A3Error`.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _lib

bf16 = torch.bfloat16
ACT_NONE, ACT_GELU, ACT_QUICKGELU = 0, 1, 2


def _L():
    return _lib.load()


def _p(t: Optional[torch.Tensor]):
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.A3Error("a3vlm_b200 ops need CUDA tensors (no CPU fallback)")
    return ctypes.c_void_p(t.data_ptr())


def _st():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def set_sm_limit(n: int) -> None:
    _lib.check(_L().a3_set_sm_limit(int(n)), "a3_set_sm_limit")


def launch_count() -> int:
    return int(_L().a3_launch_count())


def reset_launch_count() -> None:
    _L().a3_reset_launch_count()


# ----------------------------------------------------------------------------------------- GEMM
def gemm(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None, *, M=None, N=None,
         K=None, a_mn=False, b_mn=False, bias=None, residual=None, act=ACT_NONE, out_fp32=False,
         accumulate=False, alpha=1.0, rope=None, rope_cols=0, rope_period=1, rope_pos0=0):
    """C[M,N] = epilogue(alpha * A @ B^T).  a: [M,K] (or [K,M] when a_mn); b: [N,K] (or [K,N] when
    b_mn).  2-D tensors with unit inner stride; row strides are honoured."""
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    assert a.dtype == bf16 and b.dtype == bf16
    if M is None:
        M = a.shape[1] if a_mn else a.shape[0]
    if K is None:
        K = a.shape[0] if a_mn else a.shape[1]
    if N is None:
        N = b.shape[1] if b_mn else b.shape[0]
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=torch.float32 if out_fp32 else bf16)
    assert out.stride(-1) == 1 and out.dtype == (torch.float32 if out_fp32 else bf16)
    if residual is not None:
        assert residual.dtype == bf16 and residual.stride(-1) == 1
    rc = _L().a3_gemm_bf16(_p(a), _p(b), _p(out), M, N, K, a.stride(0), b.stride(0), out.stride(0),
                           int(a_mn), int(b_mn), _p(bias), _p(residual),
                           residual.stride(0) if residual is not None else 0, act, int(out_fp32),
                           int(accumulate), float(alpha), _p(rope), rope_cols if rope is not None else 0,
                           rope_period, rope_pos0, _st())
    _lib.check(rc, "a3_gemm_bf16")
    return out


def gemm_swiglu(x, w13, gu, y):
    """gu = x @ w13^T ([M][2F], g | u) and y = bf16(silu(g)) * u in one tcgen05 launch."""
    M, K = x.shape
    F = w13.shape[0] // 2
    rc = _L().a3_gemm_swiglu_bf16(_p(x), _p(w13), _p(gu), _p(y), M, F, K, x.stride(0), w13.stride(0),
                                  gu.stride(0), y.stride(0), _st())
    _lib.check(rc, "a3_gemm_swiglu_bf16")
    return gu, y


def gemm_dswiglu(d, w2, gu, dgu):
    """dgu = swiglu_bwd(gu, d @ w2) with w2 stored [Dm][F]; dy stays on chip."""
    M, Dm = d.shape
    F = w2.shape[1]
    rc = _L().a3_gemm_dswiglu_bf16(_p(d), _p(w2), _p(gu), _p(dgu), M, F, Dm, d.stride(0), w2.stride(0),
                                   gu.stride(0), dgu.stride(0), _st())
    _lib.check(rc, "a3_gemm_dswiglu_bf16")
    return dgu


def linear_small_m(x, w, out=None, residual=None, out_fp32=False):
    M, K = x.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty(M, N, device=x.device, dtype=torch.float32 if out_fp32 else bf16)
    rc = _L().a3_linear_small_m(_p(x), _p(w), _p(out), M, N, K, x.stride(0), w.stride(0),
                                out.stride(0), _p(residual),
                                residual.stride(0) if residual is not None else 0, int(out_fp32), _st())
    _lib.check(rc, "a3_linear_small_m")
    return out


def decode_linear(x, w, out, *, K=None, norm_w=None, eps=1e-5, swiglu=False, residual=None, out_fp32=False,
                  rope=None, n_heads=0, n_kv_heads=None, kcache=None, vcache=None, pos_dev=None, pos=0, max_seq=0):
    M = x.shape[0]
    N = w.shape[0]
    if K is None:
        K = w.shape[1]
    rc = _L().a3_decode_linear(_p(x), _p(w), _p(out), M, N, K, x.stride(0), w.stride(0), out.stride(0),
                               _p(norm_w), float(eps), int(swiglu), _p(residual),
                               residual.stride(0) if residual is not None else 0, int(out_fp32), _p(rope),
                               n_heads, n_heads if n_kv_heads is None else n_kv_heads, _p(kcache), _p(vcache), _p(pos_dev),
                               int(pos), max_seq, _st())
    _lib.check(rc, "a3_decode_linear")
    return out


def decode_linear_gated(x, w13, out, *, norm_w=None, eps=1e-5):
    """y [M][F] = bf16(silu(x w1^T)) * (x w3^T) with w13 = [2F][K] (w1 rows, then w3 rows); optional RMSNorm prologue."""
    M, F, K = x.shape[0], w13.shape[0] // 2, w13.shape[1]
    rc = _L().a3_decode_linear_gated(_p(x), _p(w13), _p(out), M, F, K, x.stride(0), w13.stride(0), out.stride(0),
                                     _p(norm_w), float(eps), _st())
    _lib.check(rc, "a3_decode_linear_gated")
    return out


def decode_gemm(x, w, out, residual=None, out_fp32=False):
    M, K = x.shape
    N = w.shape[0]
    rc = _L().a3_decode_gemm(_p(x), _p(w), _p(out), M, N, K, x.stride(0), w.stride(0), out.stride(0),
                             _p(residual), residual.stride(0) if residual is not None else 0, int(out_fp32), _st())
    _lib.check(rc, "a3_decode_gemm")
    return out


# ----------------------------------------------------------------------------------------- norms
def rmsnorm_fwd(x, w, eps, out=None, rstd=None):
    """x [rows][D] (row stride honoured), or [G][R][D] with x.stride(2) == 1: the G groups of R rows are
    normalised by one launch into dense out [G*R][D] (the text rows of every sample of a [B][L][D] stream)."""
    if x.dim() == 3:
        G, R, D = x.shape
        assert x.stride(2) == 1
        rows, ldx, group_rows, group_stride = G * R, x.stride(1), R, x.stride(0)
    else:
        rows, D = x.shape
        ldx, group_rows, group_stride = x.stride(0), 0, 0
    if out is None:
        out = torch.empty(rows, D, device=x.device, dtype=bf16)
    rc = _L().a3_rmsnorm_fwd(_p(x), _p(w), _p(out), _p(rstd), rows, D, ldx, out.stride(0),
                             float(eps), group_rows, group_stride, _st())
    _lib.check(rc, "a3_rmsnorm_fwd")
    return out


def rmsnorm_bwd(dy, x, w, rstd, dx, dw_partial, add_to_dx=False):
    """x / dx [rows][D] contiguous, or both [G][R][D] views with dense rows inside a group (same strides);
    dy and rstd are dense over the G*R rows."""
    if x.dim() == 3:
        G, R, D = x.shape
        assert x.stride() == dx.stride() and x.stride(2) == 1 and x.stride(1) == D and dy.is_contiguous()
        rows, group_rows, group_stride = G * R, R, x.stride(0)
    else:
        rows, D = x.shape
        assert x.is_contiguous() and dy.is_contiguous() and dx.is_contiguous()
        group_rows = group_stride = 0
    nparts = dw_partial.shape[0]
    rc = _L().a3_rmsnorm_bwd(_p(dy), _p(x), _p(w), _p(rstd), _p(dx), _p(dw_partial), nparts, rows, D,
                             int(add_to_dx), group_rows, group_stride, _st())
    _lib.check(rc, "a3_rmsnorm_bwd")
    return dx


def layernorm_fwd(x, w, b, eps, out=None, mean=None, rstd=None):
    rows, D = x.shape
    if out is None:
        out = torch.empty(rows, D, device=x.device, dtype=bf16)
    rc = _L().a3_layernorm_fwd(_p(x), _p(w), _p(b), _p(out), _p(mean), _p(rstd), rows, D,
                               x.stride(0), out.stride(0), float(eps), _st())
    _lib.check(rc, "a3_layernorm_fwd")
    return out


def layernorm_bwd(dy, x, w, mean, rstd, dx, dw_partial, db_partial):
    rows, D = x.shape
    assert x.is_contiguous() and dy.is_contiguous() and dx.is_contiguous()
    rc = _L().a3_layernorm_bwd(_p(dy), _p(x), _p(w), _p(mean), _p(rstd), _p(dx), _p(dw_partial),
                               _p(db_partial), dw_partial.shape[0], rows, D, _st())
    _lib.check(rc, "a3_layernorm_bwd")
    return dx


def colsum_partials(partial, out, accumulate=False):
    nparts, D = partial.shape
    rc = _L().a3_colsum_partials(_p(partial), nparts, D, _p(out), int(out.dtype == torch.float32),
                                 int(accumulate), _st())
    _lib.check(rc, "a3_colsum_partials")
    return out


def colsum_bf16(x, out, accumulate=False):
    M, N = x.shape
    rc = _L().a3_colsum_bf16(_p(x), M, N, x.stride(0), _p(out), int(accumulate), _st())
    _lib.check(rc, "a3_colsum_bf16")
    return out


# ----------------------------------------------------------------------------------------- RoPE
def rope_table(head_dim: int, end: int, theta: float = 10000.0, device="cpu") -> torch.Tensor:
    """float32 [end][head_dim/2][2] = (cos, sin) of the reference's precompute_freqs_cis angles
    (absent LLM/llama.py, call site llama_ens5.py:271-274) -- same torch expression so
    the angles are bit-identical to the oracle's."""
    freqs = 1.0 / (theta ** (torch.arange(0, head_dim, 2)[: head_dim // 2].float() / head_dim))
    t = torch.arange(end).float()
    ang = torch.outer(t, freqs)
    cis = torch.polar(torch.ones_like(ang), ang)
    return torch.view_as_real(cis).contiguous().to(device)


def rope_apply(x, heads, head_dim, table, period, pos0=0, inverse=False):
    """In-place on a 2-D strided view x [rows][>= heads*head_dim]."""
    rows = x.shape[0]
    rc = _L().a3_rope_apply(_p(x), x.stride(0), rows, heads, head_dim, _p(table), period, pos0,
                            int(inverse), _st())
    _lib.check(rc, "a3_rope_apply")
    return x


# ----------------------------------------------------------------------------------------- attention
def attn_fwd(q, k, v, B, H, Tq, Tk, head_dim, scale, causal, out=None, lse=None,
             q_bstride=None, kv_bstride=None):
    """q [B*Tq][ldq] view, k/v [B*Tk][ld] views (column offset already applied by slicing)."""
    if out is None:
        out = torch.empty(B * Tq, H * head_dim, device=q.device, dtype=bf16)
    ldq, ldk, ldv, ldo = q.stride(0), k.stride(0), v.stride(0), out.stride(0)
    rc = _L().a3_attn_fwd(_p(q), _p(k), _p(v), _p(out), _p(lse), B, H, Tq, Tk, head_dim, ldq, ldk, ldv,
                          ldo, q_bstride if q_bstride is not None else Tq * ldq,
                          kv_bstride if kv_bstride is not None else Tk * ldk, float(scale),
                          int(causal), _st())
    _lib.check(rc, "a3_attn_fwd")
    return out


def attn_bwd(q, k, v, o, dout, lse, delta, dq_acc, dq, dk, dv, B, H, T, head_dim, scale, causal, rope=None):
    assert q.stride(0) == k.stride(0) == v.stride(0)
    assert dq.stride(0) == dk.stride(0) == dv.stride(0)
    rc = _L().a3_attn_bwd(_p(q), _p(k), _p(v), _p(o), _p(dout), _p(lse), _p(delta), _p(dq_acc), _p(dq),
                          _p(dk), _p(dv), B, H, T, head_dim, q.stride(0), o.stride(0), dq.stride(0),
                          float(scale), int(causal), _p(rope), _st())
    _lib.check(rc, "a3_attn_bwd")


def attn_decode(q, kcache, vcache, out, scratch, B, H, head_dim, ctx, max_seq, scale, pos_dev=None, n_kv_heads=None):
    rc = _L().a3_attn_decode(_p(q), _p(kcache), _p(vcache), _p(out), _p(scratch), B, H,
                             H if n_kv_heads is None else n_kv_heads, head_dim, ctx,
                             max_seq, q.stride(0), out.stride(0), float(scale), _p(pos_dev), _st())
    _lib.check(rc, "a3_attn_decode")
    return out


def decode_rope_kv(qkv, kcache, vcache, B, H, head_dim, table, pos, max_seq, pos_dev=None, n_kv_heads=None):
    rc = _L().a3_decode_rope_kv(_p(qkv), qkv.stride(0), _p(kcache), _p(vcache), B, H,
                                H if n_kv_heads is None else n_kv_heads, head_dim, _p(table),
                                int(pos), _p(pos_dev), max_seq, _st())
    _lib.check(rc, "a3_decode_rope_kv")


def repeat_kv(qkv_compact, qkv, n_heads, n_kv_heads, head_dim):
    """[rows][(H + 2 Hkv) hd] (q | k | v) -> [rows][3 H hd]: expanded head h = group head h // (H // Hkv)."""
    rc = _L().a3_repeat_kv(_p(qkv_compact), qkv_compact.stride(0), _p(qkv), qkv.stride(0), qkv.shape[0], n_heads,
                           n_kv_heads, head_dim, _st())
    _lib.check(rc, "a3_repeat_kv")
    return qkv


def reduce_kv(dqkv, dqkv_compact, n_heads, n_kv_heads, head_dim):
    """Backward of repeat_kv: dq copied, dk / dv summed over the copies of each group."""
    rc = _L().a3_reduce_kv(_p(dqkv), dqkv.stride(0), _p(dqkv_compact), dqkv_compact.stride(0), dqkv.shape[0], n_heads,
                           n_kv_heads, head_dim, _st())
    _lib.check(rc, "a3_reduce_kv")
    return dqkv_compact


def kv_append(k, v, kcache, vcache, B, T, H, head_dim, start, max_seq):
    rc = _L().a3_kv_append(_p(k), _p(v), k.stride(0), _p(kcache), _p(vcache), B, T, H, head_dim, start,
                           max_seq, _st())
    _lib.check(rc, "a3_kv_append")


# ----------------------------------------------------------------------------------------- MLP glue
def swiglu_fwd(gu, out=None):
    rows, F2 = gu.shape
    assert gu.is_contiguous()
    if out is None:
        out = torch.empty(rows, F2 // 2, device=gu.device, dtype=bf16)
    _lib.check(_L().a3_swiglu_fwd(_p(gu), _p(out), rows, F2 // 2, _st()), "a3_swiglu_fwd")
    return out


def swiglu_bwd(gu, dy, dgu):
    rows, F2 = gu.shape
    assert gu.is_contiguous() and dy.is_contiguous() and dgu.is_contiguous()
    _lib.check(_L().a3_swiglu_bwd(_p(gu), _p(dy), _p(dgu), rows, F2 // 2, _st()), "a3_swiglu_bwd")
    return dgu


# ----------------------------------------------------------------------------------------- embed / ViT front
def embed_splice(tokens, E, img, start_img, end_img, out, B, T, Nimg, D, n_views=1):
    rc = _L().a3_embed_splice(_p(tokens), _p(E), _p(img), _p(start_img), _p(end_img), _p(out), B, T,
                              Nimg, D, img.stride(-2) if img is not None else 0, n_views, _st())
    _lib.check(rc, "a3_embed_splice")
    return out


def embed_splice_rgbd(tokens, E, img, start_img, end_img, start_depth_img, end_depth_img, out, B, T, Nimg, D,
                      n_views=1):
    """RGB stream then depth stream, each view wrapped in its own tags (llama_ens5_2images.py:517-547)."""
    rc = _L().a3_embed_splice_rgbd(_p(tokens), _p(E), _p(img), _p(start_img), _p(end_img), _p(start_depth_img),
                                   _p(end_depth_img), _p(out), B, T, Nimg, D, img.stride(-2), n_views, _st())
    _lib.check(rc, "a3_embed_splice_rgbd")
    return out


def embed_splice_bwd(tokens, dh, dE, d_img, d_start, d_end, B, T, Nimg, D, n_views=1):
    rc = _L().a3_embed_splice_bwd(_p(tokens), _p(dh), _p(dE), _p(d_img), _p(d_start), _p(d_end), B, T,
                                  Nimg, D, n_views, _st())
    _lib.check(rc, "a3_embed_splice_bwd")


def build_views(image, out, B, V):
    assert image.is_contiguous() and image.dtype in (torch.float32, bf16) and out.dtype == torch.float32
    _lib.check(_L().a3_build_views(_p(image), int(image.dtype == torch.float32), _p(out), B, V, _st()),
               "a3_build_views")
    return out


def vit_patchify(image, out, B, S, patch, Kpad):
    assert image.is_contiguous() and image.dtype in (torch.float32, bf16)
    rc = _L().a3_vit_patchify(_p(image), int(image.dtype == torch.float32), _p(out), B, S, patch, Kpad,
                              _st())
    _lib.check(rc, "a3_vit_patchify")
    return out


def vit_pad_cast(image, out, B, S, patch):
    """NCHW fp32 / bf16 image -> bf16 [B][3][S][S/patch][16] (patch columns on a 16-element pitch) for vit_patch_embed."""
    assert image.is_contiguous() and image.dtype in (torch.float32, bf16) and out.dtype == bf16 and out.is_contiguous()
    assert out.numel() == B * 3 * S * (S // patch) * 16
    rc = _L().a3_vit_pad_cast(_p(image), int(image.dtype == torch.float32), _p(out), B, S, patch, _st())
    _lib.check(rc, "a3_vit_pad_cast")
    return out


def vit_patch_embed(image_padded, w768, out, B, S, patch):
    """conv1 (kernel = stride = patch) as one tcgen05 GEMM fed by 4-D TMA boxes over the padded bf16 image."""
    assert image_padded.dtype == bf16 and image_padded.is_contiguous() and w768.is_contiguous() and out.stride(1) == 1
    rc = _L().a3_vit_patch_embed(_p(image_padded), _p(w768), _p(out), B, S, patch, w768.shape[0], out.stride(0), _st())
    _lib.check(rc, "a3_vit_patch_embed")
    return out


def cast_f32_bf16(x, out):
    assert x.dtype == torch.float32 and x.is_contiguous() and out.dtype == bf16 and out.is_contiguous()
    _lib.check(_L().a3_cast_f32_bf16(_p(x), _p(out), x.numel(), _st()), "a3_cast_f32_bf16")
    return out


def vit_assemble_lnpre(patch_out, cls, pos, ln_w, ln_b, out, B, Np, D, eps):
    rc = _L().a3_vit_assemble_lnpre(_p(patch_out), _p(cls), _p(pos), _p(ln_w), _p(ln_b), _p(out), B, Np,
                                    D, float(eps), _st())
    _lib.check(rc, "a3_vit_assemble_lnpre")
    return out


# ----------------------------------------------------------------------------------------- loss / sampling
def count_nonzero(labels, count):
    _lib.check(_L().a3_count_nonzero(_p(labels), labels.numel(), _p(count), _st()), "a3_count_nonzero")


def shift_labels(labels, shifted, count):
    B, T = labels.shape
    rc = _L().a3_shift_labels(_p(labels), labels.stride(0), _p(shifted), B, T, _p(count), _st())
    _lib.check(rc, "a3_shift_labels")


def cross_entropy(logits, labels, row_loss, count, grad_scale=1.0, want_grad=True):
    rows, V = logits.shape
    rc = _L().a3_cross_entropy(_p(logits), logits.stride(0), _p(labels), _p(row_loss), _p(count), rows,
                               V, float(grad_scale), int(want_grad), _st())
    _lib.check(rc, "a3_cross_entropy")


def loss_reduce(row_loss, count, out):
    _lib.check(_L().a3_loss_reduce(_p(row_loss), row_loss.numel(), _p(count), _p(out), _st()),
               "a3_loss_reduce")


def argmax(logits, out):
    B, V = logits.shape
    _lib.check(_L().a3_argmax(_p(logits), logits.stride(0), B, V, _p(out), _st()), "a3_argmax")
    return out


def sample_top_p(logits, temperature, top_p, u, out=None):
    """Nucleus sampling of one token per row from fp32 logits [B][V] (meta.py:456-459,568-583) with the
    uniforms u [B] in [0,1) supplied by the caller."""
    B, V = logits.shape
    assert logits.dtype == torch.float32 and logits.stride(1) == 1 and u.dtype == torch.float32 and u.numel() == B
    if out is None:
        out = torch.empty(B, device=logits.device, dtype=torch.int64)
    rc = _L().a3_sample_top_p(_p(logits), logits.stride(0), B, V, float(temperature), float(top_p), _p(u),
                              _p(out), _st())
    _lib.check(rc, "a3_sample_top_p")
    return out


def generate_step(tokens, prompt_len, next_token, cur_pos, stop_tokens, stop_offsets, n_stop, stopped, stop_pos,
                  all_stopped=None, fed_token=None):
    """meta.py:463-477 for the whole batch in one launch (see include/a3vlm_b200.h)."""
    B = tokens.shape[0]
    assert tokens.dtype == torch.int64 and tokens.stride(1) == 1 and prompt_len.dtype == torch.int32
    assert stopped.dtype in (torch.bool, torch.uint8) and stop_pos.dtype == torch.int64
    rc = _L().a3_generate_step(_p(tokens), tokens.stride(0), _p(prompt_len), _p(next_token), int(cur_pos),
                               _p(stop_tokens), _p(stop_offsets), int(n_stop), _p(stopped), _p(stop_pos),
                               _p(all_stopped), _p(fed_token), B, _st())
    _lib.check(rc, "a3_generate_step")


# ----------------------------------------------------------------------------------------- optimiser
SUMSQ_SCRATCH_BYTES = 8192  # A3_SUMSQ_SCRATCH_BYTES
_sumsq_scratch = {}


def sumsq_scratch(device) -> torch.Tensor:
    """A zeroed scratch buffer for a3_sumsq_bf16 (partials + arrival counter)."""
    return torch.zeros(SUMSQ_SCRATCH_BYTES // 4, device=device, dtype=torch.float32)


def sumsq_bf16(g, out, scratch=None):
    """out[0] += sum(g^2).  Without `scratch` one buffer per (device, current stream) is kept here, so
    calls on different streams never share the kernel's partial sums."""
    if scratch is None:
        key = (g.device, torch.cuda.current_stream(g.device).cuda_stream)
        scratch = _sumsq_scratch.get(key)
        if scratch is None:
            scratch = _sumsq_scratch[key] = sumsq_scratch(g.device)
    _lib.check(_L().a3_sumsq_bf16(_p(g), g.numel(), _p(out), _p(scratch), _st()), "a3_sumsq_bf16")


def set_adamw_blocks_per_sm(n: int) -> None:
    _lib.check(_L().a3_set_adamw_blocks_per_sm(int(n)), "a3_set_adamw_blocks_per_sm")


def adamw(master, weight_bf16, grad_bf16, m, v, lr, beta1, beta2, eps, weight_decay, step,
          norm_sq=None, max_norm=0.0):
    rc = _L().a3_adamw(_p(master), _p(weight_bf16), _p(grad_bf16), _p(m), _p(v), master.numel(),
                       float(lr), float(beta1), float(beta2), float(eps), float(weight_decay),
                       int(step), _p(norm_sq), float(max_norm), _st())
    _lib.check(rc, "a3_adamw")
