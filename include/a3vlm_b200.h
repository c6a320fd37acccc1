/*
 * a3vlm_b200 -- C ABI of the B200-native hot path (CLIP ViT-L/14 -> projector -> Llama decoder,
 * forward / backward / optimiser / greedy decode) that replaces the library kernels behind the
 * reference's `llama_type` plugin (reference model/accessory/model/meta.py:30-32 picks
 * `accessory.model.LLM.<llama_type>`; the tensor math it then runs is
 * model/accessory/model/LLM/llama_ens5.py:53-539 and model/accessory/model/components.py:10-53).
 *
 * The reference has no C ABI of its own (it is 100 % Python on top of torch / cuBLAS / cuDNN /
 * flash-attn wheels); every entry point below cites the reference lines whose arithmetic it
 * replaces.  Conventions:
 *   - all pointers are DEVICE pointers borrowed from the caller (PyTorch's allocator); the library
 *     allocates nothing and keeps no reference after the call returns;
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued, never synchronised;
 *   - matrices are row-major; `ld*` are row strides in ELEMENTS;
 *   - return value: A3_OK or an error code; a3_last_error() gives the message.  No CPU fallback
 *     exists: a call on a machine without an sm_100 device fails with A3_ERR_CUDA.
 */
#ifndef A3VLM_B200_H_
#define A3VLM_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define A3_OK 0
#define A3_ERR_ARG 1
#define A3_ERR_CUDA 2
#define A3_ERR_UNSUPPORTED 3

/* activation codes for a3_gemm_bf16 */
#define A3_ACT_NONE 0
#define A3_ACT_GELU 1       /* exact erf GELU: open_clip ViT built with pretrained=None (llama_ens5.py:299) */
#define A3_ACT_QUICKGELU 2  /* x*sigmoid(1.702x): OpenAI CLIP weights (llama_ens5.py:297) */

const char* a3_last_error(void);
int a3_version(void);
/* Number of kernels this library has launched since load / since the last reset (bench.py's
 * `gpu_launches`). */
/* Persistent kernels (GEMM) size their grids to at most n SMs (0 = all).  Used by the data-parallel
 * fine-tune step to leave SMs for the concurrent NCCL all-reduce kernels, so that no GEMM CTA has
 * to wait for a collective CTA to retire. */
int a3_set_sm_limit(int n);
long long a3_launch_count(void);
void a3_reset_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * GEMM (tcgen05 / TMEM / TMA).  C[M,N] = epilogue(alpha * sum_k A(m,k) B(n,k)).
 *   a_mn_major = 0: A stored [M][K] (lda >= K);  1: A stored [K][M] (lda >= M)
 *   b_mn_major = 0: B stored [N][K] (ldb >= K);  1: B stored [K][N] (ldb >= N)
 * Epilogue, in order: *alpha, +bias[n] (bf16, added in fp32), [round to bf16 unless out_fp32], activation,
 * [RoPE on columns < rope_cols], +residual[m,n] (bf16), [+C if accumulate], store bf16 / fp32.
 * The bf16 roundings reproduce autocast's: a Linear output is rounded before the activation and
 * before the residual add.  RoPE (rope_cols > 0): columns are heads of 128, pairs
 * (2j,2j+1) rotated by angle table[pos][j] with pos = pos0 + (m % rope_period); `rope_table` is
 * float2 (cos,sin) [max_pos][64].
 * Replaces: every fairscale Column/RowParallelLinear and nn.Linear on the path
 * (llama_ens5.py:63-90,112,169,202-217,267-269,330-333,486,530), the open_clip ViT linears and the
 * conv1 patch embedding as a GEMM (llama_ens5.py:354-370), apply_rotary_emb (llama_ens5.py:118),
 * and their autograd dgrad / wgrad forms.
 * ------------------------------------------------------------------------------------------- */
int a3_gemm_bf16(const void* A, const void* B, void* C, int M, int N, int K, long long lda,
                 long long ldb, long long ldc, int a_mn_major, int b_mn_major, const void* bias,
                 const void* residual, long long ldr, int act, int out_fp32, int accumulate,
                 float alpha, const void* rope_table, int rope_cols, int rope_period, int rope_pos0,
                 void* stream);

/* Gated-SiLU MLP with the activation fused into the tensor-core kernels (llama_ens5.py:213-217):
 *   a3_gemm_swiglu_bf16:  gu[M][2F] = X[M][K] W13[2F][K]^T (W13 = w1 rows then w3 rows) AND
 *                         y[M][F] = bf16(bf16(silu(g)) * u) in one launch (the CTA pair loads 128 rows of
 *                         w1 and the matching 128 rows of w3, so g and u of a feature share a tile).
 *   a3_gemm_dswiglu_bf16: dgu[M][2F] = swiglu'(gu) applied to dy = D[M][Dm] W2[Dm][F] -- the backward
 *                         of the gate fused behind the w2 dgrad; dy is never written to HBM. */
int a3_gemm_swiglu_bf16(const void* X, const void* W13, void* GU, void* Y, int M, int F, int K,
                        long long ldx, long long ldw, long long ldgu, long long ldy, void* stream);
int a3_gemm_dswiglu_bf16(const void* D, const void* W2, const void* GU, void* DGU, int M, int F, int Dm,
                         long long ldd, long long ldw, long long ldgu, long long lddgu, void* stream);

/* Weight-streaming GEMV-shaped linear for decode: Y[M,N] = X[M,K] W[N,K]^T (+residual), M <= 16.
 * HBM-bound (reads W once).  out_fp32 for the LM head (llama_ens5.py:530-531). */
int a3_linear_small_m(const void* X, const void* W, void* Y, int M, int N, int K, long long ldx,
                      long long ldw, long long ldy, const void* residual, long long ldr,
                      int out_fp32, void* stream);

/* Fused decode projection for M <= 8 rows (one launch per projection of a decode step):
 *   prologue  norm_w != NULL: X <- RMSNorm(X) * norm_w (components.py:39-53)
 *             swiglu != 0:    X is gu [M][2K], X <- bf16(silu(g)) * u (llama_ens5.py:213-214)
 *   epilogue  residual | fp32 logits | rope_table != NULL (QKV projection, N = (H + 2*Hkv)*128, columns q | k | v
 *             with Hkv = n_kv_heads key / value heads): RoPE of q, k at position pos (or *pos_dev) and k, v written
 *             into cache[m, pos] (llama_ens5.py:118-127).
 * Launched with programmatic dependent launch: weights are prefetched while the previous kernel of
 * the stream drains. */
int a3_decode_linear(const void* X, const void* W, void* Y, int M, int N, int K, long long ldx,
                     long long ldw, long long ldy, const void* norm_w, float eps, int swiglu,
                     const void* residual, long long ldr, int out_fp32, const float* rope_table,
                     int n_heads, int n_kv_heads, void* kcache, void* vcache, const int* pos_dev, int pos, int max_seq,
                     void* stream);

/* The w1|w3 projection of a decode step with the gated SiLU as its epilogue (llama_ens5.py:213-214): a warp streams
 * row n of w1 and row n of w3 (W13 = [2F][K], w1 rows first) and writes y[m][n] = bf16(bf16(silu(g)) * u) -- gu never
 * reaches HBM and the w2 projection needs no activation prologue.  Optional RMSNorm prologue as a3_decode_linear. */
int a3_decode_linear_gated(const void* X, const void* W13, void* Y, int M, int F, int K, long long ldx,
                           long long ldw, long long ldy, const void* norm_w, float eps, void* stream);

/* Decode projection for a batch of M <= 64 single-token rows on tcgen05 with swapped operands: the
 * weight tile is the 128-row MMA operand, the batch the narrow N side, split-K across a thread-block
 * cluster with a distributed-shared-memory reduction, 8-11 TMA stages in flight (HBM-bound weight
 * stream; llama_ens5.py:112,169,216-217,530 at seqlen 1).  Y[M][N] = X[M][K] W[N][K]^T (+residual). */
int a3_decode_gemm(const void* X, const void* W, void* Y, int M, int N, int K, long long ldx,
                   long long ldw, long long ldy, const void* residual, long long ldr, int out_fp32,
                   void* stream);

/* ---------------------------------------------------------------------------------------------
 * Norms.  x,y bf16 [rows][D]; weights bf16; statistics fp32.
 * RMSNorm: reference model/accessory/model/components.py:39-53 -- y = bf16(x*rsqrt(mean(x^2)+eps)) * w
 * (normalise in fp32, round to bf16, then multiply by the bf16 weight and round again).
 * LayerNorm: nn.LayerNorm eps 1e-5 of the ViT (ln_pre / ln_1 / ln_2 / ln_post) and of
 * visual_proj[1] (llama_ens5.py:330-333): fp32 statistics, output bf16.
 * ------------------------------------------------------------------------------------------- */
/* group_rows > 0: the input rows are grouped -- output row r reads x + (r / group_rows) * group_stride +
 * (r % group_rows) * ldx (elements).  One launch then normalises the text rows [W, L) of every sample of a
 * [B][L][D] residual stream (final norm on h[:, image_words:], llama_ens5.py:485-486).  0: dense rows. */
int a3_rmsnorm_fwd(const void* x, const void* w, void* y, float* rstd, int rows, int D,
                   long long ldx, long long ldy, float eps, int group_rows, long long group_stride,
                   void* stream);
/* dx = d/dx of the above given dy; dw_partial [nparts][D] fp32 (caller sums over parts with
 * a3_colsum_partials).  add_to_dx != 0: dx += (residual stream gradient fused).  x and dx are dense
 * [rows][D], or grouped like a3_rmsnorm_fwd's input (row stride D inside a group); dy, rstd are dense. */
int a3_rmsnorm_bwd(const void* dy, const void* x, const void* w, const float* rstd, void* dx,
                   float* dw_partial, int nparts, int rows, int D, int add_to_dx, int group_rows,
                   long long group_stride, void* stream);
int a3_layernorm_fwd(const void* x, const void* w, const void* b, void* y, float* mean,
                     float* rstd, int rows, int D, long long ldx, long long ldy, float eps,
                     void* stream);
int a3_layernorm_bwd(const void* dy, const void* x, const void* w, const float* mean,
                     const float* rstd, void* dx, float* dw_partial, float* db_partial, int nparts,
                     int rows, int D, void* stream);
/* out[D] (bf16 or fp32) (+)= sum_p partial[p][D] */
int a3_colsum_partials(const float* partial, int nparts, int D, void* out, int out_fp32,
                       int accumulate, void* stream);
/* out[N] (bf16) (+)= sum_m X[m][N] (bias gradients). */
int a3_colsum_bf16(const void* X, int M, int N, long long ldx, void* out, int accumulate,
                   void* stream);

/* ---------------------------------------------------------------------------------------------
 * RoPE (absent reference file LLM/llama.py, call sites llama_ens5.py:118,271-274):
 * table[pos][j] = (cos, sin)(pos * theta^(-2j/128)) as float2 -- built once at init on the host side
 * with the reference's own torch expression (precompute_freqs_cis) so the angles are bit-identical;
 * interleaved pairs; fp32 math, one rounding.
 * ------------------------------------------------------------------------------------------- */
/* In-place rotation of `heads` heads of head_dim at x[row*ld + h*head_dim ..]; pos = pos0 + row %
 * period.  inverse != 0 rotates by -angle (the backward pass). */
int a3_rope_apply(void* x, long long ld, int rows, int heads, int head_dim, const float* table,
                  int period, int pos0, int inverse, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Attention.  Element (b, t, h, d) of q lives at q + (b*Tq + t)*ldq + h*head_dim + d (same for
 * k, v with Tk / ldk / ldv and for o with ldo); head_dim 64 (ViT) or 128 (Llama).
 * causal != 0: right-aligned causal mask, query t sees keys <= t + (Tk - Tq)
 * (llama_ens5.py:181-185); lse [B][H][Tq] fp32 (natural log) is written when non-null.
 * Replaces flash_attn_func / F.scaled_dot_product_attention (llama_ens5.py:142-167) and the
 * nn.MultiheadAttention core of open_clip's ResidualAttentionBlock.
 * ------------------------------------------------------------------------------------------- */
int a3_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H,
                int Tq, int Tk, int head_dim, long long ldq, long long ldk, long long ldv,
                long long ldo, long long q_bstride, long long kv_bstride, float scale, int causal,
                void* stream);
/* Backward of the above for Tq == Tk, head_dim 128 (fine-tune path).  delta [B][H][T] fp32 scratch,
 * dq_acc [B*T][H*128] fp32 scratch (zeroed by the call); dq/dk/dv share the q/k/v layout.
 * rope_table != NULL: q and k were rotated by apply_rotary_emb at positions 0..T-1 (llama_ens5.py:118);
 * dq and dk are returned w.r.t. the UN-rotated projections (rotation by -angle fused into the stores). */
int a3_attn_bwd(const void* q, const void* k, const void* v, const void* o, const void* dout,
                const float* lse, float* delta, float* dq_acc, void* dq, void* dk, void* dv, int B,
                int H, int T, int head_dim, long long ldqkv, long long ldo, long long lddqkv,
                float scale, int causal, const float* rope_table, void* stream);
/* Single-query decode attention over the KV cache (llama_ens5.py:121-129,142-167 with seqlen 1):
 * cache layout [B][max_seq][Hkv][128] bf16 (Hkv = n_kv_heads; query head h reads group head h / (H / Hkv), the
 * reference's repeat_kv, llama_ens5.py:131-133); q [B][H*128]; ctx = number of valid positions. */
/* scratch: B*H*(32*130 + 1) floats, ZERO-INITIALISED once by the caller (the tail holds per-(b,h)
 * arrival counters that reset themselves; the last split to finish merges the partial results).
 * pos_dev != NULL (CUDA-graph replay): ctx = *pos_dev + 1 is read on the device and `ctx` is only an
 * upper bound used to size the split. */
int a3_attn_decode(const void* q, const void* kcache, const void* vcache, void* o, float* scratch,
                   int B, int H, int Hkv, int head_dim, int ctx, int max_seq, long long ldq, long long ldo,
                   float scale, const int* pos_dev, void* stream);
/* Decode-step fusion of apply_rotary_emb + cache write (llama_ens5.py:118-127, seqlen 1): rotates q
 * and k of qkv [B][(H + 2*Hkv)*128] in place at position pos (or *pos_dev) and writes k, v into
 * cache[b, pos]. */
int a3_decode_rope_kv(void* qkv, long long ld, void* kcache, void* vcache, int B, int H, int Hkv, int head_dim,
                      const float* table, int pos, const int* pos_dev, int max_seq, void* stream);
/* Grouped-query attention glue (n_kv_heads < n_heads; repeat_kv of the absent LLM/llama.py, call llama_ens5.py:131-133):
 * a3_repeat_kv expands compact [rows][(H + 2*Hkv)*hd] (q | k | v) to [rows][3*H*hd] -- expanded head h = group head
 * h / (H / Hkv) -- so the prefill / training attention kernels run unchanged on H heads; a3_reduce_kv is its backward
 * (dq copied, dk / dv summed over the copies of a group in fp32, one rounding). */
int a3_repeat_kv(const void* qkv_compact, long long ld_compact, void* qkv, long long ld, long long rows, int n_heads,
                 int n_kv_heads, int head_dim, void* stream);
int a3_reduce_kv(const void* dqkv, long long ld, void* dqkv_compact, long long ld_compact, long long rows, int n_heads,
                 int n_kv_heads, int head_dim, void* stream);
/* KV-cache append (llama_ens5.py:121-127): copy rows [B][T] of k and v (strided, ld) into
 * cache[:, start:start+T]. */
int a3_kv_append(const void* k, const void* v, long long ld, void* kcache, void* vcache, int B,
                 int T, int H, int head_dim, int start, int max_seq, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Gated-SiLU MLP glue (llama_ens5.py:213-217): y = bf16(bf16(silu(g)) * u), gu = [rows][2F] (g | u).
 * ------------------------------------------------------------------------------------------- */
int a3_swiglu_fwd(const void* gu, void* y, int rows, int F, void* stream);
int a3_swiglu_bwd(const void* gu, const void* dy, void* dgu, int rows, int F, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Embedding gather + image splice (llama_ens5.py:464-479,495-511).
 * h[b, 0] = E[tok[b,0]]; per view v: start_img, Nimg tokens img[(v*B + b)*Nimg + i], end_img;
 * h[b, W + t] = E[tok[b,t]], t >= 1, W = n_views * (Nimg + 2) (W = 0 when img is null).  n_views = 1:
 * single view; 5: the reference's tiling (llama_ens5.py:383-385,457).
 * ------------------------------------------------------------------------------------------- */
int a3_embed_splice(const long long* tokens, const void* E, const void* img, const void* start_img,
                    const void* end_img, void* h, int B, int T, int Nimg, int D, long long ld_img,
                    int n_views, void* stream);
/* RGB-D variant (LLM/llama_ens5_2images.py:487-501,517-547): a second image stream follows the first,
 * wrapped in its own tags: BOS | n_views x (start_img, Nimg, end_img) | n_views x (start_depth_img, Nimg,
 * end_depth_img) | text.  img holds the 2*n_views views of both streams, view-major
 * (img[((s*n_views + v)*B + b)*Nimg + i]); W = 2*n_views*(Nimg+2).  Forward only: the reference never
 * feeds depth images to the training forward (meta.py:251 passes `images` alone). */
int a3_embed_splice_rgbd(const long long* tokens, const void* E, const void* img, const void* start_img,
                         const void* end_img, const void* start_depth_img, const void* end_depth_img,
                         void* h, int B, int T, int Nimg, int D, long long ld_img, int n_views,
                         void* stream);
/* Backward: dE[tok] += dh rows of text positions (bf16 atomics), d_img, d_start, d_end. */
int a3_embed_splice_bwd(const long long* tokens, const void* dh, void* dE, void* d_img,
                        float* d_start_f32, float* d_end_f32, int B, int T, int Nimg, int D,
                        int n_views, void* stream);
/* The reference's five views of a (2V x 2V) image (llama_ens5.py:383-385): views[0] = bicubic
 * down-sample of the fp16-rounded image (fp32 accumulate, fp16 result), views[1..4] = quadrants
 * TL, TR, BL, BR.  views: fp32 [5][B][3][V][V]. */
int a3_build_views(const void* image, int image_is_fp32, float* views, int B, int V, void* stream);

/* ViT front end (llama_ens5.py:354-363).  patches: NCHW image (fp32 or bf16) -> A [B*g*g][Kpad]
 * bf16 with k = (c, ky, kx), zero padded to Kpad.  assemble: x[b,0] = cls + pos[0];
 * x[b,1+p] = patch_out[b,p] + pos[1+p]; then ln_pre -> y (bf16). */
int a3_vit_patchify(const void* image, int image_is_fp32, void* A, int B, int S, int patch, int Kpad,
                    void* stream);
/* im2col-free patch embedding (the conv1 of llama_ens5.py:354-357 as ONE tcgen05 GEMM whose A operand is loaded by
 * 4-D TMA boxes straight from the image into the UMMA shared-memory layout; no patch matrix in HBM).
 * a3_vit_pad_cast: the autocast cast of the NCHW image (fp32 or bf16) to bf16 [B][3][S][g][16], every patch column on
 * a 16-element pitch (kx >= patch zero) -- TMA needs 16-byte aligned box origins, raw 14-pixel patches are 28 bytes.
 * a3_vit_patch_embed: image = that buffer; w [D][768] bf16 = the (D, 3, patch, patch) kernel zero-padded to 16 x 16
 * windows, k = c * 256 + ky * 16 + kx; out [B * (S/patch)^2][ldo] bf16, rows in (image, py, px) order like
 * a3_vit_patchify's.  a3_cast_f32_bf16: plain fp32 -> bf16 cast of an array. */
int a3_vit_pad_cast(const void* image, int image_is_fp32, void* out, int B, int S, int patch, void* stream);
int a3_vit_patch_embed(const void* image, const void* w, void* out, int B, int S, int patch, int D, long long ldo,
                       void* stream);
int a3_cast_f32_bf16(const float* in, void* out, long long n, void* stream);
int a3_vit_assemble_lnpre(const void* patch_out, const void* cls, const void* pos, const void* ln_w,
                          const void* ln_b, void* x, int B, int Np, int D, float eps, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Loss (meta.py:256-262): CrossEntropyLoss(ignore_index=0), mean over non-ignored rows, computed
 * in fp32 from bf16 logits [rows][V].  row r uses label labels[r] (already shifted by the caller).
 * Writes per-row loss (0 for ignored) and overwrites logits with dlogits = (softmax - onehot) *
 * grad_scale / count when want_grad != 0.  `count` (device int) = number of non-ignored rows,
 * must be pre-computed with a3_count_nonzero.
 * ------------------------------------------------------------------------------------------- */
int a3_count_nonzero(const long long* labels, int n, int* count, void* stream);
/* The shift of meta.py:256-257 (logits[:, :-1] vs labels[:, 1:]) as a label buffer aligned with the
 * logit rows: shifted[b][t] = labels[b][t+1] for t < T-1, 0 (ignored) for t = T-1; labels has row
 * stride ld_labels.  count = number of non-zero shifted labels. */
int a3_shift_labels(const long long* labels, long long ld_labels, long long* shifted, int B, int T,
                    int* count, void* stream);
int a3_cross_entropy(void* logits, long long ld, const long long* labels, float* row_loss,
                     const int* count, int rows, int V, float grad_scale, int want_grad,
                     void* stream);
/* loss = sum(row_loss) / max(count,1) -> out[0] (fp32) */
int a3_loss_reduce(const float* row_loss, int rows, const int* count, float* out, void* stream);
/* Greedy sampling (meta.py:461): argmax over fp32 logits [B][V], lowest index wins ties. */
int a3_argmax(const float* logits, long long ld, int B, int V, long long* out, void* stream);
/* Nucleus sampling (meta.py:456-459,568-583) in one launch, one CTA per row: probs =
 * softmax(logits / temperature) in fp32; in descending order (ties by token id) keep the entries whose
 * exclusive cumulative sum is <= top_p, renormalise, and draw by inverse CDF with the caller's uniform
 * u[b] in [0,1) -- the distribution torch.multinomial samples in the reference; the uniform is an input
 * because multinomial's random stream is not reproducible across devices.  No sort: the crossing value
 * of the sorted cumulative sum is found by bisection over the float bit pattern.  V <= 51200. */
int a3_sample_top_p(const float* logits, long long ld, int B, int V, float temperature, float top_p,
                    const float* u, long long* out, void* stream);
/* The body of the generation loop after the token choice (meta.py:463-477) for the whole batch in one
 * launch: prompt force-feeding (prompt position = cur_pos < prompt_len[b]), tokens[b][cur_pos] write,
 * stop_pos / stopped update against n_stop stop sequences (stop_tokens flattened, sequence s =
 * stop_tokens[stop_offsets[s] .. stop_offsets[s+1])), all_stopped[0] = all(stopped) (the `break`
 * condition of :476, read by the host whenever it likes -- no per-token sync), fed_token[b] (nullable) =
 * the token written.  tokens int64 [B][ld_tokens]; stopped uint8 [B]; stop_pos int64 [B]. */
int a3_generate_step(long long* tokens, long long ld_tokens, const int* prompt_len,
                     const long long* next_token, int cur_pos, const long long* stop_tokens,
                     const int* stop_offsets, int n_stop, unsigned char* stopped, long long* stop_pos,
                     int* all_stopped, long long* fed_token, int B, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Optimiser (main_finetune.py:284-285, util/misc.py:586-599, util/clip_grad.py:187-193).
 * sumsq: out[0] += sum(g^2) over n bf16 gradients (fp32 accumulate, deterministic).  `scratch` is
 * A3_SUMSQ_SCRATCH_BYTES of caller-owned device memory, zeroed once before first use (the kernel leaves
 * it ready for the next call); calls that may be in flight concurrently (different streams) must use
 * different scratch buffers.  adamw: one fused pass over
 * a flat range -- g = bf16 grad * clip_coef(*norm_sq), m,v fp32, master fp32, shadow bf16 weight.
 * clip: coef = min(1, max_norm / (sqrt(norm_sq) + 1e-6)); max_norm <= 0 disables clipping.
 * ------------------------------------------------------------------------------------------- */
#define A3_SUMSQ_SCRATCH_BYTES 8192
int a3_sumsq_bf16(const void* g, long long n, float* out, void* scratch, void* stream);
/* Grid of a3_adamw in CTAs per SM (default 16; 1-2 when the update runs on a side stream underneath
 * the next step's forward, so that its CTAs co-reside with the persistent GEMM CTAs). */
int a3_set_adamw_blocks_per_sm(int n);
int a3_adamw(float* master, void* weight_bf16, const void* grad_bf16, float* m, float* v,
             long long n, float lr, float beta1, float beta2, float eps, float weight_decay,
             int step, const float* norm_sq, float max_norm, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* A3VLM_B200_H_ */
