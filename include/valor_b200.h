/* valor_b200 — C ABI of the B200-native kernels behind VALOR's tri-modal pretraining step.
 *
 * The reference (TXH-mercury/VALOR) has no FFI / operator registry: its only seams are the
 * Python module classes and functions named below.  Each entry point here
 * states the reference interface whose arithmetic it replaces (file:line under the
 * reference root).  INTEGRATION.md shows the Python-side binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless stated; the caller owns all memory
 *     (outputs, workspaces, saved-for-backward tensors); nothing is retained past return;
 *   - `dtype`: 0 = fp32 (parity mode), 1 = bf16 (perf mode).  Parameters that are vectors
 *     (bias, LayerNorm gamma/beta, embedding tables, bias tables) are always fp32; gradients
 *     of parameters are always accumulated (+=) into fp32 buffers;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, no host sync,
 *     no allocation: every entry point is CUDA-graph capturable and re-entrant;
 *   - return 0 on success, non-zero on error with a message in valor_last_error().
 */
#ifndef VALOR_B200_H
#define VALOR_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define VALOR_DT_F32 0
#define VALOR_DT_BF16 1

#define VALOR_ACT_NONE 0
#define VALOR_ACT_GELU 1      /* erf GELU: model/bert.py:52-57, model/transformer.py:32-38, nn.GELU (videoswin.py:58) */
#define VALOR_ACT_QUICKGELU 2 /* model/clip.py:167-169 */
#define VALOR_ACT_RELU 3      /* model/pretrain.py:105 */

#define VALOR_BACKEND_AUTO 0
#define VALOR_BACKEND_TENSOR 1 /* tcgen05 / TMEM / TMA (bf16 only) */
#define VALOR_BACKEND_SIMT 2
#define VALOR_BACKEND_MMA_SYNC 3 /* attention only: the round-1 mma.sync / ldmatrix kernels (kept as the A/B baseline) */

const char* valor_last_error(void);
int valor_version(void);
int valor_num_sms(void);

/* ---- GEMM + fused epilogue -----------------------------------------------------------------
 * Replaces every nn.Linear / torch.matmul(addmm) on the path: Swin qkv/proj/fc1/fc2/reduction
 * (videoswin.py:62-64,129-131,251), AST linears (transformer.py:109,136-137), BERT
 * query/key/value/dense (bert.py:233-235,303-305,347,360,399,412), heads (modeling.py:237,240;
 * pretrain.py:36,104-112) and their autograd gradients (dgrad / wgrad forms).
 *   C[M,N] (+)= epi( alpha * A[M,K] . B[N,K]^T )
 *   a_kmajor=1: A stored [M,K] row-major (pitch lda);  0: stored [K,M] row-major (pitch lda)
 *   b_kmajor=1: B stored [N,K] row-major (pitch ldb);  0: stored [K,N] row-major (pitch ldb)
 *   epi(x) = act(x + bias) [* act'(act_aux) instead of act when act_aux != NULL] + residual
 * force_bn / force_splits = 0: the library picks the tile width (64 / 128 / 192 / 256), the one- or two-CTA form
 * (256-wide tiles over an SM pair, tcgen05.mma.cta_group::2) and the split-K factor.  Non-zero values pin them for
 * tests and measurements: force_bn = width, + 1000 to require the two-CTA form, + 2000 to forbid it.
 */
typedef struct ValorGemmEpilogue {
  const float* bias;    /* [N] or NULL */
  const void* residual; /* [M,N] pitch ldr, dtype res_dtype, or NULL */
  const void* act_aux;  /* [M,N] pitch ld_aux: pre-activation saved by the forward, or NULL */
  void* preact_out;     /* [M,N] pitch ld_pre: receives x + bias (before act), or NULL */
  long long ldr, ld_aux, ld_pre;
  int res_dtype, aux_dtype;
  int act;
  int out_dtype;
  int accumulate; /* C += (fp32 C only); required for split-K */
  float alpha;
  float* bias_grad; /* NULL, or (a_kmajor = b_kmajor = 0, accumulate = 1: the weight-gradient form dW += dy^T x):
                       bias_grad[m] += alpha * sum_k A[k,m], i.e. the bias gradient torch autograd produces for nn.Linear,
                       fused into the same launch (no second pass over dy) */
  const float* row_scale; /* NULL, or one fp32 factor per group of `rows_per_group` consecutive output rows:
                             C = residual + row_scale[row / rows_per_group] * (alpha * A.B^T + bias)  (act = none).
                             DropPath's per-sample keep mask / keep_prob (videoswin.py:40-55,238,243) applied inside the
                             projection / fc2 GEMM that ends the residual branch */
  int rows_per_group;
} ValorGemmEpilogue;

int valor_gemm(int dtype, const void* A, long long lda, int a_kmajor, const void* B, long long ldb, int b_kmajor,
               void* C, long long ldc, int M, int N, int K, const ValorGemmEpilogue* ep, int backend, int force_bn,
               int force_splits, void* stream);

/* ---- LayerNorm: apex FusedLayerNorm (apex/apex/normalization/fused_layer_norm.py:129-161,
 * apex/csrc/layer_norm_cuda_kernel.cu) and nn.LayerNorm (videoswin.py:181,187,252,439) ------- */
int valor_layernorm_fwd(int dtype, const void* x, const float* gamma, const float* beta, void* y, float* mean,
                        float* rstd, long long M, int N, float eps, void* stream);
/* dres (optional): gradient that reaches x through a residual branch bypassing the LN; dx = LN'(dy) + dres */
int valor_layernorm_bwd(int dtype, const void* dy, const void* x, const float* gamma, const float* mean,
                        const float* rstd, const void* dres, void* dx, float* dgamma, float* dbeta, long long M,
                        int N, void* stream);

/* ---- F.normalize(dim=-1) on contrastive features (pretrain.py:276,283,289) -------------------- */
int valor_l2norm_fwd(int dtype, const void* x, void* y, float* nrm, long long M, int N, void* stream);
int valor_l2norm_bwd(int dtype, const void* dy, const void* x, const float* nrm, void* dx, long long M, int N,
                     void* stream);

/* ---- multi-head attention: BertSelfAttention / BertCrossAttention (bert.py:244-340) and AST
 * MultiHeadAttention (transformer.py:115-130).  P problems (sequences) x H heads; problem p has
 * Nq query rows starting at q_row0[p] (NULL: p*Nq) of Q and kv_len[p] (NULL: max_nk) key rows
 * starting at kv_row0[p] (NULL: p*max_nk) of K/V; head h occupies columns [h*hd,(h+1)*hd).
 * scores = q.k^T * scale + mask, mask = -10000 where key_valid[p,j]==0 or (causal[p] && j>i)
 * (bert.py:869-885).  q_key_range (NULL or int32 [Nq][2]): query i only sees keys lo <= j < hi of its problem; the
 * others do not exist for it (excluded from the softmax, not -10000).  This is how the three caption passes of one
 * sample (tva / tv / ta, pretrain.py:455-479: same text rows, video+audio / video / audio media tokens as
 * cross-attention source, bert.py:450) run as ONE problem over the sample's media tokens.
 * drop_p > 0: attention-probability dropout (bert.py:283,334; transformer.py:128): P.V uses keep * P / (1-p) with
 * keep(i,j) regenerated in the backward from rng_state = {seed, step offset} (int64[2], device), `site`, (problem, head) and
 * the (query, key) index; the log-sum-exp stays that of the undropped softmax.  Tensor-core path only.
 * lse [P,H,Nq] fp32 is saved for the backward.
 * Backward: dQ written.  dK/dV either accumulate (+=) into fp32 [rows, H*hd] buffers the caller zero-fills
 * (dK/dV: K/V rows shared by several problems add up — cross-attention), or, when every K/V row belongs to
 * exactly one problem (self-attention), are written directly in the compute dtype to dK_lp/dV_lp (pitch
 * lddkv_lp); exactly one of the two output pairs is non-NULL.  `delta` is a [P,H,Nq] fp32 scratch
 * (rowsum(dO*O), produced by the dQ kernel and consumed by the dK/dV kernel). */
int valor_mha_fwd(int dtype, const void* Q, const void* K, const void* V, long long ldq, long long ldk, long long ldv,
                  void* O, long long ldo, float* lse, int P, int H, int hd, int Nq, int max_nk, const int* q_row0,
                  const int* kv_row0, const int* kv_len, const unsigned char* key_valid, const unsigned char* causal,
                  const int* q_key_range, float scale, float drop_p, const long long* rng_state, long long site, int backend,
                  void* stream);
int valor_mha_bwd(int dtype, const void* Q, const void* K, const void* V, const void* O, const void* dO,
                  long long ldq, long long ldk, long long ldv, long long ldo, const float* lse, float* delta,
                  void* dQ, long long lddq, float* dK, float* dV, long long lddk, long long lddv, void* dK_lp,
                  void* dV_lp, long long lddkv_lp, int P, int H, int hd, int Nq, int max_nk, const int* q_row0,
                  const int* kv_row0, const int* kv_len, const unsigned char* key_valid,
                  const unsigned char* causal, const int* q_key_range, float scale, float drop_p, const long long* rng_state,
                  long long site, int backend, void* stream);

/* ---- VideoSwin shifted-window attention: WindowAttention3D.forward (videoswin.py:137-163)
 * together with torch.roll / window_partition / window_reverse / compute_mask
 * (videoswin.py:75-84,191-226,272-285), evaluated in place on the natural [B,D,H,W] token
 * order.  qkv [B*D*H*W, 3*heads*hd] (q | k | v), O [B*D*H*W, heads*hd].  (wd,wh,ww)/(sd,sh,sw)
 * are the EFFECTIVE window / shift (get_window_size, videoswin.py:86-99); (WD,WH,WW) the
 * configured window that sizes `table` = relative_position_bias_table [(2WD-1)(2WH-1)(2WW-1), heads].
 * Backward: dqkv [tokens, 3*heads*hd] receives dq | dk | dv; dtable accumulates (+=).  The SIMT
 * path needs an fp32 scratch of valor_window_attn_bwd_scratch_bytes() bytes, zero-filled by the
 * caller (the tensor-core path returns 0 and takes NULL). */
int valor_window_attn_fwd(int dtype, const void* qkv, long long ld, void* O, long long ldo, float* lse,
                          const float* table, int B, int D, int H, int W, int wd, int wh, int ww, int sd, int sh,
                          int sw, int WD, int WH, int WW, int heads, int hd, float scale, int backend, void* stream);
long long valor_window_attn_bwd_scratch_bytes(int dtype, long long tokens, int heads, int hd, long long ld, int backend);
int valor_window_attn_bwd(int dtype, const void* qkv, long long ld, const void* O, const void* dO, long long ldo,
                          const float* lse, float* delta, const float* table, void* dqkv, long long lddqkv,
                          float* scratch, float* dtable, int B, int D, int H, int W, int wd, int wh, int ww, int sd, int sh, int sw,
                          int WD, int WH, int WW, int heads, int hd, float scale, int backend, void* stream);

/* ---- data movement around the GEMMs --------------------------------------------------------- */
/* PatchEmbed3D as GEMM (videoswin.py:361-369): video [B,F,3,Hh,Ww] -> cols [B*F*Hh/4*Ww/4, 96] */
int valor_swin_im2col(int in_dtype, int dtype, const void* video, void* cols, int B, int F, int Hh, int Ww, void* stream);
/* AudioEmbeddings conv as GEMM (modeling.py:752-754): spec [BA,mel,frames] -> cols [BA*P, ps*ps] */
int valor_audio_im2col(int in_dtype, int dtype, const void* spec, void* cols, int BA, int mel, int frames, int ps, void* stream);
/* cls + position embeddings (modeling.py:755-760) */
int valor_ast_assemble_fwd(int dtype, const void* tok, const float* cls, const float* pos, void* x, int BA, int P, int Hd, void* stream);
int valor_ast_assemble_bwd(int dtype, const void* dx, void* dtok, float* dcls, float* dpos, int BA, int P, int Hd, void* stream);
/* BertEmbeddings lookup (bert.py:203-215) */
int valor_bert_embed_fwd(int dtype, const long long* tokens, const float* word, const float* pos, const float* type0, void* e, long long R, int Tn, int Hd, void* stream);
int valor_bert_embed_bwd(int dtype, const void* de, const long long* tokens, float* dword, float* dpos, float* dtype0, long long R, int Tn, int Hd, void* stream);
/* get_multimodal_forward_input_{video,audio} (modeling.py:485-502) into the cross-attention source */
int valor_media_input_fwd(int dtype, const void* in, const float* frame_emb, const float* type_emb, void* out, int B, int nf, int X, int Hd, int S_total, int row0, void* stream);
int valor_media_input_bwd(int dtype, const void* dout, void* din, float* dframe, float* dtype_emb, int B, int nf, int X, int Hd, int S_total, int row0, void* stream);
/* PatchMerging 2x2 gather (videoswin.py:261-265); inverse=1 is its gradient */
int valor_patch_merge(int dtype, const void* src, void* dst, long long BD, int H, int W, int C, int inverse, void* stream);
/* mean over the 49 spatial tokens per frame (modeling.py:389) */
int valor_mean_pool_fwd(int dtype, const void* x, void* y, long long R, int X, int C, void* stream);
int valor_mean_pool_bwd(int dtype, const void* dy, void* dx, long long R, int X, int C, void* stream);
/* bias gradient db[n] += sum_m dy[m,n] */
int valor_colsum(int dtype, const void* dy, long long ld, float* db, long long M, int N, void* stream);
/* dst[r,c] = cast(src[r,c]) over [R,C] with row pitches (flat: R=1) */
int valor_cast2d(int src_dtype, int dst_dtype, const void* src, long long sld, void* dst, long long dld, long long R, long long C, void* stream);
/* x [R,C] fp32 (pitch xld) -> out [R,3C] bf16: two-term bf16 expansion x ~ hi + lo arranged [hi|hi|lo] (side 0) or
 * [hi|lo|hi] (side 1), so one tcgen05 GEMM over K = 3C gives fp32-grade dot products of the L2-normalised contrastive
 * features: torch.einsum('atd,bvd->abtv') of compute_fine_matrix_slice (pretrain.py:200) in the reference's fp32 */
int valor_split_bf16x3(const float* x, long long xld, void* out, long long R, long long C, int side, void* stream);
/* ---- training-mode regularisation ---------------------------------------------------------------------------------
 * nn.Dropout(p) of BERT / AST hidden states (bert.py:217,353,367,418; transformer.py:78,83; modeling.py:761) fused with the
 * residual add that follows it:  out = residual + x * keep / (1-p)  (residual NULL: plain dropout; the backward of the
 * dropped branch is the same call on dy with residual NULL).  The keep mask is regenerated from Philox4x32-10 keyed by
 * rng_state = {seed, step offset} (int64[2] in DEVICE memory, rewritten by the host once per step so captured graphs see new
 * masks), the call-site id `site` and the element index r*C + c: nothing is stored between forward and backward. */
int valor_dropout(int dtype, const void* x, long long ldx, const void* residual, long long ldr, void* out, long long ldo, long long R,
                  int C, float p, const long long* rng_state, long long site, void* stream);
/* DropPath (videoswin.py:40-55): scale[b] = floor(1-p + u_b) / (1-p), one uniform per sample */
int valor_droppath_scale(float* scale, int B, float p, const long long* rng_state, long long site, void* stream);
/* out[r,:] = residual[r,:] + x[r,:] * scale[r / rows_per_group]   (videoswin.py:238,243; residual NULL in the backward) */
int valor_row_scale(int dtype, const void* x, long long ldx, const float* scale, long long rows_per_group, const void* residual,
                    long long ldr, void* out, long long ldo, long long R, int C, void* stream);
/* dh = dy * act'(h): gradient through GELU / QuickGELU / ReLU where it cannot ride a GEMM epilogue */
int valor_act_bwd(int dtype, const void* dy, const void* h, void* dh, long long n, int act, void* stream);
int valor_strided_rows(int dtype, const void* src, long long sld, void* dst, long long dld, long long R, int C, int accumulate, void* stream);

/* ---- losses ------------------------------------------------------------------------------------ */
/* F.cross_entropy over the MLM head (pretrain.py:441-444); labels -1 ignored; acc = {sum, count} */
int valor_xent_fwd(int dtype, const void* logits, long long ld, const long long* labels, float* lse, float* acc, float* loss, long long M, int V, void* stream);
int valor_xent_bwd(int dtype, const void* logits, long long ld, const long long* labels, const float* lse, const float* acc, const float* gptr, float gmul, void* dlogits, long long ldd, long long M, int V, void* stream);
/* weight softmax with -inf fill (pretrain.py:193-197) */
int valor_masked_softmax_fwd(const float* w, const unsigned char* mask, float* ws, int R, int L, void* stream);
int valor_masked_softmax_bwd(const float* ws, const float* dws, float* dw, int R, int L, void* stream);
/* compute_fine_matrix_slice reductions (pretrain.py:200-209) over L = featA.featB^T [Na*T, Nb*Vt] */
int valor_fine_reduce_fwd(const float* L, long long ldl, const unsigned char* mA, const float* wsA, const float* wsB, float* score, unsigned char* arg_v, unsigned char* arg_t, int Na, int Nb, int T, int Vt, int v0, int nv, void* stream);
int valor_fine_reduce_bwd(const float* L, long long ldl, const unsigned char* mA, const float* wsA, const float* wsB, const float* dscore, const unsigned char* arg_v, const unsigned char* arg_t, float* dL, float* dwsA, float* dwsB, int Na, int Nb, int T, int Vt, int v0, int nv, void* stream);
/* VALORModel.contrastive_loss (modeling.py:418-433) */
int valor_contrastive_fwd(const float* S, const float* temp, float* row_lse, float* col_lse, float* loss, int N, void* stream);
int valor_contrastive_bwd(const float* S, const float* temp, const float* row_lse, const float* col_lse, const float* gptr, float gmul, float* dS, float* dtemp, int N, void* stream);

/* ---- retrieval evaluation (test.py:680-775): rank of the ground-truth candidate of every query = number of candidates with a
 * higher score (replaces sort + host list.index); strides select the direction (text->video: rows, video->text: columns).
 * valor_dual_softmax: S * softmax(S / temp, over the `norm` direction) * Nnorm (compute_dualsoftmax_forward/backward). */
int valor_retrieval_rank(const float* S, long long stride_query, long long stride_cand, const int* gt, int* rank, int Nq, int Nc, void* stream);
int valor_dual_softmax(const float* S, float* out, long long stride_norm, long long stride_other, const float* temp, int Nnorm, int Nother, void* stream);

/* ---- optimizer step: optim/adamw.py:50-101 + clip_grad_norm_ (train_utils.py:359) ------------- */
int valor_grad_sumsq(const float* g, long long n, float* out, void* stream);
int valor_clip_coef(const float* sumsq, float max_norm, float* norm_out, void* stream);
/* hyper = {lr, beta1, beta2, eps, weight_decay, step_size} on device; coef = clip coefficient or NULL */
int valor_adamw(float* p, const float* g, float* m, float* v, void* p_lp, long long n, const float* hyper, const float* coef, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VALOR_B200_H */
