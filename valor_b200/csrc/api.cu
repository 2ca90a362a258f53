// valor_b200 — extern "C" boundary (include/valor_b200.h).  Plain pointers and sizes only.
#include "common.cuh"
#include "attention.cuh"
#include "../../include/valor_b200.h"
#include <stdarg.h>

namespace valor {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

// implemented in the other translation units
bool gemm_sm100_eligible(const void* A, const void* B, long long lda, long long ldb, int M, int N, int K);
int gemm_sm100(const void* A, long long lda, int a_kmajor, const void* B, long long ldb, int b_kmajor, void* C,
               long long ldc, int M, int N, int K, const GemmEpilogue& ep, int force_bn, int force_splits, cudaStream_t st);
bool gemm_sm100_fuses_bias_grad(const void* C, long long ldc, const GemmEpilogue& ep, int a_kmajor, int b_kmajor, int M, int N,
                                int force_bn);
int gemm_simt(int dtype, const void* A, long long sam, long long sak, const void* B, long long sbn, long long sbk,
              void* C, long long ldc, int M, int N, int K, const GemmEpilogue& ep, cudaStream_t st);
int layernorm_fwd(int, const void*, const float*, const float*, void*, float*, float*, long long, int, float, cudaStream_t);
int layernorm_bwd(int, const void*, const void*, const float*, const float*, const float*, const void*, void*, float*, float*, long long, int, cudaStream_t);
int l2norm_fwd(int, const void*, void*, float*, long long, int, cudaStream_t);
int l2norm_bwd(int, const void*, const void*, const float*, void*, long long, int, cudaStream_t);
int swin_im2col(int, int, const void*, void*, int, int, int, int, cudaStream_t);
int audio_im2col(int, int, const void*, void*, int, int, int, int, cudaStream_t);
int ast_assemble_fwd(int, const void*, const float*, const float*, void*, int, int, int, cudaStream_t);
int ast_assemble_bwd(int, const void*, void*, float*, float*, int, int, int, cudaStream_t);
int bert_embed_fwd(int, const long long*, const float*, const float*, const float*, void*, long long, int, int, cudaStream_t);
int bert_embed_bwd(int, const void*, const long long*, float*, float*, float*, long long, int, int, cudaStream_t);
int media_input_fwd(int, const void*, const float*, const float*, void*, int, int, int, int, int, int, cudaStream_t);
int media_input_bwd(int, const void*, void*, float*, float*, int, int, int, int, int, int, cudaStream_t);
int patch_merge(int, const void*, void*, long long, int, int, int, int, cudaStream_t);
int mean_pool_fwd(int, const void*, void*, long long, int, int, cudaStream_t);
int mean_pool_bwd(int, const void*, void*, long long, int, int, cudaStream_t);
int colsum(int, const void*, long long, float*, long long, int, cudaStream_t);
int cast2d(int, int, const void*, long long, void*, long long, long long, long long, cudaStream_t);
int act_bwd(int, const void*, const void*, void*, long long, int, cudaStream_t);
int strided_rows(int, const void*, long long, void*, long long, long long, int, int, cudaStream_t);
int xent_fwd(int, const void*, long long, const long long*, float*, float*, float*, long long, int, cudaStream_t);
int xent_bwd(int, const void*, long long, const long long*, const float*, const float*, const float*, float, void*, long long, long long, int, cudaStream_t);
int masked_softmax_fwd(const float*, const unsigned char*, float*, int, int, cudaStream_t);
int masked_softmax_bwd(const float*, const float*, float*, int, int, cudaStream_t);
int fine_reduce_fwd(const float*, long long, const unsigned char*, const float*, const float*, float*, unsigned char*, unsigned char*, int, int, int, int, int, int, cudaStream_t);
int fine_reduce_bwd(const float*, long long, const unsigned char*, const float*, const float*, const float*, const unsigned char*, const unsigned char*, float*, float*, float*, int, int, int, int, int, int, cudaStream_t);
int contrastive_fwd(const float*, const float*, float*, float*, float*, int, cudaStream_t);
int contrastive_bwd(const float*, const float*, const float*, const float*, const float*, float, float*, float*, int, cudaStream_t);
int grad_sumsq(const float*, long long, float*, cudaStream_t);
int clip_coef(const float*, float, float*, cudaStream_t);
int adamw(float*, const float*, float*, float*, void*, long long, const float*, const float*, cudaStream_t);
int split_bf16x3(const float* x, long long xld, void* out, long long R, long long C, int side, cudaStream_t st);
int dropout_apply(int, const void*, long long, const void*, long long, void*, long long, long long, int, float, const long long*, long long, cudaStream_t);
int droppath_scale(float*, int, float, const long long*, long long, cudaStream_t);
int row_scale(int, const void*, long long, const float*, long long, const void*, long long, void*, long long, long long, int, cudaStream_t);
int retrieval_rank(const float*, long long, long long, const int*, int*, int, int, cudaStream_t);
int dual_softmax(const float*, float*, long long, long long, const float*, int, int, cudaStream_t);
bool attn_mma_eligible(int dtype, int hd, long long ldq, long long ldk, long long ldv, long long ldo, const void* q,
                       const void* k, const void* v, const void* o);
int mha_mma_fwd(const MhaIndex& ix, const void* Q, const void* K, const void* V, long long ldq, long long ldk,
                long long ldv, void* O, long long ldo, float* lse, int Pn, int H, int hd, int Nq, float scale, float drop_p,
                const long long* rng, long long site, cudaStream_t st);
int mha_mma_bwd(const MhaIndex& ix, const void* Q, const void* K, const void* V, const void* O, const void* dO,
                long long ldq, long long ldk, long long ldv, long long ldo, const float* lse, float* delta, void* dQ,
                long long lddq, float* dK, float* dV, long long lddk, long long lddv, void* dK_lp, void* dV_lp,
                long long lddkv_lp, int Pn, int H, int hd, int Nq, float scale, float drop_p, const long long* rng, long long site,
                cudaStream_t st);
int window_mma_fwd(const WindowIndex& ix, const void* qkv, long long ld, void* O, long long ldo, float* lse, int Pn,
                   int H, int hd, float scale, cudaStream_t st);
int window_mma_bwd(const WindowIndex& ix, const void* qkv, long long ld, const void* O, const void* dO, long long ldo,
                   const float* lse, float* delta, void* dqkv, long long lddq, float* dtable, int Pn, int H, int hd,
                   float scale, cudaStream_t st);

bool window_cta_eligible(const WindowIndex& ix, int hd);
int window_cta_fwd(const WindowIndex& ix, const void* qkv, long long ld, void* O, long long ldo, float* lse, int Pn,
                   int H, int hd, float scale, cudaStream_t st);
int window_cta_bwd(const WindowIndex& ix, const void* qkv, long long ld, const void* O, const void* dO, long long ldo,
                   const float* lse, void* dqkv, long long lddqkv, float* dtable, int Pn, int H, int hd, float scale,
                   cudaStream_t st);

bool window_sm100_fwd_eligible(const WindowIndex& ix, int hd);
bool window_sm100_bwd_eligible(const WindowIndex& ix, int hd, bool want_dtab);
int window_sm100_bwd(const WindowIndex& ix, const void* qkv, long long ld, const void* O, const void* dO, long long ldo,
                     const float* lse, void* dqkv, long long lddqkv, float* dtable, int Pn, int H, int hd, float scale,
                     cudaStream_t st);
int window_sm100_fwd(const WindowIndex& ix, const void* qkv, long long ld, void* O, long long ldo, float* lse, int Pn,
                     int H, int hd, float scale, cudaStream_t st);

}  // namespace valor

using namespace valor;
static bool window_use_cta(const WindowIndex& ix, int hd) {
#ifdef VALOR_DEBUG   // -DVALOR_DEBUG only: VALOR_WINDOW_FLASH=1 keeps window attention on the key-blocked flash kernels (A/B)
  static int flash = -1;
  if (flash < 0) { const char* e = getenv("VALOR_WINDOW_FLASH"); flash = e ? atoi(e) : 0; }
  if (flash) return false;
#endif
  return window_cta_eligible(ix, hd);
}
#define ST ((cudaStream_t)stream)

extern "C" {

const char* valor_last_error(void) { return g_err; }
int valor_version(void) { return 100; }
int valor_num_sms(void) { return num_sms(); }

int valor_gemm(int dtype, const void* A, long long lda, int a_kmajor, const void* B, long long ldb, int b_kmajor,
               void* C, long long ldc, int M, int N, int K, const ValorGemmEpilogue* epc, int backend, int force_bn,
               int force_splits, void* stream) {
  VALOR_REQUIRE(epc != nullptr, "valor_gemm: epilogue is NULL");
  VALOR_REQUIRE(M >= 0 && N > 0 && K > 0, "valor_gemm: bad shape %d x %d x %d", M, N, K);
  if (M == 0) return 0;
  GemmEpilogue ep;
  ep.bias = epc->bias; ep.residual = epc->residual; ep.act_aux = epc->act_aux; ep.preact_out = epc->preact_out;
  ep.ldr = epc->ldr; ep.ld_aux = epc->ld_aux; ep.ld_pre = epc->ld_pre;
  ep.res_dtype = epc->res_dtype; ep.aux_dtype = epc->aux_dtype; ep.act = epc->act; ep.out_dtype = epc->out_dtype;
  ep.accumulate = epc->accumulate; ep.alpha = epc->alpha; ep.bias_grad = epc->bias_grad;
  ep.row_scale = epc->row_scale; ep.rows_per_group = epc->rows_per_group;
  VALOR_REQUIRE(ep.row_scale == nullptr || (ep.rows_per_group >= 1 && ep.act == VALOR_ACT_NONE && ep.act_aux == nullptr &&
                                             ep.preact_out == nullptr && !ep.accumulate),
                "valor_gemm: row_scale needs rows_per_group >= 1, no activation / side outputs / accumulation");
  VALOR_REQUIRE(ep.bias_grad == nullptr || (!a_kmajor && !b_kmajor && ep.accumulate), "valor_gemm: bias_grad needs the weight-gradient form (a_kmajor = b_kmajor = 0, accumulate = 1)");
  bool tensor_ok = dtype == VALOR_DT_BF16 && gemm_sm100_eligible(A, B, lda, ldb, M, N, K) &&
                   (ep.residual == nullptr || ep.res_dtype == VALOR_DT_BF16) &&
                   (ep.act_aux == nullptr || ep.aux_dtype == VALOR_DT_BF16) &&
                   (ep.preact_out == nullptr || ep.out_dtype == VALOR_DT_BF16) &&
                   (!ep.accumulate || ep.out_dtype == VALOR_DT_F32);
  if (backend == VALOR_BACKEND_TENSOR)
    VALOR_REQUIRE(tensor_ok, "valor_gemm: tensor backend requested but operands are not eligible");
  const bool use_tensor = backend == VALOR_BACKEND_TENSOR || (backend == VALOR_BACKEND_AUTO && tensor_ok);
  if (ep.bias_grad != nullptr && !(use_tensor && gemm_sm100_fuses_bias_grad(C, ldc, ep, a_kmajor, b_kmajor, M, N, force_bn))) {
    // not fusable here (fp32 operands / odd pitches): the bias gradient is its own column-sum launch over A = dY [K, M]
    VALOR_REQUIRE(ep.alpha == 1.f, "valor_gemm: unfused bias_grad needs alpha = 1");
    if (colsum(dtype, A, lda, ep.bias_grad, K, M, ST)) return 1;
    ep.bias_grad = nullptr;
  }
  if (use_tensor)
    return gemm_sm100(A, lda, a_kmajor, B, ldb, b_kmajor, C, ldc, M, N, K, ep, force_bn, force_splits, ST);
  return gemm_simt(dtype, A, a_kmajor ? lda : 1, a_kmajor ? 1 : lda, B, b_kmajor ? ldb : 1, b_kmajor ? 1 : ldb, C, ldc,
                   M, N, K, ep, ST);
}

int valor_layernorm_fwd(int dtype, const void* x, const float* gamma, const float* beta, void* y, float* mean,
                        float* rstd, long long M, int N, float eps, void* stream) {
  return layernorm_fwd(dtype, x, gamma, beta, y, mean, rstd, M, N, eps, ST);
}
int valor_layernorm_bwd(int dtype, const void* dy, const void* x, const float* gamma, const float* mean,
                        const float* rstd, const void* dres, void* dx, float* dgamma, float* dbeta, long long M, int N,
                        void* stream) {
  return layernorm_bwd(dtype, dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, M, N, ST);
}
int valor_l2norm_fwd(int dtype, const void* x, void* y, float* nrm, long long M, int N, void* stream) {
  return l2norm_fwd(dtype, x, y, nrm, M, N, ST);
}
int valor_l2norm_bwd(int dtype, const void* dy, const void* x, const float* nrm, void* dx, long long M, int N, void* stream) {
  return l2norm_bwd(dtype, dy, x, nrm, dx, M, N, ST);
}

static MhaIndex make_mha(int Nq, int max_nk, const int* q_row0, const int* kv_row0, const int* kv_len,
                         const unsigned char* key_valid, const unsigned char* causal) {
  MhaIndex ix;
  ix.max_nk = max_nk; ix.Nq = Nq; ix.q_row0 = q_row0; ix.kv_row0 = kv_row0; ix.kv_len = kv_len;
  ix.key_valid = key_valid; ix.causal = causal;
  return ix;
}

int valor_mha_fwd(int dtype, const void* Q, const void* K, const void* V, long long ldq, long long ldk, long long ldv,
                  void* O, long long ldo, float* lse, int P, int H, int hd, int Nq, int max_nk, const int* q_row0,
                  const int* kv_row0, const int* kv_len, const unsigned char* key_valid, const unsigned char* causal,
                  const int* q_key_range, float scale, float drop_p, const long long* rng_state, long long site, int backend,
                  void* stream) {
  if (P == 0) return 0;
  MhaIndex ix = make_mha(Nq, max_nk, q_row0, kv_row0, kv_len, key_valid, causal);
  ix.q_key_range = q_key_range;
  VALOR_REQUIRE(q_key_range == nullptr || max_nk < 65535, "valor_mha_fwd: per-query key ranges need max_nk < 65535");
  VALOR_REQUIRE(drop_p >= 0.f && drop_p < 1.f && (drop_p == 0.f || rng_state != nullptr), "valor_mha_fwd: bad dropout arguments");
  const bool ok = attn_mma_eligible(dtype, hd, ldq, ldk, ldv, ldo, Q, K, V, O);
  if (backend == VALOR_BACKEND_TENSOR) VALOR_REQUIRE(ok, "valor_mha_fwd: tensor backend requested but not eligible");
  if (backend != VALOR_BACKEND_SIMT && ok)
    return mha_mma_fwd(ix, Q, K, V, ldq, ldk, ldv, O, ldo, lse, P, H, hd, Nq, scale, drop_p, rng_state, site, ST);
  VALOR_REQUIRE(drop_p == 0.f, "valor_mha_fwd: attention dropout is a tensor-core-path feature (the fp32 parity path runs without)");
  return mha_ref_fwd(dtype, ix, Q, K, V, ldq, ldk, ldv, O, ldo, lse, P, H, hd, Nq, scale, ST);
}
int valor_mha_bwd(int dtype, const void* Q, const void* K, const void* V, const void* O, const void* dO,
                  long long ldq, long long ldk, long long ldv, long long ldo, const float* lse, float* delta,
                  void* dQ, long long lddq, float* dK, float* dV, long long lddk, long long lddv, void* dK_lp,
                  void* dV_lp, long long lddkv_lp, int P, int H, int hd, int Nq, int max_nk, const int* q_row0,
                  const int* kv_row0, const int* kv_len, const unsigned char* key_valid,
                  const unsigned char* causal, const int* q_key_range, float scale, float drop_p, const long long* rng_state,
                  long long site, int backend, void* stream) {
  if (P == 0) return 0;
  MhaIndex ix = make_mha(Nq, max_nk, q_row0, kv_row0, kv_len, key_valid, causal);
  ix.q_key_range = q_key_range;
  VALOR_REQUIRE(q_key_range == nullptr || max_nk < 65535, "valor_mha_bwd: per-query key ranges need max_nk < 65535");
  VALOR_REQUIRE(drop_p >= 0.f && drop_p < 1.f && (drop_p == 0.f || rng_state != nullptr), "valor_mha_bwd: bad dropout arguments");
  const bool ok = attn_mma_eligible(dtype, hd, ldq, ldk, ldv, ldo, Q, K, V, O) && (lddq % 8 == 0) &&
                  (((uintptr_t)dQ | (uintptr_t)dO) & 15) == 0 && delta != nullptr;
  if (backend == VALOR_BACKEND_TENSOR) VALOR_REQUIRE(ok, "valor_mha_bwd: tensor backend requested but not eligible");
  if (backend != VALOR_BACKEND_SIMT && ok)
    return mha_mma_bwd(ix, Q, K, V, O, dO, ldq, ldk, ldv, ldo, lse, delta, dQ, lddq, dK, dV, lddk, lddv, dK_lp, dV_lp,
                       lddkv_lp, P, H, hd, Nq, scale, drop_p, rng_state, site, ST);
  VALOR_REQUIRE(drop_p == 0.f, "valor_mha_bwd: attention dropout is a tensor-core-path feature");
  // row-per-warp path: fp32 accumulation only; in fp32 parity mode the "direct" outputs ARE fp32 buffers
  if (dK == nullptr && dtype == VALOR_DT_F32 && dK_lp != nullptr) {
    dK = (float*)dK_lp; dV = (float*)dV_lp; lddk = lddv = lddkv_lp;
  }
  VALOR_REQUIRE(dK != nullptr && dV != nullptr, "valor_mha_bwd: SIMT path needs fp32 dK/dV accumulators");
  return mha_ref_bwd(dtype, ix, Q, K, V, O, dO, ldq, ldk, ldv, ldo, lse, dQ, lddq, dK, dV, lddk, lddv, P, H, hd, Nq,
                     scale, ST);
}

static int make_window(WindowIndex& ix, const float* table, int B, int D, int H, int W, int wd, int wh, int ww, int sd,
                       int sh, int sw, int WD, int WH, int WW, int heads) {
  VALOR_REQUIRE(wd > 0 && wh > 0 && ww > 0 && D % wd == 0 && H % wh == 0 && W % ww == 0,
                "window_attn: grid %dx%dx%d is not a multiple of window %dx%dx%d (padded windows unsupported)", D, H, W,
                wd, wh, ww);
  VALOR_REQUIRE(wd <= WD && wh <= WH && ww <= WW, "window_attn: effective window exceeds configured window");
  VALOR_REQUIRE(sd >= 0 && sd < wd + (wd == 0) && sh >= 0 && sh < wh && sw >= 0 && sw < ww, "window_attn: bad shift");
  ix.N = wd * wh * ww; ix.max_nk = ix.N;
  ix.B = B; ix.D = D; ix.H = H; ix.W = W;
  ix.wd = wd; ix.wh = wh; ix.ww = ww; ix.sd = sd; ix.sh = sh; ix.sw = sw;
  ix.WD = WD; ix.WH = WH; ix.WW = WW; ix.heads = heads; ix.table = table;
  return 0;
}

int valor_window_attn_fwd(int dtype, const void* qkv, long long ld, void* O, long long ldo, float* lse,
                          const float* table, int B, int D, int H, int W, int wd, int wh, int ww, int sd, int sh,
                          int sw, int WD, int WH, int WW, int heads, int hd, float scale, int backend, void* stream) {
  WindowIndex ix;
  if (make_window(ix, table, B, D, H, W, wd, wh, ww, sd, sh, sw, WD, WH, WW, heads)) return 1;
  const int P = B * (D / wd) * (H / wh) * (W / ww);
  const char* qb = (const char*)qkv;
  const bool ok = attn_mma_eligible(dtype, hd, ld, ld, ld, ldo, qb, qb + 2 * heads * hd, qb + 4 * heads * hd, O);
  if (backend == VALOR_BACKEND_TENSOR) VALOR_REQUIRE(ok, "valor_window_attn_fwd: tensor backend requested but not eligible");
  // forward: the tcgen05 / TMEM kernel is taken on request (VALOR_BACKEND_TENSOR); AUTO keeps the round-1 mma.sync kernel,
  // which is still the faster of the two at these 392-token, head-dim-32 problems (profiles/window_attention_r2.md)
  if (backend == VALOR_BACKEND_TENSOR && ok && window_sm100_fwd_eligible(ix, hd))
    return window_sm100_fwd(ix, qkv, ld, O, ldo, lse, P, heads, hd, scale, ST);     // tcgen05 / TMEM
  if (backend != VALOR_BACKEND_SIMT && ok && window_use_cta(ix, hd))
    return window_cta_fwd(ix, qkv, ld, O, ldo, lse, P, heads, hd, scale, ST);
  if (backend != VALOR_BACKEND_SIMT && ok) return window_mma_fwd(ix, qkv, ld, O, ldo, lse, P, heads, hd, scale, ST);
  return window_ref_fwd(dtype, ix, qkv, ld, O, ldo, lse, P, heads, hd, scale, ST);
}
static bool window_bwd_tensor_ok(int dtype, int hd, long long ld, int backend) {
  return backend != VALOR_BACKEND_SIMT && dtype == VALOR_DT_BF16 && (hd == 32 || hd == 64) && (ld % 8 == 0);
}
long long valor_window_attn_bwd_scratch_bytes(int dtype, long long tokens, int heads, int hd, long long ld, int backend) {
  if (window_bwd_tensor_ok(dtype, hd, ld, backend)) return 0;
  return tokens * 2LL * heads * hd * (long long)sizeof(float);
}
int valor_window_attn_bwd(int dtype, const void* qkv, long long ld, const void* O, const void* dO, long long ldo,
                          const float* lse, float* delta, const float* table, void* dqkv, long long lddqkv,
                          float* scratch, float* dtable, int B, int D, int H, int W, int wd, int wh, int ww, int sd, int sh, int sw,
                          int WD, int WH, int WW, int heads, int hd, float scale, int backend, void* stream) {
  WindowIndex ix;
  if (make_window(ix, table, B, D, H, W, wd, wh, ww, sd, sh, sw, WD, WH, WW, heads)) return 1;
  const int P = B * (D / wd) * (H / wh) * (W / ww);
  const int C = heads * hd;
  const long long tokens = (long long)B * D * H * W;
  if (window_bwd_tensor_ok(dtype, hd, ld, backend) && backend != VALOR_BACKEND_MMA_SYNC && lddqkv % 8 == 0 && ldo % 8 == 0 &&
      ((((uintptr_t)qkv | (uintptr_t)O | (uintptr_t)dO | (uintptr_t)dqkv) & 15) == 0) &&
      window_sm100_bwd_eligible(ix, hd, dtable != nullptr))
    return window_sm100_bwd(ix, qkv, ld, O, dO, ldo, lse, dqkv, lddqkv, dtable, P, heads, hd, scale, ST);   // tcgen05 / TMEM
  if (window_bwd_tensor_ok(dtype, hd, ld, backend) && lddqkv % 8 == 0 && ldo % 8 == 0 &&
      ((((uintptr_t)qkv | (uintptr_t)O | (uintptr_t)dO | (uintptr_t)dqkv) & 15) == 0) && window_use_cta(ix, hd))
    return window_cta_bwd(ix, qkv, ld, O, dO, ldo, lse, dqkv, lddqkv, dtable, P, heads, hd, scale, ST);
  if (window_bwd_tensor_ok(dtype, hd, ld, backend) && lddqkv % 8 == 0 && ldo % 8 == 0 &&
      ((((uintptr_t)qkv | (uintptr_t)O | (uintptr_t)dO | (uintptr_t)dqkv) & 15) == 0) && delta != nullptr)
    return window_mma_bwd(ix, qkv, ld, O, dO, ldo, lse, delta, dqkv, lddqkv, dtable, P, heads, hd, scale, ST);
  VALOR_REQUIRE(scratch != nullptr, "valor_window_attn_bwd: SIMT path needs the fp32 scratch buffer");
  if (window_ref_bwd(dtype, ix, qkv, ld, O, dO, ldo, lse, dqkv, lddqkv, scratch, scratch + C, 2 * C, dtable, P, heads,
                     hd, scale, ST))
    return 1;
  const size_t es = dtype == VALOR_DT_F32 ? 4 : 2;
  return cast2d(VALOR_DT_F32, dtype, scratch, 2 * C, (char*)dqkv + C * es, lddqkv, tokens, 2 * C, ST);
}

int valor_swin_im2col(int in_dtype, int dtype, const void* video, void* cols, int B, int F, int Hh, int Ww, void* stream) {
  return swin_im2col(in_dtype, dtype, video, cols, B, F, Hh, Ww, ST);
}
int valor_audio_im2col(int in_dtype, int dtype, const void* spec, void* cols, int BA, int mel, int frames, int ps, void* stream) {
  return audio_im2col(in_dtype, dtype, spec, cols, BA, mel, frames, ps, ST);
}
int valor_ast_assemble_fwd(int dtype, const void* tok, const float* cls, const float* pos, void* x, int BA, int P, int Hd, void* stream) {
  return ast_assemble_fwd(dtype, tok, cls, pos, x, BA, P, Hd, ST);
}
int valor_ast_assemble_bwd(int dtype, const void* dx, void* dtok, float* dcls, float* dpos, int BA, int P, int Hd, void* stream) {
  return ast_assemble_bwd(dtype, dx, dtok, dcls, dpos, BA, P, Hd, ST);
}
int valor_bert_embed_fwd(int dtype, const long long* tokens, const float* word, const float* pos, const float* type0, void* e, long long R, int Tn, int Hd, void* stream) {
  return bert_embed_fwd(dtype, tokens, word, pos, type0, e, R, Tn, Hd, ST);
}
int valor_bert_embed_bwd(int dtype, const void* de, const long long* tokens, float* dword, float* dpos, float* dtype0, long long R, int Tn, int Hd, void* stream) {
  return bert_embed_bwd(dtype, de, tokens, dword, dpos, dtype0, R, Tn, Hd, ST);
}
int valor_media_input_fwd(int dtype, const void* in, const float* frame_emb, const float* type_emb, void* out, int B, int nf, int X, int Hd, int S_total, int row0, void* stream) {
  return media_input_fwd(dtype, in, frame_emb, type_emb, out, B, nf, X, Hd, S_total, row0, ST);
}
int valor_media_input_bwd(int dtype, const void* dout, void* din, float* dframe, float* dtype_emb, int B, int nf, int X, int Hd, int S_total, int row0, void* stream) {
  return media_input_bwd(dtype, dout, din, dframe, dtype_emb, B, nf, X, Hd, S_total, row0, ST);
}
int valor_patch_merge(int dtype, const void* src, void* dst, long long BD, int H, int W, int C, int inverse, void* stream) {
  return patch_merge(dtype, src, dst, BD, H, W, C, inverse, ST);
}
int valor_mean_pool_fwd(int dtype, const void* x, void* y, long long R, int X, int C, void* stream) {
  return mean_pool_fwd(dtype, x, y, R, X, C, ST);
}
int valor_mean_pool_bwd(int dtype, const void* dy, void* dx, long long R, int X, int C, void* stream) {
  return mean_pool_bwd(dtype, dy, dx, R, X, C, ST);
}
int valor_colsum(int dtype, const void* dy, long long ld, float* db, long long M, int N, void* stream) {
  return colsum(dtype, dy, ld, db, M, N, ST);
}
int valor_cast2d(int src_dtype, int dst_dtype, const void* src, long long sld, void* dst, long long dld, long long R, long long C, void* stream) {
  return cast2d(src_dtype, dst_dtype, src, sld, dst, dld, R, C, ST);
}
int valor_split_bf16x3(const float* x, long long xld, void* out, long long R, long long C, int side, void* stream) {
  return split_bf16x3(x, xld, out, R, C, side, ST);
}
int valor_dropout(int dtype, const void* x, long long ldx, const void* residual, long long ldr, void* out, long long ldo, long long R,
                  int C, float p, const long long* rng_state, long long site, void* stream) {
  return dropout_apply(dtype, x, ldx, residual, ldr, out, ldo, R, C, p, rng_state, site, ST);
}
int valor_droppath_scale(float* scale, int B, float p, const long long* rng_state, long long site, void* stream) {
  return droppath_scale(scale, B, p, rng_state, site, ST);
}
int valor_row_scale(int dtype, const void* x, long long ldx, const float* scale, long long rows_per_group, const void* residual,
                    long long ldr, void* out, long long ldo, long long R, int C, void* stream) {
  return row_scale(dtype, x, ldx, scale, rows_per_group, residual, ldr, out, ldo, R, C, ST);
}
int valor_act_bwd(int dtype, const void* dy, const void* h, void* dh, long long n, int act, void* stream) {
  return act_bwd(dtype, dy, h, dh, n, act, ST);
}
int valor_strided_rows(int dtype, const void* src, long long sld, void* dst, long long dld, long long R, int C, int accumulate, void* stream) {
  return strided_rows(dtype, src, sld, dst, dld, R, C, accumulate, ST);
}
int valor_xent_fwd(int dtype, const void* logits, long long ld, const long long* labels, float* lse, float* acc, float* loss, long long M, int V, void* stream) {
  return xent_fwd(dtype, logits, ld, labels, lse, acc, loss, M, V, ST);
}
int valor_xent_bwd(int dtype, const void* logits, long long ld, const long long* labels, const float* lse, const float* acc, const float* gptr, float gmul, void* dlogits, long long ldd, long long M, int V, void* stream) {
  return xent_bwd(dtype, logits, ld, labels, lse, acc, gptr, gmul, dlogits, ldd, M, V, ST);
}
int valor_masked_softmax_fwd(const float* w, const unsigned char* mask, float* ws, int R, int L, void* stream) {
  return masked_softmax_fwd(w, mask, ws, R, L, ST);
}
int valor_masked_softmax_bwd(const float* ws, const float* dws, float* dw, int R, int L, void* stream) {
  return masked_softmax_bwd(ws, dws, dw, R, L, ST);
}
int valor_fine_reduce_fwd(const float* L, long long ldl, const unsigned char* mA, const float* wsA, const float* wsB, float* score, unsigned char* arg_v, unsigned char* arg_t, int Na, int Nb, int T, int Vt, int v0, int nv, void* stream) {
  return fine_reduce_fwd(L, ldl, mA, wsA, wsB, score, arg_v, arg_t, Na, Nb, T, Vt, v0, nv, ST);
}
int valor_fine_reduce_bwd(const float* L, long long ldl, const unsigned char* mA, const float* wsA, const float* wsB, const float* dscore, const unsigned char* arg_v, const unsigned char* arg_t, float* dL, float* dwsA, float* dwsB, int Na, int Nb, int T, int Vt, int v0, int nv, void* stream) {
  return fine_reduce_bwd(L, ldl, mA, wsA, wsB, dscore, arg_v, arg_t, dL, dwsA, dwsB, Na, Nb, T, Vt, v0, nv, ST);
}
int valor_contrastive_fwd(const float* S, const float* temp, float* row_lse, float* col_lse, float* loss, int N, void* stream) {
  return contrastive_fwd(S, temp, row_lse, col_lse, loss, N, ST);
}
int valor_contrastive_bwd(const float* S, const float* temp, const float* row_lse, const float* col_lse, const float* gptr, float gmul, float* dS, float* dtemp, int N, void* stream) {
  return contrastive_bwd(S, temp, row_lse, col_lse, gptr, gmul, dS, dtemp, N, ST);
}
int valor_retrieval_rank(const float* S, long long stride_query, long long stride_cand, const int* gt, int* rank, int Nq, int Nc, void* stream) {
  return retrieval_rank(S, stride_query, stride_cand, gt, rank, Nq, Nc, ST);
}
int valor_dual_softmax(const float* S, float* out, long long stride_norm, long long stride_other, const float* temp, int Nnorm, int Nother, void* stream) {
  return dual_softmax(S, out, stride_norm, stride_other, temp, Nnorm, Nother, ST);
}
int valor_grad_sumsq(const float* g, long long n, float* out, void* stream) { return grad_sumsq(g, n, out, ST); }
int valor_clip_coef(const float* sumsq, float max_norm, float* norm_out, void* stream) { return clip_coef(sumsq, max_norm, norm_out, ST); }
int valor_adamw(float* p, const float* g, float* m, float* v, void* p_lp, long long n, const float* hyper, const float* coef, void* stream) {
  return adamw(p, g, m, v, p_lp, n, hyper, coef, ST);
}

}  // extern "C"
