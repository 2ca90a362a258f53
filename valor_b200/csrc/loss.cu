// valor_b200 — loss-side kernels: masked-LM cross-entropy over the 30522-wide vocabulary
// (pretrain.py:441-444), fine-grained similarity reductions (pretrain.py:191-211) and the
// symmetric contrastive loss (modeling.py:418-433).
#include "common.cuh"

namespace valor {

// ---------------------------------------------------------------------------------------
// cross entropy: one CTA per row; acc[0] += sum of row losses, acc[1] += valid rows.
// labels == -1 are ignored (the reference gathers rows with labels != -1 before the head;
// here every position is evaluated and ignored rows contribute nothing — same mean).
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
xent_fwd_kernel(const T* __restrict__ logits, long long ld, const long long* __restrict__ labels,
                float* __restrict__ lse_out, float* __restrict__ acc, int V) {
  __shared__ float sh[32];
  const long long row = blockIdx.x;
  const long long lab = labels[row];
  if (lab < 0) {
    if (threadIdx.x == 0) lse_out[row] = 0.f;
    return;
  }
  const T* lr = logits + row * ld;
  float mx = -INFINITY;
  for (int v = threadIdx.x; v < V; v += blockDim.x) mx = fmaxf(mx, to_f(lr[v]));
  mx = block_max(mx, sh);
  float s = 0.f;
  for (int v = threadIdx.x; v < V; v += blockDim.x) s += __expf(to_f(lr[v]) - mx);
  s = block_sum(s, sh);
  if (threadIdx.x == 0) {
    const float lse = mx + __logf(s);
    lse_out[row] = lse;
    atomicAdd(&acc[0], lse - to_f(lr[lab]));
    atomicAdd(&acc[1], 1.0f);
  }
}
// dlogits = (softmax - onehot) * g / count   (in place allowed)
template <typename T>
__global__ void __launch_bounds__(256)
xent_bwd_kernel(const T* __restrict__ logits, long long ld, const long long* __restrict__ labels,
                const float* __restrict__ lse, const float* __restrict__ acc, const float* __restrict__ gptr,
                float gmul, T* __restrict__ dlogits, long long ldd, int V) {
  const long long row = blockIdx.x;
  const long long lab = labels[row];
  const float cnt = fmaxf(acc[1], 1.0f);
  const float g = (gptr ? gptr[0] : 1.0f) * gmul / cnt;
  const T* lr = logits + row * ld;
  T* dr = dlogits + row * ldd;
  if (lab < 0) {
    for (int v = threadIdx.x; v < V; v += blockDim.x) dr[v] = from_f<T>(0.f);
    return;
  }
  const float L = lse[row];
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
    float p = __expf(to_f(lr[v]) - L);
    if (v == lab) p -= 1.0f;
    dr[v] = from_f<T>(p * g);
  }
}
__global__ void ratio_kernel(const float* __restrict__ acc, float* __restrict__ out) { out[0] = acc[0] / fmaxf(acc[1], 1.0f); }

int xent_fwd(int dtype, const void* logits, long long ld, const long long* labels, float* lse, float* acc, float* loss,
             long long M, int V, cudaStream_t st) {
  VALOR_CUDA(cudaMemsetAsync(acc, 0, 2 * sizeof(float), st));
  if (M > 0) {
    if (dtype == VALOR_DT_F32) xent_fwd_kernel<float><<<(unsigned)M, 256, 0, st>>>((const float*)logits, ld, labels, lse, acc, V);
    else xent_fwd_kernel<bf16><<<(unsigned)M, 256, 0, st>>>((const bf16*)logits, ld, labels, lse, acc, V);
  }
  ratio_kernel<<<1, 1, 0, st>>>(acc, loss);
  return check_launch("xent_fwd_kernel");
}
int xent_bwd(int dtype, const void* logits, long long ld, const long long* labels, const float* lse, const float* acc,
             const float* gptr, float gmul, void* dlogits, long long ldd, long long M, int V, cudaStream_t st) {
  if (M == 0) return 0;
  if (dtype == VALOR_DT_F32)
    xent_bwd_kernel<float><<<(unsigned)M, 256, 0, st>>>((const float*)logits, ld, labels, lse, acc, gptr, gmul, (float*)dlogits, ldd, V);
  else
    xent_bwd_kernel<bf16><<<(unsigned)M, 256, 0, st>>>((const bf16*)logits, ld, labels, lse, acc, gptr, gmul, (bf16*)dlogits, ldd, V);
  return check_launch("xent_bwd_kernel");
}

// ---------------------------------------------------------------------------------------
// masked softmax over short rows: ws = softmax(w.masked_fill(mask==0, -inf))  (pretrain.py:193-197)
// ---------------------------------------------------------------------------------------
__global__ void masked_softmax_fwd_kernel(const float* __restrict__ w, const unsigned char* __restrict__ mask,
                                          float* __restrict__ ws, int R, int L) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  float mx = -INFINITY;
  for (int i = 0; i < L; ++i)
    if (!mask || mask[r * L + i]) mx = fmaxf(mx, w[r * L + i]);
  float s = 0.f;
  for (int i = 0; i < L; ++i)
    if (!mask || mask[r * L + i]) s += __expf(w[r * L + i] - mx);
  for (int i = 0; i < L; ++i) ws[r * L + i] = (!mask || mask[r * L + i]) ? __expf(w[r * L + i] - mx) / s : 0.f;
}
__global__ void masked_softmax_bwd_kernel(const float* __restrict__ ws, const float* __restrict__ dws,
                                          float* __restrict__ dw, int R, int L) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  float dot = 0.f;
  for (int i = 0; i < L; ++i) dot += ws[r * L + i] * dws[r * L + i];
  for (int i = 0; i < L; ++i) dw[r * L + i] = ws[r * L + i] * (dws[r * L + i] - dot);
}
int masked_softmax_fwd(const float* w, const unsigned char* mask, float* ws, int R, int L, cudaStream_t st) {
  if (R == 0) return 0;
  masked_softmax_fwd_kernel<<<(R + 127) / 128, 128, 0, st>>>(w, mask, ws, R, L);
  return check_launch("masked_softmax_fwd_kernel");
}
int masked_softmax_bwd(const float* ws, const float* dws, float* dw, int R, int L, cudaStream_t st) {
  if (R == 0) return 0;
  masked_softmax_bwd_kernel<<<(R + 127) / 128, 128, 0, st>>>(ws, dws, dw, R, L);
  return check_launch("masked_softmax_bwd_kernel");
}

// ---------------------------------------------------------------------------------------
// fine-grained similarity reduction.  L = featA . featB^T  as a [Na*T, Nb*Vt] fp32 matrix
// (produced by the GEMM); columns [v0, v0+nv) of every b select the modality group.
//   logit[a,b,t,v] = L * mA[a,t] * mB[b,v]            (masks multiply, pretrain.py:201-202)
//   score[a,b] = 0.5 * ( sum_t wsA[a,t] max_v logit + sum_v wsB[b,v] max_t logit )
// One warp per (a,b).  argmaxes are saved for the backward.
// ---------------------------------------------------------------------------------------
#define FINE_MAX_T 64
#define FINE_MAX_V 16
__global__ void __launch_bounds__(128)
fine_reduce_fwd_kernel(const float* __restrict__ L, long long ldl, const unsigned char* __restrict__ mA,
                       const float* __restrict__ wsA, const float* __restrict__ wsB, float* __restrict__ score,
                       unsigned char* __restrict__ arg_v, unsigned char* __restrict__ arg_t, int Na, int Nb, int T,
                       int Vt, int v0, int nv) {
  const int lane = threadIdx.x & 31;
  const long long pair = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (pair >= (long long)Na * Nb) return;
  const int a = (int)(pair / Nb), b = (int)(pair % Nb);
  // lane t (and t+32) owns query row t
  float colmax[FINE_MAX_V];
  int colarg[FINE_MAX_V];
#pragma unroll
  for (int v = 0; v < FINE_MAX_V; ++v) { colmax[v] = -INFINITY; colarg[v] = 0; }
  float a2b = 0.f;
  for (int t = lane; t < T; t += 32) {
    const float m = mA ? (float)mA[a * T + t] : 1.f;
    const float* lr = L + ((long long)a * T + t) * ldl + (long long)b * Vt + v0;
    float best = -INFINITY;
    int bi = 0;
#pragma unroll
    for (int v = 0; v < FINE_MAX_V; ++v) {
      if (v < nv) {
        const float x = lr[v] * m;
        if (x > best) { best = x; bi = v; }
        if (x > colmax[v]) { colmax[v] = x; colarg[v] = t; }
      }
    }
    a2b += wsA[a * T + t] * best;
    arg_v[((long long)a * Nb + b) * T + t] = (unsigned char)bi;
  }
  a2b = warp_sum(a2b);
  float b2a = 0.f;
#pragma unroll
  for (int v = 0; v < FINE_MAX_V; ++v) {
    if (v < nv) {
      float mx = colmax[v];
      int ar = colarg[v];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, mx, o);
        const int oa = __shfl_xor_sync(0xffffffffu, ar, o);
        if (om > mx || (om == mx && oa < ar)) { mx = om; ar = oa; }
      }
      b2a += wsB[b * nv + v] * mx;
      if (lane == 0) arg_t[((long long)a * Nb + b) * nv + v] = (unsigned char)ar;
    }
  }
  if (lane == 0) score[pair] = 0.5f * (a2b + b2a);
}

// backward: scatters d score into dL (zero-filled by the caller) and reduces d wsA / d wsB.
__global__ void __launch_bounds__(128)
fine_reduce_bwd_kernel(const float* __restrict__ L, long long ldl, const unsigned char* __restrict__ mA,
                       const float* __restrict__ wsA, const float* __restrict__ wsB, const float* __restrict__ dscore,
                       const unsigned char* __restrict__ arg_v, const unsigned char* __restrict__ arg_t,
                       float* __restrict__ dL, float* __restrict__ dwsA, float* __restrict__ dwsB, int Na, int Nb,
                       int T, int Vt, int v0, int nv) {
  const int lane = threadIdx.x & 31;
  const long long pair = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (pair >= (long long)Na * Nb) return;
  const int a = (int)(pair / Nb), b = (int)(pair % Nb);
  const float g = 0.5f * dscore[pair];
  for (int t = lane; t < T; t += 32) {
    const float m = mA ? (float)mA[a * T + t] : 1.f;
    const int bv = arg_v[((long long)a * Nb + b) * T + t];
    const long long o = ((long long)a * T + t) * ldl + (long long)b * Vt + v0 + bv;
    const float best = L[o] * m;
    atomicAdd(&dwsA[a * T + t], g * best);
    atomicAdd(&dL[o], g * wsA[a * T + t] * m);
  }
  for (int v = lane; v < nv; v += 32) {
    const int bt = arg_t[((long long)a * Nb + b) * nv + v];
    const float m = mA ? (float)mA[a * T + bt] : 1.f;
    const long long o = ((long long)a * T + bt) * ldl + (long long)b * Vt + v0 + v;
    atomicAdd(&dwsB[b * nv + v], g * L[o] * m);
    atomicAdd(&dL[o], g * wsB[b * nv + v] * m);
  }
}

int fine_reduce_fwd(const float* L, long long ldl, const unsigned char* mA, const float* wsA, const float* wsB,
                    float* score, unsigned char* arg_v, unsigned char* arg_t, int Na, int Nb, int T, int Vt, int v0,
                    int nv, cudaStream_t st) {
  VALOR_REQUIRE(nv <= FINE_MAX_V && T <= 255, "fine_reduce: nv=%d (max %d), T=%d (max 255)", nv, FINE_MAX_V, T);
  const long long pairs = (long long)Na * Nb;
  if (pairs == 0) return 0;
  fine_reduce_fwd_kernel<<<(unsigned)((pairs + 3) / 4), 128, 0, st>>>(L, ldl, mA, wsA, wsB, score, arg_v, arg_t, Na, Nb, T, Vt, v0, nv);
  return check_launch("fine_reduce_fwd_kernel");
}
int fine_reduce_bwd(const float* L, long long ldl, const unsigned char* mA, const float* wsA, const float* wsB,
                    const float* dscore, const unsigned char* arg_v, const unsigned char* arg_t, float* dL, float* dwsA,
                    float* dwsB, int Na, int Nb, int T, int Vt, int v0, int nv, cudaStream_t st) {
  const long long pairs = (long long)Na * Nb;
  if (pairs == 0) return 0;
  fine_reduce_bwd_kernel<<<(unsigned)((pairs + 3) / 4), 128, 0, st>>>(L, ldl, mA, wsA, wsB, dscore, arg_v, arg_t, dL, dwsA, dwsB, Na, Nb, T, Vt, v0, nv);
  return check_launch("fine_reduce_bwd_kernel");
}

// ---------------------------------------------------------------------------------------
// contrastive loss (modeling.py:418-433): s = S / temp; loss = mean over the 2N diagonal
// entries of -log_softmax(s, dim=1) and -log_softmax(s, dim=0).  Single CTA (N <= 4096).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
contrastive_fwd_kernel(const float* __restrict__ S, const float* __restrict__ temp, float* __restrict__ row_lse,
                       float* __restrict__ col_lse, float* __restrict__ loss, int N) {
  __shared__ float sh[32];
  const float it = 1.0f / temp[0];
  float part = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    float mr = -INFINITY, mc = -INFINITY;
    for (int j = 0; j < N; ++j) {
      mr = fmaxf(mr, S[(long long)i * N + j] * it);
      mc = fmaxf(mc, S[(long long)j * N + i] * it);
    }
    float sr = 0.f, sc = 0.f;
    for (int j = 0; j < N; ++j) {
      sr += __expf(S[(long long)i * N + j] * it - mr);
      sc += __expf(S[(long long)j * N + i] * it - mc);
    }
    const float lr = mr + __logf(sr), lc = mc + __logf(sc);
    row_lse[i] = lr;
    col_lse[i] = lc;
    const float d = S[(long long)i * N + i] * it;
    part += (lr - d) + (lc - d);
  }
  part = block_sum(part, sh);
  if (threadIdx.x == 0) loss[0] = part / (2.0f * N);
}
// dS = g/(2N temp) * (softmax_row + softmax_col - 2 I);  dtemp += sum ds * (-S/temp^2)
__global__ void __launch_bounds__(256)
contrastive_bwd_kernel(const float* __restrict__ S, const float* __restrict__ temp, const float* __restrict__ row_lse,
                       const float* __restrict__ col_lse, const float* __restrict__ gptr, float gmul,
                       float* __restrict__ dS, float* __restrict__ dtemp, int N) {
  __shared__ float sh[32];
  const float t = temp[0], it = 1.0f / t;
  const float g = (gptr ? gptr[0] : 1.0f) * gmul / (2.0f * N);
  float acc = 0.f;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < (long long)N * N;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx / N), j = (int)(idx % N);
    const float s = S[idx] * it;
    float ds = __expf(s - row_lse[i]) + __expf(s - col_lse[j]);
    if (i == j) ds -= 2.0f;
    ds *= g;
    dS[idx] = ds * it;
    acc += ds * (-S[idx] * it * it);
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0 && dtemp) atomicAdd(dtemp, acc);
}
int contrastive_fwd(const float* S, const float* temp, float* row_lse, float* col_lse, float* loss, int N, cudaStream_t st) {
  contrastive_fwd_kernel<<<1, 1024, 0, st>>>(S, temp, row_lse, col_lse, loss, N);
  return check_launch("contrastive_fwd_kernel");
}
int contrastive_bwd(const float* S, const float* temp, const float* row_lse, const float* col_lse, const float* gptr,
                    float gmul, float* dS, float* dtemp, int N, cudaStream_t st) {
  long long total = (long long)N * N;
  unsigned g = (unsigned)((total + 255) / 256);
  if (g > 1024) g = 1024;
  contrastive_bwd_kernel<<<g, 256, 0, st>>>(S, temp, row_lse, col_lse, gptr, gmul, dS, dtemp, N);
  return check_launch("contrastive_bwd_kernel");
}


// ---------------------------------------------------------------------------------------
// retrieval ranks (test.py:714-775 compute_metric_ret): the reference sorts every row of the score matrix on the
// device, copies the index matrix to the host and looks the ground truth up with list.index (O(N^2) Python).  The
// rank of the ground truth is simply the number of candidates that score higher: one warp per query,
//   rank[i] = #{ j : S[i*sr + j*sc] > S[i*sr + gt[i]*sc] }       (sr/sc strides: rows or columns of the matrix)
// dual softmax (test.py:685-713, off in every shipped config) rescales the scores first:
//   S'[i,j] = S[i,j] * softmax(S[:,j] / temp)[i] * n      (forward direction: softmax down the columns)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
retrieval_rank_kernel(const float* __restrict__ S, long long sr, long long sc, const int* __restrict__ gt, int* __restrict__ rank,
                      int Nq, int Nc) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= Nq) return;
  const float* row = S + (long long)i * sr;
  const float ref = row[(long long)gt[i] * sc];
  int cnt = 0;
  for (int j = lane; j < Nc; j += 32) cnt += row[(long long)j * sc] > ref ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) rank[i] = cnt;
}
// out[i,j] = S[i,j] * softmax_over_i(S[:,j] / temp)[i] * Nr   (one block per column j; dim = 1: over j per row i via strides)
__global__ void __launch_bounds__(256)
dual_softmax_kernel(const float* __restrict__ S, float* __restrict__ out, long long sr, long long sc, const float* __restrict__ temp,
                    int Nr, int Nc) {
  __shared__ float sh[32];
  const int j = blockIdx.x;
  const float it = 1.0f / temp[0];
  float m = -INFINITY;
  for (int i = threadIdx.x; i < Nr; i += blockDim.x) m = fmaxf(m, S[(long long)i * sr + (long long)j * sc] * it);
  m = block_max(m, sh);
  float z = 0.f;
  for (int i = threadIdx.x; i < Nr; i += blockDim.x) z += __expf(S[(long long)i * sr + (long long)j * sc] * it - m);
  z = block_sum(z, sh);
  const float sc_out = (float)Nr / z;
  for (int i = threadIdx.x; i < Nr; i += blockDim.x) {
    const float v = S[(long long)i * sr + (long long)j * sc];
    out[(long long)i * sr + (long long)j * sc] = v * __expf(v * it - m) * sc_out;
  }
}
int retrieval_rank(const float* S, long long sr, long long sc, const int* gt, int* rank, int Nq, int Nc, cudaStream_t st) {
  if (Nq == 0) return 0;
  retrieval_rank_kernel<<<(Nq + 7) / 8, 256, 0, st>>>(S, sr, sc, gt, rank, Nq, Nc);
  return check_launch("retrieval_rank_kernel");
}
int dual_softmax(const float* S, float* out, long long sr, long long sc, const float* temp, int Nr, int Nc, cudaStream_t st) {
  if (Nr == 0 || Nc == 0) return 0;
  dual_softmax_kernel<<<Nc, 256, 0, st>>>(S, out, sr, sc, temp, Nr, Nc);
  return check_launch("dual_softmax_kernel");
}

}  // namespace valor
