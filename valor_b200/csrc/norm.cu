// valor_b200 — row-wise normalisation kernels (HBM-bound; one warp per row, 8-byte/16-byte
// vector accesses, fp32 statistics).
//   layernorm : apex FusedLayerNorm (apex/csrc/layer_norm_cuda_kernel.cu cuApplyLayerNorm /
//               cuComputeGradInput / cuComputeGradGammaBeta) and nn.LayerNorm in Swin
//   l2norm    : F.normalize(dim=-1) on the contrastive features (pretrain.py:276,283,289)
#include "common.cuh"

namespace valor {

template <typename T> struct Vec4;
template <> struct Vec4<float> {
  float4 v;
  __device__ __forceinline__ void load(const float* p) { v = *(const float4*)p; }
  __device__ __forceinline__ void store(float* p) const { *(float4*)p = v; }
  __device__ __forceinline__ void get(float* f) const { f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w; }
  __device__ __forceinline__ void set(const float* f) { v = make_float4(f[0], f[1], f[2], f[3]); }
};
template <> struct Vec4<bf16> {
  uint2 v;
  __device__ __forceinline__ void load(const bf16* p) { v = *(const uint2*)p; }
  __device__ __forceinline__ void store(bf16* p) const { *(uint2*)p = v; }
  __device__ __forceinline__ void get(float* f) const {
    const __nv_bfloat162* h = (const __nv_bfloat162*)&v;
    float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
  }
  __device__ __forceinline__ void set(const float* f) {
    __nv_bfloat162* h = (__nv_bfloat162*)&v;
    h[0] = __floats2bfloat162_rn(f[0], f[1]);
    h[1] = __floats2bfloat162_rn(f[2], f[3]);
  }
};

// ---------------------------------------------------------------------------------------
// LayerNorm forward: y = (x - mean) * rstd * gamma + beta ; saves mean, rstd (fp32)
// ---------------------------------------------------------------------------------------
// generic N (N % 4 == 0): one warp per row, three passes over the (L1-resident) row
template <typename T>
__global__ void __launch_bounds__(256)
layernorm_fwd_generic_kernel(const T* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                             T* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, long long M, int N,
                             float eps) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const T* xr = x + row * N;
  T* yr = y + row * N;
  const int nv = N >> 2;
  float s = 0.f;
  for (int i = lane; i < nv; i += 32) {
    Vec4<T> v; v.load(xr + 4 * i);
    float f[4]; v.get(f);
    s += f[0] + f[1] + f[2] + f[3];
  }
  const float mean = warp_sum(s) / N;
  float ss = 0.f;
  for (int i = lane; i < nv; i += 32) {
    Vec4<T> v; v.load(xr + 4 * i);
    float f[4]; v.get(f);
#pragma unroll
    for (int e = 0; e < 4; ++e) { float d = f[e] - mean; ss += d * d; }
  }
  const float rstd = rsqrtf(warp_sum(ss) / N + eps);
  for (int i = lane; i < nv; i += 32) {
    Vec4<T> v; v.load(xr + 4 * i);
    float f[4]; v.get(f);
    const float4 g = *(const float4*)(gamma + 4 * i);
    const float4 b = *(const float4*)(beta + 4 * i);
    f[0] = (f[0] - mean) * rstd * g.x + b.x;
    f[1] = (f[1] - mean) * rstd * g.y + b.y;
    f[2] = (f[2] - mean) * rstd * g.z + b.z;
    f[3] = (f[3] - mean) * rstd * g.w + b.w;
    v.set(f); v.store(yr + 4 * i);
  }
  if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
}

// N == VPL*128: the row lives in registers (one read of x), and a warp works on R rows at once so that R*VPL
// independent loads are in flight per lane -- a single 256-byte row per warp leaves HBM latency-bound.
template <typename T, int VPL, int R>
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const T* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     T* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, long long M, float eps) {
  constexpr int N = VPL * 128;
  const int lane = threadIdx.x & 31;
  const long long row0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * R;
  if (row0 >= M) return;
  float f[R][VPL][4];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const long long row = row0 + r < M ? row0 + r : M - 1;   // clamp: tail rows are recomputed, never stored
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      Vec4<T> v; v.load(x + row * N + (k * 32 + lane) * 4);
      v.get(f[r][k]);
    }
  }
  float mean[R], rstd[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) s += f[r][k][0] + f[r][k][1] + f[r][k][2] + f[r][k][3];
    mean[r] = s;
  }
#pragma unroll
  for (int r = 0; r < R; ++r) mean[r] = warp_sum(mean[r]) / N;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k)
#pragma unroll
      for (int e = 0; e < 4; ++e) { const float d = f[r][k][e] - mean[r]; ss += d * d; }
    rstd[r] = ss;
  }
#pragma unroll
  for (int r = 0; r < R; ++r) rstd[r] = rsqrtf(warp_sum(rstd[r]) / N + eps);
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int c = (k * 32 + lane) * 4;
    const float4 g = *(const float4*)(gamma + c);
    const float4 b = *(const float4*)(beta + c);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (row0 + r < M) {
        float o[4];
        o[0] = (f[r][k][0] - mean[r]) * rstd[r] * g.x + b.x;
        o[1] = (f[r][k][1] - mean[r]) * rstd[r] * g.y + b.y;
        o[2] = (f[r][k][2] - mean[r]) * rstd[r] * g.z + b.z;
        o[3] = (f[r][k][3] - mean[r]) * rstd[r] * g.w + b.w;
        Vec4<T> v; v.set(o); v.store(y + (row0 + r) * N + c);
      }
    }
  }
  if (lane < R && row0 + lane < M) {
    float mo = mean[0], ro = rstd[0];
#pragma unroll
    for (int r = 1; r < R; ++r) if (lane == r) { mo = mean[r]; ro = rstd[r]; }
    mean_out[row0 + lane] = mo; rstd_out[row0 + lane] = ro;
  }
}

// ---------------------------------------------------------------------------------------
// LayerNorm backward: dx per row; dgamma/dbeta accumulated per lane in registers over the
// rows a warp visits (each lane owns the same columns on every row), then reduced through
// shared memory and flushed with one atomicAdd per column per CTA into fp32 gradients.
// ---------------------------------------------------------------------------------------
template <typename T, int VPL /* float4-groups per lane; N == VPL*128 */, int R /* rows in flight per warp */,
          bool PARAM_GRADS /* accumulate dgamma/dbeta here (8*VPL registers); wide rows use ln_param_grad_kernel instead */>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ mean, const float* __restrict__ rstd, const T* __restrict__ dres,
                     T* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta, long long M, int N) {
  extern __shared__ float sh[];  // [2][N]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int i = threadIdx.x; i < 2 * N; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  constexpr bool kCacheGamma = VPL <= 4;   // wide rows re-read gamma from L1 instead of pinning 4*VPL registers
  constexpr int AV = PARAM_GRADS ? VPL : 1;
  float ag[AV][4], ab[AV][4], gm[kCacheGamma ? VPL : 1][4];
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    if (kCacheGamma) {
      const float4 g = *(const float4*)(gamma + (k * 32 + lane) * 4);
      gm[k][0] = g.x; gm[k][1] = g.y; gm[k][2] = g.z; gm[k][3] = g.w;
    }
    if (PARAM_GRADS) {
#pragma unroll
      for (int e = 0; e < 4; ++e) { ag[k][e] = 0.f; ab[k][e] = 0.f; }
    }
  }
  for (long long row0 = ((long long)blockIdx.x * nwarp + warp) * R; row0 < M; row0 += (long long)gridDim.x * nwarp * R) {
    float fdy[R][VPL][4], fxh[R][VPL][4], mu[R], rs[R], s1[R], s2[R];
    Vec4<T> rv[R][VPL];
#pragma unroll
    for (int r = 0; r < R; ++r) {   // all loads of the R rows are issued before the first use
      const bool ok = row0 + r < M;
      const long long row = ok ? row0 + r : M - 1;
      mu[r] = mean[row]; rs[r] = rstd[row];
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int c = (k * 32 + lane) * 4;
        Vec4<T> a, b; a.load(dy + row * N + c); b.load(x + row * N + c);
        a.get(fdy[r][k]); b.get(fxh[r][k]);
        if (dres != nullptr) rv[r][k].load(dres + row * N + c);
        if (!ok) {
#pragma unroll
          for (int e = 0; e < 4; ++e) fdy[r][k][e] = 0.f;   // tail rows contribute nothing
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float a1 = 0.f, a2 = 0.f;
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        float gk[4];
        if (kCacheGamma) { gk[0] = gm[k][0]; gk[1] = gm[k][1]; gk[2] = gm[k][2]; gk[3] = gm[k][3]; }
        else { const float4 g = *(const float4*)(gamma + (k * 32 + lane) * 4); gk[0] = g.x; gk[1] = g.y; gk[2] = g.z; gk[3] = g.w; }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          fxh[r][k][e] = (fxh[r][k][e] - mu[r]) * rs[r];
          if (PARAM_GRADS) {
            ag[k][e] += fdy[r][k][e] * fxh[r][k][e];
            ab[k][e] += fdy[r][k][e];
          }
          fdy[r][k][e] *= gk[e];  // dy * gamma
          a1 += fdy[r][k][e];
          a2 += fdy[r][k][e] * fxh[r][k][e];
        }
      }
      s1[r] = a1; s2[r] = a2;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) { s1[r] = warp_sum(s1[r]) / N; s2[r] = warp_sum(s2[r]) / N; }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (row0 + r < M) {
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
          const int c = (k * 32 + lane) * 4;
          float o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) o[e] = (fdy[r][k][e] - s1[r] - fxh[r][k][e] * s2[r]) * rs[r];
          if (dres != nullptr) {  // gradient arriving through the residual branch that bypasses this LN
            float rf[4]; rv[r][k].get(rf);
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] += rf[e];
          }
          Vec4<T> v; v.set(o); v.store(dx + (row0 + r) * N + c);
        }
      }
    }
  }
  if (!PARAM_GRADS) return;
#pragma unroll
  for (int k = 0; k < AV; ++k) {
    const int c = (k * 32 + lane) * 4;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      atomicAdd(&sh[c + e], ag[k][e]);
      atomicAdd(&sh[N + c + e], ab[k][e]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    if (dgamma) atomicAdd(&dgamma[i], sh[i]);
    if (dbeta) atomicAdd(&dbeta[i], sh[N + i]);
  }
}

// dgamma += sum_r dy*xhat, dbeta += sum_r dy for wide rows: CTA = 128-column panel x row range (the panel re-reads
// dy and x, which the dx pass has just pulled through L2)
template <typename T>
__global__ void __launch_bounds__(256)
ln_param_grad_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ mean,
                     const float* __restrict__ rstd, float* __restrict__ dgamma, float* __restrict__ dbeta, long long M, int N,
                     int rows_per_block) {
  __shared__ float sh[2][8][128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 128 + lane * 4;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > M) r1 = M;
  float ag[4] = {0.f, 0.f, 0.f, 0.f}, ab[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long r = r0 + warp; r < r1; r += 8) {
    Vec4<T> a, b; a.load(dy + r * N + c0); b.load(x + r * N + c0);
    float fd[4], fx[4]; a.get(fd); b.get(fx);
    const float mu = mean[r], rs = rstd[r];
#pragma unroll
    for (int e = 0; e < 4; ++e) { ag[e] += fd[e] * ((fx[e] - mu) * rs); ab[e] += fd[e]; }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) { sh[0][warp][lane * 4 + e] = ag[e]; sh[1][warp][lane * 4 + e] = ab[e]; }
  __syncthreads();
  if (threadIdx.x < 128) {
    float tg = 0.f, tb = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { tg += sh[0][i][threadIdx.x]; tb += sh[1][i][threadIdx.x]; }
    const int col = blockIdx.x * 128 + threadIdx.x;
    if (dgamma) atomicAdd(&dgamma[col], tg);
    if (dbeta) atomicAdd(&dbeta[col], tb);
  }
}

// generic-N fallback (N % 4 == 0): re-reads the row, shared-memory atomics for dgamma/dbeta
template <typename T>
__global__ void __launch_bounds__(256)
layernorm_bwd_generic_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ gamma,
                             const float* __restrict__ mean, const float* __restrict__ rstd,
                             const T* __restrict__ dres, T* __restrict__ dx, float* __restrict__ dgamma,
                             float* __restrict__ dbeta, long long M, int N) {
  extern __shared__ float sh[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int i = threadIdx.x; i < 2 * N; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  for (long long row = (long long)blockIdx.x * nwarp + warp; row < M; row += (long long)gridDim.x * nwarp) {
    const T* dyr = dy + row * N;
    const T* xr = x + row * N;
    const float mu = mean[row], rs = rstd[row];
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < N; c += 32) {
      const float d = to_f(dyr[c]), xh = (to_f(xr[c]) - mu) * rs, g = gamma[c];
      atomicAdd(&sh[c], d * xh);
      atomicAdd(&sh[N + c], d);
      s1 += d * g;
      s2 += d * g * xh;
    }
    s1 = warp_sum(s1) / N;
    s2 = warp_sum(s2) / N;
    for (int c = lane; c < N; c += 32) {
      const float d = to_f(dyr[c]) * gamma[c], xh = (to_f(xr[c]) - mu) * rs;
      float o = (d - s1 - xh * s2) * rs;
      if (dres != nullptr) o += to_f(dres[row * N + c]);
      dx[row * N + c] = from_f<T>(o);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    if (dgamma) atomicAdd(&dgamma[i], sh[i]);
    if (dbeta) atomicAdd(&dbeta[i], sh[N + i]);
  }
}

template <typename T>
static int ln_bwd_dispatch(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                           const void* dres, void* dx, float* dgamma, float* dbeta, long long M, int N, cudaStream_t st) {
  long long want = (M + 7) / 8;
  int grid = (int)(want < (long long)num_sms() * 8 ? want : (long long)num_sms() * 8);
  if (grid < 1) grid = 1;
  size_t smem = (size_t)2 * N * sizeof(float);
#define LN_BWD_CASE(V, RR, PG)                                                                                  \
  case V * 128:                                                                                                 \
    layernorm_bwd_kernel<T, V, RR, PG><<<grid, 256, smem, st>>>((const T*)dy, (const T*)x, gamma, mean, rstd,   \
                                                                 (const T*)dres, (T*)dx, dgamma, dbeta, M, N);  \
    if (!PG && (dgamma || dbeta)) {                                                                             \
      if (check_launch("layernorm_bwd_kernel")) return 1;                                                       \
      long long rpb = (M * (N / 128) + (long long)num_sms() * 8 - 1) / ((long long)num_sms() * 8);              \
      if (rpb < 64) rpb = 64;                                                                                   \
      dim3 g2(N / 128, (unsigned)((M + rpb - 1) / rpb));                                                        \
      ln_param_grad_kernel<T><<<g2, 256, 0, st>>>((const T*)dy, (const T*)x, mean, rstd, dgamma, dbeta, M, N, (int)rpb); \
    }                                                                                                           \
    break;
  switch (N) {
    LN_BWD_CASE(1, 4, true) LN_BWD_CASE(2, 2, true) LN_BWD_CASE(4, 1, true) LN_BWD_CASE(6, 1, true)
    LN_BWD_CASE(8, 1, false) LN_BWD_CASE(16, 1, false)
    default:
      layernorm_bwd_generic_kernel<T><<<grid, 256, smem, st>>>((const T*)dy, (const T*)x, gamma, mean, rstd,
                                                                (const T*)dres, (T*)dx, dgamma, dbeta, M, N);
  }
#undef LN_BWD_CASE
  return check_launch("layernorm_bwd_kernel");
}

int layernorm_fwd(int dtype, const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                  long long M, int N, float eps, cudaStream_t st) {
  VALOR_REQUIRE(N % 4 == 0, "layernorm: N=%d must be a multiple of 4", N);
  if (M == 0) return 0;
#define LN_FWD_CASE(TT, V, RR)                                                                                   \
  case V * 128: {                                                                                                \
    const long long blocks = (M + 8 * RR - 1) / (8 * RR);                                                        \
    VALOR_REQUIRE(blocks < 2147483647LL, "layernorm: too many rows");                                            \
    layernorm_fwd_kernel<TT, V, RR><<<(unsigned)blocks, 256, 0, st>>>((const TT*)x, gamma, beta, (TT*)y, mean, rstd, M, eps); \
    return check_launch("layernorm_fwd_kernel");                                                                 \
  }
  if (dtype == VALOR_DT_F32) {
    switch (N) { LN_FWD_CASE(float, 1, 4) LN_FWD_CASE(float, 2, 2) LN_FWD_CASE(float, 4, 2) LN_FWD_CASE(float, 6, 1) LN_FWD_CASE(float, 8, 1) default: break; }
  } else {
    switch (N) { LN_FWD_CASE(bf16, 1, 4) LN_FWD_CASE(bf16, 2, 2) LN_FWD_CASE(bf16, 4, 2) LN_FWD_CASE(bf16, 6, 1) LN_FWD_CASE(bf16, 8, 1) default: break; }
  }
#undef LN_FWD_CASE
  const long long blocks = (M + 7) / 8;
  VALOR_REQUIRE(blocks < 2147483647LL, "layernorm: too many rows");
  if (dtype == VALOR_DT_F32)
    layernorm_fwd_generic_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((const float*)x, gamma, beta, (float*)y, mean, rstd, M, N, eps);
  else
    layernorm_fwd_generic_kernel<bf16><<<(unsigned)blocks, 256, 0, st>>>((const bf16*)x, gamma, beta, (bf16*)y, mean, rstd, M, N, eps);
  return check_launch("layernorm_fwd_kernel");
}

int layernorm_bwd(int dtype, const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                  const void* dres, void* dx, float* dgamma, float* dbeta, long long M, int N, cudaStream_t st) {
  VALOR_REQUIRE(N % 4 == 0, "layernorm: N=%d must be a multiple of 4", N);
  if (M == 0) return 0;
  if (dtype == VALOR_DT_F32) return ln_bwd_dispatch<float>(dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, M, N, st);
  return ln_bwd_dispatch<bf16>(dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, M, N, st);
}

// ---------------------------------------------------------------------------------------
// L2 normalise: y = x / max(||x||, 1e-12) ; bwd dx = (dy - y (y.dy)) / max(||x||,1e-12)
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void l2norm_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, float* __restrict__ nrm, long long M, int N) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  float s = 0.f;
  for (int c = lane; c < N; c += 32) { float v = to_f(x[row * N + c]); s += v * v; }
  const float n = fmaxf(sqrtf(warp_sum(s)), 1e-12f);
  for (int c = lane; c < N; c += 32) y[row * N + c] = from_f<T>(to_f(x[row * N + c]) / n);
  if (lane == 0) nrm[row] = n;
}
template <typename T>
__global__ void l2norm_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ nrm,
                                  T* __restrict__ dx, long long M, int N) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const float n = nrm[row];
  float s = 0.f;
  for (int c = lane; c < N; c += 32) s += to_f(dy[row * N + c]) * to_f(x[row * N + c]) / n;
  s = warp_sum(s);
  for (int c = lane; c < N; c += 32) {
    const float yv = to_f(x[row * N + c]) / n;
    dx[row * N + c] = from_f<T>((to_f(dy[row * N + c]) - yv * s) / n);
  }
}

int l2norm_fwd(int dtype, const void* x, void* y, float* nrm, long long M, int N, cudaStream_t st) {
  if (M == 0) return 0;
  unsigned blocks = (unsigned)((M + 7) / 8);
  if (dtype == VALOR_DT_F32) l2norm_fwd_kernel<float><<<blocks, 256, 0, st>>>((const float*)x, (float*)y, nrm, M, N);
  else l2norm_fwd_kernel<bf16><<<blocks, 256, 0, st>>>((const bf16*)x, (bf16*)y, nrm, M, N);
  return check_launch("l2norm_fwd_kernel");
}
int l2norm_bwd(int dtype, const void* dy, const void* x, const float* nrm, void* dx, long long M, int N, cudaStream_t st) {
  if (M == 0) return 0;
  unsigned blocks = (unsigned)((M + 7) / 8);
  if (dtype == VALOR_DT_F32) l2norm_bwd_kernel<float><<<blocks, 256, 0, st>>>((const float*)dy, (const float*)x, nrm, (float*)dx, M, N);
  else l2norm_bwd_kernel<bf16><<<blocks, 256, 0, st>>>((const bf16*)dy, (const bf16*)x, nrm, (bf16*)dx, M, N);
  return check_launch("l2norm_bwd_kernel");
}

}  // namespace valor
