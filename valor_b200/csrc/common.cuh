// valor_b200 — shared device/host helpers for the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>

#define VALOR_DT_F32 0
#define VALOR_DT_BF16 1

namespace valor {

typedef __nv_bfloat16 bf16;

// ---- error plumbing (C ABI returns int, message via valor_last_error) -----------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define VALOR_REQUIRE(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      valor::set_error(__VA_ARGS__);        \
      return 1;                             \
    }                                       \
  } while (0)

#define VALOR_CUDA(expr)                                                          \
  do {                                                                            \
    cudaError_t _e = (expr);                                                      \
    if (_e != cudaSuccess) {                                                      \
      valor::set_error("%s failed: %s", #expr, cudaGetErrorString(_e));           \
      return 1;                                                                   \
    }                                                                             \
  } while (0)

// ---- dtype helpers ---------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float ld_any(const void* p, int dtype, size_t i) {
  return dtype == VALOR_DT_BF16 ? __bfloat162float(((const bf16*)p)[i]) : ((const float*)p)[i];
}
__device__ __forceinline__ void st_any(void* p, int dtype, size_t i, float v) {
  if (dtype == VALOR_DT_BF16) ((bf16*)p)[i] = __float2bfloat16_rn(v);
  else ((float*)p)[i] = v;
}

// ---- activations (reference: erf-GELU bert.py:52-57 / transformer.py:32-38 / nn.GELU;
//      QuickGELU clip.py:167-169; ReLU pretrain.py:105) --------------------------------
#define VALOR_ACT_NONE 0
#define VALOR_ACT_GELU 1
#define VALOR_ACT_QUICKGELU 2
#define VALOR_ACT_RELU 3

// erf-GELU for the bf16 tensor-core epilogues: Phi(x) through Abramowitz-Stegun 7.1.26 (|erf error| <= 1.5e-7,
// invisible under the bf16 rounding of the result) -- one MUFU.RCP, one MUFU.EX2 and ~12 FMA-pipe instructions
// instead of the ~45 of erff + expf; exp(-x^2/2) is shared between erf and the Gaussian density.
struct GeluParts { float cdf, e; };   // Phi(x), exp(-x^2/2)
__device__ __forceinline__ GeluParts gelu_parts_fast(float x) {
  const float ax = fabsf(x);
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(ax, 0.3275911f * 0.70710678118654752f, 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * x * (-0.5f * 1.4426950408889634f)));
  float p = fmaf(t, 0.5f * 1.061405429f, 0.5f * -1.453152027f);
  p = fmaf(t, p, 0.5f * 1.421413741f);
  p = fmaf(t, p, 0.5f * -0.284496736f);
  p = fmaf(t, p, 0.5f * 0.254829592f);
  const float hq = p * t * e;                       // (1 - erf(|x|/sqrt2)) / 2
  GeluParts r;
  r.cdf = x >= 0.f ? 1.0f - hq : hq;
  r.e = e;
  return r;
}
__device__ __forceinline__ float gelu_fwd_fast(float x) { return x * gelu_parts_fast(x).cdf; }
__device__ __forceinline__ float gelu_grad_fast(float x) {
  const GeluParts g = gelu_parts_fast(x);
  return fmaf(x * 0.3989422804014327f, g.e, g.cdf);
}

__device__ __forceinline__ float act_fwd(float x, int act) {
  switch (act) {
    case VALOR_ACT_GELU: return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
    case VALOR_ACT_QUICKGELU: return x / (1.0f + __expf(-1.702f * x));
    case VALOR_ACT_RELU: return x > 0.f ? x : 0.f;
    default: return x;
  }
}
__device__ __forceinline__ float act_grad(float x, int act) {
  switch (act) {
    case VALOR_ACT_GELU: {
      float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
      float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
      return cdf + x * pdf;
    }
    case VALOR_ACT_QUICKGELU: {
      float s = 1.0f / (1.0f + __expf(-1.702f * x));
      return s * (1.0f + 1.702f * x * (1.0f - s));
    }
    case VALOR_ACT_RELU: return x > 0.f ? 1.f : 0.f;
    default: return 1.f;
  }
}

// ---- warp / block reductions -----------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// blockDim.x multiple of 32, <= 1024. `sh` must hold 32 floats.
__device__ __forceinline__ float block_sum(float v, float* sh) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  float r = (lane < nw) ? sh[lane] : 0.f;
  r = warp_sum(r);
  return r;
}
__device__ __forceinline__ float block_max(float v, float* sh) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  float r = (lane < nw) ? sh[lane] : -INFINITY;
  r = warp_max(r);
  return r;
}

// ---- counter-based random numbers (Dropout / DropPath / attention-probability dropout) ----------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
  }
  return c;
}
// four uniform 32-bit words for counter g of call site `site`; state = {seed, step offset} in device memory
__device__ __forceinline__ uint4 rng4(const long long* state, long long site, unsigned long long g) {
  const unsigned long long seed = (unsigned long long)state[0], off = (unsigned long long)state[1] + (unsigned long long)site;
  return philox4x32_10(make_uint4((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)off, (uint32_t)(off >> 32)),
                       make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}
__host__ __device__ __forceinline__ uint32_t drop_threshold(float p) {   // drop iff word < threshold
  const double t = (double)p * 4294967296.0;
  return t >= 4294967295.0 ? 0xffffffffu : (uint32_t)t;
}
// cheap per-element stream under a Philox-derived key (attention probabilities: one draw per (query, key) pair inside the
// softmax loops, where a 10-round Philox per element would double the kernel): murmur3 finaliser of key + element index
__device__ __forceinline__ uint32_t mix32(uint32_t key, uint32_t idx) {
  uint32_t x = idx * 0x9E3779B1u + key;
  x ^= x >> 15; x *= 0x85EBCA77u; x ^= x >> 13; x *= 0xC2B2AE3Du; x ^= x >> 16;
  return x;
}

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// ---- GEMM epilogue description (shared by the tcgen05 and SIMT kernels) -----------------
struct GemmEpilogue {
  const float* bias;     // [N] fp32 or null
  const void* residual;  // [M,N] (ld = ldr), dtype = res_dtype, added AFTER activation
  const void* act_aux;   // [M,N] pre-activation saved by the forward; out = acc * act'(aux)
  void* preact_out;      // [M,N] optional: pre-activation (acc + bias) in out dtype
  long long ldr, ld_aux, ld_pre;
  int res_dtype, aux_dtype;
  int act;               // VALOR_ACT_*
  int out_dtype;         // VALOR_DT_*
  int accumulate;        // 1: C += result (fp32 out only; atomic when split-K)
  float alpha;           // scales the accumulator before bias
  float* bias_grad;      // wgrad form only ([K,M] x [K,N] operands): bias_grad[m] += alpha * sum_k A[k,m]  (the Linear's bias
                         // gradient = column sums of dy), taken from a ones-column MMA on the A tiles already in shared memory
  const float* row_scale;  // null, or one factor per `rows_per_group` consecutive rows, applied to (alpha * acc + bias) before the
  int rows_per_group;      // residual add (DropPath inside the GEMM that ends a residual branch); act must be NONE
};

}  // namespace valor
