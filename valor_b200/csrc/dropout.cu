// valor_b200 — training-mode stochastic regularisation of the hot path:
//   nn.Dropout(0.1) on BERT / AST hidden states (bert.py:217,353,367,418; transformer.py:78,83; modeling.py:761) and
//   DropPath (videoswin.py:40-55,238: per-sample keep mask on each residual branch, rate 0 -> 0.2 over the 24 blocks).
// Masks are never stored: forward and backward regenerate them from a counter-based generator (Philox4x32-10) keyed by
// {seed, step offset} read from DEVICE memory (so a captured CUDA graph sees fresh randomness every replay: the host
// only rewrites 16 bytes) plus a per-call-site id and the element index.
#include "common.cuh"

namespace valor {

// out[r,c] = residual[r,c] + x[r,c] * keep / (1-p)      (C % 4 == 0; element index = r*C + c)
template <typename T>
__global__ void __launch_bounds__(256)
dropout_kernel(const T* __restrict__ x, long long ldx, const T* __restrict__ res, long long ldr, T* __restrict__ out, long long ldo,
               long long R, int C, float p, const long long* __restrict__ state, long long site) {
  const uint32_t thr = drop_threshold(p);
  const float inv = 1.0f / (1.0f - p);
  const long long groups = R * (C / 4);
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (long long)gridDim.x * blockDim.x) {
    const long long r = g / (C / 4);
    const int c = (int)(g - r * (C / 4)) * 4;
    const uint4 w = rng4(state, site, (unsigned long long)g);
    const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = to_f<T>(x[r * ldx + c + j]);
      v = ws[j] < thr ? 0.f : v * inv;
      if (res != nullptr) v += to_f<T>(res[r * ldr + c + j]);
      out[r * ldo + c + j] = from_f<T>(v);
    }
  }
}

// scale[b] = floor(keep_prob + u_b) / keep_prob   (videoswin.py:45-50), one uniform per sample
__global__ void droppath_scale_kernel(float* __restrict__ scale, int B, float p, const long long* __restrict__ state, long long site) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const uint4 w = rng4(state, site, (unsigned long long)b);
  const float u = (float)(w.x >> 8) * (1.0f / 16777216.0f);     // [0,1)
  const float keep = 1.0f - p;
  scale[b] = floorf(keep + u) / keep;
}

// out[r,c] = residual[r,c] + x[r,c] * scale[r / rows_per_group]
template <typename T>
__global__ void __launch_bounds__(256)
row_scale_kernel(const T* __restrict__ x, long long ldx, const float* __restrict__ scale, long long rpg, const T* __restrict__ res,
                 long long ldr, T* __restrict__ out, long long ldo, long long R, int C) {
  constexpr int V = 16 / sizeof(T);
  const long long groups = R * (C / V);
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (long long)gridDim.x * blockDim.x) {
    const long long r = g / (C / V);
    const int c = (int)(g - r * (C / V)) * V;
    const float s = scale[r / rpg];
    uint4 xv = *(const uint4*)(x + r * ldx + c), rv = make_uint4(0, 0, 0, 0), ov;
    if (res != nullptr) rv = *(const uint4*)(res + r * ldr + c);
    const T* xp = (const T*)&xv; const T* rp = (const T*)&rv; T* op = (T*)&ov;
#pragma unroll
    for (int j = 0; j < V; ++j) op[j] = from_f<T>(to_f<T>(xp[j]) * s + (res != nullptr ? to_f<T>(rp[j]) : 0.f));
    *(uint4*)(out + r * ldo + c) = ov;
  }
}

static unsigned grid_for(long long work) {
  long long g = (work + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (g > cap) g = cap;
  return (unsigned)(g < 1 ? 1 : g);
}

int dropout_apply(int dtype, const void* x, long long ldx, const void* res, long long ldr, void* out, long long ldo, long long R,
                  int C, float p, const long long* state, long long site, cudaStream_t st) {
  VALOR_REQUIRE(C % 4 == 0 && p >= 0.f && p < 1.f, "dropout: need C %% 4 == 0 and 0 <= p < 1");
  if (R == 0) return 0;
  if (dtype == VALOR_DT_BF16)
    dropout_kernel<bf16><<<grid_for(R * (C / 4)), 256, 0, st>>>((const bf16*)x, ldx, (const bf16*)res, ldr, (bf16*)out, ldo, R, C, p, state, site);
  else
    dropout_kernel<float><<<grid_for(R * (C / 4)), 256, 0, st>>>((const float*)x, ldx, (const float*)res, ldr, (float*)out, ldo, R, C, p, state, site);
  return check_launch("dropout_kernel");
}
int droppath_scale(float* scale, int B, float p, const long long* state, long long site, cudaStream_t st) {
  if (B == 0) return 0;
  droppath_scale_kernel<<<(B + 127) / 128, 128, 0, st>>>(scale, B, p, state, site);
  return check_launch("droppath_scale_kernel");
}
int row_scale(int dtype, const void* x, long long ldx, const float* scale, long long rpg, const void* res, long long ldr, void* out,
              long long ldo, long long R, int C, cudaStream_t st) {
  const int V = dtype == VALOR_DT_BF16 ? 8 : 4;
  VALOR_REQUIRE(C % V == 0 && ldx % V == 0 && ldo % V == 0 && (res == nullptr || ldr % V == 0) &&
                (((uintptr_t)x | (uintptr_t)out | (uintptr_t)res) & 15) == 0, "row_scale: rows must be 16-byte aligned");
  if (R == 0) return 0;
  if (dtype == VALOR_DT_BF16)
    row_scale_kernel<bf16><<<grid_for(R * (C / V)), 256, 0, st>>>((const bf16*)x, ldx, scale, rpg, (const bf16*)res, ldr, (bf16*)out, ldo, R, C);
  else
    row_scale_kernel<float><<<grid_for(R * (C / V)), 256, 0, st>>>((const float*)x, ldx, scale, rpg, (const float*)res, ldr, (float*)out, ldo, R, C);
  return check_launch("row_scale_kernel");
}

}  // namespace valor
