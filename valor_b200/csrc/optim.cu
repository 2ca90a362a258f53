// valor_b200 — flat multi-tensor optimizer step (replaces the per-tensor Python loop of
// optim/adamw.py:50-101, torch clip_grad_norm_ at train_utils.py:359 and apex amp's
// master->model copy): one launch over a contiguous fp32 master/grad/moment arena, fused
// with the bf16 working-copy refresh.
#include "common.cuh"

namespace valor {

__global__ void __launch_bounds__(256)
sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ out) {
  __shared__ float sh[32];
  float s = 0.f;
  const long long n4 = n / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = ((const float4*)g)[i];
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    s += g[i] * g[i];
  s = block_sum(s, sh);
  if (threadIdx.x == 0) atomicAdd(out, s);
}
// norm_out[0] = sqrt(sumsq); norm_out[1] = clip coefficient min(1, max_norm / (norm + 1e-6))
__global__ void clip_coef_kernel(const float* __restrict__ sumsq, float max_norm, float* __restrict__ norm_out) {
  const float n = sqrtf(sumsq[0]);
  norm_out[0] = n;
  float c = 1.0f;
  if (max_norm > 0.f) c = fminf(1.0f, max_norm / (n + 1e-6f));
  norm_out[1] = c;
}

// hyper (device, refreshed by the host every step so the launch can live in a CUDA graph):
//   [0]=lr  [1]=beta1  [2]=beta2  [3]=eps  [4]=weight_decay  [5]=step_size (bias-corrected lr)
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             bf16* __restrict__ p_lp, long long n, const float* __restrict__ hyper, const float* __restrict__ coef) {
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4], step_size = hyper[5];
  const float c = coef ? coef[0] : 1.0f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * c;
    const float mi = m[i] * b1 + (1.0f - b1) * gi;
    const float vi = v[i] * b2 + (1.0f - b2) * gi * gi;
    float pi = p[i];
    pi -= step_size * mi / (sqrtf(vi) + eps);
    if (wd > 0.f) pi -= lr * wd * pi;
    m[i] = mi;
    v[i] = vi;
    p[i] = pi;
    if (p_lp) p_lp[i] = __float2bfloat16_rn(pi);
  }
}

int grad_sumsq(const float* g, long long n, float* out, cudaStream_t st) {
  if (n == 0) return 0;
  long long b = (n / 4 + 255) / 256;
  long long cap = (long long)num_sms() * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  sumsq_kernel<<<(unsigned)b, 256, 0, st>>>(g, n, out);
  return check_launch("sumsq_kernel");
}
int clip_coef(const float* sumsq, float max_norm, float* norm_out, cudaStream_t st) {
  clip_coef_kernel<<<1, 1, 0, st>>>(sumsq, max_norm, norm_out);
  return check_launch("clip_coef_kernel");
}
int adamw(float* p, const float* g, float* m, float* v, void* p_lp, long long n, const float* hyper, const float* coef,
          cudaStream_t st) {
  if (n == 0) return 0;
  long long b = (n + 255) / 256;
  long long cap = (long long)num_sms() * 16;
  if (b > cap) b = cap;
  adamw_kernel<<<(unsigned)b, 256, 0, st>>>(p, g, m, v, (bf16*)p_lp, n, hyper, coef);
  return check_launch("adamw_kernel");
}

}  // namespace valor
