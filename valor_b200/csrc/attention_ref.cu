// valor_b200 — row-per-warp attention kernels (any dtype, any shape): the fp32 parity-mode
// path and the fallback for shapes the tensor-core kernels do not cover.
//
// Two index families share one core:
//   MHA    : BERT self / cross attention (bert.py:244-340), AST attention (transformer.py:115-130)
//   WINDOW : VideoSwin shifted-window attention (videoswin.py:137-163,191-226) evaluated IN PLACE
//            on the natural [B,D,H,W] token order: cyclic shift, window partition/reverse,
//            relative-position-bias gather and the -100 shift mask are index arithmetic here,
//            not tensor copies.
#include "common.cuh"
#include "attention.cuh"

namespace valor {

template <typename T, typename IDX>
__global__ void __launch_bounds__(128)
attn_ref_fwd_kernel(IDX ix, const T* __restrict__ Q, const T* __restrict__ K, const T* __restrict__ V, long long ldq,
                    long long ldk, long long ldv, T* __restrict__ O, long long ldo, float* __restrict__ lse, int H,
                    int hd, int Nq, float scale) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p = blockIdx.y, h = blockIdx.z;
  const int i = blockIdx.x * 4 + warp;
  if (i >= Nq) return;
  const int nk = ix.nk(p);
  float* s = sm + (size_t)warp * (ix.max_nk + hd);
  float* qs = s + ix.max_nk;
  const size_t qr = ix.qrow(p, i);
  for (int d = lane; d < hd; d += 32) qs[d] = to_f(Q[qr * ldq + h * hd + d]) * ix.qscale(scale);
  __syncwarp();
  float mx = -INFINITY;
  for (int j = lane; j < nk; j += 32) {
    const size_t kr = ix.krow(p, j);
    const T* kp = K + kr * ldk + h * hd;
    float a = 0.f;
    for (int d = 0; d < hd; ++d) a = fmaf(qs[d], to_f(kp[d]), a);
    a = a * ix.sscale(scale) + ix.add(p, h, i, j);
    s[j] = a;
    mx = fmaxf(mx, a);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < nk; j += 32) {
    float e = __expf(s[j] - mx);
    s[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  const float inv = 1.f / sum;
  for (int d = lane; d < hd; d += 32) {
    float a = 0.f;
    for (int j = 0; j < nk; ++j) a = fmaf(s[j], to_f(V[(size_t)ix.krow(p, j) * ldv + h * hd + d]), a);
    O[qr * ldo + h * hd + d] = from_f<T>(a * inv);
  }
  if (lane == 0) lse[((size_t)p * H + h) * Nq + i] = mx + __logf(sum);
}

// dQ written directly; dK/dV accumulated with fp32 atomics into zero-initialised buffers
// (rows shared by several problems — the cross-attention K/V of one sample feeds the
// tva/tv/ta passes — simply add up).
template <typename T, typename IDX>
__global__ void __launch_bounds__(128)
attn_ref_bwd_kernel(IDX ix, const T* __restrict__ Q, const T* __restrict__ K, const T* __restrict__ V,
                    const T* __restrict__ O, const T* __restrict__ dO, long long ldq, long long ldk, long long ldv,
                    long long ldo, const float* __restrict__ lse, T* __restrict__ dQ, long long lddq,
                    float* __restrict__ dK, float* __restrict__ dV, long long lddk, long long lddv,
                    float* __restrict__ dtable, int H, int hd, int Nq, float scale) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p = blockIdx.y, h = blockIdx.z;
  const int i = blockIdx.x * 4 + warp;
  if (i >= Nq) return;
  const int nk = ix.nk(p);
  float* qs = sm + (size_t)warp * (3 * hd);
  float* dos = qs + hd;
  float* dqs = dos + hd;
  const size_t qr = ix.qrow(p, i);
  float delta = 0.f;
  for (int d = lane; d < hd; d += 32) {
    qs[d] = to_f(Q[qr * ldq + h * hd + d]) * ix.qscale(scale);
    float g = to_f(dO[qr * ldo + h * hd + d]);
    dos[d] = g;
    dqs[d] = 0.f;
    delta += g * to_f(O[qr * ldo + h * hd + d]);
  }
  delta = warp_sum(delta);
  __syncwarp();
  const float L = lse[((size_t)p * H + h) * Nq + i];
  // lane-private dq accumulation, reduced at the end
  float dq_acc[128 / 1];  // hd <= 128 (checked on host)
  for (int d = 0; d < hd; ++d) dq_acc[d] = 0.f;
  for (int j = lane; j < nk; j += 32) {
    const size_t kr = ix.krow(p, j);
    const T* kp = K + kr * ldk + h * hd;
    const T* vp = V + kr * ldv + h * hd;
    float a = 0.f, dp = 0.f;
    for (int d = 0; d < hd; ++d) {
      a = fmaf(qs[d], to_f(kp[d]), a);
      dp = fmaf(dos[d], to_f(vp[d]), dp);
    }
    a = a * ix.sscale(scale) + ix.add(p, h, i, j);
    const float pr = __expf(a - L);
    const float ds = pr * (dp - delta);
    ix.add_grad(dtable, p, h, i, j, ds);
    const float dsq = ds * ix.sscale(scale);
    for (int d = 0; d < hd; ++d) {
      dq_acc[d] = fmaf(dsq, to_f(kp[d]), dq_acc[d]);
      atomicAdd(&dK[kr * lddk + h * hd + d], dsq * qs[d]);
      atomicAdd(&dV[kr * lddv + h * hd + d], pr * dos[d]);
    }
  }
  for (int d = 0; d < hd; ++d) {
    float v = warp_sum(dq_acc[d]);
    if (lane == 0) dqs[d] = v;
  }
  __syncwarp();
  for (int d = lane; d < hd; d += 32) dQ[qr * lddq + h * hd + d] = from_f<T>(dqs[d] * ix.qscale(scale));
}

template <typename T, typename IDX>
static int launch_fwd(const IDX& ix, const void* Q, const void* K, const void* V, long long ldq, long long ldk,
                      long long ldv, void* O, long long ldo, float* lse, int P, int H, int hd, int Nq, float scale,
                      cudaStream_t st) {
  dim3 grid((Nq + 3) / 4, P, H);
  VALOR_REQUIRE(P <= 65535 && H <= 65535, "attention: too many problems for grid.y (%d)", P);
  size_t smem = (size_t)4 * (ix.max_nk + hd) * sizeof(float);
  VALOR_REQUIRE(smem <= 200 * 1024, "attention: key count %d too large for the row kernel", ix.max_nk);
  auto kern = attn_ref_fwd_kernel<T, IDX>;
  if (smem > 48 * 1024) VALOR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, 128, smem, st>>>(ix, (const T*)Q, (const T*)K, (const T*)V, ldq, ldk, ldv, (T*)O, ldo, lse, H, hd, Nq, scale);
  return check_launch("attn_ref_fwd_kernel");
}

template <typename T, typename IDX>
static int launch_bwd(const IDX& ix, const void* Q, const void* K, const void* V, const void* O, const void* dO,
                      long long ldq, long long ldk, long long ldv, long long ldo, const float* lse, void* dQ,
                      long long lddq, float* dK, float* dV, long long lddk, long long lddv, float* dtable, int P, int H,
                      int hd, int Nq, float scale, cudaStream_t st) {
  dim3 grid((Nq + 3) / 4, P, H);
  VALOR_REQUIRE(P <= 65535, "attention: too many problems for grid.y (%d)", P);
  VALOR_REQUIRE(hd <= 128, "attention: head dim %d > 128", hd);
  size_t smem = (size_t)4 * 3 * hd * sizeof(float);
  attn_ref_bwd_kernel<T, IDX><<<grid, 128, smem, st>>>(ix, (const T*)Q, (const T*)K, (const T*)V, (const T*)O,
                                                        (const T*)dO, ldq, ldk, ldv, ldo, lse, (T*)dQ, lddq, dK, dV,
                                                        lddk, lddv, dtable, H, hd, Nq, scale);
  return check_launch("attn_ref_bwd_kernel");
}

int mha_ref_fwd(int dtype, const MhaIndex& ix, const void* Q, const void* K, const void* V, long long ldq,
                long long ldk, long long ldv, void* O, long long ldo, float* lse, int P, int H, int hd, int Nq,
                float scale, cudaStream_t st) {
  if (dtype == VALOR_DT_F32) return launch_fwd<float>(ix, Q, K, V, ldq, ldk, ldv, O, ldo, lse, P, H, hd, Nq, scale, st);
  return launch_fwd<bf16>(ix, Q, K, V, ldq, ldk, ldv, O, ldo, lse, P, H, hd, Nq, scale, st);
}
int mha_ref_bwd(int dtype, const MhaIndex& ix, const void* Q, const void* K, const void* V, const void* O,
                const void* dO, long long ldq, long long ldk, long long ldv, long long ldo, const float* lse, void* dQ,
                long long lddq, float* dK, float* dV, long long lddk, long long lddv, int P, int H, int hd, int Nq,
                float scale, cudaStream_t st) {
  if (dtype == VALOR_DT_F32)
    return launch_bwd<float>(ix, Q, K, V, O, dO, ldq, ldk, ldv, ldo, lse, dQ, lddq, dK, dV, lddk, lddv, nullptr, P, H,
                             hd, Nq, scale, st);
  return launch_bwd<bf16>(ix, Q, K, V, O, dO, ldq, ldk, ldv, ldo, lse, dQ, lddq, dK, dV, lddk, lddv, nullptr, P, H, hd,
                          Nq, scale, st);
}
int window_ref_fwd(int dtype, const WindowIndex& ix, const void* qkv, long long ld, void* O, long long ldo, float* lse,
                   int P, int H, int hd, float scale, cudaStream_t st) {
  const int C = H * hd;
  const size_t es = dtype == VALOR_DT_F32 ? 4 : 2;
  const char* b = (const char*)qkv;
  if (dtype == VALOR_DT_F32)
    return launch_fwd<float>(ix, b, b + C * es, b + 2 * C * es, ld, ld, ld, O, ldo, lse, P, H, hd, ix.N, scale, st);
  return launch_fwd<bf16>(ix, b, b + C * es, b + 2 * C * es, ld, ld, ld, O, ldo, lse, P, H, hd, ix.N, scale, st);
}
int window_ref_bwd(int dtype, const WindowIndex& ix, const void* qkv, long long ld, const void* O, const void* dO,
                   long long ldo, const float* lse, void* dQ, long long lddq, float* dK, float* dV, long long lddkv,
                   float* dtable, int P, int H, int hd, float scale, cudaStream_t st) {
  const int C = H * hd;
  const size_t es = dtype == VALOR_DT_F32 ? 4 : 2;
  const char* b = (const char*)qkv;
  if (dtype == VALOR_DT_F32)
    return launch_bwd<float>(ix, b, b + C * es, b + 2 * C * es, O, dO, ld, ld, ld, ldo, lse, dQ, lddq, dK, dV, lddkv,
                             lddkv, dtable, P, H, hd, ix.N, scale, st);
  return launch_bwd<bf16>(ix, b, b + C * es, b + 2 * C * es, O, dO, ld, ld, ld, ldo, lse, dQ, lddq, dK, dV, lddkv, lddkv,
                          dtable, P, H, hd, ix.N, scale, st);
}

}  // namespace valor
