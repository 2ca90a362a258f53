// valor_b200 — SIMT GEMM (fp32 parity mode + shapes the TMA path cannot take, e.g. N=1
// fine-weight heads, K=96 patch-embed).  Same epilogue contract as gemm_sm100.cu.
//   C[M,N] (+)= epilogue( alpha * sum_k A(m,k) * B(n,k) ),  A(m,k)=A[m*sam+k*sak], B(n,k)=B[n*sbn+k*sbk]
#include "common.cuh"

namespace valor {

template <typename T>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const T* __restrict__ A, long long sam, long long sak, const T* __restrict__ B, long long sbn,
                 long long sbk, void* __restrict__ C, long long ldc, int M, int N, int K, GemmEpilogue ep) {
  constexpr int TM = 64, TN = 64, TK = 16;
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += TK) {
    for (int i = threadIdx.x; i < TM * TK; i += 256) {
      int m, k;
      if (sak == 1) { k = i % TK; m = i / TK; } else { m = i % TM; k = i / TM; }
      float v = 0.f;
      if (m0 + m < M && k0 + k < K) v = to_f(A[(size_t)(m0 + m) * sam + (size_t)(k0 + k) * sak]);
      As[k][m] = v;
    }
    for (int i = threadIdx.x; i < TN * TK; i += 256) {
      int n, k;
      if (sbk == 1) { k = i % TK; n = i / TK; } else { n = i % TN; k = i / TN; }
      float v = 0.f;
      if (n0 + n < N && k0 + k < K) v = to_f(B[(size_t)(n0 + n) * sbn + (size_t)(k0 + k) * sbk]);
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
    if (row >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx * 4 + j;
      if (col >= N) continue;
      float y = acc[i][j] * ep.alpha;
      if (ep.bias) y += ep.bias[col];
      if (ep.row_scale) y *= ep.row_scale[row / ep.rows_per_group];
      if (ep.preact_out) st_any(ep.preact_out, ep.out_dtype, (size_t)row * ep.ld_pre + col, y);
      if (ep.act_aux) y *= act_grad(ld_any(ep.act_aux, ep.aux_dtype, (size_t)row * ep.ld_aux + col), ep.act);
      else y = act_fwd(y, ep.act);
      if (ep.residual) y += ld_any(ep.residual, ep.res_dtype, (size_t)row * ep.ldr + col);
      const size_t o = (size_t)row * ldc + col;
      if (ep.out_dtype == VALOR_DT_BF16) ((bf16*)C)[o] = __float2bfloat16_rn(y);
      else if (ep.accumulate) ((float*)C)[o] += y;
      else ((float*)C)[o] = y;
    }
  }
}

int gemm_simt(int dtype, const void* A, long long sam, long long sak, const void* B, long long sbn, long long sbk,
              void* C, long long ldc, int M, int N, int K, const GemmEpilogue& ep, cudaStream_t st) {
  dim3 grid((N + 63) / 64, (M + 63) / 64);
  VALOR_REQUIRE(grid.y <= 65535, "gemm_simt: M too large (%d)", M);
  if (dtype == VALOR_DT_F32)
    gemm_simt_kernel<float><<<grid, 256, 0, st>>>((const float*)A, sam, sak, (const float*)B, sbn, sbk, C, ldc, M, N, K, ep);
  else
    gemm_simt_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)A, sam, sak, (const bf16*)B, sbn, sbk, C, ldc, M, N, K, ep);
  return check_launch("gemm_simt_kernel");
}

}  // namespace valor
