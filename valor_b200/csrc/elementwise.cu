// valor_b200 — data-movement / gather / reduction kernels around the GEMMs (all HBM-bound).
#include "common.cuh"

namespace valor {

static inline unsigned grid_for(long long n, int per_block) {
  long long b = (n + per_block - 1) / per_block;
  long long cap = (long long)num_sms() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

// ---------------------------------------------------------------------------------------
// PatchEmbed3D im2col (videoswin.py:361-369): video [B,F,3,Hh,Ww] (pixels fp32 or T) ->
// cols [B*F*(Hh/4)*(Ww/4), 96], column k = ((c*2 + kd)*4 + kh)*4 + kw matches
// proj.weight.view(E, 96); frame index F (the padded frame, :367) reads as zero.
// ---------------------------------------------------------------------------------------
template <typename TI, typename T>
__global__ void swin_im2col_kernel(const TI* __restrict__ video, T* __restrict__ cols, int B, int F, int Hh, int Ww) {
  const int Ho = Hh / 4, Wo = Ww / 4;
  const long long total = (long long)B * F * Ho * Wo * 24;  // 24 groups of 4 contiguous kw
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(idx % 24);
    long long t = idx / 24;
    const int j = (int)(t % Wo); t /= Wo;
    const int i = (int)(t % Ho); t /= Ho;
    const int d = (int)(t % F);
    const int b = (int)(t / F);
    const int kh = g % 4, kd = (g / 4) % 2, c = g / 8;
    const int f = d + kd;
    T* out = cols + (idx / 24) * 96 + g * 4;
    if (f < F) {
      const TI* src = video + ((((long long)b * F + f) * 3 + c) * Hh + (4 * i + kh)) * Ww + 4 * j;
#pragma unroll
      for (int e = 0; e < 4; ++e) out[e] = from_f<T>(to_f(src[e]));
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) out[e] = from_f<T>(0.f);
    }
  }
}

int swin_im2col(int in_dtype, int dtype, const void* video, void* cols, int B, int F, int Hh, int Ww, cudaStream_t st) {
  VALOR_REQUIRE(Hh % 4 == 0 && Ww % 4 == 0, "swin_im2col: resolution must be a multiple of 4");
  const long long total = (long long)B * F * (Hh / 4) * (Ww / 4) * 24;
  unsigned g = grid_for(total, 256);
  if (in_dtype == VALOR_DT_F32 && dtype == VALOR_DT_F32)
    swin_im2col_kernel<float, float><<<g, 256, 0, st>>>((const float*)video, (float*)cols, B, F, Hh, Ww);
  else if (in_dtype == VALOR_DT_F32)
    swin_im2col_kernel<float, bf16><<<g, 256, 0, st>>>((const float*)video, (bf16*)cols, B, F, Hh, Ww);
  else if (dtype == VALOR_DT_BF16)
    swin_im2col_kernel<bf16, bf16><<<g, 256, 0, st>>>((const bf16*)video, (bf16*)cols, B, F, Hh, Ww);
  else
    VALOR_REQUIRE(false, "swin_im2col: unsupported dtype combination");
  return check_launch("swin_im2col_kernel");
}

// ---------------------------------------------------------------------------------------
// AudioEmbeddings conv-as-GEMM im2col (modeling.py:752-754): spec [BA, mel, frames] ->
// cols [BA*(mel/16)*(frames/16), 256], token = i*(frames/16)+j, column = kh*16+kw.
// ---------------------------------------------------------------------------------------
template <typename TI, typename T>
__global__ void audio_im2col_kernel(const TI* __restrict__ spec, T* __restrict__ cols, int BA, int mel, int frames, int ps) {
  const int Pi = mel / ps, Pj = frames / ps;
  const long long total = (long long)BA * Pi * Pj * ps * ps;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int kw = (int)(idx % ps);
    long long t = idx / ps;
    const int kh = (int)(t % ps); t /= ps;
    const int j = (int)(t % Pj); t /= Pj;
    const int i = (int)(t % Pi);
    const int b = (int)(t / Pi);
    cols[idx] = from_f<T>(to_f(spec[((long long)b * mel + i * ps + kh) * frames + j * ps + kw]));
  }
}
int audio_im2col(int in_dtype, int dtype, const void* spec, void* cols, int BA, int mel, int frames, int ps, cudaStream_t st) {
  const long long total = (long long)BA * mel * frames;
  unsigned g = grid_for(total, 256);
  if (in_dtype == VALOR_DT_F32 && dtype == VALOR_DT_F32)
    audio_im2col_kernel<float, float><<<g, 256, 0, st>>>((const float*)spec, (float*)cols, BA, mel, frames, ps);
  else if (in_dtype == VALOR_DT_F32)
    audio_im2col_kernel<float, bf16><<<g, 256, 0, st>>>((const float*)spec, (bf16*)cols, BA, mel, frames, ps);
  else if (dtype == VALOR_DT_BF16)
    audio_im2col_kernel<bf16, bf16><<<g, 256, 0, st>>>((const bf16*)spec, (bf16*)cols, BA, mel, frames, ps);
  else
    VALOR_REQUIRE(false, "audio_im2col: unsupported dtype combination");
  return check_launch("audio_im2col_kernel");
}

// ---------------------------------------------------------------------------------------
// AST token assembly (modeling.py:755-760): x[b,0] = cls + pos[0]; x[b,1+t] = tok[b,t] + pos[1+t]
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void ast_assemble_fwd_kernel(const T* __restrict__ tok, const float* __restrict__ cls,
                                        const float* __restrict__ pos, T* __restrict__ x, int BA, int P, int Hd) {
  const long long total = (long long)BA * (P + 1) * Hd;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % Hd);
    const long long r = idx / Hd;
    const int t = (int)(r % (P + 1));
    const long long b = r / (P + 1);
    const float v = (t == 0) ? cls[c] : to_f(tok[(b * P + t - 1) * Hd + c]);
    x[idx] = from_f<T>(v + pos[(long long)t * Hd + c]);
  }
}
// dtok[b,t] = dx[b,1+t]; dcls += sum_b dx[b,0]; dpos[t] += sum_b dx[b,t]
template <typename T>
__global__ void ast_assemble_bwd_kernel(const T* __restrict__ dx, T* __restrict__ dtok, float* __restrict__ dcls,
                                        float* __restrict__ dpos, int BA, int P, int Hd) {
  const long long total = (long long)(P + 1) * Hd;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % Hd);
    const int t = (int)(idx / Hd);
    float s = 0.f;
    for (int b = 0; b < BA; ++b) {
      const T v = dx[((long long)b * (P + 1) + t) * Hd + c];
      s += to_f(v);
      if (t > 0) dtok[((long long)b * P + t - 1) * Hd + c] = v;
    }
    if (dpos) atomicAdd(&dpos[idx], s);
    if (t == 0 && dcls) atomicAdd(&dcls[c], s);
  }
}
int ast_assemble_fwd(int dtype, const void* tok, const float* cls, const float* pos, void* x, int BA, int P, int Hd, cudaStream_t st) {
  unsigned g = grid_for((long long)BA * (P + 1) * Hd, 256);
  if (dtype == VALOR_DT_F32) ast_assemble_fwd_kernel<float><<<g, 256, 0, st>>>((const float*)tok, cls, pos, (float*)x, BA, P, Hd);
  else ast_assemble_fwd_kernel<bf16><<<g, 256, 0, st>>>((const bf16*)tok, cls, pos, (bf16*)x, BA, P, Hd);
  return check_launch("ast_assemble_fwd_kernel");
}
int ast_assemble_bwd(int dtype, const void* dx, void* dtok, float* dcls, float* dpos, int BA, int P, int Hd, cudaStream_t st) {
  unsigned g = grid_for((long long)(P + 1) * Hd, 128);
  if (dtype == VALOR_DT_F32) ast_assemble_bwd_kernel<float><<<g, 128, 0, st>>>((const float*)dx, (float*)dtok, dcls, dpos, BA, P, Hd);
  else ast_assemble_bwd_kernel<bf16><<<g, 128, 0, st>>>((const bf16*)dx, (bf16*)dtok, dcls, dpos, BA, P, Hd);
  return check_launch("ast_assemble_bwd_kernel");
}

// ---------------------------------------------------------------------------------------
// BertEmbeddings gather (bert.py:203-215, token_type None): e = word[tok] + pos[t] + type[0]
// (tables read from the fp32 masters).  bwd scatters into fp32 gradients.
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void bert_embed_fwd_kernel(const long long* __restrict__ tokens, const float* __restrict__ word,
                                      const float* __restrict__ pos, const float* __restrict__ type0,
                                      T* __restrict__ e, long long R, int Tn, int Hd) {
  const long long total = R * Tn * (Hd / 4);
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % (Hd / 4)) * 4;
    const long long r = idx / (Hd / 4);
    const int t = (int)(r % Tn);
    const long long id = tokens[r];
    const float4 w = *(const float4*)(word + id * Hd + c);
    const float4 p = *(const float4*)(pos + (long long)t * Hd + c);
    const float4 y = *(const float4*)(type0 + c);
    T* o = e + r * Hd + c;
    o[0] = from_f<T>(w.x + p.x + y.x);
    o[1] = from_f<T>(w.y + p.y + y.y);
    o[2] = from_f<T>(w.z + p.z + y.z);
    o[3] = from_f<T>(w.w + p.w + y.w);
  }
}
template <typename T>
__global__ void bert_embed_bwd_kernel(const T* __restrict__ de, const long long* __restrict__ tokens,
                                      float* __restrict__ dword, float* __restrict__ dpos, float* __restrict__ dtype0,
                                      long long R, int Tn, int Hd) {
  // one block column-slice per (t, c): sums over sequences for pos/type, scatters to word rows
  const long long total = (long long)Tn * Hd;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % Hd);
    const int t = (int)(idx / Hd);
    float s = 0.f;
    for (long long r = 0; r < R; ++r) {
      const float v = to_f(de[(r * Tn + t) * Hd + c]);
      s += v;
      atomicAdd(&dword[tokens[r * Tn + t] * Hd + c], v);
    }
    atomicAdd(&dpos[(long long)t * Hd + c], s);
    atomicAdd(&dtype0[c], s);
  }
}
int bert_embed_fwd(int dtype, const long long* tokens, const float* word, const float* pos, const float* type0, void* e,
                   long long R, int Tn, int Hd, cudaStream_t st) {
  VALOR_REQUIRE(Hd % 4 == 0, "bert_embed: hidden must be a multiple of 4");
  unsigned g = grid_for(R * Tn * (Hd / 4), 256);
  if (dtype == VALOR_DT_F32) bert_embed_fwd_kernel<float><<<g, 256, 0, st>>>(tokens, word, pos, type0, (float*)e, R, Tn, Hd);
  else bert_embed_fwd_kernel<bf16><<<g, 256, 0, st>>>(tokens, word, pos, type0, (bf16*)e, R, Tn, Hd);
  return check_launch("bert_embed_fwd_kernel");
}
int bert_embed_bwd(int dtype, const void* de, const long long* tokens, float* dword, float* dpos, float* dtype0,
                   long long R, int Tn, int Hd, cudaStream_t st) {
  unsigned g = grid_for((long long)Tn * Hd, 128);
  if (dtype == VALOR_DT_F32) bert_embed_bwd_kernel<float><<<g, 128, 0, st>>>((const float*)de, tokens, dword, dpos, dtype0, R, Tn, Hd);
  else bert_embed_bwd_kernel<bf16><<<g, 128, 0, st>>>((const bf16*)de, tokens, dword, dpos, dtype0, R, Tn, Hd);
  return check_launch("bert_embed_bwd_kernel");
}

// ---------------------------------------------------------------------------------------
// get_multimodal_forward_input_{video,audio} (modeling.py:485-502):
// out[b, row0 + f*X + x, :] = in[b,f,x,:] + frame_emb[f,:] + type_emb[:]
// `out` is the per-sample cross-attention source [B, S_total, Hd]; video rows first, audio after.
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void media_input_fwd_kernel(const T* __restrict__ in, const float* __restrict__ frame_emb,
                                       const float* __restrict__ type_emb, T* __restrict__ out, int B, int nf, int X,
                                       int Hd, int S_total, int row0) {
  const long long total = (long long)B * nf * X * Hd;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % Hd);
    long long t = idx / Hd;
    const int x = (int)(t % X); t /= X;
    const int f = (int)(t % nf);
    const long long b = t / nf;
    const float v = to_f(in[idx]) + frame_emb[(long long)f * Hd + c] + type_emb[c];
    out[(b * S_total + row0 + f * X + x) * Hd + c] = from_f<T>(v);
  }
}
// din = dout slice; dframe[f] += sum_{b,x}; dtype += sum_{b,f,x}
// grid (nf, B-chunks): each CTA walks its (b-chunk, f) rows with all threads across Hd (coalesced), keeps the
// column sums in registers and issues one atomicAdd per column.
template <typename T>
__global__ void __launch_bounds__(256)
media_input_bwd_kernel(const T* __restrict__ dout, T* __restrict__ din, float* __restrict__ dframe,
                       float* __restrict__ dtype, int B, int nf, int X, int Hd, int S_total, int row0, int b_per) {
  const int f = blockIdx.x;
  const int b0 = blockIdx.y * b_per;
  const int b1 = min(B, b0 + b_per);
  for (int c = threadIdx.x; c < Hd; c += blockDim.x) {
    float s = 0.f;
    for (int b = b0; b < b1; ++b)
      for (int x = 0; x < X; ++x) {
        const T v = dout[((long long)b * S_total + row0 + f * X + x) * Hd + c];
        s += to_f(v);
        din[(((long long)b * nf + f) * X + x) * Hd + c] = v;
      }
    if (dframe) atomicAdd(&dframe[(long long)f * Hd + c], s);
    if (dtype) atomicAdd(&dtype[c], s);
  }
}
int media_input_fwd(int dtype, const void* in, const float* frame_emb, const float* type_emb, void* out, int B, int nf,
                    int X, int Hd, int S_total, int row0, cudaStream_t st) {
  unsigned g = grid_for((long long)B * nf * X * Hd, 256);
  if (dtype == VALOR_DT_F32) media_input_fwd_kernel<float><<<g, 256, 0, st>>>((const float*)in, frame_emb, type_emb, (float*)out, B, nf, X, Hd, S_total, row0);
  else media_input_fwd_kernel<bf16><<<g, 256, 0, st>>>((const bf16*)in, frame_emb, type_emb, (bf16*)out, B, nf, X, Hd, S_total, row0);
  return check_launch("media_input_fwd_kernel");
}
int media_input_bwd(int dtype, const void* dout, void* din, float* dframe, float* dtype_emb, int B, int nf, int X,
                    int Hd, int S_total, int row0, cudaStream_t st) {
  const int b_per = B >= 64 ? 4 : (B >= 16 ? 2 : 1);
  dim3 g(nf, (B + b_per - 1) / b_per);
  if (dtype == VALOR_DT_F32) media_input_bwd_kernel<float><<<g, 256, 0, st>>>((const float*)dout, (float*)din, dframe, dtype_emb, B, nf, X, Hd, S_total, row0, b_per);
  else media_input_bwd_kernel<bf16><<<g, 256, 0, st>>>((const bf16*)dout, (bf16*)din, dframe, dtype_emb, B, nf, X, Hd, S_total, row0, b_per);
  return check_launch("media_input_bwd_kernel");
}

// ---------------------------------------------------------------------------------------
// PatchMerging gather (videoswin.py:261-265): x [B*D, H, W, C] -> y [B*D, H/2, W/2, 4C],
// channel blocks ordered (x0:h even,w even | x1:h odd,w even | x2:h even,w odd | x3:h odd,w odd).
// inverse = 1: y -> x (the gradient path; a pure permutation when H, W are even).
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void patch_merge_kernel(const T* __restrict__ src, T* __restrict__ dst, long long BD, int H, int W, int C, int inverse) {
  const int H2 = H / 2, W2 = W / 2, C4 = C / 4;
  const long long total = BD * H * W * C4;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C4) * 4;
    long long t = idx / C4;
    const int w = (int)(t % W); t /= W;
    const int h = (int)(t % H);
    const long long bd = t / H;
    const int blk = (w & 1) * 2 + (h & 1);
    const long long xo = ((bd * H + h) * W + w) * C + c;
    const long long yo = ((bd * H2 + h / 2) * W2 + w / 2) * (4LL * C) + (long long)blk * C + c;
    if (inverse) {
#pragma unroll
      for (int e = 0; e < 4; ++e) dst[xo + e] = src[yo + e];
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) dst[yo + e] = src[xo + e];
    }
  }
}
int patch_merge(int dtype, const void* src, void* dst, long long BD, int H, int W, int C, int inverse, cudaStream_t st) {
  VALOR_REQUIRE(H % 2 == 0 && W % 2 == 0 && C % 4 == 0, "patch_merge: H, W must be even (got %d x %d)", H, W);
  unsigned g = grid_for(BD * H * W * (C / 4), 256);
  if (dtype == VALOR_DT_F32) patch_merge_kernel<float><<<g, 256, 0, st>>>((const float*)src, (float*)dst, BD, H, W, C, inverse);
  else patch_merge_kernel<bf16><<<g, 256, 0, st>>>((const bf16*)src, (bf16*)dst, BD, H, W, C, inverse);
  return check_launch("patch_merge_kernel");
}

// ---------------------------------------------------------------------------------------
// mean over the middle axis: x [R, X, C] -> y [R, C] (pool_video_for_contra, modeling.py:389)
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void mean_pool_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, long long R, int X, int C) {
  const long long total = R * C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const long long r = idx / C;
    float s = 0.f;
    for (int i = 0; i < X; ++i) s += to_f(x[(r * X + i) * C + c]);
    y[idx] = from_f<T>(s / X);
  }
}
template <typename T>
__global__ void mean_pool_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, long long R, int X, int C) {
  const long long total = R * X * C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const long long r = idx / ((long long)X * C);
    dx[idx] = from_f<T>(to_f(dy[r * C + c]) / X);
  }
}
int mean_pool_fwd(int dtype, const void* x, void* y, long long R, int X, int C, cudaStream_t st) {
  unsigned g = grid_for(R * C, 256);
  if (dtype == VALOR_DT_F32) mean_pool_fwd_kernel<float><<<g, 256, 0, st>>>((const float*)x, (float*)y, R, X, C);
  else mean_pool_fwd_kernel<bf16><<<g, 256, 0, st>>>((const bf16*)x, (bf16*)y, R, X, C);
  return check_launch("mean_pool_fwd_kernel");
}
int mean_pool_bwd(int dtype, const void* dy, void* dx, long long R, int X, int C, cudaStream_t st) {
  unsigned g = grid_for(R * X * C, 256);
  if (dtype == VALOR_DT_F32) mean_pool_bwd_kernel<float><<<g, 256, 0, st>>>((const float*)dy, (float*)dx, R, X, C);
  else mean_pool_bwd_kernel<bf16><<<g, 256, 0, st>>>((const bf16*)dy, (bf16*)dx, R, X, C);
  return check_launch("mean_pool_bwd_kernel");
}

// ---------------------------------------------------------------------------------------
// bias gradient: db[n] += sum_m dy[m, n]   (dy row-major, pitch ld)
// ---------------------------------------------------------------------------------------
// Each warp reads 4 consecutive columns per lane (8-byte bf16 / 16-byte fp32 vectors): a CTA covers a
// 128-column panel and a row range; partial sums meet in shared memory, one atomicAdd per column.
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ dy, long long ld, float* __restrict__ db, long long M, int N, int rows_per_block) {
  __shared__ float sh[8][128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 128 + lane * 4;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > M) r1 = M;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  const bool vec = (c0 + 4 <= N) && (ld % 4 == 0) && ((((uintptr_t)dy) & 15) == 0);
  if (vec) {
    for (long long r = r0 + warp; r < r1; r += 8) {
      float f[4];
      if (sizeof(T) == 2) {
        const uint2 v = *(const uint2*)((const bf16*)dy + r * ld + c0);
        const __nv_bfloat162* h = (const __nv_bfloat162*)&v;
        const float2 x = __bfloat1622float2(h[0]), y = __bfloat1622float2(h[1]);
        f[0] = x.x; f[1] = x.y; f[2] = y.x; f[3] = y.y;
      } else {
        const float4 v = *(const float4*)((const float*)dy + r * ld + c0);
        f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
      }
      a0 += f[0]; a1 += f[1]; a2 += f[2]; a3 += f[3];
    }
  } else {
    for (long long r = r0 + warp; r < r1; r += 8) {
      if (c0 + 0 < N) a0 += to_f(dy[r * ld + c0 + 0]);
      if (c0 + 1 < N) a1 += to_f(dy[r * ld + c0 + 1]);
      if (c0 + 2 < N) a2 += to_f(dy[r * ld + c0 + 2]);
      if (c0 + 3 < N) a3 += to_f(dy[r * ld + c0 + 3]);
    }
  }
  sh[warp][lane * 4 + 0] = a0; sh[warp][lane * 4 + 1] = a1; sh[warp][lane * 4 + 2] = a2; sh[warp][lane * 4 + 3] = a3;
  __syncthreads();
  if (threadIdx.x < 128) {
    const int col = blockIdx.x * 128 + threadIdx.x;
    if (col < N) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) t += sh[i][threadIdx.x];
      atomicAdd(&db[col], t);
    }
  }
}
int colsum(int dtype, const void* dy, long long ld, float* db, long long M, int N, cudaStream_t st) {
  if (M == 0) return 0;
  const int cb = (N + 127) / 128;
  long long want_rb = ((long long)num_sms() * 8 + cb - 1) / cb;
  long long rows_per_block = (M + want_rb - 1) / want_rb;
  if (rows_per_block < 64) rows_per_block = 64;
  const long long rb = (M + rows_per_block - 1) / rows_per_block;
  VALOR_REQUIRE(rb <= 65535, "colsum: grid too large");
  dim3 grid(cb, (unsigned)rb);
  if (dtype == VALOR_DT_F32) colsum_kernel<float><<<grid, 256, 0, st>>>((const float*)dy, ld, db, M, N, (int)rows_per_block);
  else colsum_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)dy, ld, db, M, N, (int)rows_per_block);
  return check_launch("colsum_kernel");
}

// ---------------------------------------------------------------------------------------
// 2-D strided dtype casts, activation gradient, strided row copies
// ---------------------------------------------------------------------------------------
template <typename TS, typename TD>
__global__ void cast2d_kernel(const TS* __restrict__ src, long long sld, TD* __restrict__ dst, long long dld, long long R, long long C) {
  const long long total = R * C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / C, c = idx % C;
    dst[r * dld + c] = from_f<TD>(to_f(src[r * sld + c]));
  }
}
int cast2d(int src_dtype, int dst_dtype, const void* src, long long sld, void* dst, long long dld, long long R, long long C, cudaStream_t st) {
  if (R * C == 0) return 0;
  unsigned g = grid_for(R * C, 256);
  if (src_dtype == VALOR_DT_F32 && dst_dtype == VALOR_DT_BF16)
    cast2d_kernel<float, bf16><<<g, 256, 0, st>>>((const float*)src, sld, (bf16*)dst, dld, R, C);
  else if (src_dtype == VALOR_DT_BF16 && dst_dtype == VALOR_DT_F32)
    cast2d_kernel<bf16, float><<<g, 256, 0, st>>>((const bf16*)src, sld, (float*)dst, dld, R, C);
  else if (src_dtype == VALOR_DT_F32 && dst_dtype == VALOR_DT_F32)
    cast2d_kernel<float, float><<<g, 256, 0, st>>>((const float*)src, sld, (float*)dst, dld, R, C);
  else
    cast2d_kernel<bf16, bf16><<<g, 256, 0, st>>>((const bf16*)src, sld, (bf16*)dst, dld, R, C);
  return check_launch("cast2d_kernel");
}

// x [R, C] fp32 -> out [R, 3C] bf16 holding the two-term bf16 expansion x ~ hi + lo (|x - hi - lo| <= 2^-17 |x|) laid out
// so that ONE bf16 tensor-core GEMM with K = 3C evaluates  hi.hi' + hi.lo' + lo.hi'  (fp32-grade dot products: the
// fine-grained contrastive similarity of L2-normalised features, pretrain.py:200, would otherwise lose 2^-9 per operand):
//   side 0 (left operand):  [hi | hi | lo]        side 1 (right operand):  [hi | lo | hi]
__global__ void split_bf16x3_kernel(const float* __restrict__ x, long long xld, bf16* __restrict__ out, long long C, long long R, int side) {
  const long long n = R * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C, c = i - r * C;
    const float v = x[r * xld + c];
    const bf16 hi = __float2bfloat16_rn(v);
    const bf16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    bf16* o = out + r * 3 * C + c;
    o[0] = hi;
    o[C] = side == 0 ? hi : lo;
    o[2 * C] = side == 0 ? lo : hi;
  }
}
int split_bf16x3(const float* x, long long xld, void* out, long long R, long long C, int side, cudaStream_t st) {
  if (R * C == 0) return 0;
  long long g = (R * C + 255) / 256;
  if (g > (long long)num_sms() * 16) g = (long long)num_sms() * 16;
  split_bf16x3_kernel<<<(unsigned)g, 256, 0, st>>>(x, xld, (bf16*)out, C, R, side);
  return check_launch("split_bf16x3_kernel");
}

// dh = dy * act'(h)   (flat, same dtype)
template <typename T>
__global__ void act_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ h, T* __restrict__ dh, long long n, int act) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dh[i] = from_f<T>(to_f(dy[i]) * act_grad(to_f(h[i]), act));
}
int act_bwd(int dtype, const void* dy, const void* h, void* dh, long long n, int act, cudaStream_t st) {
  if (n == 0) return 0;
  unsigned g = grid_for(n, 256);
  if (dtype == VALOR_DT_F32) act_bwd_kernel<float><<<g, 256, 0, st>>>((const float*)dy, (const float*)h, (float*)dh, n, act);
  else act_bwd_kernel<bf16><<<g, 256, 0, st>>>((const bf16*)dy, (const bf16*)h, (bf16*)dh, n, act);
  return check_launch("act_bwd_kernel");
}

// dst[r*dld + c] (=|+=) src[r*sld + c]   — cls-token select (modeling.py:399) and its gradient
template <typename T>
__global__ void strided_rows_kernel(const T* __restrict__ src, long long sld, T* __restrict__ dst, long long dld,
                                    long long R, int C, int accumulate) {
  const long long total = R * C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const long long r = idx / C;
    const float v = to_f(src[r * sld + c]);
    T* d = dst + r * dld + c;
    *d = from_f<T>(accumulate ? to_f(*d) + v : v);
  }
}
int strided_rows(int dtype, const void* src, long long sld, void* dst, long long dld, long long R, int C, int accumulate, cudaStream_t st) {
  unsigned g = grid_for(R * C, 256);
  if (dtype == VALOR_DT_F32) strided_rows_kernel<float><<<g, 256, 0, st>>>((const float*)src, sld, (float*)dst, dld, R, C, accumulate);
  else strided_rows_kernel<bf16><<<g, 256, 0, st>>>((const bf16*)src, sld, (bf16*)dst, dld, R, C, accumulate);
  return check_launch("strided_rows_kernel");
}

}  // namespace valor
