// valor_b200 — index descriptors shared by the attention kernels.
#pragma once
#include "common.cuh"

namespace valor {

// BERT / AST attention (bert.py:244-340, transformer.py:115-130).
// Problem p = one sequence; scores scaled AFTER q.k^T (bert.py:273); additive mask is
// -10000 where the key token is padding, or (causal) where j > i  (bert.py:869-885).
struct MhaIndex {
  int max_nk;
  int Nq;
  const int* q_row0;   // [P] first query row of problem p, or null -> p*Nq
  const int* kv_row0;  // [P] first key/value row, or null -> p*max_nk
  const int* kv_len;   // [P] key count, or null -> max_nk
  const unsigned char* key_valid;  // [P, max_nk] 1 = real token, or null
  const unsigned char* causal;     // [P] 1 = lower-triangular, or null
  const int* q_key_range = nullptr;  // [Nq][2] query i only sees keys lo <= j < hi of its problem (others do not exist
                                     // for it: -inf, not -10000), or null.  Lets the three caption passes of one
                                     // sample -- same media tokens, different key subsets -- run as ONE problem.

  __device__ __forceinline__ int nk(int p) const { return kv_len ? kv_len[p] : max_nk; }
  __device__ __forceinline__ size_t qrow(int p, int i) const { return (size_t)(q_row0 ? q_row0[p] : p * Nq) + i; }
  __device__ __forceinline__ size_t krow(int p, int j) const { return (size_t)(kv_row0 ? kv_row0[p] : p * max_nk) + j; }
  __device__ __forceinline__ float qscale(float) const { return 1.f; }
  __device__ __forceinline__ float sscale(float s) const { return s; }
  __device__ __forceinline__ float add(int p, int, int i, int j) const {
    if (q_key_range && (j < q_key_range[2 * i] || j >= q_key_range[2 * i + 1])) return -INFINITY;
    bool ok = true;
    if (key_valid) ok = key_valid[(size_t)p * max_nk + j] != 0;
    if (causal && causal[p] && j > i) ok = false;
    return ok ? 0.f : -10000.f;
  }
  __device__ __forceinline__ void add_grad(float*, int, int, int, int, float) const {}
};

// VideoSwin window attention (videoswin.py:137-163) on the natural token order.
// Problem p = one (batch, window); q scaled BEFORE q.k^T (videoswin.py:143).
struct WindowIndex {
  int N, max_nk;
  int B, D, H, W;       // token grid
  int wd, wh, ww;       // effective window (get_window_size, videoswin.py:86-99)
  int sd, sh, sw;       // effective cyclic shift (0,0,0 for even blocks)
  int WD, WH, WW;       // configured window that sized the bias table (8,7,7)
  int heads;
  const float* table;   // relative_position_bias_table [(2WD-1)(2WH-1)(2WW-1), heads] fp32

  __device__ __forceinline__ int nk(int) const { return N; }
  __device__ __forceinline__ void coords(int p, int i, int& cd, int& ch, int& cw, int& b) const {
    const int nWw = W / ww, nWh = H / wh, nWd = D / wd;
    int t = p;
    const int iw = t % nWw; t /= nWw;
    const int ih = t % nWh; t /= nWh;
    const int id = t % nWd; b = t / nWd;
    cd = id * wd + i / (wh * ww);
    ch = ih * wh + (i / ww) % wh;
    cw = iw * ww + i % ww;
  }
  __device__ __forceinline__ size_t row(int p, int i) const {
    int cd, ch, cw, b;
    coords(p, i, cd, ch, cw, b);
    // shifted_x = roll(x, -shift)  =>  shifted[c] = x[(c + shift) mod size]   (videoswin.py:206)
    int d = cd + sd; if (d >= D) d -= D;
    int h = ch + sh; if (h >= H) h -= H;
    int w = cw + sw; if (w >= W) w -= W;
    return (((size_t)b * D + d) * H + h) * W + w;
  }
  __device__ __forceinline__ size_t qrow(int p, int i) const { return row(p, i); }
  __device__ __forceinline__ size_t krow(int p, int j) const { return row(p, j); }
  __device__ __forceinline__ float qscale(float s) const { return s; }
  __device__ __forceinline__ float sscale(float) const { return 1.f; }
  __device__ __forceinline__ int region(int c, int S, int w, int s) const {
    // compute_mask slices (videoswin.py:276-278): [0,S-w) | [S-w,S-s) | [S-s,S)
    if (s == 0) return 0;
    return c < S - w ? 0 : (c < S - s ? 1 : 2);
  }
  __device__ __forceinline__ int rel(int i, int j) const {
    const int di = i / (wh * ww) - j / (wh * ww);
    const int hi = (i / ww) % wh - (j / ww) % wh;
    const int wi = i % ww - j % ww;
    return (di + WD - 1) * (2 * WH - 1) * (2 * WW - 1) + (hi + WH - 1) * (2 * WW - 1) + (wi + WW - 1);
  }
  __device__ __forceinline__ float add(int p, int h, int i, int j) const {
    float a = table[(size_t)rel(i, j) * heads + h];
    if (sd | sh | sw) {
      int cd, ch, cw, b, ed, eh, ew;
      coords(p, i, cd, ch, cw, b);
      coords(p, j, ed, eh, ew, b);
      const int ri = region(cd, D, wd, sd) * 9 + region(ch, H, wh, sh) * 3 + region(cw, W, ww, sw);
      const int rj = region(ed, D, wd, sd) * 9 + region(eh, H, wh, sh) * 3 + region(ew, W, ww, sw);
      if (ri != rj) a += -100.0f;  // videoswin.py:284
    }
    return a;
  }
  __device__ __forceinline__ void add_grad(float* dtable, int, int h, int i, int j, float ds) const {
    if (dtable) atomicAdd(&dtable[(size_t)rel(i, j) * heads + h], ds);
  }
};

int mha_ref_fwd(int dtype, const MhaIndex& ix, const void* Q, const void* K, const void* V, long long ldq,
                long long ldk, long long ldv, void* O, long long ldo, float* lse, int P, int H, int hd, int Nq,
                float scale, cudaStream_t st);
int mha_ref_bwd(int dtype, const MhaIndex& ix, const void* Q, const void* K, const void* V, const void* O,
                const void* dO, long long ldq, long long ldk, long long ldv, long long ldo, const float* lse, void* dQ,
                long long lddq, float* dK, float* dV, long long lddk, long long lddv, int P, int H, int hd, int Nq,
                float scale, cudaStream_t st);
int window_ref_fwd(int dtype, const WindowIndex& ix, const void* qkv, long long ld, void* O, long long ldo, float* lse,
                   int P, int H, int hd, float scale, cudaStream_t st);
int window_ref_bwd(int dtype, const WindowIndex& ix, const void* qkv, long long ld, const void* O, const void* dO,
                   long long ldo, const float* lse, void* dQ, long long lddq, float* dK, float* dV, long long lddkv,
                   float* dtable, int P, int H, int hd, float scale, cudaStream_t st);

}  // namespace valor
