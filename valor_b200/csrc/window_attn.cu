// valor_b200 — VideoSwin shifted-window attention, one CTA per (window, head).
//
// A 3-D window holds at most 8*7*7 = 392 tokens of 32 channels per head, so a whole attention
// problem (Q, K, V and, backward, dO) fits in shared memory: 4 x 196 x 64 B = 50 KB for the
// pre-training geometry.  The CTA builds the window's index / bias / mask tables ONCE, gathers the
// token rows once (cyclic shift + window partition folded into the row index, videoswin.py:206-216),
// and then every warp runs independently on its own 16-row blocks with no CTA barrier:
//
//   forward : warp owns 16 queries, sweeps all keys with an online softmax (S, P stay in registers)
//   backward: units 0..nrb-1   = 16-query blocks -> dQ (and the relative-position-bias gradient),
//             units nrb..2nrb-1 = 16-key blocks  -> dK, dV (S^T = K.Q^T so P^T / dS^T are produced
//             directly in A-operand layout); units are dealt round-robin to the warps
//
// Relative-position bias (videoswin.py:113-127,150-153) and the -100 shift mask (videoswin.py:272-285)
// are evaluated per score from one 32-bit info word per token; keys are enumerated (w,d,h) so the 32
// lanes of an MMA fragment touch 32 distinct bias slots.  The bias gradient is accumulated in shared
// memory with native 32-bit integer atomics (fp32 shared atomics are compare-and-swap loops) at a
// per-CTA power-of-two scale derived from max|dO|.max|V|, then folded into the global fp32 table.
#include "common.cuh"
#include "attention.cuh"
#include "mma_utils.cuh"

namespace valor {

struct WinParams {
  const bf16* qkv; long long ld;   // [tokens, 3C]: Q | K | V
  bf16* O; long long ldo;          // forward output / backward: saved output
  float* lse;                      // [P, H, N] natural-log sum-exp
  const bf16* dO;                  // backward, pitch ldo
  bf16* dqkv; long long lddqkv;    // [tokens, 3C]
  float* dtable;                   // [(2WD-1)(2WH-1)(2WW-1), heads] fp32, accumulated
  float scale;
  int heads;
  int NP;                          // N rounded up to 16
  int n_used, maxcode, center;     // bias slots reachable from this window: [center-maxcode, center+maxcode]
  WindowIndex win;
};

template <int HD> struct WinCfg {
  static constexpr int PITCH = HD * 2 + 16;  // bytes per staged row: 16-byte skew keeps ldmatrix conflict-free
  static constexpr int CH = HD / 8;          // 16-byte chunks per row
};

// exact i / d for 0 <= i < 4096, 1 <= d <= 4096 (multiply-shift with a rounded-up reciprocal)
__device__ __forceinline__ int small_div(int i, int inv) { return (int)(((unsigned)i * (unsigned)inv) >> 20); }
__host__ __device__ __forceinline__ int small_inv(int d) { return (int)(((1u << 20) + d - 1) / d); }

struct WinTables {
  uint32_t qrow_s, krow_s, qinfo_s, kinfo_s, tab2_s;  // shared-space addresses
};

// Per-token words:
//   bits [0,16)  : byte offset into the CTA's bias slice (query: 4*(code+maxcode); key: 4*code)
//   bits [16,24) : shift-mask region id (compute_mask, videoswin.py:272-285)
//   bit 31       : padding row (index >= N)
template <int HD>
__device__ __forceinline__ void win_build_tables(const WinParams& P, int p, int h, int* qrow, int* krow, uint32_t* qinfo,
                                                 uint32_t* kinfo, float* tab2) {
  const WindowIndex& ix = P.win;
  const int nWw = ix.W / ix.ww, nWh = ix.H / ix.wh, nWd = ix.D / ix.wd;
  int tq = p;
  const int iw = tq % nWw; tq /= nWw;
  const int ih = tq % nWh; tq /= nWh;
  const int id = tq % nWd;
  const int b = tq / nWd;
  const int od = id * ix.wd, oh = ih * ix.wh, ow = iw * ix.ww;
  const int cW = 2 * ix.WW - 1, cH = (2 * ix.WH - 1) * cW;
  const int hw = ix.wh * ix.ww, dh = ix.wd * ix.wh;
  const int inv_hw = small_inv(hw), inv_ww = small_inv(ix.ww), inv_wh = small_inv(ix.wh), inv_dh = small_inv(dh);
  const bool shifted = (ix.sd | ix.sh | ix.sw) != 0;
  for (int i = threadIdx.x; i < P.NP; i += blockDim.x) {
    if (i < ix.N) {
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        int ld, lh, lw;
        if (which == 0) {            // queries: natural (d,h,w) order
          ld = small_div(i, inv_hw);
          const int rem = i - ld * hw;
          lh = small_div(rem, inv_ww);
          lw = rem - lh * ix.ww;
        } else {                     // keys: (w,d,h) order, h fastest
          lw = small_div(i, inv_dh);
          const int rem = i - lw * dh;
          ld = small_div(rem, inv_wh);
          lh = rem - ld * ix.wh;
        }
        const int cd = od + ld, ch = oh + lh, cw = ow + lw;
        int d = cd + ix.sd; if (d >= ix.D) d -= ix.D;   // shifted[c] = x[(c + shift) mod size]  (videoswin.py:206)
        int hh = ch + ix.sh; if (hh >= ix.H) hh -= ix.H;
        int w = cw + ix.sw; if (w >= ix.W) w -= ix.W;
        const int row = ((b * ix.D + d) * ix.H + hh) * ix.W + w;
        uint32_t reg = 0;
        if (shifted) reg = (uint32_t)(ix.region(cd, ix.D, ix.wd, ix.sd) * 9 + ix.region(ch, ix.H, ix.wh, ix.sh) * 3 +
                                      ix.region(cw, ix.W, ix.ww, ix.sw));
        const int code = ld * cH + lh * cW + lw;
        if (which == 0) { qrow[i] = row; qinfo[i] = (uint32_t)(4 * (code + P.maxcode)) | (reg << 16); }
        else            { krow[i] = row; kinfo[i] = (uint32_t)(4 * code) | (reg << 16); }
      }
    } else {
      qrow[i] = -1; krow[i] = -1;
      qinfo[i] = (uint32_t)(4 * P.maxcode) | 0x80000000u;
      kinfo[i] = 0x80000000u;
    }
  }
  const int r0 = P.center - P.maxcode;
  for (int r = threadIdx.x; r < P.n_used; r += blockDim.x) tab2[r] = ix.table[(size_t)(r0 + r) * ix.heads + h] * LOG2E;
}

// gather NP rows x HD bf16 into a padded shared tile (cp.async, zero-fill for padding rows)
template <int HD>
__device__ __forceinline__ void win_load_rows(unsigned char* dst, const bf16* src, long long ld, int col0, const int* rows, int NP) {
  constexpr int PITCH = WinCfg<HD>::PITCH, CH = WinCfg<HD>::CH;
  for (int c = threadIdx.x; c < NP * CH; c += blockDim.x) {
    const int r = c / CH, ch = c % CH;
    const int gr = rows[r];
    const bf16* g = src + (size_t)(gr < 0 ? 0 : gr) * ld + col0 + ch * 8;
    cp_async16(s_u32(dst + r * PITCH + ch * 16), g, gr < 0 ? 0 : 16);
  }
}

// score in the log2 domain: s*scale*log2e + bias (+ mask)
__device__ __forceinline__ float win_score(float s, float sc2, uint32_t qaddr, uint32_t qreg, uint32_t kw, bool shifted) {
  float v = fmaf(s, sc2, lds_f32(qaddr - (kw & 0xffffu)));
  if (shifted && ((qreg ^ kw) & 0x00ff0000u)) v += M100_2;
  return v;
}

static inline size_t win_smem_bytes(int HD, int NP, int n_used, bool bwd) {
  const size_t pitch = HD * 2 + 16;
  size_t b = (size_t)(bwd ? 4 : 3) * NP * pitch;   // Q K V (dO)
  b += (size_t)4 * NP * 4;                         // qrow krow qinfo kinfo
  b += (size_t)n_used * 4;                         // bias slice
  if (bwd) b += (size_t)n_used * 4 + (size_t)2 * NP * 4 + 16;   // bias-gradient slots, lse, delta, scale words
  return b + 32;
}

// ==========================================================================================
// forward
// ==========================================================================================
template <int HD>
__global__ void __launch_bounds__(256, 2)
window_fwd_kernel(WinParams P) {
  constexpr int PITCH = WinCfg<HD>::PITCH;
  extern __shared__ __align__(16) unsigned char smem[];
  const int NP = P.NP;
  unsigned char* Qs = smem;
  unsigned char* Ks = Qs + NP * PITCH;
  unsigned char* Vs = Ks + NP * PITCH;
  int* qrow = (int*)(Vs + NP * PITCH);
  int* krow = qrow + NP;
  uint32_t* qinfo = (uint32_t*)(krow + NP);
  uint32_t* kinfo = qinfo + NP;
  float* tab2 = (float*)(kinfo + NP);
  const int p = blockIdx.x, h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  const int nwarps = blockDim.x >> 5;
  const int C = P.heads * HD, col0 = h * HD;
  win_build_tables<HD>(P, p, h, qrow, krow, qinfo, kinfo, tab2);
  __syncthreads();
  win_load_rows<HD>(Qs, P.qkv, P.ld, col0, qrow, NP);
  win_load_rows<HD>(Ks, P.qkv + C, P.ld, col0, krow, NP);
  win_load_rows<HD>(Vs, P.qkv + 2 * C, P.ld, col0, krow, NP);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  const int N = P.win.N;
  const bool shifted = (P.win.sd | P.win.sh | P.win.sw) != 0;
  const float sc2 = P.scale * LOG2E;
  const uint32_t tab2_s = s_u32(tab2), kinfo_s = s_u32(kinfo);
  const int m8 = lane >> 3, r8 = lane & 7;
  const int nrb = NP >> 4;
  const int nkb = (N + 63) >> 6;
  for (int rb = warp; rb < nrb; rb += nwarps) {
    uint32_t qf[HD / 16][4];
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks)
      ldsm_x4(qf[ks], s_u32(Qs + (rb * 16 + (m8 & 1) * 8 + r8) * PITCH + (ks * 16 + (m8 >> 1) * 8) * 2));
    const int i0 = rb * 16 + g;
    const uint32_t qw0 = qinfo[i0], qw1 = qinfo[i0 + 8];
    const uint32_t qaddr[2] = {tab2_s + (qw0 & 0xffffu), tab2_s + (qw1 & 0xffffu)};
    const uint32_t qreg[2] = {qw0, qw1};
    float o[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};
    for (int kb = 0; kb < nkb; ++kb) {
      const int k0 = kb * 64;
      const int npair = min(4, (N - k0 + 15) >> 4);   // 16-key pairs of n-tiles with at least one real key
      const unsigned char* Kb = Ks + k0 * PITCH;
      const unsigned char* Vb = Vs + k0 * PITCH;
      float s[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks)
#pragma unroll
        for (int pr = 0; pr < 4; ++pr)
          if (pr < npair) {
            uint32_t b[4];
            ldsm_x4(b, s_u32(Kb + ((pr * 2 + (m8 >> 1)) * 8 + r8) * PITCH + (ks * 16 + (m8 & 1) * 8) * 2));
            mma16816(s[pr * 2], qf[ks], b);
            mma16816(s[pr * 2 + 1], qf[ks], b + 2);
          }
      float mnew[2] = {mrow[0], mrow[1]};
      const bool ragged = k0 + npair * 16 > N;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
        if ((nt >> 1) < npair) {
          const uint2 kj = lds_v2u32(kinfo_s + 4u * (uint32_t)(k0 + nt * 8 + t4 * 2));
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const uint32_t kw = (e & 1) ? kj.y : kj.x;
            float v = win_score(s[nt][e], sc2, qaddr[e >> 1], qreg[e >> 1], kw, shifted);
            if (ragged && (kw >> 31)) v = -INFINITY;
            s[nt][e] = v;
            mnew[e >> 1] = fmaxf(mnew[e >> 1], v);
          }
        }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        mnew[r] = fmaxf(mnew[r], __shfl_xor_sync(0xffffffffu, mnew[r], 1));
        mnew[r] = fmaxf(mnew[r], __shfl_xor_sync(0xffffffffu, mnew[r], 2));
      }
      float corr[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {   // every row sees at least one real key in block 0, so mnew is finite
        corr[r] = fast_exp2(mrow[r] - mnew[r]);
        mrow[r] = mnew[r];
        lrow[r] *= corr[r];
      }
#pragma unroll
      for (int i = 0; i < HD / 8; ++i) { o[i][0] *= corr[0]; o[i][1] *= corr[0]; o[i][2] *= corr[1]; o[i][3] *= corr[1]; }
      uint32_t pf[4][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
        if ((nt >> 1) < npair) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float pv = fast_exp2(s[nt][e] - mnew[e >> 1]);
            s[nt][e] = pv;
            lrow[e >> 1] += pv;
          }
          pf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(s[nt][0], s[nt][1]);
          pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(s[nt][2], s[nt][3]);
        }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        if (kk < npair) {
#pragma unroll
          for (int dt = 0; dt < HD / 8; dt += 2) {
            uint32_t b[4];
            ldsm_x4_t(b, s_u32(Vb + (kk * 16 + (m8 & 1) * 8 + r8) * PITCH + ((dt + (m8 >> 1)) * 8) * 2));
            mma16816(o[dt], pf[kk], b);
            mma16816(o[dt + 1], pf[kk], b + 2);
          }
        }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 1);
      lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 2);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int i = i0 + r * 8;
      if (i < N) {
        const float inv = 1.f / lrow[r];
        bf16* dst = P.O + (size_t)qrow[i] * P.ldo + col0;
#pragma unroll
        for (int dt = 0; dt < HD / 8; ++dt)
          *(uint32_t*)(dst + dt * 8 + t4 * 2) = pack_bf16(o[dt][r * 2] * inv, o[dt][r * 2 + 1] * inv);
        if (t4 == 0) P.lse[((size_t)p * P.heads + h) * N + i] = (mrow[r] + log2f(lrow[r])) * LN2;
      }
    }
  }
}

// ==========================================================================================
// backward
// ==========================================================================================
template <int HD>
__global__ void __launch_bounds__(256, 2)
window_bwd_kernel(WinParams P) {
  constexpr int PITCH = WinCfg<HD>::PITCH, CH = WinCfg<HD>::CH;
  extern __shared__ __align__(16) unsigned char smem[];
  const int NP = P.NP;
  unsigned char* Qs = smem;
  unsigned char* Ks = Qs + NP * PITCH;
  unsigned char* Vs = Ks + NP * PITCH;
  unsigned char* dOs = Vs + NP * PITCH;
  int* qrow = (int*)(dOs + NP * PITCH);
  int* krow = qrow + NP;
  uint32_t* qinfo = (uint32_t*)(krow + NP);
  uint32_t* kinfo = qinfo + NP;
  float* tab2 = (float*)(kinfo + NP);
  int* dtab = (int*)(tab2 + P.n_used);
  float* lse_s = (float*)(dtab + P.n_used);   // log2 domain; +inf for padding rows
  float* del_s = lse_s + NP;
  int* mx = (int*)(del_s + NP);               // [0] max|dO row|^2, [1] max|V row|^2 (non-negative floats as ints)
  const int p = blockIdx.x, h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  const int nwarps = blockDim.x >> 5;
  const int C = P.heads * HD, col0 = h * HD;
  const int N = P.win.N;
  win_build_tables<HD>(P, p, h, qrow, krow, qinfo, kinfo, tab2);
  for (int i = threadIdx.x; i < P.n_used; i += blockDim.x) dtab[i] = 0;
  if (threadIdx.x < 2) mx[threadIdx.x] = 0;
  __syncthreads();
  win_load_rows<HD>(Qs, P.qkv, P.ld, col0, qrow, NP);
  win_load_rows<HD>(Ks, P.qkv + C, P.ld, col0, krow, NP);
  win_load_rows<HD>(Vs, P.qkv + 2 * C, P.ld, col0, krow, NP);
  win_load_rows<HD>(dOs, P.dO, P.ldo, col0, qrow, NP);
  cp_async_commit();
  for (int i = threadIdx.x; i < NP; i += blockDim.x)
    lse_s[i] = i < N ? P.lse[((size_t)p * P.heads + h) * N + i] * LOG2E : INFINITY;
  cp_async_wait<0>();
  __syncthreads();
  {
    // delta_i = dO_i . O_i (O rows straight from global), plus the row-norm maxima that size the fixed-point scale
    float mdo = 0.f, mv = 0.f;
    for (int c = threadIdx.x; c < NP * CH; c += blockDim.x) {   // NP*CH is a multiple of 32: whole warps iterate
      const int r = c / CH, ch = c % CH;
      const int gr = qrow[r];
      float d = 0.f, n2 = 0.f, v2 = 0.f;
      const uint4 a = *(const uint4*)(dOs + r * PITCH + ch * 16);
      const uint4 vv = *(const uint4*)(Vs + r * PITCH + ch * 16);
      uint4 o4 = make_uint4(0, 0, 0, 0);
      if (gr >= 0) o4 = *(const uint4*)(P.O + (size_t)gr * P.ldo + col0 + ch * 8);
      const __nv_bfloat162* pa = (const __nv_bfloat162*)&a;
      const __nv_bfloat162* po = (const __nv_bfloat162*)&o4;
      const __nv_bfloat162* pv = (const __nv_bfloat162*)&vv;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 fa = __bfloat1622float2(pa[j]), fo = __bfloat1622float2(po[j]), fv = __bfloat1622float2(pv[j]);
        d += fa.x * fo.x + fa.y * fo.y;
        n2 += fa.x * fa.x + fa.y * fa.y;
        v2 += fv.x * fv.x + fv.y * fv.y;
      }
#pragma unroll
      for (int sft = 1; sft < CH; sft <<= 1) {
        d += __shfl_xor_sync(0xffffffffu, d, sft);
        n2 += __shfl_xor_sync(0xffffffffu, n2, sft);
        v2 += __shfl_xor_sync(0xffffffffu, v2, sft);
      }
      if (ch == 0) del_s[r] = d;
      mdo = fmaxf(mdo, n2);
      mv = fmaxf(mv, v2);
    }
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) {
      mdo = fmaxf(mdo, __shfl_xor_sync(0xffffffffu, mdo, sft));
      mv = fmaxf(mv, __shfl_xor_sync(0xffffffffu, mv, sft));
    }
    if (lane == 0) {
      atomicMax(&mx[0], __float_as_int(mdo));
      atomicMax(&mx[1], __float_as_int(mv));
    }
  }
  __syncthreads();
  // |dS_ij| = |p (dp - delta)| <= 2 |dO_i| max|V|; at most N contributions share one bias slot in this CTA.
  // fixed-point scale: the largest power of two with  N * 2 |dO|max |V|max * scale < 2^30
  float fx_scale, fx_inv;
  {
    const float bound = 2.f * (float)N * sqrtf(__int_as_float(mx[0]) * __int_as_float(mx[1]));
    int e = 0;
    if (bound > 0.f) { (void)frexpf(bound, &e); }   // bound = m * 2^e, m in [0.5, 1)
    e = max(-60, min(60, 30 - e));
    fx_scale = exp2f((float)e);
    fx_inv = exp2f((float)-e);
  }
  const bool shifted = (P.win.sd | P.win.sh | P.win.sw) != 0;
  const float sc = P.scale, sc2 = P.scale * LOG2E;
  const uint32_t tab2_s = s_u32(tab2), kinfo_s = s_u32(kinfo), qinfo_s = s_u32(qinfo), dtab_s = s_u32(dtab);
  const uint32_t lse_ss = s_u32(lse_s), del_ss = s_u32(del_s);
  const int m8 = lane >> 3, r8 = lane & 7;
  const int nrb = NP >> 4;
  const int nhalf = (N + 31) >> 5;   // 32-wide sweeps over the other dimension (rows >= N are zero / masked)
  const bool want_dtab = P.dtable != nullptr;

  for (int unit = warp; unit < 2 * nrb; unit += nwarps) {
    if (unit < nrb) {
      // ------------------------------ dQ for 16 queries ------------------------------
      const int rb = unit;
      uint32_t qf[HD / 16][4], dof[HD / 16][4];
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks) {
        ldsm_x4(qf[ks], s_u32(Qs + (rb * 16 + (m8 & 1) * 8 + r8) * PITCH + (ks * 16 + (m8 >> 1) * 8) * 2));
        ldsm_x4(dof[ks], s_u32(dOs + (rb * 16 + (m8 & 1) * 8 + r8) * PITCH + (ks * 16 + (m8 >> 1) * 8) * 2));
      }
      const int i0 = rb * 16 + g;
      const uint32_t qw0 = qinfo[i0], qw1 = qinfo[i0 + 8];
      const uint32_t qoff[2] = {qw0 & 0xffffu, qw1 & 0xffffu};
      const uint32_t qreg[2] = {qw0, qw1};
      const float lse2[2] = {lse_s[i0], lse_s[i0 + 8]};   // +inf on padding rows -> p = 0
      const float del[2] = {del_s[i0], del_s[i0 + 8]};
      float dq[HD / 8][4];
#pragma unroll
      for (int i = 0; i < HD / 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
#pragma unroll 1
      for (int hb = 0; hb < nhalf; ++hb) {
        const int k0 = hb * 32;
        const unsigned char* Kb = Ks + k0 * PITCH;
        const unsigned char* Vb = Vs + k0 * PITCH;
        const int npair = min(2, (N - k0 + 15) >> 4);
        float s[4][4], dp[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) s[nt][e] = dp[nt][e] = 0.f;
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks)
#pragma unroll
          for (int pr = 0; pr < 2; ++pr)
            if (pr < npair) {
              uint32_t b[4];
              ldsm_x4(b, s_u32(Kb + ((pr * 2 + (m8 >> 1)) * 8 + r8) * PITCH + (ks * 16 + (m8 & 1) * 8) * 2));
              mma16816(s[pr * 2], qf[ks], b);
              mma16816(s[pr * 2 + 1], qf[ks], b + 2);
              ldsm_x4(b, s_u32(Vb + ((pr * 2 + (m8 >> 1)) * 8 + r8) * PITCH + (ks * 16 + (m8 & 1) * 8) * 2));
              mma16816(dp[pr * 2], dof[ks], b);
              mma16816(dp[pr * 2 + 1], dof[ks], b + 2);
            }
        uint32_t dsf[2][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          float ds[4] = {0.f, 0.f, 0.f, 0.f};
          if ((nt >> 1) < npair) {
            const uint2 kj = lds_v2u32(kinfo_s + 4u * (uint32_t)(k0 + nt * 8 + t4 * 2));
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int r = e >> 1;
              const uint32_t kw = (e & 1) ? kj.y : kj.x;
              const float v = win_score(s[nt][e], sc2, tab2_s + qoff[r], qreg[r], kw, shifted);
              float pr_ = fast_exp2(v - lse2[r]);
              if (kw >> 31) pr_ = 0.f;
              const float d = pr_ * (dp[nt][e] - del[r]);
              if (want_dtab) {
                const int q = __float2int_rn(d * fx_scale);
                asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(dtab_s + qoff[r] - (kw & 0xffffu)), "r"(q) : "memory");
              }
              ds[e] = d * sc;
            }
          }
          dsf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(ds[0], ds[1]);
          dsf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(ds[2], ds[3]);
        }
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
          if (kk < npair) {
#pragma unroll
            for (int dt = 0; dt < HD / 8; dt += 2) {
              uint32_t b[4];
              ldsm_x4_t(b, s_u32(Kb + (kk * 16 + (m8 & 1) * 8 + r8) * PITCH + ((dt + (m8 >> 1)) * 8) * 2));
              mma16816(dq[dt], dsf[kk], b);
              mma16816(dq[dt + 1], dsf[kk], b + 2);
            }
          }
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int i = i0 + r * 8;
        if (i < N) {
          bf16* dst = P.dqkv + (size_t)qrow[i] * P.lddqkv + col0;
#pragma unroll
          for (int dt = 0; dt < HD / 8; ++dt)
            *(uint32_t*)(dst + dt * 8 + t4 * 2) = pack_bf16(dq[dt][r * 2], dq[dt][r * 2 + 1]);
        }
      }
    } else {
      // ------------------------------ dK, dV for 16 keys ------------------------------
      const int jb = unit - nrb;
      uint32_t kf[HD / 16][4], vf[HD / 16][4];
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks) {
        ldsm_x4(kf[ks], s_u32(Ks + (jb * 16 + (m8 & 1) * 8 + r8) * PITCH + (ks * 16 + (m8 >> 1) * 8) * 2));
        ldsm_x4(vf[ks], s_u32(Vs + (jb * 16 + (m8 & 1) * 8 + r8) * PITCH + (ks * 16 + (m8 >> 1) * 8) * 2));
      }
      const int j0 = jb * 16 + g;
      const uint32_t kinf[2] = {kinfo[j0], kinfo[j0 + 8]};   // padding keys: K/V rows are zero, results dropped
      float dk[HD / 8][4], dv[HD / 8][4];
#pragma unroll
      for (int i = 0; i < HD / 8; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) dk[i][e] = dv[i][e] = 0.f;
#pragma unroll 1
      for (int hb = 0; hb < nhalf; ++hb) {
        const int q0 = hb * 32;
        const unsigned char* Qb = Qs + q0 * PITCH;
        const unsigned char* dOb = dOs + q0 * PITCH;
        const int npair = min(2, (N - q0 + 15) >> 4);
        float s[4][4], dp[4][4];   // rows = keys (g, g+8), cols = queries
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) s[nt][e] = dp[nt][e] = 0.f;
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks)
#pragma unroll
          for (int pr = 0; pr < 2; ++pr)
            if (pr < npair) {
              uint32_t b[4];
              ldsm_x4(b, s_u32(Qb + ((pr * 2 + (m8 >> 1)) * 8 + r8) * PITCH + (ks * 16 + (m8 & 1) * 8) * 2));
              mma16816(s[pr * 2], kf[ks], b);
              mma16816(s[pr * 2 + 1], kf[ks], b + 2);
              ldsm_x4(b, s_u32(dOb + ((pr * 2 + (m8 >> 1)) * 8 + r8) * PITCH + (ks * 16 + (m8 & 1) * 8) * 2));
              mma16816(dp[pr * 2], vf[ks], b);
              mma16816(dp[pr * 2 + 1], vf[ks], b + 2);
            }
        uint32_t pf[2][4], dsf[2][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          float pv[4] = {0.f, 0.f, 0.f, 0.f}, ds[4] = {0.f, 0.f, 0.f, 0.f};
          if ((nt >> 1) < npair) {
            const uint32_t ib = (uint32_t)(q0 + nt * 8 + t4 * 2);   // two consecutive queries (columns of S^T)
            const uint2 qi2 = lds_v2u32(qinfo_s + 4u * ib);
            const float2 l2 = lds_v2f32(lse_ss + 4u * ib);
            const float2 d2 = lds_v2f32(del_ss + 4u * ib);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int r = e >> 1;
              const uint32_t qw = (e & 1) ? qi2.y : qi2.x;
              const float v = win_score(s[nt][e], sc2, tab2_s + (qw & 0xffffu), qw, kinf[r], shifted);
              const float pr_ = fast_exp2(v - ((e & 1) ? l2.y : l2.x));   // padding queries carry lse = +inf -> 0
              pv[e] = pr_;
              ds[e] = pr_ * (dp[nt][e] - ((e & 1) ? d2.y : d2.x)) * sc;
            }
          }
          pf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(pv[0], pv[1]);
          pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(pv[2], pv[3]);
          dsf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(ds[0], ds[1]);
          dsf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(ds[2], ds[3]);
        }
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)   // contraction over the 16-query pairs of this sweep
          if (kk < npair) {
#pragma unroll
            for (int dt = 0; dt < HD / 8; dt += 2) {
              uint32_t b[4];
              ldsm_x4_t(b, s_u32(dOb + (kk * 16 + (m8 & 1) * 8 + r8) * PITCH + ((dt + (m8 >> 1)) * 8) * 2));
              mma16816(dv[dt], pf[kk], b);
              mma16816(dv[dt + 1], pf[kk], b + 2);
              ldsm_x4_t(b, s_u32(Qb + (kk * 16 + (m8 & 1) * 8 + r8) * PITCH + ((dt + (m8 >> 1)) * 8) * 2));
              mma16816(dk[dt], dsf[kk], b);
              mma16816(dk[dt + 1], dsf[kk], b + 2);
            }
          }
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int j = j0 + r * 8;
        if (j < N) {
          bf16* dstk = P.dqkv + (size_t)krow[j] * P.lddqkv + C + col0;
          bf16* dstv = dstk + C;
#pragma unroll
          for (int dt = 0; dt < HD / 8; ++dt) {
            *(uint32_t*)(dstk + dt * 8 + t4 * 2) = pack_bf16(dk[dt][r * 2], dk[dt][r * 2 + 1]);
            *(uint32_t*)(dstv + dt * 8 + t4 * 2) = pack_bf16(dv[dt][r * 2], dv[dt][r * 2 + 1]);
          }
        }
      }
    }
  }
  if (want_dtab) {
    __syncthreads();
    const int r0 = P.center - P.maxcode;
    for (int r = threadIdx.x; r < P.n_used; r += blockDim.x) {
      const int q = dtab[r];
      if (q != 0) atomicAdd(&P.dtable[(size_t)(r0 + r) * P.win.heads + h], (float)q * fx_inv);
    }
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static int win_setup(WinParams& P, const WindowIndex& ix, int H, int hd, int& nthreads) {
  P.win = ix;
  P.heads = H;
  P.NP = (ix.N + 15) / 16 * 16;
  const int cW = 2 * ix.WW - 1, cH = (2 * ix.WH - 1) * cW;
  P.center = (ix.WD - 1) * cH + (ix.WH - 1) * cW + (ix.WW - 1);
  P.maxcode = (ix.wd - 1) * cH + (ix.wh - 1) * cW + (ix.ww - 1);
  P.n_used = 2 * P.maxcode + 1;
  const int nrb = P.NP / 16;
  const int per = (nrb + 7) / 8;                 // row blocks per warp with at most 8 warps
  const int nwarps = (nrb + per - 1) / per;
  nthreads = nwarps * 32;
  VALOR_REQUIRE(ix.N <= 4095 && 4 * (2 * P.maxcode + 1) < 65536, "window_attn: window too large for the packed tables");
  (void)hd;
  return 0;
}

bool window_cta_eligible(const WindowIndex& ix, int hd) {
  const int NP = (ix.N + 15) / 16 * 16;
  const int cW = 2 * ix.WW - 1, cH = (2 * ix.WH - 1) * cW;
  const int n_used = 2 * ((ix.wd - 1) * cH + (ix.wh - 1) * cW + (ix.ww - 1)) + 1;
  return hd == 32 && win_smem_bytes(32, NP, n_used, true) <= 227 * 1024;
}

int window_cta_fwd(const WindowIndex& ix, const void* qkv, long long ld, void* O, long long ldo, float* lse, int Pn,
                   int H, int hd, float scale, cudaStream_t st) {
  WinParams P = {};
  int nthreads = 0;
  if (win_setup(P, ix, H, hd, nthreads)) return 1;
  P.qkv = (const bf16*)qkv; P.ld = ld; P.O = (bf16*)O; P.ldo = ldo; P.lse = lse; P.scale = scale;
  VALOR_REQUIRE(hd == 32 && H <= 65535, "window_cta_fwd: head dim 32 only");
  const size_t smem = win_smem_bytes(32, P.NP, P.n_used, false);
  auto kern = window_fwd_kernel<32>;
  static size_t attr = 0;
  if (smem > attr) { VALOR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = smem; }
  kern<<<dim3(Pn, H), nthreads, smem, st>>>(P);
  return check_launch("window_fwd_kernel");
}

int window_cta_bwd(const WindowIndex& ix, const void* qkv, long long ld, const void* O, const void* dO, long long ldo,
                   const float* lse, void* dqkv, long long lddqkv, float* dtable, int Pn, int H, int hd, float scale,
                   cudaStream_t st) {
  WinParams P = {};
  int nthreads = 0;
  if (win_setup(P, ix, H, hd, nthreads)) return 1;
  P.qkv = (const bf16*)qkv; P.ld = ld; P.O = (bf16*)O; P.ldo = ldo; P.lse = (float*)lse; P.scale = scale;
  P.dO = (const bf16*)dO; P.dqkv = (bf16*)dqkv; P.lddqkv = lddqkv; P.dtable = dtable;
  VALOR_REQUIRE(hd == 32 && H <= 65535, "window_cta_bwd: head dim 32 only");
  const size_t smem = win_smem_bytes(32, P.NP, P.n_used, true);
  auto kern = window_bwd_kernel<32>;
  static size_t attr = 0;
  if (smem > attr) { VALOR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = smem; }
  kern<<<dim3(Pn, H), nthreads, smem, st>>>(P);
  return check_launch("window_bwd_kernel");
}

}  // namespace valor
