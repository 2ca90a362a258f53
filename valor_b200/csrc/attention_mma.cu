// valor_b200 — key-blocked tensor-core flash attention (bf16, fp32 accumulate / softmax): BERT/AST
// self attention (32..129 tokens, hd 64), BERT cross attention (the 3 x 32 caption-pass queries of a
// sample x its 650 media keys with per-query key ranges, hd 64), and the fallback for window shapes
// window_attn.cu does not take (hd 64).  S and P never leave the SM: scores live in mma accumulators,
// the backward recomputes them from the saved log-sum-exp (the reference materialises the full
// softmax map for autograd: ~19.5 GB per step over the 24 Swin blocks).
//
// Window case: cyclic shift, window partition, relative-position bias and the -100 shift mask
// are evaluated from three small per-window tables built in shared memory at CTA start
// (row index, relative-position code, mask region) — nothing is gathered or rolled in HBM.
//
// MMA path: mma.sync.m16n8k16 (bf16) with ldmatrix operand fetch.  The softmax/bias/mask
// element work, not the MMA issue rate, bounds these kernels at hd 32.
#include "common.cuh"
#include "attention.cuh"
#include "mma_utils.cuh"

namespace valor {

static constexpr int BQ = 64, BKEY = 64;

// ---- per-problem tables in shared memory -------------------------------------------------
// One 32-bit info word per token keeps the per-element work short:
//   window : bits[0,16) relative-position code (query words carry code + center), bits[16,24) mask region
//   mha    : key words: bit0 = key is padding (-10000); query words: visible key range lo | hi << 16
//   both   : bit31 = key beyond the problem's key count (-inf, only the ragged last block)
struct Tables {
  int* qrow;            // [NqPad] global row of query i (-1 = padding)
  int* krow;            // [NkPad]
  uint32_t* qinfo;      // [NqPad]
  uint32_t* kinfo;      // [NkPad]
  float* tab2;          // window: this head's column of the bias table, pre-multiplied by log2(e)
  uint32_t qrow_s, krow_s, qinfo_s, kinfo_s, tab2_s;  // the same arrays as 32-bit shared addresses
  int causal, nq, nk, shifted, ranged;
};


// score in the log2 domain: s*scale*log2e + bias/mask
template <bool WINDOW>
__device__ __forceinline__ float score2(const Tables& t, float s, float sc2, uint32_t qi, uint32_t kj, int i, int j) {
  float v;
  if (WINDOW) {
    v = fmaf(s, sc2, lds_f32(t.tab2_s + 4u * (uint32_t)((int)(qi & 0xffffu) - (int)(kj & 0xffffu))));
    if (t.shifted && ((qi ^ kj) & 0x00ff0000u)) v += M100_2;
  } else {
    v = s * sc2;
    if ((kj & 1u) || (t.causal && j > i)) v += M10000_2;
    if (t.ranged && ((uint32_t)j < (qi & 0xffffu) || (uint32_t)j >= (qi >> 16))) v = -INFINITY;   // key outside this query's range
  }
  return v;
}

struct AttnParams {
  const bf16 *Q, *K, *V;
  long long ldq, ldk, ldv;
  bf16* O;  // fwd out / bwd: saved O
  long long ldo;
  float* lse;
  int H, hd, Nq, max_nk;
  float scale;
  // backward
  const bf16* dO;
  bf16* dQ; long long lddq;
  bf16* dK_lp; bf16* dV_lp; long long lddkv_lp;      // direct bf16 outputs (window / self)
  float* dK; float* dV; long long lddk, lddv;         // fp32 accumulate outputs (cross: shared K/V rows)
  float* dtable;
  // attention-probability dropout (bert.py:283,334; transformer.py:128): P_drop = keep * P / (1-p) feeds P.V; the row sums /
  // log-sum-exp stay those of the undropped softmax.  keep(i,j) = mix32(key(problem, head), i * max_nk + j) >= thr.
  float drop_p; const long long* rng; long long site;
  MhaIndex mha;
  WindowIndex win;
};
struct DropCtx {
  uint32_t key, thr, stride; float inv; bool on;
  __device__ __forceinline__ void init(const AttnParams& P, int p, int h) {
    on = P.drop_p > 0.f && P.rng != nullptr;
    key = 0; thr = 0; inv = 1.f; stride = (uint32_t)P.max_nk;
    if (on) { key = rng4(P.rng, P.site, (unsigned long long)((long long)p * P.H + h)).x; thr = drop_threshold(P.drop_p); inv = 1.0f / (1.0f - P.drop_p); }
  }
  __device__ __forceinline__ float apply(float v, int i, int j) const {   // v * keep / (1-p)
    return mix32(key, (uint32_t)i * stride + (uint32_t)j) < thr ? 0.f : v * inv;
  }
};

template <bool WINDOW>
__device__ __forceinline__ void build_tables(Tables& t, unsigned char* base, const AttnParams& P, int p, int h, int nq_pad, int nk_pad) {
  t.qrow = (int*)base; base += sizeof(int) * nq_pad;
  t.qinfo = (uint32_t*)base; base += sizeof(uint32_t) * nq_pad;
  if (WINDOW) {
    t.kinfo = (uint32_t*)base; base += sizeof(uint32_t) * nq_pad;
    t.krow = (int*)base; base += sizeof(int) * nq_pad;
    t.tab2 = (float*)base;
    const WindowIndex& ix = P.win;
    t.nq = t.nk = ix.N;
    t.causal = 0;
    t.ranged = 0;
    t.shifted = (ix.sd | ix.sh | ix.sw) != 0;
    const int cW = 2 * ix.WW - 1, cH = (2 * ix.WH - 1) * cW;
    const int center = (ix.WD - 1) * cH + (ix.WH - 1) * cW + (ix.WW - 1);
    // Queries keep the natural (d,h,w) order.  KEYS are enumerated (w,d,h) with h fastest: softmax does not
    // care about key order, and consecutive keys then differ by 13 in relative-position code while consecutive
    // queries differ by 1, so the 32 lanes of an MMA fragment hit 32 distinct bias-table slots (conflict-free
    // shared-memory gradient atomics and table reads).
    for (int i = threadIdx.x; i < nq_pad; i += blockDim.x) {
      if (i < ix.N) {
        int cd, ch, cw, b;
        {
          ix.coords(p, i, cd, ch, cw, b);
          t.qrow[i] = (int)ix.row(p, i);
          const int ld = i / (ix.wh * ix.ww), lh = (i / ix.ww) % ix.wh, lw = i % ix.ww;
          const uint32_t code = (uint32_t)(ld * cH + lh * cW + lw);
          const uint32_t reg = (uint32_t)(ix.region(cd, ix.D, ix.wd, ix.sd) * 9 + ix.region(ch, ix.H, ix.wh, ix.sh) * 3 +
                                          ix.region(cw, ix.W, ix.ww, ix.sw));
          t.qinfo[i] = (code + (uint32_t)center) | (reg << 16);
        }
        {
          const int lh = i % ix.wh, ld = (i / ix.wh) % ix.wd, lw = i / (ix.wh * ix.wd);
          const int nat = (ld * ix.wh + lh) * ix.ww + lw;
          ix.coords(p, nat, cd, ch, cw, b);
          t.krow[i] = (int)ix.row(p, nat);
          const uint32_t code = (uint32_t)(ld * cH + lh * cW + lw);
          const uint32_t reg = (uint32_t)(ix.region(cd, ix.D, ix.wd, ix.sd) * 9 + ix.region(ch, ix.H, ix.wh, ix.sh) * 3 +
                                          ix.region(cw, ix.W, ix.ww, ix.sw));
          t.kinfo[i] = code | (reg << 16);
        }
      } else {
        t.qrow[i] = -1; t.krow[i] = -1; t.kinfo[i] = 0x80000000u; t.qinfo[i] = (uint32_t)center;
      }
    }
    const int n_rel = (2 * ix.WD - 1) * cH;
    for (int r = threadIdx.x; r < n_rel; r += blockDim.x) t.tab2[r] = ix.table[(size_t)r * ix.heads + h] * LOG2E;
  } else {
    t.krow = (int*)base; base += sizeof(int) * nk_pad;
    t.kinfo = (uint32_t*)base;
    t.tab2 = nullptr;
    t.shifted = 0;
    const MhaIndex& ix = P.mha;
    t.nq = ix.Nq;
    t.nk = ix.nk(p);
    t.causal = ix.causal ? (int)ix.causal[p] : 0;
    t.ranged = ix.q_key_range != nullptr;
    for (int i = threadIdx.x; i < nq_pad; i += blockDim.x) {   // qinfo: visible key range lo | hi << 16
      t.qrow[i] = i < t.nq ? (int)ix.qrow(p, i) : -1;
      uint32_t w = 0xffff0000u;
      if (ix.q_key_range != nullptr && i < t.nq) w = (uint32_t)ix.q_key_range[2 * i] | ((uint32_t)ix.q_key_range[2 * i + 1] << 16);
      t.qinfo[i] = w;
    }
    for (int j = threadIdx.x; j < nk_pad; j += blockDim.x) {
      t.krow[j] = j < t.nk ? (int)ix.krow(p, j) : -1;
      uint32_t w = 0;
      if (j >= t.nk) w = 0x80000000u;
      else if (ix.key_valid != nullptr && ix.key_valid[(size_t)p * ix.max_nk + j] == 0) w = 1u;
      t.kinfo[j] = w;
    }
  }
  t.qrow_s = s_u32(t.qrow); t.krow_s = s_u32(t.krow); t.qinfo_s = s_u32(t.qinfo); t.kinfo_s = s_u32(t.kinfo);
  t.tab2_s = t.tab2 ? s_u32(t.tab2) : 0u;
}

// gather 64 rows x HD bf16 (16-byte chunks) into a padded smem tile; rows < 0 -> zeros
template <int HD>
__device__ __forceinline__ void load_tile(unsigned char* dst, const bf16* src, long long ld, int col0, uint32_t rows_s, int r0) {
  constexpr int PITCH = HD * 2 + 16, CH = HD / 8;
  for (int c = threadIdx.x; c < 64 * CH; c += blockDim.x) {
    const int r = c / CH, ch = c % CH;
    const int gr = lds_s32(rows_s + 4u * (uint32_t)(r0 + r));
    uint4 v = make_uint4(0, 0, 0, 0);
    if (gr >= 0) v = *(const uint4*)(src + (size_t)gr * ld + col0 + ch * 8);
    *(uint4*)(dst + r * PITCH + ch * 16) = v;
  }
}

// same gather, issued as cp.async (16-byte, zero-fill for padding rows): the copy of block i+1 overlaps
// the MMA / softmax work on block i
template <int HD>
__device__ __forceinline__ void load_tile_async(unsigned char* dst, const bf16* src, long long ld, int col0, uint32_t rows_s, int r0) {
  constexpr int PITCH = HD * 2 + 16, CH = HD / 8;
  for (int c = threadIdx.x; c < 64 * CH; c += blockDim.x) {
    const int r = c / CH, ch = c % CH;
    const int gr = lds_s32(rows_s + 4u * (uint32_t)(r0 + r));
    const bf16* g = src + (size_t)(gr < 0 ? 0 : gr) * ld + col0 + ch * 8;
    const uint32_t d = s_u32(dst + r * PITCH + ch * 16);
    const int nbytes = gr < 0 ? 0 : 16;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(g), "r"(nbytes) : "memory");
  }
}

// Key blocks a 64-query block has to visit: the union of its queries' visible key ranges (every warp computes the same
// pair from the query words, no barrier).  Without per-query ranges this is [0, ceil(nk/64)).
__device__ __forceinline__ void key_block_span(const Tables& t, int q0, int& kb0, int& kb1) {
  const int nkb = (t.nk + BKEY - 1) / BKEY;
  kb0 = 0; kb1 = nkb;
  if (!t.ranged) return;
  const int lane = threadIdx.x & 31;
  uint32_t lo = 0xffffu, hi = 0u;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int i = q0 + lane + 32 * r;
    if (i < t.nq) { const uint32_t w = t.qinfo[i]; lo = min(lo, w & 0xffffu); hi = max(hi, w >> 16); }
  }
#pragma unroll
  for (int sft = 16; sft > 0; sft >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, sft));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, sft));
  }
  kb0 = min((int)lo / BKEY, nkb);
  kb1 = max(kb0, min(((int)hi + BKEY - 1) / BKEY, nkb));
}

// ==========================================================================================
// forward: grid (ceil(Nq/64), P, H), 128 threads; warp w owns query rows [w*16, w*16+16)
// ==========================================================================================
template <int HD, bool WINDOW>
__global__ void __launch_bounds__(128, 4)
attn_mma_fwd_kernel(AttnParams P, int nq_pad, int nk_pad) {
  constexpr int PITCH = HD * 2 + 16, TILE = 64 * PITCH;
  extern __shared__ __align__(16) unsigned char smem[];
  unsigned char* Qs = smem;
  unsigned char* KVs = Qs + TILE;  // [2 buffers][K | V]
  const int p = blockIdx.y, h = blockIdx.z, qb = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  Tables t;
  build_tables<WINDOW>(t, KVs + 4 * TILE, P, p, h, nq_pad, nk_pad);
  __syncthreads();
  const int col0 = h * HD;
  int kb0, kb1;
  key_block_span(t, qb * BQ, kb0, kb1);
  load_tile_async<HD>(Qs, P.Q, P.ldq, col0, t.qrow_s, qb * BQ);
  load_tile_async<HD>(KVs, P.K, P.ldk, col0, t.krow_s, kb0 * BKEY);
  load_tile_async<HD>(KVs + TILE, P.V, P.ldv, col0, t.krow_s, kb0 * BKEY);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  uint32_t qf[HD / 16][4];
  {
    const int m = lane >> 3, r = lane & 7;
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks)
      ldsm_x4(qf[ks], s_u32(Qs + (warp * 16 + (m & 1) * 8 + r) * PITCH + (ks * 16 + (m >> 1) * 8) * 2));
  }
  float o[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};
  const int i0 = qb * BQ + warp * 16 + g;  // rows i0 and i0+8
  const float sc2 = P.scale * LOG2E;
  const uint32_t qinf[2] = {t.qinfo[i0], t.qinfo[i0 + 8]};
  DropCtx drop;
  drop.init(P, p, h);
  for (int kb = kb0; kb < kb1; ++kb) {
    unsigned char* Ks = KVs + ((kb - kb0) & 1) * 2 * TILE;
    unsigned char* Vs = Ks + TILE;
    if (kb > kb0) {  // block kb was prefetched during block kb-1
      cp_async_wait<0>();
      __syncthreads();
    }
    if (kb + 1 < kb1) {  // prefetch the next K/V block into the other buffer (its readers finished at the barrier above)
      unsigned char* Kn = KVs + ((kb + 1 - kb0) & 1) * 2 * TILE;
      load_tile_async<HD>(Kn, P.K, P.ldk, col0, t.krow_s, (kb + 1) * BKEY);
      load_tile_async<HD>(Kn + TILE, P.V, P.ldv, col0, t.krow_s, (kb + 1) * BKEY);
      cp_async_commit();
    }
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
    {
      const int m = lane >> 3, r = lane & 7;
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks)
#pragma unroll
        for (int nt = 0; nt < 8; nt += 2) {
          uint32_t b[4];
          ldsm_x4(b, s_u32(Ks + ((nt + (m >> 1)) * 8 + r) * PITCH + (ks * 16 + (m & 1) * 8) * 2));
          mma16816(s[nt], qf[ks], b);
          mma16816(s[nt + 1], qf[ks], b + 2);
        }
    }
    // scale + bias/mask (log2 domain), running max
    float mnew[2] = {mrow[0], mrow[1]};
    const bool ragged = (kb + 1) * BKEY > t.nk;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int jb = kb * BKEY + nt * 8 + t4 * 2;
      const uint2 kj = lds_v2u32(t.kinfo_s + 4u * (uint32_t)jb);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const uint32_t kw = (e & 1) ? kj.y : kj.x;
        float v = score2<WINDOW>(t, s[nt][e], sc2, qinf[e >> 1], kw, i0 + (e >> 1) * 8, jb + (e & 1));
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
    float corr[2], msafe[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      msafe[r] = mnew[r] == -INFINITY ? 0.f : mnew[r];
      corr[r] = exp2f(mrow[r] - msafe[r]);  // mrow=-inf -> 0
      mrow[r] = mnew[r];
      lrow[r] *= corr[r];
    }
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) { o[i][0] *= corr[0]; o[i][1] *= corr[0]; o[i][2] *= corr[1]; o[i][3] *= corr[1]; }
    uint32_t pf[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float pv = fast_exp2(s[nt][e] - msafe[e >> 1]);
        lrow[e >> 1] += pv;
        s[nt][e] = (!WINDOW && drop.on) ? drop.apply(pv, i0 + (e >> 1) * 8, kb * BKEY + nt * 8 + t4 * 2 + (e & 1)) : pv;
      }
      pf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(s[nt][0], s[nt][1]);
      pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(s[nt][2], s[nt][3]);
    }
    {
      const int m = lane >> 3, r = lane & 7;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int dt = 0; dt < HD / 8; dt += 2) {
          uint32_t b[4];
          ldsm_x4_t(b, s_u32(Vs + (kk * 16 + (m & 1) * 8 + r) * PITCH + ((dt + (m >> 1)) * 8) * 2));
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
    if (i < t.nq) {
      const float inv = 1.f / lrow[r];
      bf16* dst = P.O + (size_t)t.qrow[i] * P.ldo + col0;
#pragma unroll
      for (int dt = 0; dt < HD / 8; ++dt)
        *(uint32_t*)(dst + dt * 8 + t4 * 2) = pack_bf16(o[dt][r * 2] * inv, o[dt][r * 2 + 1] * inv);
      if (t4 == 0) P.lse[((size_t)p * P.H + h) * P.Nq + i] = mrow[r] * LN2 + __logf(lrow[r]);
    }
  }
}

// ==========================================================================================
// backward, split in two forward-shaped kernels (small shared-memory footprint -> many CTAs per
// SM hide the gather latency; no cross-warp accumulation panels, no P/dS round trip):
//   dq  kernel: CTA = 64 queries; loops key blocks;  S, P, dP, dS in registers, dQ += dS.K,
//               delta = rowsum(dO*O) computed here and published for the dkv kernel,
//               relative-position-bias gradient via shared-memory atomics (window case)
//   dkv kernel: CTA = 64 keys;    loops query blocks; S^T = K.Q^T etc. so that P^T / dS^T come out
//               of the MMA already in A-operand layout: dV += P^T.dO, dK += dS^T.Q
// ==========================================================================================
template <int HD, bool WINDOW>
__global__ void __launch_bounds__(128, 4)
attn_mma_bwd_dq_kernel(AttnParams P, float* __restrict__ delta, int nq_pad, int nk_pad) {
  constexpr int PITCH = HD * 2 + 16, TILE = 64 * PITCH;
  extern __shared__ __align__(16) unsigned char smem[];
  unsigned char* Qs = smem;
  unsigned char* dOs = Qs + TILE;
  unsigned char* KVs = dOs + TILE;             // [2 buffers][K | V]
  float* lse_s = (float*)(KVs + 4 * TILE);     // [64]
  float* del_s = lse_s + 64;                 // [64]
  // window: bias-table gradient in shared memory.  Only 32-bit INTEGER shared atomics are native (fp32 and
  // 64-bit ones compile to compare-and-swap spin loops, which dominated this kernel), so each slot is a
  // pair of int32 fixed-point words: a coarse word in units of 2^-16 (|per-CTA sum| < 32768) and the rounding
  // remainder in units of 2^-40 (64 contributions of < 2^23 each never overflow).  Two ATOMS.ADD without
  // return per element; integer adds commute, so the result is order-independent.
  uint32_t* dtab_s = (uint32_t*)(del_s + 64);  // [n_rel][2] : lo, hi words of a 64-bit fixed-point sum
  const int p = blockIdx.y, h = blockIdx.z, qb = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  int n_rel = 0;
  if (WINDOW) n_rel = (2 * P.win.WD - 1) * (2 * P.win.WH - 1) * (2 * P.win.WW - 1);
  Tables t;
  build_tables<WINDOW>(t, (unsigned char*)(((uintptr_t)(dtab_s + 2 * n_rel) + 15) & ~(uintptr_t)15), P, p, h, nq_pad, nk_pad);
  for (int i = threadIdx.x; i < 2 * n_rel; i += blockDim.x) dtab_s[i] = 0u;
  __syncthreads();
  const int col0 = h * HD;
  load_tile_async<HD>(Qs, P.Q, P.ldq, col0, t.qrow_s, qb * BQ);
  load_tile_async<HD>(dOs, P.dO, P.ldo, col0, t.qrow_s, qb * BQ);
  load_tile_async<HD>(KVs + 2 * TILE, P.O, P.ldo, col0, t.qrow_s, qb * BQ);  // O tile (buffer 1), only for delta
  int kb0, kb1;
  key_block_span(t, qb * BQ, kb0, kb1);
  load_tile_async<HD>(KVs, P.K, P.ldk, col0, t.krow_s, kb0 * BKEY);          // first K/V block (buffer 0)
  load_tile_async<HD>(KVs + TILE, P.V, P.ldv, col0, t.krow_s, kb0 * BKEY);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  {  // delta_i = sum_d dO[i,d] * O[i,d] : two threads per row
    const int r = threadIdx.x >> 1, hf = threadIdx.x & 1;
    float d = 0.f;
    const __nv_bfloat162* a = (const __nv_bfloat162*)(dOs + r * PITCH + hf * HD);
    const __nv_bfloat162* b = (const __nv_bfloat162*)(KVs + 2 * TILE + r * PITCH + hf * HD);
#pragma unroll
    for (int c = 0; c < HD / 4; ++c) {
      const float2 fa = __bfloat1622float2(a[c]), fb = __bfloat1622float2(b[c]);
      d += fa.x * fb.x + fa.y * fb.y;
    }
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    const int i = qb * BQ + r;
    if (hf == 0) {
      del_s[r] = d;
      lse_s[r] = i < t.nq ? P.lse[((size_t)p * P.H + h) * P.Nq + i] : 0.f;
      if (i < t.nq) delta[((size_t)p * P.H + h) * P.Nq + i] = d;
    }
  }
  const int m8 = lane >> 3, r8 = lane & 7;
  uint32_t qf[HD / 16][4], dof[HD / 16][4];
#pragma unroll
  for (int ks = 0; ks < HD / 16; ++ks) {
    ldsm_x4(qf[ks], s_u32(Qs + (warp * 16 + (m8 & 1) * 8 + r8) * PITCH + (ks * 16 + (m8 >> 1) * 8) * 2));
    ldsm_x4(dof[ks], s_u32(dOs + (warp * 16 + (m8 & 1) * 8 + r8) * PITCH + (ks * 16 + (m8 >> 1) * 8) * 2));
  }
  float dq[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
  const float sc = P.scale, sc2 = P.scale * LOG2E;
  const int il = warp * 16 + g;
  const int i0 = qb * BQ + il;
  __syncthreads();  // lse_s / del_s visible
  // padding query rows (i >= nq) get lse = +inf -> p = 0 -> no gradient, no bias-table contribution
  const uint32_t qinf[2] = {t.qinfo[i0], t.qinfo[i0 + 8]};
  const float lse2[2] = {i0 < t.nq ? lse_s[il] * LOG2E : INFINITY, i0 + 8 < t.nq ? lse_s[il + 8] * LOG2E : INFINITY};
  const float del[2] = {del_s[il], del_s[il + 8]};
  DropCtx drop;
  drop.init(P, p, h);
  for (int kb = kb0; kb < kb1; ++kb) {
    unsigned char* Ks = KVs + ((kb - kb0) & 1) * 2 * TILE;
    unsigned char* Vs = Ks + TILE;
    if (kb > kb0) {
      cp_async_wait<0>();
      __syncthreads();
    }
    if (kb + 1 < kb1) {  // (the barrier before this loop / above guarantees the other buffer is no longer read)
      unsigned char* Kn = KVs + ((kb + 1 - kb0) & 1) * 2 * TILE;
      load_tile_async<HD>(Kn, P.K, P.ldk, col0, t.krow_s, (kb + 1) * BKEY);
      load_tile_async<HD>(Kn + TILE, P.V, P.ldv, col0, t.krow_s, (kb + 1) * BKEY);
      cp_async_commit();
    }
    const bool ragged = (kb + 1) * BKEY > t.nk;
#pragma unroll 1
    for (int hh = 0; hh < 2; ++hh) {  // two 32-key halves: halves the live accumulator registers
      float s[4][4], dp[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) s[nt][e] = dp[nt][e] = 0.f;
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks)
#pragma unroll
        for (int nt = 0; nt < 4; nt += 2) {
          uint32_t b[4];
          ldsm_x4(b, s_u32(Ks + ((hh * 4 + nt + (m8 >> 1)) * 8 + r8) * PITCH + (ks * 16 + (m8 & 1) * 8) * 2));
          mma16816(s[nt], qf[ks], b);
          mma16816(s[nt + 1], qf[ks], b + 2);
          ldsm_x4(b, s_u32(Vs + ((hh * 4 + nt + (m8 >> 1)) * 8 + r8) * PITCH + (ks * 16 + (m8 & 1) * 8) * 2));
          mma16816(dp[nt], dof[ks], b);
          mma16816(dp[nt + 1], dof[ks], b + 2);
        }
      uint32_t dsf[2][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int jb = kb * BKEY + (hh * 4 + nt) * 8 + t4 * 2;
        const uint2 kj = lds_v2u32(t.kinfo_s + 4u * (uint32_t)jb);
        float ds[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int r = e >> 1;
          const uint32_t kw = (e & 1) ? kj.y : kj.x;
          const float v = score2<WINDOW>(t, s[nt][e], sc2, qinf[r], kw, i0 + r * 8, jb + (e & 1));
          float pr = fast_exp2(v - lse2[r]);
          if (ragged && (kw >> 31)) pr = 0.f;
          const float dpe = (!WINDOW && drop.on) ? drop.apply(dp[nt][e], i0 + r * 8, jb + (e & 1)) : dp[nt][e];
          const float d = pr * (dpe - del[r]);
          if (WINDOW) {
            // two native int32 adds without return: coarse word at 2^-16, remainder word at 2^-40
            const int hi = __float2int_rn(d * 65536.0f);
            const int lo = __float2int_rn(fmaf((float)hi, -1.0f / 65536.0f, d) * 1099511627776.0f);
            int* slot = (int*)dtab_s + 2 * ((int)(qinf[r] & 0xffffu) - (int)(kw & 0xffffu));
            atomicAdd(slot, lo);
            atomicAdd(slot + 1, hi);
          }
          ds[e] = d * sc;
        }
        dsf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(ds[0], ds[1]);
        dsf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(ds[2], ds[3]);
      }
#pragma unroll
      for (int kk = 0; kk < 2; ++kk)
#pragma unroll
        for (int dt = 0; dt < HD / 8; dt += 2) {
          uint32_t b[4];
          ldsm_x4_t(b, s_u32(Ks + ((hh * 2 + kk) * 16 + (m8 & 1) * 8 + r8) * PITCH + ((dt + (m8 >> 1)) * 8) * 2));
          mma16816(dq[dt], dsf[kk], b);
          mma16816(dq[dt + 1], dsf[kk], b + 2);
        }
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int i = i0 + r * 8;
    if (i < t.nq) {
      bf16* dst = P.dQ + (size_t)t.qrow[i] * P.lddq + col0;
#pragma unroll
      for (int dt = 0; dt < HD / 8; ++dt)
        *(uint32_t*)(dst + dt * 8 + t4 * 2) = pack_bf16(dq[dt][r * 2], dq[dt][r * 2 + 1]);
    }
  }
  if (WINDOW && P.dtable != nullptr) {
    __syncthreads();
    for (int r = threadIdx.x; r < n_rel; r += blockDim.x) {
      const int lo = (int)dtab_s[2 * r], hi = (int)dtab_s[2 * r + 1];
      if (lo != 0 || hi != 0)
        atomicAdd(&P.dtable[(size_t)r * P.win.heads + h], (float)hi * (1.0f / 65536.0f) + (float)lo * (1.0f / 1099511627776.0f));
    }
  }
}

template <int HD, bool WINDOW>
__global__ void __launch_bounds__(128, 4)
attn_mma_bwd_dkv_kernel(AttnParams P, const float* __restrict__ delta, int nq_pad, int nk_pad) {
  constexpr int PITCH = HD * 2 + 16, TILE = 64 * PITCH;
  extern __shared__ __align__(16) unsigned char smem[];
  unsigned char* Ks = smem;
  unsigned char* Vs = Ks + TILE;
  unsigned char* QDs = Vs + TILE;               // [2 buffers][Q | dO]
  float* lse_s = (float*)(QDs + 4 * TILE);      // [nq_pad]
  float* del_s = lse_s + nq_pad;              // [nq_pad]
  const int p = blockIdx.y, h = blockIdx.z, kb = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  Tables t;
  build_tables<WINDOW>(t, (unsigned char*)(del_s + nq_pad), P, p, h, nq_pad, nk_pad);
  __syncthreads();
  if (kb * BKEY >= t.nk) return;  // cross-attention: shorter key ranges than max_nk
  const int col0 = h * HD;
  for (int i = threadIdx.x; i < nq_pad; i += blockDim.x) {
    const bool ok = i < t.nq;
    lse_s[i] = ok ? P.lse[((size_t)p * P.H + h) * P.Nq + i] * LOG2E : INFINITY;  // log2 domain; +inf kills padding
    del_s[i] = ok ? delta[((size_t)p * P.H + h) * P.Nq + i] : 0.f;
  }
  // query blocks whose visible key range touches this key block (all of them without per-query ranges)
  const int nqb = (t.nq + BQ - 1) / BQ;
  uint32_t qmask = 0;
  for (int q = 0; q < nqb && q < 32; ++q) {
    int a, b;
    key_block_span(t, q * BQ, a, b);
    if (kb >= a && kb < b) qmask |= 1u << q;
  }
  if (nqb > 32) qmask = 0xffffffffu;   // (never on this path: Nq <= 2048)
  load_tile_async<HD>(Ks, P.K, P.ldk, col0, t.krow_s, kb * BKEY);
  load_tile_async<HD>(Vs, P.V, P.ldv, col0, t.krow_s, kb * BKEY);
  const int qb_first = qmask ? __ffs(qmask) - 1 : 0;
  load_tile_async<HD>(QDs, P.Q, P.ldq, col0, t.qrow_s, qb_first * BQ);
  load_tile_async<HD>(QDs + TILE, P.dO, P.ldo, col0, t.qrow_s, qb_first * BQ);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  const int m8 = lane >> 3, r8 = lane & 7;
  uint32_t kf[HD / 16][4], vf[HD / 16][4];
#pragma unroll
  for (int ks = 0; ks < HD / 16; ++ks) {
    ldsm_x4(kf[ks], s_u32(Ks + (warp * 16 + (m8 & 1) * 8 + r8) * PITCH + (ks * 16 + (m8 >> 1) * 8) * 2));
    ldsm_x4(vf[ks], s_u32(Vs + (warp * 16 + (m8 & 1) * 8 + r8) * PITCH + (ks * 16 + (m8 >> 1) * 8) * 2));
  }
  float dk[HD / 8][4], dv[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) dk[i][e] = dv[i][e] = 0.f;
  const float sc = P.scale, sc2 = P.scale * LOG2E;
  const int j0 = kb * BKEY + warp * 16 + g;  // keys j0 and j0+8 (padding keys: K/V rows are zero, results dropped)
  const uint32_t kinf[2] = {t.kinfo[j0], t.kinfo[j0 + 8]};
  DropCtx drop;
  drop.init(P, p, h);
  uint32_t todo = qmask;
  for (int it = 0; todo != 0; ++it) {
    const int qb = __ffs(todo) - 1;
    todo &= todo - 1;
    unsigned char* Qs = QDs + (it & 1) * 2 * TILE;
    unsigned char* dOs = Qs + TILE;
    if (it > 0) {
      cp_async_wait<0>();
      __syncthreads();
    }
    if (todo != 0) {   // prefetch the next relevant query block into the other buffer
      const int qn = __ffs(todo) - 1;
      unsigned char* Qn = QDs + ((it + 1) & 1) * 2 * TILE;
      load_tile_async<HD>(Qn, P.Q, P.ldq, col0, t.qrow_s, qn * BQ);
      load_tile_async<HD>(Qn + TILE, P.dO, P.ldo, col0, t.qrow_s, qn * BQ);
      cp_async_commit();
    }
#pragma unroll 1
    for (int hh = 0; hh < 2; ++hh) {  // two 32-query halves
      float s[4][4], dp[4][4];  // rows = keys (g, g+8), cols = queries
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) s[nt][e] = dp[nt][e] = 0.f;
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks)
#pragma unroll
        for (int nt = 0; nt < 4; nt += 2) {
          uint32_t b[4];
          ldsm_x4(b, s_u32(Qs + ((hh * 4 + nt + (m8 >> 1)) * 8 + r8) * PITCH + (ks * 16 + (m8 & 1) * 8) * 2));
          mma16816(s[nt], kf[ks], b);
          mma16816(s[nt + 1], kf[ks], b + 2);
          ldsm_x4(b, s_u32(dOs + ((hh * 4 + nt + (m8 >> 1)) * 8 + r8) * PITCH + (ks * 16 + (m8 & 1) * 8) * 2));
          mma16816(dp[nt], vf[ks], b);
          mma16816(dp[nt + 1], vf[ks], b + 2);
        }
      uint32_t pf[2][4], dsf[2][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int ib = qb * BQ + (hh * 4 + nt) * 8 + t4 * 2;   // two consecutive queries (columns of S^T)
        const uint2 qi2 = lds_v2u32(t.qinfo_s + 4u * (uint32_t)ib);
        const float2 l2 = *(const float2*)(lse_s + ib);
        const float2 d2 = *(const float2*)(del_s + ib);
        float pv[4], ds[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int r = e >> 1;
          const uint32_t qw = (e & 1) ? qi2.y : qi2.x;
          const float v = score2<WINDOW>(t, s[nt][e], sc2, qw, kinf[r], ib + (e & 1), j0 + r * 8);
          const float pr = fast_exp2(v - ((e & 1) ? l2.y : l2.x));   // padding queries carry lse = +inf -> 0
          const bool dr = !WINDOW && drop.on;
          pv[e] = dr ? drop.apply(pr, ib + (e & 1), j0 + r * 8) : pr;
          const float dpe = dr ? drop.apply(dp[nt][e], ib + (e & 1), j0 + r * 8) : dp[nt][e];
          ds[e] = pr * (dpe - ((e & 1) ? d2.y : d2.x)) * sc;
        }
        pf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(pv[0], pv[1]);
        pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(pv[2], pv[3]);
        dsf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(ds[0], ds[1]);
        dsf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(ds[2], ds[3]);
      }
#pragma unroll
      for (int kk = 0; kk < 2; ++kk)  // contraction over the 32 queries of this half
#pragma unroll
        for (int dt = 0; dt < HD / 8; dt += 2) {
          uint32_t b[4];
          ldsm_x4_t(b, s_u32(dOs + ((hh * 2 + kk) * 16 + (m8 & 1) * 8 + r8) * PITCH + ((dt + (m8 >> 1)) * 8) * 2));
          mma16816(dv[dt], pf[kk], b);
          mma16816(dv[dt + 1], pf[kk], b + 2);
          ldsm_x4_t(b, s_u32(Qs + ((hh * 2 + kk) * 16 + (m8 & 1) * 8 + r8) * PITCH + ((dt + (m8 >> 1)) * 8) * 2));
          mma16816(dk[dt], dsf[kk], b);
          mma16816(dk[dt + 1], dsf[kk], b + 2);
        }
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int j = j0 + r * 8;
    if (j < t.nk) {
      const size_t gr = (size_t)t.krow[j];
      if (P.dK_lp != nullptr) {
#pragma unroll
        for (int dt = 0; dt < HD / 8; ++dt) {
          *(uint32_t*)(P.dK_lp + gr * P.lddkv_lp + col0 + dt * 8 + t4 * 2) = pack_bf16(dk[dt][r * 2], dk[dt][r * 2 + 1]);
          *(uint32_t*)(P.dV_lp + gr * P.lddkv_lp + col0 + dt * 8 + t4 * 2) = pack_bf16(dv[dt][r * 2], dv[dt][r * 2 + 1]);
        }
      } else {
#pragma unroll
        for (int dt = 0; dt < HD / 8; ++dt) {
          atomicAdd(P.dK + gr * P.lddk + col0 + dt * 8 + t4 * 2, dk[dt][r * 2]);
          atomicAdd(P.dK + gr * P.lddk + col0 + dt * 8 + t4 * 2 + 1, dk[dt][r * 2 + 1]);
          atomicAdd(P.dV + gr * P.lddv + col0 + dt * 8 + t4 * 2, dv[dt][r * 2]);
          atomicAdd(P.dV + gr * P.lddv + col0 + dt * 8 + t4 * 2 + 1, dv[dt][r * 2 + 1]);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static inline int pad64(int n) { return (n + 63) / 64 * 64; }

template <int HD, bool WINDOW>
static size_t table_bytes(const AttnParams& P, int nq_pad, int nk_pad) {
  if (WINDOW) {
    const int n_rel = (2 * P.win.WD - 1) * (2 * P.win.WH - 1) * (2 * P.win.WW - 1);
    return sizeof(int) * 4 * nq_pad + sizeof(float) * n_rel;
  }
  return sizeof(int) * 2 * (nq_pad + nk_pad);
}

template <int HD, bool WINDOW>
static int launch_fwd(const AttnParams& P, int Pn, int nq, int max_nk, cudaStream_t st) {
  constexpr int PITCH = HD * 2 + 16;
  const int nq_pad = pad64(nq), nk_pad = pad64(max_nk);
  const size_t smem = 5 * 64 * PITCH + table_bytes<HD, WINDOW>(P, nq_pad, nk_pad) + 16;
  auto kern = attn_mma_fwd_kernel<HD, WINDOW>;
  if (smem > 48 * 1024) VALOR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(nq_pad / 64, Pn, P.H);
  kern<<<grid, 128, smem, st>>>(P, nq_pad, nk_pad);
  return check_launch("attn_mma_fwd_kernel");
}

template <int HD, bool WINDOW>
static int launch_bwd(const AttnParams& P, float* delta, int Pn, int nq, int max_nk, cudaStream_t st) {
  constexpr int PITCH = HD * 2 + 16;
  const int nq_pad = pad64(nq), nk_pad = pad64(max_nk);
  VALOR_REQUIRE(nq <= 2048, "attention backward: at most 2048 queries per problem (query-block bitmask)");
  size_t n_rel = 0;
  if (WINDOW) n_rel = (size_t)(2 * P.win.WD - 1) * (2 * P.win.WH - 1) * (2 * P.win.WW - 1);
  const size_t tb = table_bytes<HD, WINDOW>(P, nq_pad, nk_pad);
  {
    const size_t smem = 6 * 64 * PITCH + sizeof(float) * (128 + 2 * n_rel) + tb + 32;
    auto kern = attn_mma_bwd_dq_kernel<HD, WINDOW>;
    if (smem > 48 * 1024) VALOR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(nq_pad / 64, Pn, P.H);
    kern<<<grid, 128, smem, st>>>(P, delta, nq_pad, nk_pad);
    if (check_launch("attn_mma_bwd_dq_kernel")) return 1;
  }
  {
    const size_t smem = 6 * 64 * PITCH + sizeof(float) * 2 * nq_pad + tb + 16;
    auto kern = attn_mma_bwd_dkv_kernel<HD, WINDOW>;
    if (smem > 48 * 1024) VALOR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(nk_pad / 64, Pn, P.H);
    kern<<<grid, 128, smem, st>>>(P, delta, nq_pad, nk_pad);
    return check_launch("attn_mma_bwd_dkv_kernel");
  }
}

bool attn_mma_eligible(int dtype, int hd, long long ldq, long long ldk, long long ldv, long long ldo, const void* q,
                       const void* k, const void* v, const void* o) {
  if (dtype != VALOR_DT_BF16 || (hd != 32 && hd != 64)) return false;
  if ((ldq | ldk | ldv | ldo) % 8) return false;
  if (((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)o) & 15) return false;
  return true;
}

int mha_mma_fwd(const MhaIndex& ix, const void* Q, const void* K, const void* V, long long ldq, long long ldk,
                long long ldv, void* O, long long ldo, float* lse, int Pn, int H, int hd, int Nq, float scale,
                float drop_p, const long long* rng, long long site, cudaStream_t st) {
  AttnParams P = {};
  P.drop_p = drop_p; P.rng = rng; P.site = site;
  P.Q = (const bf16*)Q; P.K = (const bf16*)K; P.V = (const bf16*)V; P.ldq = ldq; P.ldk = ldk; P.ldv = ldv;
  P.O = (bf16*)O; P.ldo = ldo; P.lse = lse; P.H = H; P.hd = hd; P.Nq = Nq; P.max_nk = ix.max_nk; P.scale = scale;
  P.mha = ix;
  VALOR_REQUIRE(Pn <= 65535, "mha: too many problems (%d)", Pn);
  return hd == 32 ? launch_fwd<32, false>(P, Pn, Nq, ix.max_nk, st) : launch_fwd<64, false>(P, Pn, Nq, ix.max_nk, st);
}

int mha_mma_bwd(const MhaIndex& ix, const void* Q, const void* K, const void* V, const void* O, const void* dO,
                long long ldq, long long ldk, long long ldv, long long ldo, const float* lse, float* delta, void* dQ,
                long long lddq, float* dK, float* dV, long long lddk, long long lddv, void* dK_lp, void* dV_lp,
                long long lddkv_lp, int Pn, int H, int hd, int Nq, float scale, float drop_p, const long long* rng,
                long long site, cudaStream_t st) {
  AttnParams P = {};
  P.drop_p = drop_p; P.rng = rng; P.site = site;
  P.Q = (const bf16*)Q; P.K = (const bf16*)K; P.V = (const bf16*)V; P.ldq = ldq; P.ldk = ldk; P.ldv = ldv;
  P.O = (bf16*)O; P.ldo = ldo; P.lse = (float*)lse; P.H = H; P.hd = hd; P.Nq = Nq; P.max_nk = ix.max_nk; P.scale = scale;
  P.dO = (const bf16*)dO; P.dQ = (bf16*)dQ; P.lddq = lddq; P.dK = dK; P.dV = dV; P.lddk = lddk; P.lddv = lddv;
  P.dK_lp = (bf16*)dK_lp; P.dV_lp = (bf16*)dV_lp; P.lddkv_lp = lddkv_lp;
  VALOR_REQUIRE((dK != nullptr) != (dK_lp != nullptr), "mha_bwd: give exactly one of (dK,dV) / (dK_lp,dV_lp)");
  P.mha = ix;
  VALOR_REQUIRE(Pn <= 65535 && H <= 65535, "mha: too many problems/heads");
  return hd == 32 ? launch_bwd<32, false>(P, delta, Pn, Nq, ix.max_nk, st)
                  : launch_bwd<64, false>(P, delta, Pn, Nq, ix.max_nk, st);
}

int window_mma_fwd(const WindowIndex& ix, const void* qkv, long long ld, void* O, long long ldo, float* lse, int Pn,
                   int H, int hd, float scale, cudaStream_t st) {
  const int C = H * hd;
  AttnParams P = {};
  P.Q = (const bf16*)qkv; P.K = P.Q + C; P.V = P.Q + 2 * C; P.ldq = P.ldk = P.ldv = ld;
  P.O = (bf16*)O; P.ldo = ldo; P.lse = lse; P.H = H; P.hd = hd; P.Nq = ix.N; P.max_nk = ix.N; P.scale = scale;
  P.win = ix;
  VALOR_REQUIRE(Pn <= 65535, "window_attn: too many windows (%d)", Pn);
  return hd == 32 ? launch_fwd<32, true>(P, Pn, ix.N, ix.N, st) : launch_fwd<64, true>(P, Pn, ix.N, ix.N, st);
}

int window_mma_bwd(const WindowIndex& ix, const void* qkv, long long ld, const void* O, const void* dO, long long ldo,
                   const float* lse, float* delta, void* dqkv, long long lddq, float* dtable, int Pn, int H, int hd,
                   float scale, cudaStream_t st) {
  const int C = H * hd;
  AttnParams P = {};
  P.Q = (const bf16*)qkv; P.K = P.Q + C; P.V = P.Q + 2 * C; P.ldq = P.ldk = P.ldv = ld;
  P.O = (bf16*)O; P.ldo = ldo; P.lse = (float*)lse; P.H = H; P.hd = hd; P.Nq = ix.N; P.max_nk = ix.N; P.scale = scale;
  P.dO = (const bf16*)dO; P.dQ = (bf16*)dqkv; P.lddq = lddq;
  P.dK_lp = (bf16*)dqkv + C; P.dV_lp = (bf16*)dqkv + 2 * C; P.lddkv_lp = lddq;
  P.dtable = dtable;
  P.win = ix;
  VALOR_REQUIRE(Pn <= 65535 && H <= 65535, "window_attn: too many windows/heads");
  return hd == 32 ? launch_bwd<32, true>(P, delta, Pn, ix.N, ix.N, st) : launch_bwd<64, true>(P, delta, Pn, ix.N, ix.N, st);
}

}  // namespace valor
