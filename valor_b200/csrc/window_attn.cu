// valor_b200 — VideoSwin shifted-window attention, one CTA per (window, head).
//
// A 3-D window holds at most 8*7*7 = 392 tokens of 32 channels per head, so a whole attention
// problem (Q, K, V and, backward, dO) fits in shared memory: 4 x 400 x 64 B = 100 KB.  The CTA
// builds the window's index / bias / mask tables ONCE, gathers the token rows once (cyclic shift
// and window partition folded into the row index, videoswin.py:206-216), and then every warp works
// independently on 16-row blocks with no CTA barrier:
//
//   forward : a warp owns 16 queries and sweeps all keys with an online softmax; S and P stay in
//             registers; the row sums come out of the P.V MMA through an all-ones extra column.
//   backward: "dq warps" own 16-query blocks (dQ + relative-position-bias gradient), "dkv warps" own
//             16-key blocks (S^T = K.Q^T so P^T / dS^T are born in A-operand layout; dK, dV).
//
// Relative-position bias (videoswin.py:113-127,150-153): per token one code word; the bias of a
// pair is table[code_q - code_k].  Keys are enumerated (w,d,h) so the 32 lanes of an MMA fragment
// read 32 distinct table slots.  The -100 shift mask (videoswin.py:272-285) only exists in windows
// on the wrapped border of a shifted block; those CTAs run a MASKED instantiation, the others
// never test it.
//
// Bias gradient: shared-memory atomics cost ~1 LSU cycle per LANE on this part and dominated the
// backward pass.  Instead every dq warp owns a private fp32 copy of the table slice and folds its
// dS tile in row by row with plain load / add / store: inside one query row all keys map to distinct
// slots, so a warp-wide read-modify-write has no internal collision, and a warp's shared-memory
// instructions retire in order.  The copies are summed and sent to the global table once per CTA.
#include "common.cuh"
#include "attention.cuh"
#include <algorithm>

namespace valor {

namespace {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kMask = -100.0f;   // videoswin.py:284

// Shared-memory reads are volatile asm: they must stay behind the prologue barriers (a non-volatile asm without
// a memory operand may legally be hoisted above __syncthreads); ptxas still schedules them freely.
__device__ __forceinline__ uint32_t sm_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm4(uint32_t* r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t* r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// d = a.b + c, accumulator in place
__device__ __forceinline__ void mma_acc(float* c, const uint32_t* a, const uint32_t* b) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// d = a.b + {c0, c1, c2, c3}: the initial accumulator comes from other registers (no separate moves)
__device__ __forceinline__ void mma_init(float* d, const uint32_t* a, const uint32_t* b, float c0, float c1, float c2, float c3) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "f"(c0), "f"(c1), "f"(c2), "f"(c3));
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *(uint32_t*)&v;
}
__device__ __forceinline__ float ld_f32(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint2 ld_v2u32(uint32_t a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ float2 ld_v2f32(uint32_t a) { float2 v; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// Accesses to the warp-private read-modify-write regions (program order matters).
__device__ __forceinline__ float ldv_f32(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ void stv_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v)); }
__device__ __forceinline__ void stv_v2f32(uint32_t a, float x, float y) { asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(a), "f"(x), "f"(y)); }
__device__ __forceinline__ void cp16(uint32_t dst, const void* src, int nbytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
}
__device__ __forceinline__ void cp4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// exact i / d for 0 <= i < 4096 and 1 <= d <= 256 (multiply-shift with a rounded-up reciprocal)
__device__ __forceinline__ int small_div(int i, int inv) { return (int)(((unsigned)i * (unsigned)inv) >> 20); }
__host__ __device__ __forceinline__ int small_inv(int d) { return (int)(((1u << 20) + d - 1) / d); }

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
  int n_used, maxcode, center;     // bias slots reachable inside one window: [center-maxcode, center+maxcode]
  int tab_stride;                  // bytes between per-warp gradient tables (n_used*4 rounded up to 16)
  int n_dq_warps;                  // backward: warps 0..n_dq_warps-1 take the query blocks, the rest the key blocks
  WindowIndex win;
};

// Row tiles: 64 bytes per row (HD = 32 bf16), 16-byte chunk c of row r stored at chunk c ^ ((r >> 1) & 3): the 8 rows
// of an ldmatrix phase then cover all 32 banks.
constexpr int HD = 32;
constexpr int ROWB = HD * 2;
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) { return (uint32_t)(row * ROWB + ((chunk ^ ((row >> 1) & 3)) << 4)); }

struct Smem {
  unsigned char *Qs, *Ks, *Vs, *dOs;
  int *qrow, *krow;
  uint32_t *qcode, *kcode, *qreg, *kreg;
  float* tab;
};

// Per-token words: qcode = 4*(code + maxcode), kcode = 4*code (byte offsets; the pair's table slot is qcode - kcode),
// qreg / kreg = shift-mask region id (compute_mask, videoswin.py:272-285).
__device__ __forceinline__ bool win_build_tables(const WinParams& P, int p, int h, const Smem& S) {
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
  // a window carries more than one mask region only where a shifted axis wraps: the last window along that axis
  const bool masked = (ix.sd > 0 && id == nWd - 1) || (ix.sh > 0 && ih == nWh - 1) || (ix.sw > 0 && iw == nWw - 1);
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
        if (masked) reg = (uint32_t)(ix.region(cd, ix.D, ix.wd, ix.sd) * 9 + ix.region(ch, ix.H, ix.wh, ix.sh) * 3 +
                                     ix.region(cw, ix.W, ix.ww, ix.sw));
        const int code = ld * cH + lh * cW + lw;
        if (which == 0) { S.qrow[i] = row; S.qcode[i] = (uint32_t)(4 * (code + P.maxcode)); S.qreg[i] = reg; }
        else            { S.krow[i] = row; S.kcode[i] = (uint32_t)(4 * code); S.kreg[i] = reg; }
      }
    } else {
      S.qrow[i] = -1; S.krow[i] = -1;
      S.qcode[i] = (uint32_t)(4 * P.maxcode); S.kcode[i] = 0u; S.qreg[i] = 0u; S.kreg[i] = 0u;
    }
  }
  const float* src = ix.table + (size_t)(P.center - P.maxcode) * ix.heads + h;
  for (int r = threadIdx.x; r < P.n_used; r += blockDim.x) cp4(sm_addr(S.tab + r), src + (size_t)r * ix.heads);
  return masked;
}

// gather NP rows x 64 B into a swizzled shared tile (cp.async, zero-fill for padding rows)
__device__ __forceinline__ void win_load_rows(unsigned char* dst, const bf16* src, long long ld, int col0, const int* rows, int NP) {
  const uint32_t d0 = sm_addr(dst);
  for (int c = threadIdx.x; c < NP * 4; c += blockDim.x) {
    const int r = c >> 2, ch = c & 3;
    const int gr = rows[r];
    cp16(d0 + tile_off(r, ch), src + (size_t)(gr < 0 ? 0 : gr) * ld + col0 + ch * 8, gr < 0 ? 0 : 16);
  }
}

// ldmatrix addresses.  A-operand / non-transposed B-operand fragments of a 16-row block:
//   A (rows = M):   lane (m8, r8) -> row base + (m8&1)*8 + r8, chunk ks*2 + (m8>>1)
//   B (rows = N):   lane (m8, r8) -> row base + (m8>>1)*8 + r8, chunk ks*2 + (m8&1)         (two n-tiles per x4)
//   B^T (rows = K): lane (m8, r8) -> row base + (m8&1)*8 + r8, chunk dt + (m8>>1), .trans   (two n-tiles per x4)
// Row bases are multiples of 8, so the swizzle term only depends on r8.
struct Lane {
  int lane, g, t4, m8, r8;
  uint32_t a_off[2];    // A fragment, ks = 0,1 (relative to the 16-row block)
  uint32_t b_off[2];    // B fragment, ks = 0,1
  uint32_t bt_off[2];   // B^T fragment, dt = 0,2
  __device__ __forceinline__ void init() {
    lane = threadIdx.x & 31; g = lane >> 2; t4 = lane & 3; m8 = lane >> 3; r8 = lane & 7;
    const int sw = (r8 >> 1) & 3;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      a_off[ks] = (uint32_t)(((m8 & 1) * 8 + r8) * ROWB + (((ks * 2 + (m8 >> 1)) ^ sw) << 4));
      b_off[ks] = (uint32_t)(((m8 >> 1) * 8 + r8) * ROWB + (((ks * 2 + (m8 & 1)) ^ sw) << 4));
      bt_off[ks] = (uint32_t)(((m8 & 1) * 8 + r8) * ROWB + (((ks * 2 + (m8 >> 1)) ^ sw) << 4));   // dt = 2*ks
    }
  }
};

static inline size_t win_smem_bytes(int NP, int n_used, bool bwd, int n_dq_warps) {
  size_t b = (size_t)(bwd ? 4 : 3) * NP * ROWB;   // Q K V (dO)
  b += (size_t)6 * NP * 4;                        // qrow krow qcode kcode qreg kreg
  const size_t tab = ((size_t)n_used * 4 + 15) / 16 * 16;
  b += tab;                                       // bias slice
  if (bwd) b += (size_t)2 * NP * 4 + (size_t)n_dq_warps * (tab + 16 * 40 * 4);   // lse, -delta, per-warp tables + staging
  return b + 16;
}

__device__ __forceinline__ Smem carve(unsigned char* smem, int NP, bool bwd) {
  Smem S;
  S.Qs = smem;
  S.Ks = S.Qs + NP * ROWB;
  S.Vs = S.Ks + NP * ROWB;
  S.dOs = S.Vs + NP * ROWB;
  S.qrow = (int*)(bwd ? S.dOs + NP * ROWB : S.dOs);
  S.krow = S.qrow + NP;
  S.qcode = (uint32_t*)(S.krow + NP);
  S.kcode = S.qcode + NP;
  S.qreg = S.kcode + NP;
  S.kreg = S.qreg + NP;
  S.tab = (float*)(S.kreg + NP);
  return S;
}

// ==========================================================================================
// forward
// ==========================================================================================
// One 64-key block of one 16-query block.  TAIL: the block holds padding keys and fewer than four 16-key pairs.
template <bool MASKED, bool TAIL>
__device__ __forceinline__ void fwd_block(const Lane& L, uint32_t Kb, uint32_t Vb, uint32_t kcode_b, uint32_t kreg_b,
                                          int keys_left, const uint32_t (&qf)[2][4], const uint32_t (&qaddr)[2],
                                          const uint32_t (&qrg)[2], float scale, float (&o)[5][4], float (&mrow)[2]) {
  const int npair = TAIL ? min(4, (keys_left + 15) >> 4) : 4;
  float s[8][4];
#pragma unroll
  for (int pr = 0; pr < 4; ++pr)
    if (!TAIL || pr < npair) {
      uint32_t b[4];
      ldsm4(b, Kb + pr * 16 * ROWB + L.b_off[0]);
      mma_init(s[pr * 2], qf[0], b, 0.f, 0.f, 0.f, 0.f);
      mma_init(s[pr * 2 + 1], qf[0], b + 2, 0.f, 0.f, 0.f, 0.f);
      ldsm4(b, Kb + pr * 16 * ROWB + L.b_off[1]);
      mma_acc(s[pr * 2], qf[1], b);
      mma_acc(s[pr * 2 + 1], qf[1], b + 2);
    }
  float mnew[2] = {mrow[0], mrow[1]};
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
    if (!TAIL || (nt >> 1) < npair) {
      const uint2 kc = ld_v2u32(kcode_b + 4u * (uint32_t)(nt * 8 + L.t4 * 2));
      uint2 kr = make_uint2(0u, 0u);
      if (MASKED) kr = ld_v2u32(kreg_b + 4u * (uint32_t)(nt * 8 + L.t4 * 2));
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = e >> 1;
        float v = fmaf(s[nt][e], scale, ld_f32(qaddr[r] - ((e & 1) ? kc.y : kc.x)));
        if (MASKED && qrg[r] != ((e & 1) ? kr.y : kr.x)) v += kMask;
        if (TAIL && nt * 8 + L.t4 * 2 + (e & 1) >= keys_left) v = -INFINITY;
        s[nt][e] = v;
        mnew[r] = fmaxf(mnew[r], v);
      }
    }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    mnew[r] = fmaxf(mnew[r], __shfl_xor_sync(0xffffffffu, mnew[r], 1));
    mnew[r] = fmaxf(mnew[r], __shfl_xor_sync(0xffffffffu, mnew[r], 2));
  }
  float ml[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {   // block 0 always holds real keys, so mnew is finite from the first block on
    const float corr = ex2((mrow[r] - mnew[r]) * kLog2e);
    mrow[r] = mnew[r];
    ml[r] = -mnew[r] * kLog2e;
#pragma unroll
    for (int i = 0; i < 5; ++i) { o[i][r * 2] *= corr; o[i][r * 2 + 1] *= corr; }
  }
  const uint32_t ones[2] = {0x3f803f80u, 0x3f803f80u};   // bf16 1.0 pairs: column block of ones -> row sums of P
#pragma unroll
  for (int kk = 0; kk < 4; ++kk)
    if (!TAIL || kk < npair) {
      uint32_t pf[4];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int nt = kk * 2 + hf;
        float pv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) pv[e] = ex2(fmaf(s[nt][e], kLog2e, ml[e >> 1]));
        pf[hf * 2 + 0] = pack2(pv[0], pv[1]);
        pf[hf * 2 + 1] = pack2(pv[2], pv[3]);
      }
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        uint32_t b[4];
        ldsm4t(b, Vb + kk * 16 * ROWB + L.bt_off[ks]);
        mma_acc(o[ks * 2], pf, b);
        mma_acc(o[ks * 2 + 1], pf, b + 2);
      }
      mma_acc(o[4], pf, ones);
    }
}

template <bool MASKED>
__device__ __forceinline__ void fwd_rows(const WinParams& P, const Smem& S, const Lane& L, int p, int h, int rb) {
  const int N = P.win.N;
  const uint32_t Qs = sm_addr(S.Qs), Ks = sm_addr(S.Ks), Vs = sm_addr(S.Vs);
  const uint32_t tab_s = sm_addr(S.tab), kcode_s = sm_addr(S.kcode), kreg_s = sm_addr(S.kreg);
  uint32_t qf[2][4];
  ldsm4(qf[0], Qs + rb * 16 * ROWB + L.a_off[0]);
  ldsm4(qf[1], Qs + rb * 16 * ROWB + L.a_off[1]);
  const int i0 = rb * 16 + L.g;
  const uint32_t qaddr[2] = {tab_s + S.qcode[i0], tab_s + S.qcode[i0 + 8]};
  const uint32_t qrg[2] = {S.qreg[i0], S.qreg[i0 + 8]};
  float o[5][4];
#pragma unroll
  for (int i = 0; i < 5; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float mrow[2] = {-INFINITY, -INFINITY};
  const int nfull = N >> 6;
#pragma unroll 1
  for (int kb = 0; kb < nfull; ++kb)
    fwd_block<MASKED, false>(L, Ks + kb * 64 * ROWB, Vs + kb * 64 * ROWB, kcode_s + kb * 256, kreg_s + kb * 256, 64, qf,
                             qaddr, qrg, P.scale, o, mrow);
  if (N & 63)
    fwd_block<MASKED, true>(L, Ks + nfull * 64 * ROWB, Vs + nfull * 64 * ROWB, kcode_s + nfull * 256, kreg_s + nfull * 256,
                            N - nfull * 64, qf, qaddr, qrg, P.scale, o, mrow);
  const int col0 = h * HD;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int i = i0 + r * 8;
    if (i < N) {
      const float l = o[4][r * 2];   // every column of the ones block carries the row sum
      const float inv = 1.f / l;
      bf16* dst = P.O + (size_t)S.qrow[i] * P.ldo + col0;
#pragma unroll
      for (int dt = 0; dt < 4; ++dt)
        *(uint32_t*)(dst + dt * 8 + L.t4 * 2) = pack2(o[dt][r * 2] * inv, o[dt][r * 2 + 1] * inv);
      if (L.t4 == 0) P.lse[((size_t)p * P.heads + h) * N + i] = mrow[r] + log2f(l) * kLn2;
    }
  }
}

__global__ void __launch_bounds__(256, 2)
window_fwd_kernel(WinParams P) {
  extern __shared__ __align__(16) unsigned char smem[];
  const Smem S = carve(smem, P.NP, false);
  const int p = blockIdx.x, h = blockIdx.y;
  const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int C = P.heads * HD, col0 = h * HD;
  const bool masked = win_build_tables(P, p, h, S);
  __syncthreads();
  win_load_rows(S.Qs, P.qkv, P.ld, col0, S.qrow, P.NP);
  win_load_rows(S.Ks, P.qkv + C, P.ld, col0, S.krow, P.NP);
  win_load_rows(S.Vs, P.qkv + 2 * C, P.ld, col0, S.krow, P.NP);
  cp_commit();
  cp_wait_all();
  __syncthreads();
  Lane L;
  L.init();
  const int nrb = P.NP >> 4;
  if (masked) {
    for (int rb = warp; rb < nrb; rb += nwarps) fwd_rows<true>(P, S, L, p, h, rb);
  } else {
    for (int rb = warp; rb < nrb; rb += nwarps) fwd_rows<false>(P, S, L, p, h, rb);
  }
}

// ==========================================================================================
// backward
// ==========================================================================================
// dQ for one 16-query block; the block's dS also goes into this warp's private bias-gradient table.
template <bool MASKED>
__device__ __forceinline__ void bwd_dq_unit(const WinParams& P, const Smem& S, const Lane& L, int h, int rb, uint32_t lse_s,
                                            uint32_t ndel_s, uint32_t gtab_s, uint32_t stg_s) {
  const int N = P.win.N;
  const uint32_t Qs = sm_addr(S.Qs), Ks = sm_addr(S.Ks), Vs = sm_addr(S.Vs), dOs = sm_addr(S.dOs);
  const uint32_t tab_s = sm_addr(S.tab), kcode_s = sm_addr(S.kcode), kreg_s = sm_addr(S.kreg);
  uint32_t qf[2][4], dof[2][4];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    ldsm4(qf[ks], Qs + rb * 16 * ROWB + L.a_off[ks]);
    ldsm4(dof[ks], dOs + rb * 16 * ROWB + L.a_off[ks]);
  }
  const int i0 = rb * 16 + L.g;
  const uint32_t qaddr[2] = {tab_s + S.qcode[i0], tab_s + S.qcode[i0 + 8]};
  const uint32_t qrg[2] = {S.qreg[i0], S.qreg[i0 + 8]};
  const float nlse[2] = {-ld_f32(lse_s + 4u * i0) * kLog2e, -ld_f32(lse_s + 4u * (i0 + 8)) * kLog2e};   // -inf on padding rows
  const float ndel[2] = {ld_f32(ndel_s + 4u * i0), ld_f32(ndel_s + 4u * (i0 + 8))};
  // row r of the block is folded into the gradient table at byte offset qcode(row r) - kcode(key)
  const uint32_t my_qcode = S.qcode[rb * 16 + (L.lane & 15)];
  const bool want_dtab = P.dtable != nullptr;
  float dq[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
  const int nsweep = (N + 31) >> 5;
#pragma unroll 1
  for (int hb = 0; hb < nsweep; ++hb) {
    const int k0 = hb * 32;
    const int left = N - k0;                        // real keys from k0 on
    const int npair = min(2, (left + 15) >> 4);
    const uint32_t Kb = Ks + k0 * ROWB, Vb = Vs + k0 * ROWB;
    float s[4][4], dp[4][4];
#pragma unroll
    for (int pr = 0; pr < 2; ++pr)
      if (pr < npair) {
        uint32_t b[4];
        ldsm4(b, Kb + pr * 16 * ROWB + L.b_off[0]);
        mma_init(s[pr * 2], qf[0], b, 0.f, 0.f, 0.f, 0.f);
        mma_init(s[pr * 2 + 1], qf[0], b + 2, 0.f, 0.f, 0.f, 0.f);
        ldsm4(b, Kb + pr * 16 * ROWB + L.b_off[1]);
        mma_acc(s[pr * 2], qf[1], b);
        mma_acc(s[pr * 2 + 1], qf[1], b + 2);
        ldsm4(b, Vb + pr * 16 * ROWB + L.b_off[0]);   // dP - delta: the accumulator starts at -delta_i
        mma_init(dp[pr * 2], dof[0], b, ndel[0], ndel[0], ndel[1], ndel[1]);
        mma_init(dp[pr * 2 + 1], dof[0], b + 2, ndel[0], ndel[0], ndel[1], ndel[1]);
        ldsm4(b, Vb + pr * 16 * ROWB + L.b_off[1]);
        mma_acc(dp[pr * 2], dof[1], b);
        mma_acc(dp[pr * 2 + 1], dof[1], b + 2);
      }
    uint32_t dsf[2][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      float ds[4] = {0.f, 0.f, 0.f, 0.f};
      if ((nt >> 1) < npair) {
        const uint2 kc = ld_v2u32(kcode_s + 4u * (uint32_t)(k0 + nt * 8 + L.t4 * 2));
        uint2 kr = make_uint2(0u, 0u);
        if (MASKED) kr = ld_v2u32(kreg_s + 4u * (uint32_t)(k0 + nt * 8 + L.t4 * 2));
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int r = e >> 1;
          float v = fmaf(s[nt][e], P.scale, ld_f32(qaddr[r] - ((e & 1) ? kc.y : kc.x)));
          if (MASKED && qrg[r] != ((e & 1) ? kr.y : kr.x)) v += kMask;
          float pr_ = ex2(fmaf(v, kLog2e, nlse[r]));
          if (nt * 8 + L.t4 * 2 + (e & 1) >= left) pr_ = 0.f;   // padding key (only ever true in the last sweep)
          ds[e] = pr_ * dp[nt][e];
        }
      }
      if (want_dtab) {   // stage the fp32 tile [16 rows][32 keys], row pitch 40 floats
        stv_v2f32(stg_s + 4u * (uint32_t)(L.g * 40 + nt * 8 + L.t4 * 2), ds[0], ds[1]);
        stv_v2f32(stg_s + 4u * (uint32_t)((L.g + 8) * 40 + nt * 8 + L.t4 * 2), ds[2], ds[3]);
      }
      dsf[nt >> 1][(nt & 1) * 2 + 0] = pack2(ds[0], ds[1]);
      dsf[nt >> 1][(nt & 1) * 2 + 1] = pack2(ds[2], ds[3]);
    }
#pragma unroll
    for (int kk = 0; kk < 2; ++kk)
      if (kk < npair) {
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          uint32_t b[4];
          ldsm4t(b, Kb + kk * 16 * ROWB + L.bt_off[ks]);
          mma_acc(dq[ks * 2], dsf[kk], b);
          mma_acc(dq[ks * 2 + 1], dsf[kk], b + 2);
        }
      }
    if (want_dtab) {
      // fold the staged tile into the private table, one query row per step: lane l owns key k0 + l, and
      // within a row distinct keys hit distinct slots
      __syncwarp();
      const bool live = L.lane < left;
      const uint32_t kc_l = S.kcode[min(k0 + L.lane, P.NP - 1)];
#pragma unroll 4
      for (int r = 0; r < 16; ++r) {
        const uint32_t slot = gtab_s + __shfl_sync(0xffffffffu, my_qcode, r) - kc_l;
        const float add = ldv_f32(stg_s + 4u * (uint32_t)(r * 40 + L.lane));
        if (live) stv_f32(slot, ldv_f32(slot) + add);
        __syncwarp();
      }
    }
  }
  const int col0 = h * HD;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int i = i0 + r * 8;
    if (i < N) {
      bf16* dst = P.dqkv + (size_t)S.qrow[i] * P.lddqkv + col0;
#pragma unroll
      for (int dt = 0; dt < 4; ++dt)
        *(uint32_t*)(dst + dt * 8 + L.t4 * 2) = pack2(dq[dt][r * 2] * P.scale, dq[dt][r * 2 + 1] * P.scale);
    }
  }
}

// dK, dV for one 16-key block.
template <bool MASKED>
__device__ __forceinline__ void bwd_dkv_unit(const WinParams& P, const Smem& S, const Lane& L, int h, int jb, uint32_t lse_s,
                                             uint32_t ndel_s) {
  const int N = P.win.N;
  const uint32_t Qs = sm_addr(S.Qs), Ks = sm_addr(S.Ks), Vs = sm_addr(S.Vs), dOs = sm_addr(S.dOs);
  const uint32_t tab_s = sm_addr(S.tab), qcode_s = sm_addr(S.qcode), qreg_s = sm_addr(S.qreg);
  uint32_t kf[2][4], vf[2][4];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    ldsm4(kf[ks], Ks + jb * 16 * ROWB + L.a_off[ks]);
    ldsm4(vf[ks], Vs + jb * 16 * ROWB + L.a_off[ks]);
  }
  const int j0 = jb * 16 + L.g;   // keys j0, j0+8 (padding keys: K/V rows are zero, results dropped)
  const uint32_t kaddr[2] = {tab_s - S.kcode[j0], tab_s - S.kcode[j0 + 8]};
  const uint32_t krg[2] = {S.kreg[j0], S.kreg[j0 + 8]};
  float dk[4][4], dv[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) dk[i][e] = dv[i][e] = 0.f;
  const int nsweep = (N + 31) >> 5;
#pragma unroll 1
  for (int hb = 0; hb < nsweep; ++hb) {
    const int q0 = hb * 32;
    const int npair = min(2, (N - q0 + 15) >> 4);
    const uint32_t Qb = Qs + q0 * ROWB, dOb = dOs + q0 * ROWB;
    float s[4][4], dp[4][4];   // rows = keys (g, g+8), cols = queries
#pragma unroll
    for (int pr = 0; pr < 2; ++pr)
      if (pr < npair) {
        uint32_t b[4];
        const float2 nd0 = ld_v2f32(ndel_s + 4u * (uint32_t)(q0 + pr * 16 + L.t4 * 2));       // -delta of the two query columns
        const float2 nd1 = ld_v2f32(ndel_s + 4u * (uint32_t)(q0 + pr * 16 + 8 + L.t4 * 2));
        ldsm4(b, Qb + pr * 16 * ROWB + L.b_off[0]);
        mma_init(s[pr * 2], kf[0], b, 0.f, 0.f, 0.f, 0.f);
        mma_init(s[pr * 2 + 1], kf[0], b + 2, 0.f, 0.f, 0.f, 0.f);
        ldsm4(b, Qb + pr * 16 * ROWB + L.b_off[1]);
        mma_acc(s[pr * 2], kf[1], b);
        mma_acc(s[pr * 2 + 1], kf[1], b + 2);
        ldsm4(b, dOb + pr * 16 * ROWB + L.b_off[0]);
        mma_init(dp[pr * 2], vf[0], b, nd0.x, nd0.y, nd0.x, nd0.y);
        mma_init(dp[pr * 2 + 1], vf[0], b + 2, nd1.x, nd1.y, nd1.x, nd1.y);
        ldsm4(b, dOb + pr * 16 * ROWB + L.b_off[1]);
        mma_acc(dp[pr * 2], vf[1], b);
        mma_acc(dp[pr * 2 + 1], vf[1], b + 2);
      }
    uint32_t pf[2][4], dsf[2][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      float pv[4] = {0.f, 0.f, 0.f, 0.f}, ds[4] = {0.f, 0.f, 0.f, 0.f};
      if ((nt >> 1) < npair) {
        const uint32_t ib = (uint32_t)(q0 + nt * 8 + L.t4 * 2);   // two consecutive queries (columns of S^T)
        const uint2 qc = ld_v2u32(qcode_s + 4u * ib);
        const float2 l2 = ld_v2f32(lse_s + 4u * ib);              // +inf on padding queries -> p = 0
        uint2 qr = make_uint2(0u, 0u);
        if (MASKED) qr = ld_v2u32(qreg_s + 4u * ib);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int r = e >> 1;
          float v = fmaf(s[nt][e], P.scale, ld_f32(kaddr[r] + ((e & 1) ? qc.y : qc.x)));
          if (MASKED && krg[r] != ((e & 1) ? qr.y : qr.x)) v += kMask;
          const float pr_ = ex2((v - ((e & 1) ? l2.y : l2.x)) * kLog2e);
          pv[e] = pr_;
          ds[e] = pr_ * dp[nt][e];
        }
      }
      pf[nt >> 1][(nt & 1) * 2 + 0] = pack2(pv[0], pv[1]);
      pf[nt >> 1][(nt & 1) * 2 + 1] = pack2(pv[2], pv[3]);
      dsf[nt >> 1][(nt & 1) * 2 + 0] = pack2(ds[0], ds[1]);
      dsf[nt >> 1][(nt & 1) * 2 + 1] = pack2(ds[2], ds[3]);
    }
#pragma unroll
    for (int kk = 0; kk < 2; ++kk)   // contraction over the 16-query pairs of this sweep
      if (kk < npair) {
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          uint32_t b[4];
          ldsm4t(b, dOb + kk * 16 * ROWB + L.bt_off[ks]);
          mma_acc(dv[ks * 2], pf[kk], b);
          mma_acc(dv[ks * 2 + 1], pf[kk], b + 2);
          ldsm4t(b, Qb + kk * 16 * ROWB + L.bt_off[ks]);
          mma_acc(dk[ks * 2], dsf[kk], b);
          mma_acc(dk[ks * 2 + 1], dsf[kk], b + 2);
        }
      }
  }
  const int C = P.heads * HD, col0 = h * HD;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int j = j0 + r * 8;
    if (j < N) {
      bf16* dstk = P.dqkv + (size_t)S.krow[j] * P.lddqkv + C + col0;
      bf16* dstv = dstk + C;
#pragma unroll
      for (int dt = 0; dt < 4; ++dt) {
        *(uint32_t*)(dstk + dt * 8 + L.t4 * 2) = pack2(dk[dt][r * 2] * P.scale, dk[dt][r * 2 + 1] * P.scale);
        *(uint32_t*)(dstv + dt * 8 + L.t4 * 2) = pack2(dv[dt][r * 2], dv[dt][r * 2 + 1]);
      }
    }
  }
}

__global__ void __launch_bounds__(448, 1)
window_bwd_kernel(WinParams P) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int NP = P.NP;
  const Smem S = carve(smem, NP, true);
  float* lse_sm = (float*)((unsigned char*)S.tab + P.tab_stride);   // natural-log lse; +inf on padding rows
  float* ndel_sm = lse_sm + NP;                                      // -delta_i
  unsigned char* priv = (unsigned char*)(ndel_sm + NP);              // per dq warp: [tab_stride] table + [16][40] staging
  const int priv_stride = P.tab_stride + 16 * 40 * 4;
  const int p = blockIdx.x, h = blockIdx.y;
  const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int C = P.heads * HD, col0 = h * HD;
  const int N = P.win.N;
  const bool masked = win_build_tables(P, p, h, S);
  __syncthreads();
  win_load_rows(S.Qs, P.qkv, P.ld, col0, S.qrow, NP);
  win_load_rows(S.Ks, P.qkv + C, P.ld, col0, S.krow, NP);
  win_load_rows(S.Vs, P.qkv + 2 * C, P.ld, col0, S.krow, NP);
  win_load_rows(S.dOs, P.dO, P.ldo, col0, S.qrow, NP);
  cp_commit();
  for (int i = threadIdx.x; i < NP; i += blockDim.x) lse_sm[i] = i < N ? P.lse[((size_t)p * P.heads + h) * N + i] : INFINITY;
  if (P.dtable != nullptr) {
    float* z = (float*)priv;
    const int nz = P.n_dq_warps * priv_stride / 4;
    for (int i = threadIdx.x; i < nz; i += blockDim.x) z[i] = 0.f;
  }
  cp_wait_all();
  __syncthreads();
  // delta_i = dO_i . O_i: dO from the staged tile, O rows straight from global; four lanes per row
  for (int c = threadIdx.x; c < NP * 4; c += blockDim.x) {   // NP*4 is a multiple of 32: whole warps iterate
    const int r = c >> 2, ch = c & 3;
    const int gr = S.qrow[r];
    uint4 o4 = make_uint4(0, 0, 0, 0);
    if (gr >= 0) o4 = *(const uint4*)(P.O + (size_t)gr * P.ldo + col0 + ch * 8);
    const uint4 a = *(const uint4*)(S.dOs + tile_off(r, ch));
    const __nv_bfloat162* pa = (const __nv_bfloat162*)&a;
    const __nv_bfloat162* po = (const __nv_bfloat162*)&o4;
    float d = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 fa = __bfloat1622float2(pa[j]), fo = __bfloat1622float2(po[j]);
      d += fa.x * fo.x + fa.y * fo.y;
    }
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    d += __shfl_xor_sync(0xffffffffu, d, 2);
    if (ch == 0) ndel_sm[r] = -d;
  }
  __syncthreads();
  Lane L;
  L.init();
  const int nrb = NP >> 4;
  const uint32_t lse_s = sm_addr(lse_sm), ndel_s = sm_addr(ndel_sm);
  const int nq = P.n_dq_warps;
  if (warp < nq) {
    const uint32_t gtab_s = sm_addr(priv + warp * priv_stride);
    const uint32_t stg_s = gtab_s + P.tab_stride;
    if (masked) { for (int rb = warp; rb < nrb; rb += nq) bwd_dq_unit<true>(P, S, L, h, rb, lse_s, ndel_s, gtab_s, stg_s); }
    else        { for (int rb = warp; rb < nrb; rb += nq) bwd_dq_unit<false>(P, S, L, h, rb, lse_s, ndel_s, gtab_s, stg_s); }
  } else {
    const int nk = nwarps - nq;
    if (masked) { for (int jb = warp - nq; jb < nrb; jb += nk) bwd_dkv_unit<true>(P, S, L, h, jb, lse_s, ndel_s); }
    else        { for (int jb = warp - nq; jb < nrb; jb += nk) bwd_dkv_unit<false>(P, S, L, h, jb, lse_s, ndel_s); }
  }
  if (P.dtable != nullptr) {
    __syncthreads();
    const int r0 = P.center - P.maxcode;
    for (int r = threadIdx.x; r < P.n_used; r += blockDim.x) {
      float acc = 0.f;
      for (int w = 0; w < nq; ++w) acc += ((const float*)(priv + w * priv_stride))[r];
      if (acc != 0.f) atomicAdd(&P.dtable[(size_t)(r0 + r) * P.win.heads + h], acc);
    }
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static constexpr size_t kSmemLimit = 227 * 1024;

static void win_geometry(WinParams& P, const WindowIndex& ix, int H) {
  P.win = ix;
  P.heads = H;
  P.NP = (ix.N + 15) / 16 * 16;
  const int cW = 2 * ix.WW - 1, cH = (2 * ix.WH - 1) * cW;
  P.center = (ix.WD - 1) * cH + (ix.WH - 1) * cW + (ix.WW - 1);
  P.maxcode = (ix.wd - 1) * cH + (ix.wh - 1) * cW + (ix.ww - 1);
  P.n_used = 2 * P.maxcode + 1;
  P.tab_stride = (P.n_used * 4 + 15) / 16 * 16;
}

// forward: row blocks dealt round-robin to at most 8 warps
static int fwd_warps(int nrb) {
  const int per = (nrb + 7) / 8;
  return (nrb + per - 1) / per;
}

// backward: split up to 14 warps between query blocks (dQ + bias gradient, ~1.35x the work of a key block) and key
// blocks (dK, dV) so that the slower group finishes earliest; every dq warp needs a private table in shared memory
static bool bwd_warps(int NP, int n_used, int& n_dq, int& n_dkv) {
  const int nrb = NP / 16;
  double best = 1e30;
  n_dq = n_dkv = 0;
  for (int q = 1; q <= 13; ++q)
    for (int k = 1; q + k <= 14; ++k) {
      if (win_smem_bytes(NP, n_used, true, q) > kSmemLimit) continue;
      const double cost = std::max(1.35 * ((nrb + q - 1) / q), 1.0 * ((nrb + k - 1) / k)) + 1e-3 * (q + k);
      if (cost < best) { best = cost; n_dq = q; n_dkv = k; }
    }
  return n_dq > 0;
}

bool window_cta_eligible(const WindowIndex& ix, int hd) {
  if (hd != HD || ix.N > 4095) return false;
  WinParams P = {};
  win_geometry(P, ix, 1);
  if (4 * (2 * P.maxcode + 1) >= 65536) return false;
  if (ix.wh * ix.ww > 256 || ix.wd * ix.wh > 256) return false;   // small_div range
  int q, k;
  return win_smem_bytes(P.NP, P.n_used, false, 0) <= kSmemLimit && bwd_warps(P.NP, P.n_used, q, k);
}

int window_cta_fwd(const WindowIndex& ix, const void* qkv, long long ld, void* O, long long ldo, float* lse, int Pn,
                   int H, int hd, float scale, cudaStream_t st) {
  VALOR_REQUIRE(hd == HD && H <= 65535, "window_cta_fwd: head dim 32 only");
  WinParams P = {};
  win_geometry(P, ix, H);
  P.qkv = (const bf16*)qkv; P.ld = ld; P.O = (bf16*)O; P.ldo = ldo; P.lse = lse; P.scale = scale;
  const size_t smem = win_smem_bytes(P.NP, P.n_used, false, 0);
  static size_t attr = 0;
  if (smem > attr) { VALOR_CUDA(cudaFuncSetAttribute(window_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = smem; }
  window_fwd_kernel<<<dim3(Pn, H), fwd_warps(P.NP / 16) * 32, smem, st>>>(P);
  return check_launch("window_fwd_kernel");
}

int window_cta_bwd(const WindowIndex& ix, const void* qkv, long long ld, const void* O, const void* dO, long long ldo,
                   const float* lse, void* dqkv, long long lddqkv, float* dtable, int Pn, int H, int hd, float scale,
                   cudaStream_t st) {
  VALOR_REQUIRE(hd == HD && H <= 65535, "window_cta_bwd: head dim 32 only");
  WinParams P = {};
  win_geometry(P, ix, H);
  P.qkv = (const bf16*)qkv; P.ld = ld; P.O = (bf16*)O; P.ldo = ldo; P.lse = (float*)lse; P.scale = scale;
  P.dO = (const bf16*)dO; P.dqkv = (bf16*)dqkv; P.lddqkv = lddqkv; P.dtable = dtable;
  int n_dq = 0, n_dkv = 0;
  VALOR_REQUIRE(bwd_warps(P.NP, P.n_used, n_dq, n_dkv), "window_cta_bwd: window does not fit in shared memory");
  P.n_dq_warps = n_dq;
  const size_t smem = win_smem_bytes(P.NP, P.n_used, true, n_dq);
  static size_t attr = 0;
  if (smem > attr) { VALOR_CUDA(cudaFuncSetAttribute(window_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = smem; }
  window_bwd_kernel<<<dim3(Pn, H), (n_dq + n_dkv) * 32, smem, st>>>(P);
  return check_launch("window_bwd_kernel");
}

}  // namespace valor
