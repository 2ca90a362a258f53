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
// pair is table[code_q - code_k]; with the natural (d,h,w) order on both sides the 32 lanes of an
// MMA fragment read <= 14 consecutive slots (one shared-memory wavefront).  The -100 shift mask (videoswin.py:272-285) only exists in windows
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
#include <utility>

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

// dq warp w of n_dq owns the contiguous 16-query blocks [start, start + cnt): contiguous rows have a narrow span of
// query codes, so the warp's private gradient table only needs  span + maxcode + 1  slots instead of 2*maxcode + 1.
__host__ __device__ __forceinline__ void dq_range(int w, int n_dq, int nrb, int& start, int& cnt) {
  const int base = nrb / n_dq, rem = nrb % n_dq;
  start = w * base + (w < rem ? w : rem);
  cnt = base + (w < rem ? 1 : 0);
}

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
  int tab_stride;                  // bytes of the bias slice (n_used*4 rounded up to 16)
  int n_dq_warps;                  // backward: warps 0..n_dq_warps-1 take the query blocks, the rest the key blocks
  int gtab_bytes;                  // backward: bytes of one dq warp's private gradient table (see dq_range)
  WindowIndex win;
};

// Row tiles: 64 bytes per row (HD = 32 bf16), 16-byte chunk c of row r stored at chunk c ^ ((r >> 1) & 3): the 8 rows
// of an ldmatrix phase then cover all 32 banks.
constexpr int HD = 32;
constexpr int ROWB = HD * 2;
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) { return (uint32_t)(row * ROWB + ((chunk ^ ((row >> 1) & 3)) << 4)); }

struct Smem {
  unsigned char *Qs, *Ks, *Vs, *dOs;
  int* qrow;                       // global token row of local index i (queries and keys alike)
  uint32_t *qcode, *kcode, *qreg;
  float* tab;
};

// Per-token words: qcode = 4*(code + maxcode), kcode = 4*code (byte offsets; the pair's table slot is qcode - kcode),
// qreg = shift-mask region id (compute_mask, videoswin.py:272-285).
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
  const int hw = ix.wh * ix.ww;
  const int inv_hw = small_inv(hw), inv_ww = small_inv(ix.ww);
  // a window carries more than one mask region only where a shifted axis wraps: the last window along that axis
  const bool masked = (ix.sd > 0 && id == nWd - 1) || (ix.sh > 0 && ih == nWh - 1) || (ix.sw > 0 && iw == nWw - 1);
  // Queries and keys share the natural (d,h,w) enumeration: the 8 queries x 4 keys of one MMA fragment register then
  // touch at most 14 distinct, consecutive table slots (same address = broadcast), i.e. one wavefront per lookup.
  for (int i = threadIdx.x; i < P.NP; i += blockDim.x) {
    if (i < ix.N) {
      const int ld = small_div(i, inv_hw);
      const int rem = i - ld * hw;
      const int lh = small_div(rem, inv_ww);
      const int lw = rem - lh * ix.ww;
      const int cd = od + ld, ch = oh + lh, cw = ow + lw;
      int d = cd + ix.sd; if (d >= ix.D) d -= ix.D;   // shifted[c] = x[(c + shift) mod size]  (videoswin.py:206)
      int hh = ch + ix.sh; if (hh >= ix.H) hh -= ix.H;
      int w = cw + ix.sw; if (w >= ix.W) w -= ix.W;
      uint32_t reg = 0;
      if (masked) reg = (uint32_t)(ix.region(cd, ix.D, ix.wd, ix.sd) * 9 + ix.region(ch, ix.H, ix.wh, ix.sh) * 3 +
                                   ix.region(cw, ix.W, ix.ww, ix.sw));
      const int code = ld * cH + lh * cW + lw;
      S.qrow[i] = ((b * ix.D + d) * ix.H + hh) * ix.W + w;
      S.qcode[i] = (uint32_t)(4 * (code + P.maxcode));
      S.kcode[i] = (uint32_t)(4 * code);
      S.qreg[i] = reg;
    } else {
      S.qrow[i] = -1;   // padding rows reuse the last token's query code: their (zero) gradient folds stay inside the
      S.qcode[i] = (uint32_t)(8 * P.maxcode); S.kcode[i] = 0u; S.qreg[i] = 0u;   // owning warp's private table range
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

static inline size_t win_smem_bytes(int NP, int n_used, bool bwd, int n_dq_warps, int gtab_bytes = 0) {
  size_t b = (size_t)(bwd ? 4 : 3) * NP * ROWB;   // Q K V (dO)
  b += (size_t)4 * NP * 4;                        // qrow qcode kcode qreg
  const size_t tab = ((size_t)n_used * 4 + 15) / 16 * 16;
  b += tab;                                       // bias slice
  if (bwd) b += (size_t)2 * NP * 4 + 128 + (size_t)n_dq_warps * (gtab_bytes + 16 * 40 * 4 + 16);   // {lse, -delta}, table ranges, per-warp tables + staging + parking word
  return b + 16;
}

__device__ __forceinline__ Smem carve(unsigned char* smem, int NP, bool bwd) {
  Smem S;
  S.Qs = smem;
  S.Ks = S.Qs + NP * ROWB;
  S.Vs = S.Ks + NP * ROWB;
  S.dOs = S.Vs + NP * ROWB;
  S.qrow = (int*)(bwd ? S.dOs + NP * ROWB : S.dOs);
  S.qcode = (uint32_t*)(S.qrow + NP);
  S.kcode = S.qcode + NP;
  S.qreg = S.kcode + NP;
  S.tab = (float*)(S.qreg + NP);
  return S;
}

// compile-time loop: the index arrives as std::integral_constant, so it can feed "n" asm operands
template <int... I, class F>
__device__ __forceinline__ void static_for_impl(F&& f, std::integer_sequence<int, I...>) { (f(std::integral_constant<int, I>{}), ...); }
template <int N, class F>
__device__ __forceinline__ void static_for(F&& f) { static_for_impl(f, std::make_integer_sequence<int, N>{}); }

// [register + immediate] forms: one running per-lane base register serves a whole key / query sweep
template <int OFF> __device__ __forceinline__ void ldsm4_o(uint32_t* r, uint32_t base) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4+%5];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(base), "n"(OFF));
}
template <int OFF> __device__ __forceinline__ void ldsm4t_o(uint32_t* r, uint32_t base) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4+%5];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(base), "n"(OFF));
}
template <int OFF> __device__ __forceinline__ uint2 ld_v2u32_o(uint32_t base) {
  uint2 v; asm volatile("ld.shared.v2.u32 {%0,%1}, [%2+%3];" : "=r"(v.x), "=r"(v.y) : "r"(base), "n"(OFF)); return v;
}
template <int OFF> __device__ __forceinline__ float4 ld_v4f32_o(uint32_t base) {
  float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(base), "n"(OFF)); return v;
}
template <int OFF> __device__ __forceinline__ uint32_t ldv_u32_o(uint32_t base) {
  uint32_t v; asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(base), "n"(OFF)); return v;
}
template <int OFF> __device__ __forceinline__ float ldv_f32_o(uint32_t base) {
  float v; asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(base), "n"(OFF)); return v;
}
template <int OFF> __device__ __forceinline__ void stv_v2f32_o(uint32_t base, float x, float y) {
  asm volatile("st.shared.v2.f32 [%0+%1], {%2,%3};" ::"r"(base), "n"(OFF), "f"(x), "f"(y));
}

// ==========================================================================================
// forward
// ==========================================================================================
// Running state of one 16-query block while it sweeps the keys.
struct FwdRow {
  uint32_t qf[2][4];
  uint32_t qaddr[2], qrg[2];
  float o[5][4];       // o[4]: row sums (ones column)
  float mref[2];       // the exponent reference; only moved when the block maximum outgrows it by kTau
  uint32_t kb0, kb1;   // per-lane K fragment bases (ks = 0, 1), advanced 64 rows per block
  uint32_t vt0, vt1;   // per-lane V^T fragment bases (dt = 0, 2)
  uint32_t kc, kr;     // per-lane key code / region bases
};
constexpr float kTau = 8.0f;   // P <= e^8 between rescales: harmless in fp32 sums and bf16 P

// One 64-key block.  TAIL: the block holds padding keys and fewer than four 16-key pairs.
template <bool MASKED, bool TAIL>
__device__ __forceinline__ void fwd_block(const Lane& L, FwdRow& R, int keys_left, float scale) {
  const int npair = TAIL ? min(4, (keys_left + 15) >> 4) : 4;
  float s[8][4];
  static_for<4>([&](auto pr_) {
    constexpr int pr = decltype(pr_)::value;
    if (!TAIL || pr < npair) {
      uint32_t b[4];
      ldsm4_o<pr * 16 * ROWB>(b, R.kb0);
      mma_init(s[pr * 2], R.qf[0], b, 0.f, 0.f, 0.f, 0.f);
      mma_init(s[pr * 2 + 1], R.qf[0], b + 2, 0.f, 0.f, 0.f, 0.f);
      ldsm4_o<pr * 16 * ROWB>(b, R.kb1);
      mma_acc(s[pr * 2], R.qf[1], b);
      mma_acc(s[pr * 2 + 1], R.qf[1], b + 2);
    }
  });
  float mblk[2] = {-INFINITY, -INFINITY};
  static_for<8>([&](auto nt_) {
    constexpr int nt = decltype(nt_)::value;
    if (!TAIL || (nt >> 1) < npair) {
      const uint2 kc = ld_v2u32_o<nt * 32>(R.kc);
      uint2 kr = make_uint2(0u, 0u);
      if (MASKED) kr = ld_v2u32_o<nt * 32>(R.kr);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = e >> 1;
        float v = fmaf(s[nt][e], scale, ld_f32(R.qaddr[r] - ((e & 1) ? kc.y : kc.x)));
        if (MASKED && R.qrg[r] != ((e & 1) ? kr.y : kr.x)) v += kMask;
        if (TAIL && nt * 8 + L.t4 * 2 + (e & 1) >= keys_left) v = -INFINITY;
        s[nt][e] = v;
        mblk[r] = fmaxf(mblk[r], v);
      }
    }
  });
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    mblk[r] = fmaxf(mblk[r], __shfl_xor_sync(0xffffffffu, mblk[r], 1));
    mblk[r] = fmaxf(mblk[r], __shfl_xor_sync(0xffffffffu, mblk[r], 2));
  }
  // lazy rescale: the running reference only follows the maximum when it has grown by more than kTau
  if (__any_sync(0xffffffffu, mblk[0] > R.mref[0] + kTau || mblk[1] > R.mref[1] + kTau)) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float mnew = fmaxf(R.mref[r], mblk[r]);     // finite: block 0 always holds real keys
      const float corr = ex2((R.mref[r] - mnew) * kLog2e);
      R.mref[r] = mnew;
#pragma unroll
      for (int i = 0; i < 5; ++i) { R.o[i][r * 2] *= corr; R.o[i][r * 2 + 1] *= corr; }
    }
  }
  const float ml[2] = {-R.mref[0] * kLog2e, -R.mref[1] * kLog2e};
  const uint32_t ones[2] = {0x3f803f80u, 0x3f803f80u};   // bf16 1.0 pairs: column block of ones -> row sums of P
  static_for<4>([&](auto kk_) {
    constexpr int kk = decltype(kk_)::value;
    if (!TAIL || kk < npair) {
      uint32_t pf[4];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        float pv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) pv[e] = ex2(fmaf(s[kk * 2 + hf][e], kLog2e, ml[e >> 1]));
        pf[hf * 2 + 0] = pack2(pv[0], pv[1]);
        pf[hf * 2 + 1] = pack2(pv[2], pv[3]);
      }
      uint32_t b[4];
      ldsm4t_o<kk * 16 * ROWB>(b, R.vt0);
      mma_acc(R.o[0], pf, b);
      mma_acc(R.o[1], pf, b + 2);
      ldsm4t_o<kk * 16 * ROWB>(b, R.vt1);
      mma_acc(R.o[2], pf, b);
      mma_acc(R.o[3], pf, b + 2);
      mma_acc(R.o[4], pf, ones);
    }
  });
  R.kb0 += 64 * ROWB; R.kb1 += 64 * ROWB; R.vt0 += 64 * ROWB; R.vt1 += 64 * ROWB; R.kc += 256; R.kr += 256;
}

template <bool MASKED>
__device__ __forceinline__ void fwd_rows(const WinParams& P, const Smem& S, const Lane& L, int p, int h, int rb) {
  const int N = P.win.N;
  const uint32_t Qs = sm_addr(S.Qs), Ks = sm_addr(S.Ks), Vs = sm_addr(S.Vs);
  const uint32_t tab_s = sm_addr(S.tab);
  FwdRow R;
  ldsm4(R.qf[0], Qs + rb * 16 * ROWB + L.a_off[0]);
  ldsm4(R.qf[1], Qs + rb * 16 * ROWB + L.a_off[1]);
  const int i0 = rb * 16 + L.g;
  R.qaddr[0] = tab_s + S.qcode[i0]; R.qaddr[1] = tab_s + S.qcode[i0 + 8];
  R.qrg[0] = S.qreg[i0]; R.qrg[1] = S.qreg[i0 + 8];
#pragma unroll
  for (int i = 0; i < 5; ++i) R.o[i][0] = R.o[i][1] = R.o[i][2] = R.o[i][3] = 0.f;
  R.mref[0] = R.mref[1] = -INFINITY;
  R.kb0 = Ks + L.b_off[0]; R.kb1 = Ks + L.b_off[1];
  R.vt0 = Vs + L.bt_off[0]; R.vt1 = Vs + L.bt_off[1];
  R.kc = sm_addr(S.kcode) + L.t4 * 8; R.kr = sm_addr(S.qreg) + L.t4 * 8;
  const int nfull = N >> 6;
#pragma unroll 1
  for (int kb = 0; kb < nfull; ++kb) fwd_block<MASKED, false>(L, R, 64, P.scale);
  if (N & 63) fwd_block<MASKED, true>(L, R, N - nfull * 64, P.scale);
  const int col0 = h * HD;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int i = i0 + r * 8;
    if (i < N) {
      const float l = R.o[4][r * 2];   // every column of the ones block carries the row sum
      const float inv = 1.f / l;
      bf16* dst = P.O + (size_t)S.qrow[i] * P.ldo + col0;
#pragma unroll
      for (int dt = 0; dt < 4; ++dt)
        *(uint32_t*)(dst + dt * 8 + L.t4 * 2) = pack2(R.o[dt][r * 2] * inv, R.o[dt][r * 2 + 1] * inv);
      if (L.t4 == 0) P.lse[((size_t)p * P.heads + h) * N + i] = R.mref[r] + log2f(l) * kLn2;
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
  win_load_rows(S.Ks, P.qkv + C, P.ld, col0, S.qrow, P.NP);
  win_load_rows(S.Vs, P.qkv + 2 * C, P.ld, col0, S.qrow, P.NP);
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
// Running state of one 16-query block (dQ + bias gradient) while it sweeps the keys.
struct DqRow {
  uint32_t qf[2][4], dof[2][4];
  uint32_t qaddr[2], qrg[2];
  float nlse[2], ndel[2];
  float dq[4][4];
  uint32_t kb0, kb1, vb0, vb1;   // per-lane K / V fragment bases (ks = 0, 1)
  uint32_t kt0, kt1;             // per-lane K^T fragment bases (dt = 0, 2)
  uint32_t kc, kr;               // per-lane key code / region bases (fragment columns)
  uint32_t kcl;                  // code-word address of key (k0 + lane): the fold's column owner
  uint32_t my_qcode;             // qcode of row (lane & 15) of the block
  uint32_t gtab, stg_w, stg_r, dummy;   // private table, staging bases (fragment layout / lane layout), parking word
  __device__ __forceinline__ void advance(int keys) {
    kb0 += keys * ROWB; kb1 += keys * ROWB; vb0 += keys * ROWB; vb1 += keys * ROWB; kt0 += keys * ROWB; kt1 += keys * ROWB;
    kc += keys * 4; kr += keys * 4; kcl += keys * 4;
  }
};

// One 32-key sweep; SUB selects the first / second half of an unrolled pair (immediate offsets).
template <bool MASKED, bool TAIL, int SUB>
__device__ __forceinline__ void dq_sweep(const Lane& L, DqRow& R, int left, float scale, bool want_dtab) {
  constexpr int KO = SUB * 32 * ROWB;   // byte offset of this sweep inside the pair
  constexpr int CO = SUB * 128;         // code-word offset
  const int npair = TAIL ? min(2, (left + 15) >> 4) : 2;
  float s[4][4], dp[4][4];
  static_for<2>([&](auto pr_) {
    constexpr int pr = decltype(pr_)::value;
    if (!TAIL || pr < npair) {
      uint32_t b[4];
      ldsm4_o<KO + pr * 16 * ROWB>(b, R.kb0);
      mma_init(s[pr * 2], R.qf[0], b, 0.f, 0.f, 0.f, 0.f);
      mma_init(s[pr * 2 + 1], R.qf[0], b + 2, 0.f, 0.f, 0.f, 0.f);
      ldsm4_o<KO + pr * 16 * ROWB>(b, R.kb1);
      mma_acc(s[pr * 2], R.qf[1], b);
      mma_acc(s[pr * 2 + 1], R.qf[1], b + 2);
      ldsm4_o<KO + pr * 16 * ROWB>(b, R.vb0);   // dP - delta: the accumulator starts at -delta_i
      mma_init(dp[pr * 2], R.dof[0], b, R.ndel[0], R.ndel[0], R.ndel[1], R.ndel[1]);
      mma_init(dp[pr * 2 + 1], R.dof[0], b + 2, R.ndel[0], R.ndel[0], R.ndel[1], R.ndel[1]);
      ldsm4_o<KO + pr * 16 * ROWB>(b, R.vb1);
      mma_acc(dp[pr * 2], R.dof[1], b);
      mma_acc(dp[pr * 2 + 1], R.dof[1], b + 2);
    }
  });
  uint32_t dsf[2][4];
  static_for<4>([&](auto nt_) {
    constexpr int nt = decltype(nt_)::value;
    float ds[4] = {0.f, 0.f, 0.f, 0.f};
    if (!TAIL || (nt >> 1) < npair) {
      const uint2 kc = ld_v2u32_o<CO + nt * 32>(R.kc);
      uint2 kr = make_uint2(0u, 0u);
      if (MASKED) kr = ld_v2u32_o<CO + nt * 32>(R.kr);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = e >> 1;
        float v = fmaf(s[nt][e], scale, ld_f32(R.qaddr[r] - ((e & 1) ? kc.y : kc.x)));
        if (MASKED && R.qrg[r] != ((e & 1) ? kr.y : kr.x)) v += kMask;
        float pr_ = ex2(fmaf(v, kLog2e, R.nlse[r]));
        if (TAIL && nt * 8 + L.t4 * 2 + (e & 1) >= left) pr_ = 0.f;   // padding key
        ds[e] = pr_ * dp[nt][e];
      }
    }
    if (want_dtab) {   // stage the fp32 tile [16 rows][32 keys], row pitch 40 floats
      stv_v2f32_o<nt * 32>(R.stg_w, ds[0], ds[1]);
      stv_v2f32_o<8 * 160 + nt * 32>(R.stg_w, ds[2], ds[3]);
    }
    dsf[nt >> 1][(nt & 1) * 2 + 0] = pack2(ds[0], ds[1]);
    dsf[nt >> 1][(nt & 1) * 2 + 1] = pack2(ds[2], ds[3]);
  });
  static_for<2>([&](auto kk_) {
    constexpr int kk = decltype(kk_)::value;
    if (!TAIL || kk < npair) {
      uint32_t b[4];
      ldsm4t_o<KO + kk * 16 * ROWB>(b, R.kt0);
      mma_acc(R.dq[0], dsf[kk], b);
      mma_acc(R.dq[1], dsf[kk], b + 2);
      ldsm4t_o<KO + kk * 16 * ROWB>(b, R.kt1);
      mma_acc(R.dq[2], dsf[kk], b);
      mma_acc(R.dq[3], dsf[kk], b + 2);
    }
  });
  if (want_dtab) {
    // Fold the staged tile into the private table, one query row per step: lane l owns key k0 + l.  Within a row
    // distinct keys hit distinct slots; the warp is converged and its shared-memory instructions retire in
    // program order, so row r+1 observes row r's stores.  Padding keys (tail only) are parked on a dummy word.
    __syncwarp();
    const uint32_t base_l = R.gtab - ldv_u32_o<CO>(R.kcl);
    const bool live = !TAIL || L.lane < left;
    static_for<16>([&](auto r_) {
      constexpr int r = decltype(r_)::value;
      uint32_t slot = base_l + __shfl_sync(0xffffffffu, R.my_qcode, r);
      if (TAIL) slot = live ? slot : R.dummy;
      const float add = ldv_f32_o<r * 160>(R.stg_r);
      stv_f32(slot, ldv_f32(slot) + add);
    });
    __syncwarp();
  }
}

template <bool MASKED>
__device__ __forceinline__ void bwd_dq_unit(const WinParams& P, const Smem& S, const Lane& L, int h, int rb, uint32_t ln_s,
                                            uint32_t gtab_s, uint32_t stg_s) {
  const int N = P.win.N;
  const uint32_t Qs = sm_addr(S.Qs), Ks = sm_addr(S.Ks), Vs = sm_addr(S.Vs), dOs = sm_addr(S.dOs);
  const uint32_t tab_s = sm_addr(S.tab);
  DqRow R;
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    ldsm4(R.qf[ks], Qs + rb * 16 * ROWB + L.a_off[ks]);
    ldsm4(R.dof[ks], dOs + rb * 16 * ROWB + L.a_off[ks]);
  }
  const int i0 = rb * 16 + L.g;
  R.qaddr[0] = tab_s + S.qcode[i0]; R.qaddr[1] = tab_s + S.qcode[i0 + 8];
  R.qrg[0] = S.qreg[i0]; R.qrg[1] = S.qreg[i0 + 8];
  {
    const float2 a = ld_v2f32(ln_s + 8u * i0), b = ld_v2f32(ln_s + 8u * (i0 + 8));   // {lse, -delta}
    R.nlse[0] = -a.x * kLog2e; R.nlse[1] = -b.x * kLog2e;                              // -inf on padding rows -> p = 0
    R.ndel[0] = a.y; R.ndel[1] = b.y;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) R.dq[i][0] = R.dq[i][1] = R.dq[i][2] = R.dq[i][3] = 0.f;
  R.kb0 = Ks + L.b_off[0]; R.kb1 = Ks + L.b_off[1]; R.vb0 = Vs + L.b_off[0]; R.vb1 = Vs + L.b_off[1];
  R.kt0 = Ks + L.bt_off[0]; R.kt1 = Ks + L.bt_off[1];
  R.kc = sm_addr(S.kcode) + L.t4 * 8; R.kr = sm_addr(S.qreg) + L.t4 * 8;
  R.kcl = sm_addr(S.kcode) + L.lane * 4;
  R.my_qcode = S.qcode[rb * 16 + (L.lane & 15)];
  R.gtab = gtab_s;
  R.stg_w = stg_s + 4u * (uint32_t)(L.g * 40 + L.t4 * 2);
  R.stg_r = stg_s + 4u * (uint32_t)L.lane;
  R.dummy = stg_s + 16 * 40 * 4;
  const bool want_dtab = P.dtable != nullptr;
  const int nfull = N >> 5;
  int hb = 0;
#pragma unroll 1
  for (; hb + 2 <= nfull; hb += 2) {
    dq_sweep<MASKED, false, 0>(L, R, 32, P.scale, want_dtab);
    dq_sweep<MASKED, false, 1>(L, R, 32, P.scale, want_dtab);
    R.advance(64);
  }
  if (hb < nfull) {
    dq_sweep<MASKED, false, 0>(L, R, 32, P.scale, want_dtab);
    R.advance(32);
  }
  if (N & 31) dq_sweep<MASKED, true, 0>(L, R, N & 31, P.scale, want_dtab);
  const int col0 = h * HD;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int i = i0 + r * 8;
    if (i < N) {
      bf16* dst = P.dqkv + (size_t)S.qrow[i] * P.lddqkv + col0;
#pragma unroll
      for (int dt = 0; dt < 4; ++dt)
        *(uint32_t*)(dst + dt * 8 + L.t4 * 2) = pack2(R.dq[dt][r * 2] * P.scale, R.dq[dt][r * 2 + 1] * P.scale);
    }
  }
}

// dK, dV for one 16-key block: running state while it sweeps the queries.
struct DkvRow {
  uint32_t kf[2][4], vf[2][4];
  uint32_t kaddr[2], krg[2];
  float dk[4][4], dv[4][4];
  uint32_t qb0, qb1, dob0, dob1;   // per-lane Q / dO fragment bases (ks = 0, 1)
  uint32_t qt0, qt1, dot0, dot1;   // per-lane Q^T / dO^T fragment bases (dt = 0, 2)
  uint32_t qc, qr, ln;             // per-lane query code / region / {lse, -delta} bases (fragment columns)
  __device__ __forceinline__ void advance(int rows) {
    qb0 += rows * ROWB; qb1 += rows * ROWB; dob0 += rows * ROWB; dob1 += rows * ROWB;
    qt0 += rows * ROWB; qt1 += rows * ROWB; dot0 += rows * ROWB; dot1 += rows * ROWB;
    qc += rows * 4; qr += rows * 4; ln += rows * 8;
  }
};

template <bool MASKED, bool TAIL, int SUB>
__device__ __forceinline__ void dkv_sweep(const Lane& L, DkvRow& R, int left, float scale) {
  constexpr int QO = SUB * 32 * ROWB;
  constexpr int CO = SUB * 128;
  constexpr int LO = SUB * 256;
  const int npair = TAIL ? min(2, (left + 15) >> 4) : 2;
  float s[4][4], dp[4][4];   // rows = keys (g, g+8), cols = queries
  static_for<2>([&](auto pr_) {
    constexpr int pr = decltype(pr_)::value;
    if (!TAIL || pr < npair) {
      uint32_t b[4];
      const float4 l0 = ld_v4f32_o<LO + pr * 128>(R.ln);        // {lse, -delta} of the two query columns, n-tile 2*pr
      const float4 l1 = ld_v4f32_o<LO + pr * 128 + 64>(R.ln);   // n-tile 2*pr + 1
      ldsm4_o<QO + pr * 16 * ROWB>(b, R.qb0);
      mma_init(s[pr * 2], R.kf[0], b, 0.f, 0.f, 0.f, 0.f);
      mma_init(s[pr * 2 + 1], R.kf[0], b + 2, 0.f, 0.f, 0.f, 0.f);
      ldsm4_o<QO + pr * 16 * ROWB>(b, R.qb1);
      mma_acc(s[pr * 2], R.kf[1], b);
      mma_acc(s[pr * 2 + 1], R.kf[1], b + 2);
      ldsm4_o<QO + pr * 16 * ROWB>(b, R.dob0);
      mma_init(dp[pr * 2], R.vf[0], b, l0.y, l0.w, l0.y, l0.w);
      mma_init(dp[pr * 2 + 1], R.vf[0], b + 2, l1.y, l1.w, l1.y, l1.w);
      ldsm4_o<QO + pr * 16 * ROWB>(b, R.dob1);
      mma_acc(dp[pr * 2], R.vf[1], b);
      mma_acc(dp[pr * 2 + 1], R.vf[1], b + 2);
    }
  });
  uint32_t pf[2][4], dsf[2][4];
  static_for<4>([&](auto nt_) {
    constexpr int nt = decltype(nt_)::value;
    float pv[4] = {0.f, 0.f, 0.f, 0.f}, ds[4] = {0.f, 0.f, 0.f, 0.f};
    if (!TAIL || (nt >> 1) < npair) {
      const uint2 qc = ld_v2u32_o<CO + nt * 32>(R.qc);       // two consecutive queries (columns of S^T)
      const float4 l4 = ld_v4f32_o<LO + nt * 64>(R.ln);      // lse = +inf on padding queries -> p = 0
      uint2 qr = make_uint2(0u, 0u);
      if (MASKED) qr = ld_v2u32_o<CO + nt * 32>(R.qr);
      const float nl[2] = {-l4.x * kLog2e, -l4.z * kLog2e};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = e >> 1;
        float v = fmaf(s[nt][e], scale, ld_f32(R.kaddr[r] + ((e & 1) ? qc.y : qc.x)));
        if (MASKED && R.krg[r] != ((e & 1) ? qr.y : qr.x)) v += kMask;
        const float pr_ = ex2(fmaf(v, kLog2e, nl[e & 1]));
        pv[e] = pr_;
        ds[e] = pr_ * dp[nt][e];
      }
    }
    pf[nt >> 1][(nt & 1) * 2 + 0] = pack2(pv[0], pv[1]);
    pf[nt >> 1][(nt & 1) * 2 + 1] = pack2(pv[2], pv[3]);
    dsf[nt >> 1][(nt & 1) * 2 + 0] = pack2(ds[0], ds[1]);
    dsf[nt >> 1][(nt & 1) * 2 + 1] = pack2(ds[2], ds[3]);
  });
  static_for<2>([&](auto kk_) {   // contraction over the 16-query pairs of this sweep
    constexpr int kk = decltype(kk_)::value;
    if (!TAIL || kk < npair) {
      uint32_t b[4];
      ldsm4t_o<QO + kk * 16 * ROWB>(b, R.dot0);
      mma_acc(R.dv[0], pf[kk], b);
      mma_acc(R.dv[1], pf[kk], b + 2);
      ldsm4t_o<QO + kk * 16 * ROWB>(b, R.dot1);
      mma_acc(R.dv[2], pf[kk], b);
      mma_acc(R.dv[3], pf[kk], b + 2);
      ldsm4t_o<QO + kk * 16 * ROWB>(b, R.qt0);
      mma_acc(R.dk[0], dsf[kk], b);
      mma_acc(R.dk[1], dsf[kk], b + 2);
      ldsm4t_o<QO + kk * 16 * ROWB>(b, R.qt1);
      mma_acc(R.dk[2], dsf[kk], b);
      mma_acc(R.dk[3], dsf[kk], b + 2);
    }
  });
}

template <bool MASKED>
__device__ __forceinline__ void bwd_dkv_unit(const WinParams& P, const Smem& S, const Lane& L, int h, int jb, uint32_t ln_s) {
  const int N = P.win.N;
  const uint32_t Qs = sm_addr(S.Qs), Ks = sm_addr(S.Ks), Vs = sm_addr(S.Vs), dOs = sm_addr(S.dOs);
  const uint32_t tab_s = sm_addr(S.tab);
  DkvRow R;
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    ldsm4(R.kf[ks], Ks + jb * 16 * ROWB + L.a_off[ks]);
    ldsm4(R.vf[ks], Vs + jb * 16 * ROWB + L.a_off[ks]);
  }
  const int j0 = jb * 16 + L.g;   // keys j0, j0+8 (padding keys: K/V rows are zero, results dropped)
  R.kaddr[0] = tab_s - S.kcode[j0]; R.kaddr[1] = tab_s - S.kcode[j0 + 8];
  R.krg[0] = S.qreg[j0]; R.krg[1] = S.qreg[j0 + 8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) R.dk[i][e] = R.dv[i][e] = 0.f;
  R.qb0 = Qs + L.b_off[0]; R.qb1 = Qs + L.b_off[1]; R.dob0 = dOs + L.b_off[0]; R.dob1 = dOs + L.b_off[1];
  R.qt0 = Qs + L.bt_off[0]; R.qt1 = Qs + L.bt_off[1]; R.dot0 = dOs + L.bt_off[0]; R.dot1 = dOs + L.bt_off[1];
  R.qc = sm_addr(S.qcode) + L.t4 * 8; R.qr = sm_addr(S.qreg) + L.t4 * 8; R.ln = ln_s + L.t4 * 16;
  const int nfull = N >> 5;
  int hb = 0;
#pragma unroll 1
  for (; hb + 2 <= nfull; hb += 2) {
    dkv_sweep<MASKED, false, 0>(L, R, 32, P.scale);
    dkv_sweep<MASKED, false, 1>(L, R, 32, P.scale);
    R.advance(64);
  }
  if (hb < nfull) {
    dkv_sweep<MASKED, false, 0>(L, R, 32, P.scale);
    R.advance(32);
  }
  if (N & 31) dkv_sweep<MASKED, true, 0>(L, R, N & 31, P.scale);
  const int C = P.heads * HD, col0 = h * HD;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int j = j0 + r * 8;
    if (j < N) {
      bf16* dstk = P.dqkv + (size_t)S.qrow[j] * P.lddqkv + C + col0;
      bf16* dstv = dstk + C;
#pragma unroll
      for (int dt = 0; dt < 4; ++dt) {
        *(uint32_t*)(dstk + dt * 8 + L.t4 * 2) = pack2(R.dk[dt][r * 2] * P.scale, R.dk[dt][r * 2 + 1] * P.scale);
        *(uint32_t*)(dstv + dt * 8 + L.t4 * 2) = pack2(R.dv[dt][r * 2], R.dv[dt][r * 2 + 1]);
      }
    }
  }
}

__global__ void __launch_bounds__(448, 1)
window_bwd_kernel(WinParams P) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int NP = P.NP;
  const Smem S = carve(smem, NP, true);
  float* ln_sm = (float*)((unsigned char*)S.tab + P.tab_stride);    // per query {natural-log lse (+inf on padding), -delta}
  int* wrange = (int*)(ln_sm + 2 * NP);                              // [16][2] first / last slot byte offset of each dq warp's table
  unsigned char* priv = (unsigned char*)(wrange + 32);               // per dq warp: [gtab_bytes] table + [16][40] staging + parking word
  const int priv_stride = P.gtab_bytes + 16 * 40 * 4 + 16;
  const int p = blockIdx.x, h = blockIdx.y;
  const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int C = P.heads * HD, col0 = h * HD;
  const int N = P.win.N;
  const bool masked = win_build_tables(P, p, h, S);
  __syncthreads();
  win_load_rows(S.Qs, P.qkv, P.ld, col0, S.qrow, NP);
  win_load_rows(S.Ks, P.qkv + C, P.ld, col0, S.qrow, NP);
  win_load_rows(S.Vs, P.qkv + 2 * C, P.ld, col0, S.qrow, NP);
  win_load_rows(S.dOs, P.dO, P.ldo, col0, S.qrow, NP);
  cp_commit();
  for (int i = threadIdx.x; i < NP; i += blockDim.x) ln_sm[2 * i] = i < N ? P.lse[((size_t)p * P.heads + h) * N + i] : INFINITY;
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
    if (ch == 0) ln_sm[2 * r + 1] = -d;
  }
  __syncthreads();
  Lane L;
  L.init();
  const int nrb = NP >> 4;
  const uint32_t ln_s = sm_addr(ln_sm);
  const int nq = P.n_dq_warps;
  if (warp < nq) {
    int start, cnt;
    dq_range(warp, nq, nrb, start, cnt);
    // slot byte offsets this warp can touch: [qcode(first row) - 4*maxcode, qcode(last row)] (codes grow with the index)
    const int lo = (int)S.qcode[start * 16] - 4 * P.maxcode;
    const int hi = (int)S.qcode[min((start + cnt) * 16, N) - 1];
    if (L.lane == 0) { wrange[2 * warp] = lo; wrange[2 * warp + 1] = hi; }
    const uint32_t phys = sm_addr(priv + warp * priv_stride);
    const uint32_t gtab_s = phys - (uint32_t)lo;   // virtual base: slot byte offset b lives at phys + (b - lo)
    const uint32_t stg_s = phys + P.gtab_bytes;
    if (masked) { for (int rb = start; rb < start + cnt; ++rb) bwd_dq_unit<true>(P, S, L, h, rb, ln_s, gtab_s, stg_s); }
    else        { for (int rb = start; rb < start + cnt; ++rb) bwd_dq_unit<false>(P, S, L, h, rb, ln_s, gtab_s, stg_s); }
  } else {
    const int nk = nwarps - nq;
    if (masked) { for (int jb = warp - nq; jb < nrb; jb += nk) bwd_dkv_unit<true>(P, S, L, h, jb, ln_s); }
    else        { for (int jb = warp - nq; jb < nrb; jb += nk) bwd_dkv_unit<false>(P, S, L, h, jb, ln_s); }
  }
  if (P.dtable != nullptr) {
    __syncthreads();
    const int r0 = P.center - P.maxcode;
    for (int r = threadIdx.x; r < P.n_used; r += blockDim.x) {
      float acc = 0.f;
      for (int w = 0; w < nq; ++w) {
        const int lo = wrange[2 * w], hi = wrange[2 * w + 1];
        if (4 * r >= lo && 4 * r <= hi) acc += *(const float*)(priv + w * priv_stride + (4 * r - lo));
      }
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

// bytes of the largest private gradient table when n_dq warps share the query blocks contiguously
static int gtab_bytes_for(const WindowIndex& ix, int NP, int maxcode, int n_dq) {
  const int cW = 2 * ix.WW - 1, cH = (2 * ix.WH - 1) * cW;
  auto code = [&](int i) { return (i / (ix.wh * ix.ww)) * cH + ((i / ix.ww) % ix.wh) * cW + i % ix.ww; };
  const int nrb = NP / 16;
  int worst = 0;
  for (int w = 0; w < n_dq; ++w) {
    int start, cnt;
    dq_range(w, n_dq, nrb, start, cnt);
    if (cnt == 0) continue;
    const int last = std::min((start + cnt) * 16, ix.N) - 1;
    worst = std::max(worst, 4 * (code(last) - code(start * 16)) + 4 * maxcode + 4);
  }
  return (worst + 15) / 16 * 16;
}

// backward: split up to 14 warps between query blocks (dQ + bias gradient, ~1.7x the work of a key block) and key
// blocks (dK, dV) so that the slower group finishes earliest; every dq warp needs a private table in shared memory
static bool bwd_warps(const WindowIndex& ix, int NP, int n_used, int maxcode, int& n_dq, int& n_dkv, int& gtab_bytes) {
  const int nrb = NP / 16;
  double best = 1e30;
  n_dq = n_dkv = gtab_bytes = 0;
  for (int q = 1; q <= 13 && q <= nrb; ++q) {
    const int gb = gtab_bytes_for(ix, NP, maxcode, q);
    if (win_smem_bytes(NP, n_used, true, q, gb) > kSmemLimit) continue;
    for (int k = 1; q + k <= 14; ++k) {
      const double cost = std::max(1.7 * ((nrb + q - 1) / q), 1.0 * ((nrb + k - 1) / k)) + 1e-3 * (q + k);
      if (cost < best) { best = cost; n_dq = q; n_dkv = k; gtab_bytes = gb; }
    }
  }
  return n_dq > 0;
}

bool window_cta_eligible(const WindowIndex& ix, int hd) {
  if (hd != HD || ix.N > 4095) return false;
  WinParams P = {};
  win_geometry(P, ix, 1);
  if (4 * (2 * P.maxcode + 1) >= 65536) return false;
  if (ix.wh * ix.ww > 256) return false;   // small_div range
  int q, k, gb;
  return win_smem_bytes(P.NP, P.n_used, false, 0) <= kSmemLimit && bwd_warps(ix, P.NP, P.n_used, P.maxcode, q, k, gb);
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
#ifdef VALOR_DEBUG   // -DVALOR_DEBUG only: VALOR_WINDOW_NO_DTAB=1 skips the bias-table gradient (cost measurement of the fold)
  static int no_dtab = -1;
  if (no_dtab < 0) { const char* e = getenv("VALOR_WINDOW_NO_DTAB"); no_dtab = e ? atoi(e) : 0; if (no_dtab) fprintf(stderr, "valor_b200: VALOR_WINDOW_NO_DTAB active\n"); }
  if (no_dtab) P.dtable = nullptr;
#endif
  int n_dq = 0, n_dkv = 0;
  int gb = 0;
  VALOR_REQUIRE(bwd_warps(ix, P.NP, P.n_used, P.maxcode, n_dq, n_dkv, gb), "window_cta_bwd: window does not fit in shared memory");
  P.n_dq_warps = n_dq;
  P.gtab_bytes = gb;
  const size_t smem = win_smem_bytes(P.NP, P.n_used, true, n_dq, gb);
  static size_t attr = 0;
  if (smem > attr) { VALOR_CUDA(cudaFuncSetAttribute(window_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = smem; }
  window_bwd_kernel<<<dim3(Pn, H), (n_dq + n_dkv) * 32, smem, st>>>(P);
  return check_launch("window_bwd_kernel");
}

}  // namespace valor
