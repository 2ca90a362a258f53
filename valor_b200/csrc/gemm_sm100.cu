// valor_b200 — tcgen05 / TMEM / TMA GEMM with fused epilogues (sm_100a only).
//
//   C[M,N] (+)= epilogue( alpha * A[M,K] . B[N,K]^T )
//
// One kernel serves the three GEMM forms of every Linear on VALOR's hot path
// (reference: nn.Linear / torch.matmul call sites):
//   forward : A = x   [M,K]  K-major,   B = W  [N,K]  K-major
//   dgrad   : A = dy  [M,N'] K-major,   B = W  stored [N',K'] -> MN-major (contract over N')
//   wgrad   : A = dy  stored [M',N] -> MN-major, B = x stored [M',K] -> MN-major (contract M')
//
// Design: persistent, warp-specialised. warp0 = TMA producer (cp.async.bulk.tensor, 128B
// swizzle), warp1 = single-thread tcgen05.mma issuer (UMMA 128 x BLOCK_N x 16, bf16 in,
// fp32 accumulate in TMEM, 2 accumulator stages), warp2 = TMEM allocator, warps4-11 =
// epilogue: pipelined tcgen05.ld 32x32b -> bias / erf-GELU (+ pre-activation side output) /
// GELU'(aux) / residual (aux and residual tiles arrive by TMA) -> per-warp swizzled staging ->
// TMA store; fp32 split-K weight gradients leave through TMA reduce-add
// (cp.reduce.async.bulk.tensor .add).  Unaligned / fp32 outputs fall back to direct stores.
#include "common.cuh"
#include <cudaTypedefs.h>
#include <stdlib.h>

namespace valor {

static constexpr int BLOCK_M = 128;
static constexpr int BLOCK_K = 64;  // 64 bf16 = one 128-byte swizzle row
static constexpr int UMMA_K = 16;

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  uint32_t addr = smem_u32(bar);
  while (!ok) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* tm, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// ---- two-CTA (cta_group::2) forms: the pair's even CTA ("leader") owns the barriers the MMA thread waits on; a
// shared::cluster address with bit 24 cleared names the leader's copy of the same shared-memory offset
// (cute/arch/copy_sm100_tma.hpp Sm100MmaPeerBitMask)
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* tm, uint64_t* leader_bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)tm), "r"(smem_u32(leader_bar) & kPeerMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerMask) : "memory");
}
__device__ __forceinline__ void tcgen05_commit_pair(uint64_t* bar) {   // arrives on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
      " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor): 128B swizzle.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= 1ull << 46;  // version = 1 (Blackwell)
  d |= 2ull << 61;  // SWIZZLE_128B
  return d;
}

// Epilogue specialisations of the TMA-store path (compile-time: no per-element branching).
//   PLAIN    : bias
//   RES      : bias + residual
//   GELU_PRE : bias, pre-activation side output, erf-GELU
//   GELU_AUX : * GELU'(aux)   (dgrad through the activation)
//   GENERIC  : everything decided at run time (other activations, odd combinations)
//   ACC      : fp32 accumulate into global memory by TMA reduce-add (split-K weight gradients)
enum EpiMode { EPI_GENERIC = 0, EPI_PLAIN = 1, EPI_RES = 2, EPI_GELU_PRE = 3, EPI_GELU_AUX = 4, EPI_ACC = 5 };

// One 32-column slab of one accumulator row: v (raw TMEM words) -> staging row(s) in shared memory.
template <int MODE>
__device__ __forceinline__ void epi_slab(const uint32_t* v, const float* wb, const GemmEpilogue& ep, int row, int col0, int M,
                                         int N, int lane, int sub, uint8_t* st_out, uint8_t* st_pre,
                                         const uint8_t* in_tile, float rsc) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float xg[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) xg[j] = fmaf(__uint_as_float(v[g * 8 + j]), ep.alpha, wb[g * 8 + j]);
    if ((MODE == EPI_RES || MODE == EPI_GENERIC) && ep.row_scale != nullptr) {   // DropPath factor of this row (uniform branch)
#pragma unroll
      for (int j = 0; j < 8; ++j) xg[j] *= rsc;
    }
    const int col = col0 + g * 8;
    const int chunk = sub * 4 + g;  // 16-byte chunk inside the 128-byte staging row
    const uint32_t soff = (uint32_t)lane * 128u + (uint32_t)((chunk ^ (lane & 7)) << 4);
    const bool inb = row < M && col + 8 <= N;
    const bool want_pre = (MODE == EPI_GELU_PRE) || (MODE == EPI_GENERIC && ep.preact_out != nullptr);
    if (want_pre) {
      uint4 pk;
      __nv_bfloat162* h = (__nv_bfloat162*)&pk;
#pragma unroll
      for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(xg[2 * j], xg[2 * j + 1]);
      *(uint4*)(st_pre + soff) = pk;
    }
    const bool use_aux = (MODE == EPI_GELU_AUX) || (MODE == EPI_GENERIC && ep.act_aux != nullptr);
    if (MODE == EPI_GELU_AUX) {   // aux tile staged in shared memory by TMA (same 128B swizzle as the output)
      const uint4 a = *(const uint4*)(in_tile + soff);
      const __nv_bfloat162* h = (const __nv_bfloat162*)&a;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 f = __bfloat1622float2(h[j]);
        xg[2 * j] *= gelu_grad_fast(f.x);
        xg[2 * j + 1] *= gelu_grad_fast(f.y);
      }
    } else if (use_aux) {
      const int act = ep.act;
      if (inb) {
        uint4 a = *(const uint4*)((const bf16*)ep.act_aux + (size_t)row * ep.ld_aux + col);
        const __nv_bfloat162* h = (const __nv_bfloat162*)&a;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 f = __bfloat1622float2(h[j]);
          xg[2 * j] *= act_grad(f.x, act);
          xg[2 * j + 1] *= act_grad(f.y, act);
        }
      } else if (row < M) {
        for (int j = 0; j < 8 && col + j < N; ++j)
          xg[j] *= act_grad(__bfloat162float(((const bf16*)ep.act_aux)[(size_t)row * ep.ld_aux + col + j]), act);
      }
    } else if (MODE == EPI_GELU_PRE) {
#pragma unroll
      for (int j = 0; j < 8; ++j) xg[j] = gelu_fwd_fast(xg[j]);
    } else if (MODE == EPI_GENERIC) {
      if (ep.act != VALOR_ACT_NONE) {
#pragma unroll
        for (int j = 0; j < 8; ++j) xg[j] = act_fwd(xg[j], ep.act);
      }
    }
    const bool use_res = (MODE == EPI_GENERIC && ep.residual != nullptr);
    if (MODE == EPI_RES) {        // residual tile staged in shared memory by TMA
      const uint4 r = *(const uint4*)(in_tile + soff);
      const __nv_bfloat162* h = (const __nv_bfloat162*)&r;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 f = __bfloat1622float2(h[j]);
        xg[2 * j] += f.x;
        xg[2 * j + 1] += f.y;
      }
    } else if (use_res) {
      if (inb) {
        uint4 r = *(const uint4*)((const bf16*)ep.residual + (size_t)row * ep.ldr + col);
        const __nv_bfloat162* h = (const __nv_bfloat162*)&r;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 f = __bfloat1622float2(h[j]);
          xg[2 * j] += f.x;
          xg[2 * j + 1] += f.y;
        }
      } else if (row < M) {
        for (int j = 0; j < 8 && col + j < N; ++j)
          xg[j] += __bfloat162float(((const bf16*)ep.residual)[(size_t)row * ep.ldr + col + j]);
      }
    }
    uint4 pk;
    __nv_bfloat162* h = (__nv_bfloat162*)&pk;
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(xg[2 * j], xg[2 * j + 1]);
    *(uint4*)(st_out + soff) = pk;
  }
}

// BG: the weight-gradient launch also produces the bias gradient (row sums of A over the contraction) through one extra
// N = 16 MMA per k-step against a constant all-ones B tile: 2 KB of shared memory, 16 more accumulator columns per stage.
// PAIR: two CTAs of a cluster (an SM pair) share one 256 x BLOCK_N tile: each stages its own 128 rows of A and HALF of
// the B tile, the leader issues tcgen05.mma.cta_group::2 (UMMA 256 x BLOCK_N x 16) and every CTA ends up with its 128
// accumulator rows in its own tensor memory.  Per SM and k-block that is 32 KB of operands instead of 48 KB: the
// 128 x 256 single-CTA tile is bound by the L2 -> shared-memory path, not by the tensor pipe.
template <int BLOCK_N, bool BG = false, bool PAIR = false>
struct GemmCfg {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = (PAIR ? BLOCK_N / 2 : BLOCK_N) * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGING_BYTES = 4 * 16384;  // per epilogue half: 128x64 bf16 out tile + pre-activation tile
  static constexpr int BIAS_BYTES = 4096;
  static constexpr int ONES_BYTES = BG ? 768 + 2048 : 0;   // pad to 1 KB + [16 rows][64 k] of bf16 1.0 (K-major B operand)
  static constexpr int BUDGET = 227 * 1024 - 1024 /*align*/ - 256 /*barriers*/ - BIAS_BYTES - STAGING_BYTES - ONES_BYTES;
  static constexpr int STAGES = BUDGET / STAGE_BYTES > 8 ? 8 : BUDGET / STAGE_BYTES;
  // accumulator stages in tensor memory: two (the epilogue of tile i overlaps the main loop of tile i + 1), except the
  // 256-wide tiles that also carry the bias gradient: 2 x 256 + 32 columns do not exist, and a split-K weight gradient
  // spends ~40 k-blocks per work item in the main loop against one short reduce-add epilogue, so one stage costs ~4 %
  static constexpr int ACC_STAGES = (BG && BLOCK_N > 192) ? 1 : 2;
  static constexpr int ACC_COLS = ACC_STAGES * (BLOCK_N + (BG ? 16 : 0));
  static constexpr int TMEM_COLS = (ACC_COLS <= 128) ? 128 : (ACC_COLS <= 256 ? 256 : 512);
  static_assert(ACC_COLS <= 512, "accumulators do not fit tensor memory");
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 1024 + 256 + BIAS_BYTES + ONES_BYTES;
};

template <int BLOCK_N, bool A_KMAJOR, bool B_KMAJOR, int MODE, bool BG = false, bool PAIR = false>
__global__ void __launch_bounds__(384, 1)
gemm_sm100_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmP,
                  void* __restrict__ Cptr, long long ldc, int M, int N, int K, int k_splits, int vec_ok,
                  int tma_store, int dbg, GemmEpilogue ep) {
  using Cfg = GemmCfg<BLOCK_N, BG, PAIR>;
  constexpr int STAGES = Cfg::STAGES;
  static_assert(!PAIR || BLOCK_N % 128 == 0, "two-CTA tiles: BLOCK_N / 2 must be a whole number of 64-wide panels");
  constexpr int ACC_STAGES = Cfg::ACC_STAGES;
  constexpr int TILE_M = PAIR ? 2 * BLOCK_M : BLOCK_M;       // rows of one work item
  const int rank = PAIR ? (int)cluster_ctarank() : 0;        // 0 = leader (issues the MMAs)
  const int unit = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int nunits = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  static_assert(!BG || (MODE == EPI_ACC && !A_KMAJOR && !B_KMAJOR), "bias-gradient fusion belongs to the weight-gradient form");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
  uint8_t* staging = smem + STAGES * Cfg::STAGE_BYTES;  // 1024-aligned (stage sizes are multiples of 8 KB)
  uint64_t* bars = (uint64_t*)(staging + Cfg::STAGING_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;
  uint32_t* tmem_ptr_smem = (uint32_t*)(bars + 2 * STAGES + 4);
  float* bias_s = (float*)(staging + Cfg::STAGING_BYTES + 256);
  uint8_t* ones_s = staging + Cfg::STAGING_BYTES + 256 + Cfg::BIAS_BYTES + 768;   // (BG) 1024-aligned

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_blocks = (M + TILE_M - 1) / TILE_M;
  const int n_blocks = (N + BLOCK_N - 1) / BLOCK_N;
  const int kb_total = (K + BLOCK_K - 1) / BLOCK_K;
  const int kb_per = (kb_total + k_splits - 1) / k_splits;
  const int total_work = m_blocks * n_blocks * k_splits;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmB) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], PAIR ? 16 : 8);   // PAIR: the epilogue warps of both CTAs release the leader's stage
    }
    for (int i = 0; i < 8; ++i) mbar_init(&bars[2 * STAGES + 6 + i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (BG) {   // all-ones B tile (every element equal: swizzle / layout cannot matter)
    for (int i = threadIdx.x; i < 2048 / 16; i += blockDim.x) ((uint4*)ones_s)[i] = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    if (PAIR) {   // same warp, same destination offset in both CTAs (cute/arch/tmem_allocator_sm100.hpp Allocator2Sm)
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                   "r"((uint32_t)Cfg::TMEM_COLS));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                   "r"((uint32_t)Cfg::TMEM_COLS));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
  }
  tcgen05_fence_before();
  if (PAIR) cluster_sync_all();   // both CTAs' barriers exist before any remote arrive / multicast commit
  else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = unit; w < total_work; w += nunits) {
        const int n_blk = w % n_blocks;
        const int rest = w / n_blocks;
        const int m_blk = rest % m_blocks;
        const int ks = rest / m_blocks;
        const int kb0 = ks * kb_per;
        const int kb1 = min(kb0 + kb_per, kb_total);
        const int row0 = m_blk * TILE_M + rank * BLOCK_M;                       // this CTA's 128 rows of A
        const int col0b = n_blk * BLOCK_N + (PAIR ? rank * (BLOCK_N / 2) : 0);   // this CTA's share of the B tile
        constexpr int B_ROWS = PAIR ? BLOCK_N / 2 : BLOCK_N;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem_a + stage * Cfg::A_BYTES;
          uint8_t* sb = smem_b + stage * Cfg::B_BYTES;
          if (PAIR) {
            // both CTAs' loads complete on the LEADER's full barrier, which expects the bytes of the whole pair
            if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
            if (A_KMAJOR) {
              tma_load_2d_pair(&tmA, &full_bar[stage], sa, kb * BLOCK_K, row0);
            } else {
#pragma unroll
              for (int j = 0; j < BLOCK_M / 64; ++j)
                tma_load_2d_pair(&tmA, &full_bar[stage], sa + j * (BLOCK_K * 128), row0 + j * 64, kb * BLOCK_K);
            }
            if (B_KMAJOR) {
              tma_load_2d_pair(&tmB, &full_bar[stage], sb, kb * BLOCK_K, col0b);
            } else {
#pragma unroll
              for (int j = 0; j < B_ROWS / 64; ++j)
                tma_load_2d_pair(&tmB, &full_bar[stage], sb + j * (BLOCK_K * 128), col0b + j * 64, kb * BLOCK_K);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          if (A_KMAJOR) {
            tma_load_2d(&tmA, &full_bar[stage], sa, kb * BLOCK_K, m_blk * BLOCK_M);
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_M / 64; ++j)
              tma_load_2d(&tmA, &full_bar[stage], sa + j * (BLOCK_K * 128), m_blk * BLOCK_M + j * 64, kb * BLOCK_K);
          }
          if (B_KMAJOR) {
            tma_load_2d(&tmB, &full_bar[stage], sb, kb * BLOCK_K, n_blk * BLOCK_N);
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_N / 64; ++j)
              tma_load_2d(&tmB, &full_bar[stage], sb + j * (BLOCK_K * 128), n_blk * BLOCK_N + j * 64, kb * BLOCK_K);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && rank == 0) {
      // instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor); PAIR: UMMA_M = 256 over the two CTAs
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_KMAJOR ? 0u : 1u) << 15) |
                             ((B_KMAJOR ? 0u : 1u) << 16) | ((uint32_t)(BLOCK_N >> 3) << 17) |
                             ((uint32_t)(TILE_M >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      // bias-gradient MMA: A as above (MN-major), B = sixteen K-major rows of ones
      const uint32_t idesc_bg = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | ((uint32_t)(16 >> 3) << 17) |
                                ((uint32_t)(TILE_M >> 4) << 24);
      for (int w = unit; w < total_work; w += nunits) {
        const int rest = w / n_blocks;
        const int ks = rest / m_blocks;
        const int kb0 = ks * kb_per;
        const int kb1 = min(kb0 + kb_per, kb_total);
        const bool bg_tile = BG && (w % n_blocks) == 0;     // one column of tiles owns the row sums of A
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem_a + stage * Cfg::A_BYTES);
          const uint32_t sb = smem_u32(smem_b + stage * Cfg::B_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // K-major: advance 32 B inside the 128B swizzle row; SBO = 8 rows * 128 B.
            // MN-major: 16 k-rows * 128 B; SBO = 1024 (8 k-rows), LBO = BLOCK_K*128 (next 64-wide MN atom).
            const uint64_t adesc = A_KMAJOR ? make_smem_desc(sa + k * UMMA_K * 2, 0, 1024)
                                            : make_smem_desc(sa + k * UMMA_K * 128, BLOCK_K * 128, 1024);
            const uint64_t bdesc = B_KMAJOR ? make_smem_desc(sb + k * UMMA_K * 2, 0, 1024)
                                            : make_smem_desc(sb + k * UMMA_K * 128, BLOCK_K * 128, 1024);
            if (PAIR) tcgen05_mma_bf16_pair(tmem_d, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else tcgen05_mma_bf16(tmem_d, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            if (BG && bg_tile)   // bias gradient: the same A tile against sixteen columns of ones
            {
              const uint32_t d_bg = tmem_base + ACC_STAGES * BLOCK_N + acc * 16;
              const uint64_t ones_d = make_smem_desc(smem_u32(ones_s) + k * UMMA_K * 2, 0, 1024);
              if (PAIR) tcgen05_mma_bf16_pair(d_bg, adesc, ones_d, idesc_bg, (kb > kb0 || k > 0) ? 1u : 0u);
              else tcgen05_mma_bf16(d_bg, adesc, ones_d, idesc_bg, (kb > kb0 || k > 0) ? 1u : 0u);
            }
          }
          if (PAIR) {                          // both CTAs' producers / epilogues learn about it
            tcgen05_commit_pair(&empty_bar[stage]);
            if (kb == kb1 - 1) tcgen05_commit_pair(&tmem_full[acc]);
          } else {
            tcgen05_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
            if (kb == kb1 - 1) tcgen05_commit(&tmem_full[acc]);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (8 warps) =====================
    // warp e: TMEM lane quarter q = e % 4 (hardware restriction: warp_id % 4), column half = e / 4.
    const int e = warp - 4;
    const int q = e & 3;
    const int half = e >> 2;
    const int et = threadIdx.x - 128;           // 0..255
    const int r_tile = q * 32 + lane;           // row inside the 128-row tile
    // TMA-store path: every warp owns a private 2 x 4 KB staging slab (its 32 rows x 64 columns), a private
    // bias slice and issues its own bulk stores: no CTA-level barrier anywhere in the epilogue.
    uint8_t* wst = staging + e * 8192;
    float* wbias = bias_s + e * 128;
    // RES / GELU_AUX: the residual / saved pre-activation tile of each chunk arrives by TMA into the warp's
    // second staging buffer (coalesced 128-byte rows instead of 32 row-strided 16-byte loads per instruction)
    constexpr bool kInTile = (MODE == EPI_RES) || (MODE == EPI_GELU_AUX);
    uint64_t* in_bar = bars + 2 * STAGES + 6 + e;
    uint32_t in_phase = 0;
    const bool has_pre = ep.preact_out != nullptr;
    uint32_t chunk_ctr = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int w = unit; w < total_work; w += nunits) {
      const int n_blk = w % n_blocks;
      const int rest = w / n_blocks;
      const int m_blk = rest % m_blocks;
      const int ks = rest / m_blocks;
      const int n0 = n_blk * BLOCK_N;
      const int row_base = m_blk * TILE_M + rank * BLOCK_M;   // first row of this CTA's 128 accumulator rows
      const int row = row_base + r_tile;
      const float rsc = ep.row_scale != nullptr ? ep.row_scale[min(row, M - 1) / ep.rows_per_group] : 1.0f;
      if (!tma_store) {
        // stage the bias slice (only split 0 adds bias when split-K accumulates)
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int i = et; i < BLOCK_N; i += 256) {
          float b = 0.f;
          if (ep.bias != nullptr && ks == 0 && n0 + i < N) b = ep.bias[n0 + i];
          bias_s[i] = b;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      if (!tma_store || (dbg & 1)) {
        mbar_wait(&tmem_full[acc], acc_phase);
        tcgen05_fence_after();
      }
      if (dbg & 1) {
        // (diagnostic) mainloop only: release the accumulator without reading it
      } else if (tma_store) {
        if (MODE == EPI_ACC) {
          // ---------- fp32 tile += accumulator: 32x32 fp32 slabs staged in swizzled smem, folded into global
          // memory by the TMA unit (cp.reduce.async.bulk .add): coalesced 128-byte reductions at L2, no
          // per-thread atomics, OOB rows/columns clipped by the tensor map ----------
          mbar_wait(&tmem_full[acc], acc_phase);
          tcgen05_fence_after();
          const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BLOCK_N;
          if (BG && n_blk == 0 && half == 0) {   // row sums of A (bias gradient): every one of the 16 columns carries the sum
            uint32_t vb8[8];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(vb8[0]), "=r"(vb8[1]), "=r"(vb8[2]), "=r"(vb8[3]), "=r"(vb8[4]), "=r"(vb8[5]), "=r"(vb8[6]), "=r"(vb8[7])
                         : "r"(tmem_base + ((uint32_t)(q * 32) << 16) + ACC_STAGES * BLOCK_N + acc * 16));
            tmem_ld_wait();
            if (row < M) atomicAdd(ep.bias_grad + row, __uint_as_float(vb8[0]) * ep.alpha);
          }
          uint32_t va[32];
          if (n0 + half * 32 < N) tmem_ld_32x32(trow + half * 32, va);
#pragma unroll 1
          for (int c = half; c < BLOCK_N / 32; c += 2) {
            if (n0 + c * 32 >= N) break;
            uint8_t* st_out = wst + ((chunk_ctr & 1) ? 4096 : 0);
            ++chunk_ctr;
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            tmem_ld_wait();
            __syncwarp();
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const uint32_t soff = (uint32_t)lane * 128u + (uint32_t)((g ^ (lane & 7)) << 4);
              *(float4*)(st_out + soff) =
                  make_float4(__uint_as_float(va[4 * g]) * ep.alpha, __uint_as_float(va[4 * g + 1]) * ep.alpha,
                              __uint_as_float(va[4 * g + 2]) * ep.alpha, __uint_as_float(va[4 * g + 3]) * ep.alpha);
            }
            if ((c + 2 < BLOCK_N / 32) && (n0 + (c + 2) * 32 < N)) tmem_ld_32x32(trow + (c + 2) * 32, va);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
              asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                               (uint64_t)&tmC),
                           "r"(smem_u32(st_out)), "r"(n0 + c * 32), "r"(row_base + q * 32)
                           : "memory");
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          }
        } else {
        // ---------- bf16 output through swizzled smem staging + TMA store (coalesced, OOB-clipped) ----------
        // bias for all of this warp's chunks of the tile, fetched before the accumulator is ready
        __syncwarp();
#pragma unroll
        for (int cc = 0; cc < (BLOCK_N / 64 + 1) / 2; ++cc) {
          const int cb = n0 + (half + 2 * cc) * 64 + lane;
          wbias[cc * 64 + lane] = (ep.bias != nullptr && cb < N) ? ep.bias[cb] : 0.f;
          wbias[cc * 64 + 32 + lane] = (ep.bias != nullptr && cb + 32 < N) ? ep.bias[cb + 32] : 0.f;
        }
        __syncwarp();
        if (kInTile && lane == 0 && n0 + half * 64 < N) {   // first chunk's input tile, in flight while the MMAs finish
          mbar_expect_tx(in_bar, 4096);
          tma_load_2d(&tmP, in_bar, wst + 4096, n0 + half * 64, row_base + q * 32);
        }
        mbar_wait(&tmem_full[acc], acc_phase);
        tcgen05_fence_after();
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BLOCK_N;
        uint32_t va[32], vb[32];
        if (n0 + half * 64 < N) {  // first chunk's accumulator slabs in flight
          tmem_ld_32x32(trow + half * 64, va);
          tmem_ld_32x32(trow + half * 64 + 32, vb);
        }
#pragma unroll 1
        for (int c = half, cc = 0; c < BLOCK_N / 64; c += 2, ++cc) {
          if (n0 + c * 64 >= N) break;
          // staging: [out | pre] when the pre-activation side output exists (single-buffered), otherwise the
          // two buffers alternate so the TMA store of chunk i overlaps the TMEM reads / math of chunk i+1
          const bool single = has_pre || kInTile;   // second buffer taken by the side output / the input tile
          uint8_t* st_out = wst + ((!single && (chunk_ctr & 1)) ? 4096 : 0);
          uint8_t* st_pre = wst + 4096;
          ++chunk_ctr;
          if (lane == 0) {
            if (single) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          }
          tmem_ld_wait();
          __syncwarp();
          if (kInTile) {
            mbar_wait(in_bar, in_phase);
            in_phase ^= 1;
          }
          const bool more = (c + 2 < BLOCK_N / 64) && (n0 + (c + 2) * 64 < N);
          epi_slab<MODE>(va, wbias + cc * 64, ep, row, n0 + c * 64, M, N, lane, 0, st_out, st_pre, st_pre, rsc);
          if (more) tmem_ld_32x32(trow + (c + 2) * 64, va);       // next chunk's slabs stream in while this one
          epi_slab<MODE>(vb, wbias + cc * 64 + 32, ep, row, n0 + c * 64 + 32, M, N, lane, 1, st_out, st_pre, st_pre, rsc);
          if (more) tmem_ld_32x32(trow + (c + 2) * 64 + 32, vb);  // is converted, staged and stored
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (kInTile && more && lane == 0) {   // every lane has consumed the input tile: fetch the next one
            mbar_expect_tx(in_bar, 4096);
            tma_load_2d(&tmP, in_bar, st_pre, n0 + (c + 2) * 64, row_base + q * 32);
          }
          if (lane == 0 && !(dbg & 2)) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                             (uint64_t)&tmC),
                         "r"(smem_u32(st_out)), "r"(n0 + c * 64), "r"(row_base + q * 32)
                         : "memory");
            if (has_pre)
              asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                               (uint64_t)&tmP),
                           "r"(smem_u32(st_pre)), "r"(n0 + c * 64), "r"(row_base + q * 32)
                           : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        }
      } else {
        // ---------- direct stores (fp32 / split-K accumulate / unaligned outputs) ----------
#pragma unroll 1
        for (int c = half; c < BLOCK_N / 32; c += 2) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BLOCK_N + c * 32, v);
          tmem_ld_wait();
          const int col0 = n0 + c * 32;
          if (row < M && col0 < N) {
            float x[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = (__uint_as_float(v[j]) * ep.alpha + bias_s[c * 32 + j]) * rsc;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int col = col0 + g * 8;
              if (col >= N) break;
              float* xg = x + g * 8;
              const bool full = vec_ok && (col + 8 <= N);
              const int nvalid = (N - col) < 8 ? (N - col) : 8;
              if (ep.preact_out != nullptr) {
                bf16* dst = (bf16*)ep.preact_out + (size_t)row * ep.ld_pre + col;
                for (int j = 0; j < nvalid; ++j) dst[j] = __float2bfloat16_rn(xg[j]);
              }
              if (ep.act_aux != nullptr) {
                const bf16* src = (const bf16*)ep.act_aux + (size_t)row * ep.ld_aux + col;
                for (int j = 0; j < nvalid; ++j) xg[j] *= act_grad(__bfloat162float(src[j]), ep.act);
              } else if (ep.act != VALOR_ACT_NONE) {
#pragma unroll
                for (int j = 0; j < 8; ++j) xg[j] = act_fwd(xg[j], ep.act);
              }
              if (ep.residual != nullptr) {
                const bf16* src = (const bf16*)ep.residual + (size_t)row * ep.ldr + col;
                for (int j = 0; j < nvalid; ++j) xg[j] += __bfloat162float(src[j]);
              }
              if (ep.out_dtype == VALOR_DT_BF16) {
                bf16* dst = (bf16*)Cptr + (size_t)row * ldc + col;
                if (full) {
                  uint4 pk;
                  __nv_bfloat162* h = (__nv_bfloat162*)&pk;
#pragma unroll
                  for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(xg[2 * j], xg[2 * j + 1]);
                  *(uint4*)dst = pk;
                } else {
                  for (int j = 0; j < nvalid; ++j) dst[j] = __float2bfloat16_rn(xg[j]);
                }
              } else {
                float* dst = (float*)Cptr + (size_t)row * ldc + col;
                if (full && (ldc % 4 == 0)) {
                  if (ep.accumulate) {
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(xg[0]), "f"(xg[1]),
                                 "f"(xg[2]), "f"(xg[3])
                                 : "memory");
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "f"(xg[4]),
                                 "f"(xg[5]), "f"(xg[6]), "f"(xg[7])
                                 : "memory");
                  } else {
                    *(float4*)dst = make_float4(xg[0], xg[1], xg[2], xg[3]);
                    *(float4*)(dst + 4) = make_float4(xg[4], xg[5], xg[6], xg[7]);
                  }
                } else {
                  for (int j = 0; j < nvalid; ++j) {
                    if (ep.accumulate) atomicAdd(dst + j, xg[j]);
                    else dst[j] = xg[j];
                  }
                }
              }
            }
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) { if (PAIR) mbar_arrive_leader(&tmem_empty[acc]); else mbar_arrive(&tmem_empty[acc]); }
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
    if (tma_store && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }

  tcgen05_fence_before();
  if (PAIR) cluster_sync_all();   // no CTA of the pair leaves while the other may still signal its barriers
  else __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS));
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
  }
  return fn;
}

// 2-D bf16 tensor, `inner` contiguous, row pitch `ld` elements; box = {64, box_outer}, 128B swizzle.
static int make_tmap(CUtensorMap* tm, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_outer) {
  memset(tm, 0, sizeof(*tm));
  auto fn = get_encode_fn();
  VALOR_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {64, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VALOR_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d): inner=%llu outer=%llu ld=%llu ptr=%p", (int)r,
                (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld, ptr);
  return 0;
}

// 2-D fp32 tensor for the reduce-add epilogue: box = {32 floats (128 B), 32 rows}, 128B swizzle.
static int make_tmap_f32(CUtensorMap* tm, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld) {
  memset(tm, 0, sizeof(*tm));
  auto fn = get_encode_fn();
  VALOR_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 4};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VALOR_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(f32) failed (%d): inner=%llu outer=%llu ld=%llu ptr=%p", (int)r,
                (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld, ptr);
  return 0;
}

template <int BN, bool AK, bool BK, int MODE, bool BG = false, bool PAIR = false>
static int launch_cfg(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tp, void* C,
                      long long ldc, int M, int N, int K, int k_splits, int vec_ok, int tma_store,
                      const GemmEpilogue& ep, int grid, cudaStream_t st) {
  using Cfg = GemmCfg<BN, BG, PAIR>;
  auto kern = gemm_sm100_kernel<BN, AK, BK, MODE, BG, PAIR>;
  static bool attr_done = false;
  if (!attr_done) {
    VALOR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_done = true;
  }
  // release builds carry no result-changing switches; -DVALOR_DEBUG restores the diagnostic modes (bit 0: main loop
  // only, bit 1: skip the stores) used for the round-1 pipeline measurements
#ifdef VALOR_DEBUG
  static int dbg = -1;
  if (dbg < 0) { const char* e = getenv("VALOR_GEMM_DEBUG"); dbg = e ? atoi(e) : 0; if (dbg) fprintf(stderr, "valor_b200: VALOR_GEMM_DEBUG=%d active\n", dbg); }
#else
  const int dbg = 0;
#endif
  if (PAIR) {   // clusters of two CTAs (one SM pair per tile); `grid` counts CTAs and is even
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(384);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    VALOR_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, tc, tp, C, ldc, M, N, K, k_splits, vec_ok, tma_store, dbg, ep));
    return check_launch("gemm_sm100_kernel (two-CTA)");
  }
  kern<<<grid, 384, Cfg::SMEM_BYTES, st>>>(ta, tb, tc, tp, C, ldc, M, N, K, k_splits, vec_ok, tma_store, dbg, ep);
  return check_launch("gemm_sm100_kernel");
}

template <bool AK, bool BK, int MODE>
static int launch_bn(int bn, bool pair, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tp,
                     void* C, long long ldc, int M, int N, int K, int k_splits, int vec_ok, int tma_store,
                     const GemmEpilogue& ep, int grid, cudaStream_t st) {
  if (pair) return launch_cfg<256, AK, BK, MODE, false, true>(ta, tb, tc, tp, C, ldc, M, N, K, k_splits, vec_ok, tma_store, ep, grid, st);
  switch (bn) {
    case 64: return launch_cfg<64, AK, BK, MODE>(ta, tb, tc, tp, C, ldc, M, N, K, k_splits, vec_ok, tma_store, ep, grid, st);
    case 128: return launch_cfg<128, AK, BK, MODE>(ta, tb, tc, tp, C, ldc, M, N, K, k_splits, vec_ok, tma_store, ep, grid, st);
    case 192: return launch_cfg<192, AK, BK, MODE>(ta, tb, tc, tp, C, ldc, M, N, K, k_splits, vec_ok, tma_store, ep, grid, st);
    default: return launch_cfg<256, AK, BK, MODE>(ta, tb, tc, tp, C, ldc, M, N, K, k_splits, vec_ok, tma_store, ep, grid, st);
  }
}

// Can this device co-schedule the two CTAs of a pair with the kernel's shared-memory footprint?  Asked once
// (cudaOccupancyMaxActiveClusters on a representative instantiation); if not, every GEMM takes the single-CTA form.
static bool pair_supported() {
  static int ok = -1;
  if (ok < 0) {
    using Cfg = GemmCfg<256, false, true>;
    auto kern = gemm_sm100_kernel<256, true, true, EPI_PLAIN, false, true>;
    ok = 0;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) == cudaSuccess) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2);
      cfg.blockDim = dim3(384);
      cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) == cudaSuccess && n > 0) ok = 1;
    }
    if (!ok) {
      (void)cudaGetLastError();
      fprintf(stderr, "valor_b200: two-CTA GEMM tiles unavailable on this device (cluster occupancy 0): single-CTA tiles only\n");
    }
  }
  return ok == 1;
}

// Eligibility: bf16 operands, 16-byte aligned base pointers and row pitches.
bool gemm_sm100_eligible(const void* A, const void* B, long long lda, long long ldb, int M, int N, int K) {
  if (((uintptr_t)A | (uintptr_t)B) & 15) return false;
  if ((lda % 8) || (ldb % 8)) return false;
  if (M < 1 || N < 8 || K < 8) return false;
  return true;
}

// BLOCK_N: least padding waste, then enough tiles to fill the machine
static int pick_block_n(int M, int N, int accumulate, int sms) {
  const int m_blocks = (M + BLOCK_M - 1) / BLOCK_M;
  const int cands[3] = {256, 192, 128};
  int bn = 256;
  long best = -1;
  for (int i = 0; i < 3; ++i) {
    long waste = (long)((N + cands[i] - 1) / cands[i]) * cands[i] - N;
    if (best < 0 || waste < best) { best = waste; bn = cands[i]; }
  }
  if (N <= 64) bn = 64;
  while (bn > 64 && (long)m_blocks * ((N + bn - 1) / bn) < sms && !accumulate) bn = (bn == 192) ? 128 : bn / 2;
  return bn;
}

// Whether gemm_sm100 can produce the bias gradient inside the weight-gradient launch (same conditions as its TMA
// reduce-add epilogue); otherwise the caller adds a separate column-sum launch.  256-wide tiles keep ONE accumulator
// stage next to the sixteen bias-gradient columns (GemmCfg::ACC_STAGES).
bool gemm_sm100_fuses_bias_grad(const void* C, long long ldc, const GemmEpilogue& ep, int a_kmajor, int b_kmajor, int M, int N,
                                int force_bn) {
  (void)M; (void)N; (void)force_bn;
#ifdef VALOR_DEBUG
  { const char* e = getenv("VALOR_GEMM_NO_TMA_REDUCE"); if (e && atoi(e)) return false; }
#endif
  return ep.accumulate && ep.out_dtype == VALOR_DT_F32 && (ldc % 4 == 0) && (((uintptr_t)C & 15) == 0) &&
         ep.bias == nullptr && ep.residual == nullptr && ep.act_aux == nullptr && ep.preact_out == nullptr &&
         ep.act == VALOR_ACT_NONE && !a_kmajor && !b_kmajor;
}

int gemm_sm100(const void* A, long long lda, int a_kmajor, const void* B, long long ldb, int b_kmajor, void* C,
               long long ldc, int M, int N, int K, const GemmEpilogue& ep, int force_bn, int force_splits,
               cudaStream_t st) {
  VALOR_REQUIRE(gemm_sm100_eligible(A, B, lda, ldb, M, N, K), "gemm_sm100: operands not TMA-eligible");
  VALOR_REQUIRE(!ep.accumulate || ep.out_dtype == VALOR_DT_F32, "gemm_sm100: accumulate needs fp32 output");
  VALOR_REQUIRE(ep.residual == nullptr || ep.res_dtype == VALOR_DT_BF16, "gemm_sm100: residual must be bf16");
  VALOR_REQUIRE(ep.act_aux == nullptr || ep.aux_dtype == VALOR_DT_BF16, "gemm_sm100: act_aux must be bf16");
  const int sms = num_sms();
  // force_bn + 1000: require the two-CTA form, + 2000: forbid it (tests / A-B measurements)
  const int pair_req = force_bn / 1000;
  force_bn %= 1000;
  int bn = force_bn ? force_bn : pick_block_n(M, N, ep.accumulate, sms);
  // two-CTA tiles (256 x 256 over an SM pair): the 128 x 256 single-CTA tile is bound by operand traffic from L2, the
  // pair halves the B traffic per SM.  Worth it when the 256-row tiles still fill the machine.
  bool pair = bn == 256 && M > BLOCK_M && pair_supported();
  if (pair && pair_req != 1) {
    const long tiles2 = (long)((M + 2 * BLOCK_M - 1) / (2 * BLOCK_M)) * ((N + bn - 1) / bn);
    pair = ep.accumulate ? tiles2 >= 4 : tiles2 >= (long)(sms / 2);
  }
  if (pair_req == 2) pair = false;
  VALOR_REQUIRE(pair_req != 1 || pair, "gemm_sm100: the two-CTA form needs BLOCK_N = 256 and M > 128");
  const int m_blocks = pair ? (M + 2 * BLOCK_M - 1) / (2 * BLOCK_M) : (M + BLOCK_M - 1) / BLOCK_M;
  const int units = pair ? sms / 2 : sms;                 // persistent work consumers (SM pairs / SMs)
  const int n_blocks = (N + bn - 1) / bn;
  const int kb_total = (K + BLOCK_K - 1) / BLOCK_K;
  int k_splits = 1;
  if (force_splits > 0) {
    k_splits = force_splits;
  } else if (ep.accumulate && ep.out_dtype == VALOR_DT_F32) {
    // split-K chosen to minimise the makespan of the persistent schedule: rounds of `sms` work items, each
    // costing its k-blocks plus a fixed epilogue/drain term (in k-block units)
    const long tiles = (long)m_blocks * n_blocks;
    const int max_splits = kb_total / 4 > 0 ? kb_total / 4 : 1;
    const long epi_cost = 6;
    long best_cost = -1;
    for (int s = 1; s <= max_splits && s <= 512; ++s) {
      const long per = (kb_total + s - 1) / s;
      const long s_eff = (kb_total + per - 1) / per;
      const long rounds = (tiles * s_eff + units - 1) / units;
      const long cost = rounds * (per + epi_cost);
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; k_splits = (int)s_eff; }
    }
  }
  if (k_splits > kb_total) k_splits = kb_total;
  if (k_splits < 1) k_splits = 1;
  {  // every split must own >= 1 k-block
    int per = (kb_total + k_splits - 1) / k_splits;
    k_splits = (kb_total + per - 1) / per;
  }
  VALOR_REQUIRE(k_splits == 1 || ep.accumulate, "gemm_sm100: split-K requires accumulate");

  CUtensorMap ta, tb;
  if (a_kmajor) { if (make_tmap(&ta, A, K, M, lda, BLOCK_M)) return 1; }
  else          { if (make_tmap(&ta, A, M, K, lda, BLOCK_K)) return 1; }
  if (b_kmajor) { if (make_tmap(&tb, B, K, N, ldb, pair ? bn / 2 : bn)) return 1; }   // pair: each CTA stages half of the B tile
  else          { if (make_tmap(&tb, B, N, K, ldb, BLOCK_K)) return 1; }

  int vec_ok = (ldc % 8 == 0) && (((uintptr_t)C & 15) == 0);
  if (ep.residual) vec_ok = vec_ok && (ep.ldr % 8 == 0) && (((uintptr_t)ep.residual & 15) == 0);
  if (ep.act_aux) vec_ok = vec_ok && (ep.ld_aux % 8 == 0) && (((uintptr_t)ep.act_aux & 15) == 0);
  if (ep.preact_out) vec_ok = vec_ok && (ep.ld_pre % 8 == 0) && (((uintptr_t)ep.preact_out & 15) == 0);

  // bf16 outputs with 16-byte aligned pitches leave through swizzled smem staging + TMA store
  int tma_store = (ep.out_dtype == VALOR_DT_BF16) && (ldc % 8 == 0) && (((uintptr_t)C & 15) == 0) && !ep.accumulate;
  if (ep.preact_out) tma_store = tma_store && (ep.ld_pre % 8 == 0) && (((uintptr_t)ep.preact_out & 15) == 0);
  if (ep.residual) tma_store = tma_store && (ep.ldr % 8 == 0) && (((uintptr_t)ep.residual & 15) == 0);
  if (ep.act_aux) tma_store = tma_store && (ep.ld_aux % 8 == 0) && (((uintptr_t)ep.act_aux & 15) == 0);
  CUtensorMap tc, tp;
  memset(&tc, 0, sizeof(tc));
  memset(&tp, 0, sizeof(tp));
  // fp32 accumulation (weight gradients) leaves through TMA reduce-add
#ifdef VALOR_DEBUG
  static int no_acc_tma = -1;
  if (no_acc_tma < 0) { const char* e = getenv("VALOR_GEMM_NO_TMA_REDUCE"); no_acc_tma = e ? atoi(e) : 0; }
#else
  const int no_acc_tma = 0;
#endif
  const bool acc_tma = ep.accumulate && ep.out_dtype == VALOR_DT_F32 && (ldc % 4 == 0) && (((uintptr_t)C & 15) == 0) &&
                       ep.bias == nullptr && ep.residual == nullptr && ep.act_aux == nullptr && ep.preact_out == nullptr &&
                       ep.act == VALOR_ACT_NONE && !a_kmajor && !b_kmajor && !no_acc_tma;
  VALOR_REQUIRE(ep.bias_grad == nullptr || acc_tma, "gemm_sm100: bias_grad rides on the TMA reduce-add weight-gradient path only");
  if (acc_tma) {
    tma_store = 1;
    if (make_tmap_f32(&tc, C, N, M, ldc)) return 1;
  } else if (tma_store) {
    if (make_tmap(&tc, C, N, M, ldc, 32)) return 1;
    if (ep.preact_out && make_tmap(&tp, ep.preact_out, N, M, ep.ld_pre, 32)) return 1;
  }
  const long total = (long)m_blocks * n_blocks * k_splits;
  const int grid = pair ? 2 * (int)(total < units ? total : units) : (int)(total < sms ? total : sms);
  // epilogue specialisation (TMA-store path only; everything else runs the generic code)
  int mode = EPI_GENERIC;
  if (acc_tma) {
    mode = EPI_ACC;
  } else if (tma_store) {
    const bool res = ep.residual != nullptr, aux = ep.act_aux != nullptr, pre = ep.preact_out != nullptr;
    if (res && !aux && !pre && ep.act == VALOR_ACT_NONE) { if (make_tmap(&tp, ep.residual, N, M, ep.ldr, 32)) return 1; }
    if (!res && aux && !pre && ep.act == VALOR_ACT_GELU && ep.bias == nullptr) { if (make_tmap(&tp, ep.act_aux, N, M, ep.ld_aux, 32)) return 1; }
    if (!res && !aux && !pre && ep.act == VALOR_ACT_NONE && ep.row_scale == nullptr) mode = EPI_PLAIN;
    else if (res && !aux && !pre && ep.act == VALOR_ACT_NONE) mode = EPI_RES;
    else if (!res && !aux && pre && ep.act == VALOR_ACT_GELU) mode = EPI_GELU_PRE;
    else if (!res && aux && !pre && ep.act == VALOR_ACT_GELU && ep.bias == nullptr) mode = EPI_GELU_AUX;
  }
#define VALOR_LAUNCH(AK, BK, MODE) \
  return launch_bn<AK, BK, MODE>(bn, pair, ta, tb, tc, tp, C, ldc, M, N, K, k_splits, vec_ok, tma_store, ep, grid, st)
  if (a_kmajor && b_kmajor) {
    switch (mode) {
      case EPI_PLAIN: VALOR_LAUNCH(true, true, EPI_PLAIN);
      case EPI_RES: VALOR_LAUNCH(true, true, EPI_RES);
      case EPI_GELU_PRE: VALOR_LAUNCH(true, true, EPI_GELU_PRE);
      default: VALOR_LAUNCH(true, true, EPI_GENERIC);
    }
  }
  if (a_kmajor && !b_kmajor) {
    switch (mode) {
      case EPI_PLAIN: VALOR_LAUNCH(true, false, EPI_PLAIN);
      case EPI_GELU_AUX: VALOR_LAUNCH(true, false, EPI_GELU_AUX);
      default: VALOR_LAUNCH(true, false, EPI_GENERIC);
    }
  }
  if (!a_kmajor && b_kmajor) VALOR_LAUNCH(false, true, EPI_GENERIC);
  if (mode == EPI_ACC && ep.bias_grad != nullptr) {   // weight gradient + bias gradient in one launch
    if (pair) return launch_cfg<256, false, false, EPI_ACC, true, true>(ta, tb, tc, tp, C, ldc, M, N, K, k_splits, vec_ok, tma_store, ep, grid, st);
    switch (bn) {
      case 64: return launch_cfg<64, false, false, EPI_ACC, true>(ta, tb, tc, tp, C, ldc, M, N, K, k_splits, vec_ok, tma_store, ep, grid, st);
      case 128: return launch_cfg<128, false, false, EPI_ACC, true>(ta, tb, tc, tp, C, ldc, M, N, K, k_splits, vec_ok, tma_store, ep, grid, st);
      case 192: return launch_cfg<192, false, false, EPI_ACC, true>(ta, tb, tc, tp, C, ldc, M, N, K, k_splits, vec_ok, tma_store, ep, grid, st);
      default: return launch_cfg<256, false, false, EPI_ACC, true>(ta, tb, tc, tp, C, ldc, M, N, K, k_splits, vec_ok, tma_store, ep, grid, st);
    }
  }
  if (mode == EPI_ACC) VALOR_LAUNCH(false, false, EPI_ACC);
  VALOR_LAUNCH(false, false, EPI_GENERIC);
#undef VALOR_LAUNCH
}

}  // namespace valor
