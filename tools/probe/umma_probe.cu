// TEST TOOL (not part of libvalor_b200.so): validates the tcgen05 shared-memory / tensor-memory operand forms the
// attention kernels rely on, one MMA chain per launch, with every descriptor field supplied by the host so a single
// GPU call can sweep candidates.  Build: tools/probe/build.sh -> tools/probe/libumma_probe.so (ctypes from
// tools/probe/run_probe.py).
//
//   D[128, N] (fp32, TMEM) = sum over ksteps of  A_k[128 x 16] . B_k[N x 16]^T
//
// A and B are copied from global memory into shared memory by a host-chosen byte map (a_map / b_map: for every 16-byte
// chunk of the source image its destination byte offset), so the probe places data exactly like the attention kernels
// do (cp.async gathers into swizzled tiles).  a_desc_lo/hi, b_desc_*: descriptor template; the start-address field is
// added on the device: desc(k) = template + ((base + k * step) >> 4).
// a_tmem != 0: the A operand is first copied into tensor memory (tcgen05.st, one row per lane, packed bf16 pairs) and
// the MMA is issued in the "TS" form.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct ProbeArgs {
  const uint4* a_src; const int* a_map; int a_chunks;   // a_map[i] = destination byte offset of 16-byte chunk i (or -1)
  const uint4* b_src; const int* b_map; int b_chunks;
  unsigned long long a_desc, b_desc;                     // templates (start address field = 0)
  int a_step, b_step;                                    // byte advance per k-step
  int ksteps;
  unsigned int idesc;
  int N;
  int a_tmem;                                            // 1: A through tensor memory
  const uint32_t* a_tmem_words; int a_tmem_cols;         // [128][a_tmem_cols] 32-bit words (row-major), cols per k-step = 8
  float* D;                                              // [128, N]
};

__global__ void __launch_bounds__(128, 1) umma_probe_kernel(ProbeArgs P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;
  uint8_t* sb = smem + 64 * 1024;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 1024 / 16; i += 128) { ((uint4*)sa)[i] = make_uint4(0, 0, 0, 0); ((uint4*)sb)[i] = make_uint4(0, 0, 0, 0); }
  __syncthreads();
  for (int i = threadIdx.x; i < P.a_chunks; i += 128) if (P.a_map[i] >= 0) *(uint4*)(sa + P.a_map[i]) = P.a_src[i];
  for (int i = threadIdx.x; i < P.b_chunks; i += 128) if (P.b_map[i] >= 0) *(uint4*)(sb + P.b_map[i]) = P.b_src[i];
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the MMA unit
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const uint32_t tmem_a = tmem + 256;   // A staging columns (TS form)
  if (P.a_tmem) {
    // lane (row) r of warp w writes its packed words: row = warp*32 + lane
    const int row = warp * 32 + lane;
    for (int c = 0; c < P.a_tmem_cols; c += 8) {
      uint32_t v[8];
      for (int j = 0; j < 8; ++j) v[j] = P.a_tmem_words[(size_t)row * P.a_tmem_cols + c + j];
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tmem_a + ((uint32_t)(warp * 32) << 16) + c),
                   "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]));
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    for (int k = 0; k < P.ksteps; ++k) {
      const unsigned long long bd = P.b_desc + (unsigned long long)(((smem_u32(sb) + k * P.b_step) & 0x3FFFF) >> 4);
      if (P.a_tmem) {
        const uint32_t ta = tmem_a + k * 8;
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(tmem), "r"(ta), "l"(bd), "r"(P.idesc),
                     "r"(k > 0 ? 1u : 0u) : "memory");
      } else {
        const unsigned long long ad = P.a_desc + (unsigned long long)(((smem_u32(sa) + k * P.a_step) & 0x3FFFF) >> 4);
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(P.idesc),
                     "r"(k > 0 ? 1u : 0u) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  {
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int row = warp * 32 + lane;
  for (int c = 0; c < P.N; c += 8) {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) P.D[(size_t)row * P.N + c + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

extern "C" int umma_probe(const void* a_src, const int* a_map, int a_chunks, const void* b_src, const int* b_map, int b_chunks,
                          unsigned long long a_desc, unsigned long long b_desc, int a_step, int b_step, int ksteps,
                          unsigned int idesc, int N, int a_tmem, const void* a_tmem_words, int a_tmem_cols, float* D) {
  ProbeArgs P;
  P.a_src = (const uint4*)a_src; P.a_map = a_map; P.a_chunks = a_chunks;
  P.b_src = (const uint4*)b_src; P.b_map = b_map; P.b_chunks = b_chunks;
  P.a_desc = a_desc; P.b_desc = b_desc; P.a_step = a_step; P.b_step = b_step; P.ksteps = ksteps; P.idesc = idesc; P.N = N;
  P.a_tmem = a_tmem; P.a_tmem_words = (const uint32_t*)a_tmem_words; P.a_tmem_cols = a_tmem_cols; P.D = D;
  const int smem = 128 * 1024 + 1024;
  static bool done = false;
  if (!done) { cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); done = true; }
  umma_probe_kernel<<<1, 128, smem>>>(P);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { fprintf(stderr, "umma_probe: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
