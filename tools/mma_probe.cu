// Micro-probe: what paces the tcgen05 GEMM main loop on B200?  One CTA per SM runs, separately and together,
//   (M) the MMA stream of the production kernel: per k-block 4 x tcgen05.mma.cta_group::1.kind::f16 (M 128, N = BN, K 16),
//       both operands from shared memory (128B-swizzled K-major tiles), one tcgen05.commit per k-block;
//   (T) the TMA stream: per k-block one A box (128 x 64 bf16 = 16 KB) and one B box (BN x 64) into a ring of `stages`
//       slots, L2-resident sources laid out like the 16384 x 320 activations / 320 x 2880 weights of a 32x32-level conv;
// without any dependence between the two (the MMAs read whatever is in the slots: timing only).  Prints clk and ns per
// k-block for M alone, T alone and both at once.  If "both" ~ max(M, T) the streams overlap and the production loop
// (260-300 ns) loses its time in the hand-shakes; if "both" ~ M + T they share a resource (shared-memory bandwidth).
//   build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/mma_probe.cu -o makeupdiffuse_b200/build/mma_probe
//   run:    makeupdiffuse_b200/build/mma_probe            (profiles/r01_mma_probe.txt)
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_));        \
      exit(1);                                                                                  \
    }                                                                                           \
  } while (0)

constexpr int BM = 128, BK = 64, A_BYTES = BM * BK * 2;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nW1:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D1;\nbra W1;\nD1:\n}\n" ::"r"(smem_u32(b)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b),
               "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int bn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
  return t;
}

// out[cta * 4 + {0, 1}] = clk of the MMA / TMA stream, out[cta * 4 + {2, 3}] = ns of the same
template <int BN>
__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap, int iters,
                                                int mode, int stages, int kdiv, long long* out) {
  constexpr int STAGE = A_BYTES + BN * BK * 2;
  constexpr int TCOLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : BN <= 256 ? 256 : 512;
  extern __shared__ unsigned char raw[];
  unsigned char* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + stages * STAGE);  // [0] mma done, [1] commit sink, [2 + s] tma full
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2 + 2 * stages + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < stages * STAGE / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // finite bf16
  if (threadIdx.x == 0) {
    mbar_init(bars + 0, 1);
    mbar_init(bars + 1, 1 << 19);  // never completes: per-k-block commits land here like the production "empty" commits
    for (int s = 0; s < stages; ++s) mbar_init(bars + 2 + s, 1);
    for (int s = 0; s < stages; ++s) mbar_init(bars + 2 + stages + 1 + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(slot)), "r"(TCOLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = *slot;
  if (warp == 1 && lane == 0 && (mode & 1) && mode < 4) {
    constexpr uint32_t idesc = make_idesc(BN);
    const long long c0 = clock64();
    const unsigned long long g0 = gtimer();
    for (int it = 0; it < iters; ++it) {
      const uint32_t sa = smem_u32(smem + (it % stages) * STAGE);
      const uint64_t ad = make_smem_desc(sa), bd = make_smem_desc(sa + A_BYTES);
#pragma unroll
      for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (it || k) ? 1u : 0u);
      tc_commit(bars + 1);
    }
    tc_commit(bars + 0);
    mbar_wait(bars + 0, 0);
    out[blockIdx.x * 4 + 0] = clock64() - c0;
    out[blockIdx.x * 4 + 2] = (long long)(gtimer() - g0);
  }
  // mode 4: the production hand-shake — MMAs of k-block `it` wait for its TMA (full[s]), their commit releases the slot
  // (empty[s] = bars[2 + stages + s]), the TMA of k-block it + stages waits for that release
  if (mode >= 4) {
    uint64_t* full = bars + 2;
    uint64_t* empty = bars + 2 + stages + 1;
    if (warp == 1 && lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN);
      const long long c0 = clock64();
      const unsigned long long g0 = gtimer();
      for (int it = 0; it < iters; ++it) {
        const int s = it % stages;
        mbar_wait(full + s, (it / stages) & 1);
        if (mode == 4 || mode == 7) asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const uint32_t sa = smem_u32(smem + s * STAGE);
        const uint64_t ad = make_smem_desc(sa), bd = make_smem_desc(sa + A_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (it || k) ? 1u : 0u);
        if (mode == 7) {  // release through a plain arrive two k-blocks later instead of a commit per k-block
          if (it >= 2) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(empty + (it - 2) % stages)) : "memory");
        } else {
          tc_commit(empty + s);
        }
      }
      tc_commit(bars + 0);
      mbar_wait(bars + 0, 0);
      out[blockIdx.x * 4 + 0] = clock64() - c0;
      out[blockIdx.x * 4 + 2] = (long long)(gtimer() - g0);
    }
    if (warp == 2 && lane == 0) {
      const int m0 = (blockIdx.x % 128) * BM;
      for (int it = 0; it < iters; ++it) {
        const int s = it % stages;
        if (it >= stages) mbar_wait(empty + s, ((it / stages) - 1) & 1);
        unsigned char* dst = smem + s * STAGE;
        mbar_expect_tx(full + s, STAGE);
        tma_load_2d(&amap, full + s, dst, (it % kdiv) * BK, m0);
        tma_load_2d(&bmap, full + s, dst + A_BYTES, (it % 45) * BK, (blockIdx.x & 1) * BN);
      }
    }
  }
  if (warp == 2 && lane == 0 && (mode & 2) && mode < 4) {
    const int m0 = (blockIdx.x % 128) * BM;
    const long long c0 = clock64();
    const unsigned long long g0 = gtimer();
    for (int it = 0; it < iters; ++it) {
      const int s = it % stages;
      if (it >= stages) mbar_wait(bars + 2 + s, ((it / stages) - 1) & 1);  // the previous load into this slot has landed
      unsigned char* dst = smem + s * STAGE;
      mbar_expect_tx(bars + 2 + s, STAGE);
      tma_load_2d(&amap, bars + 2 + s, dst, (it % kdiv) * BK, m0);
      tma_load_2d(&bmap, bars + 2 + s, dst + A_BYTES, (it % 45) * BK, (blockIdx.x & 1) * BN);
    }
    for (int it = max(0, iters - stages); it < iters; ++it) mbar_wait(bars + 2 + (it % stages), (it / stages) & 1);
    out[blockIdx.x * 4 + 1] = clock64() - c0;
    out[blockIdx.x * 4 + 3] = (long long)(gtimer() - g0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(TCOLS));
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeFn fn, void* base, uint64_t cols, uint64_t rows, uint32_t box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {cols, rows}, str[1] = {cols * 2};
  cuuint32_t box[2] = {BK, box_rows}, es[2] = {1, 1};
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "cuTensorMapEncodeTiled failed: %d\n", (int)r);
    exit(1);
  }
  return m;
}

template <int BN>
static void run(EncodeFn fn, void* A, void* B, long long* dout, int stages, int kdiv) {
  constexpr int STAGE = A_BYTES + BN * BK * 2;
  const int iters = 400;
  CUtensorMap amap = make_map(fn, A, 64 * kdiv, 16384, BM), bmap = make_map(fn, B, 2880, 512, BN);
  const size_t smem = (size_t)stages * STAGE + (4 + 2 * stages) * 8 + 16 + 1024;
  CK(cudaFuncSetAttribute(probe<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  printf("N tile %3d  %d stages x %5.1f KB  A tile walks K = %4d:", BN, stages, STAGE / 1024.0, 64 * kdiv);
  for (int mode = 1; mode <= 7; ++mode) {
    if (mode == 6 || (mode == 7 && stages <= 2)) continue;
    std::vector<long long> h(148 * 4);
    for (int rep = 0; rep < 3; ++rep) {  // last repetition counts (warm L2, warm instruction cache)
      CK(cudaMemset(dout, 0, 148 * 4 * sizeof(long long)));
      probe<BN><<<148, 128, smem>>>(amap, bmap, iters, mode, stages, kdiv, dout);
      CK(cudaGetLastError());
      CK(cudaDeviceSynchronize());
    }
    CK(cudaMemcpy(h.data(), dout, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    auto med = [&](int col) {
      std::vector<long long> v;
      for (int c = 0; c < 148; ++c) v.push_back(h[c * 4 + col]);
      std::sort(v.begin(), v.end());
      return (double)v[74] / iters;
    };
    const char* nm = mode == 1 ? "MMA alone" : mode == 2 ? "TMA alone" : "both, independent";
    if (mode >= 4) printf("  | %s: %5.0f clk %4.0f ns", mode == 4 ? "coupled ring" : mode == 5 ? "coupled, no tcgen05 fence" : "coupled, release by plain arrive (UNSAFE, timing only)", med(0), med(2));
    else {
      if (mode & 1) printf("  | %s: MMA %5.0f clk %4.0f ns", nm, med(0), med(2));
      if (mode & 2) printf("  %s TMA %5.0f clk %4.0f ns", mode == 2 ? "| TMA alone:" : "/", med(1), med(3));
    }
  }
  printf("   (per k-block)\n");
}


// coupled ring with KB k-blocks (of 64) per stage: one full / empty hand-shake per KB * 64 of K, 2 * KB TMA boxes and
// 4 * KB MMAs per stage.  out[cta * 4 + 0 / 2] = clk / ns of the MMA stream per STAGE
template <int BN, int KB>
__global__ void __launch_bounds__(128, 1) probe_kb(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap, int iters,
                                                   int stages, long long* out) {
  constexpr int SUB = A_BYTES + BN * BK * 2, STAGE = KB * SUB;
  constexpr int TCOLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : BN <= 256 ? 256 : 512;
  extern __shared__ unsigned char raw[];
  unsigned char* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + stages * STAGE);
  uint64_t *full = bars + 1, *empty = bars + 1 + stages;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2 + 2 * stages);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < stages * STAGE / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(bars + 0, 1);
    for (int s = 0; s < 2 * stages; ++s) mbar_init(bars + 1 + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(slot)), "r"(TCOLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = *slot;
  if (warp == 1 && lane == 0) {
    constexpr uint32_t idesc = make_idesc(BN);
    const long long c0 = clock64();
    const unsigned long long g0 = gtimer();
    for (int it = 0; it < iters; ++it) {
      const int s = it % stages;
      mbar_wait(full + s, (it / stages) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        const uint32_t sa = smem_u32(smem + s * STAGE + kb * SUB);
        const uint64_t ad = make_smem_desc(sa), bd = make_smem_desc(sa + A_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (it || kb || k) ? 1u : 0u);
      }
      tc_commit(empty + s);
    }
    tc_commit(bars + 0);
    mbar_wait(bars + 0, 0);
    out[blockIdx.x * 4 + 0] = clock64() - c0;
    out[blockIdx.x * 4 + 2] = (long long)(gtimer() - g0);
  }
  if (warp == 2 && lane == 0) {
    const int m0 = (blockIdx.x % 128) * BM;
    for (int it = 0; it < iters; ++it) {
      const int s = it % stages;
      if (it >= stages) mbar_wait(empty + s, ((it / stages) - 1) & 1);
      mbar_expect_tx(full + s, STAGE);
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        unsigned char* dst = smem + s * STAGE + kb * SUB;
        tma_load_2d(&amap, full + s, dst, ((it * KB + kb) % 5) * BK, m0);
        tma_load_2d(&bmap, full + s, dst + A_BYTES, ((it * KB + kb) % 45) * BK, (blockIdx.x & 1) * BN);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(TCOLS));
  }
}

template <int BN, int KB>
static void run_kb(EncodeFn fn, void* A, void* B, long long* dout, int stages) {
  constexpr int STAGE = KB * (A_BYTES + BN * BK * 2);
  const int iters = 400 / KB;
  CUtensorMap amap = make_map(fn, A, 320, 16384, BM), bmap = make_map(fn, B, 2880, 512, BN);
  const size_t smem = (size_t)stages * STAGE + (4 + 2 * stages) * 8 + 16 + 1024;
  CK(cudaFuncSetAttribute(probe_kb<BN, KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  std::vector<long long> h(148 * 4);
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaMemset(dout, 0, 148 * 4 * sizeof(long long)));
    probe_kb<BN, KB><<<148, 128, smem>>>(amap, bmap, iters, stages, dout);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
  }
  CK(cudaMemcpy(h.data(), dout, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
  std::vector<long long> c, n;
  for (int i = 0; i < 148; ++i) { c.push_back(h[i * 4]); n.push_back(h[i * 4 + 2]); }
  std::sort(c.begin(), c.end());
  std::sort(n.begin(), n.end());
  printf("coupled ring, N tile %3d, %d x 64 of K per hand-shake, %d stages x %5.1f KB: %5.0f clk %4.0f ns per 64 of K  (MMA floor %d clk)\n", BN, KB,
         stages, STAGE / 1024.0, (double)c[74] / iters / KB, (double)n[74] / iters / KB, 2 * BN);
}

int main() {
  setvbuf(stdout, nullptr, _IOLBF, 0);
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  EncodeFn fn = reinterpret_cast<EncodeFn>(p);
  void *A, *B;
  long long* dout;
  CK(cudaMalloc(&A, (size_t)16384 * 2880 * 2));
  CK(cudaMalloc(&B, (size_t)512 * 2880 * 2));
  CK(cudaMalloc(&dout, 148 * 4 * sizeof(long long)));
  CK(cudaMemset(A, 0x3c, (size_t)16384 * 2880 * 2));
  CK(cudaMemset(B, 0x3c, (size_t)512 * 2880 * 2));
  printf("# tools/mma_probe.cu, 148 CTAs (1 per SM), 400 k-blocks each; MMA floor at N: 4 x 128 * N / 256 clk = N / 2 clk per MMA\n");
  if (getenv("MMA_PROBE_FULL")) {
    run<64>(fn, A, B, dout, 5, 5);
    run<128>(fn, A, B, dout, 5, 5);
    run<160>(fn, A, B, dout, 5, 5);
    run<160>(fn, A, B, dout, 3, 5);
    run<160>(fn, A, B, dout, 5, 45);
    run<256>(fn, A, B, dout, 3, 5);
    run<160>(fn, A, B, dout, 4, 5);
    run<160>(fn, A, B, dout, 2, 5);
  }
  run_kb<160, 1>(fn, A, B, dout, 5);
  run_kb<160, 1>(fn, A, B, dout, 3);
  run_kb<160, 2>(fn, A, B, dout, 2);
  run_kb<160, 2>(fn, A, B, dout, 3);
  run_kb<160, 3>(fn, A, B, dout, 2);
  run_kb<80, 2>(fn, A, B, dout, 3);
  run_kb<80, 4>(fn, A, B, dout, 2);
  run_kb<256, 1>(fn, A, B, dout, 4);
  run_kb<256, 2>(fn, A, B, dout, 2);
  return 0;
}
