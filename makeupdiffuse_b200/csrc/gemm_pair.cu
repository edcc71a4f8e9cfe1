// CTA-pair implicit-GEMM convolution / GEMM on the 5th-generation tensor cores (sm_100a), cta_group::2:
//
//   D[256 x UN] (fp32; rows 0-127 in the leader CTA's TMEM, rows 128-255 in its peer's)
//        += A[256 x 64] (bf16, 128 rows from each CTA's shared memory) * B[UN x 64]^T (UN / 2 weight rows from each CTA)
//
// One work unit = 256 output rows x (NSUB * UN) output channels, computed by the two SMs of a cluster of two CTAs.
// Each CTA streams only its own 128 rows of A and HALF of the unit's weight rows per k-block; with NSUB = 2 every A
// tile meets two N = 160 weight tiles, so per 64 of K an SM fetches 16 KB of activations for 640 tensor-core cycles
// where the single-CTA kernel (gemm_tcgen05.cu) fetches 16 KB for 320.  That ratio is what paced the old main loop:
// the activation tiles are distinct per CTA and arrive from L2 at ~44 B/clk per SM when all SMs pull
// (tools/mma_probe.cu: 366 clk per k-block for the TMA stream alone against 320-345 clk of MMAs).
//
// Warp roles (384 threads; both CTAs of the pair run the same code):
//   warp 0      TMA producer: A box (own rows) + NSUB B boxes (own half of the weight rows) per k-block into a
//               `stages`-deep ring; loads are the .cta_group::2 form, their bytes complete on the LEADER's full barrier.
//   warp 1      MMA issuer (leader CTA only): tcgen05.mma.cta_group::2, M = 256, N = UN; its commits are multicast to
//               both CTAs' empty / tmem_full barriers.  Owns the (cta_group::2) TMEM allocation.
//   warp 2      epilogue DMA thread (lane 0): TMA loads of the residual panels into panel slots ahead of the epilogue
//               warps, TMA stores of the finished panels (cp.async.bulk.tensor shared -> global), slot recycling by
//               bulk-group completion.  No thread ever issues a global load or store for tile data.
//   warps 4-11  epilogue: two warpgroups take alternate 32-column panels.  thread = accumulator row = TMEM lane:
//               tcgen05.ld 32 columns -> (+bias, +timestep embedding) * alpha + residual [SiLU] -> fp32 and / or bf16
//               panel in shared memory, in the 128B / 64B TMA swizzle, so every shared-memory access is conflict
//               free; optional GroupNorm statistics (column sum, sum of squares over the CTA's 128 rows) by a
//               register transpose-reduction (31 shuffles per statistic) + one 4-warp combine.
//   GEGLU mode (UN = 256, weight rows blocked [128 value | 128 gate]): value columns 0-127 come from the leader's
//   weight rows, gate columns 128-255 from the peer's; a panel = 32 value + 32 gate columns -> 32 bf16 outputs.
//
// Split-K units write raw fp32 partials (same epilogue, no operands) to the workspace [split][M][N]; the reducer of
// gemm_tcgen05.cu applies the fused epilogue.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
using namespace mkd;

namespace mkd {
int splitk_reduce(const mkd_conv_desc* d, int M, int pix_per_img, int splits, cudaStream_t stream);  // gemm_tcgen05.cu
unsigned long long* debug_trace_ptr();                                                               // gemm_tcgen05.cu
}

namespace {

constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int A_BYTES = BM * BK * 2;           // 16 KB: one CTA's activation tile per k-block
constexpr int PANEL = 32;                      // epilogue panel width (columns)
constexpr int P32_BYTES = BM * PANEL * 4;      // 16 KB: fp32 panel (128 B rows, SWIZZLE_128B)
constexpr int P16_BYTES = BM * PANEL * 2;      // 8 KB: bf16 panel (64 B rows, SWIZZLE_64B)
constexpr int MAX_STAGES = 8, MAX_SLOTS = 8;
static_assert(MAX_STAGES == MAX_SLOTS, "barrier init assumes one lane per ring / slot barrier");
constexpr int EPI_WARP0 = 4, EPI_WARPS = 8;    // EPI_WARP0 % 4 == 0: warp w reads TMEM lanes [32 (w % 4), +32)
constexpr int THREADS = (EPI_WARP0 + EPI_WARPS) * 32;
constexpr int TMEM_COLS = 512;
constexpr int SCRATCH_BYTES = 2 /*warpgroups*/ * 2 /*buffers*/ * 4 /*quadrants*/ * 32 * 8;  // GroupNorm partials
constexpr int BIASV_FLOATS = 320;  // per-unit column vector (bias [+ the tile's timestep-embedding row]): NSUB * UN floats
constexpr int BIASV_BYTES = 2 /*warpgroups*/ * 2 /*buffers*/ * BIASV_FLOATS * 4;
constexpr int BAR_BYTES = 512 + BIASV_BYTES;
constexpr int SMEM_LIMIT = 227 * 1024;

// division by a launch constant: q = umulhi(n, ceil(2^32 / d)), exact while n * d < 2^32 (checked on the host).
// Every role decodes work units; 32-bit integer divisions cost ~150 clk each, and the epilogue DMA thread ran eight of
// them per PANEL (its pace, not the tensor cores', set the epilogue's).
struct FastDiv {
  uint32_t d, m;  // m == 0: d == 1
};
__device__ __forceinline__ int fdiv(int n, FastDiv f) { return f.m ? (int)__umulhi((uint32_t)n, f.m) : n; }

struct PairP {
  // main loop
  int kblocks, kb_per_split;
  int wg_pairs, wg_rows;  // weight groups: row pairs >= wg_pairs use weight rows / bias offset by wg_rows (else wg_pairs = INT_MAX)
  int kb_main;  // k-blocks of the filter-tap walk over `x`; blocks [kb_main, kblocks) read the second 1x1 term `x2` (amap2)
  int conv, cblocks, S, pad, stride;  // stride 2: the activation map walks the input with element strides (1, 2, 2, 1)
  int Wb, Hb, Nb, tiles_w, tiles_h;
  FastDiv d_n_units, d_m_pairs, d_cblocks, d_S, d_tiles_w, d_tiles_h, d_ppi;
  int m_tiles, m_pairs, n_units, num_units;  // unit -> (n_unit fastest, m_pair, split)
  int stages, npb, slot_bytes, off16;        // shared-memory ring depth; panel slots and their layout
  // epilogue
  int M, pix_per_img, lde;
  int has_res32, has_res16, has_y32, has_y16, act;
  int emb_uniform;                           // every 128-row tile lies inside one image: its embedding row rides in the bias vector
  int partial;                               // split-K: fp32 partials [split][M][N] through a 3-D map (ragged M clips per split)
  float alpha;
  const float* bias;
  const bf16* emb;
  float2* stats;
  int stats_ld;
  const void* w;              // weights, for the L2 prefetch ahead of pdl_wait()
  unsigned long long w_bytes;  // (capped) bytes to prefetch, w_share of them per CTA
  unsigned w_share;
  unsigned long long* trace;  // debug: per-CTA %globaltimer stamps (mkd_debug_set_trace), else nullptr
};
// slot layout per CTA (16 x u64): 0 entry, 1 prologue done, 2 first TMA issued, 3 first full barrier (MMA warp), 4 unit-0 MMAs
// issued, 5 unit-0 accumulator ready (epilogue), 6 first panel computed, 7 first store issued, 8 last store issued, 9 stores
// drained, 10 last panel computed, 11 exit, 12 last unit's MMAs issued, 13 last TMA issued
// (compiled in only with -DMKD_ENABLE_TRACE, i.e. `MKD_TRACE=1 python -m makeupdiffuse_b200.build --force`)
#ifdef MKD_ENABLE_TRACE
#define PAIR_TRACE(slot)                                                  \
  do {                                                                    \
    if (p.trace) p.trace[(size_t)blockIdx.x * 16 + (slot)] = gtimer();    \
  } while (0)
// per-role progress words in shared memory, printed by a wait that times out
#define PAIR_PROG(i, v) do { prog[i] = (v); } while (0)
// detail region (after the 148 x 16 stamps): per CTA 16 panels x 8 clock64 stamps of epilogue warp 4 / 8 lane 0
#define PAIR_DETAIL(k)                                                                                       \
  do {                                                                                                       \
    if (p.trace && lane == 0 && quad == 0 && g < 16) p.trace[148 * 16 + ((size_t)blockIdx.x * 16 + g) * 8 + (k)] = clock64(); \
  } while (0)
#define PAIR_DETAIL_DEP(k, reg)                                                                              \
  do {                                                                                                       \
    if (p.trace && lane == 0 && quad == 0 && g < 16) {                                                       \
      long long c_;                                                                                          \
      asm volatile("mov.u64 %0, %%clock64;\n" : "=l"(c_) : "r"(reg));                                        \
      p.trace[148 * 16 + ((size_t)blockIdx.x * 16 + g) * 8 + (k)] = c_;                                      \
    }                                                                                                        \
  } while (0)
#else
#define PAIR_TRACE(slot) do { } while (0)
#define PAIR_PROG(i, v) do { } while (0)
#define PAIR_DETAIL(k) do { } while (0)
#define PAIR_DETAIL_DEP(k, reg) do { } while (0)
#endif

// ---- PTX wrappers -------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
  return t;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
// A wait that cannot hang the GPU: a protocol error (a barrier that is never completed) traps after a few seconds instead
// of spinning until the driver's watchdog — or the box's time limit — kills the process.  `site` names the waiting role;
// upstream roles time out first (2 s + 0.5 s x site), so the message names the wait closest to the cause.
enum { SITE_PROD_EMPTY = 0, SITE_MMA_FULL = 1, SITE_MMA_TMEM_EMPTY = 2, SITE_EPI_TMEM_FULL = 3, SITE_EPI_RES = 4, SITE_EPI_PREV = 4, SITE_DMA_COMPUTED = 5 };
__device__ int g_pair_shape[8];  // last launch: M, n cols, kblocks, num_units, stages, npb, operand flags, splits
__device__ __noinline__ void mbar_timeout(uint32_t bar, uint32_t parity, int site, int a, int b, volatile int* prog) {
  if ((threadIdx.x & 31) != 0 && site != SITE_DMA_COMPUTED) return;
  if (prog) printf("gemm_pair: block %d progress prod %d mma %d dma %d wg0 %d wg1 %d\n", (int)blockIdx.x, prog[0], prog[1], prog[2], prog[3], prog[4]);
  printf("gemm_pair: mbarrier wait timed out: site %d block %d thread %d smem 0x%x parity %u state (%d, %d); launch M %d N %d kblocks %d "
         "units %d stages %d npb %d flags 0x%x kb_per_split %d grid %d\n",
         site, (int)blockIdx.x, (int)threadIdx.x, bar, parity, a, b, g_pair_shape[0], g_pair_shape[1], g_pair_shape[2], g_pair_shape[3],
         g_pair_shape[4], g_pair_shape[5], g_pair_shape[6], g_pair_shape[7], (int)gridDim.x);
  if (site == SITE_DMA_COMPUTED) __trap();  // (the last role to time out: every other one has reported by then)
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int site, int a = 0, int b = 0, volatile int* prog = nullptr) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done, spins = 0;
  unsigned long long t0 = 0;
  for (;;) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) break;
    if ((++spins & 255u) == 0) {
      const unsigned long long now = gtimer();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 2000000000ull + 500000000ull * (unsigned)site) { mbar_timeout(addr, parity, site, a, b, prog); t0 = now; }
    }
  }
}
// Shared-window addresses carry the CTA rank of a pair in bit 24: clearing it names the same offset in the leader.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
// 2-SM TMA loads: the box lands in the ISSUING CTA's shared memory, its bytes complete on the LEADER's mbarrier
__device__ __forceinline__ void tma_load_2d_2sm(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::
          "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n" ::
          "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// CTA-local TMA load (residual panels) and store (finished panels)
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::
          "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n" ::"l"(reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];\n" ::"l"(reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
// rendezvous only (no memory ordering): both CTAs of the pair are running
__device__ __forceinline__ void cluster_sync_relaxed() {
  __syncwarp();
  asm volatile("barrier.cluster.arrive.relaxed.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  __syncwarp();  // role branches diverge lanes: reconverge before the .aligned barrier
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
// one MMA for the pair: D[256 x N], A 128 rows from each CTA's shared memory, B N/2 rows from each (same offsets in both)
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// commit of the pair's MMAs issued so far, arriving on the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void tcgen05_commit_2sm(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
// the registers of an asynchronous tcgen05.ld carry no dependence on tcgen05.wait::ld: this empty asm "rewrites" them
// after the wait (volatile asms keep their order), so that no use can be scheduled ahead of it
__device__ __forceinline__ void reg_fence32(uint32_t (&r)[32]) {
  asm volatile(""
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                 "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 lds128u(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// Blackwell packed fp32 pipe: two IEEE fp32 results per issue slot (a three-register FFMA occupies the FMA pipe of its
// scheduler for two cycles; the epilogue warps are bound by exactly that)
__device__ __forceinline__ uint64_t pk(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};\n" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ uint64_t pku(uint32_t a, uint32_t b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};\n" : "=l"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ void upk(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;\n" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;\n" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;\n" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;\n" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// erf-GELU of two values at once, same approximation as gelu_erf_fast (Abramowitz-Stegun 7.1.26): returns
// 0.5 x (1 + erf(x / sqrt 2)) for both lanes.  ~9 packed + 4 MUFU + 4 ALU instructions per pair.
__device__ __forceinline__ uint64_t gelu_erf_fast2(uint64_t x2) {
  float x0, x1;
  upk(x2, x0, x1);
  const uint64_t z2 = mul2(pk(fabsf(x0), fabsf(x1)), pk(0.70710678118654752f, 0.70710678118654752f));
  const uint64_t den2 = fma2(z2, pk(0.3275911f, 0.3275911f), pk(1.0f, 1.0f));
  const uint64_t arg2 = mul2(mul2(z2, pk(-1.4426950408889634f, -1.4426950408889634f)), z2);
  float d0, d1, a0, a1, t0, t1, e0, e1;
  upk(den2, d0, d1);
  upk(arg2, a0, a1);
  asm("rcp.approx.ftz.f32 %0, %1;\n" : "=f"(t0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;\n" : "=f"(t1) : "f"(d1));
  asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(e0) : "f"(a0));
  asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(e1) : "f"(a1));
  const uint64_t t2 = pk(t0, t1);
  // negated polynomial: q = -(a1 + t (a2 + t (a3 + t (a4 + t a5)))), so erf|x| = 1 + (q t) e
  uint64_t q2 = fma2(t2, pk(-1.061405429f, -1.061405429f), pk(1.453152027f, 1.453152027f));
  q2 = fma2(t2, q2, pk(-1.421413741f, -1.421413741f));
  q2 = fma2(t2, q2, pk(0.284496736f, 0.284496736f));
  q2 = fma2(t2, q2, pk(-0.254829592f, -0.254829592f));
  const uint64_t erf2 = fma2(mul2(q2, t2), pk(e0, e1), pk(1.0f, 1.0f));  // erf(|x| / sqrt 2)
  float r0, r1;
  upk(erf2, r0, r1);
  const uint64_t h2 = mul2(x2, pk(0.5f, 0.5f));
  return fma2(h2, pk(copysignf(r0, x0), copysignf(r1, x1)), h2);
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format): rows of 128 B, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor: D fp32, A / B bf16, both K-major
__host__ __device__ constexpr uint32_t make_idesc(int n, int m) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

struct Unit {
  int n_unit, m_pair, split, kb0, kb1;
};
__device__ __forceinline__ Unit decode_unit(const PairP& p, int u) {
  Unit t;
  const int rest = fdiv(u, p.d_n_units);
  t.n_unit = u - rest * p.n_units;
  t.split = fdiv(rest, p.d_m_pairs);
  t.m_pair = rest - t.split * p.m_pairs;
  t.kb0 = t.split * p.kb_per_split;
  t.kb1 = min(p.kblocks, t.kb0 + p.kb_per_split);
  return t;
}

enum { MODE_PLAIN = 0, MODE_GEGLU = 1 };

template <int UN, int NSUB, int MODE, int STATS>
__global__ void __launch_bounds__(THREADS, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap amap2,
                 const __grid_constant__ CUtensorMap bmap, const __grid_constant__ CUtensorMap r32map, const __grid_constant__ CUtensorMap r16map,
                 const __grid_constant__ CUtensorMap y32map, const __grid_constant__ CUtensorMap y16map, const PairP p) {
  constexpr int BH_BYTES = (UN / 2) * BK * 2;            // one CTA's half of a weight tile per k-block
  constexpr int STAGE_BYTES = A_BYTES + NSUB * BH_BYTES;
  constexpr int ACC_COLS = NSUB * UN;
  constexpr int TS = 2 * ACC_COLS <= TMEM_COLS ? 2 : 1;  // accumulator buffers in TMEM
  constexpr int PPS = UN / PANEL;                        // panels per sub-tile (plain)
  constexpr int PPU = MODE == MODE_GEGLU ? UN / 2 / PANEL : NSUB * PPS;  // panels per unit
  static_assert(MODE == MODE_PLAIN || (UN == 256 && NSUB == 1), "GEGLU: one 256-wide tile, [128 value | 128 gate]");
  static_assert(UN % 32 == 0 && UN <= 256 && ACC_COLS <= TMEM_COLS, "tile shape");

  const uint32_t rank = cluster_ctarank();
  const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* slots = smem + p.stages * STAGE_BYTES;
  float2* scratch = reinterpret_cast<float2*>(slots + p.npb * p.slot_bytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(scratch) + (STATS ? SCRATCH_BYTES : 0));
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tmem_full_bar = empty_bar + MAX_STAGES;  // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2]
  uint64_t* res_full_bar = tmem_empty_bar + 2;       // [MAX_SLOTS] slot is free and its residual (if any) has landed
  uint64_t* computed_bar = res_full_bar + MAX_SLOTS; // [MAX_SLOTS] the panel in the slot is ready to be stored
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(computed_bar + MAX_SLOTS);
  [[maybe_unused]] volatile int* prog = reinterpret_cast<volatile int*>(tmem_slot + 4);  // per-role progress (trace builds), printed on a timeout
  if (threadIdx.x < 8) prog[threadIdx.x] = -1;
  float* biasv = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(full_bar) + 512);  // [2 warpgroups][2][BIASV_FLOATS]
  // kernel parameters live in constant memory that is cold at every launch: touch every line now, so that the misses
  // (~1 us when taken one after the other in front of the first TMA) overlap the prologue's barriers instead
  asm volatile("" ::"r"(p.kblocks), "r"(p.Wb), "r"(p.num_units), "r"(p.npb), "r"(p.M), "r"(p.has_y32), "f"(p.alpha), "l"(p.bias),
               "l"(p.stats), "r"(p.stats_ld));

  if (threadIdx.x == 0) PAIR_TRACE(0);
  if (warp == 3) l2_prefetch_share(p.w, p.w_bytes, p.w_share, lane);
  if (blockIdx.x == 0 && threadIdx.x == 96) {  // (idle warp 3) what to print if a wait times out
    g_pair_shape[0] = p.M; g_pair_shape[1] = p.n_units * ACC_COLS; g_pair_shape[2] = p.kblocks; g_pair_shape[3] = p.num_units;
    g_pair_shape[4] = p.stages; g_pair_shape[5] = p.npb;
    g_pair_shape[6] = p.has_res32 | p.has_res16 << 1 | p.has_y32 << 2 | p.has_y16 << 3 | p.partial << 4 | p.conv << 5 | STATS << 6 | MODE << 7 | (p.emb != nullptr) << 8;
    g_pair_shape[7] = p.kb_per_split;
  }
  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&amap)) : "memory");
      asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&bmap)) : "memory");
      if (p.kb_main < p.kblocks) asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&amap2)) : "memory");
    }
    // one barrier per lane: full (count 2: both producers arrive on the leader's copy, both CTAs' bytes land there),
    // empty / tmem_full (1: multicast commit), tmem_empty (every epilogue warp of both CTAs, leader's copy),
    // res_full (1: the DMA thread), computed (4: the warps of the warpgroup that owns the panel)
    if (lane < MAX_STAGES) {
      mbar_init(full_bar + lane, 2);
      mbar_init(empty_bar + lane, 1);
      mbar_init(res_full_bar + lane, 1);
      mbar_init(computed_bar + lane, 4);
    }
    if (lane < 2) {
      mbar_init(tmem_full_bar + lane, 1);
      mbar_init(tmem_empty_bar + lane, 2 * EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  cluster_sync_relaxed();  // both CTAs are resident before the paired TMEM allocation
  if (threadIdx.x == 0) PAIR_TRACE(14);
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::);
  }
  tcgen05_fence_before();
  cluster_sync_all();  // the peer's barriers exist before anything arrives on them
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // from here on global memory written by the previous kernel is read and its outputs may be overwritten
  if (threadIdx.x == 0) PAIR_TRACE(1);

  if (warp == 0) {
    // ===== TMA producer (warp-uniform loop, one elected lane issues) =====
    int s = 0;
    uint32_t ph = 0;
#pragma unroll 1
    for (int u = pair_id; u < p.num_units; u += num_pairs) {
      const Unit t = decode_unit(p, u);
      const int m_tile = 2 * t.m_pair + (int)rank;
      int w0 = 0, h0 = 0, n0 = 0;
      if (p.conv) {
        const int mw = fdiv(m_tile, p.d_tiles_w), tw = m_tile - mw * p.tiles_w;
        const int tn = fdiv(mw, p.d_tiles_h), th = mw - tn * p.tiles_h;
        w0 = tw * p.Wb * p.stride;  // input coordinate of the tile's first output pixel (before the filter-tap offset)
        h0 = th * p.Hb * p.stride;
        n0 = tn * p.Nb;
      }
      const int m0 = m_tile * BM;
      // filter-tap walk (r, sx, cb) kept incrementally: no integer divisions in the k loop
      const int tap0 = fdiv(t.kb0, p.d_cblocks);
      int cb = t.kb0 - tap0 * p.cblocks, r = fdiv(tap0, p.d_S);
      int sx = tap0 - r * p.S;
      // this CTA's half of sub-tile 0's weight rows (second weight group: the same rows K further down)
      const int nb = t.n_unit * (NSUB * UN) + (int)rank * (UN / 2) + (t.m_pair >= p.wg_pairs ? p.wg_rows : 0);
#pragma unroll 1
      for (int kb = t.kb0; kb < t.kb1; ++kb) {
        if (lane == 0) PAIR_PROG(0, u * 1000 + kb);
        mbar_wait(empty_bar + s, ph ^ 1, SITE_PROD_EMPTY, u, kb, prog);  // (own copy) the MMAs that read this slot have retired
        if (elect_one()) {
          if (u == pair_id && kb == t.kb0) PAIR_TRACE(15);
          unsigned char* sa = smem + s * STAGE_BYTES;
          if (rank == 0) mbar_expect_tx(full_bar + s, 2 * STAGE_BYTES);
          else mbar_arrive_cluster(full_bar + s, 0);
          if (kb >= p.kb_main) {  // second term: the 1x1 filter over x2 (same pixels as the output tile, no tap offset)
            if (p.conv) tma_load_4d_2sm(&amap2, full_bar + s, sa, (kb - p.kb_main) * BK, w0, h0, n0);
            else tma_load_2d_2sm(&amap2, full_bar + s, sa, (kb - p.kb_main) * BK, m0);
          } else if (p.conv) tma_load_4d_2sm(&amap, full_bar + s, sa, cb * BK, w0 + sx - p.pad, h0 + r - p.pad, n0);
          else tma_load_2d_2sm(&amap, full_bar + s, sa, kb * BK, m0);
#pragma unroll
          for (int sub = 0; sub < NSUB; ++sub)
            tma_load_2d_2sm(&bmap, full_bar + s, sa + A_BYTES + sub * BH_BYTES, kb * BK, nb + sub * UN);
          if (u == pair_id && kb == t.kb0) PAIR_TRACE(2);
        }
        __syncwarp();
        if (++cb == p.cblocks) {
          cb = 0;
          if (++sx == p.S) {
            sx = 0;
            ++r;
          }
        }
        if (++s == p.stages) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: leader CTA only =====
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc(UN, 2 * BM);
      int s = 0, j = 0;
      uint32_t ph = 0;
#pragma unroll 1
      for (int u = pair_id; u < p.num_units; u += num_pairs, ++j) {
        const Unit t = decode_unit(p, u);
        const int ts = j % TS, use = j / TS;
        mbar_wait(tmem_empty_bar + ts, (use & 1) ^ 1, SITE_MMA_TMEM_EMPTY, u, j, prog);  // both CTAs' epilogue warps have drained this accumulator buffer
        tcgen05_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(ts * ACC_COLS);
#pragma unroll 1
        for (int kb = t.kb0; kb < t.kb1; ++kb) {
          if (lane == 0) PAIR_PROG(1, u * 1000 + kb);
          mbar_wait(full_bar + s, ph, SITE_MMA_FULL, u, kb, prog);
          tcgen05_fence_after();
          if (elect_one()) {
            if (j == 0 && kb == t.kb0) PAIR_TRACE(3);
            const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
            const uint64_t adesc = make_smem_desc(sa);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
#pragma unroll
              for (int sub = 0; sub < NSUB; ++sub) {
                const uint64_t bdesc = make_smem_desc(sa + A_BYTES + sub * BH_BYTES);
                // advance K inside the 128-byte swizzle atom: +32 bytes per UMMA_K (encoded >> 4)
                umma_bf16_2sm(tacc + (uint32_t)(sub * UN), adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                              (kb > t.kb0 || k) ? 1u : 0u);
              }
            }
            tcgen05_commit_2sm(empty_bar + s, (uint16_t)3);  // frees the slot in both CTAs once these MMAs retire
            if (kb == t.kb1 - 1) {
              tcgen05_commit_2sm(tmem_full_bar + ts, (uint16_t)3);
              if (j == 0) PAIR_TRACE(4);
            }
          }
          __syncwarp();
          if (++s == p.stages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 2) {
    // ===== epilogue DMA thread: residual panel loads, finished panel stores, slot recycling =====
    if (lane == 0 && pair_id < p.num_units) {
      const int my_units = (p.num_units - pair_id + num_pairs - 1) / num_pairs;
      const int total = my_units * PPU;
      const int D = p.npb >= 3 ? p.npb - 2 : p.npb - 1;  // the store of panel j is issued D panels after its slot was armed
      const bool any_res = p.has_res32 || p.has_res16;
      const uint32_t res_bytes = (p.has_res32 ? P32_BYTES : 0) + (p.has_res16 ? P16_BYTES : 0);
      constexpr int OUTW = MODE == MODE_GEGLU ? UN / 2 : NSUB * UN;  // output columns of a unit (panel q starts at column 32 q)
      // load / store cursors: unit, panel within the unit, slot; the unit's coordinates are decoded once per unit
      struct Cursor {
        int u, q, slot, row, col, valid;
      };
      auto enter_unit = [&](Cursor& c) {
        const Unit t = decode_unit(p, c.u);
        const int m_tile = 2 * t.m_pair + (int)rank;
        c.valid = m_tile < p.m_tiles;
        c.row = m_tile * BM;
        c.col = t.n_unit * OUTW;
        return t.split;
      };
      Cursor ld = {pair_id, 0, 0, 0, 0, 0}, st = {pair_id, 0, 0, 0, 0, 0};
      enter_unit(ld);
      int st_split = enter_unit(st);
      uint32_t sph = 0;
#pragma unroll 1
      for (int i = 0; i < total + D; ++i) {
        PAIR_PROG(2, i * 10);
        if (i < total) {
          // the slot's previous panel (i - npb) was stored npb - D iterations ago: at most npb - D - 1 younger groups
          if (i >= p.npb) {
            if (D == p.npb - 2) bulk_wait_read<1>();
            else bulk_wait_read<0>();
          }
          PAIR_PROG(2, i * 10 + 1);
          unsigned char* slot = slots + ld.slot * p.slot_bytes;
          if (any_res && ld.valid) {
            mbar_expect_tx(res_full_bar + ld.slot, res_bytes);
            if (p.has_res32) tma_load_2d(&r32map, res_full_bar + ld.slot, slot, ld.col + ld.q * PANEL, ld.row);
            if (p.has_res16) tma_load_2d(&r16map, res_full_bar + ld.slot, slot + p.off16, ld.col + ld.q * PANEL, ld.row);
          } else {
            mbar_arrive(res_full_bar + ld.slot);  // nothing to load: the slot is simply free
          }
          if (++ld.slot == p.npb) ld.slot = 0;
          if (++ld.q == PPU) {
            ld.q = 0;
            ld.u += num_pairs;
            if (ld.u < p.num_units) enter_unit(ld);
          }
        }
        if (i >= D) {
          PAIR_PROG(2, i * 10 + 2);
          unsigned char* slot = slots + st.slot * p.slot_bytes;
          mbar_wait(computed_bar + st.slot, sph, SITE_DMA_COMPUTED, i, st.slot, prog);
          if (st.valid) {
            if (p.partial) tma_store_3d(&y32map, slot, st.col + st.q * PANEL, st.row, st_split);
            else if (p.has_y32) tma_store_2d(&y32map, slot, st.col + st.q * PANEL, st.row);
            if (p.has_y16) tma_store_2d(&y16map, slot + p.off16, st.col + st.q * PANEL, st.row);
          }
          bulk_commit();
          PAIR_PROG(2, i * 10 + 3);
          if (i == D) PAIR_TRACE(7);
          PAIR_TRACE(8);
          if (++st.slot == p.npb) {
            st.slot = 0;
            sph ^= 1;
          }
          if (++st.q == PPU) {
            st.q = 0;
            st.u += num_pairs;
            if (st.u < p.num_units) st_split = enter_unit(st);
          }
        }
      }
      bulk_wait_read<0>();  // shared memory must outlive the stores' reads; their global writes complete with the grid
      PAIR_TRACE(9);
    }
  } else if (warp >= EPI_WARP0) {
    // ===== epilogue warps =====
    const int wg = (warp - EPI_WARP0) >> 2, quad = warp & 3;
    const int row = quad * 32 + lane;                      // accumulator row of this thread = TMEM lane
    const uint32_t lane_bits = (uint32_t)(quad * 32) << 16;
    const uint32_t sw128 = (uint32_t)(row & 7) << 4;       // SWIZZLE_128B: 16-byte chunk index ^= row % 8
    const uint32_t sw64 = (uint32_t)((row >> 1) & 3) << 4; // SWIZZLE_64B:  16-byte chunk index ^= (row / 2) % 4
    const uint32_t slots_u32 = smem_u32(slots);
    // slot_order note.  A parity wait is only sound when the waiter can be at most one phase ahead of the barrier.  With an
    // odd number of slots consecutive uses of a slot belong to ALTERNATE warpgroups, and nothing else keeps one warpgroup
    // from running two panels ahead of the other (its residual panels may simply land first): waiting for use n of
    // res_full[b] while use n - 1 has not completed passes at once (same parity as "use n done") and the warpgroup
    // overwrites a slot the other one has not even started.  So before a slot's use n a warpgroup first waits for the
    // slot's previous panel to have been computed (computed[b], use n - 1), which implies res_full[b] finished use n - 1.
    int g = 0, b = 0, j = 0;
    uint32_t rph = 0;
    [[maybe_unused]] int lp = 0;  // statistics scratch buffer parity (panels this warpgroup has reduced)
    // ---- per-unit column vector (bias [+ embedding row]) in shared memory, one unit ahead.  A global load issued while
    // the TMA ring saturates the L2 -> SM path takes ~1 us: requested at the top of a panel it put that latency on every
    // panel (the whole epilogue ran at 1.2 us per panel).  Each warpgroup keeps its own copy (no cross-group barrier).
    constexpr int ACCW = MODE == MODE_GEGLU ? UN : NSUB * UN;  // columns (weight rows) of a unit
    static_assert(ACCW <= BIASV_FLOATS && ACCW / 4 <= 128, "bias vector");
    const int wt = (warp - EPI_WARP0 - 4 * wg) * 32 + lane;     // thread index within the warpgroup
    float* bias_wg = biasv + wg * 2 * BIASV_FLOATS;
    auto load_colvec = [&](int u_) -> float4 {
      float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (wt < ACCW / 4 && u_ < p.num_units) {
        const Unit t_ = decode_unit(p, u_);
        const int col = t_.n_unit * ACCW + wt * 4;
        if (p.bias) v4 = __ldg(reinterpret_cast<const float4*>(p.bias + col + (t_.m_pair >= p.wg_pairs ? p.wg_rows : 0)));
        if (MODE == MODE_PLAIN && p.emb && p.emb_uniform) {
          const int mt = min(2 * t_.m_pair + (int)rank, p.m_tiles - 1);
          const uint2 e = __ldg(reinterpret_cast<const uint2*>(p.emb + (int64_t)fdiv(mt * BM, p.d_ppi) * p.lde + col));
          v4.x += __uint_as_float(e.x << 16);
          v4.y += __uint_as_float(e.x & 0xFFFF0000u);
          v4.z += __uint_as_float(e.y << 16);
          v4.w += __uint_as_float(e.y & 0xFFFF0000u);
        }
        if (MODE == MODE_PLAIN) {  // the epilogue computes acc * alpha + colvec
          v4.x *= p.alpha; v4.y *= p.alpha; v4.z *= p.alpha; v4.w *= p.alpha;
        }
      }
      return v4;
    };
    float4 colvec_next = load_colvec(pair_id);
    const bool emb_gather = MODE == MODE_PLAIN && p.emb && !p.emb_uniform;  // tiles spanning several images: per-row gather
#pragma unroll 1
    for (int u = pair_id; u < p.num_units; u += num_pairs, ++j) {
      const Unit t = decode_unit(p, u);
      const int ts = j % TS, use = j / TS;
      const int m_tile = 2 * t.m_pair + (int)rank;
      const bool valid = m_tile < p.m_tiles;
      const int m = m_tile * BM + row;
      const int q_last = (((g + PPU - 1) & 1) == wg) ? PPU - 1 : PPU - 2;  // this warpgroup's last panel of the unit
      // this unit's column vector -> shared memory (buffer j & 1 was last read two units ago); next unit's requested
      float* bias_s = bias_wg + (j & 1) * BIASV_FLOATS;
      if (wt < ACCW / 4) *reinterpret_cast<float4*>(bias_s + wt * 4) = colvec_next;
      asm volatile("bar.sync %0, 128;\n" ::"r"(1 + wg) : "memory");
      colvec_next = load_colvec(u + num_pairs);
      const uint32_t bias_u32 = smem_u32(bias_s);
      bool waited = false;
#pragma unroll 1
      for (int q = 0; q < PPU; ++q, ++g) {
        if ((g & 1) == wg) {
          if (quad == 0 && lane == 0) PAIR_PROG(3 + wg, g * 10);
          if (!waited) {
            mbar_wait(tmem_full_bar + ts, use & 1, SITE_EPI_TMEM_FULL, u, j, prog);
            tcgen05_fence_after();
            waited = true;
            if (j == 0 && warp == EPI_WARP0 && lane == 0) PAIR_TRACE(5);
          }
          const uint32_t s32 = slots_u32 + (uint32_t)(b * p.slot_bytes) + (uint32_t)(row * 128);
          const uint32_t s16 = slots_u32 + (uint32_t)(b * p.slot_bytes + p.off16) + (uint32_t)(row * 64);
          if constexpr (MODE == MODE_GEGLU) {
            uint32_t av[32], ag[32];
            const uint32_t taddr = tmem_base + lane_bits + (uint32_t)(ts * ACC_COLS + q * PANEL);
            tmem_ld32_nowait(taddr, av);
            tmem_ld32_nowait(taddr + UN / 2, ag);
            const uint32_t bvp = bias_u32 + (uint32_t)(q * PANEL) * 4, bgp = bvp + (UN / 2) * 4;  // value / gate bias of the panel
            if (g >= p.npb) mbar_wait(computed_bar + b, rph ^ 1, SITE_EPI_PREV, g, b, prog);  // (see slot_order note)
            mbar_wait(res_full_bar + b, rph, SITE_EPI_RES, g, b, prog);  // the slot is free
            tmem_ld_wait();
            reg_fence32(av);
            reg_fence32(ag);
            if (q == q_last) {  // every TMEM read of this unit by this warp is done
              tcgen05_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive_cluster(tmem_empty_bar + ts, 0);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {  // 8 outputs per 16-byte chunk, computed as 4 packed pairs
              uint32_t o[4];
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int i = 2 * c + h;
                // bias of this 4-column piece (uniform address: shared-memory broadcast)
                const float4 bv = lds128(bvp + i * 16), bg = lds128(bgp + i * 16);
                const uint64_t v01 = add2(pku(av[4 * i + 0], av[4 * i + 1]), pk(bv.x, bv.y));
                const uint64_t v23 = add2(pku(av[4 * i + 2], av[4 * i + 3]), pk(bv.z, bv.w));
                const uint64_t g01 = gelu_erf_fast2(add2(pku(ag[4 * i + 0], ag[4 * i + 1]), pk(bg.x, bg.y)));
                const uint64_t g23 = gelu_erf_fast2(add2(pku(ag[4 * i + 2], ag[4 * i + 3]), pk(bg.z, bg.w)));
                float r0, r1, r2, r3;
                upk(mul2(v01, g01), r0, r1);
                upk(mul2(v23, g23), r2, r3);
                o[2 * h + 0] = pack_bf16(r0, r1);
                o[2 * h + 1] = pack_bf16(r2, r3);
              }
              sts128u(s16 + (((uint32_t)c << 4) ^ sw64), o[0], o[1], o[2], o[3]);
            }
          } else {
            const int sub = q / PPS, pp = q - sub * PPS;
            const int gcol = t.n_unit * (NSUB * UN) + sub * UN + pp * PANEL;
            uint32_t acc[32];
            PAIR_DETAIL(0);
            tmem_ld32_nowait(tmem_base + lane_bits + (uint32_t)(ts * ACC_COLS + sub * UN + pp * PANEL), acc);
            uint4 eb[4];
            if (emb_gather) {
              const uint4* er = reinterpret_cast<const uint4*>(p.emb + (int64_t)fdiv(min(m, p.M - 1), p.d_ppi) * p.lde + gcol);
#pragma unroll
              for (int c = 0; c < 4; ++c) eb[c] = __ldg(er + c);
            }
            const uint32_t bcol = bias_u32 + (uint32_t)(sub * UN + pp * PANEL) * 4;
            if (g >= p.npb) mbar_wait(computed_bar + b, rph ^ 1, SITE_EPI_PREV, g, b, prog);  // (see slot_order note)
            mbar_wait(res_full_bar + b, rph, SITE_EPI_RES, g, b, prog);  // the slot is free and its residual panel (if any) has landed
            PAIR_DETAIL(1);
            uint4 rb[4];
            if (p.has_res16) {
#pragma unroll
              for (int c = 0; c < 4; ++c) rb[c] = lds128u(s16 + (((uint32_t)c << 4) ^ sw64));
            }
            tmem_ld_wait();
            reg_fence32(acc);
            PAIR_DETAIL_DEP(2, acc[31]);
            if (g == 0 && warp == EPI_WARP0 && lane == 0) PAIR_TRACE(12);
            if (q == q_last) {
              tcgen05_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive_cluster(tmem_empty_bar + ts, 0);
            }
            // The math runs in phases over all 32 columns: every phase is one straight-line block of independent
            // instructions.  (One loop over 4-column pieces with the operand tests inside compiled to a chain of short
            // basic blocks: ~350 dependent instructions at 4-5 clk each = 1400 clk per panel, the whole epilogue's pace.)
            // v = acc * alpha + colvec, colvec = (bias + embedding row) * alpha prepared per unit.
            float v[32];
            {
              float4 bb[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) bb[i] = lds128(bcol + i * 16);  // uniform address: broadcast
              const uint64_t alpha2 = pk(p.alpha, p.alpha);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                upk(fma2(pku(acc[4 * i + 0], acc[4 * i + 1]), alpha2, pk(bb[i].x, bb[i].y)), v[4 * i + 0], v[4 * i + 1]);
                upk(fma2(pku(acc[4 * i + 2], acc[4 * i + 3]), alpha2, pk(bb[i].z, bb[i].w)), v[4 * i + 2], v[4 * i + 3]);
              }
            }
            if (emb_gather) {
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const uint32_t e[4] = {eb[c].x, eb[c].y, eb[c].z, eb[c].w};
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                  v[8 * c + 2 * h + 0] = fmaf(__uint_as_float(e[h] << 16), p.alpha, v[8 * c + 2 * h + 0]);
                  v[8 * c + 2 * h + 1] = fmaf(__uint_as_float(e[h] & 0xFFFF0000u), p.alpha, v[8 * c + 2 * h + 1]);
                }
              }
            }
            if (p.has_res32) {
              float4 r[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) r[i] = lds128(s32 + (((uint32_t)i << 4) ^ sw128));
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                upk(add2(pk(v[4 * i + 0], v[4 * i + 1]), pk(r[i].x, r[i].y)), v[4 * i + 0], v[4 * i + 1]);
                upk(add2(pk(v[4 * i + 2], v[4 * i + 3]), pk(r[i].z, r[i].w)), v[4 * i + 2], v[4 * i + 3]);
              }
            }
            if (p.has_res16) {
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const uint32_t e[4] = {rb[c].x, rb[c].y, rb[c].z, rb[c].w};
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                  v[8 * c + 2 * h + 0] += __uint_as_float(e[h] << 16);
                  v[8 * c + 2 * h + 1] += __uint_as_float(e[h] & 0xFFFF0000u);
                }
              }
            }
            if (p.act == MKD_ACT_SILU) {
#pragma unroll
              for (int c = 0; c < 32; ++c) v[c] = silu_f(v[c]);
            }
            if (p.has_y32) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                sts128(s32 + (((uint32_t)i << 4) ^ sw128), v[4 * i + 0], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            }
            if (p.has_y16) {
#pragma unroll
              for (int c = 0; c < 4; ++c)
                sts128u(s16 + (((uint32_t)c << 4) ^ sw64), pack_bf16(v[8 * c + 0], v[8 * c + 1]), pack_bf16(v[8 * c + 2], v[8 * c + 3]),
                        pack_bf16(v[8 * c + 4], v[8 * c + 5]), pack_bf16(v[8 * c + 6], v[8 * c + 7]));
            }
            if constexpr (STATS != 0) {
              // column (sum, sum of squares) over this warp's 32 rows by a register transpose-reduction: after the
              // step with offset `off` a lane keeps the half of its columns selected by bit `off` of its lane id, so
              // lane l ends with column l.  Fixed order: deterministic.
              float xq[32];
#pragma unroll
              for (int c = 0; c < 32; ++c) xq[c] = v[c] * v[c];
#pragma unroll
              for (int off = 16; off >= 1; off >>= 1) {
                const bool up = (lane & off) != 0;
#pragma unroll
                for (int c = 0; c < off; ++c) {
                  const float send_s = up ? v[c] : v[c + off], keep_s = up ? v[c + off] : v[c];
                  const float send_q = up ? xq[c] : xq[c + off], keep_q = up ? xq[c + off] : xq[c];
                  v[c] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, off);
                  xq[c] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, off);
                }
              }
              float2* scr = scratch + ((wg * 2 + (lp & 1)) * 4) * 32;
              scr[quad * 32 + lane] = make_float2(v[0], xq[0]);
              asm volatile("bar.sync %0, 128;\n" ::"r"(1 + wg) : "memory");
              if (quad == 0 && valid) {
                const float2 a0 = scr[lane], a1 = scr[32 + lane], a2 = scr[64 + lane], a3 = scr[96 + lane];
                p.stats[(int64_t)m_tile * p.stats_ld + gcol + lane] =
                    make_float2(((a0.x + a1.x) + a2.x) + a3.x, ((a0.y + a1.y) + a2.y) + a3.y);
              }
              ++lp;
            }
          }
          if (g == 0 && warp == EPI_WARP0 && lane == 0) PAIR_TRACE(13);
          PAIR_DETAIL(3);
          if (quad == 0 && lane == 0) PAIR_PROG(3 + wg, g * 10 + 5);
          fence_proxy_async();  // this thread's panel writes -> visible to the TMA store
          PAIR_DETAIL(4);
          __syncwarp();
          if (lane == 0) mbar_arrive(computed_bar + b);
          PAIR_DETAIL(5);
          if (quad == 0 && lane == 0) {
            if (g == 0) PAIR_TRACE(6);
            PAIR_TRACE(10);
          }
        }
        if (++b == p.npb) {
          b = 0;
          rph ^= 1;
        }
      }
    }
  }
  tcgen05_fence_before();
  cluster_sync_all();  // no CTA leaves while its peer can still arrive on its barriers or the pair's MMAs read its smem
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
  if (threadIdx.x == 0) PAIR_TRACE(11);
}

// ---- host side ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  if (!fn) {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult r;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) == cudaSuccess && r == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(q);
  }
  return fn;
}
int encode(CUtensorMap* map, CUtensorMapDataType dt, const void* base, int rank, const cuuint64_t* dims,
           const cuuint64_t* strides_bytes, const cuuint32_t* box, CUtensorMapSwizzle sw, int spatial_stride = 1) {
  EncodeFn fn = get_encode();
  MKD_REQUIRE(fn != nullptr, MKD_E_CUDA, "cuTensorMapEncodeTiled entry point not found (driver too old?)");
  // element (traversal) strides: a box dimension of b with stride e delivers ceil(b / e) elements, every e-th one
  cuuint32_t estr[5] = {1, (cuuint32_t)spatial_stride, (cuuint32_t)spatial_stride, 1, 1};
  CUresult r = fn(map, dt, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MKD_REQUIRE(r == CUDA_SUCCESS, MKD_E_CUDA, "gemm_pair: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
  return MKD_OK;
}
// [rows, cols] row-major matrix of fp32 / bf16 with row pitch ld (elements); box = 32 columns x 128 rows
int encode_panel_map(CUtensorMap* map, bool f32, const void* base, int64_t rows, int cols, int ld) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t str[1] = {(cuuint64_t)ld * (f32 ? 4 : 2)};
  cuuint32_t box[2] = {PANEL, BM};
  return encode(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, 2, dims, str, box,
                f32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
}

FastDiv make_fastdiv(int d) {
  FastDiv f;
  f.d = (uint32_t)d;
  f.m = d <= 1 ? 0u : (uint32_t)(((1ull << 32) + (uint32_t)d - 1) / (uint32_t)d);
  return f;
}

struct PairPlan {
  int un, nsub, geglu, stats;
  int conv, M, Ktot, P, Q, Wb, Hb, Nb, tiles_w, tiles_h, m_tiles;
  int splits, kb_per_split, stages, npb, slot_bytes, off16;
};

bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// Which descriptors the pair kernel takes, and how.  Stride 1 (1x1, 3x3 / pad 1) and the stride-2 Downsample convs
// (3x3, pad 1: input pixel 2 p + r - 1, read through a tensor map with element strides 2 — no im2col pass); upsample
// convs arrive here materialised by the caller.
bool plan(const mkd_conv_desc* d, PairPlan& pl, bool forced) {
  if (d->dtype != MKD_BF16 || d->upsample) return false;
  if (d->C % BK != 0 || d->R != d->S || (d->R != 1 && d->R != 3)) return false;
  if (d->stride == 1) {
    if (d->pad != d->R / 2 || d->pad_hi_extra) return false;
  } else {
    if (d->stride != 2 || d->R != 3 || d->H % 2 || d->W % 2 || d->pad != 1 || d->pad_hi_extra || d->W > 256) return false;
  }
  if (d->ldx % 8 || !aligned16(d->x) || !aligned16(d->w)) return false;
  if (d->bias && !aligned16(d->bias)) return false;
  if (d->y && (d->ldy % 8 || !aligned16(d->y))) return false;
  if (d->y32 && (d->ldy32 % 4 || !aligned16(d->y32))) return false;
  if (d->residual && (!aligned16(d->residual) || d->ldr % (d->residual_dtype == MKD_F32 ? 4 : 8))) return false;
  if (d->emb && (d->lde % 8 || !aligned16(d->emb))) return false;
  if (d->stats && (d->act != MKD_ACT_NONE || d->stats_ld < d->K || ((uintptr_t)d->stats & 7))) return false;
  // second 1x1 term (mkd_conv_desc.x2): extra k-blocks at the end of the walk, read through a second activation map
  if (d->x2 && (d->stride != 1 || d->C2 <= 0 || d->C2 % BK || d->ldx2 % 8 || d->ldx2 < d->C2 || !aligned16(d->x2) ||
                d->act == MKD_ACT_GEGLU))
    return false;
  if (d->wgroups > 2 || d->wgroups < 0) return false;
  pl.conv = d->R == 3;
  pl.P = d->H / d->stride;
  pl.Q = d->W / d->stride;
  pl.M = d->N * pl.P * pl.Q;  // output rows (stride 2: a quarter of the input pixels)
  pl.Ktot = d->R * d->S * d->C + (d->x2 ? d->C2 : 0);
  // ragged M: TMA zero-fills the rows it loads beyond M and clips the rows it stores; statistics need whole tiles
  if (d->stats && pl.M % BM != 0) return false;
  if (pl.conv) {
    if (!is_pow2(pl.Q) || !is_pow2(pl.P)) return false;
    pl.Wb = pl.Q < BM ? pl.Q : BM;
    pl.Hb = BM / pl.Wb < pl.P ? BM / pl.Wb : pl.P;
    pl.Nb = BM / (pl.Wb * pl.Hb);
    pl.tiles_w = pl.Q / pl.Wb;
    pl.tiles_h = pl.P / pl.Hb;
    pl.m_tiles = pl.tiles_w * pl.tiles_h * ((d->N + pl.Nb - 1) / pl.Nb);
  } else {
    pl.Wb = pl.Hb = pl.Nb = pl.tiles_w = pl.tiles_h = 1;
    pl.m_tiles = (pl.M + BM - 1) / BM;
  }
  const int m_pairs = (pl.m_tiles + 1) / 2, kblocks = pl.Ktot / BK;
  // weight groups: each part of the rows is a whole number of 256-row tile pairs
  if (d->wgroups == 2 && (pl.M % (4 * BM) != 0 || pl.m_tiles % 4 != 0 || d->N % 2 != 0)) return false;
  // the kernel divides by launch constants with umulhi(n, ceil(2^32 / d)), exact while n * d < 2^32
  {
    const uint64_t lim = 1ull << 32, units_max = (uint64_t)m_pairs * (d->K / 160 + 1) * 16 + 2 * num_sms();
    if (units_max * (uint64_t)(d->K / 160 + 1) >= lim || units_max * (uint64_t)m_pairs >= lim ||
        (d->emb && (uint64_t)(pl.M + BM) * (uint64_t)(pl.P * pl.Q) >= lim) ||  // (rows / pixels-per-image: embedding rows only) (uint64_t)kblocks * (uint64_t)(d->C / BK) >= lim ||
        (uint64_t)(2 * m_pairs) * (uint64_t)(pl.tiles_w * pl.tiles_h) >= lim)
      return false;
  }
  pl.stats = d->stats != nullptr;
  pl.splits = 1;
  if (d->act == MKD_ACT_GEGLU) {
    if (d->geglu_block != 128 || d->K % 256 || !d->y || d->y32 || d->residual || d->emb || d->alpha != 1.0f || d->stats) return false;
    pl.geglu = 1;
    pl.un = 256;
    pl.nsub = 1;
  } else {
    if (d->K % 160 != 0) return false;
    if (d->act != MKD_ACT_NONE && d->act != MKD_ACT_SILU) return false;
    pl.geglu = 0;
    pl.un = 160;
    // two N = 160 sub-tiles per A tile where that still leaves enough units for the 74 pairs (deep-K 32x32-level convs),
    // or where the unit count is so small that the work is split along K anyway (8x8 / 4x4 levels)
    const int units1 = m_pairs * (d->K / 160), units2 = d->K % 320 == 0 ? m_pairs * (d->K / 320) : 0;
    const bool can_split = d->workspace && !d->stats && kblocks >= 32;
    pl.nsub = 1;
    if (units2 >= 48 && kblocks >= 16) pl.nsub = 2;
    // (a single pair of row tiles — the 4x4 level at batch 16 — splits better as 8 narrow units x 9 than as 4 wide x 15:
    // the fp32 partial epilogue of a 320-wide unit costs more than its 12 k-blocks, and the reducer reads 9 partials, not 15)
    else if (units2 > 0 && units1 >= 16 && units1 < 37 && can_split && kblocks >= 64) pl.nsub = 2;
    const int units = pl.nsub == 2 ? units2 : units1;
    if (units < 37 && can_split) {
      int s = 74 / units;
      if (s > kblocks / 12) s = kblocks / 12;
      if (s > 16) s = 16;
      while (s > 1 && (size_t)s * pl.M * d->K * sizeof(float) > d->workspace_bytes) --s;
      if (s > 1) pl.splits = s;
    }
    // too few units to be worth a cluster launch: leave tiny problems to the single-CTA kernel (more, smaller tiles) —
    // except with weight groups, where declining means TWO single-CTA launches back to back: the k-block rate does not depend on
    // the tile size, so one pair launch over both networks' rows takes the time of one of them (the middle block's K = C GEMMs at
    // batch 16: 2 x 8 units)
    if (!forced && d->wgroups != 2 && units * pl.splits < 24) return false;
  }
  pl.kb_per_split = (kblocks + pl.splits - 1) / pl.splits;
  pl.splits = (kblocks + pl.kb_per_split - 1) / pl.kb_per_split;
  // GroupNorm tail (mkd_conv_desc.gn_y): rides in the split-K reducer, one block per (image, group), one vector of 8 channels per thread
  if (d->gn_y && (pl.splits <= 1 || pl.geglu || (int64_t)pl.P * pl.Q * (d->K / d->gn_groups / 8) > 512)) return false;
  // shared memory: ring stages + panel slots
  const bool partial = pl.splits > 1;
  const bool any32 = partial || d->y32 || (d->residual && d->residual_dtype == MKD_F32);
  const bool any16 = !partial && (d->y || (d->residual && d->residual_dtype == MKD_BF16));
  pl.slot_bytes = (any32 ? P32_BYTES : 0) + (any16 ? P16_BYTES : 0);
  pl.off16 = any32 ? P32_BYTES : 0;
  const int stage_bytes = A_BYTES + pl.nsub * (pl.un / 2) * BK * 2;
  const int budget = SMEM_LIMIT - 1024 - BAR_BYTES - (pl.stats ? SCRATCH_BYTES : 0);
  pl.stages = pl.kb_per_split >= 12 ? 4 : 3;
  if (pl.stages > pl.kb_per_split && pl.kb_per_split >= 2) pl.stages = pl.kb_per_split < 3 ? 2 : 3;
  pl.npb = (budget - pl.stages * stage_bytes) / pl.slot_bytes;
  if (pl.npb > MAX_SLOTS) pl.npb = MAX_SLOTS;
  while (pl.npb < 3 && pl.stages > 2) {  // (not reached with the shapes above; keeps the slot ring workable)
    --pl.stages;
    pl.npb = (budget - pl.stages * stage_bytes) / pl.slot_bytes;
  }
  if (pl.npb > MAX_SLOTS) pl.npb = MAX_SLOTS;
  if (pl.npb < 2) return false;
  // spare shared memory goes to ring stages
  while (pl.stages < 6 && pl.kb_per_split > pl.stages && budget - (pl.stages + 1) * stage_bytes - pl.npb * pl.slot_bytes >= 0) ++pl.stages;
  return true;
}

template <int UN, int NSUB, int MODE, int STATS>
int launch(const mkd_conv_desc* d, const PairPlan& pl, cudaStream_t stream) {
  auto kernel = gemm_pair_kernel<UN, NSUB, MODE, STATS>;
  constexpr int STAGE_BYTES = A_BYTES + NSUB * (UN / 2) * BK * 2;
  const size_t smem = 1024 + (size_t)pl.stages * STAGE_BYTES + (size_t)pl.npb * pl.slot_bytes + (STATS ? SCRATCH_BYTES : 0) + BAR_BYTES;
  MKD_REQUIRE(smem <= (size_t)SMEM_LIMIT, MKD_E_INVALID, "gemm_pair: shared memory plan %zu exceeds 227 KB", smem);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
    MKD_REQUIRE(e == cudaSuccess, MKD_E_CUDA, "gemm_pair: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  const bool partial = pl.splits > 1;
  CUtensorMap amap, amap2, bmap, r32map, r16map, y32map, y16map;
  int rc;
  if (pl.conv) {
    cuuint64_t dims[4] = {(cuuint64_t)d->C, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
    cuuint64_t str[3] = {(cuuint64_t)d->ldx * 2, (cuuint64_t)d->ldx * 2 * d->W, (cuuint64_t)d->ldx * 2 * d->W * d->H};
    cuuint32_t box[4] = {BK, (cuuint32_t)(pl.Wb * d->stride), (cuuint32_t)(pl.Hb * d->stride), (cuuint32_t)pl.Nb};
    rc = encode(&amap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, d->x, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, d->stride);
  } else {
    cuuint64_t dims[2] = {(cuuint64_t)d->C, (cuuint64_t)pl.M};
    cuuint64_t str[1] = {(cuuint64_t)d->ldx * 2};
    cuuint32_t box[2] = {BK, BM};
    rc = encode(&amap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, d->x, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
  }
  if (rc) return rc;
  amap2 = amap;  // (an unused map must still be a valid kernel parameter)
  if (d->x2) {   // same boxes over the second input: its pixels are the output's (stride 1), its channels the extra k-blocks
    if (pl.conv) {
      cuuint64_t dims[4] = {(cuuint64_t)d->C2, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
      cuuint64_t str[3] = {(cuuint64_t)d->ldx2 * 2, (cuuint64_t)d->ldx2 * 2 * d->W, (cuuint64_t)d->ldx2 * 2 * d->W * d->H};
      cuuint32_t box[4] = {BK, (cuuint32_t)pl.Wb, (cuuint32_t)pl.Hb, (cuuint32_t)pl.Nb};
      rc = encode(&amap2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, d->x2, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    } else {
      cuuint64_t dims[2] = {(cuuint64_t)d->C2, (cuuint64_t)pl.M};
      cuuint64_t str[1] = {(cuuint64_t)d->ldx2 * 2};
      cuuint32_t box[2] = {BK, BM};
      rc = encode(&amap2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, d->x2, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    }
    if (rc) return rc;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)pl.Ktot, (cuuint64_t)d->K * (d->wgroups == 2 ? 2 : 1)};
    cuuint64_t str[1] = {(cuuint64_t)pl.Ktot * 2};
    cuuint32_t box[2] = {BK, (cuuint32_t)(UN / 2)};
    rc = encode(&bmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, d->w, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  const int n_out = MODE == MODE_GEGLU ? d->K / 2 : d->K;
  PairP p = {};
  p.kblocks = pl.Ktot / BK;
  p.kb_main = d->R * d->S * d->C / BK;
  p.wg_pairs = d->wgroups == 2 ? (pl.m_tiles + 1) / 4 : 0x7fffffff;
  p.wg_rows = d->K;
  p.kb_per_split = pl.kb_per_split;
  p.conv = pl.conv;
  p.cblocks = d->C / BK;
  p.S = d->S;
  p.pad = d->pad;
  p.stride = d->stride;
  p.Wb = pl.Wb; p.Hb = pl.Hb; p.Nb = pl.Nb;
  p.tiles_w = pl.tiles_w; p.tiles_h = pl.tiles_h;
  p.m_tiles = pl.m_tiles;
  p.m_pairs = (pl.m_tiles + 1) / 2;
  p.n_units = d->K / (NSUB * UN);
  p.num_units = p.m_pairs * p.n_units * pl.splits;
  p.stages = pl.stages; p.npb = pl.npb; p.slot_bytes = pl.slot_bytes; p.off16 = pl.off16;
  p.d_n_units = make_fastdiv(p.n_units);
  p.d_m_pairs = make_fastdiv(p.m_pairs);
  p.d_cblocks = make_fastdiv(p.cblocks);
  p.d_S = make_fastdiv(p.S);
  p.d_tiles_w = make_fastdiv(p.tiles_w);
  p.d_tiles_h = make_fastdiv(p.tiles_h);
  p.d_ppi = make_fastdiv(d->emb ? pl.P * pl.Q : 1);

  p.M = pl.M;
  p.pix_per_img = pl.P * pl.Q;
  p.lde = d->lde;
  p.act = partial ? MKD_ACT_NONE : d->act;
  p.alpha = partial ? 1.0f : d->alpha;
  p.bias = partial ? nullptr : d->bias;
  p.emb = partial ? nullptr : (const bf16*)d->emb;
  p.emb_uniform = (pl.P * pl.Q) % BM == 0;
  p.stats = reinterpret_cast<float2*>(d->stats);
  p.stats_ld = d->stats_ld;
  p.trace = debug_trace_ptr();
  p.w = d->w;
  r32map = r16map = y32map = y16map = amap;  // (unused maps must still be valid kernel parameters)
  if (partial) {
    p.has_y32 = 1;
    p.partial = 1;
    cuuint64_t dims[3] = {(cuuint64_t)d->K, (cuuint64_t)pl.M, (cuuint64_t)pl.splits};
    cuuint64_t str[2] = {(cuuint64_t)d->K * 4, (cuuint64_t)d->K * 4 * (cuuint64_t)pl.M};
    cuuint32_t box[3] = {PANEL, BM, 1};
    rc = encode(&y32map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, d->workspace, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  } else {
    if (d->residual && d->residual_dtype == MKD_F32) {
      p.has_res32 = 1;
      rc = encode_panel_map(&r32map, true, d->residual, pl.M, n_out, d->ldr);
      if (rc) return rc;
    }
    if (d->residual && d->residual_dtype == MKD_BF16) {
      p.has_res16 = 1;
      rc = encode_panel_map(&r16map, false, d->residual, pl.M, n_out, d->ldr);
      if (rc) return rc;
    }
    if (d->y32) {
      p.has_y32 = 1;
      rc = encode_panel_map(&y32map, true, d->y32, pl.M, n_out, d->ldy32);
      if (rc) return rc;
    }
    if (d->y) {
      p.has_y16 = 1;
      rc = encode_panel_map(&y16map, false, d->y, pl.M, n_out, d->ldy);
      if (rc) return rc;
    }
  }
  const int max_pairs = num_sms() / 2;
  const int pairs = p.num_units < max_pairs ? p.num_units : max_pairs;
  l2_prefetch_plan((unsigned long long)d->K * pl.Ktot * 2 * (d->wgroups == 2 ? 2 : 1), 2 * pairs, p.w_bytes, p.w_share);
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    MKD_LAUNCH_OK(cudaLaunchKernelEx(&cfg, kernel, amap, amap2, bmap, r32map, r16map, y32map, y16map, p));
  }
  MKD_CHECK_LAUNCH();
  if (partial) return splitk_reduce(d, pl.M, pl.P * pl.Q, pl.splits, stream);
  return MKD_OK;
}
}  // namespace

namespace mkd {
bool conv2d_pair_supported(const mkd_conv_desc* d, bool forced) {
  PairPlan pl = {};
  return plan(d, pl, forced);
}

int conv2d_pair(const mkd_conv_desc* d, bool forced, cudaStream_t stream) {
  PairPlan pl = {};
  MKD_REQUIRE(plan(d, pl, forced), MKD_E_INVALID, "gemm_pair: unsupported shape");
  if (pl.geglu) return launch<256, 1, MODE_GEGLU, 0>(d, pl, stream);
  if (pl.nsub == 2) return pl.stats ? launch<160, 2, MODE_PLAIN, 1>(d, pl, stream) : launch<160, 2, MODE_PLAIN, 0>(d, pl, stream);
  return pl.stats ? launch<160, 1, MODE_PLAIN, 1>(d, pl, stream) : launch<160, 1, MODE_PLAIN, 0>(d, pl, stream);
}
}  // namespace mkd
