// NHWC implicit-GEMM convolution / GEMM on the 5th-generation tensor cores (sm_100a):
//   D[128 x BN] (fp32, TMEM) += A[128 x 64] (bf16, smem, K-major, 128B swizzle) * B[BN x 64]^T (bf16, smem, K-major)
//
//   * A (activations) arrives by TMA.  Plain GEMM / 1x1 conv: a 2-D map over [M, K] (row pitch ldx).  3x3 conv:
//     a 4-D map over the NHWC tensor (C, W, H, N) with box (64, Wb, Hb, Nb), Wb*Hb*Nb = 128 output pixels; filter
//     tap (r, s) is the same box shifted by (s - pad, r - pad) — out-of-bounds rows/columns (the padding halo, and
//     the M tail) are zero-filled by the TMA unit, so no im2col buffer and no halo code exist anywhere.
//   * B (weights, [K_out][R*S*C] "KRSC") arrives by a 2-D TMA map, box (64, BN).
//   * persistent: one CTA per SM walks the (n_tile, m_tile, k-split) work units; warp 0 = TMA producer (one lane),
//     warp 1 = MMA issuer (one lane; owns the TMEM allocation), warps 2-9 = epilogue (tcgen05.ld -> registers ->
//     smem staging -> coalesced fused bias / timestep-embedding / alpha / residual (incl. the in-place ControlNet
//     injection) / SiLU / GEGLU -> bf16 and/or fp32 -> global, possibly into a channel slice of a concat buffer).
//   * STAGES-deep smem ring with full/empty mbarriers; tcgen05.commit releases a stage when its MMAs retire.
//   * two TMEM accumulator buffers (tmem_full / tmem_empty mbarriers): the epilogue of unit j overlaps the main loop
//     of unit j+1, and the producer prefetches across unit borders.
//   * split-K (extra work units): each split writes fp32 partials to the workspace; splitk_epilogue_kernel reduces them
//     and applies the same epilogue.  Only where the tiles cannot fill half the SMs and K is deep (the 4x4 / 8x8 levels).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
using namespace mkd;

namespace {

constexpr int BM = 128, BK = 64, UMMA_K = 16;
#ifndef MKD_MAX_STAGES
#define MKD_MAX_STAGES 8  // -DMKD_MAX_STAGES=3: shallower TMA ring (profiles/r01_gemm_stage_sweep.txt)
#endif
#ifndef MKD_GEGLU_PIPE
#define MKD_GEGLU_PIPE 1  // GEGLU on the drain / store split with two 80-column panels in flight (3 ring stages)
#endif
#ifndef MKD_EPI_PIPE
#define MKD_EPI_PIPE 1  // 0 = lock-step epilogue everywhere (A/B builds)
#endif
constexpr int A_BYTES = BM * BK * 2;  // 16 KB

struct EpiP {
  int M, N_out;      // rows, stored output channels (K or K/2 for GEGLU)
  int n_rows;        // weight rows (K)
  int ldy, ldr, lde;
  int pix_per_img;   // P*Q  (emb row = m / pix_per_img)
  int act;
  float alpha;
  bf16* y;           // bf16 output (may be null when y32 is set)
  float* y32;        // optional fp32 copy of the output
  int ldy32, res_f32;
  const float* bias;
  const bf16* emb;
  const void* res;   // bf16 or fp32 (res_f32)
  float* partial;    // split-K workspace or nullptr
  float2* stats;     // optional per-(128-row tile, channel) partial (sum, sum of squares) of the stored values
  int stats_ld;      // float2 elements per tile row of `stats`
  int wg_row;        // split-K reducer only: rows >= wg_row belong to the second weight group (bias offset by n_rows); else INT_MAX
};

struct MainP {
  int kblocks;         // total K blocks (R*S*C / 64)
  int kb_per_split;
  int conv;            // 0: plain 2-D A map, 1: 4-D tap walk
  int cblocks;         // C / 64
  int S, pad;
  int Wb, Hb, Nb;      // box
  int tiles_w, tiles_h;
  int m_tiles, n_tiles, num_units;  // persistent scheduler: unit -> (n_tile, m_tile, split)
  unsigned long long* trace;        // debug: per-CTA %globaltimer stamps (mkd_debug_set_trace), else nullptr
  int half_dw, half_dh, half_dn;    // CL == 2: coordinate offset of rank 1's half of the A box (conv mode)
  int commit_every;                 // G: smem slots are released with one tcgen05.commit per G k-blocks
  int debug;                        // debug timing experiments (results invalid): 1 = skip B loads, 2 = skip MMA issue
  const void* w;                    // weights, for the L2 prefetch ahead of pdl_wait()
  unsigned long long w_bytes;       // (capped) bytes to prefetch, w_share of them per CTA
  unsigned w_share;
};

[[maybe_unused]] __device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
  return t;
}
// slot layout per CTA (16 x u64): 0 entry, 1 prologue done, 2 first TMA issued, 3 first full barrier, 4 unit-0 MMAs
// committed, 5 unit-0 accumulator ready (epilogue), 6 unit-0 epilogue done, 7 last unit epilogue done, 8 exit
// (compiled in only with -DMKD_ENABLE_TRACE, i.e. `MKD_TRACE=1 python -m makeupdiffuse_b200.build --force`)
#ifdef MKD_ENABLE_TRACE
#define MKD_TRACE(slot)                                                     \
  do {                                                                      \
    if (mp.trace) mp.trace[(size_t)blockIdx.x * 16 + (slot)] = gtimer();    \
  } while (0)
#define MKD_DEBUG_BIT(bit) (mp.debug & (bit))  // timing experiments (tools/dbg_epilogue.sh), same build flag
#else
#define MKD_TRACE(slot) do { } while (0)
#define MKD_DEBUG_BIT(bit) false
#endif

// ---- PTX wrappers -------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::
          "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// ---- TMA multicast (CTA pair sharing the A tile) ----------------------------------------------------------------
// the box lands at the same CTA-relative smem offset, and completes on the same CTA-relative mbarrier, in every CTA of
// the cluster named by `mask`
__device__ __forceinline__ void tma_load_2d_mc(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;\n" ::
          "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_mc(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5, %6, %7}], [%2], %3;\n" ::
          "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tcgen05_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}
// one lane of a converged warp (the loops around it stay warp-uniform, so their operands live in uniform registers)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  __syncwarp();  // the role branches above diverged lanes of warps 0/1: reconverge before the .aligned barrier
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n" ::
          "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format): rows of 128 B, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);        // start address  [0,14)
  d |= (uint64_t)0 << 16;                         // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset [32,46)
  d |= (uint64_t)1 << 46;                         // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                         // layout type: SWIZZLE_128B
  return d;
}
// instruction descriptor: D fp32, A/B bf16, both K-major, M = 128, N = BN
__host__ __device__ constexpr uint32_t make_idesc(int bn, int m = BM) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}


// ---- shared epilogue math (used by the main kernel and by the split-K reducer) ------------------------------
// r[8] = accumulators of output channels [o, o+8) of row m.
__device__ __forceinline__ void epilogue_vec8(const EpiP& e, int m, int o, float (&r)[8]) {
  if (o >= e.N_out) return;
  if (o + 8 <= e.N_out) {
    if (e.bias) {
      float t[8];
      load8(e.bias + o, t);
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] += t[i];
    }
    if (e.emb) {
      float t[8];
      load8(e.emb + (int64_t)(m / e.pix_per_img) * e.lde + o, t);
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] += t[i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] *= e.alpha;
    if (e.res) {
      float t[8];
      if (e.res_f32) load8(static_cast<const float*>(e.res) + (int64_t)m * e.ldr + o, t);
      else load8(static_cast<const bf16*>(e.res) + (int64_t)m * e.ldr + o, t);
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] += t[i];
    }
    if (e.act == MKD_ACT_SILU) {
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = silu_f(r[i]);
    }
    if (e.y32) store8(e.y32 + (int64_t)m * e.ldy32 + o, r);
    if (e.y) store8(e.y + (int64_t)m * e.ldy + o, r);
  } else {  // ragged channel tail (e.g. the 4-channel `out` conv): scalar
    const int nimg = e.emb ? m / e.pix_per_img : 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (o + i < e.N_out) {
        float t = r[i] + (e.bias ? e.bias[o + i] : 0.f);
        if (e.emb) t += to_f(e.emb[(int64_t)nimg * e.lde + o + i]);
        t *= e.alpha;
        if (e.res)
          t += e.res_f32 ? static_cast<const float*>(e.res)[(int64_t)m * e.ldr + o + i]
                         : to_f(static_cast<const bf16*>(e.res)[(int64_t)m * e.ldr + o + i]);
        if (e.act == MKD_ACT_SILU) t = silu_f(t);
        if (e.y32) e.y32[(int64_t)m * e.ldy32 + o + i] = t;
        if (e.y) e.y[(int64_t)m * e.ldy + o + i] = from_f<bf16>(t);
      }
    }
  }
}
__device__ __forceinline__ void epilogue_geglu8(const EpiP& e, int m, int row_val, int row_gate, int o, float (&a)[8],
                                                float (&g)[8]) {
  float r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float av = a[i] + (e.bias ? e.bias[row_val + i] : 0.f);
    float gv = g[i] + (e.bias ? e.bias[row_gate + i] : 0.f);
    r[i] = av * gelu_erf_fast(gv);
  }
  store8(e.y + (int64_t)m * e.ldy + o, r);
}
// ---- main kernel: persistent, warp-specialised -------------------------------------------------------------
//   grid = min(#work units, #SMs) CTAs of 320 threads, 1 CTA / SM; every role walks the same unit sequence
//   unit -> (n_tile fastest, m_tile, k-split).
//   warp 0      TMA producer (one lane): STAGES-deep smem ring, full/empty mbarriers, runs ahead across tile borders
//   warp 1      MMA issuer (one lane; owns the TMEM allocation): accumulates tile j into TMEM buffer j & 1
//   warps 2-9   epilogue: drain buffer j & 1 while the MMA warp is already filling the other one
//               (tmem_full / tmem_empty mbarriers), in column panels of PW:
//                 phase 1  TMEM -> registers -> fp32 staging panel in smem (thread = accumulator row)
//                 phase 2  256 threads walk the panel row-major, 8 channels each: coalesced bias / emb / residual
//                          loads and bf16 / fp32 stores (consecutive threads -> consecutive 16 / 32 bytes)
enum { EPI_PLAIN = 0, EPI_SILU = 1, EPI_GEGLU = 2, EPI_PARTIAL = 3, EPI_STATS = 4 };
// EPI_STATS = EPI_PLAIN + GroupNorm statistics of the output: every tile also emits, per output channel, the sum and
// the sum of squares of the fp32 values it stored (column sums over its 128 rows), so that the GroupNorm that consumes
// this tensor is a single streaming pass (mkd_groupnorm_apply) instead of reduce + normalise.  Deterministic: one
// (tile, channel) slot per partial, fixed summation order, no atomics.
// Shared-memory budget: the main loop is bound by how many bytes are in flight per SM (slot round trip = TMA latency
// under load + MMA + two barrier wake-ups ~ 1800+ cycles), so the staging panel is kept narrow (40 columns, 22 KB) and
// every remaining byte of the 227 KB goes to pipeline stages.
template <int BN, int CL, int EPI, int DUAL = 0> struct Cfg {
  // PIPE: the epilogue is split into 4 DRAIN warps (TMEM -> registers -> staging panel) and 8 STORE warps (staging
  // panel -> fused epilogue math -> global), handing double-buffered panels over with named barriers, so that the TMEM
  // read of panel q + 1 (64 B/clk per SM) runs under the shared/global traffic of panel q.  The lock-step version
  // (every thread does both phases, two CTA-wide barriers per panel) overlaps nothing.  EPI_STATS keeps the lock-step
  // epilogue (its column-sum scratch plus a second staging panel would cost the 3x3 convs a pipeline stage).  EPI_GEGLU runs the
  // split scheme on two double-buffered 80-column panels (3 ring stages: FF1 582 -> 607 TFLOP/s; 32-column panels lost 3 %,
  // 16 store warps 7 %, the erf polynomial on the packed fp32 pipe 10 %: the store side is not issue-bound).
  // (split-K partials of the 256-wide tile: lock-step on 32-column panels, so that FOUR 48 KB stages fit — the coupled
  //  ring measures 516 clk per k-block with 4 stages, 593 with 3: profiles/r01_mma_probe.txt)
  static constexpr bool WIDE_PARTIAL = EPI == 3 /*EPI_PARTIAL*/ && BN == 256;
  static constexpr bool PIPE = EPI != 4 /*EPI_STATS*/ && (EPI != 2 /*EPI_GEGLU*/ || MKD_GEGLU_PIPE) && !WIDE_PARTIAL && MKD_EPI_PIPE;
  // lock-step epilogue threads: 8 warps.  (16 warps for GEGLU — ncu shows 11 300 warp-instructions per 128 x 160 tile at
  // 1.6 IPC per SM — measured 7 % SLOWER on the FF1 GEMM: 578 -> 541 TFLOP/s; the parametrisation stays for experiments.)
  static constexpr int ET = 256;
  static constexpr int THREADS = PIPE ? 512 : 64 + ET + 32;
  // staging panel width (columns); GEGLU needs value + gate groups side by side (even group count)
  static constexpr int PW = WIDE_PARTIAL ? 32 : (BN % 80 == 0) ? (EPI == 2 /*EPI_GEGLU*/ ? 80 : 40) : (BN >= 64 ? 64 : 32);
  static constexpr int NP = BN / PW;
  static constexpr int LDT = PW + 4;                                      // +4 floats: conflict-free phase-1 writes
  // DUAL (experiment, off by default — see launch()): one work unit = one A tile against TWO adjacent B tiles (both TMEM
  // accumulator buffers belong to the same unit): half the A bytes per MMA.  Background: the main loop does not speed up
  // with a deeper ring (3, 4 and 5 stages measure the same, profiles/r01_gemm_stage_sweep.txt) nor with a smaller B tile.
  static constexpr int STAGE_BYTES = A_BYTES + (1 + DUAL) * BN * BK * 2;
  static constexpr int PANEL_BYTES = BM * LDT * 4;
  static constexpr int STAGING_BYTES = PANEL_BYTES * (PIPE ? 2 : 1);
  // EPI_STATS: 16 planes (8 channels x {sum, sumsq}) of per-thread column partials, pitch 257 floats
  static constexpr int STATS_PITCH = 257;
  static constexpr int SCRATCH_BYTES = EPI == 4 /*EPI_STATS*/ ? 16 * STATS_PITCH * 4 : 0;
  static constexpr int BUDGET = 227 * 1024 - 1024 /*alignment slack*/ - 512 /*barriers*/ - STAGING_BYTES - SCRATCH_BYTES;
  static constexpr int STAGES_FIT = BUDGET / STAGE_BYTES > 8 ? 8 : BUDGET / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > MKD_MAX_STAGES ? MKD_MAX_STAGES : STAGES_FIT;  // (cap: pipeline-depth experiments)
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + STAGING_BYTES + SCRATCH_BYTES + (2 * STAGES + 4) * 8 + 16 + 1024;
  static_assert(STAGES >= 3 && SMEM <= 227 * 1024, "shared memory budget");
  static_assert(BN % PW == 0, "panels tile the N tile");
  static_assert(!DUAL || (!PIPE && CL == 1 && 2 * BN <= 512), "DUAL: lock-step epilogue, no cluster");
};
__host__ __device__ constexpr int tmem_cols2(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512; }

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}

// CL = CTAs per cluster (1 or 2).  CL == 2: the pair works on the SAME M tile and two adjacent N tiles; each CTA
// fetches half of the shared A tile (64 rows) and TMA-multicasts it into both CTAs' shared memory.
// Why A: timing the main loop with the MMAs and the B loads removed still gave 216 ns per k-block = 16 KB of A per SM
// x 148 SMs = 11.2 TB/s, the L2 -> SM bandwidth cap; B tiles are requested by many CTAs at once and dedup in L2, A
// tiles are private to an M tile.  (Sharing B instead — by multicast or by cta_group::2 — measured no gain.)
// EPI selects the ONE epilogue variant an instantiation carries (a single body holding all of them was ~10^4 SASS
// instructions and instruction-fetch bound in phase 2).
// SPEC != 0 fixes the epilogue's operand set at compile time (bit 0: fp32 residual, bit 1: fp32 output y32, bit 2: bf16
// output y, bit 3: timestep embedding): the run-time dispatch on ep.res / ep.y32 / ep.y / ep.emb cost ~115 SASS instructions
// per 8-channel piece where ~30 do the work, and with two store warps per scheduler the epilogue is issue-bound.
template <int BN, int CL, int EPI, int SPEC = 0, int DUAL = 0>
__global__ void __launch_bounds__((Cfg<BN, CL, EPI, DUAL>::THREADS), 1) gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap amap,
                                                              const __grid_constant__ CUtensorMap bmap, MainP mp, EpiP ep) {
  using C = Cfg<BN, CL, EPI, DUAL>;
  const uint32_t cta_rank = CL > 1 ? cluster_ctarank() : 0;
  const int cl_id = blockIdx.x / CL, cl_num = gridDim.x / CL;
  constexpr int STAGES = C::STAGES, STAGE_BYTES = C::STAGE_BYTES, PW = C::PW, NP = C::NP, LDT = C::LDT;
  constexpr bool PIPE = C::PIPE;
  const bool has_res = SPEC ? (SPEC & 1) != 0 : ep.res != nullptr, res_f32 = SPEC ? true : ep.res_f32 != 0;
  const bool has_y32 = SPEC ? (SPEC & 2) != 0 : ep.y32 != nullptr, has_y = SPEC ? (SPEC & 4) != 0 : ep.y != nullptr;
  const bool has_emb = SPEC ? (SPEC & 8) != 0 : ep.emb != nullptr;  // (bit 3: bf16 timestep-embedding rows, ResBlock conv1)
  // warp roles.  lock-step: 0 A producer, 1 MMA, 2-9 epilogue, 10 B producer.  PIPE: 0 A producer, 1 MMA, 2 B producer,
  // 3 idle, 4-7 drain (TMEM lane quadrant = warp % 4), 8-15 store.
  constexpr int ET = C::ET;  // epilogue (lock-step) / store (PIPE) threads
  constexpr int B_WARP = PIPE ? 2 : 2 + ET / 32, EPI_WARP0 = PIPE ? 8 : 2;
  constexpr int BAR_FULL = 2, BAR_EMPTY = 4, BAR_N = 384;  // named barriers of the panel hand-over (+ buffer index)
  constexpr int TCOLS = tmem_cols2(2 * BN);
  constexpr int NG = PW / 8;  // 8-column groups per panel
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte alignment (128B-swizzle atoms) by OFFSETTING the __shared__ array: arithmetic on it keeps the shared
  // address space, so staging accesses compile to LDS/STS instead of generic LD/ST
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* staging = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
  [[maybe_unused]] float* scratch = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + C::STAGING_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + C::STAGING_BYTES + C::SCRATCH_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) MKD_TRACE(0);
  if (warp == 3) l2_prefetch_share(mp.w, mp.w_bytes, mp.w_share, lane);  // (an epilogue / idle warp: nothing to do until the first accumulator)

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&amap)) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&bmap)) : "memory");
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(full_bar + i, 2);    // the A and the B producer warp
      // release ring: entry q is completed by the q-th release commit (one per G k-blocks, in order); CL == 2: the
      // peer multicasts into these slots too, so both MMA warps must release
      mbar_init(empty_bar + i, CL);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tmem_full_bar + i, 1);
      mbar_init(tmem_empty_bar + i, PIPE ? 4 : ET / 32);  // one arrival per warp that reads TMEM
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {  // TMEM allocation (this warp also frees it)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(TCOLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  tcgen05_fence_before();
  if (CL > 1) cluster_sync_all();  // the peer's barriers exist before anything is multicast to / arrives on them
  else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above (barrier init, TMEM allocation, tensor-map prefetch) overlapped the previous kernel's tail;
  // from here on global memory written by it is read (TMA loads, residuals) and its outputs may be overwritten
  pdl_wait();
  if (threadIdx.x == 0) MKD_TRACE(1);

  if (warp == 0 || warp == B_WARP) {
    // ===== TMA producers: warp 0 streams A, warp B_WARP streams B.  The whole warp walks the loop (warp-uniform ->
    // operands stay in uniform registers), one elected lane issues. =====
    const bool do_a = warp == 0;
    {
      int s = 0;               // smem slot, carried across work units
      int it = 0, rel = 0, rq = 0, rph = 0;  // global k-block counter; releases seen; ring index / phase of the next one
      const int G = min(mp.commit_every, STAGES - 1);  // G >= STAGES would deadlock the ring
      bool first = true;
#pragma unroll 1
      for (int unit = cl_id; unit < mp.num_units; unit += cl_num) {
        // unit -> (n group fastest, m_tile, split); a CL == 2 pair shares m_tile and takes n tiles 2*ng, 2*ng + 1
        const int n_tile = (unit % mp.n_tiles) * (DUAL ? 2 : CL) + (int)cta_rank, rest = unit / mp.n_tiles;
        const int m_tile = rest % mp.m_tiles, split = rest / mp.m_tiles;
        const int kb0 = split * mp.kb_per_split, kb1 = min(mp.kblocks, kb0 + mp.kb_per_split);
        int w0 = 0, h0 = 0, n0 = 0;
        if (mp.conv) {
          const int tw = m_tile % mp.tiles_w, th = (m_tile / mp.tiles_w) % mp.tiles_h, tn = m_tile / (mp.tiles_w * mp.tiles_h);
          w0 = tw * mp.Wb;
          h0 = th * mp.Hb;
          n0 = tn * mp.Nb;
          if (CL == 2 && cta_rank == 1) {  // this CTA fetches (and multicasts) the second half of the A box
            w0 += mp.half_dw;
            h0 += mp.half_dh;
            n0 += mp.half_dn;
          }
        }
        const int m0 = m_tile * BM + (CL == 2 ? (int)cta_rank * (BM / 2) : 0);
        // filter-tap walk (r, sx, cb) kept incrementally: no integer divisions in the k loop
        const int tap0 = kb0 / mp.cblocks;
        int cb = kb0 - tap0 * mp.cblocks, r = tap0 / mp.S;
        int sx = tap0 - r * mp.S;
        const int nb = n_tile * BN;
#pragma unroll 1
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          // slot s was last used by k-block it - STAGES; k-blocks [0, rel * G) have been released by the MMA warp
          while (it - STAGES >= rel * G) {
            mbar_wait(empty_bar + rq, rph);
            ++rel;
            if (++rq == STAGES) {
              rq = 0;
              rph ^= 1;
            }
          }
          unsigned char* sa = smem + s * STAGE_BYTES;
          if (elect_one()) {
            if (do_a) {
              mbar_expect_tx(full_bar + s, A_BYTES);  // CL == 2: my half + the peer's half both land here
              if (CL == 1) {
                if (mp.conv) tma_load_4d(&amap, full_bar + s, sa, cb * BK, w0 + sx - mp.pad, h0 + r - mp.pad, n0);
                else tma_load_2d(&amap, full_bar + s, sa, kb * BK, m0);
              } else {
                unsigned char* dst = sa + cta_rank * (A_BYTES / 2);
                if (mp.conv) tma_load_4d_mc(&amap, full_bar + s, dst, cb * BK, w0 + sx - mp.pad, h0 + r - mp.pad, n0, (uint16_t)3);
                else tma_load_2d_mc(&amap, full_bar + s, dst, kb * BK, m0, (uint16_t)3);
              }
              if (first) MKD_TRACE(2);
            } else {
              if (MKD_DEBUG_BIT(1)) mbar_arrive(full_bar + s);
              else {
                mbar_expect_tx(full_bar + s, STAGE_BYTES - A_BYTES);
                tma_load_2d(&bmap, full_bar + s, sa + A_BYTES, kb * BK, nb);
                if (DUAL) tma_load_2d(&bmap, full_bar + s, sa + A_BYTES + BN * BK * 2, kb * BK, nb + BN);
              }
            }
          }
          __syncwarp();
          first = false;
          if (++cb == mp.cblocks) {
            cb = 0;
            if (++sx == mp.S) {
              sx = 0;
              ++r;
            }
          }
          if (++s == STAGES) s = 0;
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer; warp-uniform loop, one elected lane issues =====
    {
      constexpr uint32_t idesc = make_idesc(BN, BM);
      int s = 0, ph = 0, j = 0;
      int g = 0, cq = 0;  // k-blocks since the last release commit; ring index of the next one
      const int G = min(mp.commit_every, STAGES - 1);
      bool first = true;
#pragma unroll 1
      for (int unit = cl_id; unit < mp.num_units; unit += cl_num, ++j) {
        const int split = (unit / mp.n_tiles) / mp.m_tiles;
        const int kb0 = split * mp.kb_per_split, kb1 = min(mp.kblocks, kb0 + mp.kb_per_split);
        // DUAL: the unit owns both accumulator buffers (j counts units, each buffer is used once per unit)
        const int ab = DUAL ? 0 : (j & 1), use = DUAL ? j : (j >> 1);
        mbar_wait(tmem_empty_bar + ab, (use & 1) ^ 1);  // epilogue has drained this accumulator buffer
        if (DUAL) mbar_wait(tmem_empty_bar + 1, (use & 1) ^ 1);
        tcgen05_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(ab * BN);
#pragma unroll 1
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar + s, ph);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
          const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sa + A_BYTES);
          const bool release = ++g == G;  // tcgen05.commit -> mbarrier sustains only ~1 per 200 ns: release G slots at once
          if (release) g = 0;
          if (elect_one()) {
            if (first) MKD_TRACE(3);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              if (MKD_DEBUG_BIT(2)) break;
              // advance K inside the 128-byte swizzle atom: +32 bytes per UMMA_K (encoded >> 4)
              umma_bf16(tacc, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb > kb0 || k) ? 1u : 0u);
              if (DUAL)  // same A tile, the adjacent B tile, the other accumulator buffer
                umma_bf16(tacc + (uint32_t)BN, adesc + (uint64_t)(k * 2), make_smem_desc(sa + A_BYTES + BN * BK * 2) + (uint64_t)(k * 2),
                          idesc, (kb > kb0 || k) ? 1u : 0u);
            }
            // frees the last G smem slots once the MMAs issued so far retire (they retire in order);
            // CL == 2: in both CTAs — each multicasts into the other
            if (release) {
              if (CL == 1) tcgen05_commit(empty_bar + cq);
              else tcgen05_commit_mc(empty_bar + cq, (uint16_t)3);
            }
          }
          __syncwarp();
          if (release && ++cq == STAGES) cq = 0;
          first = false;
          if (++s == STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
        // accumulator of this unit complete (CL == 2: tell both CTAs' epilogue warps)
        if (elect_one()) {
          tcgen05_commit(tmem_full_bar + ab);
          if (DUAL) tcgen05_commit(tmem_full_bar + 1);
          if (j == 0) MKD_TRACE(4);
        }
        __syncwarp();
      }
    }
  } else if (PIPE && warp >= 4 && warp < 8) {
    // ===== drain warps (PIPE): accumulator rows of TMEM lane quadrant `warp % 4`, panel by panel, into the staging
    // buffer q & 1; full/empty hand-over with the store warps through named barriers =====
    constexpr bool geglu = EPI == EPI_GEGLU;
    const int quad = warp & 3, trow_idx = quad * 32 + lane;
    int j = 0, q = 0;
#pragma unroll 1
    for (int unit = cl_id; unit < mp.num_units; unit += cl_num, ++j) {
      const int ab = j & 1, use = j >> 1;
      mbar_wait(tmem_full_bar + ab, use & 1);
      tcgen05_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(ab * BN);
#pragma unroll 1
      for (int p = 0; p < NP; ++p, ++q) {
        const int b = q & 1;
        uint32_t r[NG][8];
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          // GEGLU panels interleave NG/2 value groups with their NG/2 gate groups (gate columns start at BN/2)
          const int col = geglu ? (g < NG / 2 ? p * (PW / 2) + g * 8 : BN / 2 + p * (PW / 2) + (g - NG / 2) * 8)
                                : p * PW + g * 8;
          tmem_ld8_nowait(trow + col, r[g]);
        }
        // the store warps have consumed panel q - 2 (same buffer); waited for while the TMEM read is in flight
        if (q >= 2) asm volatile("bar.sync %0, %1;\n" ::"r"(BAR_EMPTY + b), "n"(BAR_N) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        float* dst = staging + b * (BM * LDT) + trow_idx * LDT;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          *reinterpret_cast<uint4*>(dst + g * 8) = make_uint4(r[g][0], r[g][1], r[g][2], r[g][3]);
          *reinterpret_cast<uint4*>(dst + g * 8 + 4) = make_uint4(r[g][4], r[g][5], r[g][6], r[g][7]);
        }
        if (p == NP - 1) {  // every TMEM read of this unit is done: hand the accumulator buffer back to the MMA warp
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty_bar + ab);
        }
        __threadfence_block();
        asm volatile("bar.arrive %0, %1;\n" ::"r"(BAR_FULL + b), "n"(BAR_N) : "memory");  // panel staged
      }
    }
  } else if (PIPE ? warp >= 8 : warp < 2 + ET / 32) {
    // ===== epilogue warps.  lock-step (warps 2..9): TMEM lane quadrant = warp % 4, column-group parity = (warp - 2) / 4,
    // every thread drains (phase 1) and stores (phase 2).  PIPE (warps 8..15): phase 2 only. =====
    // Phase-2 work split: thread -> ONE 8-channel column group g (so its bias vector is loaded once per panel) and
    // rows rr, rr + RPI, ... of the tile.  Consecutive threads own consecutive groups of the same row: every global
    // access of a warp is a run of consecutive 16 / 32-byte pieces.
    constexpr int RPI = ET / NG;                  // rows covered per iteration (25 when NG = 10: 6 threads idle)
    constexpr int P2_ITERS = (BM + RPI - 1) / RPI;
    [[maybe_unused]] const int ew = warp - EPI_WARP0, quad = warp & 3, half = ew >> 2;
    const int et = threadIdx.x - EPI_WARP0 * 32;  // 0..ET-1
    [[maybe_unused]] const int trow_idx = quad * 32 + lane;
    const int g2 = et % NG, rr = et / NG;         // phase-2 column group / first row
    const bool p2_active = et < RPI * NG;
    // EPI_STATS: column sums of a panel are reduced one barrier LATER (after the next panel's first barrier), out of a
    // scratch area of their own, so the statistics add no barrier to the drain loop
    [[maybe_unused]] int pend_mtile = -1, pend_ch0 = 0;
    [[maybe_unused]] auto reduce_pending = [&]() {
      constexpr int RP = C::STATS_PITCH;
      if (et < 4 * PW) {  // 4 lanes per channel: {sum, sumsq} x two row halves; whole warps (PW is a multiple of 8)
        const int c = et >> 2, stat = (et >> 1) & 1, hf = et & 1;
        const int plane = stat * 8 + (c & 7), gq = c >> 3;
        constexpr int HALF = (RPI + 1) / 2;
        const int r0 = hf * HALF, r1 = hf ? RPI : HALF;
        // fixed trip count + predicate: unrolled, so the <= 26 loads are issued back to back instead of one shared-memory
        // round trip per row (the dependent loop cost ~800 clk per panel, twice the TMEM drain it runs beside); two
        // interleaved partial sums, fixed order: deterministic
        float acc = 0.f, acc1 = 0.f;
#pragma unroll
        for (int i = 0; i < HALF; i += 2) {
          const int r = r0 + i;
          if (r < r1) acc += scratch[plane * RP + r * NG + gq];
          if (r + 1 < r1) acc1 += scratch[plane * RP + (r + 1) * NG + gq];
        }
        acc += acc1;
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);             // both row halves
        const float sq = __shfl_down_sync(0xffffffffu, acc, 2);  // lane 4c gets the sum of squares from lane 4c + 2
        const int ch = pend_ch0 + c;
        if ((et & 3) == 0 && ch < ep.N_out) ep.stats[(int64_t)pend_mtile * ep.stats_ld + ch] = make_float2(acc, sq);
      }
    };
    constexpr bool geglu = EPI == EPI_GEGLU;
    constexpr bool plain = EPI == EPI_PLAIN || EPI == EPI_SILU || EPI == EPI_STATS;
    // ---- software-pipelined global loads: bias / timestep-embedding / residual of panel q + 1 are requested while
    // panel q is processed (raw registers, unpacked at use), so no panel waits for an L2 / DRAM round trip.  Loading
    // them at the top of their own panel put that latency (~1000+ clk under load) on every panel's critical path:
    // 3 us per 128 x 160 tile, whatever the drain and the stores cost. ----
    struct Pre {
      float bias[8];              // plain: bias of this thread's 8 channels; GEGLU: bias of its value channels
      float bias2[8];             // GEGLU: bias of its gate channels
      uint4 res[P2_ITERS][2];     // residual, raw: 8 fp32 (both) or 8 bf16 ([0])
      uint4 emb[P2_ITERS];        // timestep-embedding row slice, raw bf16
    };
    auto prefetch = [&](int n_tile_, int m_base_, int p_, Pre& L) {
      if constexpr (geglu) {
        constexpr int NV = NG / 2, RPG = ET / NV;
        const int rowv = n_tile_ * BN + p_ * (PW / 2) + (et % NV) * 8;  // weight/bias row of the value channels
#pragma unroll
        for (int k = 0; k < 8; ++k) L.bias[k] = L.bias2[k] = 0.f;
        if (et < RPG * NV && ep.bias) {
          load8(ep.bias + rowv, L.bias);
          load8(ep.bias + rowv + BN / 2, L.bias2);
        }
      } else if constexpr (plain) {
        const int o_ = n_tile_ * BN + p_ * PW + g2 * 8;
        const bool full8_ = p2_active && o_ + 8 <= ep.N_out;
#pragma unroll
        for (int k = 0; k < 8; ++k) L.bias[k] = 0.f;
        if (full8_ && ep.bias) load8(ep.bias + o_, L.bias);
#pragma unroll
        for (int u = 0; u < P2_ITERS; ++u) {
          const int row = rr + u * RPI, m = m_base_ + row;
          if (full8_ && row < BM && m < ep.M) {
            if (has_res && !MKD_DEBUG_BIT(8)) {
              if (res_f32) {
                const uint4* src = reinterpret_cast<const uint4*>(static_cast<const float*>(ep.res) + (int64_t)m * ep.ldr + o_);
                L.res[u][0] = src[0];
                L.res[u][1] = src[1];
              } else {
                L.res[u][0] = *reinterpret_cast<const uint4*>(static_cast<const bf16*>(ep.res) + (int64_t)m * ep.ldr + o_);
              }
            }
            if (has_emb) L.emb[u] = *reinterpret_cast<const uint4*>(ep.emb + (int64_t)(m / ep.pix_per_img) * ep.lde + o_);
          }
        }
      }
    };
    auto unpack8 = [](const uint4& u, float (&v)[8]) {
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h[i]);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
      }
    };
    Pre cur, nxt;
    if (cl_id < mp.num_units) {
      const int n_tile0 = (cl_id % mp.n_tiles) * (DUAL ? 2 : CL) + (int)cta_rank, m_tile0 = (cl_id / mp.n_tiles) % mp.m_tiles;
      prefetch(n_tile0, m_tile0 * BM, 0, cur);
    }
    int j = 0;  // accumulator-buffer uses: one per unit, two (sub = 0, 1: adjacent N tiles) per DUAL unit
    for (int unit = cl_id; unit < mp.num_units; unit += cl_num)
    for (int sub = 0; sub <= DUAL; ++sub, ++j) {
      const int n_tile = DUAL ? (unit % mp.n_tiles) * 2 + sub : (unit % mp.n_tiles) * CL + (int)cta_rank, rest = unit / mp.n_tiles;
      const int m_tile = rest % mp.m_tiles, split = rest / mp.m_tiles;
      const int ab = j & 1, use = j >> 1;
      const int m_base = m_tile * BM;
      [[maybe_unused]] const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(ab * BN);
      [[maybe_unused]] bool acc_ready = false;
#pragma unroll 1
      for (int p = 0; p < NP; ++p) {
        const int o = n_tile * BN + p * PW + g2 * 8;  // first output channel of this thread's group (plain path)
        const bool full8 = plain && p2_active && o + 8 <= ep.N_out;
        // ---- request the NEXT panel's global operands (this panel's were requested one panel ago) ----
        if (p + 1 < NP) {
          prefetch(n_tile, m_base, p + 1, nxt);
        } else if (DUAL && sub == 0) {
          prefetch(n_tile + 1, m_base, 0, nxt);
        } else if (unit + cl_num < mp.num_units) {
          const int nu = unit + cl_num;
          prefetch((nu % mp.n_tiles) * (DUAL ? 2 : CL) + (int)cta_rank, ((nu / mp.n_tiles) % mp.m_tiles) * BM, 0, nxt);
        }
        const float* stg = staging;  // the panel phase 2 reads
        if constexpr (PIPE) {
          const int b = (j * NP + p) & 1;
          stg = staging + b * (BM * LDT);
          asm volatile("bar.sync %0, %1;\n" ::"r"(BAR_FULL + b), "n"(BAR_N) : "memory");  // the drain warps staged it
          if (j == 0 && et == 0 && p == 0) MKD_TRACE(5);
        } else {
        if (!acc_ready) {
          mbar_wait(tmem_full_bar + ab, use & 1);
          if (j == 0 && et == 0) MKD_TRACE(5);
          tcgen05_fence_after();
          acc_ready = true;
        }
        asm volatile("bar.sync 1, %0;\n" ::"n"(ET) : "memory");  // previous panel fully consumed (WAR on the staging panel)
        if (j == 0 && et == 0 && p == 0) MKD_TRACE(9);
        if constexpr (EPI == EPI_STATS) {
          if (pend_mtile >= 0) reduce_pending();  // every thread's partials of the previous panel are in `scratch`
        }
        // ---- phase 1: this thread's accumulator row, column groups g = half, half + 2, ... of the panel ----
        {
          constexpr int HS = ET / 128;  // warps per TMEM lane quadrant: column groups g = half, half + HS, ...
          uint32_t r[(NG + HS - 1) / HS][8];
#pragma unroll
          for (int gi = 0; gi < (NG + HS - 1) / HS; ++gi) {
            const int g = half + HS * gi;
            if (g < NG) {
              // GEGLU panels interleave NG/2 value groups with their NG/2 gate groups (gate columns start at BN/2)
              const int col = geglu ? (g < NG / 2 ? p * (PW / 2) + g * 8 : BN / 2 + p * (PW / 2) + (g - NG / 2) * 8)
                                    : p * PW + g * 8;
              tmem_ld8_nowait(trow + col, r[gi]);
            }
          }
          asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
          if (j == 0 && et == 0 && p == 0) MKD_TRACE(10);
#pragma unroll
          for (int gi = 0; gi < (NG + HS - 1) / HS; ++gi) {
            const int g = half + HS * gi;
            if (g < NG) {
              float* dst = staging + trow_idx * LDT + g * 8;
              *reinterpret_cast<uint4*>(dst) = make_uint4(r[gi][0], r[gi][1], r[gi][2], r[gi][3]);
              *reinterpret_cast<uint4*>(dst + 4) = make_uint4(r[gi][4], r[gi][5], r[gi][6], r[gi][7]);
            }
          }
        }
        if (p == NP - 1) {  // every TMEM read of this unit is done: hand the accumulator buffer back to the MMA warp
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty_bar + ab);
        }
        asm volatile("bar.sync 1, %0;\n" ::"n"(ET) : "memory");  // panel staged (RAW)
        }  // !PIPE
        if (j == 0 && et == 0) MKD_TRACE(p == 0 ? 11 : 13);
        // ---- phase 2: coalesced walk over the panel ----
        [[maybe_unused]] float st_s[8], st_q[8];  // EPI_STATS: this thread's column partials over its rows
        if constexpr (EPI == EPI_STATS) {
#pragma unroll
          for (int k = 0; k < 8; ++k) st_s[k] = st_q[k] = 0.f;
        }
        if constexpr (EPI == EPI_PARTIAL) {
          if (p2_active) {
            const int n = n_tile * BN + p * PW + g2 * 8;
#pragma unroll
            for (int u = 0; u < P2_ITERS; ++u) {
              const int row = rr + u * RPI, m = m_base + row;
              if (row < BM && m < ep.M && n < ep.n_rows) {
                float* dst = ep.partial + ((int64_t)split * ep.M + m) * ep.n_rows + n;
                const float* src = stg + row * LDT + g2 * 8;
                *reinterpret_cast<float4*>(dst) = *reinterpret_cast<const float4*>(src);
                *reinterpret_cast<float4*>(dst + 4) = *reinterpret_cast<const float4*>(src + 4);
              }
            }
          }
        } else if constexpr (geglu) {
          constexpr int NV = NG / 2;  // value groups of the panel; their gate groups follow in the staging row
          const int gv = et % NV, rv0 = et / NV;
          constexpr int RPG = ET / NV, G_ITERS = (BM + RPG - 1) / RPG;
          if (et < RPG * NV) {
            const float(&bv)[8] = cur.bias;
            const float(&bg)[8] = cur.bias2;
#pragma unroll
            for (int u = 0; u < G_ITERS; ++u) {
              const int row = rv0 + u * RPG, m = m_base + row;
              if (row < BM && m < ep.M) {
                float a[8], gt[8], r[8];
                load8(stg + row * LDT + gv * 8, a);
                load8(stg + row * LDT + (NV + gv) * 8, gt);
#pragma unroll
                for (int k = 0; k < 8; ++k) r[k] = (a[k] + bv[k]) * gelu_erf_fast(gt[k] + bg[k]);
                store8(ep.y + (int64_t)m * ep.ldy + n_tile * (BN / 2) + p * (PW / 2) + gv * 8, r);
              }
            }
          }
        } else if (p2_active) {
#pragma unroll
          for (int u = 0; u < P2_ITERS; ++u) {
            const int row = rr + u * RPI, m = m_base + row;
            if (row < BM && m < ep.M && o < ep.N_out) {
              float r[8];
              load8(stg + row * LDT + g2 * 8, r);
              if (full8) {
                if (has_emb) {  // per-sample timestep embedding (ResBlock conv1 only)
                  float t[8];
                  unpack8(cur.emb[u], t);
#pragma unroll
                  for (int k = 0; k < 8; ++k) r[k] += t[k];
                }
                float rs[8];
                if (!has_res || MKD_DEBUG_BIT(8)) {
#pragma unroll
                  for (int k = 0; k < 8; ++k) rs[k] = 0.f;
                } else if (res_f32) {
                  rs[0] = __uint_as_float(cur.res[u][0].x); rs[1] = __uint_as_float(cur.res[u][0].y);
                  rs[2] = __uint_as_float(cur.res[u][0].z); rs[3] = __uint_as_float(cur.res[u][0].w);
                  rs[4] = __uint_as_float(cur.res[u][1].x); rs[5] = __uint_as_float(cur.res[u][1].y);
                  rs[6] = __uint_as_float(cur.res[u][1].z); rs[7] = __uint_as_float(cur.res[u][1].w);
                } else {
                  unpack8(cur.res[u][0], rs);
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) r[k] = (r[k] + cur.bias[k]) * ep.alpha + rs[k];
                if constexpr (EPI == EPI_SILU) {
#pragma unroll
                  for (int k = 0; k < 8; ++k) r[k] = silu_f(r[k]);
                }
                if constexpr (EPI == EPI_STATS) {
#pragma unroll
                  for (int k = 0; k < 8; ++k) {
                    st_s[k] += r[k];
                    st_q[k] = fmaf(r[k], r[k], st_q[k]);
                  }
                }
                if (!MKD_DEBUG_BIT(4)) {
                  if (has_y32) store8(ep.y32 + (int64_t)m * ep.ldy32 + o, r);
                  if (has_y) store8(ep.y + (int64_t)m * ep.ldy + o, r);
                } else if (r[0] == 1234.5f) {  // (debug timing run: keep the math alive without the stores)
                  ep.y32[0] = r[1];
                }
              } else if constexpr (BN == 32) {
                // ragged channel tail (the 4-channel `out` conv, the VAE's 3-channel conv_out): scalar, statically
                // indexed.  Only K < 16 can leave a partial 8-channel group and such filters always take the N = 32
                // tile, so the wider instantiations do not carry this block (the epilogue is sensitive to code size).
                const int nimg = ep.emb ? m / ep.pix_per_img : 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                  if (o + k < ep.N_out) {
                    float t = r[k] + (ep.bias ? ep.bias[o + k] : 0.f);
                    if (ep.emb) t += to_f(ep.emb[(int64_t)nimg * ep.lde + o + k]);
                    t *= ep.alpha;
                    if (ep.res)
                      t += ep.res_f32 ? static_cast<const float*>(ep.res)[(int64_t)m * ep.ldr + o + k]
                                      : to_f(static_cast<const bf16*>(ep.res)[(int64_t)m * ep.ldr + o + k]);
                    if constexpr (EPI == EPI_SILU) t = silu_f(t);
                    if (ep.y32) ep.y32[(int64_t)m * ep.ldy32 + o + k] = t;
                    if (ep.y) ep.y[(int64_t)m * ep.ldy + o + k] = from_f<bf16>(t);
                  }
                }
              }
            }
          }
        }
        if constexpr (EPI == EPI_STATS) {
          // this thread's column partials -> scratch (read after the next barrier; the previous reduce, which read
          // the same scratch, ran before this panel's "staged" barrier, so there is no WAR hazard)
          constexpr int RP = C::STATS_PITCH;
          if (p2_active) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              scratch[k * RP + et] = st_s[k];
              scratch[(8 + k) * RP + et] = st_q[k];
            }
          }
          pend_mtile = m_tile;
          pend_ch0 = n_tile * BN + p * PW;
        }
        if constexpr (PIPE) {  // this thread's reads of the panel are done: the drain warps may refill the buffer
          asm volatile("bar.arrive %0, %1;\n" ::"r"(BAR_EMPTY + ((j * NP + p) & 1)), "n"(BAR_N) : "memory");
        }
        if (j == 0 && et == 0 && p == 0) MKD_TRACE(12);
        cur = nxt;
      }
      if (et == 0) {
        if (j == 0) MKD_TRACE(6);
        MKD_TRACE(7);
      }
    }
    if constexpr (EPI == EPI_STATS) {  // statistics of the very last panel
      asm volatile("bar.sync 1, %0;\n" ::"n"(ET) : "memory");
      if (pend_mtile >= 0) reduce_pending();
    }
  }
  tcgen05_fence_before();
  if (CL > 1) cluster_sync_all();  // no CTA leaves while its peer can still multicast into it or arrive on its barriers
  else __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(TCOLS));
  }
  if (threadIdx.x == 0) MKD_TRACE(8);
}

// split-K reducer: sums the fp32 partials and applies the epilogue; one thread per (row, 16 weight rows).
template <int BN>
__global__ void splitk_epilogue_kernel(EpiP ep, int splits) {
  pdl_wait();
  const int gw = ep.act == MKD_ACT_GEGLU ? 16 : 8;  // channels per thread
  const int groups = ep.n_rows / gw;
  const int64_t total = (int64_t)ep.M * groups;
  const float* const bias0 = ep.bias;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int m = (int)(i / groups), n = (int)(i % groups) * gw;
    ep.bias = (bias0 && m >= ep.wg_row) ? bias0 + ep.n_rows : bias0;  // second weight group (mkd_conv_desc.wgroups)
    auto gather = [&](int col, float (&v)[16]) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = 0.f;
      for (int s = 0; s < splits; ++s) {
        const float4* p = reinterpret_cast<const float4*>(ep.partial + ((int64_t)s * ep.M + m) * ep.n_rows + col);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float4 t = p[j];
          v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
        }
      }
    };
    if (ep.act == MKD_ACT_GEGLU) {
      const int tile = n / BN, c = n % BN;
      if (c >= BN / 2) continue;  // gate columns are consumed together with their value columns
      float a[16], g[16];
      gather(n, a);
      gather(n + BN / 2, g);
      float a0[8], a1[8], g0[8], g1[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { a0[i] = a[i]; a1[i] = a[8 + i]; g0[i] = g[i]; g1[i] = g[8 + i]; }
      epilogue_geglu8(ep, m, n, n + BN / 2, tile * (BN / 2) + c, a0, g0);
      epilogue_geglu8(ep, m, n + 8, n + BN / 2 + 8, tile * (BN / 2) + c + 8, a1, g1);
    } else {
      // 8 channels per thread; every partial of a batch of <= 8 splits is requested before the first add (the serial
      // split loop above costs one L2 round trip per split: ~8 us for 9 splits of a 256 x 1280 tile set).
      // The adds keep the order s = 0, 1, ...: bit-identical to the serial loop.
      {
        constexpr int h = 0;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.f;
        for (int s0 = 0; s0 < splits; s0 += 8) {
          float4 t[8][2];
#pragma unroll
          for (int s = 0; s < 8; ++s) {
            if (s0 + s < splits) {
              const float4* p = reinterpret_cast<const float4*>(ep.partial + ((int64_t)(s0 + s) * ep.M + m) * ep.n_rows + n + 8 * h);
              t[s][0] = p[0];
              t[s][1] = p[1];
            }
          }
#pragma unroll
          for (int s = 0; s < 8; ++s) {
            if (s0 + s < splits) {
              v[0] += t[s][0].x; v[1] += t[s][0].y; v[2] += t[s][0].z; v[3] += t[s][0].w;
              v[4] += t[s][1].x; v[5] += t[s][1].y; v[6] += t[s][1].z; v[7] += t[s][1].w;
            }
          }
        }
        epilogue_vec8(ep, m, n + 8 * h, v);
      }
    }
  }
}

// split-K reducer with a GroupNorm tail (mkd_conv_desc.gn_y): one block per (image, group); thread i owns row i / vpr of the image
// and channels [8 (i % vpr), +8) of the group.  The reduced value gets the ordinary epilogue (stores y / y32), then the block
// reduces (sum, sum of squares) in fp32 and every thread normalises the 8 values it still holds.
struct GnTail {
  bf16* y;
  const float* gamma;
  const float* beta;
  float eps;
  int ld, groups, silu;
  int wsplit;  // images >= wsplit read the second gamma / beta set (weight groups), else INT_MAX
};
__global__ void __launch_bounds__(512) splitk_gn_kernel(EpiP ep, int splits, GnTail gn) {
  pdl_wait();
  __shared__ float red[2][32];
  const int n = blockIdx.x / gn.groups, g = blockIdx.x - n * gn.groups;
  const int cg = ep.n_rows / gn.groups, vpr = cg / 8;
  const int items = ep.pix_per_img * vpr;
  const int i = threadIdx.x;
  const bool live = i < items;
  const int row = live ? i / vpr : 0;
  const int m = n * ep.pix_per_img + row, col = g * cg + (live ? i - row * vpr : 0) * 8;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = 0.f;
  float s = 0.f, ss = 0.f;
  if (live) {
    if (ep.bias && m >= ep.wg_row) ep.bias += ep.n_rows;  // second weight group (mkd_conv_desc.wgroups)
    // same order of additions as splitk_epilogue_kernel (s = 0, 1, ...), loads of up to 8 splits in flight
    for (int s0 = 0; s0 < splits; s0 += 8) {
      float4 t[8][2];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (s0 + q < splits) {
          const float4* p = reinterpret_cast<const float4*>(ep.partial + ((int64_t)(s0 + q) * ep.M + m) * ep.n_rows + col);
          t[q][0] = p[0];
          t[q][1] = p[1];
        }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (s0 + q < splits) {
          v[0] += t[q][0].x; v[1] += t[q][0].y; v[2] += t[q][0].z; v[3] += t[q][0].w;
          v[4] += t[q][1].x; v[5] += t[q][1].y; v[6] += t[q][1].z; v[7] += t[q][1].w;
        }
      }
    }
    epilogue_vec8(ep, m, col, v);  // bias / emb / alpha / residual, y / y32 stores; v now holds the layer's output
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s += v[j];
      ss += v[j] * v[j];
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  if (lane == 0) {
    red[0][warp] = s;
    red[1][warp] = ss;
  }
  __syncthreads();
  if (warp == 0) {
    s = lane < nw ? red[0][lane] : 0.f;
    ss = lane < nw ? red[1][lane] : 0.f;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    if (lane == 0) {
      red[0][0] = s;
      red[1][0] = ss;
    }
  }
  __syncthreads();
  if (!live) return;
  const float cnt = (float)(ep.pix_per_img * cg);
  const float mean = red[0][0] / cnt;
  const float var = fmaxf(red[1][0] / cnt - mean * mean, 0.f);
  const float rstd = rsqrtf(var + gn.eps);
  const int wo = n >= gn.wsplit ? ep.n_rows : 0;
  float ga[8], be[8], r[8];
  load8(gn.gamma + wo + col, ga);
  load8(gn.beta + wo + col, be);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float t = (v[j] - mean) * rstd * ga[j] + be[j];
    r[j] = gn.silu ? silu_f(t) : t;
  }
  store8(gn.y + (int64_t)m * gn.ld + col, r);
}

// Nearest-neighbour x2 upsample, NHWC bf16, 8 channels per thread: out[n, 2h+dy, 2w+dx, :] = in[n, h, w, :].
// (Upsample blocks: the 3x3 conv that follows then runs on the tensor cores like any other.)
__global__ void upsample2x_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int N, int H, int W, int C, int ldx) {
  pdl_wait();
  const int vpr = C / 8;
  const int64_t total = (int64_t)N * 4 * H * W * vpr;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % vpr);
    int64_t pix = i / vpr;
    const int ow = (int)(pix % (2 * W));
    pix /= 2 * W;
    const int oh = (int)(pix % (2 * H)), n = (int)(pix / (2 * H));
    const uint4 val = *reinterpret_cast<const uint4*>(x + ((int64_t)(n * H + (oh >> 1)) * W + (ow >> 1)) * ldx + v * 8);
    *reinterpret_cast<uint4*>(y + (i / vpr) * C + v * 8) = val;
  }
}
// im2col for the three stride-2 Downsample convs: out[m, (r*3+s)*C + c] = in[n, 2p-1+r, 2q-1+s, c] (0 outside).
// The outputs are 4x smaller than the inputs, so the 9x expansion costs little and the conv becomes a plain GEMM.
__global__ void im2col_s2_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int N, int H, int W, int C, int ldx,
                                 int pad) {
  pdl_wait();
  const int vpr = C / 8, P = H / 2, Q = W / 2;
  const int64_t total = (int64_t)N * P * Q * 9 * vpr;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % vpr);
    int64_t t = i / vpr;
    const int tap = (int)(t % 9);
    t /= 9;
    const int q = (int)(t % Q);
    t /= Q;
    const int p = (int)(t % P), n = (int)(t / P);
    const int ih = 2 * p - pad + tap / 3, iw = 2 * q - pad + tap % 3;  // pad 1: UNet Downsample; pad 0: VAE (pad at the far edge)
    uint4 val = make_uint4(0, 0, 0, 0);
    if (ih >= 0 && ih < H && iw >= 0 && iw < W)
      val = *reinterpret_cast<const uint4*>(x + ((int64_t)(n * H + ih) * W + iw) * ldx + v * 8);
    *reinterpret_cast<uint4*>(y + (i / vpr) * C + v * 8) = val;
  }
}

// ---- host side ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(p);
  }
  return fn;
}

int encode(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
           const cuuint32_t* box) {
  EncodeFn fn = get_encode();
  MKD_REQUIRE(fn != nullptr, MKD_E_CUDA, "cuTensorMapEncodeTiled entry point not found (driver too old?)");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MKD_REQUIRE(r == CUDA_SUCCESS, MKD_E_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
  return MKD_OK;
}

struct Geometry {
  int conv, P, Q, M, Ktot, Kout, Wb, Hb, Nb, tiles_w, tiles_h, m_tiles;
};

bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// Which shapes the tensor-core kernel takes.  Sets the error string to the reason when it declines.
bool geometry(const mkd_conv_desc* d, Geometry& g) {
  if (d->dtype != MKD_BF16) { set_error("dtype is not bf16"); return false; }
  if (d->C % BK != 0) { set_error("C=%d is not a multiple of 64", d->C); return false; }
  const bool vae_down = d->R == 3 && d->S == 3 && d->stride == 2 && d->pad == 0 && d->pad_hi_extra == 1;
  if (!vae_down && (d->R != d->S || (d->R != 1 && d->R != 3) || d->pad != d->R / 2 || d->pad_hi_extra != 0)) {
    set_error("filter is not 1x1/p0, 3x3/p1 or 3x3/s2 with far-edge padding");
    return false;
  }
  if (d->stride != 1 || d->upsample) {
    // Downsample (3x3 stride 2) and Upsample (nearest x2 + 3x3) run as: materialise (im2col / upsampled copy) into the
    // workspace, then the ordinary tensor-core kernel.  Needs the workspace and even, power-of-two sizes.
    const bool down = d->stride == 2 && !d->upsample && d->R == 3 && d->H % 2 == 0 && d->W % 2 == 0 &&
                      d->pad + d->pad_hi_extra == 1;  // p1 (UNet) or p0 + one far-edge row/column (VAE encoder)
    const bool up = d->stride == 1 && d->upsample && d->R == 3;
    if (!down && !up) { set_error("stride %d / upsample %d conv is not a Downsample/Upsample shape", d->stride, d->upsample); return false; }
    const size_t need = down ? (size_t)d->N * (d->H / 2) * (d->W / 2) * 9 * d->C * 2 : (size_t)d->N * 4 * d->H * d->W * d->C * 2;
    if (!d->workspace || d->workspace_bytes < need + (1 << 20)) { set_error("stride-2 / upsample conv needs %zu workspace bytes", need); return false; }
  }
  if (d->bias && !aligned16(d->bias)) { set_error("bias alignment"); return false; }
  if (d->ldx % 8 || !aligned16(d->x) || !aligned16(d->w)) { set_error("x/w alignment"); return false; }
  if (d->y && (d->ldy % 8 || !aligned16(d->y))) { set_error("y alignment"); return false; }
  if (d->y32 && (d->ldy32 % 8 || !aligned16(d->y32))) { set_error("y32 alignment"); return false; }
  if (d->act == MKD_ACT_GEGLU && !d->y) { set_error("GEGLU writes the bf16 output only"); return false; }
  if (d->residual && (d->ldr % 8 || !aligned16(d->residual))) { set_error("residual alignment"); return false; }
  if (d->emb && (d->lde % 8 || !aligned16(d->emb))) { set_error("emb alignment"); return false; }
  if (d->stats && (d->act != MKD_ACT_NONE || d->K % 8 || d->stats_ld < d->K || ((uintptr_t)d->stats & 7))) {
    set_error("stats needs act == NONE, K %% 8 == 0, stats_ld >= K, 8-byte aligned pointer");
    return false;
  }
  g.conv = d->R == 3;
  g.P = d->H; g.Q = d->W;
  if (d->upsample) { g.P = 2 * d->H; g.Q = 2 * d->W; }
  if (d->stride == 2) { g.P = d->H / 2; g.Q = d->W / 2; g.conv = 0; }
  g.M = d->N * g.P * g.Q;
  g.Ktot = d->R * d->S * d->C;
  g.Kout = d->act == MKD_ACT_GEGLU ? d->K / 2 : d->K;
  if (d->K % 16 != 0 && !(d->K < 16)) { set_error("K=%d is not a multiple of 16", d->K); return false; }
  // GEGLU row blocking: [80 value | 80 gate] per 160-row tile (single-CTA kernel) or [128 | 128] per 256-row tile (pair kernel)
  if (d->act == MKD_ACT_GEGLU && !((d->geglu_block == 80 && d->K % 160 == 0) || (d->geglu_block == 128 && d->K % 256 == 0 && d->R == 1))) {
    set_error("GEGLU needs geglu_block 80 (K %% 160 == 0) or 128 (K %% 256 == 0, 1x1)");
    return false;
  }
  if (g.conv) {
    if (!is_pow2(g.Q) || !is_pow2(g.P)) { set_error("conv H/W must be powers of two"); return false; }
    g.Wb = g.Q < BM ? g.Q : BM;
    g.Hb = BM / g.Wb < g.P ? BM / g.Wb : g.P;
    g.Nb = BM / (g.Wb * g.Hb);
    g.tiles_w = g.Q / g.Wb;
    g.tiles_h = g.P / g.Hb;
    g.m_tiles = g.tiles_w * g.tiles_h * ((d->N + g.Nb - 1) / g.Nb);
  } else {
    g.Wb = g.Hb = g.Nb = g.tiles_w = g.tiles_h = 1;
    g.m_tiles = (g.M + BM - 1) / BM;
  }
  return true;
}

int pick_bn(const mkd_conv_desc* d, const Geometry& g) {
  if (d->act == MKD_ACT_GEGLU) return 160;
  // deep-K convs on small maps (the 8x8 / 4x4 levels, N = 1280): they run as split-K anyway, so the tile count is free —
  // 5 tiles of 256 instead of 8 of 160 per 128 rows.  The coupled TMA / MMA ring costs ~500 clk per k-block whatever the
  // N tile (profiles/r01_mma_probe.txt): at N = 256 that is the MMA floor, at N = 160 it is 0.6 of it.  MKD_WIDE_SPLITK=0
  // disables (A/B runs).
  {
    static const int we = debug_switch("MKD_WIDE_SPLITK", 1);
    const int kblocks = g.Ktot / BK;
    // measured (tools/gemm_bench.py): 1024 x 1280 x 11520  39.5 -> 34.2 us; at M = 256 (4x4 level, 11 splits) it LOSES
    // (16.5 -> 17.9 us): only from 4 row tiles up
    if (we && d->K % 256 == 0 && d->K % 160 == 0 && d->workspace && !d->stats && d->act == MKD_ACT_NONE && kblocks >= 64 &&
        g.m_tiles >= 4 && g.m_tiles * (d->K / 160) <= 74 && (size_t)6 * g.M * d->K * sizeof(float) <= d->workspace_bytes)
      return 256;
  }
  if (d->K % 160 == 0) return 160;
  if (d->K % 80 == 0) return 80;
  // power-of-two widths (the VAE decoder's 128 / 256 / 512 channels): the widest tile that still leaves at least one
  // work unit per SM — every doubling of the N tile halves the A-tile traffic per FLOP
  if (d->K % 256 == 0 && (int64_t)g.m_tiles * (d->K / 256) >= 148) return 256;
  if (d->K % 128 == 0 && (int64_t)g.m_tiles * (d->K / 128) >= 148) return 128;
  if (d->K % 64 == 0) return 64;
  if (d->K <= 32) return 32;
  return 64;  // ragged last tile: B rows beyond K are zero-filled by TMA, stores are masked
}

unsigned long long* g_trace = nullptr;
int cluster_pref() {  // MKD_CLUSTER=2 selects the A-multicast CTA-pair kernel for BN = 160 (measured ~2 % slower: opt-in)
  static const int v = debug_switch("MKD_CLUSTER", 1);
  return v == 2 ? 2 : 1;
}

template <int BN, int CL>
int launch(const mkd_conv_desc* d, const Geometry& g, cudaStream_t stream) {
  using KernelFn = void (*)(CUtensorMap, CUtensorMap, MainP, EpiP);
  constexpr int EG = (BN == 160 ? EPI_GEGLU : EPI_PLAIN);
  // 0-4: the epilogue variants with run-time operand dispatch; 5-9: compile-time operand sets (SPEC) of the shapes that
  // dominate a UNet step (N tile 160): y32 | res32 + y32 | y | res32 + y (plain), res32 + y32 + y (statistics)
  // 10-11: the statistics variants as DUAL kernels (one A tile against two adjacent N tiles), for shapes with more than
  // one 128 x 160 tile per SM and an even number of N tiles (the 3x3 convs of the 32x32 level)
  constexpr bool SP = BN == 160 && CL == 1;
  // 12: statistics + timestep embedding + y32 (ResBlock conv1), 13: statistics + res32 + y32 (ResBlock conv2), 14: statistics +
  // res32 + y (SpatialTransformer proj_out) — the operand sets the network actually issues (tools: log of ops.conv2d kwargs)
  constexpr int NV = 15;
  const KernelFn all[NV] = {gemm_tcgen05_kernel<BN, CL, EPI_PLAIN>, gemm_tcgen05_kernel<BN, CL, EPI_SILU>,
                            gemm_tcgen05_kernel<BN, CL, EG>, gemm_tcgen05_kernel<BN, CL, EPI_PARTIAL>,
                            gemm_tcgen05_kernel<BN, CL, EPI_STATS>,
                            gemm_tcgen05_kernel<BN, CL, EPI_PLAIN, SP ? 2 : 0>, gemm_tcgen05_kernel<BN, CL, EPI_PLAIN, SP ? 3 : 0>,
                            gemm_tcgen05_kernel<BN, CL, EPI_PLAIN, SP ? 4 : 0>, gemm_tcgen05_kernel<BN, CL, EPI_PLAIN, SP ? 5 : 0>,
                            gemm_tcgen05_kernel<BN, CL, EPI_STATS, SP ? 7 : 0>,
                            gemm_tcgen05_kernel<BN, CL, EPI_STATS, 0, SP ? 1 : 0>, gemm_tcgen05_kernel<BN, CL, EPI_STATS, SP ? 7 : 0, SP ? 1 : 0>,
                            gemm_tcgen05_kernel<BN, CL, EPI_STATS, SP ? 10 : 0>, gemm_tcgen05_kernel<BN, CL, EPI_STATS, SP ? 3 : 0>,
                            gemm_tcgen05_kernel<BN, CL, EPI_STATS, SP ? 5 : 0>};
  const size_t smems[NV] = {Cfg<BN, CL, EPI_PLAIN>::SMEM, Cfg<BN, CL, EPI_SILU>::SMEM, Cfg<BN, CL, EG>::SMEM,
                            Cfg<BN, CL, EPI_PARTIAL>::SMEM, Cfg<BN, CL, EPI_STATS>::SMEM,
                            Cfg<BN, CL, EPI_PLAIN>::SMEM, Cfg<BN, CL, EPI_PLAIN>::SMEM, Cfg<BN, CL, EPI_PLAIN>::SMEM,
                            Cfg<BN, CL, EPI_PLAIN>::SMEM, Cfg<BN, CL, EPI_STATS>::SMEM,
                            Cfg<BN, CL, EPI_STATS, SP ? 1 : 0>::SMEM, Cfg<BN, CL, EPI_STATS, SP ? 1 : 0>::SMEM, Cfg<BN, CL, EPI_STATS>::SMEM,
                            Cfg<BN, CL, EPI_STATS>::SMEM, Cfg<BN, CL, EPI_STATS>::SMEM};
  const int threads[NV] = {Cfg<BN, CL, EPI_PLAIN>::THREADS, Cfg<BN, CL, EPI_SILU>::THREADS, Cfg<BN, CL, EG>::THREADS,
                           Cfg<BN, CL, EPI_PARTIAL>::THREADS, Cfg<BN, CL, EPI_STATS>::THREADS,
                           Cfg<BN, CL, EPI_PLAIN>::THREADS, Cfg<BN, CL, EPI_PLAIN>::THREADS, Cfg<BN, CL, EPI_PLAIN>::THREADS,
                           Cfg<BN, CL, EPI_PLAIN>::THREADS, Cfg<BN, CL, EPI_STATS>::THREADS,
                           Cfg<BN, CL, EPI_STATS, SP ? 1 : 0>::THREADS, Cfg<BN, CL, EPI_STATS, SP ? 1 : 0>::THREADS,
                           Cfg<BN, CL, EPI_STATS>::THREADS, Cfg<BN, CL, EPI_STATS>::THREADS, Cfg<BN, CL, EPI_STATS>::THREADS};
  static bool configured = false;
  if (!configured) {
    for (int i = 0; i < NV; ++i) {
      cudaError_t e = cudaFuncSetAttribute(all[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smems[i]);
      MKD_REQUIRE(e == cudaSuccess, MKD_E_CUDA, "gemm_tcgen05: cudaFuncSetAttribute(%zu): %s", smems[i], cudaGetErrorString(e));
    }
    configured = true;
  }
  CUtensorMap amap, bmap;
  int rc;
  int half_dw = 0, half_dh = 0, half_dn = 0;
  if (g.conv) {
    cuuint64_t dims[4] = {(cuuint64_t)d->C, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
    cuuint64_t str[3] = {(cuuint64_t)d->ldx * 2, (cuuint64_t)d->ldx * 2 * d->W, (cuuint64_t)d->ldx * 2 * d->W * d->H};
    cuuint32_t box[4] = {BK, (cuuint32_t)g.Wb, (cuuint32_t)g.Hb, (cuuint32_t)g.Nb};
    if (CL == 2) {  // each CTA of the pair fetches half of the 128-pixel box: split its outermost non-trivial dimension
      if (g.Nb >= 2) { box[3] = g.Nb / 2; half_dn = g.Nb / 2; }
      else if (g.Hb >= 2) { box[2] = g.Hb / 2; half_dh = g.Hb / 2; }
      else { box[1] = g.Wb / 2; half_dw = g.Wb / 2; }
    }
    rc = encode(&amap, d->x, 4, dims, str, box);
  } else {
    cuuint64_t dims[2] = {(cuuint64_t)d->C, (cuuint64_t)g.M};
    cuuint64_t str[1] = {(cuuint64_t)d->ldx * 2};
    cuuint32_t box[2] = {BK, BM / CL};
    rc = encode(&amap, d->x, 2, dims, str, box);
  }
  if (rc) return rc;
  {
    cuuint64_t dims[2] = {(cuuint64_t)g.Ktot, (cuuint64_t)d->K};
    cuuint64_t str[1] = {(cuuint64_t)g.Ktot * 2};
    cuuint32_t box[2] = {BK, BN};
    rc = encode(&bmap, d->w, 2, dims, str, box);
    if (rc) return rc;
  }
  MainP mp;
  mp.kblocks = g.Ktot / BK;
  mp.conv = g.conv;
  mp.cblocks = d->C / BK;
  mp.S = d->S;
  mp.pad = d->pad;
  mp.Wb = g.Wb; mp.Hb = g.Hb; mp.Nb = g.Nb;
  mp.tiles_w = g.tiles_w; mp.tiles_h = g.tiles_h;
  const int n_tiles = (d->K + BN - 1) / BN;
  // split-K: only when the tile grid leaves more than half of the SMs idle AND K is deep enough that the fp32
  // partial round trip (2 * splits * M * N * 4 bytes) is cheaper than the idle tensor cores
  int splits = 1;
  const int n_groups = (n_tiles + CL - 1) / CL;  // N tiles per cluster (a pair takes two adjacent N tiles)
  const int tiles = g.m_tiles * n_groups * CL;
  if (d->workspace && tiles <= 74 && mp.kblocks >= 32 && d->K % 16 == 0 && !d->stats) {  // (the reducer emits no stats)
    splits = 148 / tiles;
    if (splits > mp.kblocks / 16) splits = mp.kblocks / 16;  // >= 16 K blocks (1024 of K) per split
    if (splits > 16) splits = 16;
    while (splits > 1 && (size_t)splits * g.M * d->K * sizeof(float) > d->workspace_bytes) --splits;
    if (splits < 1) splits = 1;
  }
  mp.kb_per_split = (mp.kblocks + splits - 1) / splits;
  splits = (mp.kblocks + mp.kb_per_split - 1) / mp.kb_per_split;

  EpiP ep;
  ep.M = g.M; ep.N_out = g.Kout; ep.n_rows = d->K;
  ep.ldy = d->ldy; ep.ldr = d->ldr; ep.lde = d->lde;
  ep.pix_per_img = g.P * g.Q;
  ep.act = d->act; ep.alpha = d->alpha;
  ep.y = (bf16*)d->y; ep.bias = d->bias; ep.emb = (const bf16*)d->emb; ep.res = d->residual;
  ep.y32 = d->y32; ep.ldy32 = d->ldy32; ep.res_f32 = d->residual_dtype == MKD_F32;
  ep.partial = splits > 1 ? (float*)d->workspace : nullptr;
  ep.stats = reinterpret_cast<float2*>(d->stats); ep.stats_ld = d->stats_ld;
  ep.wg_row = 0x7fffffff;

  mp.m_tiles = g.m_tiles;
  mp.n_tiles = n_groups;
  mp.half_dw = half_dw; mp.half_dh = half_dh; mp.half_dn = half_dn;
  mp.trace = g_trace;
  mp.w = d->w;
  {
    static const int ge = debug_switch("MKD_COMMIT_EVERY", 0);  // overrides the release granularity (experiments)
    mp.commit_every = ge > 0 ? ge : 1;  // measured: G = 1, 2, 3 give identical k-block times
  }
  {
    // timing experiments (RESULTS INVALID): 1 skip B loads, 2 skip MMA issue, 4 skip epilogue stores, 8 skip residual loads
    mp.debug = debug_switch("MKD_DEBUG_TIMING", 0);
  }
  int variant = ep.partial ? 3 : ep.stats ? 4 : (ep.act == MKD_ACT_GEGLU ? 2 : (ep.act == MKD_ACT_SILU ? 1 : 0));
  if (SP && (variant == 0 || variant == 4) && !ep.emb && (!ep.res || ep.res_f32) && !mp.debug) {
    const int spec = (ep.res ? 1 : 0) | (ep.y32 ? 2 : 0) | (ep.y ? 4 : 0);
    if (variant == 0 && spec >= 2 && spec <= 5) variant = 3 + spec;
    else if (variant == 4 && spec == 7) variant = 9;
    else if (variant == 4 && spec == 3) variant = 13;
    else if (variant == 4 && spec == 5) variant = 14;
  } else if (SP && variant == 4 && ep.emb && !ep.res && ep.y32 && !ep.y && !mp.debug) {
    variant = 12;
  }
  {
    // DUAL units (opt-in, MKD_DUAL=1): meant for shapes where every SM has more than one 128 x 160 tile to do anyway; needs
    // an even number of N tiles.  MEASURED: no gain — 16384 x 320 x 2880 conv1 36.8 -> 39.4 us, conv2 38.0 -> 38.1 us, only
    // K = 8640 gains (82.3 -> 77.3 us); whole step 7.28 -> 7.45 ms.  A dual k-block takes ~2.5x a single one, so the main
    // loop is NOT paced by the A tile alone: time follows the bytes that cross shared memory per k-block (TMA writes + the
    // MMA's operand reads), which DUAL does not reduce per FLOP.  Off by default; kept as the experiment it is.
    static const int de = debug_switch("MKD_DUAL", 0);
    if (SP && de && (variant == 4 || variant == 9) && splits == 1 && n_tiles % 2 == 0 && g.m_tiles * n_tiles > num_sms() && !mp.debug) {
      variant = variant == 4 ? 10 : 11;
      mp.n_tiles = n_tiles / 2;
    }
  }
  mp.num_units = g.m_tiles * mp.n_tiles * splits;  // cluster-level work units
  const int max_clusters = num_sms() / CL;
  const int grid = CL * (mp.num_units < max_clusters ? mp.num_units : max_clusters);
  l2_prefetch_plan((unsigned long long)d->K * g.Ktot * 2, grid, mp.w_bytes, mp.w_share);
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads[variant]);
    cfg.dynamicSmemBytes = smems[variant];
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = CL;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = CL > 1 ? 2 : 1;
    MKD_LAUNCH_OK(cudaLaunchKernelEx(&cfg, all[variant], amap, bmap, mp, ep));
  }
  MKD_CHECK_LAUNCH();
  if (splits > 1) {
    int64_t total = (int64_t)g.M * (d->K / (d->act == MKD_ACT_GEGLU ? 16 : 8));
    int blocks = (int)((total + 127) / 128);
    if (blocks > 148 * 16) blocks = 148 * 16;
    MKD_LAUNCH_OK(launch_pdl(splitk_epilogue_kernel<BN>, dim3(blocks), dim3(128), 0, stream, ep, splits));
    MKD_CHECK_LAUNCH();
  }
  return MKD_OK;
}
}  // namespace

// debug hook (not part of the public header): device buffer of 148*16 u64 that the next GEMM launches stamp
extern "C" void mkd_debug_set_trace(void* p) { g_trace = static_cast<unsigned long long*>(p); }
// debug hook (not part of the public header): 0 = MKD_PATH_AUTO never picks the CTA-pair kernel (bisection / A-B runs)
static int g_pair_auto = 1;
extern "C" void mkd_debug_set_pair_auto(int on) { g_pair_auto = on; }

namespace {
// Downsample (3x3 stride 2) and Upsample (nearest x2 + 3x3): materialise (im2col / upsampled copy) into the head of the
// workspace and rewrite the descriptor onto it — a plain GEMM (im2col) or an ordinary stride-1 3x3 conv (upsample).
int materialise(const mkd_conv_desc* d_in, const Geometry& g, mkd_conv_desc& dd, cudaStream_t stream) {
  dd = *d_in;
  if (d_in->stride != 2 && !d_in->upsample) return MKD_OK;
  const bool down = d_in->stride == 2;
  const size_t bytes = down ? (size_t)d_in->N * g.P * g.Q * 9 * d_in->C * 2 : (size_t)d_in->N * g.P * g.Q * d_in->C * 2;
  const int64_t vecs = (int64_t)(bytes / 16);
  int blocks = (int)((vecs + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (down)
    MKD_LAUNCH_OK(launch_pdl(im2col_s2_kernel, dim3(blocks), dim3(256), 0, stream, (const bf16*)d_in->x, (bf16*)d_in->workspace, d_in->N, d_in->H, d_in->W, d_in->C, d_in->ldx, d_in->pad));
  else
    MKD_LAUNCH_OK(launch_pdl(upsample2x_kernel, dim3(blocks), dim3(256), 0, stream, (const bf16*)d_in->x, (bf16*)d_in->workspace, d_in->N, d_in->H, d_in->W, d_in->C, d_in->ldx));
  MKD_CHECK_LAUNCH();
  const size_t used = (bytes + 1023) & ~(size_t)1023;
  dd.x = d_in->workspace;
  dd.workspace = (char*)d_in->workspace + used;
  dd.workspace_bytes = d_in->workspace_bytes - used;
  dd.stride = 1;
  dd.upsample = 0;
  if (down) { dd.C = 9 * d_in->C; dd.R = dd.S = 1; dd.pad = 0; dd.pad_hi_extra = 0; dd.N = 1; dd.H = 1; dd.W = g.M; dd.ldx = dd.C; }
  else { dd.H = g.P; dd.W = g.Q; dd.ldx = d_in->C; }
  return MKD_OK;
}
}  // namespace

namespace mkd {
unsigned long long* debug_trace_ptr() { return g_trace; }
bool conv2d_pair_supported(const mkd_conv_desc* d, bool forced);          // gemm_pair.cu
int conv2d_pair(const mkd_conv_desc* d, bool forced, cudaStream_t stream);  // gemm_pair.cu

// split-K reducer for partials laid out [split][M][K] fp32 in d->workspace (also used by the pair kernel)
int splitk_reduce(const mkd_conv_desc* d, int M, int pix_per_img, int splits, cudaStream_t stream) {
  EpiP ep;
  ep.M = M; ep.N_out = d->act == MKD_ACT_GEGLU ? d->K / 2 : d->K; ep.n_rows = d->K;
  ep.ldy = d->ldy; ep.ldr = d->ldr; ep.lde = d->lde;
  ep.pix_per_img = pix_per_img;
  ep.act = d->act; ep.alpha = d->alpha;
  ep.y = (bf16*)d->y; ep.bias = d->bias; ep.emb = (const bf16*)d->emb; ep.res = d->residual;
  ep.y32 = d->y32; ep.ldy32 = d->ldy32; ep.res_f32 = d->residual_dtype == MKD_F32;
  ep.partial = (float*)d->workspace;
  ep.stats = nullptr; ep.stats_ld = 0;
  ep.wg_row = d->wgroups == 2 ? M / 2 : 0x7fffffff;
  if (d->gn_y) {  // GroupNorm tail: one block per (image, group), one vector of 8 channels per thread (gemm_pair.cu's plan checked the sizes)
    GnTail gn;
    gn.y = (bf16*)d->gn_y; gn.gamma = d->gn_gamma; gn.beta = d->gn_beta; gn.eps = d->gn_eps;
    gn.ld = d->gn_ld; gn.groups = d->gn_groups; gn.silu = d->gn_silu;
    const int images = M / pix_per_img;
    gn.wsplit = d->wgroups == 2 ? images / 2 : 0x7fffffff;
    const int items = pix_per_img * (d->K / d->gn_groups / 8);
    MKD_REQUIRE(items <= 512 && M % pix_per_img == 0, MKD_E_INVALID, "splitk_reduce: GroupNorm tail needs <= 512 vectors per (image, group)");
    MKD_LAUNCH_OK(launch_pdl(splitk_gn_kernel, dim3(images * d->gn_groups), dim3((items + 31) / 32 * 32), 0, stream, ep, splits, gn));
    MKD_CHECK_LAUNCH();
    return MKD_OK;
  }
  int64_t total = (int64_t)M * (d->K / (d->act == MKD_ACT_GEGLU ? 16 : 8));
  int blocks = (int)((total + 127) / 128);
  if (blocks > 148 * 16) blocks = 148 * 16;
  MKD_LAUNCH_OK(launch_pdl(splitk_epilogue_kernel<160>, dim3(blocks), dim3(128), 0, stream, ep, splits));
  MKD_CHECK_LAUNCH();
  return MKD_OK;
}

// the stride-2 Downsample convs the CTA-pair kernel reads in place (element-strided tensor map): no im2col, no workspace
static bool pair_takes_strided(const mkd_conv_desc* d) {
  return d->stride == 2 && !d->upsample && d->path != MKD_PATH_TCGEN05_SINGLE && (g_pair_auto || d->path == MKD_PATH_TCGEN05_PAIR) &&
         conv2d_pair_supported(d, d->path == MKD_PATH_TCGEN05_PAIR);
}

bool conv2d_tcgen05_supported(const mkd_conv_desc* d) {
  if (d->wgroups == 2) {  // weight groups: the CTA-pair kernel or nothing (the caller then launches once per part)
    if (d->path != MKD_PATH_TCGEN05_SINGLE && !d->upsample && conv2d_pair_supported(d, d->path == MKD_PATH_TCGEN05_PAIR)) return true;
    set_error("conv2d: weight groups need the CTA-pair kernel, which declined this shape");
    return false;
  }
  if (d->gn_y && !d->x2) {  // GroupNorm tail: rides in the CTA-pair kernel's split-K reducer or nowhere
    if (d->path != MKD_PATH_TCGEN05_SINGLE && d->stride == 1 && !d->upsample && conv2d_pair_supported(d, d->path == MKD_PATH_TCGEN05_PAIR)) return true;
    set_error("conv2d: the GroupNorm tail needs a split-K launch of the CTA-pair kernel, which this shape is not");
    return false;
  }
  if (d->x2) {  // second 1x1 term: the CTA-pair kernel or nothing
    if (d->path != MKD_PATH_TCGEN05_SINGLE && d->stride == 1 && !d->upsample && conv2d_pair_supported(d, d->path == MKD_PATH_TCGEN05_PAIR)) return true;
    set_error("conv2d: the x2 term needs the CTA-pair kernel, which declined this shape");
    return false;
  }
  if (pair_takes_strided(d)) return true;
  Geometry g;
  if (!geometry(d, g)) return false;
  if (d->path == MKD_PATH_TCGEN05_PAIR) {  // forced pair kernel (tests / benchmarks): only the shapes it takes
    if (d->stride != 1 || d->upsample || !conv2d_pair_supported(d, true)) {
      set_error("the CTA-pair kernel does not take this shape");
      return false;
    }
  }
  return true;
}

int conv2d_tcgen05(const mkd_conv_desc* d_in, cudaStream_t stream) {
  if (d_in->wgroups == 2 || d_in->x2 || d_in->gn_y || pair_takes_strided(d_in)) return conv2d_pair(d_in, d_in->path == MKD_PATH_TCGEN05_PAIR, stream);
  Geometry g;
  MKD_REQUIRE(geometry(d_in, g), MKD_E_INVALID, "gemm_tcgen05: unsupported shape");
  mkd_conv_desc dd;
  int rc = materialise(d_in, g, dd, stream);
  if (rc) return rc;
  const mkd_conv_desc* d = &dd;
  if (d_in->stride == 2 || d_in->upsample) MKD_REQUIRE(geometry(d, g), MKD_E_INVALID, "gemm_tcgen05: materialised shape rejected");
  // the CTA-pair kernel (cta_group::2, TMA-store epilogue) takes the shapes it is built for; the single-CTA kernel
  // keeps the rest (ragged M, narrow / odd channel counts, tiny problems)
  const bool geglu128 = d->act == MKD_ACT_GEGLU && d->geglu_block == 128;  // a row blocking only the pair kernel reads
  if ((d_in->path != MKD_PATH_TCGEN05_SINGLE && (g_pair_auto || d_in->path == MKD_PATH_TCGEN05_PAIR)) || geglu128) {
    const bool forced = d_in->path == MKD_PATH_TCGEN05_PAIR || geglu128;
    if (conv2d_pair_supported(d, forced)) return conv2d_pair(d, forced, stream);
  }
  MKD_REQUIRE(!geglu128, MKD_E_INVALID, "gemm_tcgen05: geglu_block 128 needs the CTA-pair kernel, which declined this shape");
  const int cl = (d->K > 160) ? cluster_pref() : 1;  // a pair needs two N tiles to share an A tile
  switch (pick_bn(d, g)) {
    case 160: return (cl == 2) ? launch<160, 2>(d, g, stream) : launch<160, 1>(d, g, stream);
    case 256: return launch<256, 1>(d, g, stream);
    case 128: return launch<128, 1>(d, g, stream);
    case 80: return launch<80, 1>(d, g, stream);
    case 64: return launch<64, 1>(d, g, stream);
    default: return launch<32, 1>(d, g, stream);
  }
}
}  // namespace mkd
