// SpatialTransformer attention on the 5th-generation tensor cores (sm_100a): flash-style, S = Q K^T and O += P V are
// tcgen05.mma with TMEM accumulators, Q / K / V tiles arrive by TMA, the Nq x Nkv score matrix never leaves the SM.
// (upstream CrossAttention: sim = q k^T * d^-1/2 -> softmax in fp32 -> @ v; self: Nkv = Nq = H*W, cross: Nkv = 77.)
//
//   CTA = one (batch, head) x NWG * 128 queries.  Warps:
//     0 .. 4*NWG-1   softmax warpgroups: thread = one query row = one TMEM lane, so no cross-thread reduction exists.
//                    Per 128-key tile: pull the S row out of TMEM into registers (which frees the S buffer at once),
//                    row max, p = exp2(..) -> bf16 into a 128B-swizzled K-major smem tile (the A operand of the PV MMA),
//                    fp32 row sum.  O accumulates in TMEM across the key tiles; the flash rescale of O is LAZY: the
//                    running max is only raised when it grew by more than 2^8 (p <= 256 is harmless in bf16 / fp32,
//                    the final O / l is the same number), so the TMEM read-modify-write of O is rare.
//     4*NWG          TMA producer (one lane): Q once, then a K ring and a V ring.
//     4*NWG + 1      MMA issuer (one lane; owns the 512 TMEM columns): S_wg(t+1) = Q K(t+1)^T is issued as soon as
//                    warpgroup wg has S_wg(t) in registers, so the tensor pipe runs a tile ahead of the exp2 phase;
//                    PV_wg(t) follows when P_wg(t) is published.
//   Operands: Q, K tiles are [128 rows][64 channels] boxes per 64-channel chunk, K-major for the QK^T MMA; the V tile is
//   the same box read as an MN-major B operand (N = head dim), so no transpose pass exists.  Head dims that are not a
//   multiple of 64 rely on the tensor map: dim 0 of the map is the head dim itself, so the rest of the 64-wide box is
//   zero-filled by the TMA unit (zeros add nothing to QK^T and give zero O columns that are never stored).
//   Keys beyond Nkv in the last tile are zero-filled the same way and masked to -inf before the softmax.
#include <cuda.h>
#include <stdlib.h>

#include "tcgen05.cuh"
using namespace mkd;
using namespace mkd::tc;

namespace {

constexpr int BQ = 128, BKV = 128;
constexpr int CH = 128 * 128;  // bytes of one [128 rows][64 bf16] chunk

// NWG = softmax warpgroups (128-query tiles) per CTA, KST / VST = K and V ring depths.  Two shapes are used:
//   <DN, 1, 2, 2>  112 KB of smem and 256 TMEM columns for head dims <= 64: TWO CTAs share an SM, so one CTA's start-up,
//                  barrier round trips and non-exp2 phases are covered by the other's exp2 phase (the MUFU pipe is the
//                  bound of this kernel); also what the small maps and the 77-key cross attention want (more CTAs).
//                  K and V are double-buffered: S(t+1) = Q K(t+1)^T is issued a whole tile ahead, and with single
//                  slots every TMA round trip (~1 us = one tile of exp2 work) sat on the critical path (profiled: the
//                  softmax warps spent 27 % of their samples waiting for S, 11 % for O).
//   <DN, 2, ., .>  one CTA per SM with two warpgroups sharing every K / V tile (halves the K/V traffic).
template <int DN, int NWG_, int KST_, int VST_> struct ACfg {
  static constexpr int DCH = (DN + 63) / 64;                 // 64-channel chunks of the head dim
  static constexpr int NWG = NWG_, KST = KST_, VST = VST_;
  static constexpr int OSTR = (DN + 31) / 32 * 32;           // TMEM column stride between the O accumulators
  static constexpr int TCOLS = NWG * 128 + NWG * OSTR <= 256 ? 256 : 512;
  static constexpr int THREADS = 32 * (4 * NWG + 2);
  static constexpr int NBAR = 5 * NWG + 2 * KST + 2 * VST;   // q_full, s_full, s_free, p_full, o_full | k/v full/empty
  static constexpr size_t SMEM = (size_t)(NWG * DCH + (KST + VST) * DCH + 2 * NWG) * CH + NBAR * 8 + 16;
  static_assert(NWG * 128 + NWG * OSTR <= 512, "TMEM columns");
  static_assert(SMEM <= 227 * 1024, "shared memory");
};

template <int DN, int NWG_, int KST_, int VST_>
__global__ void __launch_bounds__(ACfg<DN, NWG_, KST_, VST_>::THREADS, ACfg<DN, NWG_, KST_, VST_>::TCOLS == 256 ? 2 : 1)
    attn_tcgen05_kernel(const __grid_constant__ CUtensorMap qmap, const __grid_constant__ CUtensorMap kmap,
                        const __grid_constant__ CUtensorMap vmap, bf16* __restrict__ o, int Nq, int Nkv, int d, int ldo,
                        float sl2) {
  using C = ACfg<DN, NWG_, KST_, VST_>;
  constexpr int DCH = C::DCH, NWG = C::NWG, KST = C::KST, VST = C::VST;
  // 128B-swizzle atoms need 1024-byte alignment; the kernel has no static shared memory, so the aligned dynamic array
  // starts the CTA's window (checked: a misplaced base traps instead of corrupting operands).  No slack bytes: with
  // 112 KB + barriers two CTAs fit one SM exactly.
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw;
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
  unsigned char* Qs = smem;                         // [NWG][DCH] chunks
  unsigned char* Ks = Qs + NWG * DCH * CH;          // [KST][DCH]
  unsigned char* Vs = Ks + KST * DCH * CH;          // [VST][DCH]
  unsigned char* Ps = Vs + VST * DCH * CH;          // [NWG][2]  (128 queries x 128 keys bf16)
  uint64_t* q_full = reinterpret_cast<uint64_t*>(Ps + NWG * 2 * CH);
  uint64_t* s_full = q_full + NWG;
  uint64_t* s_free = s_full + NWG;
  uint64_t* p_full = s_free + NWG;
  uint64_t* o_full = p_full + NWG;
  uint64_t* k_full = o_full + NWG;
  uint64_t* k_empty = k_full + KST;
  uint64_t* v_full = k_empty + KST;
  uint64_t* v_empty = v_full + VST;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(v_empty + VST);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int W_PROD = 4 * NWG, W_MMA = 4 * NWG + 1;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * (BQ * NWG);
  const int T = (Nkv + BKV - 1) / BKV;
  const int nwg = (NWG == 2 && q0 + BQ < Nq) ? 2 : 1;  // warpgroups with at least one real query

  if (threadIdx.x == 0) {
    prefetch_tensormap(&qmap);
    prefetch_tensormap(&kmap);
    prefetch_tensormap(&vmap);
    for (int i = 0; i < NWG; ++i) {
      mbar_init(q_full + i, 1);
      mbar_init(s_full + i, 1);
      mbar_init(s_free + i, 4);  // one arrival per softmax warp
      mbar_init(p_full + i, 4);
      mbar_init(o_full + i, 1);
    }
    for (int i = 0; i < KST; ++i) {
      mbar_init(k_full + i, 1);
      mbar_init(k_empty + i, 1);
    }
    for (int i = 0; i < VST; ++i) {
      mbar_init(v_full + i, 1);
      mbar_init(v_empty + i, 1);
    }
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc<C::TCOLS>(tmem_slot);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  if (warp == W_PROD) {
    if (lane == 0) {
      for (int wg = 0; wg < nwg; ++wg) {
        mbar_expect_tx(q_full + wg, DCH * CH);
        for (int c = 0; c < DCH; ++c) tma_load_4d(&qmap, q_full + wg, Qs + (wg * DCH + c) * CH, c * 64, h, q0 + wg * BQ, b);
      }
      for (int t = 0; t < T; ++t) {
        const int ks = t % KST, vs = t % VST;
        if (t >= KST) mbar_wait(k_empty + ks, (t / KST - 1) & 1);
        mbar_expect_tx(k_full + ks, DCH * CH);
        for (int c = 0; c < DCH; ++c) tma_load_4d(&kmap, k_full + ks, Ks + (ks * DCH + c) * CH, c * 64, h, t * BKV, b);
        if (t >= VST) mbar_wait(v_empty + vs, (t / VST - 1) & 1);
        mbar_expect_tx(v_full + vs, DCH * CH);
        for (int c = 0; c < DCH; ++c) tma_load_4d(&vmap, v_full + vs, Vs + (vs * DCH + c) * CH, c * 64, h, t * BKV, b);
      }
    }
  } else if (warp == W_MMA) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = idesc_bf16(BQ, BKV, false);
      constexpr uint32_t idesc_o = idesc_bf16(BQ, DN, true);
      const int ksteps = (d + 15) >> 4;  // 16-channel steps of the QK^T contraction (zero-filled beyond d)
      auto issue_s = [&](int wg, int st) {
        const uint32_t a = smem_u32(Qs + wg * DCH * CH), bk = smem_u32(Ks + st * DCH * CH);
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint32_t off = (uint32_t)(ks >> 2) * CH + (uint32_t)(ks & 3) * 32;
          mma_bf16_ss(tmem_base + (uint32_t)(wg * BKV), smem_desc_sw128(a + off), smem_desc_sw128(bk + off), idesc_s, ks > 0);
        }
      };
      auto issue_pv = [&](int wg, int st, int kvalid, bool accumulate) {
        const uint32_t a = smem_u32(Ps + wg * 2 * CH), bv = smem_u32(Vs + st * DCH * CH);
        const int steps = (min(kvalid, BKV) + 15) >> 4;  // keys beyond Nkv carry P = 0: skip whole 16-key steps
        for (int kk = 0; kk < steps; ++kk) {
          const uint32_t aoff = (uint32_t)(kk >> 2) * CH + (uint32_t)(kk & 3) * 32;
          mma_bf16_ss(tmem_base + (uint32_t)(NWG * BKV + wg * C::OSTR), smem_desc_sw128(a + aoff),
                      smem_desc_sw128(bv + (uint32_t)kk * 2048, CH), idesc_o, accumulate || kk > 0);
        }
      };
      for (int wg = 0; wg < nwg; ++wg) mbar_wait(q_full + wg, 0);
      mbar_wait(k_full + 0, 0);
      fence_after();
      for (int wg = 0; wg < nwg; ++wg) {
        issue_s(wg, 0);
        commit(s_full + wg);
      }
      commit(k_empty + 0);
      for (int t = 0; t < T; ++t) {
        const int st = t % VST;
        if (t + 1 < T) {  // S(t+1) as soon as the warpgroup holds S(t) in registers
          const int st1 = (t + 1) % KST;
          mbar_wait(k_full + st1, ((t + 1) / KST) & 1);
          for (int wg = 0; wg < nwg; ++wg) {
            mbar_wait(s_free + wg, t & 1);
            fence_after();
            issue_s(wg, st1);
            commit(s_full + wg);
          }
          commit(k_empty + st1);
        }
        mbar_wait(v_full + st, (t / VST) & 1);
        for (int wg = 0; wg < nwg; ++wg) {
          mbar_wait(p_full + wg, t & 1);  // P_wg(t) is in smem and any rescale of O_wg is done
          fence_after();
          issue_pv(wg, st, Nkv - t * BKV, t > 0);
          commit(o_full + wg);
        }
        commit(v_empty + st);
      }
    }
  } else if ((warp >> 2) < nwg) {
    // ===== softmax warpgroup: thread = query row =====
    const int wg = warp >> 2, row = (warp & 3) * 32 + lane;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tS = tmem_base + lane_base + (uint32_t)(wg * BKV);
    const uint32_t tO = tmem_base + lane_base + (uint32_t)(NWG * BKV + wg * C::OSTR);
    unsigned char* Pw = Ps + wg * 2 * CH + row * 128;
    const uint32_t swz = (uint32_t)(row & 7);
    float m_used = -INFINITY, l = 0.f;  // running (lazily raised) row max, row sum of p

    for (int t = 0; t < T; ++t) {
      mbar_wait(s_full + wg, t & 1);
      fence_after();
      const int kvalid = Nkv - t * BKV;  // < 128 only in the last tile
      uint32_t v[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32_nowait(tS + c * 32, v[c]);
      tmem_ld_wait();
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free + wg);  // the S buffer may be overwritten by Q K(t+1)^T now
      if (kvalid < BKV) {  // keys beyond Nkv (zero-filled K rows): -inf -> p = 0
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[c][i] = (c * 32 + i >= kvalid) ? 0xff800000u : v[c][i];
        }
      }
      // ---- row max: 3-input max (FMNMX3), two keys per instruction, 8 independent chains ----
      float mx8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) mx8[i] = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) mx8[(i >> 1) & 7] = max3(mx8[(i >> 1) & 7], __uint_as_float(v[c][i]), __uint_as_float(v[c][i + 1]));
      }
      const float mx = fmaxf(max3(mx8[0], mx8[1], mx8[2]), max3(mx8[3], mx8[4], fmaxf(max3(mx8[5], mx8[6], mx8[7]), -INFINITY)));
      if (t > 0) {  // PV(t-1) has retired: O(t-1) is complete and the P tile may be rewritten
        mbar_wait(o_full + wg, (t - 1) & 1);
        fence_after();
      }
      // ---- lazy rescale: raise the running max only when it grew by more than 2^8 (warp-uniform decision, the
      //      TMEM accesses below are warp-collective) ----
      if (__any_sync(0xffffffffu, (mx - m_used) * sl2 > 8.0f)) {
        const float m_new = fmaxf(m_used, mx);
        const float corr = ex2((m_used - m_new) * sl2);  // first tile: ex2(-inf) = 0
        l *= corr;
        m_used = m_new;
        if (t > 0) {
#pragma unroll
          for (int g = 0; g < DN / 16; ++g) {
            uint32_t r[16];
            tmem_ld16_nowait(tO + g * 16, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * corr);
            tmem_st16(tO + g * 16, r);
          }
          tmem_st_wait();
        }
      }
      // ---- p = exp2((s - max) * scale * log2 e) -> bf16 -> swizzled smem; fp32 row sum (4 chains) ----
      // (packed fp32 pipe: one FFMA2 scales-and-shifts two keys, one FADD2 adds two p to two running sums: 3
      //  instructions per key instead of 4.5 — the loop is issue-bound, MUFU sits at ~40 %)
      const float mb = m_used * sl2;
      const uint64_t sl2x2 = pack2(sl2, sl2), nmbx2 = pack2(-mb, -mb);
      uint64_t sum2[4] = {pack2(0.f, 0.f), pack2(0.f, 0.f), pack2(0.f, 0.f), pack2(0.f, 0.f)};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float a0, a1;
          unpack2(ffma2(pack2(__uint_as_float(v[c][i]), __uint_as_float(v[c][i + 1])), sl2x2, nmbx2), a0, a1);
          const float p0 = ex2(a0), p1 = ex2(a1);
          sum2[(i >> 1) & 3] = fadd2(sum2[(i >> 1) & 3], pack2(p0, p1));
          __nv_bfloat162 hh = __floats2bfloat162_rn(p0, p1);
          pk[i >> 1] = *reinterpret_cast<uint32_t*>(&hh);
        }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const uint32_t u = (uint32_t)(c * 4 + jj);  // 16-byte unit (8 keys) of the 128-key row
          unsigned char* dst = Pw + (u >> 3) * CH + (((u & 7) ^ swz) << 4);
          *reinterpret_cast<uint4*>(dst) = make_uint4(pk[4 * jj], pk[4 * jj + 1], pk[4 * jj + 2], pk[4 * jj + 3]);
        }
      }
      {
        float s0, s1, s2, s3;
        unpack2(fadd2(sum2[0], sum2[1]), s0, s1);
        unpack2(fadd2(sum2[2], sum2[3]), s2, s3);
        l += (s0 + s1) + (s2 + s3);
      }
      fence_before();            // TMEM accesses above ordered before the MMA warp's PV(t)
      fence_proxy_async_smem();  // P visible to the tensor core's operand fetch
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full + wg);
    }
    mbar_wait(o_full + wg, (T - 1) & 1);
    fence_after();
    const int q = q0 + wg * BQ + row;
    const float inv = 1.0f / l;
    bf16* dst = o + ((int64_t)b * Nq + q) * ldo + h * d;
#pragma unroll
    for (int g = 0; g < DN / 16; ++g) {
      uint32_t r[16];
      tmem_ld16_nowait(tO + g * 16, r);  // warp-collective: every lane loads, only real rows / channels store
      tmem_ld_wait();
#pragma unroll
      for (int hlf = 0; hlf < 2; ++hlf) {
        if (q < Nq && g * 16 + hlf * 8 < d) {
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(r[hlf * 8 + i]) * inv;
          store8(dst + g * 16 + hlf * 8, f);
        }
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    fence_after();
    tmem_dealloc<C::TCOLS>(tmem_base);
  }
}

// ---- head dims <= 64: two independent 64-key half pipelines per 128 queries, two CTAs per SM -----------------------
// The keys of a row are split between two warpgroups that share nothing but Q: warpgroup g takes the 64-key halves
// g, g + 2, ... with its own issuer warp (TMA loads of its K / V halves, its S and PV MMAs), its own S buffer (64 TMEM
// columns), P buffer (16 KB), K / V rings (2 x 8 KB each), O accumulator and running (max, sum).  The two partial
// results of a row are merged once after the key loop:
//   m = max(m0, m1), a_g = 2^((m_g - m) scale), O = (a0 O0 + a1 O1) / (a0 l0 + a1 l1).
// 16 softmax warps per SM (4 per scheduler) at <= 96 registers (a thread holds 64 scores, not 128); issuers are whole
// warps walking warp-uniform loops with one elected lane issuing (an `if (lane == 0)` region costs ~15 instructions of
// uniformisation loop per MMA), descriptors built once and advanced by constants.
// Measured (profiles/r02_attention.txt): 446 us on 8 x 8 heads x 4096 tokens (385 TFLOP/s; whole-tile kernel above: 452),
// 70.8 us on 16 x 8 x 1024 (was 75).  The experiments that led here say no single pipe bounds this loop: timing-only builds
// without the exp2, without the P stores, without the MMAs, without the K / V loads each moved the time by < 3 % (all of
// them together: -32 %); in this form XU runs at 63 %, L2 tag lookups at 47 % (71 % on the busiest slice: 80-byte rows at
// head dim 40), issue slots at 50 %, commit -> barrier latency is ~110 clk, one unobstructed iteration of a warpgroup is
// 1 630 clk against 3 100-3 600 with four of them on an SM.
// Barriers of stream g, u-th half: s_full (commit), s_free (4 warps: S is in registers), p_full (4 warps), pv_done
// (commit: P may be rewritten, O holds every half up to u), k_full / v_full[2].
constexpr int HKV = 64;
constexpr int HB = HKV * 128;  // bytes of one [64 keys][64 ch] half tile

template <int DN>
__global__ void __launch_bounds__(320, 2)
    attn_split_kernel(const __grid_constant__ CUtensorMap qmap, const __grid_constant__ CUtensorMap kmap,
                      const __grid_constant__ CUtensorMap vmap, bf16* __restrict__ o, int Nq, int Nkv, int d, int ldo, float sl2) {
  static_assert(DN % 16 == 0 && DN <= 64, "one 64-channel chunk per head");
  constexpr int TCOLS = 256;          // S_g at column 64 g, O_g at column 128 + 64 g
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw;
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
  unsigned char* Qs = smem;                 // [128 queries][64 ch]
  unsigned char* Ks = Qs + CH;              // [2 streams][2 slots][64 keys][64 ch]
  unsigned char* Vs = Ks + 4 * HB;          // [2 streams][2 slots][64 keys][64 ch]
  unsigned char* Ps = Vs + 4 * HB;          // [2 streams][128 queries][64 keys]
  uint64_t* q_full = reinterpret_cast<uint64_t*>(Ps + 2 * CH);
  uint64_t* s_full = q_full + 1;   // [2]
  uint64_t* s_free = s_full + 2;   // [2]
  uint64_t* p_full = s_free + 2;   // [2]
  uint64_t* pv_done = p_full + 2;  // [2]
  uint64_t* k_full = pv_done + 2;  // [2 streams][2 slots]
  uint64_t* v_full = k_full + 4;   // [2 streams][2 slots]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(v_full + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int W_ISSUE = 8;  // warps 8, 9: issuer of stream 0 / 1
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * BQ;
  const int H = (Nkv + HKV - 1) / HKV;    // 64-key halves; stream g owns halves g, g + 2, ...

  if (threadIdx.x == 0) {
    prefetch_tensormap(&qmap);
    prefetch_tensormap(&kmap);
    prefetch_tensormap(&vmap);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(s_full + i, 1);
      mbar_init(s_free + i, 4);
      mbar_init(p_full + i, 4);
      mbar_init(pv_done + i, 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(k_full + i, 1);
      mbar_init(v_full + i, 1);
    }
    fence_barrier_init();
  }
  if (warp == W_ISSUE) tmem_alloc<TCOLS>(tmem_slot);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  if (warp >= W_ISSUE) {
    // ===== issuer of stream g: the whole warp walks the loop (warp-uniform waits), one elected lane issues; descriptors are
    // built once and advanced by constants, the k loops are unrolled at compile time =====
    const int g = warp - W_ISSUE;
    const int n = (H - g + 1) >> 1;  // halves of this stream
    constexpr uint32_t idesc_s = idesc_bf16(BQ, HKV, false);
    constexpr uint32_t idesc_o = idesc_bf16(BQ, DN, true);
    constexpr int KS = DN / 16;  // 16-channel steps of Q K^T (channels beyond d are zero-filled by the TMA unit)
    unsigned char* Kg = Ks + g * 2 * HB;
    unsigned char* Vg = Vs + g * 2 * HB;
    const uint64_t qdesc = smem_desc_sw128(smem_u32(Qs));
    const uint64_t kdesc0 = smem_desc_sw128(smem_u32(Kg));
    const uint64_t vdesc0 = smem_desc_sw128(smem_u32(Vg), CH);
    const uint64_t pdesc = smem_desc_sw128(smem_u32(Ps + g * CH));
    const uint32_t dS = tmem_base + (uint32_t)(g * HKV), dO = tmem_base + (uint32_t)(2 * HKV + g * 64);
    auto load_k = [&](int u) {  // K half g + 2 u into slot u & 1
      if (elect_one()) {
        mbar_expect_tx(k_full + 2 * g + (u & 1), HB);
        tma_load_4d(&kmap, k_full + 2 * g + (u & 1), Kg + (u & 1) * HB, 0, h, (g + 2 * u) * HKV, b);
      }
      __syncwarp();
    };
    auto load_v = [&](int u) {
      if (elect_one()) {
        mbar_expect_tx(v_full + 2 * g + (u & 1), HB);
        tma_load_4d(&vmap, v_full + 2 * g + (u & 1), Vg + (u & 1) * HB, 0, h, (g + 2 * u) * HKV, b);
      }
      __syncwarp();
    };
    auto issue_s = [&](int u) {  // S(u) = Q K(u)^T from K slot u & 1
      mbar_wait(k_full + 2 * g + (u & 1), (u >> 1) & 1);
      fence_after();
      if (elect_one()) {
        const uint64_t kd = kdesc0 + (uint64_t)((u & 1) * (HB >> 4));
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) mma_bf16_ss(dS, qdesc + (uint64_t)(ks * 2), kd + (uint64_t)(ks * 2), idesc_s, ks > 0);
        commit(s_full + g);
      }
      __syncwarp();
    };
    if (n > 0) {
      if (g == 0 && elect_one()) {
        mbar_expect_tx(q_full, CH);
        tma_load_4d(&qmap, q_full, Qs, 0, h, q0, b);
      }
      __syncwarp();
      load_k(0);
      load_v(0);
      if (n > 1) {
        load_k(1);
        load_v(1);
      }
      mbar_wait(q_full, 0);
      issue_s(0);
      // Slot reuse without a blocking wait on this stream's own commits: s_free(u) says S(u) is in registers, so K slot
      // u & 1 is free a whole iteration before S(u + 2) needs it; p_full(u) is sent only after the warpgroup has seen
      // pv_done(u - 1), so V slot (u + 1) & 1 is free one iteration before PV(u + 1) needs it.
      for (int u = 0; u < n; ++u) {
        const int s = u & 1;
        if (u + 1 < n) {
          mbar_wait(s_free + g, u & 1);
          if (u + 2 < n) load_k(u + 2);
          issue_s(u + 1);
        }
        mbar_wait(v_full + 2 * g + s, (u >> 1) & 1);
        mbar_wait(p_full + g, u & 1);    // P(u) is in smem and any rescale of O is done
        fence_after();
        if (u >= 1 && u + 1 < n) load_v(u + 1);
        if (elect_one()) {
          const uint64_t vd = vdesc0 + (uint64_t)(s * (HB >> 4));
          const int kvalid = Nkv - (g + 2 * u) * HKV;
          if (kvalid >= HKV) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) mma_bf16_ss(dO, pdesc + (uint64_t)(kk * 2), vd + (uint64_t)(kk * 128), idesc_o, u > 0 || kk > 0);
          } else {  // keys beyond Nkv carry P = 0: skip whole 16-key steps
            const int steps = (kvalid + 15) >> 4;
            for (int kk = 0; kk < steps; ++kk) mma_bf16_ss(dO, pdesc + (uint64_t)(kk * 2), vd + (uint64_t)(kk * 128), idesc_o, u > 0 || kk > 0);
          }
          commit(pv_done + g);
        }
        __syncwarp();
      }
    }
  } else {
    // ===== softmax warpgroups: thread = query row, warpgroup g = stream g =====
    const int g = warp >> 2, row = (warp & 3) * 32 + lane;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tS = tmem_base + lane_base + (uint32_t)(g * HKV);
    const uint32_t tO = tmem_base + lane_base + (uint32_t)(2 * HKV + g * 64);
    const uint32_t pw = smem_u32(Ps + g * CH + row * 128) + (((uint32_t)row & 7u) << 4);  // 16-byte unit j of the row: pw ^ (j << 4)
    float m_used = -INFINITY, l = 0.f;
    int u = 0;
    for (int hh = g; hh < H; hh += 2, ++u) {
      mbar_wait(s_full + g, u & 1);
      fence_after();
      uint32_t v[2][32];
      tmem_ld32_nowait(tS, v[0]);
      tmem_ld32_nowait(tS + 32, v[1]);
      tmem_ld_wait();
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free + g);  // S_g may be overwritten by Q K(u + 1)^T
      const int kvalid = Nkv - hh * HKV;       // < 64 only in the last half
      if (kvalid < HKV) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[c][i] = (c * 32 + i >= kvalid) ? 0xff800000u : v[c][i];
        }
      }
      float mx8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) mx8[i] = -INFINITY;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) mx8[(i >> 1) & 7] = max3(mx8[(i >> 1) & 7], __uint_as_float(v[c][i]), __uint_as_float(v[c][i + 1]));
      }
      const float mx = fmaxf(max3(mx8[0], mx8[1], mx8[2]), max3(mx8[3], mx8[4], max3(mx8[5], mx8[6], mx8[7])));
      // The round trip p_full -> issuer -> PV -> commit -> pv_done is ~1 500 clk: waiting for pv_done(u - 1) HERE (before the
      // exp2 phase) left the warpgroup idle for a third of every iteration.  Only the P stores and the (rare) rescale of O
      // depend on it, so the exp2 phase runs first, into registers (64 scores become 32 packed bf16 pairs).
      if (__any_sync(0xffffffffu, (mx - m_used) * sl2 > 8.0f)) {  // lazy rescale (warp-uniform: TMEM accesses are collective)
        const float m_new = fmaxf(m_used, mx);
        const float corr = ex2((m_used - m_new) * sl2);
        l *= corr;
        m_used = m_new;
        if (u > 0) {
          mbar_wait(pv_done + g, (u - 1) & 1);  // O_g holds every half up to u - 1
          fence_after();
#pragma unroll
          for (int gg = 0; gg < DN / 16; ++gg) {
            uint32_t r[16];
            tmem_ld16_nowait(tO + gg * 16, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * corr);
            tmem_st16(tO + gg * 16, r);
          }
          tmem_st_wait();
        }
      }
      const float mb = m_used * sl2;
      const uint64_t sl2x2 = pack2(sl2, sl2), nmbx2 = pack2(-mb, -mb);
      uint64_t sum2[4] = {pack2(0.f, 0.f), pack2(0.f, 0.f), pack2(0.f, 0.f), pack2(0.f, 0.f)};
      uint32_t pk[2][16];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float a0, a1;
          unpack2(ffma2(pack2(__uint_as_float(v[c][i]), __uint_as_float(v[c][i + 1])), sl2x2, nmbx2), a0, a1);
          const float p0 = ex2(a0), p1 = ex2(a1);
          sum2[(i >> 1) & 3] = fadd2(sum2[(i >> 1) & 3], pack2(p0, p1));
          __nv_bfloat162 hh2 = __floats2bfloat162_rn(p0, p1);
          pk[c][i >> 1] = *reinterpret_cast<uint32_t*>(&hh2);
        }
      }
      if (u > 0) {  // this stream's previous PV has retired: P_g may be rewritten
        mbar_wait(pv_done + g, (u - 1) & 1);
        fence_after();
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(pw ^ ((uint32_t)(c * 4 + jj) << 4)), "r"(pk[c][4 * jj]),
                       "r"(pk[c][4 * jj + 1]), "r"(pk[c][4 * jj + 2]), "r"(pk[c][4 * jj + 3])
                       : "memory");
      }
      {
        float s0, s1, s2, s3;
        unpack2(fadd2(sum2[0], sum2[1]), s0, s1);
        unpack2(fadd2(sum2[2], sum2[3]), s2, s3);
        l += (s0 + s1) + (s2 + s3);
      }
      fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full + g);
    }
    // merge of the two streams' partial results.  The (max, sum) pairs travel through the stream's own P buffer, dead
    // after its last PV; with a single half stream 1 has no keys (O_1 was never written)
    const bool two = H > 1;
    float2* ml = reinterpret_cast<float2*>(Ps);  // [g * CH / 8 + row]
    if (u > 0) {
      mbar_wait(pv_done + g, (u - 1) & 1);  // this stream's last PV
      ml[g * (CH / 8) + row] = make_float2(m_used, l);
    }
    fence_before();
    asm volatile("bar.sync 1, 256;\n" ::: "memory");  // both warpgroups: every PV has retired, (max, sum) pairs are published
    fence_after();
    const float2 mo = (g == 0 && !two) ? make_float2(-INFINITY, 0.f) : ml[(g ^ 1) * (CH / 8) + row];
    const float m = fmaxf(m_used, mo.x);
    const float a_own = ex2((m_used - m) * sl2), a_oth = ex2((mo.x - m) * sl2);  // (-inf - m = -inf -> 0: a stream without keys)
    const float inv = 1.0f / (l * a_own + mo.y * a_oth);
    const float c_own = a_own * inv, c_oth = a_oth * inv;
    const uint32_t tO_oth = tmem_base + lane_base + (uint32_t)(2 * HKV + (g ^ 1) * 64);
    const int q = q0 + row;
    bf16* dst = o + ((int64_t)b * Nq + q) * ldo + h * d;
#pragma unroll
    for (int gg = 0; gg < DN / 16; ++gg) {
      if ((gg & 1) != g) continue;  // the two warpgroups share the store work by 16-channel groups
      uint32_t r0[16], r1[16];
      if (g == 0 || two) tmem_ld16_nowait(tO + gg * 16, r0);  // warp-collective: every lane loads, only real rows / channels store
      if (g == 1 || two) tmem_ld16_nowait(tO_oth + gg * 16, r1);
      tmem_ld_wait();
#pragma unroll
      for (int hlf = 0; hlf < 2; ++hlf) {
        if (q < Nq && gg * 16 + hlf * 8 < d) {
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float own = (g == 0 || two) ? __uint_as_float(r0[hlf * 8 + i]) * c_own : 0.f;
            const float oth = (g == 1 || two) ? __uint_as_float(r1[hlf * 8 + i]) * c_oth : 0.f;
            f[i] = own + oth;
          }
          store8(dst + gg * 16 + hlf * 8, f);
        }
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == W_ISSUE) {
    fence_after();
    tmem_dealloc<TCOLS>(tmem_base);
  }
}

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

// (head dim, heads, tokens, batch) view of a [batch * tokens, ld] matrix whose head h occupies columns [h*d, (h+1)*d)
int head_map(CUtensorMap* map, const bf16* base, int d, int heads, int ntok, int B, int ld, int box_rows = 128) {
  cuuint64_t dims[4] = {(cuuint64_t)d, (cuuint64_t)heads, (cuuint64_t)ntok, (cuuint64_t)B};
  cuuint64_t str[3] = {(cuuint64_t)d * 2, (cuuint64_t)ld * 2, (cuuint64_t)ntok * ld * 2};
  cuuint32_t box[4] = {64, 1, (cuuint32_t)box_rows, 1};
  return tma_encode_bf16(map, base, 4, dims, str, box);
}

template <int DN, int NWG, int KST, int VST>
int launch(const bf16* q, const bf16* k, const bf16* v, bf16* o, int B, int heads, int Nq, int Nkv, int d, int ldq, int ldk,
           int ldv, int ldo, float scale, cudaStream_t st) {
  using C = ACfg<DN, NWG, KST, VST>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_tcgen05_kernel<DN, NWG, KST, VST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    MKD_REQUIRE(e == cudaSuccess, MKD_E_CUDA, "attention_tcgen05: cudaFuncSetAttribute(%zu): %s", C::SMEM, cudaGetErrorString(e));
    configured = true;
  }
  CUtensorMap qm, km, vm;
  int rc;
  if ((rc = head_map(&qm, q, d, heads, Nq, B, ldq))) return rc;
  if ((rc = head_map(&km, k, d, heads, Nkv, B, ldk))) return rc;
  if ((rc = head_map(&vm, v, d, heads, Nkv, B, ldv))) return rc;
  dim3 grid((Nq + BQ * C::NWG - 1) / (BQ * C::NWG), heads, B);
  MKD_LAUNCH_OK(launch_pdl(attn_tcgen05_kernel<DN, NWG, KST, VST>, grid, dim3(C::THREADS), C::SMEM, st, qm, km, vm, o, Nq, Nkv, d, ldo,
                           scale * 1.4426950408889634f));
  MKD_CHECK_LAUNCH();
  return MKD_OK;
}

template <int DN>
int launch_half(const bf16* q, const bf16* k, const bf16* v, bf16* o, int B, int heads, int Nq, int Nkv, int d, int ldq, int ldk,
                int ldv, int ldo, float scale, cudaStream_t st) {
  constexpr size_t SMEM = (size_t)7 * CH + 17 * 8 + 16;  // Q, 2 x 2 K and V half tiles, 2 P buffers + barriers + TMEM slot
  static_assert(2 * SMEM <= 227 * 1024, "two CTAs per SM");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_split_kernel<DN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
    MKD_REQUIRE(e == cudaSuccess, MKD_E_CUDA, "attention_tcgen05: cudaFuncSetAttribute(%zu): %s", SMEM, cudaGetErrorString(e));
    configured = true;
  }
  CUtensorMap qm, km, vm;
  int rc;
  if ((rc = head_map(&qm, q, d, heads, Nq, B, ldq))) return rc;
  if ((rc = head_map(&km, k, d, heads, Nkv, B, ldk, HKV))) return rc;
  if ((rc = head_map(&vm, v, d, heads, Nkv, B, ldv, HKV))) return rc;
  dim3 grid((Nq + BQ - 1) / BQ, heads, B);
  MKD_LAUNCH_OK(launch_pdl(attn_split_kernel<DN>, grid, dim3(320), SMEM, st, qm, km, vm, o, Nq, Nkv, d, ldo, scale * 1.4426950408889634f));
  MKD_CHECK_LAUNCH();
  return MKD_OK;
}
}  // namespace

namespace mkd {
int tma_encode_bf16(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
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

bool attention_tcgen05_supported(int d, int ldq, int ldk, int ldv) {
  return d % 8 == 0 && d >= 8 && d <= 160 && ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0;
}

// bf16 attention on tcgen05; q / k / v / o are [B * N, ld] matrices, head h in columns [h*d, (h+1)*d)
int attention_tcgen05(const bf16* q, const bf16* k, const bf16* v, bf16* o, int B, int heads, int Nq, int Nkv, int d,
                      int ldq, int ldk, int ldv, int ldo, float scale, cudaStream_t st) {
  const int dn = (d + 15) / 16 * 16;
  static const int whole = debug_switch("MKD_ATTN_WHOLE", 0);  // 1: whole-tile kernel also for head dims <= 64 (A/B measurements)
#define MKD_ATTN_ARGS q, k, v, o, B, heads, Nq, Nkv, d, ldq, ldk, ldv, ldo, scale, st
  if (dn <= 64 && !whole) {  // half-tile kernel: two softmax warpgroups per 128 queries, two CTAs per SM
    if (dn <= 16) return launch_half<16>(MKD_ATTN_ARGS);
    if (dn <= 32) return launch_half<32>(MKD_ATTN_ARGS);
    if (dn <= 48) return launch_half<48>(MKD_ATTN_ARGS);
    return launch_half<64>(MKD_ATTN_ARGS);
  }
  if (dn <= 48) return launch<48, 1, 2, 2>(MKD_ATTN_ARGS);
  if (dn <= 64) return launch<64, 1, 2, 2>(MKD_ATTN_ARGS);
  if (dn <= 80) return launch<80, 2, 2, 1>(MKD_ATTN_ARGS);
  if (dn <= 128) return launch<128, 2, 2, 1>(MKD_ATTN_ARGS);
  return launch<160, 1, 2, 1>(MKD_ATTN_ARGS);
#undef MKD_ATTN_ARGS
}
}  // namespace mkd
