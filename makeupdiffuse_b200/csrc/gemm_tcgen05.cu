// NHWC implicit-GEMM convolution / GEMM on the 5th-generation tensor cores (sm_100a):
//   D[128 x BN] (fp32, TMEM) += A[128 x 64] (bf16, smem, K-major, 128B swizzle) * B[BN x 64]^T (bf16, smem, K-major)
//
//   * A (activations) arrives by TMA.  Plain GEMM / 1x1 conv: a 2-D map over [M, K] (row pitch ldx).  3x3 conv:
//     a 4-D map over the NHWC tensor (C, W, H, N) with box (64, Wb, Hb, Nb), Wb*Hb*Nb = 128 output pixels; filter
//     tap (r, s) is the same box shifted by (s - pad, r - pad) — out-of-bounds rows/columns (the padding halo, and
//     the M tail) are zero-filled by the TMA unit, so no im2col buffer and no halo code exist anywhere.
//   * B (weights, [K_out][R*S*C] "KRSC") arrives by a 2-D TMA map, box (64, BN).
//   * warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one lane; owns the TMEM allocation),
//     warps 2-5 = epilogue (tcgen05.ld -> registers -> fused bias / timestep-embedding / alpha / residual (incl. the
//     in-place ControlNet injection) / SiLU / GEGLU -> bf16 -> global, possibly into a channel slice of a concat buffer).
//   * STAGES-deep smem ring with full/empty mbarriers; tcgen05.commit releases a stage when its MMAs retire.
//   * split-K (gridDim.z > 1): each split writes fp32 partials to the workspace; splitk_epilogue_kernel reduces them
//     and applies the same epilogue.  Used where M*N tiles alone cannot fill 148 SMs (the 8x8 and 4x4 levels).
//   * 2 CTAs per SM (<= 113 KB smem, <= 256 TMEM columns each): one CTA's epilogue overlaps the other's main loop.
#include <cuda.h>

#include "common.cuh"
using namespace mkd;

namespace {

constexpr int BM = 128, BK = 64, UMMA_K = 16;
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
};

struct MainP {
  int kblocks;         // total K blocks (R*S*C / 64)
  int kb_per_split;
  int conv;            // 0: plain 2-D A map, 1: 4-D tap walk
  int cblocks;         // C / 64
  int S, pad;
  int Wb, Hb, Nb;      // box
  int tiles_w, tiles_h;
};

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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
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
__host__ __device__ constexpr uint32_t make_idesc(int bn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__host__ __device__ constexpr int tmem_cols(int bn) { return bn <= 32 ? 32 : bn <= 64 ? 64 : bn <= 128 ? 128 : 256; }

// ---- shared epilogue math (used by the main kernel and by the split-K reducer) ------------------------------
// v[16] = accumulators of columns [n, n+16) of row m (n is a *weight-row* index).
__device__ __forceinline__ void epilogue_store16(const EpiP& e, int m, int n, float (&v)[16]) {
  const int nimg = e.emb ? m / e.pix_per_img : 0;
#pragma unroll
  for (int j = 0; j < 16; j += 8) {
    const int o = n + j;
    if (o >= e.N_out) return;
    float r[8];
    if (o + 8 <= e.N_out) {
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = v[j + i] + (e.bias ? e.bias[o + i] : 0.f);
      if (e.emb) {
        float t[8];
        load8(e.emb + (int64_t)nimg * e.lde + o, t);
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
      for (int i = 0; i < 8 && o + i < e.N_out; ++i) {
        float t = v[j + i] + (e.bias ? e.bias[o + i] : 0.f);
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
// GEGLU: a tile of BN weight rows = BN/2 value rows then BN/2 gate rows; val/gate are 16 matching columns.
__device__ __forceinline__ void epilogue_geglu16(const EpiP& e, int m, int row_val, int row_gate, int o, float (&a)[16],
                                                 float (&g)[16]) {
#pragma unroll
  for (int j = 0; j < 16; j += 8) {
    float r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float av = a[j + i] + (e.bias ? e.bias[row_val + j + i] : 0.f);
      float gv = g[j + i] + (e.bias ? e.bias[row_gate + j + i] : 0.f);
      r[i] = av * gelu_erf_f(gv);
    }
    store8(e.y + (int64_t)m * e.ldy + o + j, r);
  }
}

// ---- main kernel -----------------------------------------------------------------------------------------
template <int BN, int STAGES>
__global__ void __launch_bounds__(192) gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap amap,
                                                           const __grid_constant__ CUtensorMap bmap, MainP mp, EpiP ep) {
  constexpr int B_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int TCOLS = tmem_cols(BN);
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tile = blockIdx.x, n_tile = blockIdx.y, split = blockIdx.z;
  const int kb0 = split * mp.kb_per_split;
  const int kb1 = min(mp.kblocks, kb0 + mp.kb_per_split);
  const int nkb = kb1 - kb0;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&amap)) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&bmap)) : "memory");
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(full_bar + i, 1);
      mbar_init(empty_bar + i, 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {  // TMEM allocation (this warp also frees it)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(TCOLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int w0 = 0, h0 = 0, n0 = 0;
      if (mp.conv) {
        int tw = m_tile % mp.tiles_w, th = (m_tile / mp.tiles_w) % mp.tiles_h, tn = m_tile / (mp.tiles_w * mp.tiles_h);
        w0 = tw * mp.Wb;
        h0 = th * mp.Hb;
        n0 = tn * mp.Nb;
      }
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES, ph = (i / STAGES) & 1;
        mbar_wait(empty_bar + s, ph ^ 1);
        mbar_expect_tx(full_bar + s, STAGE_BYTES);
        unsigned char* sa = smem + s * STAGE_BYTES;
        const int kb = kb0 + i;
        if (mp.conv) {
          const int tap = kb / mp.cblocks, cb = kb - tap * mp.cblocks;
          const int r = tap / mp.S, sx = tap - r * mp.S;
          tma_load_4d(&amap, full_bar + s, sa, cb * BK, w0 + sx - mp.pad, h0 + r - mp.pad, n0);
        } else {
          tma_load_2d(&amap, full_bar + s, sa, kb * BK, m_tile * BM);
        }
        tma_load_2d(&bmap, full_bar + s, sa + A_BYTES, kb * BK, n_tile * BN);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES, ph = (i / STAGES) & 1;
        mbar_wait(full_bar + s, ph);
        tcgen05_fence_after();
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
        const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sa + A_BYTES);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // advance K inside the 128-byte swizzle atom: +32 bytes per UMMA_K (encoded >> 4)
          umma_bf16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (i | k) ? 1u : 0u);
        }
        tcgen05_commit(empty_bar + s);  // frees this smem stage once the MMAs above retire
      }
      tcgen05_commit(tmem_full_bar);    // accumulator complete
    }
  } else {
    // ===== epilogue warps 2..5: TMEM lane quadrant = warp % 4 =====
    const int quad = warp & 3;
    const int m = m_tile * BM + quad * 32 + lane;
    mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();
    const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
    const bool mvalid = m < ep.M;
    if (ep.partial) {
      float* prow = ep.partial + ((int64_t)split * ep.M + m) * ep.n_rows + n_tile * BN;
#pragma unroll 1
      for (int c = 0; c < BN; c += 16) {
        float v[16];
        tmem_ld16(trow + c, v);
        if (mvalid && n_tile * BN + c < ep.n_rows) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(prow + c + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
      }
    } else if (ep.act == MKD_ACT_GEGLU) {
#pragma unroll 1
      for (int c = 0; c < BN / 2; c += 16) {
        float a[16], g[16];
        tmem_ld16(trow + c, a);
        tmem_ld16(trow + BN / 2 + c, g);
        if (mvalid) epilogue_geglu16(ep, m, n_tile * BN + c, n_tile * BN + BN / 2 + c, n_tile * (BN / 2) + c, a, g);
      }
    } else {
#pragma unroll 1
      for (int c = 0; c < BN; c += 16) {
        float v[16];
        tmem_ld16(trow + c, v);
        if (mvalid) epilogue_store16(ep, m, n_tile * BN + c, v);
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(TCOLS));
  }
}

// split-K reducer: sums the fp32 partials and applies the epilogue; one thread per (row, 16 weight rows).
template <int BN>
__global__ void splitk_epilogue_kernel(EpiP ep, int splits) {
  const int groups = ep.n_rows / 16;
  const int64_t total = (int64_t)ep.M * groups;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int m = (int)(i / groups), n = (int)(i % groups) * 16;
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
      epilogue_geglu16(ep, m, n, n + BN / 2, tile * (BN / 2) + c, a, g);
    } else {
      float v[16];
      gather(n, v);
      epilogue_store16(ep, m, n, v);
    }
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
  if (d->stride != 1 || d->upsample) { set_error("stride/upsample convs use the generic kernel"); return false; }
  if (d->R != d->S || (d->R != 1 && d->R != 3) || d->pad != d->R / 2) { set_error("filter is not 1x1/p0 or 3x3/p1"); return false; }
  if (d->ldx % 8 || !aligned16(d->x) || !aligned16(d->w)) { set_error("x/w alignment"); return false; }
  if (d->y && (d->ldy % 8 || !aligned16(d->y))) { set_error("y alignment"); return false; }
  if (d->y32 && (d->ldy32 % 8 || !aligned16(d->y32))) { set_error("y32 alignment"); return false; }
  if (d->act == MKD_ACT_GEGLU && !d->y) { set_error("GEGLU writes the bf16 output only"); return false; }
  if (d->residual && (d->ldr % 8 || !aligned16(d->residual))) { set_error("residual alignment"); return false; }
  if (d->emb && (d->lde % 8 || !aligned16(d->emb))) { set_error("emb alignment"); return false; }
  g.conv = d->R == 3;
  g.P = d->H; g.Q = d->W;
  g.M = d->N * d->H * d->W;
  g.Ktot = d->R * d->S * d->C;
  g.Kout = d->act == MKD_ACT_GEGLU ? d->K / 2 : d->K;
  if (d->K % 16 != 0 && !(d->K < 16)) { set_error("K=%d is not a multiple of 16", d->K); return false; }
  if (d->act == MKD_ACT_GEGLU && (d->geglu_block != 80 || d->K % 160)) { set_error("GEGLU needs geglu_block 80 and K %% 160 == 0"); return false; }
  if (g.conv) {
    if (!is_pow2(d->W) || !is_pow2(d->H)) { set_error("conv H/W must be powers of two"); return false; }
    g.Wb = d->W < BM ? d->W : BM;
    g.Hb = BM / g.Wb < d->H ? BM / g.Wb : d->H;
    g.Nb = BM / (g.Wb * g.Hb);
    g.tiles_w = d->W / g.Wb;
    g.tiles_h = d->H / g.Hb;
    g.m_tiles = g.tiles_w * g.tiles_h * ((d->N + g.Nb - 1) / g.Nb);
  } else {
    g.Wb = g.Hb = g.Nb = g.tiles_w = g.tiles_h = 1;
    g.m_tiles = (g.M + BM - 1) / BM;
  }
  return true;
}

int pick_bn(const mkd_conv_desc* d) {
  if (d->act == MKD_ACT_GEGLU) return 160;
  if (d->K % 160 == 0) return 160;
  if (d->K % 80 == 0) return 80;
  if (d->K % 128 == 0) return 128;
  if (d->K % 64 == 0) return 64;
  if (d->K <= 32) return 32;
  return 64;  // ragged last tile: B rows beyond K are zero-filled by TMA, stores are masked
}

template <int BN, int STAGES>
int launch(const mkd_conv_desc* d, const Geometry& g, cudaStream_t stream) {
  constexpr int STAGE_BYTES = A_BYTES + BN * BK * 2;
  constexpr size_t smem = (size_t)STAGES * STAGE_BYTES + (2 * STAGES + 1) * 8 + 16 + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    MKD_REQUIRE(e == cudaSuccess, MKD_E_CUDA, "gemm_tcgen05: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
    configured = true;
  }
  CUtensorMap amap, bmap;
  int rc;
  if (g.conv) {
    cuuint64_t dims[4] = {(cuuint64_t)d->C, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
    cuuint64_t str[3] = {(cuuint64_t)d->ldx * 2, (cuuint64_t)d->ldx * 2 * d->W, (cuuint64_t)d->ldx * 2 * d->W * d->H};
    cuuint32_t box[4] = {BK, (cuuint32_t)g.Wb, (cuuint32_t)g.Hb, (cuuint32_t)g.Nb};
    rc = encode(&amap, d->x, 4, dims, str, box);
  } else {
    cuuint64_t dims[2] = {(cuuint64_t)d->C, (cuuint64_t)g.M};
    cuuint64_t str[1] = {(cuuint64_t)d->ldx * 2};
    cuuint32_t box[2] = {BK, BM};
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
  // split-K: only when the tile grid leaves most SMs idle and K is deep
  int splits = 1;
  const int tiles = g.m_tiles * n_tiles;
  if (d->workspace && tiles < 148 && mp.kblocks >= 8 && d->K % 16 == 0) {
    splits = (2 * 148 + tiles - 1) / tiles;              // aim at ~2 CTAs per SM
    if (splits > mp.kblocks / 4) splits = mp.kblocks / 4;  // >= 4 K blocks per split
    if (splits > 32) splits = 32;
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

  dim3 grid(g.m_tiles, n_tiles, splits);
  MKD_REQUIRE(grid.y <= 65535 && grid.z <= 65535, MKD_E_INVALID, "gemm_tcgen05: grid too large");
  gemm_tcgen05_kernel<BN, STAGES><<<grid, 192, smem, stream>>>(amap, bmap, mp, ep);
  MKD_CHECK_LAUNCH();
  if (splits > 1) {
    int64_t total = (int64_t)g.M * (d->K / 16);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    splitk_epilogue_kernel<BN><<<blocks, 256, 0, stream>>>(ep, splits);
    MKD_CHECK_LAUNCH();
  }
  return MKD_OK;
}
}  // namespace

namespace mkd {
bool conv2d_tcgen05_supported(const mkd_conv_desc* d) {
  Geometry g;
  return geometry(d, g);
}

int conv2d_tcgen05(const mkd_conv_desc* d, cudaStream_t stream) {
  Geometry g;
  MKD_REQUIRE(geometry(d, g), MKD_E_INVALID, "gemm_tcgen05: unsupported shape");
  switch (pick_bn(d)) {
    case 160: return launch<160, 3>(d, g, stream);
    case 128: return launch<128, 3>(d, g, stream);
    case 80: return launch<80, 4>(d, g, stream);
    case 64: return launch<64, 4>(d, g, stream);
    default: return launch<32, 4>(d, g, stream);
  }
}
}  // namespace mkd
