// SpatialTransformer attention (self: Nkv = Nq = H*W; cross: Nkv = 77 text tokens), flash-style:
// the Nq x Nkv score matrix never leaves the SM, softmax statistics are fp32 (upstream _ATTN_PRECISION fp32).
//
//   attention_tcgen05.cu bf16 production kernels (tcgen05 / TMEM / TMA); this file dispatches to them.
//   attn_simt_kernel<T>  one warp per query: the fp32 check mode, head dims > 160 (the VAE's 512-wide head), CLIP's causal mask.
//   attn_mma_kernel<DP>  round-0 warp-level kernel (mma.sync m16n8k16): compiled only into MKD_TRACE=1 builds, for A/B runs.
#include <stdlib.h>

#include "common.cuh"
using namespace mkd;

namespace {

// -------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void attn_simt_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                                 T* __restrict__ o, int Nq, int Nkv_all, int d, int ldq, int ldk, int ldv, int ldo,
                                 float scale, int causal) {
  pdl_wait();
  extern __shared__ float sm[];
  const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, i = blockIdx.x * warps + warp;
  float* qs = sm + (size_t)warp * (d + Nkv_all);
  // causal (the CLIP text encoder): query i sees keys 0..i, i.e. the row simply has fewer keys
  const int Nkv = causal ? min(Nkv_all, i + 1) : Nkv_all;
  float* ps = qs + d;
  if (i >= Nq) return;
  const T* qrow = q + ((int64_t)b * Nq + i) * ldq + h * d;
  for (int c = lane; c < d; c += 32) qs[c] = to_f(qrow[c]) * scale;
  __syncwarp();
  float mx = -INFINITY;
  for (int j = lane; j < Nkv; j += 32) {
    const T* krow = k + ((int64_t)b * Nkv_all + j) * ldk + h * d;
    float acc = 0.f;
    for (int c = 0; c < d; c += 8) {
      float kv[8];
      load8(krow + c, kv);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc = fmaf(qs[c + e], kv[e], acc);
    }
    ps[j] = acc;
    mx = fmaxf(mx, acc);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < Nkv; j += 32) {
    float e = __expf(ps[j] - mx);
    ps[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  const float inv = 1.0f / sum;
  T* orow = o + ((int64_t)b * Nq + i) * ldo + h * d;
  for (int c = lane; c < d; c += 32) {
    float acc = 0.f;
    const T* vcol = v + (int64_t)b * Nkv_all * ldv + h * d + c;
    for (int j = 0; j < Nkv; ++j) acc = fmaf(ps[j], to_f(vcol[(int64_t)j * ldv]), acc);
    orow[c] = from_f<T>(acc * inv);
  }
}

#ifdef MKD_ENABLE_TRACE  // the round-0 warp-level mma.sync kernel: A/B builds only (MKD_ATTN_MMA=1); the shipped library has no such path
// -------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  int sz = valid ? 16 : 0;  // src-size 0 => 16 bytes of zeros
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(dst)), "l"(src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

constexpr int AQ = 64, AK = 64;  // queries per CTA, keys per tile

// load `rows` x d bf16 (row stride ld) into smem [64][DP+8], zero-filling rows >= nrows and cols >= d
template <int DP>
__device__ __forceinline__ void load_tile(bf16* dst, const bf16* src, int nrows, int d, int ld) {
  constexpr int PITCH = DP + 8, CH = DP / 8;
  for (int idx = threadIdx.x; idx < 64 * CH; idx += 128) {
    int r = idx / CH, c = (idx % CH) * 8;
    bool ok = r < nrows && c < d;
    cp_async16(dst + r * PITCH + c, ok ? src + (int64_t)r * ld + c : src, ok);
  }
}

template <int DP>
__global__ void __launch_bounds__(128) attn_mma_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k,
                                                       const bf16* __restrict__ v, bf16* __restrict__ o, int Nq,
                                                       int Nkv, int d, int ldq, int ldk, int ldv, int ldo,
                                                       float scale_log2e) {
  pdl_wait();
  constexpr int PITCH = DP + 8, KS = DP / 16, NT = DP / 8;
  extern __shared__ __align__(16) unsigned char smraw[];
  bf16* Qs = reinterpret_cast<bf16*>(smraw);
  bf16* Ks = Qs + AQ * PITCH;       // [2][AK][PITCH]
  bf16* Vs = Ks + 2 * AK * PITCH;   // [2][AK][PITCH]
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * AQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bf16* qb = q + ((int64_t)b * Nq + q0) * ldq + h * d;
  const bf16* kb = k + (int64_t)b * Nkv * ldk + h * d;
  const bf16* vb = v + (int64_t)b * Nkv * ldv + h * d;
  const int ntiles = (Nkv + AK - 1) / AK;

  load_tile<DP>(Qs, qb, min(AQ, Nq - q0), d, ldq);
  load_tile<DP>(Ks, kb, min(AK, Nkv), d, ldk);
  load_tile<DP>(Vs, vb, min(AK, Nkv), d, ldv);
  cp_async_commit();

  float oacc[NT][4];
#pragma unroll
  for (int i = 0; i < NT; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) oacc[i][j] = 0.f;
  float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};
  uint32_t qf[KS][4];

  for (int t = 0; t < ntiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < ntiles) {  // prefetch next K/V tile into the other buffer
      const int kv1 = (t + 1) * AK;
      load_tile<DP>(Ks + (buf ^ 1) * AK * PITCH, kb + (int64_t)kv1 * ldk, min(AK, Nkv - kv1), d, ldk);
      load_tile<DP>(Vs + (buf ^ 1) * AK * PITCH, vb + (int64_t)kv1 * ldv, min(AK, Nkv - kv1), d, ldv);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (t == 0) {
#pragma unroll
      for (int ks = 0; ks < KS; ++ks)
        ldsm_x4(qf[ks], Qs + (warp * 16 + (lane & 15)) * PITCH + ks * 16 + (lane >> 4) * 8);
    }
    const bf16* Kt = Ks + buf * AK * PITCH;
    const bf16* Vt = Vs + buf * AK * PITCH;

    // S = Q K^T   (16 x 64 per warp)
    float s[AK / 8][4];
#pragma unroll
    for (int i = 0; i < AK / 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int np = 0; np < AK / 16; ++np) {
        uint32_t bfr[4];
        // lanes 0-7: keys np*16+0..7 @ col ks*16 ; 8-15: same keys @ +8 ; 16-23: keys +8..15 @ col ; 24-31: @ +8
        int row = np * 16 + (lane & 7) + ((lane >> 4) << 3);
        int col = ks * 16 + ((lane >> 3) & 1) * 8;
        ldsm_x4(bfr, Kt + row * PITCH + col);
        mma16816(s[2 * np], qf[ks], bfr[0], bfr[1]);
        mma16816(s[2 * np + 1], qf[ks], bfr[2], bfr[3]);
      }
    }
    // online softmax (rows g = lane/4 and g+8 of this warp's 16)
    const int kv0 = t * AK;
    float mnew[2] = {mrow[0], mrow[1]};
#pragma unroll
    for (int nt = 0; nt < AK / 8; ++nt) {
      int key = kv0 + nt * 8 + (lane & 3) * 2;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float val = (key + (j & 1) < Nkv) ? s[nt][j] * scale_log2e : -INFINITY;
        s[nt][j] = val;
        mnew[j >> 1] = fmaxf(mnew[j >> 1], val);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mnew[r] = fmaxf(mnew[r], __shfl_xor_sync(0xffffffffu, mnew[r], 1));
      mnew[r] = fmaxf(mnew[r], __shfl_xor_sync(0xffffffffu, mnew[r], 2));
    }
    float corr[2], psum[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      corr[r] = exp2f(mrow[r] - mnew[r]);  // first tile: exp2(-inf) = 0
      mrow[r] = mnew[r];
    }
#pragma unroll
    for (int nt = 0; nt < AK / 8; ++nt)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float e = exp2f(s[nt][j] - mnew[j >> 1]);
        s[nt][j] = e;
        psum[j >> 1] += e;
      }
#pragma unroll
    for (int r = 0; r < 2; ++r) lrow[r] = lrow[r] * corr[r] + psum[r];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      oacc[nt][0] *= corr[0];
      oacc[nt][1] *= corr[0];
      oacc[nt][2] *= corr[1];
      oacc[nt][3] *= corr[1];
    }
    // O += P V
#pragma unroll
    for (int kk = 0; kk < AK / 16; ++kk) {
      uint32_t pa[4];
      pa[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {
        uint32_t bfr[4];
        // lanes 0-7: keys kk*16+0..7 @ dcol np*16 ; 8-15: keys +8..15 @ same ; 16-23: keys 0..7 @ +8 ; 24-31: keys +8 @ +8
        int row = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        int col = np * 16 + (lane >> 4) * 8;
        ldsm_x4_t(bfr, Vt + row * PITCH + col);
        mma16816(oacc[2 * np], pa, bfr[0], bfr[1]);
        mma16816(oacc[2 * np + 1], pa, bfr[2], bfr[3]);
      }
    }
    __syncthreads();  // all warps done with this buffer before it is refilled
  }

  // finalise: quad-reduce the row sums, normalise, store
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 1);
    lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 2);
  }
  const float inv0 = 1.0f / lrow[0], inv1 = 1.0f / lrow[1];
  const int r0 = q0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
  bf16* ob = o + (int64_t)b * Nq * ldo + h * d;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    int c = nt * 8 + (lane & 3) * 2;
    if (c < d) {
      if (r0 < Nq)
        *reinterpret_cast<__nv_bfloat162*>(ob + (int64_t)r0 * ldo + c) = __floats2bfloat162_rn(oacc[nt][0] * inv0, oacc[nt][1] * inv0);
      if (r1 < Nq)
        *reinterpret_cast<__nv_bfloat162*>(ob + (int64_t)r1 * ldo + c) = __floats2bfloat162_rn(oacc[nt][2] * inv1, oacc[nt][3] * inv1);
    }
  }
}

template <int DP>
int launch_mma(const bf16* q, const bf16* k, const bf16* v, bf16* o, int B, int heads, int Nq, int Nkv, int d, int ldq,
               int ldk, int ldv, int ldo, float scale, cudaStream_t st) {
  constexpr size_t smem = (size_t)(AQ + 4 * AK) * (DP + 8) * sizeof(bf16);
  static bool configured = false;  // benign race: attribute set is idempotent
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_mma_kernel<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    MKD_REQUIRE(e == cudaSuccess, MKD_E_CUDA, "attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  dim3 grid((Nq + AQ - 1) / AQ, heads, B);
  MKD_LAUNCH_OK(launch_pdl(attn_mma_kernel<DP>, dim3(grid), dim3(128), smem, st, q, k, v, o, Nq, Nkv, d, ldq, ldk, ldv, ldo, scale * 1.4426950408889634f));
  MKD_CHECK_LAUNCH();
  return MKD_OK;
}
#endif  // MKD_ENABLE_TRACE
#ifdef MKD_ENABLE_TRACE
bool attn_force_mma() {
  static const int v = debug_switch("MKD_ATTN_MMA", 0);  // trace builds: 1 = the mma.sync kernel for every head dim (A/B runs)
  return v == 1;
}
#endif
}  // namespace

extern "C" int mkd_attention(const void* q, const void* k, const void* v, void* o, int dtype, int B, int heads, int Nq,
                             int Nkv, int d, int ldq, int ldk, int ldv, int ldo, float scale, mkd_stream_t stream) {
  MKD_REQUIRE(q && k && v && o && B > 0 && heads > 0 && Nq > 0 && Nkv > 0 && d > 0, MKD_E_INVALID, "attention: bad args");
  // head dims up to 160 (the SpatialTransformer's 40 / 80 / 160) run on the tensor cores; wider single heads (the VAE
  // decoder's 512-wide mid.attn_1) take the one-warp-per-query SIMT kernel
  MKD_REQUIRE(d % 8 == 0 && d <= 512, MKD_E_INVALID, "attention: head dim %d must be a multiple of 8, <= 512", d);
  MKD_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && aligned16(q) && aligned16(k) &&
                  aligned16(v) && aligned16(o),
              MKD_E_ALIGN, "attention: ld must be multiples of 8 and pointers 16B aligned");
  MKD_REQUIRE(B <= 65535 && heads <= 65535, MKD_E_INVALID, "attention: B/heads too large");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MKD_BF16 && d <= 160) {
    const bf16 *qq = (const bf16*)q, *kk = (const bf16*)k, *vv = (const bf16*)v;
    bf16* oo = (bf16*)o;
    // the tcgen05 / TMEM / TMA flash attention (attention_tcgen05.cu) takes every head dim admitted above
#ifdef MKD_ENABLE_TRACE
    if (attn_force_mma()) {  // A/B builds: the older warp-level mma.sync kernel
      if (d <= 16) return launch_mma<16>(qq, kk, vv, oo, B, heads, Nq, Nkv, d, ldq, ldk, ldv, ldo, scale, st);
      if (d <= 32) return launch_mma<32>(qq, kk, vv, oo, B, heads, Nq, Nkv, d, ldq, ldk, ldv, ldo, scale, st);
      if (d <= 48) return launch_mma<48>(qq, kk, vv, oo, B, heads, Nq, Nkv, d, ldq, ldk, ldv, ldo, scale, st);
      if (d <= 64) return launch_mma<64>(qq, kk, vv, oo, B, heads, Nq, Nkv, d, ldq, ldk, ldv, ldo, scale, st);
      if (d <= 80) return launch_mma<80>(qq, kk, vv, oo, B, heads, Nq, Nkv, d, ldq, ldk, ldv, ldo, scale, st);
      return launch_mma<160>(qq, kk, vv, oo, B, heads, Nq, Nkv, d, ldq, ldk, ldv, ldo, scale, st);
    }
#endif
    MKD_REQUIRE(attention_tcgen05_supported(d, ldq, ldk, ldv), MKD_E_INVALID, "attention: head dim %d not supported on the tensor-core kernel", d);
    return attention_tcgen05(qq, kk, vv, oo, B, heads, Nq, Nkv, d, ldq, ldk, ldv, ldo, scale, st);
  }
  const int warps = 4;
  size_t smem = (size_t)warps * (d + Nkv) * sizeof(float);
  MKD_REQUIRE(smem <= 200 * 1024, MKD_E_INVALID, "attention(fp32 check kernel): Nkv=%d too large", Nkv);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_simt_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_simt_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    MKD_REQUIRE(e == cudaSuccess, MKD_E_CUDA, "attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  dim3 grid((Nq + warps - 1) / warps, heads, B);
  if (dtype == MKD_BF16)
    MKD_LAUNCH_OK(launch_pdl(attn_simt_kernel<bf16>, dim3(grid), dim3(warps * 32), smem, st, (const bf16*)q, (const bf16*)k, (const bf16*)v, (bf16*)o, Nq,
                             Nkv, d, ldq, ldk, ldv, ldo, scale, 0));
  else
    MKD_LAUNCH_OK(launch_pdl(attn_simt_kernel<float>, dim3(grid), dim3(warps * 32), smem, st, (const float*)q, (const float*)k, (const float*)v, (float*)o, Nq,
                             Nkv, d, ldq, ldk, ldv, ldo, scale, 0));
  MKD_CHECK_LAUNCH();
  return MKD_OK;
}

// Causal self-attention of short sequences (the 77-token CLIP text encoder, makeup_controlnet.py:20 / yaml:109-110):
// one warp per query row on the SIMT kernel, key j visible to query i iff j <= i.  Runs once per prompt, not per step.
extern "C" int mkd_attention_causal(const void* q, const void* k, const void* v, void* o, int dtype, int B, int heads,
                                    int N, int d, int ldq, int ldk, int ldv, int ldo, float scale, mkd_stream_t stream) {
  MKD_REQUIRE(q && k && v && o && B > 0 && heads > 0 && N > 0 && d > 0, MKD_E_INVALID, "attention_causal: bad args");
  MKD_REQUIRE(d % 8 == 0 && d <= 512 && B <= 65535 && heads <= 65535, MKD_E_INVALID, "attention_causal: head dim %d / batch", d);
  MKD_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && aligned16(q) && aligned16(k) &&
                  aligned16(v) && aligned16(o),
              MKD_E_ALIGN, "attention_causal: ld must be multiples of 8 and pointers 16B aligned");
  const int warps = 4;
  const size_t smem = (size_t)warps * (d + N) * sizeof(float);
  MKD_REQUIRE(smem <= 48 * 1024, MKD_E_INVALID, "attention_causal: N=%d too long for this kernel", N);
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((N + warps - 1) / warps, heads, B);
  if (dtype == MKD_BF16)
    MKD_LAUNCH_OK(launch_pdl(attn_simt_kernel<bf16>, dim3(grid), dim3(warps * 32), smem, st, (const bf16*)q, (const bf16*)k, (const bf16*)v, (bf16*)o, N,
                             N, d, ldq, ldk, ldv, ldo, scale, 1));
  else
    MKD_LAUNCH_OK(launch_pdl(attn_simt_kernel<float>, dim3(grid), dim3(warps * 32), smem, st, (const float*)q, (const float*)k, (const float*)v, (float*)o, N,
                             N, d, ldq, ldk, ldv, ldo, scale, 1));
  MKD_CHECK_LAUNCH();
  return MKD_OK;
}
