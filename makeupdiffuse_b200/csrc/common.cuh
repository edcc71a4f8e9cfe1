// Shared device/host helpers for libmkd_b200 (sm_100a only).
#pragma once
#include <stdlib.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mkd_b200.h"

namespace mkd {

void set_error(const char* fmt, ...);
void count_launch();

#define MKD_REQUIRE(cond, code, ...)      \
  do {                                    \
    if (!(cond)) {                        \
      ::mkd::set_error(__VA_ARGS__);      \
      return (code);                      \
    }                                     \
  } while (0)

#define MKD_CHECK_LAUNCH()                                                      \
  do {                                                                          \
    cudaError_t e__ = cudaGetLastError();                                       \
    if (e__ != cudaSuccess) {                                                   \
      ::mkd::set_error("%s:%d launch failed: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return MKD_E_CUDA;                                                        \
    }                                                                           \
    ::mkd::count_launch();                                                      \
  } while (0)

// launch_pdl() result check (MKD_CHECK_LAUNCH() that follows counts the launch)
#define MKD_LAUNCH_OK(expr)                                                                        \
  do {                                                                                             \
    cudaError_t le__ = (expr);                                                                     \
    if (le__ != cudaSuccess) {                                                                     \
      ::mkd::set_error("%s:%d launch failed: %s", __FILE__, __LINE__, cudaGetErrorString(le__));   \
      return MKD_E_CUDA;                                                                           \
    }                                                                                              \
  } while (0)

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// 8 consecutive elements <-> 8 floats (16 B for bf16, 32 B for fp32); pointers must be 16 B aligned.
__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8(bf16* p, const float (&v)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
// erf-GELU for the bf16 path's fused GEGLU epilogue: erf by Abramowitz-Stegun 7.1.26 (|erf error| <= 1.5e-7, i.e. a
// GELU error <= 0.75e-7 |x| — four orders below the bf16 rounding of the result) in ~16 instructions, 2 of them MUFU,
// where erff() costs ~40 and made the FF1 epilogue issue-bound (profiles/r01_ncu_full_ff1.txt).  The fp32 check mode
// keeps erff().
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;\n" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(e) : "f"(-1.4426950408889634f * z * z));
  float p = fmaf(t, 1.061405429f, -1.453152027f);
  p = fmaf(t, p, 1.421413741f);
  p = fmaf(t, p, -0.284496736f);
  p = fmaf(t, p, 0.254829592f);
  const float erf_abs = fmaf(-p * t, e, 1.0f);  // erf(|x| / sqrt 2)
  return 0.5f * x * (1.0f + copysignf(erf_abs, x));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------------
// Every kernel of the library is launched with cudaLaunchAttributeProgrammaticStreamSerialization and executes
// pdl_wait() before its first global-memory access: the next kernel's CTAs are scheduled (launch ramp, smem carve-out,
// barrier init, TMEM allocation, tensor-map prefetch) while the previous kernel drains, then block in
// griddepcontrol.wait until that kernel has completed and flushed.  One UNet+ControlNet step is ~500 dependent
// launches of 5-50 us.  Rule that keeps it correct transitively: EVERY kernel waits, on every path, before it reads or
// writes global memory.  ON by default since round 2 (6.97 -> 6.74 ms per step; it had measured slightly negative in round 1,
// before the launch count fell and the GEMM prologues grew a cluster rendezvous); without the attribute griddepcontrol.* are no-ops.
__device__ __forceinline__ void pdl_wait() {
  asm volatile("griddepcontrol.wait;\n" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
}
bool pdl_enabled();
// L2 prefetch of this CTA's share of a launch's weights, issued by one warp BEFORE pdl_wait(): weights are cold in HBM at
// every layer (2.4 GB per step against 126 MB of L2), the prefetch touches no SM state, and it is harmless whatever the
// previous kernel is still writing (L2 is the coherence point).  So the weight stream starts while the previous kernel
// drains instead of after the first full-barrier wait of the main loop.  Capped: activations should stay L2-resident.
// `share` = bytes per CTA, computed on the host (l2_prefetch_share_bytes): the 64-bit divisions this took on the device
// (~0.5 us in front of the prologue's first cluster barrier, where warp 3's arrival is waited for) are gone.
__device__ __forceinline__ void l2_prefetch_share(const void* base, unsigned long long bytes, unsigned share, int lane) {
  constexpr unsigned CHUNK = 8192;
  const unsigned long long lo = (unsigned long long)blockIdx.x * share;
  if (lo >= bytes) return;
  const unsigned span = (unsigned)(bytes - lo < share ? bytes - lo : share);
  for (unsigned o = (unsigned)lane * CHUNK; o < span; o += 32 * CHUNK) {
    const unsigned n = span - o < CHUNK ? span - o : CHUNK;
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(reinterpret_cast<const char*>(base) + lo + o), "r"(n) : "memory");
  }
}
// host: caps the prefetched bytes (activations should stay L2-resident) and splits them over the grid in whole 8 KB chunks
inline void l2_prefetch_plan(unsigned long long w_bytes, int grid, unsigned long long& bytes, unsigned& share) {
  constexpr unsigned long long CAP = 48ull << 20, CHUNK = 8192;
  bytes = w_bytes > CAP ? CAP : w_bytes;
  share = (unsigned)(((bytes + grid - 1) / grid + CHUNK - 1) / CHUNK * CHUNK);
}
// Experiment switches (A/B runs, timing experiments) exist only in trace builds (-DMKD_ENABLE_TRACE, i.e.
// `MKD_TRACE=1 python -m makeupdiffuse_b200.build --force`); the shipped library has ONE code path and reads no environment.
#ifdef MKD_ENABLE_TRACE
inline int debug_switch(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
#else
constexpr int debug_switch(const char*, int dflt) { return dflt; }
#endif

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int dtype_size(int dtype) { return dtype == MKD_BF16 ? 2 : 4; }
inline int num_sms() { return 148; }  // B200

// entry points implemented in the individual translation units
int conv2d_generic(const mkd_conv_desc* d, cudaStream_t stream);
int conv2d_tcgen05(const mkd_conv_desc* d, cudaStream_t stream);
bool conv2d_tcgen05_supported(const mkd_conv_desc* d);
bool attention_tcgen05_supported(int d, int ldq, int ldk, int ldv);
int attention_tcgen05(const bf16* q, const bf16* k, const bf16* v, bf16* o, int B, int heads, int Nq, int Nkv, int d,
                      int ldq, int ldk, int ldv, int ldo, float scale, cudaStream_t st);

}  // namespace mkd
