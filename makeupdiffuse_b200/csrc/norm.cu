// GroupNorm(+SiLU) and LayerNorm on NHWC activations: bandwidth kernels.
//   - every global access is a 16-byte (bf16) / 32-byte (fp32) vector of 8 consecutive channels, so a warp reads
//     512 contiguous bytes of a pixel row;
//   - statistics are fp32; the cross-thread reductions are warp shuffles + one small smem pass;
//   - GroupNorm is ONE launch: a thread-block cluster (<= 8 CTAs) per sample; each CTA reduces its pixel chunk, the
//     per-group partial sums are exchanged through distributed shared memory, then every CTA normalises its own chunk
//     (the second read of x is served by L1/L2: the chunk was just touched by the same SM).  A two-launch variant
//     (partials through global memory) covers the cases a cluster cannot: tiny batches that need > 8 chunks per
//     sample to fill the GPU.
#include <cooperative_groups.h>

#include "common.cuh"
using namespace mkd;
namespace cg = cooperative_groups;

namespace {
constexpr int GN_MAX_CHUNKS = 64;  // pixel chunks per sample (partials reduced by the apply kernel)

// One thread owns 8 consecutive channels (vector column vx) and walks rows ry, ry+RY, ...
// partial[n][chunk][g] = (sum, sumsq) over the chunk's rows and the group's channels.
template <typename T>
__global__ void gn_stats_kernel(const T* __restrict__ x, float2* __restrict__ partial, int HW, int C, int groups,
                                int ldx, int rows_per_chunk, int nchunks) {
  pdl_wait();
  extern __shared__ float sm[];  // [2][RY][C]
  const int VX = C / 8, RY = blockDim.x / VX;
  const int vx = threadIdx.x % VX, ry = threadIdx.x / VX;
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int r0 = chunk * rows_per_chunk, r1 = min(HW, r0 + rows_per_chunk);
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  if (ry < RY) {
    const T* base = x + (int64_t)n * HW * ldx + vx * 8;
    for (int r = r0 + ry; r < r1; r += 4 * RY) {
      float v[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (r + u * RY < r1) load8(base + (int64_t)(r + u * RY) * ldx, v[u]);
        else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[u][j] = 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s[j] += v[u][j];
          q[j] += v[u][j] * v[u][j];
        }
    }
    float* ss = sm + (int64_t)ry * C + vx * 8;
    float* qq = sm + (int64_t)(RY + ry) * C + vx * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      ss[j] = s[j];
      qq[j] = q[j];
    }
  }
  __syncthreads();
  // warp w reduces groups w, w+nwarps, ...: lanes stride over the RY*cg values of the group
  const int cg = C / groups, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int g = warp; g < groups; g += nwarps) {
    float a = 0.f, b = 0.f;
    for (int i = lane; i < RY * cg; i += 32) {
      int r = i / cg, c = g * cg + i % cg;
      a += sm[(int64_t)r * C + c];
      b += sm[(int64_t)(RY + r) * C + c];
    }
    a = warp_sum(a);
    b = warp_sum(b);
    if (lane == 0) partial[((int64_t)n * nchunks + chunk) * groups + g] = make_float2(a, b);
  }
}

template <typename T, typename TO, bool SILU>
__global__ void gn_apply_kernel(const T* __restrict__ x, TO* __restrict__ y, const float2* __restrict__ partial,
                                const float* __restrict__ gamma, const float* __restrict__ beta, int HW, int C,
                                int groups, int ldx, int ldy, int rows_per_chunk, int nchunks, float eps, long long wsplit) {
  pdl_wait();
  extern __shared__ float sm[];  // scale[C], shift[C]
  float* scale = sm;
  float* shift = sm + C;
  const int n = blockIdx.y, chunk = blockIdx.x;
  if (n >= wsplit) { gamma += C; beta += C; }  // second weight group (mkd_groupnorm wgroups)
  const int cg = C / groups;
  const float inv_cnt = 1.0f / ((float)cg * (float)HW);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    int g = c / cg;
    float a = 0.f, b = 0.f;
    for (int k = 0; k < nchunks; ++k) {
      float2 p = partial[((int64_t)n * nchunks + k) * groups + g];
      a += p.x;
      b += p.y;
    }
    float mean = a * inv_cnt;
    float var = fmaxf(b * inv_cnt - mean * mean, 0.f);
    float rstd = rsqrtf(var + eps);
    float sc = gamma[c] * rstd;
    scale[c] = sc;
    shift[c] = beta[c] - mean * sc;
  }
  __syncthreads();
  const int VX = C / 8, RY = blockDim.x / VX;
  const int vx = threadIdx.x % VX, ry = threadIdx.x / VX;
  if (ry >= RY) return;
  const int r0 = chunk * rows_per_chunk, r1 = min(HW, r0 + rows_per_chunk);
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = scale[vx * 8 + j];
    sh[j] = shift[vx * 8 + j];
  }
  const T* xb = x + (int64_t)n * HW * ldx + vx * 8;
  TO* yb = y + (int64_t)n * HW * ldy + vx * 8;
  for (int r = r0 + ry; r < r1; r += 4 * RY) {
    float v[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (r + u * RY < r1) load8(xb + (int64_t)(r + u * RY) * ldx, v[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (r + u * RY < r1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float t = v[u][j] * sc[j] + sh[j];
          v[u][j] = SILU ? silu_f(t) : t;
        }
        store8(yb + (int64_t)(r + u * RY) * ldy, v[u]);
      }
    }
  }
}

// Single-launch GroupNorm: gridDim.x = cluster size = chunks per sample, blockIdx.y = sample.
template <typename T, typename TO, bool SILU>
__global__ void gn_cluster_kernel(const T* __restrict__ x, TO* __restrict__ y, const float* __restrict__ gamma,
                                  const float* __restrict__ beta, int HW, int C, int groups, int ldx, int ldy,
                                  int rows_per_chunk, float eps, long long wsplit) {
  pdl_wait();
  extern __shared__ float sm[];      // phase 1: [2][RY][C] ; phase 2: scale[C], shift[C]
  __shared__ float2 part[64];        // this CTA's (sum, sumsq) per group — read by the whole cluster via DSMEM
  cg::cluster_group cluster = cg::this_cluster();
  const int VX = C / 8, RY = blockDim.x / VX;
  const int vx = threadIdx.x % VX, ry = threadIdx.x / VX;
  const int n = blockIdx.y, chunk = blockIdx.x, nchunks = gridDim.x;
  if (n >= wsplit) { gamma += C; beta += C; }  // second weight group (mkd_groupnorm wgroups)
  const int r0 = chunk * rows_per_chunk, r1 = min(HW, r0 + rows_per_chunk);
  const T* xb = x + (int64_t)n * HW * ldx + vx * 8;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  if (ry < RY) {
    for (int r = r0 + ry; r < r1; r += 4 * RY) {  // 4 independent 32-byte loads in flight per thread
      float v[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (r + u * RY < r1) load8(xb + (int64_t)(r + u * RY) * ldx, v[u]);
        else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[u][j] = 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s[j] += v[u][j];
          q[j] += v[u][j] * v[u][j];
        }
    }
    float* ss = sm + (int64_t)ry * C + vx * 8;
    float* qq = sm + (int64_t)(RY + ry) * C + vx * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      ss[j] = s[j];
      qq[j] = q[j];
    }
  }
  __syncthreads();
  const int cgs = C / groups, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int g = warp; g < groups; g += nwarps) {
    float a = 0.f, b = 0.f;
    for (int i = lane; i < RY * cgs; i += 32) {
      int r = i / cgs, c = g * cgs + i % cgs;
      a += sm[(int64_t)r * C + c];
      b += sm[(int64_t)(RY + r) * C + c];
    }
    a = warp_sum(a);
    b = warp_sum(b);
    if (lane == 0) part[g] = make_float2(a, b);
  }
  cluster.sync();  // every CTA's partials are published (also a CTA-wide barrier: phase-1 smem is free again)
  float* scale = sm;
  float* shift = sm + C;
  const float inv_cnt = 1.0f / ((float)cgs * (float)HW);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cgs;
    float a = 0.f, b = 0.f;
    for (int k = 0; k < nchunks; ++k) {
      const float2 p = *cluster.map_shared_rank(&part[g], k);
      a += p.x;
      b += p.y;
    }
    const float mean = a * inv_cnt;
    const float var = fmaxf(b * inv_cnt - mean * mean, 0.f);
    const float sc = gamma[c] * rsqrtf(var + eps);
    scale[c] = sc;
    shift[c] = beta[c] - mean * sc;
  }
  cluster.sync();  // nobody may exit (or overwrite `part`) while a peer still reads it; also publishes scale/shift
  if (ry >= RY) return;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = scale[vx * 8 + j];
    sh[j] = shift[vx * 8 + j];
  }
  TO* yb = y + (int64_t)n * HW * ldy + vx * 8;
  for (int r = r0 + ry; r < r1; r += 4 * RY) {
    float v[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (r + u * RY < r1) load8(xb + (int64_t)(r + u * RY) * ldx, v[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (r + u * RY < r1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float t = v[u][j] * sc[j] + sh[j];
          v[u][j] = SILU ? silu_f(t) : t;
        }
        store8(yb + (int64_t)(r + u * RY) * ldy, v[u]);
      }
    }
  }
}

// Small feature maps (8x8, 4x4 levels: C = 1280 / 2560): one CTA per (sample, group).  The group's HW x (C/groups)
// values (at most GG_VPT vectors of 8 per thread) are read ONCE into registers, block-reduced with an exact two-pass
// variance, and normalised from registers: one read, one write, no cluster, no workspace.
constexpr int GG_THREADS = 128, GG_VPT = 5;
template <typename T, typename TO, bool SILU>
__global__ void __launch_bounds__(GG_THREADS) gn_group_kernel(const T* __restrict__ x, TO* __restrict__ y,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              int HW, int C, int groups, int ldx, int ldy, float eps, long long wsplit) {
  pdl_wait();
  __shared__ float red[2][GG_THREADS / 32];
  const int g = blockIdx.x, n = blockIdx.y;
  if (n >= wsplit) { gamma += C; beta += C; }  // second weight group (mkd_groupnorm wgroups)
  const int cgs = C / groups, vpr = cgs / 8, nv = HW * vpr;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const T* xb = x + (int64_t)n * HW * ldx + g * cgs;
  float v[GG_VPT][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < GG_VPT; ++i) {
    const int idx = threadIdx.x + i * GG_THREADS;
    if (idx < nv) {
      load8(xb + (int64_t)(idx / vpr) * ldx + (idx % vpr) * 8, v[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[i][j];
    }
  }
  s = warp_sum(s);
  if (lane == 0) red[0][warp] = s;
  __syncthreads();
  const float inv_cnt = 1.0f / (float)(nv * 8);
  const float mean = (red[0][0] + red[0][1] + red[0][2] + red[0][3]) * inv_cnt;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < GG_VPT; ++i) {
    if (threadIdx.x + i * GG_THREADS < nv) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[i][j] - mean;
        q = fmaf(d, d, q);
      }
    }
  }
  q = warp_sum(q);
  if (lane == 0) red[1][warp] = q;
  __syncthreads();
  const float rstd = rsqrtf((red[1][0] + red[1][1] + red[1][2] + red[1][3]) * inv_cnt + eps);
  TO* yb = y + (int64_t)n * HW * ldy + g * cgs;
#pragma unroll
  for (int i = 0; i < GG_VPT; ++i) {
    const int idx = threadIdx.x + i * GG_THREADS;
    if (idx < nv) {
      const int c = (idx % vpr) * 8;
      float ga[8], be[8], o[8];
      load8(gamma + g * cgs + c, ga);
      load8(beta + g * cgs + c, be);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float t = (v[i][j] - mean) * rstd * ga[j] + be[j];
        o[j] = SILU ? silu_f(t) : t;
      }
      store8(yb + (int64_t)(idx / vpr) * ldy + c, o);
    }
  }
}

// GroupNorm as one streaming pass: the producer's GEMM epilogue already emitted, per 128-row tile and channel, the
// (sum, sum of squares) of the values it stored (mkd_conv_desc.stats).  Prologue: per-channel totals of this sample
// (fixed order: deterministic) -> 32 group statistics (one warp per group) -> per-channel scale / shift in smem; then
// y = act(x * scale + shift) over this CTA's row chunk, 8 channels per thread, 4 rows per batch.
// The kernel is latency-bound (a few hundred CTAs, each with a dependent prologue and 2-3 batches of rows), so what
// matters is how many memory round trips sit one behind the other: the statistics of up to MAXR channels x 4 tiles per
// thread are requested together and ahead of everything else, gamma / beta ride along, the first batch of x follows,
// and inside the row loop batch i + 1 is requested before batch i is normalised and stored.
template <typename T, typename TO, bool SILU>
__global__ void __launch_bounds__(512) gn_apply_stats_kernel(const T* __restrict__ x, TO* __restrict__ y, const float2* __restrict__ stats,
                                      int stats_ld, int tiles, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, int HW, int C, int groups, int ldx, int ldy,
                                      int rows_per_chunk, float eps, long long wsplit) {
  pdl_wait();
  extern __shared__ float sm[];  // scale[C], shift[C], then float2 csum[C] (dead after the prologue)
  float* scale = sm;
  float* shift = sm + C;
  float2* csum = reinterpret_cast<float2*>(sm + 2 * C);
  __shared__ float2 gstat[64];   // (mean, rstd) per group
  const int n = blockIdx.y, chunk = blockIdx.x;
  if (n >= wsplit) { gamma += C; beta += C; }  // second weight group (mkd_groupnorm wgroups)
  const int cgs = C / groups, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int VX = C / 8, RY = blockDim.x / VX;
  const int vx = threadIdx.x % VX, ry = threadIdx.x / VX;
  const int r0 = chunk * rows_per_chunk, r1 = min(HW, r0 + rows_per_chunk);
  const T* xb = x + (int64_t)n * HW * ldx + vx * 8;
  constexpr int MAXR = 4;  // channel rounds of the fast prologue: C <= MAXR * blockDim
  const bool fast = C <= MAXR * (int)blockDim.x && blockDim.x < 2 * C;
  float ga[MAXR], be[MAXR];
  float2 tot[MAXR];
  float v[4][8];
  if (fast) {
    // statistics first (they head the dependent chain): rounds x 4 tiles in flight, summed in tile order per channel
    const float2* sp = stats + (int64_t)n * tiles * stats_ld;
#pragma unroll
    for (int rr = 0; rr < MAXR; ++rr) tot[rr] = make_float2(0.f, 0.f);
    for (int k0 = 0; k0 < tiles; k0 += 4) {
      float2 t[MAXR][4];
#pragma unroll
      for (int rr = 0; rr < MAXR; ++rr) {
        const int c = threadIdx.x + rr * blockDim.x;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          t[rr][k] = (c < C && k0 + k < tiles) ? sp[(int64_t)(k0 + k) * stats_ld + c] : make_float2(0.f, 0.f);
      }
      if (k0 == 0) {
#pragma unroll
        for (int rr = 0; rr < MAXR; ++rr) {
          const int c = threadIdx.x + rr * blockDim.x;
          ga[rr] = c < C ? gamma[c] : 0.f;
          be[rr] = c < C ? beta[c] : 0.f;
        }
        if (ry < RY) {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (r0 + ry + u * RY < r1) load8(xb + (int64_t)(r0 + ry + u * RY) * ldx, v[u]);
        }
      }
#pragma unroll
      for (int rr = 0; rr < MAXR; ++rr) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          tot[rr].x += t[rr][k].x;
          tot[rr].y += t[rr][k].y;
        }
      }
    }
#pragma unroll
    for (int rr = 0; rr < MAXR; ++rr) {
      const int c = threadIdx.x + rr * blockDim.x;
      if (c < C) csum[c] = tot[rr];
    }
  } else {
    // the first batch of x loads does not depend on the statistics: issue it before the prologue so that its DRAM / L2
    // latency overlaps the (dependent, three-barrier) statistics chain
    if (ry < RY) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (r0 + ry + u * RY < r1) load8(xb + (int64_t)(r0 + ry + u * RY) * ldx, v[u]);
    }
    // per-channel totals of this sample's tile partials.  When the CTA has more threads than channels, G thread groups
    // split the tiles (a 256^2 map has 512 of them); fixed partition and fixed summation order: deterministic.
    const int G = blockDim.x >= 2 * C ? blockDim.x / C : 1;
    for (int idx = threadIdx.x; idx < G * C; idx += blockDim.x) {
      const int gq = idx / C, c = idx - gq * C;
      const float2* p = stats + (int64_t)n * tiles * stats_ld + c;
      float a = 0.f, b = 0.f;
#pragma unroll 4
      for (int k = gq; k < tiles; k += G) {
        const float2 t = p[(int64_t)k * stats_ld];
        a += t.x;
        b += t.y;
      }
      csum[idx] = make_float2(a, b);
    }
    if (G > 1) {
      __syncthreads();
      float a = 0.f, b = 0.f;
      if (threadIdx.x < C) {
        for (int gq = 0; gq < G; ++gq) {
          a += csum[gq * C + threadIdx.x].x;
          b += csum[gq * C + threadIdx.x].y;
        }
      }
      __syncthreads();
      if (threadIdx.x < C) csum[threadIdx.x] = make_float2(a, b);
    }
  }
  __syncthreads();
  const float inv_cnt = 1.0f / ((float)cgs * (float)HW);
  for (int g = warp; g < groups; g += nwarps) {
    float a = 0.f, b = 0.f;
    for (int i = lane; i < cgs; i += 32) {
      const float2 t = csum[g * cgs + i];
      a += t.x;
      b += t.y;
    }
    a = warp_sum(a);
    b = warp_sum(b);
    if (lane == 0) {
      const float mean = a * inv_cnt;
      const float var = fmaxf(b * inv_cnt - mean * mean, 0.f);
      gstat[g] = make_float2(mean, rsqrtf(var + eps));
    }
  }
  __syncthreads();
  if (fast) {
#pragma unroll
    for (int rr = 0; rr < MAXR; ++rr) {
      const int c = threadIdx.x + rr * blockDim.x;
      if (c < C) {
        const float2 ms = gstat[c / cgs];
        const float sc = ga[rr] * ms.y;
        scale[c] = sc;
        shift[c] = be[rr] - ms.x * sc;
      }
    }
  } else {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const float2 ms = gstat[c / cgs];
      const float sc = gamma[c] * ms.y;
      scale[c] = sc;
      shift[c] = beta[c] - ms.x * sc;
    }
  }
  __syncthreads();
  if (ry >= RY) return;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = scale[vx * 8 + j];
    sh[j] = shift[vx * 8 + j];
  }
  TO* yb = y + (int64_t)n * HW * ldy + vx * 8;
  float w[4][8];  // the batch after the one being normalised
  for (int r = r0 + ry; r < r1; r += 4 * RY) {
    const int rn = r + 4 * RY;
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (rn + u * RY < r1) load8(xb + (int64_t)(rn + u * RY) * ldx, w[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (r + u * RY < r1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float t = v[u][j] * sc[j] + sh[j];
          v[u][j] = SILU ? silu_f(t) : t;
        }
        store8(yb + (int64_t)(r + u * RY) * ldy, v[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int j = 0; j < 8; ++j) v[u][j] = w[u][j];
  }
}

// LayerNorm: one warp per row, the row lives in registers between the mean and the variance pass
// (exact two-pass variance, like the reference).  C <= 8 * 32 * LN_VPL.
constexpr int LN_VPL = 8;
template <typename T, typename TO>
__global__ void layernorm_kernel(const T* __restrict__ x, TO* __restrict__ y, int64_t M, int C, int ldx, int ldy,
                                 const float* __restrict__ gamma, const float* __restrict__ beta, float eps, long long wsplit) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= wsplit) { gamma += C; beta += C; }  // second weight group (mkd_layernorm wgroups)
  if (row >= M) return;
  const int nv = C / 8;
  float v[LN_VPL][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LN_VPL; ++i) {
    int vi = lane + i * 32;
    if (vi < nv) {
      load8(x + row * ldx + vi * 8, v[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[i][j];
    }
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < LN_VPL; ++i) {
    int vi = lane + i * 32;
    if (vi < nv) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float d = v[i][j] - mean;
        q += d * d;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
#pragma unroll
  for (int i = 0; i < LN_VPL; ++i) {
    int vi = lane + i * 32;
    if (vi < nv) {
      float g[8], b[8], o[8];
      load8(gamma + vi * 8, g);
      load8(beta + vi * 8, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd * g[j] + b[j];
      store8(y + row * ldy + vi * 8, o);
    }
  }
}
// LayerNorm for the transformer widths of this model, C = 40 * LPR with LPR = 8 / 16 / 32 lanes per row (C = 320 / 640 /
// 1280): every lane owns exactly 5 vectors of 8 channels, so all 32 lanes are busy (the generic kernel leaves 24 of 32
// lanes idle on the second vector of a 320-wide row) and 32 / LPR rows share a warp; 5 independent 32-byte loads in
// flight per lane; sub-warp shuffles for the statistics (exact two-pass variance).
template <typename T, typename TO, int LPR>
__global__ void __launch_bounds__(256) layernorm5_kernel(const T* __restrict__ x, TO* __restrict__ y, int64_t M, int ldx,
                                                         int ldy, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, float eps, long long wsplit) {
  pdl_wait();
  constexpr int RPW = 32 / LPR, C = 40 * LPR;
  const int lane = threadIdx.x & 31, sub = lane % LPR;
  const int64_t row = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + lane / LPR;
  if (row >= wsplit) { gamma += C; beta += C; }  // second weight group (mkd_layernorm wgroups)
  const bool ok = row < M;  // inactive lanes still take part in the shuffles
  float v[5][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    if (ok) load8(x + row * ldx + (sub + i * LPR) * 8, v[i]);
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
    }
  }
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[i][j];
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.0f / (float)C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = v[i][j] - mean;
      q = fmaf(d, d, q);
    }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q * (1.0f / (float)C) + eps);
  if (!ok) return;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const int c = (sub + i * LPR) * 8;
    float g[8], b[8], o[8];
    load8(gamma + c, g);
    load8(beta + c, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd * g[j] + b[j];
    store8(y + row * ldy + c, o);
  }
}

// Row softmax y = softmax(scale * x) over the last dim, one warp per row (the VAE decoder's single-head attention runs
// as S = Q K^T (GEMM) -> this -> P V (GEMM); statistics in fp32, exp2 with the scale folded in).
template <typename T, typename TO>
__global__ void softmax_rows_kernel(const T* __restrict__ x, TO* __restrict__ y, int64_t M, int C, int ldx, int ldy,
                                    float scale_log2e) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int nv = C / 8;
  float v[LN_VPL][8];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < LN_VPL; ++i) {
    const int vi = lane + i * 32;
    if (vi < nv) {
      load8(x + row * ldx + vi * 8, v[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) mx = fmaxf(mx, v[i][j]);
    }
  }
  mx = warp_max(mx) * scale_log2e;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LN_VPL; ++i) {
    if (lane + i * 32 < nv) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[i][j] = exp2f(fmaf(v[i][j], scale_log2e, -mx));
        s += v[i][j];
      }
    }
  }
  const float inv = 1.0f / warp_sum(s);
#pragma unroll
  for (int i = 0; i < LN_VPL; ++i) {
    const int vi = lane + i * 32;
    if (vi < nv) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i][j] *= inv;
      store8(y + row * ldy + vi * 8, v[i]);
    }
  }
}

template <typename T, typename TO>
static bool layernorm5_launch(const T* x, TO* y, int64_t M, int C, int ldx, int ldy, const float* gamma, const float* beta,
                              float eps, long long wsplit, cudaStream_t st, cudaError_t* err) {
  const int warps = 8;
  auto go = [&](auto kernel, int rpw) {
    const int64_t rows_per_cta = (int64_t)warps * rpw;
    *err = launch_pdl(kernel, dim3((unsigned)((M + rows_per_cta - 1) / rows_per_cta)), dim3(warps * 32), 0, st, x, y, M, ldx, ldy, gamma, beta, eps, wsplit);
    return true;
  };
  if (C == 320) return go(layernorm5_kernel<T, TO, 8>, 4);
  if (C == 640) return go(layernorm5_kernel<T, TO, 16>, 2);
  if (C == 1280) return go(layernorm5_kernel<T, TO, 32>, 1);
  return false;
}
}  // namespace

extern "C" size_t mkd_groupnorm_workspace_bytes(int N, int groups) {
  return (size_t)N * GN_MAX_CHUNKS * groups * sizeof(float2);
}

template <typename T, typename TO>
static int groupnorm_launch(const T* x, TO* y, int N, int HW, int C, int groups, int ldx, int ldy, const float* gamma,
                            const float* beta, float eps, int silu, float2* partial, long long wsplit, cudaStream_t st) {
  const int VX = C / 8;
  int threads = VX >= 256 ? VX : (256 / VX) * VX;  // whole number of row lanes
  threads = ((threads + 31) / 32) * 32;
  const int RY = threads / VX;
  if ((C / groups) % 8 == 0 && HW * (C / groups / 8) <= GG_THREADS * GG_VPT && aligned16(gamma) && aligned16(beta)) {
    // small maps: one CTA per (sample, group), the group lives in registers
    dim3 grid(groups, N);
    if (silu)
      MKD_LAUNCH_OK(launch_pdl(gn_group_kernel<T, TO, true>, grid, dim3(GG_THREADS), 0, st, x, y, gamma, beta, HW, C, groups, ldx, ldy, eps, wsplit));
    else
      MKD_LAUNCH_OK(launch_pdl(gn_group_kernel<T, TO, false>, grid, dim3(GG_THREADS), 0, st, x, y, gamma, beta, HW, C, groups, ldx, ldy, eps, wsplit));
    MKD_CHECK_LAUNCH();
    return MKD_OK;
  }
  {
    // single-launch cluster path: P = 1, 2, 4 or 8 chunks per sample (one cluster), when that fills enough SMs
    int P = 8;
    while (P > 1 && HW / P < 2 * RY) P >>= 1;
    if ((N * P >= 64 || (int64_t)HW * C <= 64 * 1024) && groups <= 64) {
      const int rows = (HW + P - 1) / P;
      const size_t smem = (size_t)2 * RY * C * sizeof(float) > (size_t)2 * C * sizeof(float) ? (size_t)2 * RY * C * sizeof(float)
                                                                                                 : (size_t)2 * C * sizeof(float);
      MKD_REQUIRE(smem <= 48 * 1024, MKD_E_INVALID, "groupnorm: C=%d too large", C);
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(P, N);
      cfg.blockDim = dim3(threads);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = st;
      cudaLaunchAttribute attr[2];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = P;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
      cfg.attrs = attr;
      cfg.numAttrs = 2;
      cudaError_t e = silu ? cudaLaunchKernelEx(&cfg, gn_cluster_kernel<T, TO, true>, x, y, gamma, beta, HW, C, groups, ldx, ldy, rows, eps, wsplit)
                           : cudaLaunchKernelEx(&cfg, gn_cluster_kernel<T, TO, false>, x, y, gamma, beta, HW, C, groups, ldx, ldy, rows, eps, wsplit);
      MKD_REQUIRE(e == cudaSuccess, MKD_E_CUDA, "groupnorm: cluster launch failed: %s", cudaGetErrorString(e));
      MKD_CHECK_LAUNCH();
      return MKD_OK;
    }
  }
  // enough CTAs to fill the machine (~4 per SM), at least 2*RY rows per chunk
  int want = (148 * 4 + N - 1) / N;
  int nchunks = HW / (2 * RY);
  if (nchunks > want) nchunks = want;
  if (nchunks > GN_MAX_CHUNKS) nchunks = GN_MAX_CHUNKS;
  if (nchunks < 1) nchunks = 1;
  int rows_per_chunk = (HW + nchunks - 1) / nchunks;
  nchunks = (HW + rows_per_chunk - 1) / rows_per_chunk;
  dim3 grid(nchunks, N);
  size_t sm1 = (size_t)2 * RY * C * sizeof(float), sm2 = (size_t)2 * C * sizeof(float);
  MKD_REQUIRE(sm1 <= 48 * 1024 && sm2 <= 48 * 1024, MKD_E_INVALID, "groupnorm: C=%d too large", C);
  MKD_LAUNCH_OK(launch_pdl(gn_stats_kernel<T>, dim3(grid), dim3(threads), sm1, st, x, partial, HW, C, groups, ldx, rows_per_chunk, nchunks));
  MKD_CHECK_LAUNCH();
  if (silu)
    MKD_LAUNCH_OK(launch_pdl(gn_apply_kernel<T, TO, true>, dim3(grid), dim3(threads), sm2, st, x, y, partial, gamma, beta, HW, C, groups, ldx, ldy,
                                                         rows_per_chunk, nchunks, eps, wsplit));
  else
    MKD_LAUNCH_OK(launch_pdl(gn_apply_kernel<T, TO, false>, dim3(grid), dim3(threads), sm2, st, x, y, partial, gamma, beta, HW, C, groups, ldx, ldy,
                                                          rows_per_chunk, nchunks, eps, wsplit));
  MKD_CHECK_LAUNCH();
  return MKD_OK;
}

extern "C" int mkd_groupnorm(const void* x, void* y, int x_dtype, int y_dtype, int N, int HW, int C, int groups, int ldx,
                             int ldy, const float* gamma, const float* beta, float eps, int silu, void* workspace,
                             size_t workspace_bytes, int wgroups, mkd_stream_t stream) {
  MKD_REQUIRE(x && y && gamma && beta && workspace && N > 0 && HW > 0 && C > 0 && groups > 0, MKD_E_INVALID,
              "groupnorm: bad args");
  MKD_REQUIRE(wgroups == 1 || (wgroups == 2 && N % 2 == 0), MKD_E_INVALID, "groupnorm: wgroups must be 1, or 2 with an even batch");
  const long long wsplit = wgroups == 2 ? N / 2 : (1ll << 62);
  MKD_REQUIRE(C % groups == 0 && C % 8 == 0 && C <= 8 * 1024, MKD_E_INVALID,
              "groupnorm: C=%d must be a multiple of groups=%d and of 8", C, groups);
  MKD_REQUIRE(N <= 65535, MKD_E_INVALID, "groupnorm: N too large");
  MKD_REQUIRE(ldx % 8 == 0 && ldy % 8 == 0 && aligned16(x) && aligned16(y) && ldx >= C && ldy >= C, MKD_E_ALIGN,
              "groupnorm: ld must be a multiple of 8 (>= C) and pointers 16B aligned");
  MKD_REQUIRE(workspace_bytes >= mkd_groupnorm_workspace_bytes(N, groups), MKD_E_WORKSPACE,
              "groupnorm: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  float2* ws = (float2*)workspace;
  if (x_dtype == MKD_BF16 && y_dtype == MKD_BF16)
    return groupnorm_launch<bf16, bf16>((const bf16*)x, (bf16*)y, N, HW, C, groups, ldx, ldy, gamma, beta, eps, silu, ws, wsplit, st);
  if (x_dtype == MKD_F32 && y_dtype == MKD_BF16)
    return groupnorm_launch<float, bf16>((const float*)x, (bf16*)y, N, HW, C, groups, ldx, ldy, gamma, beta, eps, silu, ws, wsplit, st);
  if (x_dtype == MKD_F32 && y_dtype == MKD_F32)
    return groupnorm_launch<float, float>((const float*)x, (float*)y, N, HW, C, groups, ldx, ldy, gamma, beta, eps, silu, ws, wsplit, st);
  MKD_REQUIRE(false, MKD_E_INVALID, "groupnorm: unsupported dtype pair %d -> %d", x_dtype, y_dtype);
}

template <typename T, typename TO>
static int gn_apply_launch(const T* x, TO* y, int N, int HW, int C, int groups, int ldx, int ldy, const float* gamma,
                           const float* beta, float eps, int silu, const float2* stats, int stats_ld, int tiles,
                           long long wsplit, cudaStream_t st) {
  const int VX = C / 8;
  int threads = VX >= 256 ? VX : (256 / VX) * VX;
  threads = ((threads + 31) / 32) * 32;
  const int RY = threads / VX;
  const size_t smem = (size_t)(2 * C + 2 * (C > threads ? C : threads)) * sizeof(float);  // scale, shift, csum[G][C]
  MKD_REQUIRE(smem <= 48 * 1024, MKD_E_INVALID, "groupnorm_apply: C=%d too large", C);
  // ONE wave: as many chunks per sample as the CTAs that can be resident at once allow (a 19th chunk per sample at batch 16
  // was a second wave of 8 CTAs that doubled the kernel's time), each streaming at least 4 * RY rows
  int occ = 0;
  {
    cudaError_t e = silu ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gn_apply_stats_kernel<T, TO, true>, threads, smem)
                         : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gn_apply_stats_kernel<T, TO, false>, threads, smem);
    MKD_REQUIRE(e == cudaSuccess && occ > 0, MKD_E_CUDA, "groupnorm_apply: occupancy query failed");
  }
  int nchunks = occ * num_sms() / N;
  if (nchunks > HW / (4 * RY)) nchunks = HW / (4 * RY);
  if (nchunks < 1) nchunks = 1;
  const int rows = (HW + nchunks - 1) / nchunks;
  nchunks = (HW + rows - 1) / rows;
  dim3 grid(nchunks, N);
  if (silu)
    MKD_LAUNCH_OK(launch_pdl(gn_apply_stats_kernel<T, TO, true>, grid, dim3(threads), smem, st, x, y, stats, stats_ld, tiles, gamma, beta, HW, C,
                             groups, ldx, ldy, rows, eps, wsplit));
  else
    MKD_LAUNCH_OK(launch_pdl(gn_apply_stats_kernel<T, TO, false>, grid, dim3(threads), smem, st, x, y, stats, stats_ld, tiles, gamma, beta, HW, C,
                             groups, ldx, ldy, rows, eps, wsplit));
  MKD_CHECK_LAUNCH();
  return MKD_OK;
}

extern "C" int mkd_groupnorm_apply(const void* x, void* y, int x_dtype, int y_dtype, int N, int HW, int C, int groups,
                                   int ldx, int ldy, const float* gamma, const float* beta, float eps, int silu,
                                   const float* stats, int stats_ld, int tiles_per_sample, int wgroups, mkd_stream_t stream) {
  MKD_REQUIRE(x && y && gamma && beta && stats && N > 0 && HW > 0 && C > 0 && groups > 0 && groups <= 64, MKD_E_INVALID,
              "groupnorm_apply: bad args");
  MKD_REQUIRE(wgroups == 1 || (wgroups == 2 && N % 2 == 0), MKD_E_INVALID, "groupnorm_apply: wgroups must be 1, or 2 with an even batch");
  const long long wsplit = wgroups == 2 ? N / 2 : (1ll << 62);
  MKD_REQUIRE(C % groups == 0 && C % 8 == 0 && C <= 8 * 512, MKD_E_INVALID,
              "groupnorm_apply: C=%d must be a multiple of groups=%d and of 8", C, groups);
  MKD_REQUIRE(N <= 65535 && HW == 128 * tiles_per_sample && stats_ld >= C, MKD_E_INVALID,
              "groupnorm_apply: HW=%d must be 128 * tiles_per_sample=%d, stats_ld=%d >= C", HW, tiles_per_sample, stats_ld);
  MKD_REQUIRE(ldx % 8 == 0 && ldy % 8 == 0 && aligned16(x) && aligned16(y) && ldx >= C && ldy >= C &&
                  ((uintptr_t)stats & 7) == 0,
              MKD_E_ALIGN, "groupnorm_apply: ld must be a multiple of 8 (>= C), pointers 16B aligned, stats 8B aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const float2* sp = reinterpret_cast<const float2*>(stats);
  if (x_dtype == MKD_BF16 && y_dtype == MKD_BF16)
    return gn_apply_launch<bf16, bf16>((const bf16*)x, (bf16*)y, N, HW, C, groups, ldx, ldy, gamma, beta, eps, silu, sp, stats_ld, tiles_per_sample, wsplit, st);
  if (x_dtype == MKD_F32 && y_dtype == MKD_BF16)
    return gn_apply_launch<float, bf16>((const float*)x, (bf16*)y, N, HW, C, groups, ldx, ldy, gamma, beta, eps, silu, sp, stats_ld, tiles_per_sample, wsplit, st);
  if (x_dtype == MKD_F32 && y_dtype == MKD_F32)
    return gn_apply_launch<float, float>((const float*)x, (float*)y, N, HW, C, groups, ldx, ldy, gamma, beta, eps, silu, sp, stats_ld, tiles_per_sample, wsplit, st);
  MKD_REQUIRE(false, MKD_E_INVALID, "groupnorm_apply: unsupported dtype pair %d -> %d", x_dtype, y_dtype);
}

extern "C" int mkd_softmax_rows(const void* x, void* y, int x_dtype, int y_dtype, int64_t M, int C, int ldx, int ldy,
                                float scale, mkd_stream_t stream) {
  MKD_REQUIRE(x && y && M > 0 && C > 0, MKD_E_INVALID, "softmax_rows: bad args");
  MKD_REQUIRE(C % 8 == 0 && C <= 8 * 32 * LN_VPL && scale > 0.f, MKD_E_INVALID,
              "softmax_rows: C=%d must be a multiple of 8, <= %d; scale must be positive", C, 8 * 32 * LN_VPL);
  MKD_REQUIRE(ldx % 8 == 0 && ldy % 8 == 0 && aligned16(x) && aligned16(y), MKD_E_ALIGN,
              "softmax_rows: ld must be a multiple of 8 and pointers 16B aligned");
  const int warps = 8;
  const dim3 grid((unsigned)((M + warps - 1) / warps)), block(warps * 32);
  const float sl2 = scale * 1.4426950408889634f;
  cudaStream_t st = (cudaStream_t)stream;
  if (x_dtype == MKD_F32 && y_dtype == MKD_BF16)
    MKD_LAUNCH_OK(launch_pdl(softmax_rows_kernel<float, bf16>, grid, block, 0, st, (const float*)x, (bf16*)y, M, C, ldx, ldy, sl2));
  else if (x_dtype == MKD_F32 && y_dtype == MKD_F32)
    MKD_LAUNCH_OK(launch_pdl(softmax_rows_kernel<float, float>, grid, block, 0, st, (const float*)x, (float*)y, M, C, ldx, ldy, sl2));
  else if (x_dtype == MKD_BF16 && y_dtype == MKD_BF16)
    MKD_LAUNCH_OK(launch_pdl(softmax_rows_kernel<bf16, bf16>, grid, block, 0, st, (const bf16*)x, (bf16*)y, M, C, ldx, ldy, sl2));
  else
    MKD_REQUIRE(false, MKD_E_INVALID, "softmax_rows: unsupported dtype pair %d -> %d", x_dtype, y_dtype);
  MKD_CHECK_LAUNCH();
  return MKD_OK;
}

extern "C" int mkd_layernorm(const void* x, void* y, int x_dtype, int y_dtype, int64_t M, int C, int ldx, int ldy,
                             const float* gamma, const float* beta, float eps, int wgroups, mkd_stream_t stream) {
  MKD_REQUIRE(x && y && gamma && beta && M > 0 && C > 0, MKD_E_INVALID, "layernorm: bad args");
  MKD_REQUIRE(wgroups == 1 || (wgroups == 2 && M % 2 == 0), MKD_E_INVALID, "layernorm: wgroups must be 1, or 2 with an even row count");
  const long long wsplit = wgroups == 2 ? M / 2 : (1ll << 62);
  MKD_REQUIRE(C % 8 == 0 && C <= 8 * 32 * LN_VPL, MKD_E_INVALID, "layernorm: C=%d must be a multiple of 8, <= %d", C,
              8 * 32 * LN_VPL);
  MKD_REQUIRE(ldx % 8 == 0 && ldy % 8 == 0 && aligned16(x) && aligned16(y) && aligned16(gamma) && aligned16(beta),
              MKD_E_ALIGN, "layernorm: ld must be a multiple of 8 and pointers 16B aligned");
  const int warps = 8;
  int64_t blocks = (M + warps - 1) / warps;
  cudaStream_t st = (cudaStream_t)stream;
  {
    cudaError_t le = cudaSuccess;
    bool took = false;
    if (x_dtype == MKD_BF16 && y_dtype == MKD_BF16)
      took = layernorm5_launch((const bf16*)x, (bf16*)y, M, C, ldx, ldy, gamma, beta, eps, wsplit, st, &le);
    else if (x_dtype == MKD_F32 && y_dtype == MKD_BF16)
      took = layernorm5_launch((const float*)x, (bf16*)y, M, C, ldx, ldy, gamma, beta, eps, wsplit, st, &le);
    else if (x_dtype == MKD_F32 && y_dtype == MKD_F32)
      took = layernorm5_launch((const float*)x, (float*)y, M, C, ldx, ldy, gamma, beta, eps, wsplit, st, &le);
    if (took) {
      MKD_LAUNCH_OK(le);
      MKD_CHECK_LAUNCH();
      return MKD_OK;
    }
  }
  if (x_dtype == MKD_BF16 && y_dtype == MKD_BF16)
    MKD_LAUNCH_OK(launch_pdl(layernorm_kernel<bf16, bf16>, dim3((unsigned)blocks), dim3(warps * 32), 0, st, (const bf16*)x, (bf16*)y, M, C, ldx, ldy, gamma, beta, eps, wsplit));
  else if (x_dtype == MKD_F32 && y_dtype == MKD_BF16)
    MKD_LAUNCH_OK(launch_pdl(layernorm_kernel<float, bf16>, dim3((unsigned)blocks), dim3(warps * 32), 0, st, (const float*)x, (bf16*)y, M, C, ldx, ldy, gamma, beta, eps, wsplit));
  else if (x_dtype == MKD_F32 && y_dtype == MKD_F32)
    MKD_LAUNCH_OK(launch_pdl(layernorm_kernel<float, float>, dim3((unsigned)blocks), dim3(warps * 32), 0, st, (const float*)x, (float*)y, M, C, ldx, ldy, gamma, beta, eps, wsplit));
  else
    MKD_REQUIRE(false, MKD_E_INVALID, "layernorm: unsupported dtype pair %d -> %d", x_dtype, y_dtype);
  MKD_CHECK_LAUNCH();
  return MKD_OK;
}
