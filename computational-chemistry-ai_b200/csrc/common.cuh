// Shared device helpers for the EI-MS GCN kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/eims_b200.h"

#ifndef __CUDA_ARCH_LIST__
#endif

namespace eims {

// dims[] slots written by K1 and read by every later kernel (sizes stay on the device so
// a captured step can be replayed for batches of any size up to the plan capacity).
enum { DIM_B = 0, DIM_N = 1, DIM_E = 2, DIM_ZERO_DEG = 3, DIM_OVERFLOW = 4, DIM_COUNT = 8 };

constexpr float kBnEps = 1e-5f;       // nn.BatchNorm1d default (GCN:317)
constexpr float kBnMomentum = 0.1f;   // nn.BatchNorm1d default
constexpr float kLnEps = 1e-5f;       // nn.LayerNorm default (GCN:343)
constexpr float kCosEps = 1e-8f;      // GCN:213-214

// Dropout sites: GCN layer l -> site l (GCN:362-363); head dropouts -> L, L+1 (GCN:345,349).
struct DropCfg {
  uint32_t threshold;  // keep iff philox word >= threshold ; threshold = p * 2^32
  float scale;         // 1/(1-p)
  uint32_t seed_lo, seed_hi, step, site;
  __host__ __device__ bool active() const { return threshold != 0; }
};

static inline DropCfg make_drop(float p, uint64_t seed, int step, int site) {
  DropCfg d;
  double t = (double)p * 4294967296.0;
  d.threshold = p <= 0.f ? 0u : (t >= 4294967295.0 ? 4294967295u : (uint32_t)t);
  d.scale = p <= 0.f ? 1.f : 1.f / (1.f - p);
  d.seed_lo = (uint32_t)seed;
  d.seed_hi = (uint32_t)(seed >> 32);
  d.step = (uint32_t)step;
  d.site = (uint32_t)site;
  return d;
}

// Philox4x32-10 (Salmon et al. 2011): counter = (element/4 lo, element/4 hi, site, step).
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}

// keep-mask (scaled) for the 4 consecutive elements starting at flat index `elem` (elem % 4 == 0)
__device__ __forceinline__ float4 drop_mask4(const DropCfg& d, uint64_t elem) {
  uint64_t q = elem >> 2;
  uint4 r = philox4x32_10(make_uint4((uint32_t)q, (uint32_t)(q >> 32), d.site, d.step),
                          make_uint2(d.seed_lo, d.seed_hi));
  float4 m;
  m.x = r.x >= d.threshold ? d.scale : 0.f;
  m.y = r.y >= d.threshold ? d.scale : 0.f;
  m.z = r.z >= d.threshold ? d.scale : 0.f;
  m.w = r.w >= d.threshold ? d.scale : 0.f;
  return m;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// Grid-wide "last block reduces" ticket.  Returns true in exactly one block (the last to
// arrive) after all earlier blocks' global writes are visible; resets the counter so the
// kernel can be relaunched / replayed from a CUDA graph.
__device__ __forceinline__ bool last_block_ticket(unsigned int* counter, unsigned int nblocks) {
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = atomicAdd(counter, 1u);
    is_last = (t == nblocks - 1);
    if (is_last) *counter = 0u;
  }
  __syncthreads();
  if (is_last) __threadfence();
  return is_last;
}

}  // namespace eims
