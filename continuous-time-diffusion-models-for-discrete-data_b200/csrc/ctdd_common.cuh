// Shared device helpers: Philox4x32-10, the uniform -> Poisson / categorical maps, error plumbing.
// The arithmetic here is restated op-for-op (fp32) in oracle/rng.py; keep the two in sync.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/ctdd.h"

namespace ctdd {

// ------------------------------------------------------------------------------------------------
// error / launch accounting (host side)
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
#define CTDD_CHECK_LAUNCH(name)                                                          \
  do {                                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                \
    if (e__ != cudaSuccess) {                                                            \
      ctdd::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));           \
      return 1;                                                                          \
    }                                                                                    \
    ctdd::count_launch();                                                                \
  } while (0)

// ------------------------------------------------------------------------------------------------
// RNG streams (Philox counter word c3)
enum : uint32_t {
  STREAM_JUMP_HI = 0,   // top 16 bits of the per-(row, s) jump uniform
  STREAM_JUMP_LO = 1,   // low 16 bits, only evaluated on the slow path
  STREAM_ROW = 2,       // one 32-bit uniform per row (Euler categorical draw)
  STREAM_INIT = 3,      // initial samples
  STREAM_NOISE_XT = 4,  // forward noising x_t
  STREAM_TILDE_DIM = 5, // x~: which dimension
  STREAM_TILDE_VAL = 6  // x~: new value
};

struct Philox4 { uint32_t w[4]; };

__host__ __device__ __forceinline__ void philox_mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#ifdef __CUDA_ARCH__
  lo = a * b;
  hi = __umulhi(a, b);
#else
  uint64_t p = (uint64_t)a * b;
  lo = (uint32_t)p;
  hi = (uint32_t)(p >> 32);
#endif
}

// Standard Philox4x32-10 (Salmon et al. 2011): same constants/round function as Random123 / cuRAND.
__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                         uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, lo0, hi1, lo1;
    philox_mulhilo(M0, c0, hi0, lo0);
    philox_mulhilo(M1, c2, hi1, lo1);
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  Philox4 o; o.w[0] = c0; o.w[1] = c1; o.w[2] = c2; o.w[3] = c3;
  return o;
}

// Jump-uniform layout: one Philox call serves 8 consecutive GLOBAL rows at one state s:
//   counter = (s, grow >> 3, offset_lo, stream | offset_hi << 8), halfword (grow & 7) of the 128-bit output.
__host__ __device__ __forceinline__ Philox4 philox_jump(uint32_t s, uint64_t grow_group, uint64_t offset,
                                                       uint32_t stream, uint64_t seed) {
  return philox4x32_10(s, (uint32_t)grow_group, (uint32_t)offset,
                       stream | ((uint32_t)(offset >> 32) << 8) | ((uint32_t)(grow_group >> 32) << 24),
                       (uint32_t)seed, (uint32_t)(seed >> 32));
}
__host__ __device__ __forceinline__ uint32_t philox_half(const Philox4& p, int i) {
  return (p.w[i >> 1] >> (16 * (i & 1))) & 0xFFFFu;
}

// Per-row 32-bit draw: one Philox call serves 4 consecutive global rows: counter = (sub, grow >> 2, ...).
__host__ __device__ __forceinline__ uint32_t philox_row_word(uint64_t grow, uint32_t sub, uint64_t offset,
                                                            uint32_t stream, uint64_t seed) {
  Philox4 p = philox4x32_10(sub, (uint32_t)(grow >> 2), (uint32_t)offset,
                            stream | ((uint32_t)(offset >> 32) << 8) | ((uint32_t)(grow >> 34) << 24),
                            (uint32_t)seed, (uint32_t)(seed >> 32));
  return p.w[grow & 3];
}

// 32 random bits -> v in (0, 1]:  v = (word + 0.5) * 2^-32 evaluated in fp32 (one rounding).
__host__ __device__ __forceinline__ float u32_to_unit(uint32_t w) {
#ifdef __CUDA_ARCH__
  return __fmaf_rn(__uint2float_rn(w), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
#else
  return fmaf((float)w, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
#endif
}

// P(K >= 1) for K ~ Poisson(lam): 1 - exp(-lam), via a 3-term series below 2^-6 (no cancellation).
__host__ __device__ __forceinline__ float poisson_sf0(float lam) {
  if (lam < 0.015625f) {
    float t = fmaf(lam, -0.16666667f, 0.5f);
    float u = fmaf(-lam, t, 1.0f);
    return lam * u;
  }
  return 1.0f - expf(-lam);
}

// Upper-tail inverse CDF: k = #{ j >= 0 : v < P(K > j) }, v in (0,1] small <=> many jumps.
// lam <= 64: exact pmf recurrence in fp32 (capped); lam > 64: Cornish-Fisher normal approximation.
// Callers may skip the call when v >= lam, because P(K>=1) <= lam.
__host__ __device__ __forceinline__ int poisson_from_unit(float lam, float v) {
  if (!(lam > 0.0f)) return 0;
  float sf = poisson_sf0(lam);
  if (v >= sf) return 0;
  if (lam <= 64.0f) {
    float p = expf(-lam);
    int kmax = (int)(lam + 10.0f * sqrtf(lam) + 12.0f);
    int k = 1;
    while (k < kmax) {
      p = p * lam / (float)k;   // pmf(k)
      sf -= p;                  // P(K > k)
      if (v >= sf) break;
      ++k;
    }
    return k;
  }
  if (lam > 1.0e9f) lam = 1.0e9f;
  // z = upper-tail normal quantile of v
#ifdef __CUDA_ARCH__
  float z = -normcdfinvf(v);
#else
  float z = 0.0f;  // host build never evaluates this branch (oracle has its own implementation)
#endif
  float k = lam + sqrtf(lam) * z + (z * z - 1.0f) * 0.16666667f;
  k = rintf(k);
  if (!(k > 1.0f)) k = 1.0f;
  if (k > 2.0e9f) k = 2.0e9f;
  return (int)k;
}

// Jump contribution k*(s-x) with k saturated so row sums cannot overflow int32.
__host__ __device__ __forceinline__ int jump_contrib(int k, int s, int x) {
  int kk = k > 4096 ? 4096 : k;
  return kk * (s - x);
}

// Inverse-CDF categorical draw from unnormalised weights: first index whose sequential fp32 cumulative
// sum exceeds v * total; falls back to the last positive weight. v in (0,1].
template <class F>
__device__ __forceinline__ int inv_cdf(int n, float v, F weight) {
  float tot = 0.f;
  for (int s = 0; s < n; ++s) tot += weight(s);
  const float target = fminf(v, 0.99999994f) * tot;
  float cum = 0.f;
  int last = 0;
  for (int s = 0; s < n; ++s) {
    const float w = weight(s);
    cum += w;
    if (w > 0.f) last = s;
    if (cum > target) return s;
  }
  return last;
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

}  // namespace ctdd
