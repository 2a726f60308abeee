// Initial-state sampling and forward noising x_t ~ q_{t|0}(.|x_0), x~ one-jump proposal.
// Replaces get_initial_samples (lib/sampling/sampling.py:14-28) and the noising blocks of
// lib/losses/losses.py:46-101 (copies at :326-381, :862-874, :1213-1225, :1281-1337, :1553-1593, :1834-1874).
// All draws are inverse-CDF on Philox uniforms with a sequential fp32 cumulative sum, restated in oracle/rng.py.
#include "ctdd_common.cuh"

namespace ctdd {

// First index whose entry of the non-decreasing array cum[0..n) exceeds target; n if none.
__device__ __forceinline__ int first_above(const float* cum, int n, float target) {
  int lo = 0, hi = n;                      // invariant: cum[i] <= target for i < lo, cum[i] > target for i >= hi
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (cum[mid] > target) hi = mid; else lo = mid + 1;
  }
  return lo;
}

// Every row draws from the SAME distribution: one thread forms the sequential fp32 cumulative sums once per CTA (the
// oracle's summation order), every thread then finds its crossing by binary search.
__global__ void categorical_shared_kernel(const float* __restrict__ prob, int S, long long rows,
                                          long long row_offset, unsigned long long seed,
                                          unsigned long long offset, int* __restrict__ x_out) {
  extern __shared__ float sp[];
  __shared__ int s_last;
  for (int i = threadIdx.x; i < S; i += blockDim.x) sp[i] = prob[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    float acc = 0.f;
    int last = 0;
    for (int s = 0; s < S; ++s) {
      const float w = sp[s];
      if (w > 0.f) last = s;
      acc += w;
      sp[s] = acc;
    }
    s_last = last;
  }
  __syncthreads();
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float v = u32_to_unit(philox_row_word((uint64_t)(row_offset + r), 0, offset, STREAM_INIT, seed));
  const float target = fminf(v, 0.99999994f) * sp[S - 1];
  const int f = first_above(sp, S, target);
  x_out[r] = f < S ? f : s_last;
}

// Forward noising, LANE per (b, d) row: a warp stages the 32 rows Q[b, x0, :] of its lanes into shared memory with
// coalesced loads (row stride S + 1 floats: the per-lane walks that follow are bank-conflict free), then every lane turns
// ITS row into sequential fp32 cumulative sums in place - the oracle's summation order - and finds the first entry above
// v * total by binary search (the weights are non-negative, so this is exactly inv_cdf's first crossing; its
// "last positive weight" fallback is only reachable for an all-zero row and is kept).
__global__ void __launch_bounds__(128) noise_xt_kernel(const float* __restrict__ Q, const int* __restrict__ x0, int B,
                                                      int D, int S, long long batch_offset,
                                                      unsigned long long seed, unsigned long long offset,
                                                      int* __restrict__ xt) {
  extern __shared__ float srow[];  // [warps][32][S + 1]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const long long rows = (long long)B * D;
  const int ld = S + 1;
  float* mine = srow + (size_t)warp * 32 * ld;
  for (long long r0 = ((long long)blockIdx.x * nwarp + warp) * 32; r0 < rows; r0 += (long long)gridDim.x * nwarp * 32) {
    const long long r = r0 + lane;
    const bool ok = r < rows;
    const int xs = ok ? x0[r] : 0;
    const long long qoff = ok ? ((long long)(r / D) * S + xs) * S : 0;
    for (int j = 0; j < 32; ++j) {         // row j of the batch: all lanes copy it, 128 bytes per step
      const long long qj = __shfl_sync(0xffffffffu, qoff, j);
      const float* q = Q + qj;
      float* dst = mine + (size_t)j * ld;
      for (int s = lane; s < S; s += 32) dst[s] = __ldg(q + s);
    }
    __syncwarp();
    float* row = mine + (size_t)lane * ld;
    float acc = 0.f;
    int last = 0;
#pragma unroll 8
    for (int s = 0; s < S; ++s) {
      const float w = row[s];
      acc += w;
      if (w > 0.f) last = s;
      row[s] = acc;
    }
    if (ok) {
      const long long grow = batch_offset * D + r;
      const float v = u32_to_unit(philox_row_word((uint64_t)grow, 0, offset, STREAM_NOISE_XT, seed));
      const float target = fminf(v, 0.99999994f) * acc;
      const int f = first_above(row, S, target);
      xt[r] = f < S ? f : last;
    }
    __syncwarp();
  }
}

// Forward noising, CTA per (SAMPLE b, block of NOISE_PASS rows of Q[b]) - used when D >= S: the D rows of a sample
// share S distinct rows of Q[b].  The CTA stages rows Q[b, k0 .. k0 + NOISE_PASS, :] with coalesced loads, one thread per
// staged row turns it into its sequential fp32 cumulative sums ONCE (instead of once per (b, d) row that starts there:
// D / S times less scanning and Q traffic), and every (b, d) row whose x0 falls into the staged range does its binary
// search there.  Same sums, same crossing as noise_xt_kernel.
constexpr int NOISE_PASS = 64, NOISE_THREADS = 256;
static_assert((NOISE_PASS / (NOISE_THREADS / 32)) % 4 == 0, "rows per warp are staged four at a time");
__global__ void __launch_bounds__(NOISE_THREADS) noise_xt_sample_kernel(const float* __restrict__ Q, const int* __restrict__ x0,
                                                                      int B, int D, int S, long long batch_offset,
                                                                      unsigned long long seed, unsigned long long offset,
                                                                      int* __restrict__ xt) {
  extern __shared__ float srow[];      // [NOISE_PASS][S + 1]
  __shared__ int s_last[NOISE_PASS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int RPW = NOISE_PASS / (NOISE_THREADS / 32);      // rows staged per warp
  const int ld = S + 1;
  const int npass = (S + NOISE_PASS - 1) / NOISE_PASS;
  for (long long item = blockIdx.x; item < (long long)B * npass; item += gridDim.x) {
    const int b = (int)(item / npass), k0 = (int)(item - (long long)b * npass) * NOISE_PASS;
    const float* Qb = Q + (size_t)b * S * S;
    const int* xb = x0 + (size_t)b * D;
    const int nk = (S - k0) < NOISE_PASS ? (S - k0) : NOISE_PASS;
    // 4 rows x 8 columns-of-32 per lane in flight (the copy is latency-bound otherwise)
    for (int j0 = 0; j0 < RPW; j0 += 4) {
      for (int s0 = 0; s0 < S; s0 += 256) {
        float v[4][8];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int i = RPW * warp + j0 + jj;
          const float* q = Qb + (size_t)(k0 + (i < nk ? i : 0)) * S;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int sidx = s0 + 32 * u + lane;
            v[jj][u] = (i < nk && sidx < S) ? __ldg(q + sidx) : 0.f;
          }
        }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int i = RPW * warp + j0 + jj;
          float* dst = srow + (size_t)i * ld;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int sidx = s0 + 32 * u + lane;
            if (i < nk && sidx < S) dst[sidx] = v[jj][u];
          }
        }
      }
    }
    __syncthreads();
    if ((int)threadIdx.x < nk) {
      float* row = srow + (size_t)threadIdx.x * ld;
      float acc = 0.f;
      int last = 0;
#pragma unroll 8
      for (int s = 0; s < S; ++s) {
        const float w = row[s];
        acc += w;
        if (w > 0.f) last = s;
        row[s] = acc;
      }
      s_last[threadIdx.x] = last;
    }
    __syncthreads();
    for (int d0 = threadIdx.x; d0 < D; d0 += 8 * NOISE_THREADS) {
      int xs[8];                         // 8 independent loads in flight (one load per trip left the loop latency-bound)
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int d = d0 + u * NOISE_THREADS;
        xs[u] = d < D ? __ldg(xb + d) : -1;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int d = d0 + u * NOISE_THREADS;
        const unsigned rel = (unsigned)(xs[u] - k0);
        if (xs[u] >= 0 && rel < (unsigned)nk) {
          const float* row = srow + (size_t)rel * ld;
          const long long grow = (batch_offset + b) * D + d;
          const float v = u32_to_unit(philox_row_word((uint64_t)grow, 0, offset, STREAM_NOISE_XT, seed));
          const float target = fminf(v, 0.99999994f) * row[S - 1];
          const int f = first_above(row, S, target);
          xt[(size_t)b * D + d] = f < S ? f : s_last[rel];
        }
      }
    }
    __syncthreads();
  }
}

// one CTA per sample b: w[d] = beta_b * sum_{s != xt[d]} Rb[xt[d], s]; d* by inverse CDF over d;
// new value by inverse CDF over Rb[xt[d*], .] with the diagonal removed.  The cumulative sums are sequential fp32 sums
// (the oracle's order) made by ONE thread over shared-memory arrays with 128-bit accesses; the crossing is a binary search.
__device__ __forceinline__ int inv_cdf_smem(float* w, int n, float v) {   // w -> cumulative sums in place; n % 4 == 0 padded
  float acc = 0.f;
  int last = 0;
  float4* w4 = reinterpret_cast<float4*>(w);
#pragma unroll 4
  for (int i = 0; i < (n + 3) / 4; ++i) {
    float4 q = w4[i];
    if (q.x > 0.f) last = 4 * i;
    acc += q.x; q.x = acc;
    if (q.y > 0.f) last = 4 * i + 1;
    acc += q.y; q.y = acc;
    if (q.z > 0.f) last = 4 * i + 2;
    acc += q.z; q.z = acc;
    if (q.w > 0.f) last = 4 * i + 3;
    acc += q.w; q.w = acc;
    w4[i] = q;
  }
  const float target = fminf(v, 0.99999994f) * acc;
  const int f = first_above(w, n, target);
  return f < n ? f : last;
}

__global__ void __launch_bounds__(256) xtilde_kernel(const float* __restrict__ Rb, const float* __restrict__ beta,
                                                    const int* __restrict__ xt, int D, int S, long long batch_offset,
                                                    unsigned long long seed, unsigned long long offset,
                                                    int* __restrict__ x_tilde) {
  extern __shared__ __align__(16) float sw[];  // [S4] off-diagonal row sums, later the value weights; then [D4] weights
  const int S4 = (S + 3) & ~3, D4 = (D + 3) & ~3;
  float* soff = sw;
  float* swd = sw + S4;
  __shared__ int s_dstar;
  const int b = blockIdx.x;
  const float bt = beta[b];
  // soff[x] = sum_{s != x} Rb[x][s] * beta_b, s ascending (the oracle's order); the rows are read through a 32 x 33
  // tile per warp so that the global loads are coalesced (thread-per-row loads touched 32 lines each)
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    float* tile = sw + S4 + D4 + (size_t)warp * 32 * 33;
    for (int xb0 = 32 * warp; xb0 < S; xb0 += 32 * nwarp) {
      const int x = xb0 + lane;
      float acc = 0.f;
      for (int s0 = 0; s0 < S; s0 += 32) {
#pragma unroll 8
        for (int j = 0; j < 32; ++j)
          tile[j * 33 + lane] = (xb0 + j < S && s0 + lane < S) ? __ldg(Rb + (size_t)(xb0 + j) * S + s0 + lane) : 0.f;
        __syncwarp();
#pragma unroll 8
        for (int l = 0; l < 32; ++l) {
          const int sidx = s0 + l;
          if (sidx < S) acc += (sidx == x) ? 0.f : (tile[lane * 33 + l] * bt);
        }
        __syncwarp();
      }
      if (x < S) soff[x] = acc;
    }
  }
  __syncthreads();
  const int* xrow = xt + (size_t)b * D;
  for (int d = threadIdx.x; d < D4; d += blockDim.x) {
    swd[d] = d < D ? soff[xrow[d]] : 0.f;          // zero padding: adds nothing, never selected
    if (d < D) x_tilde[(size_t)b * D + d] = xrow[d];
  }
  __syncthreads();
  const unsigned long long gb = (unsigned long long)(batch_offset + b);
  if (threadIdx.x == 0) {
    const float v1 = u32_to_unit(philox_row_word(gb, 0, offset, STREAM_TILDE_DIM, seed));
    int dstar = inv_cdf_smem(swd, D, v1);
    if (dstar >= D) dstar = D - 1;
    s_dstar = dstar;
  }
  __syncthreads();
  const int dstar = s_dstar;
  const int xs = xrow[dstar];
  const float* rrow = Rb + (size_t)xs * S;
  for (int s = threadIdx.x; s < S4; s += blockDim.x) soff[s] = (s < S && s != xs) ? rrow[s] * bt : 0.f;
  __syncthreads();
  if (threadIdx.x == 0) {
    const float v2 = u32_to_unit(philox_row_word(gb, 0, offset, STREAM_TILDE_VAL, seed));
    int nv = inv_cdf_smem(soff, S, v2);
    if (nv >= S) nv = S - 1;
    x_tilde[(size_t)b * D + dstar] = nv;
  }
}

}  // namespace ctdd

extern "C" int ctdd_sample_categorical_shared(const float* prob, int S, int64_t rows, int64_t row_offset,
                                              uint64_t seed, uint64_t offset, int32_t* x_out, void* stream) {
  using namespace ctdd;
  if (!prob || !x_out || S <= 0 || rows <= 0) { set_error("ctdd_sample_categorical_shared: bad arguments"); return 2; }
  if (S * sizeof(float) > 48 * 1024) { set_error("ctdd_sample_categorical_shared: S too large"); return 2; }
  const int threads = 1024;      // the per-CTA cumulative-sum prologue is amortised over 1024 rows
  const unsigned blocks = (unsigned)((rows + threads - 1) / threads);
  categorical_shared_kernel<<<blocks, threads, S * sizeof(float), (cudaStream_t)stream>>>(
      prob, S, rows, row_offset, seed, offset, x_out);
  CTDD_CHECK_LAUNCH("categorical_shared_kernel");
  return 0;
}

extern "C" int ctdd_noise_xt(const float* Q, const float* Rb, const float* beta, const int32_t* x0, int B, int D,
                             int S, int64_t batch_offset, uint64_t seed, uint64_t offset, int32_t* xt_out,
                             int32_t* x_tilde_out, void* stream) {
  using namespace ctdd;
  if (!Q || !x0 || !xt_out || B <= 0 || D <= 0 || S <= 1) { set_error("ctdd_noise_xt: bad arguments"); return 2; }
  if (x_tilde_out && (!Rb || !beta)) { set_error("ctdd_noise_xt: Rb/beta required for x_tilde"); return 2; }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t per_warp = (size_t)32 * (S + 1) * sizeof(float);
  if (per_warp > 200 * 1024) { set_error("ctdd_noise_xt: S too large"); return 2; }
  int warps = (int)((100 * 1024) / per_warp);       // two CTAs per SM
  warps = warps < 1 ? 1 : (warps > 4 ? 4 : warps);
  const size_t smem1 = warps * per_warp;
  static unsigned long long attr_done = 0ull;       // function attributes live in the device's context: a bit per device
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (dev < 0 || dev >= 64 || !((attr_done >> dev) & 1ull)) {
    cudaFuncSetAttribute(noise_xt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(noise_xt_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(xtilde_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (dev >= 0 && dev < 64) attr_done |= 1ull << dev;
  }
  const size_t smem_s = (size_t)(S < NOISE_PASS ? S : NOISE_PASS) * (S + 1) * sizeof(float);
  if (D >= S && smem_s <= 200 * 1024) {        // every row of Q[b] is scanned once
    const int per_sm = smem_s > 100 * 1024 ? 1 : (smem_s > 70 * 1024 ? 2 : (smem_s > 50 * 1024 ? 3 : 4));
    const long long items = (long long)B * ((S + NOISE_PASS - 1) / NOISE_PASS);
    long long blocks_s = (long long)sms * per_sm * 4;
    if (blocks_s > items) blocks_s = items;
    noise_xt_sample_kernel<<<(unsigned)blocks_s, NOISE_THREADS, smem_s, st>>>(Q, x0, B, D, S, batch_offset, seed, offset, xt_out);
    CTDD_CHECK_LAUNCH("noise_xt_sample_kernel");
  } else {
  const long long rows = (long long)B * D;
  long long blocks = (rows + warps * 32 - 1) / (warps * 32);
  const long long cap = (long long)sms * (smem1 > 100 * 1024 ? 1 : 2) * 4;   // a few waves; the kernel strides over row batches
  if (blocks > cap) blocks = cap;
  noise_xt_kernel<<<(unsigned)blocks, warps * 32, smem1, st>>>(Q, x0, B, D, S, batch_offset, seed, offset, xt_out);
  CTDD_CHECK_LAUNCH("noise_xt_kernel");
  }
  if (x_tilde_out) {
    const size_t smem2 = (size_t)(((S + 3) & ~3) + ((D + 3) & ~3) + 8 * 32 * 33) * sizeof(float);   // + one tile per warp
    if (smem2 > 200 * 1024) { set_error("ctdd_noise_xt: S + D too large for x_tilde"); return 2; }
    xtilde_kernel<<<B, 256, smem2, st>>>(Rb, beta, xt_out, D, S, batch_offset, seed, offset, x_tilde_out);
    CTDD_CHECK_LAUNCH("xtilde_kernel");
  }
  return 0;
}
