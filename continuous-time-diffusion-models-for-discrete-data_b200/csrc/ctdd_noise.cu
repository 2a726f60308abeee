// Initial-state sampling and forward noising x_t ~ q_{t|0}(.|x_0), x~ one-jump proposal.
// Replaces get_initial_samples (lib/sampling/sampling.py:14-28) and the noising blocks of
// lib/losses/losses.py:46-101 (copies at :326-381, :862-874, :1213-1225, :1281-1337, :1553-1593, :1834-1874).
// All draws are inverse-CDF on Philox uniforms with a sequential fp32 cumulative sum, restated in oracle/rng.py.
#include "ctdd_common.cuh"

namespace ctdd {

__global__ void categorical_shared_kernel(const float* __restrict__ prob, int S, long long rows,
                                          long long row_offset, unsigned long long seed,
                                          unsigned long long offset, int* __restrict__ x_out) {
  extern __shared__ float sp[];
  for (int i = threadIdx.x; i < S; i += blockDim.x) sp[i] = prob[i];
  __syncthreads();
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float v = u32_to_unit(philox_row_word((uint64_t)(row_offset + r), 0, offset, STREAM_INIT, seed));
  x_out[r] = inv_cdf(S, v, [&](int s) { return sp[s]; });
}

// one warp per (b, d): lanes stage the row Q[b, x0, :] into shared memory with coalesced loads, then
// lane 0 walks it sequentially so the fp32 cumulative sum has the oracle's summation order.
__global__ void __launch_bounds__(256) noise_xt_kernel(const float* __restrict__ Q, const int* __restrict__ x0, int B,
                                                      int D, int S, long long batch_offset,
                                                      unsigned long long seed, unsigned long long offset,
                                                      int* __restrict__ xt) {
  extern __shared__ float srow[];  // [warps][S]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const long long r = (long long)blockIdx.x * nwarp + warp;
  if (r >= (long long)B * D) return;
  const int b = (int)(r / D);
  const int xs = x0[r];
  const float* q = Q + ((size_t)b * S + xs) * S;
  float* my = srow + (size_t)warp * S;
  for (int s = lane; s < S; s += 32) my[s] = __ldg(q + s);
  __syncwarp();
  if (lane == 0) {
    const long long grow = batch_offset * D + r;
    const float v = u32_to_unit(philox_row_word((uint64_t)grow, 0, offset, STREAM_NOISE_XT, seed));
    xt[r] = inv_cdf(S, v, [&](int s) { return my[s]; });
  }
}

// one CTA per sample b: w[d] = beta_b * sum_{s != xt[d]} Rb[xt[d], s]; d* by inverse CDF over d;
// new value by inverse CDF over Rb[xt[d*], .] with the diagonal removed.
__global__ void __launch_bounds__(256) xtilde_kernel(const float* __restrict__ Rb, const float* __restrict__ beta,
                                                    const int* __restrict__ xt, int D, int S, long long batch_offset,
                                                    unsigned long long seed, unsigned long long offset,
                                                    int* __restrict__ x_tilde) {
  extern __shared__ float sw[];  // [S] off-diagonal row sums, then [D] weights
  float* soff = sw;
  float* swd = sw + S;
  const int b = blockIdx.x;
  const float bt = beta[b];
  for (int x = threadIdx.x; x < S; x += blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < S; ++s) acc += (s == x) ? 0.f : (Rb[(size_t)x * S + s] * bt);
    soff[x] = acc;
  }
  __syncthreads();
  const int* xrow = xt + (size_t)b * D;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    swd[d] = soff[xrow[d]];
    x_tilde[(size_t)b * D + d] = xrow[d];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long gb = (unsigned long long)(batch_offset + b);
    const float v1 = u32_to_unit(philox_row_word(gb, 0, offset, STREAM_TILDE_DIM, seed));
    const int dstar = inv_cdf(D, v1, [&](int d) { return swd[d]; });
    const int xs = xrow[dstar];
    const float v2 = u32_to_unit(philox_row_word(gb, 0, offset, STREAM_TILDE_VAL, seed));
    const float* rrow = Rb + (size_t)xs * S;
    const int nv = inv_cdf(S, v2, [&](int s) { return s == xs ? 0.f : rrow[s] * bt; });
    x_tilde[(size_t)b * D + dstar] = nv;
  }
}

}  // namespace ctdd

extern "C" int ctdd_sample_categorical_shared(const float* prob, int S, int64_t rows, int64_t row_offset,
                                              uint64_t seed, uint64_t offset, int32_t* x_out, void* stream) {
  using namespace ctdd;
  if (!prob || !x_out || S <= 0 || rows <= 0) { set_error("ctdd_sample_categorical_shared: bad arguments"); return 2; }
  if (S * sizeof(float) > 48 * 1024) { set_error("ctdd_sample_categorical_shared: S too large"); return 2; }
  const int threads = 256;
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
  const int warps = 8;
  const size_t smem1 = (size_t)warps * S * sizeof(float);
  if (smem1 > 48 * 1024) { set_error("ctdd_noise_xt: S too large"); return 2; }
  const long long rows = (long long)B * D;
  noise_xt_kernel<<<(unsigned)((rows + warps - 1) / warps), warps * 32, smem1, st>>>(Q, x0, B, D, S, batch_offset,
                                                                                     seed, offset, xt_out);
  CTDD_CHECK_LAUNCH("noise_xt_kernel");
  if (x_tilde_out) {
    const size_t smem2 = (size_t)(S + D) * sizeof(float);
    if (smem2 > 48 * 1024) { set_error("ctdd_noise_xt: S + D too large for x_tilde"); return 2; }
    xtilde_kernel<<<B, 256, smem2, st>>>(Rb, beta, xt_out, D, S, batch_offset, seed, offset, x_tilde_out);
    CTDD_CHECK_LAUNCH("xtilde_kernel");
  }
  return 0;
}
