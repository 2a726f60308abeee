// Evaluation metrics on device.
//
//   pair_kernel      exp(-bd * sum_d |x_i - y_j|) over all pairs of two sample sets, either written out as the (N, M)
//                    similarity matrix (binary_exp_hamming_sim / binary_hamming_sim, reference
//                    lib/datasets/metrics.py:6-22) or summed (the three sums of binary_mmd, metrics.py:25-48).
//                    The reference materialises the (N, M, D) difference tensor (2 GB at eval_synthetic's N = 4096,
//                    D = 32); here a 64 x 64 tile of pairs lives in registers, the rows pass through shared memory once
//                    per tile, and the self-similarity sums use only the upper triangle.
//   histogram_kernel per-dimension state counts of a sample set (the distributional parity check of north_star: KL of
//                    per-dimension histograms), shared-memory privatised when a dimension block's table fits.
//
// Arithmetic: distances are sums of |differences| of small integers, exact in fp32 whatever the order; exp is expf (the
// torch CUDA op); sums are accumulated in fp64 per thread, reduced per block in a fixed order into a partials array and
// finished by one thread block in a fixed order -> deterministic, and closer to the exact sum than the reference's fp32
// reduction.
#include "ctdd_common.cuh"

namespace ctdd {
namespace {

constexpr int TILE = 64;      // pairs tile: TILE x TILE, 256 threads, 4 x 4 pairs per thread
constexpr int KC = 32;        // dimensions staged per pass

template <bool MATRIX>
__global__ void __launch_bounds__(256) pair_kernel(const float* __restrict__ X, int N, const float* __restrict__ Y, int M,
                                                   int D, float bd, int self, int hamming_sim,
                                                   float* __restrict__ K, double* __restrict__ partials) {
  __shared__ float xs[KC][TILE + 1];
  __shared__ float ys[KC][TILE + 1];
  __shared__ double red[8];
  const int tiles_j = (M + TILE - 1) / TILE;
  const int ti = blockIdx.x / tiles_j, tj = blockIdx.x % tiles_j;
  const int tid = threadIdx.x;
  double acc = 0.0;
  // self-similarity sums: only tiles on or above the diagonal carry work (block-uniform branch)
  const bool skip = !MATRIX && self && tj < ti;
  if (!skip) {
    const int i0 = ti * TILE, j0 = tj * TILE;
    const int li = tid >> 4, lj = tid & 15;       // pair (a, b) of this thread: rows li + 16a, lj + 16b
    float d[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) d[a][b] = 0.f;
    for (int k0 = 0; k0 < D; k0 += KC) {
      // stage KC dimensions of the 64 + 64 rows, transposed ([k][row]) so the inner loop reads are conflict free
      for (int e = tid; e < TILE * KC; e += 256) {
        const int r = e / KC, k = e % KC;
        const bool kin = k0 + k < D;
        xs[k][r] = (kin && i0 + r < N) ? X[(size_t)(i0 + r) * D + k0 + k] : 0.f;
        ys[k][r] = (kin && j0 + r < M) ? Y[(size_t)(j0 + r) * D + k0 + k] : 0.f;
      }
      __syncthreads();
      const int kn = min(KC, D - k0);
      for (int k = 0; k < kn; ++k) {
        float xv[4], yv[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) { xv[a] = xs[k][li + 16 * a]; yv[a] = ys[k][lj + 16 * a]; }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) d[a][b] += fabsf(xv[a] - yv[b]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int i = i0 + li + 16 * a, j = j0 + lj + 16 * b;
        if (i >= N || j >= M) continue;
        const float v = hamming_sim ? (float)D - d[a][b] : expf(-bd * d[a][b]);
        if (MATRIX) {
          K[(size_t)i * M + j] = v;
        } else if (self) {
          if (j > i) acc += 2.0 * (double)v;       // (i, j) and (j, i); the diagonal is excluded (1 - eye)
        } else {
          acc += (double)v;
        }
      }
  }
  if (!MATRIX) {
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((tid & 31) == 0) red[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
      for (int w = 0; w < 8; ++w) s += red[w];
      partials[blockIdx.x] = s;
    }
  }
}

// out[slot] = sum of partials[0..n) in a fixed order (one block)
__global__ void __launch_bounds__(256) finish_sum_kernel(const double* __restrict__ partials, int n, double* __restrict__ out) {
  __shared__ double red[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += partials[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = red[0];
}

// counts[d*S + s] += #{n : x[n, d] == s}.  One block handles a (sample chunk, block of DB dimensions) pair with a
// shared-memory table of DB*S counters; out-of-range states are counted in counts[D*S] (must stay 0).
__global__ void __launch_bounds__(256) histogram_kernel(const int32_t* __restrict__ x, long long N, int D, int S, int DB,
                                                        long long rows_per_block, int* __restrict__ counts) {
  extern __shared__ int table[];
  const int dblocks = (D + DB - 1) / DB;
  const int db = blockIdx.x % dblocks;
  const long long chunk = blockIdx.x / dblocks;
  const int d0 = db * DB, dn = min(DB, D - d0);
  for (int e = threadIdx.x; e < dn * S; e += 256) table[e] = 0;
  __syncthreads();
  const long long n0 = chunk * rows_per_block, n1 = min(N, n0 + rows_per_block);
  int bad = 0;
  for (long long e = (n0 * dn) + threadIdx.x; e < n1 * dn; e += 256) {
    const long long n = e / dn;
    const int dd = (int)(e - n * dn);
    const int s = x[n * D + d0 + dd];
    if ((unsigned)s < (unsigned)S) atomicAdd(&table[dd * S + s], 1); else ++bad;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < dn * S; e += 256) {
    const int c = table[e];
    if (c) atomicAdd(&counts[(size_t)d0 * S + e], c);
  }
  if (bad) atomicAdd(&counts[(size_t)D * S], bad);
}

}  // namespace
}  // namespace ctdd

extern "C" int64_t ctdd_pair_partials(int N, int M) {
  const int64_t ti = (N + ctdd::TILE - 1) / ctdd::TILE, tj = (M + ctdd::TILE - 1) / ctdd::TILE;
  return ti * tj;
}

extern "C" int ctdd_pair_similarity(const float* X, int N, const float* Y, int M, int D, float bd, int hamming_sim,
                                    float* K, void* stream) {
  using namespace ctdd;
  if (N < 0 || M < 0 || D < 0) { set_error("ctdd_pair_similarity: negative size"); return 2; }
  if (N == 0 || M == 0) return 0;
  if (!X || !Y || !K) { set_error("ctdd_pair_similarity: null pointer"); return 2; }
  const int64_t blocks = ctdd_pair_partials(N, M);
  if (blocks > 0x7fffffffLL) { set_error("ctdd_pair_similarity: too many tiles"); return 2; }
  pair_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(X, N, Y, M, D, bd, 0, hamming_sim, K, nullptr);
  CTDD_CHECK_LAUNCH("pair_kernel<matrix>");
  return 0;
}

extern "C" int ctdd_pair_similarity_sum(const float* X, int N, const float* Y, int M, int D, float bd, int self,
                                        int hamming_sim, double* partials, double* out, void* stream) {
  using namespace ctdd;
  if (N < 0 || M < 0 || D < 0) { set_error("ctdd_pair_similarity_sum: negative size"); return 2; }
  if (!out) { set_error("ctdd_pair_similarity_sum: null output"); return 2; }
  if (self && (X != Y || N != M)) { set_error("ctdd_pair_similarity_sum: self needs X == Y"); return 2; }
  const int64_t blocks = ctdd_pair_partials(N, M);
  if (blocks > 0x7fffffffLL) { set_error("ctdd_pair_similarity_sum: too many tiles"); return 2; }
  if (blocks > 0) {
    if (!X || !Y || !partials) { set_error("ctdd_pair_similarity_sum: null pointer"); return 2; }
    pair_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(X, N, Y, M, D, bd, self, hamming_sim, nullptr, partials);
    CTDD_CHECK_LAUNCH("pair_kernel<sum>");
  }
  finish_sum_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(partials, (int)blocks, out);
  CTDD_CHECK_LAUNCH("finish_sum_kernel");
  return 0;
}

extern "C" int ctdd_state_histogram(const int32_t* x, int64_t N, int D, int S, int32_t* counts, void* stream) {
  using namespace ctdd;
  if (N < 0 || D <= 0 || S <= 0) { set_error("ctdd_state_histogram: bad size"); return 2; }
  if (N == 0) return 0;
  if (!x || !counts) { set_error("ctdd_state_histogram: null pointer"); return 2; }
  if (S > 8192) { set_error("ctdd_state_histogram: S > 8192 not supported"); return 2; }
  int DB = 8192 / S;                    // 32 KB of counters per block
  if (DB > D) DB = D;
  const int dblocks = (D + DB - 1) / DB;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long chunks = (4LL * sms + dblocks - 1) / dblocks;       // ~4 blocks per SM in total
  if (chunks > N) chunks = N;
  if (chunks < 1) chunks = 1;
  const long long rows_per_block = (N + chunks - 1) / chunks;
  chunks = (N + rows_per_block - 1) / rows_per_block;
  histogram_kernel<<<(unsigned)(chunks * dblocks), 256, (size_t)DB * S * sizeof(int), (cudaStream_t)stream>>>(
      x, N, D, S, DB, rows_per_block, counts);
  CTDD_CHECK_LAUNCH("histogram_kernel");
  return 0;
}
