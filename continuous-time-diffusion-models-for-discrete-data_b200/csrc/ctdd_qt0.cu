// q_{t|0} builder: Q_b = U diag(exp(lam * ib_b)) Uinv, row-normalise, clamp; batched over time points.
// Replaces the two batched matmuls + diag_embed + boolean-mask clamp of
// lib/models/forward_model.py:265-287 (Gaussian), :108-126 (Uniform), :180-200 (UniformVariant), :51-75.
// Work is 2*S^3 FLOP per distinct time point (33 MFLOP at S=256) - negligible next to the reverse step, but the loss
// path builds one matrix per SAMPLE (B distinct times per training step).  Accumulation is fp64 over the fp32 factors
// so the result is the correctly rounded product.  S <= 256 runs qt0_fused_kernel: one CTA owns 16 complete rows of one
// matrix (4 x 4 fp64 register tile per thread, operands staged as doubles), so the row sums, the normalisation, the
// clamp and both the Q and the Q^T stores happen in the same kernel; larger S keeps the two-kernel form.
#include "ctdd_common.cuh"

namespace ctdd {

constexpr int QT_TILE = 16;

__global__ void __launch_bounds__(QT_TILE* QT_TILE)
qt0_eig_kernel(const float* __restrict__ U, const float* __restrict__ Uinv, const float* __restrict__ lam,
               const float* __restrict__ int_beta, int S, float* __restrict__ Q) {
  __shared__ float sU[QT_TILE][QT_TILE + 1];
  __shared__ float sV[QT_TILE][QT_TILE + 1];
  const int b = blockIdx.z;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int i = blockIdx.y * QT_TILE + ty;  // output row  (x_0)
  const int j = blockIdx.x * QT_TILE + tx;  // output col  (x_t)
  const float ib = int_beta[b];
  double acc = 0.0;
  for (int m0 = 0; m0 < S; m0 += QT_TILE) {
    const int mu = m0 + tx;  // column of U handled by this thread in the load
    const int mv = m0 + ty;  // row of Uinv
    float e = 0.f;
    if (mu < S) e = expf(lam[mu] * ib);  // same fp32 exp(adj_eigvals) as the reference
    sU[ty][tx] = (i < S && mu < S) ? U[(size_t)i * S + mu] * e : 0.f;
    sV[ty][tx] = (mv < S && j < S) ? Uinv[(size_t)mv * S + j] : 0.f;
    __syncthreads();
#pragma unroll
    for (int m = 0; m < QT_TILE; ++m) acc += (double)sU[ty][m] * (double)sV[m][tx];
    __syncthreads();
  }
  if (i < S && j < S) Q[((size_t)b * S + i) * S + j] = (float)acc;
}

// S <= 256.  grid (ceil(S/16), B); thread (tx, ty) = (tid & 63, tid >> 6) owns rows i0 + 4ty .. +3, columns 4tx .. +3.
// The m order of every dot product is ascending, as in qt0_eig_kernel: identical fp64 sums.
__global__ void __launch_bounds__(256)
qt0_fused_kernel(const float* __restrict__ U, const float* __restrict__ Uinv, const float* __restrict__ lam,
                 const float* __restrict__ int_beta, int S, int normalize, float clamp_below, float* __restrict__ Q,
                 float* __restrict__ QT) {
  __shared__ __align__(16) double sA[16][16];     // [m][row]   U[i][m] * exp(lam_m * ib)
  __shared__ __align__(16) double sB[16][256];    // [m][col]   Uinv[m][j]
  __shared__ double sred[4][2];
  const int b = blockIdx.y, i0 = blockIdx.x * 16;
  const int tid = threadIdx.x, tx = tid & 63, ty = tid >> 6;
  const float ib = int_beta[b];
  double acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
  for (int m0 = 0; m0 < S; m0 += 16) {
    {
      const int r = tid >> 4, m = tid & 15;
      float v = 0.f;
      if (i0 + r < S && m0 + m < S) v = U[(size_t)(i0 + r) * S + m0 + m] * expf(lam[m0 + m] * ib);   // fp32 product, as before
      sA[m][r] = (double)v;
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const int m = e, j = tid;
      sB[m][j] = (m0 + m < S && j < S) ? (double)Uinv[(size_t)(m0 + m) * S + j] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      const double2 a01 = *reinterpret_cast<const double2*>(&sA[m][4 * ty]);
      const double2 a23 = *reinterpret_cast<const double2*>(&sA[m][4 * ty + 2]);
      const double2 b01 = *reinterpret_cast<const double2*>(&sB[m][4 * tx]);
      const double2 b23 = *reinterpret_cast<const double2*>(&sB[m][4 * tx + 2]);
      const double av[4] = {a01.x, a01.y, a23.x, a23.y}, bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] += av[r] * bv[c];
    }
    __syncthreads();
  }
  float q[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) q[r][c] = (4 * tx + c < S) ? (float)acc[r][c] : 0.f;
  if (normalize) {
    // row sums in fp64 of the fp32-rounded entries: 64 threads (two warps) share a row group
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      double s = (double)q[r][0] + (double)q[r][1] + (double)q[r][2] + (double)q[r][3];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if ((tid & 31) == 0) sred[ty][(tid >> 5) & 1] = s;
      __syncthreads();
      const float inv = (float)(sred[ty][0] + sred[ty][1]);
      __syncthreads();
#pragma unroll
      for (int c = 0; c < 4; ++c) q[r][c] = q[r][c] / inv;
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (q[r][c] < clamp_below) q[r][c] = 0.f;
  const bool vec = (S & 3) == 0;
  const int j0 = 4 * tx, r0 = i0 + 4 * ty;
  float* Qb = Q + (size_t)b * S * S;
  float* QTb = QT ? QT + (size_t)b * S * S : nullptr;
  if (vec) {
    if (j0 < S) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (r0 + r < S) *reinterpret_cast<float4*>(Qb + (size_t)(r0 + r) * S + j0) = make_float4(q[r][0], q[r][1], q[r][2], q[r][3]);
      if (QTb && r0 < S) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<float4*>(QTb + (size_t)(j0 + c) * S + r0) = make_float4(q[0][c], q[1][c], q[2][c], q[3][c]);
      }
    }
  } else {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (r0 + r < S && j0 + c < S) {
          Qb[(size_t)(r0 + r) * S + j0 + c] = q[r][c];
          if (QTb) QTb[(size_t)(j0 + c) * S + r0 + r] = q[r][c];
        }
  }
}

// one warp per (b, row): optional normalisation by the row sum, clamp, write Q and Q^T
__global__ void qt0_finish_kernel(float* __restrict__ Q, float* __restrict__ QT, int S, int B, int normalize,
                                  float clamp_below) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B * S) return;
  const int b = warp / S, i = warp % S;
  float* row = Q + ((size_t)b * S + i) * S;
  float inv = 1.f;
  if (normalize) {
    double s = 0.0;
    for (int j = lane; j < S; j += 32) s += (double)row[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    inv = (float)s;
  }
  for (int j = lane; j < S; j += 32) {
    float v = row[j];
    if (normalize) v = v / inv;
    if (v < clamp_below) v = 0.f;
    row[j] = v;
    if (QT) QT[((size_t)b * S + j) * S + i] = v;
  }
}

__global__ void rate_scale_kernel(const float* __restrict__ Rb, const float* __restrict__ beta, int SS,
                                  float* __restrict__ out) {
  const int b = blockIdx.y;
  const float bt = beta[b];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < SS; i += gridDim.x * blockDim.x)
    out[(size_t)b * SS + i] = Rb[i] * bt;
}

}  // namespace ctdd

extern "C" int ctdd_build_qt0(const float* U, const float* Uinv, const float* lam, const float* int_beta,
                              int B, int S, int normalize, float clamp_below, float* Q_out, float* QT_out,
                              void* stream) {
  using namespace ctdd;
  if (!U || !Uinv || !lam || !int_beta || !Q_out) { set_error("ctdd_build_qt0: null pointer"); return 2; }
  if (B <= 0 || S <= 0 || S > 4096) { set_error("ctdd_build_qt0: bad sizes B=%d S=%d", B, S); return 2; }
  cudaStream_t st = (cudaStream_t)stream;
  if (S <= 256) {
    for (int b0 = 0; b0 < B; b0 += 65535) {
      const int nb = (B - b0) < 65535 ? (B - b0) : 65535;
      qt0_fused_kernel<<<dim3((S + 15) / 16, nb), 256, 0, st>>>(U, Uinv, lam, int_beta + b0, S, normalize, clamp_below,
                                                                Q_out + (size_t)b0 * S * S,
                                                                QT_out ? QT_out + (size_t)b0 * S * S : nullptr);
      CTDD_CHECK_LAUNCH("qt0_fused_kernel");
    }
    return 0;
  }
  const int tiles = (S + QT_TILE - 1) / QT_TILE;
  for (int b0 = 0; b0 < B; b0 += 65535) {
    const int nb = (B - b0) < 65535 ? (B - b0) : 65535;
    dim3 grid(tiles, tiles, nb), block(QT_TILE, QT_TILE);
    qt0_eig_kernel<<<grid, block, 0, st>>>(U, Uinv, lam, int_beta + b0, S, Q_out + (size_t)b0 * S * S);
    CTDD_CHECK_LAUNCH("qt0_eig_kernel");
  }
  const long long warps = (long long)B * S;
  const int threads = 256;
  const long long blocks = (warps * 32 + threads - 1) / threads;
  qt0_finish_kernel<<<(unsigned)blocks, threads, 0, st>>>(Q_out, QT_out, S, B, normalize, clamp_below);
  CTDD_CHECK_LAUNCH("qt0_finish_kernel");
  return 0;
}

extern "C" int ctdd_build_rate(const float* Rb, const float* beta, int B, int S, float* out, void* stream) {
  using namespace ctdd;
  if (!Rb || !beta || !out || B <= 0 || S <= 0) { set_error("ctdd_build_rate: bad arguments"); return 2; }
  dim3 grid((S * S + 255) / 256 > 64 ? 64 : (S * S + 255) / 256, B);
  rate_scale_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(Rb, beta, S * S, out);
  CTDD_CHECK_LAUNCH("rate_scale_kernel");
  return 0;
}
