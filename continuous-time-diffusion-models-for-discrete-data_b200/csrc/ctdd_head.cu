// Truncated-logistic output head: (mu, log_scale) per dimension -> logits over S bins of [-1, 1].
// Replaces sample_logistic (reference lib/models/models.py:28-74; inline copy :248-282) with one HBM-bound pass.
//
// With z_j = (edge_j - mu) * exp(2 - log_scale) for the S+1 bin edges, u = sigmoid(z), v = sigmoid(-z),
// w = z_{s+1} - z_s and kappa = 1 - exp(-w), the reference's
//     logits_1 = logsig(z_{s+1}) + log1p(-exp(logsig(z_s) - logsig(z_{s+1})) + 1e-6)
// equals   log u_{s+1} + log(kappa * v_s + 1e-6)   exactly (1 - u_s/u_{s+1} = v_s * kappa), and the mirrored
//     logits_2 = log v_s + log(kappa * u_{s+1} + 1e-6).
// This form has no cancellation (the reference's 1 - exp(.) has), so it sits inside the reference's own fp32 noise
// around the fp64 value (tests/golden/head.npz keeps both).
#include "ctdd_common.cuh"

namespace ctdd {
namespace {

// log sigmoid(z), log sigmoid(-z), sigmoid(z), sigmoid(-z): stable for every z (e <= 1, no overflow, no cancellation)
__device__ __forceinline__ void logsig_pair(float z, float& log_u, float& log_v, float& u, float& v) {
  const float e = __expf(-fabsf(z));
  const float d = 1.0f + e;
  const float r = __fdividef(1.0f, d);
  const float ld = __logf(d);
  const bool pos = z >= 0.f;
  u = pos ? r : e * r;
  v = pos ? e * r : r;
  log_u = fminf(z, 0.f) - ld;
  log_v = fminf(-z, 0.f) - ld;
}

// one thread per VEC consecutive states of a row (VEC + 1 bin edges); row scalars are recomputed per thread
template <int VEC>
__global__ void __launch_bounds__(256) logistic_logits_kernel(const float* __restrict__ mu, const float* __restrict__ log_scale,
                                                              long long rows, int D, long long batch_stride, int S, int fix,
                                                              float* __restrict__ out) {
  const int per_row = S / VEC;
  const long long total = rows * per_row;
  const float bw = 2.0f / (float)S;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long g = i / per_row;
    const int s0 = (int)(i - g * per_row) * VEC;
    long long src = g;
    if (batch_stride != (long long)D) {
      const long long n = g / D;
      src = n * batch_stride + (g - n * D);
    }
    const float m = __ldg(mu + src);
    const float inv = expf(2.0f - __ldg(log_scale + src));
    const float kap = -expm1f(-inv * bw);
    float lu[VEC + 1], lv[VEC + 1], u[VEC + 1], v[VEC + 1];
#pragma unroll
    for (int k = 0; k <= VEC; ++k) {
      // bin edges are exact in fp32 (multiples of 2/S), like the reference's centers -/+ bin_width/2
      const float z = (fmaf((float)(s0 + k), bw, -1.0f) - m) * inv;
      logsig_pair(z, lu[k], lv[k], u[k], v[k]);
    }
    float o[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      o[k] = lu[k + 1] + __logf(fmaf(kap, v[k], 1e-6f));
      if (fix) o[k] = fminf(o[k], lv[k] + __logf(fmaf(kap, u[k + 1], 1e-6f)));
    }
    if (VEC == 4) {
      *reinterpret_cast<float4*>(out + g * S + s0) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
      for (int k = 0; k < VEC; ++k) out[g * S + s0 + k] = o[k];
    }
  }
}

}  // namespace
}  // namespace ctdd

extern "C" int ctdd_logistic_logits(const float* mu, const float* log_scale, int N, int D, int64_t batch_stride, int S,
                                    int fix_logistic, float* logits_out, void* stream) {
  using namespace ctdd;
  if (!mu || !log_scale || !logits_out) { set_error("ctdd_logistic_logits: null pointer"); return 2; }
  if (N <= 0 || D <= 0 || S < 2 || batch_stride < D) { set_error("ctdd_logistic_logits: bad sizes N=%d D=%d S=%d stride=%lld", N, D, S, (long long)batch_stride); return 2; }
  const long long rows = (long long)N * D, total = rows * S;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const bool vec4 = (S % 4 == 0) && ((reinterpret_cast<uintptr_t>(logits_out) & 15) == 0);
  long long blocks = ((vec4 ? total / 4 : total) + 255) / 256;
  if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;
  if (vec4)
    logistic_logits_kernel<4><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(mu, log_scale, rows, D, batch_stride, S, fix_logistic, logits_out);
  else
    logistic_logits_kernel<1><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(mu, log_scale, rows, D, batch_stride, S, fix_logistic, logits_out);
  CTDD_CHECK_LAUNCH("logistic_logits_kernel");
  return 0;
}
