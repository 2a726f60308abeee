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

// Backward of the head: d mu = sum_s g_s dlogit_s/dmu, d log_scale = sum_s g_s dlogit_s/dlog_scale.  One warp per row,
// lane-strided 128-bit reads of the incoming gradient (each read once), butterfly reduce.  With z' = dz/dmu = -inv,
// dz/dlog_scale = -z, dkappa/dlog_scale = -w (1 - kappa), d log u/dz = v, d log v/dz = -u, dv/dz = -u v, du/dz = u v:
//   logits_1 = log u_{s+1} + log A,  A = kappa v_s + 1e-6
//     d/dmu = inv (-v_{s+1} + kappa u_s v_s / A),   d/dls = -z_{s+1} v_{s+1} + (kappa u_s v_s z_s - w (1-kappa) v_s) / A
//   logits_2 = log v_s + log B,      B = kappa u_{s+1} + 1e-6
//     d/dmu = inv (u_s - kappa u_{s+1} v_{s+1} / B), d/dls = z_s u_s - (kappa u_{s+1} v_{s+1} z_{s+1} + w (1-kappa) u_{s+1}) / B
template <int VEC>
__global__ void __launch_bounds__(256) logistic_backward_kernel(const float* __restrict__ mu, const float* __restrict__ log_scale,
                                                                const float* __restrict__ grad_logits, long long rows, int D,
                                                                long long batch_stride, int S, int fix,
                                                                float* __restrict__ dmu, float* __restrict__ dls) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const float bw = 2.0f / (float)S;
  for (long long g = warp0; g < rows; g += nwarps) {
    long long src = g;
    if (batch_stride != (long long)D) {
      const long long n = g / D;
      src = n * batch_stride + (g - n * D);
    }
    const float m = __ldg(mu + src);
    const float inv = expf(2.0f - __ldg(log_scale + src));
    const float w = inv * bw;
    const float kap = -expm1f(-w);
    const float wk = w * (1.0f - kap);
    float acc_mu = 0.f, acc_ls = 0.f;
    for (int s0 = lane * VEC; s0 < S; s0 += 32 * VEC) {
      float gr[VEC];
      if (VEC == 4) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(grad_logits + g * S + s0));
        gr[0] = q.x; gr[1] = q.y; gr[2] = q.z; gr[3] = q.w;
      } else {
#pragma unroll
        for (int k = 0; k < VEC; ++k) gr[k] = __ldg(grad_logits + g * S + s0 + k);
      }
      float z[VEC + 1], lu[VEC + 1], lv[VEC + 1], u[VEC + 1], v[VEC + 1];
#pragma unroll
      for (int k = 0; k <= VEC; ++k) {
        z[k] = (fmaf((float)(s0 + k), bw, -1.0f) - m) * inv;
        logsig_pair(z[k], lu[k], lv[k], u[k], v[k]);
      }
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const float uv_l = u[k] * v[k], uv_r = u[k + 1] * v[k + 1];
        const float A = fmaf(kap, v[k], 1e-6f);
        float d_mu = -v[k + 1] + kap * uv_l / A;
        float d_ls = -z[k + 1] * v[k + 1] + (kap * uv_l * z[k] - wk * v[k]) / A;
        if (fix) {
          const float B = fmaf(kap, u[k + 1], 1e-6f);
          const float l1 = lu[k + 1] + __logf(A), l2 = lv[k] + __logf(B);
          if (l2 < l1) {
            d_mu = u[k] - kap * uv_r / B;
            d_ls = z[k] * u[k] - (kap * uv_r * z[k + 1] + wk * u[k + 1]) / B;
          }
        }
        acc_mu = fmaf(gr[k], d_mu, acc_mu);
        acc_ls = fmaf(gr[k], d_ls, acc_ls);
      }
    }
    acc_mu = warp_sum(acc_mu);
    acc_ls = warp_sum(acc_ls);
    if (lane == 0) {
      dmu[src] = acc_mu * inv;
      dls[src] = acc_ls;
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

extern "C" int ctdd_logistic_logits_backward(const float* mu, const float* log_scale, const float* grad_logits, int N, int D,
                                             int64_t batch_stride, int S, int fix_logistic, float* grad_mu,
                                             float* grad_log_scale, void* stream) {
  using namespace ctdd;
  if (!mu || !log_scale || !grad_logits || !grad_mu || !grad_log_scale) { set_error("ctdd_logistic_logits_backward: null pointer"); return 2; }
  if (N <= 0 || D <= 0 || S < 2 || batch_stride < D) { set_error("ctdd_logistic_logits_backward: bad sizes N=%d D=%d S=%d stride=%lld", N, D, S, (long long)batch_stride); return 2; }
  const long long rows = (long long)N * D;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long blocks = (rows + 7) / 8;      // 8 warps per block, one row per warp per iteration
  if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;
  const bool vec4 = (S % 4 == 0) && ((reinterpret_cast<uintptr_t>(grad_logits) & 15) == 0);
  if (vec4)
    logistic_backward_kernel<4><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(mu, log_scale, grad_logits, rows, D, batch_stride, S, fix_logistic, grad_mu, grad_log_scale);
  else
    logistic_backward_kernel<1><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(mu, log_scale, grad_logits, rows, D, batch_stride, S, fix_logistic, grad_mu, grad_log_scale);
  CTDD_CHECK_LAUNCH("logistic_backward_kernel");
  return 0;
}
