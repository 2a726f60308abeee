// Multi-tensor EMA update of the shadow parameters: shadow <- shadow - (1 - decay) * (shadow - param) for every
// trainable tensor of the model in ONE launch.  Replaces the per-parameter Python loop of EMA.update_ema
// (reference lib/models/models.py:745-758: three torch kernels per parameter tensor, ~1500 launches for the CIFAR10
// U-Net).  HBM-bound: 12 bytes per parameter (read shadow, read param, write shadow).
//
// The caller keeps a device table of chunk descriptors (shadow pointer, param pointer, element count); a tensor is cut
// into chunks of at most CTDD_EMA_CHUNK elements so the grid is balanced whatever the tensor sizes are.  The arithmetic
// is the reference's, rounding by rounding (subtract, multiply, subtract — no fused multiply-add), so the result is
// bitwise what the torch loop produces.
#include "ctdd_common.cuh"

namespace ctdd {
namespace {

struct EmaChunk {
  float* shadow;
  const float* param;
  long long n;
};

__device__ __forceinline__ float ema_one(float s, float p, float omd) {
  return __fsub_rn(s, __fmul_rn(omd, __fsub_rn(s, p)));
}

__global__ void __launch_bounds__(256) ema_update_kernel(const EmaChunk* __restrict__ table, int n_chunks, float omd) {
  for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const EmaChunk ch = table[c];
    const bool vec = ((reinterpret_cast<uintptr_t>(ch.shadow) | reinterpret_cast<uintptr_t>(ch.param)) & 15) == 0;
    const long long n4 = vec ? ch.n >> 2 : 0;
    float4* s4 = reinterpret_cast<float4*>(ch.shadow);
    const float4* p4 = reinterpret_cast<const float4*>(ch.param);
    for (long long i = threadIdx.x; i < n4; i += blockDim.x) {
      float4 s = s4[i];
      const float4 p = __ldg(p4 + i);
      s.x = ema_one(s.x, p.x, omd); s.y = ema_one(s.y, p.y, omd);
      s.z = ema_one(s.z, p.z, omd); s.w = ema_one(s.w, p.w, omd);
      s4[i] = s;
    }
    for (long long i = 4 * n4 + threadIdx.x; i < ch.n; i += blockDim.x) ch.shadow[i] = ema_one(ch.shadow[i], ch.param[i], omd);
  }
}

}  // namespace
}  // namespace ctdd

extern "C" int64_t ctdd_ema_chunk_elems(void) { return CTDD_EMA_CHUNK; }

extern "C" int ctdd_ema_update(const void* chunk_table, int n_chunks, float one_minus_decay, void* stream) {
  using namespace ctdd;
  if (n_chunks == 0) return 0;
  if (!chunk_table || n_chunks < 0) { set_error("ctdd_ema_update: bad chunk table"); return 2; }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int blocks = n_chunks < sms * 8 ? n_chunks : sms * 8;
  ema_update_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const EmaChunk*>(chunk_table), n_chunks, one_minus_decay);
  CTDD_CHECK_LAUNCH("ema_update_kernel");
  return 0;
}
