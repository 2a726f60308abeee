// Diagnostic (not part of the product): issue rate of tcgen05.mma.cta_group::2 (M = 256 over a CTA pair, K = 16, bf16)
// with the A operand in tensor memory (TS form, what the reverse-step kernel uses) or in shared memory (SS form), for
// N = 64 / 128 / 256, alone and with four warps per CTA reading the accumulator with tcgen05.ld at the same time.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I continuous-time-diffusion-models-for-discrete-data_b200/csrc -I include \
//        -o tools/build/mma_bench tools/mma_bench.cu
#include "ctdd_tc_common.cuh"
#include <cstdio>

using namespace ctdd::tc;

__device__ __forceinline__ void umma_ss_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

struct Res { long long cyc; long long lds; };

template <int N, bool SS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1) bench(Res* out, int tiles, int with_ld, int passes_ts) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = base;                 // 2 x 64 KB: "Q hi / Q mid" halves of this CTA, K-major SW128 (4 K blocks of 16 KB)
  uint8_t* sB = base + 131072;        // 64 KB stage area
  __shared__ alignas(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int done;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  for (int i = threadIdx.x; i < (131072 + 65536) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0x3c003c00u + i;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); done = 0; fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  fence_proxy_async();
  cluster_sync_all();
  constexpr uint32_t IDESC = make_idesc(N);
  constexpr int ACC_COLS = N;         // accumulator columns per buffer
  constexpr int NACC = (256 / ACC_COLS) > 2 ? 2 : (256 / ACC_COLS);
  long long lds = 0;
  if (warp == 0) {
    if (rank == 0 && lane == 0) {
      const long long t0 = clock64();
      for (int i = 0; i < tiles; ++i) {
        const uint32_t d = tmem + 256 + (i % NACC) * ACC_COLS;
        for (int pass = 0; pass < 3; ++pass) {
          const bool ss = SS && pass < passes_ts;     // SS builds: the first `passes_ts` passes read A from shared memory
#pragma unroll
          for (int k16 = 0; k16 < 16; ++k16) {
            const uint32_t boff = (uint32_t)((k16 >> 2) * (N / 2) * 128 + (k16 & 3) * 32);
            const uint64_t bd = make_b_desc(smem_u32(sB) + (pass == 1 ? 16384u : 0u) + boff % 16384u);
            if (ss) {
              const uint64_t ad = make_b_desc(smem_u32(sA) + (uint32_t)((k16 >> 2) * 16384 + (k16 & 3) * 32));
              umma_ss_pair(d, ad, bd, IDESC, (pass | k16) ? 1u : 0u);
            } else {
              umma_ts_pair(d, tmem + (pass == 2 ? 128 : 0) + k16 * 8, (uint32_t)bd, (uint32_t)(bd >> 32), IDESC, (pass | k16) ? 1u : 0u);
            }
          }
        }
      }
      umma_commit_pair(&bar);
      mbar_wait(&bar, 0);
      const long long t1 = clock64();
      done = 1;
      out[blockIdx.x >> 1].cyc = t1 - t0;
    } else if (lane == 0) {
      mbar_wait(&bar, 0);
      done = 1;
    }
    __syncwarp();
  } else if (warp <= 4 && with_ld) {
    const int q = warp & 3;
    while (!done) {
      uint32_t r[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + 256 + (lds & 7) * 32, r);
      tmem_ld_wait();
      if (r[lane] == 0x12345678u) done = 2;     // keep the load alive
      ++lds;
    }
  }
  __syncthreads();
  if (warp == 1 && lane == 0 && rank == 0) out[blockIdx.x >> 1].lds = lds;
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem); }
}

template <int N, bool SS>
void run(const char* name, int with_ld, int passes_ss) {
  Res* d;
  cudaMalloc(&d, 74 * sizeof(Res));
  const int smem = 131072 + 65536 + 1024;
  cudaFuncSetAttribute(bench<N, SS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int tiles = 200;
  for (int rep = 0; rep < 2; ++rep) {
    cudaMemset(d, 0, 74 * sizeof(Res));
    bench<N, SS><<<148, 192, smem>>>(d, tiles, with_ld, passes_ss);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  }
  Res h[74];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double mean = 0, mx = 0, lds = 0;
  for (int i = 0; i < 74; ++i) { mean += h[i].cyc; if (h[i].cyc > mx) mx = h[i].cyc; lds += h[i].lds; }
  mean /= 74; lds /= 74;
  printf("%-28s N=%3d ld=%d ss_passes=%d: %7.1f cycles/MMA (max %7.1f), %6.0f cycles per 128 rows, tcgen05.ld per tile and warp %.1f\n", name, N,
         with_ld, passes_ss, mean / (tiles * 48.0), mx / (tiles * 48.0), mean / tiles * 128.0 / N, lds / tiles);
  cudaFree(d);
}

int main() {
  for (int ld = 0; ld < 2; ++ld) {
    run<64, false>("TS", ld, 0);
    run<128, false>("TS", ld, 0);
    run<256, false>("TS", ld, 0);
    run<64, true>("SS (all passes)", ld, 3);
    run<128, true>("SS (all passes)", ld, 3);
    run<256, true>("SS (all passes)", ld, 3);
    run<128, true>("SS hi passes, TS mid pass", ld, 2);
    run<256, true>("SS hi passes, TS mid pass", ld, 2);
  }
  return 0;
}
