// tcgen05 reverse-step kernel for S == 256 (configs C3/C4/C5): the (N*D x S)(S x S) contraction of
// lib/sampling/sampling.py:57 on 5th-generation tensor cores, fused with softmax, the q_{t|0} denominators, the
// forward-rate multiply and the Philox Poisson tau-leap (sampling.py:127-160).
//
// Formulation (transposed, "Z" in DESIGN.md):  D[s, row] = sum_k Q^T[s, k] * a[row, k]
//   * M side  = state s.  Q^T (bf16 hi + mid split, 2 x 128 TMEM columns) is loaded ONCE per CTA into tensor
//     memory and used as the A operand of tcgen05.mma (TS form) — it is never re-read from shared memory.
//   * N side  = data rows. Producer warps (one warp per row) read fp32 logits straight from HBM with coalesced
//     128-bit loads, do the row softmax with shuffles, multiply by the gathered reciprocal denominators
//     1/(Q[k,x]+eps), split to bf16 hi/mid and store the row K-major into a 128B-swizzled smem stage.
//   * 3 tensor passes  Qh*ah + Qh*am + Qm*ah  (dropped terms ~2^-16 relative, all terms non-negative) accumulate
//     in fp32 TMEM; epilogue warps read the accumulator with tcgen05.ld (lane = state s, column = row), scale by
//     the gathered forward rate R_b[s,x] and draw the per-(row,s) Poisson jump counts.
// Each CTA owns one half of the state axis (128 TMEM lanes); the two halves of a row are combined by a small
// finalize kernel.  Warp roles: 8 epilogue warps (2 per TMEM quadrant), 1 MMA-issue warp, 8 producer warps.
#include "ctdd_common.cuh"
#include <cuda_bf16.h>

namespace ctdd {
namespace tc {

constexpr int S = 256;
constexpr int NT = 64;                 // data rows per tile (= UMMA N)
constexpr int STAGES = 3;              // smem operand stages
constexpr int ACC = 2;                 // TMEM accumulator buffers
constexpr int RING = 8;                // per-tile side-info ring (>= STAGES + ACC + 1)
constexpr int NUM_EPI_WARPS = 8;       // warps e and e+4 share TMEM quadrant e&3 and split the tile's columns
constexpr int MMA_WARP = 8;
constexpr int FIRST_PROD_WARP = 9;
constexpr int NUM_PROD_WARPS = 8;
constexpr int NUM_THREADS = (FIRST_PROD_WARP + NUM_PROD_WARPS) * 32;  // 544
constexpr int ROWS_PER_PROD = NT / NUM_PROD_WARPS;                   // 8
constexpr int KBLOCK_BYTES = NT * 128;         // one 64-wide K block of one split: NT rows x 128 B
constexpr int SPLIT_BYTES = 4 * KBLOCK_BYTES;  // K = 256 -> 4 blocks
constexpr int STAGE_BYTES = 2 * SPLIT_BYTES;   // hi + mid
constexpr int TMEM_COLS = 512;
constexpr int TM_QH = 0, TM_QM = 128, TM_ACC = 256;  // TMEM column map

// per-time-point table blob (ctdd_prep_tc_tables)
constexpr size_t TAB_QH_OFF = 0;                                 // uint32 [256][128]  bf16 pairs of Q^T hi
constexpr size_t TAB_QM_OFF = TAB_QH_OFF + (size_t)S * 128 * 4;  // uint32 [256][128]  bf16 pairs of Q^T mid
constexpr size_t TAB_A_OFF = TAB_QM_OFF + (size_t)S * 128 * 4;   // float  [256][256]  tauLDR: 1/(Q[k,x]+eps); SDDM: Q[k,x]
constexpr size_t TAB_BYTES = TAB_A_OFF + (size_t)S * S * 4;
// static blob (ctdd_prep_tc_static)
constexpr size_t ST_RBZT_OFF = 0;                                // float [x][s] = Rb[s][x], zero at s == x
constexpr size_t ST_RBZ_OFF = (size_t)S * S * 4;                 // float [x][s] = Rb[x][s], zero at s == x
constexpr size_t ST_BYTES = 2 * (size_t)S * S * 4;

struct __align__(16) Side { float c1, c0; int x; float rs; };

struct Smem {
  alignas(1024) uint8_t stage[STAGES][STAGE_BYTES];
  Side side[RING][NT];
  int jump[RING][NT];
  int cnt[RING][NT];
  alignas(8) uint64_t full[STAGES];
  uint64_t empty[STAGES];
  uint64_t side_full[RING];
  uint64_t tmem_full[ACC];
  uint64_t tmem_empty[ACC];
  uint32_t tmem_base;
};

struct Args {
  int mode, branch, D, reject_multi;
  long long rows, row_offset;
  const float* logits;
  long long ld, batch_stride;
  const int* x_eval;
  const int* x_base;
  const uint8_t* tab;     // per-time-point blob
  const uint8_t* stat;    // static blob
  const float* RbT;       // [x][s] = Rb[s][x] (diagonal kept) for rr_out
  const float* Rb;        // [x][s] diagonal kept
  float beta, h;
  unsigned long long seed, offset;
  int* x_out;
  float* rr_out;
  float* ratio_out;
  unsigned long long* stats;
  int2* partial;          // [2][rows] (jump, count) per state half
  int num_tiles;
};

// ---------------------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)   // suspend-time hint: sleep in hardware, do not spin
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"((uint32_t)TMEM_COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"((uint32_t)TMEM_COLS) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]   (kind::f16, bf16 inputs, fp32 accumulate)
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, SWIZZLE_128B operand descriptor: LBO = 1 (unused), SBO = 1024 B (8 rows x 128 B), version 1
__device__ __forceinline__ uint64_t make_b_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float4 ld_stream(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// bf16 hi/mid split of two floats, packed (element 0 in the low half)
__device__ __forceinline__ void split2(float a0, float a1, uint32_t& hi, uint32_t& mid) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a0, a1);
  hi = *reinterpret_cast<uint32_t*>(&h);
  const float r0 = a0 - __uint_as_float(hi << 16);
  const float r1 = a1 - __uint_as_float(hi & 0xFFFF0000u);
  __nv_bfloat162 m = __floats2bfloat162_rn(r0, r1);
  mid = *reinterpret_cast<uint32_t*>(&m);
}

// rare path: the full 32-bit uniform is below lambda, so a jump is possible -> exact inverse CDF + accumulation
__device__ __noinline__ void jump_tail(float lam, float v, int s, int x, int* jump, int* cnt) {
  const int k = poisson_from_unit(lam, v);
  if (k) {
    atomicAdd(jump, jump_contrib(k, s, x));
    atomicAdd(cnt, k > 4096 ? 4096 : k);
  }
}

// ---------------------------------------------------------------------------------------------- the kernel
constexpr int PROD_BATCH = 4;          // rows whose loads are issued together by a producer warp
constexpr int PREFETCH_TILES = 3;      // L2 bulk-prefetch distance (tiles of this CTA's sequence)

// TAULDR: tauLDR rates (else SDDM reverse_prob); CORR: corrector adds R_t[x,:]; RATES: RATES_ONLY mode
template <bool TAULDR, bool CORR, bool RATES>
__global__ void __launch_bounds__(NUM_THREADS, 1) step_tc_kernel(const Args a) {
  extern __shared__ uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hs = blockIdx.x & 1;                 // state half owned by this CTA
  const int tile0 = blockIdx.x >> 1;
  const int tile_step = gridDim.x >> 1;
  const int my_tiles = (a.num_tiles > tile0) ? (a.num_tiles - tile0 + tile_step - 1) / tile_step : 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&sm.full[i], NUM_PROD_WARPS); mbar_init(&sm.empty[i], 1); }
    for (int i = 0; i < RING; ++i) mbar_init(&sm.side_full[i], NUM_PROD_WARPS);
    for (int i = 0; i < ACC; ++i) { mbar_init(&sm.tmem_full[i], 1); mbar_init(&sm.tmem_empty[i], NUM_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == MMA_WARP) tmem_alloc(&sm.tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;

  // Q^T halves -> tensor memory (A operand). Warp q < 4 owns TMEM lanes [32q, 32q+32).
  if (warp < 4) {
    const int srow = hs * 128 + warp * 32 + lane;
#pragma unroll 1
    for (int split = 0; split < 2; ++split) {
      const uint4* src = reinterpret_cast<const uint4*>(a.tab + (split ? TAB_QM_OFF : TAB_QH_OFF)) + (size_t)srow * 32;
#pragma unroll 1
      for (int c = 0; c < 128; c += 32) {
        uint32_t r[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint4 v = __ldg(src + (c >> 2) + i);
          r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
        }
        tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + (split ? TM_QM : TM_QH) + c, r);
      }
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp >= FIRST_PROD_WARP) {
    // ======================================================================== producers: one warp per row
    const int pw = warp - FIRST_PROD_WARP;
    // lane owns k = 4*lane .. 4*lane+3 and 128 + 4*lane .. +3: two fully coalesced 512-byte warp loads per row
    const float* tabA = reinterpret_cast<const float*>(a.tab + TAB_A_OFF) + 4 * lane;
    const float hb = RATES ? 1.0f : a.h * a.beta;   // rates-only ignores the step length
    const bool contiguous = (a.ld == S) && (a.batch_stride == (long long)a.D * S);
    const bool can_prefetch = contiguous && hs == 0 && pw == 0 && lane == 0;
    const uint32_t rows32 = (uint32_t)(a.rows < 0x7fffffffLL ? a.rows : 0x7fffffffLL);
    for (int i = 0; i < my_tiles; ++i) {
      const int tile = tile0 + i * tile_step;
      const int st = i % STAGES, slot = i % RING;
      if (can_prefetch) {   // pull a later tile of this CTA pair's sequence into L2 while this one is processed
        const long long pt = (long long)tile + (long long)PREFETCH_TILES * tile_step;
        if (pt < a.num_tiles) {
          const long long r0 = pt * NT;
          const long long nrow = (a.rows - r0) < NT ? (a.rows - r0) : NT;
          l2_prefetch_bulk(a.logits + r0 * S, (uint32_t)(nrow * S * 4));
        }
      }
      mbar_wait(&sm.empty[st], ((i / STAGES) & 1) ^ 1);
      uint8_t* stage = sm.stage[st];
#pragma unroll 1
      for (int b0 = 0; b0 < ROWS_PER_PROD; b0 += PROD_BATCH) {
        float4 v0[PROD_BATCH], v1[PROD_BATCH], t0[PROD_BATCH], t1[PROD_BATCH];
        int xr[PROD_BATCH];
        bool ok[PROD_BATCH];
        // issue every load of the batch before touching the data
#pragma unroll
        for (int j = 0; j < PROD_BATCH; ++j) {
          const int r = pw + NUM_PROD_WARPS * (b0 + j);
          const long long g = (long long)tile * NT + r;
          ok[j] = g < a.rows;
          const long long gc = ok[j] ? g : 0;
          const float* lp;
          if (contiguous) {
            lp = a.logits + gc * S + 4 * lane;
          } else {
            const uint32_t n = (uint32_t)gc / (uint32_t)a.D, d = (uint32_t)gc - n * (uint32_t)a.D;
            lp = a.logits + (long long)n * a.batch_stride + (long long)d * a.ld + 4 * lane;
          }
          v0[j] = ld_stream(lp);
          v1[j] = ld_stream(lp + 128);
          xr[j] = __ldg(a.x_eval + gc);
        }
#pragma unroll
        for (int j = 0; j < PROD_BATCH; ++j) {
          const float* tp = tabA + ((size_t)xr[j] << 8);
          t0[j] = __ldg(reinterpret_cast<const float4*>(tp));
          t1[j] = __ldg(reinterpret_cast<const float4*>(tp + 128));
        }
#pragma unroll
        for (int j = 0; j < PROD_BATCH; ++j) {
          const int r = pw + NUM_PROD_WARPS * (b0 + j);
          float v[8] = {v0[j].x, v0[j].y, v0[j].z, v0[j].w, v1[j].x, v1[j].y, v1[j].z, v1[j].w};
          const float t[8] = {t0[j].x, t0[j].y, t0[j].z, t0[j].w, t1[j].x, t1[j].y, t1[j].z, t1[j].w};
          float m = fmaxf(fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])), fmaxf(fmaxf(v[4], v[5]), fmaxf(v[6], v[7])));
          m = warp_max(m);
          const float ml = -m * 1.4426950408889634f;
          float sum = 0.f, dot = 0.f;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            v[q] = ex2_approx(fmaf(v[q], 1.4426950408889634f, ml));     // exp(v - max)
            sum += v[q];
            if (!TAULDR) dot = fmaf(v[q], t[q], dot);
          }
          sum = warp_sum(sum);
          const float rs = __frcp_rn(sum);
          Side si;
          if (TAULDR) {
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] *= t[q];             // e_k / (Q[k,x] + eps); 1/sum applied in the epilogue
            si.c1 = hb * rs * 0.0078125f;                          // lambda * 2^-7 = D * c1 * R_b[s,x]
            si.c0 = 0.f;
          } else {
            dot = warp_sum(dot);
            const float inv = __frcp_rn(fmaf(dot, rs, 1e-35f));    // 1 / (pQ[x] + 1e-35)
            si.c1 = hb * rs * inv * 0.0078125f;
            si.c0 = hb * 1e-35f * inv * 0.0078125f;
          }
          si.x = xr[j];
          si.rs = rs;
          uint32_t hi[4], mid[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) split2(v[2 * q], v[2 * q + 1], hi[q], mid[q]);
          if (!ok[j]) {
#pragma unroll
            for (int q = 0; q < 4; ++q) hi[q] = mid[q] = 0u;
            si.c1 = si.c0 = si.rs = 0.f;
            si.x = 0;
          }
          // k = 4*lane..+3 lives in K block lane/16, 16-byte chunk (lane%16)/2 (XOR-swizzled by the row), half lane&1;
          // k = 128 + 4*lane..+3 two K blocks further on
          const uint32_t off = (uint32_t)(lane >> 4) * KBLOCK_BYTES + (uint32_t)r * 128 +
                               (uint32_t)(((((lane & 15) >> 1) ^ (r & 7)) << 4) | ((lane & 1) << 3));
          *reinterpret_cast<uint2*>(stage + off) = make_uint2(hi[0], hi[1]);
          *reinterpret_cast<uint2*>(stage + off + 2 * KBLOCK_BYTES) = make_uint2(hi[2], hi[3]);
          *reinterpret_cast<uint2*>(stage + SPLIT_BYTES + off) = make_uint2(mid[0], mid[1]);
          *reinterpret_cast<uint2*>(stage + SPLIT_BYTES + off + 2 * KBLOCK_BYTES) = make_uint2(mid[2], mid[3]);
          if (lane == 0) {
            sm.side[slot][r] = si;
            sm.jump[slot][r] = 0;
            sm.cnt[slot][r] = 0;
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&sm.full[st]);
        mbar_arrive(&sm.side_full[slot]);
      }
    }
  } else if (warp == MMA_WARP) {
    // ======================================================================== MMA issue (one thread)
    if (lane == 0) {
      for (int i = 0; i < my_tiles; ++i) {
        const int st = i % STAGES, b = i % ACC;
        mbar_wait(&sm.full[st], (i / STAGES) & 1);
        mbar_wait(&sm.tmem_empty[b], ((i / ACC) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem + TM_ACC + b * NT;
        const uint32_t base = smem_u32(sm.stage[st]);
        uint32_t acc = 0;
#pragma unroll 1
        for (int pass = 0; pass < 3; ++pass) {
          const uint32_t a_tmem = tmem + (pass == 2 ? TM_QM : TM_QH);
          const uint32_t bsplit = base + (pass == 1 ? SPLIT_BYTES : 0);
#pragma unroll
          for (int k16 = 0; k16 < 16; ++k16) {
            const uint64_t bd = make_b_desc(bsplit + (k16 >> 2) * KBLOCK_BYTES + (k16 & 3) * 32);
            umma_ts(d_tmem, a_tmem + k16 * 8, bd, IDESC, acc);
            acc = 1;
          }
        }
        umma_commit(&sm.empty[st]);
        umma_commit(&sm.tmem_full[b]);
      }
    }
    __syncwarp();
  } else {
    // ======================================================================== epilogue: lane = state s
    const int q = warp & 3;                // TMEM quadrant
    const int ch = warp >> 2;              // which 32 columns of the tile
    const int s = hs * 128 + q * 32 + lane;
    const float* tabE = reinterpret_cast<const float*>(a.stat + (TAULDR ? ST_RBZT_OFF : ST_RBZ_OFF)) + s;
    const float* tabC = reinterpret_cast<const float*>(a.stat + ST_RBZ_OFF) + s;   // corrector add: Rb[x][s], zero diag
    const float* tabFull = (TAULDR ? a.RbT : a.Rb) + s;                             // diagonal kept, for rr_out
    const float hb7 = (RATES ? 1.0f : a.h * a.beta) * 0.0078125f;
    const uint32_t side_base = smem_u32(&sm.side[0][0]);
    const int c0 = 32 * ch;
    for (int i = 0; i < my_tiles; ++i) {
      const int tile = tile0 + i * tile_step;
      const int b = i % ACC, slot = i % RING;
      mbar_wait(&sm.side_full[slot], (i / RING) & 1);
      mbar_wait(&sm.tmem_full[b], (i / ACC) & 1);
      tc_fence_after();
      const long long g0 = (long long)tile * NT;
      const uint32_t side_slot = side_base + (uint32_t)(slot * NT + c0) * (uint32_t)sizeof(Side);
      uint32_t acc[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + TM_ACC + b * NT + c0, acc);
      tmem_ld_wait();
      tc_fence_before();   // accumulator is in registers: hand the buffer back to the MMA warp
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.tmem_empty[b]);
      if (RATES) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const long long g = g0 + c0 + j;
          if (g >= a.rows) continue;
          const float4 si = lds128(side_slot + (uint32_t)j * 16);
          const int x = __float_as_int(si.z);
          const float d = __uint_as_float(acc[j]);
          float ratio, rfull;
          if (TAULDR) {
            ratio = d * si.w;
            rfull = a.beta * __ldg(tabFull + ((size_t)x << 8)) * ratio;
          } else {
            // ratio = (pQ[s] + 1e-35) / (pQ[x] + 1e-35);  c1/c0 carry 2^-7
            ratio = fmaf(d, si.x, si.y) * 128.0f;
            rfull = ratio * (a.beta * __ldg(tabFull + ((size_t)x << 8)));
          }
          if (a.rr_out) a.rr_out[g * S + s] = rfull;
          if (a.ratio_out) a.ratio_out[g * S + s] = ratio;
        }
        continue;
      }
      // two passes of 16 columns keep the live register set small (544 threads -> 96 registers per thread)
#pragma unroll
      for (int h0 = 0; h0 < 32; h0 += 16) {
        // ---- phase 1 (branch-free, full ILP): lambda and the 16-bit pre-filter
        float lam7[16];
        Philox4 ph[2];
        uint32_t need = 0;
#pragma unroll
        for (int j8 = 0; j8 < 16; j8 += 8) {
          const uint64_t group = (uint64_t)(a.row_offset + g0 + c0 + h0 + j8) >> 3;
          ph[j8 >> 3] = philox_jump((uint32_t)s, group, a.offset, STREAM_JUMP_HI, a.seed);
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const float4 si = lds128(side_slot + (uint32_t)(h0 + j8 + jj) * 16);
            const size_t xo = (size_t)__float_as_int(si.z) << 8;
            float l7 = fmaf(__uint_as_float(acc[h0 + j8 + jj]), si.x, si.y) * __ldg(tabE + xo);
            if (CORR) l7 = fmaf(hb7, __ldg(tabC + xo), l7);
            lam7[j8 + jj] = l7;
            // halfword jj of the Philox output OR'ed into the mantissa of 1.0f: 1 + hi16 * 2^-23 (one PRMT)
            const uint32_t fb = __byte_perm(ph[j8 >> 3].w[jj >> 1], 0x3F800000u, (jj & 1) ? 0x7632 : 0x7610);
            // v >= hi16 * 2^-16 and P(K >= 1) <= lambda: when hi16 * 2^-16 >= lambda the count is certainly 0
            need |= ((__uint_as_float(fb) - 1.0f) < l7 ? 1u : 0u) << (j8 + jj);
          }
        }
        // ---- phase 2 (rare, grows with the jump rate): low 16 bits of the uniform, exact inverse CDF
        if (__any_sync(0xffffffffu, need != 0)) {
#pragma unroll
          for (int j8 = 0; j8 < 16; j8 += 8) {
            const uint32_t gm = (need >> j8) & 0xFFu;
            if (__any_sync(0xffffffffu, gm != 0)) {
              const uint64_t group = (uint64_t)(a.row_offset + g0 + c0 + h0 + j8) >> 3;
              const Philox4 lo = philox_jump((uint32_t)s, group, a.offset, STREAM_JUMP_LO, a.seed);
#pragma unroll
              for (int jj = 0; jj < 8; ++jj) {
                if ((gm >> jj) & 1u) {
                  const uint32_t w = (philox_half(ph[j8 >> 3], jj) << 16) | philox_half(lo, jj);
                  const float v = u32_to_unit(w);
                  const float lam = lam7[j8 + jj] * 128.0f;
                  if (v < lam) {
                    const int col = c0 + h0 + j8 + jj;
                    const int x = __float_as_int(lds128(side_slot + (uint32_t)(h0 + j8 + jj) * 16).z);
                    jump_tail(lam, v, s, x, &sm.jump[slot][col], &sm.cnt[slot][col]);
                  }
                }
              }
            }
          }
        }
      }
      epi_bar_sync();
      const int t = threadIdx.x;
      // the producers cannot reach this ring slot again before RING - (STAGES + ACC) more tiles have been drained
      if (t < NT && g0 + t < a.rows) a.partial[(size_t)hs * a.rows + g0 + t] = make_int2(sm.jump[slot][t], sm.cnt[slot][t]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem);
  }
}

// combine the two state halves of every row: x_new = clamp(x_base + jump), rejection, statistics
__global__ void step_tc_finalize_kernel(const int2* __restrict__ partial, const int* __restrict__ x_eval,
                                        const int* __restrict__ x_base, long long rows, int reject_multi,
                                        int* __restrict__ x_out, unsigned long long* stats) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int v[5] = {0, 0, 0, 0, 0};
  if (r < rows) {
    const int2 p0 = partial[r], p1 = partial[rows + r];
    int jump = p0.x + p1.x;
    const int cnt = p0.y + p1.y;
    const int xe = x_eval[r], xb = x_base ? x_base[r] : xe;
    v[3] = cnt > 0; v[4] = cnt > 1;
    if (reject_multi && cnt > 1) jump = 0;
    v[1] = jump != 0;
    int xn = xb + jump;
    xn = xn < 0 ? 0 : (xn > S - 1 ? S - 1 : xn);
    v[0] = xn != xb; v[2] = xn != xe;
    x_out[r] = xn;
  }
  if (stats) {   // block-level reduction: one atomic per counter per CTA
    __shared__ int red[5];
    if (threadIdx.x < 5) red[threadIdx.x] = 0;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      int s = v[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if ((threadIdx.x & 31) == 0 && s) atomicAdd(&red[i], s);
    }
    __syncthreads();
    if (threadIdx.x < 5 && red[threadIdx.x]) atomicAdd(stats + threadIdx.x, (unsigned long long)red[threadIdx.x]);
  }
}

// ---------------------------------------------------------------------------------------------- table prep
__global__ void prep_tables_kernel(const float* __restrict__ QT, int T, float eps, int branch, uint8_t* __restrict__ out) {
  const int t = blockIdx.y;
  const float* qt = QT + (size_t)t * S * S;
  uint8_t* blob = out + (size_t)t * TAB_BYTES;
  uint32_t* qh = reinterpret_cast<uint32_t*>(blob + TAB_QH_OFF);
  uint32_t* qm = reinterpret_cast<uint32_t*>(blob + TAB_QM_OFF);
  float* ta = reinterpret_cast<float*>(blob + TAB_A_OFF);
  // QT[s][k] = Q[k][s]: the A operand row s holds K contiguous -> pairs (k, k+1)
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < S * 128; i += gridDim.x * blockDim.x) {
    const float a0 = qt[2 * i], a1 = qt[2 * i + 1];
    uint32_t hi, mid;
    split2(a0, a1, hi, mid);
    qh[i] = hi;
    qm[i] = mid;
  }
  // TAB_A[x][k] = f(Q[k][x]) = f(QT[x][k])
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < S * S; i += gridDim.x * blockDim.x) {
    const float q = qt[i];
    ta[i] = (branch == CTDD_BRANCH_TAULDR) ? 1.0f / (q + eps) : q;
  }
}

__global__ void prep_static_kernel(const float* __restrict__ Rb, uint8_t* __restrict__ out) {
  float* rbzt = reinterpret_cast<float*>(out + ST_RBZT_OFF);
  float* rbz = reinterpret_cast<float*>(out + ST_RBZ_OFF);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < S * S; i += gridDim.x * blockDim.x) {
    const int x = i / S, s = i % S;
    rbzt[i] = (s == x) ? 0.f : Rb[(size_t)s * S + x];
    rbz[i] = (s == x) ? 0.f : Rb[i];
  }
}

}  // namespace tc

bool tc_supports(const ctdd_step_params* p) {
  if (p->S != tc::S) return false;
  if (!(p->branch == CTDD_BRANCH_TAULDR || p->branch == CTDD_BRANCH_SDDM_REVERSE_PROB)) return false;
  if (!(p->mode == CTDD_MODE_TAU_LEAP || p->mode == CTDD_MODE_TAU_LEAP_CORR || p->mode == CTDD_MODE_MIDPOINT_JUMP ||
        p->mode == CTDD_MODE_RATES_ONLY))
    return false;
  if (!p->tc_tables || !p->tc_static) return false;
  if (p->mode != CTDD_MODE_RATES_ONLY && !p->workspace) return false;
  if ((p->ld_logits & 3) || (p->batch_stride_logits & 3) || (reinterpret_cast<uintptr_t>(p->logits) & 15)) return false;
  return true;
}

long long tc_workspace_bytes(long long rows, int S) {
  if (S != tc::S) return 0;
  return 2 * rows * (long long)sizeof(int2);
}

int launch_step_tc(const ctdd_step_params* p, cudaStream_t st) {
  using namespace tc;
  static int num_sms = 0;
  static bool attr_set = false;
  const size_t smem_bytes = sizeof(Smem) + 1024;
  typedef void (*kern_t)(const Args);
  static const kern_t kerns[2][3] = {
      {step_tc_kernel<false, false, false>, step_tc_kernel<false, true, false>, step_tc_kernel<false, false, true>},
      {step_tc_kernel<true, false, false>, step_tc_kernel<true, true, false>, step_tc_kernel<true, false, true>}};
  if (!attr_set) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    for (int i = 0; i < 2; ++i)
      for (int j = 0; j < 3; ++j)
        if (cudaFuncSetAttribute(kerns[i][j], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes) != cudaSuccess) {
          set_error("ctdd_reverse_step: cannot reserve %zu bytes of shared memory for the tcgen05 kernel", smem_bytes);
          cudaGetLastError();
          return 1;
        }
    attr_set = true;
  }
  Args a;
  a.mode = p->mode; a.branch = p->branch; a.D = p->D; a.reject_multi = p->reject_multi;
  a.rows = (long long)p->N * p->D; a.row_offset = p->row_offset;
  a.logits = p->logits; a.ld = p->ld_logits; a.batch_stride = p->batch_stride_logits;
  a.x_eval = p->x_eval; a.x_base = p->x_base;
  a.tab = reinterpret_cast<const uint8_t*>(p->tc_tables);
  a.stat = reinterpret_cast<const uint8_t*>(p->tc_static);
  a.RbT = p->RbT; a.Rb = p->Rb; a.beta = p->beta; a.h = p->h; a.seed = p->seed; a.offset = p->offset;
  a.x_out = p->x_out; a.rr_out = p->rr_out; a.ratio_out = p->ratio_out;
  a.stats = reinterpret_cast<unsigned long long*>(p->stats_out);
  a.partial = reinterpret_cast<int2*>(p->workspace);
  a.num_tiles = (int)((a.rows + NT - 1) / NT);
  int grid = num_sms & ~1;                         // CTA pairs (state halves) share a tile sequence
  if (grid > 2 * a.num_tiles) grid = 2 * a.num_tiles;
  if (grid < 2) grid = 2;
  const int ki = (p->branch == CTDD_BRANCH_TAULDR) ? 1 : 0;
  const int kj = (p->mode == CTDD_MODE_RATES_ONLY) ? 2 : (p->mode == CTDD_MODE_TAU_LEAP_CORR ? 1 : 0);
  kerns[ki][kj]<<<grid, NUM_THREADS, smem_bytes, st>>>(a);
  CTDD_CHECK_LAUNCH("step_tc_kernel");
  if (p->mode != CTDD_MODE_RATES_ONLY) {
    const int threads = 1024;
    step_tc_finalize_kernel<<<(unsigned)((a.rows + threads - 1) / threads), threads, 0, st>>>(
        a.partial, a.x_eval, a.x_base, a.rows, a.reject_multi, a.x_out, a.stats);
    CTDD_CHECK_LAUNCH("step_tc_finalize_kernel");
  }
  return 0;
}

}  // namespace ctdd

extern "C" int64_t ctdd_tc_tables_bytes(int S) { return S == ctdd::tc::S ? (int64_t)ctdd::tc::TAB_BYTES : 0; }
extern "C" int64_t ctdd_tc_static_bytes(int S) { return S == ctdd::tc::S ? (int64_t)ctdd::tc::ST_BYTES : 0; }

extern "C" int ctdd_prep_tc_tables(const float* Q, const float* QT, const float* Rb, int T, int S, float eps,
                                   int branch, void* tables_out, void* stream) {
  using namespace ctdd;
  (void)Q; (void)Rb;
  if (S != tc::S) { set_error("ctdd_prep_tc_tables: the tcgen05 path needs S == 256 (got %d)", S); return 2; }
  if (!QT || !tables_out || T <= 0) { set_error("ctdd_prep_tc_tables: bad arguments"); return 2; }
  for (int t0 = 0; t0 < T; t0 += 65535) {
    const int nt = (T - t0) < 65535 ? (T - t0) : 65535;
    dim3 grid(32, nt);
    tc::prep_tables_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(QT + (size_t)t0 * S * S, nt, eps, branch,
                                                                  reinterpret_cast<uint8_t*>(tables_out) + (size_t)t0 * tc::TAB_BYTES);
    CTDD_CHECK_LAUNCH("prep_tables_kernel");
  }
  return 0;
}

extern "C" int ctdd_prep_tc_static(const float* Rb, int S, void* static_out, void* stream) {
  using namespace ctdd;
  if (S != tc::S) { set_error("ctdd_prep_tc_static: the tcgen05 path needs S == 256 (got %d)", S); return 2; }
  if (!Rb || !static_out) { set_error("ctdd_prep_tc_static: null pointer"); return 2; }
  tc::prep_static_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(Rb, reinterpret_cast<uint8_t*>(static_out));
  CTDD_CHECK_LAUNCH("prep_static_kernel");
  return 0;
}
