// bf16 GEMM on the 5th-generation tensor cores (sm_100a): TMA -> shared memory (128B swizzle) ->
// tcgen05.mma (accumulator in TMEM) -> tcgen05.ld epilogue with fused alpha / bias / ReLU / residual.
//
//   C[M,N] = relu?(alpha * op(A) op(B) + bias) + R        A,B bf16; C,R bf16 or fp32; fp32 accumulate
//
// One CTA computes one 128 x BN output tile (x one K split).  Warp roles (256 threads):
//   warp 0  : TMA producer  — one elected lane issues cp.async.bulk.tensor loads into a 4-stage ring
//   warp 1  : MMA issuer    — one elected lane issues tcgen05.mma (M=128, N=BN, K=16) x 4 per stage and
//                             releases the stage with tcgen05.commit
//   warp 2  : TMEM allocator (BN fp32 columns)
//   warps 4-7: epilogue     — each thread owns one accumulator row (TMEM lane), reads 32 columns at a time
// All four transpose forms are served by the same kernel: an operand stored with K contiguous is staged
// K-major (box 128 rows x 64 k), an operand stored with M/N contiguous is staged MN-major (boxes of
// 64 k-rows x 64 mn) and the UMMA shared-memory / instruction descriptors carry the major-ness, so no
// transposed copies of activations or weights are ever made.  Out-of-range rows/cols/k are zero-filled by TMA.
// Split-K (grid.z) with fp32 atomics covers the weight-gradient shapes (small M x N, K = T*B).
#include <cuda.h>
#include <cstdlib>

#include "common.cuh"
#include "umma.cuh"

namespace b200st {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;
constexpr int TC_THREADS = 256;

template <int BN, int STAGES, int BM = TC_BM>
struct TcSmem {
  static constexpr int A_BYTES = BM * TC_BK * 2;
  static constexpr int B_BYTES = BN * TC_BK * 2;
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE;
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 1) * 8 + 16 + 1024;   // + alignment slack
};

// Tile configurations (BN, STAGES): the deep-pipeline / narrow-N ones exist because most GEMMs on this path are
// latency-bound (M = 64 decoder steps, M ~ 2-3k Transformer rows with K = 512): what matters there is bytes in
// flight per SM and the number of CTAs, not MMA throughput.
// BM = 128, or 64 for single-M-tile problems with M <= 64 (the decoder steps): the SS-form MMA is paced by streaming
// the A tile out of shared memory, so not reading 64 zero-padded rows halves the per-MMA cost.  For M = 64 the
// accumulator row i lives in TMEM lane (i / 16) * 32 + (i % 16) (measured: scripts/probes/umma_m64_probe.cu).
template <int BN, int STAGES, bool A_MN, bool B_MN, typename TC, bool ATOMIC, int BM = TC_BM>
__global__ void __launch_bounds__(TC_THREADS, (BN * TC_BK * 2 + BM * TC_BK * 2) * STAGES <= 50 * 1024 ? 4 :
                                              ((BN * TC_BK * 2 + BM * TC_BK * 2) * STAGES <= 100 * 1024 ? 2 : 1))
gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const __grid_constant__ CUtensorMap tma_a2, const __grid_constant__ CUtensorMap tma_b2, int kb_seg1,
               TC* __restrict__ C, int64_t ldc, const TC* R, int64_t ldr, const float* __restrict__ bias,
               int relu, float alpha, int M, int N, int K, int kb_per_split) {
  using S = TcSmem<BN, STAGES, BM>;
  constexpr int TC_STAGES = STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + S::BAR_OFF);
  uint64_t* empty = full + TC_STAGES;
  uint64_t* tmem_full = empty + TC_STAGES;
  uint32_t* tmem_slot = (uint32_t*)(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kb_total = (K + TC_BK - 1) / TC_BK;
  const int kb_begin = blockIdx.z * kb_per_split;
  const int kb_end = min(kb_total, kb_begin + kb_per_split);
  const int n_iter = kb_end - kb_begin;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // PDL: everything above (barrier init, TMEM allocation) overlapped the previous kernel's tail; operands, residual
  // and the output buffer may only be touched from here on.
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % TC_STAGES;
        const uint32_t ph = (it / TC_STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* sa = smem + s * S::STAGE;
        uint8_t* sb = sa + S::A_BYTES;
        mbar_expect_tx(&full[s], S::STAGE);
        // two-segment K: k-blocks [0, kb_seg1) come from (A, B), the rest from (A2, B2) -- C = A B + A2 B2 in one
        // accumulator (the two directions' input gradients of a BLSTM layer)
        const bool seg2 = kb_begin + it >= kb_seg1;
        const CUtensorMap* pa = seg2 ? &tma_a2 : &tma_a;
        const CUtensorMap* pb = seg2 ? &tma_b2 : &tma_b;
        const int k0 = (kb_begin + it - (seg2 ? kb_seg1 : 0)) * TC_BK;
        if (A_MN) {
#pragma unroll
          for (int c = 0; c < BM / 64; ++c) tma_load_2d(sa + c * (TC_BK * 128), pa, m0 + c * 64, k0, &full[s]);
        } else {
          tma_load_2d(sa, pa, k0, m0, &full[s]);
        }
        if (B_MN) {
#pragma unroll
          for (int c = 0; c < BN / 64; ++c) tma_load_2d(sb + c * (TC_BK * 128), pb, n0 + c * 64, k0, &full[s]);
        } else {
          tma_load_2d(sb, pb, k0, n0, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(BM, BN, A_MN, B_MN);
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % TC_STAGES;
        const uint32_t ph = (it / TC_STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * S::STAGE);
        const uint32_t sb = sa + S::A_BYTES;
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k) {
          // K-major: 16 k-elements = 32 B inside the 128 B swizzled row; SBO = 8 rows * 128 B.
          // MN-major: 16 k-rows = 2048 B; LBO = next 64-wide mn chunk (BK*128 B), SBO = 8 k-rows * 128 B.
          const uint64_t da = A_MN ? umma_desc(sa + k * 2048, TC_BK * 128, 1024) : umma_desc(sa + k * 32, 16, 1024);
          const uint64_t db = B_MN ? umma_desc(sb + k * 2048, TC_BK * 128, 1024) : umma_desc(sb + k * 32, 16, 1024);
          tc_mma_f16(tmem_base, da, db, idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        tc_commit(&empty[s]);          // frees the smem stage once these MMAs have read it
      }
      tc_commit(tmem_full);            // accumulator complete
    }
  } else if (warp >= 4) {
    // ---- epilogue: TMEM -> registers -> per-warp staging tile in the (now idle) pipeline buffers -> coalesced
    // global stores.  Each warp owns the 32 accumulator rows of its TMEM lane quarter; staging makes every store
    // instruction write whole contiguous row segments instead of 32 scattered 16-byte pieces.
    const int wq = warp - 4;           // == warp % 4: the TMEM lane quarter this warp may access
    constexpr int SP = BN + 4;         // staging row pitch (floats): float4-aligned, conflict-free
    float* stg = reinterpret_cast<float*>(smem) + wq * 32 * SP;
    mbar_wait(tmem_full, 0);           // all MMAs retired: accumulator complete, smem stages no longer read
    tc_fence_after();
    if (BM == 128 && !ATOMIC && R == nullptr && n_iter > 0) {
      // no residual operand: registers -> global directly (a thread owns 32 consecutive columns of its row; 256-bit
      // stores), skipping the staging round trip -- these tiles are latency-bound and the epilogue is their tail
      const int row = m0 + wq * 32 + lane;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)c0, r);
        const int colb = n0 + c0;
        if (row >= M || colb >= N) continue;
        float v[32];
#pragma unroll
        for (int q = 0; q < 32; ++q) v[q] = alpha * __uint_as_float(r[q]);
        const bool fullw = colb + 32 <= N;
        if (bias) {
          if (fullw && (reinterpret_cast<uintptr_t>(bias + colb) & 15) == 0) {
#pragma unroll
            for (int q = 0; q < 32; q += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(bias + colb + q);
              v[q] += b4.x; v[q + 1] += b4.y; v[q + 2] += b4.z; v[q + 3] += b4.w;
            }
          } else {
#pragma unroll
            for (int q = 0; q < 32; ++q) if (colb + q < N) v[q] += bias[colb + q];
          }
        }
        if (relu == 1) {
#pragma unroll
          for (int q = 0; q < 32; ++q) v[q] = fmaxf(v[q], 0.f);
        }
        TC* cp = C + (int64_t)row * ldc + colb;
        if (fullw && (reinterpret_cast<uintptr_t>(cp) & 31) == 0) {
          if constexpr (sizeof(TC) == 4) {
#pragma unroll
            for (int q = 0; q < 32; q += 8)
              st_global_v8(reinterpret_cast<float*>(cp) + q, __float_as_uint(v[q]), __float_as_uint(v[q + 1]),
                           __float_as_uint(v[q + 2]), __float_as_uint(v[q + 3]), __float_as_uint(v[q + 4]),
                           __float_as_uint(v[q + 5]), __float_as_uint(v[q + 6]), __float_as_uint(v[q + 7]));
          } else {
#pragma unroll
            for (int q = 0; q < 32; q += 16) {
              uint32_t w[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                __nv_bfloat162 b2 = __floats2bfloat162_rn(v[q + 2 * e], v[q + 2 * e + 1]);
                w[e] = *reinterpret_cast<uint32_t*>(&b2);
              }
              st_global_v8(cp + q, w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7]);
            }
          }
        } else {
#pragma unroll
          for (int q = 0; q < 32; ++q) if (colb + q < N) cp[q] = from_f<TC>(v[q]);
        }
      }
    } else {
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)c0, r);
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(&stg[lane * SP + c0 + j]) =
            make_float4(alpha * __uint_as_float(r[j]), alpha * __uint_as_float(r[j + 1]),
                        alpha * __uint_as_float(r[j + 2]), alpha * __uint_as_float(r[j + 3]));
    }
    __syncwarp();
    if (n_iter > 0) {
      constexpr int LPR = BN / 4;        // lanes per row (4 columns each)
      constexpr int RPP = 32 / LPR;      // rows per pass
      constexpr int RPW = BM == 64 ? 16 : 32;   // accumulator rows held by one warp's TMEM lane quarter
      const int cc = (lane % LPR) * 4;
      const int col = n0 + cc;
      // residual rows are fetched UP rows ahead of their use: a dependent load per row pass would leave the whole
      // epilogue waiting on one global-memory latency per row
      constexpr int UP = 4;
      const bool has_r = R != nullptr && !ATOMIC;
#pragma unroll 1
      for (int rr0 = lane / LPR; rr0 < RPW; rr0 += RPP * UP) {
        float radd[UP][4];
#pragma unroll
        for (int u = 0; u < UP; ++u) {
          radd[u][0] = radd[u][1] = radd[u][2] = radd[u][3] = 0.f;
          const int rr = rr0 + u * RPP;
          const int row = m0 + wq * RPW + rr;
          if (has_r && rr < RPW && row < M && col < N) {
            const TC* rp = R + (int64_t)row * ldr + col;
            const int nv = min(4, N - col);
            if constexpr (sizeof(TC) == 4) {
              if (nv == 4 && (reinterpret_cast<uintptr_t>(rp) & 15) == 0) {
                const float4 q = *reinterpret_cast<const float4*>(rp);
                radd[u][0] = q.x; radd[u][1] = q.y; radd[u][2] = q.z; radd[u][3] = q.w;
              } else {
                for (int j = 0; j < nv; ++j) radd[u][j] = to_f(rp[j]);
              }
            } else {
              if (nv == 4 && (reinterpret_cast<uintptr_t>(rp) & 7) == 0) {
                const uint2 q = *reinterpret_cast<const uint2*>(rp);
                radd[u][0] = __uint_as_float(q.x << 16); radd[u][1] = __uint_as_float(q.x & 0xffff0000u);
                radd[u][2] = __uint_as_float(q.y << 16); radd[u][3] = __uint_as_float(q.y & 0xffff0000u);
              } else {
                for (int j = 0; j < nv; ++j) radd[u][j] = to_f(rp[j]);
              }
            }
          }
        }
#pragma unroll
        for (int u = 0; u < UP; ++u) {
          const int rr = rr0 + u * RPP;
          const int row = m0 + wq * RPW + rr;
          if (rr >= RPW || row >= M || col >= N) continue;
          const float4 a = *reinterpret_cast<const float4*>(&stg[rr * SP + cc]);
          float v[4] = {a.x, a.y, a.z, a.w};
          const int nv = min(4, N - col);
          TC* cp = C + (int64_t)row * ldc + col;
          if (ATOMIC) {
            float* fp = reinterpret_cast<float*>(cp);
            if (nv == 4 && (reinterpret_cast<uintptr_t>(fp) & 15) == 0) {
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(fp), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
            } else {
              for (int j = 0; j < nv; ++j) atomicAdd(fp + j, v[j]);
            }
            continue;
          }
          if (bias) {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (j < nv) v[j] += bias[col + j];
          }
          if (relu == 1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          if (relu == 2) {            // R gates the result: (R > 0) ? v : 0  (ReLU backward fused into the dX GEMM)
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = radd[u][j] > 0.f ? v[j] : 0.f;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] += radd[u][j];
          }
          if constexpr (sizeof(TC) == 4) {
            if (nv == 4 && (reinterpret_cast<uintptr_t>(cp) & 15) == 0) {
              *reinterpret_cast<float4*>(cp) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
              for (int j = 0; j < nv; ++j) reinterpret_cast<float*>(cp)[j] = v[j];
            }
          } else {
            if (nv == 4 && (reinterpret_cast<uintptr_t>(cp) & 7) == 0) {
              uint2 o;
              __nv_bfloat162* oo = reinterpret_cast<__nv_bfloat162*>(&o);
              oo[0] = __floats2bfloat162_rn(v[0], v[1]);
              oo[1] = __floats2bfloat162_rn(v[2], v[3]);
              *reinterpret_cast<uint2*>(cp) = o;
            } else {
              for (int j = 0; j < nv; ++j) cp[j] = from_f<TC>(v[j]);
            }
          }
        }
      }
    }
    }   // staged epilogue
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
  }
}

// ---- persistent variant for the throughput shapes (hundreds of output tiles) ------------------------------------
// One CTA per SM walks a static list of 128 x BN output tiles (tile = blockIdx.x + j * gridDim.x, N fastest so the
// CTAs of one wave share A rows through L2).  What the one-tile-per-CTA kernel pays per tile -- barrier init, TMEM
// allocation, an empty pipeline at the start and an idle tensor pipe during the epilogue -- is paid once per CTA here:
//   * the TMA producer runs ahead across tile boundaries (the smem ring never drains),
//   * the accumulator is DOUBLE-BUFFERED in TMEM (2 x BN fp32 columns): the issuer starts tile j+1 into the other
//     buffer while the four epilogue warps drain tile j (tmem_full / tmem_empty mbarrier pairs),
//   * the epilogue stages 32-column chunks through a private 18 KB region (the pipeline buffers are never idle), so
//     every global store still writes whole row segments.
constexpr int TCP_THREADS = 384;   // persistent kernels: warps 0-2 producer / issuer / allocator, warps 4-11 epilogue

template <int BN, int STAGES>
struct TcPersistSmem {
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;
  static constexpr int B_BYTES = BN * TC_BK * 2;
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int STG_PITCH = 36;                                    // floats per staged row (32 + pad, 16 B aligned)
  static constexpr int STG_OFF = STAGES * STAGE;
  static constexpr int STG_BYTES = 4 * 32 * STG_PITCH * 4;
  static constexpr int BAR_OFF = STG_OFF + STG_BYTES;
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 4) * 8 + 16 + 1024;
};

template <int BN, int STAGES, bool A_MN, bool B_MN, typename TC>
__global__ void __launch_bounds__(TCP_THREADS, 1)
gemm_tc_persist_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                       const __grid_constant__ CUtensorMap tma_a2, const __grid_constant__ CUtensorMap tma_b2,
                       int kb_seg1, TC* __restrict__ C, int64_t ldc, const TC* R, int64_t ldr,
                       const float* __restrict__ bias, int relu, float alpha, int M, int N, int K, int n_tiles_n,
                       int n_tiles) {
  using S = TcPersistSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + S::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;         // [2] accumulator buffer complete
  uint64_t* tempty = tfull + 2;             // [2] accumulator buffer drained by the epilogue
  uint32_t* tmem_slot = (uint32_t*)(tempty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb_total = (K + TC_BK - 1) / TC_BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)(2 * BN)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;                       // running k-block counter across tiles: ring slot and phase
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles_n) * TC_BM, n0 = (tile % n_tiles_n) * BN;
        for (int kb = 0; kb < kb_total; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
          uint8_t* sa = smem + s * S::STAGE;
          uint8_t* sb = sa + S::A_BYTES;
          mbar_expect_tx(&full[s], S::STAGE);
          const bool seg2 = kb >= kb_seg1;
          const CUtensorMap* pa = seg2 ? &tma_a2 : &tma_a;
          const CUtensorMap* pb = seg2 ? &tma_b2 : &tma_b;
          const int k0 = (kb - (seg2 ? kb_seg1 : 0)) * TC_BK;
          if (A_MN) {
#pragma unroll
            for (int c = 0; c < TC_BM / 64; ++c) tma_load_2d(sa + c * (TC_BK * 128), pa, m0 + c * 64, k0, &full[s]);
          } else {
            tma_load_2d(sa, pa, k0, m0, &full[s]);
          }
          if (B_MN) {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c) tma_load_2d(sb + c * (TC_BK * 128), pb, n0 + c * 64, k0, &full[s]);
          } else {
            tma_load_2d(sb, pb, k0, n0, &full[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(TC_BM, BN, A_MN, B_MN);
      uint32_t it = 0, j = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++j) {
        const uint32_t buf = j & 1;
        mbar_wait(&tempty[buf], ((j >> 1) & 1) ^ 1);      // the epilogue has drained this buffer (free at first use)
        tc_fence_after();
        const uint32_t acc = tmem_base + buf * BN;
        for (int kb = 0; kb < kb_total; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&full[s], (it / STAGES) & 1);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * S::STAGE);
          const uint32_t sb = sa + S::A_BYTES;
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            const uint64_t da = A_MN ? umma_desc(sa + k * 2048, TC_BK * 128, 1024) : umma_desc(sa + k * 32, 16, 1024);
            const uint64_t db = B_MN ? umma_desc(sb + k * 2048, TC_BK * 128, 1024) : umma_desc(sb + k * 32, 16, 1024);
            tc_mma_f16(acc, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          tc_commit(&empty[s]);
        }
        tc_commit(&tfull[buf]);
      }
    }
  } else if (warp >= 4) {
    // eight epilogue warps: warp w may touch TMEM lanes [32 (w % 4), +32); warps 4-7 take the even 32-column chunks,
    // warps 8-11 the odd ones, so every scheduler has two warps whose TMEM-load / store latencies overlap.  The staged
    // path (residual operand) is driven by warps 4-7 only: the staging region holds four warps' rows.
    const int wq = warp & 3;
    const int chalf = (warp - 4) >> 2;
    constexpr int SP = S::STG_PITCH;
    float* stg = reinterpret_cast<float*>(smem + S::STG_OFF) + wq * 32 * SP;
    const int cc = (lane & 7) * 4;          // 8 lanes x 4 columns cover a 32-column chunk row
    const int rl = lane >> 3;               // 4 rows per pass
    uint32_t j = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++j) {
      const int m0 = (tile / n_tiles_n) * TC_BM, n0 = (tile % n_tiles_n) * BN;
      const uint32_t buf = j & 1;
      mbar_wait(&tfull[buf], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t acc = tmem_base + buf * BN + ((uint32_t)(wq * 32) << 16);
      const bool staged = R != nullptr;
      if (staged && chalf == 1) {            // nothing to read in staged mode: release immediately
        tc_fence_before();
        if (lane == 0) mbar_arrive(&tempty[buf]);
        continue;
      }
      const int cstep = staged ? 32 : 64;
#pragma unroll 1
      for (int c0 = staged ? 0 : chalf * 32; c0 < BN; c0 += cstep) {
        uint32_t r[32];
        tmem_ld32(acc + (uint32_t)c0, r);
        if (c0 + cstep >= BN) {              // last TMEM read of this tile by this warp: hand the buffer back
          tc_fence_before();
          if (lane == 0) mbar_arrive(&tempty[buf]);
        }
        if (R == nullptr) {
          // No residual operand: registers -> global directly.  A thread owns 32 consecutive columns of its row (64 B of
          // bf16 / 128 B of fp32): four / eight 16-byte stores, ~5x fewer instructions than the staged path below, which
          // matters because the epilogue of a K = 1024 tile has to fit under that tile's 8k-cycle mainloop.
          const int row = m0 + wq * 32 + lane;
          const int colb = n0 + c0;
          if (row < M && colb < N && !(false)) {
            float v[32];
#pragma unroll
            for (int q = 0; q < 32; ++q) v[q] = alpha * __uint_as_float(r[q]);
            const bool fullw = colb + 32 <= N;
            if (bias) {
              if (fullw && (reinterpret_cast<uintptr_t>(bias + colb) & 15) == 0) {
#pragma unroll
                for (int q = 0; q < 32; q += 4) {
                  const float4 b4 = *reinterpret_cast<const float4*>(bias + colb + q);
                  v[q] += b4.x; v[q + 1] += b4.y; v[q + 2] += b4.z; v[q + 3] += b4.w;
                }
              } else {
#pragma unroll
                for (int q = 0; q < 32; ++q) if (colb + q < N) v[q] += bias[colb + q];
              }
            }
            if (relu == 1) {
#pragma unroll
              for (int q = 0; q < 32; ++q) v[q] = fmaxf(v[q], 0.f);
            }
            TC* cp = C + (int64_t)row * ldc + colb;
            if (fullw && (reinterpret_cast<uintptr_t>(cp) & 31) == 0) {
              // 256-bit stores (sm_100): every instruction writes whole 32-byte sectors
              if constexpr (sizeof(TC) == 4) {
#pragma unroll
                for (int q = 0; q < 32; q += 8)
                  st_global_v8(reinterpret_cast<float*>(cp) + q, __float_as_uint(v[q]), __float_as_uint(v[q + 1]),
                               __float_as_uint(v[q + 2]), __float_as_uint(v[q + 3]), __float_as_uint(v[q + 4]),
                               __float_as_uint(v[q + 5]), __float_as_uint(v[q + 6]), __float_as_uint(v[q + 7]));
              } else {
#pragma unroll
                for (int q = 0; q < 32; q += 16) {
                  uint32_t w[8];
#pragma unroll
                  for (int e = 0; e < 8; ++e) {
                    __nv_bfloat162 b2 = __floats2bfloat162_rn(v[q + 2 * e], v[q + 2 * e + 1]);
                    w[e] = *reinterpret_cast<uint32_t*>(&b2);
                  }
                  st_global_v8(cp + q, w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7]);
                }
              }
            } else if (fullw && (reinterpret_cast<uintptr_t>(cp) & 15) == 0) {
              if constexpr (sizeof(TC) == 4) {
#pragma unroll
                for (int q = 0; q < 32; q += 4)
                  *reinterpret_cast<float4*>(reinterpret_cast<float*>(cp) + q) = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
              } else {
#pragma unroll
                for (int q = 0; q < 32; q += 8) {
                  uint4 o4;
                  __nv_bfloat162* oo = reinterpret_cast<__nv_bfloat162*>(&o4);
                  oo[0] = __floats2bfloat162_rn(v[q], v[q + 1]);
                  oo[1] = __floats2bfloat162_rn(v[q + 2], v[q + 3]);
                  oo[2] = __floats2bfloat162_rn(v[q + 4], v[q + 5]);
                  oo[3] = __floats2bfloat162_rn(v[q + 6], v[q + 7]);
                  *reinterpret_cast<uint4*>(cp + q) = o4;
                }
              }
            } else {
#pragma unroll
              for (int q = 0; q < 32; ++q) if (colb + q < N) cp[q] = from_f<TC>(v[q]);
            }
          }
          continue;
        }
        __syncwarp();                        // the previous chunk's readers are done with the staging rows
#pragma unroll
        for (int q = 0; q < 32; q += 4)
          *reinterpret_cast<float4*>(&stg[lane * SP + q]) =
              make_float4(alpha * __uint_as_float(r[q]), alpha * __uint_as_float(r[q + 1]),
                          alpha * __uint_as_float(r[q + 2]), alpha * __uint_as_float(r[q + 3]));
        __syncwarp();
        const int col = n0 + c0 + cc;
        if (col >= N) continue;
        const int nv = min(4, N - col);
        // residual rows of the 8 passes are fetched 4 passes ahead of their use
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float radd[4][4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            radd[u][0] = radd[u][1] = radd[u][2] = radd[u][3] = 0.f;
            const int row = m0 + wq * 32 + (half * 4 + u) * 4 + rl;
            if (R != nullptr && row < M) {
              const TC* rp = R + (int64_t)row * ldr + col;
              if constexpr (sizeof(TC) == 4) {
                if (nv == 4 && (reinterpret_cast<uintptr_t>(rp) & 15) == 0) {
                  const float4 v4 = *reinterpret_cast<const float4*>(rp);
                  radd[u][0] = v4.x; radd[u][1] = v4.y; radd[u][2] = v4.z; radd[u][3] = v4.w;
                } else {
                  for (int e = 0; e < nv; ++e) radd[u][e] = to_f(rp[e]);
                }
              } else {
                if (nv == 4 && (reinterpret_cast<uintptr_t>(rp) & 7) == 0) {
                  const uint2 v2 = *reinterpret_cast<const uint2*>(rp);
                  radd[u][0] = __uint_as_float(v2.x << 16); radd[u][1] = __uint_as_float(v2.x & 0xffff0000u);
                  radd[u][2] = __uint_as_float(v2.y << 16); radd[u][3] = __uint_as_float(v2.y & 0xffff0000u);
                } else {
                  for (int e = 0; e < nv; ++e) radd[u][e] = to_f(rp[e]);
                }
              }
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int rr = (half * 4 + u) * 4 + rl;
            const int row = m0 + wq * 32 + rr;
            if (row >= M) continue;
            const float4 a4 = *reinterpret_cast<const float4*>(&stg[rr * SP + cc]);
            float v[4] = {a4.x, a4.y, a4.z, a4.w};
            if (bias) {
#pragma unroll
              for (int e = 0; e < 4; ++e) if (e < nv) v[e] += bias[col + e];
            }
            if (relu == 1) {
#pragma unroll
              for (int e = 0; e < 4; ++e) v[e] = fmaxf(v[e], 0.f);
            }
            if (relu == 2) {
#pragma unroll
              for (int e = 0; e < 4; ++e) v[e] = radd[u][e] > 0.f ? v[e] : 0.f;
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) v[e] += radd[u][e];
            }
            TC* cp = C + (int64_t)row * ldc + col;
            if constexpr (sizeof(TC) == 4) {
              if (nv == 4 && (reinterpret_cast<uintptr_t>(cp) & 15) == 0) {
                *reinterpret_cast<float4*>(cp) = make_float4(v[0], v[1], v[2], v[3]);
              } else {
                for (int e = 0; e < nv; ++e) reinterpret_cast<float*>(cp)[e] = v[e];
              }
            } else {
              if (nv == 4 && (reinterpret_cast<uintptr_t>(cp) & 7) == 0) {
                uint2 o2;
                __nv_bfloat162* oo = reinterpret_cast<__nv_bfloat162*>(&o2);
                oo[0] = __floats2bfloat162_rn(v[0], v[1]);
                oo[1] = __floats2bfloat162_rn(v[2], v[3]);
                *reinterpret_cast<uint2*>(cp) = o2;
              } else {
                for (int e = 0; e < nv; ++e) cp[e] = from_f<TC>(v[e]);
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * BN)) : "memory");
  }
}

// ---- CTA-pair variant (cta_group::2): the two CTAs of a cluster sit on the two SMs of one TPC and execute ONE
// tcgen05.mma of M = 256, N = 256 per K step.  Each CTA stages only its own 128 A rows and HALF of the B tile
// (128 of the 256 N rows), so shared-memory fill + operand-read traffic per FLOP is half that of the single-CTA kernel
// -- a single SM streaming both operands of a 128-row MMA out of its own shared memory tops out near 50 % of the
// tensor peak.  Pipeline = the persistent kernel's (smem ring across tiles, double-buffered TMEM accumulator), with the
// cross-CTA plumbing: both CTAs' TMA bytes are credited to the LEADER's full barrier, the leader issues every MMA and
// its tcgen05.commit arrives (multicast) on the empty / tmem_full barriers of BOTH CTAs, and both CTAs' epilogue warps
// release an accumulator buffer by arriving on the leader's tmem_empty barrier.
template <int STAGES, bool A_MN, bool B_MN, typename TC>
__global__ void __launch_bounds__(TCP_THREADS, 1)
gemm_tc_pair_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                       const __grid_constant__ CUtensorMap tma_a2, const __grid_constant__ CUtensorMap tma_b2,
                       int kb_seg1, TC* __restrict__ C, int64_t ldc, const TC* R, int64_t ldr,
                       const float* __restrict__ bias, int relu, float alpha, int M, int N, int K, int n_tiles_n,
                       int n_tiles, int dbg, int splits, int kb_per_split) {
  constexpr int BN = 256;                   // accumulator columns per CTA = N extent of the pair's tile
  constexpr int BNH = 128;                  // B rows (N) staged by EACH CTA; the MMA reads both halves
  using S = TcPersistSmem<BNH, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + S::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;         // [2] accumulator buffer complete
  uint64_t* tempty = tfull + 2;             // [2] accumulator buffer drained by the epilogue
  uint32_t* tmem_slot = (uint32_t*)(tempty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb_total = (K + TC_BK - 1) / TC_BK;
  const uint32_t rank = cluster_ctarank();  // 0 = leader (issues the MMAs), 1 = peer
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 16); }  // 8 epilogue warps x 2 CTAs
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {                           // the same warp of BOTH CTAs performs the pair allocation
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)(2 * BN)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                        // both CTAs' barriers exist before any cross-CTA arrival
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0 && !(dbg & 1)) {
      uint32_t it = 0;                       // running k-block counter across tiles: ring slot and phase
      for (int unit = pair; unit < n_tiles * splits; unit += n_pairs) {
        // work unit = (256 x 256 tile of the pair, K split); this CTA stages A rows [m0, m0+128) and B rows (N) [n0, n0+128)
        const int tile = unit / splits, kb0 = (unit % splits) * kb_per_split, kb1 = min(kb_total, kb0 + kb_per_split);
        const int m0 = (tile / n_tiles_n) * 256 + (int)rank * TC_BM, n0 = (tile % n_tiles_n) * BN + (int)rank * BNH;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
          uint8_t* sa = smem + s * S::STAGE;
          uint8_t* sb = sa + S::A_BYTES;
          if (rank == 0) mbar_expect_tx(&full[s], 2 * S::STAGE);      // the leader's barrier counts both CTAs' bytes
          const bool seg2 = kb >= kb_seg1;
          const CUtensorMap* pa = seg2 ? &tma_a2 : &tma_a;
          const CUtensorMap* pb = seg2 ? &tma_b2 : &tma_b;
          const int k0 = (kb - (seg2 ? kb_seg1 : 0)) * TC_BK;
          if (A_MN) {
#pragma unroll
            for (int c = 0; c < TC_BM / 64; ++c) tma_load_2d_pair(sa + c * (TC_BK * 128), pa, m0 + c * 64, k0, &full[s]);
          } else {
            tma_load_2d_pair(sa, pa, k0, m0, &full[s]);
          }
          if (B_MN) {
#pragma unroll
            for (int c = 0; c < BNH / 64; ++c) tma_load_2d_pair(sb + c * (TC_BK * 128), pb, n0 + c * 64, k0, &full[s]);
          } else {
            tma_load_2d_pair(sb, pb, k0, n0, &full[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc(256, BN, A_MN, B_MN);     // M = 256 across the pair
      uint32_t it = 0, j = 0;
      for (int unit = pair; unit < n_tiles * splits; unit += n_pairs, ++j) {
        const int kb0 = (unit % splits) * kb_per_split, kb1 = min(kb_total, kb0 + kb_per_split);
        const uint32_t buf = j & 1;
        mbar_wait(&tempty[buf], ((j >> 1) & 1) ^ 1);      // the epilogue has drained this buffer (free at first use)
        tc_fence_after();
        const uint32_t acc = tmem_base + buf * BN;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % STAGES;
          if (!(dbg & 1)) mbar_wait(&full[s], (it / STAGES) & 1);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * S::STAGE);
          const uint32_t sb = sa + S::A_BYTES;
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            const uint64_t da = A_MN ? umma_desc(sa + k * 2048, TC_BK * 128, 1024) : umma_desc(sa + k * 32, 16, 1024);
            const uint64_t db = B_MN ? umma_desc(sb + k * 2048, TC_BK * 128, 1024) : umma_desc(sb + k * 32, 16, 1024);
            tc_mma_f16_pair(acc, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          tc_commit_pair(&empty[s]);
        }
        tc_commit_pair(&tfull[buf]);
      }
    }
  } else if (warp >= 4) {
    // eight epilogue warps: warp w may touch TMEM lanes [32 (w % 4), +32); warps 4-7 take the even 32-column chunks,
    // warps 8-11 the odd ones, so every scheduler has two warps whose TMEM-load / store latencies overlap.  The staged
    // path (residual operand) is driven by warps 4-7 only: the staging region holds four warps' rows.
    const int wq = warp & 3;
    const int chalf = (warp - 4) >> 2;
    constexpr int SP = S::STG_PITCH;
    float* stg = reinterpret_cast<float*>(smem + S::STG_OFF) + wq * 32 * SP;
    const int cc = (lane & 7) * 4;          // 8 lanes x 4 columns cover a 32-column chunk row
    const int rl = lane >> 3;               // 4 rows per pass
    uint32_t j = 0;
    for (int unit = pair; unit < n_tiles * splits; unit += n_pairs, ++j) {
      const int tile = unit / splits;
      const int m0 = (tile / n_tiles_n) * 256 + (int)rank * TC_BM, n0 = (tile % n_tiles_n) * BN;
      const uint32_t buf = j & 1;
      mbar_wait(&tfull[buf], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t acc = tmem_base + buf * BN + ((uint32_t)(wq * 32) << 16);
      const bool staged = R != nullptr;
      if (staged && chalf == 1) {            // nothing to read in staged mode: release immediately
        tc_fence_before();
        if (lane == 0) mbar_arrive_leader(&tempty[buf]);
        continue;
      }
      const int cstep = staged ? 32 : 64;
#pragma unroll 1
      for (int c0 = staged ? 0 : chalf * 32; c0 < BN; c0 += cstep) {
        uint32_t r[32];
        tmem_ld32(acc + (uint32_t)c0, r);
        if (c0 + cstep >= BN) {              // last TMEM read of this tile by this warp: hand the buffer back
          tc_fence_before();
          if (lane == 0) mbar_arrive_leader(&tempty[buf]);
        }
        if (R == nullptr) {
          // No residual operand: registers -> global directly.  A thread owns 32 consecutive columns of its row (64 B of
          // bf16 / 128 B of fp32): four / eight 16-byte stores, ~5x fewer instructions than the staged path below, which
          // matters because the epilogue of a K = 1024 tile has to fit under that tile's 8k-cycle mainloop.
          const int row = m0 + wq * 32 + lane;
          const int colb = n0 + c0;
          if (row < M && colb < N && !(dbg & 2)) {
            float v[32];
#pragma unroll
            for (int q = 0; q < 32; ++q) v[q] = alpha * __uint_as_float(r[q]);
            const bool fullw = colb + 32 <= N;
            if (bias) {
              if (fullw && (reinterpret_cast<uintptr_t>(bias + colb) & 15) == 0) {
#pragma unroll
                for (int q = 0; q < 32; q += 4) {
                  const float4 b4 = *reinterpret_cast<const float4*>(bias + colb + q);
                  v[q] += b4.x; v[q + 1] += b4.y; v[q + 2] += b4.z; v[q + 3] += b4.w;
                }
              } else {
#pragma unroll
                for (int q = 0; q < 32; ++q) if (colb + q < N) v[q] += bias[colb + q];
              }
            }
            if (relu == 1) {
#pragma unroll
              for (int q = 0; q < 32; ++q) v[q] = fmaxf(v[q], 0.f);
            }
            TC* cp = C + (int64_t)row * ldc + colb;
            if (splits > 1) {                  // split-K partial: fp32 reduction into the zero-filled output
              if constexpr (sizeof(TC) == 4) {
                float* fp = reinterpret_cast<float*>(cp);
                if (fullw && (reinterpret_cast<uintptr_t>(fp) & 15) == 0) {
#pragma unroll
                  for (int q = 0; q < 32; q += 4)
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(fp + q), "f"(v[q]), "f"(v[q + 1]), "f"(v[q + 2]), "f"(v[q + 3]) : "memory");
                } else {
#pragma unroll
                  for (int q = 0; q < 32; ++q) if (colb + q < N) atomicAdd(fp + q, v[q]);
                }
              }
              continue;
            }
            if (fullw && (reinterpret_cast<uintptr_t>(cp) & 31) == 0) {
              // 256-bit stores (sm_100): every instruction writes whole 32-byte sectors
              if constexpr (sizeof(TC) == 4) {
#pragma unroll
                for (int q = 0; q < 32; q += 8)
                  st_global_v8(reinterpret_cast<float*>(cp) + q, __float_as_uint(v[q]), __float_as_uint(v[q + 1]),
                               __float_as_uint(v[q + 2]), __float_as_uint(v[q + 3]), __float_as_uint(v[q + 4]),
                               __float_as_uint(v[q + 5]), __float_as_uint(v[q + 6]), __float_as_uint(v[q + 7]));
              } else {
#pragma unroll
                for (int q = 0; q < 32; q += 16) {
                  uint32_t w[8];
#pragma unroll
                  for (int e = 0; e < 8; ++e) {
                    __nv_bfloat162 b2 = __floats2bfloat162_rn(v[q + 2 * e], v[q + 2 * e + 1]);
                    w[e] = *reinterpret_cast<uint32_t*>(&b2);
                  }
                  st_global_v8(cp + q, w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7]);
                }
              }
            } else if (fullw && (reinterpret_cast<uintptr_t>(cp) & 15) == 0) {
              if constexpr (sizeof(TC) == 4) {
#pragma unroll
                for (int q = 0; q < 32; q += 4)
                  *reinterpret_cast<float4*>(reinterpret_cast<float*>(cp) + q) = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
              } else {
#pragma unroll
                for (int q = 0; q < 32; q += 8) {
                  uint4 o4;
                  __nv_bfloat162* oo = reinterpret_cast<__nv_bfloat162*>(&o4);
                  oo[0] = __floats2bfloat162_rn(v[q], v[q + 1]);
                  oo[1] = __floats2bfloat162_rn(v[q + 2], v[q + 3]);
                  oo[2] = __floats2bfloat162_rn(v[q + 4], v[q + 5]);
                  oo[3] = __floats2bfloat162_rn(v[q + 6], v[q + 7]);
                  *reinterpret_cast<uint4*>(cp + q) = o4;
                }
              }
            } else {
#pragma unroll
              for (int q = 0; q < 32; ++q) if (colb + q < N) cp[q] = from_f<TC>(v[q]);
            }
          }
          continue;
        }
        __syncwarp();                        // the previous chunk's readers are done with the staging rows
#pragma unroll
        for (int q = 0; q < 32; q += 4)
          *reinterpret_cast<float4*>(&stg[lane * SP + q]) =
              make_float4(alpha * __uint_as_float(r[q]), alpha * __uint_as_float(r[q + 1]),
                          alpha * __uint_as_float(r[q + 2]), alpha * __uint_as_float(r[q + 3]));
        __syncwarp();
        const int col = n0 + c0 + cc;
        if (col >= N || (dbg & 2)) continue;
        const int nv = min(4, N - col);
        // residual rows of the 8 passes are fetched 4 passes ahead of their use
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float radd[4][4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            radd[u][0] = radd[u][1] = radd[u][2] = radd[u][3] = 0.f;
            const int row = m0 + wq * 32 + (half * 4 + u) * 4 + rl;
            if (R != nullptr && row < M) {
              const TC* rp = R + (int64_t)row * ldr + col;
              if constexpr (sizeof(TC) == 4) {
                if (nv == 4 && (reinterpret_cast<uintptr_t>(rp) & 15) == 0) {
                  const float4 v4 = *reinterpret_cast<const float4*>(rp);
                  radd[u][0] = v4.x; radd[u][1] = v4.y; radd[u][2] = v4.z; radd[u][3] = v4.w;
                } else {
                  for (int e = 0; e < nv; ++e) radd[u][e] = to_f(rp[e]);
                }
              } else {
                if (nv == 4 && (reinterpret_cast<uintptr_t>(rp) & 7) == 0) {
                  const uint2 v2 = *reinterpret_cast<const uint2*>(rp);
                  radd[u][0] = __uint_as_float(v2.x << 16); radd[u][1] = __uint_as_float(v2.x & 0xffff0000u);
                  radd[u][2] = __uint_as_float(v2.y << 16); radd[u][3] = __uint_as_float(v2.y & 0xffff0000u);
                } else {
                  for (int e = 0; e < nv; ++e) radd[u][e] = to_f(rp[e]);
                }
              }
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int rr = (half * 4 + u) * 4 + rl;
            const int row = m0 + wq * 32 + rr;
            if (row >= M) continue;
            const float4 a4 = *reinterpret_cast<const float4*>(&stg[rr * SP + cc]);
            float v[4] = {a4.x, a4.y, a4.z, a4.w};
            if (bias) {
#pragma unroll
              for (int e = 0; e < 4; ++e) if (e < nv) v[e] += bias[col + e];
            }
            if (relu == 1) {
#pragma unroll
              for (int e = 0; e < 4; ++e) v[e] = fmaxf(v[e], 0.f);
            }
            if (relu == 2) {
#pragma unroll
              for (int e = 0; e < 4; ++e) v[e] = radd[u][e] > 0.f ? v[e] : 0.f;
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) v[e] += radd[u][e];
            }
            TC* cp = C + (int64_t)row * ldc + col;
            if constexpr (sizeof(TC) == 4) {
              if (nv == 4 && (reinterpret_cast<uintptr_t>(cp) & 15) == 0) {
                *reinterpret_cast<float4*>(cp) = make_float4(v[0], v[1], v[2], v[3]);
              } else {
                for (int e = 0; e < nv; ++e) reinterpret_cast<float*>(cp)[e] = v[e];
              }
            } else {
              if (nv == 4 && (reinterpret_cast<uintptr_t>(cp) & 7) == 0) {
                uint2 o2;
                __nv_bfloat162* oo = reinterpret_cast<__nv_bfloat162*>(&o2);
                oo[0] = __floats2bfloat162_rn(v[0], v[1]);
                oo[1] = __floats2bfloat162_rn(v[2], v[3]);
                *reinterpret_cast<uint2*>(cp) = o2;
              } else {
                for (int e = 0; e < nv; ++e) cp[e] = from_f<TC>(v[e]);
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                        // neither CTA's shared memory / TMEM goes away while the other still uses it
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * BN)) : "memory");
  }
}

// ---- cluster split-K for the decoder-step shapes (M <= 64, long K, few N tiles) ---------------------------------
// A [64 x N] output with K = 2048 and N = 512 has 8 output tiles: 8 CTAs would each stream 256 KB of weights alone
// (~12 us, latency-bound).  Here the K range of every tile is split over a cluster of CS CTAs along grid.z (CS x more
// CTAs, each a 2-4 k-block pipeline); the fp32 partial tiles are reduce-scattered through distributed shared memory
// -- rank r receives rows [r * 64/CS, (r+1) * 64/CS) of every partial -- and each rank finishes (alpha, bias, ReLU,
// residual) and stores its rows.  No zero-filled fp32 output, no atomics, no second kernel.
template <int BN, int STAGES, bool A_MN, bool B_MN, typename TC>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_clk_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                   TC* __restrict__ C, int64_t ldc, const TC* R, int64_t ldr, const float* __restrict__ bias,
                   int relu, float alpha, int M, int N, int K, int kb_per_split) {
  constexpr int BM = 64;
  using S = TcSmem<BN, STAGES, BM>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + S::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = (uint32_t*)(tmem_full + 1);
  float* red = reinterpret_cast<float*>(smem + S::BAR_OFF + 256);   // [CS][64 / CS][BN] partial rows this rank owns

  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");   // "running": waited for before any DSMEM store
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const uint32_t cs = gridDim.z, rank = blockIdx.z;             // the cluster spans grid.z exactly
  const int kb_total = (K + TC_BK - 1) / TC_BK;
  const int kb_begin = blockIdx.z * kb_per_split;
  const int kb_end = min(kb_total, kb_begin + kb_per_split);
  const int n_iter = max(0, kb_end - kb_begin);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch_dependents();
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* sa = smem + s * S::STAGE;
        uint8_t* sb = sa + S::A_BYTES;
        mbar_expect_tx(&full[s], S::STAGE);
        const int k0 = (kb_begin + it) * TC_BK;
        if (A_MN) tma_load_2d(sa, &tma_a, 0, k0, &full[s]);
        else tma_load_2d(sa, &tma_a, k0, 0, &full[s]);
        if (B_MN) {
#pragma unroll
          for (int c = 0; c < BN / 64; ++c) tma_load_2d(sb + c * (TC_BK * 128), &tma_b, n0 + c * 64, k0, &full[s]);
        } else {
          tma_load_2d(sb, &tma_b, k0, n0, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(BM, BN, A_MN, B_MN);
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * S::STAGE);
        const uint32_t sb = sa + S::A_BYTES;
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k) {
          const uint64_t da = A_MN ? umma_desc(sa + k * 2048, TC_BK * 128, 1024) : umma_desc(sa + k * 32, 16, 1024);
          const uint64_t db = B_MN ? umma_desc(sb + k * 2048, TC_BK * 128, 1024) : umma_desc(sb + k * 32, 16, 1024);
          tc_mma_f16(tmem_base, da, db, idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        tc_commit(&empty[s]);
      }
      tc_commit(tmem_full);
    }
  } else if (warp >= 4) {
    // partial tile -> owners: accumulator row i sits in TMEM lane (i / 16) * 32 + (i % 16); lanes 0..15 of warp wq
    // hold rows 16 wq .. 16 wq + 15 and send them, 16 bytes per store, to the rank that owns the row
    const int wq = warp - 4;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const uint32_t rpo = BM / cs;                     // rows per owner
    const uint32_t row = wq * 16 + (lane & 15);
    const uint32_t owner = row / rpo, lrow = row % rpo;
    uint32_t dst = smem_u32(red + ((size_t)rank * rpo + lrow) * BN), rdst;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rdst) : "r"(dst), "r"(owner));
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)c0, r);
      if (lane < 16) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const uint32_t z = n_iter > 0 ? 0xffffffffu : 0u;     // an empty K slice contributes zeros
          asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rdst + (c0 + j) * 4), "r"(r[j] & z),
                       "r"(r[j + 1] & z), "r"(r[j + 2] & z), "r"(r[j + 3] & z) : "memory");
        }
      }
    }
  }
  tc_fence_before();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  // ---- every rank: sum the CS partials of its rows, epilogue, store
  {
    const int rpo = BM / (int)cs;
    constexpr int QPR = BN / 4;                       // float4 quads per row
    for (int item = threadIdx.x; item < rpo * QPR; item += TC_THREADS) {
      const int lrow = item / QPR, cc = (item % QPR) * 4;
      const int row = (int)rank * rpo + lrow, col = n0 + cc;
      if (row >= M || col >= N) continue;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      for (int r = 0; r < (int)cs; ++r) {
        const float4 a = *reinterpret_cast<const float4*>(&red[((size_t)r * rpo + lrow) * BN + cc]);
        v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w;
      }
      const int nv = min(4, N - col);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[j] *= alpha;
        if (bias && j < nv) v[j] += bias[col + j];
        if (relu == 1) v[j] = fmaxf(v[j], 0.f);
      }
      TC* cp = C + (int64_t)row * ldc + col;
      const TC* rp = R ? R + (int64_t)row * ldr + col : nullptr;
      if (relu == 2 && rp) {
        for (int j = 0; j < nv; ++j) cp[j] = from_f<TC>(to_f(rp[j]) > 0.f ? v[j] : 0.f);
        continue;
      }
      for (int j = 0; j < nv; ++j) cp[j] = from_f<TC>(v[j] + (rp ? to_f(rp[j]) : 0.f));
    }
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
  }
}

// ---- host side -----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// Tensor map over a row-major bf16 matrix [rows, cols] (cols contiguous, leading dim ld), box = 64 cols x box_rows.
int make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return set_error("gemm_tc: cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error("gemm_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

bool gemm_tc_eligible(int dtype_ab, int ta, int tb, int64_t M, int64_t N, int64_t K, const void* A, int64_t lda,
                      const void* B, int64_t ldb, int64_t batch) {
  if (dtype_ab != B200ST_BF16 || batch != 1) return false;
  if (M < 1 || N < 8 || K < 8) return false;
  if (lda % 8 || ldb % 8) return false;                       // TMA: 16-byte global strides
  if (((uintptr_t)A & 15) || ((uintptr_t)B & 15)) return false;
  if (M >= (1ll << 31) || N >= (1ll << 31) || K >= (1ll << 31)) return false;
  return get_encode() != nullptr;
}

template <int BN, int STAGES, bool A_MN, bool B_MN, typename TC, int BM = TC_BM>
static int launch_tc(int64_t M, int64_t N, int64_t K, float alpha, const CUtensorMap& ma, const CUtensorMap& mb,
                     const CUtensorMap& ma2, const CUtensorMap& mb2, int kb_seg1, void* C, int64_t ldc, const void* R, int64_t ldr, const float* bias, int relu, int splits,
                     int kb_per_split, cudaStream_t st) {
  using S = TcSmem<BN, STAGES, BM>;
  static_assert(S::TOTAL <= 227 * 1024, "tile configuration exceeds shared memory");
  dim3 grid((unsigned)ceil_div(N, BN), (unsigned)ceil_div(M, BM), (unsigned)splits);
  if (splits > 1) {
    if constexpr (sizeof(TC) == 4) {
      auto kern = gemm_tc_kernel<BN, STAGES, A_MN, B_MN, float, true, BM>;
      B200ST_CUDA(cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
      B200ST_CUDA(launch_pdl(kern, grid, dim3(TC_THREADS), S::TOTAL, st, ma, mb, ma2, mb2, kb_seg1, (float*)C, ldc, (const float*)nullptr,
                             (int64_t)0, (const float*)nullptr, 0, alpha, (int)M, (int)N, (int)K, kb_per_split));
    } else {
      return set_error("gemm_tc: split-K needs an fp32 output");
    }
  } else {
    auto kern = gemm_tc_kernel<BN, STAGES, A_MN, B_MN, TC, false, BM>;
    B200ST_CUDA(cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    B200ST_CUDA(launch_pdl(kern, grid, dim3(TC_THREADS), S::TOTAL, st, ma, mb, ma2, mb2, kb_seg1, (TC*)C, ldc, (const TC*)R, ldr, bias, relu,
                           alpha, (int)M, (int)N, (int)K, kb_per_split));
  }
  B200ST_LAUNCH_CHECK("gemm_tc");
  return 0;
}

template <int BN, int STAGES, bool A_MN, bool B_MN, typename TC>
static int launch_tc_clk(int64_t M, int64_t N, int64_t K, float alpha, const CUtensorMap& ma, const CUtensorMap& mb,
                         void* C, int64_t ldc, const void* R, int64_t ldr, const float* bias, int relu, int cs,
                         int kb_per_split, cudaStream_t st) {
  using S = TcSmem<BN, STAGES, 64>;
  static_assert((2 * STAGES + 1) * 8 + 16 <= 256, "barrier block");
  constexpr int SMEM = S::BAR_OFF + 256 + 64 * BN * 4 + 1024;
  auto kern = gemm_tc_clk_kernel<BN, STAGES, A_MN, B_MN, TC>;
  B200ST_CUDA(cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  dim3 grid((unsigned)ceil_div(N, BN), 1, (unsigned)cs);
  B200ST_CUDA(launch_pdl_cluster_z(kern, grid, dim3(TC_THREADS), SMEM, st, (unsigned)cs, ma, mb, (TC*)C, ldc, (const TC*)R,
                                   ldr, bias, relu, alpha, (int)M, (int)N, (int)K, kb_per_split));
  B200ST_LAUNCH_CHECK("gemm_tc_clk");
  return 0;
}

static int g_persist_enabled = 1;
int gemm_tc_set_persistent(int on) { const int old = g_persist_enabled; if (on == 0 || on == 1) g_persist_enabled = on; return old; }
static int g_sm_count = 0;
// SM budget of the persistent kernels (0 = all SMs): deferred weight-gradient GEMMs that run on side streams UNDER a
// latency-bound recurrence are launched with a budget that leaves the recurrence's SMs free -- a persistent CTA holds its
// SM for the whole GEMM, and thread-block clusters of the recurrence that find no free SMs wait for it to end.
static int g_sm_budget = 0;
int gemm_tc_set_sm_budget(int n) { const int old = g_sm_budget; g_sm_budget = n < 0 ? 0 : n; return old; }
static inline int sm_budget() { return g_sm_budget > 0 && g_sm_budget < g_sm_count ? g_sm_budget : g_sm_count; }

template <int BN, int STAGES, bool A_MN, bool B_MN, typename TC>
static int launch_tc_persist(int64_t M, int64_t N, int64_t K, float alpha, const CUtensorMap& ma, const CUtensorMap& mb,
                             const CUtensorMap& ma2, const CUtensorMap& mb2, int kb_seg1, void* C, int64_t ldc,
                             const void* R, int64_t ldr, const float* bias, int relu, cudaStream_t st) {
  using S = TcPersistSmem<BN, STAGES>;
  static_assert(S::TOTAL <= 227 * 1024, "persistent tile configuration exceeds shared memory");
  if (g_sm_count == 0) {
    int dev = 0;
    B200ST_CUDA(cudaGetDevice(&dev));
    B200ST_CUDA(cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  const int n_tiles_n = (int)ceil_div(N, BN);
  const int n_tiles = (int)(ceil_div(M, TC_BM) * n_tiles_n);
  auto kern = gemm_tc_persist_kernel<BN, STAGES, A_MN, B_MN, TC>;
  B200ST_CUDA(cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
  dim3 grid((unsigned)min(n_tiles, sm_budget()));
  B200ST_CUDA(launch_pdl(kern, grid, dim3(TCP_THREADS), S::TOTAL, st, ma, mb, ma2, mb2, kb_seg1, (TC*)C, ldc, (const TC*)R,
                         ldr, bias, relu, alpha, (int)M, (int)N, (int)K, n_tiles_n, n_tiles));
  B200ST_LAUNCH_CHECK("gemm_tc_persist");
  return 0;
}

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl_pair(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                          Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = 2;
  attr[1].val.clusterDim.y = 1;
  attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static int g_pair_enabled = 1;
static int g_pair_dbg = 0;        // probe only: bit 0 = no TMA / no full-barrier wait, bit 1 = no epilogue stores
int gemm_tc_set_pair_dbg(int v) { const int old = g_pair_dbg; g_pair_dbg = v; return old; }
int gemm_tc_set_pair(int on) { const int old = g_pair_enabled; if (on == 0 || on == 1) g_pair_enabled = on; return old; }

template <int STAGES, bool A_MN, bool B_MN, typename TC>
static int launch_tc_pair(int64_t M, int64_t N, int64_t K, float alpha, const CUtensorMap& ma, const CUtensorMap& mb,
                          const CUtensorMap& ma2, const CUtensorMap& mb2, int kb_seg1, void* C, int64_t ldc,
                          const void* R, int64_t ldr, const float* bias, int relu, cudaStream_t st, int splits = 1,
                          int kb_per_split = 0x3fffffff) {
  using S = TcPersistSmem<128, STAGES>;
  static_assert(S::TOTAL <= 227 * 1024, "pair tile configuration exceeds shared memory");
  if (splits > 1 && sizeof(TC) != 4) return set_error("gemm_tc_pair: split-K needs an fp32 output");
  if (g_sm_count == 0) {
    int dev = 0;
    B200ST_CUDA(cudaGetDevice(&dev));
    B200ST_CUDA(cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  const int n_tiles_n = (int)ceil_div(N, 256);
  const int n_tiles = (int)(ceil_div(M, 256) * n_tiles_n);
  auto kern = gemm_tc_pair_kernel<STAGES, A_MN, B_MN, TC>;
  B200ST_CUDA(cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
  dim3 grid((unsigned)(2 * min(n_tiles * splits, sm_budget() / 2)));
  B200ST_CUDA(launch_pdl_pair(kern, grid, dim3(TCP_THREADS), S::TOTAL, st, ma, mb, ma2, mb2, kb_seg1, (TC*)C, ldc,
                              (const TC*)R, ldr, bias, relu, alpha, (int)M, (int)N, (int)K, n_tiles_n, n_tiles, g_pair_dbg, splits, kb_per_split));
  B200ST_LAUNCH_CHECK("gemm_tc_pair");
  return 0;
}

static int g_clk_enabled = 1;
void gemm_tc_set_cluster_splitk(int on) { g_clk_enabled = on; }

int gemm_tc(int dtype_c, int ta, int tb, int64_t M, int64_t N, int64_t K, float alpha, const void* A, int64_t lda,
            const void* B, int64_t ldb, void* C, int64_t ldc, const void* R, int64_t ldr, const float* bias, int relu,
            cudaStream_t st, const void* A2, int64_t lda2, const void* B2, int64_t ldb2, int64_t K2) {
  // optional second operand pair (same op() forms, K2 deep): C = alpha (op(A) op(B) + op(A2) op(B2)) + ...; the
  // first segment must end on a k-block boundary
  const bool dual = A2 != nullptr && K2 > 0;
  if (dual && K % TC_BK) return set_error("gemm_tc: two-segment GEMM needs K %% 64 == 0 for the first segment");
  const int64_t K1 = K;
  if (dual) K += K2;
  // op(A) is M x K: ta=0 -> stored [M,K] (K-major); ta=1 -> stored [K,M] (M-major).
  // op(B) is K x N: tb=1 -> stored [N,K] (K-major); tb=0 -> stored [K,N] (N-major).
  const bool a_mn = ta != 0, b_mn = tb == 0;
  const int kb_total = (int)ceil_div(K, TC_BK);
  const int64_t m_tiles = ceil_div(M, TC_BM);
  // ---- tile configuration
  //  cfg 0: 128 x 128, 3 stages, 2 CTAs/SM — throughput shape (one CTA's epilogue overlaps the other's mainloop)
  //  cfg 1: 128 x  64, 4 stages, 2 CTAs/SM — mid-size problems that would not fill the 148 SMs with 128-wide tiles
  //  cfg 2: 128 x  32, 10 stages — single-M-tile (decoder step, M <= 128) with K-major B: narrow N, deep pipeline
  //  cfg 3: 128 x  64, 8 stages  — single-M-tile with N-major B (128B-swizzled MN-major boxes are 64 wide)
  // single-M-tile problems with a very wide N (the per-step vocabulary projection): keep the grid within one wave
  const bool wide = m_tiles == 1 && M <= 64 && !b_mn && ceil_div(N, 64) > 148;
  int cfg;
  if (m_tiles == 1 && N > 32) cfg = b_mn ? 3 : 2;
  else if (N <= 64 || m_tiles * ceil_div(N, 128) < 148) cfg = 1;
  else cfg = 0;
  if (N <= 32 && !b_mn) cfg = 2;
  // cfg 4: 128 x 64, 2 stages, 4 CTAs/SM -- many M tiles with a one- or two-block K (the first BLSTM layer's input
  // projection, K = acoustic dim): a tile is all prologue + epilogue latency, so what helps is more tiles in flight
  if (cfg != 2 && kb_total <= 2 && m_tiles >= 64) cfg = 4;
  // persistent kernel (one CTA per SM, double-buffered TMEM accumulator) for the throughput shapes: at least four
  // tiles per SM, fp32-atomic split-K not involved.  128 x 256 tiles when N allows, else 128 x 128.
  // Persistent kernels for the many-tile shapes (cfg 0, and cfg 4 = one/two-block K with many M tiles, which is bound by
  // its output write): at least `min_rounds` tiles per CTA.  CTA-pair kernel (256 x 256 tile per pair of SMs) when it
  // needs no more rounds than 128 x 256 single-CTA tiles would (wave quantisation; rows beyond M in the last 256-row
  // tile are wasted work) -- it halves the L2 -> shared-memory traffic per FLOP.
  static const int min_rounds = [] { const char* e = getenv("B200ST_GEMM_MIN_ROUNDS"); return e ? atoi(e) : 2; }();
  const bool many = g_persist_enabled && (cfg == 0 || cfg == 4) && N >= 128;
  const int64_t t_pair = ceil_div(M, 256) * ceil_div(N, 256), t_256 = m_tiles * ceil_div(N, 256), t_128 = m_tiles * ceil_div(N, 128);
  const bool can256 = many && N >= 256 && t_256 >= (int64_t)min_rounds * 148;
  const bool pair = many && g_pair_enabled && N >= 256 && kb_total >= 8 && t_pair >= 48 &&
                    (!can256 || ceil_div(t_pair, 74) <= ceil_div(t_256, 148));
  // split-K on the pair kernel: long-K weight gradients with few 256 x 256 tiles (fp32 output, plain sum)
  const bool pair_splitk = g_persist_enabled && g_pair_enabled && !pair && dtype_c == B200ST_F32 && !dual && !bias && !relu &&
                           !R && N >= 256 && M >= 256 && t_pair <= 37 && kb_total >= 64;
  const bool persist256 = !pair && !pair_splitk && can256;
  const bool persist128 = !pair && !pair_splitk && !persist256 && many && t_128 >= (int64_t)min_rounds * 148;
  if (pair || pair_splitk || persist256 || persist128) cfg = 0;
  const int BN = persist256 ? 256 : ((cfg == 0 || wide || pair || pair_splitk) ? 128 : (cfg == 2 ? 32 : 64));   // = B box rows
  const bool m64 = (cfg == 2 || cfg == 3) && M <= 64;       // half-height A tile for the decoder-step GEMMs
  CUtensorMap ma, mb;
  if (a_mn) { if (make_map(&ma, A, K1, M, lda, TC_BK)) return -1; }
  else      { if (make_map(&ma, A, M, K1, lda, m64 ? 64 : TC_BM)) return -1; }
  if (b_mn) { if (make_map(&mb, B, K1, N, ldb, TC_BK)) return -1; }
  else      { if (make_map(&mb, B, N, K1, ldb, BN)) return -1; }
  CUtensorMap ma2 = ma, mb2 = mb;
  int kb_seg1 = 0x7fffffff;
  if (dual) {
    if (a_mn) { if (make_map(&ma2, A2, K2, M, lda2, TC_BK)) return -1; }
    else      { if (make_map(&ma2, A2, M, K2, lda2, m64 ? 64 : TC_BM)) return -1; }
    if (b_mn) { if (make_map(&mb2, B2, K2, N, ldb2, TC_BK)) return -1; }
    else      { if (make_map(&mb2, B2, N, K2, ldb2, BN)) return -1; }
    kb_seg1 = (int)(K1 / TC_BK);
  }
  // cluster split-K: decoder-step shapes whose few N tiles would each stream a long K alone
  const int64_t tiles_1cta = ceil_div(N, BN);       // CTAs the non-split configuration would launch
  if (g_clk_enabled && !dual && m64 && !a_mn && kb_total >= 8 && (tiles_1cta <= 16 || (tiles_1cta <= 32 && kb_total >= 16))) {
    int cs = kb_total >= 16 ? 8 : 4;
    while (cs > 2 && ceil_div(N, 64) * cs > 160) cs >>= 1;
    const int kps = (int)ceil_div(kb_total, cs);
    CUtensorMap mb64 = mb;
    if (!b_mn && BN != 64) { if (make_map(&mb64, B, N, K, ldb, 64)) return -1; }
    if (b_mn) {
      if (dtype_c == B200ST_F32) return launch_tc_clk<64, 4, false, true, float>(M, N, K, alpha, ma, mb64, C, ldc, R, ldr, bias, relu, cs, kps, st);
      return launch_tc_clk<64, 4, false, true, __nv_bfloat16>(M, N, K, alpha, ma, mb64, C, ldc, R, ldr, bias, relu, cs, kps, st);
    }
    if (dtype_c == B200ST_F32) return launch_tc_clk<64, 4, false, false, float>(M, N, K, alpha, ma, mb64, C, ldc, R, ldr, bias, relu, cs, kps, st);
    return launch_tc_clk<64, 4, false, false, __nv_bfloat16>(M, N, K, alpha, ma, mb64, C, ldc, R, ldr, bias, relu, cs, kps, st);
  }
  // split-K when the output has few tiles and K is long (weight gradients): fp32 output, plain sum only.
  const int64_t tiles = m_tiles * ceil_div(N, BN);
  int splits = 1;
  if (pair_splitk) {              // units = pair tiles x K splits fill the 74 SM pairs once
    splits = (int)max((int64_t)1, 74 / t_pair);
    if (splits > kb_total / 16) splits = max(1, kb_total / 16);
  } else
  if (dtype_c == B200ST_F32 && !dual && !bias && !relu && !R && tiles < 96 && kb_total >= 16) {
    splits = (int)(296 / tiles);
    if (splits > kb_total / 4) splits = kb_total / 4;
    if (splits < 1) splits = 1;
  }
  int kb_per_split = (int)ceil_div(kb_total, splits);
  splits = (int)ceil_div(kb_total, kb_per_split);
  if (splits > 1) B200ST_CUDA(cudaMemset2DAsync(C, ldc * 4, 0, N * 4, M, st));
#define TC_PERSIST(BN_, ST_, AMN, BMN)                                                                     \
  do {                                                                                                     \
    if (dtype_c == B200ST_F32)                                                                             \
      return launch_tc_persist<BN_, ST_, AMN, BMN, float>(M, N, K, alpha, ma, mb, ma2, mb2, kb_seg1, C, ldc, R, ldr, bias, relu, st); \
    return launch_tc_persist<BN_, ST_, AMN, BMN, __nv_bfloat16>(M, N, K, alpha, ma, mb, ma2, mb2, kb_seg1, C, ldc, R, ldr, bias, relu, st); \
  } while (0)
#define TC_PAIR(AMN, BMN)                                                                                  \
  do {                                                                                                     \
    if (dtype_c == B200ST_F32)                                                                             \
      return launch_tc_pair<6, AMN, BMN, float>(M, N, K, alpha, ma, mb, ma2, mb2, kb_seg1, C, ldc, R, ldr, bias, relu, st); \
    return launch_tc_pair<6, AMN, BMN, __nv_bfloat16>(M, N, K, alpha, ma, mb, ma2, mb2, kb_seg1, C, ldc, R, ldr, bias, relu, st); \
  } while (0)
  if (pair_splitk) {
#define TC_PAIRK(AMN, BMN) return launch_tc_pair<6, AMN, BMN, float>(M, N, K, alpha, ma, mb, ma2, mb2, kb_seg1, C, ldc, R, ldr, bias, relu, st, splits, kb_per_split)
    if (!a_mn && !b_mn) TC_PAIRK(false, false);
    if (!a_mn && b_mn) TC_PAIRK(false, true);
    if (a_mn && !b_mn) TC_PAIRK(true, false);
    TC_PAIRK(true, true);
#undef TC_PAIRK
  }
  if (pair && splits == 1) {
    if (!a_mn && !b_mn) TC_PAIR(false, false);
    if (!a_mn && b_mn) TC_PAIR(false, true);
    if (a_mn && !b_mn) TC_PAIR(true, false);
    TC_PAIR(true, true);
  }
#undef TC_PAIR
  if ((persist256 || persist128) && splits == 1) {
    if (persist256) {
      if (!a_mn && !b_mn) TC_PERSIST(256, 4, false, false);
      if (!a_mn && b_mn) TC_PERSIST(256, 4, false, true);
      if (a_mn && !b_mn) TC_PERSIST(256, 4, true, false);
      TC_PERSIST(256, 4, true, true);
    }
    if (!a_mn && !b_mn) TC_PERSIST(128, 6, false, false);
    if (!a_mn && b_mn) TC_PERSIST(128, 6, false, true);
    if (a_mn && !b_mn) TC_PERSIST(128, 6, true, false);
    TC_PERSIST(128, 6, true, true);
  }
#undef TC_PERSIST
#define TC_GO(BN_, ST_, AMN, BMN)                                                                          \
  do {                                                                                                     \
    if (dtype_c == B200ST_F32)                                                                             \
      return launch_tc<BN_, ST_, AMN, BMN, float>(M, N, K, alpha, ma, mb, ma2, mb2, kb_seg1, C, ldc, R, ldr, bias, relu, splits, kb_per_split, st); \
    return launch_tc<BN_, ST_, AMN, BMN, __nv_bfloat16>(M, N, K, alpha, ma, mb, ma2, mb2, kb_seg1, C, ldc, R, ldr, bias, relu, splits, kb_per_split, st); \
  } while (0)
#define TC_GO64(BN_, ST_, AMN, BMN)                                                                        \
  do {                                                                                                     \
    if (dtype_c == B200ST_F32)                                                                             \
      return launch_tc<BN_, ST_, AMN, BMN, float, 64>(M, N, K, alpha, ma, mb, ma2, mb2, kb_seg1, C, ldc, R, ldr, bias, relu, splits, kb_per_split, st); \
    return launch_tc<BN_, ST_, AMN, BMN, __nv_bfloat16, 64>(M, N, K, alpha, ma, mb, ma2, mb2, kb_seg1, C, ldc, R, ldr, bias, relu, splits, kb_per_split, st); \
  } while (0)
#define TC_CFG(AMN, BMN)                                                   \
  do {                                                                     \
    if (cfg == 0) TC_GO(128, 3, AMN, BMN);                                 \
    if (cfg == 1) TC_GO(64, 4, AMN, BMN);                                  \
    if (cfg == 3) TC_GO(64, 8, AMN, BMN);                                  \
    if (cfg == 4) TC_GO(64, 2, AMN, BMN);                                  \
  } while (0)
  if (wide) { if (!a_mn) TC_GO64(128, 6, false, false); TC_GO64(128, 6, true, false); }
  if (m64) {                    // M <= 64: half-height tiles, deeper pipelines (12 KB / 16 KB per stage)
    if (cfg == 2) { if (!a_mn) TC_GO64(32, 12, false, false); TC_GO64(32, 12, true, false); }
    if (!a_mn && b_mn) TC_GO64(64, 10, false, true);
    if (a_mn && b_mn) TC_GO64(64, 10, true, true);
  }
  if (cfg == 2) {               // K-major B only
    if (!a_mn) TC_GO(32, 10, false, false);
    TC_GO(32, 10, true, false);
  }
  if (!a_mn && !b_mn) TC_CFG(false, false);
  if (!a_mn && b_mn) TC_CFG(false, true);
  if (a_mn && !b_mn) TC_CFG(true, false);
  TC_CFG(true, true);
  return set_error("gemm_tc: no tile configuration");
#undef TC_CFG
#undef TC_GO
#undef TC_GO64
}

}  // namespace b200st
