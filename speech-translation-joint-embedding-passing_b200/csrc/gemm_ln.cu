// GEMM with a fused residual + LayerNorm epilogue (bf16, tcgen05 / TMEM / TMA, sm_100a):
//
//   Y[M, 512]  = A[M, K] . W[512, K]^T (+ bias) + R                      (the sub-layer output, layers.py:190-197, 247-252)
//   YN[M, 512] = LayerNorm(Y; gamma, beta, eps),  mean[M], rstd[M]        (the NEXT sub-layer's pre-norm, layers.py:153, 245;
//                                                                         TFEnc.py:89, TFDec.py:127 for the final norm)
//
// Every residual sub-layer of the Transformer ends in a projection back to d_model = 512 with the skip connection added,
// and the next sub-layer starts by normalising exactly that tensor.  A LayerNorm needs whole rows, a GEMM CTA owns a
// 128 x 128 tile: the four CTAs that cover one 128-row block form a thread-block CLUSTER (4 x 1 x 1 along N), each
// computes (mean, M2) of its 128 columns per row, the partials are exchanged through distributed shared memory with one
// cluster barrier and combined with Chan's formula, and each CTA normalises its own columns.  The statistics are taken
// from the bf16-rounded Y (what the separate LayerNorm kernel would read back), in fp32.
// The accumulator tile stays in TMEM between the two passes (pass 1 writes Y and stores the rounded values back with
// tcgen05.st; pass 2 re-reads them), so no row is ever staged in registers or shared memory.
//
// Warp roles as in gemm_tc.cu: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7 epilogue (one
// accumulator row = one TMEM lane per thread).
#include <cuda.h>

#include "common.cuh"
#include "philox.cuh"
#include "umma.cuh"

namespace b200st {

constexpr int GL_BM = 128, GL_BN = 128, GL_BK = 64, GL_N = 512, GL_CL = GL_N / GL_BN, GL_STAGES = 4, GL_THREADS = 256;
constexpr int GL_A_BYTES = GL_BM * GL_BK * 2, GL_B_BYTES = GL_BN * GL_BK * 2, GL_STAGE = GL_A_BYTES + GL_B_BYTES;
constexpr int GL_BAR_OFF = GL_STAGES * GL_STAGE;                       // full[4], empty[4], tmem_full, tmem slot
constexpr int GL_STAT_OFF = GL_BAR_OFF + 128;                          // float2 stats[4 src][128 rows]
constexpr int GL_SMEM = GL_STAT_OFF + GL_CL * GL_BM * 8 + 1024;        // + alignment slack

__device__ __forceinline__ uint32_t gl_mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}

__device__ __forceinline__ float gl_round_bf16(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

__global__ void __cluster_dims__(GL_CL, 1, 1) __launch_bounds__(GL_THREADS, 1)
gemm_ln_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const __nv_bfloat16* __restrict__ R, int64_t ldr, const float* __restrict__ bias,
               __nv_bfloat16* __restrict__ Y, int64_t ldy, const float* __restrict__ gamma,
               const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ YN, int64_t ldyn,
               float* __restrict__ mean_out, float* __restrict__ rstd_out, int M, int K, float drop_p,
               const int64_t* __restrict__ rng, int64_t site) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + GL_BAR_OFF);
  uint64_t* empty = full + GL_STAGES;
  uint64_t* tmem_full = empty + GL_STAGES;
  uint32_t* tmem_slot = (uint32_t*)(tmem_full + 1);
  float2* stats = (float2*)(smem + GL_STAT_OFF);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();               // == blockIdx.x: the 128-column slice of this CTA
  const int m0 = blockIdx.y * GL_BM, n0 = blockIdx.x * GL_BN;
  const int n_iter = (K + GL_BK - 1) / GL_BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < GL_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)GL_BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                      // operands / residual / outputs may only be touched from here on (see common.cuh)
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % GL_STAGES;
        mbar_wait(&empty[s], ((it / GL_STAGES) & 1) ^ 1);
        uint8_t* sa = smem + s * GL_STAGE;
        mbar_expect_tx(&full[s], GL_STAGE);
        tma_load_2d(sa, &tma_a, it * GL_BK, m0, &full[s]);
        tma_load_2d(sa + GL_A_BYTES, &tma_b, it * GL_BK, n0, &full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(GL_BM, GL_BN, false, false);
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % GL_STAGES;
        mbar_wait(&full[s], (it / GL_STAGES) & 1);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * GL_STAGE), sb = sa + GL_A_BYTES;
#pragma unroll
        for (int k = 0; k < GL_BK / 16; ++k)
          tc_mma_f16(tmem_base, umma_desc(sa + k * 32, 16, 1024), umma_desc(sb + k * 32, 16, 1024), idesc,
                     (it > 0 || k > 0) ? 1u : 0u);
        tc_commit(&empty[s]);
      }
      tc_commit(tmem_full);
    }
  } else if (warp >= 4) {
    const int wq = warp - 4;
    const int rl = wq * 32 + lane;                 // row inside the tile == TMEM lane
    const int row = m0 + rl;
    const bool row_ok = row < M;
    const uint32_t trow = tmem_base + ((uint32_t)(wq * 32) << 16);
    // the residual row slice (128 bf16 = 256 B) is requested before the accumulator is complete
    uint4 rres[16];
    if (R != nullptr && row_ok) {
      const uint4* rp = reinterpret_cast<const uint4*>(R + (int64_t)row * ldr + n0);
#pragma unroll
      for (int i = 0; i < 16; ++i) rres[i] = rp[i];
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) rres[i] = make_uint4(0, 0, 0, 0);
    }
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    // ---- pass 1: y = acc + bias + residual -> bf16 -> Y; the rounded values go back to TMEM; row sum
    float sum = 0.f;
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      uint32_t r[32];
      tmem_ld32(trow + ch * 32, r);
      float v[32];
#pragma unroll
      for (int q = 0; q < 32; ++q) v[q] = __uint_as_float(r[q]);
      if (bias) {
#pragma unroll
        for (int q = 0; q < 32; q += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(bias + n0 + ch * 32 + q);
          v[q] += b4.x; v[q + 1] += b4.y; v[q + 2] += b4.z; v[q + 3] += b4.w;
        }
      }
      if (drop_p > 0.f) {
        // dropout on the projection's output before the skip connection (layers.py:194-195, 248-250): the same counter-based
        // mask b200st_dropout draws for a dense [M, 512] tensor (element index row * 512 + column, one Philox call per four
        // columns), applied to the bf16-rounded value the separate GEMM would have stored; backward re-draws it
        DropRng dr;
        dr.init(rng, site, drop_p);
        const uint64_t g0 = ((uint64_t)row * GL_N + (uint64_t)(n0 + ch * 32)) >> 2;
#pragma unroll
        for (int q = 0; q < 32; q += 4) {
          const Philox4 rr = dr.group(g0 + (q >> 2));
#pragma unroll
          for (int e = 0; e < 4; ++e) v[q + e] = rr.v[e] >= dr.thresh ? gl_round_bf16(v[q + e]) * dr.scale : 0.f;
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint4 rr = rres[ch * 4 + i];
        const uint32_t w[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          v[i * 8 + 2 * e] += __uint_as_float(w[e] << 16);
          v[i * 8 + 2 * e + 1] += __uint_as_float(w[e] & 0xffff0000u);
        }
      }
      uint32_t pk[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        __nv_bfloat162 b2 = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
        pk[e] = *reinterpret_cast<uint32_t*>(&b2);
        r[2 * e] = pk[e] << 16;                           // the bf16-rounded values, as fp32 bits
        r[2 * e + 1] = pk[e] & 0xffff0000u;
        sum += __uint_as_float(r[2 * e]) + __uint_as_float(r[2 * e + 1]);
      }
      tmem_st32(trow + ch * 32, r);
      if (row_ok) {
        __nv_bfloat16* yp = Y + (int64_t)row * ldy + n0 + ch * 32;
        st_global_v8(yp, pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7]);
        st_global_v8(yp + 16, pk[8], pk[9], pk[10], pk[11], pk[12], pk[13], pk[14], pk[15]);
      }
    }
    // ---- local second moment about the local mean
    const float lm = sum * (1.f / GL_BN);
    float m2 = 0.f;
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      uint32_t r[32];
      tmem_ld32(trow + ch * 32, r);
#pragma unroll
      for (int q = 0; q < 32; ++q) { const float d = __uint_as_float(r[q]) - lm; m2 = fmaf(d, d, m2); }
    }
    // ---- hand (mean, M2) of this 128-column slice to all four CTAs of the row block
    const uint32_t slot = smem_u32(&stats[rank * GL_BM + rl]);
#pragma unroll
    for (uint32_t d = 0; d < GL_CL; ++d)
      asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(gl_mapa(slot, d)), "f"(lm), "f"(m2) : "memory");
  }
  // one cluster barrier: every thread of the four CTAs arrives (the epilogue warps after their remote stores)
  __syncwarp();                    // warps 0 / 1: the elected lane's role loop has ended before the aligned barrier
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  if (warp >= 4) {
    const int wq = warp - 4;
    const int rl = wq * 32 + lane;
    const int row = m0 + rl;
    const bool row_ok = row < M;
    const uint32_t trow = tmem_base + ((uint32_t)(wq * 32) << 16);
    float pm[GL_CL], mu = 0.f, m2 = 0.f;
#pragma unroll
    for (int d = 0; d < GL_CL; ++d) { const float2 s2 = stats[d * GL_BM + rl]; pm[d] = s2.x; mu += s2.x; m2 += s2.y; }
    mu *= (1.f / GL_CL);
#pragma unroll
    for (int d = 0; d < GL_CL; ++d) { const float dd = pm[d] - mu; m2 = fmaf((float)GL_BN * dd, dd, m2); }
    const float rs = rsqrtf(m2 * (1.f / GL_N) + eps);
    if (rank == 0 && row_ok) {
      if (mean_out) mean_out[row] = mu;
      if (rstd_out) rstd_out[row] = rs;
    }
    // ---- pass 2: normalise this CTA's 128 columns
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      uint32_t r[32];
      tmem_ld32(trow + ch * 32, r);
      uint32_t pk[16];
#pragma unroll
      for (int q = 0; q < 32; q += 4) {
        const float4 g4 = *reinterpret_cast<const float4*>(gamma + n0 + ch * 32 + q);
        const float4 b4 = *reinterpret_cast<const float4*>(beta + n0 + ch * 32 + q);
        const float o0 = (__uint_as_float(r[q]) - mu) * rs * g4.x + b4.x, o1 = (__uint_as_float(r[q + 1]) - mu) * rs * g4.y + b4.y;
        const float o2 = (__uint_as_float(r[q + 2]) - mu) * rs * g4.z + b4.z, o3 = (__uint_as_float(r[q + 3]) - mu) * rs * g4.w + b4.w;
        __nv_bfloat162 lo = __floats2bfloat162_rn(o0, o1), hi = __floats2bfloat162_rn(o2, o3);
        pk[q / 2] = *reinterpret_cast<uint32_t*>(&lo);
        pk[q / 2 + 1] = *reinterpret_cast<uint32_t*>(&hi);
      }
      if (row_ok) {
        __nv_bfloat16* yp = YN + (int64_t)row * ldyn + n0 + ch * 32;
        st_global_v8(yp, pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7]);
        st_global_v8(yp + 16, pk[8], pk[9], pk[10], pk[11], pk[12], pk[13], pk[14], pk[15]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)GL_BN) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Backward twin: the input-gradient GEMM that feeds a LayerNorm backward, with the LayerNorm backward as its epilogue.
//   G[M, 512]  = A[M, K] . W[K, 512]                          (dqn = dqp . w_qs, dy = dz . w_1: the gradient of LN's output)
//   DX[M, 512] = rstd * (g - mean_c(g) - xhat * mean_c(g * xhat)) + ADD,   g = G * gamma, xhat = (X - mean) * rstd
//   PART[row block, 0:512] = sum_rows G * xhat (dgamma),  PART[row block, 512:1024] = sum_rows G (dbeta)
// Row reductions cross the 4-CTA cluster like the forward kernel's statistics; the column sums of a CTA's 128 rows go
// through a shared-memory transpose (the idle pipeline stages) and leave as per-row-block partials that the caller
// column-sums off the critical path (same format as b200st_layernorm_bwd_partial).  g and xhat live in TMEM between the
// two passes (256 columns).
constexpr int GLB_TMEM = 256;
constexpr int GLB_PITCH = 33;                         // staging pitch (floats): conflict-free row writes and column reads

__global__ void __cluster_dims__(GL_CL, 1, 1) __launch_bounds__(GL_THREADS, 1)
gemm_lnbwd_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                  const __nv_bfloat16* __restrict__ X, const float* __restrict__ gamma, const float* __restrict__ mean,
                  const float* __restrict__ rstd, const __nv_bfloat16* __restrict__ ADD, __nv_bfloat16* __restrict__ DX,
                  float* __restrict__ PART, int M, int K) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + GL_BAR_OFF);
  uint64_t* empty = full + GL_STAGES;
  uint64_t* tmem_full = empty + GL_STAGES;
  uint32_t* tmem_slot = (uint32_t*)(tmem_full + 1);
  float2* stats = (float2*)(smem + GL_STAT_OFF);
  float* stg_a = (float*)smem;                               // [128][33]  G * xhat   (pipeline stages, idle after the mainloop)
  float* stg_b = stg_a + GL_BM * GLB_PITCH;                  // [128][33]  G
  float* red = stg_b + GL_BM * GLB_PITCH;                    // [4 parts][2][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int m0 = blockIdx.y * GL_BM, n0 = blockIdx.x * GL_BN;
  const int n_iter = (K + GL_BK - 1) / GL_BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < GL_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)GLB_TMEM) : "memory");
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
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % GL_STAGES;
        mbar_wait(&empty[s], ((it / GL_STAGES) & 1) ^ 1);
        uint8_t* sa = smem + s * GL_STAGE;
        uint8_t* sb = sa + GL_A_BYTES;
        mbar_expect_tx(&full[s], GL_STAGE);
        tma_load_2d(sa, &tma_a, it * GL_BK, m0, &full[s]);
        // W is [K, 512] with the 512 output columns contiguous: staged MN-major, two 64-column boxes of 64 k-rows
        tma_load_2d(sb, &tma_b, n0, it * GL_BK, &full[s]);
        tma_load_2d(sb + GL_BK * 128, &tma_b, n0 + 64, it * GL_BK, &full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(GL_BM, GL_BN, false, true);
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % GL_STAGES;
        mbar_wait(&full[s], (it / GL_STAGES) & 1);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * GL_STAGE), sb = sa + GL_A_BYTES;
#pragma unroll
        for (int k = 0; k < GL_BK / 16; ++k)
          tc_mma_f16(tmem_base, umma_desc(sa + k * 32, 16, 1024), umma_desc(sb + k * 2048, GL_BK * 128, 1024), idesc,
                     (it > 0 || k > 0) ? 1u : 0u);
        tc_commit(&empty[s]);
      }
      tc_commit(tmem_full);
    }
  } else if (warp >= 4) {
    const int wq = warp - 4, et = threadIdx.x - 128;         // et: 0..127 among the epilogue threads
    const int rl = wq * 32 + lane;
    const int row = m0 + rl;
    const bool row_ok = row < M;
    const uint32_t trow = tmem_base + ((uint32_t)(wq * 32) << 16);
    uint4 xr[16];
    float mu = 0.f, rs = 0.f;
    if (row_ok) {
      const uint4* xp = reinterpret_cast<const uint4*>(X + (int64_t)row * GL_N + n0);
#pragma unroll
      for (int i = 0; i < 16; ++i) xr[i] = xp[i];
      mu = mean[row];
      rs = rstd[row];
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) xr[i] = make_uint4(0, 0, 0, 0);
    }
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    // ---- pass 1: g = G * gamma and xhat -> TMEM; row sums of g and g * xhat; column sums of G * xhat and G
    float s1 = 0.f, s2 = 0.f;
    const int rc = et & 31, rp = et >> 5;                    // reduce role: column rc of the chunk, rows 32 rp .. 32 rp + 31
    float cga[4], cgb[4];
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      uint32_t r[32], xh[32];
      tmem_ld32(trow + ch * 32, r);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint4 xx = xr[ch * 4 + i];
        const uint32_t w[4] = {xx.x, xx.y, xx.z, xx.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          xh[i * 8 + 2 * e] = __float_as_uint((__uint_as_float(w[e] << 16) - mu) * rs);
          xh[i * 8 + 2 * e + 1] = __float_as_uint((__uint_as_float(w[e] & 0xffff0000u) - mu) * rs);
        }
      }
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        const float G = __uint_as_float(r[q]), xv = __uint_as_float(xh[q]);
        stg_a[rl * GLB_PITCH + q] = G * xv;
        stg_b[rl * GLB_PITCH + q] = G;
      }
#pragma unroll
      for (int q = 0; q < 32; q += 4) {
        const float4 g4 = *reinterpret_cast<const float4*>(gamma + n0 + ch * 32 + q);
        const float gm[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float g = __uint_as_float(r[q + e]) * gm[e];
          r[q + e] = __float_as_uint(g);
          s1 += g;
          s2 = fmaf(g, __uint_as_float(xh[q + e]), s2);
        }
      }
      tmem_st32(trow + ch * 32, r);
      tmem_st32(trow + 128 + ch * 32, xh);
      asm volatile("bar.sync 1, 128;" ::: "memory");        // the chunk is staged by all four epilogue warps
      float a = 0.f, b = 0.f;
#pragma unroll 8
      for (int rr = 0; rr < 32; ++rr) {
        a += stg_a[(rp * 32 + rr) * GLB_PITCH + rc];
        b += stg_b[(rp * 32 + rr) * GLB_PITCH + rc];
      }
      cga[ch] = a;
      cgb[ch] = b;
      asm volatile("bar.sync 1, 128;" ::: "memory");        // staging buffers may be overwritten
    }
    // combine the four row parts per column and emit this row block's partial dgamma | dbeta
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      red[(rp * 2 + 0) * GL_BN + ch * 32 + rc] = cga[ch];
      red[(rp * 2 + 1) * GL_BN + ch * 32 + rc] = cgb[ch];
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    {
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int p = 0; p < 4; ++p) { a += red[(p * 2 + 0) * GL_BN + et]; b += red[(p * 2 + 1) * GL_BN + et]; }
      float* pp = PART + (int64_t)blockIdx.y * (2 * GL_N);
      pp[n0 + et] = a;
      pp[GL_N + n0 + et] = b;
    }
    const uint32_t slot = smem_u32(&stats[rank * GL_BM + rl]);
#pragma unroll
    for (uint32_t d = 0; d < GL_CL; ++d)
      asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(gl_mapa(slot, d)), "f"(s1), "f"(s2) : "memory");
  }
  __syncwarp();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  if (warp >= 4) {
    const int wq = warp - 4;
    const int rl = wq * 32 + lane;
    const int row = m0 + rl;
    const bool row_ok = row < M;
    const uint32_t trow = tmem_base + ((uint32_t)(wq * 32) << 16);
    uint4 ar[16];
    if (ADD != nullptr && row_ok) {
      const uint4* ap = reinterpret_cast<const uint4*>(ADD + (int64_t)row * GL_N + n0);
#pragma unroll
      for (int i = 0; i < 16; ++i) ar[i] = ap[i];
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) ar[i] = make_uint4(0, 0, 0, 0);
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int d = 0; d < GL_CL; ++d) { const float2 v = stats[d * GL_BM + rl]; s1 += v.x; s2 += v.y; }
    const float rs = row_ok ? rstd[row] : 0.f;
    const float m1 = s1 * (1.f / GL_N), m2 = s2 * (1.f / GL_N);
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      uint32_t g[32], xh[32];
      tmem_ld32(trow + ch * 32, g);
      tmem_ld32(trow + 128 + ch * 32, xh);
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint4 aa = ar[ch * 4 + i];
        const uint32_t w[4] = {aa.x, aa.y, aa.z, aa.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int q = i * 8 + 2 * e;
          const float d0 = rs * (__uint_as_float(g[q]) - m1 - __uint_as_float(xh[q]) * m2) + __uint_as_float(w[e] << 16);
          const float d1 = rs * (__uint_as_float(g[q + 1]) - m1 - __uint_as_float(xh[q + 1]) * m2) + __uint_as_float(w[e] & 0xffff0000u);
          __nv_bfloat162 b2 = __floats2bfloat162_rn(d0, d1);
          pk[i * 4 + e] = *reinterpret_cast<uint32_t*>(&b2);
        }
      }
      if (row_ok) {
        __nv_bfloat16* dp = DX + (int64_t)row * GL_N + n0 + ch * 32;
        st_global_v8(dp, pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7]);
        st_global_v8(dp + 16, pk[8], pk[9], pk[10], pk[11], pk[12], pk[13], pk[14], pk[15]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)GLB_TMEM) : "memory");
  }
}

}  // namespace b200st

using namespace b200st;

extern "C" {

int b200st_gemm_ln_eligible(int dtype, int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* W,
                            int64_t ldw, const void* R, int64_t ldr, const void* Y, int64_t ldy, const void* YN,
                            int64_t ldyn, const float* bias, const float* gamma, const float* beta) {
  if (dtype != B200ST_BF16 || N != GL_N || M < 1 || K < 64 || K % 8) return 0;
  if (lda % 8 || ldw % 8 || ldr % 8 || ldy % 16 || ldyn % 16) return 0;
  if (((uintptr_t)A & 15) || ((uintptr_t)W & 15) || ((uintptr_t)R & 15) || ((uintptr_t)Y & 31) || ((uintptr_t)YN & 31)) return 0;
  if (((uintptr_t)bias & 15) || ((uintptr_t)gamma & 15) || ((uintptr_t)beta & 15)) return 0;
  if (M >= (1ll << 31) || K >= (1ll << 31)) return 0;
  return 1;
}

int b200st_gemm_ln(int dtype, int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* W, int64_t ldw,
                   const float* bias, const void* R, int64_t ldr, void* Y, int64_t ldy, const float* gamma,
                   const float* beta, float eps, void* YN, int64_t ldyn, float* mean, float* rstd, float drop_p,
                   const int64_t* rng, int64_t site, b200st_stream_t stream) {
  if (!(drop_p >= 0.f && drop_p < 1.f) || (drop_p > 0.f && rng == nullptr))
    return set_error("gemm_ln: dropout p=%f outside [0, 1) or no rng state", (double)drop_p);
  if (drop_p > 0.f && ldy != GL_N) return set_error("gemm_ln: the fused dropout mask is defined for a dense [M, 512] output");
  if (!b200st_gemm_ln_eligible(dtype, M, N, K, A, lda, W, ldw, R, ldr, Y, ldy, YN, ldyn, bias, gamma, beta))
    return set_error("gemm_ln: needs bf16, N = 512, K %% 8 == 0, 16-byte aligned operands (got M=%lld N=%lld K=%lld)",
                     (long long)M, (long long)N, (long long)K);
  if (!gamma || !beta || !Y || !YN) return set_error("gemm_ln: gamma, beta, Y and YN are required");
  CUtensorMap ma, mb;
  if (make_map(&ma, A, M, K, lda, GL_BM)) return -1;
  if (make_map(&mb, W, N, K, ldw, GL_BN)) return -1;
  B200ST_CUDA(cudaFuncSetAttribute((const void*)gemm_ln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GL_SMEM));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(GL_CL, (unsigned)((M + GL_BM - 1) / GL_BM), 1);
  cfg.blockDim = dim3(GL_THREADS);
  cfg.dynamicSmemBytes = GL_SMEM;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  B200ST_CUDA(cudaLaunchKernelEx(&cfg, gemm_ln_kernel, ma, mb, (const __nv_bfloat16*)R, ldr, bias, (__nv_bfloat16*)Y, ldy,
                                 gamma, beta, eps, (__nv_bfloat16*)YN, ldyn, mean, rstd, (int)M, (int)K, drop_p, rng, site));
  B200ST_LAUNCH_CHECK("gemm_ln");
  return 0;
}

int b200st_gemm_lnbwd_eligible(int dtype, int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* W,
                               int64_t ldw, const void* X, const void* ADD, const void* DX, const float* gamma) {
  if (dtype != B200ST_BF16 || N != GL_N || M < 1 || K < 64 || K % 8) return 0;
  if (lda % 8 || ldw % 8) return 0;
  if (((uintptr_t)A & 15) || ((uintptr_t)W & 15) || ((uintptr_t)X & 15) || ((uintptr_t)ADD & 15) || ((uintptr_t)DX & 31) ||
      ((uintptr_t)gamma & 15))
    return 0;
  if (M >= (1ll << 31) || K >= (1ll << 31)) return 0;
  return 1;
}

int64_t b200st_gemm_lnbwd_blocks(int64_t M) { return (M + GL_BM - 1) / GL_BM; }

int b200st_gemm_lnbwd(int dtype, int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* W, int64_t ldw,
                      const void* X, const float* gamma, const float* mean, const float* rstd, const void* ADD, void* DX,
                      float* partials, b200st_stream_t stream) {
  if (!b200st_gemm_lnbwd_eligible(dtype, M, N, K, A, lda, W, ldw, X, ADD, DX, gamma))
    return set_error("gemm_lnbwd: needs bf16, N = 512, K %% 8 == 0, 16-byte aligned operands (got M=%lld N=%lld K=%lld)",
                     (long long)M, (long long)N, (long long)K);
  if (!X || !gamma || !mean || !rstd || !DX || !partials) return set_error("gemm_lnbwd: X, gamma, mean, rstd, DX and partials are required");
  CUtensorMap ma, mb;
  if (make_map(&ma, A, M, K, lda, GL_BM)) return -1;
  if (make_map(&mb, W, K, N, ldw, 64)) return -1;              // [K, N] row-major: boxes of 64 k-rows x 64 columns
  B200ST_CUDA(cudaFuncSetAttribute((const void*)gemm_lnbwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GL_SMEM));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(GL_CL, (unsigned)((M + GL_BM - 1) / GL_BM), 1);
  cfg.blockDim = dim3(GL_THREADS);
  cfg.dynamicSmemBytes = GL_SMEM;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  B200ST_CUDA(cudaLaunchKernelEx(&cfg, gemm_lnbwd_kernel, ma, mb, (const __nv_bfloat16*)X, gamma, mean, rstd,
                                 (const __nv_bfloat16*)ADD, (__nv_bfloat16*)DX, partials, (int)M, (int)K));
  B200ST_LAUNCH_CHECK("gemm_lnbwd");
  return 0;
}

}  // extern "C"
