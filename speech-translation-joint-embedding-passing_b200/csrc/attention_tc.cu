// Scaled-dot-product attention core on the tcgen05 tensor cores (bf16 operands, fp32 accumulation in TMEM) for the
// Transformer shapes of this path: head dim 64, at most 64 queries and 64 keys per (batch, head).
// Replaces ScaledDotProductAttention.forward + autograd (reference modules/layers.py:213-229).
//
// One CTA = one (batch, head).  Every operand is ONE 64 x 64 bf16 tile in shared memory, "row = token, 128 bytes =
// 64 contiguous values, 16-byte chunks XOR-swizzled by (row & 7)" -- the canonical SWIZZLE_128B layout.  The same
// bytes serve as a K-major operand (rows are the M/N index, the 128 B run along K) or as an MN-major operand (rows
// are the K index, the 128 B run along M/N); only the UMMA descriptors differ, so no transposed copy is ever made:
//
//   forward    S  = Q K^T            A = Q  K-major     B = K  K-major      -> TMEM cols [0,64)
//              O  = P V              A = P  K-major     B = V  MN-major     -> TMEM cols [64,128)
//   backward   dP = dO V^T           A = dO K-major     B = V  K-major      -> cols [0,64)
//              dV = P^T dO           A = P  MN-major    B = dO MN-major     -> cols [64,128)
//              dQ = dS K / temp      A = dS K-major     B = K  MN-major     -> cols [0,64)   (dP already consumed)
//              dK = dS^T Q / temp    A = dS MN-major    B = Q  MN-major     -> cols [64,128) (dV already stored)
//
// M = 64 accumulators: row i lives in TMEM lane (i / 16) * 32 + (i % 16), so warp w owns rows 16w .. 16w+15 in its
// lanes 0..15; such a thread holds a whole score row in registers and the softmax / dS row reductions need no
// shuffles.  128 TMEM columns and 32-48 KB of shared memory per CTA: four CTAs share an SM.
#include "umma.cuh"

namespace b200st {

constexpr int AT_THREADS = 128;
constexpr int AT_TILE = 64 * 128;          // bytes of one 64 x 64 bf16 tile

__device__ __forceinline__ uint32_t at_swz(uint32_t row, uint32_t chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }
__device__ __forceinline__ void at_sts_v4(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 at_lds_v4(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void at_sts_b16(uint32_t a, unsigned short v) { asm volatile("st.shared.b16 [%0], %1;" ::"r"(a), "h"(v) : "memory"); }
__device__ __forceinline__ uint32_t at_pack(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ void at_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// rows [0, L) of a [L x 64] bf16 matrix (row stride ld elements, 16-byte aligned rows) -> swizzled tile; rows >= L zero
__device__ __forceinline__ void at_load_tile(uint32_t tile, const __nv_bfloat16* __restrict__ src, int64_t ld, int L) {
  for (int i = threadIdx.x; i < 64 * 8; i += AT_THREADS) {
    const int r = i >> 3, c = i & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < L) v = *reinterpret_cast<const uint4*>(src + (int64_t)r * ld + c * 8);
    at_sts_v4(tile + at_swz(r, c), v);
  }
}
// The same, asynchronously (cp.async, 16 B each; rows >= L zero-filled through src-size 0): every tile of a kernel is
// requested before anything waits, so the CTA pays ONE global-memory latency for all its operands instead of one per tile
// (ncu, round 2: the three load -> st.shared phases of the synchronous version held 31 % of the forward kernel's stall samples).
__device__ __forceinline__ void at_load_tile_async(uint32_t tile, const __nv_bfloat16* __restrict__ src, int64_t ld, int L) {
  for (int i = threadIdx.x; i < 64 * 8; i += AT_THREADS) {
    const int r = i >> 3, c = i & 7;
    const bool ok = r < L;
    const int n = ok ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(tile + at_swz(r, c)), "l"(src + (int64_t)(ok ? r : 0) * ld + c * 8), "r"(n) : "memory");
  }
}
__device__ __forceinline__ void at_cp_wait_all() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}
// 64 fp32 accumulator columns [col0, col0+64) of this thread's TMEM lane
__device__ __forceinline__ void at_tmem_row(uint32_t taddr, float* v) {
  uint32_t r[32];
  tmem_ld32(taddr, r);
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  tmem_ld32(taddr + 32, r);
#pragma unroll
  for (int j = 0; j < 32; ++j) v[32 + j] = __uint_as_float(r[j]);
}
// one row of 64 fp32 values -> bf16 -> 128 B at dst (16-byte aligned)
__device__ __forceinline__ void at_store_row(__nv_bfloat16* dst, const float* v, float scale) {
#pragma unroll
  for (int c = 0; c < 8; ++c)
    *reinterpret_cast<uint4*>(dst + c * 8) =
        make_uint4(at_pack(v[8 * c] * scale, v[8 * c + 1] * scale), at_pack(v[8 * c + 2] * scale, v[8 * c + 3] * scale),
                   at_pack(v[8 * c + 4] * scale, v[8 * c + 5] * scale), at_pack(v[8 * c + 6] * scale, v[8 * c + 7] * scale));
}
// four K=16 steps of a 64 x 64 x 64 product
template <bool A_MN, bool B_MN>
__device__ __forceinline__ void at_mma64(uint32_t tmem_d, uint32_t sa, uint32_t sb) {
  constexpr uint32_t idesc = umma_idesc(64, 64, A_MN, B_MN);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint64_t da = A_MN ? umma_desc(sa + k * 2048, 8192, 1024) : umma_desc(sa + k * 32, 16, 1024);
    const uint64_t db = B_MN ? umma_desc(sb + k * 2048, 8192, 1024) : umma_desc(sb + k * 32, 16, 1024);
    tc_mma_f16(tmem_d, da, db, idesc, k > 0 ? 1u : 0u);
  }
}

struct AtCommon {
  uint8_t* smem;
  uint64_t* bar;
  uint32_t tmem_base;
};
__device__ __forceinline__ AtCommon at_prologue(uint8_t* smem_raw, int n_tiles) {
  AtCommon c;
  c.smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  c.bar = (uint64_t*)(c.smem + n_tiles * AT_TILE);
  uint32_t* tmem_slot = (uint32_t*)(c.bar + 2);
  if (threadIdx.x == 0) {
    mbar_init(&c.bar[0], 1);
    mbar_init(&c.bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  c.tmem_base = *tmem_slot;
  return c;
}
__device__ __forceinline__ void at_epilogue(const AtCommon& c) {
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(c.tmem_base), "r"(128u) : "memory");
  }
}

__global__ void __launch_bounds__(AT_THREADS, 4)
mha_fwd_tc_kernel(const __nv_bfloat16* __restrict__ q, int64_t ldq, const __nv_bfloat16* __restrict__ k, int64_t ldk,
                  const __nv_bfloat16* __restrict__ v, int64_t ldv, const uint8_t* __restrict__ mask, int64_t mask_sb,
                  int64_t mask_sq, __nv_bfloat16* __restrict__ o, int64_t ldo, __nv_bfloat16* __restrict__ p, int H,
                  int Lq, int Lk, float temperature) {
  extern __shared__ uint8_t smem_raw[];
  const AtCommon c = at_prologue(smem_raw, 4);
  const uint32_t sQ = smem_u32(c.smem), sK = sQ + AT_TILE, sV = sK + AT_TILE, sP = sV + AT_TILE;
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_wait();
  pdl_launch_dependents();
  at_load_tile_async(sQ, q + (int64_t)b * Lq * ldq + h * 64, ldq, Lq);
  at_load_tile_async(sK, k + (int64_t)b * Lk * ldk + h * 64, ldk, Lk);
  at_load_tile_async(sV, v + (int64_t)b * Lk * ldv + h * 64, ldv, Lk);
  at_cp_wait_all();
  at_fence_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    tc_fence_after();
    at_mma64<false, false>(c.tmem_base, sQ, sK);                    // S = Q K^T
    tc_commit(&c.bar[0]);
  }
  mbar_wait(&c.bar[0], 0);
  tc_fence_after();
  const int row = warp * 16 + lane;                                   // valid for lane < 16
  const uint32_t trow = c.tmem_base + ((uint32_t)(warp * 32) << 16);
  {
    float s[64];
    at_tmem_row(trow, s);                                             // warp-collective
    if (lane < 16) {
      if (row < Lq) {
        const uint8_t* mr = mask ? mask + b * mask_sb + row * mask_sq : nullptr;
        // scores are kept in units of log2(e): softmax(x) = 2^(x' - max x') with x' = x * log2e / temperature, one FMUL in front
        // and FADD + MUFU.EX2 per element (a true division and __expf's extra multiply per element were a third of the
        // instructions of this phase, which only half of the lanes of a warp can execute: M = 64 rows sit in lanes 0..15)
        const float sc = kLog2e / temperature;                        // layers.py:216
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 64; ++j) {
          float sv = s[j] * sc;
          if (j < Lk && mr && mr[j] == 0) sv = -1e9f * kLog2e;        // layers.py:224
          if (j >= Lk) sv = -INFINITY;
          s[j] = sv;
          mx = fmaxf(mx, sv);
        }
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 64; ++j) { s[j] = ex2_approx(s[j] - mx); sum += s[j]; }
        const float inv = 1.f / sum;
#pragma unroll
        for (int j = 0; j < 64; ++j) s[j] *= inv;
        if (p) {
          __nv_bfloat16* pr = p + (((int64_t)b * H + h) * Lq + row) * Lk;
          if ((Lk & 1) == 0) {
#pragma unroll
            for (int j = 0; j < 64; j += 2)
              if (j < Lk) *reinterpret_cast<uint32_t*>(pr + j) = at_pack(s[j], s[j + 1]);
          } else {
#pragma unroll
            for (int j = 0; j < 64; ++j)
              if (j < Lk) pr[j] = __float2bfloat16_rn(s[j]);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 64; ++j) s[j] = 0.f;
      }
#pragma unroll
      for (int ch = 0; ch < 8; ++ch)
        at_sts_v4(sP + at_swz(row, ch), make_uint4(at_pack(s[8 * ch], s[8 * ch + 1]), at_pack(s[8 * ch + 2], s[8 * ch + 3]),
                                                   at_pack(s[8 * ch + 4], s[8 * ch + 5]), at_pack(s[8 * ch + 6], s[8 * ch + 7])));
    }
  }
  at_fence_async();
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    tc_fence_after();
    at_mma64<false, true>(c.tmem_base + 64, sP, sV);                 // O = P V
    tc_commit(&c.bar[1]);
  }
  mbar_wait(&c.bar[1], 0);
  tc_fence_after();
  {
    float acc[64];
    at_tmem_row(trow + 64, acc);
    if (lane < 16 && row < Lq) at_store_row(o + ((int64_t)b * Lq + row) * ldo + h * 64, acc, 1.f);
  }
  at_epilogue(c);
}

__global__ void __launch_bounds__(AT_THREADS, 4)
mha_bwd_tc_kernel(const __nv_bfloat16* __restrict__ dout, int64_t ldo, const __nv_bfloat16* __restrict__ q, int64_t ldq,
                  const __nv_bfloat16* __restrict__ k, int64_t ldk, const __nv_bfloat16* __restrict__ v, int64_t ldv,
                  const __nv_bfloat16* __restrict__ p, __nv_bfloat16* __restrict__ dq, int64_t lddq,
                  __nv_bfloat16* __restrict__ dk, int64_t lddk, __nv_bfloat16* __restrict__ dv, int64_t lddv, int H,
                  int Lq, int Lk, float temperature) {
  extern __shared__ uint8_t smem_raw[];
  const AtCommon c = at_prologue(smem_raw, 6);
  const uint32_t sQ = smem_u32(c.smem), sK = sQ + AT_TILE, sV = sK + AT_TILE, sdO = sV + AT_TILE, sP = sdO + AT_TILE,
                 sdS = sP + AT_TILE;
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_wait();
  pdl_launch_dependents();
  at_load_tile_async(sQ, q + (int64_t)b * Lq * ldq + h * 64, ldq, Lq);
  at_load_tile_async(sK, k + (int64_t)b * Lk * ldk + h * 64, ldk, Lk);
  at_load_tile_async(sV, v + (int64_t)b * Lk * ldv + h * 64, ldv, Lk);
  at_load_tile_async(sdO, dout + (int64_t)b * Lq * ldo + h * 64, ldo, Lq);
  // saved probabilities: dense [Lq][Lk] bf16 -> zero-padded swizzled tile
  for (int i = threadIdx.x; i < 64 * 8; i += AT_THREADS) at_sts_v4(sP + i * 16, make_uint4(0, 0, 0, 0));
  __syncthreads();
  {
    const __nv_bfloat16* pb = p + ((int64_t)b * H + h) * Lq * Lk;
    if ((Lk & 1) == 0 && ((uintptr_t)pb & 3) == 0) {      // pairs of probabilities: 4-byte cp.async, all in flight at once
      const int half = Lk >> 1;
      for (int idx = threadIdx.x; idx < Lq * half; idx += AT_THREADS) {
        const int i = idx / half, j = 2 * (idx - i * half);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sP + at_swz(i, j >> 3) + (j & 7) * 2), "l"(pb + (int64_t)i * Lk + j) : "memory");
      }
    } else {
      const unsigned short* pu = reinterpret_cast<const unsigned short*>(pb);
      for (int idx = threadIdx.x; idx < Lq * Lk; idx += AT_THREADS) {
        const int i = idx / Lk, j = idx - i * Lk;
        at_sts_b16(sP + at_swz(i, j >> 3) + (j & 7) * 2, pu[idx]);
      }
    }
  }
  at_cp_wait_all();
  at_fence_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    tc_fence_after();
    at_mma64<false, false>(c.tmem_base, sdO, sV);                    // dP = dO V^T
    at_mma64<true, true>(c.tmem_base + 64, sP, sdO);                 // dV = P^T dO
    tc_commit(&c.bar[0]);
  }
  mbar_wait(&c.bar[0], 0);
  tc_fence_after();
  const int row = warp * 16 + lane;                                   // valid for lane < 16
  const uint32_t trow = c.tmem_base + ((uint32_t)(warp * 32) << 16);
  {
    float dp[64];
    at_tmem_row(trow, dp);
    if (lane < 16) {
      // dS = P * (dP - rowsum(P dP)); rows >= Lq and columns >= Lk have P = 0.  P is re-read from the tile in both
      // passes (16-byte shared loads) instead of being held in another 64 registers.
      float delta = 0.f;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        const uint4 u = at_lds_v4(sP + at_swz(row, ch));
        const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          delta = fmaf(__uint_as_float(w4[e] << 16), dp[8 * ch + 2 * e], delta);
          delta = fmaf(__uint_as_float(w4[e] & 0xffff0000u), dp[8 * ch + 2 * e + 1], delta);
        }
      }
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        const uint4 u = at_lds_v4(sP + at_swz(row, ch));
        const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
        uint32_t o4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
          o4[e] = at_pack(__uint_as_float(w4[e] << 16) * (dp[8 * ch + 2 * e] - delta),
                          __uint_as_float(w4[e] & 0xffff0000u) * (dp[8 * ch + 2 * e + 1] - delta));
        at_sts_v4(sdS + at_swz(row, ch), make_uint4(o4[0], o4[1], o4[2], o4[3]));
      }
    }
  }
  {
    float acc[64];
    at_tmem_row(trow + 64, acc);                                      // dV rows = keys
    if (lane < 16 && row < Lk) at_store_row(dv + ((int64_t)b * Lk + row) * lddv + h * 64, acc, 1.f);
  }
  at_fence_async();
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    tc_fence_after();
    at_mma64<false, true>(c.tmem_base, sdS, sK);                     // dQ = dS K
    at_mma64<true, true>(c.tmem_base + 64, sdS, sQ);                 // dK = dS^T Q
    tc_commit(&c.bar[1]);
  }
  mbar_wait(&c.bar[1], 0);
  tc_fence_after();
  const float inv_t = 1.f / temperature;
  {
    float acc[64];
    at_tmem_row(trow, acc);
    if (lane < 16 && row < Lq) at_store_row(dq + ((int64_t)b * Lq + row) * lddq + h * 64, acc, inv_t);
    at_tmem_row(trow + 64, acc);
    if (lane < 16 && row < Lk) at_store_row(dk + ((int64_t)b * Lk + row) * lddk + h * 64, acc, inv_t);
  }
  at_epilogue(c);
}

static bool at_aligned(const void* p, int64_t ld) { return ((uintptr_t)p & 15) == 0 && ld % 8 == 0; }

// returns 1 when the shape is not served by the tensor-core kernels (caller falls back to the SIMT tiles), 0 when launched
int mha_fwd_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const uint8_t* mask,
               int64_t mask_sb, int64_t mask_sq, void* o, int64_t ldo, void* p, int64_t B, int64_t H, int64_t Lq,
               int64_t Lk, int64_t d, float temperature, cudaStream_t st) {
  if (d != 64 || Lq > 64 || Lk > 64 || H > 65535 || B > 65535) return 1;
  if (!at_aligned(q, ldq) || !at_aligned(k, ldk) || !at_aligned(v, ldv) || !at_aligned(o, ldo)) return 1;
  if (p && ((uintptr_t)p & 3)) return 1;
  const size_t smem = 4 * AT_TILE + 64 + 1024;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute((const void*)mha_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return set_error("mha_fwd_tc: cannot reserve %zu B of shared memory", smem);
    attr = true;
  }
  B200ST_CUDA(launch_pdl(mha_fwd_tc_kernel, dim3((unsigned)H, (unsigned)B), dim3(AT_THREADS), smem, st,
                         (const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)k, ldk, (const __nv_bfloat16*)v, ldv, mask,
                         mask_sb, mask_sq, (__nv_bfloat16*)o, ldo, (__nv_bfloat16*)p, (int)H, (int)Lq, (int)Lk, temperature));
  B200ST_LAUNCH_CHECK("mha_fwd_tc");
  return 0;
}

int mha_bwd_tc(const void* dout, int64_t ldo, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
               int64_t ldv, const void* p, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
               int64_t B, int64_t H, int64_t Lq, int64_t Lk, int64_t d, float temperature, cudaStream_t st) {
  if (d != 64 || Lq > 64 || Lk > 64 || H > 65535 || B > 65535) return 1;
  if (!at_aligned(dout, ldo) || !at_aligned(q, ldq) || !at_aligned(k, ldk) || !at_aligned(v, ldv) ||
      !at_aligned(dq, lddq) || !at_aligned(dk, lddk) || !at_aligned(dv, lddv))
    return 1;
  const size_t smem = 6 * AT_TILE + 64 + 1024;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute((const void*)mha_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return set_error("mha_bwd_tc: cannot reserve %zu B of shared memory", smem);
    attr = true;
  }
  B200ST_CUDA(launch_pdl(mha_bwd_tc_kernel, dim3((unsigned)H, (unsigned)B), dim3(AT_THREADS), smem, st,
                         (const __nv_bfloat16*)dout, ldo, (const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)k, ldk,
                         (const __nv_bfloat16*)v, ldv, (const __nv_bfloat16*)p, (__nv_bfloat16*)dq, lddq,
                         (__nv_bfloat16*)dk, lddk, (__nv_bfloat16*)dv, lddv, (int)H, (int)Lq, (int)Lk, temperature));
  B200ST_LAUNCH_CHECK("mha_bwd_tc");
  return 0;
}

}  // namespace b200st
