// Persistent bidirectional LSTM recurrence on the tcgen05 tensor cores (bf16 operands, fp32 accumulate/state).
// Specialised for the reference's acoustic encoder: H = 256 hidden units per direction (Seq2seq.py:57).
//
// Layout of the work (forward):
//   grid = (8, ceil(B/16), 2): one thread-block CLUSTER of 8 CTAs per (direction, group of 16 sequences); 4 epilogue warps
//   + 1 issuer warp per CTA.  CTA `rank` owns hidden units [32*rank, 32*rank+32): the 128 matching rows of W_hh (4 gates x
//   32 units) are converted to bf16 once and stay resident in TENSOR MEMORY as the A operand of TS-form tcgen05.mma
//   (A from TMEM, B = h_{t-1} tile from shared memory: an SS-form N=16 MMA spends ~110 cycles streaming the 4 KB A tile
//   out of shared memory, the TS form ~22).
//   Every time step:  D[128 gate rows, 16 seqs] = W_slice[128,256] . h_{t-1}^T  (16 MMAs of K=16 into 4 independent TMEM
//   accumulators, one commit) -> tcgen05.ld -> + x-projection -> sigmoid/tanh (tanh.approx) -> smem transpose -> cell update
//   (c in registers) -> the CTA's 32x16 slice of h_t goes as bf16 straight into the NEXT-STEP B operand of all 8 CTAs:
//   16-byte `st.async ... mbarrier::complete_tx` stores into the peers' swizzled K-major tile; data and signal travel together,
//   there is no cluster barrier in the loop, and the issuer's wait on the local `hfull` mbarrier (8 KB expected) is the only
//   synchronisation.  Global stores of the saved state and the two-step-ahead x-projection prefetch sit behind the send.
//
// Backward keeps W_hh^T[256 units, own 128 gate rows] resident in TMEM instead: each CTA multiplies its own gate gradients
// (no all-gather) into partial dh for all 256 units, and the bf16 partials are reduce-scattered to the owning CTAs through
// distributed shared memory with the same st.async + complete_tx mechanism.
//
// What bounds a step (profiles/r02_blstm_experiments.txt; ~1880 cycles): MMA issue + completion ~630, epilogue ~830, and the
// h exchange: every CTA must receive the other 7 slices = 7 KB per step, and the SM-to-SM network moves ~21 B/clk per SM
// (B300_MICROARCH.md: "DSMEM BW 17-21 B/cyc, producer pays") -> ~340 cycles of pure transfer + ~215 latency that nothing in
// this decomposition can overlap.  Two restructurings were built and measured in round 2 and are NOT kept:
//   * 8 epilogue warps (two per TMEM lane quarter, 8 sequence columns each; every thread sends to 2 peers instead of 8):
//     tcgen05.ld + activation phase 450 -> 340 cycles, but the scattered sends land later (470 -> 800): 2370 cycles/step;
//   * per-source operand barriers (the two MMAs of source r issued as soon as r's 1 KB lands): each extra wait costs a
//     ~90-cycle mbarrier.try_wait + a 33-cycle proxy fence on the single issuer thread: 2620 cycles/step.
#include "common.cuh"

namespace b200st {

constexpr int RT_H = 256;        // hidden units per direction
constexpr int RT_C = 8;          // CTAs per cluster
constexpr int RT_UPC = 32;       // units per CTA
constexpr int RT_NB = 16;        // sequences per cluster (= UMMA N)
constexpr int RT_THREADS = 128;       // 4 epilogue warps (one TMEM lane quarter each)
constexpr int RT_BLOCK = 160;         // + 1 issuer warp: waits for operands, issues tcgen05.mma, arms barriers

// Optional in-kernel timeline (debug aid, b200st_debug_timeline): when a buffer is registered, thread 0 of the first
// CTA records clock64() at fixed points of time steps 64..71.  Costs one predictable branch per point otherwise.
__device__ long long* g_rt_timeline = nullptr;
#define RT_TL(i) do { if (tl_on && s >= 64 && s < 72) tl[(s - 64) * 16 + (i)] = clock64(); } while (0)

__device__ __forceinline__ uint32_t rt_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t rt_cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t rt_mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void rt_st_cluster_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// Remote 16-byte store whose completion is counted (complete_tx) on an mbarrier of the destination CTA: data and
// signal travel together, and unlike barrier.cluster.arrive.release nothing waits for unrelated global stores.
__device__ __forceinline__ void rt_st_async_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d), "r"(bar) : "memory");
}
// explicit shared-space accesses: pointers derived from the aligned dynamic-smem base lose their address space and
// would compile to generic LD/ST
__device__ __forceinline__ void rt_sts_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void rt_sts_b16(uint32_t a, unsigned short v) { asm volatile("st.shared.b16 [%0], %1;" ::"r"(a), "h"(v) : "memory"); }
__device__ __forceinline__ uint2 rt_lds_v2u(uint32_t a) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint4 rt_lds_v4u(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ float4 rt_lds_v4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
// volatile (position-pinned) read-only global loads / register moves for the two-step-ahead prefetch pipeline
__device__ __forceinline__ float rt_ldg_f32(const float* p) { float v; asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v; }
__device__ __forceinline__ unsigned short rt_ldg_u16(const unsigned short* p) { unsigned short v; asm volatile("ld.global.nc.u16 %0, [%1];" : "=h"(v) : "l"(p)); return v; }
__device__ __forceinline__ float rt_mov_f32(float x) { float v; asm volatile("mov.f32 %0, %1;" : "=f"(v) : "f"(x)); return v; }
__device__ __forceinline__ unsigned short rt_mov_u16(unsigned short x) { unsigned short v; asm volatile("mov.b16 %0, %1;" : "=h"(v) : "h"(x)); return v; }
__device__ __forceinline__ void rt_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rt_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rt_cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void rt_cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void rt_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void rt_tc_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void rt_tc_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void rt_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rt_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void rt_mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = rt_smem_u32(bar);
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (spin > (1u << 22)) __trap();
  }
}
__device__ __forceinline__ void rt_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(rt_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void rt_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
// Issuer-warp variants: executed by ALL 32 lanes of a convergent warp, the instruction itself is predicated on the
// elected lane.  Keeping the control flow warp-uniform lets the compiler hold descriptors / TMEM addresses in uniform
// registers (UTCHMMA operands); a divergent `if (lane == 0)` region forces a vector->uniform move before every MMA.
__device__ __forceinline__ uint32_t rt_elect() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void rt_mma_if(uint32_t leader, uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(leader) : "memory");
}
// TS form: A operand read from tensor memory (lane = row, 32-bit column = two consecutive K elements), B from smem.
// With N = 16 the SS form is bound by streaming the 4 KB A tile out of shared memory (~110 cycles per MMA measured);
// a TMEM-resident A removes that, and the recurrent weights never change during the sequence.
__device__ __forceinline__ void rt_mma_ts_if(uint32_t leader, uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc), "r"(leader) : "memory");
}
__device__ __forceinline__ void rt_tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void rt_tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
constexpr uint32_t RT_D_COLS = 64;      // accumulators live in TMEM columns [0, 64)
constexpr uint32_t RT_A_COL = 64;       // resident A operand (recurrent weights) in columns [64, 192)
constexpr uint32_t RT_TMEM_COLS = 256;

__device__ __forceinline__ void rt_commit_if(uint32_t leader, uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(rt_smem_u32(bar)), "r"(leader) : "memory");
}
__device__ __forceinline__ void rt_expect_tx_if(uint32_t leader, uint64_t* bar, uint32_t bytes) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t"
      "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}"
      ::"r"(rt_smem_u32(bar)), "r"(bytes), "r"(leader) : "memory");
}
__device__ __forceinline__ void rt_tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void rt_tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void rt_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// Sum of NACC independent accumulators (16 columns each, NACC*16 apart... consecutive) -> v[16].
// The K loop is spread round-robin over NACC accumulators so that consecutive tcgen05.mma do not form one long
// dependent accumulate chain; the partial sums are added here.
template <int NACC>
__device__ __forceinline__ void rt_tmem_ld_sum(uint32_t taddr, float* v) {
  uint32_t r[NACC][16];
#pragma unroll
  for (int a = 0; a < NACC; ++a) rt_tmem_ld16_nowait(taddr + a * 16, r[a]);
  rt_tmem_wait_ld();
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float s = __uint_as_float(r[0][i]);
#pragma unroll
    for (int a = 1; a < NACC; ++a) s += __uint_as_float(r[a][i]);
    v[i] = s;
  }
}
// K-major, 128B-swizzled operand descriptor: SBO = 8 rows * 128 B (see gemm_tc.cu / mma_sm100_desc.hpp)
__device__ __forceinline__ uint64_t rt_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
constexpr uint32_t RT_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(RT_NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__device__ __forceinline__ float rt_tanh(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rt_sigmoid(float x) { return fmaf(rt_tanh(0.5f * x), 0.5f, 0.5f); }
__device__ __forceinline__ uint32_t rt_pack(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}
// byte offset of the 16-byte chunk holding k..k+7 (k % 8 == 0) of row `row` inside one [rows x 64 k] SW128 tile
__device__ __forceinline__ uint32_t rt_swz(uint32_t row, uint32_t k_in_tile) {
  return row * 128 + ((((k_in_tile >> 3) ^ (row & 7))) << 4);
}

// Issue the 16 forward MMAs of one time step.  CONST_BASE >= 0 bakes the TMEM accumulator addresses in as
// immediates: UTCHMMA takes its TMEM address from a *uniform* register, and an address that was loaded from shared
// memory costs an ELECT + R2UR.BROADCAST round trip in front of every MMA (measured: ~116 cycles per MMA issue
// instead of ~35).  The allocator returns column 0 for the first allocation on an SM, which is the common case.
template <int CONST_BASE>
__device__ __forceinline__ void rt_issue_fwd(uint32_t leader, uint32_t tmem_base, uint32_t hb) {
  const uint32_t base = CONST_BASE >= 0 ? (uint32_t)CONST_BASE : tmem_base;
#pragma unroll
  for (int ks = 0; ks < 16; ++ks)       // A: W slice [128 rows x 256 k] = 128 columns, 8 columns per K=16 step
    rt_mma_ts_if(leader, base + (ks & 3) * 16, base + RT_A_COL + ks * 8,
                 rt_desc(hb + (ks >> 2) * 2048 + (ks & 3) * 32), RT_IDESC, ks >= 4 ? 1u : 0u);
}
template <int CONST_BASE>
__device__ __forceinline__ void rt_issue_bwd(uint32_t leader, uint32_t tmem_base, uint32_t bb) {
  const uint32_t base = CONST_BASE >= 0 ? (uint32_t)CONST_BASE : tmem_base;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)        // A: W_hh^T tile mt [128 units x 128 own gate rows] = 64 columns
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)
      rt_mma_ts_if(leader, base + (mt * 2 + (ks & 1)) * RT_NB, base + RT_A_COL + mt * 64 + ks * 8,
                   rt_desc(bb + (ks >> 2) * 2048 + (ks & 3) * 32), RT_IDESC, ks >= 2 ? 1u : 0u);
}

// shared memory map (forward): hbuf [2][4 kb][16 rows][128 B] 16 KB | actbuf float [4 gates][16 seqs][32 units] 8 KB |
// barriers + tmem slot.  (The W_hh slice lives in TMEM.)  The request is padded above half an SM's shared memory so
// that exactly one CTA is resident per SM and the TMEM allocation starts at column 0 (constant-address fast path).
constexpr int FWD_H_OFF = 0, FWD_ACT_OFF = FWD_H_OFF + 16384, FWD_BAR_OFF = FWD_ACT_OFF + 8192;
constexpr int FWD_SMEM = 120 * 1024;

__global__ void __cluster_dims__(RT_C, 1, 1) __launch_bounds__(RT_BLOCK, 1)
blstm_fwd_tc_kernel(const __nv_bfloat16* __restrict__ xproj, const float* __restrict__ w_hh_f,
                    const float* __restrict__ w_hh_r, const int32_t* __restrict__ lens,
                    __nv_bfloat16* __restrict__ out, int64_t out_ld_t, int64_t out_ld_b, int pair,
                    __nv_bfloat16* __restrict__ hs, float* __restrict__ acts, float* __restrict__ cs, int Tn, int B) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* hbuf = smem + FWD_H_OFF;
  float* actbuf = (float*)(smem + FWD_ACT_OFF);
  uint64_t* mma_bar = (uint64_t*)(smem + FWD_BAR_OFF);
  uint64_t* hfull = mma_bar + 1;                 // [2]: h buffer b holds a complete h_t (8 x 1 KB st.async landed)
  uint32_t* tmem_slot = (uint32_t*)(smem + FWD_BAR_OFF + 32);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = rt_cluster_rank();
  const int grp = blockIdx.y, dir = blockIdx.z;
  const float* w = dir ? w_hh_r : w_hh_f;

  for (int i = tid; i < 16384 / 16; i += RT_BLOCK) reinterpret_cast<uint4*>(hbuf)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    rt_mbar_init(mma_bar, 1);
    rt_mbar_init(&hfull[0], 1);
    rt_mbar_init(&hfull[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    rt_mbar_expect_tx(&hfull[1], 8192);        // step 0 fills buffer 1
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(rt_smem_u32(tmem_slot)), "r"(RT_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  rt_fence_async();
  rt_tc_before();
  __syncthreads();
  rt_tc_after();
  const uint32_t tmem_base = *tmem_slot;
  // ---- resident weights -> TMEM: thread (warp w < 4, lane l) owns TMEM lane lr = 32 w + l = gate*32 + ul, i.e. W_hh
  // row gate*256 + 32*rank + ul, packed two K elements per 32-bit column (K-major A operand of the TS-form MMA)
  if (warp < 4) {
    const float* wrow = w + (size_t)(warp * RT_H + rank * RT_UPC + lane) * RT_H;
#pragma unroll 1
    for (int c = 0; c < 8; ++c) {
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 a = *reinterpret_cast<const float4*>(wrow + c * 32 + j * 8);
        const float4 b = *reinterpret_cast<const float4*>(wrow + c * 32 + j * 8 + 4);
        pk[4 * j] = rt_pack(a.x, a.y); pk[4 * j + 1] = rt_pack(a.z, a.w);
        pk[4 * j + 2] = rt_pack(b.x, b.y); pk[4 * j + 3] = rt_pack(b.z, b.w);
      }
      rt_tmem_st16(tmem_base + ((uint32_t)(warp * 32) << 16) + RT_A_COL + c * 16, pk);
    }
    rt_tmem_wait_st();
  }
  rt_tc_before();
  __syncthreads();
  rt_tc_after();
  rt_cluster_arrive();      // every CTA has initialised its barriers and zeroed its h buffers
  rt_cluster_wait();

  long long* tl = g_rt_timeline;
  if (warp == 4) {
    // ================= issuer warp (convergent; the elected lane issues) =================
    const uint32_t leader = rt_elect();
    const bool tl_on = tl != nullptr && leader && rank == 0 && blockIdx.y == 0 && blockIdx.z == 0;
    const uint32_t hb0 = rt_smem_u32(hbuf);
    uint32_t hph[2] = {0u, 0u};
    int cur = 0;
    uint32_t phase = 0;
    for (int s = 0; s < Tn; ++s) {
      RT_TL(0);
      if (s > 0) {                             // h_{t-1} slices of all 8 CTAs have landed in hbuf[cur]
        rt_mbar_wait(&hfull[cur], hph[cur]);
        hph[cur] ^= 1;
      }
      RT_TL(1);
      rt_fence_async();
      rt_tc_after();
      RT_TL(10);
      const uint32_t hb = hb0 + cur * 8192;
      if (tmem_base == 0) rt_issue_fwd<0>(leader, 0, hb);
      else rt_issue_fwd<-1>(leader, tmem_base, hb);
      RT_TL(11);
      rt_commit_if(leader, mma_bar);
      RT_TL(2);
      // arm hbuf[cur] for h_{t+1}: its refill cannot start before every CTA has consumed this step's outputs
      if (s + 2 < Tn) rt_expect_tx_if(leader, &hfull[cur], 8192);
      rt_mbar_wait(mma_bar, phase);            // keep at most one step of MMAs in flight
      phase ^= 1;
      RT_TL(3);
      cur ^= 1;
    }
  } else {
  // ================= epilogue warps =================
  // ---- per-thread roles
  // (1) gate phase: thread owns gate row lr = tid (gate = warp, unit = lane) for all 16 sequences
  const int grow_g = warp * RT_H + rank * RT_UPC + lane;
  // (2) cell phase: thread owns sequence cb = tid / 8 and units 4*ug .. 4*ug+3, ug = tid % 8
  const int cb = tid >> 3, ug = tid & 7;
  const int b0 = grp * RT_NB;
  const int bglob = b0 + cb;
  const bool b_ok = bglob < B;
  const int len_b = b_ok ? lens[bglob] : 0;
  const int ubase = rank * RT_UPC + 4 * ug;
  float c_st[4] = {0.f, 0.f, 0.f, 0.f}, h_st[4] = {0.f, 0.f, 0.f, 0.f};
  const uint32_t hbuf_u32 = rt_smem_u32(hbuf);
  const uint32_t hfull_u32 = rt_smem_u32(hfull);
  const uint32_t act_u32 = rt_smem_u32(actbuf);
  // destination chunk (16 B = 8 units) of this thread pair inside one h buffer
  const uint32_t k0 = rank * RT_UPC + 4 * (ug & ~1);
  const uint32_t h_chunk_off = (k0 >> 6) * 2048 + rt_swz(cb, k0 & 63);

  // x-projection prefetch: 16 independent loads issued back to back one step ahead, kept as raw bf16 bits and
  // converted only when consumed, so their DRAM latency never sits on the recurrent critical path.
  unsigned short xr[16], xq[16];     // consume set / staging set (loaded two steps ahead)
  const bool full_grp = b0 + RT_NB <= B;
  auto load_x = [&](int t) {
    const unsigned short* xp = reinterpret_cast<const unsigned short*>(xproj) + (((size_t)dir * Tn + t) * B) * (4 * RT_H) + grow_g;
    if (full_grp) {                 // one base address, compile-time row offsets
#pragma unroll
      for (int b = 0; b < RT_NB; ++b) xq[b] = rt_ldg_u16(xp + (size_t)(b0 + b) * (4 * RT_H));
    } else {
#pragma unroll
      for (int b = 0; b < RT_NB; ++b) xq[b] = rt_ldg_u16(xp + (size_t)min(b0 + b, B - 1) * (4 * RT_H));
    }
  };
  auto advance_x = [&]() {
#pragma unroll
    for (int b = 0; b < RT_NB; ++b) xr[b] = rt_mov_u16(xq[b]);
  };
  auto t_of = [&](int s) { return dir ? (Tn - 1 - s) : s; };
  if (Tn > 0) { load_x(t_of(0)); advance_x(); }
  if (Tn > 1) load_x(t_of(1));

  const bool tl_on = tl != nullptr && tid == 0 && rank == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  int cur = 0;
  uint32_t phase = 0;
  for (int s = 0; s < Tn; ++s) {
    const int t = dir ? (Tn - 1 - s) : s;
    rt_mbar_wait(mma_bar, phase);
    phase ^= 1;
    rt_tc_after();
    RT_TL(9);
    float g[16];
    rt_tmem_ld_sum<4>(tmem_base + ((uint32_t)(warp * 32) << 16), g);
    rt_tc_before();
    RT_TL(4);
    // gate non-linearity: warp 2 holds the candidate gate (tanh), the others sigmoid (PyTorch order i,f,g,o)
#pragma unroll
    for (int b = 0; b < RT_NB; ++b) g[b] += __uint_as_float((uint32_t)xr[b] << 16);
#pragma unroll
    for (int b = 0; b < RT_NB; ++b)
      rt_sts_f32(act_u32 + (uint32_t)(((warp * RT_NB + b) * RT_UPC + lane) * 4), (warp == 2) ? rt_tanh(g[b]) : rt_sigmoid(g[b]));
    RT_TL(5);
    asm volatile("bar.sync 1, 128;" ::: "memory");
    RT_TL(6);
    // ---- cell update for (sequence cb, units ubase..ubase+3)
    const bool valid = t < len_b;
    const float4 gi = rt_lds_v4(act_u32 + (uint32_t)(((0 * RT_NB + cb) * RT_UPC + 4 * ug) * 4));
    const float4 gf = rt_lds_v4(act_u32 + (uint32_t)(((1 * RT_NB + cb) * RT_UPC + 4 * ug) * 4));
    const float4 gg = rt_lds_v4(act_u32 + (uint32_t)(((2 * RT_NB + cb) * RT_UPC + 4 * ug) * 4));
    const float4 go = rt_lds_v4(act_u32 + (uint32_t)(((3 * RT_NB + cb) * RT_UPC + 4 * ug) * 4));
    const float ai[4] = {gi.x, gi.y, gi.z, gi.w}, af[4] = {gf.x, gf.y, gf.z, gf.w};
    const float ag[4] = {gg.x, gg.y, gg.z, gg.w}, ao[4] = {go.x, go.y, go.z, go.w};
    float ho[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (valid) {
        c_st[q] = fmaf(af[q], c_st[q], ai[q] * ag[q]);
        h_st[q] = ao[q] * rt_tanh(c_st[q]);
      }
      ho[q] = valid ? h_st[q] : 0.f;
    }
    // ---- scatter this pair's 8 units of h_t into the next-step B operand of every CTA in the cluster
    const uint32_t p0 = rt_pack(h_st[0], h_st[1]), p1 = rt_pack(h_st[2], h_st[3]);
    const uint32_t q0 = __shfl_down_sync(0xffffffffu, p0, 1), q1 = __shfl_down_sync(0xffffffffu, p1, 1);
    if ((ug & 1) == 0 && s + 1 < Tn) {
      const uint32_t dst = hbuf_u32 + (cur ^ 1) * 8192 + h_chunk_off;
      const uint32_t bar = hfull_u32 + (cur ^ 1) * 8;
#pragma unroll
      for (uint32_t i = 0; i < RT_C; ++i) {       // rotated destination order: no receiver is hit by all 8 senders at once
        const uint32_t r = (i + rank) & (RT_C - 1);
        rt_st_async_v4(rt_mapa(dst, r), p0, p1, q0, q1, rt_mapa(bar, r));
      }
    }
    RT_TL(7);
    // x-projection pipeline (position-pinned asm): step s+1's bits, loaded a full step ago, move to the consume set;
    // step s+2's loads are issued.  No scoreboard wait on DRAM latency can reach the recurrent critical path.
    advance_x();
    if (s + 2 < Tn) load_x(t_of(s + 2));
    // ---- global stores: nothing on the recurrent critical path waits for them
    if (b_ok) {
      const size_t row = ((size_t)dir * Tn + t) * B + bglob;
      if (acts) {
        float* a = acts + row * (4 * RT_H) + ubase;
        *reinterpret_cast<float4*>(a) = gi;
        *reinterpret_cast<float4*>(a + RT_H) = gf;
        *reinterpret_cast<float4*>(a + 2 * RT_H) = gg;
        *reinterpret_cast<float4*>(a + 3 * RT_H) = go;
      }
      if (cs) *reinterpret_cast<float4*>(cs + row * RT_H + ubase) = make_float4(c_st[0], c_st[1], c_st[2], c_st[3]);
      const uint2 hv = make_uint2(rt_pack(ho[0], ho[1]), rt_pack(ho[2], ho[3]));
      *reinterpret_cast<uint2*>(out + (size_t)(t / pair) * out_ld_t + (size_t)bglob * out_ld_b +
                                (size_t)(t % pair) * 2 * RT_H + dir * RT_H + ubase) = hv;
      if (hs) *reinterpret_cast<uint2*>(hs + (((size_t)dir * (Tn + 1) + (dir ? t : t + 1)) * B + bglob) * RT_H + ubase) = hv;
    }
    RT_TL(8);
    cur ^= 1;
  }
  }  // epilogue warps
  rt_tc_before();
  rt_cluster_arrive();       // nobody exits while a peer could still address its shared memory
  rt_cluster_wait();
  if (warp == 0) {
    rt_tc_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(RT_TMEM_COLS) : "memory");
  }
}

// shared memory map (backward): B = own dG [2 kb][16][128 B] 4 KB | red bf16 [2][8 src][2 quad pairs][32 units][8] 16 KB (region sized 32 KB) |
// barriers + tmem slot.  (W_hh^T lives in TMEM.)  Padded like the forward kernel: one CTA per SM.
constexpr int BWD_B_OFF = 0, BWD_RED_OFF = BWD_B_OFF + 4096, BWD_BAR_OFF = BWD_RED_OFF + 32768;
constexpr int BWD_SMEM = 120 * 1024;

__global__ void __cluster_dims__(RT_C, 1, 1) __launch_bounds__(RT_BLOCK, 1)
blstm_bwd_tc_kernel(const __nv_bfloat16* __restrict__ dout, int64_t out_ld_t, int64_t out_ld_b, int pair,
                    const float* __restrict__ acts, const float* __restrict__ cs, const float* __restrict__ w_hh_f,
                    const float* __restrict__ w_hh_r, const int32_t* __restrict__ lens,
                    __nv_bfloat16* __restrict__ dgates, int Tn, int B) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* Bsm = smem + BWD_B_OFF;
  float* red = (float*)(smem + BWD_RED_OFF);
  uint64_t* mma_bar = (uint64_t*)(smem + BWD_BAR_OFF);
  uint64_t* rfull = mma_bar + 1;                 // [2]: red buffer b holds all 8 partial blocks (8 x 2 KB)
  uint32_t* tmem_slot = (uint32_t*)(smem + BWD_BAR_OFF + 32);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = rt_cluster_rank();
  const int grp = blockIdx.y, dir = blockIdx.z;
  const float* w = dir ? w_hh_r : w_hh_f;

  for (int i = tid; i < (4096 + 32768) / 16; i += RT_BLOCK) reinterpret_cast<uint4*>(Bsm)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    rt_mbar_init(mma_bar, 1);
    rt_mbar_init(&rfull[0], 1);
    rt_mbar_init(&rfull[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    rt_mbar_expect_tx(&rfull[1], 8192);        // step 0 sends its partials into buffer 1
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(rt_smem_u32(tmem_slot)), "r"(RT_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  rt_fence_async();
  rt_tc_before();
  __syncthreads();
  rt_tc_after();
  const uint32_t tmem_base = *tmem_slot;
  // ---- resident W_hh^T -> TMEM: tile mt, lane = unit (mt*128 + 32 w + l), K = own gate rows kl = gate*32 + ul
  // (W_hh row gate*256 + 32*rank + ul), two K elements per 32-bit column.  Lanes run along units: coalesced reads.
  if (warp < 4) {
#pragma unroll 1
    for (int mt = 0; mt < 2; ++mt) {
      const int uu = mt * 128 + warp * 32 + lane;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {           // 32 kl per pass = gate c, ul 0..31
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int kl = c * 32 + 2 * j;
          const float v0 = w[(size_t)((kl >> 5) * RT_H + rank * RT_UPC + (kl & 31)) * RT_H + uu];
          const float v1 = w[(size_t)(((kl + 1) >> 5) * RT_H + rank * RT_UPC + ((kl + 1) & 31)) * RT_H + uu];
          pk[j] = rt_pack(v0, v1);
        }
        rt_tmem_st16(tmem_base + ((uint32_t)(warp * 32) << 16) + RT_A_COL + mt * 64 + c * 16, pk);
      }
    }
    rt_tmem_wait_st();
  }
  rt_tc_before();
  __syncthreads();
  rt_tc_after();
  rt_cluster_arrive();
  rt_cluster_wait();

  long long* tl = g_rt_timeline;
  if (warp == 4) {
    // ================= issuer warp (convergent; the elected lane issues) =================
    const uint32_t leader = rt_elect();
    const uint32_t bb = rt_smem_u32(Bsm);
    int cur = 0;
    for (int s = 0; s < Tn; ++s) {
      __syncthreads();                         // the epilogue warps have written this step's gate gradients (B operand)
      if (s + 2 < Tn) rt_expect_tx_if(leader, &rfull[cur], 8192);   // red[cur] fully read: arm it for step s+1's partials
      rt_fence_async();
      rt_tc_after();
      if (tmem_base == 0) rt_issue_bwd<0>(leader, 0, bb);
      else rt_issue_bwd<-1>(leader, tmem_base, bb);
      rt_commit_if(leader, mma_bar);
      cur ^= 1;
    }
  } else {
  // ================= epilogue warps =================
  // thread owns unit ul = lane of this CTA and sequences 4*warp .. 4*warp+3 in the pointwise phase
  const int ul = lane, u = rank * RT_UPC + ul;
  const int b0 = grp * RT_NB;
  int len_b[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) len_b[q] = (b0 + 4 * warp + q < B) ? lens[b0 + 4 * warp + q] : 0;
  float dcrec[4] = {0.f, 0.f, 0.f, 0.f};
  const uint32_t red_u32 = rt_smem_u32(red);
  const uint32_t rfull_u32 = rt_smem_u32(rfull);
  const uint32_t bsm_u32 = rt_smem_u32(Bsm);
  uint32_t rph[2] = {0u, 0u};
  const size_t G4 = 4 * RT_H;

  // software-pipelined loads of the saved forward state: unconditional (clamped indices), issued TWO steps ahead into
  // the staging set (qa..), moved into the consume set (pa..) one step later while the MMAs run -- no scoreboard wait
  // for DRAM latency ever lands on the recurrent critical path.
  float pa[4][4], pct[4], pcp[4], qa[4][4], qcp[4];
  unsigned short pdy[4], qdy[4];
  const bool full_grp = b0 + RT_NB <= B;
  auto prefetch = [&](int t) {
    const int tp = min(max(dir ? t + 1 : t - 1, 0), Tn - 1);
    // pair is 1 (plain layer) or 2 (pyramid frame-pair concat) in practice: no integer division on the hot path
    const int tq = pair == 1 ? t : (pair == 2 ? (t >> 1) : t / pair), tr = pair == 1 ? 0 : (pair == 2 ? (t & 1) : t % pair);
    const unsigned short* dyp = reinterpret_cast<const unsigned short*>(dout) + (size_t)tq * out_ld_t +
                                (size_t)tr * 2 * RT_H + dir * RT_H + u;
    if (full_grp) {                 // one base address per tensor, compile-time offsets
      const size_t row = ((size_t)dir * Tn + t) * B + b0 + 4 * warp;
      const float* a = acts + row * G4 + u;
      const float* cpv = cs + (((size_t)dir * Tn + tp) * B + b0 + 4 * warp) * RT_H + u;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        qa[0][q] = rt_ldg_f32(a + q * G4); qa[1][q] = rt_ldg_f32(a + q * G4 + RT_H);
        qa[2][q] = rt_ldg_f32(a + q * G4 + 2 * RT_H); qa[3][q] = rt_ldg_f32(a + q * G4 + 3 * RT_H);
        qcp[q] = rt_ldg_f32(cpv + q * RT_H);
        qdy[q] = rt_ldg_u16(dyp + (size_t)(b0 + 4 * warp + q) * out_ld_b);
      }
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int b = min(b0 + 4 * warp + q, B - 1);
        const size_t row = ((size_t)dir * Tn + t) * B + b;
        const float* a = acts + row * G4 + u;
        qa[0][q] = rt_ldg_f32(a); qa[1][q] = rt_ldg_f32(a + RT_H); qa[2][q] = rt_ldg_f32(a + 2 * RT_H); qa[3][q] = rt_ldg_f32(a + 3 * RT_H);
        qcp[q] = rt_ldg_f32(cs + (((size_t)dir * Tn + tp) * B + b) * RT_H + u);
        qdy[q] = rt_ldg_u16(dyp + (size_t)b * out_ld_b);
      }
    }
  };
  auto advance = [&]() {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      pa[0][q] = rt_mov_f32(qa[0][q]); pa[1][q] = rt_mov_f32(qa[1][q]); pa[2][q] = rt_mov_f32(qa[2][q]);
      pa[3][q] = rt_mov_f32(qa[3][q]);
      pcp[q] = rt_mov_f32(qcp[q]); pdy[q] = rt_mov_u16(qdy[q]);
    }
  };
  auto t_of = [&](int s) { return dir ? s : (Tn - 1 - s); };
  if (Tn > 0) {
    prefetch(t_of(0)); advance();
    // c_t is loaded for the first step only: afterwards it is the previous step's c_{t-1} (carried in registers)
#pragma unroll
    for (int q = 0; q < 4; ++q)
      pct[q] = __ldg(cs + (((size_t)dir * Tn + t_of(0)) * B + min(b0 + 4 * warp + q, B - 1)) * RT_H + u);
  }
  if (Tn > 1) prefetch(t_of(1));

  // staged-tile -> global copy roles: chunk c = tid + 128 i covers sequence c / 16, k = 8 (c % 16) .. +7
  uint32_t cp_src[2], cp_dst[2];
  bool cp_ok[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const uint32_t c = tid + i * RT_THREADS, seq = c >> 4, k0 = (c & 15) * 8;
    cp_src[i] = bsm_u32 + (k0 >> 6) * 2048 + rt_swz(seq, k0 & 63);
    cp_dst[i] = seq * (uint32_t)G4 + (k0 >> 5) * RT_H + rank * RT_UPC + (k0 & 31);
    cp_ok[i] = b0 + (int)seq < B;
  }

  const bool tl_on = tl != nullptr && tid == 0 && rank == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  int cur = 0;
  uint32_t phase = 0;
  for (int s = 0; s < Tn; ++s) {
    const int t = dir ? s : (Tn - 1 - s);
    const int tp = dir ? t + 1 : t - 1;
    RT_TL(0);
    // ---- this step's saved activations were prefetched during the previous step (raw registers)
    float ai[4], af[4], ag[4], ao[4], ct[4], cp[4], dy[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      ai[q] = pa[0][q]; af[q] = pa[1][q]; ag[q] = pa[2][q]; ao[q] = pa[3][q];
      ct[q] = pct[q];
      cp[q] = (tp >= 0 && tp < Tn) ? pcp[q] : 0.f;
      pct[q] = pcp[q];                         // next step's c_t
      dy[q] = __uint_as_float((uint32_t)pdy[q] << 16);
    }
    if (s > 0) {                             // partial dh of the previous step from all 8 CTAs is in red[cur]
      rt_mbar_wait(&rfull[cur], rph[cur]);
      rph[cur] ^= 1;
    }
    // every epilogue warp has finished copying the previous step's staged tile before anyone overwrites it
    asm volatile("bar.sync 1, 128;" ::: "memory");
    RT_TL(1);
    float dg[4][4];
    float dhs[4] = {dy[0], dy[1], dy[2], dy[3]};
#pragma unroll
    for (int i = 0; i < RT_C; ++i) {           // red[buf][src][seq-quad pair][unit][2 quads x 4 bf16]: one 8 B read per source
      const uint2 r2 = rt_lds_v2u(red_u32 + (uint32_t)((((cur * RT_C + i) * 2 + (warp >> 1)) * RT_UPC + ul) * 16 + (warp & 1) * 8));
      dhs[0] += __uint_as_float(r2.x << 16); dhs[1] += __uint_as_float(r2.x & 0xffff0000u);
      dhs[2] += __uint_as_float(r2.y << 16); dhs[3] += __uint_as_float(r2.y & 0xffff0000u);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float dh = dhs[q];
      const bool valid = t < len_b[q];
      const float tc = rt_tanh(ct[q]);
      const float dc = dcrec[q] + dh * ao[q] * (1.f - tc * tc);
      dg[0][q] = valid ? dc * ag[q] * ai[q] * (1.f - ai[q]) : 0.f;
      dg[1][q] = valid ? dc * cp[q] * af[q] * (1.f - af[q]) : 0.f;
      dg[2][q] = valid ? dc * ai[q] * (1.f - ag[q] * ag[q]) : 0.f;
      dg[3][q] = valid ? dh * tc * ao[q] * (1.f - ao[q]) : 0.f;
      dcrec[q] = valid ? dc * af[q] : 0.f;
    }
    // ---- own gate gradients -> B operand [16 seqs][128 k], k = gate*32 + ul (K-major, swizzled) and -> global
#pragma unroll
    for (int gte = 0; gte < 4; ++gte) {
      const uint32_t kl = gte * 32 + ul;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t b = 4 * warp + q;
        rt_sts_b16(bsm_u32 + (kl >> 6) * 2048 + rt_swz(b, kl & 63) + (kl & 7) * 2,
                   __bfloat16_as_ushort(__float2bfloat16_rn(dg[gte][q])));
      }
    }
    RT_TL(2);
    rt_fence_async();
    rt_tc_before();
    __syncthreads();
    RT_TL(3);
    // the global copy of the gate gradients (for the weight-gradient GEMMs) is written while the MMAs run: the tile
    // just staged for the MMA ([16 seqs][128 k] bf16, swizzled) is copied out in 16-byte chunks, 2 per thread
    // (a per-thread register store would take 16 two-byte store instructions)
    {
      __nv_bfloat16* drow = dgates + (((size_t)dir * Tn + t) * B + b0) * G4;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const uint4 v = rt_lds_v4u(cp_src[i]);
        if (cp_ok[i]) *reinterpret_cast<uint4*>(drow + cp_dst[i]) = v;
      }
    }
    // next step's saved state: issued here so that no fence between now and its use has to wait for it
    RT_TL(4);
    advance();                                  // step s+1's state (loaded a full step ago) -> consume set
    if (s + 2 < Tn) prefetch(t_of(s + 2));
    RT_TL(5);
    rt_mbar_wait(mma_bar, phase);
    phase ^= 1;
    rt_tc_after();
    RT_TL(6);
    // ---- partial dh_{t-1}[unit, seq] for all 256 units: send each 32-unit block to the CTA that owns it
    const int nxt = cur ^ 1;
    if (s + 1 < Tn) {
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        float p[16];
        rt_tmem_ld_sum<2>(tmem_base + ((uint32_t)(warp * 32) << 16) + mt * 2 * RT_NB, p);
        const uint32_t owner = mt * 4 + warp;               // unit mt*128 + warp*32 + lane lives on CTA `owner`
        const uint32_t dst = rt_mapa(red_u32 + (uint32_t)((((nxt * RT_C + rank) * 2) * RT_UPC + lane) * 16), owner);
        const uint32_t bar = rt_mapa(rfull_u32 + nxt * 8, owner);
        // partial sums travel as bf16 (the SM-to-SM network moves ~24 B/clk per CTA: 8 KB instead of 16 KB per step)
#pragma unroll
        for (int j = 0; j < 2; ++j)
          rt_st_async_v4(dst + j * (RT_UPC * 16), rt_pack(p[8 * j], p[8 * j + 1]), rt_pack(p[8 * j + 2], p[8 * j + 3]),
                         rt_pack(p[8 * j + 4], p[8 * j + 5]), rt_pack(p[8 * j + 6], p[8 * j + 7]), bar);
      }
    }
    rt_tc_before();
    RT_TL(7);
    cur = nxt;
  }
  }  // epilogue warps
  rt_tc_before();
  rt_cluster_arrive();
  rt_cluster_wait();
  if (warp == 0) {
    rt_tc_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(RT_TMEM_COLS) : "memory");
  }
}

int set_timeline(void* buf) {
  long long* p = (long long*)buf;
  cudaError_t e = cudaMemcpyToSymbol(g_rt_timeline, &p, sizeof(p));
  return e == cudaSuccess ? 0 : set_error("debug_timeline: %s", cudaGetErrorString(e));
}

bool blstm_tc_eligible(int dtype, int64_t H, int64_t out_ld_t, int64_t out_ld_b) {
  return dtype == B200ST_BF16 && H == RT_H && out_ld_t % 4 == 0 && out_ld_b % 4 == 0;
}

int blstm_fwd_tc(const void* xproj, const float* w_hh_f, const float* w_hh_r, const int32_t* lens, void* out,
                 int64_t out_ld_t, int64_t out_ld_b, int pair, void* hs, float* acts, float* cs, int64_t T_, int64_t B,
                 cudaStream_t st) {
  B200ST_CUDA(cudaFuncSetAttribute((const void*)blstm_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
  dim3 grid(RT_C, (unsigned)ceil_div(B, RT_NB), 2);
  blstm_fwd_tc_kernel<<<grid, RT_BLOCK, FWD_SMEM, st>>>((const __nv_bfloat16*)xproj, w_hh_f, w_hh_r, lens,
                                                           (__nv_bfloat16*)out, out_ld_t, out_ld_b, pair,
                                                           (__nv_bfloat16*)hs, acts, cs, (int)T_, (int)B);
  B200ST_LAUNCH_CHECK("blstm_fwd_tc");
  return 0;
}

int blstm_bwd_tc(const void* dout, int64_t out_ld_t, int64_t out_ld_b, int pair, const float* acts, const float* cs,
                 const float* w_hh_f, const float* w_hh_r, const int32_t* lens, void* dgates, int64_t T_, int64_t B,
                 cudaStream_t st) {
  B200ST_CUDA(cudaFuncSetAttribute((const void*)blstm_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
  dim3 grid(RT_C, (unsigned)ceil_div(B, RT_NB), 2);
  blstm_bwd_tc_kernel<<<grid, RT_BLOCK, BWD_SMEM, st>>>((const __nv_bfloat16*)dout, out_ld_t, out_ld_b, pair, acts, cs,
                                                           w_hh_f, w_hh_r, lens, (__nv_bfloat16*)dgates, (int)T_, (int)B);
  B200ST_LAUNCH_CHECK("blstm_bwd_tc");
  return 0;
}

}  // namespace b200st
