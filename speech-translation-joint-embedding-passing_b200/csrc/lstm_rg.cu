// Persistent bidirectional LSTM recurrence with the recurrent weights resident in REGISTERS and the per-step GEMM on
// the warp-level tensor path (mma.sync.m16n8k16, bf16 operands, fp32 accumulate / state).  H = 256 per direction
// (the reference's acoustic encoder, Seq2seq.py:57; Enc.py:150-211).  ALTERNATIVE backend (b200st_set_blstm_backend(3)):
// measured on a par with the tcgen05 kernels of lstm_tc.cu (0.97 vs 1.00 us/step forward, 1.3 vs 1.16 backward at
// T = 1008, B = 64; run-to-run and box-to-box spread is larger than the difference), so lstm_tc.cu stays the default.
// What it shows (profiles/r02_blstm_experiments.txt): the step is bound by the h all-gather over distributed shared
// memory (~750-1000 cycles from the send until the last peer's slice has landed, whatever issues the stores), not by how
// the 128 x 16 x 256 product of a CTA is computed.
//
// Why try the warp-level MMA at all: one time step is D[128 gate rows, 16 seqs] per CTA -- at N = 16 a tcgen05 step pays
// fixed latencies that dwarf the math (proxy fence 35 + 16 TS-form MMAs 400 + commit/mbarrier round trip 210 +
// tcgen05.ld 120 cycles = ~770 per step, measured) and delivers the gates with one gate ROW per thread, so the cell
// update needs a shared-memory transpose and a CTA barrier.  The warp-level MMA has a lower peak (measured ~10 cycles
// per m16n8k16 per SM sub-partition, 660 cycles for the 256 HMMAs of a step, scripts/probes/hmma_probe.cu) but no fixed
// costs, and its accumulator fragment can be arranged so that all four gates of a (unit, sequence) pair land in ONE
// thread: the whole step is wait -> ldmatrix/HMMA -> activations + cell update in registers -> send.
//
// Work layout (forward): grid = (8, ceil(B/16), 2), a CLUSTER of 8 CTAs per (direction, group of 16 sequences), 8 warps
// per CTA.  CTA `rank` owns hidden units [32 rank, 32 rank + 32).  Warp w = (ub = w & 3, nt = w >> 2) owns units
// 8 ub .. 8 ub + 7 of the CTA for the 8 sequences of n-tile nt; it holds two 16-row A tiles for all K = 256:
//   tile 0 rows 0-7 = input gate of its 8 units, rows 8-15 = forget gate;  tile 1 rows 0-7 = candidate, rows 8-15 = output
// (128 registers per thread, loaded once).  In the m16n8k16 accumulator layout thread (r = lane / 4, c = lane % 4) then
// holds i, f, g, o of unit r for sequences 2c, 2c+1.
// The B operand is h_{t-1} as [k = unit][8 sequences] (16-byte rows, one 4 KB tile per n-tile), read with
// ldmatrix.x4.trans.  A thread's two h values are packed to bf16x2, the four lanes of a quad gather the 16-byte row of
// their unit with shuffles, and each lane sends it to two of the 8 CTAs with `st.async ... mbarrier::complete_tx`
// (distributed shared memory; data and signal travel together, no cluster barrier in the loop).
// The two n-tiles are INDEPENDENT recurrence chains (different sequences) with their own buffers and mbarriers.
// (16 clusters of 8 sequences -- one chain per CTA on 128 SMs -- cannot be co-resident: cudaOccupancyMaxActiveClusters
// reports 15 clusters of 8 one-CTA-per-SM blocks on the B200, scripts/probes/cluster_occ.cu.)
// Global-memory work per step is kept off the serial instruction stream: cursors advanced by constant strides instead of
// offsets recomputed from t, and the saved gates / cell states in a blocked layout (a thread's values are contiguous).
//
// Backward: same structure with A = W_hh^T[256 units, own 128 gate rows]: every CTA multiplies its own gate gradients
// into partial dh for all 256 units and the bf16 partials are reduce-scattered to the owning CTAs through DSMEM.
#include "common.cuh"

namespace b200st {

constexpr int RG_H = 256;        // hidden units per direction
constexpr int RG_C = 8;          // CTAs per cluster
constexpr int RG_UPC = 32;       // units per CTA
constexpr int RG_NB = 16;        // sequences per cluster (two n-tiles of 8)
constexpr int RG_THREADS = 256;  // 8 warps

__device__ long long* g_rg_timeline = nullptr;
static bool g_rg_timeline_on = false;     // host flag: launch the instrumented instantiation
#define RG_TL(i) do { if (tl_on && s >= 64 && s < 72) tl[(s - 64) * 16 + (i)] = clock64(); } while (0)

__device__ __forceinline__ uint32_t rg_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t rg_cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t rg_mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void rg_st_async_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d), "r"(bar) : "memory");
}
__device__ __forceinline__ void rg_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void rg_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rg_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (spin > (1u << 22)) __trap();
  }
}
__device__ __forceinline__ void rg_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void rg_ldsm4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void rg_hmma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float rg_tanh(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rg_sigmoid(float x) { return fmaf(rg_tanh(0.5f * x), 0.5f, 0.5f); }
__device__ __forceinline__ uint32_t rg_pack(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ float rg_bf(unsigned short v) { return __uint_as_float((uint32_t)v << 16); }
// position-pinned read-only loads / moves of the two-step-ahead prefetch pipeline (see lstm_tc.cu)
__device__ __forceinline__ unsigned short rg_ldg_u16(const unsigned short* p) { unsigned short v; asm volatile("ld.global.nc.u16 %0, [%1];" : "=h"(v) : "l"(p)); return v; }
__device__ __forceinline__ float rg_ldg_f32(const float* p) { float v; asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v; }
__device__ __forceinline__ void rg_ldg_v4(const float4* p, float& a, float& b, float& c, float& d) {
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "l"(p));
}
__device__ __forceinline__ void rg_ldg_v2(const float2* p, float& a, float& b) {
  asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(a), "=f"(b) : "l"(p));
}
__device__ __forceinline__ unsigned short rg_mov_u16(unsigned short x) { unsigned short v; asm volatile("mov.b16 %0, %1;" : "=h"(v) : "h"(x)); return v; }
__device__ __forceinline__ float rg_mov_f32(float x) { float v; asm volatile("mov.f32 %0, %1;" : "=f"(v) : "f"(x)); return v; }

// shared memory (forward): hbuf [2 buffers][2 n-tiles][256 k][8 seqs] bf16 = 16 KB | hfull [2][2] mbarriers
// NT = independent chains (n-tiles of 8 sequences) per CTA: 2 = 16 sequences per cluster, 256 threads; 1 = 8 sequences, 128 threads
template <bool TLINE, int NT>
__global__ void __cluster_dims__(RG_C, 1, 1) __launch_bounds__(128 * NT, 1)
blstm_fwd_rg_kernel(const __nv_bfloat16* __restrict__ xproj, const float* __restrict__ w_hh_f,
                    const float* __restrict__ w_hh_r, const int32_t* __restrict__ lens,
                    __nv_bfloat16* __restrict__ out, int64_t out_ld_t, int64_t out_ld_b, int pair,
                    __nv_bfloat16* __restrict__ hs, float* __restrict__ acts, float* __restrict__ cs, int Tn, int B) {
  constexpr int RGF_H_BYTES = 2 * NT * RG_H * 16, NTHR = 128 * NT;
  __shared__ __align__(1024) uint8_t hbuf[RGF_H_BYTES];
  __shared__ __align__(8) uint64_t hfull[2 * NT];    // [buffer][n-tile]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ub = warp & 3, nt = warp >> 2, r = lane >> 2, c = lane & 3;
  const uint32_t rank = rg_cluster_rank();
  const int grp = blockIdx.y, dir = blockIdx.z;
  const float* w = dir ? w_hh_r : w_hh_f;
  const int u = rank * RG_UPC + ub * 8 + r;          // this thread's hidden unit (direction-local index)

  for (int i = tid; i < RGF_H_BYTES / 16; i += NTHR) reinterpret_cast<uint4*>(hbuf)[i] = make_uint4(0, 0, 0, 0);
  const uint32_t hbuf_u32 = rg_smem_u32(hbuf), hfull_u32 = rg_smem_u32(hfull);
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < 2 * NT; ++i) rg_mbar_init(hfull_u32 + i * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
    for (int i = 0; i < NT; ++i) rg_mbar_expect_tx(hfull_u32 + (NT + i) * 8, 4096);      // step 0 fills buffer 1
  }
  // ---- resident weights: A fragments of the two 16-row tiles for all 16 k-steps (row-major m16k16 fragment layout:
  // a0 = (row r, k 2c..2c+1), a1 = (row r + 8, same k), a2 = (row r, k + 8), a3 = (row r + 8, k + 8))
  uint32_t afr[2][16][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    const float* wlo = w + (size_t)((2 * mt) * RG_H + u) * RG_H + 2 * c;        // gates i / g
    const float* whi = w + (size_t)((2 * mt + 1) * RG_H + u) * RG_H + 2 * c;    // gates f / o
#pragma unroll
    for (int ks = 0; ks < 16; ++ks) {
      const float2 l0 = *reinterpret_cast<const float2*>(wlo + ks * 16), l1 = *reinterpret_cast<const float2*>(wlo + ks * 16 + 8);
      const float2 h0 = *reinterpret_cast<const float2*>(whi + ks * 16), h1 = *reinterpret_cast<const float2*>(whi + ks * 16 + 8);
      afr[mt][ks][0] = rg_pack(l0.x, l0.y); afr[mt][ks][1] = rg_pack(h0.x, h0.y);
      afr[mt][ks][2] = rg_pack(l1.x, l1.y); afr[mt][ks][3] = rg_pack(h1.x, h1.y);
    }
  }
  __syncthreads();
  rg_cluster_sync();         // every CTA has initialised its barriers and zeroed its h buffers

  // ---- per-thread constants
  const int b0 = grp * (8 * NT) + nt * 8 + 2 * c;       // this thread's two sequences: b0, b0 + 1
  const bool ok0 = b0 < B, ok1 = b0 + 1 < B;
  const int len0 = ok0 ? lens[b0] : 0, len1 = ok1 ? lens[b0 + 1] : 0;
  float c_st[2] = {0.f, 0.f}, h_st[2] = {0.f, 0.f};
  // B operand rows of this warp's n-tile: lane l addresses row (32 kp + l) of the k-pair block kp
  const uint32_t ld_base = hbuf_u32 + nt * 4096 + lane * 16;
  // remote destinations: the quad's 16-byte row (unit u, n-tile nt) goes to CTAs (2c + rank) & 7 and (2c + 1 + rank) & 7
  // (mapa is offset-preserving: the peer's barrier sits at the same distance from the row as the local one)
  uint32_t dst_row[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) dst_row[i] = rg_mapa(hbuf_u32 + nt * 4096 + u * 16, (2 * c + i + rank) & (RG_C - 1));
  const uint32_t my_bar = hfull_u32 + nt * 8;
  const uint32_t bar_delta = my_bar - (hbuf_u32 + nt * 4096 + u * 16);
  const bool armer = (ub == 0 && lane == 0);

  // Global-memory cursors, advanced by a constant stride per step (a step's address arithmetic is on the warp's serial
  // instruction stream: recomputing 64-bit offsets from t cost ~150 instructions per step, as much as the rest of the step).
  // x-projection [2][T][B][4H] bf16: rows b0 and b0 + 1 are adjacent; loads of a missing sequence are predicated off.
  // Saved state for backward lives in a kernel-private BLOCKED layout (only blstm_bwd_rg_kernel reads it):
  //   acts [2][T][grp][rank][thread 256][gate 4][seq 2] fp32, cs [2][T][grp][rank][thread 256][seq 2] fp32
  // so a thread's 8 + 2 values are two 16-byte stores + one 8-byte store and a warp writes 1 KB contiguous.
  const int t0 = dir ? Tn - 1 : 0;
  const int tdir = dir ? -1 : 1;
  const int G = gridDim.y;
  const unsigned short* xp = reinterpret_cast<const unsigned short*>(xproj) + (((size_t)dir * Tn + t0) * B + b0) * (4 * RG_H) + u;
  const int x_stride = tdir * B * (4 * RG_H);
  long long blk = ((((long long)dir * Tn + t0) * G + grp) * RG_C + rank) * NTHR + tid;      // (step, thread) slot
  const int blk_stride = tdir * G * RG_C * NTHR;
  __nv_bfloat16* hs_p = hs ? hs + (((size_t)dir * (Tn + 1) + (dir ? t0 : t0 + 1)) * B + b0) * RG_H + u : nullptr;
  const int hs_stride = tdir * B * RG_H;
  __nv_bfloat16* out_b = out + (size_t)b0 * out_ld_b + dir * RG_H + u;
  const int psh = pair - 1;                          // pair is 1 or 2 (checked by the entry point)

  // x-projection prefetch, two steps ahead: [gate][seq] raw bf16 bits
  unsigned short xr[8], xq[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) xq[i] = 0;
  auto load_x = [&]() {              // loads the step xp points at, then advances the cursor
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      if (ok0) xq[2 * g] = rg_ldg_u16(xp + g * RG_H);
      if (ok1) xq[2 * g + 1] = rg_ldg_u16(xp + 4 * RG_H + g * RG_H);
    }
    xp += x_stride;
  };
  auto advance_x = [&]() {
#pragma unroll
    for (int i = 0; i < 8; ++i) xr[i] = rg_mov_u16(xq[i]);
  };
  if (Tn > 0) { load_x(); advance_x(); }
  if (Tn > 1) load_x();

  long long* tl = TLINE ? g_rg_timeline : nullptr;
  const bool tl_on = TLINE && tl != nullptr && tid == 0 && rank == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  int cur = 0;
  for (int s = 0; s < Tn; ++s) {
    const int t = dir ? (Tn - 1 - s) : s;
    RG_TL(0);
    if (s > 0) {                               // this chain's h_{t-1} rows from all 8 CTAs have landed in hbuf[cur][nt]
      rg_mbar_wait(my_bar + cur * (8 * NT), ((uint32_t)(s - 1) >> 1) & 1u);     // buffer s & 1 completes its ((s-1)/2)-th phase
    }
    RG_TL(1);
    // arm hbuf[cur][nt] for h_{t+1}: no peer can send it before it has received this step's h_t from every warp of this chain
    if (armer && s + 2 < Tn) rg_mbar_expect_tx(my_bar + cur * (8 * NT), 4096);
    // ---- gates[128 rows, 8 seqs] of this warp: accumulators start from the x-projection (biases included)
    // (one accumulator chain per tile: splitting K over two chains measured no faster, scripts/probes/hmma_probe.cu)
    float acc[2][4];                           // [tile][fragment]
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      acc[mt][0] = rg_bf(xr[4 * mt]);     acc[mt][1] = rg_bf(xr[4 * mt + 1]);       // gate 2mt (rows r), seqs 0 / 1
      acc[mt][2] = rg_bf(xr[4 * mt + 2]); acc[mt][3] = rg_bf(xr[4 * mt + 3]);       // gate 2mt + 1 (rows r + 8)
    }
    const uint32_t la = ld_base + cur * (4096 * NT);
#pragma unroll
    for (int kp = 0; kp < 8; ++kp) {
      uint32_t bfr[4];
      rg_ldsm4_t(la + kp * 512, bfr);
      rg_hmma(acc[0], afr[0][2 * kp], bfr[0], bfr[1]);
      rg_hmma(acc[1], afr[1][2 * kp], bfr[0], bfr[1]);
      rg_hmma(acc[0], afr[0][2 * kp + 1], bfr[2], bfr[3]);
      rg_hmma(acc[1], afr[1][2 * kp + 1], bfr[2], bfr[3]);
    }
    RG_TL(2);
    // ---- activations + cell update, all in registers
    float gi[2], gf[2], gg[2], go[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      gi[j] = rg_sigmoid(acc[0][j]);
      gf[j] = rg_sigmoid(acc[0][2 + j]);
      gg[j] = rg_tanh(acc[1][j]);
      go[j] = rg_sigmoid(acc[1][2 + j]);
    }
    const bool v0 = t < len0, v1 = t < len1;
    if (v0) { c_st[0] = fmaf(gf[0], c_st[0], gi[0] * gg[0]); h_st[0] = go[0] * rg_tanh(c_st[0]); }
    if (v1) { c_st[1] = fmaf(gf[1], c_st[1], gi[1] * gg[1]); h_st[1] = go[1] * rg_tanh(c_st[1]); }
    RG_TL(3);
    // ---- the quad gathers the 16-byte row (unit u, 8 sequences of this n-tile) and scatters it to the cluster
    if (s + 1 < Tn) {
      const uint32_t v = rg_pack(h_st[0], h_st[1]);
      const int q0 = lane & ~3;
      const uint32_t x0 = __shfl_sync(0xffffffffu, v, q0), x1 = __shfl_sync(0xffffffffu, v, q0 + 1);
      const uint32_t x2 = __shfl_sync(0xffffffffu, v, q0 + 2), x3 = __shfl_sync(0xffffffffu, v, q0 + 3);
      const uint32_t boff = (cur ^ 1) * (4096 * NT), moff = (cur ^ 1) * (8 * NT);
      rg_st_async_v4(dst_row[0] + boff, x0, x1, x2, x3, dst_row[0] + bar_delta + moff);
      rg_st_async_v4(dst_row[1] + boff, x0, x1, x2, x3, dst_row[1] + bar_delta + moff);
    }
    RG_TL(4);
    // ---- x-projection pipeline: step s+1's bits (loaded a full step ago) -> consume set; issue step s+2's loads
    advance_x();
    RG_TL(6);
    if (s + 2 < Tn) load_x();
    RG_TL(7);
    // ---- global stores (saved state for backward, layer output): nothing on the recurrent chain waits for them
    if (acts) {
      float4* ap = reinterpret_cast<float4*>(acts) + blk * 2;
      ap[0] = make_float4(gi[0], gi[1], gf[0], gf[1]);
      ap[1] = make_float4(gg[0], gg[1], go[0], go[1]);
      reinterpret_cast<float2*>(cs)[blk] = make_float2(c_st[0], c_st[1]);
      blk += blk_stride;
    }
    RG_TL(8);
    {
      const __nv_bfloat16 hv0 = __float2bfloat16_rn(v0 ? h_st[0] : 0.f), hv1 = __float2bfloat16_rn(v1 ? h_st[1] : 0.f);
      __nv_bfloat16* o = out_b + (size_t)(t >> psh) * out_ld_t + (size_t)(t & psh) * (2 * RG_H);
      if (ok0) o[0] = hv0;
      if (ok1) o[out_ld_b] = hv1;
      if (hs_p) {
        if (ok0) hs_p[0] = hv0;
        if (ok1) hs_p[RG_H] = hv1;
        hs_p += hs_stride;
      }
    }
    RG_TL(5);
    cur ^= 1;
  }
  rg_cluster_sync();         // nobody exits while a peer could still address its shared memory
}

// ------------------------------------------------------------------------------------------------------------------
// Backward.  CTA `rank` owns the 128 gate rows of its 32 units (k = gate * 32 + local unit).  Per step and chain:
//   wait for the 8 partial-dh blocks of the previous step -> dh = dy + sum -> pointwise LSTM backward for (unit, 2 seqs)
//   per thread -> own gate gradients as the B tile [k][8 seqs] in shared memory -> named barrier (4 warps of the chain)
//   -> partial dh_{t-1}[u' = all 256 units, 8 seqs] = W_hh^T[:, own rows] . dG: warp (mb, nt) owns units 64 mb .. 64 mb + 63
//   (4 m-tiles x 8 k-steps, A = 128 registers per thread) -> 4x4 quad transposes (6 shuffles) give every lane two
//   complete 16-byte rows (unit, 8 seqs) -> st.async to the two owning CTAs -> gate gradients to global memory.
// shared memory: red [2 buffers][2 n-tiles][8 src][32 units][8 seqs] bf16 = 16 KB | btile [2 n-tiles][128 k][8 seqs] bf16 4 KB
constexpr int RGB_RED_BYTES = 2 * 2 * RG_C * RG_UPC * 16;
constexpr int RGB_BT_BYTES = 2 * 128 * 16;

// One round of the 4x4 transpose across a quad: lane c hands the word meant for lane c ^ j to it and receives its own.
// branch-free 1-of-4 select (the ?: form compiles to divergent branches around the shuffles: ~2500 cycles per step)
__device__ __forceinline__ uint32_t rg_sel4(const uint32_t (&v)[4], int i) {
  uint32_t o;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b32 lo, hi;\n\t"
      "and.b32 lo, %5, 1;\n\tsetp.ne.b32 p, lo, 0;\n\t"
      "and.b32 hi, %5, 2;\n\tsetp.ne.b32 q, hi, 0;\n\t"
      "selp.b32 lo, %2, %1, p;\n\tselp.b32 hi, %4, %3, p;\n\tselp.b32 %0, hi, lo, q;\n\t}"
      : "=r"(o) : "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(i));
  return o;
}
// in: v[j] = this lane's word (2 seqs) of row j;  out: o[cc] = word of lane cc for row `c` (this lane's row, complete)
__device__ __forceinline__ void rg_quad_transpose(const uint32_t (&v)[4], int c, uint32_t (&o)[4]) {
  uint32_t t[4];
  t[0] = rg_sel4(v, c);
#pragma unroll
  for (int j = 1; j < 4; ++j) t[j] = __shfl_xor_sync(0xffffffffu, rg_sel4(v, c ^ j), j);    // from lane c ^ j: its word of row c
  // t[j] came from lane c ^ j -> slot (c ^ j)
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) o[cc] = rg_sel4(t, cc ^ c);
}

template <bool TLINE>
__global__ void __cluster_dims__(RG_C, 1, 1) __launch_bounds__(RG_THREADS, 1)
blstm_bwd_rg_kernel(const __nv_bfloat16* __restrict__ dout, int64_t out_ld_t, int64_t out_ld_b, int pair,
                    const float* __restrict__ acts, const float* __restrict__ cs, const float* __restrict__ w_hh_f,
                    const float* __restrict__ w_hh_r, const int32_t* __restrict__ lens,
                    __nv_bfloat16* __restrict__ dgates, int Tn, int B) {
  __shared__ __align__(1024) uint8_t red[RGB_RED_BYTES];
  __shared__ __align__(128) uint8_t btile[RGB_BT_BYTES];
  __shared__ __align__(8) uint64_t rfull[4];         // [buffer][n-tile]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ub = warp & 3, nt = warp >> 2, r = lane >> 2, c = lane & 3;      // ub doubles as mb in the GEMM phase
  const uint32_t rank = rg_cluster_rank();
  const int grp = blockIdx.y, dir = blockIdx.z;
  const float* w = dir ? w_hh_r : w_hh_f;
  const int lu = ub * 8 + r;                         // pointwise phase: local unit
  const int u = rank * RG_UPC + lu;

  for (int i = tid; i < RGB_RED_BYTES / 16; i += RG_THREADS) reinterpret_cast<uint4*>(red)[i] = make_uint4(0, 0, 0, 0);
  const uint32_t red_u32 = rg_smem_u32(red), bt_u32 = rg_smem_u32(btile), rfull_u32 = rg_smem_u32(rfull);
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) rg_mbar_init(rfull_u32 + i * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    rg_mbar_expect_tx(rfull_u32 + 2 * 8, 4096);      // step 0 sends its partials into buffer 1
    rg_mbar_expect_tx(rfull_u32 + 3 * 8, 4096);
  }
  // ---- resident W_hh^T: m-tile mt rows = units 64 ub + 16 mt + r (+8), k = own gate row (gate * 32 + local unit)
  uint32_t afr[4][8][4];
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) {
    const int u0 = ub * 64 + mt * 16 + r;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      const int k0 = ks * 16 + 2 * c;
      auto wv = [&](int k, int uu) { return w[(size_t)((k >> 5) * RG_H + rank * RG_UPC + (k & 31)) * RG_H + uu]; };
      afr[mt][ks][0] = rg_pack(wv(k0, u0), wv(k0 + 1, u0));
      afr[mt][ks][1] = rg_pack(wv(k0, u0 + 8), wv(k0 + 1, u0 + 8));
      afr[mt][ks][2] = rg_pack(wv(k0 + 8, u0), wv(k0 + 9, u0));
      afr[mt][ks][3] = rg_pack(wv(k0 + 8, u0 + 8), wv(k0 + 9, u0 + 8));
    }
  }
  __syncthreads();
  rg_cluster_sync();

  const int b0 = grp * RG_NB + nt * 8 + 2 * c;
  const bool ok0 = b0 < B, ok1 = b0 + 1 < B;
  const int len0 = ok0 ? lens[b0] : 0, len1 = ok1 ? lens[b0 + 1] : 0;
  float dcrec[2] = {0.f, 0.f};
  const size_t G4 = 4 * RG_H;
  // partial-dh rows this lane sends: (owner 2 ub, local unit 8 c + r) and (owner 2 ub + 1, same), slot = own rank
  uint32_t dst_row[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) dst_row[i] = rg_mapa(red_u32 + nt * 4096 + rank * 512 + (8 * c + r) * 16, 2 * ub + i);
  const uint32_t my_bar = rfull_u32 + nt * 8;
  const uint32_t bar_delta = my_bar - (red_u32 + nt * 4096 + rank * 512 + (8 * c + r) * 16);      // mapa is offset-preserving
  const bool armer = (ub == 0 && lane == 0);
  const uint32_t rd_base = red_u32 + nt * 4096 + lu * 16 + c * 4;          // + buffer * 8192 + src * 512
  const uint32_t bt_wr = bt_u32 + nt * 2048 + lu * 16 + c * 4;             // + gate * 512
  const uint32_t bt_rd = bt_u32 + nt * 2048 + lane * 16;                   // + kp * 512

  // Global-memory cursors (constant stride per step, see the forward kernel).  Saved state in the forward kernel's blocked
  // layout: this thread's (unit, 2 seqs) pair is exactly what the forward thread with the same index produced.
  const int t0 = dir ? 0 : Tn - 1;
  const int tdir = dir ? 1 : -1;                     // backward walks time against the forward direction
  const int G = gridDim.y;
  long long blk = ((((long long)dir * Tn + t0) * G + grp) * RG_C + rank) * RG_THREADS + tid;      // (step, thread) slot
  const int blk_stride = tdir * G * RG_C * RG_THREADS;
  const unsigned short* dy_b = reinterpret_cast<const unsigned short*>(dout) + (size_t)b0 * out_ld_b + dir * RG_H + u;
  unsigned short* dg_p = reinterpret_cast<unsigned short*>(dgates) + (((size_t)dir * Tn + t0) * B + b0) * G4 + u;
  const int dg_stride = tdir * B * (int)G4;
  const int psh = pair - 1;                          // pair is 1 or 2 (checked by the entry point)

  // saved-state prefetch, two steps ahead (see lstm_tc.cu): activations [gate][seq], c_{t-1} (= c_t of the NEXT step
  // processed), dy.  `prefetch(sp)` loads step sp (the cursors point at it) and advances them.
  float pa[8], pcp[2], pct[2], qa[8], qcp[2];
  unsigned short pdy[2], qdy[2];
  qdy[0] = qdy[1] = 0;
  qcp[0] = qcp[1] = 0.f;
  auto prefetch = [&](int sp) {
    const int t = dir ? sp : (Tn - 1 - sp);
    const float4* ap = reinterpret_cast<const float4*>(acts) + blk * 2;
    rg_ldg_v4(ap, qa[0], qa[1], qa[2], qa[3]);
    rg_ldg_v4(ap + 1, qa[4], qa[5], qa[6], qa[7]);
    blk += blk_stride;                               // -> step sp + 1, whose c_t is this step's c_{t-1}
    if (sp + 1 < Tn) rg_ldg_v2(reinterpret_cast<const float2*>(cs) + blk, qcp[0], qcp[1]);
    const unsigned short* dyp = dy_b + (size_t)(t >> psh) * out_ld_t + (size_t)(t & psh) * (2 * RG_H);
    if (ok0) qdy[0] = rg_ldg_u16(dyp);
    if (ok1) qdy[1] = rg_ldg_u16(dyp + out_ld_b);
  };
  auto advance = [&]() {
#pragma unroll
    for (int i = 0; i < 8; ++i) pa[i] = rg_mov_f32(qa[i]);
    pcp[0] = rg_mov_f32(qcp[0]); pcp[1] = rg_mov_f32(qcp[1]);
    pdy[0] = rg_mov_u16(qdy[0]); pdy[1] = rg_mov_u16(qdy[1]);
  };
  if (Tn > 0) {
    const float2 c0v = reinterpret_cast<const float2*>(cs)[blk];      // c_t of the first step; afterwards carried
    pct[0] = c0v.x; pct[1] = c0v.y;
    prefetch(0); advance();
  }
  if (Tn > 1) prefetch(1);

  long long* tl = TLINE ? g_rg_timeline : nullptr;
  const bool tl_on = TLINE && tl != nullptr && tid == 0 && rank == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  int cur = 0;
  for (int s = 0; s < Tn; ++s) {
    const int t = dir ? s : (Tn - 1 - s);
    RG_TL(0);
    float ai[2], af[2], ag[2], ao[2], ct[2], cp[2], dh[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      ai[j] = pa[j]; af[j] = pa[2 + j]; ag[j] = pa[4 + j]; ao[j] = pa[6 + j];      // [gate][seq]
      ct[j] = pct[j];
      cp[j] = (s + 1 < Tn) ? pcp[j] : 0.f;
      pct[j] = pcp[j];                           // next step's c_t
      dh[j] = rg_bf(pdy[j]);
    }
    if (s > 0) {                                 // this chain's 8 partial blocks of the previous step are in red[cur][nt]
      rg_mbar_wait(my_bar + cur * 16, ((uint32_t)(s - 1) >> 1) & 1u);
      RG_TL(1);
      const uint32_t ra = rd_base + cur * 8192;
#pragma unroll
      for (int i = 0; i < RG_C; ++i) {
        uint32_t v;
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(ra + i * 512) : "memory");
        dh[0] += __uint_as_float(v << 16); dh[1] += __uint_as_float(v & 0xffff0000u);
      }
    } else {
      RG_TL(1);
    }
    if (armer && s + 2 < Tn) rg_mbar_expect_tx(my_bar + cur * 16, 4096);
    float dg[4][2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const bool valid = t < (j ? len1 : len0);
      const float tc = rg_tanh(ct[j]);
      const float dc = dcrec[j] + dh[j] * ao[j] * (1.f - tc * tc);
      dg[0][j] = valid ? dc * ag[j] * ai[j] * (1.f - ai[j]) : 0.f;
      dg[1][j] = valid ? dc * cp[j] * af[j] * (1.f - af[j]) : 0.f;
      dg[2][j] = valid ? dc * ai[j] * (1.f - ag[j] * ag[j]) : 0.f;
      dg[3][j] = valid ? dh[j] * tc * ao[j] * (1.f - ao[j]) : 0.f;
      dcrec[j] = valid ? dc * af[j] : 0.f;
    }
    uint32_t dgp[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      dgp[g] = rg_pack(dg[g][0], dg[g][1]);
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(bt_wr + g * 512), "r"(dgp[g]) : "memory");
    }
    RG_TL(2);
    asm volatile("bar.sync %0, 128;" ::"r"(1 + nt) : "memory");      // the chain's 4 warps have written the B tile
    float acc[4][4];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) acc[mt][0] = acc[mt][1] = acc[mt][2] = acc[mt][3] = 0.f;
    if (s + 1 < Tn) {
#pragma unroll
      for (int kp = 0; kp < 4; ++kp) {
        uint32_t bfr[4];
        rg_ldsm4_t(bt_rd + kp * 512, bfr);
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) rg_hmma(acc[mt], afr[mt][2 * kp], bfr[0], bfr[1]);
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) rg_hmma(acc[mt], afr[mt][2 * kp + 1], bfr[2], bfr[3]);
      }
      RG_TL(3);
      // rows j = 2 mt + h (unit 64 ub + 8 j + r): block 0 = j 0..3 -> owner 2 ub, block 1 = j 4..7 -> owner 2 ub + 1
      const uint32_t boff = (cur ^ 1) * 8192, moff = (cur ^ 1) * 16;
#pragma unroll
      for (int hb = 0; hb < 2; ++hb) {
        uint32_t v[4], o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int mt = 2 * hb + (j >> 1), h = j & 1;
          v[j] = rg_pack(acc[mt][2 * h], acc[mt][2 * h + 1]);
        }
        rg_quad_transpose(v, c, o);
        rg_st_async_v4(dst_row[hb] + boff, o[0], o[1], o[2], o[3], dst_row[hb] + bar_delta + moff);
      }
    } else {
      RG_TL(3);
    }
    RG_TL(4);
    // ---- next step's saved state; gate gradients to global memory (for the weight-gradient / input-gradient GEMMs)
    advance();
    if (s + 2 < Tn) prefetch(s + 2);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      if (ok0) dg_p[g * RG_H] = (unsigned short)(dgp[g] & 0xffffu);
      if (ok1) dg_p[G4 + g * RG_H] = (unsigned short)(dgp[g] >> 16);
    }
    dg_p += dg_stride;
    RG_TL(5);
    cur ^= 1;
  }
  rg_cluster_sync();
}

int set_timeline_rg(void* buf) {
  long long* p = (long long*)buf;
  g_rg_timeline_on = buf != nullptr;
  cudaError_t e = cudaMemcpyToSymbol(g_rg_timeline, &p, sizeof(p));
  return e == cudaSuccess ? 0 : set_error("debug_timeline: %s", cudaGetErrorString(e));
}

bool blstm_rg_eligible(int dtype, int64_t H) { return dtype == B200ST_BF16 && H == RG_H; }

int blstm_fwd_rg(const void* xproj, const float* w_hh_f, const float* w_hh_r, const int32_t* lens, void* out,
                 int64_t out_ld_t, int64_t out_ld_b, int pair, void* hs, float* acts, float* cs, int64_t T_, int64_t B,
                 cudaStream_t st) {
  if ((acts == nullptr) != (cs == nullptr)) return set_error("blstm_fwd: acts and cs are saved together");
  dim3 grid(RG_C, (unsigned)((B + RG_NB - 1) / RG_NB), 2);
#define RG_FWD_LAUNCH(TL, NT_)                                                                                         \
  blstm_fwd_rg_kernel<TL, NT_><<<grid, 128 * NT_, 0, st>>>((const __nv_bfloat16*)xproj, w_hh_f, w_hh_r, lens,          \
                                                          (__nv_bfloat16*)out, out_ld_t, out_ld_b, pair,               \
                                                          (__nv_bfloat16*)hs, acts, cs, (int)T_, (int)B)
  if (g_rg_timeline_on) RG_FWD_LAUNCH(true, 2); else RG_FWD_LAUNCH(false, 2);
#undef RG_FWD_LAUNCH
  B200ST_LAUNCH_CHECK("blstm_fwd_rg");
  return 0;
}

int blstm_bwd_rg(const void* dout, int64_t out_ld_t, int64_t out_ld_b, int pair, const float* acts, const float* cs,
                 const float* w_hh_f, const float* w_hh_r, const int32_t* lens, void* dgates, int64_t T_, int64_t B,
                 cudaStream_t st) {
  dim3 grid(RG_C, (unsigned)((B + RG_NB - 1) / RG_NB), 2);
  if (g_rg_timeline_on)
    blstm_bwd_rg_kernel<true><<<grid, RG_THREADS, 0, st>>>((const __nv_bfloat16*)dout, out_ld_t, out_ld_b, pair, acts, cs,
                                                           w_hh_f, w_hh_r, lens, (__nv_bfloat16*)dgates, (int)T_, (int)B);
  else
    blstm_bwd_rg_kernel<false><<<grid, RG_THREADS, 0, st>>>((const __nv_bfloat16*)dout, out_ld_t, out_ld_b, pair, acts, cs,
                                                            w_hh_f, w_hh_r, lens, (__nv_bfloat16*)dgates, (int)T_, (int)B);
  B200ST_LAUNCH_CHECK("blstm_bwd_rg");
  return 0;
}

}  // namespace b200st
