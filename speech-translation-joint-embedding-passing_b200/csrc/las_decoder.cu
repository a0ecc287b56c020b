// Persistent LAS attention-LSTM decoder loop (forward): ONE launch runs all S decode steps of Dec.forward /
// forward_step / decode (reference models/Dec.py:130-233,320-438) for bf16 activations, decoder width D = 512,
// 3 uni-LSTM layers, key/value width 2H = 512 -- the shape Seq2seq constructs (Seq2seq.py:133-168).
//
// Why: a decode step is a chain of ~9 DEPENDENT small kernels (cell -> GEMM -> cell -> GEMM -> cell -> attention -> GEMM ->
// vocabulary GEMM -> arg-max), each 5-11 us of launch + prologue + drain latency for < 1 us of work: 31 steps x ~52 us.
// Here the chain stays inside one grid of 128 CTAs (one per SM, 20 SMs left to concurrent graph branches):
//   * CTA c owns hidden units [4c, 4c+4) of every LSTM layer.  Its 16 gate rows of [W_x | W_hh] of all three layers
//     (96 KB bf16) and its 4 rows of acous_ffn (8 KB) stay RESIDENT IN SHARED MEMORY for the whole loop; its 79 rows of the
//     vocabulary projection are streamed from L2 once per step.
//   * the activations every CTA needs (h of the layer below, the recurrent h, the context, the cell value: [B, 512] bf16)
//     are exchanged through L2-resident global buffers -- the same time-stacked buffers the backward pass reads -- and
//     streamed through a 4-stage cp.async ring as the A operand of mma.sync.m16n8k16 (bf16 in, fp32 accumulate).
//     Tiles are 64 x 16 per CTA: far below anything tcgen05 / TMEM pays off for (one UMMA commit round trip costs more than
//     the whole tile), so the warp-level MMA is the right instrument; the step is bound by the 6 grid-wide barriers.
//   * phases of a step, separated by a grid barrier (release/acquire on one global counter):
//       P0 layer 0: gates = gx0(token) + [cv_{s-1} | h0_{s-1}] W^T -> cell      (gx0 = row of E W_ih0[:, :E]^T + b0 gathered
//       P1 layer 1: [h0_s | h1_{s-1}] W^T + b -> cell, out = h1 + h0 (residual)   by the fed-back token, or GX0[s] when teacher forced)
//       P2 layer 2: [out1 | h2_{s-1}] W^T + b -> cell = dec_out
//       P3 attention: 2 CTAs per sequence: scores q.(W k_j), mask, softmax, context (each CTA half of the value columns)
//       P4 acous_ffn: cv = [ctx | dec_out] W_f^T            (the dynamic embedding of this step)
//       P5 vocabulary: logits = cv W_out^T + b (fp32 accumulators) -> per-row arg-max: atomicMax of an order-preserving
//          (value, ~index) key -> next step's token, EOS/PAD length rule (Dec.py:334-340)
// Saved for the (unchanged) backward pass: h / c / gate activations per layer and step, residual sums, contexts, attention
// probabilities, cell values.
#include "common.cuh"

namespace b200st {

constexpr int LP_D = 512;            // decoder width
constexpr int LP_NC = 128;           // CTAs
constexpr int LP_UPC = LP_D / LP_NC; // 4 hidden units per CTA and layer
constexpr int LP_T = 256;            // threads
constexpr int LP_KC = 64;            // K columns per staged chunk
constexpr int LP_STG = 4;            // cp.async stages
constexpr int LP_VR = 80;            // vocabulary rows per CTA slot (10 n-tiles of 8)
constexpr int LP_CROW = LP_KC * 2 + 16;          // 144 B: row stride of a staged chunk (conflict-free ldmatrix)
constexpr int LP_WROW = 2 * LP_D * 2 + 16;       // 2064 B: row stride of a resident weight row (K = 1024)
constexpr int LP_MAXB = 1024;

// shared memory map
constexpr int LP_WL_OFF = 0;                                   // [3][16][LP_WROW]
constexpr int LP_WF_OFF = LP_WL_OFF + 3 * 16 * LP_WROW;        // [8][LP_WROW] (rows 4..7 zero)
constexpr int LP_STA_OFF = LP_WF_OFF + 8 * LP_WROW;            // [STG][64][LP_CROW] activation chunks
constexpr int LP_STW_OFF = LP_STA_OFF + LP_STG * 64 * LP_CROW; // [STG][80][LP_CROW] vocabulary weight chunks
constexpr int LP_G_OFF = LP_STW_OFF + LP_STG * LP_VR * LP_CROW;   // float [64][20] gate / partial exchange
constexpr int LP_MISC_OFF = LP_G_OFF + 64 * 20 * 4;            // float [2048] scratch (attention, arg-max)
constexpr int LP_SYM_OFF = LP_MISC_OFF + 2048 * 4;             // int [LP_MAXB] tokens of this step
constexpr int LP_SMEM = LP_SYM_OFF + LP_MAXB * 4;

struct LasDecArgs {
  // step-invariant inputs
  const __nv_bfloat16* wk;        // [B, Tk, D]   projected keys W k_j (attention.py:192, hoisted)
  const __nv_bfloat16* enc;       // [B, Tk, 512] values
  const int32_t* klens;           // [B] valid keys or null
  const __nv_bfloat16* gx0;       // free running: TOK [V, 4D] = E W_ih0[:, :E]^T + b0;  teacher forcing: GX0 [S, B, 4D]
  const __nv_bfloat16* wx[3];     // input weights of layer i (layer 0: the cell-value columns W_ih0[:, E:]), row stride ldwx[i]
  int64_t ldwx[3];
  const __nv_bfloat16* whh[3];    // [4D, D]
  const float* bias[3];           // b_ih + b_hh of layers 1, 2 (layer 0's is inside gx0)
  const __nv_bfloat16* wffn;      // [D, 1024]
  const __nv_bfloat16* wout;      // [V, D]
  const float* bout;              // [V]
  // time-stacked state (L2 resident; also what backward reads)
  __nv_bfloat16* CV;              // [S+1, B, D], CV[0] = 0
  __nv_bfloat16* H[3];            // [S+1, B, D], H[i][0] = 0
  float* C[3];                    // [S+1, B, D], C[i][0] = 0
  float* ACT[3];                  // [S, B, 4D] post-activation gates
  __nv_bfloat16* RES1;            // [S, B, D] h1 + h0
  __nv_bfloat16* CTX;             // [S, B, 512]
  float* PROBS;                   // [S, B, Tk]
  __nv_bfloat16* LOGITS;          // [S, B, V] or null
  int64_t* SYM;                   // [S, B] chosen tokens
  int32_t* lengths;               // [B]
  unsigned long long* best;       // [2][B] arg-max keys (zero-filled)
  unsigned int* barrier;          // [1] (zero-filled)
  int B, Tk, S, V, teacher;
};

// Optional in-kernel timeline (debug aid, b200st_las_decoder_timeline): thread 0 of CTA 0 records %globaltimer (ns) at
// the phase boundaries of every step: [step][8] = start, after P0, P1, P2, P3, P4, P5(+barrier), unused.
__device__ long long* g_lp_timeline = nullptr;
__device__ __forceinline__ long long lp_now() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define LP_TL(i) do { if (tl) tl[s * 16 + (i)] = lp_now(); } while (0)

__device__ __forceinline__ uint32_t lp_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void lp_cp16(uint32_t dst, const void* src, bool ok) {
  const int n = ok ? 16 : 0;         // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void lp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void lp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void lp_ldsm4(uint32_t a, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void lp_ldsm2(uint32_t a, uint32_t (&r)[2]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a));
}
__device__ __forceinline__ void lp_mma(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ unsigned int lp_ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Grid-wide barrier: every CTA of the grid is resident (128 CTAs <= 148 SMs, one per SM by shared-memory footprint; any
// kernel sharing the GPU finishes independently of this grid, so late CTAs always arrive).
__device__ __forceinline__ void lp_grid_sync(unsigned int* counter, unsigned int& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    __threadfence();
    atomicAdd(counter, 1u);
    while (lp_ld_acquire(counter) < target) {}
  }
  __syncthreads();
}
__device__ __forceinline__ float lp_bf(const __nv_bfloat16* p) {      // L2-coherent scalar load (data written by other CTAs)
  unsigned short v;
  asm volatile("ld.global.cg.u16 %0, [%1];" : "=h"(v) : "l"(p));
  return __uint_as_float((uint32_t)v << 16);
}

// One A segment of a phase's K range: rows of a [B, ld] bf16 matrix in global memory.
struct LpSeg { const __nv_bfloat16* p; int64_t ld; };

// Stage chunk `kc` (64 K columns) of the A operand for batch rows [m0, m0+64) and, for the vocabulary phase, of the weight rows.
__device__ __forceinline__ void lp_issue(uint8_t* smem, int stage, int kc, const LpSeg& s0, const LpSeg& s1, int m0, int B,
                                         const __nv_bfloat16* wrows, int w0, int V) {
  const int tid = threadIdx.x;
  const LpSeg& sg = (kc < LP_D / LP_KC) ? s0 : s1;
  const int kofs = (kc % (LP_D / LP_KC)) * LP_KC;
  const uint32_t abase = lp_smem(smem + LP_STA_OFF + stage * 64 * LP_CROW);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int id = tid + i * LP_T;             // 512 x 16 B
    const int r = id >> 3, c16 = id & 7;
    const bool ok = m0 + r < B;
    lp_cp16(abase + r * LP_CROW + c16 * 16, sg.p + (int64_t)(ok ? m0 + r : 0) * sg.ld + kofs + c16 * 8, ok);
  }
  if (wrows) {
    const uint32_t wbase = lp_smem(smem + LP_STW_OFF + stage * LP_VR * LP_CROW);
    for (int id = tid; id < LP_VR * 8; id += LP_T) {
      const int r = id >> 3, c16 = id & 7;
      const bool ok = w0 + r < V;
      lp_cp16(wbase + r * LP_CROW + c16 * 16, wrows + (int64_t)(ok ? w0 + r : 0) * LP_D + kofs + c16 * 8, ok);
    }
  }
}

__device__ __forceinline__ void lp_wait_dyn(int n) {      // cp.async.wait_group with a run-time count (<= 8)
  switch (n) {
    case 0: lp_wait<0>(); break; case 1: lp_wait<1>(); break; case 2: lp_wait<2>(); break; case 3: lp_wait<3>(); break;
    case 4: lp_wait<4>(); break; case 5: lp_wait<5>(); break; case 6: lp_wait<6>(); break; default: lp_wait<7>(); break;
  }
}

// Resident-weight GEMM: acc (16 x 8 fp32 tile of m-tile `mt`, n-tile `nt`) += A[64 rows, chunks kc0 .. kc1) * W^T.
// The loop is bound by L2 latency, not bandwidth: ALL chunks of the range (<= 8 x 9 KB, the staging area doubles as 9 such
// slots outside the vocabulary phase) are requested up front and consumed as they land.  B operand from the resident weight
// block `wres` ([rows][LP_WROW], k index = global K position).  khalf < 0: the warp does all 4 k-steps of every chunk;
// khalf = 0 / 1: the first / last two (split-K across the two warp groups).  The k-steps alternate between two accumulators
// so that consecutive mma.sync do not form one dependent chain.
__device__ __forceinline__ void lp_gemm_res(uint8_t* smem, float (&acc)[1][4], int kc0, int kc1, const LpSeg& s0, const LpSeg& s1,
                                            int m0, int B, const uint8_t* wres, int nt, int mt, int khalf) {
  const int lane = threadIdx.x & 31;
  const int n = kc1 - kc0;
  float alt[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = 0; i < n; ++i) {
    lp_issue(smem, i, kc0 + i, s0, s1, m0, B, nullptr, 0, 0);
    lp_commit();
  }
  const uint32_t wbase = lp_smem(wres) + (nt * 8 + (lane & 7)) * LP_WROW + ((lane >> 3) & 1) * 16;
  const int k_lo = khalf < 0 ? 0 : khalf * 2, k_hi = khalf < 0 ? 4 : khalf * 2 + 2;
  for (int c = 0; c < n; ++c) {
    lp_wait_dyn(n - 1 - c);
    __syncthreads();
    const uint32_t abase = lp_smem(smem + LP_STA_OFF + c * 64 * LP_CROW) + (16 * mt + (lane & 15)) * LP_CROW + (lane >> 4) * 16;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      if (ks < k_lo || ks >= k_hi) continue;
      uint32_t a[4], b[2];
      lp_ldsm4(abase + ks * 32, a);
      lp_ldsm2(wbase + ((kc0 + c) * LP_KC + ks * 16) * 2, b);
      if (ks & 1) lp_mma(alt, a, b);
      else lp_mma(acc[0], a, b);
    }
  }
  __syncthreads();
#pragma unroll
  for (int e = 0; e < 4; ++e) acc[0][e] += alt[e];
}

// Vocabulary GEMM: acc[5] (m-tile mt, n-tiles nt0 .. nt0+4 of this CTA's 80-row slot) = cv[64, 512] * W_out[rows]^T, both
// operands streamed: 4 stages of (activation chunk 9 KB | weight chunk 11.25 KB).  The weight chunks do not depend on the
// other CTAs: the first four are requested BEFORE the grid barrier in front of this phase (lp_vocab_prefetch).
__device__ __forceinline__ void lp_issue_w(uint8_t* smem, int stage, int kc, const __nv_bfloat16* wrows, int w0, int wlim) {
  const uint32_t wbase = lp_smem(smem + LP_STW_OFF + stage * LP_VR * LP_CROW);
  for (int id = threadIdx.x; id < LP_VR * 8; id += LP_T) {
    const int r = id >> 3, c16 = id & 7;
    const bool ok = w0 + r < wlim;
    lp_cp16(wbase + r * LP_CROW + c16 * 16, wrows + (int64_t)(ok ? w0 + r : 0) * LP_D + kc * LP_KC + c16 * 8, ok);
  }
}
__device__ __forceinline__ void lp_vocab_prefetch(uint8_t* smem, const __nv_bfloat16* wrows, int w0, int wlim) {
  for (int i = 0; i < 4; ++i) { lp_issue_w(smem, i, i, wrows, w0, wlim); lp_commit(); }
}
__device__ __forceinline__ void lp_gemm_vocab(uint8_t* smem, float (&acc)[5][4], const LpSeg& s0, int m0, int B, int nt0, int mt,
                                              const __nv_bfloat16* wrows, int w0, int wlim) {
  const int lane = threadIdx.x & 31;
  for (int i = 0; i < 4; ++i) { lp_issue(smem, i, i, s0, s0, m0, B, nullptr, 0, 0); lp_commit(); }
  // groups so far: W0..W3 (prefetch), A0..A3; iteration c commits one more ((A, W) of chunk c + 3, or empty)
  for (int c = 0; c < 8; ++c) {
    if (c < 4) lp_wait<3>(); else lp_wait<2>();
    __syncthreads();
    if (c >= 1 && c + 3 < 8) {
      lp_issue(smem, (c + 3) & 3, c + 3, s0, s0, m0, B, nullptr, 0, 0);
      lp_issue_w(smem, (c + 3) & 3, c + 3, wrows, w0, wlim);
    }
    lp_commit();
    const int st = c & 3;
    const uint32_t abase = lp_smem(smem + LP_STA_OFF + st * 64 * LP_CROW) + (16 * mt + (lane & 15)) * LP_CROW + (lane >> 4) * 16;
    const uint32_t wb = lp_smem(smem + LP_STW_OFF + st * LP_VR * LP_CROW) + (nt0 * 8 + (lane & 7)) * LP_CROW + ((lane >> 3) & 1) * 16;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t a[4];
      lp_ldsm4(abase + ks * 32, a);
#pragma unroll
      for (int n = 0; n < 5; ++n) {
        uint32_t b[2];
        lp_ldsm2(wb + n * 8 * LP_CROW + ks * 32, b);
        lp_mma(acc[n], a, b);
      }
    }
  }
  lp_wait<0>();
  __syncthreads();
}

__device__ __forceinline__ unsigned long long lp_key(float v, int idx) {
  uint32_t u = __float_as_uint(v);
  u ^= (u >> 31) ? 0xFFFFFFFFu : 0x80000000u;                 // order-preserving map of fp32 onto uint32
  return ((unsigned long long)u << 32) | (uint32_t)(0xFFFFFFFFu - (uint32_t)idx);   // ties: the lowest index wins
}

// Split grid barrier: arrive as soon as this CTA's outputs of the phase are written, do work that does not depend on the
// other CTAs (the recurrent half of the next GEMM), then wait -- the ~1.4 us barrier round trip hides behind that work.
__device__ __forceinline__ void lp_arrive(unsigned int* counter, unsigned int& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    __threadfence();
    atomicAdd(counter, 1u);
  }
}
__device__ __forceinline__ void lp_wait_grid(unsigned int* counter, const unsigned int& target) {
  if (threadIdx.x == 0) {
    while (lp_ld_acquire(counter) < target) {}
  }
  __syncthreads();
}

constexpr int LP_MAXMB = 2;          // batch rows are processed in blocks of 64; accumulators of 2 blocks live in registers
#define LP_FOR_MB(mb) _Pragma("unroll") for (int mb = 0; mb < LP_MAXMB; ++mb) if (mb < nMB)

__global__ void __launch_bounds__(LP_T, 1) las_dec_fwd_persist_kernel(const LasDecArgs a) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ float red[32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = blockIdx.x;
  const int u0 = cta * LP_UPC;
  const int B = a.B, S = a.S, V = a.V, Tk = a.Tk;
  float* gs = reinterpret_cast<float*>(smem + LP_G_OFF);
  float* misc = reinterpret_cast<float*>(smem + LP_MISC_OFF);
  int* sym_s = reinterpret_cast<int*>(smem + LP_SYM_OFF);
  unsigned int bar_target = 0;
  const int64_t BD = (int64_t)B * LP_D;
  long long* tl = (cta == 0 && tid == 0) ? g_lp_timeline : nullptr;
  if (tl) tl[a.S * 16] = lp_now();

  // ---- resident weights: row r = gate * 4 + j of layer i  <-  [ W_x[gate*512 + u0 + j, :512] | W_hh[gate*512 + u0 + j, :512] ]
  for (int i = 0; i < 3; ++i)
    for (int id = tid; id < 16 * 128; id += LP_T) {
      const int r = id >> 7, c16 = id & 127;
      const int grow = (r >> 2) * LP_D + u0 + (r & 3);
      const __nv_bfloat16* src = (c16 < 64) ? a.wx[i] + (int64_t)grow * a.ldwx[i] + c16 * 8
                                            : a.whh[i] + (int64_t)grow * LP_D + (c16 - 64) * 8;
      uint4 v;
      if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        v = *reinterpret_cast<const uint4*>(src);
      } else {               // W_ih0[:, E:] is a column slice: rows are only 2-byte aligned in general
        __nv_bfloat16 t[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) t[e] = src[e];
        v = *reinterpret_cast<uint4*>(t);
      }
      *reinterpret_cast<uint4*>(smem + LP_WL_OFF + (i * 16 + r) * LP_WROW + c16 * 16) = v;
    }
  for (int id = tid; id < 8 * 128; id += LP_T) {
    const int r = id >> 7, c16 = id & 127;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < LP_UPC) v = *reinterpret_cast<const uint4*>(a.wffn + (int64_t)(u0 + r) * (2 * LP_D) + c16 * 8);
    *reinterpret_cast<uint4*>(smem + LP_WF_OFF + r * LP_WROW + c16 * 16) = v;
  }
  __syncthreads();

  const int vper = (V + LP_NC - 1) / LP_NC;
  const int w0 = cta * vper;                                   // this CTA's vocabulary rows [w0, w0 + vrows)
  const int vrows = max(0, min(vper, V - w0));
  const int nMB = (B + 63) / 64;
  // per-thread constants of the cell phase: thread = (row bl of the 64-row block, unit j)
  const int bl = tid >> 2, j = tid & 3;
  float bias1[4], bias2[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) { bias1[g] = a.bias[1][g * LP_D + u0 + j]; bias2[g] = a.bias[2][g * LP_D + u0 + j]; }
  // vocabulary phase: this thread's columns (n-tile n, pair element e) -> bias values, loaded once
  const int nt0v = 5 * (warp >> 2);
  float bov[5][2];
#pragma unroll
  for (int n = 0; n < 5; ++n)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int col = (nt0v + n) * 8 + 2 * (lane & 3) + e;
      bov[n][e] = col < vrows ? a.bout[w0 + col] : 0.f;
    }
  float cst[3][LP_MAXMB];                 // cell state of (layer, row block): lives in registers for the whole loop
  float h0reg[LP_MAXMB];                  // h of layer 0 of this step (the residual input of layer 1, Dec.py:417-418)
  float pre[LP_MAXMB][1][4];              // GEMM accumulators carried across a barrier
#pragma unroll
  for (int mb = 0; mb < LP_MAXMB; ++mb) {
    cst[0][mb] = cst[1][mb] = cst[2][mb] = 0.f; h0reg[mb] = 0.f;
    pre[mb][0][0] = pre[mb][0][1] = pre[mb][0][2] = pre[mb][0][3] = 0.f;      // step 0: cv_{-1} = h_{-1} = 0
  }

  // cell update of `layer` for row block mb from the accumulators (fragment -> gs exchange -> gates of one unit per thread)
  auto cell = [&](int layer, int s, int mb, const float (&acc)[1][4], const float (&add)[4]) {
    {
      const int r = 16 * (warp & 3) + (lane >> 2), c = 8 * (warp >> 2) + 2 * (lane & 3);
      gs[r * 20 + c] = acc[0][0]; gs[r * 20 + c + 1] = acc[0][1];
      gs[(r + 8) * 20 + c] = acc[0][2]; gs[(r + 8) * 20 + c + 1] = acc[0][3];
    }
    __syncthreads();
    const int bg = mb * 64 + bl;
    if (bg < B) {
      const float pi = gs[bl * 20 + j] + add[0], pf = gs[bl * 20 + 4 + j] + add[1];
      const float pg = gs[bl * 20 + 8 + j] + add[2], po = gs[bl * 20 + 12 + j] + add[3];
      const float i_ = sigmoidf_(pi), f_ = sigmoidf_(pf), g_ = tanhf(pg), o_ = sigmoidf_(po);
      const float cn = f_ * cst[layer][mb] + i_ * g_;
      const float hn = o_ * tanhf(cn);
      cst[layer][mb] = cn;
      const int64_t o1 = (int64_t)(s + 1) * BD + (int64_t)bg * LP_D + u0 + j;
      a.C[layer][o1] = cn;
      a.H[layer][o1] = __float2bfloat16_rn(hn);
      float* act = a.ACT[layer] + ((int64_t)s * B + bg) * (4 * LP_D) + u0 + j;
      act[0] = i_; act[LP_D] = f_; act[2 * LP_D] = g_; act[3 * LP_D] = o_;
      if (layer == 0) h0reg[mb] = __bfloat162float(__float2bfloat16_rn(hn));
      if (layer == 1)     // Dec.py:417-418: out = h + x on the middle layer
        a.RES1[(int64_t)s * BD + (int64_t)bg * LP_D + u0 + j] = __float2bfloat16_rn(hn + h0reg[mb]);
    }
    __syncthreads();
  };

  for (int s = 0; s < S; ++s) {
    LP_TL(0);
    unsigned long long* best_prev = a.best + (size_t)((s + 1) & 1) * B;     // written in the vocabulary phase of step s-1
    unsigned long long* best_cur = a.best + (size_t)(s & 1) * B;
    // ---------------------------------------------------------------- tokens fed at this step
    if (!a.teacher || cta == 0) {
      for (int b = tid; b < B; b += LP_T) {
        int sym = 2;  // BOS (Dec.py:158-160,199)
        if (s > 0) {
          unsigned long long key;
          asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(key) : "l"(best_prev + b));
          sym = (int)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull));
          if (cta == 0) {                                                    // Dec.py:331 + 334-340
            a.SYM[(int64_t)(s - 1) * B + b] = sym;
            if ((sym == 3 || sym == 0) && a.lengths[b] > s - 1) a.lengths[b] = s;
          }
        }
        sym_s[b] = sym;
      }
    }
    __syncthreads();
    // ---------------------------------------------------------------- layer 0: its whole GEMM ([cv_{s-1} | h0_{s-1}] W^T) was
    // accumulated while waiting for the previous step's barriers; what is left is the token's gate row and the cell
    LP_FOR_MB(mb) {
      const int bg = mb * 64 + bl;
      float add[4] = {0.f, 0.f, 0.f, 0.f};
      if (bg < B) {
        const __nv_bfloat16* g0 = a.teacher ? a.gx0 + ((int64_t)s * B + bg) * (4 * LP_D) : a.gx0 + (int64_t)sym_s[bg] * (4 * LP_D);
#pragma unroll
        for (int g = 0; g < 4; ++g) add[g] = __bfloat162float(g0[g * LP_D + u0 + j]);
      }
      cell(0, s, mb, pre[mb], add);
    }
    lp_arrive(a.barrier, bar_target);
    {   // recurrent half of layer 1 (h1_{s-1}: chunks 8..15) behind the barrier
      const LpSeg s0 = {a.H[0] + (int64_t)(s + 1) * BD, LP_D}, s1 = {a.H[1] + (int64_t)s * BD, LP_D};
      LP_FOR_MB(mb) {
        pre[mb][0][0] = pre[mb][0][1] = pre[mb][0][2] = pre[mb][0][3] = 0.f;
        lp_gemm_res(smem, pre[mb], 8, 16, s0, s1, mb * 64, B, smem + LP_WL_OFF + 16 * LP_WROW, warp >> 2, warp & 3, -1);
      }
      lp_wait_grid(a.barrier, bar_target);
      LP_TL(1);
      LP_FOR_MB(mb) {
        lp_gemm_res(smem, pre[mb], 0, 8, s0, s1, mb * 64, B, smem + LP_WL_OFF + 16 * LP_WROW, warp >> 2, warp & 3, -1);
        LP_TL(8);
        cell(1, s, mb, pre[mb], bias1);
        LP_TL(9);
      }
    }
    lp_arrive(a.barrier, bar_target);
    LP_TL(10);
    {   // layer 2
      const LpSeg s0 = {a.RES1 + (int64_t)s * BD, LP_D}, s1 = {a.H[2] + (int64_t)s * BD, LP_D};
      LP_FOR_MB(mb) {
        pre[mb][0][0] = pre[mb][0][1] = pre[mb][0][2] = pre[mb][0][3] = 0.f;
        lp_gemm_res(smem, pre[mb], 8, 16, s0, s1, mb * 64, B, smem + LP_WL_OFF + 32 * LP_WROW, warp >> 2, warp & 3, -1);
      }
      LP_TL(11);
      lp_wait_grid(a.barrier, bar_target);
      LP_TL(2);
      LP_FOR_MB(mb) {
        lp_gemm_res(smem, pre[mb], 0, 8, s0, s1, mb * 64, B, smem + LP_WL_OFF + 32 * LP_WROW, warp >> 2, warp & 3, -1);
        cell(2, s, mb, pre[mb], bias2);
      }
    }
    lp_arrive(a.barrier, bar_target);
    if (cta == 0)       // recycle the arg-max slots of step s+1 (last read in this step's token fetch, two barriers ago)
      for (int b = tid; b < B; b += LP_T) (a.best + (size_t)((s + 1) & 1) * B)[b] = 0ull;
    lp_wait_grid(a.barrier, bar_target);
    LP_TL(3);
    // ---------------------------------------------------------------- bilinear attention (attention.py:190-193,250-273)
    {
      const __nv_bfloat16* dec_out = a.H[2] + (int64_t)(s + 1) * BD;
      float* qs = misc;                 // [512]
      float* sc = misc + 512;           // [<= 512] scores / probabilities
      float* part = misc + 1024;        // [8][128] context partial sums... laid out [kg][256] with 4 key groups x 2 halves
      for (int item = cta; item < 2 * B; item += LP_NC) {
        const int b = item >> 1, half = item & 1;
        for (int c = tid; c < LP_D; c += LP_T) qs[c] = lp_bf(dec_out + (int64_t)b * LP_D + c);
        __syncthreads();
        const int klen = a.klens ? a.klens[b] : Tk;
        const __nv_bfloat16* wkb = a.wk + (int64_t)b * Tk * LP_D;
        for (int j0 = warp; j0 < Tk; j0 += 8 * 8) {     // 8 keys of this warp in flight (16 x 16 B per lane)
          uint4 v[8][2];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int jj = j0 + q * 8;
            if (jj < Tk) {
              v[q][0] = *reinterpret_cast<const uint4*>(wkb + (int64_t)jj * LP_D + lane * 8);
              v[q][1] = *reinterpret_cast<const uint4*>(wkb + (int64_t)jj * LP_D + 256 + lane * 8);
            }
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int jj = j0 + q * 8;
            if (jj < Tk) {
              float sum = 0.f;
#pragma unroll
              for (int h2 = 0; h2 < 2; ++h2) {
                const uint32_t w4[4] = {v[q][h2].x, v[q][h2].y, v[q][h2].z, v[q][h2].w};
                const float* qq = qs + h2 * 256 + lane * 8;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  sum = fmaf(qq[2 * e], __uint_as_float(w4[e] << 16), sum);
                  sum = fmaf(qq[2 * e + 1], __uint_as_float(w4[e] & 0xffff0000u), sum);
                }
              }
              sum = warp_sum(sum);
              if (lane == 0) sc[jj] = (jj >= klen) ? -1e12f : sum;
            }
          }
        }
        __syncthreads();
        float mx = -INFINITY;
        for (int jj = tid; jj < Tk; jj += LP_T) mx = fmaxf(mx, sc[jj]);
        mx = block_max(mx, red);
        float sum = 0.f;
        for (int jj = tid; jj < Tk; jj += LP_T) { const float e = expf(sc[jj] - mx); sc[jj] = e; sum += e; }
        sum = block_sum(sum, red);
        const float inv = 1.f / sum;
        __syncthreads();
        for (int jj = tid; jj < Tk; jj += LP_T) {
          const float pv = sc[jj] * inv;
          sc[jj] = pv;
          if (half == 0) a.PROBS[((int64_t)s * B + b) * Tk + jj] = pv;
        }
        __syncthreads();
        {   // context columns [256 half, 256 half + 256): thread = (key group kg of 8, 8 columns), 8 loads of 16 B in flight
          const int kg = tid >> 5, c8 = (tid & 31) * 8;
          const __nv_bfloat16* vb = a.enc + (int64_t)b * Tk * LP_D + half * 256 + c8;
          float x[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          for (int jj = kg; jj < Tk; jj += 64) {
            uint4 v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (jj + 8 * q < Tk) v[q] = *reinterpret_cast<const uint4*>(vb + (int64_t)(jj + 8 * q) * LP_D);
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (jj + 8 * q < Tk) {
                const float pv = sc[jj + 8 * q];
                const uint32_t w4[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  x[2 * e] = fmaf(pv, __uint_as_float(w4[e] << 16), x[2 * e]);
                  x[2 * e + 1] = fmaf(pv, __uint_as_float(w4[e] & 0xffff0000u), x[2 * e + 1]);
                }
              }
          }
          // reduce the 8 key groups: two rounds through `part` [4][256]
          if (kg >= 4) {
#pragma unroll
            for (int e = 0; e < 8; ++e) part[(kg - 4) * 256 + c8 + e] = x[e];
          }
          __syncthreads();
          if (kg < 4) {
#pragma unroll
            for (int e = 0; e < 8; ++e) part[kg * 256 + c8 + e] += x[e];
          }
        }
        __syncthreads();
        a.CTX[((int64_t)s * B + b) * LP_D + half * 256 + tid] =
            __float2bfloat16_rn((part[tid] + part[256 + tid]) + (part[512 + tid] + part[768 + tid]));
        __syncthreads();
      }
    }
    LP_TL(7);
    lp_arrive(a.barrier, bar_target);
    // ---------------------------------------------------------------- cell_value = acous_ffn(cat(context, dec_out)) (Dec.py:431-433):
    // the dec_out half (chunks 8..15) behind the barrier, the context half after it
    {
      const LpSeg s0 = {a.CTX + (int64_t)s * BD, LP_D}, s1 = {a.H[2] + (int64_t)(s + 1) * BD, LP_D};
      LP_FOR_MB(mb) {
        pre[mb][0][0] = pre[mb][0][1] = pre[mb][0][2] = pre[mb][0][3] = 0.f;
        lp_gemm_res(smem, pre[mb], 8, 16, s0, s1, mb * 64, B, smem + LP_WF_OFF, 0, warp & 3, warp >> 2);
      }
      lp_wait_grid(a.barrier, bar_target);
      LP_TL(4);
      LP_FOR_MB(mb) {
        const int m0 = mb * 64;
        lp_gemm_res(smem, pre[mb], 0, 8, s0, s1, m0, B, smem + LP_WF_OFF, 0, warp & 3, warp >> 2);
        const int r = 16 * (warp & 3) + (lane >> 2), c = 2 * (lane & 3);
        if (warp >= 4) {
          gs[r * 20 + c] = pre[mb][0][0]; gs[r * 20 + c + 1] = pre[mb][0][1];
          gs[(r + 8) * 20 + c] = pre[mb][0][2]; gs[(r + 8) * 20 + c + 1] = pre[mb][0][3];
        }
        __syncthreads();
        if (warp < 4 && c < LP_UPC) {
          __nv_bfloat16* cv = a.CV + (int64_t)(s + 1) * BD;
          if (m0 + r < B)
            *reinterpret_cast<__nv_bfloat162*>(cv + (int64_t)(m0 + r) * LP_D + u0 + c) =
                __floats2bfloat162_rn(pre[mb][0][0] + gs[r * 20 + c], pre[mb][0][1] + gs[r * 20 + c + 1]);
          if (m0 + r + 8 < B)
            *reinterpret_cast<__nv_bfloat162*>(cv + (int64_t)(m0 + r + 8) * LP_D + u0 + c) =
                __floats2bfloat162_rn(pre[mb][0][2] + gs[(r + 8) * 20 + c], pre[mb][0][3] + gs[(r + 8) * 20 + c + 1]);
        }
        __syncthreads();
      }
    }
    lp_arrive(a.barrier, bar_target);
    // next step's layer 0, recurrent half (h0_s: complete since the first barrier of this step) behind the barrier
    const LpSeg l0s0 = {a.CV + (int64_t)(s + 1) * BD, LP_D}, l0s1 = {a.H[0] + (int64_t)(s + 1) * BD, LP_D};
    if (s + 1 < S)
      LP_FOR_MB(mb) {
        pre[mb][0][0] = pre[mb][0][1] = pre[mb][0][2] = pre[mb][0][3] = 0.f;
        lp_gemm_res(smem, pre[mb], 8, 16, l0s0, l0s1, mb * 64, B, smem + LP_WL_OFF, warp >> 2, warp & 3, -1);
      }
    if (vrows > 0) lp_vocab_prefetch(smem, a.wout, w0, w0 + vrows);     // weight chunks 0..3: independent of the barrier
    lp_wait_grid(a.barrier, bar_target);
    LP_TL(5);
    // ---------------------------------------------------------------- vocabulary projection + arg-max (Dec.py:434-436, 331)
    LP_FOR_MB(mb) {
      const int m0 = mb * 64;
      float acc[5][4];
#pragma unroll
      for (int n = 0; n < 5; ++n) { acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f; }
      if (vrows > 0) {
        if (mb > 0) lp_vocab_prefetch(smem, a.wout, w0, w0 + vrows);
        lp_gemm_vocab(smem, acc, l0s0, m0, B, nt0v, warp & 3, a.wout, w0, w0 + vrows);
      }
      const int r = 16 * (warp & 3) + (lane >> 2);
      float mx[2] = {-INFINITY, -INFINITY};
      int mi[2] = {0x7fffffff, 0x7fffffff};
#pragma unroll
      for (int n = 0; n < 5; ++n) {
        const int col = (nt0v + n) * 8 + 2 * (lane & 3);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          if (col + e < vrows) {
            const float v0 = acc[n][e] + bov[n][e], v1 = acc[n][2 + e] + bov[n][e];
            if (v0 > mx[0]) { mx[0] = v0; mi[0] = w0 + col + e; }
            if (v1 > mx[1]) { mx[1] = v1; mi[1] = w0 + col + e; }
            if (a.LOGITS) {
              if (m0 + r < B) a.LOGITS[((int64_t)s * B + m0 + r) * V + w0 + col + e] = __float2bfloat16_rn(v0);
              if (m0 + r + 8 < B) a.LOGITS[((int64_t)s * B + m0 + r + 8) * V + w0 + col + e] = __float2bfloat16_rn(v1);
            }
          }
        }
      }
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
#pragma unroll
        for (int o = 1; o <= 2; o <<= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, mx[h2], o);
          const int oi = __shfl_xor_sync(0xffffffffu, mi[h2], o);
          if (ov > mx[h2] || (ov == mx[h2] && oi < mi[h2])) { mx[h2] = ov; mi[h2] = oi; }
        }
        const int row = m0 + r + 8 * h2;
        if ((lane & 3) == 0 && row < B && mi[h2] != 0x7fffffff) atomicMax(best_cur + row, lp_key(mx[h2], mi[h2]));
      }
    }
    lp_arrive(a.barrier, bar_target);
    // next step's layer 0, cell-value half (cv_s: complete since the barrier before the vocabulary phase)
    if (s + 1 < S)
      LP_FOR_MB(mb)
        lp_gemm_res(smem, pre[mb], 0, 8, l0s0, l0s1, mb * 64, B, smem + LP_WL_OFF, warp >> 2, warp & 3, -1);
    lp_wait_grid(a.barrier, bar_target);
    LP_TL(6);
  }
  // ---- tokens of the last step
  if (cta == 0 && S > 0) {
    unsigned long long* best_last = a.best + (size_t)((S - 1) & 1) * B;
    for (int b = tid; b < B; b += LP_T) {
      unsigned long long key;
      asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(key) : "l"(best_last + b));
      const int sym = (int)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull));
      a.SYM[(int64_t)(S - 1) * B + b] = sym;
      if ((sym == 3 || sym == 0) && a.lengths[b] > S - 1) a.lengths[b] = S;
    }
  }
}

}  // namespace b200st

using namespace b200st;

extern "C" int b200st_las_decoder_timeline(void* buf) {     // int64 [S * 8 + 1] device buffer, or NULL to switch off
  long long* p = (long long*)buf;
  B200ST_CUDA(cudaMemcpyToSymbol(g_lp_timeline, &p, sizeof(p)));
  return 0;
}

// args: host array of 42 int64 slots (pointers and sizes), see include/b200st.h
extern "C" int b200st_las_decoder_fwd(const int64_t* v, int64_t n, b200st_stream_t stream) {
  if (n != 42) return set_error("las_decoder_fwd: expected 42 argument slots, got %lld", (long long)n);
  LasDecArgs a;
  int i = 0;
  a.wk = (const __nv_bfloat16*)v[i++]; a.enc = (const __nv_bfloat16*)v[i++]; a.klens = (const int32_t*)v[i++];
  a.gx0 = (const __nv_bfloat16*)v[i++];
  for (int l = 0; l < 3; ++l) { a.wx[l] = (const __nv_bfloat16*)v[i++]; a.ldwx[l] = v[i++]; a.whh[l] = (const __nv_bfloat16*)v[i++]; a.bias[l] = (const float*)v[i++]; }
  a.wffn = (const __nv_bfloat16*)v[i++]; a.wout = (const __nv_bfloat16*)v[i++]; a.bout = (const float*)v[i++];
  a.CV = (__nv_bfloat16*)v[i++];
  for (int l = 0; l < 3; ++l) { a.H[l] = (__nv_bfloat16*)v[i++]; a.C[l] = (float*)v[i++]; a.ACT[l] = (float*)v[i++]; }
  a.RES1 = (__nv_bfloat16*)v[i++]; a.CTX = (__nv_bfloat16*)v[i++]; a.PROBS = (float*)v[i++]; a.LOGITS = (__nv_bfloat16*)v[i++];
  a.SYM = (int64_t*)v[i++]; a.lengths = (int32_t*)v[i++]; a.best = (unsigned long long*)v[i++]; a.barrier = (unsigned int*)v[i++];
  a.B = (int)v[i++]; a.Tk = (int)v[i++]; a.S = (int)v[i++]; a.V = (int)v[i++]; a.teacher = (int)v[i++];
  if (i != 42) return set_error("las_decoder_fwd: internal slot count %d", i);
  if (a.B <= 0 || a.S <= 0) return 0;
  if (a.B > 64 * LP_MAXMB) return set_error("las_decoder_fwd: batch %d > %d", a.B, 64 * LP_MAXMB);
  if (a.Tk < 1 || a.Tk > 512) return set_error("las_decoder_fwd: Tk %d out of range [1, 512]", a.Tk);
  if ((a.V + LP_NC - 1) / LP_NC > LP_VR) return set_error("las_decoder_fwd: vocabulary %d > %d", a.V, LP_NC * LP_VR);
  if (!a.teacher && !a.gx0) return set_error("las_decoder_fwd: free running needs the token gate table");
  int dev = 0, sms = 0;
  B200ST_CUDA(cudaGetDevice(&dev));
  B200ST_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (sms < LP_NC) return set_error("las_decoder_fwd: needs %d co-resident CTAs, device has %d SMs", LP_NC, sms);
  static bool attr_set = false;
  if (!attr_set) {
    B200ST_CUDA(cudaFuncSetAttribute(las_dec_fwd_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LP_SMEM));
    attr_set = true;
  }
  B200ST_CUDA(cudaMemsetAsync(a.barrier, 0, sizeof(unsigned int), (cudaStream_t)stream));
  B200ST_CUDA(cudaMemsetAsync(a.best, 0, sizeof(unsigned long long) * 2 * a.B, (cudaStream_t)stream));
  las_dec_fwd_persist_kernel<<<LP_NC, LP_T, LP_SMEM, (cudaStream_t)stream>>>(a);
  B200ST_LAUNCH_CHECK("las_decoder_fwd");
  return 0;
}
