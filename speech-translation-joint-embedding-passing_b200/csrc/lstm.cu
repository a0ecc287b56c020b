// LSTM kernels.
//   * lstm_cell_fwd / bwd: fused gate non-linearities + state update for one step (LAS decoder, Dec.py:393-419)
//   * blstm_fwd / bwd: the packed bidirectional recurrence of the pyramidal acoustic encoder
//     (Enc.py:150-211) as ONE persistent kernel per layer.
//
// Persistent recurrence design (exact fp32 variant):
//   grid = (C, G, 2): a thread-block cluster of C CTAs owns one (direction, batch group of NB sequences).
//   CTA `rank` keeps the W_hh rows of its H/C hidden units (all four gates) resident in shared memory for
//   the whole sequence, so per step only h_{t-1} (NB x H) moves: every CTA computes its slice of the gates,
//   applies the cell update, and scatters its slice of h_t into the *next-step* h buffer of every CTA in
//   the cluster through distributed shared memory; one cluster barrier per time step.
//   Sequences are independent, so batch groups never synchronise with each other.
//   Variable lengths follow PackedSequence semantics: steps with t >= len leave the state untouched and
//   emit zeros; the reverse direction therefore starts at each sequence's own last frame.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace b200st {

// ------------------------------------------------------------------------------------------------
// single-step cell
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void lstm_cell_fwd_kernel(const T* __restrict__ gates, const T* __restrict__ gates_b,
                                     const T* __restrict__ gates_c, const float* __restrict__ c_prev,
                                     T* __restrict__ h, float* __restrict__ c, float* __restrict__ acts,
                                     const T* __restrict__ residual, T* __restrict__ out_res, int64_t B,
                                     int H) {
  pdl_wait();
  pdl_launch_dependents();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * H) return;
  const int64_t b = idx / H;
  const int u = (int)(idx % H);
  const T* g = gates + b * 4 * H;
  float pi = to_f(g[u]), pf = to_f(g[H + u]), pg = to_f(g[2 * H + u]), po = to_f(g[3 * H + u]);
  if (gates_b) {   // pre-activations may arrive as partial products computed concurrently (x W_ih^T, h W_hh^T, ...)
    const T* q = gates_b + b * 4 * H;
    pi += to_f(q[u]); pf += to_f(q[H + u]); pg += to_f(q[2 * H + u]); po += to_f(q[3 * H + u]);
  }
  if (gates_c) {
    const T* q = gates_c + b * 4 * H;
    pi += to_f(q[u]); pf += to_f(q[H + u]); pg += to_f(q[2 * H + u]); po += to_f(q[3 * H + u]);
  }
  const float i_ = sigmoidf_(pi);
  const float f_ = sigmoidf_(pf);
  const float g_ = tanhf(pg);
  const float o_ = sigmoidf_(po);
  const float cp = c_prev ? c_prev[idx] : 0.f;
  const float cn = f_ * cp + i_ * g_;
  const float hn = o_ * tanhf(cn);
  c[idx] = cn;
  h[idx] = from_f<T>(hn);
  if (acts) {
    float* a = acts + b * 4 * H;
    a[u] = i_; a[H + u] = f_; a[2 * H + u] = g_; a[3 * H + u] = o_;
  }
  if (out_res) out_res[idx] = from_f<T>(hn + to_f(residual[idx]));
}

template <typename T>
__global__ void lstm_cell_bwd_kernel(const T* __restrict__ dh_a, const T* __restrict__ dh_b,
                                     const T* __restrict__ dh_c, const float* __restrict__ dc_next,
                                     const float* __restrict__ acts, const float* __restrict__ c_prev,
                                     const float* __restrict__ c, T* __restrict__ dgates,
                                     float* __restrict__ dc_prev, int64_t B, int H) {
  pdl_wait();
  pdl_launch_dependents();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * H) return;
  const int64_t b = idx / H;
  const int u = (int)(idx % H);
  float dh = 0.f;
  if (dh_a) dh += to_f(dh_a[idx]);
  if (dh_b) dh += to_f(dh_b[idx]);
  if (dh_c) dh += to_f(dh_c[idx]);
  const float* a = acts + b * 4 * H;
  const float i_ = a[u], f_ = a[H + u], g_ = a[2 * H + u], o_ = a[3 * H + u];
  const float tc = tanhf(c[idx]);
  const float cp = c_prev ? c_prev[idx] : 0.f;
  const float dc = (dc_next ? dc_next[idx] : 0.f) + dh * o_ * (1.f - tc * tc);
  T* dg = dgates + b * 4 * H;
  dg[u] = from_f<T>(dc * g_ * i_ * (1.f - i_));
  dg[H + u] = from_f<T>(dc * cp * f_ * (1.f - f_));
  dg[2 * H + u] = from_f<T>(dc * i_ * (1.f - g_ * g_));
  dg[3 * H + u] = from_f<T>(dh * tc * o_ * (1.f - o_));
  dc_prev[idx] = dc * f_;
}

// ------------------------------------------------------------------------------------------------
// persistent bidirectional recurrence, forward
// ------------------------------------------------------------------------------------------------
constexpr int BLSTM_THREADS = 256;

template <typename T, int NB>
__global__ void __launch_bounds__(BLSTM_THREADS, 1)
blstm_fwd_kernel(const T* __restrict__ xproj, const float* __restrict__ w_hh_f,
                 const float* __restrict__ w_hh_r, const int32_t* __restrict__ lens, T* __restrict__ out,
                 int64_t out_ld_t, int64_t out_ld_b, int pair, T* __restrict__ hs,
                 float* __restrict__ acts, float* __restrict__ cs, int Tn, int B, int H, int C) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int grp = blockIdx.y, dir = blockIdx.z;
  const int tid = threadIdx.x;
  const int UPC = H / C, R = 4 * UPC, WP = H + 4;
  extern __shared__ __align__(16) float smem[];
  float* Ws = smem;                    // [R][WP]   resident recurrent weights of this CTA's units
  float* hbuf = Ws + (size_t)R * WP;   // [2][NB][H] h_{t-1} (double buffered, written by all CTAs)
  float* gbuf = hbuf + 2 * NB * H;     // [R][NB]   gate pre-activations of this step
  float* cbuf = gbuf + R * NB;         // [UPC][NB] cell state

  const float* w = dir ? w_hh_r : w_hh_f;
  for (int idx = tid; idx < R * H; idx += BLSTM_THREADS) {
    const int lr = idx / H, k = idx % H;
    const int g = lr / UPC, ul = lr % UPC;
    Ws[(size_t)lr * WP + k] = w[(size_t)(g * H + rank * UPC + ul) * H + k];
  }
  for (int idx = tid; idx < 2 * NB * H; idx += BLSTM_THREADS) hbuf[idx] = 0.f;
  for (int idx = tid; idx < UPC * NB; idx += BLSTM_THREADS) cbuf[idx] = 0.f;
  cluster.sync();

  const int b0 = grp * NB;
  const int64_t G4 = 4 * (int64_t)H;
  int cur = 0;
  for (int s = 0; s < Tn; ++s) {
    const int t = dir ? (Tn - 1 - s) : s;
    const T* xp = xproj + ((int64_t)dir * Tn + t) * B * G4;
    // ---- gates = xproj + W_hh h_{t-1}: one (gate row, 4 sequences) item per thread
    for (int item = tid; item < R * (NB / 4); item += BLSTM_THREADS) {
      const int lr = item / (NB / 4), bq = item % (NB / 4);
      const int g = lr / UPC, ul = lr % UPC;
      const int grow = g * H + rank * UPC + ul;
      float x[4];
#pragma unroll
      for (int qd = 0; qd < 4; ++qd) {
        const int b = b0 + bq * 4 + qd;
        x[qd] = (b < B) ? to_f(xp[(int64_t)b * G4 + grow]) : 0.f;
      }
      const float* wrow = Ws + (size_t)lr * WP;
      const float* h0 = hbuf + (size_t)(cur * NB + bq * 4) * H;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
      for (int k = 0; k < H; k += 4) {
        const float4 wv = *reinterpret_cast<const float4*>(wrow + k);
        const float4 v0 = *reinterpret_cast<const float4*>(h0 + k);
        const float4 v1 = *reinterpret_cast<const float4*>(h0 + H + k);
        const float4 v2 = *reinterpret_cast<const float4*>(h0 + 2 * H + k);
        const float4 v3 = *reinterpret_cast<const float4*>(h0 + 3 * H + k);
        a0 = fmaf(wv.x, v0.x, a0); a0 = fmaf(wv.y, v0.y, a0); a0 = fmaf(wv.z, v0.z, a0); a0 = fmaf(wv.w, v0.w, a0);
        a1 = fmaf(wv.x, v1.x, a1); a1 = fmaf(wv.y, v1.y, a1); a1 = fmaf(wv.z, v1.z, a1); a1 = fmaf(wv.w, v1.w, a1);
        a2 = fmaf(wv.x, v2.x, a2); a2 = fmaf(wv.y, v2.y, a2); a2 = fmaf(wv.z, v2.z, a2); a2 = fmaf(wv.w, v2.w, a2);
        a3 = fmaf(wv.x, v3.x, a3); a3 = fmaf(wv.y, v3.y, a3); a3 = fmaf(wv.z, v3.z, a3); a3 = fmaf(wv.w, v3.w, a3);
      }
      float* gb = gbuf + lr * NB + bq * 4;
      gb[0] = a0 + x[0]; gb[1] = a1 + x[1]; gb[2] = a2 + x[2]; gb[3] = a3 + x[3];
    }
    __syncthreads();
    // ---- cell update for this CTA's units; scatter h_t to every CTA of the cluster
    const int nxt = cur ^ 1;
    for (int item = tid; item < UPC * NB; item += BLSTM_THREADS) {
      const int ul = item / NB, bl = item % NB;
      const int b = b0 + bl;
      if (b >= B) continue;
      const int u = rank * UPC + ul;
      const bool valid = t < lens[b];
      const float c_old = cbuf[ul * NB + bl];
      const float h_old = hbuf[(size_t)(cur * NB + bl) * H + u];
      const float i_ = sigmoidf_(gbuf[(0 * UPC + ul) * NB + bl]);
      const float f_ = sigmoidf_(gbuf[(1 * UPC + ul) * NB + bl]);
      const float g_ = tanhf(gbuf[(2 * UPC + ul) * NB + bl]);
      const float o_ = sigmoidf_(gbuf[(3 * UPC + ul) * NB + bl]);
      const float c_new = valid ? fmaf(f_, c_old, i_ * g_) : c_old;
      const float h_new = valid ? o_ * tanhf(c_new) : h_old;
      cbuf[ul * NB + bl] = c_new;
      const int64_t row = ((int64_t)dir * Tn + t) * B + b;
      if (acts) {
        float* a = acts + row * G4;
        a[u] = i_; a[H + u] = f_; a[2 * H + u] = g_; a[3 * H + u] = o_;
      }
      if (cs) cs[row * H + u] = c_new;
      const T ho = from_f<T>(valid ? h_new : 0.f);
      out[(int64_t)(t / pair) * out_ld_t + (int64_t)b * out_ld_b + (int64_t)(t % pair) * 2 * H + dir * H + u] = ho;
      if (hs) hs[(((int64_t)dir * (Tn + 1) + (dir ? t : t + 1)) * B + b) * H + u] = ho;
      for (int r = 0; r < C; ++r) {
        float* dst = cluster.map_shared_rank(hbuf, r);
        dst[(size_t)(nxt * NB + bl) * H + u] = h_new;
      }
    }
    cluster.sync();
    cur = nxt;
  }
}

// ------------------------------------------------------------------------------------------------
// persistent bidirectional recurrence, backward (BPTT).  CTA `rank` keeps W_hh[:, its units] (the
// transposed slice) resident; per step it forms the gate gradients of its units, all-gathers them across
// the cluster (NB x 4H), and reduces its slice of dh_{t-1} = dG_t W_hh.
// ------------------------------------------------------------------------------------------------
template <typename T, int NB>
__global__ void __launch_bounds__(BLSTM_THREADS, 1)
blstm_bwd_kernel(const T* __restrict__ dout, int64_t out_ld_t, int64_t out_ld_b, int pair,
                 const float* __restrict__ acts, const float* __restrict__ cs,
                 const float* __restrict__ w_hh_f, const float* __restrict__ w_hh_r,
                 const int32_t* __restrict__ lens, T* __restrict__ dgates, int Tn, int B, int H, int C) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int grp = blockIdx.y, dir = blockIdx.z;
  const int tid = threadIdx.x;
  const int UPC = H / C;
  const int G4 = 4 * H, GP = G4 + 4;
  extern __shared__ __align__(16) float smem[];
  float* WTs = smem;                         // [UPC][GP]  W_hh[k][u] for this CTA's units u
  float* dgbuf = WTs + (size_t)UPC * GP;     // [2][NB][GP] gathered gate gradients
  float* dhrec = dgbuf + (size_t)2 * NB * GP;  // [UPC][NB]
  float* dcrec = dhrec + UPC * NB;           // [UPC][NB]

  const float* w = dir ? w_hh_r : w_hh_f;
  for (int idx = tid; idx < UPC * G4; idx += BLSTM_THREADS) {
    const int k = idx / UPC, ul = idx % UPC;
    WTs[(size_t)ul * GP + k] = w[(size_t)k * H + rank * UPC + ul];
  }
  for (int idx = tid; idx < 2 * NB * GP; idx += BLSTM_THREADS) dgbuf[idx] = 0.f;
  for (int idx = tid; idx < 2 * UPC * NB; idx += BLSTM_THREADS) dhrec[idx] = 0.f;   // dhrec + dcrec
  cluster.sync();

  const int b0 = grp * NB;
  int nxt = 0;
  for (int s = 0; s < Tn; ++s) {
    const int t = dir ? s : (Tn - 1 - s);     // reverse of the forward processing order
    for (int item = tid; item < UPC * NB; item += BLSTM_THREADS) {
      const int ul = item / NB, bl = item % NB;
      const int b = b0 + bl;
      const int u = rank * UPC + ul;
      float dgi = 0.f, dgf = 0.f, dgg = 0.f, dgo = 0.f, dc_out = 0.f;
      if (b < B && t < lens[b]) {
        const int64_t row = ((int64_t)dir * Tn + t) * B + b;
        const float* a = acts + row * G4;
        const float i_ = a[u], f_ = a[H + u], g_ = a[2 * H + u], o_ = a[3 * H + u];
        const float c_t = cs[row * H + u];
        const int tp = dir ? t + 1 : t - 1;
        const float c_prev = (tp >= 0 && tp < Tn) ? cs[(((int64_t)dir * Tn + tp) * B + b) * H + u] : 0.f;
        const float dh = to_f(dout[(int64_t)(t / pair) * out_ld_t + (int64_t)b * out_ld_b +
                                   (int64_t)(t % pair) * 2 * H + dir * H + u]) + dhrec[ul * NB + bl];
        const float tc = tanhf(c_t);
        const float dc = dcrec[ul * NB + bl] + dh * o_ * (1.f - tc * tc);
        dgo = dh * tc * o_ * (1.f - o_);
        dgi = dc * g_ * i_ * (1.f - i_);
        dgf = dc * c_prev * f_ * (1.f - f_);
        dgg = dc * i_ * (1.f - g_ * g_);
        dc_out = dc * f_;
      }
      dcrec[ul * NB + bl] = dc_out;
      if (b < B) {
        T* dg = dgates + (((int64_t)dir * Tn + t) * B + b) * G4;
        dg[u] = from_f<T>(dgi); dg[H + u] = from_f<T>(dgf);
        dg[2 * H + u] = from_f<T>(dgg); dg[3 * H + u] = from_f<T>(dgo);
      }
      for (int r = 0; r < C; ++r) {
        float* dst = cluster.map_shared_rank(dgbuf, r) + (size_t)(nxt * NB + bl) * GP;
        dst[u] = dgi; dst[H + u] = dgf; dst[2 * H + u] = dgg; dst[3 * H + u] = dgo;
      }
    }
    cluster.sync();
    // ---- dh_{t-1}[b, u] = sum_k dG[b, k] W_hh[k, u]
    for (int item = tid; item < UPC * NB; item += BLSTM_THREADS) {
      const int ul = item / NB, bl = item % NB;
      const float* dg = dgbuf + (size_t)(nxt * NB + bl) * GP;
      const float* wt = WTs + (size_t)ul * GP;
      float a0 = 0.f, a1 = 0.f;
#pragma unroll 4
      for (int k = 0; k < G4; k += 4) {
        const float4 x = *reinterpret_cast<const float4*>(dg + k);
        const float4 y = *reinterpret_cast<const float4*>(wt + k);
        a0 = fmaf(x.x, y.x, a0); a1 = fmaf(x.y, y.y, a1);
        a0 = fmaf(x.z, y.z, a0); a1 = fmaf(x.w, y.w, a1);
      }
      dhrec[ul * NB + bl] = a0 + a1;
    }
    __syncthreads();
    nxt ^= 1;
  }
  cluster.sync();   // nobody exits while a peer may still address its shared memory
}

static int pick_cluster(int64_t H) {
  for (int c = 8; c >= 1; c >>= 1)
    if (H % c == 0) return c;
  return 1;
}

template <typename K, typename... Args>
static int launch_cluster(K kernel, dim3 grid, int C, size_t smem, cudaStream_t st, const char* name,
                          Args... args) {
  if (smem > 227 * 1024)
    return set_error("%s: needs %zu B of shared memory per CTA (> 227 KB); hidden size too large for "
                     "resident recurrent weights", name, smem);
  cudaError_t e = cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return set_error("%s: %s", name, cudaGetErrorString(e));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(BLSTM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kernel, args...);
  if (e != cudaSuccess) return set_error("%s: %s", name, cudaGetErrorString(e));
  count_launch();
  return 0;
}

bool blstm_tc_eligible(int dtype, int64_t H, int64_t out_ld_t, int64_t out_ld_b);
int blstm_fwd_tc(const void* xproj, const float* w_hh_f, const float* w_hh_r, const int32_t* lens, void* out,
                 int64_t out_ld_t, int64_t out_ld_b, int pair, void* hs, float* acts, float* cs, int64_t T_, int64_t B,
                 cudaStream_t st);
int blstm_bwd_tc(const void* dout, int64_t out_ld_t, int64_t out_ld_b, int pair, const float* acts, const float* cs,
                 const float* w_hh_f, const float* w_hh_r, const int32_t* lens, void* dgates, int64_t T_, int64_t B,
                 cudaStream_t st);
int set_timeline(void* buf);
// lstm_rg.cu: register-resident weights + warp-level MMA, two independent sub-chains per CTA
bool blstm_rg_eligible(int dtype, int64_t H);
int blstm_fwd_rg(const void* xproj, const float* w_hh_f, const float* w_hh_r, const int32_t* lens, void* out,
                 int64_t out_ld_t, int64_t out_ld_b, int pair, void* hs, float* acts, float* cs, int64_t T_, int64_t B,
                 cudaStream_t st);
int blstm_bwd_rg(const void* dout, int64_t out_ld_t, int64_t out_ld_b, int pair, const float* acts, const float* cs,
                 const float* w_hh_f, const float* w_hh_r, const int32_t* lens, void* dgates, int64_t T_, int64_t B,
                 cudaStream_t st);
int set_timeline_rg(void* buf);
static int g_blstm_backend = 0;   // 0 auto (tcgen05 kernels where eligible), 1 CUDA cores only, 2 = 0, 3 register-resident warp-MMA kernels

}  // namespace b200st

using namespace b200st;

extern "C" {

int b200st_debug_timeline(void* buf) { const int a = set_timeline(buf), b = set_timeline_rg(buf); return a ? a : b; }

int b200st_set_blstm_backend(int mode) {
  const int old = g_blstm_backend;
  if (mode >= 0 && mode <= 3) g_blstm_backend = mode;
  return old;
}

int b200st_blstm_saved_layout(int dtype, int64_t H) {
  return (g_blstm_backend == 3 && blstm_rg_eligible(dtype, H)) ? 1 : 0;
}

int b200st_lstm_cell_fwd(int dtype, const void* gates, const void* gates_b, const void* gates_c,
                         const float* c_prev, void* h, float* c, float* acts, const void* residual,
                         void* out_res, int64_t B, int64_t H, b200st_stream_t stream) {
  if (B * H <= 0) return 0;
  B200ST_DISPATCH(dtype, T, {
    B200ST_CUDA(launch_pdl(lstm_cell_fwd_kernel<T>, dim3((unsigned)ceil_div(B * H, 256)), dim3(256), 0, (cudaStream_t)stream,
                           (const T*)gates, (const T*)gates_b, (const T*)gates_c, c_prev, (T*)h, c, acts,
                           (const T*)residual, (T*)out_res, B, (int)H));
  });
  B200ST_LAUNCH_CHECK("lstm_cell_fwd");
  return 0;
}

int b200st_lstm_cell_bwd(int dtype, const void* dh_a, const void* dh_b, const void* dh_c,
                         const float* dc_next, const float* acts, const float* c_prev, const float* c,
                         void* dgates, float* dc_prev, int64_t B, int64_t H, b200st_stream_t stream) {
  if (B * H <= 0) return 0;
  B200ST_DISPATCH(dtype, T, {
    B200ST_CUDA(launch_pdl(lstm_cell_bwd_kernel<T>, dim3((unsigned)ceil_div(B * H, 256)), dim3(256), 0, (cudaStream_t)stream,
                           (const T*)dh_a, (const T*)dh_b, (const T*)dh_c, dc_next, acts, c_prev, c, (T*)dgates,
                           dc_prev, B, (int)H));
  });
  B200ST_LAUNCH_CHECK("lstm_cell_bwd");
  return 0;
}

int b200st_blstm_fwd(int dtype, const void* xproj, const float* w_hh_f, const float* w_hh_r,
                     const int32_t* lens, void* out, int64_t out_ld_t, int64_t out_ld_b, int pair,
                     void* hs, float* acts, float* cs, int64_t T_, int64_t B, int64_t H,
                     b200st_stream_t stream) {
  if (T_ <= 0 || B <= 0) return 0;
  if (H % 4 != 0) return set_error("blstm_fwd: hidden size %lld must be a multiple of 4", (long long)H);
  if (pair != 1 && pair != 2) return set_error("blstm_fwd: pair must be 1 or 2");
  const bool use_rg = g_blstm_backend == 3 && blstm_rg_eligible(dtype, H);
  const bool use_tc = !use_rg && g_blstm_backend != 1 && blstm_tc_eligible(dtype, H, out_ld_t, out_ld_b);
  if (use_rg || use_tc) {
    if (hs) {
      const size_t plane = (size_t)B * H * 2;
      B200ST_CUDA(cudaMemsetAsync(hs, 0, plane, (cudaStream_t)stream));
      B200ST_CUDA(cudaMemsetAsync((char*)hs + ((size_t)(T_ + 1) + T_) * plane, 0, plane, (cudaStream_t)stream));
    }
    if (use_rg)
      return blstm_fwd_rg(xproj, w_hh_f, w_hh_r, lens, out, out_ld_t, out_ld_b, pair, hs, acts, cs, T_, B,
                          (cudaStream_t)stream);
    return blstm_fwd_tc(xproj, w_hh_f, w_hh_r, lens, out, out_ld_t, out_ld_b, pair, hs, acts, cs, T_, B,
                        (cudaStream_t)stream);
  }
  constexpr int NB = 8;
  const int C = pick_cluster(H);
  const int UPC = (int)H / C, R = 4 * UPC;
  const size_t smem = ((size_t)R * (H + 4) + 2 * NB * H + (size_t)R * NB + (size_t)UPC * NB) * sizeof(float);
  dim3 grid(C, (unsigned)ceil_div(B, NB), 2);
  cudaStream_t st = (cudaStream_t)stream;
  if (hs) {   // the two zero planes that stand for the initial state
    const size_t esz = dtype == B200ST_F32 ? 4 : 2;
    const size_t plane = (size_t)B * H * esz;
    B200ST_CUDA(cudaMemsetAsync(hs, 0, plane, st));
    B200ST_CUDA(cudaMemsetAsync((char*)hs + ((size_t)(T_ + 1) + T_) * plane, 0, plane, st));
  }
  B200ST_DISPATCH(dtype, T, {
    return launch_cluster(blstm_fwd_kernel<T, NB>, grid, C, smem, st, "blstm_fwd", (const T*)xproj,
                          w_hh_f, w_hh_r, lens, (T*)out, out_ld_t, out_ld_b, pair, (T*)hs, acts, cs,
                          (int)T_, (int)B, (int)H, C);
  });
  return 0;
}

int b200st_blstm_bwd(int dtype, const void* dout, int64_t out_ld_t, int64_t out_ld_b, int pair,
                     const float* acts, const float* cs, const float* w_hh_f, const float* w_hh_r,
                     const int32_t* lens, void* dgates, int64_t T_, int64_t B, int64_t H,
                     b200st_stream_t stream) {
  if (T_ <= 0 || B <= 0) return 0;
  if (H % 4 != 0) return set_error("blstm_bwd: hidden size %lld must be a multiple of 4", (long long)H);
  if (g_blstm_backend == 3 && blstm_rg_eligible(dtype, H))
    return blstm_bwd_rg(dout, out_ld_t, out_ld_b, pair, acts, cs, w_hh_f, w_hh_r, lens, dgates, T_, B,
                        (cudaStream_t)stream);
  if (g_blstm_backend != 1 && blstm_tc_eligible(dtype, H, out_ld_t, out_ld_b))
    return blstm_bwd_tc(dout, out_ld_t, out_ld_b, pair, acts, cs, w_hh_f, w_hh_r, lens, dgates, T_, B,
                        (cudaStream_t)stream);
  constexpr int NB = 8;
  const int C = pick_cluster(H);
  const int UPC = (int)H / C;
  const size_t GP = 4 * H + 4;
  const size_t smem = ((size_t)UPC * GP + 2 * NB * GP + 2 * (size_t)UPC * NB) * sizeof(float);
  dim3 grid(C, (unsigned)ceil_div(B, NB), 2);
  B200ST_DISPATCH(dtype, T, {
    return launch_cluster(blstm_bwd_kernel<T, NB>, grid, C, smem, (cudaStream_t)stream, "blstm_bwd",
                          (const T*)dout, out_ld_t, out_ld_b, pair, acts, cs, w_hh_f, w_hh_r, lens,
                          (T*)dgates, (int)T_, (int)B, (int)H, C);
  });
  return 0;
}

}  // extern "C"
