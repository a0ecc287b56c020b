// Attention kernels (exact-fp32 CUDA-core versions).
//   * multi-head scaled-dot-product attention core, forward and backward (layers.py:213-229)
//   * LAS single-head bilinear attention step, forward and backward (attention.py:190-193,250-273)
//   * row arg-max and the LAS decode-length rule (Dec.py:320-341)
// Sequence lengths on the training path are tiny (<= 50 keys, 126 acoustic frames), so one warp owns a
// query row end to end: scores, mask, softmax and the P.V product never leave the SM.
#include "common.cuh"
#include "philox.cuh"

namespace b200st {

constexpr int MHA_WARPS = 4;

// grid: (ceil(Lq / MHA_WARPS), H, B); dynamic smem: MHA_WARPS * (d + Lk) floats.
template <typename T>
__global__ void mha_fwd_kernel(const T* __restrict__ q, int64_t ldq, const T* __restrict__ k,
                               int64_t ldk, const T* __restrict__ v, int64_t ldv,
                               const uint8_t* __restrict__ mask, int64_t mask_sb, int64_t mask_sq,
                               T* __restrict__ o, int64_t ldo, T* __restrict__ p, int H, int Lq, int Lk,
                               int d, float temperature) {
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i = blockIdx.x * MHA_WARPS + w;
  const int h = blockIdx.y, b = blockIdx.z;
  if (i >= Lq) return;
  float* qs = sm + w * (d + Lk);
  float* sc = qs + d;
  const T* qr = q + ((int64_t)b * Lq + i) * ldq + (int64_t)h * d;
  for (int c = lane; c < d; c += 32) qs[c] = to_f(qr[c]) / temperature;   // layers.py:216 (q / temperature)
  __syncwarp();
  const T* kb = k + (int64_t)b * Lk * ldk + (int64_t)h * d;
  const uint8_t* mr = mask ? mask + b * mask_sb + i * mask_sq : nullptr;
  float mx = -INFINITY;
  for (int j = lane; j < Lk; j += 32) {
    const T* kr = kb + (int64_t)j * ldk;
    float s = 0.f;
    for (int c = 0; c < d; ++c) s = fmaf(qs[c], to_f(kr[c]), s);
    if (mr && mr[j] == 0) s = -1e9f;                                     // layers.py:224
    sc[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < Lk; j += 32) {
    const float e = expf(sc[j] - mx);
    sc[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  T* pr = p ? p + (((int64_t)b * H + h) * Lq + i) * Lk : nullptr;
  for (int j = lane; j < Lk; j += 32) {
    const float pv = sc[j] * inv;
    sc[j] = pv;
    if (pr) pr[j] = from_f<T>(pv);
  }
  __syncwarp();
  const T* vb = v + (int64_t)b * Lk * ldv + (int64_t)h * d;
  T* orow = o + ((int64_t)b * Lq + i) * ldo + (int64_t)h * d;
  for (int c = lane; c < d; c += 32) {
    float acc = 0.f;
    for (int j = 0; j < Lk; ++j) acc = fmaf(sc[j], to_f(vb[(int64_t)j * ldv + c]), acc);
    orow[c] = from_f<T>(acc);
  }
}

// Pass 1 (per query row): dP = dO V^T, dS = P * (dP - sum(P dP)), dQ = dS K / temperature.
template <typename T>
__global__ void mha_bwd_q_kernel(const T* __restrict__ dout, int64_t ldo, const T* __restrict__ k,
                                 int64_t ldk, const T* __restrict__ v, int64_t ldv,
                                 const T* __restrict__ p, T* __restrict__ ds, T* __restrict__ dq,
                                 int64_t lddq, int H, int Lq, int Lk, int d, float temperature) {
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i = blockIdx.x * MHA_WARPS + w;
  const int h = blockIdx.y, b = blockIdx.z;
  if (i >= Lq) return;
  float* dos = sm + w * (d + Lk);
  float* sc = dos + d;
  const T* dor = dout + ((int64_t)b * Lq + i) * ldo + (int64_t)h * d;
  for (int c = lane; c < d; c += 32) dos[c] = to_f(dor[c]);
  __syncwarp();
  const T* vb = v + (int64_t)b * Lk * ldv + (int64_t)h * d;
  const T* pr = p + (((int64_t)b * H + h) * Lq + i) * Lk;
  float delta = 0.f;
  for (int j = lane; j < Lk; j += 32) {
    const T* vr = vb + (int64_t)j * ldv;
    float dp = 0.f;
    for (int c = 0; c < d; ++c) dp = fmaf(dos[c], to_f(vr[c]), dp);
    sc[j] = dp;
    delta += dp * to_f(pr[j]);
  }
  delta = warp_sum(delta);
  T* dsr = ds + (((int64_t)b * H + h) * Lq + i) * Lk;
  for (int j = lane; j < Lk; j += 32) {
    const float g = to_f(pr[j]) * (sc[j] - delta);
    sc[j] = g;
    dsr[j] = from_f<T>(g);
  }
  __syncwarp();
  const T* kb = k + (int64_t)b * Lk * ldk + (int64_t)h * d;
  T* dqr = dq + ((int64_t)b * Lq + i) * lddq + (int64_t)h * d;
  for (int c = lane; c < d; c += 32) {
    float acc = 0.f;
    for (int j = 0; j < Lk; ++j) acc = fmaf(sc[j], to_f(kb[(int64_t)j * ldk + c]), acc);
    dqr[c] = from_f<T>(acc / temperature);
  }
}

// Pass 2 (per key row): dV[j] = sum_i P[i,j] dO[i],  dK[j] = sum_i dS[i,j] Q[i] / temperature.
template <typename T>
__global__ void mha_bwd_kv_kernel(const T* __restrict__ dout, int64_t ldo, const T* __restrict__ q,
                                  int64_t ldq, const T* __restrict__ p, const T* __restrict__ ds,
                                  T* __restrict__ dk, int64_t lddk, T* __restrict__ dv, int64_t lddv,
                                  int H, int Lq, int Lk, int d, float temperature) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int j = blockIdx.x * MHA_WARPS + w;
  const int h = blockIdx.y, b = blockIdx.z;
  if (j >= Lk) return;
  const T* pb = p + (((int64_t)b * H + h) * Lq) * Lk + j;
  const T* dsb = ds + (((int64_t)b * H + h) * Lq) * Lk + j;
  const T* dob = dout + (int64_t)b * Lq * ldo + (int64_t)h * d;
  const T* qb = q + (int64_t)b * Lq * ldq + (int64_t)h * d;
  for (int c = lane; c < d; c += 32) {
    float av = 0.f, ak = 0.f;
    for (int i = 0; i < Lq; ++i) {
      av = fmaf(to_f(pb[(int64_t)i * Lk]), to_f(dob[(int64_t)i * ldo + c]), av);
      ak = fmaf(to_f(dsb[(int64_t)i * Lk]), to_f(qb[(int64_t)i * ldq + c]) / temperature, ak);
    }
    dv[((int64_t)b * Lk + j) * lddv + (int64_t)h * d + c] = from_f<T>(av);
    dk[((int64_t)b * Lk + j) * lddk + (int64_t)h * d + c] = from_f<T>(ak);
  }
}


// ------------------------------------------------------------------------------------------------
// Tiled variants for the training shapes (Lq, Lk <= ~128, d <= 128): one CTA per (batch, head) keeps
// Q, K, V (and dO, P, dS in backward) in shared memory as fp32, so every global element is read once with
// coalesced row loads and the three small products run out of shared memory with broadcast/conflict-free
// access (rows padded to d+1).  At L = 50, d = 64 a head is ~0.4 MFLOP: far below one tensor-core tile's
// fixed cost, so the CUDA cores are the right unit here; the grid (B*H = 512 CTAs) fills the 148 SMs.
// ------------------------------------------------------------------------------------------------
constexpr int MHA_T_THREADS = 256;

template <typename T>
__device__ __forceinline__ void mha_load_rows(float* dst, int pitch, const T* __restrict__ src, int64_t ld,
                                              int rows, int d, float scale) {
  for (int idx = threadIdx.x; idx < rows * d; idx += MHA_T_THREADS) {
    const int r = idx / d, c = idx - r * d;
    dst[r * pitch + c] = to_f(src[(int64_t)r * ld + c]) * scale;
  }
}

// Load `rows` x d row-major global rows into shared memory TRANSPOSED: dst[c * pitch + r].
template <typename T>
__device__ __forceinline__ void mha_load_rows_t(float* dst, int pitch, const T* __restrict__ src, int64_t ld,
                                                int rows, int d) {
  for (int idx = threadIdx.x; idx < rows * d; idx += MHA_T_THREADS) {
    const int r = idx / d, c = idx - r * d;
    dst[c * pitch + r] = to_f(src[(int64_t)r * ld + c]);
  }
}

// Register-blocked inner products out of shared memory.
//   ab_t:  out[i][j] = sum_c A[i][c] * Bt[c][j]   (A row-major pitch pa, Bt transposed pitch pb): 2 x 4 per thread,
//          the four j share one 16-byte load, the A values are warp broadcasts.
__device__ __forceinline__ void mha_ab_t(float* out, int po, const float* A, int pa, const float* Bt, int pb, int Li,
                                         int Lj, int d) {
  const int tj = (Lj + 3) >> 2, ti = (Li + 1) >> 1;
  for (int idx = threadIdx.x; idx < ti * tj; idx += MHA_T_THREADS) {
    const int i0 = (idx / tj) * 2, j0 = (idx % tj) * 4;
    const float* a0 = A + i0 * pa;
    const float* a1 = A + min(i0 + 1, Li - 1) * pa;
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll 4
    for (int c = 0; c < d; ++c) {
      const float4 y = *reinterpret_cast<const float4*>(Bt + c * pb + j0);
      const float x0 = a0[c], x1 = a1[c];
      acc[0][0] = fmaf(x0, y.x, acc[0][0]); acc[0][1] = fmaf(x0, y.y, acc[0][1]);
      acc[0][2] = fmaf(x0, y.z, acc[0][2]); acc[0][3] = fmaf(x0, y.w, acc[0][3]);
      acc[1][0] = fmaf(x1, y.x, acc[1][0]); acc[1][1] = fmaf(x1, y.y, acc[1][1]);
      acc[1][2] = fmaf(x1, y.z, acc[1][2]); acc[1][3] = fmaf(x1, y.w, acc[1][3]);
    }
    *reinterpret_cast<float4*>(out + i0 * po + j0) = make_float4(acc[0][0], acc[0][1], acc[0][2], acc[0][3]);
    if (i0 + 1 < Li)
      *reinterpret_cast<float4*>(out + (i0 + 1) * po + j0) = make_float4(acc[1][0], acc[1][1], acc[1][2], acc[1][3]);
  }
}
//   pv:    out[i][c] = scale * sum_j P[i][j] * V[j][c]  (P pitch pp, V row-major pitch d): 4 rows x 2 columns per thread,
//          columns adjacent across lanes (conflict-free V reads, coalesced global stores), P values broadcast.
template <typename T>
__device__ __forceinline__ void mha_pv(T* __restrict__ out, int64_t ldo, const float* P, int pp, const float* V, int Li,
                                       int Lj, int d, float scale) {
  const int ti = (Li + 3) >> 2, tc = d >> 1;
  for (int idx = threadIdx.x; idx < ti * tc; idx += MHA_T_THREADS) {
    const int i0 = (idx / tc) * 4, c0 = idx % tc;
    const float* p0 = P + i0 * pp;
    const float* p1 = P + min(i0 + 1, Li - 1) * pp;
    const float* p2 = P + min(i0 + 2, Li - 1) * pp;
    const float* p3 = P + min(i0 + 3, Li - 1) * pp;
    float acc[4][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
#pragma unroll 4
    for (int j = 0; j < Lj; ++j) {
      const float v0 = V[j * d + c0], v1 = V[j * d + c0 + tc];
      const float x0 = p0[j], x1 = p1[j], x2 = p2[j], x3 = p3[j];
      acc[0][0] = fmaf(x0, v0, acc[0][0]); acc[0][1] = fmaf(x0, v1, acc[0][1]);
      acc[1][0] = fmaf(x1, v0, acc[1][0]); acc[1][1] = fmaf(x1, v1, acc[1][1]);
      acc[2][0] = fmaf(x2, v0, acc[2][0]); acc[2][1] = fmaf(x2, v1, acc[2][1]);
      acc[3][0] = fmaf(x3, v0, acc[3][0]); acc[3][1] = fmaf(x3, v1, acc[3][1]);
    }
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      if (i0 + ii >= Li) break;
      T* r = out + (int64_t)(i0 + ii) * ldo;
      r[c0] = from_f<T>(acc[ii][0] * scale);
      r[c0 + tc] = from_f<T>(acc[ii][1] * scale);
    }
  }
}
//   ptv:   out[j][c] = sum_i P[i][j] * X[i][c]   (column j of P times rows of X): 2 key rows x 2 columns per thread.
template <typename T>
__device__ __forceinline__ void mha_ptv(T* __restrict__ out, int64_t ldo, const float* P, int pp, const float* X, int Li,
                                        int Lj, int d) {
  const int tj = (Lj + 1) >> 1, tc = d >> 1;
  for (int idx = threadIdx.x; idx < tj * tc; idx += MHA_T_THREADS) {
    const int j0 = (idx / tc) * 2, c0 = idx % tc;
    const int j1 = min(j0 + 1, Lj - 1);
    float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
#pragma unroll 4
    for (int i = 0; i < Li; ++i) {
      const float s0 = P[i * pp + j0], s1 = P[i * pp + j1];
      const float xa = X[i * d + c0], xb = X[i * d + c0 + tc];
      acc[0][0] = fmaf(s0, xa, acc[0][0]); acc[0][1] = fmaf(s0, xb, acc[0][1]);
      acc[1][0] = fmaf(s1, xa, acc[1][0]); acc[1][1] = fmaf(s1, xb, acc[1][1]);
    }
    T* r0 = out + (int64_t)j0 * ldo;
    r0[c0] = from_f<T>(acc[0][0]); r0[c0 + tc] = from_f<T>(acc[0][1]);
    if (j0 + 1 < Lj) {
      T* r1 = out + (int64_t)(j0 + 1) * ldo;
      r1[c0] = from_f<T>(acc[1][0]); r1[c0 + tc] = from_f<T>(acc[1][1]);
    }
  }
}

// shared-memory pitch of an [L x Lk] score tile: multiple of 4 floats (16-byte row alignment), >= Lk + 3 so that the
// 4-wide tiles of the last key group stay inside the row
__host__ __device__ inline int mha_lkp(int Lk) { return ((Lk + 3) & ~3) + 4; }

template <typename T>
__global__ void __launch_bounds__(MHA_T_THREADS)
mha_fwd_tiled_kernel(const T* __restrict__ q, int64_t ldq, const T* __restrict__ k, int64_t ldk,
                     const T* __restrict__ v, int64_t ldv, const uint8_t* __restrict__ mask, int64_t mask_sb,
                     int64_t mask_sq, T* __restrict__ o, int64_t ldo, T* __restrict__ p, int H, int Lq, int Lk,
                     int d, float temperature, float drop_p, const int64_t* __restrict__ rng, int64_t site) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ __align__(16) float sm[];
  const int lkp = mha_lkp(Lk);
  DropRng drop;
  if (drop_p > 0.f) drop.init(rng, site, drop_p);
  float* Qs = sm;                  // [Lq][d]    q / temperature
  float* Kt = Qs + Lq * d;         // [d][lkp]   K transposed
  float* Vs = Kt + d * lkp;        // [Lk][d]
  float* Ss = Vs + Lk * d;         // [Lq][lkp]
  const int h = blockIdx.x, b = blockIdx.y;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  // q / temperature: division (not reciprocal multiply) to follow layers.py:216 bit for bit
  for (int idx = threadIdx.x; idx < Lq * d; idx += MHA_T_THREADS) {
    const int r = idx / d, c = idx - r * d;
    Qs[idx] = to_f(q[((int64_t)b * Lq + r) * ldq + (int64_t)h * d + c]) / temperature;
  }
  for (int idx = threadIdx.x; idx < d * lkp; idx += MHA_T_THREADS) Kt[idx] = 0.f;   // padding columns must be finite
  __syncthreads();
  mha_load_rows_t(Kt, lkp, k + (int64_t)b * Lk * ldk + (int64_t)h * d, ldk, Lk, d);
  mha_load_rows(Vs, d, v + (int64_t)b * Lk * ldv + (int64_t)h * d, ldv, Lk, d, 1.f);
  __syncthreads();
  mha_ab_t(Ss, lkp, Qs, d, Kt, lkp, Lq, Lk, d);
  __syncthreads();
  for (int i = w; i < Lq; i += MHA_T_THREADS / 32) {
    float* sr = Ss + i * lkp;
    const uint8_t* mr = mask ? mask + b * mask_sb + i * mask_sq : nullptr;
    float mx = -INFINITY;
    for (int j = lane; j < Lk; j += 32) {
      float sv = sr[j];
      if (mr && mr[j] == 0) sv = -1e9f;                                  // layers.py:224
      sr[j] = sv;
      mx = fmaxf(mx, sv);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < Lk; j += 32) { const float e = expf(sr[j] - mx); sr[j] = e; sum += e; }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    T* pr = p ? p + (((int64_t)b * H + h) * Lq + i) * Lk : nullptr;
    const uint64_t row0 = (((uint64_t)b * H + h) * Lq + i) * (uint64_t)Lk;
    for (int j = lane; j < Lk; j += 32) {
      float pv = sr[j] * inv;
      if (pr) pr[j] = from_f<T>(pv);                   // the UN-dropped probabilities are what backward needs
      // attn = dropout(softmax(.)) (layers.py:226): the dropped probabilities multiply V
      if (drop_p > 0.f) pv *= drop.factor(row0 + j);
      sr[j] = pv;
    }
  }
  __syncthreads();
  mha_pv(o + (int64_t)b * Lq * ldo + (int64_t)h * d, ldo, Ss, lkp, Vs, Lq, Lk, d, 1.f);
}

// dP = dO V^T;  dS = P * (dP - rowsum(P dP));  dQ = dS K / temp;  dK = dS^T (Q / temp);  dV = P^T dO.
template <typename T>
__global__ void __launch_bounds__(MHA_T_THREADS)
mha_bwd_tiled_kernel(const T* __restrict__ dout, int64_t ldo, const T* __restrict__ q, int64_t ldq,
                     const T* __restrict__ k, int64_t ldk, const T* __restrict__ v, int64_t ldv,
                     const T* __restrict__ p, T* __restrict__ dq, int64_t lddq, T* __restrict__ dk, int64_t lddk,
                     T* __restrict__ dv, int64_t lddv, int H, int Lq, int Lk, int d, float temperature,
                     float drop_p, const int64_t* __restrict__ rng, int64_t site) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ __align__(16) float sm[];
  const int lkp = mha_lkp(Lk);
  DropRng drop;
  if (drop_p > 0.f) drop.init(rng, site, drop_p);
  float* Qs = sm;                  // [Lq][d]    q / temperature
  float* Ks = Qs + Lq * d;         // [Lk][d]
  float* Vt = Ks + Lk * d;         // [d][lkp]   V transposed
  float* dOs = Vt + d * lkp;       // [Lq][d]
  float* Ps = dOs + Lq * d;        // [Lq][lkp]
  float* dSs = Ps + Lq * lkp;      // [Lq][lkp]
  const int h = blockIdx.x, b = blockIdx.y;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int idx = threadIdx.x; idx < Lq * d; idx += MHA_T_THREADS) {
    const int r = idx / d, c = idx - r * d;
    Qs[idx] = to_f(q[((int64_t)b * Lq + r) * ldq + (int64_t)h * d + c]) / temperature;
  }
  for (int idx = threadIdx.x; idx < d * lkp; idx += MHA_T_THREADS) Vt[idx] = 0.f;
  __syncthreads();
  mha_load_rows(Ks, d, k + (int64_t)b * Lk * ldk + (int64_t)h * d, ldk, Lk, d, 1.f);
  mha_load_rows_t(Vt, lkp, v + (int64_t)b * Lk * ldv + (int64_t)h * d, ldv, Lk, d);
  mha_load_rows(dOs, d, dout + (int64_t)b * Lq * ldo + (int64_t)h * d, ldo, Lq, d, 1.f);
  const T* pb = p + ((int64_t)b * H + h) * Lq * Lk;
  for (int idx = threadIdx.x; idx < Lq * Lk; idx += MHA_T_THREADS) {
    const int i = idx / Lk, j = idx - i * Lk;
    Ps[i * lkp + j] = to_f(pb[idx]);
  }
  __syncthreads();
  mha_ab_t(dSs, lkp, dOs, d, Vt, lkp, Lq, Lk, d);          // dP
  __syncthreads();
  for (int i = w; i < Lq; i += MHA_T_THREADS / 32) {
    float delta = 0.f;
    if (drop_p > 0.f) {
      // forward used P~ = m * P (m = keep / (1 - p)): dP = m * dP~;  dS = P * (dP - sum_j P dP);  dV needs P~, so the
      // probability tile is rewritten to P~ once dS has been formed
      const uint64_t row0 = (((uint64_t)b * H + h) * Lq + i) * (uint64_t)Lk;
      for (int j = lane; j < Lk; j += 32) {
        const float dp = dSs[i * lkp + j] * drop.factor(row0 + j);
        delta += dp * Ps[i * lkp + j];
        dSs[i * lkp + j] = dp;
      }
      delta = warp_sum(delta);
      for (int j = lane; j < Lk; j += 32) {
        const float pr = Ps[i * lkp + j];
        dSs[i * lkp + j] = pr * (dSs[i * lkp + j] - delta);
        Ps[i * lkp + j] = pr * drop.factor(row0 + j);
      }
      continue;
    }
    for (int j = lane; j < Lk; j += 32) delta += dSs[i * lkp + j] * Ps[i * lkp + j];
    delta = warp_sum(delta);
    for (int j = lane; j < Lk; j += 32) dSs[i * lkp + j] = Ps[i * lkp + j] * (dSs[i * lkp + j] - delta);
  }
  __syncthreads();
  mha_pv(dq + (int64_t)b * Lq * lddq + (int64_t)h * d, lddq, dSs, lkp, Ks, Lq, Lk, d, 1.f / temperature);
  mha_ptv(dk + (int64_t)b * Lk * lddk + (int64_t)h * d, lddk, dSs, lkp, Qs, Lq, Lk, d);
  mha_ptv(dv + (int64_t)b * Lk * lddv + (int64_t)h * d, lddv, Ps, lkp, dOs, Lq, Lk, d);
}

// tensor-core variants (attention_tc.cu): return 1 when the shape is not covered
int mha_fwd_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const uint8_t* mask,
               int64_t mask_sb, int64_t mask_sq, void* o, int64_t ldo, void* p, int64_t B, int64_t H, int64_t Lq,
               int64_t Lk, int64_t d, float temperature, cudaStream_t st);
int mha_bwd_tc(const void* dout, int64_t ldo, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
               int64_t ldv, const void* p, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
               int64_t B, int64_t H, int64_t Lq, int64_t Lk, int64_t d, float temperature, cudaStream_t st);

// ------------------------------------------------------------------------------------------------
// Single-query attention over a key/value CACHE (incremental decoding): the decoder step of
// forward_translate / forward_eval (Seq2seq.py:260-393) only needs the newest position's output, whose keys and values
// for positions < i were produced by earlier steps.  One CTA per (head, hypothesis); the hypothesis' history may live
// in OTHER cache slots after a beam re-ordering, so key t of hypothesis b is read from slot anc[t * n_hyp + b]
// (ancestry table, maintained by the host loop like the reference re-orders preds_exp, Seq2seq.py:381-384) -- the
// cache itself is never moved.  Cross-attention passes anc = NULL and slot = b / bdiv (beams of one utterance share
// the encoder keys).  Same score arithmetic as the full kernel: (q / temperature) . k, masked scores SET to -1e9.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128)
mha_decode_kernel(const T* __restrict__ q, int64_t ldq, const T* __restrict__ kc, const T* __restrict__ vc,
                  int64_t stride_b, int64_t stride_t, const int32_t* __restrict__ anc, int n_hyp, int bdiv,
                  const uint8_t* __restrict__ mask, int64_t mask_sb, int mask_bdiv, T* __restrict__ o, int64_t ldo,
                  int Lk, int d, float temperature) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ __align__(16) float sm[];
  float* qs = sm;                      // [d]
  float* sc = qs + d;                  // [Lk] scores -> probabilities
  int* slot = reinterpret_cast<int*>(sc + Lk);   // [Lk] cache slot of every key
  __shared__ float red[32];
  const int h = blockIdx.x, b = blockIdx.y;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int c = threadIdx.x; c < d; c += blockDim.x) qs[c] = to_f(q[(int64_t)b * ldq + (int64_t)h * d + c]) / temperature;
  for (int t = threadIdx.x; t < Lk; t += blockDim.x) slot[t] = anc ? anc[(int64_t)t * n_hyp + b] : b / bdiv;
  __syncthreads();
  const uint8_t* mr = mask ? mask + (int64_t)(b / mask_bdiv) * mask_sb : nullptr;
  float mx = -INFINITY;
  for (int t = w; t < Lk; t += nw) {
    const T* kr = kc + (int64_t)slot[t] * stride_b + (int64_t)t * stride_t + (int64_t)h * d;
    float acc = 0.f;
    for (int c = lane; c < d; c += 32) acc = fmaf(qs[c], to_f(kr[c]), acc);
    acc = warp_sum(acc);
    if (mr && mr[t] == 0) acc = -1e9f;
    if (lane == 0) sc[t] = acc;
    mx = fmaxf(mx, acc);
  }
  mx = block_max(mx, red);
  __syncthreads();
  float sum = 0.f;
  for (int t = threadIdx.x; t < Lk; t += blockDim.x) { const float e = expf(sc[t] - mx); sc[t] = e; sum += e; }
  sum = block_sum(sum, red);
  __syncthreads();
  const float inv = 1.f / sum;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float acc = 0.f;
    for (int t = 0; t < Lk; ++t)
      acc = fmaf(sc[t] * inv, to_f(vc[(int64_t)slot[t] * stride_b + (int64_t)t * stride_t + (int64_t)h * d + c]), acc);
    o[(int64_t)b * ldo + (int64_t)h * d + c] = from_f<T>(acc);
  }
}

static int g_las_backend = 0;     // 0 = auto (key-split cluster kernels for bf16), 1 = one CTA per sequence
static int g_mha_backend = 0;     // 0 = auto (tensor cores for bf16 when the shape fits), 1 = SIMT tiles only

static size_t mha_fwd_tiled_smem(int64_t Lq, int64_t Lk, int64_t d) {
  const int64_t lkp = mha_lkp((int)Lk);
  return (Lq * d + d * lkp + Lk * d + Lq * lkp) * sizeof(float);
}
static size_t mha_bwd_tiled_smem(int64_t Lq, int64_t Lk, int64_t d) {
  const int64_t lkp = mha_lkp((int)Lk);
  return (2 * Lq * d + Lk * d + d * lkp + 2 * Lq * lkp) * sizeof(float);
}

// ------------------------------------------------------------------------------------------------
// LAS attention step: one CTA per batch row.  dynamic smem: (D + 2 * Tk) floats.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void las_attn_fwd_kernel(const T* __restrict__ q, const T* __restrict__ wk,
                                    const T* __restrict__ vals, const int32_t* __restrict__ klens,
                                    T* __restrict__ ctx, float* __restrict__ probs, int Tk, int D,
                                    int Dv) {
  extern __shared__ float sm[];
  __shared__ float scratch[32];
  float* qs = sm;
  float* sc = sm + D;
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int c = threadIdx.x; c < D; c += blockDim.x) qs[c] = to_f(q[(int64_t)b * D + c]);
  __syncthreads();
  const int klen = klens ? klens[b] : Tk;
  const T* wkb = wk + (int64_t)b * Tk * D;
  for (int j = w; j < Tk; j += nw) {
    const T* r = wkb + (int64_t)j * D;
    float s = 0.f;
    for (int c = lane; c < D; c += 32) s = fmaf(qs[c], to_f(r[c]), s);
    s = warp_sum(s);
    if (lane == 0) sc[j] = (j >= klen) ? -1e12f : s;                     // attention.py:250-252
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < Tk; j += blockDim.x) mx = fmaxf(mx, sc[j]);
  mx = block_max(mx, scratch);
  float sum = 0.f;
  for (int j = threadIdx.x; j < Tk; j += blockDim.x) {
    const float e = expf(sc[j] - mx);
    sc[j] = e;
    sum += e;
  }
  sum = block_sum(sum, scratch);
  const float inv = 1.f / sum;
  for (int j = threadIdx.x; j < Tk; j += blockDim.x) {
    const float pv = sc[j] * inv;
    sc[j] = pv;
    probs[(int64_t)b * Tk + j] = pv;
  }
  __syncthreads();
  const T* vb = vals + (int64_t)b * Tk * Dv;
  for (int c = threadIdx.x; c < Dv; c += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < Tk; ++j) acc = fmaf(sc[j], to_f(vb[(int64_t)j * Dv + c]), acc);
    ctx[(int64_t)b * Dv + c] = from_f<T>(acc);
  }
}

template <typename T>
__global__ void las_attn_bwd_kernel(const T* __restrict__ dctx, const T* __restrict__ wk,
                                    const T* __restrict__ vals, const float* __restrict__ probs,
                                    float* __restrict__ dscore, T* __restrict__ dq, int Tk, int D,
                                    int Dv) {
  extern __shared__ float sm[];
  __shared__ float scratch[32];
  float* dcs = sm;          // [Dv]
  float* sc = sm + Dv;      // [Tk]
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int c = threadIdx.x; c < Dv; c += blockDim.x) dcs[c] = to_f(dctx[(int64_t)b * Dv + c]);
  __syncthreads();
  const T* vb = vals + (int64_t)b * Tk * Dv;
  const float* pb = probs + (int64_t)b * Tk;
  for (int j = w; j < Tk; j += nw) {
    const T* r = vb + (int64_t)j * Dv;
    float s = 0.f;
    for (int c = lane; c < Dv; c += 32) s = fmaf(dcs[c], to_f(r[c]), s);
    s = warp_sum(s);
    if (lane == 0) sc[j] = s;
  }
  __syncthreads();
  float delta = 0.f;
  for (int j = threadIdx.x; j < Tk; j += blockDim.x) delta += sc[j] * pb[j];
  delta = block_sum(delta, scratch);
  for (int j = threadIdx.x; j < Tk; j += blockDim.x) {
    const float g = pb[j] * (sc[j] - delta);
    sc[j] = g;
    dscore[(int64_t)b * Tk + j] = g;
  }
  __syncthreads();
  const T* wkb = wk + (int64_t)b * Tk * D;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < Tk; ++j) acc = fmaf(sc[j], to_f(wkb[(int64_t)j * D + c]), acc);
    dq[(int64_t)b * D + c] = from_f<T>(acc);
  }
}


// ---- 16-byte vectorised row access (8 bf16 or 2 x 4 fp32 per call)
__device__ __forceinline__ void load8(const float* p, float* v) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float* v) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
__device__ __forceinline__ void store8(float* p, const float* v) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float* v) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

// Vectorised LAS attention step (D, Dv multiples of 8): one CTA (256 threads) per sequence.
//   scores: one warp per key, lanes stride over 16-byte chunks of the key row;
//   context: thread (group g, chunk c) accumulates keys j = g mod 4 for 8 output columns, groups reduced in smem.
// dynamic smem: (D + Tk + 4 * Dv) floats.
template <typename T>
__global__ void __launch_bounds__(256)
las_attn_fwd_vec_kernel(const T* __restrict__ q, const T* __restrict__ wk, const T* __restrict__ vals,
                        const int32_t* __restrict__ klens, T* __restrict__ ctx, float* __restrict__ probs, int Tk,
                        int D, int Dv) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ __align__(16) float sm[];
  __shared__ float scratch[32];
  float* qs = sm;             // [D]
  float* sc = qs + D;         // [Tk]
  float* part = sc + ((Tk + 3) & ~3);   // [4][Dv]
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int c = threadIdx.x; c < D; c += 256) qs[c] = to_f(q[(int64_t)b * D + c]);
  __syncthreads();
  const int klen = klens ? klens[b] : Tk;
  const T* wkb = wk + (int64_t)b * Tk * D;
  for (int j = w; j < Tk; j += 8) {
    float s = 0.f;
    if (j < klen) {
      const T* r = wkb + (int64_t)j * D;
      for (int c = lane * 8; c < D; c += 256) {
        float v[8];
        load8(r + c, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) s = fmaf(qs[c + i], v[i], s);
      }
      s = warp_sum(s);
    }
    if (lane == 0) sc[j] = (j >= klen) ? -1e12f : s;                     // attention.py:250-252
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < Tk; j += 256) mx = fmaxf(mx, sc[j]);
  mx = block_max(mx, scratch);
  float sum = 0.f;
  for (int j = threadIdx.x; j < Tk; j += 256) { const float e = expf(sc[j] - mx); sc[j] = e; sum += e; }
  sum = block_sum(sum, scratch);
  const float inv = 1.f / sum;
  for (int j = threadIdx.x; j < Tk; j += 256) { const float pv = sc[j] * inv; sc[j] = pv; probs[(int64_t)b * Tk + j] = pv; }
  __syncthreads();
  const T* vb = vals + (int64_t)b * Tk * Dv;
  const int nchunk = Dv >> 3;
  for (int item = threadIdx.x; item < 4 * nchunk; item += 256) {
    const int g = item / nchunk, c = (item - g * nchunk) * 8;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int j = g; j < Tk; j += 4) {
      const float pj = sc[j];
      if (pj == 0.f) continue;            // masked keys (and exact zeros) contribute nothing
      float v[8];
      load8(vb + (int64_t)j * Dv + c, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(pj, v[i], acc[i]);
    }
    store8(part + g * Dv + c, acc);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < Dv; c += 256)
    ctx[(int64_t)b * Dv + c] = from_f<T>((part[c] + part[Dv + c]) + (part[2 * Dv + c] + part[3 * Dv + c]));
}

// dynamic smem: (Dv + Tk + 4 * D) floats
template <typename T>
__global__ void __launch_bounds__(256)
las_attn_bwd_vec_kernel(const T* __restrict__ dctx, const T* __restrict__ wk, const T* __restrict__ vals,
                        const float* __restrict__ probs, float* __restrict__ dscore, T* __restrict__ dq, int Tk, int D,
                        int Dv) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ __align__(16) float sm[];
  __shared__ float scratch[32];
  float* dcs = sm;            // [Dv]
  float* sc = dcs + Dv;       // [Tk]
  float* part = sc + ((Tk + 3) & ~3);   // [4][D]
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int c = threadIdx.x; c < Dv; c += 256) dcs[c] = to_f(dctx[(int64_t)b * Dv + c]);
  __syncthreads();
  const T* vb = vals + (int64_t)b * Tk * Dv;
  const float* pb = probs + (int64_t)b * Tk;
  for (int j = w; j < Tk; j += 8) {
    float s = 0.f;
    if (pb[j] != 0.f) {                    // dscore = p * (dp - delta): rows with p == 0 need no dp
      const T* r = vb + (int64_t)j * Dv;
      for (int c = lane * 8; c < Dv; c += 256) {
        float v[8];
        load8(r + c, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) s = fmaf(dcs[c + i], v[i], s);
      }
      s = warp_sum(s);
    }
    if (lane == 0) sc[j] = s;
  }
  __syncthreads();
  float delta = 0.f;
  for (int j = threadIdx.x; j < Tk; j += 256) delta += sc[j] * pb[j];
  delta = block_sum(delta, scratch);
  for (int j = threadIdx.x; j < Tk; j += 256) {
    const float g = pb[j] * (sc[j] - delta);
    sc[j] = g;
    dscore[(int64_t)b * Tk + j] = g;
  }
  __syncthreads();
  const T* wkb = wk + (int64_t)b * Tk * D;
  const int nchunk = D >> 3;
  for (int item = threadIdx.x; item < 4 * nchunk; item += 256) {
    const int g = item / nchunk, c = (item - g * nchunk) * 8;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int j = g; j < Tk; j += 4) {
      const float gj = sc[j];
      if (gj == 0.f) continue;
      float v[8];
      load8(wkb + (int64_t)j * D + c, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(gj, v[i], acc[i]);
    }
    store8(part + g * D + c, acc);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += 256)
    dq[(int64_t)b * D + c] = from_f<T>((part[c] + part[D + c]) + (part[2 * D + c] + part[3 * D + c]));
}

// ------------------------------------------------------------------------------------------------
// Key-split LAS attention step: a cluster of LAS_CS CTAs per sequence, each owning a contiguous slice of the keys.
// One sequence reads 2 x Tk x 512 x 2 B (~260 KB) of keys + values per step; a single CTA per sequence is bound by
// its own load latency (64 CTAs, ~30 us).  Here 4 x B CTAs keep 4-8 independent 16-byte loads in flight per thread
// and the per-slice softmax statistics / partial context are merged flash-decoding style through distributed
// shared memory: slice r sends (max_r, sum_r) to every CTA and its unnormalised partial context columns to the CTA
// that owns them; after ONE cluster barrier every CTA normalises its probabilities and its quarter of the context.
// ------------------------------------------------------------------------------------------------
constexpr int LAS_CS = 4;
__device__ __forceinline__ uint32_t las_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void las_st_remote(const float* local, uint32_t rank, float v) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(local), ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
}
// split barrier: every CTA of the cluster must be running before its shared memory may be written remotely
__device__ __forceinline__ void las_cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void las_cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void las_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float dot8(const float* a, const uint4& u) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); s = fmaf(a[2 * i], f.x, s); s = fmaf(a[2 * i + 1], f.y, s); }
  return s;
}
__device__ __forceinline__ void axpy8(float* acc, float w, const uint4& u) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); acc[2 * i] = fmaf(w, f.x, acc[2 * i]); acc[2 * i + 1] = fmaf(w, f.y, acc[2 * i + 1]); }
}
// warp-per-key dot products of `vec` (smem, fp32, n floats) with rows [k0, k0+nk) of `rows` (bf16, row stride n),
// four keys (= eight 16-byte loads per lane for n = 512) in flight; keys with use(j) == false get `skip`
template <typename F>
__device__ __forceinline__ void las_key_dots(float* out, const float* vec, const __nv_bfloat16* rows, int n, int nk, float skip, F use) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int jj = w; jj < nk; jj += 32) {
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    bool on[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) on[i] = (jj + 8 * i < nk) && use(jj + 8 * i);
    for (int c = lane * 8; c < n; c += 256) {
      uint4 u[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        u[i] = on[i] ? __ldg(reinterpret_cast<const uint4*>(rows + (int64_t)(jj + 8 * i) * n + c)) : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int i = 0; i < 4; ++i) s[i] += dot8(vec + c, u[i]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float t = warp_sum(s[i]);
      if (lane == 0 && jj + 8 * i < nk) out[jj + 8 * i] = on[i] ? t : skip;
    }
  }
}
// part[g][n] (g < 4) = sum over keys j = g mod 4 of wgt[j] * rows[j][:]; thread (g, 8-column chunk), 8 loads in flight
__device__ __forceinline__ void las_weighted_rows(float* part, const float* wgt, const __nv_bfloat16* rows, int n, int nk) {
  const int nchunk = n >> 3;
  for (int item = threadIdx.x; item < 4 * nchunk; item += 256) {
    const int g = item / nchunk, c = (item - g * nchunk) * 8;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int j0 = g; j0 < nk; j0 += 32) {
      uint4 u[8];
      float wv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int j = j0 + 4 * i;
        wv[i] = j < nk ? wgt[j] : 0.f;
        u[i] = wv[i] != 0.f ? __ldg(reinterpret_cast<const uint4*>(rows + (int64_t)j * n + c)) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) axpy8(acc, wv[i], u[i]);
    }
    store8(part + g * n + c, acc);
  }
}

// dynamic smem floats: D (q) + 64 (scores) + 4*Dv (part) + Dv (red: LAS_CS x Dv/LAS_CS) + 2*LAS_CS (stat)
__global__ void __cluster_dims__(LAS_CS, 1, 1) __launch_bounds__(256)
las_attn_fwd_cl_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ wk,
                       const __nv_bfloat16* __restrict__ vals, const int32_t* __restrict__ klens,
                       __nv_bfloat16* __restrict__ ctx, float* __restrict__ probs, int Tk, int D, int Dv) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ __align__(16) float sm[];
  __shared__ float scratch[32];
  float* qs = sm;                  // [D]
  float* sc = qs + D;              // [64] scores -> exp(score - local max)
  float* part = sc + 64;           // [4][Dv]
  float* red = part + 4 * Dv;      // [LAS_CS][Dv / LAS_CS] partial contexts for the columns this CTA owns
  float* stat = red + Dv;          // [LAS_CS][2] (max, sum) of every slice
  const uint32_t rank = las_rank();
  const int b = blockIdx.y;
  const int kpc = (Tk + LAS_CS - 1) / LAS_CS;
  const int k0 = rank * kpc, nk = max(0, min(Tk, k0 + kpc) - k0);
  las_cluster_arrive();            // "I am running": waited for just before the first remote store
  for (int c = threadIdx.x; c < D; c += 256) qs[c] = to_f(q[(int64_t)b * D + c]);
  __syncthreads();
  const int klen = klens ? klens[b] : Tk;
  las_key_dots(sc, qs, wk + ((int64_t)b * Tk + k0) * D, D, nk, -1e12f, [&](int j) { return k0 + j < klen; });   // attention.py:250-252
  __syncthreads();
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < nk; j += 256) mx = fmaxf(mx, sc[j]);
  mx = block_max(mx, scratch);
  float sum = 0.f;
  for (int j = threadIdx.x; j < nk; j += 256) { const float e = expf(sc[j] - mx); sc[j] = e; sum += e; }
  sum = block_sum(sum, scratch);
  las_cluster_wait();
  if (threadIdx.x < LAS_CS) { las_st_remote(&stat[2 * rank], threadIdx.x, mx); las_st_remote(&stat[2 * rank + 1], threadIdx.x, sum); }
  __syncthreads();
  las_weighted_rows(part, sc, vals + ((int64_t)b * Tk + k0) * Dv, Dv, nk);
  __syncthreads();
  const int cpo = Dv / LAS_CS;     // columns per owner
  for (int c = threadIdx.x; c < Dv; c += 256)
    las_st_remote(&red[rank * cpo + (c % cpo)], (uint32_t)(c / cpo), (part[c] + part[Dv + c]) + (part[2 * Dv + c] + part[3 * Dv + c]));
  las_cluster_sync();
  float M = -INFINITY;
#pragma unroll
  for (int r = 0; r < LAS_CS; ++r) M = fmaxf(M, stat[2 * r]);
  float L = 0.f, f[LAS_CS];
#pragma unroll
  for (int r = 0; r < LAS_CS; ++r) { f[r] = stat[2 * r + 1] > 0.f ? expf(stat[2 * r] - M) : 0.f; L += f[r] * stat[2 * r + 1]; }
  const float inv = 1.f / L;
  for (int j = threadIdx.x; j < nk; j += 256) probs[(int64_t)b * Tk + k0 + j] = sc[j] * f[rank] * inv;
  for (int c = threadIdx.x; c < cpo; c += 256) {
    float a = 0.f;
#pragma unroll
    for (int r = 0; r < LAS_CS; ++r) a = fmaf(f[r], red[r * cpo + c], a);
    ctx[(int64_t)b * Dv + rank * cpo + c] = __float2bfloat16_rn(a * inv);
  }
}

// dynamic smem floats: Dv (dctx) + 64 (dp -> dscore) + 64 (p) + 4*D (part) + D (red) + LAS_CS (stat)
__global__ void __cluster_dims__(LAS_CS, 1, 1) __launch_bounds__(256)
las_attn_bwd_cl_kernel(const __nv_bfloat16* __restrict__ dctx, const __nv_bfloat16* __restrict__ wk,
                       const __nv_bfloat16* __restrict__ vals, const float* __restrict__ probs,
                       float* __restrict__ dscore, __nv_bfloat16* __restrict__ dq, int Tk, int D, int Dv) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ __align__(16) float sm[];
  __shared__ float scratch[32];
  float* dcs = sm;                 // [Dv]
  float* sc = dcs + Dv;            // [64]
  float* ps = sc + 64;             // [64]
  float* part = ps + 64;           // [4][D]
  float* red = part + 4 * D;       // [LAS_CS][D / LAS_CS]
  float* stat = red + D;           // [LAS_CS] partial sum_j p_j dp_j
  const uint32_t rank = las_rank();
  const int b = blockIdx.y;
  const int kpc = (Tk + LAS_CS - 1) / LAS_CS;
  const int k0 = rank * kpc, nk = max(0, min(Tk, k0 + kpc) - k0);
  las_cluster_arrive();
  for (int c = threadIdx.x; c < Dv; c += 256) dcs[c] = to_f(dctx[(int64_t)b * Dv + c]);
  for (int j = threadIdx.x; j < nk; j += 256) ps[j] = probs[(int64_t)b * Tk + k0 + j];
  __syncthreads();
  // dp_j = dctx . V_j (rows with p == 0 need no dp: dscore = p * (dp - delta))
  las_key_dots(sc, dcs, vals + ((int64_t)b * Tk + k0) * Dv, Dv, nk, 0.f, [&](int j) { return ps[j] != 0.f; });
  __syncthreads();
  float dl = 0.f;
  for (int j = threadIdx.x; j < nk; j += 256) dl += sc[j] * ps[j];
  dl = block_sum(dl, scratch);
  las_cluster_wait();
  if (threadIdx.x < LAS_CS) las_st_remote(&stat[rank], threadIdx.x, dl);
  las_cluster_sync();
  float delta = 0.f;
#pragma unroll
  for (int r = 0; r < LAS_CS; ++r) delta += stat[r];
  for (int j = threadIdx.x; j < nk; j += 256) {
    const float g = ps[j] * (sc[j] - delta);
    sc[j] = g;
    dscore[(int64_t)b * Tk + k0 + j] = g;
  }
  __syncthreads();
  las_weighted_rows(part, sc, wk + ((int64_t)b * Tk + k0) * D, D, nk);
  __syncthreads();
  const int cpo = D / LAS_CS;
  for (int c = threadIdx.x; c < D; c += 256)
    las_st_remote(&red[rank * cpo + (c % cpo)], (uint32_t)(c / cpo), (part[c] + part[D + c]) + (part[2 * D + c] + part[3 * D + c]));
  las_cluster_sync();
  for (int c = threadIdx.x; c < cpo; c += 256) {
    float a = 0.f;
#pragma unroll
    for (int r = 0; r < LAS_CS; ++r) a += red[r * cpo + c];
    dq[(int64_t)b * D + rank * cpo + c] = __float2bfloat16_rn(a);
  }
}

template <typename T>
__global__ void argmax_rows_kernel(const T* __restrict__ x, int64_t ld, int cols,
                                   int64_t* __restrict__ idx, int64_t idx_stride) {
  __shared__ float sv[32];
  __shared__ int si[32];
  const int64_t r = blockIdx.x;
  const T* xr = x + r * ld;
  float mx = -INFINITY;
  int mi = 0x7fffffff;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    const float v = to_f(xr[c]);
    if (v > mx) { mx = v; mi = c; }
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
    if (ov > mx || (ov == mx && oi < mi)) { mx = ov; mi = oi; }
  }
  if (lane == 0) { sv[w] = mx; si[w] = mi; }
  __syncthreads();
  if (w == 0) {
    float v = lane < nw ? sv[lane] : -INFINITY;
    int i = lane < nw ? si[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, i, o);
      if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
    }
    if (lane == 0) idx[r * idx_stride] = (i == 0x7fffffff) ? 0 : i;
  }
}


// Vectorised row arg-max (cols % 8 == 0, 16-byte aligned rows) with the LAS decode-length rule fused in:
// if lengths != nullptr: (sym in {EOS, PAD} and lengths[r] > step) -> lengths[r] = step + 1   (Dec.py:334-340).
template <typename T>
__global__ void __launch_bounds__(256)
argmax_rows_vec_kernel(const T* __restrict__ x, int64_t ld, int cols, int64_t* __restrict__ idx, int64_t idx_stride,
                       int32_t* __restrict__ lengths, int step, const float* __restrict__ table, T* __restrict__ emb,
                       int64_t ld_emb, int dim, const T* __restrict__ table2, T* __restrict__ out2, int64_t ld_out2,
                       int dim2) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float sv[32];
  __shared__ int si[32];
  __shared__ int s_sym;
  const int64_t r = blockIdx.x;
  const T* xr = x + r * ld;
  float mx = -INFINITY;
  int mi = 0x7fffffff;
  for (int c = threadIdx.x * 8; c < cols; c += 256 * 8) {
    float v[8];
    load8(xr + c, v);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (v[i] > mx) { mx = v[i]; mi = c + i; }
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
    if (ov > mx || (ov == mx && oi < mi)) { mx = ov; mi = oi; }
  }
  if (lane == 0) { sv[w] = mx; si[w] = mi; }
  __syncthreads();
  if (w == 0) {
    float v = lane < 8 ? sv[lane] : -INFINITY;
    int i = lane < 8 ? si[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, i, o);
      if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
    }
    if (lane == 0) {
      const int sym = (i == 0x7fffffff) ? 0 : i;
      idx[r * idx_stride] = sym;
      s_sym = sym;
      if (lengths && (sym == 3 /*EOS*/ || sym == 0 /*PAD*/) && lengths[r] > step) lengths[r] = step + 1;
    }
  }
  if (table) {        // the free-running decoder feeds this token's embedding to the next step (Dec.py:341): same launch
    __syncthreads();
    const float* src = table + (int64_t)s_sym * dim;
    T* dst = emb + r * ld_emb;
    for (int c = threadIdx.x; c < dim; c += 256) dst[c] = from_f<T>(src[c]);
  }
  if (table2) {       // ... and the token's row of a second table in the activation dtype (the pre-multiplied first-layer
    __syncthreads();  // gate contribution E W_ih^T + b of the next decoder step: a gather instead of a GEMM on the chain)
    const T* src = table2 + (int64_t)s_sym * dim2;
    T* dst = out2 + r * ld_out2;
    for (int c = threadIdx.x; c < dim2; c += 256) dst[c] = src[c];
  }
}

__global__ void las_update_lengths_kernel(const int64_t* __restrict__ sym, int64_t sym_stride,
                                          int32_t* __restrict__ lengths, int step, int64_t B) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int64_t s = sym[b * sym_stride];
  if ((s == 3 /*EOS*/ || s == 0 /*PAD*/) && lengths[b] > step) lengths[b] = step + 1;  // Dec.py:334-340
}

// ------------------------------------------------------------------------------------------------------------------
// Key / value gradients of the LAS attention over ALL decode steps at once (attention.py:203-289 in reverse):
//   out[b, t, :] = sum_s w[s, b, t] * x[s, b, :]        w fp32 [S, B, Tk] (dscore or probs), x [S, B, D], out [B, Tk, D]
// i.e. per sequence a [Tk x S] . [S x D] product with S ~ 31: far too thin for a tensor-core tile, and as a batched GEMM on
// the CUDA-core fallback it took 36 us per call on the backward critical path (the BLSTM backward waits for it).  Here a
// CTA owns (sequence, 16 keys): the weights sit in shared memory, a thread owns two adjacent columns, x is streamed once per
// CTA with 4-byte loads.  Writes dominate: B * Tk * D * sizeof.
template <typename T>
__global__ void __launch_bounds__(256) las_stack_grad_kernel(const float* __restrict__ w, const T* __restrict__ x,
                                                             T* __restrict__ out, int S, int B, int Tk, int D) {
  pdl_wait();
  pdl_launch_dependents();
  constexpr int TT = 16, SMAX = 64;
  __shared__ float ws[SMAX][TT];
  const int b = blockIdx.x, t0 = blockIdx.y * TT;
  for (int s0 = 0; s0 < S; s0 += SMAX) {
    const int sn = min(SMAX, S - s0);
    __syncthreads();
    for (int i = threadIdx.x; i < sn * TT; i += blockDim.x) {
      const int s = i / TT, tt = i % TT;
      ws[s][tt] = (t0 + tt < Tk) ? w[((int64_t)(s0 + s) * B + b) * Tk + t0 + tt] : 0.f;
    }
    __syncthreads();
    for (int c = threadIdx.x * 2; c < D; c += blockDim.x * 2) {
      float a0[TT], a1[TT];
#pragma unroll
      for (int tt = 0; tt < TT; ++tt) { a0[tt] = 0.f; a1[tt] = 0.f; }
      const bool pair = c + 1 < D;
      for (int s = 0; s < sn; ++s) {
        const T* xp = x + ((int64_t)(s0 + s) * B + b) * D + c;
        const float x0 = to_f(xp[0]), x1 = pair ? to_f(xp[1]) : 0.f;
#pragma unroll
        for (int tt = 0; tt < TT; ++tt) { a0[tt] = fmaf(ws[s][tt], x0, a0[tt]); a1[tt] = fmaf(ws[s][tt], x1, a1[tt]); }
      }
#pragma unroll
      for (int tt = 0; tt < TT; ++tt) {
        if (t0 + tt >= Tk) break;
        T* op = out + ((int64_t)b * Tk + t0 + tt) * D + c;
        if (s0 == 0) { op[0] = from_f<T>(a0[tt]); if (pair) op[1] = from_f<T>(a1[tt]); }
        else { op[0] = from_f<T>(to_f(op[0]) + a0[tt]); if (pair) op[1] = from_f<T>(to_f(op[1]) + a1[tt]); }
      }
    }
  }
}

}  // namespace b200st

using namespace b200st;

extern "C" {

static int mha_fwd_impl(int dtype, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                        int64_t ldv, const uint8_t* mask, int64_t mask_sb, int64_t mask_sq, void* o,
                        int64_t ldo, void* p, int64_t B, int64_t H, int64_t Lq, int64_t Lk, int64_t d,
                        float temperature, float drop_p, const int64_t* rng, int64_t site, b200st_stream_t stream) {
  if (B <= 0 || Lq <= 0) return 0;
  if (Lk <= 0) return set_error("mha_fwd: empty key sequence");
  if (drop_p > 0.f && (rng == nullptr || !(drop_p < 1.f))) return set_error("mha_fwd: dropout needs rng state and p < 1");
  if (dtype == B200ST_BF16 && g_mha_backend == 0 && drop_p <= 0.f) {
    const int rc = mha_fwd_tc(q, ldq, k, ldk, v, ldv, mask, mask_sb, mask_sq, o, ldo, p, B, H, Lq, Lk, d, temperature,
                              (cudaStream_t)stream);
    if (rc <= 0) return rc;           // launched (0) or failed (-1); 1 = shape not covered, use the SIMT tiles
  }
  {
    const size_t tsm = mha_fwd_tiled_smem(Lq, Lk, d);
    if (tsm <= 100 * 1024 && d % 4 == 0) {     // training shapes: whole head in shared memory
      dim3 tg((unsigned)H, (unsigned)B);
      B200ST_DISPATCH(dtype, T, {
        if (tsm > 48 * 1024)
          B200ST_CUDA(cudaFuncSetAttribute((const void*)mha_fwd_tiled_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm));
        B200ST_CUDA(launch_pdl(mha_fwd_tiled_kernel<T>, tg, dim3(MHA_T_THREADS), tsm, (cudaStream_t)stream,
                               (const T*)q, ldq, (const T*)k, ldk, (const T*)v, ldv, mask, mask_sb, mask_sq, (T*)o, ldo,
                               (T*)p, (int)H, (int)Lq, (int)Lk, (int)d, temperature, drop_p, rng, site));
      });
      B200ST_LAUNCH_CHECK("mha_fwd_tiled");
      return 0;
    }
  }
  if (drop_p > 0.f) return set_error("mha_fwd: attention dropout is implemented by the shared-memory tile kernel only (Lq=%lld Lk=%lld d=%lld does not fit)", (long long)Lq, (long long)Lk, (long long)d);
  const size_t smem = MHA_WARPS * (d + Lk) * sizeof(float);
  if (smem > 48 * 1024) return set_error("mha_fwd: Lk=%lld d=%lld exceeds the single-pass kernel", (long long)Lk, (long long)d);
  dim3 grid((unsigned)ceil_div(Lq, MHA_WARPS), (unsigned)H, (unsigned)B);
  B200ST_DISPATCH(dtype, T, {
    mha_fwd_kernel<T><<<grid, MHA_WARPS * 32, smem, (cudaStream_t)stream>>>(
        (const T*)q, ldq, (const T*)k, ldk, (const T*)v, ldv, mask, mask_sb, mask_sq, (T*)o, ldo,
        (T*)p, (int)H, (int)Lq, (int)Lk, (int)d, temperature);
  });
  B200ST_LAUNCH_CHECK("mha_fwd");
  return 0;
}

int b200st_mha_fwd(int dtype, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                   int64_t ldv, const uint8_t* mask, int64_t mask_sb, int64_t mask_sq, void* o,
                   int64_t ldo, void* p, int64_t B, int64_t H, int64_t Lq, int64_t Lk, int64_t d,
                   float temperature, b200st_stream_t stream) {
  return mha_fwd_impl(dtype, q, ldq, k, ldk, v, ldv, mask, mask_sb, mask_sq, o, ldo, p, B, H, Lq, Lk, d, temperature,
                      0.f, nullptr, 0, stream);
}

int b200st_mha_fwd_dropout(int dtype, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                           int64_t ldv, const uint8_t* mask, int64_t mask_sb, int64_t mask_sq, void* o,
                           int64_t ldo, void* p, int64_t B, int64_t H, int64_t Lq, int64_t Lk, int64_t d,
                           float temperature, float drop_p, const int64_t* rng, int64_t site,
                           b200st_stream_t stream) {
  return mha_fwd_impl(dtype, q, ldq, k, ldk, v, ldv, mask, mask_sb, mask_sq, o, ldo, p, B, H, Lq, Lk, d, temperature,
                      drop_p, rng, site, stream);
}

int b200st_set_mha_backend(int mode) {
  const int old = g_mha_backend | (g_las_backend << 1);
  if (mode >= 0 && mode <= 3) { g_mha_backend = mode & 1; g_las_backend = (mode >> 1) & 1; }
  return old;
}

static int mha_bwd_impl(int dtype, const void* dout, int64_t ldo, const void* q, int64_t ldq,
                        const void* k, int64_t ldk, const void* v, int64_t ldv, const void* p, void* ds,
                        void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, int64_t B,
                        int64_t H, int64_t Lq, int64_t Lk, int64_t d, float temperature, float drop_p,
                        const int64_t* rng, int64_t site, b200st_stream_t stream) {
  if (B <= 0 || Lq <= 0 || Lk <= 0) return 0;
  if (drop_p > 0.f && (rng == nullptr || !(drop_p < 1.f))) return set_error("mha_bwd: dropout needs rng state and p < 1");
  if (dtype == B200ST_BF16 && g_mha_backend == 0 && drop_p <= 0.f) {
    const int rc = mha_bwd_tc(dout, ldo, q, ldq, k, ldk, v, ldv, p, dq, lddq, dk, lddk, dv, lddv, B, H, Lq, Lk, d,
                              temperature, (cudaStream_t)stream);
    if (rc <= 0) return rc;
  }
  {
    const size_t tsm = mha_bwd_tiled_smem(Lq, Lk, d);
    if (tsm <= 160 * 1024 && d % 4 == 0) {
      dim3 tg((unsigned)H, (unsigned)B);
      B200ST_DISPATCH(dtype, T, {
        if (tsm > 48 * 1024)
          B200ST_CUDA(cudaFuncSetAttribute((const void*)mha_bwd_tiled_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm));
        B200ST_CUDA(launch_pdl(mha_bwd_tiled_kernel<T>, tg, dim3(MHA_T_THREADS), tsm, (cudaStream_t)stream,
                               (const T*)dout, ldo, (const T*)q, ldq, (const T*)k, ldk, (const T*)v, ldv, (const T*)p,
                               (T*)dq, lddq, (T*)dk, lddk, (T*)dv, lddv, (int)H, (int)Lq, (int)Lk, (int)d, temperature,
                               drop_p, rng, site));
      });
      B200ST_LAUNCH_CHECK("mha_bwd_tiled");
      return 0;
    }
  }
  if (drop_p > 0.f) return set_error("mha_bwd: attention dropout is implemented by the shared-memory tile kernel only (Lq=%lld Lk=%lld d=%lld does not fit)", (long long)Lq, (long long)Lk, (long long)d);
  const size_t smem = MHA_WARPS * (d + Lk) * sizeof(float);
  if (smem > 48 * 1024) return set_error("mha_bwd: Lk=%lld d=%lld exceeds the single-pass kernel", (long long)Lk, (long long)d);
  dim3 gq((unsigned)ceil_div(Lq, MHA_WARPS), (unsigned)H, (unsigned)B);
  dim3 gk((unsigned)ceil_div(Lk, MHA_WARPS), (unsigned)H, (unsigned)B);
  B200ST_DISPATCH(dtype, T, {
    mha_bwd_q_kernel<T><<<gq, MHA_WARPS * 32, smem, (cudaStream_t)stream>>>(
        (const T*)dout, ldo, (const T*)k, ldk, (const T*)v, ldv, (const T*)p, (T*)ds, (T*)dq, lddq,
        (int)H, (int)Lq, (int)Lk, (int)d, temperature);
    mha_bwd_kv_kernel<T><<<gk, MHA_WARPS * 32, 0, (cudaStream_t)stream>>>(
        (const T*)dout, ldo, (const T*)q, ldq, (const T*)p, (const T*)ds, (T*)dk, lddk, (T*)dv, lddv,
        (int)H, (int)Lq, (int)Lk, (int)d, temperature);
  });
  B200ST_LAUNCH_CHECK("mha_bwd");
  count_launch();
  return 0;
}

int b200st_mha_bwd(int dtype, const void* dout, int64_t ldo, const void* q, int64_t ldq,
                   const void* k, int64_t ldk, const void* v, int64_t ldv, const void* p, void* ds,
                   void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, int64_t B,
                   int64_t H, int64_t Lq, int64_t Lk, int64_t d, float temperature,
                   b200st_stream_t stream) {
  return mha_bwd_impl(dtype, dout, ldo, q, ldq, k, ldk, v, ldv, p, ds, dq, lddq, dk, lddk, dv, lddv, B, H, Lq, Lk, d,
                      temperature, 0.f, nullptr, 0, stream);
}

int b200st_mha_bwd_dropout(int dtype, const void* dout, int64_t ldo, const void* q, int64_t ldq,
                           const void* k, int64_t ldk, const void* v, int64_t ldv, const void* p, void* ds,
                           void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, int64_t B,
                           int64_t H, int64_t Lq, int64_t Lk, int64_t d, float temperature, float drop_p,
                           const int64_t* rng, int64_t site, b200st_stream_t stream) {
  return mha_bwd_impl(dtype, dout, ldo, q, ldq, k, ldk, v, ldv, p, ds, dq, lddq, dk, lddk, dv, lddv, B, H, Lq, Lk, d,
                      temperature, drop_p, rng, site, stream);
}

int b200st_mha_decode(int dtype, const void* q, int64_t ldq, const void* k_cache, const void* v_cache,
                      int64_t stride_b, int64_t stride_t, const int32_t* anc, int64_t n_hyp, int64_t bdiv,
                      const uint8_t* mask, int64_t mask_sb, int64_t mask_bdiv, void* o, int64_t ldo, int64_t H,
                      int64_t Lk, int64_t d, float temperature, b200st_stream_t stream) {
  if (n_hyp <= 0) return 0;
  if (Lk <= 0) return set_error("mha_decode: empty key cache");
  if (bdiv <= 0 || mask_bdiv <= 0) return set_error("mha_decode: bdiv / mask_bdiv must be >= 1");
  const size_t smem = (size_t)(d + 2 * Lk) * sizeof(float);
  if (smem > 200 * 1024) return set_error("mha_decode: Lk=%lld exceeds the shared-memory score row", (long long)Lk);
  dim3 grid((unsigned)H, (unsigned)n_hyp);
  B200ST_DISPATCH(dtype, T, {
    if (smem > 48 * 1024)
      B200ST_CUDA(cudaFuncSetAttribute((const void*)mha_decode_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B200ST_CUDA(launch_pdl(mha_decode_kernel<T>, grid, dim3(128), smem, (cudaStream_t)stream, (const T*)q, ldq,
                           (const T*)k_cache, (const T*)v_cache, stride_b, stride_t, anc, (int)n_hyp, (int)bdiv, mask,
                           mask_sb, (int)mask_bdiv, (T*)o, ldo, (int)Lk, (int)d, temperature));
  });
  B200ST_LAUNCH_CHECK("mha_decode");
  return 0;
}

int b200st_las_attn_fwd(int dtype, const void* q, const void* wk, const void* vals,
                        const int32_t* klens, void* ctx, float* probs, int64_t B, int64_t Tk, int64_t D,
                        int64_t Dv, b200st_stream_t stream) {
  if (B <= 0) return 0;
  if (dtype == B200ST_BF16 && g_las_backend == 0 && D % 8 == 0 && Dv % (8 * LAS_CS) == 0 && Tk <= 64 * LAS_CS && Tk >= LAS_CS &&
      B <= 65535 && ((uintptr_t)wk & 15) == 0 && ((uintptr_t)vals & 15) == 0) {
    const size_t cs = (D + 64 + 5 * Dv + 2 * LAS_CS) * sizeof(float);
    if (cs <= 48 * 1024) {
      B200ST_CUDA(launch_pdl(las_attn_fwd_cl_kernel, dim3(LAS_CS, (unsigned)B), dim3(256), cs, (cudaStream_t)stream,
                             (const __nv_bfloat16*)q, (const __nv_bfloat16*)wk, (const __nv_bfloat16*)vals, klens,
                             (__nv_bfloat16*)ctx, probs, (int)Tk, (int)D, (int)Dv));
      B200ST_LAUNCH_CHECK("las_attn_fwd_cl");
      return 0;
    }
  }
  {
    const size_t vs = (D + ((Tk + 3) & ~3) + 4 * Dv) * sizeof(float);
    if (D % 8 == 0 && Dv % 8 == 0 && vs <= 48 * 1024) {
      B200ST_DISPATCH(dtype, T, {
        B200ST_CUDA(launch_pdl(las_attn_fwd_vec_kernel<T>, dim3((unsigned)B), dim3(256), vs, (cudaStream_t)stream,
                               (const T*)q, (const T*)wk, (const T*)vals, klens, (T*)ctx, probs, (int)Tk, (int)D, (int)Dv));
      });
      B200ST_LAUNCH_CHECK("las_attn_fwd_vec");
      return 0;
    }
  }
  const size_t smem = (D + Tk) * sizeof(float);
  if (smem > 48 * 1024) return set_error("las_attn_fwd: D+Tk too large");
  B200ST_DISPATCH(dtype, T, {
    las_attn_fwd_kernel<T><<<(unsigned)B, 256, smem, (cudaStream_t)stream>>>(
        (const T*)q, (const T*)wk, (const T*)vals, klens, (T*)ctx, probs, (int)Tk, (int)D, (int)Dv);
  });
  B200ST_LAUNCH_CHECK("las_attn_fwd");
  return 0;
}

int b200st_las_attn_bwd(int dtype, const void* dctx, const void* wk, const void* vals,
                        const float* probs, float* dscore, void* dq, int64_t B, int64_t Tk, int64_t D,
                        int64_t Dv, b200st_stream_t stream) {
  if (B <= 0) return 0;
  if (dtype == B200ST_BF16 && g_las_backend == 0 && D % (8 * LAS_CS) == 0 && Dv % 8 == 0 && Tk <= 64 * LAS_CS && Tk >= LAS_CS &&
      B <= 65535 && ((uintptr_t)wk & 15) == 0 && ((uintptr_t)vals & 15) == 0) {
    const size_t cs = (Dv + 128 + 5 * D + LAS_CS) * sizeof(float);
    if (cs <= 48 * 1024) {
      B200ST_CUDA(launch_pdl(las_attn_bwd_cl_kernel, dim3(LAS_CS, (unsigned)B), dim3(256), cs, (cudaStream_t)stream,
                             (const __nv_bfloat16*)dctx, (const __nv_bfloat16*)wk, (const __nv_bfloat16*)vals, probs,
                             dscore, (__nv_bfloat16*)dq, (int)Tk, (int)D, (int)Dv));
      B200ST_LAUNCH_CHECK("las_attn_bwd_cl");
      return 0;
    }
  }
  {
    const size_t vs = (Dv + ((Tk + 3) & ~3) + 4 * D) * sizeof(float);
    if (D % 8 == 0 && Dv % 8 == 0 && vs <= 48 * 1024) {
      B200ST_DISPATCH(dtype, T, {
        B200ST_CUDA(launch_pdl(las_attn_bwd_vec_kernel<T>, dim3((unsigned)B), dim3(256), vs, (cudaStream_t)stream,
                               (const T*)dctx, (const T*)wk, (const T*)vals, probs, dscore, (T*)dq, (int)Tk, (int)D, (int)Dv));
      });
      B200ST_LAUNCH_CHECK("las_attn_bwd_vec");
      return 0;
    }
  }
  const size_t smem = (Dv + Tk) * sizeof(float);
  if (smem > 48 * 1024) return set_error("las_attn_bwd: Dv+Tk too large");
  B200ST_DISPATCH(dtype, T, {
    las_attn_bwd_kernel<T><<<(unsigned)B, 256, smem, (cudaStream_t)stream>>>(
        (const T*)dctx, (const T*)wk, (const T*)vals, probs, dscore, (T*)dq, (int)Tk, (int)D, (int)Dv);
  });
  B200ST_LAUNCH_CHECK("las_attn_bwd");
  return 0;
}

int b200st_argmax_rows(int dtype, const void* x, int64_t ld, int64_t rows, int64_t cols, int64_t* idx,
                       int64_t idx_stride, b200st_stream_t stream) {
  if (rows <= 0) return 0;
  B200ST_DISPATCH(dtype, T, {
    argmax_rows_kernel<T><<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>((const T*)x, ld, (int)cols,
                                                                            idx, idx_stride);
  });
  B200ST_LAUNCH_CHECK("argmax_rows");
  return 0;
}

int b200st_argmax_rows_lengths(int dtype, const void* x, int64_t ld, int64_t rows, int64_t cols, int64_t* idx,
                               int64_t idx_stride, int32_t* lengths, int step, b200st_stream_t stream) {
  return b200st_argmax_rows_embed(dtype, x, ld, rows, cols, idx, idx_stride, lengths, step, nullptr, nullptr, 0, 0, stream);
}

int b200st_argmax_rows_embed(int dtype, const void* x, int64_t ld, int64_t rows, int64_t cols, int64_t* idx,
                             int64_t idx_stride, int32_t* lengths, int step, const float* table, void* emb,
                             int64_t ld_emb, int64_t dim, b200st_stream_t stream) {
  return b200st_argmax_rows_embed2(dtype, x, ld, rows, cols, idx, idx_stride, lengths, step, table, emb, ld_emb, dim,
                                   nullptr, nullptr, 0, 0, stream);
}

int b200st_argmax_rows_embed2(int dtype, const void* x, int64_t ld, int64_t rows, int64_t cols, int64_t* idx,
                              int64_t idx_stride, int32_t* lengths, int step, const float* table, void* emb,
                              int64_t ld_emb, int64_t dim, const void* table2, void* out2, int64_t ld_out2,
                              int64_t dim2, b200st_stream_t stream) {
  if (rows <= 0) return 0;
  const bool vec = cols % 8 == 0 && ld % 8 == 0 && ((uintptr_t)x & 15) == 0;
  if (vec) {
    B200ST_DISPATCH(dtype, T, {
      B200ST_CUDA(launch_pdl(argmax_rows_vec_kernel<T>, dim3((unsigned)rows), dim3(256), 0, (cudaStream_t)stream,
                             (const T*)x, ld, (int)cols, idx, idx_stride, lengths, step, table, (T*)emb, ld_emb, (int)dim,
                             (const T*)table2, (T*)out2, ld_out2, (int)dim2));
    });
    B200ST_LAUNCH_CHECK("argmax_rows_vec");
    return 0;
  }
  if (b200st_argmax_rows(dtype, x, ld, rows, cols, idx, idx_stride, stream)) return -1;
  if (lengths && b200st_las_update_lengths(idx, idx_stride, lengths, step, rows, stream)) return -1;
  if (table2) return set_error("argmax_rows_embed2: the second gather needs the vectorised route (cols %% 8 == 0, aligned rows)");
  if (table) {
    if (idx_stride != 1) return set_error("argmax_rows_embed: the unfused route needs dense ids");
    return b200st_embedding_fwd(dtype, idx, table, emb, ld_emb, rows, dim, cols, stream);
  }
  return 0;
}

int b200st_las_update_lengths(const int64_t* sym, int64_t sym_stride, int32_t* lengths, int step,
                              int64_t B, b200st_stream_t stream) {
  if (B <= 0) return 0;
  las_update_lengths_kernel<<<(unsigned)ceil_div(B, 128), 128, 0, (cudaStream_t)stream>>>(
      sym, sym_stride, lengths, step, B);
  B200ST_LAUNCH_CHECK("las_update_lengths");
  return 0;
}

int b200st_las_stack_grad(int dtype, const float* w, const void* x, void* out, int64_t S, int64_t B, int64_t Tk, int64_t D,
                          b200st_stream_t stream) {
  if (S <= 0 || B <= 0 || Tk <= 0 || D <= 0) return 0;
  dim3 grid((unsigned)B, (unsigned)ceil_div(Tk, 16));
  B200ST_DISPATCH(dtype, T, {
    B200ST_CUDA(launch_pdl(las_stack_grad_kernel<T>, grid, dim3(256), 0, (cudaStream_t)stream, w, (const T*)x, (T*)out, (int)S,
                           (int)B, (int)Tk, (int)D));
  });
  B200ST_LAUNCH_CHECK("las_stack_grad");
  return 0;
}

}  // extern "C"
