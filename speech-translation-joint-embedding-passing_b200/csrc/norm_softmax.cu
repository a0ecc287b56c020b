// HBM-bound row kernels: LayerNorm fwd/bwd, log-softmax fwd/bwd (+argmax), masked NLL, and the fused
// softmax + NLL + gradient kernel (K17).  One row is owned by one warp (LayerNorm, D <= a few K) or one
// CTA with the row staged in shared memory (vocabulary rows), so HBM sees each element once per pass.
#include "common.cuh"

namespace b200st {

// ------------------------------------------------------------------------------------------------
// LayerNorm: y = (x - mean) * rstd * gamma + beta, biased variance, two-pass statistics in fp32.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void layernorm_fwd_kernel(const T* __restrict__ x, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, T* __restrict__ y,
                                     float* __restrict__ mean, float* __restrict__ rstd, int64_t rows,
                                     int cols, float eps) {
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const T* xr = x + row * cols;
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) s += to_f(xr[c]);
  const float mu = warp_sum(s) / cols;
  float v = 0.f;
  for (int c = lane; c < cols; c += 32) { const float d = to_f(xr[c]) - mu; v += d * d; }
  const float rs = rsqrtf(warp_sum(v) / cols + eps);
  T* yr = y + row * cols;
  for (int c = lane; c < cols; c += 32)
    yr[c] = from_f<T>((to_f(xr[c]) - mu) * rs * gamma[c] + beta[c]);
  if (lane == 0) {
    if (mean) mean[row] = mu;
    if (rstd) rstd[row] = rs;
  }
}

// Vectorised variant for cols = 256 * NV (16-byte aligned rows): the row is read ONCE with 16-byte loads and stays in
// registers for both statistics passes and the normalisation (the scalar kernel above re-reads it three times with 2-byte
// loads).  Warp per row.
__device__ __forceinline__ void ln_load8(const float* p, float* v) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ln_load8(const __nv_bfloat16* p, float* v) {
  const uint4 a = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
__device__ __forceinline__ void ln_store8(float* p, const float* v) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void ln_store8(__nv_bfloat16* p, const float* v) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&t);
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}
template <typename T, int NV>
__global__ void __launch_bounds__(256) layernorm_fwd_vec_kernel(const T* __restrict__ x, const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, T* __restrict__ y,
                                                                float* __restrict__ mean, float* __restrict__ rstd,
                                                                int64_t rows, float eps) {
  constexpr int cols = 256 * NV;
  const int lane = threadIdx.x & 31;
  // gamma / beta do not depend on the predecessor kernel: fetched before the dependency wait
  float g[NV][8], b[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i) { ln_load8(gamma + (lane + 32 * i) * 8, g[i]); ln_load8(beta + (lane + 32 * i) * 8, b[i]); }
  pdl_wait();
  pdl_launch_dependents();
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float v[NV][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    ln_load8(x + row * cols + (lane + 32 * i) * 8, v[i]);
#pragma unroll
    for (int e = 0; e < 8; ++e) s += v[i][e];
  }
  const float mu = warp_sum(s) / cols;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) { const float d = v[i][e] - mu; q += d * d; }
  const float rs = rsqrtf(warp_sum(q) / cols + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[i][e] = (v[i][e] - mu) * rs * g[i][e] + b[i][e];
    ln_store8(y + row * cols + (lane + 32 * i) * 8, v[i]);
  }
  if (lane == 0) {
    if (mean) mean[row] = mu;
    if (rstd) rstd[row] = rs;
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma;  dgamma += dy * xhat; dbeta += dy.
// Each CTA walks a strip of rows; column partials live in shared memory and are flushed once.
template <typename T>
__global__ void layernorm_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                     const float* __restrict__ gamma, const float* __restrict__ mean,
                                     const float* __restrict__ rstd, const T* __restrict__ add, T* __restrict__ dx,
                                     float* __restrict__ dgamma, float* __restrict__ dbeta,
                                     int64_t rows, int cols, int rows_per_block) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ float sm[];   // [2][cols]
  float* sg = sm;
  float* sb = sm + cols;
  for (int c = threadIdx.x; c < 2 * cols; c += blockDim.x) sm[c] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  for (int64_t row = r0 + w; row < r0 + rows_per_block && row < rows; row += nw) {
    const T* dyr = dy + row * cols;
    const T* xr = x + row * cols;
    const float mu = mean[row], rs = rstd[row];
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < cols; c += 32) {
      const float g = to_f(dyr[c]) * gamma[c];
      const float xh = (to_f(xr[c]) - mu) * rs;
      s1 += g;
      s2 += g * xh;
    }
    s1 = warp_sum(s1) / cols;
    s2 = warp_sum(s2) / cols;
    T* dxr = dx + row * cols;
    for (int c = lane; c < cols; c += 32) {
      const float d = to_f(dyr[c]);
      const float xh = (to_f(xr[c]) - mu) * rs;
      dxr[c] = from_f<T>(rs * (d * gamma[c] - s1 - xh * s2) + (add ? to_f(add[row * cols + c]) : 0.f));
      atomicAdd(&sg[c], d * xh);
      atomicAdd(&sb[c], d);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    atomicAdd(&dgamma[c], sg[c]);
    atomicAdd(&dbeta[c], sb[c]);
  }
}

// Register-accumulating variant for cols % 128 == 0, cols <= 128 * NV: lane l of every warp owns the column quads
// {(32 i + l) * 4 .. +3}, i < NV, so the dgamma / dbeta partial sums of all the rows a warp walks stay in registers
// (the strip kernel above pays two shared-memory atomics per ELEMENT); they are merged once per CTA.
template <typename T, int NV>
__global__ void __launch_bounds__(256)
layernorm_bwd_reg_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ gamma,
                         const float* __restrict__ mean, const float* __restrict__ rstd, const T* __restrict__ add,
                         T* __restrict__ dx,
                         float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t rows, int cols,
                         int rows_per_block, float* __restrict__ partials) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ float sm[];   // [2][cols]
  for (int c = threadIdx.x; c < 2 * cols; c += blockDim.x) sm[c] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int nq = cols >> 7;        // quads per lane actually used (<= NV)
  float gam[NV][4], ag[NV][4], ab[NV][4];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 g4 = i < nq ? *reinterpret_cast<const float4*>(gamma + (32 * i + lane) * 4) : make_float4(0, 0, 0, 0);
    gam[i][0] = g4.x; gam[i][1] = g4.y; gam[i][2] = g4.z; gam[i][3] = g4.w;
#pragma unroll
    for (int j = 0; j < 4; ++j) ag[i][j] = ab[i][j] = 0.f;
  }
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  for (int64_t row = r0 + w; row < r0 + rows_per_block && row < rows; row += nw) {
    const T* dyr = dy + row * cols;
    const T* xr = x + row * cols;
    const float mu = mean[row], rs = rstd[row];
    float d[NV][4], xh[NV][4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (i < nq) {
        load4(dyr + (32 * i + lane) * 4, d[i]);
        load4(xr + (32 * i + lane) * 4, xh[i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          xh[i][j] = (xh[i][j] - mu) * rs;
          const float g = d[i][j] * gam[i][j];
          s1 += g;
          s2 += g * xh[i][j];
        }
      }
    }
    s1 = warp_sum(s1) / cols;
    s2 = warp_sum(s2) / cols;
    T* dxr = dx + row * cols;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (i < nq) {
        float o[4], ad[4] = {0.f, 0.f, 0.f, 0.f};
        if (add) load4(add + row * cols + (32 * i + lane) * 4, ad);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          o[j] = rs * (d[i][j] * gam[i][j] - s1 - xh[i][j] * s2) + ad[j];
          ag[i][j] += d[i][j] * xh[i][j];
          ab[i][j] += d[i][j];
        }
        store4(dxr + (32 * i + lane) * 4, o);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (i < nq) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        atomicAdd(&sm[(32 * i + lane) * 4 + j], ag[i][j]);
        atomicAdd(&sm[cols + (32 * i + lane) * 4 + j], ab[i][j]);
      }
    }
  }
  __syncthreads();
  if (partials != nullptr) {
    // per-CTA partial sums [n_blocks][2 * cols], reduced by a separate (off-critical-path) column-sum launch: this
    // kernel then ends with one coalesced store instead of a 200-deep chain of same-address L2 atomics
    float* out = partials + (int64_t)blockIdx.x * 2 * cols;
    for (int c = threadIdx.x * 4; c < 2 * cols; c += blockDim.x * 4)
      *reinterpret_cast<float4*>(out + c) = *reinterpret_cast<const float4*>(sm + c);
    return;
  }
  // one 16-byte vector reduction per column quad: the per-address chain of same-location L2 atomics (one per CTA)
  // is what bounds this kernel once the row work is spread over many CTAs
  for (int c = threadIdx.x * 4; c < cols; c += blockDim.x * 4) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dgamma + c), "f"(sm[c]), "f"(sm[c + 1]), "f"(sm[c + 2]), "f"(sm[c + 3]) : "memory");
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dbeta + c), "f"(sm[cols + c]), "f"(sm[cols + c + 1]), "f"(sm[cols + c + 2]), "f"(sm[cols + c + 3]) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// Vocabulary rows.  The row is staged once in shared memory as fp32.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void log_softmax_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int cols,
                                       int64_t* __restrict__ argmax) {
  extern __shared__ float row[];
  __shared__ float scratch[32];
  __shared__ int s_idx[32];
  const int64_t r = blockIdx.x;
  const T* xr = x + r * cols;
  float mx = -INFINITY;
  int mi = 0x7fffffff;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    const float v = to_f(xr[c]);
    row[c] = v;
    if (v > mx) { mx = v; mi = c; }
  }
  // arg-max with first-index tie break
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
    if (ov > mx || (ov == mx && oi < mi)) { mx = ov; mi = oi; }
  }
  if (lane == 0) { scratch[w] = mx; s_idx[w] = mi; }
  __syncthreads();
  if (w == 0) {
    float v = lane < nw ? scratch[lane] : -INFINITY;
    int i = lane < nw ? s_idx[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, i, o);
      if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
    }
    if (lane == 0) { scratch[0] = v; s_idx[0] = i; }
  }
  __syncthreads();
  mx = scratch[0];
  if (argmax && threadIdx.x == 0) argmax[r] = s_idx[0];
  float s = 0.f;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) s += expf(row[c] - mx);
  s = block_sum(s, scratch);
  const float lse = mx + logf(s);
  T* yr = y + r * cols;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) yr[c] = from_f<T>(row[c] - lse);
}

// dx = dy - exp(y) * sum(dy)
template <typename T>
__global__ void log_softmax_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y,
                                       T* __restrict__ dx, int cols) {
  extern __shared__ float row[];
  __shared__ float scratch[32];
  const int64_t r = blockIdx.x;
  const T* dyr = dy + r * cols;
  float s = 0.f;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    const float v = to_f(dyr[c]);
    row[c] = v;
    s += v;
  }
  s = block_sum(s, scratch);
  const T* yr = y + r * cols;
  T* dxr = dx + r * cols;
  for (int c = threadIdx.x; c < cols; c += blockDim.x)
    dxr[c] = from_f<T>(row[c] - expf(to_f(yr[c])) * s);
}

template <typename T>
__global__ void masked_nll_fwd_kernel(const T* __restrict__ logp, int64_t ld,
                                      const int64_t* __restrict__ target,
                                      const uint8_t* __restrict__ mask, float* loss_sum, int64_t rows,
                                      int64_t cols) {
  __shared__ float scratch[32];
  float s = 0.f;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows;
       r += (int64_t)gridDim.x * blockDim.x) {
    if (mask && !mask[r]) continue;
    const int64_t t = target[r];
    if (t >= 0 && t < cols) s -= to_f(logp[r * ld + t]);
  }
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) atomicAdd(loss_sum, s);
}

template <typename T>
__global__ void masked_nll_bwd_kernel(const float* __restrict__ gscale,
                                      const int64_t* __restrict__ target,
                                      const uint8_t* __restrict__ mask, T* __restrict__ dlogp,
                                      int64_t ld, int64_t cols) {
  const int64_t r = blockIdx.x;
  const bool keep = !mask || mask[r];
  const int64_t t = target[r];
  const float g = -gscale[0];
  T* dr = dlogp + r * ld;
  for (int64_t c = threadIdx.x; c < cols; c += blockDim.x)
    dr[c] = from_f<T>((keep && c == t) ? g : 0.f);
}

// K17: loss_sum += mask * (lse - (1-eps) * x[t] - eps/V * sum(x));  dx = mask * scale * (softmax - q),
// q = (1-eps) * onehot + eps/V.  eps = 0 reproduces log_softmax + NLLLoss (loss.py:130-132).
template <typename T>
__global__ void softmax_nll_fused_kernel(const T* __restrict__ x, int64_t ld,
                                         const int64_t* __restrict__ target,
                                         const uint8_t* __restrict__ mask,
                                         const float* __restrict__ scale, float eps, float* loss_sum,
                                         T* dx, int64_t ld_d, int cols) {
  extern __shared__ float row[];
  __shared__ float scratch[32];
  const int64_t r = blockIdx.x;
  const bool keep = !mask || mask[r];
  T* dr = dx + r * ld_d;
  if (!keep) {   // masked rows contribute nothing; never read the logits
    for (int c = threadIdx.x; c < cols; c += blockDim.x) dr[c] = from_f<T>(0.f);
    return;
  }
  const T* xr = x + r * ld;
  float mx = -INFINITY, sx = 0.f;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    const float v = to_f(xr[c]);
    row[c] = v;
    mx = fmaxf(mx, v);
    sx += v;
  }
  mx = block_max(mx, scratch);
  float s = 0.f;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) s += expf(row[c] - mx);
  s = block_sum(s, scratch);
  const float lse = mx + logf(s);
  const int64_t t = target[r];
  const float u = eps / cols;
  if (eps != 0.f) sx = block_sum(sx, scratch);
  if (threadIdx.x == 0) {
    float l = lse - (1.f - eps) * row[t];
    if (eps != 0.f) l -= u * sx;
    atomicAdd(loss_sum, l);
  }
  const float sc = scale[0];
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    const float p = expf(row[c] - lse);
    const float q = (c == t ? 1.f - eps : 0.f) + u;
    dr[c] = from_f<T>(sc * (p - q));
  }
}

// Register-resident variant for rows of up to 512 threads x NV x 8 elements with 16-byte-aligned rows (the training
// shape: V = 10 000 bf16 -> 3 vectors per thread): the row is read ONCE with 16-byte loads, never staged in shared
// memory, exp() is evaluated once per element, and the gradient leaves with 16-byte stores.
template <int NV>
__global__ void __launch_bounds__(512)
softmax_nll_fused_vec_kernel(const __nv_bfloat16* __restrict__ x, int64_t ld, const int64_t* __restrict__ target,
                             const uint8_t* __restrict__ mask, const float* __restrict__ scale, float eps,
                             float* loss_sum, __nv_bfloat16* dx, int64_t ld_d, int cols) {
  __shared__ float scratch[32];
  const int64_t r = blockIdx.x;
  const bool keep = !mask || mask[r];
  __nv_bfloat16* dr = dx + r * ld_d;
  const int nvec = cols >> 3;                       // cols % 8 == 0 (checked by the launcher)
  if (!keep) {
    for (int i = threadIdx.x; i < nvec; i += 512) *reinterpret_cast<uint4*>(dr + 8 * i) = make_uint4(0, 0, 0, 0);
    return;
  }
  const __nv_bfloat16* xr = x + r * ld;
  float v[NV][8];
  float mx = -INFINITY, sx = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = k * 512 + threadIdx.x;
    if (i < nvec) {
      const uint4 q = *reinterpret_cast<const uint4*>(xr + 8 * i);
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        v[k][2 * e] = __uint_as_float(w[e] << 16);
        v[k][2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) { mx = fmaxf(mx, v[k][e]); sx += v[k][e]; }
    }
  }
  mx = block_max(mx, scratch);
  // exp(x - mx) as one FFMA + one MUFU.EX2 (relative error 2^-22, far inside the bf16 gradient this kernel writes): with expf the
  // kernel was bound by instruction issue, not by HBM
  const float nmx = -mx * kLog2e;
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    if (k * 512 + (int)threadIdx.x < nvec) {
#pragma unroll
      for (int e = 0; e < 8; ++e) { v[k][e] = ex2_approx(fmaf(v[k][e], kLog2e, nmx)); s += v[k][e]; }
    }
  }
  s = block_sum(s, scratch);
  const int64_t t = target[r];
  const float u = eps / cols;
  if (eps != 0.f) sx = block_sum(sx, scratch);
  if (threadIdx.x == 0) {
    float l = mx + logf(s) - (1.f - eps) * to_f(xr[t]);
    if (eps != 0.f) l -= u * sx;
    atomicAdd(loss_sum, l);
  }
  const float sc = scale[0], sci = sc / s;
  const float qo = -sc * u, qt = -sc * ((1.f - eps) + u);   // dx = sc * (p - q): q = u off the target, (1 - eps) + u on it
  const int ti = (int)t;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = k * 512 + threadIdx.x;
    if (i < nvec) {
      float o[8];
      const int te = ti - 8 * i;                    // position of the target inside this thread's 8 columns (or outside 0..7)
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = fmaf(v[k][e], sci, e == te ? qt : qo);
      uint4 w;
      __nv_bfloat162* ww = reinterpret_cast<__nv_bfloat162*>(&w);
#pragma unroll
      for (int e = 0; e < 4; ++e) ww[e] = __floats2bfloat162_rn(o[2 * e], o[2 * e + 1]);
      *reinterpret_cast<uint4*>(dr + 8 * i) = w;
    }
  }
}

static int set_row_smem(const void* fn, size_t bytes, const char* name) {
  if (bytes > 200 * 1024) return set_error("%s: row of %zu bytes does not fit shared memory", name, bytes);
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return set_error("%s: %s", name, cudaGetErrorString(e));
  }
  return 0;
}

}  // namespace b200st

using namespace b200st;

extern "C" {

int b200st_layernorm_fwd(int dtype, const void* x, const float* gamma, const float* beta, void* y,
                         float* mean, float* rstd, int64_t rows, int64_t cols, float eps,
                         b200st_stream_t stream) {
  if (rows <= 0) return 0;
  if ((cols == 256 || cols == 512 || cols == 1024) && ((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0 &&
      ((uintptr_t)gamma & 15) == 0 && ((uintptr_t)beta & 15) == 0) {
    const int wv = 8;
    B200ST_DISPATCH(dtype, T, {
      if (cols == 256)
        B200ST_CUDA(launch_pdl(layernorm_fwd_vec_kernel<T, 1>, dim3((unsigned)ceil_div(rows, wv)), dim3(wv * 32), 0,
                               (cudaStream_t)stream, (const T*)x, gamma, beta, (T*)y, mean, rstd, rows, eps));
      else if (cols == 512)
        B200ST_CUDA(launch_pdl(layernorm_fwd_vec_kernel<T, 2>, dim3((unsigned)ceil_div(rows, wv)), dim3(wv * 32), 0,
                               (cudaStream_t)stream, (const T*)x, gamma, beta, (T*)y, mean, rstd, rows, eps));
      else
        B200ST_CUDA(launch_pdl(layernorm_fwd_vec_kernel<T, 4>, dim3((unsigned)ceil_div(rows, wv)), dim3(wv * 32), 0,
                               (cudaStream_t)stream, (const T*)x, gamma, beta, (T*)y, mean, rstd, rows, eps));
    });
    B200ST_LAUNCH_CHECK("layernorm_fwd_vec");
    return 0;
  }
  const int wpb = 4;
  B200ST_DISPATCH(dtype, T, {
    B200ST_CUDA(launch_pdl(layernorm_fwd_kernel<T>, dim3((unsigned)ceil_div(rows, wpb)), dim3(wpb * 32), 0,
                           (cudaStream_t)stream, (const T*)x, gamma, beta, (T*)y, mean, rstd, rows, (int)cols, eps));
  });
  B200ST_LAUNCH_CHECK("layernorm_fwd");
  return 0;
}

static int layernorm_bwd_impl(int dtype, const void* dy, const void* x, const float* gamma,
                              const float* mean, const float* rstd, const void* add, void* dx, float* dgamma,
                              float* dbeta, int64_t rows, int64_t cols, float* partials, b200st_stream_t stream) {
  if (rows <= 0) return 0;
  const size_t smem = 2 * cols * sizeof(float);
  if (smem > 48 * 1024) return set_error("layernorm_bwd: cols %lld too large", (long long)cols);
  // ~2 CTAs per SM worth of row strips keeps the final atomics few while filling the chip.
  if (cols % 128 == 0 && cols <= 1024 && ((uintptr_t)dy & 15) == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)dx & 15) == 0 &&
      ((uintptr_t)gamma & 15) == 0 && ((uintptr_t)add & 15) == 0 && ((uintptr_t)dgamma & 15) == 0 && ((uintptr_t)dbeta & 15) == 0 &&
      ((uintptr_t)partials & 15) == 0) {
    const int rpb = 16;             // 8 warps x 2 rows (measured best: fewer rows per CTA lengthen the dgamma/dbeta atomic chains, more serialise the row loads)
    B200ST_DISPATCH(dtype, T, {
      if (cols <= 512) {
        B200ST_CUDA(launch_pdl(layernorm_bwd_reg_kernel<T, 4>, dim3((unsigned)ceil_div(rows, rpb)), dim3(256), smem,
                               (cudaStream_t)stream, (const T*)dy, (const T*)x, gamma, mean, rstd, (const T*)add, (T*)dx, dgamma, dbeta,
                               rows, (int)cols, rpb, partials));
      } else {
        B200ST_CUDA(launch_pdl(layernorm_bwd_reg_kernel<T, 8>, dim3((unsigned)ceil_div(rows, rpb)), dim3(256), smem,
                               (cudaStream_t)stream, (const T*)dy, (const T*)x, gamma, mean, rstd, (const T*)add, (T*)dx, dgamma, dbeta,
                               rows, (int)cols, rpb, partials));
      }
    });
    B200ST_LAUNCH_CHECK("layernorm_bwd_reg");
    return 0;
  }
  if (partials != nullptr) return set_error("layernorm_bwd_partial: needs cols %% 128 == 0, cols <= 1024 and 16-byte aligned operands");
  int rpb = (int)ceil_div(rows, 296);
  if (rpb < 4) rpb = 4;
  B200ST_DISPATCH(dtype, T, {
    B200ST_CUDA(launch_pdl(layernorm_bwd_kernel<T>, dim3((unsigned)ceil_div(rows, rpb)), dim3(128), smem,
                           (cudaStream_t)stream, (const T*)dy, (const T*)x, gamma, mean, rstd, (const T*)add, (T*)dx, dgamma, dbeta,
                           rows, (int)cols, rpb));
  });
  B200ST_LAUNCH_CHECK("layernorm_bwd");
  return 0;
}

int b200st_layernorm_bwd_add(int dtype, const void* dy, const void* x, const float* gamma,
                             const float* mean, const float* rstd, const void* add, void* dx, float* dgamma,
                             float* dbeta, int64_t rows, int64_t cols, b200st_stream_t stream) {
  return layernorm_bwd_impl(dtype, dy, x, gamma, mean, rstd, add, dx, dgamma, dbeta, rows, cols, nullptr, stream);
}

int64_t b200st_layernorm_bwd_partial_blocks(int64_t rows, int64_t cols) {
  return (cols % 128 == 0 && cols <= 1024) ? ceil_div(rows, 16) : 0;
}

int b200st_layernorm_bwd_partial(int dtype, const void* dy, const void* x, const float* gamma,
                                 const float* mean, const float* rstd, const void* add, void* dx, float* partials,
                                 int64_t rows, int64_t cols, b200st_stream_t stream) {
  if (partials == nullptr) return set_error("layernorm_bwd_partial: partials buffer required");
  return layernorm_bwd_impl(dtype, dy, x, gamma, mean, rstd, add, dx, partials, partials, rows, cols, partials, stream);
}

int b200st_layernorm_bwd(int dtype, const void* dy, const void* x, const float* gamma,
                         const float* mean, const float* rstd, void* dx, float* dgamma, float* dbeta,
                         int64_t rows, int64_t cols, b200st_stream_t stream) {
  return b200st_layernorm_bwd_add(dtype, dy, x, gamma, mean, rstd, nullptr, dx, dgamma, dbeta, rows, cols, stream);
}

int b200st_log_softmax_fwd(int dtype, const void* x, void* y, int64_t rows, int64_t cols,
                           int64_t* argmax, b200st_stream_t stream) {
  if (rows <= 0) return 0;
  const size_t smem = cols * sizeof(float);
  B200ST_DISPATCH(dtype, T, {
    if (set_row_smem((const void*)log_softmax_fwd_kernel<T>, smem, "log_softmax_fwd")) return -1;
    log_softmax_fwd_kernel<T><<<(unsigned)rows, 256, smem, (cudaStream_t)stream>>>(
        (const T*)x, (T*)y, (int)cols, argmax);
  });
  B200ST_LAUNCH_CHECK("log_softmax_fwd");
  return 0;
}

int b200st_log_softmax_bwd(int dtype, const void* dy, const void* y, void* dx, int64_t rows,
                           int64_t cols, b200st_stream_t stream) {
  if (rows <= 0) return 0;
  const size_t smem = cols * sizeof(float);
  B200ST_DISPATCH(dtype, T, {
    if (set_row_smem((const void*)log_softmax_bwd_kernel<T>, smem, "log_softmax_bwd")) return -1;
    log_softmax_bwd_kernel<T><<<(unsigned)rows, 256, smem, (cudaStream_t)stream>>>(
        (const T*)dy, (const T*)y, (T*)dx, (int)cols);
  });
  B200ST_LAUNCH_CHECK("log_softmax_bwd");
  return 0;
}

int b200st_masked_nll_fwd(int dtype, const void* logp, int64_t ld, const int64_t* target,
                          const uint8_t* mask, float* loss_sum, int64_t rows, int64_t cols,
                          b200st_stream_t stream) {
  if (rows <= 0) return 0;
  unsigned grid = (unsigned)ceil_div(rows, 256);
  if (grid > 296) grid = 296;
  B200ST_DISPATCH(dtype, T, {
    masked_nll_fwd_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)logp, ld, target, mask,
                                                                     loss_sum, rows, cols);
  });
  B200ST_LAUNCH_CHECK("masked_nll_fwd");
  return 0;
}

int b200st_masked_nll_bwd(int dtype, const float* gscale, const int64_t* target, const uint8_t* mask,
                          void* dlogp, int64_t ld, int64_t rows, int64_t cols, b200st_stream_t stream) {
  if (rows <= 0) return 0;
  B200ST_DISPATCH(dtype, T, {
    masked_nll_bwd_kernel<T><<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(gscale, target, mask,
                                                                               (T*)dlogp, ld, cols);
  });
  B200ST_LAUNCH_CHECK("masked_nll_bwd");
  return 0;
}

int b200st_softmax_nll_fused(int dtype, const void* logits, int64_t ld, const int64_t* target,
                             const uint8_t* mask, const float* scale, float eps, float* loss_sum,
                             void* dlogits, int64_t ld_d, int64_t rows, int64_t cols,
                             b200st_stream_t stream) {
  if (rows <= 0) return 0;
  if (dtype == B200ST_BF16 && cols % 8 == 0 && ld % 8 == 0 && ld_d % 8 == 0 && cols <= 512 * 8 * 3 &&
      (((uintptr_t)logits | (uintptr_t)dlogits) & 15) == 0) {
    const __nv_bfloat16* xp = (const __nv_bfloat16*)logits;
    __nv_bfloat16* dp = (__nv_bfloat16*)dlogits;
    const int nvec = (int)(cols / 8);
    if (nvec <= 512)
      softmax_nll_fused_vec_kernel<1><<<(unsigned)rows, 512, 0, (cudaStream_t)stream>>>(xp, ld, target, mask, scale, eps, loss_sum, dp, ld_d, (int)cols);
    else if (nvec <= 1024)
      softmax_nll_fused_vec_kernel<2><<<(unsigned)rows, 512, 0, (cudaStream_t)stream>>>(xp, ld, target, mask, scale, eps, loss_sum, dp, ld_d, (int)cols);
    else
      softmax_nll_fused_vec_kernel<3><<<(unsigned)rows, 512, 0, (cudaStream_t)stream>>>(xp, ld, target, mask, scale, eps, loss_sum, dp, ld_d, (int)cols);
    B200ST_LAUNCH_CHECK("softmax_nll_fused_vec");
    return 0;
  }
  const size_t smem = cols * sizeof(float);
  B200ST_DISPATCH(dtype, T, {
    if (set_row_smem((const void*)softmax_nll_fused_kernel<T>, smem, "softmax_nll_fused")) return -1;
    softmax_nll_fused_kernel<T><<<(unsigned)rows, 256, smem, (cudaStream_t)stream>>>(
        (const T*)logits, ld, target, mask, scale, eps, loss_sum, (T*)dlogits, ld_d, (int)cols);
  });
  B200ST_LAUNCH_CHECK("softmax_nll_fused");
  return 0;
}

}  // extern "C"
