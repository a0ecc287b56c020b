// Attention kernels (exact-fp32 CUDA-core versions).
//   * multi-head scaled-dot-product attention core, forward and backward (layers.py:213-229)
//   * LAS single-head bilinear attention step, forward and backward (attention.py:190-193,250-273)
//   * row arg-max and the LAS decode-length rule (Dec.py:320-341)
// Sequence lengths on the training path are tiny (<= 50 keys, 126 acoustic frames), so one warp owns a
// query row end to end: scores, mask, softmax and the P.V product never leave the SM.
#include "common.cuh"

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

__global__ void las_update_lengths_kernel(const int64_t* __restrict__ sym, int64_t sym_stride,
                                          int32_t* __restrict__ lengths, int step, int64_t B) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int64_t s = sym[b * sym_stride];
  if ((s == 3 /*EOS*/ || s == 0 /*PAD*/) && lengths[b] > step) lengths[b] = step + 1;  // Dec.py:334-340
}

}  // namespace b200st

using namespace b200st;

extern "C" {

int b200st_mha_fwd(int dtype, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                   int64_t ldv, const uint8_t* mask, int64_t mask_sb, int64_t mask_sq, void* o,
                   int64_t ldo, void* p, int64_t B, int64_t H, int64_t Lq, int64_t Lk, int64_t d,
                   float temperature, b200st_stream_t stream) {
  if (B <= 0 || Lq <= 0) return 0;
  if (Lk <= 0) return set_error("mha_fwd: empty key sequence");
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

int b200st_mha_bwd(int dtype, const void* dout, int64_t ldo, const void* q, int64_t ldq,
                   const void* k, int64_t ldk, const void* v, int64_t ldv, const void* p, void* ds,
                   void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, int64_t B,
                   int64_t H, int64_t Lq, int64_t Lk, int64_t d, float temperature,
                   b200st_stream_t stream) {
  if (B <= 0 || Lq <= 0 || Lk <= 0) return 0;
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

int b200st_las_attn_fwd(int dtype, const void* q, const void* wk, const void* vals,
                        const int32_t* klens, void* ctx, float* probs, int64_t B, int64_t Tk, int64_t D,
                        int64_t Dv, b200st_stream_t stream) {
  if (B <= 0) return 0;
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

int b200st_las_update_lengths(const int64_t* sym, int64_t sym_stride, int32_t* lengths, int step,
                              int64_t B, b200st_stream_t stream) {
  if (B <= 0) return 0;
  las_update_lengths_kernel<<<(unsigned)ceil_div(B, 128), 128, 0, (cudaStream_t)stream>>>(
      sym, sym_stride, lengths, step, B);
  B200ST_LAUNCH_CHECK("las_update_lengths");
  return 0;
}

}  // extern "C"
