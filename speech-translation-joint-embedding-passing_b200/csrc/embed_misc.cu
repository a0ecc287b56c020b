// Embedding gather/scatter, the embedding-passing concat (mix prologue), masks and small glue kernels.
// All of these are pure data movement: vectorisable, coalesced along the feature dimension, grid sized
// from the element count.
#include "common.cuh"

namespace b200st {

template <typename T>
__global__ void embedding_fwd_kernel(const int64_t* __restrict__ ids, const float* __restrict__ table,
                                     T* __restrict__ out, int64_t ld_out, int64_t n, int dim,
                                     int64_t vocab) {
  pdl_wait();
  pdl_launch_dependents();
  const int64_t i = blockIdx.x;
  int64_t id = ids[i];
  if (id < 0 || id >= vocab) id = 0;   // out-of-range ids read the PAD row instead of faulting
  const float* src = table + id * dim;
  T* dst = out + i * ld_out;
  for (int c = threadIdx.x; c < dim; c += blockDim.x) dst[c] = from_f<T>(src[c]);
}

template <typename T>
__global__ void embedding_bwd_kernel(const int64_t* __restrict__ ids, const T* __restrict__ dout,
                                     int64_t ld_dout, float* __restrict__ dtable, int64_t n, int dim,
                                     int64_t vocab, int64_t padding_idx) {
  const int64_t i = blockIdx.x;
  const int64_t id = ids[i];
  if (id == padding_idx || id < 0 || id >= vocab) return;
  const T* src = dout + i * ld_dout;
  float* dst = dtable + id * dim;
  for (int c = threadIdx.x; c < dim; c += blockDim.x) atomicAdd(&dst[c], to_f(src[c]));
}

template <typename T>
__global__ void mix_gather_concat_kernel(const int64_t* __restrict__ ids,
                                         const float* __restrict__ table, const T* __restrict__ dyn,
                                         int64_t ld_dyn, T* __restrict__ cat, int E, int D,
                                         int64_t vocab) {
  const int64_t i = blockIdx.x;
  int64_t id = ids[i];
  if (id < 0 || id >= vocab) id = 0;
  const float* src = table + id * E;
  const T* dr = dyn + i * ld_dyn;
  T* dst = cat + i * (int64_t)(E + D);
  for (int c = threadIdx.x; c < E + D; c += blockDim.x)
    dst[c] = (c < E) ? from_f<T>(src[c]) : dr[c - E];
}

template <typename T>
__global__ void add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out,
                           int64_t n) {
  pdl_wait();
  pdl_launch_dependents();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = from_f<T>(to_f(a[i]) + to_f(b[i]));
}

template <typename T>
__global__ void add_posenc_kernel(const T* __restrict__ x, const float* __restrict__ pe,
                                  T* __restrict__ out, int64_t n, int64_t LD, int D) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = from_f<T>(to_f(x[i]) + pe[i % LD]);
}

constexpr int TR_RPB = 32;      // (a, b) rows per CTA: one CTA per row meant 64 512 tiny CTAs for the acoustic input (36 us of block scheduling)
template <typename TI, typename TO>
__global__ void transpose01_kernel(const TI* __restrict__ in, TO* __restrict__ out, int64_t A,
                                   int64_t Bd, int64_t C) {
  // rows (a, b) of C contiguous elements move to (b, a); a CTA owns TR_RPB consecutive source rows
  const int64_t r0 = (int64_t)blockIdx.x * TR_RPB, rows = A * Bd;
  const int64_t n = min((int64_t)TR_RPB, rows - r0) * C;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const int64_t r = r0 + i / C, c = i % C;
    const int64_t a = r / Bd, b = r % Bd;
    out[(b * A + a) * C + c] = from_f<TO>(to_f(in[r * C + c]));
  }
}

template <typename TI, typename TO>
__global__ void cast_kernel(const TI* __restrict__ in, TO* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = from_f<TO>(to_f(in[i]));
}

// Column sums: grid.x covers column tiles of 32, grid.y splits the rows; warp lanes run along columns.
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ x, int64_t ld, float* __restrict__ out,
                              int64_t rows, int64_t cols, int64_t rows_per_block) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float part[8][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t c = (int64_t)blockIdx.x * 32 + lane;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  int64_t r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  float s = 0.f;
  if (c < cols)
    for (int64_t r = r0 + w; r < r1; r += 8) s += to_f(x[r * ld + c]);
  part[w][lane] = s;
  __syncthreads();
  if (w == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][lane];
    atomicAdd(&out[c], t);
  }
}

// bf16, cols % 8 == 0: every thread owns 8 adjacent columns (one 16-byte load per row), a warp covers 256 columns
// of one row per instruction, the 8 warps of a CTA walk 8 rows at a time with 4 rows in flight per thread.
__global__ void __launch_bounds__(256)
colsum_vec_kernel(const __nv_bfloat16* __restrict__ x, int64_t ld, float* __restrict__ out, int64_t rows, int64_t cols,
                  int64_t rows_per_block) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float part[8][256 + 8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t c = (int64_t)blockIdx.x * 256 + lane * 8;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = min(rows, r0 + rows_per_block);
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c < cols) {
    for (int64_t r = r0 + w; r < r1; r += 32) {
      uint4 u[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        u[i] = (r + 8 * i < r1) ? __ldg(reinterpret_cast<const uint4*>(x + (r + 8 * i) * ld + c)) : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t w4[4] = {u[i].x, u[i].y, u[i].z, u[i].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          s[2 * e] += __uint_as_float(w4[e] << 16);
          s[2 * e + 1] += __uint_as_float(w4[e] & 0xffff0000u);
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) part[w][lane * 8 + e] = s[e];
  __syncthreads();
  const int64_t cc = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (cc < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][threadIdx.x];
    atomicAdd(&out[cc], t);
  }
}

template <typename T>
__global__ void relu_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y, T* __restrict__ dx,
                                int64_t n) {
  pdl_wait();
  pdl_launch_dependents();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    dx[i] = to_f(y[i]) > 0.f ? dy[i] : from_f<T>(0.f);
}

__global__ void token_mask_kernel(const int64_t* __restrict__ ids, uint8_t* __restrict__ mask, int64_t B,
                                  int64_t L, int64_t pad, int causal) {
  const int64_t Lq = causal ? L : 1;
  const int64_t n = B * Lq * L;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = i % L, q = (i / L) % Lq, b = i / (L * Lq);
    mask[i] = (ids[b * L + j] != pad) && (!causal || j <= q);
  }
}

__global__ void length_mask_kernel(const int32_t* __restrict__ lengths, uint8_t* __restrict__ mask,
                                   int64_t B, int64_t L) {
  const int64_t n = B * L;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    mask[i] = (i % L) < lengths[i / L];
}

static unsigned flat_grid(int64_t n, int block) {
  int64_t g = ceil_div(n, block);
  if (g > 148 * 16) g = 148 * 16;
  return (unsigned)(g < 1 ? 1 : g);
}

}  // namespace b200st

using namespace b200st;

extern "C" {

int b200st_embedding_fwd(int dtype, const int64_t* ids, const float* table, void* out, int64_t ld_out,
                         int64_t n, int64_t dim, int64_t vocab, b200st_stream_t stream) {
  if (n <= 0) return 0;
  B200ST_DISPATCH(dtype, T, {
    B200ST_CUDA(launch_pdl(embedding_fwd_kernel<T>, dim3((unsigned)n), dim3(128), 0, (cudaStream_t)stream, ids, table,
                           (T*)out, ld_out, n, (int)dim, vocab));
  });
  B200ST_LAUNCH_CHECK("embedding_fwd");
  return 0;
}

int b200st_embedding_bwd(int dtype, const int64_t* ids, const void* dout, int64_t ld_dout,
                         float* dtable, int64_t n, int64_t dim, int64_t vocab, int64_t padding_idx,
                         b200st_stream_t stream) {
  if (n <= 0) return 0;
  B200ST_DISPATCH(dtype, T, {
    embedding_bwd_kernel<T><<<(unsigned)n, 128, 0, (cudaStream_t)stream>>>(
        ids, (const T*)dout, ld_dout, dtable, n, (int)dim, vocab, padding_idx);
  });
  B200ST_LAUNCH_CHECK("embedding_bwd");
  return 0;
}

int b200st_mix_gather_concat(int dtype, const int64_t* ids, const float* table, const void* dyn,
                             int64_t ld_dyn, void* cat, int64_t n, int64_t E, int64_t D, int64_t vocab,
                             b200st_stream_t stream) {
  if (n <= 0) return 0;
  B200ST_DISPATCH(dtype, T, {
    mix_gather_concat_kernel<T><<<(unsigned)n, 256, 0, (cudaStream_t)stream>>>(
        ids, table, (const T*)dyn, ld_dyn, (T*)cat, (int)E, (int)D, vocab);
  });
  B200ST_LAUNCH_CHECK("mix_gather_concat");
  return 0;
}

int b200st_add(int dtype, const void* a, const void* b, void* out, int64_t n, b200st_stream_t stream) {
  if (n <= 0) return 0;
  B200ST_DISPATCH(dtype, T, {
    B200ST_CUDA(launch_pdl(add_kernel<T>, dim3(flat_grid(n, 256)), dim3(256), 0, (cudaStream_t)stream, (const T*)a,
                           (const T*)b, (T*)out, n));
  });
  B200ST_LAUNCH_CHECK("add");
  return 0;
}

int b200st_add_posenc(int dtype, const void* x, const float* pe, void* out, int64_t B, int64_t L,
                      int64_t D, b200st_stream_t stream) {
  const int64_t n = B * L * D;
  if (n <= 0) return 0;
  B200ST_DISPATCH(dtype, T, {
    add_posenc_kernel<T><<<flat_grid(n, 256), 256, 0, (cudaStream_t)stream>>>((const T*)x, pe, (T*)out,
                                                                              n, L * D, (int)D);
  });
  B200ST_LAUNCH_CHECK("add_posenc");
  return 0;
}

int b200st_transpose01(int dtype_in, int dtype_out, const void* in, void* out, int64_t A, int64_t Bd,
                       int64_t C, b200st_stream_t stream) {
  if (A * Bd * C <= 0) return 0;
  const unsigned grid = (unsigned)ceil_div(A * Bd, TR_RPB);
  const int block = 256;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype_in == B200ST_F32 && dtype_out == B200ST_F32)
    transpose01_kernel<float, float><<<grid, block, 0, st>>>((const float*)in, (float*)out, A, Bd, C);
  else if (dtype_in == B200ST_F32 && dtype_out == B200ST_BF16)
    transpose01_kernel<float, __nv_bfloat16><<<grid, block, 0, st>>>((const float*)in, (__nv_bfloat16*)out, A, Bd, C);
  else if (dtype_in == B200ST_BF16 && dtype_out == B200ST_BF16)
    transpose01_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, block, 0, st>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, A, Bd, C);
  else if (dtype_in == B200ST_BF16 && dtype_out == B200ST_F32)
    transpose01_kernel<__nv_bfloat16, float><<<grid, block, 0, st>>>((const __nv_bfloat16*)in, (float*)out, A, Bd, C);
  else
    return set_error("transpose01: bad dtypes");
  B200ST_LAUNCH_CHECK("transpose01");
  return 0;
}

int b200st_cast(int dtype_in, int dtype_out, const void* in, void* out, int64_t n,
                b200st_stream_t stream) {
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = flat_grid(n, 256);
  if (dtype_in == B200ST_F32 && dtype_out == B200ST_BF16)
    cast_kernel<float, __nv_bfloat16><<<grid, 256, 0, st>>>((const float*)in, (__nv_bfloat16*)out, n);
  else if (dtype_in == B200ST_BF16 && dtype_out == B200ST_F32)
    cast_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>((const __nv_bfloat16*)in, (float*)out, n);
  else if (dtype_in == B200ST_F32 && dtype_out == B200ST_F32)
    cast_kernel<float, float><<<grid, 256, 0, st>>>((const float*)in, (float*)out, n);
  else if (dtype_in == B200ST_BF16 && dtype_out == B200ST_BF16)
    cast_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, n);
  else
    return set_error("cast: bad dtypes");
  B200ST_LAUNCH_CHECK("cast");
  return 0;
}

int b200st_colsum(int dtype, const void* x, int64_t ld, float* out, int64_t rows, int64_t cols,
                  int accumulate, b200st_stream_t stream) {
  if (cols <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate) B200ST_CUDA(cudaMemsetAsync(out, 0, cols * sizeof(float), st));
  if (rows <= 0) return 0;
  if (dtype == B200ST_BF16 && cols % 8 == 0 && ld % 8 == 0 && ((uintptr_t)x & 15) == 0) {
    const int64_t ct = ceil_div(cols, 256);
    int64_t rb = ceil_div(148 * 4, ct);
    if (rb > ceil_div(rows, 32)) rb = ceil_div(rows, 32);
    if (rb < 1) rb = 1;
    const int64_t rpb = ceil_div(rows, rb);
    B200ST_CUDA(launch_pdl(colsum_vec_kernel, dim3((unsigned)ct, (unsigned)ceil_div(rows, rpb)), dim3(256), 0, st,
                           (const __nv_bfloat16*)x, ld, out, rows, cols, rpb));
    B200ST_LAUNCH_CHECK("colsum_vec");
    return 0;
  }
  const int64_t col_tiles = ceil_div(cols, 32);
  int64_t row_blocks = ceil_div(148 * 4, col_tiles);
  if (row_blocks > ceil_div(rows, 64)) row_blocks = ceil_div(rows, 64);
  if (row_blocks < 1) row_blocks = 1;
  const int64_t rpb = ceil_div(rows, row_blocks);
  dim3 grid((unsigned)col_tiles, (unsigned)ceil_div(rows, rpb));
  B200ST_DISPATCH(dtype, T, {
    B200ST_CUDA(launch_pdl(colsum_kernel<T>, grid, dim3(256), 0, st, (const T*)x, ld, out, rows, cols, rpb));
  });
  B200ST_LAUNCH_CHECK("colsum");
  return 0;
}

int b200st_relu_bwd(int dtype, const void* dy, const void* y, void* dx, int64_t n,
                    b200st_stream_t stream) {
  if (n <= 0) return 0;
  B200ST_DISPATCH(dtype, T, {
    B200ST_CUDA(launch_pdl(relu_bwd_kernel<T>, dim3(flat_grid(n, 256)), dim3(256), 0, (cudaStream_t)stream,
                           (const T*)dy, (const T*)y, (T*)dx, n));
  });
  B200ST_LAUNCH_CHECK("relu_bwd");
  return 0;
}

int b200st_token_mask(const int64_t* ids, uint8_t* mask, int64_t B, int64_t L, int64_t pad, int causal,
                      b200st_stream_t stream) {
  const int64_t n = B * (causal ? L : 1) * L;
  if (n <= 0) return 0;
  token_mask_kernel<<<flat_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(ids, mask, B, L, pad, causal);
  B200ST_LAUNCH_CHECK("token_mask");
  return 0;
}

int b200st_length_mask(const int32_t* lengths, uint8_t* mask, int64_t B, int64_t L,
                       b200st_stream_t stream) {
  if (B * L <= 0) return 0;
  length_mask_kernel<<<flat_grid(B * L, 256), 256, 0, (cudaStream_t)stream>>>(lengths, mask, B, L);
  B200ST_LAUNCH_CHECK("length_mask");
  return 0;
}

}  // extern "C"

// ---- dropout (nn.Dropout call sites: Enc.py:159-212, Dec.py:166,386-429, Seq2seq.py:195-209, layers.py:182-249) ----
#include "philox.cuh"

namespace b200st {

// y[r, c] = x[r, c] * keep(r * ld_m + c_off + c) / (1 - p) (+ res[r, c]);  x, y, res row-strided 2-D views.
template <typename T>
__global__ void __launch_bounds__(256)
dropout_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ res, int64_t ldr, T* __restrict__ y,
               int64_t ldy, int64_t rows, int64_t cols, int64_t ld_m, int64_t c_off, float p,
               const int64_t* __restrict__ rng, int64_t site, int vec) {
  pdl_wait();
  pdl_launch_dependents();
  DropRng d;
  d.init(rng, site, p);
  if (vec) {                         // dense, every extent a multiple of 4: one Philox call per 4 elements
    const int64_t n4 = rows * cols / 4;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n4; g += (int64_t)gridDim.x * blockDim.x) {
      const Philox4 r = d.group((uint64_t)g);
      float v[4], o[4];
      load4(x + 4 * g, v);
      if (res) load4(res + 4 * g, o);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = (r.v[j] >= d.thresh ? v[j] * d.scale : 0.f) + (res ? o[j] : 0.f);
      store4(y + 4 * g, v);
    }
    return;
  }
  const int64_t n = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols, c = i - r * cols;
    float v = to_f(x[r * ldx + c]) * d.factor((uint64_t)(r * ld_m + c_off + c));
    if (res) v += to_f(res[r * ldr + c]);
    y[r * ldy + c] = from_f<T>(v);
  }
}

__global__ void rng_advance_kernel(int64_t* rng) { rng[1] += 1; }

// debug aid: the device's nanosecond timer at the point of the stream where this one-thread kernel runs
__global__ void stamp_kernel(long long* slot) {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  *slot = t;
}

}  // namespace b200st

extern "C" {

int b200st_dropout(int dtype, const void* x, int64_t ldx, const void* residual, int64_t ldr, void* y, int64_t ldy,
                   int64_t rows, int64_t cols, int64_t ld_mask, int64_t col_off, float p, const int64_t* rng,
                   int64_t site, b200st_stream_t stream) {
  if (rows * cols <= 0) return 0;
  if (!(p >= 0.f && p < 1.f)) return set_error("dropout: p=%f outside [0, 1)", (double)p);
  const bool dense = ldx == cols && ldy == cols && ld_mask == cols && col_off == 0 && (!residual || ldr == cols);
  const int a = dtype == B200ST_F32 ? 15 : 7;
  const int vec = dense && (rows * cols) % 4 == 0 && (((uintptr_t)x | (uintptr_t)y | (uintptr_t)residual) & a) == 0;
  const int64_t work = vec ? rows * cols / 4 : rows * cols;
  B200ST_DISPATCH(dtype, T, {
    B200ST_CUDA(launch_pdl(dropout_kernel<T>, dim3(flat_grid(work, 256)), dim3(256), 0, (cudaStream_t)stream,
                           (const T*)x, ldx, (const T*)residual, ldr, (T*)y, ldy, rows, cols, ld_mask, col_off, p, rng,
                           site, vec));
  });
  B200ST_LAUNCH_CHECK("dropout");
  return 0;
}

int b200st_debug_stamp(int64_t* slot, b200st_stream_t stream) {
  stamp_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((long long*)slot);
  B200ST_LAUNCH_CHECK("debug_stamp");
  return 0;
}

int b200st_rng_advance(int64_t* rng, b200st_stream_t stream) {
  rng_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(rng);
  B200ST_LAUNCH_CHECK("rng_advance");
  return 0;
}

}  // extern "C"

// ---- input stage: batch assembly of Dataset.load_acous_from_flis (utils/dataset.py:155-184) on the device -----------
// The host ships the utterances of a batch back to back (no padding over PCIe); this kernel applies the per-speaker
// mean / std normalisation (dataset.py:169-173) and writes the zero-padded [B, T_pad, F] batch (dataset.py:178-182).
namespace b200st {

__global__ void __launch_bounds__(256)
fbank_norm_pad_kernel(const float* __restrict__ packed, const int64_t* __restrict__ offsets,
                      const int32_t* __restrict__ lens, const float* __restrict__ mu, const float* __restrict__ sd,
                      float* __restrict__ out, int64_t T_pad, int F) {
  const int b = blockIdx.y;
  const int64_t n = T_pad * F;
  const int len = lens[b];
  const float* src = packed + offsets[b] * F;
  float* dst = out + (int64_t)b * n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = i / F;
    const int c = (int)(i - t * F);
    float v = 0.f;
    if (t < len) {
      v = src[i];
      if (mu != nullptr) v = __fdiv_rn(v - mu[(int64_t)b * F + c], sd[(int64_t)b * F + c]);   // 1. * (x - mu) / std
    }
    dst[i] = v;
  }
}

}  // namespace b200st

extern "C" int b200st_fbank_norm_pad(const float* packed, const int64_t* offsets, const int32_t* lens, const float* mu,
                                     const float* sd, float* out, int64_t B, int64_t T_pad, int64_t F,
                                     b200st_stream_t stream) {
  if (B <= 0 || T_pad <= 0 || F <= 0) return 0;
  if ((mu == nullptr) != (sd == nullptr)) return set_error("fbank_norm_pad: mu and sd must both be given or both be NULL");
  const int64_t n = T_pad * F;
  const int64_t gx = ceil_div(n, 256 * 4);
  dim3 grid((unsigned)(gx < 1024 ? gx : 1024), (unsigned)B);
  fbank_norm_pad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(packed, offsets, lens, mu, sd, out, T_pad, (int)F);
  B200ST_LAUNCH_CHECK("fbank_norm_pad");
  return 0;
}
