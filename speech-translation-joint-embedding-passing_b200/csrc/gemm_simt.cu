// General strided-batched GEMM on the CUDA cores with fp32 accumulation.
//
// This is the exact-arithmetic path: fp32 parity mode (1e-4 contract) cannot use bf16/tf32 tensor-core
// inputs, and it also serves every shape the tcgen05 kernel (gemm_tc.cu) does not cover (K not a
// multiple of 64, tiny M, strided operands).  C = relu?(alpha * op(A) op(B) + bias) + R.
//
// Tiling: BM x BN block tile, BK-deep smem stages, TM x TN register micro-tile per thread; both
// operands are staged k-major in shared memory so the inner product reads float4 rows.
#include "common.cuh"

namespace b200st {

template <typename TI, typename TO, bool TA, bool TB, int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
gemm_simt_kernel(int64_t M, int64_t N, int64_t K, float alpha,
                 const TI* __restrict__ A, int64_t lda, int64_t sa,
                 const TI* __restrict__ B, int64_t ldb, int64_t sb,
                 TO* C, int64_t ldc, int64_t sc,
                 const TO* R, int64_t ldr, int64_t sr,
                 const float* __restrict__ bias, int relu) {
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int PAD = 4;
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];

  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  const int64_t bz = blockIdx.z;
  A += bz * sa;
  B += bz * sb;
  C += bz * sc;
  if (R) R += bz * sr;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = 0; k0 < K; k0 += BK) {
    // ---- stage A tile (BM x BK) as As[k][m]
    for (int i = tid; i < BM * BK; i += NT) {
      int m, k;
      if (TA) { m = i % BM; k = i / BM; }      // stored [K,M]: m contiguous
      else    { k = i % BK; m = i / BK; }      // stored [M,K]: k contiguous
      const int64_t gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < K) v = to_f(TA ? A[gk * lda + gm] : A[gm * lda + gk]);
      As[k][m] = v;
    }
    // ---- stage B tile (BK x BN) as Bs[k][n]
    for (int i = tid; i < BN * BK; i += NT) {
      int n, k;
      if (TB) { k = i % BK; n = i / BK; }      // stored [N,K]: k contiguous
      else    { n = i % BN; k = i / BN; }      // stored [K,N]: n contiguous
      const int64_t gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < N && gk < K) v = to_f(TB ? B[gn * ldb + gk] : B[gk * ldb + gn]);
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        const float4 v = *reinterpret_cast<const float4*>(&As[k][ty * TM + i]);
        a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
      }
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        const float4 v = *reinterpret_cast<const float4*>(&Bs[k][tx * TN + j]);
        b[j] = v.x; b[j + 1] = v.y; b[j + 2] = v.z; b[j + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t gm = m0 + ty * TM + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int64_t gn = n0 + tx * TN + j;
      if (gn >= N) continue;
      float v = alpha * acc[i][j];
      if (bias) v += bias[gn];
      if (relu == 1) v = fmaxf(v, 0.f);
      if (R) {
        const float r = to_f(R[gm * ldr + gn]);
        v = relu == 2 ? (r > 0.f ? v : 0.f) : v + r;        // relu == 2: R gates the result (ReLU backward)
      }
      C[gm * ldc + gn] = from_f<TO>(v);
    }
  }
}

template <typename TI, typename TO, bool TA, bool TB>
static int launch_gemm(int64_t M, int64_t N, int64_t K, float alpha, const void* A, int64_t lda,
                       int64_t sa, const void* B, int64_t ldb, int64_t sb, void* C, int64_t ldc,
                       int64_t sc, const void* R, int64_t ldr, int64_t sr, const float* bias, int relu,
                       int64_t batch, cudaStream_t st) {
  // Big tiles only pay when there are enough of them to fill 148 SMs.
  const int64_t big_tiles = ceil_div(M, 128) * ceil_div(N, 128) * batch;
  if (big_tiles >= 120) {
    dim3 grid((unsigned)ceil_div(N, 128), (unsigned)ceil_div(M, 128), (unsigned)batch);
    gemm_simt_kernel<TI, TO, TA, TB, 128, 128, 8, 8, 8><<<grid, 256, 0, st>>>(
        M, N, K, alpha, (const TI*)A, lda, sa, (const TI*)B, ldb, sb, (TO*)C, ldc, sc, (const TO*)R,
        ldr, sr, bias, relu);
  } else if (ceil_div(M, 64) * ceil_div(N, 64) * batch >= 64) {
    dim3 grid((unsigned)ceil_div(N, 64), (unsigned)ceil_div(M, 64), (unsigned)batch);
    gemm_simt_kernel<TI, TO, TA, TB, 64, 64, 16, 4, 4><<<grid, 256, 0, st>>>(
        M, N, K, alpha, (const TI*)A, lda, sa, (const TI*)B, ldb, sb, (TO*)C, ldc, sc, (const TO*)R,
        ldr, sr, bias, relu);
  } else {
    dim3 grid((unsigned)ceil_div(N, 32), (unsigned)ceil_div(M, 32), (unsigned)batch);
    gemm_simt_kernel<TI, TO, TA, TB, 32, 32, 16, 4, 4><<<grid, 64, 0, st>>>(
        M, N, K, alpha, (const TI*)A, lda, sa, (const TI*)B, ldb, sb, (TO*)C, ldc, sc, (const TO*)R,
        ldr, sr, bias, relu);
  }
  B200ST_LAUNCH_CHECK("gemm_simt");
  return 0;
}

template <typename TI, typename TO>
static int dispatch_trans(int ta, int tb, int64_t M, int64_t N, int64_t K, float alpha, const void* A,
                          int64_t lda, int64_t sa, const void* B, int64_t ldb, int64_t sb, void* C,
                          int64_t ldc, int64_t sc, const void* R, int64_t ldr, int64_t sr,
                          const float* bias, int relu, int64_t batch, cudaStream_t st) {
#define GO(TA_, TB_) \
  return launch_gemm<TI, TO, TA_, TB_>(M, N, K, alpha, A, lda, sa, B, ldb, sb, C, ldc, sc, R, ldr, sr, bias, relu, batch, st)
  if (!ta && !tb) GO(false, false);
  if (!ta && tb) GO(false, true);
  if (ta && !tb) GO(true, false);
  GO(true, true);
#undef GO
}

int gemm_simt(int dtype_ab, int dtype_c, int ta, int tb, int64_t M, int64_t N, int64_t K, float alpha,
              const void* A, int64_t lda, int64_t sa, const void* B, int64_t ldb, int64_t sb, void* C,
              int64_t ldc, int64_t sc, const void* R, int64_t ldr, int64_t sr, const float* bias,
              int relu, int64_t batch, cudaStream_t st) {
  if (M <= 0 || N <= 0 || batch <= 0) return 0;
  if (batch > 65535) return set_error("gemm: batch %lld > 65535", (long long)batch);
  if (dtype_ab == B200ST_F32 && dtype_c == B200ST_F32)
    return dispatch_trans<float, float>(ta, tb, M, N, K, alpha, A, lda, sa, B, ldb, sb, C, ldc, sc, R, ldr, sr, bias, relu, batch, st);
  if (dtype_ab == B200ST_BF16 && dtype_c == B200ST_BF16)
    return dispatch_trans<__nv_bfloat16, __nv_bfloat16>(ta, tb, M, N, K, alpha, A, lda, sa, B, ldb, sb, C, ldc, sc, R, ldr, sr, bias, relu, batch, st);
  if (dtype_ab == B200ST_BF16 && dtype_c == B200ST_F32)
    return dispatch_trans<__nv_bfloat16, float>(ta, tb, M, N, K, alpha, A, lda, sa, B, ldb, sb, C, ldc, sc, R, ldr, sr, bias, relu, batch, st);
  return set_error("gemm: unsupported dtype combination ab=%d c=%d", dtype_ab, dtype_c);
}

}  // namespace b200st

namespace b200st {
bool gemm_tc_eligible(int dtype_ab, int ta, int tb, int64_t M, int64_t N, int64_t K, const void* A, int64_t lda,
                      const void* B, int64_t ldb, int64_t batch);
int gemm_tc(int dtype_c, int ta, int tb, int64_t M, int64_t N, int64_t K, float alpha, const void* A, int64_t lda,
            const void* B, int64_t ldb, void* C, int64_t ldc, const void* R, int64_t ldr, const float* bias, int relu,
            cudaStream_t st, const void* A2 = nullptr, int64_t lda2 = 0, const void* B2 = nullptr, int64_t ldb2 = 0,
            int64_t K2 = 0);
static int g_gemm_backend = 0;   // 0 auto, 1 CUDA cores only, 2 tensor cores required
}  // namespace b200st

namespace b200st { int gemm_tc_set_persistent(int on); int gemm_tc_set_pair(int on); int gemm_tc_set_pair_dbg(int v); int gemm_tc_set_sm_budget(int n); }

extern "C" int b200st_set_gemm_sm_budget(int n) { return b200st::gemm_tc_set_sm_budget(n); }
// bit 0: persistent kernels, bit 1: CTA-pair (cta_group::2) kernel for the largest shapes; returns the previous mask
extern "C" int b200st_set_gemm_persistent(int on) {
  const int old = b200st::gemm_tc_set_persistent(-1) | (b200st::gemm_tc_set_pair(-1) << 1);
  if (on >= 0 && on <= 15) { b200st::gemm_tc_set_persistent(on & 1); b200st::gemm_tc_set_pair((on >> 1) & 1); b200st::gemm_tc_set_pair_dbg((on >> 2) & 3); }
  return old;
}

extern "C" int b200st_set_gemm_backend(int mode) {
  const int old = b200st::g_gemm_backend;
  if (mode >= 0 && mode <= 2) b200st::g_gemm_backend = mode;
  return old;
}

extern "C" int b200st_gemm(int dtype_ab, int dtype_c, int trans_a, int trans_b, int64_t M, int64_t N,
                           int64_t K, float alpha, const void* A, int64_t lda, int64_t stride_a,
                           const void* B, int64_t ldb, int64_t stride_b, void* C, int64_t ldc,
                           int64_t stride_c, const void* R, int64_t ldr, int64_t stride_r,
                           const float* bias, int relu, int64_t batch, b200st_stream_t stream) {
  using namespace b200st;
  if (M <= 0 || N <= 0 || batch <= 0) return 0;
  const bool ok = g_gemm_backend != 1 &&
                  gemm_tc_eligible(dtype_ab, trans_a, trans_b, M, N, K, A, lda, B, ldb, batch);
  // bf16 operands go to the tcgen05 kernel whenever TMA can address them; tiny problems (less work than one
  // tile row of MMAs) and everything fp32 stay on the exact CUDA-core kernel.
  if (ok && (g_gemm_backend == 2 || M * N * K >= (1ll << 16)))
    return gemm_tc(dtype_c, trans_a, trans_b, M, N, K, alpha, A, lda, B, ldb, C, ldc, R, ldr, bias, relu,
                   (cudaStream_t)stream);
  if (g_gemm_backend == 2)
    return set_error("gemm: tensor-core backend forced but operands are not TMA-addressable "
                     "(need bf16, batch 1, 16-byte aligned bases and leading dims %% 8 == 0)");
  return gemm_simt(dtype_ab, dtype_c, trans_a, trans_b, M, N, K, alpha, A, lda, stride_a, B,
                   ldb, stride_b, C, ldc, stride_c, R, ldr, stride_r, bias, relu, batch,
                   (cudaStream_t)stream);
}

// C = relu?(alpha * (op(A) op(B) + op(A2) op(B2)) + bias) + R: one launch with a two-segment K loop on the tensor-core
// path (first segment K %% 64 == 0), otherwise two GEMMs chained through the residual slot.
extern "C" int b200st_gemm2(int dtype_ab, int dtype_c, int trans_a, int trans_b, int64_t M, int64_t N, int64_t K,
                            int64_t K2, float alpha, const void* A, int64_t lda, const void* B, int64_t ldb,
                            const void* A2, int64_t lda2, const void* B2, int64_t ldb2, void* C, int64_t ldc,
                            const void* R, int64_t ldr, const float* bias, int relu, b200st_stream_t stream) {
  using namespace b200st;
  if (M <= 0 || N <= 0) return 0;
  const bool ok = g_gemm_backend != 1 && K % 64 == 0 && relu != 2 &&
                  gemm_tc_eligible(dtype_ab, trans_a, trans_b, M, N, K, A, lda, B, ldb, 1) &&
                  gemm_tc_eligible(dtype_ab, trans_a, trans_b, M, N, K2, A2, lda2, B2, ldb2, 1);
  if (ok)
    return gemm_tc(dtype_c, trans_a, trans_b, M, N, K, alpha, A, lda, B, ldb, C, ldc, R, ldr, bias, relu,
                   (cudaStream_t)stream, A2, lda2, B2, ldb2, K2);
  if (relu) return set_error("gemm2: the chained fallback cannot apply an activation to the sum");
  if (b200st_gemm(dtype_ab, dtype_c, trans_a, trans_b, M, N, K, alpha, A, lda, 0, B, ldb, 0, C, ldc, 0, R, ldr, 0,
                  bias, 0, 1, stream))
    return -1;
  return b200st_gemm(dtype_ab, dtype_c, trans_a, trans_b, M, N, K2, alpha, A2, lda2, 0, B2, ldb2, 0, C, ldc, 0, C, ldc,
                     0, nullptr, 0, 1, stream);
}
