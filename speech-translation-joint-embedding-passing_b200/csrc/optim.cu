// Fused gradient-norm clip + Adam over ALL parameter tensors of the model.
//
// Replaces torch.nn.utils.clip_grad_norm_ + torch.optim.Adam.step() as called by Optimizer.step()
// (reference: modules/optim.py:31-36, trainer/trainer_base.py:422-426) -- ~270 parameter tensors, 69.5 M fp32
// elements at the benchmark config.  HBM-bound: the gradient is read once for the norm (278 MB) and once more,
// with p, m, v, for the update (4 reads + 3 writes per element = 28 B/element, 1.95 GB per step).
//
// Multi-tensor layout: a device table of int64 [n_tensors][7] = {p, g, m, v, numel, bf16 copy, bf16 copy} and a block map
// int32 [n_blocks][2] = {tensor, chunk} built once by the host (b200st/optim.py); one CTA handles one chunk of
// OPT_CHUNK elements of one tensor, so a single launch covers every tensor regardless of size.
// Three launches per step, no host synchronisation, CUDA-graph capturable:
//   multi_sqnorm  : per-chunk sum of squares -> partials[n_blocks]            (deterministic: no atomics)
//   adam_prepare  : ONE CTA sums the partials in a fixed order, forms the clip coefficient, advances the device
//                   step counter and writes {clip, lr / (1 - b1^t), sqrt(1 - b2^t), ||g||}
//   multi_adam    : m, v, p update (torch.optim.Adam arithmetic, amsgrad off) with the clip coefficient applied to
//                   the gradient on the fly; optionally rewrites the bf16 shadow copy of the parameter that the
//                   tensor-core GEMMs read, so no separate cast pass is needed after the step.
#include "common.cuh"

namespace b200st {

constexpr int OPT_CHUNK = 8192;      // elements per CTA
constexpr int OPT_THREADS = 256;
constexpr int OPT_COLS = 7;          // table row: {p, g, m, v, numel, bf16 copy 1 or 0, bf16 copy 2 or 0}     // 32 elements per thread = 8 x float4

__global__ void __launch_bounds__(OPT_THREADS)
multi_sqnorm_kernel(const int64_t* __restrict__ table, const int* __restrict__ blockmap,
                    float* __restrict__ partials) {
  __shared__ float scratch[32];
  const int t = blockmap[2 * blockIdx.x], chunk = blockmap[2 * blockIdx.x + 1];
  const float* g = reinterpret_cast<const float*>(table[OPT_COLS * t + 1]);
  const int64_t n = table[OPT_COLS * t + 4];
  const int64_t base = (int64_t)chunk * OPT_CHUNK;
  const int64_t end = min(base + OPT_CHUNK, n);
  float acc = 0.f;
  if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    for (int64_t i = base + threadIdx.x * 4; i < end; i += OPT_THREADS * 4) {
      if (i + 4 <= end) {
        const float4 v = *reinterpret_cast<const float4*>(g + i);
        acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      } else {
        for (int64_t j = i; j < end; ++j) acc += g[j] * g[j];
      }
    }
  } else {
    for (int64_t i = base + threadIdx.x; i < end; i += OPT_THREADS) acc += g[i] * g[i];
  }
  acc = block_sum(acc, scratch);
  if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

// scal[0] = clip coefficient, scal[1] = lr / (1 - beta1^t), scal[2] = sqrt(1 - beta2^t), scal[3] = ||g||_2
__global__ void __launch_bounds__(1024)
adam_prepare_kernel(const float* __restrict__ partials, int64_t n_blocks, float max_grad_norm,
                    const float* __restrict__ lr, double beta1, double beta2, float* __restrict__ step,
                    float* __restrict__ scal) {
  __shared__ double sh[1024];
  double acc = 0.0;
  if (partials != nullptr)
    for (int64_t i = threadIdx.x; i < n_blocks; i += blockDim.x) acc += (double)partials[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float norm = (float)sqrt(sh[0]);
    float clip = 1.f;
    if (max_grad_norm > 0.f && partials != nullptr) {          // torch.nn.utils.clip_grad_norm_: clamp(max/(norm+1e-6), 1)
      clip = max_grad_norm / (norm + 1e-6f);
      if (clip > 1.f) clip = 1.f;
    }
    const float t = *step + 1.f;                               // fp32 scalar like torch's state['step']
    *step = t;
    const double bc1 = 1.0 - pow(beta1, (double)t);
    const double bc2 = 1.0 - pow(beta2, (double)t);
    scal[0] = clip;
    scal[1] = (float)((double)lr[0] / bc1);
    scal[2] = (float)sqrt(bc2);
    scal[3] = norm;
  }
}

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, float clip, float step_size,
                                          float sqrt_bc2, float omb1, float b2, float omb2, float eps, float wd) {
  g *= clip;
  if (wd != 0.f) g = fmaf(wd, p, g);                            // Adam (not AdamW): L2 term joins the gradient
  m = fmaf(omb1, g - m, m);                                      // exp_avg.lerp_(grad, 1 - beta1)
  v = fmaf(b2, v, omb2 * g * g);                                // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  const float denom = sqrtf(v) / sqrt_bc2 + eps;               // (sqrt(v) / sqrt(bc2)).add_(eps)
  p -= step_size * (m / denom);                                 // param.addcdiv_(exp_avg, denom, value=-lr/bc1)
}

__global__ void __launch_bounds__(OPT_THREADS)
multi_adam_kernel(const int64_t* __restrict__ table, const int* __restrict__ blockmap,
                  const float* __restrict__ scal, float omb1, float b2, float omb2, float eps, float wd) {
  const int t = blockmap[2 * blockIdx.x], chunk = blockmap[2 * blockIdx.x + 1];
  float* p = reinterpret_cast<float*>(table[OPT_COLS * t + 0]);
  const float* g = reinterpret_cast<const float*>(table[OPT_COLS * t + 1]);
  float* m = reinterpret_cast<float*>(table[OPT_COLS * t + 2]);
  float* v = reinterpret_cast<float*>(table[OPT_COLS * t + 3]);
  const int64_t n = table[OPT_COLS * t + 4];
  __nv_bfloat16* sh = reinterpret_cast<__nv_bfloat16*>(table[OPT_COLS * t + 5]);
  __nv_bfloat16* sh2 = reinterpret_cast<__nv_bfloat16*>(table[OPT_COLS * t + 6]);
  const float clip = scal[0], step_size = scal[1], isb2 = scal[2];
  const int64_t base = (int64_t)chunk * OPT_CHUNK;
  const int64_t end = min(base + OPT_CHUNK, n);
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0 && ((reinterpret_cast<uintptr_t>(sh) | reinterpret_cast<uintptr_t>(sh2)) & 7) == 0;
  if (vec) {
    for (int64_t i = base + threadIdx.x * 4; i < end; i += OPT_THREADS * 4) {
      if (i + 4 <= end) {
        float pp[4], gg[4], mm[4], vv[4];
        load4(p + i, pp); load4(g + i, gg); load4(m + i, mm); load4(v + i, vv);
#pragma unroll
        for (int j = 0; j < 4; ++j) adam_elem(pp[j], gg[j], mm[j], vv[j], clip, step_size, isb2, omb1, b2, omb2, eps, wd);
        store4(p + i, pp); store4(m + i, mm); store4(v + i, vv);
        if (sh) store4(sh + i, pp);
        if (sh2) store4(sh2 + i, pp);
      } else {
        for (int64_t j = i; j < end; ++j) {
          float pp = p[j], mm = m[j], vv = v[j];
          adam_elem(pp, g[j], mm, vv, clip, step_size, isb2, omb1, b2, omb2, eps, wd);
          p[j] = pp; m[j] = mm; v[j] = vv;
          if (sh) sh[j] = __float2bfloat16_rn(pp);
          if (sh2) sh2[j] = __float2bfloat16_rn(pp);
        }
      }
    }
  } else {
    for (int64_t j = base + threadIdx.x; j < end; j += OPT_THREADS) {
      float pp = p[j], mm = m[j], vv = v[j];
      adam_elem(pp, g[j], mm, vv, clip, step_size, isb2, omb1, b2, omb2, eps, wd);
      p[j] = pp; m[j] = mm; v[j] = vv;
      if (sh) sh[j] = __float2bfloat16_rn(pp);
      if (sh2) sh2[j] = __float2bfloat16_rn(pp);
    }
  }
}

}  // namespace b200st

using namespace b200st;

extern "C" {

int b200st_opt_chunk(void) { return OPT_CHUNK; }
int b200st_opt_table_cols(void) { return OPT_COLS; }

int b200st_multi_sqnorm(const int64_t* table, const int32_t* blockmap, int64_t n_blocks, float* partials,
                        b200st_stream_t stream) {
  if (n_blocks <= 0) return 0;
  multi_sqnorm_kernel<<<(unsigned)n_blocks, OPT_THREADS, 0, (cudaStream_t)stream>>>(table, blockmap, partials);
  B200ST_LAUNCH_CHECK("multi_sqnorm");
  return 0;
}

int b200st_adam_prepare(const float* partials, int64_t n_blocks, float max_grad_norm, const float* lr, double beta1,
                        double beta2, float* step, float* scal, b200st_stream_t stream) {
  adam_prepare_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(partials, n_blocks, max_grad_norm, lr, beta1, beta2, step,
                                                            scal);
  B200ST_LAUNCH_CHECK("adam_prepare");
  return 0;
}

int b200st_multi_adam(const int64_t* table, const int32_t* blockmap, int64_t n_blocks, const float* scal, double beta1,
                      double beta2, double eps, double weight_decay, b200st_stream_t stream) {
  if (n_blocks <= 0) return 0;
  multi_adam_kernel<<<(unsigned)n_blocks, OPT_THREADS, 0, (cudaStream_t)stream>>>(
      table, blockmap, scal, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)eps, (float)weight_decay);
  B200ST_LAUNCH_CHECK("multi_adam");
  return 0;
}

}  // extern "C"
