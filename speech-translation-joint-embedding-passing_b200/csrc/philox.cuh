// Counter-based dropout masks (Philox4x32-10, Salmon et al. SC'11 constants).
// A mask element is a pure function of (seed, step, site, element index):
//   rng[0] = seed, rng[1] = step  -- a 2 x int64 DEVICE array; `step` is advanced once per forward pass by
//   b200st_rng_advance (a captured launch, so every replay of a CUDA graph draws fresh masks),
//   site = which dropout call inside the step (the host numbers them), index = position in the tensor.
// Backward recomputes the mask from the same four numbers: no mask tensor is stored.
#pragma once
#include <stdint.h>

namespace b200st {

struct Philox4 { uint32_t v[4]; };

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                 uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  Philox4 o;
  o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
  return o;
}

struct DropRng {
  uint32_t k0, k1, c2, c3;
  uint32_t thresh;     // keep iff random >= thresh;  thresh = p * 2^32
  float scale;         // 1 / (1 - p)
  __device__ __forceinline__ void init(const int64_t* __restrict__ rng, int64_t site, float p) {
    const uint64_t seed = (uint64_t)rng[0], step = (uint64_t)rng[1];
    k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32) ^ (uint32_t)(step >> 32);
    c2 = (uint32_t)site; c3 = (uint32_t)step;
    const double t = (double)p * 4294967296.0;
    thresh = t >= 4294967295.0 ? 0xffffffffu : (uint32_t)t;
    scale = 1.f / (1.f - p);
  }
  // the four random words of element group g (elements 4g .. 4g+3)
  __device__ __forceinline__ Philox4 group(uint64_t g) const {
    return philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), c2, c3, k0, k1);
  }
  // multiplier (0 or 1/(1-p)) of ONE element
  __device__ __forceinline__ float factor(uint64_t idx) const {
    const Philox4 r = group(idx >> 2);
    return r.v[idx & 3] >= thresh ? scale : 0.f;
  }
};

}  // namespace b200st
