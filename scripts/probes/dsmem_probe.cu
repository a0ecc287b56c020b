// Probe: cost of an all-to-all exchange inside an 8-CTA cluster (each CTA delivers CHUNK bytes to each of the 8 CTAs),
// (a) st.async.v4 per thread, (b) cp.async.bulk shared::cta -> shared::cluster, one bulk copy per destination.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dsmem_probe dsmem_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) { uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o; }
__device__ __forceinline__ void mbar_wait(uint32_t addr, uint32_t parity) {
  uint32_t done = 0;
  while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
}
template <int MODE, int CHUNK>
__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(128, 1) probe(long long* out, int iters) {
  __shared__ __align__(128) uint8_t src[8 * CHUNK];
  __shared__ __align__(128) uint8_t dst[2][8 * CHUNK];
  __shared__ __align__(8) uint64_t bar[2];
  uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int tid = threadIdx.x;
  for (int i = tid; i < 8 * CHUNK / 4; i += 128) ((uint32_t*)src)[i] = i + rank;
  if (tid == 0) {
    for (int b = 0; b < 2; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[b])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int b = 0; b < 2; ++b) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[b])), "r"(8 * CHUNK) : "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  uint32_t ph[2] = {0, 0};
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const int b = it & 1;
    if (MODE == 0) {
      // 8*CHUNK bytes out, 16 B per st.async: thread t handles pieces t, t+128, ...
      for (int p = tid; p < 8 * CHUNK / 16; p += 128) {
        const uint32_t r = (p / (CHUNK / 16) + rank) & 7, off = p % (CHUNK / 16);
        const uint4 v = *reinterpret_cast<const uint4*>(src + r * CHUNK + off * 16);
        const uint32_t d = mapa(smem_u32(&dst[b][rank * CHUNK + off * 16]), r), m = mapa(smem_u32(&bar[b]), r);
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                     ::"r"(d), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(m) : "memory");
      }
    } else {
      if (tid < 8) {
        const uint32_t r = (tid + rank) & 7;
        const uint32_t d = mapa(smem_u32(&dst[b][rank * CHUNK]), r), m = mapa(smem_u32(&bar[b]), r);
        asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(d), "r"(smem_u32(src + r * CHUNK)), "r"(CHUNK), "r"(m) : "memory");
      }
    }
    mbar_wait(smem_u32(&bar[b]), ph[b]);
    ph[b] ^= 1;
    // re-arm for the use after next: safe because every sender waits for its own barrier b^1 (which needs OUR send of
    // iteration it+1, issued after this point) before sending iteration it+2 into b
    if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[b])), "r"(8 * CHUNK) : "memory");
    __syncthreads();
  }
  long long t1 = clock64();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (tid == 0 && blockIdx.x == 0) out[0] = (t1 - t0) / iters;
  if (tid == 0 && blockIdx.x == 0) out[1] = ((uint32_t*)dst[0])[5];
}
template <int MODE, int CHUNK>
void run(const char* name) {
  long long* d; cudaMalloc(&d, 16); long long h[2] = {0, 0};
  probe<MODE, CHUNK><<<8 * 8, 128>>>(d, 2000);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%s chunk %d B (out %d B/CTA): %lld cycles/exchange (%s)\n", name, CHUNK, 8 * CHUNK, h[0], cudaGetErrorString(e));
  cudaFree(d);
}
int main() {
  run<0, 256>("st.async"); run<0, 512>("st.async"); run<0, 1024>("st.async"); 
  run<1, 256>("bulk    "); run<1, 512>("bulk    "); run<1, 1024>("bulk    "); 
  return 0;
}
