// Probe: where do the 64 accumulator rows of a cta_group::1 M=64 tcgen05.mma land in TMEM?
// A[i][k] = i + 1 for k == 0 else 0 (64 x 64, K-major SW128), B = [j][k]: B[j][0] = j + 1 -> D[i][j] = (i+1)*(j+1).
// Dump all 128 lanes x 64 columns and print, per lane, the decoded row (D[lane][0] - 1).
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
__device__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ uint32_t swz(uint32_t row, uint32_t k) { return row * 128 + (((k >> 3) ^ (row & 7)) << 4); }
__device__ uint64_t desc(uint32_t a) { uint64_t d = (a >> 4) & 0x3FFF; d |= 1ull << 16; d |= (uint64_t)(1024 >> 4) << 32; d |= 1ull << 46; d |= 2ull << 61; return d; }
__global__ void probe(float* out, int M) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* A = sm; uint8_t* B = sm + 16384;
  __shared__ uint64_t bar; __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) ((uint32_t*)sm)[i] = 0;
  __syncthreads();
  if (threadIdx.x < 128) {
    int i = threadIdx.x;
    if (i < M) *(__nv_bfloat16*)(A + swz(i, 0)) = __float2bfloat16((float)(i + 1));
    if (i < 64) *(__nv_bfloat16*)(B + swz(i, 0)) = __float2bfloat16((float)(i + 1));
  }
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&slot)), "r"(64u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t tb = slot;
  // zero the accumulator region first so untouched lanes read 0
  {
    uint32_t z[16] = {0};
    int w = threadIdx.x >> 5;
    for (int c = 0; c < 64; c += 16)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                   ::"r"(tb + ((uint32_t)(w * 32) << 16) + c), "r"(z[0]),"r"(z[1]),"r"(z[2]),"r"(z[3]),"r"(z[4]),"r"(z[5]),"r"(z[6]),"r"(z[7]),"r"(z[8]),"r"(z[9]),"r"(z[10]),"r"(z[11]),"r"(z[12]),"r"(z[13]),"r"(z[14]),"r"(z[15]));
    asm volatile("tcgen05.wait::st.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (threadIdx.x == 0) {
    uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    for (int k = 0; k < 4; ++k) {
      uint64_t da = desc(s32(A) + k * 32), db = desc(s32(B) + k * 32);
      uint32_t acc = 1;
      asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }"
                   ::"r"(tb), "l"(da), "l"(db), "r"(idesc), "r"(acc));
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)));
  }
  uint32_t done = 0;
  while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p; }" : "=r"(done) : "r"(s32(&bar)));
  asm volatile("tcgen05.fence::after_thread_sync;");
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  for (int c = 0; c < 64; c += 16) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15])
                 : "r"(tb + ((uint32_t)(w * 32) << 16) + c));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    for (int j = 0; j < 16; ++j) out[(w * 32 + l) * 64 + c + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(64u));
}
int main() {
  for (int M : {128, 64}) {
    float* d; cudaMalloc(&d, 128 * 64 * 4); cudaMemset(d, 0, 128 * 64 * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
    probe<<<1, 128, 40000>>>(d, M);
    cudaError_t e = cudaDeviceSynchronize();
    static float h[128 * 64]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("M=%d (%s): lane -> row (col0 value - 1), and col1/col0 ratio\n", M, cudaGetErrorString(e));
    for (int lane = 0; lane < 128; ++lane) {
      float v0 = h[lane * 64], v1 = h[lane * 64 + 1], v63 = h[lane * 64 + 63];
      printf("%d:%g(%g,%g) ", lane, v0 - 1, v0 != 0 ? v1 / v0 : 0.f, v0 != 0 ? v63 / v0 : 0.f);
      if (lane % 8 == 7) printf("\n");
    }
  }
  return 0;
}
