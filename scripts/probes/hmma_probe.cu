// Probe: how long does the recurrent GEMM of ONE LSTM time step take on the legacy warp-level tensor path (mma.sync, HMMA)
// with the recurrent weights resident in REGISTERS?  Per CTA: D[128 gate rows, 16 seqs] = W[128, 256] . h^T[256, 16],
// the B operand (h, 16 seqs x 256 k bf16 = 8 KB) read from shared memory with ldmatrix every step, a dependent write of
// the result back into the h tile + one bar.sync closing the step (the serial dependency of the recurrence).
//   config A: 8 warps = 4 m-tile pairs x 2 n-tiles   (A: 128 regs/thread, 32 HMMA/warp, B traffic 32 KB/step)
//   config B: 8 warps = 8 m-tiles, both n-tiles      (A:  64 regs/thread, 32 HMMA/warp, B traffic 64 KB/step)
//   config C: 4 warps = 4 m-tile pairs, both n-tiles (A: 128 regs/thread, 64 HMMA/warp, B traffic 32 KB/step)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma_probe hmma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void hmma(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}

// h tile: [16 seqs][256 k] bf16, row = 512 B, 16-byte chunks XOR-swizzled by (seq & 7) so that ldmatrix rows are conflict-free
__device__ __forceinline__ uint32_t h_off(uint32_t seq, uint32_t k) { return seq * 512 + ((((k >> 3) ^ (seq & 7)) & 31) << 4) + (k & 7) * 2; }

template <int WARPS, int MT, int NT, int KSPLIT>
__global__ void __launch_bounds__(WARPS * 32, 1) probe(long long* out, int iters, const uint32_t* __restrict__ wsrc) {
  __shared__ __align__(1024) uint8_t hbuf[2][8192];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 2 * 8192 / 4; i += WARPS * 32) ((uint32_t*)hbuf)[i] = 0x3c003c00u;
  uint32_t a[MT][16][4];
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int k = 0; k < 16; ++k)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        a[m][k][j] = wsrc[(((warp * MT + m) * 16 + k) * 4 + j) * 32 + lane];   // opaque: must stay in registers
      }
  // which n-tiles does this warp own?
  const int nt0 = (NT == 2) ? 0 : (warp / (WARPS / 2));
  __syncthreads();
  const uint32_t hb = smem_u32(hbuf);
  // ldmatrix.x4 address: matrices = (k 0-7, k 8-15) of k-step ks and of k-step ks+1 for one n-tile (8 seqs):
  // lanes 0-7 -> rows (seqs) of matrix 0, 8-15 matrix 1, ...
  const uint32_t lrow = lane & 7, lmat = lane >> 3;
  long long t0 = clock64();
  int cur = 0;
  for (int it = 0; it < iters; ++it) {
    float acc[MT][NT][KSPLIT][4];
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
      for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int s = 0; s < KSPLIT; ++s)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[m][n][s][j] = 0.f;
    const uint32_t base = hb + cur * 8192;
#pragma unroll
    for (int kp = 0; kp < 8; ++kp) {       // two k-steps per ldmatrix.x4
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        const uint32_t seq = (nt0 + n) * 8 + lrow;
        const uint32_t k = kp * 32 + lmat * 8;
        uint32_t b0, b1, b2, b3;
        ldsm4(base + h_off(seq, k), b0, b1, b2, b3);
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          hmma(acc[m][n][(2 * kp) % KSPLIT], a[m][2 * kp], b0, b1);
          hmma(acc[m][n][(2 * kp + 1) % KSPLIT], a[m][2 * kp + 1], b2, b3);
        }
      }
    }
    // dependent write-back: something derived from the accumulators goes into the other h buffer
    const uint32_t obase = hb + (cur ^ 1) * 8192;
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int s = 0; s < KSPLIT; ++s) { s0 += acc[m][n][s][0] + acc[m][n][s][2]; s1 += acc[m][n][s][1] + acc[m][n][s][3]; }
        __nv_bfloat162 v = __floats2bfloat162_rn(s0 * 1e-3f, s1 * 1e-3f);
        const uint32_t seq = (nt0 + n) * 8 + (lane & 3) * 2, k = ((warp * MT + m) * 8 + (lane >> 2)) * 2;
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(obase + h_off(seq, k & 255)), "r"(*reinterpret_cast<uint32_t*>(&v)) : "memory");
      }
    __syncthreads();
    cur ^= 1;
  }
  long long t1 = clock64();
  if (tid == 0) { out[0] = (t1 - t0) / iters; out[1] = ((uint32_t*)hbuf)[lane]; }
}

template <int WARPS, int MT, int NT, int KSPLIT>
void run(const char* name, long long* d) {
  probe<WARPS, MT, NT, KSPLIT><<<1, WARPS * 32>>>(d, 2000, (const uint32_t*)(d + 8));
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, (const void*)probe<WARPS, MT, NT, KSPLIT>);
  printf("%-52s %5lld cycles/step  (regs %d, local %zu B) %s\n", name, h[0], fa.numRegs, fa.localSizeBytes, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 64 + 16 * 2 * 16 * 4 * 32 * 4);
  cudaMemset(d, 0x3c, 64 + 16 * 2 * 16 * 4 * 32 * 4);
  run<8, 2, 1, 1>("A  8 warps, 2 m-tiles x 1 n-tile, 1 chain", d);
  run<8, 2, 1, 2>("A  8 warps, 2 m-tiles x 1 n-tile, k split 2", d);
  run<8, 2, 1, 4>("A  8 warps, 2 m-tiles x 1 n-tile, k split 4", d);
  run<8, 1, 2, 1>("B  8 warps, 1 m-tile x 2 n-tiles, 1 chain", d);
  run<8, 1, 2, 2>("B  8 warps, 1 m-tile x 2 n-tiles, k split 2", d);
  run<8, 1, 2, 4>("B  8 warps, 1 m-tile x 2 n-tiles, k split 4", d);
  run<4, 2, 2, 1>("C  4 warps, 2 m-tiles x 2 n-tiles, 1 chain", d);
  run<4, 2, 2, 2>("C  4 warps, 2 m-tiles x 2 n-tiles, k split 2", d);
  run<16, 1, 1, 1>("D 16 warps, 1 m-tile x 1 n-tile, 1 chain", d);
  run<16, 1, 1, 2>("D 16 warps, 1 m-tile x 1 n-tile, k split 2", d);
  return 0;
}
