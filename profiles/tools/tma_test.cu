#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__global__ void k(const float* x, int rows, int row_floats, int pitch, float* out, int variant) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* st = (float*)smem_raw;
  __shared__ __align__(8) uint64_t full;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t bytes = row_floats * 4;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full)), "r"(bytes * rows) : "memory");
    for (int r = 0; r < rows; ++r)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(st + r * pitch)), "l"(x + r * row_floats), "r"(bytes), "r"(smem_u32(&full)) : "memory");
  }
  long long t0 = clock64();
  bool done = false;
  while (!(done = try_wait(&full, 0))) { if (clock64() - t0 > 2000000000LL) break; }
  if (threadIdx.x == 0) {
    unsigned long long w = *(volatile unsigned long long*)&full;
    printf("variant %d rows %d bytes %u pitch %d done %d barrier word %016llx\n", variant, rows, bytes, pitch, (int)done, w);
  }
  if (done && threadIdx.x < rows) out[threadIdx.x] = st[threadIdx.x * pitch] + st[threadIdx.x * pitch + row_floats - 1];
}
int main() {
  float *x, *out;
  cudaMalloc(&x, 1 << 20); cudaMalloc(&out, 4096);
  float h[4096]; for (int i = 0; i < 4096; ++i) h[i] = i;
  cudaMemcpy(x, h, sizeof(h), cudaMemcpyHostToDevice);
  int cfg[][3] = {{1, 24, 24}, {5, 24, 24}, {5, 24, 40}, {5, 32, 32}, {1, 4, 4}, {4, 416, 432}};
  for (int i = 0; i < 6; ++i) {
    k<<<1, 256, 32 * 1024>>>(x, cfg[i][0], cfg[i][1], cfg[i][2], out, i);
    cudaError_t e0 = cudaGetLastError(); cudaError_t e = cudaDeviceSynchronize(); if (e0) printf("launch err %s\n", cudaGetErrorString(e0));
    float o[8] = {0}; cudaMemcpy(o, out, 32, cudaMemcpyDeviceToHost);
    printf("  -> %s out0 %g out1 %g\n", cudaGetErrorString(e), o[0], o[1]);
  }
  return 0;
}
