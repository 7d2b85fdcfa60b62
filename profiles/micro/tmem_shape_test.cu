// What does tcgen05.ld / st .16x64b deliver to each thread?  (B200)  Writes lane * 256 + column into 32 columns of tensor memory
// with the 32x32b shape (thread = lane), reads them back with 16x64b.x8 (16 lanes x 16 columns) and prints the mapping.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I../../wildlifemapper_b200/csrc -o tmem_shape_test tmem_shape_test.cu
#include <cstdio>
#include "common.cuh"
using namespace wm;

__global__ void __launch_bounds__(128) k(uint32_t* out) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc(&slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot;
  uint32_t v[32];
  for (int i = 0; i < 32; ++i) v[i] = (uint32_t)((warp * 32 + lane) * 256 + i);
  tmem_st32(base + ((uint32_t)(warp * 32) << 16), v);
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // read 16 lanes x 16 columns starting at lane 32 * warp + 16 * half, column 0
  for (int half = 0; half < 2; ++half) {
    uint32_t r[8];
    const uint32_t addr = base + ((uint32_t)(warp * 32 + half * 16) << 16);
    asm volatile("tcgen05.ld.sync.aligned.16x64b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(addr));
    tmem_ld_wait();
    for (int i = 0; i < 8; ++i) out[((warp * 2 + half) * 32 + lane) * 8 + i] = r[i];
  }
  // store test: 16x64b.x4 writes value 0x10000 + lane * 16 + i to columns 32.. ; read back with 32x32b
  {
    uint32_t w[4];
    for (int i = 0; i < 4; ++i) w[i] = 0x10000u + (uint32_t)lane * 16u + (uint32_t)i;
    const uint32_t addr = base + ((uint32_t)(warp * 32) << 16) + 32;
    asm volatile("tcgen05.st.sync.aligned.16x64b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]));
    tmem_st_wait();
    uint32_t rb[8];
    tmem_ld8(base + ((uint32_t)(warp * 32) << 16) + 32, rb);
    tmem_ld_wait();
    for (int i = 0; i < 8; ++i) out[8 * 32 * 8 + (warp * 32 + lane) * 8 + i] = rb[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(base, 64); }
}

int main() {
  uint32_t* d; cudaMalloc(&d, 2 * 8 * 32 * 8 * 4); cudaMemset(d, 0, 2 * 8 * 32 * 8 * 4);
  k<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  static uint32_t h[2 * 8 * 32 * 8];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  for (int half = 0; half < 2; ++half) {
    printf("warp 1, lanes +%d: thread -> (lane, col) of its 8 registers\n", half * 16);
    for (int t = 0; t < 32; ++t) {
      printf(" t%02d:", t);
      for (int i = 0; i < 8; ++i) { uint32_t x = h[((1 * 2 + half) * 32 + t) * 8 + i]; printf(" (%u,%u)", x >> 8, x & 255); }
      printf("\n");
    }
  }
  printf("store test, warp 1: TMEM lane -> 8 columns (value = 0x10000 + srclane * 16 + reg)\n");
  for (int t = 0; t < 32; ++t) {
    printf(" lane%02d:", t);
    for (int i = 0; i < 8; ++i) { uint32_t x = h[8 * 32 * 8 + (32 + t) * 8 + i]; if (x >= 0x10000u) printf(" (t%u,r%u)", (x - 0x10000u) >> 4, x & 15); else printf(" [%u]", x); }
    printf("\n");
  }
  return 0;
}
