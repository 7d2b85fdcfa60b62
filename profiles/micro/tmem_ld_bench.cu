// Microbenchmark (B200): tensor-memory read / write bandwidth seen by tcgen05.ld / tcgen05.st (32 lanes x 32 columns x 4 B
// = 4 KB per warp instruction) with 1, 2, 4 (one per scheduler) and 8 warps issuing back to back.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../wildlifemapper_b200/csrc -o tmem_ld_bench tmem_ld_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "common.cuh"

using namespace wm;

template <bool STORE>
__global__ void __launch_bounds__(256, 1) bench(long long* out, uint32_t* sink, int nwarps, int iters) {
  __shared__ uint32_t slot;
  __shared__ long long t_start[8], t_end[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(warp >> 2) * 256;
  uint32_t acc = 0;
  uint32_t v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = lane + i;
  __syncthreads();
  if (warp < nwarps) {
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (STORE) {
          tmem_st32(base + (u & 7) * 32, v);
        } else {
          tmem_ld32(base + (u & 7) * 32, v);
        }
      }
      if (STORE) tmem_st_wait();
      else {
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += v[i];
      }
    }
    const long long t1 = clock64();
    if (lane == 0) { t_start[warp] = t0; t_end[warp] = t1; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long a = t_start[0], b = t_end[0];
    for (int w = 1; w < nwarps; ++w) { a = min(a, t_start[w]); b = max(b, t_end[w]); }
    out[0] = b - a;
  }
  if (acc == 0x12345678u) sink[threadIdx.x] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

int main() {
  long long* d; uint32_t* sink;
  cudaMalloc(&d, 8); cudaMalloc(&sink, 4096);
  const int iters = 200;
  for (int st = 0; st < 2; ++st)
    for (int nw : {1, 2, 4, 8}) {
      if (st) bench<true><<<1, 256>>>(d, sink, nw, iters); else bench<false><<<1, 256>>>(d, sink, nw, iters);
      long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      const double bytes = (double)nw * iters * 8 * 4096;
      printf("tcgen05.%s x32  %d warps: %7.1f cycles per warp-instruction, %6.1f B/cycle/SM  %s\n", st ? "st" : "ld", nw,
             (double)h / (iters * 8), bytes / h, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
