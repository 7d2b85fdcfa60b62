// Microbenchmark (B200): do small tcgen05.mma instructions (128 x N x 16, bf16) that accumulate into the SAME tensor-memory
// tile serialise on a round trip, and do independent accumulators pipeline?  One CTA, one issuing thread.
//   chains = number of accumulator tiles the instruction stream alternates between (1 = every MMA depends on the previous)
//   ts     = A operand from tensor memory (the P V form) instead of shared memory
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../wildlifemapper_b200/csrc -o umma_chain_bench umma_chain_bench.cu
// run:   ./umma_chain_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "common.cuh"

using namespace wm;

template <int N, int CHAINS, bool TS, bool COMMITS = false>
__global__ void __launch_bounds__(128, 1) bench(long long* out, int n_mma) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    const uint64_t ad = make_sdesc_sw128(smem_u32(smem), 16, 1024);
    const uint64_t bd = make_sdesc_sw128(smem_u32(smem + 16384), 16, 1024);
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
#pragma unroll 1
      for (int i = 0; i < n_mma; i += 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t d = tm + (uint32_t)(((i / 4) * 4 + k) % CHAINS) * (uint32_t)N;
          if (TS) umma_bf16_ts(d, tm + 448, bd + 2 * k, idesc, 1);
          else umma_bf16(d, ad + 2 * k, bd + 2 * k, idesc, 1);
        }
        if (COMMITS) umma_commit(&bar2);  // one commit per group of four MMAs, nobody waits on it
      }
      const long long t1 = clock64();
      umma_commit(&bar);
      mbar_wait(&bar, rep & 1);
      const long long t2 = clock64();
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

template <int N, int CHAINS, bool TS, bool COMMITS = false>
void run(long long* d_out, const char* name) {
  const int n = 256;
  cudaFuncSetAttribute(bench<N, CHAINS, TS, COMMITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
  bench<N, CHAINS, TS, COMMITS><<<1, 128, 60000>>>(d_out, n);
  long long h[2];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  printf("%-34s issue %6.1f cyc/MMA   complete %6.1f cyc/MMA   (math floor %d)%s\n", name, (double)h[0] / n, (double)h[1] / n,
         128 * N / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  run<64, 1, false>(d, "128x64x16  SS 1 chain");
  run<64, 2, false>(d, "128x64x16  SS 2 chains");
  run<64, 4, false>(d, "128x64x16  SS 4 chains");
  run<64, 1, true>(d, "128x64x16  TS 1 chain");
  run<64, 2, true>(d, "128x64x16  TS 2 chains");
  run<64, 4, true>(d, "128x64x16  TS 4 chains");
  run<128, 1, false>(d, "128x128x16 SS 1 chain");
  run<128, 2, false>(d, "128x128x16 SS 2 chains");
  run<256, 1, false>(d, "128x256x16 SS 1 chain");
  run<64, 1, false, true>(d, "128x64x16  SS 1 chain + commit/4");
  run<64, 1, true, true>(d, "128x64x16  TS 1 chain + commit/4");
  return 0;
}
