// Microbenchmark (B200): how fast can 1 or 2 warps per SM sub-partition run the softmax inner loop?
//   variants: 0 MUFU.EX2 only | 1 scalar mix (FADD, MUFU, FADD, FADD, F2FP per pair -- the flash-v3 exp phase)
//             2 packed mix (add.f32x2 for the shift and the row sum) | 3 polynomial exp2 on the FMA pipe only
//             4 mix with 25 % of the exponentials on the FMA pipe | 5 FMNMX max phase (FFMA + FMNMX per score)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench mufu_bench.cu ; run: ./mufu_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float a, float b) { __nv_bfloat162 v = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&v); }
// exp2 for x <= 0 on the FMA/ALU pipes: 2^x = 2^floor(x) * p(frac), cubic minimax (FA4-style), rel err ~1e-4
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.0f);
  const float fl = floorf(x);   // FRND is a conversion-pipe op on some parts; use the magic-number trick instead
  (void)fl;
  const float t = x + 12582912.0f;              // 1.5 * 2^23: round to nearest integer in the mantissa
  const float n = t - 12582912.0f;
  const float f = x - n;                        // in [-0.5, 0.5]
  float p = fmaf(f, 0.0555041f, 0.2402265f);
  p = fmaf(p, f, 0.6931472f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

template <int VARIANT>
__global__ void __launch_bounds__(256, 1) bench(float* out, long long* cycles, int iters, float d) {
  float y[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) y[i] = -0.01f * (float)((threadIdx.x * 131 + i * 7) & 255);
  float acc[4] = {0, 0, 0, 0};
  uint32_t pk = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (VARIANT == 0) {
#pragma unroll
      for (int i = 0; i < 128; ++i) acc[i & 3] += ex2(y[i]);
    } else if (VARIANT == 1) {
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        const float e0 = ex2(y[2 * i] + d), e1 = ex2(y[2 * i + 1] + d);
        acc[i & 3] += e0 + e1;
        pk ^= pack(e0, e1);
      }
    } else if (VARIANT == 2) {
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        float2 s;
        asm("{.reg .b64 a, b, c; mov.b64 a, {%2, %3}; mov.b64 b, {%4, %4}; add.f32x2 c, a, b; mov.b64 {%0, %1}, c;}"
            : "=f"(s.x), "=f"(s.y) : "f"(y[2 * i]), "f"(y[2 * i + 1]), "f"(d));
        const float e0 = ex2(s.x), e1 = ex2(s.y);
        asm("{.reg .b64 a, b, c; mov.b64 a, {%0, %1}; mov.b64 b, {%2, %3}; add.f32x2 c, a, b; mov.b64 {%0, %1}, c;}"
            : "+f"(acc[0]), "+f"(acc[1]) : "f"(e0), "f"(e1));
        pk ^= pack(e0, e1);
      }
    } else if (VARIANT == 3) {
#pragma unroll
      for (int i = 0; i < 128; ++i) acc[i & 3] += ex2_poly(y[i] + d);
    } else if (VARIANT == 4) {
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        const float e0 = ex2(y[2 * i] + d);
        const float e1 = (i & 1) ? ex2(y[2 * i + 1] + d) : ex2_poly(y[2 * i + 1] + d);
        acc[i & 3] += e0 + e1;
        pk ^= pack(e0, e1);
      }
    } else if (VARIANT >= 6 && VARIANT <= 9) {
      // the flash-v4 rel-pos chunk: y = v*c1 + tw (FFMA2), [max (FMNMX)], y + d (FADD2), 2 x MUFU, pack, row sum (FADD2)
      //   6: as in the kernel   7: without the FMNMX   8: scalar FFMA / FADD instead of the packed forms   9: 7 + order-pinned
      float tw[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) tw[i] = 0.001f * (float)((i * 37 + threadIdx.x) & 63);
      float ymax0 = -1e30f, ymax1 = -1e30f;
      unsigned long long cs0 = 0ull, cs1 = 0ull;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int k = c * 32 + 2 * i;
          float y0, y1, a0, a1;
          if (VARIANT == 8) {
            y0 = fmaf(y[k], 0.18f, tw[k & 63]); y1 = fmaf(y[k + 1], 0.18f, tw[(k + 1) & 63]);
            a0 = y0 + d; a1 = y1 + d;
          } else {
            unsigned long long vp, twp, c1p, dp, y2, yd;
            asm("mov.b64 %0, {%1, %2};" : "=l"(vp) : "f"(y[k]), "f"(y[k + 1]));
            asm("mov.b64 %0, {%1, %2};" : "=l"(twp) : "f"(tw[k & 63]), "f"(tw[(k + 1) & 63]));
            asm("mov.b64 %0, {%1, %1};" : "=l"(c1p) : "f"(0.18f));
            asm("mov.b64 %0, {%1, %1};" : "=l"(dp) : "f"(d));
            asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(y2) : "l"(vp), "l"(c1p), "l"(twp));
            asm("mov.b64 {%0, %1}, %2;" : "=f"(y0), "=f"(y1) : "l"(y2));
            asm("add.rn.f32x2 %0, %1, %2;" : "=l"(yd) : "l"(y2), "l"(dp));
            asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(yd));
          }
          if (VARIANT == 6 || VARIANT == 8) { ymax0 = fmaxf(ymax0, y0); ymax1 = fmaxf(ymax1, y1); }
          float e0, e1;
          if (VARIANT == 9) {
            asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a0));
            asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(a1));
          } else {
            e0 = ex2(a0); e1 = ex2(a1);
          }
          unsigned long long ep;
          asm("mov.b64 %0, {%1, %2};" : "=l"(ep) : "f"(e0), "f"(e1));
          if (i & 1) asm("add.rn.f32x2 %0, %0, %1;" : "+l"(cs1) : "l"(ep));
          else asm("add.rn.f32x2 %0, %0, %1;" : "+l"(cs0) : "l"(ep));
          pk ^= pack(e0, e1);
        }
      }
      float s0, s1, s2, s3;
      asm("mov.b64 {%0, %1}, %2;" : "=f"(s0), "=f"(s1) : "l"(cs0));
      asm("mov.b64 {%0, %1}, %2;" : "=f"(s2), "=f"(s3) : "l"(cs1));
      acc[0] += s0 + s1 + s2 + s3 + ymax0 * 1e-30f + ymax1 * 1e-30f;
    } else if (VARIANT == 10 || VARIANT == 11) {
      // packed half-precision exponentials: y + d (FADD2) -> cvt.rn.f16x2.f32 -> ex2.approx.f16x2 (ONE MUFU per pair,
      // result already packed for the P operand) -> row sum in f16x2 (HADD2), one conversion per 32 scores
      //   10: f16x2    11: bf16x2
      unsigned h0 = 0u, h1 = 0u;
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        unsigned long long vp, dp, yd;
        float a0, a1;
        asm("mov.b64 %0, {%1, %2};" : "=l"(vp) : "f"(y[2 * i]), "f"(y[2 * i + 1]));
        asm("mov.b64 %0, {%1, %1};" : "=l"(dp) : "f"(d));
        asm("add.rn.f32x2 %0, %1, %2;" : "=l"(yd) : "l"(vp), "l"(dp));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(yd));
        unsigned hp, ep;
        if (VARIANT == 10) {
          asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(hp) : "f"(a1), "f"(a0));
          asm("ex2.approx.f16x2 %0, %1;" : "=r"(ep) : "r"(hp));
          if (i & 1) asm("add.rn.f16x2 %0, %0, %1;" : "+r"(h1) : "r"(ep));
          else asm("add.rn.f16x2 %0, %0, %1;" : "+r"(h0) : "r"(ep));
        } else {
          asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hp) : "f"(a1), "f"(a0));
          asm("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(ep) : "r"(hp));
          if (i & 1) asm("add.rn.bf16x2 %0, %0, %1;" : "+r"(h1) : "r"(ep));
          else asm("add.rn.bf16x2 %0, %0, %1;" : "+r"(h0) : "r"(ep));
        }
        pk ^= ep;
      }
      acc[0] += __uint_as_float(h0 << 16) + __uint_as_float(h1 << 16);
    } else if (VARIANT == 5) {
#pragma unroll
      for (int i = 0; i < 128; ++i) {
        y[i] = fmaf(y[i], 1.0001f, d);
        acc[i & 3] = fmaxf(acc[i & 3], y[i]);
      }
    }
    d += acc[0] * 1e-30f;
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc[0] + acc[1] + acc[2] + acc[3] + __uint_as_float(pk) + y[5];
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int V>
void run(const char* name, int warps_per_smsp) {
  const int threads = 128 * warps_per_smsp, iters = 200;
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&cyc, 148 * 8);
  bench<V><<<148, threads>>>(out, cyc, iters, -0.5f);
  cudaDeviceSynchronize();
  bench<V><<<148, threads>>>(out, cyc, iters, -0.5f);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  printf("%-34s %d warp/SMSP: %8.1f cycles per 128-score tile per warp-slot (%.2f cycles per score per SMSP)  %s\n", name, warps_per_smsp,
         avg / iters, avg / iters / 128.0 / warps_per_smsp * 1.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w = 1; w <= 2; ++w) {
    run<0>("MUFU.EX2 only", w);
    run<1>("scalar mix (v3 exp phase)", w);
    run<2>("packed f32x2 mix", w);
    run<3>("polynomial exp2 (FMA pipe)", w);
    run<4>("mix, 25% polynomial", w);
    run<5>("max phase (FFMA + FMNMX)", w);
    run<6>("v4 rel-pos chunk (packed, FMNMX)", w);
    run<7>("v4 rel-pos chunk without FMNMX", w);
    run<8>("v4 rel-pos chunk, scalar FFMA/FADD", w);
    run<9>("v4 rel-pos chunk, no FMNMX, pinned", w);
    run<10>("packed ex2.f16x2 path", w);
    run<11>("packed ex2.bf16x2 path", w);
  }
  return 0;
}
