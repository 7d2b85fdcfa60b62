// Internal declarations shared by the .cu translation units behind the C ABI (include/wm_b200.h).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#define WM_OK 0
#define WM_ERR_SHAPE (-1)
#define WM_ERR_ALIGN (-2)
#define WM_ERR_ARCH (-3)
#define WM_ERR_CUDA (-4)

namespace wm {

// cudaFuncSetAttribute is per DEVICE: one bit per device ordinal records where the dynamic shared-memory limit of a
// kernel has been raised (a process may drive several GPUs; the C ABI launches on the caller's current device).
template <typename Kernel>
inline int ensure_smem_attr(Kernel kernel, int bytes, std::atomic<unsigned long long>& done) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return WM_ERR_CUDA;
  const unsigned long long bit = 1ull << (dev & 63);
  if (done.load(std::memory_order_acquire) & bit) return WM_OK;
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) return WM_ERR_CUDA;
  done.fetch_or(bit, std::memory_order_release);
  return WM_OK;
}

struct GemmParams {
  int M, N, K;
  const float* bias;      // [N] fp32 or null
  const float* residual;  // fp32, row (m % res_mod), leading dim ldr; or null
  int ldr, res_mod;
  __nv_bfloat16* out_bf16;  // nullable
  int ldc_bf16;
  float* out_f32;  // nullable
  int ldc_f32;
  int act;     // 0 none, 1 erf-GELU, 2 ReLU, 3 sigmoid
  int vec_ok;  // 16-byte vector accesses to bias / residual allowed (alignment verified on host)
  int tma_out; // 1: the single output is written with TMA stores from swizzled smem staging (tc16 / tc32); 2: both outputs (CTA-pair kernel)
  int res_inplace;  // residual == out_f32 (same rows, same leading dim): fp32 output is a TMA reduce-add
  int a_mode;  // 0: A is [M,K] row-major; 1: implicit 3x3 conv over NHWC [B,64,64,conv_C]
  int conv_C;
};

// tc16 / tc32: output tensor maps (bf16 box 64 x 32, fp32 box 32 x 32, SWIZZLE_128B); only read when p.tma_out
int gemm_dispatch(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& tc16, const CUtensorMap& tc32,
                  const GemmParams& p, int bn, int num_sms, cudaStream_t st);

struct FlashParams {
  // O[b, t, h*HD + d] = softmax_k( scale * q.k + bias ) v   over Tk keys, per (image b, head h)
  int B, H, Tq, Tk;
  float scale;  // applied to q.k (the rel-pos bias uses the unscaled q)
  int q_col0, k_col0, v_col0;  // first column of head 0 inside the q / k / v tensor maps
  __nv_bfloat16* out;
  int ldo;
  int use_relpos;  // global 64x64 grid decomposed rel-pos; tables via tmap_rel ([256, HD]: rows 0..126 Rh, 128..254 Rw)
};
int flash4_dispatch(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& trel,
                    const FlashParams& p, int hd, cudaStream_t st);
int flash7_dispatch(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& trel,
                    const FlashParams& p, int hd, cudaStream_t st);  // head dim 64 (attn_flash7.cu)
int flash4_read_trace(unsigned long long* host_out);
int flash7_read_trace(unsigned long long* host_out);
int window2_read_trace(unsigned long long* host_out);  // diagnostics build (-DWM_F3_TRACE) only

struct WindowParams {
  int B, H;     // images, heads
  float scale;
  int D;        // embedding dim (q at col h*64, k at D + h*64, v at 2D + h*64 in the [B,64,64,3D] qkv tensor)
  __nv_bfloat16* out;  // [B,64,64,D]
};
int window2_dispatch(const CUtensorMap& tq, const CUtensorMap& tkv, const CUtensorMap& trel, const CUtensorMap& tout,
                     const WindowParams& p, int num_sms, cudaStream_t st);
int window3_dispatch(const CUtensorMap& tq, const CUtensorMap& tkv, const CUtensorMap& trel, const CUtensorMap& tout,
                     const WindowParams& p, int num_sms, cudaStream_t st);

}  // namespace wm
