// C ABI (include/wm_b200.h): argument validation, TMA tensor-map construction, dispatch.  No torch types.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <string>

#include "../../include/wm_b200.h"
#include "wm_internal.h"

namespace wm {
int layernorm_launch(const float*, const float*, const float*, __nv_bfloat16*, float*, const float*, int,
                     __nv_bfloat16*, int, int, float, cudaStream_t);
int patchify_launch(const float*, __nv_bfloat16*, __nv_bfloat16*, int, int, int, cudaStream_t);
int transpose_split_launch(const float*, __nv_bfloat16*, int, int, int, cudaStream_t);
int transpose_launch(const void*, void*, int, int, int, int, cudaStream_t);
int hfc_finalize_launch(const float*, const float*, __nv_bfloat16*, float*, int, cudaStream_t);
int add_cast_launch(const float*, const float*, int, __nv_bfloat16*, int, int, cudaStream_t);
int attn_small_launch(const __nv_bfloat16*, int, const __nv_bfloat16*, int, const __nv_bfloat16*, int, __nv_bfloat16*,
                      int, int, int, int, int, int, float, cudaStream_t);
int postprocess_launch(const float*, const float*, const long long*, float, int, float*, int*, long long*, int*, int,
                       int, int, cudaStream_t);
int sigmoid_topk_launch(const float*, const float*, float*, int*, float*, int*, int*, float*, int, int, int, int, int,
                        int, cudaStream_t);
int nms_launch(const float*, const float*, const long long*, int, double, int*, unsigned long long*, long long*, int*,
               cudaStream_t);
int nms_batched_small_launch(const float*, const int*, int, int, float, double, int, int*, int*, cudaStream_t);
int tiles_from_u8_launch(const uint8_t*, int, int, long long, const int*, int, int, int, const float*, const float*, float*,
                         cudaStream_t);
int merge_detections_launch(const float*, const int*, const int*, int, int, float, int*, float*, float*, long long*, int*,
                            int*, cudaStream_t);
int coco_pack_launch(const float*, const float*, const long long*, const long long*, int, float*, long long*, cudaStream_t);
int match_cost_launch(const float*, const float*, const long long*, const float*, int, int, int, float, float, float, float*,
                      cudaStream_t);
int criterion_launch(const float*, const float*, const int*, const long long*, const float*, int, const int*, const float*, int,
                     int, int, float, int*, float*, cudaStream_t);
int resize_u8_launch(const uint8_t*, long long, const int*, int, int, int, uint8_t*, uint8_t*, int, int, const int*, const int*,
                     int, const int*, const int*, int, cudaStream_t);
}  // namespace wm

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

int check_launch(int rc, const char* what) {
  if (rc == WM_OK) return WM_OK;
  if (rc == WM_ERR_CUDA) return fail(rc, "%s: CUDA error: %s", what, cudaGetErrorString(cudaGetLastError()));
  return fail(rc, "%s: unsupported shape", what);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Per-device init-once state (a process may drive several GPUs: every entry point works on the caller's CURRENT device).
struct DeviceState {
  int checked = 0;  // 0 unknown, 1 ok, -1 bad
  int num_sms = 0;
  std::string why;
};
constexpr int kMaxDevices = 64;
DeviceState g_devs[kMaxDevices];
EncodeTiledFn g_encode = nullptr;  // driver entry point (process-wide), set under g_mu
std::mutex g_mu;
thread_local int t_num_sms = 0;    // SM count of the device ensure_device() last validated on this thread

int ensure_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return fail(WM_ERR_ARCH, "no CUDA device");
  std::lock_guard<std::mutex> lk(g_mu);
  DeviceState& d = g_devs[dev];
  if (d.checked == 1) {
    t_num_sms = d.num_sms;
    return WM_OK;
  }
  if (d.checked == -1) return fail(WM_ERR_ARCH, "%s", d.why.c_str());
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    d.why = "no CUDA device";
    d.checked = -1;
    return fail(WM_ERR_ARCH, "%s", d.why.c_str());
  }
  if (prop.major != 10) {
    char b[128];
    snprintf(b, sizeof(b), "wm_b200 requires sm_100 (B200); found sm_%d%d -- there is no fallback path", prop.major,
             prop.minor);
    d.why = b;
    d.checked = -1;
    return fail(WM_ERR_ARCH, "%s", d.why.c_str());
  }
  if (g_encode == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || fn == nullptr) {
      d.why = "cuTensorMapEncodeTiled not available from the driver";
      d.checked = -1;
      return fail(WM_ERR_ARCH, "%s", d.why.c_str());
    }
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  d.num_sms = prop.multiProcessorCount;
  d.checked = 1;
  t_num_sms = d.num_sms;
  return WM_OK;
}

// bf16 tensor map, SWIZZLE_128B, inner box = 64 elements (128 bytes).  dims/strides innermost first.
int make_map(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
             const uint32_t* box, const char* what, CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16) {
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return fail(WM_ERR_ALIGN, "%s: base pointer not 16-byte aligned", what);
  cuuint64_t gd[5];
  cuuint64_t gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) {
      if (strides_bytes[i - 1] % 16 != 0) return fail(WM_ERR_ALIGN, "%s: stride %d not a multiple of 16 bytes", what, i);
      gs[i - 1] = strides_bytes[i - 1];
    }
  }
  CUresult r = g_encode(m, dtype, (cuuint32_t)rank, const_cast<void*>(ptr), gd, gs, bx, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(WM_ERR_CUDA, "%s: cuTensorMapEncodeTiled failed (%d)", what, (int)r);
  return WM_OK;
}

int make_map_2d(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                const char* what) {
  const uint64_t dims[2] = {cols, rows};
  const uint64_t strides[1] = {ld * 2};
  const uint32_t box[2] = {64, box_rows};
  return make_map(m, ptr, 2, dims, strides, box, what);
}

// A/B measurement knobs: atomics (set from any thread; every call reads each knob once)
std::atomic<int> g_flash_version{7};
std::atomic<int> g_gemm_pairs{1};

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

extern "C" {

int wm_version(void) { return 100; }
const char* wm_last_error(void) { return g_err.c_str(); }
int wm_device_check(void) { return ensure_device(); }
int wm_set_flash_version(int version) {
  if (version != 4 && version != 7) return fail(WM_ERR_SHAPE, "wm_set_flash_version: 4 or 7");
  g_flash_version.store(version);
  return WM_OK;
}
int wm_set_option(const char* name, int value) {
  const std::string n(name ? name : "");
  if (n == "flash_version") return wm_set_flash_version(value);
  if (n == "gemm_pairs") { g_gemm_pairs.store(value != 0); return WM_OK; }
  return fail(WM_ERR_SHAPE, "wm_set_option: unknown option '%s'", n.c_str());
}

int wm_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, const float* residual,
                 int64_t ldr, int res_mod, void* out_bf16, int64_t ldc_bf16, float* out_f32, int64_t ldc_f32, int M,
                 int N, int K, int act, int bn_hint, void* stream) {
  if (int rc = ensure_device()) return rc;
  if (M <= 0 || N <= 0 || K <= 0) return fail(WM_ERR_SHAPE, "wm_gemm_bf16: empty problem M=%d N=%d K=%d", M, N, K);
  if (K % 8 != 0 || lda % 8 != 0 || ldw % 8 != 0) return fail(WM_ERR_ALIGN, "wm_gemm_bf16: K, lda, ldw must be multiples of 8");
  if (out_bf16 == nullptr && out_f32 == nullptr) return fail(WM_ERR_SHAPE, "wm_gemm_bf16: no output");
  if (act < 0 || act > 3) return fail(WM_ERR_SHAPE, "wm_gemm_bf16: bad act %d", act);
  if (residual != nullptr && res_mod <= 0) return fail(WM_ERR_SHAPE, "wm_gemm_bf16: res_mod must be > 0");
  // tile selection: 512 = CTA-pair kernel (256 x 256 tile per pair of SMs) for the large GEMMs; 256 / 128 / 64 = single-CTA
  // kernel with a 128 x bn tile
  int bn = bn_hint;
  if (bn == 0) {
    if (N % 256 == 0 && M >= 256 * (t_num_sms / 2) && g_gemm_pairs.load()) bn = 512;
    else bn = (N >= 256) ? 256 : (N > 64 ? 128 : 64);
  }
  if (bn != 64 && bn != 128 && bn != 256 && bn != 512) return fail(WM_ERR_SHAPE, "wm_gemm_bf16: bn_hint must be 0/64/128/256/512");
  if (bn == 512 && N % 256 != 0) return fail(WM_ERR_SHAPE, "wm_gemm_bf16: the CTA-pair kernel needs N %% 256 == 0");
  CUtensorMap ta, tw;
  if (int rc = make_map_2d(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 128, "wm_gemm_bf16(A)")) return rc;
  if (int rc = make_map_2d(&tw, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, bn == 512 ? 128u : (uint32_t)bn, "wm_gemm_bf16(W)")) return rc;
  wm::GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.bias = bias; p.residual = residual; p.ldr = (int)ldr; p.res_mod = res_mod > 0 ? res_mod : 1;
  p.out_bf16 = reinterpret_cast<__nv_bfloat16*>(out_bf16); p.ldc_bf16 = (int)ldc_bf16;
  p.out_f32 = out_f32; p.ldc_f32 = (int)ldc_f32;
  p.act = act; p.a_mode = 0; p.conv_C = 0;
  // 16-byte vector accesses to bias / residual / outputs at columns that are multiples of 4
  p.vec_ok = (N % 4 == 0) && (bias == nullptr || aligned16(bias)) && (residual == nullptr || (aligned16(residual) && ldr % 4 == 0)) &&
             (out_bf16 == nullptr || (aligned16(out_bf16) && ldc_bf16 % 8 == 0)) &&
             (out_f32 == nullptr || (aligned16(out_f32) && ldc_f32 % 4 == 0));
  // TMA-store epilogue: exactly one output, 16-byte aligned rows; bf16 stores are 64 columns wide (tiles of >= 128 columns)
  // (with a second, bf16 output the residual has to be read: the TMA reduce-add only updates the fp32 copy)
  const bool dual = out_bf16 != nullptr && out_f32 != nullptr;
  p.res_inplace = residual != nullptr && residual == out_f32 && ldr == ldc_f32 && res_mod >= M && !dual;
  p.tma_out = (p.vec_ok && N % 8 == 0 && N >= 64 && (out_f32 != nullptr || bn >= 128) && (!dual || bn == 512)) ? (dual ? 2 : 1) : 0;
  CUtensorMap tc16 = ta, tc32 = ta;
  if (p.tma_out) {
    const uint64_t dims[2] = {(uint64_t)N, (uint64_t)M};
    if (out_bf16 != nullptr) {
      const uint64_t strides[1] = {(uint64_t)ldc_bf16 * 2};
      const uint32_t box[2] = {64, 32};
      if (int rc = make_map(&tc16, out_bf16, 2, dims, strides, box, "wm_gemm_bf16(out_bf16)")) return rc;
    }
    if (out_f32 != nullptr) {
      const uint64_t strides[1] = {(uint64_t)ldc_f32 * 4};
      const uint32_t box[2] = {32, 32};
      if (int rc = make_map(&tc32, out_f32, 2, dims, strides, box, "wm_gemm_bf16(out_f32)", CU_TENSOR_MAP_DATA_TYPE_FLOAT32)) return rc;
    }
  }
  return check_launch(wm::gemm_dispatch(ta, tw, tc16, tc32, p, bn, t_num_sms, (cudaStream_t)stream), "wm_gemm_bf16");
}

int wm_num_sms(void) {
  if (int rc = ensure_device()) return rc;
  return t_num_sms;
}

int wm_conv3x3_nhwc_bf16(const void* X, const void* W, void* out_bf16, float* out_f32, int B, int C, int N,
                         void* stream) {
  if (int rc = ensure_device()) return rc;
  if (B <= 0 || C % 64 != 0 || N <= 0 || N % 8 != 0) return fail(WM_ERR_SHAPE, "wm_conv3x3_nhwc_bf16: bad shape B=%d C=%d N=%d", B, C, N);
  if (out_bf16 == nullptr && out_f32 == nullptr) return fail(WM_ERR_SHAPE, "wm_conv3x3_nhwc_bf16: no output");
  const int bn = (N >= 256) ? 256 : (N > 64 ? 128 : 64);
  CUtensorMap ta, tw;
  const uint64_t dims[4] = {(uint64_t)C, 64, 64, (uint64_t)B};
  const uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)C * 2 * 64, (uint64_t)C * 2 * 4096};
  const uint32_t box[4] = {64, 64, 2, 1};
  if (int rc = make_map(&ta, X, 4, dims, strides, box, "wm_conv3x3_nhwc_bf16(X)")) return rc;
  if (int rc = make_map_2d(&tw, W, (uint64_t)N, (uint64_t)9 * C, (uint64_t)9 * C, (uint32_t)bn, "wm_conv3x3_nhwc_bf16(W)")) return rc;
  wm::GemmParams p{};
  p.M = B * 4096; p.N = N; p.K = 9 * C;
  p.res_mod = 1;
  p.out_bf16 = reinterpret_cast<__nv_bfloat16*>(out_bf16); p.ldc_bf16 = N;
  p.out_f32 = out_f32; p.ldc_f32 = N;
  p.a_mode = 1; p.conv_C = C;
  p.vec_ok = (N % 4 == 0) && (out_bf16 == nullptr || aligned16(out_bf16)) && (out_f32 == nullptr || aligned16(out_f32));
  p.tma_out = 0; p.res_inplace = 0;
  return check_launch(wm::gemm_dispatch(ta, tw, ta, ta, p, bn, t_num_sms, (cudaStream_t)stream), "wm_conv3x3_nhwc_bf16");
}

int wm_layernorm(const float* x, const float* gamma, const float* beta, void* y_bf16, float* y_f32, const float* add,
                 int add_mod, void* y2_bf16, int rows, int D, float eps, void* stream) {
  if (int rc = ensure_device()) return rc;
  if (!aligned16(x) || !aligned16(gamma) || !aligned16(beta) || !aligned16(y_bf16) || !aligned16(y_f32) ||
      !aligned16(add) || !aligned16(y2_bf16))
    return fail(WM_ERR_ALIGN, "wm_layernorm: pointers must be 16-byte aligned");
  if (y2_bf16 != nullptr && (add == nullptr || add_mod <= 0)) return fail(WM_ERR_SHAPE, "wm_layernorm: y2 needs add/add_mod");
  return check_launch(wm::layernorm_launch(x, gamma, beta, reinterpret_cast<__nv_bfloat16*>(y_bf16), y_f32, add,
                                           add_mod > 0 ? add_mod : 1, reinterpret_cast<__nv_bfloat16*>(y2_bf16), rows, D,
                                           eps, (cudaStream_t)stream),
                      "wm_layernorm");
}

int wm_patchify(const float* img, void* patches_bf16, void* gray_bf16, int gray_split, int B, int C, void* stream) {
  if (int rc = ensure_device()) return rc;
  if (B <= 0 || B > 65535 || (C != 1 && C != 3)) return fail(WM_ERR_SHAPE, "wm_patchify: B=%d C=%d", B, C);
  if (gray_split != 0 && gray_split != 1) return fail(WM_ERR_SHAPE, "wm_patchify: gray_split must be 0 or 1");
  if (!aligned16(img) || !aligned16(patches_bf16) || !aligned16(gray_bf16)) return fail(WM_ERR_ALIGN, "wm_patchify: alignment");
  return check_launch(wm::patchify_launch(img, reinterpret_cast<__nv_bfloat16*>(patches_bf16),
                                          reinterpret_cast<__nv_bfloat16*>(gray_bf16), gray_split, B, C, (cudaStream_t)stream),
                      "wm_patchify");
}

int wm_transpose(const void* in, void* out, int batch, int R, int C, int elt_bytes, void* stream) {
  if (int rc = ensure_device()) return rc;
  if (batch <= 0 || R <= 0 || C <= 0 || batch > 65535) return fail(WM_ERR_SHAPE, "wm_transpose: bad shape");
  return check_launch(wm::transpose_launch(in, out, batch, R, C, elt_bytes, (cudaStream_t)stream), "wm_transpose");
}

int wm_transpose_split(const float* in, void* out_bf16, int batch, int R, int C, void* stream) {
  if (int rc = ensure_device()) return rc;
  if (batch <= 0 || batch > 65535 || R <= 0 || C <= 0 || R % 64 != 0 || C % 64 != 0)
    return fail(WM_ERR_SHAPE, "wm_transpose_split: batch=%d R=%d C=%d (R and C must be multiples of 64)", batch, R, C);
  if (!aligned16(in) || !aligned16(out_bf16)) return fail(WM_ERR_ALIGN, "wm_transpose_split: alignment");
  return check_launch(wm::transpose_split_launch(in, reinterpret_cast<__nv_bfloat16*>(out_bf16), batch, R, C, (cudaStream_t)stream),
                      "wm_transpose_split");
}

int wm_hfc_finalize(const float* img, const float* low_t, void* patches_bf16, float* hfc_img, int B, void* stream) {
  if (int rc = ensure_device()) return rc;
  if (B <= 0 || B > 65535) return fail(WM_ERR_SHAPE, "wm_hfc_finalize: B=%d", B);
  return check_launch(wm::hfc_finalize_launch(img, low_t, reinterpret_cast<__nv_bfloat16*>(patches_bf16), hfc_img, B,
                                              (cudaStream_t)stream),
                      "wm_hfc_finalize");
}

int wm_add_cast(const float* a, const float* b, int b_mod, void* out_bf16, int rows, int D, void* stream) {
  if (int rc = ensure_device()) return rc;
  if (!aligned16(a) || !aligned16(b) || !aligned16(out_bf16)) return fail(WM_ERR_ALIGN, "wm_add_cast: alignment");
  if (a == nullptr && b == nullptr) return fail(WM_ERR_SHAPE, "wm_add_cast: a and b are both NULL");
  return check_launch(wm::add_cast_launch(a, b, b_mod > 0 ? b_mod : 1, reinterpret_cast<__nv_bfloat16*>(out_bf16), rows, D,
                                          (cudaStream_t)stream),
                      "wm_add_cast");
}

int wm_attn_flash(const void* q, int64_t q_rows, int64_t q_width, int64_t ldq, int q_col0, const void* k,
                  int64_t k_rows, int64_t k_width, int64_t ldk, int k_col0, const void* v, int64_t v_rows,
                  int64_t v_width, int64_t ldv, int v_col0, const void* rel_table, void* out_bf16, int64_t ldo, int B,
                  int H, int Tq, int Tk, int hd, float scale, void* stream) {
  if (int rc = ensure_device()) return rc;
  if (B <= 0 || H <= 0 || (hd != 64 && hd != 80 && hd != 128)) return fail(WM_ERR_SHAPE, "wm_attn_flash: head dim must be 64, 80 or 128");
  const bool ragged = (Tq % 256 != 0) || (Tk % 128 != 0);  // ragged tiles are masked in the kernel (no rel-pos)
  if (ragged && (rel_table != nullptr || Tq < 1 || Tk < 1))
    return fail(WM_ERR_SHAPE, "wm_attn_flash: Tq %% 256 != 0 or Tk %% 128 != 0 is not supported together with rel-pos");
  if ((int64_t)B * Tq > q_rows || (int64_t)B * Tk > k_rows || (int64_t)B * Tk > v_rows) return fail(WM_ERR_SHAPE, "wm_attn_flash: rows");
  if (q_col0 + H * hd > q_width || k_col0 + H * hd > k_width || v_col0 + H * hd > v_width) return fail(WM_ERR_SHAPE, "wm_attn_flash: columns");
  if (ldq % 8 || ldk % 8 || ldv % 8 || ldo % 8 || !aligned16(out_bf16)) return fail(WM_ERR_ALIGN, "wm_attn_flash: alignment");
  if (B > 65535 || H > 65535) return fail(WM_ERR_SHAPE, "wm_attn_flash: grid too large");
  CUtensorMap tq, tk, tv, trel;
  if (int rc = make_map_2d(&tq, q, (uint64_t)q_rows, (uint64_t)q_width, (uint64_t)ldq, 128, "wm_attn_flash(q)")) return rc;
  if (int rc = make_map_2d(&tk, k, (uint64_t)k_rows, (uint64_t)k_width, (uint64_t)ldk, 128, "wm_attn_flash(k)")) return rc;
  if (int rc = make_map_2d(&tv, v, (uint64_t)v_rows, (uint64_t)v_width, (uint64_t)ldv, 128, "wm_attn_flash(v)")) return rc;
  trel = tq;
  if (rel_table != nullptr) {
    if ((hd != 64 && hd != 80) || Tq != 4096 || Tk > 4096)
      return fail(WM_ERR_SHAPE, "wm_attn_flash: rel-pos needs head dim 64 or 80 and 64x64 tokens");
    if (int rc = make_map_2d(&trel, rel_table, 256, (uint64_t)hd, (uint64_t)hd, 16, "wm_attn_flash(rel)")) return rc;
  }
  wm::FlashParams p{};
  p.B = B; p.H = H; p.Tq = Tq; p.Tk = Tk; p.scale = scale;
  p.q_col0 = q_col0; p.k_col0 = k_col0; p.v_col0 = v_col0;
  p.out = reinterpret_cast<__nv_bfloat16*>(out_bf16); p.ldo = (int)ldo;
  p.use_relpos = rel_table != nullptr;
  // head dim 64 with rel-pos (encoder global blocks of ViT-B / ViT-L): v7 (ring of three score buffers, cheap prologue,
  // deferred publication of P: 2.29 vs 2.74 ms per launch at batch 32).  Head dims 80 / 128 (ViT-H, HFC cross-attention)
  // and the bias-free head-padded decoder attention: v4 (v7 without rel-pos measured 10 % slower per step than v4).
  const int fv = g_flash_version.load();
  if (hd == 64 && rel_table != nullptr && fv == 7)
    return check_launch(wm::flash7_dispatch(tq, tk, tv, trel, p, hd, (cudaStream_t)stream), "wm_attn_flash(v7)");
  return check_launch(wm::flash4_dispatch(tq, tk, tv, trel, p, hd, (cudaStream_t)stream), "wm_attn_flash(v4)");
}

int wm_debug_flash_trace(uint64_t* host_out_3x64x4) {
  const int rc = g_flash_version.load() == 7 ? wm::flash7_read_trace(reinterpret_cast<unsigned long long*>(host_out_3x64x4))
                                             : wm::flash4_read_trace(reinterpret_cast<unsigned long long*>(host_out_3x64x4));
  if (rc != WM_OK)
    return fail(WM_ERR_ARCH, "wm_debug_flash_trace: only available in the diagnostics build (-DWM_F3_TRACE)");
  return WM_OK;
}

int wm_debug_window_trace(uint64_t* host_out_3x64x8) {
  if (wm::window2_read_trace(reinterpret_cast<unsigned long long*>(host_out_3x64x8)) != WM_OK)
    return fail(WM_ERR_ARCH, "wm_debug_window_trace: only available in the diagnostics build (-DWM_F3_TRACE)");
  return WM_OK;
}

int wm_attn_window(const void* qkv, const void* rel_table, void* out_bf16, int B, int H, int D, float scale,
                   void* stream) {
  if (int rc = ensure_device()) return rc;
  if (B <= 0 || B > 65535 || H <= 0 || D % H != 0) return fail(WM_ERR_SHAPE, "wm_attn_window: bad B / H / D");
  const int hd = D / H;
  if (hd != 64 && hd != 80) return fail(WM_ERR_SHAPE, "wm_attn_window: head dim must be 64 or 80");
  if (!aligned16(out_bf16)) return fail(WM_ERR_ALIGN, "wm_attn_window: alignment");
  CUtensorMap tq, tkv, trel, tout;
  const uint64_t W3 = (uint64_t)3 * D;
  const uint64_t dims[4] = {W3, 64, 64, (uint64_t)B};
  const uint64_t strides[3] = {W3 * 2, W3 * 2 * 64, W3 * 2 * 4096};
  wm::WindowParams p{};
  p.B = B; p.H = H; p.scale = scale; p.D = D;
  p.out = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  // dense 14-wide boxes: a query half is 7 window rows x 14, K / V are the whole 14 x 14 window; zero padding of the
  // 70x70 grid = TMA out-of-bounds fill on loads, the crop back to 64x64 = TMA bounds check on the output store
  const uint32_t box_q[4] = {64, 14, 7, 1};
  const uint32_t box_kv[4] = {64, 14, 14, 1};
  const uint64_t odims[4] = {(uint64_t)D, 64, 64, (uint64_t)B};
  const uint64_t ostrides[3] = {(uint64_t)D * 2, (uint64_t)D * 2 * 64, (uint64_t)D * 2 * 4096};
  const uint64_t rdims[2] = {(uint64_t)hd, 64};
  const uint64_t rstrides[1] = {(uint64_t)hd * 2};
  const uint32_t rbox[2] = {64, 27};
  if (int rc = make_map(&tq, qkv, 4, dims, strides, box_q, "wm_attn_window(q)")) return rc;
  if (int rc = make_map(&tkv, qkv, 4, dims, strides, box_kv, "wm_attn_window(kv)")) return rc;
  if (int rc = make_map(&trel, rel_table, 2, rdims, rstrides, rbox, "wm_attn_window(rel)")) return rc;
  if (int rc = make_map(&tout, out_bf16, 4, odims, ostrides, box_q, "wm_attn_window(out)")) return rc;
  if (hd == 80)  // ViT-H: same structure with two 64-column sub-tiles per operand row
    return check_launch(wm::window3_dispatch(tq, tkv, trel, tout, p, t_num_sms, (cudaStream_t)stream), "wm_attn_window(hd 80)");
  return check_launch(wm::window2_dispatch(tq, tkv, trel, tout, p, t_num_sms, (cudaStream_t)stream), "wm_attn_window");
}

int wm_attn_small(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out_bf16,
                  int64_t ldo, int B, int H, int Tq, int Tk, int hd, float scale, void* stream) {
  if (int rc = ensure_device()) return rc;
  if (B <= 0 || B > 65535 || H <= 0 || Tq <= 0 || Tk <= 0) return fail(WM_ERR_SHAPE, "wm_attn_small: bad shape");
  return check_launch(wm::attn_small_launch(reinterpret_cast<const __nv_bfloat16*>(q), (int)ldq,
                                            reinterpret_cast<const __nv_bfloat16*>(k), (int)ldk,
                                            reinterpret_cast<const __nv_bfloat16*>(v), (int)ldv,
                                            reinterpret_cast<__nv_bfloat16*>(out_bf16), (int)ldo, B, H, Tq, Tk, hd, scale,
                                            (cudaStream_t)stream),
                      "wm_attn_small");
}

int wm_postprocess(const float* logits, const float* boxes, const int64_t* sizes, float thr, int from_prob,
                   float* packed, int32_t* query_idx, int64_t* labels, int32_t* counts, int B, int Q, int C1,
                   void* stream) {
  if (int rc = ensure_device()) return rc;
  if (B < 0 || Q <= 0) return fail(WM_ERR_SHAPE, "wm_postprocess: bad shape");
  return check_launch(wm::postprocess_launch(logits, boxes, reinterpret_cast<const long long*>(sizes), thr, from_prob,
                                             packed, query_idx, reinterpret_cast<long long*>(labels), counts, B, Q, C1,
                                             (cudaStream_t)stream),
                      "wm_postprocess");
}

int wm_sigmoid_topk(const float* logits, const float* boxes, float* prob_ws, int32_t* order_ws, float* scores,
                    int32_t* labels, int32_t* query, float* out_boxes, int B, int Q, int C1, int C, int K,
                    int from_prob, void* stream) {
  if (int rc = ensure_device()) return rc;
  if (B < 0 || B > 65535 || Q <= 0 || C <= 0 || C > C1) return fail(WM_ERR_SHAPE, "wm_sigmoid_topk: bad shape");
  if (!aligned16(boxes) || !aligned16(out_boxes)) return fail(WM_ERR_ALIGN, "wm_sigmoid_topk: alignment");
  return check_launch(wm::sigmoid_topk_launch(logits, boxes, prob_ws, order_ws, scores, labels, query, out_boxes, B, Q, C1,
                                              C, K, from_prob, (cudaStream_t)stream),
                      "wm_sigmoid_topk");
}

int wm_nms(const float* boxes, const float* scores, const int64_t* labels, int n, double iou_thr, int32_t* order_ws,
           uint64_t* mask_ws, int64_t* keep, int32_t* num_keep, void* stream) {
  if (int rc = ensure_device()) return rc;
  if (n < 0) return fail(WM_ERR_SHAPE, "wm_nms: n < 0");
  if (n > 0 && !aligned16(boxes)) return fail(WM_ERR_ALIGN, "wm_nms: boxes must be 16-byte aligned");
  if ((size_t)((n + 63) / 64) * 8 > 40 * 1024) return fail(WM_ERR_SHAPE, "wm_nms: n too large for the on-chip bitmap");
  return check_launch(wm::nms_launch(boxes, scores, reinterpret_cast<const long long*>(labels), n, iou_thr, order_ws,
                                     reinterpret_cast<unsigned long long*>(mask_ws), reinterpret_cast<long long*>(keep),
                                     num_keep, (cudaStream_t)stream),
                      "wm_nms");
}

int wm_nms_batched(const float* packed, const int32_t* counts, int B, int Q, float score_thr, double iou_thr,
                   int per_class, int32_t* keep_idx, int32_t* keep_cnt, void* stream) {
  if (int rc = ensure_device()) return rc;
  if (B < 0 || Q <= 0 || Q > 1024) return fail(WM_ERR_SHAPE, "wm_nms_batched: needs 0 < Q <= 1024 (use wm_nms for larger sets)");
  return check_launch(wm::nms_batched_small_launch(packed, counts, B, Q, score_thr, iou_thr, per_class, keep_idx, keep_cnt,
                                                   (cudaStream_t)stream),
                      "wm_nms_batched");
}

int wm_tiles_from_u8(const uint8_t* img, int H, int W, int64_t row_stride, const int32_t* origins, int T, int content_h,
                     int content_w, const float* mean3, const float* std3, float* out, void* stream) {
  if (int rc = ensure_device()) return rc;
  if (H <= 0 || W <= 0 || T < 0 || row_stride < (int64_t)W * 3 || content_h < 0 || content_h > 1024 || content_w < 0 ||
      content_w > 1024)
    return fail(WM_ERR_SHAPE, "wm_tiles_from_u8: bad shape H=%d W=%d T=%d stride=%lld content=%dx%d", H, W, T,
                (long long)row_stride, content_h, content_w);
  if (mean3 == nullptr || std3 == nullptr || (T > 0 && (img == nullptr || origins == nullptr || out == nullptr)))
    return fail(WM_ERR_SHAPE, "wm_tiles_from_u8: null pointer");
  if ((reinterpret_cast<uintptr_t>(out) & 15) != 0) return fail(WM_ERR_ALIGN, "wm_tiles_from_u8: out must be 16-byte aligned");
  return check_launch(wm::tiles_from_u8_launch(img, H, W, row_stride, origins, T, content_h, content_w, mean3, std3, out,
                                               (cudaStream_t)stream),
                      "wm_tiles_from_u8");
}

int wm_resize_tiles_u8(const uint8_t* img, int img_h, int img_w, int64_t row_stride, const int32_t* origins, int T, int tile_h,
                       int tile_w, uint8_t* tmp, uint8_t* out, int out_h, int out_w, const int32_t* xbounds, const int32_t* xk,
                       int xksize, const int32_t* ybounds, const int32_t* yk, int yksize, void* stream) {
  if (int rc = ensure_device()) return rc;
  if (T < 0 || T > 65535 || img_h <= 0 || img_w <= 0 || tile_h <= 0 || tile_w <= 0 || tile_h > 65535 || out_h <= 0 ||
      out_w <= 0 || out_h > 65535 || tile_h > img_h || tile_w > img_w || row_stride < (int64_t)img_w * 3 || xksize <= 0 || yksize <= 0)
    return fail(WM_ERR_SHAPE, "wm_resize_tiles_u8: bad shape (tiles must lie inside the image)");
  if (T > 0 && (img == nullptr || origins == nullptr || tmp == nullptr || out == nullptr || xbounds == nullptr || xk == nullptr ||
                ybounds == nullptr || yk == nullptr))
    return fail(WM_ERR_SHAPE, "wm_resize_tiles_u8: null pointer");
  return check_launch(wm::resize_u8_launch(img, row_stride, origins, T, tile_h, tile_w, tmp, out, out_h, out_w, xbounds, xk, xksize,
                                           ybounds, yk, yksize, (cudaStream_t)stream),
                      "wm_resize_tiles_u8");
}

int wm_merge_detections(const float* packed, const int32_t* counts, const int32_t* origins, int T, int Q, float score_thr,
                        int32_t* tile_n_ws, float* boxes, float* scores, int64_t* labels, int32_t* src, int32_t* total,
                        void* stream) {
  if (int rc = ensure_device()) return rc;
  if (T < 0 || Q <= 0 || Q > 1024) return fail(WM_ERR_SHAPE, "wm_merge_detections: needs T >= 0 and 0 < Q <= 1024");
  if (total == nullptr) return fail(WM_ERR_SHAPE, "wm_merge_detections: null pointer");
  return check_launch(wm::merge_detections_launch(packed, counts, origins, T, Q, score_thr, tile_n_ws, boxes, scores,
                                                  reinterpret_cast<long long*>(labels), src, total, (cudaStream_t)stream),
                      "wm_merge_detections");
}

int wm_match_cost(const float* logits, const float* boxes, const int64_t* tgt_ids, const float* tgt_boxes, int rows, int T,
                  int C1, float w_class, float w_bbox, float w_giou, float* cost, void* stream) {
  if (int rc = ensure_device()) return rc;
  if (rows < 0 || T < 0 || C1 < 2 || C1 > 64) return fail(WM_ERR_SHAPE, "wm_match_cost: bad shape rows=%d T=%d C1=%d", rows, T, C1);
  if ((long long)rows * T > 0 && (logits == nullptr || boxes == nullptr || tgt_ids == nullptr || tgt_boxes == nullptr || cost == nullptr))
    return fail(WM_ERR_SHAPE, "wm_match_cost: null pointer");
  return check_launch(wm::match_cost_launch(logits, boxes, reinterpret_cast<const long long*>(tgt_ids), tgt_boxes, rows, T, C1,
                                            w_class, w_bbox, w_giou, cost, (cudaStream_t)stream),
                      "wm_match_cost");
}

int wm_set_criterion(const float* logits, const float* boxes, const int32_t* m_row, const int64_t* m_label, const float* m_box,
                     int n_match, const int32_t* tgt_len, const float* empty_weight, int B, int Q, int C1, float num_boxes,
                     int32_t* tcls_ws, float* out5, void* stream) {
  if (int rc = ensure_device()) return rc;
  if (B <= 0 || Q <= 0 || C1 < 2 || C1 > 64 || n_match < 0 || !(num_boxes > 0.0f))
    return fail(WM_ERR_SHAPE, "wm_set_criterion: bad shape B=%d Q=%d C1=%d n_match=%d num_boxes=%f", B, Q, C1, n_match, num_boxes);
  if (logits == nullptr || boxes == nullptr || tgt_len == nullptr || empty_weight == nullptr || tcls_ws == nullptr || out5 == nullptr ||
      (n_match > 0 && (m_row == nullptr || m_label == nullptr || m_box == nullptr)))
    return fail(WM_ERR_SHAPE, "wm_set_criterion: null pointer");
  return check_launch(wm::criterion_launch(logits, boxes, m_row, reinterpret_cast<const long long*>(m_label), m_box, n_match,
                                           tgt_len, empty_weight, B, Q, C1, num_boxes, tcls_ws, out5, (cudaStream_t)stream),
                      "wm_set_criterion");
}

int wm_pack_coco(const float* boxes, const float* scores, const int64_t* labels, const int64_t* keep, int n_keep,
                 float* out_xywh_score, int64_t* out_category, void* stream) {
  if (int rc = ensure_device()) return rc;
  if (n_keep < 0) return fail(WM_ERR_SHAPE, "wm_pack_coco: n_keep < 0");
  return check_launch(wm::coco_pack_launch(boxes, scores, reinterpret_cast<const long long*>(labels),
                                           reinterpret_cast<const long long*>(keep), n_keep, out_xywh_score,
                                           reinterpret_cast<long long*>(out_category), (cudaStream_t)stream),
                      "wm_pack_coco");
}

}  // extern "C"
