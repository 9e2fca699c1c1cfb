// C ABI of libgadm.so (include/gadm.h): argument validation, error codes, tensor-map construction.
// No device allocation, no synchronisation, no exceptions cross this boundary.
#include <stdio.h>
#include <string.h>

#include "gadm_internal.h"

namespace gadm {

namespace {
thread_local char g_cuda_err[256] = "";
bool g_init = false;

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled g_encode = nullptr;
}  // namespace

int set_cuda_error(cudaError_t e) {
  snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", cudaGetErrorName(e), cudaGetErrorString(e));
  return GADM_ERR_CUDA;
}
int check_launch() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? GADM_OK : set_cuda_error(e);
}
bool initialised() { return g_init; }

int make_tmap_2b_3d(CUtensorMap* out, const void* base, uint64_t k, uint64_t rows, uint64_t batch,
                    uint32_t box_k, uint32_t box_rows, int dtype) {
  if (!g_encode) return GADM_ERR_NOT_INIT;
  cuuint64_t dims[3] = {k, rows, batch};
  cuuint64_t strides[2] = {k * 2, k * rows * 2};  // bytes, dims 1 and 2
  cuuint32_t box[3] = {box_k, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(out, dtype ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_cuda_err, sizeof(g_cuda_err), "cuTensorMapEncodeTiled failed: CUresult %d", int(r));
    return GADM_ERR_CUDA;
  }
  return GADM_OK;
}

}  // namespace gadm

using namespace gadm;

#define GADM_REQUIRE_INIT() \
  do { if (!initialised()) return GADM_ERR_NOT_INIT; } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" {

const char* gadm_strerror(int status) {
  switch (status) {
    case GADM_OK: return "ok";
    case GADM_ERR_BAD_ARG: return "bad argument (null pointer, non-positive size or inconsistent shapes)";
    case GADM_ERR_UNSUPPORTED: return "unsupported configuration (d % 64, M % 8, k > 32, mode ...)";
    case GADM_ERR_ALIGN: return "pointer not 16-byte aligned";
    case GADM_ERR_WORKSPACE: return "workspace too small";
    case GADM_ERR_CUDA: return "CUDA error (see gadm_last_cuda_error)";
    case GADM_ERR_ARCH: return "device is not sm_100 (B200); libgadm has no fallback path";
    case GADM_ERR_NOT_INIT: return "gadm_init() has not succeeded";
    default: return "unknown gadm status";
  }
}

int gadm_abi_version(void) { return GADM_ABI_VERSION; }

int gadm_config_set(const char* key, int value) {
  if (!key) return GADM_ERR_BAD_ARG;
  if (!strncmp(key, "knn.", 4)) return knn3d_config_set(key, value);
  return match_config_set(key, value);
}

const char* gadm_last_cuda_error(void) { return g_cuda_err; }

int gadm_init(int device) {
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return set_cuda_error(e);
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return set_cuda_error(e);
  if (prop.major != 10) return GADM_ERR_ARCH;
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess) return set_cuda_error(e);
    if (qres != cudaDriverEntryPointSuccess || fn == nullptr) {
      snprintf(g_cuda_err, sizeof(g_cuda_err), "cuTensorMapEncodeTiled not found in the driver");
      return GADM_ERR_CUDA;
    }
    g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
  }
  int rc = match_configure(device);
  if (rc != GADM_OK) return rc;
  rc = circle_configure();
  if (rc != GADM_OK) return rc;
  rc = circle_df_configure();
  if (rc != GADM_OK) return rc;
  rc = knn3d_configure();
  if (rc != GADM_OK) return rc;
  rc = knn_feat_tc_configure();
  if (rc != GADM_OK) return rc;
  rc = knn_feat_configure();
  if (rc != GADM_OK) return rc;
  g_init = true;
  return GADM_OK;
}

int gadm_operand_k(int d, int operand_mode) {
  if (d <= 0) return GADM_ERR_BAD_ARG;
  if (operand_mode == GADM_OPERAND_BF16 || operand_mode == GADM_OPERAND_BF16N) return d;
  if (operand_mode == GADM_OPERAND_BF16X3) return 3 * d;
  return GADM_ERR_UNSUPPORTED;
}

size_t gadm_aux_floats(int n_obj, int M) {
  if (n_obj <= 0 || M <= 0) return 0;
  return aux_total_floats(n_obj, M);
}

int gadm_prep_rows(const float* feat, int B, int d, int N, int operand_mode, int pad_mode, void* rows, float* rinv,
                   float* pad_sim, gadm_stream_t stream) {
  GADM_REQUIRE_INIT();
  if (!feat || !rows || !rinv || B <= 0 || d <= 0 || N <= 0) return GADM_ERR_BAD_ARG;
  if (pad_mode != GADM_PAD_NONE && !pad_sim) return GADM_ERR_BAD_ARG;
  if (pad_mode < 0 || pad_mode > GADM_PAD_E0) return GADM_ERR_UNSUPPORTED;
  if (operand_mode < GADM_OPERAND_BF16 || operand_mode > GADM_OPERAND_BF16N) return GADM_ERR_UNSUPPORTED;
  if (d % 64 != 0 || d > 256) return GADM_ERR_UNSUPPORTED;
  if (!aligned16(rows)) return GADM_ERR_ALIGN;
  return prep_rows_launch(feat, 0, nullptr, B, d, N, operand_mode, pad_mode, rows, rinv, pad_sim, (cudaStream_t)stream);
}

int gadm_compact_rows(const uint8_t* mask, int B, int N, int32_t* pos, int32_t* row_map, int32_t* n_sel,
                      gadm_stream_t stream) {
  GADM_REQUIRE_INIT();
  if (!mask || !pos || !row_map || !n_sel || B <= 0 || N <= 0) return GADM_ERR_BAD_ARG;
  return compact_rows_launch(mask, B, N, pos, row_map, n_sel, (cudaStream_t)stream);
}

int gadm_prep_rows_sel(const void* feat, int feat_is_bf16, const int32_t* pos, int B, int d, int N, int operand_mode,
                       int pad_mode, void* rows, float* rinv, float* pad_sim, gadm_stream_t stream) {
  GADM_REQUIRE_INIT();
  if (!feat || !pos || !rows || !rinv || B <= 0 || d <= 0 || N <= 0) return GADM_ERR_BAD_ARG;
  if (pad_mode != GADM_PAD_NONE && !pad_sim) return GADM_ERR_BAD_ARG;
  if (pad_mode < 0 || pad_mode > GADM_PAD_E0) return GADM_ERR_UNSUPPORTED;
  if (operand_mode < GADM_OPERAND_BF16 || operand_mode > GADM_OPERAND_BF16N) return GADM_ERR_UNSUPPORTED;
  if (feat_is_bf16 && operand_mode == GADM_OPERAND_BF16X3) return GADM_ERR_UNSUPPORTED;
  if (d % 64 != 0 || d > 256) return GADM_ERR_UNSUPPORTED;
  if (!aligned16(rows)) return GADM_ERR_ALIGN;
  return prep_rows_launch(feat, feat_is_bf16 != 0, pos, B, d, N, operand_mode, pad_mode, rows, rinv, pad_sim,
                          (cudaStream_t)stream);
}

int gadm_prep_rows_bf16(const void* feat_bf16, int B, int d, int N, int operand_mode, int pad_mode, void* rows,
                        float* rinv, float* pad_sim, gadm_stream_t stream) {
  GADM_REQUIRE_INIT();
  if (!feat_bf16 || !rows || !rinv || B <= 0 || d <= 0 || N <= 0) return GADM_ERR_BAD_ARG;
  if (pad_mode != GADM_PAD_NONE && !pad_sim) return GADM_ERR_BAD_ARG;
  if (pad_mode < 0 || pad_mode > GADM_PAD_E0) return GADM_ERR_UNSUPPORTED;
  // a bf16 source has no low part to split: BF16X3 would only triple the work
  if (operand_mode != GADM_OPERAND_BF16 && operand_mode != GADM_OPERAND_BF16N) return GADM_ERR_UNSUPPORTED;
  if (d % 64 != 0 || d > 256) return GADM_ERR_UNSUPPORTED;
  if (!aligned16(rows)) return GADM_ERR_ALIGN;
  return prep_rows_launch(feat_bf16, 1, nullptr, B, d, N, operand_mode, pad_mode, rows, rinv, pad_sim,
                          (cudaStream_t)stream);
}

int gadm_pack_match_outputs(const int64_t* idx, const float* max_sim, const float* weight, const float* soft_xyz,
                            int64_t n, int32_t* out, gadm_stream_t stream) {
  GADM_REQUIRE_INIT();
  if (!idx || !max_sim || !out || n <= 0) return GADM_ERR_BAD_ARG;
  return pack_outputs_launch(idx, max_sim, weight, soft_xyz, size_t(n), out, (cudaStream_t)stream);
}

int gadm_pack_indices_u16(const int32_t* idx, int64_t n, uint16_t* out, gadm_stream_t stream) {
  GADM_REQUIRE_INIT();
  if (!idx || !out || n <= 0) return GADM_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(idx) & 7) || (reinterpret_cast<uintptr_t>(out) & 3)) return GADM_ERR_ALIGN;
  return pack_u16_launch(idx, size_t(n), out, (cudaStream_t)stream);
}

int gadm_prep_model(const float* mesh, const float* model_xyz, int n_obj, int d, int M, int operand_mode, void* cols,
                    float* aux, gadm_stream_t stream) {
  GADM_REQUIRE_INIT();
  if (!mesh || !cols || !aux || n_obj <= 0 || d <= 0 || M <= 0) return GADM_ERR_BAD_ARG;
  if (operand_mode < GADM_OPERAND_BF16 || operand_mode > GADM_OPERAND_BF16N) return GADM_ERR_UNSUPPORTED;
  if (d % 64 != 0 || d > 256 || M % 8 != 0) return GADM_ERR_UNSUPPORTED;
  if (!aligned16(cols) || !aligned16(aux)) return GADM_ERR_ALIGN;
  return prep_model_launch(mesh, model_xyz, n_obj, d, M, operand_mode, cols, aux, (cudaStream_t)stream);
}

static int match_validate(const void* rows, const float* rinv_rows, const float* pad_sim, const void* cols,
                          const float* aux, const int32_t* obj_id, int B, int N, int M, int Kp, int n_obj, float gamma,
                          int pad_mode, int mode, const int64_t* idx, const float* max_sim, const float* weight,
                          const float* soft_xyz, const void* workspace) {
  if (!rows || !rinv_rows || !cols || !aux || !idx || !max_sim) return GADM_ERR_BAD_ARG;
  if (workspace && !aligned16(workspace)) return GADM_ERR_ALIGN;
  if (B <= 0 || N <= 0 || M <= 0 || Kp <= 0 || n_obj <= 0) return GADM_ERR_BAD_ARG;
  if (mode < GADM_MATCH_ARGMAX || mode > GADM_MATCH_ARGMAX_BF16N) return GADM_ERR_UNSUPPORTED;
  if (mode == GADM_MATCH_SOFT && (!weight || !soft_xyz)) return GADM_ERR_BAD_ARG;
  // 2^(gamma log2(e) cos) is summed without a reference exponent: keep it well inside the fp32 range
  if (mode == GADM_MATCH_SOFT && !(gamma >= -40.f && gamma <= 40.f)) return GADM_ERR_UNSUPPORTED;
  if (pad_mode < 0 || pad_mode > GADM_PAD_E0) return GADM_ERR_UNSUPPORTED;
  if (pad_mode != GADM_PAD_NONE && !pad_sim) return GADM_ERR_BAD_ARG;
  if (Kp % 64 != 0 || Kp > 768 || M % 8 != 0) return GADM_ERR_UNSUPPORTED;
  if (obj_id == nullptr && n_obj != 1 && n_obj != B) return GADM_ERR_BAD_ARG;
  if (!aligned16(rows) || !aligned16(cols) || !aligned16(aux)) return GADM_ERR_ALIGN;
  return GADM_OK;
}

int gadm_match_fwd(const void* rows, const float* rinv_rows, const float* pad_sim, const void* cols, const float* aux,
                   const uint8_t* mask, const int32_t* obj_id, int B, int N, int M, int Kp, int n_obj, float gamma,
                   int pad_mode, int mode, int64_t* idx, float* max_sim, float* weight, float* soft_xyz,
                   void* workspace, size_t workspace_bytes, gadm_stream_t stream) {
  GADM_REQUIRE_INIT();
  const int rc = match_validate(rows, rinv_rows, pad_sim, cols, aux, obj_id, B, N, M, Kp, n_obj, gamma, pad_mode, mode,
                                idx, max_sim, weight, soft_xyz, workspace);
  if (rc != GADM_OK) return rc;
  return match_launch(rows, rinv_rows, pad_sim, cols, aux, mask, obj_id, B, N, M, Kp, n_obj, gamma, pad_mode, mode,
                      idx, max_sim, weight, soft_xyz, workspace, workspace_bytes, nullptr, nullptr, N,
                      (cudaStream_t)stream);
}

int gadm_match_fwd_sel(const void* rows, const float* rinv_rows, const float* pad_sim, const void* cols,
                       const float* aux, const int32_t* n_rows, const int32_t* row_map, int N_out,
                       const int32_t* obj_id, int B, int N, int M, int Kp, int n_obj, float gamma, int pad_mode, int mode,
                       int64_t* idx, float* max_sim, float* weight, float* soft_xyz, void* workspace,
                       size_t workspace_bytes, gadm_stream_t stream) {
  GADM_REQUIRE_INIT();
  const int rc = match_validate(rows, rinv_rows, pad_sim, cols, aux, obj_id, B, N, M, Kp, n_obj, gamma, pad_mode, mode,
                                idx, max_sim, weight, soft_xyz, workspace);
  if (rc != GADM_OK) return rc;
  if (!n_rows || N_out <= 0) return GADM_ERR_BAD_ARG;
  if (row_map == nullptr && N_out < N) return GADM_ERR_BAD_ARG;
  return match_launch(rows, rinv_rows, pad_sim, cols, aux, nullptr, obj_id, B, N, M, Kp, n_obj, gamma, pad_mode, mode,
                      idx, max_sim, weight, soft_xyz, workspace, workspace_bytes, n_rows, row_map, N_out,
                      (cudaStream_t)stream);
}

int gadm_circle_loss_fwd(const void* rows, const float* rinv_rows, const float* pad_sim, const void* cols,
                         const float* aux, const float* planes_frame, const int64_t* match_idx, const int64_t* match_idx2,
                         const uint8_t* fg,
                         const int32_t* obj_id, int B, int N, int M, int Kp, int n_obj, float gamma, float margin,
                         float* loss, float* lse_p, float* lse_n, gadm_stream_t stream) {
  GADM_REQUIRE_INIT();
  if (!rows || !rinv_rows || !pad_sim || !cols || !aux || !planes_frame || !match_idx || !loss || !lse_p || !lse_n)
    return GADM_ERR_BAD_ARG;
  if (B <= 0 || N <= 0 || M <= 0 || Kp <= 0 || n_obj <= 0) return GADM_ERR_BAD_ARG;
  if (B > 65535 || Kp % 64 != 0 || Kp > 768 || M % 8 != 0) return GADM_ERR_UNSUPPORTED;
  if (obj_id == nullptr && n_obj != 1 && n_obj != B) return GADM_ERR_BAD_ARG;
  if (!(margin >= 0.f && margin < 1.f) || !(gamma > 0.f)) return GADM_ERR_BAD_ARG;
  // 2^logit is summed without a running maximum: |logit| <= gamma (2 + m)(2 - m) must stay inside the fp32 range
  if (gamma * (2.f + margin) * (2.f - margin) * 1.4426950408889634f > 120.f) return GADM_ERR_UNSUPPORTED;
  if (!aligned16(rows) || !aligned16(cols) || !aligned16(aux) || !aligned16(planes_frame)) return GADM_ERR_ALIGN;
  return circle_launch(rows, rinv_rows, pad_sim, cols, aux, planes_frame, match_idx, match_idx2, fg, obj_id, B, N, M, Kp,
                       n_obj, gamma, margin, loss, lse_p, lse_n, nullptr, nullptr, 0, nullptr, (cudaStream_t)stream);
}

static int circle_bwd_checked(const void* rows, const float* rinv_rows, const float* pad_sim, const void* cols,
                              const float* aux, const float* planes_frame, const int64_t* match_idx,
                              const int64_t* match_idx2, const int32_t* obj_id, int B, int N, int M, int Kp, int n_obj,
                              float gamma, float margin, const float* lse_p, const float* lse_n, const float* w, void* G,
                              int Mp, float* g_pad, gadm_stream_t stream) {
  GADM_REQUIRE_INIT();
  if (!rows || !rinv_rows || !pad_sim || !cols || !aux || !planes_frame || !match_idx || !lse_p || !lse_n || !w || !G)
    return GADM_ERR_BAD_ARG;
  if (B <= 0 || N <= 0 || M <= 0 || Kp <= 0 || n_obj <= 0) return GADM_ERR_BAD_ARG;
  if (B > 65535 || Kp % 64 != 0 || Kp > 768 || M % 8 != 0) return GADM_ERR_UNSUPPORTED;
  if (Mp < M + 1 || Mp % 8 != 0) return GADM_ERR_BAD_ARG;   // rows of G are written with 32-byte stores
  if (obj_id == nullptr && n_obj != 1 && n_obj != B) return GADM_ERR_BAD_ARG;
  if (!(margin >= 0.f && margin < 1.f) || !(gamma > 0.f)) return GADM_ERR_BAD_ARG;
  if (gamma * (2.f + margin) * (2.f - margin) * 1.4426950408889634f > 120.f) return GADM_ERR_UNSUPPORTED;
  if (!aligned16(rows) || !aligned16(cols) || !aligned16(aux) || !aligned16(planes_frame) ||
      (reinterpret_cast<uintptr_t>(G) & 31) != 0)
    return GADM_ERR_ALIGN;
  return circle_launch(rows, rinv_rows, pad_sim, cols, aux, planes_frame, match_idx, match_idx2, nullptr, obj_id, B, N, M,
                       Kp, n_obj, gamma, margin, nullptr, const_cast<float*>(lse_p), const_cast<float*>(lse_n), w,
                       static_cast<float*>(G), Mp, g_pad, (cudaStream_t)stream);
}

int gadm_circle_loss_bwd(const void* rows, const float* rinv_rows, const float* pad_sim, const void* cols,
                         const float* aux, const float* planes_frame, const int64_t* match_idx, const int64_t* match_idx2,
                         const int32_t* obj_id, int B, int N, int M, int Kp, int n_obj, float gamma, float margin,
                         const float* lse_p, const float* lse_n, const float* w, float* G, int Mp,
                         gadm_stream_t stream) {
  return circle_bwd_checked(rows, rinv_rows, pad_sim, cols, aux, planes_frame, match_idx, match_idx2, obj_id, B, N, M, Kp,
                            n_obj, gamma, margin, lse_p, lse_n, w, G, Mp, nullptr, stream);
}

int gadm_circle_loss_bwd_split(const void* rows, const float* rinv_rows, const float* pad_sim, const void* cols,
                               const float* aux, const float* planes_frame, const int64_t* match_idx,
                               const int64_t* match_idx2, const int32_t* obj_id, int B, int N, int M, int Kp, int n_obj,
                               float gamma, float margin, const float* lse_p, const float* lse_n, const float* w,
                               void* G2, int Mp, float* g_pad, gadm_stream_t stream) {
  if (!g_pad) return GADM_ERR_BAD_ARG;
  return circle_bwd_checked(rows, rinv_rows, pad_sim, cols, aux, planes_frame, match_idx, match_idx2, obj_id, B, N, M, Kp,
                            n_obj, gamma, margin, lse_p, lse_n, w, G2, Mp, g_pad, stream);
}

int gadm_circle_loss_bwd_fused(const void* rows, const float* rinv_rows, const float* pad_sim, const void* cols,
                               const float* aux, const float* planes_frame, const int64_t* match_idx,
                               const int64_t* match_idx2, const int32_t* obj_id, int B, int N, int M, int Kp, int n_obj,
                               float gamma, float margin, const float* lse_p, const float* lse_n, const float* w,
                               void* G2, int Mp, float* g_pad, float* dF, float* dM, gadm_stream_t stream) {
  GADM_REQUIRE_INIT();
  if (!rows || !rinv_rows || !pad_sim || !cols || !aux || !planes_frame || !match_idx || !lse_p || !lse_n || !w ||
      (!G2 && !dM) || !g_pad || !dF)
    return GADM_ERR_BAD_ARG;
  if (B <= 0 || N <= 0 || M <= 0 || Kp <= 0 || n_obj <= 0) return GADM_ERR_BAD_ARG;
  if (B > 65535 || M % 8 != 0 || !circle_df_supported(Kp)) return GADM_ERR_UNSUPPORTED;
  if (Mp < M + 1 || Mp % 8 != 0) return GADM_ERR_BAD_ARG;
  if (obj_id == nullptr && n_obj != 1 && n_obj != B) return GADM_ERR_BAD_ARG;
  if (!(margin >= 0.f && margin < 1.f) || !(gamma > 0.f)) return GADM_ERR_BAD_ARG;
  if (gamma * (2.f + margin) * (2.f - margin) * 1.4426950408889634f > 120.f) return GADM_ERR_UNSUPPORTED;
  if (!aligned16(rows) || !aligned16(cols) || !aligned16(aux) || !aligned16(planes_frame) ||
      (reinterpret_cast<uintptr_t>(G2) & 31) != 0 || (reinterpret_cast<uintptr_t>(dF) & 31) != 0 ||
      (reinterpret_cast<uintptr_t>(dM) & 15) != 0)
    return GADM_ERR_ALIGN;
  return circle_df_launch(rows, rinv_rows, pad_sim, cols, aux, planes_frame, match_idx, match_idx2, obj_id, B, N, M, Kp,
                          n_obj, gamma, margin, lse_p, lse_n, w, static_cast<float*>(G2), Mp, g_pad, dF, dM,
                          (cudaStream_t)stream);
}

size_t gadm_match_workspace_bytes(void) {
  if (!initialised()) return 0;
  return match_workspace_bytes();
}

int gadm_kabsch_moments(const int64_t* idx, const uint8_t* mask, const float* cloud, const float* aux,
                        const int32_t* obj_id, int B, int N, int M, int n_obj, double* out, gadm_stream_t stream) {
  GADM_REQUIRE_INIT();
  if (!idx || !cloud || !aux || !out || B <= 0 || N <= 0 || M <= 0 || n_obj <= 0) return GADM_ERR_BAD_ARG;
  if (obj_id == nullptr && n_obj != 1 && n_obj != B) return GADM_ERR_BAD_ARG;
  return kabsch_moments_launch(idx, mask, nullptr, cloud, aux, obj_id, B, N, M, n_obj, out, (cudaStream_t)stream);
}

int gadm_kabsch_moments_w(const int64_t* idx, const uint8_t* mask, const float* weight, const float* cloud,
                          const float* aux, const int32_t* obj_id, int B, int N, int M, int n_obj, double* out,
                          gadm_stream_t stream) {
  GADM_REQUIRE_INIT();
  if (!idx || !weight || !cloud || !aux || !out || B <= 0 || N <= 0 || M <= 0 || n_obj <= 0) return GADM_ERR_BAD_ARG;
  if (obj_id == nullptr && n_obj != 1 && n_obj != B) return GADM_ERR_BAD_ARG;
  return kabsch_moments_launch(idx, mask, weight, cloud, aux, obj_id, B, N, M, n_obj, out, (cudaStream_t)stream);
}

int gadm_kabsch_poses(const double* moments, const double* count_moments, const uint8_t* det, int B, int min_pts,
                      float* poses, gadm_stream_t stream) {
  GADM_REQUIRE_INIT();
  if (!moments || !poses || B <= 0) return GADM_ERR_BAD_ARG;
  return kabsch_pose_launch(moments, count_moments, det, B, min_pts, poses, (cudaStream_t)stream);
}

static int validate_jobs(const gadm_knn_job* jobs, int n_jobs) {
  if (!jobs || n_jobs <= 0) return GADM_ERR_BAD_ARG;
  for (int i = 0; i < n_jobs; ++i) {
    const gadm_knn_job& j = jobs[i];
    if (j.n_support <= 0 || j.n_query <= 0 || j.batch <= 0 || j.k <= 0) return GADM_ERR_BAD_ARG;
    if (j.k > 32) return GADM_ERR_UNSUPPORTED;
    if (j.k > j.n_support) return GADM_ERR_BAD_ARG;  // the reference leaves stale ids here (knn_.cxx:120-121)
    if (j.support_off < 0 || j.query_off < 0 || j.out_off < 0) return GADM_ERR_BAD_ARG;
  }
  return GADM_OK;
}

size_t gadm_knn3d_workspace_bytes(const gadm_knn_job* jobs_host, int n_jobs, int algo) {
  if (validate_jobs(jobs_host, n_jobs) != GADM_OK) return 0;
  return knn3d_workspace_bytes(jobs_host, n_jobs, algo);
}

int gadm_knn3d(const float* support, const float* query, const gadm_knn_job* jobs_host, int n_jobs, int algo,
               int32_t* idx, float* dist2, void* workspace, size_t workspace_bytes, gadm_stream_t stream) {
  GADM_REQUIRE_INIT();
  if (!support || !query || !idx) return GADM_ERR_BAD_ARG;
  int rc = validate_jobs(jobs_host, n_jobs);
  if (rc != GADM_OK) return rc;
  if (algo < GADM_KNN_BRUTE || algo > GADM_KNN_AUTO) return GADM_ERR_UNSUPPORTED;
  return knn3d_launch(support, query, jobs_host, n_jobs, algo, idx, dist2, workspace, workspace_bytes,
                      (cudaStream_t)stream);
}

int gadm_knn_feat(const float* x, int B, int C, int N, int kdim, int k, int64_t* idx, gadm_stream_t stream) {
  GADM_REQUIRE_INIT();
  if (!x || !idx || B <= 0 || C <= 0 || N <= 0 || kdim <= 0 || kdim > C || k <= 0) return GADM_ERR_BAD_ARG;
  if (k > 32) return GADM_ERR_UNSUPPORTED;
  if (k > N) return GADM_ERR_BAD_ARG;
  return knn_feat_launch(x, B, C, N, kdim, k, idx, (cudaStream_t)stream);
}

size_t gadm_knn_feat_tc_workspace_bytes(int B, int C, int N, int kdim, int k) {
  if (B <= 0 || C <= 0 || N <= 0 || !knn_feat_tc_supported(C, N, kdim, k)) return 0;
  return knn_feat_tc_workspace_bytes(B, C, N);
}

int gadm_knn_feat_tc(const float* x, int B, int C, int N, int k, int64_t* idx, void* workspace, size_t workspace_bytes,
                     gadm_stream_t stream) {
  GADM_REQUIRE_INIT();
  if (!x || !idx || B <= 0 || C <= 0 || N <= 0 || k <= 0) return GADM_ERR_BAD_ARG;
  if (k > N) return GADM_ERR_BAD_ARG;
  return knn_feat_tc_launch(x, B, C, N, k, idx, workspace, workspace_bytes, (cudaStream_t)stream);
}

size_t gadm_graph_feature_workspace_bytes(int B, int C, int N) {
  if (B <= 0 || C <= 0 || N <= 0 || C % 4 != 0) return 0;
  return size_t(B) * C * N * sizeof(float);
}

int gadm_graph_feature(const float* x, const int64_t* idx, int B, int C, int N, int k, float* out, void* workspace,
                       size_t workspace_bytes, gadm_stream_t stream) {
  GADM_REQUIRE_INIT();
  if (!x || !idx || !out || B <= 0 || C <= 0 || N <= 0 || k <= 0) return GADM_ERR_BAD_ARG;
  return graph_feature_launch(x, idx, B, C, N, k, out, workspace, workspace_bytes, (cudaStream_t)stream);
}

int gadm_group_fwd(const float* features, const int32_t* idx, int b, int c, int n, int m, int s, float* out,
                   gadm_stream_t stream) {
  GADM_REQUIRE_INIT();
  if (!features || !idx || !out || b <= 0 || c <= 0 || n <= 0 || m <= 0 || s <= 0) return GADM_ERR_BAD_ARG;
  return group_fwd_launch(features, idx, b, c, n, m, s, out, (cudaStream_t)stream);
}

int gadm_group_bwd(const float* grad_out, const int32_t* idx, int b, int c, int n, int m, int s, float* grad_features,
                   gadm_stream_t stream) {
  GADM_REQUIRE_INIT();
  if (!grad_out || !idx || !grad_features || b <= 0 || c <= 0 || n <= 0 || m <= 0 || s <= 0) return GADM_ERR_BAD_ARG;
  return group_bwd_launch(grad_out, idx, b, c, n, m, s, grad_features, (cudaStream_t)stream);
}

int gadm_gather_neighbour(const float* pc, const int64_t* idx, int B, int N, int C, int M, int K, float* out,
                          gadm_stream_t stream) {
  GADM_REQUIRE_INIT();
  if (!pc || !idx || !out || B <= 0 || N <= 0 || C <= 0 || M <= 0 || K <= 0) return GADM_ERR_BAD_ARG;
  return gather_neighbour_launch(pc, idx, B, N, C, M, K, out, (cudaStream_t)stream);
}

int gadm_gather_max(const float* feature, const int64_t* idx, int B, int C, int N, int M, int K, float* out,
                    gadm_stream_t stream) {
  GADM_REQUIRE_INIT();
  if (!feature || !idx || !out || B <= 0 || C <= 0 || N <= 0 || M <= 0 || K <= 0) return GADM_ERR_BAD_ARG;
  return gather_max_launch(feature, idx, B, C, N, M, K, out, static_cast<cudaStream_t>(stream));
}

int gadm_relative_pos_encoding(const float* xyz, const int64_t* idx, int B, int N, int K, float* out,
                               gadm_stream_t stream) {
  GADM_REQUIRE_INIT();
  if (!xyz || !idx || !out || B <= 0 || N <= 0 || K <= 0) return GADM_ERR_BAD_ARG;
  if (reinterpret_cast<uintptr_t>(out) & 7) return GADM_ERR_ALIGN;
  return relative_pos_encoding_launch(xyz, idx, B, N, K, out, static_cast<cudaStream_t>(stream));
}

int gadm_graph_feature_bwd(const float* grad_out, const int64_t* idx, int B, int C, int N, int k, float* grad_x,
                           gadm_stream_t stream) {
  GADM_REQUIRE_INIT();
  if (!grad_out || !idx || !grad_x || B <= 0 || C <= 0 || N <= 0 || k <= 0) return GADM_ERR_BAD_ARG;
  return graph_feature_bwd_launch(grad_out, idx, B, C, N, k, grad_x, static_cast<cudaStream_t>(stream));
}

int gadm_gather_neighbour_bwd(const float* grad_out, const int64_t* idx, int B, int N, int C, int M, int K,
                              float* grad_pc, gadm_stream_t stream) {
  GADM_REQUIRE_INIT();
  if (!grad_out || !idx || !grad_pc || B <= 0 || N <= 0 || C <= 0 || M <= 0 || K <= 0) return GADM_ERR_BAD_ARG;
  return gather_neighbour_bwd_launch(grad_out, idx, B, N, C, M, K, grad_pc, static_cast<cudaStream_t>(stream));
}

int gadm_gather_max_bwd(const float* feature, const int64_t* idx, const float* grad_out, int B, int C, int N, int M, int K,
                        float* grad_feature, gadm_stream_t stream) {
  GADM_REQUIRE_INIT();
  if (!feature || !idx || !grad_out || !grad_feature || B <= 0 || C <= 0 || N <= 0 || M <= 0 || K <= 0)
    return GADM_ERR_BAD_ARG;
  return gather_max_bwd_launch(feature, idx, grad_out, B, C, N, M, K, grad_feature, static_cast<cudaStream_t>(stream));
}

int gadm_seg_mask(const float* seg, int B, int N, uint8_t* mask, gadm_stream_t stream) {
  GADM_REQUIRE_INIT();
  if (!seg || !mask || B <= 0 || N <= 0) return GADM_ERR_BAD_ARG;
  return seg_mask_launch(seg, B, N, mask, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
