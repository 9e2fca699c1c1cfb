// Internal declarations shared by the translation units of libgadm.so (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include "../../include/gadm.h"

namespace gadm {

// error plumbing (gadm_api.cu)
int set_cuda_error(cudaError_t e);           // records text, returns GADM_ERR_CUDA
int check_launch();                          // cudaGetLastError() -> GADM_OK / GADM_ERR_CUDA
bool initialised();

// 3-D tensor map of 2-byte elements {inner = K, rows, batch}, box {box_k, box_rows, 1}, SWIZZLE_128B (gadm_api.cu)
// dtype: 0 = bf16, 1 = fp16
int make_tmap_2b_3d(CUtensorMap* out, const void* base, uint64_t k, uint64_t rows, uint64_t batch,
                    uint32_t box_k, uint32_t box_rows, int dtype);

// Layout of the model-side `aux` buffer (gadm_prep_model), in floats, n = n_obj * M:
//   [0, n)        1 / max(|m_j|, 1e-12)                column scales of the matcher epilogue
//   [n, 4n)       model xyz, [n_obj, M, 3] fp32        (Kabsch moments)
//   [4n, 7n)      x, y, z planes, [3, n_obj, M] fp32   (SOFT epilogue: one bulk copy per plane and tile)
__host__ __device__ inline size_t aux_total_floats(int n_obj, int M) { return size_t(n_obj) * M * 7; }
__host__ __device__ inline const float* aux_scales(const float* aux, int, int) { return aux; }
__host__ __device__ inline const float* aux_xyz(const float* aux, int n_obj, int M) { return aux + size_t(n_obj) * M; }
__host__ __device__ inline const float* aux_planes(const float* aux, int n_obj, int M) { return aux + size_t(n_obj) * M * 4; }

// match_sm100.cu
int match_configure(int device);
int match_config_set(const char* key, int value);
// circle_sm100.cu
int circle_configure();
int circle_df_configure();
bool circle_df_supported(int Kp);
int circle_df_launch(const void* rows, const float* rinv_rows, const float* pad_sim, const void* cols, const float* aux,
                     const float* planes_frame, const int64_t* match_idx, const int64_t* match_idx2,
                     const int32_t* obj_id, int B, int N, int M, int Kp, int n_obj, float gamma, float margin,
                     const float* lse_p, const float* lse_n, const float* w, float* G, int Mp, float* g_pad, float* dF,
                     float* dM, cudaStream_t stream);
int match_launch(const void* rows, const float* rinv_rows, const float* pad_sim, const void* cols, const float* aux,
                 const uint8_t* mask, const int32_t* obj_id, int B, int N, int M, int Kp, int n_obj, float gamma,
                 int pad_mode, int mode, int64_t* idx, float* max_sim, float* weight, float* soft_xyz,
                 void* workspace, size_t workspace_bytes, const int32_t* n_rows, const int32_t* row_map, int N_out,
                 cudaStream_t stream);
size_t match_workspace_bytes();
int circle_launch(const void* rows, const float* rinv_rows, const float* pad_sim, const void* cols, const float* aux,
                  const float* planes_frame, const int64_t* match_idx, const int64_t* match_idx2, const uint8_t* fg,
                  const int32_t* obj_id, int B,
                  int N, int M, int Kp, int n_obj, float gamma, float margin, float* loss, float* lse_p,
                  float* lse_n, const float* w, float* G, int Mp, float* g_pad, cudaStream_t stream);

// prep.cu
int prep_rows_launch(const void* feat, int feat_bf16, const int32_t* pos, int B, int d, int N, int operand_mode,
                     int pad_mode, void* rows, float* rinv, float* pad_sim, cudaStream_t stream);
int compact_rows_launch(const uint8_t* mask, int B, int N, int32_t* pos, int32_t* row_map, int32_t* n_sel,
                        cudaStream_t stream);
int pack_outputs_launch(const int64_t* idx, const float* max_sim, const float* weight, const float* soft_xyz, size_t n,
                        int32_t* out, cudaStream_t stream);
int pack_u16_launch(const int32_t* idx, size_t n, uint16_t* out, cudaStream_t stream);
int prep_model_launch(const float* mesh, const float* model_xyz, int n_obj, int d, int M, int operand_mode, void* cols,
                      float* aux, cudaStream_t stream);
int seg_mask_launch(const float* seg, int B, int N, uint8_t* mask, cudaStream_t stream);
int kabsch_moments_launch(const int64_t* idx, const uint8_t* mask, const float* weight, const float* cloud,
                          const float* aux, const int32_t* obj_id, int B, int N, int M, int n_obj, double* out,
                          cudaStream_t stream);
int kabsch_pose_launch(const double* mom, const double* count, const uint8_t* det, int B, int min_pts, float* poses,
                       cudaStream_t stream);

// knn3d.cu
int knn3d_configure();
int knn3d_config_set(const char* key, int value);
size_t knn3d_workspace_bytes(const gadm_knn_job* jobs, int n_jobs, int algo);
int knn3d_launch(const float* support, const float* query, const gadm_knn_job* jobs, int n_jobs, int algo,
                 int32_t* idx, float* dist2, void* workspace, size_t workspace_bytes, cudaStream_t stream);

// knn_feat.cu
int knn_feat_configure();
int knn_feat_launch(const float* x, int B, int C, int N, int kdim, int k, int64_t* idx, cudaStream_t stream);
// tensor-core path (knn_feat_tc.cu): C % 64 == 0, k <= 20, N % 4 == 0, N >= 256
int knn_feat_tc_configure();
bool knn_feat_tc_supported(int C, int N, int kdim, int k);
size_t knn_feat_tc_workspace_bytes(int B, int C, int N);
int knn_feat_tc_launch(const float* x, int B, int C, int N, int k, int64_t* idx, void* workspace, size_t workspace_bytes,
                       cudaStream_t stream);

// gather.cu
int graph_feature_launch(const float* x, const int64_t* idx, int B, int C, int N, int k, float* out,
                         void* workspace, size_t workspace_bytes, cudaStream_t stream);
int group_fwd_launch(const float* features, const int32_t* idx, int b, int c, int n, int m, int s, float* out,
                     cudaStream_t stream);
int group_bwd_launch(const float* grad_out, const int32_t* idx, int b, int c, int n, int m, int s,
                     float* grad_features, cudaStream_t stream);
int gather_neighbour_launch(const float* pc, const int64_t* idx, int B, int N, int C, int M, int K, float* out,
                            cudaStream_t stream);
int gather_max_launch(const float* feature, const int64_t* idx, int B, int C, int N, int M, int K, float* out,
                      cudaStream_t stream);
int graph_feature_bwd_launch(const float* grad_out, const int64_t* idx, int B, int C, int N, int k, float* grad_x,
                             cudaStream_t stream);
int gather_neighbour_bwd_launch(const float* grad_out, const int64_t* idx, int B, int N, int C, int M, int K,
                                float* grad_pc, cudaStream_t stream);
int gather_max_bwd_launch(const float* feature, const int64_t* idx, const float* grad_out, int B, int C, int N, int M,
                          int K, float* grad_feature, cudaStream_t stream);
int relative_pos_encoding_launch(const float* xyz, const int64_t* idx, int B, int N, int K, float* out,
                                 cudaStream_t stream);

}  // namespace gadm
