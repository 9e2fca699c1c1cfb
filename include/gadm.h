/*
 * gadm.h -- C ABI of libgadm.so: the B200 (sm_100a) implementation of the scene-to-model dense
 * correspondence path of Ray0089/geometric-aware-dense-matching (geoMatch matching head + geometric kNN).
 *
 * Every entry point below replaces one reference interface (file:line relative to the reference tree):
 *
 *   gadm_prep_rows / gadm_prep_model / gadm_match_fwd
 *       evaluator.py:77-93            cal_frame_poses: normalize rows, normalize columns, matmul, torch.max
 *       utils/pvn3d_eval_utils_kpls.py:436-444, models/geoMatch.py:117-119   "-1" pad column variant
 *       models/geoMatch_DGCNN.py:92-99                                       "e0" pad column variant
 *       (+ the soft-correspondence extension: softmax weights and soft model coordinates)
 *   gadm_circle_loss_fwd / gadm_circle_loss_bwd / gadm_circle_loss_bwd_split / gadm_circle_loss_bwd_fused
 *       models/geoMatch.py:55-83, 102-157 (geoMatch_DGCNN.py:53-78, 80-136) + models/loss.py:475-490
 *       pointwise_feature_matching + matching_loss + CircleLoss.forward, and its backward pass: dL/dsim as fp32, as two
 *       bf16 parts for tensor-core gradient products, or with the gradient products formed inside the kernel
 *   gadm_seg_mask
 *       evaluator.py:78,82            seg argmax -> foreground mask
 *   gadm_kabsch_moments
 *       utils/pvn3d_eval_utils_kpls.py:43-76   best_fit_transform (the moments of the two point sets)
 *   gadm_knn3d
 *       models/RandLA/utils/nearest_neighbors/knn_.h:11-17 / knn_.cxx:71-135   cpp_knn_batch(_omp)
 *       models/RandLA/helper_tool.py:161-170                                   DataProcessing.knn_search
 *       lib/pointops/functions/pointops.py:435-493                             knnquery / knnquery_heap
 *   gadm_knn_feat
 *       models/dgcnn.py:21-27         knn (feature-space dynamic-graph kNN)
 *   gadm_graph_feature
 *       models/dgcnn.py:30-56         get_graph_feature (gather + cat(nbr - x, x))
 *   gadm_group_fwd / gadm_group_bwd
 *       lib/pointops/functions/pointops.py:149-178   Grouping.forward / backward
 *   gadm_gather_neighbour
 *       models/RandLA/RandLANet.py:729-738           Building_block.gather_neighbour
 *
 * Conventions
 *   - All tensor pointers are DEVICE pointers unless a parameter says HOST.  The caller owns all memory,
 *     including workspaces; the library never allocates or frees device memory.  gadm_init() is the only call
 *     that may synchronise the device; every other call only enqueues work.
 *   - Every call enqueues work on `stream` and returns 0 (GADM_OK) or a negative gadm_status.  No exceptions.
 *   - Re-entrant: after gadm_init() the only mutable global state is the set of switches of gadm_config_set()
 *     (kernel selection for profiling and tests; read once per call, never from the environment).  gadm_init()
 *     keeps what it learns about a device (SM count) per device: call it once for every device that is used.
 *   - The cubin is sm_100a only.  gadm_init() fails with GADM_ERR_ARCH elsewhere; there is no fallback.
 */
#ifndef GADM_H_
#define GADM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* gadm_stream_t; /* == cudaStream_t */

typedef enum {
  GADM_OK = 0,
  GADM_ERR_BAD_ARG = -1,     /* null pointer, non-positive size, inconsistent shapes            */
  GADM_ERR_UNSUPPORTED = -2, /* d not a multiple of 64 / too large, k > 32, unknown mode ...    */
  GADM_ERR_ALIGN = -3,       /* pointer not aligned as documented                               */
  GADM_ERR_WORKSPACE = -4,   /* workspace too small                                             */
  GADM_ERR_CUDA = -5,        /* a CUDA runtime / driver call or the launch failed               */
  GADM_ERR_ARCH = -6,        /* device is not compute capability 10.x                           */
  GADM_ERR_NOT_INIT = -7     /* gadm_init() has not succeeded on this process                   */
} gadm_status;

const char* gadm_strerror(int status);
#define GADM_ABI_VERSION 3   /* bumped whenever an existing entry point changes its signature */
int gadm_abi_version(void);
/* Checks the device (cc 10.x), resolves cuTensorMapEncodeTiled, raises the kernels' shared-memory limits and records
 * the device's SM count (per device).  Makes `device` current. */
int gadm_init(int device);
/* Kernel-selection switches for profiling and tests; -1 restores the automatic choice.  Keys:
 *   "match.alt"   0 forbids the alternating persistent ARGMAX kernel        "match.pair"  1 / 0 forces / forbids the paired-row kernel
 *   "match.rt"    1 / 2 row tiles per CTA of the generic kernel             "match.ctas"  grid of the persistent kernels (<= SM count)
 *   "match.cta2"  1: CTA pairs (clusters of two, tcgen05.mma.cta_group::2) in the paired-row kernel
 *   "match.alt_cta2"  0: single CTAs instead of CTA pairs in the alternating ARGMAX kernel
 *   "knn.ppc"     target points per occupied grid cell column (1..64, default 16)
 *   "knn.grid_min" GADM_KNN_AUTO scans clouds with fewer points than this (default 128)
 * Results do not depend on any of them.  GADM_ERR_BAD_ARG for an unknown key.  The library reads no environment
 * variable. */
int gadm_config_set(const char* key, int value);
/* cudaGetLastError text of the most recent GADM_ERR_CUDA on this thread ("" if none). */
const char* gadm_last_cuda_error(void);

/* ------------------------------------------------------------------------------------------------
 * Matching head
 * ------------------------------------------------------------------------------------------------ */
enum { GADM_PAD_NONE = 0, GADM_PAD_MINUS_ONE = 1, GADM_PAD_E0 = 2 };
/* operand_mode: how fp32 descriptors become bf16 tensor-core operands.
 *   BF16   : round once to bf16 (exact if the inputs are bf16-representable); K' = d
 *   BF16X3 : hi/lo split, operands [hi|hi|lo] x [hi|lo|hi]; K' = 3d; ~fp32-faithful (error ~1e-6)
 *   BF16N  : model columns are L2-normalised in fp32 (F.normalize, evaluator.py:90) BEFORE the single rounding
 *            to bf16; scene rows as BF16; K' = d.  The column scales in `aux` become 1/||bf16 column|| = 1 +- ~2e-4,
 *            so every mode stays exact with respect to the rounded operands, and GADM_MATCH_ARGMAX_UNIT may
 *            drop the scales altogether.                                                              */
enum { GADM_OPERAND_BF16 = 0, GADM_OPERAND_BF16X3 = 1, GADM_OPERAND_BF16N = 2 };
/* outputs: ARGMAX = idx + max_sim only (the reference's path); SOFT adds weight + soft_xyz.
 * ARGMAX_UNIT = ARGMAX for operands prepared with GADM_OPERAND_BF16N, treating the column norms as exactly 1
 * DURING THE SEARCH (no per-column constant in the epilogue: the fastest kernel); the winner's similarity is
 * reported with its true scale.  The searched scores differ from ARGMAX's by the factor ||bf16 column|| =
 * 1 +- 2^-8 at most (~3e-4 rms at d = 128), so the index can differ from ARGMAX's only on rows whose top-1
 * margin is below |score| * 2^-7; falls back to ARGMAX when the unit kernel does not apply.
 * ARGMAX_BF16N = ARGMAX (same results, bit for bit) for operands prepared with GADM_OPERAND_BF16N: the caller asserts
 * that every column scale is <= 1 + 2^-8, which lets the kernel skip whole 32-column chunks whose raw maximum cannot
 * beat a running maximum (no scale loads, no multiplies for ~60 % of the chunks).                              */
enum { GADM_MATCH_ARGMAX = 0, GADM_MATCH_SOFT = 1, GADM_MATCH_ARGMAX_UNIT = 2, GADM_MATCH_ARGMAX_BF16N = 3 };

/* K' for a given d / operand_mode (d for BF16 and BF16N, 3d for BF16X3). */
int gadm_operand_k(int d, int operand_mode);

/* Scene side.  feat [B, d, N] fp32 channel-major (end_points['rgbd'], models/geoMatch.py:199).
 *   rows  [B, N, K'] bf16   point-major operand (16-byte aligned)
 *   rinv  [B, N] fp32       1 / max(||f||, 1e-12)  (F.normalize eps, evaluator.py:89)
 *   pad_sim [B, N] fp32 or NULL: similarity of each row with the pad column (pad_mode != NONE)       */
int gadm_prep_rows(const float* feat, int B, int d, int N, int operand_mode, int pad_mode, void* rows,
                   float* rinv, float* pad_sim, gadm_stream_t stream);
/* The same for descriptors that are already bf16 ([B, d, N] bf16 channel-major): identical outputs to gadm_prep_rows
 * on the fp32 values they represent, half the bytes to read (and to upload).  BF16 / BF16N operand modes only. */
int gadm_prep_rows_bf16(const void* feat_bf16, int B, int d, int N, int operand_mode, int pad_mode, void* rows,
                        float* rinv, float* pad_sim, gadm_stream_t stream);

/* Model side (once per object bank).  mesh [n_obj, d, M] fp32 channel-major (end_points['mesh']);
 * model_xyz [n_obj, M, 3] fp32 or NULL.
 *   cols [n_obj, M, K'] bf16 (16-byte aligned)
 *   aux  [n_obj, ceil(M/256), 4, 256] fp32 (16-byte aligned): per 256-vertex tile the rows
 *        {1/max(||m||,1e-12), x, y, z}; pad columns of the last tile are zero.  gadm_aux_floats() sizes it. */
size_t gadm_aux_floats(int n_obj, int M);
int gadm_prep_model(const float* mesh, const float* model_xyz, int n_obj, int d, int M, int operand_mode,
                    void* cols, float* aux, gadm_stream_t stream);

/* Fused similarity + row-wise argmax / online softmax / soft coordinates.  The [N, M] score matrix never
 * exists in HBM.  For frame b the model is obj_id[b] (NULL: b if n_obj == B, else 0).
 *   mask [B, N] uint8 or NULL: rows with mask == 0 get idx = -1 and zeros.
 *   idx [B, N] int64 (M means "pad column won"), max_sim [B, N] fp32,
 *   weight [B, N] fp32 and soft_xyz [B, N, 3] fp32 (may be NULL unless mode == GADM_MATCH_SOFT).
 * Kp = K' from gadm_operand_k(); must be a multiple of 64 and <= 768.
 * gamma (SOFT): softmax temperature, |gamma| <= 40 (the terms 2^(gamma log2(e) cos) are summed without a reference
 * exponent; GADM_ERR_UNSUPPORTED beyond that).
 * workspace (optional, 16-byte aligned, gadm_match_workspace_bytes() bytes on the current device -- about 80 KB per
 * SM: argmax scratch, arrival counters and partial results of the persistent kernels; contents irrelevant on entry
 * and exit).  The persistent kernels deal the (frame, 256-row block, model tile) units out evenly to one CTA per SM
 * and need it.  Give launches that may overlap on different streams one workspace each.  NULL selects the kernels
 * that need none; results are identical either way.                                                      */
size_t gadm_match_workspace_bytes(void);
int gadm_match_fwd(const void* rows, const float* rinv_rows, const float* pad_sim, const void* cols,
                   const float* aux, const uint8_t* mask, const int32_t* obj_id, int B, int N, int M, int Kp,
                   int n_obj, float gamma, int pad_mode, int mode, int64_t* idx, float* max_sim, float* weight,
                   float* soft_xyz, void* workspace, size_t workspace_bytes, gadm_stream_t stream);

/* Row compaction, evaluator.py:82-88 (cls_msk -> rgbd_features[cls_msk]) without the host round trip.
 *   gadm_compact_rows   mask [B, N] uint8 -> pos [B, N] int32 (rank of a selected point among the selected points of its
 *                       frame, -1 otherwise), row_map [B, N] int32 (row_map[b, j] = the point with rank j; entries
 *                       >= n_sel[b] undefined), n_sel [B] int32.  Order is preserved: compacted order IS the reference's.
 *   gadm_prep_rows_sel  as gadm_prep_rows / gadm_prep_rows_bf16 (feat_is_bf16), but point n lands in row pos[b, n] of
 *                       rows / rinv / pad_sim (still [B, N, .]: N is the capacity); unselected points are skipped.
 *   gadm_match_fwd_sel  as gadm_match_fwd over frames of n_rows[b] <= N rows (device memory; blocks of rows beyond it
 *                       cost nothing -- the persistent ARGMAX kernel rebalances its schedule on the device).  Row j of
 *                       frame b is written to position row_map[b, j] of an output frame of N_out rows (row_map NULL:
 *                       position j, the compacted order, N_out >= N).  Positions no row maps to are left untouched:
 *                       pre-fill idx with -1 for the scatter form.  B <= 2048 for the persistent kernel.        */
int gadm_compact_rows(const uint8_t* mask, int B, int N, int32_t* pos, int32_t* row_map, int32_t* n_sel,
                      gadm_stream_t stream);
int gadm_prep_rows_sel(const void* feat, int feat_is_bf16, const int32_t* pos, int B, int d, int N, int operand_mode,
                       int pad_mode, void* rows, float* rinv, float* pad_sim, gadm_stream_t stream);
int gadm_match_fwd_sel(const void* rows, const float* rinv_rows, const float* pad_sim, const void* cols,
                       const float* aux, const int32_t* n_rows, const int32_t* row_map, int N_out,
                       const int32_t* obj_id, int B, int N, int M, int Kp, int n_obj, float gamma, int pad_mode, int mode,
                       int64_t* idx, float* max_sim, float* weight, float* soft_xyz, void* workspace,
                       size_t workspace_bytes, gadm_stream_t stream);

/* Packs the matcher outputs of n scene points into records of six 32-bit words {int32 idx, max_sim, weight, x, y, z}
 * (weight / soft_xyz may be NULL: zeros), so that one contiguous copy carries them to the host -- the reference pulls
 * idx and the cloud back tensor by tensor (evaluator.py:87,99).  idx must fit int32 (it is < M + 1). */
int gadm_pack_match_outputs(const int64_t* idx, const float* max_sim, const float* weight, const float* soft_xyz,
                            int64_t n, int32_t* out, gadm_stream_t stream);

/* Narrows n int32 neighbour indices (gadm_knn3d output) to uint16 for transport: the caller guarantees that every
 * support cloud has fewer than 65536 points (indices are truncated otherwise).  The reference ships the 22 index
 * arrays of a sample as int32 numpy arrays (datasets/lm/linemod_pbr.py:534-569); here they cross the bus once, at half
 * the bytes.  idx 8-byte aligned, out 4-byte aligned (GADM_ERR_ALIGN). */
int gadm_pack_indices_u16(const int32_t* idx, int64_t n, uint16_t* out, gadm_stream_t stream);

/* Flash-style CircleLoss forward, the training-side twin of the matcher.  Replaces, per batch,
 * models/geoMatch.py:102-157 (similarity of the foreground rows with the -1-padded, normalised model), :55-83
 * (positive mask) and models/loss.py:475-490 (CircleLoss.forward) without materialising sim [n_fg, M + 1]:
 *   rows / rinv_rows / pad_sim  from gadm_prep_rows with the variant's pad column (GADM_PAD_MINUS_ONE for
 *                models/geoMatch.py:117-119, GADM_PAD_E0 for models/geoMatch_DGCNN.py:95-98); cols / aux from gadm_prep_model
 *   planes_frame [4, B, M] fp32 (16-byte aligned): per-frame x / y / z planes of the model vertices in which the
 *                vertices with visible_flag == 0 are moved to 1e18 (they can then never be positives), and the
 *                squared positive radius of every vertex (a constant for geoMatch.py:24,67; positive_r / 1000 * the
 *                vertex's camera-space depth for geoMatch_DGCNN.py:64-65)
 *   match_idx [B, N] int64: ground-truth vertex of every row, M = off the model (pad column is its only positive)
 *   fg [B, N] uint8 or NULL: rows that take part (labels == 1); the others get 0
 *   positive j for row i: |xyz[match_idx[i]] - planes_frame[0:3, b, j]|^2 + 1e-7 < planes_frame[3, b, j]   (basic_utils.py:86-89)
 *   match_idx2 [B, N] int64 or NULL.  Non-NULL selects the symmetry-aware positives of GeoMatch.matching_loss_sys
 *                (models/geoMatch.py:86-100): the positives of row i are exactly the columns match_idx[i] and
 *                match_idx2[i] (M = the pad column); planes_frame is not consulted for them
 *   loss [B, N] = softplus(LSE_p + LSE_n) per row; lse_p / lse_n [B, N]: the two natural-log LSEs (what a backward
 *   pass needs).  The mean over rows / samples (geoMatch.py:150-156) is left to the caller.
 * gamma (2 + margin)(2 - margin) log2(e) must be <= 120 (gamma = 16, margin = 0.2: 91).                          */
int gadm_circle_loss_fwd(const void* rows, const float* rinv_rows, const float* pad_sim, const void* cols,
                         const float* aux, const float* planes_frame, const int64_t* match_idx, const int64_t* match_idx2,
                         const uint8_t* fg,
                         const int32_t* obj_id, int B, int N, int M, int Kp, int n_obj, float gamma, float margin,
                         float* loss, float* lse_p, float* lse_n, gadm_stream_t stream);

/* Backward companion: dL/dsim for the rows' upstream gradients, recomputed from the same operands (the forward
 * pass keeps 8 bytes per row, not the similarity matrix).
 *   lse_p / lse_n [B, N] from gadm_circle_loss_fwd;  w [B, N] = dL/d(LSE_p + LSE_n) of every row, i.e.
 *   sigmoid(LSE_p + LSE_n) * dL/dloss_row (0 for rows that take no part)
 *   G [B, N, Mp] fp32 (32-byte aligned, Mp >= M + 1, Mp % 8 == 0): G[b, i, j] = dL/dsim_ij for j < M, column M = the
 *   pad column, columns M + 1 .. Mp - 1 = 0.  ap / an are constants as in the reference (detach, loss.py:479-480).
 * The two gradient GEMMs that follow (G M^ and G^T F^) are plain library GEMMs on the caller's side.                */
int gadm_circle_loss_bwd(const void* rows, const float* rinv_rows, const float* pad_sim, const void* cols,
                         const float* aux, const float* planes_frame, const int64_t* match_idx, const int64_t* match_idx2,
                         const int32_t* obj_id, int B, int N, int M, int Kp, int n_obj, float gamma, float margin,
                         const float* lse_p, const float* lse_n, const float* w, float* G, int Mp,
                         gadm_stream_t stream);

/* The same pass with the gradient written for tensor-core GEMMs on the forward pass's own bf16 operands.
 *   G2 [B, N, 2 Mp] bf16 (the same bytes as G above): G''[b, i, j] = dL/dsim_ij * (1/|f_i|) * (1/|m_j|) as an unevaluated
 *   sum hi + lo of two bf16 (16 mantissa bits, relative error <= 2^-17), every group of 8 columns stored as its 8 hi
 *   parts followed by its 8 lo parts: element [b, i, 16 g + 8 h + e] = part h of column 8 g + e.  Columns >= M are 0.
 *   g_pad [B, N] fp32: dL/dsim of the pad column (unscaled).
 * With K2[16 g + 8 h + e] = 8 g + e the two gradient products are plain bf16 GEMMs with fp32 accumulation:
 *   dL/df^_i = (1/rinv_i) * sum_k G2[i, k] * cols[K2[k]]  + g_pad_i * m^_pad
 *   dL/dm^_j = (1/scale_j) * sum_{h} (G2^T rows)[16 g + 8 h + e],  j = 8 g + e
 * where rows / cols are the bf16 operands of the forward pass and rinv / scale their fp32 norms' reciprocals.
 * Same argument rules as gadm_circle_loss_bwd.                                                                      */
int gadm_circle_loss_bwd_split(const void* rows, const float* rinv_rows, const float* pad_sim, const void* cols,
                               const float* aux, const float* planes_frame, const int64_t* match_idx,
                               const int64_t* match_idx2, const int32_t* obj_id, int B, int N, int M, int Kp, int n_obj,
                               float gamma, float margin, const float* lse_p, const float* lse_n, const float* w,
                               void* G2, int Mp, float* g_pad, gadm_stream_t stream);

/* gadm_circle_loss_bwd_split with the scene-side gradient product fused into the kernel (K' <= 128): the tile of
 * G'' the epilogue has just formed goes to shared memory as the A operand of a second tcgen05 MMA against the model tile
 * that is already resident (read as an MN-major B operand), accumulated over the model tiles in tensor memory:
 *   dF [B, N, K'] fp32 (32-byte aligned) = sum_j G''[b, i, j] * cols[j]  =  rinv_i * (dL/df^_i - g_pad_i * m^_pad)
 * dM == NULL: G2 and g_pad are written as by gadm_circle_loss_bwd_split (the model-side product G''^T rows is a library
 *   GEMM on the caller's side).
 * dM != NULL ([B, Mp, K'] fp32, 16-byte aligned, ZEROED by the caller): the model-side product is formed in the kernel
 *   too -- per model tile a third MMA G''^T x rows (both operands read MN-major from the buffers already in shared
 *   memory), added to dM[b] with fp32 reductions (the order of the additions, hence the last bits, varies from run to
 *   run): dM[b, j] = scale_j * dL/dm^_j of frame b.  dL/dsim then never leaves the SM: G2 is not written (may be NULL).
 * GADM_ERR_UNSUPPORTED for K' > 128.                                                                                */
int gadm_circle_loss_bwd_fused(const void* rows, const float* rinv_rows, const float* pad_sim, const void* cols,
                               const float* aux, const float* planes_frame, const int64_t* match_idx,
                               const int64_t* match_idx2, const int32_t* obj_id, int B, int N, int M, int Kp, int n_obj,
                               float gamma, float margin, const float* lse_p, const float* lse_n, const float* w,
                               void* G2, int Mp, float* g_pad, float* dF, float* dM, gadm_stream_t stream);

/* Backward passes of the gathers: the reference's torch.gather / max / cat are differentiable in the features
 * (models/dgcnn.py:41-54, models/RandLA/RandLANet.py:90-120, :729-738) and sit inside the trained networks.  Each zeroes
 * its output and scatter-adds with fp32 atomics (the summation order, hence the last bits, may vary from run to run).
 *   graph_feature_bwd     grad_out [B, 2C, N, k] -> grad_x [B, C, N]
 *   gather_neighbour_bwd  grad_out [B, M, K, C]  -> grad_pc [B, N, C]
 *   gather_max_bwd        grad_out [B, C, M]     -> grad_feature [B, C, N]; the gradient of a maximum goes to the FIRST
 *                         neighbour that attains it (feature / idx as in the forward call)                             */
int gadm_graph_feature_bwd(const float* grad_out, const int64_t* idx, int B, int C, int N, int k, float* grad_x,
                           gadm_stream_t stream);
int gadm_gather_neighbour_bwd(const float* grad_out, const int64_t* idx, int B, int N, int C, int M, int K,
                              float* grad_pc, gadm_stream_t stream);
int gadm_gather_max_bwd(const float* feature, const int64_t* idx, const float* grad_out, int B, int C, int N, int M, int K,
                        float* grad_feature, gadm_stream_t stream);

/* Foreground mask of the matcher (evaluator.py:78,82: `seg_res = argmax(seg_features, dim=0); cls_msk = seg_res == 1`)
 * without the argmax tensor: seg [B, 2, N] fp32 -> mask [B, N] uint8 = seg[b,1,n] > seg[b,0,n]  (torch.argmax returns
 * the first maximal index, so a tie is background). */
int gadm_seg_mask(const float* seg, int B, int N, uint8_t* mask, gadm_stream_t stream);

/* Moments for the least-squares pose fit that follows the matcher (best_fit_transform): per frame, over rows
 * with idx in [0, M) (and mask != 0):  n, sum A, sum B, sum A B^T  with A = model xyz[idx], B = cloud point.
 *   cloud [B, N, 3] fp32; out [B, 16] fp64 = {n, sA[3], sB[3], sAB[9]}  (zeroed by the callee)           */
int gadm_kabsch_moments(const int64_t* idx, const uint8_t* mask, const float* cloud, const float* aux,
                        const int32_t* obj_id, int B, int N, int M, int n_obj, double* out,
                        gadm_stream_t stream);
/* The same with a weight per scene point (weight [B, N] fp32, e.g. the matcher's softmax weight): out[b] =
 * {sum w, sum w A, sum w B, sum w A B^T} -- weighted Procrustes. */
int gadm_kabsch_moments_w(const int64_t* idx, const uint8_t* mask, const float* weight, const float* cloud,
                          const float* aux, const int32_t* obj_id, int B, int N, int M, int n_obj, double* out,
                          gadm_stream_t stream);
/* best_fit_transform (utils/pvn3d_eval_utils_kpls.py:43-76, torch twin utils/basic_utils.py:848-880) ON THE DEVICE, one
 * frame per thread, fp64: moments [B, 16] -> poses [B, 3, 4] fp32 = [R | t] with B ~ R A + t, 3x3 SVD by one-sided
 * Jacobi, reflection fix on the smallest singular direction (:67-69).  Frames the reference answers with its sentinel
 * (evaluator.py:69-72, :83, :96: det[b] == 0, <= 1 or fewer than min_pts matched rows) get [I | (0, 0, -1000)].
 * count_moments (NULL: the weight sum counts): moments of an UNWEIGHTED pass whose [b, 0] is the number of pairs, for
 * weighted fits.  det [B] uint8 or NULL. */
int gadm_kabsch_poses(const double* moments, const double* count_moments, const uint8_t* det, int B, int min_pts,
                      float* poses, gadm_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Exact 3-D kNN (RandLA nearest_neighbors / pointops knnquery)
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
  int64_t support_off; /* in points, relative to `support`                                       */
  int64_t query_off;   /* in points, relative to `query`                                         */
  int64_t out_off;     /* in elements, relative to `idx` / `dist2`                               */
  int64_t support_bstride, query_bstride, out_bstride; /* per batch item (points / elements)     */
  int32_t n_support, n_query, k, batch;
} gadm_knn_job;

enum { GADM_KNN_BRUTE = 0, GADM_KNN_GRID = 1, GADM_KNN_AUTO = 2 };

/* Device workspace needed by gadm_knn3d for these jobs (0 for GADM_KNN_BRUTE). */
size_t gadm_knn3d_workspace_bytes(const gadm_knn_job* jobs_host, int n_jobs, int algo);

/* For every job and batch item: the k nearest support points of each query, ascending squared distance
 * d2 = ((dx*dx) + (dy*dy)) + (dz*dz) in fp32 without FMA contraction (nanoflann.hpp:343-346); equal
 * distances ordered by ascending index.  Requires 1 <= k <= min(32, n_support).
 *   support, query: [.., 3] fp32 ; idx: int32 [.., n_query, k] ; dist2: fp32 same shape or NULL.
 *   jobs_host: HOST array (read before the call returns).                                            */
int gadm_knn3d(const float* support, const float* query, const gadm_knn_job* jobs_host, int n_jobs, int algo,
               int32_t* idx, float* dist2, void* workspace, size_t workspace_bytes, gadm_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * DGCNN dynamic-graph kNN + edge features
 * ------------------------------------------------------------------------------------------------ */
/* x [B, C, N] fp32; ranks by  -|xi|^2 + 2 xi.xj - |xj|^2  over the first `kdim` channels (kdim = C, or 3
 * for dim9=True, models/dgcnn.py:38); idx int64 [B, N, k], nearest first, ties by ascending index. k <= 32 */
int gadm_knn_feat(const float* x, int B, int C, int N, int kdim, int k, int64_t* idx, gadm_stream_t stream);

/* The same ranking on the tensor cores (bf16x3 split dot products, fp32 |x|^2, fp32 accumulate), for
 * kdim == C, C % 64 == 0, C <= 256, k <= 20, N % 4 == 0, N >= 256.  gadm_knn_feat_tc_workspace_bytes returns 0 when
 * the shape is not supported (use gadm_knn_feat), else the bytes of device scratch (256-byte aligned) the call needs
 * for the split operands: 4 C + 4 bytes per point.  Scores differ from gadm_knn_feat's by ~2^-16 |xi||xj|: neighbours
 * whose distances are closer than that may swap ranks (torch's own GEMM reorders sums on the same scale). */
size_t gadm_knn_feat_tc_workspace_bytes(int B, int C, int N, int kdim, int k);
int gadm_knn_feat_tc(const float* x, int B, int C, int N, int k, int64_t* idx, void* workspace, size_t workspace_bytes,
                     gadm_stream_t stream);

/* out [B, 2C, N, k] fp32: out[b, c, n, j] = x[b, c, idx[b,n,j]] - x[b, c, n];  out[b, C+c, n, j] = x[b, c, n]
 * workspace (optional, 16-byte aligned, gadm_graph_feature_workspace_bytes): a point-major copy of x that makes
 * the neighbour gathers sector-efficient; without it a slower direct-gather kernel runs (same result).      */
size_t gadm_graph_feature_workspace_bytes(int B, int C, int N);
int gadm_graph_feature(const float* x, const int64_t* idx, int B, int C, int N, int k, float* out, void* workspace,
                       size_t workspace_bytes, gadm_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Grouping (pointops) and neighbour gather (RandLA)
 * ------------------------------------------------------------------------------------------------ */
/* features (b, c, n) fp32, idx (b, m, s) int32 -> out (b, c, m, s): out[b,c,m,s] = features[b,c,idx[b,m,s]] */
int gadm_group_fwd(const float* features, const int32_t* idx, int b, int c, int n, int m, int s, float* out,
                   gadm_stream_t stream);
/* grad_out (b, c, m, s) -> grad_features (b, c, n), scatter-add; grad_features is zeroed by the callee. */
int gadm_group_bwd(const float* grad_out, const int32_t* idx, int b, int c, int n, int m, int s,
                   float* grad_features, gadm_stream_t stream);
/* pc (B, N, C) fp32, idx (B, M, K) int64 -> out (B, M, K, C): out[b,m,k,:] = pc[b, idx[b,m,k], :] */
int gadm_gather_neighbour(const float* pc, const int64_t* idx, int B, int N, int C, int M, int K, float* out,
                          gadm_stream_t stream);

/* RandLA pooling / interpolation (models/RandLA/RandLANet.py:90-105 random_sample, :107-120 nearest_interpolation):
 * feature (B, C, N) fp32, idx (B, M, K) int64 -> out (B, C, M): out[b,c,m] = max_k feature[b, c, idx[b,m,k]]
 * (K = 1: a plain gather).  The reference materialises the (B, C, M*K) gather first. */
int gadm_gather_max(const float* feature, const int64_t* idx, int B, int C, int N, int M, int K, float* out,
                    gadm_stream_t stream);
/* RandLA relative position encoding (models/RandLA/RandLANet.py:720-727): xyz (B, N, 3) fp32, idx (B, N, K) int64
 * -> out (B, N, K, 10) = [ |p - q|, p - q, p, q ] with p = xyz[b,n], q = xyz[b, idx[b,n,k]]; the distance is
 * sqrt((dx*dx + dy*dy) + dz*dz) in fp32. */
int gadm_relative_pos_encoding(const float* xyz, const int64_t* idx, int B, int N, int K, float* out,
                               gadm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GADM_H_ */
