// Operand preparation for the fused matcher (one pass each, HBM-bound):
//   channel-major fp32 descriptors [G, d, P]  ->  point-major bf16 operands [G, P, K'] (TMA/UMMA K-major),
//   fp32 inverse L2 norms (F.normalize eps = 1e-12, evaluator.py:89-90), the pad-column similarity
//   (geoMatch.py:117-119 / geoMatch_DGCNN.py:95-98) on the scene side, and the {x,y,z,1/|m|} aux table on the
//   model side.  Also the Kabsch moment reduction that follows the matcher
//   (utils/pvn3d_eval_utils_kpls.py:57-63).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "gadm_internal.h"

namespace gadm {

namespace {

constexpr int PTS = 32;  // points per CTA (one 128-byte line of every channel row)

// side: 0 = scene rows ([hi|hi|lo] in x3 mode), 1 = model columns ([hi|lo|hi] in x3 mode)
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }

// TIn: float, or __nv_bfloat16 for descriptors that arrive already rounded (exactly the values a float source holds
// after the rounding this kernel would apply: same operands, half the bytes)
template <int kSide, typename TIn>
__global__ void __launch_bounds__(256)
prep_kernel(const TIn* __restrict__ src, const float* __restrict__ xyz, int d, int P, int x3, int prenorm, int pad_mode,
            __nv_bfloat16* __restrict__ dst, float* __restrict__ rinv, float* __restrict__ pad_sim,
            float* __restrict__ aux_scale, float* __restrict__ aux_xyz, float* __restrict__ aux_planes, size_t plane,
            const int32_t* __restrict__ pos) {
  extern __shared__ float tile[];  // [d][PTS + 1]
  const int g = blockIdx.y;
  const int p0 = blockIdx.x * PTS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Kp = x3 ? 3 * d : d;

  const TIn* s = src + size_t(g) * d * P;
  for (int c = warp; c < d; c += 8) {
    const int p = p0 + lane;
    tile[c * (PTS + 1) + lane] = p < P ? to_f32(s[size_t(c) * P + p]) : 0.f;  // coalesced along points
  }
  __syncthreads();

  for (int pl = warp; pl < PTS; pl += 8) {  // one warp per point
    const int p = p0 + pl;
    if (p >= P) break;  // warp-uniform
    // row compaction (scene side): point p goes to row pos[g, p] of its frame, or nowhere
    const int prow = pos ? pos[size_t(g) * P + p] : p;
    if (prow < 0) continue;  // warp-uniform
    float ss = 0.f, sum = 0.f;
    __nv_bfloat16* out = dst + (size_t(g) * P + prow) * Kp;
    float pre = 1.f;
    if (kSide == 1 && prenorm) {   // BF16N: F.normalize in fp32 first, then the one rounding to bf16
      float s2 = 0.f;
      for (int c = lane; c < d; c += 32) {
        const float v = tile[c * (PTS + 1) + pl];
        s2 = fmaf(v, v, s2);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      pre = 1.f / fmaxf(sqrtf(s2), 1e-12f);
    }
    for (int c = lane * 2; c < d; c += 64) {
      const float v0 = tile[c * (PTS + 1) + pl] * pre, v1 = tile[(c + 1) * (PTS + 1) + pl] * pre;
      const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
      const float f0 = __bfloat162float(h0), f1 = __bfloat162float(h1);
      __nv_bfloat162 hi;
      hi.x = h0; hi.y = h1;
      if (x3) {
        __nv_bfloat162 lo;
        lo.x = __float2bfloat16_rn(v0 - f0);
        lo.y = __float2bfloat16_rn(v1 - f1);
        // the operand the tensor core effectively sees is hi + lo
        const float e0 = f0 + __bfloat162float(lo.x), e1 = f1 + __bfloat162float(lo.y);
        ss += e0 * e0 + e1 * e1;
        sum += e0 + e1;
        *reinterpret_cast<__nv_bfloat162*>(out + c) = hi;
        *reinterpret_cast<__nv_bfloat162*>(out + d + c) = kSide == 0 ? hi : lo;
        *reinterpret_cast<__nv_bfloat162*>(out + 2 * d + c) = kSide == 0 ? lo : hi;
      } else {
        ss += f0 * f0 + f1 * f1;
        sum += f0 + f1;
        *reinterpret_cast<__nv_bfloat162*>(out + c) = hi;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ss += __shfl_xor_sync(0xffffffffu, ss, o);
      sum += __shfl_xor_sync(0xffffffffu, sum, o);
    }
    if (lane == 0) {
      const float r = 1.f / fmaxf(sqrtf(ss), 1e-12f);
      const size_t gp = size_t(g) * P + prow;
      if (kSide == 0) {
        rinv[gp] = r;
        if (pad_mode == GADM_PAD_MINUS_ONE) {
          pad_sim[gp] = -sum * r * rsqrtf(float(d));
        } else if (pad_mode == GADM_PAD_E0) {
          const float v0 = tile[pl];
          float e0 = __bfloat162float(__float2bfloat16_rn(v0));
          if (x3) e0 += __bfloat162float(__float2bfloat16_rn(v0 - e0));
          pad_sim[gp] = e0 * r;
        }
      } else {
        aux_scale[gp] = r;   // column scale of the matcher epilogue
      }
    }
  }
  if (kSide == 1) {
    // model coordinates: fp32 copy (Kabsch) and the x / y / z planes of the SOFT epilogue
    for (int e = threadIdx.x; e < 3 * PTS; e += blockDim.x) {
      const int pl = e / 3, c = e - 3 * pl, p = p0 + pl;
      if (p >= P) continue;
      const float v = xyz ? xyz[(size_t(g) * P + p0) * 3 + e] : 0.f;
      aux_xyz[(size_t(g) * P + p0) * 3 + e] = v;
      aux_planes[c * plane + size_t(g) * P + p] = v;
    }
  }
}

// evaluator.py:78,82 in one pass: mask = (argmax over the two seg channels == 1)
__global__ void __launch_bounds__(256)
seg_mask_kernel(const float* __restrict__ seg, int N, uint8_t* __restrict__ mask) {
  const int b = blockIdx.y;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float* s = seg + size_t(b) * 2 * N;
  mask[size_t(b) * N + n] = s[N + n] > s[n] ? 1 : 0;
}

// One CTA per frame: fp64 accumulation of n, sum A, sum B, sum A B^T over matched rows.
__global__ void __launch_bounds__(256)
kabsch_moments_kernel(const int64_t* __restrict__ idx, const uint8_t* __restrict__ mask,
                      const float* __restrict__ weight, const float* __restrict__ cloud, const float* __restrict__ aux,
                      const int32_t* __restrict__ obj_id, int B, int N, int M, int n_obj, double* __restrict__ out) {
  const int b = blockIdx.x;
  const int obj = obj_id ? min(max(obj_id[b], 0), n_obj - 1) : (n_obj == B ? b : 0);   // clamped: see frame_object
  const float* tab = aux_xyz(aux, n_obj, M) + size_t(obj) * M * 3;
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.0;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const size_t gn = size_t(b) * N + n;
    const int64_t j = idx[gn];
    if (j < 0 || j >= M) continue;
    if (mask && !mask[gn]) continue;
    const float* e = tab + j * 3;
    float4 a;
    a.x = e[0]; a.y = e[1]; a.z = e[2];
    const float bx = cloud[gn * 3 + 0], by = cloud[gn * 3 + 1], bz = cloud[gn * 3 + 2];
    // weighted Procrustes: every pair counts with its weight (the matcher's softmax weight); unweighted: w = 1
    const double w = weight ? double(weight[gn]) : 1.0;
    const double wx = w * a.x, wy = w * a.y, wz = w * a.z;
    acc[0] += w;
    acc[1] += wx; acc[2] += wy; acc[3] += wz;
    acc[4] += w * bx; acc[5] += w * by; acc[6] += w * bz;
    acc[7] += wx * bx;  acc[8] += wx * by;  acc[9] += wx * bz;
    acc[10] += wy * bx; acc[11] += wy * by; acc[12] += wy * bz;
    acc[13] += wz * bx; acc[14] += wz * by; acc[15] += wz * bz;
  }
  __shared__ double red[8][16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    double v = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];  // fixed order: deterministic
    out[size_t(b) * 16 + threadIdx.x] = v;
  }
}

}  // namespace

// evaluator.py:82-88 (cls_msk -> rgbd_features[cls_msk]) as a map: one CTA per frame scans the mask; pos[b, n] = rank of
// point n among the selected points of its frame (-1: not selected), row_map[b, j] = the point that has rank j,
// n_sel[b] = how many.  Order is preserved, so rows in compacted order are exactly the reference's selected rows.
__global__ void __launch_bounds__(1024)
compact_rows_kernel(const uint8_t* __restrict__ mask, int N, int32_t* __restrict__ pos, int32_t* __restrict__ row_map,
                    int32_t* __restrict__ n_sel) {
  __shared__ int warp_sum[32];
  __shared__ int carry;
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int n0 = 0; n0 < N; n0 += 1024) {
    const int n = n0 + threadIdx.x;
    const int flag = n < N && mask[size_t(b) * N + n] != 0;
    const unsigned ballot = __ballot_sync(0xffffffffu, flag);
    const int in_warp = __popc(ballot & ((1u << lane) - 1));
    if (lane == 0) warp_sum[warp] = __popc(ballot);
    __syncthreads();
    int before = carry;
    for (int w = 0; w < warp; ++w) before += warp_sum[w];
    if (n < N) {
      const int r = before + in_warp;
      pos[size_t(b) * N + n] = flag ? r : -1;
      if (flag) row_map[size_t(b) * N + r] = n;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = carry;
      for (int w = 0; w < 32; ++w) t += warp_sum[w];
      carry = t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) n_sel[b] = carry;
}

int compact_rows_launch(const uint8_t* mask, int B, int N, int32_t* pos, int32_t* row_map, int32_t* n_sel,
                        cudaStream_t stream) {
  compact_rows_kernel<<<B, 1024, 0, stream>>>(mask, N, pos, row_map, n_sel);
  return check_launch();
}

// The common case of prep_kernel -- one bf16 operand copy (BF16: no split, no pre-normalisation), d % 64 == 0,
// P % 4 == 0 -- with a quarter of its instructions (it was bound by instruction issue, not by HBM): 16-byte loads
// along the points (four points of one channel), and in the point phase EIGHT threads per point, each converting and
// storing 8 consecutive channels with one 16-byte store per pass and reducing the norm over its own channels before
// three shuffles.  Same outputs as prep_kernel; the norm is summed in a different order (last-bit differences in rinv).
template <int kSide, typename TIn>
__global__ void __launch_bounds__(256)
prep_fast_kernel(const TIn* __restrict__ src, const float* __restrict__ xyz, int d, int P, int pad_mode,
                 __nv_bfloat16* __restrict__ dst, float* __restrict__ rinv, float* __restrict__ pad_sim,
                 float* __restrict__ aux_scale, float* __restrict__ aux_xyz, float* __restrict__ aux_planes, size_t plane,
                 const int32_t* __restrict__ pos) {
  extern __shared__ float tile[];  // [d][PTS + 1]
  const int g = blockIdx.y;
  const int p0 = blockIdx.x * PTS;
  const TIn* s = src + size_t(g) * d * P;
  // ---- load: thread -> (channel, group of 4 points); PTS / 4 = 8 groups per channel row
  for (int e = threadIdx.x; e < d * (PTS / 4); e += 256) {
    const int c = e >> 3, q = e & 7, p = p0 + q * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (p < P) {            // P % 4 == 0: the four points are valid or invalid together
      if (sizeof(TIn) == 4) {
        const float4 t = *reinterpret_cast<const float4*>(s + size_t(c) * P + p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
      } else {
        const uint2 t = *reinterpret_cast<const uint2*>(s + size_t(c) * P + p);
        v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
        v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
      }
    }
    float* row = tile + c * (PTS + 1) + q * 4;
    row[0] = v[0]; row[1] = v[1]; row[2] = v[2]; row[3] = v[3];
  }
  __syncthreads();

  // ---- points: thread -> (point, sub); sub handles the channel groups sub, sub + 8, ... of 8 channels each
  const int pl = threadIdx.x >> 3, sub = threadIdx.x & 7;
  const int p = p0 + pl;
  const bool live = p < P;
  const int prow = live ? (pos ? pos[size_t(g) * P + p] : p) : -1;   // row compaction: the row this point becomes
  float ss = 0.f, sum = 0.f;
  if (prow >= 0) {
    __nv_bfloat16* out = dst + (size_t(g) * P + prow) * d;
    for (int cg = sub; cg < d / 8; cg += 8) {
      float f[8];
      uint32_t w[4];
#pragma unroll
      for (int u = 0; u < 8; u += 2) {
        const float v0 = tile[(cg * 8 + u) * (PTS + 1) + pl], v1 = tile[(cg * 8 + u + 1) * (PTS + 1) + pl];
        const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
        w[u >> 1] = *reinterpret_cast<const uint32_t*>(&h);
        f[u] = __uint_as_float(w[u >> 1] << 16);
        f[u + 1] = __uint_as_float(w[u >> 1] & 0xffff0000u);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) { ss = fmaf(f[u], f[u], ss); sum += f[u]; }
      *reinterpret_cast<uint4*>(out + cg * 8) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) {
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
  }
  if (sub == 0 && prow >= 0) {
    const float r = 1.f / fmaxf(sqrtf(ss), 1e-12f);
    const size_t gp = size_t(g) * P + prow;
    if (kSide == 0) {
      rinv[gp] = r;
      if (pad_mode == GADM_PAD_MINUS_ONE) pad_sim[gp] = -sum * r * rsqrtf(float(d));
      else if (pad_mode == GADM_PAD_E0) pad_sim[gp] = __bfloat162float(__float2bfloat16_rn(tile[pl])) * r;
    } else {
      aux_scale[gp] = r;
    }
  }
  if (kSide == 1) {
    for (int e = threadIdx.x; e < 3 * PTS; e += blockDim.x) {
      const int pl2 = e / 3, c = e - 3 * pl2, p2 = p0 + pl2;
      if (p2 >= P) continue;
      const float v = xyz ? xyz[(size_t(g) * P + p0) * 3 + e] : 0.f;
      aux_xyz[(size_t(g) * P + p0) * 3 + e] = v;
      aux_planes[c * plane + size_t(g) * P + p2] = v;
    }
  }
}

static bool prep_fast_ok(const void* src, const void* dst, int d, int P, int elem_bytes) {
  return d % 64 == 0 && P % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 &&
         (reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (size_t(P) * elem_bytes) % 16 == 0;
}

int prep_rows_launch(const void* feat, int feat_bf16, const int32_t* pos, int B, int d, int N, int operand_mode,
                     int pad_mode, void* rows, float* rinv, float* pad_sim, cudaStream_t stream) {
  dim3 grid((N + PTS - 1) / PTS, B);
  const size_t smem = size_t(d) * (PTS + 1) * sizeof(float);
  if (operand_mode != GADM_OPERAND_BF16X3 && prep_fast_ok(feat, rows, d, N, feat_bf16 ? 2 : 4)) {
    if (feat_bf16)
      prep_fast_kernel<0, __nv_bfloat16><<<grid, 256, smem, stream>>>(
          static_cast<const __nv_bfloat16*>(feat), nullptr, d, N, pad_mode, static_cast<__nv_bfloat16*>(rows), rinv,
          pad_sim, nullptr, nullptr, nullptr, 0, pos);
    else
      prep_fast_kernel<0, float><<<grid, 256, smem, stream>>>(
          static_cast<const float*>(feat), nullptr, d, N, pad_mode, static_cast<__nv_bfloat16*>(rows), rinv, pad_sim,
          nullptr, nullptr, nullptr, 0, pos);
    return check_launch();
  }
  if (feat_bf16)
    prep_kernel<0, __nv_bfloat16><<<grid, 256, smem, stream>>>(
        static_cast<const __nv_bfloat16*>(feat), nullptr, d, N, operand_mode == GADM_OPERAND_BF16X3, 0, pad_mode,
        static_cast<__nv_bfloat16*>(rows), rinv, pad_sim, nullptr, nullptr, nullptr, 0, pos);
  else
    prep_kernel<0, float><<<grid, 256, smem, stream>>>(
        static_cast<const float*>(feat), nullptr, d, N, operand_mode == GADM_OPERAND_BF16X3, 0, pad_mode,
        static_cast<__nv_bfloat16*>(rows), rinv, pad_sim, nullptr, nullptr, nullptr, 0, pos);
  return check_launch();
}

// idx (int64 -> int32), max_sim, weight, soft_xyz -> one [n, 6] record of 32-bit words per scene point: a single
// contiguous device-to-host copy carries every matcher output of a batch
__global__ void __launch_bounds__(256)
pack_outputs_kernel(const int64_t* __restrict__ idx, const float* __restrict__ max_sim, const float* __restrict__ weight,
                    const float* __restrict__ soft_xyz, size_t n, int32_t* __restrict__ out) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int32_t* o = out + i * 6;
  o[0] = int32_t(idx[i]);
  o[1] = __float_as_int(max_sim[i]);
  o[2] = weight ? __float_as_int(weight[i]) : 0;
  o[3] = soft_xyz ? __float_as_int(soft_xyz[i * 3 + 0]) : 0;
  o[4] = soft_xyz ? __float_as_int(soft_xyz[i * 3 + 1]) : 0;
  o[5] = soft_xyz ? __float_as_int(soft_xyz[i * 3 + 2]) : 0;
}

int pack_outputs_launch(const int64_t* idx, const float* max_sim, const float* weight, const float* soft_xyz, size_t n,
                        int32_t* out, cudaStream_t stream) {
  pack_outputs_kernel<<<unsigned((n + 255) / 256), 256, 0, stream>>>(idx, max_sim, weight, soft_xyz, n, out);
  return check_launch();
}

// int32 indices -> uint16 (two per thread, one 32-bit store): halves the device-to-host bytes of the kNN pyramid when
// every support cloud has fewer than 65536 points
__global__ void __launch_bounds__(256)
pack_u16_kernel(const int32_t* __restrict__ idx, size_t n, uint16_t* __restrict__ out) {
  const size_t i = (size_t(blockIdx.x) * blockDim.x + threadIdx.x) * 2;
  if (i + 1 < n) {
    const int2 v = *reinterpret_cast<const int2*>(idx + i);
    *reinterpret_cast<uint32_t*>(out + i) = (uint32_t(v.x) & 0xFFFFu) | (uint32_t(v.y) << 16);
  } else if (i < n) {
    out[i] = uint16_t(idx[i]);
  }
}

int pack_u16_launch(const int32_t* idx, size_t n, uint16_t* out, cudaStream_t stream) {
  pack_u16_kernel<<<unsigned((n / 2 + 256) / 256), 256, 0, stream>>>(idx, n, out);
  return check_launch();
}

int prep_model_launch(const float* mesh, const float* model_xyz, int n_obj, int d, int M, int operand_mode, void* cols,
                      float* aux, cudaStream_t stream) {
  dim3 grid((M + PTS - 1) / PTS, n_obj);
  const size_t smem = size_t(d) * (PTS + 1) * sizeof(float);
  const size_t plane = size_t(n_obj) * M;
  float* a_xyz = aux + plane;
  float* a_planes = aux + plane * 4;
  if (operand_mode == GADM_OPERAND_BF16 && prep_fast_ok(mesh, cols, d, M, 4)) {
    prep_fast_kernel<1, float><<<grid, 256, smem, stream>>>(mesh, model_xyz, d, M, 0, static_cast<__nv_bfloat16*>(cols),
                                                            nullptr, nullptr, aux, a_xyz, a_planes, plane, nullptr);
    return check_launch();
  }
  prep_kernel<1, float><<<grid, 256, smem, stream>>>(mesh, model_xyz, d, M, operand_mode == GADM_OPERAND_BF16X3,
                                              operand_mode == GADM_OPERAND_BF16N, 0,
                                              static_cast<__nv_bfloat16*>(cols), nullptr, nullptr, aux, a_xyz, a_planes, plane, nullptr);
  return check_launch();
}

int seg_mask_launch(const float* seg, int B, int N, uint8_t* mask, cudaStream_t stream) {
  if (B > 65535) return GADM_ERR_UNSUPPORTED;
  seg_mask_kernel<<<dim3((N + 255) / 256, B), 256, 0, stream>>>(seg, N, mask);
  return check_launch();
}

int kabsch_moments_launch(const int64_t* idx, const uint8_t* mask, const float* weight, const float* cloud,
                          const float* aux, const int32_t* obj_id, int B, int N, int M, int n_obj, double* out,
                          cudaStream_t stream) {
  kabsch_moments_kernel<<<B, 256, 0, stream>>>(idx, mask, weight, cloud, aux, obj_id, B, N, M, n_obj, out);
  return check_launch();
}

// best_fit_transform (utils/pvn3d_eval_utils_kpls.py:43-76; torch twin utils/basic_utils.py:848-880) from the moments
// of the matched pairs, one thread per frame, fp64:
//   H = sum(w A B^T) - W cA cB^T,   H = U S V^T (one-sided Jacobi, singular values sorted descending),
//   R = V U^T, reflection fix (:67-69): det R < 0 -> last column of V negated,   t = cB - R cA.
// Frames the reference answers with its sentinel (evaluator.py:69-72, :83, :96: not detected, <= 1 or fewer than
// min_pts matched rows) get identity and t_z = -1000.  `count` = number of pairs (moments of an unweighted pass; null:
// the weight sum is taken as the count).
__global__ void __launch_bounds__(128)
kabsch_pose_kernel(const double* __restrict__ mom, const double* __restrict__ count, const uint8_t* __restrict__ det,
                   int B, int min_pts, float* __restrict__ poses) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double* m = mom + size_t(b) * 16;
  float* T = poses + size_t(b) * 12;
  const double W = m[0];
  const double npairs = count ? count[size_t(b) * 16] : W;
  if ((det && !det[b]) || npairs <= 1.0 || npairs < double(min_pts) || !(W > 0.0)) {
    for (int i = 0; i < 12; ++i) T[i] = 0.f;
    T[0] = T[5] = T[10] = 1.f;
    T[11] = -1000.f;
    return;
  }
  double ca[3], cb[3], G[3][3], V[3][3];
  for (int i = 0; i < 3; ++i) { ca[i] = m[1 + i] / W; cb[i] = m[4 + i] / W; }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) { G[i][j] = m[7 + 3 * i + j] - W * ca[i] * cb[j]; V[i][j] = i == j ? 1.0 : 0.0; }
  // one-sided Jacobi: rotate column pairs of G (and V) until the columns of G are orthogonal: G = U S, H = U S V^T
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double al = 0, be = 0, ga = 0;
        for (int i = 0; i < 3; ++i) { al += G[i][p] * G[i][p]; be += G[i][q] * G[i][q]; ga += G[i][p] * G[i][q]; }
        off = fmax(off, fabs(ga) / fmax(sqrt(al * be), 1e-300));
        if (fabs(ga) <= 1e-300 || fabs(ga) <= 1e-17 * sqrt(al * be)) continue;
        const double zeta = (be - al) / (2.0 * ga);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
        for (int i = 0; i < 3; ++i) {
          const double gp = G[i][p], gq = G[i][q];
          G[i][p] = c * gp - s * gq; G[i][q] = s * gp + c * gq;
          const double vp = V[i][p], vq = V[i][q];
          V[i][p] = c * vp - s * vq; V[i][q] = s * vp + c * vq;
        }
      }
    if (off < 1e-15) break;
  }
  double sig[3];
  int ord[3] = {0, 1, 2};
  for (int j = 0; j < 3; ++j) sig[j] = sqrt(G[0][j] * G[0][j] + G[1][j] * G[1][j] + G[2][j] * G[2][j]);
  for (int a = 0; a < 2; ++a)                         // descending, like numpy's SVD (the fix below flips the SMALLEST)
    for (int c2 = a + 1; c2 < 3; ++c2)
      if (sig[ord[c2]] > sig[ord[a]]) { const int tmp = ord[a]; ord[a] = ord[c2]; ord[c2] = tmp; }
  double U[3][3], Vs[3][3];
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) { Vs[i][j] = V[i][ord[j]]; U[i][j] = sig[ord[j]] > 0 ? G[i][ord[j]] / sig[ord[j]] : 0.0; }
  // rank-deficient H: complete U to a rotation-or-reflection basis (the fix below settles the sign)
  if (!(sig[ord[1]] > 1e-12 * sig[ord[0]])) {
    const int k = fabs(U[0][0]) < 0.9 ? 0 : 1;        // any vector not parallel to u0
    double e[3] = {0, 0, 0}; e[k] = 1.0;
    const double d0 = U[k][0];
    double nn = 0;
    for (int i = 0; i < 3; ++i) { U[i][1] = e[i] - d0 * U[i][0]; nn += U[i][1] * U[i][1]; }
    for (int i = 0; i < 3; ++i) U[i][1] /= sqrt(nn);
  }
  if (!(sig[ord[2]] > 1e-12 * sig[ord[0]])) {
    U[0][2] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
    U[1][2] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
    U[2][2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
  }
  double R[3][3];
  auto compose = [&]() {
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) R[i][j] = Vs[i][0] * U[j][0] + Vs[i][1] * U[j][1] + Vs[i][2] * U[j][2];
  };
  compose();
  const double detR = R[0][0] * (R[1][1] * R[2][2] - R[1][2] * R[2][1]) - R[0][1] * (R[1][0] * R[2][2] - R[1][2] * R[2][0]) +
                      R[0][2] * (R[1][0] * R[2][1] - R[1][1] * R[2][0]);
  if (detR < 0) {
    for (int i = 0; i < 3; ++i) Vs[i][2] = -Vs[i][2];
    compose();
  }
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) T[i * 4 + j] = float(R[i][j]);
    T[i * 4 + 3] = float(cb[i] - (R[i][0] * ca[0] + R[i][1] * ca[1] + R[i][2] * ca[2]));
  }
}

int kabsch_pose_launch(const double* mom, const double* count, const uint8_t* det, int B, int min_pts, float* poses,
                       cudaStream_t stream) {
  kabsch_pose_kernel<<<(B + 127) / 128, 128, 0, stream>>>(mom, count, det, B, min_pts, poses);
  return check_launch();
}

}  // namespace gadm
