// Index-driven gathers that consume the kNN output (HBM-bound, one pass, no intermediates):
//   graph_feature     models/dgcnn.py:41-54      out = cat(x[:, :, idx] - x, x)  ->  [B, 2C, N, k]
//   group_fwd / bwd   lib/pointops/functions/pointops.py:149-178 (Grouping)
//   gather_neighbour  models/RandLA/RandLANet.py:729-738
// Threads map to the contiguous output axis so that every store is coalesced; the gathered reads hit rows of
// N floats that stay in L1/L2.
#include "gadm_internal.h"

namespace gadm {

namespace {

// thread = one (n, j) pair of one batch item; loops over channels
__global__ void __launch_bounds__(256)
graph_feature_kernel(const float* __restrict__ x, const int64_t* __restrict__ idx, int C, int N, int k,
                     float* __restrict__ out) {
  const int b = blockIdx.y;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // n * k + j
  const long long NK = (long long)N * k;
  if (e >= NK) return;
  const int n = int(e / k);
  const int nb = int(idx[(size_t)b * NK + e]);
  const float* xb = x + (size_t)b * C * N;
  float* ob = out + (size_t)b * 2 * C * NK + e;
#pragma unroll 4
  for (int c = 0; c < C; ++c) {
    const float ctr = xb[(size_t)c * N + n];
    const float nbr = xb[(size_t)c * N + nb];
    __stcs(ob + (size_t)c * NK, nbr - ctr);        // streaming stores: the 2.7 GB output is never re-read here
    __stcs(ob + (size_t)(C + c) * NK, ctr);
  }
}

// [B, C, N] -> [B, N, C] (32x32 shared-memory tiles, both sides coalesced)
__global__ void __launch_bounds__(256)
transpose_cn_kernel(const float* __restrict__ x, int C, int N, float* __restrict__ xt) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* xb = x + (size_t)b * C * N;
  float* tb = xt + (size_t)b * C * N;
  for (int r = ty; r < 32; r += 8)
    tile[r][tx] = (c0 + r < C && n0 + tx < N) ? xb[(size_t)(c0 + r) * N + n0 + tx] : 0.f;
  __syncthreads();
  for (int r = ty; r < 32; r += 8)
    if (n0 + r < N && c0 + tx < C) tb[(size_t)(n0 + r) * C + c0 + tx] = tile[tx][r];
}

// thread = one (n, j) pair; reads the neighbour's and the centre's channel vectors from the point-major copy
// (64 contiguous bytes at a time: every fetched sector is fully used -- the direct gather uses 4 of every 32
// bytes), writes channel-major with coalesced streaming stores.
__global__ void __launch_bounds__(256)
graph_feature_t_kernel(const float* __restrict__ xt, const int64_t* __restrict__ idx, int C, int N, int k,
                       float* __restrict__ out) {
  const int b = blockIdx.y;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // n * k + j
  const long long NK = (long long)N * k;
  if (e >= NK) return;
  const int n = int(e / k);
  const int nb = int(idx[(size_t)b * NK + e]);
  const float4* pn = reinterpret_cast<const float4*>(xt + ((size_t)b * N + nb) * C);
  const float4* pc = reinterpret_cast<const float4*>(xt + ((size_t)b * N + n) * C);
  float* ob = out + (size_t)b * 2 * C * NK + e;
  for (int c4 = 0; c4 < C / 4; c4 += 4) {
    float4 vn[4], vc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (c4 + u < C / 4) { vn[u] = __ldg(pn + c4 + u); vc[u] = __ldg(pc + c4 + u); }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (c4 + u < C / 4) {
        const int c = (c4 + u) * 4;
        __stcs(ob + (size_t)(c + 0) * NK, vn[u].x - vc[u].x);
        __stcs(ob + (size_t)(c + 1) * NK, vn[u].y - vc[u].y);
        __stcs(ob + (size_t)(c + 2) * NK, vn[u].z - vc[u].z);
        __stcs(ob + (size_t)(c + 3) * NK, vn[u].w - vc[u].w);
        __stcs(ob + (size_t)(C + c + 0) * NK, vc[u].x);
        __stcs(ob + (size_t)(C + c + 1) * NK, vc[u].y);
        __stcs(ob + (size_t)(C + c + 2) * NK, vc[u].z);
        __stcs(ob + (size_t)(C + c + 3) * NK, vc[u].w);
      }
  }
}

// thread = one (m, s) pair; loops over channels
__global__ void __launch_bounds__(256)
group_fwd_kernel(const float* __restrict__ f, const int32_t* __restrict__ idx, int c, int n, int ms,
                 float* __restrict__ out) {
  const int b = blockIdx.y;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ms) return;
  const int j = idx[(size_t)b * ms + e];
  const float* fb = f + (size_t)b * c * n;
  float* ob = out + (size_t)b * c * ms + e;
  for (int ch = 0; ch < c; ++ch) ob[(size_t)ch * ms] = fb[(size_t)ch * n + j];
}

__global__ void __launch_bounds__(256)
group_bwd_kernel(const float* __restrict__ go, const int32_t* __restrict__ idx, int c, int n, int ms,
                 float* __restrict__ gf) {
  const int b = blockIdx.y;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ms) return;
  const int j = idx[(size_t)b * ms + e];
  const float* gb = go + (size_t)b * c * ms + e;
  float* fb = gf + (size_t)b * c * n;
  for (int ch = 0; ch < c; ++ch) atomicAdd(fb + (size_t)ch * n + j, gb[(size_t)ch * ms]);
}

// warp = one output row (b, m, k): C contiguous floats
__global__ void __launch_bounds__(256)
gather_neighbour_kernel(const float* __restrict__ pc, const int64_t* __restrict__ idx, int N, int C, long long rows,
                        long long rows_per_b, float* __restrict__ out) {
  const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const long long b = r / rows_per_b;
  const float* src = pc + ((size_t)b * N + (size_t)idx[r]) * C;
  float* dst = out + (size_t)r * C;
  for (int c = lane; c < C; c += 32) dst[c] = src[c];
}

// thread = one output point m of one batch item (consecutive lanes = consecutive m: coalesced stores along M);
// the K indices of the point are read once and reused for every channel of the thread's channel block.
template <int KMAX>
__global__ void __launch_bounds__(256)
gather_max_kernel(const float* __restrict__ f, const int64_t* __restrict__ idx, int C, int N, int M, int K,
                  int c_per_block, float* __restrict__ out) {
  const int b = blockIdx.z;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  int nb[KMAX];
  const int64_t* ip = idx + ((size_t)b * M + m) * K;
#pragma unroll
  for (int j = 0; j < KMAX; ++j) nb[j] = j < K ? int(ip[j]) : 0;
  const int c0 = blockIdx.y * c_per_block;
  const int c1 = min(C, c0 + c_per_block);
  for (int c = c0; c < c1; ++c) {
    const float* row = f + ((size_t)b * C + c) * N;
    float v = row[nb[0]];
#pragma unroll
    for (int j = 1; j < KMAX; ++j)
      if (j < K) v = fmaxf(v, row[nb[j]]);
    out[((size_t)b * C + c) * M + m] = v;
  }
}

// thread = one (n, k) pair: 10 floats = five float2 stores (40-byte rows are 8-byte aligned)
__global__ void __launch_bounds__(256)
relative_pos_encoding_kernel(const float* __restrict__ xyz, const int64_t* __restrict__ idx, int N, long long NK,
                             int K, float* __restrict__ out) {
  const int b = blockIdx.y;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // n * K + k
  if (e >= NK) return;
  const int n = int(e / K);
  const int q = int(idx[(size_t)b * NK + e]);
  const float* xb = xyz + (size_t)b * N * 3;
  const float px = xb[n * 3 + 0], py = xb[n * 3 + 1], pz = xb[n * 3 + 2];
  const float qx = xb[q * 3 + 0], qy = xb[q * 3 + 1], qz = xb[q * 3 + 2];
  const float dx = px - qx, dy = py - qy, dz = pz - qz;
  const float dis = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
  float2* o = reinterpret_cast<float2*>(out + ((size_t)b * NK + e) * 10);
  __stcs(o + 0, make_float2(dis, dx));
  __stcs(o + 1, make_float2(dy, dz));
  __stcs(o + 2, make_float2(px, py));
  __stcs(o + 3, make_float2(pz, qx));
  __stcs(o + 4, make_float2(qy, qz));
}

// ---- backward passes of the gathers (the reference's torch.gather / max / cat are differentiable in the features:
// models/dgcnn.py:41-54, models/RandLA/RandLANet.py:90-120, :729-738).  All three are scatter-adds with fp32 atomics.

// graph_feature: out[:, :C] = x[idx] - x, out[:, C:] = x.  thread = one (n, j) pair of one batch item, loops over channels:
//   gx[c, idx[n, j]] += g1[c, n, j];  gx[c, n] += g2[c, n, j] - g1[c, n, j]
__global__ void __launch_bounds__(256)
graph_feature_bwd_kernel(const float* __restrict__ go, const int64_t* __restrict__ idx, int C, int N, int k,
                         float* __restrict__ gx) {
  const int b = blockIdx.y;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // n * k + j
  const long long nk = (long long)N * k;
  if (e >= nk) return;
  const int n = int(e / k);
  const int q = int(idx[(size_t)b * nk + e]);
  const float* g1 = go + (size_t)b * 2 * C * nk + e;
  const float* g2 = g1 + (size_t)C * nk;
  float* gb = gx + (size_t)b * C * N;
  for (int c = 0; c < C; ++c) {
    const float a = g1[(size_t)c * nk], d = g2[(size_t)c * nk];
    atomicAdd(gb + (size_t)c * N + q, a);
    atomicAdd(gb + (size_t)c * N + n, d - a);
  }
}

// gather_neighbour: out[b, m, k, :] = pc[b, idx[b, m, k], :].  warp = one (b, m, k) row of C floats.
__global__ void __launch_bounds__(256)
gather_neighbour_bwd_kernel(const float* __restrict__ go, const int64_t* __restrict__ idx, int N, int C, long long rows,
                            long long rows_per_b, float* __restrict__ gpc) {
  const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const long long b = r / rows_per_b;
  float* dst = gpc + ((size_t)b * N + (size_t)idx[r]) * C;
  const float* src = go + (size_t)r * C;
  for (int c = lane; c < C; c += 32) atomicAdd(dst + c, src[c]);
}

// gather_max: out[b, c, m] = max_j f[b, c, idx[b, m, j]]: the gradient goes to the FIRST neighbour that attains the
// maximum (what torch.max(dim) differentiates to on the CPU).  thread = one (m) of one (b, channel block).
__global__ void __launch_bounds__(256)
gather_max_bwd_kernel(const float* __restrict__ f, const int64_t* __restrict__ idx, const float* __restrict__ go, int C,
                      int N, int M, int K, int c_per_block, float* __restrict__ gf) {
  const int b = blockIdx.z;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const int64_t* ip = idx + ((size_t)b * M + m) * K;
  const int c0 = blockIdx.y * c_per_block;
  const int c1 = min(C, c0 + c_per_block);
  for (int c = c0; c < c1; ++c) {
    const float* row = f + ((size_t)b * C + c) * N;
    int best = int(ip[0]);
    float v = row[best];
    for (int j = 1; j < K; ++j) {
      const int q = int(ip[j]);
      const float w = row[q];
      if (w > v) { v = w; best = q; }
    }
    atomicAdd(gf + ((size_t)b * C + c) * N + best, go[((size_t)b * C + c) * M + m]);
  }
}

}  // namespace

int graph_feature_bwd_launch(const float* grad_out, const int64_t* idx, int B, int C, int N, int k, float* grad_x,
                             cudaStream_t stream) {
  if (B > 65535) return GADM_ERR_UNSUPPORTED;
  cudaError_t e = cudaMemsetAsync(grad_x, 0, size_t(B) * C * N * sizeof(float), stream);
  if (e != cudaSuccess) return set_cuda_error(e);
  const long long nk = (long long)N * k;
  graph_feature_bwd_kernel<<<dim3(unsigned((nk + 255) / 256), B), 256, 0, stream>>>(grad_out, idx, C, N, k, grad_x);
  return check_launch();
}

int gather_neighbour_bwd_launch(const float* grad_out, const int64_t* idx, int B, int N, int C, int M, int K,
                                float* grad_pc, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(grad_pc, 0, size_t(B) * N * C * sizeof(float), stream);
  if (e != cudaSuccess) return set_cuda_error(e);
  const long long rows = (long long)B * M * K;
  gather_neighbour_bwd_kernel<<<unsigned((rows + 7) / 8), 256, 0, stream>>>(grad_out, idx, N, C, rows, (long long)M * K,
                                                                         grad_pc);
  return check_launch();
}

int gather_max_bwd_launch(const float* feature, const int64_t* idx, const float* grad_out, int B, int C, int N, int M,
                          int K, float* grad_feature, cudaStream_t stream) {
  if (B > 65535) return GADM_ERR_UNSUPPORTED;
  cudaError_t e = cudaMemsetAsync(grad_feature, 0, size_t(B) * C * N * sizeof(float), stream);
  if (e != cudaSuccess) return set_cuda_error(e);
  const int c_per_block = 8;
  gather_max_bwd_kernel<<<dim3((M + 255) / 256, (C + c_per_block - 1) / c_per_block, B), 256, 0, stream>>>(
      feature, idx, grad_out, C, N, M, K, c_per_block, grad_feature);
  return check_launch();
}

int gather_max_launch(const float* feature, const int64_t* idx, int B, int C, int N, int M, int K, float* out,
                      cudaStream_t stream) {
  if (B > 65535 || K > 32) return GADM_ERR_UNSUPPORTED;
  const int c_per_block = 8;
  dim3 grid((M + 255) / 256, (C + c_per_block - 1) / c_per_block, B);
  if (grid.y > 65535) return GADM_ERR_UNSUPPORTED;
  if (K == 1) gather_max_kernel<1><<<grid, 256, 0, stream>>>(feature, idx, C, N, M, K, c_per_block, out);
  else if (K <= 16) gather_max_kernel<16><<<grid, 256, 0, stream>>>(feature, idx, C, N, M, K, c_per_block, out);
  else gather_max_kernel<32><<<grid, 256, 0, stream>>>(feature, idx, C, N, M, K, c_per_block, out);
  return check_launch();
}

int relative_pos_encoding_launch(const float* xyz, const int64_t* idx, int B, int N, int K, float* out,
                                 cudaStream_t stream) {
  if (B > 65535) return GADM_ERR_UNSUPPORTED;
  const long long NK = (long long)N * K;
  dim3 grid((unsigned)((NK + 255) / 256), B);
  relative_pos_encoding_kernel<<<grid, 256, 0, stream>>>(xyz, idx, N, NK, K, out);
  return check_launch();
}

int graph_feature_launch(const float* x, const int64_t* idx, int B, int C, int N, int k, float* out,
                         void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (B > 65535) return GADM_ERR_UNSUPPORTED;
  const long long NK = (long long)N * k;
  if (C % 4 == 0 && workspace != nullptr && workspace_bytes >= size_t(B) * C * N * sizeof(float) &&
      (reinterpret_cast<uintptr_t>(workspace) & 15) == 0) {
    float* xt = static_cast<float*>(workspace);
    dim3 tg((N + 31) / 32, (C + 31) / 32, B);
    transpose_cn_kernel<<<tg, 256, 0, stream>>>(x, C, N, xt);
    dim3 grid((unsigned)((NK + 255) / 256), B);
    graph_feature_t_kernel<<<grid, 256, 0, stream>>>(xt, idx, C, N, k, out);
    return check_launch();
  }
  dim3 grid((unsigned)((NK + 255) / 256), B);
  graph_feature_kernel<<<grid, 256, 0, stream>>>(x, idx, C, N, k, out);
  return check_launch();
}

int group_fwd_launch(const float* features, const int32_t* idx, int b, int c, int n, int m, int s, float* out,
                     cudaStream_t stream) {
  if (b > 65535) return GADM_ERR_UNSUPPORTED;
  const int ms = m * s;
  dim3 grid((ms + 255) / 256, b);
  group_fwd_kernel<<<grid, 256, 0, stream>>>(features, idx, c, n, ms, out);
  return check_launch();
}

int group_bwd_launch(const float* grad_out, const int32_t* idx, int b, int c, int n, int m, int s,
                     float* grad_features, cudaStream_t stream) {
  if (b > 65535) return GADM_ERR_UNSUPPORTED;
  cudaError_t e = cudaMemsetAsync(grad_features, 0, size_t(b) * c * n * sizeof(float), stream);
  if (e != cudaSuccess) return set_cuda_error(e);
  const int ms = m * s;
  dim3 grid((ms + 255) / 256, b);
  group_bwd_kernel<<<grid, 256, 0, stream>>>(grad_out, idx, c, n, ms, grad_features);
  return check_launch();
}

int gather_neighbour_launch(const float* pc, const int64_t* idx, int B, int N, int C, int M, int K, float* out,
                            cudaStream_t stream) {
  const long long rows = (long long)B * M * K;
  gather_neighbour_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(pc, idx, N, C, rows, (long long)M * K, out);
  return check_launch();
}

}  // namespace gadm
