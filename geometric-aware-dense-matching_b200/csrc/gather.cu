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

}  // namespace

int graph_feature_launch(const float* x, const int64_t* idx, int B, int C, int N, int k, float* out,
                         cudaStream_t stream) {
  if (B > 65535) return GADM_ERR_UNSUPPORTED;
  const long long NK = (long long)N * k;
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
