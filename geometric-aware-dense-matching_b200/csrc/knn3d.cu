// Exact batched 3-D k-nearest-neighbour search for sm_100a.
//
// Replaces the reference's CPU KD-tree path: DataProcessing.knn_search (models/RandLA/helper_tool.py:161-170)
// -> knn_batch (nearest_neighbors/knn.pyx:71-109) -> cpp_knn_batch_omp (knn_.cxx:104-135, nanoflann 1.2.3), and
// the pointops knnquery contract (lib/pointops/functions/pointops.py:435-493).
//
// Semantics (bit-exact contract, see oracle/knn_oracle.c):
//   d2 = ((dx*dx) + (dy*dy)) + (dz*dz), dx = q.x - p.x, fp32, round-to-nearest, NO fma contraction
//   (nanoflann.hpp:343-346); result = the k lexicographically smallest (d2, index) pairs, ascending.
//
// Two algorithms, same result:
//   BRUTE  shared-memory-tiled all-pairs scan.  One warp owns 4 queries; every lane evaluates one candidate per
//          step for all 4; candidates that beat the current k-th entry are inserted with a warp-level sorted
//          insertion (lane l holds entry l of the list: ballot -> rank, shfl_up -> shift).
//   GRID   uniform-grid (counting-sort) acceleration: points binned into cells of side h, each query visits the
//          Chebyshev shells of cells around its own cell until the k-th best distance is provably smaller than
//          the distance to the unvisited region (conservative by a rounding slack), else falls back to a scan of
//          the whole cloud.  Selection is the same lexicographic insertion, so results are identical to BRUTE.
#include <float.h>

#include "gadm_internal.h"

namespace gadm {

namespace {

constexpr int MAX_JOBS = 24;      // per launch (kernel-parameter table); longer job lists are split
constexpr int QPW = 4;            // queries per warp (brute)
constexpr int WARPS = 8;
constexpr int QPB = QPW * WARPS;  // queries per CTA
constexpr int TS = 1024;          // support points per shared-memory tile (12 KB)
constexpr int GRID_MIN_SUPPORT = 2048;  // AUTO: smaller clouds are scanned by BRUTE
constexpr int R_MAX = 3;          // shells visited before the whole-cloud fallback
constexpr float PTS_PER_CELL = 6.f;

struct JobDev {
  long long support_off, query_off, out_off, support_bstride, query_bstride, out_bstride;
  long long ws_off, ws_item_bytes;  // GRID: byte offset of item 0 of this job's cloud, bytes per item
  int n_support, n_query, k, batch;
  int tile_begin, tiles_per_item;
  int n_cells_max, use_grid;
};
struct LaunchJobs {
  int n_jobs, total_tiles;
  JobDev jobs[MAX_JOBS];
};

struct GridHeader {  // 64 bytes at the start of each item's workspace
  float ox, oy, oz, h, inv_h, slack;
  int dx, dy, dz, n_cells;
  int pad[6];
};

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
__host__ __device__ inline size_t grid_item_bytes(int n_support, int n_cells_max) {
  // header | cell_start[n_cells_max+1] | cursor[n_cells_max] | sorted float4[n_support] | cell_of_point[n_support]
  return align_up(64 + size_t(n_cells_max + 1) * 4 + size_t(n_cells_max) * 4, 16) + size_t(n_support) * 16 +
         align_up(size_t(n_support) * 4, 16);
}
__device__ inline int* grid_cell_start(uint8_t* item) { return reinterpret_cast<int*>(item + 64); }
__device__ inline int* grid_cursor(uint8_t* item, int ncm) { return reinterpret_cast<int*>(item + 64) + ncm + 1; }
__device__ inline float4* grid_sorted(uint8_t* item, int ncm) {
  return reinterpret_cast<float4*>(item + align_up(64 + size_t(ncm + 1) * 4 + size_t(ncm) * 4, 16));
}
__device__ inline int* grid_cell_of(uint8_t* item, int ncm, int ns) {
  return reinterpret_cast<int*>(item + align_up(64 + size_t(ncm + 1) * 4 + size_t(ncm) * 4, 16) + size_t(ns) * 16);
}

// the reference metric, explicitly un-contracted
__device__ __forceinline__ float dist2_ref(float qx, float qy, float qz, float px, float py, float pz) {
  const float dx = __fsub_rn(qx, px), dy = __fsub_rn(qy, py), dz = __fsub_rn(qz, pz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

__device__ __forceinline__ bool lex_less(float d, int i, float kd, int ki) { return d < kd || (d == kd && i < ki); }

// Warp-level sorted list: lane l holds entry l.  Inserts (d, i), which the caller has checked beats entry k-1.
__device__ __forceinline__ void warp_insert(float& ld, int& li, float d, int i, int lane) {
  const bool before = lex_less(ld, li, d, i);  // my entry stays in front of the candidate (a prefix of lanes)
  const int pos = __popc(__ballot_sync(0xffffffffu, before));
  const float up_d = __shfl_up_sync(0xffffffffu, ld, 1);
  const int up_i = __shfl_up_sync(0xffffffffu, li, 1);
  if (lane > pos) { ld = up_d; li = up_i; }
  else if (lane == pos) { ld = d; li = i; }
}

// Offer one candidate per lane (valid lanes only) to the list of one query.
__device__ __forceinline__ void warp_offer(float& ld, int& li, float& kd, int& ki, float d, int i, bool valid, int k,
                                           int lane) {
  unsigned m = __ballot_sync(0xffffffffu, valid && lex_less(d, i, kd, ki));
  while (m) {
    const int src = __ffs(m) - 1;
    m &= m - 1;
    const float dc = __shfl_sync(0xffffffffu, d, src);
    const int ic = __shfl_sync(0xffffffffu, i, src);
    if (lex_less(dc, ic, kd, ki)) {  // warp-uniform; the bound may have tightened since the ballot
      warp_insert(ld, li, dc, ic, lane);
      kd = __shfl_sync(0xffffffffu, ld, k - 1);
      ki = __shfl_sync(0xffffffffu, li, k - 1);
    }
  }
}

__device__ __forceinline__ const JobDev& find_job(const LaunchJobs& L, int tile, int& item, int& qtile) {
  int j = 0;
  while (j + 1 < L.n_jobs && tile >= L.jobs[j + 1].tile_begin) ++j;
  const JobDev& job = L.jobs[j];
  const int local = tile - job.tile_begin;
  item = local / job.tiles_per_item;
  qtile = local - item * job.tiles_per_item;
  return job;
}

// --------------------------------------------------------------------------------------------- BRUTE
__global__ void __launch_bounds__(WARPS * 32)
knn_brute_kernel(const float* __restrict__ support, const float* __restrict__ query, int32_t* __restrict__ idx,
                 float* __restrict__ dist2, const __grid_constant__ LaunchJobs L) {
  __shared__ float sm[TS * 3];
  int item, qtile;
  const JobDev& job = find_job(L, blockIdx.x, item, qtile);
  if (job.use_grid) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = job.k;
  const float* S = support + (job.support_off + item * job.support_bstride) * 3;
  const float* Q = query + (job.query_off + item * job.query_bstride) * 3;
  const long long obase = job.out_off + item * job.out_bstride;

  const int q0 = qtile * QPB + warp * QPW;
  float qx[QPW], qy[QPW], qz[QPW], ld[QPW], kd[QPW];
  int li[QPW], ki[QPW];
#pragma unroll
  for (int t = 0; t < QPW; ++t) {
    const int qi = min(q0 + t, job.n_query - 1);
    qx[t] = Q[qi * 3 + 0]; qy[t] = Q[qi * 3 + 1]; qz[t] = Q[qi * 3 + 2];
    ld[t] = FLT_MAX; li[t] = INT_MAX; kd[t] = FLT_MAX; ki[t] = INT_MAX;
  }
  // FLT_MAX sentinels: a real candidate at d2 == FLT_MAX with any index still beats (FLT_MAX, INT_MAX)

  for (int s0 = 0; s0 < job.n_support; s0 += TS) {
    const int tn = min(TS, job.n_support - s0);
    __syncthreads();
    for (int e = threadIdx.x; e < tn * 3; e += blockDim.x) sm[e] = S[size_t(s0) * 3 + e];  // coalesced
    __syncthreads();
    for (int j0 = 0; j0 < tn; j0 += 32) {
      const int j = j0 + lane;
      const bool valid = j < tn;
      const int jj = valid ? j : tn - 1;
      const float px = sm[jj * 3 + 0], py = sm[jj * 3 + 1], pz = sm[jj * 3 + 2];  // stride 3: conflict-free
      const int gi = s0 + j;
#pragma unroll
      for (int t = 0; t < QPW; ++t) {
        const float d = dist2_ref(qx[t], qy[t], qz[t], px, py, pz);
        warp_offer(ld[t], li[t], kd[t], ki[t], d, gi, valid, k, lane);
      }
    }
  }
#pragma unroll
  for (int t = 0; t < QPW; ++t) {
    const int qi = q0 + t;
    if (qi < job.n_query && lane < k) {
      idx[obase + (long long)qi * k + lane] = li[t];
      if (dist2) dist2[obase + (long long)qi * k + lane] = ld[t];
    }
  }
}

// --------------------------------------------------------------------------------------------- GRID build
struct CloudDev {
  long long support_off, support_bstride, ws_off, ws_item_bytes;
  int n_support, batch, n_cells_max, item_begin;
};
struct LaunchClouds {
  int n_clouds, total_items;
  CloudDev clouds[MAX_JOBS];
};

__device__ __forceinline__ const CloudDev& find_cloud(const LaunchClouds& C, int gitem, int& item) {
  int c = 0;
  while (c + 1 < C.n_clouds && gitem >= C.clouds[c + 1].item_begin) ++c;
  item = gitem - C.clouds[c].item_begin;
  return C.clouds[c];
}

__device__ __forceinline__ int cell_coord(float p, float o, float inv_h, int dim) {
  const int c = int(floorf(__fmul_rn(__fsub_rn(p, o), inv_h)));
  return max(0, min(dim - 1, c));
}

// one CTA per cloud item: bounding box -> grid geometry; zero the counters
__global__ void __launch_bounds__(256)
grid_setup_kernel(const float* __restrict__ support, uint8_t* __restrict__ ws, const __grid_constant__ LaunchClouds C) {
  int item;
  const CloudDev& cl = find_cloud(C, blockIdx.x, item);
  const float* S = support + (cl.support_off + item * cl.support_bstride) * 3;
  uint8_t* base = ws + cl.ws_off + item * cl.ws_item_bytes;
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int i = threadIdx.x; i < cl.n_support; i += blockDim.x) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float v = S[size_t(i) * 3 + a];
      lo[a] = fminf(lo[a], v);
      hi[a] = fmaxf(hi[a], v);
    }
  }
  __shared__ float red[6][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
    if (lane == 0) { red[a][warp] = lo[a]; red[3 + a][warp] = hi[a]; }
  }
  __syncthreads();
  __shared__ GridHeader hdr;
  if (threadIdx.x == 0) {
    for (int a = 0; a < 3; ++a)
      for (int w = 1; w < 8; ++w) {
        red[a][0] = fminf(red[a][0], red[a][w]);
        red[3 + a][0] = fmaxf(red[3 + a][0], red[3 + a][w]);
      }
    float e[3] = {red[3][0] - red[0][0], red[4][0] - red[1][0], red[5][0] - red[2][0]};
    // two largest extents span the (assumed 2.5-D) surface
    float a = fmaxf(e[0], fmaxf(e[1], e[2]));
    float c = fminf(e[0], fminf(e[1], e[2]));
    float b = e[0] + e[1] + e[2] - a - c;
    float h = sqrtf(PTS_PER_CELL * fmaxf(a * b, 1e-30f) / float(cl.n_support));
    if (!(h > 0.f) || !isfinite(h)) h = 1.f;
    h = fmaxf(h, a * 1e-4f + 1e-30f);
    int dx, dy, dz;
    for (;;) {
      dx = int(e[0] / h) + 1; dy = int(e[1] / h) + 1; dz = int(e[2] / h) + 1;
      if ((long long)dx * dy * dz <= cl.n_cells_max) break;
      h *= 1.25f;
    }
    hdr.ox = red[0][0]; hdr.oy = red[1][0]; hdr.oz = red[2][0];
    hdr.h = h; hdr.inv_h = 1.f / h;
    const float scale = fmaxf(a, fmaxf(fabsf(red[0][0]), fmaxf(fabsf(red[1][0]), fabsf(red[2][0])))) + a;
    hdr.slack = scale * 3.8e-6f;  // 2^-18: >> fp32 rounding of (p - o) * inv_h, << h
    hdr.dx = dx; hdr.dy = dy; hdr.dz = dz; hdr.n_cells = dx * dy * dz;
    *reinterpret_cast<GridHeader*>(base) = hdr;
  }
  __syncthreads();
  int* cursor = grid_cursor(base, cl.n_cells_max);
  for (int i = threadIdx.x; i < hdr.n_cells; i += blockDim.x) cursor[i] = 0;
}

// grid over (cloud item, point chunk): histogram
__global__ void __launch_bounds__(256)
grid_count_kernel(const float* __restrict__ support, uint8_t* __restrict__ ws, const __grid_constant__ LaunchClouds C) {
  int item;
  const CloudDev& cl = find_cloud(C, blockIdx.y, item);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cl.n_support) return;
  const float* S = support + (cl.support_off + item * cl.support_bstride) * 3;
  uint8_t* base = ws + cl.ws_off + item * cl.ws_item_bytes;
  const GridHeader* g = reinterpret_cast<const GridHeader*>(base);
  const int cx = cell_coord(S[size_t(i) * 3 + 0], g->ox, g->inv_h, g->dx);
  const int cy = cell_coord(S[size_t(i) * 3 + 1], g->oy, g->inv_h, g->dy);
  const int cz = cell_coord(S[size_t(i) * 3 + 2], g->oz, g->inv_h, g->dz);
  const int cell = (cz * g->dy + cy) * g->dx + cx;
  grid_cell_of(base, cl.n_cells_max, cl.n_support)[i] = cell;
  atomicAdd(&grid_cursor(base, cl.n_cells_max)[cell], 1);
}

// one CTA per cloud item: exclusive scan of the histogram -> cell_start; cursor := cell_start
__global__ void __launch_bounds__(1024)
grid_scan_kernel(uint8_t* __restrict__ ws, const __grid_constant__ LaunchClouds C) {
  int item;
  const CloudDev& cl = find_cloud(C, blockIdx.x, item);
  uint8_t* base = ws + cl.ws_off + item * cl.ws_item_bytes;
  const int n = reinterpret_cast<const GridHeader*>(base)->n_cells;
  int* start = grid_cell_start(base);
  int* cursor = grid_cursor(base, cl.n_cells_max);
  __shared__ int warp_sums[32];
  __shared__ int carry;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int c0 = 0; c0 < n; c0 += 1024) {
    const int i = c0 + threadIdx.x;
    const int v = i < n ? cursor[i] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int w = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
      }
      warp_sums[lane] = w;  // inclusive
    }
    __syncthreads();
    const int excl = carry + (warp ? warp_sums[warp - 1] : 0) + x - v;
    if (i < n) { start[i] = excl; cursor[i] = excl; }
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) start[n] = carry;
}

__global__ void __launch_bounds__(256)
grid_scatter_kernel(const float* __restrict__ support, uint8_t* __restrict__ ws,
                    const __grid_constant__ LaunchClouds C) {
  int item;
  const CloudDev& cl = find_cloud(C, blockIdx.y, item);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cl.n_support) return;
  const float* S = support + (cl.support_off + item * cl.support_bstride) * 3;
  uint8_t* base = ws + cl.ws_off + item * cl.ws_item_bytes;
  const int cell = grid_cell_of(base, cl.n_cells_max, cl.n_support)[i];
  const int pos = atomicAdd(&grid_cursor(base, cl.n_cells_max)[cell], 1);
  grid_sorted(base, cl.n_cells_max)[pos] =
      make_float4(S[size_t(i) * 3 + 0], S[size_t(i) * 3 + 1], S[size_t(i) * 3 + 2], __int_as_float(i));
}

// --------------------------------------------------------------------------------------------- GRID query
// one warp per query, QPW queries per warp in sequence
__global__ void __launch_bounds__(WARPS * 32)
knn_grid_kernel(const float* __restrict__ query, int32_t* __restrict__ idx, float* __restrict__ dist2,
                uint8_t* __restrict__ ws, const __grid_constant__ LaunchJobs L) {
  int item, qtile;
  const JobDev& job = find_job(L, blockIdx.x, item, qtile);
  if (!job.use_grid) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = job.k;
  const float* Q = query + (job.query_off + item * job.query_bstride) * 3;
  const long long obase = job.out_off + item * job.out_bstride;
  uint8_t* base = ws + job.ws_off + item * job.ws_item_bytes;
  const GridHeader g = *reinterpret_cast<const GridHeader*>(base);
  const int* __restrict__ start = grid_cell_start(base);
  const float4* __restrict__ pts = grid_sorted(base, job.n_cells_max);

  for (int t = 0; t < QPW; ++t) {
    const int qi = qtile * QPB + warp * QPW + t;
    if (qi >= job.n_query) break;  // warp-uniform
    const float qx = Q[qi * 3 + 0], qy = Q[qi * 3 + 1], qz = Q[qi * 3 + 2];
    const int cx = cell_coord(qx, g.ox, g.inv_h, g.dx);
    const int cy = cell_coord(qy, g.oy, g.inv_h, g.dy);
    const int cz = cell_coord(qz, g.oz, g.inv_h, g.dz);
    float ld = FLT_MAX, kd = FLT_MAX;
    int li = INT_MAX, ki = INT_MAX;
    bool done = false;

    for (int rho = 0; rho <= R_MAX && !done; ++rho) {
      // range slots of shell rho: rows (dz, dy) in [-rho, rho]^2, two slots per row
      const int side = 2 * rho + 1;
      const int nslots = 2 * side * side;
      for (int s0 = 0; s0 < nslots; s0 += 32) {
        const int s = s0 + lane;
        int rb = 0, re = 0;
        if (s < nslots) {
          const int row = s >> 1, second = s & 1;
          const int dz = row / side - rho, dy = row % side - rho;
          const int z = cz + dz, y = cy + dy;
          if (z >= 0 && z < g.dz && y >= 0 && y < g.dy) {
            const bool rim = max(abs(dz), abs(dy)) == rho;
            int x_lo, x_hi;
            if (rim) { x_lo = cx - rho; x_hi = second ? x_lo - 1 : cx + rho; }
            else     { x_lo = second ? cx + rho : cx - rho; x_hi = x_lo; }
            if (rho == 0 && second) x_hi = x_lo - 1;
            x_lo = max(x_lo, 0); x_hi = min(x_hi, g.dx - 1);
            if (x_lo <= x_hi) {
              const int rowbase = (z * g.dy + y) * g.dx;
              rb = start[rowbase + x_lo];
              re = start[rowbase + x_hi + 1];
            }
          }
        }
        unsigned nonempty = __ballot_sync(0xffffffffu, re > rb);
        while (nonempty) {
          const int src = __ffs(nonempty) - 1;
          nonempty &= nonempty - 1;
          const int b0 = __shfl_sync(0xffffffffu, rb, src), e0 = __shfl_sync(0xffffffffu, re, src);
          for (int j0 = b0; j0 < e0; j0 += 32) {
            const int j = j0 + lane;
            const bool valid = j < e0;
            const float4 p = pts[valid ? j : e0 - 1];
            const float d = dist2_ref(qx, qy, qz, p.x, p.y, p.z);
            warp_offer(ld, li, kd, ki, d, __float_as_int(p.w), valid, k, lane);
          }
        }
      }
      // distance from q to the nearest face of the visited block behind which unvisited cells exist
      float bound = FLT_MAX;
      if (cx - rho > 0) bound = fminf(bound, qx - (g.ox + float(cx - rho) * g.h));
      if (cx + rho < g.dx - 1) bound = fminf(bound, (g.ox + float(cx + rho + 1) * g.h) - qx);
      if (cy - rho > 0) bound = fminf(bound, qy - (g.oy + float(cy - rho) * g.h));
      if (cy + rho < g.dy - 1) bound = fminf(bound, (g.oy + float(cy + rho + 1) * g.h) - qy);
      if (cz - rho > 0) bound = fminf(bound, qz - (g.oz + float(cz - rho) * g.h));
      if (cz + rho < g.dz - 1) bound = fminf(bound, (g.oz + float(cz + rho + 1) * g.h) - qz);
      if (bound == FLT_MAX) {
        done = true;  // the block covers the whole grid
      } else {
        bound -= g.slack;
        done = bound > 0.f && kd < bound * bound;  // strict: an unvisited point at exactly kd could win the index tie
      }
    }
    if (!done) {  // pathological query (far outside / sparse region): scan the whole cloud
      ld = FLT_MAX; li = INT_MAX; kd = FLT_MAX; ki = INT_MAX;
      for (int j0 = 0; j0 < job.n_support; j0 += 32) {
        const int j = j0 + lane;
        const bool valid = j < job.n_support;
        const float4 p = pts[valid ? j : job.n_support - 1];
        const float d = dist2_ref(qx, qy, qz, p.x, p.y, p.z);
        warp_offer(ld, li, kd, ki, d, __float_as_int(p.w), valid, k, lane);
      }
    }
    if (lane < k) {
      idx[obase + (long long)qi * k + lane] = li;
      if (dist2) dist2[obase + (long long)qi * k + lane] = ld;
    }
  }
}

int cells_for(int n_support) {
  long long c = 8LL * n_support;
  if (c < 4096) c = 4096;
  if (c > (1 << 18)) c = 1 << 18;
  return int(c);
}

bool job_uses_grid(const gadm_knn_job& j, int algo) {
  if (algo == GADM_KNN_BRUTE) return false;
  if (algo == GADM_KNN_GRID) return true;
  return j.n_support >= GRID_MIN_SUPPORT;
}

bool same_cloud(const gadm_knn_job& a, const gadm_knn_job& b) {
  return a.support_off == b.support_off && a.n_support == b.n_support && a.support_bstride == b.support_bstride &&
         a.batch == b.batch;
}

}  // namespace

int knn3d_configure() { return GADM_OK; }

size_t knn3d_workspace_bytes(const gadm_knn_job* jobs, int n_jobs, int algo) {
  size_t total = 0;
  for (int i = 0; i < n_jobs; ++i) {
    if (!job_uses_grid(jobs[i], algo)) continue;
    bool dup = false;
    for (int p = 0; p < i && !dup; ++p) dup = job_uses_grid(jobs[p], algo) && same_cloud(jobs[p], jobs[i]);
    if (dup) continue;
    total += grid_item_bytes(jobs[i].n_support, cells_for(jobs[i].n_support)) * size_t(jobs[i].batch);
  }
  return total;
}

int knn3d_launch(const float* support, const float* query, const gadm_knn_job* jobs, int n_jobs, int algo,
                 int32_t* idx, float* dist2, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  const size_t need = knn3d_workspace_bytes(jobs, n_jobs, algo);
  if (need > 0 && (!workspace || workspace_bytes < need)) return GADM_ERR_WORKSPACE;
  if (need > 0 && (reinterpret_cast<uintptr_t>(workspace) & 15)) return GADM_ERR_ALIGN;

  // workspace offsets per distinct cloud (deduplicated over the WHOLE job list, not per launch chunk)
  long long* ws_off = new long long[n_jobs];
  int* owner = new int[n_jobs];
  size_t off = 0;
  for (int i = 0; i < n_jobs; ++i) {
    ws_off[i] = -1; owner[i] = -1;
    if (!job_uses_grid(jobs[i], algo)) continue;
    for (int p = 0; p < i; ++p)
      if (owner[p] == p && same_cloud(jobs[p], jobs[i])) { owner[i] = p; break; }
    if (owner[i] < 0) {
      owner[i] = i;
      ws_off[i] = (long long)off;
      off += grid_item_bytes(jobs[i].n_support, cells_for(jobs[i].n_support)) * size_t(jobs[i].batch);
    } else {
      ws_off[i] = ws_off[owner[i]];
    }
  }

  int rc = GADM_OK;
  // ---- build the grids (distinct clouds), MAX_JOBS clouds per launch
  {
    LaunchClouds C;
    C.n_clouds = 0; C.total_items = 0;
    int max_pts = 0;
    auto flush = [&]() {
      if (C.n_clouds == 0) return;
      uint8_t* ws = static_cast<uint8_t*>(workspace);
      grid_setup_kernel<<<C.total_items, 256, 0, stream>>>(support, ws, C);
      dim3 gpts((max_pts + 255) / 256, C.total_items);
      grid_count_kernel<<<gpts, 256, 0, stream>>>(support, ws, C);
      grid_scan_kernel<<<C.total_items, 1024, 0, stream>>>(ws, C);
      grid_scatter_kernel<<<gpts, 256, 0, stream>>>(support, ws, C);
      C.n_clouds = 0; C.total_items = 0; max_pts = 0;
    };
    for (int i = 0; i < n_jobs; ++i) {
      if (owner[i] != i) continue;
      CloudDev& cl = C.clouds[C.n_clouds];
      cl.support_off = jobs[i].support_off; cl.support_bstride = jobs[i].support_bstride;
      cl.ws_off = ws_off[i];
      cl.n_cells_max = cells_for(jobs[i].n_support);
      cl.ws_item_bytes = (long long)grid_item_bytes(jobs[i].n_support, cl.n_cells_max);
      cl.n_support = jobs[i].n_support; cl.batch = jobs[i].batch;
      cl.item_begin = C.total_items;
      C.total_items += jobs[i].batch;
      if (jobs[i].n_support > max_pts) max_pts = jobs[i].n_support;
      if (++C.n_clouds == MAX_JOBS) flush();
    }
    flush();
    rc = check_launch();
  }

  // ---- queries, MAX_JOBS jobs per launch
  for (int j0 = 0; j0 < n_jobs && rc == GADM_OK; j0 += MAX_JOBS) {
    LaunchJobs L;
    L.n_jobs = (n_jobs - j0 < MAX_JOBS) ? n_jobs - j0 : MAX_JOBS;
    int tiles = 0;
    bool any_grid = false, any_brute = false;
    for (int i = 0; i < L.n_jobs; ++i) {
      const gadm_knn_job& s = jobs[j0 + i];
      JobDev& d = L.jobs[i];
      d.support_off = s.support_off; d.query_off = s.query_off; d.out_off = s.out_off;
      d.support_bstride = s.support_bstride; d.query_bstride = s.query_bstride; d.out_bstride = s.out_bstride;
      d.n_support = s.n_support; d.n_query = s.n_query; d.k = s.k; d.batch = s.batch;
      d.use_grid = job_uses_grid(s, algo) ? 1 : 0;
      d.n_cells_max = cells_for(s.n_support);
      d.ws_off = d.use_grid ? ws_off[j0 + i] : 0;
      d.ws_item_bytes = d.use_grid ? (long long)grid_item_bytes(s.n_support, d.n_cells_max) : 0;
      d.tile_begin = tiles;
      d.tiles_per_item = (s.n_query + QPB - 1) / QPB;
      tiles += d.tiles_per_item * s.batch;
      (d.use_grid ? any_grid : any_brute) = true;
    }
    L.total_tiles = tiles;
    if (any_brute) knn_brute_kernel<<<tiles, WARPS * 32, 0, stream>>>(support, query, idx, dist2, L);
    if (any_grid)
      knn_grid_kernel<<<tiles, WARPS * 32, 0, stream>>>(query, idx, dist2, static_cast<uint8_t*>(workspace), L);
    rc = check_launch();
  }
  delete[] ws_off;
  delete[] owner;
  return rc;
}

}  // namespace gadm
