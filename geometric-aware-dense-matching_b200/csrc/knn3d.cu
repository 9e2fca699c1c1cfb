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
#include <stdlib.h>
#include <string.h>

#include "gadm_internal.h"

namespace gadm {

namespace {

constexpr int MAX_JOBS = 24;      // per launch (kernel-parameter table); longer job lists are split
constexpr int QPW = 4;            // queries per warp (brute)
constexpr int WARPS = 8;
constexpr int QPB = QPW * WARPS;  // queries per CTA
#ifndef GADM_KNN_GRID_QPW
#define GADM_KNN_GRID_QPW 4
#endif
constexpr int GQPW = GADM_KNN_GRID_QPW;   // queries per warp of the grid kernel, one after the other (measured on the
constexpr int GQPB = GQPW * WARPS;        // 8-frame pyramid: 4 -> 0.405 ms, 8 -> 0.413 ms, 16 -> 0.442 ms)
constexpr int TS = 1024;          // support points per shared-memory tile (12 KB)
constexpr int R_MAX = 4;          // block radius (in cells) visited before the whole-cloud fallback
#ifndef GADM_KNN_GRID_E
#define GADM_KNN_GRID_E 4
#endif
constexpr int GRID_E = GADM_KNN_GRID_E;   // pending candidates per lane between extractions (grid kernel)
constexpr int BRUTE_E = 2;        // same, brute kernel (4 queries per warp: register budget)
constexpr int HALF_BLOCK_K = 4;   // grid queries with k <= this try the 2x2x2 half-cell block first

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
  float ox, oy, oz;      // bounding-box origin
  float hx, hy, hz;      // cell size per axis (the thinnest axis may get fewer, taller cells: 2.5-D surfaces)
  float ihx, ihy, ihz;   // 1 / cell size
  float slack;           // rounding slack of the cell assignment, subtracted from every face distance
  int dx, dy, dz, n_cells;
  int pad[2];
};
static_assert(sizeof(GridHeader) == 64, "GridHeader is the 64-byte workspace prefix");

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

// ---------------------------------------------------------------------------------------------------
// Warp-cooperative exact top-k selection (k <= 32) over a stream of (d2, index) candidates, one per lane per
// offer.  Keys are compared lexicographically on (bits(d2), index); d2 >= 0 so its IEEE bits are monotone.
//   * every lane keeps a short SORTED list of its own pending candidates (E entries, registers);
//   * a candidate is admitted only if it beats the current k-th best (kd, ki), known after the first extraction;
//   * when some lane's list is full (or at flush) the warp EXTRACTS the k smallest keys: k rounds of two
//     hardware warp reductions (redux.sync.min.u32 on the distance bits, then on the index among the lanes
//     that hold that distance); the winner pops its list head.  Rank r lands in lane r, which then holds it as
//     the single entry of its list, so an extraction is also the compaction step.
// Cost per query is ~16 issue slots per extracted rank instead of a serial ballot/shuffle insertion per
// admitted candidate (profiles/: the r1 kernels spent ~2400 issue slots per k=16 query on that path).
constexpr uint32_t D_EMPTY = 0xffffffffu;   // sorts after every real distance (inf = 0x7f800000)
constexpr int I_EMPTY = 0x7fffffff;

template <int E>
struct WarpSelect {
  uint32_t ed[E];
  int ei[E];
  uint32_t kd;   // k-th best so far (D_EMPTY / I_EMPTY until k candidates have been extracted)
  int ki;
  bool dirty;    // warp-uniform: something was admitted since the last extraction

  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int s = 0; s < E; ++s) { ed[s] = D_EMPTY; ei[s] = I_EMPTY; }
    kd = D_EMPTY; ki = I_EMPTY; dirty = false;
  }
  static __device__ __forceinline__ bool key_less(uint32_t d, int i, uint32_t d2, int i2) {
    return d < d2 || (d == d2 && i < i2);
  }
  // k smallest keys of everything pending -> lane r holds rank r (r < k); (kd, ki) = rank k-1
  __device__ __forceinline__ void extract(int k, int lane) {
    uint32_t od = D_EMPTY;
    int oi = I_EMPTY;
    for (int r = 0; r < k; ++r) {
      const uint32_t dmin = __reduce_min_sync(0xffffffffu, ed[0]);
      const uint32_t cand = ed[0] == dmin ? uint32_t(ei[0]) : uint32_t(I_EMPTY);
      const uint32_t imin = __reduce_min_sync(0xffffffffu, cand);
      if (lane == r) { od = dmin; oi = int(imin); }
      if (cand == imin) {  // the owner pops its head (all lanes pop an EMPTY head once the stream is exhausted)
#pragma unroll
        for (int s = 0; s + 1 < E; ++s) { ed[s] = ed[s + 1]; ei[s] = ei[s + 1]; }
        ed[E - 1] = D_EMPTY; ei[E - 1] = I_EMPTY;
      }
    }
    ed[0] = od; ei[0] = oi;
#pragma unroll
    for (int s = 1; s < E; ++s) { ed[s] = D_EMPTY; ei[s] = I_EMPTY; }
    kd = __shfl_sync(0xffffffffu, od, k - 1);
    ki = __shfl_sync(0xffffffffu, oi, k - 1);
    dirty = false;
  }
  // one candidate per lane (valid lanes only)
  __device__ __forceinline__ void offer(float d, int i, bool valid, int k, int lane) {
    const uint32_t db = __float_as_uint(d);
    const bool acc = valid && key_less(db, i, kd, ki);
    if (!__any_sync(0xffffffffu, acc)) return;
    // bubble the newcomer through the sorted list; rejected lanes carry the EMPTY key, which never swaps
    uint32_t cd = acc ? db : D_EMPTY;
    int ci = acc ? i : I_EMPTY;
#pragma unroll
    for (int s = 0; s < E; ++s) {
      const bool sw = key_less(cd, ci, ed[s], ei[s]);
      const uint32_t td = sw ? ed[s] : cd;
      const int ti = sw ? ei[s] : ci;
      ed[s] = sw ? cd : ed[s];
      ei[s] = sw ? ci : ei[s];
      cd = td; ci = ti;
    }
    dirty = true;
    if (__any_sync(0xffffffffu, ed[E - 1] != D_EMPTY)) extract(k, lane);
  }
  __device__ __forceinline__ void flush(int k, int lane) {
    if (dirty) extract(k, lane);
  }
};

// k == 1 (the 1-NN interpolation jobs: half of all queries of the pyramid): no lists, every lane keeps the best key
// it has seen, one pair of warp reductions at the end of a phase.  Same interface as WarpSelect.
struct WarpMin {
  uint32_t ed[1];
  int ei[1];
  uint32_t kd;
  int ki;
  __device__ __forceinline__ void reset() { ed[0] = D_EMPTY; ei[0] = I_EMPTY; kd = D_EMPTY; ki = I_EMPTY; }
  __device__ __forceinline__ void offer(float d, int i, bool valid, int, int) {
    const uint32_t db = __float_as_uint(d);
    if (valid && (db < ed[0] || (db == ed[0] && i < ei[0]))) { ed[0] = db; ei[0] = i; }
  }
  __device__ __forceinline__ void flush(int, int) {
    const uint32_t dmin = __reduce_min_sync(0xffffffffu, ed[0]);
    const uint32_t cand = ed[0] == dmin ? uint32_t(ei[0]) : uint32_t(I_EMPTY);
    const uint32_t imin = __reduce_min_sync(0xffffffffu, cand);
    kd = dmin; ki = int(imin);
    ed[0] = dmin; ei[0] = int(imin);     // every lane now holds the winner (lane 0 writes it out)
  }
};

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
#ifndef GADM_KNN_BRUTE_MINB
#define GADM_KNN_BRUTE_MINB 4
#endif
__global__ void __launch_bounds__(WARPS * 32, GADM_KNN_BRUTE_MINB)
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
  float qx[QPW], qy[QPW], qz[QPW];
  WarpSelect<BRUTE_E> sel[QPW];
#pragma unroll
  for (int t = 0; t < QPW; ++t) {
    const int qi = min(q0 + t, job.n_query - 1);
    qx[t] = Q[qi * 3 + 0]; qy[t] = Q[qi * 3 + 1]; qz[t] = Q[qi * 3 + 2];
    sel[t].reset();
  }

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
        sel[t].offer(d, gi, valid, k, lane);
      }
    }
  }
#pragma unroll
  for (int t = 0; t < QPW; ++t) {
    sel[t].flush(k, lane);
    const int qi = q0 + t;
    if (qi < job.n_query && lane < k) {
      idx[obase + (long long)qi * k + lane] = sel[t].ei[0];
      if (dist2) dist2[obase + (long long)qi * k + lane] = __uint_as_float(sel[t].ed[0]);
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
// Geometry: cells of side h = sqrt(ppc * a * b / n) along the two largest bounding-box extents a >= b (depth
// clouds are 2.5-D surfaces: ~ppc points per occupied column); the thinnest axis gets int(c / h) + 1 layers
// only as far as the cell budget allows (down to a single layer), so the table stays O(n) cells.
__global__ void __launch_bounds__(1024)
grid_setup_kernel(const float* __restrict__ support, uint8_t* __restrict__ ws, const __grid_constant__ LaunchClouds C,
                  float ppc) {
  int item;
  const CloudDev& cl = find_cloud(C, blockIdx.x, item);
  const float* S = support + (cl.support_off + item * cl.support_bstride) * 3;
  uint8_t* base = ws + cl.ws_off + item * cl.ws_item_bytes;
  // flat coalesced sweep over the 3n floats; element e belongs to axis e % 3
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  const int n3 = cl.n_support * 3;
  int ax = threadIdx.x % 3;
  // 1024 % 3 == 1: the axis advances by one per iteration, so three consecutive iterations touch x, y, z once each:
  // issue the three loads together (the sweep is latency bound: one CTA per cloud)
  int e = threadIdx.x;
  for (; e + 2 * 1024 < n3; e += 3 * 1024) {
    const float v0 = S[e], v1 = S[e + 1024], v2 = S[e + 2 * 1024];
    const int a1 = ax == 2 ? 0 : ax + 1, a2 = a1 == 2 ? 0 : a1 + 1;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float v = a == ax ? v0 : (a == a1 ? v1 : v2);
      lo[a] = fminf(lo[a], v); hi[a] = fmaxf(hi[a], v);
    }
  }
  for (; e < n3; e += 1024) {
    const float v = S[e];
#pragma unroll
    for (int a = 0; a < 3; ++a)
      if (a == ax) { lo[a] = fminf(lo[a], v); hi[a] = fmaxf(hi[a], v); }
    ax = ax == 2 ? 0 : ax + 1;
  }
  __shared__ float red[6][32];
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
  if (warp == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float l = red[a][lane], h = red[3 + a][lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        l = fminf(l, __shfl_xor_sync(0xffffffffu, l, o));
        h = fmaxf(h, __shfl_xor_sync(0xffffffffu, h, o));
      }
      lo[a] = l; hi[a] = h;
    }
    if (lane == 0) {
      const float e[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
      int thin = 0;  // axis of the smallest extent
      if (e[1] < e[thin]) thin = 1;
      if (e[2] < e[thin]) thin = 2;
      const float a = fmaxf(e[0], fmaxf(e[1], e[2]));
      const float b = e[0] + e[1] + e[2] - a - e[thin];
      float h = sqrtf(ppc * fmaxf(a * b, 1e-30f) / float(cl.n_support));
      if (!(h > 0.f) || !isfinite(h)) h = 1.f;
      h = fmaxf(h, a * 1e-4f + 1e-30f);
      int d[3];
      for (;;) {
        for (int x = 0; x < 3; ++x) d[x] = int(e[x] / h) + 1;
        long long plane = 1;
        for (int x = 0; x < 3; ++x) if (x != thin) plane *= d[x];
        if (plane <= cl.n_cells_max) {
          const long long layers = cl.n_cells_max / plane;
          if (d[thin] > layers) d[thin] = int(layers);
          break;
        }
        h *= 1.25f;
      }
      float hs[3];
      for (int x = 0; x < 3; ++x) hs[x] = h;
      if (d[thin] != int(e[thin] / h) + 1) hs[thin] = fmaxf(e[thin] / float(d[thin]) * 1.00001f, h);
      hdr.ox = lo[0]; hdr.oy = lo[1]; hdr.oz = lo[2];
      hdr.hx = hs[0]; hdr.hy = hs[1]; hdr.hz = hs[2];
      hdr.ihx = 1.f / hs[0]; hdr.ihy = 1.f / hs[1]; hdr.ihz = 1.f / hs[2];
      const float scale = fmaxf(a, fmaxf(fabsf(lo[0]), fmaxf(fabsf(lo[1]), fabsf(lo[2])))) + a;
      hdr.slack = scale * 3.8e-6f;  // 2^-18: >> fp32 rounding of (p - o) * inv_h, << h
      hdr.dx = d[0]; hdr.dy = d[1]; hdr.dz = d[2]; hdr.n_cells = d[0] * d[1] * d[2];
      hdr.pad[0] = hdr.pad[1] = 0;
      *reinterpret_cast<GridHeader*>(base) = hdr;
    }
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
  const int cx = cell_coord(S[size_t(i) * 3 + 0], g->ox, g->ihx, g->dx);
  const int cy = cell_coord(S[size_t(i) * 3 + 1], g->oy, g->ihy, g->dy);
  const int cz = cell_coord(S[size_t(i) * 3 + 2], g->oz, g->ihz, g->dz);
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
  constexpr int PER = 4;  // consecutive cells per thread
  for (int c0 = 0; c0 < n; c0 += 1024 * PER) {
    const int i0 = c0 + threadIdx.x * PER;
    int v[PER], tsum = 0;
#pragma unroll
    for (int u = 0; u < PER; ++u) { v[u] = i0 + u < n ? cursor[i0 + u] : 0; tsum += v[u]; }
    int x = tsum;
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
    int excl = carry + (warp ? warp_sums[warp - 1] : 0) + x - tsum;
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      if (i0 + u < n) { start[i0 + u] = excl; cursor[i0 + u] = excl; }
      excl += v[u];
    }
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl;
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
// One warp per query, QPW queries per warp in sequence.  Phase 0 visits the 3x3x3 block of cells around the
// query's cell, phase r >= 1 the Chebyshev shell of radius r + 1.  Within a phase every lane owns one cell-row
// range [rb, re) of the sorted point array; the ranges are FLATTENED (warp prefix sum + per-lane binary search
// through shuffles) so that each offer carries 32 real candidates however short the individual ranges are.
template <class Sel>
__device__ __forceinline__ void grid_offer_ranges(Sel& sel, const float4* __restrict__ pts, int rb,
                                                  int re, float qx, float qy, float qz, int k, int lane) {
  const int len = re - rb;
  int incl = len;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  const int excl = incl - len;
  for (int t0 = 0; t0 < total; t0 += 32) {
    const int t = t0 + lane;
    const bool valid = t < total;
    // owner range of flat position t = number of lanes whose inclusive prefix is <= t
    int pos = 0;
#pragma unroll
    for (int step = 16; step > 0; step >>= 1) {
      const int v = __shfl_sync(0xffffffffu, incl, pos + step - 1);
      if (v <= t) pos += step;
    }
    const int src = min(pos, 31);
    const int b0 = __shfl_sync(0xffffffffu, rb, src);
    const int x0 = __shfl_sync(0xffffffffu, excl, src);
    const float4 p = pts[valid ? b0 + (t - x0) : 0];
    const float d = dist2_ref(qx, qy, qz, p.x, p.y, p.z);
    sel.offer(d, __float_as_int(p.w), valid, k, lane);
  }
}

// one query, one warp (see knn_grid_kernel)
template <class Sel>
__device__ __forceinline__ void grid_query(const JobDev& job, const GridHeader& g, const int* __restrict__ start,
                                           const float4* __restrict__ pts, float qx, float qy, float qz, int k,
                                           int lane, int32_t* __restrict__ idx, float* __restrict__ dist2,
                                           long long o) {
  const int cx = cell_coord(qx, g.ox, g.ihx, g.dx);
  const int cy = cell_coord(qy, g.oy, g.ihy, g.dy);
  const int cz = cell_coord(qz, g.oz, g.ihz, g.dz);
  Sel sel;
  sel.reset();
  bool done = false;

  if (k <= HALF_BLOCK_K) {
    // Few neighbours wanted (the 1-NN interpolation jobs are half of all queries): look at the 2x2x2 block of
    // cells nearest to the query first -- per axis the query's own cell and the neighbour on the side of the
    // cell half the query lies in.  Every unvisited point is then at least half a cell away; with ~16 points
    // per cell column that almost always proves the result, at 4/9 of the candidates of the 3x3x3 block.
    const float fx = (qx - g.ox) * g.ihx - float(cx), fy = (qy - g.oy) * g.ihy - float(cy),
                fz = (qz - g.oz) * g.ihz - float(cz);
    const int x0 = max(fx < 0.5f ? cx - 1 : cx, 0), x1 = min(fx < 0.5f ? cx : cx + 1, g.dx - 1);
    const int y0 = max(fy < 0.5f ? cy - 1 : cy, 0), y1 = min(fy < 0.5f ? cy : cy + 1, g.dy - 1);
    const int z0 = max(fz < 0.5f ? cz - 1 : cz, 0), z1 = min(fz < 0.5f ? cz : cz + 1, g.dz - 1);
    int rb = 0, re = 0;
    if (lane < 4) {
      const int z = z0 + (lane >> 1), y = y0 + (lane & 1);
      if (z <= z1 && y <= y1) {
        const int rowbase = (z * g.dy + y) * g.dx;
        rb = start[rowbase + x0];
        re = start[rowbase + x1 + 1];
      }
    }
    grid_offer_ranges(sel, pts, rb, re, qx, qy, qz, k, lane);
    sel.flush(k, lane);
    float bound = FLT_MAX;
    if (x0 > 0) bound = fminf(bound, qx - (g.ox + float(x0) * g.hx));
    if (x1 < g.dx - 1) bound = fminf(bound, (g.ox + float(x1 + 1) * g.hx) - qx);
    if (y0 > 0) bound = fminf(bound, qy - (g.oy + float(y0) * g.hy));
    if (y1 < g.dy - 1) bound = fminf(bound, (g.oy + float(y1 + 1) * g.hy) - qy);
    if (z0 > 0) bound = fminf(bound, qz - (g.oz + float(z0) * g.hz));
    if (z1 < g.dz - 1) bound = fminf(bound, (g.oz + float(z1 + 1) * g.hz) - qz);
    if (bound == FLT_MAX) {
      done = true;
    } else {
      bound -= g.slack;
      done = bound > 0.f && sel.kd != D_EMPTY && __uint_as_float(sel.kd) < bound * bound;
    }
    if (!done) sel.reset();   // the 3x3x3 block below contains these cells again
  }

  for (int rho = 1; rho <= R_MAX && !done; ++rho) {
    // rho == 1: the whole 3x3x3 block, one range per (dz, dy) row.
    // rho >= 2: the shell of radius rho: rows (dz, dy) in [-rho, rho]^2, two range slots per row
    //           (rim rows: the full x span | nothing; inner rows: the cell at -rho | the cell at +rho).
    const int side = 2 * rho + 1;
    const int nslots = rho == 1 ? side * side : 2 * side * side;
    for (int s0 = 0; s0 < nslots; s0 += 32) {
      const int s = s0 + lane;
      int rb = 0, re = 0;
      if (s < nslots) {
        const int row = rho == 1 ? s : s >> 1, second = rho == 1 ? 0 : s & 1;
        const int dz = row / side - rho, dy = row % side - rho;
        const int z = cz + dz, y = cy + dy;
        if (z >= 0 && z < g.dz && y >= 0 && y < g.dy) {
          const bool rim = rho == 1 || max(abs(dz), abs(dy)) == rho;
          int x_lo, x_hi;
          if (rim) { x_lo = cx - rho; x_hi = second ? x_lo - 1 : cx + rho; }
          else     { x_lo = second ? cx + rho : cx - rho; x_hi = x_lo; }
          x_lo = max(x_lo, 0); x_hi = min(x_hi, g.dx - 1);
          if (x_lo <= x_hi) {
            const int rowbase = (z * g.dy + y) * g.dx;
            rb = start[rowbase + x_lo];
            re = start[rowbase + x_hi + 1];
          }
        }
      }
      grid_offer_ranges(sel, pts, rb, re, qx, qy, qz, k, lane);
    }
    sel.flush(k, lane);
    // distance from q to the nearest face of the visited block behind which unvisited cells exist
    float bound = FLT_MAX;
    if (cx - rho > 0) bound = fminf(bound, qx - (g.ox + float(cx - rho) * g.hx));
    if (cx + rho < g.dx - 1) bound = fminf(bound, (g.ox + float(cx + rho + 1) * g.hx) - qx);
    if (cy - rho > 0) bound = fminf(bound, qy - (g.oy + float(cy - rho) * g.hy));
    if (cy + rho < g.dy - 1) bound = fminf(bound, (g.oy + float(cy + rho + 1) * g.hy) - qy);
    if (cz - rho > 0) bound = fminf(bound, qz - (g.oz + float(cz - rho) * g.hz));
    if (cz + rho < g.dz - 1) bound = fminf(bound, (g.oz + float(cz + rho + 1) * g.hz) - qz);
    if (bound == FLT_MAX) {
      done = true;  // the block covers the whole grid
    } else {
      bound -= g.slack;
      // strict: an unvisited point at exactly the k-th distance could still win the index tie
      done = bound > 0.f && sel.kd != D_EMPTY && __uint_as_float(sel.kd) < bound * bound;
    }
  }
  if (!done) {  // pathological query (far outside / sparse region): scan the whole cloud
    sel.reset();
    for (int j0 = 0; j0 < job.n_support; j0 += 32) {
      const int j = j0 + lane;
      const bool valid = j < job.n_support;
      const float4 p = pts[valid ? j : job.n_support - 1];
      const float d = dist2_ref(qx, qy, qz, p.x, p.y, p.z);
      sel.offer(d, __float_as_int(p.w), valid, k, lane);
    }
    sel.flush(k, lane);
  }
  if (lane < k) {
    idx[o + lane] = sel.ei[0];
    if (dist2) dist2[o + lane] = __uint_as_float(sel.ed[0]);
  }
}


#ifndef GADM_KNN_MINB
#define GADM_KNN_MINB 6
#endif
__global__ void __launch_bounds__(WARPS * 32, GADM_KNN_MINB)
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

  for (int t = 0; t < GQPW; ++t) {
    const int qi = qtile * GQPB + warp * GQPW + t;
    if (qi >= job.n_query) break;  // warp-uniform
    const float qx = Q[qi * 3 + 0], qy = Q[qi * 3 + 1], qz = Q[qi * 3 + 2];
    const long long o = obase + (long long)qi * k;
    if (k == 1) grid_query<WarpMin>(job, g, start, pts, qx, qy, qz, k, lane, idx, dist2, o);
    else grid_query<WarpSelect<GRID_E>>(job, g, start, pts, qx, qy, qz, k, lane, idx, dist2, o);
  }
}

// cell budget of a cloud: O(n) so that the histogram scan stays cheap (grid_setup_kernel sizes the cells to fit)
int cells_for(int n_support) {
  long long c = n_support / 2;
  if (c < 1024) c = 1024;
  if (c > (1 << 18)) c = 1 << 18;
  return int(c);
}

float g_ppc = 16.f;           // target points per occupied cell column
int g_grid_min_support = 128; // AUTO: smaller clouds are scanned by BRUTE

bool job_uses_grid(const gadm_knn_job& j, int algo) {
  if (algo == GADM_KNN_BRUTE) return false;
  if (algo == GADM_KNN_GRID) return true;
  return j.n_support >= g_grid_min_support;
}

bool same_cloud(const gadm_knn_job& a, const gadm_knn_job& b) {
  return a.support_off == b.support_off && a.n_support == b.n_support && a.support_bstride == b.support_bstride &&
         a.batch == b.batch;
}

}  // namespace

int knn3d_configure() { return GADM_OK; }

// tuning knobs for experiments (gadm_config_set; results are identical for every setting; -1 restores the default)
int knn3d_config_set(const char* key, int value) {
  if (!strcmp(key, "knn.ppc")) {            // target points per occupied cell column
    if (value < 0) { g_ppc = 16.f; return GADM_OK; }
    if (value < 1 || value > 64) return GADM_ERR_BAD_ARG;
    g_ppc = float(value);
    return GADM_OK;
  }
  if (!strcmp(key, "knn.grid_min")) {       // AUTO: clouds with fewer points are scanned by BRUTE
    if (value < 0) { g_grid_min_support = 128; return GADM_OK; }
    if (value < 1) return GADM_ERR_BAD_ARG;
    g_grid_min_support = value;
    return GADM_OK;
  }
  return GADM_ERR_BAD_ARG;
}

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
      grid_setup_kernel<<<C.total_items, 1024, 0, stream>>>(support, ws, C, g_ppc);
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

  // ---- queries, MAX_JOBS jobs per launch; the scans and the grid searches have their own tile tables
  for (int j0 = 0; j0 < n_jobs && rc == GADM_OK; j0 += MAX_JOBS) {
    const int nj = (n_jobs - j0 < MAX_JOBS) ? n_jobs - j0 : MAX_JOBS;
    for (int kind = 0; kind < 2; ++kind) {          // 0: BRUTE, 1: GRID
      LaunchJobs L;
      L.n_jobs = 0;
      int tiles = 0;
      const int qpb = kind ? GQPB : QPB;
      for (int i = 0; i < nj; ++i) {
        const gadm_knn_job& s = jobs[j0 + i];
        if (int(job_uses_grid(s, algo)) != kind) continue;
        JobDev& d = L.jobs[L.n_jobs++];
        d.support_off = s.support_off; d.query_off = s.query_off; d.out_off = s.out_off;
        d.support_bstride = s.support_bstride; d.query_bstride = s.query_bstride; d.out_bstride = s.out_bstride;
        d.n_support = s.n_support; d.n_query = s.n_query; d.k = s.k; d.batch = s.batch;
        d.use_grid = kind;
        d.n_cells_max = cells_for(s.n_support);
        d.ws_off = kind ? ws_off[j0 + i] : 0;
        d.ws_item_bytes = kind ? (long long)grid_item_bytes(s.n_support, d.n_cells_max) : 0;
        d.tile_begin = tiles;
        d.tiles_per_item = (s.n_query + qpb - 1) / qpb;
        tiles += d.tiles_per_item * s.batch;
      }
      L.total_tiles = tiles;
      if (tiles == 0) continue;
      if (kind)
        knn_grid_kernel<<<tiles, WARPS * 32, 0, stream>>>(query, idx, dist2, static_cast<uint8_t*>(workspace), L);
      else
        knn_brute_kernel<<<tiles, WARPS * 32, 0, stream>>>(support, query, idx, dist2, L);
    }
    rc = check_launch();
  }
  delete[] ws_off;
  delete[] owner;
  return rc;
}

}  // namespace gadm
