// Shared definitions of the tcgen05 matcher kernels (match_sm100.cu) and the CircleLoss kernel (circle_sm100.cu).
#pragma once
#include <cuda_fp16.h>
#include <float.h>
#include <stdio.h>
#include <stdlib.h>

#include <type_traits>

#include "gadm_internal.h"
#include "ptx.cuh"

namespace gadm {
namespace {

constexpr int BM = 128;               // rows (scene points) per row tile == UMMA M
constexpr int BK = 64;                // bf16 elements per 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int A_BLK_BYTES = BM * BK * 2;     // 16 KB
constexpr int MAX_STAGES = 6;
constexpr int BN = 256;               // model vertices per accumulator tile == UMMA N
constexpr int B_STAGE_BYTES = BN * BK * 2;   // 32 KB
constexpr int AUX_SLOTS = 4;                 // per-tile {1/|m|, x, y, z} ring, decoupled from the two accumulators
constexpr int PLANE_BYTES = BN * 4;          // one fp32 plane of a tile
constexpr int EPI_WARPS = 16;
constexpr int NUM_THREADS = (EPI_WARPS + 2) * 32;   // warps 0-15 epilogue, 16 TMA, 17 UMMA
constexpr int GRP = 8;                                 // columns per argmax group (stash granularity)
constexpr int STASH_BYTES = EPI_WARPS * 32 * GRP * 4;  // per epilogue thread: the 8 scores of its best group
constexpr int STASH_PLANE = EPI_WARPS * 32 * 16;       // float4 k of thread t lives at k * STASH_PLANE + t * 16
static_assert(STASH_PLANE == 8192, "ptx::sts_stash8 hard-codes the plane stride");
constexpr int TMEM_COLS = 512;

constexpr int STASH_SLOT_BYTES = EPI_WARPS * 32 * 4 * 32;   // per-SM argmax stash slot of the workspace (64 KB)
constexpr int NO_RECORD = 0x40000000;                       // index of a track that holds no record (loses every tie)
constexpr int PART_ROWS = 2 * BM;                           // rows of a row block (two row tiles)

struct Barriers {
  uint64_t full[MAX_STAGES];
  uint64_t empty[MAX_STAGES];
  uint64_t a_full;
  uint64_t a_free;       // persistent kernels: every MMA that reads the row tiles of the segment has completed
  uint64_t s_full[2];    // accumulator a complete (UMMA commit)
  uint64_t s_free[2];    // accumulator a drained by all of its epilogue warps
  uint64_t aux_full[AUX_SLOTS];
  uint64_t aux_empty[AUX_SLOTS];
  uint32_t tmem_base;
  int merge_lo, merge_hi;   // persistent kernels: CTA range whose partial results this CTA has to merge (lo < 0: none)
  uint32_t pad;
};

struct MatchParams {
  const float* rinv_rows;  // [B, N]
  const float* pad_sim;    // [B, N] or null
  const float* scales;     // [n_obj, M]  1/|m_j|
  const float* planes;     // [3, n_obj, M] model x / y / z planes (SOFT)
  const uint8_t* mask;     // [B, N] or null
  const int32_t* obj_id;   // [B] or null
  // row compaction (evaluator.py:82-88): N is the row CAPACITY of the operand arrays (rows, rinv_rows, pad_sim are
  // [B, N] in compacted order); frame b holds n_rows[b] <= N rows (null: N), and compacted row j is written to the
  // outputs at position row_map[b, j] of a frame of N_out rows (null: position j)
  const int32_t* n_rows;   // [B] device, or null
  const int32_t* row_map;  // [B, N] device, or null
  int N_out;
  int64_t* idx;
  float* max_sim;
  float* weight;
  float* soft_xyz;
  int B, N, M, KB, n_obj, stages;
  int pad_mode;
  float gamma_log2e;
  uint8_t* stash;          // stash_slots slots of STASH_SLOT_BYTES (workspace, indexed by %smid), else null
  int stash_slots;
  // persistent kernels: the (frame, row block, model tile) units are dealt out evenly, in linear order, to the CTAs of
  // the grid; a row block whose tiles end up in several CTAs is finished by the last of them to arrive
  int T;                   // model tiles per row
  int RB;                  // row blocks (PART_ROWS rows) per frame
  long long total_units;   // B * RB * T
  unsigned int* seg_count; // [grid] arrival counters (workspace, zeroed by the launcher)
  float* partial;          // [grid][2][PART_ROWS][8] partial results (workspace)
  int unit_scales;         // 1 = GADM_MATCH_ARGMAX_UNIT: kernels that can, skip the column scales (the others apply them);
                           // 2 = GADM_MATCH_ARGMAX_BF16N: exact, scales known to be <= 1 + 2^-8 (chunk pruning)
};

__device__ __forceinline__ int frame_rows(const MatchParams& p, int b) {
  return p.n_rows ? min(p.n_rows[b], p.N) : p.N;
}
// where the results of (compacted) row `row` of frame b go
__device__ __forceinline__ size_t out_pos(const MatchParams& p, int b, int row) {
  return size_t(b) * p.N_out + (p.row_map ? p.row_map[size_t(b) * p.N + row] : row);
}

// Schedule of the persistent kernels: the units (frame, row block, model tile), in that order, are dealt out evenly to
// the CTAs.  With per-frame row counts in device memory (row compaction) the number of row blocks of a frame is only
// known on the device: `prefix` (shared memory, B + 1 entries, built by build()) holds the running count of row blocks.
struct Sched {
  long long total;       // units
  int T, RB, B;          // model tiles per row; row blocks per frame when every frame has the same number
  const int* prefix;     // or null
  int nc;                // scheduling entities: CTAs (gridDim.x), or CTA pairs
  __device__ __forceinline__ long long begin(int c) const { return (long long)c * total / (long long)nc; }
  // the CTA (pair) whose range holds unit u: the largest c with begin(c) <= u
  __device__ __forceinline__ int cta_of(long long u) const { return int(((u + 1) * (long long)nc - 1) / total); }
  // frame and row block of global row block g
  __device__ __forceinline__ void locate(int g, int& b, int& rb) const {
    if (prefix == nullptr) { b = g / RB; rb = g - b * RB; return; }
    int lo = 0, hi = B;                                    // largest b with prefix[b] <= g
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (prefix[mid] <= g) lo = mid; else hi = mid;
    }
    b = lo; rb = g - prefix[lo];
  }
};
// Called by every thread of the CTA (contains a __syncthreads when row counts are given).
__device__ __forceinline__ Sched sched_build(const MatchParams& p, int* smem_prefix, int rows_per_block, int nc) {
  Sched s;
  s.T = p.T; s.RB = p.RB; s.B = p.B; s.prefix = nullptr; s.total = p.total_units; s.nc = nc;
  if (p.n_rows != nullptr) {
    if (threadIdx.x == 0) {
      int acc = 0;
      for (int b = 0; b < p.B; ++b) {
        smem_prefix[b] = acc;
        acc += (frame_rows(p, b) + rows_per_block - 1) / rows_per_block;
      }
      smem_prefix[p.B] = acc;
    }
    __syncthreads();
    s.prefix = smem_prefix;
    s.total = (long long)smem_prefix[p.B] * p.T;
  }
  // Never more scheduling entities than units: begin() would leave CTAs with empty ranges BETWEEN the CTAs of a row
  // block, and the merge of that block counts one arrival per CTA of [c_lo, c_hi].  (Only the device knows the unit
  // count when rows were compacted.)  CTAs with index >= nc have nothing to do.
  if ((long long)s.nc > s.total) s.nc = int(s.total);
  return s;
}

// Bank slot of frame b.  A slot outside [0, n_obj) (the host wrappers reject it; a device-side id cannot be checked
// without a synchronisation) is clamped: the results of that frame are then those of a neighbouring object, but no
// table is ever read out of bounds.
__device__ __forceinline__ int frame_object(const MatchParams& p, int b) {
  if (p.obj_id) return min(max(p.obj_id[b], 0), p.n_obj - 1);
  return p.n_obj == p.B ? b : 0;
}

}  // namespace
}  // namespace gadm
