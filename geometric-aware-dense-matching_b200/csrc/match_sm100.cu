// Fused matching head for sm_100a: descriptor similarity (tcgen05 UMMA, bf16 in / fp32 accumulate in TMEM,
// operands staged by TMA) + row-wise argmax / softmax / soft model coordinates, and its training-side twin
// (CircleLoss forward / dL/dsim).  The [N, M] score matrix never leaves TMEM / registers.
//
// Replaces (reference tree): evaluator.py:89-93 (normalize, normalize, matmul, torch.max) and the padded
// variants utils/pvn3d_eval_utils_kpls.py:436-444 / models/geoMatch_DGCNN.py:92-99; models/geoMatch.py:55-83,102-157 +
// models/loss.py:475-490 (circle_kernel).  The softmax weight and soft coordinates are the extension defined in
// oracle/match_oracle.py.
//
// Kernels (what runs when: match_launch_t; measurements and bounds: DESIGN.md 3.1, profiles/SUMMARY_r1.md)
//   match_kernel<soft|argmax, RT>   thread = row; RT row tiles of 128 scene points per CTA.  RT = 1: the two TMEM
//                                   accumulators alternate between model tiles; RT = 2: accumulator r = row tile r with
//                                   8 fixed epilogue warps.  The no-workspace ARGMAX path and SOFT for K' > 128.
//   match_pair_kernel<soft|argmax>  256 rows per CTA, 128-vertex model tiles, every epilogue thread owns the same lane
//                                   of both row tiles (a per-column constant serves two scores).  Default for SOFT.
//   match_alt_kernel<exact|unit>    ARGMAX default: two row tiles per CTA, ALL 16 epilogue warps drain one accumulator
//                                   while the tensor core fills the other; stash in the per-SM workspace slot; running
//                                   maxima shared across the column slices of a row.  unit: no per-column constant.
//   match_ta_kernel, match_frag_kernel   opt-in experiments kept parity-green (A operand in tensor memory;
//                                   tcgen05.ld.16x256b fragment layout): both measured slower, see DESIGN.md.
//   circle_kernel<fwd|grad>         CircleLoss: masked exponential sums / dL/dsim in the epilogue.
//
// Common skeleton
//     warp 16     TMA producer: the row tile(s) once, then model tiles (256 vertices x 64 k, 32 KB) through an S-stage
//                 mbarrier ring, plus per tile the column scales 1/|m_j| and (SOFT / circle) the coordinate planes
//     warp 17     UMMA issuer: 128x256x16 tcgen05.mma into two 256-column TMEM accumulators, tcgen05.commit
//     warps 0..15 epilogue.  Per 32-column chunk: tcgen05.ld, score = acc * 1/|m_j| (packed f32x2), 3-input max tree
//                 per 8 columns.  The position of the maximum inside its 8-column group is NOT searched in the loop (a
//                 search is ~70 warp-divergent instructions and some lane of a warp needs one in most chunks): a
//                 thread whose running maximum rises stores the group's 8 scores to a private stash with two
//                 predicated 16-byte stores (shared memory or the L2-resident workspace), and the first maximal index
//                 is looked up there once, after the last tile.  SOFT adds p = 2^(score * g) (no reference exponent:
//                 |gamma| <= 40 keeps the sums inside the fp32 range) and fp32 sums of p and p * xyz.  The column
//                 slices of a row merge through shared memory at the end.
//   What bounds them: with two row tiles per CTA the UMMA operand reads + TMA writes alone fill the SM's shared-memory
//   data pipe (1024 wavefronts per 1024-cycle tile), so every epilogue LDS / store wavefront lengthens the tile;
//   SOFT is additionally capped by MUFU.EX2 (16/clk/SM: 2048 cycles per 128x256 tile).
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

struct Barriers {
  uint64_t full[MAX_STAGES];
  uint64_t empty[MAX_STAGES];
  uint64_t a_full;
  uint64_t s_full[2];    // accumulator a complete (UMMA commit)
  uint64_t s_free[2];    // accumulator a drained by all of its epilogue warps
  uint64_t aux_full[AUX_SLOTS];
  uint64_t aux_empty[AUX_SLOTS];
  uint32_t tmem_base;
  uint32_t pad;
};

struct MatchParams {
  const float* rinv_rows;  // [B, N]
  const float* pad_sim;    // [B, N] or null
  const float* scales;     // [n_obj, M]  1/|m_j|
  const float* planes;     // [3, n_obj, M] model x / y / z planes (SOFT)
  const uint8_t* mask;     // [B, N] or null
  const int32_t* obj_id;   // [B] or null
  int64_t* idx;
  float* max_sim;
  float* weight;
  float* soft_xyz;
  int B, N, M, KB, n_obj, stages;
  int pad_mode;
  float gamma_log2e;
  uint8_t* stash;          // fragment-layout kernel: stash_slots slots of FRAG_STASH_BYTES (workspace), else null
  int stash_slots;
  int unit_scales;         // 1 = GADM_MATCH_ARGMAX_UNIT: kernels that can, skip the column scales (the others apply them);
                           // 2 = GADM_MATCH_ARGMAX_BF16N: exact, scales known to be <= 1 + 2^-8 (chunk pruning)
  const void* rows_ptr;    // [B, N, K'] bf16 (match_ta_kernel loads its A operand from global memory)
};

__device__ __forceinline__ int frame_object(const MatchParams& p, int b) {
  if (p.obj_id) return p.obj_id[b];
  return p.n_obj == p.B ? b : 0;
}

template <bool kSoft, int RT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
match_kernel(const __grid_constant__ CUtensorMap tmap_rows, const __grid_constant__ CUtensorMap tmap_cols,
             const MatchParams p) {
  constexpr int AUX_BYTES = kSoft ? 4 * PLANE_BYTES : PLANE_BYTES;
  constexpr int SL = 4 / RT;                 // column slices per row (epilogue threads that share a row)
  constexpr int CS = BN / SL;                // columns per slice: 64 (RT = 1) or 128 (RT = 2)
  constexpr int ACC_WARPS = EPI_WARPS / RT;  // epilogue warps per accumulator (RT = 1: all of them, either one)

  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B atoms need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                                    // [RT][KB] blocks of 128 rows x 64 k
  uint8_t* smem_b = smem_a + RT * p.KB * A_BLK_BYTES;
  uint8_t* smem_aux = smem_b + p.stages * B_STAGE_BYTES;   // per slot: [1/|m| x256 | x x256 | y x256 | z x256]
  uint8_t* smem_stash = smem_aux + AUX_SLOTS * AUX_BYTES;
  Barriers* bars = reinterpret_cast<Barriers*>(smem_stash + STASH_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int row0 = blockIdx.x * (BM * RT);
  const int obj = frame_object(p, b);
  const int num_tiles = (p.M + BN - 1) / BN;

  if (warp == EPI_WARPS && lane == 0) {
    ptx::prefetch_tensormap(&tmap_rows);
    ptx::prefetch_tensormap(&tmap_cols);
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(&bars->full[s], 1);
      ptx::mbar_init(&bars->empty[s], 1);
    }
    ptx::mbar_init(&bars->a_full, 1);
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&bars->s_full[a], 1);
      ptx::mbar_init(&bars->s_free[a], ACC_WARPS);  // one arrive per epilogue warp of the accumulator
    }
    for (int a = 0; a < AUX_SLOTS; ++a) {
      ptx::mbar_init(&bars->aux_full[a], 1);
      ptx::mbar_init(&bars->aux_empty[a], EPI_WARPS);
    }
    ptx::fence_mbar_init();
  }
  if (warp == EPI_WARPS + 1) {
    ptx::tmem_alloc(&bars->tmem_base, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == EPI_WARPS) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(&bars->a_full, RT * p.KB * A_BLK_BYTES);
      for (int r = 0; r < RT; ++r)
        for (int kb = 0; kb < p.KB; ++kb)
          ptx::tma_load_3d(smem_a + (r * p.KB + kb) * A_BLK_BYTES, &tmap_rows, &bars->a_full, kb * BK,
                           row0 + r * BM, b);   // rows >= N are zero-filled by TMA
      int stage = 0;
      uint32_t phase = 0;
      const size_t plane = size_t(p.n_obj) * p.M;
      const float* sc_tab = p.scales + size_t(obj) * p.M;
      const float* xyz_tab = p.planes + size_t(obj) * p.M;
      for (int t = 0; t < num_tiles; ++t) {
        const int slot = t % AUX_SLOTS;
        const uint32_t use = uint32_t(t) / AUX_SLOTS;
        const uint32_t bytes = uint32_t(min(BN, p.M - t * BN)) * 4;   // M % 8 == 0: a multiple of 16
        ptx::mbar_wait_sleep(&bars->aux_empty[slot], (use & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&bars->aux_full[slot], kSoft ? 4 * bytes : bytes);
        uint8_t* aux = smem_aux + slot * AUX_BYTES;
        ptx::bulk_load_1d(aux, sc_tab + size_t(t) * BN, bytes, &bars->aux_full[slot]);
        if (kSoft) {
#pragma unroll
          for (int c = 0; c < 3; ++c)
            ptx::bulk_load_1d(aux + (c + 1) * PLANE_BYTES, xyz_tab + c * plane + size_t(t) * BN, bytes,
                              &bars->aux_full[slot]);
        }
        for (int kb = 0; kb < p.KB; ++kb) {
          ptx::mbar_wait_sleep(&bars->empty[stage], phase ^ 1);
          ptx::mbar_arrive_expect_tx(&bars->full[stage], B_STAGE_BYTES);
          ptx::tma_load_3d(smem_b + stage * B_STAGE_BYTES, &tmap_cols, &bars->full[stage], kb * BK, t * BN, obj);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == EPI_WARPS + 1) {
    // ============================== UMMA issuer ==============================
    // (waits with a suspend-time hint: a busy poll competes for the MIO pipe the epilogue warps depend on)
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16_f32(BM, BN);
      ptx::mbar_wait(&bars->a_full, 0);
      int stage0 = 0;            // ring position of the tile's first K block
      uint32_t phase0 = 0;
#ifdef GADM_MATCH_TRACE
      long long tr_free = 0, tr_full = 0, tr_t0 = clock64();
#endif
      for (int t = 0; t < num_tiles; ++t) {
#pragma unroll
        for (int r = 0; r < RT; ++r) {
          const int acc = RT == 1 ? (t & 1) : r;
          const uint32_t use = RT == 1 ? uint32_t(t) >> 1 : uint32_t(t);
#ifdef GADM_MATCH_TRACE
          const long long c0 = clock64();
#endif
          ptx::mbar_wait_sleep(&bars->s_free[acc], (use & 1) ^ 1);
#ifdef GADM_MATCH_TRACE
          tr_free += clock64() - c0;
#endif
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * BN;
          int stage = stage0;
          uint32_t phase = phase0;
          for (int kb = 0; kb < p.KB; ++kb) {
            if (r == 0) {      // the stages of this tile stay resident until the last row tile has used them
#ifdef GADM_MATCH_TRACE
              const long long c1 = clock64();
#endif
              ptx::mbar_wait_sleep(&bars->full[stage], phase);
#ifdef GADM_MATCH_TRACE
              tr_full += clock64() - c1;
#endif
              ptx::tc_fence_after();
            }
            const uint32_t a_addr = ptx::smem_u32(smem_a + (r * p.KB + kb) * A_BLK_BYTES);
            const uint32_t b_addr = ptx::smem_u32(smem_b + stage * B_STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              ptx::umma_bf16_ss(d_tmem, ptx::umma_desc_sw128_kmajor(a_addr + k * UMMA_K * 2),
                                ptx::umma_desc_sw128_kmajor(b_addr + k * UMMA_K * 2), idesc, (kb | k) != 0);
            }
            if (r == RT - 1) ptx::umma_commit(&bars->empty[stage]);  // frees the stage once these MMAs have read it
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
          ptx::umma_commit(&bars->s_full[acc]);     // accumulator tile complete
          if (r == RT - 1) { stage0 = stage; phase0 = phase; }
        }
      }
#ifdef GADM_MATCH_TRACE
      if ((blockIdx.x == 3 || blockIdx.x == 40) && blockIdx.y == 0)
        printf("cta %d umma thread: %lld cycles for %d tiles x %d row tiles; waiting for a free accumulator %lld, "
               "for operands %lld\n", blockIdx.x, clock64() - tr_t0, num_tiles, RT, tr_free, tr_full);
#endif
    }
  } else {
    // ============================== epilogue warps (thread == row, SL column slices per row) ==============
    const int q = warp & 3;                          // TMEM lane quarter this warp may access (warp id % 4)
    const int rt = RT == 1 ? 0 : (warp >> 2) & 1;    // row tile
    const int sub = RT == 1 ? warp >> 2 : warp >> 3; // column slice of every tile
    const int row_in_tile = q * 32 + lane;
    const int row = row0 + rt * BM + row_in_tile;
    const bool row_ok = row < p.N;
    const size_t grow = size_t(b) * p.N + (row_ok ? row : 0);
    const float rs = row_ok ? p.rinv_rows[grow] : 0.f;
    const float g = p.gamma_log2e * rs;     // exponent scale: t = (acc * 1/|m_j|) * g   (log2 units)
    const uint32_t lane_base = tmem_base + (uint32_t(q * 32) << 16) + sub * CS;
    const uint32_t stash_addr = ptx::smem_u32(smem_stash) + threadIdx.x * 16;

    float vmax = -INFINITY;                 // running maximum of this thread's slice of the row
    int vgrp = 0;                           // first column of the 8-column group that first reached it
    // packed (even | odd column) partial sums, two independent chains each
    uint64_t l2a = 0, l2b = 0, ax2a = 0, ax2b = 0, ay2a = 0, ay2b = 0, az2a = 0, az2b = 0;

#ifdef GADM_MATCH_TRACE
    long long tr_wait = 0, tr_t0 = clock64();
#endif
    for (int t = 0; t < num_tiles; ++t) {
      const int acc = RT == 1 ? (t & 1) : rt;
      const uint32_t use = RT == 1 ? uint32_t(t) >> 1 : uint32_t(t);
      const int slot = t % AUX_SLOTS;
#ifdef GADM_MATCH_TRACE
      const long long c0 = clock64();
#endif
      // both barriers are polled back to back so that their check latencies overlap
      if (!(ptx::mbar_try_wait(&bars->aux_full[slot], (uint32_t(t) / AUX_SLOTS) & 1) &
            ptx::mbar_try_wait(&bars->s_full[acc], use & 1))) {
        ptx::mbar_wait_sleep(&bars->aux_full[slot], (uint32_t(t) / AUX_SLOTS) & 1);
        ptx::mbar_wait_sleep(&bars->s_full[acc], use & 1);
      }
      ptx::tc_fence_after();
#ifdef GADM_MATCH_TRACE
      tr_wait += clock64() - c0;
#endif
      const int ncols = min(BN, p.M - t * BN) - sub * CS;   // valid columns of this slice (may be <= 0)
      const uint32_t s_tmem = lane_base + acc * BN;
      const uint32_t sc_addr = ptx::smem_u32(smem_aux + slot * AUX_BYTES) + sub * CS * 4;
      const int col_base = t * BN + sub * CS;

      // One chunk of 32 columns starting at slice column col0 (at least one of them valid).  r[] holds the raw
      // accumulators on entry.  Scores (scaled by 1/|m_j|) stay packed in pairs; per 8-column group: 3-input max
      // tree, and -- predicated, no branch -- the group's scores go to the stash when they raise the thread's
      // running maximum.  SOFT: exponentials, sums of p and p * xyz.  kGuard (ragged last tile only) masks
      // columns >= ncols.
      auto process = [&](uint32_t (&r)[32], int col0, auto guard_tag) {
        constexpr int W = 32;
        constexpr int NG = W / GRP;
        constexpr bool kGuard = decltype(guard_tag)::value;
        const uint32_t sc = sc_addr + col0 * 4;
        uint64_t v[W / 2];
#pragma unroll
        for (int j4 = 0; j4 < W / 4; ++j4) {
          const float4 cm = ptx::lds128(sc + j4 * 16);
          v[j4 * 2 + 0] = ptx::fmul2(ptx::pack2(r[j4 * 4 + 0], r[j4 * 4 + 1]), ptx::pack2f(cm.x, cm.y));
          v[j4 * 2 + 1] = ptx::fmul2(ptx::pack2(r[j4 * 4 + 2], r[j4 * 4 + 3]), ptx::pack2f(cm.z, cm.w));
        }
        if (kGuard) {  // TMA zero-fills columns >= M and the stale scales behind them are meaningless
#pragma unroll
          for (int j = 0; j < W / 2; ++j) {
            float lo, hi;
            ptx::unpack2f(v[j], lo, hi);
            if (col0 + 2 * j >= ncols) lo = -INFINITY;
            if (col0 + 2 * j + 1 >= ncols) hi = -INFINITY;
            v[j] = ptx::pack2f(lo, hi);
          }
        }
#pragma unroll
        for (int h = 0; h < NG; ++h) {
          float f[GRP];
#pragma unroll
          for (int j = 0; j < GRP / 2; ++j) ptx::unpack2f(v[h * 4 + j], f[2 * j], f[2 * j + 1]);
          const float a0 = ptx::fmax3(f[0], f[1], f[2]), a1 = ptx::fmax3(f[3], f[4], f[5]);
          const float gm = ptx::fmax3(a0, a1, fmaxf(f[6], f[7]));
          // strict: an equal value in a later group never displaces the first maximal index
          const bool up = gm > vmax;
          ptx::sts_stash8(up, stash_addr, v[h * 4 + 0], v[h * 4 + 1], v[h * 4 + 2], v[h * 4 + 3]);
          vgrp = up ? col_base + col0 + h * GRP : vgrp;
          vmax = up ? gm : vmax;
        }
        if (kSoft) {
          // p = 2^(v*g), no reference exponent: |v*g| <= |gamma| log2(e) <= 58 (gadm_match_fwd admits |gamma| <= 40),
          // so p and its sums stay inside the fp32 range; sums in packed f32x2 (even | odd column)
          const uint64_t g2 = ptx::pack2f(g, g);
#pragma unroll
          for (int j4 = 0; j4 < W / 4; ++j4) {
            const float4 X = ptx::lds128(sc + PLANE_BYTES + j4 * 16);
            const float4 Y = ptx::lds128(sc + 2 * PLANE_BYTES + j4 * 16);
            const float4 Z = ptx::lds128(sc + 3 * PLANE_BYTES + j4 * 16);
            const uint64_t p01 = ptx::ex2_2(ptx::fmul2(v[j4 * 2 + 0], g2));
            const uint64_t p23 = ptx::ex2_2(ptx::fmul2(v[j4 * 2 + 1], g2));
            l2a = ptx::fadd2(l2a, p01);
            l2b = ptx::fadd2(l2b, p23);
            ax2a = ptx::ffma2(p01, ptx::pack2f(X.x, X.y), ax2a);
            ax2b = ptx::ffma2(p23, ptx::pack2f(X.z, X.w), ax2b);
            ay2a = ptx::ffma2(p01, ptx::pack2f(Y.x, Y.y), ay2a);
            ay2b = ptx::ffma2(p23, ptx::pack2f(Y.z, Y.w), ay2b);
            az2a = ptx::ffma2(p01, ptx::pack2f(Z.x, Z.y), az2a);
            az2b = ptx::ffma2(p23, ptx::pack2f(Z.z, Z.w), az2b);
          }
        }
      };
      using guard_off = std::integral_constant<bool, false>;
      using guard_on = std::integral_constant<bool, true>;

#ifdef GADM_DBG_NOEPI
      if (false)
#endif
#pragma unroll
      for (int c2 = 0; c2 < CS / 64; ++c2) {
        const int nv = ncols - c2 * 64;      // valid columns of this pair of chunks
        if (nv <= 0) break;
        // ARGMAX keeps two 32-column chunks in flight (registers allow it), SOFT one
        uint32_t ra[32], rb[32];
        ptx::tmem_ld_32x32(s_tmem + c2 * 64, ra);
        if (!kSoft) ptx::tmem_ld_32x32(s_tmem + c2 * 64 + 32, rb);
        ptx::tmem_ld_wait();
        if (nv >= 32) process(ra, c2 * 64, guard_off{});
        else process(ra, c2 * 64, guard_on{});         // ragged last tile
        if (nv > 32) {
          if (kSoft) {
            ptx::tmem_ld_32x32(s_tmem + c2 * 64 + 32, rb);
            ptx::tmem_ld_wait();
          }
          if (nv >= 64) process(rb, c2 * 64 + 32, guard_off{});
          else process(rb, c2 * 64 + 32, guard_on{});
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(&bars->s_free[acc]);
        ptx::mbar_arrive(&bars->aux_empty[slot]);
      }
    }
#ifdef GADM_MATCH_TRACE
    if ((blockIdx.x == 3 || blockIdx.x == 40) && blockIdx.y == 0 && lane == 0 && (warp == 0 || warp == 13))
      printf("cta %d epilogue warp %d: %lld cycles, of which waiting for accumulator/aux %lld\n", blockIdx.x, warp,
             clock64() - tr_t0, tr_wait);
#endif

    // ---- first maximal index of this slice: look it up in the stashed group (own writes, no barrier needed)
    int vidx = 0;
    if (vmax > -INFINITY) {
      int j_first = GRP - 1;
#pragma unroll
      for (int k = GRP / 4 - 1; k >= 0; --k) {
        const float4 sv = ptx::lds128(stash_addr + k * STASH_PLANE);
        if (sv.w == vmax) j_first = 4 * k + 3;
        if (sv.z == vmax) j_first = 4 * k + 2;
        if (sv.y == vmax) j_first = 4 * k + 1;
        if (sv.x == vmax) j_first = 4 * k + 0;
      }
      vidx = vgrp + j_first;
    }

    // ---- merge the column slices of every row: slices 1.. publish, slice 0 combines and writes the outputs.
    // The exchange buffer reuses the row tile's own A blocks: every MMA that reads them has completed (this warp
    // has seen the last s_full of its accumulator).
    float lsum = 0.f, ax = 0.f, ay = 0.f, az = 0.f;
    if (kSoft) {
      float e, o;
      ptx::unpack2f(ptx::fadd2(l2a, l2b), e, o); lsum = e + o;
      ptx::unpack2f(ptx::fadd2(ax2a, ax2b), e, o); ax = e + o;
      ptx::unpack2f(ptx::fadd2(ay2a, ay2b), e, o); ay = e + o;
      ptx::unpack2f(ptx::fadd2(az2a, az2b), e, o); az = e + o;
    }
    float* xch = reinterpret_cast<float*>(smem_a + rt * p.KB * A_BLK_BYTES);   // (SL - 1) * 128 * 32 B <= 16 KB
    if (sub > 0) {
      float* x = xch + ((sub - 1) * BM + row_in_tile) * 8;
      x[0] = vmax; x[1] = __int_as_float(vidx); x[3] = lsum;
      x[4] = ax; x[5] = ay; x[6] = az;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
    if (sub == 0 && row_ok) {
#pragma unroll
      for (int s2 = 0; s2 < SL - 1; ++s2) {
        const float* x = xch + (s2 * BM + row_in_tile) * 8;
        const float v1 = x[0];
        const int i1 = __float_as_int(x[1]);
        if (v1 > vmax || (v1 == vmax && i1 < vidx)) { vmax = v1; vidx = i1; }
      }
      const bool keep = p.mask == nullptr || p.mask[grow] != 0;
      float best = vmax * rs;
      int64_t best_idx = vidx;
      if (p.pad_mode != GADM_PAD_NONE) {
        const float ps = p.pad_sim[grow];
        if (ps > best) { best = ps; best_idx = p.M; }  // pad column is the last one: wins only if strictly larger
      }
      p.idx[grow] = keep ? best_idx : int64_t(-1);
      p.max_sim[grow] = keep ? best : 0.f;
      if (kSoft) {
        float l = lsum, sx = ax, sy = ay, sz = az;
#pragma unroll
        for (int s2 = 0; s2 < SL - 1; ++s2) {
          const float* x = xch + (s2 * BM + row_in_tile) * 8;
          l += x[3]; sx += x[4]; sy += x[5]; sz += x[6];
        }
        const float inv = 1.f / l;
        p.weight[grow] = keep ? ptx::ex2_approx(vmax * g) * inv : 0.f;  // softmax value at the maximum
        p.soft_xyz[grow * 3 + 0] = keep ? sx * inv : 0.f;
        p.soft_xyz[grow * 3 + 1] = keep ? sy * inv : 0.f;
        p.soft_xyz[grow * 3 + 2] = keep ? sz * inv : 0.f;
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == EPI_WARPS + 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <bool kSoft>
size_t match_smem_bytes(int RT, int KB, int stages) {
  return size_t(RT) * KB * A_BLK_BYTES + size_t(stages) * B_STAGE_BYTES + AUX_SLOTS * (kSoft ? 4 : 1) * PLANE_BYTES +
         STASH_BYTES + sizeof(Barriers) + 1024;
}

// the deepest model-tile ring that fits in 227 KB next to the row tiles (0: does not fit)
template <bool kSoft>
int match_stages(int RT, int KB) {
  int stages = MAX_STAGES;
  while (stages > 0 && match_smem_bytes<kSoft>(RT, KB, stages) > 227 * 1024) --stages;
  return stages;
}

// ---------------------------------------------------------------------------------------------------------------
// Paired-row variant.  The epilogue's broadcast LDS of per-column constants share the SM's 128 B/clk shared-memory
// port with the UMMA operand reads and the TMA writes, and in SOFT they are two thirds of that traffic (DESIGN.md
// 3.1).  Here every epilogue thread owns the SAME lane of TWO row tiles, so a constant fetched from shared memory is
// used for two scores: CTA = 256 rows, model tiles of 128 vertices, four 128-column accumulators
// (2 buffers x 2 row tiles), 128x128x16 MMAs (same tensor rate, tools/umma_probe.cu).
//   warp 16  TMA: the 256 x K' row tile once, model tiles (128 vertices x 64 k, 16 KB) + the tile's aux slice
//   warp 17  UMMA: per K block the two row tiles back to back (the stage is read twice, then freed)
//   warps 0..15  epilogue: warp w owns TMEM lanes 32 (w % 4).. of both row tiles and the 32-column slice w / 4
constexpr int PBN = 128;                       // model vertices per tile
constexpr int PB_STAGE_BYTES = PBN * BK * 2;   // 16 KB
constexpr int P_MAX_STAGES = 8;
constexpr int P_PLANE_BYTES = PBN * 4;
constexpr int P_CS = PBN / 4;                  // 32 columns per warp slice
constexpr int P_STASH_BYTES = 2 * STASH_BYTES; // two rows per thread

struct PairBarriers {
  uint64_t full[P_MAX_STAGES];
  uint64_t empty[P_MAX_STAGES];
  uint64_t a_full;
  uint64_t s_full[2];
  uint64_t s_free[2];
  uint64_t aux_full[AUX_SLOTS];
  uint64_t aux_empty[AUX_SLOTS];
  uint32_t tmem_base;
  uint32_t pad;
};

template <bool kSoft>
__global__ void __launch_bounds__(NUM_THREADS, 1)
match_pair_kernel(const __grid_constant__ CUtensorMap tmap_rows, const __grid_constant__ CUtensorMap tmap_cols,
                  const MatchParams p) {
  constexpr int AUX_BYTES = kSoft ? 4 * P_PLANE_BYTES : P_PLANE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                                    // [2][KB] blocks of 128 rows x 64 k
  uint8_t* smem_b = smem_a + 2 * p.KB * A_BLK_BYTES;
  uint8_t* smem_aux = smem_b + p.stages * PB_STAGE_BYTES;  // per slot: [1/|m| x128 | x | y | z]
  uint8_t* smem_stash = smem_aux + AUX_SLOTS * AUX_BYTES;
  PairBarriers* bars = reinterpret_cast<PairBarriers*>(smem_stash + P_STASH_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int row0 = blockIdx.x * (2 * BM);
  const int obj = frame_object(p, b);
  const int num_tiles = (p.M + PBN - 1) / PBN;

  if (warp == EPI_WARPS && lane == 0) {
    ptx::prefetch_tensormap(&tmap_rows);
    ptx::prefetch_tensormap(&tmap_cols);
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(&bars->full[s], 1);
      ptx::mbar_init(&bars->empty[s], 1);
    }
    ptx::mbar_init(&bars->a_full, 1);
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&bars->s_full[a], 1);
      ptx::mbar_init(&bars->s_free[a], EPI_WARPS);
    }
    for (int a = 0; a < AUX_SLOTS; ++a) {
      ptx::mbar_init(&bars->aux_full[a], 1);
      ptx::mbar_init(&bars->aux_empty[a], EPI_WARPS);
    }
    ptx::fence_mbar_init();
  }
  if (warp == EPI_WARPS + 1) {
    ptx::tmem_alloc(&bars->tmem_base, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == EPI_WARPS) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(&bars->a_full, 2 * p.KB * A_BLK_BYTES);
      for (int r = 0; r < 2; ++r)
        for (int kb = 0; kb < p.KB; ++kb)
          ptx::tma_load_3d(smem_a + (r * p.KB + kb) * A_BLK_BYTES, &tmap_rows, &bars->a_full, kb * BK,
                           row0 + r * BM, b);   // rows >= N are zero-filled by TMA
      int stage = 0;
      uint32_t phase = 0;
      const size_t plane = size_t(p.n_obj) * p.M;
      const float* sc_tab = p.scales + size_t(obj) * p.M;
      const float* xyz_tab = p.planes + size_t(obj) * p.M;
      for (int t = 0; t < num_tiles; ++t) {
        const int slot = t % AUX_SLOTS;
        const uint32_t use = uint32_t(t) / AUX_SLOTS;
        const uint32_t bytes = uint32_t(min(PBN, p.M - t * PBN)) * 4;   // M % 8 == 0: a multiple of 16
        ptx::mbar_wait_sleep(&bars->aux_empty[slot], (use & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&bars->aux_full[slot], kSoft ? 4 * bytes : bytes);
        uint8_t* aux = smem_aux + slot * AUX_BYTES;
        ptx::bulk_load_1d(aux, sc_tab + size_t(t) * PBN, bytes, &bars->aux_full[slot]);
        if (kSoft) {
#pragma unroll
          for (int c = 0; c < 3; ++c)
            ptx::bulk_load_1d(aux + (c + 1) * P_PLANE_BYTES, xyz_tab + c * plane + size_t(t) * PBN, bytes,
                              &bars->aux_full[slot]);
        }
        for (int kb = 0; kb < p.KB; ++kb) {
          ptx::mbar_wait_sleep(&bars->empty[stage], phase ^ 1);
          ptx::mbar_arrive_expect_tx(&bars->full[stage], PB_STAGE_BYTES);
          ptx::tma_load_3d(smem_b + stage * PB_STAGE_BYTES, &tmap_cols, &bars->full[stage], kb * BK, t * PBN, obj);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == EPI_WARPS + 1) {
    // ============================== UMMA issuer ==============================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16_f32(BM, PBN);
      ptx::mbar_wait(&bars->a_full, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < num_tiles; ++t) {
        const int buf = t & 1;
        ptx::mbar_wait_sleep(&bars->s_free[buf], ((uint32_t(t) >> 1) & 1) ^ 1);
        ptx::tc_fence_after();
        for (int kb = 0; kb < p.KB; ++kb) {
          ptx::mbar_wait_sleep(&bars->full[stage], phase);
          ptx::tc_fence_after();
          const uint32_t b_addr = ptx::smem_u32(smem_b + stage * PB_STAGE_BYTES);
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const uint32_t a_addr = ptx::smem_u32(smem_a + (r * p.KB + kb) * A_BLK_BYTES);
            const uint32_t d_tmem = tmem_base + (buf * 2 + r) * PBN;
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              ptx::umma_bf16_ss(d_tmem, ptx::umma_desc_sw128_kmajor(a_addr + k * UMMA_K * 2),
                                ptx::umma_desc_sw128_kmajor(b_addr + k * UMMA_K * 2), idesc, (kb | k) != 0);
          }
          ptx::umma_commit(&bars->empty[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(&bars->s_full[buf]);
      }
    }
  } else {
    // ============================== epilogue warps: thread == the same lane of both row tiles ==============
    const int q = warp & 3;
    const int sub = warp >> 2;              // 32-column slice of every tile
    const int row_in_tile = q * 32 + lane;
    int row[2];
    bool row_ok[2];
    size_t grow[2];
    float rs[2], g[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      row[r] = row0 + r * BM + row_in_tile;
      row_ok[r] = row[r] < p.N;
      grow[r] = size_t(b) * p.N + (row_ok[r] ? row[r] : 0);
      rs[r] = row_ok[r] ? p.rinv_rows[grow[r]] : 0.f;
      g[r] = p.gamma_log2e * rs[r];
    }
    const uint32_t lane_base = tmem_base + (uint32_t(q * 32) << 16) + sub * P_CS;
    const uint32_t stash_addr = ptx::smem_u32(smem_stash) + threadIdx.x * 16;   // row r: + r * 2 * STASH_PLANE

    float vmax[2] = {-INFINITY, -INFINITY};
    int vgrp[2] = {0, 0};
    uint64_t l2[2] = {0, 0}, ax2[2] = {0, 0}, ay2[2] = {0, 0}, az2[2] = {0, 0};   // packed (even | odd column)

    for (int t = 0; t < num_tiles; ++t) {
      const int buf = t & 1;
      const uint32_t use = uint32_t(t) >> 1;
      const int slot = t % AUX_SLOTS;
      if (!(ptx::mbar_try_wait(&bars->aux_full[slot], (uint32_t(t) / AUX_SLOTS) & 1) &
            ptx::mbar_try_wait(&bars->s_full[buf], use & 1))) {
        ptx::mbar_wait_sleep(&bars->aux_full[slot], (uint32_t(t) / AUX_SLOTS) & 1);
        ptx::mbar_wait_sleep(&bars->s_full[buf], use & 1);
      }
      ptx::tc_fence_after();
      const int ncols = min(PBN, p.M - t * PBN) - sub * P_CS;   // valid columns of this slice (may be <= 0)
      const uint32_t s_tmem0 = lane_base + (buf * 2 + 0) * PBN, s_tmem1 = lane_base + (buf * 2 + 1) * PBN;
      const uint32_t sc_addr = ptx::smem_u32(smem_aux + slot * AUX_BYTES) + sub * P_CS * 4;
      const int col_base = t * PBN + sub * P_CS;

      // W columns (slice columns col0 ..) of both rows: r0 / r1 hold the raw accumulators
      auto process = [&](auto& r0, auto& r1, int col0, auto guard_tag) {
        constexpr int W = int(sizeof(r0) / sizeof(r0[0]));
        constexpr bool kGuard = decltype(guard_tag)::value;
        const uint32_t sc = sc_addr + col0 * 4;
        uint64_t v[2][W / 2];
#pragma unroll
        for (int j4 = 0; j4 < W / 4; ++j4) {
          const float4 cm = ptx::lds128(sc + j4 * 16);
          const uint64_t c01 = ptx::pack2f(cm.x, cm.y), c23 = ptx::pack2f(cm.z, cm.w);
          v[0][j4 * 2 + 0] = ptx::fmul2(ptx::pack2(r0[j4 * 4 + 0], r0[j4 * 4 + 1]), c01);
          v[0][j4 * 2 + 1] = ptx::fmul2(ptx::pack2(r0[j4 * 4 + 2], r0[j4 * 4 + 3]), c23);
          v[1][j4 * 2 + 0] = ptx::fmul2(ptx::pack2(r1[j4 * 4 + 0], r1[j4 * 4 + 1]), c01);
          v[1][j4 * 2 + 1] = ptx::fmul2(ptx::pack2(r1[j4 * 4 + 2], r1[j4 * 4 + 3]), c23);
        }
        if (kGuard) {
#pragma unroll
          for (int rr = 0; rr < 2; ++rr)
#pragma unroll
            for (int j = 0; j < W / 2; ++j) {
              float lo, hi;
              ptx::unpack2f(v[rr][j], lo, hi);
              if (col0 + 2 * j >= ncols) lo = -INFINITY;
              if (col0 + 2 * j + 1 >= ncols) hi = -INFINITY;
              v[rr][j] = ptx::pack2f(lo, hi);
            }
        }
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
#pragma unroll
          for (int h = 0; h < W / GRP; ++h) {
            float f[GRP];
#pragma unroll
            for (int j = 0; j < GRP / 2; ++j) ptx::unpack2f(v[rr][h * 4 + j], f[2 * j], f[2 * j + 1]);
            const float a0 = ptx::fmax3(f[0], f[1], f[2]), a1 = ptx::fmax3(f[3], f[4], f[5]);
            const float gm = ptx::fmax3(a0, a1, fmaxf(f[6], f[7]));
            const bool up = gm > vmax[rr];
            ptx::sts_stash8(up, stash_addr + rr * 2 * STASH_PLANE, v[rr][h * 4 + 0], v[rr][h * 4 + 1],
                            v[rr][h * 4 + 2], v[rr][h * 4 + 3]);
            vgrp[rr] = up ? col_base + col0 + h * GRP : vgrp[rr];
            vmax[rr] = up ? gm : vmax[rr];
          }
        }
        if (kSoft) {
          // p = 2^(v*g), no reference exponent (see match_kernel)
          const uint64_t g20 = ptx::pack2f(g[0], g[0]), g21 = ptx::pack2f(g[1], g[1]);
#pragma unroll
          for (int j4 = 0; j4 < W / 4; ++j4) {
            const float4 X = ptx::lds128(sc + P_PLANE_BYTES + j4 * 16);
            const float4 Y = ptx::lds128(sc + 2 * P_PLANE_BYTES + j4 * 16);
            const float4 Z = ptx::lds128(sc + 3 * P_PLANE_BYTES + j4 * 16);
            const uint64_t X01 = ptx::pack2f(X.x, X.y), X23 = ptx::pack2f(X.z, X.w);
            const uint64_t Y01 = ptx::pack2f(Y.x, Y.y), Y23 = ptx::pack2f(Y.z, Y.w);
            const uint64_t Z01 = ptx::pack2f(Z.x, Z.y), Z23 = ptx::pack2f(Z.z, Z.w);
            const uint64_t pa0 = ptx::ex2_2(ptx::fmul2(v[0][j4 * 2 + 0], g20));
            const uint64_t pb0 = ptx::ex2_2(ptx::fmul2(v[0][j4 * 2 + 1], g20));
            const uint64_t pa1 = ptx::ex2_2(ptx::fmul2(v[1][j4 * 2 + 0], g21));
            const uint64_t pb1 = ptx::ex2_2(ptx::fmul2(v[1][j4 * 2 + 1], g21));
            l2[0] = ptx::fadd2(l2[0], ptx::fadd2(pa0, pb0));
            l2[1] = ptx::fadd2(l2[1], ptx::fadd2(pa1, pb1));
            ax2[0] = ptx::ffma2(pb0, X23, ptx::ffma2(pa0, X01, ax2[0]));
            ax2[1] = ptx::ffma2(pb1, X23, ptx::ffma2(pa1, X01, ax2[1]));
            ay2[0] = ptx::ffma2(pb0, Y23, ptx::ffma2(pa0, Y01, ay2[0]));
            ay2[1] = ptx::ffma2(pb1, Y23, ptx::ffma2(pa1, Y01, ay2[1]));
            az2[0] = ptx::ffma2(pb0, Z23, ptx::ffma2(pa0, Z01, az2[0]));
            az2[1] = ptx::ffma2(pb1, Z23, ptx::ffma2(pa1, Z01, az2[1]));
          }
        }
      };
      using guard_off = std::integral_constant<bool, false>;
      using guard_on = std::integral_constant<bool, true>;

#ifdef GADM_DBG_NOEPI
      if (false)
#endif
      if (ncols > 0) {
        if (!kSoft) {
          uint32_t ra[32], rb[32];
          ptx::tmem_ld_32x32(s_tmem0, ra);
          ptx::tmem_ld_32x32(s_tmem1, rb);
          ptx::tmem_ld_wait();
          if (ncols >= P_CS) process(ra, rb, 0, guard_off{});
          else process(ra, rb, 0, guard_on{});
        } else {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (ncols - h * 16 <= 0) break;
            uint32_t ra[16], rb[16];
            ptx::tmem_ld_32x16(s_tmem0 + h * 16, ra);
            ptx::tmem_ld_32x16(s_tmem1 + h * 16, rb);
            ptx::tmem_ld_wait();
            if (ncols - h * 16 >= 16) process(ra, rb, h * 16, guard_off{});
            else process(ra, rb, h * 16, guard_on{});
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(&bars->s_free[buf]);
        ptx::mbar_arrive(&bars->aux_empty[slot]);
      }
    }

    // ---- per row: first maximal index from the stash, merge of the four column slices, outputs
    int vidx[2];
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      vidx[rr] = 0;
      if (vmax[rr] > -INFINITY) {
        int j_first = GRP - 1;
#pragma unroll
        for (int k = GRP / 4 - 1; k >= 0; --k) {
          const float4 sv = ptx::lds128(stash_addr + rr * 2 * STASH_PLANE + k * STASH_PLANE);
          if (sv.w == vmax[rr]) j_first = 4 * k + 3;
          if (sv.z == vmax[rr]) j_first = 4 * k + 2;
          if (sv.y == vmax[rr]) j_first = 4 * k + 1;
          if (sv.x == vmax[rr]) j_first = 4 * k + 0;
        }
        vidx[rr] = vgrp[rr] + j_first;
      }
    }
    float lsum[2] = {0.f, 0.f}, ax[2] = {0.f, 0.f}, ay[2] = {0.f, 0.f}, az[2] = {0.f, 0.f};
    if (kSoft) {
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        float e, o;
        ptx::unpack2f(l2[rr], e, o); lsum[rr] = e + o;
        ptx::unpack2f(ax2[rr], e, o); ax[rr] = e + o;
        ptx::unpack2f(ay2[rr], e, o); ay[rr] = e + o;
        ptx::unpack2f(az2[rr], e, o); az[rr] = e + o;
      }
    }
    // exchange buffers alias the row tiles' own A blocks (all MMAs have completed: the last s_full was seen)
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      float* xch = reinterpret_cast<float*>(smem_a + rr * p.KB * A_BLK_BYTES);   // 3 * 128 * 32 B = 12 KB <= 16 KB
      if (sub > 0) {
        float* x = xch + ((sub - 1) * BM + row_in_tile) * 8;
        x[0] = vmax[rr]; x[1] = __int_as_float(vidx[rr]); x[3] = lsum[rr];
        x[4] = ax[rr]; x[5] = ay[rr]; x[6] = az[rr];
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
    if (sub == 0) {
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        if (!row_ok[rr]) continue;
        const float* xch = reinterpret_cast<const float*>(smem_a + rr * p.KB * A_BLK_BYTES);
        float vm = vmax[rr];
        int vi = vidx[rr];
#pragma unroll
        for (int s2 = 0; s2 < 3; ++s2) {
          const float* x = xch + (s2 * BM + row_in_tile) * 8;
          const float v1 = x[0];
          const int i1 = __float_as_int(x[1]);
          if (v1 > vm || (v1 == vm && i1 < vi)) { vm = v1; vi = i1; }
        }
        const size_t gr = grow[rr];
        const bool keep = p.mask == nullptr || p.mask[gr] != 0;
        float best = vm * rs[rr];
        int64_t best_idx = vi;
        if (p.pad_mode != GADM_PAD_NONE) {
          const float ps = p.pad_sim[gr];
          if (ps > best) { best = ps; best_idx = p.M; }
        }
        p.idx[gr] = keep ? best_idx : int64_t(-1);
        p.max_sim[gr] = keep ? best : 0.f;
        if (kSoft) {
          float l = lsum[rr], sx = ax[rr], sy = ay[rr], sz = az[rr];
#pragma unroll
          for (int s2 = 0; s2 < 3; ++s2) {
            const float* x = xch + (s2 * BM + row_in_tile) * 8;
            l += x[3]; sx += x[4]; sy += x[5]; sz += x[6];
          }
          const float inv = 1.f / l;
          p.weight[gr] = keep ? ptx::ex2_approx(vm * g[rr]) * inv : 0.f;
          p.soft_xyz[gr * 3 + 0] = keep ? sx * inv : 0.f;
          p.soft_xyz[gr * 3 + 1] = keep ? sy * inv : 0.f;
          p.soft_xyz[gr * 3 + 2] = keep ? sz * inv : 0.f;
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == EPI_WARPS + 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <bool kSoft>
size_t match_pair_smem_bytes(int KB, int stages) {
  return size_t(2) * KB * A_BLK_BYTES + size_t(stages) * PB_STAGE_BYTES + AUX_SLOTS * (kSoft ? 4 : 1) * P_PLANE_BYTES +
         P_STASH_BYTES + sizeof(PairBarriers) + 1024;
}
template <bool kSoft>
int match_pair_stages(int KB) {
  int stages = P_MAX_STAGES;
  while (stages > 0 && match_pair_smem_bytes<kSoft>(KB, stages) > 227 * 1024) --stages;
  return stages;
}

// ---------------------------------------------------------------------------------------------------------------
// Fragment-layout variant.  With thread == row every per-column constant has to be delivered to all 32 lanes of a
// warp: 4 bytes of shared-memory return bandwidth per score and plane, which is what bounds the two kernels above
// (DESIGN.md 3.1).  Here the accumulators are read with tcgen05.ld.16x256b, the mma.sync fragment layout: a thread
// owns FOUR rows (TMEM lanes t/4 + {0, 8, 16, 24} of its warp's quarter) and, of every 8-column group, the column
// pair 2 (t % 4) + {0, 1}.  One 8-byte LDS then serves 8 scores (4 rows x 2 columns) instead of 2, the 4 lanes of a
// quad and the two column slices of a row merge once, after the last tile (shuffles + shared memory).
//   Tiling, TMA and UMMA roles: as match_kernel<., 2> (256 rows per CTA, accumulator r = row tile r, 128x256x16 MMAs).
//   warps 0..15: row tile (w / 4) % 2, TMEM lane quarter w % 4, 128-column slice w / 8.
//   Argmax: a row has 8 tracks (4 quad lanes x 2 slices), a thread 4 of them, so the stash of "the scores of the
//   chunk that last raised the running maximum" (32 B per track, 64 KB per CTA) lives in an L2-resident workspace
//   slot indexed by %smid (one CTA per SM: the CTA needs > half of the SM's shared memory); stores are predicated,
//   plane-major and coalesced, and only the storing thread ever reads them back.
constexpr int FRAG_STASH_BYTES = EPI_WARPS * 32 * 4 * 32;   // 512 threads x 4 rows x 8 scores = 64 KB per slot
constexpr int FRAG_STASH_PLANE = EPI_WARPS * 32 * 8;        // one packed pair per thread
constexpr int FRAG_NO_RECORD = 0x40000000;                  // chunk column of a track that holds no record

constexpr int FRAG_THREADS = 640;   // 4 epilogue warpgroups + 1 producer warpgroup (TMA, UMMA, two idle warps)

template <bool kSoft, int RT>
__global__ void __launch_bounds__(FRAG_THREADS, 1)
match_frag_kernel(const __grid_constant__ CUtensorMap tmap_rows, const __grid_constant__ CUtensorMap tmap_cols,
                  const MatchParams p) {
  constexpr int SL = 4 / RT;                 // column slices per row (warps that share a row)
  constexpr int F_CS = BN / SL;              // columns per warp slice: 64 (RT = 1) or 128 (RT = 2)
  constexpr int AUX_BYTES = kSoft ? 4 * PLANE_BYTES : PLANE_BYTES;
  constexpr int W = 32;                      // columns per chunk
  constexpr int NP = W / 8;                  // packed pairs (8-column groups) per row and chunk

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                                    // [RT][KB] blocks of 128 rows x 64 k
  uint8_t* smem_b = smem_a + RT * p.KB * A_BLK_BYTES;
  uint8_t* smem_aux = smem_b + p.stages * B_STAGE_BYTES;   // per slot: [1/|m| x256 | x x256 | y x256 | z x256]
  Barriers* bars = reinterpret_cast<Barriers*>(smem_aux + AUX_SLOTS * AUX_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int row0 = blockIdx.x * (BM * RT);
  const int obj = frame_object(p, b);
  const int num_tiles = (p.M + BN - 1) / BN;

  if (warp == EPI_WARPS && lane == 0) {
    ptx::prefetch_tensormap(&tmap_rows);
    ptx::prefetch_tensormap(&tmap_cols);
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(&bars->full[s], 1);
      ptx::mbar_init(&bars->empty[s], 1);
    }
    ptx::mbar_init(&bars->a_full, 1);
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&bars->s_full[a], 1);
      ptx::mbar_init(&bars->s_free[a], EPI_WARPS / RT);   // one arrive per epilogue warp of the accumulator
    }
    for (int a = 0; a < AUX_SLOTS; ++a) {
      ptx::mbar_init(&bars->aux_full[a], 1);
      ptx::mbar_init(&bars->aux_empty[a], EPI_WARPS);
    }
    ptx::fence_mbar_init();
  }
  if (warp == EPI_WARPS + 1) {
    ptx::tmem_alloc(&bars->tmem_base, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp >= EPI_WARPS) {
    // producer warpgroup (warps 16..19): hand registers to the epilogue warpgroups
    ptx::setmaxnreg_dec<24>();
  }
  if (warp == EPI_WARPS) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(&bars->a_full, RT * p.KB * A_BLK_BYTES);
      for (int r = 0; r < RT; ++r)
        for (int kb = 0; kb < p.KB; ++kb)
          ptx::tma_load_3d(smem_a + (r * p.KB + kb) * A_BLK_BYTES, &tmap_rows, &bars->a_full, kb * BK,
                           row0 + r * BM, b);   // rows >= N are zero-filled by TMA
      int stage = 0;
      uint32_t phase = 0;
      const size_t plane = size_t(p.n_obj) * p.M;
      const float* sc_tab = p.scales + size_t(obj) * p.M;
      const float* xyz_tab = p.planes + size_t(obj) * p.M;
      for (int t = 0; t < num_tiles; ++t) {
        const int slot = t % AUX_SLOTS;
        const uint32_t use = uint32_t(t) / AUX_SLOTS;
        const uint32_t bytes = uint32_t(min(BN, p.M - t * BN)) * 4;   // M % 8 == 0: a multiple of 16
        ptx::mbar_wait_sleep(&bars->aux_empty[slot], (use & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&bars->aux_full[slot], kSoft ? 4 * bytes : bytes);
        uint8_t* aux = smem_aux + slot * AUX_BYTES;
        ptx::bulk_load_1d(aux, sc_tab + size_t(t) * BN, bytes, &bars->aux_full[slot]);
        if (kSoft) {
#pragma unroll
          for (int c = 0; c < 3; ++c)
            ptx::bulk_load_1d(aux + (c + 1) * PLANE_BYTES, xyz_tab + c * plane + size_t(t) * BN, bytes,
                              &bars->aux_full[slot]);
        }
        for (int kb = 0; kb < p.KB; ++kb) {
          ptx::mbar_wait_sleep(&bars->empty[stage], phase ^ 1);
          ptx::mbar_arrive_expect_tx(&bars->full[stage], B_STAGE_BYTES);
          ptx::tma_load_3d(smem_b + stage * B_STAGE_BYTES, &tmap_cols, &bars->full[stage], kb * BK, t * BN, obj);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == EPI_WARPS + 1) {
    // ============================== UMMA issuer ==============================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16_f32(BM, BN);
      ptx::mbar_wait(&bars->a_full, 0);
      int stage0 = 0;            // ring position of the tile's first K block
      uint32_t phase0 = 0;
      for (int t = 0; t < num_tiles; ++t) {
#pragma unroll
        for (int r = 0; r < RT; ++r) {
          const int acc = RT == 1 ? (t & 1) : r;
          const uint32_t use = RT == 1 ? uint32_t(t) >> 1 : uint32_t(t);
          ptx::mbar_wait_sleep(&bars->s_free[acc], (use & 1) ^ 1);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * BN;
          int stage = stage0;
          uint32_t phase = phase0;
          for (int kb = 0; kb < p.KB; ++kb) {
            if (r == 0) {      // the stages of this tile stay resident until the last row tile has used them
              ptx::mbar_wait_sleep(&bars->full[stage], phase);
              ptx::tc_fence_after();
            }
            const uint32_t a_addr = ptx::smem_u32(smem_a + (r * p.KB + kb) * A_BLK_BYTES);
            const uint32_t b_addr = ptx::smem_u32(smem_b + stage * B_STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              ptx::umma_bf16_ss(d_tmem, ptx::umma_desc_sw128_kmajor(a_addr + k * UMMA_K * 2),
                                ptx::umma_desc_sw128_kmajor(b_addr + k * UMMA_K * 2), idesc, (kb | k) != 0);
            }
            if (r == RT - 1) ptx::umma_commit(&bars->empty[stage]);  // frees the stage once these MMAs have read it
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
          ptx::umma_commit(&bars->s_full[acc]);   // accumulator tile complete
          if (r == RT - 1) { stage0 = stage; phase0 = phase; }
        }
      }
    }
  } else if (warp < EPI_WARPS) {
    ptx::setmaxnreg_inc<112>();
    // ============================== epilogue warps (fragment layout: 4 rows x column pairs per thread) ==========
    const int q = warp & 3;                  // TMEM lane quarter this warp may access (warp id % 4)
    const int rt = RT == 1 ? 0 : (warp >> 2) & 1;      // row tile
    const int sub = RT == 1 ? warp >> 2 : warp >> 3;   // column slice of every tile
    const int q4 = lane & 3;                 // column pair inside every 8-column group
    const int r8 = lane >> 2;                // rows q * 32 + r8 + 8 rr, rr = 0..3
    const int rbase = row0 + rt * BM + q * 32 + r8;
    float g[4];                              // exponent scale of row rr: t = (acc * 1/|m_j|) * g   (log2 units)
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      const int row = rbase + 8 * rr;
      g[rr] = kSoft && row < p.N ? p.gamma_log2e * p.rinv_rows[size_t(b) * p.N + row] : 0.f;
    }
    const uint32_t lane_lo0 = tmem_base + (uint32_t(q * 32) << 16) + sub * F_CS;
    if (ptx::smid() >= uint32_t(p.stash_slots)) __trap();
    // stash of row rr, pair i: plane rr * NP + i of the slot, 8 bytes per thread (coalesced 256-byte warp stores)
    uint8_t* stash = p.stash + size_t(ptx::smid()) * FRAG_STASH_BYTES + threadIdx.x * 8;

    float vmax[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // running maximum of the thread's track of row rr
    int vchk[4] = {0, 0, 0, 0};              // first column of the chunk that first reached it
    // SOFT: p = 2^(score * g) WITHOUT a reference exponent: |score * g| <= |gamma| log2(e) (a cosine times gamma), and
    // gadm_match_fwd admits |gamma| <= 40 only, so p stays within 2^+-58 and the sums within fp32 range -- the online
    // maximum of a flash-style softmax (a compare, a vote and a rescale per chunk) is not needed at all.
    uint64_t l2[4] = {0, 0, 0, 0}, ax2[4] = {0, 0, 0, 0}, ay2[4] = {0, 0, 0, 0}, az2[4] = {0, 0, 0, 0};

    for (int t = 0; t < num_tiles; ++t) {
      const int slot = t % AUX_SLOTS;
      const int acc = RT == 1 ? (t & 1) : rt;
      const uint32_t use = RT == 1 ? uint32_t(t) >> 1 : uint32_t(t);
      if (!(ptx::mbar_try_wait(&bars->aux_full[slot], (uint32_t(t) / AUX_SLOTS) & 1) &
            ptx::mbar_try_wait(&bars->s_full[acc], use & 1))) {
        ptx::mbar_wait_sleep(&bars->aux_full[slot], (uint32_t(t) / AUX_SLOTS) & 1);
        ptx::mbar_wait_sleep(&bars->s_full[acc], use & 1);
      }
      ptx::tc_fence_after();
      const uint32_t lane_lo = lane_lo0 + acc * BN, lane_hi = lane_lo + (16u << 16);
      const int ncols = min(BN, p.M - t * BN) - sub * F_CS;   // valid columns of this slice (may be <= 0)
      const uint32_t sc_addr = ptx::smem_u32(smem_aux + slot * AUX_BYTES) + (sub * F_CS + 2 * q4) * 4;
      const int col_base = t * BN + sub * F_CS;

      // One chunk of W columns starting at slice column col0; d0 / d1 hold the raw accumulators of rows
      // {0, 8} / {16, 24} (+ r8), c2 the column scales of the thread's pairs.  kGuard (ragged last tile): 8-column
      // groups at or beyond ncols are masked (M % 8 == 0: a group is valid or invalid as a whole).
      auto process = [&](auto& d0, auto& d1, const uint64_t (&c2)[NP], int col0, auto guard_tag) {
        constexpr bool kGuard = decltype(guard_tag)::value;
        const uint32_t sc = sc_addr + col0 * 4;
        uint64_t v[4][NP];
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          v[0][i] = ptx::fmul2(ptx::pack2(d0[4 * i + 0], d0[4 * i + 1]), c2[i]);
          v[1][i] = ptx::fmul2(ptx::pack2(d0[4 * i + 2], d0[4 * i + 3]), c2[i]);
          v[2][i] = ptx::fmul2(ptx::pack2(d1[4 * i + 0], d1[4 * i + 1]), c2[i]);
          v[3][i] = ptx::fmul2(ptx::pack2(d1[4 * i + 2], d1[4 * i + 3]), c2[i]);
          if (kGuard && col0 + 8 * i >= ncols) {                   // warp-uniform
            const uint64_t ninf2 = ptx::pack2f(-INFINITY, -INFINITY);
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) v[rr][i] = ninf2;
          }
        }
        if (kSoft) {
#pragma unroll
          for (int i = 0; i < NP; ++i) {
            if (kGuard && col0 + 8 * i >= ncols) break;            // stale constants behind the last group
#ifdef GADM_DBG_NOLDS
            const uint64_t X2 = ptx::pack2f(1.f + i, 1.f), Y2 = ptx::pack2f(2.f + i, 1.f), Z2 = ptx::pack2f(3.f + i, 1.f);
#else
            const uint64_t X2 = ptx::lds64(sc + PLANE_BYTES + i * 32);
            const uint64_t Y2 = ptx::lds64(sc + 2 * PLANE_BYTES + i * 32);
            const uint64_t Z2 = ptx::lds64(sc + 3 * PLANE_BYTES + i * 32);
#endif
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
#ifdef GADM_DBG_NOMUFU
              const uint64_t pp = ptx::fmul2(v[rr][i], ptx::pack2f(g[rr], g[rr]));
#else
              const uint64_t pp = ptx::ex2_2(ptx::fmul2(v[rr][i], ptx::pack2f(g[rr], g[rr])));   // p = 2^(score * g)
#endif
              l2[rr] = ptx::fadd2(l2[rr], pp);
              ax2[rr] = ptx::ffma2(pp, X2, ax2[rr]);
              ay2[rr] = ptx::ffma2(pp, Y2, ay2[rr]);
              az2[rr] = ptx::ffma2(pp, Z2, az2[rr]);
            }
          }
        }
#ifdef GADM_DBG_NOMAX
        for (int rr = 0; rr < 4; ++rr) vmax[rr] += __uint_as_float(uint32_t(v[rr][0] ^ v[rr][1] ^ v[rr][2] ^ v[rr][3]));
#else
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          float f[2 * NP];
#pragma unroll
          for (int i = 0; i < NP; ++i) ptx::unpack2f(v[rr][i], f[2 * i], f[2 * i + 1]);
          const float a0 = ptx::fmax3(f[0], f[1], f[2]), a1 = ptx::fmax3(f[3], f[4], f[5]);
          const float gm = ptx::fmax3(a0, a1, fmaxf(f[6], f[7]));
          // strict: an equal value in a later chunk never displaces the first maximal index
          const bool up = gm > vmax[rr];
#pragma unroll
#ifndef GADM_DBG_NOSTASH
          for (int i = 0; i < NP; ++i) ptx::stg_pred8(up, stash + (rr * NP + i) * FRAG_STASH_PLANE, v[rr][i]);
#endif
          vchk[rr] = up ? col_base + col0 : vchk[rr];
          vmax[rr] = up ? gm : vmax[rr];
        }
#endif
      };
      using guard_off = std::integral_constant<bool, false>;
      using guard_on = std::integral_constant<bool, true>;

#ifndef GADM_DBG_NOEPI
#pragma unroll
      for (int c = 0; c < F_CS / W; ++c) {
        const int nv = ncols - c * W;        // valid columns from this chunk on
        if (nv <= 0) break;
        uint32_t d0[4 * NP], d1[4 * NP];
        uint64_t c2[NP];
        ptx::tmem_ld_frag(lane_lo + c * W, d0);
        ptx::tmem_ld_frag(lane_hi + c * W, d1);
#pragma unroll
#ifdef GADM_DBG_NOLDS
        for (int i = 0; i < NP; ++i) c2[i] = ptx::pack2f(1.f + i, 1.f);
#else
        for (int i = 0; i < NP; ++i) c2[i] = ptx::lds64(sc_addr + (c * W + 8 * i) * 4);   // overlaps the TMEM latency
#endif
        ptx::tmem_ld_wait();
#ifdef GADM_DBG_LDONLY
        vchk[0] += d0[0] + d1[0] + d0[15] + d1[15];
#else
        if (nv >= W) process(d0, d1, c2, c * W, guard_off{});
        else process(d0, d1, c2, c * W, guard_on{});
#endif
      }
#endif
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(&bars->s_free[acc]);
        ptx::mbar_arrive(&bars->aux_empty[slot]);
      }
      // The 4 tracks of a quad belong to the same rows.  Each one alone would raise its running maximum (and store
      // a stash entry: a wavefront of the SM's data pipe per store instruction with any lane on) H(n) ~ 5 times;
      // sharing the quad's maximum after tiles 0, 1, 3, 7, ... leaves ~ln(2) updates per ROW between two
      // exchanges.  A track that adopts a larger maximum than its own gives up its record: its chunk becomes a
      // sentinel that loses every tie, which is right -- the holder of that maximum sits at an earlier column.
      if ((t & (t + 1)) == 0) {
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          float m = fmaxf(vmax[rr], __shfl_xor_sync(0xffffffffu, vmax[rr], 1));
          m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
          if (vmax[rr] < m) { vmax[rr] = m; vchk[rr] = FRAG_NO_RECORD; }
        }
      }
    }

    // ---- first maximal index of every track: look it up in the stashed chunk (own stores, read back through L2)
    int vidx[4];
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      vidx[rr] = FRAG_NO_RECORD;
      if (vmax[rr] > -INFINITY && vchk[rr] != FRAG_NO_RECORD) {
        int j_first = 2 * NP - 1;
#pragma unroll
        for (int i = NP - 1; i >= 0; --i) {
          float lo, hi;
          ptx::unpack2f(ptx::ldg_cg64(stash + (rr * NP + i) * FRAG_STASH_PLANE), lo, hi);
          if (hi == vmax[rr]) j_first = 2 * i + 1;
          if (lo == vmax[rr]) j_first = 2 * i;
        }
        vidx[rr] = vchk[rr] + 8 * (j_first >> 1) + 2 * q4 + (j_first & 1);
      }
    }

    // ---- merge the 4 tracks of a quad (butterfly: afterwards every lane of the quad holds all 4 rows)
    float lsum[4], ax[4], ay[4], az[4];
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      float e, o;
      ptx::unpack2f(l2[rr], e, o); lsum[rr] = e + o;
      ptx::unpack2f(ax2[rr], e, o); ax[rr] = e + o;
      ptx::unpack2f(ay2[rr], e, o); ay[rr] = e + o;
      ptx::unpack2f(az2[rr], e, o); az[rr] = e + o;
#pragma unroll
      for (int off = 1; off <= 2; off <<= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, vmax[rr], off);
        const int oi = __shfl_xor_sync(0xffffffffu, vidx[rr], off);
        if (ov > vmax[rr] || (ov == vmax[rr] && oi < vidx[rr])) { vmax[rr] = ov; vidx[rr] = oi; }
        if (kSoft) {
          lsum[rr] += __shfl_xor_sync(0xffffffffu, lsum[rr], off);
          ax[rr] += __shfl_xor_sync(0xffffffffu, ax[rr], off);
          ay[rr] += __shfl_xor_sync(0xffffffffu, ay[rr], off);
          az[rr] += __shfl_xor_sync(0xffffffffu, az[rr], off);
        }
      }
    }
    // lane q4 of the quad finishes row rr == q4
    float my_vmax = vmax[0], my_l = lsum[0], my_ax = ax[0], my_ay = ay[0], my_az = az[0], my_g = g[0];
    int my_vidx = vidx[0];
#pragma unroll
    for (int rr = 1; rr < 4; ++rr)
      if (q4 == rr) {
        my_vmax = vmax[rr]; my_vidx = vidx[rr]; my_l = lsum[rr];
        my_ax = ax[rr]; my_ay = ay[rr]; my_az = az[rr]; my_g = g[rr];
      }
    const int row_in_tile = q * 32 + r8 + 8 * q4;
    const int row = row0 + rt * BM + row_in_tile;
    const bool row_ok = row < p.N;
    const size_t grow = size_t(b) * p.N + (row_ok ? row : 0);

    // ---- merge the two column slices of every row through shared memory (the row tile's own A blocks: every MMA
    // that reads them has completed, this warp has seen the last s_full of its accumulator)
    float* xch = reinterpret_cast<float*>(smem_a + rt * p.KB * A_BLK_BYTES);   // (SL - 1) * 128 * 32 B <= 16 KB
    if (sub > 0) {
      float* x = xch + ((sub - 1) * BM + row_in_tile) * 8;
      x[0] = my_vmax; x[1] = __int_as_float(my_vidx); x[2] = my_l;
      x[3] = my_ax; x[4] = my_ay; x[5] = my_az;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
    if (sub == 0 && row_ok) {
#pragma unroll
      for (int s2 = 0; s2 < SL - 1; ++s2) {
        const float* x = xch + (s2 * BM + row_in_tile) * 8;
        const float v1 = x[0];
        const int i1 = __float_as_int(x[1]);
        if (v1 > my_vmax || (v1 == my_vmax && i1 < my_vidx)) { my_vmax = v1; my_vidx = i1; }
        my_l += x[2]; my_ax += x[3]; my_ay += x[4]; my_az += x[5];
      }
      const float rs = p.rinv_rows[grow];
      const bool keep = p.mask == nullptr || p.mask[grow] != 0;
      float best = my_vmax * rs;
      int64_t best_idx = my_vidx;
      if (p.pad_mode != GADM_PAD_NONE) {
        const float ps = p.pad_sim[grow];
        if (ps > best) { best = ps; best_idx = p.M; }  // pad column is the last one: wins only if strictly larger
      }
      p.idx[grow] = keep ? best_idx : int64_t(-1);
      p.max_sim[grow] = keep ? best : 0.f;
      if (kSoft) {
        const float inv = 1.f / my_l;
        p.weight[grow] = keep ? ptx::ex2_approx(my_vmax * my_g) * inv : 0.f;  // softmax value at the maximum
        p.soft_xyz[grow * 3 + 0] = keep ? my_ax * inv : 0.f;
        p.soft_xyz[grow * 3 + 1] = keep ? my_ay * inv : 0.f;
        p.soft_xyz[grow * 3 + 2] = keep ? my_az * inv : 0.f;
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == EPI_WARPS + 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <bool kSoft>
size_t match_frag_smem_bytes(int RT, int KB, int stages) {
  return size_t(RT) * KB * A_BLK_BYTES + size_t(stages) * B_STAGE_BYTES + AUX_SLOTS * (kSoft ? 4 : 1) * PLANE_BYTES +
         sizeof(Barriers) + 1024;
}
template <bool kSoft>
int match_frag_stages(int RT, int KB) {
  int stages = MAX_STAGES;
  while (stages > 0 && match_frag_smem_bytes<kSoft>(RT, KB, stages) > 227 * 1024) --stages;
  return stages;
}

// ---------------------------------------------------------------------------------------------------------------
// Alternating variant of match_kernel<ARGMAX, 2>.  There, accumulator r belongs to 8 fixed epilogue warps, so an
// accumulator's MMAs wait for its own epilogue and its epilogue warps idle while it is refilled: the period of a
// model tile is T_mma + E_8warps.  Here ALL 16 epilogue warps drain accumulator 0 (row tile 0) while the tensor
// core fills accumulator 1 (row tile 1) with the same model tile, then swap: the period is 2 max(T_mma, E_16warps)
// and the epilogue never idles -- the double buffering of RT = 1 with the halved L2 operand traffic of RT = 2.
// A thread owns one row of each row tile (TMEM lane q * 32 + lane) and the 64-column slice w / 4 of every tile.
// Two stash entries per thread (32 KB per CTA) do not fit beside the operands: the stash lives in the per-SM
// workspace slot (predicated, coalesced STG.128; only the storing thread reads it back).
// kUnit (GADM_MATCH_ARGMAX_UNIT, operands from GADM_OPERAND_BF16N): the column norms are taken as 1, the epilogue
// needs no per-column constant at all -- no aux ring, no LDS, no multiply.
// kPrune (GADM_MATCH_ARGMAX_BF16N, same operands): exact scores, but a 32-column chunk is skipped -- no scale LDS, no
// multiply, no stash -- when max(raw, 0) * (1 + 2^-8) cannot beat the running maximum of any row of the warp: every
// column scale of BF16N operands is <= 1 / (1 - 2^-9), products round monotonically, so nothing is ever missed.
template <bool kUnit, bool kPrune>
__global__ void __launch_bounds__(NUM_THREADS, 1)
match_alt_kernel(const __grid_constant__ CUtensorMap tmap_rows, const __grid_constant__ CUtensorMap tmap_cols,
                 const MatchParams p) {
  constexpr int RT = 2;
  constexpr int AUX_BYTES = PLANE_BYTES;
  constexpr int SL = 4;                      // column slices per row
  constexpr int CS = BN / SL;                // 64 columns per slice

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                                    // [RT][KB] blocks of 128 rows x 64 k
  uint8_t* smem_b = smem_a + RT * p.KB * A_BLK_BYTES;
  uint8_t* smem_aux = smem_b + p.stages * B_STAGE_BYTES;   // per slot: 1/|m| x256
  float* smem_xmax = reinterpret_cast<float*>(smem_aux + AUX_SLOTS * AUX_BYTES);   // [RT][SL][128] running maxima
  Barriers* bars = reinterpret_cast<Barriers*>(smem_xmax + RT * SL * BM);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int row0 = blockIdx.x * (BM * RT);
  const int obj = frame_object(p, b);
  const int num_tiles = (p.M + BN - 1) / BN;

  if (warp == EPI_WARPS && lane == 0) {
    ptx::prefetch_tensormap(&tmap_rows);
    ptx::prefetch_tensormap(&tmap_cols);
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(&bars->full[s], 1);
      ptx::mbar_init(&bars->empty[s], 1);
    }
    ptx::mbar_init(&bars->a_full, 1);
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&bars->s_full[a], 1);
      ptx::mbar_init(&bars->s_free[a], EPI_WARPS);   // every epilogue warp drains every accumulator
    }
    for (int a = 0; a < AUX_SLOTS; ++a) {
      ptx::mbar_init(&bars->aux_full[a], 1);
      ptx::mbar_init(&bars->aux_empty[a], EPI_WARPS);
    }
    ptx::fence_mbar_init();
  }
  if (warp == EPI_WARPS + 1) {
    ptx::tmem_alloc(&bars->tmem_base, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == EPI_WARPS) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(&bars->a_full, RT * p.KB * A_BLK_BYTES);
      for (int r = 0; r < RT; ++r)
        for (int kb = 0; kb < p.KB; ++kb)
          ptx::tma_load_3d(smem_a + (r * p.KB + kb) * A_BLK_BYTES, &tmap_rows, &bars->a_full, kb * BK,
                           row0 + r * BM, b);   // rows >= N are zero-filled by TMA
      int stage = 0;
      uint32_t phase = 0;
      const float* sc_tab = p.scales + size_t(obj) * p.M;
      for (int t = 0; t < num_tiles; ++t) {
        const int slot = t % AUX_SLOTS;
        const uint32_t use = uint32_t(t) / AUX_SLOTS;
        const uint32_t bytes = uint32_t(min(BN, p.M - t * BN)) * 4;   // M % 8 == 0: a multiple of 16
        if (!kUnit) {
          ptx::mbar_wait_sleep(&bars->aux_empty[slot], (use & 1) ^ 1);
          ptx::mbar_arrive_expect_tx(&bars->aux_full[slot], bytes);
          ptx::bulk_load_1d(smem_aux + slot * AUX_BYTES, sc_tab + size_t(t) * BN, bytes, &bars->aux_full[slot]);
        }
        for (int kb = 0; kb < p.KB; ++kb) {
          ptx::mbar_wait_sleep(&bars->empty[stage], phase ^ 1);
          ptx::mbar_arrive_expect_tx(&bars->full[stage], B_STAGE_BYTES);
          ptx::tma_load_3d(smem_b + stage * B_STAGE_BYTES, &tmap_cols, &bars->full[stage], kb * BK, t * BN, obj);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == EPI_WARPS + 1) {
    // ============================== UMMA issuer ==============================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16_f32(BM, BN);
      ptx::mbar_wait(&bars->a_full, 0);
      int stage0 = 0;            // ring position of the tile's first K block
      uint32_t phase0 = 0;
      for (int t = 0; t < num_tiles; ++t) {
#pragma unroll
        for (int r = 0; r < RT; ++r) {
          ptx::mbar_wait_sleep(&bars->s_free[r], (uint32_t(t) & 1) ^ 1);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + r * BN;
          int stage = stage0;
          uint32_t phase = phase0;
          for (int kb = 0; kb < p.KB; ++kb) {
            if (r == 0) {      // the stages of this tile stay resident until the last row tile has used them
              ptx::mbar_wait_sleep(&bars->full[stage], phase);
              ptx::tc_fence_after();
            }
            const uint32_t a_addr = ptx::smem_u32(smem_a + (r * p.KB + kb) * A_BLK_BYTES);
            const uint32_t b_addr = ptx::smem_u32(smem_b + stage * B_STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              ptx::umma_bf16_ss(d_tmem, ptx::umma_desc_sw128_kmajor(a_addr + k * UMMA_K * 2),
                                ptx::umma_desc_sw128_kmajor(b_addr + k * UMMA_K * 2), idesc, (kb | k) != 0);
            }
            if (r == RT - 1) ptx::umma_commit(&bars->empty[stage]);  // frees the stage once these MMAs have read it
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
          ptx::umma_commit(&bars->s_full[r]);     // accumulator tile complete
          if (r == RT - 1) { stage0 = stage; phase0 = phase; }
        }
      }
    }
  } else {
    // ============================== epilogue warps (thread == one row of each row tile) ==============
    const int q = warp & 3;                          // TMEM lane quarter this warp may access (warp id % 4)
    const int sub = warp >> 2;                       // 64-column slice of every tile
    const int row_in_tile = q * 32 + lane;
    const uint32_t lane_base = tmem_base + (uint32_t(q * 32) << 16) + sub * CS;
    if (ptx::smid() >= uint32_t(p.stash_slots)) __trap();
    // stash entry of row tile r, float4 k: + (2 r + k) * 8192 (plane-major: coalesced 512-byte warp stores)
    uint8_t* stash = p.stash + size_t(ptx::smid()) * FRAG_STASH_BYTES + threadIdx.x * 16;

    float vmax[RT] = {-INFINITY, -INFINITY};   // running maximum of this thread's slice of its row of row tile r
    int vgrp[RT] = {0, 0};                     // first column of the 8-column group that first reached it

    for (int t = 0; t < num_tiles; ++t) {
      const int slot = t % AUX_SLOTS;
      const int ncols = min(BN, p.M - t * BN) - sub * CS;   // valid columns of this slice (may be <= 0)
      const uint32_t sc_addr = ptx::smem_u32(smem_aux + slot * AUX_BYTES) + sub * CS * 4;
      const int col_base = t * BN + sub * CS;
#pragma unroll
      for (int r = 0; r < RT; ++r) {
        if (!((kUnit || r > 0 || ptx::mbar_try_wait(&bars->aux_full[slot], (uint32_t(t) / AUX_SLOTS) & 1)) &
              ptx::mbar_try_wait(&bars->s_full[r], uint32_t(t) & 1))) {
          if (!kUnit && r == 0) ptx::mbar_wait_sleep(&bars->aux_full[slot], (uint32_t(t) / AUX_SLOTS) & 1);
          ptx::mbar_wait_sleep(&bars->s_full[r], uint32_t(t) & 1);
        }
        ptx::tc_fence_after();
        const uint32_t s_tmem = lane_base + r * BN;

        // one chunk of 32 columns starting at slice column col0 (see match_kernel)
        auto process = [&](uint32_t (&d)[32], int col0, auto guard_tag) {
          constexpr bool kGuard = decltype(guard_tag)::value;
          if (kPrune) {
            float a[11];
#pragma unroll
            for (int j = 0; j < 10; ++j)
              a[j] = ptx::fmax3(__uint_as_float(d[3 * j]), __uint_as_float(d[3 * j + 1]), __uint_as_float(d[3 * j + 2]));
            a[10] = fmaxf(__uint_as_float(d[30]), __uint_as_float(d[31]));
            const float b0 = ptx::fmax3(a[0], a[1], a[2]), b1 = ptx::fmax3(a[3], a[4], a[5]);
            const float b2 = ptx::fmax3(a[6], a[7], a[8]), b3 = fmaxf(a[9], a[10]);
            const float bound = fmaxf(ptx::fmax3(b0, b1, fmaxf(b2, b3)), 0.f) * 1.00390625f;   // * (1 + 2^-8)
            if (!__any_sync(0xffffffffu, bound > vmax[r])) return;
          }
          const uint32_t sc = sc_addr + col0 * 4;
          uint64_t v[16];
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            if (kUnit) {
              v[j4 * 2 + 0] = ptx::pack2(d[j4 * 4 + 0], d[j4 * 4 + 1]);
              v[j4 * 2 + 1] = ptx::pack2(d[j4 * 4 + 2], d[j4 * 4 + 3]);
            } else {
              const float4 cm = ptx::lds128(sc + j4 * 16);
              v[j4 * 2 + 0] = ptx::fmul2(ptx::pack2(d[j4 * 4 + 0], d[j4 * 4 + 1]), ptx::pack2f(cm.x, cm.y));
              v[j4 * 2 + 1] = ptx::fmul2(ptx::pack2(d[j4 * 4 + 2], d[j4 * 4 + 3]), ptx::pack2f(cm.z, cm.w));
            }
          }
          if (kGuard) {  // TMA zero-fills columns >= M and the stale scales behind them are meaningless
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float lo, hi;
              ptx::unpack2f(v[j], lo, hi);
              if (col0 + 2 * j >= ncols) lo = -INFINITY;
              if (col0 + 2 * j + 1 >= ncols) hi = -INFINITY;
              v[j] = ptx::pack2f(lo, hi);
            }
          }
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            float f[GRP];
#pragma unroll
            for (int j = 0; j < GRP / 2; ++j) ptx::unpack2f(v[h * 4 + j], f[2 * j], f[2 * j + 1]);
            const float a0 = ptx::fmax3(f[0], f[1], f[2]), a1 = ptx::fmax3(f[3], f[4], f[5]);
            const float gm = ptx::fmax3(a0, a1, fmaxf(f[6], f[7]));
            // strict: an equal value in a later group never displaces the first maximal index
            const bool up = gm > vmax[r];
            ptx::stg_pred32(up, stash + r * 2 * 8192, v[h * 4 + 0], v[h * 4 + 1], v[h * 4 + 2], v[h * 4 + 3]);
            vgrp[r] = up ? col_base + col0 + h * GRP : vgrp[r];
            vmax[r] = up ? gm : vmax[r];
          }
        };
        using guard_off = std::integral_constant<bool, false>;
        using guard_on = std::integral_constant<bool, true>;

        if (ncols > 0) {
          uint32_t ra[32], rb[32];
          ptx::tmem_ld_32x32(s_tmem, ra);
          ptx::tmem_ld_32x32(s_tmem + 32, rb);
          ptx::tmem_ld_wait();
          if (ncols >= 32) process(ra, 0, guard_off{});
          else process(ra, 0, guard_on{});         // ragged last tile
          if (ncols > 32) {
            if (ncols >= 64) process(rb, 32, guard_off{});
            else process(rb, 32, guard_on{});
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          ptx::mbar_arrive(&bars->s_free[r]);
          if (!kUnit && r == RT - 1) ptx::mbar_arrive(&bars->aux_empty[slot]);
        }
      }
      // A row has 4 tracks (one per column slice, in 4 different warps).  Alone, each raises its running maximum
      // (two stash stores, wavefronts of the data pipe the MMAs saturate) ~H(n) times; after tiles 0, 1, 3, 7, 15
      // the slices publish their maxima and adopt the row's: a track that adopts a larger maximum than its own
      // gives up its record (its index becomes a sentinel that loses every tie -- the holder sits at an earlier
      // column), and from then on only values above the ROW's maximum so far are recorded.
      if ((t & (t + 1)) == 0 && t < 16 && t + 1 < num_tiles) {
#pragma unroll
        for (int r = 0; r < RT; ++r) smem_xmax[(r * SL + sub) * BM + row_in_tile] = vmax[r];
        asm volatile("bar.sync 2, %0;" ::"n"(EPI_WARPS * 32) : "memory");
#pragma unroll
        for (int r = 0; r < RT; ++r) {
          float m = vmax[r];
#pragma unroll
          for (int s2 = 0; s2 < SL; ++s2) m = fmaxf(m, smem_xmax[(r * SL + s2) * BM + row_in_tile]);
          if (vmax[r] < m) { vmax[r] = m; vgrp[r] = FRAG_NO_RECORD; }
        }
        asm volatile("bar.sync 2, %0;" ::"n"(EPI_WARPS * 32) : "memory");
      }
    }

    // ---- per row tile: first maximal index of this slice from the stash (own stores, read back through L2), then
    // the merge of the 4 column slices through shared memory (the row tile's own A blocks: all MMAs have completed)
    int vidx[RT];
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      vidx[r] = FRAG_NO_RECORD;
      if (vmax[r] > -INFINITY && vgrp[r] != FRAG_NO_RECORD) {
        int j_first = GRP - 1;
#pragma unroll
        for (int k = GRP / 4 - 1; k >= 0; --k) {
          const float4 sv = ptx::ldg_cg128(stash + (r * 2 + k) * 8192);
          if (sv.w == vmax[r]) j_first = 4 * k + 3;
          if (sv.z == vmax[r]) j_first = 4 * k + 2;
          if (sv.y == vmax[r]) j_first = 4 * k + 1;
          if (sv.x == vmax[r]) j_first = 4 * k + 0;
        }
        vidx[r] = vgrp[r] + j_first;
      }
      if (sub > 0) {
        float* x = reinterpret_cast<float*>(smem_a + r * p.KB * A_BLK_BYTES) + ((sub - 1) * BM + row_in_tile) * 2;
        x[0] = vmax[r]; x[1] = __int_as_float(vidx[r]);
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
    if (sub == 0) {
#pragma unroll
      for (int r = 0; r < RT; ++r) {
        const int row = row0 + r * BM + row_in_tile;
        if (row >= p.N) continue;
        const size_t grow = size_t(b) * p.N + row;
        float vm = vmax[r];
        int vi = vidx[r];
#pragma unroll
        for (int s2 = 0; s2 < SL - 1; ++s2) {
          const float* x = reinterpret_cast<const float*>(smem_a + r * p.KB * A_BLK_BYTES) + (s2 * BM + row_in_tile) * 2;
          const float v1 = x[0];
          const int i1 = __float_as_int(x[1]);
          if (v1 > vm || (v1 == vm && i1 < vi)) { vm = v1; vi = i1; }
        }
        // kUnit searched with unit column norms; the winner's similarity is reported with its true scale
        if (kUnit) vm *= p.scales[size_t(obj) * p.M + vi];
        const bool keep = p.mask == nullptr || p.mask[grow] != 0;
        float best = vm * p.rinv_rows[grow];
        int64_t best_idx = vi;
        if (p.pad_mode != GADM_PAD_NONE) {
          const float ps = p.pad_sim[grow];
          if (ps > best) { best = ps; best_idx = p.M; }  // pad column is the last one: wins only if strictly larger
        }
        p.idx[grow] = keep ? best_idx : int64_t(-1);
        p.max_sim[grow] = keep ? best : 0.f;
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == EPI_WARPS + 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

inline size_t match_alt_smem_bytes(int KB, int stages) {
  return size_t(2) * KB * A_BLK_BYTES + size_t(stages) * B_STAGE_BYTES + AUX_SLOTS * PLANE_BYTES + 2 * 4 * BM * 4 +
         sizeof(Barriers) + 1024;
}
inline int match_alt_stages(int KB) {
  int stages = MAX_STAGES;
  while (stages > 0 && match_alt_smem_bytes(KB, stages) > 227 * 1024) --stages;
  return stages;
}

// ---------------------------------------------------------------------------------------------------------------
// TMEM-resident A.  With two row tiles per CTA the UMMA operand reads (A 4 KB + B 8 KB per 128x256x16 MMA) and the TMA
// writes of the model tile already fill the SM's shared-memory data pipe.  Here the CTA's own 256 rows
// are written into TENSOR MEMORY once (tcgen05.st by the epilogue warps, straight from global memory) and every MMA
// takes its A operand from there (tcgen05.mma [d], [a], b-desc): the MMAs read only B from shared memory, A needs no
// shared memory at all.  TMEM budget: 2 x K'/2 columns of A (K' <= 128) + two accumulators of 192 columns = 512, so
// the model tile is 192 vertices (128x192x16 MMAs).  Everything else as match_alt_kernel.
constexpr int TBN = 192;                        // model vertices per tile
constexpr int TB_STAGE_BYTES = TBN * BK * 2;    // 24 KB
constexpr int TA_COL0 = 2 * TBN;                // first TMEM column of the A operands

template <bool kUnit>
__global__ void __launch_bounds__(NUM_THREADS, 1)
match_ta_kernel(const __grid_constant__ CUtensorMap tmap_rows, const __grid_constant__ CUtensorMap tmap_cols,
                 const MatchParams p) {
  constexpr int RT = 2;
  constexpr int AUX_BYTES = PLANE_BYTES;
  constexpr int SL = 4;                      // column slices per row
  constexpr int CS = TBN / SL;               // 48 columns per slice: a 32-column and a 16-column chunk

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_b = smem;
  uint8_t* smem_aux = smem_b + p.stages * TB_STAGE_BYTES;  // per slot: 1/|m| x192
  float* smem_xmax = reinterpret_cast<float*>(smem_aux + AUX_SLOTS * AUX_BYTES);   // [RT][SL][128] running maxima
  float* smem_xch = smem_xmax + RT * SL * BM;                                      // [RT][SL - 1][128][2] slice merge
  Barriers* bars = reinterpret_cast<Barriers*>(smem_xch + RT * (SL - 1) * BM * 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int row0 = blockIdx.x * (BM * RT);
  const int obj = frame_object(p, b);
  const int num_tiles = (p.M + TBN - 1) / TBN;

  if (warp == EPI_WARPS && lane == 0) {
    ptx::prefetch_tensormap(&tmap_rows);
    ptx::prefetch_tensormap(&tmap_cols);
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(&bars->full[s], 1);
      ptx::mbar_init(&bars->empty[s], 1);
    }
    ptx::mbar_init(&bars->a_full, EPI_WARPS);      // every epilogue warp writes its part of A into TMEM
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&bars->s_full[a], 1);
      ptx::mbar_init(&bars->s_free[a], EPI_WARPS);   // every epilogue warp drains every accumulator
    }
    for (int a = 0; a < AUX_SLOTS; ++a) {
      ptx::mbar_init(&bars->aux_full[a], 1);
      ptx::mbar_init(&bars->aux_empty[a], EPI_WARPS);
    }
    ptx::fence_mbar_init();
  }
  if (warp == EPI_WARPS + 1) {
    ptx::tmem_alloc(&bars->tmem_base, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == EPI_WARPS) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const float* sc_tab = p.scales + size_t(obj) * p.M;
      for (int t = 0; t < num_tiles; ++t) {
        const int slot = t % AUX_SLOTS;
        const uint32_t use = uint32_t(t) / AUX_SLOTS;
        const uint32_t bytes = uint32_t(min(TBN, p.M - t * TBN)) * 4;   // M % 8 == 0: a multiple of 16
        if (!kUnit) {
          ptx::mbar_wait_sleep(&bars->aux_empty[slot], (use & 1) ^ 1);
          ptx::mbar_arrive_expect_tx(&bars->aux_full[slot], bytes);
          ptx::bulk_load_1d(smem_aux + slot * AUX_BYTES, sc_tab + size_t(t) * TBN, bytes, &bars->aux_full[slot]);
        }
        for (int kb = 0; kb < p.KB; ++kb) {
          ptx::mbar_wait_sleep(&bars->empty[stage], phase ^ 1);
          ptx::mbar_arrive_expect_tx(&bars->full[stage], TB_STAGE_BYTES);
          ptx::tma_load_3d(smem_b + stage * TB_STAGE_BYTES, &tmap_cols, &bars->full[stage], kb * BK, t * TBN, obj);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == EPI_WARPS + 1) {
    // ============================== UMMA issuer ==============================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16_f32(BM, TBN);
      ptx::mbar_wait(&bars->a_full, 0);
      ptx::tc_fence_after();
      int stage0 = 0;            // ring position of the tile's first K block
      uint32_t phase0 = 0;
      for (int t = 0; t < num_tiles; ++t) {
#pragma unroll
        for (int r = 0; r < RT; ++r) {
          ptx::mbar_wait_sleep(&bars->s_free[r], (uint32_t(t) & 1) ^ 1);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + r * TBN;
          const uint32_t a_tmem = tmem_base + TA_COL0 + r * (p.KB * (BK / 2));   // K'/2 columns per row tile
          int stage = stage0;
          uint32_t phase = phase0;
          for (int kb = 0; kb < p.KB; ++kb) {
            if (r == 0) {      // the stages of this tile stay resident until the last row tile has used them
              ptx::mbar_wait_sleep(&bars->full[stage], phase);
              ptx::tc_fence_after();
            }
            const uint32_t b_addr = ptx::smem_u32(smem_b + stage * TB_STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {     // 16 k = 8 TMEM columns of A
              ptx::umma_f16_ts(d_tmem, a_tmem + (kb * (BK / UMMA_K) + k) * (UMMA_K / 2),
                               ptx::umma_desc_sw128_kmajor(b_addr + k * UMMA_K * 2), idesc, (kb | k) != 0);
            }
            if (r == RT - 1) ptx::umma_commit(&bars->empty[stage]);  // frees the stage once these MMAs have read it
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
          ptx::umma_commit(&bars->s_full[r]);     // accumulator tile complete
          if (r == RT - 1) { stage0 = stage; phase0 = phase; }
        }
      }
    }
  } else {
    // ============================== epilogue warps (thread == one row of each row tile) ==============
    const int q = warp & 3;                          // TMEM lane quarter this warp may access (warp id % 4)
    const int sub = warp >> 2;                       // 48-column slice of every tile
    const int row_in_tile = q * 32 + lane;
    const uint32_t lane_base = tmem_base + (uint32_t(q * 32) << 16) + sub * CS;
    {
      // A operand: this thread's row of both row tiles, K range [sub K'/4, (sub + 1) K'/4), from global memory into
      // TMEM (lane = row, one 32-bit column = two consecutive k: the layout tcgen05.mma reads an A operand in)
      const int kq = p.KB * (BK / 4);                // bf16 elements per thread and row tile: 16 (K' = 64) or 32
#pragma unroll
      for (int r = 0; r < RT; ++r) {
        const int row = row0 + r * BM + row_in_tile;
        const uint4* src = reinterpret_cast<const uint4*>(
            static_cast<const uint8_t*>(p.rows_ptr) + ((size_t(b) * p.N + (row < p.N ? row : 0)) * (p.KB * BK) + sub * kq) * 2);
        const uint32_t dst = tmem_base + (uint32_t(q * 32) << 16) + TA_COL0 + r * (p.KB * (BK / 2)) + sub * (kq / 2);
        uint4 v0 = make_uint4(0, 0, 0, 0), v1 = v0, v2 = v0, v3 = v0;
        if (row < p.N) {
          v0 = src[0]; v1 = src[1];
          if (kq == 32) { v2 = src[2]; v3 = src[3]; }
        }
        if (kq == 32) {
          const uint32_t w[16] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w, v3.x, v3.y, v3.z, v3.w};
          ptx::tmem_st_32x16(dst, w);
        } else {
          const uint32_t w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
          ptx::tmem_st_32x8(dst, w);
        }
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bars->a_full);
    }
    if (ptx::smid() >= uint32_t(p.stash_slots)) __trap();
    // stash entry of row tile r, float4 k: + (2 r + k) * 8192 (plane-major: coalesced 512-byte warp stores)
    uint8_t* stash = p.stash + size_t(ptx::smid()) * FRAG_STASH_BYTES + threadIdx.x * 16;

    float vmax[RT] = {-INFINITY, -INFINITY};   // running maximum of this thread's slice of its row of row tile r
    int vgrp[RT] = {0, 0};                     // first column of the 8-column group that first reached it

    for (int t = 0; t < num_tiles; ++t) {
      const int slot = t % AUX_SLOTS;
      const int ncols = min(TBN, p.M - t * TBN) - sub * CS;   // valid columns of this slice (may be <= 0)
      const uint32_t sc_addr = ptx::smem_u32(smem_aux + slot * AUX_BYTES) + sub * CS * 4;
      const int col_base = t * TBN + sub * CS;
#pragma unroll
      for (int r = 0; r < RT; ++r) {
        if (!((kUnit || r > 0 || ptx::mbar_try_wait(&bars->aux_full[slot], (uint32_t(t) / AUX_SLOTS) & 1)) &
              ptx::mbar_try_wait(&bars->s_full[r], uint32_t(t) & 1))) {
          if (!kUnit && r == 0) ptx::mbar_wait_sleep(&bars->aux_full[slot], (uint32_t(t) / AUX_SLOTS) & 1);
          ptx::mbar_wait_sleep(&bars->s_full[r], uint32_t(t) & 1);
        }
        ptx::tc_fence_after();
        const uint32_t s_tmem = lane_base + r * TBN;

        // one chunk of W = 32 or 16 columns starting at slice column col0 (see match_kernel)
        auto process = [&](auto& d, int col0, auto guard_tag) {
          constexpr int W = int(sizeof(d) / sizeof(d[0]));
          constexpr bool kGuard = decltype(guard_tag)::value;
          const uint32_t sc = sc_addr + col0 * 4;
          uint64_t v[W / 2];
#pragma unroll
          for (int j4 = 0; j4 < W / 4; ++j4) {
            if (kUnit) {
              v[j4 * 2 + 0] = ptx::pack2(d[j4 * 4 + 0], d[j4 * 4 + 1]);
              v[j4 * 2 + 1] = ptx::pack2(d[j4 * 4 + 2], d[j4 * 4 + 3]);
            } else {
              const float4 cm = ptx::lds128(sc + j4 * 16);
              v[j4 * 2 + 0] = ptx::fmul2(ptx::pack2(d[j4 * 4 + 0], d[j4 * 4 + 1]), ptx::pack2f(cm.x, cm.y));
              v[j4 * 2 + 1] = ptx::fmul2(ptx::pack2(d[j4 * 4 + 2], d[j4 * 4 + 3]), ptx::pack2f(cm.z, cm.w));
            }
          }
          if (kGuard) {  // TMA zero-fills columns >= M and the stale scales behind them are meaningless
#pragma unroll
            for (int j = 0; j < W / 2; ++j) {
              float lo, hi;
              ptx::unpack2f(v[j], lo, hi);
              if (col0 + 2 * j >= ncols) lo = -INFINITY;
              if (col0 + 2 * j + 1 >= ncols) hi = -INFINITY;
              v[j] = ptx::pack2f(lo, hi);
            }
          }
#pragma unroll
          for (int h = 0; h < W / GRP; ++h) {
            float f[GRP];
#pragma unroll
            for (int j = 0; j < GRP / 2; ++j) ptx::unpack2f(v[h * 4 + j], f[2 * j], f[2 * j + 1]);
            const float a0 = ptx::fmax3(f[0], f[1], f[2]), a1 = ptx::fmax3(f[3], f[4], f[5]);
            const float gm = ptx::fmax3(a0, a1, fmaxf(f[6], f[7]));
            // strict: an equal value in a later group never displaces the first maximal index
            const bool up = gm > vmax[r];
            ptx::stg_pred32(up, stash + r * 2 * 8192, v[h * 4 + 0], v[h * 4 + 1], v[h * 4 + 2], v[h * 4 + 3]);
            vgrp[r] = up ? col_base + col0 + h * GRP : vgrp[r];
            vmax[r] = up ? gm : vmax[r];
          }
        };
        using guard_off = std::integral_constant<bool, false>;
        using guard_on = std::integral_constant<bool, true>;

#ifdef GADM_DBG_NOEPI
        if (false)
#endif
        if (ncols > 0) {
          uint32_t ra[32], rb[16];
          ptx::tmem_ld_32x32(s_tmem, ra);
          ptx::tmem_ld_32x16(s_tmem + 32, rb);
          ptx::tmem_ld_wait();
          if (ncols >= 32) process(ra, 0, guard_off{});
          else process(ra, 0, guard_on{});         // ragged last tile
          if (ncols > 32) {
            if (ncols >= CS) process(rb, 32, guard_off{});
            else process(rb, 32, guard_on{});
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          ptx::mbar_arrive(&bars->s_free[r]);
          if (!kUnit && r == RT - 1) ptx::mbar_arrive(&bars->aux_empty[slot]);
        }
      }
      // A row has 4 tracks (one per column slice, in 4 different warps).  Alone, each raises its running maximum
      // (two stash stores, wavefronts of the data pipe the MMAs saturate) ~H(n) times; after tiles 0, 1, 3, 7, 15
      // the slices publish their maxima and adopt the row's: a track that adopts a larger maximum than its own
      // gives up its record (its index becomes a sentinel that loses every tie -- the holder sits at an earlier
      // column), and from then on only values above the ROW's maximum so far are recorded.
      if ((t & (t + 1)) == 0 && t < 16 && t + 1 < num_tiles) {
#pragma unroll
        for (int r = 0; r < RT; ++r) smem_xmax[(r * SL + sub) * BM + row_in_tile] = vmax[r];
        asm volatile("bar.sync 2, %0;" ::"n"(EPI_WARPS * 32) : "memory");
#pragma unroll
        for (int r = 0; r < RT; ++r) {
          float m = vmax[r];
#pragma unroll
          for (int s2 = 0; s2 < SL; ++s2) m = fmaxf(m, smem_xmax[(r * SL + s2) * BM + row_in_tile]);
          if (vmax[r] < m) { vmax[r] = m; vgrp[r] = FRAG_NO_RECORD; }
        }
        asm volatile("bar.sync 2, %0;" ::"n"(EPI_WARPS * 32) : "memory");
      }
    }

    // ---- per row tile: first maximal index of this slice from the stash (own stores, read back through L2), then
    // the merge of the 4 column slices through shared memory
    int vidx[RT];
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      vidx[r] = FRAG_NO_RECORD;
      if (vmax[r] > -INFINITY && vgrp[r] != FRAG_NO_RECORD) {
        int j_first = GRP - 1;
#pragma unroll
        for (int k = GRP / 4 - 1; k >= 0; --k) {
          const float4 sv = ptx::ldg_cg128(stash + (r * 2 + k) * 8192);
          if (sv.w == vmax[r]) j_first = 4 * k + 3;
          if (sv.z == vmax[r]) j_first = 4 * k + 2;
          if (sv.y == vmax[r]) j_first = 4 * k + 1;
          if (sv.x == vmax[r]) j_first = 4 * k + 0;
        }
        vidx[r] = vgrp[r] + j_first;
      }
      if (sub > 0) {
        float* x = smem_xch + ((r * (SL - 1) + sub - 1) * BM + row_in_tile) * 2;
        x[0] = vmax[r]; x[1] = __int_as_float(vidx[r]);
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
    if (sub == 0) {
#pragma unroll
      for (int r = 0; r < RT; ++r) {
        const int row = row0 + r * BM + row_in_tile;
        if (row >= p.N) continue;
        const size_t grow = size_t(b) * p.N + row;
        float vm = vmax[r];
        int vi = vidx[r];
#pragma unroll
        for (int s2 = 0; s2 < SL - 1; ++s2) {
          const float* x = smem_xch + ((r * (SL - 1) + s2) * BM + row_in_tile) * 2;
          const float v1 = x[0];
          const int i1 = __float_as_int(x[1]);
          if (v1 > vm || (v1 == vm && i1 < vi)) { vm = v1; vi = i1; }
        }
        // kUnit searched with unit column norms; the winner's similarity is reported with its true scale
        if (kUnit) vm *= p.scales[size_t(obj) * p.M + vi];
        const bool keep = p.mask == nullptr || p.mask[grow] != 0;
        float best = vm * p.rinv_rows[grow];
        int64_t best_idx = vi;
        if (p.pad_mode != GADM_PAD_NONE) {
          const float ps = p.pad_sim[grow];
          if (ps > best) { best = ps; best_idx = p.M; }  // pad column is the last one: wins only if strictly larger
        }
        p.idx[grow] = keep ? best_idx : int64_t(-1);
        p.max_sim[grow] = keep ? best : 0.f;
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == EPI_WARPS + 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}


inline size_t match_ta_smem_bytes(int stages) {
  return size_t(stages) * TB_STAGE_BYTES + AUX_SLOTS * PLANE_BYTES + 2 * 4 * BM * 4 + 2 * 3 * BM * 8 + sizeof(Barriers) + 1024;
}


// ---------------------------------------------------------------------------------------------------------------
// Flash-style CircleLoss forward: the training-side twin of the matcher.
// Reference: models/geoMatch.py:102-157 (per sample: foreground rows, normalise, sim = F^ M^_pad with the -1 pad
// column), :55-83 (positive mask: model vertices that are visible AND within positive_r of the row's ground-truth
// vertex; rows whose match_idx == M have the pad column as their only positive), models/loss.py:475-490
// (ap = clamp(1 + m - s, 0), an = clamp(s + m, 0), logit_p = -ap (s - (1 - m)) gamma, logit_n = an (s - m) gamma,
// loss_row = softplus(LSE_p + LSE_n)).  The reference materialises sim [n_fg, M + 1] and runs ~12 elementwise
// passes over it; here the two masked sums are accumulated in the epilogue of the similarity GEMM, the positive
// mask is evaluated on the fly from the model coordinates (invisible vertices are moved to 1e18 in the per-frame
// planes the caller passes), and only 12 bytes per row leave the SM.
// The sums need no running maximum: |logit| <= gamma (2 + m)(2 - m) and gadm_circle_loss_fwd admits only
// gamma (2 + m)(2 - m) log2(e) <= 120, so 2^logit stays inside the fp32 range.
// Tiling: match_kernel<soft, 1> (one row tile per CTA, the two accumulators alternate between model tiles,
// thread = row x 64-column slice).
struct CircleParams {
  const float* rinv_rows;   // [B, N]
  const float* pad_sim;     // [B, N] similarity with the -1 pad column
  const float* scales;      // [n_obj, M]
  const float* planes;      // [4, B, M] per-FRAME x / y / z planes (invisible vertices at 1e18) + squared positive radius
  const float* xyz;         // [n_obj, M, 3] model coordinates (ground-truth vertex lookup)
  const int64_t* match_idx; // [B, N], M = not on the model
  const uint8_t* fg;        // [B, N] rows that take part (labels == 1)
  const int32_t* obj_id;
  float* loss;              // [B, N] softplus(LSE_p + LSE_n), 0 for rows outside fg
  float* lse_p;             // [B, N] natural-log LSE of the positive / negative logits (for a backward pass)
  float* lse_n;
  const float* w;           // kGrad: [B, N] dL/dz of every row (0 for rows that take no part)
  float* G;                 // kGrad: [B, N, Mp] dL/dsim, column M = pad column, columns M+1.. = 0
  int Mp;
  int B, N, M, KB, n_obj, stages;
  float gamma_log2e, margin;
};

// kGrad: the same pass, but instead of the two sums every score's gradient is written,
//   dL/dsim_ij = w_i * (j positive ? softmax_p(j) * (-ap_ij gamma) : softmax_n(j) * (an_ij gamma)),
// with ap / an constants (the reference detaches them, loss.py:479-480) and the row's two LSEs from the forward pass.
template <bool kGrad>
__global__ void __launch_bounds__(NUM_THREADS, 1)
circle_kernel(const __grid_constant__ CUtensorMap tmap_rows, const __grid_constant__ CUtensorMap tmap_cols,
              const CircleParams p) {
  constexpr int AUX_BYTES = 5 * PLANE_BYTES;
  constexpr int SL = 4, CS = BN / SL;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                                    // [KB] blocks of 128 rows x 64 k
  uint8_t* smem_b = smem_a + p.KB * A_BLK_BYTES;
  uint8_t* smem_aux = smem_b + p.stages * B_STAGE_BYTES;   // per slot: [1/|m| x256 | x | y | z | r^2]
  Barriers* bars = reinterpret_cast<Barriers*>(smem_aux + AUX_SLOTS * AUX_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int row0 = blockIdx.x * BM;
  const int obj = p.obj_id ? p.obj_id[b] : (p.n_obj == p.B ? b : 0);
  const int num_tiles = (p.M + BN - 1) / BN;

  if (warp == EPI_WARPS && lane == 0) {
    ptx::prefetch_tensormap(&tmap_rows);
    ptx::prefetch_tensormap(&tmap_cols);
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(&bars->full[s], 1);
      ptx::mbar_init(&bars->empty[s], 1);
    }
    ptx::mbar_init(&bars->a_full, 1);
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&bars->s_full[a], 1);
      ptx::mbar_init(&bars->s_free[a], EPI_WARPS);
    }
    for (int a = 0; a < AUX_SLOTS; ++a) {
      ptx::mbar_init(&bars->aux_full[a], 1);
      ptx::mbar_init(&bars->aux_empty[a], EPI_WARPS);
    }
    ptx::fence_mbar_init();
  }
  if (warp == EPI_WARPS + 1) {
    ptx::tmem_alloc(&bars->tmem_base, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == EPI_WARPS) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(&bars->a_full, p.KB * A_BLK_BYTES);
      for (int kb = 0; kb < p.KB; ++kb)
        ptx::tma_load_3d(smem_a + kb * A_BLK_BYTES, &tmap_rows, &bars->a_full, kb * BK, row0, b);
      int stage = 0;
      uint32_t phase = 0;
      const size_t plane = size_t(p.B) * p.M;
      const float* sc_tab = p.scales + size_t(obj) * p.M;
      const float* xyz_tab = p.planes + size_t(b) * p.M;        // per-frame planes
      for (int t = 0; t < num_tiles; ++t) {
        const int slot = t % AUX_SLOTS;
        const uint32_t use = uint32_t(t) / AUX_SLOTS;
        const uint32_t bytes = uint32_t(min(BN, p.M - t * BN)) * 4;
        ptx::mbar_wait_sleep(&bars->aux_empty[slot], (use & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&bars->aux_full[slot], 5 * bytes);
        uint8_t* aux = smem_aux + slot * AUX_BYTES;
        ptx::bulk_load_1d(aux, sc_tab + size_t(t) * BN, bytes, &bars->aux_full[slot]);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          ptx::bulk_load_1d(aux + (c + 1) * PLANE_BYTES, xyz_tab + c * plane + size_t(t) * BN, bytes,
                            &bars->aux_full[slot]);
        for (int kb = 0; kb < p.KB; ++kb) {
          ptx::mbar_wait_sleep(&bars->empty[stage], phase ^ 1);
          ptx::mbar_arrive_expect_tx(&bars->full[stage], B_STAGE_BYTES);
          ptx::tma_load_3d(smem_b + stage * B_STAGE_BYTES, &tmap_cols, &bars->full[stage], kb * BK, t * BN, obj);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == EPI_WARPS + 1) {
    // ============================== UMMA issuer ==============================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16_f32(BM, BN);
      ptx::mbar_wait(&bars->a_full, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < num_tiles; ++t) {
        const int acc = t & 1;
        ptx::mbar_wait_sleep(&bars->s_free[acc], ((uint32_t(t) >> 1) & 1) ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < p.KB; ++kb) {
          ptx::mbar_wait_sleep(&bars->full[stage], phase);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(smem_a + kb * A_BLK_BYTES);
          const uint32_t b_addr = ptx::smem_u32(smem_b + stage * B_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            ptx::umma_bf16_ss(d_tmem, ptx::umma_desc_sw128_kmajor(a_addr + k * UMMA_K * 2),
                              ptx::umma_desc_sw128_kmajor(b_addr + k * UMMA_K * 2), idesc, (kb | k) != 0);
          ptx::umma_commit(&bars->empty[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(&bars->s_full[acc]);
      }
    }
  } else {
    // ============================== epilogue warps (thread == row, 4 column slices per row) ==============
    const int q = warp & 3;
    const int sub = warp >> 2;
    const int row_in_tile = q * 32 + lane;
    const int row = row0 + row_in_tile;
    const bool row_ok = row < p.N;
    const size_t grow = size_t(b) * p.N + (row_ok ? row : 0);
    const float rs = row_ok ? p.rinv_rows[grow] : 0.f;
    const int64_t mi = row_ok ? p.match_idx[grow] : int64_t(p.M);
    const bool in_mesh = mi >= 0 && mi < p.M;
    // ground-truth vertex of the row; rows off the model sit at -1e18: no model vertex is ever within the radius
    float gx = -1e18f, gy = -1e18f, gz = -1e18f;
    if (in_mesh) {
      const float* e = p.xyz + (size_t(obj) * p.M + size_t(mi)) * 3;
      gx = e[0]; gy = e[1]; gz = e[2];
    }
    const float m = p.margin, one_m = 1.f - p.margin, one_p = 1.f + p.margin, gl = p.gamma_log2e;
    const uint32_t lane_base = tmem_base + (uint32_t(q * 32) << 16) + sub * CS;
    float sum_p = 0.f, sum_n = 0.f;
    // kGrad: row constants (log2 units) and the row of G
    const float Lp = kGrad && row_ok ? p.lse_p[grow] * 1.4426950408889634f : 0.f;
    const float Ln = kGrad && row_ok ? p.lse_n[grow] * 1.4426950408889634f : 0.f;
    const float wg = kGrad && row_ok ? p.w[grow] * (gl * 0.6931471805599453f) : 0.f;     // w_i * gamma
    float* grow_g = kGrad ? p.G + grow * size_t(p.Mp) : nullptr;

    for (int t = 0; t < num_tiles; ++t) {
      const int acc = t & 1;
      const uint32_t use = uint32_t(t) >> 1;
      const int slot = t % AUX_SLOTS;
      ptx::mbar_wait_sleep(&bars->aux_full[slot], (uint32_t(t) / AUX_SLOTS) & 1);
      ptx::mbar_wait_sleep(&bars->s_full[acc], use & 1);
      ptx::tc_fence_after();
      const int ncols = min(BN, p.M - t * BN) - sub * CS;   // valid columns of this slice (may be <= 0)
      const uint32_t s_tmem = lane_base + acc * BN;
      const uint32_t sc_addr = ptx::smem_u32(smem_aux + slot * AUX_BYTES) + sub * CS * 4;
#pragma unroll 1
      for (int c = 0; c < CS / 16; ++c) {
        if (ncols - c * 16 <= 0) break;
        uint32_t d[16];
        ptx::tmem_ld_32x16(s_tmem + c * 16, d);
        ptx::tmem_ld_wait();
        float gout[16];
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const uint32_t a = sc_addr + (c * 16 + j4 * 4) * 4;
          const float4 cm = ptx::lds128(a);
          const float4 X = ptx::lds128(a + PLANE_BYTES), Y = ptx::lds128(a + 2 * PLANE_BYTES),
                       Z = ptx::lds128(a + 3 * PLANE_BYTES), R = ptx::lds128(a + 4 * PLANE_BYTES);
          const float cs[4] = {cm.x, cm.y, cm.z, cm.w}, xs[4] = {X.x, X.y, X.z, X.w};
          const float ys[4] = {Y.x, Y.y, Y.z, Y.w}, zs[4] = {Z.x, Z.y, Z.z, Z.w}, r2s[4] = {R.x, R.y, R.z, R.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float s = (__uint_as_float(d[j4 * 4 + e]) * cs[e]) * rs;             // cosine similarity
            // (A - B).pow(2).sum(-1) as the reference evaluates it: no FMA, left to right (basic_utils.py:88)
            const float dx = __fsub_rn(gx, xs[e]), dy = __fsub_rn(gy, ys[e]), dz = __fsub_rn(gz, zs[e]);
            const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            const bool pos = __fadd_rn(d2, 1e-7f) < r2s[e];                            // sqrt(D2 + 1e-7) < positive_r[j]
            const float ap = fmaxf(one_p - s, 0.f), an = fmaxf(s + m, 0.f);             // loss.py:479-480
            const float lp = -ap * (s - one_m) * gl, ln = an * (s - m) * gl;            // loss.py:488-489 (log2 units)
            const bool valid = c * 16 + j4 * 4 + e < ncols;
            if (kGrad) {
              const float sm = ptx::ex2_approx(pos ? lp - Lp : ln - Ln);                // softmax weight inside its set
              gout[j4 * 4 + e] = wg * sm * (pos ? -ap : an);
            } else {
              const float ex = ptx::ex2_approx(pos ? lp : ln);
              sum_p += (valid && pos) ? ex : 0.f;
              sum_n += (valid && !pos) ? ex : 0.f;
            }
          }
        }
        if (kGrad && row_ok) {
          float* dst = grow_g + t * BN + sub * CS + c * 16;
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4)
            if (c * 16 + j4 * 4 < ncols)           // M % 8 == 0 and 16-byte groups: a float4 is valid or invalid as a whole
              *reinterpret_cast<float4*>(dst + j4 * 4) =
                  make_float4(gout[j4 * 4], gout[j4 * 4 + 1], gout[j4 * 4 + 2], gout[j4 * 4 + 3]);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(&bars->s_free[acc]);
        ptx::mbar_arrive(&bars->aux_empty[slot]);
      }
    }

    // ---- merge the 4 column slices (exchange buffer aliases the A blocks: all MMAs have completed), add the pad
    // column (positive exactly for the rows that are off the model: geoMatch.py:78), softplus
    if (kGrad) {
      if (sub == 0 && row_ok) {       // the pad column and the zero padding of the row
        const float s = p.pad_sim[grow];
        const float ap = fmaxf(one_p - s, 0.f), an = fmaxf(s + m, 0.f);
        const float gp = in_mesh ? wg * ptx::ex2_approx(an * (s - m) * gl - Ln) * an
                                 : wg * ptx::ex2_approx(-ap * (s - one_m) * gl - Lp) * -ap;
        for (int j = p.M; j < p.Mp; ++j) grow_g[j] = j == p.M ? gp : 0.f;
      }
    }
    float* xch = reinterpret_cast<float*>(smem_a);      // 3 * 128 * 8 B
    if (!kGrad && sub > 0) {
      xch[((sub - 1) * BM + row_in_tile) * 2 + 0] = sum_p;
      xch[((sub - 1) * BM + row_in_tile) * 2 + 1] = sum_n;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
    if (!kGrad && sub == 0 && row_ok) {
#pragma unroll
      for (int s2 = 0; s2 < SL - 1; ++s2) {
        sum_p += xch[(s2 * BM + row_in_tile) * 2 + 0];
        sum_n += xch[(s2 * BM + row_in_tile) * 2 + 1];
      }
      const float s = p.pad_sim[grow];
      const float ap = fmaxf(one_p - s, 0.f), an = fmaxf(s + m, 0.f);
      if (in_mesh) sum_n += ptx::ex2_approx(an * (s - m) * gl);
      else sum_p += ptx::ex2_approx(-ap * (s - one_m) * gl);
      const bool keep = p.fg == nullptr || p.fg[grow] != 0;
      const float lse_p = logf(sum_p), lse_n = logf(sum_n);      // log(0) = -inf: a row without positives costs 0
      const float z = lse_p + lse_n;
      const float sp = z > 20.f ? z : log1pf(expf(z));            // nn.Softplus(beta = 1, threshold = 20)
      p.loss[grow] = keep ? sp : 0.f;
      p.lse_p[grow] = keep ? lse_p : 0.f;
      p.lse_n[grow] = keep ? lse_n : 0.f;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == EPI_WARPS + 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

inline size_t circle_smem_bytes(int KB, int stages) {
  return size_t(KB) * A_BLK_BYTES + size_t(stages) * B_STAGE_BYTES + AUX_SLOTS * 5 * PLANE_BYTES + sizeof(Barriers) + 1024;
}

int g_stash_slots = 0;     // %nsmid of the device gadm_init() ran on (written once, read-only afterwards)

__global__ void nsmid_kernel(unsigned int* out) {
  unsigned int v;
  asm volatile("mov.u32 %0, %%nsmid;" : "=r"(v));
  *out = v;
}

}  // namespace

int match_configure() {
  cudaError_t e;
  e = cudaFuncSetAttribute(match_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return set_cuda_error(e);
  e = cudaFuncSetAttribute(match_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return set_cuda_error(e);
  e = cudaFuncSetAttribute(match_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return set_cuda_error(e);
  e = cudaFuncSetAttribute(match_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return set_cuda_error(e);
  e = cudaFuncSetAttribute(match_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return set_cuda_error(e);
  e = cudaFuncSetAttribute(match_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return set_cuda_error(e);
  e = cudaFuncSetAttribute(circle_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return set_cuda_error(e);
  e = cudaFuncSetAttribute(circle_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return set_cuda_error(e);
  e = cudaFuncSetAttribute(match_ta_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return set_cuda_error(e);
  e = cudaFuncSetAttribute(match_ta_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return set_cuda_error(e);
  e = cudaFuncSetAttribute(match_alt_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return set_cuda_error(e);
  e = cudaFuncSetAttribute(match_alt_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return set_cuda_error(e);
  e = cudaFuncSetAttribute(match_alt_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return set_cuda_error(e);
  e = cudaFuncSetAttribute(match_frag_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return set_cuda_error(e);
  e = cudaFuncSetAttribute(match_frag_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return set_cuda_error(e);
  e = cudaFuncSetAttribute(match_frag_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return set_cuda_error(e);
  e = cudaFuncSetAttribute(match_frag_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return set_cuda_error(e);
  int dev = 0, sms = 0;
  e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return set_cuda_error(e);
  e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return set_cuda_error(e);
  // the stash slots are indexed by %smid, whose range is [0, %nsmid) -- read it once (init may synchronise)
  unsigned int* d_n = nullptr;
  unsigned int nsmid = 0;
  e = cudaMalloc(&d_n, sizeof(unsigned int));
  if (e != cudaSuccess) return set_cuda_error(e);
  nsmid_kernel<<<1, 1>>>(d_n);
  e = cudaMemcpy(&nsmid, d_n, sizeof(unsigned int), cudaMemcpyDeviceToHost);
  cudaFree(d_n);
  if (e != cudaSuccess) return set_cuda_error(e);
  g_stash_slots = sms > int(nsmid) ? sms : int(nsmid);
  return GADM_OK;
}

// One stash slot per SM (the fragment-layout kernel runs one CTA per SM and indexes its slot by %smid).
size_t match_workspace_bytes() { return size_t(g_stash_slots) * FRAG_STASH_BYTES; }

template <bool kSoft>
static int match_launch_t(const void* rows, const void* cols, MatchParams p, int Kp, cudaStream_t stream) {
  const int KB = Kp / BK;
  {
    // fragment-layout kernel (needs the stash workspace).  RT = 1: one row tile per CTA and the two accumulators
    // alternate between consecutive model tiles; RT = 2: two row tiles per CTA (half the L2 operand traffic).
    // Measured at the BASELINE shape: ARGMAX 0.242 / 0.223 ms (RT = 1 / 2) against 0.211 ms for
    // match_kernel<., 2>, SOFT 0.445 / 0.426 ms against 0.383 ms for the paired-row kernel -- it quarters the
    // epilogue's shared-memory wavefronts but pays for them in issue slots (16 stash stores per chunk) and the
    // SOFT epilogue is MUFU / FMA-bound either way, so it is NOT the default.
    // GADM_MATCH_FRAG=1/2 selects it with RT = 1 / RT = 2 (profiling, tests).
    int frt = 0;
    if (const char* f = getenv("GADM_MATCH_FRAG")) frt = atoi(f);
    if (frt == 2 && (match_frag_stages<kSoft>(2, KB) < 2 * KB || p.N <= BM)) frt = 1;
    const int fstages = frt > 0 ? match_frag_stages<kSoft>(frt, KB) : 0;
    if (frt > 0 && p.stash != nullptr && fstages >= 2) {
      p.KB = KB; p.stages = fstages;
      CUtensorMap tmap_rows, tmap_cols;
      int rc = make_tmap_2b_3d(&tmap_rows, rows, uint64_t(Kp), uint64_t(p.N), uint64_t(p.B), BK, BM, 0);
      if (rc != GADM_OK) return rc;
      rc = make_tmap_2b_3d(&tmap_cols, cols, uint64_t(Kp), uint64_t(p.M), uint64_t(p.n_obj), BK, BN, 0);
      if (rc != GADM_OK) return rc;
      dim3 grid((p.N + frt * BM - 1) / (frt * BM), p.B);
      const size_t smem = match_frag_smem_bytes<kSoft>(frt, KB, fstages);
      if (frt == 2)
        match_frag_kernel<kSoft, 2><<<grid, FRAG_THREADS, smem, stream>>>(tmap_rows, tmap_cols, p);
      else
        match_frag_kernel<kSoft, 1><<<grid, FRAG_THREADS, smem, stream>>>(tmap_rows, tmap_cols, p);
      return check_launch();
    }
  }
  if (!kSoft) {
    // TMEM-resident A (ARGMAX, K' <= 128, needs the stash workspace): opt-in with GADM_MATCH_TA=1 (profiling, tests).
    // Parity-green but slower: A from TMEM leaves room for 192-column accumulators only, and a 128x192x16 MMA takes
    // the time of a 128x256x16 one -- with the epilogue compiled out the pipeline reaches 77 % of the bf16 peak
    // (99.7 % for match_alt_kernel's 128x256x16 MMAs from shared memory); ARGMAX 0.206 ms against 0.187-0.197 ms,
    // ARGMAX_UNIT 0.190 ms against 0.156-0.164 ms.
    bool ta = false;
    if (const char* f = getenv("GADM_MATCH_TA")) ta = atoi(f) != 0;
    if (ta && p.stash != nullptr && KB <= 2 && p.N > BM) {
      p.KB = KB; p.stages = MAX_STAGES;
      p.rows_ptr = rows;
      CUtensorMap tmap_rows, tmap_cols;
      int rc = make_tmap_2b_3d(&tmap_rows, rows, uint64_t(Kp), uint64_t(p.N), uint64_t(p.B), BK, BM, 0);
      if (rc != GADM_OK) return rc;
      rc = make_tmap_2b_3d(&tmap_cols, cols, uint64_t(Kp), uint64_t(p.M), uint64_t(p.n_obj), BK, TBN, 0);
      if (rc != GADM_OK) return rc;
      dim3 grid((p.N + 2 * BM - 1) / (2 * BM), p.B);
      if (p.unit_scales == 1)
        match_ta_kernel<true><<<grid, NUM_THREADS, match_ta_smem_bytes(p.stages), stream>>>(tmap_rows, tmap_cols, p);
      else
        match_ta_kernel<false><<<grid, NUM_THREADS, match_ta_smem_bytes(p.stages), stream>>>(tmap_rows, tmap_cols, p);
      return check_launch();
    }
  }
  if (!kSoft) {
    // alternating kernel (ARGMAX, needs the stash workspace, a ring of two resident model tiles and more than one
    // row tile per frame).  GADM_MATCH_ALT=0 forbids it (profiling, tests).
    bool alt = true;
    if (const char* f = getenv("GADM_MATCH_ALT")) alt = atoi(f) != 0;
    const int astages = match_alt_stages(KB);
    if (alt && p.stash != nullptr && astages >= 2 * KB && p.N > BM) {
      p.KB = KB; p.stages = astages;
      CUtensorMap tmap_rows, tmap_cols;
      int rc = make_tmap_2b_3d(&tmap_rows, rows, uint64_t(Kp), uint64_t(p.N), uint64_t(p.B), BK, BM, 0);
      if (rc != GADM_OK) return rc;
      rc = make_tmap_2b_3d(&tmap_cols, cols, uint64_t(Kp), uint64_t(p.M), uint64_t(p.n_obj), BK, BN, 0);
      if (rc != GADM_OK) return rc;
      dim3 grid((p.N + 2 * BM - 1) / (2 * BM), p.B);
      if (p.unit_scales == 1)
        match_alt_kernel<true, false><<<grid, NUM_THREADS, match_alt_smem_bytes(KB, astages), stream>>>(tmap_rows, tmap_cols, p);
      else if (p.unit_scales == 2)
        match_alt_kernel<false, true><<<grid, NUM_THREADS, match_alt_smem_bytes(KB, astages), stream>>>(tmap_rows, tmap_cols, p);
      else
        match_alt_kernel<false, false><<<grid, NUM_THREADS, match_alt_smem_bytes(KB, astages), stream>>>(tmap_rows, tmap_cols, p);
      return check_launch();
    }
  }
  {
    // paired-row kernel (256 rows per CTA, every epilogue thread owns two rows).  Measured at the BASELINE shape:
    // SOFT 0.384 ms against 0.400 ms (RT = 1), ARGMAX 0.241 ms against 0.211 ms (RT = 2) => default for SOFT only.
    // GADM_MATCH_PAIR=1/0 forces / forbids it (profiling, tests).
    bool pair = kSoft;
    if (const char* f = getenv("GADM_MATCH_PAIR")) pair = atoi(f) != 0;
    const int pstages = match_pair_stages<kSoft>(KB);
    if (pair && pstages >= 2 * KB && p.N > BM) {
      p.KB = KB; p.stages = pstages;
      CUtensorMap tmap_rows, tmap_cols;
      int rc = make_tmap_2b_3d(&tmap_rows, rows, uint64_t(Kp), uint64_t(p.N), uint64_t(p.B), BK, BM, 0);
      if (rc != GADM_OK) return rc;
      rc = make_tmap_2b_3d(&tmap_cols, cols, uint64_t(Kp), uint64_t(p.M), uint64_t(p.n_obj), BK, PBN, 0);
      if (rc != GADM_OK) return rc;
      dim3 grid((p.N + 2 * BM - 1) / (2 * BM), p.B);
      match_pair_kernel<kSoft><<<grid, NUM_THREADS, match_pair_smem_bytes<kSoft>(KB, pstages), stream>>>(
          tmap_rows, tmap_cols, p);
      return check_launch();
    }
  }
  // Two row tiles per CTA when a ring of at least 2 KB stages (one tile resident, one in flight) fits beside them
  // and the frame has more than one row tile; otherwise one row tile with the deepest ring.  Measured at the
  // BASELINE shape: ARGMAX 0.204 ms (RT = 2) against 0.222 ms; SOFT 0.43 ms (RT = 2) against 0.41 ms -- SOFT is
  // bound by the epilogue's shared-memory traffic, not by operand traffic, and prefers 16 warps per accumulator.
  int RT = kSoft ? 1 : 2;
  if (const char* f = getenv("GADM_MATCH_RT")) RT = atoi(f) == 1 ? 1 : 2;   // profiling aid
  int stages = match_stages<kSoft>(2, KB);
  if (RT == 2 && (stages < 2 * KB || p.N <= BM)) RT = 1;
  if (RT == 1) {
    stages = match_stages<kSoft>(1, KB);
    if (stages < 2) return GADM_ERR_UNSUPPORTED;
  }
  p.KB = KB; p.stages = stages;

  CUtensorMap tmap_rows, tmap_cols;
  int rc = make_tmap_2b_3d(&tmap_rows, rows, uint64_t(Kp), uint64_t(p.N), uint64_t(p.B), BK, BM, 0);
  if (rc != GADM_OK) return rc;
  rc = make_tmap_2b_3d(&tmap_cols, cols, uint64_t(Kp), uint64_t(p.M), uint64_t(p.n_obj), BK, BN, 0);
  if (rc != GADM_OK) return rc;

  dim3 grid((p.N + BM * RT - 1) / (BM * RT), p.B);
  const size_t smem = match_smem_bytes<kSoft>(RT, KB, stages);
  if (RT == 2)
    match_kernel<kSoft, 2><<<grid, NUM_THREADS, smem, stream>>>(tmap_rows, tmap_cols, p);
  else
    match_kernel<kSoft, 1><<<grid, NUM_THREADS, smem, stream>>>(tmap_rows, tmap_cols, p);
  return check_launch();
}

int match_launch(const void* rows, const float* rinv_rows, const float* pad_sim, const void* cols, const float* aux,
                 const uint8_t* mask, const int32_t* obj_id, int B, int N, int M, int Kp, int n_obj, float gamma,
                 int pad_mode, int mode, int64_t* idx, float* max_sim, float* weight, float* soft_xyz,
                 void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  MatchParams p;
  const bool ws_ok = workspace != nullptr && g_stash_slots > 0 && workspace_bytes >= match_workspace_bytes();
  p.stash = ws_ok ? static_cast<uint8_t*>(workspace) : nullptr;
  p.stash_slots = g_stash_slots;
  p.rinv_rows = rinv_rows; p.pad_sim = pad_sim; p.scales = aux_scales(aux, n_obj, M);
  p.planes = aux_planes(aux, n_obj, M); p.mask = mask; p.obj_id = obj_id;
  p.idx = idx; p.max_sim = max_sim; p.weight = weight; p.soft_xyz = soft_xyz;
  p.B = B; p.N = N; p.M = M; p.KB = 0; p.n_obj = n_obj; p.stages = 0; p.pad_mode = pad_mode;
  p.gamma_log2e = gamma * 1.4426950408889634f;
  p.unit_scales = mode == GADM_MATCH_ARGMAX_UNIT ? 1 : mode == GADM_MATCH_ARGMAX_BF16N ? 2 : 0;
  p.rows_ptr = rows;
  if (mode == GADM_MATCH_SOFT) return match_launch_t<true>(rows, cols, p, Kp, stream);
  return match_launch_t<false>(rows, cols, p, Kp, stream);
}

int circle_launch(const void* rows, const float* rinv_rows, const float* pad_sim, const void* cols, const float* aux,
                  const float* planes_frame, const int64_t* match_idx, const uint8_t* fg, const int32_t* obj_id, int B,
                  int N, int M, int Kp, int n_obj, float gamma, float margin, float* loss, float* lse_p,
                  float* lse_n, const float* w, float* G, int Mp, cudaStream_t stream) {
  CircleParams p;
  p.w = w; p.G = G; p.Mp = Mp;
  p.rinv_rows = rinv_rows; p.pad_sim = pad_sim; p.scales = aux_scales(aux, n_obj, M); p.planes = planes_frame;
  p.xyz = aux_xyz(aux, n_obj, M); p.match_idx = match_idx; p.fg = fg; p.obj_id = obj_id;
  p.loss = loss; p.lse_p = lse_p; p.lse_n = lse_n;
  p.B = B; p.N = N; p.M = M; p.n_obj = n_obj;
  p.gamma_log2e = gamma * 1.4426950408889634f; p.margin = margin;
  const int KB = Kp / BK;
  int stages = MAX_STAGES;
  while (stages > 0 && circle_smem_bytes(KB, stages) > 227 * 1024) --stages;
  if (stages < 2) return GADM_ERR_UNSUPPORTED;
  p.KB = KB; p.stages = stages;
  CUtensorMap tmap_rows, tmap_cols;
  int rc = make_tmap_2b_3d(&tmap_rows, rows, uint64_t(Kp), uint64_t(N), uint64_t(B), BK, BM, 0);
  if (rc != GADM_OK) return rc;
  rc = make_tmap_2b_3d(&tmap_cols, cols, uint64_t(Kp), uint64_t(M), uint64_t(n_obj), BK, BN, 0);
  if (rc != GADM_OK) return rc;
  dim3 grid((N + BM - 1) / BM, B);
  if (G != nullptr)
    circle_kernel<true><<<grid, NUM_THREADS, circle_smem_bytes(KB, stages), stream>>>(tmap_rows, tmap_cols, p);
  else
    circle_kernel<false><<<grid, NUM_THREADS, circle_smem_bytes(KB, stages), stream>>>(tmap_rows, tmap_cols, p);
  return check_launch();
}

}  // namespace gadm
