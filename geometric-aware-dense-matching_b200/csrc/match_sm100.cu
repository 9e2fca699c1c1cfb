// Fused matching head for sm_100a: descriptor similarity (tcgen05 UMMA, bf16 in / fp32 accumulate in TMEM,
// operands staged by TMA) + row-wise argmax / online softmax / soft model coordinates.
// The [N, M] score matrix never leaves TMEM/registers.
//
// Replaces (reference tree): evaluator.py:89-93 (normalize, normalize, matmul, torch.max) and the padded
// variants utils/pvn3d_eval_utils_kpls.py:436-444 / models/geoMatch_DGCNN.py:92-99.  The softmax weight
// and soft coordinates are the extension defined in oracle/match_oracle.py.
//
// Work decomposition
//   CTA  = one 128-row tile of one frame (grid = ceil(N/128) x B), 18 warps:
//     warp 16     TMA producer: the 128 x K' row tile once, then model tiles (256 vertices x 64 k, 32 KB) through an
//                 S-stage mbarrier ring, plus per tile the column scales 1/|m_j| and (SOFT) the x / y / z planes
//     warp 17     UMMA issuer: 128x256x16 tcgen05.mma, accumulators double-buffered in TMEM (2 x 256 columns)
//     warps 0..15 epilogue, thread = row: warp w owns TMEM lanes 32*(w%4).. and the 64-column slice w/4 of EVERY
//                 tile, so a tile is drained in the time one warp needs for two 32-column chunks -- the accumulator
//                 must be free again within one tile time of the tensor pipe, latency matters more than throughput.
//                 Per chunk: tcgen05.ld, score = acc * 1/|m_j| (packed f32x2), 3-input max tree, first-maximal-
//                 index search only when the chunk beats the running maximum; SOFT adds p = 2^(score*g - m_ref)
//                 against a LAZY reference exponent (raised, with a rescale of the sums, only when exceeded by
//                 more than 8), and fp32 sums of p and p * xyz.  The four slices of a row merge through shared
//                 memory at the end.
//   (A tensor-core P.V product for the coordinate sums was built and measured this round: a 128x16x16
//   tcgen05.mma costs ~120 cycles whatever its N, so 16 of them per tile cost more than the similarity GEMM.)
#include <cuda_fp16.h>
#include <float.h>
#include <stdio.h>
#include <stdlib.h>

#include <type_traits>

#include "gadm_internal.h"
#include "ptx.cuh"

namespace gadm {

namespace {

constexpr int BM = 128;               // rows (scene points) per CTA == UMMA M
constexpr int BK = 64;                // bf16 elements per 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int A_BLK_BYTES = BM * BK * 2;     // 16 KB
constexpr int MAX_STAGES = 6;
constexpr int BN = 256;               // model vertices per accumulator tile == UMMA N
constexpr int B_STAGE_BYTES = BN * BK * 2;   // 32 KB
constexpr int AUX_SLOTS = 4;                 // per-tile {1/|m|, x, y, z} ring, decoupled from the two accumulators
constexpr int PLANE_BYTES = BN * 4;          // one fp32 plane of a tile
constexpr int EPI_SUB = 4;                   // column slices per tile (epilogue warps per TMEM lane quarter)
constexpr int EPI_WARPS = 4 * EPI_SUB;
constexpr int CS = BN / EPI_SUB;             // columns per slice (64)
constexpr int NUM_THREADS = (EPI_WARPS + 2) * 32;   // warps 0-15 epilogue, 16 TMA, 17 UMMA
constexpr int XCH_BYTES = (EPI_SUB - 1) * BM * 8 * 4;  // per-row state exchange between the column slices
constexpr int BOUND_BYTES = BM * EPI_SUB * 4;          // running maxima of the four slices of every row
constexpr int TMEM_COLS = 512;
constexpr float LAZY_TAU = 8.f;              // reference exponent is raised only when exceeded by more than this

struct Barriers {
  uint64_t full[MAX_STAGES];
  uint64_t empty[MAX_STAGES];
  uint64_t a_full;
  uint64_t s_full[2];    // S accumulator of parity p complete (UMMA commit)
  uint64_t s_free[2];    // S accumulator of parity p drained by all epilogue warps
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
};

__device__ __forceinline__ int frame_object(const MatchParams& p, int b) {
  if (p.obj_id) return p.obj_id[b];
  return p.n_obj == p.B ? b : 0;
}

// first j with v[j] == m (m is the maximum of v, so one exists); W = 16 or 32
template <int W>
__device__ __forceinline__ int first_equal(const uint32_t (&v)[W], float m) {
  int j_a = W, j_b = W, j_c = W, j_d = W - 1;  // four independent select chains, W = no hit
  constexpr int Q = W / 4;
#pragma unroll
  for (int j = Q - 1; j >= 0; --j) {
    if (__uint_as_float(v[j]) == m) j_a = j;
    if (__uint_as_float(v[j + Q]) == m) j_b = j + Q;
    if (__uint_as_float(v[j + 2 * Q]) == m) j_c = j + 2 * Q;
    if (j < Q - 1 && __uint_as_float(v[j + 3 * Q]) == m) j_d = j + 3 * Q;
  }
  return min(min(j_a, j_b), min(j_c, j_d));
}

template <bool kSoft>
__global__ void __launch_bounds__(NUM_THREADS, 1)
match_kernel(const __grid_constant__ CUtensorMap tmap_rows, const __grid_constant__ CUtensorMap tmap_cols,
             const MatchParams p) {
  constexpr int AUX_BYTES = kSoft ? 4 * PLANE_BYTES : PLANE_BYTES;

  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B atoms need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem_a + p.KB * A_BLK_BYTES;
  uint8_t* smem_aux = smem_b + p.stages * B_STAGE_BYTES;   // per slot: [1/|m| x256 | x x256 | y x256 | z x256]
  uint8_t* smem_xch = smem_aux + AUX_SLOTS * AUX_BYTES;
  float* smem_bound = reinterpret_cast<float*>(smem_xch + XCH_BYTES);
  Barriers* bars = reinterpret_cast<Barriers*>(smem_xch + XCH_BYTES + BOUND_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int row0 = blockIdx.x * BM;
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
      ptx::mbar_init(&bars->s_free[a], EPI_WARPS);  // one arrive per epilogue warp
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
  if (threadIdx.x < BM * EPI_SUB) smem_bound[threadIdx.x] = -INFINITY;
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
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < num_tiles; ++t) {
        const int acc = t & 1;
        const uint32_t use = uint32_t(t) >> 1;
        ptx::mbar_wait(&bars->s_free[acc], (use & 1) ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < p.KB; ++kb) {
          ptx::mbar_wait(&bars->full[stage], phase);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(smem_a + kb * A_BLK_BYTES);
          const uint32_t b_addr = ptx::smem_u32(smem_b + stage * B_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            ptx::umma_bf16_ss(d_tmem, ptx::umma_desc_sw128_kmajor(a_addr + k * UMMA_K * 2),
                              ptx::umma_desc_sw128_kmajor(b_addr + k * UMMA_K * 2), idesc, (kb | k) != 0);
          }
          ptx::umma_commit(&bars->empty[stage]);  // frees the smem stage once these MMAs have read it
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(&bars->s_full[acc]);     // accumulator tile complete
      }
    }
  } else {
    // ============================== epilogue warps (thread == row, four column slices per row) ==============
    const int q = warp & 3;                 // TMEM lane quarter this warp may access (warp id % 4)
    const int sub = warp >> 2;              // 64-column slice of every tile
    const int row_in_tile = q * 32 + lane;
    const int row = row0 + row_in_tile;
    const bool row_ok = row < p.N;
    const size_t grow = size_t(b) * p.N + (row_ok ? row : 0);
    const float rs = row_ok ? p.rinv_rows[grow] : 0.f;
    const float g = p.gamma_log2e * rs;     // exponent scale: t = (acc * 1/|m_j|) * g   (log2 units)
    const uint32_t lane_base = tmem_base + (uint32_t(q * 32) << 16) + sub * CS;

    // The running maximum of each slice is published per tile; a chunk below the best value any slice of the row
    // has seen cannot hold the row's argmax, so the (warp-divergent) index search is skipped for it.  Stale
    // values are still valid lower bounds, no synchronisation is needed.
    const uint32_t bound_addr = ptx::smem_u32(smem_bound + row_in_tile * EPI_SUB);
    float vmax = -INFINITY, thr = -INFINITY;
    int vidx = 0;
    float mref = 0.f;
    bool have_ref = false;
    // packed (even | odd column) partial sums, two independent chains each
    uint64_t l2a = 0, l2b = 0, ax2a = 0, ax2b = 0, ay2a = 0, ay2b = 0, az2a = 0, az2b = 0;

#ifdef GADM_MATCH_TRACE
    long long tr[4][4];
#endif
    for (int t = 0; t < num_tiles; ++t) {
      const int acc = t & 1;
      const uint32_t use = uint32_t(t) >> 1;
      const int slot = t % AUX_SLOTS;
#ifdef GADM_MATCH_TRACE
      if (t >= 8 && t < 12) tr[t - 8][0] = clock64();
#endif
      // both barriers are polled back to back so that their check latencies overlap
      if (!(ptx::mbar_try_wait(&bars->aux_full[slot], (uint32_t(t) / AUX_SLOTS) & 1) &
            ptx::mbar_try_wait(&bars->s_full[acc], use & 1))) {
        ptx::mbar_wait_sleep(&bars->aux_full[slot], (uint32_t(t) / AUX_SLOTS) & 1);
        ptx::mbar_wait_sleep(&bars->s_full[acc], use & 1);
      }
      ptx::tc_fence_after();
#ifdef GADM_MATCH_TRACE
      if (t >= 8 && t < 12) tr[t - 8][1] = clock64();
#endif
      const int ncols = min(BN, p.M - t * BN) - sub * CS;   // valid columns of this slice (may be <= 0)
      const uint32_t s_tmem = lane_base + acc * BN;
      const uint32_t sc_addr = ptx::smem_u32(smem_aux + slot * AUX_BYTES) + sub * CS * 4;

      // One chunk of W (32 or 16) columns starting at slice column col0: scores, running (max, first argmax);
      // SOFT: exponentials, sums of p and p * xyz.  kGuard (ragged last tile only) masks columns >= ncols.
      auto process = [&](auto& r, int col0, auto guard_tag) {
        constexpr int W = int(sizeof(r) / sizeof(r[0]));
        constexpr bool kGuard = decltype(guard_tag)::value;
        const uint32_t sc = sc_addr + col0 * 4;
#pragma unroll
        for (int j4 = 0; j4 < W / 4; ++j4) {
          const float4 cm = ptx::lds128(sc + j4 * 16);
          const uint64_t v01 = ptx::fmul2(ptx::pack2(r[j4 * 4 + 0], r[j4 * 4 + 1]), ptx::pack2f(cm.x, cm.y));
          const uint64_t v23 = ptx::fmul2(ptx::pack2(r[j4 * 4 + 2], r[j4 * 4 + 3]), ptx::pack2f(cm.z, cm.w));
          ptx::unpack2(v01, r[j4 * 4 + 0], r[j4 * 4 + 1]);
          ptx::unpack2(v23, r[j4 * 4 + 2], r[j4 * 4 + 3]);
        }
        if (kGuard) {  // TMA zero-fills columns >= M and the stale scales behind them are meaningless
#pragma unroll
          for (int j = 0; j < W; ++j)
            if (col0 + j >= ncols) r[j] = __float_as_uint(-INFINITY);
        }
        // 3-input max tree
        float m4[W / 4];
#pragma unroll
        for (int u = 0; u < W / 4; ++u)
          m4[u] = ptx::fmax3(__uint_as_float(r[u * 4]), __uint_as_float(r[u * 4 + 1]),
                             fmaxf(__uint_as_float(r[u * 4 + 2]), __uint_as_float(r[u * 4 + 3])));
        float cmx = fmaxf(ptx::fmax3(m4[0], m4[1], m4[2]), m4[3]);
        if (W == 32) cmx = ptx::fmax3(cmx, ptx::fmax3(m4[W / 4 - 4], m4[W / 4 - 3], m4[W / 4 - 2]), m4[W / 4 - 1]);
        // strict against the own maximum (an equal value in a later chunk never displaces the first maximal
        // index), non-strict against the other slices (an equal value there may sit at a higher index)
        if (cmx > vmax && cmx >= thr) {
          vmax = cmx;
          vidx = t * BN + sub * CS + col0 + first_equal(r, cmx);
        }
        if (kSoft) {
          const float tnew = cmx * g;
          if (!have_ref) { mref = tnew; have_ref = true; }   // warp-uniform: first chunk of the first tile
          if (__any_sync(0xffffffffu, tnew > mref + LAZY_TAU)) {
            // rare: raise the reference exponent and rescale the running sums
            const bool need = tnew > mref + LAZY_TAU;
            const float s = need ? ptx::ex2_approx(mref - tnew) : 1.f;
            const uint64_t s2 = ptx::pack2f(s, s);
            l2a = ptx::fmul2(l2a, s2); l2b = ptx::fmul2(l2b, s2);
            ax2a = ptx::fmul2(ax2a, s2); ax2b = ptx::fmul2(ax2b, s2);
            ay2a = ptx::fmul2(ay2a, s2); ay2b = ptx::fmul2(ay2b, s2);
            az2a = ptx::fmul2(az2a, s2); az2b = ptx::fmul2(az2b, s2);
            mref = need ? tnew : mref;
          }
          // p = 2^(v*g - m_ref); sums in packed f32x2 (even | odd column)
          const uint64_t g2 = ptx::pack2f(g, g), nm2 = ptx::pack2f(-mref, -mref);
#pragma unroll
          for (int j4 = 0; j4 < W / 4; ++j4) {
            const float4 X = ptx::lds128(sc + PLANE_BYTES + j4 * 16);
            const float4 Y = ptx::lds128(sc + 2 * PLANE_BYTES + j4 * 16);
            const float4 Z = ptx::lds128(sc + 3 * PLANE_BYTES + j4 * 16);
            const uint64_t p01 = ptx::ex2_2(ptx::ffma2(ptx::pack2(r[j4 * 4 + 0], r[j4 * 4 + 1]), g2, nm2));
            const uint64_t p23 = ptx::ex2_2(ptx::ffma2(ptx::pack2(r[j4 * 4 + 2], r[j4 * 4 + 3]), g2, nm2));
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

      {
        const float4 bd = ptx::lds128(bound_addr);
        thr = fmaxf(fmaxf(bd.x, bd.y), fmaxf(bd.z, bd.w));
      }
      if (ncols >= CS) {
        // full slice: both 32-column chunks are requested up front
        uint32_t ra[32], rb[32];
        ptx::tmem_ld_32x32(s_tmem, ra);
        ptx::tmem_ld_32x32(s_tmem + 32, rb);
        ptx::tmem_ld_wait();
#ifdef GADM_MATCH_TRACE
        if (t >= 8 && t < 12) tr[t - 8][2] = clock64();
#endif
        process(ra, 0, guard_off{});
        process(rb, 32, guard_off{});
      } else if (ncols > 0) {
        // ragged last tile: 16-column chunks, masked
        uint32_t rc[16];
        const int n16 = (ncols + 15) / 16;
#pragma unroll 1
        for (int c = 0; c < n16; ++c) {
          ptx::tmem_ld_32x16(s_tmem + c * 16, rc);
          ptx::tmem_ld_wait();
          process(rc, c * 16, guard_on{});
        }
      }
      ptx::sts32(bound_addr + sub * 4, vmax);
      ptx::tc_fence_before();
      __syncwarp();
#ifdef GADM_MATCH_TRACE
      if (t >= 8 && t < 12) tr[t - 8][3] = clock64();
#endif
      if (lane == 0) {
        ptx::mbar_arrive(&bars->s_free[acc]);
        ptx::mbar_arrive(&bars->aux_empty[slot]);
      }
    }
#ifdef GADM_MATCH_TRACE
    if (blockIdx.x == 3 && blockIdx.y == 0 && lane == 0 && (warp == 0 || warp == 13))
      for (int i = 0; i < 4; ++i)
        printf("warp %d tile %d: start %lld  wait %lld  ldtm %lld  process %lld\n", warp, 8 + i, tr[i][0] - tr[0][0],
               tr[i][1] - tr[i][0], tr[i][2] - tr[i][1], tr[i][3] - tr[i][2]);
#endif

    // ---- merge the four column slices of every row: slices 1..3 publish, slice 0 combines and writes the outputs
    float lsum = 0.f, ax = 0.f, ay = 0.f, az = 0.f;
    if (kSoft) {
      float e, o;
      ptx::unpack2f(ptx::fadd2(l2a, l2b), e, o); lsum = e + o;
      ptx::unpack2f(ptx::fadd2(ax2a, ax2b), e, o); ax = e + o;
      ptx::unpack2f(ptx::fadd2(ay2a, ay2b), e, o); ay = e + o;
      ptx::unpack2f(ptx::fadd2(az2a, az2b), e, o); az = e + o;
      if (!have_ref) mref = -INFINITY;
    }
    float* xch = reinterpret_cast<float*>(smem_xch);
    if (sub > 0) {
      float* x = xch + ((sub - 1) * BM + row_in_tile) * 8;
      x[0] = vmax; x[1] = __int_as_float(vidx); x[2] = mref; x[3] = lsum;
      x[4] = ax; x[5] = ay; x[6] = az;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
    if (sub == 0 && row_ok) {
      float mm = mref;
#pragma unroll
      for (int s2 = 0; s2 < EPI_SUB - 1; ++s2) {
        const float* x = xch + (s2 * BM + row_in_tile) * 8;
        const float v1 = x[0];
        const int i1 = __float_as_int(x[1]);
        if (v1 > vmax || (v1 == vmax && i1 < vidx)) { vmax = v1; vidx = i1; }
        mm = fmaxf(mm, x[2]);
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
        const float s0 = ptx::ex2_approx(mref - mm);  // exp2(-inf) = 0 for a slice that saw no column
        float l = lsum * s0, sx = ax * s0, sy = ay * s0, sz = az * s0;
#pragma unroll
        for (int s2 = 0; s2 < EPI_SUB - 1; ++s2) {
          const float* x = xch + (s2 * BM + row_in_tile) * 8;
          const float s1 = ptx::ex2_approx(x[2] - mm);
          l = fmaf(x[3], s1, l); sx = fmaf(x[4], s1, sx); sy = fmaf(x[5], s1, sy); sz = fmaf(x[6], s1, sz);
        }
        const float inv = 1.f / l;
        p.weight[grow] = keep ? ptx::ex2_approx(fmaf(vmax, g, -mm)) * inv : 0.f;  // softmax value at the maximum
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
size_t match_smem_bytes(int KB, int stages) {
  return size_t(KB) * A_BLK_BYTES + size_t(stages) * B_STAGE_BYTES + AUX_SLOTS * (kSoft ? 4 : 1) * PLANE_BYTES +
         XCH_BYTES + BOUND_BYTES + sizeof(Barriers) + 1024;
}

}  // namespace

int match_configure() {
  cudaError_t e;
  e = cudaFuncSetAttribute(match_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return set_cuda_error(e);
  e = cudaFuncSetAttribute(match_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return set_cuda_error(e);
  return GADM_OK;
}

template <bool kSoft>
static int match_launch_t(const void* rows, const void* cols, MatchParams p, int Kp, cudaStream_t stream) {
  const int KB = Kp / BK;
  int stages = MAX_STAGES;  // the deepest ring that fits in 227 KB
  while (stages > 2 && match_smem_bytes<kSoft>(KB, stages) > 227 * 1024) --stages;
  if (match_smem_bytes<kSoft>(KB, stages) > 227 * 1024) return GADM_ERR_UNSUPPORTED;
  p.KB = KB; p.stages = stages;

  CUtensorMap tmap_rows, tmap_cols;
  int rc = make_tmap_2b_3d(&tmap_rows, rows, uint64_t(Kp), uint64_t(p.N), uint64_t(p.B), BK, BM, 0);
  if (rc != GADM_OK) return rc;
  rc = make_tmap_2b_3d(&tmap_cols, cols, uint64_t(Kp), uint64_t(p.M), uint64_t(p.n_obj), BK, BN, 0);
  if (rc != GADM_OK) return rc;

  dim3 grid((p.N + BM - 1) / BM, p.B);
  match_kernel<kSoft><<<grid, NUM_THREADS, match_smem_bytes<kSoft>(KB, stages), stream>>>(tmap_rows, tmap_cols, p);
  return check_launch();
}

int match_launch(const void* rows, const float* rinv_rows, const float* pad_sim, const void* cols, const float* aux,
                 const uint8_t* mask, const int32_t* obj_id, int B, int N, int M, int Kp, int n_obj, float gamma,
                 int pad_mode, int mode, int64_t* idx, float* max_sim, float* weight, float* soft_xyz,
                 cudaStream_t stream) {
  MatchParams p;
  p.rinv_rows = rinv_rows; p.pad_sim = pad_sim; p.scales = aux_scales(aux, n_obj, M);
  p.planes = aux_planes(aux, n_obj, M); p.mask = mask; p.obj_id = obj_id;
  p.idx = idx; p.max_sim = max_sim; p.weight = weight; p.soft_xyz = soft_xyz;
  p.B = B; p.N = N; p.M = M; p.KB = 0; p.n_obj = n_obj; p.stages = 0; p.pad_mode = pad_mode;
  p.gamma_log2e = gamma * 1.4426950408889634f;
  if (mode == GADM_MATCH_SOFT) return match_launch_t<true>(rows, cols, p, Kp, stream);
  return match_launch_t<false>(rows, cols, p, Kp, stream);
}

}  // namespace gadm
