// Fused matching head for sm_100a: descriptor similarity (tcgen05 UMMA, bf16 in / fp32 accumulate in TMEM,
// operands staged by TMA) + row-wise argmax / softmax / soft model coordinates, and its training-side twin
// (CircleLoss forward / dL/dsim).  The [N, M] score matrix never leaves TMEM / registers.
//
// Replaces (reference tree): evaluator.py:89-93 (normalize, normalize, matmul, torch.max) and the padded
// variants utils/pvn3d_eval_utils_kpls.py:436-444 / models/geoMatch_DGCNN.py:92-99; models/geoMatch.py:55-83,102-157 +
// models/loss.py:475-490 (circle_kernel).  The softmax weight and soft coordinates are the extension defined in
// oracle/match_oracle.py.
//
// Kernels (what runs when: match_launch_t; measurements and bounds: DESIGN.md 3.1, profiles/SUMMARY_r2.md)
//   match_kernel<soft|argmax, RT>   thread = row; RT row tiles of 128 scene points per CTA.  RT = 1: the two TMEM
//                                   accumulators alternate between model tiles; RT = 2: accumulator r = row tile r with
//                                   8 fixed epilogue warps.  The no-workspace ARGMAX path and SOFT for K' > 128.
//   match_pair_kernel<soft|argmax, cta2>  256 rows per CTA, 128-vertex model tiles, every epilogue thread owns the same
//                                   lane of both row tiles (a per-column constant serves two scores).  Default for
//                                   SOFT.  cta2: the same on CTA pairs (cluster of two, tcgen05.mma.cta_group::2).
//   match_alt_kernel<unit, prune, cta2>  ARGMAX default, PERSISTENT: the (frame, row block, model tile) units of a launch are
//                                   dealt out evenly to one CTA per SM (match_common.cuh: Sched); per segment all 16
//                                   epilogue warps drain one accumulator while the tensor core fills the other; stash in
//                                   the per-SM workspace slot; running maxima shared across the column slices of a row;
//                                   row blocks split over CTAs are merged by the last CTA to arrive.  unit: no
//                                   per-column constant; prune: chunks that cannot win are skipped (BF16N operands);
//                                   cta2 (default): CTA pairs, units = (pair of row blocks, model tile), half a model
//                                   tile per CTA, cta_group::2 MMAs.
//   (circle_sm100.cu)               CircleLoss: masked exponential sums / dL/dsim in the epilogue of the same skeleton.
//
// Common skeleton
//     warp 16     TMA producer: the row tile(s), then model tiles (256 or 128 vertices x 64 k) through an S-stage
//                 mbarrier ring, plus per tile the column scales 1/|m_j| and (SOFT / circle) the coordinate planes
//     warp 17     UMMA issuer: tcgen05.mma into TMEM accumulators, tcgen05.commit.  The loop between two MMAs is a
//                 handful of uniform-datapath instructions: descriptors are 32-bit low words + one constant high word,
//                 the accumulator address a compile-time function of (buffer, row tile)
//     warps 0..15 epilogue.  Per 32-column chunk: tcgen05.ld, score = acc * 1/|m_j| (packed f32x2), 3-input max tree
//                 per 8 columns.  The position of the maximum inside its 8-column group is NOT searched in the loop (a
//                 search is ~70 warp-divergent instructions and some lane of a warp needs one in most chunks): a
//                 thread whose running maximum rises stores the group's 8 scores to a private stash with two
//                 predicated 16-byte stores (shared memory or the L2-resident workspace), and the first maximal index
//                 is looked up there once, after the last tile.  SOFT adds p = 2^(score * g) (no reference exponent:
//                 |gamma| <= 40 keeps the sums inside the fp32 range) and fp32 sums of p and p * xyz.  The column
//                 slices of a row merge through shared memory at the end.
//   What bounds them (measured): ARGMAX -- the SM's shared-memory data pipe (UMMA operand reads + TMA
//   writes fill it with two row tiles per CTA, every epilogue LDS / store wavefront lengthens the tile), then the
//   power cap when sustained; SOFT -- the latency of its 16 epilogue instruction streams (no pipe above 61 %).
#include "match_common.cuh"

namespace gadm {

namespace {

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
  const int nrows = frame_rows(p, b);          // rows of this frame (fewer than N after row compaction)
  if (row0 >= nrows) return;                   // whole CTA, before any barrier or TMEM allocation
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
      // descriptors as 32-bit low words + one constant high word, accumulator address without the shared-memory round
      // trip (the CTA owns all 512 columns: its allocation starts at column 0) -- see match_pair_kernel
      constexpr uint64_t DESC_HI = uint64_t((1024u >> 4) | (1u << 14) | (2u << 29)) << 32;
      if (tmem_base != 0) __trap();
      const uint32_t a_lo0 = ((ptx::smem_u32(smem_a) & 0x3FFFF) >> 4) | 0x10000u;
      const uint32_t b_lo0 = ((ptx::smem_u32(smem_b) & 0x3FFFF) >> 4) | 0x10000u;
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
          const uint32_t d_tmem = uint32_t(acc) * BN;
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
            const uint32_t a_lo = a_lo0 + uint32_t(r * p.KB + kb) * (A_BLK_BYTES >> 4);
            const uint32_t b_lo = b_lo0 + uint32_t(stage) * (B_STAGE_BYTES >> 4);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              ptx::umma_bf16_ss(d_tmem, DESC_HI | (a_lo + k * (UMMA_K * 2 >> 4)), DESC_HI | (b_lo + k * (UMMA_K * 2 >> 4)),
                                idesc, (kb | k) != 0);
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
    const bool row_ok = row < nrows;
    const size_t grow = size_t(b) * p.N + (row_ok ? row : 0);      // operand arrays (compacted order)
    const size_t gout = out_pos(p, b, row_ok ? row : 0);           // outputs
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
          ptx::sts_stash8(up, stash_addr, f);
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
      p.idx[gout] = keep ? best_idx : int64_t(-1);
      p.max_sim[gout] = keep ? best : 0.f;
      if (kSoft) {
        float l = lsum, sx = ax, sy = ay, sz = az;
#pragma unroll
        for (int s2 = 0; s2 < SL - 1; ++s2) {
          const float* x = xch + (s2 * BM + row_in_tile) * 8;
          l += x[3]; sx += x[4]; sy += x[5]; sz += x[6];
        }
        const float inv = 1.f / l;
        p.weight[gout] = keep ? ptx::ex2_approx(vmax * g) * inv : 0.f;  // softmax value at the maximum
        p.soft_xyz[gout * 3 + 0] = keep ? sx * inv : 0.f;
        p.soft_xyz[gout * 3 + 1] = keep ? sy * inv : 0.f;
        p.soft_xyz[gout * 3 + 2] = keep ? sz * inv : 0.f;
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
// Pairs (of the 4 per 4 columns x 2 rows) whose 2^x runs on the FMA pipe (ptx::ex2_2_poly) instead of MUFU.EX2.
// Measured at the BASELINE shape: 0 -> 0.353 ms, 1 -> 0.377 ms, 2 -> 0.424 ms, 3 -> 0.473 ms: every exponential moved
// off the XU pipe makes the kernel slower, i.e. MUFU (59 % busy) is not what bounds the epilogue -- its 16 warps are
// latency-bound on their own instruction streams (5.9 cycles per instruction and warp), so instructions are what
// counts.  Kept at 0; the switch stays for the record and for other shapes.
#ifndef GADM_SOFT_POLY
#define GADM_SOFT_POLY 0
#endif
constexpr int P_MAX_KB = 4;                    // K blocks of a model tile (the issue loop is unrolled over them)
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

template <bool kSoft, bool kCta2>
__global__ void __launch_bounds__(NUM_THREADS, 1)
match_pair_kernel(const __grid_constant__ CUtensorMap tmap_rows, const __grid_constant__ CUtensorMap tmap_cols,
                  const MatchParams p) {
  constexpr int AUX_BYTES = kSoft ? 4 * P_PLANE_BYTES : P_PLANE_BYTES;
  // kCta2: the CTA is one half of a pair (cluster of two, cta_group::2 MMAs with M = 256).  It holds its own 256 rows
  // and HALF of every model tile (64 of the 128 vertices): the MMA reads 6 KB instead of 8 KB of operands per 64
  // cycles from this SM's shared memory and TMA writes half as much into it -- the data pipe the epilogue shares.
  constexpr int STAGE_BYTES = kCta2 ? PB_STAGE_BYTES / 2 : PB_STAGE_BYTES;
  constexpr int STAGE_ROWS = kCta2 ? PBN / 2 : PBN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                                    // [2][KB] blocks of 128 rows x 64 k
  uint8_t* smem_b = smem_a + 2 * p.KB * A_BLK_BYTES;
  uint8_t* smem_aux = smem_b + p.stages * STAGE_BYTES;  // per slot: [1/|m| x128 | x | y | z]
  uint8_t* smem_stash = smem_aux + AUX_SLOTS * AUX_BYTES;
  PairBarriers* bars = reinterpret_cast<PairBarriers*>(smem_stash + P_STASH_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int row0 = blockIdx.x * (2 * BM);
  const int nrows = frame_rows(p, b);          // rows of this frame (fewer than N after row compaction)
  const uint32_t rank = kCta2 ? ptx::cluster_ctarank() : 0;   // == blockIdx.x & 1
  // whole CTA (kCta2: whole pair -- a peer without rows still supplies its half of the model tiles), before any
  // barrier or TMEM allocation
  if ((kCta2 ? (blockIdx.x & ~1u) * (2 * BM) : row0) >= nrows) return;
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
      ptx::mbar_init(&bars->s_free[a], kCta2 ? 2 * EPI_WARPS : EPI_WARPS);   // the leader hears both CTAs
    }
    for (int a = 0; a < AUX_SLOTS; ++a) {
      ptx::mbar_init(&bars->aux_full[a], 1);
      ptx::mbar_init(&bars->aux_empty[a], EPI_WARPS);
    }
    ptx::fence_mbar_init();
  }
  if (warp == EPI_WARPS + 1) {
    if (kCta2) { ptx::tmem_alloc_pair(&bars->tmem_base, TMEM_COLS); ptx::tmem_relinquish_pair(); }
    else       { ptx::tmem_alloc(&bars->tmem_base, TMEM_COLS); ptx::tmem_relinquish(); }
  }
  ptx::tc_fence_before();
  if (kCta2) ptx::cluster_sync(); else __syncthreads();   // the peer's barriers are initialised too
  ptx::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  // barriers of the MMA issuer (the leader's), as cluster addresses
  const uint32_t a_full_ldr = kCta2 ? ptx::mapa(ptx::smem_u32(&bars->a_full), 0) : 0;

  if (warp == EPI_WARPS) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      if (rank == 0) ptx::mbar_arrive_expect_tx(&bars->a_full, (kCta2 ? 4 : 2) * p.KB * A_BLK_BYTES);
      for (int r = 0; r < 2; ++r)
        for (int kb = 0; kb < p.KB; ++kb) {    // rows >= N are zero-filled by TMA
          if (kCta2)
            ptx::tma_load_3d_pair(smem_a + (r * p.KB + kb) * A_BLK_BYTES, &tmap_rows, a_full_ldr, kb * BK,
                                  row0 + r * BM, b);
          else
            ptx::tma_load_3d(smem_a + (r * p.KB + kb) * A_BLK_BYTES, &tmap_rows, &bars->a_full, kb * BK,
                             row0 + r * BM, b);
        }
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
          // (kCta2: both halves complete on the leader's barrier; the peer's bytes may land before the leader's
          // expect_tx of the same phase -- the phase cannot end before the leader's arrive)
          if (rank == 0) ptx::mbar_arrive_expect_tx(&bars->full[stage], PB_STAGE_BYTES);
          if (kCta2)
            ptx::tma_load_3d_pair(smem_b + stage * STAGE_BYTES, &tmap_cols,
                                  ptx::mapa(ptx::smem_u32(&bars->full[stage]), 0), kb * BK,
                                  t * PBN + int(rank) * STAGE_ROWS, obj);
          else
            ptx::tma_load_3d(smem_b + stage * STAGE_BYTES, &tmap_cols, &bars->full[stage], kb * BK, t * PBN, obj);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == EPI_WARPS + 1) {
    // ============================== UMMA issuer ==============================
    // (one lane issues.  Issuing from the whole warp with an elected lane keeps the operands in uniform registers and
    // shortens the scalar code between MMAs, but 32 polling lanes cost the epilogue warps of the same scheduler more
    // than that saves: 0.403 ms against 0.381 ms at the BASELINE shape)
    if (lane == 0 && rank == 0) {
      // The issue loop is the tensor pipe's clock: a 128x128x16 MMA executes in 64 cycles, so everything between two
      // tcgen05.mma of this one thread has to stay well below that.  Descriptors are therefore kept as 32-bit low
      // words (start address >> 4 | LBO; every shared-memory address >> 4 fits the 14-bit field, so + 2 steps k
      // without a mask) next to one constant high word, and the accumulator address is a compile-time function of
      // (buf, r): the CTA owns all 512 columns, so its allocation starts at column 0 (checked below).  With the
      // address read back from shared memory the compiler emits an ELECT / R2UR.BROADCAST waterfall per MMA and the
      // loop issues one MMA per ~130 cycles (0.215 ms for the bare TMA -> MMA pipeline at the BASELINE shape).
      constexpr uint32_t idesc = ptx::umma_idesc_bf16_f32(kCta2 ? 2 * BM : BM, PBN);
      constexpr uint64_t DESC_HI = uint64_t((1024u >> 4) | (1u << 14) | (2u << 29)) << 32;   // SBO | version | SW128
      if (tmem_base != 0) __trap();
      const uint32_t a_lo0 = ((ptx::smem_u32(smem_a) & 0x3FFFF) >> 4) | 0x10000u;
      const uint32_t b_lo0 = ((ptx::smem_u32(smem_b) & 0x3FFFF) >> 4) | 0x10000u;
      auto mma = [&](uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t acc) {
        if (kCta2) ptx::umma_bf16_ss_pair(d, DESC_HI | a_lo, DESC_HI | b_lo, idesc, acc);
        else ptx::umma_bf16_ss(d, DESC_HI | a_lo, DESC_HI | b_lo, idesc, acc);
      };
      ptx::mbar_wait(&bars->a_full, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < num_tiles; ++t) {
        const int buf = t & 1;
        ptx::mbar_wait_sleep(&bars->s_free[buf], ((uint32_t(t) >> 1) & 1) ^ 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < P_MAX_KB; ++kb) {
          if (kb >= p.KB) break;
          ptx::mbar_wait_sleep(&bars->full[stage], phase);
          ptx::tc_fence_after();
          const uint32_t b_lo = b_lo0 + uint32_t(stage) * (STAGE_BYTES >> 4);
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const uint32_t a_lo = a_lo0 + uint32_t(r * p.KB + kb) * (A_BLK_BYTES >> 4);
            const uint32_t d_tmem = uint32_t(buf * 2 + r) * PBN;
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              mma(d_tmem, a_lo + k * (UMMA_K * 2 >> 4), b_lo + k * (UMMA_K * 2 >> 4), (kb | k) != 0);
          }
          if (kCta2) ptx::umma_commit_pair(&bars->empty[stage]);   // frees the stage in both CTAs
          else ptx::umma_commit(&bars->empty[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (kCta2) ptx::umma_commit_pair(&bars->s_full[buf]);
        else ptx::umma_commit(&bars->s_full[buf]);
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
      row_ok[r] = row[r] < nrows;
      grow[r] = size_t(b) * p.N + (row_ok[r] ? row[r] : 0);
      rs[r] = row_ok[r] ? p.rinv_rows[grow[r]] : 0.f;
      g[r] = p.gamma_log2e * rs[r];
    }
    const uint32_t lane_base = tmem_base + (uint32_t(q * 32) << 16) + sub * P_CS;
    const uint32_t stash_addr = ptx::smem_u32(smem_stash) + threadIdx.x * 16;   // row r: + r * 2 * STASH_PLANE
    const uint32_t s_free_ldr = kCta2 ? ptx::mapa(ptx::smem_u32(&bars->s_free[0]), 0) : 0;

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
#ifndef GADM_DBG_NOMAX   // (GADM_DBG_*: timing-only ablation builds, tools/ablate_soft.sh; results are wrong)
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
#ifndef GADM_DBG_NOSTASH
            ptx::sts_stash8(up, stash_addr + rr * 2 * STASH_PLANE, f);
#endif
            vgrp[rr] = up ? col_base + col0 + h * GRP : vgrp[rr];
            vmax[rr] = up ? gm : vmax[rr];
          }
        }
#endif
        if (kSoft) {
          // p = 2^(v*g), no reference exponent (see match_kernel)
          const uint64_t g20 = ptx::pack2f(g[0], g[0]), g21 = ptx::pack2f(g[1], g[1]);
#pragma unroll
          for (int j4 = 0; j4 < W / 4; ++j4) {
#ifdef GADM_DBG_NOXYZ
            const float4 X = make_float4(1.f, 2.f, 3.f, 4.f), Y = X, Z = X;
#else
            const float4 X = ptx::lds128(sc + P_PLANE_BYTES + j4 * 16);
            const float4 Y = ptx::lds128(sc + 2 * P_PLANE_BYTES + j4 * 16);
            const float4 Z = ptx::lds128(sc + 3 * P_PLANE_BYTES + j4 * 16);
#endif
            const uint64_t X01 = ptx::pack2f(X.x, X.y), X23 = ptx::pack2f(X.z, X.w);
            const uint64_t Y01 = ptx::pack2f(Y.x, Y.y), Y23 = ptx::pack2f(Y.z, Y.w);
            const uint64_t Z01 = ptx::pack2f(Z.x, Z.y), Z23 = ptx::pack2f(Z.z, Z.w);
#ifdef GADM_DBG_NOEXP
#define GADM_EX2(x) (x)
#else
#define GADM_EX2(x) ptx::ex2_2(x)
#endif
            // GADM_SOFT_POLY of the 4 pairs take the FMA-pipe polynomial instead of MUFU.EX2 (never in the ragged
            // tile, whose masked columns are -inf)
            const uint64_t pa0 = GADM_EX2(ptx::fmul2(v[0][j4 * 2 + 0], g20));
            const uint64_t pb0 = (!kGuard && GADM_SOFT_POLY >= 1) ? ptx::ex2_2_poly(v[0][j4 * 2 + 1], g20)
                                                                  : GADM_EX2(ptx::fmul2(v[0][j4 * 2 + 1], g20));
            const uint64_t pa1 = (!kGuard && GADM_SOFT_POLY >= 3) ? ptx::ex2_2_poly(v[1][j4 * 2 + 0], g21)
                                                                  : GADM_EX2(ptx::fmul2(v[1][j4 * 2 + 0], g21));
            const uint64_t pb1 = (!kGuard && GADM_SOFT_POLY >= 2) ? ptx::ex2_2_poly(v[1][j4 * 2 + 1], g21)
                                                                  : GADM_EX2(ptx::fmul2(v[1][j4 * 2 + 1], g21));
#undef GADM_EX2
            l2[0] = ptx::fadd2(l2[0], ptx::fadd2(pa0, pb0));
            l2[1] = ptx::fadd2(l2[1], ptx::fadd2(pa1, pb1));
#ifndef GADM_DBG_NOSUMS
            ax2[0] = ptx::ffma2(pb0, X23, ptx::ffma2(pa0, X01, ax2[0]));
            ax2[1] = ptx::ffma2(pb1, X23, ptx::ffma2(pa1, X01, ax2[1]));
            ay2[0] = ptx::ffma2(pb0, Y23, ptx::ffma2(pa0, Y01, ay2[0]));
            ay2[1] = ptx::ffma2(pb1, Y23, ptx::ffma2(pa1, Y01, ay2[1]));
            az2[0] = ptx::ffma2(pb0, Z23, ptx::ffma2(pa0, Z01, az2[0]));
            az2[1] = ptx::ffma2(pb1, Z23, ptx::ffma2(pa1, Z01, az2[1]));
#endif
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
          // (the ragged last tile has its own loop: with both variants in one unrolled body the accumulators are
          // shuffled between their register assignments after every half tile, ~20 MOVs per 32 scores)
          if (ncols >= P_CS) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              uint32_t ra[16], rb[16];
              ptx::tmem_ld_32x16(s_tmem0 + h * 16, ra);
              ptx::tmem_ld_32x16(s_tmem1 + h * 16, rb);
              ptx::tmem_ld_wait();
              process(ra, rb, h * 16, guard_off{});
            }
          } else {
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
              if (ncols - h * 16 <= 0) break;
              uint32_t ra[16], rb[16];
              ptx::tmem_ld_32x16(s_tmem0 + h * 16, ra);
              ptx::tmem_ld_32x16(s_tmem1 + h * 16, rb);
              ptx::tmem_ld_wait();
              process(ra, rb, h * 16, guard_on{});
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kCta2) ptx::mbar_arrive_cluster(s_free_ldr + buf * 8);
        else ptx::mbar_arrive(&bars->s_free[buf]);
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
        const size_t go = out_pos(p, b, row[rr]);
        p.idx[go] = keep ? best_idx : int64_t(-1);
        p.max_sim[go] = keep ? best : 0.f;
        if (kSoft) {
          float l = lsum[rr], sx = ax[rr], sy = ay[rr], sz = az[rr];
#pragma unroll
          for (int s2 = 0; s2 < 3; ++s2) {
            const float* x = xch + (s2 * BM + row_in_tile) * 8;
            l += x[3]; sx += x[4]; sy += x[5]; sz += x[6];
          }
          const float inv = 1.f / l;
          p.weight[go] = keep ? ptx::ex2_approx(vm * g[rr]) * inv : 0.f;
          p.soft_xyz[go * 3 + 0] = keep ? sx * inv : 0.f;
          p.soft_xyz[go * 3 + 1] = keep ? sy * inv : 0.f;
          p.soft_xyz[go * 3 + 2] = keep ? sz * inv : 0.f;
        }
      }
    }
  }

  ptx::tc_fence_before();
  if (kCta2) ptx::cluster_sync(); else __syncthreads();   // the pair's MMAs read both CTAs' shared memory
  if (warp == EPI_WARPS + 1) {
    ptx::tc_fence_after();
    if (kCta2) ptx::tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <bool kSoft>
size_t match_pair_smem_bytes(int KB, int stages, bool cta2 = false) {
  return size_t(2) * KB * A_BLK_BYTES + size_t(stages) * (cta2 ? PB_STAGE_BYTES / 2 : PB_STAGE_BYTES) +
         AUX_SLOTS * (kSoft ? 4 : 1) * P_PLANE_BYTES + P_STASH_BYTES + sizeof(PairBarriers) + 1024;
}
template <bool kSoft>
int match_pair_stages(int KB, bool cta2 = false) {
  int stages = P_MAX_STAGES;
  while (stages > 0 && match_pair_smem_bytes<kSoft>(KB, stages, cta2) > 227 * 1024) --stages;
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
// kCta2 (default; match.alt_cta2 = 0 switches it off): the CTAs run as pairs (clusters of two, cta_group::2 MMAs with M = 256).  A pair shares its
// units -- (pair of row blocks, model tile) --, each CTA keeps its own row block, row tiles and accumulators but only
// HALF of every model tile (128 of the 256 vertices) in shared memory: a third fewer operand wavefronts and half the TMA
// writes on the shared-memory data pipe that bounds this kernel.  The leader issues every MMA and hears both CTAs'
// epilogues on its s_free barriers; commits are multicast to both CTAs.
template <bool kUnit, bool kPrune, bool kCta2>
__global__ void __launch_bounds__(NUM_THREADS, 1)
match_alt_kernel(const __grid_constant__ CUtensorMap tmap_rows, const __grid_constant__ CUtensorMap tmap_cols,
                 const MatchParams p) {
  constexpr int RT = 2;
  constexpr int AUX_BYTES = PLANE_BYTES;
  constexpr int STAGE_BYTES = kCta2 ? B_STAGE_BYTES / 2 : B_STAGE_BYTES;
  constexpr int STAGE_ROWS = kCta2 ? BN / 2 : BN;
  constexpr int SL = 4;                      // column slices per row
  constexpr int CS = BN / SL;                // 64 columns per slice

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                                    // [RT][KB] blocks of 128 rows x 64 k
  uint8_t* smem_b = smem_a + RT * p.KB * A_BLK_BYTES;
  uint8_t* smem_aux = smem_b + p.stages * STAGE_BYTES;   // per slot: 1/|m| x256
  float* smem_xmax = reinterpret_cast<float*>(smem_aux + AUX_SLOTS * AUX_BYTES);   // [RT][SL][128] running maxima
  float2* smem_xch = reinterpret_cast<float2*>(smem_xmax + RT * SL * BM);           // [RT][SL - 1][128] slice merge
  Barriers* bars = reinterpret_cast<Barriers*>(smem_xch + RT * (SL - 1) * BM);
  int* smem_prefix = reinterpret_cast<int*>(bars + 1);                             // [B + 1] when row counts are given

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // this CTA's (pair's) share of the (frame, row block, model tile) units; a pair's row block is 512 rows, rank r of
  // the pair owns rows [256 r, 256 r + 256) of it
  const uint32_t rank = kCta2 ? ptx::cluster_ctarank() : 0;     // == blockIdx.x & 1
  const int cidx = kCta2 ? int(blockIdx.x >> 1) : int(blockIdx.x);
  constexpr int BLOCK_ROWS = BM * RT * (kCta2 ? 2 : 1);
  const Sched sch = sched_build(p, smem_prefix, BLOCK_ROWS, kCta2 ? int(gridDim.x >> 1) : int(gridDim.x));
  if (cidx >= sch.nc) return;                  // (row compaction can leave fewer units than CTAs; pair-uniform)
  const long long u_begin = sch.begin(cidx), u_end = sch.begin(cidx + 1);

  if (warp == EPI_WARPS && lane == 0) {
    ptx::prefetch_tensormap(&tmap_rows);
    ptx::prefetch_tensormap(&tmap_cols);
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(&bars->full[s], 1);
      ptx::mbar_init(&bars->empty[s], 1);
    }
    ptx::mbar_init(&bars->a_full, 1);
    ptx::mbar_init(&bars->a_free, 1);
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&bars->s_full[a], 1);
      ptx::mbar_init(&bars->s_free[a], kCta2 ? 2 * EPI_WARPS : EPI_WARPS);   // every epilogue warp (of both CTAs)
    }                                                                          // drains every accumulator
    for (int a = 0; a < AUX_SLOTS; ++a) {
      ptx::mbar_init(&bars->aux_full[a], 1);
      ptx::mbar_init(&bars->aux_empty[a], EPI_WARPS);
    }
    ptx::fence_mbar_init();
  }
  if (warp == EPI_WARPS + 1) {
    if (kCta2) { ptx::tmem_alloc_pair(&bars->tmem_base, TMEM_COLS); ptx::tmem_relinquish_pair(); }
    else       { ptx::tmem_alloc(&bars->tmem_base, TMEM_COLS); ptx::tmem_relinquish(); }
  }
  ptx::tc_fence_before();
  if (kCta2) ptx::cluster_sync(); else __syncthreads();   // the peer's barriers are initialised too
  ptx::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == EPI_WARPS) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      const uint32_t a_full_ldr = kCta2 ? ptx::mapa(ptx::smem_u32(&bars->a_full), 0) : 0;
      int stage = 0;
      uint32_t phase = 0, n = 0, seg = 0;      // n: model tiles this CTA has started so far (all segments)
      for (long long u = u_begin; u < u_end; ++seg) {
        const int rbg = int(u / p.T), ta = int(u - (long long)rbg * p.T);
        const int tb = int(min((long long)p.T, ta + (u_end - u)));
        int b, rb;
        sch.locate(rbg, b, rb);
        const int row0 = rb * BLOCK_ROWS + int(rank) * (BM * RT);
        const int obj = frame_object(p, b);
        // the row tiles of the previous segment are dead once its last MMA has completed
        if (seg > 0) ptx::mbar_wait_sleep(&bars->a_free, (seg - 1) & 1);
        if (rank == 0) ptx::mbar_arrive_expect_tx(&bars->a_full, (kCta2 ? 2 : 1) * RT * p.KB * A_BLK_BYTES);
        for (int r = 0; r < RT; ++r)
          for (int kb = 0; kb < p.KB; ++kb) {    // rows >= N are zero-filled by TMA
            if (kCta2)
              ptx::tma_load_3d_pair(smem_a + (r * p.KB + kb) * A_BLK_BYTES, &tmap_rows, a_full_ldr, kb * BK,
                                    row0 + r * BM, b);
            else
              ptx::tma_load_3d(smem_a + (r * p.KB + kb) * A_BLK_BYTES, &tmap_rows, &bars->a_full, kb * BK,
                               row0 + r * BM, b);
          }
        const float* sc_tab = p.scales + size_t(obj) * p.M;
        for (int t = ta; t < tb; ++t, ++n) {
          const int slot = n % AUX_SLOTS;
          const uint32_t use = n / AUX_SLOTS;
          const uint32_t bytes = uint32_t(min(BN, p.M - t * BN)) * 4;   // M % 8 == 0: a multiple of 16
          if (!kUnit) {
            ptx::mbar_wait_sleep(&bars->aux_empty[slot], (use & 1) ^ 1);
            ptx::mbar_arrive_expect_tx(&bars->aux_full[slot], bytes);
            ptx::bulk_load_1d(smem_aux + slot * AUX_BYTES, sc_tab + size_t(t) * BN, bytes, &bars->aux_full[slot]);
          }
          for (int kb = 0; kb < p.KB; ++kb) {
            ptx::mbar_wait_sleep(&bars->empty[stage], phase ^ 1);
            // (kCta2: both halves complete on the leader's barrier, see match_pair_kernel)
            if (rank == 0) ptx::mbar_arrive_expect_tx(&bars->full[stage], B_STAGE_BYTES);
            if (kCta2)
              ptx::tma_load_3d_pair(smem_b + stage * STAGE_BYTES, &tmap_cols,
                                    ptx::mapa(ptx::smem_u32(&bars->full[stage]), 0), kb * BK,
                                    t * BN + int(rank) * STAGE_ROWS, obj);
            else
              ptx::tma_load_3d(smem_b + stage * STAGE_BYTES, &tmap_cols, &bars->full[stage], kb * BK, t * BN, obj);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
        u += tb - ta;
      }
    }
  } else if (warp == EPI_WARPS + 1) {
    // ============================== UMMA issuer ==============================
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16_f32(kCta2 ? 2 * BM : BM, BN);
      // descriptors as 32-bit low words + one constant high word, accumulator address a function of r alone (the CTA
      // owns all 512 columns: its allocation starts at column 0) -- see match_pair_kernel
      constexpr uint64_t DESC_HI = uint64_t((1024u >> 4) | (1u << 14) | (2u << 29)) << 32;
      if (tmem_base != 0) __trap();
      const uint32_t a_lo0 = ((ptx::smem_u32(smem_a) & 0x3FFFF) >> 4) | 0x10000u;
      const uint32_t b_lo0 = ((ptx::smem_u32(smem_b) & 0x3FFFF) >> 4) | 0x10000u;
      int stage0 = 0;            // ring position of the tile's first K block
      uint32_t phase0 = 0, n = 0, seg = 0;
      for (long long u = u_begin; u < u_end; ++seg) {
        const int ntiles = int(min((long long)p.T - (u % p.T), u_end - u));
        ptx::mbar_wait(&bars->a_full, seg & 1);
        ptx::tc_fence_after();
        for (int t = 0; t < ntiles; ++t, ++n) {
#pragma unroll
          for (int r = 0; r < RT; ++r) {
            ptx::mbar_wait_sleep(&bars->s_free[r], (n & 1) ^ 1);
            ptx::tc_fence_after();
            const uint32_t d_tmem = r * BN;
            int stage = stage0;
            uint32_t phase = phase0;
            for (int kb = 0; kb < p.KB; ++kb) {
              if (r == 0) {      // the stages of this tile stay resident until the last row tile has used them
                ptx::mbar_wait_sleep(&bars->full[stage], phase);
                ptx::tc_fence_after();
              }
              const uint32_t a_lo = a_lo0 + uint32_t(r * p.KB + kb) * (A_BLK_BYTES >> 4);
              const uint32_t b_lo = b_lo0 + uint32_t(stage) * (STAGE_BYTES >> 4);
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k) {
                if (kCta2)
                  ptx::umma_bf16_ss_pair(d_tmem, DESC_HI | (a_lo + k * (UMMA_K * 2 >> 4)),
                                         DESC_HI | (b_lo + k * (UMMA_K * 2 >> 4)), idesc, (kb | k) != 0);
                else
                  ptx::umma_bf16_ss(d_tmem, DESC_HI | (a_lo + k * (UMMA_K * 2 >> 4)),
                                    DESC_HI | (b_lo + k * (UMMA_K * 2 >> 4)), idesc, (kb | k) != 0);
              }
              if (r == RT - 1) {   // frees the stage (in both CTAs) once these MMAs have read it
                if (kCta2) ptx::umma_commit_pair(&bars->empty[stage]); else ptx::umma_commit(&bars->empty[stage]);
              }
              if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
            if (kCta2) ptx::umma_commit_pair(&bars->s_full[r]); else ptx::umma_commit(&bars->s_full[r]);   // tile complete
            if (r == RT - 1) { stage0 = stage; phase0 = phase; }
          }
        }
        if (kCta2) ptx::umma_commit_pair(&bars->a_free); else ptx::umma_commit(&bars->a_free);   // row tiles are dead
        u += ntiles;
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
    uint8_t* stash = p.stash + size_t(ptx::smid()) * STASH_SLOT_BYTES + threadIdx.x * 16;
    const int rbg_first = int(u_begin / p.T);        // the row block this CTA's first segment belongs to
    const uint32_t s_free_ldr = kCta2 ? ptx::mapa(ptx::smem_u32(&bars->s_free[0]), 0) : 0;

    uint32_t n = 0;
    for (long long u = u_begin; u < u_end;) {
      const int rbg = int(u / p.T), ta = int(u - (long long)rbg * p.T);
      const int tb = int(min((long long)p.T, ta + (u_end - u)));
      int b, rb;
      sch.locate(rbg, b, rb);
      const int row0 = rb * BLOCK_ROWS + int(rank) * (BM * RT);
      const int nrows = frame_rows(p, b);
      const int obj = frame_object(p, b);
      u += tb - ta;

      float vmax[RT] = {-INFINITY, -INFINITY};   // running maximum of this thread's slice of its row of row tile r
      int vgrp[RT] = {0, 0};                     // first column of the 8-column group that first reached it

      for (int t = ta; t < tb; ++t, ++n) {
        const int slot = n % AUX_SLOTS;
        const int ncols = min(BN, p.M - t * BN) - sub * CS;   // valid columns of this slice (may be <= 0)
        const uint32_t sc_addr = ptx::smem_u32(smem_aux + slot * AUX_BYTES) + sub * CS * 4;
        const int col_base = t * BN + sub * CS;
#pragma unroll
        for (int r = 0; r < RT; ++r) {
          if (!((kUnit || r > 0 || ptx::mbar_try_wait(&bars->aux_full[slot], (n / AUX_SLOTS) & 1)) &
                ptx::mbar_try_wait(&bars->s_full[r], n & 1))) {
            if (!kUnit && r == 0) ptx::mbar_wait_sleep(&bars->aux_full[slot], (n / AUX_SLOTS) & 1);
            ptx::mbar_wait_sleep(&bars->s_full[r], n & 1);
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

#ifdef GADM_DBG_NOEPI
          if (false)
#endif
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
            if (kCta2) ptx::mbar_arrive_cluster(s_free_ldr + r * 8); else ptx::mbar_arrive(&bars->s_free[r]);
            if (!kUnit && r == RT - 1) ptx::mbar_arrive(&bars->aux_empty[slot]);
          }
        }
        // A row has 4 tracks (one per column slice, in 4 different warps).  Alone, each raises its running maximum
        // (two stash stores, wavefronts of the data pipe the MMAs saturate) ~H(n) times; after the segment's tiles
        // 0, 1, 3, 7, 15 the slices publish their maxima and adopt the row's: a track that adopts a larger maximum
        // than its own gives up its record (its index becomes a sentinel that loses every tie -- the holder sits at
        // an earlier column), and from then on only values above the ROW's maximum so far are recorded.
        const int tl = t - ta;
        if ((tl & (tl + 1)) == 0 && tl < 16 && t + 1 < tb) {
#pragma unroll
          for (int r = 0; r < RT; ++r) smem_xmax[(r * SL + sub) * BM + row_in_tile] = vmax[r];
          asm volatile("bar.sync 2, %0;" ::"n"(EPI_WARPS * 32) : "memory");
#pragma unroll
          for (int r = 0; r < RT; ++r) {
            float m = vmax[r];
#pragma unroll
            for (int s2 = 0; s2 < SL; ++s2) m = fmaxf(m, smem_xmax[(r * SL + s2) * BM + row_in_tile]);
            if (vmax[r] < m) { vmax[r] = m; vgrp[r] = NO_RECORD; }
          }
          asm volatile("bar.sync 2, %0;" ::"n"(EPI_WARPS * 32) : "memory");
        }
      }

      // ---- end of the segment.  Per row tile: first maximal index of this slice from the stash (own stores, read
      // back through L2), then the merge of the 4 column slices through shared memory
      int vidx[RT];
#pragma unroll
      for (int r = 0; r < RT; ++r) {
        vidx[r] = NO_RECORD;
        if (vmax[r] > -INFINITY && vgrp[r] != NO_RECORD) {
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
        if (sub > 0) smem_xch[(r * (SL - 1) + sub - 1) * BM + row_in_tile] = make_float2(vmax[r], __int_as_float(vidx[r]));
      }
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
      if (sub == 0) {
        float vm[RT];
        int vi[RT];
#pragma unroll
        for (int r = 0; r < RT; ++r) {
          vm[r] = vmax[r];
          vi[r] = vidx[r];
#pragma unroll
          for (int s2 = 0; s2 < SL - 1; ++s2) {
            const float2 x = smem_xch[(r * (SL - 1) + s2) * BM + row_in_tile];
            const int i1 = __float_as_int(x.y);
            if (x.x > vm[r] || (x.x == vm[r] && i1 < vi[r])) { vm[r] = x.x; vi[r] = i1; }
          }
        }
        bool finish = ta == 0 && tb == p.T;          // the whole row of model tiles was this segment
        if (!finish) {
          // Publish this segment's result; whoever arrives last at the row block merges all of its segments.
          float2* part = reinterpret_cast<float2*>(p.partial) +
                         (size_t(blockIdx.x) * 2 + (rbg == rbg_first ? 0 : 1)) * PART_ROWS;
#pragma unroll
          for (int r = 0; r < RT; ++r) part[r * BM + row_in_tile] = make_float2(vm[r], __int_as_float(vi[r]));
          __threadfence();
          asm volatile("bar.sync 3, 128;" ::: "memory");
          if (threadIdx.x == 0) {
            // (kCta2: c counts pairs; rank r of every pair holds rows [256 r, 256 r + 256) of the row block, so the two
            // ranks merge independently: own counter, own partial slots)
            const int c_lo = sch.cta_of((long long)rbg * p.T), c_hi = sch.cta_of((long long)(rbg + 1) * p.T - 1);
            const unsigned int old = atomicAdd(&p.seg_count[kCta2 ? c_lo * 2 + int(rank) : c_lo], 1u);
            bars->merge_lo = old == unsigned(c_hi - c_lo) ? c_lo : -1;
            bars->merge_hi = c_hi;
          }
          asm volatile("bar.sync 3, 128;" ::: "memory");
          const int c_lo = bars->merge_lo, c_hi = bars->merge_hi;
          if (c_lo >= 0) {
            __threadfence();
#pragma unroll
            for (int r = 0; r < RT; ++r) { vm[r] = -INFINITY; vi[r] = NO_RECORD; }
            for (int c = c_lo; c <= c_hi; ++c) {     // ascending columns: on ties the earlier segment wins
              const float2* q2 = reinterpret_cast<const float2*>(p.partial) +
                                 (size_t(kCta2 ? c * 2 + int(rank) : c) * 2 + (int(sch.begin(c) / p.T) == rbg ? 0 : 1)) *
                                     PART_ROWS;
#pragma unroll
              for (int r = 0; r < RT; ++r) {
                float x, y;
                ptx::unpack2f(ptx::ldg_cg64(q2 + r * BM + row_in_tile), x, y);
                const int i1 = __float_as_int(y);
                if (x > vm[r] || (x == vm[r] && i1 < vi[r])) { vm[r] = x; vi[r] = i1; }
              }
            }
            finish = true;
          }
        }
        if (finish) {
#pragma unroll
          for (int r = 0; r < RT; ++r) {
            const int row = row0 + r * BM + row_in_tile;
            if (row >= nrows) continue;
            const size_t grow = size_t(b) * p.N + row;
            const size_t gout = out_pos(p, b, row);
            // kUnit searched with unit column norms; the winner's similarity is reported with its true scale
            if (kUnit) vm[r] *= p.scales[size_t(obj) * p.M + vi[r]];
            const bool keep = p.mask == nullptr || p.mask[grow] != 0;
            float best = vm[r] * p.rinv_rows[grow];
            int64_t best_idx = vi[r];
            if (p.pad_mode != GADM_PAD_NONE) {
              const float ps = p.pad_sim[grow];
              if (ps > best) { best = ps; best_idx = p.M; }  // pad column is the last one: wins only if strictly larger
            }
            p.idx[gout] = keep ? best_idx : int64_t(-1);
            p.max_sim[gout] = keep ? best : 0.f;
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  if (kCta2) ptx::cluster_sync(); else __syncthreads();   // the pair's MMAs read both CTAs' shared memory
  if (warp == EPI_WARPS + 1) {
    ptx::tc_fence_after();
    if (kCta2) ptx::tmem_dealloc_pair(tmem_base, TMEM_COLS); else ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

constexpr int SCHED_MAX_FRAMES = 2048;          // frames a persistent kernel can schedule from device-side row counts
inline size_t match_alt_smem_bytes(int KB, int stages, bool cta2 = false) {
  return size_t(2) * KB * A_BLK_BYTES + size_t(stages) * (cta2 ? B_STAGE_BYTES / 2 : B_STAGE_BYTES) +
         AUX_SLOTS * PLANE_BYTES + 2 * 4 * BM * 4 + 2 * 3 * BM * 8 + sizeof(Barriers) +
         (SCHED_MAX_FRAMES + 1) * sizeof(int) + 1024;
}
inline int match_alt_stages(int KB, bool cta2 = false) {
  int stages = MAX_STAGES;
  while (stages > 0 && match_alt_smem_bytes(KB, stages, cta2) > 227 * 1024) --stages;
  return stages;
}


// ---------------------------------------------------------------------------------------------------------------
// Host side: per-device facts read once by gadm_init(), kernel selection, launch.
constexpr int MAX_DEVICES = 64;
struct DeviceInfo {
  int sms = 0;          // SM count
  int slots = 0;        // max(SM count, %nsmid): the per-SM workspace slots are indexed by %smid
};
DeviceInfo g_dev[MAX_DEVICES];

__device__ unsigned int g_nsmid_out;
__global__ void nsmid_kernel() {
  unsigned int v;
  asm volatile("mov.u32 %0, %%nsmid;" : "=r"(v));
  g_nsmid_out = v;
}

// Kernel-selection switches (gadm_config_set; profiling and tests).  -1 = automatic.
struct MatchConfig {
  int alt = -1;         // match.alt   1 / 0: allow / forbid the alternating ARGMAX kernel
  int pair = -1;        // match.pair  1 / 0: force / forbid the paired-row kernel
  int rt = -1;          // match.rt    1 / 2: row tiles per CTA of match_kernel
  int ctas = -1;        // match.ctas  grid of the persistent kernels (default: one CTA per SM)
  int cta2 = -1;        // match.cta2  1: CTA pairs (cta_group::2) in the paired-row kernel (default: single CTAs)
  int alt_cta2 = -1;    // match.alt_cta2  0: single CTAs in the alternating ARGMAX kernel (default: CTA pairs)
};
MatchConfig g_cfg;

template <typename K>
int set_smem_limit(K kernel) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  return e == cudaSuccess ? GADM_OK : set_cuda_error(e);
}

inline size_t ws_counter_bytes() { return 4096; }
inline size_t ws_partial_bytes(int slots) { return size_t(slots) * 2 * PART_ROWS * 8 * sizeof(float); }

}  // namespace

int match_config_set(const char* key, int value) {
  if (!strcmp(key, "match.alt")) { g_cfg.alt = value; return GADM_OK; }
  if (!strcmp(key, "match.pair")) { g_cfg.pair = value; return GADM_OK; }
  if (!strcmp(key, "match.rt")) { g_cfg.rt = value; return GADM_OK; }
  if (!strcmp(key, "match.ctas")) { g_cfg.ctas = value; return GADM_OK; }
  if (!strcmp(key, "match.cta2")) { g_cfg.cta2 = value; return GADM_OK; }
  if (!strcmp(key, "match.alt_cta2")) { g_cfg.alt_cta2 = value; return GADM_OK; }
  return GADM_ERR_BAD_ARG;
}

int match_configure(int device) {
  if (device < 0 || device >= MAX_DEVICES) return GADM_ERR_UNSUPPORTED;
  int rc;
  if ((rc = set_smem_limit(match_kernel<false, 1>)) != GADM_OK) return rc;
  if ((rc = set_smem_limit(match_kernel<true, 1>)) != GADM_OK) return rc;
  if ((rc = set_smem_limit(match_kernel<false, 2>)) != GADM_OK) return rc;
  if ((rc = set_smem_limit(match_kernel<true, 2>)) != GADM_OK) return rc;
  if ((rc = set_smem_limit(match_pair_kernel<false, false>)) != GADM_OK) return rc;
  if ((rc = set_smem_limit(match_pair_kernel<true, false>)) != GADM_OK) return rc;
  if ((rc = set_smem_limit(match_pair_kernel<false, true>)) != GADM_OK) return rc;
  if ((rc = set_smem_limit(match_pair_kernel<true, true>)) != GADM_OK) return rc;
  if ((rc = set_smem_limit(match_alt_kernel<false, false, false>)) != GADM_OK) return rc;
  if ((rc = set_smem_limit(match_alt_kernel<true, false, false>)) != GADM_OK) return rc;
  if ((rc = set_smem_limit(match_alt_kernel<false, true, false>)) != GADM_OK) return rc;
  if ((rc = set_smem_limit(match_alt_kernel<false, false, true>)) != GADM_OK) return rc;
  if ((rc = set_smem_limit(match_alt_kernel<true, false, true>)) != GADM_OK) return rc;
  if ((rc = set_smem_limit(match_alt_kernel<false, true, true>)) != GADM_OK) return rc;
  int sms = 0;
  cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  if (e != cudaSuccess) return set_cuda_error(e);
  // the workspace slots are indexed by %smid, whose range is [0, %nsmid): read it once (init may synchronise;
  // the result lands in a __device__ variable, nothing is allocated)
  unsigned int nsmid = 0;
  nsmid_kernel<<<1, 1>>>();
  e = cudaMemcpyFromSymbol(&nsmid, g_nsmid_out, sizeof(unsigned int));
  if (e != cudaSuccess) return set_cuda_error(e);
  g_dev[device].sms = sms;
  g_dev[device].slots = sms > int(nsmid) ? sms : int(nsmid);
  return GADM_OK;
}

// Workspace of gadm_match_fwd on the current device: one argmax stash slot per SM, the arrival counters and the
// partial results of the persistent kernels (both bounded by the SM count: at most two partial row blocks per CTA).
size_t match_workspace_bytes() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) return 0;
  const int slots = g_dev[dev].slots;
  if (slots <= 0) return 0;
  return size_t(slots) * STASH_SLOT_BYTES + ws_counter_bytes() + ws_partial_bytes(slots);
}

template <bool kSoft>
static int match_launch_t(const void* rows, const void* cols, MatchParams p, int Kp, int sms, cudaStream_t stream) {
  const int KB = Kp / BK;
  const MatchConfig cfg = g_cfg;
  if (!kSoft) {
    // alternating persistent kernel (ARGMAX; needs the workspace, a ring of two resident model tiles and more than
    // one row tile per frame)
    const int astages = match_alt_stages(KB);
    if (cfg.alt != 0 && p.stash != nullptr && astages >= 2 * KB && p.N > BM &&
        (p.n_rows == nullptr || p.B <= SCHED_MAX_FRAMES)) {
      p.KB = KB; p.stages = astages;
      p.T = (p.M + BN - 1) / BN;
      p.RB = (p.N + PART_ROWS - 1) / PART_ROWS;
      p.total_units = (long long)p.B * p.RB * p.T;
      int grid = cfg.ctas > 0 ? min(cfg.ctas, sms) : sms;
      if ((long long)grid > p.total_units) grid = int(p.total_units);
      CUtensorMap tmap_rows, tmap_cols;
      int rc = make_tmap_2b_3d(&tmap_rows, rows, uint64_t(Kp), uint64_t(p.N), uint64_t(p.B), BK, BM, 0);
      if (rc != GADM_OK) return rc;
      rc = make_tmap_2b_3d(&tmap_cols, cols, uint64_t(Kp), uint64_t(p.M), uint64_t(p.n_obj), BK, BN, 0);
      if (rc != GADM_OK) return rc;
      // CTA pairs by default: bit-identical results, 2-4 % faster at the BASELINE shape in every ARGMAX flavour, alone
      // and under the kNN pyramid of a second stream (tools/bench_match.py, tools/bench_overlap.py)
      const bool cta2 = cfg.alt_cta2 != 0 && grid >= 2 && p.N > PART_ROWS;
      if (cta2) {
        // CTA pairs: units are (pair of row blocks, model tile), dealt out to grid / 2 clusters
        p.stages = match_alt_stages(KB, true);
        p.RB = (p.N + 2 * PART_ROWS - 1) / (2 * PART_ROWS);
        p.total_units = (long long)p.B * p.RB * p.T;
        int pairs = (cfg.ctas > 0 ? min(cfg.ctas, sms) : sms) / 2;
        if ((long long)pairs > p.total_units) pairs = int(p.total_units);
        if (pairs < 1) pairs = 1;
        grid = 2 * pairs;
        rc = make_tmap_2b_3d(&tmap_cols, cols, uint64_t(Kp), uint64_t(p.M), uint64_t(p.n_obj), BK, BN / 2, 0);
        if (rc != GADM_OK) return rc;
      }
      cudaError_t e = cudaMemsetAsync(p.seg_count, 0, size_t(grid) * sizeof(unsigned int), stream);
      if (e != cudaSuccess) return set_cuda_error(e);
      if (cta2) {
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3(grid); lc.blockDim = dim3(NUM_THREADS);
        lc.dynamicSmemBytes = match_alt_smem_bytes(KB, p.stages, true);
        lc.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        lc.attrs = at; lc.numAttrs = 1;
        if (p.unit_scales == 1) e = cudaLaunchKernelEx(&lc, match_alt_kernel<true, false, true>, tmap_rows, tmap_cols, p);
        else if (p.unit_scales == 2) e = cudaLaunchKernelEx(&lc, match_alt_kernel<false, true, true>, tmap_rows, tmap_cols, p);
        else e = cudaLaunchKernelEx(&lc, match_alt_kernel<false, false, true>, tmap_rows, tmap_cols, p);
        if (e != cudaSuccess) return set_cuda_error(e);
        return check_launch();
      }
      const size_t smem = match_alt_smem_bytes(KB, astages);
      if (p.unit_scales == 1)
        match_alt_kernel<true, false, false><<<grid, NUM_THREADS, smem, stream>>>(tmap_rows, tmap_cols, p);
      else if (p.unit_scales == 2)
        match_alt_kernel<false, true, false><<<grid, NUM_THREADS, smem, stream>>>(tmap_rows, tmap_cols, p);
      else
        match_alt_kernel<false, false, false><<<grid, NUM_THREADS, smem, stream>>>(tmap_rows, tmap_cols, p);
      return check_launch();
    }
  }
  if (p.B > 65535) return GADM_ERR_UNSUPPORTED;     // the kernels below put the frame in blockIdx.y
  {
    // paired-row kernel (256 rows per CTA, every epilogue thread owns two rows).  Measured at the BASELINE shape:
    // SOFT 0.384 ms against 0.400 ms (RT = 1), ARGMAX 0.241 ms against 0.211 ms (RT = 2) => default for SOFT only.
    const bool pair = cfg.pair < 0 ? kSoft : cfg.pair != 0;
    const int pstages = match_pair_stages<kSoft>(KB);
    if (pair && pstages >= 2 * KB && KB <= P_MAX_KB && p.N > BM) {
      p.KB = KB; p.stages = pstages;
      CUtensorMap tmap_rows, tmap_cols;
      int rc = make_tmap_2b_3d(&tmap_rows, rows, uint64_t(Kp), uint64_t(p.N), uint64_t(p.B), BK, BM, 0);
      if (rc != GADM_OK) return rc;
      rc = make_tmap_2b_3d(&tmap_cols, cols, uint64_t(Kp), uint64_t(p.M), uint64_t(p.n_obj), BK, PBN, 0);
      if (rc != GADM_OK) return rc;
      dim3 grid((p.N + 2 * BM - 1) / (2 * BM), p.B);
      if (cfg.cta2 == 1) {
        // CTA pairs (opt-in): clusters of two row blocks of a frame share every model tile (half each).  Measured at
        // the BASELINE shape: shared-memory operand wavefronts -25 %, L2 -> SM traffic -35 %, time unchanged (0.357
        // against 0.353 ms) -- the epilogue's instruction streams bound this kernel, not its operand traffic.
        const int cstages = match_pair_stages<kSoft>(KB, true);
        p.stages = cstages;
        rc = make_tmap_2b_3d(&tmap_cols, cols, uint64_t(Kp), uint64_t(p.M), uint64_t(p.n_obj), BK, PBN / 2, 0);
        if (rc != GADM_OK) return rc;
        grid.x = (grid.x + 1) & ~1u;
        cudaLaunchConfig_t lc = {};
        lc.gridDim = grid; lc.blockDim = dim3(NUM_THREADS);
        lc.dynamicSmemBytes = match_pair_smem_bytes<kSoft>(KB, cstages, true);
        lc.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        lc.attrs = at; lc.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&lc, match_pair_kernel<kSoft, true>, tmap_rows, tmap_cols, p);
        if (e != cudaSuccess) return set_cuda_error(e);
        return check_launch();
      }
      match_pair_kernel<kSoft, false><<<grid, NUM_THREADS, match_pair_smem_bytes<kSoft>(KB, pstages), stream>>>(
          tmap_rows, tmap_cols, p);
      return check_launch();
    }
  }
  // Two row tiles per CTA when a ring of at least 2 KB stages (one tile resident, one in flight) fits beside them
  // and the frame has more than one row tile; otherwise one row tile with the deepest ring.  Measured at the
  // BASELINE shape: ARGMAX 0.204 ms (RT = 2) against 0.222 ms; SOFT 0.43 ms (RT = 2) against 0.41 ms -- SOFT is
  // bound by the epilogue's shared-memory traffic, not by operand traffic, and prefers 16 warps per accumulator.
  int RT = cfg.rt == 1 ? 1 : cfg.rt == 2 ? 2 : kSoft ? 1 : 2;
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
                 void* workspace, size_t workspace_bytes, const int32_t* n_rows, const int32_t* row_map, int N_out,
                 cudaStream_t stream) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return set_cuda_error(e);
  if (dev < 0 || dev >= MAX_DEVICES || g_dev[dev].slots <= 0) return GADM_ERR_NOT_INIT;   // gadm_init(dev) not run
  const int slots = g_dev[dev].slots;
  MatchParams p;
  const bool ws_ok = workspace != nullptr && workspace_bytes >= match_workspace_bytes();
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  p.stash = ws_ok ? ws : nullptr;
  p.stash_slots = slots;
  p.seg_count = ws_ok ? reinterpret_cast<unsigned int*>(ws + size_t(slots) * STASH_SLOT_BYTES) : nullptr;
  p.partial = ws_ok ? reinterpret_cast<float*>(ws + size_t(slots) * STASH_SLOT_BYTES + ws_counter_bytes()) : nullptr;
  p.T = 0; p.RB = 0; p.total_units = 0;
  p.n_rows = n_rows; p.row_map = row_map; p.N_out = N_out;
  p.rinv_rows = rinv_rows; p.pad_sim = pad_sim; p.scales = aux_scales(aux, n_obj, M);
  p.planes = aux_planes(aux, n_obj, M); p.mask = mask; p.obj_id = obj_id;
  p.idx = idx; p.max_sim = max_sim; p.weight = weight; p.soft_xyz = soft_xyz;
  p.B = B; p.N = N; p.M = M; p.KB = 0; p.n_obj = n_obj; p.stages = 0; p.pad_mode = pad_mode;
  p.gamma_log2e = gamma * 1.4426950408889634f;
  p.unit_scales = mode == GADM_MATCH_ARGMAX_UNIT ? 1 : mode == GADM_MATCH_ARGMAX_BF16N ? 2 : 0;
  if (mode == GADM_MATCH_SOFT) return match_launch_t<true>(rows, cols, p, Kp, g_dev[dev].sms, stream);
  return match_launch_t<false>(rows, cols, p, Kp, g_dev[dev].sms, stream);
}

}  // namespace gadm
