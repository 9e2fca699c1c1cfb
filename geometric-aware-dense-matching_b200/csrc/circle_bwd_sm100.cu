// CircleLoss backward with the scene-side gradient product fused into the kernel:
//   dL/df^_i = sum_j dL/dsim_ij m^_j
// is accumulated in tensor memory by a second MMA per model tile, fed from shared memory with the tile of dL/dsim the
// epilogue has just computed -- the score tile is recomputed (as in circle_kernel<grad>), nothing of size [N, M] is read
// back for this product.  Without kDm dL/dsim still leaves the SM once (the split bf16 form of
// gadm_circle_loss_bwd_split) for the model-side product G^T F^ as a library GEMM; with kDm that product is formed here
// as well (see the kernel) and dL/dsim never leaves the SM.
//
// Per CTA: one row tile (128 scene rows), model tiles of 128 vertices.
//   S   [128 x 128]  = F (rows, K-major over d)  x  M_t (K-major over d)          two accumulators, alternating
//   G'' [128 x 128]  = dL/dsim * (1/|f_i|) * (1/|m_j|) as hi + lo bf16, written by the epilogue into shared memory in the
//                      K-major SWIZZLE_128B layout of an A operand (K = the tile's 128 vertices: two blocks of 64)
//   dF  [128 x d]   += G''_hi x M_t + G''_lo x M_t, M_t read AGAIN from the stage it already occupies, now as an
//                      MN-major B operand (N = d contiguous, K = vertex rows: LBO = the 16 KB block stride between the
//                      64-wide d blocks, SBO = 1024 B between groups of 8 vertex rows)
// The stage of a model tile is released by the commit of ITS dF MMAs, the G'' buffer by the same commit; the epilogue of
// tile t + 1 reaches its first G'' store about when the 16 dF MMAs of tile t have drained, so one buffer suffices.
// TMEM: S 2 x 128 columns, dF d <= 128 columns, kDm: the tile's dM 128 more.  d <= 128 (K' <= 128) only; other shapes
// use the library path.
#include "match_common.cuh"

namespace gadm {

namespace {

constexpr int DBN = 128;                      // vertices per model tile
constexpr int DSL = 4, DCS = DBN / DSL;       // 32-column slices
constexpr int D_BLK_BYTES = DBN * BK * 2;     // 16 KB: 128 vertices x 64 d
constexpr int D_PLANE = DBN * 4;
constexpr int D_AUX_BYTES = 5 * D_PLANE;
constexpr int D_MAX_KB = 2;
constexpr int D_STAGES = 3;                   // whole model tiles
constexpr int G_BLK_BYTES = BM * 64 * 2;      // 16 KB: 128 rows x 64 vertices of one part

struct DfBarriers {
  uint64_t full[D_STAGES], empty[D_STAGES];
  uint64_t a_full;
  uint64_t s_full[2], s_free[2];
  uint64_t aux_full[AUX_SLOTS], aux_empty[AUX_SLOTS];
  uint64_t g_full, g_free, df_full;
  uint64_t dm_full, dm_free;      // kDm: the model-side tile product is complete / has been drained
  uint32_t tmem_base, pad;
};

struct DfParams {
  const float* rinv_rows;
  const float* pad_sim;
  const float* scales;
  const float* planes;
  const float* xyz;
  const int64_t* match_idx;
  const int64_t* match_idx2;
  const int32_t* obj_id;
  const float* lse_p;
  const float* lse_n;
  const float* w;
  float* G;        // [B, N, Mp] words: split bf16 form
  float* g_pad;    // [B, N]
  float* dF;       // [B, N, Kp] fp32: rinv_i * dL/df^_i without the pad column's term
  float* dM;       // kDm: [B, Mp, Kp] fp32, zeroed by the caller: scale_j * dL/dm^_j of every frame (fp32 reductions)
  int Mp, B, N, M, KB, n_obj;
  float gamma_log2e, margin;
};

// kDm: the model-side product too.  After dF += G'' M_t the issuer computes the tile's own
//   dM_t [128 vertices x d] = G''^T [vertices x rows] x F [rows x d]
// with BOTH operands MN-major -- the G'' buffer read transposed (M = vertices contiguous, K = rows: LBO = the 16 KB stride
// between its two 64-vertex blocks), the row tile read with N = d contiguous -- into a fourth 128-column accumulator;
// the epilogue drains it one tile later with 16-byte fp32 reductions into the frame's dM (every row tile of the frame
// adds to the same [M, d] array).  dL/dsim then never leaves the SM: G is not written.
template <bool kExact, bool kDm>
__global__ void __launch_bounds__(NUM_THREADS, 1)
circle_df_kernel(const __grid_constant__ CUtensorMap tmap_rows, const __grid_constant__ CUtensorMap tmap_cols,
                 const DfParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                                        // [KB] blocks of 128 rows x 64 d
  uint8_t* smem_b = smem_a + p.KB * A_BLK_BYTES;                 // [D_STAGES][KB] blocks of 128 vertices x 64 d
  uint8_t* smem_g = smem_b + D_STAGES * p.KB * D_BLK_BYTES;      // [part 2][vertex block 2] 128 rows x 64 vertices
  uint8_t* smem_aux = smem_g + 4 * G_BLK_BYTES;
  DfBarriers* bars = reinterpret_cast<DfBarriers*>(smem_aux + AUX_SLOTS * D_AUX_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int row0 = blockIdx.x * BM;
  const int obj = p.obj_id ? min(max(p.obj_id[b], 0), p.n_obj - 1) : (p.n_obj == p.B ? b : 0);
  const int num_tiles = (p.M + DBN - 1) / DBN;
  const int TILE_BYTES = p.KB * D_BLK_BYTES;

  if (warp == EPI_WARPS && lane == 0) {
    ptx::prefetch_tensormap(&tmap_rows);
    ptx::prefetch_tensormap(&tmap_cols);
    for (int s = 0; s < D_STAGES; ++s) {
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
    ptx::mbar_init(&bars->g_full, EPI_WARPS);
    ptx::mbar_init(&bars->g_free, 1);
    ptx::mbar_init(&bars->df_full, 1);
    ptx::mbar_init(&bars->dm_full, 1);
    ptx::mbar_init(&bars->dm_free, EPI_WARPS);
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
      const size_t plane = size_t(p.B) * p.M;
      const float* sc_tab = p.scales + size_t(obj) * p.M;
      const float* xyz_tab = p.planes + size_t(b) * p.M;
      for (int t = 0; t < num_tiles; ++t) {
        const int slot = t % AUX_SLOTS;
        const uint32_t use = uint32_t(t) / AUX_SLOTS;
        const uint32_t bytes = uint32_t(min(DBN, p.M - t * DBN)) * 4;
        ptx::mbar_wait_sleep(&bars->aux_empty[slot], (use & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&bars->aux_full[slot], 5 * bytes);
        uint8_t* aux = smem_aux + slot * D_AUX_BYTES;
        ptx::bulk_load_1d(aux, sc_tab + size_t(t) * DBN, bytes, &bars->aux_full[slot]);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          ptx::bulk_load_1d(aux + (c + 1) * D_PLANE, xyz_tab + c * plane + size_t(t) * DBN, bytes, &bars->aux_full[slot]);
        const int stage = t % D_STAGES;
        ptx::mbar_wait_sleep(&bars->empty[stage], ((uint32_t(t) / D_STAGES) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&bars->full[stage], TILE_BYTES);
        for (int kb = 0; kb < p.KB; ++kb)     // vertices >= M are zero-filled by TMA
          ptx::tma_load_3d(smem_b + stage * TILE_BYTES + kb * D_BLK_BYTES, &tmap_cols, &bars->full[stage], kb * BK,
                           t * DBN, obj);
      }
    }
  } else if (warp == EPI_WARPS + 1) {
    // ============================== UMMA issuer ==============================
    if (lane == 0) {
      constexpr uint32_t idesc_s = ptx::umma_idesc_bf16_f32(BM, DBN);
      const uint32_t idesc_df = ptx::umma_idesc_bf16_f32(BM, uint32_t(p.KB * BK)) | (1u << 16);   // B operand MN-major
      constexpr uint64_t DESC_HI = uint64_t((1024u >> 4) | (1u << 14) | (2u << 29)) << 32;         // SBO | version | SW128
      if (tmem_base != 0) __trap();
      const uint32_t a_lo0 = ((ptx::smem_u32(smem_a) & 0x3FFFF) >> 4) | 0x10000u;
      const uint32_t b_lo0 = ((ptx::smem_u32(smem_b) & 0x3FFFF) >> 4);
      const uint32_t g_lo0 = ((ptx::smem_u32(smem_g) & 0x3FFFF) >> 4) | 0x10000u;
      constexpr uint32_t B_LBO_K = 0x10000u;                                  // K-major: LBO unused (1)
      constexpr uint32_t B_LBO_MN = uint32_t(D_BLK_BYTES >> 4) << 16;          // MN-major: next 64-wide d block
      constexpr uint32_t D_DF = 2 * DBN;                                       // TMEM column of the dF accumulator
      constexpr uint32_t D_DM = 3 * DBN;                                       // ... of the tile's dM accumulator
      const uint32_t idesc_dm = idesc_df | (1u << 15);                         // A operand MN-major as well
      const uint32_t g_lo_mn0 = (ptx::smem_u32(smem_g) & 0x3FFFF) >> 4;
      const uint32_t a_lo_mn0 = (ptx::smem_u32(smem_a) & 0x3FFFF) >> 4;
      auto issue_df = [&](int u) {
        ptx::mbar_wait_sleep(&bars->g_full, uint32_t(u) & 1);
        ptx::tc_fence_after();
        const uint32_t b_tile = b_lo0 + uint32_t(u % D_STAGES) * uint32_t(TILE_BYTES >> 4);
#pragma unroll
        for (int part = 0; part < 2; ++part)
#pragma unroll
          for (int vb = 0; vb < 2; ++vb)
#pragma unroll
            for (int k = 0; k < 64 / UMMA_K; ++k) {
              const uint32_t a_lo = g_lo0 + uint32_t(part * 2 + vb) * (G_BLK_BYTES >> 4) + k * (UMMA_K * 2 >> 4);
              // 16 vertex rows of 128 bytes = 2048 bytes per K step
              const uint32_t b_lo = (b_tile + uint32_t((vb * 64 + k * UMMA_K) * 128 >> 4)) | B_LBO_MN;
              ptx::umma_bf16_ss(D_DF, DESC_HI | a_lo, DESC_HI | b_lo, idesc_df, (u | part | vb | k) != 0);
            }
        if (kDm) {
          if (u >= 1) ptx::mbar_wait_sleep(&bars->dm_free, uint32_t(u - 1) & 1);   // the previous tile's product is drained
          ptx::tc_fence_after();
#pragma unroll
          for (int part = 0; part < 2; ++part)
#pragma unroll
            for (int k = 0; k < BM / UMMA_K; ++k) {     // K = the 128 rows of the tile, 16 (2048 bytes) per step
              const uint32_t a_lo = (g_lo_mn0 + uint32_t(part * 2) * (G_BLK_BYTES >> 4) + uint32_t(k * UMMA_K * 128 >> 4)) |
                                    (uint32_t(G_BLK_BYTES >> 4) << 16);
              const uint32_t b_lo = (a_lo_mn0 + uint32_t(k * UMMA_K * 128 >> 4)) | (uint32_t(A_BLK_BYTES >> 4) << 16);
              ptx::umma_bf16_ss(D_DM, DESC_HI | a_lo, DESC_HI | b_lo, idesc_dm, (part | k) != 0);
            }
          ptx::umma_commit(&bars->dm_full);
        }
        ptx::umma_commit(&bars->g_free);
        ptx::umma_commit(&bars->empty[u % D_STAGES]);
      };
      ptx::mbar_wait(&bars->a_full, 0);
      for (int t = 0; t < num_tiles; ++t) {
        const int acc = t & 1;
        const int stage = t % D_STAGES;
        ptx::mbar_wait_sleep(&bars->s_free[acc], ((uint32_t(t) >> 1) & 1) ^ 1);
        ptx::mbar_wait_sleep(&bars->full[stage], (uint32_t(t) / D_STAGES) & 1);
        ptx::tc_fence_after();
        for (int kb = 0; kb < p.KB; ++kb) {
          const uint32_t a_lo = a_lo0 + uint32_t(kb) * (A_BLK_BYTES >> 4);
          const uint32_t b_lo = (b_lo0 + uint32_t(stage) * uint32_t(TILE_BYTES >> 4) + uint32_t(kb) * (D_BLK_BYTES >> 4)) |
                                B_LBO_K;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            ptx::umma_bf16_ss(uint32_t(acc) * DBN, DESC_HI | (a_lo + k * (UMMA_K * 2 >> 4)),
                              DESC_HI | (b_lo + k * (UMMA_K * 2 >> 4)), idesc_s, (kb | k) != 0);
        }
        ptx::umma_commit(&bars->s_full[acc]);
        if (t >= 1) issue_df(t - 1);
      }
      issue_df(num_tiles - 1);
      ptx::umma_commit(&bars->df_full);
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
    const int c1 = int(mi), c2 = kExact && row_ok ? int(p.match_idx2[grow]) : -1;
    const bool in_mesh = kExact ? !(c1 == p.M || c2 == p.M) : (mi >= 0 && mi < p.M);
    float gx = -1e18f, gy = -1e18f, gz = -1e18f;
    if (in_mesh) {
      const float* e = p.xyz + (size_t(obj) * p.M + size_t(mi)) * 3;
      gx = e[0]; gy = e[1]; gz = e[2];
    }
    const float m = p.margin, one_m = 1.f - p.margin, one_p = 1.f + p.margin, gl = p.gamma_log2e;
    const uint32_t lane_base = tmem_base + (uint32_t(q * 32) << 16);
    const float Lp = row_ok ? p.lse_p[grow] * 1.4426950408889634f : 0.f;
    const float Ln = row_ok ? p.lse_n[grow] * 1.4426950408889634f : 0.f;
    const float wg = row_ok ? p.w[grow] * (gl * 0.6931471805599453f) : 0.f;     // w_i * gamma
    const float wgs = wg * rs;
    float* grow_g = kDm ? nullptr : p.G + grow * size_t(p.Mp);
    // G'' in shared memory: block (part, sub >> 1), row row_in_tile, 16-byte chunk ((sub & 1) * 4 + c * 2 + {0, 1}) ^ (row & 7)
    const uint32_t g_row = ptx::smem_u32(smem_g) + uint32_t(sub >> 1) * G_BLK_BYTES + uint32_t(row_in_tile) * 128;
    const uint32_t sw = uint32_t(row_in_tile & 7);

    // kDm: tile u's model-side product: this thread holds vertex u * 128 + row_in_tile, d columns [sub * dcm, + dcm)
    const int dcm = p.KB * BK / DSL;
    auto drain_dm = [&](int u) {
      ptx::mbar_wait_sleep(&bars->dm_full, uint32_t(u) & 1);
      ptx::tc_fence_after();
      const int vtx = u * DBN + row_in_tile;
      float* dst = p.dM + (size_t(b) * p.Mp + size_t(min(vtx, p.Mp - 1))) * size_t(p.KB * BK) + sub * dcm;
      for (int c0 = 0; c0 < dcm; c0 += 16) {
        uint32_t d[16];
        ptx::tmem_ld_32x16(lane_base + 3 * DBN + sub * dcm + c0, d);
        ptx::tmem_ld_wait();
        if (vtx < p.M) {
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4)
            ptx::red_add_v4(dst + c0 + j4 * 4, __uint_as_float(d[j4 * 4]), __uint_as_float(d[j4 * 4 + 1]),
                            __uint_as_float(d[j4 * 4 + 2]), __uint_as_float(d[j4 * 4 + 3]));
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bars->dm_free);
    };

    for (int t = 0; t < num_tiles; ++t) {
      const int acc = t & 1;
      const uint32_t use = uint32_t(t) >> 1;
      const int slot = t % AUX_SLOTS;
      ptx::mbar_wait_sleep(&bars->aux_full[slot], (uint32_t(t) / AUX_SLOTS) & 1);
      ptx::mbar_wait_sleep(&bars->s_full[acc], use & 1);
      ptx::tc_fence_after();
      const int ncols = min(DBN, p.M - t * DBN) - sub * DCS;   // valid columns of this slice (may be <= 0)
      const uint32_t s_tmem = lane_base + acc * DBN + sub * DCS;
      const uint32_t sc_addr = ptx::smem_u32(smem_aux + slot * D_AUX_BYTES) + sub * DCS * 4;
#pragma unroll 1
      for (int c = 0; c < DCS / 16; ++c) {
        float gout[16];
        if (ncols - c * 16 > 0) {
          uint32_t d[16];
          ptx::tmem_ld_32x16(s_tmem + c * 16, d);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const uint32_t a = sc_addr + (c * 16 + j4 * 4) * 4;
            const float4 cm = ptx::lds128(a);
            float xs[4], ys[4], zs[4], r2s[4];
            if (!kExact) {
              const float4 X = ptx::lds128(a + D_PLANE), Y = ptx::lds128(a + 2 * D_PLANE),
                           Z = ptx::lds128(a + 3 * D_PLANE), R = ptx::lds128(a + 4 * D_PLANE);
              xs[0] = X.x; xs[1] = X.y; xs[2] = X.z; xs[3] = X.w;
              ys[0] = Y.x; ys[1] = Y.y; ys[2] = Y.z; ys[3] = Y.w;
              zs[0] = Z.x; zs[1] = Z.y; zs[2] = Z.z; zs[3] = Z.w;
              r2s[0] = R.x; r2s[1] = R.y; r2s[2] = R.z; r2s[3] = R.w;
            }
            const float cs[4] = {cm.x, cm.y, cm.z, cm.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float s = (__uint_as_float(d[j4 * 4 + e]) * cs[e]) * rs;             // cosine similarity
              bool pos;
              if (kExact) {
                const int col = t * DBN + sub * DCS + c * 16 + j4 * 4 + e;
                pos = col == c1 || col == c2;
              } else {
                const float dx = __fsub_rn(gx, xs[e]), dy = __fsub_rn(gy, ys[e]), dz = __fsub_rn(gz, zs[e]);
                const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                pos = __fadd_rn(d2, 1e-7f) < r2s[e];
              }
              const float a_ = fmaxf(pos ? one_p - s : s + m, 0.f);
              const float lg = (a_ * (s - (pos ? one_m : m))) * (pos ? -gl : gl);
              const float sm = ptx::ex2_approx(lg - (pos ? Lp : Ln));
              const bool valid = c * 16 + j4 * 4 + e < ncols;      // stale scales behind column M are meaningless
              gout[j4 * 4 + e] = valid ? ((wgs * sm) * (pos ? -a_ : a_)) * cs[e] : 0.f;
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) gout[j] = 0.f;
        }
        float hw[8], lw[8];                    // words of bf16 pairs: hi parts, lo parts
#pragma unroll
        for (int e2 = 0; e2 < 8; ++e2) {
          const float v0 = gout[e2 * 2], v1 = gout[e2 * 2 + 1];
          const uint32_t h = ptx::cvt_bf16x2(v1, v0);
          const float r0 = v0 - __uint_as_float(h << 16), r1 = v1 - __uint_as_float(h & 0xffff0000u);
          hw[e2] = __uint_as_float(h);
          lw[e2] = __uint_as_float(ptx::cvt_bf16x2(r1, r0));
        }
        if (!kDm && row_ok) {
          float* dst = grow_g + t * DBN + sub * DCS + c * 16;
#pragma unroll
          for (int j8 = 0; j8 < 2; ++j8)
            if (c * 16 + j8 * 8 < ncols)
              ptx::stg256(dst + j8 * 8, hw[j8 * 4], hw[j8 * 4 + 1], hw[j8 * 4 + 2], hw[j8 * 4 + 3], lw[j8 * 4],
                          lw[j8 * 4 + 1], lw[j8 * 4 + 2], lw[j8 * 4 + 3]);
        }
        // the dF MMAs of the previous tile have read the buffer (first store of this tile only)
        if (c == 0 && t >= 1) ptx::mbar_wait_sleep(&bars->g_free, uint32_t(t - 1) & 1);
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const uint32_t chunk = (uint32_t((sub & 1) * 4 + c * 2 + h2) ^ sw) * 16;
          ptx::sts128(g_row + chunk, hw[h2 * 4], hw[h2 * 4 + 1], hw[h2 * 4 + 2], hw[h2 * 4 + 3]);
          ptx::sts128(g_row + 2 * G_BLK_BYTES + chunk, lw[h2 * 4], lw[h2 * 4 + 1], lw[h2 * 4 + 2], lw[h2 * 4 + 3]);
        }
      }
      ptx::fence_proxy_async();          // the G'' stores must be visible to the tensor core's (async proxy) reads
      if (kDm && t >= 1) drain_dm(t - 1);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(&bars->g_full);
        ptx::mbar_arrive(&bars->s_free[acc]);
        ptx::mbar_arrive(&bars->aux_empty[slot]);
      }
    }

    if (sub == 0 && row_ok) {       // the pad column and the zero padding of the row
      const float s = p.pad_sim[grow];
      const float ap = fmaxf(one_p - s, 0.f), an = fmaxf(s + m, 0.f);
      const float gp = in_mesh ? wg * ptx::ex2_approx(an * (s - m) * gl - Ln) * an
                               : wg * ptx::ex2_approx(-ap * (s - one_m) * gl - Lp) * -ap;
      p.g_pad[grow] = gp;
      if (!kDm) for (int j = p.M; j < p.Mp; ++j) grow_g[j] = 0.f;
    }
    if (kDm) drain_dm(num_tiles - 1);
    // ---- dF: this thread's row, d columns [sub * (Kp / 4), +Kp / 4)
    ptx::mbar_wait_sleep(&bars->df_full, 0);
    ptx::tc_fence_after();
    const int Kp = p.KB * BK;
    const int dc = Kp / DSL;                       // 16 or 32 columns per thread
    float* out = p.dF + grow * size_t(Kp) + sub * dc;
    for (int c0 = 0; c0 < dc; c0 += 16) {
      uint32_t d[16];
      ptx::tmem_ld_32x16(lane_base + 2 * DBN + sub * dc + c0, d);
      ptx::tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int j8 = 0; j8 < 2; ++j8)
          ptx::stg256(out + c0 + j8 * 8, __uint_as_float(d[j8 * 8]), __uint_as_float(d[j8 * 8 + 1]),
                      __uint_as_float(d[j8 * 8 + 2]), __uint_as_float(d[j8 * 8 + 3]), __uint_as_float(d[j8 * 8 + 4]),
                      __uint_as_float(d[j8 * 8 + 5]), __uint_as_float(d[j8 * 8 + 6]), __uint_as_float(d[j8 * 8 + 7]));
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

inline size_t circle_df_smem_bytes(int KB) {
  return size_t(KB) * A_BLK_BYTES + size_t(D_STAGES) * KB * D_BLK_BYTES + 4 * G_BLK_BYTES + AUX_SLOTS * D_AUX_BYTES +
         sizeof(DfBarriers) + 1024;
}

}  // namespace

int circle_df_configure() {
  cudaError_t e = cudaFuncSetAttribute(circle_df_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return set_cuda_error(e);
  e = cudaFuncSetAttribute(circle_df_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return set_cuda_error(e);
  e = cudaFuncSetAttribute(circle_df_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return set_cuda_error(e);
  e = cudaFuncSetAttribute(circle_df_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return set_cuda_error(e);
  return GADM_OK;
}

bool circle_df_supported(int Kp) { return Kp % BK == 0 && Kp / BK >= 1 && Kp / BK <= D_MAX_KB; }

int circle_df_launch(const void* rows, const float* rinv_rows, const float* pad_sim, const void* cols, const float* aux,
                     const float* planes_frame, const int64_t* match_idx, const int64_t* match_idx2,
                     const int32_t* obj_id, int B, int N, int M, int Kp, int n_obj, float gamma, float margin,
                     const float* lse_p, const float* lse_n, const float* w, float* G, int Mp, float* g_pad, float* dF,
                     float* dM, cudaStream_t stream) {
  if (!circle_df_supported(Kp)) return GADM_ERR_UNSUPPORTED;
  DfParams p;
  p.rinv_rows = rinv_rows; p.pad_sim = pad_sim; p.scales = aux_scales(aux, n_obj, M); p.planes = planes_frame;
  p.xyz = aux_xyz(aux, n_obj, M); p.match_idx = match_idx; p.match_idx2 = match_idx2; p.obj_id = obj_id;
  p.lse_p = lse_p; p.lse_n = lse_n; p.w = w; p.G = G; p.g_pad = g_pad; p.dF = dF; p.dM = dM; p.Mp = Mp;
  p.B = B; p.N = N; p.M = M; p.KB = Kp / BK; p.n_obj = n_obj;
  p.gamma_log2e = gamma * 1.4426950408889634f; p.margin = margin;
  CUtensorMap tmap_rows, tmap_cols;
  int rc = make_tmap_2b_3d(&tmap_rows, rows, uint64_t(Kp), uint64_t(N), uint64_t(B), BK, BM, 0);
  if (rc != GADM_OK) return rc;
  rc = make_tmap_2b_3d(&tmap_cols, cols, uint64_t(Kp), uint64_t(M), uint64_t(n_obj), BK, DBN, 0);
  if (rc != GADM_OK) return rc;
  dim3 grid((N + BM - 1) / BM, B);
  const size_t smem = circle_df_smem_bytes(p.KB);
  if (dM != nullptr) {
    if (match_idx2 != nullptr) circle_df_kernel<true, true><<<grid, NUM_THREADS, smem, stream>>>(tmap_rows, tmap_cols, p);
    else circle_df_kernel<false, true><<<grid, NUM_THREADS, smem, stream>>>(tmap_rows, tmap_cols, p);
  } else {
    if (match_idx2 != nullptr) circle_df_kernel<true, false><<<grid, NUM_THREADS, smem, stream>>>(tmap_rows, tmap_cols, p);
    else circle_df_kernel<false, false><<<grid, NUM_THREADS, smem, stream>>>(tmap_rows, tmap_cols, p);
  }
  return check_launch();
}

}  // namespace gadm
