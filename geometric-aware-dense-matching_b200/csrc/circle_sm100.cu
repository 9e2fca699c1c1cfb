// Flash-style CircleLoss forward / dL/dsim on the tcgen05 similarity pipeline.
#include "match_common.cuh"

namespace gadm {

namespace {

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
  const int64_t* match_idx2;// [B, N] or null.  Non-null selects the EXACT-COLUMN positive set of matching_loss_sys
                            // (models/geoMatch.py:86-100): the positives of a row are the columns match_idx and
                            // match_idx2 themselves (M = the pad column), no radius test
  const uint8_t* fg;        // [B, N] rows that take part (labels == 1)
  const int32_t* obj_id;
  float* loss;              // [B, N] softplus(LSE_p + LSE_n), 0 for rows outside fg
  float* lse_p;             // [B, N] natural-log LSE of the positive / negative logits (for a backward pass)
  float* lse_n;
  const float* w;           // kGrad: [B, N] dL/dz of every row (0 for rows that take no part)
  float* G;                 // kGrad: [B, N, Mp] dL/dsim, column M = pad column, columns M+1.. = 0
                            // kGrad == 2: the same bytes hold bf16 pairs, see gadm_circle_loss_bwd_split
  float* g_pad;             // kGrad == 2: [B, N] dL/dsim of the pad column (fp32, unscaled)
  int Mp;
  int B, N, M, KB, n_obj, stages;
  float gamma_log2e, margin;
};

// kGrad: the same pass, but instead of the two sums every score's gradient is written,
//   dL/dsim_ij = w_i * (j positive ? softmax_p(j) * (-ap_ij gamma) : softmax_n(j) * (an_ij gamma)),
// with ap / an constants (the reference detaches them, loss.py:479-480) and the row's two LSEs from the forward pass.
// kGrad == 2 (split): G''_ij = G_ij * (1/|f_i|) * (1/|m_j|) as an unevaluated sum of two bf16 (hi + lo, 16 mantissa
// bits), every group of 8 columns stored as 8 hi then 8 lo -- the same 32 bytes the fp32 group takes.  With the norms
// folded into G'' both gradient GEMMs run on the EXACT bf16 operands the forward pass used, on the tensor cores, with
// K interleaved the same way (see gadm.h).
template <int kGrad, bool kExact>
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
  const int obj = p.obj_id ? min(max(p.obj_id[b], 0), p.n_obj - 1) : (p.n_obj == p.B ? b : 0);   // clamped
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
      // descriptors as 32-bit low words + one constant high word, accumulator address a function of acc alone (the
      // CTA owns all 512 TMEM columns: its allocation starts at column 0) -- see match_pair_kernel
      constexpr uint64_t DESC_HI = uint64_t((1024u >> 4) | (1u << 14) | (2u << 29)) << 32;
      if (tmem_base != 0) __trap();
      const uint32_t a_lo0 = ((ptx::smem_u32(smem_a) & 0x3FFFF) >> 4) | 0x10000u;
      const uint32_t b_lo0 = ((ptx::smem_u32(smem_b) & 0x3FFFF) >> 4) | 0x10000u;
      ptx::mbar_wait(&bars->a_full, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < num_tiles; ++t) {
        const int acc = t & 1;
        ptx::mbar_wait_sleep(&bars->s_free[acc], ((uint32_t(t) >> 1) & 1) ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = uint32_t(acc) * BN;
        for (int kb = 0; kb < p.KB; ++kb) {
          ptx::mbar_wait_sleep(&bars->full[stage], phase);
          ptx::tc_fence_after();
          const uint32_t a_lo = a_lo0 + uint32_t(kb) * (A_BLK_BYTES >> 4);
          const uint32_t b_lo = b_lo0 + uint32_t(stage) * (B_STAGE_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            ptx::umma_bf16_ss(d_tmem, DESC_HI | (a_lo + k * (UMMA_K * 2 >> 4)), DESC_HI | (b_lo + k * (UMMA_K * 2 >> 4)),
                              idesc, (kb | k) != 0);
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
    constexpr bool exact = kExact;      // exact-column positives (matching_loss_sys): p.match_idx2 is given
    const int c1 = int(mi), c2 = exact && row_ok ? int(p.match_idx2[grow]) : -1;
    // the pad column is a negative exactly for the rows that have a positive on the model (geoMatch.py:78); with the
    // exact-column set it is a positive iff one of the two columns IS the pad column
    const bool in_mesh = exact ? !(c1 == p.M || c2 == p.M) : (mi >= 0 && mi < p.M);
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
    constexpr bool kSplit = kGrad == 2;
    const float wg = kGrad && row_ok ? p.w[grow] * (gl * 0.6931471805599453f) : 0.f;     // w_i * gamma
    const float wgs = wg * rs;                                                            // kSplit: the row norm folded in
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
      // 16 columns of the slice.  kGuard (ragged last tile only) masks the columns >= ncols.  Only the branch of the
      // logit that the score's set needs is evaluated (the other one was computed and thrown away: 56 -> ~30
      // instructions per score); the operation order inside the branch is the reference's (loss.py:479-489).
      auto chunk = [&](int c, auto guard_tag) {
        constexpr bool kGuard = decltype(guard_tag)::value;
        uint32_t d[16];
        ptx::tmem_ld_32x16(s_tmem + c * 16, d);
        ptx::tmem_ld_wait();
        float gout[16];
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const uint32_t a = sc_addr + (c * 16 + j4 * 4) * 4;
          const float4 cm = ptx::lds128(a);
          float xs[4], ys[4], zs[4], r2s[4];
          if (!exact) {
            const float4 X = ptx::lds128(a + PLANE_BYTES), Y = ptx::lds128(a + 2 * PLANE_BYTES),
                         Z = ptx::lds128(a + 3 * PLANE_BYTES), R = ptx::lds128(a + 4 * PLANE_BYTES);
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
            if (exact) {
              const int col = t * BN + sub * CS + c * 16 + j4 * 4 + e;
              pos = col == c1 || col == c2;
            } else {
              // (A - B).pow(2).sum(-1) as the reference evaluates it: no FMA, left to right (basic_utils.py:88)
              const float dx = __fsub_rn(gx, xs[e]), dy = __fsub_rn(gy, ys[e]), dz = __fsub_rn(gz, zs[e]);
              const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
              pos = __fadd_rn(d2, 1e-7f) < r2s[e];                                     // sqrt(D2 + 1e-7) < positive_r[j]
            }
            // positive: ap = max(1 + m - s, 0), logit = (-ap (s - (1 - m))) gamma;  negative: an = max(s + m, 0),
            // logit = (an (s - m)) gamma   (loss.py:479-480, 488-489; log2 units)
            const float a_ = fmaxf(pos ? one_p - s : s + m, 0.f);
            const float lg = (a_ * (s - (pos ? one_m : m))) * (pos ? -gl : gl);
            const bool valid = !kGuard || c * 16 + j4 * 4 + e < ncols;
            if (kGrad) {
              const float sm = ptx::ex2_approx(lg - (pos ? Lp : Ln));                   // softmax weight inside its set
              gout[j4 * 4 + e] = kSplit ? ((wgs * sm) * (pos ? -a_ : a_)) * cs[e] : wg * sm * (pos ? -a_ : a_);
            } else {
              const float ex = ptx::ex2_approx(lg);
              if (valid && pos) sum_p += ex;
              if (valid && !pos) sum_n += ex;
            }
          }
        }
        if (kGrad && row_ok) {
          // 32-byte stores: a lane writes whole sectors of its row of G (the rows of a warp are 32 KB apart, so a
          // 16-byte store leaves every sector it touches half written)
          float* dst = grow_g + t * BN + sub * CS + c * 16;
#pragma unroll
          for (int j8 = 0; j8 < 2; ++j8)
            if (!kGuard || c * 16 + j8 * 8 < ncols) {  // M % 8 == 0: a group of 8 columns is valid or invalid as a whole
              if (kSplit) {
                float hw[4], lw[4];                    // 4 words of bf16 pairs each: hi parts, lo parts
#pragma unroll
                for (int e2 = 0; e2 < 4; ++e2) {
                  const float v0 = gout[j8 * 8 + e2 * 2], v1 = gout[j8 * 8 + e2 * 2 + 1];
                  const uint32_t h = ptx::cvt_bf16x2(v1, v0);                       // {hi16: bf16(v1), lo16: bf16(v0)}
                  const float r0 = v0 - __uint_as_float(h << 16), r1 = v1 - __uint_as_float(h & 0xffff0000u);
                  hw[e2] = __uint_as_float(h);
                  lw[e2] = __uint_as_float(ptx::cvt_bf16x2(r1, r0));
                }
                ptx::stg256(dst + j8 * 8, hw[0], hw[1], hw[2], hw[3], lw[0], lw[1], lw[2], lw[3]);
              } else {
                ptx::stg256(dst + j8 * 8, gout[j8 * 8], gout[j8 * 8 + 1], gout[j8 * 8 + 2], gout[j8 * 8 + 3],
                            gout[j8 * 8 + 4], gout[j8 * 8 + 5], gout[j8 * 8 + 6], gout[j8 * 8 + 7]);
              }
            }
        }
      };
      if (ncols >= CS) {
#pragma unroll 1
        for (int c = 0; c < CS / 16; ++c) chunk(c, std::integral_constant<bool, false>{});
      } else {
#pragma unroll 1
        for (int c = 0; c < CS / 16; ++c) {
          if (ncols - c * 16 <= 0) break;
          chunk(c, std::integral_constant<bool, true>{});
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
        if (kGrad == 2) {
          p.g_pad[grow] = gp;
          for (int j = p.M; j < p.Mp; ++j) grow_g[j] = 0.f;     // (bf16 zeros too)
        } else {
          for (int j = p.M; j < p.Mp; ++j) grow_g[j] = j == p.M ? gp : 0.f;
        }
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

}  // namespace

int circle_configure() {
  cudaError_t e = cudaFuncSetAttribute(circle_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return set_cuda_error(e);
  e = cudaFuncSetAttribute(circle_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return set_cuda_error(e);
  e = cudaFuncSetAttribute(circle_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return set_cuda_error(e);
  e = cudaFuncSetAttribute(circle_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return set_cuda_error(e);
  e = cudaFuncSetAttribute(circle_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return set_cuda_error(e);
  e = cudaFuncSetAttribute(circle_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return set_cuda_error(e);
  return GADM_OK;
}

int circle_launch(const void* rows, const float* rinv_rows, const float* pad_sim, const void* cols, const float* aux,
                  const float* planes_frame, const int64_t* match_idx, const int64_t* match_idx2, const uint8_t* fg,
                  const int32_t* obj_id, int B,
                  int N, int M, int Kp, int n_obj, float gamma, float margin, float* loss, float* lse_p,
                  float* lse_n, const float* w, float* G, int Mp, float* g_pad, cudaStream_t stream) {
  CircleParams p;
  p.w = w; p.G = G; p.Mp = Mp; p.g_pad = g_pad;
  p.rinv_rows = rinv_rows; p.pad_sim = pad_sim; p.scales = aux_scales(aux, n_obj, M); p.planes = planes_frame;
  p.xyz = aux_xyz(aux, n_obj, M); p.match_idx = match_idx; p.match_idx2 = match_idx2; p.fg = fg; p.obj_id = obj_id;
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
  const size_t smem = circle_smem_bytes(KB, stages);
  const bool exact = match_idx2 != nullptr;
  if (G != nullptr && g_pad != nullptr) {
    if (exact) circle_kernel<2, true><<<grid, NUM_THREADS, smem, stream>>>(tmap_rows, tmap_cols, p);
    else circle_kernel<2, false><<<grid, NUM_THREADS, smem, stream>>>(tmap_rows, tmap_cols, p);
  } else if (G != nullptr) {
    if (exact) circle_kernel<1, true><<<grid, NUM_THREADS, smem, stream>>>(tmap_rows, tmap_cols, p);
    else circle_kernel<1, false><<<grid, NUM_THREADS, smem, stream>>>(tmap_rows, tmap_cols, p);
  } else {
    if (exact) circle_kernel<0, true><<<grid, NUM_THREADS, smem, stream>>>(tmap_rows, tmap_cols, p);
    else circle_kernel<0, false><<<grid, NUM_THREADS, smem, stream>>>(tmap_rows, tmap_cols, p);
  }
  return check_launch();
}

}  // namespace gadm
