// Fused matching head for sm_100a: descriptor similarity (tcgen05 UMMA, bf16 in / fp32 accumulate in TMEM,
// operands staged by TMA) + row-wise argmax / online softmax / soft model coordinates in the epilogue.
// The [N, M] score matrix never leaves TMEM/registers.
//
// Replaces (reference tree): evaluator.py:89-93 (normalize, normalize, matmul, torch.max) and the padded
// variants utils/pvn3d_eval_utils_kpls.py:436-444 / models/geoMatch_DGCNN.py:92-99.  The softmax weight
// and soft coordinates are the extension defined in oracle/match_oracle.py.
//
// Work decomposition
//   CTA  = one 128-row tile of one frame (grid = ceil(N/128) x B), 18 warps:
//     warp 0      TMA producer: the 128 x K' row tile once, then model tiles (256 verts x 64 k, 32 KB) through
//                 an S-stage mbarrier ring, plus the per-tile aux table ({x,y,z,1/|m|} or 1/|m| only)
//     warp 1      UMMA issuer: 128x256x16 tcgen05.mma, accumulators double-buffered in TMEM (2 x 256 columns)
//     warps 2..17 epilogue: thread = row (four threads per row, one per 64-column slice of the tile);
//                 tcgen05.ld 16 columns at a time, software-pipelined; score = acc * (1/|m_j|); running max /
//                 first argmax; (soft) single-pass online softmax in base 2 against a lagged reference
//                 exponent, row scale gamma*log2e/|f_i| folded into one FFMA, sum of weights and weight * xyz
//                 accumulated in fp32; the halves are merged through shared memory at the end.
#include <float.h>

#include "gadm_internal.h"
#include "ptx.cuh"

namespace gadm {

namespace {

constexpr int BM = 128;               // rows (scene points) per CTA == UMMA M
constexpr int BN = 256;               // model vertices per accumulator tile == UMMA N
constexpr int BK = 64;                // bf16 elements per 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int A_BLK_BYTES = BM * BK * 2;     // 16 KB
constexpr int B_STAGE_BYTES = BN * BK * 2;   // 32 KB
constexpr int AUX_BYTES = BN * 16;           // 4 KB (float4 per vertex; argmax mode uses the first 1 KB)
constexpr int AUX_SLOTS = 4;                // aux ring is decoupled from the 2 accumulators so TMA can run ahead
constexpr int MAX_STAGES = 6;
constexpr int EPI_SUB = 4;                 // epilogue warps per TMEM lane quarter (column slices per tile)
constexpr int EPI_WARPS = 4 * EPI_SUB;
constexpr int NUM_THREADS = 64 + EPI_WARPS * 32;   // TMA warp, MMA warp, 16 epilogue warps
constexpr int XCH_BYTES = (EPI_SUB - 1) * BM * 8 * 4;  // per-row state exchange between the column slices
constexpr int TMEM_COLS = 512;

struct Barriers {
  uint64_t full[MAX_STAGES];
  uint64_t empty[MAX_STAGES];
  uint64_t a_full;
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t aux_full[AUX_SLOTS];
  uint64_t aux_empty[AUX_SLOTS];
  uint32_t tmem_base;
  uint32_t pad;
};

struct MatchParams {
  const float* rinv_rows;  // [B, N]
  const float* pad_sim;    // [B, N] or null
  const float* aux;        // [n_obj, M, 4] then [n_obj, M]
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

template <bool kSoft>
__global__ void __launch_bounds__(NUM_THREADS, 1)
match_kernel(const __grid_constant__ CUtensorMap tmap_rows, const __grid_constant__ CUtensorMap tmap_cols,
             const MatchParams p) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B atoms need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem_a + p.KB * A_BLK_BYTES;
  uint8_t* smem_aux = smem_b + p.stages * B_STAGE_BYTES;
  uint8_t* smem_xch = smem_aux + AUX_SLOTS * AUX_BYTES;
  Barriers* bars = reinterpret_cast<Barriers*>(smem_xch + XCH_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int row0 = blockIdx.x * BM;
  const int obj = frame_object(p, b);
  const int num_tiles = (p.M + BN - 1) / BN;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_rows);
    ptx::prefetch_tensormap(&tmap_cols);
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(&bars->full[s], 1);
      ptx::mbar_init(&bars->empty[s], 1);
    }
    ptx::mbar_init(&bars->a_full, 1);
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&bars->tmem_full[a], 1);
      ptx::mbar_init(&bars->tmem_empty[a], EPI_WARPS);  // one arrive per epilogue warp
    }
    for (int a = 0; a < AUX_SLOTS; ++a) {
      ptx::mbar_init(&bars->aux_full[a], 1);
      ptx::mbar_init(&bars->aux_empty[a], EPI_WARPS);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(&bars->tmem_base, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(&bars->a_full, p.KB * A_BLK_BYTES);
      for (int kb = 0; kb < p.KB; ++kb)
        ptx::tma_load_3d(smem_a + kb * A_BLK_BYTES, &tmap_rows, &bars->a_full, kb * BK, row0, b);
      int stage = 0;
      uint32_t phase = 0;
      // aux: per object, per 256-vertex tile, 1024 floats = [1/|m| x256 | x x256 | y x256 | z x256]
      const float* aux_tab = p.aux + size_t(obj) * num_tiles * (4 * BN);
      const uint32_t aux_bytes = kSoft ? uint32_t(AUX_BYTES) : uint32_t(BN * 4);  // ARGMAX needs only 1/|m|
      for (int t = 0; t < num_tiles; ++t) {
        const int slot = t % AUX_SLOTS;
        const uint32_t use = uint32_t(t) / AUX_SLOTS;
        ptx::mbar_wait_sleep(&bars->aux_empty[slot], (use & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&bars->aux_full[slot], aux_bytes);
        ptx::bulk_load_1d(smem_aux + slot * AUX_BYTES, aux_tab + size_t(t) * (4 * BN), aux_bytes,
                          &bars->aux_full[slot]);
        for (int kb = 0; kb < p.KB; ++kb) {
          ptx::mbar_wait_sleep(&bars->empty[stage], phase ^ 1);
          ptx::mbar_arrive_expect_tx(&bars->full[stage], B_STAGE_BYTES);
          ptx::tma_load_3d(smem_b + stage * B_STAGE_BYTES, &tmap_cols, &bars->full[stage], kb * BK, t * BN, obj);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ============================== UMMA issuer ==============================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16_f32(BM, BN);
      ptx::mbar_wait_sleep(&bars->a_full, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < num_tiles; ++t) {
        const int acc = t & 1;
        const uint32_t use = uint32_t(t) >> 1;
        ptx::mbar_wait_sleep(&bars->tmem_empty[acc], (use & 1) ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < p.KB; ++kb) {
          ptx::mbar_wait_sleep(&bars->full[stage], phase);
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
        ptx::umma_commit(&bars->tmem_full[acc]);  // accumulator tile complete
      }
    }
  } else {
    // ============================== epilogue (16 warps; thread == row, four threads per row) ==============================
    // warp w may only touch TMEM lanes [32*(w%4), +32); the four warps sharing a lane quarter each own a
    // 64-column slice of every 256-column tile and merge their per-row state at the end.  Four warps per
    // scheduler hide the TMEM-load / LDS / MUFU latencies that two could not (ncu: issue active 40 -> see profiles/).
    const int ew = warp - 2;
    const int q = warp & 3;
    const int sub = ew >> 2;                // which 64-column slice
    const int row_in_tile = q * 32 + lane;
    const int row = row0 + row_in_tile;
    const bool row_ok = row < p.N;
    const size_t grow = size_t(b) * p.N + (row_ok ? row : 0);
    const float rs = row_ok ? p.rinv_rows[grow] : 0.f;
    const float g = p.gamma_log2e * rs;  // exponent scale: t = (acc * 1/|m_j|) * g   (log2 units)

    float vmax = -INFINITY;
    int vidx = 0;
    float mrun = -INFINITY;  // set from the first chunk before any exp
    // packed (even-column | odd-column) partial sums, two independent sets (a/b) to shorten dependency chains
    uint64_t l2 = 0, ax2 = 0, ay2 = 0, az2 = 0, l2b = 0, ax2b = 0, ay2b = 0, az2b = 0;
    constexpr int CS = BN / EPI_SUB;    // columns per slice (64)
    constexpr int CW = 16;              // columns per TMEM load
    constexpr int NCH = CS / CW;        // chunks per slice per tile (4)

    for (int t = 0; t < num_tiles; ++t) {
      const int acc = t & 1;
      const uint32_t use = uint32_t(t) >> 1;
      const int slot = t % AUX_SLOTS;
      ptx::mbar_wait_sleep(&bars->aux_full[slot], (uint32_t(t) / AUX_SLOTS) & 1);
      ptx::mbar_wait_sleep(&bars->tmem_full[acc], use & 1);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + acc * BN + sub * CS;
      const int ncols = min(BN, p.M - t * BN) - sub * CS;  // valid columns in this slice (may be <= 0)
      // this slice of the aux slot (shared-space address): 1/|m| at [0,256), x at [256,512), y, z
      const uint32_t auxs = ptx::smem_u32(smem_aux + slot * AUX_BYTES) + sub * CS * 4;

      uint32_t ra[CW], rb[CW];
      if (kSoft && t == 0 && ncols > 0) {
        // reference exponent for the lagged online softmax: the first chunk's maximum
        ptx::tmem_ld_32x16(taddr, ra);
        ptx::tmem_ld_wait();
        float c0 = -FLT_MAX;
#pragma unroll
        for (int j = 0; j < CW; ++j)
          if (j < ncols) c0 = fmaxf(c0, __uint_as_float(ra[j]) * ptx::lds32(auxs + j * 4));
        mrun = c0 * g;
      }
      if (ncols > 0) ptx::tmem_ld_32x16(taddr, ra);

      auto process = [&](uint32_t (&r)[CW], int c) {
        const int cbase = c * CW;
        const uint32_t cmp = auxs + cbase * 4;
        // ---- scores: v = acc * 1/|m_j| (packed f32x2), kept in r[] for the argmax search
#pragma unroll
        for (int j4 = 0; j4 < CW / 4; ++j4) {
          const float4 cm = ptx::lds128(cmp + j4 * 16);
          const uint64_t v01 = ptx::fmul2(ptx::pack2(r[j4 * 4 + 0], r[j4 * 4 + 1]), ptx::pack2f(cm.x, cm.y));
          const uint64_t v23 = ptx::fmul2(ptx::pack2(r[j4 * 4 + 2], r[j4 * 4 + 3]), ptx::pack2f(cm.z, cm.w));
          ptx::unpack2(v01, r[j4 * 4 + 0], r[j4 * 4 + 1]);
          ptx::unpack2(v23, r[j4 * 4 + 2], r[j4 * 4 + 3]);
        }
        if (cbase + CW > ncols) {  // ragged last tile only (warp-uniform): TMA zero-fills columns >= M
#pragma unroll
          for (int j = 0; j < CW; ++j)
            if (cbase + j >= ncols) r[j] = __float_as_uint(-FLT_MAX);
        }
        // tree reduction (independent partial maxima: short dependency chains)
        float m4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          m4[u] = fmaxf(fmaxf(__uint_as_float(r[u * 4]), __uint_as_float(r[u * 4 + 1])),
                        fmaxf(__uint_as_float(r[u * 4 + 2]), __uint_as_float(r[u * 4 + 3])));
        const float cmx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        if (cmx > vmax) {  // strict: an equal value in a later chunk never displaces the first maximal index
          vmax = cmx;
          int j_lo = 7, j_hi = 15;   // two independent select chains, first hit wins
#pragma unroll
          for (int j = 6; j >= 0; --j) {
            if (__uint_as_float(r[j]) == cmx) j_lo = j;
            if (__uint_as_float(r[j + 8]) == cmx) j_hi = j + 8;
          }
          const bool lo_hit = fmaxf(m4[0], m4[1]) == cmx;
          vidx = t * BN + sub * CS + cbase + (lo_hit ? j_lo : j_hi);
        }
        if (kSoft) {
          // ---- p = 2^(v*g - m_run) against the lagged reference exponent; sums in packed f32x2 (even | odd column)
          const uint64_t g2 = ptx::pack2f(g, g), nm2 = ptx::pack2f(-mrun, -mrun);
#pragma unroll
          for (int j4 = 0; j4 < CW / 4; ++j4) {
            const float4 X = ptx::lds128(cmp + BN * 4 + j4 * 16);
            const float4 Y = ptx::lds128(cmp + 2 * BN * 4 + j4 * 16);
            const float4 Z = ptx::lds128(cmp + 3 * BN * 4 + j4 * 16);
            float t0, t1, t2, t3;
            ptx::unpack2f(ptx::ffma2(ptx::pack2(r[j4 * 4 + 0], r[j4 * 4 + 1]), g2, nm2), t0, t1);
            ptx::unpack2f(ptx::ffma2(ptx::pack2(r[j4 * 4 + 2], r[j4 * 4 + 3]), g2, nm2), t2, t3);
            const uint64_t p01 = ptx::pack2f(ptx::ex2_approx(t0), ptx::ex2_approx(t1));
            const uint64_t p23 = ptx::pack2f(ptx::ex2_approx(t2), ptx::ex2_approx(t3));
            l2 = ptx::fadd2(l2, p01);
            l2b = ptx::fadd2(l2b, p23);
            ax2 = ptx::ffma2(p01, ptx::pack2f(X.x, X.y), ax2);
            ax2b = ptx::ffma2(p23, ptx::pack2f(X.z, X.w), ax2b);
            ay2 = ptx::ffma2(p01, ptx::pack2f(Y.x, Y.y), ay2);
            ay2b = ptx::ffma2(p23, ptx::pack2f(Y.z, Y.w), ay2b);
            az2 = ptx::ffma2(p01, ptx::pack2f(Z.x, Z.y), az2);
            az2b = ptx::ffma2(p23, ptx::pack2f(Z.z, Z.w), az2b);
          }
          const float tnew = cmx * g;
          if (tnew > mrun) {  // rescale the running sums to the new reference exponent
            const float sc = ptx::ex2_approx(mrun - tnew);
            const uint64_t sc2 = ptx::pack2f(sc, sc);
            l2 = ptx::fmul2(l2, sc2); ax2 = ptx::fmul2(ax2, sc2); ay2 = ptx::fmul2(ay2, sc2); az2 = ptx::fmul2(az2, sc2);
            l2b = ptx::fmul2(l2b, sc2); ax2b = ptx::fmul2(ax2b, sc2); ay2b = ptx::fmul2(ay2b, sc2);
            az2b = ptx::fmul2(az2b, sc2);
            mrun = tnew;
          }
        }
      };

#pragma unroll
      for (int c = 0; c < NCH; c += 2) {
        if (c * CW < ncols) {
          ptx::tmem_ld_wait();
          if ((c + 1) * CW < ncols) ptx::tmem_ld_32x16(taddr + (c + 1) * CW, rb);
          process(ra, c);
        }
        if ((c + 1) * CW < ncols) {
          ptx::tmem_ld_wait();
          if ((c + 2) * CW < ncols && c + 2 < NCH) ptx::tmem_ld_32x16(taddr + (c + 2) * CW, ra);
          process(rb, c + 1);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(&bars->tmem_empty[acc]);
        ptx::mbar_arrive(&bars->aux_empty[slot]);
      }
    }

    // merge the four column slices of every row: slices 1..3 publish, slice 0 combines and writes the outputs
    float lsum, ax, ay, az;
    {
      float e, o;
      ptx::unpack2f(ptx::fadd2(l2, l2b), e, o); lsum = e + o;
      ptx::unpack2f(ptx::fadd2(ax2, ax2b), e, o); ax = e + o;
      ptx::unpack2f(ptx::fadd2(ay2, ay2b), e, o); ay = e + o;
      ptx::unpack2f(ptx::fadd2(az2, az2b), e, o); az = e + o;
    }
    float* xch = reinterpret_cast<float*>(smem_xch);
    if (sub > 0) {
      float* x = xch + ((sub - 1) * BM + row_in_tile) * 8;
      x[0] = vmax; x[1] = __int_as_float(vidx); x[2] = mrun; x[3] = lsum;
      x[4] = ax; x[5] = ay; x[6] = az;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
    if (sub == 0 && row_ok) {
      float mm = mrun;
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
        const float s0 = ptx::ex2_approx(mrun - mm);  // exp2(-inf) = 0 for a slice that saw no column
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
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

size_t match_smem_bytes(int KB, int stages) {
  return size_t(KB) * A_BLK_BYTES + size_t(stages) * B_STAGE_BYTES + AUX_SLOTS * AUX_BYTES + XCH_BYTES + sizeof(Barriers) + 1024;
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

int match_launch(const void* rows, const float* rinv_rows, const float* pad_sim, const void* cols, const float* aux,
                 const uint8_t* mask, const int32_t* obj_id, int B, int N, int M, int Kp, int n_obj, float gamma,
                 int pad_mode, int mode, int64_t* idx, float* max_sim, float* weight, float* soft_xyz,
                 cudaStream_t stream) {
  const int KB = Kp / BK;
  // pick the deepest ring that fits in 227 KB
  int stages = MAX_STAGES;
  while (stages > 2 && match_smem_bytes(KB, stages) > 227 * 1024) --stages;
  if (match_smem_bytes(KB, stages) > 227 * 1024) return GADM_ERR_UNSUPPORTED;

  CUtensorMap tmap_rows, tmap_cols;
  int rc = make_tmap_bf16_3d(&tmap_rows, rows, uint64_t(Kp), uint64_t(N), uint64_t(B), BK, BM);
  if (rc != GADM_OK) return rc;
  rc = make_tmap_bf16_3d(&tmap_cols, cols, uint64_t(Kp), uint64_t(M), uint64_t(n_obj), BK, BN);
  if (rc != GADM_OK) return rc;

  MatchParams p;
  p.rinv_rows = rinv_rows; p.pad_sim = pad_sim; p.aux = aux; p.mask = mask; p.obj_id = obj_id;
  p.idx = idx; p.max_sim = max_sim; p.weight = weight; p.soft_xyz = soft_xyz;
  p.B = B; p.N = N; p.M = M; p.KB = KB; p.n_obj = n_obj; p.stages = stages; p.pad_mode = pad_mode;
  p.gamma_log2e = gamma * 1.4426950408889634f;

  dim3 grid((N + BM - 1) / BM, B);
  const size_t smem = match_smem_bytes(KB, stages);
  if (mode == GADM_MATCH_SOFT)
    match_kernel<true><<<grid, NUM_THREADS, smem, stream>>>(tmap_rows, tmap_cols, p);
  else
    match_kernel<false><<<grid, NUM_THREADS, smem, stream>>>(tmap_rows, tmap_cols, p);
  return check_launch();
}

}  // namespace gadm
