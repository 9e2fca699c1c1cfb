// Feature-space kNN (DGCNN) on the tensor cores: models/dgcnn.py:21-27 for C % 64 == 0, k <= 20.
//
//   pd_ij = -|x_i|^2 + 2 x_i.x_j - |x_j|^2 ;  idx = pd.topk(k)[1]        (reference: a [B, N, N] fp32 matrix + topk)
//
// The dot products run as a bf16x3 split on tcgen05 -- x = hi + lo with hi = bf16(x), lo = bf16(x - hi), and
// x_i.x_j ~ hi.hi + hi.lo + lo.hi (what is dropped, lo.lo, is 2^-16 of the product) -- with fp32 accumulation in
// tensor memory; |x|^2 stays fp32 and the score keeps the reference's form ((-xx_i) + 2 dot) - xx_j.  ONE operand array
// [B, N, 2C] = [hi | lo] serves both sides: the three partial products are three passes over K blocks addressed with
// different TMA coordinates (A: hi, hi, lo;  B: hi, lo, hi), so nothing is stored twice.
//
// Skeleton of match_kernel<., 1> (match_sm100.cu): CTA = 128 query rows of one cloud, 256-column candidate tiles,
// two TMEM accumulators alternating between tiles, warp 16 TMA, warp 17 UMMA, warps 0-15 epilogue (thread = row x
// the 64-column slice w / 4).  Epilogue: every thread keeps the 20 best (value, index) of its slice in registers,
// descending.  Per 32-column chunk: scores, a max tree -- the chunk is skipped when no lane's maximum beats its 20th
// value --, a bit mask of the columns that do, and ONE loop in which every lane inserts its next marked column (a
// 5-level select tree picks the value; the insertion is a 20-step bubble on the value alone: a thread sees its
// columns in ascending order, so a later equal value never displaces an earlier one and ties resolve to the smaller
// index without comparing indices).  One insertion site keeps the loop inside the instruction cache.  The four sorted
// lists of a row go through shared memory at the end and the row's k best come out of a 4-way merge of their heads
// (value descending, index ascending).
#include <cuda_bf16.h>

#include "match_common.cuh"

namespace gadm {

namespace {

constexpr int KF_K = 20;          // list length: k <= 20
constexpr int KF_SL = 4;          // column slices per row
constexpr int KF_CS = BN / KF_SL; // 64 columns per slice
constexpr int KF_MAX_CB = 4;      // C <= 256

struct KfParams {
  const float* xx;     // [B, N]  |x_j|^2, fp32
  int64_t* idx;        // [B, N, k]
  int B, N, CB, k, stages;   // CB = C / 64
};

// x [B, C, N] fp32 -> x2 [B, N, 2C] bf16 = [hi | lo], xx [B, N]
__global__ void __launch_bounds__(256)
knn_feat_split_kernel(const float* __restrict__ x, int C, int N, __nv_bfloat16* __restrict__ x2,
                      float* __restrict__ xx) {
  extern __shared__ float tile[];   // [C][33]
  const int b = blockIdx.y, n0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* xb = x + size_t(b) * C * N;
  for (int c = ty; c < C; c += 8) tile[c * 33 + tx] = n0 + tx < N ? xb[size_t(c) * N + n0 + tx] : 0.f;
  __syncthreads();
  // thread (tx = channel lane, ty = point group): points ty, ty + 8, ...
  for (int pnt = ty; pnt < 32; pnt += 8) {
    const int n = n0 + pnt;
    float s = 0.f;
    for (int c = tx; c < C; c += 32) {
      const float v = tile[c * 33 + pnt];
      s = fmaf(v, v, s);
      if (n < N) {
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        x2[(size_t(b) * N + n) * (2 * C) + c] = hi;
        x2[(size_t(b) * N + n) * (2 * C) + C + c] = __float2bfloat16_rn(v - __bfloat162float(hi));
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (tx == 0 && n < N) xx[size_t(b) * N + n] = s;
  }
}

struct KfList {
  float v[KF_K];
  int i[KF_K];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int s = 0; s < KF_K; ++s) { v[s] = -INFINITY; i[s] = 0x7fffffff; }
  }
  // descending by value; an equal value goes BEHIND what is already there.  No carry chain: the list is sorted, so
  // "nv beats entry s" is monotone in s and every slot can be written from the OLD slots s and s - 1 alone -- 20
  // independent compares and 80 independent selects instead of a 20-step bubble (the bubble's FSETP -> SEL chain kept
  // the epilogue at 0.45 instructions per cycle and scheduler).
  __device__ __forceinline__ void insert(float nv, int ni) {
    bool sw[KF_K];
#pragma unroll
    for (int s = 0; s < KF_K; ++s) sw[s] = nv > v[s];
#pragma unroll
    for (int s = KF_K - 1; s > 0; --s) {
      v[s] = sw[s] ? (sw[s - 1] ? v[s - 1] : nv) : v[s];
      i[s] = sw[s] ? (sw[s - 1] ? i[s - 1] : ni) : i[s];
    }
    v[0] = sw[0] ? nv : v[0];
    i[0] = sw[0] ? ni : i[0];
  }
};

__device__ __forceinline__ float sel32(const float (&a)[32], int j) {
  float b[16], c[8], d[4];
#pragma unroll
  for (int t = 0; t < 16; ++t) b[t] = (j & 1) ? a[2 * t + 1] : a[2 * t];
#pragma unroll
  for (int t = 0; t < 8; ++t) c[t] = (j & 2) ? b[2 * t + 1] : b[2 * t];
#pragma unroll
  for (int t = 0; t < 4; ++t) d[t] = (j & 4) ? c[2 * t + 1] : c[2 * t];
  const float e0 = (j & 8) ? d[1] : d[0], e1 = (j & 8) ? d[3] : d[2];
  return (j & 16) ? e1 : e0;
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
knn_feat_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                   const KfParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                                        // [2 CB] blocks of 128 rows x 64 k: hi.., lo..
  uint8_t* smem_b = smem_a + 2 * p.CB * A_BLK_BYTES;             // ring of 256 x 64 k blocks
  uint8_t* smem_aux = smem_b + p.stages * B_STAGE_BYTES;         // per slot: |x_j|^2 x 256
  volatile float* smem_thr = reinterpret_cast<volatile float*>(smem_aux + AUX_SLOTS * PLANE_BYTES);   // [KF_SL][BM]
  Barriers* bars = reinterpret_cast<Barriers*>(smem_aux + AUX_SLOTS * PLANE_BYTES + KF_SL * BM * 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int row0 = blockIdx.x * BM;
  const int num_tiles = (p.N + BN - 1) / BN;
  const int KS = 3 * p.CB;                                       // K blocks per tile: hi.hi, hi.lo, lo.hi

  if (warp == EPI_WARPS && lane == 0) {
    ptx::prefetch_tensormap(&tmap_a);
    ptx::prefetch_tensormap(&tmap_b);
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
  if (threadIdx.x < KF_SL * BM) smem_thr[threadIdx.x] = -INFINITY;
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == EPI_WARPS) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(&bars->a_full, 2 * p.CB * A_BLK_BYTES);
      for (int kb = 0; kb < 2 * p.CB; ++kb)     // rows >= N are zero-filled by TMA
        ptx::tma_load_3d(smem_a + kb * A_BLK_BYTES, &tmap_a, &bars->a_full, kb * BK, row0, b);
      int stage = 0;
      uint32_t phase = 0;
      const float* xx_tab = p.xx + size_t(b) * p.N;
      for (int t = 0; t < num_tiles; ++t) {
        const int slot = t % AUX_SLOTS;
        const uint32_t use = uint32_t(t) / AUX_SLOTS;
        const uint32_t bytes = uint32_t(min(BN, p.N - t * BN)) * 4;     // N % 4 == 0: a multiple of 16
        ptx::mbar_wait_sleep(&bars->aux_empty[slot], (use & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&bars->aux_full[slot], bytes);
        ptx::bulk_load_1d(smem_aux + slot * PLANE_BYTES, xx_tab + size_t(t) * BN, bytes, &bars->aux_full[slot]);
        for (int s = 0; s < KS; ++s) {
          // B side of pass s: hi blocks, lo blocks, hi blocks again
          const int blk = s < p.CB ? s : (s < 2 * p.CB ? s : s - 2 * p.CB);
          ptx::mbar_wait_sleep(&bars->empty[stage], phase ^ 1);
          ptx::mbar_arrive_expect_tx(&bars->full[stage], B_STAGE_BYTES);
          ptx::tma_load_3d(smem_b + stage * B_STAGE_BYTES, &tmap_b, &bars->full[stage], blk * BK, t * BN, b);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == EPI_WARPS + 1) {
    // ============================== UMMA issuer ==============================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16_f32(BM, BN);
      constexpr uint64_t DESC_HI = uint64_t((1024u >> 4) | (1u << 14) | (2u << 29)) << 32;   // see match_pair_kernel
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
        for (int s = 0; s < KS; ++s) {
          // A side of pass s: hi blocks, hi blocks again, lo blocks
          const int ablk = s < p.CB ? s : (s < 2 * p.CB ? s - p.CB : s - p.CB);
          ptx::mbar_wait_sleep(&bars->full[stage], phase);
          ptx::tc_fence_after();
          const uint32_t a_lo = a_lo0 + uint32_t(ablk) * (A_BLK_BYTES >> 4);
          const uint32_t b_lo = b_lo0 + uint32_t(stage) * (B_STAGE_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            ptx::umma_bf16_ss(d_tmem, DESC_HI | (a_lo + k * (UMMA_K * 2 >> 4)), DESC_HI | (b_lo + k * (UMMA_K * 2 >> 4)),
                              idesc, (s | k) != 0);
          ptx::umma_commit(&bars->empty[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(&bars->s_full[acc]);
      }
    }
  } else {
    // ============================== epilogue warps (thread == row x 64-column slice) ==============
    const int q = warp & 3, sub = warp >> 2;
    const int row_in_tile = q * 32 + lane;
    const int row = row0 + row_in_tile;
    const bool row_ok = row < p.N;
    const float nxx = row_ok ? -p.xx[size_t(b) * p.N + row] : 0.f;
    const uint32_t lane_base = tmem_base + (uint32_t(q * 32) << 16) + sub * KF_CS;
    KfList L;
    L.init();

    for (int t = 0; t < num_tiles; ++t) {
      const int acc = t & 1;
      const uint32_t use = uint32_t(t) >> 1;
      const int slot = t % AUX_SLOTS;
      if (!(ptx::mbar_try_wait(&bars->aux_full[slot], (uint32_t(t) / AUX_SLOTS) & 1) &
            ptx::mbar_try_wait(&bars->s_full[acc], use & 1))) {
        ptx::mbar_wait_sleep(&bars->aux_full[slot], (uint32_t(t) / AUX_SLOTS) & 1);
        ptx::mbar_wait_sleep(&bars->s_full[acc], use & 1);
      }
      ptx::tc_fence_after();
      const int ncols = min(BN, p.N - t * BN) - sub * KF_CS;       // valid columns of this slice (may be <= 0)
      const uint32_t xx_addr = ptx::smem_u32(smem_aux + slot * PLANE_BYTES) + sub * KF_CS * 4;
      const int col_base = t * BN + sub * KF_CS;
#pragma unroll 1
      for (int c2 = 0; c2 < KF_CS / 32; ++c2) {
        const int nv = ncols - c2 * 32;
        if (nv <= 0) break;
        uint32_t r[32];
        ptx::tmem_ld_32x32(lane_base + acc * BN + c2 * 32, r);
        ptx::tmem_ld_wait();
        float pd[32];
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 x4 = ptx::lds128(xx_addr + (c2 * 32 + j4 * 4) * 4);
          pd[j4 * 4 + 0] = __fsub_rn(fmaf(__uint_as_float(r[j4 * 4 + 0]), 2.f, nxx), x4.x);
          pd[j4 * 4 + 1] = __fsub_rn(fmaf(__uint_as_float(r[j4 * 4 + 1]), 2.f, nxx), x4.y);
          pd[j4 * 4 + 2] = __fsub_rn(fmaf(__uint_as_float(r[j4 * 4 + 2]), 2.f, nxx), x4.z);
          pd[j4 * 4 + 3] = __fsub_rn(fmaf(__uint_as_float(r[j4 * 4 + 3]), 2.f, nxx), x4.w);
        }
        if (nv < 32) {   // ragged last tile: TMA zero-fills the rows and the |x|^2 behind them are stale
#pragma unroll
          for (int j = 0; j < 32; ++j) pd[j] = j < nv ? pd[j] : -INFINITY;
        }
        // chunk maximum: nothing to do when no lane's beats its current k-th value
        float m[11];
#pragma unroll
        for (int j = 0; j < 10; ++j) m[j] = ptx::fmax3(pd[3 * j], pd[3 * j + 1], pd[3 * j + 2]);
        m[10] = fmaxf(pd[30], pd[31]);
        const float mm = ptx::fmax3(ptx::fmax3(m[0], m[1], m[2]), ptx::fmax3(m[3], m[4], m[5]),
                                    ptx::fmax3(ptx::fmax3(m[6], m[7], m[8]), m[9], m[10]));
        // A value enters this slice's list if it beats the slice's own 20th value (strictly: the thread sees its
        // columns in ascending order) AND is not below the ROW bound: every slice publishes its 5th best value, and
        // the smallest of the four is a value that at least 4 x 5 = 20 candidates of the row reach, so nothing below
        // it can be in the row's final list.  The four slices hold statistically equal shares of the row, so this
        // bound sits close to the row's true 20th value -- the slices together then do about the insertions of ONE
        // list over the whole row instead of four.  Published and read without synchronisation: the values only
        // grow, a stale one is only a weaker bound.
        const float kth = L.v[KF_K - 1];
        const float rowb = fminf(fminf(smem_thr[row_in_tile], smem_thr[BM + row_in_tile]),
                                 fminf(smem_thr[2 * BM + row_in_tile], smem_thr[3 * BM + row_in_tile]));
        if (!__any_sync(0xffffffffu, mm > kth && mm >= rowb)) continue;
        uint32_t mask = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) mask |= (pd[j] > kth && pd[j] >= rowb) ? (1u << j) : 0u;
#pragma unroll 1
        while (__any_sync(0xffffffffu, mask != 0)) {
          const int j = mask ? __ffs(mask) - 1 : 0;
          const float v = mask ? sel32(pd, j) : -INFINITY;     // -inf never moves anything
          mask &= mask - 1;
          L.insert(v, col_base + c2 * 32 + j);
        }
        smem_thr[sub * BM + row_in_tile] = L.v[KF_K / KF_SL - 1];
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(&bars->s_free[acc]);
        ptx::mbar_arrive(&bars->aux_empty[slot]);
      }
    }

    // ---- merge the four slices of every row: every slice's list is sorted, so the row's k best come out of a 4-way
    // merge of the list heads (exchange buffer: the operand tiles -- every MMA has completed -- [KF_SL][KF_K][BM])
    float2* xch = reinterpret_cast<float2*>(smem_a);
#pragma unroll
    for (int s = 0; s < KF_K; ++s)
      xch[(sub * KF_K + s) * BM + row_in_tile] = make_float2(L.v[s], __int_as_float(L.i[s]));
    asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
    if (sub == 0 && row_ok) {
      int64_t* o = p.idx + (size_t(b) * p.N + row) * p.k;
      int h[KF_SL] = {0, 0, 0, 0};
      float2 head[KF_SL];
#pragma unroll
      for (int q2 = 0; q2 < KF_SL; ++q2) head[q2] = xch[(q2 * KF_K) * BM + row_in_tile];
      for (int r = 0; r < p.k; ++r) {
        // pick the best head: value descending, index ascending
        float bv = head[0].x;
        int bi = __float_as_int(head[0].y);
        int best = 0;
#pragma unroll
        for (int q2 = 1; q2 < KF_SL; ++q2) {
          const float cv = head[q2].x;
          const int ci = __float_as_int(head[q2].y);
          const bool w = cv > bv || (cv == bv && ci < bi);
          bv = w ? cv : bv; bi = w ? ci : bi; best = w ? q2 : best;
        }
        o[r] = bi;
        // advance the winner's list (an exhausted list shows -inf)
#pragma unroll
        for (int q2 = 0; q2 < KF_SL; ++q2) {
          if (q2 == best) {
            ++h[q2];
            head[q2] = h[q2] < KF_K ? xch[(q2 * KF_K + h[q2]) * BM + row_in_tile] : make_float2(-INFINITY, __int_as_float(0x7fffffff));
          }
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

size_t kf_smem_bytes(int CB, int stages) {
  return size_t(2) * CB * A_BLK_BYTES + size_t(stages) * B_STAGE_BYTES + AUX_SLOTS * PLANE_BYTES + KF_SL * BM * 4 +
         sizeof(Barriers) + 1024;
}
int kf_stages(int CB) {
  int stages = MAX_STAGES;
  while (stages > 0 && kf_smem_bytes(CB, stages) > 227 * 1024) --stages;
  return stages;
}

}  // namespace

int knn_feat_tc_configure() {
  cudaError_t e = cudaFuncSetAttribute(knn_feat_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return set_cuda_error(e);
  e = cudaFuncSetAttribute(knn_feat_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  return e == cudaSuccess ? GADM_OK : set_cuda_error(e);
}

bool knn_feat_tc_supported(int C, int N, int kdim, int k) {
  if (C % 64 != 0 || C > 64 * KF_MAX_CB || kdim != C || k > KF_K || N % 4 != 0 || N < BN) return false;
  const int CB = C / 64, stages = kf_stages(CB);
  // the exchange buffer of the final merge lives in the operand tiles (row tile + ring)
  return stages >= 2 && size_t(2) * CB * A_BLK_BYTES + size_t(stages) * B_STAGE_BYTES >= size_t(KF_SL) * BM * KF_K * 8;
}

size_t knn_feat_tc_workspace_bytes(int B, int C, int N) {
  return (size_t(B) * N * 2 * C * 2 + 255) / 256 * 256 + size_t(B) * N * 4;
}

int knn_feat_tc_launch(const float* x, int B, int C, int N, int k, int64_t* idx, void* workspace, size_t workspace_bytes,
                       cudaStream_t stream) {
  if (B > 65535) return GADM_ERR_UNSUPPORTED;
  if (!knn_feat_tc_supported(C, N, C, k)) return GADM_ERR_UNSUPPORTED;
  if (!workspace || workspace_bytes < knn_feat_tc_workspace_bytes(B, C, N)) return GADM_ERR_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return GADM_ERR_ALIGN;
  __nv_bfloat16* x2 = static_cast<__nv_bfloat16*>(workspace);
  float* xx = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + (size_t(B) * N * 2 * C * 2 + 255) / 256 * 256);
  knn_feat_split_kernel<<<dim3((N + 31) / 32, B), 256, size_t(C) * 33 * 4, stream>>>(x, C, N, x2, xx);
  int rc = check_launch();
  if (rc != GADM_OK) return rc;
  KfParams p;
  p.xx = xx; p.idx = idx; p.B = B; p.N = N; p.CB = C / 64; p.k = k; p.stages = kf_stages(p.CB);
  CUtensorMap ta, tb;
  rc = make_tmap_2b_3d(&ta, x2, uint64_t(2 * C), uint64_t(N), uint64_t(B), BK, BM, 0);
  if (rc != GADM_OK) return rc;
  rc = make_tmap_2b_3d(&tb, x2, uint64_t(2 * C), uint64_t(N), uint64_t(B), BK, BN, 0);
  if (rc != GADM_OK) return rc;
  knn_feat_tc_kernel<<<dim3((N + BM - 1) / BM, B), NUM_THREADS, kf_smem_bytes(p.CB, p.stages), stream>>>(ta, tb, p);
  return check_launch();
}

}  // namespace gadm
