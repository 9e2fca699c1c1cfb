// Micro-benchmark (not product code): tensor-pipe cost of the FA-style SOFT unit of the matcher.
// One "unit" = 128 rows x 128 model vertices at K' = 128:
//   8 x tcgen05.mma 128x128x16 (bf16, A from TMEM, B from shared memory) into S[r]
//   8 x tcgen05.mma 128x16x16  (fp16, A = P from TMEM aliasing S[r], B = V tile from shared memory) into 4 O accumulators
// Variants time the two groups alone and interleaved the way the kernel issues them.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/pv_probe tools/probes/pv_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../geometric-aware-dense-matching_b200/csrc/ptx.cuh"

using namespace gadm;

// load: 0 = the other warps idle, 1 = they stream tcgen05.ld x32 over the S region, 2 = ld x32 + st x16 (the epilogue's
// traffic pattern) while warp 0 issues the MMAs
__global__ void __launch_bounds__(512, 1) probe(long long* out, int reps, int load) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint64_t dummy[4];
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    for (int i = 0; i < 4; ++i) ptx::mbar_init(&dummy[i], 1);
    ptx::fence_mbar_init();
  }
  if (threadIdx.x < 32) { ptx::tmem_alloc(&tmem_slot, 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tm = tmem_slot;
  {
    uint32_t z[16];
    for (int i = 0; i < 16; ++i) z[i] = 0x3c003c00u;
    const uint32_t lane_base = tm + (uint32_t(((threadIdx.x >> 5) & 3) * 32) << 16);
    if (threadIdx.x < 128) for (int c = 0; c < 512; c += 16) ptx::tmem_st_32x16(lane_base + c, z);
    ptx::tmem_st_wait();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();

  __shared__ volatile int stop_flag;
  if (threadIdx.x == 0) stop_flag = 0;
  __syncthreads();
  if (threadIdx.x >= 128 && load > 0) {
    // background TMEM traffic: 12 warps (lane quarter = warp % 4) read (and write) 32 / 16 columns of S[0] / S[1]
    const uint32_t lane_base = tm + (uint32_t(((threadIdx.x >> 5) & 3) * 32) << 16);
    uint32_t acc = 0;
    while (!stop_flag) {
      uint32_t d[32];
      ptx::tmem_ld_32x32(lane_base + ((threadIdx.x >> 7) & 1) * 128 + 64, d);
      ptx::tmem_ld_wait();
      for (int i = 0; i < 32; ++i) acc += d[i];
      if (load > 1) {
        uint32_t z[16];
        for (int i = 0; i < 16; ++i) z[i] = 0x3c003c00u;
        ptx::tmem_st_32x16(lane_base + ((threadIdx.x >> 7) & 1) * 128 + 96, z);
        ptx::tmem_st_wait();
      }
    }
    if (acc == 0x12345678u) out[31] = acc;
  }
  if (threadIdx.x == 0) {
    const uint32_t b_addr = ptx::smem_u32(smem);                  // B stages: 2 x 16 KB (128 vertices x 64 k)
    const uint32_t v_addr = ptx::smem_u32(smem + 64 * 1024);      // V tile: 2 blocks of [16 rows x 64 k] fp16
    const uint32_t bss_a = ptx::smem_u32(smem + 32 * 1024);       // A blocks for the SS comparison
    uint32_t phase = 0;
    // TMEM map: S[0] 0..127, S[1] 128..255, A[0] 256..319, A[1] 320..383, O 384..511
    auto run = [&](int variant) -> long long {
      const uint32_t id_s = ptx::umma_idesc_bf16_f32(128, 128);
      const uint32_t id_s256 = ptx::umma_idesc_bf16_f32(128, 256);
      const uint32_t id_pv = ptx::umma_idesc_f16_f32(128, 16);
      const uint32_t id_pv32 = ptx::umma_idesc_f16_f32(128, 32);
      const uint32_t id_pv64 = ptx::umma_idesc_f16_f32(128, 64);
      const long long t0 = clock64();
      for (int rep = 0; rep < reps; ++rep) {
        const int r = rep & 1;
        const uint32_t S = tm + r * 128, A = tm + 256 + r * 64, O = tm + 384 + r * 64;
        auto s_mmas = [&]() {
          for (int k = 0; k < 8; ++k)
            ptx::umma_f16_ts(S, A + k * 8, ptx::umma_desc_sw128_kmajor(b_addr + (k >> 2) * 16384 + (k & 3) * 32), id_s, k != 0);
        };
        auto pv_mmas = [&](uint32_t idesc, int ostride) {
          for (int s = 0; s < 4; ++s)
            for (int k = 0; k < 2; ++k)
              ptx::umma_f16_ts(O + s * ostride, S + s * 32 + k * 8,
                               ptx::umma_desc_sw128_kmajor(v_addr + (s >> 1) * 2048 + ((s & 1) * 2 + k) * 32), idesc, 1);
        };
        switch (variant) {
          case 0: s_mmas(); break;                                  // S only (TS)
          case 1: pv_mmas(id_pv, 16); break;                        // PV only, 4 accumulators x 2
          case 2: pv_mmas(id_pv, 16); s_mmas(); break;              // the kernel's order
          case 3:                                                   // PV into ONE accumulator (8 dependent)
            for (int s = 0; s < 4; ++s)
              for (int k = 0; k < 2; ++k)
                ptx::umma_f16_ts(O, S + s * 32 + k * 8, ptx::umma_desc_sw128_kmajor(v_addr + (s >> 1) * 2048 + ((s & 1) * 2 + k) * 32), id_pv, 1);
            break;
          case 4:                                                   // interleaved one PV after every S MMA
            for (int k = 0; k < 8; ++k) {
              ptx::umma_f16_ts(S, A + k * 8, ptx::umma_desc_sw128_kmajor(b_addr + (k >> 2) * 16384 + (k & 3) * 32), id_s, k != 0);
              ptx::umma_f16_ts(O + (k >> 1) * 16, tm + (r ^ 1) * 128 + (k >> 1) * 32 + (k & 1) * 8,
                               ptx::umma_desc_sw128_kmajor(v_addr + (k >> 2) * 2048 + (k & 3) * 32), id_pv, 1);
            }
            break;
          case 5:                                                   // S as SS MMAs (A from shared memory), N = 128
            for (int k = 0; k < 8; ++k)
              ptx::umma_bf16_ss(S, ptx::umma_desc_sw128_kmajor(bss_a + (k >> 2) * 16384 + (k & 3) * 32),
                                ptx::umma_desc_sw128_kmajor(b_addr + (k >> 2) * 16384 + (k & 3) * 32), id_s, k != 0);
            break;
          case 6: pv_mmas(id_pv32, 8); break;                       // N = 32 (overlapping accumulators; timing only)
          case 7:                                                   // TS N = 256 over two S buffers (timing only)
            for (int k = 0; k < 8; ++k)
              ptx::umma_f16_ts(tm, A + k * 8, ptx::umma_desc_sw128_kmajor(b_addr + (k & 3) * 32), id_s256, k != 0);
            break;
          case 8:                                                   // PV with K covered by ONE accumulator per 2 slices
            for (int s = 0; s < 4; ++s)
              for (int k = 0; k < 2; ++k)
                ptx::umma_f16_ts(O + (s >> 1) * 16, S + s * 32 + k * 8, ptx::umma_desc_sw128_kmajor(v_addr + (s >> 1) * 2048 + ((s & 1) * 2 + k) * 32), id_pv, 1);
            break;
          case 9: pv_mmas(id_pv64, 0); break;                       // N = 64 (timing only)
          case 10: s_mmas(); ptx::umma_commit(&dummy[0]); break;    // S + one commit per unit (nobody waits on it)
          case 11: s_mmas(); ptx::umma_commit(&dummy[0]); ptx::umma_commit(&dummy[1]); ptx::umma_commit(&dummy[2]); break;
          case 12: pv_mmas(id_pv, 16); s_mmas(); ptx::umma_commit(&dummy[0]); ptx::umma_commit(&dummy[1]); break;
          case 13:                                                  // 8 x SS 128x256x16 (the alt kernel's unit) + 2 commits
            for (int k = 0; k < 8; ++k)
              ptx::umma_bf16_ss(tm, ptx::umma_desc_sw128_kmajor(bss_a + (k >> 2) * 16384 + (k & 3) * 32),
                                ptx::umma_desc_sw128_kmajor(b_addr + (k & 3) * 32), id_s256, k != 0);
            ptx::umma_commit(&dummy[0]); ptx::umma_commit(&dummy[1]);
            break;
          case 14:                                                  // the same without commits
            for (int k = 0; k < 8; ++k)
              ptx::umma_bf16_ss(tm, ptx::umma_desc_sw128_kmajor(bss_a + (k >> 2) * 16384 + (k & 3) * 32),
                                ptx::umma_desc_sw128_kmajor(b_addr + (k & 3) * 32), id_s256, k != 0);
            break;
          case 15: {                                                // S, commit, and WAIT for it every unit (full serialisation)
            s_mmas(); ptx::umma_commit(&dummy[3]); ptx::mbar_wait(&dummy[3], rep & 1);
          } break;
        }
      }
      ptx::umma_commit(&bar);
      ptx::mbar_wait(&bar, phase);
      phase ^= 1;
      return clock64() - t0;
    };
    for (int v = 0; v < 16; ++v) {
      if (v == 15) {                               // dummy[3]'s phase must restart at 0 for each run of variant 15
        ptx::mbar_init(&dummy[3], 1); ptx::fence_mbar_init();
      }
      const long long c = run(v);
      if (blockIdx.x == 0) out[v] = c;
    }
    stop_flag = 1;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { ptx::tc_fence_after(); ptx::tmem_dealloc(tm, 512); }
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 32 * sizeof(long long));
  const int reps = 256;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int load = 0; load < 3; ++load) {
  printf("---- background TMEM traffic from 12 warps: %s\n", load == 0 ? "none" : load == 1 ? "tcgen05.ld x32" : "tcgen05.ld x32 + st x16");
  probe<<<148, 512, 100 * 1024>>>(out, reps, load);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  const char* names[16] = {"S: 8 x TS 128x128x16 (A in TMEM)",
                           "PV: 8 x TS 128x16x16 into 4 accumulators",
                           "PV then S (kernel order)",
                           "PV: 8 x TS 128x16x16 into ONE accumulator",
                           "S and PV interleaved 1:1",
                           "S: 8 x SS 128x128x16 (A in shared memory)",
                           "PV: 8 x TS 128x32x16",
                           "8 x TS 128x256x16",
                           "PV: 8 x TS 128x16x16 into 2 accumulators",
                           "PV: 8 x TS 128x64x16",
                           "S + 1 commit per unit",
                           "S + 3 commits per unit",
                           "PV, S + 2 commits per unit",
                           "8 x SS 128x256x16 + 2 commits per unit",
                           "8 x SS 128x256x16, no commits",
                           "S + commit + wait per unit"};
  for (int v = 0; v < 16; ++v) printf("%-48s %8.1f cycles per unit\n", names[v], double(out[v]) / reps);
  }
  return 0;
}
