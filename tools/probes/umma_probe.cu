// Micro-benchmark (not product code): issue cost of small-N tcgen05.mma next to the 128x256x16 similarity MMAs.
// Answers one design question for the matcher's SOFT mode: can the sums of p and p * xyz ride on the tensor core
// (A = P from TMEM or shared memory, B = [xyz_hi | xyz_lo | 1] with N = 8 or 16) without stalling the main GEMM?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/umma_probe tools/umma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../geometric-aware-dense-matching_b200/csrc/ptx.cuh"

using namespace gadm;

__global__ void __launch_bounds__(128, 1) probe(long long* out, int reps) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); }
  if (threadIdx.x < 32) { ptx::tmem_alloc(&tmem_slot, 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tm = tmem_slot;
  // zero the TMEM region used as the A operand of the TS variants (columns 256..383)
  {
    uint32_t z[16];
    for (int i = 0; i < 16; ++i) z[i] = 0x3c003c00u;
    const uint32_t lane_base = tm + (uint32_t((threadIdx.x >> 5) * 32) << 16);
    for (int c = 0; c < 128; c += 16) ptx::tmem_st_32x16(lane_base + 256 + c, z);
    ptx::tmem_st_wait();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();

  if (threadIdx.x == 0) {
    const uint32_t a_addr = ptx::smem_u32(smem);                  // 16 KB A block
    const uint32_t b_addr = ptx::smem_u32(smem + 32 * 1024);      // 32 KB B stage
    const uint32_t v_addr = ptx::smem_u32(smem + 64 * 1024);      // small B operand (N x 64 k, 128-byte rows)
    uint32_t phase = 0;
    auto run = [&](int variant) -> long long {
      const uint32_t id_big = ptx::umma_idesc_f16_f32(128, 256);
      const uint32_t id_n8 = ptx::umma_idesc_f16_f32(128, 8);
      const uint32_t id_n16 = ptx::umma_idesc_f16_f32(128, 16);
      const uint32_t id_n32 = ptx::umma_idesc_f16_f32(128, 32);
      const long long t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        const uint32_t d_big = tm + (r & 1) * 0;   // same accumulator: dependent chain like the real kernel's k loop
        if (variant == 0 || (variant >= 4 && variant < 10)) {
          for (int k = 0; k < 8; ++k)
            ptx::umma_bf16_ss(d_big, ptx::umma_desc_sw128_kmajor(a_addr + (k & 3) * 32),
                              ptx::umma_desc_sw128_kmajor(b_addr + (k & 3) * 32), id_big, k != 0);
        }
        if (variant == 1 || variant == 4)          // 16 x TS, N = 8
          for (int k = 0; k < 16; ++k)
            ptx::umma_f16_ts(tm + 480, tm + 256 + k * 8, ptx::umma_desc_sw128_kmajor(v_addr + (k & 3) * 32), id_n8, k != 0);
        if (variant == 2 || variant == 5)          // 16 x TS, N = 16
          for (int k = 0; k < 16; ++k)
            ptx::umma_f16_ts(tm + 480, tm + 256 + k * 8, ptx::umma_desc_sw128_kmajor(v_addr + (k & 3) * 32), id_n16, k != 0);
        if (variant == 3 || variant == 6)          // 16 x SS, N = 8 (A = P staged in shared memory)
          for (int k = 0; k < 16; ++k)
            ptx::umma_bf16_ss(tm + 480, ptx::umma_desc_sw128_kmajor(a_addr + (k & 3) * 32),
                              ptx::umma_desc_sw128_kmajor(v_addr + (k & 3) * 32), id_n8, k != 0);
        if (variant == 7)                          // 16 x TS, N = 32
          for (int k = 0; k < 16; ++k)
            ptx::umma_f16_ts(tm + 448, tm + 256 + k * 8, ptx::umma_desc_sw128_kmajor(v_addr + (k & 3) * 32), id_n32, k != 0);
        if (variant == 8)                          // 4 x TS N = 8 only (per-instruction latency vs throughput)
          for (int k = 0; k < 4; ++k)
            ptx::umma_f16_ts(tm + 480, tm + 256 + k * 8, ptx::umma_desc_sw128_kmajor(v_addr + (k & 3) * 32), id_n8, k != 0);
        if (variant == 10 || variant == 11) {      // paired-row tile: 2 row tiles x 8 k-steps of 128x128x16
          const uint32_t id_128 = ptx::umma_idesc_f16_f32(128, 128);
          for (int k = 0; k < 8; ++k)
            for (int rt = 0; rt < 2; ++rt)
              ptx::umma_bf16_ss(tm + rt * 128, ptx::umma_desc_sw128_kmajor(a_addr + rt * 16384 + (k & 3) * 32),
                                ptx::umma_desc_sw128_kmajor(b_addr + (k & 3) * 32), id_128, k != 0);
          continue;
        }
        if (variant == 9)                          // 16 x TS N = 8, independent accumulators (no D dependency)
          for (int k = 0; k < 16; ++k)
            ptx::umma_f16_ts(tm + 384 + k * 8, tm + 256 + k * 8, ptx::umma_desc_sw128_kmajor(v_addr + (k & 3) * 32), id_n8, 0);
      }
      ptx::umma_commit(&bar);
      ptx::mbar_wait(&bar, phase);
      phase ^= 1;
      return clock64() - t0;
    };
    for (int v = 0; v < 12; ++v) {
      run(v);                                      // warm
      const long long c = run(v);
      if (blockIdx.x == 0) out[v] = c;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { ptx::tc_fence_after(); ptx::tmem_dealloc(tm, 512); }
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 16 * sizeof(long long));
  const int reps = 256;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  probe<<<148, 128, 100 * 1024>>>(out, reps);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  const char* names[12] = {"8 x SS 128x256x16 (one similarity tile, K=128)",
                           "16 x TS 128x8x16",
                           "16 x TS 128x16x16",
                           "16 x SS 128x8x16",
                           "tile + 16 x TS N=8",
                           "tile + 16 x TS N=16",
                           "tile + 16 x SS N=8",
                           "16 x TS 128x32x16",
                           "4 x TS 128x8x16",
                           "16 x TS 128x8x16 independent accumulators",
                           "16 x SS 128x128x16 (two row tiles x K=128, B shared)",
                           "16 x SS 128x128x16 (repeat)"};
  for (int v = 0; v < 12; ++v) printf("%-48s %8.1f cycles per repetition\n", names[v], double(out[v]) / reps);
  return 0;
}
