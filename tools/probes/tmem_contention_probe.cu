// Micro-benchmark (not product code): does tcgen05.ld slow down while the tensor pipe accumulates into the OTHER
// half of TMEM?  Warps 0..15 drain accumulator 1 (columns 256..511) in a loop, as the matcher epilogue does, while
// one thread keeps issuing 128x256x16 MMAs into accumulator 0.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/tmem_contention_probe.bin tools/tmem_contention_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../geometric-aware-dense-matching_b200/csrc/ptx.cuh"

using namespace gadm;

__global__ void __launch_bounds__(576, 1) probe(long long* out, int tiles, int ld_iters, int with_mma, int with_ld) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ long long t_ld[16];
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); }
  if (warp == 17) { ptx::tmem_alloc(&tmem_slot, 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tm = tmem_slot;
  if (warp == 17) {
    if ((threadIdx.x & 31) == 0 && with_mma) {
      const uint32_t a_addr = ptx::smem_u32(smem), b_addr = ptx::smem_u32(smem + 16 * 1024);
      const uint32_t idesc = ptx::umma_idesc_f16_f32(128, 256);
      const long long t0 = clock64();
      for (int t = 0; t < tiles; ++t)
        for (int k = 0; k < 8; ++k)
          ptx::umma_bf16_ss(tm, ptx::umma_desc_sw128_kmajor(a_addr + (k & 3) * 32),
                            ptx::umma_desc_sw128_kmajor(b_addr + (k & 3) * 32), idesc, k != 0);
      ptx::umma_commit(&bar);
      ptx::mbar_wait(&bar, 0);
      if (blockIdx.x == 0) out[0] = clock64() - t0;
    }
  } else if (warp < 16 && with_ld) {
    const uint32_t base = tm + (uint32_t((warp & 3) * 32) << 16) + 256 + (warp >> 2) * 64;
    uint32_t sink = 0;
    const long long t0 = clock64();
    for (int i = 0; i < ld_iters; ++i) {
      uint32_t r[32], s[32];
      ptx::tmem_ld_32x32(base, r);
      ptx::tmem_ld_32x32(base + 32, s);
      ptx::tmem_ld_wait();
      sink += r[0] ^ s[31];
    }
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) t_ld[warp] = t1 - t0;
    if (sink == 0x12345u) out[7] = sink;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0 && with_ld) {
    long long m = 0;
    for (int w = 0; w < 16; ++w) m = max(m, t_ld[w]);
    out[1] = m;
  }
  if (warp == 17) { ptx::tc_fence_after(); ptx::tmem_dealloc(tm, 512); }
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 16 * sizeof(long long));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  const int tiles = 512, ld_iters = 2048;
  for (int cfg = 0; cfg < 3; ++cfg) {
    const int with_mma = cfg != 1, with_ld = cfg != 0;
    for (int rep = 0; rep < 2; ++rep) {
      out[0] = out[1] = 0;
      probe<<<148, 576, 80 * 1024>>>(out, tiles, ld_iters, with_mma, with_ld);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    }
    printf("%-22s MMA: %7.1f cycles per 128x256x128 tile   LDTM: %7.1f cycles per 64-column slice per warp (16 warps)\n",
           cfg == 0 ? "MMA only" : cfg == 1 ? "tcgen05.ld only" : "both at once", double(out[0]) / tiles,
           double(out[1]) / ld_iters);
  }
  return 0;
}
