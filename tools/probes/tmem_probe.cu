// Micro-benchmark (not product code): TMEM -> register read bandwidth per SM as a function of the number of warps
// and of the tcgen05.ld width.  The matcher epilogue must drain a 128 x 256 fp32 accumulator (128 KB) per tile.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/tmem_probe.bin tools/tmem_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../geometric-aware-dense-matching_b200/csrc/ptx.cuh"

using namespace gadm;

__device__ __forceinline__ void ld_x64(uint32_t taddr, uint32_t (&r)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
      "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
      "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]),
        "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]),
        "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]),
        "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]),
        "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(taddr)
      : "memory");
}

// mode 0: x32 loads, wait after each; 1: two x32 loads per wait; 2: x64 loads; 3: x16 loads, four per wait
__global__ void __launch_bounds__(512, 1) probe(long long* out, int iters, int mode) {
  __shared__ uint32_t tmem_slot;
  __shared__ long long t_start[16], t_end[16];
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { ptx::tmem_alloc(&tmem_slot, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tm = tmem_slot;
  const uint32_t base = tm + (uint32_t((warp & 3) * 32) << 16) + (warp >> 2) * 64;
  uint32_t sink = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (mode == 0) {
      uint32_t r[32];
      ptx::tmem_ld_32x32(base + (i & 1) * 256, r);
      ptx::tmem_ld_wait();
      sink += r[0] ^ r[31];
      ptx::tmem_ld_32x32(base + (i & 1) * 256 + 32, r);
      ptx::tmem_ld_wait();
      sink += r[0] ^ r[31];
    } else if (mode == 1) {
      uint32_t r[32], s[32];
      ptx::tmem_ld_32x32(base + (i & 1) * 256, r);
      ptx::tmem_ld_32x32(base + (i & 1) * 256 + 32, s);
      ptx::tmem_ld_wait();
      sink += r[0] ^ s[31];
    } else if (mode == 2) {
      uint32_t r[64];
      ld_x64(base + (i & 1) * 256, r);
      ptx::tmem_ld_wait();
      sink += r[0] ^ r[63];
    } else {
      uint32_t a[16], b[16], c[16], d[16];
      ptx::tmem_ld_32x16(base + (i & 1) * 256, a);
      ptx::tmem_ld_32x16(base + (i & 1) * 256 + 16, b);
      ptx::tmem_ld_32x16(base + (i & 1) * 256 + 32, c);
      ptx::tmem_ld_32x16(base + (i & 1) * 256 + 48, d);
      ptx::tmem_ld_wait();
      sink += a[0] ^ b[15] ^ c[3] ^ d[7];
    }
  }
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) { t_start[warp] = t0; t_end[warp] = t1; }
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    long long lo = t_start[0], hi = t_end[0];
    for (int w = 1; w < int(blockDim.x >> 5); ++w) { lo = min(lo, t_start[w]); hi = max(hi, t_end[w]); }
    out[0] = hi - lo;
  }
  if (sink == 0x12345u) out[1] = sink;
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tm, 512); }
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 16 * sizeof(long long));
  const int iters = 2000;
  const char* names[4] = {"x32, wait each", "2 x x32 per wait", "x64 per wait", "4 x x16 per wait"};
  for (int mode = 0; mode < 4; ++mode)
    for (int warps = 4; warps <= 16; warps *= 2) {
      probe<<<148, warps * 32, 0>>>(out, iters, mode);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      probe<<<148, warps * 32, 0>>>(out, iters, mode);
      cudaDeviceSynchronize();
      const double bytes = double(iters) * warps * 32 * 64 * 4;   // every iteration: 64 columns x 32 lanes x 4 B per warp
      printf("%-18s %2d warps: %8.1f cycles / iteration, %6.1f B/clk/SM  (128 KB tile in %6.0f cycles)\n", names[mode],
             warps, double(out[0]) / iters, bytes / double(out[0]), 131072.0 / (bytes / double(out[0])));
    }
  return 0;
}
