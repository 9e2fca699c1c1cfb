// Micro-benchmark (not product code): do MUFU.EX2 (XU pipe), packed f32x2 arithmetic (FMA pipe) and FMNMX3 (ALU pipe)
// overlap on an SM sub-partition when 4 warps per scheduler issue them interleaved, as the SOFT epilogue does
// (per 32 scores: 32 MUFU.EX2, 96 packed f32x2, 16 FMNMX)?  Every slot of the unrolled loop body issues M x MUFU.EX2,
// F x FFMA2 (or scalar FFMA when kScalar) and A x FMNMX3 on independent registers.  Prints cycles per slot per
// scheduler (16 warps per SM, one CTA per SM, clock64 on every warp, averaged).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/pipe_mix_probe tools/probes/pipe_mix_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int SLOTS = 32;

template <int M, int F, int A, bool kScalar, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) probe(long long* cyc, float* sink, int iters) {
  float x[8], m[4];
  uint64_t acc[8];
  float sacc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    x[i] = -0.01f * float((threadIdx.x + i) & 31);
    acc[i] = 0;
    sacc[i] = 0.f;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) m[i] = float(i);
  uint64_t p, c;
  asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(x[0]), "f"(x[1]));
  asm volatile("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(x[2]), "f"(x[3]));
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
#pragma unroll
      for (int k = 0; k < M; ++k) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[(s * M + k) & 7]));   // 8 independent chains per thread
      }
#pragma unroll
      for (int k = 0; k < F; ++k) {
        if (kScalar) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(sacc[(s * F + k) & 7]) : "f"(m[k & 3]), "f"(m[(k + 1) & 3]));
        else asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[(s * F + k) & 7]) : "l"(p), "l"(c));
      }
#pragma unroll
      for (int k = 0; k < A; ++k)
        asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(m[(s * A + k) & 3]) : "f"(sacc[(s + k) & 7]), "f"(sacc[(s + k + 3) & 7]));
    }
  }
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * WARPS + (threadIdx.x >> 5)] = t1 - t0;
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float lo, hi;
    asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i]));
    r += lo + hi + sacc[i] + x[i];
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) r += m[i];
  if (r == 123.456f) sink[0] = r;
}

template <int M, int F, int A, bool kScalar, int WARPS = 16>
void run(const char* name, long long* d_cyc, float* d_sink) {
  const int iters = 2000, ctas = 148;
  probe<M, F, A, kScalar, WARPS><<<ctas, WARPS * 32>>>(d_cyc, d_sink, 10);
  probe<M, F, A, kScalar, WARPS><<<ctas, WARPS * 32>>>(d_cyc, d_sink, iters);
  cudaDeviceSynchronize();
  static long long h[148 * 32];
  cudaMemcpy(h, d_cyc, sizeof(long long) * ctas * WARPS, cudaMemcpyDeviceToHost);
  double s = 0;
  for (int i = 0; i < ctas * WARPS; ++i) s += double(h[i]);
  s /= double(ctas * WARPS);
  const double per_slot = s / (double(iters) * SLOTS);
  // WARPS / 4 warps per scheduler each run a slot in `per_slot` cycles
  printf("%-44s warps/sched %d  cycles per slot (per warp) %7.2f   per warp-instruction on the scheduler %6.2f\n", name,
         WARPS / 4, per_slot, per_slot / double((M + F + A) * (WARPS / 4)));
}

int main() {
  long long* d_cyc; float* d_sink;
  cudaMalloc(&d_cyc, sizeof(long long) * 148 * 32);
  cudaMalloc(&d_sink, 4);
  run<1, 0, 0, false>("MUFU.EX2 only", d_cyc, d_sink);
  run<0, 1, 0, false>("FFMA2 only", d_cyc, d_sink);
  run<0, 1, 0, true>("FFMA (scalar) only", d_cyc, d_sink);
  run<0, 0, 1, false>("FMNMX3 only", d_cyc, d_sink);
  run<1, 3, 0, false>("1 MUFU : 3 FFMA2  (SOFT ratio)", d_cyc, d_sink);
  run<1, 3, 1, false>("1 MUFU : 3 FFMA2 : 1 FMNMX3", d_cyc, d_sink);
  run<1, 2, 0, false>("1 MUFU : 2 FFMA2", d_cyc, d_sink);
  run<1, 1, 0, false>("1 MUFU : 1 FFMA2", d_cyc, d_sink);
  run<1, 6, 0, true>("1 MUFU : 6 FFMA scalar", d_cyc, d_sink);
  run<0, 3, 1, false>("3 FFMA2 : 1 FMNMX3", d_cyc, d_sink);
  run<1, 0, 0, false, 4>("MUFU.EX2 only, 1 warp per scheduler", d_cyc, d_sink);
  run<1, 3, 0, false, 4>("1 MUFU : 3 FFMA2, 1 warp per scheduler", d_cyc, d_sink);
  run<1, 3, 0, false, 32>("1 MUFU : 3 FFMA2, 8 warps per scheduler", d_cyc, d_sink);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
