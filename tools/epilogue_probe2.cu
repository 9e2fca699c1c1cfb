// Micro-benchmark (not product code): the matcher's SOFT / ARGMAX epilogue loop as the kernel has it (tcgen05.ld of a
// 32-column chunk, scales, max tree + stash, 2^x, sums of p and p * xyz from broadcast planes) on 16 warps per SM,
// without barriers / TMA / UMMA.  Template switches try restructurings in seconds; the winner is ported to
// match_sm100.cu.   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/epilogue_probe2.bin tools/epilogue_probe2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../geometric-aware-dense-matching_b200/csrc/ptx.cuh"
using namespace gadm;

// SOFT: 1 = soft, 0 = argmax.  FREE: non-volatile LDS.  EARLY: scale LDS issued before the tcgen05.ld wait.
// BOTH: two chunks requested per wait (argmax style).  NOSTASHCLOB: stash stores without a "memory" clobber.
template <int SOFT, int FREE, int EARLY, int BOTH>
__global__ void __launch_bounds__(576, 1) probe(long long* out, float* sink_out, int iters, float g, float mref) {
  __shared__ __align__(16) float aux[4 * 256];
  __shared__ __align__(16) float stash[512 * 8];
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) aux[i] = 1.0f + (i & 255) * 1e-3f;
  if (threadIdx.x < 32) { ptx::tmem_alloc(&tmem_slot, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (threadIdx.x >= 512) return;
  const int warp = threadIdx.x >> 5;
  const uint32_t tbase = tmem_slot + (uint32_t((warp & 3) * 32) << 16) + (warp >> 2) * 64;
  const uint32_t sc0 = ptx::smem_u32(aux) + ((warp >> 2) * 64) * 4;
  const uint32_t stash_addr = ptx::smem_u32(stash) + threadIdx.x * 16;
  float vmax = -1e30f; int vgrp = 0;
  uint64_t l2a = 0, l2b = 0, ax2a = 0, ax2b = 0, ay2a = 0, ay2b = 0, az2a = 0, az2b = 0;
  auto lds = [&](uint32_t a) { return FREE ? ptx::lds128_free(a) : ptx::lds128(a); };

  auto process = [&](uint32_t (&r)[32], const float4 (&cm)[8], uint32_t sc, int it) {
    uint64_t v[16];
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
      v[j4 * 2 + 0] = ptx::fmul2(ptx::pack2(r[j4 * 4 + 0], r[j4 * 4 + 1]), ptx::pack2f(cm[j4].x, cm[j4].y));
      v[j4 * 2 + 1] = ptx::fmul2(ptx::pack2(r[j4 * 4 + 2], r[j4 * 4 + 3]), ptx::pack2f(cm[j4].z, cm[j4].w));
    }
    float cmx = -1e30f;
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      float f[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) ptx::unpack2f(v[h * 4 + j], f[2 * j], f[2 * j + 1]);
      const float a0 = ptx::fmax3(f[0], f[1], f[2]), a1 = ptx::fmax3(f[3], f[4], f[5]);
      const float gm = ptx::fmax3(a0, a1, fmaxf(f[6], f[7]));
      const bool up = gm > vmax;
      ptx::sts_stash8(up, stash_addr, v[h * 4 + 0], v[h * 4 + 1], v[h * 4 + 2], v[h * 4 + 3]);
      vgrp = up ? it * 32 + h * 8 : vgrp;
      vmax = up ? gm : vmax;
      cmx = fmaxf(cmx, gm);
    }
    if (SOFT) {
      const float m = mref + cmx * 1e-9f;
      const uint64_t g2 = ptx::pack2f(g, g), nm2 = ptx::pack2f(-m, -m);
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const float4 X = lds(sc + 1024 + j4 * 16), Y = lds(sc + 2048 + j4 * 16), Z = lds(sc + 3072 + j4 * 16);
        const uint64_t p01 = ptx::ex2_2(ptx::ffma2(v[j4 * 2 + 0], g2, nm2));
        const uint64_t p23 = ptx::ex2_2(ptx::ffma2(v[j4 * 2 + 1], g2, nm2));
        l2a = ptx::fadd2(l2a, p01); l2b = ptx::fadd2(l2b, p23);
        ax2a = ptx::ffma2(p01, ptx::pack2f(X.x, X.y), ax2a); ax2b = ptx::ffma2(p23, ptx::pack2f(X.z, X.w), ax2b);
        ay2a = ptx::ffma2(p01, ptx::pack2f(Y.x, Y.y), ay2a); ay2b = ptx::ffma2(p23, ptx::pack2f(Y.z, Y.w), ay2b);
        az2a = ptx::ffma2(p01, ptx::pack2f(Z.x, Z.y), az2a); az2b = ptx::ffma2(p23, ptx::pack2f(Z.z, Z.w), az2b);
      }
    }
  };

  const long long t0 = clock64();
  for (int it = 0; it < iters; it += 2) {      // one "tile": two 32-column chunks
    const uint32_t s_tmem = tbase + (it & 2) * 128;
    const uint32_t sc = ptx::opaque(sc0);
    uint32_t ra[32], rb[32];
    float4 ca[8], cb[8];
    ptx::tmem_ld_32x32(s_tmem, ra);
    if (BOTH) ptx::tmem_ld_32x32(s_tmem + 32, rb);
    if (EARLY) {
#pragma unroll
      for (int j = 0; j < 8; ++j) ca[j] = lds(sc + j * 16);
      if (BOTH) {
#pragma unroll
        for (int j = 0; j < 8; ++j) cb[j] = lds(sc + 128 + j * 16);
      }
    }
    ptx::tmem_ld_wait();
    if (!EARLY) {
#pragma unroll
      for (int j = 0; j < 8; ++j) ca[j] = lds(sc + j * 16);
    }
    process(ra, ca, sc, it);
    if (!BOTH) {
      ptx::tmem_ld_32x32(s_tmem + 32, rb);
      if (EARLY) {
#pragma unroll
        for (int j = 0; j < 8; ++j) cb[j] = lds(sc + 128 + j * 16);
      }
      ptx::tmem_ld_wait();
    }
    if (!EARLY || (false)) {
      if (!EARLY) {
#pragma unroll
        for (int j = 0; j < 8; ++j) cb[j] = lds(sc + 128 + j * 16);
      }
    }
    process(rb, cb, sc + 128, it + 1);
  }
  const long long t1 = clock64();
  float e, o, acc = 0.f;
  ptx::unpack2f(ptx::fadd2(ptx::fadd2(l2a, l2b), ptx::fadd2(ptx::fadd2(ax2a, ax2b), ptx::fadd2(ptx::fadd2(ay2a, ay2b), ptx::fadd2(az2a, az2b)))), e, o);
  acc = e + o;
  sink_out[blockIdx.x * 512 + threadIdx.x] = acc + vmax + float(vgrp);
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  ptx::tc_fence_before();
  asm volatile("bar.sync 1, 512;" ::: "memory");
  if (threadIdx.x < 32) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_slot, 512); }
}

template <int SOFT, int FREE, int EARLY, int BOTH>
void run(long long* out, float* sink, const char* name) {
  const int iters = 4000;
  for (int rep = 0; rep < 2; ++rep) {
    probe<SOFT, FREE, EARLY, BOTH><<<148, 576>>>(out, sink, iters, 0.7f, 0.1f);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); exit(1); }
  }
  printf("%-64s %7.1f cycles per 256-column tile (16 warps)\n", name, 2.0 * double(out[0]) / iters);
}

int main() {
  long long* out; float* sink;
  cudaMallocManaged(&out, 64);
  cudaMalloc(&sink, 148 * 512 * 4);
  run<1, 0, 0, 0>(out, sink, "SOFT as shipped (volatile LDS, one chunk per wait)");
  run<1, 1, 0, 0>(out, sink, "SOFT non-volatile LDS");
  run<1, 1, 1, 0>(out, sink, "SOFT non-volatile LDS, scales before the ld wait");
  run<1, 1, 1, 1>(out, sink, "SOFT non-volatile LDS, scales early, both chunks per wait");
  run<0, 0, 0, 1>(out, sink, "ARGMAX as shipped (volatile LDS, both chunks per wait)");
  run<0, 1, 0, 1>(out, sink, "ARGMAX non-volatile LDS");
  run<0, 1, 1, 1>(out, sink, "ARGMAX non-volatile LDS, scales before the ld wait");
  run<0, 1, 1, 0>(out, sink, "ARGMAX non-volatile LDS, scales early, one chunk per wait");
  return 0;
}
