// Micro-benchmark (not product code): the SOFT epilogue's per-chunk arithmetic (scale, max tree, stash, 2^x, sum,
// fp16 pack) on register data, 16 warps per SM, no TMEM / barriers / tensor pipe.  Separates "the math is slow"
// from "the hand-offs are slow".  MODE bit 0: MUFU on; bit 1: max tree + stash on; bit 2: scale LDS on; bit 3: cvt on
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/epilogue_probe.bin tools/epilogue_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../geometric-aware-dense-matching_b200/csrc/ptx.cuh"
using namespace gadm;

template <int MODE>
__global__ void __launch_bounds__(576, 1) probe(long long* out, float* sink_out, int iters, float g, float mref) {
  __shared__ __align__(16) float scales[256];
  __shared__ __align__(16) float stash[512 * 8];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) scales[i] = 1.0f + i * 1e-3f;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  if ((MODE & (64 | 128)) && threadIdx.x < 32) { ptx::tmem_alloc(&tmem_slot, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); }
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tbase = (MODE & (64 | 128)) ? tmem_slot + (uint32_t(((threadIdx.x >> 5) & 3) * 32) << 16) + ((threadIdx.x >> 7) & 3) * 64 : 0;
  if (threadIdx.x >= 512) {
    // MODE bit 4: the two producer warps of the real kernel, parked on an mbarrier with a suspend-time hint
    // (bit 5: busy polling instead) until warp 0 is done
    if ((MODE & 16) && (threadIdx.x & 31) == 0) ptx::mbar_wait_sleep(&bar, 0);
    if ((MODE & 32) && (threadIdx.x & 31) == 0) ptx::mbar_wait(&bar, 0);
    return;
  }
  const uint32_t sc = ptx::smem_u32(scales) + ((threadIdx.x >> 7) * 64) * 4;
  const uint32_t stash_addr = ptx::smem_u32(stash) + threadIdx.x * 16;
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(0.01f * float((threadIdx.x * 37 + i * 11) & 63) - 0.3f);
  float vmax = -1e30f; int vgrp = 0;
  uint64_t l2a = 0, l2b = 0;
  uint32_t acc = 0;
  uint32_t rn[32];
  if (MODE & 256) ptx::tmem_ld_32x32(tbase, rn);
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if ((MODE & 64) && !(MODE & 256)) {           // fresh accumulators from TMEM, as in the real kernel
      ptx::tmem_ld_32x32(tbase + (it & 1) * 256 + ((it >> 1) & 1) * 32, r);
      ptx::tmem_ld_wait();
    }
    if (MODE & 256) {          // software-pipelined: wait for the chunk requested one iteration ago, request the next
      ptx::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) r[i] = rn[i];
      ptx::tmem_ld_32x32(tbase + (it & 1) * 256 + ((it >> 1) & 1) * 32, rn);
    }
    uint64_t v[16];
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
      float4 cm;
      if (MODE & 512) cm = ptx::lds128_free(sc + j4 * 16); else if (MODE & 4) cm = ptx::lds128(sc + j4 * 16); else cm = make_float4(g, g, g, g);
      v[j4 * 2 + 0] = ptx::fmul2(ptx::pack2(r[j4 * 4 + 0], r[j4 * 4 + 1]), ptx::pack2f(cm.x, cm.y));
      v[j4 * 2 + 1] = ptx::fmul2(ptx::pack2(r[j4 * 4 + 2], r[j4 * 4 + 3]), ptx::pack2f(cm.z, cm.w));
    }
    if (MODE & 2) {
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
      }
    }
    const uint64_t g2 = ptx::pack2f(g, g), nm2 = ptx::pack2f(-mref, -mref);
    uint32_t ph[16];
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
      uint64_t p01 = ptx::ffma2(v[j4 * 2 + 0], g2, nm2), p23 = ptx::ffma2(v[j4 * 2 + 1], g2, nm2);
      if (MODE & 1) { p01 = ptx::ex2_2(p01); p23 = ptx::ex2_2(p23); }
      l2a = ptx::fadd2(l2a, p01);
      l2b = ptx::fadd2(l2b, p23);
      float e0, e1, e2, e3;
      ptx::unpack2f(p01, e0, e1);
      ptx::unpack2f(p23, e2, e3);
      if (MODE & 8) { ph[j4 * 2 + 0] = ptx::cvt_f16x2(e1, e0); ph[j4 * 2 + 1] = ptx::cvt_f16x2(e3, e2); }
      else { ph[j4 * 2 + 0] = __float_as_uint(e1); ph[j4 * 2 + 1] = __float_as_uint(e3); }
    }
    if (MODE & 128) ptx::tmem_st_32x16(tbase + (it & 1) * 256 + ((it >> 1) & 1) * 16, ph);
#pragma unroll
    for (int i = 0; i < 16; ++i) acc ^= ph[i];
    // make the next iteration's inputs depend on this one (cheaply), as fresh accumulators would
#pragma unroll
    for (int i = 0; i < 32; i += 8) r[i] ^= (acc & 1u);
  }
  const long long t1 = clock64();
  float e, o;
  ptx::unpack2f(ptx::fadd2(l2a, l2b), e, o);
  sink_out[blockIdx.x * 512 + threadIdx.x] = e + o + vmax + float(vgrp) + float(acc & 7);
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  __syncwarp();
  if (threadIdx.x == 0) ptx::mbar_arrive(&bar);
  if (MODE & (64 | 128)) {
    if (MODE & 128) ptx::tmem_st_wait();
    ptx::tc_fence_before();
    asm volatile("bar.sync 1, 512;" ::: "memory");
    if (threadIdx.x < 32) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_slot, 512); }
  }
}

template <int MODE>
void run(long long* out, float* sink, const char* name, float mref = 0.1f) {
  const int iters = 4000;
  for (int rep = 0; rep < 2; ++rep) {
    probe<MODE><<<148, 576>>>(out, sink, iters, 0.7f, mref);
    cudaDeviceSynchronize();
  }
  printf("%-44s %7.1f cycles per 32-column chunk round (4 warps per scheduler)\n", name, double(out[0]) / iters);
}

int main() {
  long long* out; float* sink;
  cudaMallocManaged(&out, 64);
  cudaMalloc(&sink, 148 * 512 * 4);
  run<15>(out, sink, "everything");
  run<13>(out, sink, "no stash");
  run<13 + 64>(out, sink, "no stash + tcgen05.ld x32 per chunk");
  run<13 + 128>(out, sink, "no stash + tcgen05.st x16 per chunk");
  run<13 + 64 + 128>(out, sink, "no stash + tcgen05.ld + tcgen05.st");
  run<64>(out, sink, "ffma/fadd only + tcgen05.ld");
  run<13 + 64 + 256>(out, sink, "no stash + PREFETCHED tcgen05.ld");
  run<13 + 64 + 128 + 256>(out, sink, "no stash + PREFETCHED tcgen05.ld + tcgen05.st");
  run<15 + 64 + 128 + 256>(out, sink, "everything + PREFETCHED ld + st");
  run<15 + 64 + 128>(out, sink, "everything + ld + st");
  run<15 + 512>(out, sink, "everything, non-volatile scale LDS");
  run<15 + 64 + 128 + 512>(out, sink, "everything + ld + st, non-volatile scale LDS");
  run<13 + 64 + 512>(out, sink, "no stash + ld, non-volatile scale LDS");
  run<15>(out, sink, "everything, p ~ 2^-20 (fp16 subnormal)", 20.f);
  run<15>(out, sink, "everything, p ~ 2^-40 (fp16 zero)", 40.f);
  run<7>(out, sink, "no fp16 pack, p ~ 2^-20", 20.f);
  run<1>(out, sink, "MUFU + ffma/fadd only, p ~ 2^-20", 20.f);
  run<14>(out, sink, "no MUFU");
  run<13>(out, sink, "no max tree / stash");
  run<11>(out, sink, "no scale LDS");
  run<7>(out, sink, "no fp16 pack");
  run<1>(out, sink, "MUFU + ffma/fadd only");
  run<0>(out, sink, "ffma/fadd only");
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
