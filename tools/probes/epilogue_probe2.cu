// Micro-benchmark (not product code): the matcher's SOFT / ARGMAX epilogue loop as the kernel has it (tcgen05.ld of a
// 32-column chunk, scales, max tree + stash, 2^x, sums of p and p * xyz from broadcast planes) on 16 warps per SM,
// without barriers / TMA / UMMA.  Template switches try restructurings in seconds; the winner is ported to
// match_sm100.cu.   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/epilogue_probe2.bin tools/epilogue_probe2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../geometric-aware-dense-matching_b200/csrc/ptx.cuh"
using namespace gadm;

// SOFT: 1 = soft, 0 = argmax.  FREE: non-volatile LDS.  EARLY: scale LDS issued before the tcgen05.ld wait.
// BOTH: two chunks requested per wait (argmax style).  NOSTASHCLOB: stash stores without a "memory" clobber.
template <int SOFT, int FREE, int EARLY, int BOTH, int ABL = 0>
__global__ void __launch_bounds__(576, 1) probe(long long* out, float* sink_out, int iters, float g, float mref) {
  __shared__ __align__(16) float aux[4 * 256];
  __shared__ __align__(16) float stash[512 * 8];
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) aux[i] = 1.0f + (i & 255) * 1e-3f;
  if (threadIdx.x < 32) { ptx::tmem_alloc(&tmem_slot, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  extern __shared__ uint8_t dyn_raw[];
  __shared__ volatile int stop_flag;
  __shared__ uint64_t mbar;
  if (ABL & 32) {
    if (threadIdx.x == 0) { stop_flag = 0; ptx::mbar_init(&mbar, 1); ptx::fence_mbar_init(); }
    uint8_t* dyn = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dyn_raw) + 1023) & ~uintptr_t(1023));
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(dyn)[i] = 0x3c003c00u;
    ptx::fence_proxy_async();
    asm volatile("bar.sync 2, 576;" ::: "memory");
    if (threadIdx.x == 544) {            // warp 17, lane 0: the UMMA issuer of the real kernel
      const uint32_t a_addr = ptx::smem_u32(dyn), b_addr = ptx::smem_u32(dyn + 16 * 1024);
      const uint32_t idesc = ptx::umma_idesc_f16_f32(128, 256);
      uint32_t ph = 0; long long n = 0;
      while (!stop_flag) {
        for (int k = 0; k < 8; ++k)       // accumulate into columns 0..255 of lanes the probe does not read back
          ptx::umma_bf16_ss(tmem_slot + 256 * ((n & 1) ^ 1) * 0 + 0, ptx::umma_desc_sw128_kmajor(a_addr + (k & 3) * 32),
                            ptx::umma_desc_sw128_kmajor(b_addr + (k & 3) * 32), idesc, 1);
        ptx::umma_commit(&mbar);
        ptx::mbar_wait_sleep(&mbar, ph); ph ^= 1; ++n;
      }
      if (blockIdx.x == 0) out[2] = n;
    }
  }
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
      if (!(ABL & 8)) ptx::sts_stash8(up, stash_addr, v[h * 4 + 0], v[h * 4 + 1], v[h * 4 + 2], v[h * 4 + 3]);
      vgrp = up ? it * 32 + h * 8 : vgrp;
      vmax = up ? gm : vmax;
      cmx = fmaxf(cmx, gm);
    }
    if (SOFT) {
      const float m = mref + cmx * 1e-9f;
      const uint64_t g2 = ptx::pack2f(g, g), nm2 = ptx::pack2f(-m, -m);
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        float4 X, Y, Z;
        if (ABL & 1) { X = make_float4(g, m, g, m); Y = make_float4(m, g, m, g); Z = make_float4(g, g, m, m); }
        else { X = lds(sc + 1024 + j4 * 16); Y = lds(sc + 2048 + j4 * 16); Z = lds(sc + 3072 + j4 * 16); }
        uint64_t p01 = ptx::ffma2(v[j4 * 2 + 0], g2, nm2), p23 = ptx::ffma2(v[j4 * 2 + 1], g2, nm2);
        if (!(ABL & 4)) { p01 = ptx::ex2_2(p01); p23 = ptx::ex2_2(p23); }
        l2a = ptx::fadd2(l2a, p01); l2b = ptx::fadd2(l2b, p23);
        if (!(ABL & 2)) {
        ax2a = ptx::ffma2(p01, ptx::pack2f(X.x, X.y), ax2a); ax2b = ptx::ffma2(p23, ptx::pack2f(X.z, X.w), ax2b);
        ay2a = ptx::ffma2(p01, ptx::pack2f(Y.x, Y.y), ay2a); ay2b = ptx::ffma2(p23, ptx::pack2f(Y.z, Y.w), ay2b);
        az2a = ptx::ffma2(p01, ptx::pack2f(Z.x, Z.y), az2a); az2b = ptx::ffma2(p23, ptx::pack2f(Z.z, Z.w), az2b);
        }
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
      for (int j = 0; j < 8; ++j) ca[j] = (ABL & 16) ? make_float4(g, g, g, g) : lds(sc + j * 16);
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
        for (int j = 0; j < 8; ++j) cb[j] = (ABL & 16) ? make_float4(g, g, g, g) : lds(sc + 128 + j * 16);
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
  if (ABL & 32) {
    if (threadIdx.x == 0) stop_flag = 1;
    // give the issuer time to see the flag and drain before TMEM goes away
    __nanosleep(20000);
  }
  if (threadIdx.x < 32) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_slot, 512); }
}

// Paired rows: every thread owns the same lane of TWO row tiles; one broadcast LDS.128 of a per-column constant now
// serves 2 x 4 scores.  W = columns per step (32 or 16).
template <int SOFT, int W>
__global__ void __launch_bounds__(576, 1) probe_pair(long long* out, float* sink_out, int iters, float g, float mref) {
  __shared__ __align__(16) float aux[4 * 256];
  __shared__ __align__(16) float stash[2 * 512 * 8];
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) aux[i] = 1.0f + (i & 255) * 1e-3f;
  if (threadIdx.x < 32) { ptx::tmem_alloc(&tmem_slot, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (threadIdx.x >= 512) return;
  const int warp = threadIdx.x >> 5;
  const uint32_t tbase = tmem_slot + (uint32_t((warp & 3) * 32) << 16) + (warp >> 2) * 32;
  const uint32_t sc0 = ptx::smem_u32(aux) + ((warp >> 2) * 32) * 4;
  const uint32_t stash_addr = ptx::smem_u32(stash) + threadIdx.x * 16;
  float vmax[2] = {-1e30f, -1e30f}; int vgrp[2] = {0, 0};
  uint64_t l2[2] = {0, 0}, ax2[2] = {0, 0}, ay2[2] = {0, 0}, az2[2] = {0, 0};
  constexpr int NCH = 32 / W;
  const long long t0 = clock64();
  for (int it = 0; it < iters; it += 2) {      // one step = 2 rows x 32 columns = 64 scores per thread (as above)
    const uint32_t s_tmem = tbase + (it & 2) * 128;
    const uint32_t sc = ptx::opaque(sc0);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      uint32_t r0[W], r1[W];
      if (W == 32) { ptx::tmem_ld_32x32(s_tmem, reinterpret_cast<uint32_t(&)[32]>(r0)); ptx::tmem_ld_32x32(s_tmem + 128, reinterpret_cast<uint32_t(&)[32]>(r1)); }
      else { ptx::tmem_ld_32x16(s_tmem + c * 16, reinterpret_cast<uint32_t(&)[16]>(r0)); ptx::tmem_ld_32x16(s_tmem + 128 + c * 16, reinterpret_cast<uint32_t(&)[16]>(r1)); }
      ptx::tmem_ld_wait();
      uint64_t v0[W / 2], v1[W / 2];
#pragma unroll
      for (int j4 = 0; j4 < W / 4; ++j4) {
        const float4 cm = ptx::lds128(sc + c * W * 4 + j4 * 16);
        const uint64_t c01 = ptx::pack2f(cm.x, cm.y), c23 = ptx::pack2f(cm.z, cm.w);
        v0[j4 * 2 + 0] = ptx::fmul2(ptx::pack2(r0[j4 * 4 + 0], r0[j4 * 4 + 1]), c01);
        v0[j4 * 2 + 1] = ptx::fmul2(ptx::pack2(r0[j4 * 4 + 2], r0[j4 * 4 + 3]), c23);
        v1[j4 * 2 + 0] = ptx::fmul2(ptx::pack2(r1[j4 * 4 + 0], r1[j4 * 4 + 1]), c01);
        v1[j4 * 2 + 1] = ptx::fmul2(ptx::pack2(r1[j4 * 4 + 2], r1[j4 * 4 + 3]), c23);
      }
      float cmx[2] = {-1e30f, -1e30f};
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        uint64_t* v = rr ? v1 : v0;
#pragma unroll
        for (int h = 0; h < W / 8; ++h) {
          float f[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) ptx::unpack2f(v[h * 4 + j], f[2 * j], f[2 * j + 1]);
          const float a0 = ptx::fmax3(f[0], f[1], f[2]), a1 = ptx::fmax3(f[3], f[4], f[5]);
          const float gm = ptx::fmax3(a0, a1, fmaxf(f[6], f[7]));
          const bool up = gm > vmax[rr];
          ptx::sts_stash8(up, stash_addr + rr * 16384, v[h * 4 + 0], v[h * 4 + 1], v[h * 4 + 2], v[h * 4 + 3]);
          vgrp[rr] = up ? it * 32 + h * 8 : vgrp[rr];
          vmax[rr] = up ? gm : vmax[rr];
          cmx[rr] = fmaxf(cmx[rr], gm);
        }
      }
      if (SOFT) {
        const float m0 = mref + cmx[0] * 1e-9f, m1 = mref + cmx[1] * 1e-9f;
        const uint64_t g2 = ptx::pack2f(g, g), nm0 = ptx::pack2f(-m0, -m0), nm1 = ptx::pack2f(-m1, -m1);
#pragma unroll
        for (int j4 = 0; j4 < W / 4; ++j4) {
          const uint32_t a = sc + c * W * 4 + j4 * 16;
          const float4 X = ptx::lds128(a + 1024), Y = ptx::lds128(a + 2048), Z = ptx::lds128(a + 3072);
          const uint64_t X01 = ptx::pack2f(X.x, X.y), X23 = ptx::pack2f(X.z, X.w), Y01 = ptx::pack2f(Y.x, Y.y),
                         Y23 = ptx::pack2f(Y.z, Y.w), Z01 = ptx::pack2f(Z.x, Z.y), Z23 = ptx::pack2f(Z.z, Z.w);
          const uint64_t p0a = ptx::ex2_2(ptx::ffma2(v0[j4 * 2 + 0], g2, nm0)), p0b = ptx::ex2_2(ptx::ffma2(v0[j4 * 2 + 1], g2, nm0));
          const uint64_t p1a = ptx::ex2_2(ptx::ffma2(v1[j4 * 2 + 0], g2, nm1)), p1b = ptx::ex2_2(ptx::ffma2(v1[j4 * 2 + 1], g2, nm1));
          l2[0] = ptx::fadd2(l2[0], ptx::fadd2(p0a, p0b)); l2[1] = ptx::fadd2(l2[1], ptx::fadd2(p1a, p1b));
          ax2[0] = ptx::ffma2(p0b, X23, ptx::ffma2(p0a, X01, ax2[0])); ax2[1] = ptx::ffma2(p1b, X23, ptx::ffma2(p1a, X01, ax2[1]));
          ay2[0] = ptx::ffma2(p0b, Y23, ptx::ffma2(p0a, Y01, ay2[0])); ay2[1] = ptx::ffma2(p1b, Y23, ptx::ffma2(p1a, Y01, ay2[1]));
          az2[0] = ptx::ffma2(p0b, Z23, ptx::ffma2(p0a, Z01, az2[0])); az2[1] = ptx::ffma2(p1b, Z23, ptx::ffma2(p1a, Z01, az2[1]));
        }
      }
    }
  }
  const long long t1 = clock64();
  float e, o;
  ptx::unpack2f(ptx::fadd2(ptx::fadd2(l2[0], l2[1]), ptx::fadd2(ptx::fadd2(ax2[0], ax2[1]), ptx::fadd2(ptx::fadd2(ay2[0], ay2[1]), ptx::fadd2(az2[0], az2[1])))), e, o);
  sink_out[blockIdx.x * 512 + threadIdx.x] = e + o + vmax[0] + vmax[1] + float(vgrp[0] + vgrp[1]);
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  ptx::tc_fence_before();
  asm volatile("bar.sync 1, 512;" ::: "memory");
  if (threadIdx.x < 32) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_slot, 512); }
}

template <int SOFT, int W>
void run_pair(long long* out, float* sink, const char* name) {
  const int iters = 4000;
  for (int rep = 0; rep < 2; ++rep) {
    probe_pair<SOFT, W><<<148, 576>>>(out, sink, iters, 0.7f, 0.1f);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); exit(1); }
  }
  printf("%-64s %7.1f cycles per 128 x 256 scores (16 warps)\n", name, 2.0 * double(out[0]) / iters);
}

template <int SOFT, int FREE, int EARLY, int BOTH, int ABL = 0>
void run(long long* out, float* sink, const char* name) {
  const int iters = 4000;
  if (ABL & 32) cudaFuncSetAttribute(probe<SOFT, FREE, EARLY, BOTH, ABL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 50 * 1024);
  for (int rep = 0; rep < 2; ++rep) {
    probe<SOFT, FREE, EARLY, BOTH, ABL><<<148, 576, (ABL & 32) ? 50 * 1024 : 0>>>(out, sink, iters, 0.7f, 0.1f);
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
  run<1, 0, 0, 0, 1>(out, sink, "SOFT, xyz constants from registers (no plane LDS)");
  run<1, 0, 0, 0, 2>(out, sink, "SOFT, no p * xyz sums at all");
  run<1, 0, 0, 0, 4>(out, sink, "SOFT, no MUFU");
  run<1, 0, 0, 0, 8>(out, sink, "SOFT, no stash stores");
  run<1, 0, 0, 0, 16>(out, sink, "SOFT, no scale LDS");
  run<1, 0, 0, 0, 1 + 16>(out, sink, "SOFT, no LDS at all");
  run<1, 0, 0, 0, 1 + 8 + 16>(out, sink, "SOFT, no LDS, no stash");
  run<1, 0, 0, 0, 1 + 4 + 8 + 16>(out, sink, "SOFT, no LDS, no stash, no MUFU");
  run<1, 0, 0, 0, 32>(out, sink, "SOFT as shipped + a UMMA issuer streaming 128x256x16 MMAs from smem");
  printf("      (MMA tiles issued meanwhile: %lld)\n", out[2]);
  run<0, 0, 0, 1, 32>(out, sink, "ARGMAX as shipped + the UMMA issuer");
  printf("      (MMA tiles issued meanwhile: %lld)\n", out[2]);
  run<0, 0, 0, 1, 16>(out, sink, "ARGMAX, no scale LDS");
  run<0, 0, 0, 1, 16 + 32>(out, sink, "ARGMAX, no scale LDS + the UMMA issuer");
  printf("      (MMA tiles issued meanwhile: %lld)\n", out[2]);
  run<0, 0, 0, 1, 8 + 16>(out, sink, "ARGMAX, no scale LDS, no stash");
  run<0, 0, 0, 1, 8 + 16 + 32>(out, sink, "ARGMAX, no scale LDS, no stash + the UMMA issuer");
  printf("      (MMA tiles issued meanwhile: %lld)\n", out[2]);
  run<1, 0, 0, 0, 1 + 8 + 16 + 32>(out, sink, "SOFT, no LDS, no stash + the UMMA issuer");
  printf("      (MMA tiles issued meanwhile: %lld)\n", out[2]);
  run<1, 1, 0, 0>(out, sink, "SOFT non-volatile LDS");
  run<1, 1, 1, 0>(out, sink, "SOFT non-volatile LDS, scales before the ld wait");
  run<1, 1, 1, 1>(out, sink, "SOFT non-volatile LDS, scales early, both chunks per wait");
  run_pair<1, 32>(out, sink, "SOFT paired rows (2 rows per thread), 32 columns per step");
  run_pair<1, 16>(out, sink, "SOFT paired rows, 16 columns per step");
  run_pair<0, 32>(out, sink, "ARGMAX paired rows, 32 columns per step");
  run_pair<0, 16>(out, sink, "ARGMAX paired rows, 16 columns per step");
  run<0, 0, 0, 1>(out, sink, "ARGMAX as shipped (volatile LDS, both chunks per wait)");
  run<0, 1, 0, 1>(out, sink, "ARGMAX non-volatile LDS");
  run<0, 1, 1, 1>(out, sink, "ARGMAX non-volatile LDS, scales before the ld wait");
  run<0, 1, 1, 0>(out, sink, "ARGMAX non-volatile LDS, scales early, one chunk per wait");
  return 0;
}
