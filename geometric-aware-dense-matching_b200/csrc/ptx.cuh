// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk[.tensor]), tcgen05 (UMMA / TMEM).
// Bit layouts follow the PTX ISA "tcgen05" chapter (shared-memory matrix descriptor, instruction
// descriptor for .kind::f16).  No CUTLASS dependency.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace gadm {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ----------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (visible as a launch failure) instead of hanging the GPU box.
#ifndef GADM_SPIN_LIMIT
#define GADM_SPIN_LIMIT (1u << 24)
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > GADM_SPIN_LIMIT) __trap();
  }
}
// Same, with a suspend-time hint: the warp sleeps in hardware until the phase completes (or ~20 us pass)
// instead of re-issuing polls that steal issue slots from the epilogue warps on the same scheduler.
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait_hint(bar, parity, 20000u)) {
    if (++spins > (1u << 18)) __trap();
  }
}

// register reallocation between warpgroups (all 128 threads of the warpgroup execute it)
template <int N> __device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// ----------------------------------------------------------------------------- TMA
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 3-D tiled load, coordinates innermost first; completes `bytes` on `bar`.
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// 1-D bulk copy global -> shared (16-byte aligned, size multiple of 16).
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ----------------------------------------------------------------------------- CTA pairs (cta_group::2)
// Two CTAs of a cluster on the two SMs of a TPC execute one tcgen05.mma together: M = 256 (each CTA's TMEM receives its
// own 128 rows), each CTA supplies its 128 rows of A and HALF of the B tile (N/2 columns) from the same shared-memory
// offsets.  Only the leader (cluster rank 0) issues; the barriers the issuer waits on live in the leader's shared
// memory and the peer reaches them through the cluster window.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// (default semantics, release at CTA scope: what the arrival orders here is tcgen05.ld traffic, fenced separately;
// .release.cluster compiles to MEMBAR.ALL.GPU in front of every arrival)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// 3-D tiled load into THIS CTA's shared memory; the bytes complete on a barrier given by its cluster address
// (the leader's, for operands of a paired MMA)
__device__ __forceinline__ void tma_load_3d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr,
                                                 int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[256 x 16: 128 rows from each CTA] * B[N x 16: N/2 columns from each CTA]^T
__device__ __forceinline__ void umma_bf16_ss_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                  uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at this shared-memory offset in BOTH CTAs once the pair's MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(uint16_t(3))
      : "memory");
}

// ----------------------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle:
//   bits [0,14)  start address >> 4          bits [16,30) leading byte offset >> 4 (unused for SW128 K-major: 1)
//   bits [32,46) stride byte offset >> 4 (8 rows * 128 B = 1024 -> 64)
//   bits [46,48) descriptor version = 1 (Blackwell)      bits [61,64) layout type: 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t umma_desc_sw128_kmajor(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// Instruction descriptor, .kind::f16: D = fp32 (bits[4,6)=1), A = B = bf16 (bits[7,10)=1, [10,13)=1),
// A and B K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t umma_idesc_bf16_f32(uint32_t M, uint32_t N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread.
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once all previously issued tcgen05.mma of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns (thread t gets lane t's row).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// 16-column variant (smaller register footprint: more epilogue warps fit)
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// TMEM -> registers in the 16x256b fragment layout (the mma.sync accumulator layout): 16 TMEM lanes starting at the
// lane of `taddr`, NG groups of 8 fp32 columns.  Thread t holds, for group i:
//   r[4i + 0], r[4i + 1] = lane t / 4,     columns 8i + 2 (t % 4) + {0, 1}
//   r[4i + 2], r[4i + 3] = lane t / 4 + 8, same columns
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_frag(uint32_t taddr, uint32_t (&r)[16]) { tmem_ld_16x256b_x4(taddr, r); }
__device__ __forceinline__ void tmem_ld_frag(uint32_t taddr, uint32_t (&r)[8]) { tmem_ld_16x256b_x2(taddr, r); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }


// 8-column load / 8- and 16-column stores (registers -> TMEM), same 32x32b shape
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st(uint32_t taddr, const uint32_t (&r)[8]) { tmem_st_32x8(taddr, r); }
__device__ __forceinline__ void tmem_st(uint32_t taddr, const uint32_t (&r)[16]) { tmem_st_32x16(taddr, r); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Instruction descriptor, .kind::f16 with fp16 A and B (formats 0), fp32 D, both K-major.
__host__ __device__ constexpr uint32_t umma_idesc_f16_f32(uint32_t M, uint32_t N) {
  return (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// D[tmem] (+)= A[tmem] * B[smem]^T : the A operand (M x 16, 16-bit elements, 8 columns) is read from TMEM.
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// explicit shared-memory loads (a generic pointer makes the compiler emit LD instead of LDS)
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
// Same load WITHOUT `volatile`: the compiler may schedule it freely (hoist it over the predicated stash stores and
// the other volatile asm of the epilogue).  The caller pins it below the mbarrier wait that publishes the data by
// deriving `saddr` from a value laundered with opaque() after that wait.
__device__ __forceinline__ float4 lds128_free(uint32_t saddr) {
  float4 v;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
// makes `v` opaque to the optimiser at this point of the volatile-asm order (a scheduling pin, no instruction)
__device__ __forceinline__ uint32_t opaque(uint32_t v) {
  asm volatile("" : "+r"(v) :: "memory");
  return v;
}
// one packed pair (8 bytes) from shared memory
__device__ __forceinline__ uint64_t lds64(uint32_t saddr) {
  uint64_t v;
  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(saddr));
  return v;
}
// predicated (branch-free) 16-byte stores of packed pairs to GLOBAL memory: the fragment-layout matcher keeps its
// argmax stash in an L2-resident workspace (2048 row tracks per CTA do not fit in shared memory)
__device__ __forceinline__ void stg_pred16(bool pred, void* gaddr, uint64_t v0, uint64_t v1) {
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "setp.ne.b32 P, %0, 0;\n\t"
      "@P st.global.v2.b64 [%1], {%2, %3};\n\t}\n"
      ::"r"(uint32_t(pred)), "l"(gaddr), "l"(v0), "l"(v1)
      : "memory");
}
__device__ __forceinline__ void stg_pred32(bool pred, void* gaddr, uint64_t v0, uint64_t v1, uint64_t v2, uint64_t v3) {
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "setp.ne.b32 P, %0, 0;\n\t"
      "@P st.global.v2.b64 [%1], {%2, %3};\n\t"
      "@P st.global.v2.b64 [%1 + 8192], {%4, %5};\n\t}\n"
      ::"r"(uint32_t(pred)), "l"(gaddr), "l"(v0), "l"(v1), "l"(v2), "l"(v3)
      : "memory");
}
// 32-byte global store (sm_100: STG.256); gaddr 32-byte aligned
// fp32 reduction into global memory, 16 bytes at a time (sm_90+), no return value
__device__ __forceinline__ void red_add_v4(float* gaddr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(gaddr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t saddr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// {upper 16 bits: bf16(hi), lower 16 bits: bf16(lo)}, round to nearest even
__device__ __forceinline__ uint32_t cvt_bf16x2(float hi, float lo) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void stg256(float* gaddr, float a, float b, float c, float d, float e, float f, float g,
                                       float h) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"l"(gaddr), "f"(a), "f"(b), "f"(c), "f"(d), "f"(e), "f"(f), "f"(g), "f"(h) : "memory");
}
__device__ __forceinline__ void stg_pred8(bool pred, void* gaddr, uint64_t v) {
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "setp.ne.b32 P, %0, 0;\n\t"
      "@P st.global.b64 [%1], %2;\n\t}\n"
      ::"r"(uint32_t(pred)), "l"(gaddr), "l"(v)
      : "memory");
}
__device__ __forceinline__ uint64_t ldg_cg64(const void* gaddr) {
  uint64_t v;
  asm volatile("ld.global.cg.b64 %0, [%1];" : "=l"(v) : "l"(gaddr) : "memory");
  return v;
}
// L2 (cache-global) load of a float4 the same thread stored earlier
__device__ __forceinline__ float4 ldg_cg128(const void* gaddr) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(gaddr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t smid() {
  uint32_t v;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(v));
  return v;
}
__device__ __forceinline__ float lds32(uint32_t saddr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
  return v;
}

// ---- packed fp32x2 arithmetic (sm_100: FFMA2 / FMUL2 / FADD2): one issue slot for two lanes of work
__device__ __forceinline__ uint64_t pack2(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ uint64_t pack2f(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, uint32_t& lo, uint32_t& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}
__device__ __forceinline__ void unpack2f(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// 2^x on both halves of a packed pair, in place (no register shuffling around the MUFU results)
__device__ __forceinline__ uint64_t ex2_2(uint64_t t) {
  uint64_t d;
  asm("{\n\t.reg .f32 lo, hi;\n\t"
      "mov.b64 {lo, hi}, %1;\n\t"
      "ex2.approx.ftz.f32 lo, lo;\n\t"
      "ex2.approx.ftz.f32 hi, hi;\n\t"
      "mov.b64 %0, {lo, hi};\n\t}"
      : "=l"(d) : "l"(t));
  return d;
}
// 2^(v * g) on both halves WITHOUT the MUFU: round-to-nearest split through the 1.5 * 2^23 constant, degree-4 minimax
// polynomial of 2^f on [-0.5, 0.5] (relative error 2.7e-6), exponent added as an integer.  Seven packed FMA-pipe
// instructions and two IMAD per pair of scores, against one FMUL2 and two MUFU.EX2 (8 cycles of the XU pipe each): the
// SOFT epilogue computes part of its exponentials this way because MUFU, not the FMA pipe, is what it saturates.
// Needs |v * g| < 2^21 and finite (the epilogue guarantees |v * g| <= 58; columns masked to -inf take the MUFU path).
__device__ __forceinline__ uint64_t ex2_2_poly(uint64_t v, uint64_t g2) {
  const uint64_t magic = pack2f(12582912.f, 12582912.f), minus1 = pack2f(-1.f, -1.f);
  const uint64_t t = ffma2(v, g2, magic);          // low mantissa bits = n = round(v * g)
  const uint64_t nn = ffma2(t, minus1, magic);     // -n as a float (exact)
  const uint64_t f = ffma2(v, g2, nn);             // v * g - n in [-0.5, 0.5], one rounding
  uint64_t q = ffma2(pack2f(0.009570077061653137f, 0.009570077061653137f), f,
                     pack2f(0.055917829275131226f, 0.055917829275131226f));
  q = ffma2(q, f, pack2f(0.240247443318367f, 0.240247443318367f));
  q = ffma2(q, f, pack2f(0.6931217908859253f, 0.6931217908859253f));
  q = ffma2(q, f, pack2f(0.9999992847442627f, 0.9999992847442627f));
  uint32_t tl, th, ql, qh;
  unpack2(t, tl, th);
  unpack2(q, ql, qh);
  return pack2(ql + (tl << 23), qh + (th << 23));  // 2^n: the shift drops everything but n mod 512
}
__device__ __forceinline__ void sts32(uint32_t saddr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(saddr), "f"(v) : "memory");
}
// predicated (branch-free) store of 8 floats: float4 k goes to saddr + k * 8192 (the stash of the matcher epilogue:
// plane-major so that a full warp stores without bank conflicts).  The operands are the UNPACKED scores the max tree
// reads anyway: handing the store the packed f32x2 pairs makes ptxas copy every stored register (the pair is
// overwritten in place by the exponent multiply): 229 -> 201 instructions per 32 scores of the SOFT epilogue.
__device__ __forceinline__ void sts_stash8(bool pred, uint32_t saddr, const float (&f)[8]) {
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "setp.ne.b32 P, %0, 0;\n\t"
      "@P st.shared.v4.f32 [%1], {%2, %3, %4, %5};\n\t"
      "@P st.shared.v4.f32 [%1 + 8192], {%6, %7, %8, %9};\n\t}\n"
      ::"r"(uint32_t(pred)), "r"(saddr), "f"(f[0]), "f"(f[1]), "f"(f[2]), "f"(f[3]), "f"(f[4]), "f"(f[5]), "f"(f[6]),
        "f"(f[7])
      : "memory");
}
__device__ __forceinline__ float fmax3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }  // -> FMNMX3
// {hi, lo} fp32 -> packed f16x2 (round to nearest even); `lo` lands in the low half
__device__ __forceinline__ uint32_t cvt_f16x2(float hi, float lo) {
  uint32_t d;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

__device__ __forceinline__ uint32_t hmul2(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}

// clock read that is ordered after the producer of `dep` (profiling builds only)
__device__ __forceinline__ long long clock_after(uint32_t dep) {
  long long t;
  asm volatile("{\n\t.reg .b32 z;\n\tand.b32 z, %1, 0;\n\tmov.u64 %0, %%clock64;\n\t}" : "=l"(t) : "r"(dep) : "memory");
  return t;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

}  // namespace ptx
}  // namespace gadm
