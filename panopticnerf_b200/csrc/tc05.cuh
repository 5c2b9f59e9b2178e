// tc05.cuh — thin inline-PTX wrappers for the sm_100a features the renderer uses:
// mbarrier, bulk-async copy (TMA, UBLKCP), tcgen05 MMA / TMEM alloc / ld / st / commit.
// No CUTLASS/CuTe dependency: descriptor bit layouts follow the PTX ISA 8.6 tcgen05
// "matrix descriptor" and "instruction descriptor" tables.
#pragma once
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace pnr {

#ifndef PNR_WATCHDOG_CYCLES
#define PNR_WATCHDOG_CYCLES (8000000000LL)  // ~4 s @ 2 GHz: a stuck barrier traps instead of hanging the box
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// One lane of a fully converged warp (PTX elect.sync): the compiler knows the guarded code is single-thread.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3FFu) == 0 && (clock64() - t0) > PNR_WATCHDOG_CYCLES) {
      printf("pnr: mbarrier watchdog: block %d thread %d bar 0x%x parity %u\n", (int)blockIdx.x,
             (int)threadIdx.x, bar, parity);
      __trap();
    }
  }
}
// Same, for warps whose waits are long (epilogue, producers): back off between polls so the spinning warps
// do not take issue slots from the MMA-issuing thread that shares their scheduler.
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(40);
    if ((++spins & 0x3FFu) == 0 && (clock64() - t0) > PNR_WATCHDOG_CYCLES) {
      printf("pnr: mbarrier watchdog: block %d thread %d bar 0x%x parity %u\n", (int)blockIdx.x,
             (int)threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// Release / acquire on a shared-memory word (scout -> issuer hand-off of "stage i is ready").
__device__ __forceinline__ void st_release_smem(uint32_t addr, uint32_t v) {
  asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_smem(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
// Monotonic hand-off counter in shared memory (epilogue warps -> scout): relaxed add, no return value.
__device__ __forceinline__ void red_add_smem(uint32_t addr, uint32_t v) {
  asm volatile("red.relaxed.cta.shared::cta.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
// 256-bit global store (sm_100: STG.256): one full 32-byte sector per lane - rows written by one thread each leave the
// SM as whole sectors instead of two half-sector requests.  p must be 32-byte aligned.
__device__ __forceinline__ void st_global_v8(float* p, float a, float b, float c, float d, float e, float f, float g, float h) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d),
               "f"(e), "f"(f), "f"(g), "f"(h)
               : "memory");
}
__device__ __forceinline__ uint32_t ld_volatile_smem(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.volatile.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}

// ---------------------------------------------------------------- proxies / bulk copy
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// 1-D bulk async copy global -> shared, completion counted in bytes on an mbarrier (TMA engine).
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src_gmem, uint32_t bytes,
                                         uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          dst_smem),
      "l"(src_gmem), "r"(bytes), "r"(bar)
      : "memory");
}
// Same copy, multicast: the bytes land at the same shared-memory offset, and complete_tx is signalled on the
// mbarrier at the same offset, in every CTA of the cluster whose bit is set in cta_mask.
__device__ __forceinline__ void bulk_g2s_multicast(uint32_t dst_smem, const void* src_gmem, uint32_t bytes,
                                                   uint32_t bar, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
          dst_smem),
      "l"(src_gmem), "r"(bytes), "r"(bar), "h"(cta_mask)
      : "memory");
}

// ---------------------------------------------------------------- thread-block clusters
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---------------------------------------------------------------- TMEM management
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_result_addr) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_result_addr),
               "n"(NCOLS)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_wait_st() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// All prior tcgen05.mma of this thread arrive (count 1) on the mbarrier when they complete.
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   bar)
               : "memory");
}
// Same, arriving on the mbarrier at this offset in every CTA of the cluster selected by cta_mask.
__device__ __forceinline__ void tc_commit_multicast(uint32_t bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"(cta_mask)
      : "memory");
}

// ---------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor, K-major, no swizzle ("interleaved" 8x16B core matrices).
// Core matrix = 8 rows x 16 bytes, stored as 128 contiguous bytes.
//   lbo_bytes: distance between the two core matrices adjacent in K (one MMA consumes K=16 = 2 cores)
//   sbo_bytes: distance between 8-row groups along M/N
__device__ __forceinline__ uint64_t make_smem_desc_noswz(uint32_t saddr, uint32_t lbo_bytes,
                                                         uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version 1 (sm_100)
  return d;                             // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}
// Operand element formats of kind::f16 (instruction-descriptor encoding).
constexpr int kFmtF16 = 0, kFmtBF16 = 1;
// Instruction descriptor for kind::f16, A and B of format `fmt` (K-major both), D=f32, dense.
__host__ __device__ constexpr uint32_t make_idesc_f32acc(int M, int N, int fmt) {
  return (1u << 4)                     // c_format = F32
         | (uint32_t(fmt) << 7)        // a_format
         | (uint32_t(fmt) << 10)       // b_format
         | (uint32_t(N >> 3) << 17)    // n_dim
         | (uint32_t(M >> 4) << 24);   // m_dim
}

// ---------------------------------------------------------------- MMA issue (single thread)
// D[tmem] (+)= A[tmem] * B[smem]^T ; A is 128 x 16 bf16 held as 8 packed 32-bit TMEM columns.
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                       uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same with the weight-tile descriptor given as its low word only (address | LBO << 16); the high word is the
// constant kDescHiWord (SBO = 128 B, version 1), so advancing K is one 32-bit add.
constexpr uint32_t kDescHiWord = 0x4008u;
__device__ __forceinline__ void mma_ts_lo(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_desc_lo,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 bd;\n\t"
      "mov.b64 bd, {%2, 0x4008};\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], bd, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "r"(b_desc_lo), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                       uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---------------------------------------------------------------- TMEM <-> registers
// 32x32b: thread t of the warp touches TMEM lane (lane_base + t); register i <-> column (col + i).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// ---------------------------------------------------------------- 16-bit hi/lo split
// x = hi + lo + residual, both parts rounded to nearest even in the operand format:
//   bf16 (8-bit significand):  residual <= 2^-18 |x|, fp32 exponent range (never overflows)
//   fp16 (11-bit significand): residual <= 2^-22 |x|, requires |x| < 65504
// Two values are packed per 32-bit word with the lower K index in the low half (K-major A operand).
template <int FMT>
__device__ __forceinline__ void split_x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  if (FMT == kFmtBF16) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));  // d.hi = a, d.lo = b
    const float h0 = __uint_as_float(hi << 16);
    const float h1 = __uint_as_float(hi & 0xFFFF0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(x1 - h1), "f"(x0 - h0));
  } else {
    const __half2 h = __floats2half2_rn(x0, x1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
  }
}

}  // namespace pnr
