// probe_pair_mma.cu — bring-up for round 2: tcgen05.mma.cta_group::2 (a CTA pair drives one 256 x N x 16 MMA),
// TS form (A = this CTA's 128 rows in its own tensor memory), B split by rows between the two CTAs' shared
// memory in the same no-swizzle K-major core-matrix layout the fused MLP kernel uses.  Checks the result against
// a host reference for N = 64 / 128 / 256 and K = 64.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include "../panopticnerf_b200/csrc/tc05.cuh"
using namespace pnr;

constexpr int kK = 64;   // 4 K16 steps

__device__ __forceinline__ void mma_ts_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void commit_pair(uint32_t bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"(mask)
      : "memory");
}

// A [256, kK] fp16 row-major, B [N, kK] fp16 row-major, D [256, N] fp32
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
k(const __half* A, const __half* B, float* D, int N) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 64 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 4);
  const uint32_t rank = cluster_ctarank();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nh = N / 2;   // rows of B this CTA holds
  // B half -> shared memory, no-swizzle K-major: element (n, k) at ((k/8) * nh + n) * 16 + (k%8) * 2 bytes
  for (int i = threadIdx.x; i < nh * kK; i += blockDim.x) {
    const int n = i / kK, kk = i % kK;
    *reinterpret_cast<__half*>(smem + ((kk / 8) * nh + n) * 16 + (kk % 8) * 2) = B[(size_t)(rank * nh + n) * kK + kk];
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 32) {
    mbar_init(smem_u32(&bars[0]), 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  // A: this thread's row (TMEM lane), K values packed two per 32-bit column at columns 256 ..
  {
    const int row = rank * 128 + threadIdx.x;
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    for (int g = 0; g < kK / 16; ++g) {
      uint32_t v[8];
      for (int j = 0; j < 8; ++j) {
        const __half2 h = __halves2half2(A[(size_t)row * kK + g * 16 + 2 * j], A[(size_t)row * kK + g * 16 + 2 * j + 1]);
        v[j] = *reinterpret_cast<const uint32_t*>(&h);
      }
      tmem_st8(lane_addr + 256 + g * 8, v);
    }
    tc_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();   // both CTAs' operands are in place
  tc_fence_after();
  if (rank == 0 && warp == 1 && elect_one()) {
    const uint32_t idesc = make_idesc_f32acc(256, N, kFmtF16);
    const uint64_t bdesc0 = make_smem_desc_noswz(smem_u32(smem), (uint32_t)nh * 16u, 128);
    const uint32_t b_inc = (2u * (uint32_t)nh * 16u) >> 4;
    for (uint32_t ks = 0; ks < kK / 16; ++ks)
      mma_ts_pair(tmem, tmem + 256 + ks * 8, bdesc0 + (uint64_t)(ks * b_inc), idesc, ks == 0 ? 0u : 1u);
    commit_pair(smem_u32(&bars[0]), (uint16_t)3);
  }
  mbar_wait(smem_u32(&bars[0]), 0);
  tc_fence_after();
  {
    const int row = rank * 128 + threadIdx.x;
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    for (int g = 0; g < N / 16; ++g) {
      uint32_t r[16];
      tmem_ld16(lane_addr + g * 16, r);
      tc_wait_ld();
      for (int j = 0; j < 16; ++j) D[(size_t)row * N + g * 16 + j] = __uint_as_float(r[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  (void)lane;
}

int main() {
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65 * 1024);
  for (int N : {64, 128, 256}) {
    std::vector<__half> hA(256 * kK), hB((size_t)N * kK);
    srand(N);
    for (auto& v : hA) v = __float2half((float)(rand() % 7 - 3));
    for (auto& v : hB) v = __float2half((float)(rand() % 5 - 2));
    __half *dA, *dB; float* dD;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, (size_t)256 * N * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, (size_t)256 * N * 4);
    k<<<2, 128, 65 * 1024>>>(dA, dB, dD, N);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> hD((size_t)256 * N);
    cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0; int bad = 0;
    for (int m = 0; m < 256; ++m)
      for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int kk = 0; kk < kK; ++kk) ref += (double)__half2float(hA[m * kK + kk]) * (double)__half2float(hB[(size_t)n * kK + kk]);
        const double d = fabs(ref - (double)hD[(size_t)m * N + n]);
        if (!(d <= 1e-3)) ++bad;
        if (d > worst || d != d) worst = d;
      }
    printf("PAIR-MMA N=%3d: %s  bad=%d of %d  worst abs err %.3g  (%s)\n", N, bad == 0 ? "PASS" : "FAIL", bad, 256 * N,
           worst, cudaGetErrorString(e));
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
  }
  return 0;
}
