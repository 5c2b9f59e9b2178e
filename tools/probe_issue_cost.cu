// probe_issue_cost.cu — what does the per-stage bookkeeping of the MMA issuer cost the tensor pipe?
// 148 CTAs, one issuing thread each, stages of 12 x (N=128 TS) MMAs; variants add one element at a time.
#include <cstdio>
#include "../panopticnerf_b200/csrc/tc05.cuh"
using namespace pnr;

struct Tab { uint16_t n, acc, a_off, a_lo, lo16, flags; };
__constant__ Tab c_tab[64];

// variant bits: 1 = commit to an mbarrier after every stage, 2 = read the stage descriptor from __constant__
// with a register index, 4 = poll a shared word before each stage, 8 = mbarrier try_wait (already complete) per stage
__global__ void __launch_bounds__(128, 1) k(int variant, int stages, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 64 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 16);
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc<512>(smem_u32(slot)); tmem_relinquish(); }
  if (threadIdx.x == 32) {
    for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bars[i]), 1);
    mbar_init(smem_u32(&bars[8]), 1);
    *reinterpret_cast<volatile uint32_t*>(bars + 17) = 0xFFFFFFF0u;
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  const uint32_t word = smem_u32(bars + 17);
  if (warp == 1 && elect_one()) {
    mbar_arrive(smem_u32(&bars[8]));  // completes phase 0 -> try_wait(parity 0) is "already complete"
    long long t0 = clock64();
#pragma unroll 1
    for (int s = 0; s < stages; ++s) {
      uint32_t n = 128, acc = (s & 1) * 128, a_off = 256, a_lo = 384, lo16 = 128 * 8;
      if (variant & 2) {
        const Tab t = c_tab[s & 63];
        n = t.n; acc = t.acc; a_off = t.a_off; a_lo = t.a_lo; lo16 = t.lo16;
      }
      if (variant & 4) { while (ld_acquire_smem(word) <= (uint32_t)s) {} }
      if (variant & 8) mbar_wait(smem_u32(&bars[8]), 0);
      const uint32_t idesc = make_idesc_f32acc(128, n, kFmtF16);
      const uint32_t b_lbo = n * 16u;
      const uint64_t bdesc0 = make_smem_desc_noswz(smem_u32(smem) + (s & 1) * 32768, b_lbo, 128);
      const uint64_t bdesc0_lo = bdesc0 + (uint64_t)lo16;
      const uint32_t b_inc = (2u * b_lbo) >> 4;
#pragma unroll
      for (uint32_t ks = 0; ks < 4; ++ks) {
        mma_ts(tmem + acc, tmem + a_off + ks * 8, bdesc0 + (uint64_t)(ks * b_inc), idesc, 1u);
        mma_ts(tmem + acc, tmem + a_lo + ks * 8, bdesc0 + (uint64_t)(ks * b_inc), idesc, 1u);
        mma_ts(tmem + acc, tmem + a_off + ks * 8, bdesc0_lo + (uint64_t)(ks * b_inc), idesc, 1u);
      }
      if (variant & 1) tc_commit(smem_u32(&bars[s & 3]));
    }
    long long t1 = clock64();
    tc_commit(smem_u32(&bars[7]));
    mbar_wait(smem_u32(&bars[7]), 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  Tab h[64];
  for (int i = 0; i < 64; ++i) h[i] = Tab{128, (uint16_t)((i & 1) * 128), 256, 384, 128 * 8, 0};
  cudaMemcpyToSymbol(c_tab, h, sizeof(h));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
  const int stages = 512;
  for (int v : {0, 1, 2, 3, 4, 8, 7, 15}) {
    k<<<148, 128, 66 * 1024>>>(v, stages, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long r[2] = {0, 0}; cudaMemcpy(r, d, 16, cudaMemcpyDeviceToHost);
    printf("ISSUE variant=%2d (commit=%d const=%d poll=%d trywait=%d): issue %.0f cyc/stage, complete %.0f cyc/stage (pipe ideal 768)  %s\n",
           v, v & 1, (v >> 1) & 1, (v >> 2) & 1, (v >> 3) & 1, (double)r[0] / stages, (double)r[1] / stages, cudaGetErrorString(e));
  }
  return 0;
}
