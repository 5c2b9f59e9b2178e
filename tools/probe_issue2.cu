// probe_issue2.cu — raw issue cost of tcgen05.mma / tcgen05.commit from one thread, with and without
// ALU-saturating warps on the same SM sub-partitions (what the MLP kernel's epilogue warps are).
// Small N makes the tensor pipe fast enough that the issue path is what is measured.
#include <cstdio>
#include "../panopticnerf_b200/csrc/tc05.cuh"
using namespace pnr;

// hogs: number of warps (ids 0..hogs-1) spinning on dependent FP32 / conversion math while the issuer runs
__global__ void __launch_bounds__(512, 1) k(int n, int hogs, int commit, int stages, long long* out, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 64 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 16);
  volatile uint32_t* stop = reinterpret_cast<volatile uint32_t*>(bars + 18);
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc<512>(smem_u32(slot)); tmem_relinquish(); }
  if (threadIdx.x == 32) {
    for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bars[i]), 1);
    *stop = 0u;
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 13) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_f32acc(128, n, kFmtF16);
      const uint32_t b_lbo = n * 16u;
      const uint32_t b_inc = (2u * b_lbo) >> 4;
      long long t0 = clock64();
#pragma unroll 1
      for (int s = 0; s < stages; ++s) {
        const uint64_t bdesc0 = make_smem_desc_noswz(smem_u32(smem) + (s & 1) * 32768, b_lbo, 128);
        const uint64_t bdesc0_lo = bdesc0 + (uint64_t)(n * 8);
        const uint32_t acc = (s & 1) * 128, a_off = 256, a_lo = 384;
#pragma unroll
        for (uint32_t ks = 0; ks < 4; ++ks) {
          mma_ts(tmem + acc, tmem + a_off + ks * 8, bdesc0 + (uint64_t)(ks * b_inc), idesc, 1u);
          mma_ts(tmem + acc, tmem + a_lo + ks * 8, bdesc0 + (uint64_t)(ks * b_inc), idesc, 1u);
          mma_ts(tmem + acc, tmem + a_off + ks * 8, bdesc0_lo + (uint64_t)(ks * b_inc), idesc, 1u);
        }
        if (commit) tc_commit(smem_u32(&bars[s & 3]));
      }
      long long t1 = clock64();
      tc_commit(smem_u32(&bars[7]));
      mbar_wait(smem_u32(&bars[7]), 0);
      long long t2 = clock64();
      *stop = 1u;
      if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  } else if (warp < hogs) {
    float a = threadIdx.x * 1e-3f, b = 1.0001f, c = 0.f;
    uint32_t h = 0;
    while (*stop == 0u) {
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        a = fmaxf(a + b, 0.f);
        uint32_t hi, lo;
        split_x2<kFmtF16>(a, b, hi, lo);
        h ^= hi + lo;
        c += a * b;
      }
    }
    if (h == 0x12345u) sink[threadIdx.x] = c;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  float* sink; cudaMalloc(&sink, 4096);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
  const int stages = 512;
  for (int hogs : {0, 8, 12}) {
    for (int n : {16, 32, 64, 128}) {
      for (int commit : {0, 1}) {
        k<<<148, 512, 66 * 1024>>>(n, hogs, commit, stages, d, sink);
        cudaError_t e = cudaDeviceSynchronize();
        long long r[2] = {0, 0}; cudaMemcpy(r, d, 16, cudaMemcpyDeviceToHost);
        printf("hogs=%2d N=%3d commit=%d: issue %.0f cyc / 12-MMA stage (%.1f per MMA), complete %.0f (pipe ideal %d)  %s\n",
               hogs, n, commit, (double)r[0] / stages, (double)r[0] / stages / 12.0, (double)r[1] / stages, n * 6,
               cudaGetErrorString(e));
      }
    }
  }
  return 0;
}
