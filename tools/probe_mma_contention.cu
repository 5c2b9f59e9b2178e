// probe_mma_contention.cu — does concurrent TMA (bulk copies into shared memory) or concurrent epilogue
// TMEM traffic (tcgen05.ld / tcgen05.st from other warps) slow the tensor pipe?  One CTA per SM runs a long
// TS-form N=128 MMA chain while optional background warps generate the other traffic.
#include <cstdio>
#include <cstdlib>
#include "../panopticnerf_b200/csrc/tc05.cuh"
using namespace pnr;

__global__ void __launch_bounds__(320, 1) k(int N, int bg_tma, int bg_tmem, int count, const uint8_t* gsrc, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 160 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 8);
  volatile int* stop = reinterpret_cast<volatile int*>(bars + 9);
  const uint32_t bar = smem_u32(&bars[0]), bar_t0 = smem_u32(&bars[1]), bar_t1 = smem_u32(&bars[2]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) { tmem_alloc<512>(smem_u32(slot)); tmem_relinquish(); }
  if (threadIdx.x == 32) { mbar_init(bar, 1); mbar_init(bar_t0, 1); mbar_init(bar_t1, 1); *stop = 0; fence_mbar_init(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_f32acc(128, N, kFmtF16);
      const uint32_t b_lbo = N * 16u;
      const uint64_t bdesc = make_smem_desc_noswz(smem_u32(smem), b_lbo, 128);
      long long t0 = clock64();
      for (int i = 0; i < count; i += 4) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          mma_ts(tmem + (uint32_t)((j & 1) * N), tmem + 384 + j * 8, bdesc + (uint64_t)(j * ((2 * b_lbo) >> 4)), idesc, 1);
      }
      tc_commit(bar);
      mbar_wait(bar, 0);
      long long t1 = clock64();
      *stop = 1;
      if (blockIdx.x == 0) out[0] = t1 - t0;
    }
  } else if (warp == 2 && bg_tma) {
    if (elect_one()) {  // continuous 2 x 32 KB bulk copies (~ the weight stream of the real kernel when looped)
      uint32_t ph = 0;
      const uint8_t* src = gsrc + (size_t)blockIdx.x * 65536;
      while (!*stop) {
        mbar_arrive_expect_tx(bar_t0, 32768);
        bulk_g2s(smem_u32(smem + 64 * 1024), src, 32768, bar_t0);
        mbar_arrive_expect_tx(bar_t1, 32768);
        bulk_g2s(smem_u32(smem + 96 * 1024), src + 32768, 32768, bar_t1);
        mbar_wait(bar_t0, ph);
        mbar_wait(bar_t1, ph);
        ph ^= 1;
      }
    }
  } else if (warp >= 4 && bg_tmem) {  // 6 warps: read 16 accumulator columns, write 8+8 activation columns, repeat
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t r[16];
    while (!*stop) {
      tmem_ld16(tmem + lane_off + 256 + ((warp >> 2) & 1) * 16, r);
      tc_wait_ld();
      uint32_t h[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) h[j] = r[j] ^ r[j + 8];
      tmem_st8(tmem + lane_off + 448 + (warp & 7) * 8, h);
      tc_wait_st();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

int main() {
  long long* d; uint8_t* g;
  cudaMalloc(&d, 16); cudaMalloc(&g, 148 * 65536); cudaMemset(g, 0, 148 * 65536);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 164 * 1024);
  const int count = 4096;
  for (int N : {128, 256})
    for (int bg = 0; bg < 4; ++bg) {
      k<<<148, 320, 164 * 1024>>>(N, bg & 1, bg >> 1, count, g, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      printf("CONTENTION N=%d tma=%d tmem_traffic=%d : %.1f cyc/mma (%s)\n", N, bg & 1, bg >> 1, (double)h / count, cudaGetErrorString(e));
    }
  return 0;
}
