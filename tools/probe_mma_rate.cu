// probe_mma_rate.cu — tensor-pipe time per tcgen05.mma (M=128, K=16, kind::f16) as a function of N and of
// the A-operand source, measured with clock64 around a long back-to-back chain issued by one thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/probe_mma_rate tools/probe_mma_rate.cu
#include <cstdio>
#include <cstdlib>
#include "../panopticnerf_b200/csrc/tc05.cuh"
using namespace pnr;

// mode: 0 = TS (A in TMEM), 1 = SS (A in smem).  nacc: number of distinct accumulators cycled through.
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int mode, int nacc, int count, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 96 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 4);
  const uint32_t bar = smem_u32(&bars[0]);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // fp16 1.0
  if (warp == 0) { tmem_alloc<512>(smem_u32(slot)); tmem_relinquish(); }
  if (threadIdx.x == 32) { mbar_init(bar, 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 1 && elect_one()) {
    const uint32_t idesc = make_idesc_f32acc(128, N, kFmtF16);
    const uint32_t b_lbo = N * 16u;
    const uint64_t bdesc = make_smem_desc_noswz(smem_u32(smem), b_lbo, 128);
    const uint64_t adesc = make_smem_desc_noswz(smem_u32(smem + 64 * 1024), 128 * 16, 128);
    // warm-up
    for (int i = 0; i < 16; ++i) {
      if (mode == 0) mma_ts(tmem, tmem + 384, bdesc, idesc, 1); else mma_ss(tmem, adesc, bdesc, idesc, 1);
    }
    tc_commit(bar);
    mbar_wait(bar, 0);
    long long t0 = clock64();
    for (int i = 0; i < count; ++i) {
      const uint32_t d = tmem + (uint32_t)((i % nacc) * N);
      if (mode == 0) mma_ts(d, tmem + 384 + (i & 7) * 8, bdesc + (uint64_t)((i & 3) * ((2 * b_lbo) >> 4)), idesc, 1);
      else mma_ss(d, adesc + (uint64_t)((i & 3) * 256), bdesc + (uint64_t)((i & 3) * ((2 * b_lbo) >> 4)), idesc, 1);
    }
    long long t1 = clock64();
    tc_commit(bar);
    mbar_wait(bar, 1);
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int count = 2048;
  for (int mode = 0; mode < 2; ++mode)
    for (int N : {64, 128, 256})
      for (int nacc : {1, 2}) {
        if (nacc * N > 256) continue;
        rate_kernel<<<1, 128, 100 * 1024>>>(N, mode, nacc, count, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2] = {0, 0};
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("RATE mode=%s N=%3d nacc=%d : issue %.1f cyc/mma, complete %.1f cyc/mma  (%s)\n", mode ? "SS" : "TS", N, nacc,
               (double)h[0] / count, (double)h[1] / count, cudaGetErrorString(e));
      }
  // all SMs busy? (power / clock effects): same chain on 148 CTAs
  rate_kernel<<<148, 128, 100 * 1024>>>(128, 0, 2, count, d);
  cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("RATE 148 CTAs TS N=128 nacc=2 : issue %.1f complete %.1f cyc/mma\n", (double)h[0] / count, (double)h[1] / count);
  rate_kernel<<<148, 128, 100 * 1024>>>(256, 0, 1, count, d);
  cudaDeviceSynchronize();
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("RATE 148 CTAs TS N=256 nacc=1 : issue %.1f complete %.1f cyc/mma\n", (double)h[0] / count, (double)h[1] / count);
  return 0;
}
