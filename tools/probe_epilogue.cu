// probe_epilogue.cu — cycles per 16-column epilogue group (tcgen05.ld -> bias, ReLU -> hi/lo split ->
// tcgen05.st) with 8 warps per CTA as in the fused kernel, and which instruction class costs what.
// variant bits: 1 = skip the lo part (1-pass), 2 = replace F2FP conversions by integer packs (wrong values,
// timing only), 4 = skip tcgen05.st, 8 = skip tcgen05.ld (reuse registers), 16 = skip bias LDS
#include <cstdio>
#include "../panopticnerf_b200/csrc/tc05.cuh"
using namespace pnr;

__global__ void __launch_bounds__(256, 1) k(int variant, int iters, long long* out, float* sink) {
  __shared__ __align__(16) float bias[256];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  bias[threadIdx.x] = 0.001f * threadIdx.x;
  if (warp == 0) { tmem_alloc<512>(smem_u32(&slot)); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot + ((uint32_t)((warp & 3) * 32) << 16);
  const int ch = warp >> 2;
  uint32_t r[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(0.01f * (lane + j));
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll 1
    for (int g = ch * 4; g < ch * 4 + 4; ++g) {
      if (!(variant & 8)) { tmem_ld16(tmem + g * 16, r); tc_wait_ld(); }
      uint32_t hi[8], lo[8];
      const float4* b4 = reinterpret_cast<const float4*>(bias + g * 16);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 b = (variant & 16) ? make_float4(0.1f, 0.2f, 0.3f, 0.4f) : b4[q];
        const float v0 = fmaxf(__uint_as_float(r[4 * q + 0]) + b.x, 0.f);
        const float v1 = fmaxf(__uint_as_float(r[4 * q + 1]) + b.y, 0.f);
        const float v2 = fmaxf(__uint_as_float(r[4 * q + 2]) + b.z, 0.f);
        const float v3 = fmaxf(__uint_as_float(r[4 * q + 3]) + b.w, 0.f);
        if (variant & 2) {
          hi[2 * q] = __byte_perm(__float_as_uint(v0), __float_as_uint(v1), 0x7632);
          hi[2 * q + 1] = __byte_perm(__float_as_uint(v2), __float_as_uint(v3), 0x7632);
          lo[2 * q] = __byte_perm(__float_as_uint(v0 - 1.f), __float_as_uint(v1 - 1.f), 0x7632);
          lo[2 * q + 1] = __byte_perm(__float_as_uint(v2 - 1.f), __float_as_uint(v3 - 1.f), 0x7632);
        } else if (variant & 1) {
          const __half2 h0 = __floats2half2_rn(v0, v1), h1 = __floats2half2_rn(v2, v3);
          hi[2 * q] = *reinterpret_cast<const uint32_t*>(&h0);
          hi[2 * q + 1] = *reinterpret_cast<const uint32_t*>(&h1);
          lo[2 * q] = lo[2 * q + 1] = 0;
        } else {
          split_x2<kFmtF16>(v0, v1, hi[2 * q], lo[2 * q]);
          split_x2<kFmtF16>(v2, v3, hi[2 * q + 1], lo[2 * q + 1]);
        }
      }
      if (!(variant & 4)) {
        tmem_st8(tmem + 256 + g * 8, hi);
        if (!(variant & 1)) tmem_st8(tmem + 384 + g * 8, lo);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc += __uint_as_float(hi[j] ^ lo[j]);
      }
      if (variant & 8) {
#pragma unroll
        for (int j = 0; j < 16; ++j) r[j] ^= hi[j & 7];
      }
    }
    if (!(variant & 4)) tc_wait_st();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(slot);
}

int main() {
  long long* d; float* s; cudaMalloc(&d, 8); cudaMalloc(&s, 4);
  const int iters = 2000;
  for (int v : {0, 1, 2, 4, 8, 16, 2 | 4 | 8, 4 | 8, 2 | 16}) {
    k<<<148, 256>>>(v, iters, d, s);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("EPI variant=%2d (x1=%d noF2FP=%d noST=%d noLD=%d noLDS=%d): %.0f cycles per 16-col group per warp (4 groups/iter)  %s\n", v, v & 1,
           (v >> 1) & 1, (v >> 2) & 1, (v >> 3) & 1, (v >> 4) & 1, (double)h / iters / 4, cudaGetErrorString(e));
  }
  return 0;
}
