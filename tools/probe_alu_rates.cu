// probe_alu_rates.cu — issue cost (cycles per warp-instruction, 2 warps per scheduler as in the epilogue) of the
// instruction kinds the MLP epilogue is made of, on sm_100a.
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

template <int KIND>
__global__ void __launch_bounds__(256, 1) k(int iters, long long* out, float* sink) {
  float v[16];
  uint32_t u[8];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = 0.001f * (threadIdx.x + j);
#pragma unroll
  for (int j = 0; j < 8; ++j) u[j] = threadIdx.x * 7 + j;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (KIND == 0) v[j] = v[j] + 1.0009765625f;                       // FADD
      if (KIND == 1) v[j] = fmaxf(v[j], 0.5f * v[(j + 1) & 15]);          // FMNMX (+FMUL)
      if (KIND == 2) v[j] = fmaf(v[j], 1.0001f, 0.5f);                    // FFMA
    }
    if (KIND == 3) {  // F2FP pack: 8 per iteration
#pragma unroll
      for (int j = 0; j < 8; ++j) { __half2 h = __floats2half2_rn(v[2 * j], v[2 * j + 1]); u[j] ^= *reinterpret_cast<uint32_t*>(&h); v[2 * j] += 1.f; }
    }
    if (KIND == 4) {  // HADD2.F32 unpack: 16 per iteration
#pragma unroll
      for (int j = 0; j < 8; ++j) { __half2 h = *reinterpret_cast<__half2*>(&u[j]); float2 f = __half22float2(h); v[2 * j] += f.x; v[2 * j + 1] += f.y; u[j] += 0x00010001u; }
    }
    if (KIND == 5) {  // FADD2 (packed): 8 per iteration
#pragma unroll
      for (int j = 0; j < 8; ++j) { float2 a = make_float2(v[2 * j], v[2 * j + 1]); a = __fadd2_rn(a, make_float2(1.0009765625f, 0.5f)); v[2 * j] = a.x; v[2 * j + 1] = a.y; }
    }
    if (KIND == 6) {  // FFMA2 (packed): 8 per iteration
#pragma unroll
      for (int j = 0; j < 8; ++j) { float2 a = make_float2(v[2 * j], v[2 * j + 1]); a = __ffma2_rn(a, make_float2(1.0001f, 0.9999f), make_float2(0.5f, 0.25f)); v[2 * j] = a.x; v[2 * j + 1] = a.y; }
    }
    if (KIND == 7) {  // PRMT
#pragma unroll
      for (int j = 0; j < 8; ++j) u[j] = __byte_perm(u[j], u[(j + 1) & 7], 0x7632) + j;
    }
  }
  long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j) acc += v[j];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc += (float)u[j];
  if (acc == 1.2345f) sink[0] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <int KIND>
void run(const char* name, int n_per_iter, long long* d, float* s) {
  const int iters = 4000;
  k<KIND><<<148, 256>>>(iters, d, s);
  cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("ALU %-22s : %.2f cycles per warp-instruction (per warp; 2 warps share a scheduler)\n", name, (double)h / iters / n_per_iter);
}

int main() {
  long long* d; float* s; cudaMalloc(&d, 8); cudaMalloc(&s, 4);
  run<0>("FADD", 16, d, s);
  run<1>("FMNMX+FMUL (2 instr)", 16, d, s);
  run<2>("FFMA", 16, d, s);
  run<3>("F2FP.F16.F32.PACK (+FADD)", 8, d, s);
  run<4>("HADD2.F32 x2 (+2 FADD+IADD)", 8, d, s);
  run<5>("FADD2 (f32x2)", 8, d, s);
  run<6>("FFMA2 (f32x2)", 8, d, s);
  run<7>("PRMT+IADD", 8, d, s);
  return 0;
}
