// probe_tcgen05.cu — standalone bring-up probe for the tcgen05 building blocks used by the fused
// MLP kernel: TMEM alloc, tcgen05.st of a bf16 hi/lo split A operand, 1-D bulk-copy (TMA) ring of
// pre-packed no-swizzle K-major weight stages, tcgen05.mma (A from TMEM or SMEM), tcgen05.commit,
// tcgen05.ld epilogue.  One CTA computes OUT[128 x N] = A[128 x K] * W[N x K]^T.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o probe tools/probe_tcgen05.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cstring>
#include <cuda_bf16.h>
#include "../panopticnerf_b200/csrc/tc05.cuh"

using namespace pnr;

constexpr int kM = 128;
constexpr int kNStage = 2;
constexpr int kStageBytesMax = 256 * 64 * 2;  // N=256 rows x 64 K x bf16

struct ProbeParams {
  const float* A;          // [128, K] fp32
  const uint8_t* Wpacked;  // stage stream
  float* out;              // [128, N]
  int N, K, passes, mode_ss, swap_lbo_sbo;
};

__global__ void __launch_bounds__(192, 1) probe_kernel(ProbeParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // layout: [0, 2*32K) weight ring | [64K, 64K+128K) A hi/lo for SS mode | barriers at the end
  uint8_t* ring = smem;
  uint8_t* a_sm = smem + kNStage * kStageBytesMax;  // hi then lo, each 128*K*2 bytes
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kNStage * kStageBytesMax + 2 * kM * 256 * 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  const uint32_t bar_full0 = smem_u32(&bars[0]);
  const uint32_t bar_empty0 = smem_u32(&bars[kNStage]);
  const uint32_t bar_a_ready = smem_u32(&bars[2 * kNStage]);
  const uint32_t bar_acc_ready = smem_u32(&bars[2 * kNStage + 1]);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = p.N, K = p.K;
  const int nchunk = K / 64;
  const int nstages_total = nchunk * (p.passes == 3 ? 2 : 1);
  const uint32_t stage_bytes = N * 128;

  if (warp == 0) {
    tmem_alloc<512>(smem_u32(tmem_slot));
    tmem_relinquish();
  }
  if (threadIdx.x == 32) {
    for (int s = 0; s < kNStage; ++s) {
      mbar_init(bar_full0 + 8 * s, 1);
      mbar_init(bar_empty0 + 8 * s, 1);
    }
    mbar_init(bar_a_ready, 128);
    mbar_init(bar_acc_ready, 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t_acc = tmem_base, t_ahi = tmem_base + 256, t_alo = tmem_base + 384;

  if (warp == 0) {
    if (lane == 0) {  // ---- TMA producer
      for (int i = 0; i < nstages_total; ++i) {
        int s = i % kNStage;
        uint32_t ph = (i / kNStage) & 1;
        mbar_wait(bar_empty0 + 8 * s, ph ^ 1);
        mbar_arrive_expect_tx(bar_full0 + 8 * s, stage_bytes);
        bulk_g2s(smem_u32(ring + s * kStageBytesMax), p.Wpacked + (size_t)i * stage_bytes,
                 stage_bytes, bar_full0 + 8 * s);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ---- MMA issuer
      const uint32_t idesc = make_idesc_f32acc(kM, N, kFmtBF16);
      const uint32_t b_lbo = N * 16, b_sbo = 128;
      const uint32_t a_lbo = kM * 16, a_sbo = 128;
      mbar_wait(bar_a_ready, 0);
      tc_fence_after();
      uint32_t acc_flag = 0;
      for (int i = 0; i < nstages_total; ++i) {
        int s = i % kNStage;
        uint32_t ph = (i / kNStage) & 1;
        int chunk = (p.passes == 3) ? (i >> 1) : i;
        bool is_lo = (p.passes == 3) && (i & 1);
        mbar_wait(bar_full0 + 8 * s, ph);
        tc_fence_after();
        uint32_t sb = smem_u32(ring + s * kStageBytesMax);
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t baddr = sb + ks * 2 * b_lbo;
          uint64_t bdesc = p.swap_lbo_sbo ? make_smem_desc_noswz(baddr, b_sbo, b_lbo)
                                          : make_smem_desc_noswz(baddr, b_lbo, b_sbo);
          int kstep = chunk * 4 + ks;  // global K/16 index
          if (!p.mode_ss) {
            if (!is_lo) {
              mma_ts(t_acc, t_ahi + 8 * kstep, bdesc, idesc, acc_flag);
              acc_flag = 1;
              if (p.passes == 3) mma_ts(t_acc, t_alo + 8 * kstep, bdesc, idesc, 1);
            } else {
              mma_ts(t_acc, t_ahi + 8 * kstep, bdesc, idesc, 1);
            }
          } else {
            uint32_t ahi = smem_u32(a_sm) + kstep * 2 * a_lbo;
            uint32_t alo = ahi + kM * K * 2;
            uint64_t dhi = p.swap_lbo_sbo ? make_smem_desc_noswz(ahi, a_sbo, a_lbo)
                                          : make_smem_desc_noswz(ahi, a_lbo, a_sbo);
            uint64_t dlo = p.swap_lbo_sbo ? make_smem_desc_noswz(alo, a_sbo, a_lbo)
                                          : make_smem_desc_noswz(alo, a_lbo, a_sbo);
            if (!is_lo) {
              mma_ss(t_acc, dhi, bdesc, idesc, acc_flag);
              acc_flag = 1;
              if (p.passes == 3) mma_ss(t_acc, dlo, bdesc, idesc, 1);
            } else {
              mma_ss(t_acc, dhi, bdesc, idesc, 1);
            }
          }
        }
        tc_commit(bar_empty0 + 8 * s);
      }
      tc_commit(bar_acc_ready);
    }
  } else {
    // ---- epilogue / A-producer warps 2..5 : TMEM lane quarter = warp % 4
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const float* arow = p.A + (size_t)row * K;
    for (int k0 = 0; k0 < K; k0 += 16) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) split_x2<kFmtBF16>(arow[k0 + 2 * j], arow[k0 + 2 * j + 1], hi[j], lo[j]);
      if (!p.mode_ss) {
        tmem_st8(t_ahi + lane_off + k0 / 2, hi);
        tmem_st8(t_alo + lane_off + k0 / 2, lo);
      } else {
        // no-swizzle K-major: byte offset = (kcore*128 + row)*16 + (k%8)*2
        uint4* dh = reinterpret_cast<uint4*>(a_sm);
        uint4* dl = reinterpret_cast<uint4*>(a_sm + kM * K * 2);
        int kc = k0 / 8;
        dh[(kc)*kM + row] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        dh[(kc + 1) * kM + row] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
        dl[(kc)*kM + row] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        dl[(kc + 1) * kM + row] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
      }
    }
    if (!p.mode_ss) {
      tc_wait_st();
    } else {
      fence_proxy_async_smem();
    }
    tc_fence_before();
    mbar_arrive(bar_a_ready);

    mbar_wait(bar_acc_ready, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(t_acc + lane_off + c0, r);
      tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) p.out[(size_t)row * N + c0 + j] = __uint_as_float(r[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base);
}

static uint16_t f2bf(float x) {  // round-to-nearest-even
  uint32_t u;
  memcpy(&u, &x, 4);
  uint32_t r = u + 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(r >> 16);
}
static float bf2f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

int run(int N, int K, int passes, int mode_ss, int swap) {
  std::vector<float> A((size_t)kM * K), W((size_t)N * K);
  srand(1234);
  for (auto& v : A) v = (rand() / (float)RAND_MAX) * 2.f - 0.5f;
  for (auto& v : W) v = ((rand() / (float)RAND_MAX) * 2.f - 1.f) * 0.0625f;
  int nchunk = K / 64;
  int nst = nchunk * (passes == 3 ? 2 : 1);
  size_t stage_bytes = (size_t)N * 128;
  std::vector<uint16_t> packed(nst * stage_bytes / 2);
  for (int c = 0; c < nchunk; ++c)
    for (int part = 0; part < (passes == 3 ? 2 : 1); ++part) {
      int si = (passes == 3) ? (2 * c + part) : c;
      uint16_t* dst = packed.data() + si * stage_bytes / 2;
      for (int kc = 0; kc < 8; ++kc)
        for (int n = 0; n < N; ++n)
          for (int e = 0; e < 8; ++e) {
            float w = W[(size_t)n * K + c * 64 + kc * 8 + e];
            uint16_t h = f2bf(w);
            uint16_t v = part == 0 ? h : f2bf(w - bf2f(h));
            dst[(kc * N + n) * 8 + e] = v;
          }
    }
  float *dA, *dO;
  uint8_t* dW;
  cudaMalloc(&dA, A.size() * 4);
  cudaMalloc(&dO, (size_t)kM * N * 4);
  cudaMalloc(&dW, packed.size() * 2);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, packed.data(), packed.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dO, 0xFF, (size_t)kM * N * 4);
  ProbeParams p{dA, dW, dO, N, K, passes, mode_ss, swap};
  size_t smem = kNStage * kStageBytesMax + 2 * kM * 256 * 2 + 256;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_kernel<<<1, 192, smem>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("PROBE N=%d K=%d passes=%d ss=%d swap=%d : CUDA error %s\n", N, K, passes, mode_ss, swap,
           cudaGetErrorString(e));
    return 2;
  }
  std::vector<float> O((size_t)kM * N);
  cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < kM; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)A[(size_t)m * K + k] * (double)W[(size_t)n * K + k];
      double d = fabs(ref - (double)O[(size_t)m * N + n]);
      if (!(d <= maxerr)) maxerr = d;
      if (fabs(ref) > maxref) maxref = fabs(ref);
    }
  printf("PROBE N=%d K=%d passes=%d ss=%d swap=%d : max_abs_err=%.3e max_ref=%.3e rel=%.3e  out[0][0..3]=%g %g %g %g\n",
         N, K, passes, mode_ss, swap, maxerr, maxref, maxerr / maxref, O[0], O[1], O[2], O[3]);
  cudaFree(dA);
  cudaFree(dO);
  cudaFree(dW);
  return 0;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  printf("device %s sm_%d%d SMs=%d smem/block optin=%zu\n", prop.name, prop.major, prop.minor,
         prop.multiProcessorCount, prop.sharedMemPerBlockOptin);
  int rc = 0;
  for (int swap = 0; swap < 2; ++swap) {
    rc |= run(256, 256, 1, 0, swap);
    if (rc & 2) { cudaDeviceReset(); }
    rc |= run(256, 256, 1, 1, swap);
    if (rc & 2) { cudaDeviceReset(); }
  }
  rc |= run(256, 256, 3, 0, 0);
  rc |= run(256, 256, 3, 1, 0);
  rc |= run(128, 64, 3, 0, 0);
  rc |= run(256, 64, 3, 0, 0);
  rc |= run(64, 128, 3, 0, 0);
  rc |= run(16, 128, 3, 0, 0);
  return rc;
}
