// wgrad_tc05.cu — weight-gradient GEMM of the MLP backward on sm_100a tensor cores.
//
//   dW[o, i] = sum_s dZ[s, o] * X[s, i]        db[o] = sum_s dZ[s, o]        (s over S samples, S ~ 10^5 .. 10^7)
//
// dZ [S, No] and X [S, Ni] are the fp32 operands the fused backward kernel stashes (mlp_tc05.cu, BWD programs) or
// the inputs / output gradients of the layers after the trunk; No, Ni <= 256.  The reduction runs over the SAMPLES,
// so this is a split-K GEMM with a tiny output: every CTA (one per SM) owns every 148th slab of 32 samples,
// accumulates its 256 x 256 partial product in TENSOR MEMORY over all of its slabs (two M = 128 accumulators of
// N <= 256 columns = all 512 columns), and writes it out once; a second kernel adds the partials in CTA order
// (deterministic: no atomics).  Per slab:
//   * 16 warps, one work item each: (operand, half of its 256 feature rows, 8 of the slab's 32 samples).  A lane reads
//     4 consecutive features of 8 consecutive samples - per sample row one fully coalesced 512-byte warp request of
//     16-byte loads (ncu: with 4-byte loads the same bytes in flight gave 34 % of the DRAM peak; the limit is requests
//     in flight, not bytes) - splits each value into 16-bit hi / lo parts and writes them as 16-byte rows of the UMMA
//     no-swizzle K-major core matrices (K = samples): a thread's 8 samples of one feature ARE one core-matrix row, so
//     the transposition dZ -> dZ^T costs nothing.  The 8-row groups of the images are 144 bytes apart instead of 128
//     (the descriptor's stride byte offset is free to say so): the four feature rows of a lane and those of the
//     lanes next to it then fall into different banks.  The next slab's loads are in flight while a slab is converted
//     (two register buffers used alternately), and the slab after that is prefetched into L2: one slab per thread in
//     registers is 64 KB in flight per SM, not enough at ~2 us of loaded DRAM latency, and a prefetch costs no register.
//   * one elected lane of warp 0 (after its own share of the slab) issues tcgen05.mma kind::f16 (fp16 or bf16 operands,
//     fp32 accumulate), both operands from shared memory (512 threads = 128 registers each; a 17th warp would cost 32):
//     A = dZ^T (M = 128 output features x K = 16 samples), B = X^T (N = Ni padded to 16, K-major), three products
//     per K step: hi.hi + lo.hi + hi.lo.  bf16 parts: ~2^-17 per product, fp32 exponent range, gradients need no
//     scaling.  fp16 parts: ~2^-21 per product; dZ is multiplied by a caller-supplied power of two on load (a device
//     scalar, exact) so that the parts of ~1e-6 gradients stay normal, and the sums are divided by it at the end.
//   * a 3-deep ring of 72 KB stages (A hi, A lo, B hi, B lo images of [4 K-cores][32 row groups][144 B]); full /
//     empty mbarriers, the empty ones arrived by tcgen05.commit.
// The kernel is HBM-bound by construction: 2 KB of fp32 operands per sample against 2 * 3 * 256 * 256 tensor flops
// (~50 tensor-pipe cycles per sample and SM): algorithmic bytes = 4 (No + Ni) per sample.  Measured (B200, 256 x 256,
// profiles/r02_wgrad_final_ncu.txt): 4.1 TB/s of operand reads at 393 k samples, 4.8 TB/s at 4 M (62 % / 73 % of the
// measured 6.58 TB/s copy peak); DRAM traffic = the algorithmic bytes.
#include <cstddef>
#include <mutex>
#include "common.cuh"
#include "tc05.cuh"

namespace pnr {

constexpr int kWgProWarps = 16;
constexpr int kWgThreads = kWgProWarps * 32;
constexpr int kWgSlab = 32;                         // samples per stage
constexpr int kWgRows = 256;                        // rows of one operand image
constexpr int kWgSbo = 128 + 16;                    // 8-row groups: 128 bytes of core matrix + 16 bytes of padding (banks)
constexpr uint32_t kWgLbo = (kWgRows / 8) * kWgSbo; // K-adjacent core matrices: 4608 bytes
constexpr int kWgPart = (kWgSlab / 8) * kWgLbo;     // one 16-bit image: [4 K-cores][32 row groups][144 B] = 18 KB
constexpr int kWgStage = 4 * kWgPart;               // A hi, A lo, B hi, B lo
constexpr int kWgRing = 3;
constexpr int kWgSmemBars = kWgRing * kWgStage;     // mbarriers + tensor-memory slot
constexpr int kWgSmemDb = kWgSmemBars + 128;        // [4 sample groups][256] partial bias sums
constexpr int kWgSmemTotal = kWgSmemDb + 4 * kWgRows * 4;

struct WgradParams {
  const float* dz; int64_t ld_dz; int No;
  const float* x; int64_t ld_x; int Ni;
  int64_t S;
  int vec_a, vec_b;   // 16-byte loads possible (base and row stride 16-byte aligned)
  int mh;        // M halves of 128 output features
  int NP;        // Ni padded to a multiple of 16
  float* part;   // [grid][mh * 128][NP]
  float* dbp;    // [grid][256] or NULL
  const float* a_scale;   // device scalar (power of two) applied to dZ on load; NULL: 1
};

template <int FMT>
__global__ void __launch_bounds__(kWgThreads, 1) wgrad_kernel(const WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWgSmemBars);
  const uint32_t bar_full = smem_u32(&bars[0]);             // [kWgRing] slab written (every producer thread arrives)
  const uint32_t bar_empty = smem_u32(&bars[kWgRing]);      // [kWgRing] slab consumed (tcgen05.commit)
  const uint32_t bar_done = smem_u32(&bars[2 * kWgRing]);   // all MMAs of this CTA complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(&bars[2 * kWgRing + 1]);
  float* dbs = reinterpret_cast<float*>(smem + kWgSmemDb);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_slabs = (p.S + kWgSlab - 1) / kWgSlab;
  const int n_mine = (int)((n_slabs - (int64_t)blockIdx.x + (int64_t)gridDim.x - 1) / (int64_t)gridDim.x);   // >= 1: grid <= n_slabs

  if (warp == 0) {
    tmem_alloc<512>(smem_u32(tmem_slot));
    tmem_relinquish();
  }
  if (threadIdx.x == 32) {
    for (int s = 0; s < kWgRing; ++s) {
      mbar_init(bar_full + 8 * s, kWgProWarps * 32);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_done, 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  {
    // =============================================================== producers: fp32 rows -> 16-bit hi / lo core-matrix rows
    // warp w: operand w >> 3 (0 = A = dZ, 1 = B = X), feature half (w >> 2) & 1, sample group w & 3 of the slab.  The
    // mapping is static, so a thread's partial bias sums stay in registers over the whole launch.
    const int kg = warp & 3, fh = (warp >> 2) & 1;
    const bool isA = warp < 8;
    const int F = isA ? p.No : p.Ni;                                   // the operand's width
    const bool on = isA ? fh < p.mh : fh * 128 < p.NP;                 // this warp's rows exist in the MMA
    const bool vec = (isA ? p.vec_a : p.vec_b) != 0;
    const int f0 = fh * 128 + 4 * lane;                                // this lane's 4 features = 4 image rows
    const uint32_t idesc = make_idesc_f32acc(128, p.NP, FMT);
    const float sc = p.a_scale != nullptr ? __ldg(p.a_scale) : 1.0f;
    const int64_t ld = isA ? p.ld_dz : p.ld_x;
    // features past the operand's width: multiplied by 0 (their products only reach accumulator rows / columns nobody
    // reads); a lane wholly outside reads feature 0 instead (a valid address)
    float mul[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) mul[c] = (on && f0 + c < F) ? (isA ? sc : 1.0f) : 0.f;
    const float* src = (isA ? p.dz : p.x) + ((on && f0 < F) ? f0 : 0);
    // One slab's loads: this lane's 4 features of 8 consecutive samples, predicated on a warp-uniform count (only the
    // launch's very last slab is partial).
    auto load_slab = [&](int i, float4 (&v)[8]) {
      const int64_t s0 = ((int64_t)blockIdx.x + (int64_t)i * gridDim.x) * kWgSlab + kg * 8;
      const int64_t left = p.S - s0;
      const int nv = !on ? 0 : (left >= 8 ? 8 : (left > 0 ? (int)left : 0));
      const float* q = src + s0 * ld;
      if (vec) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (k < nv) v[k] = __ldg(reinterpret_cast<const float4*>(q));
          q += ld;
        }
      } else {      // rows not 16-byte aligned (odd widths): 4-byte loads; columns past the width are not touched
        const int nc = F - f0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (k < nv) {
            v[k].x = __ldg(q);
            if (nc > 1) v[k].y = __ldg(q + 1);
            if (nc > 2) v[k].z = __ldg(q + 2);
            if (nc > 3) v[k].w = __ldg(q + 3);
          }
          q += ld;
        }
      }
    };
    // L2 prefetch of slab i (two slabs ahead of the one being converted): registers hold one slab in flight per thread,
    // which at ~2 us of loaded DRAM latency is not enough bytes in flight to fill the memory system; a prefetch costs
    // no register.  The warp's 8 rows x 512 bytes are 32 lines: one per lane.
    const float* pf_src = (isA ? p.dz : p.x) + fh * 128 + (lane & 3) * 32;
    auto prefetch_slab = [&](int i) {
      const int64_t s = ((int64_t)blockIdx.x + (int64_t)i * gridDim.x) * kWgSlab + kg * 8 + (lane >> 2);
      if (on && s < p.S && fh * 128 + (lane & 3) * 32 < F) prefetch_l2(pf_src + s * ld);
    };
    float db[4] = {0.f, 0.f, 0.f, 0.f};
    // convert + store slab i from `cur` while slab i + 1 is on its way into `nxt`; warp 0 then issues the slab's MMAs
    auto do_slab = [&](int i, float4 (&cur)[8], float4 (&nxt)[8]) {
      const uint32_t slot = (uint32_t)i % kWgRing, ph = ((uint32_t)i / kWgRing) & 1u;
      if (i + 1 < n_mine) load_slab(i + 1, nxt);
      if (i + 2 < n_mine) prefetch_slab(i + 2);
      mbar_wait_backoff(bar_empty + 8 * slot, ph ^ 1u);
      uint8_t* stage = smem + slot * kWgStage;
      if (on) {
        // row r = f0 + c of the image: 8-row group r / 8 at kWgSbo bytes each, 16 bytes per row inside it
        uint8_t* img = stage + (isA ? 0 : 2 * kWgPart) + kg * kWgLbo + (f0 >> 3) * kWgSbo + (f0 & 7) * 16;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float m = mul[c];
          float x[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) x[k] = (c == 0 ? cur[k].x : c == 1 ? cur[k].y : c == 2 ? cur[k].z : cur[k].w);
          uint32_t h[4], l[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) split_x2<FMT>(x[2 * q] * m, x[2 * q + 1] * m, h[q], l[q]);
          *reinterpret_cast<uint4*>(img + c * 16) = make_uint4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<uint4*>(img + c * 16 + kWgPart) = make_uint4(l[0], l[1], l[2], l[3]);
          if (isA) db[c] += ((x[0] + x[1]) + (x[2] + x[3])) + ((x[4] + x[5]) + (x[6] + x[7]));
        }
      }
      fence_proxy_async_smem();            // generic-proxy stores -> visible to the tensor core's async proxy
      mbar_arrive(bar_full + 8 * slot);
      if (warp == 0) {
        // ---- MMA issue for this slab (the next slab's loads of this warp are already in flight)
        mbar_wait(bar_full + 8 * slot, ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(stage);
          for (int h = 0; h < p.mh; ++h) {
            const uint32_t d_tmem = tmem + (uint32_t)(h * 256);
#pragma unroll
            for (int ks = 0; ks < kWgSlab / 16; ++ks) {
              const uint32_t a = sa + (uint32_t)(h * 16 * kWgSbo) + (uint32_t)ks * 2u * kWgLbo;
              const uint32_t b = sa + 2u * kWgPart + (uint32_t)ks * 2u * kWgLbo;
              const uint64_t a_hi = make_smem_desc_noswz(a, kWgLbo, kWgSbo), a_lo = make_smem_desc_noswz(a + kWgPart, kWgLbo, kWgSbo);
              const uint64_t b_hi = make_smem_desc_noswz(b, kWgLbo, kWgSbo), b_lo = make_smem_desc_noswz(b + kWgPart, kWgLbo, kWgSbo);
              mma_ss(d_tmem, a_hi, b_hi, idesc, (i == 0 && ks == 0) ? 0u : 1u);
              mma_ss(d_tmem, a_lo, b_hi, idesc, 1u);
              mma_ss(d_tmem, a_hi, b_lo, idesc, 1u);
            }
          }
          tc_commit(bar_empty + 8 * slot);             // the slab's images may be overwritten once these MMAs retire
          if (i + 1 == n_mine) tc_commit(bar_done);
        }
        __syncwarp();
      }
    };
    float4 va[8], vb[8];    // two register buffers, used alternately (no copies between iterations)
    load_slab(0, va);
    if (n_mine > 1) prefetch_slab(1);
#pragma unroll 1
    for (int i = 0; i < n_mine; i += 2) {
      do_slab(i, va, vb);
      if (i + 1 < n_mine) do_slab(i + 1, vb, va);
    }
    // ---- bias partial sums: 4 sample groups per feature, added in a fixed order
    if (on && isA) {
#pragma unroll
      for (int c = 0; c < 4; ++c) dbs[kg * kWgRows + f0 + c] = db[c];
    }
    __syncthreads();
    if (p.dbp != nullptr && (int)threadIdx.x < p.mh * 128) {
      const int f = threadIdx.x;
      p.dbp[(int64_t)blockIdx.x * kWgRows + f] = ((dbs[f] + dbs[kWgRows + f]) + dbs[2 * kWgRows + f]) + dbs[3 * kWgRows + f];
    }
    // ---- drain: accumulators -> this CTA's partial product
    mbar_wait_backoff(bar_done, 0u);
    tc_fence_after();
    const int q = warp & 3, cs = warp >> 2;     // tensor-memory lane quarter ; share of the 16-column groups
    const int ngroups = p.NP >> 4;
    for (int h = 0; h < p.mh; ++h) {
      const int o = h * 128 + q * 32 + lane;
      float* dst = p.part + ((int64_t)blockIdx.x * (p.mh * 128) + o) * p.NP;
      for (int g = cs; g < ngroups; g += 4) {
        uint32_t r[16];
        tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 256 + g * 16), r);
        tc_wait_ld();
        float4* d4 = reinterpret_cast<float4*>(dst + g * 16);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          d4[c] = make_float4(__uint_as_float(r[4 * c]), __uint_as_float(r[4 * c + 1]), __uint_as_float(r[4 * c + 2]),
                              __uint_as_float(r[4 * c + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// dW[o, i] (+)= sum over the CTAs' partial products, in CTA order; db likewise.
__global__ void wgrad_reduce_kernel(const float* __restrict__ part, const float* __restrict__ dbp, int G, int rows, int NP,
                                    int No, int Ni, float* __restrict__ dW, int64_t ld_w, float* __restrict__ db,
                                    int accumulate, const float* __restrict__ a_scale) {
  const float inv = a_scale != nullptr ? 1.0f / __ldg(a_scale) : 1.0f;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < No * NP) {
    const int o = idx / NP, i = idx - o * NP;
    if (i < Ni) {
      const float* src = part + (int64_t)o * NP + i;
      const int64_t stride = (int64_t)rows * NP;
      float s = 0.f;
      int c = 0;
      for (; c + 4 <= G; c += 4) {
        const float v0 = src[(c + 0) * stride], v1 = src[(c + 1) * stride], v2 = src[(c + 2) * stride], v3 = src[(c + 3) * stride];
        s = (((s + v0) + v1) + v2) + v3;
      }
      for (; c < G; ++c) s += src[c * stride];
      float* d = dW + (int64_t)o * ld_w + i;
      s *= inv;
      *d = accumulate ? *d + s : s;
    }
  }
  if (db != nullptr && dbp != nullptr && idx < No) {
    float s = 0.f;
    for (int c = 0; c < G; ++c) s += dbp[(int64_t)c * kWgRows + idx];
    db[idx] = accumulate ? db[idx] + s : s;
  }
}

static bool g_wg_attr[kMaxDevices][2] = {};
static std::mutex g_wg_mutex;

static int wgrad_grid(int64_t S, int dev) {
  const int64_t n_slabs = (S + kWgSlab - 1) / kWgSlab;
  const int sms = num_sms(dev);
  return (int)(n_slabs < sms ? n_slabs : sms);
}

}  // namespace pnr

using namespace pnr;

extern "C" size_t pnr_wgrad_workspace_bytes(int32_t No, int32_t Ni) {
  if (No <= 0 || Ni <= 0 || No > 256 || Ni > 256) return 0;
  const int mh = No > 128 ? 2 : 1, NP = (Ni + 15) / 16 * 16;
  return (size_t)num_sms() * ((size_t)mh * 128 * NP + kWgRows) * sizeof(float);
}

template <int FMT>
static int wgrad_launch(const WgradParams& p, int grid, int dev, cudaStream_t st) {
  {
    std::lock_guard<std::mutex> lock(g_wg_mutex);
    bool& done = g_wg_attr[dev][FMT == kFmtBF16];
    if (!done) {
      PNR_CUDA(cudaFuncSetAttribute(wgrad_kernel<FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemTotal));
      done = true;
    }
  }
  wgrad_kernel<FMT><<<grid, kWgThreads, kWgSmemTotal, st>>>(p);
  PNR_LAUNCH_CHECK("wgrad_kernel");
  return PNR_OK;
}

extern "C" int pnr_wgrad(const float* dz, int64_t ld_dz, int32_t No, const float* x, int64_t ld_x, int32_t Ni, int64_t S,
                         int32_t precision, const float* dz_scale, float* dW, int64_t ld_w, float* db, int32_t accumulate,
                         void* workspace, size_t workspace_bytes, void* stream) {
  PNR_CHECK_ARG(precision == PNR_PREC_BF16X3 || precision == PNR_PREC_FP16X3, "pnr_wgrad: x3 precisions only (got %d)", precision);
  PNR_CHECK_ARG(dz != nullptr && x != nullptr && dW != nullptr, "pnr_wgrad: dz, x and dW are required");
  PNR_CHECK_ARG(No >= 1 && No <= 256 && Ni >= 1 && Ni <= 256, "pnr_wgrad: No = %d, Ni = %d must be in [1, 256] (split wider layers by columns)", No, Ni);
  PNR_CHECK_ARG(ld_dz >= No && ld_x >= Ni && ld_w >= Ni, "pnr_wgrad: leading dimensions smaller than the widths");
  PNR_CHECK_ARG(S >= 0, "pnr_wgrad: S = %lld", (long long)S);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int dev = 0;
  PNR_CUDA(cudaGetDevice(&dev));
  PNR_CHECK_ARG(dev >= 0 && dev < kMaxDevices, "pnr_wgrad: device ordinal %d >= %d", dev, kMaxDevices);
  const int mh = No > 128 ? 2 : 1, NP = (Ni + 15) / 16 * 16;
  const int grid = S > 0 ? wgrad_grid(S, dev) : 0;
  const size_t need = (size_t)grid * ((size_t)mh * 128 * NP + kWgRows) * sizeof(float);
  PNR_CHECK_ARG(grid == 0 || (workspace != nullptr && workspace_bytes >= need),
                "pnr_wgrad: workspace of %zu bytes, %zu needed (pnr_wgrad_workspace_bytes)", workspace_bytes, need);
  PNR_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "pnr_wgrad: workspace must be 16-byte aligned");
  WgradParams p;
  p.dz = dz; p.ld_dz = ld_dz; p.No = No;
  p.x = x; p.ld_x = ld_x; p.Ni = Ni;
  p.S = S; p.mh = mh; p.NP = NP;
  p.part = static_cast<float*>(workspace);
  p.dbp = db != nullptr ? p.part + (size_t)grid * mh * 128 * NP : nullptr;
  p.a_scale = dz_scale;
  p.vec_a = (reinterpret_cast<uintptr_t>(dz) & 15) == 0 && (ld_dz & 3) == 0;
  p.vec_b = (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (ld_x & 3) == 0;
  if (grid > 0) {
    const int rc = precision == PNR_PREC_FP16X3 ? wgrad_launch<kFmtF16>(p, grid, dev, st) : wgrad_launch<kFmtBF16>(p, grid, dev, st);
    if (rc != PNR_OK) return rc;
  }
  const int n = No * NP;
  wgrad_reduce_kernel<<<(n + 255) / 256, 256, 0, st>>>(p.part, p.dbp, grid, mh * 128, NP, No, Ni, dW, ld_w, db, accumulate, dz_scale);
  PNR_LAUNCH_CHECK("wgrad_reduce_kernel");
  return PNR_OK;
}
