// linear_tc05.cu — y = act(x W^T + b) for the layers AFTER the trunk on the training path:
// alpha / feature / view / rgb / the two heads, forward and input gradient (dL/dx = g W, i.e. the same product
// with the transposed matrix), on sm_100a tensor cores with the same 3-product operand split as the fused MLP
// kernel.  (The render path never comes here: there these layers are steps of the fused kernel's program.  On the
// training path they are differentiated one by one by autograd, so each is a GEMM of its own.)
//
//   y[s, n] = act( sum_k x[s, k] W[n, k] + b[n] ),   s < S (10^5 .. 10^7 samples),  K <= 512,  N <= 256
//
// Persistent, one CTA per SM, tiles of 128 samples (UMMA M = 128 = tensor-memory lanes):
//   * warps 4-11 (producers): a warp instruction reads two rows' 256-byte K chunk, 16 bytes per lane (fully
//     coalesced); the values are scaled (a power of two for gradients, exact), split into 16-bit hi / lo parts and
//     written as half core-matrix rows (8-byte stores) of the no-swizzle K-major A operand, one 64-feature K chunk per
//     stage (2 stages), the next chunk's loads in flight meanwhile.  The K-cores of the A images are 2064 bytes apart instead of 2048 (the descriptor's leading
//     byte offset is free to say so): the 16 lanes that share a row then hit 16 different 8-byte bank slots;
//   * warp 12: streams the matching K chunk of the weights - packed once per call by linear_pack_kernel into hi / lo
//     images of [8 K-cores][NP rows][16 B] - with cp.async.bulk into a 2-stage ring (64 KB stages, L2-resident source);
//   * warp 13: one elected lane issues tcgen05.mma kind::f16, M = 128, N = NP, K = 16, both operands from shared
//     memory, hi.hi + lo.hi + hi.lo, into one of TWO 256-column accumulators (tile parity);
//   * warps 0-3 (epilogue): tcgen05.ld the finished accumulator, * 1/scale, + bias, optional ReLU, 32-byte stores
//     (STG.256: one whole sector per lane; with 16-byte stores a third of the kernel's L2 sector traffic was half-used
//     sectors, ncu) to the sample's row of y - while the MMAs of the next tile fill the other accumulator.
// HBM-bound by construction for the shapes of this network (4 (K + N) bytes per sample against 6 K N tensor flops).
#include <cstddef>
#include <mutex>
#include "common.cuh"
#include "tc05.cuh"

namespace pnr {

constexpr int kLnEpiWarps = 4, kLnProWarps = 8;
constexpr int kLnThreads = (kLnEpiWarps + kLnProWarps + 2) * 32;   // + weight stream warp + MMA warp = 448
constexpr int kLnTile = 128;
constexpr int kLnChunk = 64;                             // K per stage
constexpr int kLnALbo = kLnTile * 16 + 16;               // distance of K-adjacent core matrices in the A images (padded: banks)
constexpr int kLnAPart = 8 * kLnALbo;                    // one 16-bit image of the A chunk
constexpr int kLnAStage = 2 * kLnAPart;                  // hi + lo
constexpr int kLnBStageMax = 2 * 8 * 256 * 16;           // hi + lo images of a [64 K][256 rows] weight chunk: 64 KB
constexpr int kLnRing = 2;
constexpr int kLnSmemA = 0;
constexpr int kLnSmemB = kLnRing * kLnAStage;            // 64 KB
constexpr int kLnSmemBars = kLnSmemB + kLnRing * kLnBStageMax;   // 192 KB
constexpr int kLnSmemBias = kLnSmemBars + 256;                  // [256] bias, zero-padded
constexpr int kLnSmemTotal = kLnSmemBias + 256 * 4;

struct LinearParams {
  const float* x; int64_t ld_x; int K;
  const uint8_t* wpk;     // packed weights: per 64-wide K chunk, hi image then lo image of [8][NP][16 B]
  const float* bias;      // [N] or NULL
  float* y; int64_t ld_y; int N;
  int64_t S;
  int NP, n_chunks, relu, vec_in, vec_out, vec8_out;
  const float* in_scale;  // device scalar (power of two) applied to x; the result is divided by it.  NULL: 1
};

// W [N, K] (row stride ld_w; `trans`: the matrix is given as [K, N] and read transposed) -> packed 16-bit images.
// One thread per (chunk, K-core, row): 8 consecutive K values of one row = one 16-byte core-matrix row.
template <int FMT>
__global__ void linear_pack_kernel(const float* __restrict__ W, int64_t ld_w, int N, int K, int NP, int n_chunks, int trans,
                                   uint8_t* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_chunks * 8 * NP) return;
  const int row = idx % NP, kc = (idx / NP) % 8, c = idx / (8 * NP);
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = c * kLnChunk + kc * 8 + j;
    v[j] = (row < N && k < K) ? (trans ? W[(int64_t)k * ld_w + row] : W[(int64_t)row * ld_w + k]) : 0.f;
  }
  uint32_t h[4], l[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) split_x2<FMT>(v[2 * q], v[2 * q + 1], h[q], l[q]);
  uint8_t* img = out + (size_t)c * (2 * 8 * NP * 16) + (size_t)(kc * NP + row) * 16;
  *reinterpret_cast<uint4*>(img) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(img + 8 * NP * 16) = make_uint4(l[0], l[1], l[2], l[3]);
}

template <int FMT>
__global__ void __launch_bounds__(kLnThreads, 1) linear_kernel(const LinearParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kLnSmemBars);
  const uint32_t bar_a_full = smem_u32(&bars[0]);      // [2] A chunk written (every producer thread arrives)
  const uint32_t bar_a_empty = smem_u32(&bars[2]);     // [2] consumed (tcgen05.commit)
  const uint32_t bar_b_full = smem_u32(&bars[4]);      // [2] weight chunk landed (complete_tx)
  const uint32_t bar_b_empty = smem_u32(&bars[6]);     // [2] consumed (tcgen05.commit)
  const uint32_t bar_acc_full = smem_u32(&bars[8]);    // [2] tile's accumulator complete (tcgen05.commit)
  const uint32_t bar_acc_empty = smem_u32(&bars[10]);  // [2] drained (every epilogue thread arrives)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(&bars[12]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_tiles = (p.S + kLnTile - 1) / kLnTile;
  const int n_iter = (int)((n_tiles - (int64_t)blockIdx.x + (int64_t)gridDim.x - 1) / (int64_t)gridDim.x);   // >= 1
  const int n_chunks = p.n_chunks;
  const uint32_t b_stage_bytes = (uint32_t)(2 * 8 * p.NP * 16);

  float* bias_s = reinterpret_cast<float*>(smem + kLnSmemBias);
  for (int n = threadIdx.x; n < 256; n += blockDim.x) bias_s[n] = (p.bias != nullptr && n < p.N) ? p.bias[n] : 0.f;
  if (warp == 0) {
    tmem_alloc<512>(smem_u32(tmem_slot));
    tmem_relinquish();
  }
  if (threadIdx.x == 32) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_a_full + 8 * s, kLnProWarps * 32);
      mbar_init(bar_a_empty + 8 * s, 1);
      mbar_init(bar_b_full + 8 * s, 1);
      mbar_init(bar_b_empty + 8 * s, 1);
      mbar_init(bar_acc_full + 8 * s, 1);
      mbar_init(bar_acc_empty + 8 * s, kLnEpiWarps * 32);
    }
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const float sc = p.in_scale != nullptr ? __ldg(p.in_scale) : 1.0f;

  if (warp < kLnEpiWarps) {
    // =============================================================== epilogue: accumulator -> y rows
    const int q = warp;                                  // tensor-memory lane quarter
    const float inv = 1.0f / sc;
    const int ngroups = p.NP >> 4;
    for (int it = 0; it < n_iter; ++it) {
      const int b = it & 1;
      const int64_t s = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * kLnTile + q * 32 + lane;
      mbar_wait_backoff(bar_acc_full + 8 * b, (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      float* dst = p.y + (s < p.S ? s : 0) * p.ld_y;
      uint32_t ra[16], rb[16];
      tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * 256), ra);
#pragma unroll 1
      for (int g = 0; g < ngroups; g += 2) {
        const bool two = g + 1 < ngroups;
        tc_wait_ld();
        if (two) tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * 256 + (g + 1) * 16), rb);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          if (half == 1) {
            if (!two) break;
            tc_wait_ld();
            if (g + 2 < ngroups) tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * 256 + (g + 2) * 16), ra);
          }
          const uint32_t (&r)[16] = half == 0 ? ra : rb;
          const int c0 = (g + half) * 16;
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float t = __uint_as_float(r[j]) * inv + bias_s[c0 + j];
            v[j] = p.relu ? fmaxf(t, 0.f) : t;
          }
          if (s < p.S) {
            if (p.vec8_out && c0 + 16 <= p.N) {
              st_global_v8(dst + c0, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
              st_global_v8(dst + c0 + 8, v[8], v[9], v[10], v[11], v[12], v[13], v[14], v[15]);
            } else if (p.vec_out && c0 + 16 <= p.N) {
              float4* d4 = reinterpret_cast<float4*>(dst + c0);
#pragma unroll
              for (int c = 0; c < 4; ++c) d4[c] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (c0 + j < p.N) dst[c0 + j] = v[j];
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar_acc_empty + 8 * b);
    }
  } else if (warp < kLnEpiWarps + kLnProWarps) {
    // =============================================================== producers: x rows -> A chunk images
    // warp w: rows 16 w .. 16 w + 15 of the tile, two rows per instruction; lane: row parity (lane >> 4), 16-byte
    // segment of the row's 256-byte chunk (lane & 15) = features 4 seg .. 4 seg + 3 = half (seg & 1) of K-core seg >> 1
    const int pw = warp - kLnEpiWarps;
    const int seg = lane & 15, kc = seg >> 1, half = seg & 1;
    const bool k_ragged = (p.K & 3) != 0;
    const int n_total = n_iter * n_chunks;               // chunks this CTA converts, tile after tile
    // loads of chunk (it, c): this lane's 16-byte segment of 8 rows (two rows apart), predicated per lane
    auto load_chunk = [&](int it, int c, float4 (&v)[8]) {
      const int64_t s0 = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * kLnTile + pw * 16 + (lane >> 4);
      const int64_t left = p.S - s0;                       // this lane's rows are s0 + 2 i: how many of them exist
      const int nvr = left <= 0 ? 0 : (left >= 16 ? 8 : (int)((left + 1) >> 1));
      const int k = c * kLnChunk + seg * 4;
      const int kr = p.K - k;                              // features of this lane's segment that exist (<= 0: none)
      const float* q = p.x + s0 * p.ld_x + k;              // (never dereferenced when nvr == 0 or kr <= 0)
      const int64_t step = 2 * p.ld_x;
      if (p.vec_in) {        // 16-byte loads: rows 16-byte aligned and at least round-up-4(K) floats long
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (kr > 0 && i < nvr) v[i] = __ldg(reinterpret_cast<const float4*>(q));
          q += step;
        }
        if (k_ragged && c + 1 == n_chunks && kr > 0 && kr < 4) {   // the segment that straddles K: drop what lies beyond
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (kr < 2) v[i].y = 0.f;
            if (kr < 3) v[i].z = 0.f;
            v[i].w = 0.f;
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (i < nvr) {
            if (kr > 0) v[i].x = __ldg(q);
            if (kr > 1) v[i].y = __ldg(q + 1);
            if (kr > 2) v[i].z = __ldg(q + 2);
            if (kr > 3) v[i].w = __ldg(q + 3);
          }
          q += step;
        }
      }
    };
    // L2 prefetch of the chunk after the one being loaded (registers hold one chunk in flight per thread; a prefetch
    // costs none): the warp's 16 rows x 256 bytes are 32 lines, one per lane
    auto prefetch_chunk = [&](int it, int c) {
      const int64_t s = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * kLnTile + pw * 16 + (lane >> 1);
      const int k = c * kLnChunk + (lane & 1) * 32;
      if (s < p.S && k < p.K) prefetch_l2(p.x + s * p.ld_x + k);
    };
    int it_n = 0, c_n = 0;                                 // the chunk the next load_chunk fetches
    auto advance = [&]() { if (++c_n == n_chunks) { c_n = 0; ++it_n; } };
    // convert + store chunk g from `cur` while chunk g + 1 is on its way into `nxt`
    auto do_chunk = [&](int g, float4 (&cur)[8], float4 (&nxt)[8]) {
      const uint32_t slot = (uint32_t)g & 1u, ph = ((uint32_t)g >> 1) & 1u;
      if (g + 1 < n_total) { load_chunk(it_n, c_n, nxt); advance(); }
      if (g + 2 < n_total) prefetch_chunk(it_n, c_n);
      mbar_wait_backoff(bar_a_empty + 8 * slot, ph ^ 1u);
      uint8_t* img = smem + kLnSmemA + slot * kLnAStage + kc * kLnALbo + (pw * 16 + (lane >> 4)) * 16 + half * 8;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uint32_t h0, l0, h1, l1;
        split_x2<FMT>(cur[i].x * sc, cur[i].y * sc, h0, l0);
        split_x2<FMT>(cur[i].z * sc, cur[i].w * sc, h1, l1);
        *reinterpret_cast<uint2*>(img + i * 32) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(img + i * 32 + kLnAPart) = make_uint2(l0, l1);
      }
      fence_proxy_async_smem();
      mbar_arrive(bar_a_full + 8 * slot);
    };
    float4 va[8], vb[8];     // two register buffers, used alternately
    load_chunk(0, 0, va);
    advance();
    if (n_total > 1) prefetch_chunk(it_n, c_n);
#pragma unroll 1
    for (int g = 0; g < n_total; g += 2) {
      do_chunk(g, va, vb);
      if (g + 1 < n_total) do_chunk(g + 1, vb, va);
    }
  } else if (warp == kLnEpiWarps + kLnProWarps) {
    // =============================================================== weight stream (one elected thread)
    if (elect_one()) {
      uint32_t gb = 0;
      for (int it = 0; it < n_iter; ++it) {
        for (int c = 0; c < n_chunks; ++c, ++gb) {
          const uint32_t slot = gb & 1u, ph = (gb >> 1) & 1u;
          mbar_wait_backoff(bar_b_empty + 8 * slot, ph ^ 1u);
          mbar_arrive_expect_tx(bar_b_full + 8 * slot, b_stage_bytes);
          bulk_g2s(smem_u32(smem + kLnSmemB + slot * kLnBStageMax), p.wpk + (size_t)c * b_stage_bytes, b_stage_bytes,
                   bar_b_full + 8 * slot);
        }
      }
    }
  } else {
    // =============================================================== MMA issuer
    const uint32_t idesc = make_idesc_f32acc(128, p.NP, FMT);
    const uint32_t b_lbo = (uint32_t)p.NP * 16u, b_part = 8u * (uint32_t)p.NP * 16u;
    uint32_t g = 0;
    for (int it = 0; it < n_iter; ++it) {
      const int b = it & 1;
      mbar_wait(bar_acc_empty + 8 * b, (uint32_t)(((it >> 1) & 1) ^ 1));   // the epilogue has drained this accumulator
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < n_chunks; ++c, ++g) {
        const uint32_t slot = g & 1u, ph = (g >> 1) & 1u;
        mbar_wait(bar_a_full + 8 * slot, ph);
        mbar_wait(bar_b_full + 8 * slot, ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + kLnSmemA + slot * kLnAStage);
          const uint32_t sb = smem_u32(smem + kLnSmemB + slot * kLnBStageMax);
          const uint32_t d_tmem = tmem + (uint32_t)(b * 256);
          const int kleft = p.K - c * kLnChunk;
          const int ksteps = kleft >= kLnChunk ? kLnChunk / 16 : (kleft + 15) / 16;
#pragma unroll 1
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint32_t a = sa + (uint32_t)ks * 2u * kLnALbo, bb = sb + (uint32_t)ks * 2u * b_lbo;
            const uint64_t a_hi = make_smem_desc_noswz(a, kLnALbo, 128), a_lo = make_smem_desc_noswz(a + kLnAPart, kLnALbo, 128);
            const uint64_t b_hi = make_smem_desc_noswz(bb, b_lbo, 128), b_lo = make_smem_desc_noswz(bb + b_part, b_lbo, 128);
            mma_ss(d_tmem, a_hi, b_hi, idesc, (c == 0 && ks == 0) ? 0u : 1u);
            mma_ss(d_tmem, a_lo, b_hi, idesc, 1u);
            mma_ss(d_tmem, a_hi, b_lo, idesc, 1u);
          }
          tc_commit(bar_a_empty + 8 * slot);
          tc_commit(bar_b_empty + 8 * slot);
          if (c + 1 == n_chunks) tc_commit(bar_acc_full + 8 * b);
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

static bool g_ln_attr[kMaxDevices][2] = {};
static std::mutex g_ln_mutex;

template <int FMT>
static int linear_launch(const LinearParams& p, const float* W, int64_t ld_w, int trans, uint8_t* wpk, int dev, cudaStream_t st) {
  {
    std::lock_guard<std::mutex> lock(g_ln_mutex);
    bool& done = g_ln_attr[dev][FMT == kFmtBF16];
    if (!done) {
      PNR_CUDA(cudaFuncSetAttribute(linear_kernel<FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kLnSmemTotal));
      done = true;
    }
  }
  const int n_pack = p.n_chunks * 8 * p.NP;
  linear_pack_kernel<FMT><<<(n_pack + 127) / 128, 128, 0, st>>>(W, ld_w, p.N, p.K, p.NP, p.n_chunks, trans, wpk);
  PNR_LAUNCH_CHECK("linear_pack_kernel");
  const int64_t n_tiles = (p.S + kLnTile - 1) / kLnTile;
  const int sms = num_sms(dev);
  const int grid = (int)(n_tiles < sms ? n_tiles : sms);
  if (grid > 0) {
    linear_kernel<FMT><<<grid, kLnThreads, kLnSmemTotal, st>>>(p);
    PNR_LAUNCH_CHECK("linear_kernel");
  }
  return PNR_OK;
}

}  // namespace pnr

using namespace pnr;

extern "C" size_t pnr_linear_workspace_bytes(int32_t N, int32_t K) {
  if (N <= 0 || K <= 0 || N > 256 || K > 512) return 0;
  const int NP = (N + 15) / 16 * 16, n_chunks = (K + kLnChunk - 1) / kLnChunk;
  return (size_t)n_chunks * 2 * 8 * NP * 16;
}

extern "C" int pnr_linear(const float* x, int64_t ld_x, int32_t K, const float* W, int64_t ld_w, int32_t transposed,
                          const float* bias, int32_t N, int64_t S, int32_t relu, int32_t precision, const float* in_scale,
                          float* y, int64_t ld_y, void* workspace, size_t workspace_bytes, void* stream) {
  PNR_CHECK_ARG(x != nullptr && W != nullptr && y != nullptr, "pnr_linear: x, W and y are required");
  PNR_CHECK_ARG(N >= 1 && N <= 256 && K >= 1 && K <= 512, "pnr_linear: N = %d must be in [1, 256], K = %d in [1, 512]", N, K);
  PNR_CHECK_ARG(ld_x >= K && ld_y >= N && ld_w >= (transposed ? N : K), "pnr_linear: leading dimensions smaller than the widths");
  PNR_CHECK_ARG(S >= 0, "pnr_linear: S = %lld", (long long)S);
  PNR_CHECK_ARG(precision == PNR_PREC_BF16X3 || precision == PNR_PREC_FP16X3, "pnr_linear: x3 precisions only (got %d)", precision);
  const size_t need = pnr_linear_workspace_bytes(N, K);
  PNR_CHECK_ARG(workspace != nullptr && workspace_bytes >= need && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
                "pnr_linear: workspace of %zu bytes (16-byte aligned), %zu needed (pnr_linear_workspace_bytes)", workspace_bytes, need);
  if (S == 0) return PNR_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int dev = 0;
  PNR_CUDA(cudaGetDevice(&dev));
  PNR_CHECK_ARG(dev >= 0 && dev < kMaxDevices, "pnr_linear: device ordinal %d >= %d", dev, kMaxDevices);
  LinearParams p;
  p.x = x; p.ld_x = ld_x; p.K = K;
  p.wpk = static_cast<const uint8_t*>(workspace);
  p.bias = bias;
  p.y = y; p.ld_y = ld_y; p.N = N;
  p.S = S;
  p.NP = (N + 15) / 16 * 16;
  p.n_chunks = (K + kLnChunk - 1) / kLnChunk;
  p.relu = relu != 0;
  p.vec_in = (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (ld_x & 3) == 0 && ld_x >= (K + 3) / 4 * 4;
  p.vec_out = (reinterpret_cast<uintptr_t>(y) & 15) == 0 && (ld_y & 3) == 0;
  p.vec8_out = (reinterpret_cast<uintptr_t>(y) & 31) == 0 && (ld_y & 7) == 0;
  p.in_scale = in_scale;
  uint8_t* wpk = static_cast<uint8_t*>(workspace);
  return precision == PNR_PREC_FP16X3 ? linear_launch<kFmtF16>(p, W, ld_w, transposed != 0, wpk, dev, st)
                                      : linear_launch<kFmtBF16>(p, W, ld_w, transposed != 0, wpk, dev, st);
}
