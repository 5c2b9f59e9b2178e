// mlp_tc05.cu — fused Network.forward (SURVEY.md 8(a) a7 + a8) on sm_100a tensor cores.
//
// One persistent CTA per SM walks 128-sample tiles.  Per tile the whole MLP runs on-chip:
//   * 4 producer warps form pts = o + d*z, the positional encodings gamma(x), gamma(d), split them into
//     16-bit hi/lo parts and write them into shared memory in the UMMA no-swizzle K-major operand layout;
//   * 1 TMA warp streams the pre-packed weight stages (<= 32 KB: hi image then lo image of a 128-row x 64-K
//     tile) from L2 into a 4-deep shared-memory ring with cp.async.bulk + mbarrier complete_tx; the two CTAs
//     of a cluster each fetch half of every stage and multicast it into both shared memories;
//   * 2 MMA warps take alternate stages; one elected lane issues tcgen05.mma kind::f16, M=128, N<=128, K=16:
//     the A operand is the embedding in shared memory (SS form) or the previous layer's activations
//     in TENSOR MEMORY (TS form); accumulators are fp32 in tensor memory;
//   * 1 scout thread does the mbarrier waits for "weight stage landed" / "embedding written" and publishes a
//     "stages ready" counter, so the issuers never execute an mbarrier wait;
//   * 8 epilogue warps tcgen05.ld the accumulator, add bias, ReLU, split into hi/lo and tcgen05.st the
//     result back to tensor memory as the next layer's A operand.  Activations never touch shared or
//     global memory.  The sigma head (N=1) and rgb head (N=3) are CUDA-core dot products inside the
//     epilogues; semantic / instance logits are written straight to `raw`.
//
// Overlap: every step (layer) is issued as two N-halves h0, h1 with separate accumulator columns.
// E0 (epilogue of h0) runs while the tensor pipe works on h1; E1 runs while the next step's h0 consumes
// the K-chunks E0 already produced.  Hazards:
//   MMA -> epilogue   acc_full[0/1]  mbarriers (tcgen05.commit), one phase per step;
//   MMA -> E0 stores  war_ok         mbarrier: the activation columns E0 overwrites have been read by the last
//                                    MMA of this step that needs them;
//   epilogue -> MMA   three monotonic shared-memory counters (E0 done, E1 part a done, E1 done), +1 per
//                     epilogue warp (red.release after tcgen05.wait::st + fence), polled with ld.acquire by
//                     the issuing warps: the hand-off costs one shared-memory round trip instead of
//                     256 mbarrier arrivals -> scout wake-up -> ready word -> issuer (~560 cycles, timeline v10).
//
// Precision: operands are 16-bit (fp16 or bf16), accumulation fp32.  The "x3" modes compute every product
// as A_hi*B_hi + A_lo*B_hi + A_hi*B_lo with x = hi + lo split in the operand format: ~2^-21 relative
// per product for fp16x3 (default; what the 1e-4 parity tolerance needs with margin), ~2^-17 for
// bf16x3 (fp32 exponent range).  The 1-pass modes keep the first term only (fast, out of tolerance).
// An activation outside the operand format's range (|x| > 65504 in the fp16 modes, non-finite in any) sets
// bit 0 of the context's sticky status word: overflow is reported (pnr_status), never silent.
#include <cstddef>
#include <mutex>
#include "common.cuh"
#include "composite_math.cuh"
#include "mlp_program.h"
#include "tc05.cuh"

namespace pnr {

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Write 8 consecutive K elements (one 16-byte core-matrix row) of this thread's row, split hi/lo.
template <int PASSES, int FMT>
__device__ __forceinline__ void store_core_row(uint8_t* hi_base, uint8_t* lo_base, int kcore, int row,
                                               const float (&v)[8]) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) split_x2<FMT>(v[2 * j], v[2 * j + 1], h[j], l[j]);
  const int off = (kcore * kTileM + row) * 16;
  *reinterpret_cast<uint4*>(hi_base + off) = make_uint4(h[0], h[1], h[2], h[3]);
  if (PASSES == 3) *reinterpret_cast<uint4*>(lo_base + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

// gamma(p) = [p, sin(2^0 p), cos(2^0 p), ...] padded with zeros to KPAD, streamed out 8 at a time.
template <int PASSES, int FMT, int LMAX, int KPAD>
__device__ __forceinline__ void encode_row(const float (&p)[3], int L, uint8_t* hi_base,
                                           uint8_t* lo_base, int row) {
  float v[KPAD];
#pragma unroll
  for (int i = 0; i < KPAD; ++i) v[i] = 0.f;
  v[0] = p[0]; v[1] = p[1]; v[2] = p[2];
  float f = 1.0f;
#pragma unroll
  for (int k = 0; k < LMAX; ++k) {
    if (k < L) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float sn, cs;
        sincosf(p[c] * f, &sn, &cs);
        if (3 + 6 * k + c < KPAD) v[3 + 6 * k + c] = sn;
        if (3 + 6 * k + 3 + c < KPAD) v[3 + 6 * k + 3 + c] = cs;
      }
    }
    f *= 2.0f;
  }
#pragma unroll
  for (int g = 0; g < KPAD / 8; ++g) {
    float w8[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) w8[j] = v[g * 8 + j];
    store_core_row<PASSES, FMT>(hi_base, lo_base, g, row, w8);
  }
}

// ------------------------------------------------------------------------------------------------
// Epilogue building blocks.  One thread owns one accumulator row (TMEM lane); groups are 16 columns.
// ------------------------------------------------------------------------------------------------

// activation -> next layer's A operand: v = act(acc + bias); [sigma += v . wsig]; split into hi / lo parts.
// `vmax` collects the largest hi-part bit patterns seen (two 16-bit lanes; range check of the operand format).
template <int PASSES, int FMT>
__device__ __forceinline__ void epi_group_act(const uint32_t (&r)[16], int g, const EpiDesc& ed,
                                              const float* bias, const float* wsig, float& sig, uint32_t& vmax,
                                              uint32_t (&hi)[8], uint32_t (&lo)[8]) {
  const float4* b4 = reinterpret_cast<const float4*>(bias + g * 16);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 b = b4[q];
    const float v0 = fmaxf(__uint_as_float(r[4 * q + 0]) + b.x, 0.f);
    const float v1 = fmaxf(__uint_as_float(r[4 * q + 1]) + b.y, 0.f);
    const float v2 = fmaxf(__uint_as_float(r[4 * q + 2]) + b.z, 0.f);
    const float v3 = fmaxf(__uint_as_float(r[4 * q + 3]) + b.w, 0.f);
    if (ed.sigma) {
      const float4 w = reinterpret_cast<const float4*>(wsig + g * 16)[q];
      sig += v0 * w.x + v1 * w.y + v2 * w.z + v3 * w.w;
    }
    split_x2<FMT>(v0, v1, hi[2 * q], lo[2 * q]);
    split_x2<FMT>(v2, v3, hi[2 * q + 1], lo[2 * q + 1]);
    // Range check on the packed hi parts (one 3-input 16x2 max per four values): v >= 0 after the ReLU, so the
    // 16-bit patterns order like the values, with +inf (what an overflowing conversion yields) and NaN on top.
    // (fmaxf turns a NaN accumulator into 0, but a NaN can only follow an overflow that was flagged where it
    // happened - which is why the check sits in every layer and not only on the outputs.)
#ifndef PNR_ABL_NOVMAX
    vmax = __vimax3_u16x2(vmax, hi[2 * q], hi[2 * q + 1]);
#endif
  }
}

template <int PASSES>
__device__ __forceinline__ void epi_group_store(int g, const EpiDesc& ed, uint32_t tmem_lane,
                                                const uint32_t (&hi)[8], const uint32_t (&lo)[8]) {
  tmem_st8(tmem_lane + ed.dst_col + g * 8, hi);
  if (PASSES == 3) tmem_st8(tmem_lane + ed.dst_lo_col + g * 8, lo);
}

__device__ __forceinline__ void epi_group_rgb(const uint32_t (&r)[16], int g, const EpiDesc& ed, const float* bias,
                                              const float* wr, float& c0, float& c1, float& c2) {
  // bias and the three weight rows come as 16-byte shared-memory loads (every offset is a multiple of 4 floats)
  const float4* b4 = reinterpret_cast<const float4*>(bias + g * 16);
  const float4* w0 = reinterpret_cast<const float4*>(wr + g * 16);
  const float4* w1 = reinterpret_cast<const float4*>(wr + ed.n + g * 16);
  const float4* w2 = reinterpret_cast<const float4*>(wr + 2 * ed.n + g * 16);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 b = b4[q], a0 = w0[q], a1 = w1[q], a2 = w2[q];
    const float v0 = fmaxf(__uint_as_float(r[4 * q + 0]) + b.x, 0.f);
    const float v1 = fmaxf(__uint_as_float(r[4 * q + 1]) + b.y, 0.f);
    const float v2 = fmaxf(__uint_as_float(r[4 * q + 2]) + b.z, 0.f);
    const float v3 = fmaxf(__uint_as_float(r[4 * q + 3]) + b.w, 0.f);
    c0 += v0 * a0.x; c1 += v0 * a1.x; c2 += v0 * a2.x;
    c0 += v1 * a0.y; c1 += v1 * a1.y; c2 += v1 * a2.y;
    c0 += v2 * a0.z; c1 += v2 * a1.z; c2 += v2 * a2.z;
    c0 += v3 * a0.w; c1 += v3 * a1.w; c2 += v3 * a2.w;
  }
}

// logits straight to the raw row.  `dst` / `n_valid` / `c0` describe the half the group lies in: channel of column c
// = (c - c0) relative to dst, real while < n_valid.
__device__ __forceinline__ void epi_group_logits(const uint32_t (&r)[16], int g, int c0, int n_valid,
                                                 const float* bias, float* dst) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int ch = g * 16 + j - c0;
    if (ch < n_valid) dst[ch] = __uint_as_float(r[j]) + bias[g * 16 + j];
  }
}

// ------------------------------------------------------------------------------------------------
// Backward programs (BWD kernels): the three hand-overs that write an A operand, and the gradient output.
//   EPI_RELU_TO_A   forward layer: v = relu(acc + bias); its sign pattern is saved (slot n_valid-1) for the way back
//   EPI_LOADG_TO_A  last forward layer: v = incoming gradient where acc + bias > 0 (relu' = 0 at 0, like autograd)
//   EPI_MASK_TO_A   backward layer: v = acc where the saved pattern of the layer below is set
// `mask_row` = this thread's row in slot 0 / group 0 of the pattern array; `gin` = this row of the incoming gradient.
// ------------------------------------------------------------------------------------------------
template <int PASSES, int FMT>
__device__ __forceinline__ void epi_group_bwd(const uint32_t (&r)[16], int g, const EpiDesc& ed, const float* bias,
                                              uint16_t* mask_row, const float* gin, float* stash_row, float gscale,
                                              float gunscale, uint32_t& vmax, uint32_t (&hi)[8], uint32_t (&lo)[8]) {
  float v[16];
  float sscale = gunscale;     // what the stash receives: gradients unscaled, forward activations as they are
  if (ed.kind == EPI_MASK_TO_A) {
    const uint32_t m = ed.n_valid ? mask_row[((ed.n_valid - 1) * 16 + g) * kTileM] : 0xFFFFu;
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = ((m >> j) & 1u) ? __uint_as_float(r[j]) : 0.f;
  } else {
    const float4* b4 = reinterpret_cast<const float4*>(bias + g * 16);
    uint32_t m = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 b = b4[q];
      v[4 * q + 0] = fmaxf(__uint_as_float(r[4 * q + 0]) + b.x, 0.f);
      v[4 * q + 1] = fmaxf(__uint_as_float(r[4 * q + 1]) + b.y, 0.f);
      v[4 * q + 2] = fmaxf(__uint_as_float(r[4 * q + 2]) + b.z, 0.f);
      v[4 * q + 3] = fmaxf(__uint_as_float(r[4 * q + 3]) + b.w, 0.f);
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) m |= (v[j] > 0.f ? 1u : 0u) << j;
    if (ed.kind == EPI_LOADG_TO_A) {
      const float4* g4 = reinterpret_cast<const float4*>(gin + g * 16);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 x = g4[q];
        v[4 * q + 0] = ((m >> (4 * q + 0)) & 1u) ? x.x * gscale : 0.f;
        v[4 * q + 1] = ((m >> (4 * q + 1)) & 1u) ? x.y * gscale : 0.f;
        v[4 * q + 2] = ((m >> (4 * q + 2)) & 1u) ? x.z * gscale : 0.f;
        v[4 * q + 3] = ((m >> (4 * q + 3)) & 1u) ? x.w * gscale : 0.f;
      }
    } else {
      sscale = 1.0f;
      if (ed.n_valid) mask_row[((ed.n_valid - 1) * 16 + g) * kTileM] = (uint16_t)m;
    }
  }
  if (stash_row != nullptr) {   // fp32 copy of the operand for the weight-gradient GEMMs
    float4* s4 = reinterpret_cast<float4*>(stash_row + g * 16);
#pragma unroll
    for (int q = 0; q < 4; ++q)
      s4[q] = make_float4(v[4 * q] * sscale, v[4 * q + 1] * sscale, v[4 * q + 2] * sscale, v[4 * q + 3] * sscale);
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    split_x2<FMT>(v[4 * q + 0], v[4 * q + 1], hi[2 * q], lo[2 * q]);
    split_x2<FMT>(v[4 * q + 2], v[4 * q + 3], hi[2 * q + 1], lo[2 * q + 1]);
    // gradients carry a sign: the range check compares magnitudes
    vmax = __vimax3_u16x2(vmax, hi[2 * q] & 0x7FFF7FFFu, hi[2 * q + 1] & 0x7FFF7FFFu);
  }
}

// the trunk's output activations: relu(acc + bias) of this group -> the sample's output row (16-byte stores)
__device__ __forceinline__ void epi_group_actout(const uint32_t (&r)[16], int g, const float* bias, float* dst) {
  const float4* b4 = reinterpret_cast<const float4*>(bias + g * 16);
  float4* d4 = reinterpret_cast<float4*>(dst + g * 16);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 b = b4[q];
    d4[q] = make_float4(fmaxf(__uint_as_float(r[4 * q + 0]) + b.x, 0.f), fmaxf(__uint_as_float(r[4 * q + 1]) + b.y, 0.f),
                        fmaxf(__uint_as_float(r[4 * q + 2]) + b.z, 0.f), fmaxf(__uint_as_float(r[4 * q + 3]) + b.w, 0.f));
  }
}

// gradient w.r.t. the embedded input: accumulator columns [0, n_valid) of this group -> the sample's output row
// (accumulating: all loads first - one memory latency per group, not one per column).  `vec`: the output rows are
// 16-byte aligned and padded to whole groups (row stride a multiple of 4 floats >= the step's width; the padding
// columns receive the zero-padded weights' zeros): four 16-byte accesses per group instead of sixteen 4-byte ones
// whose 32 lanes each touch a different sector.
__device__ __forceinline__ void epi_group_gradout(const uint32_t (&r)[16], int g, int n_valid, bool accumulate, bool vec,
                                                  float us, float* dst) {
  if (vec) {
    float4* d4 = reinterpret_cast<float4*>(dst + g * 16);
    float4 prev[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) prev[q] = accumulate ? d4[q] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int q = 0; q < 4; ++q)
      d4[q] = make_float4(prev[q].x + __uint_as_float(r[4 * q + 0]) * us, prev[q].y + __uint_as_float(r[4 * q + 1]) * us,
                          prev[q].z + __uint_as_float(r[4 * q + 2]) * us, prev[q].w + __uint_as_float(r[4 * q + 3]) * us);
    return;
  }
  float prev[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) prev[j] = (accumulate && g * 16 + j < n_valid) ? dst[g * 16 + j] : 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j)
    if (g * 16 + j < n_valid) dst[g * 16 + j] = prev[j] + __uint_as_float(r[j]) * us;
}

// ------------------------------------------------------------------------------------------------
// Compositing epilogue building blocks (COMP kernels).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum32(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}
// x[j] = this lane's (= this sample's) value of column j.  Returns the sum over the 32 lanes of column
// comp_col_of_lane(lane), by a fixed exchange tree (16 shuffles): lanes swap the half of the columns they give up.
__device__ __forceinline__ float comp_reduce16(const float (&x)[16], int lane) {
  float y[8], z[4], u[2];
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
#pragma unroll
  for (int j = 0; j < 8; ++j) y[j] = (b4 ? x[j + 8] : x[j]) + __shfl_xor_sync(0xffffffffu, b4 ? x[j] : x[j + 8], 16);
#pragma unroll
  for (int j = 0; j < 4; ++j) z[j] = (b3 ? y[j + 4] : y[j]) + __shfl_xor_sync(0xffffffffu, b3 ? y[j] : y[j + 4], 8);
#pragma unroll
  for (int j = 0; j < 2; ++j) u[j] = (b2 ? z[j + 2] : z[j]) + __shfl_xor_sync(0xffffffffu, b2 ? z[j] : z[j + 2], 4);
  float v = (b1 ? u[1] : u[0]) + __shfl_xor_sync(0xffffffffu, b1 ? u[0] : u[1], 2);
  return v + __shfl_xor_sync(0xffffffffu, v, 1);
}
__device__ __forceinline__ int comp_col_of_lane(int lane) { return (lane >> 1) & 15; }   // 8*b4 + 4*b3 + 2*b2 + b1

// logits of one 16-column group, weighted by this sample's compositing weight and summed over the warp's 32 samples
// (one aligned group of a ray); the even lanes write the per-quarter partial sums.
__device__ __forceinline__ void epi_group_logits_comp(const uint32_t (&r)[16], int g, int c0, int n_valid, int ch_base,
                                                      const float* bias, float w, int lane, float* qsum_q) {
  float x[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) x[j] = w * (__uint_as_float(r[j]) + bias[g * 16 + j]);
  const float tot = comp_reduce16(x, lane);
  const int ch = g * 16 + comp_col_of_lane(lane) - c0;
  if ((lane & 1) == 0 && ch < n_valid) qsum_q[ch_base + ch] = tot;
}

// Bit pattern of +inf in the operand format: a hi part >= this is an overflowed (or NaN) activation.
template <int FMT>
__device__ __forceinline__ constexpr uint32_t inf_bits16() {
  return FMT == kFmtF16 ? 0x7C00u : 0x7F80u;
}

// CTAs run as clusters of two that stream the SAME weight stages in lock step: each CTA fetches half of every
// stage from L2 and multicasts it into both shared memories, halving L2 -> SM weight traffic (ncu: lts
// throughput 28 % -> 16 %; the kernel time did not change - the weight stream is not what bounds it).
//
// COMP = false: tiles are dealt round-robin to the CTAs and the network outputs go to `raw`.
// COMP = true : every CTA owns a contiguous range of whole rays and walks it tile by tile; the epilogue warps
// composite on chip - per-sample alpha / transmittance / weight right after the sigma-producing layer (warp scan
// over aligned groups of 32 samples, carried across groups, tiles and warps through shared memory), logits and
// colours reduced per group with a fixed shuffle tree and accumulated per ray in ray order - and only weights and
// the per-ray maps leave the SM: `raw` (456 B per sample with both heads) is never written.
//
// BWD = true  : the program is a backward program (mlp_program.h): the forward trunk with its ReLU sign patterns kept
// in shared memory, then the layers in reverse with transposed weights; input `grad_in`, output `raw`.
//
// VP = true   : the program's view step (its last) has its epilogue on the producer warps (mlp_program.h).
template <int PASSES, int FMT, bool COMP, bool BWD = false, bool VP = false>
__global__ void __cluster_dims__(kClusterSize, 1, 1) __launch_bounds__(kMlpThreads, 1)
mlp_fused_kernel(const __grid_constant__ MlpLaunch L) {
  static_assert(!(COMP && BWD), "the compositing epilogue belongs to forward programs");
  static_assert(!(VP && (COMP || BWD)), "view-on-producers programs are plain forward programs");
  extern __shared__ __align__(1024) uint8_t smem[];
  const MlpParams& p = L.p;
  const MlpProgram& prog = L.prog;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  // every CTA runs the same number of tile iterations (tiles past the end are dummies whose stores are
  // masked) so that both CTAs of a cluster consume the shared weight stream the same number of times
  const int n_iter = COMP ? (int)((p.rays_per_cta * p.N + kTileM - 1) / kTileM)
                          : (p.num_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
  // first sample of this CTA's tile `it`, and the end of the samples it owns (nothing 64-bit is kept live in the
  // round-robin variant: everything is recomputed from blockIdx / gridDim / launch constants)
  auto cta_first = [&]() -> int64_t { return COMP ? (int64_t)blockIdx.x * p.rays_per_cta * p.N : 0; };
  auto tile_base = [&](int it) -> int64_t {
    return COMP ? cta_first() + (int64_t)it * kTileM : (int64_t)((int)blockIdx.x + it * (int)gridDim.x) * kTileM;
  };
  auto cta_end = [&]() -> int64_t {
    if (!COMP) return p.S;
    const int64_t e = cta_first() + p.rays_per_cta * p.N;
    return e < p.S ? e : p.S;
  };

  float* consts = reinterpret_cast<float*>(smem + kSmemConsts);
  // [2 column shares][128][4]: partial sigma / rgb sums.  Single-buffered across tiles: a share-1 warp writes the next
  // tile's sigma only in that tile's sigma step, whose MMAs waited for the previous step's counter, which every
  // share-0 warp bumped AFTER its reads of this tile's sums (program order + the fence in signal()).  racecheck does
  // not see that chain and reports a write-after-read here; double-buffering it costs registers for nothing.
  float* part = reinterpret_cast<float*>(smem + kSmemPart);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemBars);

  const uint32_t bar_full = smem_u32(&bars[0]);                 // [kRing]
  const uint32_t bar_empty = smem_u32(&bars[kRing]);            // [kRing]
  const uint32_t bar_acc_full = smem_u32(&bars[2 * kRing]);     // [2]
  const uint32_t bar_war = smem_u32(&bars[2 * kRing + 2]);      // E0 may store
  const uint32_t bar_emb_full = smem_u32(&bars[2 * kRing + 4]);
  const uint32_t bar_emb_empty = smem_u32(&bars[2 * kRing + 5]);
  const uint32_t bar_dir_full = smem_u32(&bars[2 * kRing + 6]);   // [2]
  const uint32_t bar_dir_empty = smem_u32(&bars[2 * kRing + 8]);  // [2]
  const uint32_t bar_view_full = smem_u32(&bars[2 * kRing + 10]); // VP: the view step's accumulator is complete
  // 32-bit words (each in its own 8-byte slot)
  const uint32_t cnt_e0 = smem_u32(&bars[24]);      // epilogue -> MMA counters: +1 per epilogue warp and part
  const uint32_t cnt_e1a = smem_u32(&bars[25]);
  const uint32_t cnt_e1 = smem_u32(&bars[26]);
  const uint32_t cnt_v = smem_u32(&bars[27]);       // VP: view epilogues done, +1 per producer warp and tile
  volatile uint32_t* issued_w = reinterpret_cast<volatile uint32_t*>(bars + 29);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 30);
  const uint32_t ready_word = smem_u32(bars + 31);   // number of weight stages whose operands have all landed

  // ---- one-time setup: constants to shared memory, barriers, tensor memory
  for (int i = threadIdx.x; i < prog.n_consts; i += blockDim.x) consts[i] = p.consts[i];
  // backward programs: `part` (unused there) holds this CTA's largest operand pattern per step (MlpParams::stash_absmax)
  if (BWD) for (int i = threadIdx.x; i < kMaxSteps; i += blockDim.x) reinterpret_cast<uint32_t*>(part)[i] = 0u;
  if (warp == 0) {
    tmem_alloc<512>(smem_u32(tmem_slot));
    tmem_relinquish();
  }
  if (threadIdx.x == 32) {
    for (int s = 0; s < kRing; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, kClusterSize);   // released by the MMA commits of every CTA in the cluster
    }
    for (int h = 0; h < 2; ++h) {
      mbar_init(bar_acc_full + 8 * h, 1);
      mbar_init(bar_dir_full + 8 * h, kProWarps * 32);
      mbar_init(bar_dir_empty + 8 * h, 1);
    }
    mbar_init(bar_war, 1);
    mbar_init(bar_view_full, 1);
    for (int w = 24; w < 32; ++w) bars[w] = 0ull;
    mbar_init(bar_emb_full, kProWarps * 32);
    mbar_init(bar_emb_empty, 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();   // the peer's barriers must be initialised before anything is multicast into them
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int n_stages = prog.n_stages, n_steps = prog.n_steps;
  const int n_esteps = VP ? n_steps - 1 : n_steps;   // steps whose epilogue the epilogue warps run

  if (warp < kEpiWarps) {
    // =============================================================== epilogue warps
    const int q = warp & 3, ch = warp >> 2;     // TMEM lane quarter ; which share of a column block
    constexpr int kCh = kEpiWarps / 4;          // warps sharing one lane quarter
    const int row = q * 32 + lane;
    const uint32_t tmem_lane = tmem + ((uint32_t)(q * 32) << 16);
    uint32_t gstep = 0;
    uint32_t vmax = 0;                          // largest hi-part bit patterns this thread produced (16x2)
    // compositing state (COMP): see mlp_program.h for the shared-memory map
    float* w_row = reinterpret_cast<float*>(smem + kSmemCompW);
    float* qprod = reinterpret_cast<float*>(smem + kSmemCompQ);   // [2][4]
    float* carry = qprod + 8;                                       // [2]
    float* qsum = reinterpret_cast<float*>(smem + kSmemCompS);     // [2][4][kCompChPad]
    float* racc = reinterpret_cast<float*>(smem + kSmemCompR);     // [kCompChPad]
    const int nch = 5 + p.C + p.K;
    if (COMP) {
      for (int c = threadIdx.x; c < kCompChPad; c += kEpiWarps * 32) racc[c] = 0.f;
      if (threadIdx.x == 0) carry[0] = 1.0f;
      named_bar_sync(2, kEpiWarps * 32);
    }
    for (int it = 0; it < n_iter; ++it) {
      const int64_t s_base = tile_base(it);
      const int64_t s = s_base + row;
      const int64_t s_end = cta_end();
      const bool valid = s < s_end;
      const int par = it & 1;
      float* qsum_q = qsum + (par * 4 + q) * kCompChPad;             // this quarter's partial sums of this tile
      float w_mine = 0.f;                                            // this row's compositing weight (COMP)
      float sig = 0.f;
      for (int st = 0; st < n_esteps; ++st, ++gstep) {
        const EpiDesc ed = prog.ep[st];
        const uint32_t parity = gstep & 1u;
        const float* bias = consts + ed.bias_off;
        const float* aux = consts + ed.aux_off;
        const bool to_a = BWD ? epi_writes_a(ed.kind) : ed.kind == EPI_RELU_TO_A;
        float c0 = 0.f, c1 = 0.f, c2 = 0.f;
        uint32_t smax = 0;   // backward programs: largest hi-part magnitudes of the A operand this step produces (16x2)
        // backward programs: this row's sign patterns and its row of the incoming gradient (tail rows read row S-1)
        uint16_t* mask_row = reinterpret_cast<uint16_t*>(smem + kSmemMask) + row;
        const float* gin = (BWD && ed.kind == EPI_LOADG_TO_A) ? p.grad_in + (valid ? s : p.S - 1) * (int64_t)ed.n : nullptr;
        float* stash_row = (BWD && p.stash != nullptr && ed.out_off1 != 0 && valid)
                               ? p.stash + ((int64_t)(ed.out_off1 - 1) * p.S + s) * (int64_t)ed.n : nullptr;
        if (BWD && st == 0 && p.grad_in != nullptr) {   // this row of the incoming gradient is needed D-1 steps from now: bring it into L2
          const float* g0 = p.grad_in + (valid ? s : p.S - 1) * (int64_t)prog.ep[0].n;
          for (int c = ch * 32; c < (int)prog.ep[0].n; c += kCh * 32) prefetch_l2(g0 + c);
        }
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
#ifdef PNR_TIMELINE
          const bool rec = p.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0 && it == 2;
          if (rec) p.dbg[4096 + (st * 2 + h) * 3 + 0] = clock64();
#endif
          // column blocks (in 16-column groups) of this half: part a = [pa0, pa1), part b = [pa1, pb1) (E1 only)
          const int pa0 = (h == 0 ? 0 : (int)ed.n0) >> 4;
          const int pa1 = (h == 0 ? (int)ed.n0 : (int)ed.n1a) >> 4;
          const int pb1 = (h == 0 ? (int)ed.n0 : (int)ed.n) >> 4;
          // this warp's near-equal contiguous share of each part
          const int a_lo = pa0 + (ch * (pa1 - pa0) + kCh - 1) / kCh, a_hi = pa0 + ((ch + 1) * (pa1 - pa0) + kCh - 1) / kCh;
          const int b_lo = pa1 + (ch * (pb1 - pa1) + kCh - 1) / kCh, b_hi = pa1 + ((ch + 1) * (pb1 - pa1) + kCh - 1) / kCh;
          const int na = a_hi - a_lo, nb = b_hi - b_lo;
          // (relaxed add: what is handed over lives in tensor memory - complete after tcgen05.wait::st and ordered
          //  by the tcgen05 fences on both sides; a release here is a MEMBAR.ALL.CTA per hand-off for nothing)
          auto signal = [&](uint32_t counter) {
            if (to_a) tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) red_add_smem(counter, 1u);
          };
          mbar_wait_backoff(bar_acc_full + 8 * h, parity);
          tc_fence_after();
#ifdef PNR_TIMELINE
          if (rec) p.dbg[4096 + (st * 2 + h) * 3 + 1] = clock64();
#endif
          // accumulator address of 16-column group g (odd tiles: columns XOR 128 when the program says so)
          const int flip = (prog.acc_flip && (it & 1)) ? 128 : 0;
          auto acc_of = [&](int g) { return tmem_lane + (uint32_t)(((int)ed.acc_col + g * 16) ^ flip); };
          // software pipeline: the load of the next group is in flight while a group is processed - also across
          // the part boundary, so signalling part a does not restart the load pipeline
          uint32_t ra[16], rb[16];
          if (na > 0) tmem_ld16(acc_of(a_lo), ra);
          else if (nb > 0) tmem_ld16(acc_of(b_lo), ra);
#pragma unroll
          for (int pi = 0; pi < 2; ++pi) {
            const int lo = pi == 0 ? a_lo : b_lo, hi = pi == 0 ? a_hi : b_hi;
            const int nxt = (pi == 0 && nb > 0) ? b_lo : -1;   // first group of the part that follows
            // E0 only (one part): the stores wait for the write-after-read barrier (once)
            bool war_pending = to_a && h == 0 && lo < hi;
            if (to_a) {
              // Two groups are converted before anything is stored: E0 reaches its write-after-read barrier
              // (this step's MMAs still read the columns it overwrites) with two groups of work already done.
#pragma unroll 1
              for (int g = lo; g < hi; g += 2) {
                uint32_t ha[8], la[8], hb[8], lb[8];
                const bool two = g + 1 < hi;
                const int after = (g + 2 < hi) ? g + 2 : nxt;
                tc_wait_ld();
                if (two) tmem_ld16(acc_of(g + 1), rb);
                if (BWD) epi_group_bwd<PASSES, FMT>(ra, g, ed, bias, mask_row, gin, stash_row, p.grad_scale, p.grad_unscale, smax, ha, la);
                else epi_group_act<PASSES, FMT>(ra, g, ed, bias, aux, sig, vmax, ha, la);
                if (two) {
                  tc_wait_ld();
                  if (after >= 0) tmem_ld16(acc_of(after), ra);
                  if (BWD) epi_group_bwd<PASSES, FMT>(rb, g + 1, ed, bias, mask_row, gin, stash_row, p.grad_scale, p.grad_unscale, smax, hb, lb);
                  else epi_group_act<PASSES, FMT>(rb, g + 1, ed, bias, aux, sig, vmax, hb, lb);
                } else if (after >= 0) {
                  tmem_ld16(acc_of(after), ra);
                }
                if (war_pending) {  // the columns we are about to overwrite must have been consumed by the MMAs
                  mbar_wait_backoff(bar_war, parity);
                  tc_fence_after();
                  war_pending = false;
                }
                epi_group_store<PASSES>(g, ed, tmem_lane, ha, la);
                if (two) epi_group_store<PASSES>(g + 1, ed, tmem_lane, hb, lb);
              }
            } else {
              // EPI_LOGITS: where this half's columns go (the second half may be a logit layer of its own)
              const bool own_half = !BWD && h == 1 && ed.n_valid1 > 0;
              float* out_row = p.raw + (valid ? s : 0) * p.CH + (own_half ? ed.out_off1 : ed.out_off);
              const int out_c0 = own_half ? (int)ed.n0 : 0, out_valid = own_half ? (int)ed.n_valid1 : (int)ed.n_valid;
              const int out_ch = (own_half ? (int)ed.out_off1 : (int)ed.out_off) + 1;   // composited channel of column out_c0
              const bool out_vec = BWD && (p.CH & 3) == 0 && (int)ed.n <= p.CH && (reinterpret_cast<uintptr_t>(p.raw) & 15) == 0 &&
                                   (ed.out_off & 3) == 0;
#pragma unroll 1
              for (int g = lo; g < hi; g += 2) {
                const bool two = g + 1 < hi;
                const int after = (g + 2 < hi) ? g + 2 : nxt;
                tc_wait_ld();
                if (two) tmem_ld16(acc_of(g + 1), rb);
                if (BWD) {
                  if (valid && ed.kind == EPI_ACT_OUT) epi_group_actout(ra, g, bias, out_row);
                  else if (valid) epi_group_gradout(ra, g, ed.n_valid, ed.n_valid1 != 0, out_vec, p.grad_unscale, out_row);
                } else if (ed.kind == EPI_VIEW_RGB) {
                  epi_group_rgb(ra, g, ed, bias, aux, c0, c1, c2);
                } else if (COMP) {
                  epi_group_logits_comp(ra, g, out_c0, out_valid, out_ch, bias, w_mine, lane, qsum_q);
                } else if (valid) {
                  epi_group_logits(ra, g, out_c0, out_valid, bias, out_row);
                }
                if (two) {
                  tc_wait_ld();
                  if (after >= 0) tmem_ld16(acc_of(after), ra);
                  if (BWD) {
                    if (valid && ed.kind == EPI_ACT_OUT) epi_group_actout(rb, g + 1, bias, out_row);
                    else if (valid) epi_group_gradout(rb, g + 1, ed.n_valid, ed.n_valid1 != 0, out_vec, p.grad_unscale, out_row);
                  } else if (ed.kind == EPI_VIEW_RGB) {
                    epi_group_rgb(rb, g + 1, ed, bias, aux, c0, c1, c2);
                  } else if (COMP) {
                    epi_group_logits_comp(rb, g + 1, out_c0, out_valid, out_ch, bias, w_mine, lane, qsum_q);
                  } else if (valid) {
                    epi_group_logits(rb, g + 1, out_c0, out_valid, bias, out_row);
                  }
                } else if (after >= 0) {
                  tmem_ld16(acc_of(after), ra);
                }
              }
            }
            if (pi == 0 && h == 1) signal(cnt_e1a);   // hand-off after part a of E1
          }
          if (!BWD && h == 1) {
            if (ed.sigma) {
              part[(ch * kTileM + row) * 4 + 3] = sig;
              sig = 0.f;
              if (VP) __threadfence_block();   // read by a producer warp, which synchronises through the MMA chain only
            }
            if (ed.kind == EPI_VIEW_RGB) {
              float* mine = part + (ch * kTileM + row) * 4;
              mine[0] = c0; mine[1] = c1; mine[2] = c2;
              named_bar_sync(1, kEpiWarps * 32);
              if (ch == 0 && (COMP || valid)) {
                const float* b3 = consts + prog.rgb_bias_off;
                float o0 = c0, o1 = c1, o2 = c2, o3 = mine[3];
#pragma unroll
                for (int oc = 1; oc < kEpiWarps / 4; ++oc) {   // fixed order: deterministic sums
                  const float* other = part + (oc * kTileM + row) * 4;
                  o0 += other[0]; o1 += other[1]; o2 += other[2]; o3 += other[3];
                }
                o0 += b3[0]; o1 += b3[1]; o2 += b3[2]; o3 += consts[prog.sigma_bias_off];
                if (COMP) {   // colours weighted and summed over this aligned group of 32 samples (w = 0 on dummy rows)
                  const float r0 = warp_sum32(w_mine * comp_sigmoid(o0));
                  const float r1 = warp_sum32(w_mine * comp_sigmoid(o1));
                  const float r2 = warp_sum32(w_mine * comp_sigmoid(o2));
                  if (lane == 0) { qsum_q[0] = r0; qsum_q[1] = r1; qsum_q[2] = r2; }
                } else {
                  float* dst = p.raw + s * p.CH;
                  if (p.CH == 4) {
                    *reinterpret_cast<float4*>(dst) = make_float4(o0, o1, o2, o3);
                  } else {
                    dst[0] = o0; dst[1] = o1; dst[2] = o2; dst[3] = o3;
                  }
                }
              }
            }
          }
          signal(h == 0 ? cnt_e0 : cnt_e1);
#ifdef PNR_TIMELINE
          if (rec) p.dbg[4096 + (st * 2 + h) * 3 + 2] = clock64();
#endif
        }
        if (BWD) {
          vmax = __vimax3_u16x2(vmax, smax, 0u);          // the range check sees every step
          if (p.stash_absmax != nullptr && to_a && ed.out_off1 != 0) {   // per-step maximum of what went to the stash
            uint32_t m = smax & 0xFFFFu;
            if ((smax >> 16) > m) m = smax >> 16;
            m = __reduce_max_sync(0xffffffffu, m);
            if (lane == 0) atomicMax(reinterpret_cast<uint32_t*>(part) + st, m);
          }
        }
        if (COMP && ed.sigma) {
          // ---- per-sample weights of this tile, right after the sigma-producing layer handed over (the MMAs go on
          // with the view / head steps meanwhile).  sigma = the two column shares' partial dot products + bias, in
          // the order the raw-writing kernel uses.  One warp per lane quarter = one aligned group of 32 samples.
          named_bar_sync(2, kEpiWarps * 32);            // both shares of every row's sigma are in shared memory
          if (ch == 0) {
            const int64_t sq = s_base + q * 32;         // first sample of this quarter; live or dummy as a whole
            const bool live = sq < s_end;
            float alpha = 0.f, zi = 0.f, t = 1.0f;
            if (live) {
              const int64_t r = sq / p.N;
              const int i = (int)(sq - r * p.N) + lane;
              zi = p.z[sq + lane];
              const bool has_next = i + 1 < p.N;
              const float z_next = has_next ? p.z[sq + lane + 1] : 0.f;
              const float* rr = p.rays + r * 6;
              const float dist = comp_dist(zi, z_next, has_next, comp_dnorm(rr[3], rr[4], rr[5]));
              float sraw = part[row * 4 + 3];
              sraw += part[(kTileM + row) * 4 + 3];
              sraw += consts[prog.sigma_bias_off];
              const bool masked = p.mask_outside && p.sample_box != nullptr && p.sample_box[sq + lane] < 0;
              alpha = comp_alpha(sraw, dist, masked);
              t = 1.0f - alpha + 1e-10f;
            }
            float total;
            const float excl = comp_scan32(t, lane, &total);
            if (lane == 0) qprod[par * 4 + q] = total;
            named_bar_sync(3, 4 * 32);                  // the four quarter products
            // transmittance at this quarter's first sample: carried in from the previous tile, restarted where a
            // ray starts, multiplied through the earlier quarters in order (the order is fixed by the sample index)
            float prefix = carry[par];
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
              if (qq <= q) {
                if ((s_base + 32 * qq) % p.N == 0) prefix = 1.0f;
                if (qq < q) prefix *= qprod[par * 4 + qq];
              }
            }
            const float wi = live ? alpha * (prefix * excl) : 0.f;
            w_row[row] = wi;
            if (live) p.weights[sq + lane] = wi;
            const float sd = warp_sum32(wi * zi), sa = warp_sum32(wi);
            if (lane == 0) {
              qsum_q[3] = sd;
              qsum_q[4] = sa;
              if (q == 3) carry[par ^ 1] = prefix * total;   // transmittance behind the tile's last sample
            }
          }
          named_bar_sync(2, kEpiWarps * 32);            // w_row is visible to both column shares
          w_mine = w_row[row];
        }
      }
      if (COMP) {
        // ---- end of tile: fold this tile's per-quarter sums into the running sums of the open ray, in ray order;
        // a ray is written out when the next one starts (or the CTA's range ends).  One warp, lanes stride channels.
        named_bar_sync(2, kEpiWarps * 32);              // every quarter's partial sums are in shared memory
        if (warp == 0) {
          auto flush = [&](int64_t r) {
            const float depth = racc[3], acc = racc[4];
            __syncwarp();
            for (int c = lane; c < nch; c += 32) {
              const float v = racc[c];
              if (c < 3) { if (p.rgb_map) p.rgb_map[r * 3 + c] = v + (p.white_bkgd ? (1.0f - acc) : 0.f); }
              else if (c == 3) {
                if (p.depth_map) p.depth_map[r] = v;
                if (p.disp_map) p.disp_map[r] = comp_disp(depth, acc);
              }
              else if (c == 4) { if (p.acc_map) p.acc_map[r] = v; }
              else if (c < 5 + p.C) { if (p.sem_map) p.sem_map[r * p.C + (c - 5)] = v; }
              else { if (p.inst_map) p.inst_map[r * p.K + (c - 5 - p.C)] = v; }
              racc[c] = 0.f;
            }
            __syncwarp();
          };
#pragma unroll 1
          for (int qq = 0; qq < 4; ++qq) {
            const int64_t sq = s_base + 32 * qq;
            if (sq >= s_end) break;
            if (sq > cta_first() && sq % p.N == 0) flush(sq / p.N - 1);    // the previous ray ended right before sq
            const float* src = qsum + (par * 4 + qq) * kCompChPad;
            for (int c = lane; c < nch; c += 32) racc[c] += src[c];
            __syncwarp();
          }
          if (s_base < s_end && s_base + kTileM >= s_end) flush(s_end / p.N - 1);   // the CTA's last ray
        }
      }
    }
    if (p.status != nullptr && ((vmax & 0xFFFFu) >= inf_bits16<FMT>() || (vmax >> 16) >= inf_bits16<FMT>()))
      atomicOr(p.status, 1u);
  } else if (warp < kEpiWarps + kProWarps) {
    // =============================================================== embedding producer warps
    const int row = (warp - kEpiWarps) * 32 + lane;
    uint8_t* emb_hi = smem + kSmemEmb;
    uint8_t* emb_lo = emb_hi + kEmbPartBytes;
    const int Lx = prog.Lx, Ld = prog.Ld;
    // VP: the view step's epilogue of tile t for this thread's row - what the epilogue warps do for EPI_VIEW_RGB, in
    // the same summation order (the two column shares' partial sums, share 0 first), so the results are bit-identical
    auto view_epilogue = [&](int t) {
      const EpiDesc ed = prog.ep[prog.view_step];
      const float* bias = consts + ed.bias_off;
      const float* aux = consts + ed.aux_off;
      const int64_t sv = tile_base(t) + row;
      const int q = warp & 3;                                   // this warp's TMEM lane quarter = its 32 rows
      const uint32_t tmem_lane = tmem + ((uint32_t)(q * 32) << 16);
      const int flip = (prog.acc_flip && (t & 1)) ? 128 : 0;
      const int ng = (int)ed.n >> 4, half = (ng + 1) / 2;       // share 0: groups [0, half), share 1: [half, ng)
      mbar_wait_backoff(bar_view_full, (uint32_t)(t & 1));
      tc_fence_after();
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f;
      uint32_t ra[16], rb[16];
      tmem_ld16(tmem_lane + (uint32_t)(((int)ed.acc_col) ^ flip), ra);
#pragma unroll 1
      for (int g = 0; g < ng; g += 2) {
        const bool two = g + 1 < ng;
        tc_wait_ld();
        if (two) tmem_ld16(tmem_lane + (uint32_t)(((int)ed.acc_col + (g + 1) * 16) ^ flip), rb);
        if (g < half) epi_group_rgb(ra, g, ed, bias, aux, a0, a1, a2);
        else epi_group_rgb(ra, g, ed, bias, aux, b0, b1, b2);
        if (two) {
          tc_wait_ld();
          if (g + 2 < ng) tmem_ld16(tmem_lane + (uint32_t)(((int)ed.acc_col + (g + 2) * 16) ^ flip), ra);
          if (g + 1 < half) epi_group_rgb(rb, g + 1, ed, bias, aux, a0, a1, a2);
          else epi_group_rgb(rb, g + 1, ed, bias, aux, b0, b1, b2);
        }
      }
      if (sv < p.S) {
        const float* b3 = consts + prog.rgb_bias_off;
        const float o0 = a0 + b0 + b3[0], o1 = a1 + b1 + b3[1], o2 = a2 + b2 + b3[2];
        const float o3 = part[row * 4 + 3] + part[(kTileM + row) * 4 + 3] + consts[prog.sigma_bias_off];
        float* dst = p.raw + sv * p.CH;
        if (p.CH == 4) {
          *reinterpret_cast<float4*>(dst) = make_float4(o0, o1, o2, o3);
        } else {
          dst[0] = o0; dst[1] = o1; dst[2] = o2; dst[3] = o3;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) red_add_smem(cnt_v, 1u);
    };
    for (int it = 0; it < n_iter; ++it) {
      int64_t s = tile_base(it) + row;
      if (s >= p.S) s = p.S - 1;  // clamp: tail rows compute on a valid sample, results are discarded
      float x[3], d[3];
      if (p.pts != nullptr) {
#pragma unroll
        for (int c = 0; c < 3; ++c) { x[c] = p.pts[s * 3 + c]; d[c] = BWD ? 0.f : p.viewdirs[s * 3 + c]; }
      } else {
        const int64_t ray = s / p.N;
        const float zi = p.z[s];
        const float* rr = p.rays + ray * 6;
        float dn2 = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float dc = rr[3 + c];
          x[c] = __fadd_rn(rr[c], __fmul_rn(dc, zi));  // pts = o + d*z, separately rounded like the oracle
          dn2 = (c == 0) ? __fmul_rn(dc, dc) : __fadd_rn(dn2, __fmul_rn(dc, dc));
          d[c] = dc;
        }
        const float nrm = sqrtf(dn2);
#pragma unroll
        for (int c = 0; c < 3; ++c) d[c] = __fdiv_rn(d[c], nrm);
      }
      mbar_wait_backoff(bar_emb_empty, (uint32_t)((it & 1) ^ 1));
      encode_row<PASSES, FMT, 10, 64>(x, Lx, emb_hi, emb_lo, row);
      fence_proxy_async_smem();
      mbar_arrive(bar_emb_full);
      if (!BWD) {   // (backward programs stop at the trunk: the view-direction region holds their sign patterns)
        const int b = it & 1;
        uint8_t* dir_hi = smem + kSmemDir + b * 2 * kDirPartBytes;
        mbar_wait_backoff(bar_dir_empty + 8 * b, (uint32_t)(((it >> 1) & 1) ^ 1));
        encode_row<PASSES, FMT, 4, 32>(d, Ld, dir_hi, dir_hi + kDirPartBytes, row);
        fence_proxy_async_smem();
        mbar_arrive(bar_dir_full + 8 * b);
      }
      if (VP && it >= 1) view_epilogue(it - 1);   // (this tile's embeddings first: the MMAs of its layer 0 follow the view MMAs)
    }
    if (VP && n_iter >= 1) view_epilogue(n_iter - 1);
  } else if (warp == kEpiWarps + kProWarps) {
    // =============================================================== TMA producer (one elected thread)
    if (elect_one()) {
      uint32_t gs = 0;  // global stage counter
      for (int it = 0; it < n_iter; ++it) {
        for (int si = 0; si < n_stages; ++si, ++gs) {
          const uint32_t slot = gs % kRing, ph = (gs / kRing) & 1;
          const uint32_t gofs = prog.st[si].gofs, bytes = prog.st[si].bytes;
          mbar_wait_backoff(bar_empty + 8 * slot, ph ^ 1);
#ifdef PNR_TIMELINE
          if (p.dbg != nullptr && blockIdx.x == 0 && it == 2) p.dbg[6144 + si] = clock64();
#endif
          // our barrier expects the whole stage; we fetch our 1/kClusterSize of it and multicast that part into
          // every CTA of the cluster (same shared-memory offset, same barrier offset in each of them)
          mbar_arrive_expect_tx(bar_full + 8 * slot, bytes);
          const uint32_t part_bytes = bytes / kClusterSize;
          bulk_g2s_multicast(smem_u32(smem + kSmemRing + slot * kStageBytes) + cta_rank * part_bytes,
                             p.wpacked + gofs + cta_rank * part_bytes, part_bytes, bar_full + 8 * slot,
                             (uint16_t)((1u << kClusterSize) - 1u));
        }
      }
    }
  } else if (warp == kEpiWarps + kProWarps + 2) {
    // =============================================================== scout (one elected thread)
    // Does every wait the MMA issue depends on, in stage order, and publishes "stages ready" through ONE
    // shared-memory word, so the issuers never execute an mbarrier wait (each costs ~100 cycles even when
    // already complete) and poll a single word.  Operands that arrive through the async proxy (weight stage
    // landed, embeddings written) are checked first - they are ready long before they are needed - then the
    // scout spins on the epilogue hand-off counters: epilogue warp -> counter -> scout -> ready word -> issuer
    // is two shared-memory round trips (~100 cycles; the mbarrier relay it replaces took ~560, timeline v10).
    if (elect_one()) {
      uint32_t gs = 0;
      uint32_t have_e0 = 0, have_e1a = 0, have_e1 = 0, have_v = 0;   // last values read from the hand-off counters
      auto spin = [&](uint32_t addr, uint32_t target, uint32_t& have, const char* what) {
        if ((int32_t)(have - target) >= 0) return;
        const long long t0 = clock64();
        while ((int32_t)((have = ld_volatile_smem(addr)) - target) < 0) {
          if ((clock64() - t0) > PNR_WATCHDOG_CYCLES) {
            printf("pnr: scout watchdog (%s): block %d stage %u\n", what, (int)blockIdx.x, gs);
            __trap();
          }
        }
      };
      for (int it = 0; it < n_iter; ++it) {
        const int b = it & 1;
        const uint32_t step_base = (uint32_t)(it * n_esteps) - 1u;   // needs are stored + 1
        uint32_t needs = prog.is[0].needs;
#pragma unroll 1
        for (int si = 0; si < n_stages; ++si, ++gs) {
          const uint32_t flags = prog.st[si].flags;
          const uint32_t needs_next = prog.is[si + 1 < n_stages ? si + 1 : 0].needs;   // (in flight during the waits)
          if (flags & F_WAIT_EMB) mbar_wait(bar_emb_full, (uint32_t)(it & 1));
          if (flags & F_WAIT_DIR) mbar_wait(bar_dir_full + 8 * b, (uint32_t)((it >> 1) & 1));
          const uint32_t slot = gs % kRing, ph = (gs / kRing) & 1;
          mbar_wait(bar_full + 8 * slot, ph);
          // kEpiWarps arrivals per completed epilogue part
          spin(cnt_e0, (step_base + (needs & 0xFFu)) * kEpiWarps, have_e0, "E0");
          spin(cnt_e1a, (step_base + ((needs >> 8) & 0xFFu)) * kEpiWarps, have_e1a, "E1a");
          spin(cnt_e1, (step_base + ((needs >> 16) & 0xFFu)) * kEpiWarps, have_e1, "E1");
          if (VP && (needs >> 24)) spin(cnt_v, (uint32_t)it * kProWarps, have_v, "view");   // previous tile's view epilogue
          tc_fence_before();
          st_release_smem(ready_word, gs + 1);
          needs = needs_next;
        }
      }
    }
  } else {
    // =============================================================== MMA issuer
    // The whole warp walks the stage list in lock step, so the loop state (stage words from the parameter
    // bank, ring slot, descriptors) stays on the uniform datapath; only the tcgen05 instructions sit in an
    // elect.sync branch.  The issue table (IssueDesc) holds every per-stage word ready to use: measured on
    // a stand-alone replica (tools/probe_issue3.cu) this loop needs ~570 cycles per 12-MMA stage next to
    // ALU-saturating warps, the field-by-field version it replaces ~1000 (the tensor work is 768).
    //
    // Two such warps (on different SM sub-partitions) take alternate stages.  The tensor pipe's queue is only a
    // few MMAs deep, so the ~400-600 cycles one warp spends between two bursts (commits, next stage's words,
    // loop) drain it; with two warps that work overlaps the other warp's burst.  The owner of stage g bursts
    // only after the owner of g-1 has issued (shared counter `issued`): MMAs of one accumulator keep their
    // order, and because the pipe retires a CTA's MMAs in issue order, the commit that follows a half's last
    // stage also covers the stages the other warp issued for it (hardware assumption).
    const uint32_t me = (warp == kEpiWarps + kProWarps + 1) ? 0u : 1u;
    uint32_t gs = 0, ready = 0, slot = 0, issued = 0;
    // (address field only: in a cluster the shared-window address of CTA rank > 0 carries the rank above it)
    const uint32_t ring16 = (smem_u32(smem + kSmemRing) >> 4) & 0x3FFFu;
    const uint32_t emb_hi = smem_u32(smem + kSmemEmb);
    constexpr uint32_t kFullK = PASSES == 3 ? 4u : 8u;   // K16 steps of a full stage
    for (int it = 0; it < n_iter; ++it) {
      const int b = it & 1;
      const uint32_t acc_flip = (prog.acc_flip && (it & 1)) ? 128u : 0u;
      const uint32_t dir_hi = smem_u32(smem + kSmemDir + b * 2 * kDirPartBytes);
#pragma unroll 1
      for (int si = 0; si < n_stages; ++si, ++gs, slot = (slot + 1 == (uint32_t)kRing) ? 0u : slot + 1) {
        if ((gs & 1u) != me) continue;
        const uint32_t idesc = prog.is[si].idesc, b_lo_base = prog.is[si].b_lo_base;
        const uint32_t b_inc = prog.is[si].b_inc, lo_off16 = prog.is[si].lo_off16;
        const uint32_t acc_col = prog.is[si].acc_col, a_off = prog.is[si].a_off;
        const uint32_t a_lo_off = prog.is[si].a_lo_off, fk = prog.is[si].flags_k;
        const uint32_t flags = fk & 0xFFFFu, ksteps = (fk >> 16) & 0xFFu, a_kind = fk >> 24;
#ifdef PNR_TIMELINE
        const bool rec = p.dbg != nullptr && blockIdx.x == 0 && it == 2 && lane == 0;
        if (rec) p.dbg[si * 5 + 0] = clock64();
#endif
        if (ready <= gs) {   // the scout may be several stages ahead: poll only when our copy is stale
          ready = ld_acquire_smem(ready_word);
          if (ready <= gs) {
            const long long t0 = clock64();
            while ((ready = ld_acquire_smem(ready_word)) <= gs) {
              if ((clock64() - t0) > PNR_WATCHDOG_CYCLES) {
                if (lane == 0) printf("pnr: issuer watchdog: block %d stage %u\n", (int)blockIdx.x, gs);
                __trap();
              }
            }
          }
        }
#ifdef PNR_TIMELINE
        if (rec) p.dbg[si * 5 + 1] = clock64();
#endif
        // low words of the weight-tile descriptors (K16 step 0; step ks adds ks * b_inc to the address field);
        // the high word is the same for every operand: SBO = 128 B, descriptor version 1
        const uint32_t b_hi0 = b_lo_base | (ring16 + slot * (uint32_t)(kStageBytes >> 4));
        const uint32_t b_lo0 = b_hi0 + lo_off16;
        const uint32_t d_tmem = tmem + (acc_col ^ acc_flip);
        const uint32_t acc0 = (flags & F_FIRST) ? 0u : 1u;
        const uint32_t a_hi = tmem + a_off, a_lo = tmem + a_lo_off;
        const bool fast = a_kind == A_TMEM && ksteps == kFullK;
        if (issued < gs) {   // the other warp must have issued stage gs-1
          long long t0 = clock64();
          while ((issued = *issued_w) < gs) {
            if ((clock64() - t0) > PNR_WATCHDOG_CYCLES) {
              if (lane == 0) printf("pnr: issuer hand-off watchdog: block %d stage %u\n", (int)blockIdx.x, gs);
              __trap();
            }
          }
        }
        if (elect_one()) {
          tc_fence_after();   // order our MMAs after the epilogue's tcgen05.ld / tcgen05.st (seen by the scout)
          if (fast) {
#pragma unroll
            for (uint32_t ks = 0; ks < kFullK; ++ks) {
              mma_ts_lo(d_tmem, a_hi + ks * 8, b_hi0 + ks * b_inc, idesc, ks == 0 ? acc0 : 1u);
              if (PASSES == 3) {
                mma_ts_lo(d_tmem, a_lo + ks * 8, b_hi0 + ks * b_inc, idesc, 1u);
                mma_ts_lo(d_tmem, a_hi + ks * 8, b_lo0 + ks * b_inc, idesc, 1u);
              }
            }
          } else if (a_kind == A_TMEM) {
#pragma unroll 1
            for (uint32_t ks = 0; ks < ksteps; ++ks) {
              mma_ts_lo(d_tmem, a_hi + ks * 8, b_hi0 + ks * b_inc, idesc, ks == 0 ? acc0 : 1u);
              if (PASSES == 3) {
                mma_ts_lo(d_tmem, a_lo + ks * 8, b_hi0 + ks * b_inc, idesc, 1u);
                mma_ts_lo(d_tmem, a_hi + ks * 8, b_lo0 + ks * b_inc, idesc, 1u);
              }
            }
          } else {
            const uint32_t a_base = (a_kind == A_EMB) ? emb_hi : dir_hi;
            const uint32_t a_lo_delta = (a_kind == A_EMB) ? (uint32_t)kEmbPartBytes : (uint32_t)kDirPartBytes;
            const uint64_t adesc0 = make_smem_desc_noswz(a_base, kTileM * 16, 128);
            const uint64_t adesc0_lo = make_smem_desc_noswz(a_base + a_lo_delta, kTileM * 16, 128);
            constexpr uint32_t a_inc = (2u * kTileM * 16u) >> 4;
            const uint64_t bdesc0 = ((uint64_t)kDescHiWord << 32) | b_hi0;
            const uint64_t bdesc0_lo = ((uint64_t)kDescHiWord << 32) | b_lo0;
#pragma unroll 1
            for (uint32_t ks = 0; ks < ksteps; ++ks) {
              mma_ss(d_tmem, adesc0 + (uint64_t)(ks * a_inc), bdesc0 + (uint64_t)(ks * b_inc), idesc, ks == 0 ? acc0 : 1u);
              if (PASSES == 3) {
                mma_ss(d_tmem, adesc0_lo + (uint64_t)(ks * a_inc), bdesc0 + (uint64_t)(ks * b_inc), idesc, 1u);
                mma_ss(d_tmem, adesc0 + (uint64_t)(ks * a_inc), bdesc0_lo + (uint64_t)(ks * b_inc), idesc, 1u);
              }
            }
          }
          *issued_w = gs + 1;
#ifdef PNR_TIMELINE
          if (rec) p.dbg[si * 5 + 2] = clock64();
#endif
          tc_commit_multicast(bar_empty + 8 * slot, (uint16_t)((1u << kClusterSize) - 1u));   // slot free in all CTAs
          if (flags & (F_RELEASE_EMB | F_RELEASE_DIR | F_COMMIT_WAR | F_COMMIT_ACC0 | F_COMMIT_ACC1 | F_COMMIT_VIEW)) {
            if (flags & F_RELEASE_EMB) tc_commit(bar_emb_empty);
            if (flags & F_RELEASE_DIR) tc_commit(bar_dir_empty + 8 * b);
            if (flags & F_COMMIT_WAR) tc_commit(bar_war);
            if (flags & F_COMMIT_ACC0) tc_commit(bar_acc_full);
            if (flags & F_COMMIT_ACC1) tc_commit(bar_acc_full + 8);
            if (VP && (flags & F_COMMIT_VIEW)) tc_commit(bar_view_full);
          }
#ifdef PNR_TIMELINE
          if (rec) { p.dbg[si * 5 + 3] = clock64(); p.dbg[si * 5 + 4] = p.dbg[si * 5 + 3]; }
#endif
        }
        __syncwarp();
        issued = gs + 1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (BWD && p.stash_absmax != nullptr && (int)threadIdx.x < prog.n_steps) {   // this CTA's maxima -> the launch's
    const EpiDesc& e = prog.ep[threadIdx.x];
    const uint32_t m = reinterpret_cast<const uint32_t*>(part)[threadIdx.x];
    if (epi_writes_a(e.kind) && e.out_off1 != 0 && m != 0u) atomicMax(p.stash_absmax + (e.out_off1 - 1), m);
  }
  cluster_sync();   // no CTA may exit while its peer can still multicast into it or arrive on its barriers
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// Per-device launch state: the > 48 KB dynamic shared-memory opt-in is a per-device function attribute, and so
// is the SM count the persistent grid is sized by.
struct DeviceState {
  bool attr_done[2][2][4] = {};   // [x3][bf16][plain / compositing epilogue / backward program / view on producers]
};
static DeviceState g_dev[kMaxDevices];
static std::mutex g_dev_mutex;

template <int PASSES, int FMT, bool COMP, bool BWD = false, bool VP = false>
static int launch_one(const MlpLaunch& L, int dev, int grid, cudaStream_t stream) {
  constexpr int kSmem = COMP ? kSmemTotalComp : kSmemTotal;
  {
    std::lock_guard<std::mutex> lock(g_dev_mutex);
    bool& done = g_dev[dev].attr_done[PASSES == 3][FMT == kFmtBF16][VP ? 3 : (BWD ? 2 : (COMP ? 1 : 0))];
    if (!done) {
      PNR_CUDA(cudaFuncSetAttribute(mlp_fused_kernel<PASSES, FMT, COMP, BWD, VP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    kSmem));
      done = true;
    }
  }
  mlp_fused_kernel<PASSES, FMT, COMP, BWD, VP><<<grid, kMlpThreads, kSmem, stream>>>(L);
  PNR_LAUNCH_CHECK("mlp_fused_kernel");
  return PNR_OK;
}

// Launches on the CURRENT device (the caller has made the context's device current).  mode 1: the
// compositing-epilogue variant; the launch's rays_per_cta is set here (whole rays per CTA, a multiple of the
// 128 / gcd(N, 128) rays that make whole tiles, so that every CTA's range starts on a 32-sample boundary).
// mode 2: L.prog is a backward program (x3 precisions only).
int launch_mlp(MlpLaunch& L, int passes, int fmt, int mode, cudaStream_t stream) {
  const bool composite = mode == kMlpComposite;
  int dev = 0;
  PNR_CUDA(cudaGetDevice(&dev));
  PNR_CHECK_ARG(dev >= 0 && dev < kMaxDevices, "launch_mlp: device ordinal %d >= %d", dev, kMaxDevices);
  const int sms = num_sms(dev);
  int grid = L.p.num_tiles < sms ? L.p.num_tiles : sms;
  if (grid <= 0) return PNR_OK;
  grid = (grid + kClusterSize - 1) / kClusterSize * kClusterSize;   // whole clusters (spare CTAs run dummy tiles)
  if (grid > sms) grid = sms / kClusterSize * kClusterSize;
  if (composite) {
    const int64_t R = L.p.S / L.p.N;
    L.p.rays_per_cta = (R + grid - 1) / grid;
  }
  if (mode == kMlpBackward) {
    if (passes != 3) return set_error(PNR_ERR_UNSUPPORTED, "backward programs run in the x3 precisions only");
    return fmt == kFmtF16 ? launch_one<3, kFmtF16, false, true>(L, dev, grid, stream)
                          : launch_one<3, kFmtBF16, false, true>(L, dev, grid, stream);
  }
  if (mode == kMlpForwardVP) {
    if (L.prog.view_step < 0) return set_error(PNR_ERR_STATE, "launch_mlp: not a view-on-producers program");
#define PNR_LAUNCH_VP(P, F) launch_one<P, F, false, false, true>(L, dev, grid, stream)
    if (fmt == kFmtF16) return passes == 3 ? PNR_LAUNCH_VP(3, kFmtF16) : PNR_LAUNCH_VP(1, kFmtF16);
    return passes == 3 ? PNR_LAUNCH_VP(3, kFmtBF16) : PNR_LAUNCH_VP(1, kFmtBF16);
#undef PNR_LAUNCH_VP
  }
#define PNR_LAUNCH(P, F) (composite ? launch_one<P, F, true>(L, dev, grid, stream) : launch_one<P, F, false>(L, dev, grid, stream))
  if (fmt == kFmtF16) return passes == 3 ? PNR_LAUNCH(3, kFmtF16) : PNR_LAUNCH(1, kFmtF16);
  return passes == 3 ? PNR_LAUNCH(3, kFmtBF16) : PNR_LAUNCH(1, kFmtBF16);
#undef PNR_LAUNCH
}

}  // namespace pnr
