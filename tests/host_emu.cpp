// host_emu.cpp — host build of csrc/ray_math.h (the exact per-ray routines the CUDA kernels call),
// so CPU tests can prove the fp32 operation order equals the oracle's bit for bit without a GPU.
// Built by tests/test_cpu_ray_math.py with: g++ -O2 -ffp-contract=off -shared -fPIC.
#include <cstdint>
#include "../panopticnerf_b200/csrc/ray_math.h"

extern "C" {

void emu_intersect(const float* rays, int64_t R, const float* bc, const float* bh, const float* br, int B,
                   int M, uint8_t* hit, int32_t* box_id, float* t_in, float* t_out) {
  for (int64_t r = 0; r < R; ++r) {
    const float* q = rays + r * 6;
    PnrHitList L;
    pnr_hits_init(&L);
    for (int b = 0; b < B; ++b) {
      float tmin, tmax;
      if (pnr_slab(q[0], q[1], q[2], q[3], q[4], q[5], bc + b * 3, bh + b * 3, br + b * 9, &tmin, &tmax))
        pnr_hits_insert(&L, M, tmin, tmax, b);
    }
    hit[r] = L.n > 0;
    for (int m = 0; m < M; ++m) {
      const bool v = m < L.n;
      box_id[r * M + m] = v ? L.id[m] : -1;
      t_in[r * M + m] = v ? pnr_max_nan(L.key[m], 0.f) : 0.f;
      t_out[r * M + m] = v ? L.tout[m] : 0.f;
    }
  }
}

void emu_stratified(const float* near, const float* far, const float* t_vals, const float* u, int64_t R, int N,
                    float perturb, const int32_t* box_id, const float* t_in, const float* t_out, int M, float* z,
                    int32_t* sb) {
  for (int64_t r = 0; r < R; ++r)
    for (int i = 0; i < N; ++i) {
      const float zi = perturb > 0.f ? pnr_strat_z_jitter(near[r], far[r], t_vals, i, N, u[r * N + i])
                                     : pnr_strat_z(near[r], far[r], t_vals[i]);
      z[r * N + i] = zi;
      if (sb) sb[r * N + i] = pnr_tag(zi, box_id + r * M, t_in + r * M, t_out + r * M, M);
    }
}

void emu_intervals(const float* near, const float* far, const float* t_vals, const float* u, int64_t R, int N,
                   float perturb, const int32_t* box_id, const float* t_in, const float* t_out, int M, float* z,
                   int32_t* sb) {
  for (int64_t r = 0; r < R; ++r) {
    PnrIntervalPlan P;
    pnr_interval_plan(near[r], far[r], box_id + r * M, t_in + r * M, t_out + r * M, M, N, &P);
    float* zr = z + r * N;
    if (P.kept == 0) {
      for (int i = 0; i < N; ++i)
        zr[i] = perturb > 0.f ? pnr_strat_z_jitter(near[r], far[r], t_vals, i, N, u[r * N + i])
                              : pnr_strat_z(near[r], far[r], t_vals[i]);
    } else {
      for (int k = 0; k < N; ++k) zr[k] = pnr_interval_z(&P, k, perturb > 0.f ? u[r * N + k] : 0.5f);
      for (int i = 1; i < N; ++i) {   // insertion sort: the kernel's bitonic network yields the same sorted array
        const float v = zr[i];
        int j = i - 1;
        while (j >= 0 && zr[j] > v) { zr[j + 1] = zr[j]; --j; }
        zr[j + 1] = v;
      }
    }
    for (int i = 0; i < N; ++i)
      sb[r * N + i] = P.kept == 0 ? -1 : pnr_tag(zr[i], box_id + r * M, t_in + r * M, t_out + r * M, M);
  }
}

void emu_sample_pdf(const float* z, const float* w, int64_t R, int N, int Ni, const float* u, float* z_f,
                    int64_t* idx) {
  float cdf[256];
  for (int64_t r = 0; r < R; ++r) {
    pnr_pdf_cdf(w + r * N, N, cdf);
    for (int j = 0; j < Ni; ++j) {
      const float uj = u[r * Ni + j];
      int id;
      z_f[r * Ni + j] = pnr_pdf_sample(z + r * N, cdf, N - 1, uj, &id);
      idx[r * Ni + j] = id;
    }
  }
}

void emu_fisheye(int H, int W, const float* k, float* dirs) {
  for (int v = 0; v < H; ++v)
    for (int u = 0; u < W; ++u) {
      float* o = dirs + ((size_t)v * W + u) * 3;
      pnr_fisheye_dir((float)u, (float)v, k[0], k[1], k[2], k[3], k[4], k[5], k[6], o, o + 1, o + 2);
    }
}
}
