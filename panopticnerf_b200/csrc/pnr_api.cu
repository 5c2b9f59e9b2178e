// pnr_api.cu — context management, weight packing / MLP program construction and the
// pnr_mlp_forward entry point of the C ABI (include/pnr.h).
#include <cstdlib>
#include <cstring>
#include <cuda_fp16.h>
#include <string>
#include <vector>
#include "common.cuh"
#include "mlp_program.h"
#include "tc05.cuh"

namespace pnr {

// ---------------------------------------------------------------- error / accounting plumbing
static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
void count_launch(int n) { g_launches += n; }

int launch_mlp(MlpLaunch& L, int passes, int fmt, int mode, cudaStream_t stream);  // mlp_tc05.cu

}  // namespace pnr

using namespace pnr;

struct pnr_ctx {
  pnr_config cfg;
  int passes = 3;
  int fmt = 0;   // 0 = fp16, 1 = bf16 (instruction-descriptor encoding)
  bool loaded = false;
  MlpLaunch launch;             // launch.prog = this context's program; launch.p is filled per call.  The whole
                                // struct travels as the kernel's __grid_constant__ parameter: nothing is shared
                                // between contexts, streams, devices or CUDA-graph replays.
  MlpLaunch launch_vp;          // the same stages with the view epilogue on the producer warps (networks without heads;
  bool has_vp = false;          //   pnr_mlp_forward uses it, pnr_mlp_composite keeps the standard program)
  uint8_t* d_wpacked = nullptr;
  float* d_consts = nullptr;
  uint32_t* d_status = nullptr; // sticky range-check word of the fused MLP (bit 0: activation out of operand range)
  size_t wpacked_bytes = 0;
  // backward program of the trunk (pnr_mlp_backward_trunk): built from a host copy of the trunk weights on first use
  std::vector<std::vector<float>> host_trunk;   // weight, bias per trunk layer, as given to pnr_load_weights
  std::vector<int64_t> host_trunk_shapes;
  long long* dbg_timeline = nullptr;   // development aid (pnr_debug_timeline)
  struct Aux {                         // a second program over the same weights, with its own packed stream
    bool ready = false;
    MlpLaunch launch;
    uint8_t* d_wpacked = nullptr;
    float* d_consts = nullptr;
    size_t n_w = 0, n_c = 0;           // packed 16-bit elements / constants
  };
  Aux bwd;                             // forward trunk + the layers in reverse (pnr_mlp_backward_trunk)
  Aux trunk_fwd;                       // forward trunk only, output = its activations (pnr_mlp_trunk_forward)
  // device-side weight updates (pnr_update_weights): V = all input tensors concatenated + the derived (folded) values,
  // and per program the plan "packed element p = part wpart[p] of V[widx[p]]" (Builder index mode)
  struct Plan {
    bool ready = false;
    int32_t* d_widx = nullptr; uint8_t* d_wpart = nullptr; size_t n_w = 0;
    int32_t* d_cidx = nullptr; size_t n_c = 0;
  };
  Plan plan_main, plan_bwd, plan_tf;
  float* d_V = nullptr;
  std::vector<int64_t> v_off, all_shapes;   // position of every input tensor in V; the shapes given to pnr_load_weights
  int64_t v_total = 0, v_derived = 0;
  bool device_weights = false;              // the weights in V are newer than the host copies (aux programs re-pack from V)
};

// ---------------------------------------------------------------- host-side 16-bit split (RNE, = cvt.rn.*.f32)
static inline uint16_t f2h(float x) {
  const __half h = __float2half_rn(x);
  uint16_t u;
  memcpy(&u, &h, 2);
  return u;
}
static inline float h2f(uint16_t u) {
  __half h;
  memcpy(&h, &u, 2);
  return __half2float(h);
}
static inline uint16_t f2bf(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  return (uint16_t)((u + 0x7FFFu + ((u >> 16) & 1u)) >> 16);
}
static inline float bf2f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

namespace {

struct Mat { const float* w; int out, in; };  // row-major [out, in]

struct Seg {
  uint8_t kind;        // A_TMEM / A_EMB / A_DIR
  Mat m;
  int col0, kvalid;    // columns of m feeding this segment
  int kpad;            // K rounded up (multiple of 16)
  int a_hi, a_lo;      // TMEM packed columns (A_TMEM)
  bool release;        // A_EMB/A_DIR: last reader in the tile
};

struct Builder {
  MlpProgram prog;
  std::vector<uint16_t> wbuf;   // packed weight stream (bf16 elements)
  std::vector<float> consts;
  int passes, fmt;
  // E1 in two blocks signalled separately (mlp_program.h).  Pays off when a half spans several weight stages
  // (the x3 modes, K = 64 per stage: 75.4 k -> 71.5 k cycles per tile); the 1-pass modes keep one block.
  bool split_e1;
  bool acc_flip = true;             // odd tiles use the accumulator columns XOR 128 when the program allows it
  bool view_one_half = true;        // the view step is issued as one N = W/2 half
  bool view_on_producers = false;   // the (last, one-half) view step's epilogue runs on the producer warps (mlp_program.h)
  bool out_of_fp16_range = false;   // a weight (after the feature_linear fold) exceeds 65504 or is not finite
  // INDEX MODE (device-side weight updates, weights_update.cu): the builder is run once on tensors whose VALUES are
  // their own position (+1) in the concatenation V of all input tensors followed by the derived values (the folded
  // view matrix and bias, which build_program then fills with positions instead of arithmetic).  Everything else in a
  // program's packed stream and constant table is a copy of one source value, so this run yields, per packed element,
  // where it comes from (wsrc, 0 = padding) and which 16-bit part it is (wpart) - the plan a kernel replays on device.
  bool index_mode = false;
  int64_t derived_base = 0;         // position of the first derived value in V
  std::vector<float> wsrc;
  std::vector<uint8_t> wpart;
  std::string err;

  Builder(int passes_, int fmt_) : passes(passes_), fmt(fmt_) {
    memset(&prog, 0, sizeof(prog));
    split_e1 = passes == 3;
    if (const char* v = getenv("PNR_ACC_FLIP")) acc_flip = *v != '0';   // tuning aids (A/B on the GPU)
    if (const char* v = getenv("PNR_VIEW_ONE_HALF")) view_one_half = *v != '0';
  }

  int add_consts(const float* src, int n_valid, int n_pad) {
    const int off = (int)consts.size();
    for (int i = 0; i < n_pad; ++i) consts.push_back(i < n_valid ? src[i] : 0.f);
    return off;
  }

  // Stage image of rows [row0, row0 + n_rows) x K [k0, k0 + 8*kcores): no-swizzle K-major core matrices,
  // byte offset of element (n, k) = ((k/8) * n_rows + n) * 16 + (k%8) * 2  (LBO = n_rows*16, SBO = 128).
  void pack_stage(const Mat& m, int row0, int n_rows, int col0, int kvalid, int k0, int kcores, int part) {
    const size_t base = wbuf.size();
    const int n_pad = n_rows;
    wbuf.resize(base + (size_t)n_pad * kcores * 8);
    if (index_mode) { wsrc.resize(wbuf.size(), 0.f); wpart.resize(wbuf.size(), 0); }
    for (int kc = 0; kc < kcores; ++kc)
      for (int nn = 0; nn < n_pad; ++nn)
        for (int e = 0; e < 8; ++e) {
          const int n = row0 + nn;
          const int k = k0 + kc * 8 + e;
          const float w = (n < m.out && k < kvalid) ? m.w[(size_t)n * m.in + col0 + k] : 0.f;
          if (index_mode) {
            wsrc[base + ((size_t)kc * n_pad + nn) * 8 + e] = w;
            wpart[base + ((size_t)kc * n_pad + nn) * 8 + e] = (uint8_t)part;
          } else if (!(w >= -65504.f && w <= 65504.f)) {
            out_of_fp16_range = true;   // also catches NaN
          }
          uint16_t v;
          if (fmt == 1) {
            const uint16_t h = f2bf(w);
            v = part == 0 ? h : f2bf(w - bf2f(h));
          } else {
            const uint16_t h = f2h(w);
            v = part == 0 ? h : f2h(w - h2f(h));
          }
          wbuf[base + ((size_t)kc * n_pad + nn) * 8 + e] = v;
        }
  }

  struct StepInfo { int first_stage, n_stages, n0_stage; };  // n0_stage = first stage of h1 (or end)
  std::vector<StepInfo> steps;

  // One GEMM step: acc[:, acc_col:acc_col+n_pad) = sum_seg A_seg * W_seg^T, then epilogue `ed`.
  // Issued as two N-halves (rows [0,n0) then [n0,n_pad)) when n_pad >= 64, each its own stage list.
  // `segs_h1` + `n0_split`: the second half is a GEMM of its own (own matrices, rows counted from 0, own K range) -
  // two small layers that share nothing but the step, e.g. the two logit layers (block-diagonal, zero blocks skipped).
  bool add_step(std::vector<Seg> segs, int n_pad, int acc_col, EpiDesc ed, bool first_of_tile,
                const std::vector<Seg>* segs_h1 = nullptr, int n0_split = 0) {
    if (prog.n_steps >= kMaxSteps) { err = "too many steps"; return false; }
    const int n0 = n0_split > 0 ? n0_split : (n_pad >= 64 ? n_pad / 2 : n_pad);
    const int halves = n0 < n_pad ? 2 : 1;
    StepInfo info{prog.n_stages, 0, 0};
    bool emb_waited = false;
    for (int h = 0; h < halves; ++h) {
      const int r0 = h == 0 ? 0 : n0, r1 = h == 0 ? n0 : n_pad;
      const std::vector<Seg>& hsegs = (h == 1 && segs_h1) ? *segs_h1 : segs;
      const int mrow0 = (h == 1 && segs_h1) ? 0 : r0;     // first row of the half in its weight matrix
      const int half_first = prog.n_stages;
      if (h == 1) info.n0_stage = prog.n_stages;
      for (size_t si = 0; si < hsegs.size(); ++si) {
        const Seg& sg = hsegs[si];
        // K per stage: 64 with hi+lo images (x3), 128 with the hi image only (1-pass): <= 32 KB either way
        const int chunk = passes == 3 ? 64 : 128;
        for (int k0 = 0; k0 < sg.kpad; k0 += chunk) {
          const int kcores = ((sg.kpad - k0) < chunk ? (sg.kpad - k0) : chunk) / 8;
          {
            if (prog.n_stages >= kMaxStages) { err = "too many stages"; return false; }
            const int parts = passes == 3 ? 2 : 1;
            StageDesc& sd = prog.st[prog.n_stages++];
            memset(&sd, 0, sizeof(sd));
            sd.gofs = (uint32_t)(wbuf.size() * 2);
            sd.bytes = (uint32_t)((r1 - r0) * kcores * 16 * parts);
            sd.n = (uint16_t)(r1 - r0);
            sd.acc_col = (uint16_t)(acc_col + r0);
            sd.a_off = (uint16_t)(sg.a_hi + k0 / 2);
            sd.a_lo_off = (uint16_t)(sg.a_lo + k0 / 2);
            sd.lo_off16 = (uint16_t)((r1 - r0) * kcores);
            sd.ksteps = (uint8_t)(kcores / 2);
            sd.a_kind = sg.kind;
            const bool seg_first = (k0 == 0);
            const bool seg_last = (k0 + chunk >= sg.kpad);
            const bool last_half = h == halves - 1;
            if (sg.kind == A_EMB && seg_first && first_of_tile && !emb_waited) { sd.flags |= F_WAIT_EMB; emb_waited = true; }
            if (sg.kind == A_EMB && seg_last && sg.release && last_half) sd.flags |= F_RELEASE_EMB;
            if (sg.kind == A_DIR && seg_first && h == 0) sd.flags |= F_WAIT_DIR;
            if (sg.kind == A_DIR && seg_last && sg.release && last_half) sd.flags |= F_RELEASE_DIR;
            for (int part = 0; part < parts; ++part) pack_stage(sg.m, mrow0, r1 - r0, sg.col0, sg.kvalid, k0, kcores, part);
          }
        }
      }
      prog.st[half_first].flags |= F_FIRST;
      prog.st[prog.n_stages - 1].flags |= (h == 0 ? F_COMMIT_ACC0 : F_COMMIT_ACC1);
    }
    if (halves == 1) {
      prog.st[prog.n_stages - 1].flags |= F_COMMIT_ACC1;
      info.n0_stage = prog.n_stages;
    }
    prog.st[info.first_stage].flags |= F_WAIT_E0;
    info.n_stages = prog.n_stages - info.first_stage;
    steps.push_back(info);
    ed.n = (uint16_t)n_pad;
    ed.n0 = (uint16_t)n0;
    ed.acc_col = (uint16_t)acc_col;
    const int g1 = (n_pad - n0) / 16;
    ed.n1a = (uint16_t)(n0 + ((split_e1 && g1 >= 2) ? (g1 / 2) * 16 : (n_pad - n0)));
    prog.ep[prog.n_steps++] = ed;
    return true;
  }

  static bool overlap(int a0, int a1, int b0, int b1) { return a0 < a1 && b0 < b1 && a0 < b1 && b0 < a1; }

  // Tensor-memory footprints (column intervals) used to place the cross-step hazard flags.
  struct Foot { int acc0, acc1, hi0, hi1, lo0, lo1; };
  // what the columns [c0, c1) of step s's epilogue read (accumulator) and write (activation columns)
  Foot epi_foot(int s, int c0, int c1) const {
    const EpiDesc& e = prog.ep[s];
    Foot f{e.acc_col + c0, e.acc_col + c1, 0, 0, 0, 0};
    if (epi_writes_a(e.kind) && c0 < c1) {
      f.hi0 = e.dst_col + c0 / 2; f.hi1 = e.dst_col + c1 / 2;
      if (passes == 3) { f.lo0 = e.dst_lo_col + c0 / 2; f.lo1 = e.dst_lo_col + c1 / 2; }
    }
    return f;
  }
  bool stage_touches(const StageDesc& sd, const Foot& f) const {
    const int a0 = sd.acc_col, a1 = sd.acc_col + sd.n;   // accumulator columns this stage writes
    if (overlap(a0, a1, f.acc0, f.acc1) || overlap(a0, a1, f.hi0, f.hi1) || overlap(a0, a1, f.lo0, f.lo1)) return true;
    if (sd.a_kind == A_TMEM) {
      const int h0 = sd.a_off, h1 = sd.a_off + sd.ksteps * 8;
      if (overlap(h0, h1, f.hi0, f.hi1) || overlap(h0, h1, f.acc0, f.acc1)) return true;
      if (passes == 3) {
        const int l0 = sd.a_lo_off, l1 = sd.a_lo_off + sd.ksteps * 8;
        if (overlap(l0, l1, f.lo0, f.lo1) || overlap(l0, l1, f.acc0, f.acc1)) return true;
      }
    }
    return false;
  }

  // Place the waits on the previous step's E1 parts (cyclically: the first step of a tile follows the last step
  // of the previous tile), the write-after-read commit for this step's own E0, and the per-stage hand-off counts
  // of the issue table.
  void finalize() {
    const int S = prog.n_steps;
    // accumulator flip (mlp_program.h): allowed when no activation lives inside the accumulator region (the head
    // columns of single-head programs do) and no accumulator range straddles column 128
    bool flip = acc_flip;
    for (int s = 0; s < S; ++s) flip = flip && !(epi_writes_a(prog.ep[s].kind) && prog.ep[s].dst_col < kColAHi);
    for (int i = 0; i < prog.n_stages; ++i)
      flip = flip && prog.st[i].acc_col / 128 == (prog.st[i].acc_col + prog.st[i].n - 1) / 128;
    prog.acc_flip = flip ? 1 : 0;
    // view on producers: the view step must be the tile's last, issued as one half, with at least one step before it
    const bool vp = view_on_producers && S >= 2 && prog.ep[S - 1].kind == EPI_VIEW_RGB && prog.ep[S - 1].n0 == prog.ep[S - 1].n;
    prog.view_step = vp ? S - 1 : -1;
    if (vp) {
      StageDesc& last = prog.st[steps[S - 1].first_stage + steps[S - 1].n_stages - 1];
      last.flags = (uint16_t)((last.flags & ~(F_COMMIT_ACC0 | F_COMMIT_ACC1)) | F_COMMIT_VIEW);
    }
    std::vector<int> at_a(S), at_b(S);
    for (int s = 0; s < S; ++s) {
      const StepInfo& in = steps[s];
      const int end = in.first_stage + in.n_stages;
      // the step whose E0 / E1 counts precede this one's: cyclically the tile's last step - the last one the
      // epilogue warps run, i.e. not the view step of a view-on-producers program
      const int prev_step = (s == 0) ? (vp ? S - 2 : S - 1) : s - 1;
      const EpiDesc& pe = prog.ep[prev_step];
      const bool split = in.n0_stage < end;
      auto first_touch = [&](Foot f) {
        if (s == 0 && prog.acc_flip) {   // the previous step ran in the other tile parity: its accumulator columns
          const int w = f.acc1 - f.acc0; //   are the XOR-128 image of what the program says
          f.acc0 ^= 128;
          f.acc1 = f.acc0 + w;
        }
        int at = -1;
        for (int i = in.first_stage; i < end && at < 0; ++i)
          if (stage_touches(prog.st[i], f)) at = i;
        if (at < 0) at = split ? in.n0_stage : in.first_stage;
        else if (split && at > in.n0_stage) at = in.n0_stage;   // never later than the first stage of h1
        return at;
      };
      const int prev = prev_step;
      at_b[s] = first_touch(pe.n1a < pe.n ? epi_foot(prev, pe.n1a, pe.n) : epi_foot(prev, pe.n0, pe.n));
      at_a[s] = pe.n1a < pe.n ? first_touch(epi_foot(prev, pe.n0, pe.n1a)) : at_b[s];
      if (at_a[s] > at_b[s]) at_a[s] = at_b[s];                 // "E1 done" implies "E1 part a done"
      prog.st[at_a[s]].flags |= F_WAIT_E1A;
      prog.st[at_b[s]].flags |= F_WAIT_E1;
      // E0 of this step overwrites dst columns [dst, dst + n0/2) (hi and lo): last stage reading them
      const EpiDesc& e = prog.ep[s];
      int war = in.first_stage;
      if (epi_writes_a(e.kind)) {
        Foot f = epi_foot(s, 0, e.n0);
        f.acc0 = f.acc1 = 0;   // stores only: the loads of E0 are ordered by acc_full
        for (int i = in.first_stage; i < end; ++i)
          if (stage_touches(prog.st[i], f)) war = i;
      }
      // (the epilogue warps count one write-after-read phase per step THEY run: none for a producer-run view step)
      if (!(vp && s == S - 1)) prog.st[war].flags |= F_COMMIT_WAR;
    }
    // view on producers: the first stage of the tile that touches the accumulator columns the previous tile's view
    // epilogue reads (the other tile parity's image when the program flips) and every stage after it wait for it
    int v_from = prog.n_stages;
    if (vp) {
      const EpiDesc& ve = prog.ep[S - 1];
      Foot f{ve.acc_col, ve.acc_col + ve.n, 0, 0, 0, 0};
      if (prog.acc_flip) { f.acc0 ^= 128; f.acc1 = f.acc0 + ve.n; }
      for (int i = 0; i < prog.n_stages && v_from == prog.n_stages; ++i)
        if (stage_touches(prog.st[i], f)) v_from = i;
      // no stage of the next tile touches them (narrow networks in a flipping program: the columns come up again one
      // tile later): gate the whole next tile - the count is per tile, and a done epilogue stays done
      if (v_from == prog.n_stages) v_from = 0;
    }
    for (int s = 0; s < S; ++s) {
      const StepInfo& in = steps[s];
      for (int i = in.first_stage; i < in.first_stage + in.n_stages; ++i)
        prog.is[i].needs = (uint32_t)(s + 1) | ((uint32_t)(i >= at_a[s] ? s + 1 : s) << 8) |
                           ((uint32_t)(i >= at_b[s] ? s + 1 : s) << 16) | ((uint32_t)(i >= v_from ? 1 : 0) << 24);
    }
    for (int i = 0; i < prog.n_stages; ++i) {   // issue table (flags are final now)
      const StageDesc& sd = prog.st[i];
      IssueDesc& d = prog.is[i];
      const uint32_t rows = sd.n;                         // rows of the weight tile
      d.idesc = make_idesc_f32acc(kTileM, sd.n, fmt);     // fmt: 0 = fp16, 1 = bf16
      d.b_lo_base = (uint32_t)(((rows * 16u) >> 4) & 0x3FFFu) << 16;
      d.b_inc = (2u * rows * 16u) >> 4;
      d.lo_off16 = sd.lo_off16;
      d.acc_col = sd.acc_col;
      d.a_off = sd.a_off;
      d.a_lo_off = sd.a_lo_off;
      d.flags_k = (uint32_t)sd.flags | ((uint32_t)sd.ksteps << 16) | ((uint32_t)sd.a_kind << 24);
    }
  }
};

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }


}  // namespace

extern "C" int pnr_version(void) { return PNR_VERSION; }
extern "C" const char* pnr_last_error(void) { return g_err; }
extern "C" int64_t pnr_launch_count(int32_t reset) {
  const int64_t v = g_launches;
  if (reset) g_launches = 0;
  return v;
}

static int precision_passes(int precision) {
  return (precision == PNR_PREC_BF16X3 || precision == PNR_PREC_FP16X3) ? 3 : 1;
}
static int precision_fmt(int precision) {   // instruction-descriptor operand format: 0 = fp16, 1 = bf16
  return (precision == PNR_PREC_BF16X3 || precision == PNR_PREC_BF16) ? 1 : 0;
}

static int check_config(const pnr_config* cfg) {
  PNR_CHECK_ARG(cfg->D >= 3 && cfg->D <= 16, "pnr_create: D=%d outside [3,16]", cfg->D);
  PNR_CHECK_ARG(cfg->W == 64 || cfg->W == 128 || cfg->W == 256, "pnr_create: W=%d not in {64,128,256}", cfg->W);
  PNR_CHECK_ARG(cfg->xyz_res >= 0 && cfg->xyz_res <= 10, "pnr_create: xyz_res=%d outside [0,10]", cfg->xyz_res);
  PNR_CHECK_ARG(cfg->view_res >= 0 && cfg->view_res <= 4, "pnr_create: view_res=%d outside [0,4]", cfg->view_res);
  PNR_CHECK_ARG(cfg->num_classes >= 0 && cfg->num_classes <= 128, "pnr_create: num_classes=%d outside [0,128]", cfg->num_classes);
  PNR_CHECK_ARG(cfg->num_instances >= 0 && cfg->num_instances <= 128, "pnr_create: num_instances=%d outside [0,128]", cfg->num_instances);
  PNR_CHECK_ARG(cfg->precision >= 0 && cfg->precision <= 3, "pnr_create: bad precision %d", cfg->precision);
  return PNR_OK;
}

extern "C" int pnr_create(const pnr_config* cfg, pnr_ctx** out) {
  PNR_CHECK_ARG(cfg && out, "pnr_create: null pointer");
  if (const int rc = check_config(cfg)) return rc;
  int ndev = 0;
  PNR_CUDA(cudaGetDeviceCount(&ndev));
  PNR_CHECK_ARG(cfg->device >= 0 && cfg->device < ndev, "pnr_create: device %d of %d", cfg->device, ndev);
  cudaDeviceProp prop;
  PNR_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10)
    return set_error(PNR_ERR_UNSUPPORTED, "pnr_create: device %d is sm_%d%d; libpnr is sm_100a only (no fallback path)",
                     cfg->device, prop.major, prop.minor);
  PNR_CHECK_ARG(cfg->device < kMaxDevices, "pnr_create: device ordinal %d >= %d", cfg->device, kMaxDevices);
  DeviceGuard guard(cfg->device);   // the caller's current device is restored on return
  pnr_ctx* c = new pnr_ctx();
  c->cfg = *cfg;
  c->passes = precision_passes(cfg->precision);
  c->fmt = precision_fmt(cfg->precision);
  memset(&c->launch, 0, sizeof(c->launch));
  memset(&c->launch_vp, 0, sizeof(c->launch_vp));
  cudaError_t e = cudaMalloc(&c->d_status, sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMemset(c->d_status, 0, sizeof(uint32_t));
  if (e != cudaSuccess) {
    delete c;
    return set_error(PNR_ERR_CUDA, "pnr_create: status word: %s", cudaGetErrorString(e));
  }
  *out = c;
  return PNR_OK;
}

extern "C" int pnr_destroy(pnr_ctx* ctx) {
  if (!ctx) return PNR_OK;
  DeviceGuard guard(ctx->cfg.device);
  cudaFree(ctx->d_wpacked);
  cudaFree(ctx->d_consts);
  cudaFree(ctx->bwd.d_wpacked);
  cudaFree(ctx->bwd.d_consts);
  cudaFree(ctx->trunk_fwd.d_wpacked);
  cudaFree(ctx->trunk_fwd.d_consts);
  for (pnr_ctx::Plan* pl : {&ctx->plan_main, &ctx->plan_bwd, &ctx->plan_tf}) {
    cudaFree(pl->d_widx); cudaFree(pl->d_wpart); cudaFree(pl->d_cidx);
  }
  cudaFree(ctx->d_V);
  cudaFree(ctx->d_status);
  delete ctx;
  return PNR_OK;
}

// Sticky status of the fused MLP launches enqueued so far on `stream` (synchronises that stream): bit 0 = an
// activation left the range of the 16-bit operand format (fp16 modes: |x| > 65504) or was not finite - the
// results of that launch are not trustworthy; re-run with PNR_PREC_BF16X3.  reset != 0 clears the word.
extern "C" int pnr_status(pnr_ctx* ctx, uint32_t* status_host, int32_t reset, void* stream) {
  PNR_CHECK_ARG(ctx && status_host, "pnr_status: null pointer");
  DeviceGuard guard(ctx->cfg.device);
  cudaStream_t st = (cudaStream_t)stream;
  PNR_CUDA(cudaMemcpyAsync(status_host, ctx->d_status, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  if (reset) PNR_CUDA(cudaMemsetAsync(ctx->d_status, 0, sizeof(uint32_t), st));
  PNR_CUDA(cudaStreamSynchronize(st));
  return PNR_OK;
}

// Host only (no CUDA call): checks the tensor list against cfg, builds the per-tile program, the packed
// weight stream and the constant table.  Shared by pnr_load_weights and pnr_program_host.
static int build_program(const pnr_config& c, const float* const* t, const int64_t* shapes, int32_t n,
                         Builder& bld) {
  const int D = c.D, W = c.W, W2 = W / 2, C = c.num_classes, K = c.num_instances;
  const int Ex = 3 + 6 * c.xyz_res, Ed = 3 + 6 * c.view_res, skip = D / 2;
  const int expected = 2 * D + 8 + (C > 0 ? 4 : 0) + (K > 0 ? 4 : 0);
  PNR_CHECK_ARG(n == expected, "pnr_load_weights: got %d tensors, expected %d", n, expected);
  int ti = 0;
  auto take = [&](int out, int in, Mat* m, const float** bias) -> bool {
    if (shapes[2 * ti] != out || shapes[2 * ti + 1] != in) return false;
    if (shapes[2 * ti + 2] != out || shapes[2 * ti + 3] != 1) return false;
    if (!t[ti] || !t[ti + 1]) return false;
    *m = Mat{t[ti], out, in};
    *bias = t[ti + 1];
    ti += 2;
    return true;
  };
#define PNR_TAKE(out, in, m, b)                                                                        \
  if (!take(out, in, m, b))                                                                            \
    return set_error(PNR_ERR_ARG, "pnr_load_weights: tensor %d: expected weight [%d,%d] + bias [%d,1]", \
                     ti, out, in, out)

  bld.prog.Lx = c.xyz_res;
  bld.prog.Ld = c.view_res;
  bld.prog.passes = bld.passes;

  std::vector<Mat> trunk(D);
  std::vector<const float*> trunk_b(D);
  for (int i = 0; i < D; ++i) {
    const int in = i == 0 ? Ex : (i == skip + 1 ? W + Ex : W);
    PNR_TAKE(W, in, &trunk[i], &trunk_b[i]);
  }
  Mat m_sig, m_feat, m_view, m_rgb, m_s1, m_s2, m_i1, m_i2;
  const float *b_sig, *b_feat, *b_view, *b_rgb, *b_s1 = nullptr, *b_s2 = nullptr, *b_i1 = nullptr, *b_i2 = nullptr;
  PNR_TAKE(1, W, &m_sig, &b_sig);
  PNR_TAKE(W, W, &m_feat, &b_feat);
  PNR_TAKE(W2, W + Ed, &m_view, &b_view);
  PNR_TAKE(3, W2, &m_rgb, &b_rgb);
  if (C > 0) { PNR_TAKE(W2, W, &m_s1, &b_s1); PNR_TAKE(C, W2, &m_s2, &b_s2); }
  if (K > 0) { PNR_TAKE(W2, W, &m_i1, &b_i1); PNR_TAKE(K, W2, &m_i2, &b_i2); }
#undef PNR_TAKE

  const int sig_w_off = bld.add_consts(m_sig.w, W, W);
  bld.prog.sigma_bias_off = bld.add_consts(b_sig, 1, 4);
  std::vector<float> rgbw(3 * W2);
  for (int ch = 0; ch < 3; ++ch)
    for (int k = 0; k < W2; ++k) rgbw[ch * W2 + k] = m_rgb.w[ch * W2 + k];
  const int rgb_w_off = bld.add_consts(rgbw.data(), 3 * W2, 3 * W2);
  bld.prog.rgb_bias_off = bld.add_consts(b_rgb, 3, 4);

  auto seg_tmem = [&](const Mat& m, int col0, int k, int a_hi, int a_lo) {
    return Seg{A_TMEM, m, col0, k, round_up(k, 16), a_hi, a_lo, false};
  };
  bool ok = true;
  // trunk
  for (int i = 0; i < D && ok; ++i) {
    EpiDesc ed{};
    ed.kind = EPI_RELU_TO_A;
    ed.sigma = (i == D - 1) ? 1 : 0;
    ed.dst_col = kColAHi;
    ed.dst_lo_col = kColALo;
    ed.bias_off = (uint16_t)bld.add_consts(trunk_b[i], W, W);
    ed.aux_off = (uint16_t)sig_w_off;
    std::vector<Seg> segs;
    if (i == 0) {
      segs.push_back(Seg{A_EMB, trunk[i], 0, Ex, 64, 0, 0, false});
    } else if (i == skip + 1) {
      segs.push_back(Seg{A_EMB, trunk[i], 0, Ex, 64, 0, 0, true});
      segs.push_back(seg_tmem(trunk[i], Ex, W, kColAHi, kColALo));
    } else {
      segs.push_back(seg_tmem(trunk[i], 0, W, kColAHi, kColALo));
    }
    ok = bld.add_step(segs, W, kColAcc, ed, i == 0);
  }
  // feature_linear has no activation, so it is folded into the view layer when the weights are loaded
  // (exact algebra, done in double):  W_view [feat ; gamma(d)] + b_view  with  feat = W_feat h + b_feat
  //   = (W_view[:, :W] W_feat) h + W_view[:, W:] gamma(d) + (W_view[:, :W] b_feat + b_view).
  // One 256x256 GEMM per sample (11 % of the MLP) and its epilogue disappear; the view step reads the trunk
  // output h directly.
  std::vector<float> fold((size_t)W2 * (W + Ed)), fold_b(W2);
  if (bld.index_mode) {   // derived values: positions in V's derived area (filled on device by fold_kernel)
    for (size_t i = 0; i < fold.size(); ++i) fold[i] = (float)(bld.derived_base + (int64_t)i + 1);
    for (int n_ = 0; n_ < W2; ++n_) fold_b[n_] = (float)(bld.derived_base + (int64_t)fold.size() + n_ + 1);
  }
  for (int n_ = 0; n_ < W2 && !bld.index_mode; ++n_) {
    const float* vrow = m_view.w + (size_t)n_ * (W + Ed);
    for (int k = 0; k < W; ++k) {
      double acc = 0.0;
      for (int j = 0; j < W; ++j) acc += (double)vrow[j] * (double)m_feat.w[(size_t)j * W + k];
      fold[(size_t)n_ * (W + Ed) + k] = (float)acc;
    }
    for (int e = 0; e < Ed; ++e) fold[(size_t)n_ * (W + Ed) + W + e] = vrow[W + e];
    double accb = (double)b_view[n_];
    for (int j = 0; j < W; ++j) accb += (double)vrow[j] * (double)b_feat[j];
    fold_b[n_] = (float)accb;
  }
  const Mat m_fold{fold.data(), W2, W + Ed};
  auto add_view = [&](int acc_col) {  // view branch [h, gamma(d)] -> relu -> rgb (CUDA cores) ; writes rgb + sigma
    EpiDesc ed{};
    ed.kind = EPI_VIEW_RGB;
    ed.bias_off = (uint16_t)bld.add_consts(fold_b.data(), W2, W2);
    ed.aux_off = (uint16_t)rgb_w_off;
    std::vector<Seg> segs;
    segs.push_back(seg_tmem(m_fold, 0, W, kColAHi, kColALo));
    segs.push_back(Seg{A_DIR, m_fold, W, Ed, 32, 0, 0, true});
    // (Accumulating the view step in the UPPER half of the accumulator region, so that the next tile's first layer
    // can be issued right behind the view MMAs, was measured: the stall only moves in front of the view step -
    // the tile boundary is bound by the serial epilogues of view + layer 0, 71.6 k vs 71.9 k cycles per tile.)
    // ONE N = W/2 half: as two N = 64 halves the view step's 108 MMAs cost ~66 cycles each for half the columns
    // (an M=128 N=64 K=16 MMA is no faster than 0.84 of an N=128 one: profiles/r01_mma_rate_probe.log) - 7-8 k cycles
    // of tensor pipe at the one place of the tile where nothing else can be issued.  What the halves bought, the view
    // epilogue's first part overlapping the second half's MMAs, is worth less than that.
    ok = ok && bld.add_step(segs, W2, acc_col, ed, false, nullptr, (bld.view_one_half && W2 <= 128) ? W2 : 0);
  };
  // heads: hidden layer -> logits
  auto add_head = [&](const Mat& m1, const float* b1, const Mat& m2, const float* b2, int nout, int out_off) {
    EpiDesc e1{};
    e1.kind = EPI_RELU_TO_A;
    e1.dst_col = kColHeadHi;        // head hidden activations (K <= 128) live in the upper half of the
    e1.dst_lo_col = kColHeadLo;     // accumulator region while it is free
    e1.bias_off = (uint16_t)bld.add_consts(b1, W2, W2);
    ok = ok && bld.add_step({seg_tmem(m1, 0, W, kColAHi, kColALo)}, W2, kColAcc, e1, false);
    const int npad = round_up(nout, 16);
    EpiDesc e2{};
    e2.kind = EPI_LOGITS;
    e2.n_valid = (uint16_t)nout;
    e2.out_off = (uint16_t)out_off;
    e2.bias_off = (uint16_t)bld.add_consts(b2, nout, npad);
    ok = ok && bld.add_step({seg_tmem(m2, 0, W2, kColHeadHi, kColHeadLo)}, npad, kColAcc, e2, false);
  };
  std::vector<float> hid_w, hid_b;   // must outlive add_step (Mat holds a pointer)
  if (C > 0 && K > 0 && 2 * W2 == W && W >= 128) {
    // Both heads: the view step runs first (it is the last reader of the trunk output), then ONE W-wide hidden
    // step computes both heads' hidden layers ([W_s1 ; W_i1], ReLU) in place of the trunk activations - a full-size
    // layer at trunk efficiency instead of two N = W/2 steps whose halves are issue-bound - and ONE logits step whose
    // halves are the two (block-diagonal) logit layers, zero blocks skipped: h0 = semantic logits from hidden
    // columns [0, W/2), h1 = instance logits from [W/2, W).  Per tile: 11 steps instead of 13, and the serial chain
    // hidden -> epilogue -> logits -> epilogue runs once, not twice (r2 timeline: ~30 k of a cfg3 tile's ~99 k
    // cycles went into the two head chains for ~10 k cycles of tensor work).
    // (The view step accumulating in the UPPER accumulator half, so that the hidden step's first half can start right
    // behind the view MMAs, was measured on the cfg3 frame: 411.8 vs 408.9 ms - no gain, not kept.)
    add_view(kColAcc);
    hid_w.resize((size_t)W * W);
    hid_b.resize(W);
    memcpy(hid_w.data(), m_s1.w, sizeof(float) * (size_t)W2 * W);
    memcpy(hid_w.data() + (size_t)W2 * W, m_i1.w, sizeof(float) * (size_t)W2 * W);
    memcpy(hid_b.data(), b_s1, sizeof(float) * W2);
    memcpy(hid_b.data() + W2, b_i1, sizeof(float) * W2);
    const Mat m_hid{hid_w.data(), W, W};
    EpiDesc eh{};
    eh.kind = EPI_RELU_TO_A;
    eh.dst_col = kColAHi;
    eh.dst_lo_col = kColALo;
    eh.bias_off = (uint16_t)bld.add_consts(hid_b.data(), W, W);
    ok = ok && bld.add_step({seg_tmem(m_hid, 0, W, kColAHi, kColALo)}, W, kColAcc, eh, false);
    const int n_s = round_up(C, 16), n_i = round_up(K, 16);
    EpiDesc el{};
    el.kind = EPI_LOGITS;
    el.n_valid = (uint16_t)C;
    el.out_off = 4;
    el.n_valid1 = (uint16_t)K;
    el.out_off1 = (uint16_t)(4 + C);
    el.bias_off = (uint16_t)bld.add_consts(b_s2, C, n_s);
    bld.add_consts(b_i2, K, n_i);                         // contiguous: bias of column n_s + j
    const std::vector<Seg> seg_i{seg_tmem(m_i2, 0, W2, kColAHi + W2 / 2, kColALo + W2 / 2)};
    ok = ok && bld.add_step({seg_tmem(m_s2, 0, W2, kColAHi, kColALo)}, n_s + n_i, kColAcc, el, false, &seg_i, n_s);
  } else {
    if (ok && C > 0) add_head(m_s1, b_s1, m_s2, b_s2, C, 4);
    if (ok && K > 0) add_head(m_i1, b_i1, m_i2, b_i2, K, 4 + C);
    add_view(kColAcc);
  }
  if (!ok) return set_error(PNR_ERR_UNSUPPORTED, "pnr_load_weights: program build failed: %s", bld.err.c_str());
  if ((int)bld.consts.size() > kMaxConsts)
    return set_error(PNR_ERR_UNSUPPORTED, "pnr_load_weights: %d constants > %d", (int)bld.consts.size(), kMaxConsts);
  if (bld.out_of_fp16_range && bld.fmt == kFmtF16)
    return set_error(PNR_ERR_UNSUPPORTED, "pnr_load_weights: a weight is outside the fp16 range (|w| > 65504 or not "
                     "finite): use precision bf16x3");
  bld.prog.n_consts = (int)bld.consts.size();
  bld.finalize();
  return PNR_OK;
}


// Backward program of the trunk (mlp_program.h, "BACKWARD"): t / shapes = the trunk's weight, bias pairs (the first
// 2*D tensors of pnr_load_weights' list).  Forward steps 0..D-1 (sign patterns kept, the last one loads the incoming
// gradient), then for l = D-1..0 the gradient w.r.t. layer l's input: g_l [128, W] x W_l [W, in_l], i.e. a step whose
// weight matrix is W_l transposed; the embedded-input columns of layer 0 and of the skip layer go to the output rows.
// forward_only: just the trunk, whose last layer writes its activations to the output rows (EPI_ACT_OUT).
// Stash slots (MlpParams::stash): H_i of forward layer i < D-1 in slot i; the pre-activation gradient dZ_j in slot
// 2D-2-j (the order they are produced in: dZ_{D-1} by the last forward layer's epilogue, then D-2 .. 0).
static int build_backward_program(const pnr_config& c, const float* const* t, const int64_t* shapes, int32_t n,
                                  Builder& bld, bool forward_only = false) {
  const int D = c.D, W = c.W, Ex = 3 + 6 * c.xyz_res, skip = D / 2;
  PNR_CHECK_ARG(n >= 2 * D, "backward program: got %d tensors, the trunk has %d", n, 2 * D);
  if (bld.passes != 3)
    return set_error(PNR_ERR_UNSUPPORTED, "backward program: precision must be fp16x3 or bf16x3");
  if (D - 1 > kMaxMaskSlots)
    return set_error(PNR_ERR_UNSUPPORTED, "backward program: D=%d needs %d sign-pattern slots, %d fit", D, D - 1, kMaxMaskSlots);
  std::vector<Mat> trunk(D);
  std::vector<const float*> trunk_b(D);
  for (int i = 0; i < D; ++i) {
    const int in = i == 0 ? Ex : (i == skip + 1 ? W + Ex : W);
    if (shapes[4 * i] != W || shapes[4 * i + 1] != in || shapes[4 * i + 2] != W || shapes[4 * i + 3] != 1 || !t[2 * i] || !t[2 * i + 1])
      return set_error(PNR_ERR_ARG, "backward program: trunk layer %d: expected weight [%d,%d] + bias [%d,1]", i, W, in, W);
    trunk[i] = Mat{t[2 * i], W, in};
    trunk_b[i] = t[2 * i + 1];
  }
  bld.prog.Lx = c.xyz_res;
  bld.prog.Ld = c.view_res;
  bld.prog.passes = bld.passes;
  auto seg_tmem = [&](const Mat& m, int col0, int k) {
    return Seg{A_TMEM, m, col0, k, round_up(k, 16), kColAHi, kColALo, false};
  };
  bool ok = true;
  for (int i = 0; i < D && ok; ++i) {   // forward, as in build_program (no sigma head)
    EpiDesc ed{};
    ed.kind = (i == D - 1) ? (forward_only ? EPI_ACT_OUT : EPI_LOADG_TO_A) : EPI_RELU_TO_A;
    ed.n_valid = (i == D - 1 || forward_only) ? 0 : (uint16_t)(i + 1);      // sign-pattern slot + 1
    ed.out_off1 = forward_only ? 0 : (uint16_t)(i + 1);                       // stash slot + 1: H_i, or dZ_{D-1} for i = D-1
    ed.dst_col = kColAHi;
    ed.dst_lo_col = kColALo;
    ed.bias_off = (uint16_t)bld.add_consts(trunk_b[i], W, W);
    std::vector<Seg> segs;
    if (i == 0) {
      segs.push_back(Seg{A_EMB, trunk[i], 0, Ex, 64, 0, 0, false});
    } else if (i == skip + 1) {
      segs.push_back(Seg{A_EMB, trunk[i], 0, Ex, 64, 0, 0, true});
      segs.push_back(seg_tmem(trunk[i], Ex, W));
    } else {
      segs.push_back(seg_tmem(trunk[i], 0, W));
    }
    ok = bld.add_step(segs, W, kColAcc, ed, i == 0);
  }
  std::vector<float> wt;   // W_l transposed: [in_l, W] row-major (packed inside add_step, so one buffer serves all)
  for (int l = forward_only ? -1 : D - 1; l >= 0 && ok; --l) {
    const int in = trunk[l].in;
    wt.assign((size_t)in * W, 0.f);
    for (int o = 0; o < W; ++o)
      for (int k = 0; k < in; ++k) wt[(size_t)k * W + o] = trunk[l].w[(size_t)o * in + k];
    auto grad_out = [&](const float* rows, bool accumulate) {    // embedded-input columns -> output rows
      EpiDesc eo{};
      eo.kind = EPI_GRAD_OUT;
      eo.n_valid = (uint16_t)Ex;
      eo.n_valid1 = accumulate ? 1 : 0;
      eo.out_off = 0;
      // one N = 64 half: two N = 32 halves would double the MMA count for the same tensor time per MMA
      return bld.add_step({seg_tmem(Mat{rows, Ex, W}, 0, W)}, 64, kColAcc, eo, false, nullptr, 64);
    };
    auto grad_h = [&](const float* rows) {                       // hidden columns, gated by layer l-1's sign pattern
      EpiDesc em{};
      em.kind = EPI_MASK_TO_A;
      em.n_valid = (uint16_t)l;                                  // slot (l - 1) + 1
      em.out_off1 = (uint16_t)(2 * D - 2 - (l - 1) + 1);          // stash slot + 1 of dZ_{l-1}
      em.dst_col = kColAHi;
      em.dst_lo_col = kColALo;
      return bld.add_step({seg_tmem(Mat{rows, W, W}, 0, W)}, W, kColAcc, em, false);
    };
    if (l == 0) {
      ok = grad_out(wt.data(), true);
    } else if (l == skip + 1) {   // input = [embedded xyz ; h]: the embedded part first (the next step overwrites g_l)
      ok = grad_out(wt.data(), false) && grad_h(wt.data() + (size_t)Ex * W);
    } else {
      ok = grad_h(wt.data());
    }
  }
  if (!ok) return set_error(PNR_ERR_UNSUPPORTED, "backward program: build failed: %s", bld.err.c_str());
  if ((int)bld.consts.size() > kMaxConsts)
    return set_error(PNR_ERR_UNSUPPORTED, "backward program: %d constants > %d", (int)bld.consts.size(), kMaxConsts);
  if (bld.out_of_fp16_range && bld.fmt == kFmtF16)
    return set_error(PNR_ERR_UNSUPPORTED, "backward program: a weight is outside the fp16 range: use precision bf16x3");
  bld.prog.n_consts = (int)bld.consts.size();
  bld.finalize();
  return PNR_OK;
}

extern "C" int pnr_load_weights(pnr_ctx* ctx, const float* const* t, const int64_t* shapes, int32_t n) {
  PNR_CHECK_ARG(ctx && t && shapes, "pnr_load_weights: null pointer");
  const pnr_config& c = ctx->cfg;
  Builder bld(ctx->passes, ctx->fmt);
  const int rc = build_program(c, t, shapes, n, bld);
  if (rc != PNR_OK) return rc;
  ctx->has_vp = false;
  if (c.num_classes == 0 && c.num_instances == 0 && !(getenv("PNR_VIEW_PRODUCERS") && *getenv("PNR_VIEW_PRODUCERS") == '0')) {
    Builder vp(ctx->passes, ctx->fmt);
    vp.view_on_producers = true;
    if (build_program(c, t, shapes, n, vp) == PNR_OK && vp.prog.view_step >= 0 && vp.wbuf == bld.wbuf &&
        vp.consts == bld.consts) {   // same packed stream and constants: only flags and hand-off counts differ
      ctx->launch_vp.prog = vp.prog;
      ctx->has_vp = true;
    }
  }

  DeviceGuard guard(c.device);
  // (plain cudaFree / cudaMemcpy: they synchronise with the device, so no launch still reads the old buffers)
  cudaFree(ctx->d_wpacked); cudaFree(ctx->d_consts);
  ctx->d_wpacked = nullptr; ctx->d_consts = nullptr;
  ctx->loaded = false;
  ctx->wpacked_bytes = bld.wbuf.size() * 2;
  PNR_CUDA(cudaMalloc(&ctx->d_wpacked, ctx->wpacked_bytes));
  PNR_CUDA(cudaMalloc(&ctx->d_consts, bld.consts.size() * 4));
  ctx->launch.prog = bld.prog;
  // host copy of the trunk for the backward program (built on the first pnr_mlp_backward_trunk after this load)
  ctx->bwd.ready = ctx->trunk_fwd.ready = false;
  for (pnr_ctx::Plan* pl : {&ctx->plan_main, &ctx->plan_bwd, &ctx->plan_tf}) {
    cudaFree(pl->d_widx); cudaFree(pl->d_wpart); cudaFree(pl->d_cidx);
    *pl = pnr_ctx::Plan();
  }
  cudaFree(ctx->d_V);
  ctx->d_V = nullptr;
  ctx->device_weights = false;
  ctx->all_shapes.assign(shapes, shapes + 2 * n);
  ctx->v_off.assign(n, 0);
  ctx->v_total = 0;
  for (int i = 0; i < n; ++i) { ctx->v_off[i] = ctx->v_total; ctx->v_total += shapes[2 * i] * shapes[2 * i + 1]; }
  ctx->v_derived = (int64_t)(c.W / 2) * (c.W + 3 + 6 * c.view_res) + c.W / 2;   // folded view matrix + bias
  ctx->host_trunk.clear();
  ctx->host_trunk_shapes.assign(shapes, shapes + 4 * c.D);
  for (int i = 0; i < 2 * c.D; ++i)
    ctx->host_trunk.emplace_back(t[i], t[i] + (size_t)shapes[2 * i] * (size_t)shapes[2 * i + 1]);
  PNR_CUDA(cudaMemcpy(ctx->d_wpacked, bld.wbuf.data(), ctx->wpacked_bytes, cudaMemcpyHostToDevice));
  PNR_CUDA(cudaMemcpy(ctx->d_consts, bld.consts.data(), bld.consts.size() * 4, cudaMemcpyHostToDevice));
  ctx->loaded = true;
  return PNR_OK;
}

extern "C" int pnr_program_host(const pnr_config* cfg, const float* const* t, const int64_t* shapes, int32_t n,
                                int32_t flags, void* program, size_t program_cap, size_t* program_bytes, void* wpacked,
                                size_t wpacked_cap, size_t* wpacked_bytes, float* consts, size_t consts_cap,
                                size_t* n_consts) {
  PNR_CHECK_ARG(cfg && t && shapes && program_bytes && wpacked_bytes && n_consts, "pnr_program_host: null pointer");
  if (const int rc = check_config(cfg)) return rc;
  PNR_CHECK_ARG((flags & ~(PNR_PROGRAM_SPLIT_E1 | PNR_PROGRAM_NO_SPLIT | PNR_PROGRAM_BACKWARD | PNR_PROGRAM_VIEW_PRODUCERS)) == 0,
                "pnr_program_host: unknown flags 0x%x", flags);
  Builder bld(precision_passes(cfg->precision), precision_fmt(cfg->precision));
  if (flags & PNR_PROGRAM_NO_SPLIT) bld.split_e1 = false;
  if (flags & PNR_PROGRAM_SPLIT_E1) bld.split_e1 = true;
  if (flags & PNR_PROGRAM_VIEW_PRODUCERS) bld.view_on_producers = true;
  const int rc = (flags & PNR_PROGRAM_BACKWARD) ? build_backward_program(*cfg, t, shapes, n, bld)
                                                : build_program(*cfg, t, shapes, n, bld);
  if (rc != PNR_OK) return rc;
  *program_bytes = sizeof(MlpProgram);
  *wpacked_bytes = bld.wbuf.size() * 2;
  *n_consts = bld.consts.size();
  if (program) {
    PNR_CHECK_ARG(program_cap >= sizeof(MlpProgram), "pnr_program_host: program buffer too small");
    memcpy(program, &bld.prog, sizeof(MlpProgram));
  }
  if (wpacked) {
    PNR_CHECK_ARG(wpacked_cap >= *wpacked_bytes, "pnr_program_host: weight buffer too small");
    memcpy(wpacked, bld.wbuf.data(), *wpacked_bytes);
  }
  if (consts) {
    PNR_CHECK_ARG(consts_cap >= *n_consts, "pnr_program_host: constant buffer too small");
    memcpy(consts, bld.consts.data(), *n_consts * 4);
  }
  return PNR_OK;
}

static int mlp_forward_impl(pnr_ctx* ctx, const float* pts, const float* viewdirs, const float* rays,
                            const float* z, int64_t R, int32_t N, float* raw, void* stream, long long* dbg);

extern "C" int pnr_debug_timeline(pnr_ctx* ctx, int64_t* timeline) {
  PNR_CHECK_ARG(ctx, "pnr_debug_timeline: null context");
  ctx->dbg_timeline = (long long*)timeline;
  return PNR_OK;
}

extern "C" int pnr_mlp_forward(pnr_ctx* ctx, const float* pts, const float* viewdirs, const float* rays,
                               const float* z, int64_t R, int32_t N, float* raw, void* stream) {
  return mlp_forward_impl(ctx, pts, viewdirs, rays, z, R, N, raw, stream, nullptr);
}

extern "C" int pnr_mlp_forward_timeline(pnr_ctx* ctx, const float* rays, const float* z, int64_t R, int32_t N,
                                        float* raw, int64_t* timeline, void* stream) {
  return mlp_forward_impl(ctx, nullptr, nullptr, rays, z, R, N, raw, stream, (long long*)timeline);
}

static int mlp_forward_impl(pnr_ctx* ctx, const float* pts, const float* viewdirs, const float* rays,
                            const float* z, int64_t R, int32_t N, float* raw, void* stream, long long* dbg) {
  if (R == 0) return PNR_OK;
  PNR_CHECK_ARG(ctx && raw, "pnr_mlp_forward: null pointer");
  if (!ctx->loaded) return set_error(PNR_ERR_STATE, "pnr_mlp_forward: pnr_load_weights has not been called");
  PNR_CHECK_ARG(R >= 0 && N >= 1, "pnr_mlp_forward: bad sizes R=%lld N=%d", (long long)R, N);
  PNR_CHECK_ARG((pts && viewdirs) || (!pts && rays && z), "pnr_mlp_forward: need (pts, viewdirs) or (rays, z)");
  const int64_t S = R * (int64_t)N;
  if (S == 0) return PNR_OK;
  PNR_CHECK_ARG((S + kTileM - 1) / kTileM < (int64_t)1 << 31, "pnr_mlp_forward: too many samples");
  MlpLaunch& L = ctx->has_vp ? ctx->launch_vp : ctx->launch;
  MlpParams& p = L.p;
  memset(&p, 0, sizeof(p));
  p.wpacked = ctx->d_wpacked; p.consts = ctx->d_consts;
  p.pts = pts; p.viewdirs = viewdirs; p.rays = rays; p.z = z;
  p.S = S; p.N = N; p.CH = 4 + ctx->cfg.num_classes + ctx->cfg.num_instances; p.raw = raw;
  p.num_tiles = (int32_t)((S + kTileM - 1) / kTileM);
  p.status = ctx->d_status;
  p.dbg = dbg;
  DeviceGuard guard(ctx->cfg.device);   // launch on the context's device whatever the caller's current one is
  return launch_mlp(L, ctx->passes, ctx->fmt, ctx->has_vp ? kMlpForwardVP : kMlpForward, (cudaStream_t)stream);
}

extern "C" int pnr_mlp_composite(pnr_ctx* ctx, const float* rays, const float* z, int64_t R, int32_t N,
                                 int32_t white_bkgd, int32_t mask_outside, const int32_t* sample_box,
                                 const int32_t* box_sem, const int32_t* box_inst, int32_t B,
                                 const pnr_composite_out* out, void* stream) {
  if (R == 0) return PNR_OK;
  PNR_CHECK_ARG(ctx && rays && z && out, "pnr_mlp_composite: null pointer");
  if (!ctx->loaded) return set_error(PNR_ERR_STATE, "pnr_mlp_composite: pnr_load_weights has not been called");
  PNR_CHECK_ARG(R > 0 && N >= 1, "pnr_mlp_composite: bad sizes R=%lld N=%d", (long long)R, N);
  if (N % 32 != 0)
    return set_error(PNR_ERR_UNSUPPORTED, "pnr_mlp_composite: N=%d is not a multiple of 32 (use pnr_mlp_forward + pnr_composite)", N);
  PNR_CHECK_ARG(out->weights, "pnr_mlp_composite: out->weights is required");
  PNR_CHECK_ARG(!mask_outside || sample_box, "pnr_mlp_composite: mask_outside needs sample_box");
  const int C = ctx->cfg.num_classes, K = ctx->cfg.num_instances;
  const int64_t S = R * (int64_t)N;
  PNR_CHECK_ARG((S + kTileM - 1) / kTileM < (int64_t)1 << 31, "pnr_mlp_composite: too many samples");
  MlpParams& p = ctx->launch.p;
  p.wpacked = ctx->d_wpacked; p.consts = ctx->d_consts;
  p.pts = nullptr; p.viewdirs = nullptr; p.rays = rays; p.z = z;
  p.S = S; p.N = N; p.CH = 4 + C + K; p.raw = nullptr;
  p.num_tiles = (int32_t)((S + kTileM - 1) / kTileM);
  p.status = ctx->d_status;
  p.dbg = ctx->dbg_timeline;
  p.sample_box = sample_box; p.mask_outside = mask_outside; p.white_bkgd = white_bkgd;
  p.C = C; p.K = K;
  p.weights = out->weights; p.rgb_map = out->rgb_map; p.depth_map = out->depth_map; p.acc_map = out->acc_map;
  p.disp_map = out->disp_map; p.sem_map = C > 0 ? out->semantic_map : nullptr; p.inst_map = K > 0 ? out->instance_map : nullptr;
  DeviceGuard guard(ctx->cfg.device);
  if (const int rc = launch_mlp(ctx->launch, ctx->passes, ctx->fmt, kMlpComposite, (cudaStream_t)stream)) return rc;
  const bool fs = C > 0 && out->fixed_semantic_map && sample_box && box_sem;
  const bool fi = K > 0 && out->fixed_instance_map && sample_box && box_inst;
  if (fs || fi)
    return launch_fixed_maps(out->weights, sample_box, box_sem, box_inst, R, N, C, K, B,
                             fs ? out->fixed_semantic_map : nullptr, fi ? out->fixed_instance_map : nullptr,
                             (cudaStream_t)stream);
  return PNR_OK;
}


// Build (once per weight load) the second program `aux` runs and upload its packed weights / constants.
// ------------------------------------------------------------------------------------------------ device-side updates
// A training loop changes the weights every step; re-running the host builder (three programs, ~30 ms each) and
// copying the parameters to the host and back would cost more than the step itself.  The structure of a program does
// not depend on the values, so the builder is run ONCE in index mode and the packed streams / constant tables are
// refreshed on the device from the caller's DEVICE tensors: V <- tensors, fold_kernel (the feature_linear fold, same
// double-precision sums in the same order as the host), pack / constants kernels per program.  Bit-identical to a
// fresh pnr_load_weights of the same values (tests/test_gpu_backward.py::test_update_weights_equals_fresh_load).
namespace {

__device__ __forceinline__ uint16_t dev_f2bf(float x) {   // = the host f2bf (RNE on the bit pattern)
  const uint32_t u = __float_as_uint(x);
  return (uint16_t)((u + 0x7FFFu + ((u >> 16) & 1u)) >> 16);
}

__global__ void pack_from_plan_kernel(const float* __restrict__ V, const int32_t* __restrict__ idx,
                                      const uint8_t* __restrict__ part, size_t n, int fmt, uint16_t* __restrict__ out,
                                      uint32_t* status) {
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const float w = idx[p] < 0 ? 0.f : V[idx[p]];
  uint16_t v;
  if (fmt == 1) {
    const uint16_t h = dev_f2bf(w);
    v = part[p] == 0 ? h : dev_f2bf(w - __uint_as_float((uint32_t)h << 16));
  } else {
    if (!(w >= -65504.f && w <= 65504.f) && status != nullptr) atomicOr(status, 2u);   // weight outside the fp16 range
    const __half h = __float2half_rn(w);
    const __half r = part[p] == 0 ? h : __float2half_rn(w - __half2float(h));
    v = __half_as_ushort(r);
  }
  out[p] = v;
}

__global__ void consts_from_plan_kernel(const float* __restrict__ V, const int32_t* __restrict__ idx, size_t n,
                                        float* __restrict__ out) {
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) out[p] = idx[p] < 0 ? 0.f : V[idx[p]];
}

// W' = W_view[:, :W] W_feat,  b' = W_view[:, :W] b_feat + b_view (double sums, j ascending, as on the host); the
// gamma(d) columns of W_view are copied.  One thread per element of [W2, W + Ed] (+ one column for the bias).
__global__ void fold_kernel(float* V, int64_t off_view_w, int64_t off_view_b, int64_t off_feat_w, int64_t off_feat_b,
                            int W, int W2, int Ed, int64_t off_fold, int64_t off_fold_b) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int cols = W + Ed + 1;
  if (i >= W2 * cols) return;
  const int n = i / cols, k = i % cols;
  const float* vrow = V + off_view_w + (int64_t)n * (W + Ed);
  if (k < W) {
    double acc = 0.0;
    for (int j = 0; j < W; ++j) acc += (double)vrow[j] * (double)V[off_feat_w + (int64_t)j * W + k];
    V[off_fold + (int64_t)n * (W + Ed) + k] = (float)acc;
  } else if (k < W + Ed) {
    V[off_fold + (int64_t)n * (W + Ed) + k] = vrow[k];
  } else {
    double acc = (double)V[off_view_b + n];
    for (int j = 0; j < W; ++j) acc += (double)vrow[j] * (double)V[off_feat_b + j];
    V[off_fold_b + n] = (float)acc;
  }
}

}  // namespace

// which: 0 = the forward program, 1 = backward, 2 = trunk forward.  Index-mode build -> plan on the device.
static int ensure_plan(pnr_ctx* ctx, pnr_ctx::Plan& plan, int which, size_t expect_w, size_t expect_c) {
  if (plan.ready) return PNR_OK;
  const int n = (int)ctx->v_off.size();
  if (ctx->v_total + ctx->v_derived + 1 >= (int64_t)1 << 24)
    return set_error(PNR_ERR_UNSUPPORTED, "pnr_update_weights: %lld values do not index exactly in fp32", (long long)ctx->v_total);
  std::vector<std::vector<float>> it(n);
  std::vector<const float*> tp(n);
  for (int i = 0; i < n; ++i) {
    const int64_t cnt = ctx->all_shapes[2 * i] * ctx->all_shapes[2 * i + 1];
    it[i].resize((size_t)cnt);
    for (int64_t j = 0; j < cnt; ++j) it[i][(size_t)j] = (float)(ctx->v_off[i] + j + 1);
    tp[i] = it[i].data();
  }
  Builder bld(ctx->passes, ctx->fmt);
  bld.index_mode = true;
  bld.derived_base = ctx->v_total;
  const int rc = which == 0 ? build_program(ctx->cfg, tp.data(), ctx->all_shapes.data(), n, bld)
                            : build_backward_program(ctx->cfg, tp.data(), ctx->all_shapes.data(), n, bld, which == 2);
  if (rc != PNR_OK) return rc;
  if (bld.wbuf.size() != expect_w || bld.consts.size() != expect_c)
    return set_error(PNR_ERR_STATE, "pnr_update_weights: plan / program size mismatch (%zu/%zu vs %zu/%zu)", bld.wbuf.size(),
                     bld.consts.size(), expect_w, expect_c);
  std::vector<int32_t> widx(bld.wsrc.size()), cidx(bld.consts.size());
  for (size_t i = 0; i < widx.size(); ++i) widx[i] = (int32_t)bld.wsrc[i] - 1;
  for (size_t i = 0; i < cidx.size(); ++i) cidx[i] = (int32_t)bld.consts[i] - 1;
  plan.n_w = widx.size();
  plan.n_c = cidx.size();
  PNR_CUDA(cudaMalloc(&plan.d_widx, plan.n_w * 4));
  PNR_CUDA(cudaMalloc(&plan.d_wpart, plan.n_w));
  PNR_CUDA(cudaMalloc(&plan.d_cidx, plan.n_c * 4));
  PNR_CUDA(cudaMemcpy(plan.d_widx, widx.data(), plan.n_w * 4, cudaMemcpyHostToDevice));
  PNR_CUDA(cudaMemcpy(plan.d_wpart, bld.wpart.data(), plan.n_w, cudaMemcpyHostToDevice));
  PNR_CUDA(cudaMemcpy(plan.d_cidx, cidx.data(), plan.n_c * 4, cudaMemcpyHostToDevice));
  plan.ready = true;
  return PNR_OK;
}

static int repack_from_V(pnr_ctx* ctx, pnr_ctx::Plan& plan, uint8_t* d_wpacked, float* d_consts, cudaStream_t st) {
  pack_from_plan_kernel<<<(unsigned)((plan.n_w + 255) / 256), 256, 0, st>>>(ctx->d_V, plan.d_widx, plan.d_wpart, plan.n_w, ctx->fmt,
                                                                              reinterpret_cast<uint16_t*>(d_wpacked), ctx->d_status);
  PNR_LAUNCH_CHECK("pack_from_plan_kernel");
  consts_from_plan_kernel<<<(unsigned)((plan.n_c + 255) / 256), 256, 0, st>>>(ctx->d_V, plan.d_cidx, plan.n_c, d_consts);
  PNR_LAUNCH_CHECK("consts_from_plan_kernel");
  return PNR_OK;
}

extern "C" int pnr_update_weights(pnr_ctx* ctx, const float* const* device_tensors, int32_t n, void* stream) {
  PNR_CHECK_ARG(ctx && device_tensors, "pnr_update_weights: null pointer");
  if (!ctx->loaded) return set_error(PNR_ERR_STATE, "pnr_update_weights: pnr_load_weights has not been called (it fixes the shapes)");
  PNR_CHECK_ARG(n == (int)ctx->v_off.size(), "pnr_update_weights: got %d tensors, pnr_load_weights had %d", n, (int)ctx->v_off.size());
  for (int i = 0; i < n; ++i) PNR_CHECK_ARG(device_tensors[i], "pnr_update_weights: tensor %d is null", i);
  DeviceGuard guard(ctx->cfg.device);
  cudaStream_t st = (cudaStream_t)stream;
  if (!ctx->d_V) PNR_CUDA(cudaMalloc(&ctx->d_V, (size_t)(ctx->v_total + ctx->v_derived) * 4));
  if (const int rc = ensure_plan(ctx, ctx->plan_main, 0, ctx->wpacked_bytes / 2, (size_t)ctx->launch.prog.n_consts)) return rc;
  for (int i = 0; i < n; ++i)
    PNR_CUDA(cudaMemcpyAsync(ctx->d_V + ctx->v_off[i], device_tensors[i],
                             (size_t)(ctx->all_shapes[2 * i] * ctx->all_shapes[2 * i + 1]) * 4, cudaMemcpyDeviceToDevice, st));
  const int D = ctx->cfg.D, W = ctx->cfg.W, W2 = W / 2, Ed = 3 + 6 * ctx->cfg.view_res;
  // tensor order (pnr_load_weights): trunk (w, b) x D, alpha, feature, view, rgb, heads
  const int64_t off_feat_w = ctx->v_off[2 * D + 2], off_feat_b = ctx->v_off[2 * D + 3];
  const int64_t off_view_w = ctx->v_off[2 * D + 4], off_view_b = ctx->v_off[2 * D + 5];
  const int64_t off_fold = ctx->v_total, off_fold_b = ctx->v_total + (int64_t)W2 * (W + Ed);
  fold_kernel<<<(W2 * (W + Ed + 1) + 127) / 128, 128, 0, st>>>(ctx->d_V, off_view_w, off_view_b, off_feat_w, off_feat_b, W, W2, Ed,
                                                               off_fold, off_fold_b);
  PNR_LAUNCH_CHECK("fold_kernel");
  if (const int rc = repack_from_V(ctx, ctx->plan_main, ctx->d_wpacked, ctx->d_consts, st)) return rc;
  ctx->device_weights = true;
  // the aux programs that exist follow; the others are packed from V when they are first built
  if (ctx->bwd.ready) {
    if (const int rc = ensure_plan(ctx, ctx->plan_bwd, 1, ctx->bwd.n_w, ctx->bwd.n_c)) return rc;
    if (const int rc = repack_from_V(ctx, ctx->plan_bwd, ctx->bwd.d_wpacked, ctx->bwd.d_consts, st)) return rc;
  }
  if (ctx->trunk_fwd.ready) {
    if (const int rc = ensure_plan(ctx, ctx->plan_tf, 2, ctx->trunk_fwd.n_w, ctx->trunk_fwd.n_c)) return rc;
    if (const int rc = repack_from_V(ctx, ctx->plan_tf, ctx->trunk_fwd.d_wpacked, ctx->trunk_fwd.d_consts, st)) return rc;
  }
  return PNR_OK;
}

static int ensure_aux(pnr_ctx* ctx, pnr_ctx::Aux& aux, bool forward_only, cudaStream_t st) {
  if (aux.ready) return PNR_OK;
  Builder bld(ctx->passes, ctx->fmt);
  std::vector<const float*> tp;
  for (const auto& v : ctx->host_trunk) tp.push_back(v.data());
  if (const int rc = build_backward_program(ctx->cfg, tp.data(), ctx->host_trunk_shapes.data(), (int32_t)tp.size(), bld,
                                            forward_only))
    return rc;
  cudaFree(aux.d_wpacked); cudaFree(aux.d_consts);
  aux.d_wpacked = nullptr; aux.d_consts = nullptr;
  PNR_CUDA(cudaMalloc(&aux.d_wpacked, bld.wbuf.size() * 2));
  PNR_CUDA(cudaMalloc(&aux.d_consts, bld.consts.size() * 4));
  PNR_CUDA(cudaMemcpy(aux.d_wpacked, bld.wbuf.data(), bld.wbuf.size() * 2, cudaMemcpyHostToDevice));
  PNR_CUDA(cudaMemcpy(aux.d_consts, bld.consts.data(), bld.consts.size() * 4, cudaMemcpyHostToDevice));
  memset(&aux.launch.p, 0, sizeof(MlpParams));
  aux.launch.prog = bld.prog;
  aux.n_w = bld.wbuf.size();
  aux.n_c = bld.consts.size();
  aux.ready = true;
  if (ctx->device_weights) {   // the host copies are older than the weights in V: pack this program from V
    pnr_ctx::Plan& plan = forward_only ? ctx->plan_tf : ctx->plan_bwd;
    if (const int rc = ensure_plan(ctx, plan, forward_only ? 2 : 1, aux.n_w, aux.n_c)) return rc;
    return repack_from_V(ctx, plan, aux.d_wpacked, aux.d_consts, st);
  }
  return PNR_OK;
}

static int aux_launch(pnr_ctx* ctx, pnr_ctx::Aux& aux, const char* what, const float* pts, const float* rays, const float* z,
                      int64_t R, int32_t N, const float* grad_h, float grad_scale, float* out, int32_t ld_out, float* stash,
                      uint32_t* stash_absmax, void* stream) {
  if (!ctx->loaded) return set_error(PNR_ERR_STATE, "%s: pnr_load_weights has not been called", what);
  PNR_CHECK_ARG(R > 0 && N >= 1, "%s: bad sizes R=%lld N=%d", what, (long long)R, N);
  PNR_CHECK_ARG(pts || (rays && z), "%s: need pts or (rays, z)", what);
  const int64_t S = R * (int64_t)N;
  PNR_CHECK_ARG((S + kTileM - 1) / kTileM < (int64_t)1 << 31, "%s: too many samples", what);
  PNR_CHECK_ARG(stash == nullptr || (reinterpret_cast<uintptr_t>(stash) & 15) == 0, "%s: stash must be 16-byte aligned", what);
  DeviceGuard guard(ctx->cfg.device);
  if (const int rc = ensure_aux(ctx, aux, grad_h == nullptr, (cudaStream_t)stream)) return rc;
  MlpParams& p = aux.launch.p;
  p.wpacked = aux.d_wpacked; p.consts = aux.d_consts;
  p.pts = pts; p.viewdirs = nullptr; p.rays = rays; p.z = z;
  p.S = S; p.N = N; p.CH = ld_out; p.raw = out;
  p.num_tiles = (int32_t)((S + kTileM - 1) / kTileM);
  p.status = ctx->d_status;
  p.dbg = ctx->dbg_timeline;
  p.grad_in = grad_h;
  p.stash = stash;
  p.stash_absmax = stash_absmax;
  if (stash_absmax != nullptr)
    PNR_CUDA(cudaMemsetAsync(stash_absmax, 0, sizeof(uint32_t) * (size_t)(2 * ctx->cfg.D - 1), (cudaStream_t)stream));
  p.grad_scale = grad_scale;
  p.grad_unscale = 1.0f / grad_scale;
  return launch_mlp(aux.launch, ctx->passes, ctx->fmt, kMlpBackward, (cudaStream_t)stream);
}

// dL/d(embedded xyz) through the trunk (the tensor-core part of the MLP backward): the forward
// trunk is recomputed per tile (sign patterns stay in shared memory), then the layers run in reverse on the same tiles
// with the transposed weight stream.  grad_h = dL/dh of the trunk output [R*N, W]; grad_emb [R*N, ld_emb], the first
// 3 + 6*xyz_res columns of a row are the gradient (ld_emb = 64 with a 16-byte aligned base: vector stores).
// stash (nullable) [2D-1, R*N, W] fp32 receives every A operand on the way: H_i (i < D-1) in slot i, the
// pre-activation gradient dZ_j in slot 2D-2-j - the operands of the weight-gradient GEMMs dW_j = dZ_j^T H_{j-1}.
// grad_scale: a power of two the incoming gradient is multiplied by on load (every gradient leaving the kernel is
// divided by it again): gradients of a mean-reduced loss are ~1e-6, far below the normal range of the fp16 operand
// parts; scale so that max |grad_h| * grad_scale is a few hundred (the pass is linear, the scaling exact).
extern "C" int pnr_mlp_backward_trunk(pnr_ctx* ctx, const float* pts, const float* rays, const float* z, int64_t R,
                                      int32_t N, const float* grad_h, float grad_scale, float* grad_emb, int32_t ld_emb,
                                      float* stash, uint32_t* stash_absmax, void* stream) {
  if (R == 0) return PNR_OK;
  PNR_CHECK_ARG(ctx && grad_h && grad_emb, "pnr_mlp_backward_trunk: null pointer");
  PNR_CHECK_ARG(ld_emb >= 3 + 6 * ctx->cfg.xyz_res, "pnr_mlp_backward_trunk: ld_emb=%d < %d columns", ld_emb,
                3 + 6 * ctx->cfg.xyz_res);
  int gexp = 0;
  PNR_CHECK_ARG(grad_scale > 0.f && grad_scale < 1.0e30f && grad_scale > 1.0e-30f && frexpf(grad_scale, &gexp) == 0.5f,
                "pnr_mlp_backward_trunk: grad_scale=%g must be a positive power of two", (double)grad_scale);
  PNR_CHECK_ARG(stash_absmax == nullptr || stash != nullptr, "pnr_mlp_backward_trunk: stash_absmax without a stash");
  return aux_launch(ctx, ctx->bwd, "pnr_mlp_backward_trunk", pts, rays, z, R, N, grad_h, grad_scale, grad_emb, ld_emb, stash,
                    stash_absmax, stream);
}

// The trunk's output activations h [R*N, W] (what alpha_linear, feature_linear and the heads read): the forward
// trunk on the same tiles, last layer written out.  The training forward needs it once per step: everything
// after the trunk is differentiated by the caller (torch), everything before it by pnr_mlp_backward_trunk.
extern "C" int pnr_mlp_trunk_forward(pnr_ctx* ctx, const float* pts, const float* rays, const float* z, int64_t R,
                                     int32_t N, float* h_out, void* stream) {
  if (R == 0) return PNR_OK;
  PNR_CHECK_ARG(ctx && h_out, "pnr_mlp_trunk_forward: null pointer");
  PNR_CHECK_ARG((reinterpret_cast<uintptr_t>(h_out) & 15) == 0, "pnr_mlp_trunk_forward: h_out must be 16-byte aligned");
  return aux_launch(ctx, ctx->trunk_fwd, "pnr_mlp_trunk_forward", pts, rays, z, R, N, nullptr, 1.0f, h_out, ctx->cfg.W, nullptr, nullptr, stream);
}

namespace pnr {
size_t workspace_bytes_for(int64_t R, int N, int Ni, int CH);   // render.cu
int ctx_channels(const pnr_ctx* ctx) { return 4 + ctx->cfg.num_classes + ctx->cfg.num_instances; }
int ctx_classes(const pnr_ctx* ctx, int* C, int* K) {
  *C = ctx->cfg.num_classes;
  *K = ctx->cfg.num_instances;
  return PNR_OK;
}
}  // namespace pnr

extern "C" size_t pnr_workspace_bytes(const pnr_ctx* ctx, int64_t R, int32_t N, int32_t Ni) {
  if (!ctx || R <= 0 || N < 1) return 0;
  return workspace_bytes_for(R, N, Ni > 0 ? Ni : 0, ctx_channels(ctx));
}
