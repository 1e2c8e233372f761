// jamie_b200 engine: device state of one JAMIE model (both modalities' encoder/decoder MLPs, Adam moments, BatchNorm
// statistics, resident datasets) and the C ABI declared in include/jamie_b200.h.
//
// Memory layout:
//   theta / grad / adam_m / adam_v (fp32) : one flat buffer each with the SAME padded layout. Every tensor starts on a
//       128-byte boundary and 2-D weights use a row pitch rounded up to 8 elements so that TMA can address them directly
//       in fp32 and in fp16 (tensor maps need 16-byte aligned bases and pitches). Padding stays zero forever (zero
//       gradient -> zero Adam update), so clip-norm and Adam run over the whole buffer with 128-bit accesses.
//       fc_mus.i / fc_vars.i are stored as ONE [2L, D] matrix per modality (mu rows, then logvar rows): one heads GEMM.
//   theta_hi / theta_lo (fp16)            : the GEMM operand planes of theta (hgemm.cuh), same element offsets,
//       rewritten by every Adam step.
//   activations: [B, pitch] row-major per layer and modality, pitch = width rounded up to 8 elements; GEMM outputs are
//       fp32 (with room for split-K partial sums), GEMM inputs are fp16 hi / lo planes.
// One training step = the phases of ONE persistent cooperative kernel (stepk.cuh); every step-varying scalar (plan row,
// KL anneal, Adam bias corrections, Philox stream) derives from a device-side control block, so jb_train_steps(n) is a
// single launch for n optimizer steps.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/jamie_b200.h"
#include "gemm_tf32.cuh"
#include "gemm_persistent.cuh"
#include "kernels.cuh"
#include "hgemm.cuh"
#include "stepk.cuh"
#include "metrics.cuh"

namespace {

thread_local std::string g_err;
int fail(const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return 1;
}
#define CU(x)                                                                                      \
  do {                                                                                             \
    cudaError_t e_ = (x);                                                                          \
    if (e_ != cudaSuccess) return fail("%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

inline int r4(int x) { return (x + 3) & ~3; }
inline int r8(int x) { return (x + 7) & ~7; }
inline long long r32(long long x) { return (x + 31) & ~31LL; }

struct Seg {  // one tensor of the padded flat layout
  long long off = 0;
  int rows = 0, cols = 0, ld = 0;
  long long span() const { return static_cast<long long>(rows) * ld; }
};
struct PackMap {  // packed (reference order) tensor -> location in the padded flat buffer
  long long dst_off;
  int rows, cols, ld;
};

struct ModSegs {
  Seg W1, b1, g1, be1, W2, b2, g2, be2;  // encoder: Linear(D,2D) BN(2D) | Linear(2D,D) BN(D)
  Seg Wmv, bmv;                          // heads: [2L, D] (fc_mus rows, fc_vars rows), [2L]
  Seg W3, b3, g3, be3, W4, b4, g4, be4, W5, b5;  // decoder: Linear(L,D) BN(D) | Linear(D,2D) BN(2D) | Linear(2D,D)
};

using jb::HPlanes;
struct ModActs {  // activations and gradients of one modality (device pointers; pitches in elements)
  float* x; HPlanes xp;
  float* x_stage[2];                                     // host-batch steps: two H2D landing buffers for x
  HPlanes h1, h2, cp, g1, g2;                            // forward GEMM operands
  HPlanes dxhat, dy4, dy3, dmp, dy2, dy1;                // backward GEMM operands
  float* dmulv;
  float *z, *c, *S, *g, *eps, *inj_eps, *den, *rs;
  float *bn_mean[4], *bn_inv[4], *bn_var[4];  // enc1, enc2, dec1, dec2
  unsigned char* inj_mask[4];
  float* rec_part;
  // GEMM outputs (fp32, split-K partial sums; carved per table build)
  jb::Parts y1, y2, mulv, y3, y4, xhat, dg2, dg1, dc, dh2, dh1;
  int ldD, ld2D;
};

}  // namespace

struct HostPin {
  long long cursor_val[2];   // {0, 1}: plan row of a slot, copied into the control block before the step
  int slot_val[2];           // {0, 1}
  float kl[2];
  float losses[2][8];
  int idx[1];                // [slot][modality][Bmax]
};

struct jb_engine {
  jb_config cfg{};
  int D[2]{}, L = 0, LP = 0, ldmv = 0, Bmax = 0;
  // flat parameter layout
  ModSegs ms[2];
  Seg sigma;
  long long n_flat = 0;  // padded float count (multiple of 32)
  std::vector<PackMap> packmap;
  long long n_packed = 0;
  float *theta = nullptr, *grad = nullptr, *adam_m = nullptr, *adam_v = nullptr, *theta_eval = nullptr;
  float* xg[8]{}; unsigned int* xf[8]{}; float* xmc = nullptr; int xrank = 0, xworld = 0; long long x_adam0 = 0;   // in-kernel exchange (jb_set_exchange)
  float* grad_own = nullptr;        // the engine's own gradient buffer while a caller-owned one is in use (jb_set_grad_buffer)
  __half* theta_eval_h = nullptr;   // fp16 copy of the folded inference weights (operands of the fp16 layers of the chain)
  bool eval_f16 = true;             // JB_EVAL_F16=0: the whole chain in TF32 with fp32 activations
  __half *theta_hi = nullptr, *theta_lo = nullptr;   // fp16 operand planes of theta (same element offsets)
  float* state_slab = nullptr;   // one allocation: theta | adam_m | adam_v | theta_hi, theta_lo
  size_t slab_bytes = 0;
  // BatchNorm running statistics: 8 layers in packed order (enc0.1, enc0.5, enc1.1, enc1.5, dec0.1, dec0.5, dec1.1, dec1.5)
  float* bn_run = nullptr;
  long long bn_off[8]{};
  int bn_w[8]{};
  long long n_bn = 0;
  long long nbt[8]{};  // num_batches_tracked (host side; every training step increments all 8)
  // datasets and priors
  float* data[2]{};
  long long data_n[2]{}, data_ld[2]{};
  float *p_diag = nullptr, *p_dense = nullptr, *f_dense = nullptr;
  long long pn0 = 0, pn1 = 0, p_diag_n = 0;
  // plan
  int* plan_idx[2]{};
  float* plan_kl = nullptr;
  float* out_loss = nullptr;
  int plan_cap = 0, plan_steps = 0, plan_B = 0;
  jb::Ctl* ctl = nullptr;
  double* norm_part = nullptr;
  float *norm_tile = nullptr, *norm_small = nullptr;   // fused clip norm (stepk.cuh)
  int norm_fuse = 1;                                   // JB_NORM_FUSE=0: always sweep the gradient buffer
  int norm_tile_cap = 0;
  unsigned int* bar = nullptr;          // grid barrier counters of k_step: [0] and [32], used by alternate launches
  int bar_parity = 0;
  unsigned long long* d_ts = nullptr;   // phase timestamps (profiling)
  // workspaces
  char* arena = nullptr;
  size_t arena_bytes = 0;
  char* parts_arena = nullptr;          // GEMM outputs with their split-K partials (depends on the batch size)
  ModActs act[2]{};
  float *corr = nullptr, *corr_t = nullptr, *fblk = nullptr, *fblk_t = nullptr;
  float *lat_r = nullptr, *rowpart = nullptr, *lat_coef = nullptr;
  int2* corr_hint = nullptr;
  int cosine = 0;                   // jb_set_dist_method
  float *cmax_part = nullptr, *dmax_part = nullptr, *dyn = nullptr;   // dynamic operand scales (stepk.cuh)
  // step tables at batch size step_B
  std::vector<jb::HgProblem> h_probs;
  jb::StepParams* h_prm = nullptr;   // the step kernel's parameter block (host copy; passed by value at every launch)
  jb::StepCtx& h_ctx_ref() { return h_prm->cx; }
  int step_B = 0;
  int grid = 132;            // CTAs of k_step: HG_CLUSTER x the co-resident clusters (33 x 4 on B200; set in jb_create)
  int fuse_enabled = 1;      // JB_FUSE=0: BatchNorm / reconstruction / reparameterisation as separate phases (the B > 512 path)
  int fused = 0;             // the current step tables use the cluster-fused tails
  int merge_latent = 1;      // JB_MERGE_LATENT=0: LATLOSS / LATFIN as separate phases even without F
  int fuse_ks = 1;           // JB_FUSE_KS=0: HEADS / DG3 as one CTA per M tile instead of K split over the cluster
  int accumulate = 0, accumulate_dev = 0;
  int wgrad_mode = jb::HG_MEDIUM;
  int fwd_mode = -1;                // forward / dgrad GEMM mode: -1 = by K (see build_step), else forced (JB_FWD_MODE=precise | medium)
  int wgrad_bn = 256;
  int max_ksplit = 8;
  float gs = 1.f;
  // Host-batch steps (data resident on the host): two slots so that the copies of batch k + 1 run while step k computes.
  struct HostPin* h_pin = nullptr;   // pinned: constants, per-slot kl / losses / indices
  cudaStream_t h2d_stream = nullptr;
  cudaEvent_t ev_h2d[2]{}, ev_slot_free[2]{}, ev_loss[2]{};
  bool slot_used[2]{};
  int hb_slot = 0, hb_oldest = 0, hb_outstanding = 0;
  int precision_fast = 0;    // JB_PRECISION=f16: single-pass fp16 everywhere (no parity claim)
  int num_sms = 148;
  // eval
  bool eval_dirty = true;
  int eval_chunk = 8192;     // rows per pass of the folded chain (JB_EVAL_CHUNK); activations of a chunk stay in L2
  int eval_bn256_min = 256;  // layers at least this wide use 256-column tiles (JB_EVAL_BN256_MIN)
  bool eval_persist = true;  // inference chain on the persistent GEMM (JB_EVAL_PERSIST=0: one tile per CTA)
  float *ev_a = nullptr, *ev_b = nullptr, *ev_in = nullptr, *ev_out = nullptr;
  jb::GemmProblem* d_ev_probs = nullptr;
  int ev_probs_cap = 0;
  cudaStream_t ev_stream[2]{};
  cudaEvent_t ev_done[2]{}, ev_free[2]{};
  std::vector<float> prof_gemm;     // last jb_profile_step: per GEMM phase 6 role stamps of CTA 0 (us after the phase began)
  std::vector<float> prof_detail;   // last jb_profile_step: per phase {total, longest CTA work, mean CTA work, barrier tail} us
  long long launches = 0;
};

namespace {

using jb::GemmProblem;

// ------------------------------------------------------------------------------------------- layout
void add_seg(jb_engine* e, Seg& s, int rows, int cols) {
  s.rows = rows; s.cols = cols; s.ld = rows > 1 ? r8(cols) : cols;
  s.off = e->n_flat;
  e->n_flat = r32(e->n_flat + (rows > 1 ? s.span() : cols));
}
void build_layout(jb_engine* e) {
  const int L = e->L;
  e->n_flat = 0;
  add_seg(e, e->sigma, 1, 2);
  for (int i = 0; i < 2; ++i) {
    const int D = e->D[i];
    ModSegs& m = e->ms[i];
    add_seg(e, m.W1, 2 * D, D); add_seg(e, m.b1, 1, 2 * D); add_seg(e, m.g1, 1, 2 * D); add_seg(e, m.be1, 1, 2 * D);
    add_seg(e, m.W2, D, 2 * D); add_seg(e, m.b2, 1, D); add_seg(e, m.g2, 1, D); add_seg(e, m.be2, 1, D);
  }
  for (int i = 0; i < 2; ++i) {
    const int D = e->D[i];
    ModSegs& m = e->ms[i];
    add_seg(e, m.Wmv, 2 * L, D); add_seg(e, m.bmv, 1, 2 * L);
    add_seg(e, m.W3, D, L); add_seg(e, m.b3, 1, D); add_seg(e, m.g3, 1, D); add_seg(e, m.be3, 1, D);
    add_seg(e, m.W4, 2 * D, D); add_seg(e, m.b4, 1, 2 * D); add_seg(e, m.g4, 1, 2 * D); add_seg(e, m.be4, 1, 2 * D);
    add_seg(e, m.W5, D, 2 * D); add_seg(e, m.b5, 1, D);
  }
  // packed (reference named_parameters) order -> flat locations
  auto pm = [&](const Seg& s, int row0, int rows, int cols) {
    PackMap p;
    p.dst_off = s.off + static_cast<long long>(row0) * (s.rows > 1 ? s.ld : 1);
    p.rows = rows; p.cols = cols; p.ld = s.rows > 1 ? s.ld : cols;
    e->packmap.push_back(p);
    e->n_packed += static_cast<long long>(rows) * cols;
  };
  auto whole = [&](const Seg& s) { pm(s, 0, s.rows, s.cols); };
  e->packmap.clear(); e->n_packed = 0;
  whole(e->sigma);
  for (int i = 0; i < 2; ++i) {
    ModSegs& m = e->ms[i];
    whole(m.W1); whole(m.b1); whole(m.g1); whole(m.be1); whole(m.W2); whole(m.b2); whole(m.g2); whole(m.be2);
  }
  for (int i = 0; i < 2; ++i) { pm(e->ms[i].Wmv, 0, L, e->D[i]); pm(e->ms[i].bmv, 0, 1, L); }           // fc_mus
  for (int i = 0; i < 2; ++i) { pm(e->ms[i].Wmv, L, L, e->D[i]); PackMap p; p.dst_off = e->ms[i].bmv.off + L; p.rows = 1; p.cols = L; p.ld = L; e->packmap.push_back(p); e->n_packed += L; }  // fc_vars
  for (int i = 0; i < 2; ++i) {
    ModSegs& m = e->ms[i];
    whole(m.W3); whole(m.b3); whole(m.g3); whole(m.be3); whole(m.W4); whole(m.b4); whole(m.g4); whole(m.be4);
    whole(m.W5); whole(m.b5);
  }
  // BatchNorm running stats
  const int widths[8] = {2 * e->D[0], e->D[0], 2 * e->D[1], e->D[1], e->D[0], 2 * e->D[0], e->D[1], 2 * e->D[1]};
  long long off = 0;
  for (int k = 0; k < 8; ++k) { e->bn_w[k] = widths[k]; e->bn_off[k] = off; off += 2LL * widths[k]; }
  e->n_bn = off;
}

int copy_packed(jb_engine* e, float* flat, float* packed_host, bool to_device) {
  // packed host <-> padded device, tensor by tensor (2-D copies honour the padded pitch)
  long long src = 0;
  for (const PackMap& p : e->packmap) {
    if (to_device)
      CU(cudaMemcpy2D(flat + p.dst_off, static_cast<size_t>(p.ld) * 4, packed_host + src, static_cast<size_t>(p.cols) * 4,
                      static_cast<size_t>(p.cols) * 4, p.rows, cudaMemcpyHostToDevice));
    else
      CU(cudaMemcpy2D(packed_host + src, static_cast<size_t>(p.cols) * 4, flat + p.dst_off, static_cast<size_t>(p.ld) * 4,
                      static_cast<size_t>(p.cols) * 4, p.rows, cudaMemcpyDeviceToHost));
    src += static_cast<long long>(p.rows) * p.cols;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------- arena
struct Carver {
  char* base; size_t off = 0; bool dry;
  explicit Carver(char* b) : base(b), dry(b == nullptr) {}
  template <class T> T* take(size_t n) {
    off = (off + 255) & ~size_t(255);
    T* p = dry ? nullptr : reinterpret_cast<T*>(base + off);
    off += n * sizeof(T);
    return p;
  }
};
void carve(jb_engine* e, Carver& c) {
  const size_t B = e->Bmax;
  for (int i = 0; i < 2; ++i) {
    ModActs& a = e->act[i];
    const int D = e->D[i];
    a.ldD = r8(D); a.ld2D = r8(2 * D);
    auto planes = [&](HPlanes& p, size_t n) { p.hi = c.take<__half>(n); p.lo = c.take<__half>(n); };
    a.x = c.take<float>(B * a.ldD); planes(a.xp, B * a.ldD);
    a.x_stage[0] = c.take<float>(B * a.ldD); a.x_stage[1] = c.take<float>(B * a.ldD);
    planes(a.h1, B * a.ld2D); planes(a.h2, B * a.ldD); planes(a.g1, B * a.ldD); planes(a.g2, B * a.ld2D);
    planes(a.dxhat, B * a.ldD); planes(a.dy4, B * a.ld2D); planes(a.dy3, B * a.ldD);
    a.dmulv = c.take<float>(B * e->ldmv); planes(a.dmp, B * e->ldmv); planes(a.dy2, B * a.ldD); planes(a.dy1, B * a.ld2D);
    a.z = c.take<float>(B * e->LP); a.c = c.take<float>(B * e->LP); planes(a.cp, B * e->LP); a.S = c.take<float>(B * e->LP);
    a.g = c.take<float>(B * e->LP); a.eps = c.take<float>(B * e->LP); a.inj_eps = c.take<float>(B * e->LP);
    a.den = c.take<float>(B); a.rs = c.take<float>(B);
    const int w[4] = {2 * D, D, D, 2 * D};
    for (int k = 0; k < 4; ++k) {
      a.bn_mean[k] = c.take<float>(w[k]); a.bn_inv[k] = c.take<float>(w[k]); a.bn_var[k] = c.take<float>(w[k]);
      a.inj_mask[k] = c.take<unsigned char>(B * w[k]);
    }
    a.rec_part = c.take<float>(D);   // one partial per slab item of the reconstruction phase (at most D items)
  }
  e->corr = c.take<float>(B * B); e->corr_t = c.take<float>(B * B);
  e->fblk = c.take<float>(B * B); e->fblk_t = c.take<float>(B * B);
  e->lat_r = c.take<float>(B * e->LP); e->rowpart = c.take<float>(2 * B * 8); e->lat_coef = c.take<float>(2 * B * 4); e->corr_hint = c.take<int2>(2 * B);
  e->cmax_part = c.take<float>(jb::SK_MAX_CTAS); e->dmax_part = c.take<float>(jb::SK_MAX_CTAS); e->dyn = c.take<float>(8);
}

// ------------------------------------------------------------------------------------------- step tables
// Forward GEMMs and dgrads run the fp32-class 3-pass mode on the fp16 hi / lo planes (HG_PRECISE): pre-activation errors
// flip LeakyReLU' decisions and dX errors propagate down the chain. Weight gradients run three passes without the
// accumulator drains (HG_MEDIUM, ~1e-6): their rounding stays local to the gradient tensor, but a single pass (3e-4 at
// the headline shape, up to 1.2e-3 at small ones, measured in round 1) would eat the whole 1e-3 parity budget.
struct StageSpec {   // one problem of a GEMM phase before its split-K factor is known
  HPlanes A; int lda, a_mn; HPlanes Bm; int ldb, b_mn;
  jb::Parts* out;    // activation output (partials carved later) or null ...
  float* C;          // ... for weight gradients, which go straight into the gradient buffer
  int ldc, M, N, K, bn, mode, epi;
  const float* bias;
  float out_scale;
  int acc_dynamic;
  int dyn;           // index into StepCtx::dyn of an extra dynamic output factor, or -1
  int fuse = 0, fuse_arg = 0, fuse_ks = 0;
};

int build_step(jb_engine* e, int B) {
  const int L = e->L;
  float* T = e->theta;
  float* G = e->grad;
  auto W = [&](const Seg& s) { return HPlanes{e->theta_hi + s.off, e->theta_lo + s.off}; };
  auto bias = [&](const Seg& s) { return T + s.off; };
  // loss scale of the backward pass: d xhat = gs * 2 w (xhat - x) / (B D) is O(xhat - x)
  {
    const int Dm = e->D[0] > e->D[1] ? e->D[0] : e->D[1];
    float wmax = 1.f;
    for (int k = 0; k < 4; ++k) wmax = fmaxf(wmax, fabsf(e->cfg.loss_w[k]));
    const int ex = static_cast<int>(ceil(log2(static_cast<double>(B) * Dm))) - static_cast<int>(ceil(log2(static_cast<double>(wmax))));
    e->gs = ldexpf(1.f, ex < 0 ? 0 : (ex > 30 ? 30 : ex));
    if (const char* pv = getenv("JB_LOSS_SCALE_LOG2")) e->gs = ldexpf(1.f, atoi(pv));
  }
  const float inv_gs = 1.f / e->gs;
  // Forward / dgrad mode by reduction length: the tensor core adds products into its fp32 accumulator with truncation, a
  // bias that grows with the number of accumulation steps. Up to K = 1024 (64 steps) three undrained passes (HG_MEDIUM)
  // stay at 7e-6 of the fp32 oracle at the headline shape and save 9 us per step; longer reductions (2000-wide inputs
  // without PCA: measured 6e-4 undrained) keep the drained HG_PRECISE mode.
  auto fmode_for = [&](int K) {
    if (e->precision_fast) return static_cast<int>(jb::HG_SINGLE);
    if (e->fwd_mode >= 0) return e->fwd_mode;
    return static_cast<int>(K <= 1024 ? jb::HG_MEDIUM : jb::HG_PRECISE);
  };
  const int wmode = e->precision_fast ? jb::HG_SINGLE : e->wgrad_mode;
  std::vector<std::vector<StageSpec>> st(jb::SK_NUM_GEMM);
  // Cluster-fused tails need all rows of a column block in one cluster: B <= HG_CLUSTER M tiles.
  const bool fused = e->fuse_enabled && B <= jb::HG_CLUSTER * jb::HG_BM;
  e->fused = fused ? 1 : 0;
  const int nclusters = e->grid / jb::HG_CLUSTER;
  // N tile of the forward / dgrad stages: 64 columns, or 32 when the 64-wide column blocks of both modalities would leave
  // half of the clusters idle (fused) / for narrow outputs
  auto fbn = [&](int N) { return N <= 32 ? 32 : 64; };
  auto fbn2 = [&](int N0, int N1) {
    if (!fused) return 64;
    const int blocks64 = (N0 + 63) / 64 + (N1 + 63) / 64;
    return 2 * blocks64 <= nclusters ? 32 : 64;
  };
  for (int i = 0; i < 2; ++i) {
    ModActs& a = e->act[i]; ModSegs& m = e->ms[i]; const int D = e->D[i];
    auto fwd = [&](int stage, HPlanes X, int ldx, const Seg& w, const Seg& b, jb::Parts* out, int ldy, int n_out, int n_in) {
      StageSpec sp{X, ldx, 0, W(w), w.ld, 0, out, nullptr, ldy, B, n_out, n_in, fbn(n_out), fmode_for(n_in), jb::EPI_BIAS, bias(b), 1.f, 0,
                   stage == 3 ? 0 : -1};   // the c planes carry the dynamic scale s_c
      if (fused) {
        static const int bn_layer[6] = {0, 1, -1, 2, 3, -1};
        if (bn_layer[stage] >= 0) { sp.fuse = jb::FUSE_BN_FWD; sp.fuse_arg = bn_layer[stage] * 2 + i; }
        else if (stage == 5) { sp.fuse = jb::FUSE_REC; sp.fuse_arg = i; }
        else if (stage == 2 && 2 * L <= 64) { sp.fuse = jb::FUSE_HEADS; sp.fuse_arg = i; sp.fuse_ks = e->fuse_ks && n_in >= 8 * jb::HG_BK; }
        // narrow column blocks only pay where the main loop is long (K >= 256); short-K stages keep 64
        if (stage != 2 && n_out > 32 && n_in >= 256) sp.bn = stage == 0 || stage == 4 ? fbn2(2 * e->D[0], 2 * e->D[1]) : fbn2(e->D[0], e->D[1]);
      }
      st[stage].push_back(sp);
    };
    fwd(0, a.xp, a.ldD, m.W1, m.b1, &a.y1, a.ld2D, 2 * D, D);
    fwd(1, a.h1, a.ld2D, m.W2, m.b2, &a.y2, a.ldD, D, 2 * D);
    fwd(2, a.h2, a.ldD, m.Wmv, m.bmv, &a.mulv, e->ldmv, 2 * L, D);
    fwd(3, a.cp, e->LP, m.W3, m.b3, &a.y3, a.ldD, D, L);
    fwd(4, a.g1, a.ldD, m.W4, m.b4, &a.y4, a.ld2D, 2 * D, D);
    fwd(5, a.g2, a.ld2D, m.W5, m.b5, &a.xhat, a.ldD, D, 2 * D);
    // dgrad dX[B, N_in] = dY W   (A = dY planes K-major, B = W planes MN-major, K = N_out)
    auto dgrad = [&](int stage, HPlanes dY, int lddy, const Seg& s, jb::Parts* out, int lddx, int n_out, int n_in) {
      StageSpec sp{dY, lddy, 0, W(s), s.ld, 1, out, nullptr, lddx, B, n_in, n_out, fbn(n_in), fmode_for(n_out), jb::EPI_STORE, nullptr, 1.f, 0, -1};
      if (fused && stage == 8) {   // d c: the LATBC phase becomes the tail (no F, one column block)
        if (e->f_dense == nullptr && e->merge_latent && L <= 64) { sp.fuse = jb::FUSE_LATBC; sp.fuse_arg = i; sp.fuse_ks = e->fuse_ks && n_out >= 8 * jb::HG_BK; }
      } else if (fused) {   // the dgrad result feeds a BatchNorm backward: stage 6 -> dec2, 7 -> dec1, 9 -> enc2, 10 -> enc1
        const int k = stage == 6 ? 3 : (stage == 7 ? 2 : (stage == 9 ? 1 : 0));
        sp.fuse = jb::FUSE_BN_BWD; sp.fuse_arg = k * 2 + i;
        if (n_in > 32 && n_out >= 256) sp.bn = (k == 0 || k == 3) ? fbn2(2 * e->D[0], 2 * e->D[1]) : fbn2(e->D[0], e->D[1]);
      }
      st[stage].push_back(sp);
    };
    dgrad(6, a.dxhat, a.ldD, m.W5, &a.dg2, a.ld2D, D, 2 * D);
    dgrad(7, a.dy4, a.ld2D, m.W4, &a.dg1, a.ldD, 2 * D, D);
    dgrad(8, a.dy3, a.ldD, m.W3, &a.dc, e->LP, D, L);
    dgrad(9, a.dmp, e->ldmv, m.Wmv, &a.dh2, a.ldD, 2 * L, D);
    dgrad(10, a.dy2, a.ldD, m.W2, &a.dh1, a.ld2D, D, 2 * D);
  }
  // wgrad dW[N_out, N_in] = dY^T X / gs  (A = dY planes MN-major, B = X planes MN-major, K = batch); largest first
  // The two small weight gradients do not wait for the WGRAD phase: d W3 = d y3^T c runs beside the DG3 GEMM (8 + 8 work
  // items at the headline shape) and d Wmv = d[mu|logvar]^T h2 beside DGH, on CTAs those phases leave idle; WGRAD is then
  // exactly one work item per CTA (128 items of 128 x 256 on 132 CTAs).
  auto wgrad = [&](HPlanes dY, int lddy, HPlanes X, int ldx, const Seg& s, int n_out, int n_in, int dyn, int stage = 11) {
    int bn = n_in <= 32 ? 32 : (n_in <= 64 || stage != 11 ? 64 : (n_in >= 256 && e->wgrad_bn >= 256 ? 256 : 128));
    st[stage].push_back(StageSpec{dY, lddy, 1, X, ldx, 1, nullptr, G + s.off, s.ld, n_out, n_in, B, bn, wmode, jb::EPI_STORE, nullptr, inv_gs, 1, dyn});
  };
  for (int i = 0; i < 2; ++i) { ModActs& a = e->act[i]; ModSegs& m = e->ms[i]; const int D = e->D[i];
    // encoder-side gradients carry the dynamic scale s_b of d[mu | logvar] (dyn 1); dW3's B operand is the c planes (dyn 0)
    wgrad(a.dxhat, a.ldD, a.g2, a.ld2D, m.W5, D, 2 * D, -1); wgrad(a.dy4, a.ld2D, a.g1, a.ldD, m.W4, 2 * D, D, -1);
    wgrad(a.dy2, a.ldD, a.h1, a.ld2D, m.W2, D, 2 * D, 1); wgrad(a.dy1, a.ld2D, a.xp, a.ldD, m.W1, 2 * D, D, 1); }
  for (int i = 0; i < 2; ++i) { ModActs& a = e->act[i]; ModSegs& m = e->ms[i]; const int D = e->D[i];
    wgrad(a.dmp, e->ldmv, a.h2, a.ldD, m.Wmv, 2 * L, D, 1, 9); wgrad(a.dy3, a.ldD, a.cp, e->LP, m.W3, D, L, 0, 8); }
  // split-K factor per problem: fill the grid, at least two k-blocks per partial
  size_t parts_bytes = 0;
  std::vector<std::vector<int>> ks(jb::SK_NUM_GEMM);
  for (int g = 0; g < jb::SK_NUM_GEMM; ++g) {
    int tiles = 0;
    for (const StageSpec& s : st[g]) tiles += ((s.M + jb::HG_BM - 1) / jb::HG_BM) * ((s.N + s.bn - 1) / s.bn);
    for (const StageSpec& s : st[g]) {
      int k = 1;
      if (s.out != nullptr) {
        const int kb = (s.K + jb::HG_BK - 1) / jb::HG_BK;
        k = s.fuse ? (s.fuse_ks ? jb::HG_CLUSTER : 1) : e->grid / (tiles > 0 ? tiles : 1);
        if (k > kb / 2) k = kb / 2;
        if (k > e->max_ksplit) k = e->max_ksplit;
        if (k < 1) k = 1;
        parts_bytes = ((parts_bytes + 255) & ~size_t(255)) + static_cast<size_t>(k) * s.M * s.ldc * 4;
      }
      ks[g].push_back(k);
    }
  }
  if (e->parts_arena) { cudaFree(e->parts_arena); e->parts_arena = nullptr; }
  CU(cudaMalloc(&e->parts_arena, parts_bytes + 256));
  CU(cudaMemset(e->parts_arena, 0, parts_bytes + 256));
  Carver pc(e->parts_arena);
  e->h_probs.clear();
  if (!e->h_prm) e->h_prm = new jb::StepParams();
  jb::StepCtx& cx = e->h_prm->cx;
  cx = jb::StepCtx{};
  int n_norm_tile = 0;
  for (int g = 0; g < jb::SK_NUM_GEMM; ++g) {
    const int first = static_cast<int>(e->h_probs.size());
    if (st[g].size() > static_cast<size_t>(jb::HG_MAX_PROBS) || e->h_probs.size() + st[g].size() > static_cast<size_t>(jb::SK_MAX_PROBS))
      return fail("too many problems in a GEMM phase");
    for (size_t q = 0; q < st[g].size(); ++q) {
      const StageSpec& s = st[g][q];
      const int k = ks[g][q];
      float* C = s.C;
      long long pstride = 0;
      if (s.out != nullptr) {
        pstride = static_cast<long long>(s.M) * s.ldc;
        C = pc.take<float>(static_cast<size_t>(k) * pstride);
        *s.out = jb::Parts{C, pstride, k};
      }
      jb::HgProblem hp;
      int rc = jb::hg_problem_fill(&hp, s.A, s.lda, s.a_mn, s.Bm, s.ldb, s.b_mn, C, s.ldc, s.M, s.N, s.K, s.bn, s.mode, s.epi, s.bias, k,
                                   pstride, 0, s.out_scale, jb::LRELU);
      if (rc) return fail("hgemm problem fill failed (%d) for M%d N%d K%d lda%d ldb%d bn%d mode%d", rc, s.M, s.N, s.K, s.lda, s.ldb, s.bn, s.mode);
      if (hp.ksplit != k && s.out != nullptr) s.out->n = hp.ksplit;
      if (s.fuse && s.out != nullptr) s.out->n = 1;   // fused tails store the finished tensor (K parts are summed over DSMEM)
      if (s.fuse_ks && hp.ksplit != jb::HG_CLUSTER) return fail("cluster K split needs %d parts, got %d", jb::HG_CLUSTER, hp.ksplit);
      if (s.acc_dynamic) hp.acc_flag = &e->ctl->accum;
      if (s.dyn >= 0) hp.dyn_scale = e->dyn + s.dyn;
      hp.fuse = s.fuse; hp.fuse_arg = s.fuse_arg; hp.fuse_ks = s.fuse_ks;
      if (s.out == nullptr) {   // weight gradient: per-warp sums of squares of what the epilogue stored (offset now, base below)
        hp.norm_out = reinterpret_cast<float*>(static_cast<uintptr_t>(n_norm_tile) * 4 + 4);   // +4: distinguishes offset 0 from "none"
        n_norm_tile += hp.tiles_m * hp.tiles_n * jb::HG_NEPI;
      }
      if (s.fuse && hp.tiles_m > jb::HG_CLUSTER) return fail("fused stage with %d M tiles", hp.tiles_m);
      e->h_probs.push_back(hp);
    }
    cx.gph[g] = jb::hg_phase_finalize(e->h_probs.data(), first, static_cast<int>(st[g].size()));
  }
  if (n_norm_tile > e->norm_tile_cap) {
    if (e->norm_tile) cudaFree(e->norm_tile);
    e->norm_tile = nullptr; e->norm_tile_cap = 0;
    CU(cudaMalloc(&e->norm_tile, static_cast<size_t>(n_norm_tile) * 4));
    e->norm_tile_cap = n_norm_tile;
  }
  if (n_norm_tile > 0) CU(cudaMemset(e->norm_tile, 0, static_cast<size_t>(n_norm_tile) * 4));
  for (jb::HgProblem& hp : e->h_probs)
    if (hp.norm_out != nullptr) hp.norm_out = e->norm_tile + (reinterpret_cast<uintptr_t>(hp.norm_out) - 4) / 4;
  for (size_t q = 0; q < e->h_probs.size(); ++q) e->h_prm->probs[q] = e->h_probs[q];
  // ---- the rest of the step context
  cx.B = B; cx.L = L; cx.LP = e->LP; cx.ldmv = e->ldmv;
  for (int i = 0; i < 2; ++i) {
    ModActs& a = e->act[i]; ModSegs& m = e->ms[i];
    jb::ModCtx& M = cx.m[i];
    M.data = e->data[i]; M.ld_data = e->data_ld[i]; M.stage[0] = a.x_stage[0]; M.stage[1] = a.x_stage[1];
    M.idx = e->plan_idx[i]; M.x = a.x; M.xh = a.xp.hi; M.xl = a.xp.lo;
    M.mulv = a.mulv; M.eps = a.eps; M.z = a.z; M.c = a.c; M.S = a.S; M.g = a.g; M.den = a.den; M.rs = a.rs;
    M.inj_eps = a.inj_eps; M.ch = a.cp.hi; M.cl = a.cp.lo; M.dc = a.dc; M.dmulv = a.dmulv; M.dmh = a.dmp.hi; M.dml = a.dmp.lo;
    M.xhat = a.xhat; M.dxh = a.dxhat.hi; M.dxl = a.dxhat.lo; M.db5 = G + m.b5.off; M.rec_part = a.rec_part;
    M.dbias_heads = G + m.bmv.off; M.D = e->D[i]; M.ldD = a.ldD;
    const int D = e->D[i];
    const struct { jb::Parts Y; HPlanes H; const Seg *g, *be, *b; jb::Parts dH; HPlanes dY; int N, ld; } bl[4] = {
        {a.y1, a.h1, &m.g1, &m.be1, &m.b1, a.dh1, a.dy1, 2 * D, a.ld2D},
        {a.y2, a.h2, &m.g2, &m.be2, &m.b2, a.dh2, a.dy2, D, a.ldD},
        {a.y3, a.g1, &m.g3, &m.be3, &m.b3, a.dg1, a.dy3, D, a.ldD},
        {a.y4, a.g2, &m.g4, &m.be4, &m.b4, a.dg2, a.dy4, 2 * D, a.ld2D}};
    for (int k = 0; k < 4; ++k) {
      jb::BnLayer& l = cx.bn[k][i];
      const int bnidx = k < 2 ? (2 * i + k) : (4 + 2 * i + (k - 2));
      l.Y = bl[k].Y; l.Hh = bl[k].H.hi; l.Hl = bl[k].H.lo;
      l.gamma = T + bl[k].g->off; l.beta = T + bl[k].be->off;
      l.mean = a.bn_mean[k]; l.invstd = a.bn_inv[k]; l.var = a.bn_var[k];
      l.run_mean = e->bn_run + e->bn_off[bnidx]; l.run_var = l.run_mean + e->bn_w[bnidx];
      l.mask = a.inj_mask[k];
      l.dH = bl[k].dH; l.dYh = bl[k].dY.hi; l.dYl = bl[k].dY.lo;
      l.dgamma = G + bl[k].g->off; l.dbeta = G + bl[k].be->off; l.dbias = G + bl[k].b->off;
      l.gdyn = k < 2 ? e->dyn + 1 : nullptr;   // encoder layers sit downstream of d[mu | logvar]
      l.N = bl[k].N; l.ld = bl[k].ld; l.layer_id = static_cast<unsigned>(bnidx);
    }
    M.rec_lcw = 0; M.rec_items = 0;   // set below (both modalities share the slab width of a phase)
  }
  // slab width per element-wise phase: 16 columns per CTA item unless that leaves most of the grid idle (then 8); halved
  // further until the backward slab (2 x B x cw floats) fits the operand ring
  auto slab_lcw = [&](int n0, int n1) {
    int lcw = 4;
    if (((n0 + 15) / 16 + (n1 + 15) / 16) * 10 < e->grid * 6) lcw = 3;
    while (lcw > 0 && static_cast<size_t>(B) * (1u << lcw) * 8 > static_cast<size_t>(jb::HG_RING_BYTES)) --lcw;
    return lcw;
  };
  if (static_cast<size_t>(B) * 8 > static_cast<size_t>(jb::HG_RING_BYTES)) return fail("batch size %d too large for the slab phases", B);
  for (int k = 0; k < 4; ++k) {
    const int lcw = slab_lcw(cx.bn[k][0].N, cx.bn[k][1].N);
    cx.bn[k][0].lcw = cx.bn[k][1].lcw = lcw;
  }
  {
    const int lcw = slab_lcw(e->D[0], e->D[1]);
    for (int i = 0; i < 2; ++i) { cx.m[i].rec_lcw = lcw; cx.m[i].rec_items = (e->D[i] + (1 << lcw) - 1) >> lcw; }
  }
  cx.phase_mask = ~0ull;
  if (fused) {
    for (int ph : {jb::PH_BN1, jb::PH_BN2, jb::PH_BN3, jb::PH_BN4, jb::PH_BNB1, jb::PH_BNB2, jb::PH_BNB3, jb::PH_BNB4, jb::PH_REC})
      cx.phase_mask &= ~(1ull << ph);
    if (2 * L <= 64) { cx.phase_mask &= ~(1ull << jb::PH_REPARAM); cx.eps_early = 1; }
    for (int i = 0; i < 2; ++i) {   // one squared-error partial per (column block, cluster rank) of the last decoder GEMM
      const jb::HgProblem& hp = e->h_probs[cx.gph[5].first + i];
      cx.m[i].rec_items = hp.tiles_n * jb::HG_CLUSTER;
    }
  }
  cx.norm_tile = e->norm_tile; cx.norm_small = e->norm_small; cx.n_norm_tile = n_norm_tile;
  cx.norm_fuse = e->norm_fuse ? 1 : 0;
  {
    int nr = 0;
    auto rng = [&](const Seg& sg) { cx.norm_rng[nr][0] = static_cast<int>(sg.off); cx.norm_rng[nr][1] = sg.cols; ++nr; };
    rng(e->sigma);
    for (int i = 0; i < 2; ++i) {
      const ModSegs& m = e->ms[i];
      for (const Seg* sg : {&m.b1, &m.g1, &m.be1, &m.b2, &m.g2, &m.be2, &m.bmv, &m.b3, &m.g3, &m.be3, &m.b4, &m.g4, &m.be4, &m.b5}) rng(*sg);
    }
    cx.n_norm_rng = nr;
  }
  cx.merge_latent = (e->f_dense == nullptr && e->merge_latent) ? 1 : 0;
  if (cx.merge_latent) cx.phase_mask &= ~((1ull << jb::PH_LATLOSS) | (1ull << jb::PH_LATFIN));
  if (e->h_probs[cx.gph[8].first].fuse == jb::FUSE_LATBC) cx.phase_mask &= ~(1ull << jb::PH_LATBC);
  cx.p_diag = e->p_diag; cx.p_dense = e->p_dense; cx.f_dense = e->f_dense; cx.pn1 = e->pn1;
  cx.corr = e->corr; cx.corr_t = e->corr_t; cx.fblk = e->fblk; cx.fblk_t = e->fblk_t;
  cx.pf_ratio = e->cfg.pf_ratio; cx.f_present = e->f_dense != nullptr;
  // the row shortcut of COMBINE / LATBZ exists where GATHER builds the blocks with sk_corr_row_diag (same condition as there)
  cx.corr_hint = (e->p_diag != nullptr && e->p_dense == nullptr && e->f_dense == nullptr && !getenv("JB_NO_CORR_HINT")) ? e->corr_hint : nullptr;
  cx.lat_r = e->lat_r; cx.rowpart = e->rowpart; cx.lat_coef = e->lat_coef; cx.cosine = e->cosine;
  cx.cmax_part = e->cmax_part; cx.dmax_part = e->dmax_part; cx.dyn = e->dyn;
  {
    const float ones[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
    CU(cudaMemcpy(e->dyn, ones, sizeof ones, cudaMemcpyHostToDevice));
  }
  cx.theta = e->theta; cx.grad = e->grad; cx.adam_m = e->adam_m; cx.adam_v = e->adam_v;
  cx.theta_hi = e->theta_hi; cx.theta_lo = e->theta_lo; cx.n_flat = e->n_flat;
  cx.sigma = T + e->sigma.off; cx.dsigma = G + e->sigma.off; cx.norm_part = e->norm_part;
  cx.plan_kl = e->plan_kl; cx.out_loss = e->out_loss; cx.ctl = e->ctl;
  jb::StepConsts& sc = cx.sc;
  sc.lr = e->cfg.lr; sc.beta1 = e->cfg.beta1; sc.beta2 = e->cfg.beta2; sc.adam_eps = e->cfg.adam_eps;
  sc.max_norm = e->cfg.max_grad_norm;
  for (int k = 0; k < 4; ++k) sc.w[k] = e->cfg.loss_w[k];
  sc.pf_ratio = e->cfg.pf_ratio; sc.dropout = e->cfg.dropout;
  sc.grad_scale = 1.0f / static_cast<float>(e->cfg.world_size > 0 ? e->cfg.world_size : 1);
  sc.B = B; sc.L = L; sc.D[0] = e->D[0]; sc.D[1] = e->D[1];
  cx.gs = e->gs; cx.inv_gs = inv_gs;
  for (int q = 0; q < 8; ++q) { cx.xg[q] = e->xg[q]; cx.xf[q] = e->xf[q]; }
  cx.xrank = e->xrank; cx.xworld = e->xworld; cx.xmc = e->xworld > 1 ? e->xmc : nullptr; cx.x_adam0 = e->x_adam0;
  cx.xdbg = getenv("JB_XCHG_DBG") ? atoi(getenv("JB_XCHG_DBG")) : 0;
  cx.xfence = (getenv("JB_XCHG_FENCE") && strcmp(getenv("JB_XCHG_FENCE"), "sys") == 0) ? 1 : 0;
  cx.adam_stream = 1;
  if (const char* pv = getenv("JB_ADAM_STREAM")) cx.adam_stream = atoi(pv);
  cx.prefetch_state = 0;   // measured: WGRAD +9.4 us (the prefetch competes with the operand loads), ADAM only -3.5 us
  if (const char* pv = getenv("JB_PREFETCH_STATE")) cx.prefetch_state = atoi(pv) != 0;
  cx.dbg_repeat = 1;
  if (const char* pv = getenv("JB_DBG_REPEAT")) { if (atoi(pv) >= 1) cx.dbg_repeat = atoi(pv); }
  e->step_B = B;
  return 0;
}

int ensure_step(jb_engine* e, int B) {
  if (e->step_B == B) return 0;
  CU(cudaDeviceSynchronize());
  return build_step(e, B);
}

// k_step always runs as thread-block clusters of HG_CLUSTER CTAs (cooperative: all CTAs co-resident; the grid barrier spins).
int launch_kstep_raw(jb_engine* e, void** args, cudaStream_t s) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(e->grid); cfg.blockDim = dim3(jb::SK_THREADS); cfg.dynamicSmemBytes = jb::HG_SMEM_BYTES; cfg.stream = s;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = jb::HG_CLUSTER; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeCooperative; at[1].val.cooperative = 1;
  // JB_COOP=0 (profilers that cannot replay a cooperative cluster launch): a plain cluster launch; the grid still fits the
  // device in one wave, which is all the software grid barrier needs when nothing else runs on the GPU
  static const bool coop = !(getenv("JB_COOP") && atoi(getenv("JB_COOP")) == 0);
  cfg.attrs = at; cfg.numAttrs = coop ? 2 : 1;
  CU(cudaLaunchKernelExC(&cfg, reinterpret_cast<const void*>(jb::k_step), args));
  return 0;
}

// One launch of the step kernel: phases [lo, hi) of nsteps consecutive steps, then the control-block update.
int launch_step(jb_engine* e, int lo, int hi, int nsteps, int use_stage, cudaStream_t s, unsigned long long* ts = nullptr) {
  if (e->accumulate != e->accumulate_dev) {
    jb::k_set_accum<<<1, 1, 0, s>>>(e->ctl, e->accumulate);
    e->accumulate_dev = e->accumulate;
    ++e->launches;
  }
  int row_bias = lo > jb::PH_GATHER ? -1 : 0;
  const int fwd = lo == jb::PH_GATHER ? nsteps : 0, upd = hi == jb::PH_COUNT ? nsteps : 0;
  int3 adv = make_int3(fwd, upd, fwd > 0 ? 1 : 0);
  unsigned int* bar = e->bar + (e->bar_parity ? 32 : 0);   // this launch's counter; the kernel zeroes the other one
  unsigned int* bar_next = e->bar + (e->bar_parity ? 0 : 32);
  e->bar_parity ^= 1;
  void* args[] = {e->h_prm, &lo, &hi, &nsteps, &bar, &use_stage, &row_bias, &ts, &adv, &bar_next};
  if (launch_kstep_raw(e, args, s)) return 1;
  CU(cudaGetLastError());
  e->launches += 1;
  return 0;
}

// ------------------------------------------------------------------------------------------- eval path
int prepare_eval(jb_engine* e, cudaStream_t s) {
  if (!e->eval_dirty) return 0;
  CU(cudaMemcpyAsync(e->theta_eval, e->theta, e->n_flat * 4, cudaMemcpyDeviceToDevice, s));
  for (int i = 0; i < 2; ++i) {
    ModSegs& m = e->ms[i];
    const Seg* Ws[4] = {&m.W1, &m.W2, &m.W3, &m.W4};
    const Seg* bs[4] = {&m.b1, &m.b2, &m.b3, &m.b4};
    const Seg* gs[4] = {&m.g1, &m.g2, &m.g3, &m.g4};
    const Seg* es[4] = {&m.be1, &m.be2, &m.be3, &m.be4};
    const int bnidx[4] = {2 * i, 2 * i + 1, 4 + 2 * i, 4 + 2 * i + 1};
    for (int k = 0; k < 4; ++k) {
      jb::FoldArgs fa{};
      fa.W = e->theta + Ws[k]->off; fa.b = e->theta + bs[k]->off; fa.gamma = e->theta + gs[k]->off; fa.beta = e->theta + es[k]->off;
      fa.rm = e->bn_run + e->bn_off[bnidx[k]]; fa.rv = fa.rm + e->bn_w[bnidx[k]];
      fa.Wf = e->theta_eval + Ws[k]->off; fa.bf = e->theta_eval + bs[k]->off;
      fa.rows = Ws[k]->rows; fa.cols = Ws[k]->cols; fa.ld = Ws[k]->ld;
      jb::k_fold<<<fa.rows, 128, 0, s>>>(fa);
      ++e->launches;
    }
  }
  if (e->eval_f16) { jb::k_to_half<<<296, 256, 0, s>>>(e->theta_eval, e->theta_eval_h, e->n_flat); ++e->launches; }
  CU(cudaGetLastError());
  e->eval_dirty = false;
  return 0;
}

// Runs the folded chain on `rows` rows: in [rows, ld_in] (device, TMA-addressable) -> out [rows, ld_out].
// to < 0: encoder + mu head only.
int run_chain(jb_engine* e, int from, int to, const float* in, int ld_in, int rows, float* out, int ld_out, cudaStream_t s,
              float* bufA, float* bufB, GemmProblem* d_tab) {
  const float* T = e->theta_eval;
  const int L = e->L;
  ModSegs& mf = e->ms[from];
  const int Df = e->D[from];
  std::vector<GemmProblem> tab;
  if (e->eval_persist && e->eval_f16) {
    // fp16 chain: the wide layers exchange fp16 activations and read fp16 weights (half the operand bytes); the first
    // layer of each half reads fp32 (the caller's rows, the fp32 embedding) as TF32; embeddings and outputs stay fp32
    const __half* Th = e->theta_eval_h;
    __half* hA = reinterpret_cast<__half*>(bufA);
    __half* hB = reinterpret_cast<__half*>(bufB);
    auto addc = [&](const void* A, int lda, const Seg& Wt, const Seg& bt, void* C, int ldc, int N, int K, int epi, int f16_ops, int out_f16) {
      GemmProblem g;
      int bn = N <= 32 ? 32 : (N <= 64 ? 64 : (N >= e->eval_bn256_min ? 256 : 128));
      if (out_f16 && bn < 64) bn = 64;
      const void* Bw = f16_ops ? static_cast<const void*>(Th + Wt.off) : static_cast<const void*>(T + Wt.off);
      int rc = jb::gemm_chain_problem_fill(&g, A, lda, Bw, Wt.ld, C, ldc, rows, N, K, bn, epi, T + bt.off, jb::LRELU, f16_ops, out_f16);
      if (rc) return fail("eval tensor map encode failed (%d)", rc);
      jb::gemm_table_finalize(&g, 1);
      tab.push_back(g);
      return 0;
    };
    const int h2f = r8(2 * Df), hf = r8(Df);
    if (addc(in, ld_in, mf.W1, mf.b1, hA, h2f, 2 * Df, Df, jb::EPI_BIAS_LRELU, 0, 1)) return 1;
    if (addc(hA, h2f, mf.W2, mf.b2, hB, hf, Df, 2 * Df, jb::EPI_BIAS_LRELU, 1, 1)) return 1;
    if (to < 0) {
      if (addc(hB, hf, mf.Wmv, mf.bmv, out, ld_out, L, Df, jb::EPI_BIAS, 1, 0)) return 1;
    } else {
      ModSegs& mt = e->ms[to];
      const int Dt = e->D[to];
      const int h2t = r8(2 * Dt), ht = r8(Dt);
      if (addc(hB, hf, mf.Wmv, mf.bmv, bufA, e->LP, L, Df, jb::EPI_BIAS, 1, 0)) return 1;
      if (addc(bufA, e->LP, mt.W3, mt.b3, hB, ht, Dt, L, jb::EPI_BIAS_LRELU, 0, 1)) return 1;
      if (addc(hB, ht, mt.W4, mt.b4, hA, h2t, 2 * Dt, Dt, jb::EPI_BIAS_LRELU, 1, 1)) return 1;
      if (addc(hA, h2t, mt.W5, mt.b5, out, ld_out, Dt, 2 * Dt, jb::EPI_BIAS, 1, 0)) return 1;
    }
    CU(cudaMemcpyAsync(d_tab, tab.data(), tab.size() * sizeof(GemmProblem), cudaMemcpyHostToDevice, s));
    for (size_t k = 0; k < tab.size(); ++k) {
      CU(jb::gemm_launch_persistent(d_tab + k, tab[k], e->num_sms, s, k > 0));
      ++e->launches;
    }
    return 0;
  }
  auto add = [&](const float* A, int lda, const Seg& Wt, const Seg& bt, int n_rows_w, float* C, int ldc, int N, int K, int epi) {
    GemmProblem g;
    (void)n_rows_w;
    int bn = N <= 32 ? 32 : (N <= 64 ? 64 : (N >= e->eval_bn256_min ? 256 : 128));   // single pass: wide tiles, fewer operand bytes per output
    int rc = jb::gemm_problem_fill(&g, A, lda, 0, T + Wt.off, Wt.ld, 0, C, ldc, rows, N, K, bn, epi, T + bt.off, jb::LRELU, 0);
    if (rc) return fail("eval tensor map encode failed (%d)", rc);
    if (e->eval_persist && (rc = jb::gemm_problem_set_store_map(&g))) return fail("eval output tensor map encode failed (%d)", rc);
    jb::gemm_table_finalize(&g, 1);
    tab.push_back(g);
    return 0;
  };
  const int ld2f = r4(2 * Df), ldf = r4(Df);
  if (add(in, ld_in, mf.W1, mf.b1, 0, bufA, ld2f, 2 * Df, Df, jb::EPI_BIAS_LRELU)) return 1;
  if (add(bufA, ld2f, mf.W2, mf.b2, 0, bufB, ldf, Df, 2 * Df, jb::EPI_BIAS_LRELU)) return 1;
  if (to < 0) {
    if (add(bufB, ldf, mf.Wmv, mf.bmv, 0, out, ld_out, L, Df, jb::EPI_BIAS)) return 1;   // first L rows of Wmv = fc_mus
  } else {
    ModSegs& mt = e->ms[to];
    const int Dt = e->D[to];
    const int ld2t = r4(2 * Dt), ldt = r4(Dt);
    if (add(bufB, ldf, mf.Wmv, mf.bmv, 0, bufA, e->LP, L, Df, jb::EPI_BIAS)) return 1;
    if (add(bufA, e->LP, mt.W3, mt.b3, 0, bufB, ldt, Dt, L, jb::EPI_BIAS_LRELU)) return 1;
    if (add(bufB, ldt, mt.W4, mt.b4, 0, bufA, ld2t, 2 * Dt, Dt, jb::EPI_BIAS_LRELU)) return 1;
    if (add(bufA, ld2t, mt.W5, mt.b5, 0, out, ld_out, Dt, 2 * Dt, jb::EPI_BIAS)) return 1;
  }
  CU(cudaMemcpyAsync(d_tab, tab.data(), tab.size() * sizeof(GemmProblem), cudaMemcpyHostToDevice, s));
  for (size_t k = 0; k < tab.size(); ++k) {
    // programmatic dependent launch inside the chain: GEMM k + 1 sets up while GEMM k drains (the kernel waits for its
    // predecessor before touching memory)
    if (e->eval_persist) CU(jb::gemm_launch_persistent(d_tab + k, tab[k], e->num_sms, s, k > 0));
    else CU(jb::gemm_launch(d_tab + k, 1, tab[k].tiles_m * tab[k].tiles_n, s, k > 0));
    ++e->launches;
  }
  return 0;
}

int eval_common(jb_engine* e, int from, int to, const float* X, long long n, long long ldx, float* out, long long ldo,
                int on_device, cudaStream_t s) {
  if (from < 0 || from > 1 || to > 1) return fail("modality index out of range");
  if (n <= 0) return 0;
  const int Din = e->D[from];
  const int Dout = to < 0 ? e->L : e->D[to];
  if (ldx < Din || ldo < Dout) return fail("row pitch smaller than the row width");
  if (prepare_eval(e, s)) return 1;
  const int CH = e->eval_chunk;
  const int ldin_p = r4(Din), ldout_p = r4(Dout);
  if (on_device) {
    const bool in_ok = (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
    const bool out_ok = (ldo % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    // device-resident rows need no copy / compute double-buffering: both workspace slots form one chunk of 2 CH rows
    // (measured: 111.6 vs 105.3 M rows/s)
    const long long CHd = 2LL * CH;
    int slot = 0;
    for (long long r0 = 0; r0 < n; r0 += CHd, slot ^= 1) {
      const int rows = static_cast<int>(n - r0 < CHd ? n - r0 : CHd);
      const float* in = X + r0 * ldx;
      int ld_in = static_cast<int>(ldx);
      if (!in_ok) {
        jb::k_copy2d<<<rows, 128, 0, s>>>(in, ldx, e->ev_in, ldin_p, rows, Din); ++e->launches;
        in = e->ev_in; ld_in = ldin_p;
      }
      float* o = out + r0 * ldo;
      int ld_o = static_cast<int>(ldo);
      if (!out_ok) { o = e->ev_out; ld_o = ldout_p; }
      // the per-chunk tables alternate between two device slots; a slot is reused only after the stream has
      // consumed it (same-stream ordering of the H2D table copy after the previous chunk's kernels)
      if (run_chain(e, from, to, in, ld_in, rows, o, ld_o, s, e->ev_a, e->ev_b, e->d_ev_probs + slot * 8)) return 1;
      if (!out_ok) { jb::k_copy2d<<<rows, 128, 0, s>>>(e->ev_out, ldout_p, out + r0 * ldo, ldo, rows, Dout); ++e->launches; }
      // the host-side table vector of run_chain is pageable: the async copy has been staged by the driver on return
    }
    CU(cudaGetLastError());
    return 0;
  }
  // host pointers: stream chunks through two device staging slots so that H2D, compute and D2H overlap
  CU(cudaStreamSynchronize(s));
  const size_t in_slot = static_cast<size_t>(CH) * ldin_p, out_slot = static_cast<size_t>(CH) * ldout_p;
  const size_t a_slot = static_cast<size_t>(CH) * r4(2 * (e->D[0] > e->D[1] ? e->D[0] : e->D[1]));
  int slot = 0;
  for (long long r0 = 0; r0 < n; r0 += CH, slot ^= 1) {
    cudaStream_t cs = e->ev_stream[slot];
    const int rows = static_cast<int>(n - r0 < CH ? n - r0 : CH);
    float* din = e->ev_in + slot * in_slot;
    float* dout = e->ev_out + slot * out_slot;
    CU(cudaMemcpy2DAsync(din, static_cast<size_t>(ldin_p) * 4, X + r0 * ldx, static_cast<size_t>(ldx) * 4,
                         static_cast<size_t>(Din) * 4, rows, cudaMemcpyHostToDevice, cs));
    if (run_chain(e, from, to, din, ldin_p, rows, dout, ldout_p, cs, e->ev_a + slot * a_slot, e->ev_b + slot * a_slot,
                  e->d_ev_probs + slot * 8)) return 1;
    CU(cudaMemcpy2DAsync(out + r0 * ldo, static_cast<size_t>(ldo) * 4, dout, static_cast<size_t>(ldout_p) * 4,
                         static_cast<size_t>(Dout) * 4, rows, cudaMemcpyDeviceToHost, cs));
  }
  CU(cudaStreamSynchronize(e->ev_stream[0]));
  CU(cudaStreamSynchronize(e->ev_stream[1]));
  return 0;
}


}  // namespace

// =============================================================================================== C ABI
namespace {
template <class T> struct DevArr {   // scoped device allocation of any type
  T* p = nullptr;
  ~DevArr() { if (p) cudaFree(p); }
  int alloc(size_t n) { CU(cudaMalloc(&p, (n ? n : 1) * sizeof(T))); return 0; }
  int upload(const T* h, size_t n) { if (alloc(n)) return 1; CU(cudaMemcpy(p, h, n * sizeof(T), cudaMemcpyHostToDevice)); return 0; }
};
int metric_grid(long long tiles) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long g = static_cast<long long>(sms) * 4;
  return static_cast<int>(tiles < g ? (tiles > 0 ? tiles : 1) : g);
}
}  // namespace

extern "C" {

const char* jb_last_error(void) { return g_err.c_str(); }
int jb_version(void) { return 200; }

int jb_create(const jb_config* cfg, jb_engine** out) {
  if (!cfg || !out) return fail("null argument");
  if (cfg->dims[0] <= 0 || cfg->dims[1] <= 0 || cfg->latent <= 0 || cfg->max_batch < 2)
    return fail("invalid dims/latent/batch (%d, %d, %d, %d)", cfg->dims[0], cfg->dims[1], cfg->latent, cfg->max_batch);
  if (cfg->latent > 32 * jb::LAT_MAXT) return fail("output_dim %d exceeds the supported maximum %d", cfg->latent, 32 * jb::LAT_MAXT);
  if (!(cfg->dropout >= 0.f && cfg->dropout < 1.f)) return fail("dropout must be in [0, 1)");
  CU(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10) return fail("jamie_b200 needs an sm_100 (B200) device, found sm_%d%d", prop.major, prop.minor);
  if (!prop.cooperativeLaunch) return fail("the device does not support cooperative launches");
  jb_engine* e = new jb_engine();
  e->cfg = *cfg;
  if (const char* pv = getenv("JB_PRECISION")) e->precision_fast = strcmp(pv, "f16") == 0 || strcmp(pv, "tf32") == 0;
  e->D[0] = cfg->dims[0]; e->D[1] = cfg->dims[1]; e->L = cfg->latent; e->LP = r8(cfg->latent); e->ldmv = r8(2 * cfg->latent);
  e->Bmax = cfg->max_batch;
  build_layout(e);
  const size_t fb = static_cast<size_t>(e->n_flat + 32) * 4;
  auto alloc0 = [&](float** p, size_t bytes) -> int {
    CU(cudaMalloc(p, bytes));
    CU(cudaMemset(*p, 0, bytes));
    return 0;
  };
  // theta, the Adam moments (fp32) and the fp16 operand planes live in ONE slab (fb is a multiple of 128 B)
  e->slab_bytes = 4 * fb;
  if (alloc0(&e->state_slab, e->slab_bytes) || alloc0(&e->grad, fb) || alloc0(&e->theta_eval, fb) ||
      alloc0(&e->bn_run, e->n_bn * 4)) { jb_destroy(e); return 1; }
  e->theta = e->state_slab; e->adam_m = e->theta + fb / 4; e->adam_v = e->adam_m + fb / 4;
  e->theta_hi = reinterpret_cast<__half*>(e->adam_v + fb / 4); e->theta_lo = e->theta_hi + fb / 4;
  {  // BatchNorm defaults: running_mean 0, running_var 1; gamma = 1 is set through jb_set_params
    std::vector<float> h(e->n_bn, 0.f);
    for (int k = 0; k < 8; ++k)
      for (int j = 0; j < e->bn_w[k]; ++j) h[e->bn_off[k] + e->bn_w[k] + j] = 1.f;
    CU(cudaMemcpy(e->bn_run, h.data(), e->n_bn * 4, cudaMemcpyHostToDevice));
  }
  Carver dry(nullptr);
  carve(e, dry);
  e->arena_bytes = dry.off + 256;
  CU(cudaMalloc(&e->arena, e->arena_bytes));
  CU(cudaMemset(e->arena, 0, e->arena_bytes));
  Carver real(e->arena);
  carve(e, real);
  CU(cudaMalloc(&e->ctl, sizeof(jb::Ctl)));
  jb::Ctl c0{};
  c0.seed = cfg->seed;
  CU(cudaMemcpy(e->ctl, &c0, sizeof c0, cudaMemcpyHostToDevice));
  CU(cudaMalloc(&e->norm_part, jb::SK_MAX_CTAS * sizeof(double)));
  e->norm_tile = nullptr; e->norm_tile_cap = 0;   // sized by build_step (one partial per epilogue warp and weight-gradient work item)
  CU(cudaMalloc(&e->norm_small, jb::SK_MAX_CTAS * 4 * sizeof(float)));
  CU(cudaMemset(e->norm_small, 0, jb::SK_MAX_CTAS * 4 * sizeof(float)));
  if (const char* pv = getenv("JB_NORM_FUSE")) e->norm_fuse = atoi(pv) != 0;
  CU(cudaMalloc(&e->bar, 512));
  CU(cudaMemset(e->bar, 0, 512));
  if (const char* pv = getenv("JB_WGRAD_BN")) e->wgrad_bn = atoi(pv);
  if (const char* pv = getenv("JB_FWD_MODE")) e->fwd_mode = strcmp(pv, "medium") == 0 ? jb::HG_MEDIUM : (strcmp(pv, "precise") == 0 ? jb::HG_PRECISE : -1);
  if (const char* pv = getenv("JB_WGRAD_MODE")) e->wgrad_mode = strcmp(pv, "single") == 0 ? jb::HG_SINGLE : jb::HG_MEDIUM;
  if (const char* pv = getenv("JB_MAX_KSPLIT")) { if (atoi(pv) >= 1) e->max_ksplit = atoi(pv); }
  CU(cudaFuncSetAttribute(jb::gemm_tf32_grouped_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, jb::GEMM_SMEM_BYTES));
  CU(cudaFuncSetAttribute(jb::k_step, cudaFuncAttributeMaxDynamicSharedMemorySize, jb::HG_SMEM_BYTES));
  e->num_sms = prop.multiProcessorCount;
  {  // the step kernel is cooperative and runs as clusters of HG_CLUSTER CTAs, one CTA per SM: every cluster must be
     // co-resident (33 clusters = 132 CTAs on a 148-SM B200: GPCs with 18 SMs seat four clusters of four)
    CU(cudaFuncSetAttribute(jb::k_step, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(jb::HG_CLUSTER * 64); lc.blockDim = dim3(jb::SK_THREADS); lc.dynamicSmemBytes = jb::HG_SMEM_BYTES;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = jb::HG_CLUSTER; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 1;
    int nclusters = 0;
    CU(cudaOccupancyMaxActiveClusters(&nclusters, jb::k_step, &lc));
    if (nclusters < 1) { jb_destroy(e); return fail("the step kernel does not fit on the device (shared memory / registers / clusters)"); }
    e->grid = nclusters * jb::HG_CLUSTER;
    if (e->grid > jb::SK_MAX_CTAS) e->grid = (jb::SK_MAX_CTAS / jb::HG_CLUSTER) * jb::HG_CLUSTER;
    if (const char* pv = getenv("JB_STEP_CTAS")) { const int v = atoi(pv) / jb::HG_CLUSTER * jb::HG_CLUSTER; if (v >= jb::HG_CLUSTER && v <= e->grid) e->grid = v; }
    if (const char* pv = getenv("JB_FUSE")) e->fuse_enabled = atoi(pv) != 0;
    if (const char* pv = getenv("JB_MERGE_LATENT")) e->merge_latent = atoi(pv) != 0;
    if (const char* pv = getenv("JB_FUSE_KS")) e->fuse_ks = atoi(pv) != 0;
  }
  // rows per pass of the folded chain: one 128-row M tile per SM, so every GEMM of the chain is a whole number of waves
  // (measured on B200, 1M rows 512 -> 512: 8192 rows 59.1, 9472 rows 66.7, 18944 rows 69.0 M rows/s)
  e->eval_chunk = 128 * e->num_sms;
  if (const char* pv = getenv("JB_EVAL_CHUNK")) { if (atoi(pv) >= 128) e->eval_chunk = atoi(pv); }
  if (const char* pv = getenv("JB_EVAL_BN256_MIN")) e->eval_bn256_min = atoi(pv);
  if (const char* pv = getenv("JB_EVAL_PERSIST")) e->eval_persist = atoi(pv) != 0;
  if (const char* pv = getenv("JB_EVAL_F16")) e->eval_f16 = atoi(pv) != 0;
  CU(cudaMalloc(&e->theta_eval_h, static_cast<size_t>(e->n_flat + 32) * 2));
  // eval workspaces: two slots of chunk activations
  {
    const int Dm = e->D[0] > e->D[1] ? e->D[0] : e->D[1];
    const size_t a = static_cast<size_t>(e->eval_chunk) * r4(2 * Dm);
    const size_t io = static_cast<size_t>(e->eval_chunk) * r4(Dm > e->L ? Dm : e->L);
    CU(cudaMalloc(&e->ev_a, 2 * a * 4)); CU(cudaMalloc(&e->ev_b, 2 * a * 4));
    CU(cudaMalloc(&e->ev_in, 2 * io * 4)); CU(cudaMalloc(&e->ev_out, 2 * io * 4));
    CU(cudaMalloc(&e->d_ev_probs, 16 * sizeof(GemmProblem)));
    for (int k = 0; k < 2; ++k) CU(cudaStreamCreateWithFlags(&e->ev_stream[k], cudaStreamNonBlocking));
  }
  *out = e;
  return 0;
}

void jb_destroy(jb_engine* e) {
  if (!e) return;
  cudaDeviceSynchronize();
  if (e->h_pin) cudaFreeHost(e->h_pin);
  if (e->h2d_stream) cudaStreamDestroy(e->h2d_stream);
  for (int k = 0; k < 2; ++k) {
    if (e->ev_h2d[k]) cudaEventDestroy(e->ev_h2d[k]);
    if (e->ev_slot_free[k]) cudaEventDestroy(e->ev_slot_free[k]);
    if (e->ev_loss[k]) cudaEventDestroy(e->ev_loss[k]);
  }
  void* ptrs[] = {e->state_slab, e->grad_own ? e->grad_own : e->grad, e->theta_eval, e->bn_run, e->data[0], e->data[1], e->p_diag,
                  e->p_dense, e->f_dense, e->plan_idx[0], e->plan_idx[1], e->plan_kl, e->out_loss, e->ctl, e->norm_part,
                  e->arena, e->parts_arena, e->bar, e->d_ts, e->ev_a, e->ev_b, e->ev_in, e->ev_out,
                  e->d_ev_probs, e->theta_eval_h, e->norm_tile, e->norm_small};
  for (void* p : ptrs) if (p) cudaFree(p);
  for (int k = 0; k < 2; ++k) if (e->ev_stream[k]) cudaStreamDestroy(e->ev_stream[k]);
  delete e->h_prm;
  delete e;
}

long long jb_num_params(const jb_engine* e) { return e ? e->n_packed : 0; }
long long jb_num_bn_floats(const jb_engine* e) { return e ? e->n_bn : 0; }

int jb_set_params(jb_engine* e, const float* packed, long long n) {
  if (!e || !packed) return fail("null argument");
  if (n != e->n_packed) return fail("expected %lld parameters, got %lld", e->n_packed, n);
  CU(cudaDeviceSynchronize());
  e->eval_dirty = true;
  if (copy_packed(e, e->theta, const_cast<float*>(packed), true)) return 1;
  jb::k_hsplit_flat<<<296, 256>>>(e->theta, e->theta_hi, e->theta_lo, e->n_flat);   // operand planes of the training GEMMs
  ++e->launches;
  CU(cudaGetLastError());
  CU(cudaDeviceSynchronize());
  return 0;
}
int jb_get_params(jb_engine* e, float* packed, long long n) {
  if (!e || !packed) return fail("null argument");
  if (n != e->n_packed) return fail("expected %lld parameters, got %lld", e->n_packed, n);
  CU(cudaDeviceSynchronize());
  return copy_packed(e, e->theta, packed, false);
}
int jb_get_grads(jb_engine* e, float* packed, long long n) {
  if (!e || !packed) return fail("null argument");
  if (n != e->n_packed) return fail("expected %lld parameters, got %lld", e->n_packed, n);
  CU(cudaDeviceSynchronize());
  return copy_packed(e, e->grad, packed, false);
}
int jb_get_adam_state(jb_engine* e, float* m, float* v, long long n, long long* t) {
  if (!e || !m || !v || !t) return fail("null argument");
  if (n != e->n_packed) return fail("expected %lld parameters, got %lld", e->n_packed, n);
  CU(cudaDeviceSynchronize());
  if (copy_packed(e, e->adam_m, m, false) || copy_packed(e, e->adam_v, v, false)) return 1;
  jb::Ctl c;
  CU(cudaMemcpy(&c, e->ctl, sizeof c, cudaMemcpyDeviceToHost));
  *t = c.adam_t;
  return 0;
}
int jb_set_adam_state(jb_engine* e, const float* m, const float* v, long long n, long long t) {
  if (!e || !m || !v) return fail("null argument");
  if (n != e->n_packed) return fail("expected %lld parameters, got %lld", e->n_packed, n);
  CU(cudaDeviceSynchronize());
  if (copy_packed(e, e->adam_m, const_cast<float*>(m), true) || copy_packed(e, e->adam_v, const_cast<float*>(v), true)) return 1;
  jb::Ctl c;
  CU(cudaMemcpy(&c, e->ctl, sizeof c, cudaMemcpyDeviceToHost));
  if (e->xworld > 1) { e->x_adam0 += t - c.adam_t; e->step_B = 0; }   // exchange epochs keep counting from where they were
  c.adam_t = t;
  CU(cudaMemcpy(e->ctl, &c, sizeof c, cudaMemcpyHostToDevice));
  return 0;
}
int jb_set_bn_stats(jb_engine* e, const float* packed, long long n, const long long nbt[8]) {
  if (!e || !packed) return fail("null argument");
  if (n != e->n_bn) return fail("expected %lld BatchNorm floats, got %lld", e->n_bn, n);
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(e->bn_run, packed, n * 4, cudaMemcpyHostToDevice));
  if (nbt) for (int k = 0; k < 8; ++k) e->nbt[k] = nbt[k];
  e->eval_dirty = true;
  return 0;
}
int jb_get_bn_stats(jb_engine* e, float* packed, long long n, long long nbt[8]) {
  if (!e || !packed) return fail("null argument");
  if (n != e->n_bn) return fail("expected %lld BatchNorm floats, got %lld", e->n_bn, n);
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(packed, e->bn_run, n * 4, cudaMemcpyDeviceToHost));
  if (nbt) for (int k = 0; k < 8; ++k) nbt[k] = e->nbt[k];
  return 0;
}

int jb_set_dataset(jb_engine* e, int mod, const float* X, long long n, long long ld, int on_device, void* stream) {
  if (!e || !X) return fail("null argument");
  if (mod < 0 || mod > 1) return fail("modality index out of range");
  if (ld < e->D[mod] || n <= 0) return fail("bad dataset shape");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CU(cudaDeviceSynchronize());
  if (e->data[mod]) { cudaFree(e->data[mod]); e->data[mod] = nullptr; }
  const long long ldd = r4(e->D[mod]);
  CU(cudaMalloc(&e->data[mod], static_cast<size_t>(n) * ldd * 4));
  CU(cudaMemcpy2DAsync(e->data[mod], static_cast<size_t>(ldd) * 4, X, static_cast<size_t>(ld) * 4,
                       static_cast<size_t>(e->D[mod]) * 4, n, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
  CU(cudaStreamSynchronize(s));
  e->data_n[mod] = n; e->data_ld[mod] = ldd;
  e->step_B = 0;   // pointers are baked into the step context
  return 0;
}

int jb_set_prior_diag(jb_engine* e, const float* m, long long n) {
  if (!e) return fail("null argument");
  CU(cudaDeviceSynchronize());
  if (e->p_diag) { cudaFree(e->p_diag); e->p_diag = nullptr; }
  if (e->p_dense) { cudaFree(e->p_dense); e->p_dense = nullptr; }
  if (m) {
    CU(cudaMalloc(&e->p_diag, static_cast<size_t>(n) * 4));
    CU(cudaMemcpy(e->p_diag, m, static_cast<size_t>(n) * 4, cudaMemcpyHostToDevice));
  }
  e->p_diag_n = m ? n : 0;
  e->step_B = 0;
  return 0;
}
int jb_set_prior_dense(jb_engine* e, const float* P, long long n0, long long n1) {
  if (!e || !P) return fail("null argument");
  CU(cudaDeviceSynchronize());
  if (e->p_diag) { cudaFree(e->p_diag); e->p_diag = nullptr; }
  if (e->p_dense) { cudaFree(e->p_dense); e->p_dense = nullptr; }
  if (e->f_dense && (e->pn0 != n0 || e->pn1 != n1)) return fail("P and F shapes differ");
  CU(cudaMalloc(&e->p_dense, static_cast<size_t>(n0) * n1 * 4));
  CU(cudaMemcpy(e->p_dense, P, static_cast<size_t>(n0) * n1 * 4, cudaMemcpyHostToDevice));
  e->pn0 = n0; e->pn1 = n1;
  e->step_B = 0;
  return 0;
}
int jb_set_f_dense(jb_engine* e, const float* F, long long n0, long long n1) {
  if (!e) return fail("null argument");
  CU(cudaDeviceSynchronize());
  if (e->f_dense) { cudaFree(e->f_dense); e->f_dense = nullptr; }
  if (F) {
    if (e->p_dense && (e->pn0 != n0 || e->pn1 != n1)) return fail("P and F shapes differ");
    CU(cudaMalloc(&e->f_dense, static_cast<size_t>(n0) * n1 * 4));
    CU(cudaMemcpy(e->f_dense, F, static_cast<size_t>(n0) * n1 * 4, cudaMemcpyHostToDevice));
    e->pn0 = n0; e->pn1 = n1;
  }
  e->step_B = 0;
  return 0;
}

int jb_upload_plan(jb_engine* e, const long long* idx0, const long long* idx1, const double* kl_anneal, int nsteps, int batch,
                   void* stream) {
  if (!e || !idx0 || !idx1 || !kl_anneal) return fail("null argument");
  if (batch < 2 || batch > e->Bmax) return fail("batch %d outside [2, %d]", batch, e->Bmax);
  if (nsteps <= 0) return fail("nsteps must be positive");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CU(cudaStreamSynchronize(s));
  if (nsteps > e->plan_cap || batch != e->plan_B) {
    CU(cudaDeviceSynchronize());
    for (int i = 0; i < 2; ++i) if (e->plan_idx[i]) { cudaFree(e->plan_idx[i]); e->plan_idx[i] = nullptr; }
    if (e->plan_kl) { cudaFree(e->plan_kl); e->plan_kl = nullptr; }
    if (e->out_loss) { cudaFree(e->out_loss); e->out_loss = nullptr; }
    const int cap = nsteps > e->plan_cap ? nsteps : e->plan_cap;
    for (int i = 0; i < 2; ++i) CU(cudaMalloc(&e->plan_idx[i], static_cast<size_t>(cap) * batch * 4));
    CU(cudaMalloc(&e->plan_kl, static_cast<size_t>(cap) * 4));
    CU(cudaMalloc(&e->out_loss, static_cast<size_t>(cap) * 8 * 4));
    e->plan_cap = cap; e->plan_B = batch;
    e->step_B = 0;  // plan pointers are baked into the step context
  }
  std::vector<int> h(static_cast<size_t>(nsteps) * batch);
  const long long* src[2] = {idx0, idx1};
  for (int i = 0; i < 2; ++i) {
    for (size_t k = 0; k < h.size(); ++k) {
      const long long v = src[i][k];
      if (v < 0 || (e->data[i] && v >= e->data_n[i])) return fail("batch index %lld out of range for modality %d (n = %lld)", v, i, e->data_n[i]);
      h[k] = static_cast<int>(v);
    }
    CU(cudaMemcpy(e->plan_idx[i], h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  }
  std::vector<float> kl(nsteps);
  for (int k = 0; k < nsteps; ++k) kl[k] = static_cast<float>(32 * 1e-3 * kl_anneal[k]);
  CU(cudaMemcpy(e->plan_kl, kl.data(), static_cast<size_t>(nsteps) * 4, cudaMemcpyHostToDevice));
  CU(cudaMemset(e->out_loss, 0, static_cast<size_t>(nsteps) * 8 * 4));
  long long zero = 0;
  CU(cudaMemcpy(&e->ctl->cursor, &zero, sizeof zero, cudaMemcpyHostToDevice));
  e->plan_steps = nsteps;
  return 0;
}

int jb_inject_randomness(jb_engine* e, const float* eps0, const float* eps1, const unsigned char* const masks[8], void* stream) {
  if (!e || !eps0 || !eps1 || !masks) return fail("null argument");
  if (!e->plan_B) return fail("jb_upload_plan must be called first");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int B = e->plan_B;
  const float* eps[2] = {eps0, eps1};
  for (int i = 0; i < 2; ++i) {
    CU(cudaMemcpy2DAsync(e->act[i].inj_eps, static_cast<size_t>(e->LP) * 4, eps[i], static_cast<size_t>(e->L) * 4,
                         static_cast<size_t>(e->L) * 4, B, cudaMemcpyHostToDevice, s));
    const int D = e->D[i];
    const int w[4] = {2 * D, D, D, 2 * D};
    // draw order: enc_i d1, d2 are masks[2i], masks[2i+1]; dec_i d1, d2 are masks[4+2i], masks[4+2i+1]
    const unsigned char* src[4] = {masks[2 * i], masks[2 * i + 1], masks[4 + 2 * i], masks[4 + 2 * i + 1]};
    for (int k = 0; k < 4; ++k) {
      if (!src[k]) return fail("null mask pointer");
      CU(cudaMemcpyAsync(e->act[i].inj_mask[k], src[k], static_cast<size_t>(B) * w[k], cudaMemcpyHostToDevice, s));
    }
  }
  int one = 1;
  CU(cudaMemcpyAsync(&e->ctl->inject, &one, sizeof one, cudaMemcpyHostToDevice, s));
  CU(cudaStreamSynchronize(s));
  return 0;
}

int jb_train_steps(jb_engine* e, int nsteps, void* stream) {
  if (!e) return fail("null argument");
  if (nsteps <= 0) return 0;
  if (!e->plan_B) return fail("jb_upload_plan must be called first");
  if (!e->data[0] || !e->data[1]) return fail("jb_set_dataset must be called for both modalities before jb_train_steps");
  if (ensure_step(e, e->plan_B)) return 1;
  if (launch_step(e, jb::PH_GATHER, jb::PH_COUNT, nsteps, 0, static_cast<cudaStream_t>(stream))) return 1;
  for (int k = 0; k < 8; ++k) e->nbt[k] += nsteps;
  e->eval_dirty = true;
  return 0;
}
int jb_step_backward(jb_engine* e, void* stream) {
  if (!e) return fail("null argument");
  if (!e->plan_B) return fail("jb_upload_plan must be called first");
  if (!e->data[0] || !e->data[1]) return fail("jb_set_dataset must be called for both modalities before jb_step_backward");
  if (ensure_step(e, e->plan_B)) return 1;
  if (launch_step(e, jb::PH_GATHER, jb::PH_NORM, 1, 0, static_cast<cudaStream_t>(stream))) return 1;
  for (int k = 0; k < 8; ++k) e->nbt[k] += 1;
  e->eval_dirty = true;
  return 0;
}
int jb_step_update(jb_engine* e, void* stream) {
  if (!e) return fail("null argument");
  if (!e->step_B) return fail("jb_step_backward must run before jb_step_update");
  if (launch_step(e, jb::PH_NORM, jb::PH_COUNT, 1, 0, static_cast<cudaStream_t>(stream))) return 1;
  e->eval_dirty = true;
  return 0;
}
int jb_grad_buffer(jb_engine* e, float** dev_ptr, long long* n_floats) {
  if (!e || !dev_ptr || !n_floats) return fail("null argument");
  *dev_ptr = e->grad;
  *n_floats = e->n_flat + 8;  // flat gradients + the loss scalars appended by the loss phase
  return 0;
}
int jb_set_grad_buffer(jb_engine* e, float* dev_ptr, long long n_floats) {
  if (!e) return fail("null argument");
  CU(cudaDeviceSynchronize());
  if (dev_ptr == nullptr) {
    if (e->grad_own) e->grad = e->grad_own;
  } else {
    if (n_floats < e->n_flat + 8) return fail("gradient buffer too small (%lld floats, need %lld)", n_floats, e->n_flat + 8);
    if (reinterpret_cast<uintptr_t>(dev_ptr) & 15) return fail("gradient buffer must be 16-byte aligned");
    if (!e->grad_own) e->grad_own = e->grad;
    CU(cudaMemset(dev_ptr, 0, static_cast<size_t>(e->n_flat + 8) * 4));   // the padding stays zero forever
    e->grad = dev_ptr;
  }
  e->step_B = 0;   // the step tables hold the address
  return 0;
}
long long jb_exchange_scratch_bytes(void) { return 256 + 8LL * jb::SK_MAX_CTAS * 8; }
int jb_set_exchange(jb_engine* e, int rank, int world, float* const* grad_ptrs, unsigned int* const* flag_ptrs, float* grad_multicast) {
  if (!e) return fail("null argument");
  CU(cudaDeviceSynchronize());
  e->xworld = 0; e->xmc = nullptr;
  if (world > 1 && grad_ptrs && flag_ptrs) {
    if (world > 8 || rank < 0 || rank >= world) return fail("exchange over %d ranks (rank %d): at most 8 ranks of one box", world, rank);
    if (world != e->cfg.world_size) return fail("exchange world %d != jb_config.world_size %d", world, e->cfg.world_size);
    if (grad_ptrs[rank] != e->grad) return fail("grad_ptrs[rank] must be the buffer given to jb_set_grad_buffer");
    for (int q = 0; q < world; ++q) {
      if (!grad_ptrs[q] || !flag_ptrs[q]) return fail("null peer pointer for rank %d", q);
      e->xg[q] = grad_ptrs[q]; e->xf[q] = flag_ptrs[q];
    }
    e->xrank = rank; e->xworld = world; e->xmc = grad_multicast;
    jb::Ctl c;   // exchange epochs count from the optimizer step at which the (zeroed) scratch blocks were handed over
    CU(cudaMemcpy(&c, e->ctl, sizeof c, cudaMemcpyDeviceToHost));
    e->x_adam0 = c.adam_t;
  }
  e->step_B = 0;
  return 0;
}
int jb_set_dist_method(jb_engine* e, int method) {
  if (!e) return fail("null argument");
  if (method != 0 && method != 1) return fail("dist_method %d: 0 = euclidean, 1 = cosine", method);
  CU(cudaDeviceSynchronize());
  e->cosine = method;
  e->step_B = 0;   // the step context holds the flag
  return 0;
}
int jb_set_grad_accumulate(jb_engine* e, int accumulate) {
  if (!e) return fail("null argument");
  e->accumulate = accumulate ? 1 : 0;   // reaches the device control block with the next launch (no re-build, no sync)
  return 0;
}

// Enqueues one host-batch step in slot e->hb_slot: index / kl / row copies on the copy stream, then (after an event) the
// control-block pokes, the step kernel and the loss read-back on the caller's stream. Nothing here blocks the host.
static int hostbatch_submit(jb_engine* e, const float* x0, const float* x1, const long long* idx0, const long long* idx1,
                            int batch, double kl_anneal, bool with_update, void* stream) {
  if (!e || !x0 || !x1 || !idx0 || !idx1) return fail("null argument");
  if (with_update && e->hb_outstanding >= 2) return fail("two host-batch steps are in flight: call jb_hostbatch_wait first");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (e->plan_B != batch || e->plan_cap < 2) {   // first call: allocate a two-row plan (one row per slot) through the normal path
    std::vector<long long> z(2 * static_cast<size_t>(batch), 0);
    double k0[2] = {0, 0};
    if (jb_upload_plan(e, z.data(), z.data(), k0, 2, batch, stream)) return 1;
  }
  if (ensure_step(e, batch)) return 1;
  if (!e->h_pin) {
    CU(cudaMallocHost(reinterpret_cast<void**>(&e->h_pin), sizeof(HostPin) + 4 * static_cast<size_t>(e->Bmax) * sizeof(int)));
    e->h_pin->cursor_val[0] = 0; e->h_pin->cursor_val[1] = 1;
    e->h_pin->slot_val[0] = 0; e->h_pin->slot_val[1] = 1;
    CU(cudaStreamCreateWithFlags(&e->h2d_stream, cudaStreamNonBlocking));
    for (int k = 0; k < 2; ++k) {
      CU(cudaEventCreateWithFlags(&e->ev_h2d[k], cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&e->ev_slot_free[k], cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&e->ev_loss[k], cudaEventDisableTiming));
    }
  }
  const int slot = e->hb_slot;
  HostPin* hp = e->h_pin;
  // the slot's pinned index / kl area is reused: its previous copies (two submits ago) must have been consumed
  if (e->slot_used[slot]) CU(cudaEventSynchronize(e->ev_h2d[slot]));
  int* hi = hp->idx + static_cast<size_t>(slot) * 2 * e->Bmax;
  const long long* src[2] = {idx0, idx1};
  for (int i = 0; i < 2; ++i)
    for (int k = 0; k < batch; ++k) {
      const long long v = src[i][k];
      long long lim = 1LL << 31;
      if (e->p_diag) lim = e->p_diag_n;
      else if (e->p_dense || e->f_dense) lim = i == 0 ? e->pn0 : e->pn1;
      if (v < 0 || v >= lim) return fail("cell id %lld out of range for the prior (limit %lld)", v, lim);
      hi[i * e->Bmax + k] = static_cast<int>(v);
    }
  hp->kl[slot] = static_cast<float>(32 * 1e-3 * kl_anneal);
  // copy stream: wait until the step that last used this slot's device buffers has run, then bring the batch in
  cudaStream_t cs = e->h2d_stream;
  if (e->slot_used[slot]) CU(cudaStreamWaitEvent(cs, e->ev_slot_free[slot], 0));
  const float* xs[2] = {x0, x1};
  for (int i = 0; i < 2; ++i) {
    CU(cudaMemcpyAsync(e->plan_idx[i] + static_cast<size_t>(slot) * batch, hi + i * e->Bmax, static_cast<size_t>(batch) * 4,
                       cudaMemcpyHostToDevice, cs));
    CU(cudaMemcpy2DAsync(e->act[i].x_stage[slot], static_cast<size_t>(e->act[i].ldD) * 4, xs[i], static_cast<size_t>(e->D[i]) * 4,
                         static_cast<size_t>(e->D[i]) * 4, batch, cudaMemcpyHostToDevice, cs));
  }
  CU(cudaMemcpyAsync(e->plan_kl + slot, &hp->kl[slot], 4, cudaMemcpyHostToDevice, cs));
  CU(cudaEventRecord(e->ev_h2d[slot], cs));
  // compute stream
  CU(cudaStreamWaitEvent(s, e->ev_h2d[slot], 0));
  CU(cudaMemcpyAsync(&e->ctl->cursor, &hp->cursor_val[slot], sizeof(long long), cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(&e->ctl->host_slot, &hp->slot_val[slot], sizeof(int), cudaMemcpyHostToDevice, s));
  for (int k = 0; k < 8; ++k) e->nbt[k] += 1;
  if (e->plan_steps < 2) e->plan_steps = 2;
  e->eval_dirty = true;
  if (!with_update) {   // data-parallel form: forward + backward only
    if (launch_step(e, jb::PH_GATHER, jb::PH_NORM, 1, 1, s)) return 1;
  } else {
    if (launch_step(e, jb::PH_GATHER, jb::PH_COUNT, 1, 1, s)) return 1;
    CU(cudaMemcpyAsync(hp->losses[slot], e->out_loss + static_cast<size_t>(slot) * 8, 8 * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaEventRecord(e->ev_loss[slot], s));
    ++e->hb_outstanding;
  }
  CU(cudaEventRecord(e->ev_slot_free[slot], s));
  e->slot_used[slot] = true;
  e->hb_slot ^= 1;
  return 0;
}
int jb_hostbatch_submit(jb_engine* e, const float* x0, const float* x1, const long long* idx0, const long long* idx1, int batch,
                        double kl_anneal, void* stream) {
  return hostbatch_submit(e, x0, x1, idx0, idx1, batch, kl_anneal, true, stream);
}
int jb_hostbatch_wait(jb_engine* e, float out_losses[8]) {
  if (!e || !out_losses) return fail("null argument");
  if (e->hb_outstanding <= 0) return fail("no host-batch step in flight");
  const int slot = e->hb_oldest;
  CU(cudaEventSynchronize(e->ev_loss[slot]));
  memcpy(out_losses, e->h_pin->losses[slot], 8 * 4);
  e->hb_oldest ^= 1;
  --e->hb_outstanding;
  return 0;
}
int jb_train_step_hostbatch(jb_engine* e, const float* x0, const float* x1, const long long* idx0, const long long* idx1,
                            int batch, double kl_anneal, float out_losses[8], void* stream) {
  if (!out_losses) return fail("null argument");
  if (e && e->hb_outstanding > 0) return fail("asynchronous host-batch steps are in flight: call jb_hostbatch_wait first");
  if (hostbatch_submit(e, x0, x1, idx0, idx1, batch, kl_anneal, true, stream)) return 1;
  return jb_hostbatch_wait(e, out_losses);
}
int jb_step_backward_hostbatch(jb_engine* e, const float* x0, const float* x1, const long long* idx0, const long long* idx1,
                               int batch, double kl_anneal, void* stream) {
  return hostbatch_submit(e, x0, x1, idx0, idx1, batch, kl_anneal, false, stream);
}

int jb_num_phases(void) { return jb::PH_COUNT; }
const char* jb_phase_name(int ph) {
  static const char* names[jb::PH_COUNT] = {
      "gather+corr", "gemm enc1", "bn1", "gemm enc2", "bn2", "gemm heads", "reparam", "combine", "latloss", "gemm dec1", "bn3",
      "gemm dec2", "bn4", "gemm dec3", "rec", "dgrad W5", "bnb4", "dgrad W4", "bnb3", "dgrad W3", "latbc", "latbz",
      "dmulv planes+final", "dgrad heads", "bnb2", "dgrad W2", "bnb1", "wgrad x12", "gradnorm", "adam"};
  return ph >= 0 && ph < jb::PH_COUNT ? names[ph] : "";
}

// Times ONE phase alone: `iters` launches of k_step over [phase, phase + 1) on `stream` (CUDA events on that stream; each
// launch includes the kernel's setup, so this is an upper bound of the phase's share of a step). flops: GEMM phases only.
int jb_bench_stage(jb_engine* e, int phase, int iters, float* avg_us, double* flops, void* stream) {
  if (!e || !avg_us || !flops) return fail("null argument");
  if (phase < -1 || phase >= jb::PH_COUNT || iters <= 0) return fail("bad phase / iters");   // -1: an empty launch (setup + teardown only)
  if (!e->plan_B) return fail("jb_upload_plan must be called first");
  if (ensure_step(e, e->plan_B)) return 1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  double fl = 0;
  const int gi = jb::gemm_index(phase);
  if (gi >= 0) {
    const jb::HgPhase& ph = e->h_prm->cx.gph[gi];
    for (int i = ph.first; i < ph.first + ph.count; ++i)
      fl += 2.0 * e->h_probs[i].M * e->h_probs[i].N * e->h_probs[i].K;
  }
  int lo = phase < 0 ? 0 : phase, hi = phase + 1, one = 1, zero = 0;
  unsigned int* bar = e->bar + 64;     // (a single phase: no grid barrier)
  unsigned long long* ts = nullptr;
  int3 adv0 = make_int3(0, 0, 0);
  unsigned int* bar_next = e->bar + 96;
  void* args[] = {e->h_prm, &lo, &hi, &one, &bar, &zero, &zero, &ts, &adv0, &bar_next};
  cudaEvent_t a, b;
  CU(cudaEventCreate(&a)); CU(cudaEventCreate(&b));
  int nwarm = 3;
  if (const char* pv = getenv("JB_STAGE_WARM")) nwarm = atoi(pv);   // 0 under ncu: one profiled launch per phase
  for (int i = 0; i < nwarm; ++i)
    if (launch_kstep_raw(e, args, s)) return 1;
  CU(cudaEventRecord(a, s));
  for (int i = 0; i < iters; ++i)
    if (launch_kstep_raw(e, args, s)) return 1;
  CU(cudaEventRecord(b, s));
  CU(cudaEventSynchronize(b));
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, a, b));
  cudaEventDestroy(a); cudaEventDestroy(b);
  e->launches += iters + 3;
  *avg_us = ms * 1000.f / static_cast<float>(iters);
  *flops = fl;
  return 0;
}

// In-kernel phase timeline: runs `iters` (+1 warm-up) training steps in ONE launch with CTA 0 recording the global timer at
// every phase boundary; out_us[p] = average microseconds of phase p INSIDE the persistent kernel (barrier included).
// Consumes plan row 0 for every step (the cursor is rewound) and takes optimizer steps like jb_train_steps.
int jb_profile_step(jb_engine* e, int iters, float* out_us, int cap, int* n_launches, void* stream) {
  if (!e || !out_us || !n_launches || iters <= 0) return fail("bad argument");
  if (!e->plan_B) return fail("jb_upload_plan must be called first");
  if (!e->data[0] || !e->data[1]) return fail("jb_set_dataset must be called for both modalities first");
  if (iters + 1 > e->plan_steps) iters = e->plan_steps - 1;
  if (iters < 1) return fail("the plan needs at least two rows for a profile");
  if (ensure_step(e, e->plan_B)) return 1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t nhead = 1 + static_cast<size_t>(iters + 1) * jb::PH_COUNT;
  const size_t ndet = static_cast<size_t>(iters + 1) * jb::PH_COUNT * e->grid * 3;
  const size_t ngem = static_cast<size_t>(iters + 1) * jb::SK_NUM_GEMM * 8;
  const size_t nts = nhead + ndet + ngem + static_cast<size_t>(iters + 1) * jb::PH_COUNT * 8;
  if (e->d_ts) { cudaFree(e->d_ts); e->d_ts = nullptr; }
  CU(cudaMalloc(&e->d_ts, nts * 8));
  CU(cudaMemsetAsync(e->d_ts, 0, nts * 8, s));
  CU(cudaMemsetAsync(&e->ctl->cursor, 0, sizeof(long long), s));
  if (launch_step(e, jb::PH_GATHER, jb::PH_COUNT, iters + 1, 0, s, e->d_ts)) return 1;
  std::vector<unsigned long long> ts(nts);
  CU(cudaMemcpyAsync(ts.data(), e->d_ts, nts * 8, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  std::vector<double> acc(jb::PH_COUNT, 0.0);
  for (int it = 1; it <= iters; ++it)
    for (int p = 0; p < jb::PH_COUNT; ++p) {
      const size_t k = 1 + static_cast<size_t>(it) * jb::PH_COUNT + p;
      acc[p] += static_cast<double>(ts[k] - ts[k - 1]) * 1e-3;
    }
  {  // per-CTA detail
    const int nc = e->grid;
    auto det = [&](int it, int p, int cta, int k) { return ts[nhead + ((static_cast<size_t>(it) * jb::PH_COUNT + p) * nc + cta) * 3 + k]; };
    // SM clock rate from CTA 0: cycles per ns over steps 1 .. iters
    const double cyc = static_cast<double>(det(iters, jb::PH_COUNT - 1, 0, 1) - det(1, 0, 0, 0));
    const double ns = static_cast<double>(det(iters, jb::PH_COUNT - 1, 0, 2) - ts[1 + jb::PH_COUNT - 1]);
    const double ghz = ns > 0 ? cyc / ns : 1.9;
    // GEMM role stamps of CTA 0 (cycles after the phase began): first TMA issued, last TMA issued, first stage landed,
    // last stage landed, accumulator final, tile stored
    e->prof_gemm.assign(jb::SK_NUM_GEMM * 7, 0.f);
    for (int gi = 0; gi < jb::SK_NUM_GEMM; ++gi)
      for (int it = 1; it <= iters; ++it) {
        const unsigned long long* d = &ts[nhead + ndet + (static_cast<size_t>(it) * jb::SK_NUM_GEMM + gi) * 8];
        for (int k = 0; k < 7; ++k)
          e->prof_gemm[gi * 7 + k] += static_cast<float>(static_cast<double>(static_cast<long long>(d[k] - d[7])) / ghz * 1e-3 / iters);
      }
    if (getenv("JB_PROF_ELEM")) {   // element-wise phase stamps of CTA 0 (BatchNorm forward): cycles after the phase began
      for (int p : {static_cast<int>(jb::PH_ENC1), static_cast<int>(jb::PH_ENC2), static_cast<int>(jb::PH_DEC1), static_cast<int>(jb::PH_DEC2), static_cast<int>(jb::PH_BN1), static_cast<int>(jb::PH_BN2)}) {
        double a[7] = {0, 0, 0, 0, 0, 0, 0};
        for (int it = 1; it <= iters; ++it) {
          const unsigned long long* d = &ts[nhead + ndet + ngem + (static_cast<size_t>(it) * jb::PH_COUNT + p) * 8];
          for (int k = 0; k < 7; ++k) a[k] += static_cast<double>(static_cast<long long>(d[k] - d[7])) / ghz * 1e-3 / iters;
        }
        fprintf(stderr, "bn fwd (tail / slab) phase %d CTA0, us after the phase began: enter %.2f | pass 1 + sum %.2f | pass 2 + sum %.2f | cluster sync %.2f | merge + stats %.2f | apply %.2f | %.2f\n",
                p, a[0], a[1], a[2], a[3], a[4], a[5], a[6]);
      }
    }
    if (const char* pv = getenv("JB_PROF_CTA_PHASE")) {   // per-CTA work of one phase (us, averaged over the steps)
      const int p = atoi(pv);
      if (p >= 0 && p < jb::PH_COUNT) {
        fprintf(stderr, "phase %d work per CTA (us):", p);
        for (int c = 0; c < nc; ++c) {
          double w = 0;
          for (int it = 1; it <= iters; ++it) w += static_cast<double>(det(it, p, c, 1) - det(it, p, c, 0)) / ghz * 1e-3;
          fprintf(stderr, "%s%.1f", c % 16 == 0 ? "\n  " : " ", w / iters);
        }
        fprintf(stderr, "\n");
      }
    }
    e->prof_detail.assign(jb::PH_COUNT * 4, 0.f);
    for (int p = 0; p < jb::PH_COUNT; ++p) {
      double wmax = 0, wavg = 0, tail = 0;
      for (int it = 1; it <= iters; ++it) {
        double m = 0, a = 0;
        unsigned long long latest = 0;
        for (int c = 0; c < nc; ++c) {
          const double w = static_cast<double>(det(it, p, c, 1) - det(it, p, c, 0)) / ghz * 1e-3;
          m = w > m ? w : m; a += w;
          latest = det(it, p, c, 2) > latest ? det(it, p, c, 2) : latest;
        }
        wmax += m; wavg += a / nc;
        const unsigned long long end = ts[1 + static_cast<size_t>(it) * jb::PH_COUNT + p];
        tail += end > latest ? static_cast<double>(end - latest) * 1e-3 : 0.0;
      }
      e->prof_detail[p * 4 + 0] = static_cast<float>(acc[p] / iters);
      e->prof_detail[p * 4 + 1] = static_cast<float>(wmax / iters);
      e->prof_detail[p * 4 + 2] = static_cast<float>(wavg / iters);
      e->prof_detail[p * 4 + 3] = static_cast<float>(tail / iters);
    }
  }
  for (int k = 0; k < 8; ++k) e->nbt[k] += iters + 1;
  e->eval_dirty = true;
  *n_launches = jb::PH_COUNT;
  for (int p = 0; p < jb::PH_COUNT && p < cap; ++p) out_us[p] = static_cast<float>(acc[p] / iters);
  return 0;
}

int jb_profile_detail(jb_engine* e, float* out, int cap) {
  if (!e || !out) return fail("null argument");
  if (e->prof_detail.empty()) return fail("jb_profile_step has not run");
  for (size_t k = 0; k < e->prof_detail.size() && k < static_cast<size_t>(cap); ++k) out[k] = e->prof_detail[k];
  for (size_t k = 0; k < e->prof_gemm.size() && e->prof_detail.size() + k < static_cast<size_t>(cap); ++k) out[e->prof_detail.size() + k] = e->prof_gemm[k];
  return 0;
}

int jb_read_losses(jb_engine* e, float* out, int nsteps, void* stream) {
  if (!e || !out) return fail("null argument");
  if (nsteps > e->plan_steps) return fail("only %d steps in the plan", e->plan_steps);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CU(cudaMemcpyAsync(out, e->out_loss, static_cast<size_t>(nsteps) * 8 * 4, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return 0;
}

int jb_encode(jb_engine* e, int mod, const float* X, long long n, long long ldx, float* out, long long ldo, int on_device,
              void* stream) {
  if (!e || !X || !out) return fail("null argument");
  return eval_common(e, mod, -1, X, n, ldx, out, ldo, on_device, static_cast<cudaStream_t>(stream));
}
int jb_predict(jb_engine* e, int from, int to, const float* X, long long n, long long ldx, float* out, long long ldo,
               int on_device, void* stream) {
  if (!e || !X || !out) return fail("null argument");
  if (to < 0) return fail("modality index out of range");
  return eval_common(e, from, to, X, n, ldx, out, ldo, on_device, static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------------------------------------- PCA projection
namespace {
struct DevBuf {  // scoped device allocation
  float* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  int alloc(size_t floats) {
    CU(cudaMalloc(&p, floats * 4));
    return 0;
  }
};
int launch_one(jb_engine* e, GemmProblem& g, GemmProblem* d_slot, cudaStream_t s) {
  jb::gemm_table_finalize(&g, 1);
  CU(cudaMemcpyAsync(d_slot, &g, sizeof g, cudaMemcpyHostToDevice, s));
  CU(jb::gemm_launch(d_slot, 1, g.tiles_m * g.tiles_n, s));
  ++e->launches;
  return 0;
}
// rows per chunk so that one [rows, ld] fp32 temporary stays around 256 MB
long long pca_chunk_rows(long long ld) {
  long long r = (64LL << 20) / (ld > 0 ? ld : 1);
  if (r < 128) r = 128;
  if (r > 16384) r = 16384;
  return r;
}
}  // namespace

int jb_pca_project(jb_engine* e, const float* X, long long n, long long d, const float* comp, const float* mean, int k,
                   float m, float sdev, float* out, int on_device, void* stream) {
  if (!e || !X || !comp || !mean || !out) return fail("null argument");
  if (n <= 0 || d <= 0 || k <= 0) return fail("bad PCA shape");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long ldd = r4(static_cast<int>(d));
  const int ldk = r4(k);
  const long long R = pca_chunk_rows(ldd);
  DevBuf c_raw, c_hi, c_lo, d_mean, x_raw, x_hi, x_lo, z;
  if (c_raw.alloc(static_cast<size_t>(k) * d) || c_hi.alloc(static_cast<size_t>(k) * ldd) || c_lo.alloc(static_cast<size_t>(k) * ldd) ||
      d_mean.alloc(d) || x_hi.alloc(static_cast<size_t>(R) * ldd) || x_lo.alloc(static_cast<size_t>(R) * ldd) ||
      z.alloc(static_cast<size_t>(R) * ldk)) return 1;
  if (!on_device && x_raw.alloc(static_cast<size_t>(R) * d)) return 1;
  CU(cudaMemcpyAsync(c_raw.p, comp, static_cast<size_t>(k) * d * 4, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(d_mean.p, mean, static_cast<size_t>(d) * 4, cudaMemcpyHostToDevice, s));
  jb::k_split_tf32<<<k, 256, 0, s>>>(c_raw.p, d, k, static_cast<int>(d), 2, nullptr, 1.f, 0.f, c_hi.p, c_lo.p, ldd); ++e->launches;
  DevBuf out_dev;
  if (!on_device && out_dev.alloc(static_cast<size_t>(R) * k)) return 1;
  for (long long r0 = 0; r0 < n; r0 += R) {
    const long long rows = n - r0 < R ? n - r0 : R;
    const float* src = X + r0 * d;
    if (!on_device) {
      CU(cudaMemcpyAsync(x_raw.p, src, static_cast<size_t>(rows) * d * 4, cudaMemcpyHostToDevice, s));
      src = x_raw.p;
    }
    jb::k_split_tf32<<<static_cast<unsigned>(rows), 256, 0, s>>>(src, d, rows, static_cast<int>(d), 0, d_mean.p, 1.f, 0.f, x_hi.p, x_lo.p, ldd); ++e->launches;
    {
      GemmProblem g;   // one 3xTF32 launch on the hi/lo planes
      const int bn = k <= 32 ? 32 : 64;
      int rc = jb::gemm_problem_fill(&g, x_hi.p, static_cast<int>(ldd), 0, c_hi.p, static_cast<int>(ldd), 0, z.p, ldk,
                                     static_cast<int>(rows), k, static_cast<int>(d), bn, jb::EPI_STORE, nullptr, 0.f, 0, 0,
                                     x_lo.p, c_lo.p);
      if (rc) return fail("PCA tensor map encode failed (%d)", rc);
      if (launch_one(e, g, e->d_ev_probs, s)) return 1;
    }
    float* dst = on_device ? out + r0 * k : out_dev.p;
    jb::k_standardise<<<static_cast<unsigned>(rows), 128, 0, s>>>(z.p, ldk, rows, k, m, sdev, dst, k); ++e->launches;
    if (!on_device) CU(cudaMemcpyAsync(out + r0 * k, out_dev.p, static_cast<size_t>(rows) * k * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));  // temporaries and table slots are reused by the next chunk
  }
  CU(cudaGetLastError());
  return 0;
}

// PCA fit, GPU part (jamie/jamie.py:436-452: sklearn PCA(n_components).fit): column sums and the centred Gram matrix; the
// d x d symmetric eigenproblem stays with the caller (host LAPACK, independent of n), as does the all-reduce of both
// results when the rows are sharded over ranks.
int jb_pca_colsum(jb_engine* e, const float* X, long long n, long long d, double* colsum, int on_device, void* stream) {
  if (!e || !X || !colsum) return fail("null argument");
  if (n <= 0 || d <= 0) return fail("bad PCA shape");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long R = pca_chunk_rows(d);
  const int slabs = 64;
  DevBuf x_raw;
  DevArr<double> part;
  if (part.alloc(static_cast<size_t>(slabs) * d)) return 1;
  if (!on_device && x_raw.alloc(static_cast<size_t>(R) * d)) return 1;
  std::vector<double> h(static_cast<size_t>(slabs) * d);
  for (long long c = 0; c < d; ++c) colsum[c] = 0.0;
  for (long long r0 = 0; r0 < n; r0 += R) {
    const long long rows = n - r0 < R ? n - r0 : R;
    const float* src = X + r0 * d;
    if (!on_device) {
      CU(cudaMemcpyAsync(x_raw.p, src, static_cast<size_t>(rows) * d * 4, cudaMemcpyHostToDevice, s));
      src = x_raw.p;
    }
    const long long per = (rows + slabs - 1) / slabs;
    jb::k_col_sums<<<dim3(static_cast<unsigned>((d + 127) / 128), slabs), 128, 0, s>>>(src, rows, d, per, part.p); ++e->launches;
    CU(cudaMemcpyAsync(h.data(), part.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    for (int sl = 0; sl < slabs; ++sl)   // fixed order: deterministic
      if (sl * per < rows)
        for (long long c = 0; c < d; ++c) colsum[c] += h[static_cast<size_t>(sl) * d + c];
  }
  CU(cudaGetLastError());
  return 0;
}

int jb_pca_gram(jb_engine* e, const float* X, long long n, long long d, const double* mean, double* gram, int on_device, void* stream) {
  if (!e || !X || !mean || !gram) return fail("null argument");
  if (n <= 0 || d <= 0) return fail("bad PCA shape");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long ldd = r4(static_cast<int>(d));
  const long long R = pca_chunk_rows(ldd);
  DevBuf d_mean, x_raw, x_hi, x_lo, g32;
  DevArr<double> g64;
  if (d_mean.alloc(d) || x_hi.alloc(static_cast<size_t>(R) * ldd) || x_lo.alloc(static_cast<size_t>(R) * ldd) ||
      g32.alloc(static_cast<size_t>(d) * ldd) || g64.alloc(static_cast<size_t>(d) * d)) return 1;
  if (!on_device && x_raw.alloc(static_cast<size_t>(R) * d)) return 1;
  {
    std::vector<float> mf(d);
    for (long long c = 0; c < d; ++c) mf[c] = static_cast<float>(mean[c]);
    CU(cudaMemcpyAsync(d_mean.p, mf.data(), static_cast<size_t>(d) * 4, cudaMemcpyHostToDevice, s));
    CU(cudaStreamSynchronize(s));
  }
  CU(cudaMemsetAsync(g64.p, 0, static_cast<size_t>(d) * d * sizeof(double), s));
  for (long long r0 = 0; r0 < n; r0 += R) {
    const long long rows = n - r0 < R ? n - r0 : R;
    const float* src = X + r0 * d;
    if (!on_device) {
      CU(cudaMemcpyAsync(x_raw.p, src, static_cast<size_t>(rows) * d * 4, cudaMemcpyHostToDevice, s));
      src = x_raw.p;
    }
    jb::k_split_tf32<<<static_cast<unsigned>(rows), 256, 0, s>>>(src, d, rows, static_cast<int>(d), 0, d_mean.p, 1.f, 0.f, x_hi.p, x_lo.p, ldd); ++e->launches;
    {
      GemmProblem g;   // G[d, d] = Xc^T Xc over this chunk: both operands are the same [rows][ldd] planes, MN-major (K = rows)
      int rc = jb::gemm_problem_fill(&g, x_hi.p, static_cast<int>(ldd), 1, x_hi.p, static_cast<int>(ldd), 1, g32.p, static_cast<int>(ldd),
                                     static_cast<int>(d), static_cast<int>(d), static_cast<int>(rows), 64, jb::EPI_STORE, nullptr, 0.f, 0, 0,
                                     x_lo.p, x_lo.p);
      if (rc) return fail("PCA Gram tensor map encode failed (%d)", rc);
      if (launch_one(e, g, e->d_ev_probs, s)) return 1;
    }
    jb::k_acc_f64<<<static_cast<unsigned>(d), 256, 0, s>>>(g32.p, ldd, g64.p, d); ++e->launches;
    CU(cudaStreamSynchronize(s));  // temporaries and the table slot are reused by the next chunk
  }
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(gram, g64.p, static_cast<size_t>(d) * d * sizeof(double), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return 0;
}

int jb_pca_inverse(jb_engine* e, const float* Z, long long n, int k, const float* comp, const float* mean, long long d,
                   float m, float sdev, float* out, int on_device, void* stream) {
  if (!e || !Z || !comp || !mean || !out) return fail("null argument");
  if (n <= 0 || d <= 0 || k <= 0) return fail("bad PCA shape");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long ldd = r4(static_cast<int>(d));
  const int ldk = r4(k);
  const long long R = pca_chunk_rows(ldd);
  DevBuf c_raw, c_hi, c_lo, d_mean, z_raw, z_hi, z_lo, o;
  if (c_raw.alloc(static_cast<size_t>(k) * d) || c_hi.alloc(static_cast<size_t>(k) * ldd) || c_lo.alloc(static_cast<size_t>(k) * ldd) ||
      d_mean.alloc(d) || z_hi.alloc(static_cast<size_t>(R) * ldk) || z_lo.alloc(static_cast<size_t>(R) * ldk) ||
      o.alloc(static_cast<size_t>(R) * ldd)) return 1;
  if (!on_device && z_raw.alloc(static_cast<size_t>(R) * k)) return 1;
  CU(cudaMemcpyAsync(c_raw.p, comp, static_cast<size_t>(k) * d * 4, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(d_mean.p, mean, static_cast<size_t>(d) * 4, cudaMemcpyHostToDevice, s));
  jb::k_split_tf32<<<k, 256, 0, s>>>(c_raw.p, d, k, static_cast<int>(d), 2, nullptr, 1.f, 0.f, c_hi.p, c_lo.p, ldd); ++e->launches;
  for (long long r0 = 0; r0 < n; r0 += R) {
    const long long rows = n - r0 < R ? n - r0 : R;
    const float* src = Z + r0 * k;
    if (!on_device) {
      CU(cudaMemcpyAsync(z_raw.p, src, static_cast<size_t>(rows) * k * 4, cudaMemcpyHostToDevice, s));
      src = z_raw.p;
    }
    jb::k_split_tf32<<<static_cast<unsigned>(rows), 128, 0, s>>>(src, k, rows, k, 1, nullptr, sdev, m, z_hi.p, z_lo.p, ldk); ++e->launches;
    {
      GemmProblem g;   // out[rows, d] = A[rows, k] * comp[k, d]: B is logically [N = d, K = k] stored [k][d] -> MN-major
      int rc = jb::gemm_problem_fill(&g, z_hi.p, ldk, 0, c_hi.p, static_cast<int>(ldd), 1, o.p, static_cast<int>(ldd),
                                     static_cast<int>(rows), static_cast<int>(d), k, 64, jb::EPI_BIAS, d_mean.p, 0.f, 0, 0,
                                     z_lo.p, c_lo.p);
      if (rc) return fail("PCA tensor map encode failed (%d)", rc);
      if (launch_one(e, g, e->d_ev_probs, s)) return 1;
    }
    CU(cudaMemcpy2DAsync(out + r0 * d, static_cast<size_t>(d) * 4, o.p, static_cast<size_t>(ldd) * 4, static_cast<size_t>(d) * 4,
                         rows, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
  }
  CU(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------- evaluation metrics (N4)

int jb_metric_foscttm(const float* emb0, const float* emb1, long long n, int L, int device, unsigned long long* raw_count_closer) {
  if (!emb0 || !emb1 || !raw_count_closer) return fail("null argument");
  if (n <= 0 || L <= 0) return fail("bad embedding shape");
  CU(cudaSetDevice(device));
  DevArr<float> a, b; DevArr<double> diag; DevArr<unsigned long long> cnt;
  if (a.upload(emb0, static_cast<size_t>(n) * L) || b.upload(emb1, static_cast<size_t>(n) * L) || diag.alloc(n) || cnt.alloc(1)) return 1;
  CU(cudaMemset(cnt.p, 0, sizeof(unsigned long long)));
  jb::k_pair_diag<<<static_cast<unsigned>((n + 255) / 256), 256>>>(a.p, b.p, n, L, diag.p);
  const long long tiles = (n + jb::MT_TILE - 1) / jb::MT_TILE;
  jb::k_foscttm<<<metric_grid(tiles * tiles), jb::MT_THREADS>>>(a.p, b.p, n, L, diag.p, cnt.p);
  CU(cudaGetLastError());
  CU(cudaMemcpy(raw_count_closer, cnt.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return 0;
}

int jb_metric_knn_vote(const float* query, long long nq, const float* ref, const int* ref_class, long long nr, int L, int k,
                       int n_classes, int device, int* pred_class) {
  if (!query || !ref || !ref_class || !pred_class) return fail("null argument");
  if (nq <= 0 || nr <= 0 || L <= 0) return fail("bad embedding shape");
  if (k < 1 || k > nr) return fail("k = %d neighbours of %lld reference rows", k, nr);
  if (n_classes < 1 || n_classes > 8192) return fail("%d classes (1 .. 8192)", n_classes);
  for (long long j = 0; j < nr; ++j)
    if (ref_class[j] < 0 || ref_class[j] >= n_classes) return fail("class index %d of row %lld out of range", ref_class[j], j);
  CU(cudaSetDevice(device));
  long long Q = (512LL << 20) / (4 * nr);   // query rows per chunk: a 512 MB distance slab
  if (Q < 1) Q = 1;
  if (Q > nq) Q = nq;
  DevArr<float> q, r, D; DevArr<int> cls, pred;
  if (q.upload(query, static_cast<size_t>(nq) * L) || r.upload(ref, static_cast<size_t>(nr) * L) || cls.upload(ref_class, nr) ||
      D.alloc(static_cast<size_t>(Q) * nr) || pred.alloc(nq)) return 1;
  for (long long q0 = 0; q0 < nq; q0 += Q) {
    const long long rows = nq - q0 < Q ? nq - q0 : Q;
    const long long tiles = ((rows + jb::MT_TILE - 1) / jb::MT_TILE) * ((nr + jb::MT_TILE - 1) / jb::MT_TILE);
    jb::k_dist_rows<<<metric_grid(tiles), jb::MT_THREADS>>>(q.p + q0 * L, rows, r.p, nr, L, D.p);
    jb::k_knn_vote<<<static_cast<unsigned>(rows), jb::KV_THREADS, static_cast<size_t>(n_classes) * sizeof(int)>>>(D.p, nr, k, cls.p, n_classes, pred.p + q0);
  }
  CU(cudaGetLastError());
  CU(cudaMemcpy(pred_class, pred.p, static_cast<size_t>(nq) * sizeof(int), cudaMemcpyDeviceToHost));
  return 0;
}

int jb_metric_feature_pearson(const float* x, const float* y, long long n, long long d, int device, double* r) {
  if (!x || !y || !r) return fail("null argument");
  if (n <= 0 || d <= 0) return fail("bad matrix shape");
  CU(cudaSetDevice(device));
  const int slabs = static_cast<int>(n < 64 ? 1 : (n / 64 < 256 ? n / 64 : 256));
  const long long per = (n + slabs - 1) / slabs;
  DevArr<float> dx, dy; DevArr<double> part;
  if (dx.upload(x, static_cast<size_t>(n) * d) || dy.upload(y, static_cast<size_t>(n) * d) || part.alloc(static_cast<size_t>(slabs) * d * 5)) return 1;
  jb::k_col_moments<<<dim3(static_cast<unsigned>((d + 127) / 128), slabs), 128>>>(dx.p, dy.p, n, d, per, part.p);
  CU(cudaGetLastError());
  std::vector<double> h(static_cast<size_t>(slabs) * d * 5);
  CU(cudaMemcpy(h.data(), part.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost));
  const double N = static_cast<double>(n);
  for (long long c = 0; c < d; ++c) {
    double m[5] = {0, 0, 0, 0, 0};
    for (int sl = 0; sl < slabs; ++sl)   // fixed order: deterministic
      for (int q = 0; q < 5; ++q) m[q] += h[(static_cast<size_t>(sl) * d + c) * 5 + q];
    const double cxy = m[4] - m[0] * m[1] / N, cxx = m[2] - m[0] * m[0] / N, cyy = m[3] - m[1] * m[1] / N;
    r[c] = cxy / std::sqrt(cxx * cyy);   // constant feature: 0 / 0 = nan, as the host definition gives
  }
  return 0;
}

long long jb_debug_read(jb_engine* e, const char* name, float* out, long long cap) {
  if (!e || !name || !out) { fail("null argument"); return -1; }
  cudaDeviceSynchronize();
  const int B = e->step_B ? e->step_B : e->plan_B;
  // a tap is an fp32 array (optionally the sum of split-K partials) or a pair of fp16 operand planes (hi + lo / 2^11);
  // backward taps carry the loss scale, which is divided out here
  // (parts that the consuming phase reduces in place hold the sum in partial 0: read with parts = 1)
  struct Tap { const char* n; const float* p; const __half* hi; const __half* lo; long long rows, cols, ld; int parts; long long pstride; float scale; };
  std::vector<Tap> taps;
  char nm[2][24][16];
  const float ig = 1.f / e->gs;
  for (int i = 0; i < 2; ++i) {
    ModActs& a = e->act[i];
    const int D = e->D[i], L = e->L;
    int k = 0;
    auto f32 = [&](const char* base, const float* p, int cols, int ld, float scale = 1.f) {
      snprintf(nm[i][k], sizeof nm[i][k], "%s%d", base, i);
      taps.push_back({nm[i][k], p, nullptr, nullptr, B, cols, ld, 1, 0, scale}); ++k;
    };
    auto prt = [&](const char* base, const jb::Parts& q, int cols, int ld, float scale = 1.f, bool reduced = true) {
      snprintf(nm[i][k], sizeof nm[i][k], "%s%d", base, i);
      taps.push_back({nm[i][k], q.ptr, nullptr, nullptr, B, cols, ld, reduced ? 1 : q.n, q.stride, scale}); ++k;
    };
    auto pl = [&](const char* base, const HPlanes& h, int cols, int ld, float scale = 1.f) {
      snprintf(nm[i][k], sizeof nm[i][k], "%s%d", base, i);
      taps.push_back({nm[i][k], nullptr, h.hi, h.lo, B, cols, ld, 1, 0, scale}); ++k;
    };
    f32("x", a.x, D, a.ldD); prt("y1_", a.y1, 2 * D, a.ld2D); pl("h1_", a.h1, 2 * D, a.ld2D); prt("y2_", a.y2, D, a.ldD);
    pl("h2_", a.h2, D, a.ldD); prt("mulv", a.mulv, 2 * L, e->ldmv); f32("z", a.z, L, e->LP); f32("c", a.c, L, e->LP);
    f32("eps", a.eps, L, e->LP); pl("g1_", a.g1, D, a.ldD); pl("g2_", a.g2, 2 * D, a.ld2D); prt("xhat", a.xhat, D, a.ldD);
    pl("dxhat", a.dxhat, D, a.ldD, ig); prt("dg2_", a.dg2, 2 * D, a.ld2D, ig, false); pl("dy4_", a.dy4, 2 * D, a.ld2D, ig);
    prt("dg1_", a.dg1, D, a.ldD, ig, false); pl("dy3_", a.dy3, D, a.ldD, ig); prt("dc", a.dc, L, e->LP, ig);
    f32("dmulv", a.dmulv, 2 * L, e->ldmv, ig); prt("dh2_", a.dh2, D, a.ldD, ig, false); pl("dy2_", a.dy2, D, a.ldD, ig);
    prt("dh1_", a.dh1, 2 * D, a.ld2D, ig, false); pl("dy1_", a.dy1, 2 * D, a.ld2D, ig); f32("S", a.S, L, e->LP);
  }
  taps.push_back({"corr", e->corr, nullptr, nullptr, B, B, B, 1, 0, 1.f});
  taps.push_back({"fblk", e->fblk, nullptr, nullptr, B, B, B, 1, 0, 1.f});
  taps.push_back({"grad", e->grad, nullptr, nullptr, 1, e->n_flat, e->n_flat, 1, 0, 1.f});
  taps.push_back({"theta", e->theta, nullptr, nullptr, 1, e->n_flat, e->n_flat, 1, 0, 1.f});
  for (const Tap& t : taps) {
    if (strcmp(t.n, name) != 0) continue;
    const long long need = t.rows * t.cols;
    if (need > cap) { fail("buffer too small: need %lld floats", need); return -1; }
    if (t.p) {
      if (!e->step_B && t.parts > 0 && t.p == nullptr) { fail("no step has run yet"); return -1; }
      std::vector<float> tmp(static_cast<size_t>(need));
      for (long long q = 0; q < need; ++q) out[q] = 0.f;
      for (int p = 0; p < t.parts; ++p) {
        cudaError_t ce = cudaMemcpy2D(tmp.data(), static_cast<size_t>(t.cols) * 4, t.p + p * t.pstride, static_cast<size_t>(t.ld) * 4,
                                      static_cast<size_t>(t.cols) * 4, t.rows, cudaMemcpyDeviceToHost);
        if (ce != cudaSuccess) { fail("debug read failed: %s", cudaGetErrorString(ce)); return -1; }
        for (long long q = 0; q < need; ++q) out[q] += tmp[static_cast<size_t>(q)];
      }
    } else {
      std::vector<__half> hi(static_cast<size_t>(need)), lo(static_cast<size_t>(need));
      cudaError_t ce = cudaMemcpy2D(hi.data(), static_cast<size_t>(t.cols) * 2, t.hi, static_cast<size_t>(t.ld) * 2,
                                    static_cast<size_t>(t.cols) * 2, t.rows, cudaMemcpyDeviceToHost);
      if (ce == cudaSuccess)
        ce = cudaMemcpy2D(lo.data(), static_cast<size_t>(t.cols) * 2, t.lo, static_cast<size_t>(t.ld) * 2,
                          static_cast<size_t>(t.cols) * 2, t.rows, cudaMemcpyDeviceToHost);
      if (ce != cudaSuccess) { fail("debug read failed: %s", cudaGetErrorString(ce)); return -1; }
      for (long long q = 0; q < need; ++q) out[q] = __half2float(hi[static_cast<size_t>(q)]) + __half2float(lo[static_cast<size_t>(q)]) * jb::HG_LO_INV;
    }
    if (t.scale != 1.f) for (long long q = 0; q < need; ++q) out[q] *= t.scale;
    return need;
  }
  fail("unknown tap '%s'", name);
  return -1;
}

long long jb_launch_count(const jb_engine* e) { return e ? e->launches : 0; }

}  // extern "C"
