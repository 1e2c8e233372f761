// jamie_b200 engine: device state of one JAMIE model (both modalities' encoder/decoder MLPs, Adam moments, BatchNorm
// statistics, resident datasets) and the C ABI declared in include/jamie_b200.h.
//
// Memory layout (all fp32, one cudaMalloc arena each):
//   theta / grad / adam_m / adam_v : one flat buffer each with the SAME padded layout. Every tensor starts on a
//       128-byte boundary and 2-D weights use a row pitch rounded up to 4 floats so that TMA can address them
//       directly (tensor maps need 16-byte aligned bases and pitches). Padding stays zero forever (zero gradient ->
//       zero Adam update), so clip-norm and Adam run over the whole buffer with 128-bit accesses.
//       fc_mus.i / fc_vars.i are stored as ONE [2L, D] matrix per modality (mu rows, then logvar rows): one heads GEMM.
//   activations: [B, pitch] row-major per layer and modality, pitch = width rounded up to 4 floats.
// One training step = one CUDA graph (built once per batch size) of ~35 kernels; every step-varying scalar (plan row,
// KL anneal, Adam bias corrections, Philox stream) lives in a device-side control block.
#include <cuda.h>
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

struct Planes {  // a GEMM operand as TF32 hi / lo planes (same shape and pitch); see gemm_tf32.cuh
  float *hi = nullptr, *lo = nullptr;
};
struct ModActs {  // activations and gradients of one modality (device pointers, pitches in floats)
  float *x, *y1, *y2, *mulv, *y3, *y4, *xhat;           // fp32: gathered input and GEMM outputs
  float* x_stage[2];                                     // host-batch steps: two H2D landing buffers for x
  Planes xp, h1, h2, cp, g1, g2;                         // forward GEMM operands
  float *dg2, *dg1, *dc, *dmulv, *dh2, *dh1;             // fp32: dgrad outputs (+ dmulv from the latent backward)
  Planes dxhat, dy4, dy3, dmp, dy2, dy1;                 // backward GEMM operands
  float *z, *c, *S, *g, *eps, *inj_eps, *den, *rs;
  float *bn_mean[4], *bn_inv[4];  // enc1, enc2, dec1, dec2
  unsigned char* inj_mask[4];
  float* rec_part;
  int ldD, ld2D, ldmv, LP;
};

struct GemmStage {
  int first = 0, count = 0, ctas = 0;
  int ck = 1;   // split-K factor = cluster size of the launch (gemm_tf32.cuh)
};

}  // namespace

struct HostPin {
  long long cursor_val[2];   // {0, 1}: plan row of a slot, copied into the control block before the step graph
  int slot_val[2];           // {0, 1}
  float kl[2];
  float losses[2][8];
  int idx[1];                // [slot][modality][Bmax]
};

struct jb_engine {
  jb_config cfg{};
  int D[2]{}, L = 0, LP = 0, Bmax = 0;
  // flat parameter layout
  ModSegs ms[2];
  Seg sigma;
  long long n_flat = 0;  // padded float count (multiple of 4)
  long long n_enc = 0;   // floats [0, n_enc): sigma + both encoders; [n_enc, n_flat): heads + decoders
  std::vector<PackMap> packmap;
  long long n_packed = 0;
  float *theta = nullptr, *grad = nullptr, *adam_m = nullptr, *adam_v = nullptr, *theta_eval = nullptr;
  float *theta_hi = nullptr, *theta_lo = nullptr;   // TF32 planes of theta (same layout), rewritten by every Adam step
  float* state_slab = nullptr;   // one allocation: theta | adam_m | adam_v | theta_hi | theta_lo
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
  int *plan_idx[2]{};
  float* plan_kl = nullptr;
  float* out_loss = nullptr;
  int plan_cap = 0, plan_steps = 0, plan_B = 0;
  jb::Ctl* ctl = nullptr;
  double* norm_part = nullptr;
  // workspaces
  char* arena = nullptr;
  size_t arena_bytes = 0;
  ModActs act[2]{};
  float *corr = nullptr, *corr_t = nullptr, *fblk = nullptr, *fblk_t = nullptr, *rs_p = nullptr, *rs_f = nullptr;
  float *lat_r = nullptr, *rowpart = nullptr;
  // GEMM tables (device) for the training step at batch size graph_B
  jb::GemmProblem* d_probs = nullptr;
  std::vector<jb::GemmProblem> h_probs;
  GemmStage st_f[6], st_b[7];   // st_b[5]: all wgrads, or (data-parallel) heads + decoder wgrads with st_b[6] = encoder wgrads
  cudaGraphExec_t g_bwd_part[2]{};   // data-parallel step in two halves (see build_layout)
  bool dp_split = false;             // set by the first jb_step_backward_part: wgrads in two launches
  int launches_bwd_part[2]{};
  int graph_B = 0;
  bool graph_accum = false;
  cudaGraphExec_t g_full = nullptr, g_bwd = nullptr, g_upd = nullptr, g_host = nullptr, g_host_bwd = nullptr;
  // Host-batch steps (data resident on the host): two slots so that the copies of batch k + 1 run while step k computes.
  struct HostPin* h_pin = nullptr;   // pinned: constants, per-slot kl / losses / indices
  cudaStream_t h2d_stream = nullptr;
  cudaEvent_t ev_h2d[2]{}, ev_slot_free[2]{}, ev_loss[2]{};
  bool slot_used[2]{};
  int hb_slot = 0, hb_oldest = 0, hb_outstanding = 0;
  int launches_host = 0, launches_host_bwd = 0;
  cudaStream_t cap_stream = nullptr, side_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool use_side = true;      // P / F block build on a forked branch of the step graph (JB_SIDE=0 disables)
  int wgrad_bn = 256;        // widest N tile of the batched wgrad launch
  int dgrad_split = 1;       // dgrads in 3xTF32 (1) or one TF32 pass on the hi planes (JB_DGRAD_SPLIT=0: exploration)
  int adam_blocks = 888;     // grid of k_adam: 6 blocks of 256 threads per SM (JB_ADAM_BLOCKS; 592: +2.7 us/step)
  int slab_cw = 16;          // columns per block of the BatchNorm / reconstruction slab kernels: 16 (1024 threads) or 8 (512)
  int accumulate = 0;
  int precision_fast = 0;    // JB_PRECISION=tf32: single-pass TF32 everywhere (no parity claim)
  bool pending_inject = false;
  bool use_pdl = true;       // programmatic dependent launch between the kernels of the step graph (JB_PDL=0 disables)
  bool use_splitk = true;    // split-K clusters for stages with few tiles (JB_SPLITK=0 disables)
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
  long long launches = 0;
  int launches_per_step = 0, launches_bwd = 0, launches_upd = 0;
};

namespace {

using jb::GemmProblem;

// ------------------------------------------------------------------------------------------- layout
void add_seg(jb_engine* e, Seg& s, int rows, int cols) {
  s.rows = rows; s.cols = cols; s.ld = rows > 1 ? r4(cols) : cols;
  s.off = e->n_flat;
  e->n_flat = r32(e->n_flat + (rows > 1 ? s.span() : cols));
}
void build_layout(jb_engine* e) {
  const int L = e->L;
  e->n_flat = 0;
  add_seg(e, e->sigma, 1, 2);
  // Encoder tensors of both modalities first, then heads + decoders: the gradients of the second part are complete
  // half-way through the backward pass, so the data-parallel step can all-reduce that bucket while the encoder backward
  // still runs (jb_step_backward_part / jb_grad_bucket).
  for (int i = 0; i < 2; ++i) {
    const int D = e->D[i];
    ModSegs& m = e->ms[i];
    add_seg(e, m.W1, 2 * D, D); add_seg(e, m.b1, 1, 2 * D); add_seg(e, m.g1, 1, 2 * D); add_seg(e, m.be1, 1, 2 * D);
    add_seg(e, m.W2, D, 2 * D); add_seg(e, m.b2, 1, D); add_seg(e, m.g2, 1, D); add_seg(e, m.be2, 1, D);
  }
  e->n_enc = e->n_flat;
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
    a.ldD = r4(D); a.ld2D = r4(2 * D); a.ldmv = r4(2 * e->L); a.LP = e->LP;
    auto planes = [&](Planes& p, size_t n) { p.hi = c.take<float>(n); p.lo = c.take<float>(n); };
    a.x = c.take<float>(B * a.ldD); planes(a.xp, B * a.ldD);
    a.x_stage[0] = c.take<float>(B * a.ldD); a.x_stage[1] = c.take<float>(B * a.ldD);
    a.y1 = c.take<float>(B * a.ld2D); planes(a.h1, B * a.ld2D);
    a.y2 = c.take<float>(B * a.ldD); planes(a.h2, B * a.ldD); a.mulv = c.take<float>(B * a.ldmv);
    a.y3 = c.take<float>(B * a.ldD); planes(a.g1, B * a.ldD); a.y4 = c.take<float>(B * a.ld2D);
    planes(a.g2, B * a.ld2D); a.xhat = c.take<float>(B * a.ldD);
    planes(a.dxhat, B * a.ldD); a.dg2 = c.take<float>(B * a.ld2D); planes(a.dy4, B * a.ld2D);
    a.dg1 = c.take<float>(B * a.ldD); planes(a.dy3, B * a.ldD); a.dc = c.take<float>(B * a.LP);
    a.dmulv = c.take<float>(B * a.ldmv); planes(a.dmp, B * a.ldmv); a.dh2 = c.take<float>(B * a.ldD); planes(a.dy2, B * a.ldD);
    a.dh1 = c.take<float>(B * a.ld2D); planes(a.dy1, B * a.ld2D);
    a.z = c.take<float>(B * a.LP); a.c = c.take<float>(B * a.LP); planes(a.cp, B * a.LP); a.S = c.take<float>(B * a.LP);
    a.g = c.take<float>(B * a.LP); a.eps = c.take<float>(B * a.LP); a.inj_eps = c.take<float>(B * a.LP);
    a.den = c.take<float>(B); a.rs = c.take<float>(B);
    const int w[4] = {2 * D, D, D, 2 * D};
    for (int k = 0; k < 4; ++k) {
      a.bn_mean[k] = c.take<float>(w[k]); a.bn_inv[k] = c.take<float>(w[k]);
      a.inj_mask[k] = c.take<unsigned char>(B * w[k]);
    }
    a.rec_part = c.take<float>((D + 7) / 8);
  }
  e->corr = c.take<float>(B * B); e->corr_t = c.take<float>(B * B);
  e->fblk = c.take<float>(B * B); e->fblk_t = c.take<float>(B * B);
  e->rs_p = c.take<float>(B); e->rs_f = c.take<float>(B);
  e->lat_r = c.take<float>(B * e->LP); e->rowpart = c.take<float>(2 * B * 8);
}

// ------------------------------------------------------------------------------------------- GEMM tables
// Forward GEMMs and dgrads run the error-compensated 3xTF32 mode on hi/lo planes (fp32-class accuracy): pre-activation
// errors flip LeakyReLU' decisions and dX errors propagate down the chain. A wgrad's TF32 rounding (~3e-4 relative,
// unbiased) stays local to that gradient tensor, so wgrads are single-pass on the hi planes.
int add_prob(jb_engine* e, Planes A, int lda, int a_mn, Planes Bm, int ldb, int b_mn, float* C, int ldc, int M, int N, int K,
             int bn, int epi, const float* bias, int accumulate, int split) {
  GemmProblem g;
  if (e->precision_fast) split = 0;
  int rc = jb::gemm_problem_fill(&g, A.hi, lda, a_mn, Bm.hi, ldb, b_mn, C, ldc, M, N, K, bn, epi, bias, jb::LRELU, accumulate, 0,
                                 split ? A.lo : nullptr, split ? Bm.lo : nullptr);
  if (rc) return fail("cuTensorMapEncodeTiled failed (%d) for M%d N%d K%d lda%d ldb%d", rc, M, N, K, lda, ldb);
  e->h_probs.push_back(g);
  return 0;
}
int choose_bn(int N) {
  if (N <= 32) return 32;
  return 64;  // more CTAs beat wider tiles at these problem sizes
}
// Closes the stage made of h_probs[first ..]: picks its split-K factor and assigns CTA ranges.
void close_stage(jb_engine* e, GemmStage& st, int first) {
  st.first = first;
  st.count = static_cast<int>(e->h_probs.size()) - first;
  st.ck = e->use_splitk ? jb::gemm_pick_splitk(e->h_probs.data() + first, st.count, e->num_sms) : 1;
  st.ctas = jb::gemm_table_finalize(e->h_probs.data() + first, st.count, st.ck);
}

int build_train_tables(jb_engine* e, int B, int accum) {
  e->h_probs.clear();
  float* T = e->theta;
  float* G = e->grad;
  const int L = e->L;
  auto W = [&](const Seg& s) { return Planes{e->theta_hi + s.off, e->theta_lo + s.off}; };
  auto bias = [&](const Seg& s) { return T + s.off; };
  auto dW = [&](const Seg& s) { return G + s.off; };
  // ---- forward:  Y[B, N_out] = X W^T + b   (A = X planes K-major, B = W planes K-major)
  auto fwd = [&](GemmStage& st, auto pick) {
    const int f0 = static_cast<int>(e->h_probs.size());
    for (int i = 0; i < 2; ++i) if (pick(i)) return 1;
    close_stage(e, st, f0);
    return 0;
  };
  auto lin = [&](Planes X, int ldx, const Seg& w, const Seg& b, float* Y, int ldy, int n_out, int n_in) {
    return add_prob(e, X, ldx, 0, W(w), w.ld, 0, Y, ldy, B, n_out, n_in, choose_bn(n_out), jb::EPI_BIAS, bias(b), 0, 1);
  };
  if (fwd(e->st_f[0], [&](int i) { ModActs& a = e->act[i]; ModSegs& m = e->ms[i]; const int D = e->D[i];
        return lin(a.xp, a.ldD, m.W1, m.b1, a.y1, a.ld2D, 2 * D, D); })) return 1;
  if (fwd(e->st_f[1], [&](int i) { ModActs& a = e->act[i]; ModSegs& m = e->ms[i]; const int D = e->D[i];
        return lin(a.h1, a.ld2D, m.W2, m.b2, a.y2, a.ldD, D, 2 * D); })) return 1;
  if (fwd(e->st_f[2], [&](int i) { ModActs& a = e->act[i]; ModSegs& m = e->ms[i]; const int D = e->D[i];
        return lin(a.h2, a.ldD, m.Wmv, m.bmv, a.mulv, a.ldmv, 2 * L, D); })) return 1;
  if (fwd(e->st_f[3], [&](int i) { ModActs& a = e->act[i]; ModSegs& m = e->ms[i]; const int D = e->D[i];
        return lin(a.cp, a.LP, m.W3, m.b3, a.y3, a.ldD, D, L); })) return 1;
  if (fwd(e->st_f[4], [&](int i) { ModActs& a = e->act[i]; ModSegs& m = e->ms[i]; const int D = e->D[i];
        return lin(a.g1, a.ldD, m.W4, m.b4, a.y4, a.ld2D, 2 * D, D); })) return 1;
  if (fwd(e->st_f[5], [&](int i) { ModActs& a = e->act[i]; ModSegs& m = e->ms[i]; const int D = e->D[i];
        return lin(a.g2, a.ld2D, m.W5, m.b5, a.xhat, a.ldD, D, 2 * D); })) return 1;
  // ---- backward: dgrad dX[B, N_in]     = dY W    (A = dY planes K-major, B = W planes MN-major, K = N_out): one stage
  //                per layer on the critical path;
  //                wgrad dW[N_out, N_in] = dY^T X  (A = dY hi MN-major, B = X hi MN-major, K = batch, single pass): nothing
  //                downstream of a wgrad but the optimizer, so all twelve run as ONE launch at the end of the backward pass.
  auto wgrad = [&](Planes dY, int lddy, Planes X, int ldx, const Seg& s, int n_out, int n_in) {
    // single pass: wide tiles cut the CTA count and the operand bytes per output; with 256-wide tiles the twelve wgrads
    // of the headline shapes are ONE wave of 140 CTAs instead of 272 CTAs in two (profiles/README.md; JB_WGRAD_BN=128
    // restores the narrower tiles)
    const int bn = n_in <= 32 ? 32 : (n_in <= 64 ? 64 : (n_in >= 256 && e->wgrad_bn >= 256 ? 256 : 128));
    return add_prob(e, dY, lddy, 1, X, ldx, 1, dW(s), s.ld, n_out, n_in, B, bn, jb::EPI_STORE, nullptr, accum, 0);
  };
  auto dgrad = [&](Planes dY, int lddy, const Seg& s, float* dX, int lddx, int n_out, int n_in) {
    return add_prob(e, dY, lddy, 0, W(s), s.ld, 1, dX, lddx, B, n_in, n_out, choose_bn(n_in), jb::EPI_STORE, nullptr, 0, e->dgrad_split);
  };
  int first = static_cast<int>(e->h_probs.size());   // B6: last decoder Linear(2D -> D)
  for (int i = 0; i < 2; ++i) { ModActs& a = e->act[i]; ModSegs& m = e->ms[i]; const int D = e->D[i];
    if (dgrad(a.dxhat, a.ldD, m.W5, a.dg2, a.ld2D, D, 2 * D)) return 1; }
  close_stage(e, e->st_b[0], first);
  first = static_cast<int>(e->h_probs.size());       // B5: decoder Linear(D -> 2D)
  for (int i = 0; i < 2; ++i) { ModActs& a = e->act[i]; ModSegs& m = e->ms[i]; const int D = e->D[i];
    if (dgrad(a.dy4, a.ld2D, m.W4, a.dg1, a.ldD, 2 * D, D)) return 1; }
  close_stage(e, e->st_b[1], first);
  first = static_cast<int>(e->h_probs.size());       // B4: decoder Linear(L -> D)
  for (int i = 0; i < 2; ++i) { ModActs& a = e->act[i]; ModSegs& m = e->ms[i]; const int D = e->D[i];
    if (dgrad(a.dy3, a.ldD, m.W3, a.dc, a.LP, D, L)) return 1; }
  close_stage(e, e->st_b[2], first);
  first = static_cast<int>(e->h_probs.size());       // B3: heads Linear(D -> 2L)
  for (int i = 0; i < 2; ++i) { ModActs& a = e->act[i]; ModSegs& m = e->ms[i]; const int D = e->D[i];
    if (dgrad(a.dmp, a.ldmv, m.Wmv, a.dh2, a.ldD, 2 * L, D)) return 1; }
  close_stage(e, e->st_b[3], first);
  first = static_cast<int>(e->h_probs.size());       // B2: encoder Linear(2D -> D); Linear(D -> 2D) needs no input gradient
  for (int i = 0; i < 2; ++i) { ModActs& a = e->act[i]; ModSegs& m = e->ms[i]; const int D = e->D[i];
    if (dgrad(a.dy2, a.ldD, m.W2, a.dh1, a.ld2D, D, 2 * D)) return 1; }
  close_stage(e, e->st_b[4], first);
  // weight gradients (largest problems first): all twelve in one launch at the end. Two-part data-parallel backward
  // (jb_step_backward_part): the heads + decoder wgrads right after the latent backward (their gradient bucket is
  // all-reduced while the encoder backward runs), the encoder wgrads at the end.
  const bool split_w = e->dp_split;
  auto wgrad_stage = [&](GemmStage& st, bool dec, bool enc) {
    const int f0 = static_cast<int>(e->h_probs.size());
    for (int i = 0; i < 2; ++i) { ModActs& a = e->act[i]; ModSegs& m = e->ms[i]; const int D = e->D[i];
      if (dec && (wgrad(a.dxhat, a.ldD, a.g2, a.ld2D, m.W5, D, 2 * D) || wgrad(a.dy4, a.ld2D, a.g1, a.ldD, m.W4, 2 * D, D))) return 1;
      if (enc && (wgrad(a.dy2, a.ldD, a.h1, a.ld2D, m.W2, D, 2 * D) || wgrad(a.dy1, a.ld2D, a.xp, a.ldD, m.W1, 2 * D, D))) return 1; }
    if (dec)
      for (int i = 0; i < 2; ++i) { ModActs& a = e->act[i]; ModSegs& m = e->ms[i]; const int D = e->D[i];
        if (wgrad(a.dmp, a.ldmv, a.h2, a.ldD, m.Wmv, 2 * L, D) || wgrad(a.dy3, a.ldD, a.cp, a.LP, m.W3, D, L)) return 1; }
    close_stage(e, st, f0);
    st.ck = 1;   // wide tiles, about one wave of CTAs: no split-K
    st.ctas = jb::gemm_table_finalize(e->h_probs.data() + f0, st.count, 1);
    return 0;
  };
  e->st_b[6] = GemmStage{};
  if (split_w) { if (wgrad_stage(e->st_b[5], true, false) || wgrad_stage(e->st_b[6], false, true)) return 1; }
  else if (wgrad_stage(e->st_b[5], true, true)) return 1;
  if (e->d_probs) cudaFree(e->d_probs);
  CU(cudaMalloc(&e->d_probs, e->h_probs.size() * sizeof(GemmProblem)));
  CU(cudaMemcpy(e->d_probs, e->h_probs.data(), e->h_probs.size() * sizeof(GemmProblem), cudaMemcpyHostToDevice));
  return 0;
}

// ------------------------------------------------------------------------------------------- step recording
struct Rec {  // launches kernels on a stream and counts them
  jb_engine* e; cudaStream_t s; int n = 0; cudaError_t err = cudaSuccess;
  // Capture only: a second stream forked off the step's stream so that kernels nothing upstream depends on (the P / F
  // block build) run beside the encoder instead of in front of it. Null: everything is launched in order on s.
  cudaStream_t side = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool no_pdl_next = false;   // the next launch has a cross-stream dependency: plain (full) serialization
  bool mute = false;          // recording the other half of a two-part backward: launches are skipped


  cudaEvent_t* ev = nullptr;        // profiling: ev[k] is recorded after launch k - 1 (ev[0] before the first launch)
  const char** names = nullptr;
  void mark(const char* name) {
    if (ev && n < 63) { names[n] = name; cudaEventRecord(ev[n + 1], s); }
  }
  void gemm(const GemmStage& st) {
    if (mute) return;
    if (err == cudaSuccess)
      err = jb::gemm_launch(e->d_probs + st.first, st.count, st.ctas, s, e->use_pdl && !no_pdl_next, st.ck, e->h_probs.data() + st.first);
    no_pdl_next = false;
    mark("gemm");
    ++n;
  }
};

// Every kernel of the step starts with griddepcontrol.launch_dependents + griddepcontrol.wait (kernels.cuh), so with
// programmatic stream serialization the next kernel's launch latency and prologue overlap this kernel's execution while
// all data dependencies (transitively) still see completed, flushed predecessors.
template <typename... KP, typename... A>
void launchk(Rec& r, void (*kern)(KP...), dim3 grid, dim3 block, A... args) {
  if (r.mute) return;
  if (r.err != cudaSuccess) { ++r.n; return; }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = r.s;
  cudaLaunchAttribute at[1];
  int na = 0;
  if (r.e->use_pdl && !r.no_pdl_next) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = at;
  cfg.numAttrs = na;
  r.no_pdl_next = false;
  r.err = cudaLaunchKernelEx(&cfg, kern, static_cast<KP>(args)...);
  r.mark(nullptr);
  ++r.n;
}

jb::StepConsts make_consts(const jb_engine* e, int B) {
  jb::StepConsts sc{};
  sc.lr = e->cfg.lr; sc.beta1 = e->cfg.beta1; sc.beta2 = e->cfg.beta2; sc.adam_eps = e->cfg.adam_eps;
  sc.max_norm = e->cfg.max_grad_norm;
  for (int k = 0; k < 4; ++k) sc.w[k] = e->cfg.loss_w[k];
  sc.pf_ratio = e->cfg.pf_ratio; sc.dropout = e->cfg.dropout;
  sc.grad_scale = 1.0f / static_cast<float>(e->cfg.world_size > 0 ? e->cfg.world_size : 1);
  sc.B = B; sc.L = e->L; sc.D[0] = e->D[0]; sc.D[1] = e->D[1];
  return sc;
}

jb::Latent make_latent(jb_engine* e) {
  jb::Latent a{};
  for (int i = 0; i < 2; ++i) {
    ModActs& m = e->act[i];
    a.mulv[i] = m.mulv; a.eps[i] = m.eps; a.inj_eps[i] = m.inj_eps; a.z[i] = m.z; a.c[i] = m.c; a.S[i] = m.S;
    a.g[i] = m.g; a.den[i] = m.den; a.rs[i] = m.rs; a.dc_dec[i] = m.dc; a.dmulv[i] = m.dmulv;
    a.ch[i] = m.cp.hi; a.cl[i] = m.cp.lo; a.dmh[i] = m.dmp.hi; a.dml[i] = m.dmp.lo;
  }
  a.ldmv = e->act[0].ldmv; a.r = e->lat_r; a.rowpart = e->rowpart;
  a.corr = e->corr; a.corr_t = e->corr_t; a.fblk = e->fblk; a.fblk_t = e->fblk_t;
  a.sigma = e->theta + e->sigma.off; a.LP = e->LP; a.f_present = e->f_dense != nullptr;
  return a;
}

// part < 0: the whole forward + backward. Data-parallel halves: part 0 = everything up to the heads + decoder wgrads,
// part 1 = encoder backward and encoder wgrads.
void record_backward(jb_engine* e, Rec& r, int B, bool gather = true, int part = -1) {
  const int L = e->L;
  r.mute = part == 1;
  const float p = e->cfg.dropout;
  const int accum = e->accumulate;
  const jb::StepConsts sc = make_consts(e, B);
  float* T = e->theta;
  float* G = e->grad;
  // inputs (the first kernel also derives the step's control scalars)
  jb::GatherArgs ga{};
  for (int i = 0; i < 2; ++i) {
    ga.data[i] = e->data[i]; ga.ld_data[i] = e->data_ld[i]; ga.x[i] = e->act[i].x; ga.ldx[i] = e->act[i].ldD;
    ga.xh[i] = e->act[i].xp.hi; ga.xl[i] = e->act[i].xp.lo;
    ga.D[i] = e->D[i]; ga.idx[i] = e->plan_idx[i];
    ga.stage[0][i] = e->act[i].x_stage[0]; ga.stage[1][i] = e->act[i].x_stage[1];
  }
  if (gather) launchk(r, jb::k_gather, dim3(B, 2), dim3(128), ga, e->ctl, e->plan_kl, sc, B);
  else launchk(r, jb::k_split_x, dim3(B, 2), dim3(128), ga, e->ctl, e->plan_kl, sc, B);   // host batch: x was copied in
  // P / F blocks: needed first by k_combine, so on the side stream they overlap the encoder
  jb::CorrArgs ca{};
  ca.p_diag = e->p_diag; ca.p_dense = e->p_dense; ca.f_dense = e->f_dense; ca.n1 = e->pn1;
  ca.idx[0] = e->plan_idx[0]; ca.idx[1] = e->plan_idx[1]; ca.rs_p = e->rs_p; ca.rs_f = e->rs_f;
  ca.corr = e->corr; ca.corr_t = e->corr_t; ca.fblk = e->fblk; ca.fblk_t = e->fblk_t; ca.pf_ratio = e->cfg.pf_ratio;
  {
    cudaStream_t main_s = r.s;
    const bool fork = r.side != nullptr && r.err == cudaSuccess && !r.mute;
    if (fork) {
      if ((r.err = cudaEventRecord(r.ev_fork, main_s)) == cudaSuccess) r.err = cudaStreamWaitEvent(r.side, r.ev_fork, 0);
      r.s = r.side;
      r.no_pdl_next = true;
    }
    launchk(r, jb::k_corr_rowsum, dim3(B), dim3(128), ca, e->ctl, B);
    launchk(r, jb::k_corr_build, dim3((B + 31) / 32, (B + 31) / 32), dim3(32, 8), ca, e->ctl, B);
    if (fork) {
      if (r.err == cudaSuccess) r.err = cudaEventRecord(r.ev_join, r.side);
      r.s = main_s;
    }
  }

  auto bnf = [&](int k, int which /*0 enc1,1 enc2,2 dec1,3 dec2*/) {
    (void)k;
    jb::BnFwdPair pr{};
    for (int i = 0; i < 2; ++i) {
      ModActs& a = e->act[i]; ModSegs& m = e->ms[i]; const int D = e->D[i];
      jb::BnFwd& l = pr.l[i];
      const int bnidx = which < 2 ? (2 * i + which) : (4 + 2 * i + (which - 2));
      switch (which) {
        case 0: l.Y = a.y1; l.ldy = a.ld2D; l.Hh = a.h1.hi; l.Hl = a.h1.lo; l.ldh = a.ld2D; l.gamma = T + m.g1.off; l.beta = T + m.be1.off; l.N = 2 * D; break;
        case 1: l.Y = a.y2; l.ldy = a.ldD; l.Hh = a.h2.hi; l.Hl = a.h2.lo; l.ldh = a.ldD; l.gamma = T + m.g2.off; l.beta = T + m.be2.off; l.N = D; break;
        case 2: l.Y = a.y3; l.ldy = a.ldD; l.Hh = a.g1.hi; l.Hl = a.g1.lo; l.ldh = a.ldD; l.gamma = T + m.g3.off; l.beta = T + m.be3.off; l.N = D; break;
        default: l.Y = a.y4; l.ldy = a.ld2D; l.Hh = a.g2.hi; l.Hl = a.g2.lo; l.ldh = a.ld2D; l.gamma = T + m.g4.off; l.beta = T + m.be4.off; l.N = 2 * D; break;
      }
      l.mean = a.bn_mean[which]; l.invstd = a.bn_inv[which];
      l.run_mean = e->bn_run + e->bn_off[bnidx]; l.run_var = l.run_mean + e->bn_w[bnidx];
      l.mask = a.inj_mask[which]; l.ldm = l.N; l.layer_id = static_cast<unsigned>(bnidx); l.blocks = (l.N + 31) / 32;
    }
    if (B <= 512 && !getenv("JB_DEBUG_GENERIC_BN")) {
      const int cw = e->slab_cw;
      const dim3 grid((pr.l[0].N + cw - 1) / cw + (pr.l[1].N + cw - 1) / cw);
      if (cw == 8) launchk(r, jb::k_bn_fwd_slab<8, 512>, grid, dim3(512), pr, e->ctl, B, p);
      else launchk(r, jb::k_bn_fwd_slab<16, 1024>, grid, dim3(1024), pr, e->ctl, B, p);
    }
    else launchk(r, jb::k_bn_fwd, dim3(pr.l[0].blocks + pr.l[1].blocks), dim3(256), pr, e->ctl, B, p);
  };
  // ---- forward
  r.gemm(e->st_f[0]); bnf(0, 0);
  r.gemm(e->st_f[1]); bnf(1, 1);
  r.gemm(e->st_f[2]);
  jb::Latent lat = make_latent(e);
  launchk(r, jb::k_reparam, dim3((2 * B * L + 255) / 256), dim3(256), lat, e->ctl, B, L);
  const int wblocks = (2 * B * 32 + 255) / 256;
  if (r.side && r.err == cudaSuccess && !r.mute) {   // join: the P / F blocks are complete
    r.err = cudaStreamWaitEvent(r.s, r.ev_join, 0);
    r.no_pdl_next = true;
  }
  const int fuse_loss = lat.f_present ? 0 : 1;   // without F the loss partials need nothing of another row
  launchk(r, jb::k_combine, dim3(wblocks), dim3(256), lat, B, L, fuse_loss);
  if (!fuse_loss) launchk(r, jb::k_latent_loss, dim3(wblocks), dim3(256), lat, B, L);
  r.gemm(e->st_f[3]); bnf(2, 2);
  r.gemm(e->st_f[4]); bnf(3, 3);
  r.gemm(e->st_f[5]);
  // ---- losses + backward
  jb::RecPair rp{};
  for (int i = 0; i < 2; ++i) {
    ModActs& a = e->act[i]; ModSegs& m = e->ms[i];
    jb::RecArgs& q = rp.m[i];
    q.xhat = a.xhat; q.ldxh = a.ldD; q.x = a.x; q.ldx = a.ldD; q.dxh = a.dxhat.hi; q.dxl = a.dxhat.lo; q.lddx = a.ldD;
    const bool slab = B <= 512;
    const int cw = slab ? e->slab_cw : 32;
    q.dbias = G + m.b5.off; q.part = a.rec_part; q.D = e->D[i]; q.blocks = (e->D[i] + cw - 1) / cw;
  }
  if (B <= 512 && e->slab_cw == 8) launchk(r, jb::k_rec_slab<8, 512>, dim3(rp.m[0].blocks + rp.m[1].blocks), dim3(512), rp, B, sc.w[1], accum);
  else if (B <= 512) launchk(r, jb::k_rec_slab<16, 1024>, dim3(rp.m[0].blocks + rp.m[1].blocks), dim3(1024), rp, B, sc.w[1], accum);
  else launchk(r, jb::k_rec, dim3(rp.m[0].blocks + rp.m[1].blocks), dim3(256), rp, B, sc.w[1], accum);
  auto bnb = [&](int which) {
    jb::BnBwdPair pr{};
    for (int i = 0; i < 2; ++i) {
      ModActs& a = e->act[i]; ModSegs& m = e->ms[i]; const int D = e->D[i];
      jb::BnBwd& l = pr.l[i];
      const int bnidx = which < 2 ? (2 * i + which) : (4 + 2 * i + (which - 2));
      switch (which) {
        case 0: l.dH = a.dh1; l.lddh = a.ld2D; l.Y = a.y1; l.ldy = a.ld2D; l.dYh = a.dy1.hi; l.dYl = a.dy1.lo; l.lddy = a.ld2D; l.N = 2 * D;
                l.gamma = T + m.g1.off; l.beta = T + m.be1.off; l.dgamma = G + m.g1.off; l.dbeta = G + m.be1.off; l.dbias = G + m.b1.off; break;
        case 1: l.dH = a.dh2; l.lddh = a.ldD; l.Y = a.y2; l.ldy = a.ldD; l.dYh = a.dy2.hi; l.dYl = a.dy2.lo; l.lddy = a.ldD; l.N = D;
                l.gamma = T + m.g2.off; l.beta = T + m.be2.off; l.dgamma = G + m.g2.off; l.dbeta = G + m.be2.off; l.dbias = G + m.b2.off; break;
        case 2: l.dH = a.dg1; l.lddh = a.ldD; l.Y = a.y3; l.ldy = a.ldD; l.dYh = a.dy3.hi; l.dYl = a.dy3.lo; l.lddy = a.ldD; l.N = D;
                l.gamma = T + m.g3.off; l.beta = T + m.be3.off; l.dgamma = G + m.g3.off; l.dbeta = G + m.be3.off; l.dbias = G + m.b3.off; break;
        default: l.dH = a.dg2; l.lddh = a.ld2D; l.Y = a.y4; l.ldy = a.ld2D; l.dYh = a.dy4.hi; l.dYl = a.dy4.lo; l.lddy = a.ld2D; l.N = 2 * D;
                l.gamma = T + m.g4.off; l.beta = T + m.be4.off; l.dgamma = G + m.g4.off; l.dbeta = G + m.be4.off; l.dbias = G + m.b4.off; break;
      }
      l.mean = a.bn_mean[which]; l.invstd = a.bn_inv[which];
      l.mask = a.inj_mask[which]; l.ldm = l.N; l.layer_id = static_cast<unsigned>(bnidx); l.blocks = (l.N + 31) / 32;
    }
    if (B <= 512 && !getenv("JB_DEBUG_GENERIC_BN")) {
      const int cw = e->slab_cw;
      const dim3 grid((pr.l[0].N + cw - 1) / cw + (pr.l[1].N + cw - 1) / cw);
      if (cw == 8) launchk(r, jb::k_bn_bwd_slab<8, 512>, grid, dim3(512), pr, e->ctl, B, p, accum);
      else launchk(r, jb::k_bn_bwd_slab<16, 1024>, grid, dim3(1024), pr, e->ctl, B, p, accum);
    }
    else launchk(r, jb::k_bn_bwd, dim3(pr.l[0].blocks + pr.l[1].blocks), dim3(256), pr, e->ctl, B, p, accum);
  };
  r.gemm(e->st_b[0]); bnb(3);
  r.gemm(e->st_b[1]); bnb(2);
  r.gemm(e->st_b[2]);
  const float k_cos = sc.w[2] * 32.f * 2.f / (static_cast<float>(B) * static_cast<float>(L));
  const float k_f = sc.w[3] * 2.f / (static_cast<float>(B) * static_cast<float>(L));
  launchk(r, jb::k_latent_bwd_c, dim3(wblocks), dim3(256), lat, B, L, k_cos, k_f);
  launchk(r, jb::k_latent_bwd_z, dim3(wblocks), dim3(256), lat, e->ctl, B, L, k_cos);
  jb::FinalArgs fa{};
  fa.rowpart = e->rowpart;
  for (int i = 0; i < 2; ++i) {
    fa.rs[i] = e->act[i].rs; fa.rec_part[i] = e->act[i].rec_part; fa.rec_blocks[i] = rp.m[i].blocks;
    fa.dmulv[i] = e->act[i].dmulv; fa.dbias_heads[i] = G + e->ms[i].bmv.off; fa.D[i] = e->D[i];
  }
  fa.mulv1 = e->act[1].mulv; fa.ldmv = e->act[0].ldmv; fa.dsigma = G + e->sigma.off; fa.out_loss = e->out_loss;
  fa.grad_tail = G + e->n_flat;
  launchk(r, jb::k_latent_final, dim3(1 + (4 * L + jb::SLAB_CW - 1) / jb::SLAB_CW), dim3(jb::SLAB_THREADS), fa, e->ctl, B, L, sc, accum);
  const bool split_w = e->st_b[6].count > 0;
  if (split_w) r.gemm(e->st_b[5]);   // heads + decoder wgrads: that gradient bucket is now complete
  r.mute = part == 0;
  r.gemm(e->st_b[3]); bnb(1);
  r.gemm(e->st_b[4]); bnb(0);
  r.gemm(split_w ? e->st_b[6] : e->st_b[5]);
  r.mute = false;
}

void record_update(jb_engine* e, Rec& r, int B) {
  const jb::StepConsts sc = make_consts(e, B);
  const long long n4 = e->n_flat / 4;
  launchk(r, jb::k_gradnorm, dim3(jb::NORM_BLOCKS), dim3(256), e->grad, n4, e->norm_part);
  launchk(r, jb::k_adam, e->adam_blocks, 256, e->theta, e->theta_hi, e->theta_lo, e->grad, e->adam_m, e->adam_v, n4, e->norm_part, jb::NORM_BLOCKS, e->ctl, sc, e->out_loss);
}

int capture(jb_engine* e, int B, int what /*0 full, 1 bwd, 2 upd, 3 host, 4 host bwd, 5 / 6 bwd halves*/, cudaGraphExec_t* out,
            int* nlaunch) {
  cudaGraph_t g = nullptr;
  CU(cudaStreamBeginCapture(e->cap_stream, cudaStreamCaptureModeThreadLocal));
  Rec r{e, e->cap_stream};
  if (e->use_side) { r.side = e->side_stream; r.ev_fork = e->ev_fork; r.ev_join = e->ev_join; }
  if (what == 0 || what == 1) record_backward(e, r, B);
  if (what == 5 || what == 6) record_backward(e, r, B, true, what - 5);
  if (what == 3 || what == 4) record_backward(e, r, B, false);
  if (what == 0 || what == 2 || what == 3) record_update(e, r, B);
  cudaError_t ce = cudaStreamEndCapture(e->cap_stream, &g);
  if (r.err != cudaSuccess) { if (g) cudaGraphDestroy(g); return fail("kernel launch failed during capture: %s", cudaGetErrorString(r.err)); }
  if (ce != cudaSuccess) return fail("cudaStreamEndCapture: %s", cudaGetErrorString(ce));
  if (*out) { cudaGraphExecDestroy(*out); *out = nullptr; }
  ce = cudaGraphInstantiate(out, g, 0);
  cudaGraphDestroy(g);
  if (ce != cudaSuccess) return fail("cudaGraphInstantiate: %s", cudaGetErrorString(ce));
  *nlaunch = r.n;
  return 0;
}

int ensure_graphs(jb_engine* e, int B) {
  const bool acc = e->accumulate != 0;
  if (e->graph_B == B && e->graph_accum == acc && e->g_upd) return 0;
  if (build_train_tables(e, B, e->accumulate)) return 1;
  if (e->data[0] && e->data[1]) {   // the gathering graphs need resident datasets; the host-batch graph does not
    if (capture(e, B, 0, &e->g_full, &e->launches_per_step)) return 1;
    if (capture(e, B, 1, &e->g_bwd, &e->launches_bwd)) return 1;
    if (e->st_b[6].count > 0)
      for (int h = 0; h < 2; ++h)
        if (capture(e, B, 5 + h, &e->g_bwd_part[h], &e->launches_bwd_part[h])) return 1;
  }
  if (capture(e, B, 2, &e->g_upd, &e->launches_upd)) return 1;
  if (capture(e, B, 3, &e->g_host, &e->launches_host)) return 1;
  if (capture(e, B, 4, &e->g_host_bwd, &e->launches_host_bwd)) return 1;
  e->graph_B = B; e->graph_accum = acc;
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
extern "C" {

const char* jb_last_error(void) { return g_err.c_str(); }
int jb_version(void) { return 100; }

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
  jb_engine* e = new jb_engine();
  e->cfg = *cfg;
  if (const char* pv = getenv("JB_PRECISION")) e->precision_fast = strcmp(pv, "tf32") == 0;
  e->D[0] = cfg->dims[0]; e->D[1] = cfg->dims[1]; e->L = cfg->latent; e->LP = r4(cfg->latent); e->Bmax = cfg->max_batch;
  build_layout(e);
  const size_t fb = static_cast<size_t>(e->n_flat + 32) * 4;
  auto alloc0 = [&](float** p, size_t bytes) -> int {
    CU(cudaMalloc(p, bytes));
    CU(cudaMemset(*p, 0, bytes));
    return 0;
  };
  // theta, the Adam moments and the operand planes live in ONE slab (fb is a multiple of 128 B). A persisting-L2
  // access-policy window over it for k_adam was measured on B200: 268.8 vs 266.5 us/step without (the set-aside L2 is
  // missed by the activations), so none is set.
  e->slab_bytes = 5 * fb;
  if (alloc0(&e->state_slab, e->slab_bytes) || alloc0(&e->grad, fb) || alloc0(&e->theta_eval, fb) ||
      alloc0(&e->bn_run, e->n_bn * 4)) { jb_destroy(e); return 1; }
  e->theta = e->state_slab; e->adam_m = e->theta + fb / 4; e->adam_v = e->adam_m + fb / 4;
  e->theta_hi = e->adam_v + fb / 4; e->theta_lo = e->theta_hi + fb / 4;
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
  CU(cudaMalloc(&e->norm_part, jb::NORM_BLOCKS * sizeof(double)));
  CU(cudaStreamCreateWithFlags(&e->cap_stream, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&e->side_stream, cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
  if (const char* pv = getenv("JB_SIDE")) e->use_side = atoi(pv) != 0;
  if (const char* pv = getenv("JB_WGRAD_BN")) e->wgrad_bn = atoi(pv);
  if (const char* pv = getenv("JB_DGRAD_SPLIT")) e->dgrad_split = atoi(pv) != 0;
  if (const char* pv = getenv("JB_ADAM_BLOCKS")) { if (atoi(pv) > 0) e->adam_blocks = atoi(pv); }
  if (const char* pv = getenv("JB_SLAB_CW")) e->slab_cw = atoi(pv) == 8 ? 8 : 16;
  CU(cudaFuncSetAttribute(jb::gemm_tf32_grouped_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, jb::GEMM_SMEM_BYTES));
  if (const char* pv = getenv("JB_PDL")) e->use_pdl = atoi(pv) != 0;
  if (const char* pv = getenv("JB_SPLITK")) e->use_splitk = atoi(pv) != 0;
  e->num_sms = prop.multiProcessorCount;
  // rows per pass of the folded chain: one 128-row M tile per SM, so every GEMM of the chain is a whole number of waves
  // (measured on B200, 1M rows 512 -> 512: 8192 rows 59.1, 9472 rows 66.7, 18944 rows 69.0 M rows/s)
  e->eval_chunk = 128 * e->num_sms;
  if (const char* pv = getenv("JB_EVAL_CHUNK")) { if (atoi(pv) >= 128) e->eval_chunk = atoi(pv); }
  if (const char* pv = getenv("JB_EVAL_BN256_MIN")) e->eval_bn256_min = atoi(pv);
  if (const char* pv = getenv("JB_EVAL_PERSIST")) e->eval_persist = atoi(pv) != 0;
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
  if (e->g_full) cudaGraphExecDestroy(e->g_full);
  if (e->g_bwd) cudaGraphExecDestroy(e->g_bwd);
  if (e->g_upd) cudaGraphExecDestroy(e->g_upd);
  if (e->g_host) cudaGraphExecDestroy(e->g_host);
  if (e->g_host_bwd) cudaGraphExecDestroy(e->g_host_bwd);
  for (auto& g : e->g_bwd_part) if (g) cudaGraphExecDestroy(g);
  if (e->h_pin) cudaFreeHost(e->h_pin);
  if (e->h2d_stream) cudaStreamDestroy(e->h2d_stream);
  for (int k = 0; k < 2; ++k) {
    if (e->ev_h2d[k]) cudaEventDestroy(e->ev_h2d[k]);
    if (e->ev_slot_free[k]) cudaEventDestroy(e->ev_slot_free[k]);
    if (e->ev_loss[k]) cudaEventDestroy(e->ev_loss[k]);
  }
  void* ptrs[] = {e->state_slab, e->grad, e->theta_eval, e->bn_run, e->data[0], e->data[1], e->p_diag,
                  e->p_dense, e->f_dense, e->plan_idx[0], e->plan_idx[1], e->plan_kl, e->out_loss, e->ctl, e->norm_part,
                  e->arena, e->d_probs, e->ev_a, e->ev_b, e->ev_in, e->ev_out, e->d_ev_probs};
  for (void* p : ptrs) if (p) cudaFree(p);
  if (e->cap_stream) cudaStreamDestroy(e->cap_stream);
  if (e->side_stream) cudaStreamDestroy(e->side_stream);
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  if (e->ev_join) cudaEventDestroy(e->ev_join);
  for (int k = 0; k < 2; ++k) if (e->ev_stream[k]) cudaStreamDestroy(e->ev_stream[k]);
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
  jb::k_split_flat<<<296, 256>>>(e->theta, e->theta_hi, e->theta_lo, e->n_flat);   // operand planes of the training GEMMs
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
  if (e->g_upd) { cudaGraphExecDestroy(e->g_upd); e->g_upd = nullptr; }  // pointers are baked into the graphs
  e->graph_B = 0;
  return 0;
}

static int reset_graphs(jb_engine* e) {
  if (e->g_full) { cudaGraphExecDestroy(e->g_full); e->g_full = nullptr; }
  if (e->g_bwd) { cudaGraphExecDestroy(e->g_bwd); e->g_bwd = nullptr; }
  if (e->g_upd) { cudaGraphExecDestroy(e->g_upd); e->g_upd = nullptr; }
  e->graph_B = 0;
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
  return reset_graphs(e);
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
  return reset_graphs(e);
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
  return reset_graphs(e);
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
    reset_graphs(e);  // plan pointers are baked into the graph
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
  if (!e->plan_B) return fail("jb_upload_plan must be called first");
  if (ensure_graphs(e, e->plan_B)) return 1;
  if (!e->g_full) return fail("jb_set_dataset must be called for both modalities before jb_train_steps");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  for (int k = 0; k < nsteps; ++k) CU(cudaGraphLaunch(e->g_full, s));
  e->launches += static_cast<long long>(nsteps) * e->launches_per_step;
  for (int k = 0; k < 8; ++k) e->nbt[k] += nsteps;
  e->eval_dirty = true;
  return 0;
}
int jb_step_backward(jb_engine* e, void* stream) {
  if (!e) return fail("null argument");
  if (!e->plan_B) return fail("jb_upload_plan must be called first");
  if (ensure_graphs(e, e->plan_B)) return 1;
  if (!e->g_bwd) return fail("jb_set_dataset must be called for both modalities before jb_step_backward");
  CU(cudaGraphLaunch(e->g_bwd, static_cast<cudaStream_t>(stream)));
  e->launches += e->launches_bwd;
  for (int k = 0; k < 8; ++k) e->nbt[k] += 1;
  e->eval_dirty = true;
  return 0;
}
int jb_step_backward_part(jb_engine* e, int part, void* stream) {
  if (!e) return fail("null argument");
  if (part < 0 || part > 1) return fail("part must be 0 or 1");
  if (!e->plan_B) return fail("jb_upload_plan must be called first");
  if (e->cfg.world_size <= 1) return fail("the two-part backward is the data-parallel form: it needs world_size > 1");
  if (!e->dp_split) { e->dp_split = true; reset_graphs(e); }   // rebuild the tables with the wgrads in two launches
  if (ensure_graphs(e, e->plan_B)) return 1;
  if (!e->g_bwd_part[part]) return fail("jb_set_dataset must be called for both modalities before jb_step_backward_part");
  CU(cudaGraphLaunch(e->g_bwd_part[part], static_cast<cudaStream_t>(stream)));
  e->launches += e->launches_bwd_part[part];
  if (part == 0) for (int k = 0; k < 8; ++k) e->nbt[k] += 1;
  e->eval_dirty = true;
  return 0;
}
int jb_grad_bucket(jb_engine* e, int part, float** dev_ptr, long long* n_floats) {
  if (!e || !dev_ptr || !n_floats) return fail("null argument");
  if (part < 0 || part > 1) return fail("part must be 0 or 1");
  // part 0 finishes the heads + decoder gradients (and the loss scalars behind the buffer), part 1 the rest
  *dev_ptr = part == 0 ? e->grad + e->n_enc : e->grad;
  *n_floats = part == 0 ? e->n_flat + 8 - e->n_enc : e->n_enc;
  return 0;
}
int jb_step_update(jb_engine* e, void* stream) {
  if (!e) return fail("null argument");
  if (!e->g_upd) return fail("jb_step_backward must run before jb_step_update");
  CU(cudaGraphLaunch(e->g_upd, static_cast<cudaStream_t>(stream)));
  e->launches += e->launches_upd;
  e->eval_dirty = true;
  return 0;
}
int jb_grad_buffer(jb_engine* e, float** dev_ptr, long long* n_floats) {
  if (!e || !dev_ptr || !n_floats) return fail("null argument");
  *dev_ptr = e->grad;
  *n_floats = e->n_flat + 8;  // flat gradients + the loss scalars appended by the loss kernel
  return 0;
}
int jb_set_grad_accumulate(jb_engine* e, int accumulate) {
  if (!e) return fail("null argument");
  e->accumulate = accumulate ? 1 : 0;
  return 0;
}

// Enqueues one host-batch step in slot e->hb_slot: index / kl / row copies on the copy stream, then (after an event) the
// control-block pokes, the step graph and the loss read-back on the caller's stream. Nothing here blocks the host.
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
  if (ensure_graphs(e, batch)) return 1;
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
    CU(cudaGraphLaunch(e->g_host_bwd, s));
    e->launches += e->launches_host_bwd;
  } else {
    CU(cudaGraphLaunch(e->g_host, s));
    CU(cudaMemcpyAsync(hp->losses[slot], e->out_loss + static_cast<size_t>(slot) * 8, 8 * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaEventRecord(e->ev_loss[slot], s));
    e->launches += e->launches_host;
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

int jb_bench_stage(jb_engine* e, int stage, int iters, float* avg_us, double* flops, void* stream) {
  if (!e || !avg_us || !flops) return fail("null argument");
  if (stage < 0 || stage > 11 || iters <= 0) return fail("bad stage / iters");
  if (!e->plan_B) return fail("jb_upload_plan must be called first");
  if (ensure_graphs(e, e->plan_B)) return 1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const GemmStage& st = stage < 6 ? e->st_f[stage] : e->st_b[stage - 6];
  double fl = 0;
  for (int i = st.first; i < st.first + st.count; ++i)
    fl += 2.0 * e->h_probs[i].M * e->h_probs[i].N * e->h_probs[i].K;
  cudaEvent_t a, b;
  CU(cudaEventCreate(&a)); CU(cudaEventCreate(&b));
  for (int i = 0; i < 3; ++i) CU(jb::gemm_launch(e->d_probs + st.first, st.count, st.ctas, s, false, st.ck, e->h_probs.data() + st.first));
  CU(cudaEventRecord(a, s));
  for (int i = 0; i < iters; ++i) CU(jb::gemm_launch(e->d_probs + st.first, st.count, st.ctas, s, false, st.ck, e->h_probs.data() + st.first));
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

int jb_profile_step(jb_engine* e, int iters, float* out_us, int cap, int* n_launches, void* stream) {
  if (!e || !out_us || !n_launches || iters <= 0) return fail("bad argument");
  if (!e->plan_B) return fail("jb_upload_plan must be called first");
  if (ensure_graphs(e, e->plan_B)) return 1;
  if (!e->data[0] || !e->data[1]) return fail("jb_set_dataset must be called for both modalities first");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaEvent_t ev[65];
  const char* names[64] = {};
  for (auto& x : ev) CU(cudaEventCreate(&x));
  std::vector<double> acc(64, 0.0);
  int n = 0;
  const bool pdl = e->use_pdl;
  e->use_pdl = false;   // events between launches serialise the stream anyway
  for (int it = 0; it < iters + 1; ++it) {
    // rewind the plan cursor so that the profile can run any number of iterations
    CU(cudaMemsetAsync(&e->ctl->cursor, 0, sizeof(long long), s));
    jb::k_spin<<<1, 1, 0, s>>>(400000);   // 0.4 ms head start: the host enqueues the whole step behind it
    Rec r{e, s};
    r.ev = ev; r.names = names;
    CU(cudaEventRecord(ev[0], s));
    record_backward(e, r, e->plan_B);
    record_update(e, r, e->plan_B);
    CU(cudaStreamSynchronize(s));
    if (r.err != cudaSuccess) return fail("launch failed while profiling: %s", cudaGetErrorString(r.err));
    n = r.n < 64 ? r.n : 64;
    if (it == 0) continue;   // warm-up
    for (int k = 0; k < n; ++k) {
      float ms = 0;
      CU(cudaEventElapsedTime(&ms, ev[k], ev[k + 1]));
      acc[k] += ms * 1e3;
    }
  }
  e->use_pdl = pdl;
  for (auto& x : ev) cudaEventDestroy(x);
  for (int k = 0; k < 8; ++k) e->nbt[k] += iters + 1;
  e->launches += static_cast<long long>(iters + 1) * (n + 1);
  e->eval_dirty = true;
  *n_launches = n;
  for (int k = 0; k < n && k < cap; ++k) out_us[k] = static_cast<float>(acc[k] / iters);
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

long long jb_debug_read(jb_engine* e, const char* name, float* out, long long cap) {
  if (!e || !name || !out) { fail("null argument"); return -1; }
  cudaDeviceSynchronize();
  const int B = e->graph_B ? e->graph_B : e->plan_B;
  struct Tap { const char* n; const float* p; long long rows, cols, ld; const float* lo = nullptr; };
  std::vector<Tap> taps;
  char nm[2][24][16];
  for (int i = 0; i < 2; ++i) {
    ModActs& a = e->act[i];
    const int D = e->D[i], L = e->L;
    // tensors that only exist as operand planes are returned as hi + lo
    const struct { const char* base; const float* p; int cols, ld; const float* lo; } t[] = {
        {"x", a.x, D, a.ldD, nullptr}, {"y1_", a.y1, 2 * D, a.ld2D, nullptr}, {"h1_", a.h1.hi, 2 * D, a.ld2D, a.h1.lo},
        {"y2_", a.y2, D, a.ldD, nullptr}, {"h2_", a.h2.hi, D, a.ldD, a.h2.lo}, {"mulv", a.mulv, 2 * L, a.ldmv, nullptr},
        {"z", a.z, L, a.LP, nullptr}, {"c", a.c, L, a.LP, nullptr}, {"eps", a.eps, L, a.LP, nullptr},
        {"g1_", a.g1.hi, D, a.ldD, a.g1.lo}, {"g2_", a.g2.hi, 2 * D, a.ld2D, a.g2.lo}, {"xhat", a.xhat, D, a.ldD, nullptr},
        {"dxhat", a.dxhat.hi, D, a.ldD, a.dxhat.lo}, {"dg2_", a.dg2, 2 * D, a.ld2D, nullptr},
        {"dy4_", a.dy4.hi, 2 * D, a.ld2D, a.dy4.lo}, {"dg1_", a.dg1, D, a.ldD, nullptr}, {"dy3_", a.dy3.hi, D, a.ldD, a.dy3.lo},
        {"dc", a.dc, L, a.LP, nullptr}, {"dmulv", a.dmulv, 2 * L, a.ldmv, nullptr}, {"dh2_", a.dh2, D, a.ldD, nullptr},
        {"dy2_", a.dy2.hi, D, a.ldD, a.dy2.lo}, {"dh1_", a.dh1, 2 * D, a.ld2D, nullptr},
        {"dy1_", a.dy1.hi, 2 * D, a.ld2D, a.dy1.lo}, {"S", a.S, L, a.LP, nullptr}};
    int k = 0;
    for (const auto& q : t) {
      snprintf(nm[i][k], sizeof nm[i][k], "%s%d", q.base, i);
      taps.push_back({nm[i][k], q.p, B, q.cols, q.ld, q.lo});
      ++k;
    }
  }
  taps.push_back({"corr", e->corr, B, B, B});
  taps.push_back({"fblk", e->fblk, B, B, B});
  taps.push_back({"grad", e->grad, 1, e->n_flat, e->n_flat});
  taps.push_back({"theta", e->theta, 1, e->n_flat, e->n_flat});
  for (const Tap& t : taps) {
    if (strcmp(t.n, name) == 0) {
      const long long need = t.rows * t.cols;
      if (need > cap) { fail("buffer too small: need %lld floats", need); return -1; }
      cudaError_t ce = cudaMemcpy2D(out, static_cast<size_t>(t.cols) * 4, t.p, static_cast<size_t>(t.ld) * 4,
                                    static_cast<size_t>(t.cols) * 4, t.rows, cudaMemcpyDeviceToHost);
      if (ce != cudaSuccess) { fail("debug read failed: %s", cudaGetErrorString(ce)); return -1; }
      if (t.lo) {
        std::vector<float> lo(static_cast<size_t>(need));
        ce = cudaMemcpy2D(lo.data(), static_cast<size_t>(t.cols) * 4, t.lo, static_cast<size_t>(t.ld) * 4,
                          static_cast<size_t>(t.cols) * 4, t.rows, cudaMemcpyDeviceToHost);
        if (ce != cudaSuccess) { fail("debug read failed: %s", cudaGetErrorString(ce)); return -1; }
        for (long long q = 0; q < need; ++q) out[q] += lo[static_cast<size_t>(q)];
      }
      return need;
    }
  }
  fail("unknown tap '%s'", name);
  return -1;
}

long long jb_launch_count(const jb_engine* e) { return e ? e->launches : 0; }

}  // extern "C"
