// Non-GEMM kernels of the JAMIE train step: batch gather, P/F block build, BatchNorm(+LeakyReLU+Dropout) forward and
// backward over column slabs, reparameterisation, correspondence-weighted latent combination and its backward,
// the fused latent loss/gradient kernels, reconstruction loss, global-norm + clip + Adam over the flat buffer.
// All reductions are fixed-order (no floating-point atomics): a step is bit-reproducible run to run.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "ptx.cuh"

namespace jb {

constexpr float BN_EPS = 1e-5f;
constexpr float BN_MOM = 0.1f;
constexpr float LRELU = 0.01f;
constexpr int NORM_BLOCKS = 592;  // 4 x 148 SMs: blocks (and partials) of k_gradnorm

// ------------------------------------------------------------------------------------------------ step control
// Device-resident per-step scalars so that one captured CUDA graph serves every step.
struct Ctl {
  long long cursor;   // next plan row to run
  long long adam_t;   // optimizer steps taken so far
  int row;            // plan row of the step in flight
  int inject;         // use injected eps / masks for this step (cleared at the end of the step)
  float kl_coef;      // w_KL * 0.032 * anneal(epoch) for this step
  float step_size;    // lr / (1 - beta1^t)
  float inv_bc2_sqrt; // 1 / sqrt(1 - beta2^t)
  float kl_base;      // 0.032 * anneal(epoch) (the reference's KL scale, before loss_weights)
  unsigned long long seed;
  unsigned long long stream_id;  // philox counter word: distinct per step
  int host_slot;      // host-batch steps: which of the two device staging buffers holds this step's rows
  int accum;          // batch_step=False: gradients of this backward pass are added to the buffer (jamie/jamie.py:744-749)
};

struct StepConsts {
  float lr, beta1, beta2, adam_eps, max_norm;
  float w[4];
  float pf_ratio;
  float dropout;
  float grad_scale;   // 1 / world_size
  int B, L;
  int D[2];
};

// ------------------------------------------------------------------------------------------------ Philox4x32-10
__device__ __forceinline__ uint4 philox4x32(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ uint2 philox_key(const Ctl* ctl) {
  const unsigned long long s = ctl->seed ^ (ctl->stream_id * 0x9E3779B97F4A7C15ull);
  return make_uint2(static_cast<uint32_t>(s), static_cast<uint32_t>(s >> 32));
}

// First statement of every kernel of the step: let the next kernel of the stream start launching, then wait until
// every earlier kernel has completed and flushed (both are no-ops without programmatic stream serialization).
__device__ __forceinline__ void pdl_prologue() {
  grid_dep_launch();
  grid_dep_wait();
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------ step control
// The first kernel of a step (k_gather / k_split_x) derives the step's scalars: every block reads the plan cursor
// (nobody writes it while that kernel runs) and one thread publishes the derived values for the later kernels. The
// cursor and the Adam step count advance later in the step (k_reparam); the injection flag is cleared by k_adam.
__device__ __forceinline__ void step_begin(Ctl* ctl, const float* __restrict__ plan_kl, const StepConsts& sc) {
  const long long row = ctl->cursor;
  const long long t = ctl->adam_t + 1;
  ctl->row = static_cast<int>(row);
  ctl->kl_base = plan_kl[row];
  ctl->kl_coef = sc.w[0] * plan_kl[row];
  const double bc1 = 1.0 - pow(static_cast<double>(sc.beta1), static_cast<double>(t));
  const double bc2 = 1.0 - pow(static_cast<double>(sc.beta2), static_cast<double>(t));
  ctl->step_size = static_cast<float>(static_cast<double>(sc.lr) / bc1);
  ctl->inv_bc2_sqrt = static_cast<float>(1.0 / sqrt(bc2));
  ctl->stream_id = static_cast<unsigned long long>(t);
}

// busy-wait (profiling only): gives the host a head start so a whole step is enqueued behind it
__global__ void k_spin(long long ns) {
  const uint64_t t0 = globaltimer_ns();
  while (globaltimer_ns() - t0 < static_cast<uint64_t>(ns)) {}
}

// ------------------------------------------------------------------------------------------------ gather
// x_i[b, :] = data_i[idx_i[row][b], :]   (jamie/jamie.py:583).  grid (B, 2), 128 threads.
struct GatherArgs {
  const float* data[2];
  long long ld_data[2];
  float* x[2];
  float* xh[2];       // TF32 hi / lo planes of x (operands of the first encoder GEMM), same pitch as x
  float* xl[2];
  int ldx[2];
  int D[2];
  const int* idx[2];  // plan index arrays [nsteps][B]
  const float* stage[2][2];  // host-batch steps: [slot][modality] staging buffers the H2D copies land in (pitch ldx)
};
__global__ void k_gather(GatherArgs a, Ctl* ctl, const float* __restrict__ plan_kl, StepConsts sc, int B) {
  pdl_prologue();
  const int i = blockIdx.y, b = blockIdx.x;
  const long long row = ctl->cursor;
  if (i == 0 && b == 0 && threadIdx.x == 0) step_begin(ctl, plan_kl, sc);
  const int src = a.idx[i][row * B + b];
  const float* s = a.data[i] + static_cast<long long>(src) * a.ld_data[i];
  const long long o = static_cast<long long>(b) * a.ldx[i];
  float* d = a.x[i] + o;
  float* dh = a.xh[i] + o;
  float* dl = a.xl[i] + o;
  const int D = a.D[i];
  int j0 = 0;
  if ((a.ld_data[i] & 3) == 0 && (reinterpret_cast<uintptr_t>(a.data[i]) & 15) == 0) {
    const int nv = D >> 2;
    for (int j = threadIdx.x; j < nv; j += blockDim.x) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(s) + j);
      float4 h, l;
      tf32_split(v.x, h.x, l.x); tf32_split(v.y, h.y, l.y); tf32_split(v.z, h.z, l.z); tf32_split(v.w, h.w, l.w);
      reinterpret_cast<float4*>(d)[j] = v;
      reinterpret_cast<float4*>(dh)[j] = h;
      reinterpret_cast<float4*>(dl)[j] = l;
    }
    j0 = nv << 2;
  }
  for (int j = j0 + threadIdx.x; j < D; j += blockDim.x) {
    const float v = __ldg(s + j);
    d[j] = v;
    tf32_split(v, dh[j], dl[j]);
  }
}
// Host-batch step: the rows arrived by H2D copy in staging slot ctl->host_slot; this writes x and its operand planes.
// grid (B, 2), 128 threads.
__global__ void k_split_x(GatherArgs a, Ctl* ctl, const float* __restrict__ plan_kl, StepConsts sc, int B) {
  pdl_prologue();
  const int i = blockIdx.y, b = blockIdx.x;
  if (i == 0 && b == 0 && threadIdx.x == 0) step_begin(ctl, plan_kl, sc);
  const long long o = static_cast<long long>(b) * a.ldx[i];
  const float* s = a.stage[ctl->host_slot != 0 ? 1 : 0][i] + o;
  float* d = a.x[i] + o;
  float* dh = a.xh[i] + o;
  float* dl = a.xl[i] + o;
  for (int j = threadIdx.x; j < a.D[i]; j += blockDim.x) {
    const float v = s[j];
    d[j] = v;
    tf32_split(v, dh[j], dl[j]);
  }
  (void)B;
}

// ------------------------------------------------------------------------------------------------ P / F blocks
// corr = r * rownorm(P[idx0][:, idx1]) + (1 - r) * rownorm(F[idx0][:, idx1])   (jamie/jamie.py:586-604)
// P is either diag(m) (never materialised) or a dense matrix; F dense or absent.
struct CorrArgs {
  const float* p_diag;    // m[n] or null
  const float* p_dense;   // [n0, n1] or null
  const float* f_dense;   // [n0, n1] or null
  long long n1;
  const int* idx[2];
  float* rs_p;            // [B] row sums of the P block
  float* rs_f;            // [B]
  float* corr;            // [B, B]
  float* corr_t;          // [B, B] transposed copy
  float* fblk;            // [B, B] normalised F block (F loss)
  float* fblk_t;
  float pf_ratio;
};
__device__ __forceinline__ float corr_p_entry(const CorrArgs& a, int i0, int i1) {
  if (a.p_dense) return __ldg(a.p_dense + static_cast<long long>(i0) * a.n1 + i1);
  if (a.p_diag) return i0 == i1 ? __ldg(a.p_diag + i0) : 0.f;
  return 0.f;
}
// one block per block-row a: row sums (fixed-order tree)
__global__ void k_corr_rowsum(CorrArgs a, const Ctl* __restrict__ ctl, int B) {
  pdl_prologue();
  __shared__ float sp[128], sf[128];
  const long long base = static_cast<long long>(ctl->row) * B;
  const int ra = blockIdx.x;
  const int i0 = a.idx[0][base + ra];
  float p = 0.f, f = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const int i1 = a.idx[1][base + b];
    p += corr_p_entry(a, i0, i1);
    if (a.f_dense) f += __ldg(a.f_dense + static_cast<long long>(i0) * a.n1 + i1);
  }
  sp[threadIdx.x] = p; sf[threadIdx.x] = f;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (threadIdx.x < o) { sp[threadIdx.x] += sp[threadIdx.x + o]; sf[threadIdx.x] += sf[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    a.rs_p[ra] = sp[0] == 0.f ? 1.f : sp[0];
    a.rs_f[ra] = sf[0] == 0.f ? 1.f : sf[0];
  }
}
// 32x32 tiles, block (32, 8): writes corr, corr^T, F block and its transpose with coalesced stores.
__global__ void k_corr_build(CorrArgs a, const Ctl* __restrict__ ctl, int B) {
  pdl_prologue();
  __shared__ float tc[32][33], tf[32][33];
  const long long base = static_cast<long long>(ctl->row) * B;
  const int a0 = blockIdx.y * 32, b0 = blockIdx.x * 32;
  const int bb = b0 + threadIdx.x;
  const int i1 = bb < B ? a.idx[1][base + bb] : 0;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int ra = a0 + r;
    float c = 0.f, f = 0.f;
    if (ra < B && bb < B) {
      const int i0 = a.idx[0][base + ra];
      const float pv = corr_p_entry(a, i0, i1) / a.rs_p[ra];
      if (a.f_dense) f = __ldg(a.f_dense + static_cast<long long>(i0) * a.n1 + i1) / a.rs_f[ra];
      c = a.pf_ratio * pv + (1.f - a.pf_ratio) * f;
      a.corr[static_cast<long long>(ra) * B + bb] = c;
      a.fblk[static_cast<long long>(ra) * B + bb] = f;
    }
    tc[r][threadIdx.x] = c;
    tf[r][threadIdx.x] = f;
  }
  __syncthreads();
  const int ca = a0 + threadIdx.x;  // transposed: row index = b, column = a
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int rb = b0 + r;
    if (rb < B && ca < B) {
      a.corr_t[static_cast<long long>(rb) * B + ca] = tc[threadIdx.x][r];
      a.fblk_t[static_cast<long long>(rb) * B + ca] = tf[threadIdx.x][r];
    }
  }
}

// ------------------------------------------------------------------------------------------------ BatchNorm slabs
// Linear output Y [B, N] -> BatchNorm1d (batch statistics) -> LeakyReLU(0.01) -> Dropout(p)
// (jamie/model.py:151-154 and siblings). One block owns 32 feature columns and all B rows, so the batch
// statistics are block-local; 8 warps stride over rows in groups of 4 (one Philox call = 4 rows of one column).
struct BnFwd {
  const float* Y; int ldy;
  float* Hh; float* Hl; int ldh;      // output as TF32 hi / lo planes (the next GEMM's operand; h itself is never needed)
  const float* gamma; const float* beta;
  float* mean; float* invstd;         // saved for backward
  float* run_mean; float* run_var;    // running statistics (momentum 0.1, unbiased variance)
  const unsigned char* mask; int ldm; // injected keep-mask or null
  int N;
  unsigned layer_id;
  int blocks;                         // ceil(N / 32)
};
struct BnFwdPair { BnFwd l[2]; };

__device__ __forceinline__ void block_colsum2(float& a, float& b, float (*sh)[2][32], int warp, int lane) {
  sh[warp][0][lane] = a; sh[warp][1][lane] = b;
  __syncthreads();
  if (warp == 0) {
    float x = 0.f, y = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { x += sh[w][0][lane]; y += sh[w][1][lane]; }
    sh[0][0][lane] = x; sh[0][1][lane] = y;
  }
  __syncthreads();
  a = sh[0][0][lane]; b = sh[0][1][lane];
  __syncthreads();
}

// 4 keep decisions (rows 4*rgroup .. 4*rgroup+3 of column col)
__device__ __forceinline__ uint4 rand4(uint2 key, unsigned layer_id, int col, int rgroup) {
  return philox4x32(make_uint4(static_cast<uint32_t>(rgroup), static_cast<uint32_t>(col), layer_id, 0x4A4Du), key);
}

__global__ void __launch_bounds__(256) k_bn_fwd(BnFwdPair pr, const Ctl* __restrict__ ctl, int B, float p) {
  pdl_prologue();
  __shared__ float sh[8][2][32];
  const int which = blockIdx.x >= pr.l[0].blocks ? 1 : 0;
  const BnFwd& L = pr.l[which];
  const int cb = blockIdx.x - (which ? pr.l[0].blocks : 0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = cb * 32 + lane;
  const bool cok = c < L.N;
  const float* Y = L.Y + (cok ? c : 0);
  // pass 1: mean
  float s = 0.f, dummy = 0.f;
  for (int r = warp; r < B; r += 8) s += cok ? __ldg(Y + static_cast<long long>(r) * L.ldy) : 0.f;
  block_colsum2(s, dummy, sh, warp, lane);
  const float mean = s / static_cast<float>(B);
  // pass 2: biased variance around the mean
  float q = 0.f;
  for (int r = warp; r < B; r += 8) {
    const float d = (cok ? __ldg(Y + static_cast<long long>(r) * L.ldy) : 0.f) - mean;
    q += d * d;
  }
  dummy = 0.f;
  block_colsum2(q, dummy, sh, warp, lane);
  const float var = q / static_cast<float>(B);
  const float invx = 1.0f / sqrtf(var + BN_EPS);
  if (warp == 0 && cok) {
    L.mean[c] = mean;
    L.invstd[c] = invx;
    const float unb = B > 1 ? var * (static_cast<float>(B) / static_cast<float>(B - 1)) : var;
    L.run_mean[c] = (1.f - BN_MOM) * L.run_mean[c] + BN_MOM * mean;
    L.run_var[c] = (1.f - BN_MOM) * L.run_var[c] + BN_MOM * unb;
  }
  // pass 3: normalise, LeakyReLU, dropout
  const float g = cok ? __ldg(L.gamma + c) : 0.f, be = cok ? __ldg(L.beta + c) : 0.f;
  const float scale = p > 0.f ? 1.f / (1.f - p) : 1.f;
  const bool inject = ctl->inject != 0 && L.mask != nullptr;
  const uint32_t thresh = p > 0.f ? static_cast<uint32_t>(fminf(p * 4294967296.0f, 4294967040.0f)) : 0u;
  const uint2 key = philox_key(ctl);
  const int ngroups = (B + 3) >> 2;
  for (int gq = warp; gq < ngroups; gq += 8) {
    uint4 rnd = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    if (p > 0.f && !inject) rnd = rand4(key, L.layer_id, c, gq);
    const uint32_t rr[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = gq * 4 + k;
      if (r < B && cok) {
        const float y = __ldg(Y + static_cast<long long>(r) * L.ldy);
        const float a = g * ((y - mean) * invx) + be;
        float o = a > 0.f ? a : LRELU * a;
        if (p > 0.f) {
          const bool keep = inject ? (L.mask[static_cast<long long>(r) * L.ldm + c] != 0) : (rr[k] >= thresh);
          o = keep ? o * scale : 0.f;
        }
        tf32_split(o, L.Hh[static_cast<long long>(r) * L.ldh + c], L.Hl[static_cast<long long>(r) * L.ldh + c]);
      }
    }
  }
}

// Backward of the slab: dH -> dY (through dropout, LeakyReLU, BatchNorm), dgamma, dbeta; the pre-BN bias gradient is
// identically zero (BN subtracts the batch mean) and is written as 0.
struct BnBwd {
  const float* dH; int lddh;
  const float* Y; int ldy;
  float* dYh; float* dYl; int lddy;   // dY as TF32 hi / lo planes (dgrad reads both, wgrad the hi plane)
  const float* gamma; const float* beta;
  const float* mean; const float* invstd;
  float* dgamma; float* dbeta; float* dbias;
  const unsigned char* mask; int ldm;
  int N;
  unsigned layer_id;
  int blocks;
};
struct BnBwdPair { BnBwd l[2]; };

__global__ void __launch_bounds__(256) k_bn_bwd(BnBwdPair pr, const Ctl* __restrict__ ctl, int B, float p, int accum) {
  pdl_prologue();
  __shared__ float sh[8][2][32];
  const int which = blockIdx.x >= pr.l[0].blocks ? 1 : 0;
  const BnBwd& L = pr.l[which];
  const int cb = blockIdx.x - (which ? pr.l[0].blocks : 0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = cb * 32 + lane;
  const bool cok = c < L.N;
  const int cc = cok ? c : 0;
  const float mean = __ldg(L.mean + cc), inv = __ldg(L.invstd + cc);
  const float g = __ldg(L.gamma + cc), be = __ldg(L.beta + cc);
  const float scale = p > 0.f ? 1.f / (1.f - p) : 1.f;
  const bool inject = ctl->inject != 0 && L.mask != nullptr;
  const uint32_t thresh = p > 0.f ? static_cast<uint32_t>(fminf(p * 4294967296.0f, 4294967040.0f)) : 0u;
  const uint2 key = philox_key(ctl);
  const int ngroups = (B + 3) >> 2;
  float s1 = 0.f, s2 = 0.f;
  for (int pass = 0; pass < 2; ++pass) {
    for (int gq = warp; gq < ngroups; gq += 8) {
      uint4 rnd = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
      if (p > 0.f && !inject) rnd = rand4(key, L.layer_id, c, gq);
      const uint32_t rr[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = gq * 4 + k;
        if (r < B && cok) {
          const float y = __ldg(L.Y + static_cast<long long>(r) * L.ldy + c);
          const float yh = (y - mean) * inv;
          const float a = g * yh + be;
          float d = __ldg(L.dH + static_cast<long long>(r) * L.lddh + c);
          if (p > 0.f) {
            const bool keep = inject ? (L.mask[static_cast<long long>(r) * L.ldm + c] != 0) : (rr[k] >= thresh);
            d = keep ? d * scale : 0.f;
          }
          const float da = a > 0.f ? d : LRELU * d;
          if (pass == 0) {
            s1 += da;
            s2 += da * yh;
          } else {
            const float fb = static_cast<float>(B);
            const long long o = static_cast<long long>(r) * L.lddy + c;
            tf32_split((inv * g / fb) * (fb * da - s1 - yh * s2), L.dYh[o], L.dYl[o]);
          }
        }
      }
    }
    if (pass == 0) {
      block_colsum2(s1, s2, sh, warp, lane);
      if (warp == 0 && cok) {
        if (accum) { L.dbeta[c] += s1; L.dgamma[c] += s2; }
        else { L.dbeta[c] = s1; L.dgamma[c] = s2; L.dbias[c] = 0.f; }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ narrow slabs
// Fast path for B <= 512: a block owns 16 feature columns and all B rows with 1024 threads (64 threads per column,
// 2 row groups of 4 rows each), so a [512, 1024] layer becomes 128 blocks with 8 (forward) / 16 (backward)
// independent 4-byte loads in flight per thread instead of a latency-bound walk down the rows.
constexpr int SLAB_CW = 16;
constexpr int SLAB_THREADS = 1024;
constexpr int SLAB_SLOTS = SLAB_THREADS / SLAB_CW;   // 64 threads per column
constexpr int SLAB_G = 2;                            // row groups (of 4 rows) per thread: B <= 4 * 64 * 2 = 512

// column sums of two per-thread partials over the 64 slots of each column (fixed order), broadcast to all threads.
// CW columns per block (16 or 8), THREADS = 64 CW threads: thread t owns column t % CW, slot t / CW.
template <int CW, int THREADS>
__device__ __forceinline__ void slab_colsum2(float& a, float& b, float (*sh)[2][CW], int warp, int lane) {
#pragma unroll
  for (int o = 16; o >= CW; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if (lane < CW) { sh[warp][0][lane] = a; sh[warp][1][lane] = b; }
  __syncthreads();
  if (warp == 0 && lane < CW) {
    float x = 0.f, y = 0.f;
#pragma unroll
    for (int w = 0; w < THREADS / 32; ++w) { x += sh[w][0][lane]; y += sh[w][1][lane]; }
    sh[0][0][lane] = x; sh[0][1][lane] = y;
  }
  __syncthreads();
  a = sh[0][0][lane & (CW - 1)]; b = sh[0][1][lane & (CW - 1)];
  __syncthreads();
}

template <int CW, int THREADS>
__global__ void __launch_bounds__(THREADS) k_bn_fwd_slab(BnFwdPair pr, const Ctl* __restrict__ ctl, int B, float p) {
  pdl_prologue();
  __shared__ float sh[THREADS / 32][2][CW];
  const int nb0 = (pr.l[0].N + CW - 1) / CW;
  const int which = blockIdx.x >= nb0 ? 1 : 0;
  const BnFwd& L = pr.l[which];
  const int cb = blockIdx.x - (which ? nb0 : 0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = cb * CW + (threadIdx.x & (CW - 1));
  const int slot = threadIdx.x / CW;
  const bool cok = c < L.N;
  const float* Y = L.Y + (cok ? c : 0);
  const int ldy = L.ldy;
  float v[4 * SLAB_G];
#pragma unroll
  for (int t = 0; t < SLAB_G; ++t)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = 4 * (slot + (THREADS / CW) * t) + k;
      v[4 * t + k] = (cok && r < B) ? __ldg(Y + static_cast<long long>(r) * ldy) : 0.f;
    }
  float s = 0.f, dummy = 0.f;
#pragma unroll
  for (int i = 0; i < 4 * SLAB_G; ++i) s += v[i];
  slab_colsum2<CW, THREADS>(s, dummy, sh, warp, lane);
  const float mean = s / static_cast<float>(B);
  float q = 0.f;
#pragma unroll
  for (int t = 0; t < SLAB_G; ++t)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = 4 * (slot + (THREADS / CW) * t) + k;
      const float d = v[4 * t + k] - mean;
      q += (r < B) ? d * d : 0.f;
    }
  dummy = 0.f;
  slab_colsum2<CW, THREADS>(q, dummy, sh, warp, lane);
  const float var = q / static_cast<float>(B);
  const float invx = 1.0f / sqrtf(var + BN_EPS);
  if (slot == 0 && cok) {
    L.mean[c] = mean;
    L.invstd[c] = invx;
    const float unb = B > 1 ? var * (static_cast<float>(B) / static_cast<float>(B - 1)) : var;
    L.run_mean[c] = (1.f - BN_MOM) * L.run_mean[c] + BN_MOM * mean;
    L.run_var[c] = (1.f - BN_MOM) * L.run_var[c] + BN_MOM * unb;
  }
  const float g = cok ? __ldg(L.gamma + c) : 0.f, be = cok ? __ldg(L.beta + c) : 0.f;
  const float scale = p > 0.f ? 1.f / (1.f - p) : 1.f;
  const bool inject = ctl->inject != 0 && L.mask != nullptr;
  const uint32_t thresh = p > 0.f ? static_cast<uint32_t>(fminf(p * 4294967296.0f, 4294967040.0f)) : 0u;
  const uint2 key = philox_key(ctl);
  float* Hh = L.Hh + (cok ? c : 0);
  float* Hl = L.Hl + (cok ? c : 0);
  const int ldh = L.ldh;
#pragma unroll
  for (int t = 0; t < SLAB_G; ++t) {
    const int gq = slot + (THREADS / CW) * t;
    uint4 rnd = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    if (p > 0.f && !inject && 4 * gq < B) rnd = rand4(key, L.layer_id, c, gq);
    const uint32_t rr[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = gq * 4 + k;
      if (r < B && cok) {
        const float a = g * ((v[4 * t + k] - mean) * invx) + be;
        float o = a > 0.f ? a : LRELU * a;
        if (p > 0.f) {
          const bool keep = inject ? (L.mask[static_cast<long long>(r) * L.ldm + c] != 0) : (rr[k] >= thresh);
          o = keep ? o * scale : 0.f;
        }
        tf32_split(o, Hh[static_cast<long long>(r) * ldh], Hl[static_cast<long long>(r) * ldh]);
      }
    }
  }
}

template <int CW, int THREADS>
__global__ void __launch_bounds__(THREADS) k_bn_bwd_slab(BnBwdPair pr, const Ctl* __restrict__ ctl, int B, float p,
                                                               int accum) {
  pdl_prologue();
  __shared__ float sh[THREADS / 32][2][CW];
  const int nb0 = (pr.l[0].N + CW - 1) / CW;
  const int which = blockIdx.x >= nb0 ? 1 : 0;
  const BnBwd& L = pr.l[which];
  const int cb = blockIdx.x - (which ? nb0 : 0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = cb * CW + (threadIdx.x & (CW - 1));
  const int slot = threadIdx.x / CW;
  const bool cok = c < L.N;
  const int cc = cok ? c : 0;
  const float mean = __ldg(L.mean + cc), inv = __ldg(L.invstd + cc);
  const float g = __ldg(L.gamma + cc), be = __ldg(L.beta + cc);
  const float scale = p > 0.f ? 1.f / (1.f - p) : 1.f;
  const bool inject = ctl->inject != 0 && L.mask != nullptr;
  const uint32_t thresh = p > 0.f ? static_cast<uint32_t>(fminf(p * 4294967296.0f, 4294967040.0f)) : 0u;
  const uint2 key = philox_key(ctl);
  float yh[4 * SLAB_G], da[4 * SLAB_G];
#pragma unroll
  for (int t = 0; t < SLAB_G; ++t)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = 4 * (slot + (THREADS / CW) * t) + k;
      const bool ok = cok && r < B;
      yh[4 * t + k] = ok ? __ldg(L.Y + static_cast<long long>(r) * L.ldy + cc) : mean;
      da[4 * t + k] = ok ? __ldg(L.dH + static_cast<long long>(r) * L.lddh + cc) : 0.f;
    }
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int t = 0; t < SLAB_G; ++t) {
    const int gq = slot + (THREADS / CW) * t;
    uint4 rnd = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    if (p > 0.f && !inject && 4 * gq < B) rnd = rand4(key, L.layer_id, c, gq);
    const uint32_t rr[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = gq * 4 + k;
      const float h = (yh[4 * t + k] - mean) * inv;
      const float a = g * h + be;
      float d = da[4 * t + k];
      if (p > 0.f && r < B && cok) {
        const bool keep = inject ? (L.mask[static_cast<long long>(r) * L.ldm + c] != 0) : (rr[k] >= thresh);
        d = keep ? d * scale : 0.f;
      }
      d = a > 0.f ? d : LRELU * d;
      yh[4 * t + k] = h;
      da[4 * t + k] = d;
      s1 += d;
      s2 += d * h;
    }
  }
  slab_colsum2<CW, THREADS>(s1, s2, sh, warp, lane);
  if (slot == 0 && cok) {
    if (accum) { L.dbeta[c] += s1; L.dgamma[c] += s2; }
    else { L.dbeta[c] = s1; L.dgamma[c] = s2; L.dbias[c] = 0.f; }
  }
  const float fb = static_cast<float>(B);
  const float k0 = inv * g / fb;
#pragma unroll
  for (int t = 0; t < SLAB_G; ++t)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = 4 * (slot + (THREADS / CW) * t) + k;
      if (r < B && cok) {
        const long long o = static_cast<long long>(r) * L.lddy + c;
        tf32_split(k0 * (fb * da[4 * t + k] - s1 - yh[4 * t + k] * s2), L.dYh[o], L.dYl[o]);
      }
    }
}

// ------------------------------------------------------------------------------------------------ reconstruction loss
// dxhat = w_rec * 2 (xhat - x) / (B D);  per-block partial of sum (xhat - x)^2;  bias gradient of the last decoder
// Linear = column sums of dxhat   (jamie/jamie.py:637-643).
struct RecArgs {
  const float* xhat; int ldxh;
  const float* x; int ldx;
  float* dxh; float* dxl; int lddx;   // d loss / d xhat as TF32 hi / lo planes
  float* dbias;
  float* part;   // [blocks] partial sums of squares
  int D;
  int blocks;
};
struct RecPair { RecArgs m[2]; };
__global__ void __launch_bounds__(256) k_rec(RecPair pr, int B, float w_rec, int accum) {
  pdl_prologue();
  __shared__ float sh[8][2][32];
  const int which = blockIdx.x >= pr.m[0].blocks ? 1 : 0;
  const RecArgs& A = pr.m[which];
  const int cb = blockIdx.x - (which ? pr.m[0].blocks : 0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = cb * 32 + lane;
  const bool cok = c < A.D;
  const float k = w_rec * 2.f / (static_cast<float>(B) * static_cast<float>(A.D));
  float sq = 0.f, cs = 0.f;
#pragma unroll 8
  for (int r = warp; r < B; r += 8) {
    if (cok) {
      const float d = __ldg(A.xhat + static_cast<long long>(r) * A.ldxh + c) - __ldg(A.x + static_cast<long long>(r) * A.ldx + c);
      sq += d * d;
      const float gx = k * d;
      cs += gx;
      tf32_split(gx, A.dxh[static_cast<long long>(r) * A.lddx + c], A.dxl[static_cast<long long>(r) * A.lddx + c]);
    }
  }
  block_colsum2(sq, cs, sh, warp, lane);
  if (warp == 0) {
    if (cok) A.dbias[c] = accum ? A.dbias[c] + cs : cs;
    const float tot = warp_sum(cok ? sq : 0.f);
    if (lane == 0) A.part[cb] = tot;
  }
}

template <int CW, int THREADS>
__global__ void __launch_bounds__(THREADS) k_rec_slab(RecPair pr, int B, float w_rec, int accum) {
  pdl_prologue();
  __shared__ float sh[THREADS / 32][2][CW];
  const int nb0 = (pr.m[0].D + CW - 1) / CW;
  const int which = blockIdx.x >= nb0 ? 1 : 0;
  const RecArgs& A = pr.m[which];
  const int cb = blockIdx.x - (which ? nb0 : 0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = cb * CW + (threadIdx.x & (CW - 1));
  const int slot = threadIdx.x / CW;
  const bool cok = c < A.D;
  const float kk = w_rec * 2.f / (static_cast<float>(B) * static_cast<float>(A.D));
  float xh[4 * SLAB_G], xx[4 * SLAB_G];
#pragma unroll
  for (int t = 0; t < SLAB_G; ++t)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = 4 * (slot + (THREADS / CW) * t) + k;
      const bool ok = cok && r < B;
      xh[4 * t + k] = ok ? __ldg(A.xhat + static_cast<long long>(r) * A.ldxh + c) : 0.f;
      xx[4 * t + k] = ok ? __ldg(A.x + static_cast<long long>(r) * A.ldx + c) : 0.f;
    }
  float sq = 0.f, cs = 0.f;
#pragma unroll
  for (int t = 0; t < SLAB_G; ++t)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = 4 * (slot + (THREADS / CW) * t) + k;
      const float d = xh[4 * t + k] - xx[4 * t + k];
      sq += d * d;
      const float gx = kk * d;
      cs += gx;
      if (r < B && cok) tf32_split(gx, A.dxh[static_cast<long long>(r) * A.lddx + c], A.dxl[static_cast<long long>(r) * A.lddx + c]);
    }
  slab_colsum2<CW, THREADS>(sq, cs, sh, warp, lane);
  if (warp == 0) {
    if (lane < CW && cok) A.dbias[c] = accum ? A.dbias[c] + cs : cs;
    float t = (lane < CW && cok) ? sq : 0.f;
    t = warp_sum(t);
    if (lane == 0) A.part[cb] = t;
  }
}

// ------------------------------------------------------------------------------------------------ latent stage
struct Latent {
  // per modality
  const float* mulv[2]; int ldmv;   // heads output [B, 2L]: mu | logvar
  float* eps[2];                    // [B, LP]
  const float* inj_eps[2];
  float* z[2]; float* c[2]; float* S[2]; float* g[2];   // [B, LP]
  float* ch[2]; float* cl[2];       // TF32 hi / lo planes of c (first decoder GEMM operand)
  float* dmh[2]; float* dml[2];     // TF32 hi / lo planes of dmulv (heads dgrad / wgrad operand)
  float* den[2]; float* rs[2];      // [B]
  float* r;                         // [B, LP] F-loss residual c0 - F c1
  const float* dc_dec[2];           // [B, LP] decoder dgrad wrt c
  float* dmulv[2];                  // [B, 2L] gradient wrt heads output
  float* rowpart;                   // [2][B][8] per-row partial sums
  const float* corr; const float* corr_t; const float* fblk; const float* fblk_t;   // [B, B]
  const float* sigma;               // 2 parameters
  int LP;
  int f_present;
};
constexpr int LAT_MAXT = 4;  // latent width up to 128

// eps (injected or Philox Box-Muller) and z = mu + (exp(logvar/2) + 1e-7) eps   (jamie/model.py:230-240)
__global__ void k_reparam(Latent a, Ctl* ctl, int B, int L) {
  pdl_prologue();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t == 0) { ctl->cursor += 1; ctl->adam_t += 1; }   // this step's scalars were derived by the first kernel
  if (t >= 2 * B * L) return;
  const int i = t / (B * L), rem = t - i * B * L, b = rem / L, l = rem - b * L;
  float e;
  if (ctl->inject) {
    e = a.inj_eps[i][static_cast<long long>(b) * a.LP + l];
  } else {
    const uint4 r = philox4x32(make_uint4(static_cast<uint32_t>(b), static_cast<uint32_t>(l), 0xE950u + i, 0x4A4Du),
                               philox_key(ctl));
    const float u1 = (static_cast<float>(r.x >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float u2 = (static_cast<float>(r.y >> 8) + 0.5f) * (1.0f / 16777216.0f);
    e = sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
  }
  const float mu = a.mulv[i][static_cast<long long>(b) * a.ldmv + l];
  const float lv = a.mulv[i][static_cast<long long>(b) * a.ldmv + L + l];
  a.eps[i][static_cast<long long>(b) * a.LP + l] = e;
  a.z[i][static_cast<long long>(b) * a.LP + l] = mu + (expf(lv * 0.5f) + 1e-7f) * e;
}

// out[l] (per lane, LAT_MAXT strided) = sum_b M[row, b] * V[b, l], skipping zero entries; also returns the row sum.
// The row is fetched 512 entries at a time (16 independent loads per lane) before the nonzeros are visited in
// ascending column order, so the scan is bandwidth- rather than latency-bound and the sum order is fixed.
__device__ __forceinline__ float row_times(const float* __restrict__ Mrow, const float* __restrict__ V, int B, int LP,
                                           int L, int lane, float (&acc)[LAT_MAXT]) {
#pragma unroll
  for (int t = 0; t < LAT_MAXT; ++t) acc[t] = 0.f;
  float rs = 0.f;
  for (int sup = 0; sup < B; sup += 512) {
    float m[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int b = sup + 32 * i + lane;
      m[i] = b < B ? __ldg(Mrow + b) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      unsigned nz = __ballot_sync(0xffffffffu, m[i] != 0.f);
      while (nz) {
        const int src = __ffs(nz) - 1;
        nz &= nz - 1;
        const float mv = __shfl_sync(0xffffffffu, m[i], src);
        rs += mv;
        const float* v = V + static_cast<long long>(sup + 32 * i + src) * LP;
#pragma unroll
        for (int t = 0; t < LAT_MAXT; ++t) {
          const int l = lane + 32 * t;
          if (l < L) acc[t] += mv * v[l];
        }
      }
    }
  }
  return rs;
}

// combine (jamie/model.py:245-259): c_i = (s_i z_i + s_j C_i z_j) / (s_i + s_j rowsum(C_i)), C_0 = corr, C_1 = corr^T.
// One warp per (modality, row).
// fuse_loss (F absent: the F residual is r = c0, nothing of another row is needed): also emits the row partial sums of
// k_latent_loss, which is then not launched.
__global__ void k_combine(Latent a, int B, int L, int fuse_loss) {
  pdl_prologue();
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= 2 * B) return;
  const int i = w / B, row = w - i * B, j = 1 - i;
  const float si = __ldg(a.sigma + i), sj = __ldg(a.sigma + j);
  const float* Ci = (i == 0 ? a.corr : a.corr_t) + static_cast<long long>(row) * B;
  float acc[LAT_MAXT];
  const float rs = row_times(Ci, a.z[j], B, a.LP, L, lane, acc);
  const float den = si + sj * rs;
  if (lane == 0) { a.den[i][row] = den; a.rs[i][row] = rs; }
  float smu = 0.f, scs = 0.f, sr = 0.f;
#pragma unroll
  for (int t = 0; t < LAT_MAXT; ++t) {
    const int l = lane + 32 * t;
    if (l < L) {
      const long long o = static_cast<long long>(row) * a.LP + l;
      a.S[i][o] = acc[t];
      const float zv = a.z[i][o];
      const float cv = (si * zv + sj * acc[t]) / den;
      a.c[i][o] = cv;
      tf32_split(cv, a.ch[i][o], a.cl[i][o]);
      if (fuse_loss) {
        const float mu = a.mulv[i][static_cast<long long>(row) * a.ldmv + l];
        smu += mu * mu;
        const float d = zv - cv;
        scs += d * d;
        if (i == 0) { a.r[o] = cv; sr += cv * cv; }
      }
    }
  }
  if (fuse_loss) {
    smu = warp_sum(smu); scs = warp_sum(scs); sr = warp_sum(sr);
    if (lane == 0) {
      float* rp = a.rowpart + (static_cast<long long>(i) * B + row) * 8;
      rp[0] = smu; rp[1] = scs; rp[2] = sr;
    }
  }
}

// Row partial sums (rowpart[i][row][k]):
//   0: sum_l mu^2   1: sum_l (z - c)^2   2: sum_l r^2 (i = 0)   3: sum_l g z   4: sum_l g c   5: sum_l g S
// F residual r = c0 - F c1 (jamie/jamie.py:663-665).
__global__ void k_latent_loss(Latent a, int B, int L) {
  pdl_prologue();
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= 2 * B) return;
  const int i = w / B, row = w - i * B;
  float acc[LAT_MAXT];
  if (i == 0 && a.f_present) row_times(a.fblk + static_cast<long long>(row) * B, a.c[1], B, a.LP, L, lane, acc);
  else {
#pragma unroll
    for (int t = 0; t < LAT_MAXT; ++t) acc[t] = 0.f;
  }
  float smu = 0.f, scs = 0.f, sr = 0.f;
#pragma unroll
  for (int t = 0; t < LAT_MAXT; ++t) {
    const int l = lane + 32 * t;
    if (l < L) {
      const long long o = static_cast<long long>(row) * a.LP + l;
      const float mu = a.mulv[i][static_cast<long long>(row) * a.ldmv + l];
      smu += mu * mu;
      const float d = a.z[i][o] - a.c[i][o];
      scs += d * d;
      if (i == 0) {
        const float r = a.c[0][o] - acc[t];
        a.r[o] = r;
        sr += r * r;
      }
    }
  }
  smu = warp_sum(smu); scs = warp_sum(scs); sr = warp_sum(sr);
  if (lane == 0) {
    float* rp = a.rowpart + (static_cast<long long>(i) * B + row) * 8;
    rp[0] = smu; rp[1] = scs; rp[2] = sr;
  }
}

// g_i = d(loss)/dc_i / den_i with d/dc_i = decoder dgrad - k_cos (z_i - c_i) + F term.
__global__ void k_latent_bwd_c(Latent a, int B, int L, float k_cos, float k_f) {
  pdl_prologue();
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= 2 * B) return;
  const int i = w / B, row = w - i * B;
  float acc[LAT_MAXT];
  if (i == 1 && a.f_present) row_times(a.fblk_t + static_cast<long long>(row) * B, a.r, B, a.LP, L, lane, acc);
  else {
#pragma unroll
    for (int t = 0; t < LAT_MAXT; ++t) acc[t] = 0.f;
  }
  const float den = a.den[i][row];
  float p3 = 0.f, p4 = 0.f, p5 = 0.f;
#pragma unroll
  for (int t = 0; t < LAT_MAXT; ++t) {
    const int l = lane + 32 * t;
    if (l < L) {
      const long long o = static_cast<long long>(row) * a.LP + l;
      const float z = a.z[i][o], c = a.c[i][o];
      float dc = a.dc_dec[i][o] - k_cos * (z - c);
      dc += i == 0 ? k_f * a.r[o] : -k_f * acc[t];
      const float g = dc / den;
      a.g[i][o] = g;
      p3 += g * z; p4 += g * c; p5 += g * a.S[i][o];
    }
  }
  p3 = warp_sum(p3); p4 = warp_sum(p4); p5 = warp_sum(p5);
  if (lane == 0) {
    float* rp = a.rowpart + (static_cast<long long>(i) * B + row) * 8;
    rp[3] = p3; rp[4] = p4; rp[5] = p5;
  }
}

// dz_i = k_cos (z_i - c_i) + s_i g_i + s_i C_i g_j ; then through the reparameterisation and the KL term
// (jamie/jamie.py:619-632 with the reference's logvar quirk: only rows 0 and 1 of modality 1's logvar get KL
// gradient, each scaled by the broadcast over the batch).
__global__ void k_latent_bwd_z(Latent a, const Ctl* __restrict__ ctl, int B, int L, float k_cos) {
  pdl_prologue();
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= 2 * B) return;
  const int i = w / B, row = w - i * B, j = 1 - i;
  const float si = __ldg(a.sigma + i);
  const float* Ci = (i == 0 ? a.corr : a.corr_t) + static_cast<long long>(row) * B;
  float acc[LAT_MAXT];
  row_times(Ci, a.g[j], B, a.LP, L, lane, acc);
  const float kkl = ctl->kl_coef;
  const float fbl = static_cast<float>(B) * static_cast<float>(L);
#pragma unroll
  for (int t = 0; t < LAT_MAXT; ++t) {
    const int l = lane + 32 * t;
    if (l < L) {
      const long long o = static_cast<long long>(row) * a.LP + l;
      const long long om = static_cast<long long>(row) * a.ldmv;
      const float mu = a.mulv[i][om + l], lv = a.mulv[i][om + L + l];
      const float dz = k_cos * (a.z[i][o] - a.c[i][o]) + si * a.g[i][o] + si * acc[t];
      float dmu = dz + kkl * mu / fbl;
      float dlv = dz * a.eps[i][o] * 0.5f * expf(lv * 0.5f);
      if (i == 1 && row < 2) dlv += kkl * -0.5f * (1.f - expf(lv)) / static_cast<float>(L);
      a.dmulv[i][om + l] = dmu;
      a.dmulv[i][om + L + l] = dlv;
      tf32_split(dmu, a.dmh[i][om + l], a.dml[i][om + l]);
      tf32_split(dlv, a.dmh[i][om + L + l], a.dml[i][om + L + l]);
    }
  }
}

// One block: loss scalars, d sigma, head bias gradients (fixed-order sums).
struct FinalArgs {
  const float* rowpart;          // [2][B][8]
  const float* rs[2];
  const float* rec_part[2]; int rec_blocks[2];
  const float* mulv1; int ldmv;  // modality 1 heads output (logvar rows 0, 1 for the KL value)
  const float* dmulv[2];
  float* dsigma;                 // 2
  float* dbias_heads[2];         // [2L] each: mu bias | var bias
  float* out_loss;               // [nsteps][8]
  float* grad_tail;              // 8 floats after the flat gradients (all-reduce piggy-back)
  int D[2];
};
// grid 1 + ceil(4L / 16) blocks of 1024 threads: block 0 reduces the loss scalars and d sigma; block 1 + k owns 16 of
// the 4L head-bias columns (both modalities: mu bias | var bias) with 64 row slots per column.
__global__ void __launch_bounds__(SLAB_THREADS) k_latent_final(FinalArgs a, const Ctl* __restrict__ ctl, int B, int L,
                                                               StepConsts sc, int accum) {
  pdl_prologue();
  __shared__ float tot[14];   // [i*7 + k]: k = 0..5 the rowpart sums, k = 6: sum_r (g.c)[r] * rowsum_i[r]
  __shared__ float aux[4];    // [0,1]: sum_l (1 + lv - exp lv) of logvar rows 0/1 (modality 1); [2,3]: sum (xhat - x)^2
  __shared__ float sh[SLAB_THREADS / 32][2][SLAB_CW];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (blockIdx.x > 0) {
    // head bias gradients: column sums of dmulv over the batch (fixed order)
    const int col = (blockIdx.x - 1) * SLAB_CW + (tid & (SLAB_CW - 1));
    const int slot = tid / SLAB_CW;
    const bool cok = col < 4 * L;
    const int i = cok ? col / (2 * L) : 0, cidx = cok ? col - i * 2 * L : 0;
    const float* src = a.dmulv[i] + cidx;
    float s = 0.f, dummy = 0.f;
    if (cok)
#pragma unroll 8
      for (int r = slot; r < B; r += SLAB_SLOTS) s += __ldg(src + static_cast<long long>(r) * a.ldmv);
    slab_colsum2<SLAB_CW, SLAB_THREADS>(s, dummy, sh, warp, lane);
    if (slot == 0 && cok) a.dbias_heads[i][cidx] = accum ? a.dbias_heads[i][cidx] + s : s;
    return;
  }
  if (warp >= 14 && warp < 16) {
    const int i = warp - 14;
    float t1 = 0.f;
    for (int l = lane; l < L; l += 32) {
      const float lv = a.mulv1[static_cast<long long>(i) * a.ldmv + L + l];
      t1 += 1.f + lv - expf(lv);
    }
    t1 = warp_sum(t1);
    if (lane == 0) aux[i] = t1;
  } else if (warp >= 16 && warp < 18) {
    const int i = warp - 16;
    float sacc = 0.f;
    for (int b = lane; b < a.rec_blocks[i]; b += 32) sacc += a.rec_part[i][b];
    sacc = warp_sum(sacc);
    if (lane == 0) aux[2 + i] = sacc;
  }
  // 14 row reductions, one warp each (fixed order: lane-strided partials, then the shuffle tree)
  if (warp < 14) {
    const int i = warp / 7, k = warp % 7;
    float s = 0.f;
    for (int r = lane; r < B; r += 32) {
      const float* rp = a.rowpart + (static_cast<long long>(i) * B + r) * 8;
      s += k < 6 ? rp[k] : rp[4] * a.rs[i][r];
    }
    s = warp_sum(s);
    if (lane == 0) tot[warp] = s;
  }
  __syncthreads();
  if (tid == 0) {
    const float fB = static_cast<float>(B), fL = static_cast<float>(L);
    // KL value (jamie/jamie.py:619-628) with logvars = rows 0/1 of modality 1's logvar
    float kl = 0.f;
    for (int i = 0; i < 2; ++i) kl += -0.5f * (aux[i] / fL - tot[i * 7 + 0] / (fB * fL));
    const float l_kl = ctl->kl_base * kl;
    float rec = 0.f;
    for (int i = 0; i < 2; ++i) rec += aux[2 + i] / (fB * static_cast<float>(a.D[i]));
    const float l_cos = 32.f * (tot[1] + tot[7 + 1]) / (fB * fL);
    const float l_f = tot[2] / (fB * fL);
    // d sigma (combine backward)
    // i = 0: d s0 += sum g0.z0 - sum g0.c0 ; d s1 += sum g0.S0 - sum (g0.c0) rs0     (and symmetrically for i = 1)
    const float ds0 = (tot[3] - tot[4]) + (tot[7 + 5] - tot[7 + 6]);
    const float ds1 = (tot[7 + 3] - tot[7 + 4]) + (tot[5] - tot[6]);
    a.dsigma[0] = accum ? a.dsigma[0] + ds0 : ds0;
    a.dsigma[1] = accum ? a.dsigma[1] + ds1 : ds1;
    const float total = sc.w[0] * l_kl + sc.w[1] * rec + sc.w[2] * l_cos + sc.w[3] * l_f;
    float* o = a.out_loss + static_cast<long long>(ctl->row) * 8;
    o[0] = l_kl; o[1] = rec; o[2] = l_cos; o[3] = l_f; o[4] = total; o[6] = 0.f; o[7] = 0.f;
    a.grad_tail[0] = l_kl; a.grad_tail[1] = rec; a.grad_tail[2] = l_cos; a.grad_tail[3] = l_f; a.grad_tail[4] = total;
  }
}

// ------------------------------------------------------------------------------------------------ clip + Adam
// Phase 1: per-block partial of sum g^2 over the padded flat buffer (padding is zero).
__global__ void __launch_bounds__(256) k_gradnorm(const float* __restrict__ g, long long n4, double* __restrict__ part) {
  pdl_prologue();
  __shared__ double red[256];
  double s = 0.0;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * 256) {
    const float4 v = __ldg(g4 + i);
    s += static_cast<double>(v.x) * v.x + static_cast<double>(v.y) * v.y + static_cast<double>(v.z) * v.z +
         static_cast<double>(v.w) * v.w;
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = red[0];
}
// Phase 2: every block re-reduces the partials in the same order (identical clip coefficient everywhere), then
// g *= grad_scale * clip;  m, v, theta updated with torch.optim.Adam's formulas (jamie/jamie.py:739-741).
__global__ void __launch_bounds__(256, 6) k_adam(float* __restrict__ theta, float* __restrict__ theta_hi,
                                              float* __restrict__ theta_lo, const float* __restrict__ g, float* __restrict__ m,
                                              float* __restrict__ v, long long n4, const double* __restrict__ part,
                                              int nparts, Ctl* ctl, StepConsts sc,
                                              float* __restrict__ out_loss) {
  pdl_prologue();
  __shared__ double red[256];
  __shared__ float s_coef;
  double s = 0.0;
  for (int i = threadIdx.x; i < nparts; i += 256) s += part[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double norm = sqrt(red[0]) * static_cast<double>(sc.grad_scale);
    const double coef = fmin(1.0, static_cast<double>(sc.max_norm) / (norm + 1e-6));
    s_coef = static_cast<float>(coef) * sc.grad_scale;
    if (blockIdx.x == 0) {
      out_loss[static_cast<long long>(ctl->row) * 8 + 5] = static_cast<float>(norm);
      ctl->inject = 0;   // injected eps / masks serve exactly one step
    }
  }
  __syncthreads();
  const float coef = s_coef;
  const float b1 = sc.beta1, b2 = sc.beta2, eps = sc.adam_eps;
  const float step = ctl->step_size, ibc2 = ctl->inv_bc2_sqrt;
  float4* t4 = reinterpret_cast<float4*>(theta);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * 256) {
    const float4 gg = __ldg(g4 + i);
    float4 mm = m4[i], vv = v4[i], tt = t4[i];
    const float gx[4] = {gg.x * coef, gg.y * coef, gg.z * coef, gg.w * coef};
    float* mp = reinterpret_cast<float*>(&mm);
    float* vp = reinterpret_cast<float*>(&vv);
    float* tp = reinterpret_cast<float*>(&tt);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      mp[k] = mp[k] + (gx[k] - mp[k]) * (1.f - b1);
      vp[k] = vp[k] * b2 + gx[k] * gx[k] * (1.f - b2);
      const float denom = sqrtf(vp[k]) * ibc2 + eps;
      tp[k] = tp[k] - step * (mp[k] / denom);
    }
    m4[i] = mm; v4[i] = vv; t4[i] = tt;
    float4 th, tl;   // the updated weights as GEMM operand planes for the next step
    tf32_split(tt.x, th.x, tl.x); tf32_split(tt.y, th.y, tl.y); tf32_split(tt.z, th.z, tl.z); tf32_split(tt.w, th.w, tl.w);
    reinterpret_cast<float4*>(theta_hi)[i] = th;
    reinterpret_cast<float4*>(theta_lo)[i] = tl;
  }
}
// theta -> (theta_hi, theta_lo) over the whole flat buffer (after jb_set_params)
__global__ void k_split_flat(const float* __restrict__ src, float* __restrict__ hi, float* __restrict__ lo, long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x)
    tf32_split(src[i], hi[i], lo[i]);
}

// ------------------------------------------------------------------------------------------------ eval-mode folding
// W'[r, :] = W[r, :] * s[r],  b'[r] = (b[r] - running_mean[r]) * s[r] + beta[r],  s = gamma / sqrt(running_var + eps)
struct FoldArgs {
  const float* W; const float* b; const float* gamma; const float* beta; const float* rm; const float* rv;
  float* Wf; float* bf;
  int rows, cols, ld;
};
__global__ void k_fold(FoldArgs a) {
  const int r = blockIdx.x;
  const float s = a.gamma[r] / sqrtf(a.rv[r] + BN_EPS);
  for (int c = threadIdx.x; c < a.cols; c += blockDim.x)
    a.Wf[static_cast<long long>(r) * a.ld + c] = a.W[static_cast<long long>(r) * a.ld + c] * s;
  if (threadIdx.x == 0) a.bf[r] = (a.b[r] - a.rm[r]) * s + a.beta[r];
}

// strided 2-D copy (pads rows to a TMA-friendly pitch)
__global__ void k_copy2d(const float* __restrict__ src, long long lds, float* __restrict__ dst, long long ldd, long long rows,
                         int cols) {
  const long long r = blockIdx.x;
  if (r >= rows) return;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) dst[r * ldd + c] = src[r * lds + c];
}

// ------------------------------------------------------------------------------------------------ PCA helpers
// Error-compensated TF32 split (tf32_split) of a [rows, cols] matrix with a fused pre-transform; the split GEMM then
// reproduces the fp32 product to ~2^-22 relative (used for the PCA projection, which the reference computes in
// float64 before casting to float32).
//   mode 0: v = x - colmean[c]      (centre before projecting, jamie/utilities.py:663)
//   mode 1: v = x * s + m           (un-standardise before the inverse projection, jamie/utilities.py:674-675)
//   mode 2: v = x
__global__ void k_split_tf32(const float* __restrict__ src, long long lds, long long rows, int cols, int mode,
                             const float* __restrict__ colmean, float s, float m, float* __restrict__ hi,
                             float* __restrict__ lo, long long ldd) {
  const long long r = blockIdx.x;
  if (r >= rows) return;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    float v = src[r * lds + c];
    if (mode == 0) v -= colmean[c];
    else if (mode == 1) v = v * s + m;
    tf32_split(v, hi[r * ldd + c], lo[r * ldd + c]);
  }
}
// out = (z - m) / s with NaN -> 0 (jamie/utilities.py:664-669), written with the caller's pitch
__global__ void k_standardise(const float* __restrict__ z, long long ldz, long long rows, int cols, float m, float s,
                              float* __restrict__ out, long long ldo) {
  const long long r = blockIdx.x;
  if (r >= rows) return;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    float v = (z[r * ldz + c] - m) / s;
    if (v != v) v = 0.f;
    out[r * ldo + c] = v;
  }
}

}  // namespace jb
