// The whole JAMIE training step as ONE persistent cooperative kernel (round 2).
//
// Round 1 ran the step as a CUDA graph of 31 kernels, 29 of them on the critical path; its own trace showed that ~8 of
// every 12-14 us GEMM link were launch latency, prologue (barrier init, TMEM allocation, tensor-map fetch), pipeline fill
// and epilogue, and that the BatchNorm / latent / optimizer kernels were latency chains of a few microseconds each
// (profiles/README.md, round 1). Here one CTA per SM (148 x 512 threads) stays resident for any number of optimizer
// steps and walks the phases of the step, separated by grid barriers (ptx.cuh: grid_barrier, ~1.2 us): the operand ring,
// the mbarriers, the TMEM allocation and the tensor maps live for the whole kernel, the GEMM phases run the fp16-split
// tcgen05 pipeline of hgemm.cuh (4 B per operand element instead of round 1's 8), and the element-wise phases use all
// 16 warps of all SMs.
//
// Phases of one step (StepPhase). "slab" = a CTA owns 16 feature columns and all B rows, so BatchNorm statistics are
// CTA-local; "row" = one warp per (modality, batch row). Split-K partial sums of a GEMM are reduced by the phase that
// consumes them (fixed order: deterministic, no atomics anywhere).
//   GATHER   rows: x = data[idx] + fp16 operand planes; CTA strips: P / F blocks (row sums + normalise + transposes)
//   ENC1     GEMM  y1 = x W1^T + b1                        BN1  slab: batch stats, LeakyReLU, dropout -> h1 planes
//   ENC2     GEMM  y2 = h1 W2^T + b2 (split-K)             BN2  slab -> h2 planes
//   HEADS    GEMM  [mu | logvar] = h2 Wmv^T + bmv (split-K)
//   REPARAM  elements: eps (Philox or injected), z         COMBINE rows: sigma-weighted combine over the nonzeros of
//            the correspondence rows (+ latent loss partials when F is absent)     LATLOSS rows (only with F)
//   DEC1..3  GEMMs + BN3, BN4 slabs                         REC  slab: reconstruction loss, d xhat planes, bias grad
//   DG5, DG4, DG3 dgrad GEMMs + BNB4, BNB3 backward slabs   LATBC, LATBZ rows: latent backward (KL with the reference's
//            logvar quirk, "CosSim", F loss, combine, reparameterisation) -> d[mu | logvar] planes
//   DGH      dgrad GEMM of the heads; idle CTAs: loss scalars, d sigma, head bias gradients (FINAL)
//   BNB2, DG2, BNB1                                         WGRAD all 12 weight gradients (3-pass, 128 x 256 tiles)
//   NORM     sum g^2 partials                               ADAM  clip + Adam over the flat buffer + fp16 weight planes
// The backward pass carries the loss scale `gs` (a power of two): d xhat and the latent loss gradients are multiplied by
// it, every write into the gradient buffer divides it out (exactly), so small gradients stay in fp16's normal range.
#pragma once
#include "hgemm.cuh"
#include "kernels.cuh"

namespace jb {

constexpr int SK_THREADS = 512;
constexpr int SK_WARPS = SK_THREADS / 32;
constexpr int SK_CW = 16;                          // slab: columns per CTA item
constexpr int SK_SLOTS = SK_THREADS / SK_CW;       // 32 row slots per column
constexpr int SK_RG = 4;                           // register path: row groups (4 rows each) per thread, B <= 512
constexpr int SK_MAX_CTAS = 160;

enum StepPhase : int {
  PH_GATHER = 0, PH_ENC1, PH_BN1, PH_ENC2, PH_BN2, PH_HEADS, PH_REPARAM, PH_COMBINE, PH_LATLOSS, PH_DEC1, PH_BN3, PH_DEC2,
  PH_BN4, PH_DEC3, PH_REC, PH_DG5, PH_BNB4, PH_DG4, PH_BNB3, PH_DG3, PH_LATBC, PH_LATBZ, PH_DGH, PH_BNB2, PH_DG2, PH_BNB1,
  PH_WGRAD, PH_NORM, PH_ADAM, PH_COUNT
};
// index of a GEMM phase in StepCtx::gph, or -1
__host__ __device__ inline int gemm_index(int ph) {
  switch (ph) {
    case PH_ENC1: return 0; case PH_ENC2: return 1; case PH_HEADS: return 2; case PH_DEC1: return 3; case PH_DEC2: return 4;
    case PH_DEC3: return 5; case PH_DG5: return 6; case PH_DG4: return 7; case PH_DG3: return 8; case PH_DGH: return 9;
    case PH_DG2: return 10; case PH_WGRAD: return 11; default: return -1;
  }
}
constexpr int SK_NUM_GEMM = 12;

struct Parts {       // a GEMM output as split-K partial sums: element e of partial p at ptr[p * stride + e]
  float* ptr; long long stride; int n;
};
__device__ __forceinline__ float ld_parts(const Parts& q, long long e) {
  float s = __ldcg(q.ptr + e);
  for (int p = 1; p < q.n; ++p) s += __ldcg(q.ptr + p * q.stride + e);
  return s;
}

struct BnLayer {     // one BatchNorm(+LeakyReLU+Dropout) layer of one modality, forward and backward views; one pitch
  Parts Y;                         // pre-BN Linear output (reduced in place into partial 0 by the forward slab)
  __half *Hh, *Hl;                 // post-dropout activation, operand planes
  const float *gamma, *beta;
  float *mean, *invstd, *run_mean, *run_var;
  const unsigned char* mask;       // injected keep mask [B, N] or null
  Parts dH;                        // gradient wrt the layer output (dgrad result)
  __half *dYh, *dYl;               // gradient wrt the pre-BN output, operand planes
  float *dgamma, *dbeta, *dbias;
  int N, ld;
  unsigned layer_id;
};

struct ModCtx {      // per modality
  const float* data; long long ld_data;
  const float* stage[2];           // host-batch steps: the two H2D landing buffers
  const int* idx;                  // plan [nsteps][B]
  float* x; __half *xh, *xl;       // gathered rows [B, ldD] + planes
  Parts mulv;                      // heads output [B, ldmv]: mu | logvar
  float *eps, *z, *c, *S, *g, *den, *rs;   // latent [B, LP] / [B]
  const float* inj_eps;
  __half *ch, *cl;                 // planes of c
  Parts dc;                        // decoder dgrad wrt c [B, LP]
  float* dmulv; __half *dmh, *dml; // gradient wrt the heads output + planes
  Parts xhat;                      // reconstruction [B, ldD]
  __half *dxh, *dxl;               // d loss / d xhat planes
  float* db5;                      // bias gradient of the last decoder Linear
  float* rec_part;                 // [ceil(D / 16)] partial sums of squares
  float* dbias_heads;              // [2L]
  int D, ldD;
};

struct StepCtx {
  int B, L, LP, ldmv;
  ModCtx m[2];
  BnLayer bn[4][2];                // enc1, enc2, dec1, dec2
  // correspondence blocks
  const float *p_diag, *p_dense, *f_dense; long long pn1;
  float *corr, *corr_t, *fblk, *fblk_t;
  float pf_ratio;
  int f_present;
  // latent scratch
  float* lat_r; float* rowpart;
  // parameters / optimizer
  float *theta, *grad, *adam_m, *adam_v; __half *theta_hi, *theta_lo;
  long long n_flat;
  const float* sigma; float* dsigma;
  double* norm_part;               // [SK_MAX_CTAS]
  // plan / control / outputs
  const float* plan_kl; float* out_loss; Ctl* ctl;
  StepConsts sc;
  float gs, inv_gs;                // loss scale of the backward pass and its inverse
  // GEMM tables
  const HgProblem* probs;
  HgPhase gph[SK_NUM_GEMM];
};

struct StepVars {    // per-step scalars (one copy per CTA in shared memory)
  long long row;
  float kl_base, kl_coef, step_size, inv_bc2_sqrt;
  uint2 key;
  int inject, accum, host_slot;
};

// ------------------------------------------------------------------------------------------------ helpers
// column sums of two per-thread partials over the 32 row slots of each of the 16 slab columns (fixed order), broadcast.
// Thread t owns column t & 15, slot t >> 4. Must be called by all 512 threads.
__device__ __forceinline__ void sk_colsum2(float& a, float& b, float* sh, int warp, int lane) {
  a += __shfl_xor_sync(0xffffffffu, a, 16);
  b += __shfl_xor_sync(0xffffffffu, b, 16);
  if (lane < 16) { sh[(2 * warp) * 16 + lane] = a; sh[(2 * warp + 1) * 16 + lane] = b; }
  __syncthreads();
  if (warp == 0 && lane < 16) {
    float x = 0.f, y = 0.f;
#pragma unroll
    for (int w = 0; w < SK_WARPS; ++w) { x += sh[(2 * w) * 16 + lane]; y += sh[(2 * w + 1) * 16 + lane]; }
    sh[lane] = x; sh[16 + lane] = y;
  }
  __syncthreads();
  a = sh[lane & 15]; b = sh[16 + (lane & 15)];
  __syncthreads();
}
__device__ __forceinline__ void st_h4(__half* p, float a, float b, float c, float d, bool lo_of = false) { (void)lo_of;
  const __half2 u = __floats2half2_rn(a, b), v = __floats2half2_rn(c, d);
  uint2 w;
  w.x = *reinterpret_cast<const uint32_t*>(&u); w.y = *reinterpret_cast<const uint32_t*>(&v);
  *reinterpret_cast<uint2*>(p) = w;
}
// split four values into the hi / lo planes with two 8-byte stores
__device__ __forceinline__ void split4_store(__half* hi, __half* lo, float a, float b, float c, float d) {
  const __half ha = __float2half_rn(a), hb = __float2half_rn(b), hc = __float2half_rn(c), hd = __float2half_rn(d);
  const __half2 h01 = __halves2half2(ha, hb), h23 = __halves2half2(hc, hd);
  const __half2 l01 = __floats2half2_rn((a - __half2float(ha)) * HG_LO_SCALE, (b - __half2float(hb)) * HG_LO_SCALE);
  const __half2 l23 = __floats2half2_rn((c - __half2float(hc)) * HG_LO_SCALE, (d - __half2float(hd)) * HG_LO_SCALE);
  uint2 wh, wl;
  wh.x = *reinterpret_cast<const uint32_t*>(&h01); wh.y = *reinterpret_cast<const uint32_t*>(&h23);
  wl.x = *reinterpret_cast<const uint32_t*>(&l01); wl.y = *reinterpret_cast<const uint32_t*>(&l23);
  *reinterpret_cast<uint2*>(hi) = wh;
  *reinterpret_cast<uint2*>(lo) = wl;
}

// ------------------------------------------------------------------------------------------------ GATHER
__device__ __forceinline__ float sk_p_entry(const StepCtx& cx, int i0, int i1) {
  if (cx.p_dense) return __ldg(cx.p_dense + static_cast<long long>(i0) * cx.pn1 + i1);
  if (cx.p_diag) return i0 == i1 ? __ldg(cx.p_diag + i0) : 0.f;
  return 0.f;
}
// P / F blocks of the step (jamie/jamie.py:586-604): one CTA per strip of 32 block rows: row sums, then 32 x 32 tiles.
__device__ void sk_corr_strip(const StepCtx& cx, const StepVars& sv, int strip, float* scratch, float* tiles, int warp, int lane) {
  const int B = cx.B;
  const long long base = sv.row * B;
  const int a0 = strip * 32;
  float* rs_p = scratch;        // [32]
  float* rs_f = scratch + 32;   // [32]
  for (int rr = warp; rr < 32; rr += SK_WARPS) {
    const int ra = a0 + rr;
    float p = 0.f, f = 0.f;
    if (ra < B) {
      const int i0 = cx.m[0].idx[base + ra];
      for (int b = lane; b < B; b += 32) {
        const int i1 = cx.m[1].idx[base + b];
        p += sk_p_entry(cx, i0, i1);
        if (cx.f_dense) f += __ldg(cx.f_dense + static_cast<long long>(i0) * cx.pn1 + i1);
      }
    }
    p = warp_sum(p); f = warp_sum(f);
    if (lane == 0) { rs_p[rr] = p == 0.f ? 1.f : p; rs_f[rr] = f == 0.f ? 1.f : f; }
  }
  __syncthreads();
  float* tc = tiles + warp * (2 * 32 * 33);
  float* tf = tc + 32 * 33;
  for (int b0 = warp * 32; b0 < B; b0 += SK_WARPS * 32) {
    const int bb = b0 + lane;
    const int i1 = bb < B ? cx.m[1].idx[base + bb] : 0;
    for (int r = 0; r < 32; ++r) {
      const int ra = a0 + r;
      float c = 0.f, f = 0.f;
      if (ra < B && bb < B) {
        const int i0 = cx.m[0].idx[base + ra];
        const float pv = sk_p_entry(cx, i0, i1) / rs_p[r];
        if (cx.f_dense) f = __ldg(cx.f_dense + static_cast<long long>(i0) * cx.pn1 + i1) / rs_f[r];
        c = cx.pf_ratio * pv + (1.f - cx.pf_ratio) * f;
        cx.corr[static_cast<long long>(ra) * B + bb] = c;
        cx.fblk[static_cast<long long>(ra) * B + bb] = f;
      }
      tc[r * 33 + lane] = c;
      tf[r * 33 + lane] = f;
    }
    __syncwarp();
    const int ca = a0 + lane;   // transposed: row index = b, column = a
    for (int r = 0; r < 32; ++r) {
      const int rb = b0 + r;
      if (rb < B && ca < B) {
        cx.corr_t[static_cast<long long>(rb) * B + ca] = tc[lane * 33 + r];
        cx.fblk_t[static_cast<long long>(rb) * B + ca] = tf[lane * 33 + r];
      }
    }
    __syncwarp();
  }
  fence_proxy_async_smem();   // the tile scratch is operand-ring memory: generic writes before the next TMA writes
  __syncthreads();
}
// x_i[b, :] = data_i[idx_i[row][b], :] (jamie/jamie.py:583) + operand planes; one warp per (modality, row)
__device__ void sk_gather_rows(const StepCtx& cx, const StepVars& sv, int use_stage, int gw, int nw, int lane) {
  const int B = cx.B;
  for (int it = gw; it < 2 * B; it += nw) {
    const int i = it / B, b = it - i * B;
    const ModCtx& M = cx.m[i];
    const float* s;
    if (use_stage) s = M.stage[sv.host_slot != 0 ? 1 : 0] + static_cast<long long>(b) * M.ldD;
    else s = M.data + static_cast<long long>(M.idx[sv.row * B + b]) * M.ld_data;
    const long long o = static_cast<long long>(b) * M.ldD;
    float* d = M.x + o;
    __half* dh = M.xh + o;
    __half* dl = M.xl + o;
    const int D = M.D;
    int j0 = 0;
    if ((reinterpret_cast<uintptr_t>(s) & 15) == 0) {
      const int nv = D >> 2;
      for (int j = lane; j < nv; j += 32) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(s) + j);
        reinterpret_cast<float4*>(d)[j] = v;
        split4_store(dh + 4 * j, dl + 4 * j, v.x, v.y, v.z, v.w);
      }
      j0 = nv << 2;
    }
    for (int j = j0 + lane; j < D; j += 32) {
      const float v = __ldg(s + j);
      d[j] = v;
      h_split(v, dh[j], dl[j]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ BatchNorm slabs
// Linear output Y [B, N] -> BatchNorm1d (batch statistics) -> LeakyReLU(0.01) -> Dropout(p)  (jamie/model.py:151-154 and
// siblings). A CTA item owns 16 feature columns and all B rows: thread t has column t & 15 and row slot t >> 4; rows in
// groups of 4 (one Philox call = the keep decisions of 4 rows of one column). REG: B <= 512, the column stays in
// registers between the statistics and the normalisation; otherwise three passes over L2.
// apply the BatchNorm affine + LeakyReLU + dropout to one element and write its operand planes
struct BnFwdApply {
  float mean, invx, g, be, scale, p;
  uint32_t thresh;
  bool inject, reduce_y;
};
__device__ __forceinline__ void sk_bn_fwd_elem(const BnLayer& L, const BnFwdApply& A, int r, int c, int ld, float y, uint32_t rk) {
  const long long o = static_cast<long long>(r) * ld + c;
  if (A.reduce_y) L.Y.ptr[o] = y;   // the backward slab reads one array
  const float a = A.g * ((y - A.mean) * A.invx) + A.be;
  float out = a > 0.f ? a : LRELU * a;
  if (A.p > 0.f) {
    const bool keep = A.inject ? (L.mask[static_cast<long long>(r) * L.N + c] != 0) : (rk >= A.thresh);
    out = keep ? out * A.scale : 0.f;
  }
  h_split(out, L.Hh[o], L.Hl[o]);
}
template <bool REG>
__device__ void sk_bn_fwd_item(const BnLayer& L, const StepCtx& cx, const StepVars& sv, int cb, float* sh, int tid, int warp, int lane) {
  const int B = cx.B;
  const float p = cx.sc.dropout;
  const int c = cb * SK_CW + (tid & (SK_CW - 1));
  const int slot = tid / SK_CW;
  const bool cok = c < L.N;
  const int cc = cok ? c : 0;
  const int ld = L.ld;
  const int ngroups = (B + 3) >> 2;
  float v[REG ? 4 * SK_RG : 1];
  float s = 0.f, dummy = 0.f;
  if constexpr (REG) {
#pragma unroll
    for (int t = 0; t < SK_RG; ++t)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = 4 * (slot + SK_SLOTS * t) + k;
        v[4 * t + k] = (cok && r < B) ? ld_parts(L.Y, static_cast<long long>(r) * ld + cc) : 0.f;
        s += v[4 * t + k];
      }
  } else {
    for (int r = slot; r < B; r += SK_SLOTS) s += cok ? ld_parts(L.Y, static_cast<long long>(r) * ld + cc) : 0.f;
  }
  sk_colsum2(s, dummy, sh, warp, lane);
  const float mean = s / static_cast<float>(B);
  float q = 0.f;
  if constexpr (REG) {
#pragma unroll
    for (int t = 0; t < SK_RG; ++t)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = 4 * (slot + SK_SLOTS * t) + k;
        const float d = v[4 * t + k] - mean;
        q += (r < B) ? d * d : 0.f;
      }
  } else {
    for (int r = slot; r < B; r += SK_SLOTS) {
      const float d = (cok ? ld_parts(L.Y, static_cast<long long>(r) * ld + cc) : mean) - mean;
      q += d * d;
    }
  }
  dummy = 0.f;
  sk_colsum2(q, dummy, sh, warp, lane);
  const float var = q / static_cast<float>(B);
  const float invx = 1.0f / sqrtf(var + BN_EPS);
  if (slot == 0 && cok) {
    L.mean[c] = mean;
    L.invstd[c] = invx;
    const float unb = B > 1 ? var * (static_cast<float>(B) / static_cast<float>(B - 1)) : var;
    L.run_mean[c] = (1.f - BN_MOM) * L.run_mean[c] + BN_MOM * mean;
    L.run_var[c] = (1.f - BN_MOM) * L.run_var[c] + BN_MOM * unb;
  }
  BnFwdApply A;
  A.mean = mean; A.invx = invx;
  A.g = cok ? __ldg(L.gamma + c) : 0.f; A.be = cok ? __ldg(L.beta + c) : 0.f;
  A.p = p; A.scale = p > 0.f ? 1.f / (1.f - p) : 1.f;
  A.inject = sv.inject != 0 && L.mask != nullptr;
  A.thresh = p > 0.f ? static_cast<uint32_t>(fminf(p * 4294967296.0f, 4294967040.0f)) : 0u;
  A.reduce_y = L.Y.n > 1;
  if constexpr (REG) {
#pragma unroll
    for (int t = 0; t < SK_RG; ++t) {
      const int gq = slot + SK_SLOTS * t;
      uint4 rnd = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
      if (p > 0.f && !A.inject && 4 * gq < B) rnd = rand4(sv.key, L.layer_id, c, gq);
      const uint32_t rr[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = gq * 4 + k;
        if (r < B && cok) sk_bn_fwd_elem(L, A, r, c, ld, v[4 * t + k], rr[k]);
      }
    }
  } else {
    for (int gq = slot; gq < ngroups; gq += SK_SLOTS) {
      uint4 rnd = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
      if (p > 0.f && !A.inject) rnd = rand4(sv.key, L.layer_id, c, gq);
      const uint32_t rr[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = gq * 4 + k;
        if (r < B && cok) sk_bn_fwd_elem(L, A, r, c, ld, ld_parts(L.Y, static_cast<long long>(r) * ld + c), rr[k]);
      }
    }
  }
}

// Backward of the slab: dH -> dY (through dropout, LeakyReLU, BatchNorm), dgamma, dbeta; the pre-BN bias gradient is
// identically zero (BN subtracts the batch mean) and is written as 0.
struct BnBwdApply {
  float mean, inv, g, be, scale, p;
  uint32_t thresh;
  bool inject;
};
// da through dropout and LeakyReLU'; x_hat
__device__ __forceinline__ void sk_bn_bwd_elem(const BnLayer& L, const BnBwdApply& A, int r, int c, bool ok, float y, float d, uint32_t rk,
                                               float& h_out, float& d_out) {
  const float h = (y - A.mean) * A.inv;
  const float a = A.g * h + A.be;
  if (A.p > 0.f && ok) {
    const bool keep = A.inject ? (L.mask[static_cast<long long>(r) * L.N + c] != 0) : (rk >= A.thresh);
    d = keep ? d * A.scale : 0.f;
  }
  d = a > 0.f ? d : LRELU * d;
  h_out = h; d_out = d;
}
template <bool REG>
__device__ void sk_bn_bwd_item(const BnLayer& L, const StepCtx& cx, const StepVars& sv, int cb, float* sh, int tid, int warp, int lane) {
  const int B = cx.B;
  const float p = cx.sc.dropout;
  const int c = cb * SK_CW + (tid & (SK_CW - 1));
  const int slot = tid / SK_CW;
  const bool cok = c < L.N;
  const int cc = cok ? c : 0;
  const int ld = L.ld;
  const int ngroups = (B + 3) >> 2;
  BnBwdApply A;
  A.mean = __ldcg(L.mean + cc); A.inv = __ldcg(L.invstd + cc);
  A.g = __ldg(L.gamma + cc); A.be = __ldg(L.beta + cc);
  A.p = p; A.scale = p > 0.f ? 1.f / (1.f - p) : 1.f;
  A.inject = sv.inject != 0 && L.mask != nullptr;
  A.thresh = p > 0.f ? static_cast<uint32_t>(fminf(p * 4294967296.0f, 4294967040.0f)) : 0u;
  float yh[REG ? 4 * SK_RG : 1], da[REG ? 4 * SK_RG : 1];
  float s1 = 0.f, s2 = 0.f;
  if constexpr (REG) {
#pragma unroll
    for (int t = 0; t < SK_RG; ++t)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = 4 * (slot + SK_SLOTS * t) + k;
        const bool ok = cok && r < B;
        const long long o = static_cast<long long>(r) * ld + cc;
        yh[4 * t + k] = ok ? __ldcg(L.Y.ptr + o) : A.mean;
        da[4 * t + k] = ok ? ld_parts(L.dH, o) : 0.f;
      }
#pragma unroll
    for (int t = 0; t < SK_RG; ++t) {
      const int gq = slot + SK_SLOTS * t;
      uint4 rnd = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
      if (p > 0.f && !A.inject && 4 * gq < B) rnd = rand4(sv.key, L.layer_id, c, gq);
      const uint32_t rr[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = gq * 4 + k;
        float h, dd;
        sk_bn_bwd_elem(L, A, r, c, cok && r < B, yh[4 * t + k], da[4 * t + k], rr[k], h, dd);
        yh[4 * t + k] = h; da[4 * t + k] = dd;
        s1 += dd;
        s2 += dd * h;
      }
    }
  } else {
    for (int gq = slot; gq < ngroups; gq += SK_SLOTS) {
      uint4 rnd = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
      if (p > 0.f && !A.inject) rnd = rand4(sv.key, L.layer_id, c, gq);
      const uint32_t rr[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = gq * 4 + k;
        const bool ok = cok && r < B;
        const long long o = static_cast<long long>(r) * ld + cc;
        float h, dd;
        sk_bn_bwd_elem(L, A, r, c, ok, ok ? __ldcg(L.Y.ptr + o) : A.mean, ok ? ld_parts(L.dH, o) : 0.f, rr[k], h, dd);
        s1 += dd;
        s2 += dd * h;
      }
    }
  }
  sk_colsum2(s1, s2, sh, warp, lane);
  if (slot == 0 && cok) {
    const float ig = cx.inv_gs;
    if (sv.accum) { L.dbeta[c] += s1 * ig; L.dgamma[c] += s2 * ig; }
    else { L.dbeta[c] = s1 * ig; L.dgamma[c] = s2 * ig; L.dbias[c] = 0.f; }
  }
  const float fb = static_cast<float>(B);
  const float k0 = A.inv * A.g / fb;
  if constexpr (REG) {
#pragma unroll
    for (int t = 0; t < SK_RG; ++t)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = 4 * (slot + SK_SLOTS * t) + k;
        if (r < B && cok) {
          const long long o = static_cast<long long>(r) * ld + c;
          h_split(k0 * (fb * da[4 * t + k] - s1 - yh[4 * t + k] * s2), L.dYh[o], L.dYl[o]);
        }
      }
  } else {
    for (int gq = slot; gq < ngroups; gq += SK_SLOTS) {
      uint4 rnd = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
      if (p > 0.f && !A.inject) rnd = rand4(sv.key, L.layer_id, c, gq);
      const uint32_t rr[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = gq * 4 + k;
        if (r < B && cok) {
          const long long o = static_cast<long long>(r) * ld + c;
          float h, dd;
          sk_bn_bwd_elem(L, A, r, c, true, __ldcg(L.Y.ptr + o), ld_parts(L.dH, o), rr[k], h, dd);
          h_split(k0 * (fb * dd - s1 - h * s2), L.dYh[o], L.dYl[o]);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ reconstruction loss
// d xhat = gs * w_rec * 2 (xhat - x) / (B D); per-item partial of sum (xhat - x)^2; bias gradient of the last decoder
// Linear = column sums of d xhat / gs   (jamie/jamie.py:637-643).
template <bool REG>
__device__ void sk_rec_item(const ModCtx& M, const StepCtx& cx, const StepVars& sv, int cb, float* sh, int tid, int warp, int lane) {
  const int B = cx.B;
  const int c = cb * SK_CW + (tid & (SK_CW - 1));
  const int slot = tid / SK_CW;
  const bool cok = c < M.D;
  const int cc = cok ? c : 0;
  const int ld = M.ldD;
  const float kk = cx.gs * cx.sc.w[1] * 2.f / (static_cast<float>(B) * static_cast<float>(M.D));
  const bool reduce = M.xhat.n > 1;
  float sq = 0.f, cs = 0.f;
  for (int r0 = slot * 4; r0 < B; r0 += SK_SLOTS * 4) {
    float xh[4], xx[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = r0 + k;
      const bool ok = cok && r < B;
      const long long o = static_cast<long long>(r) * ld + cc;
      xh[k] = ok ? ld_parts(M.xhat, o) : 0.f;
      xx[k] = ok ? __ldcg(M.x + o) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = r0 + k;
      const float d = xh[k] - xx[k];
      sq += d * d;
      const float gx = kk * d;
      cs += gx;
      if (r < B && cok) {
        const long long o = static_cast<long long>(r) * ld + c;
        if (reduce) M.xhat.ptr[o] = xh[k];
        h_split(gx, M.dxh[o], M.dxl[o]);
      }
    }
  }
  sk_colsum2(sq, cs, sh, warp, lane);
  if (warp == 0) {
    if (lane < SK_CW && cok) {
      const float v = cs * cx.inv_gs;
      M.db5[c] = sv.accum ? M.db5[c] + v : v;
    }
    float t = (lane < SK_CW && cok) ? sq : 0.f;
    t = warp_sum(t);
    if (lane == 0) M.rec_part[cb] = t;
  }
  (void)REG;
}

// ------------------------------------------------------------------------------------------------ latent stage
// eps (injected or Philox Box-Muller) and z = mu + (exp(logvar/2) + 1e-7) eps   (jamie/model.py:230-240); also reduces
// the split-K partials of the heads GEMM into partial 0.
__device__ void sk_reparam(const StepCtx& cx, const StepVars& sv, int gt, int nt) {
  const int B = cx.B, L = cx.L;
  for (int t = gt; t < 2 * B * L; t += nt) {
    const int i = t / (B * L), rem = t - i * B * L, b = rem / L, l = rem - b * L;
    const ModCtx& M = cx.m[i];
    float e;
    if (sv.inject) {
      e = M.inj_eps[static_cast<long long>(b) * cx.LP + l];
    } else {
      const uint4 r = philox4x32(make_uint4(static_cast<uint32_t>(b), static_cast<uint32_t>(l), 0xE950u + i, 0x4A4Du), sv.key);
      const float u1 = (static_cast<float>(r.x >> 8) + 0.5f) * (1.0f / 16777216.0f);
      const float u2 = (static_cast<float>(r.y >> 8) + 0.5f) * (1.0f / 16777216.0f);
      e = sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
    }
    const long long om = static_cast<long long>(b) * cx.ldmv;
    const float mu = ld_parts(M.mulv, om + l);
    const float lv = ld_parts(M.mulv, om + L + l);
    if (M.mulv.n > 1) { M.mulv.ptr[om + l] = mu; M.mulv.ptr[om + L + l] = lv; }
    M.eps[static_cast<long long>(b) * cx.LP + l] = e;
    M.z[static_cast<long long>(b) * cx.LP + l] = mu + (expf(lv * 0.5f) + 1e-7f) * e;
  }
}

// out[l] (per lane, LAT_MAXT strided) = sum_b M[row, b] * V[b, l], skipping zero entries; also returns the row sum.
__device__ __forceinline__ float sk_row_times(const float* __restrict__ Mrow, const float* __restrict__ V, int B, int LP, int L, int lane,
                                              float (&acc)[LAT_MAXT]) {
#pragma unroll
  for (int t = 0; t < LAT_MAXT; ++t) acc[t] = 0.f;
  float rs = 0.f;
  for (int sup = 0; sup < B; sup += 512) {
    float m[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int b = sup + 32 * i + lane;
      m[i] = b < B ? __ldcg(Mrow + b) : 0.f;
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
          if (l < L) acc[t] += mv * __ldcg(v + l);
        }
      }
    }
  }
  return rs;
}

// combine (jamie/model.py:245-259): c_i = (s_i z_i + s_j C_i z_j) / (s_i + s_j rowsum(C_i)), C_0 = corr, C_1 = corr^T.
// Without F the latent loss partials need nothing of another row and are emitted here (fuse_loss).
__device__ void sk_combine(const StepCtx& cx, int gw, int nw, int lane) {
  const int B = cx.B, L = cx.L, LP = cx.LP;
  const int fuse_loss = cx.f_present ? 0 : 1;
  for (int w = gw; w < 2 * B; w += nw) {
    const int i = w / B, row = w - i * B, j = 1 - i;
    const ModCtx& M = cx.m[i];
    const float si = __ldcg(cx.sigma + i), sj = __ldcg(cx.sigma + j);
    const float* Ci = (i == 0 ? cx.corr : cx.corr_t) + static_cast<long long>(row) * B;
    float acc[LAT_MAXT];
    const float rs = sk_row_times(Ci, cx.m[j].z, B, LP, L, lane, acc);
    const float den = si + sj * rs;
    if (lane == 0) { M.den[row] = den; M.rs[row] = rs; }
    float smu = 0.f, scs = 0.f, sr = 0.f;
#pragma unroll
    for (int t = 0; t < LAT_MAXT; ++t) {
      const int l = lane + 32 * t;
      if (l < L) {
        const long long o = static_cast<long long>(row) * LP + l;
        M.S[o] = acc[t];
        const float zv = __ldcg(M.z + o);
        const float cv = (si * zv + sj * acc[t]) / den;
        M.c[o] = cv;
        h_split(cv, M.ch[o], M.cl[o]);
        if (fuse_loss) {
          const float mu = __ldcg(M.mulv.ptr + static_cast<long long>(row) * cx.ldmv + l);
          smu += mu * mu;
          const float d = zv - cv;
          scs += d * d;
          if (i == 0) { cx.lat_r[o] = cv; sr += cv * cv; }
        }
      }
    }
    if (fuse_loss) {
      smu = warp_sum(smu); scs = warp_sum(scs); sr = warp_sum(sr);
      if (lane == 0) {
        float* rp = cx.rowpart + (static_cast<long long>(i) * B + row) * 8;
        rp[0] = smu; rp[1] = scs; rp[2] = sr;
      }
    }
  }
}
// Row partial sums (rowpart[i][row][k]): 0: sum mu^2  1: sum (z - c)^2  2: sum r^2 (i = 0)  3: sum g z  4: sum g c  5: sum g S
// F residual r = c0 - F c1 (jamie/jamie.py:663-665).
__device__ void sk_latloss(const StepCtx& cx, int gw, int nw, int lane) {
  const int B = cx.B, L = cx.L, LP = cx.LP;
  for (int w = gw; w < 2 * B; w += nw) {
    const int i = w / B, row = w - i * B;
    const ModCtx& M = cx.m[i];
    float acc[LAT_MAXT];
    if (i == 0) sk_row_times(cx.fblk + static_cast<long long>(row) * B, cx.m[1].c, B, LP, L, lane, acc);
    else {
#pragma unroll
      for (int t = 0; t < LAT_MAXT; ++t) acc[t] = 0.f;
    }
    float smu = 0.f, scs = 0.f, sr = 0.f;
#pragma unroll
    for (int t = 0; t < LAT_MAXT; ++t) {
      const int l = lane + 32 * t;
      if (l < L) {
        const long long o = static_cast<long long>(row) * LP + l;
        const float mu = __ldcg(M.mulv.ptr + static_cast<long long>(row) * cx.ldmv + l);
        smu += mu * mu;
        const float d = __ldcg(M.z + o) - __ldcg(M.c + o);
        scs += d * d;
        if (i == 0) {
          const float r = __ldcg(M.c + o) - acc[t];
          cx.lat_r[o] = r;
          sr += r * r;
        }
      }
    }
    smu = warp_sum(smu); scs = warp_sum(scs); sr = warp_sum(sr);
    if (lane == 0) {
      float* rp = cx.rowpart + (static_cast<long long>(i) * B + row) * 8;
      rp[0] = smu; rp[1] = scs; rp[2] = sr;
    }
  }
}
// g_i = d(loss)/dc_i / den_i with d/dc_i = decoder dgrad - k_cos (z_i - c_i) + F term (everything times the loss scale).
__device__ void sk_latbc(const StepCtx& cx, int gw, int nw, int lane) {
  const int B = cx.B, L = cx.L, LP = cx.LP;
  const float k_cos = cx.gs * cx.sc.w[2] * 32.f * 2.f / (static_cast<float>(B) * static_cast<float>(L));
  const float k_f = cx.gs * cx.sc.w[3] * 2.f / (static_cast<float>(B) * static_cast<float>(L));
  for (int w = gw; w < 2 * B; w += nw) {
    const int i = w / B, row = w - i * B;
    const ModCtx& M = cx.m[i];
    float acc[LAT_MAXT];
    if (i == 1 && cx.f_present) sk_row_times(cx.fblk_t + static_cast<long long>(row) * B, cx.lat_r, B, LP, L, lane, acc);
    else {
#pragma unroll
      for (int t = 0; t < LAT_MAXT; ++t) acc[t] = 0.f;
    }
    const float den = __ldcg(M.den + row);
    float p3 = 0.f, p4 = 0.f, p5 = 0.f;
#pragma unroll
    for (int t = 0; t < LAT_MAXT; ++t) {
      const int l = lane + 32 * t;
      if (l < L) {
        const long long o = static_cast<long long>(row) * LP + l;
        const float z = __ldcg(M.z + o), c = __ldcg(M.c + o);
        const float dcd = ld_parts(M.dc, o);
        if (M.dc.n > 1) M.dc.ptr[o] = dcd;
        float dc = dcd - k_cos * (z - c);
        dc += i == 0 ? k_f * __ldcg(cx.lat_r + o) : -k_f * acc[t];
        const float g = dc / den;
        M.g[o] = g;
        p3 += g * z; p4 += g * c; p5 += g * __ldcg(M.S + o);
      }
    }
    p3 = warp_sum(p3); p4 = warp_sum(p4); p5 = warp_sum(p5);
    if (lane == 0) {
      float* rp = cx.rowpart + (static_cast<long long>(i) * B + row) * 8;
      rp[3] = p3; rp[4] = p4; rp[5] = p5;
    }
  }
}
// dz_i = k_cos (z_i - c_i) + s_i g_i + s_i C_i g_j ; then through the reparameterisation and the KL term
// (jamie/jamie.py:619-632 with the reference's logvar quirk: only rows 0 and 1 of modality 1's logvar get KL
// gradient, each scaled by the broadcast over the batch).
__device__ void sk_latbz(const StepCtx& cx, const StepVars& sv, int gw, int nw, int lane) {
  const int B = cx.B, L = cx.L, LP = cx.LP;
  const float k_cos = cx.gs * cx.sc.w[2] * 32.f * 2.f / (static_cast<float>(B) * static_cast<float>(L));
  const float kkl = cx.gs * sv.kl_coef;
  const float fbl = static_cast<float>(B) * static_cast<float>(L);
  for (int w = gw; w < 2 * B; w += nw) {
    const int i = w / B, row = w - i * B, j = 1 - i;
    const ModCtx& M = cx.m[i];
    const float si = __ldcg(cx.sigma + i);
    const float* Ci = (i == 0 ? cx.corr : cx.corr_t) + static_cast<long long>(row) * B;
    float acc[LAT_MAXT];
    sk_row_times(Ci, cx.m[j].g, B, LP, L, lane, acc);
#pragma unroll
    for (int t = 0; t < LAT_MAXT; ++t) {
      const int l = lane + 32 * t;
      if (l < L) {
        const long long o = static_cast<long long>(row) * LP + l;
        const long long om = static_cast<long long>(row) * cx.ldmv;
        const float mu = __ldcg(M.mulv.ptr + om + l), lv = __ldcg(M.mulv.ptr + om + L + l);
        const float dz = k_cos * (__ldcg(M.z + o) - __ldcg(M.c + o)) + si * __ldcg(M.g + o) + si * acc[t];
        const float dmu = dz + kkl * mu / fbl;
        float dlv = dz * __ldcg(M.eps + o) * 0.5f * expf(lv * 0.5f);
        if (i == 1 && row < 2) dlv += kkl * -0.5f * (1.f - expf(lv)) / static_cast<float>(L);
        M.dmulv[om + l] = dmu;
        M.dmulv[om + L + l] = dlv;
        h_split(dmu, M.dmh[om + l], M.dml[om + l]);
        h_split(dlv, M.dmh[om + L + l], M.dml[om + L + l]);
      }
    }
  }
}
// FINAL, CTA item 0: loss scalars and d sigma (fixed-order sums); items 1 ..: head bias gradients (16 of the 4L columns).
__device__ void sk_final_item(const StepCtx& cx, const StepVars& sv, int item, float* sh, int tid, int warp, int lane) {
  const int B = cx.B, L = cx.L;
  if (item > 0) {
    const int col = (item - 1) * SK_CW + (tid & (SK_CW - 1));
    const int slot = tid / SK_CW;
    const bool cok = col < 4 * L;
    const int i = cok ? col / (2 * L) : 0, cidx = cok ? col - i * 2 * L : 0;
    const float* src = cx.m[i].dmulv + cidx;
    float s = 0.f, dummy = 0.f;
    if (cok)
      for (int r = slot; r < B; r += SK_SLOTS) s += __ldcg(src + static_cast<long long>(r) * cx.ldmv);
    sk_colsum2(s, dummy, sh, warp, lane);
    if (slot == 0 && cok) {
      float* dst = cx.m[i].dbias_heads + cidx;
      const float v = s * cx.inv_gs;
      *dst = sv.accum ? *dst + v : v;
    }
    return;
  }
  float* tot = sh + 64;    // [14]: [i * 7 + k], k = 0..5 the rowpart sums, k = 6: sum_r (g.c)[r] * rowsum_i[r]
  float* aux = sh + 80;    // [0,1]: sum_l (1 + lv - exp lv) of logvar rows 0 / 1 (modality 1); [2,3]: sum (xhat - x)^2
  if (warp < 14) {
    const int i = warp / 7, k = warp % 7;
    float s = 0.f;
    for (int r = lane; r < B; r += 32) {
      const float* rp = cx.rowpart + (static_cast<long long>(i) * B + r) * 8;
      s += k < 6 ? __ldcg(rp + k) : __ldcg(rp + 4) * __ldcg(cx.m[i].rs + r);
    }
    s = warp_sum(s);
    if (lane == 0) tot[warp] = s;
  } else if (warp == 14) {
    for (int i = 0; i < 2; ++i) {
      float t1 = 0.f;
      for (int l = lane; l < L; l += 32) {
        const float lv = __ldcg(cx.m[1].mulv.ptr + static_cast<long long>(i) * cx.ldmv + L + l);
        t1 += 1.f + lv - expf(lv);
      }
      t1 = warp_sum(t1);
      if (lane == 0) aux[i] = t1;
    }
  } else if (warp == 15) {
    for (int i = 0; i < 2; ++i) {
      float sacc = 0.f;
      const int nb = (cx.m[i].D + SK_CW - 1) / SK_CW;
      for (int b = lane; b < nb; b += 32) sacc += __ldcg(cx.m[i].rec_part + b);
      sacc = warp_sum(sacc);
      if (lane == 0) aux[2 + i] = sacc;
    }
  }
  __syncthreads();
  if (tid == 0) {
    const float fB = static_cast<float>(B), fL = static_cast<float>(L);
    // KL value (jamie/jamie.py:619-628) with logvars = rows 0/1 of modality 1's logvar
    float kl = 0.f;
    for (int i = 0; i < 2; ++i) kl += -0.5f * (aux[i] / fL - tot[i * 7 + 0] / (fB * fL));
    const float l_kl = sv.kl_base * kl;
    float rec = 0.f;
    for (int i = 0; i < 2; ++i) rec += aux[2 + i] / (fB * static_cast<float>(cx.m[i].D));
    const float l_cos = 32.f * (tot[1] + tot[7 + 1]) / (fB * fL);
    const float l_f = tot[2] / (fB * fL);
    // d sigma (combine backward): i = 0: d s0 += sum g0.z0 - sum g0.c0 ; d s1 += sum g0.S0 - sum (g0.c0) rs0 (and symmetrically)
    const float ds0 = ((tot[3] - tot[4]) + (tot[7 + 5] - tot[7 + 6])) * cx.inv_gs;
    const float ds1 = ((tot[7 + 3] - tot[7 + 4]) + (tot[5] - tot[6])) * cx.inv_gs;
    cx.dsigma[0] = sv.accum ? cx.dsigma[0] + ds0 : ds0;
    cx.dsigma[1] = sv.accum ? cx.dsigma[1] + ds1 : ds1;
    const StepConsts& sc = cx.sc;
    const float total = sc.w[0] * l_kl + sc.w[1] * rec + sc.w[2] * l_cos + sc.w[3] * l_f;
    float* o = cx.out_loss + sv.row * 8;
    o[0] = l_kl; o[1] = rec; o[2] = l_cos; o[3] = l_f; o[4] = total; o[6] = 0.f; o[7] = 0.f;
    float* gt = cx.grad + cx.n_flat;   // 8 floats behind the flat gradients: the data-parallel all-reduce carries them
    gt[0] = l_kl; gt[1] = rec; gt[2] = l_cos; gt[3] = l_f; gt[4] = total;
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------ clip + Adam
__device__ void sk_norm(const StepCtx& cx, int cta, int ncta, double* shd, int tid) {
  const long long n4 = cx.n_flat / 4;
  double s = 0.0;
  const float4* g4 = reinterpret_cast<const float4*>(cx.grad);
  for (long long i = static_cast<long long>(cta) * SK_THREADS + tid; i < n4; i += static_cast<long long>(ncta) * SK_THREADS) {
    const float4 v = __ldcg(g4 + i);
    s += static_cast<double>(v.x) * v.x + static_cast<double>(v.y) * v.y + static_cast<double>(v.z) * v.z + static_cast<double>(v.w) * v.w;
  }
  shd[tid] = s;
  __syncthreads();
  for (int o = SK_THREADS / 2; o > 0; o >>= 1) {
    if (tid < o) shd[tid] += shd[tid + o];
    __syncthreads();
  }
  if (tid == 0) cx.norm_part[cta] = shd[0];
  __syncthreads();
}
// every CTA re-reduces the partials in the same order (identical clip coefficient everywhere), then
// g *= grad_scale * clip;  m, v, theta updated with torch.optim.Adam's formulas (jamie/jamie.py:739-741).
__device__ void sk_adam(const StepCtx& cx, const StepVars& sv, int cta, int ncta, double* shd, int tid) {
  double s = 0.0;
  for (int i = tid; i < ncta; i += SK_THREADS) s += __ldcg(cx.norm_part + i);
  shd[tid] = s;
  __syncthreads();
  for (int o = SK_THREADS / 2; o > 0; o >>= 1) {
    if (tid < o) shd[tid] += shd[tid + o];
    __syncthreads();
  }
  const StepConsts& sc = cx.sc;
  const double norm = sqrt(shd[0]) * static_cast<double>(sc.grad_scale);
  const float coef = static_cast<float>(fmin(1.0, static_cast<double>(sc.max_norm) / (norm + 1e-6))) * sc.grad_scale;
  if (cta == 0 && tid == 0) cx.out_loss[sv.row * 8 + 5] = static_cast<float>(norm);
  __syncthreads();
  const float b1 = sc.beta1, b2 = sc.beta2, eps = sc.adam_eps;
  const float step = sv.step_size, ibc2 = sv.inv_bc2_sqrt;
  const long long n4 = cx.n_flat / 4;
  float4* t4 = reinterpret_cast<float4*>(cx.theta);
  const float4* g4 = reinterpret_cast<const float4*>(cx.grad);
  float4* m4 = reinterpret_cast<float4*>(cx.adam_m);
  float4* v4 = reinterpret_cast<float4*>(cx.adam_v);
  for (long long i = static_cast<long long>(cta) * SK_THREADS + tid; i < n4; i += static_cast<long long>(ncta) * SK_THREADS) {
    const float4 gg = __ldcg(g4 + i);
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
    split4_store(cx.theta_hi + 4 * i, cx.theta_lo + 4 * i, tt.x, tt.y, tt.z, tt.w);   // next step's GEMM operand planes
  }
}

// theta -> fp16 planes over the whole flat buffer (after jb_set_params)
__global__ void k_hsplit_flat(const float* __restrict__ src, __half* __restrict__ hi, __half* __restrict__ lo, long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x)
    h_split(src[i], hi[i], lo[i]);
}
__global__ void k_set_accum(Ctl* ctl, int v) { ctl->accum = v; }
// after a launch of k_step: advance the plan cursor / optimizer step count / Philox stream, clear the injection flag
__global__ void k_ctl_advance(Ctl* ctl, int d_cursor, int d_adam, int clear_inject) {
  ctl->cursor += d_cursor;
  ctl->stream_id += static_cast<unsigned long long>(d_cursor);
  ctl->adam_t += d_adam;
  if (clear_inject) ctl->inject = 0;
}

// ------------------------------------------------------------------------------------------------ the kernel
// Runs phases [ph_lo, ph_hi) of `nsteps` consecutive steps (plan rows ctl->cursor ...). A launch that contains PH_ADAM
// uses optimizer step counts ctl->adam_t + 1 ...; the host advances ctl with k_ctl_advance after the launch.
// use_stage: the batch rows were copied into ModCtx::stage[ctl->host_slot] (host-batch step). row_bias: -1 for an
// update-only launch (the cursor was already advanced by the backward launch). ts (optional): CTA 0 records the global
// timer at kernel start (ts[0]) and at the end of every phase (ts[1 + step * PH_COUNT + phase]).
__global__ void __launch_bounds__(SK_THREADS, 1) k_step(const StepCtx* __restrict__ cxp, int ph_lo, int ph_hi, int nsteps,
                                                         unsigned int* bar, int use_stage, int row_bias,
                                                         unsigned long long* ts) {
  extern __shared__ uint8_t sk_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sk_smem_raw) + 1023) & ~uintptr_t(1023));
  HgCtrl* ctrl = reinterpret_cast<HgCtrl*>(smem);
  StepVars* svp = reinterpret_cast<StepVars*>(smem + 512);
  uint8_t* ring = smem + HG_CTRL_BYTES;
  uint8_t* stage = ring + HG_RING_BYTES;
  float* sh = reinterpret_cast<float*>(stage);            // reduction scratch of the element-wise phases (GEMM idle)
  double* shd = reinterpret_cast<double*>(stage + 4096);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = blockIdx.x, ncta = gridDim.x;
  const int gw = cta * SK_WARPS + warp, nw = ncta * SK_WARPS;
  const StepCtx& cx = *cxp;
  const uint32_t tmem_d = hg_setup(ctrl, warp, lane);
  HgPipe pp;
  unsigned int target = 0;
  const int B = cx.B;
  const bool reg = B <= 4 * SK_SLOTS * SK_RG;
  // control block: read once (nothing in this launch writes it)
  const long long cursor0 = cx.ctl->cursor, adam0 = cx.ctl->adam_t;
  const unsigned long long stream0 = cx.ctl->stream_id, seed = cx.ctl->seed;
  const int inject = cx.ctl->inject, accum = cx.ctl->accum, host_slot = cx.ctl->host_slot;

  if (ts != nullptr && cta == 0 && tid == 0) ts[0] = globaltimer_ns();
  for (int s = 0; s < nsteps; ++s) {
    if (tid == 0) {
      StepVars v;
      v.row = cursor0 + s + row_bias;
      const long long t = adam0 + s + 1;
      v.kl_base = cx.plan_kl[v.row];
      v.kl_coef = cx.sc.w[0] * v.kl_base;
      const double bc1 = 1.0 - pow(static_cast<double>(cx.sc.beta1), static_cast<double>(t));
      const double bc2 = 1.0 - pow(static_cast<double>(cx.sc.beta2), static_cast<double>(t));
      v.step_size = static_cast<float>(static_cast<double>(cx.sc.lr) / bc1);
      v.inv_bc2_sqrt = static_cast<float>(1.0 / sqrt(bc2));
      const unsigned long long sid = stream0 + static_cast<unsigned long long>(s) + 1ull;
      const unsigned long long k = seed ^ (sid * 0x9E3779B97F4A7C15ull);
      v.key = make_uint2(static_cast<uint32_t>(k), static_cast<uint32_t>(k >> 32));
      v.inject = inject; v.accum = accum; v.host_slot = host_slot;
      *svp = v;
    }
    __syncthreads();
    const StepVars& sv = *svp;

    for (int ph = ph_lo; ph < ph_hi; ++ph) {
      if (ph == PH_LATLOSS && !cx.f_present) {
        if (ts != nullptr && cta == 0 && tid == 0) ts[1 + s * PH_COUNT + ph] = globaltimer_ns();
        continue;
      }
      const int gi = gemm_index(ph);
      if (gi >= 0) {
        hg_run_phase(cx.probs, cx.gph[gi], cta, ncta, ctrl, ring, stage, tmem_d, pp, warp, lane);
        if (ph == PH_DGH) {
          // FINAL on the CTAs from the top down (idle in this phase at the headline shapes), after their own tiles
          const int nitems = 1 + (4 * cx.L + SK_CW - 1) / SK_CW;
          bool any = false;
          for (int it = 0; it < nitems; ++it) any = any || (ncta - 1 - (it % ncta)) == cta;
          if (any) {
            __syncthreads();
            for (int it = 0; it < nitems; ++it)
              if ((ncta - 1 - (it % ncta)) == cta) sk_final_item(cx, sv, it, sh, tid, warp, lane);
          }
        }
      } else {
        switch (ph) {
          case PH_GATHER: {
            const int nstrips = (B + 31) / 32;
            for (int it = 0; it < nstrips; ++it)
              if ((ncta - 1 - (it % ncta)) == cta) sk_corr_strip(cx, sv, it, sh, reinterpret_cast<float*>(ring), warp, lane);
            sk_gather_rows(cx, sv, use_stage, gw, nw, lane);
            break;
          }
          case PH_BN1: case PH_BN2: case PH_BN3: case PH_BN4: {
            const int which = ph == PH_BN1 ? 0 : (ph == PH_BN2 ? 1 : (ph == PH_BN3 ? 2 : 3));
            const int nb0 = (cx.bn[which][0].N + SK_CW - 1) / SK_CW, nb1 = (cx.bn[which][1].N + SK_CW - 1) / SK_CW;
            for (int it = cta; it < nb0 + nb1; it += ncta) {
              const int i = it >= nb0 ? 1 : 0;
              if (reg) sk_bn_fwd_item<true>(cx.bn[which][i], cx, sv, it - (i ? nb0 : 0), sh, tid, warp, lane);
              else sk_bn_fwd_item<false>(cx.bn[which][i], cx, sv, it - (i ? nb0 : 0), sh, tid, warp, lane);
            }
            break;
          }
          case PH_BNB1: case PH_BNB2: case PH_BNB3: case PH_BNB4: {
            const int which = ph == PH_BNB1 ? 0 : (ph == PH_BNB2 ? 1 : (ph == PH_BNB3 ? 2 : 3));
            const int nb0 = (cx.bn[which][0].N + SK_CW - 1) / SK_CW, nb1 = (cx.bn[which][1].N + SK_CW - 1) / SK_CW;
            for (int it = cta; it < nb0 + nb1; it += ncta) {
              const int i = it >= nb0 ? 1 : 0;
              if (reg) sk_bn_bwd_item<true>(cx.bn[which][i], cx, sv, it - (i ? nb0 : 0), sh, tid, warp, lane);
              else sk_bn_bwd_item<false>(cx.bn[which][i], cx, sv, it - (i ? nb0 : 0), sh, tid, warp, lane);
            }
            break;
          }
          case PH_REC: {
            const int nb0 = (cx.m[0].D + SK_CW - 1) / SK_CW, nb1 = (cx.m[1].D + SK_CW - 1) / SK_CW;
            for (int it = cta; it < nb0 + nb1; it += ncta) {
              const int i = it >= nb0 ? 1 : 0;
              sk_rec_item<true>(cx.m[i], cx, sv, it - (i ? nb0 : 0), sh, tid, warp, lane);
            }
            break;
          }
          case PH_REPARAM: sk_reparam(cx, sv, cta * SK_THREADS + tid, ncta * SK_THREADS); break;
          case PH_COMBINE: sk_combine(cx, gw, nw, lane); break;
          case PH_LATLOSS: sk_latloss(cx, gw, nw, lane); break;
          case PH_LATBC: sk_latbc(cx, gw, nw, lane); break;
          case PH_LATBZ: sk_latbz(cx, sv, gw, nw, lane); break;
          case PH_NORM: sk_norm(cx, cta, ncta, shd, tid); break;
          case PH_ADAM: sk_adam(cx, sv, cta, ncta, shd, tid); break;
          default: break;
        }
      }
      const bool last = (ph == ph_hi - 1) && (s == nsteps - 1);
      if (!last) {
        target += static_cast<unsigned int>(ncta);
        grid_barrier(bar, target);
      }
      if (ts != nullptr && cta == 0 && tid == 0) ts[1 + s * PH_COUNT + ph] = globaltimer_ns();
    }
  }
  hg_teardown(tmem_d, warp);
}

}  // namespace jb
