// Shared definitions of the step kernel (control block, constants, Philox4x32-10) and the small stand-alone kernels of the
// inference / ingest paths (BatchNorm folding, strided copy, TF32 split, standardisation). Round 1's per-phase training
// kernels lived here; the whole step is now stepk.cuh's persistent kernel.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "ptx.cuh"

namespace jb {

constexpr float BN_EPS = 1e-5f;
constexpr float BN_MOM = 0.1f;
constexpr float LRELU = 0.01f;

// ------------------------------------------------------------------------------------------------ step control
// Device-resident control block: the step kernel derives every step-varying scalar from it and advances it itself.
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

// 4 keep decisions (rows 4*rgroup .. 4*rgroup+3 of column col)
__device__ __forceinline__ uint4 rand4(uint2 key, unsigned layer_id, int col, int rgroup) {
  return philox4x32(make_uint4(static_cast<uint32_t>(rgroup), static_cast<uint32_t>(col), layer_id, 0x4A4Du), key);
}

constexpr int LAT_MAXT = 4;  // latent width up to 128

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

// ------------------------------------------------------------------------------------------------ PCA fit (Gram matrix)
// column sums of a slab of rows in float64: part[slab][c]
__global__ void k_col_sums(const float* __restrict__ x, long long n, long long d, long long rows_per_slab, double* __restrict__ part) {
  const long long c = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (c >= d) return;
  const long long r0 = blockIdx.y * rows_per_slab, r1 = r0 + rows_per_slab < n ? r0 + rows_per_slab : n;
  double s = 0.0;
  for (long long r = r0; r < r1; ++r) s += x[r * d + c];
  part[static_cast<long long>(blockIdx.y) * d + c] = s;
}
// G64[r, c] += G32[r, c] (the float64 accumulator over row chunks of the fp32-class chunk products)
__global__ void k_acc_f64(const float* __restrict__ g32, long long ld32, double* __restrict__ g64, long long d) {
  const long long r = blockIdx.x;
  for (long long c = threadIdx.x; c < d; c += blockDim.x) g64[r * d + c] += static_cast<double>(g32[r * ld32 + c]);
}

}  // namespace jb
