// Evaluation metrics of the reference on the GPU: FOSCTTM (jamie/evaluation.py:65-85, class method
// jamie/jamie.py:892-913), kNN label transfer (jamie/evaluation.py:114-132, jamie/jamie.py:943-961) and the per-feature
// Pearson correlation of imputed vs measured values (jamie/evaluation.py:491-513). All three are O(n^2 L) / O(n d)
// streaming problems over small embeddings: CUDA-core fp64 accumulation (counts must agree with the float64 host
// definition), shared-memory tiles, integer atomics only (deterministic).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace jb {

constexpr int MT_TILE = 64;      // rows of a x rows of b per block
constexpr int MT_LC = 32;        // latent columns per shared-memory chunk
constexpr int MT_THREADS = 256;  // 16 x 16 threads, 4 x 4 pairs each

// squared distance of matched pairs: diag[i] = |a_i - b_i|^2. The same float64 operations in the same order as a tile
// element of mt_tile (sequential fma over the latent columns), so that D[i, i] of the tile equals diag[i] bit for bit.
__global__ void k_pair_diag(const float* __restrict__ a, const float* __restrict__ b, long long n, int L, double* __restrict__ diag) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int l = 0; l < L; ++l) {
    const double d = static_cast<double>(a[i * L + l]) - static_cast<double>(b[i * L + l]);
    s = fma(d, d, s);
  }
  diag[i] = s;
}

// 64 x 64 tile of D[i, j] = |a_i - b_j|^2 in registers (4 x 4 per thread); `fn(i, j, d)` sees every valid pair
template <class F>
__device__ __forceinline__ void mt_tile(const float* __restrict__ a, long long na, const float* __restrict__ b, long long nb, int L,
                                        long long i0, long long j0, float (*As)[MT_LC + 1], float (*Bs)[MT_LC + 1], F&& fn) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double acc[4][4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[u][v] = 0.0;
  for (int l0 = 0; l0 < L; l0 += MT_LC) {
    __syncthreads();
    for (int t = threadIdx.x; t < MT_TILE * MT_LC; t += MT_THREADS) {
      const int r = t / MT_LC, c = t % MT_LC;
      As[r][c] = (i0 + r < na && l0 + c < L) ? a[(i0 + r) * L + l0 + c] : 0.f;
      Bs[r][c] = (j0 + r < nb && l0 + c < L) ? b[(j0 + r) * L + l0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int c = 0; c < MT_LC; ++c) {
      double av[4], bv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { av[u] = As[ty + 16 * u][c]; bv[u] = Bs[tx + 16 * u][c]; }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const double d = av[u] - bv[v];
          acc[u][v] = fma(d, d, acc[u][v]);
        }
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const long long i = i0 + ty + 16 * u, j = j0 + tx + 16 * v;
      if (i < na && j < nb) fn(i, j, acc[u][v]);
    }
}

// FOSCTTM raw count: #{(i, j): D[i, j] < D[i, i]} (a -> b) + #{(i, j): D[i, j] < D[j, j]} (b -> a), one pass over D
__global__ void __launch_bounds__(MT_THREADS) k_foscttm(const float* __restrict__ a, const float* __restrict__ b, long long n, int L,
                                                        const double* __restrict__ diag, unsigned long long* __restrict__ count) {
  __shared__ float As[MT_TILE][MT_LC + 1], Bs[MT_TILE][MT_LC + 1];
  __shared__ unsigned int wsum[MT_THREADS / 32];
  const long long tiles = (n + MT_TILE - 1) / MT_TILE;
  unsigned int local = 0;
  for (long long t = blockIdx.x; t < tiles * tiles; t += gridDim.x) {
    const long long i0 = (t / tiles) * MT_TILE, j0 = (t % tiles) * MT_TILE;
    mt_tile(a, n, b, n, L, i0, j0, As, Bs, [&](long long i, long long j, double d) {
      local += (d < diag[i]) ? 1u : 0u;
      local += (d < diag[j]) ? 1u : 0u;
    });
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long s = 0;
    for (int w = 0; w < MT_THREADS / 32; ++w) s += wsum[w];
    atomicAdd(count, s);
  }
}

// distances of a chunk of query rows to every reference row, as fp32 (rounded from the float64 sum): D[q, j]
__global__ void __launch_bounds__(MT_THREADS) k_dist_rows(const float* __restrict__ q, long long nq, const float* __restrict__ r, long long nr,
                                                          int L, float* __restrict__ D) {
  __shared__ float As[MT_TILE][MT_LC + 1], Bs[MT_TILE][MT_LC + 1];
  const long long ti = (nq + MT_TILE - 1) / MT_TILE, tj = (nr + MT_TILE - 1) / MT_TILE;
  for (long long t = blockIdx.x; t < ti * tj; t += gridDim.x) {
    const long long i0 = (t / tj) * MT_TILE, j0 = (t % tj) * MT_TILE;
    mt_tile(q, nq, r, nr, L, i0, j0, As, Bs, [&](long long i, long long j, double d) { D[i * nr + j] = static_cast<float>(d); });
  }
}

// One block per query row: the k nearest reference rows (ties by the lowest index, as a stable argsort gives), uniform
// votes over their classes, ties between classes to the lowest class (np.argmax). Radix select over the bit patterns of
// the non-negative fp32 distances (4 passes of 8 bits), then one ordered pass that admits the ties in index order.
constexpr int KV_THREADS = 256;
__global__ void __launch_bounds__(KV_THREADS) k_knn_vote(const float* __restrict__ D, long long nr, int k, const int* __restrict__ cls,
                                                         int n_classes, int* __restrict__ pred) {
  extern __shared__ int votes[];   // [n_classes]
  __shared__ unsigned int hist[256];
  __shared__ unsigned int s_prefix, s_need, s_run;
  __shared__ unsigned int wcnt[KV_THREADS / 32];
  const unsigned int* row = reinterpret_cast<const unsigned int*>(D + static_cast<long long>(blockIdx.x) * nr);
  const int tid = threadIdx.x;
  if (tid == 0) { s_prefix = 0u; s_need = static_cast<unsigned int>(k); }
  for (int c = tid; c < n_classes; c += KV_THREADS) votes[c] = 0;
  // the k-th smallest bit pattern T; s_need = how many elements equal to T are admitted
  for (int shift = 24; shift >= 0; shift -= 8) {
    hist[tid] = 0u;
    __syncthreads();
    const unsigned int prefix = s_prefix, mask = shift == 24 ? 0u : (0xFFFFFFFFu << (shift + 8));
    for (long long j = tid; j < nr; j += KV_THREADS) {
      const unsigned int v = row[j];
      if ((v & mask) == prefix) atomicAdd(&hist[(v >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      unsigned int need = s_need, b = 0;
      for (; b < 256u; ++b) {
        if (hist[b] >= need) break;
        need -= hist[b];
      }
      s_prefix = prefix | (b << shift);
      s_need = need;
    }
    __syncthreads();
  }
  const unsigned int T = s_prefix;
  if (tid == 0) s_run = 0u;
  __syncthreads();
  for (long long j0 = 0; j0 < nr; j0 += KV_THREADS) {
    const long long j = j0 + tid;
    const unsigned int v = j < nr ? row[j] : 0xFFFFFFFFu;
    if (j < nr && v < T) atomicAdd(&votes[cls[j]], 1);
    const bool tie = j < nr && v == T;
    const unsigned int bal = __ballot_sync(0xffffffffu, tie);
    if ((tid & 31) == 0) wcnt[tid >> 5] = __popc(bal);
    __syncthreads();
    if (tie) {   // ties in index order: the first s_need of them are among the k nearest
      unsigned int before = s_run;
      for (int w = 0; w < (tid >> 5); ++w) before += wcnt[w];
      before += __popc(bal & ((1u << (tid & 31)) - 1u));
      if (before < s_need) atomicAdd(&votes[cls[j]], 1);
    }
    __syncthreads();
    if (tid == 0) {
      unsigned int s = 0;
      for (int w = 0; w < KV_THREADS / 32; ++w) s += wcnt[w];
      s_run += s;
    }
    __syncthreads();
  }
  if (tid == 0) {
    int best = 0;
    for (int c = 1; c < n_classes; ++c)
      if (votes[c] > votes[best]) best = c;
    pred[blockIdx.x] = best;
  }
}

// per-feature moments of two [n, d] matrices over a slab of rows: part[slab][c] = {sum x, sum y, sum xx, sum yy, sum xy}
__global__ void k_col_moments(const float* __restrict__ x, const float* __restrict__ y, long long n, long long d, long long rows_per_slab,
                              double* __restrict__ part) {
  const long long c = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (c >= d) return;
  const long long r0 = blockIdx.y * rows_per_slab, r1 = r0 + rows_per_slab < n ? r0 + rows_per_slab : n;
  double sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
  for (long long r = r0; r < r1; ++r) {
    const double a = x[r * d + c], b = y[r * d + c];
    sx += a; sy += b; sxx += a * a; syy += b * b; sxy += a * b;
  }
  double* p = part + (static_cast<long long>(blockIdx.y) * d + c) * 5;
  p[0] = sx; p[1] = sy; p[2] = sxx; p[3] = syy; p[4] = sxy;
}

}  // namespace jb
