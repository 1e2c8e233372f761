// Grouped TF32 GEMM on the 5th-gen tensor cores (tcgen05.mma kind::tf32, fp32 accumulators in TMEM),
// operands staged in shared memory by TMA with the 128-byte swizzle, one 128 x BN output tile per CTA.
//
//   C[M,N] = A * B^T   where A is logically [M,K], B is logically [N,K]
//
// Each operand can be stored K-major (row-major [rows,K], the forward-pass case) or MN-major
// (stored [K, rows] row-major, i.e. the transposed view of a row-major activation / weight) so that
// the same kernel serves the forward GEMM (K,K), dgrad (K,MN) and wgrad (MN,MN) of a Linear layer
// without materialising transposed copies. One launch covers a whole table of problems (both
// modalities, several layers): blockIdx.x is a global tile id.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (TMEM -> registers -> global), one TMEM lane quadrant each.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include "ptx.cuh"

namespace jb {

constexpr int GEMM_BM = 128;           // UMMA M (TMEM lanes)
constexpr int GEMM_BK = 32;            // fp32 elements per k-block = one 128-byte swizzle row
constexpr int GEMM_UMMA_K = 8;         // tf32: 32 bytes of K per instruction
constexpr int GEMM_THREADS = 192;
constexpr int GEMM_A_STAGE_BYTES = GEMM_BM * GEMM_BK * 4;  // 16 KB
constexpr int GEMM_TILE_SMEM = 192 * 1024;                 // ring of operand stages
constexpr int GEMM_CTRL_SMEM = 1024;                       // barriers + tmem slot
constexpr int GEMM_SMEM_BYTES = GEMM_TILE_SMEM + GEMM_CTRL_SMEM + 1024;  // + alignment slack
constexpr int GEMM_MAX_STAGES = 8;

enum GemmEpilogue : int {
  EPI_STORE = 0,       // C = acc
  EPI_BIAS = 1,        // C = acc + bias[n]
  EPI_BIAS_LRELU = 2,  // C = leaky_relu(acc + bias[n], slope)      (inference, BatchNorm folded)
};

struct alignas(128) GemmProblem {
  CUtensorMap tmA;  // 128 B each
  CUtensorMap tmB;
  float* C;
  const float* bias;
  int M, N, K, ldc;
  int bn;          // N tile = UMMA N (32, 64, 128 or 256)
  int a_mn, b_mn;  // 0 = K-major operand, 1 = MN-major operand
  int epi;
  int tiles_m, tiles_n, tile_base;
  int round_out;  // round the stored value to tf32 (it feeds another tensor-core GEMM)
  float slope;
  uint32_t mn_lbo, mn_sbo;  // MN-major descriptor byte offsets (4096 / 512 for this tiling)
  int accumulate;           // C += result (fp32 atomics-free: one CTA owns the tile)
  uint32_t mn_layout;       // UMMA layout type of MN-major operands (1 = SWIZZLE_128B_BASE32B)
  int pad_[2];
  long long* dbg;           // optional: per-CTA phase timestamps (bring-up lab only)
};

struct GemmCtrl {
  uint64_t full[GEMM_MAX_STAGES];
  uint64_t empty[GEMM_MAX_STAGES];
  uint64_t tmem_full;
  uint32_t tmem_base;
};

__device__ __forceinline__ float leaky(float x, float slope) { return x > 0.f ? x : x * slope; }

__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tf32_grouped_kernel(const GemmProblem* __restrict__ probs, int nprobs) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle (TMA and UMMA descriptors agree on it).
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  GemmCtrl* ctrl = reinterpret_cast<GemmCtrl*>(smem);
  uint8_t* tiles = smem + GEMM_CTRL_SMEM;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  int p = 0;
  {
    const int tile = blockIdx.x;
    while (p + 1 < nprobs && tile >= probs[p + 1].tile_base) ++p;
  }
  const GemmProblem& P = probs[p];
  const int t = blockIdx.x - P.tile_base;
  const int tm = t / P.tiles_n, tn = t % P.tiles_n;
  const int m0 = tm * GEMM_BM;
  const int bn = P.bn;
  const int n0 = tn * bn;
  const int num_kb = (P.K + GEMM_BK - 1) / GEMM_BK;
  const int b_stage_bytes = bn * GEMM_BK * 4;
  const int stage_bytes = GEMM_A_STAGE_BYTES + b_stage_bytes;
  int nstages = GEMM_TILE_SMEM / stage_bytes;
  if (nstages > GEMM_MAX_STAGES) nstages = GEMM_MAX_STAGES;
  const uint32_t tmem_cols = bn < 32 ? 32u : static_cast<uint32_t>(bn);  // power of two >= 32

  long long* dbg = P.dbg ? P.dbg + static_cast<size_t>(blockIdx.x) * 8 : nullptr;
  if (dbg && threadIdx.x == 0) { dbg[0] = clock64(); dbg[6] = static_cast<long long>(globaltimer_ns()); }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.tmA);
    tma_prefetch_desc(&P.tmB);
    for (int s = 0; s < nstages; ++s) {
      mbar_init(&ctrl->full[s], 1);
      mbar_init(&ctrl->empty[s], 1);
    }
    mbar_init(&ctrl->tmem_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&ctrl->tmem_base, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = ctrl->tmem_base;
  if (dbg && threadIdx.x == 0) dbg[1] = clock64();

  if (warp == 0) {
    // ------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % nstages;
        const uint32_t ph = (kb / nstages) & 1;
        mbar_wait(&ctrl->empty[s], ph ^ 1);
        uint8_t* sa = tiles + s * stage_bytes;
        uint8_t* sb = sa + GEMM_A_STAGE_BYTES;
        mbar_arrive_expect_tx(&ctrl->full[s], static_cast<uint32_t>(stage_bytes));
        const int k0 = kb * GEMM_BK;
        if (!P.a_mn) {
          tma_load_2d(sa, &P.tmA, &ctrl->full[s], k0, m0);  // box {32 k, 128 rows}
        } else {
          for (int i = 0; i < GEMM_BM / 32; ++i)  // box {32 rows(contiguous), 32 k}
            tma_load_2d(sa + i * 4096, &P.tmA, &ctrl->full[s], m0 + 32 * i, k0);
        }
        if (!P.b_mn) {
          tma_load_2d(sb, &P.tmB, &ctrl->full[s], k0, n0);  // box {32 k, bn rows}
        } else {
          for (int i = 0; i < bn / 32; ++i) tma_load_2d(sb + i * 4096, &P.tmB, &ctrl->full[s], n0 + 32 * i, k0);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (single thread)
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(GEMM_BM, bn, P.a_mn, P.b_mn);
      const uint32_t a_step = P.a_mn ? 1024u : 32u;  // bytes per UMMA_K step
      const uint32_t b_step = P.b_mn ? 1024u : 32u;
      const uint32_t a_lbo = P.a_mn ? P.mn_lbo : 16u, a_sbo = P.a_mn ? P.mn_sbo : 1024u;
      const uint32_t b_lbo = P.b_mn ? P.mn_lbo : 16u, b_sbo = P.b_mn ? P.mn_sbo : 1024u;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % nstages;
        const uint32_t ph = (kb / nstages) & 1;
        mbar_wait(&ctrl->full[s], ph);
        tc_fence_after();
        if (dbg && kb == 0) dbg[2] = clock64();
        const uint32_t sa = smem_u32(tiles + s * stage_bytes);
        const uint32_t sb = sa + GEMM_A_STAGE_BYTES;
#pragma unroll
        for (int k = 0; k < GEMM_BK / GEMM_UMMA_K; ++k) {
          const uint64_t da = umma_smem_desc(sa + k * a_step, a_lbo, a_sbo, P.a_mn ? P.mn_layout : 2u);
          const uint64_t db = umma_smem_desc(sb + k * b_step, b_lbo, b_sbo, P.b_mn ? P.mn_layout : 2u);
          umma_tf32(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&ctrl->empty[s]);  // frees the smem slot when these MMAs have read it
      }
      umma_commit(&ctrl->tmem_full);  // accumulator complete
      if (dbg) dbg[3] = clock64();
    }
    __syncwarp();
  } else {
    // ------------------------------------------------ epilogue: TMEM -> registers -> global
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const int row = m0 + q * 32 + lane;
    mbar_wait(&ctrl->tmem_full, 0);
    tc_fence_after();
    if (dbg && warp == 2 && lane == 0) dbg[4] = clock64();
    const bool row_ok = row < P.M;
    float* crow = P.C + static_cast<size_t>(row_ok ? row : 0) * P.ldc;
    const bool vec_ok = (P.ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(P.C) & 15) == 0;
    for (int c0 = 0; c0 < bn; c0 += 32) {
      float v[32];
      tmem_ld_32x32(tmem_d + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c0), v);
      tmem_ld_wait();
      const int nbase = n0 + c0;
      if (nbase >= P.N) break;  // warp-uniform
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float x = v[j];
        const int n = nbase + j;
        if (P.epi != EPI_STORE && n < P.N) x += __ldg(P.bias + n);
        if (P.epi == EPI_BIAS_LRELU) x = leaky(x, P.slope);
        v[j] = x;
      }
      if (row_ok) {
        if (P.accumulate) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (nbase + j < P.N) v[j] += crow[nbase + j];
        }
        if (P.round_out) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = tf32_rna(v[j]);
        }
        if (vec_ok && nbase + 32 <= P.N) {
          float4* dst = reinterpret_cast<float4*>(crow + nbase);
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (nbase + j < P.N) crow[nbase + j] = v[j];
        }
      }
    }
  }
  if (dbg && warp == 2 && lane == 0) dbg[5] = clock64();
  tc_fence_before();
  __syncthreads();
  if (dbg && threadIdx.x == 0) dbg[7] = static_cast<long long>(globaltimer_ns());
  if (warp == 1) tmem_dealloc(tmem_d, tmem_cols);
}

// =========================================================================== host side
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);

inline PFN_tmapEncodeTiled tmap_encode_fn() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return nullptr;
    fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  }
  return fn;
}

// 2-D fp32 tensor map: `inner` contiguous elements, `outer` rows of `ld` elements, 128B swizzle,
// out-of-bounds box elements read as zero.
inline int make_tmap_2d(CUtensorMap* tm, const float* base, uint64_t inner, uint64_t outer, uint64_t ld,
                        uint32_t box_inner, uint32_t box_outer, int dtype_tf32 = 1, int swizzle_atom32 = 0) {
  PFN_tmapEncodeTiled fn = tmap_encode_fn();
  if (!fn) return -1;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * sizeof(float)};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, dtype_tf32 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                  const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : static_cast<int>(r);
}

inline int pick_bn(int N) {
  if (N <= 32) return 32;
  if (N <= 64) return 64;
  return 128;
}

// Fill one problem entry. A is logically [M,K]: K-major => memory [M][lda]; MN-major => memory [K][lda].
// Same for B with N. Returns 0 on success.
inline int gemm_problem_fill(GemmProblem* g, const float* A, int lda, int a_mn, const float* B, int ldb, int b_mn,
                             float* C, int ldc, int M, int N, int K, int bn, int epi, const float* bias,
                             float slope, int round_out, int accumulate, int dtype_tf32 = 1) {
  *g = GemmProblem{};
  int rc;
  if (!a_mn) rc = make_tmap_2d(&g->tmA, A, K, M, lda, GEMM_BK, GEMM_BM, dtype_tf32);
  else rc = make_tmap_2d(&g->tmA, A, M, K, lda, 32, GEMM_BK, dtype_tf32, 1);
  if (rc) return rc;
  if (!b_mn) rc = make_tmap_2d(&g->tmB, B, K, N, ldb, GEMM_BK, bn, dtype_tf32);
  else rc = make_tmap_2d(&g->tmB, B, N, K, ldb, 32, GEMM_BK, dtype_tf32, 1);
  if (rc) return rc;
  g->C = C; g->bias = bias;
  g->M = M; g->N = N; g->K = K; g->ldc = ldc;
  g->bn = bn; g->a_mn = a_mn; g->b_mn = b_mn; g->epi = epi;
  g->tiles_m = (M + GEMM_BM - 1) / GEMM_BM;
  g->tiles_n = (N + bn - 1) / bn;
  g->tile_base = 0;
  g->round_out = round_out; g->slope = slope;
  g->mn_lbo = 4096; g->mn_sbo = 512; g->mn_layout = 1;
  g->accumulate = accumulate;
  return 0;
}

// Assign tile ranges; returns the total tile count (= grid size).
inline int gemm_table_finalize(GemmProblem* g, int n) {
  int base = 0;
  for (int i = 0; i < n; ++i) {
    g[i].tile_base = base;
    base += g[i].tiles_m * g[i].tiles_n;
  }
  return base;
}

inline cudaError_t gemm_launch(const GemmProblem* dev_table, int nprobs, int total_tiles, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tf32_grouped_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         GEMM_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  gemm_tf32_grouped_kernel<<<total_tiles, GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(dev_table, nprobs);
  return cudaGetLastError();
}

}  // namespace jb
