// Grouped TF32 GEMM on the 5th-gen tensor cores (tcgen05.mma kind::tf32, fp32 accumulators in TMEM),
// operands staged in shared memory by TMA, one 128 x BN output tile per CTA.
//
//   C[M,N] (+)= A * B^T (+ bias)   where A is logically [M,K], B is logically [N,K]
//
// Each operand can be stored K-major (row-major [rows,K], the forward-pass case) or MN-major
// (stored [K, rows] row-major, i.e. the transposed view of a row-major activation / weight) so that
// the same kernel serves the forward GEMM (K,K), dgrad (K,MN) and wgrad (MN,MN) of a Linear layer
// without materialising transposed copies. One launch covers a whole table of problems (both
// modalities, wgrad + dgrad of a stage, ...): blockIdx.x is a global tile id.
//
// Numerics: operands are fp32 in global memory; the tensor maps are typed TFLOAT32 so TMA rounds
// (round-to-nearest) to TF32 on the way into shared memory (verified on B200: a FLOAT32 map leaves the low
// mantissa bits and the tensor core then truncates, which biases every product low). Accumulation is fp32.
//
// Shared-memory layouts (verified on B200 with tools/gemm_lab.cu):
//   K-major operand : TMA SWIZZLE_128B, box {32 k, rows};  UMMA layout SWIZZLE_128B, SBO 1024, +32 B per K=8 step
//   MN-major operand: TMA SWIZZLE_128B_ATOM_32B, boxes {32 rows, 32 k} (4 KB each);  UMMA layout
//                     SWIZZLE_128B_BASE32B (the only MN-major layout for 32-bit operands), LBO 4096, SBO 512,
//                     +1024 B per K=8 step
//
// split = 1 (training GEMMs): error-compensated 3xTF32. The tensor maps are typed FLOAT32 (raw fp32 lands in shared
// memory) and the tensor core itself truncates every operand to TF32, so the raw tile doubles as the HIGH part
// hi = trunc(x) for free; the four epilogue warps, idle during the main loop, write the exact remainders
// lo = x - trunc(x) into a second tile, and each K step issues three MMAs  hi*hi + lo*hi + hi*lo  (lo*lo ~ 2^-22
// dropped). Products are then exact to ~2^-20, i.e. fp32-class results from the TF32 pipe. Needed because a single
// TF32 pass perturbs pre-activations by ~5e-4, which flips LeakyReLU' at ~4e-4 of the elements and costs ~2e-2 of
// relative gradient error (measured, profiles/parity_r1.md) -- far outside the 1e-3 parity tolerance.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (TMEM -> registers -> padded smem transpose -> coalesced 128-bit global stores).
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
constexpr int GEMM_EPI_PITCH = 36;                         // floats; 144 B rows: conflict-free 128-bit access
constexpr int GEMM_EPI_SMEM = 4 * 32 * GEMM_EPI_PITCH * 4; // 18 KB: one 32x32 transpose buffer per epilogue warp
constexpr int GEMM_SMEM_BYTES = GEMM_TILE_SMEM + GEMM_CTRL_SMEM + GEMM_EPI_SMEM + 1024;  // + alignment slack
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
  int accumulate;  // C += result (one CTA owns the tile: no atomics)
  float slope;
  uint32_t mn_lbo, mn_sbo;  // MN-major descriptor byte offsets (4096 / 512 for this tiling)
  uint32_t mn_layout;       // UMMA layout type of MN-major operands (1 = SWIZZLE_128B_BASE32B)
  int ks;                   // k-blocks (32 floats of K each) per pipeline stage: 1, 2 or 4
  long long* dbg;           // optional: per-CTA phase timestamps (bring-up lab only)
  int dbg_mode;             // lab only: 1 = TMA only (no MMA), 2 = MMA only (no TMA)
  int split;                // 1: error-compensated 3xTF32 (fp32-level accuracy), see the header comment
};

struct GemmCtrl {
  uint64_t full[GEMM_MAX_STAGES];
  uint64_t empty[GEMM_MAX_STAGES];
  uint64_t ready[GEMM_MAX_STAGES];   // split mode: lo tiles written by the splitter warps
  uint64_t tmem_full;
  uint32_t tmem_base;
};

__device__ __forceinline__ float leaky(float x, float slope) { return x > 0.f ? x : x * slope; }

// kLab = true adds the bring-up instrumentation (phase timestamps, TMA-only / MMA-only modes); the product
// instantiation (kLab = false) carries none of it.
template <bool kLab>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tf32_grouped_kernel(const GemmProblem* __restrict__ probs, int nprobs) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle (TMA and UMMA descriptors agree on it).
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  GemmCtrl* ctrl = reinterpret_cast<GemmCtrl*>(smem);
  uint8_t* tiles = smem + GEMM_CTRL_SMEM;
  float* epi_stage = reinterpret_cast<float*>(tiles + GEMM_TILE_SMEM);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  int p = 0;
  {
    const int tile = blockIdx.x;
    while (p + 1 < nprobs && tile >= probs[p + 1].tile_base) ++p;
  }
  const GemmProblem& P = probs[p];
  // Everything the pipelines need is read ONCE into registers: the inline-asm barriers below carry "memory"
  // clobbers, so any P.field left inside a loop would be re-fetched from global memory on every iteration.
  const int t = blockIdx.x - P.tile_base;
  const int tiles_n = P.tiles_n;
  const int tm = t / tiles_n, tn = t % tiles_n;
  const int m0 = tm * GEMM_BM;
  const int bn = P.bn;
  const int n0 = tn * bn;
  const int a_mn = P.a_mn, b_mn = P.b_mn;
  const int pM = P.M, pN = P.N, ldc = P.ldc, epi = P.epi, accumulate = P.accumulate;
  float* const pC = P.C;
  const float* const pbias = P.bias;
  const float slope = P.slope;
  const uint32_t mn_lbo = P.mn_lbo, mn_sbo = P.mn_sbo, mn_layout = P.mn_layout;
  const int dbg_mode = kLab ? P.dbg_mode : 0;
  const CUtensorMap* const tmA = &P.tmA;
  const CUtensorMap* const tmB = &P.tmB;
  const int num_kb = (P.K + GEMM_BK - 1) / GEMM_BK;
  const int kblock_bytes = GEMM_A_STAGE_BYTES + bn * GEMM_BK * 4;  // one k-block: A sub-tile then B sub-tile
  const int ks = P.ks;
  const int split = P.split;
  const int raw_bytes = ks * kblock_bytes;                 // raw (hi) k-blocks of a stage, contiguous
  const int stage_bytes = raw_bytes * (split ? 2 : 1);     // split: the lo k-blocks follow the raw ones
  int nstages = GEMM_TILE_SMEM / stage_bytes;
  if (nstages > GEMM_MAX_STAGES) nstages = GEMM_MAX_STAGES;
  const int num_st = (num_kb + ks - 1) / ks;  // pipeline iterations
  uint32_t tmem_cols = bn < 32 ? 32u : static_cast<uint32_t>(bn);  // power of two >= 32
  if (kLab && dbg_mode == 6) tmem_cols *= 2;  // lab: two independent accumulators

  long long* dbg = (kLab && P.dbg) ? P.dbg + static_cast<size_t>(blockIdx.x) * 8 : nullptr;
  if (dbg && threadIdx.x == 0) dbg[0] = clock64();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(tmA);
    tma_prefetch_desc(tmB);
    for (int s = 0; s < nstages; ++s) {
      mbar_init(&ctrl->full[s], 1);
      mbar_init(&ctrl->empty[s], 1);
      mbar_init(&ctrl->ready[s], 4);   // one arrival per splitter warp
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
    // ------------------------------------------------ TMA producer (whole warp loops, one elected lane issues)
    int s = 0;          // ring slot and its phase parity, advanced incrementally (no integer divisions in the loop)
    uint32_t ph = 0;
    for (int it = 0; it < ((kLab && dbg_mode >= 3) ? 0 : num_st); ++it) {
      mbar_wait(&ctrl->empty[s], ph ^ 1);
      const int kb0 = it * ks;
      const int nk = min(ks, num_kb - kb0);
      if (elect_one()) {
        if (kLab && dbg_mode == 2) {
          mbar_arrive(&ctrl->full[s]);
        } else {
          mbar_arrive_expect_tx(&ctrl->full[s], static_cast<uint32_t>(nk * kblock_bytes));
          for (int j = 0; j < nk; ++j) {
            uint8_t* sa = tiles + s * stage_bytes + j * kblock_bytes;
            uint8_t* sb = sa + GEMM_A_STAGE_BYTES;
            const int k0 = (kb0 + j) * GEMM_BK;
            if (!a_mn) {
              tma_load_2d(sa, tmA, &ctrl->full[s], k0, m0);  // box {32 k, 128 rows}
            } else {
              for (int i = 0; i < GEMM_BM / 32; ++i)  // box {32 rows(contiguous), 32 k}
                tma_load_2d(sa + i * 4096, tmA, &ctrl->full[s], m0 + 32 * i, k0);
            }
            if (!b_mn) {
              tma_load_2d(sb, tmB, &ctrl->full[s], k0, n0);  // box {32 k, bn rows}
            } else {
              for (int i = 0; i < bn / 32; ++i) tma_load_2d(sb + i * 4096, tmB, &ctrl->full[s], n0 + 32 * i, k0);
            }
          }
        }
      }
      __syncwarp();
      if (++s == nstages) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (whole warp loops, one elected lane issues)
    const uint32_t idesc = umma_idesc_tf32(GEMM_BM, bn, a_mn, b_mn);
    const uint32_t a_step16 = a_mn ? 64u : 2u;  // descriptor start-address units (16 B) per UMMA_K step
    const uint32_t b_step16 = b_mn ? 64u : 2u;
    // descriptor bits that do not change across the k loop (everything but the start address)
    const uint64_t da_hi = umma_smem_desc(0u, a_mn ? mn_lbo : 16u, a_mn ? mn_sbo : 1024u, a_mn ? mn_layout : 2u);
    const uint64_t db_hi = umma_smem_desc(0u, b_mn ? mn_lbo : 16u, b_mn ? mn_sbo : 1024u, b_mn ? mn_layout : 2u);
    const uint32_t tiles_u32 = smem_u32(tiles);
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < num_st; ++it) {
      if (!(kLab && dbg_mode >= 3)) mbar_wait(split ? &ctrl->ready[s] : &ctrl->full[s], ph);     // lab modes 3/4: no pipeline at all
      if (!(kLab && dbg_mode == 5)) tc_fence_after();
      if (kLab && dbg && it == 0 && lane == 0) dbg[2] = clock64();
      const int kb0 = it * ks;
      const int nk = min(ks, num_kb - kb0);
      if (elect_one()) {
        if (kLab && dbg_mode == 1) {
          mbar_arrive(&ctrl->empty[s]);
        } else {
          for (int j = 0; j < nk; ++j) {
            const uint32_t sa = tiles_u32 + s * stage_bytes + j * kblock_bytes;  // 1024-aligned: (addr >> 4) + k*step
            const uint32_t sb = sa + GEMM_A_STAGE_BYTES;                        // never carries out of the 14-bit field
            const uint64_t da0 = da_hi | static_cast<uint64_t>((sa >> 4) & 0x3FFFu);
            const uint64_t db0 = db_hi | static_cast<uint64_t>((sb >> 4) & 0x3FFFu);
            if (split) {
              const uint32_t lo16 = static_cast<uint32_t>(raw_bytes) >> 4;   // lo tile = raw tile + raw_bytes
#pragma unroll
              for (int k = 0; k < GEMM_BK / GEMM_UMMA_K; ++k) {
                const uint64_t da = da0 + k * a_step16, db = db0 + k * b_step16;
                umma_tf32(tmem_d, da + lo16, db, idesc, (kb0 | j | k) != 0 ? 1u : 0u);   // lo(A) * hi(B)
                umma_tf32(tmem_d, da, db + lo16, idesc, 1u);                            // hi(A) * lo(B)
                umma_tf32(tmem_d, da, db, idesc, 1u);                                   // hi(A) * hi(B)
              }
            } else {
#pragma unroll
              for (int k = 0; k < GEMM_BK / GEMM_UMMA_K; ++k)
                umma_tf32(tmem_d + ((kLab && dbg_mode == 6 && (k & 1)) ? bn : 0), da0 + k * a_step16, db0 + k * b_step16,
                          idesc, (kb0 | j | k) != 0 ? 1u : 0u);
            }
          }
          if (!(kLab && (dbg_mode == 3 || dbg_mode == 5 || dbg_mode == 6))) umma_commit(&ctrl->empty[s]);  // frees the smem slot when these MMAs have read it
        }
      }
      __syncwarp();
      if (++s == nstages) { s = 0; ph ^= 1; }
    }
    if (elect_one()) {
      if (kLab && dbg_mode == 1) mbar_arrive(&ctrl->tmem_full);
      else umma_commit(&ctrl->tmem_full);  // accumulator complete
      if (kLab && dbg) dbg[3] = clock64();
    }
    __syncwarp();
  } else {
    // ------------------------------------------------ epilogue: TMEM -> registers -> smem transpose -> global
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    float* st = epi_stage + q * 32 * GEMM_EPI_PITCH;
    if (split) {
      // ---- splitter: lo = x - trunc_tf32(x), elementwise over the raw k-blocks of every stage (layout-agnostic)
      const int et = (warp - 2) * 32 + lane;   // 0..127
      int s = 0;
      uint32_t ph = 0;
      for (int it = 0; it < num_st; ++it) {
        mbar_wait(&ctrl->full[s], ph);
        const int nk = min(ks, num_kb - it * ks);
        const float4* src = reinterpret_cast<const float4*>(tiles + s * stage_bytes);
        float4* dst = reinterpret_cast<float4*>(tiles + s * stage_bytes + raw_bytes);
        const int n4 = (nk * kblock_bytes) >> 4;
#pragma unroll 4
        for (int i = et; i < n4; i += 128) {
          const float4 x = src[i];
          float4 l;
          l.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
          l.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
          l.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u);
          l.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
          dst[i] = l;
        }
        fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctrl->ready[s]);
        if (++s == nstages) { s = 0; ph ^= 1; }
      }
    }
    mbar_wait(&ctrl->tmem_full, 0);
    tc_fence_after();
    if (dbg && warp == 2 && lane == 0) dbg[4] = clock64();
    const bool vec_ok = (ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(pC) & 15) == 0;
    const int rsub = lane >> 3, ch = lane & 7;  // read-back mapping: 4 rows x 8 float4 per pass
    for (int c0 = 0; c0 < bn; c0 += 32) {
      const int nbase = n0 + c0;
      if (nbase >= pN) break;  // warp-uniform
      float v[32];
      tmem_ld_32x32(tmem_d + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c0), v);
      tmem_ld_wait();
      if (epi != EPI_STORE) {
        const float bl = (nbase + lane < pN) ? __ldg(pbias + nbase + lane) : 0.f;  // one coalesced load, then shuffles
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = v[j] + __shfl_sync(0xffffffffu, bl, j);
          if (epi == EPI_BIAS_LRELU) x = leaky(x, slope);
          v[j] = x;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(st + lane * GEMM_EPI_PITCH + 4 * j) =
            make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int r = it * 4 + rsub;
        const int grow = m0 + q * 32 + r;
        const int n = nbase + ch * 4;
        float4 x = *reinterpret_cast<const float4*>(st + r * GEMM_EPI_PITCH + ch * 4);
        if (grow < pM && n < pN) {
          float* dst = pC + static_cast<size_t>(grow) * ldc + n;
          if (vec_ok && n + 4 <= pN) {
            if (accumulate) {
              const float4 o = *reinterpret_cast<const float4*>(dst);
              x.x += o.x; x.y += o.y; x.z += o.z; x.w += o.w;
            }
            *reinterpret_cast<float4*>(dst) = x;
          } else {
            const float xs[4] = {x.x, x.y, x.z, x.w};
            for (int j = 0; j < 4; ++j)
              if (n + j < pN) dst[j] = accumulate ? dst[j] + xs[j] : xs[j];
          }
        }
      }
      __syncwarp();
    }
  }
  if (dbg && warp == 2 && lane == 0) dbg[5] = clock64();
  tc_fence_before();
  __syncthreads();
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

// 2-D fp32 tensor map: `inner` contiguous elements, `outer` rows of `ld` elements, out-of-bounds box elements read
// as zero. dtype_tf32: TMA rounds to TF32 on load. swizzle_atom32: the 32-byte-atom flavour of the 128-byte swizzle
// (MN-major operands).  Requirements: base 16-byte aligned, ld a multiple of 4 floats.
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
                             float slope, int accumulate, int dtype_tf32 = 1, int split = 0) {
  *g = GemmProblem{};
  if (split) dtype_tf32 = 0;   // raw fp32 in shared memory: the tensor core truncates (hi), the splitter adds lo
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
  g->slope = slope;
  g->mn_lbo = 4096; g->mn_sbo = 512; g->mn_layout = 1;
  g->accumulate = accumulate;
  g->ks = 2;
  g->split = split;
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

template <bool kLab = false>
inline cudaError_t gemm_launch(const GemmProblem* dev_table, int nprobs, int total_tiles, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tf32_grouped_kernel<kLab>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         GEMM_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  gemm_tf32_grouped_kernel<kLab><<<total_tiles, GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(dev_table, nprobs);
  return cudaGetLastError();
}

}  // namespace jb
