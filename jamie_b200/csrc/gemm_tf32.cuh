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
// Numerics: accumulation is fp32 in TMEM. Two operand modes:
//   split = 0 (wgrad, inference): one TF32 pass. Operands are either fp32 arrays read through TFLOAT32 tensor maps
//       (TMA rounds to nearest on the way into shared memory; a FLOAT32 map would leave the low mantissa bits and the
//       tensor core then truncates, which biases every product low -- verified on B200) or planes that already hold
//       TF32-representable values.
//   split = 1 (forward and dgrad of the training step): error-compensated 3xTF32 on PRE-SPLIT operands. Every producer
//       kernel of the step (gather, BatchNorm slabs, latent kernels, Adam) writes its result x as two planes
//       hi = rna_tf32(x), lo = rna_tf32(x - hi), so |x - hi - lo| <= 2^-23 |x| and both planes are exactly
//       representable (the tensor core's truncation is then a no-op and nothing is biased). Per 8-wide K step the MMA
//       warp issues TWO instructions:  D[:, 0:2bn] += A_hi * [B_hi ; B_lo]^T   (B_hi and B_lo tiles are adjacent in
//       shared memory, so one N = 2 bn instruction reads A_hi once) and  D[:, 0:bn] += A_lo * B_hi^T. The epilogue adds
//       the two column groups. Only lo*lo (<= 2^-22 relative, unbiased) is dropped: fp32-class products from the TF32
//       pipe. Needed because a single TF32 pass perturbs pre-activations by ~5e-4, which flips LeakyReLU' at ~4e-4 of
//       the elements and costs ~2e-2 of relative gradient error -- far outside the 1e-3 parity tolerance.
//
// Shared-memory layouts (verified on B200 with tools/gemm_lab.cu):
//   K-major operand : TMA SWIZZLE_128B, box {32 k, rows};  UMMA layout SWIZZLE_128B, SBO 1024, +32 B per K=8 step
//   MN-major operand: TMA SWIZZLE_128B_ATOM_32B, boxes {32 rows, 32 k} (4 KB each);  UMMA layout
//                     SWIZZLE_128B_BASE32B (the only MN-major layout for 32-bit operands), LBO 4096, SBO 512,
//                     +1024 B per K=8 step
// One k-block (32 floats of K) in shared memory: [A_hi 16 KB][A_lo 16 KB][B_hi bn*128 B][B_lo bn*128 B] (split) or
// [A][B]; the ring holds as many k-blocks as fit in 192 KB (4 for split bn = 64, 8 for single-pass).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (TMEM -> registers -> padded smem transpose -> coalesced 128-bit global stores).
//
// Split-K (ck = 2 or 4, a launch attribute): measured on B200, a CTA ingests its operand k-blocks at ~100 GB/s whatever
// the grid size (profiles/gemm_r1_s2_multicast_lab.md), so a stage is as slow as its busiest CTA. Stages with few tiles
// (K = 1024 layers: 64 tiles; heads and latent dgrads: 8 tiles) therefore launch a cluster of ck CTAs per tile, each
// accumulating 1/ck of K, and reduce-scatter the partial tiles through distributed shared memory: the rows of epilogue
// warp q are finished by rank q * ck / 4, which adds the deposits in ascending rank order (deterministic).
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
constexpr int GEMM_DRAIN_KB = 2;       // split: k-blocks accumulated in TMEM between two drains (8 accumulation steps)

enum GemmEpilogue : int {
  EPI_STORE = 0,       // C = acc
  EPI_BIAS = 1,        // C = acc + bias[n]
  EPI_BIAS_LRELU = 2,  // C = leaky_relu(acc + bias[n], slope)      (inference, BatchNorm folded)
};

struct alignas(128) GemmProblem {
  CUtensorMap tmA;     // 128 B each; split: the hi planes
  CUtensorMap tmB;
  CUtensorMap tmA_lo;  // split only
  CUtensorMap tmB_lo;
  CUtensorMap tmC;     // persistent inference GEMM only: 32 x 32 output boxes, SWIZZLE_128B (TMA store epilogue)
  float* C;
  const float* bias;
  int M, N, K, ldc;
  int bn;          // N tile (32, 64 or 128); split problems issue UMMA N = 2 bn
  int a_mn, b_mn;  // 0 = K-major operand, 1 = MN-major operand
  int epi;
  int tiles_m, tiles_n, tile_base;
  int accumulate;  // C += result (one CTA owns the tile: no atomics)
  float slope;
  int split;       // 1: error-compensated 3xTF32 on pre-split hi/lo planes, see the header comment
  // persistent inference GEMM only (gemm_persistent.cuh):
  int f16_ops;     // A and B are fp16 tensors (kind::f16, 64 elements per k-block) instead of fp32 read as TF32
  int out_f16;     // C is an fp16 tensor (the next layer's operand): 64 x 32 output boxes
};

// First CTA of every problem of a launch, passed BY VALUE (constant bank): the CTA -> problem lookup costs no dependent
// global loads (the wgrad launch has 12 problems).
constexpr int GEMM_MAX_PROBS = 16;
struct GemmBases { int base[GEMM_MAX_PROBS]; };

struct GemmCtrl {
  uint64_t full[GEMM_MAX_STAGES];
  uint64_t empty[GEMM_MAX_STAGES];
  uint64_t tmem_full;
  uint64_t acc_full[2];    // split: a k-block's worth of hi*hi products is complete in accumulator buffer b
  uint64_t acc_empty[2];   // split: the epilogue warps have drained buffer b (4 arrivals)
  uint32_t tmem_base;
};

__device__ __forceinline__ float leaky(float x, float slope) { return x > 0.f ? x : x * slope; }

__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tf32_grouped_kernel(const GemmProblem* __restrict__ probs, int nprobs, int ck, const GemmBases bases) {
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
    while (p + 1 < nprobs && tile >= bases.base[p + 1]) ++p;
  }
  const GemmProblem& P = probs[p];
  // Everything the pipelines need is read ONCE into registers: the inline-asm barriers below carry "memory"
  // clobbers, so any P.field left inside a loop would be re-fetched from global memory on every iteration.
  // Split-K: a cluster of ck CTAs (launch attribute) owns one output tile; rank r accumulates its share of the
  // k-blocks and the partial tiles are reduced through distributed shared memory (see the epilogue). tile_base counts
  // CTAs, a multiple of ck for every problem, so blockIdx.x % ck is the cluster rank.
  const int cta = blockIdx.x - bases.base[p];
  const int crank = cta % ck;
  const int t = cta / ck;
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
  const int split = P.split;
  const int kb_total = (P.K + GEMM_BK - 1) / GEMM_BK;   // the host guarantees kb_total >= ck
  const int kb_begin = crank * (kb_total / ck) + (crank < kb_total % ck ? crank : kb_total % ck);
  const int num_kb = kb_total / ck + (crank < kb_total % ck ? 1 : 0);
  const int b_bytes = bn * GEMM_BK * 4;
  const int kb_bytes = (GEMM_A_STAGE_BYTES + b_bytes) * (split ? 2 : 1);   // one ring slot = one k-block
  int nstages = GEMM_TILE_SMEM / kb_bytes;
  if (nstages > GEMM_MAX_STAGES) nstages = GEMM_MAX_STAGES;
  // split: [big0 | small0 | big1 | small1], bn columns each (see the MMA issuer); bn in {32, 64}: a power of two >= 32
  const uint32_t tmem_cols = static_cast<uint32_t>(split ? 4 * bn : bn);
  // offsets inside a ring slot
  const int off_alo = GEMM_A_STAGE_BYTES;                        // split only
  const int off_b = split ? 2 * GEMM_A_STAGE_BYTES : GEMM_A_STAGE_BYTES;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.tmA);
    tma_prefetch_desc(&P.tmB);
    if (split) { tma_prefetch_desc(&P.tmA_lo); tma_prefetch_desc(&P.tmB_lo); }
    for (int s = 0; s < nstages; ++s) {
      mbar_init(&ctrl->full[s], 1);
      mbar_init(&ctrl->empty[s], 1);
    }
    mbar_init(&ctrl->tmem_full, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&ctrl->acc_full[b], 1);
      mbar_init(&ctrl->acc_empty[b], 4);
    }
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
  // Everything above touched only this kernel's own parameters (the problem table and tensor maps are written by the
  // host, never by a kernel of the step): with programmatic dependent launch it overlaps the previous kernel's tail.
  grid_dep_wait();
  grid_dep_launch();

  // running sum of the drained big buffers (split): bn <= 64 columns of this thread's row; epilogue warps only
  float run[2][32];

  if (warp == 0) {
    // ------------------------------------------------ TMA producer (whole warp loops, one elected lane issues)
    const CUtensorMap* const tmA = &P.tmA;
    const CUtensorMap* const tmB = &P.tmB;
    const CUtensorMap* const tmAl = &P.tmA_lo;
    const CUtensorMap* const tmBl = &P.tmB_lo;
    int s = 0;          // ring slot and its phase parity, advanced incrementally (no integer divisions in the loop)
    uint32_t ph = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(&ctrl->empty[s], ph ^ 1);
      if (elect_one()) {
        uint64_t* bar = &ctrl->full[s];
        mbar_arrive_expect_tx(bar, static_cast<uint32_t>(kb_bytes));
        uint8_t* sa = tiles + s * kb_bytes;
        uint8_t* sb = sa + off_b;
        const int k0 = (kb_begin + kb) * GEMM_BK;
        if (!a_mn) {
          tma_load_2d(sa, tmA, bar, k0, m0);  // box {32 k, 128 rows}
          if (split) tma_load_2d(sa + off_alo, tmAl, bar, k0, m0);
        } else {
          for (int i = 0; i < GEMM_BM / 32; ++i) {  // box {32 rows(contiguous), 32 k}
            tma_load_2d(sa + i * 4096, tmA, bar, m0 + 32 * i, k0);
            if (split) tma_load_2d(sa + off_alo + i * 4096, tmAl, bar, m0 + 32 * i, k0);
          }
        }
        if (!b_mn) {
          tma_load_2d(sb, tmB, bar, k0, n0);  // box {32 k, bn rows}
          if (split) tma_load_2d(sb + b_bytes, tmBl, bar, k0, n0);
        } else {
          for (int i = 0; i < bn / 32; ++i) {
            tma_load_2d(sb + i * 4096, tmB, bar, n0 + 32 * i, k0);
            if (split) tma_load_2d(sb + b_bytes + i * 4096, tmBl, bar, n0 + 32 * i, k0);
          }
        }
      }
      __syncwarp();
      if (++s == nstages) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (whole warp loops, one elected lane issues)
    const uint32_t idesc = umma_idesc_tf32(GEMM_BM, bn, a_mn, b_mn);
    const uint32_t idesc2 = umma_idesc_tf32(GEMM_BM, 2 * bn, a_mn, b_mn);   // A_hi x [B_hi ; B_lo]
    const uint32_t a_step16 = a_mn ? 64u : 2u;  // descriptor start-address units (16 B) per UMMA_K step
    const uint32_t b_step16 = b_mn ? 64u : 2u;
    // descriptor bits that do not change across the k loop (everything but the start address)
    const uint64_t da_hi = umma_smem_desc(0u, a_mn ? 4096u : 16u, a_mn ? 512u : 1024u, a_mn ? 1u : 2u);
    const uint64_t db_hi = umma_smem_desc(0u, b_mn ? 4096u : 16u, b_mn ? 512u : 1024u, b_mn ? 1u : 2u);
    const uint32_t tiles_u32 = smem_u32(tiles);
    const uint32_t alo16 = static_cast<uint32_t>(off_alo) >> 4;
    int s = 0;
    uint32_t ph = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(&ctrl->full[s], ph);
      const int chunk = kb / GEMM_DRAIN_KB;
      const bool chunk_start = kb % GEMM_DRAIN_KB == 0;
      if (split && chunk_start && chunk >= 2) mbar_wait(&ctrl->acc_empty[chunk & 1], ((chunk >> 1) - 1) & 1);   // buffer drained
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = tiles_u32 + s * kb_bytes;   // 1024-aligned: (addr >> 4) + k*step never carries out of
        const uint32_t sb = sa + off_b;                 // the 14-bit start-address field
        const uint64_t da0 = da_hi | static_cast<uint64_t>((sa >> 4) & 0x3FFFu);
        const uint64_t db0 = db_hi | static_cast<uint64_t>((sb >> 4) & 0x3FFFu);
        if (split) {
          // The tensor core adds into its fp32 accumulator with truncation, a bias that grows with the number of
          // accumulation steps (measured: 2.4e-6 relative at K = 512, 9e-6 at K = 2000). So the dominant hi*hi sum is
          // accumulated in TMEM for a chunk of GEMM_DRAIN_KB k-blocks (8 steps) only: "big" buffer b = chunk & 1 is
          // overwritten at the start of every chunk and drained by the epilogue warps, which keep the running sum in
          // registers with round-to-nearest fp32 adds (measured: 1.1-1.8e-7 relative for K = 32 .. 2000). The
          // correction terms (2^-11 smaller, truncation irrelevant) accumulate over the whole K in the "small" buffer
          // next to it.
          const uint32_t big = tmem_d + static_cast<uint32_t>((chunk & 1) * 2 * bn), small = big + static_cast<uint32_t>(bn);
          const uint32_t blo16 = static_cast<uint32_t>(b_bytes) >> 4;
#pragma unroll
          for (int k = 0; k < GEMM_BK / GEMM_UMMA_K; ++k) {
            const uint64_t da = da0 + k * a_step16, db = db0 + k * b_step16;
            if (k == 0 && chunk_start && chunk >= 2) {
              umma_tf32(big, da, db, idesc, 0u);                          // hi(A) * hi(B): restart the big buffer
              umma_tf32(small, da, db + blo16, idesc, 1u);                // hi(A) * lo(B)
            } else {
              // hi(A) * [hi(B) ; lo(B)] -> big | small in one N = 2 bn instruction (first use of a buffer: overwrite)
              umma_tf32(big, da, db, idesc2, (k != 0 || !chunk_start) ? 1u : 0u);
            }
            umma_tf32(small, da + alo16, db, idesc, 1u);                  // lo(A) * hi(B)
          }
          umma_commit(&ctrl->empty[s]);             // frees the ring slot when these MMAs have read it
          if (kb % GEMM_DRAIN_KB == GEMM_DRAIN_KB - 1 || kb == num_kb - 1)
            umma_commit(&ctrl->acc_full[chunk & 1]);   // ... and hands the big buffer to the epilogue warps
        } else {
#pragma unroll
          for (int k = 0; k < GEMM_BK / GEMM_UMMA_K; ++k)
            umma_tf32(tmem_d, da0 + k * a_step16, db0 + k * b_step16, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        if (!split) umma_commit(&ctrl->empty[s]);  // frees the ring slot when these MMAs have read it
      }
      __syncwarp();
      if (++s == nstages) { s = 0; ph ^= 1; }
    }
    if (!split && elect_one()) umma_commit(&ctrl->tmem_full);  // accumulator complete
    __syncwarp();
  } else {
    // ------------------------------------------------ epilogue warps, phase 1: wait for (and, split, drain) the MMAs
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const uint32_t lane_base = tmem_d + (static_cast<uint32_t>(q * 32) << 16);
    if (split) {
#pragma unroll
      for (int j = 0; j < 32; ++j) { run[0][j] = 0.f; run[1][j] = 0.f; }
      const int num_chunks = (num_kb + GEMM_DRAIN_KB - 1) / GEMM_DRAIN_KB;
      for (int c = 0; c < num_chunks; ++c) {
        mbar_wait(&ctrl->acc_full[c & 1], (c >> 1) & 1);
        tc_fence_after();
        const uint32_t big = lane_base + static_cast<uint32_t>((c & 1) * 2 * bn);
        float v[32], w[32];
        tmem_ld_32x32(big, v);
        if (bn > 32) tmem_ld_32x32(big + 32u, w);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctrl->acc_empty[c & 1]);   // values are in registers: the buffer may be overwritten
#pragma unroll
        for (int j = 0; j < 32; ++j) run[0][j] += v[j];
        if (bn > 32) {
#pragma unroll
          for (int j = 0; j < 32; ++j) run[1][j] += w[j];
        }
      }
      // the last acc_full commit covers every MMA of the tile: the small buffers are final as well
    } else {
      mbar_wait(&ctrl->tmem_full, 0);
      tc_fence_after();
    }
  }

  // Split-K, barrier A: every CTA of the cluster has finished reading its operand ring (the epilogue warps arrive after
  // the last MMA has completed), so a mate may now deposit its partial tile into it.
  if (ck > 1) cluster_sync_all();

  if (warp >= 2) {
    // ------------------------------------------------ epilogue warps, phase 2: TMEM -> registers -> (reduce) -> global
    const int q = warp & 3;
    const uint32_t lane_base = tmem_d + (static_cast<uint32_t>(q * 32) << 16);
    float* st = epi_stage + q * 32 * GEMM_EPI_PITCH;
    // this thread's row of the CTA's partial tile, columns [c0, c0 + 32)
    auto load_chunk = [&](int c0, float (&v)[32]) {
      const uint32_t taddr = lane_base + static_cast<uint32_t>(c0);
      if (split) {
        tmem_ld_32x32(taddr + static_cast<uint32_t>(bn), v);   // small0: correction terms of the even chunks
        tmem_ld_wait();
        if (num_kb > GEMM_DRAIN_KB) {   // more than one chunk: buffer 1 was used
          float w[32];
          tmem_ld_32x32(taddr + static_cast<uint32_t>(3 * bn), w);   // small1: odd chunks
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += w[j];
        }
        if (c0 == 0) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += run[0][j];
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += run[1][j];
        }
      } else {
        tmem_ld_32x32(taddr, v);
        tmem_ld_wait();
      }
    };
    // Split-K reduce-scatter: the 32-row slab of epilogue warp q is finished by cluster rank owner = q * ck / 4; the
    // other ranks deposit their partial slab in the owner's (now idle) operand ring: slot (sender rank, skipping the
    // owner), rows padded to bn + 4 floats (conflict-free 128-bit accesses on both sides).
    const int owner = (q * ck) >> 2;
    const int pitch = bn + 4;
    const int slab_floats = 32 * pitch;
    const int ql = q - ((owner * 4) / ck);                 // slab index among the slabs this owner finishes
    const int slabs_per_owner = 4 / ck;
    if (ck > 1 && owner != crank) {
      const int slot = crank < owner ? crank : crank - 1;
      float* dst_local = reinterpret_cast<float*>(tiles) + (slot * slabs_per_owner + ql) * slab_floats + lane * pitch;
      const uint32_t dst = mapa_shared(smem_u32(dst_local), static_cast<uint32_t>(owner));
      for (int c0 = 0; c0 < bn; c0 += 32) {
        if (n0 + c0 >= pN) break;  // warp-uniform
        float v[32];
        load_chunk(c0, v);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          st_shared_cluster_f4(dst + static_cast<uint32_t>((c0 + 4 * j) * 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
    }
    if (ck > 1) cluster_sync_all();   // barrier B: the deposits are visible to their owners
    if (ck == 1 || owner == crank) {
      const bool vec_ok = (ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(pC) & 15) == 0;
      const int rsub = lane >> 3, ch = lane & 7;  // read-back mapping: 4 rows x 8 float4 per pass
      for (int c0 = 0; c0 < bn; c0 += 32) {
        const int nbase = n0 + c0;
        if (nbase >= pN) break;  // warp-uniform
        float v[32];
        load_chunk(c0, v);
        for (int sl = 0; sl < ck - 1; ++sl) {   // fixed order: ascending sender rank
          const float* src = reinterpret_cast<const float*>(tiles) + (sl * slabs_per_owner + ql) * slab_floats + lane * pitch + c0;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 x = *reinterpret_cast<const float4*>(src + 4 * j);
            v[4 * j] += x.x; v[4 * j + 1] += x.y; v[4 * j + 2] += x.z; v[4 * j + 3] += x.w;
          }
        }
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
  } else if (ck > 1) {
    cluster_sync_all();   // barrier B (producer and MMA warps only take part in the cluster barriers)
  }
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
// Same for B with N. split: A/B are the hi planes and A_lo/B_lo the lo planes (same shape and pitch).
// dtype_tf32: single-pass problems on raw fp32 arrays let TMA round to TF32 on load. Returns 0 on success.
inline int gemm_problem_fill(GemmProblem* g, const float* A, int lda, int a_mn, const float* B, int ldb, int b_mn,
                             float* C, int ldc, int M, int N, int K, int bn, int epi, const float* bias,
                             float slope, int accumulate, int dtype_tf32 = 1, const float* A_lo = nullptr,
                             const float* B_lo = nullptr) {
  *g = GemmProblem{};
  const int split = (A_lo && B_lo) ? 1 : 0;
  if (split) dtype_tf32 = 0;   // the planes are already TF32-representable
  auto mk = [&](CUtensorMap* tm, const float* base, int rows, int ld, int mn, int box_rows) {
    if (!mn) return make_tmap_2d(tm, base, K, rows, ld, GEMM_BK, box_rows, dtype_tf32);
    return make_tmap_2d(tm, base, rows, K, ld, 32, GEMM_BK, dtype_tf32, 1);
  };
  int rc;
  if ((rc = mk(&g->tmA, A, M, lda, a_mn, GEMM_BM))) return rc;
  if ((rc = mk(&g->tmB, B, N, ldb, b_mn, bn))) return rc;
  if (split) {
    if ((rc = mk(&g->tmA_lo, A_lo, M, lda, a_mn, GEMM_BM))) return rc;
    if ((rc = mk(&g->tmB_lo, B_lo, N, ldb, b_mn, bn))) return rc;
  }
  g->C = C; g->bias = bias;
  g->M = M; g->N = N; g->K = K; g->ldc = ldc;
  g->bn = bn; g->a_mn = a_mn; g->b_mn = b_mn; g->epi = epi;
  g->tiles_m = (M + GEMM_BM - 1) / GEMM_BM;
  g->tiles_n = (N + bn - 1) / bn;
  g->tile_base = 0;
  g->slope = slope;
  g->accumulate = accumulate;
  g->split = split;
  if (split && bn > 64) return -2;   // the split epilogue keeps bn <= 64 running sums in registers
  return 0;
}

// Split-K factor of a stage: the largest ck in {4, 2, 1} that keeps the launch within one wave of `sms` CTAs and leaves
// every cluster rank at least `min_kb_per_rank` k-blocks.
inline int gemm_pick_splitk(const GemmProblem* g, int n, int sms = 148, int min_kb_per_rank = 4) {
  int tiles = 0, min_kb = 1 << 30;
  for (int i = 0; i < n; ++i) {
    tiles += g[i].tiles_m * g[i].tiles_n;
    const int kb = (g[i].K + GEMM_BK - 1) / GEMM_BK;
    if (kb < min_kb) min_kb = kb;
  }
  for (int ck = 4; ck > 1; ck >>= 1)
    if (tiles * ck <= sms && min_kb >= min_kb_per_rank * ck) return ck;
  return 1;
}

// Assign CTA ranges (ck CTAs per tile); returns the grid size.
inline int gemm_table_finalize(GemmProblem* g, int n, int ck = 1) {
  int base = 0;
  for (int i = 0; i < n; ++i) {
    g[i].tile_base = base;
    base += g[i].tiles_m * g[i].tiles_n * ck;
  }
  return base;
}

// use_pdl: launch with programmatic stream serialization (the kernel calls griddepcontrol.wait itself), so that its
// prologue (barrier init, TMEM allocation, tensor-map prefetch) overlaps the tail of the previous kernel in the stream.
// total_ctas = gemm_table_finalize(..., ck). ck > 1 launches clusters of ck CTAs (split-K, see the kernel).
// host_table: the host copy of the same table entries (tile_base of every problem), or null for a single problem.
inline cudaError_t gemm_launch(const GemmProblem* dev_table, int nprobs, int total_ctas, cudaStream_t st, bool use_pdl = false,
                               int ck = 1, const GemmProblem* host_table = nullptr) {
  if (nprobs > GEMM_MAX_PROBS || (nprobs > 1 && !host_table)) return cudaErrorInvalidValue;
  GemmBases bases{};
  for (int i = 0; i < nprobs; ++i) bases.base[i] = host_table ? host_table[i].tile_base : 0;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tf32_grouped_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         GEMM_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(total_ctas);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = GEMM_SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (use_pdl) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (ck > 1) {
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = static_cast<unsigned>(ck);
    at[na].val.clusterDim.y = 1;
    at[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = at;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, gemm_tf32_grouped_kernel, dev_table, nprobs, ck, bases);
}

}  // namespace jb
