// fp16-split GEMM of the training step on the 5th-gen tensor cores (tcgen05.mma kind::f16, fp32 accumulators in TMEM,
// operands staged in shared memory by TMA), written as DEVICE FUNCTIONS with persistent pipeline state so that the same
// code runs inside the whole-step kernel (stepk.cuh: one CTA per SM walks the tiles of every GEMM phase; barriers, TMEM
// and the operand ring live for the whole kernel) and inside a plain one-phase kernel (tools/hgemm_lab.cu).
//
//   C[M,N] (+)= out_scale * (A * B^T) (+ bias)      A logically [M,K], B logically [N,K]
//
// Operand format (round 2; round 1 shipped TF32 hi/lo planes in fp32 containers = 8 B per element through L2 -> SM,
// measured to be THE bound of every GEMM of the step, profiles/ncu_gemm_r1_s32.md): every GEMM operand x is stored as
// two fp16 planes
//       hi = fp16(x),   lo = fp16((x - hi) * 2^11)
// so |x - hi - lo 2^-11| <= 2^-22 |x| (same 11 + 11 significand bits as the TF32 split) at 4 B per element, and the
// products run on kind::f16 (twice the TF32 rate, K = 16 per instruction). fp16 range (6e-5 .. 65504 normal) is kept by
// construction: activations are O(1) after BatchNorm, weights O(1/sqrt(fan_in)), and the backward pass carries a
// power-of-two loss scale (StepConsts::gscale) that the gradient epilogues divide out exactly.
//
// Modes (HgProblem::mode):
//   HG_SINGLE  one pass on the hi planes (11-bit operands, like one TF32 pass)
//   HG_PRECISE fp32-class: D = Ah Bh + 2^-11 (Ah Bl + Al Bh). The dominant Ah Bh sum is accumulated in TMEM for chunks of
//              HG_DRAIN_KB k-blocks only ("big" buffer b = chunk & 1, drained by the epilogue warps into registers with
//              round-to-nearest adds) because the tensor core adds into its accumulator with truncation (round 1
//              measured a bias of 2.4e-6 at K = 512, enough to flip LeakyReLU decisions); the cross terms accumulate
//              in the "small" buffer next to it. bn <= 64. Forward and dgrad GEMMs.
//   HG_MEDIUM  three passes without drains (big | small over the whole K): ~1e-6, bn <= 256. Weight gradients.
// Operand majors: K-major (row-major [rows, K]) or MN-major (stored [K, rows]): forward (K,K), dgrad (K,MN) and wgrad
// (MN,MN) read the same activation / weight buffers, nothing is transposed in memory.
//   K-major : TMA SWIZZLE_128B box {64 k, rows};   UMMA SWIZZLE_128B, SBO 1024, +32 B per K = 16 step
//   MN-major: TMA SWIZZLE_128B boxes {64 mn, 64 k} (8 KB each); UMMA SWIZZLE_128B, LBO 8192 (next 64 mn), SBO 1024
//             (next 8 k rows), +2048 B per K = 16 step
// Split-K: HgProblem::ksplit partial results go to C + p * part_stride; the CONSUMER phase (BatchNorm slab, latent,
// reconstruction ...) sums them in a fixed order, so there is no DSMEM / atomic reduction and the result is deterministic.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include "ptx.cuh"
#include "gemm_tf32.cuh"   // tmap_encode_fn, GemmEpilogue, leaky

namespace jb {

constexpr int HG_BM = 128;
constexpr int HG_BK = 64;                       // fp16 elements per k-block = one 128-byte swizzle row
constexpr int HG_UMMA_K = 16;
constexpr int HG_A_BYTES = HG_BM * HG_BK * 2;   // 16 KB per plane
constexpr int HG_RING_BYTES = 192 * 1024;
constexpr int HG_CTRL_BYTES = 2048;                // barriers, per-phase first tiles, step scalars | 1 KB of dropout keep bits (stepk.cuh)
constexpr int HG_NEPI = 8;                      // epilogue warps (two per TMEM lane quadrant)
constexpr int HG_STAGE_BYTES = HG_NEPI * 4096;  // one swizzled 32 x 32 fp32 block per epilogue warp
constexpr int HG_SMEM_BYTES = HG_CTRL_BYTES + HG_RING_BYTES + HG_STAGE_BYTES + 1024;   // + alignment slack
constexpr int HG_MAX_STAGES = 8;
constexpr int HG_DRAIN_KB = 2;                  // precise: k-blocks (8 accumulation steps) between two drains
constexpr int HG_TMEM_COLS = 512;
constexpr float HG_LO_SCALE = 2048.f;           // lo planes hold (x - hi) * 2^11
constexpr float HG_LO_INV = 1.f / 2048.f;
constexpr int HG_WARP_TMA = 0, HG_WARP_MMA = 1, HG_WARP_EPI0 = 4;   // warp roles inside a CTA of >= 12 warps
constexpr int HG_MAX_PROBS = 16;
constexpr int HG_CLUSTER = 4;                   // CTAs per cluster of the step kernel (fused problems: one M tile each)

enum HgMode : int { HG_SINGLE = 0, HG_PRECISE = 1, HG_MEDIUM = 2 };

struct alignas(128) HgProblem {
  CUtensorMap tmA_hi, tmA_lo, tmB_hi, tmB_lo;
  float* C;
  const float* bias;
  long long part_stride;   // floats between split-K partial outputs
  int M, N, K, ldc;
  int bn;                  // N tile: 32 / 64 (precise), 32 .. 256 (single, medium)
  int a_mn, b_mn;          // 0 = K-major operand, 1 = MN-major operand
  int epi;                 // GemmEpilogue (bias is added by split-K partial 0 only)
  int mode;                // HgMode
  int ksplit;              // >= 1
  int tiles_m, tiles_n, tile_base;   // CTA work items of this problem: [tile_base, tile_base + tiles_m tiles_n ksplit)
  int accumulate;          // C += result (one work item owns the tile: no atomics)
  const int* acc_flag;     // optional device flag: accumulate if *acc_flag != 0 (gradient accumulation over batches)
  float out_scale;         // exact power of two (1 / loss scale for weight gradients)
  float slope;
  const float* dyn_scale;  // optional device scalar multiplied into out_scale (inverse of a dynamic operand scale)
  // Cluster-fused epilogue (stepk.cuh): the work items of the problem are (column block, M tile) with the M tile index
  // fastest and padded to HG_CLUSTER, so that the HG_CLUSTER CTAs of one thread-block cluster hold all rows of a column
  // block (CTA rank in the cluster = M tile). The accumulators are not stored: they are left in the staging blocks for
  // the caller's tail (BatchNorm statistics over the cluster, activation, dropout, loss ...). Requires tiles_m <=
  // HG_CLUSTER, bn <= 64, ksplit = 1.
  float* norm_out;         // optional (weight gradients): sum of squares of the values this work item stored, one partial per
                           // epilogue warp at norm_out[item * HG_NEPI + e] (the clip norm without a sweep over the gradient buffer)
  int fuse;                // 0 = plain epilogue; otherwise the kind of tail (StepFuse)
  int fuse_arg;            // which layer / modality the tail works on
  int fuse_ks;             // fused problem with ONE column block whose K range is split over the HG_CLUSTER ranks instead
                           // (ksplit = HG_CLUSTER; work items M tile major, rank = K part): the tail sums the four staged
                           // partial tiles through distributed shared memory. For GEMMs with a handful of M tiles.
};

struct HgPhase {           // one GEMM phase = a table of problems
  int first, count, total_tiles;
  int base[HG_MAX_PROBS];
};

struct HgCtrl {
  uint64_t full[HG_MAX_STAGES];
  uint64_t empty[HG_MAX_STAGES];
  uint64_t accf[2];        // MMA -> epilogue: a chunk (precise) / the tile (other modes) is complete in TMEM
  uint64_t acce[2];        // epilogue -> MMA: the buffer has been read out of TMEM (HG_NEPI arrivals)
  uint32_t tmem_base;
};

// Per-thread pipeline state that survives from tile to tile and from phase to phase (every role walks the same tile
// sequence, so the counters of the roles agree without communication).
struct HgPipe {
  int s = 0;               // next ring slot
  uint32_t par = 0;        // bit i: parity of the number of fills of slot i (per-slot, because the slot count varies)
  int geom = 0;            // slot size of the previous tile; the ring is re-cut (and drained) when it changes
  uint32_t nf0 = 0, nf1 = 0;   // completed uses of accumulator buffer 0 / 1
  long long* dbg = nullptr;    // profiling: SM clock stamps of this CTA's first tile of the phase (8 slots) or null
};
__device__ __forceinline__ void hg_stamp(const HgPipe& pp, int k) { if (pp.dbg != nullptr) pp.dbg[k] = clock64(); }

// fp16 split of one value (producers of GEMM operands call this)
// Conversions SATURATE to +-65504 instead of overflowing to inf (cvt.rn.satfinite: same single instruction). Operand
// magnitudes are bounded by construction or by the dynamic scales, except in blow-ups (measured on the noise benchmark:
// one latent row at 1e16 crushes the other rows' planes to zero, a BatchNorm layer sees zero variance and amplifies its
// gradient 316x past fp16's range): an inf there became NaN in the lo plane and destroyed every parameter, where the fp32
// reference takes one clipped step and carries on. Saturated values are wrong but finite, and the global-norm clip
// that such a step always triggers scales them away.
__device__ __forceinline__ __half h_sat(float x) {
  unsigned short r;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(x));
  return __ushort_as_half(r);
}
__device__ __forceinline__ __half2 h2_sat(float a, float b) {   // .x = a, .y = b
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return *reinterpret_cast<__half2*>(&r);
}
__device__ __forceinline__ void h_split(float x, __half& hi, __half& lo) {
  hi = h_sat(x);
  lo = h_sat((x - __half2float(hi)) * HG_LO_SCALE);
}
__device__ __forceinline__ float h_join(__half hi, __half lo) { return __half2float(hi) + __half2float(lo) * HG_LO_INV; }

// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16 (fp16 operands, fp32 accumulate), issued by ONE thread.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Instruction descriptor, kind::f16 with F16 operands, fp32 accumulate, M = 128 (formats: a/b 0 = F16, c 1 = F32).
__host__ __device__ __forceinline__ uint32_t umma_idesc_f16(int m, int n, int a_mn, int b_mn) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= static_cast<uint32_t>(a_mn & 1) << 15;
  d |= static_cast<uint32_t>(b_mn & 1) << 16;
  d |= static_cast<uint32_t>(n >> 3) << 17;
  d |= static_cast<uint32_t>(m >> 4) << 24;
  return d;
}

// ------------------------------------------------------------------------------------------------ setup / teardown
// Called by every thread of the CTA once per kernel (contains __syncthreads).
__device__ __forceinline__ uint32_t hg_setup(HgCtrl* ctrl, int warp, int lane) {
  if (warp == HG_WARP_TMA && lane == 0) {
    for (int s = 0; s < HG_MAX_STAGES; ++s) {
      mbar_init(&ctrl->full[s], 1);
      mbar_init(&ctrl->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&ctrl->accf[b], 1);
      mbar_init(&ctrl->acce[b], HG_NEPI);
    }
    fence_mbar_init();
  }
  if (warp == HG_WARP_MMA) {
    tmem_alloc(&ctrl->tmem_base, HG_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return ctrl->tmem_base;
}
__device__ __forceinline__ void hg_teardown(uint32_t tmem_d, int warp) {
  tc_fence_before();
  __syncthreads();
  if (warp == HG_WARP_MMA) tmem_dealloc(tmem_d, HG_TMEM_COLS);
}

// geometry of one work item, derived identically by every role
struct HgTile {
  int p, m0, n0, ks, kb_begin, num_kb, bn, b_bytes, slot_bytes, nstages, planes;
  int lt;                  // index of the work item inside its problem
  int null;                // fused problems: this CTA's M tile lies beyond M (no GEMM work, the CTA still joins the tail)
};
__device__ __forceinline__ HgTile hg_decode(const HgProblem* __restrict__ probs, const HgPhase& ph, int t) {
  HgTile T;
  int p = 0;
  while (p + 1 < ph.count && t >= ph.base[p + 1]) ++p;
  const HgProblem& P = probs[ph.first + p];
  T.p = ph.first + p;
  const int local = t - ph.base[p];
  T.lt = local;
  const int ksplit = P.ksplit;
  T.ks = local % ksplit;
  const int tt = local / ksplit;
  const int tiles_n = P.tiles_n;
  T.bn = P.bn;
  if (P.fuse && !P.fuse_ks) {
    T.m0 = (tt % HG_CLUSTER) * HG_BM;
    T.n0 = (tt / HG_CLUSTER) * T.bn;
  } else {
    T.m0 = (tt / tiles_n) * HG_BM;
    T.n0 = (tt % tiles_n) * T.bn;
  }
  T.null = T.m0 >= P.M ? 1 : 0;
  const int kb_total = (P.K + HG_BK - 1) / HG_BK;   // host guarantees kb_total >= ksplit
  T.kb_begin = T.ks * (kb_total / ksplit) + (T.ks < kb_total % ksplit ? T.ks : kb_total % ksplit);
  T.num_kb = kb_total / ksplit + (T.ks < kb_total % ksplit ? 1 : 0);
  const int brows = P.b_mn ? ((T.bn + 63) & ~63) : T.bn;   // MN-major operands arrive in 64-wide boxes
  T.b_bytes = brows * HG_BK * 2;
  T.planes = P.mode == HG_SINGLE ? 1 : 2;
  T.slot_bytes = (HG_A_BYTES + T.b_bytes) * T.planes;
  int ns = HG_RING_BYTES / T.slot_bytes;
  T.nstages = ns > HG_MAX_STAGES ? HG_MAX_STAGES : ns;
  return T;
}
// The ring is re-cut when the slot size changes (it depends on bn and mode): both roles restart at slot 0, and the
// producer first waits until every slot of the old geometry has been released (the new slots overlap other old slots).

// ------------------------------------------------------------------------------------------------ TMA producer warp
// The three roles below process ONE work item; hg_run_phase walks the items of the calling CTA (every role walks the
// same sequence, so the pipeline counters in HgPipe agree without communication).
__device__ __forceinline__ void hg_produce_tile(const HgProblem& P, const HgTile& T, HgCtrl* ctrl, uint8_t* ring, HgPipe& pp_io,
                                             bool stamp) {   // stamp: the CTA's first work item of the phase
  HgPipe pp = pp_io;   // the roles are separate functions (own register allocation); the pipeline state travels by value
  const int a_mn = P.a_mn, b_mn = P.b_mn;
  const CUtensorMap* const tmAh = &P.tmA_hi;
  const CUtensorMap* const tmAl = &P.tmA_lo;
  const CUtensorMap* const tmBh = &P.tmB_hi;
  const CUtensorMap* const tmBl = &P.tmB_lo;
  const bool two = T.planes == 2;
  const int off_b = two ? 2 * HG_A_BYTES : HG_A_BYTES;
  if (T.slot_bytes != pp.geom) {
    // the ring is re-cut: every slot of the old geometry must have been released. At the first work item of a phase that
    // is known (the previous phase ended with its accumulators complete, then a grid barrier); the eight waits cost 0.7 us
    if (!stamp)
      for (int i = 0; i < HG_MAX_STAGES; ++i) mbar_wait(&ctrl->empty[i], ((pp.par >> i) & 1u) ^ 1u);
    pp.s = 0;
    pp.geom = T.slot_bytes;
  }
  for (int kb = 0; kb < T.num_kb; ++kb) {
    mbar_wait(&ctrl->empty[pp.s], ((pp.par >> pp.s) & 1u) ^ 1u);
    if (elect_one()) {
      uint64_t* bar = &ctrl->full[pp.s];
      mbar_arrive_expect_tx(bar, static_cast<uint32_t>(T.slot_bytes));
      uint8_t* sa = ring + pp.s * T.slot_bytes;
      uint8_t* sb = sa + off_b;
      const int k0 = (T.kb_begin + kb) * HG_BK;
      if (!a_mn) {
        tma_load_2d(sa, tmAh, bar, k0, T.m0);                 // box {64 k, 128 rows}
        if (two) tma_load_2d(sa + HG_A_BYTES, tmAl, bar, k0, T.m0);
      } else {
#pragma unroll
        for (int i = 0; i < HG_BM / 64; ++i) {                // boxes {64 rows (contiguous), 64 k}
          tma_load_2d(sa + i * 8192, tmAh, bar, T.m0 + 64 * i, k0);
          if (two) tma_load_2d(sa + HG_A_BYTES + i * 8192, tmAl, bar, T.m0 + 64 * i, k0);
        }
      }
      if (!b_mn) {
        tma_load_2d(sb, tmBh, bar, k0, T.n0);                 // box {64 k, bn rows}
        if (two) tma_load_2d(sb + T.b_bytes, tmBl, bar, k0, T.n0);
      } else {
        for (int i = 0; i < (T.bn + 63) / 64; ++i) {
          tma_load_2d(sb + i * 8192, tmBh, bar, T.n0 + 64 * i, k0);
          if (two) tma_load_2d(sb + T.b_bytes + i * 8192, tmBl, bar, T.n0 + 64 * i, k0);
        }
      }
    }
    __syncwarp();
    if (stamp && lane_id() == 0) { if (kb == 0) hg_stamp(pp, 1); if (kb == T.num_kb - 1) hg_stamp(pp, 2); }
    pp.par ^= 1u << pp.s;
    if (++pp.s == T.nstages) pp.s = 0;
  }
  pp_io = pp;
}

// ------------------------------------------------------------------------------------------------ MMA issuer warp
__device__ __forceinline__ void hg_mma_tile(const HgProblem& P, const HgTile& T, HgCtrl* ctrl, uint8_t* ring, uint32_t tmem_d,
                                         HgPipe& pp_io, bool stamp) {
  HgPipe pp = pp_io;
  const uint32_t ring_u32 = smem_u32(ring);
  const int a_mn = P.a_mn, b_mn = P.b_mn, mode = P.mode, bn = T.bn;
  const uint32_t idesc = umma_idesc_f16(HG_BM, bn, a_mn, b_mn);
  const uint32_t idesc2 = umma_idesc_f16(HG_BM, 2 * bn, a_mn, b_mn);   // A_hi x [B_hi ; B_lo]
  // [B_hi ; B_lo] are adjacent N rows in shared memory when B is K-major, or whole 64-wide blocks when MN-major
  const bool cat = mode != HG_SINGLE && 2 * bn <= 256 && (!b_mn || (bn & 63) == 0);
  const uint32_t a_step = a_mn ? 128u : 2u;   // descriptor start-address units (16 B) per UMMA_K step
  const uint32_t b_step = b_mn ? 128u : 2u;
  const uint64_t da_hi = umma_smem_desc(0u, a_mn ? 8192u : 16u, 1024u, 2u);
  const uint64_t db_hi = umma_smem_desc(0u, b_mn ? 8192u : 16u, 1024u, 2u);
  const bool two = T.planes == 2;
  const int off_b = two ? 2 * HG_A_BYTES : HG_A_BYTES;
  const uint32_t alo16 = static_cast<uint32_t>(HG_A_BYTES) >> 4;
  const uint32_t blo16 = static_cast<uint32_t>(T.b_bytes) >> 4;
  if (T.slot_bytes != pp.geom) { pp.s = 0; pp.geom = T.slot_bytes; }
  if (mode != HG_PRECISE) {
    // the previous tile's accumulator (any mode uses columns from 0) must have been read by the epilogue warps
    if (pp.nf0 > 0) mbar_wait(&ctrl->acce[0], (pp.nf0 - 1) & 1);
    if (pp.nf1 > 0) mbar_wait(&ctrl->acce[1], (pp.nf1 - 1) & 1);
  }
  for (int kb = 0; kb < T.num_kb; ++kb) {
    const int chunk = kb / HG_DRAIN_KB;
    const bool chunk_start = kb % HG_DRAIN_KB == 0;
    const int buf = chunk & 1;
    if (mode == HG_PRECISE && chunk_start) {
      // buffer `buf` is about to be restarted: its previous use (this tile or an earlier one) must be drained; the
      // first chunks of a tile also wait for the OTHER buffer's previous tile (its small columns are restarted too)
      const uint32_t nf = buf ? pp.nf1 : pp.nf0;
      if (nf > 0) mbar_wait(&ctrl->acce[buf], (nf - 1) & 1);
    }
    mbar_wait(&ctrl->full[pp.s], (pp.par >> pp.s) & 1u);
    tc_fence_after();
    if (stamp && lane_id() == 0) { if (kb == 0) hg_stamp(pp, 3); if (kb == T.num_kb - 1) hg_stamp(pp, 4); }
    if (elect_one()) {
      const uint32_t sa = ring_u32 + pp.s * T.slot_bytes;
      const uint32_t sb = sa + off_b;
      const uint64_t da0 = da_hi | static_cast<uint64_t>((sa >> 4) & 0x3FFFu);
      const uint64_t db0 = db_hi | static_cast<uint64_t>((sb >> 4) & 0x3FFFu);
      if (mode == HG_PRECISE) {
        const uint32_t big = tmem_d + static_cast<uint32_t>(buf * 2 * bn), small = big + static_cast<uint32_t>(bn);
        const bool first_use = chunk < 2;   // first chunk of this tile in this buffer: the small columns restart too
#pragma unroll
        for (int k = 0; k < HG_BK / HG_UMMA_K; ++k) {
          const uint64_t da = da0 + k * a_step, db = db0 + k * b_step;
          const uint32_t acc_big = (k != 0 || !chunk_start) ? 1u : 0u;
          const uint32_t acc_small = (k != 0 || !chunk_start || !first_use) ? 1u : 0u;
          if (cat && acc_big == acc_small) {
            umma_f16(big, da, db, idesc2, acc_big);                    // hi(A) * [hi(B) ; lo(B)] -> big | small
          } else {
            umma_f16(big, da, db, idesc, acc_big);                     // hi(A) * hi(B)
            umma_f16(small, da, db + blo16, idesc, acc_small);         // hi(A) * lo(B)
          }
          umma_f16(small, da + alo16, db, idesc, 1u);                  // lo(A) * hi(B)
        }
        umma_commit(&ctrl->empty[pp.s]);
        if (kb % HG_DRAIN_KB == HG_DRAIN_KB - 1 || kb == T.num_kb - 1) {
          umma_commit(&ctrl->accf[buf]);                               // hands the big buffer to the epilogue warps
        }
      } else if (mode == HG_MEDIUM) {
        const uint32_t big = tmem_d, small = tmem_d + static_cast<uint32_t>(bn);
#pragma unroll
        for (int k = 0; k < HG_BK / HG_UMMA_K; ++k) {
          const uint64_t da = da0 + k * a_step, db = db0 + k * b_step;
          const uint32_t acc = (kb | k) != 0 ? 1u : 0u;
          if (cat) {
            umma_f16(big, da, db, idesc2, acc);
          } else {
            umma_f16(big, da, db, idesc, acc);
            umma_f16(small, da, db + blo16, idesc, acc);
          }
          umma_f16(small, da + alo16, db, idesc, 1u);
        }
        umma_commit(&ctrl->empty[pp.s]);
        if (kb == T.num_kb - 1) umma_commit(&ctrl->accf[0]);
      } else {
#pragma unroll
        for (int k = 0; k < HG_BK / HG_UMMA_K; ++k)
          umma_f16(tmem_d, da0 + k * a_step, db0 + k * b_step, idesc, (kb | k) != 0 ? 1u : 0u);
        umma_commit(&ctrl->empty[pp.s]);
        if (kb == T.num_kb - 1) umma_commit(&ctrl->accf[0]);
      }
    }
    __syncwarp();
    if (mode == HG_PRECISE) {
      if (kb % HG_DRAIN_KB == HG_DRAIN_KB - 1 || kb == T.num_kb - 1) { if (buf) ++pp.nf1; else ++pp.nf0; }
    } else if (kb == T.num_kb - 1) {
      ++pp.nf0;
    }
    pp.par ^= 1u << pp.s;
    if (++pp.s == T.nstages) pp.s = 0;
  }
  pp_io = pp;
}

// ------------------------------------------------------------------------------------------------ epilogue warps
// Staging blocks: epilogue warp e owns a swizzled 32 x 32 fp32 block (16-byte unit j of row r at unit j ^ (r & 7):
// conflict-free both ways) at stage_base + e * 4096, covering rows [32 (e & 3), +32) and columns [32 (e >> 2), +32) of
// a 128 x 64 tile. Fused problems leave the scaled accumulators there for the caller's tail.
__device__ __forceinline__ const float4* hg_stage_ptr(const uint8_t* stage_base, int r, int c) {   // c a multiple of 4
  return reinterpret_cast<const float4*>(stage_base + ((r >> 5) + ((c >> 5) << 2)) * 4096 + (r & 31) * 128 + ((((c & 31) >> 2) ^ (r & 7)) << 4));
}
// e = epilogue warp index 0 .. 7: TMEM lane quadrant q = e & 3 (== warp index & 3), column half = e >> 2.
__device__ __forceinline__ void hg_epilogue_tile(const HgProblem& P, const HgTile& T, HgCtrl* ctrl, uint8_t* stage_base, uint32_t tmem_d,
                                              HgPipe& pp_io, int e, int lane, bool stamp) {
  HgPipe pp = pp_io;
  const int q = e & 3, half = e >> 2;
  uint8_t* const stw = stage_base + e * 4096;
  const uint32_t lane_base = tmem_d + (static_cast<uint32_t>(q * 32) << 16);
  const int rsub = lane >> 3, ch = lane & 7;   // read-back mapping: 4 rows x 8 float4 per pass
  const int bn = T.bn, mode = P.mode, pM = P.M, pN = P.N, ldc = P.ldc, epi = P.epi, fuse = P.fuse;
  const int accumulate = P.accumulate | (P.acc_flag != nullptr ? __ldcg(P.acc_flag) : 0);
  float* const pC = P.C + static_cast<long long>(T.ks) * P.part_stride;
  const float* const pbias = T.ks == 0 ? P.bias : nullptr;
  const float slope = P.slope, out_scale = P.dyn_scale != nullptr ? P.out_scale * __ldcg(P.dyn_scale) : P.out_scale;
  const bool vec_ok = (ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(pC) & 15) == 0;
  const int m0 = T.m0, n0 = T.n0;
  float nsq = 0.f;         // sum of squares of the values this thread stored (HgProblem::norm_out)

  // finish one 32-column block held in v (this thread: row q * 32 + lane of the tile): the raw accumulators go through
  // the warp's staging block; scale, bias and activation are applied on the way out, four columns per lane, in a ROLLED
  // loop (the step kernel is bound by instruction fetch: this used to be 32-wide unrolled code in every epilogue
  // variant). Fused problems stop after the staging write (scaled; the tail adds the bias).
  auto finish = [&](float (&v)[32], int c0) {
    const int nbase = n0 + c0;
    if (fuse) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(stw + lane * 128 + ((j ^ (lane & 7)) << 4)) =
            make_float4(v[4 * j] * out_scale, v[4 * j + 1] * out_scale, v[4 * j + 2] * out_scale, v[4 * j + 3] * out_scale);
      return;
    }
    if (nbase >= pN || m0 + q * 32 >= pM) return;   // warp-uniform
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<float4*>(stw + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    __syncwarp();
    const int n = nbase + ch * 4;
    float b4[4] = {0.f, 0.f, 0.f, 0.f};
    if (epi != EPI_STORE && pbias != nullptr) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n + j < pN) b4[j] = __ldg(pbias + n + j);
    }
    const bool vec = vec_ok && n + 4 <= pN;
    float* dst = pC + static_cast<size_t>(m0 + q * 32 + rsub) * ldc + n;
#pragma unroll 1
    for (int it = 0; it < 8; ++it, dst += 4 * static_cast<size_t>(ldc)) {
      const int r = it * 4 + rsub;
      const float4 xr = *reinterpret_cast<const float4*>(stw + r * 128 + ((ch ^ (r & 7)) << 4));
      if (m0 + q * 32 + r < pM && n < pN) {
        float xs[4] = {xr.x * out_scale + b4[0], xr.y * out_scale + b4[1], xr.z * out_scale + b4[2], xr.w * out_scale + b4[3]};
        if (epi == EPI_BIAS_LRELU) {
#pragma unroll
          for (int j = 0; j < 4; ++j) xs[j] = leaky(xs[j], slope);
        }
        if (vec) {
          if (accumulate) {
            const float4 o = *reinterpret_cast<const float4*>(dst);
            xs[0] += o.x; xs[1] += o.y; xs[2] += o.z; xs[3] += o.w;
          }
          *reinterpret_cast<float4*>(dst) = make_float4(xs[0], xs[1], xs[2], xs[3]);
          nsq += xs[0] * xs[0] + xs[1] * xs[1] + xs[2] * xs[2] + xs[3] * xs[3];
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (n + j < pN) { const float w = accumulate ? dst[j] + xs[j] : xs[j]; dst[j] = w; nsq += w * w; }
        }
      }
    }
    __syncwarp();
  };

  if (mode == HG_PRECISE) {
    const bool mine = 32 * half < bn;   // this warp owns columns [32 half, 32 half + 32) (bn = 32: half 1 only syncs)
    float run[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) run[j] = 0.f;
    const int num_chunks = (T.num_kb + HG_DRAIN_KB - 1) / HG_DRAIN_KB;
    for (int c = 0; c < num_chunks; ++c) {
      const int buf = c & 1;
      const uint32_t nf = buf ? pp.nf1 : pp.nf0;
      mbar_wait(&ctrl->accf[buf], nf & 1);
      if (buf) ++pp.nf1; else ++pp.nf0;
      tc_fence_after();
      if (mine) {
        float v[32];
        tmem_ld_32x32(lane_base + static_cast<uint32_t>(buf * 2 * bn + 32 * half), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) run[j] += v[j];
      }
      // the last use of each buffer in this tile is released only after the small columns have been read
      if (c < num_chunks - 2) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctrl->acce[buf]);
      }
    }
    if (mine) {
      {
        float v[32];
        tmem_ld_32x32(lane_base + static_cast<uint32_t>(bn + 32 * half), v);   // small0
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) run[j] += v[j] * HG_LO_INV;
        if (num_chunks > 1) {
          tmem_ld_32x32(lane_base + static_cast<uint32_t>(3 * bn + 32 * half), v);   // small1
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) run[j] += v[j] * HG_LO_INV;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&ctrl->acce[(num_chunks - 1) & 1]);
        if (num_chunks > 1) mbar_arrive(&ctrl->acce[(num_chunks - 2) & 1]);
      }
      if (stamp && e == 0 && lane == 0) hg_stamp(pp, 5);
      finish(run, 32 * half);
      if (stamp && e == 0 && lane == 0) hg_stamp(pp, 6);
    } else {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&ctrl->acce[(num_chunks - 1) & 1]);
        if (num_chunks > 1) mbar_arrive(&ctrl->acce[(num_chunks - 2) & 1]);
      }
    }
  } else {
    mbar_wait(&ctrl->accf[0], pp.nf0 & 1);
    ++pp.nf0;
    tc_fence_after();
    // column blocks c0 = 32 (2 j + half) of this warp; the TMEM reads of all of them come first so that the
    // accumulator is released before the (slow) global stores
    for (int c0 = 32 * half; c0 < bn; c0 += 64) {
      float v[32];
      tmem_ld_32x32(lane_base + static_cast<uint32_t>(c0), v);
      tmem_ld_wait();
      if (mode == HG_MEDIUM) {
        float w[32];
        tmem_ld_32x32(lane_base + static_cast<uint32_t>(bn + c0), w);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += w[j] * HG_LO_INV;
      }
      if (c0 + 64 >= bn) {   // last TMEM read of this warp for the tile
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctrl->acce[0]);
      }
      finish(v, c0);
    }
    if (32 * half >= bn) {   // no column block of this warp in a narrow tile: still one arrival per tile
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctrl->acce[0]);
    }
  }
  if (P.norm_out != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nsq += __shfl_xor_sync(0xffffffffu, nsq, o);
    if (lane == 0) P.norm_out[T.lt * HG_NEPI + e] = nsq;
  }
  pp_io = pp;
}

// One GEMM phase for the calling CTA: every warp calls this. `tail(T, P)` is called by ALL threads of the CTA after every
// work item of a fused problem (after a __syncthreads: the item's scaled accumulators are in the staging blocks).
// `side(T, P, i)` is called by the warps without a GEMM role (i = 0 .. HG_NSIDE - 1) while the roles work on the item.
constexpr int HG_NSIDE = 6;   // warps 2, 3, 12, 13, 14, 15 of a 16-warp CTA
struct HgNoTail { __device__ __forceinline__ void operator()(const HgTile&, const HgProblem&, bool) const {} };
struct HgNoSide { __device__ __forceinline__ void operator()(const HgTile&, const HgProblem&, int) const {} };
template <class Tail, class Side>
__device__ __forceinline__ void hg_run_phase(const HgProblem* __restrict__ probs, const HgPhase& ph, int cta, int ncta,
                                             HgCtrl* ctrl, uint8_t* ring, uint8_t* stage_base, uint32_t tmem_d, HgPipe& pp,
                                             int warp, int lane, const HgTile* first_tile, Tail&& tail, Side&& side) {
  if (warp == HG_WARP_TMA) {
    if (lane == 0) hg_stamp(pp, 0);
    fence_proxy_async_global();   // operands written by generic stores of earlier phases (any CTA) -> TMA reads
  }
  for (int t = cta; t < ph.total_tiles; t += ncta) {
    const HgTile T = (t == cta && first_tile != nullptr) ? *first_tile : hg_decode(probs, ph, t);
    const HgProblem& P = probs[T.p];
    const bool stamp = t == cta;
    if (!T.null) {
      if (warp == HG_WARP_TMA) hg_produce_tile(P, T, ctrl, ring, pp, stamp);
      else if (warp == HG_WARP_MMA) hg_mma_tile(P, T, ctrl, ring, tmem_d, pp, stamp);
      else if (warp >= HG_WARP_EPI0 && warp < HG_WARP_EPI0 + HG_NEPI) hg_epilogue_tile(P, T, ctrl, stage_base, tmem_d, pp, warp - HG_WARP_EPI0, lane, stamp);
      else if (P.fuse) side(T, P, warp < HG_WARP_EPI0 ? warp - 2 : warp - (HG_WARP_EPI0 + HG_NEPI) + 2);
    }
    if (P.fuse) {
      __syncthreads();
      // more: the cluster has another work item in this phase (uniform over the cluster's CTAs: ncta is a multiple of
      // HG_CLUSTER and the fused items come first)
      tail(T, P, (t - (t % HG_CLUSTER)) + ncta < ph.total_tiles);
    }
  }
}

// ------------------------------------------------------------------------------------------------ stand-alone kernel
constexpr int HG_THREADS = 512;
__global__ void __launch_bounds__(HG_THREADS, 1) hgemm_phase_kernel(const HgProblem* __restrict__ probs, const HgPhase ph) {
  extern __shared__ uint8_t hg_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(hg_smem_raw) + 1023) & ~uintptr_t(1023));
  HgCtrl* ctrl = reinterpret_cast<HgCtrl*>(smem);
  uint8_t* ring = smem + HG_CTRL_BYTES;
  uint8_t* stage = ring + HG_RING_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem_d = hg_setup(ctrl, warp, lane);
  HgPipe pp;
  hg_run_phase(probs, ph, blockIdx.x, gridDim.x, ctrl, ring, stage, tmem_d, pp, warp, lane, nullptr, HgNoTail{}, HgNoSide{});
  hg_teardown(tmem_d, warp);
}

// =========================================================================== host side
// 2-D fp16 tensor map: `inner` contiguous elements, `outer` rows of `ld` elements; out-of-bounds box elements read as
// zero. Requirements: base 16-byte aligned, ld a multiple of 8 halves.
inline int make_tmap_f16(CUtensorMap* tm, const __half* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                         uint32_t box_outer) {
  PFN_tmapEncodeTiled fn = tmap_encode_fn();
  if (!fn) return -1;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * sizeof(__half)};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : static_cast<int>(r);
}

struct HPlanes {   // a GEMM operand as fp16 hi / lo planes (same shape and pitch)
  __half* hi = nullptr;
  __half* lo = nullptr;
};

// A is logically [M,K]: K-major => memory [M][lda]; MN-major => memory [K][lda]. Same for B with N. Returns 0 on success.
inline int hg_problem_fill(HgProblem* g, HPlanes A, int lda, int a_mn, HPlanes B, int ldb, int b_mn, float* C, int ldc, int M,
                           int N, int K, int bn, int mode, int epi, const float* bias, int ksplit, long long part_stride,
                           int accumulate, float out_scale, float slope = 0.01f) {
  *g = HgProblem{};
  auto mk = [&](CUtensorMap* tm, const __half* base, int rows, int ld, int mn, int box_rows) {
    if (!mn) return make_tmap_f16(tm, base, K, rows, ld, HG_BK, box_rows);
    return make_tmap_f16(tm, base, rows, K, ld, 64, HG_BK);
  };
  if (mode == HG_PRECISE && bn > 64) return -2;
  if (bn > 256 || (bn & 31) != 0) return -3;
  if (mode == HG_MEDIUM && 2 * bn > HG_TMEM_COLS) return -4;
  int rc;
  if ((rc = mk(&g->tmA_hi, A.hi, M, lda, a_mn, HG_BM))) return rc;
  if ((rc = mk(&g->tmB_hi, B.hi, N, ldb, b_mn, bn))) return rc;
  if (mode != HG_SINGLE) {
    if (!A.lo || !B.lo) return -5;
    if ((rc = mk(&g->tmA_lo, A.lo, M, lda, a_mn, HG_BM))) return rc;
    if ((rc = mk(&g->tmB_lo, B.lo, N, ldb, b_mn, bn))) return rc;
  }
  const int kb_total = (K + HG_BK - 1) / HG_BK;
  if (ksplit < 1) ksplit = 1;
  if (ksplit > kb_total) ksplit = kb_total;
  g->C = C; g->bias = bias; g->part_stride = part_stride;
  g->M = M; g->N = N; g->K = K; g->ldc = ldc;
  g->bn = bn; g->a_mn = a_mn; g->b_mn = b_mn; g->epi = epi; g->mode = mode; g->ksplit = ksplit;
  g->tiles_m = (M + HG_BM - 1) / HG_BM;
  g->tiles_n = (N + bn - 1) / bn;
  g->tile_base = 0;
  g->accumulate = accumulate;
  g->out_scale = out_scale;
  g->slope = slope;
  return 0;
}

// Assign work-item ranges to the problems of one phase.
inline HgPhase hg_phase_finalize(HgProblem* all, int first, int count) {
  HgPhase ph{};
  ph.first = first; ph.count = count;
  int base = 0;
  for (int i = 0; i < count; ++i) {
    all[first + i].tile_base = base;
    ph.base[i] = base;
    base += (all[first + i].fuse && !all[first + i].fuse_ks) ? HG_CLUSTER * all[first + i].tiles_n : all[first + i].tiles_m * all[first + i].tiles_n * all[first + i].ksplit;
  }
  ph.total_tiles = base;
  return ph;
}

inline cudaError_t hgemm_launch_phase(const HgProblem* dev_table, const HgPhase& ph, int max_ctas, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(hgemm_phase_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HG_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int grid = ph.total_tiles < max_ctas ? ph.total_tiles : max_ctas;
  hgemm_phase_kernel<<<grid, HG_THREADS, HG_SMEM_BYTES, st>>>(dev_table, ph);
  return cudaGetLastError();
}

}  // namespace jb
