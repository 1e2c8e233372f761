// Persistent single-pass GEMM for the inference chain (modal_predict / transform): C = act(A * B^T + bias) with
// A [M, K] and B [N, K] both K-major, M in the tens of thousands (a chunk of cells), N and K the layer widths.
// Operands: fp32 read as TF32 (TFLOAT32 tensor maps: TMA rounds to nearest on load) or fp16 (kind::f16: the same 128-byte
// k-block rows hold 64 elements instead of 32, so a layer moves half the operand bytes through L2 -> shared memory, the
// measured bound of the TF32 chain). Output: fp32, or fp16 when the next layer takes fp16 operands: the wide layers of
// the chain exchange fp16 activations (11-bit significand, like the TF32 rounding they replace), the first layer reads
// the caller's fp32 rows as TF32 and the last one writes fp32.
//
// One CTA per SM walks the output tiles (tile = blockIdx.x, += gridDim.x; the N tiles of one M tile are adjacent, so
// neighbouring SMs share the A rows in L2). The operand ring, the barriers and the TMEM allocation live for the whole
// kernel, and the accumulator is double-buffered in TMEM (2 x bn columns): while the four epilogue warps drain tile i
// (tcgen05.ld -> bias / LeakyReLU -> swizzled 32 x 32 block in shared memory -> one TMA store, which also clips the
// ragged edges) the TMA and MMA warps already run the main loop of tile i + 1. Compared with the one-tile-per-CTA kernel
// of the training step this removes the per-tile prologue (barrier init, TMEM allocation, pipeline fill) and the exposed
// epilogue from every tile but the last.
#pragma once
#include "gemm_tf32.cuh"
#include "hgemm.cuh"   // umma_f16, umma_idesc_f16, make_tmap_f16

namespace jb {

struct GemmPersistCtrl {
  uint64_t full[GEMM_MAX_STAGES];
  uint64_t empty[GEMM_MAX_STAGES];
  uint64_t acc_full[2];    // the MMAs of a tile have completed in accumulator buffer b
  uint64_t acc_empty[2];   // the epilogue warps have read buffer b out of TMEM (4 arrivals)
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tf32_persistent_kernel(const GemmProblem* __restrict__ prob) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  GemmPersistCtrl* ctrl = reinterpret_cast<GemmPersistCtrl*>(smem);
  uint8_t* tiles = smem + GEMM_CTRL_SMEM;
  float* epi_stage = reinterpret_cast<float*>(tiles + GEMM_TILE_SMEM);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const GemmProblem& P = *prob;
  // read once into registers: the inline-asm barriers carry "memory" clobbers
  const int tiles_n = P.tiles_n;
  const int total_tiles = P.tiles_m * tiles_n;
  const int bn = P.bn;
  const int pM = P.M, pN = P.N, epi = P.epi;
  const float* const pbias = P.bias;
  const float slope = P.slope;
  const int f16 = P.f16_ops, out_f16 = P.out_f16;
  const int bk = f16 ? 2 * GEMM_BK : GEMM_BK;   // elements per k-block (128 bytes per row either way)
  const int num_kb = (P.K + bk - 1) / bk;
  const int b_bytes = bn * GEMM_BK * 4;
  const int kb_bytes = GEMM_A_STAGE_BYTES + b_bytes;
  int nstages = GEMM_TILE_SMEM / kb_bytes;
  if (nstages > GEMM_MAX_STAGES) nstages = GEMM_MAX_STAGES;
  const uint32_t tmem_cols = static_cast<uint32_t>(2 * bn);   // bn in {32, 64, 128, 256}: a power of two >= 64

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.tmA);
    tma_prefetch_desc(&P.tmB);
    for (int s = 0; s < nstages; ++s) {
      mbar_init(&ctrl->full[s], 1);
      mbar_init(&ctrl->empty[s], 1);
    }
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
  grid_dep_wait();      // everything above touched only kernel parameters: it overlaps the previous kernel's tail
  grid_dep_launch();

  if (warp == 0) {
    // ------------------------------------------------ TMA producer: one continuous stream of k-blocks over all tiles
    const CUtensorMap* const tmA = &P.tmA;
    const CUtensorMap* const tmB = &P.tmB;
    int s = 0;
    uint32_t ph = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int m0 = (t / tiles_n) * GEMM_BM, n0 = (t % tiles_n) * bn;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&ctrl->empty[s], ph ^ 1);
        if (elect_one()) {
          uint64_t* bar = &ctrl->full[s];
          mbar_arrive_expect_tx(bar, static_cast<uint32_t>(kb_bytes));
          uint8_t* sa = tiles + s * kb_bytes;
          tma_load_2d(sa, tmA, bar, kb * bk, m0);                        // box {128 bytes of k, 128 rows}
          tma_load_2d(sa + GEMM_A_STAGE_BYTES, tmB, bar, kb * bk, n0);   // box {128 bytes of k, bn rows}
        }
        __syncwarp();
        if (++s == nstages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer: accumulator buffer j & 1 for the CTA's j-th tile
    const uint32_t idesc = f16 ? umma_idesc_f16(GEMM_BM, bn, 0, 0) : umma_idesc_tf32(GEMM_BM, bn, 0, 0);
    const uint64_t d_hi = umma_smem_desc(0u, 16u, 1024u, 2u);   // K-major, SWIZZLE_128B
    const uint32_t tiles_u32 = smem_u32(tiles);
    int s = 0;
    uint32_t ph = 0;
    int j = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++j) {
      const int buf = j & 1;
      mbar_wait(&ctrl->acc_empty[buf], ((j >> 1) & 1) ^ 1);   // first use of a buffer passes at once
      tc_fence_after();
      const uint32_t acc = tmem_d + static_cast<uint32_t>(buf * bn);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&ctrl->full[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = tiles_u32 + s * kb_bytes;
          const uint64_t da0 = d_hi | static_cast<uint64_t>((sa >> 4) & 0x3FFFu);
          const uint64_t db0 = d_hi | static_cast<uint64_t>(((sa + GEMM_A_STAGE_BYTES) >> 4) & 0x3FFFu);
#pragma unroll
          for (int k = 0; k < GEMM_BK / GEMM_UMMA_K; ++k) {   // four instructions of 32 bytes of K each
            if (f16) umma_f16(acc, da0 + 2u * k, db0 + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
            else umma_tf32(acc, da0 + 2u * k, db0 + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&ctrl->empty[s]);                              // frees the ring slot when these MMAs have read it
          if (kb == num_kb - 1) umma_commit(&ctrl->acc_full[buf]);   // ... and hands the accumulator to the epilogue
        }
        __syncwarp();
        if (++s == nstages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp < 6) {
    // ------------------------------------------------ epilogue warps: TMEM -> registers -> swizzled smem block -> TMA store
    const int q = warp & 3;   // TMEM lane quadrant this warp may access
    // 32 rows x 128 B per warp in the SWIZZLE_128B pattern the output tensor map expects: the 16-byte unit j of row r
    // lives at unit j ^ (r & 7), so the row-per-lane 128-bit writes are bank-conflict free
    uint8_t* const stw = reinterpret_cast<uint8_t*>(epi_stage) + q * 4096;
    const CUtensorMap* const tmC = &P.tmC;
    if (lane == 0) tma_prefetch_desc(tmC);
    int j = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++j) {
      const int buf = j & 1;
      const int m0 = (t / tiles_n) * GEMM_BM, n0 = (t % tiles_n) * bn;
      mbar_wait(&ctrl->acc_full[buf], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t lane_base = tmem_d + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(buf * bn);
      if (out_f16) {
        // fp16 output: 64 columns per store (32 rows x 128 bytes in the same swizzle)
        for (int c0 = 0; c0 < bn; c0 += 64) {
          const int nbase = n0 + c0;
          float v[64];
          tmem_ld_32x32(lane_base + static_cast<uint32_t>(c0), *reinterpret_cast<float(*)[32]>(&v[0]));
          tmem_ld_32x32(lane_base + static_cast<uint32_t>(c0 + 32), *reinterpret_cast<float(*)[32]>(&v[32]));
          tmem_ld_wait();
          if (c0 + 64 >= bn) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&ctrl->acc_empty[buf]);
          }
          if (nbase >= pN || m0 + q * 32 >= pM) continue;
          if (epi != EPI_STORE) {
            const float bl0 = (nbase + lane < pN) ? __ldg(pbias + nbase + lane) : 0.f;
            const float bl1 = (nbase + 32 + lane < pN) ? __ldg(pbias + nbase + 32 + lane) : 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float x0 = v[i] + __shfl_sync(0xffffffffu, bl0, i), x1 = v[32 + i] + __shfl_sync(0xffffffffu, bl1, i);
              if (epi == EPI_BIAS_LRELU) { x0 = leaky(x0, slope); x1 = leaky(x1, slope); }
              v[i] = x0; v[32 + i] = x1;
            }
          }
          if (lane == 0) tma_store_wait_read();
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {   // 16-byte unit i = columns 8 i .. 8 i + 7
            const __half2 h0 = __floats2half2_rn(v[8 * i], v[8 * i + 1]), h1 = __floats2half2_rn(v[8 * i + 2], v[8 * i + 3]);
            const __half2 h2 = __floats2half2_rn(v[8 * i + 4], v[8 * i + 5]), h3 = __floats2half2_rn(v[8 * i + 6], v[8 * i + 7]);
            *reinterpret_cast<uint4*>(stw + lane * 128 + ((i ^ (lane & 7)) << 4)) =
                make_uint4(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1),
                           *reinterpret_cast<const uint32_t*>(&h2), *reinterpret_cast<const uint32_t*>(&h3));
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(tmC, stw, nbase, m0 + q * 32);
            tma_store_commit();
          }
        }
        continue;
      }
      for (int c0 = 0; c0 < bn; c0 += 32) {
        const int nbase = n0 + c0;
        float v[32];
        tmem_ld_32x32(lane_base + static_cast<uint32_t>(c0), v);
        tmem_ld_wait();
        if (c0 + 32 >= bn) {   // last read of this buffer: the next tile's MMAs may overwrite it
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ctrl->acc_empty[buf]);
        }
        if (nbase >= pN || m0 + q * 32 >= pM) continue;   // warp-uniform: nothing of this block lies inside C
        if (epi != EPI_STORE) {
          const float bl = (nbase + lane < pN) ? __ldg(pbias + nbase + lane) : 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float x = v[i] + __shfl_sync(0xffffffffu, bl, i);
            if (epi == EPI_BIAS_LRELU) x = leaky(x, slope);
            v[i] = x;
          }
        }
        if (lane == 0) tma_store_wait_read();   // the previous store of this warp has read the staging block
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i)
          *reinterpret_cast<float4*>(stw + lane * 128 + ((i ^ (lane & 7)) << 4)) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        fence_proxy_async_smem();   // generic-proxy writes -> visible to the TMA (async proxy)
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(tmC, stw, nbase, m0 + q * 32);   // rows / columns beyond C are clipped by the tensor map
          tma_store_commit();
        }
      }
    }
    if (lane == 0) tma_store_wait_all();   // global writes complete before the kernel ends
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_d, tmem_cols);
}

// Output tensor map of a filled problem (C [M, N] fp32, pitch ldc): 32 x 32 boxes in the 128-byte swizzle.
inline int gemm_problem_set_store_map(GemmProblem* g) {
  PFN_tmapEncodeTiled fn = tmap_encode_fn();
  if (!fn) return -1;
  if (g->out_f16) {   // C [M, N] fp16, pitch ldc halves: 64 x 32 boxes in the 128-byte swizzle
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(g->N), static_cast<cuuint64_t>(g->M)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(g->ldc) * sizeof(__half)};
    cuuint32_t box[2] = {64, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(&g->tmC, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, g->C, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : static_cast<int>(r);
  }
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(g->N), static_cast<cuuint64_t>(g->M)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(g->ldc) * sizeof(float)};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(&g->tmC, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, g->C, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : static_cast<int>(r);
}

// A problem of the inference chain: K-major operands, fp32-as-TF32 or fp16 (A, B: [rows, K] with pitches in elements of
// their type); C fp32 or fp16 (ldc in elements of its type).
inline int gemm_chain_problem_fill(GemmProblem* g, const void* A, int lda, const void* B, int ldb, void* C, int ldc, int M, int N, int K,
                                   int bn, int epi, const float* bias, float slope, int f16_ops, int out_f16) {
  int rc;
  if (!f16_ops) {
    if ((rc = gemm_problem_fill(g, static_cast<const float*>(A), lda, 0, static_cast<const float*>(B), ldb, 0, static_cast<float*>(C), ldc, M, N, K, bn,
                                epi, bias, slope, 0)))
      return rc;
  } else {
    *g = GemmProblem{};
    if ((rc = make_tmap_f16(&g->tmA, static_cast<const __half*>(A), K, M, lda, 64, GEMM_BM))) return rc;
    if ((rc = make_tmap_f16(&g->tmB, static_cast<const __half*>(B), K, N, ldb, 64, bn))) return rc;
    g->C = static_cast<float*>(C); g->bias = bias;
    g->M = M; g->N = N; g->K = K; g->ldc = ldc;
    g->bn = bn; g->epi = epi;
    g->tiles_m = (M + GEMM_BM - 1) / GEMM_BM;
    g->tiles_n = (N + bn - 1) / bn;
    g->slope = slope;
  }
  g->f16_ops = f16_ops; g->out_f16 = out_f16;
  return gemm_problem_set_store_map(g);
}

// One problem (filled by gemm_problem_fill with K-major operands, no split, + gemm_problem_set_store_map), one CTA per
// SM at most.
inline cudaError_t gemm_launch_persistent(const GemmProblem* dev_prob, const GemmProblem& host_prob, int sms, cudaStream_t st,
                                          bool use_pdl) {
  if (host_prob.a_mn || host_prob.b_mn || host_prob.split || host_prob.accumulate || host_prob.bn > 256) return cudaErrorInvalidValue;
  if ((host_prob.ldc & (host_prob.out_f16 ? 7 : 3)) != 0 || (reinterpret_cast<uintptr_t>(host_prob.C) & 15) != 0) return cudaErrorInvalidValue;   // TMA store
  if (host_prob.out_f16 && (host_prob.bn & 63) != 0) return cudaErrorInvalidValue;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tf32_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int tiles = host_prob.tiles_m * host_prob.tiles_n;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(tiles < sms ? tiles : sms);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = GEMM_SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = use_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, gemm_tf32_persistent_kernel, dev_prob);
}

}  // namespace jb
