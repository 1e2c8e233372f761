// Persistent single-pass TF32 GEMM for the inference chain (modal_predict / transform): C = act(A * B^T + bias) with
// A [M, K] and B [N, K] both K-major fp32 (TFLOAT32 tensor maps: TMA rounds to nearest on load), M in the tens of
// thousands (a chunk of cells), N and K the layer widths.
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
  const int num_kb = (P.K + GEMM_BK - 1) / GEMM_BK;
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
          tma_load_2d(sa, tmA, bar, kb * GEMM_BK, m0);                        // box {32 k, 128 rows}
          tma_load_2d(sa + GEMM_A_STAGE_BYTES, tmB, bar, kb * GEMM_BK, n0);   // box {32 k, bn rows}
        }
        __syncwarp();
        if (++s == nstages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer: accumulator buffer j & 1 for the CTA's j-th tile
    const uint32_t idesc = umma_idesc_tf32(GEMM_BM, bn, 0, 0);
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
          for (int k = 0; k < GEMM_BK / GEMM_UMMA_K; ++k)
            umma_tf32(acc, da0 + 2u * k, db0 + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
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
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(g->N), static_cast<cuuint64_t>(g->M)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(g->ldc) * sizeof(float)};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(&g->tmC, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, g->C, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : static_cast<int>(r);
}

// One problem (filled by gemm_problem_fill with K-major operands, no split, + gemm_problem_set_store_map), one CTA per
// SM at most.
inline cudaError_t gemm_launch_persistent(const GemmProblem* dev_prob, const GemmProblem& host_prob, int sms, cudaStream_t st,
                                          bool use_pdl) {
  if (host_prob.a_mn || host_prob.b_mn || host_prob.split || host_prob.accumulate || host_prob.bn > 256) return cudaErrorInvalidValue;
  if ((host_prob.ldc & 3) != 0 || (reinterpret_cast<uintptr_t>(host_prob.C) & 15) != 0) return cudaErrorInvalidValue;   // TMA store
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
