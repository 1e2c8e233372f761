// Standalone bring-up harness for the tcgen05/TMA TF32 GEMM (no torch, no python):
//   * checks every operand-major combination and edge shape against a double-precision CPU result,
//   * probes whether a TFLOAT32 tensor map rounds on load,
//   * times the headline layer shape.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/gemm_lab tools/gemm_lab.cu
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../jamie_b200/csrc/gemm_tf32.cuh"

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      return 1;                                                                        \
    }                                                                                  \
  } while (0)

static float host_tf32_rna(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u += 0x1000u;
  u &= 0xFFFFE000u;
  float y;
  memcpy(&y, &u, 4);
  return y;
}
static float host_tf32_trunc(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u &= 0xFFFFE000u;
  float y;
  memcpy(&y, &u, 4);
  return y;
}
static uint32_t rng_state = 12345u;
static float frand() {  // uniform (-1, 1)
  rng_state = rng_state * 1664525u + 1013904223u;
  return ((rng_state >> 8) * (1.0f / 8388608.0f)) - 1.0f;
}

struct Case {
  int M, N, K, a_mn, b_mn, bn;
  uint32_t lbo, sbo;
  int round_inputs;  // 1: host pre-rounds to tf32 (exact check); 0: raw fp32 inputs
  int tmap_tf32;     // tensor-map dtype TFLOAT32 instead of FLOAT32
  int epi;
  uint32_t mn_layout = 1;
  int ks = 1;
  int split = 0;
};

static int run_case(const Case& c, FILE* out) {
  const int M = c.M, N = c.N, K = c.K;
  // logical A[M,K], B[N,K]; storage depends on major. Leading dims padded to a multiple of 4 floats.
  const int lda = c.a_mn ? ((M + 3) & ~3) : ((K + 3) & ~3);
  const int ldb = c.b_mn ? ((N + 3) & ~3) : ((K + 3) & ~3);
  const int ldc = (N + 3) & ~3;
  const size_t a_rows = c.a_mn ? K : M, b_rows = c.b_mn ? K : N;
  std::vector<float> hA(a_rows * lda, 0.f), hB(b_rows * ldb, 0.f), hBias(N), hC((size_t)M * ldc, -777.f);
  std::vector<float> lA((size_t)M * K), lB((size_t)N * K);
  for (auto& v : lA) { v = frand(); if (c.round_inputs) v = host_tf32_rna(v); }
  for (auto& v : lB) { v = frand(); if (c.round_inputs) v = host_tf32_rna(v); }
  for (auto& v : hBias) v = frand();
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < K; ++k) {
      if (c.a_mn) hA[(size_t)k * lda + m] = lA[(size_t)m * K + k];
      else hA[(size_t)m * lda + k] = lA[(size_t)m * K + k];
    }
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      if (c.b_mn) hB[(size_t)k * ldb + n] = lB[(size_t)n * K + k];
      else hB[(size_t)n * ldb + k] = lB[(size_t)n * K + k];
    }
  float *dA, *dB, *dC, *dBias;
  jb::GemmProblem* dT;
  CK(cudaMalloc(&dA, hA.size() * 4)); CK(cudaMalloc(&dB, hB.size() * 4));
  CK(cudaMalloc(&dC, hC.size() * 4)); CK(cudaMalloc(&dBias, N * 4));
  CK(cudaMalloc(&dT, sizeof(jb::GemmProblem)));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dC, hC.data(), hC.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dBias, hBias.data(), N * 4, cudaMemcpyHostToDevice));
  jb::GemmProblem g;
  int rc = jb::gemm_problem_fill(&g, dA, lda, c.a_mn, dB, ldb, c.b_mn, dC, ldc, M, N, K, c.bn, c.epi, dBias, 0.01f, 0,
                                 c.tmap_tf32, c.split);
  if (rc) { fprintf(out, "tensor map encode failed rc=%d\n", rc); return 1; }
  g.mn_lbo = c.lbo; g.mn_sbo = c.sbo; g.mn_layout = c.mn_layout; g.ks = c.ks;
  int tiles = jb::gemm_table_finalize(&g, 1);
  CK(cudaMemcpy(dT, &g, sizeof(g), cudaMemcpyHostToDevice));
  cudaError_t e = jb::gemm_launch<true>(dT, 1, tiles, 0);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    fprintf(out, "M%d N%d K%d a_mn%d b_mn%d bn%d lbo%u sbo%u : LAUNCH ERROR %s\n", M, N, K, c.a_mn, c.b_mn, c.bn, c.lbo,
            c.sbo, cudaGetErrorString(e));
    return 2;  // context is dead after a trap
  }
  CK(cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost));
  double max_err = 0, max_ref = 0, max_err_tr = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double acc = 0, acc_tr = 0;
      const float* a = &lA[(size_t)m * K];
      const float* b = &lB[(size_t)n * K];
      for (int k = 0; k < K; ++k) {
        acc += (double)a[k] * (double)b[k];
        if (!c.round_inputs) acc_tr += (double)host_tf32_trunc(a[k]) * (double)host_tf32_trunc(b[k]);
      }
      if (c.epi != jb::EPI_STORE) { acc += hBias[n]; acc_tr += hBias[n]; }
      if (c.epi == jb::EPI_BIAS_LRELU) { acc = acc > 0 ? acc : 0.01 * acc; acc_tr = acc_tr > 0 ? acc_tr : 0.01 * acc_tr; }
      double got = hC[(size_t)m * ldc + n];
      max_err = fmax(max_err, fabs(got - acc));
      max_err_tr = fmax(max_err_tr, fabs(got - acc_tr));
      max_ref = fmax(max_ref, fabs(acc));
    }
  fprintf(out,
          "M%-4d N%-4d K%-4d a_mn%d b_mn%d bn%-3d ks%d sbo%-4u rnd%d tmtf32 %d epi%d : max_err %.3e (vs trunc-ref %.3e) "
          "max_ref %.3e  %s\n",
          M, N, K, c.a_mn, c.b_mn, c.bn, c.ks, c.sbo, c.round_inputs, c.tmap_tf32, c.epi, max_err, max_err_tr, max_ref,
          (max_err < 2e-5 * max_ref * ((c.round_inputs || c.split) ? 1 : 200)) ? "OK" : "MISMATCH");
  if (c.split) fprintf(out, "    split: max_err / max_ref = %.3e\n", max_err / max_ref);
  cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dBias); cudaFree(dT);
  return 0;
}

// ---- TMA rounding probe: load one 8x32 fp32 box with a given tensor-map dtype and dump shared memory.
__global__ void tma_probe_kernel(const CUtensorMap* tm, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  float* tile = reinterpret_cast<float*>(smem + 1024);
  if (threadIdx.x == 0) {
    jb::mbar_init(bar, 1);
    jb::fence_mbar_init();
    jb::mbar_arrive_expect_tx(bar, 8 * 128);
    jb::tma_load_2d(tile, tm, bar, 0, 0);
    jb::mbar_wait(bar, 0);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 256; i += blockDim.x) out[i] = tile[i];
}

static int tma_probe(FILE* out) {
  std::vector<float> h(8 * 32);
  for (auto& v : h) v = frand() * 3.f;
  float *d, *o;
  CUtensorMap* dtm;
  CK(cudaMalloc(&d, 1024)); CK(cudaMalloc(&o, 1024)); CK(cudaMalloc(&dtm, sizeof(CUtensorMap)));
  CK(cudaMemcpy(d, h.data(), 1024, cudaMemcpyHostToDevice));
  for (int dt = 0; dt < 2; ++dt) {
    CUtensorMap tm;
    if (jb::make_tmap_2d(&tm, d, 32, 8, 32, 32, 8, dt)) { fprintf(out, "probe: encode failed\n"); return 1; }
    CK(cudaMemcpy(dtm, &tm, sizeof(tm), cudaMemcpyHostToDevice));
    tma_probe_kernel<<<1, 128, 4096>>>(dtm, o);
    CK(cudaDeviceSynchronize());
    std::vector<float> g(256);
    CK(cudaMemcpy(g.data(), o, 1024, cudaMemcpyDeviceToHost));
    int exact = 0, rna = 0, trunc = 0, other = 0, swz_ok = 0;
    for (int r = 0; r < 8; ++r)
      for (int cidx = 0; cidx < 32; ++cidx) {
        // SW128: 16-byte chunk index XOR (row % 8)
        int chunk = cidx / 4, within = cidx % 4;
        float got = g[r * 32 + ((chunk ^ r) * 4 + within)];
        float src = h[r * 32 + cidx];
        if (got == src) { ++exact; ++swz_ok; }
        else if (got == host_tf32_rna(src)) { ++rna; ++swz_ok; }
        else if (got == host_tf32_trunc(src)) { ++trunc; ++swz_ok; }
        else ++other;
      }
    fprintf(out, "tma probe dtype=%s: exact %d  rna %d  trunc %d  other %d  (swizzle model ok for %d/256)\n",
            dt ? "TFLOAT32" : "FLOAT32", exact, rna, trunc, other, swz_ok);
  }
  return 0;
}

static int time_case(FILE* out, int M, int N, int K, int a_mn, int b_mn, int bn, int nprob, int dump_dbg = 0,
                     int ks = 1, int dbg_mode = 0, int split = 0) {
  const int lda = a_mn ? M : K, ldb = b_mn ? N : K;
  float *dA, *dB, *dC;
  jb::GemmProblem* dT;
  size_t asz = (size_t)M * K, bsz = (size_t)N * K, csz = (size_t)M * N;
  CK(cudaMalloc(&dA, asz * 4 * nprob)); CK(cudaMalloc(&dB, bsz * 4 * nprob)); CK(cudaMalloc(&dC, csz * 4 * nprob));
  CK(cudaMemset(dA, 0, asz * 4 * nprob)); CK(cudaMemset(dB, 0, bsz * 4 * nprob));
  CK(cudaMalloc(&dT, sizeof(jb::GemmProblem) * nprob));
  std::vector<jb::GemmProblem> g(nprob);
  for (int i = 0; i < nprob; ++i)
    if (jb::gemm_problem_fill(&g[i], dA + asz * i, lda, a_mn, dB + bsz * i, ldb, b_mn, dC + csz * i, N, M, N, K, bn, 0,
                              nullptr, 0.f, 0, 1, split)) return 1;
  int tiles = jb::gemm_table_finalize(g.data(), nprob);
  long long* dDbg = nullptr;
  if (dump_dbg) {
    CK(cudaMalloc(&dDbg, sizeof(long long) * 8 * tiles));
    CK(cudaMemset(dDbg, 0, sizeof(long long) * 8 * tiles));
    for (int i = 0; i < nprob; ++i) g[i].dbg = dDbg;
  }
  for (int i = 0; i < nprob; ++i) { g[i].ks = ks; g[i].dbg_mode = dbg_mode; }
  CK(cudaMemcpy(dT, g.data(), sizeof(jb::GemmProblem) * nprob, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 5; ++i) jb::gemm_launch<true>(dT, nprob, tiles, 0);
  CK(cudaDeviceSynchronize());
  const int iters = 50;
  cudaEventRecord(e0);
  for (int i = 0; i < iters; ++i) jb::gemm_launch<true>(dT, nprob, tiles, 0);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double us = ms * 1000.0 / iters;
  double tf = 2.0 * M * N * K * nprob / (us * 1e-6) / 1e12;
  fprintf(out, "time: %d x [M%d N%d K%d] a_mn%d b_mn%d bn%d ks%d mode%d split%d tiles %d : %.2f us/launch  %.1f TFLOP/s (tf32)\n",
          nprob, M, N, K, a_mn, b_mn, bn, ks, dbg_mode, split, tiles, us, tf);
  if (dump_dbg) {
    std::vector<long long> h(8 * tiles);
    CK(cudaMemcpy(h.data(), dDbg, sizeof(long long) * 8 * tiles, cudaMemcpyDeviceToHost));
    for (int t = 0; t < tiles; t += (tiles > 8 ? tiles / 8 : 1)) {
      long long* d = &h[8 * t];
      fprintf(out, "  cta %3d: setup %lld  first_full %lld  mma_issued %lld  tmem_full %lld  epi_done %lld (clk)\n",
              t, d[1] - d[0], d[2] - d[0], d[3] - d[0], d[4] - d[0], d[5] - d[0]);
    }
    cudaFree(dDbg);
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dT);
  return 0;
}

int main(int argc, char** argv) {
  FILE* out = stdout;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  fprintf(out, "device: %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  int which = argc > 1 ? atoi(argv[1]) : 0;
  if (which == 0) {
    if (tma_probe(out)) return 1;
    // K-major / K-major first: the forward-pass layout.
    const Case base[] = {
        {128, 128, 32, 0, 0, 128, 4096, 1024, 1, 0, 0},
        {128, 128, 128, 0, 0, 128, 4096, 1024, 1, 0, 0},
        {512, 1024, 512, 0, 0, 128, 4096, 1024, 1, 0, 0},
        {512, 512, 1024, 0, 0, 64, 4096, 1024, 1, 0, 1},
        {512, 32, 512, 0, 0, 32, 4096, 1024, 1, 0, 1},
        {512, 512, 32, 0, 0, 128, 4096, 1024, 1, 0, 2},
        {300, 1000, 2000, 0, 0, 128, 4096, 1024, 1, 0, 0},
        {512, 39, 78, 0, 0, 64, 4096, 1024, 1, 0, 1},
        {512, 78, 39, 0, 0, 128, 4096, 1024, 1, 0, 0},
        {512, 1024, 512, 0, 0, 256, 4096, 1024, 1, 0, 0},
        // raw fp32 inputs: how does the tensor core / TMA treat the low mantissa bits?
        {512, 512, 512, 0, 0, 128, 4096, 1024, 0, 0, 0},
        {512, 512, 512, 0, 0, 128, 4096, 1024, 0, 1, 0},
    };
    for (int ks = 1; ks <= 4; ks *= 2)
      for (Case c : base) {
        c.ks = ks;
        if (run_case(c, out) == 2) return 2;
      }
  } else if (which == 1) {
    // dgrad layout: A K-major, B MN-major.
    const Case cs[] = {
        {128, 128, 32, 0, 1, 128, 4096, 512, 1, 1, 0},
        {512, 512, 1024, 0, 1, 128, 4096, 512, 1, 1, 0},
        {300, 2000, 1000, 0, 1, 128, 4096, 512, 1, 1, 0},
        {512, 39, 78, 0, 1, 64, 4096, 512, 1, 1, 0},
        {512, 32, 512, 0, 1, 32, 4096, 512, 1, 1, 0},
    };
    for (int ks = 1; ks <= 2; ks *= 2)
      for (Case c : cs) {
        c.ks = ks;
        if (run_case(c, out) == 2) return 2;
      }
  } else if (which == 2) {
    // wgrad layout: both MN-major.
    const Case cs[] = {
        {128, 128, 32, 1, 1, 128, 4096, 512, 1, 1, 0},
        {1024, 512, 512, 1, 1, 128, 4096, 512, 1, 1, 0},
        {2000, 1000, 300, 1, 1, 128, 4096, 512, 1, 1, 0},
        {78, 39, 512, 1, 1, 64, 4096, 512, 1, 1, 0},
        {512, 32, 512, 1, 1, 32, 4096, 512, 1, 1, 0},
    };
    for (const Case& c : cs)
      if (run_case(c, out) == 2) return 2;
  } else if (which == 3) {
    // alternative MN descriptor conventions in case (1)/(2) mismatch
    const Case cs[] = {
        {128, 128, 32, 0, 1, 128, 512, 4096, 1, 1, 0, 1},
        {128, 128, 32, 0, 1, 128, 4096, 1024, 1, 1, 0, 1},
        {128, 128, 32, 0, 1, 128, 1024, 512, 1, 1, 0, 1},
        {128, 32, 32, 0, 1, 32, 4096, 512, 1, 1, 0, 1},
        {128, 32, 8, 0, 1, 32, 4096, 512, 1, 1, 0, 1},
        {128, 32, 8, 1, 0, 32, 4096, 512, 1, 1, 0, 1},
    };
    for (const Case& c : cs)
      if (run_case(c, out) == 2) return 2;
  } else if (which == 7) {
    // error-compensated 3xTF32 on raw fp32 inputs, all operand layouts, ks 1 and 2
    const Case cs[] = {
        {128, 128, 32, 0, 0, 128, 4096, 512, 0, 0, 0}, {512, 1024, 512, 0, 0, 64, 4096, 512, 0, 0, 1},
        {512, 512, 1024, 0, 1, 64, 4096, 512, 0, 0, 0}, {1024, 512, 512, 1, 1, 64, 4096, 512, 0, 0, 0},
        {300, 1000, 2000, 0, 0, 64, 4096, 512, 0, 0, 1}, {78, 39, 512, 1, 1, 64, 4096, 512, 0, 0, 0},
        {512, 39, 78, 0, 1, 64, 4096, 512, 0, 0, 0}, {512, 32, 512, 0, 0, 32, 4096, 512, 0, 0, 1},
        {2000, 1000, 300, 1, 1, 64, 4096, 512, 0, 0, 0}, {512, 64, 512, 0, 0, 64, 4096, 512, 0, 0, 1},
    };
    for (int ks = 1; ks <= 2; ++ks)
      for (Case c : cs) {
        c.ks = ks; c.split = 1;
        if (run_case(c, out) == 2) return 2;
      }
    for (int sp = 0; sp < 2; ++sp) {
      time_case(out, 512, 1024, 512, 0, 0, 64, 2, 1, 2, 0, sp);
      time_case(out, 512, 512, 1024, 0, 1, 64, 2, 1, 2, 0, sp);
      time_case(out, 1024, 512, 512, 1, 1, 64, 2, 1, 2, 0, sp);
      time_case(out, 512, 1024, 512, 0, 0, 128, 2, 1, sp ? 1 : 2, 0, sp);
      time_case(out, 512, 1024, 512, 0, 0, 32, 2, 1, 2, 0, sp);
    }
  } else if (which == 5) {
    for (int mode = 0; mode < 1; ++mode)
      for (int ks = 1; ks <= 4; ks *= 2) {
        time_case(out, 512, 1024, 512, 0, 0, 128, 2, 1, ks, mode);
        time_case(out, 512, 1024, 512, 0, 0, 64, 2, 1, ks, mode);
      }
    for (int mode = 3; mode <= 6; mode += 3) {
      time_case(out, 512, 1024, 512, 0, 0, 128, 2, 1, 1, mode);
      time_case(out, 512, 1024, 512, 0, 0, 64, 2, 1, 1, mode);
      time_case(out, 512, 1024, 512, 0, 0, 128, 2, 1, 4, mode);
    }
    time_case(out, 1024, 512, 512, 1, 1, 128, 2, 1, 1, 0);
    time_case(out, 1024, 512, 512, 1, 1, 128, 2, 1, 2, 0);
    time_case(out, 512, 512, 1024, 0, 1, 128, 2, 1, 2, 0);
    time_case(out, 65536, 1024, 512, 0, 0, 256, 1, 1, 1, 0);
    time_case(out, 65536, 1024, 512, 0, 0, 256, 1, 1, 2, 0);
    time_case(out, 65536, 1024, 512, 0, 0, 128, 1, 1, 2, 0);
  } else if (which == 4) {
    time_case(out, 512, 1024, 512, 0, 0, 128, 2);
    time_case(out, 512, 1024, 512, 0, 0, 64, 2);
    time_case(out, 512, 1024, 512, 0, 0, 256, 2);
    time_case(out, 512, 512, 1024, 0, 0, 64, 2);
    time_case(out, 512, 512, 1024, 0, 1, 64, 2);
    time_case(out, 1024, 512, 512, 1, 1, 64, 2);
    time_case(out, 1024, 512, 512, 1, 1, 128, 2);
    time_case(out, 8192, 1024, 512, 0, 0, 128, 1);
    time_case(out, 65536, 1024, 512, 0, 0, 128, 1);
    time_case(out, 65536, 1024, 512, 0, 0, 256, 1);
  }
  return 0;
}
