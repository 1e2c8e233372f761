// Standalone bring-up harness for the tcgen05/TMA TF32 GEMM (no torch, no python):
//   * checks every operand-major combination, edge shapes, and both numeric modes (single-pass TF32 on TFLOAT32 maps,
//     3xTF32 on pre-split hi/lo planes) against a double-precision CPU result,
//   * times the headline layer shapes of the training step (warm L2, back-to-back launches).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/gemm_lab tools/gemm_lab.cu -lcuda
// Run:   tools/gemm_lab [check|time|all]
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
static uint32_t rng_state = 12345u;
static float frand() {  // uniform (-1, 1)
  rng_state = rng_state * 1664525u + 1013904223u;
  return ((rng_state >> 8) * (1.0f / 8388608.0f)) - 1.0f;
}

struct Case {
  int M, N, K, a_mn, b_mn, bn, split, epi, accumulate;
  int ck = 1;           // split-K cluster size
};

static int run_case(const Case& c) {
  const int M = c.M, N = c.N, K = c.K;
  const int lda = c.a_mn ? ((M + 3) & ~3) : ((K + 3) & ~3);
  const int ldb = c.b_mn ? ((N + 3) & ~3) : ((K + 3) & ~3);
  const int ldc = (N + 3) & ~3;
  const size_t na = static_cast<size_t>(c.a_mn ? K : M) * lda, nb = static_cast<size_t>(c.b_mn ? K : N) * ldb;
  std::vector<float> A(na, 0.f), B(nb, 0.f), bias(N), C0(static_cast<size_t>(M) * ldc);
  auto ai = [&](int m, int k) { return c.a_mn ? static_cast<size_t>(k) * lda + m : static_cast<size_t>(m) * lda + k; };
  auto bi = [&](int n, int k) { return c.b_mn ? static_cast<size_t>(k) * ldb + n : static_cast<size_t>(n) * ldb + k; };
  for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) A[ai(m, k)] = frand();
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) B[bi(n, k)] = frand();
  for (int n = 0; n < N; ++n) bias[n] = frand();
  for (auto& v : C0) v = frand();
  std::vector<float> Ah(na), Al(na), Bh(nb), Bl(nb);
  for (size_t i = 0; i < na; ++i) { Ah[i] = host_tf32_rna(A[i]); Al[i] = host_tf32_rna(A[i] - Ah[i]); }
  for (size_t i = 0; i < nb; ++i) { Bh[i] = host_tf32_rna(B[i]); Bl[i] = host_tf32_rna(B[i] - Bh[i]); }
  float *dA, *dAl, *dB, *dBl, *dC, *dbias;
  CK(cudaMalloc(&dA, na * 4)); CK(cudaMalloc(&dAl, na * 4)); CK(cudaMalloc(&dB, nb * 4)); CK(cudaMalloc(&dBl, nb * 4));
  CK(cudaMalloc(&dC, C0.size() * 4)); CK(cudaMalloc(&dbias, N * 4));
  CK(cudaMemcpy(dA, c.split ? Ah.data() : A.data(), na * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dAl, Al.data(), na * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, c.split ? Bh.data() : B.data(), nb * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dBl, Bl.data(), nb * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dC, C0.data(), C0.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dbias, bias.data(), N * 4, cudaMemcpyHostToDevice));
  jb::GemmProblem g;
  int rc = jb::gemm_problem_fill(&g, dA, lda, c.a_mn, dB, ldb, c.b_mn, dC, ldc, M, N, K, c.bn, c.epi, dbias, 0.01f, c.accumulate,
                                 1, c.split ? dAl : nullptr, c.split ? dBl : nullptr);
  if (rc) { printf("tensor map encode failed %d\n", rc); return 1; }
  const int tiles = jb::gemm_table_finalize(&g, 1, c.ck);
  jb::GemmProblem* dg;
  CK(cudaMalloc(&dg, sizeof g));
  CK(cudaMemcpy(dg, &g, sizeof g, cudaMemcpyHostToDevice));
  CK(jb::gemm_launch(dg, 1, tiles, 0, false, c.ck));
  CK(cudaDeviceSynchronize());
  std::vector<float> C(C0.size());
  CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
  double err2 = 0, ref2 = 0, emax = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += static_cast<double>(A[ai(m, k)]) * B[bi(n, k)];
      if (c.epi != jb::EPI_STORE) s += bias[n];
      if (c.epi == jb::EPI_BIAS_LRELU) s = s > 0 ? s : 0.01 * s;
      if (c.accumulate) s += C0[static_cast<size_t>(m) * ldc + n];
      const double d = C[static_cast<size_t>(m) * ldc + n] - s;
      err2 += d * d; ref2 += s * s; emax = fmax(emax, fabs(d));
    }
  const double rel = sqrt(err2 / (ref2 + 1e-30));
  const double tol = c.split ? 5e-7 : 6e-4;
  printf("%s M%-4d N%-4d K%-4d a_mn%d b_mn%d bn%-3d split%d epi%d acc%d ck%d : rel %.3e max %.3e %s\n", rel < tol ? "ok  " : "FAIL",
         M, N, K, c.a_mn, c.b_mn, c.bn, c.split, c.epi, c.accumulate, c.ck, rel, emax, rel < tol ? "" : "<<<<<<");
  cudaFree(dA); cudaFree(dAl); cudaFree(dB); cudaFree(dBl); cudaFree(dC); cudaFree(dbias); cudaFree(dg);
  return rel < tol ? 0 : 2;
}

// nprob identical problems in one launch (the two modalities of a stage), L2-warm, back-to-back
static int time_case(int nprob, int M, int N, int K, int a_mn, int b_mn, int bn, int split, bool pdl, int ck = 1) {
  const int lda = a_mn ? M : K, ldb = b_mn ? N : K;
  std::vector<jb::GemmProblem> g(nprob);
  std::vector<float*> bufs;
  for (int i = 0; i < nprob; ++i) {
    float *dA, *dAl, *dB, *dBl, *dC;
    CK(cudaMalloc(&dA, static_cast<size_t>(M) * K * 4)); CK(cudaMalloc(&dAl, static_cast<size_t>(M) * K * 4));
    CK(cudaMalloc(&dB, static_cast<size_t>(N) * K * 4)); CK(cudaMalloc(&dBl, static_cast<size_t>(N) * K * 4));
    CK(cudaMalloc(&dC, static_cast<size_t>(M) * N * 4));
    CK(cudaMemset(dA, 0, static_cast<size_t>(M) * K * 4)); CK(cudaMemset(dAl, 0, static_cast<size_t>(M) * K * 4));
    CK(cudaMemset(dB, 0, static_cast<size_t>(N) * K * 4)); CK(cudaMemset(dBl, 0, static_cast<size_t>(N) * K * 4));
    bufs.insert(bufs.end(), {dA, dAl, dB, dBl, dC});
    if (jb::gemm_problem_fill(&g[i], dA, lda, a_mn, dB, ldb, b_mn, dC, N, M, N, K, bn, jb::EPI_STORE, nullptr, 0.f, 0, 1,
                              split ? dAl : nullptr, split ? dBl : nullptr)) { printf("encode failed\n"); return 1; }
  }
  const int tiles = jb::gemm_table_finalize(g.data(), nprob, ck);
  jb::GemmProblem* dg;
  CK(cudaMalloc(&dg, nprob * sizeof(jb::GemmProblem)));
  CK(cudaMemcpy(dg, g.data(), nprob * sizeof(jb::GemmProblem), cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 5; ++i) CK(jb::gemm_launch(dg, nprob, tiles, 0, pdl, ck, g.data()));
  const int iters = 200;
  CK(cudaEventRecord(e0));
  for (int i = 0; i < iters; ++i) CK(jb::gemm_launch(dg, nprob, tiles, 0, pdl, ck, g.data()));
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  const double us = ms * 1e3 / iters;
  const double tf = 2.0 * nprob * M * N * K / (us * 1e-6) / 1e12;
  const double mb = static_cast<double>(tiles / ck) * ((K + 31) / 32) * (128 + bn) * 128 * (split ? 2 : 1) / 1e6;
  printf("time: %d x [M%d N%d K%d] a_mn%d b_mn%d bn%-3d split%d pdl%d ck%d ctas %-3d : %6.2f us/launch  %6.1f TFLOP/s (algorithmic)  L2->SM %.1f MB = %.2f TB/s\n",
         nprob, M, N, K, a_mn, b_mn, bn, split, pdl ? 1 : 0, ck, tiles, us, tf, mb, mb / us);
  for (float* p : bufs) cudaFree(p);
  cudaFree(dg);
  return 0;
}

int main(int argc, char** argv) {
  const char* what = argc > 1 ? argv[1] : "all";
  CK(cudaFuncSetAttribute(jb::gemm_tf32_grouped_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, jb::GEMM_SMEM_BYTES));
  int bad = 0;
  if (!strcmp(what, "check") || !strcmp(what, "all")) {
    const Case cases[] = {
        // single pass, every major combination, ragged edges
        {256, 128, 96, 0, 0, 64, 0, jb::EPI_BIAS, 0},   {200, 100, 72, 0, 1, 64, 0, jb::EPI_STORE, 0},
        {130, 39, 300, 1, 1, 64, 0, jb::EPI_STORE, 1},  {128, 32, 32, 1, 0, 32, 0, jb::EPI_BIAS_LRELU, 0},
        {300, 2000, 40, 0, 0, 128, 0, jb::EPI_BIAS_LRELU, 0},
        // 3xTF32 on planes
        {512, 1024, 512, 0, 0, 64, 1, jb::EPI_BIAS, 0}, {512, 512, 1024, 0, 0, 64, 1, jb::EPI_BIAS, 0},
        {512, 64, 512, 0, 0, 64, 1, jb::EPI_BIAS, 0},   {512, 512, 32, 0, 0, 64, 1, jb::EPI_BIAS, 0},
        {512, 1024, 512, 0, 1, 64, 1, jb::EPI_STORE, 0}, {512, 32, 512, 0, 1, 32, 1, jb::EPI_STORE, 0},
        {300, 1000, 2000, 0, 0, 64, 1, jb::EPI_BIAS, 0}, {300, 78, 39, 0, 1, 64, 1, jb::EPI_STORE, 0},
        {700, 64, 1302, 0, 0, 64, 1, jb::EPI_STORE, 0}, {700, 1302, 64, 0, 1, 64, 1, jb::EPI_BIAS, 0},
        {130, 39, 300, 1, 1, 64, 1, jb::EPI_STORE, 1},  {100, 130, 70, 1, 0, 64, 1, jb::EPI_STORE, 0},
        // split-K clusters with the DSMEM reduce-scatter: ck = 2, 4; both numeric modes; ragged edges; uneven k-block shares
        {512, 512, 1024, 0, 0, 64, 1, jb::EPI_BIAS, 0, 2},   {512, 64, 512, 0, 0, 64, 1, jb::EPI_BIAS, 0, 4},
        {512, 32, 512, 0, 1, 32, 1, jb::EPI_STORE, 0, 4},    {512, 512, 1024, 0, 1, 64, 1, jb::EPI_STORE, 0, 2},
        {300, 1000, 2000, 0, 0, 64, 1, jb::EPI_BIAS, 0, 2},  {300, 78, 200, 0, 1, 64, 1, jb::EPI_STORE, 0, 4},
        {700, 64, 1302, 0, 0, 64, 1, jb::EPI_BIAS_LRELU, 0, 4}, {130, 39, 300, 1, 1, 64, 1, jb::EPI_STORE, 1, 2},
        {64, 512, 512, 1, 1, 64, 0, jb::EPI_STORE, 0, 4},    {1024, 512, 512, 1, 1, 128, 0, jb::EPI_STORE, 1, 2},
        {256, 128, 160, 0, 0, 64, 0, jb::EPI_BIAS, 0, 4},    {100, 130, 70, 1, 0, 32, 1, jb::EPI_STORE, 0, 2},
        // 256-wide single-pass tiles (wgrad launch in one wave)
        {1024, 512, 512, 1, 1, 256, 0, jb::EPI_STORE, 1},    {512, 1024, 512, 1, 1, 256, 0, jb::EPI_STORE, 0},
        {300, 700, 130, 0, 0, 256, 0, jb::EPI_BIAS, 0},      {200, 300, 72, 0, 1, 256, 0, jb::EPI_STORE, 0},
    };
    for (const Case& c : cases) bad += run_case(c) != 0;
    printf("check: %d failing case(s)\n", bad);
  }
  if (!strcmp(what, "time") || !strcmp(what, "all")) {
    for (int pdl = 0; pdl < 2; ++pdl) {
      time_case(2, 512, 1024, 512, 0, 0, 64, 1, pdl);    // encoder / decoder wide layer, forward
      time_case(2, 512, 512, 1024, 0, 0, 64, 1, pdl);    // narrow layer, forward (K = 1024)
      time_case(2, 512, 512, 1024, 0, 0, 32, 1, pdl);
      time_case(2, 512, 1024, 512, 0, 1, 64, 1, pdl);    // dgrad
      time_case(2, 1024, 512, 512, 1, 1, 64, 0, pdl);    // wgrad (single pass)
      time_case(2, 512, 64, 512, 0, 0, 64, 1, pdl);      // heads
      time_case(2, 512, 512, 32, 0, 0, 64, 1, pdl);      // first decoder layer
      time_case(2, 512, 1024, 512, 0, 0, 64, 0, pdl);    // single pass for comparison
      time_case(2, 1024, 512, 512, 1, 1, 128, 0, pdl);   // wgrad, 128-wide tiles
      time_case(2, 1024, 512, 512, 1, 1, 256, 0, pdl);   // wgrad, 256-wide tiles
      time_case(8, 1024, 512, 512, 1, 1, 128, 0, pdl);   // the eight big wgrads of a step, 256 CTAs
      time_case(8, 1024, 512, 512, 1, 1, 256, 0, pdl);   // ... in one wave of 128 CTAs
      // split-K
      time_case(2, 512, 512, 1024, 0, 0, 64, 1, pdl, 2);
      time_case(2, 512, 512, 1024, 0, 1, 64, 1, pdl, 2);
      time_case(2, 512, 64, 512, 0, 0, 64, 1, pdl, 4);
      time_case(2, 512, 32, 512, 0, 1, 32, 1, pdl, 4);
      time_case(2, 512, 64, 512, 0, 0, 64, 1, pdl, 2);
    }
  }
  return bad ? 1 : 0;
}
