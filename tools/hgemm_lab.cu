// Stand-alone bring-up harness for the fp16-split tcgen05 GEMM (jamie_b200/csrc/hgemm.cuh) and the grid barrier of the
// whole-step kernel (no torch, no python):
//   check : every operand-major combination x mode x edge shape against a double-precision CPU result of the fp32
//           inputs (so the reported error includes the split), with full grids and with tiny grids (several tiles and
//           several ring geometries per CTA: the persistent pipeline state)
//   time  : the layer shapes of the training step, L2-warm back-to-back launches
//   bar   : microseconds per grid barrier of a 148 x 512-thread cooperative kernel
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/hgemm_lab tools/hgemm_lab.cu -lcuda
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../jamie_b200/csrc/hgemm.cuh"

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      return 1;                                                                        \
    }                                                                                  \
  } while (0)

static uint32_t rng_state = 12345u;
static float frand() {  // uniform (-1, 1)
  rng_state = rng_state * 1664525u + 1013904223u;
  return ((rng_state >> 8) * (1.0f / 8388608.0f)) - 1.0f;
}

struct Case {
  int M, N, K, a_mn, b_mn, bn, mode, epi, accumulate, ksplit;
  float scale_a, scale_b;   // magnitude of the operands (range behaviour of the fp16 split)
};

struct DevCase {
  std::vector<float> A, B, bias, C0;
  __half *dAh = nullptr, *dAl = nullptr, *dBh = nullptr, *dBl = nullptr;
  float *dC = nullptr, *dbias = nullptr;
  int lda, ldb, ldc;
  long long part_stride;
};

static int r8(int x) { return (x + 7) & ~7; }

static int setup_case(const Case& c, DevCase& d, jb::HgProblem* g) {
  const int M = c.M, N = c.N, K = c.K;
  d.lda = c.a_mn ? r8(M) : r8(K);
  d.ldb = c.b_mn ? r8(N) : r8(K);
  d.ldc = (N + 3) & ~3;
  const size_t na = static_cast<size_t>(c.a_mn ? K : M) * d.lda, nb = static_cast<size_t>(c.b_mn ? K : N) * d.ldb;
  d.A.assign(na, 0.f); d.B.assign(nb, 0.f); d.bias.resize(N);
  d.part_stride = static_cast<long long>(M) * d.ldc;
  d.C0.resize(static_cast<size_t>(d.part_stride) * c.ksplit);
  auto ai = [&](int m, int k) { return c.a_mn ? static_cast<size_t>(k) * d.lda + m : static_cast<size_t>(m) * d.lda + k; };
  auto bi = [&](int n, int k) { return c.b_mn ? static_cast<size_t>(k) * d.ldb + n : static_cast<size_t>(n) * d.ldb + k; };
  for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) d.A[ai(m, k)] = frand() * c.scale_a;
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) d.B[bi(n, k)] = frand() * c.scale_b;
  for (int n = 0; n < N; ++n) d.bias[n] = frand();
  for (auto& v : d.C0) v = frand();
  std::vector<__half> Ah(na), Al(na), Bh(nb), Bl(nb);
  for (size_t i = 0; i < na; ++i) { Ah[i] = __float2half_rn(d.A[i]); Al[i] = __float2half_rn((d.A[i] - __half2float(Ah[i])) * 2048.f); }
  for (size_t i = 0; i < nb; ++i) { Bh[i] = __float2half_rn(d.B[i]); Bl[i] = __float2half_rn((d.B[i] - __half2float(Bh[i])) * 2048.f); }
  CK(cudaMalloc(&d.dAh, na * 2)); CK(cudaMalloc(&d.dAl, na * 2)); CK(cudaMalloc(&d.dBh, nb * 2)); CK(cudaMalloc(&d.dBl, nb * 2));
  CK(cudaMalloc(&d.dC, d.C0.size() * 4)); CK(cudaMalloc(&d.dbias, N * 4));
  CK(cudaMemcpy(d.dAh, Ah.data(), na * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d.dAl, Al.data(), na * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d.dBh, Bh.data(), nb * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d.dBl, Bl.data(), nb * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d.dC, d.C0.data(), d.C0.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d.dbias, d.bias.data(), N * 4, cudaMemcpyHostToDevice));
  int rc = jb::hg_problem_fill(g, jb::HPlanes{d.dAh, d.dAl}, d.lda, c.a_mn, jb::HPlanes{d.dBh, d.dBl}, d.ldb, c.b_mn, d.dC, d.ldc, M, N, K,
                               c.bn, c.mode, c.epi, d.dbias, c.ksplit, d.part_stride, c.accumulate, 0.5f);
  if (rc) { printf("problem fill failed %d\n", rc); return 1; }
  return 0;
}
static void free_case(DevCase& d) {
  cudaFree(d.dAh); cudaFree(d.dAl); cudaFree(d.dBh); cudaFree(d.dBl); cudaFree(d.dC); cudaFree(d.dbias);
}

// compares the summed split-K partials with the double-precision result; returns relative error or < 0 on CUDA failure
static double verify_case(const Case& c, DevCase& d, const jb::HgProblem& g) {
  const int M = c.M, N = c.N, K = c.K;
  std::vector<float> C(d.C0.size());
  if (cudaMemcpy(C.data(), d.dC, C.size() * 4, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  auto ai = [&](int m, int k) { return c.a_mn ? static_cast<size_t>(k) * d.lda + m : static_cast<size_t>(m) * d.lda + k; };
  auto bi = [&](int n, int k) { return c.b_mn ? static_cast<size_t>(k) * d.ldb + n : static_cast<size_t>(n) * d.ldb + k; };
  double err2 = 0, ref2 = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += static_cast<double>(d.A[ai(m, k)]) * d.B[bi(n, k)];
      s *= 0.5;   // out_scale
      if (c.epi != jb::EPI_STORE) s += d.bias[n];
      if (c.epi == jb::EPI_BIAS_LRELU) s = s > 0 ? s : 0.01 * s;
      double got = 0;
      for (int p = 0; p < g.ksplit; ++p) got += C[static_cast<size_t>(p) * d.part_stride + static_cast<size_t>(m) * d.ldc + n];
      if (c.accumulate)
        for (int p = 0; p < g.ksplit; ++p) s += d.C0[static_cast<size_t>(p) * d.part_stride + static_cast<size_t>(m) * d.ldc + n];
      const double e = got - s;
      err2 += e * e; ref2 += s * s;
    }
  return sqrt(err2 / (ref2 + 1e-300));
}

static const char* mode_name(int m) { return m == jb::HG_SINGLE ? "single " : (m == jb::HG_PRECISE ? "precise" : "medium "); }

static int run_case(const Case& c, int max_ctas) {
  DevCase d;
  jb::HgProblem g;
  if (setup_case(c, d, &g)) return 1;
  jb::HgPhase ph = jb::hg_phase_finalize(&g, 0, 1);
  jb::HgProblem* dg;
  CK(cudaMalloc(&dg, sizeof g));
  CK(cudaMemcpy(dg, &g, sizeof g, cudaMemcpyHostToDevice));
  CK(jb::hgemm_launch_phase(dg, ph, max_ctas, 0));
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
  const double rel = verify_case(c, d, g);
  // LeakyReLU epilogues amplify relative error near zero crossings a little; accumulate adds fp32 rounding of C0
  const double tol = c.mode == jb::HG_SINGLE ? 8e-4 : (c.mode == jb::HG_PRECISE ? 4e-7 : 3e-6);
  const bool ok = rel >= 0 && rel < tol;
  printf("%s %s M%-4d N%-4d K%-4d a_mn%d b_mn%d bn%-3d epi%d acc%d ks%d ctas%-3d sa %.0e sb %.0e : rel %.3e %s\n", ok ? "ok  " : "FAIL",
         mode_name(c.mode), c.M, c.N, c.K, c.a_mn, c.b_mn, c.bn, c.epi, c.accumulate, g.ksplit, max_ctas, c.scale_a, c.scale_b, rel,
         ok ? "" : "<<<<<<");
  free_case(d); cudaFree(dg);
  return ok ? 0 : 2;
}

// several problems of different geometry in ONE phase on a tiny grid: every CTA walks tiles of several ring geometries
static int run_mixed(int max_ctas) {
  const Case cs[] = {
      {300, 200, 520, 0, 0, 64, jb::HG_PRECISE, jb::EPI_BIAS, 0, 2, 1.f, 1.f},
      {256, 512, 200, 1, 1, 256, jb::HG_MEDIUM, jb::EPI_STORE, 1, 1, 1.f, 1.f},
      {200, 96, 130, 0, 1, 32, jb::HG_PRECISE, jb::EPI_STORE, 0, 1, 1.f, 1.f},
      {384, 320, 72, 0, 0, 128, jb::HG_SINGLE, jb::EPI_BIAS_LRELU, 0, 1, 1.f, 1.f},
      {130, 130, 700, 1, 0, 64, jb::HG_PRECISE, jb::EPI_STORE, 0, 3, 1.f, 1.f},
      {512, 64, 512, 1, 1, 64, jb::HG_MEDIUM, jb::EPI_STORE, 0, 1, 1.f, 1.f},
  };
  const int n = sizeof(cs) / sizeof(cs[0]);
  std::vector<DevCase> d(n);
  std::vector<jb::HgProblem> g(n);
  for (int i = 0; i < n; ++i) if (setup_case(cs[i], d[i], &g[i])) return 1;
  jb::HgPhase ph = jb::hg_phase_finalize(g.data(), 0, n);
  jb::HgProblem* dg;
  CK(cudaMalloc(&dg, n * sizeof(jb::HgProblem)));
  CK(cudaMemcpy(dg, g.data(), n * sizeof(jb::HgProblem), cudaMemcpyHostToDevice));
  CK(jb::hgemm_launch_phase(dg, ph, max_ctas, 0));
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("mixed phase failed: %s\n", cudaGetErrorString(e)); return 1; }
  int bad = 0;
  for (int i = 0; i < n; ++i) {
    const double rel = verify_case(cs[i], d[i], g[i]);
    const double tol = cs[i].mode == jb::HG_SINGLE ? 8e-4 : (cs[i].mode == jb::HG_PRECISE ? 4e-7 : 3e-6);
    const bool ok = rel >= 0 && rel < tol;
    printf("%s mixed[%d] %s ctas%-3d tiles %d : rel %.3e\n", ok ? "ok  " : "FAIL", i, mode_name(cs[i].mode), max_ctas, ph.total_tiles, rel);
    bad += !ok;
    free_case(d[i]);
  }
  cudaFree(dg);
  return bad;
}

// nprob identical problems in one launch (the two modalities of a stage), L2-warm, back-to-back
static int time_case(int nprob, int M, int N, int K, int a_mn, int b_mn, int bn, int mode, int ksplit) {
  const int lda = a_mn ? M : K, ldb = b_mn ? N : K;
  std::vector<jb::HgProblem> g(nprob);
  std::vector<void*> bufs;
  for (int i = 0; i < nprob; ++i) {
    __half *dA, *dAl, *dB, *dBl; float* dC;
    CK(cudaMalloc(&dA, static_cast<size_t>(M) * K * 2)); CK(cudaMalloc(&dAl, static_cast<size_t>(M) * K * 2));
    CK(cudaMalloc(&dB, static_cast<size_t>(N) * K * 2)); CK(cudaMalloc(&dBl, static_cast<size_t>(N) * K * 2));
    CK(cudaMalloc(&dC, static_cast<size_t>(M) * N * 4 * ksplit));
    CK(cudaMemset(dA, 0, static_cast<size_t>(M) * K * 2)); CK(cudaMemset(dAl, 0, static_cast<size_t>(M) * K * 2));
    CK(cudaMemset(dB, 0, static_cast<size_t>(N) * K * 2)); CK(cudaMemset(dBl, 0, static_cast<size_t>(N) * K * 2));
    bufs.insert(bufs.end(), {dA, dAl, dB, dBl, dC});
    if (jb::hg_problem_fill(&g[i], jb::HPlanes{dA, dAl}, lda, a_mn, jb::HPlanes{dB, dBl}, ldb, b_mn, dC, N, M, N, K, bn, mode, jb::EPI_STORE,
                            nullptr, ksplit, static_cast<long long>(M) * N, 0, 1.f)) { printf("fill failed\n"); return 1; }
  }
  jb::HgPhase ph = jb::hg_phase_finalize(g.data(), 0, nprob);
  jb::HgProblem* dg;
  CK(cudaMalloc(&dg, nprob * sizeof(jb::HgProblem)));
  CK(cudaMemcpy(dg, g.data(), nprob * sizeof(jb::HgProblem), cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 5; ++i) CK(jb::hgemm_launch_phase(dg, ph, 148, 0));
  const int iters = 200;
  CK(cudaEventRecord(e0));
  for (int i = 0; i < iters; ++i) CK(jb::hgemm_launch_phase(dg, ph, 148, 0));
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  const double us = ms * 1e3 / iters;
  const double tf = 2.0 * nprob * M * N * K / (us * 1e-6) / 1e12;
  const int planes = mode == jb::HG_SINGLE ? 1 : 2;
  const double mb = static_cast<double>(ph.total_tiles) * ((K + 63) / 64 / ksplit) * (128 + bn) * 128.0 * planes / 1e6;
  printf("time: %d x [M%d N%d K%d] a_mn%d b_mn%d bn%-3d %s ks%d tiles %-3d : %6.2f us/launch  %6.1f TFLOP/s (algorithmic)  L2->SM %.1f MB = %.2f TB/s\n",
         nprob, M, N, K, a_mn, b_mn, bn, mode_name(mode), ksplit, ph.total_tiles, us, tf, mb, mb / us);
  for (void* p : bufs) cudaFree(p);
  cudaFree(dg);
  return 0;
}

// ------------------------------------------------------------------------------------------------ grid barrier bench
__global__ void __launch_bounds__(512, 1) k_bar_bench(unsigned int* counter, int n, float* sink, int variant, float* scratch) {
  extern __shared__ uint8_t dummy_smem[];
  unsigned int target = 0;
  float acc = 0.f;
  for (int i = 0; i < n; ++i) {
    acc += static_cast<float>(i) * 1e-9f;
    target += gridDim.x;
    if (variant >= 2) scratch[(blockIdx.x * 512 + threadIdx.x) * 4 + (i & 3)] = acc;   // a global store in flight per thread
    if (variant == 0) jb::grid_barrier_v0(counter, target);
    else jb::grid_barrier(counter, target);
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) *sink = acc + dummy_smem[0];
}
static int bar_bench() {
  unsigned int* ctr; float* sink; float* scratch;
  CK(cudaMalloc(&ctr, 128)); CK(cudaMalloc(&sink, 4)); CK(cudaMalloc(&scratch, 148 * 512 * 16 * 4));
  CK(cudaFuncSetAttribute(k_bar_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, jb::HG_SMEM_BYTES));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  for (int variant = 0; variant < 3; ++variant)
  for (int grid : {sms, 32}) {
    for (int rep = 0; rep < 2; ++rep) {
      int n = 2000;
      CK(cudaMemset(ctr, 0, 128));
      cudaEvent_t e0, e1;
      CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      void* args[] = {&ctr, &n, &sink, &variant, &scratch};
      CK(cudaEventRecord(e0));
      CK(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(k_bar_bench), dim3(grid), dim3(512), args, jb::HG_SMEM_BYTES, 0));
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep) printf("bar: variant %d grid %3d x 512 threads: %.3f us per grid barrier (%d barriers)\n", variant, grid, ms * 1e3 / n, n);
    }
  }
  cudaFree(ctr); cudaFree(sink);
  return 0;
}

int main(int argc, char** argv) {
  const char* what = argc > 1 ? argv[1] : "all";
  int bad = 0;
  if (!strcmp(what, "check") || !strcmp(what, "all")) {
    const Case cases[] = {
        // single pass, every major combination, ragged edges
        {256, 128, 96, 0, 0, 64, jb::HG_SINGLE, jb::EPI_BIAS, 0, 1, 1.f, 1.f},
        {200, 100, 72, 0, 1, 64, jb::HG_SINGLE, jb::EPI_STORE, 0, 1, 1.f, 1.f},
        {130, 39, 300, 1, 1, 64, jb::HG_SINGLE, jb::EPI_STORE, 1, 1, 1.f, 1.f},
        {128, 32, 32, 1, 0, 32, jb::HG_SINGLE, jb::EPI_BIAS_LRELU, 0, 1, 1.f, 1.f},
        {300, 2000, 40, 0, 0, 128, jb::HG_SINGLE, jb::EPI_BIAS_LRELU, 0, 1, 1.f, 1.f},
        {300, 700, 130, 0, 0, 256, jb::HG_SINGLE, jb::EPI_BIAS, 0, 1, 1.f, 1.f},
        // precise: the shapes of the training step (forward K,K; dgrad K,MN), split-K
        {512, 1024, 512, 0, 0, 64, jb::HG_PRECISE, jb::EPI_BIAS, 0, 1, 1.f, 0.04f},
        {512, 512, 1024, 0, 0, 64, jb::HG_PRECISE, jb::EPI_BIAS, 0, 2, 1.f, 0.03f},
        {512, 64, 512, 0, 0, 64, jb::HG_PRECISE, jb::EPI_BIAS, 0, 4, 1.f, 0.04f},
        {512, 512, 32, 0, 0, 64, jb::HG_PRECISE, jb::EPI_BIAS, 0, 1, 1.f, 0.17f},
        {512, 1024, 512, 0, 1, 64, jb::HG_PRECISE, jb::EPI_STORE, 0, 1, 1.f, 0.03f},
        {512, 32, 512, 0, 1, 32, jb::HG_PRECISE, jb::EPI_STORE, 0, 8, 1.f, 0.17f},
        {512, 512, 64, 0, 1, 64, jb::HG_PRECISE, jb::EPI_STORE, 0, 1, 1.f, 0.04f},
        {300, 1000, 2000, 0, 0, 64, jb::HG_PRECISE, jb::EPI_BIAS, 0, 2, 1.f, 0.02f},
        {300, 78, 39, 0, 1, 64, jb::HG_PRECISE, jb::EPI_STORE, 0, 1, 1.f, 1.f},
        {700, 64, 1302, 0, 0, 64, jb::HG_PRECISE, jb::EPI_STORE, 0, 3, 1.f, 1.f},
        {700, 1302, 64, 0, 1, 64, jb::HG_PRECISE, jb::EPI_BIAS, 0, 1, 1.f, 1.f},
        {130, 39, 300, 1, 1, 64, jb::HG_PRECISE, jb::EPI_STORE, 1, 1, 1.f, 1.f},
        {100, 130, 70, 1, 0, 32, jb::HG_PRECISE, jb::EPI_STORE, 0, 1, 1.f, 1.f},
        {64, 40, 24, 0, 0, 32, jb::HG_PRECISE, jb::EPI_BIAS, 0, 1, 1.f, 1.f},
        // operand range: small gradients (loss-scaled), large activations
        {512, 512, 512, 0, 1, 64, jb::HG_PRECISE, jb::EPI_STORE, 0, 1, 1e-3f, 0.04f},
        {512, 512, 512, 0, 0, 64, jb::HG_PRECISE, jb::EPI_STORE, 0, 1, 3e3f, 0.04f},
        {512, 512, 512, 0, 0, 64, jb::HG_PRECISE, jb::EPI_STORE, 0, 1, 1.f, 2e-3f},
        // medium: weight gradients (MN,MN; K = batch), wide tiles
        {1024, 512, 512, 1, 1, 256, jb::HG_MEDIUM, jb::EPI_STORE, 1, 1, 1.f, 1.f},
        {512, 1024, 512, 1, 1, 256, jb::HG_MEDIUM, jb::EPI_STORE, 0, 1, 1.f, 1.f},
        {64, 512, 512, 1, 1, 128, jb::HG_MEDIUM, jb::EPI_STORE, 0, 1, 1.f, 1.f},
        {512, 32, 512, 1, 1, 32, jb::HG_MEDIUM, jb::EPI_STORE, 0, 1, 1.f, 1.f},
        {78, 39, 300, 1, 1, 64, jb::HG_MEDIUM, jb::EPI_STORE, 0, 1, 1.f, 1.f},
        {2000, 1000, 300, 1, 1, 256, jb::HG_MEDIUM, jb::EPI_STORE, 0, 1, 1.f, 1.f},
        {300, 200, 130, 0, 0, 128, jb::HG_MEDIUM, jb::EPI_BIAS, 0, 2, 1.f, 1.f},
    };
    for (const Case& c : cases) bad += run_case(c, 148) != 0;
    // persistent pipeline state: few CTAs, many tiles each
    const Case pc[] = {
        {512, 1024, 512, 0, 0, 64, jb::HG_PRECISE, jb::EPI_BIAS, 0, 1, 1.f, 0.04f},
        {512, 512, 1024, 0, 1, 64, jb::HG_PRECISE, jb::EPI_STORE, 0, 2, 1.f, 0.03f},
        {1024, 512, 512, 1, 1, 256, jb::HG_MEDIUM, jb::EPI_STORE, 0, 1, 1.f, 1.f},
        {512, 512, 200, 0, 0, 128, jb::HG_SINGLE, jb::EPI_BIAS, 0, 1, 1.f, 1.f},
        {512, 64, 150, 0, 0, 32, jb::HG_PRECISE, jb::EPI_BIAS, 0, 1, 1.f, 1.f},
    };
    for (const Case& c : pc) { bad += run_case(c, 5) != 0; bad += run_case(c, 1) != 0; }
    bad += run_mixed(148);
    bad += run_mixed(3);
    bad += run_mixed(1);
    printf("check: %d failing case(s)\n", bad);
  }
  if (!strcmp(what, "time") || !strcmp(what, "all")) {
    time_case(2, 512, 1024, 512, 0, 0, 64, jb::HG_PRECISE, 1);    // encoder / decoder wide layer, forward
    time_case(2, 512, 512, 1024, 0, 0, 64, jb::HG_PRECISE, 2);    // narrow layer, forward (K = 1024), split-K 2
    time_case(2, 512, 512, 1024, 0, 0, 64, jb::HG_PRECISE, 1);
    time_case(2, 512, 1024, 512, 0, 1, 64, jb::HG_PRECISE, 1);    // dgrad
    time_case(2, 512, 512, 1024, 0, 1, 64, jb::HG_PRECISE, 2);
    time_case(2, 512, 64, 512, 0, 0, 64, jb::HG_PRECISE, 8);      // heads
    time_case(2, 512, 64, 512, 0, 0, 64, jb::HG_PRECISE, 4);
    time_case(2, 512, 512, 32, 0, 0, 64, jb::HG_PRECISE, 1);      // first decoder layer
    time_case(2, 512, 32, 512, 0, 1, 32, jb::HG_PRECISE, 8);      // its dgrad
    time_case(2, 512, 1024, 512, 0, 0, 64, jb::HG_SINGLE, 1);     // single pass for comparison
    time_case(2, 512, 1024, 512, 0, 0, 128, jb::HG_SINGLE, 1);
    time_case(2, 1024, 512, 512, 1, 1, 256, jb::HG_MEDIUM, 1);    // wgrad
    time_case(2, 1024, 512, 512, 1, 1, 128, jb::HG_MEDIUM, 1);
    time_case(8, 1024, 512, 512, 1, 1, 256, jb::HG_MEDIUM, 1);    // the eight big wgrads of a step in one wave
    time_case(8, 1024, 512, 512, 1, 1, 128, jb::HG_MEDIUM, 1);
    time_case(8, 1024, 512, 512, 1, 1, 256, jb::HG_SINGLE, 1);
  }
  if (!strcmp(what, "bar") || !strcmp(what, "all")) bad += bar_bench();
  return bad ? 1 : 0;
}
