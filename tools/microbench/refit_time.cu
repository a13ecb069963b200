// Stand-alone timing of the refit's latency kernels on one SPD matrix (sm_100a):
//   * the diagonal-block kernel (modes 0 / 1 / 2) with clock64 stage stamps (GPO_DIAG_TIMING)
//   * trsm_panel_kernel and rank128_kernel<4,4>, each timed back to back (warm) and alternating with the others (what the chain does)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -DGPO_DIAG_TIMING -o refit_time refit_time.cu
// Output: JSON lines on stdout.  Timing: CUDA events around R launches.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../../gpyopt_b200/csrc/kernfun.cuh"
#include "../../gpyopt_b200/csrc/kernels.cuh"
#include "../../gpyopt_b200/csrc/refit_kernels.cuh"
using namespace gpo;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

int main(int argc, char** argv) {
  const int n_pad = argc > 1 ? atoi(argv[1]) : 2048, nb = n_pad / TILE;
  const size_t nn = (size_t)n_pad * n_pad;
  std::vector<double> h(nn);
  // SPD: exp(-|i-j|/50) + 0.1 I
  for (int i = 0; i < n_pad; ++i)
    for (int j = 0; j < n_pad; ++j) h[(size_t)i * n_pad + j] = exp(-fabs((double)i - j) / 50.0) + (i == j ? 0.1 : 0.0);
  double *A, *A0, *X, *U, *rdiag; int* info;
  CK(cudaMalloc(&A, nn * 8)); CK(cudaMalloc(&A0, nn * 8)); CK(cudaMalloc(&X, nn * 8)); CK(cudaMalloc(&U, nn * 8));
  CK(cudaMalloc(&rdiag, n_pad * 8)); CK(cudaMalloc(&info, 4));
  CK(cudaMemcpy(A0, h.data(), nn * 8, cudaMemcpyHostToDevice));
  CK(cudaMemset(X, 0, nn * 8)); CK(cudaMemset(U, 0, nn * 8)); CK(cudaMemset(info, 0, 4));
  CK(cudaFuncSetAttribute(diag_chol_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DIAG_SMEM_BYTES));
  CK(cudaFuncSetAttribute(trsm_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSM_SMEM_BYTES));
  CK(cudaFuncSetAttribute(rank128_kernel<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, rank128_smem_bytes<4, 4>()));
  cudaStream_t st; CK(cudaStreamCreate(&st));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int R = 20;
  auto reset = [&]() { CK(cudaMemcpyAsync(A, A0, nn * 8, cudaMemcpyDeviceToDevice, st)); };
  R128Params c; memset(&c, 0, sizeof(c));
  c.alpha = -1.0; c.beta = 1.0;
  c.A = A; c.lda = n_pad; c.a_row0 = TILE; c.a_col0 = 0;
  c.B = A; c.ldb = n_pad; c.b_row0 = TILE; c.b_col0 = 0;
  c.C = A; c.ldc = n_pad; c.c_row0 = TILE; c.c_col0 = TILE;
  c.tiles_m = (nb - 1) * 2; c.tiles_n = 2; c.skip_upper = 1;
  auto L_chol = [&](int mode) { diag_chol_inv_kernel<<<1, DIAG_THREADS, DIAG_SMEM_BYTES, st>>>(A, X, U, n_pad, 0, TILE, info, mode, rdiag); };
  auto L_trsm = [&]() { trsm_panel_kernel<<<dim3((nb - 1) * 4, 1), 256, TRSM_SMEM_BYTES, st>>>(A, X, n_pad, 0, 1); };
  auto L_upd = [&]() { rank128_kernel<4, 4><<<dim3(c.tiles_m * c.tiles_n, 1), 256, rank128_smem_bytes<4, 4>(), st>>>(c); };
  // --- each kernel back to back (the matrix degenerates after the first pass; timing does not depend on the values except
  //     through NaN/denormals, so the diagonal kernel always gets a fresh copy of its block) ---
  for (int mode = 0; mode < 3; ++mode) {
    float best = 1e9f;
    for (int rep = 0; rep < R; ++rep) {
      if (mode != 2) { CK(cudaMemcpy2DAsync(A, n_pad * 8, A0, n_pad * 8, TILE * 8, TILE, cudaMemcpyDeviceToDevice, st)); }
      CK(cudaEventRecord(e0, st)); L_chol(mode); CK(cudaEventRecord(e1, st)); CK(cudaStreamSynchronize(st));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = fminf(best, ms);
    }
    long long t[64]; CK(cudaMemcpyFromSymbol(t, g_diag_t, sizeof(t)));
    printf("{\"kernel\": \"diag mode %d\", \"best_us\": %.2f, \"load_cyc\": %lld, \"total_cyc\": %lld, \"stages\": [", mode, best * 1e3, t[0] - t[63], t[41] - t[63]);
    const int s0 = mode == 2 ? 4 : 0, s1 = mode == 1 ? 5 : 8;
    long long prev = t[0];
    for (int s = s0; s < s1; ++s) {
      if (s < 4) { printf("{\"s\": %d, \"store\": %lld, \"chol32\": %lld, \"solve\": %lld, \"update\": %lld}%s", s, t[1 + 4 * s] - prev, t[2 + 4 * s] - t[1 + 4 * s], t[3 + 4 * s] - t[2 + 4 * s], t[4 + 4 * s] - t[3 + 4 * s], s + 1 < s1 ? ", " : ""); }
      else { printf("{\"s\": %d, \"pre+solve\": %lld, \"update\": %lld}%s", s, t[3 + 4 * s] - prev, t[4 + 4 * s] - t[3 + 4 * s], s + 1 < s1 ? ", " : ""); }
      prev = t[4 + 4 * s];
    }
    printf("], \"store_cyc\": %lld}\n", t[41] - t[40]);
  }
  // leave a valid L_00 / D / rdiag behind for the panel kernel
  CK(cudaMemcpy2DAsync(A, n_pad * 8, A0, n_pad * 8, TILE * 8, TILE, cudaMemcpyDeviceToDevice, st));
  L_chol(1);
  auto time_it = [&](const char* name, auto&& fn, int reps) {
    for (int w = 0; w < 3; ++w) fn();
    CK(cudaEventRecord(e0, st));
    for (int rep = 0; rep < reps; ++rep) fn();
    CK(cudaEventRecord(e1, st)); CK(cudaStreamSynchronize(st));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("{\"kernel\": \"%s\", \"us_per_iteration\": %.2f}\n", name, ms * 1e3 / reps);
  };
  time_it("trsm back to back", [&]() { L_trsm(); }, R);
  time_it("rank128<4,4> column update back to back", [&]() { L_upd(); }, R);
  time_it("chol(mode 1) back to back", [&]() { L_chol(1); }, R);
  time_it("chol(1) + trsm + update (the chain of one block column)", [&]() { L_chol(1); L_trsm(); L_upd(); }, R);
  time_it("trsm + update alternating", [&]() { L_trsm(); L_upd(); }, R);
  // --- the chain kernel of block column 1 in isolation ---
  {
    double *raw, *Pg, *Tg; int* flags;
    CK(cudaMalloc(&raw, TILE * TILE * 8)); CK(cudaMalloc(&Pg, TILE * TILE * 8)); CK(cudaMalloc(&Tg, TILE * TILE * 8)); CK(cudaMalloc(&flags, 8));
    CK(cudaFuncSetAttribute(chol_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CHAIN_SMEM_BYTES));
    float best = 1e9f;
    for (int rep = 0; rep < R; ++rep) {
      CK(cudaMemcpyAsync(A, A0, nn * 8, cudaMemcpyDeviceToDevice, st));
      L_chol(1);
      tile_copy_kernel<<<dim3(16, 1), 256, 0, st>>>(A, raw, n_pad, 1, 0);
      CK(cudaMemsetAsync(flags, 0, 8, st));
      CK(cudaEventRecord(e0, st));
      chol_chain_kernel<<<CHAIN_CTAS, DIAG_THREADS, CHAIN_SMEM_BYTES, st>>>(A, X, U, n_pad, 1, TILE, info, rdiag, raw, Pg, Tg, flags, 1);
      CK(cudaEventRecord(e1, st)); CK(cudaStreamSynchronize(st));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = fminf(best, ms);
    }
    CK(cudaGetLastError());
    long long t[64]; CK(cudaMemcpyFromSymbol(t, g_diag_t, sizeof(t)));
    printf("{\"kernel\": \"chain kernel (block column 1)\", \"best_us\": %.2f", best * 1e3);
    const char* names[7] = {"load", "trsm", "publish+wait", "gather", "syrk", "publish", "wait"};
    for (int c = 0; c < 2; ++c) {
      const int o = c == 0 ? 44 : 54;
      printf(", \"cta%d\": {", c == 0 ? 0 : 7);
      for (int i = 0; i < (c == 0 ? 7 : 6); ++i) printf("\"%s\": %lld%s", names[i], t[o + i + 1] - t[o + i], i + 1 < (c == 0 ? 7 : 6) ? ", " : "");
      printf("}");
    }
    printf(", \"chol_cyc\": %lld}\n", t[41] - t[44 + 7]);
  }
  CK(cudaGetLastError());
  printf("{\"done\": true}\n");
  return 0;
}
