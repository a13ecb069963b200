// Stated baseline for the refit: cuSOLVER dpotrf + dpotri on the same box, timing only.  This
// probe is never linked into libgpo_b200.so; the product's refit is refit_kernels.cuh.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cusolver_potrf cusolver_potrf.cu -lcusolver -lcublas
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <algorithm>
#include <cuda_runtime.h>
#include <cusolverDn.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
#define SK(x) do { cusolverStatus_t s = (x); if (s != CUSOLVER_STATUS_SUCCESS) { fprintf(stderr, "%s: status %d\n", #x, (int)s); exit(1); } } while (0)

int main(int argc, char** argv) {
  std::vector<int> ns;
  for (int i = 1; i < argc; ++i) ns.push_back(atoi(argv[i]));
  if (ns.empty()) ns = {512, 1024, 2048, 3072, 4096};
  cusolverDnHandle_t h; SK(cusolverDnCreate(&h));
  cudaStream_t st; CK(cudaStreamCreate(&st)); SK(cusolverDnSetStream(h, st));
  cudaEvent_t e0, e1, e2; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
  for (int n : ns) {
    // SPD test matrix: RBF Gram matrix of random points in [1, 10]^6, lengthscale 2, + 1e-2 I (the bench's refit problem)
    std::vector<double> X((size_t)n * 6), K((size_t)n * n);
    srand(1);
    for (auto& v : X) v = 1.0 + 9.0 * rand() / (double)RAND_MAX;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        double r2 = 0; for (int k = 0; k < 6; ++k) { double t = (X[(size_t)i * 6 + k] - X[(size_t)j * 6 + k]) / 2.0; r2 += t * t; }
        K[(size_t)i * n + j] = exp(-0.5 * r2) + (i == j ? 1e-2 : 0.0);
      }
    double *dK, *dA, *work; int* info; int lw1 = 0, lw2 = 0;
    CK(cudaMalloc(&dK, (size_t)n * n * 8)); CK(cudaMalloc(&dA, (size_t)n * n * 8)); CK(cudaMalloc(&info, 4));
    CK(cudaMemcpy(dK, K.data(), (size_t)n * n * 8, cudaMemcpyHostToDevice));
    SK(cusolverDnDpotrf_bufferSize(h, CUBLAS_FILL_MODE_LOWER, n, dA, n, &lw1));
    SK(cusolverDnDpotri_bufferSize(h, CUBLAS_FILL_MODE_LOWER, n, dA, n, &lw2));
    const int lw = std::max(lw1, lw2);
    CK(cudaMalloc(&work, (size_t)lw * 8));
    float best_f = 1e30f, best_i = 1e30f;
    for (int it = 0; it < 8; ++it) {
      CK(cudaMemcpyAsync(dA, dK, (size_t)n * n * 8, cudaMemcpyDeviceToDevice, st));
      CK(cudaEventRecord(e0, st));
      SK(cusolverDnDpotrf(h, CUBLAS_FILL_MODE_LOWER, n, dA, n, work, lw, info));
      CK(cudaEventRecord(e1, st));
      SK(cusolverDnDpotri(h, CUBLAS_FILL_MODE_LOWER, n, dA, n, work, lw, info));
      CK(cudaEventRecord(e2, st));
      CK(cudaStreamSynchronize(st));
      float a, b; CK(cudaEventElapsedTime(&a, e0, e1)); CK(cudaEventElapsedTime(&b, e1, e2));
      if (it >= 2) { best_f = std::min(best_f, a); best_i = std::min(best_i, b); }
    }
    int hinfo = -1; CK(cudaMemcpy(&hinfo, info, 4, cudaMemcpyDeviceToHost));
    printf("{\"n\": %d, \"cusolver_dpotrf_ms\": %.4f, \"cusolver_dpotri_ms\": %.4f, \"sum_ms\": %.4f, \"info\": %d, \"note\": \"device events, best of 6; dpotrf = L only, dpotri = Ky^-1 (no explicit L^-1)\"}\n",
           n, best_f, best_i, best_f + best_i, hinfo);
    cudaFree(dK); cudaFree(dA); cudaFree(work); cudaFree(info);
  }
  return 0;
}
