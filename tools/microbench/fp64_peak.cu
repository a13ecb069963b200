// FP64 peak probes for B200 (sm_100a): cuBLAS DGEMM, raw DMMA (mma.sync f64) shapes, raw DFMA, mixed.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o fp64_peak fp64_peak.cu -lcublas
// Output: one JSON object on stdout. Timing: CUDA events, best of R after warm-up.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cublas_v2.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
__device__ __forceinline__ void dmma1684(double (&c)[4], const double (&a)[2], double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(b));
}

// ILP independent accumulator chains per warp, iters iterations. flops/warp/iter = ILP*2*M*N*K
template <int ILP>
__global__ void k_dmma884(double *out, int iters, double seed) {
  double c0[ILP], c1[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) { c0[i] = 0.0; c1[i] = 0.0; }
  double a = seed + threadIdx.x * 1e-9, b = seed * 0.5;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) dmma884(c0[i], c1[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += c0[i] + c1[i];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void k_dmma16816(double *out, int iters, double seed) {
  double c[ILP][4];
#pragma unroll
  for (int i = 0; i < ILP; i++) for (int j = 0; j < 4; j++) c[i][j] = 0.0;
  double a[8], b[4];
  for (int j = 0; j < 8; j++) a[j] = seed + threadIdx.x * 1e-9 + j;
  for (int j = 0; j < 4; j++) b[j] = seed * 0.5 + j;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) dmma16816(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void k_dmma1688(double *out, int iters, double seed) {
  double c[ILP][4];
#pragma unroll
  for (int i = 0; i < ILP; i++) for (int j = 0; j < 4; j++) c[i][j] = 0.0;
  double a[4], b[2];
  for (int j = 0; j < 4; j++) a[j] = seed + threadIdx.x * 1e-9 + j;
  for (int j = 0; j < 2; j++) b[j] = seed * 0.5 + j;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) dmma1688(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void k_dmma1684(double *out, int iters, double seed) {
  double c[ILP][4];
#pragma unroll
  for (int i = 0; i < ILP; i++) for (int j = 0; j < 4; j++) c[i][j] = 0.0;
  double a[2], b;
  for (int j = 0; j < 2; j++) a[j] = seed + threadIdx.x * 1e-9 + j;
  b = seed * 0.5;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) dmma1684(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void k_dfma(double *out, int iters, double seed) {
  double c[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) c[i] = i;
  double a = seed + threadIdx.x * 1e-9, b = seed * 0.5;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) c[i] = fma(a, c[i], b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += c[i];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mixed: per iteration ILP dmma884 + NF dfma per thread
template <int ILP, int NF>
__global__ void k_mixed(double *out, int iters, double seed) {
  double c0[ILP], c1[ILP], f[NF];
#pragma unroll
  for (int i = 0; i < ILP; i++) { c0[i] = 0.0; c1[i] = 0.0; }
#pragma unroll
  for (int i = 0; i < NF; i++) f[i] = i;
  double a = seed + threadIdx.x * 1e-9, b = seed * 0.5;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) dmma884(c0[i], c1[i], a, b);
#pragma unroll
    for (int i = 0; i < NF; i++) f[i] = fma(a, f[i], b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += c0[i] + c1[i];
#pragma unroll
  for (int i = 0; i < NF; i++) s += f[i];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// smem-fed m8n8k4: warp tile 64x32 (8 A frags, 4 B frags, 32 mma per k4), fragments read from smem each step.
__global__ void k_dmma_smem(double *out, int iters, double seed) {
  extern __shared__ double sm[];
  // A tile: 128 rows x 16 k (pitch 20), B tile: 128 rows x 16 k (pitch 20)
  const int PITCH = 20;
  double *sA = sm, *sB = sm + 128 * PITCH;
  for (int i = threadIdx.x; i < 2 * 128 * PITCH; i += blockDim.x) sm[i] = seed + i * 1e-6;
  __syncthreads();
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int wm = (warp & 1) * 64, wn = (warp >> 1) * 32;
  int g = lane >> 2, t = lane & 3;
  double c[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) { c[i][j][0] = 0; c[i][j][1] = 0; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int ks = 0; ks < 4; ks++) {
      double a[8], b[4];
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = sA[(wm + i * 8 + g) * PITCH + ks * 4 + t];
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = sB[(wn + j * 8 + g) * PITCH + ks * 4 + t];
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) dmma884(c[i][j][0], c[i][j][1], a[i], b[j]);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) s += c[i][j][0] + c[i][j][1];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double time_best(F launch, int reps) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch(); CK(cudaDeviceSynchronize());
  double best = 1e30;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best * 1e-3;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  double *out; CK(cudaMalloc(&out, 1 << 24));
  printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
  const int iters = 20000;
  // raw probes: blocks = sms * bps, threads = 32 * wpb
  struct Cfg { int bps, wpb; } cfgs[] = {{1, 4}, {1, 8}, {1, 16}, {2, 16}};
  for (auto c : cfgs) {
    int blocks = sms * c.bps, thr = 32 * c.wpb; double warps = (double)blocks * c.wpb;
    double t;
    t = time_best([&] { k_dmma884<8><<<blocks, thr>>>(out, iters, 1.0); }, 5);
    printf(",\n \"dmma884_ilp8_w%d_tf\": %.3f", c.bps * c.wpb, warps * iters * 8 * 2.0 * 8 * 8 * 4 / t * 1e-12);
    t = time_best([&] { k_dmma884<16><<<blocks, thr>>>(out, iters, 1.0); }, 5);
    printf(", \"dmma884_ilp16_w%d_tf\": %.3f", c.bps * c.wpb, warps * iters * 16 * 2.0 * 8 * 8 * 4 / t * 1e-12);
    t = time_best([&] { k_dmma1684<8><<<blocks, thr>>>(out, iters, 1.0); }, 5);
    printf(", \"dmma1684_ilp8_w%d_tf\": %.3f", c.bps * c.wpb, warps * iters * 8 * 2.0 * 16 * 8 * 4 / t * 1e-12);
    t = time_best([&] { k_dmma1688<8><<<blocks, thr>>>(out, iters, 1.0); }, 5);
    printf(", \"dmma1688_ilp8_w%d_tf\": %.3f", c.bps * c.wpb, warps * iters * 8 * 2.0 * 16 * 8 * 8 / t * 1e-12);
    t = time_best([&] { k_dmma16816<8><<<blocks, thr>>>(out, iters / 2, 1.0); }, 5);
    printf(", \"dmma16816_ilp8_w%d_tf\": %.3f", c.bps * c.wpb, warps * (iters / 2) * 8 * 2.0 * 16 * 8 * 16 / t * 1e-12);
    t = time_best([&] { k_dfma<16><<<blocks, thr>>>(out, iters, 1.0); }, 5);
    printf(", \"dfma_ilp16_w%d_tf\": %.3f", c.bps * c.wpb, warps * 32 * iters * 16 * 2.0 / t * 1e-12);
    t = time_best([&] { k_mixed<8, 8><<<blocks, thr>>>(out, iters, 1.0); }, 5);
    printf(", \"mixed_8mma_8fma_w%d_tf\": %.3f", c.bps * c.wpb, warps * iters * (8 * 2.0 * 256 + 32 * 8 * 2.0) / t * 1e-12);
    t = time_best([&] { k_mixed<8, 32><<<blocks, thr>>>(out, iters, 1.0); }, 5);
    printf(", \"mixed_8mma_32fma_w%d_tf\": %.3f", c.bps * c.wpb, warps * iters * (8 * 2.0 * 256 + 32 * 32 * 2.0) / t * 1e-12);
  }
  {
    int smem = 2 * 128 * 20 * 8;
    CK(cudaFuncSetAttribute(k_dmma_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int bps = 1; bps <= 2; bps++) {
      int blocks = sms * bps; int it2 = 4000;
      double t = time_best([&] { k_dmma_smem<<<blocks, 256, smem>>>(out, it2, 1.0); }, 5);
      printf(",\n \"dmma884_smemfed_64x32_8w_bps%d_tf\": %.3f", bps, (double)blocks * 8 * it2 * 4 * 32 * 512.0 / t * 1e-12);
    }
  }
  // cuBLAS DGEMM
  {
    cublasHandle_t h; cublasCreate(&h);
    for (int n : {4096, 8192}) {
      double *A, *B, *C; size_t bytes = (size_t)n * n * 8;
      CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&B, bytes)); CK(cudaMalloc(&C, bytes));
      CK(cudaMemset(A, 0, bytes)); CK(cudaMemset(B, 0, bytes));
      std::vector<double> hA((size_t)n * n);
      for (size_t i = 0; i < hA.size(); i++) hA[i] = (double)rand() / RAND_MAX - 0.5;
      CK(cudaMemcpy(A, hA.data(), bytes, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(B, hA.data(), bytes, cudaMemcpyHostToDevice));
      double one = 1.0, zero = 0.0;
      double t = time_best([&] { cublasDgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n); }, 10);
      printf(",\n \"cublas_dgemm_tn_%d_tf\": %.3f", n, 2.0 * n * n * n / t * 1e-12);
      t = time_best([&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n); }, 10);
      printf(", \"cublas_dgemm_nn_%d_tf\": %.3f", n, 2.0 * n * n * n / t * 1e-12);
      // sustained 3 s
      if (n == 8192) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        int reps = 100; cudaEventRecord(e0);
        for (int r = 0; r < reps; r++) cublasDgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n);
        cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf(", \"cublas_dgemm_tn_8192_sustained_tf\": %.3f", reps * 2.0 * n * n * n / (ms * 1e-3) * 1e-12);
      }
      // our shape: M=2048 (n_obs), N=16384 (chunk), K=2048
      cudaFree(A); cudaFree(B); cudaFree(C);
    }
    {
      int M = 2048, N = 16384, K = 2048; double *A, *B, *C;
      CK(cudaMalloc(&A, (size_t)M * K * 8)); CK(cudaMalloc(&B, (size_t)K * N * 8)); CK(cudaMalloc(&C, (size_t)M * N * 8));
      CK(cudaMemset(A, 0, (size_t)M * K * 8)); CK(cudaMemset(B, 0, (size_t)K * N * 8));
      double one = 1.0, zero = 0.0;
      double t = time_best([&] { cublasDgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, M, N, K, &one, A, K, B, K, &zero, C, M); }, 10);
      printf(",\n \"cublas_dgemm_tn_2048x16384x2048_tf\": %.3f", 2.0 * M * N * K / t * 1e-12);
    }
  }
  printf("\n}\n");
  return 0;
}
