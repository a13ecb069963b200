// Common device helpers for the sm_100a kernels: mbarrier / TMA (cp.async.bulk.tensor) PTX wrappers,
// FP64 tensor-core MMA (mma.sync m8n8k4 f64 -> SASS DMMA.8x8x4), error macros.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

namespace gpo {

constexpr int TILE = 128;  // every matrix dimension in HBM is padded to a multiple of this

// ---- mbarrier ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// ---- TMA: 3-D tiled tensor load global -> shared, completion on an mbarrier ---------------------------
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// ---- TMA: 1-D bulk copy global -> shared (cp.async.bulk), completion on an mbarrier -------------------------
// dst / src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void tma_bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// L2 prefetch of a contiguous range (cp.async.bulk.prefetch.L2): a hint, no completion to wait for.  16-byte aligned, size a
// multiple of 16.
__device__ __forceinline__ void l2_prefetch_bulk(const void* gsrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// ---- FP64 tensor-core MMA -----------------------------------------------------------------------------
// D(8x8) += A(8x4, row) * B(4x8, col).  Lane (g = lane>>2, t = lane&3) supplies A[g][t], B[t][g] and owns
// D[g][2t], D[g][2t+1].
__device__ __forceinline__ void dmma_884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Row permutation inside an 8-row fragment group that makes 16-byte fragment loads from a
// SWIZZLE_128B tile bank-conflict free (lanes g, g+1 of a quarter-warp must differ in row bit 2).
__device__ __host__ __forceinline__ int frag_row(int g) { return ((g & 1) << 2) | (g >> 1); }

}  // namespace gpo

#define GPO_CUDA_OK(expr)                                                                         \
  do {                                                                                            \
    cudaError_t e__ = (expr);                                                                     \
    if (e__ != cudaSuccess) {                                                                     \
      snprintf(eng->last_error, sizeof(eng->last_error), "%s:%d %s: %s", __FILE__, __LINE__, #expr, \
               cudaGetErrorString(e__));                                                          \
      return GPO_ERR_CUDA;                                                                        \
    }                                                                                             \
  } while (0)
