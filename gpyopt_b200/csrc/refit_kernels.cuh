// Latency kernels of the refit (K4): what sits on, or right next to, the critical chain of the blocked Cholesky
//   diagonal block kb  ->  panel (rows below it)  ->  update of block column kb+1  ->  diagonal block kb+1 ...
// The persistent TMA/DMMA GEMM (dgemm_tma.cuh) is built for long products; a K = 128 product of a handful of tiles spends
// two thirds of its 24-31 us in fixed cost (descriptor fetch, ring fill, 128-wide tiles on one SM each).  The two kernels
// here take those products with small tiles, plain cp.async staging and register-tiled FP64 FMAs (the DFMA pipe reaches
// the DMMA rate on B200: tools/microbench/fp64_peak.cu), so that a launch is one short wave over many SMs.
//
// Reference arithmetic they replace: GPy's jitchol (LAPACK dpotrf) and dtrtri / dpotri behind
// /root/reference/GPyOpt/models/gpmodel.py:78-93 (model.optimize -> every likelihood evaluation refactorises Ky).
#pragma once
#include "common.cuh"

namespace gpo {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Shared-memory row pitch (doubles) of the K-major operand tiles: rows stay 16-byte aligned and sixteen consecutive rows
// land in eight distinct 16-byte slots of a 128-byte line, i.e. the LDS.128 of a quarter-warp reading (row tx, k..k+1) is
// conflict-free (1040 B = 8 x 128 + 16).
constexpr int R128_LP = 130;

// ---------------------------------------------------------------------------------------------------------------
// C[m][n] = beta C[m][n] + alpha sum_{k < 128} A[m][k] B[n][k]      (both operands K-major, K = 128 exactly)
// optionally also Ct[n][m] = the same value (the transposed copy the next product reads as its K-major operand).
// CTA tile (16 MI) x (16 NJ), 256 threads; thread (ty, tx) owns rows ty + 16 i, columns tx + 16 j.
//   <4, 4>: 64 x 64   (block-column updates of the Cholesky, right-hand-side updates of the interleaved inverse)
//   <8, 2>: 128 x 32  (products that overwrite their own B operand through Ct: the CTA reads all 128 k of its 32 rows
//                      before it writes all 128 m of the same rows)
// The operands arrive in two cp.async groups (k < 64, k >= 64); the second lands under the first half's FMAs.
// Accumulation order is fixed (k ascending, one FMA chain per output): bit-reproducible.
// ---------------------------------------------------------------------------------------------------------------
struct R128Params {
  const double* A; long long lda, a_z; int a_row0, a_col0;    // A[m][k] = A[a_z z + (a_row0 + m) lda + a_col0 + k]
  const double* B; long long ldb, b_z; int b_row0, b_col0;    // B[n][k]
  double* C; long long ldc, c_z; int c_row0, c_col0;          // may be null
  double* Ct; long long ldct, ct_z; int ct_row0, ct_col0;     // may be null: Ct[ct_z z + (ct_row0 + n) ldct + ct_col0 + m]
  double alpha, beta;
  int tiles_m, tiles_n;    // in units of the CTA tile
  int skip_upper;          // skip CTA tiles that lie entirely above the diagonal of C (symmetric updates: lower part only)
};

template <int MI, int NJ>
__host__ __device__ constexpr int rank128_smem_bytes() { return 16 * (MI + NJ) * R128_LP * 8; }

template <int MI, int NJ>
__global__ void __launch_bounds__(256) rank128_kernel(const R128Params p) {
  constexpr int BM = 16 * MI, BN = 16 * NJ, LP = R128_LP;
  extern __shared__ __align__(16) double r128_sm[];
  double* As = r128_sm;
  double* Bs = r128_sm + BM * LP;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15, z = blockIdx.y;
  const int tile_m = blockIdx.x % p.tiles_m, tile_n = blockIdx.x / p.tiles_m;
  const int m0 = tile_m * BM, n0 = tile_n * BN;
  if (p.skip_upper && p.c_col0 + n0 > p.c_row0 + m0 + BM - 1) return;
  const double* Ag = p.A + p.a_z * z + (long long)(p.a_row0 + m0) * p.lda + p.a_col0;
  const double* Bg = p.B + p.b_z * z + (long long)(p.b_row0 + n0) * p.ldb + p.b_col0;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    for (int e = tid; e < BM * 32; e += 256) {
      const int r = e >> 5, c = ((e & 31) << 1) + 64 * h;
      cp_async16(As + r * LP + c, Ag + (long long)r * p.lda + c);
    }
    for (int e = tid; e < BN * 32; e += 256) {
      const int r = e >> 5, c = ((e & 31) << 1) + 64 * h;
      cp_async16(Bs + r * LP + c, Bg + (long long)r * p.ldb + c);
    }
    cp_async_commit();
  }
  double acc[MI][NJ];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[i][j] = 0.0;
  const double* a_base = As + ty * LP;
  const double* b_base = Bs + tx * LP;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    if (h == 0) cp_async_wait<1>();
    else cp_async_wait<0>();
    __syncthreads();
#pragma unroll 4
    for (int k = 64 * h; k < 64 * h + 64; k += 2) {
      double2 a[MI], b[NJ];
#pragma unroll
      for (int i = 0; i < MI; ++i) a[i] = *reinterpret_cast<const double2*>(a_base + 16 * i * LP + k);
#pragma unroll
      for (int j = 0; j < NJ; ++j) b[j] = *reinterpret_cast<const double2*>(b_base + 16 * j * LP + k);
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j] = fma(a[i].x, b[j].x, acc[i][j]);
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j] = fma(a[i].y, b[j].y, acc[i][j]);
    }
  }
  // epilogue: rows of C are written in 128-byte runs per half-warp straight from the registers; the transposed copy goes
  // through shared memory (over the A tile, which nobody reads any more after the barrier) so that its rows are runs too
  if (p.C) {
    double* Cg = p.C + p.c_z * z + (long long)(p.c_row0 + m0 + ty) * p.ldc + p.c_col0 + n0 + tx;
    if (p.beta != 0.0) {
      double cin[MI][NJ];
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) cin[i][j] = Cg[(long long)16 * i * p.ldc + 16 * j];
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j] = p.alpha * acc[i][j] + p.beta * cin[i][j];
    } else {
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j] = p.alpha * acc[i][j];
    }
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
      for (int j = 0; j < NJ; ++j) Cg[(long long)16 * i * p.ldc + 16 * j] = acc[i][j];
  } else {
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
      for (int j = 0; j < NJ; ++j) acc[i][j] = p.alpha * acc[i][j];
  }
  if (p.Ct) {
    constexpr int CP = BM + 1;
    double* Cs = r128_sm;   // [BN][BM + 1]
    __syncthreads();
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
      for (int j = 0; j < NJ; ++j) Cs[(tx + 16 * j) * CP + ty + 16 * i] = acc[i][j];
    __syncthreads();
    double* Ctg = p.Ct + p.ct_z * z + (long long)(p.ct_row0 + n0) * p.ldct + p.ct_col0 + m0;
    for (int e = tid; e < BN * BM; e += 256) {
      const int n = e / BM, m = e % BM;
      Ctg[(long long)n * p.ldct + m] = Cs[n * CP + m];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Panel of the blocked Cholesky by block substitution: rows below diagonal block kb,  P = A[kb+1.., kb] L_kk^-T.
// L_kk (128 x 128 lower) is split into 32-wide blocks; with D_p = (L_kk)_pp^-1 (the diagonal-block kernel leaves the
// four of them in the diagonal 32 x 32 blocks of X_kk):
//     P_p = (A_p - sum_{q < p} P_q L_pq^T) D_p^T,      p = 0 .. 3
// so the panel does not wait for the 128 x 128 inverse of the diagonal block (which moves off the critical chain).
// One CTA per strip of 32 rows (grid.x = 4 (nb - kb - 1), grid.y = S), 256 threads; thread (ty, tx) owns rows ty + 16 i,
// block columns tx + 16 j (2 x 2), two FMA chains per output (even / odd k), fixed order: bit-reproducible.
// ---------------------------------------------------------------------------------------------------------------
constexpr int TRSM_RS = 32;                 // rows per CTA
constexpr int TRSM_DP = 34;                 // pitch of the 32 x 32 blocks (D_p, T): 272 B = 2 x 128 + 16, conflict-free like R128_LP
constexpr int TRSM_SMEM_BYTES = (TILE * R128_LP + TRSM_RS * R128_LP + 4 * 32 * TRSM_DP + TRSM_RS * TRSM_DP) * 8;

__global__ void __launch_bounds__(256) trsm_panel_kernel(double* A, const double* X, int n_pad, int kb) {
  constexpr int LP = R128_LP, DP = TRSM_DP;
  extern __shared__ __align__(16) double trsm_sm[];
  double* Ls = trsm_sm;                       // [128][LP]  L_kk
  double* As = Ls + TILE * LP;                // [32][LP]   the strip: A on entry, P on exit
  double* Ds = As + TRSM_RS * LP;             // [4][32][DP]
  double* Ts = Ds + 4 * 32 * DP;              // [32][DP]
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15, z = blockIdx.y;
  const long long zoff = (long long)z * n_pad * n_pad;
  const double* Lg = A + zoff + (long long)kb * TILE * n_pad + (long long)kb * TILE;
  const double* Xg = X + zoff + (long long)kb * TILE * n_pad + (long long)kb * TILE;
  double* Ag = A + zoff + ((long long)(kb + 1) * TILE + (long long)blockIdx.x * TRSM_RS) * n_pad + (long long)kb * TILE;
  for (int e = tid; e < TRSM_RS * 64; e += 256) {
    const int r = e >> 6, c = (e & 63) << 1;
    cp_async16(As + r * LP + c, Ag + (long long)r * n_pad + c);
  }
  for (int e = tid; e < 4 * 32 * 16; e += 256) {
    const int pr = e >> 4, c = (e & 15) << 1;    // pr = 32 p + row
    cp_async16(Ds + pr * DP + c, Xg + (long long)pr * n_pad + (pr & ~31) + c);
  }
  for (int e = tid; e < (TILE - 32) * 64; e += 256) {   // block row 0 of L_kk is not needed (only D_0 is)
    const int r = 32 + (e >> 6), c = (e & 63) << 1;
    cp_async16(Ls + r * LP + c, Lg + (long long)r * n_pad + c);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
#pragma unroll 1
  for (int p = 0; p < 4; ++p) {
    const int c0 = 32 * p;
    double t[2][2], u[2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        t[i][j] = As[(ty + 16 * i) * LP + c0 + tx + 16 * j];
        u[i][j] = 0.0;
      }
#pragma unroll 4
    for (int q = 0; q < c0; q += 2) {
      double2 a[2], l[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) a[i] = *reinterpret_cast<const double2*>(As + (ty + 16 * i) * LP + q);
#pragma unroll
      for (int j = 0; j < 2; ++j) l[j] = *reinterpret_cast<const double2*>(Ls + (c0 + tx + 16 * j) * LP + q);
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          t[i][j] = fma(-a[i].x, l[j].x, t[i][j]);
          u[i][j] = fma(-a[i].y, l[j].y, u[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) Ts[(ty + 16 * i) * DP + tx + 16 * j] = t[i][j] + u[i][j];
    __syncthreads();
    double o[2][2], o2[2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) o[i][j] = o2[i][j] = 0.0;
#pragma unroll 4
    for (int q = 0; q < 32; q += 2) {
      double2 tt[2], dd[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) tt[i] = *reinterpret_cast<const double2*>(Ts + (ty + 16 * i) * DP + q);
#pragma unroll
      for (int j = 0; j < 2; ++j) dd[j] = *reinterpret_cast<const double2*>(Ds + (c0 + tx + 16 * j) * DP + q);
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          o[i][j] = fma(tt[i].x, dd[j].x, o[i][j]);
          o2[i][j] = fma(tt[i].y, dd[j].y, o2[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) As[(ty + 16 * i) * LP + c0 + tx + 16 * j] = o[i][j] + o2[i][j];
    __syncthreads();
  }
  for (int e = tid; e < TRSM_RS * 64; e += 256) {
    const int r = e >> 6, c = (e & 63) << 1;
    *reinterpret_cast<double2*>(Ag + (long long)r * n_pad + c) = *reinterpret_cast<const double2*>(As + r * LP + c);
  }
}

}  // namespace gpo
