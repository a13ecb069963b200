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
  double* C2; long long c2_z;   // may be null: dense copy [128][128] of output rows < 128, columns < 128 (the tile the next chain
                                // kernel factors against: the panel kernel overwrites it in C while that kernel still reads it)
  int kchunks;                  // rank128_dmma_kernel only: K in units of 32 (0 = 4, i.e. K = 128)
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
    if (p.C2 && m0 < TILE && n0 < TILE) {
      double* C2g = p.C2 + p.c2_z * z + (long long)(m0 + ty) * TILE + n0 + tx;
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) C2g[16 * i * TILE + 16 * j] = acc[i][j];
    }
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
// The same product on the FP64 tensor pipe: 64 x 64 tile per CTA of four warps (warp tile 32 x 32 = 4 x 4 DMMA m8n8k4
// accumulators), operands staged in 32-wide k chunks through a two-stage cp.async ring (80 KB: two CTAs per SM, so one CTA's
// loads / epilogue run under the other's DMMAs).  The register-tiled FMA kernel above is bound by shared-memory bandwidth
// (an LDS.128 per four FMAs, measured 2.2 x off the FP64 pipe); a DMMA consumes 256 FMAs per two fragment loads.
// Fragment loads are 16-byte: lane (g, t) reads k = 8 s + 2 t, 2 t + 1 of row g and feeds .x / .y to two DMMAs (the same k
// permutation for both operands); the row pitch of 320 B = 2 x 128 + 64 makes them conflict-free.  Not persistent: a launch
// is many short CTAs, so kernels of higher-priority streams get SMs within a few microseconds (no CTA caps needed).
// skip_upper also suppresses Ct on diagonal tiles (a symmetric update writes both triangles of those through C).
// Accumulation order is fixed: bit-reproducible.
// ---------------------------------------------------------------------------------------------------------------
constexpr int RD_KC = 32;                        // k extent of a stage
constexpr int RD_LP = RD_KC + 8;                 // row pitch (doubles)
constexpr int RD_STAGE = 128 * RD_LP;            // 64 A rows + 64 B rows
constexpr int RD_SMEM_BYTES = 2 * RD_STAGE * 8;  // 81,920

__global__ void __launch_bounds__(128, 2) rank128_dmma_kernel(const R128Params p) {
  extern __shared__ __align__(16) double rd_sm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3, z = blockIdx.y;
  const int tile_m = blockIdx.x % p.tiles_m, tile_n = blockIdx.x / p.tiles_m;
  const int m0 = tile_m * 64, n0 = tile_n * 64;
  if (p.skip_upper && p.c_col0 + n0 > p.c_row0 + m0 + 63) return;
  const int wm = (warp & 1) * 32, wn = (warp >> 1) * 32;
  const double* Ag = p.A + p.a_z * z + (long long)(p.a_row0 + m0) * p.lda + p.a_col0;
  const double* Bg = p.B + p.b_z * z + (long long)(p.b_row0 + n0) * p.ldb + p.b_col0;
  auto load_chunk = [&](int c, int buf) {
    double* st = rd_sm + buf * RD_STAGE;
    for (int e = tid; e < 128 * 16; e += 128) {
      const int row = e >> 4, ch = (e & 15) << 1;
      const double* src = row < 64 ? Ag + (long long)row * p.lda + RD_KC * c + ch : Bg + (long long)(row - 64) * p.ldb + RD_KC * c + ch;
      cp_async16(st + row * RD_LP + ch, src);
    }
    cp_async_commit();
  };
  const int nch = p.kchunks > 0 ? p.kchunks : TILE / RD_KC;   // (>= 2)
  load_chunk(0, 0);
  load_chunk(1, 1);
  if (p.C && p.beta != 0.0) {
    // the read-modify-write epilogue fetches its 64 x 64 tile of C at the very end, from HBM when the matrix is larger than the L2
    // (W^-1 at n = 4096: 134 MB): pull the 256 lines into the L2 now, a whole main loop ahead
    const double* Cp = p.C + p.c_z * z + (long long)(p.c_row0 + m0 + (tid >> 1)) * p.ldc + p.c_col0 + n0 + (tid & 1) * 32;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(Cp));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(Cp + 16));
  }
  double acc[4][4][2];
#pragma unroll
  for (int f = 0; f < 4; ++f)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[f][j][0] = acc[f][j][1] = 0.0;
#pragma unroll 1
  for (int c = 0; c < nch; ++c) {
    if (c < nch - 1) cp_async_wait<1>();
    else cp_async_wait<0>();
    __syncthreads();
    const double* As = rd_sm + (c & 1) * RD_STAGE + (wm + g) * RD_LP + 2 * t;
    const double* Bs = rd_sm + (c & 1) * RD_STAGE + (64 + wn + g) * RD_LP + 2 * t;
#pragma unroll
    for (int s = 0; s < RD_KC / 8; ++s) {
      double2 a[4], b[4];
#pragma unroll
      for (int f = 0; f < 4; ++f) a[f] = *reinterpret_cast<const double2*>(As + 8 * f * RD_LP + 8 * s);
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const double2*>(Bs + 8 * j * RD_LP + 8 * s);
#pragma unroll
      for (int f = 0; f < 4; ++f)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma_884(acc[f][j][0], acc[f][j][1], a[f].x, b[j].x);
#pragma unroll
      for (int f = 0; f < 4; ++f)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma_884(acc[f][j][0], acc[f][j][1], a[f].y, b[j].y);
    }
    __syncthreads();
    if (c + 2 < nch) load_chunk(c + 2, c & 1);
  }
  // accumulator (f, j, e) of lane (g, t): row wm + 8 f + g, column wn + 8 j + 2 t + e of the tile
  if (p.C) {
    double* Cg = p.C + p.c_z * z + (long long)(p.c_row0 + m0 + wm + g) * p.ldc + p.c_col0 + n0 + wn + 2 * t;
    if (p.beta != 0.0) {
      double2 cin[4][4];
#pragma unroll
      for (int f = 0; f < 4; ++f)
#pragma unroll
        for (int j = 0; j < 4; ++j) cin[f][j] = *reinterpret_cast<const double2*>(Cg + (long long)8 * f * p.ldc + 8 * j);
#pragma unroll
      for (int f = 0; f < 4; ++f)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[f][j][0] = p.alpha * acc[f][j][0] + p.beta * cin[f][j].x;
          acc[f][j][1] = p.alpha * acc[f][j][1] + p.beta * cin[f][j].y;
        }
    } else {
#pragma unroll
      for (int f = 0; f < 4; ++f)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[f][j][0] *= p.alpha; acc[f][j][1] *= p.alpha; }
    }
#pragma unroll
    for (int f = 0; f < 4; ++f)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<double2*>(Cg + (long long)8 * f * p.ldc + 8 * j) = make_double2(acc[f][j][0], acc[f][j][1]);
    if (p.C2 && m0 < TILE && n0 < TILE) {
      double* C2g = p.C2 + p.c2_z * z + (long long)(m0 + wm + g) * TILE + n0 + wn + 2 * t;
#pragma unroll
      for (int f = 0; f < 4; ++f)
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<double2*>(C2g + 8 * f * TILE + 8 * j) = make_double2(acc[f][j][0], acc[f][j][1]);
    }
  } else {
#pragma unroll
    for (int f = 0; f < 4; ++f)
#pragma unroll
      for (int j = 0; j < 4; ++j) { acc[f][j][0] *= p.alpha; acc[f][j][1] *= p.alpha; }
  }
  if (p.Ct && !(p.skip_upper && p.c_col0 + n0 + 63 >= p.c_row0 + m0)) {   // (the main loop ended with a barrier: the ring is free)
    constexpr int CP = 65;
    double* Cs = rd_sm;   // [64][65]
#pragma unroll
    for (int f = 0; f < 4; ++f)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        Cs[(wn + 8 * j + 2 * t) * CP + wm + 8 * f + g] = acc[f][j][0];
        Cs[(wn + 8 * j + 2 * t + 1) * CP + wm + 8 * f + g] = acc[f][j][1];
      }
    __syncthreads();
    double* Ctg = p.Ct + p.ct_z * z + (long long)(p.ct_row0 + n0) * p.ldct + p.ct_col0 + m0;
    for (int e = tid; e < 64 * 64; e += 128) {
      const int n = e >> 6, m = e & 63;
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

// The substitution itself, for a strip of 16 RI rows held in As (A on entry, P on exit); thread (ty, tx) owns rows ty + 16 i,
// block columns tx + 16 j.  Eight FMA chains per output (k mod 8; an FP64 operation has ~45 cycles of latency on this part and
// the kernel has two warps per scheduler to hide it), combined pairwise in a fixed order: the arithmetic of an output does not
// depend on RI, so the panel kernel (RI = 2) and the chain kernel (RI = 1) produce identical bits.
// Ls holds rows ls_row0 .. 127 of L_kk (rows < 32 are never touched).  Ends with a barrier.
template <int RI>
__device__ __forceinline__ void trsm_strip(double* As, const double* Ls, int ls_row0, const double* Ds, double* Ts, int ty, int tx) {
  constexpr int LP = R128_LP, DP = TRSM_DP;
#pragma unroll 1
  for (int p = 0; p < 4; ++p) {
    const int c0 = 32 * p;
    double acc[RI][2][8];
#pragma unroll
    for (int i = 0; i < RI; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        acc[i][j][0] = As[(ty + 16 * i) * LP + c0 + tx + 16 * j];
#pragma unroll
        for (int c = 1; c < 8; ++c) acc[i][j][c] = 0.0;
      }
#pragma unroll 1
    for (int q = 0; q < c0; q += 8) {
      double2 a[RI][4], l[2][4];
#pragma unroll
      for (int i = 0; i < RI; ++i)
#pragma unroll
        for (int h = 0; h < 4; ++h) a[i][h] = *reinterpret_cast<const double2*>(As + (ty + 16 * i) * LP + q + 2 * h);
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int h = 0; h < 4; ++h) l[j][h] = *reinterpret_cast<const double2*>(Ls + (c0 + tx + 16 * j - ls_row0) * LP + q + 2 * h);
#pragma unroll
      for (int i = 0; i < RI; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            acc[i][j][2 * h] = fma(-a[i][h].x, l[j][h].x, acc[i][j][2 * h]);
            acc[i][j][2 * h + 1] = fma(-a[i][h].y, l[j][h].y, acc[i][j][2 * h + 1]);
          }
    }
#pragma unroll
    for (int i = 0; i < RI; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j)
        Ts[(ty + 16 * i) * DP + tx + 16 * j] = ((acc[i][j][0] + acc[i][j][4]) + (acc[i][j][2] + acc[i][j][6])) +
                                               ((acc[i][j][1] + acc[i][j][5]) + (acc[i][j][3] + acc[i][j][7]));
    __syncthreads();
#pragma unroll
    for (int i = 0; i < RI; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[i][j][c] = 0.0;
#pragma unroll
    for (int q = 0; q < 32; q += 8) {
      double2 t[RI][4], d[2][4];
#pragma unroll
      for (int i = 0; i < RI; ++i)
#pragma unroll
        for (int h = 0; h < 4; ++h) t[i][h] = *reinterpret_cast<const double2*>(Ts + (ty + 16 * i) * DP + q + 2 * h);
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int h = 0; h < 4; ++h) d[j][h] = *reinterpret_cast<const double2*>(Ds + (c0 + tx + 16 * j) * DP + q + 2 * h);
#pragma unroll
      for (int i = 0; i < RI; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            acc[i][j][2 * h] = fma(t[i][h].x, d[j][h].x, acc[i][j][2 * h]);
            acc[i][j][2 * h + 1] = fma(t[i][h].y, d[j][h].y, acc[i][j][2 * h + 1]);
          }
    }
#pragma unroll
    for (int i = 0; i < RI; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j)
        As[(ty + 16 * i) * LP + c0 + tx + 16 * j] = ((acc[i][j][0] + acc[i][j][4]) + (acc[i][j][2] + acc[i][j][6])) +
                                                    ((acc[i][j][1] + acc[i][j][5]) + (acc[i][j][3] + acc[i][j][7]));
    __syncthreads();
  }
}

// stage L_kk (rows 32 .. 127, to Ls + (row - ls_row0) LP) and the four diagonal 32 x 32 inverses of block kb into shared memory
// (cp.async, no wait)
__device__ __forceinline__ void trsm_stage_factor(double* Ls, int ls_row0, double* Ds, const double* Lg, const double* Xg, int n_pad,
                                                  int tid) {
  constexpr int LP = R128_LP, DP = TRSM_DP;
  for (int e = tid; e < 4 * 32 * 16; e += 256) {
    const int pr = e >> 4, c = (e & 15) << 1;    // pr = 32 p + row
    cp_async16(Ds + pr * DP + c, Xg + (long long)pr * n_pad + (pr & ~31) + c);
  }
  for (int e = tid; e < (TILE - 32) * 64; e += 256) {   // block row 0 of L_kk is not needed (only D_0 is)
    const int r = 32 + (e >> 6), c = (e & 63) << 1;
    cp_async16(Ls + (r - ls_row0) * LP + c, Lg + (long long)r * n_pad + c);
  }
}

// first_block: index of the first 128-row block of the panel to solve (kb + 1 for the whole panel)
__global__ void __launch_bounds__(256) trsm_panel_kernel(double* A, const double* X, int n_pad, int kb, int first_block) {
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
  double* Ag = A + zoff + ((long long)first_block * TILE + (long long)blockIdx.x * TRSM_RS) * n_pad + (long long)kb * TILE;
  for (int e = tid; e < TRSM_RS * 64; e += 256) {
    const int r = e >> 6, c = (e & 63) << 1;
    cp_async16(As + r * LP + c, Ag + (long long)r * n_pad + c);
  }
  trsm_stage_factor(Ls, 0, Ds, Lg, Xg, n_pad, tid);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  trsm_strip<2>(As, Ls, 0, Ds, Ts, ty, tx);
  for (int e = tid; e < TRSM_RS * 64; e += 256) {
    const int r = e >> 6, c = (e & 63) << 1;
    *reinterpret_cast<double2*>(Ag + (long long)r * n_pad + c) = *reinterpret_cast<const double2*>(As + r * LP + c);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// The critical chain of block column kb >= 1 in ONE launch of eight CTAs per hyper-parameter set (any eight SMs; the CTAs meet
// through two arrival counters in global memory):
//   A  CTA r solves rows 16 r .. 16 r + 15 of the first panel block, P = Raw L^-T with L = L_(kb-1,kb-1)  (trsm_strip, the same
//      bits the panel kernel writes into the matrix), and publishes them in Pg
//   B  it then forms its rows of the updated diagonal block, T = A_kk - P P^T (lower part), and publishes them in Tg
//   C  CTA 0 factors T (diag_chol_inv_body, mode 1: L_kk, the diagonal 32 x 32 inverses, the reciprocal pivots; mode 0 for the
//      last block column: the 128 x 128 inverse as well)
// `Raw` is the tile (kb, kb-1) as the previous updates left it (dense [128][128] copy made by whoever wrote it last: the panel
// kernel overwrites the matrix's own tile concurrently).  Nothing else of the matrix is written.  A CTA that is scheduled late
// only delays the others (they spin); every CTA of the launch is eventually resident, so the waits terminate.
// ---------------------------------------------------------------------------------------------------------------
constexpr int CHAIN_CTAS = 8;
constexpr int CHAIN_ROWS = TILE / CHAIN_CTAS;     // 16
constexpr int CHAIN_OFF_DS = TILE * R128_LP;                       // after Pl [128][LP] (phase A: Ls rows 32.., Ts)
constexpr int CHAIN_OFF_AS_MIN = CHAIN_OFF_DS + TILE * TRSM_DP;
constexpr int CHAIN_OFF_AS = CHAIN_OFF_AS_MIN > ((DIAG_SMEM_BYTES / 8 + 1) & ~1) ? CHAIN_OFF_AS_MIN : ((DIAG_SMEM_BYTES / 8 + 1) & ~1);  // own strip [16][LP]
constexpr int CHAIN_SMEM_BYTES = (CHAIN_OFF_AS + CHAIN_ROWS * R128_LP) * 8;
static_assert(DIAG_SMEM_BYTES <= CHAIN_OFF_AS * 8, "the factorisation's workspace must not reach the strip");

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Phase B of the chain kernel for one thread.  CTA r owns the row groups G1 = 8 r .. 8 r + 7 and G2 = 120 - 8 r .. 127 - 8 r of
// the 128 x 128 block (a light and a heavy one: every CTA forms 9 x 128 outputs of the lower triangle).  Thread (rho, h, tx)
// takes row rho of both groups and the column groups {tx + 16 c : c = h, h + 2, ...}: c < NJ2 = 9 - NJ1 for the G2 row, c < NJ1
// for the G1 row (NJ1 = r / 2 + 1), so the P rows it reads as the second operand serve both of its rows.  Four FMA chains per
// output (k mod 4), fixed order.  P must be complete in Pl (barrier before the call); ends with the stores to T (no barrier).
template <int NJ1, int H>
__device__ __forceinline__ void chain_syrk_rows(const double* Pl, const double* Akk, int n_pad, double* Tz, int r, int tid) {
  constexpr int LP = R128_LP, NJ2 = 9 - NJ1;
  constexpr int N2 = (NJ2 - H + 1) / 2, N1 = (NJ1 - H + 1) / 2;   // column groups of this thread for the G2 / G1 row
  const int rho = (tid >> 4) & 7, tx = tid & 15;
  const int row1 = 8 * r + rho, row2 = 120 - 8 * r + rho;
  double tin2[N2 > 0 ? N2 : 1], tin1[N1 > 0 ? N1 : 1];
#pragma unroll
  for (int u = 0; u < N2; ++u) tin2[u] = Akk[(long long)row2 * n_pad + tx + 16 * (2 * u + H)];
#pragma unroll
  for (int u = 0; u < N1; ++u) tin1[u] = Akk[(long long)row1 * n_pad + tx + 16 * (2 * u + H)];
  cp_async_wait<0>();
  __syncthreads();
  double acc2[N2 > 0 ? N2 : 1][4], acc1[N1 > 0 ? N1 : 1][4];
#pragma unroll
  for (int u = 0; u < N2; ++u) acc2[u][0] = acc2[u][1] = acc2[u][2] = acc2[u][3] = 0.0;
#pragma unroll
  for (int u = 0; u < N1; ++u) acc1[u][0] = acc1[u][1] = acc1[u][2] = acc1[u][3] = 0.0;
  const double* a2p = Pl + row2 * LP;
  const double* a1p = Pl + row1 * LP;
  const double* bp = Pl + (tx + 16 * H) * LP;
#pragma unroll 2
  for (int k = 0; k < TILE; k += 4) {
    const double2 a20 = *reinterpret_cast<const double2*>(a2p + k), a21 = *reinterpret_cast<const double2*>(a2p + k + 2);
    double2 a10, a11;
    if (N1 > 0) { a10 = *reinterpret_cast<const double2*>(a1p + k); a11 = *reinterpret_cast<const double2*>(a1p + k + 2); }
#pragma unroll
    for (int u = 0; u < N2; ++u) {
      const double2 b0 = *reinterpret_cast<const double2*>(bp + 32 * u * LP + k), b1 = *reinterpret_cast<const double2*>(bp + 32 * u * LP + k + 2);
      acc2[u][0] = fma(a20.x, b0.x, acc2[u][0]);
      acc2[u][1] = fma(a20.y, b0.y, acc2[u][1]);
      acc2[u][2] = fma(a21.x, b1.x, acc2[u][2]);
      acc2[u][3] = fma(a21.y, b1.y, acc2[u][3]);
      if (u < N1) {
        acc1[u][0] = fma(a10.x, b0.x, acc1[u][0]);
        acc1[u][1] = fma(a10.y, b0.y, acc1[u][1]);
        acc1[u][2] = fma(a11.x, b1.x, acc1[u][2]);
        acc1[u][3] = fma(a11.y, b1.y, acc1[u][3]);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < N2; ++u)
    Tz[(long long)row2 * TILE + tx + 16 * (2 * u + H)] = tin2[u] - ((acc2[u][0] + acc2[u][2]) + (acc2[u][1] + acc2[u][3]));
#pragma unroll
  for (int u = 0; u < N1; ++u)
    Tz[(long long)row1 * TILE + tx + 16 * (2 * u + H)] = tin1[u] - ((acc1[u][0] + acc1[u][2]) + (acc1[u][1] + acc1[u][3]));
}

#ifdef GPO_DIAG_TIMING
#define CHAIN_T(i) do { if (tid == 0 && z == 0 && (r == 0 || r == CHAIN_CTAS - 1)) g_diag_t[(r == 0 ? 44 : 54) + (i)] = clock64(); } while (0)
#else
#define CHAIN_T(i) do { } while (0)
#endif
__global__ void __launch_bounds__(DIAG_THREADS) chol_chain_kernel(double* A, double* X, double* U, int n_pad, int kb, int n_valid,
                                                                  int* info, double* rdiag, const double* Raw, double* Pg,
                                                                  double* Tg, int* flags, int mode) {
  constexpr int LP = R128_LP;
  extern __shared__ __align__(16) double chain_sm[];
  double* sm = chain_sm;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int z = blockIdx.x / CHAIN_CTAS, r = blockIdx.x % CHAIN_CTAS;
  const long long zoff = (long long)z * n_pad * n_pad;
  const long long tt = (long long)TILE * TILE;
  const double* Rawz = Raw + tt * z;
  double* Pz = Pg + tt * z;
  double* Tz = Tg + tt * z;
  int* flag = flags + 2 * z;
  double* Pl = sm;                                 // phase B: rows 0 .. 16 r + 15 of P
  double* Ls = sm;                                 // phase A: rows 32 .. 127 of L
  double* Ts = sm + (TILE - 32) * LP;              // phase A: [16][DP], behind the factor
  double* Ds = sm + CHAIN_OFF_DS;
  double* As = sm + CHAIN_OFF_AS;
  CHAIN_T(0);
  // ---- A: the strip of the first panel block ----
  {
    const double* Lg = A + zoff + (long long)(kb - 1) * TILE * n_pad + (long long)(kb - 1) * TILE;
    const double* Xg = X + zoff + (long long)(kb - 1) * TILE * n_pad + (long long)(kb - 1) * TILE;
    const double* Rg = Rawz + (long long)r * CHAIN_ROWS * TILE;
    for (int e = tid; e < CHAIN_ROWS * 64; e += 256) {
      const int rr = e >> 6, c = (e & 63) << 1;
      cp_async16(As + rr * LP + c, Rg + rr * TILE + c);
    }
    trsm_stage_factor(Ls, 32, Ds, Lg, Xg, n_pad, tid);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    CHAIN_T(1);
    trsm_strip<1>(As, Ls, 32, Ds, Ts, ty, tx);
    CHAIN_T(2);
    for (int e = tid; e < CHAIN_ROWS * 64; e += 256) {
      const int rr = e >> 6, c = (e & 63) << 1;
      *reinterpret_cast<double2*>(Pz + (long long)(r * CHAIN_ROWS + rr) * TILE + c) = *reinterpret_cast<const double2*>(As + rr * LP + c);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      red_release_gpu_add(flag, 1);
      while (ld_acquire_gpu(flag) < CHAIN_CTAS) { }
      __threadfence();
    }
    __syncthreads();
    CHAIN_T(3);
  }
  // ---- B: this CTA's rows of  T = A_kk - P P^T  (lower part) ----
  {
    for (int e = tid; e < TILE * 64; e += 256) {   // all of P as the second operand (L2 loads: other CTAs wrote it)
      const int rr = e >> 6, c = (e & 63) << 1;
      cp_async16(Pl + rr * LP + c, Pz + (long long)rr * TILE + c);
    }
    cp_async_commit();
    const double* Akk = A + zoff + (long long)kb * TILE * n_pad + (long long)kb * TILE;
    const int h = tid >> 7;
    cp_async_wait<0>();
    __syncthreads();
    CHAIN_T(4);
    switch (r >> 1) {
      case 0: if (h == 0) chain_syrk_rows<1, 0>(Pl, Akk, n_pad, Tz, r, tid); else chain_syrk_rows<1, 1>(Pl, Akk, n_pad, Tz, r, tid); break;
      case 1: if (h == 0) chain_syrk_rows<2, 0>(Pl, Akk, n_pad, Tz, r, tid); else chain_syrk_rows<2, 1>(Pl, Akk, n_pad, Tz, r, tid); break;
      case 2: if (h == 0) chain_syrk_rows<3, 0>(Pl, Akk, n_pad, Tz, r, tid); else chain_syrk_rows<3, 1>(Pl, Akk, n_pad, Tz, r, tid); break;
      default: if (h == 0) chain_syrk_rows<4, 0>(Pl, Akk, n_pad, Tz, r, tid); else chain_syrk_rows<4, 1>(Pl, Akk, n_pad, Tz, r, tid); break;
    }
    CHAIN_T(5);
    __threadfence();
    __syncthreads();
    if (tid == 0) red_release_gpu_add(flag + 1, 1);
    CHAIN_T(6);
    if (r != 0) return;
    if (tid == 0) {
      while (ld_acquire_gpu(flag + 1) < CHAIN_CTAS) { }
      __threadfence();
    }
    __syncthreads();
  }
  CHAIN_T(7);
  // ---- C: CTA 0 factors the updated block ----
  diag_chol_inv_body<true>(sm, Tz, TILE, A, X, U, n_pad, kb, n_valid, info, mode, rdiag, z);
}

// dense [128][128] copy of tile (row_block, col_block) of every batch: the first chain kernel's `Raw`
__global__ void tile_copy_kernel(const double* A, double* out, int n_pad, int row_block, int col_block) {
  const int z = blockIdx.y;
  const double* src = A + (long long)z * n_pad * n_pad + (long long)row_block * TILE * n_pad + (long long)col_block * TILE;
  double* dst = out + (long long)z * TILE * TILE;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < TILE * TILE; e += gridDim.x * blockDim.x)
    dst[e] = src[(long long)(e >> 7) * n_pad + (e & (TILE - 1))];
}

}  // namespace gpo
