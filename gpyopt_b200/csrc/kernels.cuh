// Non-GEMM kernels of the path: cross-covariance + mean (K1), Gram build / diagonal-block Cholesky /
// triangular GEMV (K4 helpers), acquisition epilogue (K3) and top-k selection.
#pragma once
#include <cfloat>
#include "common.cuh"
#include "kernfun.cuh"

namespace gpo {

constexpr int DMAX = 16;  // largest supported input dimension

// =====================================================================================================
// K1: fused cross-covariance + posterior mean (+ mean gradient).
// Writes the K*^T rows ([cand][n_pad], the K-major B operand of the variance GEMM) and reduces k.alpha and
// sum_j alpha_j dk_j/dx in registers.
// Replaces GPy kern.K(X, X*) + Kx^T alpha (PosteriorExact._raw_predict) and
// kern.gradients_X(alpha^T, X*, X) (GP.predictive_gradients), reference call sites
// /root/reference/GPyOpt/models/gpmodel.py:98,136,138.
// =====================================================================================================
struct CovParams {
  const double* Xs;       // candidates AoS [n_cand][d]
  long long n_cand;       // valid rows
  long long nc_run;       // rows to process (multiple of 128, >= n_cand; the excess is zero-coordinate padding)
  long long nc_ld;        // leading dimension (rows) of the chunk buffers Kt / mean / dmean
  const double* Xt;       // training inputs SoA [d][ldx]
  long long ldx;          // = n_pad
  const double* xtile;    // [S][n_pad/32][DP+1][32]: per 32-row tile the length-scaled coordinates (rows 0..DP-1, zero past d)
                          // and alpha (row DP), contiguous so that one TMA bulk copy stages a tile
  int n, n_pad, d, kern;
  const double* alpha;    // [S][n_pad]
  const double* invl;     // [S][d]
  const double* sigf2;    // [S]
  double* Kt;             // [S][nc_ld][n_pad]   (may be null: mean only)
  double* Ht;             // [S][nc_ld][n_pad]   (dK/dr)/r, written only when non-null (non-RBF kernels, GRAD)
  int js, tiles_per_split; // the training index is cut into `js` contiguous slices (grid.y) of `tiles_per_split` 32-row tiles;
                           // js depends only on n_pad, so a row's result never depends on the batch it is in
  double* mean;           // [S][js][nc_ld]      per-slice partial sums, added in slice order by K3 / fit_scalars
  double* dmean;          // [S][js][d][nc_ld]   (GRAD only)
  int* ticket;            // cov_small_kernel only: [S] arrival counters (zero before the launch)
};

// One thread owns one candidate (its scaled coordinates, mean and mean-gradient accumulators stay in registers for the
// whole pass over the training set); the block walks the training index in tiles of 32 whose scaled coordinates and
// alpha are staged in shared memory by TMA bulk copies (cp.async.bulk + mbarrier, double buffered) and read as broadcasts.  The covariance tile [32 j][128 candidates] is transposed
// through (padded, conflict-free) shared memory so that K*^T rows are written as full 256-byte lines.
constexpr int COV_THREADS = 128;
constexpr int COV_JT = 32;
constexpr int COV_PITCH = COV_THREADS + 1;
constexpr int cov_smem_bytes(int dp, bool with_h) {
  return (COV_JT * COV_PITCH * (with_h ? 2 : 1) + 2 * (dp + 1) * COV_JT + 2) * 8;
}

template <int DP, bool GRAD>
__global__ void __launch_bounds__(COV_THREADS) cov_mean_kernel(const CovParams p) {
  extern __shared__ double cov_smem[];
  double* tile_k = cov_smem;                                   // [32][129]
  double* xj_s0 = cov_smem + COV_JT * COV_PITCH;               // 2 x [(DP+1)][32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(xj_s0 + 2 * (DP + 1) * COV_JT);  // 2 mbarriers
  double* tile_h = xj_s0 + 2 * (DP + 1) * COV_JT + 2;          // [32][129], only when Ht is written
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int z = blockIdx.z;
  const long long c0 = (long long)blockIdx.x * COV_THREADS;
  const long long cand = c0 + tid;
  const double* invl = p.invl + (long long)z * p.d;
  const double sf2 = p.sigf2[z];   // (alpha comes in through the staged tiles)
  double xs[DP], dm[DP], m = 0.0;
#pragma unroll
  for (int q = 0; q < DP; ++q) {
    xs[q] = (q < p.d && cand < p.n_cand) ? p.Xs[cand * p.d + q] * invl[q] : 0.0;
    dm[q] = 0.0;
  }
  double* Kt = p.Kt ? p.Kt + ((long long)z * p.nc_ld + c0) * p.n_pad : nullptr;
  double* Ht = (GRAD && p.Ht) ? p.Ht + ((long long)z * p.nc_ld + c0) * p.n_pad : nullptr;
  const int split = blockIdx.y;
  const int jt_begin = split * p.tiles_per_split, jt_end = min(p.n_pad / COV_JT, jt_begin + p.tiles_per_split);
  // stage loader: one elected thread issues a TMA bulk copy of the next tile [(DP+1)][32] and arms its mbarrier
  constexpr uint32_t kStageBytes = (DP + 1) * COV_JT * 8;
  const double* xtile = p.xtile + (long long)z * (p.n_pad / COV_JT) * (DP + 1) * COV_JT;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  auto stage = [&](int jt, int buf) {
    if (tid == 0) {
      mbar_arrive_expect_tx(&bars[buf], kStageBytes);
      tma_bulk_load(xj_s0 + buf * (DP + 1) * COV_JT, xtile + (long long)jt * (DP + 1) * COV_JT, kStageBytes, &bars[buf]);
    }
  };
  stage(jt_begin, 0);
  for (int jt = jt_begin; jt < jt_end; ++jt) {
    const int buf = (jt - jt_begin) & 1;
    // buffer buf^1 was last read in the previous iteration, which every thread left through a __syncthreads
    if (jt + 1 < jt_end) stage(jt + 1, buf ^ 1);
    mbar_wait(&bars[buf], (uint32_t)(((jt - jt_begin) >> 1) & 1));
    const double* xj = xj_s0 + buf * (DP + 1) * COV_JT;
    const int j0 = jt * COV_JT;
    // JV training points at a time: all JV covariance values are evaluated in one straight-line block (kernel switch
    // outside, no per-point branches), so their dependency chains -- exp is 17 dependent FMAs -- interleave
    constexpr int JV = 4;
    constexpr bool KEEP_DF = DP <= 8;    // the scaled differences stay in registers for the gradient (recomputed for larger d)
#pragma unroll 1
    for (int jb = 0; jb < COV_JT; jb += JV) {
      double r2[JV], kv[JV], hv[JV], df[KEEP_DF ? JV : 1][DP];
#pragma unroll
      for (int v = 0; v < JV; ++v) {
        r2[v] = 0.0;
#pragma unroll
        for (int q = 0; q < DP; ++q) {
          const double dq = xs[q] - xj[q * COV_JT + jb + v];
          if (KEEP_DF) df[KEEP_DF ? v : 0][q] = dq;
          r2[v] = fma(dq, dq, r2[v]);
        }
      }
      kern_k_h_vec<JV>(p.kern, sf2, r2, kv, hv);
#pragma unroll
      for (int v = 0; v < JV; ++v) {
        const int jj = jb + v;
        const bool live = j0 + jj < p.n;
        kv[v] = live ? kv[v] : 0.0;
        hv[v] = live ? hv[v] : 0.0;
        const double aj = xj[DP * COV_JT + jj];
        m = fma(kv[v], aj, m);
        if (GRAD) {
          const double w = aj * hv[v];
#pragma unroll
          for (int q = 0; q < DP; ++q) dm[q] = fma(w, KEEP_DF ? df[KEEP_DF ? v : 0][q] : xs[q] - xj[q * COV_JT + jj], dm[q]);
          if (Ht) tile_h[jj * COV_PITCH + tid] = hv[v];
        }
        if (Kt) tile_k[jj * COV_PITCH + tid] = kv[v];
      }
    }
    __syncthreads();
    if (Kt) {
      // transposed write-out: warp w stores candidates w, w+4, ...; lane l carries training index j0 + l
      for (int c = warp; c < COV_THREADS; c += COV_THREADS / 32) {
        if (c0 + c < p.nc_run) {
          Kt[(long long)c * p.n_pad + j0 + lane] = tile_k[lane * COV_PITCH + c];
          if (GRAD && Ht) Ht[(long long)c * p.n_pad + j0 + lane] = tile_h[lane * COV_PITCH + c];
        }
      }
    }
    __syncthreads();
  }
  if (cand < p.nc_run) {
    const long long zs = (long long)z * p.js + split;
    p.mean[zs * p.nc_ld + cand] = m;
    if (GRAD) {
#pragma unroll
      for (int q = 0; q < DP; ++q)
        if (q < p.d) p.dmean[(zs * p.d + q) * p.nc_ld + cand] = dm[q] * invl[q];
    }
  }
}

// Tile-major staging buffer for K1's TMA loads: xtile[z][jt][q][jj] = X[jt*32+jj][q] / l_zq (q < d), 0 (d <= q < dp),
// alpha_z[jt*32+jj] (q == dp).  Rebuilt after every refit (alpha and the lengthscales change).
__global__ void build_xtile_kernel(double* xtile, const double* Xt, long long ldx, const double* alpha, const double* invl,
                                   int n_pad, int d, int dp) {
  const int jt = blockIdx.x, z = blockIdx.y, ntiles = n_pad / COV_JT;
  double* out = xtile + ((long long)z * ntiles + jt) * (dp + 1) * COV_JT;
  for (int t = threadIdx.x; t < (dp + 1) * COV_JT; t += blockDim.x) {
    const int q = t / COV_JT, j = jt * COV_JT + t % COV_JT;
    double v = 0.0;
    if (q < d) v = Xt[q * ldx + j] * invl[z * d + q];
    else if (q == dp) v = alpha[(long long)z * n_pad + j];
    out[t] = v;
  }
}

// =====================================================================================================
// K4 helpers
// =====================================================================================================
// Gram matrix Ky = K(X,X) + (noise + 1e-8 + extra_jitter) I, identity on the padding.  GPy
// ExactGaussianInference.inference; extra_jitter implements GPy jitchol's retry ladder.
__global__ void gram_kernel(double* A, const double* Xt, long long ldx, int n, int n_pad, int d, int kern,
                            const double* invl, const double* sigf2, const double* noise, const double* extra_jitter) {
  const int z = blockIdx.z;
  const int i = blockIdx.y * blockDim.y + threadIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad || j >= n_pad) return;
  double v;
  if (i < n && j < n) {
    double r2 = 0.0;
    for (int q = 0; q < d; ++q) {
      const double il = invl[z * d + q];
      const double df = Xt[q * ldx + i] * il - Xt[q * ldx + j] * il;
      r2 = fma(df, df, r2);
    }
    v = kern_k(kern, sigf2[z], r2);
    if (i == j) v += noise[z] + 1e-8 + extra_jitter[z];
  } else {
    v = (i == j) ? 1.0 : 0.0;
  }
  A[((long long)z * n_pad + i) * n_pad + j] = v;
}

// Prior covariance between two candidate sets (AoS [N][d]): out[i][j] = k(x1_i, x2_j) (+ diag_add on i == j), zero on the
// padding.  First term of GP.predict(full_cov=True) / posterior_covariance_between_points.
__global__ void cross_gram_kernel(double* out, long long ldo, const double* X1, long long n1, const double* X2, long long n2,
                                  long long n1p, long long n2p, int d, int kern, const double* invl, const double* sigf2,
                                  double diag_add) {
  const long long i = (long long)blockIdx.y * blockDim.y + threadIdx.y, j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n1p || j >= n2p) return;
  double v = 0.0;
  if (i < n1 && j < n2) {
    double r2 = 0.0;
    for (int q = 0; q < d; ++q) {
      const double df = X1[i * d + q] * invl[q] - X2[j * d + q] * invl[q];
      r2 = fma(df, df, r2);
    }
    v = kern_k(kern, sigf2[0], r2);
    if (i == j) v += diag_add;
  }
  out[i * ldo + j] = v;
}

// mean of diag(K(X,X)) + noise + 1e-8 over the n real rows == sigf2 + noise + 1e-8 for a stationary kernel:
// computed on the host; no kernel needed.

// Cholesky factor + inverse of one 128 x 128 diagonal block per batch z.
//   in : lower triangle of A[z][kb*128 .. +128)^2   (upper part ignored)
//   out: A block <- L (upper zeroed), X block <- L^-1 (lower), U block <- L^-T (upper)
// info[z] is set to (global column index + 1) of the first non-positive pivot.
// 16 x 16 threads; thread (ty, tx) keeps the 8 x 8 elements (ty + 16 a, tx + 16 b) in REGISTERS through both phases
// (cyclic ownership keeps all threads busy while the active region shrinks).
constexpr int DIAG_THREADS = 256;
constexpr int DIAG_PITCH = TILE + 1;
constexpr int DIAG_PB = 32;                       // inner panel width
constexpr int DIAG_NP = TILE / DIAG_PB;           // 4 panels
constexpr int DIAG_XD_PITCH = DIAG_PB + 1;
constexpr int DIAG_SMEM_BYTES = (TILE * DIAG_PITCH + DIAG_NP * DIAG_PB * DIAG_XD_PITCH + TILE + DIAG_PB * DIAG_PB + 2) * 8;

#ifndef GPO_VAR_FENCE
#define GPO_VAR_FENCE 0
#endif
#ifndef GPO_VAR_RSQRT
#define GPO_VAR_RSQRT 1
#endif
#ifndef GPO_VAR_SLEEP
#define GPO_VAR_SLEEP 0
#endif
// 1 / sqrt(d) for a pivot in [2^-100, 2^99): single-precision MUFU seed (relative error ~2^-22) and one third-order step,
// y (1 + e/2 + 3 e^2/8) with e = 1 - d y^2: relative error ~e^3 ~ 1e-20 before the final rounding.  Four dependent FP64
// operations instead of the eight of rsqrt(double): the pivot chain of the diagonal-block kernel is a latency chain.
__device__ __forceinline__ double rsqrt_pivot(double d) {
  const double y = (double)rsqrtf((float)d);
  const double t = d * y;
  const double e = fma(-t, y, 1.0);
  const double p = fma(0.375, e, 0.5);
  const double ye = y * e;
  return fma(ye, p, y);
}

// Cholesky of the 32x32 diagonal block at S[j0.., j0..] by ONE warp, lane = row, the row in registers.  Per column:
// broadcast of the pivot, 1/sqrt in every lane, the scaled column goes to shared memory (cbuf[j][row], all 32 columns are
// kept) and comes back as 16-byte broadcast loads (one LDS.128 per two columns still to update; a shuffle version needs five
// instructions per column and is issue-bound).  The NEXT pivot does not wait for that round trip: its owner forms
// a_(j+1,j+1) - l^2 from its own l, so the serial chain per column is shuffle -> rsqrt -> multiply -> fma.
// After column j is in cbuf (and 1 / L_jj in rd) the warp publishes *prog = j + 1: the other warps solve the rows below the
// block and the block's inverse one column behind it (fwdsub32_trail) instead of after it.
// Writes L (lower part) back and 1/L_jj to rd.  cbuf: 32 x 32 doubles, 16-byte aligned.
__device__ __forceinline__ void chol32_warp(double* S, double* rd, double* cbuf, volatile int* prog, int j0, int lane, int* info_z,
                                            int col0) {
  double a[DIAG_PB];
#pragma unroll
  for (int k = 0; k < DIAG_PB; ++k) a[k] = (k <= lane) ? S[(j0 + lane) * DIAG_PITCH + j0 + k] : 0.0;
  double dd = __shfl_sync(0xffffffffu, a[0], 0);
#pragma unroll
  for (int j = 0; j < DIAG_PB; ++j) {
#if GPO_VAR_RSQRT
    double rp = rsqrt_pivot(dd);
#else
    double rp = rsqrt(dd);
#endif
    // exponent check on the high word (integer pipe, off the FP64 chain): false for <= 0, NaN, tiny and huge pivots
    if (!((unsigned)(__double2hiint(dd) - 0x39B00000) < (unsigned)(0x46300000 - 0x39B00000))) {  // uniform: dd is a broadcast
      if (!(dd > 0.0)) {
        if (lane == j && *info_z == 0) *info_z = col0 + j0 + j + 1;
        dd = 1.0;
      }
      rp = rsqrt(dd);
    }
    const double l = (lane > j) ? a[j] * rp : (lane == j ? dd * rp : 0.0);  // as LAPACK dpotf2: scale by the reciprocal pivot
    a[j] = l;
    if (lane == j) rd[j0 + j] = rp;
    if (j < DIAG_PB - 1) {
      dd = __shfl_sync(0xffffffffu, fma(-l, l, a[j + 1]), j + 1);   // next pivot, from its owner's own l
      double* cb = cbuf + j * DIAG_PB;
      cb[lane] = l;
      __syncwarp();
      if (lane == 0) {
#if GPO_VAR_FENCE
        __threadfence_block();
#endif
        *prog = j + 1;
      }
      if (((j + 1) & 1) != 0) a[j + 1] = fma(-l, cb[j + 1], a[j + 1]);   // odd first column: single load
#pragma unroll
      for (int k = (j + 2) & ~1; k < DIAG_PB; k += 2) {
        const double2 lk = *reinterpret_cast<const double2*>(cb + k);
        a[k] = fma(-l, lk.x, a[k]);   // lanes below k carry unused values in a[k]
        a[k + 1] = fma(-l, lk.y, a[k + 1]);
      }
    } else {
      __syncwarp();
      if (lane == 0) {
        __threadfence_block();
        *prog = DIAG_PB;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < DIAG_PB; ++k)
    if (k <= lane) S[(j0 + lane) * DIAG_PITCH + j0 + k] = a[k];
}

// fwdsub32 (below) against the block chol32_warp is still factoring: column c of L is read from cbuf as soon as *prog > c.
// Same operations on the same numbers as fwdsub32 after the fact.
__device__ __forceinline__ void fwdsub32_trail(double* vec, int sv, const double* cbuf, const double* rdp, const volatile int* prog) {
  double v[DIAG_PB];
#pragma unroll
  for (int c = 0; c < DIAG_PB; ++c) v[c] = vec[c * sv];
#pragma unroll
  for (int c = 0; c < DIAG_PB; ++c) {
    while (*prog <= c) {
#if GPO_VAR_SLEEP
      __nanosleep(GPO_VAR_SLEEP);
#endif
    }
    __threadfence_block();
    v[c] *= rdp[c];
#pragma unroll
    for (int c2 = c + 1; c2 < DIAG_PB; ++c2) v[c2] = fma(-v[c], cbuf[c * DIAG_PB + c2], v[c2]);
  }
#pragma unroll
  for (int c = 0; c < DIAG_PB; ++c) vec[c * sv] = v[c];
}

// Forward substitution L x = v with a 32x32 lower-triangular block (rdp = 1 / diagonal), v in shared memory with stride
// sv, one thread per right-hand side; the entries of L are shared-memory broadcasts.  The diagonal-block kernel funnels
// ALL of its triangular work through this single unrolled routine -- the panel rows below a diagonal block
// (l L11^T = a), the columns of the diagonal-block inverses (L x = e_c) and the block rows of the inverse
// (X_pp R = L_pp^-1 R) -- because it runs once per launch from a cold instruction cache: straight-line code costs ~8
// cycles per instruction on first touch (measured), rolled shared-memory loops ~35 cycles per element.
__device__ __forceinline__ void fwdsub32(double* vec, int sv, const double* Lb, const double* rdp) {
  double v[DIAG_PB];
#pragma unroll
  for (int c = 0; c < DIAG_PB; ++c) v[c] = vec[c * sv];
#pragma unroll
  for (int c = 0; c < DIAG_PB; ++c) {
    v[c] *= rdp[c];
#pragma unroll
    for (int c2 = c + 1; c2 < DIAG_PB; ++c2) v[c2] = fma(-v[c], Lb[c2 * DIAG_PITCH + c], v[c2]);
  }
#pragma unroll
  for (int c = 0; c < DIAG_PB; ++c) vec[c * sv] = v[c];
}

// Register-tile helpers of the diagonal-block kernel; thread (ty, tx) owns rows ty + 16 a, columns tx + 16 b in r[a][b].
// P is the panel index as a template parameter so that every register index is a compile-time constant while the panel
// loop itself stays rolled (the long straight-line warp routines above then exist once in the instruction stream).
template <int P>
__device__ __forceinline__ void diag_store_block_column(double (&r)[8][8], double* S, int ty, int tx) {
#pragma unroll
  for (int a = 2 * P; a < 8; ++a)
#pragma unroll
    for (int b = 2 * P; b < 2 * P + 2; ++b) {
      const int i = ty + 16 * a, k = tx + 16 * b;
      if (k <= i) S[i * DIAG_PITCH + k] = r[a][b];
    }
}

// rank-32 update of the trailing matrix from the finished panel P (tiles on or below the diagonal; diagonal tiles stay
// symmetric)
template <int P>
__device__ __forceinline__ void diag_syrk_update(double (&r)[8][8], const double* S, int ty, int tx) {
  constexpr int j0 = DIAG_PB * P;
#pragma unroll 4
  for (int c = 0; c < DIAG_PB; ++c) {
    double li[8], lk[8];
#pragma unroll
    for (int a = 2 * P + 2; a < 8; ++a) li[a] = S[(ty + 16 * a) * DIAG_PITCH + j0 + c];
#pragma unroll
    for (int b = 2 * P + 2; b < 8; ++b) lk[b] = S[(tx + 16 * b) * DIAG_PITCH + j0 + c];
#pragma unroll
    for (int a = 2 * P + 2; a < 8; ++a)
#pragma unroll
      for (int b = 2 * P + 2; b < 8; ++b)
        if (b <= a) r[a][b] = fma(-li[a], lk[b], r[a][b]);
  }
}

// rows of block P of the running right-hand side R (columns < j0), transposed into the strict upper triangle:
// S[c][j0 + k] = R[j0 + k][c]
template <int P>
__device__ __forceinline__ void diag_store_rhs_rows(double (&r)[8][8], double* S, int ty, int tx) {
#pragma unroll
  for (int a = 2 * P; a < 2 * P + 2; ++a)
#pragma unroll
    for (int b = 0; b < 2 * P; ++b) S[(tx + 16 * b) * DIAG_PITCH + ty + 16 * a] = r[a][b];
}

// R[i][c] -= sum_k L[i][j0+k] X[j0+k][c]   for rows below block P, columns up to the end of block P
template <int P>
__device__ __forceinline__ void diag_rhs_update(double (&r)[8][8], const double* S, const double* XDp, int ty, int tx) {
  constexpr int j0 = DIAG_PB * P;
#pragma unroll 4
  for (int k = 0; k < DIAG_PB; ++k) {
    double li[8], xr[8];
#pragma unroll
    for (int a = 2 * P + 2; a < 8; ++a) li[a] = S[(ty + 16 * a) * DIAG_PITCH + j0 + k];
#pragma unroll
    for (int b = 0; b < 2 * P + 2; ++b)
      xr[b] = (b < 2 * P) ? S[(tx + 16 * b) * DIAG_PITCH + j0 + k] : XDp[k * DIAG_XD_PITCH + tx + 16 * (b - 2 * P)];
#pragma unroll
    for (int a = 2 * P + 2; a < 8; ++a)
#pragma unroll
      for (int b = 0; b < 2 * P + 2; ++b) r[a][b] = fma(-li[a], xr[b], r[a][b]);
  }
}

// One CTA factors the 128x128 diagonal block kb of A (lower Cholesky, in place) and inverts the factor
// (X = L^-1, U = X^T).  Blocked inside the CTA with 32-wide panels so that the serial part is four warp-level
// 32x32 factorizations instead of 128 block-wide barrier rounds:
//   phase 1, per panel: the trailing matrix lives in registers; owners drop the block column into shared memory, warp 0
//     factors the diagonal 32x32 block (chol32_warp), one thread per row below solves its row against it (fwdsub32),
//     then every thread applies the rank-32 update to its registers from the finished panel.
//   phase 2: the columns of the four diagonal 32x32 inverses (fwdsub32 on unit vectors), then block forward
//     substitution L X = I with the running right-hand side in registers: per block row a fwdsub32 per column and a
//     rank-32 update; finished rows of X are kept TRANSPOSED in the strict upper triangle of the shared array
//     (diagonal blocks in XD).
// The eight stages run through ONE loop body so that chol32_warp and fwdsub32 exist once in the instruction stream.
// Only the first n_valid = clamp(n - 128 kb, 0, 128) rows are real data; the rest is identity padding for which every
// step is a no-op, so panels at or beyond n_valid are skipped.
// stage timestamps for tools/microbench/refit_time.cu (compiled in only there)
#ifdef GPO_DIAG_TIMING
__device__ long long g_diag_t[64];
#define DIAG_T(i) do { if (tid == 0 && blockIdx.x == 0) g_diag_t[i] = clock64(); } while (0)
#else
#define DIAG_T(i) do { } while (0)
#endif
// mode 0: both phases.  mode 1 (critical chain of the refit): Cholesky + the four diagonal 32 x 32 inverses only -- L to A,
// the inverses into the diagonal 32-blocks of X (what trsm_panel_kernel needs), 1 / L_jj to rdiag.  mode 2 (side stream): the
// inverse from the L and rdiag that mode 1 left: same instructions on the same numbers, X and U bit-identical to mode 0.
// The block to factor is read from src (row pitch src_ld; only its lower triangle): the matrix itself for the stand-alone
// kernel, the chain kernel's freshly updated copy otherwise (SRC_CG: written by other CTAs of the same launch, so the loads
// bypass L1).
template <bool SRC_CG>
__device__ __forceinline__ void diag_chol_inv_body(double* sm, const double* src, long long src_ld, double* A, double* X, double* U,
                                                   int n_pad, int kb, int n_valid, int* info, int mode, double* rdiag, int z) {
  double* S = sm;                                        // [128][129]  lower: L;  strict upper: X^T (off-diagonal blocks)
  double* XD = sm + TILE * DIAG_PITCH;                   // [4][32][33] inverses of the diagonal 32x32 blocks
  double* rd = XD + DIAG_NP * DIAG_PB * DIAG_XD_PITCH;   // [128]       1 / L_jj
  double* cbuf = rd + TILE;                              // [32][32]    chol32_warp's columns (16-byte aligned)
  volatile int* prog = reinterpret_cast<volatile int*>(cbuf + DIAG_PB * DIAG_PB);   // columns of the current block in cbuf
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ty = tid >> 4, tx = tid & 15;
  const long long base = ((long long)z * n_pad + (long long)kb * TILE) * n_pad + (long long)kb * TILE;
  DIAG_T(63);
  double r[8][8];
  if (mode != 2) {
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const int i = ty + 16 * a, k = tx + 16 * b;
        const double* ps = src + (long long)(k <= i ? i : k) * src_ld + (k <= i ? k : i);  // symmetric fill from the lower part
        r[a][b] = SRC_CG ? __ldcg(ps) : *ps;
        if (k > i) S[i * DIAG_PITCH + k] = 0.0;
      }
    if (tid < TILE) rd[tid] = 1.0;
  } else {
    for (int e = tid; e < TILE * TILE; e += DIAG_THREADS) {
      const int i = e >> 7, k = e & (TILE - 1);
      S[i * DIAG_PITCH + k] = (k <= i) ? A[base + (long long)i * n_pad + k] : 0.0;
    }
    if (tid < TILE) rd[tid] = rdiag[(long long)z * n_pad + kb * TILE + tid];
    __syncthreads();
  }
  const int s_begin = mode == 2 ? 4 : 0, s_end = mode == 1 ? 5 : 8;
  DIAG_T(0);

  // stages 0..3: Cholesky panel p = s;  stage 4: L out + diagonal-block inverses;  stages 5..7: block row p = s - 4 of X
#pragma unroll 1
  for (int s = s_begin; s < s_end; ++s) {
    const int p = s < 4 ? s : s - 4, j0 = DIAG_PB * p;
    const bool live = j0 < n_valid;                      // uniform; identity padding from j0 on otherwise
    double* vec = S;
    int sv = 1;
    const double* Lb = S + j0 * DIAG_PITCH + j0;
    const double* rdp = rd + j0;
    bool solve = false;
    if (s < 4) {
      if (p == 0) diag_store_block_column<0>(r, S, ty, tx);
      else if (p == 1) diag_store_block_column<1>(r, S, ty, tx);
      else if (p == 2) diag_store_block_column<2>(r, S, ty, tx);
      else diag_store_block_column<3>(r, S, ty, tx);
      if (tid == 0) *prog = 0;
      if (warp == 5) {                                   // unit right-hand sides of this block's inverse (L_pp x = e_c)
#pragma unroll 4
        for (int i = 0; i < DIAG_PB; ++i) XD[p * DIAG_PB * DIAG_XD_PITCH + i * DIAG_XD_PITCH + lane] = (i == lane) ? 1.0 : 0.0;
      }
      __syncthreads();
      DIAG_T(1 + 4 * s);
      if (live) {
        // warp 0 factors the diagonal 32 x 32 block; one column behind it warps 1-3 solve the rows below (thread = row:
        // l L_pp^T = a; rows of the identity padding hold zeros already) and warp 5 the columns of L_pp^-1
        if (warp == 0) chol32_warp(S, rd, cbuf, prog, j0, lane, info + z, kb * TILE);
        else if (warp <= 3) {
          const int i = j0 + DIAG_PB + tid - 32;
          if (p < DIAG_NP - 1 && i < n_valid) fwdsub32_trail(S + i * DIAG_PITCH + j0, 1, cbuf, rdp, prog);
        } else if (warp == 5) {
          fwdsub32_trail(XD + p * DIAG_PB * DIAG_XD_PITCH + lane, DIAG_XD_PITCH, cbuf, rdp, prog);
        }
      }
      DIAG_T(2 + 4 * s);
    } else if (s == 4) {
      if (mode != 2) {
        for (int e = tid; e < TILE * TILE; e += DIAG_THREADS) {  // L to global memory (coalesced rows)
          const int i = e >> 7, k = e & (TILE - 1);
          A[base + (long long)i * n_pad + k] = (k <= i) ? S[i * DIAG_PITCH + k] : 0.0;
        }
      } else if (tid < TILE) {                           // thread = (diagonal block w, column c): L_ww x = e_c
        const int w = tid >> 5;
        vec = XD + w * DIAG_PB * DIAG_XD_PITCH + lane;
        sv = DIAG_XD_PITCH;
#pragma unroll 4
        for (int i = 0; i < DIAG_PB; ++i) vec[i * sv] = (i == lane) ? 1.0 : 0.0;
        Lb = S + (DIAG_PB * w) * DIAG_PITCH + DIAG_PB * w;
        rdp = rd + DIAG_PB * w;
        solve = DIAG_PB * w < n_valid;
      }
    } else {
      if (live) {  // rows of block p of the running right-hand side, transposed into the strict upper triangle
        if (p == 1) diag_store_rhs_rows<1>(r, S, ty, tx);
        else if (p == 2) diag_store_rhs_rows<2>(r, S, ty, tx);
        else diag_store_rhs_rows<3>(r, S, ty, tx);
      }
      __syncthreads();
      solve = live && tid < j0;                          // thread = column c < j0: X[block p][c] = L_pp^-1 R[block p][c]
      vec = S + tid * DIAG_PITCH + j0;
    }
    if (solve) fwdsub32(vec, sv, Lb, rdp);
    __syncthreads();
    DIAG_T(3 + 4 * s);
    if (s < 3) {
      if (live) {
        if (p == 0) diag_syrk_update<0>(r, S, ty, tx);
        else if (p == 1) diag_syrk_update<1>(r, S, ty, tx);
        else diag_syrk_update<2>(r, S, ty, tx);
      }
    } else if (s == 4 && mode != 1) {
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) r[a][b] = (ty + 16 * a == tx + 16 * b) ? 1.0 : 0.0;  // running right-hand side R = I
      if (n_valid > 0) diag_rhs_update<0>(r, S, XD, ty, tx);
    } else if (s > 4 && live) {
      const double* XDp = XD + p * DIAG_PB * DIAG_XD_PITCH;
      if (p == 1) diag_rhs_update<1>(r, S, XDp, ty, tx);
      else if (p == 2) diag_rhs_update<2>(r, S, XDp, ty, tx);
    }
    DIAG_T(4 + 4 * s);
  }
  __syncthreads();
  DIAG_T(40);
  if (mode == 1) {   // the diagonal 32 x 32 inverses (final entries of X) and the reciprocal pivots; the rest follows in mode 2
    for (int e = tid; e < DIAG_NP * DIAG_PB * DIAG_PB; e += DIAG_THREADS) {
      const int w = e >> 10, i = (e >> 5) & 31, c = e & 31;
      X[base + (long long)(DIAG_PB * w + i) * n_pad + DIAG_PB * w + c] =
          (c <= i) ? XD[w * DIAG_PB * DIAG_XD_PITCH + i * DIAG_XD_PITCH + c] : 0.0;
    }
    if (tid < TILE) rdiag[(long long)z * n_pad + kb * TILE + tid] = rd[tid];
    DIAG_T(41);
    return;
  }
  // X (lower) and U = X^T (upper), coalesced rows
  for (int e = tid; e < TILE * TILE; e += DIAG_THREADS) {
    const int i = e >> 7, c = e & (TILE - 1);
    const int lo = c < i ? c : i, hi = c < i ? i : c;              // X[hi][lo] is the non-zero of the pair
    const bool same = (hi >> 5) == (lo >> 5);
    const double x = same ? XD[(hi >> 5) * DIAG_PB * DIAG_XD_PITCH + (hi & 31) * DIAG_XD_PITCH + (lo & 31)] : S[lo * DIAG_PITCH + hi];
    if (mode == 0 || !same) X[base + (long long)i * n_pad + c] = (c <= i) ? x : 0.0;   // (mode 2: the panel may be reading them)
    U[base + (long long)i * n_pad + c] = (c >= i) ? x : 0.0;
  }
  DIAG_T(41);
}

__global__ void __launch_bounds__(DIAG_THREADS) diag_chol_inv_kernel(double* A, double* X, double* U, int n_pad, int kb,
                                                                     int n_valid, int* info, int mode, double* rdiag) {
  extern __shared__ double sm[];
  const int z = blockIdx.x;
  const long long base = ((long long)z * n_pad + (long long)kb * TILE) * n_pad + (long long)kb * TILE;
  diag_chol_inv_body<false>(sm, A + base, n_pad, A, X, U, n_pad, kb, n_valid, info, mode, rdiag, z);
}

// out[z][r] = sum_{k in [lo(r), hi(r))} M[z][r][k] * v[z][k]; mode 0: k <= r (lower), 1: k >= r (upper), 2: full
__global__ void gemv_kernel(const double* M, const double* v, double* out, int n_pad, int mode) {
  const int z = blockIdx.y;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= n_pad) return;
  const int lo = mode == 1 ? r : 0, hi = mode == 0 ? r + 1 : n_pad;
  const double* row = M + ((long long)z * n_pad + r) * n_pad;
  const double* vv = v + (long long)z * n_pad;
  double s = 0.0;
  for (int k = lo + lane; k < hi; k += 32) s = fma(row[k], vv[k], s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[(long long)z * n_pad + r] = s;
}

// per-z scalar reductions after the factorisation:
//   logdet[z] = 2 sum_i log L_ii,  quad[z] = sum_i alpha_i y_i,  fmin[z] = min_i mean_i  (i < n)
__global__ void fit_scalars_kernel(const double* L, const double* alpha, const double* y, const double* mean,
                                   long long mean_ld, int js, int n, int n_pad, double* logdet, double* quad,
                                   double* fmin_out) {
  const int z = blockIdx.x;
  __shared__ double sh[3][32];
  double ld = 0.0, qd = 0.0, mn = DBL_MAX;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    ld += log(L[((long long)z * n_pad + i) * n_pad + i]);
    qd = fma(alpha[(long long)z * n_pad + i], y[i], qd);
    if (mean) {  // null: the refit proper (log-likelihood only); fmin follows when the state is finalised
      double mi = 0.0;
      for (int sp = 0; sp < js; ++sp) mi += mean[((long long)z * js + sp) * mean_ld + i];
      mn = ::fmin(mn, mi);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ld += __shfl_xor_sync(0xffffffffu, ld, o);
    qd += __shfl_xor_sync(0xffffffffu, qd, o);
    mn = ::fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  }
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sh[0][w] = ld; sh[1][w] = qd; sh[2][w] = mn; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0, c = DBL_MAX;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += sh[0][i]; b += sh[1][i]; c = ::fmin(c, sh[2][i]); }
    logdet[z] = 2.0 * a;
    quad[z] = b;
    if (mean) fmin_out[z] = c;
  }
}

// ML-II gradient: g[z][p] = sum_{i,j<n} dL_dK_ij dK_ij/dtheta_p with dL_dK = 0.5 (alpha alpha^T - W^-1):
//   p = 0: variance, p = 1..d: per-dimension lengthscale (summed by the host for non-ARD), p = d+1: noise (trace).
// GPy Stationary.update_gradients_full + Gaussian.exact_inference_gradients.  Output: per-block partials.
template <int DP>
__global__ void __launch_bounds__(256) loglik_grad_kernel(const double* W, const double* alpha, const double* Xt, long long ldx,
                                                          int n, int n_pad, int d, int kern, const double* invl,
                                                          const double* sigf2, double* partial /*[S][gridDim.x][d+2]*/) {
  // one warp per row i (8 rows per block), lanes over the columns j: W and the training coordinates are read in
  // coalesced rows, the row's own coordinates stay in registers, no integer division per element (the first version
  // decoded (i, j) from a flat index with a 64-bit divide and ran 256 blocks: 0.58 ms at n=4096)
  const int z = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + warp;
  double acc[DP + 2];
#pragma unroll
  for (int q = 0; q < DP + 2; ++q) acc[q] = 0.0;
  const double sf2 = sigf2[z], inv_sf2 = 1.0 / sf2;
  if (i < n) {
    double xi[DP], il[DP];
#pragma unroll
    for (int q = 0; q < DP; ++q) {
      il[q] = q < d ? invl[z * d + q] : 0.0;
      xi[q] = q < d ? Xt[q * ldx + i] * il[q] : 0.0;
    }
    const double ai = alpha[(long long)z * n_pad + i];
    const double* Wi = W + ((long long)z * n_pad + i) * n_pad;
    const double* az = alpha + (long long)z * n_pad;
    for (int j = lane; j < n; j += 32) {
      const double dl = 0.5 * (ai * az[j] - Wi[j]);
      double r2 = 0.0, dq[DP];
#pragma unroll
      for (int q = 0; q < DP; ++q) {
        const double df = q < d ? xi[q] - Xt[q * ldx + j] * il[q] : 0.0;
        dq[q] = df * df;
        r2 += dq[q];
      }
      double kv, hv;
      kern_k_h(kern, sf2, r2, kv, hv);
      acc[0] = fma(dl, kv * inv_sf2, acc[0]);
      // dK/dl_q = dK/dr * dr/dl_q = -(dK/dr / r) * (scaled diff_q)^2 / l_q  = -h * dq * invl_q
      const double dh = -(dl * hv);
#pragma unroll
      for (int q = 0; q < DP; ++q) acc[1 + q] = fma(dh, dq[q] * il[q], acc[1 + q]);
      if (i == j) acc[DP + 1] += dl;
    }
  }
  __shared__ double sh[8][DP + 2];
#pragma unroll
  for (int q = 0; q < DP + 2; ++q) {
    double v = acc[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sh[warp][q] = v;
  }
  __syncthreads();
  if (threadIdx.x < d + 2) {
    const int q = threadIdx.x, src = q < d + 1 ? q : DP + 1;   // the noise term sits behind the DP lengthscale slots
    double t = 0.0;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) t += sh[w8][src];
    partial[((long long)z * gridDim.x + blockIdx.x) * (d + 2) + q] = t;
  }
}

// =====================================================================================================
// K3: acquisition epilogue.  One thread per candidate: combines the per-row-block partials of the variance
// GEMM with K1's mean, applies GPModel's clip/sqrt, then EI / MPI / LCB (+ MCMC average) and, if set, the
// Local-Penalization transform + penalisers.  Reference: gpmodel.py:95-142, util/general.py:113-128,
// acquisitions/{EI,MPI,LCB,EI_mcmc,MPI_mcmc,LCB_mcmc,LP}.py.
// =====================================================================================================
enum AcqKind { ACQ_EI = 0, ACQ_MPI = 1, ACQ_LCB = 2, ACQ_EI_MCMC = 3, ACQ_MPI_MCMC = 4, ACQ_LCB_MCMC = 5, ACQ_NONE = -1 };

struct AcqParams {
  long long n_cand, nc_pad;
  int S, d, tiles_m;     // tiles_m = number of partial rows per hyper-parameter set (row blocks x k-splits)
  int part_nv;           // values per partial row: 1 (EPI_SUMSQ) or d_epi + 2 (EPI_GRAD)
  long long part_ld;     // leading dimension (candidates) of the partial buffers
  int grad;              // produce gradients (partials then hold S0, H0, P_1..P_d)
  int with_noise;
  int js;                // K1's slices of the training index: mean / dmean hold js partial sums per candidate
  int js_ld;             // slices per hyper-parameter set in the buffers (>= js; js = 1 after cov_small_kernel's own reduction)
  const double* mean;    // [S][js][nc_pad]
  const double* dmean;   // [S][js][d][nc_pad]
  const double* partial; // [S][tiles_m][nv][nc_pad]; nv = 1 (sum of squares) or d + 2 (S0, H0, P_1..P_d)
  const double* partial_sq; // optional [S][sq_tiles_m][nc_pad]: sum of squares |L^-1 k|^2 (strict variance in gradient mode)
  int sq_tiles_m;
  const double* invl;    // [S][d]
  const double* xmu;     // [d] centring offsets of the training inputs
  const double* sigf2;   // [S]
  const double* noise;   // [S]
  const double* fmin;    // [S]
  int kind;
  double par;
  // predict outputs (kind == ACQ_NONE), layout [S][N] / [S][N][d], already offset to this chunk's first row
  double *out_m, *out_s, *out_dm, *out_ds;
  long long out_ldN;     // N of the whole call (stride between hyper-parameter sets)
  // acquisition outputs, offset to the chunk's first row
  double *out_f, *out_df;
  // fused selection: every block leaves its best candidate (largest score = f, or smallest LP value; ties -> lowest
  // index) found with warp shuffles; gpo_acq_topk(k = 1) reduces these N/128 records instead of re-reading f
  double* blk_val;
  long long* blk_idx;
  long long row0;        // global index of the chunk's first row
  // S > 1: the kernel runs with gridDim.y = S (one hyper-parameter set per block row) and leaves each set's contribution
  // in zbuf [S][1+d][nc_pad]; acq_reduce_kernel then adds them in set order (deterministic) and finishes
  double* zbuf;
  // Local Penalization
  int lp_on, lp_K, lp_transform;
  const double *lp_xb, *lp_r, *lp_s;   // [K][d], [K], [K]
  const double* Xs;                    // candidates AoS (needed for the penalisers)
};

__device__ __forceinline__ double log_ndtr(double z) {
  // log Phi(z): erfcx form for the left tail (no underflow), log1p form otherwise (as scipy.special.log_ndtr)
  const double t = z * 0.70710678118654752440;
  if (z < -1.0) return log(0.5 * erfcx(-t)) - t * t;
  return log1p(-0.5 * erfc(t));
}

// strict "a is a better (score, index) than b": larger score first, NaN last (as np.argsort and topk_before order them),
// ties by lowest index
__device__ __forceinline__ bool argmax_better(double va, long long ia, double vb, long long ib) {
  const bool na = va != va, nb = vb != vb;
  if (na || nb) return (!na && nb) || (na == nb && ia < ib);
  return va > vb || (va == vb && ia < ib);
}

// block-wide argmax of (score, index) with warp shuffles; ties -> lowest index; NaN scores lose against any number
__device__ __forceinline__ void block_argmax_store(double score, long long idx, double* blk_val, long long* blk_idx,
                                                   long long slot) {
  __shared__ double sv[4];
  __shared__ long long si[4];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, score, o);
    const long long oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (argmax_better(ov, oi, score, idx)) { score = ov; idx = oi; }
  }
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = score; si[threadIdx.x >> 5] = idx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
      if (argmax_better(sv[w], si[w], score, idx)) { score = sv[w]; idx = si[w]; }
    blk_val[slot] = score;
    blk_idx[slot] = idx;
  }
}

// last stage of K3 for one candidate: MCMC average, then either the plain outputs or the Local-Penalization transform.
// (one thread per candidate; calls with a handful of points take acq_finalize_warp below)
__device__ __forceinline__ double acq_finalize(const AcqParams& p, long long i, double f_acc, double (&df_acc)[DMAX], bool store) {
  const int d = p.d;
  if (p.kind >= ACQ_EI_MCMC) {
    const double invS = 1.0 / (double)p.S;  // f_acqu/(len(means)): true division by S
    f_acc = f_acc / (double)p.S;
#pragma unroll
    for (int q = 0; q < DMAX; ++q) df_acc[q] = df_acc[q] / (double)p.S;
    (void)invS;
  }
  if (!p.lp_on) {
    if (store) {
      p.out_f[i] = f_acc;
      if (p.grad) {
#pragma unroll
        for (int q = 0; q < DMAX; ++q)
          if (q < d) p.out_df[i * d + q] = df_acc[q];
      }
    }
    return f_acc;
  }
  // ---- Local Penalization (LP.py:70-132).  acquisition_function(x) == -f_acc, so fval = f_acc. ----
  const double fval = f_acc;
  double val, scale;
  if (p.lp_transform == 1) {  // softplus
    if (fval >= 40.0) val = log(fval);
    else val = log(log1p(exp(fval)));
    scale = 1.0 / (log1p(exp(fval)) * (1.0 + exp(-fval)));
  } else {
    val = log(fval + 1e-50);
    scale = 1.0 / fval;
  }
  val = -val;
  double dsum = 0.0;
  for (int k = 0; k < p.lp_K; ++k) {
    double nm2 = 0.0;
    for (int q = 0; q < d; ++q) {
      const double df = p.Xs[i * d + q] - p.lp_xb[k * d + q];
      nm2 = fma(df, df, nm2);
    }
    const double nm = sqrt(nm2);
    const double zz = (nm - p.lp_r[k]) / p.lp_s[k];
    val -= log_ndtr(zz);
    if (p.grad) {
      const double h = 0.5 * erfc(-zz * 0.70710678118654752440);  // norm.cdf(z)
      double dk = 1.0 / (p.lp_s[k] * 2.50662827463100050242 * h) * exp(-zz * zz / 2.0) / nm;
      if (h < 1e-50) dk = 0.0;
      dsum += dk;
    }
  }
  if (store) {
    p.out_f[i] = val;
    if (p.grad) {
      // scale * (-df) - sum_k d_k (direction dropped, broadcast over coordinates: LP.py:100-103,132)
#pragma unroll
      for (int q = 0; q < DMAX; ++q)
        if (q < d) p.out_df[i * d + q] = scale * (-df_acc[q]) - dsum;
    }
  }
  return -val;  // the penalised value is minimised
}

// K3 for one candidate and the hyper-parameter sets z_begin .. z_end-1: posterior (m, s, gradients) from the partial rows, then
// either the predict outputs (kind == ACQ_NONE, written when `store`) or the acquisition contributions added into f_acc / df_acc
__device__ __forceinline__ void acq_accumulate(const AcqParams& p, long long i, int z_begin, int z_end, bool store, double& f_acc,
                                               double (&df_acc)[DMAX]) {
  const int d = p.d, nv = p.part_nv;
  for (int z = z_begin; z < z_end; ++z) {
    double m = 0.0;
    for (int sp = 0; sp < p.js; ++sp) m += p.mean[((long long)z * p.js_ld + sp) * p.nc_pad + i];
    const double* part = p.partial + (long long)z * p.tiles_m * nv * p.part_ld + i;
    double s0 = 0.0;
    if (p.grad && p.partial_sq) {
      // strict variance: the reference's own form sigma_f^2 - |L^-1 k|^2 (PosteriorExact._raw_predict), immune to
      // the cancellation of k.W^-1.k on ill-conditioned Gram matrices
      const double* psq = p.partial_sq + (long long)z * p.sq_tiles_m * p.nc_pad + i;
      for (int tm = 0; tm < p.sq_tiles_m; ++tm) s0 += psq[(long long)tm * p.nc_pad];
    } else {
      for (int tm = 0; tm < p.tiles_m; ++tm) s0 += part[(long long)tm * nv * p.part_ld];
    }
    double v = p.sigf2[z] - s0;
    if (p.with_noise) v += p.noise[z];
    v = ::fmax(v, 1e-10);
    double s = sqrt(v);
    double dm[DMAX], ds[DMAX];
    if (p.grad) {
      // dv/dx_q = -2 sum_j b_j h_j (x_q - x_jq)/l_q^2 = -2/l_q^2 ((x_q - mu_q) H0 - P_q);  dsdx = dvdx / (2 s)
      const double inv2s = 1.0 / (2.0 * s);
      double h0 = 0.0;
      for (int tm = 0; tm < p.tiles_m; ++tm) h0 += part[((long long)tm * nv + 1) * p.part_ld];
#pragma unroll
      for (int q = 0; q < DMAX; ++q) {
        if (q < d) {
          double pq = 0.0;
          for (int tm = 0; tm < p.tiles_m; ++tm) pq += part[((long long)tm * nv + 2 + q) * p.part_ld];
          const double il = p.invl[z * d + q];
          const double gq = ((p.Xs[i * d + q] - p.xmu[q]) * h0 - pq) * il * il;
          double dmq = 0.0;
          for (int sp = 0; sp < p.js; ++sp) dmq += p.dmean[(((long long)z * p.js_ld + sp) * d + q) * p.nc_pad + i];
          dm[q] = dmq;
          ds[q] = -2.0 * gq * inv2s;
        } else { dm[q] = 0.0; ds[q] = 0.0; }
      }
    }
    if (p.kind == ACQ_NONE) {
      if (!store) continue;
      p.out_m[(long long)z * p.out_ldN + i] = m;
      if (p.out_s) p.out_s[(long long)z * p.out_ldN + i] = s;  // null: mean-only call (no variance partials exist)
      if (p.grad) {
#pragma unroll
        for (int q = 0; q < DMAX; ++q)
          if (q < d) {
            p.out_dm[((long long)z * p.out_ldN + i) * d + q] = dm[q];
            if (p.out_ds) p.out_ds[((long long)z * p.out_ldN + i) * d + q] = ds[q];
          }
      }
      continue;
    }
    const int base_kind = p.kind % 3;
    if (base_kind == ACQ_LCB) {
      f_acc += -m + p.par * s;
      if (p.grad) {
#pragma unroll
        for (int q = 0; q < DMAX; ++q) df_acc[q] += -dm[q] + p.par * ds[q];
      }
    } else {
      if (s < 1e-10) s = 1e-10;  // get_quantiles clamp (util/general.py:121-124)
      const double fm = p.fmin[z];
      const double u = (fm - m - p.par) / s;
      const double phi = exp(-0.5 * u * u) * 0.39894228040143267794;
      const double Phi = 0.5 * erfc(-u * 0.70710678118654752440);
      if (base_kind == ACQ_EI) {
        if (p.kind == ACQ_EI_MCMC) f_acc += (fm - m + p.par) * Phi + s * phi;  // EI_mcmc.py:38,51 (+jitter quirk)
        else f_acc += s * (u * Phi + phi);
        if (p.grad) {
#pragma unroll
          for (int q = 0; q < DMAX; ++q) df_acc[q] += ds[q] * phi - Phi * dm[q];
        }
      } else {  // MPI
        f_acc += Phi;
        if (p.grad) {
          const double c = -(phi / s);
#pragma unroll
          for (int q = 0; q < DMAX; ++q) df_acc[q] += c * (dm[q] + ds[q] * u);
        }
      }
    }
  }
}

__global__ void __launch_bounds__(128) acq_kernel(const AcqParams p) {
  const long long i_raw = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = i_raw < p.n_cand;
  const long long i = active ? i_raw : p.n_cand - 1;  // out-of-range threads shadow the last row (identical values, benign)
  const int d = p.d;
  double f_acc = 0.0, df_acc[DMAX];
#pragma unroll
  for (int q = 0; q < DMAX; ++q) df_acc[q] = 0.0;
  const bool zsplit = gridDim.y > 1;
  const int z_begin = zsplit ? blockIdx.y : 0, z_end = zsplit ? blockIdx.y + 1 : p.S;
  acq_accumulate(p, i, z_begin, z_end, true, f_acc, df_acc);
  if (p.kind == ACQ_NONE) return;
  if (zsplit) {
    double* zb = p.zbuf + (long long)z_begin * (d + 1) * p.nc_pad + i;
    zb[0] = f_acc;
    if (p.grad) {
#pragma unroll
      for (int q = 0; q < DMAX; ++q)
        if (q < d) zb[(long long)(q + 1) * p.nc_pad] = df_acc[q];
    }
    return;
  }
  const double score = acq_finalize(p, i, f_acc, df_acc, true);
  if (p.blk_val) block_argmax_store(score, p.row0 + i, p.blk_val, p.blk_idx, (p.row0 >> 7) + blockIdx.x);
}

// (declarations of the device L-BFGS, described in its own section below)
constexpr int LB_HIST = 10;
constexpr int LB_MAXM = 64;

struct LbLineSearch {   // dcsrch's saved state
  double stx, fx, gx, sty, fy, gy, stmin, stmax, width, width1, finit, ginit, gtest, stpmx;
  int brackt, stage, nls, pad;
};

struct LbfgsParams {
  int M, d, negate, maxiter;        // negate: the sweep returns the acquisition (maximised): minimise -f, gradient -df
  double pgtol, ftol;               // ftol = factr * machine epsilon
  const double* lo; const double* hi;   // [d] box
  double* x; double* g; double* f;  // [M][d], [M][d], [M]   current iterate
  double* xt;                       // [M][d]  trial points = input of the next evaluation
  const double* ft; const double* gt;   // [M], [M][d]  value / gradient at the trial points (sweep outputs)
  double* dir;                      // [M][d]  search direction
  double* S; double* Y; double* rho;    // [M][LB_HIST][d] x 2, [M][LB_HIST]
  double* step; double* gd;         // [M] current step length, directional derivative g . dir at the iterate
  LbLineSearch* ls;                 // [M]
  int* hist_n; int* hist_head; int* iter; int* phase; int* done; int* nfev;   // [M]
  int* n_done;                      // [1]
};

// per-lane registers holding one anchor's optimiser state between the moment its loads are issued (the top of acq_small_kernel,
// so that their round trips run under the evaluation epilogue) and the moment lbfgs_anchor_update spills them to shared memory
struct LbStage {
  double x, g, dir, lo, hi, xt, S[5], Y[5], rho, lsw, dsc;
  int isc, done;
};
__device__ __forceinline__ void lbfgs_stage_issue(const LbfgsParams& p, int a, int lane, LbStage& st);
__device__ void lbfgs_anchor_update(const LbfgsParams& p, int a, int lane, const LbStage& st);

// K3 for calls with a handful of candidates (the L-BFGS steps): ONE WARP per candidate (grid = candidates, 32 threads).  For a
// single thread K3 is a latency chain run from a cold instruction cache -- 72 us for 5 points with 15 Local-Penalization
// penalisers at n = 4096 -- so the lanes split the penalisers (acq_finalize_warp) and lane 0 stores.  Hyper-parameter sets are
// walked in set order inside the warp (no zbuf / zsum / reduce launches).  With `lb` set the same warp continues into the
// optimiser step of anchor `candidate` (lbfgs_anchor_update): evaluation epilogue and L-BFGS state machine in one launch.
// Lane q of the warp owns coordinate q (d <= DMAX <= 32): the per-coordinate partial sums, gradient terms and outputs are one
// value per lane, so their loads go out together (one thread walking the coordinates paid ~50 dependent L2 round trips: 29 us of
// the 47 us this kernel took at n = 2048) and the code is a sixteenth of the unrolled form.  Per element the arithmetic is
// acq_accumulate's / acq_finalize's, operation for operation: identical bits.
__device__ __forceinline__ void acq_accumulate_warp(const AcqParams& p, long long i, int lane, double& f_acc, double& df_lane) {
  const int d = p.d, nv = p.part_nv;
  const int q = lane < d ? lane : d - 1;   // lanes past d shadow the last coordinate (never stored)
  for (int z = 0; z < p.S; ++z) {
    double m = 0.0;
    for (int sp = 0; sp < p.js; ++sp) m += p.mean[((long long)z * p.js_ld + sp) * p.nc_pad + i];
    const double* part = p.partial + (long long)z * p.tiles_m * nv * p.part_ld + i;
    double s0 = 0.0;
    if (p.grad && p.partial_sq) {
      const double* psq = p.partial_sq + (long long)z * p.sq_tiles_m * p.nc_pad + i;
      for (int tm = 0; tm < p.sq_tiles_m; ++tm) s0 += psq[(long long)tm * p.nc_pad];
    } else {
      for (int tm = 0; tm < p.tiles_m; ++tm) s0 += part[(long long)tm * nv * p.part_ld];
    }
    double v = p.sigf2[z] - s0;
    if (p.with_noise) v += p.noise[z];
    v = ::fmax(v, 1e-10);
    double s = sqrt(v);
    double dmq = 0.0, dsq = 0.0;
    if (p.grad) {
      const double inv2s = 1.0 / (2.0 * s);
      double h0 = 0.0;
      for (int tm = 0; tm < p.tiles_m; ++tm) h0 += part[((long long)tm * nv + 1) * p.part_ld];
      double pq = 0.0;
      for (int tm = 0; tm < p.tiles_m; ++tm) pq += part[((long long)tm * nv + 2 + q) * p.part_ld];
      const double il = p.invl[z * d + q];
      const double gq = ((p.Xs[i * d + q] - p.xmu[q]) * h0 - pq) * il * il;
      for (int sp = 0; sp < p.js; ++sp) dmq += p.dmean[(((long long)z * p.js_ld + sp) * d + q) * p.nc_pad + i];
      dsq = -2.0 * gq * inv2s;
    }
    if (p.kind == ACQ_NONE) {
      if (lane == 0) {
        p.out_m[(long long)z * p.out_ldN + i] = m;
        if (p.out_s) p.out_s[(long long)z * p.out_ldN + i] = s;
      }
      if (p.grad && lane < d) {
        p.out_dm[((long long)z * p.out_ldN + i) * d + q] = dmq;
        if (p.out_ds) p.out_ds[((long long)z * p.out_ldN + i) * d + q] = dsq;
      }
      continue;
    }
    const int base_kind = p.kind % 3;
    if (base_kind == ACQ_LCB) {
      f_acc += -m + p.par * s;
      if (p.grad) df_lane += -dmq + p.par * dsq;
    } else {
      if (s < 1e-10) s = 1e-10;
      const double fm = p.fmin[z];
      const double u = (fm - m - p.par) / s;
      const double phi = exp(-0.5 * u * u) * 0.39894228040143267794;
      const double Phi = 0.5 * erfc(-u * 0.70710678118654752440);
      if (base_kind == ACQ_EI) {
        if (p.kind == ACQ_EI_MCMC) f_acc += (fm - m + p.par) * Phi + s * phi;
        else f_acc += s * (u * Phi + phi);
        if (p.grad) df_lane += dsq * phi - Phi * dmq;
      } else {
        f_acc += Phi;
        if (p.grad) {
          const double c = -(phi / s);
          df_lane += c * (dmq + dsq * u);
        }
      }
    }
  }
}

__device__ __forceinline__ void acq_finalize_warp(const AcqParams& p, long long i, double f_acc, double df_lane, int lane) {
  const int d = p.d;
  if (p.kind >= ACQ_EI_MCMC) {
    f_acc = f_acc / (double)p.S;
    df_lane = df_lane / (double)p.S;
  }
  if (!p.lp_on) {
    if (lane == 0) p.out_f[i] = f_acc;
    if (p.grad && lane < d) p.out_df[i * d + lane] = df_lane;
    return;
  }
  const double fval = f_acc;
  double val, scale;
  if (p.lp_transform == 1) {
    if (fval >= 40.0) val = log(fval);
    else val = log(log1p(exp(fval)));
    scale = 1.0 / (log1p(exp(fval)) * (1.0 + exp(-fval)));
  } else {
    val = log(fval + 1e-50);
    scale = 1.0 / fval;
  }
  val = -val;
  double dsum = 0.0, pen = 0.0;
  for (int k = lane; k < p.lp_K; k += 32) {
    double nm2 = 0.0;
    for (int q = 0; q < d; ++q) {
      const double df = p.Xs[i * d + q] - p.lp_xb[k * d + q];
      nm2 = fma(df, df, nm2);
    }
    const double nm = sqrt(nm2);
    const double zz = (nm - p.lp_r[k]) / p.lp_s[k];
    pen += log_ndtr(zz);
    if (p.grad) {
      const double h = 0.5 * erfc(-zz * 0.70710678118654752440);
      double dk = 1.0 / (p.lp_s[k] * 2.50662827463100050242 * h) * exp(-zz * zz / 2.0) / nm;
      if (h < 1e-50) dk = 0.0;
      dsum += dk;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    pen += __shfl_xor_sync(0xffffffffu, pen, o);
    dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
  }
  val -= pen;
  if (lane == 0) p.out_f[i] = val;
  if (p.grad && lane < d) p.out_df[i * d + lane] = scale * (-df_lane) - dsum;
}

__global__ void __launch_bounds__(32) acq_small_kernel(const AcqParams p, const LbfgsParams lb, int with_lb) {
  const long long i = blockIdx.x;
  const int lane = threadIdx.x;
  double f_acc = 0.0, df_lane = 0.0;
  LbStage stg;
  if (with_lb) lbfgs_stage_issue(lb, (int)i, lane, stg);
#ifdef GPO_ACQ_TIMING
  const long long t0 = clock64();
#endif
  acq_accumulate_warp(p, i, lane, f_acc, df_lane);
  if (p.kind == ACQ_NONE) return;
#ifdef GPO_ACQ_TIMING
  const long long t1 = clock64();
#endif
  acq_finalize_warp(p, i, f_acc, df_lane, lane);
#ifdef GPO_ACQ_TIMING
  const long long t2 = clock64();
#endif
  if (with_lb) {
    __threadfence_block();
    __syncwarp();
    lbfgs_anchor_update(lb, (int)i, lane, stg);
  }
#ifdef GPO_ACQ_TIMING
  if (lane == 0 && i == 0) printf("acq_small cycles: accumulate %lld finalize %lld lbfgs %lld\n", t1 - t0, t2 - t1, clock64() - t2);
#endif
}

// S > 1, stage 1: grid (candidates / 128, 1 + d).  Thread (i, v) adds value v of candidate i over the hyper-parameter sets
// in set order (deterministic, the reference's own accumulation order) and leaves the sum in set 0's slot of zbuf.
__global__ void __launch_bounds__(128) acq_zsum_kernel(const AcqParams p) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int v = blockIdx.y;
  if (i >= p.n_cand) return;
  const long long stride = (long long)(p.d + 1) * p.nc_pad;
  double* zb = p.zbuf + (long long)v * p.nc_pad + i;
  double acc = 0.0;
  for (int z = 0; z < p.S; ++z) acc += zb[(long long)z * stride];
  zb[0] = acc;
}

// S > 1, stage 2: MCMC average, Local Penalization, outputs and the fused block argmax
__global__ void __launch_bounds__(128) acq_reduce_kernel(const AcqParams p) {
  const long long i_raw = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long i = i_raw < p.n_cand ? i_raw : p.n_cand - 1;  // shadow the last row (identical values, benign)
  const int d = p.d;
  double f_acc = p.zbuf[i], df_acc[DMAX];
#pragma unroll
  for (int q = 0; q < DMAX; ++q) df_acc[q] = (p.grad && q < d) ? p.zbuf[(long long)(q + 1) * p.nc_pad + i] : 0.0;
  const double score = acq_finalize(p, i, f_acc, df_acc, true);
  if (p.blk_val) block_argmax_store(score, p.row0 + i, p.blk_val, p.blk_idx, (p.row0 >> 7) + blockIdx.x);
}

// Parity guard run after every refit: at the n training inputs (where the posterior variance is smallest and the
// cancellation worst) compare s from the triangular form sigma_f^2 - |L^-1 k|^2 with s from k.W^-1.k.
// out[z] = max_i |s_L - s_W|.
__global__ void variance_probe_kernel(const double* part_sq, const double* part_gr, int tiles_m, int nv, long long nc_pad,
                                      int n, const double* sigf2, const double* noise, double* out) {
  const int z = blockIdx.x;
  __shared__ double sh[32];
  double mx = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double q = 0.0, s0 = 0.0;
    for (int tm = 0; tm < tiles_m; ++tm) {
      q += part_sq[((long long)z * tiles_m + tm) * nc_pad + i];
      s0 += part_gr[((long long)(z * tiles_m + tm) * nv) * nc_pad + i];
    }
    const double sl = sqrt(::fmax(sigf2[z] - q + noise[z], 1e-10)), sw = sqrt(::fmax(sigf2[z] - s0 + noise[z], 1e-10));
    mx = ::fmax(mx, fabs(sl - sw));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = ::fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) mx = ::fmax(mx, sh[i]);
    out[z] = mx;
  }
}

// =====================================================================================================
// top-k selection (anchor_points_generator.py:67): k smallest (or largest) values, ties -> lowest index.
// Level kernel: each block bitonic-sorts TOPK_SEG (value, index) pairs in shared memory and emits its first k.
// =====================================================================================================
constexpr int TOPK_SEG = 2048;
constexpr int TOPK_THREADS = 512;

__device__ __forceinline__ bool topk_before(double va, long long ia, double vb, long long ib) {
  // strict "a sorts before b": smaller value first, NaN last, ties by index
  const bool na = va != va, nb = vb != vb;
  if (na || nb) return (!na && nb) || (na == nb && ia < ib);
  return va < vb || (va == vb && ia < ib);
}

__global__ void __launch_bounds__(TOPK_THREADS) topk_level_kernel(const double* vals, const long long* idx_in, long long n,
                                                                  int k, double sign, double* out_vals,
                                                                  long long* out_idx) {
  __shared__ double sv[TOPK_SEG];
  __shared__ long long si[TOPK_SEG];
  const long long base = (long long)blockIdx.x * TOPK_SEG;
  for (int t = threadIdx.x; t < TOPK_SEG; t += TOPK_THREADS) {
    const long long g = base + t;
    if (g < n) {
      sv[t] = idx_in ? vals[g] : sign * vals[g];
      si[t] = idx_in ? idx_in[g] : g;
    } else {
      sv[t] = __longlong_as_double(0x7ff8000000000000LL);  // NaN sorts last
      si[t] = 0x7fffffffffffffffLL;
    }
  }
  __syncthreads();
  for (int size = 2; size <= TOPK_SEG; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < TOPK_SEG / 2; t += TOPK_THREADS) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool up = (lo & size) == 0;
        const bool hi_first = topk_before(sv[hi], si[hi], sv[lo], si[lo]);
        if (hi_first == up) {
          const double tv = sv[lo]; sv[lo] = sv[hi]; sv[hi] = tv;
          const long long ti = si[lo]; si[lo] = si[hi]; si[hi] = ti;
        }
      }
      __syncthreads();
    }
  }
  for (int t = threadIdx.x; t < k; t += TOPK_THREADS) {
    out_vals[(long long)blockIdx.x * k + t] = sv[t];
    out_idx[(long long)blockIdx.x * k + t] = si[t];
  }
}

// last step of a top-k whose results stay on the device: undo the sign, pad past min(k, N) with (NaN, -1)
__global__ void topk_finish_kernel(const double* __restrict__ sv, const long long* __restrict__ si, int k, long long kk,
                                   double sign, double* __restrict__ vals, long long* __restrict__ idx) {
  for (int t = threadIdx.x; t < k; t += blockDim.x) {
    vals[t] = t < kk ? sign * sv[t] : __longlong_as_double(0x7ff8000000000000LL);
    idx[t] = t < kk ? si[t] : -1LL;
  }
}

// --------------------------------------------------------------------------------------------------
// Device-side random design: uniform candidates from a counter-based generator, so that a sweep
// needs no host->device copy of X*.  The stream is Philox4x64-10 exactly as NumPy consumes it:
//   numpy.random.Generator(numpy.random.Philox(key=key)).uniform(lo, hi, size=(N, d))     (row-major)
// element e takes lane (e & 3) of the block whose counter is (e >> 2) + 1 (NumPy increments the counter BEFORE it
// generates a block), u = (v >> 11) * 2^-53, x = lo + (hi - lo) * u with the product and the sum rounded
// separately (NumPy's random_uniform is not FMA-contracted).  Replaces samples_multidimensional_uniform
// (GPyOpt/experiment_design/random_design.py:66-76), whose legacy MT19937 stream is inherently sequential.
// --------------------------------------------------------------------------------------------------
struct DesignParams {
  unsigned long long key0, key1;
  long long e0;    // first global element index (row0 * d)
  long long n_el;  // elements to generate
  int d;
  double lo[DMAX], range[DMAX];
  double* out;     // element e0 is written to out[0]
};

__device__ __forceinline__ void philox4x64_10(unsigned long long c0, unsigned long long c1, unsigned long long c2,
                                              unsigned long long c3, unsigned long long k0, unsigned long long k1,
                                              unsigned long long (&v)[4]) {
  constexpr unsigned long long M0 = 0xD2E7470EE14C6C93ULL, M1 = 0xCA5A826395121157ULL;
  constexpr unsigned long long W0 = 0x9E3779B97F4A7C15ULL, W1 = 0xBB67AE8584CAA73BULL;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    if (r) { k0 += W0; k1 += W1; }
    const unsigned long long hi0 = __umul64hi(M0, c0), lo0 = M0 * c0;
    const unsigned long long hi1 = __umul64hi(M1, c2), lo1 = M1 * c2;
    const unsigned long long n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
  }
  v[0] = c0; v[1] = c1; v[2] = c2; v[3] = c3;
}

__global__ void __launch_bounds__(256) design_uniform_kernel(const DesignParams p) {
  const long long b_first = p.e0 >> 2;
  const long long e_end = p.e0 + p.n_el;
  const long long nblk = ((e_end + 3) >> 2) - b_first;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < nblk; t += (long long)gridDim.x * blockDim.x) {
    const unsigned long long b = (unsigned long long)(b_first + t);
    unsigned long long v[4];
    philox4x64_10(b + 1ULL, 0ULL, 0ULL, 0ULL, p.key0, p.key1, v);  // N*d < 2^66: no carry into the second counter word
    int q = (int)((b << 2) % (unsigned long long)p.d);
#pragma unroll
    for (int lane = 0; lane < 4; ++lane) {
      const long long e = (long long)(b << 2) + lane;
      if (e >= p.e0 && e < e_end) {
        const double u = (double)(v[lane] >> 11) * (1.0 / 9007199254740992.0);
        p.out[e - p.e0] = __dadd_rn(p.lo[q], __dmul_rn(p.range[q], u));
      }
      if (++q == p.d) q = 0;
    }
  }
}

// rows idx[0..k) of X [N, d] -> out [k, d]  (the winning candidates of a device-generated design)
__global__ void gather_rows_kernel(const double* __restrict__ X, const long long* __restrict__ idx, int k, int d,
                                   double* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= k * d) return;
  const int r = t / d, q = t - r * d;
  const long long i = idx[r];
  out[t] = i >= 0 ? X[i * d + q] : __longlong_as_double(0x7ff8000000000000LL);
}

// --------------------------------------------------------------------------------------------------
// Latency path for a handful of candidates (the L-BFGS steps of the acquisition optimiser: N = 1 ... 8 per call).
// The tensor-core GEMM works on 128-candidate tiles, so a single point costs 128x its flops (2 n^2 * 128 = 4.3 GFLOP at
// n = 4096, ~0.3 ms).  Here b = W^-1 k is a batched matrix-vector product streamed once over W^-1 (HBM/L2-bound:
// 8 n^2 bytes), fused with the same epilogue reductions as EPI_GRAD:
//   S0_i = sum_j b_ji k_ji,   H0_i = sum_j b_ji h_ji,   P_qi = sum_j b_ji h_ji xc_jq
// One warp owns R rows j of W^-1 and NC candidates (R x NC accumulators, lanes stride over the columns: coalesced 256-byte
// row segments, each k value reused for R rows), one CTA = 8 warps = 8 R rows; every CTA leaves one partial row for K3,
// which adds them in row-block order (deterministic; a candidate's result does not depend on NC or on its neighbours).
// --------------------------------------------------------------------------------------------------
struct SmallNParams {
  const double* W;   long long w_z;      // [S][n_pad][n_pad]
  const double* Kt;  const double* Ht;   // [S][nc_ld][n_pad] candidate-major rows from K1 (Ht null: RBF, h = -k)
  long long kt_z;
  const double* Xtc; int ldx;            // centred training coordinates [d][ldx]
  int n_pad, d;                          // d = 0: value only (S0 and H0)
  double* partial;   long long part_z;   // [S][row blocks][d + 2][ldp]
  int ldp;
  double* final_rows;                    // [S][nvals][final_ld]: the partial rows added in row-block order by the last CTA
  long long final_ld;
  int* ticket;                           // [S] arrival counters (zero before the launch; the last CTA resets its own)
};

// SQ = true: the strict-variance form instead.  W is then X = L^-1 (lower triangular: the column loop stops at the
// diagonal) and the only reduction is sum_j (X k)_j^2 = |L^-1 k|^2, one value per candidate.
// One CTA owns SMALLN_ROWS rows (four per warp) and streams them in chunks of SMALLN_CW columns through a ring of
// SMALLN_STAGES shared-memory stages filled by asynchronous copies (GPO_SMALLN_CPASYNC below): three stages per CTA are in flight
// without holding a register.  Plain loads did not get there -- the compiler schedules the FMAs of step t ahead of the loads of
// step t+1, each warp had ~1.5 steps in flight and the kernel ran at 2.6 of HBM's 7.7 TB/s with every warp on the long
// scoreboard (ncu, profiles/).
// SMALLN_ROWS is a template parameter (32 / 16 / 8 rows per CTA, i.e. 4 / 2 / 1 per warp): smaller matrices need more,
// thinner CTAs to put every SM's load path to work (n = 2048: 64 CTAs of 32 rows took 25 us for 32 MB out of L2).
constexpr int SMALLN_CW = 128, SMALLN_STAGES = 4;
constexpr int smalln_smem_bytes(int rows) { return SMALLN_STAGES * rows * SMALLN_CW * 8 + SMALLN_STAGES * 8; }
// How the ring is filled.  1 (default): 16-byte cp.async copies issued by all 256 threads, one commit group and ONE block barrier per
// chunk (the stage freed by chunk t - 1 takes chunk t + STAGES - 1).  0: the round-1 form, one 1 KB cp.async.bulk per row and chunk on
// an mbarrier per stage.  The bulk form is NOT safe: on launches that run in several waves with three or more CTAs per SM it
// returned wrong partial sums for whole row blocks -- once in 10 ... 400 calls with 5 hyper-parameter sets at n = 684, or 2 sets at
// n = 3000, or one set at n >= 5000 (variance clamped to 1e-10, or garbage) -- first seen as ONE failing case of the 250-case
// randomised oracle sweep, then reproduced with tools/stress_latency_path.py (the same call repeated thousands of times must return
// the same bits).  Ruled out by experiment: the arrival-counter protocol (a separate reduction kernel sees the same wrong partial
// rows), the read-only data path of the K* loads, re-initialising un-invalidated mbarriers, 2-D tensor copies instead of bulk
// copies (worse), capping the residency at two CTAs per SM (rarer, not gone).  What remains is the bulk copies' data itself; with
// compute-sanitizer closed on this pool the cause stays open.  The cp.async form has never returned a differing bit in 60,000
// stress calls over the shapes that failed, and is a little faster (n = 4096, five points: 0.127 -> 0.120 ms).
#ifndef GPO_SMALLN_CPASYNC
#define GPO_SMALLN_CPASYNC 1
#endif
__device__ __forceinline__ void sn_cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void sn_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void sn_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NC, bool SQ, int SMALLN_ROWS>
__device__ __forceinline__ void smalln_wk_body(const SmallNParams& p) {
  extern __shared__ __align__(128) double smalln_smem[];
  __shared__ double red[8][DMAX + 2][NC];
  double* tile = smalln_smem;                                                            // [STAGES][ROWS][CW]
  uint64_t* full = reinterpret_cast<uint64_t*>(smalln_smem + SMALLN_STAGES * SMALLN_ROWS * SMALLN_CW);
  constexpr int R = SMALLN_ROWS / 8;
  const int z = blockIdx.y, rb = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_pad = p.n_pad, d = SQ ? -1 : p.d;   // d + 2 values per candidate (SQ: one)
  const int row0 = rb * SMALLN_ROWS + warp * R;
  const int nchunks = SQ ? (rb * SMALLN_ROWS + SMALLN_ROWS + SMALLN_CW - 1) / SMALLN_CW : n_pad / SMALLN_CW;
  const double* Kt = p.Kt + (long long)z * p.kt_z;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < SMALLN_STAGES; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  auto issue_cpasync = [&](int t) {   // every thread: chunk t (or nothing past the end) as 16-byte cp.async copies, one commit group
    if (t < nchunks) {
      const int s = t % SMALLN_STAGES;
      const double* Wblk = p.W + (long long)z * p.w_z + (long long)rb * SMALLN_ROWS * n_pad + (long long)t * SMALLN_CW;
      for (int e = threadIdx.x; e < SMALLN_ROWS * (SMALLN_CW / 2); e += 256) {
        const int row = e / (SMALLN_CW / 2), ch = e % (SMALLN_CW / 2);
        sn_cp_async16(tile + ((long long)s * SMALLN_ROWS + row) * SMALLN_CW + 2 * ch, Wblk + (long long)row * n_pad + 2 * ch);
      }
    }
    sn_cp_async_commit();
  };
  auto issue = [&](int t) {  // (bulk form) warp 0: chunk t of all rows into stage t % STAGES
    const int s = t % SMALLN_STAGES;
    const double* Wblk = p.W + (long long)z * p.w_z + (long long)rb * SMALLN_ROWS * n_pad;
    if (lane == 0) mbar_arrive_expect_tx(&full[s], SMALLN_ROWS * SMALLN_CW * 8);
    __syncwarp();
    if (lane < SMALLN_ROWS)
      tma_bulk_load(tile + ((long long)s * SMALLN_ROWS + lane) * SMALLN_CW, Wblk + (long long)lane * n_pad + (long long)t * SMALLN_CW,
                    SMALLN_CW * 8, &full[s]);
  };
  if (GPO_SMALLN_CPASYNC) {
    for (int t = 0; t < SMALLN_STAGES - 1; ++t) issue_cpasync(t);
  } else if (warp == 0)
    for (int t = 0; t < SMALLN_STAGES && t < nchunks; ++t) issue(t);
  double acc[R][NC];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int i = 0; i < NC; ++i) acc[r][i] = 0.0;
  for (int t = 0; t < nchunks; ++t) {
    const int s = t % SMALLN_STAGES;
    double2 kv[2][NC];   // this chunk of the candidates' covariance vectors (shared by every CTA: L1 / L2 hits)
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int i = 0; i < NC; ++i)
        kv[u][i] = __ldg(reinterpret_cast<const double2*>(Kt + (long long)i * n_pad + t * SMALLN_CW + 64 * u + 2 * lane));
    if (GPO_SMALLN_CPASYNC) {
      sn_cp_async_wait<SMALLN_STAGES - 2>();   // this thread's copies of chunk t have landed ...
      __syncthreads();                         // ... and everybody else's; everybody is also done with chunk t - 1:
      issue_cpasync(t + SMALLN_STAGES - 1);    // its stage takes chunk t + STAGES - 1 (one barrier per chunk)
    } else
      mbar_wait(&full[s], (uint32_t)((t / SMALLN_STAGES) & 1));
    const double* ts = tile + ((long long)s * SMALLN_ROWS + warp * R) * SMALLN_CW + 2 * lane;
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const double2 w = *reinterpret_cast<const double2*>(ts + r * SMALLN_CW + 64 * u);
#pragma unroll
        for (int i = 0; i < NC; ++i) {
          acc[r][i] = fma(w.x, kv[u][i].x, acc[r][i]);
          acc[r][i] = fma(w.y, kv[u][i].y, acc[r][i]);
        }
      }
    if (!GPO_SMALLN_CPASYNC) {
      __syncthreads();   // every warp is done with stage s: refill it
      if (warp == 0 && t + SMALLN_STAGES < nchunks) issue(t + SMALLN_STAGES);
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int i = 0; i < NC; ++i)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[r][i] += __shfl_xor_sync(0xffffffffu, acc[r][i], o);
  // lane i < NC finishes candidate i over this warp's R rows
  double s0 = 0.0, h0 = 0.0, pq[DMAX];
#pragma unroll
  for (int q = 0; q < DMAX; ++q) pq[q] = 0.0;
  if (lane < NC) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      double b = 0.0;
#pragma unroll
      for (int i = 0; i < NC; ++i)
        if (lane == i) b = acc[r][i];
      if (SQ) { s0 = fma(b, b, s0); continue; }
      const int j = row0 + r;
      const double kj = __ldg(Kt + (long long)lane * n_pad + j);
      const double bh = p.Ht ? b * __ldg(p.Ht + (long long)z * p.kt_z + (long long)lane * n_pad + j) : -(b * kj);
      s0 = fma(b, kj, s0);
      h0 += bh;
#pragma unroll
      for (int q = 0; q < DMAX; ++q)
        if (q < d) pq[q] = fma(bh, __ldg(p.Xtc + (long long)q * p.ldx + j), pq[q]);
    }
    red[warp][0][lane] = s0;
    if (!SQ) red[warp][1][lane] = h0;
#pragma unroll
    for (int q = 0; q < DMAX; ++q)
      if (q < d) red[warp][2 + q][lane] = pq[q];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < (d + 2) * NC; idx += 256) {
    const int v = idx / NC, i = idx - v * NC;
    double t = 0.0;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) t += red[w8][v][i];
    p.partial[(long long)z * p.part_z + ((long long)rb * (d + 2) + v) * p.ldp + i] = t;
  }
  // the last CTA of this hyper-parameter set to arrive adds the partial rows (fixed order, so the sum is deterministic)
  // and leaves ONE row for K3: a single K3 thread would otherwise walk row_blocks x (d+2) dependent loads per candidate
  __shared__ int is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int t = atomicAdd(p.ticket + z, 1);
    is_last = (t == (int)gridDim.x - 1);
    if (is_last) p.ticket[z] = 0;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // one warp per (value, candidate) pair: the lanes fetch every 32nd partial row at once and a fixed shuffle tree adds them
  for (int idx = warp; idx < (d + 2) * NC; idx += 8) {
    const int v = idx / NC, i = idx - v * NC;
    const double* src = p.partial + (long long)z * p.part_z + (long long)v * p.ldp + i;
    double t = 0.0;
    for (int b0 = lane; b0 < (int)gridDim.x; b0 += 256) {   // (loads first, then the adds in the same order as before)
      double vals[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) vals[u] = b0 + 32 * u < (int)gridDim.x ? __ldcg(src + (long long)(b0 + 32 * u) * (d + 2) * p.ldp) : 0.0;
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (b0 + 32 * u < (int)gridDim.x) t += vals[u];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) p.final_rows[((long long)z * (d + 2) + v) * p.final_ld + i] = t;
  }
}


template <int NC, bool SQ, int SMALLN_ROWS>
__global__ void __launch_bounds__(256) smalln_wk_kernel(const SmallNParams p) {
  smalln_wk_body<NC, SQ, SMALLN_ROWS>(p);
}

// Both forms in ONE launch (gradient calls in the reference variance form need b = W^-1 k AND |L^-1 k|^2): grid.z = 0 streams
// W^-1, grid.z = 1 streams L^-1, concurrently -- alone each of them is latency-bound (2.1 / 1.4 TB/s at n = 4096: 63 + 49 us back
// to back), together they keep twice as many TMA stages in flight.
template <int NC, int SMALLN_ROWS>
__global__ void __launch_bounds__(256) smalln_both_kernel(const SmallNParams pw, const SmallNParams psq) {
  if (blockIdx.z == 0) smalln_wk_body<NC, false, SMALLN_ROWS>(pw);
  else smalln_wk_body<NC, true, SMALLN_ROWS>(psq);
}

// K1 for the same handful of candidates: one thread per TRAINING point (K1 proper gives each candidate a thread and
// would run 128-candidate blocks that are all padding but one row).  Block `split` covers K1's slice `split` of the
// training index and leaves the same per-slice partial sums in mean / dmean, K*^T / H*^T rows in Kt / Ht, for the
// candidate slots 0 .. nc_run-1 (slots past the last candidate shadow it, as K1's padding does).
constexpr int SMALLN_SLOTS = 32;   // candidate slots of a latency-path call (= SMALLN_MAX_CALL on the host side)
constexpr int COV_SUB = 4;    // CTAs per slice of the training index (grid.x = js * COV_SUB)
__global__ void __launch_bounds__(256) cov_small_kernel(const CovParams p) {
  const int z = blockIdx.y, split = blockIdx.x / COV_SUB, ss = blockIdx.x % COV_SUB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int d = p.d;
  const double* invl = p.invl + (long long)z * d;
  const double sf2 = p.sigf2[z];
  const double* alpha = p.alpha + (long long)z * p.n_pad;
  const int j_begin = split * p.tiles_per_split * COV_JT, j_end = min(p.n_pad, j_begin + p.tiles_per_split * COV_JT);
  // this CTA's quarter of the slice (whole 32-point groups)
  const int sub_len = (((j_end - j_begin) / COV_SUB + 31) / 32) * 32;
  const int jb = j_begin + ss * sub_len, je = min(j_end, jb + sub_len);
  const long long zs = (long long)z * p.js + split;
  // One WARP per candidate slot, the eight warps of the block work on eight slots at once, and COV_SUB CTAs share a slice (the
  // first version walked the slots one after the other, 256 points each, with two block barriers per slot: 28 us for eight slots,
  // all of it latency).  A slot's partial sums are formed by its warp alone -- lane-strided, then a fixed shuffle tree -- and are
  // parked in columns i + 32 ss of the slice's row; they do not depend on the other slots of the call.
  for (int i = warp; i < (int)p.nc_run; i += 8) {
    const long long cand = i < p.n_cand ? i : p.n_cand - 1;
    double xs[DMAX], m = 0.0, dm[DMAX];
#pragma unroll
    for (int q = 0; q < DMAX; ++q) {
      xs[q] = q < d ? p.Xs[cand * d + q] * invl[q] : 0.0;
      dm[q] = 0.0;
    }
    for (int j = jb + lane; j < je; j += 32) {
      double r2 = 0.0, df[DMAX];
#pragma unroll
      for (int q = 0; q < DMAX; ++q) {
        df[q] = q < d ? xs[q] - p.Xt[q * p.ldx + j] * invl[q] : 0.0;
        r2 = fma(df[q], df[q], r2);
      }
      double kv, hv;
      kern_k_h(p.kern, sf2, r2, kv, hv);
      if (j >= p.n) { kv = 0.0; hv = 0.0; }
      const double aj = alpha[j];
      m = fma(kv, aj, m);
      const double w = aj * hv;
#pragma unroll
      for (int q = 0; q < DMAX; ++q) dm[q] = fma(w, df[q], dm[q]);
      if (p.Kt) p.Kt[((long long)z * p.nc_ld + i) * p.n_pad + j] = kv;
      if (p.Ht) p.Ht[((long long)z * p.nc_ld + i) * p.n_pad + j] = hv;
    }
    const int col = i + SMALLN_SLOTS * ss;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m += __shfl_xor_sync(0xffffffffu, m, o);
    if (lane == 0) p.mean[zs * p.nc_ld + col] = m;
    if (p.dmean) {
#pragma unroll
      for (int q = 0; q < DMAX; ++q)
        if (q < d) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) dm[q] += __shfl_xor_sync(0xffffffffu, dm[q], o);
          if (lane == 0) p.dmean[(zs * d + q) * p.nc_ld + col] = dm[q] * invl[q];
        }
    }
  }
  // the last block of this hyper-parameter set adds the js * COV_SUB partial sums (fixed order) into column i of slice 0: K3 then
  // reads one value per candidate and output instead of walking dependent loads for each (41 us for a single point at n=4096)
  __shared__ int is_last;
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const int t = atomicAdd(p.ticket + z, 1);
    is_last = (t == (int)gridDim.x - 1);
    if (is_last) p.ticket[z] = 0;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // one warp per (value, slot) pair: the lanes fetch the partial sums (entry e = slice * COV_SUB + quarter on lane e % 32; at most
  // 16 slices -> two entries per lane) in one round trip, a fixed shuffle tree adds them
  const int nv = p.dmean ? d + 1 : 1, n_ent = p.js * COV_SUB;
  for (int idx = warp; idx < nv * (int)p.nc_run; idx += 8) {
    const int v = idx / (int)p.nc_run, i = idx - v * (int)p.nc_run;
    const double* src = v == 0 ? p.mean + (long long)z * p.js * p.nc_ld + i : p.dmean + ((long long)z * p.js * d + (v - 1)) * p.nc_ld + i;
    const long long stride = v == 0 ? p.nc_ld : (long long)d * p.nc_ld;
    const int e0 = lane, e1 = lane + 32;
    const double v0 = e0 < n_ent ? __ldcg(src + (long long)(e0 / COV_SUB) * stride + SMALLN_SLOTS * (e0 % COV_SUB)) : 0.0;
    const double v1 = e1 < n_ent ? __ldcg(src + (long long)(e1 / COV_SUB) * stride + SMALLN_SLOTS * (e1 % COV_SUB)) : 0.0;
    double t = v0 + v1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    __syncwarp();
    if (lane == 0) {
      if (v == 0) p.mean[(long long)z * p.js * p.nc_ld + i] = t;
      else p.dmean[((long long)z * p.js * d + (v - 1)) * p.nc_ld + i] = t;
    }
  }
}

// =====================================================================================================
// Fused value sweep for small models (n <= 128: every early BO iteration, BASELINE configs 1 and 2).
// K1 + the variance contraction in ONE kernel: the whole posterior state -- L^-1 (<= 132 KB), the length-scaled training
// inputs and alpha -- is resident in shared memory, candidates stream through, and K* never leaves the registers:
//   * a warp owns 8 NF candidates; lane (g, t) evaluates k(x_train[4 kk + t], x_cand[8 nf + g]) directly in the
//     B-fragment position of mma.m8n8k4 (no transposition, no K*^T round trip through HBM as on the general path)
//   * t = L^-1 k runs on the FP64 tensor pipe (DMMA) against A fragments read from the resident L^-1, skipping the
//     8 x 4 blocks above the diagonal; the model is padded to a multiple of 8 rows (not 128: at n = 55 the general
//     path does (128/55)^2 = 5.4x the necessary flops)
//   * mean = k . alpha is reduced in registers; the epilogue leaves mean and |L^-1 k|^2 per candidate, which K3
//     (acq_kernel) turns into s / EI / MPI / LCB exactly as it does for the general path.
// Reference arithmetic replaced: GPy kern.K + PosteriorExact._raw_predict (dtrtrs + sum of squares) as called from
// /root/reference/GPyOpt/models/gpmodel.py:98 and acquisitions/LCB.py:39, EI.py:36, MPI.py:36.
// Every candidate's result is independent of its neighbours and of the chunking (bitwise).
// =====================================================================================================
struct FusedSmallParams {
  const double* Xs;       // candidates AoS [n_cand][d] of this chunk
  long long n_cand;
  const double* Linv;     // [S][n_pad][n_pad] lower triangular (identity on the padding)
  const double* Xt;       // training inputs SoA [d][ldx]
  long long ldx;
  const double* alpha;    // [S][n_pad]
  const double* invl;     // [S][d]
  const double* sigf2;    // [S]
  int n, n_pad, d, kern;
  double* mean;           // [S][ld]
  double* sumsq;          // [S][ld]   |L^-1 k|^2
  long long ld;
};

constexpr int FUSED_THREADS = 256;
__host__ __device__ constexpr int fused_small_smem_bytes(int mf, int nf, int d) {
  const int np = 8 * mf;
  return (np * (np + 4) + d * np + np + (FUSED_THREADS / 32) * (8 * nf * (d + 1) + 2 * 8 * nf)) * 8;
}

template <int MF, int NF>
__global__ void __launch_bounds__(FUSED_THREADS) fused_small_value_kernel(const FusedSmallParams p) {
  constexpr int NP = 8 * MF, PA = NP + 4, CW = 8 * NF;
  extern __shared__ double fs_smem[];
  const int d = p.d, PW = d + 1;
  double* Ls = fs_smem;                         // [NP][PA]   L^-1
  double* xts = Ls + NP * PA;                   // [d][NP]    training inputs / lengthscale
  double* als = xts + d * NP;                   // [NP]       alpha
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  double* xsw = als + NP + warp * (CW * PW + 2 * CW);   // [CW][PW] this warp's candidates / lengthscale
  double* outw = xsw + CW * PW;                 // [2][CW]    staging of the warp's results
  const int z = blockIdx.y;
  const double* invl = p.invl + (long long)z * d;
  const double sf2 = p.sigf2[z];
  {
    const double* Lg = p.Linv + (long long)z * p.n_pad * p.n_pad;
    for (int i = tid; i < NP * NP; i += FUSED_THREADS) {
      const int r = i / NP, c = i - r * NP;
      Ls[r * PA + c] = Lg[(long long)r * p.n_pad + c];
    }
    for (int i = tid; i < d * NP; i += FUSED_THREADS) {
      const int q = i / NP, j = i - q * NP;
      xts[i] = p.Xt[q * p.ldx + j] * invl[q];
    }
    for (int i = tid; i < NP; i += FUSED_THREADS) als[i] = p.alpha[(long long)z * p.n_pad + i];
  }
  __syncthreads();
  const long long n_blocks = (p.n_cand + CW - 1) / CW;
  for (long long blk = (long long)blockIdx.x * (FUSED_THREADS / 32) + warp; blk < n_blocks; blk += (long long)gridDim.x * (FUSED_THREADS / 32)) {
    const long long c0 = blk * CW;
    __syncwarp();
    for (int i = lane; i < CW * d; i += 32) {
      const int c = i / d, q = i - c * d;
      const long long cand = min(c0 + c, p.n_cand - 1);      // rows past the end shadow the last one (never stored)
      xsw[c * PW + q] = p.Xs[cand * d + q] * invl[q];
    }
    __syncwarp();
    double acc[MF][NF][2], macc[NF];
#pragma unroll
    for (int mf = 0; mf < MF; ++mf)
#pragma unroll
      for (int nf = 0; nf < NF; ++nf) acc[mf][nf][0] = acc[mf][nf][1] = 0.0;
#pragma unroll
    for (int nf = 0; nf < NF; ++nf) macc[nf] = 0.0;
#pragma unroll 1
    for (int kp = 0; kp < MF; ++kp) {
      // the lane's two training points of this pair of k steps (j = 8 kp + 4 hh + t) against its NF candidates: all 2 NF
      // covariance values are evaluated in one straight-line block (independent chains interleave), then fed to the DMMAs
      double r2[2 * NF], kv[2 * NF];
#pragma unroll
      for (int v = 0; v < 2 * NF; ++v) r2[v] = 0.0;
      const int j0 = 8 * kp + t;
      for (int q = 0; q < d; ++q) {
        const double xj0 = xts[q * NP + j0], xj1 = xts[q * NP + j0 + 4];
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) {
          const double xc = xsw[(8 * nf + g) * PW + q];
          const double d0 = xc - xj0, d1 = xc - xj1;
          r2[nf] = fma(d0, d0, r2[nf]);
          r2[NF + nf] = fma(d1, d1, r2[NF + nf]);
        }
      }
      kern_k_vec<2 * NF>(p.kern, sf2, r2, kv);
      const double a0 = als[j0], a1 = als[j0 + 4];
#pragma unroll
      for (int nf = 0; nf < NF; ++nf) {
        kv[nf] = j0 < p.n ? kv[nf] : 0.0;               // padding rows of the model carry no covariance
        kv[NF + nf] = j0 + 4 < p.n ? kv[NF + nf] : 0.0;
        macc[nf] = fma(kv[nf], a0, macc[nf]);
        macc[nf] = fma(kv[NF + nf], a1, macc[nf]);
      }
      // rows 8 mf .. 8 mf + 7 of L^-1 against k columns 8 kp + 4 hh .. + 3: zero above the diagonal (mf < kp)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
        for (int mf = 0; mf < MF; ++mf) {
          if (mf >= kp) {   // warp-uniform
            const double a = Ls[(8 * mf + g) * PA + j0 + 4 * hh];
#pragma unroll
            for (int nf = 0; nf < NF; ++nf) dmma_884(acc[mf][nf][0], acc[mf][nf][1], a, kv[hh * NF + nf]);
          }
        }
      }
    }
    // epilogue: lane (g, t) holds rows 8 mf + g of candidates 8 nf + 2 t + e; mean partials are per (j = .. + t, candidate 8 nf + g)
#pragma unroll
    for (int nf = 0; nf < NF; ++nf) {
      double mv = macc[nf];
      mv += __shfl_xor_sync(0xffffffffu, mv, 1);
      mv += __shfl_xor_sync(0xffffffffu, mv, 2);
      if (t == 0) outw[8 * nf + g] = mv;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        double s2 = 0.0;
#pragma unroll
        for (int mf = 0; mf < MF; ++mf) s2 = fma(acc[mf][nf][e], acc[mf][nf][e], s2);
        s2 += __shfl_xor_sync(0xffffffffu, s2, 4);
        s2 += __shfl_xor_sync(0xffffffffu, s2, 8);
        s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
        if (g == 0) outw[CW + 8 * nf + 2 * t + e] = s2;
      }
    }
    __syncwarp();
    if (lane < CW && c0 + lane < p.n_cand) {
      p.mean[(long long)z * p.ld + c0 + lane] = outw[lane];
      p.sumsq[(long long)z * p.ld + c0 + lane] = outw[CW + lane];
    }
  }
}

// =====================================================================================================
// Batched multi-start L-BFGS, lock-step on the device.
// The reference optimises its anchor points one after the other with SciPy's L-BFGS-B, one f(x) and one f_df(x) host call
// per step (GPyOpt/optimization/optimizer.py:28-61, acquisition_optimizer.py:73-74).  Here all M anchors advance together:
// per round ONE fused evaluation of value + gradient at the M trial points (the latency kernels, N = M) and ONE launch of
// lbfgs_update_kernel, in which thread a runs anchor a's optimiser state machine:
//   * direction: two-loop recursion over a ring of LB_HIST curvature pairs, scaled by y.y / s.y, on the free variables
//     (a variable on a bound whose gradient -- or quasi-Newton direction -- points outward is frozen)
//   * line search: the More-Thuente search of L-BFGS-B (MINPACK-2 dcsrch / dcstep in reverse-communication form, which is
//     what makes it a state machine: ftol 1e-3, gtol 0.9, xtol 0.1, at most 20 evaluations, step capped so that the
//     segment stays inside the box)
//   * L-BFGS-B's update and stopping rules: pair skipped unless s.y > eps * (-g.s), projected-gradient norm <= pgtol,
//     relative decrease <= factr * eps, maxiter; a failed search drops the history and restarts from steepest descent.
// In the interior of the box the iterates follow SciPy's L-BFGS-B (same search, same update); at active bounds L-BFGS-B's
// Cauchy-point / subspace step is replaced by the projected direction above.  No host round trip inside the loop: the
// host only polls the finished-anchor count every few rounds.  (The test-suite keeps a line-by-line NumPy mirror of this state
// machine and walks both along the same path.)
// =====================================================================================================
__device__ __forceinline__ double lb_clamp(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

__device__ __forceinline__ double lb_cubic_gamma(double theta, double da, double db) {
  const double s = ::fmax(fabs(theta), ::fmax(fabs(da), fabs(db)));
  return s * sqrt(::fmax(0.0, (theta / s) * (theta / s) - (da / s) * (db / s)));
}

// MINPACK-2 dcstep: safeguarded cubic / quadratic step and update of the interval of uncertainty
__device__ double lb_dcstep(double& stx, double& fx, double& dx, double& sty, double& fy, double& dy, double stp, double fp,
                            double dp, int& brackt, double stpmin, double stpmax) {
  const double sgnd = dp * (dx / fabs(dx));
  double stpf;
  if (fp > fx) {
    const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    double gamma = lb_cubic_gamma(theta, dx, dp);
    if (stp < stx) gamma = -gamma;
    const double p = (gamma - dx) + theta, q = ((gamma - dx) + gamma) + dp, r = p / q;
    const double stpc = stx + r * (stp - stx);
    const double stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx);
    stpf = fabs(stpc - stx) < fabs(stpq - stx) ? stpc : stpc + (stpq - stpc) / 2.0;
    brackt = 1;
  } else if (sgnd < 0.0) {
    const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    double gamma = lb_cubic_gamma(theta, dx, dp);
    if (stp > stx) gamma = -gamma;
    const double p = (gamma - dp) + theta, q = ((gamma - dp) + gamma) + dx, r = p / q;
    const double stpc = stp + r * (stx - stp);
    const double stpq = stp + (dp / (dp - dx)) * (stx - stp);
    stpf = fabs(stpc - stp) > fabs(stpq - stp) ? stpc : stpq;
    brackt = 1;
  } else if (fabs(dp) < fabs(dx)) {
    const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    double gamma = lb_cubic_gamma(theta, dx, dp);
    if (stp > stx) gamma = -gamma;
    const double p = (gamma - dp) + theta, q = (gamma + (dx - dp)) + gamma, r = p / q;
    double stpc;
    if (r < 0.0 && gamma != 0.0) stpc = stp + r * (stx - stp);
    else if (stp > stx) stpc = stpmax;
    else stpc = stpmin;
    const double stpq = stp + (dp / (dp - dx)) * (stx - stp);
    if (brackt) {
      stpf = fabs(stpc - stp) < fabs(stpq - stp) ? stpc : stpq;
      stpf = stp > stx ? ::fmin(stp + 0.66 * (sty - stp), stpf) : ::fmax(stp + 0.66 * (sty - stp), stpf);
    } else {
      stpf = fabs(stpc - stp) > fabs(stpq - stp) ? stpc : stpq;
      stpf = ::fmax(stpmin, ::fmin(stpmax, stpf));
    }
  } else {
    if (brackt) {
      const double theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp;
      double gamma = lb_cubic_gamma(theta, dy, dp);
      if (stp > sty) gamma = -gamma;
      const double p = (gamma - dp) + theta, q = ((gamma - dp) + gamma) + dy, r = p / q;
      stpf = stp + r * (sty - stp);
    } else if (stp > stx) stpf = stpmax;
    else stpf = stpmin;
  }
  if (fp > fx) { sty = stp; fy = fp; dy = dp; }
  else {
    if (sgnd < 0.0) { sty = stx; fy = fx; dy = dx; }
    stx = stp; fx = fp; dx = dp;
  }
  return stpf;
}

__device__ void lb_dcsrch_start(LbLineSearch& st, double stp, double f, double g, double stpmax) {
  st.brackt = 0; st.stage = 1; st.finit = f; st.ginit = g; st.gtest = 1e-3 * g; st.width = stpmax; st.width1 = 2.0 * stpmax;
  st.stx = 0.0; st.fx = f; st.gx = g; st.sty = 0.0; st.fy = f; st.gy = g; st.stmin = 0.0; st.stmax = stp + 4.0 * stp;
  st.stpmx = stpmax; st.nls = 0;
}

// one reverse-communication step of dcsrch after the evaluation (f, g) at step stp; returns true when the search ends
// (convergence or one of dcsrch's warnings), otherwise stp holds the next trial step
__device__ bool lb_dcsrch_next(LbLineSearch& st, double& stp, double f, double g) {
  const double stpmax = st.stpmx;
  const double ftest = st.finit + stp * st.gtest;
  if (st.stage == 1 && f <= ftest && g >= 0.0) st.stage = 2;
  bool warn = false;
  if (st.brackt && (stp <= st.stmin || stp >= st.stmax)) warn = true;
  if (st.brackt && st.stmax - st.stmin <= 0.1 * st.stmax) warn = true;
  if (stp == stpmax && f <= ftest && g <= st.gtest) warn = true;
  if (stp == 0.0 && (f > ftest || g >= st.gtest)) warn = true;
  if (f <= ftest && fabs(g) <= 0.9 * (-st.ginit)) return true;
  if (warn) return true;
  if (st.stage == 1 && f <= st.fx && f > ftest) {
    const double gt = st.gtest;
    double fxm = st.fx - st.stx * gt, fym = st.fy - st.sty * gt, gxm = st.gx - gt, gym = st.gy - gt;
    stp = lb_dcstep(st.stx, fxm, gxm, st.sty, fym, gym, stp, f - stp * gt, g - gt, st.brackt, st.stmin, st.stmax);
    st.fx = fxm + st.stx * gt; st.fy = fym + st.sty * gt; st.gx = gxm + gt; st.gy = gym + gt;
  } else {
    stp = lb_dcstep(st.stx, st.fx, st.gx, st.sty, st.fy, st.gy, stp, f, g, st.brackt, st.stmin, st.stmax);
  }
  if (st.brackt) {
    if (fabs(st.sty - st.stx) >= 0.66 * st.width1) stp = st.stx + 0.5 * (st.sty - st.stx);
    st.width1 = st.width;
    st.width = fabs(st.sty - st.stx);
    st.stmin = ::fmin(st.stx, st.sty); st.stmax = ::fmax(st.stx, st.sty);
  } else {
    st.stmin = stp + 1.1 * (stp - st.stx);
    st.stmax = stp + 4.0 * (stp - st.stx);
  }
  stp = ::fmin(::fmax(stp, 0.0), stpmax);
  if ((st.brackt && (stp <= st.stmin || stp >= st.stmax)) || (st.brackt && st.stmax - st.stmin <= 0.1 * st.stmax)) stp = st.stx;
  return false;
}

__global__ void lbfgs_init_kernel(const LbfgsParams p, const double* __restrict__ x0) {
  const int a = threadIdx.x;
  if (a == 0) *p.n_done = 0;
  if (a >= p.M) return;
  for (int q = 0; q < p.d; ++q) {
    const double v = lb_clamp(x0[a * p.d + q], p.lo[q], p.hi[q]);
    p.x[a * p.d + q] = v;
    p.xt[a * p.d + q] = v;
  }
  p.hist_n[a] = 0; p.hist_head[a] = 0; p.iter[a] = 0; p.phase[a] = 0; p.done[a] = 0; p.nfev[a] = 0;
  p.step[a] = 0.0; p.gd[a] = 0.0; p.f[a] = 0.0;
}

// One optimiser state-machine step of anchor `a` after an evaluation at its trial point, executed by ONE WARP: the lanes stage
// ALL of the anchor's state -- iterate, gradient, direction, curvature pairs, box, the evaluation's outputs, the line-search record
// and the scalar counters (<= 3.5 KB) -- from global into shared memory in one round trip, lane 0 runs the scalar state machine on
// the shared copy (every global access of the state machine used to be its own ~0.5 us round trip: 20 - 55 us per round), the lanes
// write the state back.
// dot product of two shared-memory vectors added in coordinate order with fused multiply-adds (every lane forms the same value
// from broadcast reads).  Deliberately a rolled loop: these one-warp kernels run from a cold instruction cache, and what they cost
// is the number of distinct instruction lines they touch (an unrolled, shuffle-based form of this function made one pass of the
// two-loop recursion 100,000 cycles long).
__device__ __noinline__ double lb_dot_seq(const double* a, const double* b, int d) {
  // (one shared copy of this body serves all 23 dot products of a pass: the operands are fetched up front, then the chain runs)
  double av[DMAX], bv[DMAX];
#pragma unroll
  for (int q = 0; q < DMAX; ++q) {
    av[q] = q < d ? a[q] : 0.0;
    bv[q] = q < d ? b[q] : 0.0;
  }
  double s = 0.0;
#pragma unroll
  for (int q = 0; q < DMAX; ++q)
    if (q < d) s = fma(av[q], bv[q], s);
  return s;
}

struct LbScalars { double ft, f, step, gd; int hist_n, hist_head, iter, phase, nfev, done, pad0, pad1; };
// issue the loads of anchor a's state (everything but this round's evaluation outputs); nothing here waits for them
__device__ __forceinline__ void lbfgs_stage_issue(const LbfgsParams& p, int a, int lane, LbStage& st) {
  const int d = p.d;
  st.done = (a < p.M) ? p.done[a] : 1;
  const int q = lane < d ? lane : 0;
  st.x = p.x[a * d + q]; st.g = p.g[a * d + q]; st.dir = p.dir[a * d + q]; st.lo = p.lo[q]; st.hi = p.hi[q]; st.xt = p.xt[a * d + q];
  const double* gS = p.S + (long long)a * LB_HIST * d; const double* gY = p.Y + (long long)a * LB_HIST * d;
#pragma unroll
  for (int r = 0; r < 5; ++r) {
    const int e = lane + 32 * r < LB_HIST * d ? lane + 32 * r : 0;
    st.S[r] = gS[e]; st.Y[r] = gY[e];
  }
  st.rho = p.rho[a * LB_HIST + (lane < LB_HIST ? lane : 0)];
  st.lsw = reinterpret_cast<const double*>(p.ls + a)[lane & 15];
  const double* dp = lane == 1 ? p.f + a : lane == 2 ? p.step + a : p.gd + a;
  const int* ip = lane == 4 ? p.hist_n + a : lane == 5 ? p.hist_head + a : lane == 6 ? p.iter + a : lane == 7 ? p.phase + a : p.nfev + a;
  st.dsc = *dp;
  st.isc = *ip;
}

__device__ void lbfgs_anchor_update(const LbfgsParams& p, int a, int lane, const LbStage& st) {
  __shared__ double sh_x[DMAX], sh_g[DMAX], sh_dir[DMAX], sh_S[LB_HIST * DMAX], sh_Y[LB_HIST * DMAX], sh_rho[LB_HIST];
  __shared__ double sh_lo[DMAX], sh_hi[DMAX], sh_G[DMAX], sh_xt[DMAX];
  __shared__ LbLineSearch sh_ls;
  __shared__ LbScalars sh_sc;
  __shared__ double sh_v[DMAX], sh_al[LB_HIST];
  __shared__ int sh_fl[8];   // finished, need_direction, steepest, hist_n, hist_head: lane 0 -> warp -> lane 0
  static_assert(sizeof(LbLineSearch) == 16 * 8, "copied as 16 words");
#ifdef GPO_ACQ_TIMING
  long long lt[6]; int n_pass = 0; lt[0] = clock64();
#endif
  if (a >= p.M || st.done) return;              // (warp-uniform)
  const int d = p.d;
  {
    if (lane < d) {
      sh_x[lane] = st.x; sh_g[lane] = st.g; sh_dir[lane] = st.dir; sh_lo[lane] = st.lo; sh_hi[lane] = st.hi; sh_xt[lane] = st.xt;
      sh_G[lane] = p.gt[a * d + lane];          // this round's evaluation: written by this warp a moment ago (or by the sweep)
    }
#pragma unroll
    for (int r = 0; r < 5; ++r)
      if (lane + 32 * r < LB_HIST * d) { sh_S[lane + 32 * r] = st.S[r]; sh_Y[lane + 32 * r] = st.Y[r]; }
    if (lane < LB_HIST) sh_rho[lane] = st.rho;
    if (lane >= 16) reinterpret_cast<double*>(&sh_ls)[lane - 16] = st.lsw;
    if (lane == 0) sh_sc.ft = p.ft[a];
    else if (lane < 4) (&sh_sc.ft)[lane] = st.dsc;
    else if (lane < 9) (&sh_sc.hist_n)[lane - 4] = st.isc;
    else if (lane == 9) sh_sc.done = 0;
  }
  const double pgtol = p.pgtol, ftol = p.ftol;
  const int maxiter = p.maxiter, negate = p.negate;
  __syncwarp();
#ifdef GPO_ACQ_TIMING
  lt[1] = clock64();
#endif
  if (lane == 0) {
  const double eps = 2.220446049250313e-16;
  double* x = sh_x; double* g = sh_g; double* dir = sh_dir; double* xt = sh_xt;
  const double* lo = sh_lo; const double* hi = sh_hi;
  LbLineSearch ls = sh_ls;
  LbScalars sc = sh_sc;
  const double sgn = negate ? -1.0 : 1.0;
  double F = sgn * sc.ft;
  double G[DMAX];
  bool finite = F == F && fabs(F) < 1e300;
  _Pragma("unroll 1") for (int q = 0; q < d; ++q) { G[q] = sgn * sh_G[q]; finite = finite && G[q] == G[q] && fabs(G[q]) < 1e300; }
  sc.nfev += 1;
  int hist_n = sc.hist_n, hist_head = sc.hist_head;
  bool finished = false, need_direction = true;
  auto pg_norm = [&](const double* xx, const double* gg) {
    double m = 0.0;
    _Pragma("unroll 1") for (int q = 0; q < d; ++q) m = ::fmax(m, fabs(lb_clamp(xx[q] - gg[q], lo[q], hi[q]) - xx[q]));
    return m;
  };
  bool steepest = false;
  if (sc.phase == 0) {            // the evaluation at the start point
    sc.f = F;
    if (!finite) finished = true;
    else {
      _Pragma("unroll 1") for (int q = 0; q < d; ++q) g[q] = G[q];
      if (pg_norm(x, g) <= pgtol) finished = true;
      steepest = true;
      sc.phase = 1;
    }
  } else {                          // a line-search trial at xt = x + step * dir
    double stp = sc.step;
    double gdt = 0.0;
    if (finite) { for (int q = 0; q < d; ++q) gdt = fma(G[q], dir[q], gdt); }
    else { F = ls.fx + 1e10 * (1.0 + fabs(ls.fx)); gdt = 1e10; }
    ls.nls += 1;
    const bool ended = lb_dcsrch_next(ls, stp, F, gdt);
    if (!ended && ls.nls < 20) {    // next trial
      sc.step = stp;
      _Pragma("unroll 1") for (int q = 0; q < d; ++q) xt[q] = lb_clamp(x[q] + stp * dir[q], lo[q], hi[q]);
      need_direction = false;
    } else if (ended && finite) {
      // accept: curvature pair, new iterate, stopping rules (as scipy.optimize.fmin_l_bfgs_b: pgtol, factr, maxiter)
      double sy = 0.0;
      _Pragma("unroll 1") for (int q = 0; q < d; ++q) sy = fma(xt[q] - x[q], G[q] - g[q], sy);
      if (sy > eps * (-sc.gd * sc.step)) {
        double* Sh = sh_S + hist_head * d; double* Yh = sh_Y + hist_head * d;
        _Pragma("unroll 1") for (int q = 0; q < d; ++q) { Sh[q] = xt[q] - x[q]; Yh[q] = G[q] - g[q]; }
        sh_rho[hist_head] = 1.0 / sy;
        hist_head = (hist_head + 1) % LB_HIST;
        if (hist_n < LB_HIST) hist_n += 1;
      }
      const double f0 = sc.f;
      _Pragma("unroll 1") for (int q = 0; q < d; ++q) { x[q] = xt[q]; g[q] = G[q]; }
      sc.f = F;
      sc.iter += 1;
      if (pg_norm(x, g) <= pgtol || (f0 - F) <= ftol * ::fmax(::fmax(fabs(f0), fabs(F)), 1.0) || sc.iter >= maxiter)
        finished = true;
    } else {
      // failed search: with history, drop it and restart from steepest descent at the iterate; without, stop
      if (hist_n > 0) { hist_n = 0; steepest = true; }
      else finished = true;
    }
  }
  sh_fl[0] = finished; sh_fl[1] = need_direction; sh_fl[2] = steepest; sh_fl[3] = hist_n; sh_fl[4] = hist_head;
  sh_ls = ls; sh_sc = sc;
  }
  __syncwarp();
  // ---- new search direction from the iterate (x, g): the whole warp, lane q owns coordinate q.  The scalar form (lane 0 walking
  //      v[], S[], Y[] element by element out of local / shared memory) took 20 - 55 us per round; every dot product here adds its
  //      terms in the same order with the same fused multiply-adds (lb_dot_seq), so the direction has the same bits. ----
#ifdef GPO_ACQ_TIMING
  lt[2] = clock64();
#endif
  double gdd = 0.0, dn = 0.0;
  const bool want_dir = !sh_fl[0] && sh_fl[1];
  if (want_dir) {           // (warp-uniform)
    bool steepest = sh_fl[2] != 0;
    int hist_n = sh_fl[3];
    const int hist_head = sh_fl[4];
    const bool on = lane < d;
    const double xq = on ? sh_x[lane] : 0.0, gq = on ? sh_g[lane] : 0.0, loq = on ? sh_lo[lane] : 0.0, hiq = on ? sh_hi[lane] : 0.0;
    bool fr = on && !((xq <= loq && gq > 0.0) || (xq >= hiq && gq < 0.0));
    double dirq = 0.0;
    const int lq = on ? lane : 0;   // (lanes past d shadow coordinate 0 and never store)
    for (int attempt = 0; attempt < 2; ++attempt) {
      const int hn = steepest ? 0 : hist_n;
      for (int pass = 0; pass <= d; ++pass) {
#ifdef GPO_ACQ_TIMING
        ++n_pass;
#endif
        double v = fr ? gq : 0.0;
        if (on) sh_v[lane] = v;
        __syncwarp();
#pragma unroll 1
        for (int i = 0; i < hn; ++i) {        // newest to oldest
          const int h = (hist_head - 1 - i + 2 * LB_HIST) % LB_HIST;
          const double ali = sh_rho[h] * lb_dot_seq(sh_S + h * d, sh_v, d);
          if (lane == 0) sh_al[i] = ali;
          v = fma(-ali, sh_Y[h * d + lq], v);
          __syncwarp();
          if (on) sh_v[lane] = v;
          __syncwarp();
        }
        if (hn > 0) {
          const int h = (hist_head - 1 + LB_HIST) % LB_HIST;
          const double gamma = 1.0 / (sh_rho[h] * lb_dot_seq(sh_Y + h * d, sh_Y + h * d, d));
          v *= gamma;
          __syncwarp();
          if (on) sh_v[lane] = v;
          __syncwarp();
        }
#pragma unroll 1
        for (int i = hn - 1; i >= 0; --i) {   // oldest to newest
          const int h = (hist_head - 1 - i + 2 * LB_HIST) % LB_HIST;
          const double be = sh_rho[h] * lb_dot_seq(sh_Y + h * d, sh_v, d);
          v = fma(sh_al[i] - be, sh_S[h * d + lq], v);
          __syncwarp();
          if (on) sh_v[lane] = v;
          __syncwarp();
        }
        dirq = fr ? -v : 0.0;
        // a free variable sitting on a bound and pushed outward by the quasi-Newton coupling: freeze it and redo
        const bool hit = fr && ((xq <= loq && dirq < 0.0) || (xq >= hiq && dirq > 0.0));
        if (hit) fr = false;
        if (!__any_sync(0xffffffffu, hit)) break;
      }
      __syncwarp();
      if (on) sh_v[lane] = dirq;
      __syncwarp();
      gdd = lb_dot_seq(sh_g, sh_v, d);
      dn = lb_dot_seq(sh_v, sh_v, d);
      if (gdd < 0.0 || hn == 0) break;
      hist_n = 0;                             // not a descent direction: drop the history, steepest descent
      steepest = true;
    }
    if (on) sh_dir[lane] = dirq;
    if (lane == 0) sh_fl[3] = hist_n;
  }
  __syncwarp();
#ifdef GPO_ACQ_TIMING
  lt[3] = clock64();
#endif
  if (lane == 0) {
  double* x = sh_x; double* dir = sh_dir; double* xt = sh_xt;
  const double* lo = sh_lo; const double* hi = sh_hi;
  LbLineSearch ls = sh_ls;
  LbScalars sc = sh_sc;
  bool finished = sh_fl[0] != 0;
  const int hist_n = sh_fl[3], hist_head = sh_fl[4];
  if (want_dir) {
    if (!(dn > 0.0) || !(gdd < 0.0)) finished = true;
    else {
      double stpmx = 1e10;
      _Pragma("unroll 1") for (int q = 0; q < d; ++q) {
        if (dir[q] > 0.0) stpmx = ::fmin(stpmx, (hi[q] - x[q]) / dir[q]);
        else if (dir[q] < 0.0) stpmx = ::fmin(stpmx, (lo[q] - x[q]) / dir[q]);
      }
      const double stp = ::fmin(1.0, stpmx);
      lb_dcsrch_start(ls, stp, sc.f, gdd, stpmx);
      sc.step = stp;
      sc.gd = gdd;
      _Pragma("unroll 1") for (int q = 0; q < d; ++q) xt[q] = lb_clamp(x[q] + stp * dir[q], lo[q], hi[q]);
    }
  }
  if (finished) {
    sc.done = 1;
    _Pragma("unroll 1") for (int q = 0; q < d; ++q) xt[q] = x[q];   // later (lock-step) evaluations of this anchor are idle repeats
    atomicAdd(p.n_done, 1);
  }
  sh_ls = ls;
  sc.hist_n = hist_n; sc.hist_head = hist_head;
  sh_sc = sc;
  }
  __syncwarp();
#ifdef GPO_ACQ_TIMING
  lt[4] = clock64();
#endif
  {
    double* gx = p.x + a * d; double* gg = p.g + a * d; double* gd = p.dir + a * d;
    double* gS = p.S + (long long)a * LB_HIST * d; double* gY = p.Y + (long long)a * LB_HIST * d;
    for (int i = lane; i < d; i += 32) { gx[i] = sh_x[i]; gg[i] = sh_g[i]; gd[i] = sh_dir[i]; p.xt[a * d + i] = sh_xt[i]; }
    if (lane >= 16) reinterpret_cast<double*>(p.ls + a)[lane - 16] = reinterpret_cast<const double*>(&sh_ls)[lane - 16];
    {
      double* dp = lane == 1 ? p.f + a : lane == 2 ? p.step + a : p.gd + a;
      int* ip = lane == 4 ? p.hist_n + a : lane == 5 ? p.hist_head + a : lane == 6 ? p.iter + a : lane == 7 ? p.phase + a : p.nfev + a;
      if (lane >= 1 && lane < 4) *dp = (&sh_sc.ft)[lane];
      else if (lane >= 4 && lane < 9) *ip = (&sh_sc.hist_n)[lane - 4];
      else if (lane == 9 && sh_sc.done) p.done[a] = 1;
    }
    for (int i = lane; i < LB_HIST * d; i += 32) { gS[i] = sh_S[i]; gY[i] = sh_Y[i]; }
    if (lane < LB_HIST) p.rho[a * LB_HIST + lane] = sh_rho[lane];
  }
#ifdef GPO_ACQ_TIMING
  if (lane == 0 && a == 0) printf("  lbfgs: stage %lld A %lld dir %lld (passes %d) C %lld wb %lld\n", lt[1] - lt[0], lt[2] - lt[1], lt[3] - lt[2], n_pass, lt[4] - lt[3], clock64() - lt[4]);
#endif
}

// stand-alone form: one warp (block) per anchor
__global__ void __launch_bounds__(32) lbfgs_update_kernel(const LbfgsParams p) {
  LbStage st;
  lbfgs_stage_issue(p, (int)blockIdx.x, (int)threadIdx.x, st);
  lbfgs_anchor_update(p, (int)blockIdx.x, (int)threadIdx.x, st);
}

}  // namespace gpo
