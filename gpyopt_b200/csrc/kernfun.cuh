// Stationary covariance functions of the scaled squared distance r2 = sum_q ((x_q - x'_q)/l_q)^2.
// Mirrors GPy kern.src.{rbf.RBF, stationary.Matern32, stationary.Matern52}.K_of_r / dK_dr
//  as used through /root/reference/GPyOpt/models/gpmodel.py:58,221.
#pragma once

namespace gpo {

enum KernelKind { KERN_RBF = 0, KERN_MATERN52 = 1, KERN_MATERN32 = 2 };

// Taylor coefficients 1/13! ... 1/2! of exp_nonpos.  In constant memory on purpose: as literals every FP64 coefficient costs
// two UMOV instructions per use (ncu on the fused small-model kernel: as many UMOVs as DFMAs issued); as c[bank][offset]
// operands they ride along with the DFMA.
__constant__ double GPO_EXP_C[12] = {1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0,
                                     1.0 / 40320.0,      1.0 / 5040.0,      1.0 / 720.0,      1.0 / 120.0,     1.0 / 24.0,
                                     1.0 / 6.0,          0.5};

// exp(x) for x <= 0, branch-free (17 FP64 instructions, no divergent slow paths, so the compiler can interleave several
// evaluations -- K1 is bound by exactly this dependency chain).  Cody-Waite reduction x = n ln2 + r with the fdlibm
// ln2_hi / ln2_lo split, degree-13 Taylor polynomial on |r| <= ln2/2 (truncation 4e-18), 2^n applied to the exponent
// field.  Measured against long-double expl on 2e7 points in [-700, 0]: max error 0.87 ulp.  Results below
// exp(-708) = 3.3e-308 are flushed to zero.
__device__ __forceinline__ double exp_nonpos(double x) {
  const double SHIFT = 6755399441055744.0;  // 1.5 * 2^52: adding it leaves round-to-nearest(x log2 e) in the low word
  const double fn = fma(x, 1.4426950408889634, SHIFT);
  const int n = __double2loint(fn);
  const double nf = fn - SHIFT;
  double r = fma(nf, -6.93147180369123816490e-01, x);
  r = fma(nf, -1.90821492927058770002e-10, r);
  double p = GPO_EXP_C[0];
#pragma unroll
  for (int i = 1; i < 12; ++i) p = fma(p, r, GPO_EXP_C[i]);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const double res = __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
  return x < -708.0 ? 0.0 : res;
}

// sqrt(x) for x >= 0 without the slow-path branch of CUDA's sqrt() (BSSY / CALL around every evaluation: it splits the
// instruction stream into one basic block per evaluation, so independent evaluations are not interleaved).  Single-precision
// reciprocal-square-root seed, two Newton steps in double and one correction of the product x * y: within 1 ulp for
// 1e-30 <= x <= 1e30 (scaled squared distances); below that the result is 0 (|error| < 1e-15).
__device__ __forceinline__ double sqrt_nobranch(double x) {
  double y = (double)rsqrtf((float)x);
  y = y * fma(-0.5 * x, y * y, 1.5);
  y = y * fma(-0.5 * x, y * y, 1.5);
  double r = x * y;
  r = fma(fma(-r, r, x), 0.5 * y, r);
  return x < 1e-30 ? 0.0 : r;
}

// k = K_of_r(r),  h = dK_dr(r) / r   (the factor multiplying (x_q - x'_q)/l_q^2 in gradients_X; finite at r = 0,
// where GPy substitutes 1/r := 0 -- the coordinate difference is zero there, so the product agrees)
__device__ __forceinline__ void kern_k_h(int kind, double sf2, double r2, double& k, double& h) {
  if (kind == KERN_RBF) {
    k = sf2 * exp_nonpos(-0.5 * r2);
    h = -k;
  } else if (kind == KERN_MATERN52) {
    const double s5 = 2.23606797749978969641;
    const double r = sqrt_nobranch(r2);
    const double e = sf2 * exp_nonpos(-s5 * r);
    k = (1.0 + s5 * r + (5.0 / 3.0) * r2) * e;
    h = -(5.0 / 3.0) * (1.0 + s5 * r) * e;
  } else {
    const double s3 = 1.73205080756887729353;
    const double r = sqrt_nobranch(r2);
    const double e = sf2 * exp_nonpos(-s3 * r);
    k = (1.0 + s3 * r) * e;
    h = -3.0 * e;
  }
}

__device__ __forceinline__ double kern_k(int kind, double sf2, double r2) {
  double k, h;
  kern_k_h(kind, sf2, r2, k, h);
  return k;
}

// NV (k, h) pairs at once with the (warp-uniform) kernel switch outside the loop: the same arithmetic as kern_k_h, element by
// element (bitwise), in one straight-line block whose independent chains the compiler interleaves
template <int NV>
__device__ __forceinline__ void kern_k_h_vec(int kind, double sf2, const double (&r2)[NV], double (&k)[NV], double (&h)[NV]) {
  if (kind == KERN_RBF) {
#pragma unroll
    for (int v = 0; v < NV; ++v) { k[v] = sf2 * exp_nonpos(-0.5 * r2[v]); h[v] = -k[v]; }
  } else if (kind == KERN_MATERN52) {
    const double s5 = 2.23606797749978969641;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const double r = sqrt_nobranch(r2[v]);
      const double e = sf2 * exp_nonpos(-s5 * r);
      k[v] = (1.0 + s5 * r + (5.0 / 3.0) * r2[v]) * e;
      h[v] = -(5.0 / 3.0) * (1.0 + s5 * r) * e;
    }
  } else {
    const double s3 = 1.73205080756887729353;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const double r = sqrt_nobranch(r2[v]);
      const double e = sf2 * exp_nonpos(-s3 * r);
      k[v] = (1.0 + s3 * r) * e;
      h[v] = -3.0 * e;
    }
  }
}

// NV covariance values at once, the (warp-uniform) kernel switch outside the loop: straight-line code in which the
// compiler interleaves the NV independent dependency chains
template <int NV>
__device__ __forceinline__ void kern_k_vec(int kind, double sf2, const double (&r2)[NV], double (&k)[NV]) {
  if (kind == KERN_RBF) {
#pragma unroll
    for (int v = 0; v < NV; ++v) k[v] = sf2 * exp_nonpos(-0.5 * r2[v]);
  } else if (kind == KERN_MATERN52) {
    const double s5 = 2.23606797749978969641;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const double r = sqrt_nobranch(r2[v]);
      k[v] = (1.0 + s5 * r + (5.0 / 3.0) * r2[v]) * (sf2 * exp_nonpos(-s5 * r));
    }
  } else {
    const double s3 = 1.73205080756887729353;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const double r = sqrt_nobranch(r2[v]);
      k[v] = (1.0 + s3 * r) * (sf2 * exp_nonpos(-s3 * r));
    }
  }
}

}  // namespace gpo
