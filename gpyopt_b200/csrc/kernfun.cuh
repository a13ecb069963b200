// Stationary covariance functions of the scaled squared distance r2 = sum_q ((x_q - x'_q)/l_q)^2.
// Mirrors GPy kern.src.{rbf.RBF, stationary.Matern32, stationary.Matern52}.K_of_r / dK_dr
//  as used through /root/reference/GPyOpt/models/gpmodel.py:58,221.
#pragma once

namespace gpo {

enum KernelKind { KERN_RBF = 0, KERN_MATERN52 = 1, KERN_MATERN32 = 2 };

// k = K_of_r(r),  h = dK_dr(r) / r   (the factor multiplying (x_q - x'_q)/l_q^2 in gradients_X; finite at r = 0,
// where GPy substitutes 1/r := 0 -- the coordinate difference is zero there, so the product agrees)
__device__ __forceinline__ void kern_k_h(int kind, double sf2, double r2, double& k, double& h) {
  if (kind == KERN_RBF) {
    k = sf2 * exp(-0.5 * r2);
    h = -k;
  } else if (kind == KERN_MATERN52) {
    const double s5 = 2.23606797749978969641;
    const double r = sqrt(r2);
    const double e = sf2 * exp(-s5 * r);
    k = (1.0 + s5 * r + (5.0 / 3.0) * r2) * e;
    h = -(5.0 / 3.0) * (1.0 + s5 * r) * e;
  } else {
    const double s3 = 1.73205080756887729353;
    const double r = sqrt(r2);
    const double e = sf2 * exp(-s3 * r);
    k = (1.0 + s3 * r) * e;
    h = -3.0 * e;
  }
}

__device__ __forceinline__ double kern_k(int kind, double sf2, double r2) {
  double k, h;
  kern_k_h(kind, sf2, r2, k, h);
  return k;
}

}  // namespace gpo
