// Gaussian quadratures of the CDA imaginary-axis integration - host mirror of
// xtp/src/libxtp/gaussian_quadrature/gauss_legendre_quadrature.{h,cc} (Gauss_Legendre_Quadrature and
// Gauss_modified_Legendre_Quadrature, the two schemes ImaginaryAxisIntegration is used with).  No device dependency.
#pragma once
#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>

#include "matrix.h"

namespace votca {
namespace xtp {

// Gauss-Legendre nodes/weights on [-1,1] by Newton iteration on P_n (the reference ships the same
// numbers as 50-digit tables, gaussian_quadrature/gauss_legendre_quadrature.cc:28-561).
inline void gauss_legendre(Index n, std::vector<double>& x, std::vector<double>& w) {
  x.assign(n, 0.0);
  w.assign(n, 0.0);
  const double pi = 3.14159265358979323846;
  for (Index i = 0; i < (n + 1) / 2; ++i) {
    double z = std::cos(pi * (double(i) + 0.75) / (double(n) + 0.5));
    double pp = 0.0;
    for (int it = 0; it < 100; ++it) {
      double p1 = 1.0, p2 = 0.0;
      for (Index j = 0; j < n; ++j) {
        const double p3 = p2;
        p2 = p1;
        p1 = ((2.0 * double(j) + 1.0) * z * p2 - double(j) * p3) / double(j + 1);
      }
      pp = double(n) * (z * p1 - p2) / (z * z - 1.0);
      const double z1 = z;
      z = z1 - p1 / pp;
      if (std::abs(z - z1) < 1e-16) break;
    }
    x[i] = -z;
    x[n - 1 - i] = z;
    w[i] = 2.0 / ((1.0 - z * z) * pp * pp);
    w[n - 1 - i] = w[i];
  }
}

// Points / weights mapped to the integration domain (gauss_legendre_quadrature.h:57-67, 81-92):
// "legendre":          x' = tan(pi x / 2),        w' = w (pi/2) / cos^2(pi x / 2)   on (-inf, inf), no symmetry factor
// "modified_legendre": x' = (1 + x) / (2 (1 - x)), w' = w / (1 - x)^2               on (0, inf), integrand doubled
inline void mapped_gauss_legendre(const std::string& scheme, Index order, std::vector<double>& pts,
                                  std::vector<double>& wts, bool& symmetry) {
  std::vector<double> gx, gw;
  gauss_legendre(order, gx, gw);
  const double halfpi = 0.5 * 3.14159265358979323846;
  pts.clear();
  wts.clear();
  if (scheme == "legendre") {
    symmetry = false;
    for (Index j = 0; j < order; ++j) {
      pts.push_back(std::tan(halfpi * gx[j]));
      const double c = std::cos(halfpi * gx[j]);
      wts.push_back(gw[j] * halfpi / (c * c));
    }
  } else if (scheme == "modified_legendre") {
    symmetry = true;
    for (Index j = 0; j < order; ++j) {
      pts.push_back(0.5 * (1.0 + gx[j]) / (1.0 - gx[j]));
      wts.push_back(gw[j] / ((1.0 - gx[j]) * (1.0 - gx[j])));
    }
  } else {
    throw std::runtime_error("quadrature scheme '" + scheme + "' is not available in this build");
  }
}

}  // namespace xtp
}  // namespace votca
