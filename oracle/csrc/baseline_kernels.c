/* CPU baseline kernels of the oracle (TEST / BENCH INFRASTRUCTURE, never linked into the product).
 *
 * Compiled restatements of the reference's hot scalar loops so that the CPU baseline clock is not
 * inflated by Python overhead.  Loop bodies follow the cited reference lines; OpenMP parallelism is over
 * the same index the reference parallelises (qp levels: gw.cc:344, sigma_base.cc:57,68).
 *   gcc -O3 -march=native -fopenmp -shared -fPIC baseline_kernels.c -o ../lib/liboracle_baseline.so
 */
#include <math.h>
#include <stddef.h>

/* Sigma_PPM::CalcCorrelationDiagElement, sigma_ppm.cc:37-66.
 * M: one level slice stored [naux][n] (column i_aux of the reference's n x naux matrix is contiguous). */
static double sigma_c_ppm_one(const double* M, int n, int naux, int lumo, double eta2, const double* weight,
                              const double* pfreq, const double* energies, double frequency) {
  double sigma = 0.0;
  for (int i_aux = 0; i_aux < naux; ++i_aux) {
    if (weight[i_aux] < 1.e-9) continue;
    const double ppm_freq = pfreq[i_aux];
    const double fac = 0.5 * weight[i_aux] * ppm_freq;
    const double* col = M + (size_t)i_aux * n;
    double s = 0.0;
    for (int k = 0; k < lumo; ++k) {
      const double t = frequency - energies[k] + ppm_freq;
      s += col[k] * col[k] * t / (t * t + eta2);
    }
    for (int k = lumo; k < n; ++k) {
      const double t = frequency - energies[k] - ppm_freq;
      s += col[k] * col[k] * t / (t * t + eta2);
    }
    sigma += fac * s;
  }
  return sigma;
}

/* nreq independent (level, frequency) evaluations, one OpenMP thread each (as the per-level QP searches).
 * M: [nlevels][naux][n]. */
void sigma_c_ppm_diag_batch(const double* M, int n, int naux, int lumo, double eta, const double* weight,
                            const double* pfreq, const double* energies, int nreq, const int* levels,
                            const double* freqs, double* out) {
  const double eta2 = eta * eta;
#pragma omp parallel for schedule(dynamic)
  for (int r = 0; r < nreq; ++r)
    out[r] = sigma_c_ppm_one(M + (size_t)levels[r] * naux * n, n, naux, lumo, eta2, weight, pfreq, energies,
                             freqs[r]);
}

/* Sigma_PPM::CalcCorrelationOffDiagElement, sigma_ppm.cc:93-126, for npairs level pairs */
void sigma_c_ppm_offdiag_batch(const double* M, int n, int naux, int lumo, double eta, const double* weight,
                               const double* pfreq, const double* energies, int npairs, const int* l1,
                               const int* l2, const double* f1, const double* f2, double* out) {
  const double eta2 = eta * eta;
#pragma omp parallel for schedule(dynamic)
  for (int r = 0; r < npairs; ++r) {
    const double* M1 = M + (size_t)l1[r] * naux * n;
    const double* M2 = M + (size_t)l2[r] * naux * n;
    double sigma_c = 0.0;
    for (int i_aux = 0; i_aux < naux; ++i_aux) {
      if (weight[i_aux] < 1.e-9) continue;
      const double ppm_freq = pfreq[i_aux];
      const double fac = 0.25 * weight[i_aux] * ppm_freq;
      const double* c1 = M1 + (size_t)i_aux * n;
      const double* c2 = M2 + (size_t)i_aux * n;
      double s = 0.0;
      for (int k = 0; k < n; ++k) {
        const double shift = k < lumo ? -ppm_freq : ppm_freq;
        const double t1 = f1[r] - (energies[k] + shift);
        const double t2 = f2[r] - (energies[k] + shift);
        s += (t1 / (t1 * t1 + eta2) + t2 / (t2 * t2 + eta2)) * c1[k] * c2[k];
      }
      sigma_c += fac * s;
    }
    out[r] = sigma_c;
  }
}

/* Sigma_base::CalcExchangeMatrix inner product, sigma_base.cc:45-47, for npairs level pairs; M [nlevels][naux][n] */
void sigma_x_pairs(const double* M, int n, int naux, int occ, int npairs, const int* l1, const int* l2, double* out) {
#pragma omp parallel for schedule(dynamic)
  for (int r = 0; r < npairs; ++r) {
    const double* M1 = M + (size_t)l1[r] * naux * n;
    const double* M2 = M + (size_t)l2[r] * naux * n;
    double s = 0.0;
    for (int i_aux = 0; i_aux < naux; ++i_aux)
      for (int k = 0; k < occ; ++k) s += M1[(size_t)i_aux * n + k] * M2[(size_t)i_aux * n + k];
    out[r] = -s;
  }
}
