// TCMatrix_gwbse::Fill(auxbasis, dftbasis, dft_orbitals) - the reference's argument list (threecenter.cc:72-90) - driven
// from C for tests/test_host_layer_on_mock_cpu.py (linked against the test-only CPU mock of the kernel library).
#include <cstring>

#include "../../votca_b200/host/aobasis.h"

using namespace votca;
using namespace votca::xtp;

static AOBasisData make(int nshell, const int* l, const int* nprim, const double* centers, const double* exps,
                        const double* contractions) {
  AOBasisData d;
  d.l.assign(l, l + nshell);
  d.nprim.assign(nprim, nprim + nshell);
  d.centers.assign(centers, centers + 3 * nshell);
  size_t np = 0;
  for (int s = 0; s < nshell; ++s) np += (size_t)nprim[s];
  d.exps.assign(exps, exps + np);
  d.NormalizeFromRawContractions(std::vector<double>(contractions, contractions + np));
  return d;
}

extern "C" int fill_from_bases(int ns_d, const int* l_d, const int* np_d, const double* cen_d, const double* ex_d,
                               const double* raw_d, int ns_a, const int* l_a, const int* np_a, const double* cen_a,
                               const double* ex_a, const double* raw_a, const double* mos, long nbasis, long mmax,
                               double* out /* [m][chi][n] */, long* removed, char* err, int cap) {
  try {
    Device dev(0);
    const AOBasisData dd = make(ns_d, l_d, np_d, cen_d, ex_d, raw_d), ad = make(ns_a, l_a, np_a, cen_a, ex_a, raw_a);
    DeviceAOBasis dft(dev, dd), aux(dev, ad);
    TCMatrix_gwbse Mmn(dev);
    Mmn.Initialize(aux.AOBasisSize(), 0, mmax, 0, nbasis - 1);
    const MatrixXd C(mos, nbasis, nbasis, nbasis);
    Mmn.Fill(aux, dft, C, 17);
    for (long m = 0; m <= mmax; ++m) {
      const MatrixXd s = Mmn[m];
      std::memcpy(out + (size_t)m * s.rows() * s.cols(), s.data(), sizeof(double) * s.rows() * s.cols());
    }
    *removed = Mmn.Removedfunctions();
    Mmn.Rebuild();  // device snapshot path
    const MatrixXd again = Mmn[0];
    for (long i = 0; i < again.rows() * again.cols(); ++i)
      if (again.data()[i] != out[i]) throw std::runtime_error("Rebuild changed the tensor");
    return 0;
  } catch (const std::exception& e) {
    std::strncpy(err, e.what(), cap - 1);
    err[cap - 1] = 0;
    return 1;
  }
}
