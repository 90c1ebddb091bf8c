// AO integrals produced on the GPU: the host-layer side of gwbse_basis_* / gwbse_ao3c_block_dev /
// gwbse_ao_coulomb2c.  DeviceAOIntegrals is an AOIntegralSource whose blocks never exist on the host:
// TCMatrix_gwbse::Fill3cMO (threecenter.h) asks DeviceBlock() first and contracts straight from the device buffer.
// It stands where the reference calls libint: ComputeAO3cBlock (xtp/src/libxtp/libint2_calls.cc:544-593) and
// AOCoulomb::Fill (libint2_calls.cc:224-271).
#pragma once
#include <algorithm>
#include <vector>

#include <array>
#include <cmath>

#include "threecenter.h"

namespace votca {
namespace xtp {

// The numbers AOBasis::Fill / AOShell::LibintShell hold per shell (aobasis.cc:85-105, aoshell.cc:65-89): shells in
// atom order, pure functions m = -l..l; coefs include libint's primitive normalisation and VOTCA's shell norm.
struct AOBasisData {
  std::vector<int> l, nprim;
  std::vector<double> centers;  // 3 per shell, bohr
  std::vector<double> exps, coefs;
  Index AOBasisSize() const {
    Index n = 0;
    for (int x : l) n += 2 * x + 1;
    return n;
  }
  // AOBasis::getFuncPerAtom as a map: the atom (shells sharing a centre, in order of appearance) of every function
  std::vector<Index> AtomOfFunction() const {
    std::vector<Index> out;
    std::vector<std::array<double, 3>> atoms;
    for (size_t sh = 0; sh < l.size(); ++sh) {
      const std::array<double, 3> c = {centers[3 * sh], centers[3 * sh + 1], centers[3 * sh + 2]};
      Index a = -1;
      for (size_t k = 0; k < atoms.size(); ++k)
        if (std::abs(atoms[k][0] - c[0]) + std::abs(atoms[k][1] - c[1]) + std::abs(atoms[k][2] - c[2]) < 1e-10)
          a = static_cast<Index>(k);
      if (a < 0) {
        atoms.push_back(c);
        a = static_cast<Index>(atoms.size()) - 1;
      }
      for (int m = 0; m < 2 * l[sh] + 1; ++m) out.push_back(a);
    }
    return out;
  }
  // raw contraction factors as they stand in a basis-set XML (basisset.cc:150-199) -> coefs
  void NormalizeFromRawContractions(const std::vector<double>& contractions) {
    if (contractions.size() != exps.size()) throw std::runtime_error("inconsistent basis description");
    coefs.resize(exps.size());
    if (gwbse_basis_normalize((int)l.size(), l.data(), nprim.data(), exps.data(), contractions.data(), coefs.data()))
      throw std::runtime_error("invalid shell in basis description");
  }
};

class DeviceAOBasis {
 public:
  DeviceAOBasis(const Device& dev, const AOBasisData& d) : dev_(dev) {
    if (d.l.size() != d.nprim.size() || d.centers.size() != 3 * d.l.size() || d.exps.size() != d.coefs.size())
      throw std::runtime_error("inconsistent basis description");
    dev_.check(gwbse_basis_create(dev_.ctx(), (int)d.l.size(), d.l.data(), d.nprim.data(), d.centers.data(),
                                  d.exps.data(), d.coefs.data(), &h_));
  }
  ~DeviceAOBasis() {
    if (h_) gwbse_basis_destroy(dev_.ctx(), h_);
  }
  DeviceAOBasis(const DeviceAOBasis&) = delete;
  DeviceAOBasis& operator=(const DeviceAOBasis&) = delete;
  const gwbse_basis* handle() const { return h_; }
  // AODipole::Fill: <mu | r_k | nu> about the origin, k = x, y, z
  std::vector<MatrixXd> Dipoles() const {
    const Index n = AOBasisSize();
    std::vector<double> buf(static_cast<size_t>(3 * n * n));
    dev_.check(gwbse_ao_dipole(dev_.ctx(), h_, buf.data(), (int)n));
    std::vector<MatrixXd> out;
    for (int k = 0; k < 3; ++k) out.emplace_back(buf.data() + static_cast<size_t>(k) * n * n, n, n, n);
    return out;
  }
  // AOOverlap::Fill: <mu | nu>
  MatrixXd Overlap() const {
    const Index n = AOBasisSize();
    MatrixXd S(n, n);
    dev_.check(gwbse_ao_overlap(dev_.ctx(), h_, S.data(), (int)n));
    return S;
  }
  Index AOBasisSize() const { return gwbse_basis_size(h_); }

 private:
  const Device& dev_;
  gwbse_basis* h_ = nullptr;
};

class DeviceAOIntegrals : public AOIntegralSource {
 public:
  // aux overlap (AOOverlap::Fill) and aux Coulomb matrix (AOCoulomb::Fill) are computed on the device on first
  // use unless supplied
  DeviceAOIntegrals(const Device& dev, const DeviceAOBasis& aux, const DeviceAOBasis& dft,
                    const MatrixXd* aux_overlap = nullptr, const MatrixXd* aux_coulomb = nullptr)
      : dev_(dev), aux_(aux), dft_(dft) {
    if (aux_overlap) {
      S_ = *aux_overlap;
      have_S_ = true;
    }
    if (aux_coulomb) {
      V_ = *aux_coulomb;
      have_V_ = true;
    }
  }
  Index AuxSize() const override { return aux_.AOBasisSize(); }
  Index BasisSize() const override { return dft_.AOBasisSize(); }
  void ComputeAO3cBlock(Index aux_offset, Index aux_count, double* out) const override {
    dev_.check(gwbse_ao3c_block(dev_.ctx(), aux_.handle(), dft_.handle(), (int)aux_offset, (int)aux_count, out));
  }
  const double* DeviceBlock(Index aux_offset, Index aux_count) const override {
    const size_t need = static_cast<size_t>(aux_count) * BasisSize() * BasisSize();
    if (block_.size() < need) block_ = dev_.alloc(need);
    dev_.check(gwbse_ao3c_block_dev(dev_.ctx(), aux_.handle(), dft_.handle(), (int)aux_offset, (int)aux_count,
                                    block_.get()));
    return block_.get();
  }
  const MatrixXd& AuxOverlap() const override {
    if (!have_S_) {
      const Index n = AuxSize();
      S_ = MatrixXd(n, n);
      dev_.check(gwbse_ao_overlap(dev_.ctx(), aux_.handle(), S_.data(), (int)n));
      have_S_ = true;
    }
    return S_;
  }
  const MatrixXd& AuxCoulomb() const override {
    if (!have_V_) {
      const Index n = AuxSize();
      V_ = MatrixXd(n, n);
      dev_.check(gwbse_ao_coulomb2c(dev_.ctx(), aux_.handle(), V_.data(), (int)n));
      have_V_ = true;
    }
    return V_;
  }

 private:
  const Device& dev_;
  const DeviceAOBasis& aux_;
  const DeviceAOBasis& dft_;
  mutable MatrixXd S_, V_;
  mutable bool have_S_ = false, have_V_ = false;
  mutable Device::Buffer block_;
};

inline void TCMatrix_gwbse::Fill(const DeviceAOBasis& auxbasis, const DeviceAOBasis& dftbasis,
                                 const MatrixXd& dft_orbitals, Index aux_block) {
  owned_ints_ = std::make_unique<DeviceAOIntegrals>(dev_, auxbasis, dftbasis);
  Fill(*owned_ints_, dft_orbitals, aux_block);
}

}  // namespace xtp
}  // namespace votca
