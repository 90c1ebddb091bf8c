// TCMatrix_gwbse - host mirror of xtp/include/votca/xtp/threecenter.h:41-142 with the tensor resident
// on the GPU.  Same public methods and argument meaning; the libint-backed AOBasis arguments of Fill()
// are replaced by an AOIntegralSource (the reference-side shim implements it with ComputeAO3cBlock,
// libint2_calls.cc:544-593).
#pragma once
#include <functional>
#include <memory>
#include <vector>

#include "device.h"

namespace votca {
namespace xtp {

// Host-side producer of AO integrals (stays on the host per the north star: libint).
struct AOIntegralSource {
  virtual ~AOIntegralSource() = default;
  virtual Index AuxSize() const = 0;
  virtual Index BasisSize() const = 0;
  // aux_count symmetric N x N matrices for aux functions [aux_offset, aux_offset+aux_count)
  virtual void ComputeAO3cBlock(Index aux_offset, Index aux_count, double* out) const = 0;
  // optional: host pointer to the block if the integrals are already stored contiguously (no copy is made)
  virtual const double* HostBlock(Index /*aux_offset*/, Index /*aux_count*/) const { return nullptr; }
  // optional: device pointer to the same block if the integrals already live on the GPU (nullptr otherwise)
  virtual const double* DeviceBlock(Index /*aux_offset*/, Index /*aux_count*/) const { return nullptr; }
  virtual const MatrixXd& AuxOverlap() const = 0;  // AOOverlap::Fill(auxbasis)
  virtual const MatrixXd& AuxCoulomb() const = 0;  // AOCoulomb::Fill(auxbasis)
};

class DeviceAOBasis;  // aobasis.h

class TCMatrix_gwbse {
 public:
  explicit TCMatrix_gwbse(const Device& dev) : dev_(dev) {}

  // threecenter.cc:29-48
  void Initialize(Index basissize, Index mmin, Index mmax, Index nmin, Index nmax) {
    nmin_ = nmin;
    nmax_ = nmax;
    ntotal_ = nmax - nmin + 1;
    mmin_ = mmin;
    mmax_ = mmax;
    mtotal_ = mmax - mmin + 1;
    auxbasissize_ = basissize;
    dev_.check(gwbse_mmn_alloc(dev_.ctx(), (int)basissize, (int)mmin, (int)mmax, (int)nmin, (int)nmax));
  }

  Index auxsize() const { return auxbasissize_; }
  Index get_mmin() const { return mmin_; }
  Index get_mmax() const { return mmax_; }
  Index get_nmin() const { return nmin_; }
  Index get_nmax() const { return nmax_; }
  Index msize() const { return mtotal_; }
  Index nsize() const { return ntotal_; }
  Index Removedfunctions() const { return removedfunctions_; }
  const Device& device() const { return dev_; }

  // operator[] (threecenter.h:125-134): the tensor lives on the GPU, so this hands out a host copy of
  // slice i fetched on demand (the reference returns a reference into host storage).
  MatrixXd operator[](Index i) const {
    MatrixXd out(ntotal_, auxbasissize_);
    dev_.check(gwbse_mmn_get_slice(dev_.ctx(), (int)i, out.data(), (int)ntotal_));
    return out;
  }
  void set_slice(Index i, const MatrixXd& m) {
    dev_.check(gwbse_mmn_set_slice(dev_.ctx(), (int)i, m.data(), (int)m.rows()));
  }

  // threecenter.cc:72-90.  (Running V^-1/2 on a second context beside the fill was tried in round 2 and withdrawn:
  // the TMA-staged GEMM produced wrong tiles when another stream's grids shared the SMs)
  void Fill(const AOIntegralSource& ints, const MatrixXd& dft_orbitals, Index aux_block = 64) {
    ints_ = &ints;
    dft_orbitals_ = &dft_orbitals;
    aux_block_ = aux_block;
    Fill3cMO(ints, dft_orbitals);
    const MatrixXd& S = ints.AuxOverlap();
    const MatrixXd& V = ints.AuxCoulomb();
    if (S.rows() != auxbasissize_ || V.rows() != auxbasissize_)
      throw std::runtime_error("aux overlap / Coulomb matrices do not match the aux basis size");
    int removed = 0;
    // V^-1/2 stays on the device and is applied from there
    dev_.check(gwbse_pseudo_invsqrt(dev_.ctx(), (int)auxbasissize_, S.data(), V.data(), 5e-7, nullptr, &removed));
    removedfunctions_ = removed;
    MultiplyRightWithAuxMatrix_dev(gwbse_pseudo_invsqrt_result_dev(dev_.ctx()), auxbasissize_);
    if (keep_snapshot_) dev_.check(gwbse_mmn_snapshot(dev_.ctx()));
    have_snapshot_ = keep_snapshot_;
  }

  // threecenter.cc:72-90 with the reference's own argument list (auxbasis, dftbasis, dft_orbitals): overlap, two-
  // and three-centre Coulomb integrals are all produced on the device (defined in aobasis.h)
  void Fill(const DeviceAOBasis& auxbasis, const DeviceAOBasis& dftbasis, const MatrixXd& dft_orbitals,
            Index aux_block = 64);

  // gw.cc:242-246 calls Rebuild() every reset_3c iterations.  With a device snapshot this is a D2D copy;
  // otherwise the integrals are contracted again exactly as the reference does.
  void Rebuild() {
    if (have_snapshot_) {
      dev_.check(gwbse_mmn_restore(dev_.ctx()));
    } else {
      if (!ints_) throw std::runtime_error("TCMatrix_gwbse::Rebuild called before Fill");
      Fill(*ints_, *dft_orbitals_, aux_block_);
    }
  }
  void KeepSnapshot(bool on) { keep_snapshot_ = on; }

  // threecenter.cc:108-131 (QSGW): rotate the n rows of the QP window in the QP-window m slices
  void Rotate(const MatrixXd& U, Index qpmin, Index qpmax) {
    const Index qptotal = qpmax - qpmin + 1;
    if (U.rows() != qptotal || U.cols() != qptotal)
      throw std::runtime_error("TCMatrix_gwbse::Rotate: rotation matrix does not match the QP window");
    dev_.check(gwbse_mmn_rotate(dev_.ctx(), U.data(), (int)U.rows(), (int)qpmin, (int)qpmax));
  }

  // threecenter.cc:54-65
  void MultiplyRightWithAuxMatrix(const MatrixXd& matrix) {
    if (matrix.rows() != auxbasissize_ || matrix.cols() != auxbasissize_)
      throw std::runtime_error("Shape mismatch in MultiplyRightWithAuxMatrix");
    dev_.check(gwbse_mmn_mul_right(dev_.ctx(), matrix.data(), (int)matrix.rows()));
  }
  void MultiplyRightWithAuxMatrix_dev(const double* R_dev, Index ld) {
    dev_.check(gwbse_mmn_mul_right_dev(dev_.ctx(), R_dev, (int)ld));
  }
  // the same product for a caller that goes on to read rows [n_lo, n_hi) (absolute level indices) only: the device
  // rotates those now and the rest when something else needs it (gwbse_mmn_mul_right_window_dev)
  void MultiplyRightWithAuxMatrix_dev(const double* R_dev, Index ld, Index n_lo, Index n_hi) {
    dev_.check(gwbse_mmn_mul_right_window_dev(dev_.ctx(), R_dev, (int)ld, (int)(n_lo - nmin_), (int)(n_hi - nmin_)));
  }

 private:
  // libint2_calls.cc:595-651: aux functions are processed block by block (the reference goes shell by shell)
  void Fill3cMO(const AOIntegralSource& ints, const MatrixXd& dft_orbitals) {
    const Index N = ints.BasisSize();
    if (dft_orbitals.rows() != N) throw std::runtime_error("MO coefficient matrix does not match the basis size");
    dev_.check(gwbse_mmn_set_mos(dev_.ctx(), dft_orbitals.data(), (int)dft_orbitals.rows(), (int)N,
                                 (int)dft_orbitals.cols()));
    // two page-locked blocks: the producer fills one while the other is on its way to the GPU
    struct Pinned {
      double* p = nullptr;
      ~Pinned() { gwbse_host_free(p); }
    } block[2];
    int cur = 0;
    // several GPUs: every rank contracts its share of the aux functions for all m (the reference's parallel
    // loop over aux shells, libint2_calls.cc:621-622); fill_end redistributes to the m-sharded layout
    const bool sharded = dev_.world() > 1;
    int lo = 0, hi = (int)auxbasissize_;
    dev_.check(gwbse_mmn_fill_begin(dev_.ctx(), sharded ? 1 : 0));
    if (sharded) dev_.check(gwbse_shard_aux_range(dev_.ctx(), dev_.rank(), &lo, &hi));
    for (Index a0 = lo; a0 < hi; a0 += aux_block_) {
      const Index cnt = std::min<Index>(aux_block_, hi - a0);
      if (const double* d = ints.DeviceBlock(a0, cnt)) {
        dev_.check(gwbse_mmn_fill_block_dev(dev_.ctx(), (int)a0, (int)cnt, d));
        continue;
      }
      if (const double* h = ints.HostBlock(a0, cnt)) {
        dev_.check(gwbse_mmn_fill_block(dev_.ctx(), (int)a0, (int)cnt, h));
        continue;
      }
      if (!block[cur].p) {
        void* p = nullptr;
        if (gwbse_host_malloc(sizeof(double) * static_cast<size_t>(aux_block_ * N * N), &p))
          throw std::runtime_error("cannot allocate page-locked staging memory for the AO integral blocks");
        block[cur].p = static_cast<double*>(p);
      }
      ints.ComputeAO3cBlock(a0, cnt, block[cur].p);
      dev_.check(gwbse_mmn_fill_block(dev_.ctx(), (int)a0, (int)cnt, block[cur].p));
      cur ^= 1;
    }
    dev_.check(gwbse_mmn_fill_end(dev_.ctx()));
  }

  const Device& dev_;
  Index auxbasissize_ = 0, mmin_ = 0, mmax_ = 0, nmin_ = 0, nmax_ = 0, ntotal_ = 0, mtotal_ = 0;
  Index removedfunctions_ = 0;
  const AOIntegralSource* ints_ = nullptr;
  std::unique_ptr<AOIntegralSource> owned_ints_;  // producer created by Fill(auxbasis, dftbasis, ...)
  const MatrixXd* dft_orbitals_ = nullptr;
  Index aux_block_ = 64;
  bool keep_snapshot_ = true, have_snapshot_ = false;
};

}  // namespace xtp
}  // namespace votca
