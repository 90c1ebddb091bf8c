// BSECoupling - host mirror of xtp/include/votca/xtp/bsecoupling.h and xtp/src/libxtp/bsecoupling.cc (SURVEY.md 8f,
// N4): exciton couplings J_AB between two monomers from a GW-BSE run on their dimer.
//
// Same interface (Initialize / CalculateCouplings / Addoutput / get{Singlet,Triplet}CouplingElement), same numbers.
// What differs is where the work happens.  The reference assembles the projection P = [Frenkel | charge transfer]
// (bse_size x n) on the host and forms J = P^T (H P) with the row-rebuilding BSE operator.  Here P never exists on
// the host: the monomer -> dimer orbital projections, every Frenkel column (two GEMMs per exciton) and every
// charge-transfer column (a rank-1 GEMM) are written by the DMMA GEMM straight into one device buffer, H P is one
// call of the factorised BSE operator (csrc/capi_bse.cu) on all columns at once, J and S = P^T P are two more GEMMs,
// and only the n x n matrices come back for the Loewdin / perturbation / reduction algebra (n = a few dozen).
#pragma once
#include <array>
#include <cstdio>
#include <limits>

#include "bse.h"

namespace votca {
namespace xtp {

// What BSECoupling reads from an Orbitals object (orbitals.h): a monomer needs the first block, the dimer all of it
struct CouplingOrbitals {
  const MatrixXd* mos = nullptr;  // MOs().eigenvectors(), AO basis x levels
  Index bse_vmin = 0, bse_vmax = 0, bse_cmin = 0, bse_cmax = 0;
  const MatrixXd* singlets = nullptr;  // BSESinglets().eigenvectors(): (vtotal * ctotal) x states, index ctotal * v + c
  const MatrixXd* triplets = nullptr;
  const VectorXd* singlet_energies = nullptr;
  const VectorXd* triplet_energies = nullptr;
  // dimer
  Index homo = 0, rpamin = 0, rpamax = 0, qpmin = 0, qpmax = 0;
  const MatrixXd* Hqp = nullptr;  // QPdiag().eigenvectors() * eigenvalues * eigenvectors^T
  const VectorXd* rpa_input_energies = nullptr;
  bool use_Hqp_offdiag = true;
  const MatrixXd* overlap = nullptr;  // AO overlap of the dimer basis (CouplingBase::CalculateOverlapMatrix)

  Index basis_size() const { return mos ? mos->rows() : 0; }
  Index vtotal() const { return bse_vmax - bse_vmin + 1; }
  Index ctotal() const { return bse_cmax - bse_cmin + 1; }
};

class BSECoupling {
 public:
  struct options {  // share/xtp/xml/subpackages/bsecoupling.xml
    std::string spin = "all";
    bool use_perturbation = true;
    bool output_tb = false;
    Index statesA = 5, occLevelsA = 5, unoccLevelsA = 5;
    Index statesB = 5, occLevelsB = 5, unoccLevelsB = 5;
  };
  struct Diagnostics {  // bsecoupling.h:70-76
    double xi = 0.0, pt_rm_discrepancy = 0.0;
    bool downfolding_safe = false;
  };
  struct Channel {
    std::array<MatrixXd, 2> JAB;  // [0] perturbation, [1] reduction; Hartree, (levA + levB)^2
    MatrixXd J_dimer, S_dimer;    // raw projected Hamiltonian / overlap, FE block first
    Diagnostics diag;
    VectorXd monomerA_energies, monomerB_energies;  // eV
  };

  BSECoupling(const Device& dev, Logger& log) : dev_(dev), log_(log) {}
  std::string Identify() const { return "bsecoupling"; }

  // bsecoupling.cc:38-69
  void Initialize(const options& opt) {
    if (opt.spin == "all") {
      doSinglets_ = doTriplets_ = true;
    } else if (opt.spin == "singlet" || opt.spin == "triplet") {
      doSinglets_ = opt.spin == "singlet";
      doTriplets_ = !doSinglets_;
    } else {
      throw std::runtime_error("Choice " + opt.spin + " for type not known. Available singlet,triplet,all");
    }
    output_perturbation_ = opt.use_perturbation;
    output_tb_ = opt.output_tb;
    levA_ = opt.statesA;
    levB_ = opt.statesB;
    occA_ = opt.occLevelsA;
    occB_ = opt.occLevelsB;
    unoccA_ = opt.unoccLevelsA;
    unoccB_ = opt.unoccLevelsB;
  }

  // bsecoupling.cc:356-612
  void CalculateCouplings(const CouplingOrbitals& A, const CouplingOrbitals& B, const CouplingOrbitals& AB,
                          const AOIntegralSource& dimer_integrals) {
    log_("  Calculating exciton couplings");
    const Index basisA = A.basis_size(), basisB = B.basis_size();
    if (basisA == 0 || basisB == 0) throw std::runtime_error("Basis set size is not stored in monomers");
    if (!AB.mos || !AB.Hqp || !AB.rpa_input_energies || !AB.overlap)
      throw std::runtime_error("BSECoupling: dimer orbitals, Hqp, RPA input energies and AO overlap are needed");
    if (basisA + basisB > AB.basis_size())
      throw std::runtime_error("BSECoupling: the dimer basis is smaller than the two monomer bases together");
    ClampToAvailable(A, B);
    const Index ab_vtotal = AB.vtotal(), ab_ctotal = AB.ctotal(), ab_total = ab_vtotal + ab_ctotal;
    const Index ab_size = ab_vtotal * ab_ctotal;
    log_("   levels used for BSE of molA: " + std::to_string(A.bse_vmin) + " to " + std::to_string(A.bse_cmax));
    log_("   levels used for BSE of molB: " + std::to_string(B.bse_vmin) + " to " + std::to_string(B.bse_cmax));
    log_("   levels used for BSE of dimer AB: " + std::to_string(AB.bse_vmin) + " to " + std::to_string(AB.bse_cmax));

    // monomer orbitals in the dimer's MO basis: X_AB = (S MOs_AB)[rows of X]^T MOs_X, on the device
    const Index nAB = AB.basis_size();
    Device::Buffer S = dev_.upload(*AB.overlap);
    Device::Buffer Cab = dev_.upload(AB.mos->block(0, AB.bse_vmin, nAB, ab_total));
    Device::Buffer SC = dev_.alloc(static_cast<size_t>(nAB * ab_total));
    dev_.gemm('N', 'N', nAB, ab_total, nAB, 1.0, S.get(), nAB, Cab.get(), nAB, 0.0, SC.get(), nAB);
    Projected pA = ProjectMonomer(A, SC.get(), nAB, 0, ab_total, "A");
    Projected pB = ProjectMonomer(B, SC.get(), nAB, nAB - basisB, ab_total, "B");

    // dimer Mmn and BSE operator (TDA), bsecoupling.cc:488-512
    TCMatrix_gwbse Mmn(dev_);
    Mmn.Initialize(dimer_integrals.AuxSize(), AB.rpamin, AB.qpmax, AB.rpamin, AB.rpamax);
    Mmn.Fill(dimer_integrals, *AB.mos);
    BSE::options bopt;
    bopt.cmax = AB.bse_cmax;
    bopt.homo = AB.homo;
    bopt.qpmin = AB.qpmin;
    bopt.qpmax = AB.qpmax;
    bopt.rpamax = AB.rpamax;
    bopt.rpamin = AB.rpamin;
    bopt.useTDA = true;
    bopt.vmin = AB.bse_vmin;
    bopt.use_Hqp_offdiag = AB.use_Hqp_offdiag;
    BSE bse(log_, Mmn);
    bse.configure(bopt, *AB.rpa_input_energies, *AB.Hqp);
    log_(" Setup BSE operator");

    const Index n_fe = levA_ + levB_;
    const Index n_ct = occA_ * unoccB_ + unoccA_ * occB_;
    for (int spin = 0; spin < 2; ++spin) {
      if (!(spin == 0 ? doSinglets_ : doTriplets_)) continue;
      log_(spin == 0 ? "   Evaluating singlets" : "   Evaluating triplets");
      const MatrixXd* excA = spin == 0 ? A.singlets : A.triplets;
      const MatrixXd* excB = spin == 0 ? B.singlets : B.triplets;
      Channel& ch = spin == 0 ? singlet_ : triplet_;
      const VectorXd* eA = spin == 0 ? A.singlet_energies : A.triplet_energies;
      const VectorXd* eB = spin == 0 ? B.singlet_energies : B.triplet_energies;
      if (eA) ch.monomerA_energies = hrt2ev * eA->head(std::min<Index>(levA_, eA->size()));
      if (eB) ch.monomerB_energies = hrt2ev * eB->head(std::min<Index>(levB_, eB->size()));
      // projection columns: [Frenkel A | Frenkel B | CT A+B- | CT A-B+], built in place on the device
      Device::Buffer P = dev_.alloc(static_cast<size_t>(ab_size * (n_fe + n_ct)));
      ProjectFrenkelExcitons(*excA, levA_, pA, A, ab_vtotal, ab_ctotal, ab_total, P.get());
      ProjectFrenkelExcitons(*excB, levB_, pB, B, ab_vtotal, ab_ctotal, ab_total, P.get() + levA_ * ab_size);
      SetupCTStates(pA, A, pB, B, ab_vtotal, ab_ctotal, ab_total, P.get() + n_fe * ab_size);
      if (spin == 0) {
        SingletOperator_TDA H = bse.getSingletOperator_TDA();
        ProjectExcitons(P.get(), ab_size, n_fe + n_ct, H, ch);
      } else {
        TripletOperator_TDA H = bse.getTripletOperator_TDA();
        ProjectExcitons(P.get(), ab_size, n_fe + n_ct, H, ch);
      }
      log_(spin == 0 ? "   calculated singlet couplings " : "   calculated triplet couplings ");
    }
    log_("  Done with exciton couplings");
  }

  // bsecoupling.cc:256-266: eV; methodindex 0 = perturbation, 1 = reduction
  double getSingletCouplingElement(Index levelA, Index levelB, Index methodindex) const {
    return singlet_.JAB[methodindex](levelA, levelB + levA_) * hrt2ev;
  }
  double getTripletCouplingElement(Index levelA, Index levelB, Index methodindex) const {
    return triplet_.JAB[methodindex](levelA, levelB + levA_) * hrt2ev;
  }
  const Channel& singlet() const { return singlet_; }
  const Channel& triplet() const { return triplet_; }
  bool doSinglets() const { return doSinglets_; }
  bool doTriplets() const { return doTriplets_; }
  Index levA() const { return levA_; }
  Index levB() const { return levB_; }

  // bsecoupling.cc:127-254: the <bsecoupling> subtree of the job's XML output (attribute formats "%1.6e" / "%1.4f")
  std::string Addoutput() const {
    std::string x = "<" + Identify() + ">\n";
    const std::string algorithm = output_perturbation_ ? "j_pert" : "j_diag";
    for (int spin = 0; spin < 2; ++spin) {
      if (!(spin == 0 ? doSinglets_ : doTriplets_)) continue;
      const Channel& ch = spin == 0 ? singlet_ : triplet_;
      const std::string name = spin == 0 ? "singlet" : "triplet";
      const char tag = spin == 0 ? 's' : 't';
      x += "  <" + name + " algorithm=\"" + algorithm + "\">\n";
      for (Index a = 0; a < levA_; ++a)
        for (Index b = 0; b < levB_; ++b)
          x += "    <coupling stateA=\"" + StateName(tag, a) + "\" stateB=\"" + StateName(tag, b) + "\" j_pert=\"" +
               Sci(ch.JAB[0](a, b + levA_) * hrt2ev) + "\" j_diag=\"" + Sci(ch.JAB[1](a, b + levA_) * hrt2ev) +
               "\"/>\n";
      if (output_tb_) {
        x += "    <monomer_energies>\n";
        for (int f = 0; f < 2; ++f) {
          const VectorXd& e = f == 0 ? ch.monomerA_energies : ch.monomerB_energies;
          x += std::string("      <fragment") + (f == 0 ? "A" : "B") + ">\n";
          for (Index i = 0; i < e.size(); ++i)
            x += "        <energy state=\"" + StateName(tag, i) + "\" eV=\"" + Sci(e(i)) + "\"/>\n";
          x += std::string("      </fragment") + (f == 0 ? "A" : "B") + ">\n";
        }
        x += "    </monomer_energies>\n";
        char xi[64];
        std::snprintf(xi, sizeof xi, "%1.4f", ch.diag.xi);
        x += std::string("    <diagnostics xi=\"") + xi + "\" pt_rm_discrepancy_eV=\"" +
             Sci(ch.diag.pt_rm_discrepancy * hrt2ev) + "\" downfolding_safe=\"" +
             (ch.diag.downfolding_safe ? "true" : "false") + "\"/>\n";
        const Index n_fe = levA_ + levB_, n_ct = ch.J_dimer.rows() - n_fe;
        x += "    <tb_matrices n_FE=\"" + std::to_string(n_fe) + "\" n_CT=\"" + std::to_string(n_ct) + "\" n_occA=\"" +
             std::to_string(occA_) + "\" n_unoccA=\"" + std::to_string(unoccA_) + "\" n_occB=\"" +
             std::to_string(occB_) + "\" n_unoccB=\"" + std::to_string(unoccB_) + "\">\n";
        x += MatrixNode("H_FE_FE", ch.J_dimer.block(0, 0, n_fe, n_fe), hrt2ev);
        x += MatrixNode("S_FE_FE", ch.S_dimer.block(0, 0, n_fe, n_fe), 1.0);
        if (n_ct > 0) {
          x += MatrixNode("H_FE_CT", ch.J_dimer.block(0, n_fe, n_fe, n_ct), hrt2ev);
          x += MatrixNode("S_FE_CT", ch.S_dimer.block(0, n_fe, n_fe, n_ct), 1.0);
          x += MatrixNode("H_CT_CT", ch.J_dimer.block(n_fe, n_fe, n_ct, n_ct), hrt2ev);
          x += MatrixNode("S_CT_CT", ch.S_dimer.block(n_fe, n_fe, n_ct, n_ct), 1.0);
        }
        x += "    </tb_matrices>\n";
      }
      x += "  </" + name + ">\n";
    }
    return x + "</" + Identify() + ">\n";
  }

  static constexpr double hrt2ev = 27.21138602;  // votca::tools::conv::hrt2ev

 private:
  // X_AB on the device (ab_total x x_total, leading dimension ab_total)
  struct Projected {
    Device::Buffer X;
    Index ld = 0, vtotal = 0, ctotal = 0;
    const double* occ_occ(Index col = 0) const { return X.get() + col * ld; }  // rows 0.., occupied columns
    // rows of the dimer's unoccupied levels, column `col` counted from the monomer's first unoccupied level
    const double* unocc_unocc(Index ab_vtotal, Index col = 0) const { return X.get() + ab_vtotal + (vtotal + col) * ld; }
  };

  // bsecoupling.cc:408-468
  void ClampToAvailable(const CouplingOrbitals& A, const CouplingOrbitals& B) {
    auto states = [&](const MatrixXd* a, const MatrixXd* b, const char* what) {
      if (!a || !b) throw std::runtime_error(std::string("BSECoupling: monomer ") + what + " are missing");
      if (levA_ > a->cols()) {
        log_("  Number of excitons you want is greater than stored for molecule A. Setting to max number available");
        levA_ = a->cols();
      }
      if (levB_ > b->cols()) {
        log_("  Number of excitons you want is greater than stored for molecule B. Setting to max number available");
        levB_ = b->cols();
      }
    };
    if (doSinglets_) states(A.singlets, B.singlets, "singlets");
    if (doTriplets_) states(A.triplets, B.triplets, "triplets");
    auto orbitals = [&](Index& n, Index available, const char* what) {
      if (n > available || n < 0) {
        log_(std::string("  Number of ") + what + " for CT creation exceeds number of KS-orbitals in BSE");
        n = available;
      }
    };
    orbitals(unoccA_, A.ctotal(), "unoccupied orbitals in molecule A");
    orbitals(unoccB_, B.ctotal(), "unoccupied orbitals in molecule B");
    orbitals(occA_, A.vtotal(), "occupied orbitals in molecule A");
    orbitals(occB_, B.vtotal(), "occupied orbitals in molecule B");
  }

  // bsecoupling.cc:472-486: rows [row0, row0 + basisX) of S * MOs_AB are the monomer's AO functions
  Projected ProjectMonomer(const CouplingOrbitals& X, const double* SC, Index ldsc, Index row0, Index ab_total,
                           const char* name) const {
    const Index x_total = X.vtotal() + X.ctotal(), nX = X.basis_size();
    Device::Buffer Cx = dev_.upload(X.mos->block(0, X.bse_vmin, nX, x_total));
    Projected p;
    p.X = dev_.alloc(static_cast<size_t>(ab_total * x_total));
    p.ld = ab_total;
    p.vtotal = X.vtotal();
    p.ctotal = X.ctotal();
    dev_.gemm('T', 'N', ab_total, x_total, nX, 1.0, SC + row0, ldsc, Cx.get(), nX, 0.0, p.X.get(), ab_total);
    // how much of every monomer orbital the dimer levels span (the reference's warning uses the same norms)
    const MatrixXd host = dev_.download(p.X.get(), ab_total, x_total);
    double smallest = std::numeric_limits<double>::max();
    for (Index j = 0; j < x_total; ++j) smallest = std::min(smallest, host.col(j).dot(host.col(j)));
    if (smallest < 0.95)
      log_(std::string("Warning: Projection of orbitals of monomer ") + name + " on dimer is insufficient,mag=" +
           std::to_string(smallest));
    return p;
  }

  // bsecoupling.cc:321-345.  Column i of P, read as an (ab_ctotal x ab_vtotal) matrix, is
  // X_unocc_unocc * C_i * X_occ_occ^T with C_i the exciton's (ctotal x vtotal) coefficient matrix
  void ProjectFrenkelExcitons(const MatrixXd& coeffs, Index states, const Projected& p, const CouplingOrbitals& X,
                              Index ab_vtotal, Index ab_ctotal, Index ab_total, double* P) const {
    if (states == 0) return;
    if (coeffs.rows() != X.vtotal() * X.ctotal())
      throw std::runtime_error("BSECoupling: exciton coefficients do not match the monomer's BSE window");
    (void)ab_total;
    Device::Buffer C = dev_.upload(coeffs.leftCols(states));
    Device::Buffer half = dev_.alloc(static_cast<size_t>(ab_ctotal * X.vtotal()));
    for (Index i = 0; i < states; ++i) {
      dev_.gemm('N', 'N', ab_ctotal, X.vtotal(), X.ctotal(), 1.0, p.unocc_unocc(ab_vtotal), p.ld,
                C.get() + i * coeffs.rows(), X.ctotal(), 0.0, half.get(), ab_ctotal);
      dev_.gemm('N', 'T', ab_ctotal, ab_vtotal, X.vtotal(), 1.0, half.get(), ab_ctotal, p.occ_occ(), p.ld, 0.0,
                P + i * ab_ctotal * ab_vtotal, ab_ctotal);
    }
  }

  // bsecoupling.cc:268-319: products |unoccupied of one monomer> <occupied of the other|, A+B- first
  void SetupCTStates(const Projected& pA, const CouplingOrbitals& A, const Projected& pB, const CouplingOrbitals& B,
                     Index ab_vtotal, Index ab_ctotal, Index ab_total, double* P) const {
    (void)ab_total;
    const Index ab_size = ab_vtotal * ab_ctotal;
    auto outer = [&](const double* unocc, const double* occ, double* col) {
      dev_.gemm('N', 'T', ab_ctotal, ab_vtotal, 1, 1.0, unocc, ab_ctotal, occ, ab_vtotal, 0.0, col, ab_ctotal);
    };
    log_("   Setting up CT-states");
    for (Index a = 0; a < occA_; ++a)
      for (Index b = 0; b < unoccB_; ++b)
        outer(pB.unocc_unocc(ab_vtotal, b), pA.occ_occ(A.vtotal() - occA_ + a), P + (a * unoccB_ + b) * ab_size);
    log_("  " + std::to_string(occA_ * unoccB_) + " CT states A+B- created");
    double* P2 = P + occA_ * unoccB_ * ab_size;
    for (Index b = 0; b < occB_; ++b)
      for (Index a = 0; a < unoccA_; ++a)
        outer(pA.unocc_unocc(ab_vtotal, a), pB.occ_occ(B.vtotal() - occB_ + b), P2 + (b * unoccA_ + a) * ab_size);
    log_("  " + std::to_string(unoccA_ * occB_) + " CT states A-B+ created");
  }

  // symmetric eigen-decomposition on the device; returns eigenvalues, M becomes the eigenvectors
  VectorXd Eig(MatrixXd& M) const { return dev_.sym_eig(M); }
  MatrixXd InverseSqrt(const MatrixXd& M, double* smallest = nullptr) const {
    MatrixXd U = M;
    const VectorXd w = Eig(U);
    if (smallest) *smallest = w(0);
    MatrixXd scaled = U;
    for (Index j = 0; j < U.cols(); ++j)
      for (Index i = 0; i < U.rows(); ++i) scaled(i, j) = U(i, j) / std::sqrt(w(j));
    return scaled * U.transpose();
  }

  // bsecoupling.cc:789-835 (ProjectExcitons) with CalcJ_dimer (:683-735); OrthogonalizeCTs (:614-669) only merges
  template <class BSE_OPERATOR>
  void ProjectExcitons(const double* P, Index ab_size, Index n, const BSE_OPERATOR& H, Channel& ch) const {
    log_("   Setting up coupling matrix size " + std::to_string(n));
    Device::Buffer HP = dev_.alloc(static_cast<size_t>(ab_size * n));
    H.apply_dev(P, ab_size, n, HP.get(), ab_size);
    Device::Buffer small = dev_.alloc(static_cast<size_t>(n * n));
    dev_.gemm('T', 'N', n, n, ab_size, 1.0, P, ab_size, HP.get(), ab_size, 0.0, small.get(), n);
    ch.J_dimer = dev_.download(small.get(), n, n);
    dev_.gemm('T', 'N', n, n, ab_size, 1.0, P, ab_size, P, ab_size, 0.0, small.get(), n);
    ch.S_dimer = dev_.download(small.get(), n, n);
    double smallest = 0.0;
    const MatrixXd Sm1 = InverseSqrt(ch.S_dimer, &smallest);
    log_("   Smallest value of dimer overlapmatrix is " + std::to_string(smallest));
    const MatrixXd J_ortho = Sm1 * ch.J_dimer * Sm1;
    log_("   Running Perturbation algorithm");
    ch.JAB[0] = Perturbation(J_ortho);
    log_("    Running Projection algorithm");
    ch.JAB[1] = Fulldiag(J_ortho);
    ch.diag = ComputeDiagnostics(ch.J_dimer, ch.JAB[0], ch.JAB[1]);
  }

  // bsecoupling.cc:846-916: CT block diagonalised, then the CT-mediated second-order term with the symmetric
  // 1/2 (1/(Ea - Ek) + 1/(Eb - Ek)) denominator (eq. 28 of Wehner, Baumeier, JCTC 2017)
  MatrixXd Perturbation(const MatrixXd& J_dimer) const {
    const Index n_fe = levA_ + levB_, n = J_dimer.rows(), n_ct = n - n_fe;
    MatrixXd J = J_dimer;
    if (n_ct > 0) {
      MatrixXd ct = J_dimer.block(n_fe, n_fe, n_ct, n_ct);
      Eig(ct);
      MatrixXd T = MatrixXd::Identity(n, n);
      T.setBlock(n_fe, n_fe, ct);
      J = T.transpose() * J_dimer * T;
    }
    MatrixXd out = MatrixXd::Zero(n_fe, n_fe);
    for (Index a = 0; a < levA_; ++a)
      for (Index b = 0; b < levB_; ++b) {
        const Index bd = b + levA_;
        const double Ea = J(a, a), Eb = J(bd, bd);
        double j = J(a, bd);
        for (Index k = n_fe; k < n; ++k) {
          const double Ek = J(k, k);
          if (std::abs(Ek - Ea) < 0.001 || std::abs(Ek - Eb) < 0.001)
            log_("Energydifference between a Frenkel state and CT state " + std::to_string(k + 1) + " is below 1 mHrt");
          j += 0.5 * J(k, a) * J(k, bd) * (1.0 / (Ea - Ek) + 1.0 / (Eb - Ek));
        }
        out(a, bd) = out(bd, a) = j;
      }
    return out;
  }

  // bsecoupling.cc:918-1016: per pair, the two eigenstates of the full FE + CT problem that look most like the two
  // Frenkel states, restricted to those two rows, Loewdin-orthogonalised, rotated back: H_eff = T E T^T
  MatrixXd Fulldiag(const MatrixXd& J_dimer) const {
    const Index n_fe = levA_ + levB_, n = J_dimer.rows();
    MatrixXd U = J_dimer;
    const VectorXd w = Eig(U);
    auto dominant = [&](Index row, Index skip) {
      Index best = -1;
      for (Index k = 0; k < n; ++k)
        if (k != skip && (best < 0 || std::abs(U(row, k)) > std::abs(U(row, best)))) best = k;
      return best;
    };
    MatrixXd out = MatrixXd::Zero(n_fe, n_fe);
    for (Index a = 0; a < levA_; ++a)
      for (Index b = 0; b < levB_; ++b) {
        const Index bd = b + levA_;
        Index pick[2] = {dominant(a, -1), dominant(bd, -1)};
        if (pick[0] == pick[1]) pick[1] = dominant(bd, pick[1]);
        const Index row[2] = {a, bd};
        double T[2][2], E[2] = {w(pick[0]), w(pick[1])};
        for (int i = 0; i < 2; ++i) {
          const double lead = U(row[i], pick[i]);
          const double sign = lead < 0 ? -1.0 : (lead > 0 ? 1.0 : 0.0);
          T[0][i] = sign * U(a, pick[i]);
          T[1][i] = sign * U(bd, pick[i]);
          const double norm = std::hypot(T[0][i], T[1][i]);
          if (norm > 0) {  // Eigen's normalize() leaves a zero vector alone
            T[0][i] /= norm;
            T[1][i] /= norm;
          }
        }
        if (T[0][0] * T[1][1] - T[0][1] * T[1][0] < 0) {
          T[0][1] = -T[0][1];
          T[1][1] = -T[1][1];
        }
        // S = T T^T, s = S^-1/2 from the closed-form 2 x 2 eigen-decomposition
        const double s00 = T[0][0] * T[0][0] + T[0][1] * T[0][1], s11 = T[1][0] * T[1][0] + T[1][1] * T[1][1];
        const double s01 = T[0][0] * T[1][0] + T[0][1] * T[1][1];
        const double mean = 0.5 * (s00 + s11), gap = std::hypot(0.5 * (s00 - s11), s01);
        const double l0 = mean - gap, l1 = mean + gap;
        // eigenvector of l1: (cos t, sin t) with tan 2t = 2 s01 / (s00 - s11)
        const double t = 0.5 * std::atan2(2.0 * s01, s00 - s11), c = std::cos(t), sn = std::sin(t);
        const double r0 = 1.0 / std::sqrt(l0), r1 = 1.0 / std::sqrt(l1);
        const double m00 = r1 * c * c + r0 * sn * sn, m11 = r1 * sn * sn + r0 * c * c, m01 = (r1 - r0) * c * sn;
        const double sm[2][2] = {{m00, m01}, {m01, m11}};
        // E' = s E s, T' = T s, J = T' E' T'^T
        double Es[2][2], Ts[2][2], J2[2][2];
        for (int i = 0; i < 2; ++i)
          for (int j = 0; j < 2; ++j) {
            Es[i][j] = sm[i][0] * E[0] * sm[0][j] + sm[i][1] * E[1] * sm[1][j];
            Ts[i][j] = T[i][0] * sm[0][j] + T[i][1] * sm[1][j];
          }
        for (int i = 0; i < 2; ++i)
          for (int j = 0; j < 2; ++j) {
            J2[i][j] = 0.0;
            for (int k = 0; k < 2; ++k)
              for (int l = 0; l < 2; ++l) J2[i][j] += Ts[i][k] * Es[k][l] * Ts[j][l];
          }
        out(a, bd) = J2[0][1];
        out(bd, a) = J2[1][0];
      }
    return out;
  }

  // bsecoupling.cc:748-787
  Diagnostics ComputeDiagnostics(const MatrixXd& J_dimer, const MatrixXd& J_pert, const MatrixXd& J_diag) const {
    const Index n_fe = levA_ + levB_, n = J_dimer.rows();
    Diagnostics d;
    for (Index i = 0; i < n_fe; ++i)
      for (Index k = n_fe; k < n; ++k) {
        const double dE = std::abs(J_dimer(i, i) - J_dimer(k, k));
        d.xi = dE > 1e-10 ? std::max(d.xi, std::abs(J_dimer(i, k)) / dE) : std::numeric_limits<double>::infinity();
      }
    for (Index i = 0; i < levA_; ++i)
      for (Index j = 0; j < levB_; ++j)
        d.pt_rm_discrepancy = std::max(d.pt_rm_discrepancy, std::abs(J_pert(i, j + levA_) - J_diag(i, j + levA_)));
    d.downfolding_safe = std::isfinite(d.xi) && d.xi < 0.3 && d.pt_rm_discrepancy < 1e-4;
    return d;
  }

  static std::string Sci(double v) {
    char buf[64];
    std::snprintf(buf, sizeof buf, "%1.6e", v);
    return buf;
  }
  static std::string StateName(char type, Index i) { return std::string(1, type) + std::to_string(i + 1); }  // QMState::ToString
  static std::string MatrixNode(const std::string& name, const MatrixXd& m, double conversion) {  // :77-92
    std::string x = "      <" + name + " rows=\"" + std::to_string(m.rows()) + "\" cols=\"" + std::to_string(m.cols()) + "\"";
    for (Index i = 0; i < m.rows(); ++i) {
      x += " row_" + std::to_string(i) + "=\"";
      for (Index j = 0; j < m.cols(); ++j) x += (j ? " " : "") + Sci(m(i, j) * conversion);
      x += "\"";
    }
    return x + "/>\n";
  }

  const Device& dev_;
  Logger& log_;
  bool doSinglets_ = true, doTriplets_ = true, output_perturbation_ = true, output_tb_ = false;
  Index levA_ = 0, levB_ = 0, occA_ = 0, occB_ = 0, unoccA_ = 0, unoccB_ = 0;
  Channel singlet_, triplet_;
};

}  // namespace xtp
}  // namespace votca
