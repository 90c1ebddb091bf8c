// BSE - host mirror of xtp/include/votca/xtp/bse.h:40-169 and xtp/src/libxtp/gwbse/bse.cc:43-360,489-716,
// plus the transition-dipole post-processing of xtp/src/libxtp/orbitals.cc:643-674,742-795.
#pragma once
#include <chrono>
#include <cmath>

#include "bse_operator.h"
#include "davidsonsolver.h"
#include "rpa.h"

namespace votca {
namespace xtp {

// tools::EigenSystem (tools/include/votca/tools/eigensystem.h:27-54)
struct EigenSystem {
  VectorXd eigenvalues;
  MatrixXd eigenvectors;
  MatrixXd eigenvectors2;
  bool success = false;
};

// bse_initialization.h:47-93
inline MatrixXd BuildFullBSEXRankedInitialGuess(const VectorXd& adiag, const VectorXd& bdiag, Index nroots) {
  if (adiag.size() != bdiag.size()) throw std::runtime_error("BuildFullBSEXRankedInitialGuess: size mismatch.");
  const Index n = adiag.size();
  const Index nguess = std::min<Index>(n, std::max<Index>(4 * nroots, 8));
  struct RankedMode {
    double omega, a;
    Index idx;
  };
  std::vector<RankedMode> ranked;
  ranked.reserve(n);
  for (Index i = 0; i < n; ++i) {
    const double a = adiag(i), b = bdiag(i);
    const double disc = std::max(0.0, (a - b) * (a + b));
    ranked.push_back({std::sqrt(disc), a, i});
  }
  std::sort(ranked.begin(), ranked.end(), [](const RankedMode& l, const RankedMode& r) {
    if (l.omega != r.omega) return l.omega < r.omega;
    return l.a < r.a;
  });
  MatrixXd guess = MatrixXd::Zero(2 * n, nguess);
  for (Index col = 0; col < nguess; ++col) guess(ranked[col].idx, col) = 1.0;
  return guess;
}

class BSE {
 public:
  BSE(Logger& log, TCMatrix_gwbse& Mmn) : log_(log), Mmn_(Mmn) {}

  struct options {
    bool useTDA = true;
    Index homo = 0, rpamin = 0, rpamax = 0, qpmin = 0, qpmax = 0, vmin = 0, cmax = 0;
    Index nmax = 5;
    std::string davidson_correction = "DPR";
    std::string davidson_tolerance = "normal";
    std::string davidson_update = "safe";
    Index davidson_maxiter = 50;
    double min_print_weight = 0.5;
    bool use_Hqp_offdiag = true;
    Index max_dyn_iter = 0;
    double dyn_tolerance = 1e-5;
  };

  // bse.cc:43-59
  void configure(const options& opt, const VectorXd& RPAInputEnergies, const MatrixXd& Hqp_in) {
    opt_ = opt;
    bse_vmax_ = opt_.homo;
    bse_cmin_ = opt_.homo + 1;
    bse_vtotal_ = bse_vmax_ - opt_.vmin + 1;
    bse_ctotal_ = opt_.cmax - bse_cmin_ + 1;
    bse_size_ = bse_vtotal_ * bse_ctotal_;
    max_dyn_iter_ = opt_.max_dyn_iter;
    dyn_tolerance_ = opt_.dyn_tolerance;
    Hqp_ = AdjustHqpSize(Hqp_in, RPAInputEnergies);
    if (!opt_.use_Hqp_offdiag) Hqp_ = asDiagonal(Hqp_.diagonal());
    SetupDirectInteractionOperator(RPAInputEnergies, 0.0);
  }

  // bse.cc:147-186 (and bse_uks.cc:92-133, the same routine with the spin channel's homo in `o`): Hqp cut or
  // extended to the BSE window; levels outside the GW window get their RPA input energies on the diagonal
  static MatrixXd AdjustHqpSizeFor(const options& o, const MatrixXd& Hqp, const VectorXd& RPAInputEnergies) {
    const Index vtotal = o.homo - o.vmin + 1, ctotal = o.cmax - o.homo;
    Index hqp_size = vtotal + ctotal;
    Index gwsize = o.qpmax - o.qpmin + 1;
    Index RPAoffset = o.vmin - o.rpamin;
    MatrixXd Hqp_BSE = MatrixXd::Zero(hqp_size, hqp_size);
    if (o.vmin >= o.qpmin) {
      Index start = o.vmin - o.qpmin;
      if (o.cmax <= o.qpmax) {
        Hqp_BSE = Hqp.block(start, start, hqp_size, hqp_size);
      } else {
        Index virtoffset = gwsize - start;
        Hqp_BSE.setBlock(0, 0, Hqp.block(start, start, virtoffset, virtoffset));
        Index virt_extra = o.cmax - o.qpmax;
        for (Index i = 0; i < virt_extra; ++i)
          Hqp_BSE(hqp_size - virt_extra + i, hqp_size - virt_extra + i) = RPAInputEnergies(RPAoffset + virtoffset + i);
      }
    }
    if (o.vmin < o.qpmin) {
      Index occ_extra = o.qpmin - o.vmin;
      for (Index i = 0; i < occ_extra; ++i) Hqp_BSE(i, i) = RPAInputEnergies(RPAoffset + i);
      Hqp_BSE.setBlock(occ_extra, occ_extra, Hqp.block(0, 0, gwsize, gwsize));
      if (o.cmax > o.qpmax) {
        Index virtoffset = occ_extra + gwsize;
        Index virt_extra = o.cmax - o.qpmax;
        for (Index i = 0; i < virt_extra; ++i)
          Hqp_BSE(hqp_size - virt_extra + i, hqp_size - virt_extra + i) = RPAInputEnergies(RPAoffset + virtoffset + i);
      }
    }
    return Hqp_BSE;
  }

  const MatrixXd& getHqp() const { return Hqp_; }
  // bse.cc:206-232: the TDA operators as BSECoupling uses them (they refer to this object's screening and Hqp)
  SingletOperator_TDA getSingletOperator_TDA() const {
    SingletOperator_TDA H(epsilon_0_inv_, Mmn_, Hqp_);
    configureBSEOperator(H);
    return H;
  }
  TripletOperator_TDA getTripletOperator_TDA() const {
    TripletOperator_TDA H(epsilon_0_inv_, Mmn_, Hqp_);
    configureBSEOperator(H);
    return H;
  }
  const VectorXd& getEpsilonInv() const { return epsilon_0_inv_; }
  Index last_davidson_iterations() const { return last_iterations_; }
  Index last_operator_columns() const { return last_op_columns_; }

  EigenSystem Solve_singlets() const { return opt_.useTDA ? Solve_singlets_TDA() : Solve_singlets_BTDA(); }
  EigenSystem Solve_triplets() const { return opt_.useTDA ? Solve_triplets_TDA() : Solve_triplets_BTDA(); }

  struct ExpectationValues {
    VectorXd direct_term, cross_term;
  };
  struct Interaction {
    VectorXd exchange_contrib, direct_contrib, qp_contrib;
  };

  // bse.cc:553-606
  Interaction Analyze_eh_interaction(bool singlet, const EigenSystem& es) const {
    Interaction analysis;
    {
      HqpOperator hqp(epsilon_0_inv_, Mmn_, Hqp_);
      configureBSEOperator(hqp);
      analysis.qp_contrib = ExpectationValue_Operator(es, hqp).direct_term;
    }
    {
      HdOperator hd(epsilon_0_inv_, Mmn_, Hqp_);
      configureBSEOperator(hd);
      analysis.direct_contrib = ExpectationValue_Operator(es, hd).direct_term;
    }
    if (!opt_.useTDA) {
      Hd2Operator hd2(epsilon_0_inv_, Mmn_, Hqp_);
      configureBSEOperator(hd2);
      analysis.direct_contrib += ExpectationValue_Operator(es, hd2).cross_term;
    }
    const double xpref = singlet ? 2.0 : 0.0;
    if (xpref != 0.0) {
      HxOperator hx(epsilon_0_inv_, Mmn_, Hqp_);
      configureBSEOperator(hx);
      ExpectationValues ev = ExpectationValue_Operator(es, hx);
      analysis.exchange_contrib = xpref * ev.direct_term;
      if (!opt_.useTDA) analysis.exchange_contrib += xpref * ev.cross_term;
    } else {
      analysis.exchange_contrib = VectorXd::Zero(analysis.direct_contrib.size());
    }
    return analysis;
  }

  // bse.cc:608-716
  VectorXd Perturbative_DynamicalScreening(const EigenSystem& es, const VectorXd& RPAInputEnergies) {
    SetupDirectInteractionOperator(RPAInputEnergies, 0.0);
    VectorXd Hd_static_contribution;
    {
      HdOperator Hd_static(epsilon_0_inv_, Mmn_, Hqp_);
      configureBSEOperator(Hd_static);
      Hd_static_contribution = ExpectationValue_Operator(es, Hd_static).direct_term;
    }
    if (!opt_.useTDA) {
      Hd2Operator Hd2_static(epsilon_0_inv_, Mmn_, Hqp_);
      configureBSEOperator(Hd2_static);
      Hd_static_contribution += ExpectationValue_Operator(es, Hd2_static).cross_term;
    }
    const VectorXd& BSEenergies = es.eigenvalues;
    VectorXd BSEenergies_dynamic = BSEenergies;
    for (Index i_exc = 0; i_exc < BSEenergies.size(); i_exc++) {
      for (Index iter = 0; iter < max_dyn_iter_; iter++) {
        double old_energy = BSEenergies_dynamic(i_exc);
        SetupDirectInteractionOperator(RPAInputEnergies, old_energy);
        VectorXd Hd_dynamic_contribution;
        {
          HdOperator Hd_dyn(epsilon_0_inv_, Mmn_, Hqp_);
          configureBSEOperator(Hd_dyn);
          Hd_dynamic_contribution = ExpectationValue_Operator_State(i_exc, es, Hd_dyn).direct_term;
        }
        if (!opt_.useTDA) {
          Hd2Operator Hd2_dyn(epsilon_0_inv_, Mmn_, Hqp_);
          configureBSEOperator(Hd2_dyn);
          Hd_dynamic_contribution += ExpectationValue_Operator_State(i_exc, es, Hd2_dyn).cross_term;
        }
        BSEenergies_dynamic(i_exc) = BSEenergies(i_exc) + Hd_static_contribution(i_exc) - Hd_dynamic_contribution(0);
        if (std::abs(BSEenergies_dynamic(i_exc) - old_energy) < dyn_tolerance_) break;
      }
    }
    return BSEenergies_dynamic;
  }

  // orbitals.cc:764-795: d_s = -sqrt(2) sum_vc (X+Y)_vc <c|r|v>; interlevel[i] is ctotal x vtotal
  std::vector<VectorXd> CalcCoupledTransition_Dipoles(const EigenSystem& es,
                                                      const std::vector<MatrixXd>& interlevel_dipoles) const {
    std::vector<VectorXd> out;
    const double sqrt2 = std::sqrt(2.0);
    for (Index i_exc = 0; i_exc < es.eigenvalues.size(); ++i_exc) {
      VectorXd coeffs = es.eigenvectors.col(i_exc);
      if (!opt_.useTDA) coeffs += es.eigenvectors2.col(i_exc);
      VectorXd tdipole(3, 0.0);
      for (Index i = 0; i < 3; ++i) {
        double s = 0.0;
        for (Index v = 0; v < bse_vtotal_; ++v)
          for (Index c = 0; c < bse_ctotal_; ++c) s += coeffs(c + bse_ctotal_ * v) * interlevel_dipoles[i](c, v);
        tdipole(i) = -sqrt2 * s;
      }
      out.push_back(tdipole);
    }
    return out;
  }
  // orbitals.cc:643-674
  static VectorXd Oscillatorstrengths(const std::vector<VectorXd>& tdip, const VectorXd& energies) {
    const Index size = std::min<Index>(tdip.size(), energies.size());
    VectorXd oscs(size);
    for (Index i = 0; i < size; ++i) oscs(i) = tdip[i].dot(tdip[i]) * 2.0 / 3.0 * energies(i);
    return oscs;
  }

 private:
  // bse.cc:147-186
  MatrixXd AdjustHqpSize(const MatrixXd& Hqp, const VectorXd& RPAInputEnergies) {
    return AdjustHqpSizeFor(opt_, Hqp, RPAInputEnergies);
  }

  // bse.cc:188-204: eps(energy) -> eigen-decomposition -> rotate Mmn, all on the device
  void SetupDirectInteractionOperator(const VectorXd& RPAInputEnergies, double energy) {
    const Device& dev = Mmn_.device();
    RPA rpa(log_, Mmn_);
    rpa.configure(opt_.homo, opt_.rpamin, opt_.rpamax);
    rpa.setRPAInputEnergies(RPAInputEnergies);
    const Index n = Mmn_.auxsize();
    double* eps = rpa.calculate_epsilon_r_dev(energy);
    Device::Buffer U = dev.alloc(static_cast<size_t>(n * n));
    dev.check(gwbse_d2d(dev.ctx(), U.get(), eps, static_cast<size_t>(n * n)));
    VectorXd ev(n);
    dev.check(gwbse_sym_eig_dev(dev.ctx(), (int)n, U.get(), (int)n, ev.data()));
    // the operators built on this tensor read the rows of the (v, c) window only
    Mmn_.MultiplyRightWithAuxMatrix_dev(U.get(), n, opt_.vmin, opt_.cmax + 1);
    epsilon_0_inv_ = VectorXd::Zero(n);
    for (Index i = 0; i < n; ++i)
      if (ev(i) > 1e-8) epsilon_0_inv_(i) = 1 / ev(i);
  }

  template <typename BSE_OPERATOR>
  void configureBSEOperator(BSE_OPERATOR& H) const {
    BSEOperator_Options opt;
    opt.cmax = opt_.cmax;
    opt.homo = opt_.homo;
    opt.qpmin = opt_.qpmin;
    opt.rpamin = opt_.rpamin;
    opt.vmin = opt_.vmin;
    H.configure(opt);
  }

  void configureDavidson(DavidsonSolver& DS) const {
    DS.set_correction(opt_.davidson_correction);
    DS.set_tolerance(opt_.davidson_tolerance);
    DS.set_size_update(opt_.davidson_update);
    DS.set_iter_max(opt_.davidson_maxiter);
    DS.set_max_search_space(10 * opt_.nmax);
  }

  EigenSystem Solve_singlets_TDA() const {
    SingletOperator_TDA Hs(epsilon_0_inv_, Mmn_, Hqp_);
    configureBSEOperator(Hs);
    log_(" Setup TDA singlet hamiltonian ");
    return solve_hermitian(Hs);
  }
  EigenSystem Solve_triplets_TDA() const {
    TripletOperator_TDA Ht(epsilon_0_inv_, Mmn_, Hqp_);
    configureBSEOperator(Ht);
    return solve_hermitian(Ht);
  }
  EigenSystem Solve_singlets_BTDA() const {
    SingletOperator_TDA A(epsilon_0_inv_, Mmn_, Hqp_);
    configureBSEOperator(A);
    SingletOperator_BTDA_B B(epsilon_0_inv_, Mmn_, Hqp_);
    configureBSEOperator(B);
    log_(" Setup Full singlet hamiltonian ");
    return Solve_nonhermitian_Davidson(A, B);
  }
  EigenSystem Solve_triplets_BTDA() const {
    TripletOperator_TDA A(epsilon_0_inv_, Mmn_, Hqp_);
    configureBSEOperator(A);
    Hd2Operator B(epsilon_0_inv_, Mmn_, Hqp_);
    configureBSEOperator(B);
    log_(" Setup Full triplet hamiltonian ");
    return Solve_nonhermitian_Davidson(A, B);
  }

  // bse.cc:266-293
  template <typename BSE_OPERATOR>
  EigenSystem solve_hermitian(BSE_OPERATOR& h) const {
    auto start = std::chrono::system_clock::now();
    EigenSystem result;
    DavidsonSolver DS(log_);
    configureDavidson(DS);
    DS.solve(h, opt_.nmax);
    result.eigenvalues = DS.eigenvalues();
    result.eigenvectors = DS.eigenvectors();
    result.success = DS.success();
    last_iterations_ = DS.num_iterations();
    last_op_columns_ = DS.num_operator_columns();
    std::chrono::duration<double> el = std::chrono::system_clock::now() - start;
    log_(" Diagonalization done in " + std::to_string(el.count()) + " secs");
    return result;
  }

  // bse.cc:315-360
  template <typename BSE_OPERATOR_A, typename BSE_OPERATOR_B>
  EigenSystem Solve_nonhermitian_Davidson(BSE_OPERATOR_A& Aop, BSE_OPERATOR_B& Bop) const {
    auto start = std::chrono::system_clock::now();
    HamiltonianOperator<BSE_OPERATOR_A, BSE_OPERATOR_B> Hop(Aop, Bop);
    DavidsonSolver DS(log_);
    configureDavidson(DS);
    DS.set_matrix_type("HAM");
    MatrixXd initial_guess = BuildFullBSEXRankedInitialGuess(Aop.diagonal(), Bop.diagonal(), opt_.nmax);
    DS.solve(Hop, opt_.nmax, initial_guess);
    EigenSystem result;
    result.eigenvalues = DS.eigenvalues();
    result.success = DS.success();
    last_iterations_ = DS.num_iterations();
    last_op_columns_ = DS.num_operator_columns();
    const MatrixXd ev = DS.eigenvectors();
    const Index half = Aop.rows(), n = ev.cols();
    result.eigenvectors = MatrixXd(half, n);
    result.eigenvectors2 = MatrixXd(half, n);
    for (Index j = 0; j < n; ++j) {
      double normX = 0.0, normY = 0.0;
      for (Index i = 0; i < half; ++i) {
        normX += ev(i, j) * ev(i, j);
        normY += ev(half + i, j) * ev(half + i, j);
      }
      const double sqinvnorm = std::sqrt(1.0 / (normX - normY));
      for (Index i = 0; i < half; ++i) {
        result.eigenvectors(i, j) = ev(i, j) * sqinvnorm;
        result.eigenvectors2(i, j) = ev(half + i, j) * sqinvnorm;
      }
    }
    std::chrono::duration<double> el = std::chrono::system_clock::now() - start;
    log_(" Diagonalization done in " + std::to_string(el.count()) + " secs");
    return result;
  }

  static VectorXd ExpValue(const MatrixXd& a, const MatrixXd& b) {
    VectorXd r(a.cols());
    for (Index j = 0; j < a.cols(); ++j) {
      double s = 0.0;
      for (Index i = 0; i < a.rows(); ++i) s += a(i, j) * b(i, j);
      r(j) = s;
    }
    return r;
  }
  // bse.cc:500-551
  template <typename BSE_OPERATOR>
  ExpectationValues ExpectationValue_Operator(const EigenSystem& es, const BSE_OPERATOR& H) const {
    ExpectationValues ev;
    const MatrixXd temp = H * es.eigenvectors;
    ev.direct_term = ExpValue(es.eigenvectors, temp);
    if (!opt_.useTDA) {
      ev.direct_term += ExpValue(es.eigenvectors2, H * es.eigenvectors2);
      ev.cross_term = 2 * ExpValue(es.eigenvectors2, temp);
    }
    return ev;
  }
  template <typename BSE_OPERATOR>
  ExpectationValues ExpectationValue_Operator_State(Index state, const EigenSystem& es, const BSE_OPERATOR& H) const {
    EigenSystem one;
    one.eigenvectors = es.eigenvectors.block(0, state, es.eigenvectors.rows(), 1);
    if (!opt_.useTDA) one.eigenvectors2 = es.eigenvectors2.block(0, state, es.eigenvectors2.rows(), 1);
    return ExpectationValue_Operator(one, H);
  }

  options opt_;
  Logger& log_;
  Index bse_vmax_ = 0, bse_cmin_ = 0, bse_size_ = 0, bse_vtotal_ = 0, bse_ctotal_ = 0;
  Index max_dyn_iter_ = 0;
  double dyn_tolerance_ = 1e-5;
  VectorXd epsilon_0_inv_;
  TCMatrix_gwbse& Mmn_;
  MatrixXd Hqp_;
  mutable Index last_iterations_ = 0, last_op_columns_ = 0;
};

}  // namespace xtp
}  // namespace votca
