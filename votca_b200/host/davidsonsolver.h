// DavidsonSolver - host mirror of xtp/include/votca/xtp/davidsonsolver.h:47-334 and
// xtp/src/libxtp/davidsonsolver.cc:149-555.  Same options, same iteration (Ritz / harmonic Ritz, DPR or Olsen
// correction, twice-applied Gram-Schmidt, restart at max_search_space).  The search space V, A*V (and A*A*V)
// live on the GPU; only the small projected matrices T, B and their eigen-decompositions touch the host.
#pragma once
#include <chrono>
#include <cstdio>
#include <numeric>

#include "device.h"

namespace votca {
namespace xtp {

class DavidsonSolver {
 public:
  explicit DavidsonSolver(Logger& log) : log_(log) {}

  void set_iter_max(Index N) { iter_max_ = N; }
  void set_max_search_space(Index N) { max_search_space_ = N; }
  void set_tolerance(std::string tol) {
    if (tol == "loose") tol_ = 1E-3;
    else if (tol == "normal") tol_ = 1E-4;
    else if (tol == "strict") tol_ = 1E-5;
    else if (tol == "lapack") tol_ = 1E-9;
    else throw std::runtime_error(tol + " is not a valid Davidson tolerance");
  }
  void set_correction(std::string method) {
    if (method == "DPR") davidson_correction_ = CORR::DPR;
    else if (method == "OLSEN") davidson_correction_ = CORR::OLSEN;
    else throw std::runtime_error(method + " is not a valid Davidson correction method");
  }
  void set_size_update(std::string update_size) {
    if (update_size == "min") davidson_update_ = UPDATE::MIN;
    else if (update_size == "safe") davidson_update_ = UPDATE::SAFE;
    else if (update_size == "max") davidson_update_ = UPDATE::MAX;
    else throw std::runtime_error(update_size + " is not a valid Davidson update");
  }
  void set_matrix_type(std::string mt) {
    if (mt == "HAM") matrix_type_ = MATRIX_TYPE::HAM;
    else if (mt == "SYMM") matrix_type_ = MATRIX_TYPE::SYMM;
    else throw std::runtime_error(mt + " is not a valid Davidson matrix type");
  }

  bool success() const { return success_; }
  VectorXd eigenvalues() const { return eigenvalues_; }
  MatrixXd eigenvectors() const { return eigenvectors_; }
  Index num_iterations() const { return i_iter_; }
  Index num_operator_columns() const { return op_columns_; }

  template <typename MatrixReplacement>
  void solve(const MatrixReplacement& A, Index neigen, Index size_initial_guess = 0) {
    prepare(A, neigen);
    if (size_initial_guess == 0) size_initial_guess = 2 * neigen;
    restart_size_ = size_initial_guess;
    MatrixXd guess = setupInitialEigenvectors(size_initial_guess);
    iterate(A, neigen, guess, false);
  }

  template <typename MatrixReplacement>
  void solve(const MatrixReplacement& A, Index neigen, const MatrixXd& initial_guess) {
    prepare(A, neigen);
    if (initial_guess.rows() != op_size_)
      throw std::runtime_error("DavidsonSolver::solve initial_guess has wrong number of rows.");
    if (initial_guess.cols() < neigen)
      throw std::runtime_error("DavidsonSolver::solve initial_guess has fewer columns than neigen.");
    restart_size_ = std::min<Index>(initial_guess.cols(), max_search_space_);
    iterate(A, neigen, initial_guess, true);
  }

 private:
  enum CORR { DPR, OLSEN };
  enum UPDATE { MIN, SAFE, MAX };
  enum MATRIX_TYPE { SYMM, HAM };

  struct RitzEigenPair {
    VectorXd lambda;
    MatrixXd U;             // eigenvectors of the small subspace (host)
    Device::Buffer q, res;  // Ritz vectors and residues (device, op_size x ncols)
    Index ncols = 0;
    VectorXd res_norm;
  };

  struct ProjectedSpace {
    Device::Buffer V, AV, AAV;  // op_size x capacity
    Index ncols = 0;            // columns in V
    Index nav = 0;              // columns in AV
    MatrixXd T, B;
    Index size_update = 0;
    std::vector<char> root_converged;
    Index search_space() const { return ncols; }
  };

  template <typename MatrixReplacement>
  void prepare(const MatrixReplacement& A, Index neigen) {
    dev_ = &A.device();
    if (max_search_space_ < neigen) max_search_space_ = neigen * 5;
    start_ = std::chrono::system_clock::now();
    op_size_ = A.rows();
    op_columns_ = 0;
    checkOptions(op_size_);
    printOptions(op_size_);
    Adiag_ = A.diagonal();
    Adiag_dev_ = dev_->upload(Adiag_);
  }

  // davidsonsolver.h:64-203 (both overloads share this loop)
  template <typename MatrixReplacement>
  void iterate(const MatrixReplacement& A, Index neigen, const MatrixXd& guess, bool orthogonalize_guess) {
    ProjectedSpace proj;
    proj.size_update = getSizeUpdate(neigen);
    capacity_ = std::max<Index>(guess.cols(), max_search_space_) + 2 * proj.size_update + restart_size_ + 2;
    const size_t cap = static_cast<size_t>(op_size_ * capacity_);
    proj.V = dev_->alloc(cap);
    proj.AV = dev_->alloc(cap);
    if (matrix_type_ == MATRIX_TYPE::HAM) proj.AAV = dev_->alloc(cap);
    dev_->check(gwbse_h2d(dev_->ctx(), proj.V.get(), guess.data(), static_cast<size_t>(guess.size())));
    proj.ncols = guess.cols();
    if (orthogonalize_guess) orthogonalize(proj, proj.ncols);
    proj.root_converged.assign(proj.size_update, 0);
    RitzEigenPair rep;
    log_(" iter\tSearch Space\tNorm");
    for (i_iter_ = 0; i_iter_ < iter_max_; i_iter_++) {
      updateProjection(A, proj);
      rep = getRitzEigenPairs(proj);
      bool converged = checkConvergence(rep, proj, neigen);
      printIterationData(rep, proj, neigen);
      bool last_iter = i_iter_ == (iter_max_ - 1);
      if (converged) {
        storeConvergedData(rep, neigen);
        break;
      } else if (last_iter) {
        storeNotConvergedData(rep, proj.root_converged, neigen);
        break;
      }
      Index extension_size = extendProjection(rep, proj);
      bool do_restart = (proj.search_space() > max_search_space_);
      if (do_restart) restart(rep, proj, extension_size);
    }
    printTiming();
  }

  double* col(const Device::Buffer& b, Index j) const { return b.get() + j * op_size_; }

  // Vt(cols a..a+na) ^T * W(cols b..b+nb) -> host (na x nb)
  MatrixXd project(const Device::Buffer& Vb, Index a, Index na, const Device::Buffer& Wb, Index b, Index nb) const {
    Device::Buffer out = dev_->alloc(static_cast<size_t>(std::max<Index>(na * nb, 1)));
    dev_->gemm('T', 'N', na, nb, op_size_, 1.0, col(Vb, a), op_size_, col(Wb, b), op_size_, 0.0, out.get(), na);
    return dev_->download(out.get(), na, nb);
  }

  // davidsonsolver.h:239-280
  template <typename MatrixReplacement>
  void updateProjection(const MatrixReplacement& A, ProjectedSpace& proj) {
    if (i_iter_ == 0) {
      A.apply_dev(proj.V.get(), op_size_, proj.ncols, proj.AV.get(), op_size_);
      op_columns_ += proj.ncols;
      proj.nav = proj.ncols;
      proj.T = project(proj.V, 0, proj.ncols, proj.AV, 0, proj.ncols);
      if (matrix_type_ == MATRIX_TYPE::HAM) {
        A.apply_dev(proj.AV.get(), op_size_, proj.ncols, proj.AAV.get(), op_size_);
        op_columns_ += proj.ncols;
        proj.B = project(proj.V, 0, proj.ncols, proj.AAV, 0, proj.ncols);
      }
      return;
    }
    const Index old_dim = proj.nav;
    const Index new_dim = proj.ncols;
    const Index nvec = new_dim - old_dim;
    A.apply_dev(col(proj.V, old_dim), op_size_, nvec, col(proj.AV, old_dim), op_size_);
    op_columns_ += nvec;
    proj.nav = new_dim;
    MatrixXd T(new_dim, new_dim);
    T.setBlock(0, 0, proj.T);
    T.setBlock(0, old_dim, project(proj.V, 0, new_dim, proj.AV, old_dim, nvec));
    if (matrix_type_ == MATRIX_TYPE::SYMM) {
      for (Index i = 0; i < nvec; ++i)
        for (Index j = 0; j < old_dim; ++j) T(old_dim + i, j) = T(j, old_dim + i);
    } else {
      T.setBlock(old_dim, 0, project(proj.V, old_dim, nvec, proj.AV, 0, old_dim));
      A.apply_dev(col(proj.AV, old_dim), op_size_, nvec, col(proj.AAV, old_dim), op_size_);
      op_columns_ += nvec;
      MatrixXd B(new_dim, new_dim);
      B.setBlock(0, 0, proj.B);
      B.setBlock(0, old_dim, project(proj.V, 0, new_dim, proj.AAV, old_dim, nvec));
      B.setBlock(old_dim, 0, project(proj.V, old_dim, nvec, proj.AAV, 0, old_dim));
      proj.B = B;
    }
    proj.T = T;
  }

  RitzEigenPair getRitzEigenPairs(const ProjectedSpace& proj) const {
    return matrix_type_ == MATRIX_TYPE::SYMM ? getRitz(proj) : getHarmonicRitz(proj);
  }

  // q = V U, res = AV U - q diag(lambda), residual norms
  void finishRitz(const ProjectedSpace& proj, RitzEigenPair& rep) const {
    const Index dim = proj.T.cols(), n = rep.U.cols();
    rep.ncols = n;
    Device::Buffer U = dev_->upload(rep.U);
    rep.q = dev_->alloc(static_cast<size_t>(op_size_ * n));
    rep.res = dev_->alloc(static_cast<size_t>(op_size_ * n));
    dev_->gemm('N', 'N', op_size_, n, dim, 1.0, proj.V.get(), op_size_, U.get(), dim, 0.0, rep.q.get(), op_size_);
    dev_->gemm('N', 'N', op_size_, n, dim, 1.0, proj.AV.get(), op_size_, U.get(), dim, 0.0, rep.res.get(), op_size_);
    Device::Buffer tmp = dev_->alloc(static_cast<size_t>(op_size_ * n));
    dev_->check(gwbse_d2d(dev_->ctx(), tmp.get(), rep.q.get(), static_cast<size_t>(op_size_ * n)));
    dev_->check(gwbse_scale_cols_dev(dev_->ctx(), (int)op_size_, (int)n, tmp.get(), (int)op_size_, rep.lambda.data()));
    dev_->check(gwbse_axpy_dev(dev_->ctx(), (int)op_size_, (int)n, -1.0, tmp.get(), (int)op_size_, rep.res.get(),
                               (int)op_size_));
    rep.res_norm = VectorXd(n);
    dev_->check(gwbse_colnorms_dev(dev_->ctx(), (int)op_size_, (int)n, rep.res.get(), (int)op_size_,
                                   rep.res_norm.data()));
  }

  // davidsonsolver.cc:221-239
  RitzEigenPair getRitz(const ProjectedSpace& proj) const {
    RitzEigenPair rep;
    MatrixXd T = proj.T;
    VectorXd ev;
    try {
      ev = dev_->sym_eig(T);
    } catch (const std::exception&) {
      throw std::runtime_error("Small hermitian eigenvalue problem failed.");
    }
    Index needed_pairs = std::min(proj.T.cols(), std::max(restart_size_, proj.size_update));
    rep.lambda = ev.head(needed_pairs);
    rep.U = T.leftCols(needed_pairs);
    finishRitz(proj, rep);
    return rep;
  }

  // davidsonsolver.cc:241-332
  RitzEigenPair getHarmonicRitz(const ProjectedSpace& proj) const {
    RitzEigenPair rep;
    const Index dim = proj.T.cols();
    VectorXd wr(dim), wi(dim);
    MatrixXd VR(dim, dim);
    try {
      dev_->check(gwbse_gen_eig_host(dev_->ctx(), (int)dim, proj.T.data(), proj.B.data(), wr.data(), wi.data(),
                                     VR.data()));
    } catch (const std::exception&) {
      throw std::runtime_error("Small generalized eigenvalue problem failed.");
    }
    std::vector<std::pair<Index, Index>> complex_pairs;
    for (Index i = 0; i < dim; i++) {
      if (wi(i) != 0) {
        bool found_partner = false;
        for (auto& pair : complex_pairs) {
          if (pair.second > -1) continue;
          bool are_pair = (std::abs(wr(pair.first) - wr(i)) < 1e-9) && (std::abs(wi(pair.first) + wi(i)) < 1e-9);
          if (are_pair) {
            pair.second = i;
            found_partner = true;
          }
        }
        if (!found_partner) complex_pairs.emplace_back(i, -1);
      }
    }
    for (const auto& pair : complex_pairs)
      if (pair.second < 0)
        throw std::runtime_error("Eigenvalue:" + std::to_string(pair.first) + " is complex but has no partner.");
    if (!complex_pairs.empty())
      log_(" Found " + std::to_string(complex_pairs.size()) + " complex pairs in eigenvalue problem");
    const Index nreal = dim - Index(complex_pairs.size());
    VectorXd eigenvalues(nreal);
    MatrixXd eigenvectors(dim, nreal);
    Index j = 0;
    for (Index i = 0; i < dim; i++) {
      bool is_second = false;
      for (const auto& pair : complex_pairs) is_second = is_second || pair.second == i;
      if (is_second) continue;
      eigenvalues(j) = wr(i);
      // real part of the eigenvector: LAPACK stores (re, im) of a complex pair in columns (i, i+1)
      double nrm = 0.0;
      for (Index r = 0; r < dim; ++r) nrm += VR(r, i) * VR(r, i);
      nrm = std::sqrt(nrm);
      for (Index r = 0; r < dim; ++r) eigenvectors(r, j) = VR(r, i) / nrm;
      j++;
    }
    Index needed_pairs = std::min(proj.T.cols(), std::max(restart_size_, proj.size_update));
    needed_pairs = std::min(needed_pairs, nreal);
    // largest values first (inverse problem), davidsonsolver.cc:321-325
    std::vector<Index> idx = argsort(eigenvalues);
    std::reverse(idx.begin(), idx.end());
    rep.U = MatrixXd(dim, needed_pairs);
    for (Index c = 0; c < needed_pairs; ++c)
      for (Index r = 0; r < dim; ++r) rep.U(r, c) = eigenvectors(r, idx[c]);
    const MatrixXd UtTU = rep.U.transpose() * proj.T * rep.U;
    rep.lambda = UtTU.diagonal();
    finishRitz(proj, rep);
    return rep;
  }

  // davidsonsolver.cc:347-352
  bool checkConvergence(const RitzEigenPair& rep, ProjectedSpace& proj, Index neigen) const {
    bool all = true;
    for (Index j = 0; j < proj.size_update; ++j) {
      // pairs beyond the projected dimension do not exist (tiny operators): nothing to extend for them
      const bool c = j >= rep.res_norm.size() || rep.res_norm(j) < tol_;
      proj.root_converged[j] = c;
      if (j < neigen) all = all && c;
    }
    return all;
  }

  // davidsonsolver.cc:354-378
  Index extendProjection(const RitzEigenPair& rep, ProjectedSpace& proj) {
    Index nupdate = 0;
    for (Index j = 0; j < proj.size_update; ++j) nupdate += proj.root_converged[j] ? 0 : 1;
    Index oldsize = proj.ncols;
    if (oldsize + nupdate > capacity_) throw std::runtime_error("Davidson search space exceeds its allocation");
    Index k = 0;
    for (Index j = 0; j < proj.size_update; j++) {
      if (proj.root_converged[j]) continue;
      const double lam = rep.lambda(j);
      dev_->check(gwbse_davidson_correction_dev(dev_->ctx(), (int)op_size_, 1, davidson_correction_ == CORR::OLSEN,
                                                Adiag_dev_.get(), &lam, col(rep.res, j), (int)op_size_,
                                                col(rep.q, j), (int)op_size_, col(proj.V, oldsize + k),
                                                (int)op_size_));
      k++;
    }
    proj.ncols = oldsize + nupdate;
    orthogonalize(proj, nupdate);
    return nupdate;
  }

  void orthogonalize(ProjectedSpace& proj, Index nupdate) const {
    dev_->check(gwbse_gramschmidt_dev(dev_->ctx(), (int)op_size_, (int)proj.ncols, (int)(proj.ncols - nupdate),
                                      proj.V.get(), (int)op_size_));
  }

  // davidsonsolver.cc:490-513
  void restart(const RitzEigenPair& rep, ProjectedSpace& proj, Index newvectors) const {
    const Index rs = restart_size_;
    const Index oldV = proj.ncols - newvectors;  // columns of V that AV covers
    Device::Buffer newV = dev_->alloc(static_cast<size_t>(op_size_ * capacity_));
    dev_->check(gwbse_d2d(dev_->ctx(), col(newV, rs), col(proj.V, oldV), static_cast<size_t>(op_size_ * newvectors)));
    const Index dim = rep.U.rows();
    MatrixXd Urs = rep.U.leftCols(rs);
    Device::Buffer tmp = dev_->alloc(static_cast<size_t>(op_size_ * rs));
    if (matrix_type_ == MATRIX_TYPE::SYMM) {
      dev_->check(gwbse_d2d(dev_->ctx(), newV.get(), rep.q.get(), static_cast<size_t>(op_size_ * rs)));
      Device::Buffer U = dev_->upload(Urs);
      dev_->gemm('N', 'N', op_size_, rs, dim, 1.0, proj.AV.get(), op_size_, U.get(), dim, 0.0, tmp.get(), op_size_);
      dev_->check(gwbse_d2d(dev_->ctx(), proj.AV.get(), tmp.get(), static_cast<size_t>(op_size_ * rs)));
    } else {
      MatrixXd orthonormal = qr(Urs);
      Device::Buffer Q = dev_->upload(orthonormal);
      dev_->gemm('N', 'N', op_size_, rs, dim, 1.0, proj.V.get(), op_size_, Q.get(), dim, 0.0, newV.get(), op_size_);
      dev_->gemm('N', 'N', op_size_, rs, dim, 1.0, proj.AV.get(), op_size_, Q.get(), dim, 0.0, tmp.get(), op_size_);
      dev_->check(gwbse_d2d(dev_->ctx(), proj.AV.get(), tmp.get(), static_cast<size_t>(op_size_ * rs)));
      dev_->gemm('N', 'N', op_size_, rs, dim, 1.0, proj.AAV.get(), op_size_, Q.get(), dim, 0.0, tmp.get(), op_size_);
      dev_->check(gwbse_d2d(dev_->ctx(), proj.AAV.get(), tmp.get(), static_cast<size_t>(op_size_ * rs)));
      dev_->check(gwbse_sync(dev_->ctx()));
      proj.B = project(newV, 0, rs, proj.AAV, 0, rs);
    }
    dev_->check(gwbse_sync(dev_->ctx()));
    proj.T = project(newV, 0, rs, proj.AV, 0, rs);
    proj.V = std::move(newV);
    proj.ncols = rs + newvectors;
    proj.nav = rs;
  }

  // DavidsonSolver::qr, davidsonsolver.cc:480-488: thin Q of a Householder QR
  static MatrixXd qr(const MatrixXd& A) {
    const Index m = A.rows(), n = std::min(A.rows(), A.cols());
    MatrixXd R = A;
    std::vector<VectorXd> vs;
    for (Index k = 0; k < n; ++k) {
      VectorXd v(m, 0.0);
      double norm = 0.0;
      for (Index i = k; i < m; ++i) norm += R(i, k) * R(i, k);
      norm = std::sqrt(norm);
      if (norm == 0.0) {
        vs.push_back(v);
        continue;
      }
      const double alpha = R(k, k) > 0 ? -norm : norm;
      for (Index i = k; i < m; ++i) v(i) = R(i, k);
      v(k) -= alpha;
      double vn = 0.0;
      for (Index i = k; i < m; ++i) vn += v(i) * v(i);
      vn = std::sqrt(vn);
      if (vn > 0)
        for (Index i = k; i < m; ++i) v(i) /= vn;
      for (Index j = k; j < R.cols(); ++j) {
        double d = 0.0;
        for (Index i = k; i < m; ++i) d += v(i) * R(i, j);
        for (Index i = k; i < m; ++i) R(i, j) -= 2.0 * d * v(i);
      }
      vs.push_back(v);
    }
    MatrixXd Q = MatrixXd::Identity(m, n);
    for (Index k = n - 1; k >= 0; --k)
      for (Index j = 0; j < n; ++j) {
        double d = 0.0;
        for (Index i = k; i < m; ++i) d += vs[k](i) * Q(i, j);
        for (Index i = k; i < m; ++i) Q(i, j) -= 2.0 * d * vs[k](i);
      }
    return Q;
  }

  void storeConvergedData(const RitzEigenPair& rep, Index neigen) {
    storeEigenPairs(rep, neigen);
    log_(" Davidson converged after " + std::to_string(i_iter_) + " iterations.");
    success_ = true;
  }
  void storeNotConvergedData(const RitzEigenPair& rep, const std::vector<char>& root_converged, Index neigen) {
    storeEigenPairs(rep, neigen);
    double percent_converged = 0;
    for (Index i = 0; i < neigen; i++) {
      if (!root_converged[i]) {
        eigenvalues_(i) = 0;
        for (Index r = 0; r < eigenvectors_.rows(); ++r) eigenvectors_(r, i) = 0.0;
      } else {
        percent_converged += 1.;
      }
    }
    percent_converged = 100. * percent_converged / double(neigen);
    char buf[128];
    std::snprintf(buf, sizeof(buf), "- Warning : Davidson %5.2f%% converged after %ld iterations.", percent_converged,
                  (long)i_iter_);
    log_(buf);
    success_ = false;
  }
  // davidsonsolver.cc:549-555
  void storeEigenPairs(const RitzEigenPair& rep, Index neigen) {
    eigenvalues_ = rep.lambda.head(neigen);
    eigenvectors_ = dev_->download(rep.q.get(), op_size_, neigen);
    for (Index j = 0; j < neigen; ++j) {
      double n = 0.0;
      for (Index r = 0; r < op_size_; ++r) n += eigenvectors_(r, j) * eigenvectors_(r, j);
      n = std::sqrt(n);
      for (Index r = 0; r < op_size_; ++r) eigenvectors_(r, j) /= n;
    }
  }

  // davidsonsolver.cc:149-170
  Index getSizeUpdate(Index neigen) const {
    switch (davidson_update_) {
      case UPDATE::MIN: return neigen;
      case UPDATE::SAFE: return neigen < 20 ? static_cast<Index>(1.5 * double(neigen)) : neigen + 10;
      default: return 2 * neigen;
    }
  }
  std::vector<Index> argsort(const VectorXd& V) const {
    std::vector<Index> idx(V.size());
    std::iota(idx.begin(), idx.end(), 0);
    std::sort(idx.begin(), idx.end(), [&](Index i1, Index i2) { return V[i1] < V[i2]; });
    return idx;
  }
  // davidsonsolver.cc:180-206
  MatrixXd setupInitialEigenvectors(Index size_initial_guess) const {
    MatrixXd guess = MatrixXd::Zero(Adiag_.size(), size_initial_guess);
    std::vector<Index> idx = argsort(Adiag_);
    if (matrix_type_ == MATRIX_TYPE::SYMM) {
      for (Index j = 0; j < size_initial_guess; j++) guess(idx[j], j) = 1.0;
    } else {
      Index ind0 = Adiag_.size() / 2;
      for (Index j = 0; j < size_initial_guess; j++) guess(idx[ind0 + j], j) = 1.0;
    }
    return guess;
  }
  // davidsonsolver.cc:44-63
  void checkOptions(Index operator_size) {
    if (max_search_space_ > operator_size) {
      log_(" == Warning : Max search space (" + std::to_string(max_search_space_) + ") larger than system size (" +
           std::to_string(operator_size) + ")");
      max_search_space_ = operator_size;
      log_(" == Warning : Max search space set to " + std::to_string(operator_size));
      log_(" == Warning : If problems appear, try asking for less than " + std::to_string(operator_size / 10) +
           " eigenvalues");
    }
  }
  void printOptions(Index operator_size) const {
    log_(" Davidson Solver on the GPU (device-resident search space).");
    log_(" Tolerance : " + std::to_string(tol_));
    log_(davidson_correction_ == CORR::DPR ? " DPR Correction" : " Olsen Correction");
    log_(" Matrix size : " + std::to_string(operator_size) + "x" + std::to_string(operator_size));
  }
  void printIterationData(const RitzEigenPair& rep, const ProjectedSpace& proj, Index neigen) const {
    Index converged_roots = 0;
    double maxres = 0.0;
    for (Index j = 0; j < neigen; ++j) {
      converged_roots += proj.root_converged[j] ? 1 : 0;
      maxres = std::max(maxres, rep.res_norm(j));
    }
    char buf[160];
    std::snprintf(buf, sizeof(buf), " %4ld %12ld \t %4.2e \t %5.2f%% converged", (long)i_iter_,
                  (long)proj.search_space(), maxres, 100.0 * double(converged_roots) / double(neigen));
    log_(buf);
  }
  void printTiming() const {
    std::chrono::duration<double> el = std::chrono::system_clock::now() - start_;
    log_("-----------------------------------");
    log_("- Davidson ran for " + std::to_string(el.count()) + "secs.");
    log_("-----------------------------------");
  }

  Logger& log_;
  const Device* dev_ = nullptr;
  Index iter_max_ = 50;
  Index i_iter_ = 0;
  double tol_ = 1E-4;
  Index max_search_space_ = 0;
  Index op_size_ = 0, capacity_ = 0;
  Index op_columns_ = 0;
  VectorXd Adiag_;
  Device::Buffer Adiag_dev_;
  Index restart_size_ = 0;
  CORR davidson_correction_ = CORR::DPR;
  UPDATE davidson_update_ = UPDATE::SAFE;
  MATRIX_TYPE matrix_type_ = MATRIX_TYPE::SYMM;
  VectorXd eigenvalues_;
  MatrixXd eigenvectors_;
  bool success_ = false;
  std::chrono::time_point<std::chrono::system_clock> start_;
};

}  // namespace xtp
}  // namespace votca
