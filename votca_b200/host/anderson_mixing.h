// Anderson mixing of the evGW / QSGW fixed-point iteration - host mirror of
// xtp/src/libxtp/anderson_mixing.cc:28-95 (xtp/include/votca/xtp/anderson_mixing.h).  No device dependency.
#pragma once
#include <cmath>
#include <vector>

#include "matrix.h"

namespace votca {
namespace xtp {

// anderson_mixing.cc:28-95
class Anderson {
 public:
  void Configure(const Index order, const double alpha) {
    order_ = order + 1;
    alpha_ = alpha;
  }
  void UpdateOutput(const VectorXd& newOutput) {
    if (Index(output_.size()) > order_ - 1) output_.erase(output_.begin());
    output_.push_back(newOutput);
  }
  void UpdateInput(const VectorXd& newInput) {
    if (Index(output_.size()) > order_ - 1) input_.erase(input_.begin());
    input_.push_back(newInput);
  }
  const VectorXd MixHistory() {
    const Index iteration = output_.size();
    const Index used_history = iteration - 1;
    VectorXd OutMixed = output_.back();
    VectorXd InMixed = input_.back();
    if (iteration > 1 && order_ > 1) {
      VectorXd DeltaN = OutMixed - InMixed;
      MatrixXd A(used_history, used_history);
      VectorXd c(used_history);
      for (Index m = 1; m < iteration; m++) {
        const VectorXd dm = DeltaN - output_[used_history - m] + input_[used_history - m];
        c(m - 1) = dm.dot(DeltaN);
        for (Index j = 1; j < iteration; j++)
          A(m - 1, j - 1) = dm.dot(DeltaN - output_[used_history - j] + input_[used_history - j]);
      }
      VectorXd coefficients = SolveFullPivQR(A, c);
      for (Index n = 1; n < iteration; n++) {
        OutMixed += coefficients(n - 1) * (output_[used_history - n] - output_[used_history]);
        InMixed += coefficients(n - 1) * (input_[used_history - n] - input_[used_history]);
      }
    }
    return alpha_ * OutMixed + (1 - alpha_) * InMixed;
  }

  // rank-revealing solve standing in for Eigen's fullPivHouseholderQr().solve (anderson_mixing.cc:72):
  // Gaussian elimination with complete pivoting, free variables of a rank-deficient system set to zero.
  static VectorXd SolveFullPivQR(MatrixXd A, VectorXd b) {
    const Index n = A.rows();
    std::vector<Index> colperm(n);
    for (Index i = 0; i < n; ++i) colperm[i] = i;
    double maxpiv = 0.0;
    Index rank = 0;
    for (Index k = 0; k < n; ++k) {
      Index pi = k, pj = k;
      double best = 0.0;
      for (Index j = k; j < n; ++j)
        for (Index i = k; i < n; ++i)
          if (std::abs(A(i, j)) > best) {
            best = std::abs(A(i, j));
            pi = i;
            pj = j;
          }
      if (k == 0) maxpiv = best;
      if (best <= maxpiv * 1e-14 * double(n) || best == 0.0) break;
      for (Index j = 0; j < n; ++j) std::swap(A(k, j), A(pi, j));
      std::swap(b(k), b(pi));
      for (Index i = 0; i < n; ++i) std::swap(A(i, k), A(i, pj));
      std::swap(colperm[k], colperm[pj]);
      for (Index i = k + 1; i < n; ++i) {
        const double f = A(i, k) / A(k, k);
        for (Index j = k; j < n; ++j) A(i, j) -= f * A(k, j);
        b(i) -= f * b(k);
      }
      ++rank;
    }
    VectorXd y(n, 0.0);
    for (Index k = rank - 1; k >= 0; --k) {
      double s = b(k);
      for (Index j = k + 1; j < rank; ++j) s -= A(k, j) * y(j);
      y(k) = s / A(k, k);
    }
    VectorXd x(n, 0.0);
    for (Index k = 0; k < n; ++k) x(colperm[k]) = y(k);
    return x;
  }

 private:
  std::vector<VectorXd> input_, output_;
  double alpha_ = 0.7;
  Index order_ = 25;
};

}  // namespace xtp
}  // namespace votca
