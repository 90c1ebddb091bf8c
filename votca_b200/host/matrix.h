// Minimal column-major dense containers for the host layer.
//
// The reference's interfaces take Eigen::MatrixXd / Eigen::VectorXd (xtp/include/votca/xtp/eigen.h).
// Eigen is not available in this build environment, so the host mirror uses these containers, which have
// the same memory layout (column-major, contiguous, data()/rows()/cols()) - an Eigen::Map<MatrixXd> over
// data() is a zero-copy view in either direction.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <stdexcept>
#include <vector>

namespace votca {
using Index = long;
namespace xtp {

class VectorXd {
 public:
  VectorXd() = default;
  explicit VectorXd(Index n, double v = 0.0) : d_(static_cast<size_t>(n), v) {}
  VectorXd(const double* p, Index n) : d_(p, p + n) {}
  static VectorXd Zero(Index n) { return VectorXd(n, 0.0); }
  Index size() const { return static_cast<Index>(d_.size()); }
  double& operator()(Index i) { return d_[static_cast<size_t>(i)]; }
  double operator()(Index i) const { return d_[static_cast<size_t>(i)]; }
  double& operator[](Index i) { return d_[static_cast<size_t>(i)]; }
  double operator[](Index i) const { return d_[static_cast<size_t>(i)]; }
  double* data() { return d_.data(); }
  const double* data() const { return d_.data(); }
  void resize(Index n) { d_.assign(static_cast<size_t>(n), 0.0); }
  VectorXd segment(Index start, Index n) const { return VectorXd(d_.data() + start, n); }
  VectorXd head(Index n) const { return segment(0, n); }
  double sum() const {
    double s = 0;
    for (double v : d_) s += v;
    return s;
  }
  double maxAbs() const {
    double s = 0;
    for (double v : d_) s = std::max(s, std::abs(v));
    return s;
  }
  double dot(const VectorXd& o) const {
    double s = 0;
    for (size_t i = 0; i < d_.size(); ++i) s += d_[i] * o.d_[i];
    return s;
  }
  VectorXd& operator+=(const VectorXd& o) {
    for (size_t i = 0; i < d_.size(); ++i) d_[i] += o.d_[i];
    return *this;
  }
  VectorXd& operator-=(const VectorXd& o) {
    for (size_t i = 0; i < d_.size(); ++i) d_[i] -= o.d_[i];
    return *this;
  }
  VectorXd& operator*=(double a) {
    for (double& v : d_) v *= a;
    return *this;
  }
  friend VectorXd operator+(VectorXd a, const VectorXd& b) { return a += b; }
  friend VectorXd operator-(VectorXd a, const VectorXd& b) { return a -= b; }
  friend VectorXd operator*(double s, VectorXd a) { return a *= s; }

 private:
  std::vector<double> d_;
};

class MatrixXd {
 public:
  MatrixXd() = default;
  MatrixXd(Index r, Index c, double v = 0.0) : r_(r), c_(c), d_(static_cast<size_t>(r * c), v) {}
  MatrixXd(const double* p, Index r, Index c, Index ld) : r_(r), c_(c), d_(static_cast<size_t>(r * c)) {
    for (Index j = 0; j < c; ++j) std::copy(p + j * ld, p + j * ld + r, d_.begin() + j * r);
  }
  static MatrixXd Zero(Index r, Index c) { return MatrixXd(r, c, 0.0); }
  static MatrixXd Identity(Index r, Index c) {
    MatrixXd m(r, c, 0.0);
    for (Index i = 0; i < std::min(r, c); ++i) m(i, i) = 1.0;
    return m;
  }
  Index rows() const { return r_; }
  Index cols() const { return c_; }
  Index size() const { return r_ * c_; }
  double& operator()(Index i, Index j) { return d_[static_cast<size_t>(i + j * r_)]; }
  double operator()(Index i, Index j) const { return d_[static_cast<size_t>(i + j * r_)]; }
  double* data() { return d_.data(); }
  const double* data() const { return d_.data(); }
  double* colptr(Index j) { return d_.data() + j * r_; }
  const double* colptr(Index j) const { return d_.data() + j * r_; }
  void resize(Index r, Index c) {
    r_ = r;
    c_ = c;
    d_.assign(static_cast<size_t>(r * c), 0.0);
  }
  VectorXd diagonal() const {
    VectorXd v(std::min(r_, c_));
    for (Index i = 0; i < v.size(); ++i) v(i) = (*this)(i, i);
    return v;
  }
  VectorXd col(Index j) const { return VectorXd(colptr(j), r_); }
  MatrixXd block(Index i0, Index j0, Index nr, Index nc) const {
    MatrixXd b(nr, nc);
    for (Index j = 0; j < nc; ++j)
      for (Index i = 0; i < nr; ++i) b(i, j) = (*this)(i0 + i, j0 + j);
    return b;
  }
  void setBlock(Index i0, Index j0, const MatrixXd& b) {
    for (Index j = 0; j < b.cols(); ++j)
      for (Index i = 0; i < b.rows(); ++i) (*this)(i0 + i, j0 + j) = b(i, j);
  }
  MatrixXd leftCols(Index n) const { return block(0, 0, r_, n); }
  MatrixXd transpose() const {
    MatrixXd t(c_, r_);
    for (Index j = 0; j < c_; ++j)
      for (Index i = 0; i < r_; ++i) t(j, i) = (*this)(i, j);
    return t;
  }
  double trace() const { return diagonal().sum(); }
  // small host product (control-plane sized matrices only; the contractions live on the GPU)
  MatrixXd operator*(const MatrixXd& o) const {
    if (c_ != o.r_) throw std::runtime_error("host matrix product: shape mismatch");
    MatrixXd out(r_, o.c_);
    for (Index j = 0; j < o.c_; ++j)
      for (Index k = 0; k < c_; ++k) {
        const double b = o(k, j);
        if (b == 0.0) continue;
        const double* a = colptr(k);
        double* y = out.colptr(j);
        for (Index i = 0; i < r_; ++i) y[i] += a[i] * b;
      }
    return out;
  }
  MatrixXd& operator+=(const MatrixXd& o) {
    for (size_t i = 0; i < d_.size(); ++i) d_[i] += o.d_[i];
    return *this;
  }
  MatrixXd& operator-=(const MatrixXd& o) {
    for (size_t i = 0; i < d_.size(); ++i) d_[i] -= o.d_[i];
    return *this;
  }
  MatrixXd& operator*=(double a) {
    for (double& v : d_) v *= a;
    return *this;
  }
  friend MatrixXd operator+(MatrixXd a, const MatrixXd& b) { return a += b; }
  friend MatrixXd operator-(MatrixXd a, const MatrixXd& b) { return a -= b; }
  friend MatrixXd operator*(double s, MatrixXd a) { return a *= s; }

 private:
  Index r_ = 0, c_ = 0;
  std::vector<double> d_;
};

inline MatrixXd asDiagonal(const VectorXd& v) {
  MatrixXd m(v.size(), v.size());
  for (Index i = 0; i < v.size(); ++i) m(i, i) = v(i);
  return m;
}

}  // namespace xtp
}  // namespace votca
