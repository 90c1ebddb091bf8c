// BSE_OPERATOR<cqp,cx,cd,cd2> and HamiltonianOperator<A,B> - host mirror of
// xtp/include/votca/xtp/bse_operator.h:32-87, matrixfreeoperator.h:45-70 and
// bseoperator_btda.h:59-149.  matmul() has the reference signature (host in, host out); apply_dev() is the
// device-resident entry the Davidson solver uses so trial vectors never cross PCIe.
#pragma once
#include "threecenter.h"

namespace votca {
namespace xtp {

struct BSEOperator_Options {
  Index homo;
  Index rpamin;
  Index qpmin;
  Index vmin;
  Index cmax;
};

class MatrixFreeOperator {
 public:
  virtual ~MatrixFreeOperator() = default;
  Index rows() const { return size_; }
  Index cols() const { return size_; }
  Index size() const { return size_; }
  void set_size(Index size) { size_ = size; }
  virtual VectorXd diagonal() const = 0;
  virtual MatrixXd matmul(const MatrixXd& input) const = 0;
  // Y_dev(size x k) = H * X_dev(size x k)
  virtual void apply_dev(const double* X_dev, Index ldx, Index k, double* Y_dev, Index ldy) const = 0;
  virtual const Device& device() const = 0;
  MatrixXd operator*(const MatrixXd& x) const { return matmul(x); }

 private:
  Index size_ = 0;
};

template <Index cqp, Index cx, Index cd, Index cd2>
class BSE_OPERATOR final : public MatrixFreeOperator {
 public:
  BSE_OPERATOR(const VectorXd& Hd_operator, const TCMatrix_gwbse& Mmn, const MatrixXd& Hqp)
      : epsilon_0_inv_(Hd_operator), Mmn_(Mmn), Hqp_(Hqp) {
    static_assert(!(cd2 != 0 && cd != 0), "Hamiltonian cannot contain Hd and Hd2 at the same time");
  }

  // bse_operator.cc:29-38
  void configure(BSEOperator_Options opt) {
    opt_ = opt;
    Index bse_vmax = opt_.homo;
    bse_cmin_ = opt_.homo + 1;
    bse_vtotal_ = bse_vmax - opt_.vmin + 1;
    bse_ctotal_ = opt_.cmax - bse_cmin_ + 1;
    bse_size_ = bse_vtotal_ * bse_ctotal_;
    this->set_size(bse_size_);
  }

  VectorXd diagonal() const override {
    bind();
    VectorXd d(bse_size_);
    device().check(gwbse_bse_diagonal(device().ctx(), (int)cqp, (int)cx, (int)cd, (int)cd2, d.data()));
    return d;
  }

  MatrixXd matmul(const MatrixXd& input) const override {
    if (input.rows() != bse_size_) throw std::runtime_error("Shape mismatch in BSE_OPERATOR::matmul");
    bind();
    MatrixXd out(bse_size_, input.cols());
    device().check(gwbse_bse_matmul(device().ctx(), (int)cqp, (int)cx, (int)cd, (int)cd2, (int)input.cols(),
                                    input.data(), (int)input.rows(), out.data(), (int)out.rows()));
    return out;
  }

  void apply_dev(const double* X_dev, Index ldx, Index k, double* Y_dev, Index ldy) const override {
    bind();
    device().check(gwbse_bse_matmul_dev(device().ctx(), (int)cqp, (int)cx, (int)cd, (int)cd2, (int)k, X_dev, (int)ldx,
                                        Y_dev, (int)ldy));
  }
  const Device& device() const override { return Mmn_.device(); }

 private:
  // the device holds one (eps_inv, Hqp, ranges) binding at a time; (re)bind before use
  void bind() const {
    const Index hs = bse_vtotal_ + bse_ctotal_;
    if (Hqp_.rows() < hs || Hqp_.cols() < hs) throw std::runtime_error("Hqp is smaller than the BSE window");
    device().check(gwbse_bse_configure(device().ctx(), (int)opt_.homo, (int)opt_.rpamin, (int)opt_.vmin,
                                       (int)opt_.cmax, epsilon_0_inv_.data(), Hqp_.data(), (int)Hqp_.rows()));
  }
  BSEOperator_Options opt_;
  Index bse_size_ = 0, bse_vtotal_ = 0, bse_ctotal_ = 0, bse_cmin_ = 0;
  const VectorXd& epsilon_0_inv_;
  const TCMatrix_gwbse& Mmn_;
  const MatrixXd& Hqp_;
};

typedef BSE_OPERATOR<1, 2, 1, 0> SingletOperator_TDA;
typedef BSE_OPERATOR<1, 0, 1, 0> TripletOperator_TDA;
typedef BSE_OPERATOR<0, 2, 0, 1> SingletOperator_BTDA_B;
typedef BSE_OPERATOR<1, 0, 0, 0> HqpOperator;
typedef BSE_OPERATOR<0, 1, 0, 0> HxOperator;
typedef BSE_OPERATOR<0, 0, 1, 0> HdOperator;
typedef BSE_OPERATOR<0, 0, 0, 1> Hd2Operator;

// [A B; -B -A] * [X; Y], bseoperator_btda.h:116-149
template <typename MatrixReplacementA, typename MatrixReplacementB>
class HamiltonianOperator {
 public:
  HamiltonianOperator(const MatrixReplacementA& A, const MatrixReplacementB& B) : A_(A), B_(B) {
    size_ = 2 * A.cols();
    VectorXd d = A_.diagonal();
    diag_ = VectorXd(size_);
    for (Index i = 0; i < size_ / 2; ++i) {
      diag_(i) = d(i);
      diag_(i + size_ / 2) = -d(i);
    }
  }
  Index rows() const { return size_; }
  Index cols() const { return size_; }
  VectorXd diagonal() const { return diag_; }
  const Device& device() const { return A_.device(); }

  void apply_dev(const double* X_dev, Index ldx, Index k, double* Y_dev, Index ldy) const {
    const Device& dev = device();
    const Index half = size_ / 2;
    // the halves of every column, side by side: (half x 2k) with the column pairs interleaved, exactly the
    // Map reshape of bseoperator_btda.h:139 when ldx == size
    Device::Buffer in = dev.alloc(static_cast<size_t>(half * 2 * k));
    Device::Buffer tA = dev.alloc(static_cast<size_t>(half * 2 * k));
    Device::Buffer tB = dev.alloc(static_cast<size_t>(half * 2 * k));
    for (Index j = 0; j < k; ++j) {
      dev.check(gwbse_d2d(dev.ctx(), in.get() + (2 * j) * half, X_dev + j * ldx, (size_t)half));
      dev.check(gwbse_d2d(dev.ctx(), in.get() + (2 * j + 1) * half, X_dev + j * ldx + half, (size_t)half));
    }
    A_.apply_dev(in.get(), half, 2 * k, tA.get(), half);
    B_.apply_dev(in.get(), half, 2 * k, tB.get(), half);
    // top = A Xt + B Xb ; bottom = -A Xb - B Xt
    for (Index j = 0; j < k; ++j) {
      double* top = Y_dev + j * ldy;
      double* bot = top + half;
      dev.check(gwbse_d2d(dev.ctx(), top, tA.get() + (2 * j) * half, (size_t)half));
      dev.check(gwbse_axpy_dev(dev.ctx(), (int)half, 1, 1.0, tB.get() + (2 * j + 1) * half, (int)half, top, (int)half));
      dev.check(gwbse_dev_memset_zero(dev.ctx(), bot, (size_t)half));
      dev.check(gwbse_axpy_dev(dev.ctx(), (int)half, 1, -1.0, tA.get() + (2 * j + 1) * half, (int)half, bot, (int)half));
      dev.check(gwbse_axpy_dev(dev.ctx(), (int)half, 1, -1.0, tB.get() + (2 * j) * half, (int)half, bot, (int)half));
    }
    dev.check(gwbse_sync(dev.ctx()));
  }

  MatrixXd matmul(const MatrixXd& x) const {
    const Device& dev = device();
    Device::Buffer X = dev.upload(x);
    Device::Buffer Y = dev.alloc(static_cast<size_t>(x.size()));
    apply_dev(X.get(), x.rows(), x.cols(), Y.get(), x.rows());
    return dev.download(Y.get(), x.rows(), x.cols());
  }
  MatrixXd operator*(const MatrixXd& x) const { return matmul(x); }

  const MatrixReplacementA& A_;
  const MatrixReplacementB& B_;

 private:
  Index size_;
  VectorXd diag_;
};

}  // namespace xtp
}  // namespace votca
