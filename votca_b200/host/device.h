// RAII access to the C ABI from the C++ host layer.  Every non-zero status is rethrown as
// std::runtime_error carrying gwbse_last_error(), preserving the reference's throw-on-error convention
// (xtp/src/libxtp/cudamatrix.cc:25-37).
#pragma once
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>

#include "../../include/gwbse_b200.h"
#include "matrix.h"

namespace votca {
namespace xtp {

// Logger stand-in for votca::xtp::Logger + XTP_LOG (xtp/include/votca/xtp/logger.h:37-41): collects the
// same messages (the "... calculation took X seconds" lines xtp_benchmark parses) in a string buffer.
class Logger {
 public:
  void operator()(const std::string& line) {
    buf_ << line << '\n';
  }
  std::string str() const { return buf_.str(); }

 private:
  std::ostringstream buf_;
};

class Device {
 public:
  explicit Device(int device = 0) : index_(device) {
    if (gwbse_ctx_create(device, &ctx_) != 0) throw std::runtime_error(gwbse_create_error());
  }
  ~Device() { gwbse_ctx_destroy(ctx_); }
  Device(const Device&) = delete;
  Device& operator=(const Device&) = delete;
  int index() const { return index_; }  // CUDA device ordinal: a second context on the same GPU (UKS beta channel)
  gwbse_ctx* ctx() const { return ctx_; }
  // multi-GPU: one process per GPU, Mmn sharded m-cyclically (gwbse_comm_init)
  int rank() const { return gwbse_comm_rank(ctx_); }
  int world() const { return gwbse_comm_world(ctx_); }
  bool owns_slice(Index m) const { return gwbse_shard_owner((int)m, world()) == rank(); }
  void allreduce(double* buf, size_t n) const { check(gwbse_comm_allreduce_host(ctx_, buf, n)); }
  void check(int rc) const {
    if (rc != 0) throw std::runtime_error(gwbse_last_error(ctx_));
  }
  void sync() const { check(gwbse_sync(ctx_)); }

  // device buffer with RAII (CudaMatrix analogue, cudamatrix.h:95-190)
  class Buffer {
   public:
    Buffer() = default;
    Buffer(const Device& d, size_t n) : dev_(&d), n_(n) { d.check(gwbse_dev_malloc(d.ctx(), n * sizeof(double), &p_)); }
    Buffer(Buffer&& o) noexcept : dev_(o.dev_), p_(o.p_), n_(o.n_) { o.p_ = nullptr; }
    Buffer& operator=(Buffer&& o) noexcept {
      release();
      dev_ = o.dev_;
      p_ = o.p_;
      n_ = o.n_;
      o.p_ = nullptr;
      return *this;
    }
    ~Buffer() { release(); }
    double* get() const { return p_; }
    size_t size() const { return n_; }

   private:
    void release() {
      if (p_ && dev_) gwbse_dev_free(dev_->ctx(), p_);
      p_ = nullptr;
    }
    const Device* dev_ = nullptr;
    double* p_ = nullptr;
    size_t n_ = 0;
  };

  Buffer alloc(size_t n) const { return Buffer(*this, n); }
  Buffer upload(const MatrixXd& m) const {
    Buffer b(*this, static_cast<size_t>(std::max<Index>(m.size(), 1)));
    if (m.size()) check(gwbse_h2d(ctx_, b.get(), m.data(), static_cast<size_t>(m.size())));
    return b;
  }
  Buffer upload(const VectorXd& v) const {
    Buffer b(*this, static_cast<size_t>(std::max<Index>(v.size(), 1)));
    if (v.size()) check(gwbse_h2d(ctx_, b.get(), v.data(), static_cast<size_t>(v.size())));
    return b;
  }
  MatrixXd download(const double* p, Index r, Index c) const {
    MatrixXd m(r, c);
    if (m.size()) check(gwbse_d2h(ctx_, m.data(), p, static_cast<size_t>(m.size())));
    return m;
  }
  void gemm(char ta, char tb, Index m, Index n, Index k, double alpha, const double* A, Index lda, const double* B,
            Index ldb, double beta, double* C, Index ldc) const {
    check(gwbse_dgemm_dev(ctx_, ta, tb, (int)m, (int)n, (int)k, alpha, A, (int)lda, B, (int)ldb, beta, C, (int)ldc));
  }
  // Eigen::SelfAdjointEigenSolver on the device: returns eigenvalues, overwrites A with eigenvectors
  VectorXd sym_eig(MatrixXd& A) const {
    const Index n = A.rows();
    Buffer d = upload(A);
    VectorXd w(n);
    check(gwbse_sym_eig_dev(ctx_, (int)n, d.get(), (int)n, w.data()));
    A = download(d.get(), n, n);
    return w;
  }

 private:
  int index_ = 0;
  gwbse_ctx* ctx_ = nullptr;
};

}  // namespace xtp
}  // namespace votca
