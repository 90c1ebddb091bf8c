// vc2index - compound transition index of the BSE blocks, as xtp/include/votca/xtp/vc2index.h:36-54:
// I = ctotal * (v - vmin) + (c - cmin); the layout every B x k trial / product block of BSE_OPERATOR::matmul uses
// (bse_operator.cc:40-119) and the row order of the device-side factorised products (capi_bse.cu).
#pragma once
#include "matrix.h"

namespace votca {
namespace xtp {

class vc2index {
 public:
  vc2index(Index vmin, Index cmin, Index ctotal) : vmin_(vmin), cmin_(cmin), ctotal_(ctotal) {}

  inline Index I(Index v, Index c) const { return ctotal_ * (v - vmin_) + (c - cmin_); }
  inline Index v(Index index) const { return index / ctotal_ + vmin_; }
  inline Index c(Index index) const { return index % ctotal_ + cmin_; }

 private:
  Index vmin_, cmin_, ctotal_;
};

}  // namespace xtp
}  // namespace votca
