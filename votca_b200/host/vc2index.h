// vc2index - compound transition index of the BSE blocks (the public helper xtp/include/votca/xtp/vc2index.h:36-54
// provides under this name): I = ctotal * (v - vmin) + (c - cmin), c running fastest.  It is the row order of every
// B x k trial / product block of BSE_OPERATOR::matmul (bse_operator.cc:40-119) and of the device-side factorised
// products (capi_bse.cu).  Same constructor and I / v / c accessors; kept as one precomputed origin.
#pragma once
#include <utility>

#include "matrix.h"

namespace votca {
namespace xtp {

class vc2index {
 public:
  vc2index(Index vmin, Index cmin, Index ctotal) : width_(ctotal), vfirst_(vmin), cfirst_(cmin) {}

  // (v, c) -> I
  Index I(Index v, Index c) const { return (v - vfirst_) * width_ + (c - cfirst_); }
  // I -> v, I -> c, and both at once
  Index v(Index index) const { return vfirst_ + index / width_; }
  Index c(Index index) const { return cfirst_ + index % width_; }
  std::pair<Index, Index> vc(Index index) const { return {v(index), c(index)}; }

 private:
  Index width_;            // number of virtual levels: the stride of v in the compound index
  Index vfirst_, cfirst_;  // lowest occupied / virtual level of the block
};

}  // namespace xtp
}  // namespace votca
