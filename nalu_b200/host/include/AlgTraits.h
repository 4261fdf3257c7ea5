// AlgTraits.h -- include/AlgTraits.h:18-25 (only the Hex8 traits are on this path)
#pragma once
#include "stk_shim.h"
namespace sierra {
namespace nalu {
struct AlgTraitsHex8 {
  static constexpr int nDim_ = 3;
  static constexpr int nodesPerElement_ = 8;
  static constexpr int numScsIp_ = 12;
  static constexpr int numScvIp_ = 8;
  static constexpr int numGp_ = 8;
  static constexpr stk::topology::topology_t topo_ = stk::topology::HEX_8;
};
} // namespace nalu
} // namespace sierra
