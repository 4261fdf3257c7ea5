// AlgTraits.h -- include/AlgTraits.h:18-34 (the Hex8 traits, and Hex27 for the momentum kernels)
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
// include/AlgTraits.h:27-34: BASELINE config 5; MomentumAdvDiff, MomentumMass, ContinuityAdv and ScalarAdvDiff are instantiated for it
struct AlgTraitsHex27 {
  static constexpr int nDim_ = 3;
  static constexpr int nodesPerElement_ = 27;
  static constexpr int numScsIp_ = 216;
  static constexpr int numScvIp_ = 216;
  static constexpr int numGp_ = 27;
  static constexpr stk::topology::topology_t topo_ = stk::topology::HEX_27;
};
// include/AlgTraits.h:81-88 (AlgTraitsQuad4), :157-176 (AlgTraitsFaceElem), :199
struct AlgTraitsQuad4 {
  static constexpr int nDim_ = 3;
  static constexpr int nodesPerElement_ = 4;
  static constexpr int numScsIp_ = 4;
  static constexpr int numGp_ = 4;
  static constexpr stk::topology::topology_t topo_ = stk::topology::QUAD_4;
};
template <typename AlgTraitsFace, typename AlgTraitsElem>
struct AlgTraitsFaceElem {
  static constexpr int nDim_ = AlgTraitsElem::nDim_;
  static constexpr int nodesPerElement_ = AlgTraitsElem::nodesPerElement_;
  static constexpr int nodesPerFace_ = AlgTraitsFace::nodesPerElement_;
  static constexpr int numScsIp_ = AlgTraitsElem::numScsIp_;
  static constexpr int numFaceIp_ = AlgTraitsFace::numScsIp_;
  static constexpr stk::topology::topology_t elemTopo_ = AlgTraitsElem::topo_;
  static constexpr stk::topology::topology_t faceTopo_ = AlgTraitsFace::topo_;
};
using AlgTraitsQuad4Hex8 = AlgTraitsFaceElem<AlgTraitsQuad4, AlgTraitsHex8>;
} // namespace nalu
} // namespace sierra
