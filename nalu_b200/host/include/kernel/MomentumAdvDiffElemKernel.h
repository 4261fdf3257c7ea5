// include/kernel/MomentumAdvDiffElemKernel.h:27-66, src/kernel/MomentumAdvDiffElemKernel.C:26-64
#pragma once
#include "Kernel.h"
namespace sierra {
namespace nalu {
/** Advection diffusion term for momentum equation (velocity DOF) */
template <typename AlgTraits>
class MomentumAdvDiffElemKernel : public Kernel {
public:
  MomentumAdvDiffElemKernel(const stk::mesh::BulkData &, const SolutionOptions &, VectorFieldType *, ScalarFieldType *,
                            ElemDataRequests &);
  virtual ~MomentumAdvDiffElemKernel() {}
  bool device_descriptor(nalu_b200_kernel_desc &) override;
  const char *name() const override { return "MomentumAdvDiffElemKernel"; }
  // v_shape_function_ / v_adv_shape_function_ of the reference (filled, not used on the host)
  double v_shape_function_[AlgTraits::numScsIp_][AlgTraits::nodesPerElement_];
  double v_adv_shape_function_[AlgTraits::numScsIp_][AlgTraits::nodesPerElement_];
private:
  MomentumAdvDiffElemKernel() = delete;
  VectorFieldType *velocityNp1_{nullptr};
  VectorFieldType *coordinates_{nullptr};
  ScalarFieldType *viscosity_{nullptr};
  GenericFieldType *massFlowRate_{nullptr};
  const int *lrscv_;
  const double includeDivU_;
  const bool shiftedGradOp_;
  bool skewSymmetric_;
};
template <>
struct has_hex27_instantiation<MomentumAdvDiffElemKernel> : std::true_type {};
} // namespace nalu
} // namespace sierra
