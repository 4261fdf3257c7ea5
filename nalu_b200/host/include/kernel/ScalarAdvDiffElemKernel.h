// include/kernel/ScalarAdvDiffElemKernel.h, src/kernel/ScalarAdvDiffElemKernel.C:26-63;
// ScalarDiffElemKernel: include/kernel/ScalarDiffElemKernel.h, src/kernel/ScalarDiffElemKernel.C:26-62
#pragma once
#include "Kernel.h"
namespace sierra {
namespace nalu {
template <typename AlgTraits>
class ScalarAdvDiffElemKernel : public Kernel {
public:
  ScalarAdvDiffElemKernel(const stk::mesh::BulkData &, const SolutionOptions &, ScalarFieldType *, ScalarFieldType *,
                          ElemDataRequests &);
  virtual ~ScalarAdvDiffElemKernel() {}
  bool device_descriptor(nalu_b200_kernel_desc &) override;
  const char *name() const override { return "ScalarAdvDiffElemKernel"; }
  double v_shape_function_[AlgTraits::numScsIp_][AlgTraits::nodesPerElement_];
  double v_adv_shape_function_[AlgTraits::numScsIp_][AlgTraits::nodesPerElement_];
private:
  ScalarAdvDiffElemKernel() = delete;
  ScalarFieldType *scalarQ_{nullptr};
  ScalarFieldType *diffFluxCoeff_{nullptr};
  VectorFieldType *coordinates_{nullptr};
  GenericFieldType *massFlowRate_{nullptr};
  const int *lrscv_;
  const bool shiftedGradOp_;
  bool skewSymmetric_;
};

template <typename AlgTraits>
class ScalarDiffElemKernel : public Kernel {
public:
  ScalarDiffElemKernel(const stk::mesh::BulkData &, const SolutionOptions &, ScalarFieldType *, ScalarFieldType *,
                       ElemDataRequests &);
  virtual ~ScalarDiffElemKernel() {}
  bool device_descriptor(nalu_b200_kernel_desc &) override;
  const char *name() const override { return "ScalarDiffElemKernel"; }
  double v_shape_function_[AlgTraits::numScsIp_][AlgTraits::nodesPerElement_];
private:
  ScalarDiffElemKernel() = delete;
  ScalarFieldType *scalarQ_{nullptr};
  ScalarFieldType *diffFluxCoeff_{nullptr};
  VectorFieldType *coordinates_{nullptr};
  const int *lrscv_;
  const bool shiftedGradOp_;
};
template <>
struct has_hex27_instantiation<ScalarAdvDiffElemKernel> : std::true_type {};
} // namespace nalu
} // namespace sierra
