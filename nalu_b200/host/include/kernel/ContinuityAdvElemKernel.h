// include/kernel/ContinuityAdvElemKernel.h, src/kernel/ContinuityAdvElemKernel.C:26-84
#pragma once
#include "Kernel.h"
namespace sierra {
namespace nalu {
template <typename AlgTraits>
class ContinuityAdvElemKernel : public Kernel {
public:
  ContinuityAdvElemKernel(const stk::mesh::BulkData &, const SolutionOptions &, ElemDataRequests &);
  virtual ~ContinuityAdvElemKernel() {}
  void setup(const TimeIntegrator &) override;
  bool device_descriptor(nalu_b200_kernel_desc &) override;
  const char *name() const override { return "ContinuityAdvElemKernel"; }
  double v_shape_function_[AlgTraits::numScsIp_][AlgTraits::nodesPerElement_];
private:
  ContinuityAdvElemKernel() = delete;
  VectorFieldType *velocityRTM_{nullptr};
  VectorFieldType *Gpdx_{nullptr};
  ScalarFieldType *pressure_{nullptr};
  ScalarFieldType *densityNp1_{nullptr};
  VectorFieldType *coordinates_{nullptr};
  double projTimeScale_{1.0};
  double dt_{1.0}, gamma1_{1.0};
  const bool meshMotion_, shiftMdot_, shiftPoisson_, reducedSensitivities_;
  const int *lrscv_;
};
template <>
struct has_hex27_instantiation<ContinuityAdvElemKernel> : std::true_type {};
} // namespace nalu
} // namespace sierra
