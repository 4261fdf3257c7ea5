// MomentumNSOElemKernel -- host side of the reference's include/nso/MomentumNSOElemKernel.h /
// src/nso/MomentumNSOElemKernel.C:27-125: same constructor signature (velocity, Gju, viscosity, fourthFac, altResFac),
// same ElemDataRequests registrations (SCS_AREAV, SCS_(SHIFTED_)GRAD_OP, SCS_GIJ); execute() is lowered to a device
// descriptor.
#pragma once
#include "../kernel/Kernel.h"
namespace sierra {
namespace nalu {

template <typename AlgTraits>
class MomentumNSOElemKernel : public Kernel {
public:
  MomentumNSOElemKernel(const stk::mesh::BulkData &, const SolutionOptions &, VectorFieldType *, GenericFieldType *, ScalarFieldType *,
                        const double fourthFac, const double altResFac, ElemDataRequests &);
  virtual ~MomentumNSOElemKernel() {}
  void setup(const TimeIntegrator &) override;
  bool device_descriptor(nalu_b200_kernel_desc &) override;
  const char *name() const override { return "MomentumNSOElemKernel"; }
  double v_shape_function_[AlgTraits::numScsIp_][AlgTraits::nodesPerElement_];
private:
  MomentumNSOElemKernel() = delete;
  ScalarFieldType *viscosity_{nullptr};
  GenericFieldType *Gju_{nullptr};
  VectorFieldType *velocityNm1_{nullptr}, *velocityN_{nullptr}, *velocityNp1_{nullptr}, *velocityRTM_{nullptr}, *coordinates_{nullptr};
  ScalarFieldType *densityNm1_{nullptr}, *densityN_{nullptr}, *densityNp1_{nullptr}, *pressure_{nullptr};
  const int *lrscv_;
  double dt_{0.0}, gamma1_{0.0}, gamma2_{0.0}, gamma3_{0.0};
  const double Cupw_{0.1};
  const double fourthFac_, altResFac_, om_altResFac_;
  const double nonConservedForm_{0.0};
  const double includeDivU_;
  const double small_{1.0e-16};
  const bool shiftedGradOp_;
};

} // namespace nalu
} // namespace sierra
