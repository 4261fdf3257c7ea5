// TurbKineticEnergySSTSrcElemKernel / SpecificDissipationRateSSTSrcElemKernel -- host side of the reference's
// include/kernel/TurbKineticEnergySSTSrcElemKernel.h and SpecificDissipationRateSSTSrcElemKernel.h: same
// constructor signatures and ElemDataRequests registrations; execute() is lowered to a device descriptor.
#pragma once
#include "Kernel.h"
namespace sierra {
namespace nalu {

template <typename AlgTraits>
class TurbKineticEnergySSTSrcElemKernel : public Kernel {
public:
  TurbKineticEnergySSTSrcElemKernel(const stk::mesh::BulkData &, const SolutionOptions &, ElemDataRequests &, const bool lumpedMass);
  virtual ~TurbKineticEnergySSTSrcElemKernel() {}
  bool device_descriptor(nalu_b200_kernel_desc &) override;
  const char *name() const override { return "TurbKineticEnergySSTSrcElemKernel"; }
  double v_shape_function_[AlgTraits::numScvIp_][AlgTraits::nodesPerElement_];
private:
  TurbKineticEnergySSTSrcElemKernel() = delete;
  ScalarFieldType *tkeNp1_{nullptr}, *sdrNp1_{nullptr}, *densityNp1_{nullptr}, *tvisc_{nullptr};
  VectorFieldType *velocityNp1_{nullptr}, *coordinates_{nullptr};
  const bool lumpedMass_, shiftedGradOp_;
  const double betaStar_, tkeProdLimitRatio_;
  const int *ipNodeMap_;
};

template <typename AlgTraits>
class SpecificDissipationRateSSTSrcElemKernel : public Kernel {
public:
  SpecificDissipationRateSSTSrcElemKernel(const stk::mesh::BulkData &, const SolutionOptions &, ElemDataRequests &,
                                          const bool lumpedMass);
  virtual ~SpecificDissipationRateSSTSrcElemKernel() {}
  bool device_descriptor(nalu_b200_kernel_desc &) override;
  const char *name() const override { return "SpecificDissipationRateSSTSrcElemKernel"; }
  double v_shape_function_[AlgTraits::numScvIp_][AlgTraits::nodesPerElement_];
private:
  SpecificDissipationRateSSTSrcElemKernel() = delete;
  ScalarFieldType *tkeNp1_{nullptr}, *sdrNp1_{nullptr}, *densityNp1_{nullptr}, *tvisc_{nullptr}, *fOneBlend_{nullptr};
  VectorFieldType *velocityNp1_{nullptr}, *coordinates_{nullptr};
  const bool lumpedMass_, shiftedGradOp_;
  const double betaStar_, sigmaWTwo_, betaOne_, betaTwo_, gammaOne_, gammaTwo_, tkeProdLimitRatio_;
  const int *ipNodeMap_;
};

// include/kernel/TurbKineticEnergySSTDESSrcElemKernel.h
template <typename AlgTraits>
class TurbKineticEnergySSTDESSrcElemKernel : public Kernel {
public:
  TurbKineticEnergySSTDESSrcElemKernel(const stk::mesh::BulkData &, const SolutionOptions &, ElemDataRequests &, const bool lumpedMass);
  virtual ~TurbKineticEnergySSTDESSrcElemKernel() {}
  bool device_descriptor(nalu_b200_kernel_desc &) override;
  const char *name() const override { return "TurbKineticEnergySSTDESSrcElemKernel"; }
  double v_shape_function_[AlgTraits::numScvIp_][AlgTraits::nodesPerElement_];
private:
  TurbKineticEnergySSTDESSrcElemKernel() = delete;
  ScalarFieldType *tkeNp1_{nullptr}, *sdrNp1_{nullptr}, *densityNp1_{nullptr}, *tvisc_{nullptr}, *maxLengthScale_{nullptr},
      *fOneBlend_{nullptr};
  VectorFieldType *velocityNp1_{nullptr}, *coordinates_{nullptr};
  const bool lumpedMass_, shiftedGradOp_;
  const double betaStar_, tkeProdLimitRatio_, cDESke_, cDESkw_;
  const int *ipNodeMap_;
};

// include/kernel/TurbKineticEnergyKsgsSrcElemKernel.h
template <typename AlgTraits>
class TurbKineticEnergyKsgsSrcElemKernel : public Kernel {
public:
  TurbKineticEnergyKsgsSrcElemKernel(const stk::mesh::BulkData &, const SolutionOptions &, ElemDataRequests &, double lrksgsfac);
  virtual ~TurbKineticEnergyKsgsSrcElemKernel() {}
  bool device_descriptor(nalu_b200_kernel_desc &) override;
  const char *name() const override { return "TurbKineticEnergyKsgsSrcElemKernel"; }
private:
  TurbKineticEnergyKsgsSrcElemKernel() = delete;
  ScalarFieldType *tkeNp1_{nullptr}, *densityNp1_{nullptr}, *tvisc_{nullptr}, *dualNodalVolume_{nullptr}, *cEps_{nullptr},
      *visc_{nullptr}, *dsqrtkSq_{nullptr};
  VectorFieldType *coordinates_{nullptr};
  GenericFieldType *Gju_{nullptr};
  const double lrksgsfac_, tkeProdLimitRatio_;
  const int *ipNodeMap_;
};

} // namespace nalu
} // namespace sierra
