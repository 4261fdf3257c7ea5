// Element mass kernels: include/kernel/{Momentum,Continuity,Scalar}MassElemKernel.h;
// src/kernel/MomentumMassElemKernel.C:27-96, ContinuityMassElemKernel.C:27-82, ScalarMassElemKernel.C:27-95
#pragma once
#include "Kernel.h"
namespace sierra {
namespace nalu {

template <typename AlgTraits>
class MomentumMassElemKernel : public Kernel {
public:
  MomentumMassElemKernel(const stk::mesh::BulkData &, const SolutionOptions &, ElemDataRequests &, const bool lumpedMass);
  virtual ~MomentumMassElemKernel() {}
  void setup(const TimeIntegrator &) override;
  bool device_descriptor(nalu_b200_kernel_desc &) override;
  const char *name() const override { return "MomentumMassElemKernel"; }
  double v_shape_function_[AlgTraits::numScvIp_][AlgTraits::nodesPerElement_];
private:
  MomentumMassElemKernel() = delete;
  VectorFieldType *velocityNm1_{nullptr}, *velocityN_{nullptr}, *velocityNp1_{nullptr};
  ScalarFieldType *densityNm1_{nullptr}, *densityN_{nullptr}, *densityNp1_{nullptr};
  VectorFieldType *Gjp_{nullptr}, *coordinates_{nullptr};
  double dt_{0.0}, gamma1_{0.0}, gamma2_{0.0}, gamma3_{0.0};
  const bool lumpedMass_;
  const double densFac_, om_densFac_;
  const int *ipNodeMap_;
};

template <typename AlgTraits>
class ContinuityMassElemKernel : public Kernel {
public:
  ContinuityMassElemKernel(const stk::mesh::BulkData &, const SolutionOptions &, ElemDataRequests &, const bool lumpedMass);
  virtual ~ContinuityMassElemKernel() {}
  void setup(const TimeIntegrator &) override;
  bool device_descriptor(nalu_b200_kernel_desc &) override;
  const char *name() const override { return "ContinuityMassElemKernel"; }
  double v_shape_function_[AlgTraits::numScvIp_][AlgTraits::nodesPerElement_];
private:
  ContinuityMassElemKernel() = delete;
  ScalarFieldType *densityNm1_{nullptr}, *densityN_{nullptr}, *densityNp1_{nullptr};
  VectorFieldType *coordinates_{nullptr};
  double dt_{0.0}, gamma1_{0.0}, gamma2_{0.0}, gamma3_{0.0};
  const bool lumpedMass_;
  const int *ipNodeMap_;
};

template <typename AlgTraits>
class ScalarMassElemKernel : public Kernel {
public:
  ScalarMassElemKernel(const stk::mesh::BulkData &, const SolutionOptions &, ScalarFieldType *, ElemDataRequests &,
                       const bool lumpedMass);
  virtual ~ScalarMassElemKernel() {}
  void setup(const TimeIntegrator &) override;
  bool device_descriptor(nalu_b200_kernel_desc &) override;
  const char *name() const override { return "ScalarMassElemKernel"; }
  double v_shape_function_[AlgTraits::numScvIp_][AlgTraits::nodesPerElement_];
private:
  ScalarMassElemKernel() = delete;
  ScalarFieldType *scalarQNm1_{nullptr}, *scalarQN_{nullptr}, *scalarQNp1_{nullptr};
  ScalarFieldType *densityNm1_{nullptr}, *densityN_{nullptr}, *densityNp1_{nullptr};
  VectorFieldType *coordinates_{nullptr};
  double dt_{0.0}, gamma1_{0.0}, gamma2_{0.0}, gamma3_{0.0};
  const bool lumpedMass_;
  const int *ipNodeMap_;
};

template <>
struct has_hex27_instantiation<MomentumMassElemKernel> : std::true_type {};

} // namespace nalu
} // namespace sierra
