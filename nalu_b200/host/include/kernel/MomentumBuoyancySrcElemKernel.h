// MomentumBuoyancySrcElemKernel / MomentumBuoyancyBoussinesqSrcElemKernel -- host side of the reference's
// include/kernel/MomentumBuoyancySrcElemKernel.h:29-60 and MomentumBuoyancyBoussinesqSrcElemKernel.h: same
// constructor signatures and ElemDataRequests registrations; execute() is lowered to a device descriptor.
#pragma once
#include "Kernel.h"
namespace sierra {
namespace nalu {

template <typename AlgTraits>
class MomentumBuoyancySrcElemKernel : public Kernel {
public:
  MomentumBuoyancySrcElemKernel(const stk::mesh::BulkData &, const SolutionOptions &, ElemDataRequests &);
  virtual ~MomentumBuoyancySrcElemKernel() {}
  bool device_descriptor(nalu_b200_kernel_desc &) override;
  const char *name() const override { return "MomentumBuoyancySrcElemKernel"; }
  double v_shape_function_[AlgTraits::numScvIp_][AlgTraits::nodesPerElement_];
private:
  MomentumBuoyancySrcElemKernel() = delete;
  ScalarFieldType *densityNp1_{nullptr};
  VectorFieldType *coordinates_{nullptr};
  double rhoRef_;
  double gravity_[3];
  const int *ipNodeMap_;
};

template <typename AlgTraits>
class MomentumBuoyancyBoussinesqSrcElemKernel : public Kernel {
public:
  MomentumBuoyancyBoussinesqSrcElemKernel(const stk::mesh::BulkData &, const SolutionOptions &, ElemDataRequests &);
  virtual ~MomentumBuoyancyBoussinesqSrcElemKernel() {}
  bool device_descriptor(nalu_b200_kernel_desc &) override;
  const char *name() const override { return "MomentumBuoyancyBoussinesqSrcElemKernel"; }
  double v_shape_function_[AlgTraits::numScvIp_][AlgTraits::nodesPerElement_];
private:
  MomentumBuoyancyBoussinesqSrcElemKernel() = delete;
  ScalarFieldType *temperatureNp1_{nullptr};
  VectorFieldType *coordinates_{nullptr};
  double rhoRef_, tRef_, beta_;
  double gravity_[3];
  const int *ipNodeMap_;
};

} // namespace nalu
} // namespace sierra
