// include/kernel/ScalarUpwAdvDiffElemKernel.h:27-96, src/kernel/ScalarUpwAdvDiffElemKernel.C:28-108;
// include/kernel/MomentumUpwAdvDiffElemKernel.h:27-99, src/kernel/MomentumUpwAdvDiffElemKernel.C:28-106.
// Same constructor signatures and ElemDataRequests registrations as the reference; the Peclet function the
// reference obtains from EquationSystem::create_peclet_function (include/EquationSystem.h:306-325) is captured
// as its kind and parameters, since it is evaluated on the device.
#pragma once
#include "Kernel.h"
#include "../Realm.h"
namespace sierra {
namespace nalu {
struct PecletFunctionSpec {
  bool tanh_form = false;          // realm.get_tanh_functional_form(dofName) != "classic"
  double hybridFactor = 0.0;       // ClassicPecletFunction(A = 5, hybridFactor)
  double c1 = 2.0, c2 = 4.0;       // TanhFunction(c1 = tanh_trans, c2 = tanh_width)
};

template <typename AlgTraits>
class ScalarUpwAdvDiffElemKernel : public Kernel {
public:
  ScalarUpwAdvDiffElemKernel(const stk::mesh::BulkData &, const SolutionOptions &, EquationSystem *, ScalarFieldType *,
                             VectorFieldType *, ScalarFieldType *, ElemDataRequests &);
  virtual ~ScalarUpwAdvDiffElemKernel() {}
  void setup(const TimeIntegrator &) override;
  bool device_descriptor(nalu_b200_kernel_desc &) override;
  const char *name() const override { return "ScalarUpwAdvDiffElemKernel"; }
  double v_shape_function_[AlgTraits::numScsIp_][AlgTraits::nodesPerElement_];
  double v_adv_shape_function_[AlgTraits::numScsIp_][AlgTraits::nodesPerElement_];
private:
  ScalarUpwAdvDiffElemKernel() = delete;
  const SolutionOptions &solnOpts_;
  ScalarFieldType *scalarQ_{nullptr};
  VectorFieldType *Gjq_{nullptr};
  ScalarFieldType *diffFluxCoeff_{nullptr};
  VectorFieldType *velocityRTM_{nullptr};
  VectorFieldType *coordinates_{nullptr};
  ScalarFieldType *density_{nullptr};
  GenericFieldType *massFlowRate_{nullptr};
  const int *lrscv_;
  const std::string dofName_;
  double alpha_, alphaUpw_, hoUpwind_;
  bool useLimiter_;
  double om_alpha_, om_alphaUpw_;
  const bool shiftedGradOp_;
  bool skewSymmetric_;
  const double small_{1.0e-16};
  PecletFunctionSpec pecletFunction_;
};

template <typename AlgTraits>
class MomentumUpwAdvDiffElemKernel : public Kernel {
public:
  MomentumUpwAdvDiffElemKernel(const stk::mesh::BulkData &, const SolutionOptions &, EquationSystem *, VectorFieldType *,
                               ScalarFieldType *, GenericFieldType *, ElemDataRequests &);
  virtual ~MomentumUpwAdvDiffElemKernel() {}
  void setup(const TimeIntegrator &) override;
  bool device_descriptor(nalu_b200_kernel_desc &) override;
  const char *name() const override { return "MomentumUpwAdvDiffElemKernel"; }
  double v_shape_function_[AlgTraits::numScsIp_][AlgTraits::nodesPerElement_];
  double v_adv_shape_function_[AlgTraits::numScsIp_][AlgTraits::nodesPerElement_];
private:
  MomentumUpwAdvDiffElemKernel() = delete;
  const SolutionOptions &solnOpts_;
  VectorFieldType *velocityNp1_{nullptr};
  VectorFieldType *velocityRTM_{nullptr};
  VectorFieldType *coordinates_{nullptr};
  ScalarFieldType *density_{nullptr};
  ScalarFieldType *viscosity_{nullptr};
  GenericFieldType *Gju_{nullptr};
  GenericFieldType *massFlowRate_{nullptr};
  const int *lrscv_;
  const std::string dofName_;
  double alpha_, alphaUpw_, hoUpwind_;
  bool useLimiter_;
  double om_alpha_, om_alphaUpw_;
  const double includeDivU_;
  const bool shiftedGradOp_;
  bool skewSymmetric_;
  const double small_{1.0e-16};
  PecletFunctionSpec pecletFunction_;
};
} // namespace nalu
} // namespace sierra
