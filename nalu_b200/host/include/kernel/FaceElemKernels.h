// FaceElemKernels.h -- the face/element boundary-condition kernels of the reference for Quad4 faces of Hex8 elements:
//   MomentumSymmetryElemKernel     include/kernel/MomentumSymmetryElemKernel.h:28-67,     src/kernel/MomentumSymmetryElemKernel.C:23-63
//   MomentumOpenAdvDiffElemKernel  include/kernel/MomentumOpenAdvDiffElemKernel.h:32-99,  src/kernel/MomentumOpenAdvDiffElemKernel.C:26-100
//   ScalarOpenAdvElemKernel        include/kernel/ScalarOpenAdvElemKernel.h:32-92,        src/kernel/ScalarOpenAdvElemKernel.C:25-82
//   ContinuityOpenElemKernel       include/kernel/ContinuityOpenElemKernel.h:28-80,       src/kernel/ContinuityOpenElemKernel.C:25-97
// Same constructor signatures (face and element ElemDataRequests) and registrations as the reference.  Instead of
// execute(lhs, rhs, faceScratchViews, elemScratchViews, elemFaceOrdinal) each kernel exposes device_face_descriptor():
// AssembleFaceElemSolverAlgorithm::execute lowers the registered list to one nalu_b200_assemble_face_elem call.
#pragma once
#include "Kernel.h"
#include "UpwAdvDiffElemKernels.h"
namespace sierra {
namespace nalu {

// get_face_shape_fn_data (include/kernel/Kernel.h:58-72)
template <typename BcAlgTraits, typename LambdaFunction>
void get_face_shape_fn_data(LambdaFunction lambdaFunction, double (&view)[BcAlgTraits::numFaceIp_][BcAlgTraits::nodesPerFace_])
{
  lambdaFunction(&view[0][0]);
}

/** A kernel registered on an AssembleFaceElemSolverAlgorithm.  first_face: index of the part's first face in the rank's
 *  face numbering -- the face fields handed to the device start there. */
class FaceElemKernel : public Kernel {
public:
  virtual bool device_face_descriptor(nalu_b200_face_kernel_desc &, size_t first_face) = 0;
protected:
  static void clear(nalu_b200_face_kernel_desc &d);
  static void bind(nalu_b200_face_kernel_desc &d, int slot, stk::mesh::FieldBase *f);
  static void bind_face(nalu_b200_face_kernel_desc &d, int slot, stk::mesh::FieldBase *f, size_t first_face);
};

template <typename BcAlgTraits>
class MomentumSymmetryElemKernel : public FaceElemKernel {
public:
  MomentumSymmetryElemKernel(const stk::mesh::MetaData &metaData, const SolutionOptions &solnOpts, VectorFieldType *velocity,
                             ScalarFieldType *viscosity, ElemDataRequests &faceDataPreReqs, ElemDataRequests &elemDataPreReqs);
  virtual ~MomentumSymmetryElemKernel() {}
  bool device_face_descriptor(nalu_b200_face_kernel_desc &, size_t first_face) override;
  const char *name() const override { return "MomentumSymmetryElemKernel"; }
  double vf_shape_function_[BcAlgTraits::numFaceIp_][BcAlgTraits::nodesPerFace_];
private:
  MomentumSymmetryElemKernel() = delete;
  ScalarFieldType *viscosity_{nullptr};
  VectorFieldType *velocityNp1_{nullptr};
  VectorFieldType *coordinates_{nullptr};
  GenericFieldType *exposedAreaVec_{nullptr};
  const double includeDivU_;
  const bool shiftedGradOp_;
  MasterElement *meSCS_{nullptr};
};

template <typename BcAlgTraits>
class MomentumOpenAdvDiffElemKernel : public FaceElemKernel {
public:
  MomentumOpenAdvDiffElemKernel(const stk::mesh::MetaData &metaData, const SolutionOptions &solnOpts, EquationSystem *eqSystem,
                                VectorFieldType *velocity, GenericFieldType *Gjui, ScalarFieldType *viscosity,
                                ElemDataRequests &faceDataPreReqs, ElemDataRequests &elemDataPreReqs);
  virtual ~MomentumOpenAdvDiffElemKernel() {}
  bool device_face_descriptor(nalu_b200_face_kernel_desc &, size_t first_face) override;
  const char *name() const override { return "MomentumOpenAdvDiffElemKernel"; }
  double vf_shape_function_[BcAlgTraits::numFaceIp_][BcAlgTraits::nodesPerFace_];
  double vf_adv_shape_function_[BcAlgTraits::numFaceIp_][BcAlgTraits::nodesPerFace_];
private:
  MomentumOpenAdvDiffElemKernel() = delete;
  ScalarFieldType *viscosity_{nullptr};
  GenericFieldType *Gjui_{nullptr};
  VectorFieldType *velocityNp1_{nullptr}, *velocityRTM_{nullptr}, *meshVelocity_{nullptr}, *coordinates_{nullptr}, *velocityBc_{nullptr};
  ScalarFieldType *density_{nullptr};
  GenericFieldType *exposedAreaVec_{nullptr}, *openMassFlowRate_{nullptr};
  const double alphaUpw_, om_alphaUpw_, hoUpwind_, includeDivU_, meshVelocityCorrection_;
  const bool shiftedGradOp_;
  bool skewSymmetric_;
  const double small_{1.0e-16};
  const int *faceIpNodeMap_{nullptr};
  MasterElement *meSCS_{nullptr};
  PecletFunctionSpec pecletFunction_;
};

template <typename BcAlgTraits>
class ScalarOpenAdvElemKernel : public FaceElemKernel {
public:
  ScalarOpenAdvElemKernel(const stk::mesh::MetaData &metaData, const SolutionOptions &solnOpts, EquationSystem *eqSystem,
                          ScalarFieldType *scalarQ, ScalarFieldType *bcScalarQ, VectorFieldType *Gjq, ScalarFieldType *diffFluxCoeff,
                          ElemDataRequests &faceDataPreReqs, ElemDataRequests &elemDataPreReqs);
  virtual ~ScalarOpenAdvElemKernel() {}
  bool device_face_descriptor(nalu_b200_face_kernel_desc &, size_t first_face) override;
  const char *name() const override { return "ScalarOpenAdvElemKernel"; }
  double vf_adv_shape_function_[BcAlgTraits::numFaceIp_][BcAlgTraits::nodesPerFace_];
private:
  ScalarOpenAdvElemKernel() = delete;
  ScalarFieldType *scalarQ_{nullptr}, *bcScalarQ_{nullptr};
  VectorFieldType *Gjq_{nullptr};
  ScalarFieldType *diffFluxCoeff_{nullptr};
  VectorFieldType *velocityRTM_{nullptr}, *coordinates_{nullptr};
  ScalarFieldType *density_{nullptr};
  GenericFieldType *openMassFlowRate_{nullptr};
  const double alphaUpw_, om_alphaUpw_, hoUpwind_;
  bool skewSymmetric_;
  const double small_{1.0e-16};
  const int *faceIpNodeMap_{nullptr};
  MasterElement *meSCS_{nullptr};
  PecletFunctionSpec pecletFunction_;
};

template <typename BcAlgTraits>
class ContinuityOpenElemKernel : public FaceElemKernel {
public:
  ContinuityOpenElemKernel(const stk::mesh::MetaData &metaData, const SolutionOptions &solnOpts, ElemDataRequests &faceDataPreReqs,
                           ElemDataRequests &elemDataPreReqs);
  virtual ~ContinuityOpenElemKernel() {}
  /** projTimeScale = dt / gamma1 (src/kernel/ContinuityOpenElemKernel.C:89-97) */
  void setup(const TimeIntegrator &) override;
  bool device_face_descriptor(nalu_b200_face_kernel_desc &, size_t first_face) override;
  const char *name() const override { return "ContinuityOpenElemKernel"; }
  double vf_shape_function_[BcAlgTraits::numFaceIp_][BcAlgTraits::nodesPerFace_];
private:
  ContinuityOpenElemKernel() = delete;
  VectorFieldType *velocityRTM_{nullptr}, *Gpdx_{nullptr}, *coordinates_{nullptr};
  ScalarFieldType *pressure_{nullptr}, *pressureBc_{nullptr}, *density_{nullptr};
  GenericFieldType *exposedAreaVec_{nullptr}, *dynamicPressure_{nullptr};
  const bool shiftedGradOp_, reducedSensitivities_, shiftMdot_;
  double dt_{1.0}, gamma1_{1.0};
  const double penaltyFac_{2.0};
  MasterElement *meSCS_{nullptr};
};

// include/kernel/MomentumWallFunctionElemKernel.h:26-70.  The reference registers it on an AssembleElemSolverAlgorithm over the
// wall part's faces (BcAlgTraits = AlgTraitsQuad4, one ElemDataRequests); here it is a kernel of the face/element algorithm of
// that part (the face set already knows the element attached to each face, whose slot map carries the scatter), so BcAlgTraits =
// AlgTraitsQuad4Hex8 and the face requirements go to faceDataNeeded_.
template <typename BcAlgTraits>
class MomentumWallFunctionElemKernel : public FaceElemKernel {
public:
  MomentumWallFunctionElemKernel(const stk::mesh::BulkData &bulkData, const SolutionOptions &solnOpts, ElemDataRequests &dataPreReqs);
  virtual ~MomentumWallFunctionElemKernel() {}
  bool device_face_descriptor(nalu_b200_face_kernel_desc &, size_t first_face) override;
  const char *name() const override { return "MomentumWallFunctionElemKernel"; }
  double vf_shape_function_[BcAlgTraits::numFaceIp_][BcAlgTraits::nodesPerFace_];
private:
  MomentumWallFunctionElemKernel() = delete;
  const double elog_, kappa_, yplusCrit_;
  const int *ipNodeMap_;
  VectorFieldType *velocityNp1_{nullptr}, *bcVelocity_{nullptr}, *coordinates_{nullptr};
  ScalarFieldType *density_{nullptr}, *viscosity_{nullptr};
  GenericFieldType *exposedAreaVec_{nullptr}, *wallFrictionVelocityBip_{nullptr}, *wallNormalDistanceBip_{nullptr};
};

} // namespace nalu
} // namespace sierra
