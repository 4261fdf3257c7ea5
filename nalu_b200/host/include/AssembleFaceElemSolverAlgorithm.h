// AssembleFaceElemSolverAlgorithm -- the reference class of the same name (include/AssembleFaceElemSolverAlgorithm.h:30-146,
// src/AssembleFaceElemSolverAlgorithm.C:48-108): boundary contributions that need the face AND its attached element.
// Same constructor, public members (faceDataNeeded_, elemDataNeeded_, numDof_, nodesPerFace_, nodesPerElem_, rhsSize_) and
// call sequence; execute() lowers the active kernels to ONE nalu_b200_assemble_face_elem call over the part's faces.
// The reference's run_face_elem_algorithm(bulk, lambda) hands a host lambda SharedMemData_FaceElem views of one SIMD group
// of faces; here run_face_elem_algorithm calls the lambda once per face with the (face nodes, element nodes, face ordinal)
// connectivity -- what unit_tests/kernels/UnitTestFaceElemBasic.C:83-125 checks -- the arithmetic being on the device.
#pragma once
#include "AssembleElemSolverAlgorithm.h"
#include "kernel/FaceElemKernels.h"

namespace sierra {
namespace nalu {

struct FaceElemConnectivity {
  const int32_t *elemNodes; // the attached element's nodes (local indices), nodesPerElem
  int32_t faceNodes[4];     // the face's nodes, in the order of the element's side_node_ordinals
  int elemFaceOrdinal;
  int64_t elem, face;       // local element; face index within the rank's face numbering
};

class AssembleFaceElemSolverAlgorithm : public SolverAlgorithm {
public:
  AssembleFaceElemSolverAlgorithm(Realm &realm, stk::mesh::Part *part, EquationSystem *eqSystem, unsigned nodesPerFace,
                                  unsigned nodesPerElem, bool interleaveMeViews = true);
  virtual ~AssembleFaceElemSolverAlgorithm();
  virtual void initialize_connectivity(); // eqSystem_->linsys_->buildFaceElemToNodeGraph(partVec_)
  virtual void execute();

  template <typename LambdaFunction>
  void run_face_elem_algorithm(stk::mesh::BulkData &bulk, LambdaFunction lambdaFunc)
  {
    for (stk::mesh::Part *part : partVec_) {
      const auto &fe = part->face_elems();
      const auto &fo = part->face_ordinals();
      for (size_t f = 0; f < fe.size(); ++f) {
        FaceElemConnectivity c;
        c.elem = fe[f];
        c.face = int64_t(part->first_face() + f);
        c.elemFaceOrdinal = fo[f];
        c.elemNodes = bulk.begin_nodes(stk::mesh::Entity{stk::topology::ELEM_RANK, fe[f]});
        int ord[4];
        stk::topology(stk::topology::HEX_8).face_node_ordinals(unsigned(fo[f]), ord);
        for (int i = 0; i < 4; ++i) c.faceNodes[i] = c.elemNodes[ord[i]];
        lambdaFunc(c);
      }
    }
  }

  ElemDataRequests faceDataNeeded_;
  ElemDataRequests elemDataNeeded_;
  unsigned numDof_;
  unsigned nodesPerFace_;
  unsigned nodesPerElem_;
  int rhsSize_;
  const bool interleaveMEViews_;

private:
  std::vector<nalu_b200_face_set *> faceSets_; // one per part, created at the first execute()
};

} // namespace nalu
} // namespace sierra
