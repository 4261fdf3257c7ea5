// Algorithm / SolverAlgorithm / AssembleElemSolverAlgorithm / AssembleNodeSolverAlgorithm -- the
// reference classes of the same names (include/Algorithm.h:28-59, include/SolverAlgorithm.h,
// include/AssembleElemSolverAlgorithm.h:38-112, include/AssembleNodeSolverAlgorithm.h).
#pragma once
#include "ElemDataRequests.h"
#include "LinearSystem.h"
#include "Realm.h"
#include "kernel/Kernel.h"
#include <vector>

namespace sierra {
namespace nalu {

class Algorithm {
public:
  Algorithm(Realm &realm, stk::mesh::Part *part) : realm_(realm) { partVec_.push_back(part); }
  Algorithm(Realm &realm, stk::mesh::PartVector &partVec) : realm_(realm), partVec_(partVec) {}
  virtual ~Algorithm(); // deletes the active kernels, src/Algorithm.C:46-55
  virtual void execute() = 0;
  virtual void pre_work() {}
  Realm &realm_;
  stk::mesh::PartVector partVec_;
  std::vector<Kernel *> activeKernels_;
};

class SolverAlgorithm : public Algorithm {
public:
  SolverAlgorithm(Realm &realm, stk::mesh::Part *part, EquationSystem *eqSystem) : Algorithm(realm, part), eqSystem_(eqSystem) {}
  virtual ~SolverAlgorithm() {}
  virtual void execute() = 0;
  virtual void initialize_connectivity() = 0;
  EquationSystem *eqSystem_;
};

class AssembleElemSolverAlgorithm : public SolverAlgorithm {
public:
  AssembleElemSolverAlgorithm(Realm &realm, stk::mesh::Part *part, EquationSystem *eqSystem, stk::mesh::EntityRank entityRank,
                              unsigned nodesPerEntity, bool interleaveMeViews = true);
  virtual ~AssembleElemSolverAlgorithm() {}
  virtual void initialize_connectivity(); // eqSystem_->linsys_->buildElemToNodeGraph(partVec_)
  /** kernel->setup(timeIntegrator) for every active kernel, then the whole registered list over all
   *  locally-owned elements in one fused device launch (src/AssembleElemSolverAlgorithm.C:74-100) */
  virtual void execute();
  /** the tiled kernel may store instead of add when no earlier algorithm touched the system */
  bool firstAlgorithmAfterZero_ = false;

  ElemDataRequests dataNeededByKernels_;
  stk::mesh::EntityRank entityRank_;
  unsigned nodesPerEntity_;
  int rhsSize_;
  const bool interleaveMEViews_;
};

/** Lumped mass through AssembleNodeSolverAlgorithm + {Momentum,Continuity,Scalar}MassBDF2NodeSuppAlg
 *  (src/AssembleNodeSolverAlgorithm.C:57-111, src/MomentumMassBDF2NodeSuppAlg.C:80-103): one-node
 *  "elements", diagonal block + rhs.  `kind` is the mass kernel kind of the C ABI. */
class AssembleNodeSolverAlgorithm : public SolverAlgorithm {
public:
  AssembleNodeSolverAlgorithm(Realm &realm, stk::mesh::Part *part, EquationSystem *eqSystem);
  virtual void initialize_connectivity(); // buildNodeGraph
  virtual void execute();
  int massKind_ = -1;                   // NALU_K_{MOMENTUM,CONTINUITY,SCALAR}_MASS
  stk::mesh::FieldBase *scalarQ_ = nullptr; // scalar mass only
};

} // namespace nalu
} // namespace sierra
