// Realm.h / EquationSystem.h -- only what the assembly path dereferences: realm_.bulk_data(),
// realm_.timeIntegrator_, realm_.solutionOptions_, eqSystem_->linsys_, eqSystem_->name_
// (src/AssembleElemSolverAlgorithm.C:66-82, src/SolverAlgorithm.C:57-68)
#pragma once
#include "SolutionOptions.h"
#include "TimeIntegrator.h"
#include "stk_shim.h"
#include <memory>
#include <string>

namespace sierra {
namespace nalu {

class LinearSystem;

class Realm {
public:
  Realm() : solutionOptions_(new SolutionOptions()), timeIntegrator_(new TimeIntegrator()) {}
  ~Realm()
  {
    delete solutionOptions_;
    delete timeIntegrator_;
  }
  stk::mesh::BulkData &bulk_data() { return *bulkData_; }
  stk::mesh::MetaData &meta_data() { return bulkData_->mesh_meta_data(); }
  std::shared_ptr<stk::mesh::BulkData> bulkData_;
  SolutionOptions *solutionOptions_;
  TimeIntegrator *timeIntegrator_;
};

class EquationSystem {
public:
  EquationSystem(Realm &realm, const std::string name = "no_name", const std::string eqnTypeName = "")
      : realm_(realm), name_(name), eqnTypeName_(eqnTypeName), linsys_(nullptr)
  {
  }
  virtual ~EquationSystem();
  Realm &realm_;
  std::string name_;
  const std::string eqnTypeName_;
  LinearSystem *linsys_; // owned, like the reference's equation system owns its linear system
};

} // namespace nalu
} // namespace sierra
