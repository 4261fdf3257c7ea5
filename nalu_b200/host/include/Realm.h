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
  // src/Realm.C:4231-4241,4416-4458
  double get_hybrid_factor(const std::string dofName) { return SolutionOptions::lookup(solutionOptions_->hybridMap_, dofName, solutionOptions_->hybridDefault_); }
  std::string get_tanh_functional_form(const std::string dofName)
  {
    auto it = solutionOptions_->tanhFormMap_.find(dofName);
    return it == solutionOptions_->tanhFormMap_.end() ? solutionOptions_->tanhFormDefault_ : it->second;
  }
  double get_tanh_trans(const std::string dofName) { return SolutionOptions::lookup(solutionOptions_->tanhTransMap_, dofName, solutionOptions_->tanhTransDefault_); }
  double get_tanh_width(const std::string dofName) { return SolutionOptions::lookup(solutionOptions_->tanhWidthMap_, dofName, solutionOptions_->tanhWidthDefault_); }
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
