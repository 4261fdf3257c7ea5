// KernelBuilder.h -- include/kernel/KernelBuilder.h:45-68,170-191,290-328: name-gated, topology
// dispatched kernel creation.  A kernel is built iff its name is listed for the equation under
// `element_source_terms` (src/EquationSystem.C:462-475); creation order = call order in the equation
// system's source, not YAML order.
#pragma once
#include "../AssembleElemSolverAlgorithm.h"
#include <algorithm>
#include <map>
#include <string>
#include <utility>
#include <vector>

namespace sierra {
namespace nalu {

/** element_source_terms of the input deck: eqnTypeName -> kernel names */
typedef std::map<std::string, std::vector<std::string>> ElemSrcTermsMap;

inline bool supp_alg_is_requested(const ElemSrcTermsMap &m, const std::string &eqnType, const std::string &name)
{
  auto it = m.find(eqnType);
  return it != m.end() && std::find(it->second.begin(), it->second.end(), name) != it->second.end();
}

template <template <typename> class T, typename... Args>
Kernel *build_topo_kernel(stk::topology topo, Args &&...args)
{
  switch (topo.t_) {
  case stk::topology::HEX_8:
    return new T<AlgTraitsHex8>(std::forward<Args>(args)...);
  case stk::topology::HEX_27:
    if constexpr (has_hex27_instantiation<T>::value) return new T<AlgTraitsHex27>(std::forward<Args>(args)...);
    else throw std::runtime_error("build_topo_kernel: this kernel is not instantiated for HEXAHEDRON_27 (the momentum list, ContinuityAdv and ScalarAdvDiff are)");
  default:
    throw std::runtime_error("build_topo_kernel: only Hex8 (and Hex27 for four kernels) is on this path (topology " + topo.name() + ")");
  }
}

template <template <typename> class T, typename... Args>
bool build_topo_kernel_if_requested(stk::topology topo, const ElemSrcTermsMap &srcTerms, const std::string &eqnTypeName,
                                    std::vector<Kernel *> &kernelVec, std::string name, Args &&...args)
{
  const bool isCreated = supp_alg_is_requested(srcTerms, eqnTypeName, name);
  if (isCreated) kernelVec.push_back(build_topo_kernel<T>(topo, std::forward<Args>(args)...));
  return isCreated;
}

/** one AssembleElemSolverAlgorithm per (equation system, topology): eqSys.name_ + "_AssembleElemSolverAlg_" + topo.name() */
inline std::pair<AssembleElemSolverAlgorithm *, bool>
build_or_add_part_to_solver_alg(EquationSystem &eqSys, stk::mesh::Part &part,
                                std::map<std::string, SolverAlgorithm *> &solverAlgs)
{
  const stk::topology topo = part.topology();
  const std::string algName = eqSys.name_ + "_AssembleElemSolverAlg_" + topo.name();
  auto itc = solverAlgs.find(algName);
  const bool createNewAlg = itc == solverAlgs.end();
  if (createNewAlg) {
    // Hex8 runs the templated (non-interleaved) master-element path in the reference: :299-311
    auto *theSolverAlg = new AssembleElemSolverAlgorithm(eqSys.realm_, &part, &eqSys, topo.rank(), topo.num_nodes(), false);
    solverAlgs.insert({algName, theSolverAlg});
    return {theSolverAlg, true};
  }
  auto *theSolverAlg = dynamic_cast<AssembleElemSolverAlgorithm *>(itc->second);
  theSolverAlg->partVec_.push_back(&part);
  return {theSolverAlg, false};
}

} // namespace nalu
} // namespace sierra
