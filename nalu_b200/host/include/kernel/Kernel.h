// Kernel.h -- the reference's computational-kernel plugin interface (include/kernel/Kernel.h:95-125).
//
// A Kernel subclass is constructed with the reference's own argument list, registers its needs in
// the algorithm's ElemDataRequests exactly as the reference does, and gets setup(TimeIntegrator) before
// every assembly.  What changes on this path: the per-element arithmetic is not a host virtual call.
// AssembleElemSolverAlgorithm::execute() asks every active kernel for a POD descriptor
// (device_descriptor) and evaluates the whole registered list in ONE fused CUDA launch through the C
// ABI (nalu_b200_assemble_elem).  There is no CPU execute(): the product path has no CPU fallback.
#pragma once
#include <type_traits>
#include "../../../../include/nalu_b200.h"
#include "../AlgTraits.h"
#include "../ElemDataRequests.h"
#include "../FieldTypeDef.h"
#include "../SolutionOptions.h"
#include "../TimeIntegrator.h"
#include "../master_element/MasterElement.h"

namespace sierra {
namespace nalu {

// get_scs_shape_fn_data / get_scv_shape_fn_data (include/kernel/Kernel.h:26-56): fill a
// [numIp][nodesPerElement] table from a MasterElement callback
template <typename AlgTraits, typename LambdaFunction>
void get_scs_shape_fn_data(LambdaFunction lambdaFunction, double (&view)[AlgTraits::numScsIp_][AlgTraits::nodesPerElement_])
{
  lambdaFunction(&view[0][0]);
}
template <typename AlgTraits, typename LambdaFunction>
void get_scv_shape_fn_data(LambdaFunction lambdaFunction, double (&view)[AlgTraits::numScvIp_][AlgTraits::nodesPerElement_])
{
  lambdaFunction(&view[0][0]);
}

// kernels that are also instantiated for AlgTraitsHex27 (build_topo_kernel dispatches HEX_27 parts to them)
template <template <typename> class T>
struct has_hex27_instantiation : std::false_type {};

class Kernel {
public:
  Kernel() = default;
  virtual ~Kernel() {}
  /** Perform pre-timestep work for the computational kernel */
  virtual void setup(const TimeIntegrator &) {}
  /** What the fused device launch needs to evaluate this kernel: kind, option bits, scalars captured
   *  by the constructor / setup(), device addresses of the gathered fields.  Returns false for a
   *  kernel that has no device form (the algorithm then refuses to run rather than fall back). */
  virtual bool device_descriptor(nalu_b200_kernel_desc &) { return false; }
  virtual const char *name() const { return "Kernel"; }

protected:
  static void clear(nalu_b200_kernel_desc &d);
  static void bind(nalu_b200_kernel_desc &d, int slot, stk::mesh::FieldBase *f);
};

#define NALU_B200_DECLARE_KERNEL(NAME) template class NAME<AlgTraitsHex8>;

} // namespace nalu
} // namespace sierra
