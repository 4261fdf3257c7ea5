// SolutionOptions.h -- the option fields the kernels on this path read (defaults from
// src/SolutionOptions.C:40-95, accessor semantics from :824-845)
#pragma once
#include <map>
#include <string>
namespace sierra {
namespace nalu {
class SolutionOptions {
public:
  double includeDivU_ = 0.0;                // src/SolutionOptions.C:49
  bool cvfemShiftMdot_ = false;             // :40
  bool cvfemReducedSensPoisson_ = false;    // :41
  bool balancedForce_ = false;
  bool meshMotion_ = false, meshDeformation_ = false, externalMeshDeformation_ = false;
  std::map<std::string, bool> shiftedGradOpMap_, skewSymmetricMap_;
  bool does_mesh_move() const { return meshMotion_ || meshDeformation_ || externalMeshDeformation_; }
  std::string get_coordinates_name() const { return does_mesh_move() ? "current_coordinates" : "coordinates"; }
  bool get_shifted_grad_op(const std::string &dofName) const
  {
    auto it = shiftedGradOpMap_.find(dofName);
    return it == shiftedGradOpMap_.end() ? false : it->second;
  }
  bool get_skew_symmetric(const std::string &dofName) const
  {
    auto it = skewSymmetricMap_.find(dofName);
    return it == skewSymmetricMap_.end() ? false : it->second;
  }
};
} // namespace nalu
} // namespace sierra
