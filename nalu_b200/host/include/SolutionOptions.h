// SolutionOptions.h -- the option fields the kernels on this path read (defaults from
// src/SolutionOptions.C:40-95, accessor semantics from :824-845)
#pragma once
#include <array>
#include <map>
#include <string>
#include <vector>
namespace sierra {
namespace nalu {
class SolutionOptions {
public:
  double includeDivU_ = 0.0;                // src/SolutionOptions.C:49
  bool cvfemShiftMdot_ = false;             // :40
  bool cvfemReducedSensPoisson_ = false;    // :41
  bool balancedForce_ = false;
  // buoyancy sources (src/SolutionOptions.C: gravity_, referenceDensity_, referenceTemperature_,
  // thermalExpansionCoeff_; get_gravity_vector pads to three components)
  std::vector<double> gravity_ = std::vector<double>(3, 0.0);
  double referenceDensity_ = 0.0, referenceTemperature_ = 298.0, thermalExpansionCoeff_ = 1.0;
  std::array<double, 3> get_gravity_vector() const
  {
    std::array<double, 3> g{{0.0, 0.0, 0.0}};
    for (size_t i = 0; i < gravity_.size() && i < 3; ++i) g[i] = gravity_[i];
    return g;
  }
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
