// SolutionOptions.h -- the option fields the kernels on this path read (defaults from
// src/SolutionOptions.C:40-95, accessor semantics from :824-845)
#pragma once
#include <array>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>
namespace sierra {
namespace nalu {
// include/Enums.h:156-208 (the members this path reads)
enum TurbulenceModel { LAMINAR = 0, KSGS = 1, SMAGORINSKY = 2, WALE = 3, SST = 4, SST_DES = 5, KEPS = 6, LRKSGS = 7 };
enum TurbulenceModelConstant {
  TM_cMu = 0, TM_kappa = 1, TM_cDESke = 2, TM_cDESkw = 3, TM_tkeProdLimitRatio = 4, TM_cmuEps = 5, TM_cEps = 6,
  TM_betaStar = 7, TM_aOne = 8, TM_betaOne = 9, TM_betaTwo = 10, TM_gammaOne = 11, TM_gammaTwo = 12,
  TM_sigmaKOne = 13, TM_sigmaKTwo = 14, TM_sigmaWOne = 15, TM_sigmaWTwo = 16, TM_elog = 23, TM_yplus_crit = 24, TM_END = 27
};
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
  // src/SolutionOptions.C:742-770 (values of the constants the SST source kernels read), :824-832
  TurbulenceModel turbulenceModel_ = LAMINAR;
  std::map<TurbulenceModelConstant, double> turbModelConstantMap_;
  void initialize_turbulence_constants()
  {
    turbModelConstantMap_[TM_cMu] = 0.09;
    turbModelConstantMap_[TM_kappa] = 0.41;
    turbModelConstantMap_[TM_cDESke] = 0.61;
    turbModelConstantMap_[TM_cDESkw] = 0.78;
    turbModelConstantMap_[TM_tkeProdLimitRatio] = (turbulenceModel_ == SST || turbulenceModel_ == SST_DES) ? 10.0 : 500.0;
    turbModelConstantMap_[TM_betaStar] = 0.09;
    turbModelConstantMap_[TM_aOne] = 0.31;
    turbModelConstantMap_[TM_betaOne] = 0.075;
    turbModelConstantMap_[TM_betaTwo] = 0.0828;
    turbModelConstantMap_[TM_gammaOne] = 5.0 / 9.0;
    turbModelConstantMap_[TM_gammaTwo] = 0.44;
    turbModelConstantMap_[TM_sigmaKOne] = 0.85;
    turbModelConstantMap_[TM_sigmaKTwo] = 1.0;
    turbModelConstantMap_[TM_sigmaWOne] = 0.50;
    turbModelConstantMap_[TM_sigmaWTwo] = 0.856;
    turbModelConstantMap_[TM_elog] = 9.8;         // src/SolutionOptions.C:769-770
    turbModelConstantMap_[TM_yplus_crit] = 11.63;
  }
  double get_turb_model_constant(TurbulenceModelConstant turbModelEnum) const
  {
    auto it = turbModelConstantMap_.find(turbModelEnum);
    if (it == turbModelConstantMap_.end()) throw std::runtime_error("unknown (not initialized) turbulence model constant");
    return it->second;
  }
  std::array<double, 3> get_gravity_vector() const
  {
    std::array<double, 3> g{{0.0, 0.0, 0.0}};
    for (size_t i = 0; i < gravity_.size() && i < 3; ++i) g[i] = gravity_[i];
    return g;
  }
  // upwind / Peclet blending (defaults src/SolutionOptions.C:32-44; accessors :776-821; the hybrid / tanh
  // accessors live on Realm in the reference, src/Realm.C:4231-4458, and read these maps)
  double hybridDefault_ = 0.0, alphaDefault_ = 0.0, alphaUpwDefault_ = 1.0, upwDefault_ = 1.0;
  std::string tanhFormDefault_ = "classic";
  double tanhTransDefault_ = 2.0, tanhWidthDefault_ = 4.0;
  std::map<std::string, double> hybridMap_, alphaMap_, alphaUpwMap_, upwMap_, tanhTransMap_, tanhWidthMap_;
  std::map<std::string, bool> limiterMap_;
  std::map<std::string, std::string> tanhFormMap_;
  static double lookup(const std::map<std::string, double> &m, const std::string &dofName, double dflt)
  {
    auto it = m.find(dofName);
    return it == m.end() ? dflt : it->second;
  }
  double get_alpha_factor(const std::string &dofName) const { return lookup(alphaMap_, dofName, alphaDefault_); }
  double get_alpha_upw_factor(const std::string &dofName) const { return lookup(alphaUpwMap_, dofName, alphaUpwDefault_); }
  double get_upw_factor(const std::string &dofName) const { return lookup(upwMap_, dofName, upwDefault_); }
  bool primitive_uses_limiter(const std::string &dofName) const
  {
    auto it = limiterMap_.find(dofName);
    return it == limiterMap_.end() ? false : it->second;
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
