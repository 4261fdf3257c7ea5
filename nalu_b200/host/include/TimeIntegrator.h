// TimeIntegrator.h -- what Kernel::setup reads (src/TimeIntegrator.C:411-419: BE gammas 1,-1,0;
// BDF2 (1+2tau)/(1+tau), -(1+tau), tau^2/(1+tau))
#pragma once
namespace sierra {
namespace nalu {
class TimeIntegrator {
public:
  double timeStepN_ = 1.0, timeStepNm1_ = 1.0;
  double gamma1_ = 1.0, gamma2_ = -1.0, gamma3_ = 0.0;
  double get_time_step() const { return timeStepN_; }
  double get_gamma1() const { return gamma1_; }
  double get_gamma2() const { return gamma2_; }
  double get_gamma3() const { return gamma3_; }
  void compute_gamma_bdf2()
  {
    const double tau = timeStepN_ / timeStepNm1_;
    gamma1_ = (1.0 + 2.0 * tau) / (1.0 + tau);
    gamma2_ = -(1.0 + tau);
    gamma3_ = tau * tau / (1.0 + tau);
  }
};
} // namespace nalu
} // namespace sierra
