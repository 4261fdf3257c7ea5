// CPU check of the device master-element functions (nalu_b200/csrc/hex8_dev.cuh compiled as
// plain C++) against the oracle.  Test scaffolding: built and run by tests/test_device_math_cpu.py.
#define __host__
#define __device__
#define __forceinline__ inline
#include "../../nalu_b200/csrc/hex8_dev.cuh"
#include "../../oracle/nalu_oracle.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>

using namespace nalu_b200::hex8;

int main()
{
  double maxerr = 0;
  auto upd = [&](double a, double b) { maxerr = std::fmax(maxerr, std::fabs(a - b)); };
  const int *lr = orc_hex8_lrscv();
  for (int ip = 0; ip < 12; ++ip)
    if (scs_left(ip) != lr[2 * ip] || scs_right(ip) != lr[2 * ip + 1]) { printf("lrscv mismatch %d\n", ip); return 1; }
  for (int sh = 0; sh < 2; ++sh) {
    double loc[36], locv[24];
    orc_hex8_scs_intg_loc(sh, loc);
    orc_hex8_scv_intg_loc(sh, locv);
    for (int ip = 0; ip < 12; ++ip) {
      double s1, s2, s3;
      scs_intg_loc(ip, sh ? 0.5 : 0.25, s1, s2, s3);
      if (s1 != loc[3 * ip] || s2 != loc[3 * ip + 1] || s3 != loc[3 * ip + 2]) { printf("scs loc mismatch %d\n", ip); return 1; }
    }
    for (int ip = 0; ip < 8; ++ip) {
      double s1, s2, s3;
      scv_intg_loc(ip, sh ? 0.5 : 0.25, s1, s2, s3);
      if (s1 != locv[3 * ip] || s2 != locv[3 * ip + 1] || s3 != locv[3 * ip + 2]) { printf("scv loc mismatch %d\n", ip); return 1; }
    }
    double shp[12][8], der[12][8][3];
    orc_hex8_shape_fcn(12, loc, &shp[0][0]);
    orc_hex8_derivative(12, loc, &der[0][0][0]);
    for (int ip = 0; ip < 12; ++ip)
      for (int n = 0; n < 8; ++n) {
        if (shape_fcn(n, loc[3 * ip], loc[3 * ip + 1], loc[3 * ip + 2]) != shp[ip][n]) { printf("shape mismatch\n"); return 1; }
        double d[3];
        shape_deriv(n, loc[3 * ip], loc[3 * ip + 1], loc[3 * ip + 2], d);
        for (int j = 0; j < 3; ++j)
          if (d[j] != der[ip][n][j]) { printf("deriv mismatch\n"); return 1; }
      }
  }
  srand(12345);
  for (int trial = 0; trial < 200; ++trial) {
    double c[8][3];
    const double amp = trial == 0 ? 0.0 : 0.2;
    for (int n = 0; n < 8; ++n) {
      c[n][0] = 0.5 + 1.3 * sgn_x(n) * 0.5 + amp * (rand() / (double)RAND_MAX - 0.5);
      c[n][1] = 0.5 + 0.9 * sgn_y(n) * 0.5 + amp * (rand() / (double)RAND_MAX - 0.5);
      c[n][2] = 0.5 + 1.1 * sgn_z(n) * 0.5 + amp * (rand() / (double)RAND_MAX - 0.5);
    }
    double areav[12][3], vol[8], g[12][8][3], gs[12][8][3];
    orc_hex8_scs_areav(&c[0][0], &areav[0][0]);
    orc_hex8_scv_volume(&c[0][0], vol);
    orc_hex8_scs_grad_op(0, &c[0][0], &g[0][0][0], nullptr);
    orc_hex8_scs_grad_op(1, &c[0][0], &gs[0][0][0], nullptr);
    for (int ip = 0; ip < 12; ++ip) {
      double a[3];
      scs_areav_ip(c, ip, a);
      for (int d = 0; d < 3; ++d) upd(a[d], areav[ip][d]);
      double s1, s2, s3, dn[8][3];
      scs_intg_loc(ip, 0.25, s1, s2, s3);
      grad_op_ip(c, s1, s2, s3, dn);
      for (int n = 0; n < 8; ++n) for (int d = 0; d < 3; ++d) upd(dn[n][d], g[ip][n][d]);
      scs_intg_loc(ip, 0.5, s1, s2, s3);
      grad_op_ip(c, s1, s2, s3, dn);
      for (int n = 0; n < 8; ++n) for (int d = 0; d < 3; ++d) upd(dn[n][d], gs[ip][n][d]);
    }
    for (int ip = 0; ip < 8; ++ip) upd(scv_volume_ip(c, ip), vol[ip]);
  }
  printf("maxerr %.3e\n", maxerr);
  return maxerr < 1e-13 ? 0 : 2;
}
