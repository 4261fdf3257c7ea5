// scalar_tile.cuh -- tiled, write-once ScalarAdvDiffElemKernel / ScalarDiffElemKernel assembly (one dof per node)
// for structured Hex8 blocks: src/kernel/ScalarAdvDiffElemKernel.C:69-127, src/kernel/ScalarDiffElemKernel.C:64-110
// through AssembleElemSolverAlgorithm::execute + TpetraLinearSystem::sumInto.  Same work decomposition as
// cont_tile.cuh (thread = element in phase 1, thread = node row with 54 register accumulators in phase 2, every
// CSR value written once, no atomics); the per-point record is
//   b_k = -Gamma_ip (sum_i adj(J)[i][k] A_i)/det J   so that lhsfacDiff(ic) = dN_ic/dxi . b,
//   mdot (0 for the pure diffusion kernel)           so that lhsfacAdv(ic)  = N_adv,ic mdot,
//   s = sum_ic (lhsfacAdv + lhsfacDiff) q_ic = mdot q_ip + b . dq/dxi   (the rhs flux, :121-124).
#pragma once
#include "cont_tile.cuh"

namespace nalu_b200 {
namespace scal {

using box::bit_x;
using box::bit_y;
using box::bit_z;
using box::ip_of;
using box::node_of_bits;
using cont::copy8_or_zero;
using cont::grad_dot_area;
using cont::stage_wait;

struct ScalarParams {
  int nx, ny, nz;    // elements of the local block
  int NX, NY, NZ;    // nodes
  int z_chunk, tiles_x, tiles_y;
  int z_first, z_last;
  const double *coords, *q, *gam; // [nNodes][3], [1] scalarQ, [1] diffFluxCoeff
  const double *mdot;             // [nElem][12] mass_flow_rate_scs; null: ScalarDiffElemKernel
  const int *node_row;
  const long long *browptr;
  const unsigned char *slot27;
  double *values;
  double *rhs;
  int *err;
  int options;     // NALU_OPT_SHIFTED_GRAD_OP | NALU_OPT_SKEW_SYMMETRIC
  int overwrite;
};

struct Cfg {
  static constexpr int TBX = 15, TBY = 7;
  static constexpr int ELX = 16, ELY = 8, NEL = ELX * ELY;
  static constexpr int NT = 128;
  static constexpr int RNX = ELX + 1, RNY = ELY + 1, RN = RNX * RNY;
  static constexpr int NF = 5;                            // x[3], q, Gamma
  static constexpr int PLN = RN * NF;
  static constexpr int RD = 6;                            // record: b[3], mdot, s, pad
  static constexpr int ES = 12 * RD + 2;                  // element stride: 592 bytes = 80 mod 128
  static constexpr int OFF_REC = 0, OFF_RING = NEL * ES;
  static constexpr int SMEM_DOUBLES = OFF_RING + 3 * PLN;
};

__host__ __device__ inline void stage_plane(const ScalarParams &P, double *ring_plane, int tid, int x0, int y0, int zp)
{
  using C = Cfg;
  const long long pn = (long long)P.NX * P.NY;
  const bool zin = zp >= 0 && zp < P.NZ;
#pragma unroll 1
  for (int rn = tid; rn < C::RN; rn += C::NT) {
    const int x = x0 - 1 + rn % C::RNX, y = y0 - 1 + rn / C::RNX;
    const bool valid = zin && x >= 0 && y >= 0 && x < P.NX && y < P.NY;
    const long long n = valid ? x + (long long)P.NX * y + pn * zp : 0;
    double *d = ring_plane + rn * C::NF;
    const double *xc = P.coords + n * 3;
    copy8_or_zero(d + 0, xc + 0, valid);
    copy8_or_zero(d + 1, xc + 1, valid);
    copy8_or_zero(d + 2, xc + 2, valid);
    copy8_or_zero(d + 3, P.q + n, valid);
    copy8_or_zero(d + 4, P.gam + n, valid);
  }
}

// ---- phase 1: thread = element el of layer ez ---------------------------------------------------------------
__host__ __device__ inline void p1_element(const ScalarParams &P, double *sm, int el, int x0, int y0, int ez, const double *pl0,
                                           const double *pl1)
{
  using C = Cfg;
  const int elx = el % C::ELX, ely = el / C::ELX;
  const int ex = x0 - 1 + elx, ey = y0 - 1 + ely;
  double *rec = sm + C::OFF_REC + el * C::ES;
  if (ex < 0 || ey < 0 || ex >= P.nx || ey >= P.ny || ez < 0 || ez >= P.nz) {
#pragma unroll
    for (int i = 0; i < 12 * C::RD; ++i) rec[i] = 0.0;
    return;
  }
  const double *md = P.mdot ? P.mdot + 12 * (ex + (long long)P.nx * (ey + (long long)P.ny * ez)) : nullptr;
  double c[8][3], qn[8], gn[8];
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const double *nd = (bit_z(n) ? pl1 : pl0) + ((elx + bit_x(n)) + C::RNX * (ely + bit_y(n))) * C::NF;
    c[n][0] = nd[0];
    c[n][1] = nd[1];
    c[n][2] = nd[2];
    qn[n] = nd[3];
    gn[n] = nd[4];
  }
  const double qg = (P.options & NALU_OPT_SHIFTED_GRAD_OP) ? 0.5 : 0.25, qa = (P.options & NALU_OPT_SKEW_SYMMETRIC) ? 0.5 : 0.25;
#pragma unroll 1
  for (int ip = 0; ip < 12; ++ip) {
    double A[3], b[3], s1, s2, s3;
    hex8::scs_areav_ip(c, ip, A);
    hex8::scs_intg_loc(ip, qg, s1, s2, s3);
    const double det = grad_dot_area(c, s1, s2, s3, A, b);
    if (!(det > 0.0)) *P.err = 1; // MasterElementFunctions.h:91-94
    double dq[3] = {0, 0, 0};
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      double d[3];
      hex8::shape_deriv(n, s1, s2, s3, d);
      dq[0] += d[0] * qn[n];
      dq[1] += d[1] * qn[n];
      dq[2] += d[2] * qn[n];
    }
    // diffFluxCoeffIp with the (never shifted) shape function, q_ip with the advection shape function (:92-110)
    hex8::scs_intg_loc(ip, 0.25, s1, s2, s3);
    double gam = 0.0;
#pragma unroll
    for (int n = 0; n < 8; ++n) gam += hex8::shape_fcn(n, s1, s2, s3) * gn[n];
    hex8::scs_intg_loc(ip, qa, s1, s2, s3);
    double qip = 0.0;
#pragma unroll
    for (int n = 0; n < 8; ++n) qip += hex8::shape_fcn(n, s1, s2, s3) * qn[n];
    const double tmdot = md ? md[ip] : 0.0;
    const double b0 = -gam * b[0], b1 = -gam * b[1], b2 = -gam * b[2];
    rec[ip * C::RD + 0] = b0;
    rec[ip * C::RD + 1] = b1;
    rec[ip * C::RD + 2] = b2;
    rec[ip * C::RD + 3] = tmdot;
    rec[ip * C::RD + 4] = tmdot * qip + (b0 * dq[0] + b1 * dq[1] + b2 * dq[2]);
    rec[ip * C::RD + 5] = 0.0;
  }
}

// ---- phase 1 in coefficient form (cont_tile.cuh: Walsh-Hadamard coefficients of the trilinear fields) -----------
struct ScalElem {
  double X[3][8], Q[8], G[8];
  double qg, qa;
  const double *md;
};
template <int T, int SU, int SV>
__host__ __device__ __forceinline__ void p1_point_c(const ScalarParams &P, const ScalElem &E, double *rec)
{
  using C = Cfg;
  constexpr int U = (T + 1) % 3, V = (T + 2) % 3;
  double orient;
  const int ip = cont::point_index<T, SU, SV>(orient);
  double A[3], b[3];
  cont::area_plus_t<T, SU, SV>(E.X, A);
  A[0] *= orient;
  A[1] *= orient;
  A[2] *= orient;
  const double det = cont::grad_dot_area_c<T, SU, SV>(E.X, E.qg, A, b);
  if (!(det > 0.0)) *P.err = 1; // MasterElementFunctions.h:91-94
  double dq[3];
  cont::grad_c<T, SU, SV>(E.Q, E.qg, dq);
  const double gam = cont::value_c<T, SU, SV>(E.G, 0.25);  // diffFluxCoeffIp: never shifted
  const double qip = cont::value_c<T, SU, SV>(E.Q, E.qa);  // advection shape function
  const double tmdot = E.md ? E.md[ip] : 0.0;
  const double b0 = -gam * b[0], b1 = -gam * b[1], b2 = -gam * b[2];
  rec[ip * C::RD + T] = b0;
  rec[ip * C::RD + U] = b1;
  rec[ip * C::RD + V] = b2;
  rec[ip * C::RD + 3] = tmdot;
  rec[ip * C::RD + 4] = tmdot * qip + (b0 * dq[0] + b1 * dq[1] + b2 * dq[2]);
  rec[ip * C::RD + 5] = 0.0;
}
template <int T>
__host__ __device__ __forceinline__ void p1_edge_dir_c(const ScalarParams &P, const ScalElem &E, double *rec)
{
  p1_point_c<T, 0, 0>(P, E, rec);
  p1_point_c<T, 1, 0>(P, E, rec);
  p1_point_c<T, 0, 1>(P, E, rec);
  p1_point_c<T, 1, 1>(P, E, rec);
}
__host__ __device__ inline void p1_element_c(const ScalarParams &P, double *sm, int el, int x0, int y0, int ez, const double *pl0,
                                             const double *pl1)
{
  using C = Cfg;
  const int elx = el % C::ELX, ely = el / C::ELX;
  const int ex = x0 - 1 + elx, ey = y0 - 1 + ely;
  double *rec = sm + C::OFF_REC + el * C::ES;
  if (ex < 0 || ey < 0 || ex >= P.nx || ey >= P.ny || ez < 0 || ez >= P.nz) {
#pragma unroll
    for (int i = 0; i < 12 * C::RD; ++i) rec[i] = 0.0;
    return;
  }
  const int n000 = (elx + C::RNX * ely) * C::NF;
  ScalElem E;
  E.md = P.mdot ? P.mdot + 12 * (ex + (long long)P.nx * (ey + (long long)P.ny * ez)) : nullptr;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int m = 0; m < 8; ++m)
      E.X[i][m] = ((m & 4) ? pl1 : pl0)[n000 + ((m & 1) + C::RNX * ((m >> 1) & 1)) * C::NF + i];
    const double o = E.X[i][0]; // relative to the element's first node (see cont_tile.cuh)
#pragma unroll
    for (int m = 0; m < 8; ++m) E.X[i][m] -= o;
    box::butterfly8(E.X[i]);
  }
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const double *nd = ((m & 4) ? pl1 : pl0) + n000 + ((m & 1) + C::RNX * ((m >> 1) & 1)) * C::NF;
    E.Q[m] = nd[3];
    E.G[m] = nd[4];
  }
  box::butterfly8(E.Q);
  box::butterfly8(E.G);
  E.qg = (P.options & NALU_OPT_SHIFTED_GRAD_OP) ? 0.5 : 0.25;
  E.qa = (P.options & NALU_OPT_SKEW_SYMMETRIC) ? 0.5 : 0.25;
  p1_edge_dir_c<0>(P, E, rec);
  p1_edge_dir_c<1>(P, E, rec);
  p1_edge_dir_c<2>(P, E, rec);
}

// ---- phase 2: thread = node (lx, ly) of the tile (see cont_tile.cuh) ----------------------------------------
// SH: shifted grad-op, SK: skew-symmetric (shifted advection shape function)
template <bool SH, bool SK, int SX, int SY, int DIR, int LVL>
__host__ __device__ __forceinline__ void p2_point(const double *er, double (&acc)[27], double &r)
{
  using C = Cfg;
  constexpr int bx = !SX, by = !SY, bz = LVL;
  const int k = node_of_bits(bx, by, bz);
  const int ip = ip_of(k, DIR);
  const double sgn = hex8::scs_left(ip) == k ? 1.0 : -1.0;
  double s1, s2, s3, a1, a2, a3;
  hex8::scs_intg_loc(ip, SH ? 0.5 : 0.25, s1, s2, s3);
  hex8::scs_intg_loc(ip, SK ? 0.5 : 0.25, a1, a2, a3);
  const double b0 = er[ip * C::RD + 0], b1 = er[ip * C::RD + 1], b2 = er[ip * C::RD + 2], md = er[ip * C::RD + 3],
               s = er[ip * C::RD + 4];
#pragma unroll
  for (int ic = 0; ic < 8; ++ic) {
    double d[3];
    hex8::shape_deriv(ic, s1, s2, s3, d);
    const int idx = (bit_z(ic) - bz + 1) * 9 + (bit_y(ic) - by + 1) * 3 + (bit_x(ic) - bx + 1);
    acc[idx] += (sgn * d[0]) * b0;
    acc[idx] += (sgn * d[1]) * b1;
    acc[idx] += (sgn * d[2]) * b2;
    acc[idx] += (sgn * hex8::shape_fcn(ic, a1, a2, a3)) * md;
  }
  r -= sgn * s;
}
template <bool SH, bool SK, int SX, int SY>
__host__ __device__ __forceinline__ void p2_element(const double *sm, int lx, int ly, double (&accA)[27], double &rA,
                                                    double (&accB)[27], double &rB)
{
  using C = Cfg;
  const double *er = sm + C::OFF_REC + ((lx + SX) + C::ELX * (ly + SY)) * C::ES;
  p2_point<SH, SK, SX, SY, 0, 0>(er, accA, rA);
  p2_point<SH, SK, SX, SY, 1, 0>(er, accA, rA);
  p2_point<SH, SK, SX, SY, 2, 0>(er, accA, rA);
  p2_point<SH, SK, SX, SY, 0, 1>(er, accB, rB);
  p2_point<SH, SK, SX, SY, 1, 1>(er, accB, rB);
  p2_point<SH, SK, SX, SY, 2, 1>(er, accB, rB);
}
template <bool SH, bool SK>
__host__ __device__ __forceinline__ void p2_layer(const double *sm, int lx, int ly, double (&accA)[27], double &rA,
                                                  double (&accB)[27], double &rB)
{
  p2_element<SH, SK, 0, 0>(sm, lx, ly, accA, rA, accB, rB);
  p2_element<SH, SK, 1, 0>(sm, lx, ly, accA, rA, accB, rB);
  p2_element<SH, SK, 0, 1>(sm, lx, ly, accA, rA, accB, rB);
  p2_element<SH, SK, 1, 1>(sm, lx, ly, accA, rA, accB, rB);
}

} // namespace scal
} // namespace nalu_b200
