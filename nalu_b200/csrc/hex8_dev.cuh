// hex8_dev.cuh -- Hex8 CVFEM master-element operators as device functions.
//
// What they compute is defined by the reference's HexSCS / HexSCV
// (src/master_element/Hex8CVFEM.C, include/master_element/Hex8GeometryFunctions.h,
// include/master_element/MasterElementFunctions.h); how they compute it is not:
//   * shape functions and their parametric derivatives are evaluated in product form
//     N_n = (1/2 + sx_n s1)(1/2 + sy_n s2)(1/2 + sz_n s3) from node sign bits instead of the
//     reference's expanded polynomials / per-element derivative table (Hex8CVFEM.C:24-70);
//     all values are dyadic rationals, so both forms are exact;
//   * an SCS area vector is half the cross product of the facet's diagonals -- identical, as
//     a vector area of a closed polygon, to the reference's four-triangle fan about the facet
//     centroid (Hex8GeometryFunctions.h:30-76) -- and the facet corners are written directly
//     as node averages instead of materialising the 27-point subdivision.
#pragma once
#include <cstdint>

namespace nalu_b200 {
namespace hex8 {

// Exodus / STK Hex8 node signs on the [-1/2,1/2]^3 reference element (Hex8CVFEM.C:396-412)
__host__ __device__ __forceinline__ double sgn_x(int n) { return ((n ^ (n >> 1)) & 1) ? 1.0 : -1.0; }
__host__ __device__ __forceinline__ double sgn_y(int n) { return ((n >> 1) & 1) ? 1.0 : -1.0; }
__host__ __device__ __forceinline__ double sgn_z(int n) { return ((n >> 2) & 1) ? 1.0 : -1.0; }

// left/right node of each sub-control surface (Hex8CVFEM.C:264-276)
__host__ __device__ __forceinline__ int scs_left(int ip)
{
  // {0,1,2,0, 4,5,6,4, 0,1,2,3}
  return ip < 8 ? ((ip & 3) == 3 ? (ip & 4) : ip) : ip - 8;
}
__host__ __device__ __forceinline__ int scs_right(int ip)
{
  // {1,2,3,3, 5,6,7,7, 4,5,6,7}
  return ip < 8 ? ((ip & 3) == 3 ? ip : ip + 1) : ip - 4;
}

// SCS integration point ip, parametric location (Hex8CVFEM.C:318-345); q = 1/4 standard, 1/2 shifted
__host__ __device__ __forceinline__ void scs_intg_loc(int ip, double q, double &s1, double &s2, double &s3)
{
  if (ip < 8) {
    const int e = ip & 3; // edge within the bottom / top face
    s3 = (ip & 4) ? q : -q;
    s1 = (e == 0 || e == 2) ? 0.0 : (e == 1 ? q : -q);
    s2 = (e == 1 || e == 3) ? 0.0 : (e == 0 ? -q : q);
  } else {
    const int n = ip - 8; // vertical edge above bottom node n
    s1 = sgn_x(n) * q;
    s2 = sgn_y(n) * q;
    s3 = 0.0;
  }
}

// SCV integration point = quarter point next to node ip (Hex8CVFEM.C:88-107)
__host__ __device__ __forceinline__ void scv_intg_loc(int ip, double q, double &s1, double &s2, double &s3)
{
  s1 = sgn_x(ip) * q;
  s2 = sgn_y(ip) * q;
  s3 = sgn_z(ip) * q;
}

// trilinear shape function of node n (MasterElementWork.F:2344-2402 hex_shape_fcn)
__host__ __device__ __forceinline__ double shape_fcn(int n, double s1, double s2, double s3)
{
  return (0.5 + sgn_x(n) * s1) * (0.5 + sgn_y(n) * s2) * (0.5 + sgn_z(n) * s3);
}

// parametric derivatives of node n's shape function (Hex8CVFEM.C:24-70 hex8_derivative)
__host__ __device__ __forceinline__ void shape_deriv(int n, double s1, double s2, double s3, double d[3])
{
  const double fx = 0.5 + sgn_x(n) * s1, fy = 0.5 + sgn_y(n) * s2, fz = 0.5 + sgn_z(n) * s3;
  d[0] = sgn_x(n) * fy * fz;
  d[1] = sgn_y(n) * fx * fz;
  d[2] = sgn_z(n) * fx * fy;
}

// generic_grad_op for one integration point (MasterElementFunctions.h:51-108):
// dndx[n][i] = (sum_j adj[i][j] deriv[n][j]) / det.  coords is [8][3].  Returns det.
template <typename CoordT>
__host__ __device__ __forceinline__ double grad_op_ip(const CoordT &coords, double s1, double s2, double s3,
                                             double dndx[8][3])
{
  double jact[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  double deriv[8][3];
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    shape_deriv(n, s1, s2, s3, deriv[n]);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) jact[i][j] += deriv[n][j] * coords[n][i];
  }
  double adj[3][3];
  adj[0][0] = jact[1][1] * jact[2][2] - jact[2][1] * jact[1][2];
  adj[0][1] = jact[1][2] * jact[2][0] - jact[2][2] * jact[1][0];
  adj[0][2] = jact[1][0] * jact[2][1] - jact[2][0] * jact[1][1];
  adj[1][0] = jact[0][2] * jact[2][1] - jact[2][2] * jact[0][1];
  adj[1][1] = jact[0][0] * jact[2][2] - jact[2][0] * jact[0][2];
  adj[1][2] = jact[0][1] * jact[2][0] - jact[2][1] * jact[0][0];
  adj[2][0] = jact[0][1] * jact[1][2] - jact[1][1] * jact[0][2];
  adj[2][1] = jact[0][2] * jact[1][0] - jact[1][2] * jact[0][0];
  adj[2][2] = jact[0][0] * jact[1][1] - jact[1][0] * jact[0][1];
  const double det = jact[0][0] * adj[0][0] + jact[1][0] * adj[1][0] + jact[2][0] * adj[2][0];
  const double inv = 1.0 / det;
#pragma unroll
  for (int n = 0; n < 8; ++n)
#pragma unroll
    for (int i = 0; i < 3; ++i)
      dndx[n][i] = (adj[i][0] * deriv[n][0] + adj[i][1] * deriv[n][1] + adj[i][2] * deriv[n][2]) * inv;
  return det;
}

// Metric tensors at a parametric point: HexSCS::gij (Hex8CVFEM.C:800-808) + generic_gij_3d
// (MasterElementFunctions.h:171-236): g^ij = J J^T with J[i][k] = sum_n dN_n/dxi_k x_n,i, g_ij = its inverse.  Both are
// symmetric: G[0..5] = upper 00 01 02 11 12 22, G[6..11] = lower in the same order.
template <typename CoordT>
__host__ __device__ __forceinline__ void gij_ip(const CoordT &coords, double s1, double s2, double s3, double *G)
{
  double jac[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    double d[3];
    shape_deriv(n, s1, s2, s3, d);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) jac[i][j] += d[j] * coords[n][i];
  }
  const double u00 = jac[0][0] * jac[0][0] + jac[0][1] * jac[0][1] + jac[0][2] * jac[0][2];
  const double u01 = jac[0][0] * jac[1][0] + jac[0][1] * jac[1][1] + jac[0][2] * jac[1][2];
  const double u02 = jac[0][0] * jac[2][0] + jac[0][1] * jac[2][1] + jac[0][2] * jac[2][2];
  const double u11 = jac[1][0] * jac[1][0] + jac[1][1] * jac[1][1] + jac[1][2] * jac[1][2];
  const double u12 = jac[1][0] * jac[2][0] + jac[1][1] * jac[2][1] + jac[1][2] * jac[2][2];
  const double u22 = jac[2][0] * jac[2][0] + jac[2][1] * jac[2][1] + jac[2][2] * jac[2][2];
  const double inv = 1.0 / (u00 * (u11 * u22 - u12 * u12) - u01 * (u01 * u22 - u12 * u02) + u02 * (u01 * u12 - u11 * u02));
  G[0] = u00; G[1] = u01; G[2] = u02; G[3] = u11; G[4] = u12; G[5] = u22;
  G[6] = inv * (u11 * u22 - u12 * u12);
  G[7] = inv * (u02 * u12 - u01 * u22);
  G[8] = inv * (u01 * u12 - u02 * u11);
  G[9] = inv * (u00 * u22 - u02 * u02);
  G[10] = inv * (u02 * u01 - u00 * u12);
  G[11] = inv * (u00 * u11 - u01 * u01);
}
// entry (i, j) of a symmetric 3x3 tensor stored as 00 01 02 11 12 22
__host__ __device__ __forceinline__ double sym3(const double *g, int i, int j)
{
  const int a = i < j ? i : j, b = i < j ? j : i;
  return g[a == 0 ? b : (a == 1 ? 2 + b : 5)];
}

// Physical position of the trilinear map at a parametric point (node averages when the
// point is an edge / face / body midpoint)
template <typename CoordT>
__host__ __device__ __forceinline__ void map_point(const CoordT &coords, double s1, double s2, double s3, double x[3])
{
  x[0] = x[1] = x[2] = 0.0;
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const double w = shape_fcn(n, s1, s2, s3);
    x[0] += w * coords[n][0];
    x[1] += w * coords[n][1];
    x[2] += w * coords[n][2];
  }
}

// SCS area vector of integration point ip (HexSCS::determinant, Hex8CVFEM.C:547-581).
// The facet of edge (L,R) has corners: edge midpoint, the two adjacent face centroids and the
// element centroid; orientation such that the vector points from L to R.
template <typename CoordT>
__host__ __device__ __forceinline__ void scs_areav_ip(const CoordT &coords, int ip, double a[3])
{
  double e1, e2, e3; // parametric location of the edge midpoint: the ip location at q = 1/2
  scs_intg_loc(ip, 0.5, e1, e2, e3);
  // edge direction axis: the one whose midpoint coordinate is 0
  // facet corners in parametric space: m=(e), f1 = e with one transverse coord zeroed,
  // f2 = e with the other transverse coord zeroed, c = 0
  double m[3], f1[3], f2[3], c[3];
  map_point(coords, e1, e2, e3, m);
  map_point(coords, 0.0, 0.0, 0.0, c);
  double t1[3], t2[3]; // transverse parametric points
  if (e1 == 0.0) {        // x-edge: transverse (y,z)
    t1[0] = 0; t1[1] = e2; t1[2] = 0;  // face centroid with z zeroed
    t2[0] = 0; t2[1] = 0;  t2[2] = e3; // face centroid with y zeroed
  } else if (e2 == 0.0) { // y-edge: transverse (z,x)
    t1[0] = 0;  t1[1] = 0; t1[2] = e3;
    t2[0] = e1; t2[1] = 0; t2[2] = 0;
  } else {                // z-edge: transverse (x,y)
    t1[0] = e1; t1[1] = 0;  t1[2] = 0;
    t2[0] = 0;  t2[1] = e2; t2[2] = 0;
  }
  map_point(coords, t1[0], t1[1], t1[2], f1);
  map_point(coords, t2[0], t2[1], t2[2], f2);
  // quad (m, f1, c, f2): vector area = 1/2 (c - m) x (f2 - f1); for a right-handed triple
  // (edge, t1, t2) with both transverse coordinates of m negative this points along -edge.
  const double d1[3] = {c[0] - m[0], c[1] - m[1], c[2] - m[2]};
  const double d2[3] = {f2[0] - f1[0], f2[1] - f1[1], f2[2] - f1[2]};
  double ax = 0.5 * (d1[1] * d2[2] - d1[2] * d2[1]);
  double ay = 0.5 * (d1[2] * d2[0] - d1[0] * d2[2]);
  double az = 0.5 * (d1[0] * d2[1] - d1[1] * d2[0]);
  // sign: the product of the two transverse midpoint coordinates tells which quadrant the facet
  // sits in; the left->right direction along the edge axis is fixed by lrscv
  double sq, dir;
  if (e1 == 0.0) { sq = e2 * e3; dir = sgn_x(scs_right(ip)) ; }
  else if (e2 == 0.0) { sq = e3 * e1; dir = sgn_y(scs_right(ip)); }
  else { sq = e1 * e2; dir = sgn_z(scs_right(ip)); }
  const double s = (sq > 0.0 ? -1.0 : 1.0) * dir;
  a[0] = s * ax;
  a[1] = s * ay;
  a[2] = s * az;
}

// Grandy volume of the sub-control volume next to node ip (HexSCV::determinant,
// Hex8CVFEM.C:149-178; hex_volume_grandy, Hex8GeometryFunctions.h:78-163).
// The sub-hex is the image of the parametric octant of node ip; its 8 corners in the
// reference's subDivisionTable order are the octant corners in Hex8 node order.
template <typename CoordT>
__host__ __device__ __forceinline__ double scv_volume_ip(const CoordT &coords, int ip)
{
  const double ox = sgn_x(ip) * 0.25, oy = sgn_y(ip) * 0.25, oz = sgn_z(ip) * 0.25; // octant centre
  double p[14][3];
#pragma unroll
  for (int n = 0; n < 8; ++n)
    map_point(coords, ox + 0.25 * sgn_x(n), oy + 0.25 * sgn_y(n), oz + 0.25 * sgn_z(n), p[n]);
  const int face_nodes[6][4] = {{0, 3, 2, 1}, {4, 5, 6, 7}, {0, 1, 5, 4}, {2, 3, 7, 6}, {1, 2, 6, 5}, {0, 4, 3, 7}};
#pragma unroll
  for (int k = 0; k < 6; ++k)
#pragma unroll
    for (int d = 0; d < 3; ++d)
      p[8 + k][d] = 0.25 * (p[face_nodes[k][0]][d] + p[face_nodes[k][1]][d] + p[face_nodes[k][2]][d] +
                            p[face_nodes[k][3]][d]);
  const int tri[24][3] = {{0, 8, 1},  {8, 2, 1},  {3, 2, 8},  {3, 8, 0},  {6, 9, 5},  {7, 9, 6},
                          {4, 9, 7},  {4, 5, 9},  {10, 0, 1}, {5, 10, 1}, {4, 10, 5}, {4, 0, 10},
                          {7, 6, 11}, {6, 2, 11}, {2, 3, 11}, {3, 7, 11}, {6, 12, 2}, {5, 12, 6},
                          {5, 1, 12}, {1, 2, 12}, {0, 4, 13}, {4, 7, 13}, {7, 3, 13}, {3, 0, 13}};
  double vol = 0.0;
#pragma unroll
  for (int k = 0; k < 24; ++k) {
    const double *P = p[tri[k][0]], *Q = p[tri[k][1]], *R = p[tri[k][2]];
    const double m0 = P[0] + Q[0] + R[0], m1 = P[1] + Q[1] + R[1], m2 = P[2] + Q[2] + R[2];
    const double q0 = Q[0] - P[0], q1 = Q[1] - P[1], q2 = Q[2] - P[2];
    const double r0 = R[0] - P[0], r1 = R[1] - P[1], r2 = R[2] - P[2];
    vol += m0 * (q1 * r2 - r1 * q2) + m1 * (r0 * q2 - q0 * r2) + m2 * (q0 * r1 - r0 * q1);
  }
  return vol / 18.0;
}

} // namespace hex8
} // namespace nalu_b200
