/*
 * hex8_me.c -- oracle restatement of the Hex8 CVFEM master elements.
 * TEST INFRASTRUCTURE ONLY (see nalu_oracle.h).  Scalar, reference loop order.
 */
#include "nalu_oracle.h"
#include <string.h>

/* src/master_element/Hex8CVFEM.C:264-276 */
static const int lrscv_[24] = {0, 1, 1, 2, 2, 3, 0, 3, 4, 5, 5, 6, 6, 7, 4, 7, 0, 4, 1, 5, 2, 6, 3, 7};
/* src/master_element/Hex8CVFEM.C:83-85 */
static const int scvIpNodeMap_[8] = {0, 1, 2, 3, 4, 5, 6, 7};

const int *orc_hex8_lrscv(void) { return lrscv_; }
const int *orc_hex8_scv_ip_node_map(void) { return scvIpNodeMap_; }

/* src/master_element/Hex8CVFEM.C:318-345 */
void orc_hex8_scs_intg_loc(int shifted, double loc[36])
{
  static const double base[36] = {
      0.00,  -0.25, -0.25, /* surf 1    1->2 */
      0.25,  0.00,  -0.25, /* surf 2    2->3 */
      0.00,  0.25,  -0.25, /* surf 3    3->4 */
      -0.25, 0.00,  -0.25, /* surf 4    1->4 */
      0.00,  -0.25, 0.25,  /* surf 5    5->6 */
      0.25,  0.00,  0.25,  /* surf 6    6->7 */
      0.00,  0.25,  0.25,  /* surf 7    7->8 */
      -0.25, 0.00,  0.25,  /* surf 8    5->8 */
      -0.25, -0.25, 0.00,  /* surf 9    1->5 */
      0.25,  -0.25, 0.00,  /* surf 10   2->6 */
      0.25,  0.25,  0.00,  /* surf 11   3->7 */
      -0.25, 0.25,  0.00   /* surf 12   4->8 */
  };
  /* the shifted table (:333-345) is the same pattern with 0.25 -> 0.50 */
  for (int i = 0; i < 36; ++i) loc[i] = shifted ? 2.0 * base[i] : base[i];
}

/* src/master_element/Hex8CVFEM.C:88-107 */
void orc_hex8_scv_intg_loc(int shifted, double loc[24])
{
  static const double base[24] = {-0.25, -0.25, -0.25, +0.25, -0.25, -0.25, +0.25, +0.25,
                                  -0.25, -0.25, +0.25, -0.25, -0.25, -0.25, +0.25, +0.25,
                                  -0.25, +0.25, +0.25, +0.25, +0.25, -0.25, +0.25, +0.25};
  for (int i = 0; i < 24; ++i) loc[i] = shifted ? 2.0 * base[i] : base[i];
}

/* src/master_element/MasterElementWork.F:2344-2402 (hex_shape_fcn) */
void orc_hex8_shape_fcn(int npts, const double *par_coord, double *shape)
{
  const double half = 1.0 / 2.0, one4th = 1.0 / 4.0, one8th = 1.0 / 8.0;
  for (int j = 0; j < npts; ++j) {
    const double s1 = par_coord[3 * j + 0];
    const double s2 = par_coord[3 * j + 1];
    const double s3 = par_coord[3 * j + 2];
    double *f = shape + 8 * j;
    f[0] = one8th + one4th * (-s1 - s2 - s3) + half * (s2 * s3 + s3 * s1 + s1 * s2) - s1 * s2 * s3;
    f[1] = one8th + one4th * (s1 - s2 - s3) + half * (s2 * s3 - s3 * s1 - s1 * s2) + s1 * s2 * s3;
    f[2] = one8th + one4th * (s1 + s2 - s3) + half * (-s2 * s3 - s3 * s1 + s1 * s2) - s1 * s2 * s3;
    f[3] = one8th + one4th * (-s1 + s2 - s3) + half * (-s2 * s3 + s3 * s1 - s1 * s2) + s1 * s2 * s3;
    f[4] = one8th + one4th * (-s1 - s2 + s3) + half * (-s2 * s3 - s3 * s1 + s1 * s2) + s1 * s2 * s3;
    f[5] = one8th + one4th * (s1 - s2 + s3) + half * (-s2 * s3 + s3 * s1 - s1 * s2) - s1 * s2 * s3;
    f[6] = one8th + one4th * (s1 + s2 + s3) + half * (s2 * s3 + s3 * s1 + s1 * s2) + s1 * s2 * s3;
    f[7] = one8th + one4th * (-s1 + s2 + s3) + half * (s2 * s3 - s3 * s1 - s1 * s2) - s1 * s2 * s3;
  }
}

/* src/master_element/Hex8CVFEM.C:24-70 (hex8_derivative) */
void orc_hex8_derivative(int npts, const double *intgLoc, double *deriv)
{
  const double half = 0.50, one4th = 0.25;
#define D(ip, n, d) deriv[((ip)*8 + (n)) * 3 + (d)]
  for (int ip = 0; ip < npts; ++ip) {
    const double s1 = intgLoc[ip * 3];
    const double s2 = intgLoc[ip * 3 + 1];
    const double s3 = intgLoc[ip * 3 + 2];
    const double s1s2 = s1 * s2;
    const double s2s3 = s2 * s3;
    const double s1s3 = s1 * s3;

    D(ip, 0, 0) = half * (s3 + s2) - s2s3 - one4th;
    D(ip, 1, 0) = half * (-s3 - s2) + s2s3 + one4th;
    D(ip, 2, 0) = half * (-s3 + s2) - s2s3 + one4th;
    D(ip, 3, 0) = half * (s3 - s2) + s2s3 - one4th;
    D(ip, 4, 0) = half * (-s3 + s2) + s2s3 - one4th;
    D(ip, 5, 0) = half * (s3 - s2) - s2s3 + one4th;
    D(ip, 6, 0) = half * (s3 + s2) + s2s3 + one4th;
    D(ip, 7, 0) = half * (-s3 - s2) - s2s3 - one4th;

    D(ip, 0, 1) = half * (s3 + s1) - s1s3 - one4th;
    D(ip, 1, 1) = half * (s3 - s1) + s1s3 - one4th;
    D(ip, 2, 1) = half * (-s3 + s1) - s1s3 + one4th;
    D(ip, 3, 1) = half * (-s3 - s1) + s1s3 + one4th;
    D(ip, 4, 1) = half * (-s3 + s1) + s1s3 - one4th;
    D(ip, 5, 1) = half * (-s3 - s1) - s1s3 - one4th;
    D(ip, 6, 1) = half * (s3 + s1) + s1s3 + one4th;
    D(ip, 7, 1) = half * (s3 - s1) - s1s3 + one4th;

    D(ip, 0, 2) = half * (s2 + s1) - s1s2 - one4th;
    D(ip, 1, 2) = half * (s2 - s1) + s1s2 - one4th;
    D(ip, 2, 2) = half * (-s2 - s1) - s1s2 - one4th;
    D(ip, 3, 2) = half * (-s2 + s1) + s1s2 - one4th;
    D(ip, 4, 2) = half * (-s2 - s1) + s1s2 + one4th;
    D(ip, 5, 2) = half * (-s2 + s1) - s1s2 + one4th;
    D(ip, 6, 2) = half * (s2 + s1) + s1s2 + one4th;
    D(ip, 7, 2) = half * (s2 - s1) - s1s2 + one4th;
  }
#undef D
}

/* include/master_element/MasterElementFunctions.h:31-43 */
static void cofactorMatrix3(double adjJac[3][3], const double jact[3][3])
{
  adjJac[0][0] = jact[1][1] * jact[2][2] - jact[2][1] * jact[1][2];
  adjJac[0][1] = jact[1][2] * jact[2][0] - jact[2][2] * jact[1][0];
  adjJac[0][2] = jact[1][0] * jact[2][1] - jact[2][0] * jact[1][1];

  adjJac[1][0] = jact[0][2] * jact[2][1] - jact[2][2] * jact[0][1];
  adjJac[1][1] = jact[0][0] * jact[2][2] - jact[2][0] * jact[0][2];
  adjJac[1][2] = jact[0][1] * jact[2][0] - jact[2][1] * jact[0][0];

  adjJac[2][0] = jact[0][1] * jact[1][2] - jact[1][1] * jact[0][2];
  adjJac[2][1] = jact[0][2] * jact[1][0] - jact[1][2] * jact[0][0];
  adjJac[2][2] = jact[0][0] * jact[1][1] - jact[1][0] * jact[0][1];
}

/* include/master_element/MasterElementFunctions.h:51-108 (generic_grad_op<AlgTraitsHex8>) */
int orc_generic_grad_op(int npts, const double *deriv, const double *coords, double *gradop,
                        double *det_j)
{
  int nbad = 0;
  for (int ip = 0; ip < npts; ++ip) {
    double jact[3][3];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) jact[i][j] = 0.0;

    double refGrad[8][3];
    for (int n = 0; n < 8; ++n) {
      for (int i = 0; i < 3; ++i) refGrad[n][i] = deriv[(ip * 8 + n) * 3 + i];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) jact[i][j] += refGrad[n][j] * coords[n * 3 + i];
    }

    double adjJac[3][3];
    cofactorMatrix3(adjJac, jact);

    double det = 0.0;
    for (int i = 0; i < 3; ++i) det += jact[i][0] * adjJac[i][0];
    if (!(det > 1.0e-300)) ++nbad; /* reference: STK_ThrowAssertMsg(det > tiny_positive_value()) */
    if (det_j) det_j[ip] = det;

    const double inv_detj = 1.0 / det;

    for (int n = 0; n < 8; ++n) {
      for (int i = 0; i < 3; ++i) {
        double w = 0.0;
        for (int j = 0; j < 3; ++j) w += adjJac[i][j] * refGrad[n][j];
        w *= inv_detj;
        gradop[(ip * 8 + n) * 3 + i] = w;
      }
    }
  }
  return nbad;
}

/* include/master_element/Hex8GeometryFunctions.h:263-321 */
void orc_subdivide_hex_8(const double *c, double coordv[27][3])
{
#define C(n, d) c[(n)*3 + (d)]
  for (int n = 0; n < 8; ++n) {
    coordv[n][0] = C(n, 0);
    coordv[n][1] = C(n, 1);
    coordv[n][2] = C(n, 2);
  }
  for (int d = 0; d < 3; ++d) {
    /* Face 0 */
    coordv[8][d] = 0.5 * (C(0, d) + C(1, d));
    coordv[9][d] = 0.5 * (C(1, d) + C(2, d));
    coordv[10][d] = 0.5 * (C(2, d) + C(3, d));
    coordv[11][d] = 0.5 * (C(3, d) + C(0, d));
    coordv[12][d] = 0.25 * (C(0, d) + C(1, d) + C(2, d) + C(3, d));
    /* Face 1 */
    coordv[13][d] = 0.5 * (C(4, d) + C(5, d));
    coordv[14][d] = 0.5 * (C(5, d) + C(6, d));
    coordv[15][d] = 0.5 * (C(6, d) + C(7, d));
    coordv[16][d] = 0.5 * (C(7, d) + C(4, d));
    coordv[17][d] = 0.25 * (C(4, d) + C(5, d) + C(6, d) + C(7, d));
    /* Face 2 */
    coordv[18][d] = 0.5 * (C(1, d) + C(5, d));
    coordv[19][d] = 0.5 * (C(0, d) + C(4, d));
    coordv[20][d] = 0.25 * (C(0, d) + C(1, d) + C(4, d) + C(5, d));
    /* Face 3 */
    coordv[21][d] = 0.5 * (C(3, d) + C(7, d));
    coordv[22][d] = 0.5 * (C(2, d) + C(6, d));
    coordv[23][d] = 0.25 * (C(2, d) + C(3, d) + C(6, d) + C(7, d));
    /* Face 4 */
    coordv[24][d] = 0.25 * (C(1, d) + C(2, d) + C(5, d) + C(6, d));
    /* Face 5 */
    coordv[25][d] = 0.25 * (C(0, d) + C(3, d) + C(4, d) + C(7, d));
    /* Volume centroid */
    coordv[26][d] = 0.;
    for (int n = 0; n < 8; ++n) coordv[26][d] += C(n, d);
    coordv[26][d] *= 0.125;
  }
#undef C
}

/* include/master_element/Hex8GeometryFunctions.h:30-76 */
static void quad_area_by_triangulation(int ics, const double areacoords[4][3], double *area)
{
  area[ics * 3 + 0] = 0.0;
  area[ics * 3 + 1] = 0.0;
  area[ics * 3 + 2] = 0.0;

  const double xmid[3] = {
      0.25 * (areacoords[0][0] + areacoords[1][0] + areacoords[2][0] + areacoords[3][0]),
      0.25 * (areacoords[0][1] + areacoords[1][1] + areacoords[2][1] + areacoords[3][1]),
      0.25 * (areacoords[0][2] + areacoords[1][2] + areacoords[2][2] + areacoords[3][2])};

  double r1[3] = {areacoords[0][0] - xmid[0], areacoords[0][1] - xmid[1], areacoords[0][2] - xmid[2]};
  for (int itriangle = 0; itriangle < 4; ++itriangle) {
    const int t_index = (itriangle + 1) % 4;
    const double r2[3] = {areacoords[t_index][0] - xmid[0], areacoords[t_index][1] - xmid[1],
                          areacoords[t_index][2] - xmid[2]};

    area[ics * 3 + 0] += r1[1] * r2[2] - r2[1] * r1[2];
    area[ics * 3 + 1] += r1[2] * r2[0] - r2[2] * r1[0];
    area[ics * 3 + 2] += r1[0] * r2[1] - r2[0] * r1[1];

    r1[0] = r2[0];
    r1[1] = r2[1];
    r1[2] = r2[2];
  }
  area[ics * 3 + 0] *= 0.5;
  area[ics * 3 + 1] *= 0.5;
  area[ics * 3 + 2] *= 0.5;
}

/* src/master_element/Hex8CVFEM.C:547-581 (HexSCS::determinant, SharedMemView overload) */
void orc_hex8_scs_areav(const double *coords, double *areav)
{
  static const int hex_edge_facet_table[12][4] = {
      {20, 8, 12, 26},  {24, 9, 12, 26},  {10, 12, 26, 23}, {11, 25, 26, 12},
      {13, 20, 26, 17}, {17, 14, 24, 26}, {17, 15, 23, 26}, {16, 17, 26, 25},
      {19, 20, 26, 25}, {20, 18, 24, 26}, {22, 23, 26, 24}, {21, 25, 26, 23}};

  double coordv[27][3];
  orc_subdivide_hex_8(coords, coordv);

  for (int ics = 0; ics < 12; ++ics) {
    double scscoords[4][3];
    for (int inode = 0; inode < 4; ++inode) {
      const int itrianglenode = hex_edge_facet_table[ics][inode];
      for (int d = 0; d < 3; ++d) scscoords[inode][d] = coordv[itrianglenode][d];
    }
    quad_area_by_triangulation(ics, scscoords, areav);
  }
}

/* include/master_element/Hex8GeometryFunctions.h:78-163 (hex_volume_grandy) */
double orc_hex_volume_grandy(const double scvcoords[8][3])
{
  double coordv[14][3];
  for (int n = 0; n < 8; ++n) {
    coordv[n][0] = scvcoords[n][0];
    coordv[n][1] = scvcoords[n][1];
    coordv[n][2] = scvcoords[n][2];
  }
  static const int face_nodes[6][4] = {{0, 3, 2, 1}, {4, 5, 6, 7}, {0, 1, 5, 4},
                                       {2, 3, 7, 6}, {1, 2, 6, 5}, {0, 4, 3, 7}};
  for (int k = 0; k < 6; ++k) {
    const int coordIndex = k + 8;
    for (int d = 0; d < 3; ++d) {
      coordv[coordIndex][d] = 0.25 * (coordv[face_nodes[k][0]][d] + coordv[face_nodes[k][1]][d] +
                                      coordv[face_nodes[k][2]][d] + coordv[face_nodes[k][3]][d]);
    }
  }
  static const int triangular_facets[24][3] = {
      {0, 8, 1},  {8, 2, 1},  {3, 2, 8},  {3, 8, 0},  {6, 9, 5},  {7, 9, 6},  {4, 9, 7},  {4, 5, 9},
      {10, 0, 1}, {5, 10, 1}, {4, 10, 5}, {4, 0, 10}, {7, 6, 11}, {6, 2, 11}, {2, 3, 11}, {3, 7, 11},
      {6, 12, 2}, {5, 12, 6}, {5, 1, 12}, {1, 2, 12}, {0, 4, 13}, {4, 7, 13}, {7, 3, 13}, {3, 0, 13}};

  double volume = 0.0;
  for (int k = 0; k < 24; ++k) {
    const int p = triangular_facets[k][0];
    const int q = triangular_facets[k][1];
    const int r = triangular_facets[k][2];

    const double triFaceMid[3] = {coordv[p][0] + coordv[q][0] + coordv[r][0],
                                  coordv[p][1] + coordv[q][1] + coordv[r][1],
                                  coordv[p][2] + coordv[q][2] + coordv[r][2]};
    enum { XC = 0, YC = 1, ZC = 2 };
    double dxv[3];
    dxv[0] = (coordv[q][YC] - coordv[p][YC]) * (coordv[r][ZC] - coordv[p][ZC]) -
             (coordv[r][YC] - coordv[p][YC]) * (coordv[q][ZC] - coordv[p][ZC]);
    dxv[1] = (coordv[r][XC] - coordv[p][XC]) * (coordv[q][ZC] - coordv[p][ZC]) -
             (coordv[q][XC] - coordv[p][XC]) * (coordv[r][ZC] - coordv[p][ZC]);
    dxv[2] = (coordv[q][XC] - coordv[p][XC]) * (coordv[r][YC] - coordv[p][YC]) -
             (coordv[r][XC] - coordv[p][XC]) * (coordv[q][YC] - coordv[p][YC]);

    volume += triFaceMid[0] * dxv[0] + triFaceMid[1] * dxv[1] + triFaceMid[2] * dxv[2];
  }
  volume /= 18.0;
  return volume;
}

/* src/master_element/Hex8CVFEM.C:149-178 (HexSCV::determinant, SharedMemView overload) */
void orc_hex8_scv_volume(const double *coords, double *volume)
{
  static const int subDivisionTable[8][8] = {
      {0, 8, 12, 11, 19, 20, 26, 25},  {8, 1, 9, 12, 20, 18, 24, 26},
      {12, 9, 2, 10, 26, 24, 22, 23},  {11, 12, 10, 3, 25, 26, 23, 21},
      {19, 20, 26, 25, 4, 13, 17, 16}, {20, 18, 24, 26, 13, 5, 14, 17},
      {26, 24, 22, 23, 17, 14, 6, 15}, {25, 26, 23, 21, 16, 17, 15, 7}};

  double coordv[27][3];
  orc_subdivide_hex_8(coords, coordv);

  for (int ip = 0; ip < 8; ++ip) {
    double scvHex[8][3];
    for (int n = 0; n < 8; ++n) {
      const int subIndex = subDivisionTable[ip][n];
      for (int d = 0; d < 3; ++d) scvHex[n][d] = coordv[subIndex][d];
    }
    volume[ip] = orc_hex_volume_grandy((const double(*)[3])scvHex);
  }
}

/* src/master_element/Hex8CVFEM.C:523-542 */
int orc_hex8_scs_grad_op(int shifted, const double *coords, double *gradop, double *deriv)
{
  double loc[36];
  double dloc[12 * 8 * 3];
  if (!deriv) deriv = dloc;
  orc_hex8_scs_intg_loc(shifted, loc);
  orc_hex8_derivative(12, loc, deriv);
  return orc_generic_grad_op(12, deriv, coords, gradop, NULL);
}

/* src/master_element/Hex8CVFEM.C:182-190,221-229 */
int orc_hex8_scv_grad_op(int shifted, const double *coords, double *gradop, double *deriv)
{
  double loc[24];
  double dloc[8 * 8 * 3];
  if (!deriv) deriv = dloc;
  orc_hex8_scv_intg_loc(shifted, loc);
  orc_hex8_derivative(8, loc, deriv);
  return orc_generic_grad_op(8, deriv, coords, gradop, NULL);
}

/* src/master_element/Hex8CVFEM.C:800-808 (HexSCS::gij, SharedMemView overload: always the unshifted integration
 * points) + include/master_element/MasterElementFunctions.h:171-236 (generic_gij_3d): contravariant metric
 * g^ij = J J^T with J[i][k] = sum_n dN_n/dxi_k x_n,i; covariant metric = its inverse (symmetric cofactor form) */
void orc_hex8_scs_gij(const double *coords, double *gupper, double *glower)
{
  double loc[36];
  double deriv[12 * 8 * 3];
  orc_hex8_scs_intg_loc(0, loc);
  orc_hex8_derivative(12, loc, deriv);
#define GUP(ip, i, j) gupper[((ip)*3 + (i)) * 3 + (j)]
#define GLO(ip, i, j) glower[((ip)*3 + (i)) * 3 + (j)]
  for (int ip = 0; ip < 12; ++ip) {
    double jac[3][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
    for (int n = 0; n < 8; ++n) {
      const double *d = &deriv[(ip * 8 + n) * 3];
      jac[0][0] += d[0] * coords[n * 3 + 0];
      jac[0][1] += d[1] * coords[n * 3 + 0];
      jac[0][2] += d[2] * coords[n * 3 + 0];

      jac[1][0] += d[0] * coords[n * 3 + 1];
      jac[1][1] += d[1] * coords[n * 3 + 1];
      jac[1][2] += d[2] * coords[n * 3 + 1];

      jac[2][0] += d[0] * coords[n * 3 + 2];
      jac[2][1] += d[1] * coords[n * 3 + 2];
      jac[2][2] += d[2] * coords[n * 3 + 2];
    }
    GUP(ip, 0, 0) = jac[0][0] * jac[0][0] + jac[0][1] * jac[0][1] + jac[0][2] * jac[0][2];
    GUP(ip, 0, 1) = jac[0][0] * jac[1][0] + jac[0][1] * jac[1][1] + jac[0][2] * jac[1][2];
    GUP(ip, 0, 2) = jac[0][0] * jac[2][0] + jac[0][1] * jac[2][1] + jac[0][2] * jac[2][2];

    GUP(ip, 1, 0) = GUP(ip, 0, 1);
    GUP(ip, 1, 1) = jac[1][0] * jac[1][0] + jac[1][1] * jac[1][1] + jac[1][2] * jac[1][2];
    GUP(ip, 1, 2) = jac[1][0] * jac[2][0] + jac[1][1] * jac[2][1] + jac[1][2] * jac[2][2];

    GUP(ip, 2, 0) = GUP(ip, 0, 2);
    GUP(ip, 2, 1) = GUP(ip, 1, 2);
    GUP(ip, 2, 2) = jac[2][0] * jac[2][0] + jac[2][1] * jac[2][1] + jac[2][2] * jac[2][2];

    const double inv_detj = 1.0 / (GUP(ip, 0, 0) * (GUP(ip, 1, 1) * GUP(ip, 2, 2) - GUP(ip, 1, 2) * GUP(ip, 1, 2)) -
                                   GUP(ip, 0, 1) * (GUP(ip, 0, 1) * GUP(ip, 2, 2) - GUP(ip, 1, 2) * GUP(ip, 0, 2)) +
                                   GUP(ip, 0, 2) * (GUP(ip, 0, 1) * GUP(ip, 1, 2) - GUP(ip, 1, 1) * GUP(ip, 0, 2)));

    GLO(ip, 0, 0) = inv_detj * (GUP(ip, 1, 1) * GUP(ip, 2, 2) - GUP(ip, 1, 2) * GUP(ip, 1, 2));
    GLO(ip, 0, 1) = inv_detj * (GUP(ip, 0, 2) * GUP(ip, 1, 2) - GUP(ip, 0, 1) * GUP(ip, 2, 2));
    GLO(ip, 0, 2) = inv_detj * (GUP(ip, 0, 1) * GUP(ip, 1, 2) - GUP(ip, 0, 2) * GUP(ip, 1, 1));

    GLO(ip, 1, 0) = GLO(ip, 0, 1);
    GLO(ip, 1, 1) = inv_detj * (GUP(ip, 0, 0) * GUP(ip, 2, 2) - GUP(ip, 0, 2) * GUP(ip, 0, 2));
    GLO(ip, 1, 2) = inv_detj * (GUP(ip, 0, 2) * GUP(ip, 0, 1) - GUP(ip, 0, 0) * GUP(ip, 1, 2));

    GLO(ip, 2, 0) = GLO(ip, 0, 2);
    GLO(ip, 2, 1) = GLO(ip, 1, 2);
    GLO(ip, 2, 2) = inv_detj * (GUP(ip, 0, 0) * GUP(ip, 1, 1) - GUP(ip, 0, 1) * GUP(ip, 0, 1));
  }
#undef GUP
#undef GLO
}

/* unit_tests/UnitTestUtils.C:41-54 (Lcg inside perturb_coord_hex_8) */
double orc_lcg_next(uint32_t *state)
{
  const uint32_t a = 1103515245u, c = 12345u, m = 2147483647u;
  *state = (uint32_t)(((uint64_t)a * (uint64_t)(*state) + c) % m);
  return (2. * (*state / (double)m) - 1.);
}
