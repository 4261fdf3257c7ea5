// box_tile.cuh -- tiled, write-once MomentumAdvDiff assembly for structured Hex8 blocks.
//
// What is computed is what AssembleElemSolverAlgorithm::execute + MomentumAdvDiffElemKernel::execute
// + TpetraLinearSystem::sumInto produce (src/AssembleElemSolverAlgorithm.C:74-100,
// src/kernel/MomentumAdvDiffElemKernel.C:70-186, src/TpetraLinearSystem.C:1457-1522); how it is
// computed is turned inside out so that every CSR value is written exactly once, already summed:
//
//   * a CTA owns a TBX x TBY column of NODES and sweeps it along z; each node row is owned by
//     8 threads ("classes" (a,b,c) in {0,1}^3: the thread owns the column nodes whose offset from
//     the row node is (+-a, +-b, +-c), so every one of the 8 elements around the node hands each
//     class exactly one column node -> all 8 threads do identical work, 24 (element, ip, column)
//     triples per node plane);
//   * phase 1 evaluates, once per element layer, the per-integration-point data the rows need:
//     G' = -mu_ip adj(J)/det(J)   (so that -mu dN_ic/dx = G' dN_ic/dxi), the SCS area vector A,
//     the mass flow rate and the div(u) stress scalar -- 14 doubles per ip in shared memory.
//     The trilinear map is expanded in its 8 coefficient vectors, which makes the Jacobian at an
//     SCS ip 5 vector FMAs and the area vector one cross product (the reference subdivides the hex
//     into 27 points and recomputes a derivative table per element, Hex8CVFEM.C:24-70,547-581);
//   * phase 2 accumulates the 3x3 blocks  sum_{e,ip} +-( (G' dN_ic) (x) A + (N_ic mdot + (G' dN_ic).A) I )
//     in registers (the L and R rows of an ip differ only in sign, MomentumAdvDiffElemKernel.C:152-183),
//     lays the finished row out in shared memory exactly as the CSR stores it and streams it to HBM
//     with full-width coalesced stores;
//   * rhs = -(assembled row) . u - divU stress: algebraically what the reference accumulates entry by
//     entry (rhs -= lhsfac*u, :171-182) because the advection lhs and aflux share one shape function.
//
// Every phase is a plain function of the thread index so that tests/cpu_harness can execute the
// kernel thread by thread on the CPU (no GPU in the development container).
#pragma once
#include "../../include/nalu_b200.h"
#include "hex8_dev.cuh"

namespace nalu_b200 {
namespace box {

struct BoxParams {
  int nx, ny, nz;    // elements of the local block
  int NX, NY, NZ;    // nodes
  int z_chunk;       // node planes per CTA
  int tiles_x, tiles_y;
  const double *coords; // [nNodes][3]
  const double *vel;    // [nNodes][3]
  const double *visc;   // [nNodes]
  const double *mdot;   // [nElem][12]
  const int *node_row;
  const long long *browptr;
  const unsigned char *slot27;
  double *values;
  double *rhs;
  int *err;
  double include_divu;
  int options;
  int overwrite;
  int z_first, z_last; // node planes [z_first, z_last) this launch assembles (column-pair kernel; default all)
  long long *trace; // profiling aid (NALU_BOX_TRACE): per-warp clock stamps of one CTA, else null
};

template <int TBX_, int TBY_>
struct Cfg {
  static constexpr int TBX = TBX_, TBY = TBY_;
  static constexpr int NNODE = TBX * TBY;
  static constexpr int NT = NNODE * 8;
  static constexpr int ELX = TBX + 1, ELY = TBY + 1, NEL = ELX * ELY;
  static constexpr int IPD = 14;            // doubles per ip: G'[9], A[3], mdot, w
  static constexpr int ES = 12 * IPD + 2;   // element stride (doubles); +2 spreads the 16-byte bank groups
  static constexpr int SROW = 248;          // staging row stride (doubles) >= 27*9
  static constexpr int LBUF = (NEL * ES > NNODE * SROW) ? NEL * ES : NNODE * SROW;
  static constexpr int TAB = 8 * 3 * 2 * 8 * 2; // [octant][edge dir][half][class][2]
  static constexpr int RNX = TBX + 2, RNY = TBY + 2, RN = RNX * RNY; // node neighbourhood of the tile
  static constexpr int NF = 7;              // staged nodal doubles: x[3], mu, u[3]
  static constexpr int RING = 3 * RN * NF;  // three node planes
  static constexpr int MDB = NEL * 12;      // mass flow rates of one element layer
  static constexpr int NPL = (RN * NF + NT - 1) / NT;  // per-thread loads to stage one node plane
  static constexpr int NML = (MDB + NT - 1) / NT;      // per-thread loads to stage one mdot layer
  // doubles: two layer buffers, table, nodal ring, mdot layer, row base (as 8-byte ints) and row length
  static constexpr int OFF_TAB = 2 * LBUF, OFF_RING = OFF_TAB + TAB, OFF_MD = OFF_RING + RING,
                       OFF_RB = OFF_MD + MDB, OFF_LEN = OFF_RB + NNODE;
  static constexpr int SMEM_DOUBLES = OFF_LEN + (NNODE + 1) / 2;
  // phase-1 work items: (element, edge direction), split once more in two when that still fits one
  // round of the CTA's threads
  static constexpr bool SPLIT = NEL * 6 <= NT;
  static constexpr int NITEM = NEL * 3 * (SPLIT ? 2 : 1);
};

// ---- small integer helpers on the Hex8 numbering (Hex8CVFEM.C:264-276, 396-412) -------------
__host__ __device__ __forceinline__ int node_of_bits(int bx, int by, int bz) { return bz * 4 + by * 2 + (bx ^ by); }
__host__ __device__ __forceinline__ int bit_x(int n) { return (n ^ (n >> 1)) & 1; }
__host__ __device__ __forceinline__ int bit_y(int n) { return (n >> 1) & 1; }
__host__ __device__ __forceinline__ int bit_z(int n) { return (n >> 2) & 1; }
// the SCS ip on the edge that leaves node k in direction dir (0 x, 1 y, 2 z)
__host__ __device__ __forceinline__ int ip_of(int k, int dir)
{
  const int bz = bit_z(k);
  if (dir == 0) return bit_y(k) ? (bz ? 6 : 2) : (bz ? 4 : 0);
  if (dir == 1) return bit_x(k) ? (bz ? 5 : 1) : (bz ? 7 : 3);
  return 8 + (k & 3);
}

// ---- phase 2 constant table -------------------------------------------------------------------
// For octant s (bits: 1 = +x, 2 = +y, 4 = +z: the element lies on that side of the row node),
// edge direction pp and class cls: parametric derivative of the class's column node at the ip, in
// the ip's local axis order (t,u,v) = (pp, pp+1, pp+2), and the advection shape function value.
template <class C>
__host__ __device__ inline void build_table_entry(int idx, int options, double *tab)
{
  const int cls = idx & 7, pp = (idx >> 3) % 3, s = idx / 24;
  const int sxp = s & 1, syp = (s >> 1) & 1, szp = (s >> 2) & 1;
  // row node sits at the corner of the element opposite to the octant direction
  const int k = node_of_bits(!sxp, !syp, !szp);
  const int a = cls & 1, b = (cls >> 1) & 1, c = (cls >> 2) & 1;
  const int ic = node_of_bits(a ? sxp : !sxp, b ? syp : !syp, c ? szp : !szp);
  const int ip = ip_of(k, pp);
  double s1, s2, s3, d[3];
  hex8::scs_intg_loc(ip, (options & NALU_OPT_SHIFTED_GRAD_OP) ? 0.5 : 0.25, s1, s2, s3);
  hex8::shape_deriv(ic, s1, s2, s3, d);
  hex8::scs_intg_loc(ip, (options & NALU_OPT_SKEW_SYMMETRIC) ? 0.5 : 0.25, s1, s2, s3);
  const double nadv = hex8::shape_fcn(ic, s1, s2, s3);
  double *t0 = tab + ((s * 3 + pp) * 2 + 0) * 16 + cls * 2;
  double *t1 = tab + ((s * 3 + pp) * 2 + 1) * 16 + cls * 2;
  t0[0] = d[pp];
  t0[1] = d[(pp + 1) % 3];
  t1[0] = d[(pp + 2) % 3];
  t1[1] = nadv;
}

// ---- staging of nodal fields / mass flow rates (global -> registers -> shared) ---------------------
// Slot k of thread tid covers flat index tid + k*NT of the [RN][NF] plane image; the loads are
// issued one plane ahead of their use and parked in registers so their latency is off the path.
template <class C>
__host__ __device__ inline void plane_load(const BoxParams &P, int tid, int x0, int y0, int zp, double (&reg)[C::NPL])
{
#pragma unroll
  for (int k = 0; k < C::NPL; ++k) {
    const int idx = tid + k * C::NT;
    reg[k] = 0.0;
    if (idx >= C::RN * C::NF || zp < 0 || zp >= P.NZ) continue;
    const int rn = idx / C::NF, f = idx - rn * C::NF;
    const int x = x0 - 1 + rn % C::RNX, y = y0 - 1 + rn / C::RNX;
    if (x < 0 || y < 0 || x >= P.NX || y >= P.NY) continue;
    const long long n = x + (long long)P.NX * (y + (long long)P.NY * zp);
    reg[k] = f < 3 ? P.coords[n * 3 + f] : (f == 3 ? P.visc[n] : P.vel[n * 3 + (f - 4)]);
  }
}
template <class C>
__host__ __device__ inline void plane_store(int tid, const double (&reg)[C::NPL], double *ring_plane)
{
#pragma unroll
  for (int k = 0; k < C::NPL; ++k) {
    const int idx = tid + k * C::NT;
    if (idx < C::RN * C::NF) ring_plane[idx] = reg[k];
  }
}
template <class C>
__host__ __device__ inline void mdot_load(const BoxParams &P, int tid, int x0, int y0, int ez, double (&reg)[C::NML])
{
#pragma unroll
  for (int k = 0; k < C::NML; ++k) {
    const int idx = tid + k * C::NT;
    reg[k] = 0.0;
    if (idx >= C::MDB || ez < 0 || ez >= P.nz) continue;
    const int el = idx / 12, ip = idx - el * 12;
    const int ex = x0 - 1 + el % C::ELX, ey = y0 - 1 + el / C::ELX;
    if (ex < 0 || ey < 0 || ex >= P.nx || ey >= P.ny) continue;
    reg[k] = P.mdot[(ex + (long long)P.nx * (ey + (long long)P.ny * ez)) * 12 + ip];
  }
}
template <class C>
__host__ __device__ inline void mdot_store(int tid, const double (&reg)[C::NML], double *mdb)
{
#pragma unroll
  for (int k = 0; k < C::NML; ++k) {
    const int idx = tid + k * C::NT;
    if (idx < C::MDB) mdb[idx] = reg[k];
  }
}

// ---- phase 1: integration-point data of one (element, edge direction[, half]) item -------------
// Local axes (t,u,v) = (g, g+1, g+2) mod 3; local node m has bits (b0,b1,b2) along (t,u,v).
// pl0 / pl1: staged node planes below / above the element layer; mdb: the layer's mass flow rates.
// An item covers the 4 ips of one edge direction, or (C::SPLIT) the 2 of them on one v side so that
// every warp of a large tile has phase-1 work.
__host__ __device__ __forceinline__ void butterfly8(double *f)
{
#pragma unroll
  for (int h = 1; h < 8; h <<= 1)
#pragma unroll
    for (int m = 0; m < 8; ++m)
      if (!(m & h)) {
        const double lo = f[m], hi = f[m | h];
        f[m] = hi + lo;
        f[m | h] = hi - lo;
      }
}

template <class C>
__host__ __device__ inline void phase1_item(const BoxParams &P, int item, int x0, int y0, const double *pl0,
                                            const double *pl1, const double *mdb, double *buf)
{
  constexpr int NH = C::SPLIT ? 2 : 1, NIP = 4 / NH;
  const int half = C::SPLIT ? (item & 1) : 0;
  const int ge = item / NH;
  const int g = ge / C::NEL, el = ge - g * C::NEL;
  const int elx = el % C::ELX, ely = el / C::ELX;
  const int ex = x0 - 1 + elx, ey = y0 - 1 + ely;
  if (ex < 0 || ey < 0 || ex >= P.nx || ey >= P.ny) return;
  // strides of the local axes (t,u,v) in the staged image: x -> NF, y -> RNX*NF, z -> other plane
  const int sx = C::NF, sy = C::RNX * C::NF;
  const int s_t = g == 0 ? sx : (g == 1 ? sy : 0), s_u = g == 0 ? sy : (g == 1 ? 0 : sx),
            s_v = g == 0 ? 0 : (g == 1 ? sx : sy);
  const int zbit = g == 0 ? 4 : (g == 1 ? 2 : 1); // which local bit is the z axis
  const int n000 = (elx + C::RNX * ely) * C::NF;
  const bool divu = P.include_divu != 0.0;
  enum { T = 1, Uu = 2, V = 4, TU = 3, TV = 5, UV = 6, TUV = 7 };
  const double q = (P.options & NALU_OPT_SHIFTED_GRAD_OP) ? 0.5 : 0.25;
  // original-axis bit positions of the local axes
  const int at = g, au = (g + 1) % 3, av = (g + 2) % 3;
  double *const oe = buf + el * C::ES;
  int ipn[NIP];

  { // ---- geometry: G' = -mu adj(J)/det, area vector
    // butterfly: C[mask] = sum_m prod_{axis in mask} sign_axis(m) f_m   (mask bits: 1 = t, 2 = u, 4 = v)
    double X[3][8], M[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const double *nd = ((m & zbit) ? pl1 : pl0) + n000 + (m & 1) * s_t + ((m >> 1) & 1) * s_u + ((m >> 2) & 1) * s_v;
      X[0][m] = nd[0];
      X[1][m] = nd[1];
      X[2][m] = nd[2];
      M[m] = nd[3];
    }
    butterfly8(X[0]);
    butterfly8(X[1]);
    butterfly8(X[2]);
    butterfly8(M);
    // area vector of the facet through the t = 0 plane, pointing along +t:
    //   A = Cu x Cv / 64 + Cuv x (sig2 Cv - sig1 Cu) / 128
    double base[3];
    base[0] = (X[1][Uu] * X[2][V] - X[2][Uu] * X[1][V]) * (1.0 / 64.0);
    base[1] = (X[2][Uu] * X[0][V] - X[0][Uu] * X[2][V]) * (1.0 / 64.0);
    base[2] = (X[0][Uu] * X[1][V] - X[1][Uu] * X[0][V]) * (1.0 / 64.0);
#pragma unroll
    for (int ff = 0; ff < NIP; ++ff) {
      const int b1 = ff & 1, b2 = C::SPLIT ? half : (ff >> 1);
      const double sg1 = b1 ? 1.0 : -1.0, sg2 = b2 ? 1.0 : -1.0;
      // which ip: the edge along t at (u,v) side (b1,b2); take its node with bit_t = 0
      const int obx = at == 0 ? 0 : (au == 0 ? b1 : b2), oby = at == 1 ? 0 : (au == 1 ? b1 : b2),
                obz = at == 2 ? 0 : (au == 2 ? b1 : b2);
      (void)av;
      const int k0 = node_of_bits(obx, oby, obz);
      const int ip = ip_of(k0, g);
      ipn[ff] = ip;
      const double orient = (hex8::scs_left(ip) == k0) ? 1.0 : -1.0;

      double J[3][3]; // J[i][j] = d x_i / d xi_j, local axis order
      const double c1 = 0.5 * q * sg1, c2 = 0.5 * q * sg2, c12 = q * q * sg1 * sg2;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        J[i][0] = 0.25 * X[i][T] + c1 * X[i][TU] + c2 * X[i][TV] + c12 * X[i][TUV];
        J[i][1] = 0.25 * X[i][Uu] + c2 * X[i][UV];
        J[i][2] = 0.25 * X[i][V] + c1 * X[i][UV];
      }
      double adj[3][3];
      adj[0][0] = J[1][1] * J[2][2] - J[2][1] * J[1][2];
      adj[0][1] = J[1][2] * J[2][0] - J[2][2] * J[1][0];
      adj[0][2] = J[1][0] * J[2][1] - J[2][0] * J[1][1];
      adj[1][0] = J[0][2] * J[2][1] - J[2][2] * J[0][1];
      adj[1][1] = J[0][0] * J[2][2] - J[2][0] * J[0][2];
      adj[1][2] = J[0][1] * J[2][0] - J[2][1] * J[0][0];
      adj[2][0] = J[0][1] * J[1][2] - J[1][1] * J[0][2];
      adj[2][1] = J[0][2] * J[1][0] - J[1][2] * J[0][0];
      adj[2][2] = J[0][0] * J[1][1] - J[1][0] * J[0][1];
      const double det = J[0][0] * adj[0][0] + J[1][0] * adj[1][0] + J[2][0] * adj[2][0];
      if (!(det > 0.0)) *P.err = 1; // MasterElementFunctions.h:91-94 asserts this in debug builds
      const double inv = 1.0 / det;
      // mu at the (never shifted) ip location, MomentumAdvDiffElemKernel.C:107-110
      const double mu = 0.125 * M[0] + 0.0625 * (sg1 * M[Uu] + sg2 * M[V]) + 0.03125 * sg1 * sg2 * M[UV];
      const double fac = -mu * inv;

      double *o = oe + ip * C::IPD;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) o[i * 3 + j] = fac * adj[i][j];
      double W[3], A[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) W[i] = sg2 * X[i][V] - sg1 * X[i][Uu];
      A[0] = base[0] + (X[1][UV] * W[2] - X[2][UV] * W[1]) * (1.0 / 128.0);
      A[1] = base[1] + (X[2][UV] * W[0] - X[0][UV] * W[2]) * (1.0 / 128.0);
      A[2] = base[2] + (X[0][UV] * W[1] - X[1][UV] * W[0]) * (1.0 / 128.0);
      o[9] = orient * A[0];
      o[10] = orient * A[1];
      o[11] = orient * A[2];
      o[12] = mdb[el * 12 + ip];
      if (!divu) o[13] = 0.0;
    }
  }
  if (divu) {
    // ---- divU stress scalar w = 2/3 mu divU includeDivU (:111-114,:127), with
    //   mu divU = mu sum_ij (adj_ij/det) d u_i/d xi_j = - sum_ij G'_ij d u_i/d xi_j   (G' read back)
    double U[3][8];
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const double *nd = ((m & zbit) ? pl1 : pl0) + n000 + (m & 1) * s_t + ((m >> 1) & 1) * s_u + ((m >> 2) & 1) * s_v;
      U[0][m] = nd[4];
      U[1][m] = nd[5];
      U[2][m] = nd[6];
    }
    butterfly8(U[0]);
    butterfly8(U[1]);
    butterfly8(U[2]);
    const double wf = -2.0 / 3.0 * P.include_divu;
#pragma unroll
    for (int ff = 0; ff < NIP; ++ff) {
      const int b1 = ff & 1, b2 = C::SPLIT ? half : (ff >> 1);
      const double sg1 = b1 ? 1.0 : -1.0, sg2 = b2 ? 1.0 : -1.0;
      const double c1 = 0.5 * q * sg1, c2 = 0.5 * q * sg2, c12 = q * q * sg1 * sg2;
      double *o = oe + ipn[ff] * C::IPD;
      double dv = 0.0;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const double d0 = 0.25 * U[i][T] + c1 * U[i][TU] + c2 * U[i][TV] + c12 * U[i][TUV];
        const double d1 = 0.25 * U[i][Uu] + c2 * U[i][UV];
        const double d2 = 0.25 * U[i][V] + c1 * U[i][UV];
        dv += o[i * 3 + 0] * d0 + o[i * 3 + 1] * d1 + o[i * 3 + 2] * d2;
      }
      o[13] = wf * dv;
    }
  }
}

// ---- phase 2 per-thread state ---------------------------------------------------------------------
struct P2State {
  double acc[4][9]; // 3x3 blocks per (sx,sy) quadrant, row-major [i][j]
  double D[4];      // diagonal part of each block
  double rr[3];     // rhs partial of this thread
  double dv;        // divU stress component carried by classes 0..2
};

struct NodeCtx {
  int active;        // node inside the domain
  int x, y, z;
  int lx, ly;        // position in the tile
  long long n;       // lattice node index
  long long rowbase; // values offset of the node's first dof row
  int nb;            // blocks in the row
  int row;           // row node index (rhs offset / 3)
  int cls;
  unsigned slots_mid, slots_fin; // CSR block slot (0xFF: no such neighbour) of the 4 quadrant blocks
  int vmask;         // bit q: the (sx,sy) quadrant q holds an element column
  int eoff;          // offset of element (lx,ly) of the layer tile in a layer buffer
};

template <class C>
__host__ __device__ inline void node_ctx_init(const BoxParams &P, NodeCtx &N, int tid, int x0, int y0)
{
  const int nl = tid >> 3;
  N.cls = tid & 7;
  N.lx = nl % C::TBX;
  N.ly = nl / C::TBX;
  N.x = x0 + N.lx;
  N.y = y0 + N.ly;
  N.active = N.x < P.NX && N.y < P.NY;
  N.vmask = 0;
  for (int q = 0; q < 4; ++q) {
    const int ex = N.x - 1 + (q & 1), ey = N.y - 1 + (q >> 1);
    if (ex >= 0 && ey >= 0 && ex < P.nx && ey < P.ny) N.vmask |= 1 << q;
  }
  N.eoff = (N.lx + C::ELX * N.ly) * C::ES;
}

// CSR slots of the blocks this class flushes: byte q = quadrant (sx,sy) -> offset (a sx, b sy, dz)
__host__ __device__ inline unsigned load_slots(const unsigned char *slots, int cls, int dz)
{
  const int a = cls & 1, b = (cls >> 1) & 1;
  unsigned r = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int sxp = q & 1, syp = q >> 1;
    unsigned v = 0xFF;
    if (!((!a && sxp) || (!b && syp))) {
      const int dx = a ? (sxp ? 1 : -1) : 0, dy = b ? (syp ? 1 : -1) : 0;
      v = slots[(dz + 1) * 9 + (dy + 1) * 3 + (dx + 1)];
    }
    r |= v << (8 * q);
  }
  return r;
}

__host__ __device__ inline void p2_zero(P2State &S)
{
#pragma unroll
  for (int q = 0; q < 4; ++q) {
#pragma unroll
    for (int t = 0; t < 9; ++t) S.acc[q][t] = 0.0;
    S.D[q] = 0.0;
  }
}
__host__ __device__ inline void p2_begin(P2State &S)
{
  p2_zero(S);
  S.rr[0] = S.rr[1] = S.rr[2] = 0.0;
  S.dv = 0.0;
}

// one z-half (SZP = 0: elements below the node plane, 1: above) of the 8 octants.
// tabc = table + 2*cls; buf = that layer's ip data
template <class C, int SZP>
__host__ __device__ inline void p2_pass(const BoxParams &P, const NodeCtx &N, P2State &S, const double *buf,
                                        const double *tabc, bool divu)
{
  const int ez = N.z - 1 + SZP;
  if (ez < 0 || ez >= P.nz) return;
  const double *ebase = buf + N.eoff;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int sxp = q & 1, syp = q >> 1;
    if (!(N.vmask & (1 << q))) continue; // no element in this quadrant (domain boundary)
    const int s = q + 4 * SZP;
    const int k = node_of_bits(!sxp, !syp, !SZP);
    const double *eb = ebase + (sxp + C::ELX * syp) * C::ES;
#pragma unroll
    for (int pp = 0; pp < 3; ++pp) {
      const int ip = ip_of(k, pp);
      const bool isL = hex8::scs_left(ip) == k;
      const double *o = eb + ip * C::IPD;
      const double *t0 = tabc + ((s * 3 + pp) * 2 + 0) * 16;
      const double *t1 = tabc + ((s * 3 + pp) * 2 + 1) * 16;
      const double d0 = t0[0], d1 = t0[1], d2 = t1[0], nadv = t1[1];
      double h[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) h[i] = o[i * 3 + 0] * d0 + o[i * 3 + 1] * d1 + o[i * 3 + 2] * d2;
      const double A0 = o[9], A1 = o[10], A2 = o[11];
      const double dsum = h[0] * A0 + h[1] * A1 + h[2] * A2 + nadv * o[12];
      if (isL) {
        S.D[q] += dsum;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          S.acc[q][i * 3 + 0] += h[i] * A0;
          S.acc[q][i * 3 + 1] += h[i] * A1;
          S.acc[q][i * 3 + 2] += h[i] * A2;
        }
      } else {
        S.D[q] -= dsum;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          S.acc[q][i * 3 + 0] -= h[i] * A0;
          S.acc[q][i * 3 + 1] -= h[i] * A1;
          S.acc[q][i * 3 + 2] -= h[i] * A2;
        }
      }
      if (divu) {
        // rhs(L) -= divUstress, rhs(R) += divUstress (:127-135).  Every class meets each (element, ip)
        // next to the node exactly once; classes 0,1,2 each carry one component of w A (the other
        // classes' copies are dropped when the partials are combined)
        const double Asel = o[9 + (N.cls < 3 ? N.cls : 0)];
        if (isL) S.dv -= o[13] * Asel; else S.dv += o[13] * Asel;
      }
    }
  }
}

// fold the quadrant accumulators of the axes the class does not distinguish, then write the
// finished blocks into the staging row (CSR layout) and take their share of rhs = -row . u.
// `slots`: packed CSR slots of the 4 quadrant blocks; `uplane`: staged node plane of the column nodes
template <class C>
__host__ __device__ inline void p2_flush(const NodeCtx &N, P2State &S, unsigned slots, const double *uplane, double *stg)
{
  const int a = N.cls & 1, b = (N.cls >> 1) & 1;
  if (!a) {
#pragma unroll
    for (int sy = 0; sy < 2; ++sy) {
#pragma unroll
      for (int t = 0; t < 9; ++t) S.acc[sy * 2][t] += S.acc[sy * 2 + 1][t];
      S.D[sy * 2] += S.D[sy * 2 + 1];
    }
  }
  if (!b) {
#pragma unroll
    for (int sx = 0; sx < 2; ++sx) {
#pragma unroll
      for (int t = 0; t < 9; ++t) S.acc[sx][t] += S.acc[2 + sx][t];
      S.D[sx] += S.D[2 + sx];
    }
  }
  double *row = stg + (N.lx + C::TBX * N.ly) * C::SROW;
  const int rs = N.nb * 3; // doubles per dof row
  const double *uc = uplane + ((N.lx + 1) + C::RNX * (N.ly + 1)) * C::NF + 4;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int sxp = q & 1, syp = q >> 1;
    const unsigned slot = (slots >> (8 * q)) & 0xFF;
    if (slot == 0xFF) continue; // block not owned by this class, or no such neighbour node
    const int dx = a ? (sxp ? 1 : -1) : 0, dy = b ? (syp ? 1 : -1) : 0;
    const double *u = uc + (dx + C::RNX * dy) * C::NF;
    const double u0 = u[0], u1 = u[1], u2 = u[2];
    double *bp = row + slot * 3;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double v0 = S.acc[q][i * 3 + 0], v1 = S.acc[q][i * 3 + 1], v2 = S.acc[q][i * 3 + 2];
      if (i == 0) v0 += S.D[q];
      if (i == 1) v1 += S.D[q];
      if (i == 2) v2 += S.D[q];
      bp[i * rs + 0] = v0;
      bp[i * rs + 1] = v1;
      bp[i * rs + 2] = v2;
      S.rr[i] -= v0 * u0 + v1 * u1 + v2 * u2;
    }
  }
}

} // namespace box
} // namespace nalu_b200
