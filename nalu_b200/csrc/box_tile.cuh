// box_tile.cuh -- tiled, write-once MomentumAdvDiff assembly for structured Hex8 blocks.
//
// What is computed is what AssembleElemSolverAlgorithm::execute + MomentumAdvDiffElemKernel::execute
// + TpetraLinearSystem::sumInto produce (src/AssembleElemSolverAlgorithm.C:74-100,
// src/kernel/MomentumAdvDiffElemKernel.C:70-186, src/TpetraLinearSystem.C:1457-1522); how it is
// computed is turned inside out so that every CSR value is written exactly once, already summed:
//
//   * a CTA owns a TBX x TBY column of NODES and sweeps it along z; each node row is owned by
//     8 threads.  Column nodes fall into 8 "classes" (a,b,c) in {0,1}^3 -- the class holds the column
//     nodes whose offset from the row node is (+-a, +-b, +-c), so every one of the 8 elements around
//     the node hands each class exactly one column node.  A thread owns the class pair (a,b,0),(a,b,1)
//     on one half H of the 4 (x,y) quadrants (split along x if a, else along y): per loaded integration
//     point it serves two column nodes, and all 8 threads do identical work (24 (element, ip, column)
//     triples per node plane) with 4 block accumulators each;
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
  const unsigned char *slot27; // [nNodes][27] CSR slot of each lattice neighbour (host emulation / table build)
  const unsigned *slotpack;    // [nNodes][8][2] slot_pack(.., w = 0, 1) of every thread of the node
  double *values;
  double *rhs;
  int *err;
  double include_divu;
  int options;
  int overwrite;
};

template <int TBX_, int TBY_>
struct Cfg {
  static constexpr int TBX = TBX_, TBY = TBY_;
  static constexpr int NNODE = TBX * TBY;
  static constexpr int TPN = 8;             // threads per node row
  static constexpr int NT = NNODE * TPN;
  static constexpr int ELX = TBX + 1, ELY = TBY + 1, NEL = ELX * ELY;
  static constexpr int IPD = 14;            // doubles per ip: G'[9], A[3], mdot, w
  static constexpr int ES = 12 * IPD + 2;   // element stride (doubles); +2 spreads the 16-byte bank groups
  static constexpr int SROW = 248;          // staging row stride (doubles) >= 27*9
  static constexpr int LBUF = (NEL * ES > NNODE * SROW) ? NEL * ES : NNODE * SROW;
  static constexpr int TABE = 6;            // d0,d1,d2 (ip-local axes), N_adv, sign, pad -- all pre-multiplied by the L/R sign
  static constexpr int TAB = 8 * 3 * 8 * TABE; // [octant][edge dir][class][TABE]
  static constexpr int RNX = TBX + 2, RNY = TBY + 2, RN = RNX * RNY; // node neighbourhood of the tile
  static constexpr int NF = 7;              // staged nodal doubles: x[3], mu, u[3]
  static constexpr int RING = 3 * RN * NF;  // three node planes
  static constexpr int MDB = NEL * 12;      // mass flow rates of one element layer
  static constexpr int NPL = (RN * NF + NT - 1) / NT;  // per-thread loads to stage one node plane
  static constexpr int NML = (MDB + NT - 1) / NT;      // per-thread loads to stage one mdot layer
  // doubles: two layer buffers, table, nodal ring, mdot layer, parked half blocks, row base (8-byte ints), row length
  static constexpr int PARK = NNODE * 2 * 9; // half blocks of the two centre-column classes
  static constexpr int OFF_TAB = 2 * LBUF, OFF_RING = OFF_TAB + TAB, OFF_MD = OFF_RING + RING,
                       OFF_PARK = OFF_MD + MDB, OFF_RB = OFF_PARK + PARK, OFF_LEN = OFF_RB + NNODE;
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
  // rows L and R of an ip receive +X and -X (MomentumAdvDiffElemKernel.C:152-183): fold the sign here
  const double sg = hex8::scs_left(ip) == k ? 1.0 : -1.0;
  double s1, s2, s3, d[3];
  hex8::scs_intg_loc(ip, (options & NALU_OPT_SHIFTED_GRAD_OP) ? 0.5 : 0.25, s1, s2, s3);
  hex8::shape_deriv(ic, s1, s2, s3, d);
  hex8::scs_intg_loc(ip, (options & NALU_OPT_SKEW_SYMMETRIC) ? 0.5 : 0.25, s1, s2, s3);
  const double nadv = hex8::shape_fcn(ic, s1, s2, s3);
  double *t = tab + ((s * 3 + pp) * 8 + cls) * C::TABE;
  t[0] = sg * d[pp];
  t[1] = sg * d[(pp + 1) % 3];
  t[2] = sg * d[(pp + 2) % 3];
  t[3] = sg * nadv;
  t[4] = sg;
  t[5] = 0.0;
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
  double acc[2][2][9]; // [cc][quadrant slot] 3x3 blocks, row-major [i][j]; cc = c bit of the class
  double D[2][2];      // diagonal part of each block
  double rr[3];        // rhs partial of this thread
  double dv;           // divU stress component carried by the (a,b) != (1,1) threads
};

// Per-thread constants and the per-plane row context.  sub = tid % 8 = ab + 4 H, ab = a + 2 b.
struct NodeCtx {
  int active;        // node inside the domain
  int x, y, z;
  int lx, ly, nl;    // position in the tile
  long long n;       // lattice node index
  long long rowbase; // values offset of the node's first dof row
  int nb;            // blocks in the row
  int row;           // row node index (rhs offset / 3)
  int ab, H;
  int valid[2];      // quadrant slot j holds an element column
  int eoff[2];       // its element's offset in a layer buffer
  int toff[2];       // its table offset for class (a,b,0), lower layer (+4*TABE: c = 1; +96*TABE: upper layer)
  int ipo[2][3];     // ip*IPD of the x-, y-, z-edge ip next to the row node, lower layer (x,y: -4*IPD above)
  int fold;          // the two quadrant slots belong to one block
  int pair;          // ... which the H = 0 and H = 1 threads share (a = b = 0)
  int uoff[2];       // offset of the block's column node in a staged node plane (velocity at +4)
  unsigned slots_mid; // CSR block slots (0xFF: none): byte j = slot of block j, dz = -1 (class c = 1)
  unsigned slots_fin; // byte j: dz = 0 (class c = 0); byte 2+j: dz = +1 (class c = 1)
};

// offsets (dx,dy) of the block finished from quadrant slot j by thread (ab, H)
__host__ __device__ inline void block_offset(int ab, int H, int j, int &dx, int &dy)
{
  const int a = ab & 1, b = (ab >> 1) & 1;
  if (a) {
    dx = H ? 1 : -1;
    dy = b ? (j ? 1 : -1) : 0;
  } else {
    dx = 0;
    dy = b ? (H ? 1 : -1) : 0;
  }
}

// CSR slots of the blocks a thread finishes, from the node's 27-neighbour slot row (0xFF: no such
// neighbour / not a block of this thread): w = 0 -> slots_mid, w = 1 -> slots_fin
__host__ __device__ inline unsigned slot_pack(const unsigned char *s27, int sub, int w)
{
  const int ab = sub & 3, H = sub >> 2;
  const int a = ab & 1, b = (ab >> 1) & 1;
  const int fold = a ? !b : 1, pair = !a && !b;
  const int nblk = fold ? 1 : 2;
  unsigned r = 0;
  for (int g = 0; g < 2; ++g) { // w = 0: g = 0 only (dz = -1); w = 1: g = 0 -> dz = 0, g = 1 -> dz = +1
    const int dz = w == 0 ? -1 : g;
    for (int j = 0; j < 2; ++j) {
      unsigned v = 0xFF;
      if (j < nblk && !(pair && H == 1) && !(w == 0 && g == 1)) {
        int dx, dy;
        block_offset(ab, H, j, dx, dy);
        v = s27[(dz + 1) * 9 + (dy + 1) * 3 + (dx + 1)];
      }
      r |= v << (8 * (2 * g + j));
    }
  }
  return r;
}

template <class C>
__host__ __device__ inline void node_ctx_init(const BoxParams &P, NodeCtx &N, int tid, int x0, int y0)
{
  N.nl = tid / C::TPN;
  const int sub = tid % C::TPN;
  N.ab = sub & 3;
  N.H = sub >> 2;
  N.lx = N.nl % C::TBX;
  N.ly = N.nl / C::TBX;
  N.x = x0 + N.lx;
  N.y = y0 + N.ly;
  N.active = N.x < P.NX && N.y < P.NY;
  const int a = N.ab & 1, b = (N.ab >> 1) & 1;
  N.fold = a ? !b : 1;
  N.pair = !a && !b;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    // threads that distinguish x split their quadrants by x (H = sx, j = sy); the others by y
    const int sxp = (a || !b) ? N.H : j, syp = (a || !b) ? j : N.H;
    const int q = sxp + 2 * syp;
    const int ex = N.x - 1 + sxp, ey = N.y - 1 + syp;
    N.valid[j] = N.active && ex >= 0 && ey >= 0 && ex < P.nx && ey < P.ny;
    N.eoff[j] = ((N.lx + sxp) + C::ELX * (N.ly + syp)) * C::ES;
    N.toff[j] = (q * 3 * 8 + N.ab) * C::TABE;
    // ips next to the row node in the element BELOW the plane (row node = a top corner, bz = 1)
    const int k = node_of_bits(!sxp, !syp, 1);
    N.ipo[j][0] = ip_of(k, 0) * C::IPD;
    N.ipo[j][1] = ip_of(k, 1) * C::IPD;
    N.ipo[j][2] = ip_of(k, 2) * C::IPD;
    int dx, dy;
    block_offset(N.ab, N.H, j, dx, dy);
    N.uoff[j] = ((N.lx + 1 + dx) + C::RNX * (N.ly + 1 + dy)) * C::NF + 4;
  }
}

template <int CC>
__host__ __device__ inline void p2_zero(P2State &S)
{
#pragma unroll
  for (int j = 0; j < 2; ++j) {
#pragma unroll
    for (int t = 0; t < 9; ++t) S.acc[CC][j][t] = 0.0;
    S.D[CC][j] = 0.0;
  }
}
__host__ __device__ inline void p2_begin(P2State &S)
{
  p2_zero<0>(S);
  p2_zero<1>(S);
  S.rr[0] = S.rr[1] = S.rr[2] = 0.0;
  S.dv = 0.0;
}

// one z-half (SZP = 0: elements below the node plane, 1: above): the thread's two quadrants, 3 ips each,
// two column nodes (c = 0, 1) per ip.  The x-/y-edge ips next to the row node lie on the element's top
// face below the plane (ip 4..7) and on its bottom face above it (ip 0..3): 4*IPD doubles earlier; the
// z-edge ip is the same.
template <class C, int SZP>
__host__ __device__ inline void p2_pass(const BoxParams &P, const NodeCtx &N, P2State &S, const double *buf,
                                        const double *tab, bool divu)
{
  const int ez = N.z - 1 + SZP;
  if (ez < 0 || ez >= P.nz) return;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    if (!N.valid[j]) continue; // no element in this quadrant (domain boundary / node outside the domain)
    const double *eb = buf + N.eoff[j];
    const double *tb = tab + N.toff[j] + SZP * (4 * 3 * 8 * C::TABE);
#pragma unroll
    for (int pp = 0; pp < 3; ++pp) {
      const double *o = eb + N.ipo[j][pp] - ((SZP && pp < 2) ? 4 * C::IPD : 0);
      const double G00 = o[0], G01 = o[1], G02 = o[2], G10 = o[3], G11 = o[4], G12 = o[5], G20 = o[6], G21 = o[7],
                   G22 = o[8];
      const double A0 = o[9], A1 = o[10], A2 = o[11], md = o[12];
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const double *t = tb + pp * (8 * C::TABE) + cc * (4 * C::TABE);
        const double d0 = t[0], d1 = t[1], d2 = t[2], nadv = t[3];
        // signed: +-(-mu dN_ic/dx)
        const double h0 = G00 * d0 + G01 * d1 + G02 * d2;
        const double h1 = G10 * d0 + G11 * d1 + G12 * d2;
        const double h2 = G20 * d0 + G21 * d1 + G22 * d2;
        S.D[cc][j] += h0 * A0 + h1 * A1 + h2 * A2 + nadv * md;
        S.acc[cc][j][0] += h0 * A0; S.acc[cc][j][1] += h0 * A1; S.acc[cc][j][2] += h0 * A2;
        S.acc[cc][j][3] += h1 * A0; S.acc[cc][j][4] += h1 * A1; S.acc[cc][j][5] += h1 * A2;
        S.acc[cc][j][6] += h2 * A0; S.acc[cc][j][7] += h2 * A1; S.acc[cc][j][8] += h2 * A2;
      }
      if (divu) {
        // rhs(L) -= divUstress, rhs(R) += divUstress (:127-135).  Every (a,b) thread pair (H = 0, 1) meets
        // each (element, ip) next to the node exactly once; (a,b) = (0,0),(1,0),(0,1) each carry one
        // component of w A, the (1,1) threads' sums are dropped when the partials are combined
        S.dv -= tb[pp * (8 * C::TABE) + 4] * o[13] * o[9 + (N.ab < 3 ? N.ab : 0)];
      }
    }
  }
}

// Finish the blocks of class c = CC, in two steps with a warp barrier between them; every thread of the
// warp calls both (`on` = takes part).  Step a: fold the two quadrant slots when they belong to one
// block; the H = 1 thread of the centre-column classes (a = b = 0) parks its half of the shared block.
template <class C, int CC>
__host__ __device__ inline void p2_flush_a(const NodeCtx &N, P2State &S, bool on, double *parkbase)
{
  if (!on) return;
  if (N.fold) {
#pragma unroll
    for (int t = 0; t < 9; ++t) S.acc[CC][0][t] += S.acc[CC][1][t];
    S.D[CC][0] += S.D[CC][1];
  }
  if (N.pair && N.H == 1) {
    double *park = parkbase + (N.nl * 2 + CC) * 9;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) park[i * 3 + j] = S.acc[CC][0][i * 3 + j] + (i == j ? S.D[CC][0] : 0.0);
  }
}
// Step b: write the finished blocks into the staging row (CSR layout) and take their share of
// rhs = -row . u.  `slots`: byte j = CSR slot of block j; `uplane` = staged plane of the column nodes.
template <class C, int CC>
__host__ __device__ inline void p2_flush_b(const NodeCtx &N, P2State &S, bool on, unsigned slots, const double *uplane,
                                           double *stg, const double *parkbase)
{
  if (!on) return;
  double *row = stg + N.nl * C::SROW;
  const double *park = parkbase + (N.nl * 2 + CC) * 9;
  const int rs = N.nb * 3; // doubles per dof row
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const unsigned slot = (slots >> (8 * j)) & 0xFF;
    if (slot == 0xFF) continue; // not a block this thread finishes, or no such neighbour node
    const double *u = uplane + N.uoff[j];
    const double u0 = u[0], u1 = u[1], u2 = u[2];
    double *bp = row + slot * 3;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double v0 = S.acc[CC][j][i * 3 + 0], v1 = S.acc[CC][j][i * 3 + 1], v2 = S.acc[CC][j][i * 3 + 2];
      if (i == 0) v0 += S.D[CC][j];
      if (i == 1) v1 += S.D[CC][j];
      if (i == 2) v2 += S.D[CC][j];
      if (N.pair) {
        v0 += park[i * 3 + 0];
        v1 += park[i * 3 + 1];
        v2 += park[i * 3 + 2];
      }
      bp[i * rs + 0] = v0;
      bp[i * rs + 1] = v1;
      bp[i * rs + 2] = v2;
      S.rr[i] -= v0 * u0 + v1 * u1 + v2 * u2;
    }
  }
}

} // namespace box
} // namespace nalu_b200
