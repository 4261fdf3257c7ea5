// box2_tile.cuh -- second-generation tiled, write-once MomentumAdvDiff assembly for structured
// Hex8 blocks ("column-pair" kernel).  Same contract as box_tile.cuh (what
// AssembleElemSolverAlgorithm::execute + MomentumAdvDiffElemKernel::execute +
// TpetraLinearSystem::sumInto produce: src/AssembleElemSolverAlgorithm.C:74-100,
// src/kernel/MomentumAdvDiffElemKernel.C:70-186, src/TpetraLinearSystem.C:1457-1522), new work
// decomposition, driven by what was measured on the B200 (shared-memory
// wavefronts, instruction-cache footprint, per-warp latency, cost of a bulk-copy issue):
//
//   * a CTA owns an 8 x 4 column of nodes and sweeps z; a node row is served by 4 threads, one
//     per WARP-UNIFORM class (a,b) in {0,1}^2: the thread owns the column nodes at offset
//     (+-a, +-b, *) and computes, for every integration point it visits, BOTH the column in the
//     row's plane (c = 0) and the one in the neighbouring plane (c = 1) from one 7 x LDS.128 read
//     (half the shared-memory traffic of one column per thread).  The shape-function constants of
//     a class are per-thread registers fetched from shared memory (classes are data, not code:
//     per-class instruction streams overflow the instruction cache -- measured 43 % no-instruction
//     stalls); the outward signs tau_x, tau_y of a face are compile-time (quadrants are unrolled)
//     and applied as FMA operand modifiers;
//   * only 8 of an element's 12 integration points are resident at a time: the 4 z-edge points
//     (used by the node plane below and the one above) and the 4 points of one face (bottom face
//     for the rows below the layer, top face for the rows above it) -- phase 1 runs twice per plane;
//     records of elements outside the domain are zero, so phase 2 carries no boundary tests;
//   * finished 3x3 blocks go to a staging image of the row's dz = 0 group followed by its dz = -1
//     (lower half) or dz = +1 (upper half) group, laid out like the CSR row and phase-aligned with
//     it modulo 16 bytes, and leave through cp.async.bulk: one 27-double run per dof row for dz = -1,
//     one 54-double run for dz = 0 and dz = +1 together;
//   * nodal fields and mass flow rates are staged with cp.async one phase ahead, node metadata
//     (row offsets, the 27 column slots) with loads split across phases: no global-memory latency
//     sits on a phase's critical path;
//   * the div(u) stress needs no integration-point scalar: with acc = sum tau h (x) A accumulated for
//     the viscous block anyway, rhs_i -= 2/3 mu divU A_i  ==  rhs_i += 2/3 sum_c sum_j acc_c[j][i] u_cj
//     (MomentumAdvDiffElemKernel.C:105-136 uses the same dndx and mu for divU and the viscous lhs).
//
// Every phase is a plain function of the thread index so that tests/cpu_harness can execute the
// kernel thread by thread on the CPU (no GPU in the development container).
#pragma once
#include "box_tile.cuh"

namespace nalu_b200 {
namespace box2 {

using box::BoxParams;
using box::bit_x;
using box::bit_y;
using box::butterfly8;
using box::ip_of;
using box::node_of_bits;

struct Cfg {
  static constexpr int TBX = 8, TBY = 4, NNODE = TBX * TBY, NCLS = 4, NT = NCLS * NNODE;
  static constexpr int ELX = TBX + 1, ELY = TBY + 1, NEL = ELX * ELY;
  static constexpr int IPD = 14;          // doubles per ip record: G~[9], A[3], mdot, pad
  static constexpr int ES = 4 * IPD + 2;  // element stride; ES/2 odd: the 8 lanes of an LDS.128 phase hit 8 bank groups
  static constexpr int RNX = TBX + 2, RNY = TBY + 2, RN = RNX * RNY, NF = 7, PLN = RN * NF;
  static constexpr int MDB = NEL * 12;
  // staging of one node: 3 dof rows x [parity pad | dz=0 group (27) | dz=-1 or dz=+1 group (27)]; the two
  // parts are adjacent like in the CSR row, so that dz=0 and dz=+1 leave as ONE 54-double run
  static constexpr int SROWP = 56;
  static constexpr int SNODE = 3 * SROWP + 1;  // odd: spreads the nodes of a warp over the banks
  static constexpr int NP1A = 4 * NEL, NP1B = 2 * NEL;
  // shared-memory map (doubles)
  static constexpr int OFF_F = 0, OFF_ZE = OFF_F + NEL * ES, OFF_S = OFF_ZE + NEL * ES,
                       OFF_RING = OFF_S + NNODE * SNODE, OFF_MD = OFF_RING + 3 * PLN, OFF_RP = OFF_MD + MDB,
                       OFF_META = OFF_RP + NCLS * NNODE * 3;
  // node metadata of one plane: rowbase int64[NNODE]; nb int[NNODE], row int[NNODE];
  // slots uint8[NNODE][28]; gp uint16[NNODE][4]: [g] = first slot of dz group g | its staging offset << 8,
  // [3] = parity bits of the three dof rows | merged flag << 3; cnt uint8[NNODE][4]: columns of group g
  static constexpr int META_RB = 0, META_NB = NNODE, META_ROW = META_NB + NNODE / 2, META_SLOT = META_ROW + NNODE / 2,
                       META_GP = META_SLOT + NNODE * 28 / 8, META_CNT = META_GP + NNODE * 8 / 8,
                       META = META_CNT + NNODE * 4 / 8;
  static constexpr int OFF_QK = OFF_META + 2 * META;  // class constants: [NCLS][24]
  // run descriptors of one plane (descriptor copy-out): [3 dz groups][3 dof rows][NNODE] x {int64 first double in
  // `values`, int32 staging offset, int32 doubles}
  static constexpr int NRUN = 9 * NNODE, OFF_RD = OFF_QK + NCLS * 24;
  static constexpr int SMEM_DOUBLES = OFF_RD + 2 * NRUN;
};

struct Meta {
  long long *rowbase;
  int *nb, *row;
  unsigned char *slot;
  unsigned short *gp;
  unsigned char *cnt;
};
__host__ __device__ inline Meta meta_at(double *sm, int buf)
{
  double *m = sm + Cfg::OFF_META + buf * Cfg::META;
  Meta M;
  M.rowbase = reinterpret_cast<long long *>(m + Cfg::META_RB);
  M.nb = reinterpret_cast<int *>(m + Cfg::META_NB);
  M.row = reinterpret_cast<int *>(m + Cfg::META_ROW);
  M.slot = reinterpret_cast<unsigned char *>(m + Cfg::META_SLOT);
  M.gp = reinterpret_cast<unsigned short *>(m + Cfg::META_GP);
  M.cnt = reinterpret_cast<unsigned char *>(m + Cfg::META_CNT);
  return M;
}

// ---- staging of nodal fields / mass flow rates: global -> shared with cp.async (no registers) ------
// plane image [RN][NF] = (x, y, z, mu, u, v, w) of the tile's node neighbourhood at node plane zp;
// mass-flow image [NEL][12] of element layer ez.  Entries outside the mesh read zero (cp.async
// with source size 0).  The caller waits (stage_wait) and passes a CTA barrier before use.
__host__ __device__ __forceinline__ void copy8_or_zero(double *dst, const double *src, bool valid)
{
#ifdef __CUDA_ARCH__
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src),
               "r"(valid ? 8 : 0)
               : "memory");
#else
  *dst = valid ? *src : 0.0;
#endif
}
__host__ __device__ __forceinline__ void stage_wait()
{
#ifdef __CUDA_ARCH__
  asm volatile("cp.async.wait_all;" ::: "memory");
#endif
}
__host__ __device__ inline void stage_plane(const BoxParams &P, double *ring_plane, int tid, int x0, int y0, int zp)
{
  using C = Cfg;
  const long long pn = (long long)P.NX * P.NY;
  const bool zin = zp >= 0 && zp < P.NZ;
  // one node per thread: the index arithmetic is paid once per node, not once per double
  if (tid < C::RN) {
    const int rn = tid;
    const int x = x0 - 1 + rn % C::RNX, y = y0 - 1 + rn / C::RNX;
    const bool valid = zin && x >= 0 && y >= 0 && x < P.NX && y < P.NY;
    const long long n = valid ? x + (long long)P.NX * y + pn * zp : 0;
    double *d = ring_plane + rn * C::NF;
    const double *xc = P.coords + n * 3, *u = P.vel + n * 3;
    copy8_or_zero(d + 0, xc + 0, valid);
    copy8_or_zero(d + 1, xc + 1, valid);
    copy8_or_zero(d + 2, xc + 2, valid);
    copy8_or_zero(d + 3, P.visc + n, valid);
    copy8_or_zero(d + 4, u + 0, valid);
    copy8_or_zero(d + 5, u + 1, valid);
    copy8_or_zero(d + 6, u + 2, valid);
  }
}
// 16 bytes at a time: the 12 mass flow rates of an element are 96 contiguous, 16-byte aligned bytes on both sides
__host__ __device__ __forceinline__ void copy16_or_zero(double *dst, const double *src, bool valid)
{
#ifdef __CUDA_ARCH__
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src),
               "r"(valid ? 16 : 0)
               : "memory");
#else
  dst[0] = valid ? src[0] : 0.0;
  dst[1] = valid ? src[1] : 0.0;
#endif
}
__host__ __device__ inline void stage_mdot(const BoxParams &P, double *mdb, int tid, int x0, int y0, int ez)
{
  using C = Cfg;
  const long long ln = (long long)P.nx * P.ny;
  const bool zin = ez >= 0 && ez < P.nz;
  const bool wide = (((unsigned long long)P.mdot) & 15) == 0; // uniform
#pragma unroll 1
  for (int idx = tid; idx < C::NEL * 6; idx += C::NT) {
    const int el = idx / 6, pr = idx - el * 6;
    const int ex = x0 - 1 + el % C::ELX, ey = y0 - 1 + el / C::ELX;
    const bool valid = zin && ex >= 0 && ey >= 0 && ex < P.nx && ey < P.ny;
    const long long e = valid ? ex + (long long)P.nx * ey + ln * ez : 0;
    const double *src = P.mdot + e * 12 + 2 * pr;
    double *dst = mdb + el * 12 + 2 * pr;
    if (wide) copy16_or_zero(dst, src, valid);
    else {
      copy8_or_zero(dst, src, valid);
      copy8_or_zero(dst + 1, src + 1, valid);
    }
  }
}

// ---- phase 1: one (element, edge direction g, side) item = 2 integration points ------------------
// Local axes (t,u,v) = (g, g+1, g+2) mod 3.  g = 0,1: `side` is the z side of the face (0 bottom:
// ips 0-3 -> F, 1 top: ips 4-7 -> F); g = 2: `side` is the y side, ips 8-11 -> ZE.
// Record: G~[i][j] = -mu adj(J)[i][j]/det with columns u, v pre-multiplied by the ip's side signs
// (so that a viewer only needs class constants), A = area vector pointing along +t (NOT left->right),
// mdot counted along +t.
__host__ __device__ inline void p1_item(const BoxParams &P, int el, int g, int side, int x0, int y0, const double *pl0,
                                        const double *pl1, const double *mdb, double *F, double *ZE)
{
  using C = Cfg;
  const int elx = el % C::ELX, ely = el / C::ELX;
  const int ex = x0 - 1 + elx, ey = y0 - 1 + ely;
  if (ex < 0 || ey < 0 || ex >= P.nx || ey >= P.ny) return;
  const int sx = C::NF, sy = C::RNX * C::NF;
  const int s_t = g == 0 ? sx : (g == 1 ? sy : 0), s_u = g == 0 ? sy : (g == 1 ? 0 : sx),
            s_v = g == 0 ? 0 : (g == 1 ? sx : sy);
  const int zbit = g == 0 ? 4 : (g == 1 ? 2 : 1);
  const int n000 = (elx + C::RNX * ely) * C::NF;
  enum { T = 1, Uu = 2, V = 4, TU = 3, TV = 5, UV = 6, TUV = 7 };
  const double q = (P.options & NALU_OPT_SHIFTED_GRAD_OP) ? 0.5 : 0.25;
  const int at = g, au = (g + 1) % 3;
  double X[3][8], M[8];
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const double *nd = ((m & zbit) ? pl1 : pl0) + n000 + (m & 1) * s_t + ((m >> 1) & 1) * s_u + ((m >> 2) & 1) * s_v;
    X[0][m] = nd[0];
    X[1][m] = nd[1];
    X[2][m] = nd[2];
    M[m] = nd[3];
  }
  // coordinates relative to the item's first node: only difference coefficients are used below, and differences
  // of partial sums of large coordinates would cost |x|/h ulps
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double o = X[i][0];
#pragma unroll
    for (int m = 0; m < 8; ++m) X[i][m] -= o;
  }
  butterfly8(X[0]);
  butterfly8(X[1]);
  butterfly8(X[2]);
  butterfly8(M);
  double base[3];
  base[0] = (X[1][Uu] * X[2][V] - X[2][Uu] * X[1][V]) * (1.0 / 64.0);
  base[1] = (X[2][Uu] * X[0][V] - X[0][Uu] * X[2][V]) * (1.0 / 64.0);
  base[2] = (X[0][Uu] * X[1][V] - X[1][Uu] * X[0][V]) * (1.0 / 64.0);
#pragma unroll
  for (int ff = 0; ff < 2; ++ff) {
    const int b1 = (g == 1) ? side : ff, b2 = (g == 1) ? ff : side;
    const double sg1 = b1 ? 1.0 : -1.0, sg2 = b2 ? 1.0 : -1.0;
    // the node of the ip's edge with bit_t = 0
    const int obx = at == 0 ? 0 : (au == 0 ? b1 : b2), oby = at == 1 ? 0 : (au == 1 ? b1 : b2),
              obz = at == 2 ? 0 : (au == 2 ? b1 : b2);
    const int k0 = node_of_bits(obx, oby, obz);
    const int ip = ip_of(k0, g);
    const double orient = (hex8::scs_left(ip) == k0) ? 1.0 : -1.0;
    double J[3][3];
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
    // mu at the (never shifted) ip location, MomentumAdvDiffElemKernel.C:107-110
    const double mu = 0.125 * M[0] + 0.0625 * (sg1 * M[Uu] + sg2 * M[V]) + 0.03125 * sg1 * sg2 * M[UV];
    const double fac = -mu * (1.0 / det); // reciprocal + multiply: half the instructions of a full FP64 division
    const double fu = fac * sg1, fv = fac * sg2;
    double *o = (g == 2 ? ZE + el * C::ES + (ip - 8) * C::IPD : F + el * C::ES + (ip & 3) * C::IPD);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      o[i * 3 + 0] = fac * adj[i][0];
      o[i * 3 + 1] = fu * adj[i][1];
      o[i * 3 + 2] = fv * adj[i][2];
    }
    double W[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) W[i] = sg2 * X[i][V] - sg1 * X[i][Uu];
    o[9] = base[0] + (X[1][UV] * W[2] - X[2][UV] * W[1]) * (1.0 / 128.0);
    o[10] = base[1] + (X[2][UV] * W[0] - X[0][UV] * W[2]) * (1.0 / 128.0);
    o[11] = base[2] + (X[0][UV] * W[1] - X[1][UV] * W[0]) * (1.0 / 128.0);
    o[12] = orient * mdb[el * 12 + ip];
    o[13] = 0.0;
  }
}

// ---- phase 2 ---------------------------------------------------------------------------------------
// Uniform constants.  In the frame of the row node an ip sits at 0 along its edge and at distance q
// towards the row node on the two transverse axes; phi = 1/2 + q for a column on the row's side of
// a transverse axis, 1/2 - q otherwise (Hex8CVFEM.C:24-70,318-345).
struct P2K {
  double kt[3]; // phi_u phi_v for 0, 1, 2 transverse axes on which the column differs from the row
  double kh[2]; // 1/2 phi
  double kn[3]; // advection shape function 1/2 psi_u psi_v (psi: q = 1/2 when skew-symmetric)
  double f23;   // 2/3 includeDivU
};
__host__ __device__ inline void p2k_init(P2K &K, int options, double include_divu)
{
  const double q = (options & NALU_OPT_SHIFTED_GRAD_OP) ? 0.5 : 0.25, qa = (options & NALU_OPT_SKEW_SYMMETRIC) ? 0.5 : 0.25;
  const double p0 = 0.5 + q, p1 = 0.5 - q, a0 = 0.5 + qa, a1 = 0.5 - qa;
  K.kt[0] = p0 * p0; K.kt[1] = p0 * p1; K.kt[2] = p1 * p1;
  K.kh[0] = 0.5 * p0; K.kh[1] = 0.5 * p1;
  K.kn[0] = 0.5 * a0 * a0; K.kn[1] = 0.5 * a0 * a1; K.kn[2] = 0.5 * a1 * a1;
  K.f23 = 2.0 / 3.0 * include_divu;
}

__host__ __device__ __forceinline__ double flip_if(double v, bool neg)
{
#ifdef __CUDA_ARCH__
  return __hiloint2double(__double2hiint(v) ^ (neg ? 0x80000000 : 0), __double2loint(v));
#else
  return neg ? -v : v;
#endif
}

struct D2 {
  double x, y;
};

// per-class constants.  Class (a,b): the thread's two column nodes differ from the row node along
// x iff a, along y iff b, and along z iff c (c = 0: the column in the row's plane, c = 1: the one in
// the neighbouring plane).  For the ip on the row's edge along axis pp, local axes (t,u,v) =
// (pp, pp+1, pp+2):  tau d_t = c0[c][pp],  tau d_u = tau c1[c][pp],  tau d_v = tau c2[c][pp],
// tau N_adv = tau cn[c][pp], where tau is the outward sign of the face for the row node.
// Classes are data, not code: all warps run one instruction stream (per-class instruction streams
// overflow the instruction cache: measured 43 % no-instruction stalls).
struct ClsK {
  double c0[2][3], c1[2][3], c2[2][3], cn[2][3];
  int a, b;
};
__host__ __device__ inline void cls_init(ClsK &Q, const P2K &K, int cls)
{
  Q.a = cls & 1;
  Q.b = (cls >> 1) & 1;
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int rho[3] = {Q.a, Q.b, c};
#pragma unroll
    for (int pp = 0; pp < 3; ++pp) {
      const int rt = rho[pp], ru = rho[(pp + 1) % 3], rv = rho[(pp + 2) % 3];
      Q.c0[c][pp] = rt ? K.kt[ru + rv] : -K.kt[ru + rv];
      Q.c1[c][pp] = ru ? -K.kh[rv] : K.kh[rv];
      Q.c2[c][pp] = rv ? -K.kh[ru] : K.kh[ru];
      Q.cn[c][pp] = K.kn[ru + rv];
    }
  }
}
__host__ __device__ inline void cls_store(const ClsK &Q, double *qk)
{
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int pp = 0; pp < 3; ++pp) {
      qk[c * 12 + pp] = Q.c0[c][pp];
      qk[c * 12 + 3 + pp] = Q.c1[c][pp];
      qk[c * 12 + 6 + pp] = Q.c2[c][pp];
      qk[c * 12 + 9 + pp] = Q.cn[c][pp];
    }
}
__host__ __device__ inline void cls_fetch(ClsK &Q, const double *qk, int cls)
{
  Q.a = cls & 1;
  Q.b = (cls >> 1) & 1;
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int pp = 0; pp < 3; ++pp) {
      Q.c0[c][pp] = qk[c * 12 + pp];
      Q.c1[c][pp] = qk[c * 12 + 3 + pp];
      Q.c2[c][pp] = qk[c * 12 + 6 + pp];
      Q.cn[c][pp] = qk[c * 12 + 9 + pp];
    }
}

// one integration point: acc += (tau h) (x) A, dm += tau N mdot.  Load and arithmetic are separate
// so that the caller can keep the next record in flight while this one is consumed.
struct Rec {
  double o[14];
};
__host__ __device__ __forceinline__ void rec_load(Rec &R, const double *rec)
{
#ifdef __CUDA_ARCH__
  const double2 *r2 = reinterpret_cast<const double2 *>(rec);
#else
  const D2 *r2 = reinterpret_cast<const D2 *>(rec);
#endif
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    R.o[2 * k] = r2[k].x;
    R.o[2 * k + 1] = r2[k].y;
  }
}
// NEG: the outward sign tau of the face is -1 (applies to the u, v derivative terms and to mdot;
// tau d_t is sign-free, see ClsK).  Written so that the sign is an operand modifier of the FMA.
template <bool NEG>
__host__ __device__ __forceinline__ void rec_apply(const Rec &R, double c0, double c1, double c2, double cn,
                                                   double (&acc)[9], double &dm)
{
  const double *o = R.o;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double w = o[i * 3 + 1] * c1 + o[i * 3 + 2] * c2;
    const double h = NEG ? o[i * 3 + 0] * c0 - w : o[i * 3 + 0] * c0 + w;
    acc[i * 3 + 0] += h * o[9];
    acc[i * 3 + 1] += h * o[10];
    acc[i * 3 + 2] += h * o[11];
  }
  if (NEG) dm -= cn * o[12];
  else dm += cn * o[12];
}

struct T2 {
  int lx, ly, node;
  int active, vmask;
  double rr[3];
};

__host__ __device__ inline void t2_init(const BoxParams &P, T2 &N, int lane, int x0, int y0)
{
  using C = Cfg;
  N.node = lane;
  N.lx = lane % C::TBX;
  N.ly = lane / C::TBX;
  const int x = x0 + N.lx, y = y0 + N.ly;
  N.active = x < P.NX && y < P.NY;
  N.vmask = 0;
  for (int q = 0; q < 4; ++q) {
    const int ex = x - 1 + (q & 1), ey = y - 1 + (q >> 1);
    if (ex >= 0 && ey >= 0 && ex < P.nx && ey < P.ny) N.vmask |= 1 << q;
  }
}

// ---- flush of the finished blocks of one (sx,sy) group -> staging; rhs share ------------------------
// H = 0: elements below the plane (column c = 1 fills the dz = -1 group, c = 0 the first half of
// dz = 0), H = 1: above (c = 1 -> dz = +1, c = 0 completes dz = 0 on top of the partial left in the
// staging row when `has_lower`).  Split in three so that no shared-memory latency sits between the
// last FMA of a block and its stores (the dependent slot -> offset -> operand chain was measured at
// 600-1700 cycles per flush): flush_prep (start of the pass: slot bytes -> staging offsets),
// flush_fetch (start of the quadrant: u of the column nodes, partial of the dz = 0 block),
// flush_commit (arithmetic and stores only).
struct FlushPrep {
  int s[2];   // staging offset of the (i = 0) row of the block of column c, < 0: no such neighbour
  int uo[2];  // offset of the column node's velocity in the ring
};
struct FlushOps {
  double u[2][3];
  double part[9];
};
__host__ __device__ __forceinline__ void flush_prep(FlushPrep &F, const T2 &N, const Meta &M, const ClsK &Q, int H, int sx,
                                                    int sy, int z, unsigned gp0, unsigned gp1)
{
  using C = Cfg;
  const int dx = Q.a ? (sx ? 1 : -1) : 0, dy = Q.b ? (sy ? 1 : -1) : 0;
  const int idx = (dy + 1) * 3 + (dx + 1);
  const int g1 = H ? 2 : 0;
  const unsigned slot0 = M.slot[N.node * 28 + 9 + idx], slot1 = M.slot[N.node * 28 + g1 * 9 + idx];
  F.s[0] = slot0 == 0xFF ? -1 : (int)(N.node * C::SNODE + (gp0 >> 8) + (slot0 - (gp0 & 0xFF)) * 3);
  F.s[1] = slot1 == 0xFF ? -1 : (int)(N.node * C::SNODE + (gp1 >> 8) + (slot1 - (gp1 & 0xFF)) * 3);
  const int uo = ((N.lx + 1 + dx) + C::RNX * (N.ly + 1 + dy)) * C::NF + 4;
  F.uo[0] = C::OFF_RING + (z % 3) * C::PLN + uo;
  F.uo[1] = C::OFF_RING + ((z + (H ? 1 : 2)) % 3) * C::PLN + uo;
}
__host__ __device__ __forceinline__ void flush_fetch(FlushOps &O, const FlushPrep &F, const double *sm, unsigned pb, bool add)
{
  using C = Cfg;
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int j = 0; j < 3; ++j) O.u[c][j] = sm[F.uo[c] + j];
#pragma unroll
  for (int t = 0; t < 9; ++t) O.part[t] = 0.0;
  if (add && F.s[0] >= 0) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double *d = sm + C::OFF_S + F.s[0] + i * C::SROWP + ((pb >> i) & 1);
      O.part[i * 3 + 0] = d[0];
      O.part[i * 3 + 1] = d[1];
      O.part[i * 3 + 2] = d[2];
    }
  }
}
__host__ __device__ __forceinline__ void flush_commit(T2 &N, double *sm, const FlushPrep &F, const FlushOps &O, unsigned pb,
                                                      const double (&acc)[2][9], const double (&dm)[2], double f23)
{
  using C = Cfg;
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    if (F.s[c] < 0) continue; // no such neighbour
    const double u0 = O.u[c][0], u1 = O.u[c][1], u2 = O.u[c][2];
    const double D = dm[c] + (acc[c][0] + acc[c][4] + acc[c][8]);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double v0 = acc[c][i * 3 + 0], v1 = acc[c][i * 3 + 1], v2 = acc[c][i * 3 + 2];
      if (i == 0) v0 += D;
      if (i == 1) v1 += D;
      if (i == 2) v2 += D;
      N.rr[i] -= v0 * u0 + v1 * u1 + v2 * u2;
      N.rr[i] += f23 * (acc[c][0 + i] * u0 + acc[c][3 + i] * u1 + acc[c][6 + i] * u2);
      double *d = sm + C::OFF_S + F.s[c] + i * C::SROWP + ((pb >> i) & 1);
      if (c == 0) { // the partial of the lower half (zero when there is none)
        v0 += O.part[i * 3 + 0];
        v1 += O.part[i * 3 + 1];
        v2 += O.part[i * 3 + 2];
      }
      d[0] = v0;
      d[1] = v1;
      d[2] = v2;
    }
  }
}

// record k (0..2: x edge, y edge, z edge ip) of quadrant (SX,SY: the element lies on the +x / +y side
// of the row node).  The row node is corner (!sx, !sy, !H) of the element (Hex8CVFEM.C:264-276).
template <int SX, int SY, int K>
__host__ __device__ __forceinline__ const double *rec_addr(const double *eF, const double *eZ)
{
  using C = Cfg;
  constexpr int eo = (SX + C::ELX * SY) * C::ES;
  constexpr int fx = SY ? 0 : 2;               // x edge of the face: ip & 3
  constexpr int fy = SX ? 3 : 1;               // y edge
  constexpr int fz = (SY ? 0 : 2) + (SX ^ SY); // z edge above/below bottom node (k & 3)
  return K == 0 ? eF + eo + fx * C::IPD : (K == 1 ? eF + eo + fy * C::IPD : eZ + eo + fz * C::IPD);
}
// the outward signs tau_x, tau_y are compile-time, tau_z comes in through z1, z2, zn; one record feeds
// both column nodes of the thread
template <int SX, int SY, int K>
__host__ __device__ __forceinline__ void rec_use(const Rec &R, const ClsK &Q, const double (&z1)[2], const double (&z2)[2],
                                                 const double (&zn)[2], double (&acc)[2][9], double (&dm)[2])
{
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    if (K == 0) rec_apply<!SX>(R, Q.c0[c][0], Q.c1[c][0], Q.c2[c][0], Q.cn[c][0], acc[c], dm[c]);
    else if (K == 1) rec_apply<!SY>(R, Q.c0[c][1], Q.c1[c][1], Q.c2[c][1], Q.cn[c][1], acc[c], dm[c]);
    else rec_apply<false>(R, Q.c0[c][2], z1[c], z2[c], zn[c], acc[c], dm[c]);
  }
}

// one z half (H = 0 below / 1 above the plane) of the node's octants.  Quadrants that fold into the
// same column node are visited back to back so that one 3x3 accumulator is live; the flush points
// are a function of the (warp-uniform) class: (a,b) = (1,1): every quadrant is its own column;
// (0,1): pairs differing in sx fold; (1,0): pairs differing in sy fold (SWAP: visit sy fastest);
// (0,0): all four fold.  No test for quadrants / nodes outside the domain: their records are zero
// (never written by phase 1 after the prologue cleared F / ZE) and their flushes see slot 0xFF.
// The 12 records are double-buffered in registers: record k+1 is loaded before record k is used.
template <bool SWAP>
__host__ __device__ inline void p2_half_t(T2 &N, double *sm, const ClsK &Q, double f23, int H, int z, bool has_lower)
{
  using C = Cfg;
  const double *eF = sm + C::OFF_F + (N.lx + C::ELX * N.ly) * C::ES;
  const double *eZ = sm + C::OFF_ZE + (N.lx + C::ELX * N.ly) * C::ES;
  const unsigned fm = (Q.a & Q.b) ? 0xFu : ((Q.a | Q.b) ? 0xAu : 0x8u);
  // z-edge ips: tau_z = +1 for the elements above the plane
  double z1[2], z2[2], zn[2];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    z1[c] = flip_if(Q.c1[c][2], !H);
    z2[c] = flip_if(Q.c2[c][2], !H);
    zn[c] = flip_if(Q.cn[c][2], !H);
  }
  double acc[2][9], dm[2];
#pragma unroll
  for (int t = 0; t < 9; ++t) acc[0][t] = acc[1][t] = 0.0;
  dm[0] = dm[1] = 0.0;
  constexpr int X1 = SWAP ? 0 : 1, Y1 = SWAP ? 1 : 0, X2 = SWAP ? 1 : 0, Y2 = SWAP ? 0 : 1;
  // staging offsets and velocity addresses of the (up to) four flushes of this pass
  const Meta M = meta_at(sm, z & 1);
  const unsigned gp0 = M.gp[N.node * 4 + 1], gp1 = M.gp[N.node * 4 + (H ? 2 : 0)], pb = M.gp[N.node * 4 + 3];
  const bool add = H && has_lower;
  FlushPrep FP0, FP1, FP2, FP3;
  flush_prep(FP0, N, M, Q, H, 0, 0, z, gp0, gp1);
  flush_prep(FP1, N, M, Q, H, X1, Y1, z, gp0, gp1);
  flush_prep(FP2, N, M, Q, H, X2, Y2, z, gp0, gp1);
  flush_prep(FP3, N, M, Q, H, 1, 1, z, gp0, gp1);
  FlushOps FO;
  Rec RA, RB;
  rec_load(RA, rec_addr<0, 0, 0>(eF, eZ));
  // the record buffers swap roles from quadrant to quadrant (three records each: A, B, A then B, A, B): the
  // first record of the next quadrant is loaded into the buffer the next quadrant starts with, no copy
#define NALU_P2_QUAD(J, SX, SY, NX_, NY_, LAST, RA, RB)                                           \
  if (fm & (1u << J)) flush_fetch(FO, FP##J, sm, pb, add);                                        \
  rec_load(RB, rec_addr<SX, SY, 1>(eF, eZ));                                                      \
  rec_use<SX, SY, 0>(RA, Q, z1, z2, zn, acc, dm);                                                 \
  rec_load(RA, rec_addr<SX, SY, 2>(eF, eZ));                                                      \
  rec_use<SX, SY, 1>(RB, Q, z1, z2, zn, acc, dm);                                                 \
  if (!LAST) rec_load(RB, rec_addr<NX_, NY_, 0>(eF, eZ));                                         \
  rec_use<SX, SY, 2>(RA, Q, z1, z2, zn, acc, dm);                                                 \
  if (fm & (1u << J)) {                                                                           \
    flush_commit(N, sm, FP##J, FO, pb, acc, dm, f23);                                             \
    _Pragma("unroll") for (int t = 0; t < 9; ++t) acc[0][t] = acc[1][t] = 0.0;                    \
    dm[0] = dm[1] = 0.0;                                                                          \
  }
  NALU_P2_QUAD(0, 0, 0, X1, Y1, false, RA, RB)
  NALU_P2_QUAD(1, X1, Y1, X2, Y2, false, RB, RA)
  NALU_P2_QUAD(2, X2, Y2, 1, 1, false, RA, RB)
  NALU_P2_QUAD(3, 1, 1, 1, 1, true, RB, RA)
#undef NALU_P2_QUAD
}

__host__ __device__ inline void p2_half(T2 &N, double *sm, const ClsK &Q, double f23, int H, int z, bool has_lower)
{
  // nodes outside the domain run the same code on zero records; their flushes see slot 0xFF
  if (Q.a && !Q.b) p2_half_t<true>(N, sm, Q, f23, H, z, has_lower);
  else p2_half_t<false>(N, sm, Q, f23, H, z, has_lower);
}

// ---- node metadata of plane z: thread (node, g) reads the 9 slots of dz group g ------------------
// One thread per node.  Split in three so that the two dependent global loads (node_row -> browptr)
// and the slot bytes are in flight while other work runs: meta_row one phase ahead, meta_fetch
// before phase 1a, meta_store after.
struct MetaRaw {
  unsigned s[27];
  long long b0, b1;
  int row;
};
__host__ __device__ inline int meta_row(const BoxParams &P, int node, int x0, int y0, int z)
{
  using C = Cfg;
  const int x = x0 + node % C::TBX, y = y0 + node / C::TBX;
  if (!(x < P.NX && y < P.NY && z >= 0 && z < P.NZ)) return -1;
  return P.node_row[x + (long long)P.NX * (y + (long long)P.NY * z)];
}
__host__ __device__ inline void meta_fetch(const BoxParams &P, MetaRaw &R, int node, int x0, int y0, int z, int row)
{
  using C = Cfg;
  R.row = row;
  R.b0 = R.b1 = 0;
#pragma unroll
  for (int k = 0; k < 27; ++k) R.s[k] = 0xFF;
  if (row < 0) return;
  const int x = x0 + node % C::TBX, y = y0 + node / C::TBX;
  const long long n = x + (long long)P.NX * (y + (long long)P.NY * z);
  const unsigned char *sl = P.slot27 + n * 27;
#pragma unroll
  for (int k = 0; k < 27; ++k) R.s[k] = sl[k];
  R.b0 = P.browptr[row];
  R.b1 = P.browptr[row + 1];
}
__host__ __device__ inline void meta_store(const BoxParams &P, const Meta &M, const MetaRaw &R, int node)
{
  using C = Cfg;
  unsigned mn[3], cnt[3];
#pragma unroll
  for (int g = 0; g < 3; ++g) {
    unsigned lo = 255, hi = 0, c = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const unsigned s = R.s[g * 9 + k];
      M.slot[node * 28 + g * 9 + k] = (unsigned char)s;
      if (s != 0xFF) {
        lo = s < lo ? s : lo;
        hi = s > hi ? s : hi;
        ++c;
      }
    }
    if (c && hi - lo + 1 != c) *P.err = 2; // the columns of a dz group must be contiguous in the row
    mn[g] = c ? lo : 0;
    cnt[g] = c;
  }
  const int nb = (int)(R.b1 - R.b0);
  const long long rowbase = 9 * R.b0;
  // parity of dof row i: staging offset of the dz=0 group = 16-byte phase of its place in `values`
  unsigned pb = 0;
#pragma unroll
  for (int i = 0; i < 3; ++i) pb |= (unsigned)((rowbase + ((long long)i * nb + mn[1]) * 3 + node) & 1) << i;
  // the dz=-1 / dz=+1 group sits right behind the dz=0 group (as in the row); one more double if its
  // own 16-byte phase demands it (only when the groups are not neighbours in the row)
  unsigned base[3], fix2 = 0;
  base[1] = 0;
#pragma unroll
  for (int g = 0; g < 3; g += 2) {
    const unsigned b = cnt[1] * 3;
    const unsigned fix = (unsigned)((node + (pb & 1) + b + rowbase + (long long)mn[g] * 3) & 1);
    base[g] = b + fix;
    if (g == 2) fix2 = fix;
  }
  const unsigned merged = cnt[2] && cnt[1] && mn[2] == mn[1] + cnt[1] && !fix2;
#pragma unroll
  for (int g = 0; g < 3; ++g) {
    M.gp[node * 4 + g] = (unsigned short)(mn[g] | (base[g] << 8));
    M.cnt[node * 4 + g] = (unsigned char)cnt[g];
  }
  M.gp[node * 4 + 3] = (unsigned short)(pb | (merged << 3));
  M.rowbase[node] = rowbase;
  M.nb[node] = nb;
  M.row[node] = R.row < 0 ? 0 : R.row;
  (void)C::SROWP;
}
__host__ __device__ inline void meta_load(const BoxParams &P, const Meta &M, int node, int x0, int y0, int z)
{
  MetaRaw R;
  meta_fetch(P, R, node, x0, y0, z, meta_row(P, node, x0, y0, z));
  meta_store(P, M, R, node);
}

// ---- copy-out of dof row i of dz group g of one node: where it lives -------------------------------
struct Run {
  long long G;   // first double in values
  int L;         // doubles
  int soff;      // offset in the group's staging array
};
__host__ __device__ inline Run run_of(const Meta &M, int node, int i, int g)
{
  using C = Cfg;
  Run r;
  const unsigned gp = M.gp[node * 4 + g], nf = M.gp[node * 4 + 3];
  const bool merged = (nf >> 3) & 1;
  r.L = 3 * M.cnt[node * 4 + g];
  if (merged && g == 1) r.L += 3 * M.cnt[node * 4 + 2]; // dz=0 and dz=+1 leave as one run
  if (merged && g == 2) r.L = 0;
  r.G = M.rowbase[node] + ((long long)i * M.nb[node] + (gp & 0xFF)) * 3;
  r.soff = node * C::SNODE + i * C::SROWP + ((nf >> i) & 1) + (gp >> 8);
  return r;
}

// descriptor copy-out: the (never merged, <= 27 doubles) run of dz group g of dof row i of a node, stored once
// per plane so that the copy loop is descriptor load -> staging load -> store with no index arithmetic
struct RunDesc {
  double *dst; // first double of the run in `values` (the address itself: nothing left to compute in the copy loop)
  int soff, L;
};
__host__ __device__ inline void build_run_desc(const Meta &M, double *sm, int tid, double *values)
{
  using C = Cfg;
  RunDesc *rd = reinterpret_cast<RunDesc *>(sm + C::OFF_RD);
  for (int idx = tid; idx < C::NRUN; idx += C::NT) {
    const int g = idx / (3 * C::NNODE), rr = idx - g * 3 * C::NNODE;
    const int i = rr / C::NNODE, node = rr - i * C::NNODE;
    const unsigned gp = M.gp[node * 4 + g], nf = M.gp[node * 4 + 3];
    RunDesc d;
    d.L = 3 * M.cnt[node * 4 + g];
    d.dst = values + (M.rowbase[node] + ((long long)i * M.nb[node] + (gp & 0xFF)) * 3);
    d.soff = node * C::SNODE + i * C::SROWP + ((nf >> i) & 1) + (gp >> 8);
    rd[idx] = d;
  }
}

// columns of the row that no local element touches (remote columns announced by the ranks sharing
// the node): in overwrite mode they must read zero.  Thread (node, i) handles dof row i.
__host__ __device__ inline void zero_uncovered(const Meta &M, int node, int i, double *values)
{
  const int nb = M.nb[node];
  const int c0 = M.cnt[node * 4 + 0], c1 = M.cnt[node * 4 + 1], c2 = M.cnt[node * 4 + 2];
  if (nb == c0 + c1 + c2) return;
  const int s0 = M.gp[node * 4 + 0] & 0xFF, s1 = M.gp[node * 4 + 1] & 0xFF, s2 = M.gp[node * 4 + 2] & 0xFF;
  double *row = values + M.rowbase[node] + (long long)i * nb * 3;
  for (int s = 0; s < nb; ++s) {
    if ((s >= s0 && s < s0 + c0) || (s >= s1 && s < s1 + c1) || (s >= s2 && s < s2 + c2)) continue;
    row[s * 3 + 0] = row[s * 3 + 1] = row[s * 3 + 2] = 0.0;
  }
}

} // namespace box2
} // namespace nalu_b200
