// box2_tile.cuh -- second-generation tiled, write-once MomentumAdvDiff assembly for structured
// Hex8 blocks ("column-pair" kernel).  Same contract as box_tile.cuh (what
// AssembleElemSolverAlgorithm::execute + MomentumAdvDiffElemKernel::execute +
// TpetraLinearSystem::sumInto produce: src/AssembleElemSolverAlgorithm.C:74-100,
// src/kernel/MomentumAdvDiffElemKernel.C:70-186, src/TpetraLinearSystem.C:1457-1522), new
// work decomposition, chosen from the measured machine balance (tools/ubench.cu: one LDS.128 per
// clock per SM whatever the address pattern, DFMA needs 4 independent instructions in flight per
// scheduler, 208-byte cp.async.bulk stores sustain ~6 TB/s):
//
//   * a CTA owns an 8 x 4 column of nodes and sweeps z; a node row is served by 8 threads, one
//     per WARP-UNIFORM class (a,b,c) in {0,1}^3: the thread owns the column nodes at offset
//     (+-a, +-b, +-c) and reads an integration-point record with 7 x LDS.128; the shape-function
//     constants of a class are 12 per-thread registers (classes are data, not code: per-class
//     instruction streams overflow the instruction cache -- measured 43 % no-instruction stalls),
//     the only other sign is the outward direction tau of the face with respect to the row node;
//   * only 8 of an element's 12 integration points are resident at a time: the 4 z-edge points
//     (used by the node plane below and the one above) and the 4 points of one face (bottom face
//     for the rows below the layer, top face for the rows above it) -- phase 1 runs twice per plane;
//   * finished 3x3 blocks go to a staging image of ONE dz group of the row (9 columns, 27 doubles
//     per dof row) which is streamed to HBM by cp.async.bulk (one 208-byte copy per dof row and
//     group); the dz = 0 group is completed in two halves (elements below / above the plane);
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
  static constexpr int TBX = 8, TBY = 4, NNODE = TBX * TBY, NCLS = 8, NT = NCLS * NNODE;
  static constexpr int ELX = TBX + 1, ELY = TBY + 1, NEL = ELX * ELY;
  static constexpr int IPD = 14;          // doubles per ip record: G~[9], A[3], mdot, pad
  static constexpr int ES = 4 * IPD + 2;  // element stride; ES/2 odd: the 8 lanes of an LDS.128 phase hit 8 bank groups
  static constexpr int RNX = TBX + 2, RNY = TBY + 2, RN = RNX * RNY, NF = 7, PLN = RN * NF;
  static constexpr int MDB = NEL * 12;
  static constexpr int NPL = (PLN + NT - 1) / NT, NML = (MDB + NT - 1) / NT;
  static constexpr int SROWP = 28;             // staging pitch of a dof row of one dz group (27 + parity slack)
  static constexpr int SNODE = 3 * SROWP + 1;  // odd: conflict-free 64-bit stores across the nodes of a warp
  static constexpr int NP1A = 4 * NEL, NP1B = 2 * NEL;
  // shared-memory map (doubles)
  static constexpr int OFF_F = 0, OFF_ZE = OFF_F + NEL * ES, OFF_S0 = OFF_ZE + NEL * ES, OFF_S1 = OFF_S0 + NNODE * SNODE,
                       OFF_RING = OFF_S1 + NNODE * SNODE, OFF_MD = OFF_RING + 3 * PLN, OFF_RP = OFF_MD + MDB,
                       OFF_META = OFF_RP + NCLS * NNODE * 3;
  // node metadata of one plane: rowbase int64[NNODE]; nb int[NNODE], row int[NNODE];
  // slots uint8[NNODE][28]; gp uint16[NNODE][4] (group start | parity bits << 8); cnt uint8[NNODE][4]
  static constexpr int META_RB = 0, META_NB = NNODE, META_ROW = META_NB + NNODE / 2, META_SLOT = META_ROW + NNODE / 2,
                       META_GP = META_SLOT + NNODE * 28 / 8, META_CNT = META_GP + NNODE * 8 / 8,
                       META = META_CNT + NNODE * 4 / 8;
  static constexpr int OFF_QK = OFF_META + 2 * META;  // class constants: [NCLS][12]
  static constexpr int SMEM_DOUBLES = OFF_QK + NCLS * 12;
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

// ---- staging of nodal fields / mass flow rates (global -> registers -> shared) ---------------------
// Slot k of thread tid covers flat index tid + k*NT of the [RN][NF] plane image (box::plane_store)
// and of the [NEL][12] mass-flow image (box::mdot_store).  What does not depend on z is folded once
// into a packed index: pk = (x + NX y) * 8 + field, mk = (ex + nx ey) * 12 + ip, -1 = outside.
struct StageIdx {
  int pk[Cfg::NPL];
  int mk[Cfg::NML];
};
__host__ __device__ inline void stage_init(const BoxParams &P, StageIdx &S, int tid, int x0, int y0)
{
  using C = Cfg;
#pragma unroll
  for (int k = 0; k < C::NPL; ++k) {
    const int idx = tid + k * C::NT;
    S.pk[k] = -1;
    if (idx >= C::PLN) continue;
    const int rn = idx / C::NF, f = idx - rn * C::NF;
    const int x = x0 - 1 + rn % C::RNX, y = y0 - 1 + rn / C::RNX;
    if (x < 0 || y < 0 || x >= P.NX || y >= P.NY) continue;
    S.pk[k] = (x + P.NX * y) * 8 + f;
  }
#pragma unroll
  for (int k = 0; k < C::NML; ++k) {
    const int idx = tid + k * C::NT;
    S.mk[k] = -1;
    if (idx >= C::MDB) continue;
    const int el = idx / 12, ip = idx - el * 12;
    const int ex = x0 - 1 + el % C::ELX, ey = y0 - 1 + el / C::ELX;
    if (ex < 0 || ey < 0 || ex >= P.nx || ey >= P.ny) continue;
    S.mk[k] = (ex + P.nx * ey) * 12 + ip;
  }
}
__host__ __device__ inline void plane_fetch(const BoxParams &P, const StageIdx &S, int zp, double (&reg)[Cfg::NPL])
{
  const long long pn = (long long)P.NX * P.NY;
#pragma unroll
  for (int k = 0; k < Cfg::NPL; ++k) {
    reg[k] = 0.0;
    if (S.pk[k] < 0 || zp < 0 || zp >= P.NZ) continue;
    const long long n = (S.pk[k] >> 3) + pn * zp;
    const int f = S.pk[k] & 7;
    reg[k] = f < 3 ? P.coords[n * 3 + f] : (f == 3 ? P.visc[n] : P.vel[n * 3 + (f - 4)]);
  }
}
__host__ __device__ inline void mdot_fetch(const BoxParams &P, const StageIdx &S, int ez, double (&reg)[Cfg::NML])
{
  const long long ln = (long long)P.nx * P.ny * 12;
#pragma unroll
  for (int k = 0; k < Cfg::NML; ++k) {
    reg[k] = 0.0;
    if (S.mk[k] < 0 || ez < 0 || ez >= P.nz) continue;
    reg[k] = P.mdot[S.mk[k] + ln * ez];
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
    const double fac = -mu / det;
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

// per-class constants (class = (a,b,c): the column node differs from the row node along x / y / z).
// For the ip on the row's edge along axis pp, with local axes (t,u,v) = (pp, pp+1, pp+2):
//   tau d_t = c0[pp],  tau d_u = tau c1[pp],  tau d_v = tau c2[pp],  tau N_adv = tau cn[pp]
// where tau is the outward sign of the face for the row node.  Classes are data, not code: all
// warps run one instruction stream (the instruction cache does not hold per-class variants).
struct ClsK {
  double c0[3], c1[3], c2[3], cn[3];
  int a, b, c;
};
__host__ __device__ inline void cls_init(ClsK &Q, const P2K &K, int cls)
{
  Q.a = cls & 1;
  Q.b = (cls >> 1) & 1;
  Q.c = (cls >> 2) & 1;
  const int rho[3] = {Q.a, Q.b, Q.c};
#pragma unroll
  for (int pp = 0; pp < 3; ++pp) {
    const int rt = rho[pp], ru = rho[(pp + 1) % 3], rv = rho[(pp + 2) % 3];
    Q.c0[pp] = rt ? K.kt[ru + rv] : -K.kt[ru + rv];
    Q.c1[pp] = ru ? -K.kh[rv] : K.kh[rv];
    Q.c2[pp] = rv ? -K.kh[ru] : K.kh[ru];
    Q.cn[pp] = K.kn[ru + rv];
  }
}

__host__ __device__ inline void cls_store(const ClsK &Q, double *qk)
{
#pragma unroll
  for (int pp = 0; pp < 3; ++pp) {
    qk[pp] = Q.c0[pp];
    qk[3 + pp] = Q.c1[pp];
    qk[6 + pp] = Q.c2[pp];
    qk[9 + pp] = Q.cn[pp];
  }
}
__host__ __device__ inline void cls_fetch(ClsK &Q, const double *qk, int cls)
{
  Q.a = cls & 1;
  Q.b = (cls >> 1) & 1;
  Q.c = (cls >> 2) & 1;
#pragma unroll
  for (int pp = 0; pp < 3; ++pp) {
    Q.c0[pp] = qk[pp];
    Q.c1[pp] = qk[3 + pp];
    Q.c2[pp] = qk[6 + pp];
    Q.cn[pp] = qk[9 + pp];
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
__host__ __device__ __forceinline__ void rec_apply(const Rec &R, double c0, double c1, double c2, double cn,
                                                   double (&acc)[9], double &dm)
{
  const double *o = R.o;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double h = o[i * 3 + 0] * c0 + o[i * 3 + 1] * c1 + o[i * 3 + 2] * c2;
    acc[i * 3 + 0] += h * o[9];
    acc[i * 3 + 1] += h * o[10];
    acc[i * 3 + 2] += h * o[11];
  }
  dm += cn * o[12];
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

// finished block of one (sx,sy) group -> staging; rhs share.  H = 0: elements below the plane
// (class c = 1 fills the dz = -1 group, c = 0 the first half of dz = 0), H = 1: above (c = 1 ->
// dz = +1, c = 0 completes dz = 0 on top of the partial left in the staging row when `has_lower`).
__host__ __device__ __forceinline__ void p2_flush(T2 &N, const Meta &M, const ClsK &Q, int H, int sx, int sy, int z,
                                                  const double *ring, double *S0, double *S1, const double (&acc)[9],
                                                  double dm, double f23, bool has_lower)
{
  using C = Cfg;
  const int dx = Q.a ? (sx ? 1 : -1) : 0, dy = Q.b ? (sy ? 1 : -1) : 0;
  const int dz = Q.c ? (H ? 1 : -1) : 0;
  const int g = dz + 1;
  const unsigned slot = M.slot[N.node * 28 + g * 9 + (dy + 1) * 3 + (dx + 1)];
  if (slot == 0xFF) return; // no such neighbour
  const unsigned gp = M.gp[N.node * 4 + g];
  const unsigned gs = gp & 0xFF, pb = gp >> 8;
  const double *u = ring + ((z + dz + 3) % 3) * C::PLN + ((N.lx + 1 + dx) + C::RNX * (N.ly + 1 + dy)) * C::NF + 4;
  const double u0 = u[0], u1 = u[1], u2 = u[2];
  const double D = dm + (acc[0] + acc[4] + acc[8]);
  double *Sb = (Q.c ? S1 : S0) + N.node * C::SNODE + (slot - gs) * 3;
  const bool add = !Q.c && H && has_lower;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double v0 = acc[i * 3 + 0], v1 = acc[i * 3 + 1], v2 = acc[i * 3 + 2];
    if (i == 0) v0 += D;
    if (i == 1) v1 += D;
    if (i == 2) v2 += D;
    N.rr[i] -= v0 * u0 + v1 * u1 + v2 * u2;
    N.rr[i] += f23 * (acc[0 + i] * u0 + acc[3 + i] * u1 + acc[6 + i] * u2);
    double *d = Sb + i * C::SROWP + ((pb >> i) & 1);
    if (add) {
      v0 += d[0];
      v1 += d[1];
      v2 += d[2];
    }
    d[0] = v0;
    d[1] = v1;
    d[2] = v2;
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
// the outward signs tau_x, tau_y are compile-time, tau_z comes in through z1, z2, zn
template <int SX, int SY, int K>
__host__ __device__ __forceinline__ void rec_use(const Rec &R, const ClsK &Q, double z1, double z2, double zn,
                                                 double (&acc)[9], double &dm)
{
  if (K == 0) rec_apply(R, Q.c0[0], SX ? Q.c1[0] : -Q.c1[0], SX ? Q.c2[0] : -Q.c2[0], SX ? Q.cn[0] : -Q.cn[0], acc, dm);
  else if (K == 1) rec_apply(R, Q.c0[1], SY ? Q.c1[1] : -Q.c1[1], SY ? Q.c2[1] : -Q.c2[1], SY ? Q.cn[1] : -Q.cn[1], acc, dm);
  else rec_apply(R, Q.c0[2], z1, z2, zn, acc, dm);
}

// one z half (H = 0 below / 1 above the plane) of the node's octants.  Quadrants that fold into the
// same column node are visited back to back so that one 3x3 accumulator is live; the flush points
// are a function of the (warp-uniform) class: (a,b) = (1,1): every quadrant is its own column;
// (0,1): pairs differing in sx fold; (1,0): pairs differing in sy fold (SWAP: visit sy fastest);
// (0,0): all four fold.  No test for quadrants / nodes outside the domain: their records are zero
// (never written by phase 1 after the prologue cleared F / ZE) and their flushes see slot 0xFF.
// The 12 records are double-buffered in registers: record k+1 is loaded before record k is used.
template <bool SWAP>
__host__ __device__ inline void p2_half_t(T2 &N, const Meta &M, const ClsK &Q, double f23, int H, int z, const double *F,
                                          const double *ZE, const double *ring, double *S0, double *S1, bool has_lower)
{
  using C = Cfg;
  const double *eF = F + (N.lx + C::ELX * N.ly) * C::ES;
  const double *eZ = ZE + (N.lx + C::ELX * N.ly) * C::ES;
  const unsigned fm = (Q.a & Q.b) ? 0xFu : ((Q.a | Q.b) ? 0xAu : 0x8u);
  // z-edge ips: tau_z = +1 for the elements above the plane
  const double z1 = flip_if(Q.c1[2], !H), z2 = flip_if(Q.c2[2], !H), zn = flip_if(Q.cn[2], !H);
  double acc[9], dm = 0.0;
#pragma unroll
  for (int t = 0; t < 9; ++t) acc[t] = 0.0;
  constexpr int X1 = SWAP ? 0 : 1, Y1 = SWAP ? 1 : 0, X2 = SWAP ? 1 : 0, Y2 = SWAP ? 0 : 1;
  Rec RA, RB;
  rec_load(RA, rec_addr<0, 0, 0>(eF, eZ));
#define NALU_P2_QUAD(J, SX, SY, NX_, NY_, LAST)                                                   \
  rec_load(RB, rec_addr<SX, SY, 1>(eF, eZ));                                                      \
  rec_use<SX, SY, 0>(RA, Q, z1, z2, zn, acc, dm);                                                 \
  rec_load(RA, rec_addr<SX, SY, 2>(eF, eZ));                                                      \
  rec_use<SX, SY, 1>(RB, Q, z1, z2, zn, acc, dm);                                                 \
  if (!LAST) rec_load(RB, rec_addr<NX_, NY_, 0>(eF, eZ));                                         \
  rec_use<SX, SY, 2>(RA, Q, z1, z2, zn, acc, dm);                                                 \
  if (fm & (1u << J)) {                                                                           \
    p2_flush(N, M, Q, H, SX, SY, z, ring, S0, S1, acc, dm, f23, has_lower);                       \
    _Pragma("unroll") for (int t = 0; t < 9; ++t) acc[t] = 0.0;                                   \
    dm = 0.0;                                                                                     \
  }                                                                                               \
  if (!LAST) RA = RB;
  NALU_P2_QUAD(0, 0, 0, X1, Y1, false)
  NALU_P2_QUAD(1, X1, Y1, X2, Y2, false)
  NALU_P2_QUAD(2, X2, Y2, 1, 1, false)
  NALU_P2_QUAD(3, 1, 1, 1, 1, true)
#undef NALU_P2_QUAD
}

__host__ __device__ inline void p2_half(T2 &N, const Meta &M, const ClsK &Q, double f23, int H, int z, const double *F,
                                        const double *ZE, const double *ring, double *S0, double *S1, bool has_lower)
{
  // nodes outside the domain run the same code on zero records; their flushes see slot 0xFF
  if (Q.a && !Q.b) p2_half_t<true>(N, M, Q, f23, H, z, F, ZE, ring, S0, S1, has_lower);
  else p2_half_t<false>(N, M, Q, f23, H, z, F, ZE, ring, S0, S1, has_lower);
}

// ---- node metadata of plane z: thread (node, g) reads the 9 slots of dz group g ------------------
__host__ __device__ inline void meta_load(const BoxParams &P, const Meta &M, int node, int g, int x0, int y0, int z)
{
  using C = Cfg;
  const int x = x0 + node % C::TBX, y = y0 + node / C::TBX;
  unsigned mn = 255, mx = 0, cnt = 0;
  long long rowbase = 0;
  int nb = 0, row = 0;
  if (x < P.NX && y < P.NY && z < P.NZ) {
    const long long n = x + (long long)P.NX * (y + (long long)P.NY * z);
    const unsigned char *sl = P.slot27 + n * 27 + g * 9;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const unsigned s = sl[k];
      M.slot[node * 28 + g * 9 + k] = (unsigned char)s;
      if (s != 0xFF) {
        mn = s < mn ? s : mn;
        mx = s > mx ? s : mx;
        ++cnt;
      }
    }
    if (cnt && mx - mn + 1 != cnt) *P.err = 2; // the columns of a dz group must be contiguous in the row
    row = P.node_row[n];
    const long long b0 = P.browptr[row];
    nb = (int)(P.browptr[row + 1] - b0);
    rowbase = 9 * b0;
  } else {
#pragma unroll
    for (int k = 0; k < 9; ++k) M.slot[node * 28 + g * 9 + k] = 0xFF;
  }
  if (!cnt) mn = 0;
  unsigned pb = 0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const long long G = rowbase + ((long long)i * nb + mn) * 3;
    pb |= (unsigned)((G ^ node) & 1) << i;
  }
  M.gp[node * 4 + g] = (unsigned short)(mn | (pb << 8));
  M.cnt[node * 4 + g] = (unsigned char)cnt;
  if (g == 0) {
    M.rowbase[node] = rowbase;
    M.nb[node] = nb;
    M.row[node] = row;
  }
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
  const unsigned gp = M.gp[node * 4 + g];
  r.L = 3 * M.cnt[node * 4 + g];
  r.G = M.rowbase[node] + ((long long)i * M.nb[node] + (gp & 0xFF)) * 3;
  r.soff = node * C::SNODE + i * C::SROWP + ((gp >> (8 + i)) & 1);
  return r;
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
