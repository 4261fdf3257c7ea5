// box5_tile.cuh -- row-thread tiled, write-once MomentumAdvDiff assembly for structured Hex8 blocks.
//
// Contract: what AssembleElemSolverAlgorithm::execute + MomentumAdvDiffElemKernel::execute (+ the lumped
// MomentumMassElemKernel when fused) + TpetraLinearSystem::sumInto produce (src/AssembleElemSolverAlgorithm.C:74-100,
// src/kernel/MomentumAdvDiffElemKernel.C:70-186, src/kernel/MomentumMassElemKernel.C:98-171,
// src/TpetraLinearSystem.C:1457-1522), every CSR value written exactly once, already summed.
//
// Work decomposition (what the round-1 profiles asked for: fewer instructions per node, no CTA barrier between
// accumulate and output, coalesced stores that drain while other warps compute):
//   * a CTA (128 threads) owns an 8 x 4 column of nodes and sweeps z; a warp owns one x line of 8 nodes, a node is
//     served by a QUAD of lanes: lane i = 0..2 keeps dof row i of the node's blocks in registers, lane 3 keeps the
//     diagonal term D_c = sum tau (N_c mdot + h_c . A) of every block (MomentumAdvDiffElemKernel.C:160-176);
//     all four run ONE instruction stream: lane 3 reads the row B' = G~^T A -+ mdot/2 and the "area vector" (1,*,*).
//     (A variant with two quads per node, one per column plane, 27 accumulators and 16 warps per SM was measured in
//     round 2: both quads load the same records, the shared-memory pipe -- already 70 % busy -- saturates, no gain);
//   * per (integration point, column node) a lane spends 6 FMAs: h = G~_i . (tau dN_c/dxi) (3), acc[c][.] += h A (3);
//     all shape-function derivatives are compile-time signs times 7 run-time magnitudes, quadrants / points / columns
//     are unrolled, the lower / upper z half differ in a run-time sign only (one copy of the code);
//   * lane i stores its accumulators ROTATED (local index k' <-> dof column (i + k') mod 3, the area vector is read
//     through three rotated pointers), so that the diagonal entry is register 0 for every lane and D_c is added
//     with 9 adds after one quad broadcast;
//   * rhs_i = -(assembled row) . u + 2/3 includeDivU sum_c sum_j acc_c[j][i] u_cj (the div(u) stress expressed through
//     the accumulated viscous blocks, as in box2_tile.cuh), the transposed terms travel by two quad shuffles;
//   * output: the 24 finished 27-double runs of a warp (8 nodes x 3 dof rows of one dz group) are transposed through
//     a warp-private 5 KB scratch and leave as 24 store instructions of 216 contiguous bytes, addresses =
//     one warp-uniform base + compile-time offsets on interior x lines; only __syncwarp, so the warps of a CTA drift
//     apart and the stores of one overlap the FMAs of the others;
//   * phase 1 (thread = (element, 4 integration points): bottom-face + z-edge units before the upper half, top-face
//     units after it): one Walsh-Hadamard butterfly of the 8 nodes per unit, Jacobian = 5 vector FMAs per point,
//     area vector one cross product; 17 doubles per point in shared memory, 8 points per element resident.
//
// Every phase is a plain function of the thread index so that tests/cpu_harness/box5_host_check.cpp can execute the
// kernel thread by thread on the CPU (no GPU in the development container).
#pragma once
#include "box_tile.cuh"

namespace nalu_b200 {
namespace box5 {

using box::BoxParams;
using box::butterfly8;
using box::ip_of;
using box::node_of_bits;

struct Cfg {
  static constexpr int TBX = 8, TBY = 4, NNODE = TBX * TBY, NT = 4 * NNODE;
  static constexpr int ELX = TBX + 1, ELY = TBY + 1, NEL = ELX * ELY;
  // record of one integration point: G~[3][3] (rows = dof, columns = local axes t,u,v), B_t - mdot/2, B_u, B_v,
  // B_t + mdot/2, A[3] (along +t), 1.0
  // Layout: field-major, [8 points: 4 face, 4 z-edge][17 fields][FS elements]: the threads of phase 1 (thread = element)
  // write consecutive doubles; a 64-bit shared load is served half a warp at a time: the 16 lanes (4 consecutive
  // elements x lanes i = 0..3 reading fields 3i + const) must hit 16 different 8-byte bank slots, i.e. 3 FS = 4 mod 16
  static constexpr int REC = 17;
  static constexpr int FS = 60;           // field stride >= NEL, = 12 mod 16
  static constexpr int PS = REC * FS;     // point stride
  static constexpr int RECSZ = 8 * PS;
  // staged node plane: geometry part [RN][6] = x y z mu - - (48-byte stride: the 16-byte loads of 8 consecutive
  // elements cover all banks), velocity part [RN][3] (the quads of a warp read 24 consecutive doubles)
  static constexpr int RNX = TBX + 2, RNY = TBY + 2, RN = RNX * RNY, GEO = 6, VEL = 3, OFF_VEL = RN * GEO, PLN = RN * (GEO + VEL);
  static constexpr int SCRW = 24 * 27; // doubles of scratch per warp: 24 runs of 27
  static constexpr int NWARP = NT / 32;
  static constexpr int MDB = NEL * 12;
  // node metadata of one plane = the graph's row descriptors (graph.h, node_desc) of the tile's 32 nodes: [3][32] words
  static constexpr int META = 3 * NNODE;
  // shared-memory map (doubles); SCV: the 8 sub-control-volume volumes of every element of the layer, as two partial sums
  // (fused lumped mass)
  static constexpr int OFF_REC = 0, OFF_SCR = OFF_REC + RECSZ, OFF_RING = OFF_SCR + NWARP * SCRW, OFF_MD = OFF_RING + 3 * PLN,
                       OFF_META = OFF_MD + MDB, OFF_SCV = OFF_META + 2 * META;
  static constexpr int SMEM_DOUBLES = OFF_SCV + NEL * 16;
};
static_assert(Cfg::FS % 16 == 12 && Cfg::FS >= Cfg::NEL, "field stride must spread the quads of a warp over the banks");
static_assert((Cfg::OFF_RING % 2) == 0 && (Cfg::OFF_MD % 2) == 0, "16-byte aligned staging areas");

struct Meta {
  const unsigned long long *w; // [3][NNODE]
  __host__ __device__ __forceinline__ long long rowbase(int node) const { return (long long)w[node]; }
  __host__ __device__ __forceinline__ int row(int node) const { return (int)(unsigned)w[Cfg::NNODE + node]; }
  __host__ __device__ __forceinline__ int nb(int node) const { return (int)((w[Cfg::NNODE + node] >> 32) & 0xFF); }
  __host__ __device__ __forceinline__ bool natural(int node) const { return (w[Cfg::NNODE + node] >> 40) & 1; }
  __host__ __device__ __forceinline__ bool valid(int node) const { return (w[Cfg::NNODE + node] >> 41) & 1; }
  __host__ __device__ __forceinline__ int s0(int node, int g) const { return (int)((w[2 * Cfg::NNODE + node] >> (8 * g)) & 0xFF); }
  __host__ __device__ __forceinline__ int cnt(int node, int g) const { return (int)((w[2 * Cfg::NNODE + node] >> (24 + 8 * g)) & 0xFF); }
  __host__ __device__ __forceinline__ unsigned order(int node) const { return (unsigned)(w[2 * Cfg::NNODE + node] & 0xFFFFFF); }
};
__host__ __device__ inline Meta meta_at(double *sm, int buf)
{
  Meta M;
  M.w = reinterpret_cast<const unsigned long long *>(sm + Cfg::OFF_META + buf * Cfg::META);
  return M;
}

// ---- staging: global -> shared with cp.async (no registers); entries outside the mesh read zero -------------------
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
__host__ __device__ __forceinline__ void stage_wait()
{
#ifdef __CUDA_ARCH__
  asm volatile("cp.async.wait_all;" ::: "memory");
#endif
}
// node plane zp of the tile's neighbourhood; `t` = 0..RN-1 (one node per calling thread)
__host__ __device__ inline void stage_plane(const BoxParams &P, double *ring_plane, int t, int x0, int y0, int zp)
{
  using C = Cfg;
  if (t >= C::RN) return;
  const long long pn = (long long)P.NX * P.NY;
  const int x = x0 - 1 + t % C::RNX, y = y0 - 1 + t / C::RNX;
  const bool valid = zp >= 0 && zp < P.NZ && x >= 0 && y >= 0 && x < P.NX && y < P.NY;
  const long long n = valid ? x + (long long)P.NX * y + pn * zp : 0;
  double *d = ring_plane + t * C::GEO, *v = ring_plane + C::OFF_VEL + t * C::VEL;
  const double *xc = P.coords + n * 3, *u = P.vel + n * 3;
  copy8_or_zero(d + 0, xc + 0, valid);
  copy8_or_zero(d + 1, xc + 1, valid);
  copy8_or_zero(d + 2, xc + 2, valid);
  copy8_or_zero(d + 3, P.visc + n, valid);
  copy8_or_zero(v + 0, u + 0, valid);
  copy8_or_zero(v + 1, u + 1, valid);
  copy8_or_zero(v + 2, u + 2, valid);
}
// mass flow rates of element layer ez: [NEL][12] as 270 16-byte copies; thread t makes copies t, t + NT, t + 2 NT
__host__ __device__ inline void stage_mdot(const BoxParams &P, double *mdb, int t, int x0, int y0, int ez)
{
  using C = Cfg;
  const bool zin = ez >= 0 && ez < P.nz;
  const double *layer = P.mdot + (zin ? (long long)P.nx * P.ny * ez * 12 : 0);
  const bool wide = (((unsigned long long)P.mdot) & 15) == 0; // uniform
#pragma unroll
  for (int k = 0; k < (C::NEL * 6 + C::NT - 1) / C::NT; ++k) {
    const int idx = t + k * C::NT;
    if (idx >= C::NEL * 6) continue;
    const int el = idx / 6, pr = idx - el * 6;
    const int ex = x0 - 1 + el % C::ELX, ey = y0 - 1 + el / C::ELX;
    const bool valid = zin && ex >= 0 && ey >= 0 && ex < P.nx && ey < P.ny;
    const double *src = layer + (valid ? (ex + P.nx * ey) * 12 + 2 * pr : 0);
    double *dst = mdb + el * 12 + 2 * pr;
    if (wide) copy16_or_zero(dst, src, valid);
    else {
      copy8_or_zero(dst, src, valid);
      copy8_or_zero(dst + 1, src + 1, valid);
    }
  }
}

// ---- phase 1 ---------------------------------------------------------------------------------------------------------
// Walsh-Hadamard coefficients of the element's coordinates (relative to its first node: only difference coefficients
// are used, and differences of partial sums of large coordinates would cost |x|/h ulps) and viscosity; mask bits:
// 1 = x, 2 = y, 4 = z (the element's own axes = the lattice axes).
struct Coef {
  double X[3][8], M[8];
};
struct D2 {
  double x, y;
};
__host__ __device__ __forceinline__ void ld4(const double *p, double &a, double &b, double &c, double &d)
{
#ifdef __CUDA_ARCH__
  const double2 v0 = reinterpret_cast<const double2 *>(p)[0], v1 = reinterpret_cast<const double2 *>(p)[1];
  a = v0.x; b = v0.y; c = v1.x; d = v1.y;
#else
  a = p[0]; b = p[1]; c = p[2]; d = p[3];
#endif
}
__host__ __device__ __forceinline__ void coef_load(Coef &C, const double *pl0, const double *pl1, int elx, int ely)
{
  using K = Cfg;
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const double *nd = ((m & 4) ? pl1 : pl0) + ((elx + (m & 1)) + K::RNX * (ely + ((m >> 1) & 1))) * K::GEO;
    ld4(nd, C.X[0][m], C.X[1][m], C.X[2][m], C.M[m]);
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double o = C.X[i][0];
#pragma unroll
    for (int m = 0; m < 8; ++m) C.X[i][m] -= o;
  }
  butterfly8(C.X[0]);
  butterfly8(C.X[1]);
  butterfly8(C.X[2]);
  butterfly8(C.M);
}
// base area vector of direction T: C_u x C_v / 64
template <int T>
__host__ __device__ __forceinline__ void area_base(const Coef &C, double (&b)[3])
{
  constexpr int mu = 1 << ((T + 1) % 3), mv = 1 << ((T + 2) % 3);
  b[0] = (C.X[1][mu] * C.X[2][mv] - C.X[2][mu] * C.X[1][mv]) * (1.0 / 64.0);
  b[1] = (C.X[2][mu] * C.X[0][mv] - C.X[0][mu] * C.X[2][mv]) * (1.0 / 64.0);
  b[2] = (C.X[0][mu] * C.X[1][mv] - C.X[1][mu] * C.X[0][mv]) * (1.0 / 64.0);
}
// one integration point on an edge along axis T (local axes t,u,v = T, T+1, T+2), on side (sgu, sgv) of the element;
// md = mass flow rate counted along +t.  Jacobian / cofactors / determinant as generic_grad_op
// (MasterElementFunctions.h:51-108) in the coefficient basis; area vector = the reference's 27-point subdivision +
// 4-triangle fan (Hex8GeometryFunctions.h:30-76,263-321) as the vector area of a closed polygon.
template <int T, int F = Cfg::FS>
__host__ __device__ __forceinline__ void p1_point(const Coef &C, const double (&base)[3], double sgu, double sgv, double q, double md,
                                                  double *o, int *err)
{
  constexpr int mt = 1 << T, mu = 1 << ((T + 1) % 3), mv = 1 << ((T + 2) % 3);
  const double c1 = 0.5 * q * sgu, c2 = 0.5 * q * sgv, c12 = q * q * sgu * sgv;
  double J[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    J[i][0] = 0.25 * C.X[i][mt] + c1 * C.X[i][mt | mu] + c2 * C.X[i][mt | mv] + c12 * C.X[i][7];
    J[i][1] = 0.25 * C.X[i][mu] + c2 * C.X[i][mu | mv];
    J[i][2] = 0.25 * C.X[i][mv] + c1 * C.X[i][mu | mv];
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
  if (!(det > 0.0)) *err = 1; // MasterElementFunctions.h:91-94 asserts this in debug builds
  // mu at the (never shifted) ip location, MomentumAdvDiffElemKernel.C:107-110
  const double mu_ip = 0.125 * C.M[0] + 0.0625 * (sgu * C.M[mu] + sgv * C.M[mv]) + 0.03125 * sgu * sgv * C.M[mu | mv];
  const double fac = -mu_ip * (1.0 / det);
  double G[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int k = 0; k < 3; ++k) G[i][k] = fac * adj[i][k];
  double W[3], A[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) W[i] = sgv * C.X[i][mv] - sgu * C.X[i][mu];
  A[0] = base[0] + (C.X[1][mu | mv] * W[2] - C.X[2][mu | mv] * W[1]) * (1.0 / 128.0);
  A[1] = base[1] + (C.X[2][mu | mv] * W[0] - C.X[0][mu | mv] * W[2]) * (1.0 / 128.0);
  A[2] = base[2] + (C.X[0][mu | mv] * W[1] - C.X[1][mu | mv] * W[0]) * (1.0 / 128.0);
  double B[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) B[k] = G[0][k] * A[0] + G[1][k] * A[1] + G[2][k] * A[2];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int k = 0; k < 3; ++k) o[(i * 3 + k) * F] = G[i][k];
  o[9 * F] = B[0] - 0.5 * md;
  o[10 * F] = B[1];
  o[11 * F] = B[2];
  o[12 * F] = B[0] + 0.5 * md;
  o[13 * F] = A[0];
  o[14 * F] = A[1];
  o[15 * F] = A[2];
  o[16 * F] = 1.0;
}
// Volumes of the 8 sub-control volumes of an element (HexSCV::determinant, Hex8CVFEM.C:149-178; hex_volume_grandy,
// Hex8GeometryFunctions.h:78-163), from the coefficient vectors.  The sub-hex of corner s = (s1,s2,s3) is the image of the
// parametric octant xi_k = s_k/4 + eta_k/2, eta in [-1/2,1/2]^3; in eta the map is again trilinear,
//   dx/d eta_k = (B_k + B_kl eta_l + B_km eta_m + C_klm eta_l eta_m) / 8,
//   B_k = C_k + s_l C_kl/2 + s_m C_km/2 + s_l s_m C_klm/4,   B_kl = C_kl + s_m C_klm/2,
// and the integral of its Jacobian determinant over the cube (what Grandy's 24-triangle formula evaluates exactly for
// a hex with bilinear faces) keeps only the even monomials:
//   V(s) = [ det(B1,B2,B3) + ( det(B13,B23,B3) + det(B12,B2,B23) + det(B1,B12,B13) ) / 12 ] / 512.
// As a function of the corner, V(s) = sum_mask w_mask prod_{k in mask} s_k (s_k^2 = 1), and expanding the determinants
// leaves 29 triple products [abc] = C_a . (C_b x C_c) of the coefficient vectors (indices = axis masks 1 x, 2 y, 4 z):
//   512 w0 =  [124] + ([135] - [236] + [456])/3          512 w4 = -[146]/2 - [157]/6 + [245]/2 + [267]/6
//   512 w1 = ([125] + [134])/2 + ([457] - [237])/6       512 w5 = (-[147] - [156] + [345])/4 + [367]/12
//   512 w2 = ([126] - [234])/2 + ([137] - [467])/6       512 w6 = ( [247] - [256] - [346])/4 - [357]/12
//   512 w3 = ([127] + [136] - [235])/4 - [567]/12        512 w7 = -[356]/4
// -- a third of the arithmetic of eight separate corner formulas.  The triple products are grouped by a shared cross
// product and split into two parts of about equal cost: PART 0 (cross products 12, 13, 23, 45; the face units of phase
// 1a) and PART 1 (67, 56, 47, 46, 57; the z units).  Each part synthesises its share of the 8 volumes
// (out[part * 8 + Hex8 node of the corner]); the consumer adds the two.
__host__ __device__ __forceinline__ void cross_c(const Coef &C, int a, int b, double (&p)[3])
{
  p[0] = C.X[1][a] * C.X[2][b] - C.X[2][a] * C.X[1][b];
  p[1] = C.X[2][a] * C.X[0][b] - C.X[0][a] * C.X[2][b];
  p[2] = C.X[0][a] * C.X[1][b] - C.X[1][a] * C.X[0][b];
}
__host__ __device__ __forceinline__ double dot_c(const Coef &C, const double (&p)[3], int c)
{
  return p[0] * C.X[0][c] + p[1] * C.X[1][c] + p[2] * C.X[2][c];
}
template <int PART>
__host__ __device__ __forceinline__ void scv_volumes_part(const Coef &C, double *out)
{
  constexpr double h = 0.5, q = 0.25, t = 1.0 / 3.0, s = 1.0 / 6.0, w12 = 1.0 / 12.0;
  double w[8], p[3];
  if (PART == 0) {
    cross_c(C, 1, 2, p);
    w[0] = dot_c(C, p, 4); w[1] = h * dot_c(C, p, 5); w[2] = h * dot_c(C, p, 6); w[3] = q * dot_c(C, p, 7);
    cross_c(C, 1, 3, p);
    w[0] += t * dot_c(C, p, 5); w[1] += h * dot_c(C, p, 4); w[2] += s * dot_c(C, p, 7); w[3] += q * dot_c(C, p, 6);
    cross_c(C, 2, 3, p);
    w[0] -= t * dot_c(C, p, 6); w[1] -= s * dot_c(C, p, 7); w[2] -= h * dot_c(C, p, 4); w[3] -= q * dot_c(C, p, 5);
    cross_c(C, 4, 5, p);
    w[0] += t * dot_c(C, p, 6); w[1] += s * dot_c(C, p, 7); w[4] = h * dot_c(C, p, 2); w[5] = q * dot_c(C, p, 3);
    w[6] = 0.0; w[7] = 0.0;
  } else {
    cross_c(C, 6, 7, p);
    w[2] = -s * dot_c(C, p, 4); w[3] = -w12 * dot_c(C, p, 5); w[4] = s * dot_c(C, p, 2); w[5] = w12 * dot_c(C, p, 3);
    cross_c(C, 5, 6, p);
    w[5] -= q * dot_c(C, p, 1); w[6] = -q * dot_c(C, p, 2); w[7] = -q * dot_c(C, p, 3);
    cross_c(C, 4, 7, p);
    w[5] -= q * dot_c(C, p, 1); w[6] += q * dot_c(C, p, 2);
    cross_c(C, 4, 6, p);
    w[4] -= h * dot_c(C, p, 1); w[6] -= q * dot_c(C, p, 3);
    cross_c(C, 5, 7, p);
    w[4] -= s * dot_c(C, p, 1); w[6] -= w12 * dot_c(C, p, 3);
    w[0] = 0.0; w[1] = 0.0;
  }
  // synthesis: V[bits] = sum_mask w_mask prod s_k, s_k = +1 where bit k of `bits` is set
#pragma unroll
  for (int b = 1; b < 8; b <<= 1)
#pragma unroll
    for (int m = 0; m < 8; ++m)
      if (!(m & b)) {
        const double lo = w[m], hi = w[m | b];
        w[m] = lo - hi;
        w[m | b] = lo + hi;
      }
#pragma unroll
  for (int m = 0; m < 8; ++m) out[PART * 8 + node_of_bits(m & 1, (m >> 1) & 1, m >> 2)] = w[m] * (1.0 / 512.0);
}

// face unit: the 4 points of the bottom (top = 0) or top (top = 1) face of element `el` -> record slots 0..3
// (slot = ip & 3: Hex8CVFEM.C:264-276: ip 0: x edge at y-, 1: y edge at x+, 2: x edge at y+ (L -> R runs along -x),
// 3: y edge at x-)
__host__ __device__ inline void p1_face_unit(const BoxParams &P, int el, int top, int x0, int y0, const double *pl0, const double *pl1,
                                             const double *mdb, double *rec, double *scv)
{
  using K = Cfg;
  const int elx = el % K::ELX, ely = el / K::ELX;
  const int ex = x0 - 1 + elx, ey = y0 - 1 + ely;
  if (ex < 0 || ey < 0 || ex >= P.nx || ey >= P.ny) return;
  const double q = (P.options & NALU_OPT_SHIFTED_GRAD_OP) ? 0.5 : 0.25;
  Coef C;
  coef_load(C, pl0, pl1, elx, ely);
  const double sz = top ? 1.0 : -1.0;
  const double *md = mdb + el * 12 + (top ? 4 : 0);
  double *o = rec + el;
  if (P.fuse_mass) scv_volumes_part<0>(C, scv + el * 16);
  double bx[3], by[3];
  area_base<0>(C, bx);
  area_base<1>(C, by);
  // T = 0: (u,v) = (y,z); T = 1: (u,v) = (z,x).  A loop, not four inlined copies: phase 1 shares 128 registers per
  // thread with the accumulate sweep's budget, and 16 resident warps hide the latency an unrolled version would
#pragma unroll 1
  for (int s = 0; s < 2; ++s) {
    const double sg = s ? 1.0 : -1.0;
    p1_point<0>(C, bx, sg, sz, q, s ? -md[2] : md[0], o + (2 * s) * K::PS, P.err);  // ip 0 (y-), ip 2 (y+, L -> R along -x)
    p1_point<1>(C, by, sz, -sg, q, s ? md[3] : md[1], o + (1 + 2 * s) * K::PS, P.err); // ip 1 (x+), ip 3 (x-)
  }
}
// half of a face unit: the 2 points of edge direction T (0: x edges = ips 0, 2; 1: y edges = ips 1, 3) of the bottom /
// top face.  Phase 1b spreads these over four warps (one butterfly per 2 points, but half the critical path)
template <int T>
__host__ __device__ inline void p1_face_half(const BoxParams &P, int el, int top, int x0, int y0, const double *pl0, const double *pl1,
                                             const double *mdb, double *rec)
{
  using K = Cfg;
  const int elx = el % K::ELX, ely = el / K::ELX;
  const int ex = x0 - 1 + elx, ey = y0 - 1 + ely;
  if (ex < 0 || ey < 0 || ex >= P.nx || ey >= P.ny) return;
  const double q = (P.options & NALU_OPT_SHIFTED_GRAD_OP) ? 0.5 : 0.25;
  Coef C;
  coef_load(C, pl0, pl1, elx, ely);
  const double sz = top ? 1.0 : -1.0;
  const double *md = mdb + el * 12 + (top ? 4 : 0);
  double *o = rec + el;
  double b[3];
  area_base<T>(C, b);
#pragma unroll 1
  for (int s = 0; s < 2; ++s) {
    const double sg = s ? 1.0 : -1.0;
    if (T == 0) p1_point<0>(C, b, sg, sz, q, s ? -md[2] : md[0], o + (2 * s) * K::PS, P.err);   // ip 0 (y-), ip 2 (y+, L -> R along -x)
    else p1_point<1>(C, b, sz, -sg, q, s ? md[3] : md[1], o + (1 + 2 * s) * K::PS, P.err);     // ip 1 (x+), ip 3 (x-)
  }
}
// z unit: the 4 z-edge points (ip 8 + n above bottom node n) -> record slots 4..7
__host__ __device__ inline void p1_z_unit(const BoxParams &P, int el, int x0, int y0, const double *pl0, const double *pl1,
                                          const double *mdb, double *rec, double *scv)
{
  using K = Cfg;
  const int elx = el % K::ELX, ely = el / K::ELX;
  const int ex = x0 - 1 + elx, ey = y0 - 1 + ely;
  if (ex < 0 || ey < 0 || ex >= P.nx || ey >= P.ny) return;
  const double q = (P.options & NALU_OPT_SHIFTED_GRAD_OP) ? 0.5 : 0.25;
  Coef C;
  coef_load(C, pl0, pl1, elx, ely);
  const double *md = mdb + el * 12 + 8;
  double *o = rec + el + 4 * K::PS;
  if (P.fuse_mass) scv_volumes_part<1>(C, scv + el * 16);
  double bz[3];
  area_base<2>(C, bz);
  // T = 2: (u,v) = (x,y); bottom nodes 0..3 = (x-,y-), (x+,y-), (x+,y+), (x-,y+)
#pragma unroll 1
  for (int n = 0; n < 4; ++n)
    p1_point<2>(C, bz, (n == 1 || n == 2) ? 1.0 : -1.0, n >= 2 ? 1.0 : -1.0, q, md[n], o + n * K::PS, P.err);
}

// ---- phase 2 -----------------------------------------------------------------------------------------------------------
// Run-time magnitudes of the shape-function derivatives.  In the frame of the row node an integration point sits at 0
// along its edge and at distance q towards the row node on the two transverse axes; phi = 1/2 + q for a column on the
// row's side of a transverse axis, 1/2 - q otherwise (Hex8CVFEM.C:24-70,318-345).
// kt[n] = phi_u phi_v for n = 0, 1, 2 transverse axes on which the column differs from the row; kh = 1/2 phi; f23 = 2/3
// includeDivU.  They live in BoxParams (constant bank), filled by the launcher.
__host__ __device__ inline void p2k_init(BoxParams &P)
{
  const double q = (P.options & NALU_OPT_SHIFTED_GRAD_OP) ? 0.5 : 0.25;
  const double p0 = 0.5 + q, p1 = 0.5 - q;
  P.kt[0] = p0 * p0; P.kt[1] = p0 * p1; P.kt[2] = p1 * p1;
  P.kh[0] = 0.5 * p0; P.kh[1] = 0.5 * p1;
  P.f23 = 2.0 / 3.0 * P.include_divu;
}
// the advection shape function must sit at the same point as the gradient operator (N_c = |dN_c/dxi_t| / 2 is what
// lets lane 3 fold N_c mdot into its row): both options on or both off
__host__ __device__ inline bool options_supported(int options)
{
  return ((options & NALU_OPT_SHIFTED_GRAD_OP) != 0) == ((options & NALU_OPT_SKEW_SYMMETRIC) != 0);
}

// who a thread is: a pure function of its index, rebuilt at the head of every phase
struct Thr {
  int lx, ly, i;       // node of the tile, lane role (0..2 dof row, 3 diagonal term)
  int node;            // lx + 8 ly
  int rot[3];          // (i + k') mod 3 for i < 3; lane 3: 0, 1, 2
  int active;          // node inside the domain
  // element (lx, ly) (= quadrant (0,0)) record base + lane offsets, in doubles from the start of shared memory
  int oG, oGp, oA[3];
};
__host__ __device__ __forceinline__ Thr thr_make(const BoxParams &P, int tid, int x0, int y0)
{
  using C = Cfg;
  Thr N;
  const int lane = tid & 31;
  N.ly = tid >> 5;
  N.lx = lane >> 2;
  N.i = lane & 3;
  N.node = N.lx + C::TBX * N.ly;
  const int r = N.i < 3 ? N.i : 0;
  N.rot[0] = r;
  N.rot[1] = r == 2 ? 0 : r + 1;
  N.rot[2] = r == 0 ? 2 : r - 1;
  N.active = (x0 + N.lx < P.NX) && (y0 + N.ly < P.NY);
  const int rec = C::OFF_REC + (N.lx + C::ELX * N.ly);
  N.oG = rec + 3 * N.i * C::FS;                           // lane 3: field 9 = B_t - mdot/2
  N.oGp = N.i < 3 ? N.oG : rec + 12 * C::FS;
  N.oA[0] = rec + (N.i < 3 ? 13 + N.rot[0] : 16) * C::FS; // lane 3: 1.0, acc[c][0] accumulates D_c
  N.oA[1] = rec + (N.i < 3 ? 13 + N.rot[1] : 16) * C::FS; // lane 3: anything finite on a bank slot of its own
  N.oA[2] = rec + (N.i < 3 ? 13 + N.rot[2] : 16) * C::FS;
  return N;
}

// record slot of point K (0 x edge, 1 y edge, 2 z edge) of the row node in quadrant (SX,SY: the element lies on the
// +x / +y side of the node): the node is corner (!SX, !SY, .) of the element
template <int SX, int SY, int K>
__host__ __device__ constexpr int rec_off()
{
  return (SX + Cfg::ELX * SY) + (K == 0 ? (SY ? 0 : 2) : (K == 1 ? (SX ? 3 : 1) : 4 + (SY ? 0 : 2) + (SX ^ SY))) * Cfg::PS;
}

// one integration point into the 8 column nodes of its element.  sH = rho_z = +1 for the lower half (the row node is
// a top corner of the element), -1 for the upper half.
//   e_t = (rt ? + : -) kt[ru+rv]
//   e_u = -rho_t (ru ? -rho_u : rho_u) kh[rv],  e_v = -rho_t (rv ? -rho_v : rho_v) kh[ru]     (rho = side of the row node)
//   Gt: B_t + sigma_t mdot/2 for lane 3 -> the `far` copy for columns with sigma_t = +1
template <int SX, int SY, int K>
__host__ __device__ __forceinline__ void sweep_point(const double *pT0, const double *pT1, const double *pG, const double *pA0,
                                                     const double *pA1, const double *pA2, const BoxParams &Kc, const double (&khz)[2],
                                                     double (&acc1)[9][3], double (&acc2)[9][3])
{
  constexpr int off = rec_off<SX, SY, K>();
  const double gt0 = pT0[off], gt1 = pT1[off], gu = pG[off + Cfg::FS], gv = pG[off + 2 * Cfg::FS];
  const double a0 = pA0[off], a1 = pA1[off], a2 = pA2[off];
  constexpr int T = K, U = (K + 1) % 3, V = (K + 2) % 3;
  // compile-time part of rho along the natural axes (rho_z = sH is applied through khz)
  constexpr int rx = SX ? -1 : 1, ry = SY ? -1 : 1;
  constexpr int rhoT = T == 0 ? rx : (T == 1 ? ry : 1), rhoU = U == 0 ? rx : (U == 1 ? ry : 1), rhoV = V == 0 ? rx : (V == 1 ? ry : 1);
  constexpr bool zu = (T == 2) || (U == 2), zv = (T == 2) || (V == 2); // which transverse terms carry sH
#pragma unroll
  for (int cc = 0; cc < 2; ++cc)
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const int rho[3] = {a, b, cc};
        const int rt = rho[T], ru = rho[U], rv = rho[V];
        const double gt = rt ? gt1 : gt0;
        const double ct = rt ? Kc.kt[ru + rv] : -Kc.kt[ru + rv];
        const double mu_ = zu ? khz[rv] : Kc.kh[rv], mv_ = zv ? khz[ru] : Kc.kh[ru];
        const int su = -rhoT * (ru ? -rhoU : rhoU), sv = -rhoT * (rv ? -rhoV : rhoV);
        const double cu = su > 0 ? mu_ : -mu_, cv = sv > 0 ? mv_ : -mv_;
        const double h = gt * ct + gu * cu + gv * cv;
        const int dx = a ? (SX ? 1 : -1) : 0, dy = b ? (SY ? 1 : -1) : 0;
        const int col = (dy + 1) * 3 + (dx + 1);
        if (cc) {
          acc2[col][0] += h * a0;
          acc2[col][1] += h * a1;
          acc2[col][2] += h * a2;
        } else {
          acc1[col][0] += h * a0;
          acc1[col][1] += h * a1;
          acc1[col][2] += h * a2;
        }
      }
}

// one z half of the node's octants: H = 0 the 4 elements below the plane (acc2 = dz -1 group), H = 1 above (acc2 =
// dz +1 group); acc1 = dz 0 group.  No test for elements / nodes outside the domain: their records are zero.
__host__ __device__ __forceinline__ void sweep_half(const Thr &N, const double *sm, const BoxParams &Kc, int H, double (&acc1)[9][3],
                                                    double (&acc2)[9][3])
{
  const double *pG = sm + N.oG, *pGp = sm + N.oGp, *pA0 = sm + N.oA[0], *pA1 = sm + N.oA[1], *pA2 = sm + N.oA[2];
  // columns on the row's side of an axis have sigma = rho: x / y edges: compile-time; z edges: rho_z = sH
  const double *pZ0 = H ? pG : pGp, *pZ1 = H ? pGp : pG;
  double khz[2];
  khz[0] = H ? -Kc.kh[0] : Kc.kh[0];
  khz[1] = H ? -Kc.kh[1] : Kc.kh[1];
#define NALU_B5_Q(SX, SY)                                                                                          \
  sweep_point<SX, SY, 0>(SX ? pG : pGp, SX ? pGp : pG, pG, pA0, pA1, pA2, Kc, khz, acc1, acc2);                     \
  sweep_point<SX, SY, 1>(SY ? pG : pGp, SY ? pGp : pG, pG, pA0, pA1, pA2, Kc, khz, acc1, acc2);                     \
  sweep_point<SX, SY, 2>(pZ0, pZ1, pG, pA0, pA1, pA2, Kc, khz, acc1, acc2);
  NALU_B5_Q(0, 0) NALU_B5_Q(1, 0) NALU_B5_Q(0, 1) NALU_B5_Q(1, 1)
#undef NALU_B5_Q
}

// ---- fused lumped MomentumMassElemKernel (src/kernel/MomentumMassElemKernel.C:98-171) -----------------------------------
// With the shifted (lumped) SCV shape function every sub-control volume talks to its own node only, so the 8 elements
// around a node contribute (sum of their sub-control volumes at the node) x nodal quantities to the node's diagonal
// entries and residual.  Share of one z half of the dual volume: the 4 elements of the half, corner (!SX, !SY, !H).
__host__ __device__ __forceinline__ double dual_volume_half(const Thr &N, const double *sm, int H)
{
  using C = Cfg;
  const double *s = sm + C::OFF_SCV + (N.lx + C::ELX * N.ly) * 16;
  const int bz = H ? 0 : 4;
  // node_of_bits(bx, by, bz) = bz*4 + by*2 + (bx ^ by): quadrant (SX,SY) sees the node as corner (!SX, !SY); every
  // volume is stored as two partial sums (scv_volumes_part)
  const double *e0 = s + bz + 2, *e1 = s + 16 + bz + 3, *e2 = s + C::ELX * 16 + bz + 1, *e3 = s + (C::ELX + 1) * 16 + bz;
  return ((e0[0] + e0[8]) + (e1[0] + e1[8])) + ((e2[0] + e2[8]) + (e3[0] + e3[8]));
}
// lane (node, i < 3): diagonal entry (i,i) of the node's own block and residual of dof row i.  The nodal values are
// fetched (mass_fetch) right after the sweep and used (mass_apply) after the dz=0 group is finalized: the global-memory
// latency hides under the finalize arithmetic.
struct MassNode {
  double rp, rn, rm, up, un, um, gp;
};
__host__ __device__ __forceinline__ void mass_fetch(const BoxParams &P, long long n, int i, MassNode &m)
{
  m.rp = P.rho_np1[n]; m.rn = P.rho_n[n]; m.rm = P.rho_nm1[n];
  m.up = P.vel[n * 3 + i]; m.un = P.vel_n[n * 3 + i]; m.um = P.vel_nm1[n * 3 + i]; m.gp = P.dpdx[n * 3 + i];
}
// the same addresses, touched before the sweep so that mass_fetch finds them in L1 / L2 (no registers held meanwhile)
__host__ __device__ __forceinline__ void mass_prefetch(const BoxParams &P, long long n, int i)
{
#ifdef __CUDA_ARCH__
  asm volatile("prefetch.global.L1 [%0];" ::"l"(P.rho_np1 + n));
  asm volatile("prefetch.global.L1 [%0];" ::"l"(P.rho_n + n));
  asm volatile("prefetch.global.L1 [%0];" ::"l"(P.rho_nm1 + n));
  asm volatile("prefetch.global.L1 [%0];" ::"l"(P.vel + n * 3 + i));
  asm volatile("prefetch.global.L1 [%0];" ::"l"(P.vel_n + n * 3 + i));
  asm volatile("prefetch.global.L1 [%0];" ::"l"(P.vel_nm1 + n * 3 + i));
  asm volatile("prefetch.global.L1 [%0];" ::"l"(P.dpdx + n * 3 + i));
#else
  (void)P; (void)n; (void)i;
#endif
}
__host__ __device__ __forceinline__ void mass_apply(const BoxParams &P, const MassNode &m, double V, double &diag, double &rr)
{
  diag += P.g1_dt * m.rp * V;
  rr += -(P.g1_dt * m.rp * m.up + P.g2_dt * m.rn * m.un + P.g3_dt * m.rm * m.um) * V - m.gp * V * (m.rp * P.dens_fac + P.om_dens_fac);
}

// ---- finalize of one dz group: diagonal term, rhs share ---------------------------------------------------------------
// offset (doubles from the start of shared memory) of the staged velocity of the row node in ring plane `slot`
__host__ __device__ __forceinline__ int vel_off(const Thr &N, int slot)
{
  using C = Cfg;
  return C::OFF_RING + slot * C::PLN + C::OFF_VEL + ((N.lx + 1) + C::RNX * (N.ly + 1)) * C::VEL;
}
// D[c]: the quad's diagonal terms (acc[c][0] of lane 3).  Adds them, takes the group's share of rhs_i = -(row . u) and
// returns in T what the other dof rows need for the div(u) stress: T[k'] = sum_c acc[c][k'] u_c[i] before the diagonal
// term (T[0] stays, T[1] goes to dof row (i+1) mod 3, T[2] to (i+2) mod 3).  Returns row . u.
__host__ __device__ __forceinline__ double fin_apply(const Thr &N, const double *sm, int uoff, double (&acc)[9][3], const double (&D)[9],
                                                     double (&T)[3])
{
  using C = Cfg;
  const double *pu0 = sm + uoff + N.rot[0], *pu1 = sm + uoff + N.rot[1], *pu2 = sm + uoff + N.rot[2];
  double r0 = 0.0, r1 = 0.0, r2 = 0.0;
  T[0] = T[1] = T[2] = 0.0;
#pragma unroll
  for (int c = 0; c < 9; ++c) {
    const int o = ((c % 3 - 1) + C::RNX * (c / 3 - 1)) * C::VEL;
    const double u0 = pu0[o], u1 = pu1[o], u2 = pu2[o];
    T[0] += acc[c][0] * u0;
    T[1] += acc[c][1] * u0;
    T[2] += acc[c][2] * u0;
    acc[c][0] += D[c];
    r0 += acc[c][0] * u0;
    r1 += acc[c][1] * u1;
    r2 += acc[c][2] * u2;
  }
  return r0 + (r1 + r2);
}

// ---- output ------------------------------------------------------------------------------------------------------------
// lane i < 3 of node lx lays dof row i of dz group g out like its CSR run in the warp's scratch: run r = 3 lx + i
__host__ __device__ __forceinline__ void out_stage_natural(const Thr &N, double *scr, const double (&acc)[9][3])
{
  if (N.i >= 3) return;
  double *s = scr + (3 * N.lx + N.i) * 27;
#pragma unroll
  for (int c = 0; c < 9; ++c) {
    s[c * 3 + N.rot[0]] = acc[c][0];
    s[c * 3 + N.rot[1]] = acc[c][1];
    s[c * 3 + N.rot[2]] = acc[c][2];
  }
}
// rows that are not natural (domain boundary, rank interface): the slot of every column comes from the graph's
// node_slot27; `sl` = the 27 slot bytes of the node
__host__ __device__ __forceinline__ void out_stage_slots(const Thr &N, const Meta &M, const unsigned char *sl, int g, double *scr,
                                                         const double (&acc)[9][3])
{
  if (N.i >= 3) return;
  double *s = scr + (3 * N.lx + N.i) * 27;
  const int s0 = M.s0(N.node, g);
#pragma unroll
  for (int c = 0; c < 9; ++c) {
    const int v = sl[g * 9 + c];
    if (v == 0xFF) continue;
    double *d = s + (v - s0) * 3;
    d[N.rot[0]] = acc[c][0];
    d[N.rot[1]] = acc[c][1];
    d[N.rot[2]] = acc[c][2];
  }
}
// run of dof row i of dz group g of a node: first double in `values`
__host__ __device__ __forceinline__ long long run_first(const Meta &M, int node, int i, int g)
{
  return M.rowbase(node) + ((long long)i * M.nb(node) + M.s0(node, g)) * 3;
}
// columns of the row that no local element touches (remote columns announced by the ranks sharing the node): in
// overwrite mode they must read zero.  Lane (node, i < 3) handles dof row i.
__host__ __device__ inline void zero_uncovered(const Meta &M, int node, int i, double *values)
{
  const int nb = M.nb(node);
  const int c0 = M.cnt(node, 0), c1 = M.cnt(node, 1), c2 = M.cnt(node, 2);
  if (nb == c0 + c1 + c2) return;
  const int s0 = M.s0(node, 0), s1 = M.s0(node, 1), s2 = M.s0(node, 2);
  double *row = values + M.rowbase(node) + (long long)i * nb * 3;
  for (int s = 0; s < nb; ++s) {
    if ((s >= s0 && s < s0 + c0) || (s >= s1 && s < s1 + c1) || (s >= s2 && s < s2 + c2)) continue;
    row[s * 3 + 0] = row[s * 3 + 1] = row[s * 3 + 2] = 0.0;
  }
}

// ---- node metadata of plane z: one lane per node ---------------------------------------------------------------------
__host__ __device__ inline void meta_load(const BoxParams &P, double *sm, int buf, int node, int x0, int y0, int z)
{
  using C = Cfg;
  unsigned long long *w = reinterpret_cast<unsigned long long *>(sm + C::OFF_META + buf * C::META);
  const int x = x0 + node % C::TBX, y = y0 + node / C::TBX;
  const bool in = x < P.NX && y < P.NY && z >= 0 && z < P.NZ;
  unsigned long long w0 = 0, w1 = 0, w2 = 0;
  if (in) {
    const unsigned long long *d = P.node_desc + 3 * (x + (long long)P.NX * (y + (long long)P.NY * z));
    w0 = d[0]; w1 = d[1]; w2 = d[2];
  }
  w[node] = w0;
  w[C::NNODE + node] = w1;
  w[2 * C::NNODE + node] = w2;
}
// an x line is "natural" when its 8 nodes are interior rows in lattice order, stored back to back with the same group
// order: the 24 runs of a group then sit at compile-time offsets from the first one
__host__ __device__ inline bool node_in_natural_line(const Meta &M, int node)
{
  const int n0 = node & ~(Cfg::TBX - 1), lx = node - n0;
  return M.natural(node) && M.rowbase(node) == M.rowbase(n0) + 243LL * lx && M.order(node) == M.order(n0);
}
__host__ __device__ inline bool line_natural(const Meta &M, int ly)
{
  bool ok = true;
  for (int lx = 0; lx < Cfg::TBX; ++lx) ok = ok && node_in_natural_line(M, Cfg::TBX * ly + lx);
  return ok;
}

} // namespace box5
} // namespace nalu_b200
