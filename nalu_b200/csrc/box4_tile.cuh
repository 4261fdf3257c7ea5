// box4_tile.cuh -- row-thread phase 2 for the tiled write-once MomentumAdvDiff assembly.
//
// Same contract, tile (8 x 4 nodes, z sweep), phase 1, records, field / mass-flow / metadata staging and phase
// structure as the column-pair kernel (box2_tile.cuh); what changes is who owns what in phase 2:
//   * a node is served by the threads (node, i), i = 0..2 (warps 0-2, lane = node), and one thread for the
//     advection operator M_c = sum +-N_c mdot (warp 3).  Thread (node, i) keeps dof row i of the 18 column
//     nodes of a z half (9 in the row's plane, 9 in the neighbouring plane) in 54 registers:
//     acc[c][j] += h_i A_j with h_i = G~_i . dN_c (box2_tile.cuh, rec_apply) -- 6 FMAs per (point, column)
//     and thread, every index compile-time (classes, quadrants and points are unrolled; the class constants
//     are the 8 runtime numbers of P2K);
//   * the diagonal term of a block, M_c + tr(acc_c) (MomentumAdvDiffElemKernel.C:160-176), and the div(u)
//     transpose term need the other rows of the block: exchanged once per half through shared memory;
//   * the dz=0 partial of the lower half waits in shared memory for the upper half; finished groups go from
//     registers straight to the CSR through the slot map -- no staging image, no copy-out loop, no
//     read-modify-write of staged blocks.
// Shared memory: the staging region of box2 (Cfg::OFF_S, 5408 doubles) holds the parked partials (2688) and the
// exchange buffer (2592).
#pragma once
#include "box2_tile.cuh"

namespace nalu_b200 {
namespace box4 {

using box::BoxParams;
using C = box2::Cfg;
using box2::Meta;
using box2::P2K;

static constexpr int OFF_PARK = C::OFF_S;                 // [i][28][node]: dz=0 partial (27) + rhs partial
static constexpr int PARKN = 28;
static constexpr int OFF_EX = OFF_PARK + C::NNODE * 3 * PARKN; // [4][18][node] diag exchange, then [3][3][node] divU exchange
static constexpr int OFF_EXQ = OFF_EX + 4 * 18 * C::NNODE;
static_assert(OFF_EXQ + 3 * 3 * C::NNODE <= C::OFF_S + C::NNODE * C::SNODE, "park + exchange must fit the staging region");

struct Row {
  int lx, ly, node, i; // i = 0..2: dof row; 3: advection-operator thread
};

__host__ __device__ __forceinline__ const double *rec_ptr(const double *sm, const Row &N)
{
  return sm + (N.lx + C::ELX * N.ly) * C::ES;
}

// one record (quadrant SX,SY; point K: 0 x edge, 1 y edge, 2 z edge; half H) into the 8 columns of the element
template <int H, int SX, int SY, int K>
__host__ __device__ __forceinline__ void rec_row(const double *sm, const Row &N, const P2K &Kc, double (&acc)[18][3])
{
  const double *eF = sm + C::OFF_F + (N.lx + C::ELX * N.ly) * C::ES;
  const double *eZ = sm + C::OFF_ZE + (N.lx + C::ELX * N.ly) * C::ES;
  const double *o = box2::rec_addr<SX, SY, K>(eF, eZ);
  const double o0 = o[N.i * 3 + 0], o1 = o[N.i * 3 + 1], o2 = o[N.i * 3 + 2];
  const double A0 = o[9], A1 = o[10], A2 = o[11];
  constexpr bool NEG = K == 0 ? !SX : (K == 1 ? !SY : false);
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const int rho[3] = {a, b, c};
        const int rt = rho[K], ru = rho[(K + 1) % 3], rv = rho[(K + 2) % 3];
        const double c0 = rt ? Kc.kt[ru + rv] : -Kc.kt[ru + rv];
        double c1 = ru ? -Kc.kh[rv] : Kc.kh[rv];
        double c2 = rv ? -Kc.kh[ru] : Kc.kh[ru];
        if (K == 2 && !H) { // z-edge points: tau_z = +1 for the elements above the plane
          c1 = -c1;
          c2 = -c2;
        }
        const double w = o1 * c1 + o2 * c2;
        const double h = NEG ? o0 * c0 - w : o0 * c0 + w;
        const int dx = a ? (SX ? 1 : -1) : 0, dy = b ? (SY ? 1 : -1) : 0;
        const int col = c * 9 + (dy + 1) * 3 + (dx + 1);
        acc[col][0] += h * A0;
        acc[col][1] += h * A1;
        acc[col][2] += h * A2;
      }
}
template <int H, int SX, int SY, int K>
__host__ __device__ __forceinline__ void rec_adv(const double *sm, const Row &N, const P2K &Kc, double (&dm)[18])
{
  const double *eF = sm + C::OFF_F + (N.lx + C::ELX * N.ly) * C::ES;
  const double *eZ = sm + C::OFF_ZE + (N.lx + C::ELX * N.ly) * C::ES;
  const double md = box2::rec_addr<SX, SY, K>(eF, eZ)[12];
  constexpr bool NEG = K == 0 ? !SX : (K == 1 ? !SY : (H == 0));
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const int rho[3] = {a, b, c};
        const int ru = rho[(K + 1) % 3], rv = rho[(K + 2) % 3];
        const double cn = Kc.kn[ru + rv];
        const int dx = a ? (SX ? 1 : -1) : 0, dy = b ? (SY ? 1 : -1) : 0;
        const int col = c * 9 + (dy + 1) * 3 + (dx + 1);
        if (NEG) dm[col] -= cn * md;
        else dm[col] += cn * md;
      }
}

// the 12 records of one z half (H = 0: elements below the plane, 1: above)
template <int H>
__host__ __device__ __forceinline__ void half_rows(const double *sm, const Row &N, const P2K &Kc, double (&acc)[18][3])
{
#define NALU_B4_Q(SX, SY)                \
  rec_row<H, SX, SY, 0>(sm, N, Kc, acc); \
  rec_row<H, SX, SY, 1>(sm, N, Kc, acc); \
  rec_row<H, SX, SY, 2>(sm, N, Kc, acc);
  NALU_B4_Q(0, 0) NALU_B4_Q(1, 0) NALU_B4_Q(0, 1) NALU_B4_Q(1, 1)
#undef NALU_B4_Q
}
template <int H>
__host__ __device__ __forceinline__ void half_adv(const double *sm, const Row &N, const P2K &Kc, double (&dm)[18])
{
#define NALU_B4_Q(SX, SY)               \
  rec_adv<H, SX, SY, 0>(sm, N, Kc, dm); \
  rec_adv<H, SX, SY, 1>(sm, N, Kc, dm); \
  rec_adv<H, SX, SY, 2>(sm, N, Kc, dm);
  NALU_B4_Q(0, 0) NALU_B4_Q(1, 0) NALU_B4_Q(0, 1) NALU_B4_Q(1, 1)
#undef NALU_B4_Q
}

// offset (doubles, from the start of shared memory) of the velocity of column `col` of the half
template <int H>
__host__ __device__ __forceinline__ int vel_off(const Row &N, int col, int z)
{
  const int c = col / 9, idx = col - c * 9;
  const int dx = idx % 3 - 1, dy = idx / 3 - 1;
  const int plane = c ? (z + (H ? 1 : 2)) % 3 : z % 3;
  return C::OFF_RING + plane * C::PLN + ((N.lx + 1 + dx) + C::RNX * (N.ly + 1 + dy)) * C::NF + 4;
}

// ---- exchange, part 1 (before the barrier): what the other threads of the node need from this one ---------------
template <int H>
__host__ __device__ __forceinline__ void exchange_put_row(double *sm, const Row &N, const double (&acc)[18][3], int z)
{
  // divU transpose term: thread r contributes q_i = sum_col acc[col][i] u_col[r] to rhs_i (box2_tile.cuh, flush_commit)
  double q[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int col = 0; col < 18; ++col) {
    const double u = sm[vel_off<H>(N, col, z) + N.i];
    q[0] += acc[col][0] * u;
    q[1] += acc[col][1] * u;
    q[2] += acc[col][2] * u;
    // the diagonal entry of this row: its share of tr(acc_col); N.i is warp-uniform
    const double d = N.i == 0 ? acc[col][0] : (N.i == 1 ? acc[col][1] : acc[col][2]);
    sm[OFF_EX + (N.i * 18 + col) * C::NNODE + N.node] = d;
  }
#pragma unroll
  for (int ii = 0; ii < 3; ++ii) sm[OFF_EXQ + (N.i * 3 + ii) * C::NNODE + N.node] = q[ii];
}
__host__ __device__ __forceinline__ void exchange_put_adv(double *sm, const Row &N, const double (&dm)[18])
{
#pragma unroll
  for (int col = 0; col < 18; ++col) sm[OFF_EX + (3 * 18 + col) * C::NNODE + N.node] = dm[col];
}

// ---- exchange, part 2 (after the barrier): diagonal term, rhs share of the half ---------------------------------
template <int H>
__host__ __device__ __forceinline__ void finalize_row(const double *sm, const Row &N, double (&acc)[18][3], double &rr, double f23,
                                                      int z)
{
#pragma unroll
  for (int col = 0; col < 18; ++col) {
    const double D = (sm[OFF_EX + (0 * 18 + col) * C::NNODE + N.node] + sm[OFF_EX + (1 * 18 + col) * C::NNODE + N.node]) +
                     (sm[OFF_EX + (2 * 18 + col) * C::NNODE + N.node] + sm[OFF_EX + (3 * 18 + col) * C::NNODE + N.node]);
    if (N.i == 0) acc[col][0] += D;
    else if (N.i == 1) acc[col][1] += D;
    else acc[col][2] += D;
    const double *u = sm + vel_off<H>(N, col, z);
    rr -= acc[col][0] * u[0] + acc[col][1] * u[1] + acc[col][2] * u[2];
  }
  rr += f23 * (sm[OFF_EXQ + (0 * 3 + N.i) * C::NNODE + N.node] + sm[OFF_EXQ + (1 * 3 + N.i) * C::NNODE + N.node] +
               sm[OFF_EXQ + (2 * 3 + N.i) * C::NNODE + N.node]);
}

// ---- registers -> CSR: the 9 blocks (row i) of dz group g held in acc[c9 .. c9+9) ----------------------------------
__host__ __device__ __forceinline__ void store_group(const BoxParams &P, const Meta &M, const Row &N, int g, int c9,
                                                     const double (&acc)[18][3])
{
  const long long rowbase = M.rowbase[N.node];
  const int nb = M.nb[N.node];
  double *row = P.values + rowbase + (long long)N.i * nb * 3;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const unsigned s = M.slot[N.node * 28 + g * 9 + k];
    if (s == 0xFF) continue;
    double *d = row + s * 3;
    if (P.overwrite) {
      d[0] = acc[c9 + k][0];
      d[1] = acc[c9 + k][1];
      d[2] = acc[c9 + k][2];
    } else {
      d[0] += acc[c9 + k][0];
      d[1] += acc[c9 + k][1];
      d[2] += acc[c9 + k][2];
    }
  }
}

// ---- coalesced output: the group is first laid out like the CSR run in the warp's own park slice (the slice is
// private to the warp of dof row i, so only warp-level synchronisation is needed), then streamed out one run per
// warp instruction (assemble_box4.cu, copy_staged_warp).  A register write-out (store_group) costs 243 sector
// requests per node -- lanes are 1 944 bytes apart -- and the LSU takes about one per clock.
struct RunW {
  long long G; // first double of the run in `values`
  int L;       // doubles
};
__host__ __device__ __forceinline__ double *warp_slice(double *sm, const Row &N) { return sm + OFF_PARK + N.i * PARKN * C::NNODE; }
__host__ __device__ __forceinline__ RunW stage_group(double *sm, const Meta &M, const Row &N, int g, int c9,
                                                     const double (&acc)[18][3])
{
  double *S = warp_slice(sm, N) + N.node * 27;
  const unsigned s0 = M.gp[N.node * 4 + g] & 0xFF;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const unsigned s = M.slot[N.node * 28 + g * 9 + k];
    if (s == 0xFF) continue;
    double *d = S + (s - s0) * 3;
    d[0] = acc[c9 + k][0];
    d[1] = acc[c9 + k][1];
    d[2] = acc[c9 + k][2];
  }
  RunW r;
  r.L = 3 * M.cnt[N.node * 4 + g];
  r.G = M.rowbase[N.node] + ((long long)N.i * M.nb[N.node] + s0) * 3;
  return r;
}

__host__ __device__ __forceinline__ void park(double *sm, const Row &N, const double (&acc)[18][3], double rr)
{
  double *p = sm + OFF_PARK + N.i * PARKN * C::NNODE + N.node;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    p[(k * 3 + 0) * C::NNODE] = acc[k][0];
    p[(k * 3 + 1) * C::NNODE] = acc[k][1];
    p[(k * 3 + 2) * C::NNODE] = acc[k][2];
  }
  p[27 * C::NNODE] = rr;
}
// the finished lower-half share (blocks with their diagonal term, rhs share) joins the upper-half share
__host__ __device__ __forceinline__ void unpark_add(const double *sm, const Row &N, double (&acc)[18][3], double &rr)
{
  const double *p = sm + OFF_PARK + N.i * PARKN * C::NNODE + N.node;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    acc[k][0] += p[(k * 3 + 0) * C::NNODE];
    acc[k][1] += p[(k * 3 + 1) * C::NNODE];
    acc[k][2] += p[(k * 3 + 2) * C::NNODE];
  }
  rr += p[27 * C::NNODE];
}

} // namespace box4
} // namespace nalu_b200
