// box3_tile.cuh -- warp-specialised variant of the column-pair kernel (box2_tile.cuh): the same
// phase-1 / phase-2 arithmetic, but producers (phase 1), consumers (phase 2) and the copy-out /
// staging warps of ONE CTA per SM run concurrently on consecutive half planes behind mbarriers
// instead of alternating at CTA barriers.  That needs the buffers a half plane reads and the next
// one writes to be distinct:
//   records   F_BF (bottom-face points of layer z, read by the upper octants of plane z),
//             F_TF (top-face points of layer z, read by the lower octants of plane z+1),
//             ZE[2] (z-edge points of layer z, read by both);
//   staging   SA  dz=0 | dz=+1 groups of plane z (written by the upper half, leaves as 54-double runs),
//             SB  dz=-1 group (written by the lower half, leaves as 27-double runs while the upper half runs),
//             SC  partial of the dz=0 group (lower half -> upper half, private to the thread);
//   ring      4 node planes, 2 mass-flow layers, 3 node-metadata buffers.
// What is computed, and every formula, is box2_tile.cuh's (same references into the Nalu sources).
#pragma once
#include "box2_tile.cuh"

namespace nalu_b200 {
namespace box3 {

using box::BoxParams;
using box2::ClsK;
using box2::Meta;
using box2::MetaRaw;
using box2::Rec;
using box2::T2;

struct Cfg3 {
  using B = box2::Cfg;
  static constexpr int NNODE = B::NNODE, NEL = B::NEL, ES = B::ES, PLN = B::PLN, MDB = B::MDB;
  static constexpr int NCONS = 128, NPROD = 128, NAUX = 128, NT = NCONS + NPROD + NAUX;
  static constexpr int SA_ROW = 56, SA_NODE = 3 * SA_ROW + 1; // [parity pad | dz=0 (27) | dz=+1 (27)]
  static constexpr int SB_ROW = 28, SB_NODE = 3 * SB_ROW + 1; // [parity pad | 27]
  static constexpr int OFF_BF = 0, OFF_TF = OFF_BF + NEL * ES, OFF_ZE = OFF_TF + NEL * ES, // ZE: two buffers
                       OFF_SA = OFF_ZE + 2 * NEL * ES, OFF_SB = OFF_SA + NNODE * SA_NODE,
                       OFF_SC = OFF_SB + NNODE * SB_NODE, OFF_RING = OFF_SC + NNODE * SB_NODE,
                       OFF_MD = OFF_RING + 4 * PLN, OFF_RP = OFF_MD + 2 * MDB, OFF_META = OFF_RP + 4 * NNODE * 3,
                       OFF_QK = OFF_META + 3 * B::META, OFF_BAR = OFF_QK + 4 * 24;
  static constexpr int NBAR = 8;
  static constexpr int SMEM_DOUBLES = OFF_BAR + NBAR;
};

__host__ __device__ inline Meta meta3_at(double *sm, int z)
{
  using B = box2::Cfg;
  double *m = sm + Cfg3::OFF_META + ((z % 3 + 3) % 3) * B::META;
  Meta M;
  M.rowbase = reinterpret_cast<long long *>(m + B::META_RB);
  M.nb = reinterpret_cast<int *>(m + B::META_NB);
  M.row = reinterpret_cast<int *>(m + B::META_ROW);
  M.slot = reinterpret_cast<unsigned char *>(m + B::META_SLOT);
  M.gp = reinterpret_cast<unsigned short *>(m + B::META_GP);
  M.cnt = reinterpret_cast<unsigned char *>(m + B::META_CNT);
  return M;
}

// gp[g] = first slot of dz group g | staging offset inside its array << 8 (g = 0: SB, 0; g = 1: SA, 0;
// g = 2: SA, behind the dz=0 group); gp[3] = parity bits of the dz=0 rows | merged << 3 | parity bits of
// the dz=-1 rows << 4
__host__ __device__ inline void meta_store3(const BoxParams &P, const Meta &M, const MetaRaw &R, int node)
{
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
  unsigned pb = 0, pbB = 0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    pb |= (unsigned)((rowbase + ((long long)i * nb + mn[1]) * 3 + node) & 1) << i;
    pbB |= (unsigned)((rowbase + ((long long)i * nb + mn[0]) * 3 + node) & 1) << i;
  }
  const unsigned b2 = cnt[1] * 3;
  const unsigned fix2 = (unsigned)((node + (pb & 1) + b2 + rowbase + (long long)mn[2] * 3) & 1);
  const unsigned merged = cnt[2] && cnt[1] && mn[2] == mn[1] + cnt[1] && !fix2;
  M.gp[node * 4 + 0] = (unsigned short)mn[0];
  M.gp[node * 4 + 1] = (unsigned short)mn[1];
  M.gp[node * 4 + 2] = (unsigned short)(mn[2] | ((b2 + fix2) << 8));
  M.gp[node * 4 + 3] = (unsigned short)(pb | (merged << 3) | (pbB << 4));
#pragma unroll
  for (int g = 0; g < 3; ++g) M.cnt[node * 4 + g] = (unsigned char)cnt[g];
  M.rowbase[node] = rowbase;
  M.nb[node] = nb;
  M.row[node] = R.row < 0 ? 0 : R.row;
}

// kind 0: dz=-1 run in SB; 1: dz=0 (+ dz=+1 when merged) run in SA; 2: dz=+1 run in SA when not merged
__host__ __device__ inline box2::Run run_of3(const Meta &M, int node, int i, int kind)
{
  using C = Cfg3;
  box2::Run r;
  const unsigned gp = M.gp[node * 4 + kind], nf = M.gp[node * 4 + 3];
  const bool merged = (nf >> 3) & 1;
  r.L = 3 * M.cnt[node * 4 + kind];
  if (merged && kind == 1) r.L += 3 * M.cnt[node * 4 + 2];
  if (merged && kind == 2) r.L = 0;
  r.G = M.rowbase[node] + ((long long)i * M.nb[node] + (gp & 0xFF)) * 3;
  if (kind == 0) r.soff = C::OFF_SB + node * C::SB_NODE + i * C::SB_ROW + ((nf >> (4 + i)) & 1);
  else r.soff = C::OFF_SA + node * C::SA_NODE + i * C::SA_ROW + ((nf >> i) & 1) + (gp >> 8);
  return r;
}

// ---- flush (see box2_tile.cuh): offsets into SA / SB / SC -----------------------------------------
struct FlushPrep3 {
  int s[2];   // smem offset of row 0 of the block of column c (parity of row 0 not yet added), < 0: none
  int ps;     // H = 1: smem offset of the partial of the c = 0 block in SC
  int uo[2];
};
// final0: the dz=0 block of this half is final (upper half, or lower half of the top plane) -> SA
__host__ __device__ __forceinline__ void flush_prep3(FlushPrep3 &F, const T2 &N, const Meta &M, const ClsK &Q, int H, int sx,
                                                     int sy, int z, unsigned gp0, unsigned gp1, bool final0)
{
  using C = Cfg3;
  using B = box2::Cfg;
  const int dx = Q.a ? (sx ? 1 : -1) : 0, dy = Q.b ? (sy ? 1 : -1) : 0;
  const int idx = (dy + 1) * 3 + (dx + 1);
  const int g1 = H ? 2 : 0;
  const unsigned slot0 = M.slot[N.node * 28 + 9 + idx], slot1 = M.slot[N.node * 28 + g1 * 9 + idx];
  const int o0 = (int)(slot0 - (gp0 & 0xFF)) * 3;
  // c = 0: dz=0 block: lower half -> SC (partial), upper half -> SA
  F.ps = C::OFF_SC + N.node * C::SB_NODE + o0;
  F.s[0] = slot0 == 0xFF ? -1 : (final0 ? C::OFF_SA + N.node * C::SA_NODE + o0 : F.ps);
  // c = 1: lower half -> SB (dz=-1), upper half -> SA behind the dz=0 group (dz=+1)
  const int o1 = (int)(slot1 - (gp1 & 0xFF)) * 3;
  F.s[1] = slot1 == 0xFF ? -1 : (H ? C::OFF_SA + N.node * C::SA_NODE + (int)(gp1 >> 8) + o1 : C::OFF_SB + N.node * C::SB_NODE + o1);
  const int uo = ((N.lx + 1 + dx) + B::RNX * (N.ly + 1 + dy)) * B::NF + 4;
  F.uo[0] = C::OFF_RING + (z & 3) * C::PLN + uo;
  F.uo[1] = C::OFF_RING + ((z + (H ? 1 : 3)) & 3) * C::PLN + uo;
}
// row pitch / parity of the array the block of column c goes to: (H=0,c=0) SC: pitch 28, no parity;
// (H=0,c=1) SB: pitch 28, parity bits pbB; (H=1, both) SA: pitch 56, parity bits pb
__host__ __device__ __forceinline__ int row_off3(int H, int c, int i, unsigned nf, bool final0)
{
  using C = Cfg3;
  if (H || (c == 0 && final0)) return i * C::SA_ROW + ((nf >> i) & 1);
  return i * C::SB_ROW + (c ? ((nf >> (4 + i)) & 1) : 0);
}
struct FlushOps3 {
  double u[2][3];
  double part[9];
};
__host__ __device__ __forceinline__ void flush_fetch3(FlushOps3 &O, const FlushPrep3 &F, const double *sm, bool add)
{
  using C = Cfg3;
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int j = 0; j < 3; ++j) O.u[c][j] = sm[F.uo[c] + j];
#pragma unroll
  for (int t = 0; t < 9; ++t) O.part[t] = 0.0;
  if (add && F.s[0] >= 0) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double *d = sm + F.ps + i * C::SB_ROW;
      O.part[i * 3 + 0] = d[0];
      O.part[i * 3 + 1] = d[1];
      O.part[i * 3 + 2] = d[2];
    }
  }
}
__host__ __device__ __forceinline__ void flush_commit3(T2 &N, double *sm, const FlushPrep3 &F, const FlushOps3 &O, int H,
                                                       unsigned nf, const double (&acc)[2][9], const double (&dm)[2],
                                                       double f23, bool final0)
{
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
      double *d = sm + F.s[c] + row_off3(H, c, i, nf, final0);
      if (c == 0) {
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

// one z half of the node's octants (box2::p2_half_t with this kernel's buffers)
template <bool SWAP>
__host__ __device__ inline void p2_half3_t(T2 &N, double *sm, const ClsK &Q, double f23, int H, int z, bool has_lower,
                                           bool has_upper)
{
  const bool final0 = H || !has_upper;
  using C = Cfg3;
  using B = box2::Cfg;
  // lower half: top-face points of layer z-1 and its z-edge points; upper half: bottom-face points of layer z
  const double *eF = sm + (H ? C::OFF_BF : C::OFF_TF) + (N.lx + B::ELX * N.ly) * B::ES;
  const double *eZ = sm + C::OFF_ZE + (((H ? z : z + 1) & 1) * C::NEL + N.lx + B::ELX * N.ly) * B::ES;
  const unsigned fm = (Q.a & Q.b) ? 0xFu : ((Q.a | Q.b) ? 0xAu : 0x8u);
  double z1[2], z2[2], zn[2];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    z1[c] = box2::flip_if(Q.c1[c][2], !H);
    z2[c] = box2::flip_if(Q.c2[c][2], !H);
    zn[c] = box2::flip_if(Q.cn[c][2], !H);
  }
  double acc[2][9], dm[2];
#pragma unroll
  for (int t = 0; t < 9; ++t) acc[0][t] = acc[1][t] = 0.0;
  dm[0] = dm[1] = 0.0;
  constexpr int X1 = SWAP ? 0 : 1, Y1 = SWAP ? 1 : 0, X2 = SWAP ? 1 : 0, Y2 = SWAP ? 0 : 1;
  const Meta M = meta3_at(sm, z);
  const unsigned gp0 = M.gp[N.node * 4 + 1], gp1 = M.gp[N.node * 4 + (H ? 2 : 0)], nf = M.gp[N.node * 4 + 3];
  const bool add = H && has_lower;
  FlushPrep3 FP0, FP1, FP2, FP3;
  flush_prep3(FP0, N, M, Q, H, 0, 0, z, gp0, gp1, final0);
  flush_prep3(FP1, N, M, Q, H, X1, Y1, z, gp0, gp1, final0);
  flush_prep3(FP2, N, M, Q, H, X2, Y2, z, gp0, gp1, final0);
  flush_prep3(FP3, N, M, Q, H, 1, 1, z, gp0, gp1, final0);
  FlushOps3 FO;
  Rec RA, RB;
  box2::rec_load(RA, box2::rec_addr<0, 0, 0>(eF, eZ));
#define NALU_P3_QUAD(J, SX, SY, NX_, NY_, LAST)                                                   \
  if (fm & (1u << J)) flush_fetch3(FO, FP##J, sm, add);                                           \
  box2::rec_load(RB, box2::rec_addr<SX, SY, 1>(eF, eZ));                                          \
  box2::rec_use<SX, SY, 0>(RA, Q, z1, z2, zn, acc, dm);                                           \
  box2::rec_load(RA, box2::rec_addr<SX, SY, 2>(eF, eZ));                                          \
  box2::rec_use<SX, SY, 1>(RB, Q, z1, z2, zn, acc, dm);                                           \
  if (!LAST) box2::rec_load(RB, box2::rec_addr<NX_, NY_, 0>(eF, eZ));                             \
  box2::rec_use<SX, SY, 2>(RA, Q, z1, z2, zn, acc, dm);                                           \
  if (fm & (1u << J)) {                                                                           \
    flush_commit3(N, sm, FP##J, FO, H, nf, acc, dm, f23, final0);                                         \
    _Pragma("unroll") for (int t = 0; t < 9; ++t) acc[0][t] = acc[1][t] = 0.0;                    \
    dm[0] = dm[1] = 0.0;                                                                          \
  }                                                                                               \
  if (!LAST) RA = RB;
  NALU_P3_QUAD(0, 0, 0, X1, Y1, false)
  NALU_P3_QUAD(1, X1, Y1, X2, Y2, false)
  NALU_P3_QUAD(2, X2, Y2, 1, 1, false)
  NALU_P3_QUAD(3, 1, 1, 1, 1, true)
#undef NALU_P3_QUAD
}
__host__ __device__ inline void p2_half3(T2 &N, double *sm, const ClsK &Q, double f23, int H, int z, bool has_lower,
                                         bool has_upper)
{
  if (Q.a && !Q.b) p2_half3_t<true>(N, sm, Q, f23, H, z, has_lower, has_upper);
  else p2_half3_t<false>(N, sm, Q, f23, H, z, has_lower, has_upper);
}

// phase 1 item list of one producer step.  kind 0: prologue (top face + z edges of layer z), 1: bottom
// face + z edges, 2: top face
__host__ __device__ inline void p1_step3(const BoxParams &P, double *sm, int kind, int x0, int y0, int z, int t, int nt)
{
  using C = Cfg3;
  using B = box2::Cfg;
  const int n = kind == 2 ? 2 * B::NEL : 4 * B::NEL;
  const double *pl0 = sm + C::OFF_RING + (z & 3) * C::PLN, *pl1 = sm + C::OFF_RING + ((z + 1) & 3) * C::PLN;
  const double *mdb = sm + C::OFF_MD + (z & 1) * C::MDB;
  double *ZE = sm + C::OFF_ZE + (z & 1) * C::NEL * B::ES;
  for (int it = t; it < n; it += nt) {
    const int k = it / B::NEL, el = it - k * B::NEL;
    const int g = k < 2 ? k : 2;
    const int side = k < 2 ? (kind == 1 ? 0 : 1) : k - 2;
    box2::p1_item(P, el, g, side, x0, y0, pl0, pl1, mdb, sm + (side && g < 2 ? C::OFF_TF : C::OFF_BF), ZE);
  }
}

} // namespace box3
} // namespace nalu_b200
