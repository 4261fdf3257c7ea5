// CPU emulation of the tiled write-once scalar kernel-LIST kernel (nalu_b200/csrc/slist_tile.cuh compiled as plain C++):
// [ScalarMass(lumped)]? + {ScalarUpwAdvDiff | ScalarAdvDiff | ScalarDiff} + [TKE SST src | SDR SST src]? in one sweep.
// The phases are executed thread by thread with the kernel's own control flow (assemble_slist_box.cu), the CTA barriers
// being the boundaries between the loops, and the assembled CSR values / rhs are compared with the oracle's literal
// restatement of the reference kernels run as ONE kernel list on the same mesh and state.  Covers 1..3 slab ranks, the
// option combinations of every member of the list, several z chunkings, overwrite and accumulate modes, and a
// stretched-z (ABL-style) mesh.  Test scaffolding.
#define __host__
#define __device__
#define __forceinline__ inline
#include <cmath>
#include <cstdint>
#include "../../nalu_b200/csrc/slist_tile.cuh"
#include "../../nalu_b200/csrc/graph.h"
#include "../../oracle/nalu_oracle.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace nalu_b200;
using slist::SlistParams;
using CC = slist::Cfg;

struct RankMesh {
  int nx, ny, nz, k0;
  std::vector<int64_t> gid;
  std::vector<int32_t> owner;
  std::vector<int32_t> conn;
  std::vector<double> coords, vel, rho, rho_n, rho_nm1, q, q_n, q_nm1, gam, gjq, other, tvisc, f1, mdot;
};

static double hash01(uint64_t x)
{
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  x ^= x >> 31;
  return (double)(x >> 11) / (double)(1ull << 53);
}

static RankMesh make_rank(int nx, int ny, int nzg, int rank, int nranks, bool stretched)
{
  RankMesh m;
  const int base = nzg / nranks, rem = nzg % nranks;
  int k0 = 0;
  for (int r = 0; r < rank; ++r) k0 += base + (r < rem);
  const int lz = base + (rank < rem);
  m.nx = nx; m.ny = ny; m.nz = lz; m.k0 = k0;
  const int NX = nx + 1, NY = ny + 1, NZ = lz + 1;
  const int64_t sxy = (int64_t)NX * NY;
  for (int z = 0; z < NZ; ++z)
    for (int y = 0; y < NY; ++y)
      for (int x = 0; x < NX; ++x) {
        const int64_t g = 1 + x + NX * (int64_t)y + sxy * (z + k0);
        m.gid.push_back(g);
        m.owner.push_back((rank > 0 && z == 0) ? rank - 1 : rank);
        // ABL-style: geometric stretching in z (ratio 1.15), cell aspect ratio up to ~10
        const double zg = z + k0;
        const double zs = stretched ? 0.1 * (pow(1.15, zg) - 1.0) / 0.15 : zg;
        const double hz = stretched ? 0.1 * pow(1.15, zg > 0 ? zg - 1 : 0) : 1.0;
        const double c[3] = {x + 0.24 * (hash01(3 * g) - 0.5), y + 0.24 * (hash01(3 * g + 1) - 0.5),
                             zs + 0.24 * hz * (hash01(3 * g + 2) - 0.5)};
        for (int d = 0; d < 3; ++d) m.coords.push_back(c[d]);
        m.vel.push_back(-cos(0.9 * c[0]) * sin(0.8 * c[1]) + 0.1 * c[2]);
        m.vel.push_back(sin(0.7 * c[0]) * cos(0.9 * c[1]));
        m.vel.push_back(0.3 * sin(0.5 * c[2] + c[0]));
        const double rho = 1.0 + 0.2 * cos(0.4 * c[0] + 0.3 * c[1]);
        m.rho.push_back(rho);
        m.rho_n.push_back(0.97 * rho + 0.01);
        m.rho_nm1.push_back(0.95 * rho - 0.01);
        const double q = 0.5 + 0.4 * cos(0.6 * c[0]) * sin(0.5 * c[1] + 0.2) + 0.05 * sin(0.4 * c[2]);
        m.q.push_back(q);
        m.q_n.push_back(0.9 * q + 0.02);
        m.q_nm1.push_back(0.85 * q - 0.01);
        m.gam.push_back(2.0e-2 * (1.0 + 0.3 * sin(c[0]) * cos(c[1] + c[2])));
        // a projected nodal gradient close to, but not equal to, grad q (as the reference's Gjq is)
        m.gjq.push_back(-0.24 * sin(0.6 * c[0]) * sin(0.5 * c[1] + 0.2) + 0.01);
        m.gjq.push_back(0.2 * cos(0.6 * c[0]) * cos(0.5 * c[1] + 0.2) - 0.02);
        m.gjq.push_back(0.02 * cos(0.4 * c[2]) + 0.005);
        m.other.push_back(1.2 + 0.5 * sin(0.5 * c[0] + 0.3) * cos(0.7 * c[1]) + 0.1 * cos(0.3 * c[2]));
        m.tvisc.push_back(0.05 * (1.0 + 0.4 * cos(0.8 * c[0] - 0.2 * c[1])));
        m.f1.push_back(0.5 + 0.45 * sin(0.9 * c[0] + 0.4 * c[1] + 0.1 * c[2]));
      }
  for (int k = 0; k < lz; ++k)
    for (int j = 0; j < ny; ++j)
      for (int i = 0; i < nx; ++i) {
        const int n0 = i + NX * (j + NY * k);
        const int s = (int)sxy;
        const int c[8] = {n0, n0 + 1, n0 + 1 + NX, n0 + NX, n0 + s, n0 + 1 + s, n0 + 1 + NX + s, n0 + NX + s};
        m.conn.insert(m.conn.end(), c, c + 8);
      }
  m.mdot.resize((size_t)nx * ny * lz * 12);
  orc_calc_mass_flow_rate_scs((int64_t)nx * ny * lz, m.conn.data(), m.coords.data(), m.rho.data(), m.vel.data(),
                              m.mdot.data());
  return m;
}

// assemble_slist_box_kernel, one CTA at a time
template <bool SH, bool SK>
static void emulate(const SlistParams &P0, int z_chunk)
{
  SlistParams P = P0;
  P.tiles_x = (P.NX + CC::TBX - 1) / CC::TBX;
  P.tiles_y = (P.NY + CC::TBY - 1) / CC::TBY;
  P.z_chunk = z_chunk;
  const int ntile = P.tiles_x * P.tiles_y;
  const int nplanes = P.z_last - P.z_first;
  const int zchunks = (nplanes + z_chunk - 1) / z_chunk;
  std::vector<double> smv(CC::SMEM_DOUBLES);
  double *sm = smv.data();
  double *ring = sm + CC::OFF_RING;
  struct Th { double accA[27], accB[27], rA, rB, volA, volB; };
  std::vector<Th> th(CC::NT);
  const bool scv_pass = P.mass || P.src;
  const bool src_lumped = P.src_lumped != 0;
  const slist::UpwK K = {P.alpha, P.alpha_upw};
  for (int cta = 0; cta < ntile * zchunks; ++cta) {
    for (int i = 0; i < CC::SMEM_DOUBLES; ++i) sm[i] = NAN; // uninitialised shared memory must never be consumed
    const int tile = cta % ntile, zc = cta / ntile;
    const int x0 = (tile % P.tiles_x) * CC::TBX, y0 = (tile / P.tiles_x) * CC::TBY;
    const int zs = P.z_first + zc * P.z_chunk;
    const int ze = (zs + P.z_chunk < P.z_last) ? zs + P.z_chunk : P.z_last;
    for (auto &t : th) {
      for (int k = 0; k < 27; ++k) t.accA[k] = t.accB[k] = 0.0;
      t.rA = t.rB = t.volA = t.volB = 0.0;
    }
    for (int tid = 0; tid < CC::NT; ++tid) {
      slist::stage_plane(P, ring + ((zs + 1) & 1) * CC::PLN, tid, x0, y0, zs - 1);
      slist::stage_plane(P, ring + (zs & 1) * CC::PLN, tid, x0, y0, zs);
    }
    // ---- barrier
    for (int ez = zs - 1; ez < ze; ++ez) {
      const double *pl0 = ring + ((ez + 2) & 1) * CC::PLN, *pl1 = ring + ((ez + 1) & 1) * CC::PLN;
      if (scv_pass) {
        for (int tid = 0; tid < CC::NT; ++tid) slist::p1b_element(P, sm, tid, x0, y0, ez, pl0, pl1);
        // ---- barrier
        for (int tid = 0; tid < CC::NT; ++tid) {
          const int lx = tid % CC::ELX, ly = tid / CC::ELX;
          Th &t = th[tid];
          if (lx < CC::TBX && ly < CC::TBY) slist::p2b_layer(sm, lx, ly, src_lumped, t.accA, t.rA, t.volA, t.accB, t.rB, t.volB);
        }
        // ---- barrier
      }
      for (int tid = 0; tid < CC::NT; ++tid) slist::p1a_element(P, sm, tid, x0, y0, ez, pl0, pl1);
      // ---- barrier
      // on the device the cp.async of plane ez+2 lands (asynchronously) in the slot of plane ez while phase 2a runs;
      // phase 2a reads the records only, so running it first here is the same schedule
      for (int tid = 0; tid < CC::NT; ++tid) {
        const int lx = tid % CC::ELX, ly = tid / CC::ELX;
        const bool in_tile = lx < CC::TBX && ly < CC::TBY;
        const bool active = in_tile && x0 + lx < P.NX && y0 + ly < P.NY;
        Th &t = th[tid];
        if (in_tile) slist::p2a_layer<SH, SK>(K, sm, lx, ly, t.accA, t.rA, t.accB, t.rB);
        if (active && ez >= zs) {
          if (P.mass) slist::mass_apply(P, x0 + lx, y0 + ly, ez, t.volA, t.accA[13], t.rA);
          cont::write_row(P, x0 + lx, y0 + ly, ez, t.accA, t.rA);
        }
        for (int k = 0; k < 27; ++k) {
          t.accA[k] = t.accB[k];
          t.accB[k] = 0.0;
        }
        t.rA = t.rB; t.rB = 0.0;
        t.volA = t.volB; t.volB = 0.0;
      }
      for (int tid = 0; tid < CC::NT; ++tid) slist::stage_plane(P, ring + ((ez + 2) & 1) * CC::PLN, tid, x0, y0, ez + 2);
      // ---- barrier
    }
  }
}

static int exchange_and_finalize(std::vector<Graph> &gs, std::vector<std::vector<uint8_t>> &slot27)
{
  std::string err;
  const int R = (int)gs.size();
  for (int r = 0; r < R; ++r)
    if (graph_number_and_adjacency(gs[r], err) || graph_build_send_messages(gs[r], err)) {
      printf("graph: %s\n", err.c_str());
      return 1;
    }
  for (int r = 0; r < R; ++r)
    for (int p = 0; p < R; ++p)
      if (p != r) gs[p].recv_msg[r] = gs[r].send_msg[p];
  slot27.resize(R);
  for (int r = 0; r < R; ++r) {
    std::vector<uint8_t> elem_slot;
    if (graph_finalize_host(gs[r], elem_slot, slot27[r], err)) {
      printf("finalize: %s\n", err.c_str());
      return 1;
    }
  }
  return 0;
}

// adv: 0 ScalarDiff, 1 ScalarAdvDiff, 2 ScalarUpwAdvDiff;  src: 0 none, 1 tke, 2 sdr
struct Case {
  int nx, ny, nz, R, zc;
  int adv, adv_opt; // NALU_OPT_SHIFTED_GRAD_OP | SKEW_SYMMETRIC | USE_LIMITER | PECLET_TANH
  int mass, src, src_opt; // NALU_OPT_LUMPED_MASS | SHIFTED_GRAD_OP
  int overwrite, stretched;
};

static double run_case(const Case &c)
{
  const int R = c.R;
  std::vector<RankMesh> ms;
  std::vector<Graph> gs(R);
  for (int r = 0; r < R; ++r) {
    ms.push_back(make_rank(c.nx, c.ny, c.nz, r, R, c.stretched));
    Graph &g = gs[r];
    g.rank = r; g.nranks = R; g.ndof = 1;
    g.n_nodes = (int)ms[r].gid.size();
    g.n_elems = (int64_t)ms[r].conn.size() / 8;
    g.box[0] = c.nx; g.box[1] = c.ny; g.box[2] = ms[r].nz;
    g.gid = ms[r].gid; g.owner = ms[r].owner; g.elem_nodes = ms[r].conn;
  }
  std::vector<std::vector<uint8_t>> slot27;
  if (exchange_and_finalize(gs, slot27)) return 1.0;

  orc_linsys *ls = orc_linsys_create(R, 1);
  for (int r = 0; r < R; ++r)
    orc_linsys_set_rank_mesh(ls, r, gs[r].n_nodes, ms[r].gid.data(), ms[r].owner.data(), gs[r].n_elems, ms[r].conn.data());
  orc_linsys_build_elem_to_node_graph(ls);
  orc_linsys_finalize(ls);
  orc_linsys_zero(ls);

  const double dt = 0.05, g1 = 1.5, g2 = -2.0, g3 = 0.5;
  const double alpha = 0.3, alpha_upw = 0.8, ho_upwind = 0.9, hybrid = 1.0, tanh_trans = 2.0, tanh_width = 4.0;
  const double turb[10] = {0.09, 10.0, 0.856, 0.075, 0.0828, 5.0 / 9.0, 0.44, 0, 0, 0};
  double worst = 0.0;
  for (int r = 0; r < R; ++r) {
    const RankMesh &m = ms[r];
    orc_kernel_desc o[3];
    int nk = 0;
    memset(o, 0, sizeof(o));
    if (c.mass) {
      orc_kernel_desc &k = o[nk++];
      k.kind = ORC_SCALAR_MASS;
      k.lumpedMass = 1;
      k.dt = dt; k.gamma1 = g1; k.gamma2 = g2; k.gamma3 = g3;
      k.coordinates = m.coords.data();
      k.scalarQ = m.q.data(); k.scalarQN = m.q_n.data(); k.scalarQNm1 = m.q_nm1.data();
      k.density = m.rho.data(); k.densityN = m.rho_n.data(); k.densityNm1 = m.rho_nm1.data();
    }
    {
      orc_kernel_desc &k = o[nk++];
      k.kind = c.adv == 2 ? ORC_SCALAR_UPW_ADV_DIFF : (c.adv == 1 ? ORC_SCALAR_ADV_DIFF : ORC_SCALAR_DIFF);
      k.shiftedGradOp = (c.adv_opt & NALU_OPT_SHIFTED_GRAD_OP) != 0;
      k.skewSymmetric = (c.adv_opt & NALU_OPT_SKEW_SYMMETRIC) != 0;
      k.useLimiter = (c.adv_opt & NALU_OPT_USE_LIMITER) != 0;
      k.pecletTanh = (c.adv_opt & NALU_OPT_PECLET_TANH) != 0;
      k.om_densFac = 1; k.dt = dt; k.gamma1 = g1; k.gamma2 = g2; k.gamma3 = g3;
      k.coordinates = m.coords.data();
      k.scalarQ = m.q.data();
      k.viscosity = m.gam.data();
      k.massFlowRate = m.mdot.data();
      k.velocity = m.vel.data();
      k.density = m.rho.data();
      k.Gjq = m.gjq.data();
      k.alpha = alpha; k.alphaUpw = alpha_upw; k.hoUpwind = ho_upwind;
      k.hybridFactor = hybrid; k.tanhTrans = tanh_trans; k.tanhWidth = tanh_width;
    }
    if (c.src) {
      orc_kernel_desc &k = o[nk++];
      k.kind = c.src == 1 ? ORC_TKE_SST_SRC : ORC_SDR_SST_SRC;
      k.lumpedMass = (c.src_opt & NALU_OPT_LUMPED_MASS) != 0;
      k.shiftedGradOp = (c.src_opt & NALU_OPT_SHIFTED_GRAD_OP) != 0;
      k.coordinates = m.coords.data();
      k.tke = c.src == 1 ? m.q.data() : m.other.data();
      k.sdr = c.src == 1 ? m.other.data() : m.q.data();
      k.velocity = m.vel.data();
      k.density = m.rho.data();
      k.tvisc = m.tvisc.data();
      k.fOneBlend = m.f1.data();
      k.betaStar = turb[0]; k.tkeProdLimitRatio = turb[1]; k.sigmaWTwo = turb[2]; k.betaOne = turb[3]; k.betaTwo = turb[4];
      k.gammaOne = turb[5]; k.gammaTwo = turb[6];
    }
    orc_assemble_elem(ls, r, o, nk);

    const Graph &g = gs[r];
    std::vector<long long> bp(g.browptr.begin(), g.browptr.end());
    SlistParams P;
    memset(&P, 0, sizeof(P));
    P.nx = c.nx; P.ny = c.ny; P.nz = m.nz;
    P.NX = c.nx + 1; P.NY = c.ny + 1; P.NZ = m.nz + 1;
    P.coords = m.coords.data(); P.q = m.q.data(); P.gam = m.gam.data();
    P.mdot = c.adv ? m.mdot.data() : nullptr;
    P.options = c.adv_opt & (NALU_OPT_SHIFTED_GRAD_OP | NALU_OPT_SKEW_SYMMETRIC);
    if (c.adv == 2) {
      P.upw = 1;
      P.upw_options = c.adv_opt & (NALU_OPT_USE_LIMITER | NALU_OPT_PECLET_TANH);
      P.vel = m.vel.data(); P.rho = m.rho.data(); P.gjq = m.gjq.data();
      P.alpha = alpha; P.alpha_upw = alpha_upw; P.ho_upwind = ho_upwind; P.hybrid = hybrid;
      P.tanh_trans = tanh_trans; P.tanh_width = tanh_width;
    }
    if (c.mass) {
      P.mass = 1;
      P.rho = m.rho.data(); P.q_n = m.q_n.data(); P.q_nm1 = m.q_nm1.data(); P.rho_n = m.rho_n.data(); P.rho_nm1 = m.rho_nm1.data();
      P.g1_dt = g1 / dt; P.g2_dt = g2 / dt; P.g3_dt = g3 / dt;
    }
    if (c.src) {
      P.src = c.src;
      P.src_lumped = (c.src_opt & NALU_OPT_LUMPED_MASS) != 0;
      P.src_shifted = (c.src_opt & NALU_OPT_SHIFTED_GRAD_OP) != 0;
      P.other = m.other.data(); P.vel = m.vel.data(); P.rho = m.rho.data(); P.tvisc = m.tvisc.data();
      P.f1 = c.src == 2 ? m.f1.data() : nullptr;
      for (int i = 0; i < 10; ++i) P.turb[i] = turb[i];
    }
    P.node_row = g.node_row.data(); P.browptr = bp.data(); P.slot27 = slot27[r].data();
    int errflag = 0;
    P.err = &errflag;
    P.overwrite = c.overwrite;
    P.z_first = 0; P.z_last = P.NZ;
    const size_t nv = (size_t)(g.nnz_owned() + g.nnz_shared()), nr = (size_t)g.n_row_nodes();
    std::vector<double> values(nv, c.overwrite ? 777.0 : 0.0), rhs(nr, c.overwrite ? -55.0 : 0.0);
    P.values = values.data(); P.rhs = rhs.data();
    const bool sh = P.options & NALU_OPT_SHIFTED_GRAD_OP, sk = P.options & NALU_OPT_SKEW_SYMMETRIC;
    if (sh && sk) emulate<true, true>(P, c.zc);
    else if (sh) emulate<true, false>(P, c.zc);
    else if (sk) emulate<false, true>(P, c.zc);
    else emulate<false, false>(P, c.zc);
    if (errflag) { printf("error flag %d\n", errflag); return 1.0; }

    for (int shared = 0; shared < 2; ++shared) {
      const int nrows = shared ? orc_linsys_num_shared_rows(ls, r) : orc_linsys_num_owned_rows(ls, r);
      const size_t *rp = orc_linsys_rowptr(ls, r, shared);
      const double *ov = orc_linsys_values(ls, r, shared), *orr = orc_linsys_rhs(ls, r, shared);
      const double *v = values.data() + (shared ? g.nnz_owned() : 0);
      const double *rr = rhs.data() + (shared ? (size_t)g.n_owned_nodes : 0);
      if ((int64_t)rp[nrows] != (shared ? g.nnz_shared() : g.nnz_owned())) {
        printf("nnz mismatch rank %d shared %d\n", r, shared);
        return 1.0;
      }
      // per-row scale (stretched meshes span orders of magnitude)
      for (int q = 0; q < nrows; ++q) {
        double vs = 0;
        for (size_t p = rp[q]; p < rp[q + 1]; ++p) vs = fmax(vs, fabs(ov[p]));
        double absrow = 0;
        for (size_t p = rp[q]; p < rp[q + 1]; ++p) {
          const double e = fabs(ov[p] - v[p]) / vs;
          if (!(e <= worst)) worst = std::isnan(e) ? 1.0 : e;
          absrow += fabs(ov[p]);
        }
        // the residual is a difference of O(|row| * |q|) terms
        const double e = fabs(orr[q] - rr[q]) / fmax(fabs(orr[q]), absrow);
        if (!(e <= worst)) worst = std::isnan(e) ? 1.0 : e;
      }
    }
  }
  orc_linsys_destroy(ls);
  return worst;
}

int main()
{
  const int SH = NALU_OPT_SHIFTED_GRAD_OP, SK = NALU_OPT_SKEW_SYMMETRIC, LIM = NALU_OPT_USE_LIMITER, TANH = NALU_OPT_PECLET_TANH,
            LM = NALU_OPT_LUMPED_MASS;
  const Case cases[] = {
      // nx ny nz R zc | adv opt | mass src src_opt | ow stretched
      {1, 1, 1, 1, 16, 2, 0, 0, 0, 0, 1, 0},                   // upwind alone
      {3, 2, 2, 1, 16, 2, LIM, 1, 0, 0, 1, 0},                 // mass + limited upwind
      {4, 3, 3, 1, 2, 2, LIM | TANH | SH, 1, 1, LM, 1, 0},     // config-4 tke list
      {4, 3, 3, 1, 16, 2, LIM | SK, 1, 2, LM, 1, 0},           // config-4 sdr list
      {17, 9, 3, 1, 3, 2, 0, 1, 1, 0, 1, 0},                   // consistent source, two tiles in x and y
      {5, 4, 4, 2, 16, 2, SH | SK, 1, 2, SH, 1, 0},            // shifted SCV grad-op, 2 ranks
      {6, 5, 7, 3, 2, 1, 0, 1, 1, LM | SH, 0, 0},              // central adv-diff + mass + tke, accumulate, 3 ranks
      {5, 5, 3, 1, 16, 0, SH, 1, 0, 0, 1, 0},                  // ScalarDiff + mass
      {4, 4, 3, 1, 16, 1, SK, 0, 2, LM, 1, 0},                 // adv-diff + sdr, no mass
      {6, 4, 12, 2, 5, 2, LIM, 1, 1, LM, 1, 1},                // stretched z, tke list
      {6, 4, 12, 1, 16, 2, LIM, 1, 2, LM, 1, 1},               // stretched z, sdr list
  };
  double worst = 0;
  for (const Case &c : cases) {
    const double e = run_case(c);
    printf("slist %dx%dx%d ranks %d zchunk %d adv %d opt %d mass %d src %d srcopt %d overwrite %d stretched %d: relerr %.3e\n", c.nx,
           c.ny, c.nz, c.R, c.zc, c.adv, c.adv_opt, c.mass, c.src, c.src_opt, c.overwrite, c.stretched, e);
    worst = fmax(worst, e);
  }
  printf("worst %.3e\n", worst);
  if (!(worst < 1e-12)) {
    printf("FAIL\n");
    return 1;
  }
  printf("PASS\n");
  return 0;
}
