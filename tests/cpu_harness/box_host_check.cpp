// CPU emulation of the tiled write-once kernel (nalu_b200/csrc/box_tile.cuh compiled as plain
// C++): the phases are executed thread by thread with the kernel's own control flow
// (assemble_box.cu), the CTA barriers being the boundaries between the loops, and the assembled
// CSR values / rhs are compared with the oracle's literal restatement of the reference on the
// same mesh and state.  Covers 1 rank and both ranks of a 2-slab decomposition (shared-not-owned
// rows, remote columns), several tile shapes and z chunkings.  Test scaffolding.
#define __host__
#define __device__
#define __forceinline__ inline
#include "../../nalu_b200/csrc/box_tile.cuh"
#include "../../nalu_b200/csrc/graph.h"
#include "../../oracle/nalu_oracle.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace nalu_b200;
using box::BoxParams;

struct RankMesh {
  int nx, ny, nz, k0;
  std::vector<int64_t> gid;
  std::vector<int32_t> owner;
  std::vector<int32_t> conn;
  std::vector<double> coords, vel, visc, rho, mdot;
};

static double hash01(uint64_t x)
{
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  x ^= x >> 31;
  return (double)(x >> 11) / (double)(1ull << 53);
}

static RankMesh make_rank(int nx, int ny, int nzg, int rank, int nranks)
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
        const double c[3] = {x + 0.24 * (hash01(3 * g) - 0.5), y + 0.24 * (hash01(3 * g + 1) - 0.5),
                             z + k0 + 0.24 * (hash01(3 * g + 2) - 0.5)};
        for (int d = 0; d < 3; ++d) m.coords.push_back(c[d]);
        m.vel.push_back(-cos(0.9 * c[0]) * sin(0.8 * c[1]) + 0.1 * c[2]);
        m.vel.push_back(sin(0.7 * c[0]) * cos(0.9 * c[1]));
        m.vel.push_back(0.3 * sin(0.5 * c[2] + c[0]));
        m.visc.push_back(0.1 * (1.0 + 0.3 * sin(c[0]) * cos(c[1] + c[2])));
        m.rho.push_back(1.0 + 0.2 * cos(0.4 * c[0] + 0.3 * c[1]));
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

template <class C>
static void emulate(const BoxParams &P0, int z_chunk, double *values, double *rhs)
{
  BoxParams P = P0;
  P.values = values;
  P.rhs = rhs;
  P.tiles_x = (P.NX + C::TBX - 1) / C::TBX;
  P.tiles_y = (P.NY + C::TBY - 1) / C::TBY;
  P.z_chunk = z_chunk;
  const int zchunks = (P.NZ + z_chunk - 1) / z_chunk;
  constexpr int PLN = C::RN * C::NF;
  constexpr int NT = C::NT;
  std::vector<double> sm(C::SMEM_DOUBLES);
  std::vector<box::P2State> S(NT);
  std::vector<box::NodeCtx> N(NT);
  std::vector<int> ncov(NT), row_n(NT);
  std::vector<long long> b0_n(NT), b1_n(NT);
  std::vector<unsigned> sm_n(NT), sf_n(NT);
  struct Regs { double p[C::NPL]; double m[C::NML]; };
  std::vector<Regs> R(NT);
  const bool divu = P.include_divu != 0.0;
  const long long plane_nodes = (long long)P.NX * P.NY;
  for (int cta = 0; cta < P.tiles_x * P.tiles_y * zchunks; ++cta) {
    for (auto &v : sm) v = NAN; // uninitialised shared memory must never be consumed
    double *tab = sm.data() + C::OFF_TAB, *ring = sm.data() + C::OFF_RING, *mdb = sm.data() + C::OFF_MD;
    long long *s_rb = reinterpret_cast<long long *>(sm.data() + C::OFF_RB);
    int *s_len = reinterpret_cast<int *>(sm.data() + C::OFF_LEN);
    const int ntile = P.tiles_x * P.tiles_y;
    const int tile = cta % ntile, zc = cta / ntile;
    const int x0 = (tile % P.tiles_x) * C::TBX, y0 = (tile / P.tiles_x) * C::TBY;
    const int zs = zc * P.z_chunk;
    const int ze = (zs + P.z_chunk < P.NZ) ? zs + P.z_chunk : P.NZ;
    for (int i = 0; i < 192; ++i) box::build_table_entry<C>(i, P.options, tab);
    for (int pl = zs - 1; pl <= zs + 1; ++pl)
      for (int tid = 0; tid < NT; ++tid) {
        box::plane_load<C>(P, tid, x0, y0, pl, R[tid].p);
        box::plane_store<C>(tid, R[tid].p, ring + ((pl + 3) % 3) * PLN);
      }
    for (int tid = 0; tid < NT; ++tid) {
      box::mdot_load<C>(P, tid, x0, y0, zs - 1, R[tid].m);
      box::mdot_store<C>(tid, R[tid].m, mdb);
    }
    if (zs >= 1)
      for (int it = 0; it < C::NITEM; ++it)
        box::phase1_item<C>(P, it, x0, y0, ring + ((zs - 1) % 3) * PLN, ring + (zs % 3) * PLN, mdb,
                            sm.data() + ((zs - 1) & 1) * C::LBUF);
    for (int tid = 0; tid < NT; ++tid) {
      box::mdot_load<C>(P, tid, x0, y0, zs, R[tid].m);
      box::mdot_store<C>(tid, R[tid].m, mdb);
      box::NodeCtx &n = N[tid];
      box::node_ctx_init<C>(P, n, tid, x0, y0);
      n.n = n.x + (long long)P.NX * (n.y + (long long)P.NY * zs);
      n.row = 0; n.nb = 0; n.rowbase = 0; n.slots_mid = n.slots_fin = 0xFFFFFFFFu;
      if (n.active) {
        n.row = P.node_row[n.n];
        const long long b0 = P.browptr[n.row];
        n.nb = (int)(P.browptr[n.row + 1] - b0);
        n.rowbase = 9 * b0;
        n.slots_mid = box::load_slots(P.slot27 + n.n * 27, n.cls, -1);
        n.slots_fin = box::load_slots(P.slot27 + n.n * 27, n.cls, (n.cls & 4) ? 1 : 0);
      }
    }
    for (int z = zs; z < ze; ++z) {
      const bool more = z + 1 < ze;
      for (int tid = 0; tid < NT; ++tid) {
        box::plane_load<C>(P, tid, x0, y0, z + 2, R[tid].p);
        box::mdot_load<C>(P, tid, x0, y0, z + 1, R[tid].m);
        row_n[tid] = (N[tid].active && more) ? P.node_row[N[tid].n + plane_nodes] : 0;
      }
      const double *pl_m = ring + ((z + 2) % 3) * PLN, *pl_0 = ring + (z % 3) * PLN, *pl_p = ring + ((z + 1) % 3) * PLN;
      if (z < P.nz)
        for (int it = 0; it < C::NITEM; ++it) box::phase1_item<C>(P, it, x0, y0, pl_0, pl_p, mdb, sm.data() + (z & 1) * C::LBUF);
      // ---- barrier S1
      const double *lower = sm.data() + ((z + 1) & 1) * C::LBUF;
      const double *upper = sm.data() + (z & 1) * C::LBUF;
      double *stg = sm.data() + ((z + 1) & 1) * C::LBUF;
      for (int tid = 0; tid < NT; ++tid) {
        box::mdot_store<C>(tid, R[tid].m, mdb);
        box::NodeCtx &n = N[tid];
        const int nl = tid >> 3;
        if (n.cls == 0) {
          s_rb[nl] = n.rowbase;
          s_len[nl] = n.active ? n.nb * 9 : 0;
        }
        b0_n[tid] = b1_n[tid] = 0;
        sm_n[tid] = sf_n[tid] = 0xFFFFFFFFu;
        if (n.active && more) {
          b0_n[tid] = P.browptr[row_n[tid]];
          b1_n[tid] = P.browptr[row_n[tid] + 1];
          const unsigned char *sl = P.slot27 + (n.n + plane_nodes) * 27;
          sm_n[tid] = box::load_slots(sl, n.cls, -1);
          sf_n[tid] = box::load_slots(sl, n.cls, (n.cls & 4) ? 1 : 0);
        }
        box::p2_begin(S[tid]);
        n.z = z;
        if (n.active) {
          ncov[tid] = ((n.x > 0) + 1 + (n.x < P.NX - 1)) * ((n.y > 0) + 1 + (n.y < P.NY - 1)) * ((z > 0) + 1 + (z < P.NZ - 1));
          box::p2_pass<C, 0>(P, n, S[tid], lower, tab + 2 * n.cls, divu);
        }
      }
      // ---- barrier S2: the lower layer's buffer becomes the staging area
      for (int q = 0; q < C::LBUF; ++q) stg[q] = NAN;
      for (int tid = 0; tid < NT; ++tid) {
        box::NodeCtx &n = N[tid];
        if (n.active && n.nb != ncov[tid]) {
          double *row = stg + (tid >> 3) * C::SROW;
          for (int q = n.cls; q < n.nb * 9; q += 8) row[q] = 0.0;
        }
      }
      for (int tid = 0; tid < NT; ++tid) {
        box::NodeCtx &n = N[tid];
        if (!n.active) continue;
        const int dz_fin = (n.cls & 4) ? 1 : 0;
        if (n.cls & 4) {
          box::p2_flush<C>(n, S[tid], n.slots_mid, pl_m, stg);
          box::p2_zero(S[tid]);
        }
        box::p2_pass<C, 1>(P, n, S[tid], upper, tab + 2 * n.cls, divu);
        box::p2_flush<C>(n, S[tid], n.slots_fin, dz_fin ? pl_p : pl_0, stg);
      }
      for (int nl = 0; nl < C::NNODE; ++nl) {
        const box::NodeCtx &n = N[nl * 8];
        if (!n.active) continue;
        for (int i = 0; i < 3; ++i) {
          double v = 0;
          for (int c = 0; c < 8; ++c) v += S[nl * 8 + c].rr[i] + (c == i ? S[nl * 8 + c].dv : 0.0);
          double *r = P.rhs + (long long)n.row * 3 + i;
          *r = P.overwrite ? v : *r + v;
        }
      }
      // ---- barrier S3
      for (int tid = 0; tid < NT; ++tid) box::plane_store<C>(tid, R[tid].p, ring + ((z + 2) % 3) * PLN);
      for (int m = 0; m < C::NNODE; ++m) {
        const int len = s_len[m];
        double *dst = P.values + s_rb[m];
        const double *src = stg + m * C::SROW;
        for (int q = 0; q < len; ++q) dst[q] = P.overwrite ? src[q] : dst[q] + src[q];
      }
      // ---- barrier S4
      for (int tid = 0; tid < NT; ++tid) {
        box::NodeCtx &n = N[tid];
        n.n += plane_nodes;
        n.row = row_n[tid];
        n.nb = (int)(b1_n[tid] - b0_n[tid]);
        n.rowbase = 9 * b0_n[tid];
        n.slots_mid = sm_n[tid];
        n.slots_fin = sf_n[tid];
      }
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

template <class C>
static double run_case(int nx, int ny, int nzg, int R, int z_chunk, int options, double include_divu, int overwrite)
{
  std::vector<RankMesh> ms;
  std::vector<Graph> gs(R);
  for (int r = 0; r < R; ++r) {
    ms.push_back(make_rank(nx, ny, nzg, r, R));
    Graph &g = gs[r];
    g.rank = r; g.nranks = R; g.ndof = 3;
    g.n_nodes = (int)ms[r].gid.size();
    g.n_elems = (int64_t)ms[r].conn.size() / 8;
    g.box[0] = nx; g.box[1] = ny; g.box[2] = ms[r].nz;
    g.gid = ms[r].gid; g.owner = ms[r].owner; g.elem_nodes = ms[r].conn;
  }
  std::vector<std::vector<uint8_t>> slot27;
  if (exchange_and_finalize(gs, slot27)) return 1.0;

  orc_linsys *ls = orc_linsys_create(R, 3);
  for (int r = 0; r < R; ++r)
    orc_linsys_set_rank_mesh(ls, r, gs[r].n_nodes, ms[r].gid.data(), ms[r].owner.data(), gs[r].n_elems, ms[r].conn.data());
  orc_linsys_build_elem_to_node_graph(ls);
  orc_linsys_finalize(ls);
  orc_linsys_zero(ls);

  double worst = 0.0;
  for (int r = 0; r < R; ++r) {
    orc_kernel_desc o;
    memset(&o, 0, sizeof(o));
    o.kind = ORC_MOMENTUM_ADV_DIFF;
    o.shiftedGradOp = (options & NALU_OPT_SHIFTED_GRAD_OP) != 0;
    o.skewSymmetric = (options & NALU_OPT_SKEW_SYMMETRIC) != 0;
    o.includeDivU = include_divu;
    o.om_densFac = 1; o.dt = 1; o.gamma1 = 1; o.gamma2 = -1;
    o.coordinates = ms[r].coords.data();
    o.velocity = ms[r].vel.data();
    o.viscosity = ms[r].visc.data();
    o.massFlowRate = ms[r].mdot.data();
    orc_assemble_elem(ls, r, &o, 1);

    const Graph &g = gs[r];
    std::vector<long long> bp(g.browptr.begin(), g.browptr.end());
    BoxParams P;
    memset(&P, 0, sizeof(P));
    P.nx = nx; P.ny = ny; P.nz = ms[r].nz;
    P.NX = nx + 1; P.NY = ny + 1; P.NZ = ms[r].nz + 1;
    P.coords = ms[r].coords.data(); P.vel = ms[r].vel.data(); P.visc = ms[r].visc.data(); P.mdot = ms[r].mdot.data();
    P.node_row = g.node_row.data(); P.browptr = bp.data(); P.slot27 = slot27[r].data();
    int errflag = 0;
    P.err = &errflag;
    P.include_divu = include_divu; P.options = options; P.overwrite = overwrite;
    const size_t nv = (size_t)(g.nnz_owned() + g.nnz_shared()), nr = (size_t)g.n_row_nodes() * 3;
    // overwrite mode must not depend on what the arrays held; accumulate mode starts from zero
    std::vector<double> values(nv, overwrite ? 777.0 : 0.0), rhs(nr, overwrite ? -55.0 : 0.0);
    emulate<C>(P, z_chunk, values.data(), rhs.data());
    if (errflag) return 1.0;

    for (int shared = 0; shared < 2; ++shared) {
      const int nrows = shared ? orc_linsys_num_shared_rows(ls, r) : orc_linsys_num_owned_rows(ls, r);
      const size_t *rp = orc_linsys_rowptr(ls, r, shared);
      const double *ov = orc_linsys_values(ls, r, shared), *orr = orc_linsys_rhs(ls, r, shared);
      const double *v = values.data() + (shared ? g.nnz_owned() : 0);
      const double *rr = rhs.data() + (shared ? (size_t)g.n_owned_nodes * 3 : 0);
      if ((int64_t)rp[nrows] != (shared ? g.nnz_shared() : g.nnz_owned())) {
        printf("nnz mismatch rank %d shared %d\n", r, shared);
        return 1.0;
      }
      double vs = 0, rs = 0;
      for (size_t q = 0; q < rp[nrows]; ++q) vs = fmax(vs, fabs(ov[q]));
      for (int q = 0; q < nrows; ++q) rs = fmax(rs, fabs(orr[q]));
      for (size_t q = 0; q < rp[nrows]; ++q) {
        const double e = fabs(ov[q] - v[q]) / vs;
        if (!(e <= worst)) worst = std::isnan(e) ? 1.0 : e;
      }
      for (int q = 0; q < nrows; ++q) {
        const double e = fabs(orr[q] - rr[q]) / rs;
        if (!(e <= worst)) worst = std::isnan(e) ? 1.0 : e;
      }
    }
  }
  orc_linsys_destroy(ls);
  return worst;
}

int main()
{
  double worst = 0;
  struct Case { int nx, ny, nz, R, zc, opt; double divu; int ow; };
  const Case cases[] = {
      {1, 1, 1, 1, 16, 0, 0.0, 1},  {2, 3, 2, 1, 16, 0, 1.0, 1},  {7, 5, 4, 1, 2, 0, 1.0, 1},
      {13, 9, 5, 1, 3, NALU_OPT_SHIFTED_GRAD_OP | NALU_OPT_SKEW_SYMMETRIC, 1.0, 1},
      {6, 4, 6, 2, 16, 0, 1.0, 1},  {9, 7, 5, 2, 2, NALU_OPT_SKEW_SYMMETRIC, 0.0, 0},
      {5, 5, 7, 3, 1, NALU_OPT_SHIFTED_GRAD_OP, 1.0, 1},
  };
  for (const Case &c : cases) {
    const double e64 = run_case<box::Cfg<6, 4>>(c.nx, c.ny, c.nz, c.R, c.zc, c.opt, c.divu, c.ow);
    const double e84 = run_case<box::Cfg<8, 4>>(c.nx, c.ny, c.nz, c.R, c.zc, c.opt, c.divu, c.ow);
    const double e86 = run_case<box::Cfg<8, 6>>(c.nx, c.ny, c.nz, c.R, c.zc, c.opt, c.divu, c.ow);
    printf("box %dx%dx%d ranks %d zchunk %d opt %d divu %.0f overwrite %d: relerr 6x4 %.3e  8x4 %.3e  8x6 %.3e\n", c.nx,
           c.ny, c.nz, c.R, c.zc, c.opt, c.divu, c.ow, e64, e84, e86);
    worst = fmax(worst, fmax(e64, fmax(e84, e86)));
  }
  printf("worst %.3e\n", worst);
  if (!(worst < 1e-12)) {
    printf("FAIL\n");
    return 1;
  }
  printf("PASS\n");
  return 0;
}
