// CPU emulation of the tiled write-once continuity kernel (nalu_b200/csrc/cont_tile.cuh compiled as plain
// C++): the phases are executed thread by thread with the kernel's own control flow (assemble_cont_box.cu),
// the CTA barriers being the boundaries between the loops, and the assembled CSR values / rhs are compared
// with the oracle's literal restatement of the reference on the same mesh and state.  Covers 1..3 slab ranks
// (shared-not-owned rows, remote columns), every option combination, several z chunkings, overwrite and
// accumulate modes.  Test scaffolding.
#define __host__
#define __device__
#define __forceinline__ inline
#include "../../nalu_b200/csrc/cont_tile.cuh"
#include "../../nalu_b200/csrc/graph.h"
#include "../../oracle/nalu_oracle.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace nalu_b200;
using cont::ContParams;
using CC = cont::Cfg;

struct RankMesh {
  int nx, ny, nz, k0;
  std::vector<int64_t> gid;
  std::vector<int32_t> owner;
  std::vector<int32_t> conn;
  std::vector<double> coords, vel, rho, p, gp;
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
        m.rho.push_back(1.0 + 0.2 * cos(0.4 * c[0] + 0.3 * c[1]));
        m.p.push_back(-0.25 * (cos(0.6 * c[0]) + cos(0.6 * c[1])) + 0.1 * sin(0.4 * c[2]));
        m.gp.push_back(0.15 * sin(0.6 * c[0]) + 0.01);
        m.gp.push_back(0.15 * sin(0.6 * c[1]) - 0.02);
        m.gp.push_back(0.04 * cos(0.4 * c[2]));
      }
  for (int k = 0; k < lz; ++k)
    for (int j = 0; j < ny; ++j)
      for (int i = 0; i < nx; ++i) {
        const int n0 = i + NX * (j + NY * k);
        const int s = (int)sxy;
        const int c[8] = {n0, n0 + 1, n0 + 1 + NX, n0 + NX, n0 + s, n0 + 1 + s, n0 + 1 + NX + s, n0 + NX + s};
        m.conn.insert(m.conn.end(), c, c + 8);
      }
  return m;
}

// assemble_cont_box_kernel, one CTA at a time
template <bool SH>
static void emulate(const ContParams &P0, int z_chunk)
{
  ContParams P = P0;
  P.tiles_x = (P.NX + CC::TBX - 1) / CC::TBX;
  P.tiles_y = (P.NY + CC::TBY - 1) / CC::TBY;
  P.z_chunk = z_chunk;
  const int ntile = P.tiles_x * P.tiles_y;
  const int nplanes = P.z_last - P.z_first;
  const int zchunks = (nplanes + z_chunk - 1) / z_chunk;
  std::vector<double> smv(CC::SMEM_DOUBLES);
  double *sm = smv.data();
  double *ring = sm + CC::OFF_RING;
  struct Th { double accA[27], accB[27], rA, rB; };
  std::vector<Th> th(CC::NT);
  for (int cta = 0; cta < ntile * zchunks; ++cta) {
    for (int i = 0; i < CC::SMEM_DOUBLES; ++i) sm[i] = NAN; // uninitialised shared memory must never be consumed
    const int tile = cta % ntile, zc = cta / ntile;
    const int x0 = (tile % P.tiles_x) * CC::TBX, y0 = (tile / P.tiles_x) * CC::TBY;
    const int zs = P.z_first + zc * P.z_chunk;
    const int ze = (zs + P.z_chunk < P.z_last) ? zs + P.z_chunk : P.z_last;
    for (auto &t : th) {
      for (int k = 0; k < 27; ++k) t.accA[k] = t.accB[k] = 0.0;
      t.rA = t.rB = 0.0;
    }
    for (int tid = 0; tid < CC::NT; ++tid) {
      cont::stage_plane(P, ring + ((zs + 2) % 3) * CC::PLN, tid, x0, y0, zs - 1);
      cont::stage_plane(P, ring + (zs % 3) * CC::PLN, tid, x0, y0, zs);
    }
    // ---- barrier
    for (int ez = zs - 1; ez < ze; ++ez) {
      for (int tid = 0; tid < CC::NT; ++tid) {
        cont::stage_plane(P, ring + ((ez + 2 + 3) % 3) * CC::PLN, tid, x0, y0, ez + 2);
        cont::p1_element_c(P, sm, tid, x0, y0, ez, ring + ((ez + 3) % 3) * CC::PLN, ring + ((ez + 1 + 3) % 3) * CC::PLN);
      }
      // ---- barrier
      for (int tid = 0; tid < CC::NT; ++tid) {
        const int lx = tid % CC::ELX, ly = tid / CC::ELX;
        const bool in_tile = lx < CC::TBX && ly < CC::TBY;
        const bool active = in_tile && x0 + lx < P.NX && y0 + ly < P.NY;
        Th &t = th[tid];
        if (in_tile) cont::p2_layer<SH>(sm, lx, ly, t.accA, t.rA, t.accB, t.rB);
        if (active && ez >= zs) cont::write_row(P, x0 + lx, y0 + ly, ez, t.accA, t.rA);
        for (int k = 0; k < 27; ++k) {
          t.accA[k] = t.accB[k];
          t.accB[k] = 0.0;
        }
        t.rA = t.rB;
        t.rB = 0.0;
      }
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


static double run_case(int nx, int ny, int nzg, int R, int z_chunk, int options, int overwrite)
{
  std::vector<RankMesh> ms;
  std::vector<Graph> gs(R);
  for (int r = 0; r < R; ++r) {
    ms.push_back(make_rank(nx, ny, nzg, r, R));
    Graph &g = gs[r];
    g.rank = r; g.nranks = R; g.ndof = 1;
    g.n_nodes = (int)ms[r].gid.size();
    g.n_elems = (int64_t)ms[r].conn.size() / 8;
    g.box[0] = nx; g.box[1] = ny; g.box[2] = ms[r].nz;
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

  const double dt = 0.05, gamma1 = 1.5;
  double worst = 0.0;
  for (int r = 0; r < R; ++r) {
    orc_kernel_desc o;
    memset(&o, 0, sizeof(o));
    o.kind = ORC_CONTINUITY_ADV;
    o.shiftMdot = (options & NALU_OPT_SHIFT_MDOT) != 0;
    o.shiftPoisson = (options & NALU_OPT_SHIFT_POISSON) != 0;
    o.reducedSensitivities = (options & NALU_OPT_REDUCED_SENS) != 0;
    o.om_densFac = 1; o.dt = dt; o.gamma1 = gamma1; o.gamma2 = -2.0; o.gamma3 = 0.5;
    o.coordinates = ms[r].coords.data();
    o.velocity = ms[r].vel.data();
    o.density = ms[r].rho.data();
    o.pressure = ms[r].p.data();
    o.Gpdx = ms[r].gp.data();
    orc_assemble_elem(ls, r, &o, 1);

    const Graph &g = gs[r];
    std::vector<long long> bp(g.browptr.begin(), g.browptr.end());
    ContParams P;
    memset(&P, 0, sizeof(P));
    P.nx = nx; P.ny = ny; P.nz = ms[r].nz;
    P.NX = nx + 1; P.NY = ny + 1; P.NZ = ms[r].nz + 1;
    P.coords = ms[r].coords.data(); P.vel = ms[r].vel.data(); P.rho = ms[r].rho.data(); P.p = ms[r].p.data(); P.gp = ms[r].gp.data();
    P.node_row = g.node_row.data(); P.browptr = bp.data(); P.slot27 = slot27[r].data();
    int errflag = 0;
    P.err = &errflag;
    P.tau = dt / gamma1; P.options = options; P.overwrite = overwrite;
    P.z_first = 0; P.z_last = P.NZ;
    const size_t nv = (size_t)(g.nnz_owned() + g.nnz_shared()), nr = (size_t)g.n_row_nodes();
    // overwrite mode must not depend on what the arrays held; accumulate mode starts from zero
    std::vector<double> values(nv, overwrite ? 777.0 : 0.0), rhs(nr, overwrite ? -55.0 : 0.0);
    P.values = values.data(); P.rhs = rhs.data();
    if (options & (NALU_OPT_SHIFT_POISSON | NALU_OPT_REDUCED_SENS)) emulate<true>(P, z_chunk);
    else emulate<false>(P, z_chunk);
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
  struct Case { int nx, ny, nz, R, zc, opt, ow; };
  const Case cases[] = {
      {1, 1, 1, 1, 16, 0, 1},  {2, 3, 2, 1, 16, 0, 1},  {7, 5, 4, 1, 2, NALU_OPT_SHIFT_MDOT, 1},
      {17, 9, 3, 1, 3, NALU_OPT_SHIFT_POISSON, 1},  {16, 8, 3, 1, 16, NALU_OPT_REDUCED_SENS, 1},
      {6, 4, 6, 2, 16, NALU_OPT_REDUCED_SENS | NALU_OPT_SHIFT_MDOT, 1},  {9, 7, 5, 2, 2, 0, 0},
      {5, 5, 7, 3, 1, NALU_OPT_SHIFT_POISSON | NALU_OPT_REDUCED_SENS, 1},  {31, 15, 2, 1, 16, 0, 0},
  };
  for (const Case &c : cases) {
    const double e = run_case(c.nx, c.ny, c.nz, c.R, c.zc, c.opt, c.ow);
    printf("cont %dx%dx%d ranks %d zchunk %d opt %d overwrite %d: relerr %.3e\n", c.nx, c.ny, c.nz, c.R, c.zc, c.opt, c.ow, e);
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
