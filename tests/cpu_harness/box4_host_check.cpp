// CPU emulation of the row-thread tiled momentum kernel (nalu_b200/csrc/box4_tile.cuh compiled as plain
// C++): the phases are executed thread by thread with the kernel's own control flow
// (assemble_box4.cu), the CTA barriers being the boundaries between the loops, and the assembled
// CSR values / rhs are compared with the oracle's literal restatement of the reference on the
// same mesh and state.  Covers 1 rank and both ranks of a 2-slab decomposition (shared-not-owned
// rows, remote columns), several tile shapes and z chunkings.  Test scaffolding.
#define __host__
#define __device__
#define __forceinline__ inline
#include "../../nalu_b200/csrc/box4_tile.cuh"
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


using C2 = box2::Cfg;

static int g_align_fail = 0;
// copy_staged_warp of assemble_box4.cu
static void copy_staged_host(const BoxParams &P, double *sm, const box4::Row &N, const box4::RunW (&runs)[32])
{
  const double *S = box4::warp_slice(sm, N);
  for (int r = 0; r < 32; ++r)
    for (int lane = 0; lane < 32; ++lane)
      if (lane < runs[r].L) {
        double *d = P.values + runs[r].G + lane;
        *d = P.overwrite ? S[r * 27 + lane] : *d + S[r * 27 + lane];
      }
}

static void emulate2(const BoxParams &P0, int z_chunk, double *values, double *rhs, bool bulk)
{
  using C = C2;
  BoxParams P = P0;
  P.values = values;
  P.rhs = rhs;
  P.tiles_x = (P.NX + C::TBX - 1) / C::TBX;
  P.tiles_y = (P.NY + C::TBY - 1) / C::TBY;
  P.z_chunk = z_chunk;
  P.z_first = 0;
  P.z_last = P.NZ;
  const int zchunks = (P.NZ + z_chunk - 1) / z_chunk;
  constexpr int NT = C::NT;
  double *sm = nullptr;
  if (posix_memalign((void **)&sm, 16, sizeof(double) * C::SMEM_DOUBLES)) abort();
  struct Th { box4::Row N; double acc[18][3], dm[18], rr; };
  std::vector<Th> T(NT);
  box2::P2K K;
  box2::p2k_init(K, P.options, P.include_divu);
  for (int cta = 0; cta < P.tiles_x * P.tiles_y * zchunks; ++cta) {
    for (int i = 0; i < C::SMEM_DOUBLES; ++i) sm[i] = NAN; // uninitialised shared memory must never be consumed
    for (int i = 0; i < 2 * C::NEL * C::ES; ++i) sm[C::OFF_F + i] = 0.0;
    double *F = sm + C::OFF_F, *ZE = sm + C::OFF_ZE;
    double *ring = sm + C::OFF_RING, *mdb = sm + C::OFF_MD;
    const int ntile = P.tiles_x * P.tiles_y;
    const int tile = cta % ntile, zc = cta / ntile;
    const int x0 = (tile % P.tiles_x) * C::TBX, y0 = (tile / P.tiles_x) * C::TBY;
    const int zs = zc * P.z_chunk;
    const int ze = (zs + P.z_chunk < P.NZ) ? zs + P.z_chunk : P.NZ;
    for (int tid = 0; tid < NT; ++tid) {
      box4::Row &N = T[tid].N;
      N.node = tid & 31; N.lx = N.node % C::TBX; N.ly = N.node / C::TBX; N.i = tid >> 5;
    }
    for (int tid = 0; tid < NT; ++tid) {
      box2::stage_plane(P, ring + ((zs + 2) % 3) * C::PLN, tid, x0, y0, zs - 1);
      box2::stage_plane(P, ring + (zs % 3) * C::PLN, tid, x0, y0, zs);
      box2::stage_mdot(P, mdb, tid, x0, y0, zs - 1);
      if ((tid >> 5) == 3) box2::meta_load(P, box2::meta_at(sm, zs & 1), tid & 31, x0, y0, zs);
    }
    // ---- barrier
    if (zs >= 1)
      for (int it = 0; it < 4 * C::NEL; ++it) {
        const int k = it / C::NEL, el = it - k * C::NEL;
        box2::p1_item(P, el, k < 2 ? k : 2, k < 2 ? 1 : k - 2, x0, y0, ring + ((zs - 1) % 3) * C::PLN,
                      ring + (zs % 3) * C::PLN, mdb, F, ZE);
      }
    for (int tid = 0; tid < NT; ++tid) box2::stage_plane(P, ring + ((zs + 1) % 3) * C::PLN, tid, x0, y0, zs + 1);
    // ---- barrier
    for (int z = zs; z < ze; ++z) {
      const box2::Meta M = box2::meta_at(sm, z & 1);
      const double *pl_0 = ring + (z % 3) * C::PLN, *pl_p = ring + ((z + 1) % 3) * C::PLN;
      const bool lower = z >= 1, upper = z < P.nz;
      // ---- A1
      for (int tid = 0; tid < NT; ++tid) box2::stage_mdot(P, mdb, tid, x0, y0, z);
      for (int tid = 0; tid < NT; ++tid) {
        Th &t = T[tid];
        t.rr = 0.0;
        for (int c = 0; c < 18; ++c) t.acc[c][0] = t.acc[c][1] = t.acc[c][2] = t.dm[c] = 0.0;
        if (!lower) continue;
        if (t.N.i < 3) {
          box4::half_rows<0>(sm, t.N, K, t.acc);
          box4::exchange_put_row<0>(sm, t.N, t.acc, z);
        } else {
          box4::half_adv<0>(sm, t.N, K, t.dm);
          box4::exchange_put_adv(sm, t.N, t.dm);
        }
      }
      // ---- barrier; A2 + B
      for (int w = 0; w < 3; ++w) {
        box4::RunW runs[32];
        for (int tid = 32 * w; tid < 32 * w + 32; ++tid) {
          Th &t = T[tid];
          if (lower) {
            box4::finalize_row<0>(sm, t.N, t.acc, t.rr, K.f23, z);
            if (bulk) runs[tid & 31] = box4::stage_group(sm, M, t.N, 0, 9, t.acc);
            else box4::store_group(P, M, t.N, 0, 9, t.acc);
          } else {
            for (int c = 0; c < 9; ++c) t.acc[c][0] = t.acc[c][1] = t.acc[c][2] = 0.0;
          }
        }
        if (lower && bulk) copy_staged_host(P, sm, T[32 * w].N, runs);
        for (int tid = 32 * w; tid < 32 * w + 32; ++tid) box4::park(sm, T[tid].N, T[tid].acc, T[tid].rr);
      }
      if (upper)
        for (int it = 0; it < C::NP1A; ++it) {
          const int k = it / C::NEL, el = it - k * C::NEL;
          box2::p1_item(P, el, k < 2 ? k : 2, k < 2 ? 0 : k - 2, x0, y0, pl_0, pl_p, mdb, F, ZE);
        }
      if (z + 1 < ze)
        for (int tid = 0; tid < 32; ++tid) box2::meta_load(P, box2::meta_at(sm, (z + 1) & 1), tid, x0, y0, z + 1);
      // the exchange buffer is rewritten by C1: poison it so that stale reads show
      for (int i = 0; i < 4 * 18 * C::NNODE + 9 * C::NNODE; ++i) sm[box4::OFF_EX + i] = NAN;
      // ---- barrier; C1
      for (int tid = 0; tid < NT; ++tid) {
        Th &t = T[tid];
        t.rr = 0.0;
        for (int c = 0; c < 18; ++c) t.acc[c][0] = t.acc[c][1] = t.acc[c][2] = t.dm[c] = 0.0;
        if (!upper) continue;
        if (t.N.i < 3) {
          box4::half_rows<1>(sm, t.N, K, t.acc);
          box4::exchange_put_row<1>(sm, t.N, t.acc, z);
        } else {
          box4::half_adv<1>(sm, t.N, K, t.dm);
          box4::exchange_put_adv(sm, t.N, t.dm);
        }
      }
      // ---- barrier; C2 + D
      for (int tid = 0; tid < NT; ++tid) box2::stage_plane(P, ring + ((z + 2) % 3) * C::PLN, tid, x0, y0, z + 2);
      for (int w = 0; w < 3; ++w) {
        box4::RunW runs[32];
        for (int tid = 32 * w; tid < 32 * w + 32; ++tid) {
          Th &t = T[tid];
          if (upper) box4::finalize_row<1>(sm, t.N, t.acc, t.rr, K.f23, z);
          box4::unpark_add(sm, t.N, t.acc, t.rr);
          if (!bulk) {
            box4::store_group(P, M, t.N, 1, 0, t.acc);
            if (upper) box4::store_group(P, M, t.N, 2, 9, t.acc);
          }
        }
        if (bulk) {
          for (int k = 0; k < box4::PARKN * C::NNODE; ++k) box4::warp_slice(sm, T[32 * w].N)[k] = NAN;
          for (int tid = 32 * w; tid < 32 * w + 32; ++tid) runs[tid & 31] = box4::stage_group(sm, M, T[tid].N, 1, 0, T[tid].acc);
          copy_staged_host(P, sm, T[32 * w].N, runs);
          if (upper) {
            for (int tid = 32 * w; tid < 32 * w + 32; ++tid) runs[tid & 31] = box4::stage_group(sm, M, T[tid].N, 2, 9, T[tid].acc);
            copy_staged_host(P, sm, T[32 * w].N, runs);
          }
        }
      }
      for (int tid = 0; tid < 96; ++tid) {
        Th &t = T[tid];
        const int lane = tid & 31, warp = tid >> 5;
        if (P.overwrite) box2::zero_uncovered(M, lane, warp, P.values);
        if (M.cnt[lane * 4 + 1]) {
          double *r = P.rhs + (long long)M.row[lane] * 3 + warp;
          *r = P.overwrite ? t.rr : *r + t.rr;
        }
      }
      if (upper)
        for (int it = 0; it < C::NP1B; ++it) {
          const int k = it / C::NEL, el = it - k * C::NEL;
          box2::p1_item(P, el, k, 1, x0, y0, pl_0, pl_p, mdb, F, ZE);
        }
      for (int i = 0; i < C::NNODE * 3 * box4::PARKN + 4 * 18 * C::NNODE + 9 * C::NNODE; ++i) sm[box4::OFF_PARK + i] = NAN;
      // ---- barrier
    }
  }
  free(sm);
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

static double run_case(int nx, int ny, int nzg, int R, int z_chunk, int options, double include_divu, int overwrite, bool bulk)
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
    std::vector<double> rhs(nr, overwrite ? -55.0 : 0.0);
    double *vbuf = nullptr;
    if (posix_memalign((void **)&vbuf, 16, sizeof(double) * (nv + 2))) abort();
    struct Free { double *p; ~Free() { free(p); } } freer{vbuf};
    for (size_t q = 0; q < nv; ++q) vbuf[q] = overwrite ? 777.0 : 0.0;
    struct { double *p; double *data() { return p; } } values{vbuf};
    emulate2(P, z_chunk, values.data(), rhs.data(), bulk);
    if (errflag) { printf("error flag %d\n", errflag); return 1.0; }

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
      {5, 5, 7, 3, 1, NALU_OPT_SHIFTED_GRAD_OP, 1.0, 1}, {17, 8, 3, 1, 16, 0, 1.0, 0},
  };
  for (const Case &c : cases) {
    const double e0 = run_case(c.nx, c.ny, c.nz, c.R, c.zc, c.opt, c.divu, c.ow, false);
    const double e1 = run_case(c.nx, c.ny, c.nz, c.R, c.zc, c.opt, c.divu, c.ow, true);
    printf("box4 %dx%dx%d ranks %d zchunk %d opt %d divu %.0f overwrite %d: relerr plain %.3e  bulk %.3e\n", c.nx, c.ny,
           c.nz, c.R, c.zc, c.opt, c.divu, c.ow, e0, e1);
    worst = fmax(worst, fmax(e0, e1));
  }
  printf("worst %.3e  bulk alignment failures %d\n", worst, g_align_fail);
  if (!(worst < 1e-12) || g_align_fail) {
    printf("FAIL\n");
    return 1;
  }
  printf("PASS\n");
  return 0;
}
