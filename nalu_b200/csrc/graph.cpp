// graph.cpp -- host half of the CSR graph builder (see graph.h).
//
// Result contract (checked bit-for-bit against the oracle's literal restatement of
// src/TpetraLinearSystem.C in tests/test_graph.py):
//   * row order: locally-owned node ids ascending, then shared-not-owned ids ascending,
//     numDof consecutive rows per node                          (:190-327)
//   * column map: owned, shared-not-owned, then remote ids ordered by (owner, id)
//                                                               (:1051-1094)
//   * each row: sorted unique column lids; every dof row of a node has the same columns
//     (LocalGraphArrays::insertIndices, fill_in_extra_dof_rows_per_node :1227-1246)
//   * owned rows of nodes that other ranks share also hold the columns those ranks
//     connect to them (communicate_remote_columns :700-738)
#include "graph.h"

#include <algorithm>
#include <cstring>
#include <numeric>
#include <omp.h>

namespace nalu_b200 {

namespace {

struct GidNode {
  int64_t gid;
  int32_t node;
};

inline bool gid_less(const GidNode &a, const GidNode &b)
{
  return a.gid < b.gid || (a.gid == b.gid && a.node < b.node);
}

// index of gid in a sorted unique list, or -1
inline int32_t find_gid(const int64_t *v, int32_t n, int64_t gid)
{
  const int64_t *p = std::lower_bound(v, v + n, gid);
  return (p != v + n && *p == gid) ? int32_t(p - v) : -1;
}

} // namespace

int graph_number_and_adjacency(Graph &g, std::string &err)
{
  const int32_t nn = g.n_nodes;
  // ---- numbering -----------------------------------------------------------------
  std::vector<GidNode> owned, shared;
  owned.reserve(nn);
  for (int32_t n = 0; n < nn; ++n) {
    if (g.owner[n] < 0 || g.owner[n] >= g.nranks) {
      err = "node_owner out of range";
      return 1;
    }
    if (g.gid[n] < 1) {
      err = "node ids are 1-based";
      return 1;
    }
    (g.owner[n] == g.rank ? owned : shared).push_back({g.gid[n], n});
  }
  if (!std::is_sorted(owned.begin(), owned.end(), gid_less)) std::sort(owned.begin(), owned.end(), gid_less);
  if (!std::is_sorted(shared.begin(), shared.end(), gid_less)) std::sort(shared.begin(), shared.end(), gid_less);
  auto same = [](const GidNode &a, const GidNode &b) { return a.gid == b.gid; };
  owned.erase(std::unique(owned.begin(), owned.end(), same), owned.end());
  shared.erase(std::unique(shared.begin(), shared.end(), same), shared.end());
  g.n_owned_nodes = int32_t(owned.size());
  g.n_shared_nodes = int32_t(shared.size());
  const int32_t nrow = g.n_row_nodes();
  if ((int64_t)nrow * g.ndof >= 2000000000LL) {
    err = "more than 2e9 local dof rows: exceeds 32-bit local ordinals";
    return 1;
  }
  g.row_node_gid.resize(nrow);
  g.shared_owner.resize(g.n_shared_nodes);
  for (int32_t i = 0; i < g.n_owned_nodes; ++i) g.row_node_gid[i] = owned[i].gid;
  for (int32_t i = 0; i < g.n_shared_nodes; ++i) {
    g.row_node_gid[g.n_owned_nodes + i] = shared[i].gid;
    g.shared_owner[i] = g.owner[shared[i].node];
  }
  g.node_row.resize(nn);
  const int64_t *og = g.row_node_gid.data();
  const int64_t *sg = og + g.n_owned_nodes;
  int bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
  for (int32_t n = 0; n < nn; ++n) {
    int32_t r;
    if (g.owner[n] == g.rank) {
      r = find_gid(og, g.n_owned_nodes, g.gid[n]);
    } else {
      r = find_gid(sg, g.n_shared_nodes, g.gid[n]);
      if (r >= 0) r += g.n_owned_nodes;
    }
    if (r < 0) ++bad;
    g.node_row[n] = r;
  }
  if (bad) {
    err = "a node id is owned by this rank on one alias and by another rank on another";
    return 1;
  }
  // the node the reference's node algorithms visit for a row: periodic subjects are excluded from their selectors
  // (src/AssembleNodeSolverAlgorithm.C:57-111), which leaves the monarch = the lowest local node of the id
  g.row_rep.assign(nrow, -1);
  for (int32_t n = nn - 1; n >= 0; --n) g.row_rep[g.node_row[n]] = n;

  // ---- element incidence per row node ------------------------------------------------
  const int64_t ne = g.n_elems;
  const int npe = g.npe; // 8 (Hex8) or 27 (Hex27)
  std::vector<int32_t> erow(size_t(ne) * npe);
#pragma omp parallel for schedule(static)
  for (int64_t q = 0; q < ne * npe; ++q) {
    const int32_t n = g.elem_nodes[q];
    erow[q] = (n >= 0 && n < nn) ? g.node_row[n] : -1;
  }
  for (int64_t q = 0; q < ne * npe; ++q)
    if (erow[q] < 0) {
      err = "elem_nodes entry out of range";
      return 1;
    }
  std::vector<int64_t> inc_ptr(size_t(nrow) + 1, 0);
  for (int64_t q = 0; q < ne * npe; ++q) ++inc_ptr[erow[q] + 1];
  for (int32_t r = 0; r < nrow; ++r) inc_ptr[r + 1] += inc_ptr[r];
  std::vector<int64_t> inc(size_t(ne) * npe);
  {
    std::vector<int64_t> pos(inc_ptr.begin(), inc_ptr.end() - 1);
    for (int64_t e = 0; e < ne; ++e)
      for (int k = 0; k < npe; ++k) inc[pos[erow[e * npe + k]]++] = e;
  }

  // ---- element colouring (greedy over the row-node incidence; parity of (i,j,k) on a generated box) ----------
  g.color_elems.clear();
  g.color_ptr.clear();
  if (ne > 0 && ne < (int64_t(1) << 31)) {
    std::vector<int8_t> color(size_t(ne), -1);
    bool ok = true;
    int ncolor = 0;
    bool aliased = false; // an element that touches one row through two of its nodes races with itself
#pragma omp parallel for schedule(static) reduction(|| : aliased)
    for (int64_t e = 0; e < ne; ++e)
      for (int a = 0; a < npe; ++a)
        for (int b = a + 1; b < npe; ++b) aliased = aliased || erow[e * npe + a] == erow[e * npe + b];
    if (aliased) ok = false;
    if (ok && npe == 8 && g.box[0] > 0 && nrow == nn) {
      const int64_t nx = g.box[0], ny = g.box[1];
#pragma omp parallel for schedule(static)
      for (int64_t e = 0; e < ne; ++e) color[e] = int8_t(((e % nx) & 1) | ((((e / nx) % ny) & 1) << 1) | (((e / (nx * ny)) & 1) << 2));
      ncolor = 8;
    } else if (ok) {
      for (int64_t e = 0; e < ne && ok; ++e) {
        uint64_t used = 0;
        for (int k = 0; k < npe; ++k) {
          const int32_t r = erow[e * npe + k];
          for (int64_t p = inc_ptr[r]; p < inc_ptr[r + 1]; ++p) {
            const int8_t c = color[inc[p]];
            if (c >= 0) used |= uint64_t(1) << c;
          }
        }
        if (~used == 0) { ok = false; break; }
        const int c = __builtin_ctzll(~used);
        color[e] = int8_t(c);
        ncolor = std::max(ncolor, c + 1);
      }
    }
    if (ok) {
      // sort key = (chunk of consecutive element ids, colour): the general kernel runs one launch per (chunk, colour) in
      // this order (graph.h, kColorChunk: one chunk today)
      const int64_t kChunk = g.color_chunk > 0 ? g.color_chunk : kColorChunk;
      const int64_t nchunk = (ne + kChunk - 1) / kChunk;
      g.n_colors = ncolor;
      g.color_ptr.assign(size_t(nchunk * ncolor) + 1, 0);
      for (int64_t e = 0; e < ne; ++e) ++g.color_ptr[(e / kChunk) * ncolor + color[e] + 1];
      for (int64_t c = 0; c < nchunk * ncolor; ++c) g.color_ptr[c + 1] += g.color_ptr[c];
      g.color_elems.resize(size_t(ne));
      std::vector<int64_t> pos(g.color_ptr.begin(), g.color_ptr.end() - 1);
      for (int64_t e = 0; e < ne; ++e) g.color_elems[pos[(e / kChunk) * ncolor + color[e]]++] = int32_t(e);
    }
  }

  // ---- adjacency: nodes of incident elements, sorted unique --------------------------
  const int nchunk = std::max(1, omp_get_max_threads() * 8);
  std::vector<std::vector<int32_t>> chunk_adj(nchunk);
  std::vector<int32_t> row_len(size_t(nrow), 0);
#pragma omp parallel for schedule(dynamic, 1)
  for (int c = 0; c < nchunk; ++c) {
    const int32_t r0 = int32_t(int64_t(nrow) * c / nchunk), r1 = int32_t(int64_t(nrow) * (c + 1) / nchunk);
    std::vector<int32_t> &out = chunk_adj[c];
    out.reserve(size_t(r1 - r0) * 27);
    std::vector<int32_t> cand;
    for (int32_t r = r0; r < r1; ++r) {
      cand.clear();
      for (int64_t p = inc_ptr[r]; p < inc_ptr[r + 1]; ++p) {
        const int32_t *en = &erow[inc[p] * npe];
        cand.insert(cand.end(), en, en + npe);
      }
      std::sort(cand.begin(), cand.end());
      cand.erase(std::unique(cand.begin(), cand.end()), cand.end());
      row_len[r] = int32_t(cand.size());
      out.insert(out.end(), cand.begin(), cand.end());
    }
  }
  g.adj_ptr.assign(size_t(nrow) + 1, 0);
  for (int32_t r = 0; r < nrow; ++r) g.adj_ptr[r + 1] = g.adj_ptr[r] + row_len[r];
  g.adj.resize(size_t(g.adj_ptr[nrow]));
#pragma omp parallel for schedule(dynamic, 1)
  for (int c = 0; c < nchunk; ++c) {
    const int32_t r0 = int32_t(int64_t(nrow) * c / nchunk);
    if (!chunk_adj[c].empty())
      std::memcpy(&g.adj[size_t(g.adj_ptr[r0])], chunk_adj[c].data(), chunk_adj[c].size() * sizeof(int32_t));
    std::vector<int32_t>().swap(chunk_adj[c]);
  }
  return 0;
}

// Handshake message to the owner `p` of some of my shared-not-owned row nodes (int64 stream):
//   [nNodes] then per node (ascending id):
//   [nodeId, sharedNodeIndex, sharedBlockStart, nCols, (colNodeId, colOwner) x nCols]
// Columns are listed in my CSR order, so the owner knows my row layout.
int graph_build_send_messages(Graph &g, std::string &err)
{
  (void)err;
  g.send_msg.assign(g.nranks, {});
  g.recv_msg.assign(g.nranks, {});
  g.recv_have.assign(g.nranks, 0);
  std::vector<int64_t> count(g.nranks, 0);
  for (int32_t s = 0; s < g.n_shared_nodes; ++s) ++count[g.shared_owner[s]];
  for (int p = 0; p < g.nranks; ++p)
    if (count[p]) g.send_msg[p].push_back(count[p]);
  const int64_t shared0 = g.adj_ptr[g.n_owned_nodes];
  for (int32_t s = 0; s < g.n_shared_nodes; ++s) {
    const int32_t r = g.n_owned_nodes + s;
    std::vector<int64_t> &m = g.send_msg[g.shared_owner[s]];
    m.push_back(g.row_node_gid[r]);
    m.push_back(s);
    m.push_back(g.adj_ptr[r] - shared0);
    m.push_back(g.adj_ptr[r + 1] - g.adj_ptr[r]);
    for (int64_t p = g.adj_ptr[r]; p < g.adj_ptr[r + 1]; ++p) {
      const int32_t c = g.adj[p];
      m.push_back(g.row_node_gid[c]);
      m.push_back(c < g.n_owned_nodes ? g.rank : g.shared_owner[c - g.n_owned_nodes]);
    }
  }
  return 0;
}

namespace {
struct OwnerGid {
  int32_t owner;
  int64_t gid;
  bool operator<(const OwnerGid &o) const { return owner < o.owner || (owner == o.owner && gid < o.gid); }
  bool operator==(const OwnerGid &o) const { return owner == o.owner && gid == o.gid; }
};
struct RowCol {
  int32_t row, col;
  bool operator<(const RowCol &o) const { return row < o.row || (row == o.row && col < o.col); }
  bool operator==(const RowCol &o) const { return row == o.row && col == o.col; }
};
} // namespace

int graph_finalize_host(Graph &g, std::vector<uint8_t> &elem_slot, std::vector<uint8_t> &node_slot27,
                        std::string &err)
{
  const int32_t nrow = g.n_row_nodes();
  const int32_t nown = g.n_owned_nodes;
  const int ndof = g.ndof;
  const int64_t *og = g.row_node_gid.data();
  const int64_t *sg = og + nown;

  // ---- remote columns announced by the ranks that share my nodes ---------------------
  std::vector<OwnerGid> remote;
  for (int q = 0; q < g.nranks; ++q) {
    const std::vector<int64_t> &m = g.recv_msg[q];
    if (m.empty()) continue;
    size_t pos = 0;
    const int64_t nnodes = m[pos++];
    for (int64_t i = 0; i < nnodes; ++i) {
      if (pos + 4 > m.size()) { err = "truncated handshake message"; return 1; }
      const int64_t rowId = m[pos];
      const int64_t ncols = m[pos + 3];
      pos += 4;
      if (find_gid(og, nown, rowId) < 0) { err = "handshake row not owned by this rank"; return 1; }
      for (int64_t c = 0; c < ncols; ++c) {
        const int64_t cg = m[pos++];
        const int32_t co = int32_t(m[pos++]);
        if (find_gid(og, nown, cg) < 0 && find_gid(sg, g.n_shared_nodes, cg) < 0) remote.push_back({co, cg});
      }
    }
  }
  std::sort(remote.begin(), remote.end());
  remote.erase(std::unique(remote.begin(), remote.end()), remote.end());
  g.remote_gid.resize(remote.size());
  g.remote_owner.resize(remote.size());
  for (size_t i = 0; i < remote.size(); ++i) {
    g.remote_gid[i] = remote[i].gid;
    g.remote_owner[i] = remote[i].owner;
    if (remote[i].owner == g.rank) { err = "remote column claims to be owned by this rank"; return 1; }
  }
  auto col_of = [&](int32_t owner, int64_t gid) -> int32_t {
    int32_t c = find_gid(og, nown, gid);
    if (c >= 0) return c;
    c = find_gid(sg, g.n_shared_nodes, gid);
    if (c >= 0) return nown + c;
    auto it = std::lower_bound(remote.begin(), remote.end(), OwnerGid{owner, gid});
    return nrow + int32_t(it - remote.begin());
  };

  // ---- extra (row, col) pairs for owned rows -------------------------------------------
  std::vector<RowCol> extra;
  for (int q = 0; q < g.nranks; ++q) {
    const std::vector<int64_t> &m = g.recv_msg[q];
    if (m.empty()) continue;
    size_t pos = 0;
    const int64_t nnodes = m[pos++];
    for (int64_t i = 0; i < nnodes; ++i) {
      const int32_t r = find_gid(og, nown, m[pos]);
      const int64_t ncols = m[pos + 3];
      pos += 4;
      for (int64_t c = 0; c < ncols; ++c) {
        extra.push_back({r, col_of(int32_t(m[pos + 1]), m[pos])});
        pos += 2;
      }
    }
  }
  std::sort(extra.begin(), extra.end());
  extra.erase(std::unique(extra.begin(), extra.end()), extra.end());

  // ---- final block graph -----------------------------------------------------------------
  g.browptr.assign(size_t(nrow) + 1, 0);
  if (extra.empty()) {
    g.browptr = g.adj_ptr;
    g.bcol.swap(g.adj);
  } else {
    std::vector<int32_t> merged;
    g.bcol.clear();
    g.bcol.reserve(g.adj.size() + extra.size());
    size_t xp = 0;
    for (int32_t r = 0; r < nrow; ++r) {
      const int32_t *a0 = &g.adj[size_t(g.adj_ptr[r])], *a1 = &g.adj[size_t(g.adj_ptr[r + 1])];
      size_t x0 = xp;
      while (xp < extra.size() && extra[xp].row == r) ++xp;
      if (x0 == xp) {
        g.bcol.insert(g.bcol.end(), a0, a1);
      } else {
        merged.clear();
        merged.insert(merged.end(), a0, a1);
        for (size_t x = x0; x < xp; ++x) merged.push_back(extra[x].col);
        std::sort(merged.begin(), merged.end());
        merged.erase(std::unique(merged.begin(), merged.end()), merged.end());
        g.bcol.insert(g.bcol.end(), merged.begin(), merged.end());
      }
      g.browptr[r + 1] = int64_t(g.bcol.size());
    }
    std::vector<int32_t>().swap(g.adj);
  }
  g.nblk_owned = g.browptr[nown];
  g.nblk_shared = g.browptr[nrow] - g.browptr[nown];
  if (g.nnz_owned() + g.nnz_shared() < 0) { err = "nnz overflow"; return 1; }

  auto find_slot = [&](int32_t r, int32_t c) -> int {
    const int32_t *b0 = &g.bcol[size_t(g.browptr[r])], *b1 = &g.bcol[size_t(g.browptr[r + 1])];
    const int32_t *p = std::lower_bound(b0, b1, c);
    return (p != b1 && *p == c) ? int(p - b0) : -1;
  };

  // ---- element scatter map: slot of column node ic inside the row of node k -----------
  const int64_t ne = g.n_elems;
  const int npe = g.npe;
  elem_slot.resize(size_t(ne) * npe * npe);
  int overflow = 0;
#pragma omp parallel for schedule(static) reduction(+ : overflow)
  for (int64_t e = 0; e < ne; ++e) {
    int32_t rn[27];
    for (int k = 0; k < npe; ++k) rn[k] = g.node_row[g.elem_nodes[e * npe + k]];
    for (int k = 0; k < npe; ++k)
      for (int ic = 0; ic < npe; ++ic) {
        const int s = find_slot(rn[k], rn[ic]);
        if (s < 0 || s >= 255) ++overflow;
        elem_slot[(size_t(e) * npe + k) * npe + ic] = uint8_t(s);
      }
  }
  if (overflow) { err = "row with more than 254 node blocks (or missing connection)"; return 1; }

  const int64_t d2 = int64_t(ndof) * ndof;
  // ---- box meshes: slot of each lattice neighbour ----------------------------------------
  node_slot27.clear();
  if (g.box[0] > 0) {
    const int64_t sx = g.box[0] + 1, sy = g.box[1] + 1, sz = g.box[2] + 1;
    if (sx * sy * sz != g.n_nodes || int64_t(g.box[0]) * g.box[1] * g.box[2] != ne) {
      err = "box dimensions do not match node / element counts";
      return 1;
    }
    node_slot27.assign(size_t(g.n_nodes) * 27, 0xFF);
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < g.n_nodes; ++n) {
      const int64_t x = n % sx, y = (n / sx) % sy, z = n / (sx * sy);
      const int32_t r = g.node_row[n];
      for (int dz = -1; dz <= 1; ++dz)
        for (int dy = -1; dy <= 1; ++dy)
          for (int dx = -1; dx <= 1; ++dx) {
            const int64_t xx = x + dx, yy = y + dy, zz = z + dz;
            if (xx < 0 || yy < 0 || zz < 0 || xx >= sx || yy >= sy || zz >= sz) continue;
            const int s = find_slot(r, g.node_row[xx + sx * (yy + sy * zz)]);
            if (s >= 0) node_slot27[size_t(n) * 27 + (dz + 1) * 9 + (dy + 1) * 3 + (dx + 1)] = uint8_t(s);
          }
    }
    // row descriptors of the tiled kernels.  They rely on one property of the column ordering: the columns of a row
    // that share dz are contiguous in the row (true for every z-slab ownership: owned, shared and remote columns keep
    // lattice order inside a plane).  A box declared with another ownership pattern (e.g. an x / y decomposition, where
    // a node plane mixes owned and shared nodes) does not have it: the block then loses its box attribute and is
    // assembled by the general kernel.
    g.node_desc.assign(size_t(g.n_nodes) * 3, 0);
    int scattered = 0;
#pragma omp parallel for schedule(static) reduction(+ : scattered)
    for (int64_t n = 0; n < g.n_nodes; ++n) {
      const int32_t r = g.node_row[n];
      const int64_t nb = g.browptr[r + 1] - g.browptr[r];
      bool natural = nb == 27;
      uint64_t w2 = 0;
      for (int gq = 0; gq < 3; ++gq) {
        unsigned lo = 255, hi = 0, c = 0;
        for (int k = 0; k < 9; ++k) {
          const unsigned v = node_slot27[size_t(n) * 27 + gq * 9 + k];
          if (v == 0xFF) continue;
          lo = v < lo ? v : lo;
          hi = v > hi ? v : hi;
          ++c;
        }
        if (c && hi - lo + 1 != c) ++scattered;
        for (int k = 0; k < 9; ++k) natural = natural && node_slot27[size_t(n) * 27 + gq * 9 + k] == lo + k;
        w2 |= uint64_t(c ? lo : 0) << (8 * gq) | uint64_t(c) << (24 + 8 * gq);
      }
      g.node_desc[size_t(n) * 3 + 0] = uint64_t(d2 * g.browptr[r]);
      g.node_desc[size_t(n) * 3 + 1] = uint64_t(uint32_t(r)) | uint64_t(nb & 0xFF) << 32 | uint64_t(natural) << 40 | uint64_t(1) << 41;
      g.node_desc[size_t(n) * 3 + 2] = w2;
    }
    // the tiled kernel assumes Ioss `generated:` connectivity
    const int64_t nx = g.box[0], ny = g.box[1];
    int badconn = 0;
#pragma omp parallel for schedule(static) reduction(+ : badconn)
    for (int64_t e = 0; e < ne; ++e) {
      const int64_t i = e % nx, j = (e / nx) % ny, k = e / (nx * ny);
      const int64_t n0 = i + sx * (j + sy * k);
      const int64_t ref[8] = {n0, n0 + 1, n0 + 1 + sx, n0 + sx, n0 + sx * sy, n0 + 1 + sx * sy,
                              n0 + 1 + sx + sx * sy, n0 + sx + sx * sy};
      for (int q = 0; q < 8; ++q)
        if (g.elem_nodes[e * 8 + q] != ref[q]) ++badconn;
    }
    if (badconn) { err = "box declared but connectivity is not the generated-box pattern"; return 1; }
    if (scattered) {
      g.box[0] = g.box[1] = g.box[2] = 0;
      node_slot27.clear();
      g.node_desc.clear();
    }
  }

  // ---- halo maps: send side -------------------------------------------------------------
  g.send_peers.clear();
  g.recv_peers.clear();
  for (int p = 0; p < g.nranks; ++p) {
    if (g.send_msg[p].empty()) continue;
    HaloPeer hp;
    hp.rank = p;
    for (int32_t s = 0; s < g.n_shared_nodes; ++s)
      if (g.shared_owner[s] == p) hp.send_nodes.push_back(s);
    for (int32_t s : hp.send_nodes) {
      const int32_t r = nown + s;
      hp.n_value_doubles += (g.browptr[r + 1] - g.browptr[r]) * d2;
    }
    hp.n_rhs_doubles = int64_t(hp.send_nodes.size()) * ndof;
    hp.h_send_node_packed_off.push_back(0);
    for (int32_t s : hp.send_nodes) {
      const int32_t r = nown + s;
      hp.h_send_node_packed_off.push_back(hp.h_send_node_packed_off.back() +
                                          (g.browptr[r + 1] - g.browptr[r]) * d2);
    }
    g.send_peers.push_back(std::move(hp));
  }

  // ---- halo maps: receive (owner) side ----------------------------------------------------
  // one entry per (peer row node, block): where the ndof x ndof block sits in the sender's
  // layout (packed buffer / in place) and in my owned row
  for (int q = 0; q < g.nranks; ++q) {
    const std::vector<int64_t> &m = g.recv_msg[q];
    if (m.empty()) continue;
    HaloPeer hp;
    hp.rank = q;
    size_t pos = 0;
    const int64_t nnodes = m[pos++];
    int64_t packed = 0;
    bool first = true;
    for (int64_t i = 0; i < nnodes; ++i) {
      const int32_t r = find_gid(og, nown, m[pos]);
      const int64_t sidx = m[pos + 1];
      const int64_t sblk = m[pos + 2];
      const int64_t ncols = m[pos + 3];
      pos += 4;
      if (first) {
        hp.peer_values_off = sblk * d2;
        hp.peer_rhs_off = sidx * ndof;
        first = false;
      }
      const int64_t nb = g.browptr[r + 1] - g.browptr[r];
      for (int64_t c = 0; c < ncols; ++c) {
        const int32_t col = col_of(int32_t(m[pos + 1]), m[pos]);
        pos += 2;
        const int slot = find_slot(r, col);
        if (slot < 0) { err = "halo: column announced by a sharer is missing from the owned row"; return 1; }
        hp.h_src_packed.push_back(packed + c * ndof);
        hp.h_src_inplace.push_back(sblk * d2 + c * ndof);
        hp.h_src_stride.push_back(int32_t(ncols * ndof));
        hp.h_dst.push_back(g.browptr[r] * d2 + int64_t(slot) * ndof);
        hp.h_dst_stride.push_back(int32_t(nb * ndof));
      }
      packed += ncols * d2;
      hp.h_rhs_src_inplace.push_back(sidx * ndof);
      hp.h_rhs_dst.push_back(int64_t(r) * ndof);
    }
    hp.n_value_doubles = packed;
    hp.n_rhs_doubles = nnodes * ndof;
    hp.n_segments = int64_t(hp.h_dst.size());
    hp.n_rhs_nodes = nnodes;
    g.recv_peers.push_back(std::move(hp));
  }
  return 0;
}

} // namespace nalu_b200
