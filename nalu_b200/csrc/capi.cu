// capi.cu -- the extern "C" boundary declared in include/nalu_b200.h.
#include "internal.h"
#include <omp.h>
#include <cstdlib>

#include <atomic>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

namespace nalu_b200 {

static std::mutex g_err_mutex;
static std::string g_last_error;
static std::atomic<long long> g_launches{0};
static std::atomic<int> g_last_mode{0};

void set_error(const std::string &msg)
{
  std::lock_guard<std::mutex> lk(g_err_mutex);
  g_last_error = msg;
}
int cuda_fail(cudaError_t e, const char *what)
{
  set_error(std::string(what) + ": " + cudaGetErrorString(e));
  return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? NALU_B200_ERR_NO_DEVICE : NALU_B200_ERR_CUDA;
}
void count_launch(int n) { g_launches += n; }
static Tuning &tuning_rw()
{
  static Tuning t = [] {
    Tuning v;
    if (const char *e = getenv("NALU_BOX_KERNEL")) v.box_kernel = atoi(e) == 2 ? 2 : 5;
    if (const char *e = getenv("NALU_BOX_ZCHUNK")) v.box_zchunk = atoi(e) > 0 ? atoi(e) : 0;
    if (const char *e = getenv("NALU_SPLIT_LISTS")) v.split_lists = atoi(e) != 0;
    if (const char *e = getenv("NALU_CONT_BOX")) v.cont_tiled = atoi(e) != 0;
    if (const char *e = getenv("NALU_GENERAL_MINB")) v.general_minb = atoi(e) < 2 ? 2 : (atoi(e) > 4 ? 4 : atoi(e));
    return v;
  }();
  return t;
}
const Tuning &tuning() { return tuning_rw(); }
void set_last_mode(int m) { g_last_mode = m; }

template <typename T>
static int upload(T **dst, const std::vector<T> &src)
{
  *dst = nullptr;
  if (src.empty()) return NALU_B200_OK;
  NALU_CUDA(cudaMalloc((void **)dst, src.size() * sizeof(T)));
  NALU_CUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
  return NALU_B200_OK;
}

// dof-level CSR expansion of the node-block graph (fill_in_extra_dof_rows_per_node,
// src/TpetraLinearSystem.C:1227-1246, done as one pass on the device)
__global__ void expand_rowptr_kernel(int r0, int r1, int ndof, const long long *browptr, unsigned long long *rowptr)
{
  const long long base = browptr[r0];
  const long long d2 = (long long)ndof * ndof;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < (long long)(r1 - r0) * ndof;
       t += (long long)gridDim.x * blockDim.x) {
    const int r = r0 + (int)(t / ndof), i = (int)(t % ndof);
    const long long nb = browptr[r + 1] - browptr[r];
    rowptr[t] = (unsigned long long)(d2 * (browptr[r] - base) + (long long)i * nb * ndof);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) rowptr[(long long)(r1 - r0) * ndof] = (unsigned long long)(d2 * (browptr[r1] - base));
}

__global__ void expand_colidx_kernel(int r0, int r1, int ndof, const long long *browptr, const int *bcol, int *colidx)
{
  const long long base = browptr[r0];
  const long long d2 = (long long)ndof * ndof;
  for (int r = r0 + blockIdx.x; r < r1; r += gridDim.x) {
    const long long b0 = browptr[r], nb = browptr[r + 1] - b0;
    int *out = colidx + d2 * (b0 - base);
    const long long n = nb * d2;
    for (long long q = threadIdx.x; q < n; q += blockDim.x) {
      // q = i*(nb*ndof) + b*ndof + j
      const long long within = q % (nb * ndof);
      const long long b = within / ndof;
      const int j = (int)(within - b * ndof);
      out[q] = bcol[b0 + b] * ndof + j;
    }
  }
}

static void free_graph_device(Graph &g)
{
  auto F = [](auto *&p) {
    if (p) cudaFree(p);
    p = nullptr;
  };
  F(g.d_elem_nodes); F(g.d_node_row); F(g.d_row_rep); F(g.d_color_elems); F(g.d_browptr); F(g.d_bcol); F(g.d_elem_slot); F(g.d_node_slot27); F(g.d_node_desc);
  F(g.d_rowelem_ptr); F(g.d_rowelem); F(g.d_elem_nodes27); F(g.d_elem_slot27);
  F(g.d_rowptr_owned); F(g.d_rowptr_shared); F(g.d_colidx_owned); F(g.d_colidx_shared); F(g.d_err);
  for (auto *peers : {&g.send_peers, &g.recv_peers})
    for (HaloPeer &p : *peers) {
      F(p.d_send_nodes); F(p.d_send_node_packed_off); F(p.d_src_packed); F(p.d_src_inplace); F(p.d_src_stride);
      F(p.d_dst); F(p.d_dst_stride); F(p.d_rhs_src_inplace); F(p.d_rhs_dst); F(p.d_buf);
    }
}

static int upload_graph(Graph &g, const std::vector<uint8_t> &elem_slot, const std::vector<uint8_t> &node_slot27)
{
  int rc;
  if ((rc = upload(g.npe == 27 ? &g.d_elem_nodes27 : &g.d_elem_nodes, g.elem_nodes))) return rc;
  if ((rc = upload(&g.d_node_row, g.node_row))) return rc;
  if ((rc = upload(&g.d_row_rep, g.row_rep))) return rc;
  if ((rc = upload(&g.d_color_elems, g.color_elems))) return rc;
  if ((rc = upload(&g.d_browptr, g.browptr))) return rc;
  if ((rc = upload(&g.d_bcol, g.bcol))) return rc;
  if ((rc = upload(g.npe == 27 ? &g.d_elem_slot27 : &g.d_elem_slot, elem_slot))) return rc;
  if ((rc = upload(&g.d_node_slot27, node_slot27))) return rc;
  if ((rc = upload(&g.d_node_desc, g.node_desc))) return rc;
  std::vector<uint64_t>().swap(g.node_desc);
  if (g.npe == 8 && g.box[0] == 0 && g.n_elems > 0 && g.n_elems < (int64_t(1) << 28)) {
    // row node -> (element, local node), elements in ascending order: a counting sort of the connectivity by row
    const int nrow = g.n_row_nodes();
    std::vector<int64_t> ptr(size_t(nrow) + 1, 0);
    for (size_t q = 0; q < g.elem_nodes.size(); ++q) ++ptr[size_t(g.node_row[g.elem_nodes[q]]) + 1];
    for (int r = 0; r < nrow; ++r) ptr[r + 1] += ptr[r];
    std::vector<int32_t> re(g.elem_nodes.size());
    std::vector<int64_t> pos(ptr.begin(), ptr.end() - 1);
    for (size_t q = 0; q < g.elem_nodes.size(); ++q) re[size_t(pos[g.node_row[g.elem_nodes[q]]]++)] = int32_t(q); // q = e * 8 + local
    if ((rc = upload(&g.d_rowelem_ptr, ptr))) return rc;
    if ((rc = upload(&g.d_rowelem, re))) return rc;
  }
  NALU_CUDA(cudaMalloc((void **)&g.d_err, sizeof(int)));
  NALU_CUDA(cudaMemset(g.d_err, 0, sizeof(int)));
  const int nown = g.n_owned_nodes, nrow = g.n_row_nodes(), ndof = g.ndof;
  NALU_CUDA(cudaMalloc((void **)&g.d_rowptr_owned, (size_t(nown) * ndof + 1) * sizeof(uint64_t)));
  NALU_CUDA(cudaMalloc((void **)&g.d_rowptr_shared, (size_t(nrow - nown) * ndof + 1) * sizeof(uint64_t)));
  if (g.nnz_owned()) NALU_CUDA(cudaMalloc((void **)&g.d_colidx_owned, size_t(g.nnz_owned()) * sizeof(int32_t)));
  if (g.nnz_shared()) NALU_CUDA(cudaMalloc((void **)&g.d_colidx_shared, size_t(g.nnz_shared()) * sizeof(int32_t)));
  const long long *bp = (const long long *)g.d_browptr;
  expand_rowptr_kernel<<<296, 256>>>(0, nown, ndof, bp, (unsigned long long *)g.d_rowptr_owned);
  expand_rowptr_kernel<<<296, 256>>>(nown, nrow, ndof, bp, (unsigned long long *)g.d_rowptr_shared);
  if (nown > 0) expand_colidx_kernel<<<(nown < 148 * 32 ? nown : 148 * 32), 128>>>(0, nown, ndof, bp, g.d_bcol, g.d_colidx_owned);
  if (nrow > nown)
    expand_colidx_kernel<<<((nrow - nown) < 148 * 32 ? (nrow - nown) : 148 * 32), 128>>>(nown, nrow, ndof, bp, g.d_bcol,
                                                                                      g.d_colidx_shared);
  count_launch(4);
  NALU_CUDA(cudaGetLastError());
  for (HaloPeer &p : g.send_peers) {
    if ((rc = upload(&p.d_send_nodes, p.send_nodes))) return rc;
    if ((rc = upload(&p.d_send_node_packed_off, p.h_send_node_packed_off))) return rc;
  }
  for (HaloPeer &p : g.recv_peers) {
    if ((rc = upload(&p.d_src_packed, p.h_src_packed))) return rc;
    if ((rc = upload(&p.d_src_inplace, p.h_src_inplace))) return rc;
    if ((rc = upload(&p.d_src_stride, p.h_src_stride))) return rc;
    if ((rc = upload(&p.d_dst, p.h_dst))) return rc;
    if ((rc = upload(&p.d_dst_stride, p.h_dst_stride))) return rc;
    if ((rc = upload(&p.d_rhs_src_inplace, p.h_rhs_src_inplace))) return rc;
    if ((rc = upload(&p.d_rhs_dst, p.h_rhs_dst))) return rc;
    for (auto *v : {&p.h_src_packed, &p.h_src_inplace, &p.h_dst, &p.h_rhs_src_inplace, &p.h_rhs_dst}) std::vector<int64_t>().swap(*v);
    for (auto *v : {&p.h_src_stride, &p.h_dst_stride}) std::vector<int32_t>().swap(*v);
  }
  NALU_CUDA(cudaDeviceSynchronize());
  return NALU_B200_OK;
}

} // namespace nalu_b200

using namespace nalu_b200;

extern "C" {

const char *nalu_b200_last_error(void)
{
  static thread_local std::string copy;
  std::lock_guard<std::mutex> lk(g_err_mutex);
  copy = g_last_error;
  return copy.c_str();
}
int nalu_b200_version(void) { return 100; }
int64_t nalu_b200_launch_count(void) { return g_launches.load(); }
int nalu_b200_last_assemble_mode(void) { return g_last_mode.load(); }
int nalu_b200_tune(const char *name, int32_t value)
{
  if (!name) return NALU_B200_ERR_ARG;
  Tuning &t = tuning_rw();
  const std::string n(name);
  if (n == "box_kernel") t.box_kernel = value == 2 ? 2 : 5;
  else if (n == "box_zchunk") t.box_zchunk = value > 0 ? value : 0;
  else if (n == "split_lists") t.split_lists = value != 0;
  else if (n == "cont_tiled") t.cont_tiled = value != 0;
  else if (n == "general_minb") t.general_minb = value < 2 ? 2 : (value > 4 ? 4 : value);
  else if (n == "general_scatter") t.general_scatter = value != 0;
  else if (n == "urow") t.urow = value != 0;
  else if (n == "urow_owner") t.urow_owner = value != 0;
  else if (n == "urow_minb") t.urow_minb = value < 2 ? 2 : (value > 4 ? 4 : value);
  else if (n == "color_chunk") t.color_chunk = value > 0 ? value : 0;
  else if (n == "host_threads") {
    // OpenMP threads of the host graph builder for the calling thread (launchers such as torchrun export OMP_NUM_THREADS=1)
    if (value > 0) omp_set_num_threads(value);
  }
  else {
    set_error("tune: unknown knob " + n);
    return NALU_B200_ERR_ARG;
  }
  return NALU_B200_OK;
}

static int graph_create_impl(nalu_b200_graph **out, int32_t rank, int32_t nranks, int32_t num_dof, int32_t npe, int32_t n_nodes,
                             const int64_t *node_gid, const int32_t *node_owner, int64_t n_elems, const int32_t *elem_nodes,
                             int32_t box_nx, int32_t box_ny, int32_t box_nz);

int nalu_b200_graph_create(nalu_b200_graph **out, int32_t rank, int32_t nranks, int32_t num_dof, int32_t n_nodes,
                           const int64_t *node_gid, const int32_t *node_owner, int64_t n_elems,
                           const int32_t *elem_nodes, int32_t box_nx, int32_t box_ny, int32_t box_nz)
{
  return graph_create_impl(out, rank, nranks, num_dof, 8, n_nodes, node_gid, node_owner, n_elems, elem_nodes, box_nx, box_ny,
                           box_nz);
}

int nalu_b200_graph_create_topo(nalu_b200_graph **out, int32_t rank, int32_t nranks, int32_t num_dof, int32_t nodes_per_elem,
                                int32_t n_nodes, const int64_t *node_gid, const int32_t *node_owner, int64_t n_elems,
                                const int32_t *elem_nodes)
{
  if (nodes_per_elem != 8 && nodes_per_elem != 27) {
    set_error("graph_create_topo: nodes_per_elem must be 8 (Hex8) or 27 (Hex27)");
    return NALU_B200_ERR_ARG;
  }
  return graph_create_impl(out, rank, nranks, num_dof, nodes_per_elem, n_nodes, node_gid, node_owner, n_elems, elem_nodes, 0, 0, 0);
}

static int graph_create_impl(nalu_b200_graph **out, int32_t rank, int32_t nranks, int32_t num_dof, int32_t npe, int32_t n_nodes,
                             const int64_t *node_gid, const int32_t *node_owner, int64_t n_elems, const int32_t *elem_nodes,
                             int32_t box_nx, int32_t box_ny, int32_t box_nz)
{
  if (!out || nranks < 1 || rank < 0 || rank >= nranks || (num_dof != 1 && num_dof != 3) || n_nodes < 0 || n_elems < 0 ||
      (n_nodes && (!node_gid || !node_owner)) || (n_elems && !elem_nodes)) {
    set_error("graph_create: bad argument (numDof must be 1 or 3)");
    return NALU_B200_ERR_ARG;
  }
  nalu_b200_graph *h = nullptr;
  try {
    h = new nalu_b200_graph();
    Graph &g = h->g;
    g.rank = rank; g.nranks = nranks; g.ndof = num_dof; g.n_nodes = n_nodes; g.n_elems = n_elems;
    if (tuning().color_chunk > 0) g.color_chunk = tuning().color_chunk;
    g.box[0] = box_nx; g.box[1] = box_ny; g.box[2] = box_nz;
    if (!(box_nx > 0 && box_ny > 0 && box_nz > 0)) g.box[0] = g.box[1] = g.box[2] = 0;
    g.gid.assign(node_gid, node_gid + n_nodes);
    g.owner.assign(node_owner, node_owner + n_nodes);
    g.npe = npe;
    g.elem_nodes.assign(elem_nodes, elem_nodes + n_elems * npe);
    std::string err;
    if (graph_number_and_adjacency(g, err) || graph_build_send_messages(g, err)) {
      set_error("graph_create: " + err);
      delete h;
      return NALU_B200_ERR_ARG;
    }
    *out = h;
    return NALU_B200_OK;
  } catch (const std::exception &e) {
    set_error(std::string("graph_create: ") + e.what());
    delete h;
    return NALU_B200_ERR_ARG;
  }
}

void nalu_b200_graph_destroy(nalu_b200_graph *h)
{
  if (!h) return;
  free_graph_device(h->g);
  delete h;
}

int nalu_b200_graph_send_size(const nalu_b200_graph *h, int32_t dest, int64_t *nbytes)
{
  if (!h || !nbytes || dest < 0 || dest >= h->g.nranks) return NALU_B200_ERR_ARG;
  *nbytes = (int64_t)h->g.send_msg[dest].size() * 8;
  return NALU_B200_OK;
}
int nalu_b200_graph_send_data(const nalu_b200_graph *h, int32_t dest, void *buf)
{
  if (!h || dest < 0 || dest >= h->g.nranks) return NALU_B200_ERR_ARG;
  const auto &m = h->g.send_msg[dest];
  if (!m.empty()) std::memcpy(buf, m.data(), m.size() * 8);
  return NALU_B200_OK;
}
int nalu_b200_graph_recv_data(nalu_b200_graph *h, int32_t src, const void *buf, int64_t nbytes)
{
  if (!h || src < 0 || src >= h->g.nranks || nbytes < 0 || nbytes % 8) return NALU_B200_ERR_ARG;
  if (h->g.finalized) {
    set_error("graph_recv_data after finalize");
    return NALU_B200_ERR_STATE;
  }
  auto &m = h->g.recv_msg[src];
  m.resize(size_t(nbytes / 8));
  if (nbytes) std::memcpy(m.data(), buf, size_t(nbytes));
  return NALU_B200_OK;
}

int nalu_b200_graph_finalize(nalu_b200_graph *h)
{
  if (!h) return NALU_B200_ERR_ARG;
  Graph &g = h->g;
  if (g.finalized) {
    set_error("graph_finalize called twice");
    return NALU_B200_ERR_STATE;
  }
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0) {
    set_error("graph_finalize: no CUDA device; the assembly path has no CPU fallback");
    return NALU_B200_ERR_NO_DEVICE;
  }
  try {
    std::vector<uint8_t> elem_slot, node_slot27;
    std::string err;
    if (graph_finalize_host(g, elem_slot, node_slot27, err)) {
      set_error("graph_finalize: " + err);
      return NALU_B200_ERR_ARG;
    }
    int rc = upload_graph(g, elem_slot, node_slot27);
    if (rc) return rc;
    g.finalized = true;
    return NALU_B200_OK;
  } catch (const std::exception &e) {
    set_error(std::string("graph_finalize: ") + e.what());
    return NALU_B200_ERR_ARG;
  }
}

/* host-only variant used by the CPU test-suite to check the graph without a device */
int nalu_b200_graph_finalize_host_only(nalu_b200_graph *h)
{
  if (!h) return NALU_B200_ERR_ARG;
  Graph &g = h->g;
  if (g.finalized) return NALU_B200_ERR_STATE;
  std::vector<uint8_t> elem_slot, node_slot27;
  std::string err;
  if (graph_finalize_host(g, elem_slot, node_slot27, err)) {
    set_error("graph_finalize: " + err);
    return NALU_B200_ERR_ARG;
  }
  g.finalized = true;
  return NALU_B200_OK;
}

int nalu_b200_graph_get_info(const nalu_b200_graph *h, nalu_b200_graph_info *info)
{
  if (!h || !info) return NALU_B200_ERR_ARG;
  const Graph &g = h->g;
  info->num_dof = g.ndof;
  info->n_owned_rows = g.n_owned_nodes * g.ndof;
  info->n_shared_rows = g.n_shared_nodes * g.ndof;
  info->n_cols = (g.n_row_nodes() + (int32_t)g.remote_gid.size()) * g.ndof;
  info->nnz_owned = g.nnz_owned();
  info->nnz_shared = g.nnz_shared();
  info->n_elems = g.n_elems;
  info->n_nodes = g.n_nodes;
  info->is_box = g.box[0] > 0;
  return NALU_B200_OK;
}

int nalu_b200_graph_copy_rowptr(const nalu_b200_graph *h, int shared, uint64_t *rowptr)
{
  if (!h || !rowptr || !h->g.finalized) return NALU_B200_ERR_STATE;
  const Graph &g = h->g;
  const int r0 = shared ? g.n_owned_nodes : 0, r1 = shared ? g.n_row_nodes() : g.n_owned_nodes;
  const int64_t d2 = int64_t(g.ndof) * g.ndof, base = g.browptr[r0];
  size_t t = 0;
  for (int r = r0; r < r1; ++r) {
    const int64_t nb = g.browptr[r + 1] - g.browptr[r];
    for (int i = 0; i < g.ndof; ++i) rowptr[t++] = uint64_t(d2 * (g.browptr[r] - base) + int64_t(i) * nb * g.ndof);
  }
  rowptr[t] = uint64_t(d2 * (g.browptr[r1] - base));
  return NALU_B200_OK;
}

int nalu_b200_graph_copy_colidx(const nalu_b200_graph *h, int shared, int32_t *colidx)
{
  if (!h || !h->g.finalized) return NALU_B200_ERR_STATE;
  const Graph &g = h->g;
  const int r0 = shared ? g.n_owned_nodes : 0, r1 = shared ? g.n_row_nodes() : g.n_owned_nodes;
  size_t t = 0;
  for (int r = r0; r < r1; ++r)
    for (int i = 0; i < g.ndof; ++i)
      for (int64_t b = g.browptr[r]; b < g.browptr[r + 1]; ++b)
        for (int j = 0; j < g.ndof; ++j) colidx[t++] = g.bcol[size_t(b)] * g.ndof + j;
  return NALU_B200_OK;
}

int nalu_b200_graph_copy_row_gids(const nalu_b200_graph *h, int shared, int64_t *gids)
{
  if (!h || !gids) return NALU_B200_ERR_ARG;
  const Graph &g = h->g;
  const int r0 = shared ? g.n_owned_nodes : 0, r1 = shared ? g.n_row_nodes() : g.n_owned_nodes;
  size_t t = 0;
  for (int r = r0; r < r1; ++r)
    for (int i = 0; i < g.ndof; ++i) gids[t++] = int64_t(g.ndof) * (g.row_node_gid[r] - 1) + i + 1; /* GID_ */
  return NALU_B200_OK;
}

int nalu_b200_graph_copy_col_gids(const nalu_b200_graph *h, int64_t *gids)
{
  if (!h || !gids || !h->g.finalized) return NALU_B200_ERR_STATE;
  const Graph &g = h->g;
  size_t t = 0;
  for (int r = 0; r < g.n_row_nodes(); ++r)
    for (int i = 0; i < g.ndof; ++i) gids[t++] = int64_t(g.ndof) * (g.row_node_gid[r] - 1) + i + 1;
  for (size_t r = 0; r < g.remote_gid.size(); ++r)
    for (int i = 0; i < g.ndof; ++i) gids[t++] = int64_t(g.ndof) * (g.remote_gid[r] - 1) + i + 1;
  return NALU_B200_OK;
}

int nalu_b200_graph_copy_node_lids(const nalu_b200_graph *h, int32_t *row_lid, int32_t *col_lid)
{
  if (!h) return NALU_B200_ERR_ARG;
  const Graph &g = h->g;
  for (int32_t n = 0; n < g.n_nodes; ++n) {
    if (row_lid) row_lid[n] = g.node_row[n] * g.ndof;
    if (col_lid) col_lid[n] = g.node_row[n] * g.ndof; /* local nodes: column lid == row lid */
  }
  return NALU_B200_OK;
}

int nalu_b200_graph_copy_elem_colors(const nalu_b200_graph *h, int32_t *n_colors, int32_t *color_of_elem)
{
  if (!h || !n_colors) return NALU_B200_ERR_ARG;
  const Graph &g = h->g;
  *n_colors = g.color_ptr.empty() ? 0 : g.n_colors;
  if (color_of_elem && *n_colors)
    for (size_t grp = 0; grp + 1 < g.color_ptr.size(); ++grp)
      for (int64_t q = g.color_ptr[grp]; q < g.color_ptr[grp + 1]; ++q) color_of_elem[g.color_elems[size_t(q)]] = int32_t(grp % g.n_colors);
  return NALU_B200_OK;
}

int nalu_b200_graph_device_ptrs(const nalu_b200_graph *h, int shared, const uint64_t **d_rowptr, const int32_t **d_colidx)
{
  if (!h || !h->g.finalized || !h->g.d_browptr) return NALU_B200_ERR_STATE;
  if (d_rowptr) *d_rowptr = shared ? h->g.d_rowptr_shared : h->g.d_rowptr_owned;
  if (d_colidx) *d_colidx = shared ? h->g.d_colidx_shared : h->g.d_colidx_owned;
  return NALU_B200_OK;
}

int nalu_b200_zero_system(const nalu_b200_graph *h, double *d_values, double *d_rhs, void *stream)
{
  if (!h || !h->g.d_browptr) return NALU_B200_ERR_STATE;
  const Graph &g = h->g;
  cudaStream_t s = (cudaStream_t)stream;
  if (d_values) NALU_CUDA(cudaMemsetAsync(d_values, 0, size_t(g.nnz_owned() + g.nnz_shared()) * sizeof(double), s));
  if (d_rhs) NALU_CUDA(cudaMemsetAsync(d_rhs, 0, size_t(g.n_row_nodes()) * g.ndof * sizeof(double), s));
  return NALU_B200_OK;
}

int nalu_b200_assemble_elem(const nalu_b200_graph *h, const nalu_b200_kernel_desc *kernels, int32_t nk, double *d_values,
                            double *d_rhs, int32_t mode, int32_t overwrite, void *stream)
{
  // an empty part (no elements, hence no matrix entries) may come with null arrays: nothing is dereferenced
  if (!h || !kernels || (!d_values && h->g.nnz_owned() + h->g.nnz_shared() > 0) || (!d_rhs && h->g.n_row_nodes() > 0)) {
    set_error("assemble_elem: null argument");
    return NALU_B200_ERR_ARG;
  }
  if (!h->g.d_browptr) {
    set_error("assemble_elem: graph not finalized on a device");
    return NALU_B200_ERR_STATE;
  }
  const Graph &g = h->g;
  cudaStream_t s = (cudaStream_t)stream;
  if (g.n_elems == 0) { // AssembleElemSolverAlgorithm::execute over empty buckets (src/AssembleElemSolverAlgorithm.C:74-100)
    set_last_mode(0);
    return NALU_B200_OK;
  }
  if (g.npe == 27) { // Hex27 graphs (nalu_b200_graph_create_topo): one kernel, assemble_hex27.cu
    if (!hex27_supported(g, kernels, nk)) {
      set_error("assemble_elem: Hex27 graphs take [MomentumMassElemKernel(lumped)]? + MomentumAdvDiffElemKernel (default options, "
                "one velocity field) on a 3-dof system, [ContinuityAdvElemKernel] or [ScalarAdvDiffElemKernel] (default options) "
                "on a 1-dof system");
      return NALU_B200_ERR_ARG;
    }
    for (int i = 0; i < nk; ++i)
      if (const char *missing = missing_field(kernels[i])) {
        set_error(std::string("assemble_elem: field missing for kernel ") + std::to_string(i) + ": " + missing);
        return NALU_B200_ERR_ARG;
      }
    set_last_mode(5);
    return launch_assemble_hex27(g, kernels, nk, d_values, d_rhs, s);
  }
  // A list with one kernel the tiled write-once kernels know plus others (e.g. [MomentumMass(lumped), MomentumAdvDiff],
  // the momentum list of the low-Mach equation system, src/LowMachEquationSystem.C:1304-1398): the tiled kernel
  // takes its kernel first (it may overwrite), the others accumulate on top through the general kernel -- the sum the
  // reference forms in one scratch lhs/rhs (src/AssembleElemSolverAlgorithm.C:86-91), in another order.
  const bool split_lists = tuning().split_lists;
  // scalar transport lists (config 4: lumped mass + upwind adv-diff + SST source) in one tiled write-once launch
  if (mode != NALU_ASSEMBLE_GENERAL && tuning().cont_tiled && slist_box_supported(g, kernels, nk)) {
    for (int i = 0; i < nk; ++i)
      if (const char *missing = missing_field(kernels[i])) {
        set_error(std::string("assemble_elem: field missing for kernel ") + std::to_string(i) + ": " + missing);
        return NALU_B200_ERR_ARG;
      }
    set_last_mode(2);
    return launch_assemble_slist_box(g, kernels, nk, d_values, d_rhs, overwrite, s);
  }
  if (mode == NALU_ASSEMBLE_AUTO && nk > 1 && nk <= kMaxKernels && split_lists && !box_supported(g, kernels, nk)) {
    int t = -1;
    for (int i = 0; i < nk && t < 0; ++i)
      if (kernels[i].kind != NALU_K_TEST_FIXED_BLOCK &&
          (box_supported(g, kernels + i, 1) || cont_box_supported(g, kernels + i, 1) || scalar_box_supported(g, kernels + i, 1)))
        t = i;
    bool fixed = false; // TestKernel assigns its block (unit_tests/UnitTestTpetra.C:125-139): order matters, no split
    for (int i = 0; i < nk; ++i) fixed = fixed || kernels[i].kind == NALU_K_TEST_FIXED_BLOCK;
    if (t >= 0 && !fixed) {
      int rc = nalu_b200_assemble_elem(h, kernels + t, 1, d_values, d_rhs, NALU_ASSEMBLE_AUTO, overwrite, stream);
      if (rc != NALU_B200_OK) return rc;
      nalu_b200_kernel_desc rest[kMaxKernels];
      int nr = 0;
      for (int i = 0; i < nk; ++i)
        if (i != t) rest[nr++] = kernels[i];
      // lumped mass kernels need no element scratch: one thread per sub-control volume
      rc = scv_lumped_supported(g, rest, nr) ? launch_scv_lumped(g, rest, nr, d_values, d_rhs, s)
                                             : launch_assemble_general(g, rest, nr, d_values, d_rhs, s);
      set_last_mode(3);
      return rc;
    }
  }
  bool use_box = false;
  // the one-dof tiled kernel (ContinuityAdvElemKernel on a structured block); NALU_CONT_BOX=0 keeps the general kernel
  const bool cont_tiled = tuning().cont_tiled;
  if (mode != NALU_ASSEMBLE_GENERAL && cont_tiled && cont_box_supported(g, kernels, nk)) {
    set_last_mode(2);
    return launch_assemble_cont_box(g, kernels, nk, d_values, d_rhs, overwrite, s);
  }
  if (mode != NALU_ASSEMBLE_GENERAL && cont_tiled && scalar_box_supported(g, kernels, nk)) {
    set_last_mode(2);
    return launch_assemble_scalar_box(g, kernels, nk, d_values, d_rhs, overwrite, s);
  }
  if (mode == NALU_ASSEMBLE_BOX) {
    if (!box_supported(g, kernels, nk)) {
      set_error("assemble_elem: tiled kernel does not support this mesh / kernel list");
      return NALU_B200_ERR_ARG;
    }
    use_box = true;
  } else if (mode == NALU_ASSEMBLE_AUTO) {
    use_box = box_supported(g, kernels, nk);
  } else if (mode != NALU_ASSEMBLE_GENERAL) {
    set_error("assemble_elem: unknown mode");
    return NALU_B200_ERR_ARG;
  }
  if (use_box) {
    set_last_mode(2);
    return launch_assemble_box(g, kernels, nk, d_values, d_rhs, overwrite, s);
  }
  // any Hex8 connectivity with an element colouring: the momentum list through the row-thread unstructured kernel
  if (mode == NALU_ASSEMBLE_AUTO && tuning().urow && urow_supported(g, kernels, nk)) {
    set_last_mode(4);
    return launch_assemble_urow(g, kernels, nk, d_values, d_rhs, s);
  }
  set_last_mode(1);
  return launch_assemble_general(g, kernels, nk, d_values, d_rhs, s);
}

int nalu_b200_assemble_elem_zrange(const nalu_b200_graph *h, const nalu_b200_kernel_desc *kernels, int32_t nk,
                                   double *d_values, double *d_rhs, int32_t overwrite, int32_t z_begin, int32_t z_end,
                                   void *stream)
{
  if (!h || !kernels || !d_values || !d_rhs) {
    set_error("assemble_elem_zrange: null argument");
    return NALU_B200_ERR_ARG;
  }
  if (!h->g.d_browptr) {
    set_error("assemble_elem_zrange: graph not finalized on a device");
    return NALU_B200_ERR_STATE;
  }
  if (!box_supported(h->g, kernels, nk)) {
    set_error("assemble_elem_zrange: tiled kernel does not support this mesh / kernel list");
    return NALU_B200_ERR_ARG;
  }
  if (z_begin < 0) {
    set_error("assemble_elem_zrange: bad node-plane range");
    return NALU_B200_ERR_ARG;
  }
  set_last_mode(2);
  return launch_assemble_box(h->g, kernels, nk, d_values, d_rhs, overwrite, (cudaStream_t)stream, z_begin, z_end);
}

int nalu_b200_geometry_errors(const nalu_b200_graph *h, int32_t *flag)
{
  if (!h || !flag || !h->g.d_err) return NALU_B200_ERR_STATE;
  NALU_CUDA(cudaDeviceSynchronize());
  NALU_CUDA(cudaMemcpy(flag, h->g.d_err, sizeof(int), cudaMemcpyDeviceToHost));
  NALU_CUDA(cudaMemset(h->g.d_err, 0, sizeof(int)));
  return NALU_B200_OK;
}

int nalu_b200_assemble_node_mass(const nalu_b200_graph *h, const nalu_b200_kernel_desc *k, double *d_values, double *d_rhs,
                                 void *stream)
{
  if (!h || !k || !d_values || !d_rhs || !h->g.d_browptr) return NALU_B200_ERR_STATE;
  return launch_node_mass(h->g, *k, d_values, d_rhs, (cudaStream_t)stream);
}

int nalu_b200_sum_into(const nalu_b200_graph *h, int32_t n_entities, const int32_t *nodes, const double *rhs, const double *lhs,
                       double *d_values, double *d_rhs, void *stream)
{
  if (!h || !nodes || !rhs || !lhs || !d_values || !d_rhs || !h->g.d_browptr) return NALU_B200_ERR_STATE;
  return launch_sum_into(h->g, n_entities, nodes, rhs, lhs, d_values, d_rhs, (cudaStream_t)stream);
}

int nalu_b200_apply_dirichlet_bcs(const nalu_b200_graph *h, int32_t n_nodes, const int32_t *d_nodes, const double *d_solution,
                                  const double *d_bc_values, int32_t begin_pos, int32_t end_pos, double *d_values, double *d_rhs,
                                  void *stream)
{
  if (!h || !d_values || !d_rhs || !h->g.d_browptr) return NALU_B200_ERR_STATE;
  if (n_nodes > 0 && (!d_nodes || !d_solution || !d_bc_values)) return NALU_B200_ERR_ARG;
  return launch_row_fixup(h->g, n_nodes, d_nodes, begin_pos, end_pos, d_solution, d_bc_values, 1, d_values, d_rhs,
                          (cudaStream_t)stream);
}

int nalu_b200_reset_rows(const nalu_b200_graph *h, int32_t n_nodes, const int32_t *d_nodes, int32_t begin_pos, int32_t end_pos,
                         double *d_values, double *d_rhs, void *stream)
{
  if (!h || !d_values || !d_rhs || !h->g.d_browptr) return NALU_B200_ERR_STATE;
  if (n_nodes > 0 && !d_nodes) return NALU_B200_ERR_ARG;
  return launch_row_fixup(h->g, n_nodes, d_nodes, begin_pos, end_pos, nullptr, nullptr, 0, d_values, d_rhs, (cudaStream_t)stream);
}

int nalu_b200_mdot_scs(const nalu_b200_graph *h, const double *d_coords, const double *d_density, const double *d_velocity,
                       double *d_mdot, void *stream)
{
  if (h && d_coords && d_density && d_velocity && d_mdot && h->g.npe == 27 && h->g.d_elem_nodes27)
    return launch_mdot_hex27(h->g, d_coords, d_density, d_velocity, d_mdot, (cudaStream_t)stream);
  if (!h || !d_coords || !d_density || !d_velocity || !d_mdot || !h->g.d_elem_nodes) return NALU_B200_ERR_STATE;
  return launch_mdot(h->g, d_coords, d_density, d_velocity, d_mdot, (cudaStream_t)stream);
}

int nalu_b200_assemble_nodal_grad_elem(const nalu_b200_graph *h, const double *d_coords, const double *d_field, int32_t ncomp,
                                       const double *d_dual_volume, int32_t use_shifted, double *d_grad, void *stream)
{
  if (h && d_coords && d_field && d_dual_volume && d_grad && h->g.npe == 27 && h->g.d_elem_nodes27) {
    if (use_shifted) {
      set_error("assemble_nodal_grad_elem: Hex27 graphs take the unshifted shape functions only");
      return NALU_B200_ERR_ARG;
    }
    return launch_nodal_grad_hex27(h->g, d_coords, d_field, ncomp, d_dual_volume, d_grad, (cudaStream_t)stream);
  }
  if (!h || !d_coords || !d_field || !d_dual_volume || !d_grad || !h->g.d_elem_nodes) return NALU_B200_ERR_STATE;
  return launch_nodal_grad(h->g, d_coords, d_field, ncomp, d_dual_volume, use_shifted, d_grad, (cudaStream_t)stream);
}

int nalu_b200_compute_mdot_elem(const nalu_b200_graph *h, const double *d_coords, const double *d_density, const double *d_velocity,
                                const double *d_pressure, const double *d_gpdx, double proj_time_scale, int32_t options,
                                double *d_mdot, void *stream)
{
  if (h && d_coords && d_density && d_velocity && d_pressure && d_gpdx && d_mdot && h->g.npe == 27 && h->g.d_elem_nodes27) {
    if (options != 0) {
      set_error("compute_mdot_elem: Hex27 graphs take the default options only");
      return NALU_B200_ERR_ARG;
    }
    return launch_mdot_hex27(h->g, d_coords, d_density, d_velocity, d_mdot, (cudaStream_t)stream, d_pressure, d_gpdx, proj_time_scale);
  }
  if (!h || !d_coords || !d_density || !d_velocity || !d_pressure || !d_gpdx || !d_mdot || !h->g.d_elem_nodes) return NALU_B200_ERR_STATE;
  return launch_compute_mdot(h->g, d_coords, d_density, d_velocity, d_pressure, d_gpdx, proj_time_scale, options, d_mdot,
                             (cudaStream_t)stream);
}

int nalu_b200_compute_mdot_elem_range(const nalu_b200_graph *h, const double *d_coords, const double *d_density,
                                      const double *d_velocity, const double *d_pressure, const double *d_gpdx, double proj_time_scale,
                                      int32_t options, int64_t e_begin, int64_t e_end, double *d_mdot, void *stream)
{
  if (h && d_coords && d_density && d_velocity && d_pressure && d_gpdx && d_mdot && h->g.npe == 27 && h->g.d_elem_nodes27) {
    if (options != 0) {
      set_error("compute_mdot_elem_range: Hex27 graphs take the default options only");
      return NALU_B200_ERR_ARG;
    }
    return launch_mdot_hex27(h->g, d_coords, d_density, d_velocity, d_mdot, (cudaStream_t)stream, d_pressure, d_gpdx, proj_time_scale, e_begin,
                             e_end);
  }
  if (!h || !d_coords || !d_density || !d_velocity || !d_pressure || !d_gpdx || !d_mdot || !h->g.d_elem_nodes) return NALU_B200_ERR_STATE;
  return launch_compute_mdot(h->g, d_coords, d_density, d_velocity, d_pressure, d_gpdx, proj_time_scale, options, d_mdot,
                             (cudaStream_t)stream, e_begin, e_end);
}

int nalu_b200_halo_num_peers(const nalu_b200_graph *h, int32_t *ns, int32_t *nr)
{
  if (!h || !h->g.finalized) return NALU_B200_ERR_STATE;
  if (ns) *ns = (int32_t)h->g.send_peers.size();
  if (nr) *nr = (int32_t)h->g.recv_peers.size();
  return NALU_B200_OK;
}
int nalu_b200_halo_send_peer(const nalu_b200_graph *h, int32_t i, int32_t *peer, int64_t *n)
{
  if (!h || i < 0 || i >= (int)h->g.send_peers.size()) return NALU_B200_ERR_ARG;
  if (peer) *peer = h->g.send_peers[i].rank;
  if (n) *n = h->g.send_peers[i].n_value_doubles + h->g.send_peers[i].n_rhs_doubles;
  return NALU_B200_OK;
}
int nalu_b200_halo_recv_peer(const nalu_b200_graph *h, int32_t i, int32_t *peer, int64_t *n)
{
  if (!h || i < 0 || i >= (int)h->g.recv_peers.size()) return NALU_B200_ERR_ARG;
  if (peer) *peer = h->g.recv_peers[i].rank;
  if (n) *n = h->g.recv_peers[i].n_value_doubles + h->g.recv_peers[i].n_rhs_doubles;
  return NALU_B200_OK;
}
int nalu_b200_halo_pack(const nalu_b200_graph *h, int32_t i, const double *d_values, const double *d_rhs, double *d_buf,
                        void *stream)
{
  if (!h || i < 0 || i >= (int)h->g.send_peers.size() || !d_values || !d_rhs || !d_buf) return NALU_B200_ERR_ARG;
  return launch_halo_pack(h->g, h->g.send_peers[i], d_values, d_rhs, d_buf, (cudaStream_t)stream);
}
int nalu_b200_halo_unpack_add(const nalu_b200_graph *h, int32_t i, const double *d_buf, double *d_values, double *d_rhs,
                              void *stream)
{
  if (!h || i < 0 || i >= (int)h->g.recv_peers.size() || !d_values || !d_rhs || !d_buf) return NALU_B200_ERR_ARG;
  const HaloPeer &p = h->g.recv_peers[i];
  return launch_halo_add(h->g, p, d_buf, d_buf + p.n_value_doubles, false, d_values, d_rhs, (cudaStream_t)stream);
}
int nalu_b200_halo_pull_add(const nalu_b200_graph *h, int32_t i, const double *d_peer_values, const double *d_peer_rhs,
                            double *d_values, double *d_rhs, void *stream)
{
  if (!h || i < 0 || i >= (int)h->g.recv_peers.size() || !d_values || !d_rhs || !d_peer_values || !d_peer_rhs)
    return NALU_B200_ERR_ARG;
  return launch_halo_add(h->g, h->g.recv_peers[i], d_peer_values, d_peer_rhs, true, d_values, d_rhs, (cudaStream_t)stream);
}
int nalu_b200_halo_peer_offsets(const nalu_b200_graph *h, int32_t i, int64_t *vo, int64_t *ro)
{
  if (!h || i < 0 || i >= (int)h->g.recv_peers.size()) return NALU_B200_ERR_ARG;
  if (vo) *vo = h->g.recv_peers[i].peer_values_off;
  if (ro) *ro = h->g.recv_peers[i].peer_rhs_off;
  return NALU_B200_OK;
}

int nalu_b200_ipc_get_handle(const void *d_ptr, void *handle64)
{
  cudaIpcMemHandle_t hnd;
  NALU_CUDA(cudaIpcGetMemHandle(&hnd, const_cast<void *>(d_ptr)));
  static_assert(sizeof(hnd) == 64, "cudaIpcMemHandle_t is 64 bytes");
  std::memcpy(handle64, &hnd, 64);
  return NALU_B200_OK;
}
int nalu_b200_ipc_open_handle(const void *handle64, void **d_ptr)
{
  cudaIpcMemHandle_t hnd;
  std::memcpy(&hnd, handle64, 64);
  NALU_CUDA(cudaIpcOpenMemHandle(d_ptr, hnd, cudaIpcMemLazyEnablePeerAccess));
  return NALU_B200_OK;
}
int nalu_b200_ipc_close(void *d_ptr)
{
  NALU_CUDA(cudaIpcCloseMemHandle(d_ptr));
  return NALU_B200_OK;
}

int nalu_b200_malloc(void **d_ptr, size_t nbytes)
{
  NALU_CUDA(cudaMalloc(d_ptr, nbytes ? nbytes : 8));
  return NALU_B200_OK;
}
int nalu_b200_free(void *d_ptr)
{
  NALU_CUDA(cudaFree(d_ptr));
  return NALU_B200_OK;
}
int nalu_b200_memcpy_h2d(void *d, const void *hsrc, size_t n, void *stream)
{
  NALU_CUDA(cudaMemcpyAsync(d, hsrc, n, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return NALU_B200_OK;
}
int nalu_b200_memcpy_d2h(void *hdst, const void *d, size_t n, void *stream)
{
  NALU_CUDA(cudaMemcpyAsync(hdst, d, n, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return NALU_B200_OK;
}
int nalu_b200_stream_sync(void *stream)
{
  NALU_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return NALU_B200_OK;
}

} // extern "C"
