// comm.cpp -- loadComplete inside the library: the shared-row halo sum over NCCL send / recv, the one collective of the
// path (TpetraLinearSystem::loadComplete = fillComplete + Export(ADD) of matrix and rhs, src/TpetraLinearSystem.C:1757-1778).
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy a host process such as PyTorch has already loaded, else the
// system one), so the library loads and every single-rank entry point works on a box without NCCL.  The communicator is
// the library's own (ncclCommInitRank from a unique id the caller distributes), the exchange buffers are persistent in the
// graph handle, and everything is enqueued on the caller's stream: the caller overlaps it with the interior assembly by
// assembling the interface planes first (nalu_b200_assemble_elem_zrange) and handing this call a second stream.
#include "internal.h"

#include <dlfcn.h>
#include <cstring>
#include <mutex>

namespace nalu_b200 {

namespace {
// the handful of NCCL entry points used, with NCCL's own signatures (nccl.h)
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
constexpr int kNcclFloat64 = 8; // ncclDouble

struct Nccl {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

Nccl &nccl()
{
  static Nccl n;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
      n.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (n.lib) break;
    }
    if (!n.lib) return;
    auto sym = [&](const char *s) { return dlsym(n.lib, s); };
    n.GetUniqueId = (decltype(n.GetUniqueId))sym("ncclGetUniqueId");
    n.CommInitRank = (decltype(n.CommInitRank))sym("ncclCommInitRank");
    n.CommDestroy = (decltype(n.CommDestroy))sym("ncclCommDestroy");
    n.GroupStart = (decltype(n.GroupStart))sym("ncclGroupStart");
    n.GroupEnd = (decltype(n.GroupEnd))sym("ncclGroupEnd");
    n.Send = (decltype(n.Send))sym("ncclSend");
    n.Recv = (decltype(n.Recv))sym("ncclRecv");
    n.GetErrorString = (decltype(n.GetErrorString))sym("ncclGetErrorString");
    n.ok = n.GetUniqueId && n.CommInitRank && n.CommDestroy && n.GroupStart && n.GroupEnd && n.Send && n.Recv;
  });
  return n;
}

int nccl_fail(ncclResult_t r, const char *what)
{
  Nccl &n = nccl();
  set_error(std::string(what) + ": " + (n.GetErrorString ? n.GetErrorString(r) : "NCCL error"));
  return NALU_B200_ERR_CUDA;
}
#define NALU_NCCL(call)                                   \
  do {                                                    \
    ncclResult_t _r = (call);                             \
    if (_r != 0) return nccl_fail(_r, #call);             \
  } while (0)
} // namespace

} // namespace nalu_b200

struct nalu_b200_comm {
  nalu_b200::ncclComm_t comm = nullptr;
  int nranks = 0, rank = 0;
};

using namespace nalu_b200;

extern "C" {

int nalu_b200_comm_get_unique_id(void *id128)
{
  if (!id128) return NALU_B200_ERR_ARG;
  Nccl &n = nccl();
  if (!n.ok) {
    set_error("comm_get_unique_id: libnccl.so.2 not found");
    return NALU_B200_ERR_STATE;
  }
  ncclUniqueId id;
  NALU_NCCL(n.GetUniqueId(&id));
  std::memcpy(id128, &id, sizeof(id));
  return NALU_B200_OK;
}

int nalu_b200_comm_create(nalu_b200_comm **out, int32_t nranks, int32_t rank, const void *id128)
{
  if (!out || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return NALU_B200_ERR_ARG;
  Nccl &n = nccl();
  if (!n.ok) {
    set_error("comm_create: libnccl.so.2 not found");
    return NALU_B200_ERR_STATE;
  }
  ncclUniqueId id;
  std::memcpy(&id, id128, sizeof(id));
  nalu_b200_comm *c = new nalu_b200_comm();
  c->nranks = nranks;
  c->rank = rank;
  ncclResult_t r = n.CommInitRank(&c->comm, nranks, id, rank);
  if (r != 0) {
    delete c;
    return nccl_fail(r, "ncclCommInitRank");
  }
  *out = c;
  return NALU_B200_OK;
}

void nalu_b200_comm_destroy(nalu_b200_comm *c)
{
  if (!c) return;
  if (c->comm) nccl().CommDestroy(c->comm);
  delete c;
}

int nalu_b200_load_complete(nalu_b200_graph *h, nalu_b200_comm *c, double *d_values, double *d_rhs, void *stream)
{
  if (!h || !d_values || !d_rhs || !h->g.d_browptr) return NALU_B200_ERR_STATE;
  Graph &g = h->g;
  if (g.send_peers.empty() && g.recv_peers.empty()) return NALU_B200_OK; // a single rank owns every row it touches
  if (!c || !c->comm) {
    set_error("load_complete: rows are shared with other ranks but no communicator was given");
    return NALU_B200_ERR_ARG;
  }
  Nccl &n = nccl();
  cudaStream_t s = (cudaStream_t)stream;
  // persistent exchange buffers, allocated at the first call
  for (auto *peers : {&g.send_peers, &g.recv_peers})
    for (HaloPeer &p : *peers)
      if (!p.d_buf) NALU_CUDA(cudaMalloc((void **)&p.d_buf, sizeof(double) * (size_t)(p.n_value_doubles + p.n_rhs_doubles + 1)));
  for (HaloPeer &p : g.send_peers) {
    int rc = launch_halo_pack(g, p, d_values, d_rhs, p.d_buf, s);
    if (rc) return rc;
  }
  NALU_NCCL(n.GroupStart());
  for (HaloPeer &p : g.send_peers)
    NALU_NCCL(n.Send(p.d_buf, (size_t)(p.n_value_doubles + p.n_rhs_doubles), kNcclFloat64, p.rank, c->comm, s));
  for (HaloPeer &p : g.recv_peers)
    NALU_NCCL(n.Recv(p.d_buf, (size_t)(p.n_value_doubles + p.n_rhs_doubles), kNcclFloat64, p.rank, c->comm, s));
  NALU_NCCL(n.GroupEnd());
  for (HaloPeer &p : g.recv_peers) {
    int rc = launch_halo_add(g, p, p.d_buf, p.d_buf + p.n_value_doubles, false, d_values, d_rhs, s);
    if (rc) return rc;
  }
  return NALU_B200_OK;
}

} // extern "C"
