// halo.cu -- loadComplete: add the shared-not-owned rows into their owners' rows.
//
// Reference: TpetraLinearSystem::loadComplete (src/TpetraLinearSystem.C:1757-1778) does
// ownedMatrix_->doExport(*sharedNotOwnedMatrix_, *exporter_, Tpetra::ADD) and the same for the
// RHS: pack on the sender, MPI, unpack + combine on the owner, with global-id lookups per entry.
// Here the row/column correspondence is resolved once, during the graph handshake, into
// (source offset, destination offset) pairs per ndof x ndof block, so the data path is:
//   * pull:  the owner's kernel reads the sharer's CSR values directly through a peer
//            (NVLink P2P / CUDA-IPC mapped) pointer and adds them in place -- one kernel,
//            no pack, no staging, no unpack; or
//   * packed: sender gathers its rows into a contiguous buffer (halo_pack_kernel), any
//            transport moves it (NCCL send/recv), the owner runs the same add kernel on it.
// Each destination entry is touched by exactly one source entry per peer, so plain
// read-modify-write suffices (launches for different peers are ordered on the stream).
#include "internal.h"

namespace nalu_b200 {

__global__ void halo_pack_kernel(int n_send_nodes, const int *send_nodes, const long long *packed_off,
                                 const long long *browptr, int n_owned_nodes, int ndof, long long n_value_doubles,
                                 const double *values, const double *rhs, double *buf)
{
  for (int t = blockIdx.x; t < n_send_nodes; t += gridDim.x) {
    const int r = n_owned_nodes + send_nodes[t];
    const long long src = (long long)ndof * ndof * browptr[r];
    const long long n = packed_off[t + 1] - packed_off[t];
    double *dst = buf + packed_off[t];
    for (long long q = threadIdx.x; q < n; q += blockDim.x) dst[q] = values[src + q];
    if (threadIdx.x < ndof) buf[n_value_doubles + (long long)t * ndof + threadIdx.x] = rhs[(long long)r * ndof + threadIdx.x];
  }
}

__global__ void halo_add_kernel(long long n_segments, int ndof, const long long *src_off, const int *src_stride,
                                const long long *dst_off, const int *dst_stride, const double *__restrict__ src,
                                double *dst)
{
  const int d2 = ndof * ndof;
  const long long total = n_segments * d2;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long m = t / d2;
    const int ij = (int)(t - m * d2);
    const int i = ij / ndof, j = ij - i * ndof;
    dst[dst_off[m] + (long long)i * dst_stride[m] + j] += src[src_off[m] + (long long)i * src_stride[m] + j];
  }
}

__global__ void halo_add_rhs_kernel(long long n_nodes, int ndof, const long long *src_off, long long packed_base,
                                    const long long *dst_off, const double *__restrict__ src, double *dst)
{
  const long long total = n_nodes * ndof;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long m = t / ndof;
    const int i = (int)(t - m * ndof);
    const long long s = src_off ? src_off[m] + i : packed_base + t;
    dst[dst_off[m] + i] += src[s];
  }
}

int launch_halo_pack(const Graph &g, const HaloPeer &p, const double *d_values, const double *d_rhs, double *d_buf,
                     cudaStream_t stream)
{
  const int n = (int)p.send_nodes.size();
  if (n == 0) return NALU_B200_OK;
  const int ctas = n < 148 * 8 ? n : 148 * 8;
  halo_pack_kernel<<<ctas, 128, 0, stream>>>(n, p.d_send_nodes, (const long long *)p.d_send_node_packed_off,
                                             (const long long *)g.d_browptr, g.n_owned_nodes, g.ndof,
                                             p.n_value_doubles, d_values, d_rhs, d_buf);
  count_launch();
  NALU_CUDA(cudaGetLastError());
  return NALU_B200_OK;
}

int launch_halo_add(const Graph &g, const HaloPeer &p, const double *src_values, const double *src_rhs, bool inplace,
                    double *d_values, double *d_rhs, cudaStream_t stream)
{
  if (p.n_segments > 0) {
    const long long total = p.n_segments * g.ndof * g.ndof;
    long long ctas = (total + 255) / 256;
    if (ctas > 148 * 16) ctas = 148 * 16;
    halo_add_kernel<<<(unsigned)ctas, 256, 0, stream>>>(p.n_segments, g.ndof,
                                                        (const long long *)(inplace ? p.d_src_inplace : p.d_src_packed),
                                                        p.d_src_stride, (const long long *)p.d_dst, p.d_dst_stride,
                                                        src_values, d_values);
    count_launch();
  }
  if (p.n_rhs_nodes > 0) {
    const long long total = p.n_rhs_nodes * g.ndof;
    long long ctas = (total + 255) / 256;
    if (ctas > 148 * 16) ctas = 148 * 16;
    halo_add_rhs_kernel<<<(unsigned)ctas, 256, 0, stream>>>(p.n_rhs_nodes, g.ndof,
                                                            inplace ? (const long long *)p.d_rhs_src_inplace : nullptr,
                                                            0, (const long long *)p.d_rhs_dst, src_rhs, d_rhs);
    count_launch();
  }
  NALU_CUDA(cudaGetLastError());
  return NALU_B200_OK;
}

} // namespace nalu_b200
