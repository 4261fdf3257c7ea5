// graph.h -- internal representation behind the opaque nalu_b200_graph handle.
//
// Replaces TpetraLinearSystem's DOF numbering and CSR graph construction
// (src/TpetraLinearSystem.C:190-327, 392-420, 1257-1353 of the reference) with a sort-based
// node-block builder: the reference grows per-row std::vectors and does sorted inserts of
// every (row, col) pair; here each row node gathers the nodes of its incident elements once,
// sorts, uniques, and the dof-level rowptr/colidx are expanded on the device.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace nalu_b200 {

struct HaloPeer {
  int32_t rank = -1;
  int64_t n_value_doubles = 0; // CSR values exchanged with this peer
  int64_t n_rhs_doubles = 0;
  // send side: my shared row nodes owned by `rank` (indices into the shared row-node list)
  std::vector<int32_t> send_nodes;
  int32_t *d_send_nodes = nullptr;
  int64_t *d_send_node_packed_off = nullptr; // [n_send_nodes+1] offsets (doubles) in the packed buffer
  // recv side: one entry per (peer row node, block)
  int64_t n_segments = 0;
  int64_t *d_src_packed = nullptr;  // offset of the (i=0) ndof-segment in the packed buffer
  int64_t *d_src_inplace = nullptr; // same, relative to the start of the peer's shared values
  int32_t *d_src_stride = nullptr;  // doubles between dof rows on the source side
  int64_t *d_dst = nullptr;         // offset of the (i=0) segment in my values
  int32_t *d_dst_stride = nullptr;
  int64_t n_rhs_nodes = 0;
  int64_t *d_rhs_src_inplace = nullptr; // per node: offset in the peer's shared rhs
  int64_t *d_rhs_dst = nullptr;         // per node: offset in my rhs
  int64_t peer_values_off = 0, peer_rhs_off = 0; // first row of this peer pair inside peer's shared block
  double *d_buf = nullptr; // persistent exchange buffer of nalu_b200_load_complete (values, then rhs), allocated on first use
  // host staging of the maps above (uploaded by graph_upload, then released)
  std::vector<int64_t> h_send_node_packed_off;
  std::vector<int64_t> h_src_packed, h_src_inplace, h_dst, h_rhs_src_inplace, h_rhs_dst;
  std::vector<int32_t> h_src_stride, h_dst_stride;
};

// elements per chunk of the (chunk, colour) launch order of the general kernel.  Measured on B200 at 128^3 (momentum
// adv-diff, general kernel): one chunk (8 launches, every colour streams the whole matrix) 60 Melem/s; 16 K-element chunks
// (L2-resident rows, 512 small launches) 46 Melem/s -- the kernel is bound by its per-element latency, not by DRAM, so
// the launches must stay large.  One chunk it is.
constexpr int64_t kColorChunk = int64_t(1) << 40;

struct Graph {
  int32_t rank = 0, nranks = 1, ndof = 1;
  int32_t n_nodes = 0;
  int64_t n_elems = 0;
  int32_t npe = 8; // nodes per element: 8 (Hex8) or 27 (Hex27, nalu_b200_graph_create_topo; assemble_hex27.cu only)
  int32_t box[3] = {0, 0, 0};
  std::vector<int64_t> gid;
  std::vector<int32_t> owner;
  std::vector<int32_t> elem_nodes;

  // numbering: row nodes = unique owned ids (ascending) then unique shared-not-owned ids
  int32_t n_owned_nodes = 0, n_shared_nodes = 0;
  std::vector<int32_t> node_row;     // [n_nodes] local node -> row node
  std::vector<int32_t> row_rep;      // [n_row_nodes] lowest local node of the row (several nodes alias a row on periodic meshes)
  std::vector<int64_t> row_node_gid; // [n_owned+n_shared]
  std::vector<int32_t> shared_owner; // [n_shared]

  // element colouring for the general kernel's scatter: elements of one colour share no row node, so a colour is summed
  // into the CSR with plain read-modify-writes, colours in a fixed order (deterministic; no atomics).  color_elems =
  // element ids sorted by (chunk of kColorChunk consecutive ids, colour), color_ptr[g] .. color_ptr[g+1] the elements of
  // group g = chunk * n_colors + colour; empty: no colouring (an element touches one row twice, or more than 64
  // colours) -> atomic scatter.
  std::vector<int32_t> color_elems;
  std::vector<int64_t> color_ptr;
  int32_t n_colors = 0;
  int64_t color_chunk = kColorChunk; // elements per chunk of the (chunk, colour) launch order; set before the numbering phase
  int32_t *d_color_elems = nullptr;

  // local adjacency (row node -> sorted unique row nodes), before the handshake
  std::vector<int64_t> adj_ptr;
  std::vector<int32_t> adj;

  // handshake
  std::vector<std::vector<int64_t>> send_msg, recv_msg;
  std::vector<char> recv_have;
  bool finalized = false;

  // column map: [owned row nodes][shared row nodes][remote nodes sorted by (owner,gid)]
  std::vector<int64_t> remote_gid;
  std::vector<int32_t> remote_owner;

  // final node-block graph; owned rows [0,n_owned), shared rows after
  std::vector<int64_t> browptr; // [n_row_nodes+1]
  std::vector<int32_t> bcol;    // block column = column LID / ndof
  int64_t nblk_owned = 0, nblk_shared = 0;

  // device
  int32_t *d_elem_nodes = nullptr; // [n_elems][8]; null on Hex27 graphs, so that every Hex8 entry point refuses them
  int32_t *d_elem_nodes27 = nullptr; // Hex27 graphs: [n_elems][27] (STK node order)
  uint8_t *d_elem_slot27 = nullptr;  // Hex27 graphs: [n_elems][27][27] slot of column node ic in the row of node k
  int32_t *d_node_row = nullptr;   // [n_nodes]
  int32_t *d_row_rep = nullptr;    // [n_row_nodes]
  int64_t *d_browptr = nullptr;    // [n_row_nodes+1]
  int32_t *d_bcol = nullptr;
  uint8_t *d_elem_slot = nullptr;  // [n_elems][8][8]
  // transpose of the connectivity for the row-owner kernel (assemble_urow.cu), non-lattice 3-dof graphs only: the elements
  // around row node r are rowelem[rowelem_ptr[r] .. rowelem_ptr[r+1]), each packed as element << 3 | local node
  int64_t *d_rowelem_ptr = nullptr;
  int32_t *d_rowelem = nullptr;
  uint8_t *d_node_slot27 = nullptr; // box only: [n_nodes][27] slot of lattice neighbour (dz,dy,dx)
  // box only: [n_nodes][3] row descriptor of the tiled kernels: [0] first value of the node's rows (ndof^2 browptr[row]),
  // [1] row | blocks << 32 | natural << 40 (all 27 lattice neighbours present, in lattice order) | 1 << 41,
  // [2] bytes 0..2: first slot of the dz = -1, 0, +1 column group, bytes 3..5: columns in the group
  std::vector<uint64_t> node_desc;
  uint64_t *d_node_desc = nullptr;
  uint64_t *d_rowptr_owned = nullptr, *d_rowptr_shared = nullptr;
  int32_t *d_colidx_owned = nullptr, *d_colidx_shared = nullptr;
  int32_t *d_err = nullptr; // device error flag (negative Jacobians)

  std::vector<HaloPeer> send_peers, recv_peers;

  int64_t nnz_owned() const { return nblk_owned * ndof * ndof; }
  int64_t nnz_shared() const { return nblk_shared * ndof * ndof; }
  int32_t n_row_nodes() const { return n_owned_nodes + n_shared_nodes; }
};

// host-side phases (graph.cpp)
int graph_number_and_adjacency(Graph &g, std::string &err);
int graph_build_send_messages(Graph &g, std::string &err);
int graph_finalize_host(Graph &g, std::vector<uint8_t> &elem_slot, std::vector<uint8_t> &node_slot27,
                        std::string &err);

} // namespace nalu_b200
