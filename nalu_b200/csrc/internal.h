// internal.h -- glue shared by the translation units of libnalu_b200.so
#pragma once
#include "../../include/nalu_b200.h"
#include "graph.h"
#include <cuda_runtime.h>
#include <string>

struct nalu_b200_graph {
  nalu_b200::Graph g;
};

namespace nalu_b200 {

void set_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what);
void count_launch(int n = 1);
void set_last_mode(int m);

#define NALU_CUDA(call)                                                        \
  do {                                                                         \
    cudaError_t _e = (call);                                                   \
    if (_e != cudaSuccess) return ::nalu_b200::cuda_fail(_e, #call);           \
  } while (0)

// assemble_general.cu
int launch_assemble_general(const Graph &g, const nalu_b200_kernel_desc *kernels, int nk, double *d_values,
                            double *d_rhs, cudaStream_t stream);
const char *missing_field(const nalu_b200_kernel_desc &k);
int launch_node_mass(const Graph &g, const nalu_b200_kernel_desc &k, double *d_values, double *d_rhs,
                     cudaStream_t stream);
bool cont_box_supported(const Graph &g, const nalu_b200_kernel_desc *k, int nk);
int launch_assemble_cont_box(const Graph &g, const nalu_b200_kernel_desc *k, int nk, double *d_values, double *d_rhs,
                             int overwrite, cudaStream_t stream);
bool scalar_box_supported(const Graph &g, const nalu_b200_kernel_desc *k, int nk);
int launch_assemble_scalar_box(const Graph &g, const nalu_b200_kernel_desc *k, int nk, double *d_values, double *d_rhs,
                               int overwrite, cudaStream_t stream);
// assemble_slist_box.cu: scalar kernel lists (lumped mass + (upwind) adv-diff + SST source) in one tiled launch
bool slist_box_supported(const Graph &g, const nalu_b200_kernel_desc *k, int nk);
int launch_assemble_slist_box(const Graph &g, const nalu_b200_kernel_desc *k, int nk, double *d_values, double *d_rhs,
                              int overwrite, cudaStream_t stream);
// assemble_urow.cu: [MomentumMass(lumped)]? + MomentumAdvDiff on any coloured Hex8 connectivity, thread = (element, row node)
bool urow_supported(const Graph &g, const nalu_b200_kernel_desc *k, int nk);
int launch_assemble_urow(const Graph &g, const nalu_b200_kernel_desc *k, int nk, double *d_values, double *d_rhs, cudaStream_t stream);
bool scv_lumped_supported(const Graph &g, const nalu_b200_kernel_desc *k, int nk);
int launch_scv_lumped(const Graph &g, const nalu_b200_kernel_desc *k, int nk, double *d_values, double *d_rhs, cudaStream_t stream);
int launch_nodal_grad(const Graph &g, const double *coords, const double *field, int ncomp, const double *dual_volume, int use_shifted,
                      double *grad, cudaStream_t stream);
int launch_compute_mdot(const Graph &g, const double *coords, const double *density, const double *velocity, const double *pressure,
                        const double *gpdx, double tau, int options, double *mdot, cudaStream_t stream, long long e0 = 0,
                        long long e1 = -1);
int launch_row_fixup(const Graph &g, int n_list, const int *d_list, int begin_pos, int end_pos, const double *d_solution,
                     const double *d_bc_values, int dirichlet, double *d_values, double *d_rhs, cudaStream_t stream);
int launch_sum_into(const Graph &g, int n_entities, const int *h_nodes, const double *h_rhs, const double *h_lhs,
                    double *d_values, double *d_rhs, cudaStream_t stream);
int launch_mdot(const Graph &g, const double *coords, const double *density, const double *velocity,
                double *mdot, cudaStream_t stream);
// assemble_box.cu
bool box_supported(const Graph &g, const nalu_b200_kernel_desc *kernels, int nk);
int launch_assemble_box(const Graph &g, const nalu_b200_kernel_desc *kernels, int nk, double *d_values,
                        double *d_rhs, int overwrite, cudaStream_t stream, int z_begin = -1, int z_end = -1);
// assemble_box2.cu (column-pair kernel: option combinations the row-thread kernel does not take), assemble_box5.cu
// (row-thread kernel, the default)
namespace box { struct BoxParams; }
int launch_assemble_box2(box::BoxParams &P, cudaStream_t stream);
bool box5_supported(const box::BoxParams &P);
int launch_assemble_box5(box::BoxParams &P, cudaStream_t stream);
// assemble_hex27.cu: MomentumAdvDiffElemKernel<AlgTraitsHex27>, CTA = element, coloured value-less reductions
bool hex27_supported(const Graph &g, const nalu_b200_kernel_desc *k, int nk);
int launch_assemble_hex27(const Graph &g, const nalu_b200_kernel_desc *k, int nk, double *d_values, double *d_rhs, cudaStream_t stream);
int launch_mdot_hex27(const Graph &g, const double *coords, const double *density, const double *velocity, double *mdot,
                      cudaStream_t stream, const double *pressure = nullptr, const double *gpdx = nullptr, double tau = 0.0,
                      long long e0 = 0, long long e1 = -1);
int launch_nodal_grad_hex27(const Graph &g, const double *coords, const double *field, int ncomp, const double *dual_volume, double *grad,
                            cudaStream_t stream);
// halo.cu
int launch_halo_pack(const Graph &g, const HaloPeer &p, const double *d_values, const double *d_rhs,
                     double *d_buf, cudaStream_t stream);
int launch_halo_add(const Graph &g, const HaloPeer &p, const double *src_values, const double *src_rhs,
                    bool inplace, double *d_values, double *d_rhs, cudaStream_t stream);

constexpr int kMaxKernels = 8;

// Tuning knobs for A/B measurements, read ONCE from the environment when the library is first used (never on the
// launch path): NALU_BOX_KERNEL=2 column-pair kernel instead of the row-thread kernel, NALU_BOX_ZCHUNK=n node planes
// per CTA, NALU_SPLIT_LISTS=0 multi-kernel lists through the general kernel only, NALU_CONT_BOX=0 one-dof lists through
// the general kernel.  None of them changes what is computed.
struct Tuning {
  int box_kernel = 5;
  int box_zchunk = 0;
  bool split_lists = true;
  bool cont_tiled = true;
  int general_minb = 4; // CTAs per SM the general kernel's registers are bounded for (2, 3 or 4)
  // coloured scatter of the general kernel: 0 = load / add / store, 1 = fire-and-forget reduction (red.global.add.f64).
  // Either way exactly one element of a colour touches an address, so the result is the same sum in the same order.
  int general_scatter = 1;
  int color_chunk = 0; // elements per (chunk, colour) launch group of graphs created from now on; 0 = one chunk
  int urow_minb = 2;   // CTAs per SM the row-thread unstructured kernel's registers are bounded for
  bool urow_owner = true; // row-owner form (one launch, thread = row node) when the graph carries the transposed connectivity
  bool urow = true; // momentum lists on non-lattice meshes: row-thread unstructured kernel (assemble_urow.cu) instead of the general one
};
const Tuning &tuning();

} // namespace nalu_b200
