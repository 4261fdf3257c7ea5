// assemble_general.cu -- general Hex8 element assembly: any connectivity, any kernel list.
//
// One warp per element (grid-stride), element lhs/rhs in the warp's shared-memory scratch
// (general_elem.cuh), then the sumInto step: every lhs entry goes straight to its CSR slot
// through the per-element slot map built with the graph -- the device analogue of
// TpetraLinearSystem::sumInto's sorted row merge (src/TpetraLinearSystem.C:1376-1522), with the sort
// and the linear column search replaced by the precomputed map.  The elements are processed colour
// by colour (graph.cpp: elements of one colour share no row node), one launch per colour, so the sum
// is a plain read-modify-write in a fixed order: deterministic, no atomics.  (Meshes the colouring
// gives up on -- an element touching one row twice -- fall back to atomicAdd.)  This is the path for
// unstructured meshes, periodic aliasing and for kernel lists the tiled kernels do not cover; it
// always accumulates.
#include "general_elem.cuh"
#include "box_tile.cuh"
#include "internal.h"
#include <initializer_list>
#include <string>

namespace nalu_b200 {

struct GeneralParams {
  long long n_elems;          // elements this launch processes
  const int *elem_list;       // their ids (one colour), or null: 0 .. n_elems-1 with an atomic scatter
  int red_scatter;            // coloured scatter by fire-and-forget reductions instead of load / add / store
  const int *elem_nodes;
  const int *node_row;
  const long long *browptr;
  const unsigned char *elem_slot;
  double *values;
  double *rhs;
  int *err;
  int nk;
  nalu_b200_kernel_desc k[kMaxKernels];
};

constexpr int kWarpsPerCta = 4;

// value-less reduction: one add at the L2, nothing returned to the SM
__device__ __forceinline__ void red_add(double *p, double v) { asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }

// MINB = CTAs per SM the register allocation is bounded for (2: 255 registers, 3: 168, 4: 128)
template <int NDOF, int MINB>
__global__ void __launch_bounds__(kWarpsPerCta * 32, MINB) assemble_general_kernel(const __grid_constant__ GeneralParams p)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ElemScratch<NDOF> *all = reinterpret_cast<ElemScratch<NDOF> *>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  ElemScratch<NDOF> &ws = all[warp];
  constexpr int R = 8 * NDOF;
  const long long warp0 = (long long)blockIdx.x * kWarpsPerCta + warp;
  const long long nwarps = (long long)gridDim.x * kWarpsPerCta;

  MeNeeds need = {0, 0, 0, 0};
  const double *coord_field = nullptr;
  for (int i = 0; i < p.nk; ++i) {
    me_needs_of(p.k[i], need);
    if (p.k[i].nodal[NALU_F_COORDINATES]) coord_field = p.k[i].nodal[NALU_F_COORDINATES];
  }

  for (long long t = warp0; t < p.n_elems; t += nwarps) {
    const long long e = p.elem_list ? p.elem_list[t] : t;
    if (lane < 8) {
      const int node = p.elem_nodes[e * 8 + lane];
      const int rn = p.node_row[node];
      const long long b0 = p.browptr[rn], b1 = p.browptr[rn + 1];
      ws.nodes[lane] = node;
      ws.rownode[lane] = rn;
      ws.rownb[lane] = (int)(b1 - b0);
      ws.rowbase[lane] = (long long)NDOF * NDOF * b0;
      if (coord_field) {
        ws.coords[lane][0] = coord_field[(size_t)node * 3 + 0];
        ws.coords[lane][1] = coord_field[(size_t)node * 3 + 1];
        ws.coords[lane][2] = coord_field[(size_t)node * 3 + 2];
      }
    }
    if (lane == 0) ws.bad_det = 0;
    for (int q = lane; q < R * R; q += 32) ws.lhs[q] = 0.0; // set_zero(simdlhs)
    if (lane < R) ws.rhs[lane] = 0.0;                        // set_zero(simdrhs)
    __syncwarp();
    stage_geometry(lane, ws, need);
    __syncwarp();
    for (int i = 0; i < p.nk; ++i) {
      const nalu_b200_kernel_desc &k = p.k[i];
      stage_gather(lane, ws, k);
      __syncwarp();
      stage_ip(lane, ws, k, e);
      __syncwarp();
      stage_lhs(lane, ws, k);
      stage_rhs(lane, ws, k);
      __syncwarp();
    }
    if (lane == 0 && ws.bad_det) atomicExch(p.err, 1);
    // sumInto: lhs[r][c] -> values[rowbase(k) + i*nb*NDOF + slot(k,ic)*NDOF + j]
    const unsigned char *slots = p.elem_slot + e * 64;
    auto slot_of = [&](int q) -> long long {
      const int r = q / R, c = q - r * R;
      const int kn = r / NDOF, i = r - kn * NDOF;
      const int ic = c / NDOF, j = c - ic * NDOF;
      return ws.rowbase[kn] + (long long)i * ws.rownb[kn] * NDOF + slots[kn * 8 + ic] * NDOF + j;
    };
    if (p.elem_list && p.red_scatter) {
      // No other element of this colour touches these rows and the entries of one element are distinct: every address
      // receives exactly one add per launch, launches (colours) are stream-ordered -> the same sum in the same order as
      // the read-modify-write below, bit for bit.  The reduction instruction does not return a value, so the warp does
      // not wait a DRAM round trip per batch of entries (the RMW form's long-scoreboard stall: 32 % of its samples).
      for (int q = lane; q < R * R; q += 32) {
        const double v = ws.lhs[q];
        if (v != 0.0) red_add(p.values + slot_of(q), v);
      }
    } else if (p.elem_list) {
      // plain read-modify-write, the loads of a batch issued together (nothing aliases inside a batch)
      constexpr int NIT = (R * R + 31) / 32, B = NIT % 6 == 0 ? 6 : 2;
#pragma unroll 1
      for (int it = 0; it < NIT; it += B) {
        double v[B], old[B];
        long long pos[B];
#pragma unroll
        for (int u = 0; u < B; ++u) {
          const int q = lane + 32 * (it + u);
          v[u] = q < R * R ? ws.lhs[q] : 0.0;
          pos[u] = v[u] != 0.0 ? slot_of(q) : -1;
        }
#pragma unroll
        for (int u = 0; u < B; ++u)
          if (pos[u] >= 0) old[u] = p.values[pos[u]];
#pragma unroll
        for (int u = 0; u < B; ++u)
          if (pos[u] >= 0) p.values[pos[u]] = old[u] + v[u];
      }
    } else {
      for (int q = lane; q < R * R; q += 32) {
        const double v = ws.lhs[q];
        if (v != 0.0) atomicAdd(p.values + slot_of(q), v);
      }
    }
    if (lane < R) {
      const double v = ws.rhs[lane];
      double *r = p.rhs + (long long)ws.rownode[lane / NDOF] * NDOF + lane % NDOF;
      if (v != 0.0) {
        if (p.elem_list && p.red_scatter) red_add(r, v);
        else if (p.elem_list) *r += v;
        else atomicAdd(r, v);
      }
    }
    __syncwarp();
  }
}

// the nodal / element fields each kernel kind dereferences (its add_gathered_nodal_field / add_element_field
// registrations in the reference's constructors); a missing one is an argument error on the host, not an illegal
// address on the device.  Returns the name of the first missing field, or null.
const char *missing_field(const nalu_b200_kernel_desc &k)
{
  static const char *names[NALU_F_COUNT] = {"coordinates", "velocity", "velocity(N)", "velocity(NM1)", "viscosity", "density",
                                            "density(N)", "density(NM1)", "pressure", "dpdx", "scalarQ", "scalarQ(N)",
                                            "scalarQ(NM1)", "dual_nodal_volume", "temperature", "turbulent_ke",
                                            "specific_dissipation_rate", "turbulent_viscosity", "sst_f_one_blending", "Gjq",
                                            "dudx", "velocity_rtm", "sst_max_length_scale", "c_epsilon", "dsqrtk_dx_sq", ""};
  int need[12];
  int n = 0;
  bool mdot = false;
  auto add = [&](std::initializer_list<int> l) { for (int f : l) need[n++] = f; };
  switch (k.kind) {
  case NALU_K_MOMENTUM_ADV_DIFF: add({NALU_F_COORDINATES, NALU_F_VELOCITY, NALU_F_VISCOSITY}); mdot = true; break;
  case NALU_K_CONTINUITY_ADV: add({NALU_F_COORDINATES, NALU_F_VELOCITY, NALU_F_DENSITY, NALU_F_PRESSURE, NALU_F_DPDX}); break;
  case NALU_K_SCALAR_ADV_DIFF: add({NALU_F_COORDINATES, NALU_F_SCALARQ, NALU_F_VISCOSITY}); mdot = true; break;
  case NALU_K_SCALAR_DIFF: add({NALU_F_COORDINATES, NALU_F_SCALARQ, NALU_F_VISCOSITY}); break;
  case NALU_K_SCALAR_UPW_ADV_DIFF:
    add({NALU_F_COORDINATES, NALU_F_VELOCITY, NALU_F_GJQ, NALU_F_SCALARQ, NALU_F_DENSITY, NALU_F_VISCOSITY}); mdot = true; break;
  case NALU_K_MOMENTUM_UPW_ADV_DIFF:
    add({NALU_F_COORDINATES, NALU_F_VELOCITY, NALU_F_VISCOSITY, NALU_F_DENSITY, NALU_F_GJU}); mdot = true; break;
  case NALU_K_MOMENTUM_MASS:
    add({NALU_F_COORDINATES, NALU_F_VELOCITY, NALU_F_VELOCITY_N, NALU_F_VELOCITY_NM1, NALU_F_DPDX, NALU_F_DENSITY, NALU_F_DENSITY_N,
         NALU_F_DENSITY_NM1}); break;
  case NALU_K_SCALAR_MASS:
    add({NALU_F_COORDINATES, NALU_F_SCALARQ, NALU_F_SCALARQ_N, NALU_F_SCALARQ_NM1, NALU_F_DENSITY, NALU_F_DENSITY_N, NALU_F_DENSITY_NM1}); break;
  case NALU_K_CONTINUITY_MASS: add({NALU_F_COORDINATES, NALU_F_DENSITY, NALU_F_DENSITY_N, NALU_F_DENSITY_NM1}); break;
  case NALU_K_MOMENTUM_BUOYANCY_SRC: add({NALU_F_COORDINATES, NALU_F_DENSITY}); break;
  case NALU_K_MOMENTUM_BUOYANCY_BOUSSINESQ_SRC: add({NALU_F_COORDINATES, NALU_F_TEMPERATURE}); break;
  case NALU_K_TKE_SST_SRC:
  case NALU_K_SDR_SST_SRC:
    add({NALU_F_COORDINATES, NALU_F_VELOCITY, NALU_F_DENSITY, NALU_F_TKE, NALU_F_SDR, NALU_F_TVISC});
    if (k.kind == NALU_K_SDR_SST_SRC) add({NALU_F_FONE_BLEND});
    break;
  case NALU_K_TKE_SST_DES_SRC:
    add({NALU_F_COORDINATES, NALU_F_VELOCITY, NALU_F_DENSITY, NALU_F_TKE, NALU_F_SDR, NALU_F_TVISC, NALU_F_FONE_BLEND,
         NALU_F_SST_MAX_LENGTH_SCALE});
    break;
  case NALU_K_MOMENTUM_NSO:
    add({NALU_F_COORDINATES, NALU_F_GJU, NALU_F_VELOCITY, NALU_F_VELOCITY_N, NALU_F_VELOCITY_NM1, NALU_F_DENSITY, NALU_F_DENSITY_N,
         NALU_F_DENSITY_NM1, NALU_F_VISCOSITY, NALU_F_PRESSURE});
    break;
  case NALU_K_TKE_KSGS_SRC:
    add({NALU_F_COORDINATES, NALU_F_TKE, NALU_F_DENSITY, NALU_F_TVISC, NALU_F_DUAL_VOLUME, NALU_F_C_EPSILON, NALU_F_GJU,
         NALU_F_VISCOSITY, NALU_F_DSQRTK_DX_SQ});
    break;
  case NALU_K_TEST_FIXED_BLOCK: return k.fixed_block ? nullptr : "fixed_block";
  default: return "unknown kernel kind";
  }
  const bool upw = k.kind == NALU_K_SCALAR_UPW_ADV_DIFF || k.kind == NALU_K_MOMENTUM_UPW_ADV_DIFF;
  for (int q = 0; q < n; ++q) {
    if (need[q] == NALU_F_VELOCITY && upw && k.kind == NALU_K_SCALAR_UPW_ADV_DIFF && k.nodal[NALU_F_VELOCITY_RTM]) continue;
    if (!k.nodal[need[q]]) return names[need[q]];
  }
  if (mdot && !k.mass_flow_rate) return "mass_flow_rate_scs";
  return nullptr;
}

int launch_assemble_general(const Graph &g, const nalu_b200_kernel_desc *kernels, int nk, double *d_values,
                            double *d_rhs, cudaStream_t stream)
{
  if (nk < 1 || nk > kMaxKernels) {
    set_error("assemble_elem: between 1 and 8 kernels per algorithm");
    return NALU_B200_ERR_ARG;
  }
  GeneralParams p;
  p.n_elems = g.n_elems;
  p.elem_nodes = g.d_elem_nodes;
  p.node_row = g.d_node_row;
  p.browptr = (const long long *)g.d_browptr;
  p.elem_slot = g.d_elem_slot;
  p.values = d_values;
  p.rhs = d_rhs;
  p.err = g.d_err;
  p.nk = nk;
  bool need_coords = false;
  for (int i = 0; i < nk; ++i) {
    p.k[i] = kernels[i];
    const int kind = kernels[i].kind;
    if (kind < 0 || kind >= NALU_K_COUNT) {
      set_error("assemble_elem: unknown kernel kind");
      return NALU_B200_ERR_ARG;
    }
    const bool vec = kind == NALU_K_MOMENTUM_ADV_DIFF || kind == NALU_K_MOMENTUM_MASS ||
                     kind == NALU_K_MOMENTUM_BUOYANCY_SRC || kind == NALU_K_MOMENTUM_BUOYANCY_BOUSSINESQ_SRC ||
                     kind == NALU_K_MOMENTUM_UPW_ADV_DIFF || kind == NALU_K_MOMENTUM_NSO;
    if (kind != NALU_K_TEST_FIXED_BLOCK) {
      need_coords = true;
      if ((vec && g.ndof != 3) || (!vec && g.ndof != 1)) {
        set_error("assemble_elem: kernel kind does not match the linear system's numDof");
        return NALU_B200_ERR_ARG;
      }
      if (!kernels[i].nodal[NALU_F_COORDINATES]) {
        set_error("assemble_elem: coordinates field missing (add_coordinates_field must precede master element calls)");
        return NALU_B200_ERR_ARG;
      }
    }
    if (const char *missing = missing_field(kernels[i])) {
      set_error(std::string("assemble_elem: field missing for kernel ") + std::to_string(i) + ": " + missing);
      return NALU_B200_ERR_ARG;
    }
  }
  (void)need_coords;
  if (g.n_elems == 0) return NALU_B200_OK;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (g.ndof != 1 && g.ndof != 3) {
    set_error("assemble_elem: numDof must be 1 or 3");
    return NALU_B200_ERR_ARG;
  }
  const size_t smem = (g.ndof == 3 ? sizeof(ElemScratch<3>) : sizeof(ElemScratch<1>)) * kWarpsPerCta;
  const int minb = tuning().general_minb;
  void (*kern)(const GeneralParams) =
      g.ndof == 3 ? (minb == 2 ? assemble_general_kernel<3, 2> : minb == 3 ? assemble_general_kernel<3, 3> : assemble_general_kernel<3, 4>)
                  : (minb == 2 ? assemble_general_kernel<1, 2> : minb == 3 ? assemble_general_kernel<1, 3> : assemble_general_kernel<1, 4>);
  NALU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  NALU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  // one launch per (chunk, colour) group (same stream: the groups run in order), or a single atomic launch when there
  // is no colouring
  const long long ncolor = g.d_color_elems ? (long long)g.color_ptr.size() - 1 : 1;
  for (long long c = 0; c < ncolor; ++c) {
    p.elem_list = g.d_color_elems ? g.d_color_elems + g.color_ptr[c] : nullptr;
    p.red_scatter = tuning().general_scatter;
    p.n_elems = g.d_color_elems ? g.color_ptr[c + 1] - g.color_ptr[c] : g.n_elems;
    if (p.n_elems == 0) continue;
    long long ctas = (p.n_elems + kWarpsPerCta - 1) / kWarpsPerCta;
    if (ctas > (long long)sms * 16) ctas = (long long)sms * 16;
    kern<<<(unsigned)ctas, kWarpsPerCta * 32, smem, stream>>>(p);
    count_launch();
  }
  NALU_CUDA(cudaGetLastError());
  return NALU_B200_OK;
}

// ---- lumped mass kernels on their own: thread = (element, sub-control volume) -------------------------------------
// MomentumMassElemKernel / ContinuityMassElemKernel / ScalarMassElemKernel with the lumped (shifted) shape function
// (src/kernel/MomentumMassElemKernel.C:119-170, ContinuityMassElemKernel.C:98-117, ScalarMassElemKernel.C:117-154):
// the shape function is 1 at the SCV's own node and 0 elsewhere, so the element contributes the SCV volume times
// nodal quantities to the diagonal entry and the residual of that node -- 6 (momentum) or 2 adds per SCV instead of
// a 24 x 24 scratch block per element.  Used for the lumped mass kernels of a list whose adv-diff kernel went through
// a tiled kernel (capi.cu).  Colour by colour like the general kernel (elem_list = the elements of one colour: plain adds
// in a fixed order); without a colouring (elem_list null) the adds are atomic.
#define NALU_ADD(ptr, val)                 \
  do {                                     \
    if (elem_list) *(ptr) += (val);        \
    else atomicAdd((ptr), (val));          \
  } while (0)
__global__ void __launch_bounds__(256) scv_lumped_mass_kernel(long long n_elems, const int *elem_list, int ndof, const int *elem_nodes,
                                                              const int *node_row, const long long *browptr, const unsigned char *elem_slot,
                                                              nalu_b200_kernel_desc k, double *values, double *rhs)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int ip = (int)(t & 7);
  if ((t >> 3) >= n_elems) return;
  const long long e = elem_list ? elem_list[t >> 3] : (t >> 3);
  double c[8][3];
  const double *X = k.nodal[NALU_F_COORDINATES];
  for (int n = 0; n < 8; ++n) {
    const int nd = elem_nodes[e * 8 + n];
    c[n][0] = X[(size_t)nd * 3 + 0];
    c[n][1] = X[(size_t)nd * 3 + 1];
    c[n][2] = X[(size_t)nd * 3 + 2];
  }
  const double scV = hex8::scv_volume_ip(c, ip);
  const int node = elem_nodes[e * 8 + ip];
  const int rn = node_row[node];
  const long long b0 = browptr[rn];
  const int nb = (int)(browptr[rn + 1] - b0), slot = elem_slot[e * 64 + ip * 8 + ip];
  const double rNp1 = k.nodal[NALU_F_DENSITY][node], rN = k.nodal[NALU_F_DENSITY_N][node], rNm1 = k.nodal[NALU_F_DENSITY_NM1][node];
  if (k.kind == NALU_K_MOMENTUM_MASS) {
    const double densFac = rNp1 * k.dens_fac + k.om_dens_fac;
    const double lhsfac = k.gamma1 * rNp1 * scV / k.dt;
    for (int j = 0; j < 3; ++j) {
      const double uNp1 = k.nodal[NALU_F_VELOCITY][(size_t)node * 3 + j], uN = k.nodal[NALU_F_VELOCITY_N][(size_t)node * 3 + j],
                   uNm1 = k.nodal[NALU_F_VELOCITY_NM1][(size_t)node * 3 + j], Gp = k.nodal[NALU_F_DPDX][(size_t)node * 3 + j];
      NALU_ADD(rhs + (long long)rn * 3 + j,
               -(k.gamma1 * rNp1 * uNp1 + k.gamma2 * rN * uN + k.gamma3 * rNm1 * uNm1) * scV / k.dt - Gp * scV * densFac);
      NALU_ADD(values + 9 * b0 + (long long)j * nb * 3 + slot * 3 + j, lhsfac);
    }
  } else if (k.kind == NALU_K_CONTINUITY_MASS) {
    const double tau = k.dt / k.gamma1;
    NALU_ADD(rhs + rn, -(k.gamma1 * rNp1 + k.gamma2 * rN + k.gamma3 * rNm1) * scV / k.dt / tau);
  } else {
    const double qNp1 = k.nodal[NALU_F_SCALARQ][node], qN = k.nodal[NALU_F_SCALARQ_N][node], qNm1 = k.nodal[NALU_F_SCALARQ_NM1][node];
    NALU_ADD(rhs + rn, -(k.gamma1 * rNp1 * qNp1 + k.gamma2 * rN * qN + k.gamma3 * rNm1 * qNm1) * scV / k.dt);
    NALU_ADD(values + b0 + slot, k.gamma1 * rNp1 * scV / k.dt);
  }
  (void)ndof;
}
#undef NALU_ADD

bool scv_lumped_supported(const Graph &g, const nalu_b200_kernel_desc *k, int nk)
{
  if (!g.d_elem_slot || nk < 1) return false;
  for (int i = 0; i < nk; ++i) {
    const int kind = k[i].kind;
    const bool mass = kind == NALU_K_MOMENTUM_MASS || kind == NALU_K_CONTINUITY_MASS || kind == NALU_K_SCALAR_MASS;
    if (!mass || !(k[i].options & NALU_OPT_LUMPED_MASS) || !k[i].nodal[NALU_F_COORDINATES]) return false;
    if ((kind == NALU_K_MOMENTUM_MASS) != (g.ndof == 3)) return false;
  }
  return true;
}

int launch_scv_lumped(const Graph &g, const nalu_b200_kernel_desc *k, int nk, double *d_values, double *d_rhs, cudaStream_t stream)
{
  if (!scv_lumped_supported(g, k, nk)) {
    set_error("scv_lumped: unsupported kernel list");
    return NALU_B200_ERR_ARG;
  }
  for (int i = 0; i < nk; ++i)
    if (const char *missing = missing_field(k[i])) {
      set_error(std::string("assemble_elem: field missing for a lumped mass kernel: ") + missing);
      return NALU_B200_ERR_ARG;
    }
  if (g.n_elems == 0) return NALU_B200_OK;
  const long long ncolor = g.d_color_elems ? (long long)g.color_ptr.size() - 1 : 1;
  for (int i = 0; i < nk; ++i)
    for (long long c = 0; c < ncolor; ++c) {
      const int *list = g.d_color_elems ? g.d_color_elems + g.color_ptr[c] : nullptr;
      const long long n = g.d_color_elems ? g.color_ptr[c + 1] - g.color_ptr[c] : g.n_elems;
      if (n == 0) continue;
      scv_lumped_mass_kernel<<<(unsigned)((n * 8 + 255) / 256), 256, 0, stream>>>(n, list, g.ndof, g.d_elem_nodes, g.d_node_row,
                                                                                 (const long long *)g.d_browptr, g.d_elem_slot, k[i],
                                                                                 d_values, d_rhs);
      count_launch();
    }
  NALU_CUDA(cudaGetLastError());
  return NALU_B200_OK;
}

// ---- node mass algorithms ----------------------------------------------------------------
// AssembleNodeSolverAlgorithm::execute (src/AssembleNodeSolverAlgorithm.C:57-111) with
// {Momentum,Continuity,Scalar}MassBDF2NodeSuppAlg::node_execute; one thread per locally-owned ROW: the node visited is the
// row's representative (periodic subjects, which alias their monarch's row, are excluded from the reference's selector), so
// every row is touched by exactly one thread -- plain adds, no atomics.  The 1-node sumInto targets the diagonal block.
__global__ void node_mass_kernel(int n_owned_nodes, int ndof, const int *row_rep, const long long *browptr, const int *bcol,
                                 nalu_b200_kernel_desc k, double *values, double *rhs)
{
  const int rn = blockIdx.x * blockDim.x + threadIdx.x;
  if (rn >= n_owned_nodes) return;
  const int node = row_rep[rn];
  if (node < 0) return;
  const long long b0 = browptr[rn], b1 = browptr[rn + 1];
  // diagonal block: column rn (binary search, rows are sorted)
  long long lo = b0, hi = b1;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (bcol[mid] < rn) lo = mid + 1; else hi = mid;
  }
  if (lo >= b1 || bcol[lo] != rn) return;
  const int slot = (int)(lo - b0), nb = (int)(b1 - b0);
  const double dualVolume = k.nodal[NALU_F_DUAL_VOLUME][node];
  const double rhoNm1 = k.nodal[NALU_F_DENSITY_NM1][node], rhoN = k.nodal[NALU_F_DENSITY_N][node],
               rhoNp1 = k.nodal[NALU_F_DENSITY][node];
  const long long base = (long long)ndof * ndof * b0;
  if (k.kind == NALU_K_MOMENTUM_MASS) {
    const double lhsfac = k.gamma1 * rhoNp1 * dualVolume / k.dt;
    for (int i = 0; i < 3; ++i) {
      const double uNm1 = k.nodal[NALU_F_VELOCITY_NM1][(size_t)node * 3 + i], uN = k.nodal[NALU_F_VELOCITY_N][(size_t)node * 3 + i],
                   uNp1 = k.nodal[NALU_F_VELOCITY][(size_t)node * 3 + i], dpdx = k.nodal[NALU_F_DPDX][(size_t)node * 3 + i];
      const double r = -(k.gamma1 * rhoNp1 * uNp1 + k.gamma2 * rhoN * uN + k.gamma3 * rhoNm1 * uNm1) * dualVolume / k.dt -
                       dpdx * dualVolume;
      rhs[(long long)rn * 3 + i] += r;
      values[base + (long long)i * nb * 3 + slot * 3 + i] += lhsfac;
    }
  } else if (k.kind == NALU_K_CONTINUITY_MASS) {
    const double tau = k.dt / k.gamma1;
    rhs[rn] += -(k.gamma1 * rhoNp1 + k.gamma2 * rhoN + k.gamma3 * rhoNm1) * dualVolume / k.dt / tau;
  } else {
    const double qNm1 = k.nodal[NALU_F_SCALARQ_NM1][node], qN = k.nodal[NALU_F_SCALARQ_N][node], qNp1 = k.nodal[NALU_F_SCALARQ][node];
    rhs[rn] += -(k.gamma1 * rhoNp1 * qNp1 + k.gamma2 * qN * rhoN + k.gamma3 * qNm1 * rhoNm1) * dualVolume / k.dt;
    values[base + slot] += k.gamma1 * rhoNp1 * dualVolume / k.dt;
  }
}

int launch_node_mass(const Graph &g, const nalu_b200_kernel_desc &k, double *d_values, double *d_rhs, cudaStream_t stream)
{
  const bool vec = k.kind == NALU_K_MOMENTUM_MASS;
  if (!(k.kind == NALU_K_MOMENTUM_MASS || k.kind == NALU_K_CONTINUITY_MASS || k.kind == NALU_K_SCALAR_MASS) ||
      (vec ? g.ndof != 3 : g.ndof != 1) || !k.nodal[NALU_F_DUAL_VOLUME]) {
    set_error("assemble_node_mass: needs a mass kernel kind matching numDof and the dual_nodal_volume field");
    return NALU_B200_ERR_ARG;
  }
  if (const char *missing = missing_field(k)) {
    set_error(std::string("assemble_node_mass: nodal field missing: ") + missing);
    return NALU_B200_ERR_ARG;
  }
  if (g.n_owned_nodes == 0) return NALU_B200_OK;
  node_mass_kernel<<<(g.n_owned_nodes + 255) / 256, 256, 0, stream>>>(g.n_owned_nodes, g.ndof, g.d_row_rep,
                                                               (const long long *)g.d_browptr, g.d_bcol, k, d_values, d_rhs);
  count_launch();
  NALU_CUDA(cudaGetLastError());
  return NALU_B200_OK;
}

// ---- sumInto for host callers (src/TpetraLinearSystem.C:1457-1522) ---------------------------------
// One thread per lhs entry; the column search the reference does by walking the sorted row is a
// binary search over the row's block columns; absent columns are skipped like the reference does.
__global__ void sum_into_kernel(int n_entities, int ndof, const int *nodes, const double *rhs, const double *lhs,
                                const int *node_row, const long long *browptr, const int *bcol, double *values, double *grhs)
{
  const int n = n_entities * ndof;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * n) return;
  const int r = t / n, c = t - r * n;
  const int rn = node_row[nodes[r / ndof]], i = r % ndof;
  const int cn = node_row[nodes[c / ndof]], j = c % ndof;
  const long long b0 = browptr[rn], b1 = browptr[rn + 1];
  long long lo = b0, hi = b1;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (bcol[mid] < cn) lo = mid + 1; else hi = mid;
  }
  if (lo < b1 && bcol[lo] == cn) {
    const long long nb = b1 - b0;
    atomicAdd(values + (long long)ndof * ndof * b0 + (long long)i * nb * ndof + (lo - b0) * ndof + j, lhs[t]);
  }
  if (c == 0) atomicAdd(grhs + (long long)rn * ndof + i, rhs[r]);
}

int launch_sum_into(const Graph &g, int n_entities, const int *h_nodes, const double *h_rhs, const double *h_lhs,
                    double *d_values, double *d_rhs, cudaStream_t stream)
{
  if (n_entities <= 0) return NALU_B200_OK;
  for (int e = 0; e < n_entities; ++e)
    if (h_nodes[e] < 0 || h_nodes[e] >= g.n_nodes) {
      set_error("sum_into: node index out of range");
      return NALU_B200_ERR_ARG;
    }
  const int n = n_entities * g.ndof;
  const size_t bytes_nodes = (size_t(n_entities) * sizeof(int) + 7) & ~size_t(7);
  const size_t bytes = bytes_nodes + (size_t(n) + size_t(n) * n) * sizeof(double);
  char *d = nullptr;
  NALU_CUDA(cudaMallocAsync((void **)&d, bytes, stream));
  NALU_CUDA(cudaMemcpyAsync(d, h_nodes, size_t(n_entities) * sizeof(int), cudaMemcpyHostToDevice, stream));
  NALU_CUDA(cudaMemcpyAsync(d + bytes_nodes, h_rhs, size_t(n) * sizeof(double), cudaMemcpyHostToDevice, stream));
  NALU_CUDA(cudaMemcpyAsync(d + bytes_nodes + size_t(n) * sizeof(double), h_lhs, size_t(n) * n * sizeof(double),
                            cudaMemcpyHostToDevice, stream));
  sum_into_kernel<<<(n * n + 127) / 128, 128, 0, stream>>>(n_entities, g.ndof, (const int *)d, (const double *)(d + bytes_nodes),
                                                         (const double *)(d + bytes_nodes + size_t(n) * sizeof(double)), g.d_node_row,
                                                         (const long long *)g.d_browptr, g.d_bcol, d_values, d_rhs);
  count_launch();
  NALU_CUDA(cudaGetLastError());
  NALU_CUDA(cudaFreeAsync(d, stream));
  // the host buffers may be reused by the caller as soon as we return
  NALU_CUDA(cudaStreamSynchronize(stream));
  return NALU_B200_OK;
}

// ---- applyDirichletBCs / resetRows (src/TpetraLinearSystem.C:1579-1653,1706-1755) ---------------------
// One warp per (listed node, dof row d in [beginPos, endPos)): the whole CSR row is overwritten -- 1 on the
// diagonal of a locally-owned row, 0 elsewhere (a shared-not-owned row is cleared completely, its owner
// holds the 1) -- and the residual entry becomes bc - solution (owned), 0 (shared-not-owned), or 0 for
// resetRows.  solution / bc_values are nodal fields [nNodes][numDof] indexed by the local node id.
__global__ void row_fixup_kernel(int n_list, int begin_pos, int end_pos, int ndof, int n_owned_nodes, const int *list,
                                 const int *node_row, const long long *browptr, const int *bcol, const double *solution,
                                 const double *bc_values, int dirichlet, double *values, double *rhs)
{
  const int nd = end_pos - begin_pos;
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= (long long)n_list * nd) return;
  const int node = list[w / nd], d = begin_pos + (int)(w % nd);
  const int rn = node_row[node];
  const bool owned = rn < n_owned_nodes;
  const long long b0 = browptr[rn], b1 = browptr[rn + 1];
  const int nb = (int)(b1 - b0);
  int diag = -1;
  if (dirichlet && owned) { // column of the row's own node: the block graph row is sorted
    long long lo = b0, hi = b1;
    while (lo < hi) {
      const long long mid = (lo + hi) >> 1;
      if (bcol[mid] < rn) lo = mid + 1; else hi = mid;
    }
    if (lo < b1 && bcol[lo] == rn) diag = (int)(lo - b0) * ndof + d;
  }
  double *row = values + (long long)ndof * ndof * b0 + (long long)d * nb * ndof;
  for (int e = lane; e < nb * ndof; e += 32) row[e] = e == diag ? 1.0 : 0.0;
  if (lane == 0)
    rhs[(long long)rn * ndof + d] =
        (dirichlet && owned) ? bc_values[(long long)node * ndof + d] - solution[(long long)node * ndof + d] : 0.0;
}

int launch_row_fixup(const Graph &g, int n_list, const int *d_list, int begin_pos, int end_pos, const double *d_solution,
                     const double *d_bc_values, int dirichlet, double *d_values, double *d_rhs, cudaStream_t stream)
{
  if (begin_pos < 0 || end_pos > g.ndof || begin_pos > end_pos) {
    set_error("applyDirichletBCs / resetRows: [beginPos, endPos) outside the system's dofs");
    return NALU_B200_ERR_ARG;
  }
  const long long warps = (long long)n_list * (end_pos - begin_pos);
  if (warps <= 0) return NALU_B200_OK;
  row_fixup_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, stream>>>(n_list, begin_pos, end_pos, g.ndof, g.n_owned_nodes, d_list,
                                                                           g.d_node_row, (const long long *)g.d_browptr, g.d_bcol,
                                                                           d_solution, d_bc_values, dirichlet, d_values, d_rhs);
  count_launch();
  NALU_CUDA(cudaGetLastError());
  return NALU_B200_OK;
}

// ---- mass_flow_rate_scs producer (unit_tests/kernels/UnitTestKernelUtils.C:547-628) --------
__global__ void mdot_kernel(long long n_elems, const int *elem_nodes, const double *coords, const double *density,
                            const double *velocity, double *mdot)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long e = t / 12;
  const int ip = (int)(t - e * 12);
  if (e >= n_elems) return;
  double c[8][3], ru[8][3];
  // coordinates relative to the element's first node: the area vector is translation invariant, its rounding error in
  // absolute coordinates (|x|/h ulps) is not
  const size_t n0 = (size_t)elem_nodes[e * 8] * 3;
  for (int n = 0; n < 8; ++n) {
    const int node = elem_nodes[e * 8 + n];
    const double rho = density[node];
    for (int d = 0; d < 3; ++d) {
      c[n][d] = coords[(size_t)node * 3 + d] - coords[n0 + d];
      ru[n][d] = rho * velocity[(size_t)node * 3 + d];
    }
  }
  double a[3], s1, s2, s3;
  hex8::scs_areav_ip(c, ip, a);
  hex8::scs_intg_loc(ip, 0.25, s1, s2, s3);
  double rhoU[3] = {0, 0, 0};
  for (int n = 0; n < 8; ++n) {
    const double r = hex8::shape_fcn(n, s1, s2, s3);
    for (int d = 0; d < 3; ++d) rhoU[d] += r * ru[n][d];
  }
  mdot[e * 12 + ip] = rhoU[0] * a[0] + rhoU[1] * a[1] + rhoU[2] * a[2];
}

// AssembleNodalGradElemAlgorithm / AssembleNodalGradUElemAlgorithm (src/AssembleNodalGradElemAlgorithm.C:77-133,
// src/AssembleNodalGradUElemAlgorithm.C:120-195): interior Green-Gauss contributions to the projected nodal gradient
// (dpdx, dqdx, dudx: inputs of the continuity, upwind and mass kernels).  One thread per (element, point).
template <int NCOMP>
__global__ void nodal_grad_kernel(long long n_elems, const int *elem_nodes, const double *coords, const double *field,
                                  const double *dual_volume, int use_shifted, double *grad)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long e = t / 12;
  const int ip = (int)(t - e * 12);
  if (e >= n_elems) return;
  double c[8][3];
  int nd[8];
  for (int n = 0; n < 8; ++n) {
    nd[n] = elem_nodes[e * 8 + n];
    for (int d = 0; d < 3; ++d) c[n][d] = coords[(size_t)nd[n] * 3 + d];
  }
  double a[3], s1, s2, s3, q[NCOMP];
  hex8::scs_areav_ip(c, ip, a);
  hex8::scs_intg_loc(ip, use_shifted ? 0.5 : 0.25, s1, s2, s3);
  for (int i = 0; i < NCOMP; ++i) q[i] = 0.0;
  for (int n = 0; n < 8; ++n) {
    const double r = hex8::shape_fcn(n, s1, s2, s3);
    for (int i = 0; i < NCOMP; ++i) q[i] += r * field[(size_t)nd[n] * NCOMP + i];
  }
  const int nl = nd[hex8::scs_left(ip)], nr = nd[hex8::scs_right(ip)];
  const double ivl = 1.0 / dual_volume[nl], ivr = 1.0 / dual_volume[nr];
  for (int i = 0; i < NCOMP; ++i)
    for (int j = 0; j < 3; ++j) {
      const double fac = q[i] * a[j];
      atomicAdd(grad + (size_t)nl * NCOMP * 3 + i * 3 + j, fac * ivl);
      atomicAdd(grad + (size_t)nr * NCOMP * 3 + i * 3 + j, -fac * ivr);
    }
}

int launch_nodal_grad(const Graph &g, const double *coords, const double *field, int ncomp, const double *dual_volume, int use_shifted,
                      double *grad, cudaStream_t stream)
{
  if (ncomp != 1 && ncomp != 3) {
    set_error("assemble_nodal_grad_elem: ncomp must be 1 (scalar) or 3 (vector)");
    return NALU_B200_ERR_ARG;
  }
  if (g.n_elems == 0) return NALU_B200_OK;
  const long long nt = g.n_elems * 12;
  const unsigned nb = (unsigned)((nt + 127) / 128);
  if (ncomp == 1) nodal_grad_kernel<1><<<nb, 128, 0, stream>>>(g.n_elems, g.d_elem_nodes, coords, field, dual_volume, use_shifted, grad);
  else nodal_grad_kernel<3><<<nb, 128, 0, stream>>>(g.n_elems, g.d_elem_nodes, coords, field, dual_volume, use_shifted, grad);
  count_launch();
  NALU_CUDA(cudaGetLastError());
  return NALU_B200_OK;
}

// ComputeMdotElemAlgorithm::execute (src/ComputeMdotElemAlgorithm.C:85-254): the producer of mass_flow_rate_scs in a
// time step, mdot_ip = (rho u_ip - projTimeScale (dpdx_ip - Gpdx_ip)) . areav_ip.  One thread per element, all 12 points:
// the trilinear fields (coordinates relative to the first node, w = rho u + tau Gp, p) are expanded once in their
// Walsh-Hadamard coefficients (box_tile.cuh butterfly8); at a point on an edge along t (local axes t,u,v) the Jacobian
// is 5 vector FMAs, the area vector one cross product (as in box5_tile.cuh p1_point), and
//   dpdx . A = (adj^T A) . dp/dxi / det,   w_ip = W_0/8 + q_m (s_u W_u + s_v W_v)/4 + q_m^2 s_u s_v W_uv/2.
// Elements [e0, e1): a caller that streams the nodal fields in (bench.py end-to-end step) produces the field range by range.
template <int T>
__device__ __forceinline__ double mdot_point(const double (&X)[3][8], const double (&W)[3][8], const double (&Pc)[8], double sgu,
                                             double sgv, double qp, double qm, double tau, int *err)
{
  constexpr int mt = 1 << T, mu = 1 << ((T + 1) % 3), mv = 1 << ((T + 2) % 3);
  const double c1 = 0.5 * qp * sgu, c2 = 0.5 * qp * sgv, c12 = qp * qp * sgu * sgv;
  double J[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    J[i][0] = 0.25 * X[i][mt] + c1 * X[i][mt | mu] + c2 * X[i][mt | mv] + c12 * X[i][7];
    J[i][1] = 0.25 * X[i][mu] + c2 * X[i][mu | mv];
    J[i][2] = 0.25 * X[i][mv] + c1 * X[i][mu | mv];
  }
  double adj[3][3];
  adj[0][0] = J[1][1] * J[2][2] - J[2][1] * J[1][2];
  adj[0][1] = J[1][2] * J[2][0] - J[2][2] * J[1][0];
  adj[0][2] = J[1][0] * J[2][1] - J[2][0] * J[1][1];
  adj[1][0] = J[0][2] * J[2][1] - J[2][2] * J[0][1];
  adj[1][1] = J[0][0] * J[2][2] - J[2][0] * J[0][2];
  adj[1][2] = J[0][1] * J[2][0] - J[2][1] * J[0][0];
  adj[2][0] = J[0][1] * J[1][2] - J[1][1] * J[0][2];
  adj[2][1] = J[0][2] * J[1][0] - J[1][2] * J[0][0];
  adj[2][2] = J[0][0] * J[1][1] - J[1][0] * J[0][1];
  const double det = J[0][0] * adj[0][0] + J[1][0] * adj[1][0] + J[2][0] * adj[2][0];
  if (!(det > 0.0)) *err = 1;
  // area vector along +t (geometry of the sub-control surface: never shifted)
  double A[3], Wv[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) Wv[i] = sgv * X[i][mv] - sgu * X[i][mu];
  A[0] = (X[1][mu] * X[2][mv] - X[2][mu] * X[1][mv]) * (1.0 / 64.0) + (X[1][mu | mv] * Wv[2] - X[2][mu | mv] * Wv[1]) * (1.0 / 128.0);
  A[1] = (X[2][mu] * X[0][mv] - X[0][mu] * X[2][mv]) * (1.0 / 64.0) + (X[2][mu | mv] * Wv[0] - X[0][mu | mv] * Wv[2]) * (1.0 / 128.0);
  A[2] = (X[0][mu] * X[1][mv] - X[1][mu] * X[0][mv]) * (1.0 / 64.0) + (X[0][mu | mv] * Wv[1] - X[1][mu | mv] * Wv[0]) * (1.0 / 128.0);
  // dp/dxi at the grad-op point, local axis order
  const double pt = 0.25 * Pc[mt] + c1 * Pc[mt | mu] + c2 * Pc[mt | mv] + c12 * Pc[7];
  const double pu = 0.25 * Pc[mu] + c2 * Pc[mu | mv], pv = 0.25 * Pc[mv] + c1 * Pc[mu | mv];
  double gA = 0.0, wA = 0.0;
  const double m1 = 0.25 * qm * sgu, m2 = 0.25 * qm * sgv, m12 = 0.5 * qm * qm * sgu * sgv;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    gA += (adj[j][0] * pt + adj[j][1] * pu + adj[j][2] * pv) * A[j];
    wA += (0.125 * W[j][0] + m1 * W[j][mu] + m2 * W[j][mv] + m12 * W[j][mu | mv]) * A[j];
  }
  return wA - tau * gA / det;
}

__global__ void compute_mdot_kernel(long long e0, long long e1, const int *elem_nodes, const double *coords, const double *density,
                                    const double *velocity, const double *pressure, const double *gpdx, double tau,
                                    int shift_mdot, int shift_poisson, double *mdot, int *err)
{
  const long long e = e0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= e1) return;
  double X[3][8], W[3][8], Pc[8];
  // Hex8 node n -> lattice bits (x, y, z): box_tile.cuh bit_x / bit_y / bit_z; coefficient index m = x + 2 y + 4 z
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const int node = elem_nodes[e * 8 + n];
    const int m = box::bit_x(n) + 2 * box::bit_y(n) + 4 * box::bit_z(n);
    const double rho = density[node];
    Pc[m] = pressure[node];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      X[d][m] = coords[(size_t)node * 3 + d];
      W[d][m] = rho * velocity[(size_t)node * 3 + d] + tau * gpdx[(size_t)node * 3 + d]; // interpolated with one shape function
    }
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double o = X[d][0];
#pragma unroll
    for (int m = 0; m < 8; ++m) X[d][m] -= o;
    box::butterfly8(X[d]);
    box::butterfly8(W[d]);
  }
  box::butterfly8(Pc);
  const double qp = shift_poisson ? 0.5 : 0.25, qm = shift_mdot ? 0.5 : 0.25;
  double *o = mdot + e * 12;
  // ip numbering and L -> R orientation: Hex8CVFEM.C:264-276 (ip 2 and 6 run along -x)
#pragma unroll 1
  for (int f = 0; f < 2; ++f) { // bottom / top face
    const double sz = f ? 1.0 : -1.0;
    o[4 * f + 0] = mdot_point<0>(X, W, Pc, -1.0, sz, qp, qm, tau, err);
    o[4 * f + 1] = mdot_point<1>(X, W, Pc, sz, 1.0, qp, qm, tau, err);
    o[4 * f + 2] = -mdot_point<0>(X, W, Pc, 1.0, sz, qp, qm, tau, err);
    o[4 * f + 3] = mdot_point<1>(X, W, Pc, sz, -1.0, qp, qm, tau, err);
  }
#pragma unroll 1
  for (int n = 0; n < 4; ++n)
    o[8 + n] = mdot_point<2>(X, W, Pc, (n == 1 || n == 2) ? 1.0 : -1.0, n >= 2 ? 1.0 : -1.0, qp, qm, tau, err);
}

int launch_compute_mdot(const Graph &g, const double *coords, const double *density, const double *velocity, const double *pressure,
                        const double *gpdx, double tau, int options, double *mdot, cudaStream_t stream, long long e0, long long e1)
{
  if (e1 < 0) e1 = g.n_elems;
  if (e0 < 0 || e0 > e1 || e1 > g.n_elems) {
    set_error("compute_mdot_elem: bad element range");
    return NALU_B200_ERR_ARG;
  }
  if (e1 == e0) return NALU_B200_OK;
  compute_mdot_kernel<<<(unsigned)((e1 - e0 + 127) / 128), 128, 0, stream>>>(e0, e1, g.d_elem_nodes, coords, density, velocity, pressure,
                                                                           gpdx, tau, (options & NALU_OPT_SHIFT_MDOT) != 0,
                                                                           (options & NALU_OPT_SHIFT_POISSON) != 0, mdot, g.d_err);
  count_launch();
  NALU_CUDA(cudaGetLastError());
  return NALU_B200_OK;
}

int launch_mdot(const Graph &g, const double *coords, const double *density, const double *velocity, double *mdot,
                cudaStream_t stream)
{
  if (g.n_elems == 0) return NALU_B200_OK;
  const long long nt = g.n_elems * 12;
  mdot_kernel<<<(unsigned)((nt + 127) / 128), 128, 0, stream>>>(g.n_elems, g.d_elem_nodes, coords, density, velocity, mdot);
  count_launch();
  NALU_CUDA(cudaGetLastError());
  return NALU_B200_OK;
}

} // namespace nalu_b200
