/*
 * nalu_b200.h -- C ABI of the B200-native Hex8 CVFEM element-assembly path.
 *
 * One shared library (nalu_b200/libnalu_b200.so) exports exactly these entry points.
 * Plain pointers and sizes only; no C++/torch types; every function returns an int status
 * (NALU_B200_OK == 0) and never throws across the boundary.  Device pointers are plain
 * CUDA device addresses (cudaMalloc, torch.Tensor.data_ptr(), ...); the caller owns every
 * buffer.  `stream` is a cudaStream_t passed as void* (NULL = default stream); calls are
 * asynchronous on that stream unless stated otherwise.
 *
 * What each entry point replaces in the reference (paths relative to the Nalu tree):
 *
 *   nalu_b200_graph_*         TpetraLinearSystem::beginLinearSystemConstruction /
 *                             buildElemToNodeGraph / finalizeLinearSystem
 *                             (src/TpetraLinearSystem.C:190-327, 392-420, 1257-1353) and
 *                             LocalGraphArrays (include/LinearSolver.h:54-120)
 *   nalu_b200_zero_system     TpetraLinearSystem::zeroSystem (src/TpetraLinearSystem.C:1355-1372)
 *   nalu_b200_assemble_elem   AssembleElemSolverAlgorithm::execute: gather, master-element
 *                             views, every registered Kernel::execute, apply_coeff ->
 *                             TpetraLinearSystem::sumInto
 *                             (src/AssembleElemSolverAlgorithm.C:74-100,
 *                              include/AssembleElemSolverAlgorithm.h:52-105,
 *                              src/ScratchViews.C:408-509, src/kernel/ *ElemKernel.C,
 *                              src/SolverAlgorithm.C:57-68, src/TpetraLinearSystem.C:1376-1522)
 *   nalu_b200_sum_into        TpetraLinearSystem::sumInto for host callers
 *                             (src/TpetraLinearSystem.C:1457-1522)
 *   nalu_b200_assemble_node_mass  AssembleNodeSolverAlgorithm::execute with the
 *                             {Momentum,Continuity,Scalar}MassBDF2NodeSuppAlg
 *                             (src/AssembleNodeSolverAlgorithm.C:57-111,
 *                              src/MomentumMassBDF2NodeSuppAlg.C:80-103)
 *   nalu_b200_halo_*          TpetraLinearSystem::loadComplete: Export(ADD) of the
 *                             shared-not-owned rows into the owners' rows
 *                             (src/TpetraLinearSystem.C:1757-1778)
 *   nalu_b200_mdot_scs        the mass_flow_rate_scs producer used by the reference's tests
 *                             (unit_tests/kernels/UnitTestKernelUtils.C:547-628)
 */
#ifndef NALU_B200_H
#define NALU_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  NALU_B200_OK = 0,
  NALU_B200_ERR_ARG = 1,       /* bad argument / unsupported combination            */
  NALU_B200_ERR_CUDA = 2,      /* a CUDA call failed (see nalu_b200_last_error())   */
  NALU_B200_ERR_NO_DEVICE = 3, /* no CUDA device: this path has no CPU fallback     */
  NALU_B200_ERR_STATE = 4,     /* call sequence violated (e.g. finalize before begin)*/
  NALU_B200_ERR_GEOMETRY = 5   /* non-positive Jacobian determinant detected        */
};

const char *nalu_b200_last_error(void);
int nalu_b200_version(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
int64_t nalu_b200_launch_count(void);

/* ------------------------------------------------------------------------------------ */
/* Kernel descriptors: what the reference's Kernel constructors + setup() capture.       */
/* ------------------------------------------------------------------------------------ */
enum nalu_b200_kernel_kind {
  NALU_K_MOMENTUM_ADV_DIFF = 0, /* src/kernel/MomentumAdvDiffElemKernel.C:70-186 */
  NALU_K_CONTINUITY_ADV = 1,    /* src/kernel/ContinuityAdvElemKernel.C:86-148   */
  NALU_K_SCALAR_ADV_DIFF = 2,   /* src/kernel/ScalarAdvDiffElemKernel.C:69-127   */
  NALU_K_MOMENTUM_MASS = 3,     /* src/kernel/MomentumMassElemKernel.C:98-171    */
  NALU_K_CONTINUITY_MASS = 4,   /* src/kernel/ContinuityMassElemKernel.C:82-119  */
  NALU_K_SCALAR_MASS = 5,       /* src/kernel/ScalarMassElemKernel.C:96-156      */
  NALU_K_SCALAR_DIFF = 6,       /* src/kernel/ScalarDiffElemKernel.C:64-110      */
  NALU_K_TEST_FIXED_BLOCK = 7,  /* unit_tests/UnitTestTpetra.C:116-144 TestKernel */
  NALU_K_MOMENTUM_BUOYANCY_SRC = 8,            /* src/kernel/MomentumBuoyancySrcElemKernel.C:61-90           */
  NALU_K_MOMENTUM_BUOYANCY_BOUSSINESQ_SRC = 9, /* src/kernel/MomentumBuoyancyBoussinesqSrcElemKernel.C:66-94 */
  NALU_K_TKE_SST_SRC = 10, /* src/kernel/TurbKineticEnergySSTSrcElemKernel.C:95-174       */
  NALU_K_SDR_SST_SRC = 11, /* src/kernel/SpecificDissipationRateSSTSrcElemKernel.C:100-204 */
  NALU_K_SCALAR_UPW_ADV_DIFF = 12,   /* src/kernel/ScalarUpwAdvDiffElemKernel.C:110-256   */
  NALU_K_MOMENTUM_UPW_ADV_DIFF = 13, /* src/kernel/MomentumUpwAdvDiffElemKernel.C:108-311 */
  NALU_K_TKE_SST_DES_SRC = 14, /* src/kernel/TurbKineticEnergySSTDESSrcElemKernel.C:106-198 */
  NALU_K_TKE_KSGS_SRC = 15,    /* src/kernel/TurbKineticEnergyKsgsSrcElemKernel.C:78-128    */
  NALU_K_MOMENTUM_NSO = 16,    /* src/nso/MomentumNSOElemKernel.C:127-310                   */
  NALU_K_COUNT = 17
};

/* option bits (SolutionOptions fields the kernels read) */
enum {
  NALU_OPT_SHIFTED_GRAD_OP = 1 << 0, /* get_shifted_grad_op(dofName)                       */
  NALU_OPT_SKEW_SYMMETRIC = 1 << 1,  /* get_skew_symmetric(dofName): shifted adv shape fn   */
  NALU_OPT_LUMPED_MASS = 1 << 2,     /* mass kernels: shifted SCV shape fn                  */
  NALU_OPT_SHIFT_MDOT = 1 << 3,      /* cvfemShiftMdot_                                     */
  NALU_OPT_SHIFT_POISSON = 1 << 4,   /* get_shifted_grad_op("pressure")                     */
  NALU_OPT_REDUCED_SENS = 1 << 5,    /* cvfemReducedSensPoisson_                            */
  NALU_OPT_USE_LIMITER = 1 << 6,     /* primitive_uses_limiter(dofName): van Leer limiter   */
  NALU_OPT_PECLET_TANH = 1 << 7      /* get_tanh_functional_form(dofName) != "classic"      */
};

/* nodal-field slots of a descriptor; every field is [nNodes][ncomp] doubles, ncomp
 * contiguous per node (the layout of an STK nodal field), device memory */
enum nalu_b200_field_slot {
  NALU_F_COORDINATES = 0, /* 3 */
  NALU_F_VELOCITY = 1,    /* 3  velocity NP1 (velocity_rtm for continuity with mesh motion) */
  NALU_F_VELOCITY_N = 2,  /* 3 */
  NALU_F_VELOCITY_NM1 = 3,/* 3 */
  NALU_F_VISCOSITY = 4,   /* 1  viscosity / diffFluxCoeff */
  NALU_F_DENSITY = 5,     /* 1  density NP1 */
  NALU_F_DENSITY_N = 6,   /* 1 */
  NALU_F_DENSITY_NM1 = 7, /* 1 */
  NALU_F_PRESSURE = 8,    /* 1 */
  NALU_F_DPDX = 9,        /* 3 */
  NALU_F_SCALARQ = 10,    /* 1  scalarQ NP1 */
  NALU_F_SCALARQ_N = 11,  /* 1 */
  NALU_F_SCALARQ_NM1 = 12,/* 1 */
  NALU_F_DUAL_VOLUME = 13,/* 1  dual_nodal_volume (node mass algorithms only) */
  NALU_F_TEMPERATURE = 14,/* 1  temperature NP1 (Boussinesq buoyancy) */
  NALU_F_TKE = 15,        /* 1  turbulent_ke NP1 */
  NALU_F_SDR = 16,        /* 1  specific_dissipation_rate NP1 */
  NALU_F_TVISC = 17,      /* 1  turbulent_viscosity */
  NALU_F_FONE_BLEND = 18, /* 1  sst_f_one_blending */
  NALU_F_GJQ = 19,        /* 3  projected nodal gradient of scalarQ (upwind kernels) */
  NALU_F_GJU = 20,        /* 9  dudx: Gju(node,i,j) at [3*i+j] (upwind kernels) */
  NALU_F_VELOCITY_RTM = 21, /* 3  velocity_rtm of the upwind kernels when the mesh moves; NULL: NALU_F_VELOCITY */
  NALU_F_SST_MAX_LENGTH_SCALE = 22, /* 1  sst_max_length_scale (SST-DES source) */
  NALU_F_C_EPSILON = 23,    /* 1  c_epsilon (Ksgs source) */
  NALU_F_DSQRTK_DX_SQ = 24, /* 1  dsqrtk_dx_sq (Ksgs source; the reference binds the viscosity field when lrksgsfac = 0) */
  NALU_F_COUNT = 26
};

typedef struct nalu_b200_kernel_desc {
  int32_t kind;    /* nalu_b200_kernel_kind */
  int32_t options; /* NALU_OPT_* bits */
  double include_divu;          /* SolutionOptions::includeDivU_                     */
  double dens_fac, om_dens_fac; /* balancedForce_ ? (1,0) : (0,1)                    */
  double dt, gamma1, gamma2, gamma3; /* Kernel::setup(TimeIntegrator)                */
  const double *nodal[NALU_F_COUNT]; /* device pointers, NULL when unused            */
  const double *mass_flow_rate;      /* device [nElem][12], mass_flow_rate_scs       */
  const double *fixed_block;         /* device [rhsSize*rhsSize], NALU_K_TEST_FIXED_BLOCK */
  /* buoyancy source kernels: SolutionOptions::gravity_, referenceDensity_, referenceTemperature_,
   * thermalExpansionCoeff_ (src/kernel/MomentumBuoyancySrcElemKernel.C:31-41) */
  double gravity[3];
  double rho_ref, t_ref, beta;
  /* SST source kernels: SolutionOptions::get_turb_model_constant (src/SolutionOptions.C:742-770) */
  double turb[10]; /* NALU_TM_* */
  /* upwind adv-diff kernels: SolutionOptions::get_alpha_factor / get_alpha_upw_factor / get_upw_factor and the
   * Peclet function EquationSystem::create_peclet_function builds (include/EquationSystem.h:306-325):
   * classic A = 5 with Realm::get_hybrid_factor, or tanh with get_tanh_trans / get_tanh_width */
  double upw[8]; /* NALU_UPW_* */
  /* MomentumNSOElemKernel constructor arguments fourthFac / altResFac (src/nso/MomentumNSOElemKernel.C:28-36): the
   * kernel reads NALU_F_GJU, the three velocity and density states, NALU_F_VELOCITY_RTM (NULL: NALU_F_VELOCITY),
   * NALU_F_VISCOSITY, NALU_F_PRESSURE, dt / gamma1..3, include_divu and NALU_OPT_SHIFTED_GRAD_OP */
  double nso_fourth_fac, nso_alt_res_fac;
} nalu_b200_kernel_desc;
enum { NALU_UPW_ALPHA = 0, NALU_UPW_ALPHA_UPW = 1, NALU_UPW_HO_UPWIND = 2, NALU_UPW_HYBRID_FACTOR = 3,
       NALU_UPW_TANH_TRANS = 4, NALU_UPW_TANH_WIDTH = 5 };
enum { NALU_TM_BETA_STAR = 0, NALU_TM_TKE_PROD_LIMIT_RATIO = 1, NALU_TM_SIGMA_W_TWO = 2, NALU_TM_BETA_ONE = 3,
       NALU_TM_BETA_TWO = 4, NALU_TM_GAMMA_ONE = 5, NALU_TM_GAMMA_TWO = 6, NALU_TM_CDES_KW = 7, NALU_TM_CDES_KE = 8,
       NALU_TM_LRKSGS_FAC = 9 /* TurbKineticEnergyKsgsSrcElemKernel.C:54-57: 1 with the low-Reynolds correction, else 0 */ };

/* ------------------------------------------------------------------------------------ */
/* Graph: DOF numbering + CSR graph, built once per mesh.                                */
/* ------------------------------------------------------------------------------------ */
typedef struct nalu_b200_graph nalu_b200_graph; /* opaque, host + device resident */

/* Local mesh of this rank (host pointers, copied): nodes carry the 1-based naluGlobalId
 * (periodic subjects alias their monarch's id) and the owning rank; all elements are
 * locally owned.  box_n{x,y,z} > 0 declares that the nodes are the lexicographic
 * (x fastest) numbering of an nx*ny*nz-element structured block with Ioss `generated:`
 * connectivity, which enables the tiled write-once kernel; pass 0,0,0 otherwise. */
int nalu_b200_graph_create(nalu_b200_graph **out, int32_t rank, int32_t nranks, int32_t num_dof,
                           int32_t n_nodes, const int64_t *node_gid, const int32_t *node_owner,
                           int64_t n_elems, const int32_t *elem_nodes,
                           int32_t box_nx, int32_t box_ny, int32_t box_nz);
/* Same, for a named element topology: nodes_per_elem = 8 (Hex8, AlgTraitsHex8) or 27 (Hex27, AlgTraitsHex27:
 * include/AlgTraits.h:27-34; elem_nodes[n_elems][27] in STK's Hex27 node order, the one
 * HexahedralP2Element::stkNodeMap_ maps to the tensor grid, src/master_element/Hex27CVFEM.C:36-47).  Numbering, graph,
 * handshake, zeroSystem, loadComplete, Dirichlet / resetRows are topology independent and work as for Hex8.  On a Hex27
 * graph nalu_b200_assemble_elem takes [MomentumMassElemKernel(lumped)]? + MomentumAdvDiffElemKernel (3 dof, default options,
 * either order, the mass fused into the same launches; mass_flow_rate is [n_elems][216], the Hex27SCS point order of
 * Hex27CVFEM.C:887-1050) and nalu_b200_mdot_scs produces that field; on a 1-dof Hex27 graph it takes [ContinuityAdvElemKernel]
 * or [ScalarAdvDiffElemKernel] (default options); nalu_b200_compute_mdot_elem(_range) is ComputeMdotElemAlgorithm and
 * nalu_b200_assemble_nodal_grad_elem is AssembleNodalGrad(U)ElemAlgorithm on Hex27 (default options).  Other lists return NALU_B200_ERR_ARG and every other element-level entry point returns
 * NALU_B200_ERR_STATE. */
int nalu_b200_graph_create_topo(nalu_b200_graph **out, int32_t rank, int32_t nranks, int32_t num_dof,
                                int32_t nodes_per_elem, int32_t n_nodes, const int64_t *node_gid,
                                const int32_t *node_owner, int64_t n_elems, const int32_t *elem_nodes);
void nalu_b200_graph_destroy(nalu_b200_graph *g);

/* Neighbour handshake (TpetraLinearSystem.C:678-738,924-952).  After create(), for every
 * rank p != rank: query the byte size of the message for p, fetch it, exchange the messages
 * with whatever transport the host has (MPI, torch.distributed, ...), and hand the received
 * ones back.  Single-rank jobs skip straight to finalize. */
int nalu_b200_graph_send_size(const nalu_b200_graph *g, int32_t dest_rank, int64_t *nbytes);
int nalu_b200_graph_send_data(const nalu_b200_graph *g, int32_t dest_rank, void *buf);
int nalu_b200_graph_recv_data(nalu_b200_graph *g, int32_t src_rank, const void *buf, int64_t nbytes);
/* builds the owned and shared-not-owned CSR graphs, uploads them and the scatter maps */
int nalu_b200_graph_finalize(nalu_b200_graph *g);
/* same graph construction without touching a device: lets host-only tools (and the CPU test
 * suite) inspect the graph; such a graph cannot be assembled into */
int nalu_b200_graph_finalize_host_only(nalu_b200_graph *g);

typedef struct nalu_b200_graph_info {
  int32_t num_dof;
  int32_t n_owned_rows, n_shared_rows; /* dof rows: maxOwnedRowId_, numSharedNotOwned*numDof */
  int32_t n_cols;                      /* size of the column map (totalColsMap_)              */
  int64_t nnz_owned, nnz_shared;
  int64_t n_elems;
  int32_t n_nodes;
  int32_t is_box;
} nalu_b200_graph_info;
int nalu_b200_graph_get_info(const nalu_b200_graph *g, nalu_b200_graph_info *info);

/* Host copies for parity checks / hand-off (caller-allocated; `shared` selects the
 * shared-not-owned graph).  rowptr is size_t like the reference's (n_rows+1 entries,
 * starting at 0 for each graph), colidx is int32 local column ids. */
int nalu_b200_graph_copy_rowptr(const nalu_b200_graph *g, int shared, uint64_t *rowptr);
int nalu_b200_graph_copy_colidx(const nalu_b200_graph *g, int shared, int32_t *colidx);
int nalu_b200_graph_copy_row_gids(const nalu_b200_graph *g, int shared, int64_t *gids);
int nalu_b200_graph_copy_col_gids(const nalu_b200_graph *g, int64_t *gids);
/* entityToLID_ / entityToColLID_ for the local nodes */
int nalu_b200_graph_copy_node_lids(const nalu_b200_graph *g, int32_t *row_lid, int32_t *col_lid);
/* Element colouring used by the general kernel's scatter (elements of one colour share no row node: a colour is summed
 * into the CSR with plain read-modify-writes, colours in ascending order -- deterministic, no atomics).  n_colors = 0:
 * the mesh has no colouring (an element touches one row through two of its nodes; atomic scatter).  color_of_elem
 * (host, [n_elems]) may be NULL. */
int nalu_b200_graph_copy_elem_colors(const nalu_b200_graph *g, int32_t *n_colors, int32_t *color_of_elem);
/* device-resident graph arrays (for a device solver): owned graph */
int nalu_b200_graph_device_ptrs(const nalu_b200_graph *g, int shared, const uint64_t **d_rowptr,
                                const int32_t **d_colidx);

/* ------------------------------------------------------------------------------------ */
/* Matrix / RHS storage.  `values` holds nnz_owned + nnz_shared doubles (owned CSR values */
/* first, shared-not-owned after), `rhs` holds n_owned_rows + n_shared_rows doubles.      */
/* ------------------------------------------------------------------------------------ */
int nalu_b200_zero_system(const nalu_b200_graph *g, double *d_values, double *d_rhs, void *stream);

enum {
  NALU_ASSEMBLE_AUTO = 0,    /* tiled write-once kernel when the mesh is a box and the kernel
                                list is supported by it, else the general kernel            */
  NALU_ASSEMBLE_GENERAL = 1, /* any Hex8 mesh, any kernel list: accumulates (sumInto) into a
                                zeroed or partially assembled system                        */
  NALU_ASSEMBLE_BOX = 2      /* force the tiled kernel (error if unsupported)               */
};
/* AssembleElemSolverAlgorithm::execute for all locally-owned elements.
 * `overwrite` != 0 tells the tiled kernel that the system holds no earlier contributions,
 * so it may store instead of add (zero_system is then unnecessary for the matrix values and
 * RHS rows it covers).  The general kernel always adds.  A graph without elements (an empty
 * part / rank) is a no-op and may be given null value / rhs arrays. */
int nalu_b200_assemble_elem(const nalu_b200_graph *g, const nalu_b200_kernel_desc *kernels,
                            int32_t n_kernels, double *d_values, double *d_rhs, int32_t mode,
                            int32_t overwrite, void *stream);
/* The same for the rows of node planes [z_begin, z_end) of a structured block only (tiled kernel;
 * error if the mesh / kernel list is not supported by it).  The reference's execute() loops over
 * element buckets (src/AssembleElemSolverAlgorithm.C:74-100) and any partition of that loop is a
 * valid schedule; here the partition is by ROW plane, each row being written once by exactly one
 * launch, so that a caller can overlap the host->device copy of the fields of the next planes with
 * the assembly of the current ones.  Needs the nodal fields of planes z_begin-1 .. z_end and the
 * element field of layers z_begin-1 .. z_end-1 to be valid at launch. */
int nalu_b200_assemble_elem_zrange(const nalu_b200_graph *g, const nalu_b200_kernel_desc *kernels,
                                   int32_t n_kernels, double *d_values, double *d_rhs, int32_t overwrite,
                                   int32_t z_begin, int32_t z_end, void *stream);
/* which kernel family the last nalu_b200_assemble_elem call used: 1 general (coloured scatter), 2 tiled write-once,
 * 3 tiled write-once for one kernel of the list + general for the others, 4 row-thread unstructured momentum kernel
 * (any Hex8 connectivity with an element colouring, coloured value-less reductions), 5 Hex27 kernel (CTA per element) */
int nalu_b200_last_assemble_mode(void);
/* Launch-configuration knobs for A/B measurements and tests; none changes what is computed.  name / value:
 *   "box_kernel"  5 row-thread tiled momentum kernel (default), 2 column-pair kernel
 *   "box_zchunk"  node planes per CTA of the tiled kernels, 0 = automatic
 *   "split_lists" 1 (default) a list with one tiled-capable kernel runs tiled + accumulate, 0 general kernel only
 *   "cont_tiled"  1 (default) one-dof lists on structured blocks use the tiled kernels
 *   "urow"            1 (default) momentum lists on non-lattice meshes use the row-thread unstructured kernel, 0 general kernel
 *   "general_minb"    CTAs per SM the general kernel's registers are bounded for: 2, 3 or 4 (default)
 *   "general_scatter" coloured scatter of the general kernel: 1 (default) value-less reductions (red.global.add.f64), 0
 *                     load / add / store; one element of a colour touches an address either way: same sums, bit for bit
 *   "host_threads"    OpenMP threads of the host graph builder (graph_create / graph_finalize) for the calling thread; launchers
 *                     such as torchrun export OMP_NUM_THREADS=1, which makes a 17 M-node graph take 27 s instead of 9
 * The initial values come from the environment variables NALU_BOX_KERNEL, NALU_BOX_ZCHUNK, NALU_SPLIT_LISTS,
 * NALU_CONT_BOX, read once when the library is first used -- nothing on the launch path reads the environment.
 * Returns NALU_B200_ERR_ARG for an unknown name. */
int nalu_b200_tune(const char *name, int32_t value);
/* synchronises, then reports (and clears) the device flag raised when an element had a
 * non-positive Jacobian determinant -- the reference asserts on this in debug builds
 * (include/master_element/MasterElementFunctions.h:91-94) */
int nalu_b200_geometry_errors(const nalu_b200_graph *g, int32_t *flag);

/* TpetraLinearSystem::sumInto for host callers (src/TpetraLinearSystem.C:1457-1522; node algorithms,
 * boundary conditions): `nodes` are local node indices, `rhs` [n] and `lhs` [n][n] row-major with
 * n = n_entities*num_dof, all HOST pointers (copied to the device on `stream`).  Columns absent from
 * a row are skipped, like the reference (:1397,1439-1451).  The element algorithm never calls this:
 * it scatters in bulk inside nalu_b200_assemble_elem. */
int nalu_b200_sum_into(const nalu_b200_graph *g, int32_t n_entities, const int32_t *nodes, const double *rhs,
                       const double *lhs, double *d_values, double *d_rhs, void *stream);

/* AssembleNodeSolverAlgorithm::execute with the mass BDF2/backward-Euler node algorithms;
 * `k` uses kind NALU_K_{MOMENTUM,CONTINUITY,SCALAR}_MASS and NALU_F_DUAL_VOLUME. */
int nalu_b200_assemble_node_mass(const nalu_b200_graph *g, const nalu_b200_kernel_desc *k,
                                 double *d_values, double *d_rhs, void *stream);

/* AssembleNodalGradElemAlgorithm::execute / AssembleNodalGradUElemAlgorithm::execute
 * (src/AssembleNodalGradElemAlgorithm.C:77-133, src/AssembleNodalGradUElemAlgorithm.C:120-195): interior Green-Gauss
 * contributions of every locally-owned element to the projected nodal gradient -- dpdx, dqdx (ncomp = 1 -> [nNodes][3])
 * or dudx (ncomp = 3 -> [nNodes][9], (i,j) at 3*i+j) -- the inputs of the continuity, upwind and mass kernels.
 * d_grad is accumulated into (zero it first; boundary contributions and the parallel sum are the caller's, as in the
 * reference); use_shifted: shifted shape function. */
int nalu_b200_assemble_nodal_grad_elem(const nalu_b200_graph *g, const double *d_coords, const double *d_field,
                                       int32_t ncomp, const double *d_dual_volume, int32_t use_shifted, double *d_grad,
                                       void *stream);
/* ComputeMdotElemAlgorithm::execute (src/ComputeMdotElemAlgorithm.C:85-254): the producer of the element field
 * mass_flow_rate_scs [nElem][12] the adv-diff kernels read, mdot_ip = (rho u_ip - projTimeScale (dpdx_ip - Gpdx_ip)) . areav_ip
 * with projTimeScale = dt / gamma1; options: NALU_OPT_SHIFT_MDOT (shifted shape function), NALU_OPT_SHIFT_POISSON
 * (shifted grad-op).  With it the field never leaves the device between two assemblies. */
int nalu_b200_compute_mdot_elem(const nalu_b200_graph *g, const double *d_coords, const double *d_density,
                                const double *d_velocity, const double *d_pressure, const double *d_gpdx,
                                double proj_time_scale, int32_t options, double *d_mdot, void *stream);
/* the same for the elements [e_begin, e_end) only (a caller that streams the nodal fields onto the device produces the
 * element field range by range, ahead of nalu_b200_assemble_elem_zrange) */
int nalu_b200_compute_mdot_elem_range(const nalu_b200_graph *g, const double *d_coords, const double *d_density,
                                      const double *d_velocity, const double *d_pressure, const double *d_gpdx,
                                      double proj_time_scale, int32_t options, int64_t e_begin, int64_t e_end,
                                      double *d_mdot, void *stream);
/* TpetraLinearSystem::applyDirichletBCs (src/TpetraLinearSystem.C:1579-1653): for every listed local node and
 * dof d in [begin_pos, end_pos) the CSR row is replaced by the identity row (locally-owned rows; a
 * shared-not-owned row is cleared, its owner holds the 1) and the residual by bc - solution (0 for a
 * shared-not-owned row).  d_nodes: device [n_nodes] local node ids (the nodes the reference's selector picks:
 * owned or shared, on the BC parts, with the solution field); d_solution / d_bc_values: device nodal fields
 * [nNodes][numDof].  Call after the assembly, before nalu_b200 loadComplete -- the order of
 * EquationSystem::assemble_and_solve (src/EquationSystem.C:271-283). */
int nalu_b200_apply_dirichlet_bcs(const nalu_b200_graph *g, int32_t n_nodes, const int32_t *d_nodes,
                                  const double *d_solution, const double *d_bc_values, int32_t begin_pos,
                                  int32_t end_pos, double *d_values, double *d_rhs, void *stream);
/* TpetraLinearSystem::resetRows (src/TpetraLinearSystem.C:1706-1755): rows of the listed nodes, dofs
 * [begin_pos, end_pos): every entry (diagonal included) and the residual become 0. */
int nalu_b200_reset_rows(const nalu_b200_graph *g, int32_t n_nodes, const int32_t *d_nodes, int32_t begin_pos,
                         int32_t end_pos, double *d_values, double *d_rhs, void *stream);

/* calc_mass_flow_rate_scs: d_mdot[nElem][12] (Hex27 graphs: [nElem][216]) from coordinates, density, velocity */
int nalu_b200_mdot_scs(const nalu_b200_graph *g, const double *d_coords, const double *d_density,
                       const double *d_velocity, double *d_mdot, void *stream);

/* ------------------------------------------------------------------------------------ */
/* Face/element boundary-condition algorithms.                             */
/* ------------------------------------------------------------------------------------ */
/* Replaces AssembleFaceElemSolverAlgorithm::execute / run_face_elem_algorithm (src/AssembleFaceElemSolverAlgorithm.C:78-108,
 * include/AssembleFaceElemSolverAlgorithm.h:45-133) for Quad4 faces of Hex8 elements (AlgTraitsQuad4Hex8): per exposed face
 * the gather of the face's and the attached element's nodes, HexSCS::face_grad_op / shifted_face_grad_op
 * (src/master_element/Hex8CVFEM.C:668-742), the registered face kernels and apply_coeff with the ELEMENT's connected nodes.
 * A face is (local element, face ordinal); its nodes are the element's side_node_ordinals(ordinal) (Hex8CVFEM.C:414-422),
 * the identity unit_tests/kernels/UnitTestFaceElemBasic.C:48-66 asserts.  Face fields are [n_faces][4 ips][ncomp] arrays in
 * the order the faces were given to nalu_b200_face_set_create. */
enum nalu_b200_face_kernel_kind {
  NALU_FK_MOMENTUM_SYMMETRY = 0,      /* src/kernel/MomentumSymmetryElemKernel.C:68-163     */
  NALU_FK_SCALAR_OPEN_ADV = 1,        /* src/kernel/ScalarOpenAdvElemKernel.C:90-195        */
  NALU_FK_CONTINUITY_OPEN = 2,        /* src/kernel/ContinuityOpenElemKernel.C:80-199       */
  NALU_FK_MOMENTUM_OPEN_ADV_DIFF = 3, /* src/kernel/MomentumOpenAdvDiffElemKernel.C:110-325 */
  /* src/kernel/MomentumWallFunctionElemKernel.C:73-155.  The reference runs it on an AssembleElemSolverAlgorithm over the wall
   * part's FACES (a 12 x 12 matrix on the face's nodes); here it is one more kernel of the face set, adding the same values to
   * the same CSR entries through the attached element's slot map. */
  NALU_FK_MOMENTUM_WALL_FUNCTION = 4,
  NALU_FK_COUNT = 5
};
/* nodal field slots of the face kernels */
enum nalu_b200_face_nodal_slot {
  NALU_FN_COORDINATES = 0,  /* 3 */
  NALU_FN_VELOCITY = 1,     /* 3  velocity NP1 */
  NALU_FN_VELOCITY_RTM = 2, /* 3  velocity_rtm (mesh motion); NULL: NALU_FN_VELOCITY */
  NALU_FN_MESH_VELOCITY = 3,/* 3  mesh_velocity (mesh motion); NULL: NALU_FN_VELOCITY (MomentumOpenAdvDiffElemKernel.C:53-60) */
  NALU_FN_VISCOSITY = 4,    /* 1  viscosity / diffFluxCoeff */
  NALU_FN_DENSITY = 5,      /* 1 */
  NALU_FN_PRESSURE = 6,     /* 1 */
  NALU_FN_PRESSURE_BC = 7,  /* 1  pressure_bc */
  NALU_FN_DPDX = 8,         /* 3 */
  NALU_FN_SCALARQ = 9,      /* 1 */
  NALU_FN_BC_SCALARQ = 10,  /* 1  the open bc's entrainment value */
  NALU_FN_GJQ = 11,         /* 3 */
  NALU_FN_GJUI = 12,        /* 9  dudx */
  NALU_FN_VELOCITY_BC = 13, /* 3  open_velocity_bc */
  NALU_FN_WALL_VELOCITY_BC = 14, /* 3  wall_velocity_bc */
  NALU_FN_COUNT = 16
};
enum nalu_b200_face_field_slot {
  NALU_FF_EXPOSED_AREA_VEC = 0,    /* [4][3] exposed_area_vector (nalu_b200_compute_exposed_area) */
  NALU_FF_OPEN_MASS_FLOW_RATE = 1, /* [4]    open_mass_flow_rate */
  NALU_FF_DYNAMIC_PRESSURE = 2,    /* [4]    dynamic_pressure */
  NALU_FF_WALL_FRICTION_VELOCITY_BIP = 3, /* [4] wall_friction_velocity_bip */
  NALU_FF_WALL_NORMAL_DISTANCE_BIP = 4,   /* [4] wall_normal_distance_bip */
  NALU_FF_COUNT = 6
};
#define NALU_OPT_MESH_MOVES (1 << 8) /* solnOpts.does_mesh_move(): meshVelocityCorrection_ of the momentum open kernel */
typedef struct nalu_b200_face_kernel_desc {
  int32_t kind;    /* nalu_b200_face_kernel_kind */
  int32_t options; /* NALU_OPT_SHIFTED_GRAD_OP, _SKEW_SYMMETRIC, _SHIFT_MDOT, _REDUCED_SENS, _PECLET_TANH, NALU_OPT_MESH_MOVES */
  double include_divu;
  double dt, gamma1; /* ContinuityOpenElemKernel::setup: projTimeScale = dt / gamma1 */
  double upw[8];     /* NALU_UPW_ALPHA_UPW, _HO_UPWIND, _HYBRID_FACTOR, _TANH_TRANS, _TANH_WIDTH */
  const double *nodal[NALU_FN_COUNT]; /* device pointers, NULL when unused */
  const double *face[NALU_FF_COUNT];  /* device pointers */
  double wall[4]; /* wall function: NALU_WALL_ELOG, _KAPPA, _YPLUS_CRIT = get_turb_model_constant(TM_elog / TM_kappa / TM_yplus_crit) */
} nalu_b200_face_kernel_desc;
enum { NALU_WALL_ELOG = 0, NALU_WALL_KAPPA = 1, NALU_WALL_YPLUS_CRIT = 2 };

/* The locally owned faces of a boundary part (host arrays, copied).  The set colours the faces (two faces of one colour
 * share no node, so a colour is summed with plain read-modify-writes: deterministic, no atomics). */
typedef struct nalu_b200_face_set nalu_b200_face_set;
int nalu_b200_face_set_create(nalu_b200_face_set **out, const nalu_b200_graph *g, int64_t n_faces, const int32_t *face_elem,
                              const int32_t *face_ordinal);
void nalu_b200_face_set_destroy(nalu_b200_face_set *fs);
int nalu_b200_face_set_num_colors(const nalu_b200_face_set *fs, int32_t *n_colors);
/* ComputeGeometryBoundaryAlgorithm::execute (src/ComputeGeometryBoundaryAlgorithm.C:62-123): exposed_area_vector of every
 * face of the set = Quad3DSCS::determinant (quad3d_scs_det, MasterElementWork.F:1467-1586); d_area_vec [n_faces][4][3] */
int nalu_b200_compute_exposed_area(const nalu_b200_graph *g, const nalu_b200_face_set *fs, const double *d_coords,
                                   double *d_area_vec, void *stream);
/* AssembleFaceElemSolverAlgorithm::execute: every kernel of the list on every face, summed into the element's rows */
int nalu_b200_assemble_face_elem(const nalu_b200_graph *g, const nalu_b200_face_set *fs, const nalu_b200_face_kernel_desc *k,
                                 int32_t nk, double *d_values, double *d_rhs, void *stream);

/* ------------------------------------------------------------------------------------ */
/* loadComplete: shared-not-owned rows -> owners, ADD.                                   */
/* ------------------------------------------------------------------------------------ */
/* Per neighbour the exchange is a flat array of doubles: for each of my shared-not-owned
 * dof rows owned by `peer` (ascending gid) its CSR values followed, after all rows, by the
 * rows' RHS entries.  The owner adds them through a map built during the graph handshake. */
int nalu_b200_halo_num_peers(const nalu_b200_graph *g, int32_t *n_send_peers, int32_t *n_recv_peers);
int nalu_b200_halo_send_peer(const nalu_b200_graph *g, int32_t i, int32_t *peer_rank, int64_t *n_doubles);
int nalu_b200_halo_recv_peer(const nalu_b200_graph *g, int32_t i, int32_t *peer_rank, int64_t *n_doubles);
/* pack my shared rows for send-peer i into d_buf (contiguous copy: the shared rows of one
 * owner are contiguous in `values` except for the RHS tail) */
int nalu_b200_halo_pack(const nalu_b200_graph *g, int32_t i, const double *d_values,
                        const double *d_rhs, double *d_buf, void *stream);
/* owner side: add the doubles received from recv-peer i (device buffer, which may be a
 * peer-GPU pointer mapped through CUDA IPC / P2P: the kernel then reads straight across
 * NVLink) into my owned rows */
int nalu_b200_halo_unpack_add(const nalu_b200_graph *g, int32_t i, const double *d_buf,
                              double *d_values, double *d_rhs, void *stream);
/* owner side, fused: read recv-peer i's shared rows directly out of the peer's arrays and add
 * them -- no pack, no staging buffer.  d_peer_values / d_peer_rhs are peer device pointers to the
 * START OF THE PEER'S SHARED-NOT-OWNED BLOCK (its values + nnz_owned, its rhs + n_owned_rows) */
int nalu_b200_halo_pull_add(const nalu_b200_graph *g, int32_t i, const double *d_peer_values,
                            const double *d_peer_rhs, double *d_values, double *d_rhs,
                            void *stream);
/* the handshake also tells each owner where recv-peer i's rows start inside the peer's
 * values / rhs arrays (element offsets), needed by halo_pull_add */
int nalu_b200_halo_peer_offsets(const nalu_b200_graph *g, int32_t i, int64_t *values_offset,
                                int64_t *rhs_offset);

/* loadComplete inside the library (TpetraLinearSystem::loadComplete, src/TpetraLinearSystem.C:1757-1778: Export(ADD) of
 * the shared-not-owned matrix and rhs rows to their owners): pack kernel -> ncclSend / ncclRecv per neighbour in one
 * group -> unpack-add kernel, all enqueued on `stream`, exchange buffers persistent in the graph handle.  NCCL is bound
 * at run time (dlopen of libnccl.so.2); the communicator is the library's own: rank 0 asks for a unique id (128 bytes),
 * the caller distributes it (MPI_Bcast, torch.distributed, a file), every rank calls nalu_b200_comm_create.  To overlap
 * the exchange with the assembly, assemble the interface node planes first (nalu_b200_assemble_elem_zrange), call this
 * on a second stream that waits for them, and assemble the interior meanwhile.  A graph without neighbours returns OK
 * at once (comm may be NULL). */
typedef struct nalu_b200_comm nalu_b200_comm;
int nalu_b200_comm_get_unique_id(void *id128);
int nalu_b200_comm_create(nalu_b200_comm **comm, int32_t nranks, int32_t rank, const void *id128);
void nalu_b200_comm_destroy(nalu_b200_comm *comm);
int nalu_b200_load_complete(nalu_b200_graph *g, nalu_b200_comm *comm, double *d_values, double *d_rhs, void *stream);

/* CUDA IPC helpers so that separate processes (one per GPU) can map each other's buffers */
int nalu_b200_ipc_get_handle(const void *d_ptr, void *handle64 /* 64 bytes */);
int nalu_b200_ipc_open_handle(const void *handle64, void **d_ptr);
int nalu_b200_ipc_close(void *d_ptr);

/* device memory helpers (the C++ host layer owns its CSR storage through these) */
int nalu_b200_malloc(void **d_ptr, size_t nbytes);
int nalu_b200_free(void *d_ptr);
int nalu_b200_memcpy_h2d(void *d_dst, const void *h_src, size_t nbytes, void *stream);
int nalu_b200_memcpy_d2h(void *h_dst, const void *d_src, size_t nbytes, void *stream);
int nalu_b200_stream_sync(void *stream);

#ifdef __cplusplus
}
#endif
#endif /* NALU_B200_H */
