/*
 * nalu_oracle.h -- CPU restatement of NaluCFD/Nalu's Hex8 CVFEM element assembly.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under nalu_b200/ (the product) may include,
 * link or call this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, as the checker / CPU baseline.
 *
 * Every function restates, in plain scalar C and in the reference's own loop
 * order, the reference routine cited above it (paths relative to the Nalu
 * source tree).  Parity status: PINNED -- tests/test_oracle_golden.py checks
 * this code against every golden vector the reference's unit tests hold for the
 * path (tests/golden/ *.json, extracted verbatim by tests/golden/make_golden.py).
 *
 * The reference itself cannot be compiled here (needs Trilinos: STK, Kokkos,
 * Tpetra; yaml-cpp; MPI; a Fortran compiler) so there is no oracle/_ref binary.
 *
 * Third-party arithmetic restated from its published algorithm:
 *   Tpetra::Details::shellSortKeysAndValues (Trilinos, version unpinned by the
 *   reference: SQA/buildAndTestScript.sh tracks master/develop) -- Shell sort,
 *   gaps n/2, n/4, ...; the keys it sorts here are distinct column ids, so the
 *   result does not depend on the gap sequence.
 */
#ifndef NALU_ORACLE_H
#define NALU_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ */
/* Hex8 master elements (src/master_element/Hex8CVFEM.C)               */
/* ------------------------------------------------------------------ */
void orc_hex8_scs_intg_loc(int shifted, double loc[36]);      /* Hex8CVFEM.C:318-345 */
void orc_hex8_scv_intg_loc(int shifted, double loc[24]);      /* Hex8CVFEM.C:88-107  */
const int *orc_hex8_lrscv(void);                              /* Hex8CVFEM.C:264-276 */
const int *orc_hex8_scv_ip_node_map(void);                    /* Hex8CVFEM.C:83-85   */

/* MasterElementWork.F:2344-2402 (hex_shape_fcn); shape is [npts][8] */
void orc_hex8_shape_fcn(int npts, const double *par_coord, double *shape);
/* Hex8CVFEM.C:24-70 (hex8_derivative); deriv is [npts][8][3] */
void orc_hex8_derivative(int npts, const double *intgLoc, double *deriv);
/* MasterElementFunctions.h:31-43,51-108 (cofactorMatrix, generic_grad_op);
 * returns the number of ips with det <= tiny (the reference asserts on these) */
int orc_generic_grad_op(int npts, const double *deriv, const double *coords /*[8][3]*/,
                        double *gradop /*[npts][8][3]*/, double *det_j /*[npts] or NULL*/);
/* Hex8GeometryFunctions.h:263-321 */
void orc_subdivide_hex_8(const double *coords /*[8][3]*/, double coordv[27][3]);
/* Hex8CVFEM.C:547-581 + Hex8GeometryFunctions.h:30-76 */
void orc_hex8_scs_areav(const double *coords, double *areav /*[12][3]*/);
/* Hex8CVFEM.C:149-178 + Hex8GeometryFunctions.h:78-163 (Grandy) */
void orc_hex8_scv_volume(const double *coords, double *volume /*[8]*/);
double orc_hex_volume_grandy(const double scvcoords[8][3]);
/* HexSCS::grad_op / shifted_grad_op (Hex8CVFEM.C:523-542), HexSCV::grad_op / shifted (:182-190,:221-229) */
int orc_hex8_scs_grad_op(int shifted, const double *coords, double *gradop, double *deriv);
int orc_hex8_scv_grad_op(int shifted, const double *coords, double *gradop, double *deriv);
void orc_hex8_scs_gij(const double *coords, double *gupper, double *glower); /* [12][3][3] each */

/* ------------------------------------------------------------------ */
/* Element kernels (src/kernel/ *.C)                                    */
/* ------------------------------------------------------------------ */
enum orc_kernel_kind {
  ORC_MOMENTUM_ADV_DIFF = 0, /* src/kernel/MomentumAdvDiffElemKernel.C */
  ORC_CONTINUITY_ADV = 1,    /* src/kernel/ContinuityAdvElemKernel.C   */
  ORC_SCALAR_ADV_DIFF = 2,   /* src/kernel/ScalarAdvDiffElemKernel.C   */
  ORC_MOMENTUM_MASS = 3,     /* src/kernel/MomentumMassElemKernel.C    */
  ORC_CONTINUITY_MASS = 4,   /* src/kernel/ContinuityMassElemKernel.C  */
  ORC_SCALAR_MASS = 5,       /* src/kernel/ScalarMassElemKernel.C      */
  ORC_SCALAR_DIFF = 6,       /* src/kernel/ScalarDiffElemKernel.C      */
  ORC_TEST_FIXED_BLOCK = 7,  /* unit_tests/UnitTestTpetra.C:116-144 TestKernel */
  ORC_MOMENTUM_BUOYANCY_SRC = 8,            /* src/kernel/MomentumBuoyancySrcElemKernel.C           */
  ORC_MOMENTUM_BUOYANCY_BOUSSINESQ_SRC = 9, /* src/kernel/MomentumBuoyancyBoussinesqSrcElemKernel.C */
  ORC_TKE_SST_SRC = 10,                     /* src/kernel/TurbKineticEnergySSTSrcElemKernel.C       */
  ORC_SDR_SST_SRC = 11,                     /* src/kernel/SpecificDissipationRateSSTSrcElemKernel.C */
  ORC_SCALAR_UPW_ADV_DIFF = 12,             /* src/kernel/ScalarUpwAdvDiffElemKernel.C              */
  ORC_MOMENTUM_UPW_ADV_DIFF = 13,           /* src/kernel/MomentumUpwAdvDiffElemKernel.C            */
  ORC_TKE_SST_DES_SRC = 14,                 /* src/kernel/TurbKineticEnergySSTDESSrcElemKernel.C    */
  ORC_TKE_KSGS_SRC = 15,                    /* src/kernel/TurbKineticEnergyKsgsSrcElemKernel.C      */
  ORC_MOMENTUM_NSO = 16                     /* src/nso/MomentumNSOElemKernel.C                      */
};

/* what the kernel constructors/ setup() capture (options + time integrator) */
typedef struct orc_kernel_desc {
  int kind;
  /* options */
  int shiftedGradOp;      /* solnOpts.get_shifted_grad_op(name)                      */
  int skewSymmetric;      /* solnOpts.get_skew_symmetric(name): adv shape fn shifted */
  int lumpedMass;         /* mass kernels                                            */
  int shiftMdot;          /* solnOpts.cvfemShiftMdot_                                */
  int shiftPoisson;       /* solnOpts.get_shifted_grad_op("pressure")                */
  int reducedSensitivities; /* solnOpts.cvfemReducedSensPoisson_                     */
  double includeDivU;     /* solnOpts.includeDivU_                                   */
  double densFac, om_densFac; /* balancedForce_ ? (1,0) : (0,1)                      */
  double dt, gamma1, gamma2, gamma3; /* Kernel::setup(TimeIntegrator)                */
  /* nodal fields, [nNode][ncomp] AoS like an STK field; NULL when unused */
  const double *coordinates; /* [3] */
  const double *velocity;    /* [3]  velocityNp1 (or velocityRTM for continuity) */
  const double *velocityN;   /* [3] */
  const double *velocityNm1; /* [3] */
  const double *viscosity;   /* [1]  viscosity / diffFluxCoeff */
  const double *density;     /* [1]  densityNp1 */
  const double *densityN;    /* [1] */
  const double *densityNm1;  /* [1] */
  const double *pressure;    /* [1] */
  const double *Gpdx;        /* [3]  dpdx */
  const double *scalarQ;     /* [1]  scalarQNp1 */
  const double *scalarQN;    /* [1] */
  const double *scalarQNm1;  /* [1] */
  /* element fields */
  const double *massFlowRate; /* [nElem][12] mass_flow_rate_scs */
  const double *fixedBlock;   /* ORC_TEST_FIXED_BLOCK: rhsSize*rhsSize values written (not added) */
  /* buoyancy sources: solnOpts.gravity_, referenceDensity_, referenceTemperature_, thermalExpansionCoeff_ */
  const double *temperature;  /* [1]  temperatureNp1 */
  double gravity[3];
  double rhoRef, tRef, beta;
  /* SST sources */
  const double *tke, *sdr, *tvisc, *fOneBlend; /* [1] each */
  double betaStar, tkeProdLimitRatio, sigmaWTwo, betaOne, betaTwo, gammaOne, gammaTwo; /* get_turb_model_constant */
  /* upwind adv-diff kernels (ScalarUpwAdvDiffElemKernel.C:28-90, MomentumUpwAdvDiffElemKernel.C:28-90) */
  const double *Gjq;         /* [3]  projected nodal gradient of scalarQ              */
  const double *Gju;         /* [9]  dudx, Gju(node,i,j) at [3*i+j]                   */
  const double *velocityRTM; /* [3]  velocity_rtm when the mesh moves; NULL: velocity */
  double alpha, alphaUpw, hoUpwind; /* get_alpha_factor / get_alpha_upw_factor / get_upw_factor */
  int useLimiter;                   /* primitive_uses_limiter                          */
  int pecletTanh;                   /* 0: ClassicPecletFunction(A=5, hybridFactor), 1: TanhFunction(c1, c2) */
  double hybridFactor, tanhTrans, tanhWidth; /* EquationSystem.h:306-325                */
  /* SST-DES / Ksgs sources */
  const double *maxLengthScale; /* [1] sst_max_length_scale                                  */
  const double *cEps;           /* [1] c_epsilon                                             */
  const double *dsqrtkSq;       /* [1] dsqrtk_dx_sq (the viscosity field when lrksgsfac = 0) */
  const double *dualNodalVolume;/* [1] dual_nodal_volume                                     */
  double cDESke, cDESkw, lrksgsfac;
  /* MomentumNSOElemKernel ctor arguments (src/nso/MomentumNSOElemKernel.C:28-36) */
  double fourthFac, altResFac;
} orc_kernel_desc;

/* One element through the registered kernels, exactly as
 * AssembleElemSolverAlgorithm::execute does for one SIMD lane
 * (src/AssembleElemSolverAlgorithm.C:74-100): gather, ME views, zero lhs/rhs,
 * run every kernel in order.  lhs is [rhsSize][rhsSize] row-major. */
void orc_element_lhs_rhs(const orc_kernel_desc *kernels, int nk, int numDof,
                         const int32_t *elemNodes /*[8]*/, int64_t elemIndex,
                         double *lhs, double *rhs);

/* unit_tests/kernels/UnitTestKernelUtils.C:547-628 (calc_mass_flow_rate_scs) */
void orc_calc_mass_flow_rate_scs(int64_t nElem, const int32_t *elemNodes, const double *coords,
                                 const double *density, const double *velocity, double *mdot);

/* src/AssembleNodalGradElemAlgorithm.C:77-133, src/AssembleNodalGradUElemAlgorithm.C:120-195 */
void orc_assemble_nodal_grad_elem(int64_t nElem, const int32_t *elemNodes, const double *coords, const double *field,
                                  int ncomp, const double *dualNodalVolume, int useShifted, double *grad);
/* src/ComputeMdotElemAlgorithm.C:85-254 */
void orc_compute_mdot_elem(int64_t nElem, const int32_t *elemNodes, const double *coords, const double *density,
                           const double *velocity, const double *pressure, const double *Gpdx, double projTimeScale,
                           int shiftMdot, int shiftPoisson, double *mdot);

/* ------------------------------------------------------------------ */
/* Linear system: graph, sumInto, loadComplete (src/TpetraLinearSystem.C)*/
/* ------------------------------------------------------------------ */
typedef struct orc_linsys orc_linsys; /* all ranks of one (simulated) MPI job */

orc_linsys *orc_linsys_create(int nRanks, int numDof);
void orc_linsys_destroy(orc_linsys *ls);
/* per-rank local mesh: nodes carry naluGlobalId (1-based; periodic subjects alias their
 * monarch's id) and the owning rank; every element listed is locally owned. */
void orc_linsys_set_rank_mesh(orc_linsys *ls, int rank, int32_t nNodes, const int64_t *nodeGid,
                              const int32_t *nodeOwner, int64_t nElem, const int32_t *elemNodes);
/* buildElemToNodeGraph on every rank, then finalizeLinearSystem incl. the
 * neighbour handshake (TpetraLinearSystem.C:190-327,329-367,392-413,1257-1353) */
void orc_linsys_build_elem_to_node_graph(orc_linsys *ls);
/* buildNodeGraph (TpetraLinearSystem.C:369-390): call before finalize if wanted */
void orc_linsys_finalize(orc_linsys *ls);

/* sizes / arrays after finalize */
int32_t orc_linsys_num_owned_rows(const orc_linsys *ls, int rank);
int32_t orc_linsys_num_shared_rows(const orc_linsys *ls, int rank);
int32_t orc_linsys_num_cols(const orc_linsys *ls, int rank);
const size_t *orc_linsys_rowptr(const orc_linsys *ls, int rank, int shared);   /* size rows+1 */
const int32_t *orc_linsys_colidx(const orc_linsys *ls, int rank, int shared);
double *orc_linsys_values(orc_linsys *ls, int rank, int shared);
double *orc_linsys_rhs(orc_linsys *ls, int rank, int shared);
const int64_t *orc_linsys_row_gids(const orc_linsys *ls, int rank, int shared); /* ownedRowsMap_ / sharedNotOwnedRowsMap_ */
const int64_t *orc_linsys_col_gids(const orc_linsys *ls, int rank);             /* totalColsMap_ */
const int32_t *orc_linsys_node_row_lid(const orc_linsys *ls, int rank);         /* entityToLID_    [nNodes] */
const int32_t *orc_linsys_node_col_lid(const orc_linsys *ls, int rank);         /* entityToColLID_ [nNodes] */

void orc_linsys_zero(orc_linsys *ls);                                            /* :1355-1372 */
/* TpetraLinearSystem::sumInto (:1457-1522) for one entity set */
void orc_linsys_sum_into(orc_linsys *ls, int rank, int numEntities, const int32_t *entities,
                         const double *rhs, const double *lhs);
/* AssembleElemSolverAlgorithm::execute over every locally-owned element of `rank` */
void orc_assemble_elem(orc_linsys *ls, int rank, const orc_kernel_desc *kernels, int nk);
/* loadComplete: Export(ADD) of shared-not-owned rows to their owners (:1757-1778) */
void orc_linsys_load_complete(orc_linsys *ls);

/* Node mass supplemental algorithms through AssembleNodeSolverAlgorithm
 * (src/AssembleNodeSolverAlgorithm.C:57-111, src/MomentumMassBDF2NodeSuppAlg.C:80-103,
 *  src/ContinuityMassBDF2NodeSuppAlg.C, src/ScalarMassBDF2NodeSuppAlg.C) */
void orc_assemble_node_mass(orc_linsys *ls, int rank, const orc_kernel_desc *k,
                            const double *dualNodalVolume);

/* ------------------------------------------------------------------ */
/* Face/element boundary-condition path (oracle/face_kernels.c)         */
/* ------------------------------------------------------------------ */
enum orc_face_kernel_kind {
  ORC_FK_MOMENTUM_SYMMETRY = 0,     /* src/kernel/MomentumSymmetryElemKernel.C     */
  ORC_FK_SCALAR_OPEN_ADV = 1,       /* src/kernel/ScalarOpenAdvElemKernel.C        */
  ORC_FK_CONTINUITY_OPEN = 2,       /* src/kernel/ContinuityOpenElemKernel.C       */
  ORC_FK_MOMENTUM_OPEN_ADV_DIFF = 3,/* src/kernel/MomentumOpenAdvDiffElemKernel.C  */
  ORC_FK_MOMENTUM_WALL_FUNCTION = 4 /* src/kernel/MomentumWallFunctionElemKernel.C */
};
/* what the face kernels' constructors / setup() capture; nodal fields [nNode][ncomp], face fields [nFaces][4 ips][ncomp] */
typedef struct orc_face_kernel_desc {
  int kind;
  int shiftedGradOp, skewSymmetric, shiftMdot, reducedSensitivities, meshMoves;
  int pecletTanh;
  double includeDivU, dt, gamma1, alphaUpw, hoUpwind, hybridFactor, tanhTrans, tanhWidth;
  const double *coordinates, *velocity, *velocityRTM, *meshVelocity, *viscosity, *density, *pressure, *pressureBc, *Gpdx;
  const double *scalarQ, *bcScalarQ, *Gjq, *Gjui, *velocityBc;
  const double *exposedAreaVec, *openMassFlowRate, *dynamicPressure; /* face fields: [.][4][3], [.][4], [.][4] */
  /* MomentumWallFunctionElemKernel: wall_velocity_bc (nodal), wall_friction_velocity_bip / wall_normal_distance_bip ([.][4]),
   * TM_elog / TM_kappa / TM_yplus_crit (src/SolutionOptions.C) */
  const double *wallVelocityBc, *wallFrictionVelocityBip, *wallNormalDistanceBip;
  double elog, kappa, yplusCrit;
} orc_face_kernel_desc;

const int *orc_hex8_side_node_ordinals(int ordinal); /* Hex8CVFEM.C:414-422,587-591 */
const int *orc_hex8_face_ip_node_map(int ordinal);   /* Hex8CVFEM.C:381-393 */
int orc_hex8_opposing_node(int ordinal, int ip);     /* Hex8CVFEM.C:288-300,853-858 */
int orc_hex8_face_grad_op(int face_ordinal, int shifted, const double *coords /*[8][3]*/, double *gradop /*[4][8][3]*/);
void orc_quad3d_scs_areav(const double *coords /*[4][3]*/, double *area /*[4][3]*/);
void orc_quad3d_shape_fcn(int shifted, double shp[4][4]);
/* src/ComputeGeometryBoundaryAlgorithm.C:62-123; a face = (element, face ordinal) */
void orc_compute_exposed_area(int64_t nFaces, const int32_t *faceElem, const int32_t *faceOrdinal, const int32_t *elemNodes,
                              const double *coords, double *areaVec);
void orc_face_elem_lhs_rhs(const orc_face_kernel_desc *kernels, int nk, int numDof, const int32_t *elemNodes, int elemFaceOrdinal,
                           int64_t faceIndex, double *lhs, double *rhs);
/* AssembleFaceElemSolverAlgorithm::execute (src/AssembleFaceElemSolverAlgorithm.C:78-108) over the listed faces */
void orc_assemble_face_elem(orc_linsys *ls, int rank, int numDof, const int32_t *elemNodes, const orc_face_kernel_desc *kernels,
                            int nk, int64_t nFaces, const int32_t *faceElem, const int32_t *faceOrdinal);

/* LocalGraphArrays test hooks (include/LinearSolver.h:62-120) */
void *orc_graph_arrays_create(const size_t *rowLens, size_t n);
void orc_graph_arrays_insert(void *h, size_t localRow, size_t numInds, const int32_t *inds, int numDof);
const int32_t *orc_graph_arrays_colidx(void *h);
const size_t *orc_graph_arrays_rowptr(void *h);
void orc_graph_arrays_destroy(void *h);

/* Tpetra::Details::shellSortKeysAndValues restatement */
void orc_shell_sort_keys_and_values(int *keys, int *values, int n);

/* unit_tests/UnitTestUtils.C:37-72 perturb_coord_hex_8 LCG */
double orc_lcg_next(uint32_t *state);

#ifdef __cplusplus
}
#endif
#endif
