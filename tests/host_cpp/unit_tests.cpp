// C++ unit tests of the host API, written the way the reference's own kernel / linear-system unit tests
// are (unit_tests/kernels/UnitTestMomentumAdvDiffElem.C:274-305, UnitTestContinuityAdvElem.C,
// UnitTestScalarAdvDiffElem.C, UnitTestMomentumMassElem.C, UnitTestTpetra.C:86-203,
// UnitTestKokkosViews.C): build a generated Hex8 mesh and the fixture's analytic fields, make the
// helper objects (realm, equation system, linear system, AssembleElemSolverAlgorithm), construct the
// kernel with the reference's constructor arguments, push it into activeKernels_, execute(), compare the
// captured lhs/rhs with the reference's golden tables.
//
//   unit_tests <golden_dir> cpu     registration / error-convention tests, no device needed
//   unit_tests <golden_dir> gpu     everything (runs the CUDA path through the C ABI)
#include "AssembleElemSolverAlgorithm.h"
#include "AssembleFaceElemSolverAlgorithm.h"
#include "ScratchViews.h"
#include "kernel/ContinuityAdvElemKernel.h"
#include "kernel/KernelBuilder.h"
#include "kernel/MassElemKernels.h"
#include "kernel/MomentumAdvDiffElemKernel.h"
#include "kernel/MomentumBuoyancySrcElemKernel.h"
#include "kernel/SSTSrcElemKernels.h"
#include "kernel/ScalarAdvDiffElemKernel.h"
#include "kernel/UpwAdvDiffElemKernels.h"
#include "nso/MomentumNSOElemKernel.h"

#include <cmath>
#include <cstdio>
#include <fstream>
#include <functional>
#include <iostream>
#include <memory>
#include <random>

using namespace sierra::nalu;

static int g_failures = 0, g_checks = 0;
#define EXPECT_TRUE(c)                                                                   \
  do {                                                                                   \
    ++g_checks;                                                                          \
    if (!(c)) {                                                                          \
      ++g_failures;                                                                      \
      std::printf("  FAILED %s:%d: %s\n", __FILE__, __LINE__, #c);                       \
    }                                                                                    \
  } while (0)
#define EXPECT_NEAR(a, b, tol)                                                           \
  do {                                                                                   \
    ++g_checks;                                                                          \
    if (!(std::fabs((a) - (b)) <= (tol))) {                                              \
      ++g_failures;                                                                      \
      std::printf("  FAILED %s:%d: |%.17g - %.17g| > %g\n", __FILE__, __LINE__, (double)(a), (double)(b), (double)(tol)); \
    }                                                                                    \
  } while (0)
#define EXPECT_THROW(stmt)                                                               \
  do {                                                                                   \
    ++g_checks;                                                                          \
    bool thrown = false;                                                                 \
    try { stmt; } catch (const std::exception &) { thrown = true; }                      \
    if (!thrown) { ++g_failures; std::printf("  FAILED %s:%d: no throw: %s\n", __FILE__, __LINE__, #stmt); } \
  } while (0)

static std::string g_golden;
struct Table { std::vector<int> dims; std::vector<double> v; };
static Table gold(const std::string &file, const std::string &name)
{
  std::ifstream f(g_golden + "/" + file + "__" + name + ".txt");
  if (!f) throw std::runtime_error("missing golden table " + file + " " + name);
  Table t;
  int nd; f >> nd;
  t.dims.resize(nd);
  size_t n = 1;
  for (int &d : t.dims) { f >> d; n *= d; }
  t.v.resize(n);
  for (double &x : t.v) f >> x;
  return t;
}

namespace unit_test_utils {
const double pi = std::acos(-1.0), a = 0.3;

// unit_tests/UnitTestUtils.C:37-72 perturb_coord_hex_8: three LCG draws per node, ascending id
void perturb_coord_hex_8(stk::mesh::BulkData &bulk, stk::mesh::Field<double> &coords, double scale = 0.125)
{
  uint32_t prev = bulk.parallel_rank() + 1;
  auto lcg = [&]() {
    prev = (uint32_t)(((uint64_t)1103515245u * prev + 12345u) % 2147483647u);
    return 2.0 * ((double)prev / 2147483647.0) - 1.0;
  };
  for (size_t n = 0; n < bulk.num_nodes(); ++n)
    for (int d = 0; d < 3; ++d) coords.host_data()[3 * n + d] += scale * lcg();
  coords.modify_on_host();
}

// unit_tests/kernels/UnitTestKernelUtils.h:217-352 (LowMachKernelHex8Mesh / MomentumKernelHex8Mesh)
struct LowMachKernelHex8Mesh {
  LowMachKernelHex8Mesh(int nx = 1, int ny = 1, int nz = 1) : meta_(3), bulk_(std::make_shared<stk::mesh::BulkData>(meta_))
  {
    part_ = &meta_.declare_part_with_topology("block_1", stk::topology::HEX_8);
    partVec_ = {part_};
    coordinates_ = &meta_.declare_field<double>(stk::topology::NODE_RANK, "coordinates");
    velocity_ = &meta_.declare_field<double>(stk::topology::NODE_RANK, "velocity", 2);
    dpdx_ = &meta_.declare_field<double>(stk::topology::NODE_RANK, "dpdx", 2);
    density_ = &meta_.declare_field<double>(stk::topology::NODE_RANK, "density", 2);
    pressure_ = &meta_.declare_field<double>(stk::topology::NODE_RANK, "pressure", 2);
    viscosity_ = &meta_.declare_field<double>(stk::topology::NODE_RANK, "viscosity");
    scalarQ_ = &meta_.declare_field<double>(stk::topology::NODE_RANK, "scalarQ", 2);
    diffFluxCoeff_ = &meta_.declare_field<double>(stk::topology::NODE_RANK, "diffFluxCoeff");
    dualNodalVolume_ = &meta_.declare_field<double>(stk::topology::NODE_RANK, "dual_nodal_volume");
    temperature_ = &meta_.declare_field<double>(stk::topology::NODE_RANK, "temperature"); // UnitTestKernelUtils.h:330-350
    // SSTKernelHex8Mesh (UnitTestKernelUtils.h:362-405)
    tke_ = &meta_.declare_field<double>(stk::topology::NODE_RANK, "turbulent_ke");
    sdr_ = &meta_.declare_field<double>(stk::topology::NODE_RANK, "specific_dissipation_rate");
    tvisc_ = &meta_.declare_field<double>(stk::topology::NODE_RANK, "turbulent_viscosity");
    fOneBlend_ = &meta_.declare_field<double>(stk::topology::NODE_RANK, "sst_f_one_blending");
    maxLengthScale_ = &meta_.declare_field<double>(stk::topology::NODE_RANK, "sst_max_length_scale");
    // projected nodal gradients of the upwind kernels (dqdx / dudx in the reference's equation systems)
    dqdx_ = &meta_.declare_field<double>(stk::topology::NODE_RANK, "dqdx");
    dudx_ = &meta_.declare_field<double>(stk::topology::NODE_RANK, "dudx");
    massFlowRate_ = &meta_.declare_field<double>(stk::topology::ELEM_RANK, "mass_flow_rate_scs");
    stk::mesh::put_field_on_mesh(*coordinates_, *part_, 3);
    stk::mesh::put_field_on_mesh(*velocity_, *part_, 3);
    stk::mesh::put_field_on_mesh(*dpdx_, *part_, 3);
    stk::mesh::put_field_on_mesh(*density_, *part_, 1);
    stk::mesh::put_field_on_mesh(*pressure_, *part_, 1);
    stk::mesh::put_field_on_mesh(*viscosity_, *part_, 1);
    stk::mesh::put_field_on_mesh(*scalarQ_, *part_, 1);
    stk::mesh::put_field_on_mesh(*diffFluxCoeff_, *part_, 1);
    stk::mesh::put_field_on_mesh(*dualNodalVolume_, *part_, 1);
    stk::mesh::put_field_on_mesh(*temperature_, *part_, 1);
    for (auto *f : {tke_, sdr_, tvisc_, fOneBlend_, maxLengthScale_}) stk::mesh::put_field_on_mesh(*f, *part_, 1);
    stk::mesh::put_field_on_mesh(*dqdx_, *part_, 3);
    stk::mesh::put_field_on_mesh(*dudx_, *part_, 9);
    stk::mesh::put_field_on_mesh(*massFlowRate_, *part_, 12);
    bulk_->fill_generated_mesh(nx, ny, nz, *part_);
  }
  // steady Taylor vortex etc. (unit_tests/kernels/UnitTestKernelUtils.C:23-274); only StateNP1 is filled
  void fill_mesh_and_init_fields(bool doPerturb = false)
  {
    if (doPerturb) perturb_coord_hex_8(*bulk_, *coordinates_);
    const double *x = coordinates_->host_data();
    for (size_t n = 0; n < bulk_->num_nodes(); ++n) {
      const double X = x[3 * n], Y = x[3 * n + 1];
      double *u = velocity_->host_data() + 3 * n;
      u[0] = -std::cos(a * pi * X) * std::sin(a * pi * Y);
      u[1] = +std::sin(a * pi * X) * std::cos(a * pi * Y);
      u[2] = 0.0;
      pressure_->host_data()[n] = -0.25 * (std::cos(2 * a * pi * X) + std::cos(2 * a * pi * Y));
      double *g = dpdx_->host_data() + 3 * n;
      g[0] = 0.5 * a * pi * std::sin(2 * a * pi * X);
      g[1] = 0.5 * a * pi * std::sin(2 * a * pi * Y);
      g[2] = 0.0;
      density_->host_data()[n] = 1.0;
      viscosity_->host_data()[n] = 0.1;
      temperature_->host_data()[n] = 300.0; // stk::mesh::field_fill(300.0, *temperature_), :350
    }
    for (auto *f : {velocity_, pressure_, dpdx_, density_, viscosity_, temperature_}) f->modify_on_host();
  }
  // SSTKernelHex8Mesh::fill_mesh_and_init_fields (UnitTestKernelUtils.h:385-398) on top of the above:
  // tvisc = 0.3 and the TrigFieldFunction density / tke / sdr / f-one-blending (UnitTestKernelUtils.C:93-220)
  void fill_sst_fields()
  {
    const double *x = coordinates_->host_data();
    for (size_t n = 0; n < bulk_->num_nodes(); ++n) {
      const double X = x[3 * n], Y = x[3 * n + 1], Z = x[3 * n + 2];
      tvisc_->host_data()[n] = 0.3;
      density_->host_data()[n] = std::cos(a * pi * X) * std::cos(a * pi * Y) * std::cos(a * pi * Z);
      tke_->host_data()[n] = 2.0 + std::cos(a * pi * X) * std::sin(a * pi * Y) * std::cos(a * pi * Z);
      sdr_->host_data()[n] = 2.0 + std::cos(a * pi * X) * std::sin(a * pi * Y) * std::sin(a * pi * Z);
      fOneBlend_->host_data()[n] = std::sin(a * pi * X) * std::cos(a * pi * Y) * std::sin(a * pi * Z);
      maxLengthScale_->host_data()[n] = 0.5; // stk::mesh::field_fill(0.5, *maxLengthScale_), :391
    }
    for (auto *f : {density_, tke_, sdr_, tvisc_, fOneBlend_, maxLengthScale_}) f->modify_on_host();
  }
  stk::mesh::MetaData meta_;
  std::shared_ptr<stk::mesh::BulkData> bulk_;
  stk::mesh::Part *part_;
  stk::mesh::PartVector partVec_;
  SolutionOptions solnOpts_;
  VectorFieldType *coordinates_, *velocity_, *dpdx_, *dqdx_;
  GenericFieldType *dudx_;
  ScalarFieldType *density_, *pressure_, *viscosity_, *scalarQ_, *diffFluxCoeff_, *dualNodalVolume_, *temperature_, *tke_, *sdr_, *tvisc_, *fOneBlend_, *maxLengthScale_;
  GenericFieldType *massFlowRate_;
};

// unit_tests/UnitTestHelperObjects.h:16-54
struct HelperObjects {
  HelperObjects(std::shared_ptr<stk::mesh::BulkData> bulk, stk::topology topo, int numDof, stk::mesh::Part *part)
      : eqSystem(realm, "eq")
  {
    realm.bulkData_ = bulk;
    linsys = dynamic_cast<DeviceCsrLinearSystem *>(LinearSystem::create(realm, numDof, &eqSystem, nullptr));
    eqSystem.linsys_ = linsys;
    assembleElemSolverAlg = new AssembleElemSolverAlgorithm(realm, part, &eqSystem, topo.rank(), topo.num_nodes(), false);
  }
  ~HelperObjects()
  {
    assembleElemSolverAlg->activeKernels_.clear(); // the tests own their kernels, like the reference's
    delete assembleElemSolverAlg;
  }
  /** initialize_connectivity + finalize + zero: what EquationSystem::initialize does before the first assembly */
  void build_system()
  {
    assembleElemSolverAlg->initialize_connectivity();
    linsys->finalizeLinearSystem();
    linsys->zeroSystem();
  }
  /** what TestLinearSystem::sumInto captures in the reference (unit_tests/UnitTestLinearSystem.h:40-61): the
   *  element lhs/rhs, here rebuilt from the assembled one-element CSR */
  void element_lhs_rhs(std::vector<double> &lhs, std::vector<double> &rhs)
  {
    const int ndof = linsys->numDof(), n = 8 * ndof;
    const auto rp = linsys->ownedRowPtr();
    const auto ci = linsys->ownedColIdx();
    const auto v = linsys->ownedValues();
    const auto r = linsys->ownedRhs();
    const int32_t *en = realm.bulk_data().begin_nodes(stk::mesh::Entity{stk::topology::ELEM_RANK, 0});
    std::vector<int> rows;
    for (int k = 0; k < 8; ++k)
      for (int d = 0; d < ndof; ++d) rows.push_back(linsys->entityToLID(stk::mesh::Entity{stk::topology::NODE_RANK, en[k]}) + d);
    lhs.assign(size_t(n) * n, 0.0);
    rhs.resize(n);
    for (int i = 0; i < n; ++i) {
      rhs[i] = r[rows[i]];
      for (uint64_t q = rp[rows[i]]; q < rp[rows[i] + 1]; ++q)
        for (int j = 0; j < n; ++j)
          if (ci[q] == rows[j]) lhs[size_t(i) * n + j] = v[q];
    }
  }
  Realm realm;
  EquationSystem eqSystem;
  DeviceCsrLinearSystem *linsys;
  AssembleElemSolverAlgorithm *assembleElemSolverAlg;
};

void expect_all_near(const std::vector<double> &calc, const Table &gold, double tol = 1.0e-15)
{
  EXPECT_TRUE(calc.size() == gold.v.size());
  for (size_t i = 0; i < calc.size() && i < gold.v.size(); ++i) EXPECT_NEAR(calc[i], gold.v[i], tol);
}

// calc_mass_flow_rate_scs (unit_tests/kernels/UnitTestKernelUtils.C:547-628), on the device
void calc_mass_flow_rate_scs(HelperObjects &h, LowMachKernelHex8Mesh &fx)
{
  void *d = nullptr;
  const size_t n = fx.bulk_->num_elems() * 12;
  if (nalu_b200_malloc(&d, n * 8) != 0) throw std::runtime_error("malloc");
  if (nalu_b200_mdot_scs(h.linsys->graph(), fx.coordinates_->device_data(), fx.density_->device_data(), fx.velocity_->device_data(),
                         (double *)d, nullptr) != 0)
    throw std::runtime_error(nalu_b200_last_error());
  nalu_b200_memcpy_d2h(fx.massFlowRate_->host_data(), d, n * 8, nullptr);
  nalu_b200_stream_sync(nullptr);
  nalu_b200_free(d);
  fx.massFlowRate_->modify_on_host();
}
} // namespace unit_test_utils

using unit_test_utils::HelperObjects;
using unit_test_utils::LowMachKernelHex8Mesh;

// ---------------------------------------------------------------------------------------------------
// CPU-only: registration and error conventions
// ---------------------------------------------------------------------------------------------------
static void test_elem_data_requests()
{
  std::printf("ElemDataRequests: registration order and conflicting sizes (include/ElemDataRequests.h:89-101, src/ElemDataRequests.C:21-27)\n");
  LowMachKernelHex8Mesh fx;
  ElemDataRequests req;
  EXPECT_THROW(req.add_master_element_call(SCS_AREAV, CURRENT_COORDINATES)); // coordinates must come first
  req.add_coordinates_field(*fx.coordinates_, 3, CURRENT_COORDINATES);
  req.add_master_element_call(SCS_AREAV, CURRENT_COORDINATES);
  req.add_gathered_nodal_field(*fx.velocity_, 3);
  req.add_gathered_nodal_field(*fx.velocity_, 3); // same size twice is fine
  EXPECT_THROW(req.add_gathered_nodal_field(*fx.velocity_, 1));
  EXPECT_THROW(req.add_gathered_nodal_field(*fx.massFlowRate_, 12)); // element field through the nodal door
  EXPECT_THROW(req.add_element_field(*fx.velocity_, 3));
  req.add_element_field(*fx.massFlowRate_, 12);
  EXPECT_TRUE(req.get_fields().size() == 3);
  EXPECT_TRUE(req.get_data_enums(CURRENT_COORDINATES).count(SCS_AREAV) == 1);
  EXPECT_THROW(req.get_coordinates_field(MODEL_COORDINATES));
}

// the reference's per-element scratch on the host: view shapes, lookup by field ordinal, gather, master-element views
// (include/ScratchViews.h:111-175,584-704; src/ScratchViews.C:408-484; unit_tests/UnitTestScratchViews.C)
static void test_scratch_views()
{
  std::printf("ScratchViews: shapes, gather and master-element views of one Hex8 element (include/ScratchViews.h:111-175)\n");
  LowMachKernelHex8Mesh fx;
  fx.fill_mesh_and_init_fields();
  ElemDataRequests req;
  MomentumAdvDiffElemKernel<AlgTraitsHex8> k(*fx.bulk_, fx.solnOpts_, fx.velocity_, fx.viscosity_, req);
  ScratchViews<double> sv(TeamHandleType(), *fx.bulk_, 8, req);
  // coords 24 + velocity 24 + viscosity 8 + mdot 12 + areav 36 + dndx 288 + deriv 288 + det_j 12 = 692 doubles
  EXPECT_TRUE(sv.total_bytes() == 692 * 8);
  EXPECT_THROW(sv.get_scratch_view_1D(*fx.pressure_));              // not a pre-req field of this kernel
  EXPECT_THROW(sv.get_scratch_view_1D(*fx.velocity_));              // registered as a vector
  stk::mesh::Entity elem;
  elem.rank = stk::topology::ELEM_RANK;
  elem.index = 0;
  fill_pre_req_data(req, *fx.bulk_, elem, sv);
  auto &u = sv.get_scratch_view_2D(*fx.velocity_);
  auto &mu = sv.get_scratch_view_1D(*fx.viscosity_);
  EXPECT_TRUE(u.extent(0) == 8 && u.extent(1) == 3 && mu.extent(0) == 8);
  const int32_t *nodes = fx.bulk_->begin_nodes(elem);
  EXPECT_TRUE(sv.elemNodes == nodes);
  double err = 0;
  for (int n = 0; n < 8; ++n) {
    err = std::fmax(err, std::fabs(mu(n) - fx.viscosity_->host_data()[nodes[n]]));
    for (int d = 0; d < 3; ++d) err = std::fmax(err, std::fabs(u(n, d) - fx.velocity_->host_data()[nodes[n] * 3 + d]));
  }
  EXPECT_TRUE(err == 0.0);
  // unit cube: every sub-control surface has area 1/4 along its edge, from the left to the right node
  // (HexSCS::determinant, Hex8CVFEM.C:547-581); the gradient operator sums to zero over the nodes
  auto &me = sv.get_me_views(CURRENT_COORDINATES);
  HexSCS scs;
  const int *lr = scs.adjacentNodes();
  auto &x = sv.get_scratch_view_2D(*fx.coordinates_);
  for (int ip = 0; ip < 12; ++ip) {
    double dot = 0, a2 = 0, sum[3] = {0, 0, 0};
    for (int d = 0; d < 3; ++d) {
      dot += me.scs_areav(ip, d) * (x(lr[2 * ip + 1], d) - x(lr[2 * ip], d));
      a2 += me.scs_areav(ip, d) * me.scs_areav(ip, d);
      for (int n = 0; n < 8; ++n) sum[d] += me.dndx(ip, n, d);
    }
    EXPECT_NEAR(dot, 0.25, 1e-15);
    EXPECT_NEAR(a2, 0.0625, 1e-15);
    EXPECT_NEAR(std::fabs(sum[0]) + std::fabs(sum[1]) + std::fabs(sum[2]), 0.0, 1e-14);
    EXPECT_NEAR(me.det_j(ip), 1.0, 1e-15);
  }
  EXPECT_THROW(sv.get_me_views(MODEL_COORDINATES));
}

static void test_kernel_registration()
{
  std::printf("Kernel constructors register what the reference's do (src/kernel/*ElemKernel.C ctors)\n");
  LowMachKernelHex8Mesh fx;
  {
    ElemDataRequests req;
    MomentumAdvDiffElemKernel<AlgTraitsHex8> k(*fx.bulk_, fx.solnOpts_, fx.velocity_, fx.viscosity_, req);
    EXPECT_TRUE(req.get_fields().size() == 4); // coordinates, velocity, viscosity, mass_flow_rate_scs
    EXPECT_TRUE(req.get_data_enums(CURRENT_COORDINATES) == (std::set<ELEM_DATA_NEEDED>{SCS_AREAV, SCS_GRAD_OP}));
    EXPECT_TRUE(req.get_cvfem_surface_me() != nullptr && req.get_cvfem_volume_me() == nullptr);
    EXPECT_NEAR(k.v_shape_function_[0][0], 0.28125, 0.0); // (1/2)(3/4)(3/4): ip 0 sits next to nodes 0,1
  }
  {
    SolutionOptions so;
    so.shiftedGradOpMap_["velocity"] = true;
    ElemDataRequests req;
    MomentumAdvDiffElemKernel<AlgTraitsHex8> k(*fx.bulk_, so, fx.velocity_, fx.viscosity_, req);
    EXPECT_TRUE(req.get_data_enums(CURRENT_COORDINATES) == (std::set<ELEM_DATA_NEEDED>{SCS_AREAV, SCS_SHIFTED_GRAD_OP}));
  }
  {
    SolutionOptions so;
    so.cvfemReducedSensPoisson_ = true;
    ElemDataRequests req;
    ContinuityAdvElemKernel<AlgTraitsHex8> k(*fx.bulk_, so, req);
    EXPECT_TRUE(req.get_fields().size() == 5);
    EXPECT_TRUE(req.get_data_enums(CURRENT_COORDINATES) == (std::set<ELEM_DATA_NEEDED>{SCS_AREAV, SCS_GRAD_OP, SCS_SHIFTED_GRAD_OP}));
  }
  {
    ElemDataRequests req;
    MomentumMassElemKernel<AlgTraitsHex8> k(*fx.bulk_, fx.solnOpts_, req, true);
    // density N/NP1, velocity N/NP1 (NM1 aliases N with two states), dpdx, coordinates
    EXPECT_TRUE(req.get_fields().size() == 6);
    EXPECT_TRUE(req.get_data_enums(CURRENT_COORDINATES) == (std::set<ELEM_DATA_NEEDED>{SCV_VOLUME}));
    EXPECT_NEAR(k.v_shape_function_[3][3], 1.0, 0.0); // lumped: shifted SCV shape functions are the identity
    EXPECT_NEAR(k.v_shape_function_[3][4], 0.0, 0.0);
  }
  {
    // shared requests: two kernels on one algorithm take the union; a size conflict throws
    ElemDataRequests req;
    MomentumMassElemKernel<AlgTraitsHex8> m(*fx.bulk_, fx.solnOpts_, req, false);
    MomentumAdvDiffElemKernel<AlgTraitsHex8> k(*fx.bulk_, fx.solnOpts_, fx.velocity_, fx.viscosity_, req);
    EXPECT_TRUE(req.get_fields().size() == 8);
    EXPECT_TRUE(req.get_data_enums(CURRENT_COORDINATES) == (std::set<ELEM_DATA_NEEDED>{SCS_AREAV, SCS_GRAD_OP, SCV_VOLUME}));
  }
  {
    // KernelBuilder: name-gated creation in call order (include/kernel/KernelBuilder.h:170-191)
    ElemSrcTermsMap terms;
    terms["momentum"] = {"advection_diffusion", "lumped_momentum_time_derivative"};
    std::vector<Kernel *> kernels;
    ElemDataRequests req;
    const stk::topology topo = stk::topology::HEX_8;
    bool made = build_topo_kernel_if_requested<MomentumMassElemKernel>(topo, terms, "momentum", kernels, "momentum_time_derivative",
                                                                        *fx.bulk_, fx.solnOpts_, req, false);
    EXPECT_TRUE(!made);
    made = build_topo_kernel_if_requested<MomentumMassElemKernel>(topo, terms, "momentum", kernels, "lumped_momentum_time_derivative",
                                                                   *fx.bulk_, fx.solnOpts_, req, true);
    EXPECT_TRUE(made);
    made = build_topo_kernel_if_requested<MomentumAdvDiffElemKernel>(topo, terms, "momentum", kernels, "advection_diffusion", *fx.bulk_,
                                                                      fx.solnOpts_, fx.velocity_, fx.viscosity_, req);
    EXPECT_TRUE(made && kernels.size() == 2);
    EXPECT_TRUE(std::string(kernels[0]->name()) == "MomentumMassElemKernel"); // source order, not YAML order
    EXPECT_THROW(build_topo_kernel<MomentumAdvDiffElemKernel>(stk::topology(stk::topology::HEX_27), *fx.bulk_, fx.solnOpts_, fx.velocity_,
                                                               fx.viscosity_, req));
    for (Kernel *k : kernels) delete k;
  }
  {
    MasterElement *scs = MasterElementRepo::get_surface_master_element(stk::topology::HEX_8);
    const int lr[24] = {0, 1, 1, 2, 2, 3, 0, 3, 4, 5, 5, 6, 6, 7, 4, 7, 0, 4, 1, 5, 2, 6, 3, 7};
    for (int i = 0; i < 24; ++i) EXPECT_TRUE(scs->adjacentNodes()[i] == lr[i]);
    double shp[96];
    scs->shifted_shape_fcn(shp); // edge midpoints: 1/2 on the two edge nodes
    EXPECT_NEAR(shp[0], 0.5, 0.0);
    EXPECT_NEAR(shp[1], 0.5, 0.0);
    EXPECT_NEAR(shp[2], 0.0, 0.0);
    EXPECT_THROW(scs->determinant(1, nullptr, nullptr, nullptr));
  }
}

// ---------------------------------------------------------------------------------------------------
// GPU: golden kernel tests
// ---------------------------------------------------------------------------------------------------
static void test_momentum_advection_diffusion()
{
  std::printf("MomentumKernelHex8Mesh.advection_diffusion (unit_tests/kernels/UnitTestMomentumAdvDiffElem.C:274-305)\n");
  LowMachKernelHex8Mesh fx;
  fx.fill_mesh_and_init_fields();
  fx.solnOpts_.includeDivU_ = 0.0;
  HelperObjects helperObjs(fx.bulk_, stk::topology::HEX_8, 3, fx.partVec_[0]);
  helperObjs.build_system();
  unit_test_utils::calc_mass_flow_rate_scs(helperObjs, fx);
  std::unique_ptr<Kernel> kernel(new MomentumAdvDiffElemKernel<AlgTraitsHex8>(*fx.bulk_, fx.solnOpts_, fx.velocity_, fx.viscosity_,
                                                                              helperObjs.assembleElemSolverAlg->dataNeededByKernels_));
  helperObjs.assembleElemSolverAlg->activeKernels_.push_back(kernel.get());
  helperObjs.assembleElemSolverAlg->execute();
  std::vector<double> lhs, rhs;
  helperObjs.element_lhs_rhs(lhs, rhs);
  EXPECT_TRUE(lhs.size() == 24u * 24u && rhs.size() == 24u);
  unit_test_utils::expect_all_near(rhs, gold("momentum_advdiff", "hex8_golds::advection_diffusion::rhs"), 2e-15);
  unit_test_utils::expect_all_near(lhs, gold("momentum_advdiff", "hex8_golds::advection_diffusion::lhs"), 2e-15);
}

static void test_continuity_advection(const char *variant)
{
  std::printf("ContinuityKernelHex8Mesh.%s (unit_tests/kernels/UnitTestContinuityAdvElem.C:59-170)\n", variant);
  LowMachKernelHex8Mesh fx;
  fx.fill_mesh_and_init_fields();
  const std::string v(variant);
  fx.solnOpts_.cvfemShiftMdot_ = false;
  fx.solnOpts_.shiftedGradOpMap_["pressure"] = v == "advection_reduced_shift_cvfem_poisson";
  fx.solnOpts_.cvfemReducedSensPoisson_ = v == "advection_reduced_sens_cvfem_poisson";
  HelperObjects helperObjs(fx.bulk_, stk::topology::HEX_8, 1, fx.partVec_[0]);
  helperObjs.build_system();
  helperObjs.realm.timeIntegrator_->timeStepN_ = 1.0;
  helperObjs.realm.timeIntegrator_->gamma1_ = 1.0;
  std::unique_ptr<Kernel> kernel(
      new ContinuityAdvElemKernel<AlgTraitsHex8>(*fx.bulk_, fx.solnOpts_, helperObjs.assembleElemSolverAlg->dataNeededByKernels_));
  helperObjs.assembleElemSolverAlg->activeKernels_.push_back(kernel.get());
  helperObjs.assembleElemSolverAlg->execute();
  std::vector<double> lhs, rhs;
  helperObjs.element_lhs_rhs(lhs, rhs);
  const std::string ns = v == "advection_default" ? "advection_default" : "advection_reduced_sensitivities";
  unit_test_utils::expect_all_near(rhs, gold("continuity_adv", "hex8_golds::" + ns + "::rhs"), 2e-15);
  unit_test_utils::expect_all_near(lhs, gold("continuity_adv", "hex8_golds::" + ns + "::lhs"), 2e-15);
}

static void test_momentum_mass(bool lumped)
{
  std::printf("MomentumKernelHex8Mesh.%s (unit_tests/kernels/UnitTestMomentumMassElem.C:144-215)\n",
              lumped ? "momentum_time_derivative_lumped" : "momentum_time_derivative");
  LowMachKernelHex8Mesh fx;
  fx.fill_mesh_and_init_fields();
  HelperObjects helperObjs(fx.bulk_, stk::topology::HEX_8, 3, fx.partVec_[0]);
  helperObjs.build_system();
  TimeIntegrator &ti = *helperObjs.realm.timeIntegrator_;
  ti.timeStepN_ = 0.1; ti.timeStepNm1_ = 0.1; ti.gamma1_ = 1.0; ti.gamma2_ = -1.0; ti.gamma3_ = 0.0;
  std::unique_ptr<Kernel> kernel(new MomentumMassElemKernel<AlgTraitsHex8>(*fx.bulk_, fx.solnOpts_,
                                                                           helperObjs.assembleElemSolverAlg->dataNeededByKernels_, lumped));
  helperObjs.assembleElemSolverAlg->activeKernels_.push_back(kernel.get());
  helperObjs.assembleElemSolverAlg->execute();
  std::vector<double> lhs, rhs;
  helperObjs.element_lhs_rhs(lhs, rhs);
  if (lumped) {
    unit_test_utils::expect_all_near(rhs, gold("momentum_mass", "hex8_golds::momentum_time_derivative_lumped::rhs"), 1e-14);
    for (int i = 0; i < 24; ++i)
      for (int j = 0; j < 24; ++j) EXPECT_NEAR(lhs[i * 24 + j], i == j ? 1.25 : 0.0, 1e-14);
  } else {
    unit_test_utils::expect_all_near(rhs, gold("momentum_mass", "hex8_golds::momentum_time_derivative::rhs"), 1e-14);
    unit_test_utils::expect_all_near(lhs, gold("momentum_mass", "hex8_golds::momentum_time_derivative::lhs"), 1e-14);
  }
}

static void test_momentum_nso(bool alt)
{
  std::printf("MomentumKernelHex8Mesh.%s (unit_tests/kernels/UnitTestMomentumNSOElem.C:%s)\n", alt ? "nso_2nd_alt" : "nso_2nd",
              alt ? "361-404" : "318-359");
  LowMachKernelHex8Mesh fx;
  fx.fill_mesh_and_init_fields();
  // MomentumKernelHex8Mesh::fill_mesh_and_init_fields: dudx_test_function = TrigFieldFunction::dudx
  // (unit_tests/kernels/UnitTestKernelUtils.h:343-351, UnitTestKernelUtils.C:38-63)
  const double *x = fx.coordinates_->host_data();
  for (size_t n = 0; n < fx.bulk_->num_nodes(); ++n) {
    const double a_pi = unit_test_utils::a * unit_test_utils::pi, X = x[3 * n], Y = x[3 * n + 1];
    double *g = fx.dudx_->host_data() + 9 * n;
    for (int q = 0; q < 9; ++q) g[q] = 0.0;
    g[0] = a_pi * std::sin(a_pi * X) * std::sin(a_pi * Y);
    g[1] = -a_pi * std::cos(a_pi * X) * std::cos(a_pi * Y);
    g[3] = a_pi * std::cos(a_pi * X) * std::cos(a_pi * Y);
    g[4] = -a_pi * std::sin(a_pi * X) * std::sin(a_pi * Y);
  }
  fx.dudx_->modify_on_host();
  fx.solnOpts_.meshMotion_ = false;
  fx.solnOpts_.meshDeformation_ = false;
  fx.solnOpts_.externalMeshDeformation_ = false;
  fx.solnOpts_.includeDivU_ = 0.0;
  HelperObjects helperObjs(fx.bulk_, stk::topology::HEX_8, 3, fx.partVec_[0]);
  helperObjs.build_system();
  std::unique_ptr<Kernel> kernel(new MomentumNSOElemKernel<AlgTraitsHex8>(*fx.bulk_, fx.solnOpts_, fx.velocity_, fx.dudx_, fx.viscosity_, 0.0,
                                                                          alt ? 1.0 : 0.0,
                                                                          helperObjs.assembleElemSolverAlg->dataNeededByKernels_));
  helperObjs.assembleElemSolverAlg->activeKernels_.push_back(kernel.get());
  TimeIntegrator &ti = *helperObjs.realm.timeIntegrator_;
  ti.timeStepN_ = 0.1; ti.timeStepNm1_ = 0.1; ti.gamma1_ = 1.0; ti.gamma2_ = -1.0; ti.gamma3_ = 0.0;
  helperObjs.assembleElemSolverAlg->execute();
  std::vector<double> lhs, rhs;
  helperObjs.element_lhs_rhs(lhs, rhs);
  EXPECT_TRUE(rhs.size() == 24u && lhs.size() == 576u);
  const std::string ns = alt ? "hex8_golds::nso_2nd_alt" : "hex8_golds::nso_2nd";
  // The fixture's z velocity is zero: with the pde residual the z rows of the gold (1e-10 .. 1e-9) are the rounding noise of
  // a residual that cancels to zero, divided by sqrt(small_) = 1e-8 (nu = |residual| / sqrt(0 + 1e-16),
  // src/nso/MomentumNSOElemKernel.C:262-270): another summation order gets other noise of the same size.  x and y rows
  // at the reference's tolerance, z rows at 2e-9.
  const Table gr = gold("momentum_nso", ns + "::rhs"), gl = gold("momentum_nso", ns + "::lhs");
  for (int i = 0; i < 24; ++i) {
    const double tol = i % 3 == 2 ? 2e-9 : 1e-15;
    EXPECT_NEAR(rhs[i], gr.v[i], tol);
    for (int j = 0; j < 24; ++j) EXPECT_NEAR(lhs[i * 24 + j], gl.v[i * 24 + j], tol);
  }
}

static void test_momentum_buoyancy(bool boussinesq)
{
  std::printf("MomentumKernelHex8Mesh.%s (unit_tests/kernels/UnitTestMomentumBuoyancy%sSrcElem.C)\n",
              boussinesq ? "buoyancy_boussinesq" : "buoyancy", boussinesq ? "Boussinesq" : "");
  std::mt19937 rng;
  rng.seed(0); // fixed seed
  std::uniform_real_distribution<double> ref_densities(0.8, 1.3);
  LowMachKernelHex8Mesh fx;
  fx.fill_mesh_and_init_fields();
  fx.solnOpts_.gravity_[2] = -9.81;
  fx.solnOpts_.referenceDensity_ = ref_densities(rng);
  fx.solnOpts_.referenceTemperature_ = 298;
  fx.solnOpts_.thermalExpansionCoeff_ = 1.0;
  HelperObjects helperObjs(fx.bulk_, stk::topology::HEX_8, 3, fx.partVec_[0]);
  helperObjs.build_system();
  std::unique_ptr<Kernel> kernel;
  if (boussinesq)
    kernel.reset(new MomentumBuoyancyBoussinesqSrcElemKernel<AlgTraitsHex8>(*fx.bulk_, fx.solnOpts_,
                                                                            helperObjs.assembleElemSolverAlg->dataNeededByKernels_));
  else
    kernel.reset(new MomentumBuoyancySrcElemKernel<AlgTraitsHex8>(*fx.bulk_, fx.solnOpts_,
                                                                  helperObjs.assembleElemSolverAlg->dataNeededByKernels_));
  helperObjs.assembleElemSolverAlg->activeKernels_.push_back(kernel.get());
  helperObjs.assembleElemSolverAlg->execute();
  std::vector<double> lhs, rhs;
  helperObjs.element_lhs_rhs(lhs, rhs);
  EXPECT_TRUE(rhs.size() == 24u && lhs.size() == 576u);
  std::vector<double> rhsExact(24, 0.0);
  for (size_t i = 2; i < 24; i += 3)
    rhsExact[i] = boussinesq ? -0.125 * fx.solnOpts_.gravity_[2] * fx.solnOpts_.referenceDensity_ * 1.0 * (300.0 - 298.0)
                             : 0.125 * fx.solnOpts_.gravity_[2] * (1.0 - fx.solnOpts_.referenceDensity_);
  Table exact;
  exact.dims = {24};
  exact.v = rhsExact;
  unit_test_utils::expect_all_near(rhs, exact, 4e-15);
  for (double v : lhs) EXPECT_TRUE(v == 0.0);
}

static void test_sst_src(int which)
{
  const bool sdr = which == 1, des = which == 2;
  std::printf("SSTKernelHex8Mesh.%s (unit_tests/kernels/UnitTestSSTSrcElem.C:%s)\n",
              sdr ? "specificdissipationratesstsrcelem" : des ? "turbkineticenergysstdessrcelem" : "turbkineticenergysstsrcelem",
              sdr ? "259-297" : des ? "219-257" : "179-217");
  LowMachKernelHex8Mesh fx;
  fx.fill_mesh_and_init_fields();
  fx.fill_sst_fields();
  fx.solnOpts_.initialize_turbulence_constants();
  HelperObjects helperObjs(fx.bulk_, stk::topology::HEX_8, 1, fx.partVec_[0]);
  helperObjs.build_system();
  std::unique_ptr<Kernel> kernel;
  if (sdr)
    kernel.reset(new SpecificDissipationRateSSTSrcElemKernel<AlgTraitsHex8>(*fx.bulk_, fx.solnOpts_,
                                                                            helperObjs.assembleElemSolverAlg->dataNeededByKernels_, false));
  else if (des)
    kernel.reset(new TurbKineticEnergySSTDESSrcElemKernel<AlgTraitsHex8>(*fx.bulk_, fx.solnOpts_,
                                                                         helperObjs.assembleElemSolverAlg->dataNeededByKernels_, false));
  else
    kernel.reset(new TurbKineticEnergySSTSrcElemKernel<AlgTraitsHex8>(*fx.bulk_, fx.solnOpts_,
                                                                      helperObjs.assembleElemSolverAlg->dataNeededByKernels_, false));
  helperObjs.assembleElemSolverAlg->activeKernels_.push_back(kernel.get());
  helperObjs.assembleElemSolverAlg->execute();
  std::vector<double> lhs, rhs;
  helperObjs.element_lhs_rhs(lhs, rhs);
  EXPECT_TRUE(rhs.size() == 8u && lhs.size() == 64u);
  const std::string ns = sdr ? "SpecificDissipationRateSSTSrcElemKernel"
                             : des ? "TurbKineticEnergySSTDESSrcElemKernel" : "TurbKineticEnergySSTSrcElemKernel";
  unit_test_utils::expect_all_near(rhs, gold("sst_src", "hex8_golds::" + ns + "::rhs"), 2e-15);
  unit_test_utils::expect_all_near(lhs, gold("sst_src", "hex8_golds::" + ns + "::lhs"), 2e-15);
}

// The reference has no unit test for the upwind kernels.  With the SolutionOptions defaults (alpha = 0, hybrid factor
// 0 -> classic Peclet function = 0, src/SolutionOptions.C:32-44) every upwind term vanishes, so the kernel must
// reproduce the golden table of MomentumAdvDiffElemKernel whatever the limiter / upwind settings are.
static void test_momentum_upw_defaults_reproduce_central_gold()
{
  std::printf("MomentumUpwAdvDiffElemKernel, default blending == MomentumKernelHex8Mesh.advection_diffusion gold\n");
  LowMachKernelHex8Mesh fx;
  fx.fill_mesh_and_init_fields();
  fx.solnOpts_.includeDivU_ = 0.0;
  fx.solnOpts_.limiterMap_["velocity"] = true;
  for (size_t n = 0; n < 9 * fx.bulk_->num_nodes(); ++n) fx.dudx_->host_data()[n] = 0.1 * std::sin(1.0 + n);
  fx.dudx_->modify_on_host();
  HelperObjects helperObjs(fx.bulk_, stk::topology::HEX_8, 3, fx.partVec_[0]);
  helperObjs.build_system();
  unit_test_utils::calc_mass_flow_rate_scs(helperObjs, fx);
  std::unique_ptr<Kernel> kernel(new MomentumUpwAdvDiffElemKernel<AlgTraitsHex8>(
      *fx.bulk_, fx.solnOpts_, &helperObjs.eqSystem, fx.velocity_, fx.viscosity_, fx.dudx_,
      helperObjs.assembleElemSolverAlg->dataNeededByKernels_));
  helperObjs.assembleElemSolverAlg->activeKernels_.push_back(kernel.get());
  helperObjs.assembleElemSolverAlg->execute();
  std::vector<double> lhs, rhs;
  helperObjs.element_lhs_rhs(lhs, rhs);
  EXPECT_TRUE(lhs.size() == 24u * 24u && rhs.size() == 24u);
  unit_test_utils::expect_all_near(rhs, gold("momentum_advdiff", "hex8_golds::advection_diffusion::rhs"), 2e-15);
  unit_test_utils::expect_all_near(lhs, gold("momentum_advdiff", "hex8_golds::advection_diffusion::lhs"), 2e-15);
}

// upwind blending switched on (hybrid factor 1, limiter): structure the reference's loop guarantees -- every ip adds
// +v to the left row and -v to the right row, so lhs columns and the rhs sum to zero over the element; and
// setup() must pick up option changes made after construction (ScalarUpwAdvDiffElemKernel.C:96-108)
static void test_scalar_upw_blending()
{
  std::printf("ScalarUpwAdvDiffElemKernel, upwind blending: conservation structure, setup() re-reads the options\n");
  LowMachKernelHex8Mesh fx;
  fx.fill_mesh_and_init_fields(true);
  const double *x = fx.coordinates_->host_data();
  for (size_t n = 0; n < fx.bulk_->num_nodes(); ++n) {
    fx.scalarQ_->host_data()[n] = std::cos(x[3 * n]) * std::sin(0.5 + x[3 * n + 1]) + 0.2 * x[3 * n + 2];
    fx.diffFluxCoeff_->host_data()[n] = 1.0e-3;
    for (int j = 0; j < 3; ++j) fx.dqdx_->host_data()[3 * n + j] = 0.3 * std::cos(1.0 + j + x[3 * n + j]);
  }
  for (stk::mesh::FieldBase *f : {(stk::mesh::FieldBase *)fx.scalarQ_, (stk::mesh::FieldBase *)fx.diffFluxCoeff_, (stk::mesh::FieldBase *)fx.dqdx_})
    f->modify_on_host();
  HelperObjects helperObjs(fx.bulk_, stk::topology::HEX_8, 1, fx.partVec_[0]);
  helperObjs.realm.solutionOptions_->hybridMap_["scalarQ"] = 1.0;
  fx.solnOpts_.limiterMap_["scalarQ"] = true;
  helperObjs.build_system();
  unit_test_utils::calc_mass_flow_rate_scs(helperObjs, fx);
  std::unique_ptr<Kernel> kernel(new ScalarUpwAdvDiffElemKernel<AlgTraitsHex8>(
      *fx.bulk_, fx.solnOpts_, &helperObjs.eqSystem, fx.scalarQ_, fx.dqdx_, fx.diffFluxCoeff_,
      helperObjs.assembleElemSolverAlg->dataNeededByKernels_));
  helperObjs.assembleElemSolverAlg->activeKernels_.push_back(kernel.get());
  helperObjs.assembleElemSolverAlg->execute();
  std::vector<double> lhs, rhs;
  helperObjs.element_lhs_rhs(lhs, rhs);
  EXPECT_TRUE(lhs.size() == 64u && rhs.size() == 8u);
  double scale = 0, rs = 0;
  for (double v : lhs) scale = std::fmax(scale, std::fabs(v));
  for (int c = 0; c < 8; ++c) {
    double s = 0;
    for (int r = 0; r < 8; ++r) s += lhs[r * 8 + c];
    EXPECT_TRUE(std::fabs(s) <= 1e-14 * scale);
  }
  for (double v : rhs) rs += v;
  EXPECT_TRUE(std::fabs(rs) <= 1e-14);
  // the upwind operator is not the central one
  nalu_b200_kernel_desc d;
  kernel->device_descriptor(d);
  EXPECT_TRUE(d.kind == NALU_K_SCALAR_UPW_ADV_DIFF && (d.options & NALU_OPT_USE_LIMITER) && d.upw[NALU_UPW_HYBRID_FACTOR] == 1.0);
  fx.solnOpts_.alphaUpwMap_["scalarQ"] = 0.5;
  kernel->setup(*helperObjs.realm.timeIntegrator_);
  kernel->device_descriptor(d);
  EXPECT_TRUE(d.upw[NALU_UPW_ALPHA_UPW] == 0.5);
}

// LinearSystem::applyDirichletBCs / resetRows (src/TpetraLinearSystem.C:1579-1653,1706-1755) after an assembly:
// identity rows + (bc - solution) residuals on a node-set part, zero rows on a node list, every other row untouched
static void test_dirichlet_and_reset_rows()
{
  std::printf("DeviceCsrLinearSystem::applyDirichletBCs / resetRows\n");
  LowMachKernelHex8Mesh fx(3, 3, 3);
  fx.fill_mesh_and_init_fields(true);
  fx.solnOpts_.includeDivU_ = 1.0;
  HelperObjects h(fx.bulk_, stk::topology::HEX_8, 3, fx.partVec_[0]);
  h.build_system();
  unit_test_utils::calc_mass_flow_rate_scs(h, fx);
  std::unique_ptr<Kernel> kernel(new MomentumAdvDiffElemKernel<AlgTraitsHex8>(*fx.bulk_, fx.solnOpts_, fx.velocity_, fx.viscosity_,
                                                                              h.assembleElemSolverAlg->dataNeededByKernels_));
  h.assembleElemSolverAlg->activeKernels_.push_back(kernel.get());
  h.assembleElemSolverAlg->execute();
  const std::vector<double> v0 = h.linsys->ownedValues(), r0 = h.linsys->ownedRhs();
  const std::vector<uint64_t> rp = h.linsys->ownedRowPtr();
  const std::vector<int32_t> ci = h.linsys->ownedColIdx();
  // a wall: the nodes of the first node plane (ids 1..16 of generated:3x3x3), given as a node-set part
  stk::mesh::Part &wall = fx.meta_.declare_part_with_topology("surface_1", stk::topology::NODE);
  std::vector<int32_t> wallNodes;
  for (int n = 0; n < 16; ++n) wallNodes.push_back(n);
  wall.set_nodes(wallNodes);
  VectorFieldType &bc = fx.meta_.declare_field<double>(stk::topology::NODE_RANK, "velocity_bc");
  stk::mesh::put_field_on_mesh(bc, *fx.part_, 3);
  bc.set_num_entities(fx.bulk_->num_nodes());
  for (size_t n = 0; n < 3 * fx.bulk_->num_nodes(); ++n) bc.host_data()[n] = 0.25 + 0.01 * n;
  bc.modify_on_host();
  h.linsys->applyDirichletBCs(fx.velocity_, &bc, {&wall}, 0, 3);
  std::vector<stk::mesh::Entity> resetList;
  for (int n : {20, 37, 63}) {
    stk::mesh::Entity e;
    e.rank = stk::topology::NODE_RANK;
    e.index = n;
    resetList.push_back(e);
  }
  h.linsys->resetRows(resetList, 1, 2);
  const std::vector<double> v1 = h.linsys->ownedValues(), r1 = h.linsys->ownedRhs();
  const double *u = fx.velocity_->host_data();
  int bad = 0, ones = 0;
  for (int node = 0; node < int(fx.bulk_->num_nodes()); ++node) {
    const int lid0 = h.linsys->entityToLID({stk::topology::NODE_RANK, node});
    for (int d = 0; d < 3; ++d) {
      const int row = lid0 + d;
      const bool dir = node < 16, rst = (node == 20 || node == 37 || node == 63) && d == 1;
      for (uint64_t j = rp[row]; j < rp[row + 1]; ++j) {
        const double want = dir ? (ci[j] == row ? 1.0 : 0.0) : (rst ? 0.0 : v0[j]);
        if (v1[j] != want) ++bad;
        if (dir && ci[j] == row) ++ones;
      }
      const double rw = dir ? bc.host_data()[3 * node + d] - u[3 * node + d] : (rst ? 0.0 : r0[row]);
      if (r1[row] != rw) ++bad;
    }
  }
  EXPECT_TRUE(bad == 0);
  EXPECT_TRUE(ones == 48);
  EXPECT_THROW(h.linsys->applyDirichletBCs(fx.velocity_, &bc, {&wall}, 0, 4));
}

static void test_scalar_diff()
{
  std::printf("HeatCondKernelHex8Mesh.cvfem_diff (unit_tests/kernels/UnitTestScalarDiffElem.C:30-58)\n");
  LowMachKernelHex8Mesh fx;
  fx.fill_mesh_and_init_fields();
  for (size_t n = 0; n < fx.bulk_->num_nodes(); ++n) {
    fx.scalarQ_->host_data()[n] = std::cos(fx.coordinates_->host_data()[3 * n]);
    fx.diffFluxCoeff_->host_data()[n] = 1.0;
  }
  fx.scalarQ_->modify_on_host();
  fx.diffFluxCoeff_->modify_on_host();
  HelperObjects helperObjs(fx.bulk_, stk::topology::HEX_8, 1, fx.partVec_[0]);
  helperObjs.build_system();
  std::unique_ptr<Kernel> kernel(new ScalarDiffElemKernel<AlgTraitsHex8>(*fx.bulk_, fx.solnOpts_, fx.scalarQ_, fx.diffFluxCoeff_,
                                                                         helperObjs.assembleElemSolverAlg->dataNeededByKernels_));
  helperObjs.assembleElemSolverAlg->activeKernels_.push_back(kernel.get());
  helperObjs.assembleElemSolverAlg->execute();
  std::vector<double> lhs, rhs;
  helperObjs.element_lhs_rhs(lhs, rhs);
  unit_test_utils::expect_all_near(lhs, gold("scalar_diff", "hex8_golds::heatcond::cvfem_diff::lhs"), 2e-15);
}

// unit_tests/UnitTestTpetra.C:86-203: 1x1x2 mesh, a TestKernel writing a fixed 8x8 block, row lengths and the
// assembled values looked up by global ids; plus sumInto for a host caller and the two-kernel accumulation order
static void test_tpetra_basic()
{
  std::printf("Tpetra.basic, np=1 (unit_tests/UnitTestTpetra.C:86-203)\n");
  struct TestKernel : public Kernel {
    const double *block;
    explicit TestKernel(const double *b) : block(b) {}
    bool device_descriptor(nalu_b200_kernel_desc &d) override
    {
      clear(d);
      d.kind = NALU_K_TEST_FIXED_BLOCK;
      d.fixed_block = block;
      return true;
    }
  };
  LowMachKernelHex8Mesh fx(1, 1, 2);
  HelperObjects helperObjs(fx.bulk_, stk::topology::HEX_8, 1, fx.partVec_[0]);
  helperObjs.build_system();
  EXPECT_TRUE(helperObjs.linsys->numOwnedRows() == 12);
  EXPECT_TRUE(helperObjs.linsys->numSharedNotOwnedRows() == 0);
  const auto rp = helperObjs.linsys->ownedRowPtr();
  const int lens[12] = {8, 8, 8, 8, 12, 12, 12, 12, 8, 8, 8, 8};
  for (int r = 0; r < 12; ++r) EXPECT_TRUE(int(rp[r + 1] - rp[r]) == lens[r]);
  const Table elemVals = gold("tpetra_2elem", "elemVals"), lhsVals = gold("tpetra_2elem", "lhsVals");
  void *dblk = nullptr;
  nalu_b200_malloc(&dblk, 64 * 8);
  nalu_b200_memcpy_h2d(dblk, elemVals.v.data(), 64 * 8, nullptr);
  TestKernel *tk = new TestKernel((const double *)dblk);
  helperObjs.assembleElemSolverAlg->activeKernels_.push_back(tk);
  helperObjs.assembleElemSolverAlg->execute();
  helperObjs.linsys->loadComplete();
  const auto ci = helperObjs.linsys->ownedColIdx();
  const auto v = helperObjs.linsys->ownedValues();
  const auto rg = helperObjs.linsys->ownedRowGids();
  const auto cg = helperObjs.linsys->colGids();
  for (int r = 0; r < 12; ++r)
    for (uint64_t q = rp[r]; q < rp[r + 1]; ++q) EXPECT_NEAR(lhsVals.v[size_t(rg[r] - 1) * 12 + size_t(cg[ci[q]] - 1)], v[q], 1e-14);
  // sumInto from a host caller (one node, like AssembleNodeSolverAlgorithm does)
  stk::mesh::Entity node = fx.bulk_->get_entity(stk::topology::NODE_RANK, 5);
  std::vector<int> ids;
  std::vector<double> vals, rhs1 = {2.5}, lhs1 = {7.0};
  helperObjs.linsys->sumInto(std::vector<stk::mesh::Entity>{node}, ids, vals, rhs1, lhs1, "unit test");
  const auto v2 = helperObjs.linsys->ownedValues();
  const auto r2 = helperObjs.linsys->ownedRhs();
  const int lid = helperObjs.linsys->entityToLID(node);
  EXPECT_NEAR(r2[lid], 2.5, 1e-14);
  for (uint64_t q = rp[lid]; q < rp[lid + 1]; ++q) EXPECT_NEAR(v2[q] - v[q], ci[q] == lid ? 7.0 : 0.0, 1e-14);
  delete tk;
  nalu_b200_free(dblk);
}

// UnitTestSimdMomentum-style smoke on generated:20x20x20 with two kernels on one algorithm: the fused launch must
// equal the sum of the two kernels run alone (accumulation into one zeroed lhs/rhs,
// src/AssembleElemSolverAlgorithm.C:86-91), and the big mesh takes the tiled kernel when it can
static void test_two_kernels_one_algorithm()
{
  std::printf("Hex8Mesh 20x20x20: mass + advection_diffusion on one algorithm (unit_tests/UnitTestSimdMomentum.C:16-63)\n");
  auto run = [](bool mass, bool advdiff, std::vector<double> &values, std::vector<double> &rhs, int &family) {
    LowMachKernelHex8Mesh fx(20, 20, 20);
    fx.fill_mesh_and_init_fields(true);
    fx.solnOpts_.includeDivU_ = 1.0;
    HelperObjects h(fx.bulk_, stk::topology::HEX_8, 3, fx.partVec_[0]);
    h.build_system();
    h.realm.timeIntegrator_->timeStepN_ = 0.05;
    unit_test_utils::calc_mass_flow_rate_scs(h, fx);
    std::vector<std::unique_ptr<Kernel>> own;
    auto &req = h.assembleElemSolverAlg->dataNeededByKernels_;
    if (mass) own.emplace_back(new MomentumMassElemKernel<AlgTraitsHex8>(*fx.bulk_, fx.solnOpts_, req, true));
    if (advdiff) own.emplace_back(new MomentumAdvDiffElemKernel<AlgTraitsHex8>(*fx.bulk_, fx.solnOpts_, fx.velocity_, fx.viscosity_, req));
    for (auto &k : own) h.assembleElemSolverAlg->activeKernels_.push_back(k.get());
    h.assembleElemSolverAlg->execute();
    values = h.linsys->ownedValues();
    rhs = h.linsys->ownedRhs();
    family = h.linsys->lastAssembleKernelFamily();
  };
  std::vector<double> vb, rb, vm, rm, va, ra;
  int fb, fm, fa;
  run(true, true, vb, rb, fb);
  run(true, false, vm, rm, fm);
  run(false, true, va, ra, fa);
  EXPECT_TRUE(fa == 2); // advection-diffusion alone on a generated box: tiled write-once kernel
  EXPECT_TRUE(fb == 2 && fm == 1); // lumped mass + adv-diff: ONE launch of the row-thread tiled kernel (fused list); mass alone: general kernel
  double scale = 0, err = 0;
  for (size_t i = 0; i < vb.size(); ++i) {
    scale = std::fmax(scale, std::fabs(vb[i]));
    err = std::fmax(err, std::fabs(vb[i] - vm[i] - va[i]));
  }
  EXPECT_TRUE(err <= 1e-12 * scale);
  double rs = 0, re = 0;
  for (size_t i = 0; i < rb.size(); ++i) {
    rs = std::fmax(rs, std::fabs(rb[i]));
    re = std::fmax(re, std::fabs(rb[i] - rm[i] - ra[i]));
  }
  EXPECT_TRUE(re <= 1e-12 * rs);
}

// ---------------------------------------------------------------------------------------------------
// Hex27 (BASELINE config 5): master-element tables on the host, the momentum list through the registration API
// ---------------------------------------------------------------------------------------------------
static void test_hex27_master_elements()
{
  std::printf("MasterElementHexSerial.hex27_* (unit_tests/UnitTestHexMasterElements.C:297-321) on the host tables\n");
  MasterElement *scs = MasterElementRepo::get_surface_master_element(stk::topology::HEX_27);
  MasterElement *scv = MasterElementRepo::get_volume_master_element(stk::topology::HEX_27);
  EXPECT_TRUE(scs->numIntPoints_ == AlgTraitsHex27::numScsIp_ && scv->numIntPoints_ == AlgTraitsHex27::numScvIp_);
  EXPECT_TRUE(scs->nodesPerElement_ == 27 && scv->nodesPerElement_ == 27);
  std::vector<double> sf(216 * 27), sfs(216 * 27);
  scs->shape_fcn(sf.data());
  const int *lr = scs->adjacentNodes();
  // reference element (unit_tests/UnitTestUtils.C:218-233): node coordinate along x = its position in the tensor line; a
  // linear field is reproduced at every point, and every point separates two different nodes
  static const double X[27] = {-1, 1, 1, -1, -1, 1, 1, -1, 0, 1, 0, -1, -1, 1, 1, -1, 0, 1, 0, -1, 0, 0, 0, -1, 1, 0, 0};
  for (int ip = 0; ip < 216; ++ip) {
    double sum = 0.0, x = 0.0;
    for (int n = 0; n < 27; ++n) { sum += sf[ip * 27 + n]; x += sf[ip * 27 + n] * X[n]; }
    EXPECT_NEAR(sum, 1.0, 1e-14);
    EXPECT_TRUE(std::fabs(x) <= 1.0);
    EXPECT_TRUE(lr[2 * ip] != lr[2 * ip + 1] && lr[2 * ip] >= 0 && lr[2 * ip] < 27 && lr[2 * ip + 1] >= 0 && lr[2 * ip + 1] < 27);
  }
  // s-direction points (the last 72) sit at x = -+1/sqrt(3) and point from the low-x to the high-x node
  for (int ip = 144; ip < 216; ++ip) {
    double x = 0.0;
    for (int n = 0; n < 27; ++n) x += sf[ip * 27 + n] * X[n];
    EXPECT_NEAR(std::fabs(x), 1.0 / std::sqrt(3.0), 1e-14);
    EXPECT_TRUE(x < 0 ? X[lr[2 * ip]] < X[lr[2 * ip + 1]] : X[lr[2 * ip]] > X[lr[2 * ip + 1]]); // left node on the outer side
  }
  scv->shifted_shape_fcn(sfs.data());
  const int *nn = scv->ipNodeMap();
  int count[27] = {0};
  for (int ip = 0; ip < 216; ++ip) {
    ++count[nn[ip]];
    for (int n = 0; n < 27; ++n) EXPECT_NEAR(sfs[ip * 27 + n], n == nn[ip] ? 1.0 : 0.0, 1e-15); // lumping: nodal
  }
  for (int n = 0; n < 27; ++n) EXPECT_TRUE(count[n] == 8);
  // topology dispatch of the kernel builder: the momentum list exists for HEX_27, other kernels say so
  stk::mesh::MetaData meta(3);
  bool threw = false;
  try {
    std::vector<Kernel *> v;
    ElemDataRequests req;
    std::shared_ptr<stk::mesh::BulkData> bulk = std::make_shared<stk::mesh::BulkData>(meta);
    SolutionOptions so;
    build_topo_kernel<MomentumBuoyancySrcElemKernel>(stk::topology::HEX_27, *bulk, so, req);
  } catch (const std::exception &e) {
    threw = std::string(e.what()).find("HEXAHEDRON_27") != std::string::npos;
  }
  EXPECT_TRUE(threw);
}

static void test_hex27_momentum_list()
{
  std::printf("Hex27 element: [MomentumMass(lumped), MomentumAdvDiff] through build_topo_kernel / AssembleElemSolverAlgorithm "
              "(gold: tests/golden/make_hex27_host_gold.py)\n");
  stk::mesh::MetaData meta(3);
  auto bulk = std::make_shared<stk::mesh::BulkData>(meta);
  stk::mesh::Part *part = &meta.declare_part_with_topology("block_1", stk::topology::HEX_27);
  VectorFieldType *coords = &meta.declare_field<double>(stk::topology::NODE_RANK, "coordinates");
  VectorFieldType *velocity = &meta.declare_field<double>(stk::topology::NODE_RANK, "velocity", 3);
  VectorFieldType *dpdx = &meta.declare_field<double>(stk::topology::NODE_RANK, "dpdx");
  ScalarFieldType *density = &meta.declare_field<double>(stk::topology::NODE_RANK, "density", 3);
  ScalarFieldType *viscosity = &meta.declare_field<double>(stk::topology::NODE_RANK, "viscosity");
  GenericFieldType *mdot = &meta.declare_field<double>(stk::topology::ELEM_RANK, "mass_flow_rate_scs");
  stk::mesh::put_field_on_mesh(*coords, *part, 3);
  stk::mesh::put_field_on_mesh(*velocity, *part, 3);
  stk::mesh::put_field_on_mesh(*dpdx, *part, 3);
  stk::mesh::put_field_on_mesh(*density, *part, 1);
  stk::mesh::put_field_on_mesh(*viscosity, *part, 1);
  stk::mesh::put_field_on_mesh(*mdot, *part, AlgTraitsHex27::numScsIp_);
  bulk->fill_generated_hex27_mesh(1, 1, 1, *part);
  EXPECT_TRUE(bulk->num_nodes() == 27u && bulk->num_elems() == 1u && bulk->nodes_per_elem() == 27);
  // closed-form mesh and fields of the gold script
  auto &uN = velocity->field_of_state(stk::mesh::StateN), &uNm1 = velocity->field_of_state(stk::mesh::StateNM1);
  auto &rN = density->field_of_state(stk::mesh::StateN), &rNm1 = density->field_of_state(stk::mesh::StateNM1);
  for (size_t n = 0; n < bulk->num_nodes(); ++n) {
    double *x = coords->host_data() + 3 * n;
    const double X0 = x[0], Y0 = x[1], Z0 = x[2];
    x[0] = X0 + 0.04 * std::sin(3.1 * Y0 + 1.7 * Z0 + 0.3);
    x[1] = Y0 + 0.04 * std::cos(2.3 * X0 - 1.1 * Z0);
    x[2] = Z0 + 0.04 * std::sin(1.9 * X0 + 2.9 * Y0 + 0.5);
    const double X = x[0], Y = x[1], Z = x[2];
    const double u[3] = {std::sin(0.9 * Y) + 0.3 * Z, std::cos(0.7 * X) * Z, 0.2 * X * Y - 0.5};
    for (int j = 0; j < 3; ++j) {
      velocity->host_data()[3 * n + j] = u[j];
      uN.host_data()[3 * n + j] = 0.9 * u[j] + 0.01;
      uNm1.host_data()[3 * n + j] = 0.8 * u[j] - 0.02;
    }
    viscosity->host_data()[n] = 0.3 + 0.1 * std::cos(X + 2.0 * Y);
    const double rho = 1.0 + 0.2 * std::sin(Z);
    density->host_data()[n] = rho;
    rN.host_data()[n] = 0.98 * rho;
    rNm1.host_data()[n] = 0.95 * rho;
    dpdx->host_data()[3 * n + 0] = 0.1 * X;
    dpdx->host_data()[3 * n + 1] = -0.2 * Y;
    dpdx->host_data()[3 * n + 2] = 0.05 * Z;
  }
  for (stk::mesh::FieldBase *f : std::initializer_list<stk::mesh::FieldBase *>{coords, velocity, &uN, &uNm1, viscosity, density, &rN, &rNm1, dpdx})
    f->modify_on_host();
  SolutionOptions solnOpts;
  solnOpts.includeDivU_ = 1.0;
  HelperObjects h(bulk, stk::topology::HEX_27, 3, part);
  h.build_system();
  TimeIntegrator &ti = *h.realm.timeIntegrator_;
  ti.timeStepN_ = 0.05; ti.timeStepNm1_ = 0.05; ti.gamma1_ = 1.5; ti.gamma2_ = -2.0; ti.gamma3_ = 0.5;
  // mass_flow_rate_scs on the device (216 per element), checked against the gold before it is used
  {
    void *d = nullptr;
    const size_t n = bulk->num_elems() * 216;
    if (nalu_b200_malloc(&d, n * 8) != 0) throw std::runtime_error("malloc");
    if (nalu_b200_mdot_scs(h.linsys->graph(), coords->device_data(), density->device_data(), velocity->device_data(), (double *)d, nullptr) != 0)
      throw std::runtime_error(nalu_b200_last_error());
    nalu_b200_memcpy_d2h(mdot->host_data(), d, n * 8, nullptr);
    nalu_b200_stream_sync(nullptr);
    nalu_b200_free(d);
    mdot->modify_on_host();
    const Table g = gold("hex27_host", "mdot");
    for (int ip = 0; ip < 216; ++ip) EXPECT_NEAR(mdot->host_data()[ip], g.v[ip], 1e-14);
  }
  auto &req = h.assembleElemSolverAlg->dataNeededByKernels_;
  std::unique_ptr<Kernel> mass(build_topo_kernel<MomentumMassElemKernel>(stk::topology::HEX_27, *bulk, solnOpts, req, true));
  std::unique_ptr<Kernel> advdiff(build_topo_kernel<MomentumAdvDiffElemKernel>(stk::topology::HEX_27, *bulk, solnOpts, velocity, viscosity, req));
  h.assembleElemSolverAlg->activeKernels_.push_back(mass.get());
  h.assembleElemSolverAlg->activeKernels_.push_back(advdiff.get());
  h.assembleElemSolverAlg->execute();
  EXPECT_TRUE(h.linsys->lastAssembleKernelFamily() == 5);
  // element lhs / rhs in (STK node, dof) order from the one-element CSR
  const int n = 81;
  const auto rp = h.linsys->ownedRowPtr();
  const auto ci = h.linsys->ownedColIdx();
  const auto v = h.linsys->ownedValues();
  const auto r = h.linsys->ownedRhs();
  const int32_t *en = bulk->begin_nodes(stk::mesh::Entity{stk::topology::ELEM_RANK, 0});
  std::vector<int> rows;
  for (int k = 0; k < 27; ++k)
    for (int d = 0; d < 3; ++d) rows.push_back(h.linsys->entityToLID(stk::mesh::Entity{stk::topology::NODE_RANK, en[k]}) + d);
  std::vector<int> pos(n, -1);
  for (int j = 0; j < n; ++j) pos[rows[j]] = j;
  std::vector<double> lhs(size_t(n) * n, 0.0), rhs(n);
  for (int i = 0; i < n; ++i) {
    rhs[i] = r[rows[i]];
    EXPECT_TRUE(rp[rows[i] + 1] - rp[rows[i]] == uint64_t(n));
    for (uint64_t q = rp[rows[i]]; q < rp[rows[i] + 1]; ++q) lhs[size_t(i) * n + pos[ci[q]]] = v[q];
  }
  unit_test_utils::expect_all_near(lhs, gold("hex27_host", "momentum_list::lhs"), 2e-13);
  unit_test_utils::expect_all_near(rhs, gold("hex27_host", "momentum_list::rhs"), 2e-13);
}

static void test_hex27_continuity()
{
  std::printf("Hex27 element: ContinuityAdvElemKernel through build_topo_kernel / AssembleElemSolverAlgorithm (oracle gold)\n");
  stk::mesh::MetaData meta(3);
  auto bulk = std::make_shared<stk::mesh::BulkData>(meta);
  stk::mesh::Part *part = &meta.declare_part_with_topology("block_1", stk::topology::HEX_27);
  VectorFieldType *coords = &meta.declare_field<double>(stk::topology::NODE_RANK, "coordinates");
  VectorFieldType *velocity = &meta.declare_field<double>(stk::topology::NODE_RANK, "velocity", 2);
  VectorFieldType *dpdx = &meta.declare_field<double>(stk::topology::NODE_RANK, "dpdx");
  ScalarFieldType *density = &meta.declare_field<double>(stk::topology::NODE_RANK, "density", 2);
  ScalarFieldType *pressure = &meta.declare_field<double>(stk::topology::NODE_RANK, "pressure");
  stk::mesh::put_field_on_mesh(*coords, *part, 3);
  stk::mesh::put_field_on_mesh(*velocity, *part, 3);
  stk::mesh::put_field_on_mesh(*dpdx, *part, 3);
  stk::mesh::put_field_on_mesh(*density, *part, 1);
  stk::mesh::put_field_on_mesh(*pressure, *part, 1);
  bulk->fill_generated_hex27_mesh(1, 1, 1, *part);
  for (size_t n = 0; n < bulk->num_nodes(); ++n) {
    double *x = coords->host_data() + 3 * n;
    const double X0 = x[0], Y0 = x[1], Z0 = x[2];
    x[0] = X0 + 0.04 * std::sin(3.1 * Y0 + 1.7 * Z0 + 0.3);
    x[1] = Y0 + 0.04 * std::cos(2.3 * X0 - 1.1 * Z0);
    x[2] = Z0 + 0.04 * std::sin(1.9 * X0 + 2.9 * Y0 + 0.5);
    const double X = x[0], Y = x[1], Z = x[2];
    velocity->host_data()[3 * n + 0] = std::sin(0.9 * Y) + 0.3 * Z;
    velocity->host_data()[3 * n + 1] = std::cos(0.7 * X) * Z;
    velocity->host_data()[3 * n + 2] = 0.2 * X * Y - 0.5;
    density->host_data()[n] = 1.0 + 0.2 * std::sin(Z);
    pressure->host_data()[n] = 0.5 * std::cos(0.8 * X) * std::sin(0.6 * Y) + 0.1 * Z;
    dpdx->host_data()[3 * n + 0] = 0.1 * X;
    dpdx->host_data()[3 * n + 1] = -0.2 * Y;
    dpdx->host_data()[3 * n + 2] = 0.05 * Z;
  }
  for (stk::mesh::FieldBase *f : std::initializer_list<stk::mesh::FieldBase *>{coords, velocity, density, pressure, dpdx}) f->modify_on_host();
  SolutionOptions solnOpts;
  HelperObjects h(bulk, stk::topology::HEX_27, 1, part);
  h.build_system();
  TimeIntegrator &ti = *h.realm.timeIntegrator_;
  ti.timeStepN_ = 0.05; ti.timeStepNm1_ = 0.05; ti.gamma1_ = 1.5; ti.gamma2_ = -2.0; ti.gamma3_ = 0.5;
  std::unique_ptr<Kernel> k(build_topo_kernel<ContinuityAdvElemKernel>(stk::topology::HEX_27, *bulk, solnOpts,
                                                                         h.assembleElemSolverAlg->dataNeededByKernels_));
  h.assembleElemSolverAlg->activeKernels_.push_back(k.get());
  h.assembleElemSolverAlg->execute();
  EXPECT_TRUE(h.linsys->lastAssembleKernelFamily() == 5);
  const auto rp = h.linsys->ownedRowPtr();
  const auto ci = h.linsys->ownedColIdx();
  const auto v = h.linsys->ownedValues();
  const auto r = h.linsys->ownedRhs();
  const int32_t *en = bulk->begin_nodes(stk::mesh::Entity{stk::topology::ELEM_RANK, 0});
  std::vector<int> rows(27), pos(27, -1);
  for (int kk = 0; kk < 27; ++kk) rows[kk] = h.linsys->entityToLID(stk::mesh::Entity{stk::topology::NODE_RANK, en[kk]});
  for (int j = 0; j < 27; ++j) pos[rows[j]] = j;
  std::vector<double> lhs(27 * 27, 0.0), rhs(27);
  for (int i = 0; i < 27; ++i) {
    rhs[i] = r[rows[i]];
    for (uint64_t q = rp[rows[i]]; q < rp[rows[i] + 1]; ++q) lhs[size_t(i) * 27 + pos[ci[q]]] = v[q];
  }
  unit_test_utils::expect_all_near(lhs, gold("hex27_host", "continuity::lhs"), 1e-13);
  unit_test_utils::expect_all_near(rhs, gold("hex27_host", "continuity::rhs"), 2e-12); // entries ~ mdot / tau = 30 mdot
}

// ---------------------------------------------------------------------------------------------------
// Face/element boundary-condition algorithm (unit_tests/kernels/UnitTestFaceElemBasic.C)
// ---------------------------------------------------------------------------------------------------
// Hex8ElementWithBCFields (unit_tests/UnitTestUtils.h:181-256): one reference element, every side in "all_surfaces",
// the fields the BC kernels read, initialised to the fixture's constants (density is left zero-initialised there; 1 here
// so that the entrainment velocity of the open kernel is finite and the results can be checked)
struct Hex8ElementWithBCFields {
  Hex8ElementWithBCFields(int n = 1) : meta_(3), bulk_(std::make_shared<stk::mesh::BulkData>(meta_))
  {
    block_ = &meta_.declare_part_with_topology("block_1", stk::topology::HEX_8);
    auto node = [&](const char *name, unsigned scalars, double init) {
      auto *f = &meta_.declare_field<double>(stk::topology::NODE_RANK, name);
      stk::mesh::put_field_on_mesh(*f, *block_, scalars);
      inits_.push_back({f, init});
      return f;
    };
    coordinates_ = node("coordinates", 3, 0.0);
    velocity_ = node("velocity", 3, 1.0);
    density_ = node("density", 1, 1.0);
    viscosity_ = node("viscosity", 1, 1.0);
    bcVelocityOpen_ = node("open_velocity_bc", 3, 1.0);
    bcVelocityWall_ = node("wall_velocity_bc", 3, 0.0);
    Gjui_ = node("dudx", 9, 1.0);
    scalarQ_ = node("scalar_q", 1, 1.0);
    bcScalarQ_ = node("bc_scalar_q", 1, 1.0);
    Gjq_ = node("Gjq", 3, 1.0);
    pressure_ = node("pressure", 1, 0.0);
    pressureBc_ = node("pressure_bc", 1, 0.0);
    dpdx_ = node("dpdx", 3, 0.0);
    exposedAreaVec_ = &meta_.declare_field<double>(meta_.side_rank(), "exposed_area_vector");
    openMdot_ = &meta_.declare_field<double>(meta_.side_rank(), "open_mass_flow_rate");
    dynamicP_ = &meta_.declare_field<double>(meta_.side_rank(), "dynamic_pressure");
    stk::mesh::put_field_on_mesh(*exposedAreaVec_, *block_, 12);
    stk::mesh::put_field_on_mesh(*openMdot_, *block_, 4);
    stk::mesh::put_field_on_mesh(*dynamicP_, *block_, 4);
    wallUtau_ = &meta_.declare_field<double>(meta_.side_rank(), "wall_friction_velocity_bip");
    wallYp_ = &meta_.declare_field<double>(meta_.side_rank(), "wall_normal_distance_bip");
    stk::mesh::put_field_on_mesh(*wallUtau_, *block_, 4);
    stk::mesh::put_field_on_mesh(*wallYp_, *block_, 4);
    bulk_->fill_generated_mesh(n, n, n, *block_);
    bulk_->create_generated_sidesets("xXyYzZ");
    for (auto &fi : inits_) {
      if (fi.first == coordinates_) continue;
      std::fill(fi.first->host_data(), fi.first->host_data() + fi.first->num_entities() * fi.first->max_size(), fi.second);
      fi.first->modify_on_host();
    }
    // oneVecFour = {1, 1, -1, -1}: two leaving and two entraining integration points per face (UnitTestUtils.h:188,227)
    for (size_t f = 0; f < bulk_->num_faces(); ++f) {
      double *m = openMdot_->host_data() + 4 * f;
      m[0] = m[1] = 1.0;
      m[2] = m[3] = -1.0;
    }
    openMdot_->modify_on_host();
  }
  stk::mesh::MetaData meta_;
  std::shared_ptr<stk::mesh::BulkData> bulk_;
  stk::mesh::Part *block_;
  SolutionOptions solnOpts_;
  VectorFieldType *coordinates_, *velocity_, *bcVelocityOpen_, *bcVelocityWall_, *Gjq_, *dpdx_;
  ScalarFieldType *density_, *viscosity_, *scalarQ_, *bcScalarQ_, *pressure_, *pressureBc_;
  GenericFieldType *Gjui_, *exposedAreaVec_, *openMdot_, *dynamicP_, *wallUtau_, *wallYp_;
  std::vector<std::pair<stk::mesh::Field<double> *, double>> inits_;
};

// CPU: the connectivity the face/element algorithm hands to its kernels (UnitTestFaceElemBasic.C:83-125, TestFaceElemKernel)
static void test_face_elem_basic()
{
  std::printf("Hex8Mesh.faceElemBasic\n");
  Hex8ElementWithBCFields fx;
  EXPECT_TRUE(fx.bulk_->num_faces() == 6); // verify_faces_exist
  EXPECT_TRUE(fx.meta_.get_part("surface_1") && fx.meta_.get_part("surface_6") && !fx.meta_.get_part("surface_7"));
  EXPECT_TRUE(fx.meta_.get_part("surface_1")->face_elems().size() == 1);
  stk::topology faceTopo = stk::topology::QUAD_4, elemTopo = stk::topology::HEX_8;
  EXPECT_TRUE(faceTopo.num_nodes() == 4 && faceTopo.rank() == stk::topology::FACE_RANK);
  MasterElement *meFC = MasterElementRepo::get_surface_master_element(faceTopo);
  MasterElement *meSCS = MasterElementRepo::get_surface_master_element(elemTopo);
  MasterElement *meSCV = MasterElementRepo::get_volume_master_element(elemTopo);
  EXPECT_TRUE(meFC->numIntPoints_ == 4 && meFC->nodesPerElement_ == 4);
  Realm realm;
  realm.bulkData_ = fx.bulk_;
  EquationSystem eq(realm, "eq");
  eq.linsys_ = LinearSystem::create(realm, 3, &eq, nullptr);
  AssembleFaceElemSolverAlgorithm faceElemAlg(realm, fx.meta_.get_part("all_surfaces"), &eq, faceTopo.num_nodes(), elemTopo.num_nodes());
  faceElemAlg.faceDataNeeded_.add_cvfem_face_me(meFC);
  faceElemAlg.elemDataNeeded_.add_cvfem_surface_me(meSCS);
  faceElemAlg.elemDataNeeded_.add_cvfem_volume_me(meSCV);
  EXPECT_TRUE(faceElemAlg.faceDataNeeded_.get_cvfem_face_me() == meFC && faceElemAlg.rhsSize_ == 12 && faceElemAlg.numDof_ == 3);
  unsigned numTimesExecuted = 0;
  std::set<int> ordinals;
  faceElemAlg.run_face_elem_algorithm(*fx.bulk_, [&](const FaceElemConnectivity &c) {
    int faceNodeOrdinals[4];
    elemTopo.face_node_ordinals(unsigned(c.elemFaceOrdinal), faceNodeOrdinals);
    for (unsigned i = 0; i < faceTopo.num_nodes(); ++i) {
      EXPECT_TRUE(c.faceNodes[i] == c.elemNodes[faceNodeOrdinals[i]]);
      EXPECT_TRUE(faceNodeOrdinals[i] == meSCS->side_node_ordinals(c.elemFaceOrdinal)[i]);
      EXPECT_TRUE(faceNodeOrdinals[i] == meSCS->ipNodeMap(c.elemFaceOrdinal)[i]);
    }
    ordinals.insert(c.elemFaceOrdinal);
    ++numTimesExecuted;
  });
  EXPECT_TRUE(numTimesExecuted == 6 && ordinals.size() == 6);
  // opposing nodes: across the element, never on the face (Hex8CVFEM.C:288-300)
  const int opp0[4] = {3, 2, 6, 7}, opp5[4] = {0, 1, 2, 3};
  for (int k = 0; k < 4; ++k) EXPECT_TRUE(meSCS->opposingNodes(0, k) == opp0[k] && meSCS->opposingNodes(5, k) == opp5[k]);
  // Quad3DSCS shape functions: partition of unity; shifted = collocation
  double shp[16], shs[16];
  meFC->shape_fcn(shp);
  meFC->shifted_shape_fcn(shs);
  for (int p = 0; p < 4; ++p) {
    EXPECT_NEAR(shp[4 * p] + shp[4 * p + 1] + shp[4 * p + 2] + shp[4 * p + 3], 1.0, 1e-15);
    EXPECT_NEAR(shp[4 * p + p], 0.5625, 1e-15);
    EXPECT_NEAR(shs[4 * p + p], 1.0, 0.0);
  }
  // a face field registered with (numFaceIp, nDim), as the BC kernels do
  ElemDataRequests req;
  req.add_face_field(*fx.exposedAreaVec_, 4, 3);
  EXPECT_THROW(req.add_face_field(*fx.velocity_, 4, 3));
}

struct FaceHelper {
  FaceHelper(Hex8ElementWithBCFields &fx, int numDof) : eq(realm, "eq")
  {
    realm.bulkData_ = fx.bulk_;
    linsys = dynamic_cast<DeviceCsrLinearSystem *>(LinearSystem::create(realm, numDof, &eq, nullptr));
    eq.linsys_ = linsys;
    alg.reset(new AssembleFaceElemSolverAlgorithm(realm, fx.meta_.get_part("all_surfaces"), &eq, 4, 8));
    alg->initialize_connectivity();
    linsys->finalizeLinearSystem();
    linsys->zeroSystem();
    // ComputeGeometryBoundaryAlgorithm: exposed_area_vector on the device, adopted by the field
    nalu_b200_face_set *fs = nullptr;
    stk::mesh::Part *all = fx.meta_.get_part("all_surfaces");
    if (nalu_b200_face_set_create(&fs, linsys->graph(), int64_t(all->face_elems().size()), all->face_elems().data(),
                                  all->face_ordinals().data()) != 0)
      throw std::runtime_error(nalu_b200_last_error());
    void *d = nullptr;
    nalu_b200_malloc(&d, fx.bulk_->num_faces() * 12 * 8);
    if (nalu_b200_compute_exposed_area(linsys->graph(), fs, fx.coordinates_->device_data(), (double *)d, nullptr) != 0)
      throw std::runtime_error(nalu_b200_last_error());
    nalu_b200_memcpy_d2h(fx.exposedAreaVec_->host_data(), d, fx.bulk_->num_faces() * 12 * 8, nullptr);
    nalu_b200_stream_sync(nullptr);
    nalu_b200_free(d);
    nalu_b200_face_set_destroy(fs);
    fx.exposedAreaVec_->modify_on_host();
  }
  ~FaceHelper() { alg->activeKernels_.clear(); }
  Realm realm;
  EquationSystem eq;
  DeviceCsrLinearSystem *linsys;
  std::unique_ptr<AssembleFaceElemSolverAlgorithm> alg;
};

// GPU: UnitTestFaceElemBasic.C:148-229 (faceElemMomentumSymmetry, faceElemMomentumOpen, faceElemScalarOpen) -- the
// reference only executes these; here the uniform fields of the fixture give closed forms
static void test_face_elem_kernels()
{
  std::printf("Hex8ElementWithBCFields.faceElemMomentumSymmetry / faceElemMomentumOpen / faceElemScalarOpen / ContinuityOpen\n");
  {
    Hex8ElementWithBCFields fx;
    FaceHelper h(fx, 3);
    // the exposed area vectors of a unit cube: 1/4 per integration point, outward
    double asum[3] = {0, 0, 0}, amag = 0;
    for (size_t i = 0; i < fx.bulk_->num_faces() * 4; ++i) {
      const double *a = fx.exposedAreaVec_->host_data() + 3 * i;
      for (int j = 0; j < 3; ++j) asum[j] += a[j];
      amag += std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    }
    EXPECT_NEAR(amag, 6.0, 1e-14);
    EXPECT_NEAR(std::fabs(asum[0]) + std::fabs(asum[1]) + std::fabs(asum[2]), 0.0, 1e-15);
    MomentumSymmetryElemKernel<AlgTraitsQuad4Hex8> k(fx.meta_, fx.solnOpts_, fx.velocity_, fx.viscosity_, h.alg->faceDataNeeded_,
                                                    h.alg->elemDataNeeded_);
    h.alg->activeKernels_.push_back(&k);
    h.alg->execute();
    // uniform velocity: no strain, the residual vanishes and every lhs row annihilates a constant
    const auto v = h.linsys->ownedValues();
    const auto r = h.linsys->ownedRhs();
    const auto rp = h.linsys->ownedRowPtr();
    double vmax = 0, rmax = 0, rowsum = 0;
    for (double x : v) vmax = std::fmax(vmax, std::fabs(x));
    for (double x : r) rmax = std::fmax(rmax, std::fabs(x));
    for (size_t row = 0; row + 1 < rp.size(); ++row) {
      double s = 0;
      for (uint64_t q = rp[row]; q < rp[row + 1]; ++q) s += v[q];
      rowsum = std::fmax(rowsum, std::fabs(s));
    }
    EXPECT_TRUE(vmax > 0.1);
    EXPECT_NEAR(rmax, 0.0, 1e-14);
    EXPECT_NEAR(rowsum, 0.0, 1e-14);
  }
  {
    // MomentumWallFunctionElemKernel on every side of the unit cube, u = (1,1,1), wall at rest, rho = mu = 1: a point below
    // y+_crit (y_p = 0.5, u_tau = 1: y+ = 0.5) pulls with lambda = mu / y_p |A| = 0.5, one in the log layer (y_p = 0.5,
    // u_tau = 40: y+ = 20) with rho kappa u_tau / ln(E y+) |A|.  Per point the residual is -lambda (I - n n^T) u: the two
    // wall-parallel components, -lambda each
    Hex8ElementWithBCFields fx;
    FaceHelper h(fx, 3);
    const double utau[2] = {1.0, 40.0};
    for (size_t f = 0; f < fx.bulk_->num_faces(); ++f)
      for (int ip = 0; ip < 4; ++ip) {
        fx.wallUtau_->host_data()[4 * f + ip] = utau[ip & 1];
        fx.wallYp_->host_data()[4 * f + ip] = 0.5;
      }
    fx.wallUtau_->modify_on_host();
    fx.wallYp_->modify_on_host();
    fx.solnOpts_.initialize_turbulence_constants(); // TM_elog, TM_kappa, TM_yplus_crit (src/SolutionOptions.C:742-770)
    MomentumWallFunctionElemKernel<AlgTraitsQuad4Hex8> k(*fx.bulk_, fx.solnOpts_, h.alg->faceDataNeeded_);
    h.alg->activeKernels_.push_back(&k);
    h.alg->execute();
    const double lam[2] = {1.0 / 0.5 * 0.25, 0.41 * 40.0 / std::log(9.8 * 20.0) * 0.25};
    const auto r = h.linsys->ownedRhs();
    const auto v = h.linsys->ownedValues();
    double rs = 0, vs = 0;
    for (double x : r) rs += x;
    for (double x : v) vs += x;
    // 6 faces x 4 points, two of each kind per face, two tangential components each
    EXPECT_NEAR(rs, -6.0 * 2.0 * 2.0 * (lam[0] + lam[1]), 1e-12);
    // lhs: lambda (I - n n^T) x shape functions (summing to 1): trace 2 per point, the off-diagonal n_i n_j vanish on a cube
    EXPECT_NEAR(vs, 6.0 * 2.0 * 2.0 * (lam[0] + lam[1]), 1e-12);
  }
  {
    Hex8ElementWithBCFields fx;
    FaceHelper h(fx, 3);
    MomentumOpenAdvDiffElemKernel<AlgTraitsQuad4Hex8> k(fx.meta_, fx.solnOpts_, &h.eq, fx.velocity_, fx.Gjui_, fx.viscosity_,
                                                       h.alg->faceDataNeeded_, h.alg->elemDataNeeded_);
    h.alg->activeKernels_.push_back(&k);
    h.alg->execute();
    // uniform u = (1,1,1): no viscous traction.  Leaving points (mdot = 1; alphaUpw = 1 but hybrid factor 0 -> pecfac = 0):
    // rhs_i -= u_i and the lhs gets the advection shape functions: 6 faces x 2 points x 3 components -> sum lhs = 36,
    // sum rhs = -36.  Entraining points (mdot = -1): rhs_i -= mdot (uEntrain n_i + uspec_i - (uspec . n) n_i) with
    // uEntrain = mdot / (rho |A|) = -4 and uspec = (1,1,1): summed over i, -2 per point on a +side face and +6 on a -side
    // face -> 3 x 2 x (-2) + 3 x 2 x 6 = +24.  Total -12.
    const auto v = h.linsys->ownedValues();
    const auto r = h.linsys->ownedRhs();
    double vs = 0, rs = 0;
    bool finite = true;
    for (double x : v) { vs += x; finite = finite && std::isfinite(x); }
    for (double x : r) { rs += x; finite = finite && std::isfinite(x); }
    EXPECT_TRUE(finite);
    EXPECT_NEAR(vs, 36.0, 1e-12);
    EXPECT_NEAR(rs, -12.0, 1e-12);
  }
  {
    Hex8ElementWithBCFields fx;
    FaceHelper h(fx, 1);
    ScalarOpenAdvElemKernel<AlgTraitsQuad4Hex8> k(fx.meta_, fx.solnOpts_, &h.eq, fx.scalarQ_, fx.bcScalarQ_, fx.Gjq_, fx.viscosity_,
                                                 h.alg->faceDataNeeded_, h.alg->elemDataNeeded_);
    h.alg->activeKernels_.push_back(&k);
    h.alg->execute();
    // q = q_bc = 1: every leaving point takes 1 out of its node, every entraining point brings 1 in
    const auto v = h.linsys->ownedValues();
    const auto r = h.linsys->ownedRhs();
    double vs = 0, rs = 0, rabs = 0;
    for (double x : v) vs += x;
    for (double x : r) { rs += x; rabs += std::fabs(x); }
    EXPECT_NEAR(vs, 12.0, 1e-13); // 6 faces x 2 leaving points x sum of the shape functions
    EXPECT_NEAR(rs, 0.0, 1e-13);
    EXPECT_TRUE(rabs <= 24.0 + 1e-12);
  }
  {
    Hex8ElementWithBCFields fx;
    FaceHelper h(fx, 1);
    h.realm.timeIntegrator_->timeStepN_ = 0.5;
    ContinuityOpenElemKernel<AlgTraitsQuad4Hex8> k(fx.meta_, fx.solnOpts_, h.alg->faceDataNeeded_, h.alg->elemDataNeeded_);
    h.alg->activeKernels_.push_back(&k);
    h.alg->execute();
    // p = p_bc = 0, uniform rho u = (1,1,1): mdot = u . A; over the closed surface sum rhs = -sum mdot / tau = 0;
    // every row's gradient columns cancel, leaving the penalty 2 |A| / L = 2 x 1/4 x 1 per point -> total 6 x 4 x 1/2
    const auto v = h.linsys->ownedValues();
    const auto r = h.linsys->ownedRhs();
    double vs = 0, rs = 0, rabs = 0;
    for (double x : v) vs += x;
    for (double x : r) { rs += x; rabs += std::fabs(x); }
    EXPECT_NEAR(vs, 12.0, 1e-12);
    EXPECT_NEAR(rs, 0.0, 1e-12);
    EXPECT_TRUE(rabs > 1.0);
  }
}

int main(int argc, char **argv)
{
  if (argc < 3) {
    std::printf("usage: unit_tests <golden_dir> cpu|gpu\n");
    return 2;
  }
  g_golden = argv[1];
  const bool gpu = std::string(argv[2]) == "gpu";
  try {
    test_elem_data_requests();
    test_kernel_registration();
    test_scratch_views();
    test_face_elem_basic();
    test_hex27_master_elements();
    if (gpu) {
      test_hex27_momentum_list();
      test_hex27_continuity();
      test_face_elem_kernels();
      test_momentum_advection_diffusion();
      test_continuity_advection("advection_default");
      test_continuity_advection("advection_reduced_sens_cvfem_poisson");
      test_continuity_advection("advection_reduced_shift_cvfem_poisson");
      test_momentum_mass(false);
      test_momentum_mass(true);
      test_momentum_buoyancy(false);
      test_momentum_buoyancy(true);
      test_momentum_nso(false);
      test_momentum_nso(true);
      test_sst_src(0);
      test_sst_src(1);
      test_sst_src(2);
      test_scalar_diff();
      test_momentum_upw_defaults_reproduce_central_gold();
      test_scalar_upw_blending();
      test_dirichlet_and_reset_rows();
      test_tpetra_basic();
      test_two_kernels_one_algorithm();
    } else {
      // no device: building a system must fail loudly, never fall back
      LowMachKernelHex8Mesh fx;
      HelperObjects h(fx.bulk_, stk::topology::HEX_8, 3, fx.partVec_[0]);
      bool threw = false;
      try { h.build_system(); } catch (const std::exception &e) { threw = true; std::printf("no-device error (expected): %s\n", e.what()); }
      EXPECT_TRUE(threw);
    }
  } catch (const std::exception &e) {
    std::printf("UNCAUGHT: %s\n", e.what());
    return 1;
  }
  std::printf("%d checks, %d failures\n", g_checks, g_failures);
  return g_failures ? 1 : 0;
}
