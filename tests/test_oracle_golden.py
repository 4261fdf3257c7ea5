"""Pin the CPU oracle against every golden vector the reference's unit tests hold for the
Hex8 element-assembly path (tests/golden/*.json, extracted verbatim from the reference by
tests/golden/make_golden.py).  Tolerance 1e-15 abs, as in the reference
(unit_tests/kernels/UnitTestKernelUtils.h:158-205), except where the reference says otherwise."""
import numpy as np
import pytest

import oracle
from nalu_b200 import fields, mesh
from golden_io import load_golden, tab
from nalu_fixtures import LowMachFixture, MixtureFractionFixture, SSTFixture

TOL = 1.0e-15

# unit_tests/UnitTestUtils.C:361-372 create_one_perturbed_element(HEXAHEDRON_8): test INPUT
PERTURBED_HEX8 = np.array([
    [-0.453578, -0.327867, -0.321027], [0.673626, -0.438218, -0.557809],
    [0.398767, 0.278356, -0.613672], [-0.511167, 0.656084, -0.510011],
    [-0.553608, -0.331961, 0.418698], [0.574086, -0.565879, 0.728578],
    [0.320175, 0.685044, 0.486804], [-0.349545, 0.510239, 0.58944]])


def test_momentum_advdiff_gold():
    g = load_golden("momentum_advdiff")
    fx = LowMachFixture()
    fx.calc_mdot(oracle)
    keep = []
    d = oracle.make_desc("momentum_adv_diff", keep, coordinates=fx.coords, velocity=fx.velocity,
                         viscosity=fx.viscosity, massFlowRate=fx.mdot, includeDivU=0.0)
    lhs, rhs = oracle.element_lhs_rhs([d], 3, fx.mesh.elem_nodes[0])
    np.testing.assert_allclose(rhs, tab(g, "hex8_golds::advection_diffusion::rhs"), rtol=0, atol=TOL)
    np.testing.assert_allclose(lhs, tab(g, "hex8_golds::advection_diffusion::lhs"), rtol=0, atol=TOL)


@pytest.mark.parametrize("variant", ["default", "reduced_sens", "shifted"])
def test_continuity_adv_gold(variant):
    g = load_golden("continuity_adv")
    fx = LowMachFixture()
    keep = []
    opts = dict(default=dict(), reduced_sens=dict(reducedSensitivities=1), shifted=dict(shiftPoisson=1))[variant]
    d = oracle.make_desc("continuity_adv", keep, coordinates=fx.coords, velocity=fx.velocity,
                         density=fx.density, pressure=fx.pressure, Gpdx=fx.dpdx, dt=1.0, gamma1=1.0, **opts)
    lhs, rhs = oracle.element_lhs_rhs([d], 1, fx.mesh.elem_nodes[0])
    ns = "advection_default" if variant == "default" else "advection_reduced_sensitivities"
    np.testing.assert_allclose(rhs, tab(g, f"hex8_golds::{ns}::rhs"), rtol=0, atol=TOL)
    np.testing.assert_allclose(lhs, tab(g, f"hex8_golds::{ns}::lhs"), rtol=0, atol=TOL)


def test_scalar_advdiff_gold_perturbed():
    g = load_golden("scalar_advdiff")
    fx = MixtureFractionFixture(perturb=True)
    fx.calc_mdot(oracle)
    keep = []
    d = oracle.make_desc("scalar_adv_diff", keep, coordinates=fx.coords, scalarQ=fx.mixFraction,
                         viscosity=fx.viscosity, massFlowRate=fx.mdot)
    lhs, rhs = oracle.element_lhs_rhs([d], 1, fx.mesh.elem_nodes[0])
    np.testing.assert_allclose(rhs, tab(g, "hex8_golds::advection_diffusion::rhs"), rtol=0, atol=TOL)
    np.testing.assert_allclose(lhs, tab(g, "hex8_golds::advection_diffusion::lhs"), rtol=0, atol=TOL)


def test_scalar_diff_gold():
    # HeatCondKernelHex8Mesh: thermal conductivity 1.0 (UnitTestKernelUtils.h:446-478)
    g = load_golden("scalar_diff")
    fx = LowMachFixture()
    keep = []
    temperature = np.cos(fx.coords[:, 0])  # rhs is not checked by the reference test
    d = oracle.make_desc("scalar_diff", keep, coordinates=fx.coords, scalarQ=temperature,
                         viscosity=np.ones(8))
    lhs, _ = oracle.element_lhs_rhs([d], 1, fx.mesh.elem_nodes[0])
    np.testing.assert_allclose(lhs, tab(g, "hex8_golds::heatcond::cvfem_diff::lhs"), rtol=0, atol=TOL)


@pytest.mark.parametrize("lumped", [False, True])
def test_momentum_mass_gold(lumped):
    g = load_golden("momentum_mass")
    fx = LowMachFixture()
    keep = []
    d = oracle.make_desc("momentum_mass", keep, coordinates=fx.coords, velocity=fx.velocity,
                         velocityN=fx.velocityN, velocityNm1=fx.velocityN, density=fx.density,
                         densityN=fx.densityN, densityNm1=fx.densityN, Gpdx=fx.dpdx,
                         lumpedMass=int(lumped), dt=0.1, gamma1=1.0, gamma2=-1.0, gamma3=0.0)
    lhs, rhs = oracle.element_lhs_rhs([d], 3, fx.mesh.elem_nodes[0])
    if not lumped:
        np.testing.assert_allclose(rhs, tab(g, "hex8_golds::momentum_time_derivative::rhs"), rtol=0, atol=TOL)
        np.testing.assert_allclose(lhs, tab(g, "hex8_golds::momentum_time_derivative::lhs"), rtol=0, atol=TOL)
    else:
        np.testing.assert_allclose(rhs, tab(g, "hex8_golds::momentum_time_derivative_lumped::rhs"), rtol=0, atol=TOL)
        # UnitTestMomentumMassElem.C:222-232: lumped lhs is 1.25 on the diagonal, 0 elsewhere
        np.testing.assert_allclose(lhs, 1.25 * np.eye(24), rtol=0, atol=TOL)


@pytest.mark.parametrize("lumped", [False, True])
def test_continuity_mass_gold(lumped):
    # unit_tests/kernels/UnitTestContinuityMassElem.C:14-97: rhs == -12.5, lhs == 0
    fx = LowMachFixture()
    keep = []
    d = oracle.make_desc("continuity_mass", keep, coordinates=fx.coords, density=fx.density,
                         densityN=fx.densityN, densityNm1=fx.densityN, lumpedMass=int(lumped),
                         dt=0.1, gamma1=1.0, gamma2=-1.0, gamma3=0.0)
    lhs, rhs = oracle.element_lhs_rhs([d], 1, fx.mesh.elem_nodes[0])
    np.testing.assert_allclose(rhs, -12.5, rtol=0, atol=1e-13)
    np.testing.assert_allclose(lhs, 0.0, rtol=0, atol=TOL)


def reference_density_of_the_reference_tests():
    """std::mt19937 seeded with 0 through std::uniform_real_distribution<double>(0.8, 1.3)
    (unit_tests/kernels/UnitTestMomentumBuoyancySrcElem.C:18-31): libstdc++ draws two 32-bit words,
    (g0 + g1 2^32) / 2^64.  tests/golden/make_golden.py checks the value against a compiled C++ draw."""
    bg = np.random.MT19937()
    bg._legacy_seeding(0)
    g0, g1 = (float(x) for x in bg.random_raw(2))
    return 0.8 + 0.5 * ((g0 + g1 * 4294967296.0) / 18446744073709551616.0)


def test_momentum_buoyancy_gold():
    # unit_tests/kernels/UnitTestMomentumBuoyancySrcElem.C:16-59: rhs_z = 0.125 g_z (1 - rhoRef), lhs == 0
    rho_ref = reference_density_of_the_reference_tests()
    assert abs(rho_ref - 1.0964223082583413) < 1e-16
    fx = LowMachFixture()
    keep = []
    d = oracle.make_desc("momentum_buoyancy_src", keep, coordinates=fx.coords, density=fx.density,
                         gravity=(0.0, 0.0, -9.81), rhoRef=rho_ref)
    lhs, rhs = oracle.element_lhs_rhs([d], 3, fx.mesh.elem_nodes[0])
    exact = np.zeros(24)
    exact[2::3] = 0.125 * -9.81 * (1.0 - rho_ref)
    np.testing.assert_allclose(rhs, exact, rtol=0, atol=TOL)
    np.testing.assert_allclose(lhs, 0.0, rtol=0, atol=0)


def test_momentum_buoyancy_boussinesq_gold():
    # unit_tests/kernels/UnitTestMomentumBuoyancyBoussinesqSrcElem.C:16-66: T = 300, Tref = 298, beta = 1
    rho_ref = reference_density_of_the_reference_tests()
    fx = LowMachFixture()
    keep = []
    d = oracle.make_desc("momentum_buoyancy_boussinesq_src", keep, coordinates=fx.coords,
                         temperature=np.full(fx.mesh.n_nodes, 300.0), gravity=(0.0, 0.0, -9.81), rhoRef=rho_ref,
                         tRef=298.0, beta=1.0)
    lhs, rhs = oracle.element_lhs_rhs([d], 3, fx.mesh.elem_nodes[0])
    exact = np.zeros(24)
    exact[2::3] = -0.125 * -9.81 * rho_ref * 1.0 * (300.0 - 298.0)
    np.testing.assert_allclose(rhs, exact, rtol=0, atol=4 * TOL)  # |rhs| ~ 2.7: two ulps
    np.testing.assert_allclose(lhs, 0.0, rtol=0, atol=0)


@pytest.mark.parametrize("kind,ns", [("tke_sst_src", "TurbKineticEnergySSTSrcElemKernel"),
                                     ("sdr_sst_src", "SpecificDissipationRateSSTSrcElemKernel")])
def test_sst_src_gold(kind, ns):
    # unit_tests/kernels/UnitTestSSTSrcElem.C:179-217, 259-297 (lumpedMass = false)
    g = load_golden("sst_src")
    fx = SSTFixture()
    keep = []
    d = oracle.make_desc(kind, keep, coordinates=fx.coords, velocity=fx.velocity, density=fx.density, tke=fx.tke,
                         sdr=fx.sdr, tvisc=fx.tvisc, fOneBlend=fx.fOneBlend, **SSTFixture.TURB)
    lhs, rhs = oracle.element_lhs_rhs([d], 1, fx.mesh.elem_nodes[0])
    np.testing.assert_allclose(rhs, tab(g, f"hex8_golds::{ns}::rhs"), rtol=0, atol=TOL)
    np.testing.assert_allclose(lhs, tab(g, f"hex8_golds::{ns}::lhs"), rtol=0, atol=TOL)


def test_sst_des_src_gold():
    # unit_tests/kernels/UnitTestSSTSrcElem.C:219-257: sst_max_length_scale = 0.5 (UnitTestKernelUtils.h:391),
    # cDESke / cDESkw from SolutionOptions::initialize_turbulence_constants (src/SolutionOptions.C)
    g = load_golden("sst_src")
    fx = SSTFixture()
    keep = []
    d = oracle.make_desc("tke_sst_des_src", keep, coordinates=fx.coords, velocity=fx.velocity, density=fx.density,
                         tke=fx.tke, sdr=fx.sdr, tvisc=fx.tvisc, fOneBlend=fx.fOneBlend,
                         maxLengthScale=np.full(fx.mesh.n_nodes, 0.5), cDESke=0.61, cDESkw=0.78,
                         betaStar=SSTFixture.TURB["betaStar"], tkeProdLimitRatio=SSTFixture.TURB["tkeProdLimitRatio"])
    lhs, rhs = oracle.element_lhs_rhs([d], 1, fx.mesh.elem_nodes[0])
    np.testing.assert_allclose(rhs, tab(g, "hex8_golds::TurbKineticEnergySSTDESSrcElemKernel::rhs"), rtol=0, atol=TOL)
    np.testing.assert_allclose(lhs, tab(g, "hex8_golds::TurbKineticEnergySSTDESSrcElemKernel::lhs"), rtol=0, atol=TOL)


@pytest.mark.parametrize("ns,alt", [("nso_2nd", 0.0), ("nso_2nd_alt", 1.0)])
def test_momentum_nso_gold(ns, alt):
    # unit_tests/kernels/UnitTestMomentumNSOElem.C:318-404: MomentumKernelHex8Mesh (velocity / density with two states, only
    # NP1 filled; dudx = the analytic gradient of the Taylor vortex), fourthFac = 0, altResFac = 0 / 1, dt = 0.1, BE
    g = load_golden("momentum_nso")
    fx = LowMachFixture()
    keep = []
    d = oracle.make_desc("momentum_nso", keep, coordinates=fx.coords, velocity=fx.velocity, velocityN=fx.velocityN,
                         velocityNm1=fx.velocityN, density=fx.density, densityN=fx.densityN, densityNm1=fx.densityN,
                         viscosity=fx.viscosity, pressure=fx.pressure, Gju=fields.dudx(fx.coords), fourthFac=0.0, altResFac=alt,
                         dt=0.1, gamma1=1.0, gamma2=-1.0, gamma3=0.0, includeDivU=0.0)
    lhs, rhs = oracle.element_lhs_rhs([d], 3, fx.mesh.elem_nodes[0])
    # the gold tables are printed with 16 decimals
    np.testing.assert_allclose(rhs, tab(g, f"hex8_golds::{ns}::rhs"), rtol=0, atol=1e-15)
    np.testing.assert_allclose(lhs, tab(g, f"hex8_golds::{ns}::lhs"), rtol=0, atol=1e-15)


def test_ksgs_src_unit_cube_closed_form():
    # src/kernel/TurbKineticEnergyKsgsSrcElemKernel.C:78-128 has no gold table in the reference's unit tests: on the unit
    # cube every SCV is 1/8 and every quantity is the nearest node's, so each node's rhs / diagonal has a closed form
    fx = SSTFixture()
    n = fx.mesh.n_nodes
    rng = np.random.default_rng(3)
    gju = rng.standard_normal((n, 9))
    ceps = np.full(n, 0.845)
    dvol = 0.125 + 0.01 * rng.random(n)
    visc = 0.02 + 0.01 * rng.random(n)
    dsq = rng.random(n)
    rho = 1.0 + 0.1 * rng.random(n)
    for lrk, limit in ((0.0, 500.0), (1.0, 500.0), (1.0, 0.5)):
        keep = []
        d = oracle.make_desc("tke_ksgs_src", keep, coordinates=fx.coords, density=rho, tke=fx.tke, tvisc=fx.tvisc,
                             Gju=gju, viscosity=visc, cEps=ceps, dsqrtkSq=dsq, dualNodalVolume=dvol, lrksgsfac=lrk,
                             tkeProdLimitRatio=limit)
        lhs, rhs = oracle.element_lhs_rhs([d], 1, fx.mesh.elem_nodes[0])
        nodes = fx.mesh.elem_nodes[0]
        G = gju[nodes].reshape(8, 3, 3)
        pk = fx.tvisc[nodes] * np.einsum("nij,nij->n", G, G + G.transpose(0, 2, 1))
        fac = ceps[nodes] * rho[nodes] * np.sqrt(fx.tke[nodes]) / np.cbrt(dvol[nodes])
        dk = fac * fx.tke[nodes] + lrk * 2.0 * visc[nodes] * dsq[nodes]
        pk = np.minimum(pk, limit * dk)
        np.testing.assert_allclose(rhs, (pk - dk) * 0.125, rtol=1e-13, atol=1e-14)
        np.testing.assert_allclose(lhs.reshape(8, 8), np.diag(1.5 * fac * 0.125), rtol=1e-13, atol=1e-15)


def test_scv_grad_op_gold_perturbed():
    # unit_tests/UnitTestKokkosME.C + UnitTestKokkosMEGold.h:13-79, tol 1e-10 (UnitTestUtils.h:67)
    g = load_golden("kokkos_me_gold")
    c = PERTURBED_HEX8
    gop, _, nbad = oracle.scv_grad_op(c, shifted=False)
    assert nbad == 0
    np.testing.assert_allclose(gop.ravel(), tab(g, "kokkos_me_gold::hex8_scv_grad_op"), rtol=0, atol=1e-10)
    gops, _, _ = oracle.scv_grad_op(c, shifted=True)
    np.testing.assert_allclose(gops.ravel(), tab(g, "kokkos_me_gold::hex8_scv_shifted_grad_op"), rtol=0, atol=1e-10)


def test_scv_volumes_unit_cube_and_affine_map():
    # unit_tests/UnitTestHexSCVDeterminant.C:62-135
    m = mesh.generated_box(1, 1, 1)
    c = m.coords[m.elem_nodes[0]]
    np.testing.assert_allclose(oracle.scv_volume(c), 0.125, rtol=0, atol=1e-15)
    rng = np.random.default_rng(7)
    Q = np.eye(3) + 0.3 * rng.standard_normal((3, 3))
    if np.linalg.det(Q) < 0:
        Q[:, 0] *= -1
    assert abs(oracle.hex_volume_grandy(c @ Q.T) - np.linalg.det(Q)) < 1e-13
    np.testing.assert_allclose(oracle.scv_volume(c @ Q.T).sum(), np.linalg.det(Q), rtol=1e-13)


def test_shape_functions_partition_of_unity_and_lumping():
    for shifted in (False, True):
        assert np.allclose(oracle.scs_shape_fcn(shifted).sum(axis=1), 1.0, atol=1e-15)
    # shifted SCV shape functions are the identity (lumped mass), Hex8CVFEM.C:99-107
    np.testing.assert_allclose(oracle.scv_shape_fcn(True), np.eye(8), rtol=0, atol=0)


def test_scs_areav_closed_surface_and_unit_cube():
    m = mesh.generated_box(1, 1, 1)
    c = m.coords[m.elem_nodes[0]]
    a = oracle.scs_areav(c)
    # each SCS of the unit cube has area 1/4 along its edge direction
    np.testing.assert_allclose(np.abs(a).sum(axis=1), 0.25, atol=1e-15)
    lr = [0, 1, 1, 2, 2, 3, 0, 3, 4, 5, 5, 6, 6, 7, 4, 7, 0, 4, 1, 5, 2, 6, 3, 7]
    for ip in range(12):
        edge = c[lr[2 * ip + 1]] - c[lr[2 * ip]]
        assert a[ip] @ edge > 0  # area vector points from left to right node


def test_lcg_matches_python_stream():
    np.testing.assert_array_equal(oracle.lcg_sequence(1, 24), mesh.lcg_stream(1, 24))


# ---- upwind adv-diff kernels (ScalarUpwAdvDiffElemKernel.C, MomentumUpwAdvDiffElemKernel.C) ---------------
# The reference has no unit test for them.  Anchors: (1) alpha = 0 and hybridFactor = 0 give pecfac = 0, every
# upwind term vanishes and the kernels must reproduce the golden tables of the central kernels; (2) with
# pecfac = 1, alphaUpw = 1, hoUpwind = 0 the advective flux is first-order upwind, restated here in numpy.
def _grad_like(c, ncomp):
    a = 0.3 * np.pi
    return np.stack([np.cos(a * c[:, j % 3] + 0.1 * j) * (1.0 + 0.05 * j) for j in range(ncomp)], axis=1)


def test_scalar_upw_reduces_to_central_gold():
    g = load_golden("scalar_advdiff")
    fx = MixtureFractionFixture(perturb=True)
    fx.calc_mdot(oracle)
    keep = []
    d = oracle.make_desc("scalar_upw_adv_diff", keep, coordinates=fx.coords, scalarQ=fx.mixFraction,
                         viscosity=fx.viscosity, density=fx.density, velocity=fx.velocity, Gjq=_grad_like(fx.coords, 3),
                         massFlowRate=fx.mdot, alpha=0.0, alphaUpw=1.0, hoUpwind=1.0, useLimiter=1, hybridFactor=0.0)
    lhs, rhs = oracle.element_lhs_rhs([d], 1, fx.mesh.elem_nodes[0])
    np.testing.assert_allclose(rhs, tab(g, "hex8_golds::advection_diffusion::rhs"), rtol=0, atol=TOL)
    np.testing.assert_allclose(lhs, tab(g, "hex8_golds::advection_diffusion::lhs"), rtol=0, atol=TOL)


def test_momentum_upw_reduces_to_central_gold():
    g = load_golden("momentum_advdiff")
    fx = LowMachFixture()
    fx.calc_mdot(oracle)
    keep = []
    d = oracle.make_desc("momentum_upw_adv_diff", keep, coordinates=fx.coords, velocity=fx.velocity,
                         viscosity=fx.viscosity, density=fx.density, Gju=_grad_like(fx.coords, 9), massFlowRate=fx.mdot,
                         includeDivU=0.0, alpha=0.0, alphaUpw=1.0, hoUpwind=1.0, useLimiter=1, hybridFactor=0.0)
    lhs, rhs = oracle.element_lhs_rhs([d], 3, fx.mesh.elem_nodes[0])
    np.testing.assert_allclose(rhs, tab(g, "hex8_golds::advection_diffusion::rhs"), rtol=0, atol=TOL)
    np.testing.assert_allclose(lhs, tab(g, "hex8_golds::advection_diffusion::lhs"), rtol=0, atol=TOL)


LRSCV = [(0, 1), (1, 2), (2, 3), (0, 3), (4, 5), (5, 6), (6, 7), (4, 7), (0, 4), (1, 5), (2, 6), (3, 7)]  # Hex8CVFEM.C:264-276


@pytest.mark.parametrize("ndof", [1, 3])
def test_upw_first_order_upwind_limit(ndof):
    """pecfac = 1 (tanh function far above its transition), alphaUpw = 1, hoUpwind = 0:
    aflux = mdot * q(upwind node); lhs gets (mdot + |mdot|)/2 on the left column, (mdot - |mdot|)/2 on the right."""
    fx = MixtureFractionFixture(perturb=True)
    mdot = fx.calc_mdot(oracle)[0] * np.array([1, -1, 1, -1, 1, 1, -1, -1, 1, -1, -1, 1.0])  # both signs
    mdot = mdot[None, :].copy()
    keep = []
    common = dict(coordinates=fx.coords, viscosity=fx.viscosity, density=fx.density, velocity=fx.velocity,
                  massFlowRate=mdot, alpha=0.0, alphaUpw=1.0, hoUpwind=0.0, pecletTanh=1, tanhTrans=-1.0e3, tanhWidth=1.0)
    if ndof == 1:
        q = fx.mixFraction[:, None]
        d = oracle.make_desc("scalar_upw_adv_diff", keep, scalarQ=fx.mixFraction, Gjq=_grad_like(fx.coords, 3), **common)
        dd = oracle.make_desc("scalar_diff", keep, coordinates=fx.coords, scalarQ=fx.mixFraction, viscosity=fx.viscosity)
    else:
        q = fx.velocity
        d = oracle.make_desc("momentum_upw_adv_diff", keep, Gju=_grad_like(fx.coords, 9), includeDivU=1.0, **common)
        dd = oracle.make_desc("momentum_adv_diff", keep, coordinates=fx.coords, velocity=fx.velocity, viscosity=fx.viscosity,
                              massFlowRate=np.zeros((1, 12)), includeDivU=1.0)
    lhs, rhs = oracle.element_lhs_rhs([d], ndof, fx.mesh.elem_nodes[0])
    lhs_d, rhs_d = oracle.element_lhs_rhs([dd], ndof, fx.mesh.elem_nodes[0])
    n = 8 * ndof
    q = q[fx.mesh.elem_nodes[0]]  # element-local node order
    elhs, erhs = lhs_d.reshape(n, n).copy(), rhs_d.copy()
    for ip, (il, ir) in enumerate(LRSCV):
        m = mdot[0, ip]
        up = il if m > 0 else ir
        for i in range(ndof):
            L, R = il * ndof + i, ir * ndof + i
            erhs[L] -= m * q[up, i]
            erhs[R] += m * q[up, i]
            elhs[L, L] += 0.5 * (m + abs(m)); elhs[R, L] -= 0.5 * (m + abs(m))
            elhs[R, R] -= 0.5 * (m - abs(m)); elhs[L, R] += 0.5 * (m - abs(m))
    scale = np.abs(elhs).max()
    np.testing.assert_allclose(lhs.reshape(n, n), elhs, rtol=0, atol=1e-14 * scale)
    np.testing.assert_allclose(rhs, erhs, rtol=0, atol=1e-14 * max(1.0, np.abs(erhs).max()))


@pytest.mark.parametrize("variant", ["default", "shifted"])
def test_compute_mdot_elem_matches_continuity_gold(variant):
    """ComputeMdotElemAlgorithm (src/ComputeMdotElemAlgorithm.C:85-254) evaluates the mass flow rate
    ContinuityAdvElemKernel.C:135-147 forms: rhs(il) -= mdot/tau, rhs(ir) += mdot/tau must reproduce the reference's
    continuity golds (unit_tests/kernels/UnitTestContinuityAdvElem.C, dt = gamma1 = 1)."""
    g = load_golden("continuity_adv")
    fx = LowMachFixture()
    sp = variant == "shifted"
    mdot = oracle.compute_mdot_elem(fx.mesh.elem_nodes, fx.coords, fx.density, fx.velocity, fx.pressure, fx.dpdx, 1.0,
                                    shift_mdot=False, shift_poisson=sp)[0]
    rhs = np.zeros(8)
    for ip, (il, ir) in enumerate(LRSCV):
        rhs[il] -= mdot[ip]
        rhs[ir] += mdot[ip]
    ns = "advection_default" if variant == "default" else "advection_reduced_sensitivities"
    np.testing.assert_allclose(rhs, tab(g, f"hex8_golds::{ns}::rhs"), rtol=0, atol=2e-15)


@pytest.mark.parametrize("ncomp", [1, 3])
def test_nodal_grad_elem_is_exact_for_linear_fields(ncomp):
    """AssembleNodalGrad(U)ElemAlgorithm: known answer -- on the (affine) generated box the interior Green-Gauss sum
    over the closed dual-volume surface of an interior node returns the gradient of a linear field exactly
    (dual nodal volume of an interior node of the unit lattice = 1)."""
    m = mesh.generated_box(4, 3, 3)
    c = m.coords
    Gm = np.array([[0.3, -1.2, 0.7], [2.0, 0.1, -0.4], [-0.6, 0.9, 1.5]])[:ncomp]
    f = c @ Gm.T + 0.25
    field = f[:, 0] if ncomp == 1 else f
    grad = oracle.assemble_nodal_grad_elem(m.elem_nodes, c, field, np.ones(m.n_nodes))
    interior = np.all((c > 0.5) & (c < np.array([3.5, 2.5, 2.5])), axis=1)
    assert interior.sum() == 3 * 2 * 2
    np.testing.assert_allclose(grad[interior], np.tile(Gm.reshape(-1), (interior.sum(), 1)), rtol=0, atol=1e-13)
