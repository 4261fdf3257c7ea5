"""CUDA parity tests (run on the B200 box): the product path, called through the C ABI, against
(1) the reference's golden vectors and (2) the CPU oracle on the same seeded inputs.
Bar: CSR graph bit-exact (tests/test_graph.py), LHS/RHS values within 1e-12 relative (FP64)."""
import numpy as np
import pytest

import oracle
from golden_io import load_golden, tab
from nalu_fixtures import LowMachFixture, MixtureFractionFixture
from parity_util import State, oracle_desc, perturbed_box, product_desc, rel_err, rel_err_rows

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

RTOL = 1.0e-12  # north_star: FP64 values within 1e-12 relative of the reference
# relative-error denominator floor for a matrix entry, as a fraction of the largest |entry| of ITS ROW.  An entry is a sum of
# up to 24 (element, integration point) contributions whose magnitudes can exceed the row's largest entry (the advective
# N_c mdot terms of opposite faces cancel on a divergence-free field): measured on the 64^3 scalar case, the entries that
# differ most from the oracle differ by 6e-15 of their row's largest entry (28 ulps), i.e. 6e-13 against this floor.
ROW_FLOOR = 1.0e-2
# same for a residual entry, as a fraction of the largest |rhs|: a residual is -(row . u) - flux, its rounding error scales with
# sum_j |a_ij u_j|, which on smooth fields is far larger than the residual itself
RHS_FLOOR = 1.0e-2


@pytest.fixture
def tune():
    """nalu_b200_tune knobs, restored to their defaults after the test."""
    from nalu_b200 import lib
    yield lib.tune
    lib.tune(box_kernel=5, box_zchunk=0, split_lists=1, cont_tiled=1, urow=1, urow_owner=1, general_scatter=1)


def _cuda():
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required: the assembly path has no CPU fallback")
    return "cuda"


def _linsys(m, ndof, **kw):
    from nalu_b200.linsys import DeviceCsrLinearSystem
    return DeviceCsrLinearSystem(m, ndof, **kw).allocate(_cuda())


def _one_element_lhs_rhs(ls, ndof):
    """TestLinearSystem::sumInto capture (unit_tests/UnitTestLinearSystem.h:40-61): rebuild the
    element lhs/rhs from the assembled one-element CSR."""
    rp, ci = ls.rowptr(), ls.colidx()
    v = ls.owned_values().cpu().numpy()
    n = 8 * ndof
    A = np.zeros((n, n))
    for r in range(n):
        A[r, ci[int(rp[r]):int(rp[r + 1])]] = v[int(rp[r]):int(rp[r + 1])]
    rl, cl = ls.node_lids()
    en = ls.mesh.elem_nodes[0]
    rows = np.concatenate([rl[en[k]] + np.arange(ndof) for k in range(8)])
    return A[np.ix_(rows, rows)], ls.owned_rhs().cpu().numpy()[rows]


def _dev_fields(fx, dev):
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device=dev)
    return t


def test_momentum_advdiff_golden():
    g = load_golden("momentum_advdiff")
    fx = LowMachFixture()
    fx.calc_mdot(oracle)
    dev = _cuda()
    t = _dev_fields(fx, dev)
    ls = _linsys(fx.mesh, 3)
    ls.zero_system()
    d = ls.make_desc("momentum_adv_diff", fields=dict(coordinates=t(fx.coords), velocity=t(fx.velocity),
                                                      viscosity=t(fx.viscosity)),
                     mass_flow_rate=t(fx.mdot), include_divu=0.0)
    ls.assemble_elem([d])
    lhs, rhs = _one_element_lhs_rhs(ls, 3)
    np.testing.assert_allclose(rhs, tab(g, "hex8_golds::advection_diffusion::rhs"), rtol=0, atol=2e-15)
    np.testing.assert_allclose(lhs, tab(g, "hex8_golds::advection_diffusion::lhs"), rtol=0, atol=2e-15)
    assert ls.geometry_errors() == 0


@pytest.mark.parametrize("variant", ["default", "reduced_sens", "shifted"])
def test_continuity_adv_golden(variant):
    g = load_golden("continuity_adv")
    fx = LowMachFixture()
    dev = _cuda()
    t = _dev_fields(fx, dev)
    ls = _linsys(fx.mesh, 1)
    ls.zero_system()
    opts = dict(default=(), reduced_sens=("reduced_sens",), shifted=("shift_poisson",))[variant]
    d = ls.make_desc("continuity_adv", options=opts, dt=1.0, gamma1=1.0,
                     fields=dict(coordinates=t(fx.coords), velocity=t(fx.velocity), density=t(fx.density),
                                 pressure=t(fx.pressure), dpdx=t(fx.dpdx)))
    ls.assemble_elem([d])
    lhs, rhs = _one_element_lhs_rhs(ls, 1)
    ns = "advection_default" if variant == "default" else "advection_reduced_sensitivities"
    np.testing.assert_allclose(rhs, tab(g, f"hex8_golds::{ns}::rhs"), rtol=0, atol=2e-15)
    np.testing.assert_allclose(lhs, tab(g, f"hex8_golds::{ns}::lhs"), rtol=0, atol=2e-15)


def test_scalar_advdiff_golden_perturbed():
    g = load_golden("scalar_advdiff")
    fx = MixtureFractionFixture(perturb=True)
    dev = _cuda()
    t = _dev_fields(fx, dev)
    ls = _linsys(fx.mesh, 1)
    ls.zero_system()
    # mass_flow_rate_scs computed on the device too (calc_mass_flow_rate_scs)
    mdot = ls.mdot_scs(t(fx.coords), t(fx.density), t(fx.velocity))
    np.testing.assert_allclose(mdot.cpu().numpy(), fx.calc_mdot(oracle), rtol=0, atol=1e-18 + 1e-14 * np.abs(fx.mdot).max())
    d = ls.make_desc("scalar_adv_diff", fields=dict(coordinates=t(fx.coords), scalarq=t(fx.mixFraction),
                                                    viscosity=t(fx.viscosity)), mass_flow_rate=mdot)
    ls.assemble_elem([d])
    lhs, rhs = _one_element_lhs_rhs(ls, 1)
    np.testing.assert_allclose(rhs, tab(g, "hex8_golds::advection_diffusion::rhs"), rtol=0, atol=1e-15)
    np.testing.assert_allclose(lhs, tab(g, "hex8_golds::advection_diffusion::lhs"), rtol=0, atol=1e-15)


def test_scalar_diff_golden():
    g = load_golden("scalar_diff")
    fx = LowMachFixture()
    dev = _cuda()
    t = _dev_fields(fx, dev)
    ls = _linsys(fx.mesh, 1)
    ls.zero_system()
    d = ls.make_desc("scalar_diff", fields=dict(coordinates=t(fx.coords), scalarq=t(np.cos(fx.coords[:, 0])),
                                                viscosity=t(np.ones(8))))
    ls.assemble_elem([d])
    lhs, _ = _one_element_lhs_rhs(ls, 1)
    np.testing.assert_allclose(lhs, tab(g, "hex8_golds::heatcond::cvfem_diff::lhs"), rtol=0, atol=2e-15)


@pytest.mark.parametrize("lumped", [False, True])
def test_momentum_mass_golden(lumped):
    g = load_golden("momentum_mass")
    fx = LowMachFixture()
    dev = _cuda()
    t = _dev_fields(fx, dev)
    ls = _linsys(fx.mesh, 3)
    ls.zero_system()
    d = ls.make_desc("momentum_mass", options=(("lumped_mass",) if lumped else ()), dt=0.1, gamma1=1.0, gamma2=-1.0,
                     gamma3=0.0,
                     fields=dict(coordinates=t(fx.coords), velocity=t(fx.velocity), velocity_n=t(fx.velocityN),
                                 velocity_nm1=t(fx.velocityN), density=t(fx.density), density_n=t(fx.densityN),
                                 density_nm1=t(fx.densityN), dpdx=t(fx.dpdx)))
    ls.assemble_elem([d])
    lhs, rhs = _one_element_lhs_rhs(ls, 3)
    if lumped:
        np.testing.assert_allclose(rhs, tab(g, "hex8_golds::momentum_time_derivative_lumped::rhs"), rtol=0, atol=1e-14)
        np.testing.assert_allclose(lhs, 1.25 * np.eye(24), rtol=0, atol=1e-14)
    else:
        np.testing.assert_allclose(rhs, tab(g, "hex8_golds::momentum_time_derivative::rhs"), rtol=0, atol=1e-14)
        np.testing.assert_allclose(lhs, tab(g, "hex8_golds::momentum_time_derivative::lhs"), rtol=0, atol=1e-14)


def test_continuity_mass_golden():
    fx = LowMachFixture()
    dev = _cuda()
    t = _dev_fields(fx, dev)
    ls = _linsys(fx.mesh, 1)
    ls.zero_system()
    d = ls.make_desc("continuity_mass", dt=0.1, gamma1=1.0, gamma2=-1.0, gamma3=0.0,
                     fields=dict(coordinates=t(fx.coords), density=t(fx.density), density_n=t(fx.densityN),
                                 density_nm1=t(fx.densityN)))
    ls.assemble_elem([d])
    lhs, rhs = _one_element_lhs_rhs(ls, 1)
    np.testing.assert_allclose(rhs, -12.5, rtol=0, atol=1e-13)
    np.testing.assert_allclose(lhs, 0.0, rtol=0, atol=0)


@pytest.mark.parametrize("boussinesq", [False, True])
def test_momentum_buoyancy_golden(boussinesq):
    """unit_tests/kernels/UnitTestMomentumBuoyancySrcElem.C:16-59, ...BoussinesqSrcElem.C:16-66, through the
    CUDA path; then on a perturbed box with a varying density / temperature against the oracle."""
    from test_oracle_golden import reference_density_of_the_reference_tests
    rho_ref = reference_density_of_the_reference_tests()
    fx = LowMachFixture()
    dev = _cuda()
    t = _dev_fields(fx, dev)
    ls = _linsys(fx.mesh, 3)
    ls.zero_system()
    if boussinesq:
        d = ls.make_desc("momentum_buoyancy_boussinesq_src", gravity=(0.0, 0.0, -9.81), rho_ref=rho_ref, t_ref=298.0,
                         beta=1.0, fields=dict(coordinates=t(fx.coords), temperature=t(np.full(fx.mesh.n_nodes, 300.0))))
        exact_z = -0.125 * -9.81 * rho_ref * (300.0 - 298.0)
    else:
        d = ls.make_desc("momentum_buoyancy_src", gravity=(0.0, 0.0, -9.81), rho_ref=rho_ref,
                         fields=dict(coordinates=t(fx.coords), density=t(fx.density)))
        exact_z = 0.125 * -9.81 * (1.0 - rho_ref)
    ls.assemble_elem([d])
    lhs, rhs = _one_element_lhs_rhs(ls, 3)
    exact = np.zeros(24)
    exact[2::3] = exact_z
    np.testing.assert_allclose(rhs, exact, rtol=0, atol=4e-15)
    np.testing.assert_allclose(lhs, 0.0, rtol=0, atol=0)

    # perturbed 7x5x4 box, all three gravity components, varying field: vs the oracle
    m = perturbed_box(7, 5, 4)
    c = m.coords
    s = 1.0 + 0.3 * np.sin(0.7 * c[:, 0]) * np.cos(0.4 * c[:, 1] + 0.2 * c[:, 2])
    g = (0.3, -1.7, -9.81)
    keep = []
    orc = oracle.LinSys([m], 3)
    if boussinesq:
        od = oracle.make_desc("momentum_buoyancy_boussinesq_src", keep, coordinates=c, temperature=300.0 * s, gravity=g,
                              rhoRef=1.1, tRef=298.0, beta=3.4e-3)
    else:
        od = oracle.make_desc("momentum_buoyancy_src", keep, coordinates=c, density=s, gravity=g, rhoRef=1.1)
    orc.assemble_elem(0, [od])
    ls = _linsys(m, 3)
    ls.zero_system()
    tt = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
    if boussinesq:
        d = ls.make_desc("momentum_buoyancy_boussinesq_src", gravity=g, rho_ref=1.1, t_ref=298.0, beta=3.4e-3,
                         fields=dict(coordinates=tt(c), temperature=tt(300.0 * s)))
    else:
        d = ls.make_desc("momentum_buoyancy_src", gravity=g, rho_ref=1.1, fields=dict(coordinates=tt(c), density=tt(s)))
    assert ls.assemble_elem([d]) == 1
    _compare_system(ls, orc)


@pytest.mark.parametrize("kind,ns", [("tke_sst_src", "TurbKineticEnergySSTSrcElemKernel"),
                                     ("sdr_sst_src", "SpecificDissipationRateSSTSrcElemKernel")])
def test_sst_src_golden(kind, ns):
    """unit_tests/kernels/UnitTestSSTSrcElem.C:179-217, 259-297 through the CUDA path; then a perturbed box with
    the shifted SCV grad-op and lumped shape functions against the oracle."""
    from nalu_b200 import fields
    from nalu_fixtures import SSTFixture
    g = load_golden("sst_src")
    fx = SSTFixture()
    dev = _cuda()
    tt = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
    turb = dict(beta_star=0.09, tke_prod_limit_ratio=500.0, sigma_w_two=0.856, beta_one=0.075, beta_two=0.0828,
                gamma_one=5.0 / 9.0, gamma_two=0.44)
    ls = _linsys(fx.mesh, 1)
    ls.zero_system()
    f = dict(coordinates=tt(fx.coords), velocity=tt(fx.velocity), density=tt(fx.density), tke=tt(fx.tke), sdr=tt(fx.sdr),
             tvisc=tt(fx.tvisc), f_one_blend=tt(fx.fOneBlend))
    ls.assemble_elem([ls.make_desc(kind, fields=f, turb=turb)])
    lhs, rhs = _one_element_lhs_rhs(ls, 1)
    np.testing.assert_allclose(rhs, tab(g, f"hex8_golds::{ns}::rhs"), rtol=0, atol=2e-15)
    np.testing.assert_allclose(lhs, tab(g, f"hex8_golds::{ns}::lhs"), rtol=0, atol=2e-15)

    m = perturbed_box(6, 5, 4)
    c = m.coords
    a = 0.3 * np.pi
    x, y, z = c[:, 0], c[:, 1], c[:, 2]
    fld = dict(velocity=fields.velocity(c), density=1.0 + 0.2 * np.cos(a * x) * np.cos(a * y),
               tke=2.0 + np.cos(a * x) * np.sin(a * y) * np.cos(a * z), sdr=2.0 + np.cos(a * x) * np.sin(a * y) * np.sin(a * z),
               tvisc=0.3 + 0.1 * np.sin(a * z), fOneBlend=0.5 + 0.4 * np.sin(a * x) * np.cos(a * y))
    keep = []
    orc = oracle.LinSys([m], 1)
    orc.assemble_elem(0, [oracle.make_desc(kind, keep, coordinates=c, shiftedGradOp=1, lumpedMass=1, **fld,
                                           **SSTFixture.TURB)])
    ls = _linsys(m, 1)
    ls.zero_system()
    f = dict(coordinates=tt(c), velocity=tt(fld["velocity"]), density=tt(fld["density"]), tke=tt(fld["tke"]),
             sdr=tt(fld["sdr"]), tvisc=tt(fld["tvisc"]), f_one_blend=tt(fld["fOneBlend"]))
    assert ls.assemble_elem([ls.make_desc(kind, fields=f, options=("shifted_grad_op", "lumped_mass"), turb=turb)]) == 1
    _compare_system(ls, orc)


def test_sst_des_src_golden():
    """unit_tests/kernels/UnitTestSSTSrcElem.C:219-257 (sst_max_length_scale = 0.5) through the CUDA path; then a
    perturbed box where both branches of the DES length-scale switch occur, against the oracle."""
    from nalu_b200 import fields
    from nalu_fixtures import SSTFixture
    g = load_golden("sst_src")
    fx = SSTFixture()
    dev = _cuda()
    tt = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
    turb = dict(beta_star=0.09, tke_prod_limit_ratio=500.0, cdes_ke=0.61, cdes_kw=0.78)
    ls = _linsys(fx.mesh, 1)
    ls.zero_system()
    f = dict(coordinates=tt(fx.coords), velocity=tt(fx.velocity), density=tt(fx.density), tke=tt(fx.tke), sdr=tt(fx.sdr),
             tvisc=tt(fx.tvisc), f_one_blend=tt(fx.fOneBlend), sst_max_length_scale=tt(np.full(fx.mesh.n_nodes, 0.5)))
    ls.assemble_elem([ls.make_desc("tke_sst_des_src", fields=f, turb=turb)])
    lhs, rhs = _one_element_lhs_rhs(ls, 1)
    ns = "TurbKineticEnergySSTDESSrcElemKernel"
    np.testing.assert_allclose(rhs, tab(g, f"hex8_golds::{ns}::rhs"), rtol=0, atol=2e-15)
    np.testing.assert_allclose(lhs, tab(g, f"hex8_golds::{ns}::lhs"), rtol=0, atol=2e-15)

    m = perturbed_box(6, 5, 4)
    c = m.coords
    a = 0.3 * np.pi
    x, y, z = c[:, 0], c[:, 1], c[:, 2]
    fld = dict(velocity=fields.velocity(c), density=1.0 + 0.2 * np.cos(a * x) * np.cos(a * y),
               tke=2.0 + np.cos(a * x) * np.sin(a * y) * np.cos(a * z), sdr=2.0 + np.cos(a * x) * np.sin(a * y) * np.sin(a * z),
               tvisc=0.3 + 0.1 * np.sin(a * z), fOneBlend=0.5 + 0.4 * np.sin(a * x) * np.cos(a * y),
               maxLengthScale=4.0 + 3.9 * np.sin(1.3 * x + 0.4 * z))  # l_SST ~ 3 .. 15: both sides of min()
    lsst = np.sqrt(fld["tke"]) / 0.09 / fld["sdr"]
    cdes = fld["fOneBlend"] * 0.78 + (1 - fld["fOneBlend"]) * 0.61
    assert (lsst < cdes * fld["maxLengthScale"]).any() and (lsst > cdes * fld["maxLengthScale"]).any()
    for opts, oopts in (((), {}), (("shifted_grad_op", "lumped_mass"), dict(shiftedGradOp=1, lumpedMass=1))):
        keep = []
        orc = oracle.LinSys([m], 1)
        orc.assemble_elem(0, [oracle.make_desc("tke_sst_des_src", keep, coordinates=c, betaStar=0.09, tkeProdLimitRatio=10.0,
                                               cDESke=0.61, cDESkw=0.78, **oopts, **fld)])
        ls = _linsys(m, 1)
        ls.zero_system()
        f = dict(coordinates=tt(c), velocity=tt(fld["velocity"]), density=tt(fld["density"]), tke=tt(fld["tke"]),
                 sdr=tt(fld["sdr"]), tvisc=tt(fld["tvisc"]), f_one_blend=tt(fld["fOneBlend"]),
                 sst_max_length_scale=tt(fld["maxLengthScale"]))
        assert ls.assemble_elem([ls.make_desc("tke_sst_des_src", fields=f, options=opts,
                                              turb=dict(turb, tke_prod_limit_ratio=10.0))]) == 1
        _compare_system(ls, orc)


@pytest.mark.parametrize("ns,alt", [("nso_2nd", 0.0), ("nso_2nd_alt", 1.0)])
def test_momentum_nso_golden(ns, alt):
    """unit_tests/kernels/UnitTestMomentumNSOElem.C:318-404 through the CUDA path; then a perturbed box with the BDF2 time
    term, the fourth-order term, a blended residual, divU, the shifted grad-op and a separate velocity_rtm against the oracle."""
    from nalu_b200 import fields
    g = load_golden("momentum_nso")
    fx = LowMachFixture()
    dev = _cuda()
    tt = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
    ls = _linsys(fx.mesh, 3)
    ls.zero_system()
    f = dict(coordinates=tt(fx.coords), velocity=tt(fx.velocity), velocity_n=tt(fx.velocityN), velocity_nm1=tt(fx.velocityN),
             density=tt(fx.density), density_n=tt(fx.densityN), density_nm1=tt(fx.densityN), viscosity=tt(fx.viscosity),
             pressure=tt(fx.pressure), gju=tt(fields.dudx(fx.coords)))
    ls.assemble_elem([ls.make_desc("momentum_nso", fields=f, fourth_fac=0.0, alt_res_fac=alt, dt=0.1, gamma1=1.0, gamma2=-1.0,
                                   gamma3=0.0, include_divu=0.0)])
    lhs, rhs = _one_element_lhs_rhs(ls, 3)
    # The fixture's z velocity is zero: the z rows of the pde-residual gold (entries of 1e-10 .. 1e-9) are the ROUNDING
    # NOISE of a residual that cancels to zero, divided by sqrt(small_) = 1e-8 (nu = |residual| / sqrt(0 + 1e-16),
    # MomentumNSOElemKernel.C:262-270).  The oracle sums in the reference's order and reproduces even those to 1e-15
    # (tests/test_oracle_golden.py); another summation order gets other noise of the same size: x and y rows at the
    # reference's tolerance, z rows at 1e-16 / 1e-8 of the row scale.
    gl, gr = np.asarray(tab(g, f"hex8_golds::{ns}::lhs")).reshape(24, 24), np.asarray(tab(g, f"hex8_golds::{ns}::rhs"))
    lhs = np.asarray(lhs).reshape(24, 24)
    xy = np.arange(24) % 3 != 2
    np.testing.assert_allclose(rhs[xy], gr[xy], rtol=0, atol=1e-15)
    np.testing.assert_allclose(lhs[xy], gl[xy], rtol=0, atol=1e-15)
    np.testing.assert_allclose(rhs[~xy], gr[~xy], rtol=0, atol=2e-9)
    np.testing.assert_allclose(lhs[~xy], gl[~xy], rtol=0, atol=2e-9)

    m = perturbed_box(6, 5, 4)
    st = State(m)
    d = st.to_device(dev)
    rtm = st.velocity - 0.1 * np.cos(st.coordinates)
    keep = []
    sc = dict(dt=0.05, gamma1=1.5, gamma2=-2.0, gamma3=0.5)
    orc = oracle.LinSys([m], 3)
    dd = oracle_desc(st, "momentum_nso", ("shifted_grad_op",), keep, include_divu=1.0, **sc)
    dd.fourthFac, dd.altResFac = 1.0, 0.4
    arr = np.ascontiguousarray(rtm)
    keep.append(arr)
    dd.velocityRTM = arr.ctypes.data
    orc.assemble_elem(0, [dd])
    ls = _linsys(m, 3)
    ls.zero_system()
    f = {k: v for k, v in d.items() if k != "mdot"}
    f["velocity_rtm"] = tt(rtm)
    assert ls.assemble_elem([ls.make_desc("momentum_nso", fields=f, options=("shifted_grad_op",), include_divu=1.0, fourth_fac=1.0,
                                          alt_res_fac=0.4, **sc)]) == 1
    _compare_system(ls, orc)


@pytest.mark.parametrize("lrksgsfac", [0.0, 1.0])
def test_ksgs_src_vs_oracle(lrksgsfac):
    """src/kernel/TurbKineticEnergyKsgsSrcElemKernel.C:78-128 (no reference unit test; the oracle restatement is
    checked against a closed form in tests/test_oracle_golden.py): perturbed box, the production limiter active on
    part of the nodes, with and without the low-Reynolds term."""
    dev = _cuda()
    tt = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
    m = perturbed_box(6, 5, 4)
    st = State(m)
    c = m.coords
    a = 0.3 * np.pi
    x, y, z = c[:, 0], c[:, 1], c[:, 2]
    tke = 2.0 + np.cos(a * x) * np.sin(a * y) * np.cos(a * z)
    tvisc = 0.3 + 0.1 * np.sin(a * z)
    ceps = 0.845 + 0.1 * np.cos(a * y)
    dsq = 0.5 + 0.4 * np.sin(a * x + 0.3)
    gju = st.gju * (1.0 + 8.0 * (x > 3.0))[:, None]  # production above the limit where x > 3
    limit = 2.0
    keep = []
    orc = oracle.LinSys([m], 1)
    orc.assemble_elem(0, [oracle.make_desc("tke_ksgs_src", keep, coordinates=c, density=st.density, tke=tke, tvisc=tvisc, Gju=gju,
                                           viscosity=st.viscosity, cEps=ceps, dsqrtkSq=dsq if lrksgsfac else st.viscosity,
                                           dualNodalVolume=st.dual_volume, lrksgsfac=lrksgsfac, tkeProdLimitRatio=limit)])
    ls = _linsys(m, 1)
    ls.zero_system()
    f = dict(coordinates=tt(c), density=tt(st.density), tke=tt(tke), tvisc=tt(tvisc), gju=tt(gju), viscosity=tt(st.viscosity),
             c_epsilon=tt(ceps), dsqrtk_dx_sq=tt(dsq if lrksgsfac else st.viscosity), dual_volume=tt(st.dual_volume))
    d = ls.make_desc("tke_ksgs_src", fields=f, turb=dict(tke_prod_limit_ratio=limit, lrksgs_fac=lrksgsfac))
    assert ls.assemble_elem([d]) == 1
    _compare_system(ls, orc)
    # a field the kernel gathers is an argument error when missing, not an illegal address
    f2 = dict(f)
    del f2["c_epsilon"]
    with pytest.raises(RuntimeError, match="c_epsilon"):
        ls.assemble_elem([ls.make_desc("tke_ksgs_src", fields=f2, turb=dict(tke_prod_limit_ratio=limit))])


# ---------------------------------------------------------------------------------------------
# whole-system parity against the oracle
# ---------------------------------------------------------------------------------------------
CASES = {
    "momentum": (3, [("momentum_adv_diff", (), dict(include_divu=1.0))]),
    "momentum_nodivu": (3, [("momentum_adv_diff", (), dict(include_divu=0.0))]),
    "momentum_shifted_skew": (3, [("momentum_adv_diff", ("shifted_grad_op", "skew_symmetric"), dict(include_divu=1.0))]),
    "momentum_mass+advdiff": (3, [("momentum_mass", ("lumped_mass",), dict(dt=0.05, gamma1=1.5, gamma2=-2.0, gamma3=0.5)),
                                  ("momentum_adv_diff", (), dict(include_divu=1.0))]),
    "momentum_consistent_mass": (3, [("momentum_mass", (), dict(dt=0.05, gamma1=1.5, gamma2=-2.0, gamma3=0.5, dens_fac=1.0, om_dens_fac=0.0))]),
    "continuity": (1, [("continuity_adv", (), dict(dt=0.05, gamma1=1.5))]),
    "continuity_reduced_shiftmdot": (1, [("continuity_adv", ("reduced_sens", "shift_mdot"), dict(dt=0.05, gamma1=1.5))]),
    "continuity_mass+adv": (1, [("continuity_mass", ("lumped_mass",), dict(dt=0.05, gamma1=1.5, gamma2=-2.0, gamma3=0.5)),
                                ("continuity_adv", ("shift_poisson",), dict(dt=0.05, gamma1=1.5))]),
    "scalar": (1, [("scalar_adv_diff", (), {})]),
    "scalar_mass+advdiff": (1, [("scalar_mass", (), dict(dt=0.05, gamma1=1.5, gamma2=-2.0, gamma3=0.5)),
                                ("scalar_adv_diff", ("skew_symmetric",), {})]),
    "scalar_diff": (1, [("scalar_diff", ("shifted_grad_op",), {})]),
    # upwind adv-diff kernels (BASELINE configs 1 and 4): classic Peclet blending, limiter, tanh blending
    "scalar_upw": (1, [("scalar_upw_adv_diff", ("use_limiter",),
                        dict(upw=dict(alpha=0.0, alpha_upw=1.0, ho_upwind=1.0, hybrid_factor=1.0)))]),
    "scalar_mass+upw_tanh": (1, [("scalar_mass", ("lumped_mass",), dict(dt=0.05, gamma1=1.5, gamma2=-2.0, gamma3=0.5)),
                                 ("scalar_upw_adv_diff", ("peclet_tanh", "skew_symmetric"),
                                  dict(upw=dict(alpha=0.1, alpha_upw=0.75, ho_upwind=0.5, tanh_trans=2.0, tanh_width=4.0)))]),
    "momentum_upw": (3, [("momentum_upw_adv_diff", ("use_limiter",),
                          dict(include_divu=1.0, upw=dict(alpha=0.0, alpha_upw=1.0, ho_upwind=1.0, hybrid_factor=1.0)))]),
    "momentum_mass+upw_tanh+buoyancy-free": (3, [
        ("momentum_mass", ("lumped_mass",), dict(dt=0.05, gamma1=1.5, gamma2=-2.0, gamma3=0.5)),
        ("momentum_upw_adv_diff", ("peclet_tanh", "shifted_grad_op"),
         dict(include_divu=0.0, upw=dict(alpha=0.1, alpha_upw=0.75, ho_upwind=0.5, tanh_trans=2.0, tanh_width=4.0)))]),
}


def _compare_system(ls, orc, rank=0, tol=RTOL, sol_scale=None):
    """sol_scale: when given, a residual entry is compared relative to max(|a|, |b|, RHS_FLOOR * sol_scale * max_rows
    sum_j |a_ij|): a residual is a sum of terms of the size |a_ij| |u_j| whose rounding error does not shrink when the terms
    cancel (upwinded fluxes on a nearly divergence-free field), the form tests/sampled_parity.py uses"""
    v = ls.owned_values().cpu().numpy()
    r = ls.owned_rhs().cpu().numpy()
    ov, orr = orc.values(rank), orc.rhs(rank)
    # Elementwise relative error with a per-row floor: an entry is a sum of O(row scale) terms of both
    # signs, so its absolute rounding error is O(eps * row scale) whatever the summation order; an entry that cancels
    # to << its row's scale cannot be compared relative to itself.
    ev = rel_err_rows(v, ov, orc.rowptr(rank), ROW_FLOOR)
    rfloor = RHS_FLOOR * max(np.abs(orr).max(), 1e-300)
    if sol_scale is not None:
        rp = orc.rowptr(rank).astype(np.int64)
        rfloor = max(rfloor, RHS_FLOOR * sol_scale * float(np.add.reduceat(np.abs(ov), rp[:-1]).max()))
    er = rel_err(r, orr, rfloor)
    assert ev < tol, f"values rel err {ev:.3e}"
    assert er < tol, f"rhs rel err {er:.3e}"
    if ls.n_shared_rows:
        sv = ls.shared_values().cpu().numpy()
        assert rel_err_rows(sv, orc.values(rank, True), orc.rowptr(rank, True), ROW_FLOOR) < tol
    return ev, er


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("mode", ["general", "auto"])
def test_system_parity_vs_oracle(case, mode):
    ndof, klist = CASES[case]
    m = perturbed_box(9, 7, 6)
    st = State(m)
    dev = st.to_device(_cuda())
    orc = oracle.LinSys([m], ndof)
    keep = []
    orc.assemble_elem(0, [oracle_desc(st, k, o, keep, **s) for k, o, s in klist])
    ls = _linsys(m, ndof)
    ls.zero_system()
    used = ls.assemble_elem([product_desc(ls, dev, k, o, **s) for k, o, s in klist], mode=mode, overwrite=True)
    assert used in (1, 2, 3, 4)
    tiled = {"momentum_adv_diff", "continuity_adv", "scalar_adv_diff", "scalar_diff"}
    if mode == "auto" and len(klist) > 1:
        # [lumped MomentumMass, MomentumAdvDiff]: one launch of the row-thread kernel (fused); another tiled-capable
        # kernel plus others: tiled + accumulate; otherwise the general kernel
        # ... and [lumped ScalarMass, ScalarUpwAdvDiff]: one launch of the scalar-list tiled kernel
        fused = case in ("momentum_mass+advdiff", "scalar_mass+upw_tanh")
        # ... and the momentum upwind list, which no tiled kernel takes: the row-owner kernel, on the lattice too
        rowowner = case == "momentum_mass+upw_tanh+buoyancy-free"
        assert used == (2 if fused else (4 if rowowner else (3 if any(k in tiled for k, _, _ in klist) else 1)))
    _compare_system(ls, orc)
    assert ls.geometry_errors() == 0


def test_system_parity_64cubed_scalar():
    """BASELINE config 2 size (64^3 = 262144 Hex8), scalar adv-diff, full comparison with the oracle."""
    m = perturbed_box(64, 64, 64)
    st = State(m)
    dev = st.to_device(_cuda())
    orc = oracle.LinSys([m], 1)
    keep = []
    orc.assemble_elem(0, [oracle_desc(st, "scalar_adv_diff", (), keep)])
    ls = _linsys(m, 1)
    ls.zero_system()
    ls.assemble_elem([product_desc(ls, dev, "scalar_adv_diff")], overwrite=True)
    _compare_system(ls, orc)


def test_system_parity_40cubed_momentum():
    m = perturbed_box(40, 40, 40)
    st = State(m)
    dev = st.to_device(_cuda())
    orc = oracle.LinSys([m], 3)
    keep = []
    orc.assemble_elem(0, [oracle_desc(st, "momentum_adv_diff", (), keep, include_divu=1.0)])
    ls = _linsys(m, 3)
    ls.zero_system()
    ls.assemble_elem([product_desc(ls, dev, "momentum_adv_diff", include_divu=1.0)], overwrite=True)
    _compare_system(ls, orc)


def _periodic_state(nx, ny, nz):
    """Fully periodic box (BASELINE config 2): subjects keep their coordinates and carry their monarch's field values;
    the coordinate perturbation is keyed on the (aliased) node id, so it is periodic too."""
    from nalu_b200 import mesh as M
    m = M.periodic_box(nx, ny, nz)
    M.perturb_coords_hash(m, 0.1)
    st = State(m)
    for k in st.names():
        if k != "coordinates":
            setattr(st, k, getattr(st, k)[m.periodic_monarch])
    st.mdot = oracle.calc_mass_flow_rate_scs(m.elem_nodes, st.coordinates, st.density, st.velocity)
    return m, st


@pytest.mark.parametrize("n", [(6, 5, 4), (64, 64, 64)])
def test_periodic_box_scalar_parity(n):
    """BASELINE config 2: scalar advection-diffusion on a periodic box (64^3 = 262 144 Hex8, 262 144 rows), every entry
    against the oracle."""
    m, st = _periodic_state(*n)
    dev = st.to_device(_cuda())
    orc = oracle.LinSys([m], 1)
    keep = []
    orc.assemble_elem(0, [oracle_desc(st, "scalar_adv_diff", (), keep)])
    ls = _linsys(m, 1)
    assert ls.n_owned_rows == n[0] * n[1] * n[2]
    np.testing.assert_array_equal(ls.rowptr(), orc.rowptr(0))
    np.testing.assert_array_equal(ls.colidx(), orc.colidx(0))
    ls.zero_system()
    ls.assemble_elem([product_desc(ls, dev, "scalar_adv_diff")])
    _compare_system(ls, orc)


@pytest.mark.parametrize("kind", ["shuffled", "periodic", "box"])
def test_general_kernel_is_deterministic(kind):
    """The general kernel's scatter is colour by colour, not atomic: two assemblies of the same state are bit-identical
    (unstructured connectivity, periodic aliasing, and a kernel list the tiled kernels do not take)."""
    if kind == "periodic":
        m, st = _periodic_state(9, 8, 7)
    else:
        m = perturbed_box(9, 8, 7)
        if kind == "shuffled":
            m.node_gid = (np.random.default_rng(1).permutation(m.n_nodes) + 1).astype(np.int64)
            m.box = None
        st = State(m)
    dev = st.to_device(_cuda())
    ls = _linsys(m, 3)
    assert ls.elem_colors()[0] >= 8
    descs = [product_desc(ls, dev, "momentum_mass", (), dt=0.05, gamma1=1.5, gamma2=-2.0, gamma3=0.5),
             product_desc(ls, dev, "momentum_upw_adv_diff", ("use_limiter",), include_divu=1.0,
                          upw=dict(alpha=0.0, alpha_upw=1.0, ho_upwind=1.0, hybrid_factor=1.0))]
    out = []
    for _ in range(3):
        ls.zero_system()
        assert ls.assemble_elem(descs, mode="general") == 1
        out.append((ls.values.clone(), ls.rhs.clone()))
    assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][0], out[2][0])
    assert torch.equal(out[0][1], out[1][1]) and torch.equal(out[0][1], out[2][1])
    orc = oracle.LinSys([m], 3)
    keep = []
    orc.assemble_elem(0, [oracle_desc(st, "momentum_mass", (), keep, dt=0.05, gamma1=1.5, gamma2=-2.0, gamma3=0.5),
                          oracle_desc(st, "momentum_upw_adv_diff", ("use_limiter",), keep, include_divu=1.0,
                                      upw=dict(alpha=0.0, alpha_upw=1.0, ho_upwind=1.0, hybrid_factor=1.0))])
    _compare_system(ls, orc)


@pytest.mark.parametrize("kind,ndof", [("momentum_mass", 3), ("scalar_mass", 1)])
def test_node_mass_algorithms_periodic(kind, ndof):
    """AssembleNodeSolverAlgorithm::execute excludes the periodic subject parts from its selector
    (src/AssembleNodeSolverAlgorithm.C:57-111): a row whose gid several local nodes alias gets the mass term once."""
    m, st = _periodic_state(5, 4, 3)
    dev = st.to_device(_cuda())
    sc = dict(dt=0.05, gamma1=1.5, gamma2=-2.0, gamma3=0.5)
    orc = oracle.LinSys([m], ndof)
    keep = []
    # the oracle is handed the selector's nodes: one node (the monarch) per row
    dv = st.dual_volume.copy()
    subjects = m.periodic_monarch != np.arange(m.n_nodes)
    dv_sel = np.where(subjects, 0.0, dv)
    orc.assemble_node_mass(0, oracle_desc(st, kind, (), keep, **sc), dv_sel)
    ls = _linsys(m, ndof)
    ls.zero_system()
    f = {k: v for k, v in dev.items() if k != "mdot"}
    ls.assemble_node_mass(ls.make_desc(kind, fields=f, **sc))
    _compare_system(ls, orc)


def test_momentum_256cubed_sampled_rows_vs_oracle():
    """BASELINE config 3 at full size: the tiled assembly of the 256^3 box against the oracle on every complete row of
    sub-boxes cut out at the corners, edges, faces and interior (sampled rows at 256^3), plus the row-sum
    identity of every row on the device."""
    import sampled_parity as sp
    from nalu_b200 import fields
    n = 256
    m = perturbed_box(n, n, n)
    dev = _cuda()

    def field_fn(c):
        return {"coordinates": c, "velocity": fields.velocity(c), "viscosity": fields.smooth_viscosity(c),
                "density": fields.smooth_density(c)}

    def kernel_fn(sub, f):
        keep = []
        mdot = oracle.calc_mass_flow_rate_scs(sub.elem_nodes, f["coordinates"], f["density"], f["velocity"])
        kernel_fn.keep = keep
        return [oracle.make_desc("momentum_adv_diff", keep, coordinates=f["coordinates"], velocity=f["velocity"],
                                 viscosity=f["viscosity"], massFlowRate=mdot, includeDivU=1.0)]

    h = field_fn(m.coords)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
    d = {k: t(v) for k, v in h.items()}
    ls = _linsys(m, 3)
    mdot = ls.mdot_scs(d["coordinates"], d["density"], d["velocity"])
    ls.values.fill_(float("nan"))
    ls.rhs.fill_(float("nan"))
    desc = ls.make_desc("momentum_adv_diff", fields={k: d[k] for k in ("coordinates", "velocity", "viscosity")}, mass_flow_rate=mdot,
                        include_divu=1.0)
    assert ls.assemble_elem([desc], mode="auto", overwrite=True) == 2
    res = sp.check_sampled_rows(ls, m, n, field_fn, kernel_fn, ndof=3, size=5, n_random=24, row_floor=ROW_FLOOR)
    assert res["rows"] >= 3 * 1000, res
    assert res["max_rel_err_values"] < RTOL and res["max_rel_err_rhs"] < RTOL, res
    # against the reference arithmetic in absolute coordinates (|x|/h up to 256): the reference's own rounding error shows
    raw = sp.check_sampled_rows(ls, m, n, field_fn, kernel_fn, ndof=3, size=5, n_random=24, row_floor=ROW_FLOOR, shift=False)
    assert raw["max_rel_err_values"] < 256 * RTOL and raw["max_rel_err_rhs"] < 256 * RTOL, raw
    print("256^3 sampled rows:", res, "unshifted oracle:", raw["max_rel_err_values"], raw["max_rel_err_rhs"])
    assert sp.row_sum_identity(ls, m, mdot) < 1e-11
    assert bool(torch.isfinite(ls.values).all()) and bool(torch.isfinite(ls.rhs).all())
    assert ls.geometry_errors() == 0


@pytest.mark.parametrize("kind,ndof", [("momentum_mass", 3), ("continuity_mass", 1), ("scalar_mass", 1)])
def test_node_mass_algorithms(kind, ndof):
    m = perturbed_box(5, 4, 3)
    st = State(m)
    dev = st.to_device(_cuda())
    sc = dict(dt=0.05, gamma1=1.5, gamma2=-2.0, gamma3=0.5)
    orc = oracle.LinSys([m], ndof)
    keep = []
    orc.assemble_node_mass(0, oracle_desc(st, kind, (), keep, **sc), st.dual_volume)
    ls = _linsys(m, ndof)
    ls.zero_system()
    ls.assemble_node_mass(product_desc(ls, dev, kind, **sc))
    _compare_system(ls, orc)


# ---------------------------------------------------------------------------------------------
# multi-rank: graph handshake + shared-not-owned rows + loadComplete, ranks simulated on one GPU
# ---------------------------------------------------------------------------------------------
def _build_ranks(parts, ndof):
    from test_graph import build_product_systems  # CPU helper: routes the handshake in-process
    from nalu_b200.linsys import DeviceCsrLinearSystem
    host = build_product_systems(parts, ndof)
    sends = {p.mesh.rank: p.handshake_send for p in host}
    out = []
    for m in parts:
        ex = lambda send, r=m.rank: {src: sends[src][r] for src in sends if r in sends[src]}
        out.append(DeviceCsrLinearSystem(m, ndof, exchange=ex).allocate(_cuda()))
    return out


@pytest.mark.parametrize("nranks,transport", [(2, "packed"), (2, "pull"), (3, "packed"), (3, "pull")])
def test_tpetra_two_hex_and_slabs_load_complete(nranks, transport):
    """UnitTestTpetra.C `basic` at np=2 (1x1x2) and a 3-slab variant: fixed element block, shared rows
    exported to the owner with ADD; values looked up by global ids as in verify_matrix_for_2_hex8_mesh."""
    from nalu_b200 import mesh as M
    g = load_golden("tpetra_2elem")
    elemVals = tab(g, "elemVals")
    full = M.generated_box(1, 1, 2) if nranks == 2 else M.generated_box(2, 2, 6)
    parts = M.partition_slabs(full, nranks)
    systems = _build_ranks(parts, 1)
    blk = torch.as_tensor(elemVals, device="cuda").contiguous()
    for ls in systems:
        ls.zero_system()
        ls.assemble_elem([ls.make_desc("test_fixed_block", fixed_block=blk)])
    torch.cuda.synchronize()
    # loadComplete
    for ls in systems:
        send, recv = ls.halo_peers()
        for i, (peer, n) in enumerate(recv):
            src = systems[peer]
            if transport == "packed":
                ssend, _ = src.halo_peers()
                j = [q for q, (pr, _) in enumerate(ssend) if pr == ls.mesh.rank][0]
                buf = torch.empty(ssend[j][1], dtype=torch.float64, device="cuda")
                assert ssend[j][1] == n
                src.halo_pack(j, buf)
                ls.halo_unpack_add(i, buf)
            else:
                ls.halo_pull_add(i, src.shared_values().data_ptr(), src.shared_rhs().data_ptr())
    torch.cuda.synchronize()
    # oracle for the same job
    orc = oracle.LinSys(parts, 1)
    keep = []
    d = oracle.make_desc("test_fixed_block", keep, fixedBlock=elemVals)
    for r in range(nranks):
        orc.assemble_elem(r, [d])
    orc.load_complete()
    for r, ls in enumerate(systems):
        np.testing.assert_allclose(ls.owned_values().cpu().numpy(), orc.values(r), rtol=0, atol=1e-13)
        np.testing.assert_allclose(ls.owned_rhs().cpu().numpy(), orc.rhs(r), rtol=0, atol=1e-13)
    if nranks == 2:
        lhsVals = tab(g, "lhsVals")
        for ls in systems:
            rp, ci, v, rg, cg = ls.rowptr(), ls.colidx(), ls.owned_values().cpu().numpy(), ls.row_gids(), ls.col_gids()
            for row in range(ls.n_owned_rows):
                for j in range(int(rp[row]), int(rp[row + 1])):
                    assert abs(lhsVals[rg[row] - 1][cg[ci[j]] - 1] - v[j]) < 1e-9


@pytest.mark.parametrize("ndof,kind", [(3, "momentum_adv_diff"), (1, "continuity_adv")])
def test_np4_physics_parity_after_load_complete(ndof, kind):
    from nalu_b200 import mesh as M
    full = perturbed_box(6, 5, 9)
    parts = M.partition_slabs(full, 4)
    systems = _build_ranks(parts, ndof)
    orc = oracle.LinSys(parts, ndof)
    keep = []
    sc = dict(include_divu=1.0) if ndof == 3 else dict(dt=0.05, gamma1=1.5)
    for r, (m, ls) in enumerate(zip(parts, systems)):
        st = State(m)
        dev = st.to_device("cuda")
        orc.assemble_elem(r, [oracle_desc(st, kind, (), keep, **sc)])
        ls.zero_system()
        ls.assemble_elem([product_desc(ls, dev, kind, **sc)], overwrite=True)
    torch.cuda.synchronize()
    for ls in systems:
        _, recv = ls.halo_peers()
        for i, (peer, n) in enumerate(recv):
            ls.halo_pull_add(i, systems[peer].shared_values().data_ptr(), systems[peer].shared_rhs().data_ptr())
    torch.cuda.synchronize()
    orc.load_complete()
    for r, ls in enumerate(systems):
        _compare_system(ls, orc, rank=r)


def test_missing_device_pointer_is_an_error():
    from nalu_b200.lib import NaluB200Error
    m = perturbed_box(2, 2, 2)
    ls = _linsys(m, 3)
    d = ls.make_desc("momentum_adv_diff")  # no coordinates
    with pytest.raises(NaluB200Error):
        ls.assemble_elem([d])
    ls1 = _linsys(m, 1)
    st = State(m)
    dev = st.to_device("cuda")
    with pytest.raises(NaluB200Error):
        ls1.assemble_elem([product_desc(ls1, dev, "momentum_adv_diff")])  # numDof mismatch


# ---------------------------------------------------------------------------------------------
# tiled write-once momentum kernels (assemble_box5.cu, assemble_box2.cu): z chunking, options, overwrite semantics, determinism
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kernel", [5, 2])
@pytest.mark.parametrize("zchunk", [1, 3, 0])
@pytest.mark.parametrize("options,divu", [(("skew_symmetric",), 1.0), ((), 1.0), (("shifted_grad_op", "skew_symmetric"), 1.0),
                                          (("shifted_grad_op",), 0.0)])
def test_box_kernel_tiles_and_chunks(kernel, zchunk, options, divu, tune):
    # kernel 5: row-thread kernel (assemble_box5.cu, the default); 2: column-pair kernel (assemble_box2.cu), which also
    # serves the option combinations the row-thread kernel does not take (it is then selected whatever the knob says)
    tune(box_kernel=kernel, box_zchunk=zchunk)
    m = perturbed_box(19, 9, 7)
    st = State(m)
    dev = st.to_device(_cuda())
    orc = oracle.LinSys([m], 3)
    keep = []
    orc.assemble_elem(0, [oracle_desc(st, "momentum_adv_diff", options, keep, include_divu=divu)])
    ls = _linsys(m, 3)
    # overwrite mode must not depend on what the arrays held: no zeroSystem, poison instead
    ls.values.fill_(float("nan"))
    ls.rhs.fill_(float("nan"))
    d = product_desc(ls, dev, "momentum_adv_diff", options, include_divu=divu)
    assert ls.assemble_elem([d], mode="box", overwrite=True) == 2
    _compare_system(ls, orc)
    v1, r1 = ls.values.clone(), ls.rhs.clone()
    # bit-identical on a second run (no atomics, fixed summation order)
    ls.values.fill_(7.0)
    ls.rhs.fill_(7.0)
    ls.assemble_elem([d], mode="box", overwrite=True)
    assert torch.equal(v1, ls.values) and torch.equal(r1, ls.rhs)
    # accumulate mode: adds to what is there (sumInto semantics)
    ls.assemble_elem([d], mode="box", overwrite=False)
    torch.testing.assert_close(ls.values, 2 * v1, rtol=1e-15, atol=0)
    torch.testing.assert_close(ls.rhs, 2 * r1, rtol=1e-15, atol=0)
    assert ls.geometry_errors() == 0


@pytest.mark.parametrize("zchunk", [0, 2])
@pytest.mark.parametrize("options,divu,order", [((), 1.0, 0), (("shifted_grad_op", "skew_symmetric"), 0.0, 1)])
def test_fused_lumped_mass_list(options, divu, order, zchunk, tune):
    """The momentum list of the low-Mach equation system, [MomentumMassElemKernel(lumped), MomentumAdvDiffElemKernel]
    (src/LowMachEquationSystem.C:1304-1398), in ONE launch of the row-thread kernel: SCV volumes from the coefficient
    form of the trilinear map, diagonal + residual terms added to the dz=0 group before it is written."""
    tune(box_zchunk=zchunk)
    m = perturbed_box(19, 9, 7)
    st = State(m)
    dev = st.to_device(_cuda())
    sc = dict(dt=0.05, gamma1=1.5, gamma2=-2.0, gamma3=0.5, dens_fac=0.25, om_dens_fac=0.75)
    klist = [("momentum_mass", ("lumped_mass",), sc), ("momentum_adv_diff", options, dict(include_divu=divu))]
    if order:
        klist.reverse()
    orc = oracle.LinSys([m], 3)
    keep = []
    orc.assemble_elem(0, [oracle_desc(st, k, o, keep, **s) for k, o, s in klist])
    ls = _linsys(m, 3)
    ls.values.fill_(float("nan"))
    ls.rhs.fill_(float("nan"))
    descs = [product_desc(ls, dev, k, o, **s) for k, o, s in klist]
    launches0 = ls.L.nalu_b200_launch_count()
    assert ls.assemble_elem(descs, mode="auto", overwrite=True) == 2
    assert ls.L.nalu_b200_launch_count() - launches0 == 1
    _compare_system(ls, orc)
    assert ls.geometry_errors() == 0
    v1 = ls.values.clone()
    ls.assemble_elem(descs, mode="auto", overwrite=True)
    assert torch.equal(v1, ls.values)  # bit-reproducible


def test_row_thread_and_column_pair_kernels_agree_64cubed(tune):
    """The two tiled momentum kernels on a block with many interior (natural) x lines and every kind of boundary line."""
    m = perturbed_box(67, 37, 21)
    st = State(m)
    dev = st.to_device(_cuda())
    ls = _linsys(m, 3)
    d = product_desc(ls, dev, "momentum_adv_diff", (), include_divu=1.0)
    tune(box_kernel=2)
    ls.values.fill_(float("nan"))
    ls.rhs.fill_(float("nan"))
    assert ls.assemble_elem([d], mode="box", overwrite=True) == 2
    v2, r2 = ls.values.clone(), ls.rhs.clone()
    tune(box_kernel=5)
    ls.values.fill_(float("nan"))
    ls.rhs.fill_(float("nan"))
    assert ls.assemble_elem([d], mode="box", overwrite=True) == 2
    assert float((ls.values - v2).abs().max() / v2.abs().max()) < 1e-14
    assert float((ls.rhs - r2).abs().max() / r2.abs().max()) < 1e-13
    assert ls.geometry_errors() == 0


def test_box_kernel_zrange_pieces_equal_one_launch():
    """nalu_b200_assemble_elem_zrange: the rows of consecutive node-plane ranges, assembled by separate
    launches, are bit-identical to one launch over the block (each row is written by exactly one launch)."""
    m = perturbed_box(11, 10, 37)
    st = State(m)
    dev = st.to_device(_cuda())
    ls = _linsys(m, 3)
    d = product_desc(ls, dev, "momentum_adv_diff", ("skew_symmetric",), include_divu=1.0)
    ls.values.fill_(float("nan"))
    ls.rhs.fill_(float("nan"))
    assert ls.assemble_elem([d], mode="box", overwrite=True) == 2
    v1, r1 = ls.values.clone(), ls.rhs.clone()
    ls.values.fill_(float("nan"))
    ls.rhs.fill_(float("nan"))
    for z0, z1 in ((0, 1), (1, 17), (17, 20), (20, 38)):
        ls.assemble_elem_zrange([d], z0, z1, overwrite=True)
    assert torch.equal(v1, ls.values) and torch.equal(r1, ls.rhs)
    with pytest.raises(Exception):
        ls.assemble_elem_zrange([d], 5, 39, overwrite=True)
    assert ls.geometry_errors() == 0


def test_box_kernel_matches_general_kernel_96cubed():
    """Beyond oracle-sized meshes: the two CUDA kernel families (atomic sumInto vs tiled write-once)
    are independent implementations; they must agree on a mesh too large for the CPU oracle."""
    m = perturbed_box(96, 96, 96)
    st = State(m)
    dev = st.to_device(_cuda())
    ls = _linsys(m, 3)
    d = product_desc(ls, dev, "momentum_adv_diff", include_divu=1.0)
    ls.zero_system()
    assert ls.assemble_elem([d], mode="general") == 1
    vg, rg = ls.values.clone(), ls.rhs.clone()
    assert ls.assemble_elem([d], mode="box", overwrite=True) == 2
    scale = float(vg.abs().max())
    assert float((ls.values - vg).abs().max()) / scale < 1e-12
    assert float((ls.rhs - rg).abs().max()) / float(rg.abs().max()) < 1e-12
    assert ls.geometry_errors() == 0


# ---------------------------------------------------------------------------------------------
# applyDirichletBCs / resetRows (src/TpetraLinearSystem.C:1579-1653,1706-1755)
# ---------------------------------------------------------------------------------------------
def _row_fixup_expected(ls, m, v, r, nodes, ndof, begin, end, sol=None, bc=None):
    """numpy restatement of the two reference loops on this rank's owned + shared CSR."""
    rowptr = np.concatenate([ls.rowptr(), ls.rowptr()[-1] + ls.rowptr(shared=True)[1:]]) if ls.n_shared_rows else ls.rowptr()
    colidx = np.concatenate([ls.colidx(), ls.colidx(shared=True)]) if ls.n_shared_rows else ls.colidx()
    row_lid, _ = ls.node_lids()
    n_owned = ls.n_owned_rows
    v, r = v.copy(), r.copy()
    for node in nodes:
        for d in range(begin, end):
            lid = int(row_lid[node]) + d
            owned = lid < n_owned
            b, e = int(rowptr[lid]), int(rowptr[lid + 1])
            v[b:e] = np.where((colidx[b:e] == lid) & owned & (sol is not None), 1.0, 0.0)
            r[lid] = (bc[node, d] - sol[node, d]) if (owned and sol is not None) else 0.0
    return v, r


@pytest.mark.parametrize("ndof", [1, 3])
@pytest.mark.parametrize("nranks", [1, 2])
def test_dirichlet_and_reset_rows(ndof, nranks):
    from nalu_b200 import mesh as M
    meshes = M.partition_slabs(perturbed_box(6, 5, 4), nranks)
    systems = _build_ranks(meshes, ndof)
    for rank, (m, ls) in enumerate(zip(meshes, systems)):
        dev = _cuda()
        g = torch.Generator().manual_seed(5 + rank)
        nnz = ls.values.numel()
        ls.values.copy_(torch.rand(nnz, generator=g, dtype=torch.float64).to(dev) + 0.5)
        ls.rhs.copy_(torch.rand(ls.rhs.numel(), generator=g, dtype=torch.float64).to(dev) + 0.5)
        v0, r0 = ls.values.cpu().numpy(), ls.rhs.cpu().numpy()
        # the nodes of the x = 0 face (owned and shared alike, like the reference's selector)
        x0 = m.coords[:, 0].min()
        nodes = np.nonzero(np.abs(m.coords[:, 0] - x0) < 0.3)[0].astype(np.int32)
        sol = np.cos(m.coords[:, :1] + np.arange(ndof)[None, :])
        bc = 2.0 + np.sin(m.coords[:, 1:2] + np.arange(ndof)[None, :])
        t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
        begin, end = (0, 1) if ndof == 1 else (1, 3)
        ls.apply_dirichlet_bcs(t(nodes), t(sol), t(bc), begin, end)
        ev, er = _row_fixup_expected(ls, m, v0, r0, nodes, ndof, begin, end, sol, bc)
        np.testing.assert_array_equal(ls.values.cpu().numpy(), ev)
        np.testing.assert_array_equal(ls.rhs.cpu().numpy(), er)
        assert (ev == 1.0).sum() > 0
        # resetRows on a few other nodes
        others = np.setdiff1d(np.arange(m.n_nodes, dtype=np.int32), nodes)[::7].astype(np.int32)
        ls.reset_rows(t(others), 0, ndof)
        ev2, er2 = _row_fixup_expected(ls, m, ev, er, others, ndof, 0, ndof)
        np.testing.assert_array_equal(ls.values.cpu().numpy(), ev2)
        np.testing.assert_array_equal(ls.rhs.cpu().numpy(), er2)


# ---------------------------------------------------------------------------------------------
# tiled write-once continuity kernel (assemble_cont_box.cu)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("zchunk", [0, 1, 5])
@pytest.mark.parametrize("options", [(), ("shift_mdot",), ("reduced_sens",), ("shift_poisson", "reduced_sens")])
def test_cont_box_kernel_options_and_chunks(options, zchunk, tune):
    tune(box_zchunk=zchunk)
    m = perturbed_box(33, 16, 9)  # more than two tiles in x and y, partial tiles at the far sides
    st = State(m)
    dev = st.to_device(_cuda())
    orc = oracle.LinSys([m], 1)
    keep = []
    sc = dict(dt=0.05, gamma1=1.5)
    orc.assemble_elem(0, [oracle_desc(st, "continuity_adv", options, keep, **sc)])
    ls = _linsys(m, 1)
    ls.values.fill_(float("nan"))  # overwrite mode: every entry must be written
    ls.rhs.fill_(float("nan"))
    assert ls.assemble_elem([product_desc(ls, dev, "continuity_adv", options, **sc)], mode="auto", overwrite=True) == 2
    _compare_system(ls, orc)
    assert ls.geometry_errors() == 0
    # accumulate mode (sumInto semantics) on top of the overwritten system: twice the values
    v1, r1 = ls.owned_values().clone(), ls.owned_rhs().clone()
    assert ls.assemble_elem([product_desc(ls, dev, "continuity_adv", options, **sc)], mode="auto", overwrite=False) == 2
    assert torch.equal(ls.owned_values(), 2 * v1) and torch.equal(ls.owned_rhs(), 2 * r1)
    # bit-reproducible run to run
    ls.assemble_elem([product_desc(ls, dev, "continuity_adv", options, **sc)], mode="auto", overwrite=True)
    assert torch.equal(ls.owned_values(), v1) and torch.equal(ls.owned_rhs(), r1)


def test_cont_box_matches_general_kernel_96cubed():
    m = perturbed_box(96, 96, 96)
    st = State(m)
    dev = st.to_device(_cuda())
    sc = dict(dt=0.05, gamma1=1.5)
    ls = _linsys(m, 1)
    ls.zero_system()
    assert ls.assemble_elem([product_desc(ls, dev, "continuity_adv", **sc)], mode="general") == 1
    v0, r0 = ls.owned_values().clone(), ls.owned_rhs().clone()
    assert ls.assemble_elem([product_desc(ls, dev, "continuity_adv", **sc)], mode="auto", overwrite=True) == 2
    assert float((ls.owned_values() - v0).abs().max() / v0.abs().max()) < 1e-13
    assert float((ls.owned_rhs() - r0).abs().max() / r0.abs().max()) < 1e-13


# ---------------------------------------------------------------------------------------------
# tiled write-once scalar adv-diff / diffusion kernel (assemble_scalar_box.cu)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("zchunk", [0, 3])
@pytest.mark.parametrize("kind,options", [("scalar_adv_diff", ()), ("scalar_adv_diff", ("skew_symmetric",)),
                                          ("scalar_adv_diff", ("shifted_grad_op", "skew_symmetric")),
                                          ("scalar_diff", ()), ("scalar_diff", ("shifted_grad_op",))])
def test_scalar_box_kernel_options_and_chunks(kind, options, zchunk, tune):
    tune(box_zchunk=zchunk)
    m = perturbed_box(33, 16, 9)
    st = State(m)
    dev = st.to_device(_cuda())
    orc = oracle.LinSys([m], 1)
    keep = []
    orc.assemble_elem(0, [oracle_desc(st, kind, options, keep)])
    ls = _linsys(m, 1)
    ls.values.fill_(float("nan"))
    ls.rhs.fill_(float("nan"))
    assert ls.assemble_elem([product_desc(ls, dev, kind, options)], mode="auto", overwrite=True) == 2
    _compare_system(ls, orc)
    assert ls.geometry_errors() == 0
    v1, r1 = ls.owned_values().clone(), ls.owned_rhs().clone()
    assert ls.assemble_elem([product_desc(ls, dev, kind, options)], mode="auto", overwrite=False) == 2
    assert torch.equal(ls.owned_values(), 2 * v1) and torch.equal(ls.owned_rhs(), 2 * r1)


@pytest.mark.parametrize("options", [(), ("shift_mdot",), ("shift_poisson",), ("shift_mdot", "shift_poisson")])
def test_compute_mdot_elem(options):
    """nalu_b200_compute_mdot_elem vs the oracle's restatement of ComputeMdotElemAlgorithm; and the adv-diff assembly
    fed with the device-produced field equals the one fed with the oracle's."""
    m = perturbed_box(9, 7, 6)
    st = State(m)
    dev = st.to_device(_cuda())
    tau = 0.05 / 1.5
    ref = oracle.compute_mdot_elem(m.elem_nodes, st.coordinates, st.density, st.velocity, st.pressure, st.dpdx, tau,
                                   shift_mdot="shift_mdot" in options, shift_poisson="shift_poisson" in options)
    ls = _linsys(m, 3)
    got = ls.compute_mdot_elem(dev["coordinates"], dev["density"], dev["velocity"], dev["pressure"], dev["dpdx"], tau, options)
    scale = np.abs(ref).max()
    assert np.abs(got.cpu().numpy() - ref).max() <= 1e-13 * scale
    assert ls.geometry_errors() == 0


@pytest.mark.parametrize("ncomp,shifted", [(1, False), (1, True), (3, False)])
def test_assemble_nodal_grad_elem(ncomp, shifted):
    m = perturbed_box(9, 7, 6)
    st = State(m)
    dev = st.to_device(_cuda())
    field = st.pressure if ncomp == 1 else st.velocity
    ref = oracle.assemble_nodal_grad_elem(m.elem_nodes, st.coordinates, field, st.dual_volume, use_shifted=shifted)
    ls = _linsys(m, 1)
    grad = torch.zeros((m.n_nodes, 3 * ncomp), dtype=torch.float64, device=_cuda())
    ls.assemble_nodal_grad_elem(dev["coordinates"], dev["pressure"] if ncomp == 1 else dev["velocity"], dev["dual_volume"], grad,
                                use_shifted=shifted)
    assert np.abs(grad.cpu().numpy() - ref).max() <= 1e-13 * np.abs(ref).max()


# ---------------------------------------------------------------------------------------------
# real meshes of the reference's regression suite (tests/golden/meshes, read with nalu_b200.exodus)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["momentum_mass+advdiff", "momentum_shifted_skew", "momentum_nodivu", "continuity_mass+adv",
                                  "scalar_mass+upw_tanh", "momentum_upw", "scalar_diff"])
def test_unstructured_exodus_mesh_parity(case):
    """reg_tests/mesh/thex8_8.g (1 852 Hex8, 2 to 46 elements per node, rows of 12 to 141 node blocks): the row-thread
    unstructured kernel (momentum lists, mode 4) and the coloured general kernel (everything else, mode 1) against the
    oracle, every entry."""
    from nalu_b200 import exodus
    import os
    m = exodus.read_exodus_hex8(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "meshes", "thex8_8.g"))
    ndof, klist = CASES[case]
    st = State(m)
    dev = st.to_device(_cuda())
    orc = oracle.LinSys([m], ndof)
    keep = []
    orc.assemble_elem(0, [oracle_desc(st, k, o, keep, **s) for k, o, s in klist])
    ls = _linsys(m, ndof)
    assert ls.elem_colors()[0] >= 46
    ls.zero_system()
    urow = case in ("momentum_mass+advdiff", "momentum_shifted_skew", "momentum_nodivu", "scalar_diff", "scalar_mass+upw_tanh", "momentum_upw")
    assert ls.assemble_elem([product_desc(ls, dev, k, o, **s) for k, o, s in klist], mode="auto") == (4 if urow else 1)
    _compare_system(ls, orc)
    assert ls.geometry_errors() == 0
    # deterministic: colour-by-colour plain read-modify-writes
    v0 = ls.owned_values().clone()
    ls.zero_system()
    ls.assemble_elem([product_desc(ls, dev, k, o, **s) for k, o, s in klist], mode="auto")
    assert torch.equal(v0, ls.owned_values())


@pytest.mark.parametrize("ndof,kind", [(3, "momentum_adv_diff"), (1, "continuity_adv")])
def test_decomposed_exodus_mesh_parity_after_load_complete(ndof, kind):
    """reg_tests/mesh/ether_conform_coarse_mks.g.4.[0-3] (the SEACAS decomposition the reference's np=4 decks read): four
    ranks, irregular interfaces, assembly + loadComplete against the oracle's export-add."""
    from nalu_b200 import exodus
    import os
    base = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "meshes", "ether_conform_coarse_mks.g")
    parts = exodus.read_exodus_pieces(base, 4)
    for m in parts:
        # the deck is 7.5 cm across (mks units); the analytic test fields vary over O(1) lengths: rescale so that the
        # residual is not a 1e-4 remainder of cancelling terms (the comparison floor is 1 % of the largest residual)
        m.coords = m.coords * (10.0 / 0.075)
    systems = _build_ranks(parts, ndof)
    orc = oracle.LinSys(parts, ndof)
    keep = []
    sc = dict(include_divu=1.0) if ndof == 3 else dict(dt=0.05, gamma1=1.5)
    for r, (m, ls) in enumerate(zip(parts, systems)):
        st = State(m)
        dev = st.to_device("cuda")
        orc.assemble_elem(r, [oracle_desc(st, kind, (), keep, **sc)])
        ls.zero_system()
        assert ls.assemble_elem([product_desc(ls, dev, kind, **sc)]) == 4  # row-owner kernels, momentum and continuity
    torch.cuda.synchronize()
    for ls in systems:
        send, recv = ls.halo_peers()
        for i, (peer, n) in enumerate(recv):
            src = systems[peer]
            ssend, _ = src.halo_peers()
            j = [q for q, (pr, _) in enumerate(ssend) if pr == ls.mesh.rank][0]
            buf = torch.empty(ssend[j][1], dtype=torch.float64, device="cuda")
            src.halo_pack(j, buf)
            ls.halo_unpack_add(i, buf)
    torch.cuda.synchronize()
    orc.load_complete()
    for r, ls in enumerate(systems):
        _compare_system(ls, orc, rank=r)


def _config4_state(m, eq):
    """ABL-style stretched box (cell heights growing geometrically from the wall) + SST trig fields (tests/config4_util.py)"""
    import config4_util as c4
    m.coords = c4.abl_coords(m.coords)
    return c4.state(m.coords, eq)


def _config4_oracle(orc, rank, m, f, eq, src_lumped):
    import config4_util as c4
    descs, keep = c4.oracle_list(m, f, eq, src_lumped=src_lumped)
    orc.assemble_elem(rank, descs)


def _config4_descs(ls, m, f, eq, src_lumped, shared_dof_field, dev):
    import config4_util as c4
    d = {k: torch.as_tensor(np.ascontiguousarray(v), device=dev) for k, v in f.items()}
    if shared_dof_field:
        # as in Nalu: the equation's dof field IS turbulent_ke / specific_dissipation_rate (one field, one pointer)
        d[eq] = d["scalarq"]
    mdot = torch.as_tensor(oracle.calc_mass_flow_rate_scs(m.elem_nodes, f["coordinates"], f["density"], f["velocity"]), device=dev)
    return c4.product_list(ls, d, eq, mdot, src_lumped=src_lumped)


@pytest.mark.parametrize("eq", ["tke", "sdr"])
@pytest.mark.parametrize("path", ["tiled", "tiled_lumped_src", "general"])
def test_config4_sst_kernel_lists_on_abl_style_box(eq, path):
    """BASELINE config 4 at test size: a box stretched in z like an atmospheric-boundary-layer mesh, SST trig fields, and
    the kernel lists of the turbulence equations [ScalarMass(lumped), ScalarUpwAdvDiff, TurbKineticEnergySSTSrc |
    SpecificDissipationRateSSTSrc] on ONE algorithm (src/TurbKineticEnergyEquationSystem.C /
    SpecificDissipationRateEquationSystem.C register them in this order).  `tiled`: the dof field is shared between the
    kernels as in Nalu -> one launch of the scalar-list write-once kernel (mode 2); `general`: separate copies of the dof
    field -> the coloured general kernel (mode 1)."""
    m = perturbed_box(8, 6, 10, amp=0.05)
    f = _config4_state(m, eq)
    lumped = path == "tiled_lumped_src"
    orc = oracle.LinSys([m], 1)
    _config4_oracle(orc, 0, m, f, eq, lumped)
    ls = _linsys(m, 1)
    ls.zero_system()
    descs = _config4_descs(ls, m, f, eq, lumped, path != "general", _cuda())
    assert ls.assemble_elem(descs, mode="auto", overwrite=path != "general") == (1 if path == "general" else 2)
    _compare_system(ls, orc)
    assert ls.geometry_errors() == 0


@pytest.mark.parametrize("eq", ["tke", "sdr"])
def test_config4_lists_tiled_many_tiles_and_ranks(eq, tune):
    """The scalar-list tiled kernel across tile seams (3 x 3 tiles of 15 x 7 nodes), short z chunks and three slab ranks
    (shared rows + loadComplete) against the oracle; NaN-poisoned output arrays (write-once: no zeroSystem)."""
    import config4_util as c4
    from nalu_b200 import mesh as M
    full = perturbed_box(33, 17, 14, amp=0.05)
    full.coords = c4.abl_coords(full.coords)
    parts = M.partition_slabs(full, 3)
    systems = _build_ranks(parts, 1)
    orc = oracle.LinSys(parts, 1)
    tune(box_zchunk=3)
    for r, (m, ls) in enumerate(zip(parts, systems)):
        f = c4.state(m.coords, eq)  # analytic fields: every rank evaluates them on its own nodes
        _config4_oracle(orc, r, m, f, eq, True)
        ls.values.fill_(float("nan"))
        ls.rhs.fill_(float("nan"))
        assert ls.assemble_elem(_config4_descs(ls, m, f, eq, True, True, "cuda"), mode="auto", overwrite=True) == 2
        assert ls.geometry_errors() == 0
    torch.cuda.synchronize()
    for ls in systems:
        _, recv = ls.halo_peers()
        for i, (peer, n) in enumerate(recv):
            ls.halo_pull_add(i, systems[peer].shared_values().data_ptr(), systems[peer].shared_rhs().data_ptr())
    torch.cuda.synchronize()
    orc.load_complete()
    for r, ls in enumerate(systems):
        _compare_system(ls, orc, rank=r, sol_scale=3.0)  # tke, sdr <= 3


def test_solver_handoff_device_solution_equals_oracle_solution():
    """SURVEY 8f-3, the consumer of the path: the device-assembled momentum system [MomentumMass(lumped), MomentumAdvDiff]
    solved in place on the device (cuSPARSE products through torch.sparse_csr, Jacobi-BiCGStab), against a direct sparse
    solve (scipy) of the ORACLE-assembled system: the solutions agree, so the CSR the kernels leave in HBM is the matrix."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    from nalu_b200 import mesh as M
    m = perturbed_box(10, 9, 8)
    st = State(m)
    dev = st.to_device(_cuda())
    bdf2 = dict(dt=0.05, gamma1=1.5, gamma2=-2.0, gamma3=0.5)
    klist = [("momentum_mass", ("lumped_mass",), bdf2), ("momentum_adv_diff", (), dict(include_divu=1.0))]
    keep = []
    orc = oracle.LinSys([m], 3)
    orc.assemble_elem(0, [oracle_desc(st, k, o, keep, **s) for k, o, s in klist])
    ls = _linsys(m, 3)
    assert ls.assemble_elem([product_desc(ls, dev, k, o, **s) for k, o, s in klist], overwrite=True) == 2
    x, iters, res = ls.solve(tol=1e-12, max_iter=400)
    assert res < 1e-11 and 0 < iters < 400, (iters, res)
    n = orc.n_owned_rows(0)
    A = sp.csr_matrix((orc.values(0), orc.colidx(0), orc.rowptr(0).astype(np.int64)), shape=(n, n))
    xo = spla.spsolve(A.tocsc(), orc.rhs(0))
    err = np.abs(x.cpu().numpy() - xo).max() / np.abs(xo).max()
    assert err < 1e-9, err


@pytest.mark.parametrize("kind", ["shuffled", "periodic", "periodic_one_cell_wide"])
@pytest.mark.parametrize("opts,divu,mass", [((), 1.0, True), (("shifted_grad_op", "skew_symmetric"), 0.0, True), ((), 1.0, False)])
def test_row_thread_unstructured_momentum_kernel(kind, opts, divu, mass, tune):
    """assemble_urow.cu: [MomentumMass(lumped)]? + MomentumAdvDiff on connectivity the tiled kernels do not take (shuffled
    node ids, periodic aliasing) against the oracle: the row-owner form (ONE launch, thread = row node walking the elements
    around it, plain adds) and the element-colour form (thread = (element, row node), value-less reductions); both bit
    identical run to run."""
    if kind == "periodic":
        m, st = _periodic_state(9, 8, 7)
    elif kind == "periodic_one_cell_wide":
        # an element reaches a row through two of its nodes: no element colouring exists; the row-owner form visits the
        # element twice and needs none.  The reference itself is not defined here: sum_into_row_vec_3
        # (src/TpetraLinearSystem.C:1386-1413) walks the row forward once per sorted column id, so the second of two equal
        # ids is never found and it returns, dropping the rest of the element row -- the oracle restates that.  The check
        # is against the general kernel (atomic fallback: every contribution summed).
        m, st = _periodic_state(1, 5, 4)
    else:
        m = perturbed_box(9, 8, 7)
        m.node_gid = (np.random.default_rng(1).permutation(m.n_nodes) + 1).astype(np.int64)
        m.box = None
        st = State(m)
    dev = st.to_device(_cuda())
    bdf2 = dict(dt=0.05, gamma1=1.5, gamma2=-2.0, gamma3=0.5)
    klist = ([("momentum_mass", ("lumped_mass",), bdf2)] if mass else []) + [("momentum_adv_diff", opts, dict(include_divu=divu))]
    orc = oracle.LinSys([m], 3)
    keep = []
    orc.assemble_elem(0, [oracle_desc(st, k, o, keep, **s) for k, o, s in klist])
    ls = _linsys(m, 3)
    descs = [product_desc(ls, dev, k, o, **s) for k, o, s in klist]
    out = []
    for _ in range(2):
        launches0 = ls.L.nalu_b200_launch_count()
        ls.zero_system()
        assert ls.assemble_elem(descs, mode="auto") == 4
        assert ls.L.nalu_b200_launch_count() - launches0 == 1
        out.append((ls.values.clone(), ls.rhs.clone()))
    assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1])
    assert ls.geometry_errors() == 0
    if kind == "periodic_one_cell_wide":
        assert ls.elem_colors()[0] == 0
        tune(urow=0)
        ls.zero_system()
        assert ls.assemble_elem(descs, mode="auto") == 1
        tune(urow=1)
        gv, gr = ls.values, ls.rhs
        assert float((gv - out[0][0]).abs().max() / gv.abs().max()) < 1e-13
        assert float((gr - out[0][1]).abs().max() / gr.abs().max()) < 1e-12
        return
    _compare_system(ls, orc)
    # the element-colour form of the same kernel (one launch per colour, value-less reductions)
    tune(urow_owner=0)
    launches0 = ls.L.nalu_b200_launch_count()
    ls.zero_system()
    assert ls.assemble_elem(descs, mode="auto") == 4
    assert ls.L.nalu_b200_launch_count() - launches0 >= 8
    _compare_system(ls, orc)
    c0 = ls.values.clone()
    ls.zero_system()
    ls.assemble_elem(descs, mode="auto")
    assert torch.equal(c0, ls.values)
    tune(urow_owner=1)
    # the knob takes the list back to the general kernel
    tune(urow=0)
    ls.zero_system()
    assert ls.assemble_elem(descs, mode="auto") == 1
    tune(urow=1)
    _compare_system(ls, orc)


_ONEDOF_LISTS = {
    "scalar_adv_diff": [("scalar_adv_diff", (), {})],
    "scalar_adv_diff_shifted_skew": [("scalar_adv_diff", ("shifted_grad_op", "skew_symmetric"), {})],
    "scalar_diff_shifted": [("scalar_diff", ("shifted_grad_op",), {})],
    "continuity": [("continuity_adv", (), dict(dt=0.05, gamma1=1.5))],
    "continuity_reduced_shiftmdot": [("continuity_adv", ("reduced_sens", "shift_mdot"), dict(dt=0.05, gamma1=1.5))],
    "continuity_shift_poisson": [("continuity_adv", ("shift_poisson",), dict(dt=0.05, gamma1=1.5))],
    "scalar_upw_limiter": [("scalar_upw_adv_diff", ("use_limiter",), dict(upw=dict(alpha=0.0, alpha_upw=1.0, ho_upwind=1.0, hybrid_factor=1.0)))],
    "scalar_mass+upw_tanh": [("scalar_mass", ("lumped_mass",), dict(dt=0.05, gamma1=1.5, gamma2=-2.0, gamma3=0.5)),
                             ("scalar_upw_adv_diff", ("peclet_tanh", "skew_symmetric", "shifted_grad_op"),
                              dict(upw=dict(alpha=0.1, alpha_upw=0.75, ho_upwind=0.5, tanh_trans=2.0, tanh_width=4.0)))],
    "scalar_mass+adv_diff": [("scalar_mass", ("lumped_mass",), dict(dt=0.05, gamma1=1.5, gamma2=-2.0, gamma3=0.5)),
                             ("scalar_adv_diff", (), {})],
}


@pytest.mark.parametrize("kind", ["shuffled", "periodic", "periodic_one_cell_wide"])
@pytest.mark.parametrize("case", sorted(_ONEDOF_LISTS))
def test_row_owner_unstructured_onedof_kernels(kind, case, tune):
    """assemble_urow.cu, one dof per node: [ScalarMass(lumped)]? + ScalarAdvDiff / ScalarDiff / ScalarUpwAdvDiff, and
    ContinuityAdv, on non-lattice connectivity (the periodic box of BASELINE config 2) in ONE launch, thread = row node, no
    atomics; against the oracle, and against the general kernel where the reference is not defined (see
    test_row_thread_unstructured_momentum_kernel)."""
    if kind == "periodic":
        m, st = _periodic_state(9, 8, 7)
    elif kind == "periodic_one_cell_wide":
        m, st = _periodic_state(1, 5, 4)
    else:
        m = perturbed_box(9, 8, 7)
        m.node_gid = (np.random.default_rng(1).permutation(m.n_nodes) + 1).astype(np.int64)
        m.box = None
        st = State(m)
    dev = st.to_device(_cuda())
    klist = _ONEDOF_LISTS[case]
    orc = oracle.LinSys([m], 1)
    keep = []
    orc.assemble_elem(0, [oracle_desc(st, k, o, keep, **sc) for k, o, sc in klist])
    ls = _linsys(m, 1)
    descs = [product_desc(ls, dev, k, o, **sc) for k, o, sc in klist]
    out = []
    for _ in range(2):
        launches0 = ls.L.nalu_b200_launch_count()
        ls.zero_system()
        assert ls.assemble_elem(descs, mode="auto") == 4
        assert ls.L.nalu_b200_launch_count() - launches0 == 1
        out.append((ls.values.clone(), ls.rhs.clone()))
    assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1])
    assert ls.geometry_errors() == 0
    if kind != "periodic_one_cell_wide":
        _compare_system(ls, orc, sol_scale=1.0 if "upw" in case else None)
    tune(urow=0)
    ls.zero_system()
    assert ls.assemble_elem(descs, mode="auto") == 1
    tune(urow=1)
    assert float((ls.values - out[0][0]).abs().max() / ls.values.abs().max()) < 1e-13
    assert float((ls.rhs - out[0][1]).abs().max() / ls.rhs.abs().max()) < 1e-12


@pytest.mark.parametrize("eq", ["tke", "sdr"])
@pytest.mark.parametrize("src_lumped", [True, False])
def test_config4_lists_on_unstructured_connectivity(eq, src_lumped, tune):
    """The turbulence lists of BASELINE config 4 [ScalarMass(lumped), ScalarUpwAdvDiff, SST source] on a mesh WITHOUT the
    lattice numbering (shuffled node ids, stretched ABL-style coordinates): one launch of the row-owner kernel
    (assemble_urow.cu, mode 4) against the oracle and against the general kernel."""
    m = perturbed_box(8, 6, 10, amp=0.05)
    f = _config4_state(m, eq)
    m.node_gid = (np.random.default_rng(3).permutation(m.n_nodes) + 1).astype(np.int64)
    m.box = None
    orc = oracle.LinSys([m], 1)
    _config4_oracle(orc, 0, m, f, eq, src_lumped)
    ls = _linsys(m, 1)
    descs = _config4_descs(ls, m, f, eq, src_lumped, True, _cuda())
    launches0 = ls.L.nalu_b200_launch_count()
    ls.zero_system()
    assert ls.assemble_elem(descs, mode="auto") == 4
    assert ls.L.nalu_b200_launch_count() - launches0 == 1
    _compare_system(ls, orc, sol_scale=3.0)
    assert ls.geometry_errors() == 0
    v0, r0 = ls.values.clone(), ls.rhs.clone()
    tune(urow=0)
    ls.zero_system()
    assert ls.assemble_elem(descs, mode="auto") == 1
    tune(urow=1)
    assert float((ls.values - v0).abs().max() / v0.abs().max()) < 1e-13
    assert float((ls.rhs - r0).abs().max() / r0.abs().max()) < 1e-12


@pytest.mark.parametrize("kind", ["box", "shuffled", "periodic"])
@pytest.mark.parametrize("variant", ["upw_limiter", "mass+upw_tanh+buoyancy", "mass+advdiff+boussinesq"])
def test_row_owner_momentum_upwind_and_buoyancy_lists(kind, variant, tune):
    """The momentum list of BASELINE config 1 in 3-D -- [MomentumMass(lumped), MomentumUpwAdvDiff, MomentumBuoyancySrc]
    (reg_tests/test_files/heatedWaterChannelElem/heatedWaterChannelElem.i:152-155) -- and its relatives through the row-owner
    kernel (mode 4) on a lattice (where no tiled kernel takes them), shuffled and periodic connectivity; against the oracle
    and against the general kernel."""
    if kind == "periodic":
        m, st = _periodic_state(9, 8, 7)
    else:
        m = perturbed_box(9, 8, 7)
        if kind == "shuffled":
            m.node_gid = (np.random.default_rng(1).permutation(m.n_nodes) + 1).astype(np.int64)
            m.box = None
        st = State(m)
    dev = st.to_device(_cuda())
    bdf2 = dict(dt=0.05, gamma1=1.5, gamma2=-2.0, gamma3=0.5)
    g = (0.3, -1.7, -9.81)
    temp = 300.0 * (1.0 + 0.01 * st.density)
    klist, extra_o, extra_p = [], [], []
    if variant == "upw_limiter":
        klist = [("momentum_upw_adv_diff", ("use_limiter",), dict(include_divu=1.0, upw=dict(alpha=0.0, alpha_upw=1.0, ho_upwind=1.0, hybrid_factor=1.0)))]
    elif variant == "mass+upw_tanh+buoyancy":
        klist = [("momentum_mass", ("lumped_mass",), bdf2),
                 ("momentum_upw_adv_diff", ("peclet_tanh", "shifted_grad_op", "skew_symmetric"),
                  dict(include_divu=1.0, upw=dict(alpha=0.1, alpha_upw=0.75, ho_upwind=0.5, tanh_trans=2.0, tanh_width=4.0)))]
        extra_o = [("momentum_buoyancy_src", dict(density=st.density, gravity=g, rhoRef=1.1))]
        extra_p = [("momentum_buoyancy_src", dict(gravity=g, rho_ref=1.1), "density", st.density)]
    else:
        klist = [("momentum_mass", ("lumped_mass",), bdf2), ("momentum_adv_diff", (), dict(include_divu=0.0))]
        extra_o = [("momentum_buoyancy_boussinesq_src", dict(temperature=temp, gravity=g, rhoRef=1.1, tRef=298.0, beta=3.4e-3))]
        extra_p = [("momentum_buoyancy_boussinesq_src", dict(gravity=g, rho_ref=1.1, t_ref=298.0, beta=3.4e-3), "temperature", temp)]
    keep = []
    orc = oracle.LinSys([m], 3)
    orc.assemble_elem(0, [oracle_desc(st, k, o, keep, **s) for k, o, s in klist] +
                      [oracle.make_desc(k, keep, coordinates=st.coordinates, **kw) for k, kw in extra_o])
    ls = _linsys(m, 3)
    descs = [product_desc(ls, dev, k, o, **s) for k, o, s in klist]
    for k, kw, fname, fval in extra_p:
        descs.append(ls.make_desc(k, fields={"coordinates": dev["coordinates"], fname: torch.as_tensor(np.ascontiguousarray(fval), device="cuda")}, **kw))
    # on a lattice a list with a tiled-capable member keeps the tiled write-once kernel for it (+ the others on top: mode 3)
    want = 3 if (kind == "box" and variant == "mass+advdiff+boussinesq") else 4
    out = []
    for _ in range(2):
        launches0 = ls.L.nalu_b200_launch_count()
        ls.zero_system()
        assert ls.assemble_elem(descs, mode="auto") == want
        assert want == 3 or ls.L.nalu_b200_launch_count() - launches0 == 1
        out.append((ls.values.clone(), ls.rhs.clone()))
    assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1])
    assert ls.geometry_errors() == 0
    _compare_system(ls, orc, sol_scale=1.0)
    tune(urow=0)
    ls.zero_system()
    assert ls.assemble_elem(descs, mode="auto") in (1, 3)
    tune(urow=1)
    assert float((ls.values - out[0][0]).abs().max() / ls.values.abs().max()) < 1e-13
    assert float((ls.rhs - out[0][1]).abs().max() / ls.rhs.abs().max()) < 1e-12
