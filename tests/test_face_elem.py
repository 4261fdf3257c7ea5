"""Face/element boundary-condition path: AssembleFaceElemSolverAlgorithm with MomentumSymmetry,
MomentumOpenAdvDiff, ScalarOpenAdv and ContinuityOpen kernels on Quad4 faces of Hex8 elements.

The reference's own tests of these kernels (unit_tests/kernels/UnitTestFaceElemBasic.C) only execute them -- they hold no
gold values -- so the oracle restatement (oracle/face_kernels.c) is anchored here by closed forms (CPU tests) and the CUDA
path is compared with the oracle on perturbed boxes (GPU tests)."""
import numpy as np
import pytest

import oracle
from nalu_b200 import mesh as M
from parity_util import State, perturbed_box, rel_err, rel_err_rows

UNIT = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], float)
OUTWARD = np.array([[0, -1, 0], [1, 0, 0], [0, 1, 0], [-1, 0, 0], [0, 0, -1], [0, 0, 1]], float)


def _skewed():
    rng = np.random.default_rng(3)
    return UNIT * np.array([1.3, 0.8, 1.1]) + 0.12 * rng.uniform(-1, 1, (8, 3))


# ----------------------------------------------------------------------------------------------------------------------
# oracle, closed forms (CPU)
# ----------------------------------------------------------------------------------------------------------------------
def test_face_tables_match_the_hex8_side_definition():
    """unit_tests/kernels/UnitTestFaceElemBasic.C:48-66 (face nodes = element nodes at the side's ordinals): the nearest
    node of face ip k is the face's node k, the opposing node is its neighbour across the element."""
    sno, ipn, opp = oracle.hex8_face_tables()
    assert np.array_equal(sno, M.HEX8_SIDE_NODES) and np.array_equal(ipn, sno)
    for f in range(6):
        axis = np.argmax(np.abs(OUTWARD[f]))
        for k in range(4):
            d = UNIT[sno[f, k]] - UNIT[opp[f, k]]
            assert np.allclose(d, OUTWARD[f]), (f, k)  # the edge opposing -> nearest points out of the element
            assert abs(d[axis]) == 1.0


def test_exposed_area_vectors_unit_cube_and_closed_surface():
    sno, _, _ = oracle.hex8_face_tables()
    for f in range(6):
        a = oracle.quad3d_scs_areav(UNIT[sno[f]])
        np.testing.assert_allclose(a, np.tile(0.25 * OUTWARD[f], (4, 1)), atol=1e-16)
    # a closed surface: the area vectors of the six (warped) faces of any hexahedron sum to zero
    c = _skewed()
    tot = sum(oracle.quad3d_scs_areav(c[sno[f]]).sum(0) for f in range(6))
    np.testing.assert_allclose(tot, 0, atol=1e-15)
    fe, fo = np.zeros(6, np.int32), np.arange(6, dtype=np.int32)
    ev = oracle.compute_exposed_area(fe, fo, np.arange(8, dtype=np.int32)[None, :], c)
    for f in range(6):
        np.testing.assert_array_equal(ev[f], oracle.quad3d_scs_areav(c[sno[f]]))


@pytest.mark.parametrize("shifted", [False, True])
def test_face_grad_op_reproduces_linear_fields(shifted):
    c = _skewed()
    g = np.array([0.7, -1.9, 0.4])
    fld = c @ g + 2.0
    for f in range(6):
        d, bad = oracle.hex8_face_grad_op(f, c, shifted)
        assert bad == 0
        np.testing.assert_allclose(np.einsum("pnj,n->pj", d, fld), np.tile(g, (4, 1)), atol=2e-14)
        np.testing.assert_allclose(d.sum(axis=1), 0, atol=1e-14)


def test_quad_shape_functions_partition_unity_and_collocate():
    s = oracle.quad3d_shape_fcn(False)
    np.testing.assert_allclose(s.sum(1), 1.0, atol=1e-16)
    np.testing.assert_allclose(np.diag(s), 0.5625)  # (3/4)^2
    np.testing.assert_array_equal(oracle.quad3d_shape_fcn(True), np.eye(4))


def _one_elem_fields(c):
    return dict(coordinates=c, velocity=np.tile([0.3, -0.2, 0.5], (8, 1)), viscosity=np.full(8, 0.1), density=np.full(8, 1.2),
                pressure=np.zeros(8), pressureBc=np.zeros(8), Gpdx=np.zeros((8, 3)), scalarQ=np.linspace(0.2, 0.9, 8),
                bcScalarQ=np.full(8, 0.33), Gjq=np.zeros((8, 3)), Gjui=np.zeros((8, 9)), velocityBc=np.zeros((8, 3)))


def test_continuity_open_uniform_flow_closed_form():
    """p = p_bc = G = 0, uniform rho u: mdot = rho u . A, rhs(nearest) = -mdot / tau; the lhs row of the nearest node sums
    to the penalty 2 aMag / L (the gradient columns of a row sum to zero)."""
    c = _skewed()
    sno, ipn, _ = oracle.hex8_face_tables()
    keep = []
    for f in range(6):
        area = oracle.quad3d_scs_areav(c[sno[f]])[None]
        d = oracle.make_face_desc("continuity_open", keep, dt=0.05, gamma1=1.5, exposedAreaVec=area, dynamicPressure=np.zeros((1, 4)),
                                  **_one_elem_fields(c))
        lhs, rhs = oracle.face_elem_lhs_rhs([d], 1, np.arange(8), f)
        tau = 0.05 / 1.5
        g, _ = oracle.hex8_face_grad_op(f, c)
        for k in range(4):
            mdot = 1.2 * np.dot([0.3, -0.2, 0.5], area[0, k])
            assert rhs[ipn[f, k]] == pytest.approx(-mdot / tau, rel=1e-13)
            invL = sum(g[k, n] @ area[0, k] for n in sno[f]) / np.linalg.norm(area[0, k])
            assert lhs[ipn[f, k]].sum() == pytest.approx(2.0 * invL * np.linalg.norm(area[0, k]), rel=1e-12)
        others = np.setdiff1d(np.arange(8), sno[f])
        assert not lhs[others].any() and not rhs[others].any()


def test_scalar_open_adv_leaving_and_entraining_closed_form():
    c = _skewed()
    sno, ipn, _ = oracle.hex8_face_tables()
    keep = []
    f = 1
    fl = _one_elem_fields(c)
    shp = oracle.quad3d_shape_fcn(False)
    for sign in (+1.0, -1.0):
        mf = sign * np.array([[0.4, 0.1, 0.25, 0.3]])
        # alphaUpw = 0: pure central value whatever the Peclet number
        d = oracle.make_face_desc("scalar_open_adv", keep, alphaUpw=0.0, hoUpwind=1.0, hybridFactor=1.0, openMassFlowRate=mf, **fl)
        lhs, rhs = oracle.face_elem_lhs_rhs([d], 1, np.arange(8), f)
        for k in range(4):
            row = ipn[f, k]
            if sign > 0:
                qip = shp[k] @ fl["scalarQ"][sno[f]]
                assert rhs[row] == pytest.approx(-mf[0, k] * qip, rel=1e-14)
                np.testing.assert_allclose(lhs[row, sno[f]], shp[k] * mf[0, k], rtol=1e-14)
            else:
                assert rhs[row] == pytest.approx(-mf[0, k] * 0.33, rel=1e-14)
                assert not lhs[row].any()
    # full upwinding (alphaUpw = 1, pecfac -> 1 at a huge Peclet number): the nearest node's own value
    fl2 = dict(fl, viscosity=np.full(8, 1e-12))
    mf = np.array([[0.4, 0.1, 0.25, 0.3]])
    d = oracle.make_face_desc("scalar_open_adv", keep, alphaUpw=1.0, hoUpwind=1.0, hybridFactor=1.0, openMassFlowRate=mf, **fl2)
    lhs, rhs = oracle.face_elem_lhs_rhs([d], 1, np.arange(8), f)
    for k in range(4):
        row = ipn[f, k]
        assert lhs[row, row] == pytest.approx(mf[0, k], rel=1e-9)
        assert rhs[row] == pytest.approx(-mf[0, k] * fl["scalarQ"][row], rel=1e-9)


def test_momentum_symmetry_plus_open_diffusion_is_the_full_traction():
    """n n^T + (I - n n^T) = I: with no open-boundary mass flow the symmetry kernel and the open kernel together apply the
    whole viscous traction -mu (grad u + grad u^T) . A to the nearest node, for any face normal."""
    c = _skewed()
    sno, ipn, _ = oracle.hex8_face_tables()
    rng = np.random.default_rng(5)
    u = rng.uniform(-1, 1, (8, 3))
    fl = dict(_one_elem_fields(c), velocity=u)
    keep = []
    for f in range(6):
        area = oracle.quad3d_scs_areav(c[sno[f]])[None]
        common = dict(includeDivU=0.0, exposedAreaVec=area, **fl)
        ds = oracle.make_face_desc("momentum_symmetry", keep, **common)
        do = oracle.make_face_desc("momentum_open_adv_diff", keep, alphaUpw=1.0, hoUpwind=1.0, hybridFactor=1.0,
                                   openMassFlowRate=np.zeros((1, 4)), **common)
        lhs, rhs = oracle.face_elem_lhs_rhs([ds, do], 3, np.arange(8), f)
        g, _ = oracle.hex8_face_grad_op(f, c)
        for k in range(4):
            A = area[0, k]
            ref = np.zeros((3, 24))
            for ic in range(8):
                for i in range(3):
                    ref[i, 3 * ic + i] += -0.1 * (g[k, ic] @ A)
                    for j in range(3):
                        ref[i, 3 * ic + j] += -0.1 * g[k, ic, i] * A[j]
            rows = 3 * ipn[f, k] + np.arange(3)
            np.testing.assert_allclose(lhs[rows], ref, atol=1e-14)
            np.testing.assert_allclose(rhs[rows], -ref @ u.ravel(), atol=1e-14)


def test_momentum_wall_function_closed_form():
    """flat wall (bottom face of the unit cube), uniform velocity U with a wall-normal part, wall at rest: every face point
    pulls its nearest node with -lambda (I - n n^T) U, lambda = mu / y_p |A| below y+_crit and rho kappa u_tau / ln(E y+) |A|
    above; the lhs rows are lambda (I - n n^T) times the face shape functions, on the face's nodes only
    (src/kernel/MomentumWallFunctionElemKernel.C:73-155)"""
    c = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], dtype=np.float64)
    sno, ipn, _ = oracle.hex8_face_tables()
    f = 4  # ordinal 4: nodes 0 3 2 1, outward normal -z
    area = oracle.quad3d_scs_areav(c[sno[f]])[None]
    assert np.allclose(area[0], [[0, 0, -0.25]] * 4)
    U = np.array([0.3, -0.2, 0.5])
    rho, mu, kappa, elog, ycrit = 1.2, 0.1, 0.41, 9.8, 11.63
    shp = oracle.quad3d_shape_fcn()
    for utau, yp in ((0.2, 0.05), (4.0, 0.5)):
        keep = []
        fld = _one_elem_fields(c)
        d = oracle.make_face_desc("momentum_wall_function", keep, exposedAreaVec=area, wallVelocityBc=np.zeros((8, 3)),
                                  wallFrictionVelocityBip=np.full((1, 4), utau), wallNormalDistanceBip=np.full((1, 4), yp),
                                  elog=elog, kappa=kappa, yplusCrit=ycrit, **fld)
        lhs, rhs = oracle.face_elem_lhs_rhs([d], 3, np.arange(8), f)
        yplus = rho * yp * utau / mu
        lam = (rho * kappa * utau / np.log(elog * yplus) if yplus > ycrit else mu / yp) * 0.25
        assert (yplus > ycrit) == (utau > 1.0)
        P = np.diag([1.0, 1.0, 0.0])  # I - n n^T for n = -z
        L = lhs.reshape(8, 3, 8, 3)
        for k in range(4):
            node = sno[f][k]
            assert np.allclose(rhs.reshape(8, 3)[node], -lam * P @ U, rtol=0, atol=1e-15)
            for ic in range(4):
                assert np.allclose(L[node, :, sno[f][ic], :], lam * P * shp[k][ic], rtol=0, atol=1e-15)
        others = np.setdiff1d(np.arange(8), sno[f])
        assert not L[others].any() and not L[:, :, others, :].any()
        # total wall force = -lambda_total (tangential velocity): the four points cover the whole face
        assert np.allclose(rhs.reshape(8, 3).sum(axis=0), -4 * lam * P @ U)


def test_exposed_faces_of_a_box():
    m = M.generated_box(4, 3, 2)
    fe, fo = M.exposed_faces(m)
    assert fe.size == 2 * (4 * 3 + 3 * 2 + 4 * 2)
    got = set(zip(fe.tolist(), fo.tolist()))
    want = set()
    for s in "xXyYzZ":
        e, o = M.box_side_faces(m, s)
        want |= set(zip(e.tolist(), o.tolist()))
    assert got == want
    # outward area vectors: they sum to zero over the closed skin
    ev = oracle.compute_exposed_area(fe, fo, m.elem_nodes, m.coords)
    np.testing.assert_allclose(ev.sum((0, 1)), 0, atol=1e-13)
    assert ev.reshape(-1, 3)[fo.repeat(4) == 5][:, 2].min() > 0


# ----------------------------------------------------------------------------------------------------------------------
# CUDA path vs oracle (GPU)
# ----------------------------------------------------------------------------------------------------------------------
torch = pytest.importorskip("torch")


def _face_state(m, fe, fo, seed=2):
    st = State(m)
    c = m.coords
    rng = np.random.default_rng(seed)
    a = 0.3 * np.pi
    nf = fe.size
    area = oracle.compute_exposed_area(fe, fo, m.elem_nodes, c)
    fld = dict(coordinates=c, velocity=st.velocity, velocityRTM=st.velocity - 0.1 * np.cos(c), meshVelocity=0.1 * np.cos(c),
               viscosity=st.viscosity, density=st.density, pressure=st.pressure, pressureBc=0.9 * st.pressure + 0.05,
               Gpdx=st.dpdx, scalarQ=st.scalarq, bcScalarQ=0.5 + 0.2 * np.sin(a * c[:, 1]), Gjq=st.gjq, Gjui=st.gju,
               velocityBc=0.5 * st.velocity + 0.02, wallVelocityBc=0.1 * st.velocity - 0.03)
    # open mass flow of both signs on every side; a dynamic pressure
    mf = (area * (st.density[:, None] * st.velocity)[m.elem_nodes[fe][np.arange(nf)[:, None], M.HEX8_SIDE_NODES[fo]]]).sum(2)
    mf += 0.05 * rng.uniform(-1, 1, mf.shape) * np.abs(mf).max()
    # wall distances / friction velocities that put face points on both sides of y+_crit
    ff = dict(exposedAreaVec=area, openMassFlowRate=mf, dynamicPressure=0.01 * rng.uniform(0, 1, (nf, 4)),
              wallFrictionVelocityBip=rng.uniform(0.05, 2.0, (nf, 4)), wallNormalDistanceBip=rng.uniform(0.02, 0.3, (nf, 4)))
    return fld, ff


_PN = dict(coordinates="coordinates", velocity="velocity", velocityRTM="velocity_rtm", meshVelocity="mesh_velocity",
           viscosity="viscosity", density="density", pressure="pressure", pressureBc="pressure_bc", Gpdx="dpdx", scalarQ="scalarq",
           bcScalarQ="bc_scalarq", Gjq="gjq", Gjui="gjui", velocityBc="velocity_bc", wallVelocityBc="wall_velocity_bc")
_PF = dict(exposedAreaVec="exposed_area_vector", openMassFlowRate="open_mass_flow_rate", dynamicPressure="dynamic_pressure",
           wallFrictionVelocityBip="wall_friction_velocity_bip", wallNormalDistanceBip="wall_normal_distance_bip")
_WALL = dict(elog=9.8, kappa=0.41, yplusCrit=11.63)   # SolutionOptions defaults of TM_elog, TM_kappa, TM_yplus_crit
_WALL_ABI = dict(elog=9.8, kappa=0.41, yplus_crit=11.63)
_UPW = dict(alphaUpw=0.75, hoUpwind=0.9, hybridFactor=1.0, tanhTrans=2.0, tanhWidth=4.0)
_UPW_ABI = dict(alpha_upw=0.75, ho_upwind=0.9, hybrid_factor=1.0, tanh_trans=2.0, tanh_width=4.0)

FACE_CASES = {
    "momentum_symmetry": (3, [("momentum_symmetry", dict(includeDivU=1.0), ())]),
    "momentum_symmetry_shifted": (3, [("momentum_symmetry", dict(includeDivU=0.0, shiftedGradOp=1), ("shifted_grad_op",))]),
    "momentum_open": (3, [("momentum_open_adv_diff", dict(includeDivU=1.0, **_UPW), ())]),
    "momentum_open_moving_tanh": (3, [("momentum_open_adv_diff", dict(includeDivU=0.0, meshMoves=1, pecletTanh=1, skewSymmetric=1,
                                                                        shiftedGradOp=1, **_UPW),
                                       ("mesh_moves", "peclet_tanh", "skew_symmetric", "shifted_grad_op"))]),
    "momentum_symmetry+open": (3, [("momentum_symmetry", dict(includeDivU=1.0), ()),
                                   ("momentum_open_adv_diff", dict(includeDivU=1.0, **_UPW), ())]),
    "momentum_wall_function": (3, [("momentum_wall_function", dict(**_WALL), ())]),
    "momentum_symmetry+wall_function": (3, [("momentum_symmetry", dict(includeDivU=1.0), ()),
                                            ("momentum_wall_function", dict(**_WALL), ())]),
    "scalar_open": (1, [("scalar_open_adv", dict(**_UPW), ())]),
    "scalar_open_skew_tanh": (1, [("scalar_open_adv", dict(skewSymmetric=1, pecletTanh=1, **_UPW), ("skew_symmetric", "peclet_tanh"))]),
    "continuity_open": (1, [("continuity_open", dict(dt=0.05, gamma1=1.5), ())]),
    "continuity_open_reduced_shiftmdot": (1, [("continuity_open", dict(dt=0.05, gamma1=1.5, reducedSensitivities=1, shiftMdot=1),
                                               ("reduced_sens", "shift_mdot"))]),
    "continuity_open_shifted": (1, [("continuity_open", dict(dt=0.05, gamma1=1.5, shiftedGradOp=1), ("shifted_grad_op",))]),
}


def _product_face_descs(ls, klist, fld, ff, dev):
    tt = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
    dn = {_PN[k]: tt(v) for k, v in fld.items()}
    df = {_PF[k]: tt(v) for k, v in ff.items()}
    out = []
    for kind, sc, opts in klist:
        if "mesh_moves" not in opts and kind == "momentum_open_adv_diff":
            f = {k: v for k, v in dn.items() if k not in ("velocity_rtm", "mesh_velocity")}
        else:
            f = dn
        out.append(ls.make_face_desc(kind, fields=f, face_fields=df, options=opts, include_divu=sc.get("includeDivU", 0.0),
                                     dt=sc.get("dt", 1.0), gamma1=sc.get("gamma1", 1.0), upw=_UPW_ABI, wall=_WALL_ABI))
    return out


def _compare(ls, orc, rank=0):
    v, r = ls.owned_values().cpu().numpy(), ls.owned_rhs().cpu().numpy()
    ov, orr = orc.values(rank), orc.rhs(rank)
    assert np.abs(ov).max() > 0 and np.abs(orr).max() > 0
    ev = rel_err_rows(v, ov, orc.rowptr(rank), 1e-2)
    rp = orc.rowptr(rank).astype(np.int64)
    # residual entries: sums of |a_ij u_j|-sized terms (see tests/sampled_parity.py)
    er = rel_err(r, orr, 1e-2 * max(np.abs(orr).max(), float(np.add.reduceat(np.abs(ov), rp[:-1]).max())))
    assert ev < 1e-12, f"values rel err {ev:.3e}"
    assert er < 1e-12, f"rhs rel err {er:.3e}"


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(FACE_CASES))
def test_face_elem_algorithm_vs_oracle(case):
    from nalu_b200.linsys import DeviceCsrLinearSystem
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required: the assembly path has no CPU fallback")
    ndof, klist = FACE_CASES[case]
    m = perturbed_box(6, 5, 4)
    fe, fo = M.exposed_faces(m)
    fld, ff = _face_state(m, fe, fo)
    keep = []
    orc = oracle.LinSys([m], ndof)
    orc.assemble_face_elem(0, [oracle.make_face_desc(k, keep, **sc, **fld, **ff) for k, sc, _ in klist], fe, fo)
    ls = DeviceCsrLinearSystem(m, ndof).allocate("cuda")
    fs = ls.face_set(fe, fo)
    assert 4 <= fs.num_colors() <= 12
    # exposed_area_vector on the device (ComputeGeometryBoundaryAlgorithm)
    area = fs.exposed_area(torch.as_tensor(m.coords, device="cuda"))
    np.testing.assert_allclose(area.cpu().numpy(), ff["exposedAreaVec"], rtol=0, atol=1e-15)
    descs = _product_face_descs(ls, klist, fld, ff, "cuda")
    ls.zero_system()
    ls.assemble_face_elem(fs, descs)
    torch.cuda.synchronize()
    _compare(ls, orc)
    # deterministic: colour by colour, plain read-modify-writes
    v0, r0 = ls.owned_values().clone(), ls.owned_rhs().clone()
    ls.zero_system()
    ls.assemble_face_elem(fs, descs)
    assert torch.equal(v0, ls.owned_values()) and torch.equal(r0, ls.owned_rhs())


@pytest.mark.gpu
def test_face_elem_on_top_of_interior_assembly_and_slab_ranks():
    """The call sequence of a time step: zeroSystem -> interior algorithm (tiled write-once) -> boundary face algorithm ->
    loadComplete, on two z-slab ranks (open side X on both ranks, symmetry side z on the lower rank), vs the oracle."""
    from test_gpu_parity import _build_ranks
    from parity_util import oracle_desc, product_desc
    full = perturbed_box(7, 5, 6)
    parts = M.partition_slabs(full, 2)
    systems = _build_ranks(parts, 3)
    orc = oracle.LinSys(parts, 3)
    keep = []
    for r, (m, ls) in enumerate(zip(parts, systems)):
        st = State(m)
        dev = st.to_device("cuda")
        orc.assemble_elem(r, [oracle_desc(st, "momentum_adv_diff", (), keep, include_divu=1.0)])
        assert ls.assemble_elem([product_desc(ls, dev, "momentum_adv_diff", include_divu=1.0)], overwrite=True) == 2
        for side, kind, sc in (("X", "momentum_open_adv_diff", dict(includeDivU=1.0, **_UPW)), ("z", "momentum_symmetry", dict(includeDivU=1.0))):
            fe, fo = M.box_side_faces(m, side)
            if fe.size == 0:
                assert side == "z" and r == 1
                continue
            fld, ff = _face_state(m, fe, fo, seed=5 + r)
            orc.assemble_face_elem(r, [oracle.make_face_desc(kind, keep, **sc, **fld, **ff)], fe, fo)
            fs = ls.face_set(fe, fo)
            ls.assemble_face_elem(fs, _product_face_descs(ls, [(kind, sc, ())], fld, ff, "cuda"))
    torch.cuda.synchronize()
    for ls in systems:
        _, recv = ls.halo_peers()
        for i, (peer, n) in enumerate(recv):
            ls.halo_pull_add(i, systems[peer].shared_values().data_ptr(), systems[peer].shared_rhs().data_ptr())
    torch.cuda.synchronize()
    orc.load_complete()
    for r, ls in enumerate(systems):
        _compare(ls, orc, rank=r)


@pytest.mark.gpu
def test_face_elem_missing_field_is_an_argument_error():
    from nalu_b200.linsys import DeviceCsrLinearSystem
    m = perturbed_box(3, 3, 2)
    fe, fo = M.exposed_faces(m)
    fld, ff = _face_state(m, fe, fo)
    ls = DeviceCsrLinearSystem(m, 1).allocate("cuda")
    fs = ls.face_set(fe, fo)
    del fld["bcScalarQ"]
    ls.zero_system()
    with pytest.raises(RuntimeError, match="bcScalarQ"):
        ls.assemble_face_elem(fs, _product_face_descs(ls, [("scalar_open_adv", {}, ())], fld, ff, "cuda"))
    with pytest.raises(RuntimeError, match="numDof"):
        ls.assemble_face_elem(fs, _product_face_descs(ls, [("momentum_symmetry", {}, ())], fld, ff, "cuda"))
