"""Hex27 (P2 CVFEM, BASELINE config 5): oracle pinned by the reference's own Hex27 tests (they are property tests: the
reference holds no Hex27 gold table), graph on Hex27 connectivity, and the CUDA kernel against the oracle.

Restated reference tests:
  unit_tests/UnitTestHexMasterElements.C:297-321  hex27_scs_interpolation / hex27_scv_interpolation / hex27_scs_derivatives
  unit_tests/UnitTestMasterElements.C:185-229     volume_integration
"""
import numpy as np
import pytest

from nalu_b200 import mesh as M
from oracle import hex27 as H

# unit_tests/UnitTestUtils.C:218-233 create_one_reference_hex27_element
REF_NODES = np.array([
    [-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1],
    [0, -1, -1], [1, 0, -1], [0, 1, -1], [-1, 0, -1], [-1, -1, 0], [1, -1, 0], [1, 1, 0], [-1, 1, 0],
    [0, -1, 1], [1, 0, 1], [0, 1, 1], [-1, 0, 1], [0, 0, 0], [0, 0, -1], [0, 0, 1], [-1, 0, 0], [1, 0, 0], [0, -1, 0],
    [0, 1, 0]], dtype=np.float64)


def _poly(rng, order=2):
    return rng.uniform(-1.0, 1.0, size=(3, order + 1))


def _poly_val(c, x):
    x = np.atleast_2d(x)
    out = np.ones(x.shape[0])
    for d in range(3):
        out = out * sum(c[d][j] * x[:, d] ** j for j in range(c.shape[1]))
    return out


def _poly_der(c, x, direction):
    x = np.atleast_2d(x)
    out = np.ones(x.shape[0])
    for d in range(3):
        if d == direction:
            out = out * sum(j * c[d][j] * x[:, d] ** (j - 1) for j in range(1, c.shape[1]))
        else:
            out = out * sum(c[d][j] * x[:, d] ** j for j in range(c.shape[1]))
    return out


def _perturbed_box(n, amp=0.08, seed=3):
    m = M.hex27_box(*n)
    rng = np.random.default_rng(seed)
    m.coords = m.coords + amp * rng.uniform(-1.0, 1.0, size=m.coords.shape)  # every node: curved (truly quadratic) elements
    return m


def _fields(m, seed=5):
    rng = np.random.default_rng(seed)
    x = m.coords
    vel = np.stack([np.sin(0.9 * x[:, 1]) + 0.3 * x[:, 2], np.cos(0.7 * x[:, 0]) * x[:, 2], 0.2 * x[:, 0] * x[:, 1] - 0.5], axis=1)
    vel += 0.05 * rng.uniform(-1, 1, size=vel.shape)
    visc = 0.3 + 0.1 * np.cos(x[:, 0] + 2.0 * x[:, 1]) + 0.02 * rng.uniform(-1, 1, size=x.shape[0])
    rho = 1.0 + 0.2 * np.sin(x[:, 2]) + 0.02 * rng.uniform(-1, 1, size=x.shape[0])
    return vel, visc, rho


# ---- oracle pinning (CPU) ------------------------------------------------------------------------------------------------

def test_shape_functions_are_nodal_and_partition_unity():
    f = H.hex27_shape_fcn(REF_NODES)
    assert np.allclose(f, np.eye(27), atol=1e-15)  # the literal formulas are the Lagrange basis of REF_NODES in STK order
    rng = np.random.default_rng(0)
    p = rng.uniform(-1, 1, size=(50, 3))
    assert np.allclose(H.hex27_shape_fcn(p).sum(axis=1), 1.0, atol=1e-14)
    assert np.allclose(H.hex27_shape_deriv(p).sum(axis=1), 0.0, atol=1e-13)


def test_shape_deriv_is_the_derivative_of_the_literal_shape_functions():
    rng = np.random.default_rng(1)
    p = rng.uniform(-1, 1, size=(40, 3))
    d = H.hex27_shape_deriv(p)
    h = 1e-6
    for k in range(3):
        e = np.zeros(3)
        e[k] = h
        fd = (H.hex27_shape_fcn(p + e) - H.hex27_shape_fcn(p - e)) / (2 * h)
        assert np.abs(fd - d[:, :, k]).max() < 1e-9


@pytest.mark.parametrize("me", ["scs", "scv"])
def test_hex27_interpolation(me):
    """hex27_scs_interpolation / hex27_scv_interpolation: a random quadratic (per direction) polynomial is interpolated
    exactly to the integration points"""
    m = H.Hex27SCS() if me == "scs" else H.Hex27SCV()
    c = _poly(np.random.default_rng(0))
    nodal = _poly_val(c, REF_NODES)
    assert np.abs(m.shape_fcn @ nodal - _poly_val(c, m.intg_loc)).max() < 1e-12


def test_hex27_scs_derivatives():
    """hex27_scs_derivatives: grad_op on the reference element = shape derivatives, det = 1, polynomial derivative exact"""
    m = H.Hex27SCS()
    dndx, det = m.grad_op(REF_NODES[None])
    assert np.abs(dndx[0] - m.shape_deriv).max() < 1e-12
    assert np.abs(det - 1.0).max() < 1e-12
    c = _poly(np.random.default_rng(0))
    nodal = _poly_val(c, REF_NODES)
    for d in range(3):
        assert np.abs(dndx[0, :, :, d] @ nodal - _poly_der(c, m.intg_loc, d)).max() < 1e-12


def test_hex27_volume_integration():
    """volume_integration: SCV weights scale with det of a linear map; and they sum to the element volume"""
    m = H.Hex27SCV()
    rng = np.random.default_rng(0)
    q, r = np.linalg.qr(rng.uniform(-1, 1, size=(3, 3)))
    qr = q @ (np.eye(3) + 0.3 * np.triu(rng.uniform(-1, 1, size=(3, 3))))
    if np.linalg.det(qr) < 0:
        qr[:, 0] *= -1
    v0 = m.volumes(REF_NODES[None])[0]
    v1 = m.volumes((REF_NODES @ qr.T)[None])[0]
    assert np.abs(np.linalg.det(qr) * v0 - v1).max() < 1e-13
    assert abs(v0.sum() - 8.0) < 1e-13
    # every node owns 8 points; 1-D dual lengths: 1 - 1/sqrt(3) (end nodes), 2/sqrt(3) (mid node)
    per_node = np.bincount(m.ip_node_map, weights=v0, minlength=27)
    l = np.array([1 - H.SCS_DIST, 2 * H.SCS_DIST, 1 - H.SCS_DIST])
    expect = np.empty(27)
    for cc in range(3):
        for b in range(3):
            for a in range(3):
                expect[H.tensor_product_node_map(a, b, cc)] = l[a] * l[b] * l[cc]
    assert np.abs(per_node - expect).max() < 1e-14


def test_scs_area_vectors_closed_forms():
    m = H.Hex27SCS()
    a = m.area_vectors(REF_NODES[None])[0]
    # a constant-u surface of the reference cube has area 4, normal +-z; the vector points from the left to the right node
    for direction, axis in ((H.U_DIRECTION, 2), (H.T_DIRECTION, 1), (H.S_DIRECTION, 0)):
        ips = np.nonzero(m.direction == direction)[0]
        assert np.allclose(np.delete(a[ips], axis, axis=1), 0.0, atol=1e-15)
        assert abs(np.abs(a[ips, axis]).sum() - 8.0) < 1e-13  # two surfaces of area 4
        lr = REF_NODES[m.lrscv[ips, 1]] - REF_NODES[m.lrscv[ips, 0]]
        assert np.all(np.sign(lr[:, axis]) == np.sign(a[ips, axis]))
    # on a curved element every interior sub-control volume is closed: sum of outward area vectors of the centre node's
    # control volume vanishes (all its faces are sub-control surfaces of this element)
    rng = np.random.default_rng(2)
    xc = REF_NODES + 0.1 * rng.uniform(-1, 1, size=REF_NODES.shape)
    ac = m.area_vectors(xc[None])[0]
    out = ac[m.lrscv[:, 0] == 20].sum(axis=0) - ac[m.lrscv[:, 1] == 20].sum(axis=0)
    assert np.abs(out).max() < 1e-13


def test_hex27_with_centred_mid_nodes_reproduces_trilinear_geometry():
    """mid-edge / face / body nodes at the averages of a perturbed Hex8's corners: the quadratic map IS the trilinear one,
    so Jacobians at the SCS points agree with the Hex8 shape derivatives evaluated there"""
    rng = np.random.default_rng(4)
    x = REF_NODES.copy()
    x[:8] += 0.15 * rng.uniform(-1, 1, size=(8, 3))
    corner_ref = REF_NODES[:8]
    def tri(p):  # trilinear shape functions of the 8 corners at p
        return np.prod(1.0 + corner_ref[None, :, :] * p[:, None, :], axis=2) / 8.0
    x[8:] = tri(REF_NODES[8:]) @ x[:8]
    m = H.Hex27SCS()
    jact = np.einsum("pnj,ni->pij", m.shape_deriv, x)
    h = 1e-6
    for k in range(3):
        e = np.zeros(3)
        e[k] = h
        fd = (tri(m.intg_loc + e) - tri(m.intg_loc - e)) @ x[:8] / (2 * h)
        assert np.abs(fd - jact[:, :, k]).max() < 1e-9


def test_momentum_adv_diff_identities():
    """row sums of the diffusion + stress operator vanish (constants are in its null space), and a uniform velocity with
    consistent mdot gives a residual = -(net advective flux) that sums to zero over an element"""
    m = _perturbed_box((1, 1, 1))
    vel, visc, rho = _fields(m)
    mdot0 = np.zeros((1, 216))
    lhs, rhs = H.momentum_adv_diff(m.elem_nodes, m.coords, vel, visc, mdot0, include_divu=1.0)
    L = lhs.reshape(81, 27, 3)
    assert np.abs(L.sum(axis=1)).max() < 1e-12 * np.abs(L).max() * 27
    assert np.abs(rhs.reshape(27, 3).sum(axis=0)).max() < 1e-12  # fluxes telescope: every point adds to il what it takes from ir
    # uniform velocity: no viscous force, no divergence
    u0 = np.tile([0.3, -0.2, 0.7], (m.coords.shape[0], 1))
    mdot = H.calc_mass_flow_rate_scs(m.elem_nodes, m.coords, np.ones(m.coords.shape[0]), u0)
    lhs, rhs = H.momentum_adv_diff(m.elem_nodes, m.coords, u0, visc, mdot, include_divu=1.0)
    # rhs(node, i) = -u0_i * (net mass flux out of the node's sub-control volume through interior surfaces)
    me = H.scs()
    net = np.zeros(27)
    np.add.at(net, me.lrscv[:, 0], mdot[0])
    np.add.at(net, me.lrscv[:, 1], -mdot[0])
    assert np.abs(rhs.reshape(27, 3) + net[:, None] * u0[0][None, :]).max() < 1e-13


def test_scv_shifted_points_are_nodal_and_lumped_mass_closed_form():
    """scv_shifted_ips_are_nodal (unit_tests/UnitTestMasterElements.C:161-183, run by the reference for P1 elements only; the
    Hex27 tables have the same property) and the lumped mass of a unit-density box: diagonal = gamma1 * dual volume / dt"""
    me = H.scv()
    assert np.array_equal(me.shape_fcn_shift, np.eye(27)[me.ip_node_map])
    m = M.hex27_box(1, 1, 1)
    n = m.n_nodes
    one, zero3 = np.ones(n), np.zeros((n, 3))
    u = np.tile([1.0, 2.0, 3.0], (n, 1))
    lhs, rhs = H.momentum_mass(m.elem_nodes, m.coords, u, u, u, one, one, one, zero3, dt=0.5, gamma1=1.5, gamma2=-2.0, gamma3=0.5)
    vol = np.bincount(me.ip_node_map, weights=me.volumes(m.coords[m.elem_nodes])[0], minlength=27)
    assert abs(vol.sum() - 1.0) < 1e-14
    assert np.allclose(np.diag(lhs[0]), np.repeat(1.5 * vol / 0.5, 3), rtol=0, atol=1e-15)
    assert np.abs(lhs[0] - np.diag(np.diag(lhs[0]))).max() == 0.0
    assert np.abs(rhs).max() < 1e-14  # gamma1 + gamma2 + gamma3 = 0 on a steady state


def test_hex27_box_mesh():
    m = M.hex27_box(2, 3, 2)
    assert m.elem_nodes.shape == (12, 27) and m.coords.shape == (5 * 7 * 5, 3)
    # element 0 in STK order = the reference element scaled by 1/2 and shifted
    assert np.allclose(m.coords[m.elem_nodes[0]], 0.5 * (REF_NODES + 1.0))
    v = H.Hex27SCV().volumes(m.coords[m.elem_nodes])
    assert np.allclose(v.sum(axis=1), 1.0)


def test_hex27_graph_host_only_matches_oracle_csr():
    from nalu_b200.linsys import DeviceCsrLinearSystem
    m = M.hex27_box(3, 2, 2)
    ls = DeviceCsrLinearSystem(m, 3, host_only=True)
    E = m.n_elems
    rowptr, colidx, _, _ = H.assemble_csr(m.node_gid, m.elem_nodes, np.zeros((E, 81, 81)), np.zeros((E, 81)))
    assert ls.n_owned_rows == 3 * m.n_nodes and ls.n_shared_rows == 0
    assert np.array_equal(ls.rowptr().astype(np.int64), rowptr)
    assert np.array_equal(ls.colidx().astype(np.int64), colidx)
    nc, col = ls.elem_colors()
    assert nc == 8
    # elements of one colour share no node
    for c in range(nc):
        nodes = m.elem_nodes[col == c].ravel()
        assert np.unique(nodes).size == nodes.size
    # the widest row: a corner node shared by 8 elements couples to 5^3 nodes
    assert int(np.diff(rowptr).max()) == 125 * 3


def test_hex8_entry_points_refuse_hex27_graphs():
    import ctypes as C
    from nalu_b200 import lib as L
    from nalu_b200.linsys import DeviceCsrLinearSystem
    m = M.hex27_box(1, 1, 1)
    ls = DeviceCsrLinearSystem(m, 3, host_only=True)
    buf = np.zeros(8)
    # the face/element boundary algorithms exist for Quad4 faces of Hex8 elements only
    fs = C.c_void_p()
    fe, fo = np.zeros(1, dtype=np.int32), np.zeros(1, dtype=np.int32)
    rc = ls.L.nalu_b200_face_set_create(C.byref(fs), ls.h, C.c_int64(1), L.ptr(fe), L.ptr(fo))
    assert rc != 0


def _global_csr_lookup(m):
    """sorted (row dof gid, col dof gid) keys of the one-rank Hex27 CSR, GID_ = 3 (id - 1) + i + 1"""
    E = m.n_elems
    rowptr, colidx, _, _ = H.assemble_csr(m.node_gid, m.elem_nodes, np.zeros((E, 81, 81)), np.zeros((E, 81)))
    n = rowptr.size - 1
    rows = np.repeat(np.arange(n), np.diff(rowptr))
    return (rows + 1) * (n + 1) + (colidx + 1), n


def _rank_keys(ls, n, shared=False):
    rp, ci = ls.rowptr(shared).astype(np.int64), ls.colidx(shared).astype(np.int64)
    rg, cg = ls.row_gids(shared), ls.col_gids()
    rows = np.repeat(rg, np.diff(rp))
    return rows * (n + 1) + cg[ci]


@pytest.mark.parametrize("nranks", [2, 3])
def test_hex27_graph_on_partitioned_mesh(nranks):
    """owned rows of every rank = the rows of the one-rank graph (same columns, by global id); shared-not-owned rows hold
    what the rank's own elements touch (src/TpetraLinearSystem.C:190-327, 678-738 on 27-node elements)"""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_graph import build_product_systems
    m = M.hex27_box(3, 2, 3)
    rng = np.random.default_rng(11)
    parts = M.partition_elements(m, rng.integers(0, nranks, size=m.n_elems), nranks)
    systems = build_product_systems(parts, 3)
    keys, n = _global_csr_lookup(m)
    seen = []
    for ls in systems:
        k = _rank_keys(ls, n)
        assert np.all(np.diff(k.reshape(-1)) != 0)
        assert np.isin(k, keys).all()
        seen.append(k)
        ks = _rank_keys(ls, n, shared=True)
        assert np.isin(ks, keys).all()
    allk = np.sort(np.concatenate(seen))
    assert np.array_equal(allk, keys)  # every row of the global graph is owned exactly once, with all its columns


# ---- CUDA kernel vs oracle (GPU) -----------------------------------------------------------------------------------------

def _device_assemble(m, vel, visc, rho, include_divu):
    import torch
    from nalu_b200.linsys import DeviceCsrLinearSystem
    dev = torch.device("cuda:0")
    ls = DeviceCsrLinearSystem(m, 3)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
    coords, v, mu, r = t(m.coords), t(vel), t(visc), t(rho)
    mdot = ls.mdot_scs(coords, r, v)
    ls.allocate(dev)
    ls.zero_system()
    d = ls.make_desc("momentum_adv_diff", fields=dict(coordinates=coords, velocity=v, viscosity=mu), mass_flow_rate=mdot,
                     include_divu=include_divu)
    mode = ls.assemble_elem([d])
    torch.cuda.synchronize()
    return ls, mode, mdot.cpu().numpy()


def _rel_err_rows(a, b, rowptr):
    rowmax = np.maximum.reduceat(np.maximum(np.abs(a), np.abs(b)), rowptr[:-1])
    floor = np.repeat(rowmax, np.diff(rowptr)) * 1e-2
    return float((np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), floor + 1e-300)).max())


@pytest.mark.gpu
@pytest.mark.parametrize("shape,include_divu", [((1, 1, 1), 1.0), ((3, 2, 2), 1.0), ((2, 3, 4), 0.0), ((4, 4, 3), 1.0)])
def test_hex27_momentum_adv_diff_vs_oracle(shape, include_divu):
    m = _perturbed_box(shape)
    vel, visc, rho = _fields(m)
    ls, mode, mdot_dev = _device_assemble(m, vel, visc, rho, include_divu)
    assert mode == 5
    assert ls.geometry_errors() == 0
    mdot = H.calc_mass_flow_rate_scs(m.elem_nodes, m.coords, rho, vel)
    assert np.abs(mdot_dev - mdot).max() <= 1e-12 * np.abs(mdot).max()
    lhs, rhs = H.momentum_adv_diff(m.elem_nodes, m.coords, vel, visc, mdot, include_divu=include_divu)
    rowptr, colidx, values, b = H.assemble_csr(m.node_gid, m.elem_nodes, lhs, rhs)
    got_v = ls.owned_values().cpu().numpy()
    got_b = ls.owned_rhs().cpu().numpy()
    assert got_v.shape == values.shape
    # tolerance: 1e-12 relative, floor 1 % of the largest entry of the same row
    assert _rel_err_rows(got_v, values, rowptr) < 1e-12
    scale = np.abs(values).max() * np.abs(vel).max()
    assert np.abs(got_b - b).max() < 1e-12 * scale * 27


@pytest.mark.gpu
@pytest.mark.parametrize("order", ["mass_first", "advdiff_first"])
def test_hex27_momentum_list_with_fused_lumped_mass_vs_oracle(order):
    """[MomentumMassElemKernel(lumped), MomentumAdvDiffElemKernel] (the low-Mach momentum list, src/LowMachEquationSystem.C:1304-1398)
    on Hex27 in one launch family"""
    import torch
    from nalu_b200.linsys import DeviceCsrLinearSystem
    m = _perturbed_box((3, 2, 3))
    vel, visc, rho = _fields(m)
    rng = np.random.default_rng(9)
    vn, vm = 0.9 * vel + 0.01, 0.8 * vel - 0.02
    rn, rm = 0.98 * rho, 0.95 * rho
    gp = 0.3 * rng.uniform(-1, 1, size=vel.shape)
    bdf2 = dict(dt=0.05, gamma1=1.5, gamma2=-2.0, gamma3=0.5)
    dev = torch.device("cuda:0")
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
    ls = DeviceCsrLinearSystem(m, 3).allocate(dev)
    f = dict(coordinates=t(m.coords), velocity=t(vel), velocity_n=t(vn), velocity_nm1=t(vm), viscosity=t(visc), density=t(rho),
             density_n=t(rn), density_nm1=t(rm), dpdx=t(gp))
    mdot_dev = ls.mdot_scs(f["coordinates"], f["density"], f["velocity"])
    descs = [ls.make_desc("momentum_mass", fields=f, options=("lumped_mass",), **bdf2),
             ls.make_desc("momentum_adv_diff", fields=f, mass_flow_rate=mdot_dev, include_divu=1.0)]
    if order == "advdiff_first":
        descs = descs[::-1]
    ls.zero_system()
    assert ls.assemble_elem(descs) == 5
    torch.cuda.synchronize()
    assert ls.geometry_errors() == 0
    mdot = H.calc_mass_flow_rate_scs(m.elem_nodes, m.coords, rho, vel)
    l1, r1 = H.momentum_adv_diff(m.elem_nodes, m.coords, vel, visc, mdot, include_divu=1.0)
    l2, r2 = H.momentum_mass(m.elem_nodes, m.coords, vel, vn, vm, rho, rn, rm, gp, lumped=True, **bdf2)
    rowptr, colidx, values, b = H.assemble_csr(m.node_gid, m.elem_nodes, l1 + l2, r1 + r2)
    got_v, got_b = ls.owned_values().cpu().numpy(), ls.owned_rhs().cpu().numpy()
    assert _rel_err_rows(got_v, values, rowptr) < 1e-12
    assert np.abs(got_b - b).max() < 1e-12 * max(np.abs(b).max(), np.abs(values).max() * np.abs(vel).max())
    # a consistent-mass list is refused, not silently mis-assembled
    bad = [ls.make_desc("momentum_mass", fields=f, **bdf2), descs[0] if order == "advdiff_first" else descs[1]]
    with pytest.raises(Exception):
        ls.assemble_elem(bad)


def test_one_dof_oracle_identities():
    """continuity: the Laplacian-like operator annihilates constants and, with p = 0 and Gp = 0 on a uniform rho u, the
    residual is -(net mass flux of the sub-control volume)/tau; scalar: a constant q has no diffusive flux"""
    m = _perturbed_box((1, 1, 1))
    n = m.n_nodes
    u0 = np.tile([0.3, -0.2, 0.7], (n, 1))
    lhs, rhs = H.continuity_adv(m.elem_nodes, m.coords, u0, np.ones(n), np.zeros(n), np.zeros((n, 3)), dt=0.05, gamma1=1.5)
    assert np.abs(lhs[0].sum(axis=1)).max() < 1e-12 * np.abs(lhs).max() * 27
    mdot = H.calc_mass_flow_rate_scs(m.elem_nodes, m.coords, np.ones(n), u0)
    me = H.scs()
    net = np.zeros(27)
    np.add.at(net, me.lrscv[:, 0], mdot[0])
    np.add.at(net, me.lrscv[:, 1], -mdot[0])
    assert np.abs(rhs[0] + net / (0.05 / 1.5)).max() < 1e-12
    # a linear pressure with its exact gradient as Gp: the two pressure terms cancel
    g = np.array([0.4, -0.3, 0.2])
    lhs2, rhs2 = H.continuity_adv(m.elem_nodes, m.coords, u0, np.ones(n), m.coords @ g, np.tile(g, (n, 1)), dt=0.05, gamma1=1.5)
    assert np.abs(rhs2 - rhs).max() < 1e-11
    l3, r3 = H.scalar_adv_diff(m.elem_nodes, m.coords, np.full(n, 2.5), np.full(n, 0.1), mdot)
    assert np.abs(r3[0] + 2.5 * net).max() < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["continuity_adv", "scalar_adv_diff"])
def test_hex27_one_dof_kernels_vs_oracle(kind):
    import torch
    from nalu_b200.linsys import DeviceCsrLinearSystem
    m = _perturbed_box((3, 2, 3))
    vel, visc, rho = _fields(m)
    x = m.coords
    p = 0.5 * np.cos(0.8 * x[:, 0]) * np.sin(0.6 * x[:, 1]) + 0.1 * x[:, 2]
    gp = np.stack([-0.4 * np.sin(0.8 * x[:, 0]), 0.3 * np.cos(0.6 * x[:, 1]), 0.1 + 0.05 * x[:, 2]], axis=1)
    q = 1.0 + 0.5 * np.sin(x[:, 0] + 0.5 * x[:, 1]) * np.cos(0.7 * x[:, 2])
    dev = torch.device("cuda:0")
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
    ls = DeviceCsrLinearSystem(m, 1).allocate(dev)
    ls.zero_system()
    mdot = H.calc_mass_flow_rate_scs(m.elem_nodes, m.coords, rho, vel)
    if kind == "continuity_adv":
        d = ls.make_desc("continuity_adv", fields=dict(coordinates=t(x), velocity=t(vel), density=t(rho), pressure=t(p), dpdx=t(gp)),
                         dt=0.05, gamma1=1.5)
        lhs, rhs = H.continuity_adv(m.elem_nodes, x, vel, rho, p, gp, dt=0.05, gamma1=1.5)
    else:
        d = ls.make_desc("scalar_adv_diff", fields=dict(coordinates=t(x), scalarq=t(q), viscosity=t(visc)), mass_flow_rate=t(mdot))
        lhs, rhs = H.scalar_adv_diff(m.elem_nodes, x, q, visc, mdot)
    assert ls.assemble_elem([d]) == 5
    torch.cuda.synchronize()
    assert ls.geometry_errors() == 0
    rowptr, colidx, values, b = H.assemble_csr(m.node_gid, m.elem_nodes, lhs, rhs, num_dof=1)
    assert np.array_equal(ls.rowptr().astype(np.int64), rowptr) and np.array_equal(ls.colidx().astype(np.int64), colidx)
    got_v, got_b = ls.owned_values().cpu().numpy(), ls.owned_rhs().cpu().numpy()
    assert _rel_err_rows(got_v, values, rowptr) < 1e-12
    rowsum = np.add.reduceat(np.abs(values), rowptr[:-1])
    assert np.abs(got_b - b).max() < 1e-12 * max(np.abs(b).max(), rowsum.max() * np.abs(q).max())
    # deterministic
    v0, r0 = ls.owned_values().clone(), ls.owned_rhs().clone()
    ls.zero_system()
    ls.assemble_elem([d])
    assert torch.equal(v0, ls.owned_values()) and torch.equal(r0, ls.owned_rhs())


@pytest.mark.gpu
def test_hex27_compute_mdot_elem_vs_oracle():
    """ComputeMdotElemAlgorithm on Hex27: the pressure-stabilised mass flow rate, whole mesh and an element range; it is the flux
    whose divergence the continuity kernel puts in its residual"""
    import torch
    from nalu_b200.linsys import DeviceCsrLinearSystem
    m = _perturbed_box((3, 2, 2))
    vel, visc, rho = _fields(m)
    x = m.coords
    p = 0.5 * np.cos(0.8 * x[:, 0]) * np.sin(0.6 * x[:, 1]) + 0.1 * x[:, 2]
    gp = np.stack([-0.4 * np.sin(0.8 * x[:, 0]), 0.3 * np.cos(0.6 * x[:, 1]), 0.1 + 0.05 * x[:, 2]], axis=1)
    tau = 0.05 / 1.5
    dev = torch.device("cuda:0")
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
    ls = DeviceCsrLinearSystem(m, 1)
    ref = H.compute_mdot_elem(m.elem_nodes, x, rho, vel, p, gp, tau)
    got = ls.compute_mdot_elem(t(x), t(rho), t(vel), t(p), t(gp), tau)
    assert got.shape == (m.n_elems, 216)
    assert np.abs(got.cpu().numpy() - ref).max() < 1e-12 * np.abs(ref).max()
    part = torch.full((m.n_elems, 216), np.nan, dtype=torch.float64, device=dev)
    ls.compute_mdot_elem(t(x), t(rho), t(vel), t(p), t(gp), tau, out=part, e_range=(4, 9))
    pn = part.cpu().numpy()
    assert np.isnan(pn[:4]).all() and np.isnan(pn[9:]).all() and np.abs(pn[4:9] - ref[4:9]).max() < 1e-12 * np.abs(ref).max()
    # consistency with the continuity kernel: rhs = -(net flux of every sub-control volume) / tau
    lhs, rhs = H.continuity_adv(m.elem_nodes, x, vel, rho, p, gp, dt=0.05, gamma1=1.5)
    me = H.scs()
    net = np.zeros((m.n_elems, 27))
    for e in range(m.n_elems):
        np.add.at(net[e], me.lrscv[:, 0], ref[e])
        np.add.at(net[e], me.lrscv[:, 1], -ref[e])
    assert np.abs(rhs + net / tau).max() < 1e-11 * np.abs(rhs).max()
    with pytest.raises(Exception):
        ls.compute_mdot_elem(t(x), t(rho), t(vel), t(p), t(gp), tau, options=("shift_mdot",))


def test_nodal_gradient_oracle_is_exact_for_linear_fields_inside():
    """Green-Gauss over closed interior control volumes reproduces the gradient of a linear field exactly (the known-answer test of
    the Hex8 path) on affinely mapped Hex27 elements (on curved ones the 2 x 2 Gauss rule no longer integrates p A exactly: the
    reference has the same quadrature error); boundary nodes lack the boundary contribution"""
    m = M.hex27_box(2, 2, 2)
    m.coords = m.coords @ np.array([[1.0, 0.2, -0.1], [0.15, 0.9, 0.25], [-0.05, 0.1, 1.1]]).T
    dv = H.dual_nodal_volume(m.elem_nodes, m.coords, m.n_nodes)
    assert abs(dv.sum() - H.Hex27SCV().volumes(m.coords[m.elem_nodes]).sum()) < 1e-12
    g = np.array([0.7, -0.4, 0.25])
    grad = H.assemble_nodal_grad_elem(m.elem_nodes, m.coords, m.coords @ g + 0.3, dv)
    lat = np.rint(M.hex27_box(2, 2, 2).coords * 2).astype(int)
    interior = np.all((lat > 0) & (lat < 4), axis=1)
    assert interior.sum() == 27 and np.abs(grad[interior] - g).max() < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("ncomp", [1, 3])
def test_hex27_nodal_gradient_vs_oracle(ncomp):
    import torch
    from nalu_b200.linsys import DeviceCsrLinearSystem
    m = _perturbed_box((3, 2, 2))
    vel, visc, rho = _fields(m)
    field = visc if ncomp == 1 else vel
    dv = H.dual_nodal_volume(m.elem_nodes, m.coords, m.n_nodes)
    ref = H.assemble_nodal_grad_elem(m.elem_nodes, m.coords, field, dv)
    dev = torch.device("cuda:0")
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
    ls = DeviceCsrLinearSystem(m, 1)
    grad = torch.zeros((m.n_nodes, 3) if ncomp == 1 else (m.n_nodes, 3, 3), dtype=torch.float64, device=dev)
    ls.assemble_nodal_grad_elem(t(m.coords), t(field), t(dv), grad)
    torch.cuda.synchronize()
    assert np.abs(grad.cpu().numpy() - ref).max() < 1e-12 * max(np.abs(ref).max(), 1.0) * 27
    with pytest.raises(Exception):
        ls.assemble_nodal_grad_elem(t(m.coords), t(field), t(dv), grad, use_shifted=True)


@pytest.mark.gpu
def test_hex27_periodic_aliasing_takes_the_atomic_path():
    """a box one element wide and periodic in x (the subject plane adopts the monarch's ids, src/PeriodicManager.C:698-724): every
    element reaches rows through two of its nodes, no colouring exists, the kernel sums with atomics -- every contribution is
    kept (the reference's sum_into_row drops part of a row with repeated column ids, as for Hex8: DESIGN 1.2)"""
    import torch
    from nalu_b200.linsys import DeviceCsrLinearSystem
    m = _perturbed_box((1, 2, 2))
    sx = 3
    ix = np.arange(m.n_nodes) % sx
    m.node_gid = np.where(ix == 2, m.node_gid - 2, m.node_gid)   # x = 1 plane -> ids of the x = 0 plane
    vel, visc, rho = _fields(m)
    # periodic fields: the aliased nodes carry the monarch's values (coordinates stay their own)
    mon = np.where(ix == 2, np.arange(m.n_nodes) - 2, np.arange(m.n_nodes))
    vel, visc, rho = vel[mon], visc[mon], rho[mon]
    ls, mode, _ = _device_assemble(m, vel, visc, rho, 1.0)
    assert mode == 5 and ls.elem_colors()[0] == 0
    mdot = H.calc_mass_flow_rate_scs(m.elem_nodes, m.coords, rho, vel)
    lhs, rhs = H.momentum_adv_diff(m.elem_nodes, m.coords, vel, visc, mdot, include_divu=1.0)
    rowptr, colidx, values, b = H.assemble_csr(m.node_gid, m.elem_nodes, lhs, rhs)
    assert np.array_equal(ls.rowptr().astype(np.int64), rowptr) and np.array_equal(ls.colidx().astype(np.int64), colidx)
    got_v, got_b = ls.owned_values().cpu().numpy(), ls.owned_rhs().cpu().numpy()
    assert np.abs(got_v - values).max() < 1e-12 * np.abs(values).max()
    assert np.abs(got_b - b).max() < 1e-12 * np.abs(values).max() * np.abs(vel).max() * 27


@pytest.mark.gpu
def test_hex27_assembly_is_deterministic_and_accumulates():
    m = _perturbed_box((3, 3, 2))
    vel, visc, rho = _fields(m)
    ls1, _, _ = _device_assemble(m, vel, visc, rho, 1.0)
    ls2, _, _ = _device_assemble(m, vel, visc, rho, 1.0)
    import torch
    assert torch.equal(ls1.owned_values(), ls2.owned_values()) and torch.equal(ls1.owned_rhs(), ls2.owned_rhs())


@pytest.mark.gpu
@pytest.mark.parametrize("nranks", [2, 3])
def test_hex27_partitioned_assembly_after_load_complete(nranks):
    """random element partition over simulated ranks on one GPU: assemble per rank, export-add the shared-not-owned rows
    (loadComplete, src/TpetraLinearSystem.C:1757-1778), compare every owned entry with the one-rank oracle by global ids"""
    import os, sys
    import torch
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_graph import build_product_systems
    from nalu_b200.linsys import DeviceCsrLinearSystem
    m = _perturbed_box((3, 3, 2))
    vel, visc, rho = _fields(m)
    rng = np.random.default_rng(13)
    elem_rank = rng.integers(0, nranks, size=m.n_elems)
    parts = M.partition_elements(m, elem_rank, nranks)
    host = build_product_systems(parts, 3)
    sends = {p.mesh.rank: p.handshake_send for p in host}
    dev = torch.device("cuda:0")
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
    systems = []
    for pm in parts:
        ex = lambda send, r=pm.rank: {src: sends[src][r] for src in sends if r in sends[src]}
        ls = DeviceCsrLinearSystem(pm, 3, exchange=ex).allocate(dev)
        # nodal fields of the rank's nodes: the partition keeps global ids, fields are functions of the global node
        g = pm.node_gid - 1
        coords, v, mu, r = t(m.coords[g]), t(vel[g]), t(visc[g]), t(rho[g])
        mdot = ls.mdot_scs(coords, r, v)
        ls.zero_system()
        assert ls.assemble_elem([ls.make_desc("momentum_adv_diff", fields=dict(coordinates=coords, velocity=v, viscosity=mu),
                                              mass_flow_rate=mdot, include_divu=1.0)]) == 5
        systems.append(ls)
    torch.cuda.synchronize()
    for ls in systems:
        send, recv = ls.halo_peers()
        for i, (peer, n) in enumerate(recv):
            src = systems[peer]
            ssend, _ = src.halo_peers()
            j = [q for q, (pr, _) in enumerate(ssend) if pr == ls.mesh.rank][0]
            buf = torch.empty(ssend[j][1], dtype=torch.float64, device=dev)
            src.halo_pack(j, buf)
            ls.halo_unpack_add(i, buf)
    torch.cuda.synchronize()
    mdot = H.calc_mass_flow_rate_scs(m.elem_nodes, m.coords, rho, vel)
    lhs, rhs = H.momentum_adv_diff(m.elem_nodes, m.coords, vel, visc, mdot, include_divu=1.0)
    rowptr, colidx, values, b = H.assemble_csr(m.node_gid, m.elem_nodes, lhs, rhs)
    n = rowptr.size - 1
    keys = (np.repeat(np.arange(n), np.diff(rowptr)) + 1) * (n + 1) + (colidx + 1)
    scale = np.abs(values).max()
    for ls in systems:
        k = _rank_keys(ls, n)
        pos = np.searchsorted(keys, k)
        assert np.array_equal(keys[pos], k)
        got = ls.owned_values().cpu().numpy()
        assert np.abs(got - values[pos]).max() < 1e-12 * scale
        rg = ls.row_gids()
        assert np.abs(ls.owned_rhs().cpu().numpy() - b[rg - 1]).max() < 1e-12 * scale * np.abs(vel).max() * 27


# ---- empty inputs: a rank / part without elements (the reference's bucket loops simply do not run) --------------------------

@pytest.mark.parametrize("npe", [8, 27])
@pytest.mark.parametrize("n_nodes", [0, 5])
def test_empty_part_graph(npe, n_nodes):
    from nalu_b200.linsys import DeviceCsrLinearSystem
    m = M.HexMesh(coords=np.zeros((n_nodes, 3)), node_gid=np.arange(1, n_nodes + 1, dtype=np.int64),
                  node_owner=np.zeros(n_nodes, dtype=np.int32), elem_nodes=np.zeros((0, npe), dtype=np.int32))
    ls = DeviceCsrLinearSystem(m, 3, host_only=True)
    assert ls.n_owned_rows == 3 * n_nodes and ls.nnz_owned == 0
    assert np.array_equal(ls.rowptr(), np.zeros(3 * n_nodes + 1, dtype=np.uint64))


@pytest.mark.gpu
@pytest.mark.parametrize("npe", [8, 27])
def test_empty_part_assembly_is_a_no_op(npe):
    import torch
    from nalu_b200.linsys import DeviceCsrLinearSystem
    dev = torch.device("cuda:0")
    n = 5
    m = M.HexMesh(coords=np.zeros((n, 3)), node_gid=np.arange(1, n + 1, dtype=np.int64), node_owner=np.zeros(n, dtype=np.int32),
                  elem_nodes=np.zeros((0, npe), dtype=np.int32))
    ls = DeviceCsrLinearSystem(m, 3).allocate(dev)
    ls.zero_system()
    z3, z1 = torch.zeros((n, 3), dtype=torch.float64, device=dev), torch.zeros(n, dtype=torch.float64, device=dev)
    mdot = torch.zeros((0, 12 if npe == 8 else 216), dtype=torch.float64, device=dev)
    # a zero-size tensor has a null data pointer in torch: give the descriptor a real (unused) address
    mdot_ptr = torch.zeros(1, dtype=torch.float64, device=dev)
    d = ls.make_desc("momentum_adv_diff", fields=dict(coordinates=z3, velocity=z3, viscosity=z1), mass_flow_rate=mdot_ptr, include_divu=1.0)
    ls.assemble_elem([d])
    torch.cuda.synchronize()
    assert ls.geometry_errors() == 0
    assert float(ls.owned_rhs().abs().sum()) == 0.0 and ls.owned_values().numel() == 0
    del mdot
