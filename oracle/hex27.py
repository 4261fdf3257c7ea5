"""CPU restatement (numpy) of the reference's Hex27 (P2 CVFEM) momentum assembly.  TEST INFRASTRUCTURE ONLY: imported by
tests/, `__graft_entry__.smoke()` and the checker / cpu_baseline leg of tools/bench_hex27.py, never by the product path.

Follows, function by function (paths relative to /root/reference):
  * HexahedralP2Element tables        src/master_element/Hex27CVFEM.C:25-75 (stkNodeMap_, scsEndLoc_, quadrature rule),
                                      :192-262 (tensor_product_node_map, gauss_point_location, tensor_product_weight),
                                      :355-421 hex27_shape_fcn (restated literally, all 27 formulas),
                                      :423-632 hex27_shape_deriv (restated as the derivative of the same tensor-product
                                      polynomials; checked against the literal shape functions by differences and by the
                                      polynomial-exactness tests of the reference)
  * Hex27SCS::set_interior_info       :887-1050 (216 points: location, lrscv, signed weight, direction; loop order kept)
  * Hex27SCS area vectors             :1428-1470 area_vector<direction> + weighted_area_vectors
                                      (include/master_element/Hex27CVFEM.h:378-415)
  * Hex27SCS::grad_op                 generic_grad_op, include/master_element/MasterElementFunctions.h:52-108
  * Hex27SCV::set_interior_info / determinant   :660-707, :743-763, :809-845
  * MomentumAdvDiffElemKernel<AlgTraitsHex27>::execute   src/kernel/MomentumAdvDiffElemKernel.C:95-186
  * calc_mass_flow_rate_scs           unit_tests/kernels/UnitTestKernelUtils.C:547-628
  * sumInto into the CSR              src/TpetraLinearSystem.C:1457-1522 (rows / columns sorted by global id)

Parity pinning: the reference holds NO gold table for Hex27 (its Hex27 tests are property tests:
unit_tests/UnitTestHexMasterElements.C:297-321 interpolation / derivative exactness for quadratic polynomials,
unit_tests/UnitTestMasterElements.C:185-229 volume integration); tests/test_hex27.py restates those tests on these tables, adds
closed forms (unit-cube areas, closed control-volume surfaces, uniform-flow and linear-field identities) and checks that on an
element whose mid nodes sit at the edge / face / body centres the Hex27 geometry reproduces the trilinear (Hex8) one.
"""
import numpy as np

NPE, NIP = 27, 216
STK_NODE_MAP = np.array([0, 8, 1, 11, 21, 9, 3, 10, 2, 12, 25, 13, 23, 20, 24, 15, 26, 14, 4, 16, 5, 19, 22, 17, 7, 18, 6])
SCS_DIST = np.sqrt(3.0) / 3.0
SCS_END_LOC = np.array([-1.0, -SCS_DIST, SCS_DIST, 1.0])
GAUSS_ABSCISSAE = np.array([-np.sqrt(3.0) / 3.0, np.sqrt(3.0) / 3.0])
GAUSS_WEIGHT = np.array([0.5, 0.5])
S_DIRECTION, T_DIRECTION, U_DIRECTION = 0, 1, 2


def tensor_product_node_map(i, j, k):
    return int(STK_NODE_MAP[i + 3 * (j + 3 * k)])


def isoparametric_mapping(b, a, xi):
    return xi * (b - a) / 2.0 + (a + b) / 2.0


def gauss_point_location(node_ordinal, gauss_point_ordinal):
    return isoparametric_mapping(SCS_END_LOC[node_ordinal + 1], SCS_END_LOC[node_ordinal], GAUSS_ABSCISSAE[gauss_point_ordinal])


def tensor_product_weight2(s1_node, s2_node, s1_ip, s2_ip):
    ls1 = SCS_END_LOC[s1_node + 1] - SCS_END_LOC[s1_node]
    ls2 = SCS_END_LOC[s2_node + 1] - SCS_END_LOC[s2_node]
    return ls1 * ls2 * GAUSS_WEIGHT[s1_ip] * GAUSS_WEIGHT[s2_ip]


def tensor_product_weight3(s1_node, s2_node, s3_node, s1_ip, s2_ip, s3_ip):
    ls1 = SCS_END_LOC[s1_node + 1] - SCS_END_LOC[s1_node]
    ls2 = SCS_END_LOC[s2_node + 1] - SCS_END_LOC[s2_node]
    ls3 = SCS_END_LOC[s3_node + 1] - SCS_END_LOC[s3_node]
    return ls1 * ls2 * ls3 * GAUSS_WEIGHT[s1_ip] * GAUSS_WEIGHT[s2_ip] * GAUSS_WEIGHT[s3_ip]


def hex27_shape_fcn(intg_loc):
    """Hex27CVFEM.C:355-421, literally; intg_loc [n,3] -> [n,27]"""
    p = np.asarray(intg_loc, dtype=np.float64)
    s, t, u = p[:, 0], p[:, 1], p[:, 2]
    one, half, one4th, one8th = 1.0, 1.0 / 2.0, 1.0 / 4.0, 1.0 / 8.0
    stu, st, su, tu = s * t * u, s * t, s * u, t * u
    one_m_s, one_p_s = one - s, one + s
    one_m_t, one_p_t = one - t, one + t
    one_m_u, one_p_u = one - u, one + u
    one_m_ss, one_m_tt, one_m_uu = one - s * s, one - t * t, one - u * u
    f = np.empty((p.shape[0], 27))
    f[:, 0] = -one8th * stu * one_m_s * one_m_t * one_m_u
    f[:, 1] = one8th * stu * one_p_s * one_m_t * one_m_u
    f[:, 2] = -one8th * stu * one_p_s * one_p_t * one_m_u
    f[:, 3] = one8th * stu * one_m_s * one_p_t * one_m_u
    f[:, 4] = one8th * stu * one_m_s * one_m_t * one_p_u
    f[:, 5] = -one8th * stu * one_p_s * one_m_t * one_p_u
    f[:, 6] = one8th * stu * one_p_s * one_p_t * one_p_u
    f[:, 7] = -one8th * stu * one_m_s * one_p_t * one_p_u
    f[:, 8] = one4th * tu * one_m_ss * one_m_t * one_m_u
    f[:, 9] = -one4th * su * one_p_s * one_m_tt * one_m_u
    f[:, 10] = -one4th * tu * one_m_ss * one_p_t * one_m_u
    f[:, 11] = one4th * su * one_m_s * one_m_tt * one_m_u
    f[:, 12] = one4th * st * one_m_s * one_m_t * one_m_uu
    f[:, 13] = -one4th * st * one_p_s * one_m_t * one_m_uu
    f[:, 14] = one4th * st * one_p_s * one_p_t * one_m_uu
    f[:, 15] = -one4th * st * one_m_s * one_p_t * one_m_uu
    f[:, 16] = -one4th * tu * one_m_ss * one_m_t * one_p_u
    f[:, 17] = one4th * su * one_p_s * one_m_tt * one_p_u
    f[:, 18] = one4th * tu * one_m_ss * one_p_t * one_p_u
    f[:, 19] = -one4th * su * one_m_s * one_m_tt * one_p_u
    f[:, 20] = one_m_ss * one_m_tt * one_m_uu
    f[:, 21] = -half * u * one_m_ss * one_m_tt * one_m_u
    f[:, 22] = half * u * one_m_ss * one_m_tt * one_p_u
    f[:, 23] = -half * s * one_m_s * one_m_tt * one_m_uu
    f[:, 24] = half * s * one_p_s * one_m_tt * one_m_uu
    f[:, 25] = -half * t * one_m_ss * one_m_t * one_m_uu
    f[:, 26] = half * t * one_m_ss * one_p_t * one_m_uu
    return f


def _lagrange(x):
    x = np.asarray(x, dtype=np.float64)
    val = np.stack([-0.5 * x * (1.0 - x), (1.0 - x) * (1.0 + x), 0.5 * x * (1.0 + x)], axis=-1)
    der = np.stack([x - 0.5, -2.0 * x, x + 0.5], axis=-1)
    return val, der


def hex27_shape_deriv(intg_loc):
    """Hex27CVFEM.C:423-632: derivatives of the 27 shape functions, [n,27,3].  The reference spells out 81 formulas; they
    are the derivatives of N_(a,b,c) = l_a(s) l_b(t) l_c(u) (the factorisation visible in hex27_shape_fcn), node (a,b,c) at
    STK ordinal stkNodeMap_[a + 3 (b + 3 c)]."""
    p = np.asarray(intg_loc, dtype=np.float64)
    ls, ds = _lagrange(p[:, 0])
    lt, dt = _lagrange(p[:, 1])
    lu, du = _lagrange(p[:, 2])
    out = np.empty((p.shape[0], 27, 3))
    for c in range(3):
        for b in range(3):
            for a in range(3):
                n = tensor_product_node_map(a, b, c)
                out[:, n, 0] = ds[:, a] * lt[:, b] * lu[:, c]
                out[:, n, 1] = ls[:, a] * dt[:, b] * lu[:, c]
                out[:, n, 2] = ls[:, a] * lt[:, b] * du[:, c]
    return out


class Hex27SCS:
    """Hex27SCS::set_interior_info (Hex27CVFEM.C:887-1050) + tables at the integration points"""

    def __init__(self):
        surfaces_per_direction, nodes1d, num_quad = 2, 3, 2
        self.lrscv = np.zeros((NIP, 2), dtype=np.int64)
        self.intg_loc = np.zeros((NIP, 3))
        self.weight = np.zeros(NIP)
        self.direction = np.zeros(NIP, dtype=np.int64)
        scs_loc = [-SCS_DIST, SCS_DIST]
        orientation = [-1.0, +1.0]
        idx = 0
        # u direction: bottom-top
        for m in range(surfaces_per_direction):
            for l in range(nodes1d):
                for k in range(nodes1d):
                    if m == 0:
                        left, right = tensor_product_node_map(k, l, m), tensor_product_node_map(k, l, m + 1)
                    else:
                        left, right = tensor_product_node_map(k, l, m + 1), tensor_product_node_map(k, l, m)
                    for j in range(num_quad):
                        for i in range(num_quad):
                            self.lrscv[idx] = (left, right)
                            self.intg_loc[idx] = (gauss_point_location(k, i), gauss_point_location(l, j), scs_loc[m])
                            self.weight[idx] = orientation[m] * tensor_product_weight2(k, l, i, j)
                            self.direction[idx] = U_DIRECTION
                            idx += 1
        # t direction: front-back
        for m in range(surfaces_per_direction):
            for l in range(nodes1d):
                for k in range(nodes1d):
                    if m == 0:
                        left, right = tensor_product_node_map(k, m, l), tensor_product_node_map(k, m + 1, l)
                    else:
                        left, right = tensor_product_node_map(k, m + 1, l), tensor_product_node_map(k, m, l)
                    for j in range(num_quad):
                        for i in range(num_quad):
                            self.lrscv[idx] = (left, right)
                            self.intg_loc[idx] = (gauss_point_location(k, i), scs_loc[m], gauss_point_location(l, j))
                            self.weight[idx] = orientation[m] * tensor_product_weight2(k, l, i, j)
                            self.direction[idx] = T_DIRECTION
                            idx += 1
        # s direction: left-right
        for m in range(surfaces_per_direction):
            for l in range(nodes1d):
                for k in range(nodes1d):
                    if m == 0:
                        left, right = tensor_product_node_map(m, k, l), tensor_product_node_map(m + 1, k, l)
                    else:
                        left, right = tensor_product_node_map(m + 1, k, l), tensor_product_node_map(m, k, l)
                    for j in range(num_quad):
                        for i in range(num_quad):
                            self.lrscv[idx] = (left, right)
                            self.intg_loc[idx] = (scs_loc[m], gauss_point_location(k, i), gauss_point_location(l, j))
                            self.weight[idx] = -orientation[m] * tensor_product_weight2(k, l, i, j)
                            self.direction[idx] = S_DIRECTION
                            idx += 1
        assert idx == NIP
        self.shape_fcn = hex27_shape_fcn(self.intg_loc)       # [216,27]
        self.shape_deriv = hex27_shape_deriv(self.intg_loc)   # [216,27,3]

    def area_vectors(self, coords):
        """weighted_area_vectors (Hex27CVFEM.h:378-415,435-470): coords [E,27,3] -> areav [E,216,3]"""
        coords = np.asarray(coords, dtype=np.float64)
        out = np.empty((coords.shape[0], NIP, 3))
        for direction in (U_DIRECTION, T_DIRECTION, S_DIRECTION):
            s1 = S_DIRECTION if direction == T_DIRECTION else T_DIRECTION
            s2 = S_DIRECTION if direction == U_DIRECTION else U_DIRECTION
            ips = np.nonzero(self.direction == direction)[0]
            # sjac[i][0|1] = sum_n dn_ds1|2 * coords(n, i)
            d1 = np.einsum("pn,eni->epi", self.shape_deriv[ips, :, s1], coords)
            d2 = np.einsum("pn,eni->epi", self.shape_deriv[ips, :, s2], coords)
            out[:, ips, 0] = d1[..., 1] * d2[..., 2] - d1[..., 2] * d2[..., 1]
            out[:, ips, 1] = d1[..., 2] * d2[..., 0] - d1[..., 0] * d2[..., 2]
            out[:, ips, 2] = d1[..., 0] * d2[..., 1] - d1[..., 1] * d2[..., 0]
        return out * self.weight[None, :, None]

    def grad_op(self, coords):
        return generic_grad_op(self.shape_deriv, coords)


def generic_grad_op(ref_grad, coords):
    """MasterElementFunctions.h:52-108: ref_grad [P,27,3], coords [E,27,3] -> dndx [E,P,27,3], det [E,P]"""
    coords = np.asarray(coords, dtype=np.float64)
    jact = np.einsum("pnj,eni->epij", ref_grad, coords)  # jact[i][j] += refGrad[n][j] * coords(n, i)
    adj = np.empty_like(jact)                            # cofactorMatrix
    adj[..., 0, 0] = jact[..., 1, 1] * jact[..., 2, 2] - jact[..., 2, 1] * jact[..., 1, 2]
    adj[..., 0, 1] = jact[..., 1, 2] * jact[..., 2, 0] - jact[..., 2, 2] * jact[..., 1, 0]
    adj[..., 0, 2] = jact[..., 1, 0] * jact[..., 2, 1] - jact[..., 2, 0] * jact[..., 1, 1]
    adj[..., 1, 0] = jact[..., 0, 2] * jact[..., 2, 1] - jact[..., 2, 2] * jact[..., 0, 1]
    adj[..., 1, 1] = jact[..., 0, 0] * jact[..., 2, 2] - jact[..., 2, 0] * jact[..., 0, 2]
    adj[..., 1, 2] = jact[..., 0, 1] * jact[..., 2, 0] - jact[..., 2, 1] * jact[..., 0, 0]
    adj[..., 2, 0] = jact[..., 0, 1] * jact[..., 1, 2] - jact[..., 1, 1] * jact[..., 0, 2]
    adj[..., 2, 1] = jact[..., 0, 2] * jact[..., 1, 0] - jact[..., 1, 2] * jact[..., 0, 0]
    adj[..., 2, 2] = jact[..., 0, 0] * jact[..., 1, 1] - jact[..., 1, 0] * jact[..., 0, 1]
    det = np.einsum("epi,epi->ep", jact[..., :, 0], adj[..., :, 0])
    dndx = np.einsum("epij,pnj->epni", adj, ref_grad) / det[..., None, None]
    return dndx, det


class Hex27SCV:
    """Hex27SCV::set_interior_info (Hex27CVFEM.C:660-707), determinant (:743-763)"""

    def __init__(self):
        self.ip_node_map = np.zeros(NIP, dtype=np.int64)
        self.intg_loc = np.zeros((NIP, 3))
        self.ip_weight = np.zeros(NIP)
        idx = 0
        for n in range(3):
            for m in range(3):
                for l in range(3):
                    node = tensor_product_node_map(l, m, n)
                    for k in range(2):
                        for j in range(2):
                            for i in range(2):
                                self.intg_loc[idx] = (gauss_point_location(l, i), gauss_point_location(m, j), gauss_point_location(n, k))
                                self.ip_weight[idx] = tensor_product_weight3(l, m, n, i, j, k)
                                self.ip_node_map[idx] = node
                                idx += 1
        self.shape_fcn = hex27_shape_fcn(self.intg_loc)
        self.shape_deriv = hex27_shape_deriv(self.intg_loc)
        # shifted points (Hex27CVFEM.C:214-222,688-690): gaussAbscissaeShift_ = {-1,-1,0,0,+1,+1}[2 node + ip], i.e. the node itself
        shift = np.array([-1.0, -1.0, 0.0, 0.0, 1.0, 1.0])
        self.intg_loc_shift = np.zeros((NIP, 3))
        idx = 0
        for n in range(3):
            for m in range(3):
                for l in range(3):
                    for k in range(2):
                        for j in range(2):
                            for i in range(2):
                                self.intg_loc_shift[idx] = (shift[2 * l + i], shift[2 * m + j], shift[2 * n + k])
                                idx += 1
        self.shape_fcn_shift = hex27_shape_fcn(self.intg_loc_shift)

    def volumes(self, coords):
        _, det = generic_grad_op(self.shape_deriv, coords)
        return det * self.ip_weight[None, :]


_SCS = None


def scs():
    global _SCS
    if _SCS is None:
        _SCS = Hex27SCS()
    return _SCS


def calc_mass_flow_rate_scs(elem_nodes, coords, density, velocity):
    """unit_tests/kernels/UnitTestKernelUtils.C:547-628 with AlgTraitsHex27: mdot[E,216]"""
    me = scs()
    c = np.asarray(coords)[elem_nodes]
    rho = np.asarray(density)[elem_nodes]
    u = np.asarray(velocity)[elem_nodes]
    areav = me.area_vectors(c)
    rho_u = np.einsum("pn,en,enj->epj", me.shape_fcn, rho, u)
    return np.einsum("epj,epj->ep", rho_u, areav)


def momentum_adv_diff(elem_nodes, coords, velocity, viscosity, mdot, include_divu=0.0):
    """MomentumAdvDiffElemKernel<AlgTraitsHex27>::execute (src/kernel/MomentumAdvDiffElemKernel.C:95-186), default options
    (no shifted gradient operator, no skew-symmetric advection): lhs [E,81,81], rhs [E,81]; the loop over integration
    points is kept, the loops over (ic, i, j) are array operations over all elements."""
    me = scs()
    E = elem_nodes.shape[0]
    c = np.asarray(coords)[elem_nodes]
    u = np.asarray(velocity)[elem_nodes]            # v_uNp1 [E,27,3]
    visc = np.asarray(viscosity)[elem_nodes]        # [E,27]
    areav = me.area_vectors(c)                      # v_scs_areav
    dndx, _ = me.grad_op(c)                         # v_dndx [E,216,27,3]
    sf = me.shape_fcn                               # v_shape_function_ = v_adv_shape_function_
    lhs = np.zeros((E, 27, 3, 27, 3))
    rhs = np.zeros((E, 27, 3))
    ar = np.arange(E)
    for ip in range(NIP):
        il, ir = me.lrscv[ip]
        tmdot = mdot[:, ip]
        mu_ip = visc @ sf[ip]
        u_ip = np.einsum("n,enj->ej", sf[ip], u)
        div_u = np.einsum("enj,enj->e", u, dndx[:, ip])
        a = areav[:, ip]                            # [E,3]
        # rhs: advective flux + divU stress
        flux = tmdot[:, None] * u_ip + (2.0 / 3.0) * (mu_ip * div_u)[:, None] * a * include_divu
        rhs[:, il] -= flux
        rhs[:, ir] += flux
        # lhs
        lhsfac_adv = sf[ip][None, :] * tmdot[:, None]                                  # [E,ic]
        lhs_ric = -mu_ip[:, None] * np.einsum("ecj,ej->ec", dndx[:, ip], a)           # sum_j -mu dndx(ic,j) A_j
        diag = lhsfac_adv + lhs_ric                                                    # -> lhs(row i, ic*3+i)
        # lhsfacDiff_j at (row i, col ic*3+j) = -mu dndx(ic,i) A_j
        blk = -mu_ip[:, None, None, None] * dndx[:, ip][:, :, :, None] * a[:, None, None, :]   # [E,ic,i,j]
        blk = blk.transpose(0, 2, 1, 3).copy()                                         # [E,i,ic,j]
        for i in range(3):
            blk[:, i, :, i] += diag
        lhs[:, il] += blk
        lhs[:, ir] -= blk
        # rhs -= lhsfacDiff_j * uj + lhs_riC_i * ui  (the advective lhs is NOT applied: aflux above is its action)
        blk_nodiag = blk.copy()
        for i in range(3):
            blk_nodiag[:, i, :, i] -= lhsfac_adv
        r = np.einsum("eicj,ecj->ei", blk_nodiag, u)
        rhs[:, il] -= r
        rhs[:, ir] += r
    del ar
    return lhs.reshape(E, 81, 81), rhs.reshape(E, 81)


def continuity_adv(elem_nodes, coords, velocity, density, pressure, gpdx, dt, gamma1):
    """ContinuityAdvElemKernel<AlgTraitsHex27>::execute (src/kernel/ContinuityAdvElemKernel.C:95-175), default options (no
    shifted mdot / Poisson operator, full sensitivities): lhs [E,27,27], rhs [E,27]"""
    me = scs()
    E = elem_nodes.shape[0]
    proj_time_scale = dt / gamma1
    c = np.asarray(coords)[elem_nodes]
    u = np.asarray(velocity)[elem_nodes]
    rho = np.asarray(density)[elem_nodes]
    p = np.asarray(pressure)[elem_nodes]
    gp = np.asarray(gpdx)[elem_nodes]
    areav = me.area_vectors(c)
    dndx, _ = me.grad_op(c)
    sf = me.shape_fcn
    lhs = np.zeros((E, 27, 27))
    rhs = np.zeros((E, 27))
    for ip in range(NIP):
        il, ir = me.lrscv[ip]
        a = areav[:, ip]
        w_gpdx_ip = np.einsum("n,enj->ej", sf[ip], gp)
        w_rho_u_ip = np.einsum("n,en,enj->ej", sf[ip], rho, u)
        w_dpdx_ip = np.einsum("enj,en->ej", dndx[:, ip], p)
        lhsfac = -np.einsum("enj,ej->en", dndx[:, ip], a)
        lhs[:, il] += lhsfac
        lhs[:, ir] -= lhsfac
        mdot = np.einsum("ej,ej->e", w_rho_u_ip - proj_time_scale * (w_dpdx_ip - w_gpdx_ip), a)
        rhs[:, il] -= mdot / proj_time_scale
        rhs[:, ir] += mdot / proj_time_scale
    return lhs, rhs


def compute_mdot_elem(elem_nodes, coords, density, velocity, pressure, gpdx, proj_time_scale):
    """ComputeMdotElemAlgorithm::execute (src/ComputeMdotElemAlgorithm.C:85-254) with the Hex27 master element, default options:
    mdot[E,216] = (rho u_ip - projTimeScale (dpdx_ip - Gpdx_ip)) . areav_ip"""
    me = scs()
    c = np.asarray(coords)[elem_nodes]
    u = np.asarray(velocity)[elem_nodes]
    rho = np.asarray(density)[elem_nodes]
    p = np.asarray(pressure)[elem_nodes]
    gp = np.asarray(gpdx)[elem_nodes]
    areav = me.area_vectors(c)
    dndx, _ = me.grad_op(c)
    sf = me.shape_fcn
    rho_u_ip = np.einsum("pn,en,enj->epj", sf, rho, u)
    gp_ip = np.einsum("pn,enj->epj", sf, gp)
    dpdx_ip = np.einsum("epnj,en->epj", dndx, p)
    return np.einsum("epj,epj->ep", rho_u_ip - proj_time_scale * (dpdx_ip - gp_ip), areav)


def dual_nodal_volume(elem_nodes, coords, n_nodes):
    """dual_nodal_volume: the SCV volumes of every element summed into their nodes (src/ComputeGeometryInteriorAlgorithm.C:72-150)"""
    me = scv()
    vol = me.volumes(np.asarray(coords)[elem_nodes])                        # [E,216]
    out = np.zeros(n_nodes)
    np.add.at(out, elem_nodes[:, me.ip_node_map], vol)
    return out


def assemble_nodal_grad_elem(elem_nodes, coords, field, dual_volume):
    """AssembleNodalGradElemAlgorithm / AssembleNodalGradUElemAlgorithm::execute (src/AssembleNodalGradElemAlgorithm.C:64-150,
    src/AssembleNodalGradUElemAlgorithm.C:120-195) with the Hex27 master element: grad[n, i, j] += q_i,ip A_j / V_n at the left
    node, minus at the right node; field [nNodes] or [nNodes, ncomp]"""
    me = scs()
    f = np.asarray(field, dtype=np.float64)
    f2 = f[:, None] if f.ndim == 1 else f
    c = np.asarray(coords)[elem_nodes]
    areav = me.area_vectors(c)                                               # [E,216,3]
    q_ip = np.einsum("pn,eni->epi", me.shape_fcn, f2[elem_nodes])            # [E,216,ncomp]
    fac = q_ip[:, :, :, None] * areav[:, :, None, :]                          # [E,216,ncomp,3]
    grad = np.zeros((f2.shape[0], f2.shape[1], 3))
    nl, nr = elem_nodes[:, me.lrscv[:, 0]], elem_nodes[:, me.lrscv[:, 1]]
    dv = np.asarray(dual_volume)
    np.add.at(grad, nl, fac / dv[nl][:, :, None, None])
    np.add.at(grad, nr, -fac / dv[nr][:, :, None, None])
    return grad[:, 0, :] if f.ndim == 1 else grad


def scalar_adv_diff(elem_nodes, coords, scalar_q, diff_flux_coeff, mdot):
    """ScalarAdvDiffElemKernel<AlgTraitsHex27>::execute (src/kernel/ScalarAdvDiffElemKernel.C:69-127), default options:
    lhs [E,27,27], rhs [E,27]"""
    me = scs()
    E = elem_nodes.shape[0]
    c = np.asarray(coords)[elem_nodes]
    q = np.asarray(scalar_q)[elem_nodes]
    gam = np.asarray(diff_flux_coeff)[elem_nodes]
    areav = me.area_vectors(c)
    dndx, _ = me.grad_op(c)
    sf = me.shape_fcn
    lhs = np.zeros((E, 27, 27))
    rhs = np.zeros((E, 27))
    for ip in range(NIP):
        il, ir = me.lrscv[ip]
        tmdot = mdot[:, ip]
        gam_ip = gam @ sf[ip]
        lhsfac_adv = sf[ip][None, :] * tmdot[:, None]
        lhsfac_diff = -gam_ip[:, None] * np.einsum("enj,ej->en", dndx[:, ip], areav[:, ip])
        tot = lhsfac_adv + lhsfac_diff
        lhs[:, il] += tot
        lhs[:, ir] -= tot
        flux = np.einsum("en,en->e", tot, q)   # qAdv + qDiff
        rhs[:, il] -= flux
        rhs[:, ir] += flux
    return lhs, rhs


_SCV = None


def scv():
    global _SCV
    if _SCV is None:
        _SCV = Hex27SCV()
    return _SCV


def momentum_mass(elem_nodes, coords, vel_np1, vel_n, vel_nm1, rho_np1, rho_n, rho_nm1, gpdx, dt, gamma1, gamma2, gamma3,
                  lumped=True, dens_fac=0.0, om_dens_fac=1.0):
    """MomentumMassElemKernel<AlgTraitsHex27>::execute (src/kernel/MomentumMassElemKernel.C:98-171): lhs [E,81,81], rhs [E,81];
    lumped = the shifted SCV shape functions (:60-63)"""
    me = scv()
    E = elem_nodes.shape[0]
    sf = me.shape_fcn_shift if lumped else me.shape_fcn       # v_shape_function_ [216,27]
    c = np.asarray(coords)[elem_nodes]
    scv_vol = me.volumes(c)                                   # v_scv_volume [E,216]
    g = lambda a: np.asarray(a)[elem_nodes]
    rho_m_ip, rho_n_ip, rho_p_ip = g(rho_nm1) @ sf.T, g(rho_n) @ sf.T, g(rho_np1) @ sf.T            # [E,216]
    u_m = np.einsum("pn,enj->epj", sf, g(vel_nm1))
    u_n = np.einsum("pn,enj->epj", sf, g(vel_n))
    u_p = np.einsum("pn,enj->epj", sf, g(vel_np1))
    gjp = np.einsum("pn,enj->epj", sf, g(gpdx))
    lhs = np.zeros((E, 27, 3, 27, 3))
    rhs = np.zeros((E, 27, 3))
    for ip in range(NIP):
        nn = me.ip_node_map[ip]
        v = scv_vol[:, ip]
        fac = rho_p_ip[:, ip] * dens_fac + om_dens_fac
        rhs[:, nn] += -(gamma1 * rho_p_ip[:, ip, None] * u_p[:, ip] + gamma2 * rho_n_ip[:, ip, None] * u_n[:, ip]
                        + gamma3 * rho_m_ip[:, ip, None] * u_m[:, ip]) * (v / dt)[:, None] - gjp[:, ip] * (v * fac)[:, None]
        lhsfac = sf[ip][None, :] * (gamma1 * rho_p_ip[:, ip] * v / dt)[:, None]                      # [E,ic]
        for j in range(3):
            lhs[:, nn, j, :, j] += lhsfac
    return lhs.reshape(E, 81, 81), rhs.reshape(E, 81)


def assemble_csr(node_gid, elem_nodes, lhs, rhs, num_dof=3):
    """sumInto of every element into the single-rank CSR the reference builds: rows = node ids ascending (numDof rows per
    node), columns sorted by id within a row (src/TpetraLinearSystem.C:190-327,1257-1353,1457-1522).
    Returns (rowptr, colidx, values, rhs_vector) with aliases (periodic ids) merged."""
    gid = np.asarray(node_gid)
    uniq, inv = np.unique(gid, return_inverse=True)
    rows_node = inv[elem_nodes]                                          # [E,27]
    npe = elem_nodes.shape[1]
    dof = (rows_node[:, :, None] * num_dof + np.arange(num_dof)[None, None, :]).reshape(elem_nodes.shape[0], npe * num_dof)
    n = uniq.size * num_dof
    key = (dof[:, :, None].astype(np.int64) * n + dof[:, None, :]).ravel()
    ukey, kinv = np.unique(key, return_inverse=True)                     # sorted by (row, column): the CSR order
    values = np.bincount(kinv, weights=lhs.ravel(), minlength=ukey.size)
    rowptr = np.searchsorted(ukey // n, np.arange(n + 1)).astype(np.int64)
    b = np.zeros(n)
    np.add.at(b, dof.ravel(), rhs.ravel())
    return rowptr, (ukey % n).astype(np.int64), values, b
