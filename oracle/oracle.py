"""ctypes binding of liboracle.so (oracle/*.c).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KIND = dict(momentum_adv_diff=0, continuity_adv=1, scalar_adv_diff=2, momentum_mass=3,
            continuity_mass=4, scalar_mass=5, scalar_diff=6, test_fixed_block=7, momentum_buoyancy_src=8,
            momentum_buoyancy_boussinesq_src=9, tke_sst_src=10, sdr_sst_src=11, scalar_upw_adv_diff=12,
            momentum_upw_adv_diff=13, tke_sst_des_src=14, tke_ksgs_src=15, momentum_nso=16)


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("hex8_me.c", "elem_kernels.c", "linsys.c", "face_kernels.c", "nalu_oracle.h")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_linsys_create.restype = C.c_void_p
        for name in ("orc_linsys_rowptr", "orc_linsys_colidx", "orc_linsys_values", "orc_linsys_rhs",
                     "orc_linsys_row_gids", "orc_linsys_col_gids", "orc_linsys_node_row_lid",
                     "orc_linsys_node_col_lid"):
            getattr(_LIB, name).restype = C.c_void_p
        for name in ("orc_graph_arrays_create", "orc_graph_arrays_colidx", "orc_graph_arrays_rowptr"):
            getattr(_LIB, name).restype = C.c_void_p
        _LIB.orc_hex_volume_grandy.restype = C.c_double
        _LIB.orc_lcg_next.restype = C.c_double
    return _LIB


class KernelDesc(C.Structure):
    _fields_ = [("kind", C.c_int), ("shiftedGradOp", C.c_int), ("skewSymmetric", C.c_int),
                ("lumpedMass", C.c_int), ("shiftMdot", C.c_int), ("shiftPoisson", C.c_int),
                ("reducedSensitivities", C.c_int),
                ("includeDivU", C.c_double), ("densFac", C.c_double), ("om_densFac", C.c_double),
                ("dt", C.c_double), ("gamma1", C.c_double), ("gamma2", C.c_double), ("gamma3", C.c_double)] + \
               [(n, C.c_void_p) for n in ("coordinates", "velocity", "velocityN", "velocityNm1", "viscosity",
                                          "density", "densityN", "densityNm1", "pressure", "Gpdx", "scalarQ",
                                          "scalarQN", "scalarQNm1", "massFlowRate", "fixedBlock", "temperature")] + \
               [("gravity", C.c_double * 3), ("rhoRef", C.c_double), ("tRef", C.c_double), ("beta", C.c_double)] + \
               [(n, C.c_void_p) for n in ("tke", "sdr", "tvisc", "fOneBlend")] + \
               [(n, C.c_double) for n in ("betaStar", "tkeProdLimitRatio", "sigmaWTwo", "betaOne", "betaTwo",
                                          "gammaOne", "gammaTwo")] + \
               [(n, C.c_void_p) for n in ("Gjq", "Gju", "velocityRTM")] + \
               [(n, C.c_double) for n in ("alpha", "alphaUpw", "hoUpwind")] + \
               [("useLimiter", C.c_int), ("pecletTanh", C.c_int)] + \
               [(n, C.c_double) for n in ("hybridFactor", "tanhTrans", "tanhWidth")] + \
               [(n, C.c_void_p) for n in ("maxLengthScale", "cEps", "dsqrtkSq", "dualNodalVolume")] + \
               [(n, C.c_double) for n in ("cDESke", "cDESkw", "lrksgsfac", "fourthFac", "altResFac")]


_PTR_FIELDS = ("coordinates", "velocity", "velocityN", "velocityNm1", "viscosity", "density", "densityN",
               "densityNm1", "pressure", "Gpdx", "scalarQ", "scalarQN", "scalarQNm1", "massFlowRate",
               "fixedBlock", "temperature", "tke", "sdr", "tvisc", "fOneBlend", "Gjq", "Gju", "velocityRTM",
               "maxLengthScale", "cEps", "dsqrtkSq", "dualNodalVolume")


def _f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a


def make_desc(kind, keep, **kw):
    """Build a KernelDesc; numpy arrays given for pointer fields are kept alive in `keep`."""
    d = KernelDesc()
    d.kind = KIND[kind] if isinstance(kind, str) else int(kind)
    d.om_densFac = 1.0
    d.dt = 1.0
    d.gamma1 = 1.0
    d.gamma2 = -1.0
    d.gamma3 = 0.0
    for k, v in kw.items():
        if k in _PTR_FIELDS:
            if v is None:
                continue
            arr = _f64(v)
            keep.append(arr)
            setattr(d, k, arr.ctypes.data)
        elif k == "gravity":
            for j in range(3):
                d.gravity[j] = float(v[j])
        else:
            setattr(d, k, v)
    return d


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def scs_shape_fcn(shifted=False):
    loc = np.zeros(36)
    lib().orc_hex8_scs_intg_loc(int(shifted), _p(loc))
    out = np.zeros((12, 8))
    lib().orc_hex8_shape_fcn(12, _p(loc), _p(out))
    return out


def scv_shape_fcn(shifted=False):
    loc = np.zeros(24)
    lib().orc_hex8_scv_intg_loc(int(shifted), _p(loc))
    out = np.zeros((8, 8))
    lib().orc_hex8_shape_fcn(8, _p(loc), _p(out))
    return out


def scs_areav(coords):
    c = _f64(coords).reshape(8, 3)
    out = np.zeros((12, 3))
    lib().orc_hex8_scs_areav(_p(c), _p(out))
    return out


def scv_volume(coords):
    c = _f64(coords).reshape(8, 3)
    out = np.zeros(8)
    lib().orc_hex8_scv_volume(_p(c), _p(out))
    return out


def hex_volume_grandy(coords):
    c = _f64(coords).reshape(8, 3)
    return lib().orc_hex_volume_grandy(_p(c))


def scs_grad_op(coords, shifted=False):
    c = _f64(coords).reshape(8, 3)
    g = np.zeros((12, 8, 3))
    d = np.zeros((12, 8, 3))
    nbad = lib().orc_hex8_scs_grad_op(int(shifted), _p(c), _p(g), _p(d))
    return g, d, nbad


def scv_grad_op(coords, shifted=False):
    c = _f64(coords).reshape(8, 3)
    g = np.zeros((8, 8, 3))
    d = np.zeros((8, 8, 3))
    nbad = lib().orc_hex8_scv_grad_op(int(shifted), _p(c), _p(g), _p(d))
    return g, d, nbad


def lcg_sequence(seed, n):
    st = C.c_uint32(seed)
    out = np.empty(n)
    f = lib().orc_lcg_next
    for i in range(n):
        out[i] = f(C.byref(st))
    return out


def element_lhs_rhs(descs, num_dof, elem_nodes, elem_index=0):
    arr = (KernelDesc * len(descs))(*descs)
    n = 8 * num_dof
    lhs = np.zeros((n, n))
    rhs = np.zeros(n)
    en = np.ascontiguousarray(elem_nodes, dtype=np.int32)
    lib().orc_element_lhs_rhs(arr, len(descs), num_dof, _p(en), C.c_int64(elem_index), _p(lhs), _p(rhs))
    return lhs, rhs


def calc_mass_flow_rate_scs(elem_nodes, coords, density, velocity):
    en = np.ascontiguousarray(elem_nodes, dtype=np.int32).reshape(-1, 8)
    mdot = np.zeros((en.shape[0], 12))
    c, r, v = _f64(coords), _f64(density), _f64(velocity)
    lib().orc_calc_mass_flow_rate_scs(C.c_int64(en.shape[0]), _p(en), _p(c), _p(r), _p(v), _p(mdot))
    return mdot


def assemble_nodal_grad_elem(elem_nodes, coords, field, dual_volume, use_shifted=False, grad=None):
    """AssembleNodalGradElemAlgorithm / AssembleNodalGradUElemAlgorithm: interior Green-Gauss nodal gradient."""
    en = np.ascontiguousarray(elem_nodes, dtype=np.int32).reshape(-1, 8)
    f = _f64(field)
    ncomp = 1 if f.ndim == 1 else f.shape[1]
    n_nodes = f.shape[0]
    if grad is None:
        grad = np.zeros((n_nodes, 3 * ncomp))
    c, v = _f64(coords), _f64(dual_volume)
    lib().orc_assemble_nodal_grad_elem(C.c_int64(en.shape[0]), _p(en), _p(c), _p(f), int(ncomp), _p(v), int(use_shifted), _p(grad))
    return grad


def compute_mdot_elem(elem_nodes, coords, density, velocity, pressure, Gpdx, proj_time_scale, shift_mdot=False,
                      shift_poisson=False):
    """ComputeMdotElemAlgorithm::execute (src/ComputeMdotElemAlgorithm.C:85-254)."""
    en = np.ascontiguousarray(elem_nodes, dtype=np.int32).reshape(-1, 8)
    mdot = np.zeros((en.shape[0], 12))
    a = [_f64(x) for x in (coords, density, velocity, pressure, Gpdx)]
    lib().orc_compute_mdot_elem(C.c_int64(en.shape[0]), _p(en), *[_p(x) for x in a], C.c_double(proj_time_scale),
                                int(shift_mdot), int(shift_poisson), _p(mdot))
    return mdot


# ---- face/element boundary-condition path (oracle/face_kernels.c) ---------------------------------------------------
FACE_KIND = {"momentum_symmetry": 0, "scalar_open_adv": 1, "continuity_open": 2, "momentum_open_adv_diff": 3,
             "momentum_wall_function": 4}
_FACE_PTR_FIELDS = ("coordinates", "velocity", "velocityRTM", "meshVelocity", "viscosity", "density", "pressure", "pressureBc",
                    "Gpdx", "scalarQ", "bcScalarQ", "Gjq", "Gjui", "velocityBc", "exposedAreaVec", "openMassFlowRate",
                    "dynamicPressure")
_FACE_WALL_PTR_FIELDS = ("wallVelocityBc", "wallFrictionVelocityBip", "wallNormalDistanceBip")


class FaceKernelDesc(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("kind", "shiftedGradOp", "skewSymmetric", "shiftMdot", "reducedSensitivities", "meshMoves",
                                       "pecletTanh")] + \
               [(n, C.c_double) for n in ("includeDivU", "dt", "gamma1", "alphaUpw", "hoUpwind", "hybridFactor", "tanhTrans",
                                          "tanhWidth")] + \
               [(n, C.c_void_p) for n in _FACE_PTR_FIELDS] + \
               [(n, C.c_void_p) for n in _FACE_WALL_PTR_FIELDS] + [(n, C.c_double) for n in ("elog", "kappa", "yplusCrit")]


def make_face_desc(kind, keep, **kw):
    """Build a FaceKernelDesc; numpy arrays given for pointer fields are kept alive in `keep`."""
    d = FaceKernelDesc()
    d.kind = FACE_KIND[kind] if isinstance(kind, str) else int(kind)
    d.dt = 1.0
    d.gamma1 = 1.0
    for k, v in kw.items():
        if k in _FACE_PTR_FIELDS or k in _FACE_WALL_PTR_FIELDS:
            if v is None:
                continue
            arr = _f64(v)
            keep.append(arr)
            setattr(d, k, arr.ctypes.data)
        else:
            setattr(d, k, v)
    return d


def quad3d_scs_areav(coords):
    c = _f64(coords).reshape(4, 3)
    out = np.zeros((4, 3))
    lib().orc_quad3d_scs_areav(_p(c), _p(out))
    return out


def quad3d_shape_fcn(shifted=False):
    out = np.zeros((4, 4))
    lib().orc_quad3d_shape_fcn(int(shifted), _p(out))
    return out


def hex8_face_grad_op(face_ordinal, coords, shifted=False):
    c = _f64(coords).reshape(8, 3)
    out = np.zeros((4, 8, 3))
    bad = lib().orc_hex8_face_grad_op(int(face_ordinal), int(shifted), _p(c), _p(out))
    return out, bad


def hex8_face_tables():
    """(side_node_ordinals[6][4], face ipNodeMap[6][4], opposingNodes[6][4])"""
    L = lib()
    L.orc_hex8_side_node_ordinals.restype = C.POINTER(C.c_int)
    L.orc_hex8_face_ip_node_map.restype = C.POINTER(C.c_int)
    sno = np.array([[L.orc_hex8_side_node_ordinals(f)[k] for k in range(4)] for f in range(6)])
    ipn = np.array([[L.orc_hex8_face_ip_node_map(f)[k] for k in range(4)] for f in range(6)])
    opp = np.array([[L.orc_hex8_opposing_node(f, k) for k in range(4)] for f in range(6)])
    return sno, ipn, opp


def compute_exposed_area(face_elem, face_ordinal, elem_nodes, coords):
    fe = np.ascontiguousarray(face_elem, dtype=np.int32)
    fo = np.ascontiguousarray(face_ordinal, dtype=np.int32)
    en = np.ascontiguousarray(elem_nodes, dtype=np.int32)
    c = _f64(coords)
    out = np.zeros((fe.size, 4, 3))
    lib().orc_compute_exposed_area(C.c_int64(fe.size), _p(fe), _p(fo), _p(en), _p(c), _p(out))
    return out


def face_elem_lhs_rhs(descs, num_dof, elem_nodes, face_ordinal, face_index=0):
    arr = (FaceKernelDesc * len(descs))(*descs)
    en = np.ascontiguousarray(elem_nodes, dtype=np.int32)
    n = 8 * num_dof
    lhs, rhs = np.zeros((n, n)), np.zeros(n)
    lib().orc_face_elem_lhs_rhs(arr, len(descs), int(num_dof), _p(en), int(face_ordinal), C.c_int64(face_index), _p(lhs), _p(rhs))
    return lhs, rhs


class LinSys:
    """All ranks of one simulated MPI job (orc_linsys)."""

    def __init__(self, rank_meshes, num_dof):
        self.L = lib()
        self.nranks = len(rank_meshes)
        self.num_dof = num_dof
        self.h = C.c_void_p(self.L.orc_linsys_create(self.nranks, num_dof))
        self.meshes = rank_meshes
        for r, m in enumerate(rank_meshes):
            gid = np.ascontiguousarray(m.node_gid, dtype=np.int64)
            own = np.ascontiguousarray(m.node_owner, dtype=np.int32)
            en = np.ascontiguousarray(m.elem_nodes, dtype=np.int32)
            self.L.orc_linsys_set_rank_mesh(self.h, r, C.c_int32(gid.size), _p(gid), _p(own),
                                            C.c_int64(en.shape[0]), _p(en))
        self.L.orc_linsys_build_elem_to_node_graph(self.h)
        self.L.orc_linsys_finalize(self.h)

    def __del__(self):
        try:
            self.L.orc_linsys_destroy(self.h)
        except Exception:
            pass

    def _arr(self, ptr, n, dtype):
        if n == 0:
            return np.zeros(0, dtype=dtype)
        buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype, count=n)

    def n_owned_rows(self, r):
        return self.L.orc_linsys_num_owned_rows(self.h, r)

    def n_shared_rows(self, r):
        return self.L.orc_linsys_num_shared_rows(self.h, r)

    def n_cols(self, r):
        return self.L.orc_linsys_num_cols(self.h, r)

    def rowptr(self, r, shared=False):
        n = self.n_shared_rows(r) if shared else self.n_owned_rows(r)
        return self._arr(self.L.orc_linsys_rowptr(self.h, r, int(shared)), n + 1, np.uint64)

    def colidx(self, r, shared=False):
        nnz = int(self.rowptr(r, shared)[-1])
        return self._arr(self.L.orc_linsys_colidx(self.h, r, int(shared)), nnz, np.int32)

    def values(self, r, shared=False):
        nnz = int(self.rowptr(r, shared)[-1])
        return self._arr(self.L.orc_linsys_values(self.h, r, int(shared)), nnz, np.float64)

    def rhs(self, r, shared=False):
        n = self.n_shared_rows(r) if shared else self.n_owned_rows(r)
        return self._arr(self.L.orc_linsys_rhs(self.h, r, int(shared)), n, np.float64)

    def row_gids(self, r, shared=False):
        n = self.n_shared_rows(r) if shared else self.n_owned_rows(r)
        return self._arr(self.L.orc_linsys_row_gids(self.h, r, int(shared)), n, np.int64)

    def col_gids(self, r):
        return self._arr(self.L.orc_linsys_col_gids(self.h, r), self.n_cols(r), np.int64)

    def node_row_lid(self, r):
        return self._arr(self.L.orc_linsys_node_row_lid(self.h, r), self.meshes[r].node_gid.size, np.int32)

    def node_col_lid(self, r):
        return self._arr(self.L.orc_linsys_node_col_lid(self.h, r), self.meshes[r].node_gid.size, np.int32)

    def zero(self):
        self.L.orc_linsys_zero(self.h)

    def assemble_elem(self, r, descs):
        arr = (KernelDesc * len(descs))(*descs)
        self.L.orc_assemble_elem(self.h, r, arr, len(descs))

    def assemble_face_elem(self, r, descs, face_elem, face_ordinal):
        """AssembleFaceElemSolverAlgorithm::execute over the faces (element, ordinal) of rank r"""
        arr = (FaceKernelDesc * len(descs))(*descs)
        fe = np.ascontiguousarray(face_elem, dtype=np.int32)
        fo = np.ascontiguousarray(face_ordinal, dtype=np.int32)
        en = np.ascontiguousarray(self.meshes[r].elem_nodes, dtype=np.int32)
        self.L.orc_assemble_face_elem(self.h, r, int(self.num_dof), _p(en), arr, len(descs), C.c_int64(fe.size), _p(fe), _p(fo))

    def assemble_node_mass(self, r, desc, dual_nodal_volume):
        dv = _f64(dual_nodal_volume)
        self.L.orc_assemble_node_mass(self.h, r, C.byref(desc), _p(dv))

    def sum_into(self, r, entities, rhs, lhs):
        en = np.ascontiguousarray(entities, dtype=np.int32)
        rr, ll = _f64(rhs), _f64(lhs)
        self.L.orc_linsys_sum_into(self.h, r, en.size, _p(en), _p(rr), _p(ll))

    def load_complete(self):
        self.L.orc_linsys_load_complete(self.h)


class GraphArrays:
    """LocalGraphArrays test hook."""

    def __init__(self, row_lens):
        rl = np.ascontiguousarray(row_lens, dtype=np.uint64)
        self.n = rl.size
        self.nnz = int(rl.sum())
        self.h = C.c_void_p(lib().orc_graph_arrays_create(_p(rl), C.c_size_t(rl.size)))

    def insert(self, row, inds, num_dof):
        a = np.ascontiguousarray(inds, dtype=np.int32)
        lib().orc_graph_arrays_insert(self.h, C.c_size_t(row), C.c_size_t(a.size), _p(a), num_dof)

    def colidx(self):
        ptr = lib().orc_graph_arrays_colidx(self.h)
        return np.frombuffer((C.c_char * (4 * self.nnz)).from_address(ptr), dtype=np.int32).copy()

    def rowptr(self):
        ptr = lib().orc_graph_arrays_rowptr(self.h)
        return np.frombuffer((C.c_char * (8 * (self.n + 1))).from_address(ptr), dtype=np.uint64).copy()

    def __del__(self):
        try:
            lib().orc_graph_arrays_destroy(self.h)
        except Exception:
            pass
