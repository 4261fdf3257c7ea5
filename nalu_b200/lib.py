"""ctypes binding of libnalu_b200.so -- the C ABI declared in include/nalu_b200.h.

This is the only door into the product's compute path.  If the library is missing it is
built in-tree (nvcc, sm_100a); if that fails the import fails loudly: there is no CPU
fallback for the assembly path.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

_LIB = None

OK = 0
ERR_NO_DEVICE = 3

K = dict(momentum_adv_diff=0, continuity_adv=1, scalar_adv_diff=2, momentum_mass=3, continuity_mass=4,
         scalar_mass=5, scalar_diff=6, test_fixed_block=7, momentum_buoyancy_src=8,
         momentum_buoyancy_boussinesq_src=9, tke_sst_src=10, sdr_sst_src=11, scalar_upw_adv_diff=12,
         momentum_upw_adv_diff=13, tke_sst_des_src=14, tke_ksgs_src=15, momentum_nso=16)
OPT = dict(shifted_grad_op=1, skew_symmetric=2, lumped_mass=4, shift_mdot=8, shift_poisson=16, reduced_sens=32,
           use_limiter=64, peclet_tanh=128)
F = dict(coordinates=0, velocity=1, velocity_n=2, velocity_nm1=3, viscosity=4, density=5, density_n=6,
         density_nm1=7, pressure=8, dpdx=9, scalarq=10, scalarq_n=11, scalarq_nm1=12, dual_volume=13, temperature=14, tke=15, sdr=16,
         tvisc=17, f_one_blend=18, gjq=19, gju=20, velocity_rtm=21, sst_max_length_scale=22, c_epsilon=23,
         dsqrtk_dx_sq=24)
F_COUNT = 26
MODE = dict(auto=0, general=1, box=2)


class KernelDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("options", C.c_int32), ("include_divu", C.c_double),
                ("dens_fac", C.c_double), ("om_dens_fac", C.c_double), ("dt", C.c_double),
                ("gamma1", C.c_double), ("gamma2", C.c_double), ("gamma3", C.c_double),
                ("nodal", C.c_void_p * F_COUNT), ("mass_flow_rate", C.c_void_p), ("fixed_block", C.c_void_p),
                ("gravity", C.c_double * 3), ("rho_ref", C.c_double), ("t_ref", C.c_double), ("beta", C.c_double),
                ("turb", C.c_double * 10), ("upw", C.c_double * 8),
                ("nso_fourth_fac", C.c_double), ("nso_alt_res_fac", C.c_double)]


UPW = dict(alpha=0, alpha_upw=1, ho_upwind=2, hybrid_factor=3, tanh_trans=4, tanh_width=5)
TM = dict(beta_star=0, tke_prod_limit_ratio=1, sigma_w_two=2, beta_one=3, beta_two=4, gamma_one=5, gamma_two=6, cdes_kw=7, cdes_ke=8,
          lrksgs_fac=9)


# face/element boundary-condition kernels (nalu_b200_face_kernel_desc)
FK = dict(momentum_symmetry=0, scalar_open_adv=1, continuity_open=2, momentum_open_adv_diff=3, momentum_wall_function=4)
FN = dict(coordinates=0, velocity=1, velocity_rtm=2, mesh_velocity=3, viscosity=4, density=5, pressure=6, pressure_bc=7, dpdx=8,
          scalarq=9, bc_scalarq=10, gjq=11, gjui=12, velocity_bc=13, wall_velocity_bc=14)
FN_COUNT = 16
FF = dict(exposed_area_vector=0, open_mass_flow_rate=1, dynamic_pressure=2, wall_friction_velocity_bip=3,
          wall_normal_distance_bip=4)
FF_COUNT = 6
WALL = dict(elog=0, kappa=1, yplus_crit=2)
OPT_MESH_MOVES = 1 << 8


class FaceKernelDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("options", C.c_int32), ("include_divu", C.c_double), ("dt", C.c_double),
                ("gamma1", C.c_double), ("upw", C.c_double * 8), ("nodal", C.c_void_p * FN_COUNT), ("face", C.c_void_p * FF_COUNT),
                ("wall", C.c_double * 4)]


class GraphInfo(C.Structure):
    _fields_ = [("num_dof", C.c_int32), ("n_owned_rows", C.c_int32), ("n_shared_rows", C.c_int32),
                ("n_cols", C.c_int32), ("nnz_owned", C.c_int64), ("nnz_shared", C.c_int64),
                ("n_elems", C.c_int64), ("n_nodes", C.c_int32), ("is_box", C.c_int32)]


# every symbol include/nalu_b200.h declares (checked by tests/test_capi_symbols.py)
SYMBOLS = [
    "nalu_b200_last_error", "nalu_b200_version", "nalu_b200_launch_count", "nalu_b200_graph_create", "nalu_b200_graph_create_topo",
    "nalu_b200_graph_destroy", "nalu_b200_graph_send_size", "nalu_b200_graph_send_data",
    "nalu_b200_graph_recv_data", "nalu_b200_graph_finalize", "nalu_b200_graph_finalize_host_only",
    "nalu_b200_graph_get_info", "nalu_b200_graph_copy_rowptr", "nalu_b200_graph_copy_colidx",
    "nalu_b200_graph_copy_row_gids", "nalu_b200_graph_copy_col_gids", "nalu_b200_graph_copy_node_lids",
    "nalu_b200_graph_device_ptrs", "nalu_b200_graph_copy_elem_colors", "nalu_b200_zero_system", "nalu_b200_assemble_elem", "nalu_b200_assemble_elem_zrange",
    "nalu_b200_last_assemble_mode", "nalu_b200_tune", "nalu_b200_geometry_errors", "nalu_b200_sum_into", "nalu_b200_assemble_node_mass",
    "nalu_b200_mdot_scs", "nalu_b200_compute_mdot_elem", "nalu_b200_compute_mdot_elem_range", "nalu_b200_assemble_nodal_grad_elem", "nalu_b200_apply_dirichlet_bcs", "nalu_b200_reset_rows", "nalu_b200_halo_num_peers", "nalu_b200_halo_send_peer", "nalu_b200_halo_recv_peer",
    "nalu_b200_halo_pack", "nalu_b200_halo_unpack_add", "nalu_b200_halo_pull_add",
    "nalu_b200_halo_peer_offsets", "nalu_b200_comm_get_unique_id", "nalu_b200_comm_create", "nalu_b200_comm_destroy",
    "nalu_b200_load_complete", "nalu_b200_ipc_get_handle", "nalu_b200_ipc_open_handle",
    "nalu_b200_ipc_close", "nalu_b200_malloc", "nalu_b200_free", "nalu_b200_memcpy_h2d",
    "nalu_b200_memcpy_d2h", "nalu_b200_stream_sync",
    "nalu_b200_face_set_create", "nalu_b200_face_set_destroy", "nalu_b200_face_set_num_colors", "nalu_b200_compute_exposed_area",
    "nalu_b200_assemble_face_elem",
]


class NaluB200Error(RuntimeError):
    pass


def lib():
    global _LIB
    if _LIB is None:
        so = _build.OUT
        if not os.path.exists(so):
            so = _build.build()
        L = C.CDLL(so)
        L.nalu_b200_last_error.restype = C.c_char_p
        L.nalu_b200_launch_count.restype = C.c_int64
        L.nalu_b200_graph_destroy.restype = None
        L.nalu_b200_face_set_destroy.restype = None
        L.nalu_b200_comm_destroy.restype = None
        _LIB = L
    return _LIB


def tune(**knobs):
    """nalu_b200_tune: launch-configuration knobs (box_kernel, box_zchunk, split_lists, cont_tiled)."""
    for k, v in knobs.items():
        check(lib().nalu_b200_tune(k.encode(), C.c_int32(int(v))), "tune")


def check(rc, what=""):
    if rc != OK:
        msg = lib().nalu_b200_last_error().decode()
        raise NaluB200Error(f"{what} failed (status {rc}): {msg}")


def ptr(a):
    """void* of a numpy array or torch tensor (device or host)."""
    if a is None:
        return None
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    return a.ctypes.data_as(C.c_void_p)
