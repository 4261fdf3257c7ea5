"""DeviceCsrLinearSystem -- Python face of the C ABI's linear-system half.

Mirrors the call sequence the reference drives through its LinearSystem interface
(include/LinearSystem.h:39-167; EquationSystem::assemble_and_solve, src/EquationSystem.C:263-307):

    buildElemToNodeGraph(parts) ; finalizeLinearSystem()      once per mesh
    zeroSystem() ; <assemble algorithms> ; loadComplete()     every nonlinear iteration

torch is used only to hold device memory, streams and (multi-GPU) the process group.
"""
import ctypes as C

import numpy as np

from . import lib as _lib
from .lib import F, FF, FK, FN, K, MODE, OPT, FaceKernelDesc, GraphInfo, KernelDesc, check, ptr


def _exchange_single(send):
    return {}


class DeviceCsrLinearSystem:
    def __init__(self, mesh, num_dof, exchange=None, host_only=False, use_box=True):
        """mesh: nalu_b200.mesh.HexMesh of this rank.  exchange(send: {dest: bytes}) -> {src: bytes}
        performs the neighbour handshake (None for a single rank)."""
        self.L = _lib.lib()
        self.mesh = mesh
        self.num_dof = num_dof
        self.h = C.c_void_p()
        gid = np.ascontiguousarray(mesh.node_gid, dtype=np.int64)
        own = np.ascontiguousarray(mesh.node_owner, dtype=np.int32)
        en = np.ascontiguousarray(mesh.elem_nodes, dtype=np.int32)
        box = (mesh.box.nx, mesh.box.ny, mesh.box.nz) if (use_box and mesh.box is not None) else (0, 0, 0)
        self.nodes_per_elem = int(en.shape[1])
        if self.nodes_per_elem == 8:
            check(self.L.nalu_b200_graph_create(C.byref(self.h), mesh.rank, mesh.nranks, num_dof, gid.size, ptr(gid),
                                                ptr(own), C.c_int64(en.shape[0]), ptr(en), *box), "graph_create")
        else:  # Hex27 (mesh.hex27_box): AlgTraitsHex27
            check(self.L.nalu_b200_graph_create_topo(C.byref(self.h), mesh.rank, mesh.nranks, num_dof, self.nodes_per_elem,
                                                     gid.size, ptr(gid), ptr(own), C.c_int64(en.shape[0]), ptr(en)),
                  "graph_create_topo")
        # --- handshake -------------------------------------------------------------------
        send = {}
        for p in range(mesh.nranks):
            if p == mesh.rank:
                continue
            n = C.c_int64()
            check(self.L.nalu_b200_graph_send_size(self.h, p, C.byref(n)), "graph_send_size")
            if n.value:
                buf = np.empty(n.value // 8, dtype=np.int64)
                check(self.L.nalu_b200_graph_send_data(self.h, p, ptr(buf)), "graph_send_data")
                send[p] = buf
        self.handshake_send = send
        recv = (exchange or _exchange_single)(send) if mesh.nranks > 1 else {}
        for src, buf in recv.items():
            buf = np.ascontiguousarray(buf, dtype=np.int64)
            check(self.L.nalu_b200_graph_recv_data(self.h, src, ptr(buf), C.c_int64(buf.size * 8)), "graph_recv_data")
        self.host_only = host_only
        if host_only:
            check(self.L.nalu_b200_graph_finalize_host_only(self.h), "graph_finalize_host_only")
        else:
            check(self.L.nalu_b200_graph_finalize(self.h), "graph_finalize")
        self.info = GraphInfo()
        check(self.L.nalu_b200_graph_get_info(self.h, C.byref(self.info)), "graph_get_info")
        self.values = None
        self.rhs = None
        self._keep = []

    def __del__(self):
        try:
            if self.h:
                self.L.nalu_b200_graph_destroy(self.h)
                self.h = C.c_void_p()
        except Exception:
            pass

    # ---- graph inspection (host copies) ---------------------------------------------------
    @property
    def n_owned_rows(self):
        return self.info.n_owned_rows

    @property
    def n_shared_rows(self):
        return self.info.n_shared_rows

    @property
    def nnz_owned(self):
        return self.info.nnz_owned

    @property
    def nnz_shared(self):
        return self.info.nnz_shared

    def rowptr(self, shared=False):
        n = self.n_shared_rows if shared else self.n_owned_rows
        out = np.empty(n + 1, dtype=np.uint64)
        check(self.L.nalu_b200_graph_copy_rowptr(self.h, int(shared), ptr(out)), "copy_rowptr")
        return out

    def colidx(self, shared=False):
        out = np.empty(self.nnz_shared if shared else self.nnz_owned, dtype=np.int32)
        check(self.L.nalu_b200_graph_copy_colidx(self.h, int(shared), ptr(out)), "copy_colidx")
        return out

    def row_gids(self, shared=False):
        out = np.empty(self.n_shared_rows if shared else self.n_owned_rows, dtype=np.int64)
        check(self.L.nalu_b200_graph_copy_row_gids(self.h, int(shared), ptr(out)), "copy_row_gids")
        return out

    def col_gids(self):
        out = np.empty(self.info.n_cols, dtype=np.int64)
        check(self.L.nalu_b200_graph_copy_col_gids(self.h, ptr(out)), "copy_col_gids")
        return out

    def node_lids(self):
        r = np.empty(self.mesh.n_nodes, dtype=np.int32)
        c = np.empty(self.mesh.n_nodes, dtype=np.int32)
        check(self.L.nalu_b200_graph_copy_node_lids(self.h, ptr(r), ptr(c)), "copy_node_lids")
        return r, c

    def elem_colors(self):
        """(number of colours, colour of every element) of the general kernel's scatter; (0, None) without a colouring."""
        n = C.c_int32()
        out = np.full(self.mesh.n_elems, -1, dtype=np.int32)
        check(self.L.nalu_b200_graph_copy_elem_colors(self.h, C.byref(n), ptr(out)), "copy_elem_colors")
        return n.value, (out if n.value else None)

    def device_graph(self, shared=False):
        """(rowptr, colidx) device addresses for a device-side consumer (solver)."""
        rp, ci = C.c_void_p(), C.c_void_p()
        check(self.L.nalu_b200_graph_device_ptrs(self.h, int(shared), C.byref(rp), C.byref(ci)), "device_ptrs")
        return rp.value, ci.value

    def device_graph_tensors(self, shared=False):
        """(rowptr int64[nRows+1], colidx int32[nnz]) as torch views of the device-resident graph (no copy)."""
        import torch
        rp, ci = self.device_graph(shared)
        n = self.n_shared_rows if shared else self.n_owned_rows
        nnz = self.nnz_shared if shared else self.nnz_owned

        def view(addr, count, typestr):
            class _Mem:
                pass
            m = _Mem()
            m.__cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr, "data": (int(addr), False), "version": 3}
            return torch.as_tensor(m, device=self.device)
        return view(rp, n + 1, "<i8"), view(ci, max(nnz, 1), "<i4")[:nnz]

    # ---- storage -----------------------------------------------------------------------------
    def allocate(self, device="cuda", raw=False):
        """values / rhs storage.  raw=True takes the memory straight from cudaMalloc (through the C ABI)
        instead of torch's caching allocator, so that a CUDA-IPC handle of the tensor maps exactly this
        buffer in a peer process (dist.PeerWindows)."""
        import torch
        self.device = torch.device(device)
        nv, nr = self.nnz_owned + self.nnz_shared, self.n_owned_rows + self.n_shared_rows
        if raw:
            self._raw = []
            self.values = self._raw_tensor(nv)
            self.rhs = self._raw_tensor(nr)
        else:
            self.values = torch.empty(nv, dtype=torch.float64, device=self.device)
            self.rhs = torch.empty(nr, dtype=torch.float64, device=self.device)
        return self

    def _raw_tensor(self, n):
        import torch
        p = C.c_void_p()
        check(self.L.nalu_b200_malloc(C.byref(p), C.c_size_t(max(n, 1) * 8)), "malloc")

        class _Mem:
            pass
        m = _Mem()
        m.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (p.value, False), "version": 3}
        t = torch.as_tensor(m, device=self.device)
        self._raw.append((p, m))
        return t

    def _stream(self):
        import torch
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def zero_system(self):
        """TpetraLinearSystem::zeroSystem"""
        check(self.L.nalu_b200_zero_system(self.h, ptr(self.values), ptr(self.rhs), self._stream()), "zero_system")

    # ---- assembly ------------------------------------------------------------------------------
    def make_desc(self, kind, fields=None, options=(), mass_flow_rate=None, fixed_block=None, **scalars):
        """fields: {slot name: torch tensor on the device}."""
        d = KernelDesc()
        d.kind = K[kind]
        opt = 0
        for o in options:
            opt |= OPT[o]
        d.options = opt
        d.include_divu = scalars.get("include_divu", 0.0)
        d.dens_fac = scalars.get("dens_fac", 0.0)
        d.om_dens_fac = scalars.get("om_dens_fac", 1.0)
        d.dt = scalars.get("dt", 1.0)
        d.gamma1 = scalars.get("gamma1", 1.0)
        d.gamma2 = scalars.get("gamma2", -1.0)
        d.gamma3 = scalars.get("gamma3", 0.0)
        g = scalars.get("gravity", (0.0, 0.0, 0.0))
        d.gravity[0], d.gravity[1], d.gravity[2] = float(g[0]), float(g[1]), float(g[2])
        d.rho_ref = scalars.get("rho_ref", 0.0)
        d.t_ref = scalars.get("t_ref", 0.0)
        d.beta = scalars.get("beta", 0.0)
        d.nso_fourth_fac = scalars.get("fourth_fac", 0.0)
        d.nso_alt_res_fac = scalars.get("alt_res_fac", 0.0)
        for name, v in (scalars.get("turb") or {}).items():
            d.turb[_lib.TM[name]] = float(v)
        for name, v in (scalars.get("upw") or {}).items():
            d.upw[_lib.UPW[name]] = float(v)
        for name, t in (fields or {}).items():
            if t is None:
                continue
            self._keep.append(t)
            d.nodal[F[name]] = t.data_ptr()
        if mass_flow_rate is not None:
            self._keep.append(mass_flow_rate)
            d.mass_flow_rate = mass_flow_rate.data_ptr()
        if fixed_block is not None:
            self._keep.append(fixed_block)
            d.fixed_block = fixed_block.data_ptr()
        return d

    def assemble_elem(self, descs, mode="auto", overwrite=False):
        """AssembleElemSolverAlgorithm::execute for the registered kernel list."""
        arr = (KernelDesc * len(descs))(*descs)
        check(self.L.nalu_b200_assemble_elem(self.h, arr, len(descs), ptr(self.values), ptr(self.rhs), MODE[mode],
                                             int(overwrite), self._stream()), "assemble_elem")
        return self.L.nalu_b200_last_assemble_mode()

    def assemble_elem_zrange(self, descs, z_begin, z_end, overwrite=True, stream=None):
        """Rows of node planes [z_begin, z_end) of a structured block (tiled kernel only)."""
        arr = (KernelDesc * len(descs))(*descs)
        st = self._stream() if stream is None else C.c_void_p(stream)
        check(self.L.nalu_b200_assemble_elem_zrange(self.h, arr, len(descs), ptr(self.values), ptr(self.rhs),
                                                    int(overwrite), int(z_begin), int(z_end), st), "assemble_elem_zrange")

    def assemble_node_mass(self, desc):
        check(self.L.nalu_b200_assemble_node_mass(self.h, C.byref(desc), ptr(self.values), ptr(self.rhs),
                                                  self._stream()), "assemble_node_mass")

    def assemble_nodal_grad_elem(self, coords, field, dual_volume, grad, use_shifted=False):
        """AssembleNodalGrad(U)ElemAlgorithm::execute: interior contributions accumulated into `grad`."""
        ncomp = 1 if field.dim() == 1 else int(field.shape[1])
        check(self.L.nalu_b200_assemble_nodal_grad_elem(self.h, ptr(coords), ptr(field), ncomp, ptr(dual_volume), int(use_shifted),
                                                        ptr(grad), self._stream()), "assemble_nodal_grad_elem")
        return grad

    def compute_mdot_elem(self, coords, density, velocity, pressure, dpdx, proj_time_scale, options=(), out=None, e_range=None):
        """ComputeMdotElemAlgorithm::execute: mass_flow_rate_scs of every locally-owned element (or of elements e_range)."""
        import torch
        if out is None:
            out = torch.empty((self.mesh.n_elems, 12 if self.nodes_per_elem == 8 else 216), dtype=torch.float64, device=coords.device)
        opt = 0
        for o in options:
            opt |= OPT[o]
        e0, e1 = e_range if e_range is not None else (0, self.mesh.n_elems)
        check(self.L.nalu_b200_compute_mdot_elem_range(self.h, ptr(coords), ptr(density), ptr(velocity), ptr(pressure), ptr(dpdx),
                                                       C.c_double(proj_time_scale), opt, C.c_int64(e0), C.c_int64(e1), ptr(out),
                                                       self._stream()), "compute_mdot_elem")
        return out

    def apply_dirichlet_bcs(self, nodes, solution, bc_values, begin_pos, end_pos):
        """LinearSystem::applyDirichletBCs; nodes: int32 device tensor of local node ids."""
        check(self.L.nalu_b200_apply_dirichlet_bcs(self.h, int(nodes.numel()), ptr(nodes), ptr(solution), ptr(bc_values),
                                                   int(begin_pos), int(end_pos), ptr(self.values), ptr(self.rhs),
                                                   self._stream()), "apply_dirichlet_bcs")

    def reset_rows(self, nodes, begin_pos, end_pos):
        """LinearSystem::resetRows."""
        check(self.L.nalu_b200_reset_rows(self.h, int(nodes.numel()), ptr(nodes), int(begin_pos), int(end_pos),
                                          ptr(self.values), ptr(self.rhs), self._stream()), "reset_rows")

    # ---- face/element boundary-condition algorithms (AssembleFaceElemSolverAlgorithm) ---------------------------------
    def face_set(self, face_elem, face_ordinal):
        """The locally owned faces of a boundary part: (local element, Hex8 face ordinal) pairs (host arrays)."""
        return FaceSet(self, face_elem, face_ordinal)

    def make_face_desc(self, kind, fields=None, face_fields=None, options=(), **scalars):
        """fields: {nodal slot name: device tensor}; face_fields: {face field name: device tensor [n_faces][4][..]}"""
        d = FaceKernelDesc()
        d.kind = FK[kind]
        opt = 0
        for o in options:
            opt |= _lib.OPT_MESH_MOVES if o == "mesh_moves" else OPT[o]
        d.options = opt
        d.include_divu = scalars.get("include_divu", 0.0)
        d.dt = scalars.get("dt", 1.0)
        d.gamma1 = scalars.get("gamma1", 1.0)
        for name, v in (scalars.get("upw") or {}).items():
            d.upw[_lib.UPW[name]] = float(v)
        for name, v in (scalars.get("wall") or {}).items():
            d.wall[_lib.WALL[name]] = float(v)
        for name, t in (fields or {}).items():
            if t is None or name not in FN:
                continue
            self._keep.append(t)
            d.nodal[FN[name]] = t.data_ptr()
        for name, t in (face_fields or {}).items():
            self._keep.append(t)
            d.face[FF[name]] = t.data_ptr()
        return d

    def assemble_face_elem(self, face_set, descs):
        """AssembleFaceElemSolverAlgorithm::execute: every kernel of the list on every face of the set."""
        arr = (FaceKernelDesc * len(descs))(*descs)
        check(self.L.nalu_b200_assemble_face_elem(self.h, face_set.h, arr, len(descs), ptr(self.values), ptr(self.rhs),
                                                  self._stream()), "assemble_face_elem")

    def mdot_scs(self, coords, density, velocity, out=None):
        import torch
        if out is None:
            out = torch.empty((self.mesh.n_elems, 12 if self.nodes_per_elem == 8 else 216), dtype=torch.float64, device=coords.device)
        check(self.L.nalu_b200_mdot_scs(self.h, ptr(coords), ptr(density), ptr(velocity), ptr(out), self._stream()),
              "mdot_scs")
        return out

    def solve(self, tol=1e-10, max_iter=1000):
        """LinearSystem::solve (src/TpetraLinearSystem.C:1841-1926) on the device-resident CSR, one rank: returns
        (solution tensor, iterations, relative residual); see nalu_b200/solve.py"""
        from . import solve as _solve
        return _solve.solve(self, tol=tol, max_iter=max_iter)

    def geometry_errors(self):
        f = C.c_int32()
        check(self.L.nalu_b200_geometry_errors(self.h, C.byref(f)), "geometry_errors")
        return f.value

    # ---- views ---------------------------------------------------------------------------------
    def owned_values(self):
        return self.values[: self.nnz_owned]

    def shared_values(self):
        return self.values[self.nnz_owned:]

    def owned_rhs(self):
        return self.rhs[: self.n_owned_rows]

    def shared_rhs(self):
        return self.rhs[self.n_owned_rows:]

    # ---- loadComplete ----------------------------------------------------------------------------
    def halo_peers(self):
        ns, nr = C.c_int32(), C.c_int32()
        check(self.L.nalu_b200_halo_num_peers(self.h, C.byref(ns), C.byref(nr)), "halo_num_peers")
        send, recv = [], []
        for i in range(ns.value):
            p, n = C.c_int32(), C.c_int64()
            check(self.L.nalu_b200_halo_send_peer(self.h, i, C.byref(p), C.byref(n)), "halo_send_peer")
            send.append((p.value, n.value))
        for i in range(nr.value):
            p, n = C.c_int32(), C.c_int64()
            check(self.L.nalu_b200_halo_recv_peer(self.h, i, C.byref(p), C.byref(n)), "halo_recv_peer")
            recv.append((p.value, n.value))
        return send, recv

    def halo_pack(self, i, buf):
        check(self.L.nalu_b200_halo_pack(self.h, i, ptr(self.values), ptr(self.rhs), ptr(buf), self._stream()),
              "halo_pack")

    def halo_unpack_add(self, i, buf):
        check(self.L.nalu_b200_halo_unpack_add(self.h, i, ptr(buf), ptr(self.values), ptr(self.rhs), self._stream()),
              "halo_unpack_add")

    def halo_pull_add(self, i, peer_shared_values_ptr, peer_shared_rhs_ptr):
        check(self.L.nalu_b200_halo_pull_add(self.h, i, C.c_void_p(peer_shared_values_ptr),
                                             C.c_void_p(peer_shared_rhs_ptr), ptr(self.values), ptr(self.rhs),
                                             self._stream()), "halo_pull_add")

    def load_complete(self, comm, stream=None):
        """LinearSystem::loadComplete inside the library (nalu_b200_load_complete): `comm` from nalu_b200.dist.make_comm."""
        st = self._stream() if stream is None else C.c_void_p(stream)
        check(self.L.nalu_b200_load_complete(self.h, comm, ptr(self.values), ptr(self.rhs), st), "load_complete")

    def load_complete_nccl(self, group=None):
        """loadComplete over torch.distributed: pack -> batched isend/irecv -> unpack-add."""
        import torch
        import torch.distributed as dist
        send, recv = self.halo_peers()
        if not send and not recv:
            return
        sbufs = [torch.empty(n, dtype=torch.float64, device=self.values.device) for _, n in send]
        rbufs = [torch.empty(n, dtype=torch.float64, device=self.values.device) for _, n in recv]
        for i in range(len(send)):
            self.halo_pack(i, sbufs[i])
        ops = [dist.P2POp(dist.isend, sbufs[i], send[i][0], group=group) for i in range(len(send))]
        ops += [dist.P2POp(dist.irecv, rbufs[i], recv[i][0], group=group) for i in range(len(recv))]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        for i in range(len(recv)):
            self.halo_unpack_add(i, rbufs[i])
        self._keep_halo = (sbufs, rbufs)


class FaceSet:
    """nalu_b200_face_set: the faces of one boundary part, coloured for the atomic-free scatter."""

    def __init__(self, ls, face_elem, face_ordinal):
        self.ls = ls
        self.L = ls.L
        fe = np.ascontiguousarray(face_elem, dtype=np.int32)
        fo = np.ascontiguousarray(face_ordinal, dtype=np.int32)
        assert fe.shape == fo.shape
        self.n_faces = int(fe.size)
        self.h = C.c_void_p()
        check(self.L.nalu_b200_face_set_create(C.byref(self.h), ls.h, C.c_int64(fe.size), ptr(fe), ptr(fo)), "face_set_create")

    def __del__(self):
        try:
            if self.h:
                self.L.nalu_b200_face_set_destroy(self.h)
                self.h = C.c_void_p()
        except Exception:
            pass

    def num_colors(self):
        n = C.c_int32()
        check(self.L.nalu_b200_face_set_num_colors(self.h, C.byref(n)), "face_set_num_colors")
        return n.value

    def exposed_area(self, coords, out=None):
        """ComputeGeometryBoundaryAlgorithm::execute: exposed_area_vector [n_faces][4][3] on the device"""
        import torch
        if out is None:
            out = torch.empty((self.n_faces, 4, 3), dtype=torch.float64, device=coords.device)
        check(self.L.nalu_b200_compute_exposed_area(self.ls.h, self.h, ptr(coords), ptr(out), self.ls._stream()),
              "compute_exposed_area")
        return out
