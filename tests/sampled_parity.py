"""Sampled-row parity at sizes where the oracle cannot assemble the whole system ("sampled rows +
global checksums at 256^3").  TEST INFRASTRUCTURE: used by the `-m gpu` tests and by bench.py's `checks` entry, never on
the product path.

A row of the assembled matrix depends only on the (at most) 8 elements around its node, so the oracle can reproduce it
exactly from a small sub-box that contains those elements: sub-boxes are cut out of the GLOBAL block at its corners,
edges, faces, interior and (multi-rank) across the rank interfaces, re-created from their global node ids (coordinates:
lattice + the id-keyed perturbation; fields: the analytic functions of the coordinates), assembled with the oracle
(oracle/elem_kernels.c + oracle/linsys.c: the reference's gather -> kernel -> sumInto restated), and every COMPLETE row
of a sub-box (all of the node's elements inside it) that this rank owns is compared entry by entry with the device CSR,
columns matched by global id.  On several ranks this is a check of the assembly AND of loadComplete on the timed mesh.

Plus one global invariant of every row, computed on the device with torch: sum_c lhs[(r,i),(c,j)] = delta_ij * (net mass
flow out of the node's control volume), because the shape-function derivatives of an element sum to zero and the
advection shape functions to one (src/kernel/MomentumAdvDiffElemKernel.C:138-184).
"""
import numpy as np

import oracle
from nalu_b200 import mesh as nmesh


def subbox_origins(nx, ny, nz, size, n_random=16, seed=7, z_interfaces=()):
    """corners, edge mid-points, face centres, random interior sub-boxes and sub-boxes centred on the rank interfaces
    (element origin of each)."""
    def pos(n):
        return [0, max(0, (n - size) // 2), max(0, n - size)]
    out = set()
    for i in pos(nx):
        for j in pos(ny):
            for k in pos(nz):
                out.add((i, j, k))
    rng = np.random.default_rng(seed)
    for _ in range(n_random):
        out.add(tuple(int(rng.integers(0, max(1, n - size + 1))) for n in (nx, ny, nz)))
    for kz in z_interfaces:
        k0 = min(max(0, kz - size // 2 - 1), max(0, nz - size))
        for i in pos(nx):
            for j in pos(ny):
                out.add((i, j, k0))
    return sorted(out)


def make_subbox(nx, ny, nz, origin, size, amp, coord_map=None):
    """the sub-box [origin, origin + size) of the global generated box as a one-rank mesh with the id-keyed coordinate
    perturbation of the benchmark / test meshes; returns (mesh, global node ids, which nodes have all their elements in
    the sub-box)"""
    i0, j0, k0 = origin
    sx, sy, sz = min(size, nx - i0), min(size, ny - j0), min(size, nz - k0)
    sub = nmesh.generated_box(sx, sy, sz)
    n = np.arange(sub.n_nodes)
    ii, jj, kk = n % (sx + 1), (n // (sx + 1)) % (sy + 1), n // ((sx + 1) * (sy + 1))
    gi, gj, gk = i0 + ii, j0 + jj, k0 + kk
    gid = (1 + gi + (nx + 1) * (gj + (ny + 1) * gk.astype(np.int64))).astype(np.int64)
    sub.coords = np.stack([gi, gj, gk], axis=1).astype(np.float64)
    true_gid, sub.node_gid = sub.node_gid, gid
    nmesh.perturb_coords_hash(sub, amp)
    sub.node_gid = true_gid  # the oracle's graph of the sub-box uses local ids; `gid` maps them back
    if coord_map is not None:
        sub.coords = coord_map(sub.coords)  # e.g. the stretched-z map of config 4 (lattice -> physical coordinates)
    complete = np.ones(sub.n_nodes, dtype=bool)
    for a, a0, s, ng in ((ii, i0, sx, nx), (jj, j0, sy, ny), (kk, k0, sz, nz)):
        complete &= ((a > 0) | (a0 == 0)) & ((a < s) | (a0 + s == ng))
    return sub, gid, complete


def check_sampled_rows(ls, m, nz_global, field_fn, kernel_fn, ndof=3, size=5, n_random=16, row_floor=1e-2, amp=0.12, shift=True,
                       sol_scale=1.0, coord_map=None):
    """ls: DeviceCsrLinearSystem holding the assembled (and, on several ranks, loadComplete'd) system of rank mesh `m`, a
    z-slab of the global nx x ny x nz_global generated box.  field_fn(coords) -> {name: nodal array} recreates the nodal
    fields of any set of nodes; kernel_fn(sub_mesh, fields) -> list of oracle descriptors (keeps its arrays alive).
    Relative error of a matrix entry is taken against max(|a|, |b|, row_floor * max|row|); of a residual
    entry against max(|a|, |b|, row_floor * sum|row| * sol_scale): the residual is -(row . u) - (time terms), every term of
    the size of a row entry times the solution scale, the sum of them far smaller on smooth fields.  The oracle sees every
    sub-box in coordinates relative to its first node unless shift=False: the assembled system is translation invariant,
    the reference's rounding error is not (it differences absolute coordinates: |x|/h ulps)."""
    import torch
    assert m.box is not None
    nx, ny = m.box.nx, m.box.ny
    kz0 = m.box.k0
    dev = ls.values.device
    rowptr, colidx = ls.device_graph_tensors(False)
    row_lid, _ = ls.node_lids()
    col_gids = ls.col_gids()
    n_owned = ls.n_owned_rows
    interfaces = [k for (k, _) in nmesh.slab_ranges(nz_global, m.nranks)][1:] if m.nranks > 1 else []
    origins = subbox_origins(nx, ny, nz_global, size, n_random, z_interfaces=interfaces)
    worst_v = worst_r = 0.0
    nrows = nsub = 0
    NXl, NYl = nx + 1, ny + 1
    for origin in origins:
        sub, gid, complete = make_subbox(nx, ny, nz_global, origin, size, amp, coord_map)
        # global -> local node of this rank; rows this rank owns
        gi, gj, gk = (gid - 1) % NXl, ((gid - 1) // NXl) % NYl, (gid - 1) // (NXl * NYl)
        lk = gk - kz0
        local = (lk >= 0) & (lk <= m.box.nz)
        ln = np.where(local, gi + NXl * (gj + NYl * np.clip(lk, 0, m.box.nz)), 0)
        owned = local & (row_lid[ln] < n_owned)
        nodes = np.nonzero(complete & owned)[0]
        if nodes.size == 0:
            continue
        nsub += 1
        f = field_fn(sub.coords)
        if shift:
            f = dict(f, coordinates=sub.coords - sub.coords[0])
        descs = kernel_fn(sub, f)
        orc = oracle.LinSys([sub], ndof)
        orc.assemble_elem(0, descs)
        orp, oci, ov, orhs = orc.rowptr(0).astype(np.int64), orc.colidx(0), orc.values(0), orc.rhs(0)
        prow = (row_lid[ln[nodes]].astype(np.int64)[:, None] + np.arange(ndof)[None, :]).ravel()
        prow_t = torch.as_tensor(prow, device=dev)
        p0 = rowptr[prow_t].cpu().numpy()
        p1 = rowptr[prow_t + 1].cpu().numpy()
        idx_t = torch.as_tensor(np.concatenate([np.arange(a, b) for a, b in zip(p0, p1)]), device=dev)
        pv = ls.values[idx_t].cpu().numpy()
        pc = col_gids[colidx[idx_t].cpu().numpy()]
        prhs = ls.rhs[prow_t].cpu().numpy()
        off = np.concatenate([[0], np.cumsum(p1 - p0)])
        sub_col_gid = (ndof * (gid - 1)[:, None] + np.arange(ndof)[None, :] + 1).ravel()  # GID_ of the sub-box's dof columns
        t = 0
        for s in nodes:
            for i in range(ndof):
                r = s * ndof + i
                oc = sub_col_gid[oci[orp[r]:orp[r + 1]]]
                ovals = ov[orp[r]:orp[r + 1]]
                cols, vals = pc[off[t]:off[t + 1]], pv[off[t]:off[t + 1]]
                o1, o2 = np.argsort(oc, kind="stable"), np.argsort(cols, kind="stable")
                assert cols.size == oc.size and np.array_equal(oc[o1], cols[o2]), "row columns differ from the oracle's graph"
                x, y = ovals[o1], vals[o2]
                scale = max(np.abs(x).max(), 1e-300)
                worst_v = max(worst_v, float(np.max(np.abs(x - y) / np.maximum(np.maximum(np.abs(x), np.abs(y)), row_floor * scale))))
                rs = max(abs(orhs[r]), abs(prhs[t]), row_floor * sol_scale * float(np.abs(x).sum()), 1e-300)
                worst_r = max(worst_r, float(abs(orhs[r] - prhs[t]) / rs))
                t += 1
                nrows += 1
    return {"rows": nrows, "subboxes": nsub, "subbox_elements": size, "max_rel_err_values": worst_v,
            "max_rel_err_rhs": worst_r, "row_floor": row_floor}


def row_sum_identity(ls, m, mdot_dev, chunk_planes=16):
    """Momentum adv-diff alone (any viscosity, any divU option) on a one-rank box: for every dof row (r,i) and column
    dof j, sum_c lhs[(r,i),(c,j)] = delta_ij * M_r with M_r = sum over the node's integration points of the outward mass
    flow.  Returns the largest violation relative to the row's largest absolute entry.  Device-side (torch), chunked
    over node planes."""
    import torch
    assert m.box is not None and m.nranks == 1
    nx, ny, nz = m.box.nx, m.box.ny, m.box.nz
    NX, NY, NZ = nx + 1, ny + 1, nz + 1
    dev = ls.values.device
    # net outward mass flow of every node: ip (L,R): +mdot to L, -mdot to R (Hex8CVFEM.C:264-276)
    L = torch.tensor([0, 1, 2, 0, 4, 5, 6, 4, 0, 1, 2, 3], device=dev)
    R = torch.tensor([1, 2, 3, 3, 5, 6, 7, 7, 4, 5, 6, 7], device=dev)
    en = torch.as_tensor(m.elem_nodes.astype(np.int64), device=dev)
    M = torch.zeros(m.n_nodes, dtype=torch.float64, device=dev)
    M.index_add_(0, en[:, L].reshape(-1), mdot_dev.reshape(-1))
    M.index_add_(0, en[:, R].reshape(-1), -mdot_dev.reshape(-1))
    del en
    rowptr, _ = ls.device_graph_tensors(False)
    worst = 0.0
    rows_per_plane = 3 * NX * NY
    for z0 in range(0, NZ, chunk_planes):
        z1 = min(NZ, z0 + chunk_planes)
        r0, r1 = z0 * rows_per_plane, z1 * rows_per_plane
        a, b = int(rowptr[r0]), int(rowptr[r1])
        v = ls.values[a:b].view(-1, 3)
        lens = ((rowptr[r0 + 1:r1 + 1] - rowptr[r0:r1]) // 3)
        S = torch.segment_reduce(v, "sum", lengths=lens, axis=0)            # [rows, 3]
        amax = torch.segment_reduce(v.abs().amax(dim=1), "max", lengths=lens)  # [rows]
        rows = torch.arange(r0, r1, device=dev)
        node, i = rows // 3, rows % 3
        expect = torch.zeros_like(S)
        expect[torch.arange(rows.numel(), device=dev), i] = M[node]
        viol = ((S - expect).abs().amax(dim=1) / amax.clamp_min(1e-300)).max()
        worst = max(worst, float(viol))
    return worst
