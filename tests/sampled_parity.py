"""Sampled-row parity at sizes where the oracle cannot assemble the whole system ("sampled rows +
global checksums at 256^3").  TEST INFRASTRUCTURE: used by the `-m gpu` tests and by bench.py's `checks` entry, never on
the product path.

A row of the assembled matrix depends only on the (at most) 8 elements around its node, so the oracle can reproduce it
exactly from a small sub-box that contains those elements: sub-boxes are cut out of the block at its corners, edges,
faces and interior, assembled with the oracle (oracle/elem_kernels.c + oracle/linsys.c: the reference's gather ->
kernel -> sumInto restated), and every COMPLETE row of a sub-box (all of the node's elements inside it) is compared
entry by entry with the device CSR, columns matched by node.  Plus one global invariant of every row, computed on the
device with torch: sum_c lhs[(r,i),(c,j)] = delta_ij * (net mass flow out of the node's control volume), because the
shape-function derivatives of an element sum to zero and the advection shape functions to one
(src/kernel/MomentumAdvDiffElemKernel.C:138-184).
"""
import numpy as np

import oracle
from nalu_b200 import mesh as nmesh


def subbox_origins(nx, ny, nz, size, n_random=16, seed=7):
    """corners, edge mid-points, face centres, and random interior sub-boxes (element origin of each)."""
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
    return sorted(out)


def check_sampled_rows(ls, m, host_fields, mdot_dev, kernels, ndof=3, size=5, n_random=16, row_floor=1e-2, shift=True):
    """ls: DeviceCsrLinearSystem holding the assembled system of one-rank box mesh `m`; host_fields: {oracle field
    name: numpy nodal array}; mdot_dev: device [nElem][12] (or None); kernels: [(oracle kind, {oracle scalar/option
    kwargs})] in list order.  Returns {"rows", "subboxes", "max_rel_err_values", "max_rel_err_rhs"}: relative error of
    an entry is taken against max(|a|, |b|, row_floor * max|row|) (per-row floor so that entries that cancel
    to zero do not explode)."""
    import torch
    assert m.box is not None and m.nranks == 1
    nx, ny, nz = m.box.nx, m.box.ny, m.box.nz
    NX, NY = nx + 1, ny + 1
    dev = ls.values.device
    rowptr, colidx = ls.device_graph_tensors(False)
    row_lid, _ = ls.node_lids()
    worst_v = worst_r = 0.0
    nrows = 0
    origins = subbox_origins(nx, ny, nz, size, n_random)
    for (i0, j0, k0) in origins:
        sx, sy, sz = min(size, nx - i0), min(size, ny - j0), min(size, nz - k0)
        sub = nmesh.generated_box(sx, sy, sz)
        # local node / element indices of the sub-box inside the block
        ii, jj, kk = np.meshgrid(np.arange(sx + 1), np.arange(sy + 1), np.arange(sz + 1), indexing="ij")
        ii, jj, kk = (a.transpose(2, 1, 0).ravel() for a in (ii, jj, kk))
        gn = (i0 + ii) + NX * ((j0 + jj) + NY * (k0 + kk))
        ei, ej, ek = np.meshgrid(np.arange(sx), np.arange(sy), np.arange(sz), indexing="ij")
        ei, ej, ek = (a.transpose(2, 1, 0).ravel() for a in (ei, ej, ek))
        ge = (i0 + ei) + nx * ((j0 + ej) + ny * (k0 + ek))
        kw = {k: np.ascontiguousarray(v[gn]) for k, v in host_fields.items()}
        # The assembled system is translation invariant, the reference's rounding error is not: it differences absolute
        # coordinates (Jacobian = sum_n dN_n x_n), which costs |x|/h ulps -- 256 of them at the far corner of this box.
        # The oracle therefore sees the sub-box in coordinates relative to its first node (the product shifts every
        # element to its own first node, box5_tile.cuh coef_load); `shift=False` measures the unshifted difference.
        if shift:
            kw["coordinates"] = kw["coordinates"] - kw["coordinates"][0]
        sub.coords = kw["coordinates"]
        if mdot_dev is not None:
            kw["massFlowRate"] = np.ascontiguousarray(mdot_dev[torch.as_tensor(ge, device=dev)].cpu().numpy())
        keep = []
        descs = [oracle.make_desc(kind, keep, **dict(kw, **extra)) for kind, extra in kernels]
        orc = oracle.LinSys([sub], ndof)
        orc.assemble_elem(0, descs)
        orp, oci, ov, orhs = orc.rowptr(0).astype(np.int64), orc.colidx(0), orc.values(0), orc.rhs(0)
        # complete nodes: every element around the node (that exists in the block) lies inside the sub-box
        def complete(a, a0, s, n):
            return ((a > 0) | (a0 == 0)) & ((a < s) | (a0 + s == n))
        ok = complete(ii, i0, sx, nx) & complete(jj, j0, sy, ny) & complete(kk, k0, sz, nz)
        nodes = np.nonzero(ok)[0]
        # product rows of those nodes, fetched in one gather
        prow = (row_lid[gn[nodes]].astype(np.int64)[:, None] + np.arange(ndof)[None, :]).ravel()
        prow_t = torch.as_tensor(prow, device=dev)
        p0 = rowptr[prow_t].cpu().numpy()
        p1 = rowptr[prow_t + 1].cpu().numpy()
        lens = p1 - p0
        idx = np.concatenate([np.arange(a, b) for a, b in zip(p0, p1)])
        idx_t = torch.as_tensor(idx, device=dev)
        pv = ls.values[idx_t].cpu().numpy()
        pc = colidx[idx_t].cpu().numpy()
        prhs = ls.rhs[prow_t].cpu().numpy()
        off = np.concatenate([[0], np.cumsum(lens)])
        sub_col_to_prod = (row_lid[gn].astype(np.int64)[:, None] + np.arange(ndof)[None, :]).ravel()  # sub col lid -> product col lid (node_lids are dof-level)
        t = 0
        for s in nodes:
            for i in range(ndof):
                r = s * ndof + i
                oc = sub_col_to_prod[oci[orp[r]:orp[r + 1]]]
                ovals = ov[orp[r]:orp[r + 1]]
                a, b = off[t], off[t + 1]
                cols, vals = pc[a:b], pv[a:b]
                assert cols.size == oc.size, "row length differs from the oracle's graph"
                o1, o2 = np.argsort(oc, kind="stable"), np.argsort(cols, kind="stable")
                assert np.array_equal(oc[o1], cols[o2]), "row columns differ from the oracle's graph"
                x, y = ovals[o1], vals[o2]
                scale = max(np.abs(x).max(), 1e-300)
                e = np.max(np.abs(x - y) / np.maximum(np.maximum(np.abs(x), np.abs(y)), row_floor * scale))
                worst_v = max(worst_v, float(e))
                # rhs_i = -(row . u) + ...: its rounding error scales with the row, not with itself
                rs = max(abs(orhs[r]), abs(prhs[t]), row_floor * float(np.abs(x).sum()) * 1e-2, 1e-300)
                worst_r = max(worst_r, abs(orhs[r] - prhs[t]) / rs)
                t += 1
                nrows += 1
    return {"rows": nrows, "subboxes": len(origins), "subbox_elements": size, "max_rel_err_values": worst_v,
            "max_rel_err_rhs": worst_r, "row_floor": row_floor}


def row_sum_identity(ls, m, mdot_dev, chunk_planes=16):
    """Momentum adv-diff (any viscosity, any divU option) on a one-rank box: for every dof row (r,i) and column dof j,
    sum_c lhs[(r,i),(c,j)] = delta_ij * M_r with M_r = sum over the node's integration points of the outward mass flow.
    Returns the largest violation relative to the row's largest absolute entry.  Device-side (torch), chunked over node planes."""
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
