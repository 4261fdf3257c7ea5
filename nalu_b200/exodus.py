"""Exodus-II (netCDF-3 classic) Hex8 mesh reader -> flat arrays of `nalu_b200.mesh.HexMesh`.

The reference reads its meshes through STK-IO / Ioss (src/Realm.C `ioBroker_->create_input_mesh`); what the assembly
path sees of a mesh is node ids, ownership, element->node connectivity and coordinates.  This module produces those
arrays straight from the file: every element block whose `elem_type` is HEX / HEX8, nodes numbered through
`node_num_map` (the ids STK uses as `naluGlobalId` when no periodic aliasing is active), elements through
`elem_num_map`.  A decomposed file (`name.g.P.r`, written by the SEACAS `decomp` tool the reference's regression
decks use) is one rank's piece: its maps carry the global ids, so `read_exodus_pieces` can stitch node ownership
(lowest rank that holds the node, the STK rule seen in unit_tests/UnitTestTpetra.C:149-156).

netCDF-4 / HDF5 based files (`*.exo` of newer SEACAS versions) are not netCDF-3 and are rejected with a clear error.
`write_minimal_exodus` writes the variables this reader consumes (used for the committed test fixtures).
"""
import numpy as np

from .mesh import HexMesh


def _open(path):
    from scipy.io import netcdf_file
    try:
        return netcdf_file(path, "r", mmap=False)
    except Exception as e:  # HDF5 signature or garbage
        raise ValueError(f"{path}: not a netCDF-3 classic Exodus file ({e})") from e


def _coords(f, n):
    v = f.variables
    if "coord" in v:  # older files: one [num_dim][num_nodes] array
        c = np.array(v["coord"][:], dtype=np.float64)
        return np.ascontiguousarray(c.T)
    cols = [np.array(v[k][:], dtype=np.float64) for k in ("coordx", "coordy", "coordz") if k in v]
    if len(cols) != 3:
        raise ValueError("Exodus file is not three-dimensional")
    return np.ascontiguousarray(np.stack(cols, axis=1))


def read_exodus_hex8(path, rank=0, nranks=1):
    """All HEX8 element blocks of one Exodus file as a one-rank HexMesh (node_owner = rank).  Elements keep the file's
    block order (the order STK buckets them in for a one-block mesh), nodes the file's node order."""
    f = _open(path)
    try:
        d = f.dimensions
        if int(d.get("num_dim", 0)) != 3:
            raise ValueError(f"{path}: num_dim = {d.get('num_dim')}, the Hex8 path needs 3")
        nn = int(d["num_nodes"])
        coords = _coords(f, nn)
        v = f.variables
        gid = np.array(v["node_num_map"][:], dtype=np.int64) if "node_num_map" in v else np.arange(1, nn + 1, dtype=np.int64)
        conn = []
        b = 1
        while f"connect{b}" in v:
            c = v[f"connect{b}"]
            et = getattr(c, "elem_type", b"")
            et = (et.decode() if isinstance(et, bytes) else str(et)).strip().upper()
            arr = np.array(c[:], dtype=np.int64)
            if et in ("HEX", "HEX8") and arr.shape[1] == 8:
                conn.append(arr - 1)
            elif arr.size:
                raise ValueError(f"{path}: element block {b} is {et or '?'} with {arr.shape[1]} nodes; only HEX8 is supported")
            b += 1
        if not conn:
            raise ValueError(f"{path}: no HEX8 element block")
        conn = np.concatenate(conn, axis=0)
        ne = conn.shape[0]
        egid = np.array(v["elem_num_map"][:], dtype=np.int64) if "elem_num_map" in v else np.arange(1, ne + 1, dtype=np.int64)
    finally:
        f.close()
    if conn.min() < 0 or conn.max() >= nn:
        raise ValueError(f"{path}: connectivity out of range")
    return HexMesh(coords=coords, node_gid=gid, node_owner=np.full(nn, rank, dtype=np.int32),
                   elem_nodes=np.ascontiguousarray(conn.astype(np.int32)), rank=rank, nranks=nranks, box=None,
                   elem_gid=egid[:ne])


def read_exodus_pieces(base, nranks):
    """The `base.P.r` pieces of a decomposed mesh (r = 0..P-1, zero padded like SEACAS does) as per-rank meshes with
    node ownership = lowest rank holding the node."""
    width = len(str(nranks))
    parts = [read_exodus_hex8(f"{base}.{nranks}.{str(r).zfill(width)}", r, nranks) for r in range(nranks)]
    owner = {}
    for m in parts:
        for g in m.node_gid.tolist():
            if g not in owner:
                owner[g] = m.rank
    for m in parts:
        m.node_owner = np.array([owner[g] for g in m.node_gid.tolist()], dtype=np.int32)
        order = np.argsort(m.node_gid, kind="stable")  # ranks hold their nodes in ascending id order (mesh.partition_elements)
        inv = np.empty_like(order)
        inv[order] = np.arange(order.size)
        m.coords = m.coords[order]
        m.node_gid = m.node_gid[order]
        m.node_owner = m.node_owner[order]
        m.elem_nodes = inv[m.elem_nodes].astype(np.int32)
    return parts


def write_minimal_exodus(path, mesh, title="nalu_b200 fixture"):
    """coordx/y/z, connect1 (elem_type HEX8), node_num_map, elem_num_map: what read_exodus_hex8 consumes."""
    from scipy.io import netcdf_file
    f = netcdf_file(path, "w", version=1)
    f.title = title
    nn, ne = mesh.n_nodes, mesh.n_elems
    f.createDimension("num_dim", 3)
    f.createDimension("num_nodes", nn)
    f.createDimension("num_elem", ne)
    f.createDimension("num_el_blk", 1)
    f.createDimension("num_el_in_blk1", ne)
    f.createDimension("num_nod_per_el1", 8)
    for k, name in enumerate(("coordx", "coordy", "coordz")):
        v = f.createVariable(name, "d", ("num_nodes",))
        v[:] = mesh.coords[:, k]
    c = f.createVariable("connect1", "i", ("num_el_in_blk1", "num_nod_per_el1"))
    c.elem_type = "HEX8"
    c[:] = mesh.elem_nodes.astype(np.int32) + 1
    v = f.createVariable("node_num_map", "i", ("num_nodes",))
    v[:] = mesh.node_gid.astype(np.int32)
    v = f.createVariable("elem_num_map", "i", ("num_elem",))
    v[:] = (mesh.elem_gid if mesh.elem_gid is not None else np.arange(1, ne + 1)).astype(np.int32)
    f.close()
