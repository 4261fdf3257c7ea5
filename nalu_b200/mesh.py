"""Flat-array Hex8 mesh substrate: generators, perturbation, partitioning.

The reference reads its meshes through STK-IO / Ioss; the hot path only ever sees, per
rank, node ids (`naluGlobalId`), node ownership, element->node connectivity and nodal /
element fields.  This module produces exactly those arrays.

  * `generated_box`   -- Ioss `generated:NxMxK` convention (unit-spaced nodes, id
                         1+i+(N+1)j+(N+1)(M+1)k, Exodus Hex8 node order; confirmed by
                         unit_tests/UnitTestTpetra.C:97-98 "elem 1 nodes: 1 2 4 3 5 6 8 7").
  * `perturb_coords_lcg` -- unit_tests/UnitTestUtils.C:37-72 `perturb_coord_hex_8`.
  * `partition_slabs` -- z-slab element decomposition (Ioss default for generated meshes,
                         unit_tests/UnitTestUtils.C:30-35), shared nodes owned by the lowest
                         sharing rank (STK rule, visible in unit_tests/UnitTestTpetra.C:149-156).
  * `periodic_box`    -- periodic in x,y,z by id aliasing: a subject node keeps its own
                         coordinates but adopts its monarch's id (src/PeriodicManager.C:698-724).
"""
from dataclasses import dataclass, field
from typing import Optional, Tuple

import numpy as np


@dataclass
class BoxInfo:
    """Logical structure of a generated box (local to one rank): node (i,j,k) has local
    index i + (nx+1)*j + (nx+1)*(ny+1)*k; element (i,j,k) has local index i + nx*j + nx*ny*k."""
    nx: int
    ny: int
    nz: int
    k0: int = 0          # global z-index of this rank's first element layer
    nz_global: int = 0


@dataclass
class HexMesh:
    coords: np.ndarray          # [nNode,3] f64
    node_gid: np.ndarray        # [nNode] int64, 1-based naluGlobalId (aliases allowed: periodic)
    node_owner: np.ndarray      # [nNode] int32
    elem_nodes: np.ndarray      # [nElem,8] int32 local node indices, Exodus Hex8 order
    rank: int = 0
    nranks: int = 1
    box: Optional[BoxInfo] = None
    elem_gid: Optional[np.ndarray] = None

    @property
    def n_nodes(self):
        return int(self.node_gid.size)

    @property
    def n_elems(self):
        return int(self.elem_nodes.shape[0])


def _box_conn(nx, ny, nz):
    i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    # element order: i fastest, then j, then k
    i = i.transpose(2, 1, 0).ravel()
    j = j.transpose(2, 1, 0).ravel()
    k = k.transpose(2, 1, 0).ravel()
    sx, sy = nx + 1, (nx + 1) * (ny + 1)
    n0 = i + sx * j + sy * k
    conn = np.stack([n0, n0 + 1, n0 + 1 + sx, n0 + sx,
                     n0 + sy, n0 + 1 + sy, n0 + 1 + sx + sy, n0 + sx + sy], axis=1)
    return conn.astype(np.int32)


def _box_coords(nx, ny, nz, k0=0):
    x = np.arange(nx + 1, dtype=np.float64)
    y = np.arange(ny + 1, dtype=np.float64)
    z = np.arange(k0, k0 + nz + 1, dtype=np.float64)
    Z, Y, X = np.meshgrid(z, y, x, indexing="ij")
    return np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)


def generated_box(nx, ny, nz) -> HexMesh:
    """Ioss `generated:nx x ny x nz` on one rank."""
    nn = (nx + 1) * (ny + 1) * (nz + 1)
    return HexMesh(coords=_box_coords(nx, ny, nz),
                   node_gid=np.arange(1, nn + 1, dtype=np.int64),
                   node_owner=np.zeros(nn, dtype=np.int32),
                   elem_nodes=_box_conn(nx, ny, nz),
                   box=BoxInfo(nx, ny, nz, 0, nz),
                   elem_gid=np.arange(1, nx * ny * nz + 1, dtype=np.int64))


def slab_ranges(nz, nranks):
    """Element-layer ranges [k0,k1) per rank: as even as possible, extra layers to low ranks."""
    base, rem = divmod(nz, nranks)
    out, k = [], 0
    for r in range(nranks):
        n = base + (1 if r < rem else 0)
        out.append((k, k + n))
        k += n
    return out


def generated_box_rank(nx, ny, nz, rank, nranks) -> HexMesh:
    """Rank-local piece of `generated:nx x ny x nz` under the z-slab decomposition, built
    directly (no global mesh is materialised: 256^3 on 8 ranks must not need 17M-node globals)."""
    k0, k1 = slab_ranges(nz, nranks)[rank]
    lz = k1 - k0
    assert lz > 0, "more ranks than element layers"
    sxy = (nx + 1) * (ny + 1)
    nn = sxy * (lz + 1)
    gid = np.arange(1, nn + 1, dtype=np.int64) + sxy * k0
    owner = np.full(nn, rank, dtype=np.int32)
    if rank > 0:
        owner[:sxy] = rank - 1  # bottom plane shared with (and owned by) the lower rank
    eg = np.arange(1, nx * ny * lz + 1, dtype=np.int64) + nx * ny * k0
    return HexMesh(coords=_box_coords(nx, ny, lz, k0), node_gid=gid, node_owner=owner,
                   elem_nodes=_box_conn(nx, ny, lz), rank=rank, nranks=nranks,
                   box=BoxInfo(nx, ny, lz, k0, nz), elem_gid=eg)


def partition_slabs(mesh: HexMesh, nranks: int):
    """Split a one-rank generated box into z-slab rank meshes (general path: works from the
    arrays, keeps whatever coordinates the mesh carries)."""
    assert mesh.box is not None
    nx, ny, nz = mesh.box.nx, mesh.box.ny, mesh.box.nz
    sxy = (nx + 1) * (ny + 1)
    out = []
    for r, (k0, k1) in enumerate(slab_ranges(nz, nranks)):
        lo, hi = sxy * k0, sxy * (k1 + 1)
        m = generated_box_rank(nx, ny, nz, r, nranks)
        m.coords = mesh.coords[lo:hi].copy()
        out.append(m)
    return out


def partition_elements(mesh: HexMesh, elem_rank: np.ndarray, nranks: int):
    """General element partition of an arbitrary one-rank Hex8 or Hex27 mesh: rank r gets the elements
    with elem_rank==r and the closure of their nodes (ascending gid); a shared node is owned by
    the lowest rank that touches it."""
    ne = mesh.n_elems
    node_min_rank = np.full(mesh.n_nodes, nranks, dtype=np.int32)
    np.minimum.at(node_min_rank, mesh.elem_nodes.ravel(), np.repeat(elem_rank.astype(np.int32), mesh.elem_nodes.shape[1]))
    out = []
    for r in range(nranks):
        es = np.nonzero(elem_rank == r)[0]
        conn = mesh.elem_nodes[es]
        nodes = np.unique(conn.ravel())
        nodes = nodes[np.argsort(mesh.node_gid[nodes], kind="stable")]
        g2l = np.full(mesh.n_nodes, -1, dtype=np.int32)
        g2l[nodes] = np.arange(nodes.size, dtype=np.int32)
        out.append(HexMesh(coords=mesh.coords[nodes].copy(), node_gid=mesh.node_gid[nodes].copy(),
                           node_owner=node_min_rank[nodes].copy(), elem_nodes=g2l[conn],
                           rank=r, nranks=nranks, box=None,
                           elem_gid=(mesh.elem_gid[es].copy() if mesh.elem_gid is not None else es + 1)))
    assert ne == sum(m.n_elems for m in out)
    return out


def periodic_box(nx, ny, nz) -> HexMesh:
    """Box periodic in all three directions: nodes on the +x/+y/+z faces are subjects whose
    naluGlobalId is their monarch's; they keep their own coordinates (so geometry is sane) and
    must carry copies of the monarch's field values."""
    m = generated_box(nx, ny, nz)
    sx, sy = nx + 1, (nx + 1) * (ny + 1)
    idx = np.arange(m.n_nodes)
    i, j, k = idx % sx, (idx // sx) % (ny + 1), idx // sy
    mi, mj, mk = i % nx, j % ny, k % nz
    monarch = mi + sx * mj + sy * mk
    m.node_gid = m.node_gid[monarch].copy()
    m.box = None  # aliasing breaks the lexicographic id structure the box fast path relies on
    m.periodic_monarch = monarch.astype(np.int32)
    return m


_LCG_A, _LCG_C, _LCG_M = 1103515245, 12345, 2147483647


def lcg_stream(seed: int, n: int) -> np.ndarray:
    """n values of the reference's LCG (unit_tests/UnitTestUtils.C:41-54) in [-1,1)."""
    out = np.empty(n, dtype=np.float64)
    prev = int(seed)
    for t in range(n):
        prev = (_LCG_A * prev + _LCG_C) % _LCG_M
        out[t] = 2.0 * (prev / float(_LCG_M)) - 1.0
    return out


def perturb_coords_lcg(mesh: HexMesh, amplitude=0.125, seed=None):
    """perturb_coord_hex_8: three LCG draws per locally-owned node, nodes in ascending id,
    seed = rank+1.  Pure-Python loop: meant for the small fixture meshes."""
    seed = mesh.rank + 1 if seed is None else seed
    owned = np.nonzero(mesh.node_owner == mesh.rank)[0]
    owned = owned[np.argsort(mesh.node_gid[owned], kind="stable")]
    r = lcg_stream(seed, 3 * owned.size).reshape(-1, 3)
    mesh.coords[owned] += amplitude * r
    return mesh


def perturb_coords_hash(mesh: HexMesh, amplitude=0.125):
    """Deterministic, rank-independent perturbation keyed on the node id (vectorised): every
    rank that holds a node computes the same displacement, so no owned->shared copy is needed.
    Used for the large synthetic boxes, where the serial LCG stream would take minutes."""
    g = mesh.node_gid.astype(np.uint64)
    out = np.empty((g.size, 3))
    for d in range(3):
        x = (g * np.uint64(3) + np.uint64(d) + np.uint64(0x9E3779B97F4A7C15)) & np.uint64(0xFFFFFFFFFFFFFFFF)
        x ^= x >> np.uint64(30)
        x = (x * np.uint64(0xBF58476D1CE4E5B9)) & np.uint64(0xFFFFFFFFFFFFFFFF)
        x ^= x >> np.uint64(27)
        x = (x * np.uint64(0x94D049BB133111EB)) & np.uint64(0xFFFFFFFFFFFFFFFF)
        x ^= x >> np.uint64(31)
        out[:, d] = (x >> np.uint64(11)).astype(np.float64) * (2.0 / float(1 << 53)) - 1.0
    mesh.coords += amplitude * out
    return mesh


# ---- boundary faces (sidesets) ---------------------------------------------------------------------------------------
# Hex8 side -> element-local node ordinals (src/master_element/Hex8CVFEM.C:414-422; the STK / Exodus HEX_8 side table)
HEX8_SIDE_NODES = np.array([[0, 1, 5, 4], [1, 2, 6, 5], [2, 3, 7, 6], [0, 4, 7, 3], [0, 3, 2, 1], [4, 5, 6, 7]], dtype=np.int32)


def exposed_faces(mesh: HexMesh):
    """(face_elem, face_ordinal) of every element side that no other element of this mesh shares: the skin of a one-rank
    mesh (what `generated:...|sideset:xXyYzZ` or an Exodus side set lists), ordered by (element, ordinal)."""
    en = mesh.elem_nodes
    ne = en.shape[0]
    gid = mesh.node_gid[en[:, HEX8_SIDE_NODES]]          # [ne, 6, 4]
    key = np.sort(gid.reshape(ne * 6, 4), axis=1)
    _, inv, cnt = np.unique(key, axis=0, return_inverse=True, return_counts=True)
    single = np.nonzero(cnt[inv.reshape(-1)] == 1)[0]
    return (single // 6).astype(np.int32), (single % 6).astype(np.int32)


def box_side_faces(mesh: HexMesh, side: str):
    """The faces of sideset `side` (one of x X y Y z Z: low / high end of an axis, Ioss `generated:` naming) of a generated
    box or of this rank's z-slab of it; empty on a slab that does not touch the side."""
    b = mesh.box
    assert b is not None and side in "xXyYzZ"
    nx, ny, lz = b.nx, b.ny, b.nz
    i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(lz), indexing="ij")
    e = (i + nx * (j + ny * k)).astype(np.int32)
    # ordinal of the side facing -x, +x, -y, +y, -z, +z
    sel, ordinal = {"x": (i == 0, 3), "X": (i == nx - 1, 1), "y": (j == 0, 0), "Y": (j == ny - 1, 2),
                    "z": ((k == 0) & (b.k0 == 0), 4), "Z": ((k == lz - 1) & (b.k0 + lz == b.nz_global), 5)}[side]
    fe = np.sort(e[sel])
    return fe, np.full(fe.size, ordinal, dtype=np.int32)


# ---- Hex27 (P2 CVFEM) boxes: BASELINE config 5 ---------------------------------------------------------------------
# tensor-grid node (i, j, k) of a Hex27 element -> STK Hex27 ordinal (HexahedralP2Element::stkNodeMap_,
# src/master_element/Hex27CVFEM.C:36-47)
HEX27_STK_NODE_MAP = np.array([0, 8, 1, 11, 21, 9, 3, 10, 2, 12, 25, 13, 23, 20, 24, 15, 26, 14, 4, 16, 5, 19, 22, 17,
                               7, 18, 6], dtype=np.int32)


def hex27_box(nx, ny, nz) -> HexMesh:
    """nx*ny*nz Hex27 elements on the (2nx+1)(2ny+1)(2nz+1) node lattice of spacing 1/2 (what promoting a `generated:`
    Hex8 box to P2 produces, up to the node numbering: here lexicographic, x fastest), elem_nodes[nElem, 27] in STK order."""
    sx, sy, sz = 2 * nx + 1, 2 * ny + 1, 2 * nz + 1
    zz, yy, xx = np.meshgrid(np.arange(sz), np.arange(sy), np.arange(sx), indexing="ij")
    coords = 0.5 * np.stack([xx.ravel(), yy.ravel(), zz.ravel()], axis=1).astype(np.float64)
    ek, ej, ei = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    n0 = (2 * ei + sx * (2 * ej + sy * 2 * ek)).ravel()
    conn = np.empty((n0.size, 27), dtype=np.int32)
    for c in range(3):
        for b in range(3):
            for a in range(3):
                conn[:, HEX27_STK_NODE_MAP[a + 3 * (b + 3 * c)]] = n0 + a + sx * (b + sy * c)
    n = sx * sy * sz
    return HexMesh(coords=coords, node_gid=np.arange(1, n + 1, dtype=np.int64), node_owner=np.zeros(n, dtype=np.int32),
                   elem_nodes=conn)


def hex27_box_rank(nx, ny, nz, rank, nranks) -> HexMesh:
    """Rank-local z-slab of the nx*ny*nz Hex27 box (`hex27_box` numbering: id = 1 + lexicographic index on the global
    (2nx+1)(2ny+1)(2nz+1) lattice); the node plane shared with the lower rank is owned by the lower rank."""
    k0, k1 = slab_ranges(nz, nranks)[rank]
    lz = k1 - k0
    assert lz > 0, "more ranks than element layers"
    sx, sy = 2 * nx + 1, 2 * ny + 1
    loc = hex27_box(nx, ny, lz)
    loc.coords[:, 2] += k0
    loc.node_gid = loc.node_gid + sx * sy * 2 * k0
    owner = np.full(loc.n_nodes, rank, dtype=np.int32)
    if rank > 0:
        owner[:sx * sy] = rank - 1
    loc.node_owner = owner
    loc.rank, loc.nranks = rank, nranks
    loc.elem_gid = np.arange(1, nx * ny * lz + 1, dtype=np.int64) + nx * ny * k0
    return loc


def perturb_coords_by_gid(mesh: HexMesh, amplitude):
    """deterministic perturbation that depends on the node id only (every rank that holds a node moves it identically)"""
    g = mesh.node_gid.astype(np.float64)
    for d, (a, b) in enumerate(((12.9898, 43758.5453), (78.233, 24634.6345), (39.3468, 11938.1459))):
        h = np.sin(g * a + d) * b
        mesh.coords[:, d] += amplitude * (2.0 * (h - np.floor(h)) - 1.0)
    return mesh
