"""The product's CSR graph builder (nalu_b200/csrc/graph.cpp, through the C ABI) must equal the
oracle's literal restatement of TpetraLinearSystem's graph bit-for-bit: rowptr, colidx, row
maps, column map -- np=1 and simulated multi-rank.  Runs without a GPU (host-only finalize)."""
import numpy as np
import pytest

import oracle
from nalu_b200 import mesh
from nalu_b200.linsys import DeviceCsrLinearSystem


def build_product_systems(parts, ndof):
    """Create all ranks in one process and route the handshake messages between them."""
    # phase 1: create every rank with a deferred exchange (collect sends first)
    sends = {}
    pending = []

    class Deferred(Exception):
        pass

    # two-pass: first gather all send messages, then build for real with a lookup exchange
    for m in parts:
        def capture(send, r=m.rank):
            sends[r] = send
            raise Deferred()
        try:
            DeviceCsrLinearSystem(m, ndof, exchange=capture, host_only=True)
        except Deferred:
            pass
        if m.nranks == 1:
            sends[m.rank] = {}
    out = []
    for m in parts:
        def exchange(send, r=m.rank):
            return {src: sends[src][r] for src in sends if r in sends[src]}
        out.append(DeviceCsrLinearSystem(m, ndof, exchange=exchange, host_only=True))
    return out


def assert_graph_equal(prod, orc, nranks):
    for r in range(nranks):
        p = prod[r]
        assert p.n_owned_rows == orc.n_owned_rows(r)
        assert p.n_shared_rows == orc.n_shared_rows(r)
        assert p.info.n_cols == orc.n_cols(r)
        for shared in (False, True):
            np.testing.assert_array_equal(p.rowptr(shared), orc.rowptr(r, shared))
            np.testing.assert_array_equal(p.colidx(shared), orc.colidx(r, shared))
            np.testing.assert_array_equal(p.row_gids(shared), orc.row_gids(r, shared))
        np.testing.assert_array_equal(p.col_gids(), orc.col_gids(r))
        rl, cl = p.node_lids()
        np.testing.assert_array_equal(rl, orc.node_row_lid(r))
        np.testing.assert_array_equal(cl, orc.node_col_lid(r))


@pytest.mark.parametrize("ndof", [1, 3])
@pytest.mark.parametrize("dims,nranks", [((1, 1, 2), 1), ((1, 1, 2), 2), ((3, 2, 5), 1), ((3, 2, 5), 3),
                                         ((4, 4, 4), 4), ((6, 5, 7), 2)])
def test_box_graph_bit_exact(dims, nranks, ndof):
    full = mesh.generated_box(*dims)
    parts = [full] if nranks == 1 else mesh.partition_slabs(full, nranks)
    orc = oracle.LinSys(parts, ndof)
    prod = build_product_systems(parts, ndof)
    assert_graph_equal(prod, orc, nranks)


@pytest.mark.parametrize("ndof", [1, 3])
def test_unstructured_partition_graph_bit_exact(ndof):
    """Random element->rank assignment: irregular interfaces, several owners per rank, nodes
    shared by up to 4 ranks, remote columns from several owners."""
    full = mesh.generated_box(5, 4, 3)
    rng = np.random.default_rng(3)
    # shuffle node ids so that local order != id order
    perm = rng.permutation(full.n_nodes)
    full.node_gid = (perm + 1).astype(np.int64)
    nranks = 4
    elem_rank = rng.integers(0, nranks, full.n_elems)
    parts = mesh.partition_elements(full, elem_rank, nranks)
    orc = oracle.LinSys(parts, ndof)
    prod = build_product_systems(parts, ndof)
    assert_graph_equal(prod, orc, nranks)


def test_periodic_box_graph_np1():
    full = mesh.periodic_box(4, 3, 3)
    for ndof in (1, 3):
        orc = oracle.LinSys([full], ndof)
        prod = build_product_systems([full], ndof)
        assert_graph_equal(prod, orc, 1)
        # every row of a fully periodic box with >=3 cells per direction has 27 node blocks
        assert set(np.diff(prod[0].rowptr()).astype(int)) == {27 * ndof}


@pytest.mark.parametrize("kind", ["box", "shuffled", "periodic", "partition"])
def test_element_colouring_is_conflict_free(kind):
    """Elements of one colour share no ROW node (periodic aliases count as one row): the general kernel sums a colour
    into the CSR without atomics."""
    rng = np.random.default_rng(5)
    if kind == "periodic":
        meshes = [mesh.periodic_box(5, 4, 3)]
    else:
        full = mesh.generated_box(6, 5, 4)
        if kind == "shuffled":
            full.node_gid = (rng.permutation(full.n_nodes) + 1).astype(np.int64)
            full.box = None
        meshes = mesh.partition_elements(full, rng.integers(0, 3, full.n_elems), 3) if kind == "partition" else [full]
    for p, m in zip(build_product_systems(meshes, 3), meshes):
        ncol, col = p.elem_colors()
        assert ncol >= 1
        assert col.min() >= 0
        rows, _ = p.node_lids()
        er = rows[m.elem_nodes]                      # [nElem, 8] row of every element node
        for c in range(ncol):
            r = er[col == c].ravel()
            assert np.unique(r).size == r.size, f"colour {c} touches a row twice"


def test_tiny_periodic_box_has_no_colouring():
    # 1 cell in x: an element reaches the same row through two of its nodes -> atomic fallback
    p = build_product_systems([mesh.periodic_box(1, 3, 3)], 1)[0]
    assert p.elem_colors()[0] == 0


def test_row_lengths_of_a_box():
    # interior 27, face 18, edge 12, corner 8 node blocks (UnitTestTpetra.C:146-158 for 1x1x2)
    p = build_product_systems([mesh.generated_box(4, 4, 4)], 3)[0]
    lens = np.diff(p.rowptr()).astype(int) // 3
    vals, counts = np.unique(lens, return_counts=True)
    assert dict(zip(vals, counts // 3)) == {8: 8, 12: 36, 18: 54, 27: 27}
    assert p.nnz_owned == 9 * (3 * 4 + 1) ** 3


def test_bad_arguments_are_reported():
    from nalu_b200.lib import NaluB200Error
    m = mesh.generated_box(2, 2, 2)
    with pytest.raises(NaluB200Error):
        DeviceCsrLinearSystem(m, 2, host_only=True)  # numDof must be 1 or 3
    m.elem_nodes = m.elem_nodes.copy()
    m.elem_nodes[0, 0] = 10 ** 6
    with pytest.raises(NaluB200Error):
        DeviceCsrLinearSystem(m, 1, host_only=True)
