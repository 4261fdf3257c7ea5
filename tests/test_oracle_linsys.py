"""Pin the oracle's graph / sumInto / loadComplete restatement against the reference's
own linear-system tests: unit_tests/UnitTestTpetra.C (2-element mesh, np=1 and np=2) and
unit_tests/UnitTestLocalGraphArrays.C."""
import numpy as np
import pytest

import oracle
from golden_io import load_golden, tab
from nalu_b200 import mesh


def test_local_graph_arrays_numdof1():
    # UnitTestLocalGraphArrays.C:45-71
    g = oracle.GraphArrays([2, 3, 4])
    assert list(np.diff(g.rowptr())) == [2, 3, 4]
    g.insert(0, [0, 1], 1)
    g.insert(1, [2, 3, 4], 1)
    g.insert(2, [7, 8, 6, 5], 1)
    assert list(g.colidx()) == list(range(9))
    g.insert(0, [0, 1], 1)
    g.insert(1, [2, 3, 4], 1)
    assert list(g.colidx()) == list(range(9))


def test_local_graph_arrays_numdof3():
    # UnitTestLocalGraphArrays.C:73-105
    g = oracle.GraphArrays([6, 9, 12])
    g.insert(0, [0, 3], 3)
    g.insert(1, [6, 9, 12], 3)
    g.insert(2, [21, 18, 24, 15], 3)
    assert list(g.colidx()) == list(range(27))
    g.insert(0, [0, 3], 3)
    g.insert(1, [6, 9, 12], 3)
    assert list(g.colidx()) == list(range(27))


def _two_hex_system(nprocs):
    full = mesh.generated_box(1, 1, 2)
    # UnitTestTpetra.C:97-98: elem 1 nodes 1 2 4 3 5 6 8 7 ; elem 2 nodes 5 6 8 7 9 10 12 11
    assert list(full.node_gid[full.elem_nodes[0]]) == [1, 2, 4, 3, 5, 6, 8, 7]
    assert list(full.node_gid[full.elem_nodes[1]]) == [5, 6, 8, 7, 9, 10, 12, 11]
    parts = [full] if nprocs == 1 else mesh.partition_slabs(full, nprocs)
    return oracle.LinSys(parts, 1), parts


@pytest.mark.parametrize("nprocs", [1, 2])
def test_tpetra_basic_graph_and_matrix(nprocs):
    g = load_golden("tpetra_2elem")
    elemVals, lhsVals = tab(g, "elemVals"), tab(g, "lhsVals")
    ls, parts = _two_hex_system(nprocs)
    # verify_graph_for_2_hex8_mesh (UnitTestTpetra.C:146-176)
    gold_len = {1: [[8, 8, 8, 8, 12, 12, 12, 12, 8, 8, 8, 8]],
                2: [[8, 8, 8, 8, 12, 12, 12, 12], [8, 8, 8, 8]]}[nprocs]
    assert sum(ls.n_owned_rows(r) for r in range(nprocs)) == 12
    for r in range(nprocs):
        assert ls.n_owned_rows(r) == len(gold_len[r])
        assert list(np.diff(ls.rowptr(r)).astype(int)) == gold_len[r]
    # execute() + loadComplete(), then verify_matrix_for_2_hex8_mesh (:178-203)
    keep = []
    d = oracle.make_desc("test_fixed_block", keep, fixedBlock=elemVals)
    for r in range(nprocs):
        ls.assemble_elem(r, [d])
    ls.load_complete()
    for r in range(nprocs):
        rp, ci, v = ls.rowptr(r), ls.colidx(r), ls.values(r)
        rg, cg = ls.row_gids(r), ls.col_gids(r)
        for row in range(ls.n_owned_rows(r)):
            for j in range(int(rp[row]), int(rp[row + 1])):
                assert abs(lhsVals[rg[row] - 1][cg[ci[j]] - 1] - v[j]) < 1e-9


def test_np2_ownership_and_colmap_order():
    ls, parts = _two_hex_system(2)
    # proc 0 owns 8 of 12 rows (lowest sharing rank owns shared nodes)
    assert list(ls.row_gids(0)) == list(range(1, 9))
    assert list(ls.row_gids(1)) == [9, 10, 11, 12]
    assert list(ls.row_gids(1, shared=True)) == [5, 6, 7, 8]
    # column map: owned, then shared-not-owned, then remote ordered by (owner, gid)
    assert list(ls.col_gids(1)) == [9, 10, 11, 12, 5, 6, 7, 8]
    assert list(ls.col_gids(0)) == list(range(1, 9)) + [9, 10, 11, 12]
    # proc 0's rows for the shared nodes include the remote columns 9..12
    rp, ci, cg = ls.rowptr(0), ls.colidx(0), ls.col_gids(0)
    for row in range(4, 8):
        cols = sorted(cg[ci[int(rp[row]):int(rp[row + 1])]])
        assert cols == list(range(1, 13))


@pytest.mark.parametrize("ndof", [1, 3])
def test_np_invariance_of_assembled_system(ndof):
    """np=1 vs np=3 slabs on a 3x2x5 perturbed box: after loadComplete the union of owned rows
    is the same matrix (by global ids)."""
    full = mesh.generated_box(3, 2, 5)
    mesh.perturb_coords_hash(full, 0.1)
    rng = np.random.default_rng(0)
    n = 8 * ndof
    blk = rng.standard_normal((n, n))
    keep = []
    d = oracle.make_desc("test_fixed_block", keep, fixedBlock=blk)

    def dense(parts):
        ls = oracle.LinSys(parts, ndof)
        for r in range(len(parts)):
            ls.assemble_elem(r, [d])
        ls.load_complete()
        N = full.n_nodes * ndof
        A = np.zeros((N, N))
        for r in range(len(parts)):
            rp, ci, v, rg, cg = ls.rowptr(r), ls.colidx(r), ls.values(r), ls.row_gids(r), ls.col_gids(r)
            for row in range(ls.n_owned_rows(r)):
                for j in range(int(rp[row]), int(rp[row + 1])):
                    A[rg[row] - 1, cg[ci[j]] - 1] += v[j]
        return A

    # test_fixed_block overwrites (lhs = block) per element, sumInto adds: same on any np
    A1 = dense([full])
    A3 = dense(mesh.partition_slabs(full, 3))
    np.testing.assert_allclose(A3, A1, rtol=0, atol=1e-12)
