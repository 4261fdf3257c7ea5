"""Exodus reader (nalu_b200/exodus.py) and the graph builder on REAL meshes of the reference's regression suite
(tests/golden/meshes/, made by tests/golden/make_mesh_fixtures.py): an unstructured one-block Hex8 mesh with 2 to 46
elements per node (thex8_8.g) and a mesh decomposed into four pieces by the SEACAS tool chain (ether_conform_coarse_mks.g.4.*).
No GPU needed (host-only graph finalize)."""
import os

import numpy as np
import pytest

import oracle
from nalu_b200 import exodus, mesh
from test_graph import assert_graph_equal, build_product_systems

MESHES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "meshes")
REF = "/root/reference/reg_tests/mesh"


def thex8():
    return exodus.read_exodus_hex8(os.path.join(MESHES, "thex8_8.g"))


def ether_pieces():
    return exodus.read_exodus_pieces(os.path.join(MESHES, "ether_conform_coarse_mks.g"), 4)


def test_reader_counts_maps_and_orientation():
    m = thex8()
    assert (m.n_nodes, m.n_elems) == (2365, 1852)
    assert m.elem_nodes.dtype == np.int32 and m.elem_nodes.min() == 0 and m.elem_nodes.max() == m.n_nodes - 1
    assert np.unique(m.node_gid).size == m.n_nodes and m.node_gid.min() >= 1
    # every element is a right-handed Hex8 in Exodus node order: all 12 SCS Jacobians positive
    for e in range(m.n_elems):
        assert oracle.scs_grad_op(m.coords[m.elem_nodes[e]])[2] == 0
    # total volume of the sub-control volumes = volume of the meshed unit-ish domain, the same from both formulas
    vol = sum(oracle.scv_volume(m.coords[m.elem_nodes[e]]).sum() for e in range(m.n_elems))
    grandy = sum(oracle.hex_volume_grandy(m.coords[m.elem_nodes[e]]) for e in range(m.n_elems))
    assert abs(vol - grandy) < 1e-12 * abs(grandy)


def test_write_read_round_trip(tmp_path):
    m = mesh.generated_box(3, 2, 2)
    mesh.perturb_coords_lcg(m)
    m.node_gid = (np.random.default_rng(0).permutation(m.n_nodes) + 1).astype(np.int64)
    p = str(tmp_path / "box.g")
    exodus.write_minimal_exodus(p, m)
    r = exodus.read_exodus_hex8(p)
    np.testing.assert_array_equal(r.coords, m.coords)
    np.testing.assert_array_equal(r.elem_nodes, m.elem_nodes)
    np.testing.assert_array_equal(r.node_gid, m.node_gid)


def test_rejects_what_it_cannot_read(tmp_path):
    p = tmp_path / "garbage.g"
    p.write_bytes(b"\x89HDF\r\n\x1a\n" + b"\0" * 64)  # a netCDF-4 / HDF5 signature
    with pytest.raises(ValueError, match="netCDF-3"):
        exodus.read_exodus_hex8(str(p))


@pytest.mark.parametrize("ndof", [1, 3])
def test_unstructured_mesh_graph_bit_exact(ndof):
    m = thex8()
    orc = oracle.LinSys([m], ndof)
    prod = build_product_systems([m], ndof)
    assert_graph_equal(prod, orc, 1)
    lens = np.diff(prod[0].rowptr()).astype(int) // ndof
    assert lens.min() == 12 and lens.max() == 141  # node blocks per row (a lattice has 8..27): 2 to 46 elements per node
    ncol, col = prod[0].elem_colors()
    assert ncol >= 46 and col.min() >= 0           # at least one colour per element of the busiest node
    rows, _ = prod[0].node_lids()
    er = rows[m.elem_nodes]
    for c in range(ncol):
        r = er[col == c].ravel()
        assert np.unique(r).size == r.size
    # a random 3-way element partition of the same mesh: shared nodes with several owners
    parts = mesh.partition_elements(m, np.random.default_rng(1).integers(0, 3, m.n_elems), 3)
    assert_graph_equal(build_product_systems(parts, ndof), oracle.LinSys(parts, ndof), 3)


@pytest.mark.parametrize("ndof", [1, 3])
def test_decomposed_mesh_graph_bit_exact(ndof):
    parts = ether_pieces()
    assert [m.n_elems for m in parts] == [299, 299, 299, 299]
    allg = np.unique(np.concatenate([m.node_gid for m in parts]))
    assert allg.size == 1505
    for m in parts:  # ownership: lowest rank that holds the node
        for q in parts[:m.rank]:
            assert not np.isin(m.node_gid[m.node_owner == m.rank], q.node_gid).any()
    assert_graph_equal(build_product_systems(parts, ndof), oracle.LinSys(parts, ndof), 4)


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference's mesh directory exists in the build container only")
def test_fixtures_match_the_reference_files_and_reader_handles_the_suite():
    for name in ["thex8_8.g"] + [f"ether_conform_coarse_mks.g.4.{r}" for r in range(4)]:
        a = exodus.read_exodus_hex8(os.path.join(REF, name))
        b = exodus.read_exodus_hex8(os.path.join(MESHES, name))
        np.testing.assert_array_equal(a.coords, b.coords)
        np.testing.assert_array_equal(a.elem_nodes, b.elem_nodes)
        np.testing.assert_array_equal(a.node_gid, b.node_gid)
    # other Hex8 decks of the regression suite: structured, unstructured, multi-block
    for name, nn, ne in (("periodic3d.g", 9261, 8000), ("cube_32.g", 35937, 32768), ("1cm_ped_35K.g", 35532, 34215),
                         ("hex8_20_10_1.g", 60903, 40000)):
        m = exodus.read_exodus_hex8(os.path.join(REF, name))
        assert (m.n_nodes, m.n_elems) == (nn, ne)
    with pytest.raises(ValueError, match="num_dim"):
        exodus.read_exodus_hex8(os.path.join(REF, "quad4_8el.g"))
