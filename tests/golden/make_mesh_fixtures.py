"""Real (unstructured / decomposed) Hex8 meshes of the reference's regression suite as small test fixtures.

Run in the build container, where /root/reference exists:  python tests/golden/make_mesh_fixtures.py
Reads reg_tests/mesh/{thex8_8.g, ether_conform_coarse_mks.g.4.[0-3]} with nalu_b200.exodus and re-writes the variables
the assembly path consumes (coordinates, Hex8 connectivity, node / element id maps) as minimal netCDF-3 Exodus files under
tests/golden/meshes/ -- the GPU box has no /root/reference.  Side sets, names, QA records are dropped.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from nalu_b200 import exodus  # noqa: E402

SRC = "/root/reference/reg_tests/mesh"
OUT = os.path.join(HERE, "meshes")
FILES = ["thex8_8.g"] + [f"ether_conform_coarse_mks.g.4.{r}" for r in range(4)]

if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for name in FILES:
        m = exodus.read_exodus_hex8(os.path.join(SRC, name))
        exodus.write_minimal_exodus(os.path.join(OUT, name), m, title=f"from reg_tests/mesh/{name}")
        print(name, m.n_nodes, "nodes", m.n_elems, "elements")
