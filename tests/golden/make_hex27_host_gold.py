"""Gold tables for the Hex27 case of tests/host_cpp/unit_tests.cpp (test_hex27_momentum_list): the element matrix and residual of
[MomentumMassElemKernel(lumped), MomentumAdvDiffElemKernel(includeDivU=1)] on one curved Hex27 element, from the numpy oracle
(oracle/hex27.py).  The reference holds no Hex27 kernel gold; mesh and fields are closed-form functions that the C++ test
re-evaluates.  Run from the repo root:  python tests/golden/make_hex27_host_gold.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from nalu_b200 import mesh as M  # noqa: E402
from oracle import hex27 as H  # noqa: E402


def state(c0):
    X0, Y0, Z0 = c0[:, 0], c0[:, 1], c0[:, 2]
    c = c0 + 0.04 * np.stack([np.sin(3.1 * Y0 + 1.7 * Z0 + 0.3), np.cos(2.3 * X0 - 1.1 * Z0), np.sin(1.9 * X0 + 2.9 * Y0 + 0.5)], axis=1)
    X, Y, Z = c[:, 0], c[:, 1], c[:, 2]
    u = np.stack([np.sin(0.9 * Y) + 0.3 * Z, np.cos(0.7 * X) * Z, 0.2 * X * Y - 0.5], axis=1)
    mu = 0.3 + 0.1 * np.cos(X + 2.0 * Y)
    rho = 1.0 + 0.2 * np.sin(Z)
    gp = np.stack([0.1 * X, -0.2 * Y, 0.05 * Z], axis=1)
    return c, u, 0.9 * u + 0.01, 0.8 * u - 0.02, mu, rho, 0.98 * rho, 0.95 * rho, gp


def main():
    m = M.hex27_box(1, 1, 1)
    c, u, un, um, mu, rho, rn, rm, gp = state(m.coords)
    mdot = H.calc_mass_flow_rate_scs(m.elem_nodes, c, rho, u)
    l1, r1 = H.momentum_adv_diff(m.elem_nodes, c, u, mu, mdot, include_divu=1.0)
    l2, r2 = H.momentum_mass(m.elem_nodes, c, u, un, um, rho, rn, rm, gp, dt=0.05, gamma1=1.5, gamma2=-2.0, gamma3=0.5, lumped=True)
    # element ordering of the gold = (STK node ordinal, dof)
    lhs, rhs = (l1 + l2)[0], (r1 + r2)[0]
    X, Y, Z = c[:, 0], c[:, 1], c[:, 2]
    pr = 0.5 * np.cos(0.8 * X) * np.sin(0.6 * Y) + 0.1 * Z
    lc, rc = H.continuity_adv(m.elem_nodes, c, u, rho, pr, gp, dt=0.05, gamma1=1.5)
    out = {"source": "oracle/hex27.py via tests/golden/make_hex27_host_gold.py (no reference gold exists for Hex27 kernels)",
           "tables": {"momentum_list::lhs": {"dims": [81, 81], "values": [float(x) for x in lhs.ravel()]},
                      "momentum_list::rhs": {"dims": [81], "values": [float(x) for x in rhs]},
                      "mdot": {"dims": [216], "values": [float(x) for x in mdot[0]]},
                      "continuity::lhs": {"dims": [27, 27], "values": [float(x) for x in lc[0].ravel()]},
                      "continuity::rhs": {"dims": [27], "values": [float(x) for x in rc[0]]}}}
    with open(os.path.join(ROOT, "tests", "golden", "hex27_host.json"), "w") as fh:
        json.dump(out, fh)
    print("wrote hex27_host.json", np.abs(lhs).max(), np.abs(rhs).max())


if __name__ == "__main__":
    main()
