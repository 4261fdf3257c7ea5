#!/usr/bin/env python3
"""One-dof assembly (ContinuityAdvElemKernel or ScalarAdvDiffElemKernel) on an n^3 box: tiled write-once kernel vs
the general (atomic) kernel, device-resident inputs, CUDA events.
usage: bench_continuity.py [n] [steps] [continuity_adv|scalar_adv_diff]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nalu_b200 import fields, mesh  # noqa: E402
from nalu_b200.linsys import DeviceCsrLinearSystem  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
kind = sys.argv[3] if len(sys.argv) > 3 else "continuity_adv"
dev = torch.device("cuda", 0)
m = mesh.generated_box(n, n, n)
mesh.perturb_coords_hash(m, 0.12)
if len(sys.argv) > 4 and sys.argv[4] == "nobox":
    m.box = None  # the same elements as a general Hex8 mesh: "auto" is the row-owner kernel (mode 4)
c = m.coords
t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
f = dict(coordinates=t(c), velocity=t(fields.velocity(c)), density=t(fields.smooth_density(c)),
         pressure=t(fields.pressure(c)), dpdx=t(fields.dpdx(c)))
ls = DeviceCsrLinearSystem(m, 1).allocate(dev)
if kind == "continuity_adv":
    desc = ls.make_desc("continuity_adv", fields=f, dt=0.05, gamma1=1.5)
    nodal_comps, elem_comps = 11, 0
else:
    f = dict(coordinates=f["coordinates"], scalarq=t(fields.mixture_fraction(c, 1.0, 0.37)), viscosity=t(fields.smooth_viscosity(c)))
    mdot = DeviceCsrLinearSystem.mdot_scs(ls, f["coordinates"], t(fields.smooth_density(c)), t(fields.velocity(c)))
    desc = ls.make_desc("scalar_adv_diff", fields=f, mass_flow_rate=mdot)
    nodal_comps, elem_comps = 5, 12
out = {"n": n, "elements": m.n_elems, "kind": kind}
ref = None
for mode in ("general", "auto"):
    ow = ls.assemble_elem([desc], mode=mode, overwrite=True) == 2  # only the tiled write-once kernels need no zeroSystem

    def step(ow=ow):
        if not ow:
            ls.zero_system()
        return ls.assemble_elem([desc], mode=mode, overwrite=ow)

    for _ in range(3):
        fam = step()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        a.record()
        step()
        b.record()
    torch.cuda.synchronize()
    ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    v, r = ls.owned_values().clone(), ls.owned_rhs().clone()
    if ref is None:
        ref = (v, r)
    else:
        out["tiled_vs_general_values_relerr"] = float((v - ref[0]).abs().max() / ref[0].abs().max())
        out["tiled_vs_general_rhs_relerr"] = float((r - ref[1]).abs().max() / ref[1].abs().max())
    nnz = ls.nnz_owned
    bytes_alg = 8 * nodal_comps * m.n_nodes + (32 + 8 * elem_comps) * m.n_elems + 8 * nnz + 8 * m.n_nodes
    out[mode] = {"kernel_family": fam, "ms": ms, "Melem_per_s": m.n_elems / ms / 1e3, "GBps_algorithmic": bytes_alg / ms / 1e6,
                 "bytes_per_element": bytes_alg / m.n_elems}
out["geometry_errors"] = ls.geometry_errors()
print(json.dumps(out))
