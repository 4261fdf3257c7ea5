#!/usr/bin/env python3
"""The momentum kernel list of the low-Mach equation system, [MomentumMassElemKernel(lumped), MomentumAdvDiffElemKernel]
(BASELINE config 3), on an n^3 box: tiled + general split (auto) vs the general kernel alone.  usage: [n] [steps]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nalu_b200 import fields, mesh  # noqa: E402
from nalu_b200.linsys import DeviceCsrLinearSystem  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
if len(sys.argv) > 3:  # A/B of the general kernel's coloured scatter: 0 = load/add/store, 1 = red.global.add
    from nalu_b200 import lib as _lib
    _lib.tune(general_scatter=int(sys.argv[3]))
dev = torch.device("cuda", 0)
m = mesh.generated_box(n, n, n)
mesh.perturb_coords_hash(m, 0.12)
if len(sys.argv) > 7:  # 1 row-owner form (one launch), 0 element-colour form of the unstructured kernel
    from nalu_b200 import lib as _lib
    _lib.tune(urow_owner=int(sys.argv[7]))
if len(sys.argv) > 6:  # CTAs per SM the row-thread unstructured kernel is register-bounded for
    from nalu_b200 import lib as _lib
    _lib.tune(urow_minb=int(sys.argv[6]))
if len(sys.argv) > 5:  # elements per (chunk, colour) launch group
    from nalu_b200 import lib as _lib
    _lib.tune(color_chunk=int(sys.argv[5]))
if len(sys.argv) > 4 and sys.argv[4] == "nobox":
    # the same elements declared as a general Hex8 mesh: "auto" is then the row-thread unstructured kernel (mode 4)
    m.box = None
c = m.coords
t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
vel, rho = fields.velocity(c), fields.smooth_density(c)
f = dict(coordinates=t(c), velocity=t(vel), velocity_n=t(0.9 * vel), velocity_nm1=t(0.8 * vel), viscosity=t(fields.smooth_viscosity(c)),
         density=t(rho), density_n=t(0.98 * rho), density_nm1=t(0.95 * rho), dpdx=t(fields.dpdx(c)))
ls = DeviceCsrLinearSystem(m, 3).allocate(dev)
mdot = ls.mdot_scs(f["coordinates"], f["density"], f["velocity"])
descs = [ls.make_desc("momentum_mass", fields=f, options=("lumped_mass",), dt=0.05, gamma1=1.5, gamma2=-2.0, gamma3=0.5),
         ls.make_desc("momentum_adv_diff", fields=f, mass_flow_rate=mdot, include_divu=1.0)]
out = {"n": n, "elements": m.n_elems, "kernels": ["MomentumMassElemKernel(lumped)", "MomentumAdvDiffElemKernel(includeDivU=1)"]}
ref = None
for mode in ("general", "auto"):
    fam = ls.assemble_elem(descs, mode=mode, overwrite=(mode == "auto"))

    def step(fam=fam):
        if fam != 2:  # only the tiled write-once kernels need no zeroSystem
            ls.zero_system()
        return ls.assemble_elem(descs, mode=mode, overwrite=(fam == 2))
    for _ in range(2):
        fam = step()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        a.record()
        step()
        b.record()
    torch.cuda.synchronize()
    ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    v, r = ls.owned_values(), ls.owned_rhs()
    if ref is None:
        ref = (float(v.abs().max()), v.clone(), r.clone())
    else:
        out["split_vs_general_values_relerr"] = float((v - ref[1]).abs().max() / ref[0])
        out["split_vs_general_rhs_relerr"] = float((r - ref[2]).abs().max() / ref[2].abs().max())
    out[mode] = {"kernel_family": fam, "ms": ms, "Melem_per_s": m.n_elems / ms / 1e3}
print(json.dumps(out))
