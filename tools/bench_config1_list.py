#!/usr/bin/env python3
"""The momentum kernel list of BASELINE config 1 (heatedWaterChannelElem.i:152-155) in 3-D,
[MomentumMassElemKernel(lumped), MomentumUpwAdvDiffElemKernel(limiter), MomentumBuoyancySrcElemKernel], on an n^3 Hex8 box:
the row-owner kernel (auto: no tiled kernel takes this list) vs the general kernel.  usage: [n=128] [steps=5]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nalu_b200 import fields, mesh  # noqa: E402
from nalu_b200.linsys import DeviceCsrLinearSystem  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda", 0)
m = mesh.generated_box(n, n, n)
mesh.perturb_coords_hash(m, 0.12)
c = m.coords
t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
vel, rho = fields.velocity(c), fields.smooth_density(c)
a = 0.3 * np.pi
gju = np.stack([np.sin(a * c[:, j % 3] + 0.2 * j) * (0.3 + 0.02 * j) for j in range(9)], axis=1)
f = dict(coordinates=t(c), velocity=t(vel), velocity_n=t(0.9 * vel), velocity_nm1=t(0.8 * vel), viscosity=t(fields.smooth_viscosity(c)),
         density=t(rho), density_n=t(0.98 * rho), density_nm1=t(0.95 * rho), dpdx=t(fields.dpdx(c)), gju=t(gju))
ls = DeviceCsrLinearSystem(m, 3).allocate(dev)
mdot = ls.mdot_scs(f["coordinates"], f["density"], f["velocity"])
descs = [ls.make_desc("momentum_mass", fields=f, options=("lumped_mass",), dt=0.05, gamma1=1.5, gamma2=-2.0, gamma3=0.5),
         ls.make_desc("momentum_upw_adv_diff", fields=f, options=("use_limiter",), mass_flow_rate=mdot, include_divu=1.0,
                      upw=dict(alpha=0.0, alpha_upw=1.0, ho_upwind=1.0, hybrid_factor=1.0)),
         ls.make_desc("momentum_buoyancy_src", fields=dict(coordinates=f["coordinates"], density=f["density"]), gravity=(0.0, 0.0, -9.81),
                      rho_ref=1.0)]
out = {"n": n, "elements": m.n_elems, "kernels": ["MomentumMassElemKernel(lumped)", "MomentumUpwAdvDiffElemKernel(limiter)",
                                                 "MomentumBuoyancySrcElemKernel"]}
ref = None
for mode in ("general", "auto"):
    def step():
        ls.zero_system()
        return ls.assemble_elem(descs, mode=mode)
    for _ in range(2):
        fam = step()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for x, y in ev:
        x.record()
        step()
        y.record()
    torch.cuda.synchronize()
    ms = float(np.mean([x.elapsed_time(y) for x, y in ev]))
    v, r = ls.owned_values(), ls.owned_rhs()
    if ref is None:
        ref = (v.clone(), r.clone())
    else:
        out["auto_vs_general_values_relerr"] = float((v - ref[0]).abs().max() / ref[0].abs().max())
        out["auto_vs_general_rhs_relerr"] = float((r - ref[1]).abs().max() / ref[1].abs().max())
    out[mode] = {"kernel_family": fam, "ms_incl_zero_system": ms, "Melem_per_s": m.n_elems / ms / 1e3}
print(json.dumps(out))
