#!/usr/bin/env python3
"""BASELINE config 2: scalar advection-diffusion (ScalarAdvDiffElemKernel, mixture-fraction fixture fields) on a fully
PERIODIC 64^3 Hex8 box: 262 144 elements, 262 144 unique node ids (subjects alias their monarchs in x, y, z), one GPU,
device-resident inputs, CUDA events.  The periodic aliasing breaks the lattice numbering the tiled kernels rely on, so the
system is assembled by the coloured general kernel (deterministic scatter); the same box without periodicity is timed
through the tiled write-once kernel beside it.  Every entry of the periodic system is compared with the oracle.
usage: bench_config2.py [n=64] [steps=10]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402  (checker only)
from nalu_b200 import fields, mesh  # noqa: E402
from nalu_b200.linsys import DeviceCsrLinearSystem  # noqa: E402
from parity_util import rel_err, rel_err_rows  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda", 0)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
out = {"workload": f"config2_periodic_box_{n}^3_scalar_adv_diff", "elements": n ** 3, "kernels": ["ScalarAdvDiffElemKernel"]}


def state(m, monarch=None):
    c = m.coords
    q, gam = fields.mixture_fraction(c, 1.0, 0.37), fields.smooth_viscosity(c)
    rho, vel = fields.smooth_density(c), fields.velocity(c)
    if monarch is not None:  # subjects carry their monarch's values; coordinates stay their own
        q, gam, rho, vel = q[monarch], gam[monarch], rho[monarch], vel[monarch]
    return c, q, gam, rho, vel


def run(m, monarch, label):
    c, q, gam, rho, vel = state(m, monarch)
    ls = DeviceCsrLinearSystem(m, 1).allocate(dev)
    mdot = ls.mdot_scs(t(c), t(rho), t(vel))
    d = ls.make_desc("scalar_adv_diff", fields=dict(coordinates=t(c), scalarq=t(q), viscosity=t(gam)), mass_flow_rate=mdot)
    fam = ls.assemble_elem([d], mode="auto", overwrite=True)
    ow = fam == 2

    def step():
        if not ow:
            ls.zero_system()
        ls.assemble_elem([d], mode="auto", overwrite=ow)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        a.record()
        step()
        b.record()
    torch.cuda.synchronize()
    ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    nn = ls.n_owned_rows
    bytes_alg = 8 * 5 * nn + (96 + 32) * m.n_elems + 8 * int(ls.values.numel()) + 8 * nn
    r = {"kernel_family": fam, "launch": "tiled write-once" if ow else "general, coloured scatter (+ zeroSystem)", "ms": ms,
         "Melem_per_s": m.n_elems / ms / 1e3, "rows": nn, "bytes_per_element": bytes_alg / m.n_elems,
         "roofline_frac": bytes_alg / (ms * 1e-3) / 1e9 / peak}
    if label == "periodic":
        orc = oracle.LinSys([m], 1)
        keep = []
        orc.assemble_elem(0, [oracle.make_desc("scalar_adv_diff", keep, coordinates=c, scalarQ=q, viscosity=gam,
                                               massFlowRate=mdot.cpu().numpy())])
        assert np.array_equal(ls.rowptr(), orc.rowptr(0)) and np.array_equal(ls.colidx(), orc.colidx(0))
        ov, orr = orc.values(0), orc.rhs(0)
        r["vs_oracle"] = {"graph": "bit-exact", "max_rel_err_values": rel_err_rows(ls.owned_values().cpu().numpy(), ov, orc.rowptr(0), 1e-2),
                          "max_rel_err_rhs": rel_err(ls.owned_rhs().cpu().numpy(), orr, 1e-2 * np.abs(orr).max()), "entries": int(ov.size)}
    out[label] = r


mp = mesh.periodic_box(n, n, n)
mesh.perturb_coords_hash(mp, 0.1)
run(mp, mp.periodic_monarch, "periodic")
mb = mesh.generated_box(n, n, n)
mesh.perturb_coords_hash(mb, 0.1)
run(mb, None, "non_periodic_same_size")
print(json.dumps(out))
