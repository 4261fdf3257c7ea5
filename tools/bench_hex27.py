"""BASELINE config 5: N^3 Hex27 (P2 CVFEM) box, MomentumAdvDiffElemKernel<AlgTraitsHex27> + sumInto on one B200.

    python tools/bench_hex27.py [N=48] [steps=10] [warmup=3]

A step = zeroSystem + one nalu_b200_assemble_elem (8 colour launches).  Inputs resident in HBM, CUDA events on the launching
stream, working set (4.2 GB of CSR values at N = 48) far larger than L2.  After the timed region the complete rows of a 3^3
corner sub-box and of one in the middle are compared with the oracle (oracle/hex27.py), and the oracle is timed on a bounded
sample (cpu_baseline, numpy, one core).  Prints one JSON line.
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def fields(coords):
    x = coords
    vel = np.stack([np.sin(0.9 * x[:, 1]) + 0.3 * np.cos(x[:, 2]), np.cos(0.7 * x[:, 0]) * np.sin(0.5 * x[:, 2]),
                    0.2 * np.sin(x[:, 0]) * np.cos(x[:, 1]) - 0.5], axis=1)
    visc = 0.3 + 0.1 * np.cos(x[:, 0] + 2.0 * x[:, 1])
    rho = 1.0 + 0.2 * np.sin(x[:, 2])
    return vel, visc, rho


def subbox_check(m, n, i0, vel, visc, rho, got_v, got_b, grp, gci):
    """complete rows of the 3^3-element sub-box whose first element is (i0, i0, i0): nodes strictly inside its node range
    (or on the domain boundary) have all their elements in the sub-box"""
    from oracle import hex27 as H
    sx = 2 * n + 1
    ei = np.arange(i0, i0 + 3)
    elems = (ei[None, None, :] + n * (ei[None, :, None] + n * ei[:, None, None])).ravel()
    conn = m.elem_nodes[elems]
    nodes = np.unique(conn)
    local = np.full(m.n_nodes, -1, dtype=np.int64)
    local[nodes] = np.arange(nodes.size)
    lconn = local[conn].astype(np.int32)
    c = m.coords[nodes]
    c = c - c[0]  # the assembled system is translation invariant; the reference's rounding error is not
    mdot = H.calc_mass_flow_rate_scs(lconn, c, rho[nodes], vel[nodes])
    lhs, rhs = H.momentum_adv_diff(lconn, c, vel[nodes], visc[nodes], mdot, include_divu=1.0)
    rp, ci, vals, b = H.assemble_csr(np.arange(1, nodes.size + 1), lconn, lhs, rhs)
    # complete nodes: lattice index in (2 i0, 2 i0 + 6) exclusive unless at the domain boundary
    ix, iy, iz = nodes % sx, (nodes // sx) % sx, nodes // (sx * sx)
    def inside(q):
        lo, hi = 2 * i0, 2 * i0 + 6
        return ((q > lo) | (q == 0)) & ((q < hi) | (q == sx - 1))
    complete = np.nonzero(inside(ix) & inside(iy) & inside(iz))[0]
    worst_v = worst_b = 0.0
    nrows = 0
    for ln in complete:
        gn = nodes[ln]
        for i in range(3):
            r, gr = 3 * ln + i, 3 * gn + i
            ref_cols = ci[rp[r]:rp[r + 1]]
            ref_vals = vals[rp[r]:rp[r + 1]]
            assert np.array_equal(nodes[ref_cols // 3] * 3 + ref_cols % 3, gci[grp[gr]:grp[gr + 1]]), "row structure differs"
            mine = got_v[grp[gr]:grp[gr + 1]]
            scale = max(np.abs(ref_vals).max(), 1e-300)
            worst_v = max(worst_v, float((np.abs(mine - ref_vals) / np.maximum(np.abs(ref_vals), 1e-2 * scale)).max()))
            rs = float(np.abs(ref_vals).sum() * np.abs(vel).max())
            worst_b = max(worst_b, abs(got_b[gr] - b[r]) / max(abs(b[r]), 1e-2 * rs))
            nrows += 1
    return nrows, worst_v, worst_b


def run(n=48, steps=10, warmup=3):
    import torch
    from nalu_b200 import mesh as M
    from nalu_b200.linsys import DeviceCsrLinearSystem
    dev = torch.device("cuda:0")
    m = M.hex27_box(n, n, n)
    rng = np.random.default_rng(7)
    m.coords = m.coords + 0.04 * rng.uniform(-1, 1, size=m.coords.shape)
    vel, visc, rho = fields(m.coords)
    t0 = time.time()
    ls = DeviceCsrLinearSystem(m, 3)
    graph_s = time.time() - t0
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
    coords, v, mu, r = t(m.coords), t(vel), t(visc), t(rho)
    mdot = ls.mdot_scs(coords, r, v)
    ls.allocate(dev)
    d = ls.make_desc("momentum_adv_diff", fields=dict(coordinates=coords, velocity=v, viscosity=mu), mass_flow_rate=mdot,
                     include_divu=1.0)
    l0 = ls.L.nalu_b200_launch_count()
    for _ in range(warmup):
        ls.zero_system()
        mode = ls.assemble_elem([d])
    torch.cuda.synchronize()
    l1 = ls.L.nalu_b200_launch_count()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    asm_ms = 0.0
    tot_ms = 0.0
    for _ in range(steps):
        e0.record()
        ls.zero_system()
        e1.record()
        ls.assemble_elem([d])
        e2.record()
        torch.cuda.synchronize()
        asm_ms += e1.elapsed_time(e2)
        tot_ms += e0.elapsed_time(e2)
    asm_ms /= steps
    tot_ms /= steps
    ne = m.n_elems
    nnz = int(ls.nnz_owned)
    # algorithmic bytes of a step: every CSR value and rhs entry written once, nodal fields (x, u: 3, mu: 1) and mdot read once,
    # connectivity and slot map read once
    alg_bytes = nnz * 8 + m.n_nodes * 3 * 8 + m.n_nodes * 7 * 8 + ne * (216 * 8 + 27 * 4 + 729)
    peak = 6456.2
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    got_v = ls.owned_values().cpu().numpy()
    got_b = ls.owned_rhs().cpu().numpy()
    rp = ls.rowptr().astype(np.int64)
    ci = ls.colidx().astype(np.int64)
    checks = {}
    worst = [0, 0.0, 0.0]
    for i0 in ([0, max(0, n // 2 - 1), n - 3] if n >= 6 else [0]):
        nrows, wv, wb = subbox_check(m, n, i0, vel, visc, rho, got_v, got_b, rp, ci)
        worst = [worst[0] + nrows, max(worst[1], wv), max(worst[2], wb)]
    checks["sampled_rows_vs_oracle"] = {"rows": worst[0], "values_max_rel_err": worst[1], "rhs_max_rel_err": worst[2],
                                        "tolerance": 1e-12}
    checks["finite"] = bool(np.isfinite(got_v).all() and np.isfinite(got_b).all())
    checks["geometry_errors"] = int(ls.geometry_errors())
    # cpu baseline: the oracle on a bounded sample
    from oracle import hex27 as H
    ns = min(ne, 256)
    conn = m.elem_nodes[:ns]
    tc = time.time()
    md = H.calc_mass_flow_rate_scs(conn, m.coords, rho, vel)
    H.momentum_adv_diff(conn, m.coords, vel, visc, md, include_divu=1.0)
    cpu_s = time.time() - tc
    out = {
        "metric": "hex27_momentum_assembly_throughput", "unit": "Melem/s", "value": ne / tot_ms / 1e3, "n_gpus": 1,
        "steps": steps, "warmup": warmup, "ms_per_step": tot_ms, "assemble_ms": asm_ms, "higher_is_better": True,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "BASELINE config 5: %d^3 Hex27 box (perturbed, curved elements), [MomentumAdvDiffElemKernel(includeDivU=1)]"
                               " on AlgTraitsHex27, zeroSystem + assemble" % n,
                   "elements": ne, "nodes": m.n_nodes, "dof_rows": 3 * m.n_nodes, "nnz": nnz, "l2": "working set %.1f GB >> L2" % (nnz * 8 / 1e9),
                   "graph_build_s": graph_s, "colours": int(ls.elem_colors()[0])},
        "kernel_family": mode, "gpu_launches": int(l1 - l0) // max(warmup, 1),
        "nodes_per_s_M": m.n_nodes / tot_ms / 1e3,
        "roofline": {"bound": "hbm", "achieved": alg_bytes / tot_ms / 1e6, "peak": peak, "unit": "GB/s",
                     "frac": alg_bytes / tot_ms / 1e6 / peak, "traffic": None,
                     "algorithmic_bytes_per_elem": alg_bytes / ne,
                     "fp64": {"flop_per_elem_static": 2 * (216 * (27 * 18 + 27 + 60) + 243 * 16 * 3 * 22), "note": "FMA = 2 flop, static count of phase B + C"}},
        "cpu_baseline": {"value": ns / cpu_s / 1e6, "unit": "Melem/s", "cores": 1, "kind": "port",
                         "sample": "%d elements through oracle/hex27.py (numpy)" % ns},
        "checks": checks,
    }
    return out


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    warmup = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    print(json.dumps(run(n, steps, warmup)))


if __name__ == "__main__":
    main()
