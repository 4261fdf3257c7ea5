"""BASELINE config 5: Hex27 (P2 CVFEM) box, MomentumAdvDiffElemKernel<AlgTraitsHex27> + sumInto, on 1..8 B200.

    python tools/bench_hex27.py [N=48] [steps=10] [warmup=3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 tools/bench_hex27.py [N] [steps] [warmup]

One GPU: an N^3 box.  G GPUs: weak scaling, N x N x (N G) elements in z-slabs of N layers, shared node planes owned by the lower
rank, loadComplete inside the library (nalu_b200_load_complete: pack -> ncclSend / ncclRecv -> unpack-add).  A step =
zeroSystem + one nalu_b200_assemble_elem (8 colour launches) [+ loadComplete].  Inputs resident in HBM, CUDA events on the
launching stream, barrier + synchronize on both sides, max over ranks; working set (4.2 GB of CSR values per GPU at N = 48) far
larger than L2.  After the timed region the complete rows of 3^3-element sub-boxes (domain corners, slab middle and -- with
several ranks -- across every rank's upper interface, i.e. rows completed by the export-add) are compared with the oracle
(oracle/hex27.py); on one GPU the oracle is also timed on a bounded sample (cpu_baseline, numpy, one core).  Rank 0 prints one
JSON line.
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

AMPLITUDE = 0.04


def fields(coords):
    x = coords
    vel = np.stack([np.sin(0.9 * x[:, 1]) + 0.3 * np.cos(x[:, 2]), np.cos(0.7 * x[:, 0]) * np.sin(0.5 * x[:, 2]),
                    0.2 * np.sin(x[:, 0]) * np.cos(x[:, 1]) - 0.5], axis=1)
    visc = 0.3 + 0.1 * np.cos(x[:, 0] + 2.0 * x[:, 1])
    rho = 1.0 + 0.2 * np.sin(x[:, 2])
    return vel, visc, rho


def subbox_check(nx, ny, nz, origin, row_gids, rp, ci, col_gids, got_v, got_b):
    """complete rows of the 3^3-element sub-box whose first element is `origin` (global element indices), rebuilt from global
    node ids alone: nodes strictly inside its node range (or on the domain boundary) have all their elements in the sub-box.
    Rows this rank does not own are skipped.  Returns (rows compared, worst values error, worst rhs error)."""
    from nalu_b200 import mesh as M
    from oracle import hex27 as H
    sx, sy, sz = 2 * nx + 1, 2 * ny + 1, 2 * nz + 1
    i0, j0, k0 = origin
    sub = M.hex27_box(3, 3, 3)
    lat = np.rint(sub.coords * 2).astype(np.int64)                       # local lattice indices 0..6
    gx, gy, gz = lat[:, 0] + 2 * i0, lat[:, 1] + 2 * j0, lat[:, 2] + 2 * k0
    sub.node_gid = 1 + gx + sx * (gy + sy * gz)
    sub.coords = 0.5 * np.stack([gx, gy, gz], axis=1).astype(np.float64)
    M.perturb_coords_by_gid(sub, AMPLITUDE)
    vel, visc, rho = fields(sub.coords)
    c = sub.coords - sub.coords[0]  # the assembled system is translation invariant; the reference's rounding error is not
    mdot = H.calc_mass_flow_rate_scs(sub.elem_nodes, c, rho, vel)
    lhs, rhs = H.momentum_adv_diff(sub.elem_nodes, c, vel, visc, mdot, include_divu=1.0)
    orp, oci, vals, b = H.assemble_csr(np.arange(1, sub.n_nodes + 1), sub.elem_nodes, lhs, rhs)

    def inside(q, lo, top):
        return ((q > lo) | (q == 0)) & ((q < lo + 6) | (q == top))
    complete = np.nonzero(inside(gx, 2 * i0, sx - 1) & inside(gy, 2 * j0, sy - 1) & inside(gz, 2 * k0, sz - 1))[0]
    worst_v = worst_b = 0.0
    nrows = 0
    vmax = float(np.abs(vel).max())
    for ln in complete:
        for i in range(3):
            dof_gid = 3 * (sub.node_gid[ln] - 1) + i + 1
            gr = int(np.searchsorted(row_gids, dof_gid))
            if gr >= row_gids.size or row_gids[gr] != dof_gid:
                continue  # owned by another rank
            r = 3 * ln + i
            ref_cols = oci[orp[r]:orp[r + 1]]
            ref_vals = vals[orp[r]:orp[r + 1]]
            ref_gid = 3 * (sub.node_gid[ref_cols // 3] - 1) + ref_cols % 3 + 1
            mine_gid = col_gids[ci[rp[gr]:rp[gr + 1]]]
            mine = got_v[rp[gr]:rp[gr + 1]]
            # this rank's row may list columns in another order (owned, shared, remote): compare by global id
            o1, o2 = np.argsort(ref_gid), np.argsort(mine_gid)
            assert np.array_equal(ref_gid[o1], mine_gid[o2]), "row structure differs"
            scale = max(np.abs(ref_vals).max(), 1e-300)
            worst_v = max(worst_v, float((np.abs(mine[o2] - ref_vals[o1]) / np.maximum(np.abs(ref_vals[o1]), 1e-2 * scale)).max()))
            rs = float(np.abs(ref_vals).sum() * vmax)
            worst_b = max(worst_b, abs(got_b[gr] - b[r]) / max(abs(b[r]), 1e-2 * rs))
            nrows += 1
    return nrows, worst_v, worst_b


def run(n=48, steps=10, warmup=3):
    import torch
    import torch.distributed as dist
    from nalu_b200 import dist as nd
    from nalu_b200 import mesh as M
    from nalu_b200.linsys import DeviceCsrLinearSystem
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    own_group = False
    if world > 1:
        torch.cuda.set_device(local_rank)
        if not dist.is_initialized():
            if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
                os.environ["NCCL_DEBUG"] = "WARN"
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            own_group = True
    dev = torch.device("cuda", local_rank if world > 1 else torch.cuda.current_device())
    if world > 1:
        from nalu_b200 import lib as _l
        _l.tune(host_threads=max(1, len(os.sched_getaffinity(0)) // int(os.environ.get("LOCAL_WORLD_SIZE", world))))
    nzg = n * world
    m = M.hex27_box_rank(n, n, nzg, rank, world)
    M.perturb_coords_by_gid(m, AMPLITUDE)
    vel, visc, rho = fields(m.coords)
    t0 = time.time()
    ls = DeviceCsrLinearSystem(m, 3, exchange=nd.make_exchange(rank, world) if world > 1 else None)
    graph_s = time.time() - t0
    comm = nd.make_comm(rank, world) if world > 1 else None
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
    coords, v, mu, r = t(m.coords), t(vel), t(visc), t(rho)
    mdot = ls.mdot_scs(coords, r, v)
    ls.allocate(dev)
    d = ls.make_desc("momentum_adv_diff", fields=dict(coordinates=coords, velocity=v, viscosity=mu), mass_flow_rate=mdot,
                     include_divu=1.0)

    def step():
        ls.zero_system()
        mode = ls.assemble_elem([d])
        if comm is not None:
            ls.load_complete(comm)
        return mode

    l0 = ls.L.nalu_b200_launch_count()
    for _ in range(max(warmup, 1)):
        mode = step()
    torch.cuda.synchronize()
    l1 = ls.L.nalu_b200_launch_count()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    tot_ms = e0.elapsed_time(e1) / steps
    if world > 1:
        tt = torch.tensor([tot_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        tot_ms = float(tt.item())
    # assembly alone (no zeroSystem, no exchange), this rank
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ls.zero_system()
    a0.record()
    ls.assemble_elem([d])
    a1.record()
    torch.cuda.synchronize()
    asm_ms = a0.elapsed_time(a1)
    if comm is not None:
        ls.load_complete(comm)
        torch.cuda.synchronize()
    ne = m.n_elems
    nnz = int(ls.nnz_owned)
    n_owned_nodes = int(ls.n_owned_rows) // 3
    # algorithmic bytes of a step on one GPU: every owned CSR value and rhs entry written once, nodal fields (x, u: 3, mu: 1)
    # and mdot read once, connectivity and slot map read once
    alg_bytes = nnz * 8 + n_owned_nodes * 3 * 8 + m.n_nodes * 7 * 8 + ne * (216 * 8 + 27 * 4 + 729)
    peak = 6456.2
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    traffic = None
    try:
        tj = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")))["assemble_hex27"]
        if tj["elements"] == ne:
            traffic = {"bytes": tj["bytes"], "source": "static: " + tj["source"]}
    except Exception:
        pass
    got_v = ls.owned_values().cpu().numpy()
    got_b = ls.owned_rhs().cpu().numpy()
    rp, ci = ls.rowptr().astype(np.int64), ls.colidx().astype(np.int64)
    row_gids, col_gids = ls.row_gids(), ls.col_gids()
    k_lo = n * rank
    origins = [(0, 0, k_lo + max(0, n // 2 - 1))]                 # a domain edge, middle of the slab
    if rank == 0:
        origins.append((0, 0, 0))                                    # domain corner
    if rank == world - 1:
        origins.append((n - 3, n - 3, nzg - 3))                      # far corner
    else:
        # straddles this rank's upper interface: the shared node plane is owned here, and its rows are complete only after
        # the upper rank's contributions arrived through loadComplete
        origins.append((max(0, n // 2 - 1), 0, k_lo + n - 2))
    worst = [0, 0.0, 0.0]
    if n >= 3:
        for o in origins:
            nrows, wv, wb = subbox_check(n, n, nzg, o, row_gids, rp, ci, col_gids, got_v, got_b)
            worst = [worst[0] + nrows, max(worst[1], wv), max(worst[2], wb)]
    finite = bool(np.isfinite(got_v).all() and np.isfinite(got_b).all())
    geom = int(ls.geometry_errors())
    if world > 1:
        allw = [None] * world
        dist.all_gather_object(allw, (worst, finite, geom))
        worst = [sum(w[0][0] for w in allw), max(w[0][1] for w in allw), max(w[0][2] for w in allw)]
        finite = all(w[1] for w in allw)
        geom = max(w[2] for w in allw)
    checks = {"sampled_rows_vs_oracle": {"rows": worst[0], "values_max_rel_err": worst[1], "rhs_max_rel_err": worst[2],
                                         "tolerance": 1e-12, "passed": bool(worst[0] > 0 and worst[1] < 1e-12 and worst[2] < 1e-12),
                                         "across_rank_interfaces": world > 1},
              "finite": finite, "geometry_errors": geom}
    mlist = None
    if world == 1:
        # the low-Mach momentum list [MomentumMass(lumped), MomentumAdvDiff] with the mass fused into the same launches
        f = dict(coordinates=coords, velocity=v, velocity_n=t(0.9 * vel + 0.01), velocity_nm1=t(0.8 * vel - 0.02), viscosity=mu,
                 density=r, density_n=t(0.98 * rho), density_nm1=t(0.95 * rho), dpdx=t(0.1 * vel))
        dl = [ls.make_desc("momentum_mass", fields=f, options=("lumped_mass",), dt=0.05, gamma1=1.5, gamma2=-2.0, gamma3=0.5), d]
        for _ in range(2):
            ls.zero_system()
            ls.assemble_elem(dl)
        torch.cuda.synchronize()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        for _ in range(steps):
            ls.zero_system()
            ls.assemble_elem(dl)
        b1.record()
        torch.cuda.synchronize()
        lms = b0.elapsed_time(b1) / steps
        mlist = {"kernels": ["MomentumMassElemKernel(lumped)", "MomentumAdvDiffElemKernel(includeDivU=1)"], "value": ne / lms / 1e3,
                 "unit": "Melem/s", "ms_per_step": lms, "finite": bool(torch.isfinite(ls.values).all())}
    onedof = None
    if world == 1:
        # the one-dof Hex27 kernels on the same mesh: ContinuityAdvElemKernel and ScalarAdvDiffElemKernel (their own 1-dof graph)
        ls1 = DeviceCsrLinearSystem(m, 1)
        ls1.allocate(dev)
        x = m.coords
        pr = t(0.5 * np.cos(0.8 * x[:, 0]) * np.sin(0.6 * x[:, 1]) + 0.1 * x[:, 2])
        gpx = t(0.1 * vel)
        qq = t(1.0 + 0.5 * np.sin(x[:, 0] + 0.5 * x[:, 1]) * np.cos(0.7 * x[:, 2]))
        dk = {"continuity_adv": ls1.make_desc("continuity_adv", fields=dict(coordinates=coords, velocity=v, density=r, pressure=pr, dpdx=gpx),
                                              dt=0.05, gamma1=1.5),
              "scalar_adv_diff": ls1.make_desc("scalar_adv_diff", fields=dict(coordinates=coords, scalarq=qq, viscosity=mu), mass_flow_rate=mdot)}
        onedof = {}
        for name, dd in dk.items():
            for _ in range(2):
                ls1.zero_system()
                ls1.assemble_elem([dd])
            torch.cuda.synchronize()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for _ in range(steps):
                ls1.zero_system()
                ls1.assemble_elem([dd])
            c1.record()
            torch.cuda.synchronize()
            oms = c0.elapsed_time(c1) / steps
            onedof[name] = {"value": ne / oms / 1e3, "unit": "Melem/s", "ms_per_step": oms, "finite": bool(torch.isfinite(ls1.values).all()),
                            "nnz": int(ls1.nnz_owned)}
        del ls1
    cpu = None
    if world == 1:
        # cpu baseline: the oracle on a bounded sample
        from oracle import hex27 as H
        ns = min(ne, 256)
        conn = m.elem_nodes[:ns]
        tc = time.time()
        md = H.calc_mass_flow_rate_scs(conn, m.coords, rho, vel)
        H.momentum_adv_diff(conn, m.coords, vel, visc, md, include_divu=1.0)
        cpu_s = time.time() - tc
        cpu = {"value": ns / cpu_s / 1e6, "unit": "Melem/s", "cores": 1, "kind": "port",
               "sample": "%d elements through oracle/hex27.py (numpy)" % ns}
    out = {
        "metric": "hex27_momentum_assembly_throughput", "unit": "Melem/s", "value": ne * world / tot_ms / 1e3, "n_gpus": world,
        "steps": steps, "warmup": max(warmup, 1), "ms_per_step": tot_ms, "assemble_ms": asm_ms, "higher_is_better": True,
        "scaling": "weak", "dtype": "f64", "data": "synthetic",
        "config": {"workload": "BASELINE config 5: %dx%dx%d Hex27 box (perturbed, curved elements), [MomentumAdvDiffElemKernel(includeDivU=1)]"
                               " on AlgTraitsHex27, zeroSystem + assemble%s" % (n, n, nzg, " + loadComplete" if world > 1 else ""),
                   "elements_per_gpu": ne, "nodes_per_gpu": m.n_nodes, "dof_rows_per_gpu": int(ls.n_owned_rows), "nnz_per_gpu": nnz,
                   "partition": "z-slabs x%d" % world, "l2": "working set %.1f GB per GPU >> L2" % (nnz * 8 / 1e9),
                   "graph_build_s": graph_s, "colours": int(ls.elem_colors()[0])},
        "kernel_family": mode, "gpu_launches": int(l1 - l0) // max(warmup, 1),
        "nodes_per_s_M": n_owned_nodes * world / tot_ms / 1e3,
        "roofline": {"bound": "hbm", "achieved": alg_bytes / tot_ms / 1e6, "peak": peak, "unit": "GB/s",
                     "frac": alg_bytes / tot_ms / 1e6 / peak, "traffic": traffic,
                     "algorithmic_bytes_per_elem": alg_bytes / ne,
                     "fp64": {"flop_per_elem_static": 2 * (216 * (27 * 23 + 60) + 243 * 16 * 81),
                              "note": "FMA = 2 flop, static count of phase B + C"}},
        "cpu_baseline": cpu, "momentum_list": mlist, "one_dof_kernels": onedof,
        "checks": checks,
    }
    if own_group:
        dist.barrier()
        dist.destroy_process_group()
    return out if rank == 0 else None


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    warmup = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    out = run(n, steps, warmup)
    if out is not None:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
