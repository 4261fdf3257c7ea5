#!/usr/bin/env python3
"""bench.py -- Hex8 momentum assembly throughput on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # reference algorithm on host cores

One "step" = one AssembleElemSolverAlgorithm::execute of MomentumAdvDiffElemKernel over all
locally-owned Hex8 elements of a perturbed structured box (LHS + RHS computed AND scattered into
the CSR), followed, when N > 1, by loadComplete (the shared-row halo sum).  Weak scaling: every
GPU holds an n^3-element slab (n=256 by default: BASELINE config 3 at N=1; 8 x 256^3 = the 512^3
element count of config 4 at N=8).  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "hex8_momentum_assembly_throughput"
UNIT = "Melem/s"
# algorithmic bytes per element, BASELINE.md section 3 rule:
#   8*(nodal comps read)*nNodes + 8*(elem comps)*nElem + 32*nElem + 8*nnz + 8*ndof*nNodes


def algorithmic_bytes(n_nodes, n_elems, nnz, ndof, nodal_comps, elem_comps):
    return 8 * nodal_comps * n_nodes + 8 * elem_comps * n_elems + 32 * n_elems + 8 * nnz + 8 * ndof * n_nodes


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, STREAM-style copy)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                self.samples.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx.append(float(s[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle (literal restatement of the reference algorithm)
# timed on the host cores, P slab partitions on P threads standing in for P MPI ranks
# --------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, cores=None, nx=64, layers_per_rank=16):
    import oracle
    from nalu_b200 import fields, mesh
    P = cores or min(os.cpu_count() or 1, 64)
    parts = [mesh.generated_box_rank(nx, nx, layers_per_rank * P, r, P) for r in range(P)]
    for m in parts:
        mesh.perturb_coords_hash(m, 0.12)
    ls = oracle.LinSys(parts, 3)  # graph construction is outside the timed region, as for the GPU arm
    keep = []
    descs = []
    for m in parts:
        c = m.coords
        vel = fields.velocity(c)
        rho = fields.smooth_density(c)
        mdot = oracle.calc_mass_flow_rate_scs(m.elem_nodes, c, rho, vel)
        descs.append(oracle.make_desc("momentum_adv_diff", keep, coordinates=c, velocity=vel,
                                      viscosity=fields.smooth_viscosity(c), massFlowRate=mdot, includeDivU=1.0))
    n_elems = sum(m.n_elems for m in parts)

    def one_step():
        ls.zero()
        t0 = time.perf_counter()
        if P == 1:
            ls.assemble_elem(0, [descs[0]])
        else:
            th = [threading.Thread(target=ls.assemble_elem, args=(r, [descs[r]])) for r in range(P)]
            for t in th:
                t.start()
            for t in th:
                t.join()
        ls.load_complete()
        return time.perf_counter() - t0

    for _ in range(warmup):
        one_step()
    times = [one_step() for _ in range(steps)]
    t = float(np.mean(times))
    return {"value": n_elems / t / 1e6, "unit": UNIT, "cores": P, "kind": "port",
            "sample": f"{nx}x{nx}x{layers_per_rank * P} Hex8 box in {P} z-slabs ({n_elems} elements), "
                      f"momentum adv-diff + sumInto + export-add, {steps} steps, oracle C restatement -O3",
            "ms_per_step": t * 1e3, "n_elems": n_elems}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_run(max(1, args.steps), max(1, min(args.warmup, 2)))
    line = {"metric": METRIC, "value": r["value"], "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "hex8_box_momentum_advdiff_assembly", "sample": r["sample"]},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from nalu_b200 import fields, lib, mesh
    from nalu_b200.linsys import DeviceCsrLinearSystem

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the assembly path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout at NCCL_DEBUG=VERSION; stdout carries exactly one JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    n = args.n
    L = lib.lib()
    lib.tune(box_kernel=args.box_kernel)

    # ---- mesh, graph (once per mesh; outside the timed region) ------------------------------
    t0 = time.perf_counter()
    m = mesh.generated_box_rank(n, n, n * world, rank, world)
    mesh.perturb_coords_hash(m, 0.12)

    from nalu_b200 import dist as nd
    halo = args.halo if world > 1 else "none"
    ls = DeviceCsrLinearSystem(m, 3, exchange=nd.make_exchange(rank, world) if world > 1 else None)
    ls.allocate(dev, raw=(halo == "pull"))
    windows = None
    if halo == "pull":
        try:
            windows = nd.PeerWindows(ls, rank, world)
        except Exception as e:  # CUDA IPC unavailable in this container: say so and use NCCL
            sys.stderr.write(f"[rank {rank}] peer windows unavailable ({e}); falling back to NCCL send/recv\n")
            halo = "nccl"
        flag = torch.tensor([1 if halo == "pull" else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            halo, windows = "nccl", None
    t_graph = time.perf_counter() - t0

    # ---- state: host (pinned) + device ----------------------------------------------------------
    c = m.coords
    h = {"coordinates": c, "velocity": fields.velocity(c), "viscosity": fields.smooth_viscosity(c)}
    rho = fields.smooth_density(c)
    d = {k: torch.as_tensor(np.ascontiguousarray(v), device=dev) for k, v in h.items()}
    d_rho = torch.as_tensor(rho, device=dev)
    mdot = ls.mdot_scs(d["coordinates"], d_rho, d["velocity"])
    desc = ls.make_desc("momentum_adv_diff", fields=d, mass_flow_rate=mdot, include_divu=args.divu)

    mode = args.mode
    probe = ls.assemble_elem([desc], mode=mode, overwrite=True)  # also a first warm-up
    overwrite = probe == 2
    send_peers, recv_peers = ls.halo_peers()

    def step():
        if not overwrite:
            ls.zero_system()   # the general kernel accumulates: zeroSystem is part of its step
        ls.assemble_elem([desc], mode=mode, overwrite=overwrite)
        if world > 1:
            load_complete()

    def load_complete():
        if windows is not None:
            # owners read the sharers' rows straight across NVLink; the barrier orders "my neighbour finished
            # assembling" before "I pull its rows", a second one keeps it from re-assembling under my reads
            windows.load_complete(lambda: (torch.cuda.synchronize(), dist.barrier()))
            torch.cuda.synchronize()
            dist.barrier()
        else:
            ls.load_complete_nccl()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # ---- timed region: device-resident inputs ---------------------------------------------------
    launches0 = L.nalu_b200_launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    asm_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        ev[i][0].record()
        if not overwrite:
            ls.zero_system()
        asm_ev[i][0].record()
        ls.assemble_elem([desc], mode=mode, overwrite=overwrite)
        asm_ev[i][1].record()
        if world > 1:
            load_complete()
        ev[i][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = L.nalu_b200_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    step_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    asm_ms = float(np.mean([a.elapsed_time(b) for a, b in asm_ev]))
    tt = torch.tensor([step_ms, asm_ms, t_wall * 1e3 / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    step_ms, asm_ms, wall_ms = [float(x) for x in tt.cpu()]
    step_ms = max(step_ms, 1e-9)

    n_elems_rank = m.n_elems
    n_elems_total = n_elems_rank * world
    value = n_elems_total / (step_ms * 1e-3) / 1e6

    # ---- roofline of the dominant kernel (the assembly kernel), per launch ---------------------
    peak, peak_src = load_peaks()
    nnz = ls.nnz_owned + ls.nnz_shared
    bytes_alg = algorithmic_bytes(m.n_nodes, n_elems_rank, nnz, 3, nodal_comps=7, elem_comps=12)
    achieved = bytes_alg / (asm_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src, "kernel": (f"assemble_box{args.box_kernel}" if overwrite else "assemble_general"),
                "algorithmic_bytes_per_launch": bytes_alg, "bytes_per_element": bytes_alg / n_elems_rank,
                "kernel_ms": asm_ms}
    tf = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tf):
        try:
            with open(tf) as f:
                roofline["traffic"] = json.load(f).get(roofline["kernel"])
        except Exception:
            pass

    # ---- e2e: host buffers in, residual (RHS) + matrix checksum out, every step ------------------
    pinned = {k: torch.as_tensor(np.ascontiguousarray(v)).pin_memory() for k, v in h.items()}
    pinned["mdot"] = mdot.cpu().pin_memory()
    rhs_host = torch.empty(ls.n_owned_rows, dtype=torch.float64).pin_memory()
    h2d = sum(t.numel() * 8 for t in pinned.values())
    d2h = rhs_host.numel() * 8
    e2e_steps = max(1, min(args.steps, 5))

    def e2e_step():
        for k in ("coordinates", "velocity", "viscosity"):
            d[k].copy_(pinned[k], non_blocking=True)
        mdot.copy_(pinned["mdot"], non_blocking=True)
        step()
        rhs_host.copy_(ls.owned_rhs(), non_blocking=True)
        torch.cuda.synchronize()

    def matrix_fingerprint():  # outside the timed region: serial and pipelined schedules must agree bit for bit
        v = ls.owned_values()
        return torch.stack([v.sum(), torch.linalg.vector_norm(v, 1), ls.owned_rhs().sum()]).cpu()

    # The same step with the host->device copies of the next node planes overlapped with the assembly of
    # the current ones (nalu_b200_assemble_elem_zrange on the compute stream, copies on a second stream).
    NZl = n + 1
    nxy_nodes, nxy_elems = (n + 1) * (n + 1), n * n
    nchunk = max(1, min(args.e2e_chunks, n // 16))
    zb = [round(i * NZl / nchunk) for i in range(nchunk + 1)]
    copy_stream = torch.cuda.Stream(device=dev)
    d2h_stream = torch.cuda.Stream(device=dev)
    rows_per_plane = 3 * nxy_nodes  # one rank: owned rows are the nodes in lattice order

    def e2e_step_pipelined():
        cur = torch.cuda.current_stream(dev)
        copy_stream.wait_stream(cur)
        planes_done, layers_done = 0, 0
        for c in range(nchunk):
            z0, z1 = zb[c], zb[c + 1]
            p_hi = min(z1 + 1, NZl)        # node planes z0-1 .. z1
            l_hi = min(z1, n)              # element layers z0-1 .. z1-1
            with torch.cuda.stream(copy_stream):
                if p_hi > planes_done:
                    a, b = planes_done * nxy_nodes, p_hi * nxy_nodes
                    for k in ("coordinates", "velocity", "viscosity"):
                        d[k][a:b].copy_(pinned[k][a:b], non_blocking=True)
                    planes_done = p_hi
                if l_hi > layers_done:
                    a, b = layers_done * nxy_elems, l_hi * nxy_elems
                    mdot[a:b].copy_(pinned["mdot"][a:b], non_blocking=True)
                    layers_done = l_hi
                evc = torch.cuda.Event()
                evc.record(copy_stream)
            cur.wait_event(evc)
            ls.assemble_elem_zrange([desc], z0, z1, overwrite=True)
            if world == 1:
                # the residual rows of a finished range go home while the next range is assembled (full duplex)
                evk = torch.cuda.Event()
                evk.record(cur)
                with torch.cuda.stream(d2h_stream):
                    d2h_stream.wait_event(evk)
                    a, b = z0 * rows_per_plane, z1 * rows_per_plane
                    rhs_host[a:b].copy_(ls.owned_rhs()[a:b], non_blocking=True)
        if world > 1:
            load_complete()
            rhs_host.copy_(ls.owned_rhs(), non_blocking=True)
        torch.cuda.synchronize()

    def time_e2e(fn):
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            fn()
        barrier()
        ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
        te = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return float(te.cpu()[0])

    serial_ms = time_e2e(e2e_step)
    fp_serial, rhs_serial = matrix_fingerprint(), rhs_host.clone()
    use_pipe = overwrite and nchunk > 1
    e2e_ms = time_e2e(e2e_step_pipelined) if use_pipe else serial_ms
    if use_pipe:  # same system, whichever schedule
        assert torch.equal(fp_serial, matrix_fingerprint()) and torch.equal(rhs_serial, rhs_host), \
            "pipelined e2e step changed the assembled system"
    e2e = {"value": n_elems_total / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "serial_ms_per_step": serial_ms,
           "schedule": (f"{nchunk} node-plane ranges: host->device copy of range c+1 overlaps the assembly of range c "
                        "(nalu_b200_assemble_elem_zrange)") if use_pipe else "copy everything, then assemble",
           "note": "inputs (coords, velocity, viscosity, mdot) from pinned host memory every step; the RHS "
                   "(nonlinear residual) is read back to the host; the CSR values stay device-resident for a "
                   "device solver"}

    # ---- size-independent correctness properties on the timed result --------------------------
    vals = ls.owned_values()
    sum_v, sum_abs = float(vals.sum()), float(vals.abs().sum())
    props = {"sum_values_over_sum_abs": sum_v / sum_abs if sum_abs else None,
             "finite": bool(torch.isfinite(vals).all()) and bool(torch.isfinite(ls.rhs).all()),
             "geometry_errors": ls.geometry_errors()}

    if world > 1:
        props["multi_rank_selfcheck"] = multi_rank_selfcheck(rank, world, dev, halo)

    # ---- the other pass of BASELINE config 3: continuity (ContinuityAdvElemKernel) on the same mesh, device-resident
    # inputs, outside the timed region of the headline metric; reported beside it, not folded into `value`
    continuity = None
    if world == 1 and not args.no_continuity:
        ls1 = DeviceCsrLinearSystem(m, 1).allocate(dev)
        f1 = {"coordinates": d["coordinates"], "velocity": d["velocity"], "density": d_rho,
              "pressure": torch.as_tensor(fields.pressure(c), device=dev), "dpdx": torch.as_tensor(np.ascontiguousarray(fields.dpdx(c)), device=dev)}
        desc1 = ls1.make_desc("continuity_adv", fields=f1, dt=0.05, gamma1=1.5)
        fam = 0
        for _ in range(3):
            fam = ls1.assemble_elem([desc1], mode=mode if mode != "box" else "auto", overwrite=True)
        torch.cuda.synchronize()
        cev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
        for a, b in cev:
            a.record()
            if fam != 2:
                ls1.zero_system()
            ls1.assemble_elem([desc1], mode=mode if mode != "box" else "auto", overwrite=(fam == 2))
            b.record()
        torch.cuda.synchronize()
        c_ms = float(np.mean([a.elapsed_time(b) for a, b in cev]))
        c_bytes = algorithmic_bytes(m.n_nodes, n_elems_rank, ls1.nnz_owned, 1, nodal_comps=11, elem_comps=0)
        continuity = {"kernel": "assemble_cont_box" if fam == 2 else "assemble_general", "value": n_elems_rank / c_ms / 1e3,
                      "unit": UNIT, "ms": c_ms, "bytes_per_element": c_bytes / n_elems_rank,
                      "achieved_GBps": c_bytes / c_ms / 1e6, "frac_of_hbm": c_bytes / c_ms / 1e6 / peak, "bound": "fp64",
                      "finite": bool(torch.isfinite(ls1.values).all()), "geometry_errors": ls1.geometry_errors(),
                      "momentum_plus_continuity_two_passes": {"value": n_elems_rank / (asm_ms + c_ms) / 1e3, "unit": UNIT,
                                                               "frac_of_hbm": (bytes_alg + c_bytes) / (asm_ms + c_ms) / 1e6 / peak}}
        del ls1

    # ---- the momentum kernel LIST of the low-Mach equation system (BASELINE config 3: lumped mass + adv-diff) through
    # the public call: tiled write-once kernel for adv-diff, per-SCV kernel for the lumped mass on top
    momentum_list = None
    if world == 1 and not args.no_continuity and overwrite:
        vel_h = h["velocity"]
        fl = dict(d)
        fl.update(velocity_n=torch.as_tensor(np.ascontiguousarray(0.9 * vel_h), device=dev),
                  velocity_nm1=torch.as_tensor(np.ascontiguousarray(0.8 * vel_h), device=dev), density=d_rho,
                  density_n=0.98 * d_rho, density_nm1=0.95 * d_rho,
                  dpdx=torch.as_tensor(np.ascontiguousarray(fields.dpdx(c)), device=dev))
        dl = [ls.make_desc("momentum_mass", fields=fl, options=("lumped_mass",), dt=0.05, gamma1=1.5, gamma2=-2.0, gamma3=0.5), desc]
        fam = 0
        for _ in range(2):
            fam = ls.assemble_elem(dl, mode="auto", overwrite=True)
        torch.cuda.synchronize()
        lev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
        for a, b in lev:
            a.record()
            ls.assemble_elem(dl, mode="auto", overwrite=True)
            b.record()
        torch.cuda.synchronize()
        l_ms = float(np.mean([a.elapsed_time(b) for a, b in lev]))
        momentum_list = {"kernels": ["MomentumMassElemKernel(lumped)", f"MomentumAdvDiffElemKernel(includeDivU={args.divu:g})"],
                         "kernel_family": fam, "value": n_elems_rank / l_ms / 1e3, "unit": UNIT, "ms": l_ms,
                         "finite": bool(torch.isfinite(ls.values).all())}
        ls.assemble_elem([desc], mode=mode, overwrite=True)  # leave the headline system in place

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(steps=2, warmup=1)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"hex8_box_{n}x{n}x{n * world}_momentum_advdiff_assembly",
                           "elements_per_gpu": n_elems_rank, "kernels": [f"MomentumAdvDiffElemKernel(includeDivU={args.divu:g})"],
                           "partition": f"z-slabs x{world}", "halo": {"nccl": "pack kernel -> NCCL send/recv -> unpack-add kernel", "pull": "owner kernel reads peer rows over NVLink (CUDA IPC), no pack/staging", "none": "none"}[halo],
                           "l2_policy": "working set (>=36 GB CSR per GPU at n=256) exceeds the 126 MB L2",
                           "assemble_kernel": roofline["kernel"], "graph_build_s": t_graph},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
                "clocks": clocks, "wall_ms_per_step": wall_ms, "checks": props, "continuity": continuity,
                "momentum_list": momentum_list}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def multi_rank_selfcheck(rank, world, dev, halo, n=10):
    """Distributed assembly + loadComplete over the real transport against a single-rank assembly of the same
    global box (done redundantly by every rank): owned rows must agree entry by entry, matched by global ids.
    Returns the largest relative difference over all ranks."""
    import scipy.sparse as sp
    import torch
    import torch.distributed as dist
    from nalu_b200 import dist as nd
    from nalu_b200 import fields, mesh
    from nalu_b200.linsys import DeviceCsrLinearSystem

    def assemble(m, exchange, raw):
        mesh.perturb_coords_hash(m, 0.12)
        ls = DeviceCsrLinearSystem(m, 3, exchange=exchange).allocate(dev, raw=raw)
        c = m.coords
        t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
        d = {"coordinates": t(c), "velocity": t(fields.velocity(c)), "viscosity": t(fields.smooth_viscosity(c))}
        mdot = ls.mdot_scs(d["coordinates"], t(fields.smooth_density(c)), d["velocity"])
        ls.assemble_elem([ls.make_desc("momentum_adv_diff", fields=d, mass_flow_rate=mdot, include_divu=1.0)], overwrite=True)
        return ls

    def to_global(ls):
        rp, ci = ls.rowptr().astype(np.int64), ls.colidx()
        cg, rg = ls.col_gids(), ls.row_gids()
        v = ls.owned_values().cpu().numpy()
        ncol = int(3 * (n + 1) * (n + 1) * (n * world + 1)) + 1
        rows = np.repeat(rg, np.diff(rp))
        return sp.csr_matrix((v, (rows, cg[ci])), shape=(ncol, ncol)), ls.owned_rhs().cpu().numpy(), rg

    part = assemble(mesh.generated_box_rank(n, n, n * world, rank, world), nd.make_exchange(rank, world), halo == "pull")
    if halo == "pull":
        w = nd.PeerWindows(part, rank, world)
        w.load_complete(lambda: (torch.cuda.synchronize(), dist.barrier()))
        torch.cuda.synchronize()
        dist.barrier()
        w.close()
    else:
        part.load_complete_nccl()
    torch.cuda.synchronize()
    full = assemble(mesh.generated_box(n, n, n * world), None, False)
    A, b, rg = to_global(part)
    Af, bf, rgf = to_global(full)
    sel = np.zeros(A.shape[0], dtype=bool)
    sel[rg] = True
    D = (A - Af.multiply(sp.csr_matrix(sel[:, None].astype(np.float64)))).tocsr()
    err = abs(D).max() / abs(Af).max() if D.nnz else 0.0
    bfull = np.zeros(A.shape[0])
    bfull[rgf] = bf
    errb = np.abs(b - bfull[rg]).max() / np.abs(bf).max()
    e = torch.tensor([max(float(err), float(errb))], dtype=torch.float64, device=dev)
    dist.all_reduce(e, op=dist.ReduceOp.MAX)
    return float(e.item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=int(os.environ.get("NALU_BENCH_N", "256")),
                    help="elements per box edge per GPU")
    ap.add_argument("--mode", default="auto", choices=["auto", "general", "box"])
    ap.add_argument("--box-kernel", type=int, default=5, choices=[5, 2],
                    help="tiled momentum kernel: 5 row-thread (default), 2 column-pair (round-1 kernel, for A/B runs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-continuity", action="store_true", help="skip the continuity pass reported beside the headline metric")
    ap.add_argument("--e2e-chunks", type=int, default=8, help="node-plane ranges of the pipelined end-to-end step")
    ap.add_argument("--halo", default="nccl", choices=["nccl", "pull"], help="loadComplete transport at N>1")
    ap.add_argument("--divu", type=float, default=1.0, help="SolutionOptions::includeDivU_ (reference default 0; 1 is the costlier variant)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
