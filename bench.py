#!/usr/bin/env python3
"""bench.py -- Hex8 momentum assembly throughput on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # reference algorithm on host cores

One "step" = one AssembleElemSolverAlgorithm::execute of the momentum kernel list of the low-Mach
equation system, [MomentumMassElemKernel(lumped), MomentumAdvDiffElemKernel]
(src/LowMachEquationSystem.C:1304-1398; BASELINE config 3), over all locally-owned Hex8 elements of a
perturbed structured box (LHS + RHS computed AND scattered into the CSR), followed, when N > 1, by
loadComplete (the shared-row halo sum).  --scaling weak (default): every GPU holds an n^3-element
slab (n=256: BASELINE config 3 at N=1; 8 x 256^3 = the 512^3 element count of config 4 at N=8);
--scaling strong: ONE n x n x n box cut into N z-slabs (config 3 "at 1/2/4/8 B200").  Prints ONE JSON
line (rank 0).
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


MASS = dict(dt=0.05, gamma1=1.5, gamma2=-2.0, gamma3=0.5, dens_fac=0.0, om_dens_fac=1.0)


def workload_fields(c):
    """nodal state of the benchmark as analytic functions of the coordinates (host, numpy): any rank -- and the
    sampled-row checker -- can recreate the fields of any node"""
    from nalu_b200 import fields
    vel, rho = fields.velocity(c), fields.smooth_density(c)
    return {"coordinates": c, "velocity": vel, "viscosity": fields.smooth_viscosity(c), "density": rho,
            "density_n": 0.98 * rho, "density_nm1": 0.95 * rho, "velocity_n": 0.9 * vel, "velocity_nm1": 0.8 * vel,
            "dpdx": fields.dpdx(c), "pressure": fields.pressure(c)}


def oracle_kernels(kernels, divu):
    """kernel_fn of tests/sampled_parity.check_sampled_rows: the oracle's descriptors of the benchmark's kernel list"""
    import oracle

    def fn(sub, f):
        keep = []
        mdot = oracle.calc_mass_flow_rate_scs(sub.elem_nodes, f["coordinates"], f["density"], f["velocity"])
        out = []
        for k in kernels:
            if k == "momentum_mass":
                out.append(oracle.make_desc("momentum_mass", keep, coordinates=f["coordinates"], velocity=f["velocity"],
                                            velocityN=f["velocity_n"], velocityNm1=f["velocity_nm1"], density=f["density"],
                                            densityN=f["density_n"], densityNm1=f["density_nm1"], Gpdx=f["dpdx"], lumpedMass=1,
                                            dt=MASS["dt"], gamma1=MASS["gamma1"], gamma2=MASS["gamma2"], gamma3=MASS["gamma3"],
                                            densFac=MASS["dens_fac"], om_densFac=MASS["om_dens_fac"]))
            else:
                out.append(oracle.make_desc("momentum_adv_diff", keep, coordinates=f["coordinates"], velocity=f["velocity"],
                                            viscosity=f["viscosity"], massFlowRate=mdot, includeDivU=divu))
        fn.keep = keep
        return out
    return fn


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
def cpu_reference_run(steps, warmup, kernels, divu, cores=None, nx=64, layers_per_rank=16):
    import oracle
    from nalu_b200 import mesh
    P = cores or min(os.cpu_count() or 1, 64)
    parts = [mesh.generated_box_rank(nx, nx, layers_per_rank * P, r, P) for r in range(P)]
    for m in parts:
        mesh.perturb_coords_hash(m, 0.12)
    ls = oracle.LinSys(parts, 3)  # graph construction is outside the timed region, as for the GPU arm
    fns, descs = [], []
    for m in parts:
        fn = oracle_kernels(kernels, divu)
        descs.append(fn(m, workload_fields(m.coords)))
        fns.append(fn)
    n_elems = sum(m.n_elems for m in parts)

    def one_step():
        ls.zero()
        t0 = time.perf_counter()
        if P == 1:
            ls.assemble_elem(0, descs[0])
        else:
            th = [threading.Thread(target=ls.assemble_elem, args=(r, descs[r])) for r in range(P)]
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
            "sample": f"{nx}x{nx}x{layers_per_rank * P} Hex8 box in {P} z-slabs ({n_elems} elements; the GPU arm's workload is "
                      f"the same kernel list on a larger box: same_config=false), kernels {kernels}, gather + kernels + sumInto "
                      f"+ export-add, {steps} steps; oracle C restatement of the reference, scalar code (no SIMD across "
                      "elements), gcc -O3 -march=x86-64-v3 -ffp-contract=off, one thread per slab",
            "ms_per_step": t * 1e3, "n_elems": n_elems}


def kernel_names(args):
    return ["momentum_mass", "momentum_adv_diff"] if args.kernels == "list" else ["momentum_adv_diff"]


def kernel_labels(args):
    adv = f"MomentumAdvDiffElemKernel(includeDivU={args.divu:g})"
    return ["MomentumMassElemKernel(lumped)", adv] if args.kernels == "list" else [adv]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_run(max(1, args.steps), max(1, min(args.warmup, 2)), kernel_names(args), args.divu)
    line = {"metric": METRIC, "value": r["value"], "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "hex8_box_momentum_list_assembly" if args.kernels == "list" else "hex8_box_momentum_advdiff_assembly",
                       "kernels": kernel_labels(args), "sample": r["sample"], "same_config": False},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def bind_to_gpu_numa_node(index):
    """CPU affinity of this rank on the NUMA node its GPU hangs off, set before any pinned host buffer exists, so that the
    pinned buffers are first-touched in memory local to the GPU's PCIe root: N ranks streaming host->device copies no
    longer share one socket's memory controllers and the inter-socket link (the e2e collapse of the previous round)."""
    try:
        import torch
        p = torch.cuda.get_device_properties(index)
        bdf = None
        if all(hasattr(p, a) for a in ("pci_domain_id", "pci_bus_id", "pci_device_id")):
            bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        else:
            import pynvml
            pynvml.nvmlInit()
            for i in range(pynvml.nvmlDeviceGetCount()):
                hnd = pynvml.nvmlDeviceGetHandleByIndex(i)
                uuid = pynvml.nvmlDeviceGetUUID(hnd)
                uuid = uuid.decode() if isinstance(uuid, bytes) else uuid
                if str(getattr(p, "uuid", "")) in uuid:
                    b = pynvml.nvmlDeviceGetPciInfo(hnd).busId
                    b = b.decode() if isinstance(b, bytes) else b
                    bdf = b.lower()[-12:]
        if bdf is None:
            return {"bound": False, "why": "no PCI address"}
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read())
        if node < 0:
            return {"bound": False, "why": "numa_node = -1 (no NUMA information in this container)", "pci": bdf}
        cpus = set()
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return {"bound": False, "why": "no allowed CPU on the GPU's node", "numa_node": node, "pci": bdf}
        os.sched_setaffinity(0, allowed)
        return {"bound": True, "numa_node": node, "cpus": len(allowed), "pci": bdf}
    except Exception as e:  # affinity is an optimisation, never a requirement
        return {"bound": False, "why": f"{type(e).__name__}: {e}"}


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
    numa = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout at NCCL_DEBUG=VERSION; stdout carries exactly one JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    n = args.n
    strong = args.scaling == "strong"
    nzg = n if strong else n * world   # element layers of the global box
    L = lib.lib()
    lib.tune(box_kernel=args.box_kernel)
    if world > 1:
        # torchrun exports OMP_NUM_THREADS=1: give the host graph builder this rank's share of the cores back
        lib.tune(host_threads=max(1, len(os.sched_getaffinity(0)) // int(os.environ.get("LOCAL_WORLD_SIZE", world))))

    # ---- mesh, graph (once per mesh; outside the timed region) ------------------------------
    t0 = time.perf_counter()
    m = mesh.generated_box_rank(n, n, nzg, rank, world)
    mesh.perturb_coords_hash(m, 0.12)

    from nalu_b200 import dist as nd
    halo = args.halo if world > 1 else "none"
    ls = DeviceCsrLinearSystem(m, 3, exchange=nd.make_exchange(rank, world) if world > 1 else None)
    ls.allocate(dev, raw=(halo == "pull"))
    windows = None
    comm = nd.make_comm(rank, world) if halo == "capi" else None
    halo_stream = torch.cuda.Stream(device=dev) if halo == "capi" else None
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
    h = workload_fields(m.coords)
    d = {k: torch.as_tensor(np.ascontiguousarray(v), device=dev) for k, v in h.items()}
    mdot = ls.mdot_scs(d["coordinates"], d["density"], d["velocity"])
    adv_fields = {k: d[k] for k in ("coordinates", "velocity", "viscosity")}
    mass_fields = {k: d[k] for k in ("coordinates", "velocity", "velocity_n", "velocity_nm1", "density", "density_n",
                                     "density_nm1", "dpdx")}
    desc_adv = ls.make_desc("momentum_adv_diff", fields=adv_fields, mass_flow_rate=mdot, include_divu=args.divu)
    desc_mass = ls.make_desc("momentum_mass", fields=mass_fields, options=("lumped_mass",), **MASS)
    descs = [desc_mass, desc_adv] if args.kernels == "list" else [desc_adv]
    nodal_comps = 19 if args.kernels == "list" else 7   # doubles read per node: x3 u3 mu (+ u_n3 u_nm13 rho x3 dpdx3)

    mode = args.mode
    probe = ls.assemble_elem(descs, mode=mode, overwrite=True)  # also a first warm-up
    overwrite = probe == 2
    kernel_name = (f"assemble_box{args.box_kernel}" if probe == 2 else
                   ("assemble_box + scv_lumped_mass (two launches)" if probe == 3 else "assemble_general"))

    NZl = m.box.nz + 1
    overlap = halo == "capi" and overwrite and NZl >= 4

    def assemble(ev0=None, ev1=None):
        """the assembly launches of one step; with the library's loadComplete the interface node planes (my shared plane
        0, owned by the rank below; my top plane, which receives the rank above's rows) are assembled first and the
        exchange runs on a second stream while the interior is assembled"""
        if not overwrite:
            ls.zero_system()   # kernels that accumulate: zeroSystem is part of their step
        if ev0 is not None:
            ev0.record()
        if not overlap:
            ls.assemble_elem(descs, mode=mode, overwrite=overwrite)
            if ev1 is not None:
                ev1.record()
            return None
        lo, hi = (1 if rank > 0 else 0), (NZl - 1 if rank < world - 1 else NZl)
        if lo > 0:
            ls.assemble_elem_zrange(descs, 0, lo, overwrite=True)
        if hi < NZl:
            ls.assemble_elem_zrange(descs, hi, NZl, overwrite=True)
        cur = torch.cuda.current_stream(dev)
        e1 = torch.cuda.Event()
        e1.record(cur)
        halo_stream.wait_event(e1)
        ls.load_complete(comm, stream=halo_stream.cuda_stream)
        e2 = torch.cuda.Event()
        e2.record(halo_stream)
        ls.assemble_elem_zrange(descs, lo, hi, overwrite=True)
        if ev1 is not None:
            ev1.record()
        return e2

    def step(ev0=None, ev1=None):
        e2 = assemble(ev0, ev1)
        if e2 is not None:
            torch.cuda.current_stream(dev).wait_event(e2)
        elif world > 1:
            load_complete()

    def load_complete():
        if comm is not None:
            ls.load_complete(comm)
        elif windows is not None:
            # owners read the sharers' rows straight across NVLink; the barriers inside order "my neighbour finished
            # assembling" before "I pull its rows" and keep it from re-assembling under my reads
            windows.load_complete(lambda: (torch.cuda.synchronize(), dist.barrier()))
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
        step(asm_ev[i][0], asm_ev[i][1])
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
    n_elems_total = n * n * nzg
    value = n_elems_total / (step_ms * 1e-3) / 1e6

    # ---- roofline of the dominant kernel (the assembly kernel), per launch ---------------------
    peak, peak_src = load_peaks()
    nnz = ls.nnz_owned + ls.nnz_shared
    bytes_alg = algorithmic_bytes(m.n_nodes, n_elems_rank, nnz, 3, nodal_comps=nodal_comps, elem_comps=12)
    achieved = bytes_alg / (asm_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "traffic_source": None, "peak_source": peak_src, "kernel": kernel_name,
                "algorithmic_bytes_per_launch": bytes_alg, "bytes_per_element": bytes_alg / n_elems_rank,
                "kernel_ms": asm_ms, "exposed_halo_ms": step_ms - asm_ms if world > 1 else 0.0}
    workload_name = f"hex8_box_{n}x{n}x{nzg}_{'momentum_list' if args.kernels == 'list' else 'momentum_advdiff'}_assembly"
    tf = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tf):
        try:
            with open(tf) as f:
                t = json.load(f).get(kernel_name)
            if isinstance(t, dict):  # valid for the workload and per-rank element count it was captured on only
                ok = t.get("workload") == workload_name and t.get("elements") == n_elems_rank
                if ok and "fp64" in t:
                    # the co-limiting resource: FP64 work of one launch, counted once with ncu (static), over
                    # the kernel time measured in this run
                    f = t["fp64"]
                    roofline["fp64"] = {"flops_per_launch": f["flops"], "flops_per_element": f["flops"] / n_elems_rank,
                                        "achieved": f["flops"] / (asm_ms * 1e-3) / 1e12, "peak": f["peak_tflops"], "unit": "TFLOP/s",
                                        "frac": f["flops"] / (asm_ms * 1e-3) / 1e12 / f["peak_tflops"],
                                        "pipe_busy_pct_ncu": f["pipe_busy_pct"], "source": "static count: " + f["source"]}
                t = t["bytes"] if ok else None
            if t is not None:
                roofline["traffic"] = t
                roofline["traffic_source"] = "static: dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel (profiles/traffic.json), not measured in this run"
        except Exception:
            pass

    # ---- correctness of the timed result (outside the timed region) -----------------------------------------------
    props = {"finite": bool(torch.isfinite(ls.owned_values()).all()) and bool(torch.isfinite(ls.owned_rhs()).all()),
             "geometry_errors": ls.geometry_errors()}
    if not args.no_oracle_check:
        # sampled rows of THIS system (after loadComplete on several ranks) against the oracle: sub-boxes at the corners,
        # edges, faces, interior and across the rank interfaces of the timed mesh (tests/sampled_parity.py)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import sampled_parity as sp
        res = sp.check_sampled_rows(ls, m, nzg, workload_fields, oracle_kernels(kernel_names(args), args.divu), ndof=3, size=5,
                                    n_random=16)
        agg = torch.tensor([res["max_rel_err_values"], res["max_rel_err_rhs"], float(res["rows"])], dtype=torch.float64, device=dev)
        if world > 1:
            mx = agg.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(agg, op=dist.ReduceOp.SUM)
            res["max_rel_err_values"], res["max_rel_err_rhs"], res["rows"] = float(mx[0]), float(mx[1]), int(agg[2])
        res["tolerance"] = 1e-12
        res["passed"] = bool(res["max_rel_err_values"] < 1e-12 and res["max_rel_err_rhs"] < 1e-12)
        props["sampled_rows_vs_oracle"] = res

    # ---- e2e: host buffers in, residual (RHS) out, every step ------------------------------------------------------
    # every nodal field the step reads comes from the host; the element field mass_flow_rate_scs is produced on the device
    # from them (ComputeMdotElemAlgorithm::execute, what a time step runs between the continuity solve and this assembly)
    e2e_keys = ["coordinates", "velocity", "viscosity", "density", "pressure", "dpdx"] + \
               (["velocity_n", "velocity_nm1", "density_n", "density_nm1"] if args.kernels == "list" else [])
    pinned = {k: torch.as_tensor(np.ascontiguousarray(h[k])).pin_memory() for k in e2e_keys}
    copy_keys = list(e2e_keys)  # the fields a step copies; the closures below read this list
    rhs_host = torch.empty(ls.n_owned_rows, dtype=torch.float64).pin_memory()
    h2d = sum(t.numel() * 8 for t in pinned.values())
    d2h = rhs_host.numel() * 8
    e2e_steps = max(1, min(args.steps, 5))
    tau = MASS["dt"] / MASS["gamma1"]

    def produce_mdot(e_range=None):
        ls.compute_mdot_elem(d["coordinates"], d["density"], d["velocity"], d["pressure"], d["dpdx"], tau, out=mdot, e_range=e_range)

    def e2e_step():
        for k in copy_keys:
            d[k].copy_(pinned[k], non_blocking=True)
        produce_mdot()
        step()
        rhs_host.copy_(ls.owned_rhs(), non_blocking=True)
        torch.cuda.synchronize()

    def matrix_fingerprint():  # outside the timed region: serial and pipelined schedules must agree bit for bit
        v = ls.owned_values()
        return torch.stack([v.sum(), torch.linalg.vector_norm(v, 1), ls.owned_rhs().sum()]).cpu()

    # The same step with the host->device copies of the next node planes overlapped with the assembly of
    # the current ones (nalu_b200_assemble_elem_zrange on the compute stream, copies on a second stream).
    lz = m.box.nz
    NZl = lz + 1
    nxy_nodes, nxy_elems = (n + 1) * (n + 1), n * n
    nchunk = max(1, min(args.e2e_chunks, lz // 16))
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
            l_hi = min(z1, lz)             # element layers z0-1 .. z1-1
            with torch.cuda.stream(copy_stream):
                if p_hi > planes_done:
                    a, b = planes_done * nxy_nodes, p_hi * nxy_nodes
                    for k in copy_keys:
                        d[k][a:b].copy_(pinned[k][a:b], non_blocking=True)
                    planes_done = p_hi
                evc = torch.cuda.Event()
                evc.record(copy_stream)
            cur.wait_event(evc)
            if l_hi > layers_done:
                produce_mdot((layers_done * nxy_elems, l_hi * nxy_elems))
                layers_done = l_hi
            ls.assemble_elem_zrange(descs, z0, z1, overwrite=True)
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
           "note": "every nodal field the step reads (coordinates, velocity states, viscosity, density states, pressure, dpdx) "
                   "comes from pinned host memory every step; mass_flow_rate_scs is produced on the device from them "
                   "(nalu_b200_compute_mdot_elem_range); the RHS (nonlinear residual) is read back to the host; the CSR "
                   "values stay device-resident for a device solver"}

    # Beside it (never the headline): the same step inside a time step's nonlinear loop, where the host marks only the
    # fields an iteration changes as modified (velocity, viscosity, density, pressure, dpdx at NP1); coordinates and the
    # old states N, NM1 were synced by the first iteration and stay device resident -- the modify_on_host /
    # sync_to_device contract of the host mirror's fields (nalu_b200/host/include/stk_shim.h), as in STK's NGP fields.
    if use_pipe:
        copy_keys[:] = ["velocity", "viscosity", "density", "pressure", "dpdx"]
        it_ms = time_e2e(e2e_step_pipelined)
        assert torch.equal(fp_serial, matrix_fingerprint()), "iteration-only copies changed the assembled system"
        e2e["nonlinear_iteration"] = {"value": n_elems_total / (it_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": it_ms,
                                      "h2d_bytes_per_step": sum(pinned[k].numel() * 8 for k in copy_keys),
                                      "note": "fields unchanged since the previous assembly of the time step (coordinates, "
                                              "old states) are not copied again; not the headline e2e"}
        copy_keys[:] = list(e2e_keys)

    # ---- what explains the headline: the adv-diff kernel alone, the continuity pass of BASELINE config 3 (one rank,
    # device-resident inputs, outside the timed region of the headline metric)
    def time_launches(fn, reps=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for a, b in evs:
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        return float(np.mean([a.elapsed_time(b) for a, b in evs]))

    if world == 1 and not args.no_continuity:
        if args.kernels == "list":
            fam1 = ls.assemble_elem([desc_adv], mode=mode, overwrite=True)
            s_ms = time_launches(lambda: ls.assemble_elem([desc_adv], mode=mode, overwrite=(fam1 == 2)))
            s_bytes = algorithmic_bytes(m.n_nodes, n_elems_rank, nnz, 3, nodal_comps=7, elem_comps=12)
            roofline["single_kernel"] = {"kernels": [f"MomentumAdvDiffElemKernel(includeDivU={args.divu:g})"], "value": n_elems_rank / s_ms / 1e3,
                                         "unit": UNIT, "ms": s_ms, "bytes_per_element": s_bytes / n_elems_rank,
                                         "frac": s_bytes / s_ms / 1e6 / peak, "list_over_single": s_ms / asm_ms}
            if not args.no_oracle_check and fam1 == 2:
                props["row_sum_identity_advdiff"] = sp.row_sum_identity(ls, m, mdot)
        ls1 = DeviceCsrLinearSystem(m, 1).allocate(dev)
        f1 = {k: d[k] for k in ("coordinates", "velocity", "density", "pressure", "dpdx")}
        desc1 = ls1.make_desc("continuity_adv", fields=f1, dt=0.05, gamma1=1.5)
        cmode = mode if mode != "box" else "auto"
        fam = ls1.assemble_elem([desc1], mode=cmode, overwrite=True)

        def cont():
            if fam != 2:
                ls1.zero_system()
            ls1.assemble_elem([desc1], mode=cmode, overwrite=(fam == 2))
        c_ms = time_launches(cont)
        c_bytes = algorithmic_bytes(m.n_nodes, n_elems_rank, ls1.nnz_owned, 1, nodal_comps=11, elem_comps=0)
        roofline["continuity"] = {"kernel": "assemble_cont_box" if fam == 2 else "assemble_general", "value": n_elems_rank / c_ms / 1e3,
                                  "unit": UNIT, "ms": c_ms, "bytes_per_element": c_bytes / n_elems_rank,
                                  "frac": c_bytes / c_ms / 1e6 / peak, "bound": "fp64 / instruction issue",
                                  "finite": bool(torch.isfinite(ls1.values).all()), "geometry_errors": ls1.geometry_errors()}
        roofline["momentum_plus_continuity"] = {"value": n_elems_rank / (asm_ms + c_ms) / 1e3, "unit": UNIT,
                                                "frac": (bytes_alg + c_bytes) / (asm_ms + c_ms) / 1e6 / peak,
                                                "note": "two passes (momentum list, then continuity), BASELINE config 3; target 0.60"}
        del ls1
        ls.assemble_elem(descs, mode=mode, overwrite=overwrite)  # leave the headline system in place

    other = {}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # BASELINE config 5 (48^3 Hex27 box), a parity-and-stress case beside the headline: tools/bench_hex27.py, 2-3 s
        try:
            import importlib.util
            spec = importlib.util.spec_from_file_location("bench_hex27", os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools", "bench_hex27.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            h = mod.run(48, 5, 3)
            other["config5_hex27_48cubed"] = {"value": h["value"], "unit": h["unit"], "ms_per_step": h["ms_per_step"],
                                              "roofline_frac": h["roofline"]["frac"], "checks": h["checks"],
                                              "kernel_family": h["kernel_family"], "workload": h["config"]["workload"]}
        except Exception as ex:  # never lose the headline line to a side case
            other["config5_hex27_48cubed"] = {"error": repr(ex)}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(2, 1, kernel_names(args), args.divu)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True, "scaling": args.scaling,
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name,
                           "elements_per_gpu": n_elems_rank, "kernels": kernel_labels(args),
                           "partition": f"z-slabs x{world}",
                           "halo": {"capi": "nalu_b200_load_complete: pack kernel -> ncclSend/ncclRecv -> unpack-add kernel on a second stream" + (", interface planes assembled first, exchange overlapped with the interior assembly" if overlap else ""),
                                    "nccl": "pack kernel -> torch.distributed send/recv -> unpack-add kernel",
                                    "pull": "owner kernel reads peer rows over NVLink (CUDA IPC), no pack/staging", "none": "none"}[halo],
                           "l2_policy": "working set (>=36 GB CSR per GPU at n=256) exceeds the 126 MB L2",
                           "assemble_kernel": kernel_name, "graph_build_s": t_graph, "numa_binding": numa},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
                "clocks": clocks, "wall_ms_per_step": wall_ms, "checks": props, "other_configs": other}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


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
    ap.add_argument("--kernels", default="list", choices=["list", "advdiff"],
                    help="list: [MomentumMass(lumped), MomentumAdvDiff], the momentum list of the low-Mach system (default); advdiff: MomentumAdvDiff alone")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every GPU holds an n^3 slab; strong: one n^3 box cut into --gpus z-slabs")
    ap.add_argument("--no-oracle-check", action="store_true", help="skip the sampled-row comparison with the oracle after the timed region")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-continuity", action="store_true", help="skip the continuity pass reported beside the headline metric")
    ap.add_argument("--e2e-chunks", type=int, default=8, help="node-plane ranges of the pipelined end-to-end step")
    ap.add_argument("--halo", default="capi", choices=["capi", "nccl", "pull"],
                    help="loadComplete transport at N>1: capi = nalu_b200_load_complete (library-owned NCCL communicator, overlapped)")
    ap.add_argument("--divu", type=float, default=1.0, help="SolutionOptions::includeDivU_ (reference default 0; 1 is the costlier variant)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
